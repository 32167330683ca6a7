/* lcr_frag.h — fragment-matrix and phasing stage: shared argument blocks and launchers (internal). */
#ifndef LCR_FRAG_H
#define LCR_FRAG_H

#include "lcr_device.h"
#include "lcr_pipeline.h"

struct LcrPairEntry { /* open-addressing table of co-observed SNP pairs of one region */
    unsigned long long key; /* (i << 32) | j, i < j; ~0 = empty */
    uint32_t cis;           /* reads carrying ref/ref or alt/alt        (LD_Pair counts AB + ab) */
    uint32_t trans;         /* reads carrying ref/alt or alt/ref        (Ab + aB)                */
};

struct FragArgs {
    lcr_params P;
    uint32_t n_slots;
    const lcr_region *regions;
    const uint32_t *slot_off, *slot_region;
    const uint8_t *slot_flags;
    const int32_t *pos;
    const uint64_t *seq_off, *cig_off;
    const uint8_t *seq, *qual;
    const uint32_t *cigar;
    LcrRegionState *rstate;
    const lcr_candidate *cand;
    lcr_stats *stats;
    struct LcrCounters *ctr;
    uint32_t elem_cap, pair_cap_total, adj_cap; /* capacities of the element arrays, the pair table and the adjacency */
    /* per slot */
    uint32_t *frag_flag, *elem_count; /* count pass outputs (u32 so they can be scanned in place) */
    const uint32_t *frag_scan, *elem_scan;
    /* per candidate (global index) */
    uint32_t *cover_count;
    const uint32_t *cover_off;
    uint32_t *cover_cursor;
    uint32_t *cover_frag; /* fragment index within the region */
    int8_t *cover_cell;
    /* per fragment (global index); the totals are frag_scan[n_slots] and elem_scan[n_slots], known on the device only */
    uint32_t *frag_slot, *frag_elem_off, *frag_links;
    /* per element */
    uint32_t *elem_snp;
    int8_t *elem_cell;
    uint8_t *elem_base;
    uint8_t *is_fragment;
};

/* everything the phasing kernel reads and writes */
#define LCR_COL_PIECE 128u

struct PhaseArgs {
    lcr_params P;
    uint32_t n_regions;
    const lcr_region *regions;
    const uint32_t *slot_off;
    LcrRegionState *rstate;
    lcr_candidate *cand;
    const LcrDeviceTables *tables;
    lcr_stats *stats;
    /* fragment matrix */
    const uint32_t *frag_slot, *frag_elem_off, *frag_links;
    const uint32_t *elem_snp;
    const int8_t *elem_cell;
    const uint32_t *cover_off, *cover_frag;
    const int8_t *cover_cell;
    /* LD graph (regions with more than max_enum_snps candidates) */
    const uint32_t *adj_off, *adj;
    /* state, per candidate (global index) */
    char4 *st; /* delta, eta, for_phasing at the start of phase(), conserved */
    int8_t *best_hap, *best_gen;
    uint32_t *label, *rank, *work; /* work: adj_total + n_cand entries per region segment */
    long long *blk_q, *blk_qflip;
    /* cooperative kernel: the delta / eta sweep cuts every column into pieces of LCR_COL_PIECE covering reads so that a deep site does
       not hold the whole grid at the barrier; piece_col[p] = site of piece p, piece_off[i] = first piece of site i, col_acc = 5 sums per site */
    uint32_t *piece_off, *piece_col;
    long long *col_acc;
    /* state, per fragment (global index) */
    int8_t *tag, *best_tag;
    uint8_t *fp, *assign;
    uint8_t *ds; /* --downsample: 1 = fragment inside the sampled set of its region (k_downsample); null when the flag is off */
    /* outputs per read: (region << 2 | HP) and (region << 32 | PS) under atomicMin, so that a read shared by several regions
       keeps the entry of the lowest region whatever the CTA order; unpacked by k_finalize_reads */
    uint32_t *hp_key;
    unsigned long long *ps_key;
    /* the warp-per-configuration enumeration search (phase_enum.cu): work lists and winners per (region, chunk) */
    struct LcrCounters *ctr;
    uint32_t big_frag_threshold; /* LD-path regions with at least this many fragments run on the whole GPU (k_phase_grid) */
    uint32_t *es_base;       /* [n_regions+1] first result slot of every region */
    long long *es_prob;
    uint32_t *es_cfg;
    uint32_t *es_done;       /* [n_regions] chunks finished; 0x80000000: the winner's final state is in best_hap / best_gen / best_tag */
    uint32_t *work_region, *work_chunk;
};

void lcr_launch_frag_count(const FragArgs &a, bool long_cigars, cudaStream_t st);
void lcr_launch_region_frag_ranges(const FragArgs &a, uint32_t n_regions, cudaStream_t st);
void lcr_launch_frag_fill(const FragArgs &a, bool long_cigars, cudaStream_t st);
void lcr_launch_pair_plan(const FragArgs &a, uint32_t n_regions, uint32_t *region_cap, const uint32_t *region_cap_off, bool assign, cudaStream_t st);
void lcr_launch_pair_build(const FragArgs &a, uint32_t n_regions, LcrPairEntry *table, uint32_t *entry_region, int sm_count, cudaStream_t st);
void lcr_launch_ld_edges(bool fill, const FragArgs &a, const LcrPairEntry *table, const uint32_t *entry_region, uint32_t *deg, const uint32_t *adj_off, uint32_t *adj_cursor,
                         uint32_t *adj, int sm_count, cudaStream_t st);
void lcr_launch_adj_finish(const FragArgs &a, const uint32_t *adj_off, uint32_t *adj, bool sort, int sm_count, cudaStream_t st);
/* which: 0 every region, 1 the regions outside the enumeration search's plan (es_base), 2 the regions inside it */
void lcr_launch_phase(const PhaseArgs &a, int which, cudaStream_t st);
/* marks PhaseArgs.ds for the regions that downsample (scratch: one uint32 per fragment, for regions too large for shared memory) */
void lcr_launch_downsample(const PhaseArgs &a, uint32_t *scratch, cudaStream_t st);
int lcr_launch_phase_grid(const PhaseArgs &a, const uint32_t *big_list, uint32_t n_big_list, void *bcast_scratch, int sm_count, cudaStream_t st);
size_t lcr_phase_bcast_bytes();
void lcr_launch_enum_plan(const PhaseArgs &a, uint32_t work_cap, int sm_count, cudaStream_t st);
int lcr_launch_enum_search(int bin, const PhaseArgs &a, int sm_count, cudaStream_t st);
/* false when no region of the batch can fall into the bin (its fragment class needs more reads than any region has) */
bool lcr_enum_bin_possible(int bin, uint32_t max_region_reads);

#endif
