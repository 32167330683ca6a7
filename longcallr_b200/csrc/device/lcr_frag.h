/* lcr_frag.h — fragment-matrix and phasing stage: shared argument blocks and launchers (internal). */
#ifndef LCR_FRAG_H
#define LCR_FRAG_H

#include "lcr_device.h"

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
    /* per slot */
    uint32_t *frag_flag, *elem_count; /* count pass outputs (u32 so they can be scanned in place) */
    const uint32_t *frag_scan, *elem_scan;
    /* per candidate (global index) */
    uint32_t *cover_count;
    const uint32_t *cover_off;
    uint32_t *cover_cursor;
    uint32_t *cover_frag; /* fragment index within the region */
    int8_t *cover_cell;
    /* per fragment (global index) */
    uint32_t n_frag_total, n_elem_total;
    uint32_t *frag_slot, *frag_elem_off, *frag_links;
    /* per element */
    uint32_t *elem_snp;
    int8_t *elem_cell;
    uint8_t *elem_base;
    uint8_t *is_fragment;
};

/* everything the phasing kernel reads and writes */
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
    /* state, per fragment (global index) */
    int8_t *tag, *best_tag;
    uint8_t *fp, *assign;
    /* outputs per read */
    int8_t *hp;
    uint32_t *ps;
    /* winners of the warp-per-configuration enumeration search (phase_enum.cu), per (region, chunk) */
    const uint8_t *big_region; /* [n_regions] 1: handled by the cooperative whole-GPU kernel, or null */
    const uint32_t *es_base; /* [n_regions+1] or null */
    const long long *es_prob;
    const uint32_t *es_cfg;
};

void lcr_launch_frag_count(const FragArgs &a, bool long_cigars, cudaStream_t st);
void lcr_launch_region_frag_ranges(uint32_t n_regions, const uint32_t *slot_off, const uint32_t *frag_scan, LcrRegionState *rstate, cudaStream_t st);
void lcr_launch_frag_fill(const FragArgs &a, bool long_cigars, cudaStream_t st);
void lcr_launch_pair_count(const FragArgs &a, LcrPairEntry *table, cudaStream_t st);
void lcr_launch_ld_edges(bool fill, uint32_t thr, uint32_t n_regions, const LcrRegionState *rstate, const LcrPairEntry *table, uint64_t table_size,
                         const uint32_t *entry_region, uint32_t *deg, const uint32_t *adj_off, uint32_t *adj_cursor, uint32_t *adj, cudaStream_t st);
void lcr_launch_adj_sort(uint32_t n_cand, const uint32_t *adj_off, uint32_t *adj, cudaStream_t st);
void lcr_launch_fill_entry_region(uint32_t n_regions, const LcrRegionState *rstate, uint32_t *entry_region, cudaStream_t st);
void lcr_launch_phase(const PhaseArgs &a, cudaStream_t st);
int lcr_launch_phase_grid(const PhaseArgs &a, uint32_t reg, void *bcast_scratch, int sm_count, cudaStream_t st);
size_t lcr_phase_bcast_bytes();
int lcr_enum_shape_for(uint32_t n_cand);
uint32_t lcr_enum_cfgs_per_cta(int shape);
int lcr_launch_enum_search(int shape, bool pre, const PhaseArgs &a, uint32_t n_work, const uint32_t *work_region, const uint32_t *work_chunk, uint32_t nf_cap,
                           long long *out_prob, uint32_t *out_cfg, cudaStream_t st);

#endif
