/*
 * lcr_device.h — internal layout of the device pipeline (sm_100a).
 *
 * One lcr_submit / lcr_run_device call = the per-region worker body of
 * src/thread.rs:78-221 for every region of the batch:
 *
 *   slot prep        read filter + fetch window (util.rs:636-668), tile work items and segments (read-end trim / poly-A mask applied)
 *   tile pileup      util.rs:650-948 fused with the per-site filter cascade and
 *                    genotype likelihood of candidate.rs:75-463
 *   cand finalize    position sort + dense-cluster filters (candidate.rs:465-526)
 *   fragment build   fragment.rs:28-308 (CSR by read, CSC by SNP, LD pair table)
 *   phase            phase.rs:1087-1296 + snpfrags.rs:191-733, one CTA per region
 *
 * Vocabulary: a *slot* is one (region, read) pair; a *tile* is TILE consecutive
 * positions of one region; an *item* is the part of one read that falls in one tile (a row of the
 * tile's pileup); a *segment* is a run of unmasked aligned bases / deleted / intron positions of an item.
 */
#ifndef LCR_DEVICE_H
#define LCR_DEVICE_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "longcallr_b200.h"
#include "lcr_contract.h"
#include "lcr_pipeline.h"

#define LCR_TILE 512        /* positions per pileup tile == threads per pileup CTA */

/* per-region scalars produced on the device */
struct LcrRegionState {
    int32_t status;
    uint32_t n_cand;
    uint32_t cand_begin;    /* first candidate in the sorted candidate array        */
    uint32_t n_frag;
    uint32_t frag_begin;
    uint32_t n_ld_pairs_cap; /* upper bound of LD pair instances (sum k(k-1)/2)     */
    uint32_t pair_begin;    /* first slot of this region in the pair hash table     */
    uint32_t pair_cap;
};

struct LcrDeviceTables {    /* what the kernels need from lcr_luts, in device-friendly form */
    int64_t gl_fx_err[32], gl_fx_ok[32];
    int64_t fx_err[32], fx_ok[32];
    int64_t fx_prior_homref, fx_prior_homvar, fx_prior_het, fx_log10_2;
    double log10_2;
    double gl_prior_log[3];
    float sor_threshold;
    uint32_t binom_reject[32]; /* bit k of [n]: two-tailed binomial(k; n, 0.5) < 0.05 (candidate.rs:37-47, n <= 30) */
};

struct lcr_ctx {
    lcr_params P;
    int device;
    cudaStream_t stream;
    cudaStream_t side[4];      /* independent launches of one stage fork onto these and join back */
    cudaStream_t copy_stream;  /* lcr_submit: the next chunk's host-to-device copies run here while the current chunk computes */
    lcr_timing last_submit;    /* accounting of the last lcr_submit (lcr_last_submit_timing) */
    std::recursive_mutex mu;   /* entry points serialise on the context: concurrent callers (rayon workers) are safe, not parallel */
    cudaEvent_t ev_fork, ev_join[4];
    lcr_luts luts;
    LcrDeviceTables *d_tables; /* device copy */
    std::vector<uint8_t *> d_ref;
    std::vector<uint64_t> ref_len;
    uint8_t **d_ref_table;     /* device array of contig pointers */
    uint64_t *d_ref_len;
    int ref_table_cap;
    bool ref_dirty;
    std::string last_error;
    int sticky;
    int sm_count;
    /* one device run at a time per context (the entry points serialise on `mu`): its scratch, counters and timing events */
    LcrArena arena;
    LcrCounters *d_ctr;        /* device counter block of the current run */
    LcrCounters *h_ctr;        /* pinned host copies read once at the end of a run */
    lcr_stats *h_stats;
    cudaEvent_t ev_t[8];       /* tile kernel begin / end, run begin, pileup stage end, fragment build end, run end, enumeration end, phase kernels end */
    /* experiment / test knobs read from the environment when the context is created */
    uint32_t big_frag_threshold; /* LCR_BIG_REGION_FRAGS: fragments from which an LD-path region takes the cooperative kernel */
    int frag_walk_mode;          /* LCR_FRAG_WALK: 0 by ops per read, 1 thread per read, 2 warp per read */
    size_t submit_chunk_bytes;   /* LCR_SUBMIT_CHUNK_MB: seq + qual bytes per chunk of lcr_submit */
    int tile_variant;            /* LCR_TILE_VARIANT: launch shape of the tile pileup kernel (pileup.cu) */
    LcrCaps caps_hint;           /* largest capacities any batch of this context has needed: first guess for the next upload */
    /* what the batches of this context needed per unit of input (x 1024): pre-candidates per position, elements and items per slot,
       segments per CIGAR op; a new upload (every chunk of lcr_submit is one) sizes its first attempt with them instead of overflowing again */
    uint64_t hint_pre_q10, hint_elems_q10, hint_items_q10, hint_segs_q10;
    int debug_sync;              /* LCR_DEBUG_SYNC: synchronise and check after every launch group (bring-up only) */
    /* page-locked staging blocks for the small host-side tables of an upload (pageable sources would make every copy wait for the stream) */
    std::vector<std::pair<char *, size_t>> stage_free;
};

struct lcr_device_batch {
    /* inputs, resident */
    uint32_t n_regions, n_reads, n_slots, n_tiles;
    uint64_t n_pos, n_bases, n_cigar;
    lcr_region *regions;
    int32_t *pos;
    uint16_t *flag;
    uint8_t *mapq;
    int8_t *ts;
    float *de;
    uint64_t *seq_off, *cig_off;
    uint8_t *seq, *qual;
    bool qual_on_host;     /* LCR_FLAG_QUAL_ON_DEMAND: `qual` is the device alias of the caller's page-locked array, not a copy */
    /* --exon-only mask (lcr_batch.exon_off / exon_iv): per region the union of its intervals, sorted, as (start, stop) pairs; null = no mask */
    uint32_t *exon_off;
    uint2 *exon_iv;
    /* -v (lcr_batch.ext_*): imported candidate positions per region; null = candidates are called from the pileup */
    uint32_t *ext_off, *ext_pos;
    uint8_t *ext_gt;
    float *ext_qual;
    uint32_t *cigar;
    /* derived on the host at upload: cheap prefix sums over region lengths / read ranges */
    uint32_t *slot_off;    /* [n_regions+1] */
    uint32_t *slot_region; /* [n_slots]     */
    uint32_t *tile_base;   /* [n_regions+1] */
    uint32_t *tile_region; /* [n_tiles]     */
    uint64_t *pos_off;     /* [n_regions+1] */
    std::vector<uint64_t> h_pos_off;
    std::vector<uint32_t> h_slot_off;
    uint64_t h2d_bytes;
    /* results, resident after lcr_run_device */
    LcrRegionState *rstate;   /* [n_regions] */
    lcr_candidate *cand;      /* [n_cand] sorted by (region, pos) */
    uint32_t n_cand;
    int8_t *hp;               /* [n_reads] */
    uint32_t *ps;             /* [n_reads] */
    uint8_t *is_fragment;     /* [n_reads] */
    lcr_stats *d_stats;
    /* optional debug outputs */
    uint32_t *pl_acgt, *pl_fwd, *pl_d, *pl_n, *pl_ts;
    uint32_t n_frag;
    uint64_t n_elem;
    uint32_t *fr_frag_off, *fr_frag_read, *fr_elem_snp;
    uint64_t *fr_elem_off;
    int8_t *fr_elem_cell;
    uint8_t *fr_elem_base;
    bool ran;
    lcr_timing timing;
    /* capacities of the data-dependent scratch of a run (grown when a run reports an overflow) and of the result buffers */
    LcrCaps caps;
    uint64_t cand_alloc;       /* candidates `cand` can hold */
    uint8_t *slot_flags;       /* [n_slots] bit 0: read passes the filter and overlaps its region's window */
    int32_t *status0;          /* [n_regions] region statuses known at upload */
    uint32_t max_region_slots; /* largest read range of a region */
    uint32_t *big_list;        /* regions with enough reads to need the cooperative phasing kernel */
    uint32_t n_big_list;
    LcrCounters counters;      /* host copy of the last run's counter block */
    /* asynchronous upload (lcr_submit): the small tables are ready at ev_meta, seq / qual at ev_seq (null: synchronous upload) */
    cudaEvent_t ev_meta, ev_seq;
    bool seq_wait_pending; /* the run has not waited for ev_seq yet */
    /* lcr_batch.seq4 of an asynchronous upload: the packed bases wait here until the first run expands them into `seq` (after ev_seq) */
    uint8_t *seq4;
    uint64_t *seq4_off;
    bool seq4_pending;
};

/* rust-htslib CigarStringView::leading_softclips / trailing_softclips (util.rs:682-690, fragment.rs:59): the soft clip at that
   end of the alignment, looking through one hard clip (`5H10S...` has 10 leading soft clips) */
#if defined(__CUDACC__)
__device__ __forceinline__ int64_t lcr_leading_softclips(const uint32_t *cigar, uint64_t c0, uint64_t c1) {
    if (c1 <= c0) return 0;
    const uint32_t o0 = cigar[c0];
    if ((o0 & 0xf) == 4) return (int64_t)(o0 >> 4);
    if ((o0 & 0xf) == 5 && c0 + 1 < c1) {
        const uint32_t o1 = cigar[c0 + 1];
        if ((o1 & 0xf) == 4) return (int64_t)(o1 >> 4);
    }
    return 0;
}
__device__ __forceinline__ int64_t lcr_trailing_softclips(const uint32_t *cigar, uint64_t c0, uint64_t c1) {
    if (c1 <= c0) return 0;
    const uint32_t o0 = cigar[c1 - 1];
    if ((o0 & 0xf) == 4) return (int64_t)(o0 >> 4);
    if ((o0 & 0xf) == 5 && c1 - c0 >= 2) {
        const uint32_t o1 = cigar[c1 - 2];
        if ((o1 & 0xf) == 4) return (int64_t)(o1 >> 4);
    }
    return 0;
}
#endif

/* expands db->seq4 into db->seq on `st` and releases the packed copy (api.cu); called where a run first needs the bases */
int lcr_unpack_seq4(lcr_ctx *ctx, lcr_device_batch *db, cudaStream_t st);

/* bring-up aid: with LCR_DEBUG_SYNC set, wait for the stream after a launch group and attribute a failure to it */
#define LCR_DEBUG_CHECK(ctx, what)                                                      \
    do {                                                                                \
        if ((ctx)->debug_sync) {                                                        \
            cudaError_t e__ = cudaStreamSynchronize((ctx)->stream);                     \
            if (e__ == cudaSuccess) e__ = cudaGetLastError();                           \
            if (e__ != cudaSuccess) {                                                   \
                (ctx)->last_error = std::string(what) + ": " + cudaGetErrorString(e__); \
                (ctx)->sticky = LCR_ERR_CUDA;                                           \
                return (ctx)->sticky;                                                   \
            }                                                                           \
        }                                                                               \
    } while (0)

/* error plumbing */
#define LCR_CUDA_TRY(ctx, expr)                                                         \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            (ctx)->last_error = std::string(#expr) + ": " + cudaGetErrorString(e__);    \
            (ctx)->sticky = (e__ == cudaErrorMemoryAllocation) ? LCR_ERR_OOM : LCR_ERR_CUDA; \
            return (ctx)->sticky;                                                       \
        }                                                                               \
    } while (0)

#endif
