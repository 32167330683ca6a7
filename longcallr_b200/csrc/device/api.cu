/*
 * api.cu — the C ABI of include/longcallr_b200.h: contexts, reference upload, batch upload,
 * the device run (pileup/genotype stage, fragment + phasing stage) and result fetch.
 *
 * The call sequence mirrors the reference worker (src/thread.rs:78-221): lcr_set_reference is
 * `ref_seqs.get(chr)` (:59,:79), lcr_submit is the closure body, and the returned records are
 * what the three queues receive (:204-221).  There is no CPU fallback: without a device every
 * compute entry point returns LCR_ERR_NO_DEVICE.
 *
 * A run is one uninterrupted stream of kernels: every data-dependent size stays in a device counter
 * block (lcr_pipeline.h), scratch comes from a per-context arena sized by capacities, and the host
 * synchronises once, at the end, to read the counters.  A run whose counters report an overflow is
 * repeated with larger capacities.
 */
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <new>

#include "lcr_frag.h"

int lcr_stage_pileup(lcr_ctx *ctx, lcr_device_batch *db, LcrArena &A, LcrCounters *ctr, bool launch);

namespace {

struct ResultBox {
    lcr_result res{};
    std::vector<uint32_t> cand_off;
    std::vector<lcr_candidate> cand;
    std::vector<int32_t> region_status;
    std::vector<int8_t> hp;
    std::vector<uint32_t> ps;
    std::vector<uint8_t> is_fragment;
    std::vector<uint64_t> pos_off;
    std::vector<uint32_t> acgt, fwd, d, n, ts;
    std::vector<uint32_t> frag_off, frag_read, elem_snp;
    std::vector<uint64_t> elem_off;
    std::vector<int8_t> elem_cell;
    std::vector<uint8_t> elem_base;
};

/* debug copy of the fragment matrix, filled by the run when LCR_FLAG_EMIT_FRAGMENTS is set */
struct FragDebug {
    std::vector<uint32_t> frag_slot, frag_elem_off, elem_snp;
    std::vector<int8_t> elem_cell;
    std::vector<uint8_t> elem_base;
};

struct DbExtra {
    FragDebug fragdbg;
    char *stage = nullptr; /* page-locked staging block of this upload (returned to the context at release) */
    size_t stage_cap = 0, stage_off = 0;
    std::vector<lcr_region> h_regions;
    std::vector<int32_t> h_status0;
};

template <class T>
int dalloc(lcr_ctx *ctx, T **p, size_t n) {
    LCR_CUDA_TRY(ctx, cudaMallocAsync((void **)p, sizeof(T) * (n ? n : 1), ctx->stream));
    return 0;
}
#define DALLOC(p, n)                          \
    do {                                      \
        int rc__ = dalloc(ctx, &(p), (n));    \
        if (rc__) return rc__;                \
    } while (0)
#define TRY(expr) LCR_CUDA_TRY(ctx, expr)
#define DFREE(p)                                               \
    do {                                                       \
        if (p) { cudaFreeAsync((void *)(p), ctx->stream); (p) = nullptr; } \
    } while (0)

template <class T>
int h2d(lcr_ctx *ctx, T **dst, const T *src, size_t n, uint64_t *bytes) {
    int rc = dalloc(ctx, dst, n);
    if (rc) return rc;
    if (n) {
        LCR_CUDA_TRY(ctx, cudaMemcpyAsync(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice, ctx->stream));
        if (bytes) *bytes += sizeof(T) * n;
    }
    return 0;
}

template <class T>
int h2d_padded(lcr_ctx *ctx, T **dst, const T *src, size_t n, size_t pad, uint64_t *bytes) {
    int rc = dalloc(ctx, dst, n + pad);
    if (rc) return rc;
    LCR_CUDA_TRY(ctx, cudaMemsetAsync(*dst + n, 0, sizeof(T) * pad, ctx->stream));
    if (n) {
        LCR_CUDA_TRY(ctx, cudaMemcpyAsync(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice, ctx->stream));
        if (bytes) *bytes += sizeof(T) * n;
    }
    return 0;
}

/* page-locked staging: take a block of at least `need` bytes from the context's free list (or allocate one) */
static bool stage_take(lcr_ctx *ctx, DbExtra &x, size_t need) {
    for (size_t i = 0; i < ctx->stage_free.size(); ++i)
        if (ctx->stage_free[i].second >= need) {
            x.stage = ctx->stage_free[i].first; x.stage_cap = ctx->stage_free[i].second; x.stage_off = 0;
            ctx->stage_free.erase(ctx->stage_free.begin() + i);
            return true;
        }
    const size_t cap = need + need / 4 + (1u << 20);
    void *p = nullptr;
    if (cudaMallocHost(&p, cap) != cudaSuccess) { cudaGetLastError(); return false; }
    x.stage = static_cast<char *>(p); x.stage_cap = cap; x.stage_off = 0;
    return true;
}
static void stage_give_back(lcr_ctx *ctx, DbExtra &x) {
    if (!x.stage) return;
    if (ctx->stage_free.size() < 16) ctx->stage_free.emplace_back(x.stage, x.stage_cap);
    else cudaFreeHost(x.stage);
    x.stage = nullptr; x.stage_cap = 0;
}
static bool host_is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}
/* a copy of [src, src + bytes) inside the staging block, or src itself when the block is absent or full */
static const void *stage_copy(DbExtra &x, const void *src, size_t bytes) {
    const size_t a = (x.stage_off + 63) & ~(size_t)63;
    if (!x.stage || !bytes || a + bytes > x.stage_cap) return src;
    memcpy(x.stage + a, src, bytes);
    x.stage_off = a + bytes;
    return x.stage + a;
}

/* lcr_batch.seq4 -> the ASCII letters rust-htslib's Seq::as_bytes() yields (util.rs:693): one warp per read, one packed byte per lane and step */
__global__ void __launch_bounds__(256) k_unpack_seq4(uint32_t n_reads, const uint64_t *seq_off, const uint64_t *seq4_off, const uint8_t *seq4, uint8_t *seq) {
    const unsigned long long LO = 0x565352474d43413dull, HI = 0x4e42444b48595754ull; /* "=ACMGRSV" "TWYHKDBN", first letter in the low byte */
    const uint32_t lane = threadIdx.x & 31, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_reads; r += nwarps) {
        const uint64_t s0 = seq_off[r], l = seq_off[r + 1] - s0;
        const uint8_t *src = seq4 + seq4_off[r];
        uint8_t *dst = seq + s0;
        for (uint64_t k = lane; 2 * k < l; k += 32) {
            const uint32_t b = src[k], hi = b >> 4, lo = b & 15u;
            dst[2 * k] = (uint8_t)((hi < 8 ? LO : HI) >> (8 * (hi & 7)));
            if (2 * k + 1 < l) dst[2 * k + 1] = (uint8_t)((lo < 8 ? LO : HI) >> (8 * (lo & 7)));
        }
    }
}

/* unpack the per-read (region, value) keys: hp -1 / ps 0 = no entry */
__global__ void k_finalize_reads(uint32_t n, const uint32_t *hp_key, const unsigned long long *ps_key, int8_t *hp, uint32_t *ps) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t h = hp_key[i];
    hp[i] = h == 0xffffffffu ? (int8_t)-1 : (int8_t)(h & 3u);
    const unsigned long long p = ps_key[i];
    ps[i] = p == ~0ull ? 0u : (uint32_t)p;
}

/* start of a run: region states from the statuses known at upload */
__global__ void k_init_rstate(uint32_t n_regions, const int32_t *status0, LcrRegionState *rstate) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_regions) return;
    LcrRegionState s;
    memset(&s, 0, sizeof s);
    s.status = status0[r];
    rstate[r] = s;
}

} // namespace

int lcr_unpack_seq4(lcr_ctx *ctx, lcr_device_batch *db, cudaStream_t st) {
    if (!db->seq4_pending) return LCR_OK;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(((uint64_t)db->n_reads + 7) / 8, (uint64_t)(ctx->sm_count > 0 ? ctx->sm_count : 148) * 8);
    if (db->n_reads) k_unpack_seq4<<<grid, 256, 0, st>>>(db->n_reads, db->seq_off, db->seq4_off, db->seq4, db->seq);
    LCR_CUDA_TRY(ctx, cudaGetLastError());
    cudaFreeAsync(db->seq4, st); db->seq4 = nullptr;
    cudaFreeAsync(db->seq4_off, st); db->seq4_off = nullptr;
    db->seq4_pending = false;
    return LCR_OK;
}


struct lcr_device_batch_full : lcr_device_batch {
    DbExtra extra;
};

/* fragment matrix, LD graph, phasing, read / SNP assignment, phase sets: carve from the arena, enqueue when `launch` */
static int stage_fragments_phase(lcr_ctx *ctx, lcr_device_batch_full *db, LcrArena &A, LcrCounters *ctr, bool launch) {
    cudaStream_t st = ctx->stream;
    const uint32_t n_slots = db->n_slots, n_regions = db->n_regions, n_reads = db->n_reads;
    const LcrCaps &C = db->caps;
    const size_t sn = (size_t)n_slots + 1, cn = (size_t)std::min<uint64_t>(C.pre, 0xfffffff0u) + 1, rn = (size_t)n_regions + 1;
    /* zeroed block: cover_count | cover_cursor | deg | adj_cursor (each cn), es_done (rn) */
    uint32_t *zero_blk = A.take<uint32_t>(4 * cn + rn);
    uint32_t *frag_flag = A.take<uint32_t>(sn), *elem_count = A.take<uint32_t>(sn), *frag_scan = A.take<uint32_t>(sn), *elem_scan = A.take<uint32_t>(sn);
    uint32_t *cover_off = A.take<uint32_t>(cn), *adj_off = A.take<uint32_t>(cn);
    uint32_t *frag_slot = A.take<uint32_t>(sn), *frag_elem_off = A.take<uint32_t>(sn + 1), *frag_links = A.take<uint32_t>(sn);
    uint32_t *elem_snp = A.take<uint32_t>(C.elems), *cover_frag = A.take<uint32_t>(C.elems);
    int8_t *elem_cell = A.take<int8_t>(C.elems), *cover_cell = A.take<int8_t>(C.elems);
    uint8_t *elem_base = A.take<uint8_t>(C.elems);
    LcrPairEntry *table = A.take<LcrPairEntry>(C.pairs);
    uint32_t *entry_region = A.take<uint32_t>(C.pairs / 16 + 1);
    uint32_t *region_cap = A.take<uint32_t>(rn), *region_cap_off = A.take<uint32_t>(rn);
    uint32_t *adj = A.take<uint32_t>(C.adj);
    PhaseArgs pa{};
    pa.st = A.take<char4>(cn); pa.best_hap = A.take<int8_t>(cn); pa.best_gen = A.take<int8_t>(cn);
    pa.label = A.take<uint32_t>(cn); pa.rank = A.take<uint32_t>(cn);
    pa.work = A.take<uint32_t>(C.adj + cn + 1);
    pa.blk_q = A.take<long long>(cn); pa.blk_qflip = A.take<long long>(cn);
    pa.piece_off = A.take<uint32_t>(cn + 1); pa.piece_col = A.take<uint32_t>(C.elems / LCR_COL_PIECE + cn + 1); pa.col_acc = A.take<long long>(5 * cn);
    pa.tag = A.take<int8_t>(sn); pa.best_tag = A.take<int8_t>(sn); pa.fp = A.take<uint8_t>(sn); pa.assign = A.take<uint8_t>(sn);
    const bool ds_on = (ctx->P.flags & LCR_FLAG_DOWNSAMPLE) && ctx->P.downsample_depth > 0;
    pa.ds = ds_on ? A.take<uint8_t>(sn) : nullptr; /* --downsample: the sampled set, and the shuffle's index scratch for very large regions */
    uint32_t *ds_scratch = ds_on ? A.take<uint32_t>(sn) : nullptr;
    pa.hp_key = A.take<uint32_t>(n_reads); pa.ps_key = A.take<unsigned long long>(n_reads);
    pa.es_base = A.take<uint32_t>(rn); pa.es_cfg = A.take<uint32_t>(C.enum_work); pa.es_prob = A.take<long long>(C.enum_work);
    pa.work_region = A.take<uint32_t>(C.enum_work); pa.work_chunk = A.take<uint32_t>(C.enum_work);
    void *bcast = A.take<char>(lcr_phase_bcast_bytes());
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)std::max(std::max(sn, cn), rn), st);
    void *tmp = A.take<char>(tmp_bytes + 256);
    /* the debug copy of the matrix is taken from these after the run */
    db->fr_frag_off = frag_slot; db->fr_frag_read = frag_elem_off; db->fr_elem_snp = elem_snp; db->fr_elem_cell = elem_cell; db->fr_elem_base = elem_base;
    if (!launch) return LCR_OK;

    uint32_t *cover_count = zero_blk, *cover_cursor = zero_blk + cn, *deg = zero_blk + 2 * cn, *adj_cursor = zero_blk + 3 * cn;
    pa.es_done = zero_blk + 4 * cn;
    TRY(cudaMemsetAsync(zero_blk, 0, sizeof(uint32_t) * (4 * cn + rn), st));
    TRY(cudaMemsetAsync(pa.hp_key, 0xff, sizeof(uint32_t) * (size_t)(n_reads ? n_reads : 1), st));
    TRY(cudaMemsetAsync(pa.ps_key, 0xff, sizeof(unsigned long long) * (size_t)(n_reads ? n_reads : 1), st));
    TRY(cudaMemsetAsync(frag_flag + n_slots, 0, sizeof(uint32_t), st));
    TRY(cudaMemsetAsync(elem_count + n_slots, 0, sizeof(uint32_t), st));

    FragArgs fa{};
    fa.P = ctx->P;
    fa.n_slots = n_slots;
    fa.regions = db->regions;
    fa.slot_off = db->slot_off; fa.slot_region = db->slot_region; fa.slot_flags = db->slot_flags;
    fa.pos = db->pos; fa.seq_off = db->seq_off; fa.cig_off = db->cig_off; fa.seq = db->seq; fa.qual = db->qual; fa.cigar = db->cigar;
    fa.rstate = db->rstate; fa.cand = db->cand; fa.stats = db->d_stats; fa.ctr = ctr;
    fa.elem_cap = (uint32_t)std::min<uint64_t>(C.elems, 0xfffffff0u); fa.pair_cap_total = (uint32_t)std::min<uint64_t>(C.pairs, 0xfffffff0u);
    fa.adj_cap = (uint32_t)std::min<uint64_t>(C.adj, 0xfffffff0u);
    fa.frag_flag = frag_flag; fa.elem_count = elem_count; fa.frag_scan = frag_scan; fa.elem_scan = elem_scan;
    fa.cover_count = cover_count; fa.cover_off = cover_off; fa.cover_cursor = cover_cursor;
    fa.cover_frag = cover_frag; fa.cover_cell = cover_cell;
    fa.frag_slot = frag_slot; fa.frag_elem_off = frag_elem_off; fa.frag_links = frag_links;
    fa.elem_snp = elem_snp; fa.elem_cell = elem_cell; fa.elem_base = elem_base;
    fa.is_fragment = db->is_fragment;
    const int walk_mode = ctx->frag_walk_mode; /* 1: thread per read, 2: warp per read */
    const bool long_cigars = walk_mode == 2 || (walk_mode == 0 && db->n_cigar > 24ull * (db->n_reads ? db->n_reads : 1));
    LCR_DEBUG_CHECK(ctx, "fragment stage memsets");
    lcr_launch_frag_count(fa, long_cigars, st);
    LCR_DEBUG_CHECK(ctx, "frag_count");
    TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, frag_flag, frag_scan, (int)sn, st));
    TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, elem_count, elem_scan, (int)sn, st));
    TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cover_count, cover_off, (int)cn, st));
    LCR_DEBUG_CHECK(ctx, "fragment scans");
    lcr_launch_region_frag_ranges(fa, n_regions, st);
    LCR_DEBUG_CHECK(ctx, "region_frag_ranges");
    lcr_launch_frag_fill(fa, long_cigars, st);
    LCR_DEBUG_CHECK(ctx, "frag_fill");
    db->timing.kernel_launches += 3; /* own kernels only; cub scans are library code */

    /* LD graph of the regions with more than max_enum_snps candidates: pair table segments, perfect-LD edges, sorted adjacency */
    lcr_launch_pair_plan(fa, n_regions, region_cap, region_cap_off, false, st);
    TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, region_cap, region_cap_off, (int)rn, st));
    lcr_launch_pair_plan(fa, n_regions, region_cap, region_cap_off, true, st);
    LCR_DEBUG_CHECK(ctx, "pair_plan");
    lcr_launch_pair_build(fa, n_regions, table, entry_region, ctx->sm_count, st);
    LCR_DEBUG_CHECK(ctx, "pair_build");
    lcr_launch_ld_edges(false, fa, table, entry_region, deg, nullptr, nullptr, nullptr, ctx->sm_count, st);
    TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, deg, adj_off, (int)cn, st));
    lcr_launch_adj_finish(fa, adj_off, adj, false, ctx->sm_count, st);
    lcr_launch_ld_edges(true, fa, table, entry_region, deg, adj_off, adj_cursor, adj, ctx->sm_count, st);
    lcr_launch_adj_finish(fa, adj_off, adj, true, ctx->sm_count, st);
    db->timing.kernel_launches += 9;
    LCR_DEBUG_CHECK(ctx, "ld graph");
    TRY(cudaEventRecord(ctx->ev_t[4], st));

    /* phasing state */
    pa.P = ctx->P;
    pa.n_regions = n_regions;
    pa.regions = db->regions; pa.slot_off = db->slot_off; pa.rstate = db->rstate; pa.cand = db->cand;
    pa.tables = ctx->d_tables; pa.stats = db->d_stats;
    pa.frag_slot = frag_slot; pa.frag_elem_off = frag_elem_off; pa.frag_links = frag_links;
    pa.elem_snp = elem_snp; pa.elem_cell = elem_cell;
    pa.cover_off = cover_off; pa.cover_frag = cover_frag; pa.cover_cell = cover_cell;
    pa.adj_off = adj_off; pa.adj = adj;
    pa.ctr = ctr;
    pa.big_frag_threshold = ctx->big_frag_threshold;
    if (ds_on) { /* thread.rs:144-151: mark the sampled fragments of the regions deep enough to downsample */
        TRY(cudaMemsetAsync(pa.ds, 0, sn, st));
        lcr_launch_downsample(pa, ds_scratch, st);
        db->timing.kernel_launches += 1;
        LCR_DEBUG_CHECK(ctx, "k_downsample");
    }
    /* enumeration search (regions with at most min(max_enum_snps, 10) candidates): work lists on the device, one persistent launch
       per (shape, class) bin; the bins are independent: fork them onto side streams so that small and large shapes overlap */
    lcr_launch_enum_plan(pa, (uint32_t)std::min<uint64_t>(C.enum_work, 0xfffffff0u), ctx->sm_count, st);
    db->timing.kernel_launches += 1;
    LCR_DEBUG_CHECK(ctx, "enum_plan");
    TRY(cudaEventRecord(ctx->ev_fork, st));
    for (int i = 0; i < 4; ++i) TRY(cudaStreamWaitEvent(ctx->side[i], ctx->ev_fork, 0));
    /* the LD-path regions do not wait for the search: their CTAs (long, latency-bound chains of sweeps) run on the main stream beside it */
    lcr_launch_phase(pa, 1, st);
    db->timing.kernel_launches += 1;
    for (int b = 0, nl = 0; b < LCR_ENUM_BINS; ++b) {
        if (!lcr_enum_bin_possible(b, db->max_region_slots)) continue;
        int e = lcr_launch_enum_search(b, pa, ctx->sm_count, ctx->side[nl++ % 4]);
        if (e) { ctx->last_error = std::string("k_enum_search: ") + cudaGetErrorString((cudaError_t)e); ctx->sticky = LCR_ERR_CUDA; return ctx->sticky; }
        db->timing.kernel_launches += 1;
    }
    for (int i = 0; i < 4; ++i) {
        TRY(cudaEventRecord(ctx->ev_join[i], ctx->side[i]));
        TRY(cudaStreamWaitEvent(st, ctx->ev_join[i], 0));
    }
    LCR_DEBUG_CHECK(ctx, "enum_search");
    TRY(cudaEventRecord(ctx->ev_t[6], st));
    lcr_launch_phase(pa, 2, st);
    db->timing.kernel_launches += 1;
    LCR_DEBUG_CHECK(ctx, "k_phase");
    /* regions too large for one CTA take the whole GPU, one after the other */
    if (db->n_big_list) {
        int e = lcr_launch_phase_grid(pa, db->big_list, db->n_big_list, bcast, ctx->sm_count, st);
        if (e) { ctx->last_error = std::string("k_phase_grid: ") + cudaGetErrorString((cudaError_t)e); ctx->sticky = LCR_ERR_CUDA; return ctx->sticky; }
        db->timing.kernel_launches += 1;
    }
    TRY(cudaEventRecord(ctx->ev_t[7], st));
    if (n_reads) {
        k_finalize_reads<<<(n_reads + 255) / 256, 256, 0, st>>>(n_reads, pa.hp_key, pa.ps_key, db->hp, db->ps);
        db->timing.kernel_launches += 1;
    }
    TRY(cudaGetLastError());
    return LCR_OK;
}

/* sizes a run needs from the arena with the batch's current capacities */
static int plan_or_launch(lcr_ctx *ctx, lcr_device_batch_full *db, bool launch) {
    LcrArena &A = ctx->arena;
    A.off = 0;
    A.dry = !launch;
    int rc = lcr_stage_pileup(ctx, db, A, ctx->d_ctr, launch);
    if (rc) return rc;
    if (launch) TRY(cudaEventRecord(ctx->ev_t[3], ctx->stream));
    if (!(ctx->P.flags & LCR_FLAG_SKIP_PHASING)) rc = stage_fragments_phase(ctx, db, A, ctx->d_ctr, launch);
    else if (launch) { TRY(cudaEventRecord(ctx->ev_t[4], ctx->stream)); TRY(cudaEventRecord(ctx->ev_t[6], ctx->stream)); TRY(cudaEventRecord(ctx->ev_t[7], ctx->stream)); }
    return rc;
}

extern "C" {

int lcr_abi_version(void) { return LCR_ABI_VERSION; }

int lcr_pin_host(void *ptr, size_t bytes) {
    if (!ptr || !bytes) return LCR_ERR_INVALID_ARG;
    const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorMemoryAllocation ? LCR_ERR_OOM : LCR_ERR_CUDA; }
    return LCR_OK;
}
int lcr_unpin_host(void *ptr) {
    if (!ptr) return LCR_ERR_INVALID_ARG;
    if (cudaHostUnregister(ptr) != cudaSuccess) { cudaGetLastError(); return LCR_ERR_CUDA; }
    return LCR_OK;
}

const char *lcr_strerror(int status) {
    switch (status) {
        case LCR_OK: return "ok";
        case LCR_ERR_INVALID_ARG: return "invalid argument";
        case LCR_ERR_CUDA: return "CUDA error";
        case LCR_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback on this path)";
        case LCR_ERR_OOM: return "out of memory";
        case LCR_ERR_BAD_CIGAR: return "unknown or inconsistent CIGAR operation";
        case LCR_ERR_NO_REFERENCE: return "region on a contig without reference sequence";
        case LCR_ERR_BASEQ_ZERO: return "base quality 0 at a phase site";
        case LCR_ERR_INTERNAL: return "internal invariant violated";
        default: return "unknown status";
    }
}

const char *lcr_last_error(lcr_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int lcr_create(const lcr_params *p, int device, lcr_ctx **out) {
    if (!p || !out) return LCR_ERR_INVALID_ARG;
    /* `1 << n` configurations are enumerated for up to max_enum_snps candidates (phase.rs:1097-1122) */
    if (p->max_enum_snps > 20u) return LCR_ERR_INVALID_ARG;
    if (p->platform != 0 && p->platform != 1) return LCR_ERR_INVALID_ARG;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return LCR_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return LCR_ERR_NO_DEVICE;
    lcr_ctx *ctx = new (std::nothrow) lcr_ctx();
    if (!ctx) return LCR_ERR_OOM;
    ctx->P = *p;
    ctx->device = device;
    ctx->sticky = 0;
    ctx->d_ref_table = nullptr;
    ctx->d_ref_len = nullptr;
    ctx->ref_table_cap = 0;
    ctx->ref_dirty = true;
    ctx->d_ctr = nullptr; ctx->h_ctr = nullptr; ctx->h_stats = nullptr;
    memset(&ctx->caps_hint, 0, sizeof ctx->caps_hint);
    ctx->hint_pre_q10 = ctx->hint_elems_q10 = ctx->hint_items_q10 = ctx->hint_segs_q10 = 0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    ctx->sm_count = prop.multiProcessorCount;
    /* experiment / test knobs, read once per context (DESIGN.md "Environment knobs") */
    {
        const char *e = getenv("LCR_BIG_REGION_FRAGS");
        ctx->big_frag_threshold = (e && *e) ? (uint32_t)strtoul(e, nullptr, 10) : 8192u;
        e = getenv("LCR_FRAG_WALK");
        ctx->frag_walk_mode = (e && *e) ? atoi(e) : 0;
        e = getenv("LCR_SUBMIT_CHUNK_MB");
        ctx->submit_chunk_bytes = (e && *e) ? (size_t)strtoull(e, nullptr, 10) << 20 : (size_t)512 << 20;
        e = getenv("LCR_TILE_VARIANT");
        ctx->tile_variant = (e && *e) ? atoi(e) : 0;
        e = getenv("LCR_DEBUG_SYNC");
        ctx->debug_sync = (e && *e) ? atoi(e) : 0;
    }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return LCR_ERR_CUDA; }
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    memset(&ctx->last_submit, 0, sizeof ctx->last_submit);
    for (int i = 0; i < 4; ++i) {
        cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming);
    }
    for (int i = 0; i < 8; ++i) cudaEventCreate(&ctx->ev_t[i]);
    /* keep freed blocks in the stream-ordered pool between runs */
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    lcr_build_luts(&ctx->luts);
    LcrDeviceTables t{};
    for (int q = 0; q <= LCR_MAX_BASE_QUALITY; ++q) {
        t.gl_fx_err[q] = ctx->luts.gl_fx_err[q];
        t.gl_fx_ok[q] = ctx->luts.gl_fx_ok[q];
        t.fx_err[q] = ctx->luts.fx_err[q];
        t.fx_ok[q] = ctx->luts.fx_ok[q];
    }
    t.fx_prior_homref = ctx->luts.fx_prior_homref; t.fx_prior_homvar = ctx->luts.fx_prior_homvar;
    t.fx_prior_het = ctx->luts.fx_prior_het; t.fx_log10_2 = ctx->luts.fx_log10_2;
    t.log10_2 = ctx->luts.log10_2;
    for (int i = 0; i < 3; ++i) t.gl_prior_log[i] = ctx->luts.gl_prior_log[i];
    t.sor_threshold = ctx->luts.sor_threshold;
    for (uint32_t n = 0; n <= 30; ++n) {
        t.binom_reject[n] = 0;
        for (uint32_t k = 0; k <= n; ++k)
            if (lcr_binom_two_tailed_lt_0p05(k, n)) t.binom_reject[n] |= 1u << k;
    }
    bool ok = cudaMalloc(&ctx->d_tables, sizeof t) == cudaSuccess && cudaMemcpy(ctx->d_tables, &t, sizeof t, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_ctr, sizeof(LcrCounters)) == cudaSuccess;
    ok = ok && cudaMallocHost(&ctx->h_ctr, sizeof(LcrCounters)) == cudaSuccess && cudaMallocHost(&ctx->h_stats, sizeof(lcr_stats)) == cudaSuccess;
    if (!ok) {
        lcr_destroy(ctx);
        return LCR_ERR_CUDA;
    }
    *out = ctx;
    return LCR_OK;
}

void lcr_destroy(lcr_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (uint8_t *p : ctx->d_ref) if (p) cudaFree(p);
    if (ctx->d_ref_table) cudaFree(ctx->d_ref_table);
    if (ctx->d_ref_len) cudaFree(ctx->d_ref_len);
    if (ctx->d_tables) cudaFree(ctx->d_tables);
    if (ctx->d_ctr) cudaFree(ctx->d_ctr);
    if (ctx->h_ctr) cudaFreeHost(ctx->h_ctr);
    if (ctx->h_stats) cudaFreeHost(ctx->h_stats);
    if (ctx->arena.base) cudaFree(ctx->arena.base);
    for (auto &sb : ctx->stage_free) cudaFreeHost(sb.first);
    ctx->stage_free.clear();
    for (int i = 0; i < 4; ++i) { cudaStreamDestroy(ctx->side[i]); cudaEventDestroy(ctx->ev_join[i]); }
    for (int i = 0; i < 8; ++i) cudaEventDestroy(ctx->ev_t[i]);
    cudaEventDestroy(ctx->ev_fork);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int lcr_set_reference(lcr_ctx *ctx, int32_t tid, const uint8_t *seq, uint64_t len) {
    std::unique_lock<std::recursive_mutex> ctx_lock__;
    if (ctx) ctx_lock__ = std::unique_lock<std::recursive_mutex>(ctx->mu);
    if (!ctx || tid < 0 || (!seq && len)) return LCR_ERR_INVALID_ARG;
    if (ctx->sticky) return ctx->sticky;
    TRY(cudaSetDevice(ctx->device));
    if ((size_t)tid >= ctx->d_ref.size()) { ctx->d_ref.resize(tid + 1, nullptr); ctx->ref_len.resize(tid + 1, 0); }
    if (ctx->d_ref[tid]) { TRY(cudaFree(ctx->d_ref[tid])); ctx->d_ref[tid] = nullptr; }
    TRY(cudaMalloc(&ctx->d_ref[tid], len + 64)); /* slack: the tile kernel copies whole 16-byte groups */
    if (len) TRY(cudaMemcpyAsync(ctx->d_ref[tid], seq, len, cudaMemcpyHostToDevice, ctx->stream));
    TRY(cudaStreamSynchronize(ctx->stream));
    ctx->ref_len[tid] = len;
    ctx->ref_dirty = true;
    return LCR_OK;
}

/* async: allocate and copy on the context's copy stream and record `ready` instead of waiting */
static int upload_impl(lcr_ctx *ctx, const lcr_batch *b, lcr_device_batch **out, bool async) {
    std::unique_lock<std::recursive_mutex> ctx_lock__;
    if (ctx) ctx_lock__ = std::unique_lock<std::recursive_mutex>(ctx->mu);
    if (!ctx || !b || !out) return LCR_ERR_INVALID_ARG;
    struct StreamSwap { /* the helpers above issue work on ctx->stream */
        lcr_ctx *c; cudaStream_t saved; bool on;
        StreamSwap(lcr_ctx *c_, bool on_) : c(c_), saved(c_->stream), on(on_) { if (on) c->stream = c->copy_stream; }
        ~StreamSwap() { if (on) c->stream = saved; }
    } swap_guard(ctx, async);
    if (ctx->sticky) return ctx->sticky;
    if (b->n_regions && !b->regions) return LCR_ERR_INVALID_ARG;
    if (b->exon_off && b->n_regions && b->exon_off[b->n_regions] > b->exon_off[0] && !b->exon_iv) return LCR_ERR_INVALID_ARG;
    if (b->ext_off) { /* imported candidates: offsets non-decreasing, positions ascending inside a region (the reference keeps one record per position) */
        for (uint32_t r = 0; r < b->n_regions; ++r) {
            if (b->ext_off[r + 1] < b->ext_off[r]) return LCR_ERR_INVALID_ARG;
            if (b->ext_off[r + 1] > b->ext_off[r] && (!b->ext_pos || !b->ext_gt || !b->ext_qual)) return LCR_ERR_INVALID_ARG;
            for (uint32_t e = b->ext_off[r] + 1; e < b->ext_off[r + 1]; ++e)
                if (b->ext_pos[e] <= b->ext_pos[e - 1]) return LCR_ERR_INVALID_ARG;
        }
    }
    if (b->n_reads && (!b->pos || !b->flag || !b->mapq || !b->ts || !b->de || !b->seq_off || !b->cig_off)) return LCR_ERR_INVALID_ARG;
    /* the kernels index the pools with these offsets: they must be non-decreasing and the pools present */
    const uint64_t n_bases = b->n_reads ? b->seq_off[b->n_reads] : 0, n_cig = b->n_reads ? b->cig_off[b->n_reads] : 0;
    if ((n_bases && ((!b->seq && !b->seq4) || !b->qual)) || (n_cig && !b->cigar)) return LCR_ERR_INVALID_ARG;
    if (n_bases >= (1ull << 47) || n_cig >= (1ull << 47)) return LCR_ERR_INVALID_ARG;
    for (uint32_t i = 0; i < b->n_reads; ++i)
        if (b->seq_off[i + 1] < b->seq_off[i] || b->cig_off[i + 1] < b->cig_off[i]) return LCR_ERR_INVALID_ARG;
    const bool packed = n_bases && b->seq4;
    if (packed) { /* every read's packed bytes must lie inside its own span of the pool */
        if (!b->seq4_off) return LCR_ERR_INVALID_ARG;
        for (uint32_t i = 0; i < b->n_reads; ++i)
            if (b->seq4_off[i + 1] < b->seq4_off[i] || b->seq4_off[i + 1] - b->seq4_off[i] < (b->seq_off[i + 1] - b->seq_off[i] + 1) / 2) return LCR_ERR_INVALID_ARG;
    }
    TRY(cudaSetDevice(ctx->device));
    lcr_device_batch_full *db = new (std::nothrow) lcr_device_batch_full();
    if (!db) return LCR_ERR_OOM;
    db->ev_meta = nullptr; db->ev_seq = nullptr; db->seq_wait_pending = false;
    db->seq4 = nullptr; db->seq4_off = nullptr; db->seq4_pending = false;
    db->exon_off = nullptr; db->exon_iv = nullptr;
    db->ext_off = nullptr; db->ext_pos = nullptr; db->ext_gt = nullptr; db->ext_qual = nullptr;
    db->n_regions = b->n_regions;
    db->n_reads = b->n_reads;
    db->ran = false;
    memset(&db->timing, 0, sizeof db->timing);
    /* prefix sums over region lengths and read ranges (64-bit: read ranges may overlap, so slots are not bounded by the
       read count); regions that cannot run get their status here */
    std::vector<uint32_t> slot_off(b->n_regions + 1, 0), tile_base(b->n_regions + 1, 0), big_list;
    std::vector<uint64_t> pos_off(b->n_regions + 1, 0);
    db->extra.h_status0.assign(b->n_regions, 0);
    uint64_t slots64 = 0, tiles64 = 0;
    uint32_t max_slots = 0;
    for (uint32_t r = 0; r < b->n_regions; ++r) {
        const lcr_region &g = b->regions[r];
        int32_t stt = 0;
        if (g.end < g.start || g.start < 1 || g.read_end < g.read_begin || g.read_end > b->n_reads) stt = LCR_ERR_INVALID_ARG;
        else if (g.tid < 0 || (size_t)g.tid >= ctx->d_ref.size() || !ctx->d_ref[g.tid]) stt = LCR_ERR_NO_REFERENCE;
        else if ((uint64_t)g.end - 1 > ctx->ref_len[g.tid]) stt = LCR_ERR_INVALID_ARG;
        else if (b->exon_off && b->exon_off[r + 1] < b->exon_off[r]) stt = LCR_ERR_INVALID_ARG;
        else if (b->exon_off && b->exon_off[r + 1] == b->exon_off[r]) stt = LCR_REGION_NO_EXON; /* thread.rs:88-91: the region is skipped */
        db->extra.h_status0[r] = stt;
        const uint64_t len = stt ? 0 : (uint64_t)(g.end - g.start);
        const uint32_t nreads = stt ? 0 : g.read_end - g.read_begin;
        slots64 += nreads;
        tiles64 += (len + LCR_TILE - 1) / LCR_TILE;
        if (slots64 > 0xfffffff0ull || tiles64 > 0xfffffff0ull) { delete db; return LCR_ERR_INVALID_ARG; }
        slot_off[r + 1] = (uint32_t)slots64;
        tile_base[r + 1] = (uint32_t)tiles64;
        pos_off[r + 1] = pos_off[r] + len;
        max_slots = std::max(max_slots, nreads);
        if (nreads >= ctx->big_frag_threshold) big_list.push_back(r);
    }
    db->n_slots = slot_off[b->n_regions];
    db->n_tiles = tile_base[b->n_regions];
    db->n_pos = pos_off[b->n_regions];
    db->max_region_slots = max_slots;
    db->n_big_list = (uint32_t)big_list.size();
    std::vector<uint32_t> slot_region(db->n_slots), tile_region(db->n_tiles);
    for (uint32_t r = 0; r < b->n_regions; ++r) {
        std::fill(slot_region.begin() + slot_off[r], slot_region.begin() + slot_off[r + 1], r);
        std::fill(tile_region.begin() + tile_base[r], tile_region.begin() + tile_base[r + 1], r);
    }
    db->h_pos_off = pos_off;
    db->h_slot_off = slot_off;
    db->extra.h_regions.assign(b->regions, b->regions + b->n_regions);
    db->n_bases = n_bases;
    db->n_cigar = n_cig;
    /* first guesses of the data-dependent capacities; a run that needs more says so in its counters */
    {
        const uint64_t factor = std::max<uint64_t>(1, (db->n_slots + (uint64_t)std::max<uint32_t>(b->n_reads, 1) - 1) / std::max<uint32_t>(b->n_reads, 1));
        LcrCaps &C = db->caps;
        C.items = 3ull * db->n_slots + 4ull * db->n_tiles + 1024;
        C.segs = n_cig * factor + 32ull * db->n_slots + 1024;
        C.pre = db->n_pos / 64 + 65536;
        if (b->ext_off && b->n_regions) C.pre = std::max<uint64_t>(C.pre, 2ull * (b->ext_off[b->n_regions] - b->ext_off[0]) + 65536); /* -v: one pre-candidate per imported record */
        C.elems = 16ull * db->n_slots + (1ull << 20);
        C.pairs = 1ull << 20;
        C.adj = 1ull << 20;
        C.enum_work = 64ull * b->n_regions + 1024;
        /* what earlier batches of this context needed (lcr_submit cuts a run into similar chunks; callers resubmit similar regions) */
        const LcrCaps &Hc = ctx->caps_hint;
        C.items = std::max(C.items, Hc.items); C.segs = std::max(C.segs, Hc.segs); C.pre = std::max(C.pre, Hc.pre); C.elems = std::max(C.elems, Hc.elems);
        C.pairs = std::max(C.pairs, Hc.pairs); C.adj = std::max(C.adj, Hc.adj);
        /* densities seen so far on this context (+ 25 %) */
        auto scaled = [](uint64_t units, uint64_t q10) { return (units * q10 >> 10) + (units * q10 >> 12) + 4096; };
        if (ctx->hint_pre_q10) C.pre = std::max(C.pre, scaled(db->n_pos, ctx->hint_pre_q10));
        if (ctx->hint_elems_q10) C.elems = std::max(C.elems, scaled(db->n_slots, ctx->hint_elems_q10));
        if (ctx->hint_items_q10) C.items = std::max(C.items, scaled(db->n_slots, ctx->hint_items_q10));
        if (ctx->hint_segs_q10) C.segs = std::max(C.segs, scaled(n_cig * factor, ctx->hint_segs_q10));
    }
    uint64_t bytes = 0;
    int rc = 0;
    /* the tables built above and any pageable per-read array go through a page-locked staging block: a pageable source would make
       cudaMemcpyAsync wait for everything queued on the stream before it, i.e. for the previous chunk's bases and qualities */
    const bool pin_soff = !b->n_reads || host_is_pinned(b->seq_off), pin_coff = !b->n_reads || host_is_pinned(b->cig_off), pin_pos = !b->n_reads || host_is_pinned(b->pos),
               pin_flag = !b->n_reads || host_is_pinned(b->flag), pin_mapq = !b->n_reads || host_is_pinned(b->mapq), pin_ts = !b->n_reads || host_is_pinned(b->ts),
               pin_de = !b->n_reads || host_is_pinned(b->de), pin_cig = !n_cig || host_is_pinned(b->cigar), pin_s4off = !packed || (host_is_pinned(b->seq4_off) && b->seq4_off[0] == 0);
    {
        size_t need = 4 * (slot_off.size() + slot_region.size() + tile_base.size() + tile_region.size() + big_list.size() + db->extra.h_status0.size()) + 8 * pos_off.size() +
                      sizeof(lcr_region) * (size_t)b->n_regions + 64 * 16 +
                      (b->exon_off ? 4 * ((size_t)b->n_regions + 1) + 8 * (size_t)(b->exon_off[b->n_regions] - b->exon_off[0]) + 128 : 0) +
                      (b->ext_off ? 4 * ((size_t)b->n_regions + 1) + 9 * (size_t)(b->ext_off[b->n_regions] - b->ext_off[0]) + 512 : 0);
        const size_t nr = b->n_reads;
        need += (pin_soff ? 0 : 8 * (nr + 1)) + (pin_coff ? 0 : 8 * (nr + 1)) + (pin_pos ? 0 : 4 * nr) + (pin_flag ? 0 : 2 * nr) + (pin_mapq ? 0 : nr) + (pin_ts ? 0 : nr) + (pin_de ? 0 : 4 * nr) +
                (pin_s4off ? 0 : 8 * (nr + 1)) + ((pin_cig || n_cig * 4 > (64ull << 20)) ? 0 : 4 * (size_t)n_cig);
        if (need <= (512ull << 20)) stage_take(ctx, db->extra, need);
    }
    DbExtra &X = db->extra;
#define UP(field, src, n) if (!rc) rc = h2d(ctx, &db->field, src, (size_t)(n), &bytes)
#define UPS(field, src, n) if (!rc) rc = h2d(ctx, &db->field, static_cast<decltype(src)>(stage_copy(X, src, sizeof(*(src)) * (size_t)(n))), (size_t)(n), &bytes)
    /* the small host-side tables first (pageable sources: those copies wait for the stream), the caller's large arrays last,
       so that an asynchronous upload returns while seq / qual are still in flight */
    UPS(slot_off, (const uint32_t *)slot_off.data(), slot_off.size());
    UPS(slot_region, (const uint32_t *)slot_region.data(), slot_region.size());
    UPS(tile_base, (const uint32_t *)tile_base.data(), tile_base.size());
    UPS(tile_region, (const uint32_t *)tile_region.data(), tile_region.size());
    UPS(pos_off, (const uint64_t *)pos_off.data(), pos_off.size());
    UPS(status0, (const int32_t *)db->extra.h_status0.data(), db->extra.h_status0.size());
    UPS(big_list, (const uint32_t *)big_list.data(), big_list.size());
    UPS(regions, b->regions, b->n_regions);
    if (b->exon_off && !rc) {
        /* --exon-only: per region the union of its intervals, sorted by start: membership of a position is one binary search on the device
           (the reference asks an interval tree whether anything overlaps [pos + 1, pos + 2): the same set of positions) */
        std::vector<uint32_t> eoff((size_t)b->n_regions + 1, 0);
        std::vector<uint2> eiv;
        std::vector<std::pair<uint32_t, uint32_t>> tmp;
        for (uint32_t r = 0; r < b->n_regions; ++r) {
            tmp.clear();
            if (db->extra.h_status0[r] == 0)
                for (uint32_t e = b->exon_off[r]; e < b->exon_off[r + 1]; ++e)
                    if (b->exon_iv[2 * e + 1] > b->exon_iv[2 * e]) tmp.emplace_back(b->exon_iv[2 * e], b->exon_iv[2 * e + 1]);
            std::sort(tmp.begin(), tmp.end());
            for (const auto &iv : tmp) {
                if (eiv.size() > eoff[r] && iv.first <= eiv.back().y) eiv.back().y = std::max(eiv.back().y, iv.second);
                else eiv.push_back(make_uint2(iv.first, iv.second));
            }
            eoff[r + 1] = (uint32_t)eiv.size();
        }
        rc = h2d(ctx, &db->exon_off, static_cast<const uint32_t *>(stage_copy(X, eoff.data(), 4 * eoff.size())), eoff.size(), &bytes);
        if (!rc) rc = h2d(ctx, &db->exon_iv, static_cast<const uint2 *>(stage_copy(X, eiv.data(), 8 * eiv.size())), eiv.size(), &bytes);
    }
    if (b->ext_off && !rc) { /* -v: the regions' slices of the imported records, offsets rebased to the first one */
        const uint32_t e0 = b->n_regions ? b->ext_off[0] : 0, ne = b->n_regions ? b->ext_off[b->n_regions] - e0 : 0;
        std::vector<uint32_t> xoff((size_t)b->n_regions + 1, 0);
        for (uint32_t r = 0; r <= b->n_regions && b->n_regions; ++r) xoff[r] = b->ext_off[r] - e0;
        rc = h2d(ctx, &db->ext_off, static_cast<const uint32_t *>(stage_copy(X, xoff.data(), 4 * xoff.size())), xoff.size(), &bytes);
        if (!rc) rc = h2d(ctx, &db->ext_pos, static_cast<const uint32_t *>(stage_copy(X, ne ? b->ext_pos + e0 : nullptr, 4 * (size_t)ne)), (size_t)ne, &bytes);
        if (!rc) rc = h2d(ctx, &db->ext_gt, static_cast<const uint8_t *>(stage_copy(X, ne ? b->ext_gt + e0 : nullptr, (size_t)ne)), (size_t)ne, &bytes);
        if (!rc) rc = h2d(ctx, &db->ext_qual, static_cast<const float *>(stage_copy(X, ne ? b->ext_qual + e0 : nullptr, 4 * (size_t)ne)), (size_t)ne, &bytes);
    }
    if (b->n_reads) {
        if (pin_soff) { UP(seq_off, b->seq_off, (size_t)b->n_reads + 1); } else { UPS(seq_off, b->seq_off, (size_t)b->n_reads + 1); }
        if (pin_coff) { UP(cig_off, b->cig_off, (size_t)b->n_reads + 1); } else { UPS(cig_off, b->cig_off, (size_t)b->n_reads + 1); }
    } else { static const uint64_t zero = 0; UP(seq_off, &zero, 1); UP(cig_off, &zero, 1); }
    if (pin_pos) { UP(pos, b->pos, b->n_reads); } else { UPS(pos, b->pos, b->n_reads); }
    if (pin_flag) { UP(flag, b->flag, b->n_reads); } else { UPS(flag, b->flag, b->n_reads); }
    if (pin_mapq) { UP(mapq, b->mapq, b->n_reads); } else { UPS(mapq, b->mapq, b->n_reads); }
    if (pin_ts) { UP(ts, b->ts, b->n_reads); } else { UPS(ts, b->ts, b->n_reads); }
    if (pin_de) { UP(de, b->de, b->n_reads); } else { UPS(de, b->de, b->n_reads); }
    if (pin_cig) { UP(cigar, b->cigar, n_cig); } else { UPS(cigar, b->cigar, n_cig); }
    /* result buffers live as long as the handle */
    if (!rc) rc = dalloc(ctx, &db->rstate, db->n_regions);
    if (!rc) rc = dalloc(ctx, &db->hp, db->n_reads);
    if (!rc) rc = dalloc(ctx, &db->ps, db->n_reads);
    if (!rc) rc = dalloc(ctx, &db->is_fragment, db->n_reads);
    if (!rc) rc = dalloc(ctx, &db->d_stats, 1);
    if (!rc) rc = dalloc(ctx, &db->slot_flags, db->n_slots);
    if (!rc && async) {
        cudaError_t e = cudaEventCreateWithFlags(&db->ev_meta, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(db->ev_meta, ctx->stream);
        if (e != cudaSuccess) { ctx->last_error = cudaGetErrorString(e); ctx->sticky = LCR_ERR_CUDA; rc = ctx->sticky; }
    }
    /* seq / qual carry 64 bytes of slack: bulk copies and block loads read whole 16-byte groups */
    if (!rc && !packed) rc = h2d_padded(ctx, &db->seq, b->seq, (size_t)n_bases, 64, &bytes);
    if (!rc && packed) {
        /* bases in the BAM record's 4-bit form: half the bytes over the bus, expanded here to the letters the kernels compare */
        const uint64_t o0 = b->seq4_off[0], n4 = b->seq4_off[b->n_reads] - o0;
        std::vector<uint64_t> rel;
        const uint64_t *offs = b->seq4_off;
        if (o0) { rel.resize((size_t)b->n_reads + 1); for (uint32_t i = 0; i <= b->n_reads; ++i) rel[i] = b->seq4_off[i] - o0; offs = rel.data(); }
        /* a pageable `rel` may die at the end of this block: cudaMemcpyAsync returns once a pageable source has been staged */
        if (pin_s4off) rc = h2d(ctx, &db->seq4_off, offs, (size_t)b->n_reads + 1, &bytes);
        else rc = h2d(ctx, &db->seq4_off, static_cast<const uint64_t *>(stage_copy(X, offs, 8 * ((size_t)b->n_reads + 1))), (size_t)b->n_reads + 1, &bytes);
        if (!rc) rc = h2d(ctx, &db->seq4, b->seq4 + o0, (size_t)n4, &bytes);
        if (!rc) rc = dalloc(ctx, &db->seq, (size_t)n_bases + 64);
        if (!rc) {
            const cudaError_t e = cudaMemsetAsync(db->seq + n_bases, 0, 64, ctx->stream);
            if (e != cudaSuccess) { ctx->last_error = cudaGetErrorString(e); ctx->sticky = LCR_ERR_CUDA; rc = ctx->sticky; }
        }
        db->seq4_pending = !rc;
        /* an asynchronous upload leaves the expansion to the first run (on the compute stream, after ev_seq): the copy stream only copies */
        if (!rc && !async) rc = lcr_unpack_seq4(ctx, db, ctx->stream);
    }
    db->qual_on_host = false;
    if (!rc && n_bases && (ctx->P.flags & LCR_FLAG_QUAL_ON_DEMAND)) {
        /* qualities are read at candidate sites only: leave a page-locked array where it is and let the kernels fetch those bytes */
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, b->qual) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
            db->qual = static_cast<uint8_t *>(at.devicePointer);
            db->qual_on_host = true;
        } else cudaGetLastError();
    }
    if (!rc && !db->qual_on_host) rc = h2d_padded(ctx, &db->qual, b->qual, (size_t)n_bases, 64, &bytes);
#undef UP
#undef UPS
    if (!rc) {
        cudaError_t e = cudaSuccess;
        if (async) {
            e = cudaEventCreateWithFlags(&db->ev_seq, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventRecord(db->ev_seq, ctx->stream);
        } else e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->last_error = cudaGetErrorString(e); ctx->sticky = LCR_ERR_CUDA; rc = ctx->sticky; }
    }
    if (rc) { lcr_release(ctx, db); return rc; }
    db->h2d_bytes = bytes;
    *out = db;
    return LCR_OK;
}

int lcr_upload(lcr_ctx *ctx, const lcr_batch *b, lcr_device_batch **out) { return upload_impl(ctx, b, out, false); }

int lcr_run_device(lcr_ctx *ctx, lcr_device_batch *dbb) {
    std::unique_lock<std::recursive_mutex> ctx_lock__;
    if (ctx) ctx_lock__ = std::unique_lock<std::recursive_mutex>(ctx->mu);
    if (!ctx || !dbb) return LCR_ERR_INVALID_ARG;
    if (ctx->sticky) return ctx->sticky;
    lcr_device_batch_full *db = static_cast<lcr_device_batch_full *>(dbb);
    TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (db->ev_meta) TRY(cudaStreamWaitEvent(st, db->ev_meta, 0));
    db->ran = false;
    const uint64_t h2d_keep = db->h2d_bytes;
    if (ctx->ref_dirty) {
        const int n = (int)ctx->d_ref.size();
        if (ctx->d_ref_table) { TRY(cudaFree(ctx->d_ref_table)); ctx->d_ref_table = nullptr; }
        TRY(cudaMalloc(&ctx->d_ref_table, sizeof(uint8_t *) * (n ? n : 1)));
        if (n) TRY(cudaMemcpy(ctx->d_ref_table, ctx->d_ref.data(), sizeof(uint8_t *) * n, cudaMemcpyHostToDevice));
        ctx->ref_dirty = false;
    }
    if ((ctx->P.flags & LCR_FLAG_EMIT_PLANES) && !db->pl_acgt) {
        DALLOC(db->pl_acgt, db->n_pos * 4); DALLOC(db->pl_fwd, db->n_pos * 4); DALLOC(db->pl_d, db->n_pos); DALLOC(db->pl_n, db->n_pos); DALLOC(db->pl_ts, db->n_pos * 2);
    }
    uint32_t attempts = 0;
    for (int attempt = 0;; ++attempt) {
        ++attempts;
        /* the asynchronous upload's seq / qual copies may still be in flight: the pileup stage waits where it first reads them */
        db->seq_wait_pending = db->ev_seq != nullptr;
        memset(&db->timing, 0, sizeof db->timing);
        db->timing.h2d_bytes = h2d_keep;
        /* candidates: at most one per pre-candidate */
        if (db->cand_alloc < db->caps.pre) {
            DFREE(db->cand);
            DALLOC(db->cand, db->caps.pre);
            db->cand_alloc = db->caps.pre;
        }
        int rc = plan_or_launch(ctx, db, false);
        if (rc) return rc;
        if (ctx->arena.off > ctx->arena.cap) {
            TRY(cudaStreamSynchronize(st));
            if (ctx->arena.base) { TRY(cudaFree(ctx->arena.base)); ctx->arena.base = nullptr; ctx->arena.cap = 0; }
            const size_t want = ctx->arena.off + ctx->arena.off / 8;
            TRY(cudaMalloc((void **)&ctx->arena.base, want));
            ctx->arena.cap = want;
        }
        TRY(cudaEventRecord(ctx->ev_t[2], st));
        TRY(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(LcrCounters), st));
        TRY(cudaMemsetAsync(db->d_stats, 0, sizeof(lcr_stats), st));
        TRY(cudaMemsetAsync(db->hp, 0xff, db->n_reads ? db->n_reads : 1, st));
        TRY(cudaMemsetAsync(db->ps, 0, sizeof(uint32_t) * (db->n_reads ? db->n_reads : 1), st));
        TRY(cudaMemsetAsync(db->is_fragment, 0, db->n_reads ? db->n_reads : 1, st));
        if (db->n_regions) k_init_rstate<<<(db->n_regions + 127) / 128, 128, 0, st>>>(db->n_regions, db->status0, db->rstate);
        if (db->pl_acgt) {
            const size_t np1 = db->n_pos ? db->n_pos : 1; /* regions that fail on the device keep zeroed planes */
            TRY(cudaMemsetAsync(db->pl_acgt, 0, 16 * np1 / (db->n_pos ? 1 : 4), st)); TRY(cudaMemsetAsync(db->pl_fwd, 0, 16 * np1 / (db->n_pos ? 1 : 4), st));
            TRY(cudaMemsetAsync(db->pl_d, 0, 4 * np1, st)); TRY(cudaMemsetAsync(db->pl_n, 0, 4 * np1, st)); TRY(cudaMemsetAsync(db->pl_ts, 0, 8 * np1 / (db->n_pos ? 1 : 2), st));
        }
        rc = plan_or_launch(ctx, db, true);
        if (rc) return rc;
        TRY(cudaEventRecord(ctx->ev_t[5], st));
        TRY(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(LcrCounters), cudaMemcpyDeviceToHost, st));
        TRY(cudaMemcpyAsync(ctx->h_stats, db->d_stats, sizeof(lcr_stats), cudaMemcpyDeviceToHost, st));
        TRY(cudaStreamSynchronize(st)); /* the one host synchronisation of a run */
        TRY(cudaGetLastError());
        const LcrCounters &K = *ctx->h_ctr;
        db->counters = K;
        if (!K.overflow) break;
        if (attempt >= 4) { ctx->last_error = "run scratch overflow after repeated growth"; return LCR_ERR_OOM; }
        /* grow what overflowed to what the counters ask for (plus slack) and run again */
        LcrCaps &C = db->caps;
        auto grow = [](uint64_t &cap, uint64_t need) { cap = std::max<uint64_t>(cap * 2, need + need / 4 + 1024); };
        if (K.overflow & LCR_OVF_ITEMS) grow(C.items, K.n_items);
        if (K.overflow & LCR_OVF_SEGS) grow(C.segs, K.n_segs);
        if (K.overflow & LCR_OVF_PRE) grow(C.pre, K.n_pre);
        if (K.overflow & LCR_OVF_ELEMS) grow(C.elems, K.n_elem);
        if (K.overflow & LCR_OVF_PAIRS) grow(C.pairs, K.pair_need);
        if (K.overflow & LCR_OVF_ADJ) grow(C.adj, K.adj_total);
        if (K.overflow & LCR_OVF_ENUM) grow(C.enum_work, K.enum_work);
        /* only the capacities that depend on the data's depth and variant density are remembered, not the ones that scale with the batch */
        ctx->caps_hint.pairs = std::max(ctx->caps_hint.pairs, C.pairs);
        ctx->caps_hint.adj = std::max(ctx->caps_hint.adj, C.adj);
    }
    const LcrCounters &K = db->counters;
    { /* remember the densities of this batch for the first attempt of the next upload */
        auto ratio = [](uint64_t need, uint64_t units) { return units ? (need * 1024 + units - 1) / units : 0; };
        const uint64_t cig_units = db->n_cigar * std::max<uint64_t>(1, (db->n_slots + (uint64_t)std::max<uint32_t>(db->n_reads, 1) - 1) / std::max<uint32_t>(db->n_reads, 1));
        ctx->hint_pre_q10 = std::max(ctx->hint_pre_q10, ratio(K.n_pre, db->n_pos));
        ctx->hint_elems_q10 = std::max(ctx->hint_elems_q10, ratio(K.n_elem, db->n_slots));
        ctx->hint_items_q10 = std::max(ctx->hint_items_q10, ratio(K.n_items, db->n_slots));
        ctx->hint_segs_q10 = std::max(ctx->hint_segs_q10, ratio(K.n_segs, cig_units));
    }
    if (getenv("LCR_TILE_PROF")) {
        fprintf(stderr, "tile prof (cycles, consumer warp 0 / producer warp 0, summed over CTAs):");
        for (int i = 0; i < 16; ++i) fprintf(stderr, " [%d]=%llu", i, K.prof[i]);
        fprintf(stderr, " tiles=%u\n", K.n_tiles_done);
    }
    db->n_cand = K.n_cand;
    db->n_frag = (ctx->P.flags & LCR_FLAG_SKIP_PHASING) ? 0 : K.n_frag;
    db->n_elem = (ctx->P.flags & LCR_FLAG_SKIP_PHASING) ? 0 : K.n_elem;
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev_t[0], ctx->ev_t[1]); db->timing.ms_pileup_kernel = ms;
    cudaEventElapsedTime(&ms, ctx->ev_t[2], ctx->ev_t[3]); db->timing.ms_pileup = ms;
    cudaEventElapsedTime(&ms, ctx->ev_t[3], ctx->ev_t[4]); db->timing.ms_fragments = ms;
    cudaEventElapsedTime(&ms, ctx->ev_t[4], ctx->ev_t[5]); db->timing.ms_phase = ms;
    cudaEventElapsedTime(&ms, ctx->ev_t[2], ctx->ev_t[5]); db->timing.ms_total = ms;
    cudaEventElapsedTime(&ms, ctx->ev_t[2], ctx->ev_t[0]); db->timing.ms_prep = ms;
    cudaEventElapsedTime(&ms, ctx->ev_t[4], ctx->ev_t[6]); db->timing.ms_enum = ms;
    cudaEventElapsedTime(&ms, ctx->ev_t[6], ctx->ev_t[7]); db->timing.ms_phase_kernel = ms;
    db->timing.run_attempts = attempts;
    db->timing.n_segments = K.n_segs_used; db->timing.n_items = K.n_items_used; db->timing.n_tiles = K.n_tiles_done;
    /* one sweep iteration (sigma pass + delta / eta pass) touches every phase-site cell twice (1 B cell + 4 B index), every fragment's row
       pointer and haplotag, every site's column pointer and state: B_sweep of SURVEY 8(d), times the iterations executed */
    db->timing.phase_alg_bytes = (10ull * ctx->h_stats->nnz_phase + 6ull * db->n_frag + 8ull * K.n_cand) * (ctx->h_stats->n_cross_optimize ? ctx->h_stats->n_sweep_iters / std::max<uint64_t>(1, ctx->h_stats->n_cross_optimize) : 0) ;
    /* algorithmic bytes of the tile kernel = what it has to read and write: the base of every aligned base (it does not read qualities:
       those are fetched by k_site_ll at the surviving sites only), the segment and item descriptors it stages, one tile descriptor and the
       reference bytes per processed tile, the surviving sites it writes */
    db->timing.pileup_alg_bytes = 1ull * ctx->h_stats->n_aligned_bases + 16ull * K.n_segs_used + 16ull * K.n_items_used + 48ull * K.n_tiles_done + K.n_pos_done +
                                  72ull * std::min<uint64_t>(K.n_pre, db->caps.pre);
    /* LCR_FLAG_QUAL_ON_DEMAND: the qualities fetched from the caller's page-locked array cross the bus in 32-byte sectors: k_site_ll's
       reads are counted on the device, the fragment build reads one per element in each of its two passes */
    if (db->qual_on_host) db->timing.h2d_bytes += 32ull * (K.qual_reads + 2ull * K.n_elem);
    if ((ctx->P.flags & LCR_FLAG_EMIT_FRAGMENTS) && !(ctx->P.flags & LCR_FLAG_SKIP_PHASING)) {
        FragDebug &fd = db->extra.fragdbg;
        const uint32_t nf = db->n_frag;
        const uint64_t ne = db->n_elem;
        fd.frag_slot.resize(nf); fd.frag_elem_off.resize((size_t)nf + 1);
        fd.elem_snp.resize(ne); fd.elem_cell.resize(ne); fd.elem_base.resize(ne);
        if (nf) TRY(cudaMemcpyAsync(fd.frag_slot.data(), db->fr_frag_off, 4ull * nf, cudaMemcpyDeviceToHost, st));
        TRY(cudaMemcpyAsync(fd.frag_elem_off.data(), db->fr_frag_read, 4ull * ((size_t)nf + 1), cudaMemcpyDeviceToHost, st));
        if (ne) {
            TRY(cudaMemcpyAsync(fd.elem_snp.data(), db->fr_elem_snp, 4ull * ne, cudaMemcpyDeviceToHost, st));
            TRY(cudaMemcpyAsync(fd.elem_cell.data(), db->fr_elem_cell, ne, cudaMemcpyDeviceToHost, st));
            TRY(cudaMemcpyAsync(fd.elem_base.data(), db->fr_elem_base, ne, cudaMemcpyDeviceToHost, st));
        }
        TRY(cudaStreamSynchronize(st));
        if (!nf) fd.frag_elem_off[0] = 0;
    }
    db->ran = true;
    return LCR_OK;
}

int lcr_fetch(lcr_ctx *ctx, lcr_device_batch *dbb, lcr_result **out) {
    std::unique_lock<std::recursive_mutex> ctx_lock__;
    if (ctx) ctx_lock__ = std::unique_lock<std::recursive_mutex>(ctx->mu);
    if (!ctx || !dbb || !out) return LCR_ERR_INVALID_ARG;
    if (ctx->sticky) return ctx->sticky;
    lcr_device_batch_full *db = static_cast<lcr_device_batch_full *>(dbb);
    if (!db->ran) return LCR_ERR_INVALID_ARG;
    TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ResultBox *box = new (std::nothrow) ResultBox();
    if (!box) return LCR_ERR_OOM;
    const uint32_t nr = db->n_regions, nreads = db->n_reads, nc = db->n_cand;
    std::vector<LcrRegionState> hrs(nr);
    box->cand.resize(nc);
    box->hp.resize(nreads); box->ps.resize(nreads); box->is_fragment.resize(nreads);
    uint64_t bytes = 0;
    auto d2h = [&](void *dst, const void *src, size_t n) -> cudaError_t {
        if (!n) return cudaSuccess;
        bytes += n;
        return cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, st);
    };
    lcr_stats hs{};
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = d2h(hrs.data(), db->rstate, sizeof(LcrRegionState) * nr);
    if (e == cudaSuccess) e = d2h(box->cand.data(), db->cand, sizeof(lcr_candidate) * nc);
    if (e == cudaSuccess) e = d2h(box->hp.data(), db->hp, nreads);
    if (e == cudaSuccess) e = d2h(box->ps.data(), db->ps, 4ull * nreads);
    if (e == cudaSuccess) e = d2h(box->is_fragment.data(), db->is_fragment, nreads);
    if (e == cudaSuccess) e = d2h(&hs, db->d_stats, sizeof hs);
    const bool planes = db->pl_acgt != nullptr;
    if (planes) {
        const uint64_t np = db->n_pos;
        box->acgt.resize(np * 4); box->fwd.resize(np * 4); box->d.resize(np); box->n.resize(np); box->ts.resize(np * 2);
        if (e == cudaSuccess) e = d2h(box->acgt.data(), db->pl_acgt, 16 * np);
        if (e == cudaSuccess) e = d2h(box->fwd.data(), db->pl_fwd, 16 * np);
        if (e == cudaSuccess) e = d2h(box->d.data(), db->pl_d, 4 * np);
        if (e == cudaSuccess) e = d2h(box->n.data(), db->pl_n, 4 * np);
        if (e == cudaSuccess) e = d2h(box->ts.data(), db->pl_ts, 8 * np);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { delete box; ctx->last_error = cudaGetErrorString(e); ctx->sticky = LCR_ERR_CUDA; return ctx->sticky; }
    db->timing.d2h_bytes = bytes;
    box->cand_off.assign(nr + 1, 0);
    box->region_status.assign(nr, 0);
    for (uint32_t r = 0; r < nr; ++r) {
        box->region_status[r] = hrs[r].status;
        box->cand_off[r + 1] = box->cand_off[r] + (hrs[r].status == 0 || true ? hrs[r].n_cand : 0);
    }
    lcr_result &res = box->res;
    res.n_regions = nr; res.n_reads = nreads; res.n_cand = nc;
    res.cand_off = box->cand_off.data();
    res.cand = box->cand.data();
    res.region_status = box->region_status.data();
    res.hp = box->hp.data(); res.ps = box->ps.data(); res.is_fragment = box->is_fragment.data();
    hs.n_positions = db->n_pos;
    hs.n_candidates = nc;
    hs.n_fragments = db->n_frag;
    res.stats = hs;
    if (planes) {
        box->pos_off = db->h_pos_off;
        res.planes.n_pos = db->n_pos;
        res.planes.pos_off = box->pos_off.data();
        res.planes.acgt = box->acgt.data(); res.planes.fwd = box->fwd.data(); res.planes.d = box->d.data();
        res.planes.n = box->n.data(); res.planes.ts = box->ts.data();
    }
    if ((ctx->P.flags & LCR_FLAG_EMIT_FRAGMENTS) && !(ctx->P.flags & LCR_FLAG_SKIP_PHASING)) {
        const FragDebug &fd = db->extra.fragdbg;
        /* fragments of regions that failed are not reported (the reference would have panicked there) */
        box->frag_off.assign(nr + 1, 0);
        box->elem_off.assign(1, 0);
        for (uint32_t r = 0; r < nr; ++r) {
            if (hrs[r].status == 0) {
                for (uint32_t f = hrs[r].frag_begin; f < hrs[r].frag_begin + hrs[r].n_frag; ++f) {
                    const uint32_t slot = fd.frag_slot[f];
                    box->frag_read.push_back(db->extra.h_regions[r].read_begin + (slot - db->h_slot_off[r]));
                    for (uint32_t e = fd.frag_elem_off[f]; e < fd.frag_elem_off[f + 1]; ++e) {
                        box->elem_snp.push_back(fd.elem_snp[e]);
                        box->elem_cell.push_back(fd.elem_cell[e]);
                        box->elem_base.push_back(fd.elem_base[e]);
                    }
                    box->elem_off.push_back(box->elem_snp.size());
                }
            }
            box->frag_off[r + 1] = (uint32_t)box->frag_read.size();
        }
        const size_t nf = box->frag_read.size();
        res.fragments.n_frag = nf;
        res.fragments.n_elem = box->elem_snp.size();
        res.fragments.frag_off = box->frag_off.data();
        res.fragments.frag_read = box->frag_read.data();
        res.fragments.elem_off = box->elem_off.data();
        res.fragments.elem_snp = box->elem_snp.data();
        res.fragments.elem_cell = box->elem_cell.data();
        res.fragments.elem_base = box->elem_base.data();
    }
    *out = &box->res;
    return LCR_OK;
}

void lcr_free_result(lcr_result *res) {
    if (res) delete reinterpret_cast<ResultBox *>(res);
}

void lcr_release(lcr_ctx *ctx, lcr_device_batch *dbb) {
    std::unique_lock<std::recursive_mutex> ctx_lock__;
    if (ctx) ctx_lock__ = std::unique_lock<std::recursive_mutex>(ctx->mu);
    if (!ctx || !dbb) return;
    lcr_device_batch_full *db = static_cast<lcr_device_batch_full *>(dbb);
    cudaSetDevice(ctx->device);
    DFREE(db->rstate); DFREE(db->cand); DFREE(db->hp); DFREE(db->ps); DFREE(db->is_fragment); DFREE(db->d_stats);
    DFREE(db->pl_acgt); DFREE(db->pl_fwd); DFREE(db->pl_d); DFREE(db->pl_n); DFREE(db->pl_ts);
    DFREE(db->slot_flags); DFREE(db->status0); DFREE(db->big_list);
    DFREE(db->regions); DFREE(db->pos); DFREE(db->flag); DFREE(db->mapq); DFREE(db->ts); DFREE(db->de);
    DFREE(db->seq_off); DFREE(db->cig_off); DFREE(db->seq); if (db->qual_on_host) db->qual = nullptr; DFREE(db->qual); DFREE(db->cigar);
    DFREE(db->slot_off); DFREE(db->slot_region); DFREE(db->tile_base); DFREE(db->tile_region); DFREE(db->pos_off);
    DFREE(db->seq4); DFREE(db->seq4_off); DFREE(db->exon_off); DFREE(db->exon_iv); DFREE(db->ext_off); DFREE(db->ext_pos); DFREE(db->ext_gt); DFREE(db->ext_qual);
    cudaStreamSynchronize(ctx->stream);
    if (db->ev_seq) { cudaEventSynchronize(db->ev_seq); cudaEventDestroy(db->ev_seq); }
    stage_give_back(ctx, db->extra); /* after the copies out of it have completed */
    if (db->ev_meta) cudaEventDestroy(db->ev_meta);
    delete db;
}

int lcr_device_results(lcr_ctx *ctx, lcr_device_batch *db, lcr_device_view *out) {
    if (!ctx || !db || !out || !db->ran) return LCR_ERR_INVALID_ARG;
    out->cand = db->cand; out->hp = db->hp; out->ps = db->ps;
    out->n_cand = db->n_cand; out->n_reads = db->n_reads;
    return LCR_OK;
}

int lcr_get_timing(lcr_ctx *ctx, lcr_device_batch *db, lcr_timing *out) {
    if (!ctx || !db || !out) return LCR_ERR_INVALID_ARG;
    *out = db->timing;
    return LCR_OK;
}

int lcr_last_submit_timing(lcr_ctx *ctx, lcr_timing *out) {
    std::unique_lock<std::recursive_mutex> ctx_lock__;
    if (ctx) ctx_lock__ = std::unique_lock<std::recursive_mutex>(ctx->mu);
    if (!ctx || !out) return LCR_ERR_INVALID_ARG;
    *out = ctx->last_submit;
    return LCR_OK;
}

static void add_timing(lcr_timing &acc, const lcr_timing &t) {
    acc.ms_total += t.ms_total; acc.ms_pileup += t.ms_pileup; acc.ms_pileup_kernel += t.ms_pileup_kernel;
    acc.ms_fragments += t.ms_fragments; acc.ms_phase += t.ms_phase; acc.kernel_launches += t.kernel_launches;
    acc.pileup_alg_bytes += t.pileup_alg_bytes; acc.h2d_bytes += t.h2d_bytes; acc.d2h_bytes += t.d2h_bytes;
}

static int submit_one(lcr_ctx *ctx, const lcr_batch *batch, lcr_result **out) {
    lcr_device_batch *db = nullptr;
    int rc = upload_impl(ctx, batch, &db, true); /* copies on the copy stream; the run waits on events, not on the host */
    if (rc) return rc;
    rc = lcr_run_device(ctx, db);
    if (!rc) rc = lcr_fetch(ctx, db, out);
    if (!rc) add_timing(ctx->last_submit, db->timing);
    lcr_release(ctx, db);
    return rc;
}

/* One chunk of a large submit: consecutive regions and the read rows they span, offsets rebased to the chunk. */
struct SubmitChunk {
    uint32_t r0, r1, read_lo, read_hi;
    std::vector<lcr_region> regions;
    std::vector<uint64_t> seq_off, cig_off;
    lcr_batch view;
};

/* The worker body for a batch of regions, host buffers in, host results out.  Large batches are cut into chunks of
   consecutive regions; the host-to-device copies of chunk k+1 run on the copy stream while chunk k computes (regions are
   independent, so the result is the same as one pass over the whole batch). */
int lcr_submit(lcr_ctx *ctx, const lcr_batch *batch, lcr_result **out) {
    std::unique_lock<std::recursive_mutex> ctx_lock__;
    if (ctx) ctx_lock__ = std::unique_lock<std::recursive_mutex>(ctx->mu);
    if (!ctx || !batch || !out) return LCR_ERR_INVALID_ARG;
    if (ctx->sticky) return ctx->sticky;
    memset(&ctx->last_submit, 0, sizeof ctx->last_submit);
    const size_t chunk_bytes = ctx->submit_chunk_bytes; /* seq + qual bytes per chunk (LCR_SUBMIT_CHUNK_MB at lcr_create: tests) */
    const uint64_t total_bases = batch->n_reads && batch->seq_off ? batch->seq_off[batch->n_reads] : 0;
    const bool debug_out = ctx->P.flags & (LCR_FLAG_EMIT_PLANES | LCR_FLAG_EMIT_FRAGMENTS);
    bool chunkable = !debug_out && batch->n_regions > 1 && 2 * total_bases > chunk_bytes + chunk_bytes / 2 && batch->regions && batch->seq_off && batch->cig_off;
    std::vector<SubmitChunk> chunks;
    if (chunkable) {
        /* greedy cut by the bases of the reads each region spans; regions with unusable read ranges ride along */
        uint32_t r = 0;
        while (r < batch->n_regions) {
            SubmitChunk ck;
            ck.r0 = r;
            ck.read_lo = 0xffffffffu; ck.read_hi = 0;
            uint64_t bases = 0;
            /* the last chunk's run is the only one no copy hides: when what is left fits 1.25 chunks, cut it so that a quarter chunk comes last */
            uint64_t limit = chunk_bytes;
            {
                const lcr_region &g0 = batch->regions[r];
                const bool ok0 = g0.read_end >= g0.read_begin && g0.read_end <= batch->n_reads;
                const uint64_t left = ok0 ? 2 * (total_bases - batch->seq_off[g0.read_begin]) : 0;
                if (ok0 && left > chunk_bytes / 2 && left <= chunk_bytes + chunk_bytes / 4) limit = left - chunk_bytes / 4;
            }
            while (r < batch->n_regions) {
                const lcr_region &g = batch->regions[r];
                const bool ok = g.read_end >= g.read_begin && g.read_end <= batch->n_reads;
                if (ok && g.read_end > g.read_begin) {
                    const uint32_t lo = std::min(ck.read_lo, g.read_begin), hi = std::max(ck.read_hi, g.read_end);
                    const uint64_t nb = batch->seq_off[hi] - batch->seq_off[lo];
                    if (r > ck.r0 && 2 * nb > limit) break;
                    ck.read_lo = lo; ck.read_hi = hi; bases = nb;
                }
                ++r;
            }
            (void)bases;
            ck.r1 = r;
            if (ck.read_lo > ck.read_hi) { ck.read_lo = 0; ck.read_hi = 0; }
            chunks.push_back(std::move(ck));
        }
        if (chunks.size() < 2) chunkable = false;
    }
    if (!chunkable) return submit_one(ctx, batch, out);

    for (SubmitChunk &ck : chunks) {
        const uint32_t nreads = ck.read_hi - ck.read_lo;
        ck.regions.assign(batch->regions + ck.r0, batch->regions + ck.r1);
        for (lcr_region &g : ck.regions) {
            const bool ok = g.read_end >= g.read_begin && g.read_end <= batch->n_reads;
            if (ok && g.read_end > g.read_begin) { g.read_begin -= ck.read_lo; g.read_end -= ck.read_lo; }
            else if (ok) { g.read_begin = 0; g.read_end = 0; }
            else { g.read_begin = 1; g.read_end = 0; } /* stays invalid */
        }
        ck.seq_off.resize((size_t)nreads + 1);
        ck.cig_off.resize((size_t)nreads + 1);
        const uint64_t sb = batch->seq_off[ck.read_lo], cb = batch->cig_off[ck.read_lo];
        for (uint32_t i = 0; i <= nreads; ++i) { ck.seq_off[i] = batch->seq_off[ck.read_lo + i] - sb; ck.cig_off[i] = batch->cig_off[ck.read_lo + i] - cb; }
        lcr_batch &v = ck.view;
        v.n_regions = ck.r1 - ck.r0; v.n_reads = nreads; v.regions = ck.regions.data();
        v.pos = batch->pos + ck.read_lo; v.flag = batch->flag + ck.read_lo; v.mapq = batch->mapq + ck.read_lo;
        v.ts = batch->ts + ck.read_lo; v.de = batch->de + ck.read_lo;
        v.seq_off = ck.seq_off.data(); v.cig_off = ck.cig_off.data();
        v.seq = batch->seq ? batch->seq + sb : nullptr; v.qual = batch->qual ? batch->qual + sb : nullptr;
        v.seq4 = batch->seq4; v.seq4_off = batch->seq4 && batch->seq4_off ? batch->seq4_off + ck.read_lo : nullptr; /* absolute offsets: upload_impl rebases */
        v.exon_off = batch->exon_off ? batch->exon_off + ck.r0 : nullptr; v.exon_iv = batch->exon_iv; /* absolute interval offsets */
        v.ext_off = batch->ext_off ? batch->ext_off + ck.r0 : nullptr; v.ext_pos = batch->ext_pos; v.ext_gt = batch->ext_gt; v.ext_qual = batch->ext_qual;
        v.cigar = batch->cigar ? batch->cigar + cb : nullptr;
    }

    ResultBox *all = new (std::nothrow) ResultBox();
    if (!all) return LCR_ERR_OOM;
    all->cand_off.assign((size_t)batch->n_regions + 1, 0);
    all->region_status.assign(batch->n_regions, 0);
    all->hp.assign(batch->n_reads, (int8_t)-1);
    all->ps.assign(batch->n_reads, 0);
    all->is_fragment.assign(batch->n_reads, 0);
    lcr_stats st{};
    int rc = 0;
    /* uploads run ahead of the runs on the copy stream (bounded by the bytes they pin on the device): the bus stays busy while a chunk computes */
    std::vector<lcr_device_batch *> up(chunks.size(), nullptr);
    std::vector<uint64_t> up_bytes(chunks.size(), 0);
    for (size_t k = 0; k < chunks.size(); ++k) up_bytes[k] = 3 * (chunks[k].seq_off.empty() ? 0 : chunks[k].seq_off.back());
    const uint64_t ahead_cap = 12ull << 30;
    uint64_t ahead = 0;
    size_t next_up = 0;
    auto pump = [&](size_t k_cur) {
        while (!rc && next_up < chunks.size() && (next_up <= k_cur + 1 || ahead + up_bytes[next_up] <= ahead_cap)) {
            rc = upload_impl(ctx, &chunks[next_up].view, &up[next_up], true);
            ahead += up_bytes[next_up];
            ++next_up;
        }
    };
    static const bool prof = getenv("LCR_SUBMIT_PROF") != nullptr; /* bring-up aid: host wall time per phase of the chunk loop */
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_up = 0, t_run = 0, t_fetch = 0, t_merge = 0, t0 = now();
    for (size_t k = 0; k < chunks.size() && !rc; ++k) {
        t0 = now();
        pump(k);
        lcr_device_batch *cur = up[k];
        t_up += now() - t0; t0 = now();
        lcr_result *part = nullptr;
        if (!rc) rc = lcr_run_device(ctx, cur);
        t_run += now() - t0; t0 = now();
        if (!rc) rc = lcr_fetch(ctx, cur, &part);
        t_fetch += now() - t0; t0 = now();
        if (!rc) {
            const SubmitChunk &ck = chunks[k];
            add_timing(ctx->last_submit, cur->timing);
            const uint32_t base = (uint32_t)all->cand.size();
            for (uint32_t i = 0; i < part->n_cand; ++i) {
                lcr_candidate c = part->cand[i];
                c.region += ck.r0;
                all->cand.push_back(c);
            }
            for (uint32_t r = 0; r < part->n_regions; ++r) {
                all->region_status[ck.r0 + r] = part->region_status[r];
                all->cand_off[ck.r0 + r + 1] = base + part->cand_off[r + 1];
            }
            for (uint32_t i = 0; i < part->n_reads; ++i) {
                const size_t g = (size_t)ck.read_lo + i;
                /* chunks are consecutive regions: the first chunk with an entry is the lowest region (same rule as on the device) */
                if (part->hp[i] != -1 && all->hp[g] == -1) all->hp[g] = part->hp[i];
                if (part->ps[i] && !all->ps[g]) all->ps[g] = part->ps[i];
                if (part->is_fragment[i]) all->is_fragment[g] = 1;
            }
            st.n_reads_pass += part->stats.n_reads_pass; st.n_aligned_bases += part->stats.n_aligned_bases;
            st.n_positions += part->stats.n_positions; st.n_candidates += part->stats.n_candidates;
            st.n_fragments += part->stats.n_fragments; st.nnz_phase += part->stats.nnz_phase;
            st.n_cross_optimize += part->stats.n_cross_optimize; st.n_sweep_iters += part->stats.n_sweep_iters;
            lcr_free_result(part);
        }
        if (cur) lcr_release(ctx, cur);
        up[k] = nullptr;
        ahead -= up_bytes[k];
        t_merge += now() - t0;
    }
    if (prof) fprintf(stderr, "lcr_submit: %zu chunks, host ms: upload %.2f run %.2f fetch %.2f merge+release %.2f\n", chunks.size(), t_up, t_run, t_fetch, t_merge);
    for (lcr_device_batch *h : up) if (h) lcr_release(ctx, h);
    if (rc) { delete all; return rc; }
    lcr_result &res = all->res;
    res.n_regions = batch->n_regions; res.n_reads = batch->n_reads; res.n_cand = (uint32_t)all->cand.size();
    res.cand_off = all->cand_off.data(); res.cand = all->cand.data(); res.region_status = all->region_status.data();
    res.hp = all->hp.data(); res.ps = all->ps.data(); res.is_fragment = all->is_fragment.data();
    res.stats = st;
    *out = &all->res;
    return LCR_OK;
}

} /* extern "C" */
