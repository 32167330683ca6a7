/*
 * regions.cu — isolated-region discovery on the device.
 *
 * Replaces find_isolated_regions_with_depth (reference src/util.rs:236-332), the reference's BAM pass #1:
 *   - depth over [reference_start, reference_end) of every read that passes the filter of util.rs:262-279
 *     (introns included), built here as a difference array (+1 / -1 per read) and one prefix sum over the
 *     concatenated contigs instead of the reference's per-position increments (util.rs:283-285);
 *   - a position is "covered" when depth > 0 and not (truncation && depth > truncation_coverage) (util.rs:294-296);
 *     maximal covered runs are found in one pass over the depth vector;
 *   - the reference's run-to-region state machine (util.rs:297-330) pushes a region only when region_end > region_start
 *     and resets its state only when it pushes, so a covered run of ONE position is not dropped: it stays the start of
 *     the next region, which then extends to the end of the following run.  Regions are therefore one run (length >= 2)
 *     or a length-1 run joined with the run after it; a length-1 run that is the last of its contig is dropped.  Which
 *     length-1 runs open a region is the parity of the chain of length-1 runs before them;
 *   - max_coverage is the largest depth since the previous push up to and including the position that triggered this one
 *     (util.rs:291-293 runs before the push test), so it can exceed truncation_coverage;
 *   - read_begin / read_end: the contiguous superset of fetch((chr, start, end)) that lcr_host_find_regions returns
 *     (prefix maximum of the read end positions, then two binary searches per region).
 *
 * Layout: the depth vector holds contig_len + 1 int32 per contig, contigs back to back; the extra element of every
 * contig always ends at depth 0, so runs never cross contigs and one global scan equals the per-contig scans.
 */
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "lcr_device.h"

namespace {

struct RgIn {
    uint32_t n_reads, n_contigs;
    const int32_t *tid, *pos;
    const uint16_t *flag;
    const uint8_t *mapq;
    const float *de;
    const uint64_t *seq_off, *cig_off;
    const uint32_t *cigar;
    const uint64_t *base; /* [n_contigs+1] first element of every contig in the depth vector */
    int32_t min_mapq, min_read_length;
    float divergence;
    int truncation;
    uint32_t trunc_cov;
};

struct RgCtr {
    uint32_t n_start, n_end, n_regions, pad;
};

__device__ __forceinline__ bool rg_covered(int32_t d, int trunc, uint32_t cov) { return d > 0 && !(trunc && (uint32_t)d > cov); }

/* largest c with base[c] <= g */
__device__ __forceinline__ uint32_t rg_contig_of(const uint64_t *base, uint32_t n_contigs, uint64_t g) {
    uint32_t lo = 0, hi = n_contigs;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (base[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

/* one thread per read: the filter of util.rs:262-279, the reference span, +1 / -1 into the difference array */
__global__ void k_rg_span(RgIn a, int32_t *diff, unsigned long long *endkey) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_reads; i += gridDim.x * blockDim.x) {
        int64_t span = 0;
        for (uint64_t c = a.cig_off[i]; c < a.cig_off[i + 1]; ++c) {
            const uint32_t op = a.cigar[c] & 0xf, len = a.cigar[c] >> 4;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += len;
        }
        const int64_t p = a.pos[i], endpos = p + span;
        int64_t e = endpos > p ? endpos : p + 1;
        if (e < 0) e = 0;
        const int32_t t = a.tid[i];
        endkey[i] = ((unsigned long long)(uint32_t)(t + 1) << 40) | (unsigned long long)e;
        if (t < 0 || (uint32_t)t >= a.n_contigs) continue;
        const uint64_t l_seq = a.seq_off[i + 1] - a.seq_off[i];
        const uint16_t f = a.flag[i];
        if ((int32_t)a.mapq[i] < a.min_mapq || l_seq < (uint64_t)a.min_read_length || (f & 0x4) || (f & 0x100) || (f & 0x800)) continue;
        const float de = a.de[i];
        if (!(de != de) && de >= a.divergence) continue;
        const int64_t L = (int64_t)(a.base[t + 1] - a.base[t]) - 1;
        const int64_t s = p > 0 ? p : 0, b = endpos < L ? endpos : L;
        if (b > s) {
            atomicAdd(diff + a.base[t] + s, 1);
            atomicAdd(diff + a.base[t] + b, -1);
        }
    }
}

/* one pass over the depth vector, four positions per thread: the first and last position of every covered run */
__global__ void k_rg_runs(const int32_t *depth, uint64_t G, int trunc, uint32_t cov, unsigned long long *starts, unsigned long long *ends, uint32_t cap, RgCtr *ctr) {
    const uint64_t n4 = (G + 3) / 4;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g0 = q * 4;
        int32_t d[6];
        d[0] = g0 ? depth[g0 - 1] : 0;
        if (g0 + 4 <= G) {
            const int4 v = *reinterpret_cast<const int4 *>(depth + g0);
            d[1] = v.x; d[2] = v.y; d[3] = v.z; d[4] = v.w;
        } else {
            for (int k = 0; k < 4; ++k) d[1 + k] = g0 + k < G ? depth[g0 + k] : 0;
        }
        if ((d[1] | d[2] | d[3] | d[4]) == 0) continue; /* run starts and ends are recorded at covered positions only */
        d[5] = g0 + 4 < G ? depth[g0 + 4] : 0;
        bool c[6];
        for (int k = 0; k < 6; ++k) c[k] = rg_covered(d[k], trunc, cov);
        for (int k = 1; k <= 4; ++k) {
            if (!c[k]) continue;
            if (!c[k - 1]) {
                const uint32_t s = atomicAdd(&ctr->n_start, 1u);
                if (s < cap) starts[s] = g0 + k - 1;
            }
            if (!c[k + 1]) {
                const uint32_t s = atomicAdd(&ctr->n_end, 1u);
                if (s < cap) ends[s] = g0 + k - 1;
            }
        }
    }
}

/* one thread per run: does it open a region (parity of the chain of length-1 runs before it), and which run closes it */
__global__ void k_rg_group(const unsigned long long *S, const unsigned long long *E, const uint64_t *base, uint32_t n_contigs, const RgCtr *ctr, uint32_t *flag) {
    const uint32_t n_runs = ctr->n_start;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_runs; k += gridDim.x * blockDim.x) {
        const uint32_t c = rg_contig_of(base, n_contigs, S[k]);
        uint32_t m = 0; /* length-1 runs of this contig immediately before run k */
        for (uint32_t j = k; j > 0; --j) {
            if (S[j - 1] < base[c] || S[j - 1] != E[j - 1]) break;
            ++m;
        }
        uint32_t f = 0;
        if ((m & 1u) == 0) {
            if (S[k] != E[k]) f = 1;                                     /* a run of two or more positions is a region on its own */
            else if (k + 1 < n_runs && S[k + 1] < base[c + 1]) f = 1;    /* a single position waits for the next run of its contig */
        }
        flag[k] = f;
    }
}

__global__ void k_rg_emit(const unsigned long long *S, const unsigned long long *E, const uint64_t *base, uint32_t n_contigs, RgCtr *ctr, const uint32_t *flag, const uint32_t *scan,
                          lcr_region *regions, unsigned long long *push_g, uint32_t *max_cov) {
    const uint32_t n_runs = ctr->n_start;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_runs; k += gridDim.x * blockDim.x) {
        if (k == n_runs - 1) ctr->n_regions = scan[k] + (flag[k] ? 1u : 0u);
        if (!flag[k]) continue;
        const uint32_t c = rg_contig_of(base, n_contigs, S[k]);
        const uint64_t a = S[k] - base[c], b = E[S[k] == E[k] ? k + 1 : k] - base[c];
        lcr_region rg;
        rg.tid = (int32_t)c;
        rg.start = (uint32_t)(a + 1);
        rg.end = (uint32_t)(b + 2);
        rg.read_begin = rg.read_end = 0;
        regions[scan[k]] = rg;
        push_g[scan[k]] = base[c] + b + 1; /* the position whose depth test pushes the region (the contig's spare element when the run reaches the contig end) */
        max_cov[scan[k]] = 0;
    }
}

/* max_coverage: every position belongs to the first region pushed at or after it in its contig; 16 positions per thread */
__global__ void k_rg_maxcov(const int32_t *depth, uint64_t G, const uint64_t *base, uint32_t n_contigs, const RgCtr *ctr, const lcr_region *regions, const unsigned long long *push_g, uint32_t *max_cov) {
    const uint32_t n = ctr->n_regions;
    if (!n) return;
    const uint64_t nchunk = (G + 15) / 16;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nchunk; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g0 = q * 16, g1 = g0 + 16 < G ? g0 + 16 : G;
        int32_t d[16];
        if (g0 + 16 <= G) {
            for (int k = 0; k < 4; ++k) {
                const int4 v = reinterpret_cast<const int4 *>(depth + g0)[k];
                d[4 * k] = v.x; d[4 * k + 1] = v.y; d[4 * k + 2] = v.z; d[4 * k + 3] = v.w;
            }
        } else {
            for (int k = 0; k < 16; ++k) d[k] = g0 + k < G ? depth[g0 + k] : 0;
        }
        int32_t any = 0;
        for (int k = 0; k < 16; ++k) any |= d[k];
        if (!any) continue;
        uint32_t lo = 0, hi = n; /* first region with push_g >= g0 */
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (push_g[mid] < g0) lo = mid + 1; else hi = mid;
        }
        uint32_t r = lo, best = 0;
        for (uint64_t g = g0; g < g1 && r < n; ++g) {
            while (r < n && push_g[r] < g) {
                if (best) atomicMax(max_cov + r, best);
                best = 0;
                ++r;
            }
            if (r >= n) break;
            if (g < base[regions[r].tid]) continue; /* the tail of the previous contig after its last push counts for nobody */
            const uint32_t dv = (uint32_t)d[g - g0];
            if (dv > best) best = dv;
        }
        if (r < n && best) atomicMax(max_cov + r, best);
    }
}

/* rows [read_begin, read_end) of every region */
__global__ void k_rg_reads(const RgCtr *ctr, lcr_region *regions, const int32_t *tid, const int32_t *pos, const unsigned long long *pmax, uint32_t n_reads) {
    const uint32_t n = ctr->n_regions;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        lcr_region rg = regions[r];
        auto key = [&](uint32_t i) { const int32_t t = tid[i]; return t < 0 ? 0x7fffffff : t; };
        uint32_t lo = 0, hi = n_reads;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (key(mid) < rg.tid) lo = mid + 1; else hi = mid; }
        const uint32_t c0 = lo;
        hi = n_reads;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (key(mid) <= rg.tid) lo = mid + 1; else hi = mid; }
        const uint32_t c1 = lo;
        uint32_t b0 = c0, b1 = c1; /* first read whose prefix-maximum end is above start */
        while (b0 < b1) { const uint32_t mid = (b0 + b1) >> 1; if ((pmax[mid] & 0xffffffffffull) <= (unsigned long long)rg.start) b0 = mid + 1; else b1 = mid; }
        uint32_t e0 = c0, e1 = c1; /* first read that starts at or after end */
        while (e0 < e1) { const uint32_t mid = (e0 + e1) >> 1; if ((int64_t)pos[mid] < (int64_t)rg.end) e0 = mid + 1; else e1 = mid; }
        if (e0 < b0) e0 = b0;
        regions[r].read_begin = b0;
        regions[r].read_end = e0;
    }
}

struct MaxU64 {
    __device__ __forceinline__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a > b ? a : b; }
};

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

} // namespace

extern "C" int lcr_discover_regions(lcr_ctx *ctx, const lcr_align_index *in, int truncation, uint32_t truncation_coverage, lcr_region **out_regions, uint32_t **out_max_coverage,
                                    uint32_t *out_n, float *out_ms) {
    std::unique_lock<std::recursive_mutex> ctx_lock__;
    if (ctx) ctx_lock__ = std::unique_lock<std::recursive_mutex>(ctx->mu);
    if (!ctx || !in || !out_regions || !out_max_coverage || !out_n) return LCR_ERR_INVALID_ARG;
    if (ctx->sticky) return ctx->sticky;
    *out_regions = nullptr; *out_max_coverage = nullptr; *out_n = 0;
    if (out_ms) *out_ms = 0.f;
    const uint32_t n = in->n_reads, nc = in->n_contigs;
    if (nc && !in->contig_lens) return LCR_ERR_INVALID_ARG;
    if (n && (!in->tid || !in->pos || !in->flag || !in->mapq || !in->de || !in->seq_off || !in->cig_off)) return LCR_ERR_INVALID_ARG;
    if (!n || !nc) return LCR_OK;
    for (uint32_t i = 0; i < n; ++i)
        if (in->cig_off[i + 1] < in->cig_off[i] || in->seq_off[i + 1] < in->seq_off[i]) return LCR_ERR_INVALID_ARG;
    const uint64_t n_cig = in->cig_off[n] - in->cig_off[0];
    if (n_cig && !in->cigar) return LCR_ERR_INVALID_ARG;
    std::vector<uint64_t> base(nc + 1, 0);
    for (uint32_t c = 0; c < nc; ++c) {
        if (in->contig_lens[c] >= 0xfffffffeull) return LCR_ERR_INVALID_ARG; /* Region.start / end are u32 (util.rs:21-32) */
        base[c + 1] = base[c] + in->contig_lens[c] + 1;
    }
    const uint64_t G = base[nc];
    const uint32_t cap = 2 * n + 2; /* a run starts where a read starts or where depth falls back under the truncation level */
    cudaStream_t st = ctx->stream;
    LCR_CUDA_TRY(ctx, cudaSetDevice(ctx->device));

    DevBuf b_tid, b_pos, b_flag, b_mapq, b_de, b_soff, b_coff, b_cig, b_base, b_depth, b_key, b_S, b_E, b_S2, b_E2, b_flagk, b_scan, b_reg, b_push, b_max, b_ctr, b_tmp;
    auto up = [&](DevBuf &b, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(&b.p, bytes ? bytes : 16);
        if (e != cudaSuccess) return e;
        return bytes ? cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess;
    };
    /* offsets are shipped relative to the first element so that a view into a larger pool works */
    std::vector<uint64_t> coff(n + 1);
    for (uint32_t i = 0; i <= n; ++i) coff[i] = in->cig_off[i] - in->cig_off[0];
    LCR_CUDA_TRY(ctx, up(b_tid, in->tid, (size_t)n * 4));
    LCR_CUDA_TRY(ctx, up(b_pos, in->pos, (size_t)n * 4));
    LCR_CUDA_TRY(ctx, up(b_flag, in->flag, (size_t)n * 2));
    LCR_CUDA_TRY(ctx, up(b_mapq, in->mapq, (size_t)n));
    LCR_CUDA_TRY(ctx, up(b_de, in->de, (size_t)n * 4));
    LCR_CUDA_TRY(ctx, up(b_soff, in->seq_off, (size_t)(n + 1) * 8));
    LCR_CUDA_TRY(ctx, up(b_coff, coff.data(), (size_t)(n + 1) * 8));
    LCR_CUDA_TRY(ctx, up(b_cig, in->cigar ? in->cigar + in->cig_off[0] : nullptr, (size_t)n_cig * 4));
    LCR_CUDA_TRY(ctx, up(b_base, base.data(), (size_t)(nc + 1) * 8));
    LCR_CUDA_TRY(ctx, cudaMalloc(&b_depth.p, (G + 4) * 4));
    LCR_CUDA_TRY(ctx, cudaMalloc(&b_key.p, (size_t)n * 8));
    LCR_CUDA_TRY(ctx, cudaMalloc(&b_S.p, (size_t)cap * 8));
    LCR_CUDA_TRY(ctx, cudaMalloc(&b_E.p, (size_t)cap * 8));
    LCR_CUDA_TRY(ctx, cudaMalloc(&b_S2.p, (size_t)cap * 8));
    LCR_CUDA_TRY(ctx, cudaMalloc(&b_E2.p, (size_t)cap * 8));
    LCR_CUDA_TRY(ctx, cudaMalloc(&b_flagk.p, (size_t)cap * 4));
    LCR_CUDA_TRY(ctx, cudaMalloc(&b_scan.p, (size_t)cap * 4));
    LCR_CUDA_TRY(ctx, cudaMalloc(&b_reg.p, (size_t)cap * sizeof(lcr_region)));
    LCR_CUDA_TRY(ctx, cudaMalloc(&b_push.p, (size_t)cap * 8));
    LCR_CUDA_TRY(ctx, cudaMalloc(&b_max.p, (size_t)cap * 4));
    LCR_CUDA_TRY(ctx, cudaMalloc(&b_ctr.p, sizeof(RgCtr)));
    size_t t_scan = 0, t_key = 0, t_sort = 0, t_flag = 0;
    cub::DeviceScan::InclusiveSum(nullptr, t_scan, b_depth.as<int32_t>(), b_depth.as<int32_t>(), (size_t)G, st);
    cub::DeviceScan::InclusiveScan(nullptr, t_key, b_key.as<unsigned long long>(), b_key.as<unsigned long long>(), MaxU64(), (int)n, st);
    cub::DeviceRadixSort::SortKeys(nullptr, t_sort, b_S.as<unsigned long long>(), b_S2.as<unsigned long long>(), (int)cap, 0, 64, st);
    cub::DeviceScan::ExclusiveSum(nullptr, t_flag, b_flagk.as<uint32_t>(), b_scan.as<uint32_t>(), (int)cap, st);
    const size_t t_max = std::max(std::max(t_scan, t_key), std::max(t_sort, t_flag));
    LCR_CUDA_TRY(ctx, cudaMalloc(&b_tmp.p, t_max ? t_max : 16));

    RgIn a;
    a.n_reads = n; a.n_contigs = nc;
    a.tid = b_tid.as<int32_t>(); a.pos = b_pos.as<int32_t>(); a.flag = b_flag.as<uint16_t>(); a.mapq = b_mapq.as<uint8_t>(); a.de = b_de.as<float>();
    a.seq_off = b_soff.as<uint64_t>(); a.cig_off = b_coff.as<uint64_t>(); a.cigar = b_cig.as<uint32_t>(); a.base = b_base.as<uint64_t>();
    a.min_mapq = ctx->P.min_mapq; a.min_read_length = ctx->P.min_read_length; a.divergence = ctx->P.divergence;
    a.truncation = truncation ? 1 : 0; a.trunc_cov = truncation_coverage;

    cudaEvent_t ev0 = ctx->ev_t[0], ev1 = ctx->ev_t[1];
    LCR_CUDA_TRY(ctx, cudaEventRecord(ev0, st));
    LCR_CUDA_TRY(ctx, cudaMemsetAsync(b_depth.p, 0, (G + 4) * 4, st));
    LCR_CUDA_TRY(ctx, cudaMemsetAsync(b_ctr.p, 0, sizeof(RgCtr), st));
    LCR_CUDA_TRY(ctx, cudaMemsetAsync(b_S.p, 0xff, (size_t)cap * 8, st));
    LCR_CUDA_TRY(ctx, cudaMemsetAsync(b_E.p, 0xff, (size_t)cap * 8, st));
    LCR_CUDA_TRY(ctx, cudaMemsetAsync(b_flagk.p, 0, (size_t)cap * 4, st));
    const int sm = ctx->sm_count > 0 ? ctx->sm_count : 148;
    const int rb = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)sm * 8);
    k_rg_span<<<rb, 256, 0, st>>>(a, b_depth.as<int32_t>(), b_key.as<unsigned long long>());
    size_t tb = t_max;
    cub::DeviceScan::InclusiveSum(b_tmp.p, tb, b_depth.as<int32_t>(), b_depth.as<int32_t>(), (size_t)G, st);
    tb = t_max;
    cub::DeviceScan::InclusiveScan(b_tmp.p, tb, b_key.as<unsigned long long>(), b_key.as<unsigned long long>(), MaxU64(), (int)n, st);
    k_rg_runs<<<sm * 8, 256, 0, st>>>(b_depth.as<int32_t>(), G, a.truncation, a.trunc_cov, b_S.as<unsigned long long>(), b_E.as<unsigned long long>(), cap, b_ctr.as<RgCtr>());
    tb = t_max;
    cub::DeviceRadixSort::SortKeys(b_tmp.p, tb, b_S.as<unsigned long long>(), b_S2.as<unsigned long long>(), (int)cap, 0, 64, st);
    tb = t_max;
    cub::DeviceRadixSort::SortKeys(b_tmp.p, tb, b_E.as<unsigned long long>(), b_E2.as<unsigned long long>(), (int)cap, 0, 64, st);
    const int kb = (int)std::min<uint64_t>((cap + 255) / 256, (uint64_t)sm * 8);
    k_rg_group<<<kb, 256, 0, st>>>(b_S2.as<unsigned long long>(), b_E2.as<unsigned long long>(), a.base, nc, b_ctr.as<RgCtr>(), b_flagk.as<uint32_t>());
    tb = t_max;
    cub::DeviceScan::ExclusiveSum(b_tmp.p, tb, b_flagk.as<uint32_t>(), b_scan.as<uint32_t>(), (int)cap, st);
    k_rg_emit<<<kb, 256, 0, st>>>(b_S2.as<unsigned long long>(), b_E2.as<unsigned long long>(), a.base, nc, b_ctr.as<RgCtr>(), b_flagk.as<uint32_t>(), b_scan.as<uint32_t>(), b_reg.as<lcr_region>(),
                                  b_push.as<unsigned long long>(), b_max.as<uint32_t>());
    k_rg_maxcov<<<sm * 8, 256, 0, st>>>(b_depth.as<int32_t>(), G, a.base, nc, b_ctr.as<RgCtr>(), b_reg.as<lcr_region>(), b_push.as<unsigned long long>(), b_max.as<uint32_t>());
    k_rg_reads<<<kb, 256, 0, st>>>(b_ctr.as<RgCtr>(), b_reg.as<lcr_region>(), a.tid, a.pos, b_key.as<unsigned long long>(), n);
    LCR_CUDA_TRY(ctx, cudaEventRecord(ev1, st));
    RgCtr h;
    LCR_CUDA_TRY(ctx, cudaMemcpyAsync(&h, b_ctr.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    LCR_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    LCR_CUDA_TRY(ctx, cudaGetLastError());
    if (h.n_start != h.n_end || h.n_start > cap) {
        ctx->last_error = "lcr_discover_regions: run start / end lists disagree";
        return LCR_ERR_INTERNAL;
    }
    if (out_ms) LCR_CUDA_TRY(ctx, cudaEventElapsedTime(out_ms, ev0, ev1));
    if (h.n_regions) {
        lcr_region *r = (lcr_region *)malloc((size_t)h.n_regions * sizeof(lcr_region));
        uint32_t *m = (uint32_t *)malloc((size_t)h.n_regions * 4);
        if (!r || !m) { free(r); free(m); return LCR_ERR_OOM; }
        cudaError_t e = cudaMemcpyAsync(r, b_reg.p, (size_t)h.n_regions * sizeof(lcr_region), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(m, b_max.p, (size_t)h.n_regions * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { free(r); free(m); LCR_CUDA_TRY(ctx, e); }
        *out_regions = r; *out_max_coverage = m; *out_n = h.n_regions;
    }
    return LCR_OK;
}

extern "C" void lcr_free_regions(lcr_region *regions, uint32_t *max_coverage) {
    free(regions);
    free(max_coverage);
}
