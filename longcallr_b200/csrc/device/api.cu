/*
 * api.cu — the C ABI of include/longcallr_b200.h: contexts, reference upload, batch upload,
 * the device run (pileup/genotype stage, fragment + phasing stage) and result fetch.
 *
 * The call sequence mirrors the reference worker (src/thread.rs:78-221): lcr_set_reference is
 * `ref_seqs.get(chr)` (:59,:79), lcr_submit is the closure body, and the returned records are
 * what the three queues receive (:204-221).  There is no CPU fallback: without a device every
 * compute entry point returns LCR_ERR_NO_DEVICE.
 */
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "lcr_frag.h"

int lcr_stage_pileup_impl(lcr_ctx *ctx, lcr_device_batch *db, uint8_t *slot_flags);

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

/* fragment count from which an LD-path region gets the cooperative whole-GPU kernel (LCR_BIG_REGION_FRAGS overrides: tests) */
static uint32_t big_frag_threshold() {
    const char *e = getenv("LCR_BIG_REGION_FRAGS");
    if (e && *e) return (uint32_t)strtoul(e, nullptr, 10);
    return 8192;
}

__global__ void k_scatter_winners(uint32_t n, const uint32_t *slot, const long long *prob, const uint32_t *cfg, long long *out_prob, uint32_t *out_cfg) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out_prob[slot[i]] = prob[i]; out_cfg[slot[i]] = cfg[i]; }
}

__global__ void k_init_pairs(LcrPairEntry *t, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { t[i].key = ~0ull; t[i].cis = 0; t[i].trans = 0; }
}

int exclusive_scan_u32(lcr_ctx *ctx, const uint32_t *in, uint32_t *out, size_t n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, ctx->stream);
    void *tmp = nullptr;
    TRY(cudaMallocAsync(&tmp, bytes ? bytes : 16, ctx->stream));
    TRY(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, (int)n, ctx->stream));
    TRY(cudaFreeAsync(tmp, ctx->stream));
    return 0;
}

} // namespace

struct lcr_device_batch_full : lcr_device_batch {
    DbExtra extra;
    uint8_t *slot_flags = nullptr;
};

static int stage_fragments_phase(lcr_ctx *ctx, lcr_device_batch_full *db) {
    cudaStream_t st = ctx->stream;
    const uint32_t n_slots = db->n_slots, n_cand = db->n_cand, n_regions = db->n_regions;
    uint32_t *frag_flag = nullptr, *elem_count = nullptr, *frag_scan = nullptr, *elem_scan = nullptr;
    uint32_t *cover_count = nullptr, *cover_off = nullptr, *cover_cursor = nullptr, *cover_frag = nullptr;
    int8_t *cover_cell = nullptr;
    uint32_t *frag_slot = nullptr, *frag_elem_off = nullptr, *frag_links = nullptr, *elem_snp = nullptr;
    int8_t *elem_cell = nullptr;
    uint8_t *elem_base = nullptr;
    DALLOC(frag_flag, (size_t)n_slots + 1);
    DALLOC(elem_count, (size_t)n_slots + 1);
    DALLOC(frag_scan, (size_t)n_slots + 1);
    DALLOC(elem_scan, (size_t)n_slots + 1);
    DALLOC(cover_count, (size_t)n_cand + 1);
    DALLOC(cover_off, (size_t)n_cand + 1);
    DALLOC(cover_cursor, (size_t)n_cand + 1);
    TRY(cudaMemsetAsync(frag_flag, 0, sizeof(uint32_t) * ((size_t)n_slots + 1), st));
    TRY(cudaMemsetAsync(elem_count, 0, sizeof(uint32_t) * ((size_t)n_slots + 1), st));
    TRY(cudaMemsetAsync(cover_count, 0, sizeof(uint32_t) * ((size_t)n_cand + 1), st));
    TRY(cudaMemsetAsync(cover_cursor, 0, sizeof(uint32_t) * ((size_t)n_cand + 1), st));

    FragArgs fa{};
    fa.P = ctx->P;
    fa.n_slots = n_slots;
    fa.regions = db->regions;
    fa.slot_off = db->slot_off; fa.slot_region = db->slot_region; fa.slot_flags = db->slot_flags;
    fa.pos = db->pos; fa.seq_off = db->seq_off; fa.cig_off = db->cig_off; fa.seq = db->seq; fa.qual = db->qual; fa.cigar = db->cigar;
    fa.rstate = db->rstate; fa.cand = db->cand; fa.stats = db->d_stats;
    fa.frag_flag = frag_flag; fa.elem_count = elem_count; fa.frag_scan = frag_scan; fa.elem_scan = elem_scan;
    fa.cover_count = cover_count; fa.cover_off = cover_off; fa.cover_cursor = cover_cursor;
    fa.is_fragment = db->is_fragment;
    static const int walk_mode = [] { const char *e = getenv("LCR_FRAG_WALK"); return e && *e ? atoi(e) : 0; }(); /* 1: thread per read, 2: warp per read */
    const bool long_cigars = walk_mode == 2 || (walk_mode == 0 && db->n_cigar > 24ull * (db->n_reads ? db->n_reads : 1));
    lcr_launch_frag_count(fa, long_cigars, st);
    int rc;
    if ((rc = exclusive_scan_u32(ctx, frag_flag, frag_scan, (size_t)n_slots + 1))) return rc;
    if ((rc = exclusive_scan_u32(ctx, elem_count, elem_scan, (size_t)n_slots + 1))) return rc;
    if ((rc = exclusive_scan_u32(ctx, cover_count, cover_off, (size_t)n_cand + 1))) return rc;
    db->timing.kernel_launches += 1; /* own kernels only; cub scans are library code */
    lcr_launch_region_frag_ranges(n_regions, db->slot_off, frag_scan, db->rstate, st);
    db->timing.kernel_launches += 1;
    uint32_t n_frag_total = 0, n_elem_total = 0;
    std::vector<LcrRegionState> hrs(n_regions);
    TRY(cudaMemcpyAsync(&n_frag_total, frag_scan + n_slots, 4, cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(&n_elem_total, elem_scan + n_slots, 4, cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(hrs.data(), db->rstate, sizeof(LcrRegionState) * n_regions, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
    /* LD pair tables: one open-addressing segment per region that takes the LD path */
    uint64_t table_size = 0;
    for (uint32_t r = 0; r < n_regions; ++r) {
        LcrRegionState &s = hrs[r];
        s.pair_begin = 0; s.pair_cap = 0;
        if (s.status != 0 || s.n_cand <= ctx->P.max_enum_snps || !s.n_ld_pairs_cap) continue;
        uint64_t bound = std::min<uint64_t>(s.n_ld_pairs_cap, (uint64_t)s.n_cand * (s.n_cand - 1) / 2);
        uint64_t cap = 16;
        while (cap < 2 * bound) cap <<= 1;
        if (table_size + cap > 0xfffffff0ull) { ctx->last_error = "LD pair table too large"; return LCR_ERR_OOM; }
        s.pair_begin = (uint32_t)table_size;
        s.pair_cap = (uint32_t)cap;
        table_size += cap;
    }
    TRY(cudaMemcpyAsync(db->rstate, hrs.data(), sizeof(LcrRegionState) * n_regions, cudaMemcpyHostToDevice, st));
    db->n_frag = n_frag_total;
    db->n_elem = n_elem_total;
    DALLOC(frag_slot, (size_t)n_frag_total + 1);
    DALLOC(frag_elem_off, (size_t)n_frag_total + 1);
    DALLOC(frag_links, (size_t)n_frag_total + 1);
    DALLOC(elem_snp, n_elem_total);
    DALLOC(elem_cell, n_elem_total);
    DALLOC(elem_base, n_elem_total);
    DALLOC(cover_frag, n_elem_total);
    DALLOC(cover_cell, n_elem_total);
    if (!n_frag_total) TRY(cudaMemsetAsync(frag_elem_off, 0, sizeof(uint32_t), st));
    fa.n_frag_total = n_frag_total; fa.n_elem_total = n_elem_total;
    fa.frag_slot = frag_slot; fa.frag_elem_off = frag_elem_off; fa.frag_links = frag_links;
    fa.elem_snp = elem_snp; fa.elem_cell = elem_cell; fa.elem_base = elem_base;
    fa.cover_frag = cover_frag; fa.cover_cell = cover_cell;
    lcr_launch_frag_fill(fa, long_cigars, st);
    db->timing.kernel_launches += 1;

    /* LD graph */
    LcrPairEntry *table = nullptr;
    uint32_t *entry_region = nullptr, *deg = nullptr, *adj_off = nullptr, *adj_cursor = nullptr, *adj = nullptr;
    DALLOC(deg, (size_t)n_cand + 1);
    DALLOC(adj_off, (size_t)n_cand + 1);
    DALLOC(adj_cursor, (size_t)n_cand + 1);
    TRY(cudaMemsetAsync(deg, 0, sizeof(uint32_t) * ((size_t)n_cand + 1), st));
    TRY(cudaMemsetAsync(adj_cursor, 0, sizeof(uint32_t) * ((size_t)n_cand + 1), st));
    uint32_t adj_total = 0;
    if (table_size) {
        DALLOC(table, table_size);
        DALLOC(entry_region, table_size / 16);
        k_init_pairs<<<(uint32_t)((table_size + 255) / 256), 256, 0, st>>>(table, table_size);
        lcr_launch_fill_entry_region(n_regions, db->rstate, entry_region, st);
        lcr_launch_pair_count(fa, table, st);
        lcr_launch_ld_edges(false, ctx->P.ld_weight_threshold, n_regions, db->rstate, table, table_size, entry_region, deg, nullptr, nullptr, nullptr, st);
        db->timing.kernel_launches += 4;
    }
    if ((rc = exclusive_scan_u32(ctx, deg, adj_off, (size_t)n_cand + 1))) return rc;
    if (table_size) {
        TRY(cudaMemcpyAsync(&adj_total, adj_off + n_cand, 4, cudaMemcpyDeviceToHost, st));
        TRY(cudaStreamSynchronize(st));
    }
    DALLOC(adj, adj_total);
    if (table_size && adj_total) {
        lcr_launch_ld_edges(true, ctx->P.ld_weight_threshold, n_regions, db->rstate, table, table_size, entry_region, deg, adj_off, adj_cursor, adj, st);
        lcr_launch_adj_sort(n_cand, adj_off, adj, st);
        db->timing.kernel_launches += 2;
    }
    cudaEvent_t ev_frag;
    TRY(cudaEventCreate(&ev_frag));
    TRY(cudaEventRecord(ev_frag, st));

    /* phasing state */
    PhaseArgs pa{};
    pa.P = ctx->P;
    pa.n_regions = n_regions;
    pa.regions = db->regions; pa.slot_off = db->slot_off; pa.rstate = db->rstate; pa.cand = db->cand;
    pa.tables = ctx->d_tables; pa.stats = db->d_stats;
    pa.frag_slot = frag_slot; pa.frag_elem_off = frag_elem_off; pa.frag_links = frag_links;
    pa.elem_snp = elem_snp; pa.elem_cell = elem_cell;
    pa.cover_off = cover_off; pa.cover_frag = cover_frag; pa.cover_cell = cover_cell;
    pa.adj_off = adj_off; pa.adj = adj;
    DALLOC(pa.st, n_cand); DALLOC(pa.best_hap, n_cand); DALLOC(pa.best_gen, n_cand);
    DALLOC(pa.label, n_cand); DALLOC(pa.rank, n_cand);
    DALLOC(pa.work, (size_t)adj_total + n_cand + 1);
    DALLOC(pa.blk_q, n_cand); DALLOC(pa.blk_qflip, n_cand);
    DALLOC(pa.tag, n_frag_total); DALLOC(pa.best_tag, n_frag_total); DALLOC(pa.fp, n_frag_total); DALLOC(pa.assign, n_frag_total);
    pa.hp = db->hp; pa.ps = db->ps;
    /* enumeration search (regions with at most min(max_enum_snps, 10) candidates): one warp per configuration */
    uint32_t *es_base = nullptr, *es_cfg = nullptr, *work_region = nullptr, *work_chunk = nullptr;
    long long *es_prob = nullptr;
    {
        /* bins: launch shape (by number of configurations) x fragment-count class (shared memory footprint) */
        const uint32_t NF_TINY = 384, NF_SMALL = 1024, NF_BIG = 16384; /* fragment-count classes: shared-memory footprint */
        const int NBIN = 15;
        std::vector<uint32_t> base(n_regions + 1, 0), wr[NBIN], wc[NBIN];
        uint32_t nfmax[NBIN] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        /* regions with 5+ sites: 64 configurations per CTA when that still fills the GPU, else 16 (more, shorter CTAs) */
        uint64_t big_cfgs = 0;
        for (uint32_t r = 0; r < n_regions; ++r) {
            const LcrRegionState &s = hrs[r];
            if (s.status == 0 && s.n_cand > 4 && s.n_cand <= ctx->P.max_enum_snps && s.n_cand <= 10 && s.n_frag <= NF_BIG) big_cfgs += 1ull << s.n_cand;
        }
        const bool small_batch = big_cfgs / 64 < 8ull * (uint64_t)(ctx->sm_count > 0 ? ctx->sm_count : 148);
        std::vector<int> bin(n_regions, -1);
        for (uint32_t r = 0; r < n_regions; ++r) {
            const LcrRegionState &s = hrs[r];
            uint32_t chunks = 0;
            if (s.status == 0 && s.n_cand && s.n_cand <= ctx->P.max_enum_snps && s.n_cand <= 10 && s.n_frag <= NF_BIG) {
                int shape = lcr_enum_shape_for(s.n_cand);
                if (shape == 3 && small_batch) shape = 4;
                const uint32_t per_cta = lcr_enum_cfgs_per_cta(shape);
                chunks = ((1u << s.n_cand) + per_cta - 1) / per_cta;
                bin[r] = shape * 3 + (s.n_frag <= NF_TINY ? 0 : s.n_frag <= NF_SMALL ? 1 : 2);
            }
            base[r + 1] = base[r] + chunks;
        }
        const uint32_t n_work_total = base[n_regions];
        if (n_work_total) {
            for (uint32_t r = 0; r < n_regions; ++r)
                if (bin[r] >= 0) {
                    nfmax[bin[r]] = std::max(nfmax[bin[r]], hrs[r].n_frag);
                    for (uint32_t ck = 0; ck < base[r + 1] - base[r]; ++ck) { wr[bin[r]].push_back(r); wc[bin[r]].push_back(ck); }
                }
            DALLOC(es_base, (size_t)n_regions + 1);
            DALLOC(es_cfg, n_work_total);
            DALLOC(es_prob, n_work_total);
            DALLOC(work_region, n_work_total);
            DALLOC(work_chunk, n_work_total);
            TRY(cudaMemcpyAsync(es_base, base.data(), sizeof(uint32_t) * (n_regions + 1), cudaMemcpyHostToDevice, st));
            /* each bin writes its winners in launch order; they are scattered to (region, chunk) order afterwards */
            std::vector<uint32_t> order_region, order_chunk;
            for (int b = 0; b < NBIN; ++b) { order_region.insert(order_region.end(), wr[b].begin(), wr[b].end()); order_chunk.insert(order_chunk.end(), wc[b].begin(), wc[b].end()); }
            TRY(cudaMemcpyAsync(work_region, order_region.data(), sizeof(uint32_t) * n_work_total, cudaMemcpyHostToDevice, st));
            TRY(cudaMemcpyAsync(work_chunk, order_chunk.data(), sizeof(uint32_t) * n_work_total, cudaMemcpyHostToDevice, st));
            long long *tmp_prob = nullptr;
            uint32_t *tmp_cfg = nullptr, *d_slot = nullptr;
            DALLOC(tmp_prob, n_work_total);
            DALLOC(tmp_cfg, n_work_total);
            DALLOC(d_slot, n_work_total);
            std::vector<uint32_t> slot(n_work_total);
            for (uint32_t i = 0; i < n_work_total; ++i) slot[i] = base[order_region[i]] + order_chunk[i];
            TRY(cudaMemcpyAsync(d_slot, slot.data(), sizeof(uint32_t) * n_work_total, cudaMemcpyHostToDevice, st));
            TRY(cudaStreamSynchronize(st)); /* the host vectors above go out of scope */
            /* the bins are independent: fork them onto side streams so that small and large shapes overlap */
            uint32_t off = 0;
            TRY(cudaEventRecord(ctx->ev_fork, st));
            int used = 0;
            for (int b = 0; b < NBIN; ++b) {
                const uint32_t nw = (uint32_t)wr[b].size();
                if (!nw) continue;
                cudaStream_t ss = ctx->side[used % 4];
                TRY(cudaStreamWaitEvent(ss, ctx->ev_fork, 0));
                int e = lcr_launch_enum_search(b / 3, b % 3 == 0, pa, nw, work_region + off, work_chunk + off, std::max<uint32_t>(nfmax[b], 32), tmp_prob + off, tmp_cfg + off, ss);
                if (e) { ctx->last_error = "k_enum_search launch failed"; ctx->sticky = LCR_ERR_CUDA; return ctx->sticky; }
                db->timing.kernel_launches += 1;
                off += nw;
                ++used;
            }
            for (int i = 0; i < 4 && i < used; ++i) {
                TRY(cudaEventRecord(ctx->ev_join[i], ctx->side[i]));
                TRY(cudaStreamWaitEvent(st, ctx->ev_join[i], 0));
            }
            k_scatter_winners<<<(n_work_total + 127) / 128, 128, 0, st>>>(n_work_total, d_slot, tmp_prob, tmp_cfg, es_prob, es_cfg);
            db->timing.kernel_launches += 1;
            DFREE(tmp_prob); DFREE(tmp_cfg); DFREE(d_slot);
            pa.es_base = es_base; pa.es_prob = es_prob; pa.es_cfg = es_cfg;
        }
    }
    /* regions too large for one CTA take the whole GPU, one after the other */
    uint8_t *big_region = nullptr;
    void *bcast = nullptr;
    std::vector<uint32_t> big_list;
    {
        std::vector<uint8_t> big(n_regions, 0);
        for (uint32_t r = 0; r < n_regions; ++r)
            if (hrs[r].status == 0 && hrs[r].n_cand > ctx->P.max_enum_snps && hrs[r].n_frag >= big_frag_threshold()) { big[r] = 1; big_list.push_back(r); }
        if (!big_list.empty()) {
            DALLOC(big_region, n_regions);
            TRY(cudaMemcpyAsync(big_region, big.data(), n_regions, cudaMemcpyHostToDevice, st));
            TRY(cudaStreamSynchronize(st));
            TRY(cudaMallocAsync(&bcast, lcr_phase_bcast_bytes(), st));
            pa.big_region = big_region;
        }
    }
    lcr_launch_phase(pa, st);
    db->timing.kernel_launches += 1;
    for (uint32_t r : big_list) {
        int e = lcr_launch_phase_grid(pa, r, bcast, ctx->sm_count, st);
        if (e) { ctx->last_error = std::string("k_phase_grid: ") + cudaGetErrorString((cudaError_t)e); ctx->sticky = LCR_ERR_CUDA; return ctx->sticky; }
        db->timing.kernel_launches += 1;
    }
    cudaEvent_t ev_end;
    TRY(cudaEventCreate(&ev_end));
    TRY(cudaEventRecord(ev_end, st));

    if (ctx->P.flags & LCR_FLAG_EMIT_FRAGMENTS) {
        FragDebug &fd = db->extra.fragdbg;
        fd.frag_slot.resize(n_frag_total); fd.frag_elem_off.resize((size_t)n_frag_total + 1);
        fd.elem_snp.resize(n_elem_total); fd.elem_cell.resize(n_elem_total); fd.elem_base.resize(n_elem_total);
        if (n_frag_total) TRY(cudaMemcpyAsync(fd.frag_slot.data(), frag_slot, 4ull * n_frag_total, cudaMemcpyDeviceToHost, st));
        TRY(cudaMemcpyAsync(fd.frag_elem_off.data(), frag_elem_off, 4ull * ((size_t)n_frag_total + (n_frag_total ? 1 : 0)), cudaMemcpyDeviceToHost, st));
        if (n_elem_total) {
            TRY(cudaMemcpyAsync(fd.elem_snp.data(), elem_snp, 4ull * n_elem_total, cudaMemcpyDeviceToHost, st));
            TRY(cudaMemcpyAsync(fd.elem_cell.data(), elem_cell, n_elem_total, cudaMemcpyDeviceToHost, st));
            TRY(cudaMemcpyAsync(fd.elem_base.data(), elem_base, n_elem_total, cudaMemcpyDeviceToHost, st));
        }
        if (!n_frag_total) fd.frag_elem_off[0] = 0;
    }
    TRY(cudaStreamSynchronize(st));
    TRY(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_frag, ev_end);
    db->timing.ms_phase = ms;
    cudaEventDestroy(ev_frag);
    cudaEventDestroy(ev_end);

    DFREE(frag_flag); DFREE(elem_count); DFREE(frag_scan); DFREE(elem_scan);
    DFREE(cover_count); DFREE(cover_off); DFREE(cover_cursor); DFREE(cover_frag); DFREE(cover_cell);
    DFREE(frag_slot); DFREE(frag_elem_off); DFREE(frag_links); DFREE(elem_snp); DFREE(elem_cell); DFREE(elem_base);
    DFREE(table); DFREE(entry_region); DFREE(deg); DFREE(adj_off); DFREE(adj_cursor); DFREE(adj);
    DFREE(pa.st); DFREE(pa.best_hap); DFREE(pa.best_gen);
    DFREE(pa.label); DFREE(pa.rank); DFREE(pa.work); DFREE(pa.blk_q); DFREE(pa.blk_qflip);
    DFREE(pa.tag); DFREE(pa.best_tag); DFREE(pa.fp); DFREE(pa.assign);
    DFREE(es_base); DFREE(es_cfg); DFREE(es_prob); DFREE(work_region); DFREE(work_chunk);
    DFREE(big_region); DFREE(bcast);
    return LCR_OK;
}

extern "C" {

int lcr_abi_version(void) { return LCR_ABI_VERSION; }

const char *lcr_strerror(int status) {
    switch (status) {
        case LCR_OK: return "ok";
        case LCR_ERR_INVALID_ARG: return "invalid argument";
        case LCR_ERR_CUDA: return "CUDA error";
        case LCR_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback on this path)";
        case LCR_ERR_OOM: return "out of memory";
        case LCR_ERR_BAD_CIGAR: return "unknown or inconsistent CIGAR operation";
        case LCR_ERR_NO_REFERENCE: return "region on a contig without reference sequence";
        case LCR_ERR_BASEQ_ZERO: return "base quality 0 at a fragment site";
        default: return "unknown status";
    }
}

const char *lcr_last_error(lcr_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int lcr_create(const lcr_params *p, int device, lcr_ctx **out) {
    if (!p || !out) return LCR_ERR_INVALID_ARG;
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
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return LCR_ERR_CUDA; }
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    memset(&ctx->last_submit, 0, sizeof ctx->last_submit);
    for (int i = 0; i < 4; ++i) {
        cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming);
    }
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
    if (cudaMalloc(&ctx->d_tables, sizeof t) != cudaSuccess || cudaMemcpy(ctx->d_tables, &t, sizeof t, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
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
    for (int i = 0; i < 4; ++i) { cudaStreamDestroy(ctx->side[i]); cudaEventDestroy(ctx->ev_join[i]); }
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
    TRY(cudaMalloc(&ctx->d_ref[tid], len ? len : 1));
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
    if (b->n_reads && (!b->pos || !b->flag || !b->mapq || !b->ts || !b->de || !b->seq_off || !b->cig_off)) return LCR_ERR_INVALID_ARG;
    TRY(cudaSetDevice(ctx->device));
    lcr_device_batch_full *db = new (std::nothrow) lcr_device_batch_full();
    if (!db) return LCR_ERR_OOM;
    db->ev_meta = nullptr; db->ev_seq = nullptr; db->seq_wait_pending = false;
    db->n_regions = b->n_regions;
    db->n_reads = b->n_reads;
    db->ran = false;
    memset(&db->timing, 0, sizeof db->timing);
    /* prefix sums over region lengths and read ranges; regions that cannot run get their status here */
    std::vector<uint32_t> slot_off(b->n_regions + 1, 0), tile_base(b->n_regions + 1, 0);
    std::vector<uint64_t> pos_off(b->n_regions + 1, 0);
    db->extra.h_status0.assign(b->n_regions, 0);
    for (uint32_t r = 0; r < b->n_regions; ++r) {
        const lcr_region &g = b->regions[r];
        int32_t stt = 0;
        if (g.end < g.start || g.start < 1 || g.read_end < g.read_begin || g.read_end > b->n_reads) stt = LCR_ERR_INVALID_ARG;
        else if (g.tid < 0 || (size_t)g.tid >= ctx->d_ref.size() || !ctx->d_ref[g.tid]) stt = LCR_ERR_NO_REFERENCE;
        else if ((uint64_t)g.end - 1 > ctx->ref_len[g.tid]) stt = LCR_ERR_INVALID_ARG;
        db->extra.h_status0[r] = stt;
        const uint64_t len = stt ? 0 : (uint64_t)(g.end - g.start);
        const uint32_t nreads = stt ? 0 : g.read_end - g.read_begin;
        slot_off[r + 1] = slot_off[r] + nreads;
        tile_base[r + 1] = tile_base[r] + (uint32_t)((len + LCR_TILE - 1) / LCR_TILE);
        pos_off[r + 1] = pos_off[r] + len;
    }
    db->n_slots = slot_off[b->n_regions];
    db->n_tiles = tile_base[b->n_regions];
    db->n_pos = pos_off[b->n_regions];
    std::vector<uint32_t> slot_region(db->n_slots), tile_region(db->n_tiles);
    for (uint32_t r = 0; r < b->n_regions; ++r) {
        std::fill(slot_region.begin() + slot_off[r], slot_region.begin() + slot_off[r + 1], r);
        std::fill(tile_region.begin() + tile_base[r], tile_region.begin() + tile_base[r + 1], r);
    }
    db->h_pos_off = pos_off;
    db->h_slot_off = slot_off;
    db->extra.h_regions.assign(b->regions, b->regions + b->n_regions);
    const uint64_t n_bases = b->n_reads ? b->seq_off[b->n_reads] : 0, n_cig = b->n_reads ? b->cig_off[b->n_reads] : 0;
    db->n_bases = n_bases;
    db->n_cigar = n_cig;
    uint64_t bytes = 0;
    int rc = 0;
#define UP(field, src, n) if (!rc) rc = h2d(ctx, &db->field, src, (size_t)(n), &bytes)
    /* the small host-side tables first (pageable sources: those copies wait for the stream), the caller's large arrays last,
       so that an asynchronous upload returns while seq / qual are still in flight */
    UP(slot_off, slot_off.data(), slot_off.size());
    UP(slot_region, slot_region.data(), slot_region.size());
    UP(tile_base, tile_base.data(), tile_base.size());
    UP(tile_region, tile_region.data(), tile_region.size());
    UP(pos_off, pos_off.data(), pos_off.size());
    UP(regions, b->regions, b->n_regions);
    if (b->n_reads) { UP(seq_off, b->seq_off, (size_t)b->n_reads + 1); UP(cig_off, b->cig_off, (size_t)b->n_reads + 1); }
    else { static const uint64_t zero = 0; UP(seq_off, &zero, 1); UP(cig_off, &zero, 1); }
    UP(pos, b->pos, b->n_reads);
    UP(flag, b->flag, b->n_reads);
    UP(mapq, b->mapq, b->n_reads);
    UP(ts, b->ts, b->n_reads);
    UP(de, b->de, b->n_reads);
    UP(cigar, b->cigar, n_cig);
    if (!rc && async) {
        cudaError_t e = cudaEventCreateWithFlags(&db->ev_meta, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(db->ev_meta, ctx->stream);
        if (e != cudaSuccess) { ctx->last_error = cudaGetErrorString(e); ctx->sticky = LCR_ERR_CUDA; rc = ctx->sticky; }
    }
    /* seq / qual carry 32 bytes of slack: the tile kernel reads aligned 16-byte blocks plus the following word */
    if (!rc) rc = h2d_padded(ctx, &db->seq, b->seq, (size_t)n_bases, 32, &bytes);
    if (!rc) rc = h2d_padded(ctx, &db->qual, b->qual, (size_t)n_bases, 32, &bytes);
#undef UP
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

static void free_results(lcr_ctx *ctx, lcr_device_batch_full *db) {
    DFREE(db->rstate); DFREE(db->cand); DFREE(db->hp); DFREE(db->ps); DFREE(db->is_fragment); DFREE(db->d_stats);
    DFREE(db->pl_acgt); DFREE(db->pl_fwd); DFREE(db->pl_d); DFREE(db->pl_n); DFREE(db->pl_ts);
    DFREE(db->slot_flags);
    db->n_cand = 0;
    db->ran = false;
}

int lcr_run_device(lcr_ctx *ctx, lcr_device_batch *dbb) {
    std::unique_lock<std::recursive_mutex> ctx_lock__;
    if (ctx) ctx_lock__ = std::unique_lock<std::recursive_mutex>(ctx->mu);
    if (!ctx || !dbb) return LCR_ERR_INVALID_ARG;
    if (ctx->sticky) return ctx->sticky;
    lcr_device_batch_full *db = static_cast<lcr_device_batch_full *>(dbb);
    TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    db->seq_wait_pending = false;
    if (db->ev_meta) TRY(cudaStreamWaitEvent(st, db->ev_meta, 0));
    if (db->ev_seq) {
        /* ONT presets: nothing before the tile kernel reads bases (the read-end trim needs no sequence), so the seq / qual
           copies keep running under the read filter and segment build; the pileup stage waits right before the tile kernel */
        if (ctx->P.platform == 1) db->seq_wait_pending = true;
        else TRY(cudaStreamWaitEvent(st, db->ev_seq, 0));
    }
    free_results(ctx, db);
    const uint64_t h2d_keep = db->h2d_bytes;
    memset(&db->timing, 0, sizeof db->timing);
    db->timing.h2d_bytes = h2d_keep;
    if (ctx->ref_dirty) {
        const int n = (int)ctx->d_ref.size();
        if (ctx->d_ref_table) { TRY(cudaFree(ctx->d_ref_table)); ctx->d_ref_table = nullptr; }
        TRY(cudaMalloc(&ctx->d_ref_table, sizeof(uint8_t *) * (n ? n : 1)));
        if (n) TRY(cudaMemcpy(ctx->d_ref_table, ctx->d_ref.data(), sizeof(uint8_t *) * n, cudaMemcpyHostToDevice));
        ctx->ref_dirty = false;
    }
    cudaEvent_t ev0, ev1, ev2;
    TRY(cudaEventCreate(&ev0)); TRY(cudaEventCreate(&ev1)); TRY(cudaEventCreate(&ev2));
    TRY(cudaEventRecord(ev0, st));
    /* result buffers */
    DALLOC(db->rstate, db->n_regions);
    {
        std::vector<LcrRegionState> init(db->n_regions);
        for (uint32_t r = 0; r < db->n_regions; ++r) { memset(&init[r], 0, sizeof init[r]); init[r].status = db->extra.h_status0[r]; }
        if (db->n_regions) TRY(cudaMemcpyAsync(db->rstate, init.data(), sizeof(LcrRegionState) * db->n_regions, cudaMemcpyHostToDevice, st));
        TRY(cudaStreamSynchronize(st));
    }
    DALLOC(db->hp, db->n_reads); DALLOC(db->ps, db->n_reads); DALLOC(db->is_fragment, db->n_reads);
    DALLOC(db->d_stats, 1);
    DALLOC(db->slot_flags, db->n_slots);
    TRY(cudaMemsetAsync(db->hp, 0xff, db->n_reads ? db->n_reads : 1, st));
    TRY(cudaMemsetAsync(db->ps, 0, sizeof(uint32_t) * (db->n_reads ? db->n_reads : 1), st));
    TRY(cudaMemsetAsync(db->is_fragment, 0, db->n_reads ? db->n_reads : 1, st));
    TRY(cudaMemsetAsync(db->d_stats, 0, sizeof(lcr_stats), st));
    TRY(cudaMemsetAsync(db->slot_flags, 0, db->n_slots ? db->n_slots : 1, st));
    if (ctx->P.flags & LCR_FLAG_EMIT_PLANES) {
        DALLOC(db->pl_acgt, db->n_pos * 4); DALLOC(db->pl_fwd, db->n_pos * 4); DALLOC(db->pl_d, db->n_pos); DALLOC(db->pl_n, db->n_pos); DALLOC(db->pl_ts, db->n_pos * 2);
        const size_t np1 = db->n_pos ? db->n_pos : 1; /* regions that fail on the device keep zeroed planes */
        TRY(cudaMemsetAsync(db->pl_acgt, 0, 16 * np1 / (db->n_pos ? 1 : 4), st)); TRY(cudaMemsetAsync(db->pl_fwd, 0, 16 * np1 / (db->n_pos ? 1 : 4), st));
        TRY(cudaMemsetAsync(db->pl_d, 0, 4 * np1, st)); TRY(cudaMemsetAsync(db->pl_n, 0, 4 * np1, st)); TRY(cudaMemsetAsync(db->pl_ts, 0, 8 * np1 / (db->n_pos ? 1 : 2), st));
    }
    int rc = lcr_stage_pileup_impl(ctx, db, db->slot_flags);
    if (rc) return rc;
    TRY(cudaEventRecord(ev1, st));
    if (!(ctx->P.flags & LCR_FLAG_SKIP_PHASING)) {
        rc = stage_fragments_phase(ctx, db);
        if (rc) return rc;
    }
    TRY(cudaEventRecord(ev2, st));
    TRY(cudaStreamSynchronize(st));
    TRY(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1); db->timing.ms_pileup = ms;
    cudaEventElapsedTime(&ms, ev1, ev2); db->timing.ms_fragments = ms - db->timing.ms_phase;
    cudaEventElapsedTime(&ms, ev0, ev2); db->timing.ms_total = ms;
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
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
    free_results(ctx, db);
    DFREE(db->regions); DFREE(db->pos); DFREE(db->flag); DFREE(db->mapq); DFREE(db->ts); DFREE(db->de);
    DFREE(db->seq_off); DFREE(db->cig_off); DFREE(db->seq); DFREE(db->qual); DFREE(db->cigar);
    DFREE(db->slot_off); DFREE(db->slot_region); DFREE(db->tile_base); DFREE(db->tile_region); DFREE(db->pos_off);
    cudaStreamSynchronize(ctx->stream);
    if (db->ev_seq) { cudaEventSynchronize(db->ev_seq); cudaEventDestroy(db->ev_seq); }
    if (db->ev_meta) cudaEventDestroy(db->ev_meta);
    delete db;
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

static size_t submit_chunk_bytes() { /* seq + qual bytes per chunk (LCR_SUBMIT_CHUNK_MB overrides: tests) */
    const char *e = getenv("LCR_SUBMIT_CHUNK_MB");
    if (e && *e) return (size_t)strtoull(e, nullptr, 10) << 20;
    return (size_t)256 << 20;
}

/* The worker body for a batch of regions, host buffers in, host results out.  Large batches are cut into chunks of
   consecutive regions; the host-to-device copies of chunk k+1 run on the copy stream while chunk k computes (regions are
   independent, so the result is the same as one pass over the whole batch). */
int lcr_submit(lcr_ctx *ctx, const lcr_batch *batch, lcr_result **out) {
    std::unique_lock<std::recursive_mutex> ctx_lock__;
    if (ctx) ctx_lock__ = std::unique_lock<std::recursive_mutex>(ctx->mu);
    if (!ctx || !batch || !out) return LCR_ERR_INVALID_ARG;
    if (ctx->sticky) return ctx->sticky;
    memset(&ctx->last_submit, 0, sizeof ctx->last_submit);
    const size_t chunk_bytes = submit_chunk_bytes();
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
            while (r < batch->n_regions) {
                const lcr_region &g = batch->regions[r];
                const bool ok = g.read_end >= g.read_begin && g.read_end <= batch->n_reads;
                if (ok && g.read_end > g.read_begin) {
                    const uint32_t lo = std::min(ck.read_lo, g.read_begin), hi = std::max(ck.read_hi, g.read_end);
                    const uint64_t nb = batch->seq_off[hi] - batch->seq_off[lo];
                    if (r > ck.r0 && 2 * nb > chunk_bytes) break;
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
    lcr_device_batch *cur = nullptr, *nxt = nullptr;
    rc = upload_impl(ctx, &chunks[0].view, &cur, true);
    for (size_t k = 0; k < chunks.size() && !rc; ++k) {
        if (k + 1 < chunks.size()) rc = upload_impl(ctx, &chunks[k + 1].view, &nxt, true); /* overlaps the run below */
        lcr_result *part = nullptr;
        if (!rc) rc = lcr_run_device(ctx, cur);
        if (!rc) rc = lcr_fetch(ctx, cur, &part);
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
                if (part->hp[i] != -1) all->hp[g] = part->hp[i];
                if (part->ps[i]) all->ps[g] = part->ps[i];
                if (part->is_fragment[i]) all->is_fragment[g] = 1;
            }
            st.n_reads_pass += part->stats.n_reads_pass; st.n_aligned_bases += part->stats.n_aligned_bases;
            st.n_positions += part->stats.n_positions; st.n_candidates += part->stats.n_candidates;
            st.n_fragments += part->stats.n_fragments; st.nnz_phase += part->stats.nnz_phase;
            st.n_cross_optimize += part->stats.n_cross_optimize; st.n_sweep_iters += part->stats.n_sweep_iters;
            lcr_free_result(part);
        }
        if (cur) lcr_release(ctx, cur);
        cur = nxt;
        nxt = nullptr;
    }
    if (cur) lcr_release(ctx, cur);
    if (nxt) lcr_release(ctx, nxt);
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
