/*
 * phase.cu — haplotype phasing of one region per CTA, all iterations on the device.
 *
 * Replaces:
 *   src/phase.rs:32-49      aki                                  (aki_fx)
 *   src/phase.rs:810-976    SNPFrag::cross_optimize              (cross_optimize)
 *   src/phase.rs:257-276    cal_overall_probability              (objective)
 *   src/phase.rs:600-691    init_haplotypes_LD2 / init_assignment / init_genotype
 *   src/phase.rs:1087-1296  SNPFrag::phase                       (phase_enum / phase_ld)
 *   src/phase.rs:1298-1394  cross_optimize_by_block              (cross_optimize_by_block)
 *   src/snpfrags.rs:548-625 assign_reads_haplotype               (assign_reads)
 *   src/snpfrags.rs:378-546 assign_snp_haplotype_genotype        (assign_snps)
 *   src/snpfrags.rs:191-376 eval_rna_edit_var_phase / eval_low_frac_var_phase (rescue)
 *   src/snpfrags.rs:628-733 assign_phase_set                     (phase_sets)
 *
 * All log10 sums are int64 fixed point (LCR_FX_FRAC fractional bits), so every sweep is an
 * order-independent integer reduction: reads flip on the sign of sum(p * sigma * delta * W[q])
 * over their heterozygous sites, SNPs pick the arg-max of four integer column sums.
 */
#include <cooperative_groups.h>

#include "lcr_frag.h"

namespace cg = cooperative_groups;

#define PB 128          /* threads per CTA of the per-region kernel (small regions dominate: more of them in flight per SM) */
#define PBG 512         /* threads per CTA of the cooperative (whole-GPU) kernel */
#define NONE32 0xffffffffu
#define FA(x, k) (((x).fp[k] & (x).need) == (x).need) /* fragment k takes part in the current stage */

namespace {

/* values every thread of the team must agree on; lives in shared memory (one CTA per region) or in
   global memory (cooperative kernel: the whole grid works on one large region) */
struct TeamBcast {
    int flag[3];
    long long acc[3];
    uint32_t root0;
    int decision;
};

struct Ctx {
    const PhaseArgs &a;
    const LcrDeviceTables &T;
    uint32_t reg, n, nf, cb, fb, tid;
    uint32_t nthreads;  /* team size: PB, or gridDim.x * PBG */
    bool grid;          /* team is the whole cooperative grid */
    uint32_t epoch_any, epoch_sum; /* number of tany() / tsum() calls so far (same in every thread) */
    TeamBcast *bc;
    lcr_candidate *c;
    uint64_t region_key;
    uint32_t slot0;
    /* region views */
    char4 *st;   /* per SNP: x = delta (haplotype), y = eta (genotype), z = for_phasing when phase() started, w = conserved */
    int8_t *best_hap, *best_gen, *tag, *best_tag;
    uint8_t *fp, *assign;
    const long long *OK, *ERR, *W; /* shared-memory copies of the fixed-point tables; W = ok - err */
    uint32_t *label, *rank, *work;
    long long *blk_q, *blk_qflip;
    const uint32_t *frag_slot, *frag_elem_off, *frag_links, *cover_off, *adj_off;
    long long *sh; /* shared scratch, PB/32 * 8 entries */
    uint32_t *piece_off, *piece_col; /* grid teams: column pieces of the delta / eta sweep */
    long long *col_acc;
    uint32_t n_pieces;
    /* fp[k]: bit 0 = fragment used for phasing (num_hete_links >= min_linkers, or promoted by a rescue pass), bit 1 = inside the
       --downsample set (always set when the region does not downsample).  need = 3 while the reference passes apply_downsampling,
       1 for the last assignment round and the phase sets (thread.rs:166-182) */
    uint8_t need;
    bool apply_ds; /* the region downsamples */
    unsigned long long n_iters;
};

__device__ __forceinline__ int64_t aki_tab(const long long *OK, const long long *ERR, int sigma, int delta, int eta, int p, int q) {
    const int x = eta == 0 ? sigma * delta : eta;
    return p == x ? OK[q] : ERR[q];
}
__device__ __forceinline__ int cell_p(int8_t cell) { return cell > 0 ? 1 : -1; }
__device__ __forceinline__ int cell_q(int8_t cell) { return (cell > 0 ? cell : -cell) - 1; }

/* team barrier */
__device__ __forceinline__ void tsync(Ctx &x) {
    if (x.grid) cg::this_grid().sync();
    else __syncthreads();
}
/* team-wide OR (also a barrier).  Grid teams rotate three flag slots so that a reset never races a later write. */
__device__ int tany(Ctx &x, int v) {
    if (!x.grid) return __syncthreads_or(v);
    const uint32_t k = x.epoch_any++;
    if (v) x.bc->flag[k % 3] = 1;
    cg::this_grid().sync();
    const int r = *(volatile int *)&x.bc->flag[k % 3];
    if (x.tid == 0) x.bc->flag[(k + 2) % 3] = 0;
    return r;
}
/* team-wide sum (also a barrier) */
__device__ long long tsum(Ctx &x, long long v) {
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    const uint32_t lt = threadIdx.x;
    if ((lt & 31) == 0) x.sh[lt >> 5] = v;
    __syncthreads();
    long long r = 0;
    const uint32_t nw = blockDim.x / 32;
    for (uint32_t w = 0; w < nw; ++w) r += x.sh[w];
    __syncthreads();
    if (!x.grid) return r;
    const uint32_t k = x.epoch_sum++;
    if (lt == 0 && r) atomicAdd((unsigned long long *)&x.bc->acc[k % 3], (unsigned long long)r);
    cg::this_grid().sync();
    r = *(volatile long long *)&x.bc->acc[k % 3];
    if (x.tid == 0) x.bc->acc[(k + 2) % 3] = 0;
    return r;
}

__device__ __forceinline__ double uniform(const Ctx &x, uint32_t stream, uint32_t call, uint32_t idx) {
    return lcr_uniform(x.a.P.seed, x.region_key, stream, call, idx);
}
__device__ __forceinline__ uint32_t read_rel(const Ctx &x, uint32_t k) { return x.frag_slot[k] - x.slot0; }
__device__ __forceinline__ int64_t aki_fx(const Ctx &x, int sigma, int delta, int eta, int p, int q) { return aki_tab(x.OK, x.ERR, sigma, delta, eta, p, q); }

struct ColFx {
    long long het_d = 0, het_nd = 0, homref = 0, homvar = 0;
    uint32_t cov = 0;
    const long long *OK, *ERR;
    __device__ __forceinline__ ColFx(const Ctx &x) : OK(x.OK), ERR(x.ERR) {}
};
__device__ __forceinline__ void col_add(const LcrDeviceTables &T, ColFx &c, int sigma, int delta, int p, int q) {
    c.het_d += aki_tab(c.OK, c.ERR, sigma, delta, 0, p, q);
    c.het_nd += aki_tab(c.OK, c.ERR, sigma, -delta, 0, p, q);
    c.homref += aki_tab(c.OK, c.ERR, sigma, delta, 1, p, q);
    c.homvar += aki_tab(c.OK, c.ERR, sigma, delta, -1, p, q);
    c.cov++;
}
__device__ __forceinline__ void col_reduce(ColFx &c) { /* all lanes end up with the warp totals */
    for (int o = 16; o; o >>= 1) {
        c.het_d += __shfl_xor_sync(0xffffffffu, c.het_d, o);
        c.het_nd += __shfl_xor_sync(0xffffffffu, c.het_nd, o);
        c.homref += __shfl_xor_sync(0xffffffffu, c.homref, o);
        c.homvar += __shfl_xor_sync(0xffffffffu, c.homvar, o);
        c.cov += __shfl_xor_sync(0xffffffffu, c.cov, o);
    }
}
__device__ __forceinline__ void col_L(const LcrDeviceTables &T, const ColFx &c, long long L[4], long long &D) {
    const long long ph = T.fx_prior_het - (long long)c.cov * T.fx_log10_2;
    L[0] = c.het_d + ph;
    L[1] = c.het_nd + ph;
    L[2] = c.homref + T.fx_prior_homref;
    L[3] = c.homvar + T.fx_prior_homvar;
    D = L[3] + L[0] + L[2] + L[1];
}

#ifdef LCR_PHASE_PROF /* cycles of one thread of the cooperative grid per part of cross_optimize (prof[10..15] of the counter block) */
#define GP_T(var) const long long var = clock64()
#define GP_ADD(slot, t0, t1) do { if (x.grid && x.tid == 0) x.a.ctr->prof[slot] += (unsigned long long)((t1) - (t0)); } while (0)
#else
#define GP_T(var)
#define GP_ADD(slot, t0, t1)
#endif

/* cal_overall_probability (phase.rs:257-276) */
__device__ long long objective(Ctx &x) {
    long long s = 0;
    /* eight lanes per fragment row (four on the cooperative grid: more rows in flight per pass) */
    const uint32_t lsh = x.grid ? 2u : 3u, lpr = 1u << lsh;
    for (uint32_t k = x.tid >> lsh; k < x.nf; k += x.nthreads >> lsh) {
        if (!FA(x, k) || x.tag[k] == 0) continue;
        const int sg = x.tag[k];
        for (uint32_t e = x.frag_elem_off[k] + (x.tid & (lpr - 1)); e < x.frag_elem_off[k + 1]; e += lpr) {
            const char4 st = x.st[x.a.elem_snp[e]];
            if (!st.z) continue;
            const int8_t cell = x.a.elem_cell[e];
            s += aki_fx(x, sg, st.x, st.y, cell_p(cell), cell_q(cell));
        }
    }
    return tsum(x, s);
}

/* cross_optimize (phase.rs:810-976) */
__device__ long long cross_optimize(Ctx &x, bool keep_conserved, bool with_genotype) {
    bool hg_increase = true, ht_increase = true;
    int num_iters = 0;
    long long obj_iter = 0; /* cooperative grid: this thread's share of the objective of the state the last delta / eta sweep left */
    while (hg_increase | ht_increase) {
        x.n_iters++;
        int better = 0;
        GP_T(g0);
        /* sigma sweep (phase.rs:823-868) */
        /* eight lanes per fragment row (four on the cooperative grid); the row sum is shuffled together inside the group */
        const uint32_t lsh = x.grid ? 2u : 3u, lpr = 1u << lsh;
        for (uint32_t k0 = 0; k0 < x.nf; k0 += x.nthreads >> lsh) {
            const uint32_t k = k0 + (x.tid >> lsh);
            long long diff = 0;
            int sg = 0;
            if (k < x.nf && FA(x, k)) sg = x.tag[k];
            if (sg != 0) {
                for (uint32_t e = x.frag_elem_off[k] + (x.tid & (lpr - 1)); e < x.frag_elem_off[k + 1]; e += lpr) {
                    const char4 st = x.st[x.a.elem_snp[e]];
                    if (!st.z || st.y != 0) continue;
                    const int8_t cell = x.a.elem_cell[e];
                    const long long W = x.W[cell_q(cell)];
                    diff += (cell_p(cell) == sg * st.x) ? W : -W;
                }
            }
            diff += __shfl_xor_sync(0xffffffffu, diff, 1);
            diff += __shfl_xor_sync(0xffffffffu, diff, 2);
            if (lpr == 8) diff += __shfl_xor_sync(0xffffffffu, diff, 4);
            if (sg != 0 && diff < 0) {
                if ((x.tid & (lpr - 1)) == 0) x.tag[k] = (int8_t)(-sg);
                better = 1;
            }
        }
        GP_T(g1);
        better = tany(x, better);
        GP_T(g2);
        GP_ADD(10, g0, g1); GP_ADD(11, g1, g2);
        if (!better) ht_increase = false;
        else { ht_increase = true; hg_increase = true; }
        /* delta / eta sweep (phase.rs:872-958) */
        better = 0;
        auto decide = [&](uint32_t i, const ColFx &col, int d, int eta) {
            long long L[4], D;
            col_L(x.T, col, L, D);
            const long long L_old = eta == 0 ? L[0] : (eta == 1 ? L[2] : L[3]);
            long long L_new;
            int nd = d, ne = eta;
            if (with_genotype) {
                long long mx = L[0] > L[1] ? L[0] : L[1];
                const long long m2 = L[2] > L[3] ? L[2] : L[3];
                mx = mx > m2 ? mx : m2;
                if (L[0] == mx) { nd = d; ne = 0; L_new = L[0]; }
                else if (L[1] == mx) { nd = -d; ne = 0; L_new = L[1]; }
                else if (L[2] == mx) { nd = d; ne = 1; L_new = L[2]; }
                else { nd = d; ne = -1; L_new = L[3]; }
            } else if (eta == 0) {
                if (L[0] >= L[1]) { nd = d; ne = 0; L_new = L[0]; } else { nd = -d; ne = 0; L_new = L[1]; }
            } else {
                if (L[2] >= L[3]) { nd = d; ne = 1; L_new = L[2]; } else { nd = d; ne = -1; L_new = L[3]; }
            }
            x.st[i].x = (int8_t)nd;
            x.st[i].y = (int8_t)ne;
            if (L_new > L_old) better = 1;
        };
        if (x.grid) {
            /* the whole GPU on one region: one warp per PIECE of a column (a deep site would otherwise hold every CTA at the barrier),
               partial sums by 64-bit atomics (exact integers: any order), then one thread per site decides */
            for (uint32_t p = x.tid >> 5; p < x.n_pieces; p += x.nthreads >> 5) {
                const uint32_t i = x.piece_col[p];
                const char4 sti = x.st[i];
                if (!sti.z) continue; /* conserved sites are summed too: they are not decided, but they count in the objective */
                const uint32_t w0 = x.cover_off[i] + (p - x.piece_off[i]) * LCR_COL_PIECE;
                const uint32_t w1 = min(w0 + LCR_COL_PIECE, x.cover_off[i + 1]);
                ColFx col(x);
                /* the piece's four 32-element trips with their loads issued together (each is an L2 round trip) */
                uint32_t kk[LCR_COL_PIECE / 32];
                int tg[LCR_COL_PIECE / 32];
                int8_t cl[LCR_COL_PIECE / 32];
#pragma unroll
                for (uint32_t j = 0; j < LCR_COL_PIECE / 32; ++j) {
                    const uint32_t w = w0 + (x.tid & 31) + 32 * j;
                    kk[j] = w < w1 ? x.a.cover_frag[w] : NONE32;
                    cl[j] = w < w1 ? x.a.cover_cell[w] : (int8_t)1;
                }
#pragma unroll
                for (uint32_t j = 0; j < LCR_COL_PIECE / 32; ++j) tg[j] = kk[j] != NONE32 ? (int)x.tag[kk[j]] : 0;
                if (x.apply_ds) { /* without downsampling a haplotag is non-zero only on a fragment that takes part (init_assignment, sign flips) */
#pragma unroll
                    for (uint32_t j = 0; j < LCR_COL_PIECE / 32; ++j)
                        if (tg[j] != 0 && !FA(x, kk[j])) tg[j] = 0;
                }
#pragma unroll
                for (uint32_t j = 0; j < LCR_COL_PIECE / 32; ++j)
                    if (tg[j] != 0) col_add(x.T, col, tg[j], sti.x, cell_p(cl[j]), cell_q(cl[j]));
                col_reduce(col);
                if ((x.tid & 31) == 0 && col.cov) {
                    unsigned long long *acc = reinterpret_cast<unsigned long long *>(x.col_acc + 5 * (size_t)i);
                    atomicAdd(acc + 0, (unsigned long long)col.het_d);
                    atomicAdd(acc + 1, (unsigned long long)col.het_nd);
                    atomicAdd(acc + 2, (unsigned long long)col.homref);
                    atomicAdd(acc + 3, (unsigned long long)col.homvar);
                    atomicAdd(acc + 4, (unsigned long long)col.cov);
                }
            }
            tsync(x);
            obj_iter = 0;
            for (uint32_t i = x.tid; i < x.n; i += x.nthreads) {
                const char4 sti = x.st[i];
                if (!sti.z) continue;
                long long *acc = x.col_acc + 5 * (size_t)i;
                ColFx col(x);
                col.het_d = __ldcg(acc + 0); col.het_nd = __ldcg(acc + 1); col.homref = __ldcg(acc + 2); col.homvar = __ldcg(acc + 3);
                col.cov = (uint32_t)__ldcg(acc + 4);
                if (!col.cov) continue;
                acc[0] = 0; acc[1] = 0; acc[2] = 0; acc[3] = 0; acc[4] = 0;
                int nd = sti.x, ne = sti.y;
                if (!(keep_conserved && sti.w)) {
                    decide(i, col, sti.x, sti.y);
                    nd = x.st[i].x; ne = x.st[i].y;
                }
                /* cal_overall_probability (phase.rs:257-276) column by column: the state this pass leaves is the state the call returns,
                   and its column sums were taken with the haplotags the last sigma sweep left */
                obj_iter += ne == 0 ? (nd == sti.x ? col.het_d : col.het_nd) : (ne == 1 ? col.homref : col.homvar);
            }
        } else {
            /* one warp per SNP: lanes stride the column, five shuffled sums */
            for (uint32_t i = x.tid >> 5; i < x.n; i += x.nthreads >> 5) {
                if (!x.st[i].z) continue;
                if (keep_conserved && x.st[i].w) continue;
                const int d = x.st[i].x, eta = x.st[i].y;
                ColFx col(x);
                for (uint32_t w = x.cover_off[i] + (x.tid & 31); w < x.cover_off[i + 1]; w += 32) {
                    const uint32_t k = x.a.cover_frag[w];
                    if (!FA(x, k) || x.tag[k] == 0) continue;
                    const int8_t cell = x.a.cover_cell[w];
                    col_add(x.T, col, x.tag[k], d, cell_p(cell), cell_q(cell));
                }
                col_reduce(col);
                if ((x.tid & 31) != 0) continue;
                if (!col.cov) continue;
                decide(i, col, d, eta);
            }
        }
        GP_T(g3);
        better = tany(x, better);
        GP_T(g4);
        GP_ADD(12, g2, g3); GP_ADD(13, g3, g4);
        if (!better) hg_increase = false;
        else { hg_increase = true; ht_increase = true; }
        if (++num_iters > 20) break;
    }
    GP_T(g5);
    const long long obj = x.grid ? tsum(x, obj_iter) : objective(x);
    GP_T(g6);
    GP_ADD(14, g5, g6);
    return obj;
}

__device__ void save_best(Ctx &x) {
    for (uint32_t i = x.tid; i < x.n; i += x.nthreads) { x.best_hap[i] = x.st[i].x; x.best_gen[i] = x.st[i].y; }
    for (uint32_t k = x.tid; k < x.nf; k += x.nthreads) x.best_tag[k] = x.tag[k];
    tsync(x);
}
__device__ void load_best(Ctx &x) {
    for (uint32_t i = x.tid; i < x.n; i += x.nthreads) { x.st[i].x = x.best_hap[i]; x.st[i].y = x.best_gen[i]; }
    for (uint32_t k = x.tid; k < x.nf; k += x.nthreads) x.tag[k] = x.best_tag[k];
    tsync(x);
}
__device__ void init_genotype(Ctx &x) { /* phase.rs:682-691 */
    for (uint32_t i = x.tid; i < x.n; i += x.nthreads) {
        const int vt = x.c[i].variant_type;
        x.st[i].y = (int8_t)(vt == 0 ? 1 : (vt == 1 ? 0 : ((vt == 2 || vt == 3) ? -1 : x.st[i].y)));
    }
}
__device__ void init_assignment(Ctx &x, uint32_t call) { /* phase.rs:673-680 */
    for (uint32_t k = x.tid; k < x.nf; k += x.nthreads)
        if (x.fp[k] & 1) x.tag[k] = uniform(x, LCR_RNG_INIT_SIGMA, call, read_rel(x, k)) < 0.5 ? -1 : 1; /* every fragment used for phasing, sampled or not */
}

/* phase.rs:1097-1122: all 2^n starting haplotypes */
__device__ bool phase_enum(Ctx &x) {
    /* the search kernel (phase_enum.cu) already ran every configuration and left the winner's final state in the best_* arrays */
    if (x.a.es_done && x.a.es_done[x.reg] == 0x80000000u) {
        load_best(x);
        return true;
    }
    long long best = 0;
    bool have = false;
    const uint32_t n_cfg = 1u << x.n;
    for (uint32_t cfg = 0; cfg < n_cfg; ++cfg) {
        for (uint32_t i = x.tid; i < x.n; i += x.nthreads) x.st[i].x = ((cfg >> i) & 1u) ? -1 : 1;
        init_assignment(x, cfg);
        init_genotype(x);
        tsync(x);
        const long long prob = cross_optimize(x, false, true);
        if (!have || prob > best) { best = prob; have = true; save_best(x); }
    }
    load_best(x);
    return false;
}

/* does fragment k carry, before SNP idx in its row, an element outside the block rooted at `root`? (phase.rs:1334-1338) */
__device__ __forceinline__ bool flip_read_of(const Ctx &x, uint32_t k, uint32_t idx, uint32_t root) {
    for (uint32_t e = x.frag_elem_off[k]; e < x.frag_elem_off[k + 1]; ++e) {
        const uint32_t s = x.a.elem_snp[e];
        if (s >= idx) break;
        if (x.label[s] != root) return false;
    }
    return true;
}

/* cross_optimize_by_block (phase.rs:1298-1394); block scores are sums of round((1 - L1/D) * 2^40) */
__device__ long long cross_optimize_by_block(Ctx &x, uint32_t root0) {
    for (uint32_t i = x.tid; i < x.n; i += x.nthreads) { x.blk_q[i] = 0; x.blk_qflip[i] = 0; }
    tsync(x);
    for (uint32_t i = x.tid; i < x.n; i += x.nthreads) {
        const uint32_t root = x.label[i];
        if (root == NONE32) continue;
        const int d = x.st[i].x, eta = x.st[i].y;
        ColFx c0(x), c1(x);
        for (uint32_t w = x.cover_off[i]; w < x.cover_off[i + 1]; ++w) {
            const uint32_t k = x.a.cover_frag[w];
            if (!FA(x, k) || x.tag[k] == 0) continue;
            const int8_t cell = x.a.cover_cell[w];
            const int sg = x.tag[k];
            const int sf = flip_read_of(x, k, i, root) ? -sg : sg;
            col_add(x.T, c0, sg, d, cell_p(cell), cell_q(cell));
            col_add(x.T, c1, sf, -d, cell_p(cell), cell_q(cell));
        }
        long long L[4], D;
        col_L(x.T, c0, L, D);
        const double t0 = 1.0 - (double)(eta == 0 ? L[0] : (eta == 1 ? L[2] : L[3])) / (double)D;
        col_L(x.T, c1, L, D);
        const double t1 = 1.0 - (double)(eta == 0 ? L[0] : (eta == 1 ? L[2] : L[3])) / (double)D;
        atomicAdd((unsigned long long *)&x.blk_q[root], (unsigned long long)__double2ll_rn(t0 * 1099511627776.0));
        atomicAdd((unsigned long long *)&x.blk_qflip[root], (unsigned long long)__double2ll_rn(t1 * 1099511627776.0));
    }
    tsync(x);
    /* haplotags: only the last block of ld_blocks (the component of the lowest LD node) survives the
       per-block rewrite of tmp_haplotag (phase.rs:1365-1378) */
    if (root0 != NONE32 && x.blk_q[root0] < x.blk_qflip[root0]) {
        for (uint32_t k = x.tid; k < x.nf; k += x.nthreads) {
            if (!FA(x, k) || x.tag[k] == 0) continue;
            uint32_t best_idx = NONE32, best_rank = 0;
            for (uint32_t e = x.frag_elem_off[k]; e < x.frag_elem_off[k + 1]; ++e) {
                const uint32_t s = x.a.elem_snp[e];
                if (x.label[s] != root0) continue;
                if (best_idx == NONE32 || x.rank[s] > best_rank) { best_idx = s; best_rank = x.rank[s]; }
            }
            if (best_idx == NONE32) continue;
            if (flip_read_of(x, k, best_idx, root0)) x.tag[k] = (int8_t)(-x.tag[k]);
        }
    }
    for (uint32_t i = x.tid; i < x.n; i += x.nthreads) {
        const uint32_t root = x.label[i];
        if (root != NONE32 && x.blk_q[root] < x.blk_qflip[root]) x.st[i].x = (int8_t)(-x.st[i].x);
    }
    tsync(x);
    return objective(x);
}

/* phase.rs:1123-1294: LD-seeded start, block flips, random perturbation restarts */
__device__ void phase_ld(Ctx &x) {
    const uint32_t *adj_off = x.adj_off;
    const uint32_t adj_base = adj_off[0];
    const uint32_t *adj = x.a.adj;
    GP_T(gb0);
    /* init_haplotypes_LD2 (phase.rs:600-652) */
    for (uint32_t i = x.tid; i < x.n; i += x.nthreads) {
        x.st[i].x = uniform(x, LCR_RNG_INIT_DELTA, 0, i) < 0.5 ? 1 : -1;
        x.label[i] = NONE32;
        x.rank[i] = 0;
        x.st[i].w = adj_off[i + 1] > adj_off[i] ? 1 : 0;
    }
    tsync(x);
    if (x.tid < 32) {
        /* Bfs from the first node of every block: a node takes its sign from the neighbour that was dequeued first, which is
           the node that discovered it.  One warp walks the queue in the serial order; the lanes look at 32 neighbours of the
           dequeued node at a time and append the undiscovered ones in adjacency order (prefix counts of the ballot), which is
           exactly what the one-by-one loop does because a node's adjacency list holds no duplicates. */
        const uint32_t lane = x.tid, lt = (1u << lane) - 1u;
        uint32_t root0 = NONE32;
        uint32_t *queue = x.work;
        for (uint32_t r = 0; r < x.n; ++r) {
            if (adj_off[r + 1] == adj_off[r] || *(volatile uint32_t *)&x.label[r] != NONE32) continue;
            if (root0 == NONE32) root0 = r;
            uint32_t qh = 0, qt = 1;
            if (lane == 0) { x.label[r] = r; x.st[r].x = 1; queue[0] = r; }
            __syncwarp();
            while (qh < qt) {
                const uint32_t nx = *(volatile uint32_t *)&queue[qh++];
                const int8_t sx = *(volatile int8_t *)&x.st[nx].x;
                const uint32_t a0 = adj_off[nx], a1 = adj_off[nx + 1];
                for (uint32_t w0 = a0; w0 < a1; w0 += 128) { /* four 32-wide steps with their loads issued together (each is an L2 round trip) */
                    uint32_t av[4];
                    bool fresh[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) { const uint32_t w = w0 + 32 * c + lane; av[c] = w < a1 ? adj[w] : NONE32; }
#pragma unroll
                    for (int c = 0; c < 4; ++c) fresh[c] = av[c] != NONE32 && *(volatile uint32_t *)&x.label[av[c] & 0x7fffffffu] == NONE32; /* no duplicates in one list */
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t m = __ballot_sync(0xffffffffu, fresh[c]);
                        if (fresh[c]) {
                            const uint32_t v = av[c] & 0x7fffffffu;
                            x.label[v] = r;
                            x.st[v].x = (av[c] & 0x80000000u) ? (int8_t)(-sx) : sx;
                            queue[qt + __popc(m & lt)] = v;
                        }
                        qt += __popc(m);
                    }
                    __syncwarp();
                }
            }
        }
        /* order of block[0..] for the last block: petgraph Dfs from its first node (pop; skip if seen; number it; push its unseen
           neighbours in adjacency order) with the same 32-wide pushes */
        if (root0 != NONE32) {
            uint32_t *stack = x.work;
            uint32_t sp = 1, order = 0;
            if (lane == 0) stack[0] = root0;
            __syncwarp();
            while (sp) {
                const uint32_t node = *(volatile uint32_t *)&stack[--sp];
                if (*(volatile uint32_t *)&x.rank[node]) continue;
                ++order;
                if (lane == 0) x.rank[node] = order;
                __syncwarp();
                const uint32_t a0 = adj_off[node], a1 = adj_off[node + 1];
                for (uint32_t w0 = a0; w0 < a1; w0 += 128) {
                    uint32_t av[4];
                    bool push[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) { const uint32_t w = w0 + 32 * c + lane; av[c] = w < a1 ? (adj[w] & 0x7fffffffu) : NONE32; }
#pragma unroll
                    for (int c = 0; c < 4; ++c) push[c] = av[c] != NONE32 && !*(volatile uint32_t *)&x.rank[av[c]];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t m = __ballot_sync(0xffffffffu, push[c]);
                        if (push[c]) stack[sp + __popc(m & lt)] = av[c];
                        sp += __popc(m);
                    }
                    __syncwarp();
                }
            }
        }
        if (lane == 0) x.bc->root0 = root0;
        (void)adj_base;
    }
    tsync(x);
    GP_T(gb1);
    GP_ADD(9, gb0, gb1);
    const uint32_t root0 = *(volatile uint32_t *)&x.bc->root0;
    init_genotype(x);
    init_assignment(x, 0);
    tsync(x);
    long long best = cross_optimize(x, true, false);
    save_best(x);
    load_best(x);
    GP_T(gb2);
    long long prob = cross_optimize_by_block(x, root0);
    GP_T(gb3);
    GP_ADD(8, gb2, gb3);
    if (prob > best) { best = prob; save_best(x); }
    load_best(x);
    for (uint32_t t = 0; t <= x.n / 4; ++t) {
        const bool flip = (t & 1u) == 1u;
        for (uint32_t i = x.tid; i < x.n; i += x.nthreads) {
            const double rg = uniform(x, LCR_RNG_PERTURB_DELTA, t, i);
            if (rg < 0.1) x.st[i].x = flip ? 1 : -1;
            else if (rg >= 0.9) x.st[i].x = flip ? -1 : 1;
        }
        tsync(x);
        prob = cross_optimize(x, false, false);
        if (prob > best) { best = prob; save_best(x); }
        load_best(x);
        for (uint32_t k = x.tid; k < x.nf; k += x.nthreads) {
            if (!(x.fp[k] & 1) || x.tag[k] == 0) continue; /* phase.rs:1218-1225: no downsampling test here */
            if (uniform(x, LCR_RNG_PERTURB_SIGMA, t, read_rel(x, k)) < 0.1) x.tag[k] = (int8_t)(-x.tag[k]);
        }
        tsync(x);
        prob = cross_optimize(x, false, false);
        if (prob > best) { best = prob; save_best(x); }
        load_best(x);
    }
}

/* assign_reads_haplotype (snpfrags.rs:548-625) */
__device__ void assign_reads(Ctx &x, bool record) {
    for (uint32_t k = x.tid; k < x.nf; k += x.nthreads) {
        if (!FA(x, k)) continue;
        const int sg = x.tag[k];
        long long A = 0, B = 0;
        uint32_t cnt = 0;
        bool q0_used = false;
        for (uint32_t e = x.frag_elem_off[k]; e < x.frag_elem_off[k + 1]; ++e) {
            const lcr_candidate &s = x.c[x.a.elem_snp[e]];
            if (!(s.flags & LCR_CF_FOR_PHASING) || s.haplotype == 0 || s.genotype != 0) continue;
            const int8_t cell = x.a.elem_cell[e];
            cnt++;
            if (cell_q(cell) == 0) { q0_used = true; continue; } /* only at a rescued site: see below */
            A += aki_fx(x, sg, s.haplotype, 0, cell_p(cell), cell_q(cell));
            B += aki_fx(x, -sg, s.haplotype, 0, cell_p(cell), cell_q(cell));
        }
        int asg = 0;
        if (sg == 0 || cnt == 0) { x.tag[k] = 0; }
        /* a quality-0 element has prob = 1.0: one of the two sums holds log10(0) = -inf, one of q, qn is NaN, and
           `(q - qn).abs() >= cutoff` sends the read to the unknown group (snpfrags.rs:580-618) */
        else if (q0_used) { x.tag[k] = 0; }
        else {
            const double den = lcr_fx_to_f64(A + B);
            const double q = 1.0 - lcr_fx_to_f64(A) / den;
            const double qn = 1.0 - lcr_fx_to_f64(B) / den;
            if (fabs(q - qn) >= x.a.P.read_assignment_cutoff) {
                if (q >= qn) asg = sg == 1 ? 1 : 2;
                else { asg = sg == 1 ? 2 : 1; x.tag[k] = (int8_t)(sg == 1 ? -1 : 1); }
            } else x.tag[k] = 0;
        }
        x.assign[k] = (uint8_t)asg;
        if (record) {
            /* a read shared by several regions keeps the entry of the lowest region (thread.rs:308-314 keeps the first per qname) */
            const uint32_t read = x.a.regions[x.reg].read_begin + read_rel(x, k);
            atomicMin(&x.a.hp_key[read], (x.reg << 2) | (uint32_t)asg);
        }
    }
    tsync(x);
}

/* -10 log10(1 - cal_phase_score_log) from three column sums (phase.rs:238-255, snpfrags.rs:483) */
__device__ __forceinline__ double phase_score_from(long long L1, long long L2, long long L3) {
    const double v = 1.0 - lcr_fx_to_f64(L1) / lcr_fx_to_f64(L2 + L3);
    return -10.0 * lcr_log10(1.0 - v);
}

/* assign_snp_haplotype_genotype (snpfrags.rs:378-546): one warp per SNP, lane 0 writes the record */
__device__ void assign_snps(Ctx &x) {
    const uint32_t lane = x.tid & 31;
    for (uint32_t i = x.tid >> 5; i < x.n; i += x.nthreads >> 5) {
        lcr_candidate &s = x.c[i];
        const uint16_t flags0 = s.flags;
        const int vt0 = s.variant_type;
        const int d = s.haplotype;
        __syncwarp();
        if (!(flags0 & LCR_CF_FOR_PHASING)) { if (lane == 0) s.flags = flags0 | LCR_CF_NON_SELECTED; continue; }
        if (x.cover_off[i + 1] == x.cover_off[i]) { if (lane == 0) s.flags = flags0 | LCR_CF_SINGLE; continue; }
        ColFx col(x);
        long long Lp = 0, Lm = 0; /* het sums for delta = +1 / -1 */
        int hap1 = 0, hap2 = 0;
        for (uint32_t w = x.cover_off[i] + lane; w < x.cover_off[i + 1]; w += 32) {
            const uint32_t k = x.a.cover_frag[w];
            if (!FA(x, k) || x.frag_links[k] < x.a.P.min_linkers) continue;
            if (vt0 == 1 && x.assign[k] == 0) continue;
            if (x.assign[k] == 1) hap1++; else if (x.assign[k] == 2) hap2++;
            const int8_t cell = x.a.cover_cell[w];
            const int p = cell_p(cell), q = cell_q(cell), sg = x.tag[k];
            col_add(x.T, col, sg, d, p, q);
            Lp += aki_fx(x, sg, 1, 0, p, q);
            Lm += aki_fx(x, sg, -1, 0, p, q);
        }
        col_reduce(col);
        for (int o = 16; o; o >>= 1) {
            Lp += __shfl_xor_sync(0xffffffffu, Lp, o);
            Lm += __shfl_xor_sync(0xffffffffu, Lm, o);
            hap1 += __shfl_xor_sync(0xffffffffu, hap1, o);
            hap2 += __shfl_xor_sync(0xffffffffu, hap2, o);
        }
        if (lane != 0) continue;
        if (!col.cov) { s.flags = flags0 | LCR_CF_NON_SELECTED; continue; }
        long long L[4], D;
        col_L(x.T, col, L, D);
        long long mx = L[0] > L[1] ? L[0] : L[1];
        const long long m2 = L[2] > L[3] ? L[2] : L[3];
        mx = mx > m2 ? mx : m2;
        if (L[0] == mx) { s.haplotype = (int8_t)d; s.genotype = 0; s.variant_type = 1; }
        else if (L[1] == mx) { s.haplotype = (int8_t)(-d); s.genotype = 0; s.variant_type = 1; }
        else if (L[2] == mx) { s.haplotype = (int8_t)d; s.genotype = 1; s.variant_type = 0; }
        else { s.haplotype = (int8_t)d; s.genotype = -1; if (vt0 != 2 && vt0 != 3) s.variant_type = 2; }
        if (s.genotype != 0) { s.flags = flags0 | LCR_CF_NON_SELECTED; continue; }
        if (hap1 >= 1 && hap2 >= 1) s.phase_score = phase_score_from(s.haplotype == 1 ? Lp : Lm, Lp, Lm);
        else s.phase_score = 0.19940219;
    }
    tsync(x);
}

/* eval_rna_edit_var_phase / eval_low_frac_var_phase (snpfrags.rs:191-376), candidates visited in list order */
__device__ void rescue(Ctx &x, uint16_t list_flag, bool low_frac) {
    const float mps = x.a.P.min_phase_score - 3.0f;
    for (uint32_t ti = 0; ti < x.n; ++ti) {
        lcr_candidate &s = x.c[ti];
        if (!(s.flags & list_flag)) continue;
        if (x.cover_off[ti + 1] == x.cover_off[ti]) { if (x.tid == 0) s.flags |= LCR_CF_SINGLE; tsync(x); continue; }
        if (s.variant_type != 1) { if (x.tid == 0) s.flags |= LCR_CF_NON_SELECTED; tsync(x); continue; }
        long long Lp = 0, Lm = 0, h1 = 0, h2 = 0, cnt = 0, zz = 0;
        for (uint32_t w = x.cover_off[ti] + x.tid; w < x.cover_off[ti + 1]; w += x.nthreads) {
            const uint32_t k = x.a.cover_frag[w];
            if (!FA(x, k) || x.assign[k] == 0 || x.frag_links[k] < x.a.P.min_linkers) continue;
            if (x.assign[k] == 1) h1++; else if (x.assign[k] == 2) h2++;
            const int8_t cell = x.a.cover_cell[w];
            const int p = cell_p(cell), q = cell_q(cell), sg = x.tag[k];
            cnt++;
            if (q == 0) { zz += (p == sg) ? 1ll : (1ll << 32); continue; } /* aki = 0 for delta = +1 (low half) / delta = -1 (high half) */
            Lp += aki_fx(x, sg, 1, 0, p, q);
            Lm += aki_fx(x, sg, -1, 0, p, q);
        }
        Lp = tsum(x, Lp); Lm = tsum(x, Lm); h1 = tsum(x, h1); h2 = tsum(x, h2); cnt = tsum(x, cnt); zz = tsum(x, zz);
        if (x.tid == 0) {
            int decision = 0;
            if (cnt == 0 || h1 < 2 || h2 < 2) s.flags |= LCR_CF_SINGLE;
            else {
                double s1, s2;
                if (zz) {
                    /* quality-0 elements: log_q2 and / or log_q3 of cal_phase_score_log (phase.rs:238-255) is -inf; the score of the delta
                       whose own sum is infinite is NaN (inf / inf), the other one is -10 log10(1 - 1) = +inf */
                    const double qnan = lcr_u2d(0x7ff8000000000000ULL), pinf = lcr_u2d(0x7ff0000000000000ULL);
                    s1 = (zz & 0xffffffffll) ? qnan : pinf;
                    s2 = (zz >> 32) ? qnan : pinf;
                } else { s1 = phase_score_from(Lp, Lp, Lm); s2 = phase_score_from(Lm, Lp, Lm); }
                s.flags &= ~LCR_CF_SINGLE;
                const double best = fmax(s1, s2);
                if (best >= (double)mps) {
                    s.flags &= ~(LCR_CF_NON_SELECTED | LCR_CF_RNA_EDITING);
                    if (low_frac) s.flags &= ~LCR_CF_CAND_SOMATIC;
                    s.flags |= LCR_CF_FOR_PHASING;
                    s.haplotype = s1 >= s2 ? 1 : -1;
                    s.genotype = 0;
                    s.variant_type = 1;
                    s.phase_score = best;
                    decision = 1;
                } else {
                    s.flags |= LCR_CF_NON_SELECTED;
                    if (low_frac) { s.flags |= LCR_CF_CAND_SOMATIC; s.flags &= ~LCR_CF_FOR_PHASING; }
                    else s.flags |= LCR_CF_RNA_EDITING;
                }
            }
            x.bc->decision = decision;
        }
        tsync(x);
        if (*(volatile int *)&x.bc->decision) {
            for (uint32_t w = x.cover_off[ti] + x.tid; w < x.cover_off[ti + 1]; w += x.nthreads) {
                const uint32_t k = x.a.cover_frag[w];
                x.fp[k] |= 1;
                if (x.tag[k] == 0 || x.assign[k] == 0) x.tag[k] = uniform(x, LCR_RNG_RESCUE_SIGMA, ti, read_rel(x, k)) < 0.5 ? -1 : 1;
            }
        }
        tsync(x);
    }
}

/* assign_phase_set (snpfrags.rs:628-733): a read links the nodes it sees with the same p * delta */
__device__ void phase_sets(Ctx &x) {
    const double mps = (double)x.a.P.min_phase_score;
    for (uint32_t i = x.tid; i < x.n; i += x.nthreads) {
        const lcr_candidate &s = x.c[i];
        const bool node = s.genotype == 0 && s.variant_type == 1 && !(s.flags & (LCR_CF_DENSE | LCR_CF_RNA_EDITING)) && !(s.phase_score < mps);
        x.label[i] = node ? i : NONE32;
    }
    tsync(x);
    for (;;) {
        int changed = 0;
        for (uint32_t k = x.tid; k < x.nf; k += x.nthreads) {
            if (!(x.fp[k] & 1) || x.assign[k] == 0) continue;
            uint32_t mn[2] = {NONE32, NONE32};
            for (uint32_t e = x.frag_elem_off[k]; e < x.frag_elem_off[k + 1]; ++e) {
                const uint32_t i = x.a.elem_snp[e];
                const uint32_t l = x.label[i];
                if (l == NONE32) continue;
                const int cls = (cell_p(x.a.elem_cell[e]) * x.c[i].haplotype) > 0 ? 0 : 1;
                if (l < mn[cls]) mn[cls] = l;
            }
            for (uint32_t e = x.frag_elem_off[k]; e < x.frag_elem_off[k + 1]; ++e) {
                const uint32_t i = x.a.elem_snp[e];
                const uint32_t l = x.label[i];
                if (l == NONE32) continue;
                const int cls = (cell_p(x.a.elem_cell[e]) * x.c[i].haplotype) > 0 ? 0 : 1;
                if (mn[cls] < l) { atomicMin(&x.label[i], mn[cls]); changed = 1; }
            }
        }
        /* pointer jumping */
        tsync(x);
        for (uint32_t i = x.tid; i < x.n; i += x.nthreads) {
            uint32_t l = x.label[i];
            if (l == NONE32) continue;
            while (x.label[l] < l) l = x.label[l];
            if (l < x.label[i]) { x.label[i] = l; changed = 1; }
        }
        if (!tany(x, changed)) break;
    }
    for (uint32_t i = x.tid; i < x.n; i += x.nthreads)
        if (x.label[i] != NONE32) x.c[i].phase_set = (uint32_t)(x.c[x.label[i]].pos + 1);
    /* components are visited in descending order of their first node; a read keeps the first id it meets */
    for (uint32_t k = x.tid; k < x.nf; k += x.nthreads) {
        if (!(x.fp[k] & 1) || x.assign[k] == 0) continue;
        uint32_t cnt[2] = {0, 0}, root[2] = {0, 0}, nn = 0;
        for (uint32_t e = x.frag_elem_off[k]; e < x.frag_elem_off[k + 1]; ++e) {
            const uint32_t i = x.a.elem_snp[e];
            const uint32_t l = x.label[i];
            if (l == NONE32) continue;
            const int cls = (cell_p(x.a.elem_cell[e]) * x.c[i].haplotype) > 0 ? 0 : 1;
            cnt[cls]++; root[cls] = l; nn++;
        }
        uint32_t best = NONE32;
        if (nn == 1) best = cnt[0] ? root[0] : root[1];
        else {
            if (cnt[0] >= 2) best = root[0];
            if (cnt[1] >= 2 && (best == NONE32 || root[1] > best)) best = root[1];
        }
        if (best != NONE32) {
            const uint32_t read = x.a.regions[x.reg].read_begin + read_rel(x, k);
            atomicMin(&x.a.ps_key[read], ((unsigned long long)x.reg << 32) | (unsigned long long)(uint32_t)(x.c[best].pos + 1));
        }
    }
    tsync(x);
}

__device__ __forceinline__ void load_tables(const PhaseArgs &a, long long (*tabs)[32]) {
    if (threadIdx.x < 32) {
        const uint32_t q = threadIdx.x < 31 ? threadIdx.x : 30;
        tabs[0][threadIdx.x] = a.tables->fx_ok[q];
        tabs[1][threadIdx.x] = a.tables->fx_err[q];
        tabs[2][threadIdx.x] = a.tables->fx_ok[q] - a.tables->fx_err[q];
    }
    __syncthreads();
}

/* the whole worker body after the fragment matrix exists: phase(), then thread.rs:168-201 */
__device__ void run_region(Ctx &x) {
    const PhaseArgs &a = x.a;
    const LcrRegionState rs = a.rstate[x.reg];
    x.n = rs.n_cand; x.nf = rs.n_frag; x.cb = rs.cand_begin; x.fb = rs.frag_begin;
    x.c = a.cand + x.cb;
    x.region_key = lcr_region_key(a.regions[x.reg].tid, a.regions[x.reg].start);
    x.slot0 = a.slot_off[x.reg];
    x.st = a.st + x.cb; x.best_hap = a.best_hap + x.cb; x.best_gen = a.best_gen + x.cb;
    x.label = a.label + x.cb; x.rank = a.rank + x.cb;
    x.blk_q = a.blk_q + x.cb; x.blk_qflip = a.blk_qflip + x.cb;
    x.tag = a.tag + x.fb; x.best_tag = a.best_tag + x.fb; x.fp = a.fp + x.fb; x.assign = a.assign + x.fb;
    x.frag_slot = a.frag_slot + x.fb; x.frag_elem_off = a.frag_elem_off + x.fb; x.frag_links = a.frag_links + x.fb;
    x.cover_off = a.cover_off + x.cb;
    x.adj_off = a.adj_off ? a.adj_off + x.cb : nullptr;
    x.work = (a.adj_off && a.work) ? a.work + (a.adj_off[x.cb] + x.cb) : nullptr;
    x.n_iters = 0;
    x.epoch_any = 0;
    x.epoch_sum = 0;
    x.n_pieces = 0;
    if (x.grid) { /* column pieces of the delta / eta sweep: offsets by one warp (a few thousand sites), the list by everybody */
        x.piece_off = a.piece_off + x.cb; x.piece_col = a.piece_col; x.col_acc = a.col_acc + 5 * (size_t)x.cb;
        if (x.tid < 32) {
            uint32_t run = 0;
            for (uint32_t i0 = 0; i0 < x.n; i0 += 32) {
                const uint32_t i = i0 + x.tid;
                const uint32_t np = i < x.n ? (x.cover_off[i + 1] - x.cover_off[i] + LCR_COL_PIECE - 1) / LCR_COL_PIECE : 0u;
                uint32_t incl = np;
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if ((int)x.tid >= o) incl += v;
                }
                if (i < x.n) x.piece_off[i] = run + incl - np;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (x.tid == 0) x.piece_off[x.n] = run;
        }
        for (uint32_t i = x.tid; i < 5 * x.n; i += x.nthreads) x.col_acc[i] = 0;
        tsync(x);
        x.n_pieces = __ldcg(&x.piece_off[x.n]);
        for (uint32_t i = x.tid; i < x.n; i += x.nthreads)
            for (uint32_t p = __ldcg(&x.piece_off[i]); p < __ldcg(&x.piece_off[i + 1]); ++p) x.piece_col[p] = i;
    }

    for (uint32_t i = x.tid; i < x.n; i += x.nthreads) {
        x.st[i].x = 0;
        x.st[i].y = x.c[i].genotype;
        x.st[i].z = (x.c[i].flags & LCR_CF_FOR_PHASING) ? 1 : 0;
        x.st[i].w = 0;
    }
    /* thread.rs:144-151: --downsample applies to a region with at least downsample_depth fragments (k_downsample marked the sampled ones) */
    const bool apply_ds = (a.P.flags & LCR_FLAG_DOWNSAMPLE) && a.P.downsample_depth > 0 && x.nf >= a.P.downsample_depth && a.ds;
    x.need = 3;
    x.apply_ds = apply_ds;
    for (uint32_t k = x.tid; k < x.nf; k += x.nthreads) {
        x.tag[k] = 0;
        x.assign[k] = 0;
        x.fp[k] = (uint8_t)((x.frag_links[k] >= a.P.min_linkers ? 1 : 0) | ((!apply_ds || a.ds[x.fb + k]) ? 2 : 0));
    }
    tsync(x);
    uint64_t n_calls;
    bool counted_elsewhere = false;
    GP_T(gr0);
    if (x.n <= a.P.max_enum_snps) { counted_elsewhere = phase_enum(x); n_calls = 1ull << x.n; }
    else { phase_ld(x); n_calls = 1ull + 2ull * (x.n / 4 + 1); }
    GP_T(gr1);
    GP_ADD(7, gr0, gr1);
    for (uint32_t i = x.tid; i < x.n; i += x.nthreads) { x.c[i].haplotype = x.st[i].x; x.c[i].genotype = x.st[i].y; }
    tsync(x);
    /* thread.rs:168-201 */
    assign_reads(x, false);
    assign_snps(x);
    assign_reads(x, false);
    assign_snps(x);
    rescue(x, LCR_CF_EDIT_LIST, false);
    rescue(x, LCR_CF_SOMATIC_LIST, true);
    x.need = 1; /* thread.rs:181-182: the last round runs over every fragment used for phasing */
    assign_reads(x, true);
    assign_snps(x);
    phase_sets(x);
    if (x.tid == 0 && !counted_elsewhere) {
        atomicAdd((unsigned long long *)&a.stats->n_cross_optimize, (unsigned long long)n_calls);
        atomicAdd((unsigned long long *)&a.stats->n_sweep_iters, x.n_iters);
    }
}

/* one CTA per region (everything except the few regions handed to the cooperative kernel) */
__global__ void __launch_bounds__(PB) k_phase(PhaseArgs a, int which) {
    __shared__ long long sh[32];
    __shared__ long long tabs[3][32];
    __shared__ TeamBcast bc;
    const uint32_t reg = blockIdx.x;
    if (a.ctr->overflow) return; /* a capacity was exceeded: the run is repeated with larger buffers */
    const LcrRegionState rs = a.rstate[reg];
    if (rs.status != 0) return;
    /* which: 0 every region; 1 the regions the enumeration search does not cover (they can run beside it); 2 the ones it covers */
    if (which && a.es_base && ((a.es_base[reg + 1] > a.es_base[reg]) != (which == 2))) return;
    if (rs.n_cand == 0) { /* phase() still runs its single (empty) configuration: one cross_optimize call of one iteration (phase.rs:1097-1122) */
        if (threadIdx.x == 0) {
            atomicAdd((unsigned long long *)&a.stats->n_cross_optimize, 1ull);
            atomicAdd((unsigned long long *)&a.stats->n_sweep_iters, 1ull);
        }
        return;
    }
    if (rs.n_cand > a.P.max_enum_snps && rs.n_frag >= a.big_frag_threshold) return; /* k_phase_grid takes it */
    Ctx x{a, *a.tables};
    x.reg = reg; x.tid = threadIdx.x; x.nthreads = PB; x.grid = false; x.bc = &bc; x.sh = sh;
    load_tables(a, tabs);
    x.OK = tabs[0]; x.ERR = tabs[1]; x.W = tabs[2];
    run_region(x);
}

/* the whole GPU on one large region: every sweep is a grid-wide pass over the fragment matrix in L2,
   grid.sync() between passes, no host round trips (SURVEY.md section 7 "ragged work", BASELINE config 5) */
__global__ void __launch_bounds__(PBG) k_phase_grid(PhaseArgs a, const uint32_t *big_list, uint32_t n_big_list, TeamBcast *gbc) {
    __shared__ long long sh[32];
    __shared__ long long tabs[3][32];
    Ctx x{a, *a.tables};
    load_tables(a, tabs);
    x.OK = tabs[0]; x.ERR = tabs[1]; x.W = tabs[2];
    x.tid = blockIdx.x * blockDim.x + threadIdx.x; x.nthreads = gridDim.x * blockDim.x; x.grid = true; x.bc = gbc; x.sh = sh;
    if (a.ctr->overflow) return; /* uniform over the grid: a capacity was exceeded, the run is repeated with larger buffers */
    /* the regions that can be this large are known from their read counts at upload; whether one is decided here, by
       the same test k_phase uses to leave it alone */
    for (uint32_t bi = 0; bi < n_big_list; ++bi) {
        const uint32_t reg = big_list[bi];
        const LcrRegionState rs = a.rstate[reg];
        if (rs.status != 0 || rs.n_cand <= a.P.max_enum_snps || rs.n_frag < a.big_frag_threshold) continue;
        if (x.tid == 0) { TeamBcast z{}; *gbc = z; }
        cg::this_grid().sync();
        x.reg = reg;
        GP_T(gq0);
        run_region(x);
        GP_T(gq1);
        GP_ADD(6, gq0, gq1);
        cg::this_grid().sync();
    }
}

} // namespace

/* downsample_fragments (phase.rs:693-701): one CTA per region that downsamples; thread 0 runs the reference's seeded Fisher-Yates shuffle
   (sequential by construction: the position in the ChaCha12 stream depends on the rejections before it) over an index array in shared
   memory (regions of up to DS_SMEM_IDX fragments) or in global scratch, then the CTA marks the first downsample_depth fragments */
#define DS_SMEM_IDX 51200u
__global__ void __launch_bounds__(256) k_downsample(PhaseArgs a, uint32_t *scratch) {
    extern __shared__ uint32_t ds_idx_smem[];
    const uint32_t reg = blockIdx.x;
    if (a.ctr->overflow) return;
    const LcrRegionState rs = a.rstate[reg];
    const uint32_t depth = a.P.downsample_depth, nf = rs.n_frag;
    if (rs.status != 0 || depth == 0 || nf < depth) return;
    uint32_t *idx = nf <= DS_SMEM_IDX ? ds_idx_smem : scratch + rs.frag_begin;
    for (uint32_t i = threadIdx.x; i < nf; i += blockDim.x) idx[i] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
        lcr_chacha12 g;
        lcr_stdrng_seed_from_u64(&g, 2025ull); /* thread.rs:149 */
        lcr_stdrng_shuffle(&g, idx, nf);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < depth; i += blockDim.x) a.ds[rs.frag_begin + idx[i]] = 1;
}

void lcr_launch_downsample(const PhaseArgs &a, uint32_t *scratch, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(k_downsample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(DS_SMEM_IDX * 4)); attr_set = true; }
    if (a.n_regions) k_downsample<<<a.n_regions, 256, DS_SMEM_IDX * 4, st>>>(a, scratch);
}

void lcr_launch_phase(const PhaseArgs &a, int which, cudaStream_t st) {
    if (a.n_regions) k_phase<<<a.n_regions, PB, 0, st>>>(a, which);
}

int lcr_launch_phase_grid(const PhaseArgs &a, const uint32_t *big_list, uint32_t n_big_list, void *bcast_scratch, int sm_count, cudaStream_t st) {
    static int occ_cached = 0;
    int occ = occ_cached;
    cudaError_t e = cudaSuccess;
    if (!occ) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_phase_grid, PBG, 0);
        if (e != cudaSuccess) return (int)e;
        if (occ < 1) occ = 1;
        if (occ > 2) occ = 2;
        occ_cached = occ;
    }
    PhaseArgs args = a;
    TeamBcast *gbc = reinterpret_cast<TeamBcast *>(bcast_scratch);
    void *params[] = {&args, &big_list, &n_big_list, &gbc};
    e = cudaLaunchCooperativeKernel((void *)k_phase_grid, dim3((unsigned)(sm_count * occ)), dim3(PBG), params, 0, st);
    return (int)e;
}
size_t lcr_phase_bcast_bytes() { return sizeof(TeamBcast); }
