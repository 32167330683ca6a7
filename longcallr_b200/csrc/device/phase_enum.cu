/*
 * phase_enum.cu — the 2^n starting-haplotype enumeration of SNPFrag::phase (src/phase.rs:1097-1122),
 * one warp per configuration.
 *
 * Every configuration is an independent cross_optimize run (src/phase.rs:810-976) from its own
 * random haplotags, so the 2^n runs of a region spread over warps and CTAs (persistent kernels that
 * draw (region, chunk) work items from device-built lists, k_enum_plan).  The winner is the highest
 * objective, lowest index on ties (strict `>` in enumeration order); the CTA that finishes a region's
 * last chunk runs the winning configuration once more and leaves its final state (haplotypes,
 * genotypes, haplotags) for k_phase, which continues from there.
 *
 * The region's fragment matrix is staged once per CTA in shared memory as one 64-bit word per
 * read (n <= 10 sites x 6 bits: sign and capped quality + 1), haplotags are one bit per read per
 * warp, SNP state lives in registers.  With sigma, delta, p in {-1, +1}:
 *   read k flips to   sign( sum_i p_ki * delta_i * W_ki )      over its heterozygous phase sites
 *   site i compares   C_i +- delta_i * M_i  (M_i = sum_k p_ki * sigma_k * W_ki), R_i, V_i
 * where W = log10(1 - eps) - log10(eps) in fixed point and C, R, V are per-column constants.
 */
#include "lcr_frag.h"
#include "lcr_pipeline.h"
#include "lcr_async.h"

#define EMAXN 10
#define EW_MAX 8
#define NF_PRE 384 /* rows of the signed-term table (PRE bins: regions of at most that many fragments) */
/* PRE bins: the region's tile of the fragment matrix (CSR slice: row offsets, site indices, cells) is staged in shared memory by
   the TMA engine (1-D bulk copies, completion on an mbarrier) before the rows are built from it */
#define TILE_ELEMS (NF_PRE * EMAXN)
#define TILE_OFF_BYTES ((((NF_PRE + 1) * 4 + 32) + 15) & ~15)
#define TILE_SNP_BYTES (TILE_ELEMS * 4 + 32)
#define TILE_CELL_BYTES (TILE_ELEMS + 32)
#define TILE_BYTES (TILE_OFF_BYTES + TILE_SNP_BYTES + TILE_CELL_BYTES)

namespace {

__device__ __forceinline__ int lcr_enum_shape_for_dev(uint32_t n_cand) { return n_cand <= 2 ? 0 : (n_cand == 3 ? 1 : (n_cand == 4 ? 2 : 3)); }

struct EnumShared {
    long long W[32], OK[32], ERR[32];
    long long C[EMAXN], R[EMAXN], V[EMAXN];
    uint32_t cov[EMAXN];
    long long best_prob[EW_MAX];
    uint32_t best_cfg[EW_MAX];
    unsigned long long iters[EW_MAX];
    uint32_t work, last, win_cfg;
    uint32_t t_off, t_snp, t_cell; /* staged tile: byte offsets of the region's first row offset / site index / cell inside the staged windows */
};

/* EW warps per CTA, ECFG_PER_WARP configurations per warp (strided over the chunk so that warps stay balanced);
   small regions use small CTAs so that many of them share an SM */
template <int EW, int ECFG_PER_WARP, bool PRE>
__global__ void __launch_bounds__(EW * 32, EW == 8 ? 3 : 1) k_enum_search(PhaseArgs a, int bin, uint32_t nf_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ EnumShared S;
    __shared__ __align__(8) unsigned long long s_tile_bar;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t tile_par = 0;
    if (PRE) {
        if (tid == 0) { mbar_init(smem_u32(&s_tile_bar), 1); mbar_fence_init(); }
        __syncthreads();
    }
    /* the bin's work list is built on the device (k_enum_plan): CTAs draw (region, chunk) items from its ticket counter */
  for (;;) {
    __syncthreads();
    if (tid == 0) S.work = atomicAdd(&a.ctr->enum_ticket[bin], 1u);
    __syncthreads();
    if (S.work >= a.ctr->enum_cnt[bin] || a.ctr->overflow) return;
    const uint32_t witem = a.ctr->enum_off[bin] + S.work;
    const uint32_t reg = a.work_region[witem], chunk = a.work_chunk[witem];
    const LcrRegionState rs = a.rstate[reg];
    const uint32_t n = rs.n_cand, nf = rs.n_frag, cb = rs.cand_begin, fb = rs.frag_begin;
    const uint32_t words = (nf_cap + 31) / 32;
    unsigned long long *rows = reinterpret_cast<unsigned long long *>(smem_raw);
    uint32_t *rrel = reinterpret_cast<uint32_t *>(rows + nf_cap);
    uint32_t *wm = rrel + nf_cap;            /* per 32 rows: sites some phasing row of the word carries (warp-uniform skips) */
    uint32_t *sig = wm + words + warp * words;
    /* PRE (regions with few fragments): p * W of every cell, site-major, so the sweeps load a signed term instead of decoding it */
    long long *Tm = reinterpret_cast<long long *>(smem_raw + (((size_t)nf_cap * 12 + (size_t)(EW + 1) * words * 4 + 7) & ~(size_t)7));
    unsigned char *tile = smem_raw + (((size_t)(reinterpret_cast<unsigned char *>(Tm + (size_t)NF_PRE * EMAXN) - smem_raw) + 15) & ~(size_t)15); /* PRE: 16-byte aligned */
    const lcr_candidate *c = a.cand + cb;
    const LcrDeviceTables &T = *a.tables;
    const uint64_t region_key = lcr_region_key(a.regions[reg].tid, a.regions[reg].start);
    const uint32_t slot0 = a.slot_off[reg];

    if (tid < 32) {
        const uint32_t q = tid < 31 ? tid : 30;
        S.OK[tid] = T.fx_ok[q];
        S.ERR[tid] = T.fx_err[q];
        S.W[tid] = T.fx_ok[q] - T.fx_err[q];
    }
    if (tid < EMAXN) { S.C[tid] = 0; S.R[tid] = 0; S.V[tid] = 0; S.cov[tid] = 0; }
    for (uint32_t w = tid; w < words; w += EW * 32) wm[w] = 0;
    if (PRE) {
        if (tid == 0) { /* windows from the 16-byte boundary below the region's first row offset / element to the one above its last */
            const uint32_t e0 = a.frag_elem_off[fb], e1 = a.frag_elem_off[fb + nf];
            const uint32_t bar = smem_u32(&s_tile_bar), dst = smem_u32(tile);
            const unsigned char *g_off = reinterpret_cast<const unsigned char *>(a.frag_elem_off + fb);
            const unsigned char *g_snp = reinterpret_cast<const unsigned char *>(a.elem_snp + e0);
            const unsigned char *g_cell = reinterpret_cast<const unsigned char *>(a.elem_cell + e0);
            const uint32_t o_off = (uint32_t)((uintptr_t)g_off & 15u), o_snp = (uint32_t)((uintptr_t)g_snp & 15u), o_cell = (uint32_t)((uintptr_t)g_cell & 15u);
            const uint32_t b_off = (o_off + 4u * (nf + 1u) + 15u) & ~15u, b_snp = (o_snp + 4u * (e1 - e0) + 15u) & ~15u, b_cell = (o_cell + (e1 - e0) + 15u) & ~15u;
            S.t_off = o_off; S.t_snp = TILE_OFF_BYTES + o_snp; S.t_cell = TILE_OFF_BYTES + TILE_SNP_BYTES + o_cell;
            bulk_g2s(dst, g_off - o_off, b_off, bar);
            if (e1 > e0) {
                bulk_g2s(dst + TILE_OFF_BYTES, g_snp - o_snp, b_snp, bar);
                bulk_g2s(dst + TILE_OFF_BYTES + TILE_SNP_BYTES, g_cell - o_cell, b_cell, bar);
            }
            mbar_arrive_expect_tx(bar, b_off + (e1 > e0 ? b_snp + b_cell : 0u));
        }
    }
    __syncthreads();
    if (PRE) {
        mbar_wait(smem_u32(&s_tile_bar), tile_par);
        tile_par ^= 1u;
    }
    const uint32_t *t_off = reinterpret_cast<const uint32_t *>(tile + S.t_off);
    const uint32_t *t_snp = reinterpret_cast<const uint32_t *>(tile + S.t_snp);
    const int8_t *t_cell = reinterpret_cast<const int8_t *>(tile + S.t_cell);
    const uint32_t t_e0 = PRE ? t_off[0] : 0u;
    /* stage the rows: bit 63 = fragment takes part in the sweeps (used for phasing and, when the region downsamples, inside the sampled
       set), bit 62 = used for phasing at all, 6 bits per site: (q + 1) | 32 when p < 0 */
    const bool apply_ds = (a.P.flags & LCR_FLAG_DOWNSAMPLE) && a.P.downsample_depth > 0 && nf >= a.P.downsample_depth && a.ds;
    for (uint32_t k = tid; k < nf; k += EW * 32) {
        const uint32_t f = fb + k;
        unsigned long long row = 0ull;
        if (a.frag_links[f] >= a.P.min_linkers) row = (1ull << 62) | ((!apply_ds || a.ds[f]) ? (1ull << 63) : 0ull);
        uint32_t present = 0;
        if (PRE) {
#pragma unroll
            for (int i = 0; i < EMAXN; ++i) Tm[i * NF_PRE + k] = 0;
        }
        const uint32_t ea = PRE ? t_off[k] - t_e0 : a.frag_elem_off[f], eb = PRE ? t_off[k + 1] - t_e0 : a.frag_elem_off[f + 1];
        for (uint32_t e = ea; e < eb; ++e) {
            const int8_t cell = PRE ? t_cell[e] : a.elem_cell[e];
            const uint32_t snp = PRE ? t_snp[e] : a.elem_snp[e];
            const uint32_t code = cell > 0 ? (uint32_t)cell : ((uint32_t)(-cell) | 32u);
            row |= (unsigned long long)code << (6 * snp);
            if (code) present |= 1u << snp;
            if (PRE && code) Tm[snp * NF_PRE + k] = cell > 0 ? S.W[(code & 31u) - 1u] : -S.W[(code & 31u) - 1u];
        }
        rows[k] = row;
        if ((row >> 63) && present) atomicOr(&wm[k >> 5], present);
        rrel[k] = a.frag_slot[f] - slot0;
    }
    if (PRE) { /* the table is read in whole words of 32 rows: rows past the last fragment hold zeros */
        for (uint32_t k = nf + tid; k < ((nf + 31u) & ~31u); k += EW * 32) {
#pragma unroll
            for (int i = 0; i < EMAXN; ++i) Tm[i * NF_PRE + k] = 0;
        }
    }
    uint32_t phase0 = 0; /* for_phasing mask */
    uint32_t het0 = 0, pos0 = 0xffffffffu; /* init_genotype, phase.rs:682-691: type 0 -> eta 1, type 1 -> 0, else -1 */
#pragma unroll
    for (int i = 0; i < EMAXN; ++i) {
        if ((uint32_t)i < n) {
            if (c[i].flags & LCR_CF_FOR_PHASING) phase0 |= 1u << i;
            const int vt = c[i].variant_type;
            if (vt == 1) het0 |= 1u << i;
            if (vt != 0) pos0 &= ~(1u << i);
        }
    }
    __syncthreads();
    /* column constants over the fragments used for phasing: C = sum(ok + err), R = homref sum, V = homvar sum */
    for (uint32_t i = warp; i < n; i += EW) {
        long long C = 0, R = 0, V = 0;
        uint32_t cov = 0;
        for (uint32_t k = lane; k < nf; k += 32) {
            const unsigned long long row = rows[k];
            const uint32_t code = (uint32_t)(row >> (6 * i)) & 63u;
            if (!(row >> 63) || !code) continue;
            const uint32_t q = (code & 31u) - 1u;
            const bool neg = code & 32u;
            C += S.OK[q] + S.ERR[q];
            R += neg ? S.ERR[q] : S.OK[q];
            V += neg ? S.OK[q] : S.ERR[q];
            cov++;
        }
        for (int o = 16; o; o >>= 1) {
            C += __shfl_xor_sync(0xffffffffu, C, o);
            R += __shfl_xor_sync(0xffffffffu, R, o);
            V += __shfl_xor_sync(0xffffffffu, V, o);
            cov += __shfl_xor_sync(0xffffffffu, cov, o);
        }
        if (lane == 0) { S.C[i] = C; S.R[i] = R; S.V[i] = V; S.cov[i] = cov; }
    }
    __syncthreads();

    const uint32_t n_cfg = 1u << n;
    long long best_prob = 0;
    uint32_t best_cfg = 0xffffffffu;
    unsigned long long iters = 0;
    const uint32_t nwords = (nf + 31) / 32;
    /* one cross_optimize run (phase.rs:810-976) from configuration cfg by this warp; leaves the final haplotags in sig[] and
       the final site state in (dneg, eta_het, eta_pos) */
    uint32_t dneg = 0, eta_het = 0, eta_pos = 0;
    auto run_cfg = [&](uint32_t cfg) -> long long {
        dneg = cfg; /* bit i set: delta_i = -1 */
        eta_het = het0; eta_pos = pos0; /* bit i: eta_i == 0 / eta_i == +1 (neither: -1) */
        /* init_assignment (phase.rs:673-680) */
        for (uint32_t w = 0; w < nwords; ++w) {
            const uint32_t k = w * 32 + lane;
            bool neg = false;
            if (k < nf && (rows[k] >> 63)) neg = lcr_uniform(a.P.seed, region_key, LCR_RNG_INIT_SIGMA, cfg, rrel[k]) < 0.5;
            const uint32_t word = __ballot_sync(0xffffffffu, neg);
            if (lane == 0) sig[w] = word;
        }
        __syncwarp();
        long long M[EMAXN];
        bool hg_increase = true, ht_increase = true;
        int num_iters = 0;
        while (hg_increase | ht_increase) {
            ++iters;
#pragma unroll
            for (int i = 0; i < EMAXN; ++i) M[i] = 0;
            bool any_flip = false;
            const uint32_t hetmask = eta_het & phase0;
            for (uint32_t w = 0; w < nwords; ++w) {
                const uint32_t k = w * 32 + lane;
                const uint32_t word = sig[w];
                const uint32_t wsites = wm[w];                 /* the same for the whole warp */
                const uint32_t vs = wsites & hetmask, ms = wsites & phase0;
                bool neg = (word >> lane) & 1u;
                unsigned long long row = 0;
                if (k < nf) row = rows[k];
                const bool active = row >> 63;
                long long v = 0; /* sum_i p * delta * W over heterozygous phase sites */
                if (PRE) {
                    /* one load per (site of the word, row): the signed term p * W feeds the read's decision and the site's column sum */
                    const long long *Tk = Tm + k;
                    long long tv[EMAXN];
#pragma unroll
                    for (int i = 0; i < EMAXN; ++i) tv[i] = ((ms >> i) & 1u) ? Tk[i * NF_PRE] : 0ll;
#pragma unroll
                    for (int i = 0; i < EMAXN; ++i)
                        if ((vs >> i) & 1u) v += ((dneg >> i) & 1u) ? -tv[i] : tv[i];
                    if (active && v != 0) {
                        const bool nneg = v < 0;
                        if (nneg != neg) { any_flip = true; neg = nneg; }
                    }
                    if (active) {
#pragma unroll
                        for (int i = 0; i < EMAXN; ++i)
                            if ((ms >> i) & 1u) M[i] += neg ? -tv[i] : tv[i];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < EMAXN; ++i) {
                        if ((vs >> i) & 1u) {
                            const uint32_t code = (uint32_t)(row >> (6 * i)) & 63u;
                            if (code) {
                                const long long Wq = S.W[(code & 31u) - 1u];
                                const bool minus = ((code >> 5) ^ (dneg >> i)) & 1u;
                                v += minus ? -Wq : Wq;
                            }
                        }
                    }
                    if (active && v != 0) {
                        const bool nneg = v < 0;
                        if (nneg != neg) { any_flip = true; neg = nneg; }
                    }
#pragma unroll
                    for (int i = 0; i < EMAXN; ++i) {
                        if ((ms >> i) & 1u) {
                            const uint32_t code = (uint32_t)(row >> (6 * i)) & 63u;
                            if (active && code) {
                                const long long Wq = S.W[(code & 31u) - 1u];
                                const bool minus = ((code >> 5) & 1u) ^ (neg ? 1u : 0u);
                                M[i] += minus ? -Wq : Wq;
                            }
                        }
                    }
                }
                const uint32_t nword = __ballot_sync(0xffffffffu, neg);
                if (lane == 0) sig[w] = nword;
            }
            __syncwarp();
            any_flip = __any_sync(0xffffffffu, any_flip);
#pragma unroll
            for (int i = 0; i < EMAXN; ++i)
                if ((uint32_t)i < n)
                    for (int o = 16; o; o >>= 1) M[i] += __shfl_xor_sync(0xffffffffu, M[i], o);
            if (!any_flip) ht_increase = false;
            else { ht_increase = true; hg_increase = true; }
            /* delta / eta sweep with_genotype (phase.rs:905-921): all lanes hold the same sums, lane i decides site i */
            bool better = false, flip = false;
            bool nhet = (eta_het >> lane) & 1u, npos = (eta_pos >> lane) & 1u;
            if (lane < n && ((phase0 >> lane) & 1u) && S.cov[lane]) {
                long long Mi = M[0];
#pragma unroll
                for (int i = 1; i < EMAXN; ++i) Mi = (int)lane == i ? M[i] : Mi;
                const long long dM = ((dneg >> lane) & 1u) ? -Mi : Mi;
                const long long ph2 = 2 * (T.fx_prior_het - (long long)S.cov[lane] * T.fx_log10_2);
                const long long L0 = S.C[lane] + dM + ph2, L1 = S.C[lane] - dM + ph2;
                const long long L2 = 2 * (S.R[lane] + T.fx_prior_homref), L3 = 2 * (S.V[lane] + T.fx_prior_homvar);
                const long long L_old = nhet ? L0 : (npos ? L2 : L3);
                long long mx = L0 > L1 ? L0 : L1;
                const long long m2 = L2 > L3 ? L2 : L3;
                mx = mx > m2 ? mx : m2;
                long long L_new;
                if (L0 == mx) { nhet = true; npos = false; L_new = L0; }
                else if (L1 == mx) { flip = true; nhet = true; npos = false; L_new = L1; }
                else if (L2 == mx) { nhet = false; npos = true; L_new = L2; }
                else { nhet = false; npos = false; L_new = L3; }
                better = L_new > L_old;
            }
            dneg ^= __ballot_sync(0xffffffffu, flip);
            eta_het = __ballot_sync(0xffffffffu, nhet);
            eta_pos = __ballot_sync(0xffffffffu, npos);
            better = __any_sync(0xffffffffu, better);
            if (!better) hg_increase = false;
            else { hg_increase = true; ht_increase = true; }
            if (++num_iters > 20) break;
        }
        /* cal_overall_probability from the column sums of the final state */
        long long prob2 = 0;
        if (lane < n && ((phase0 >> lane) & 1u) && S.cov[lane]) {
            long long Mi = M[0];
#pragma unroll
            for (int i = 1; i < EMAXN; ++i) Mi = (int)lane == i ? M[i] : Mi;
            const long long dM = ((dneg >> lane) & 1u) ? -Mi : Mi;
            prob2 = ((eta_het >> lane) & 1u) ? (S.C[lane] + dM) : (((eta_pos >> lane) & 1u) ? 2 * S.R[lane] : 2 * S.V[lane]);
        }
        for (int o = 16; o; o >>= 1) prob2 += __shfl_xor_sync(0xffffffffu, prob2, o);
        return prob2 / 2;
    };
    for (uint32_t j = 0; j < ECFG_PER_WARP; ++j) {
        const uint32_t cfg = chunk * (EW * ECFG_PER_WARP) + j * EW + warp;
        if (cfg >= n_cfg) break;
        const long long prob = run_cfg(cfg);
        if (best_cfg == 0xffffffffu || prob > best_prob) { best_prob = prob; best_cfg = cfg; }
    }
    if (lane == 0) { S.best_prob[warp] = best_prob; S.best_cfg[warp] = best_cfg; S.iters[warp] = iters; }
    __syncthreads();
    const uint32_t es0 = a.es_base[reg], n_chunks = a.es_base[reg + 1] - es0;
    if (tid == 0) {
        long long bp = 0;
        uint32_t bc = 0xffffffffu;
        unsigned long long it = 0;
        uint32_t ncfg_done = 0;
        for (int w = 0; w < EW; ++w) {
            it += S.iters[w];
            if (S.best_cfg[w] == 0xffffffffu) continue;
            if (bc == 0xffffffffu || S.best_prob[w] > bp || (S.best_prob[w] == bp && S.best_cfg[w] < bc)) { bp = S.best_prob[w]; bc = S.best_cfg[w]; }
        }
        const uint32_t first = chunk * (EW * ECFG_PER_WARP);
        ncfg_done = n_cfg > first ? (n_cfg - first < EW * ECFG_PER_WARP ? n_cfg - first : EW * ECFG_PER_WARP) : 0;
        a.es_prob[es0 + chunk] = bp;
        a.es_cfg[es0 + chunk] = bc;
        atomicAdd((unsigned long long *)&a.stats->n_sweep_iters, it);
        atomicAdd((unsigned long long *)&a.stats->n_cross_optimize, (unsigned long long)ncfg_done);
        /* the CTA that finishes the region's last chunk materialises the winner: its rows are already staged here */
        __threadfence();
        S.last = atomicAdd(&a.es_done[reg], 1u) + 1u == n_chunks ? 1u : 0u;
        if (S.last) {
            __threadfence();
            long long wp = 0;
            uint32_t wc = 0xffffffffu;
            for (uint32_t w = 0; w < n_chunks; ++w) { /* strict `>` in enumeration order (phase.rs:1108-1122): highest objective, lowest index on ties */
                const uint32_t cw = *(volatile uint32_t *)&a.es_cfg[es0 + w];
                const long long pw = *(volatile long long *)&a.es_prob[es0 + w];
                if (cw == 0xffffffffu) continue;
                if (wc == 0xffffffffu || pw > wp || (pw == wp && cw < wc)) { wp = pw; wc = cw; }
            }
            S.win_cfg = wc;
        }
    }
    __syncthreads();
    if (S.last && warp == 0 && S.win_cfg != 0xffffffffu) {
        run_cfg(S.win_cfg);
        /* final state of the winning run: what phase() leaves in the candidates and fragments before thread.rs:168 */
        if (lane < n) {
            a.best_hap[cb + lane] = ((dneg >> lane) & 1u) ? -1 : 1;
            a.best_gen[cb + lane] = ((eta_het >> lane) & 1u) ? 0 : (((eta_pos >> lane) & 1u) ? 1 : -1);
        }
        __syncwarp();
        for (uint32_t w = 0; w < nwords; ++w) {
            const uint32_t k = w * 32 + lane;
            if (k < nf) {
                int8_t t = 0;
                if (rows[k] >> 63) t = ((sig[w] >> lane) & 1u) ? -1 : 1;
                else if ((rows[k] >> 62) & 1u) /* used for phasing but outside the sampled set: it keeps the haplotag init_assignment drew for the winner */
                    t = lcr_uniform(a.P.seed, region_key, LCR_RNG_INIT_SIGMA, S.win_cfg, rrel[k]) < 0.5 ? -1 : 1;
                a.best_tag[fb + k] = t;
            }
        }
        if (lane == 0) a.es_done[reg] = 0x80000000u; /* replayed: k_phase loads the state instead of running the configuration again */
    }
  }
}

/* ---- work lists of the enumeration search, built on the device ----
   One CTA walks the regions: every region with at most min(max_enum_snps, 10) candidates and at most 16384 fragments gets a
   launch shape by its number of configurations and a class by its fragment count (shared-memory footprint); its 2^n
   configurations are cut into chunks of one CTA each.  Output: es_base (first result slot of every region), the
   per-bin (region, chunk) lists and their sizes in the counter block. */
#define EP_THREADS 1024
__global__ void __launch_bounds__(EP_THREADS) k_enum_plan(PhaseArgs a, uint32_t work_cap, int sm_count) {
    __shared__ unsigned long long s_big;
    __shared__ uint32_t s_cnt[LCR_ENUM_BINS], s_off[LCR_ENUM_BINS], s_cur[LCR_ENUM_BINS];
    __shared__ uint32_t s_scan[EP_THREADS / 32], s_carry, s_ovf;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_regions = a.n_regions;
    if (tid == 0) { s_big = 0; s_carry = 0; s_ovf = 0; }
    if (tid < LCR_ENUM_BINS) { s_cnt[tid] = 0; s_cur[tid] = 0; }
    __syncthreads();
    auto eligible = [&](const LcrRegionState &rs) {
        return rs.status == 0 && rs.n_cand && rs.n_cand <= a.P.max_enum_snps && rs.n_cand <= 10u && rs.n_frag <= 16384u;
    };
    /* regions with 5+ sites: 64 configurations per CTA when that still fills the GPU, else 16 (more, shorter CTAs) */
    unsigned long long big = 0;
    for (uint32_t r = tid; r < n_regions; r += EP_THREADS) {
        const LcrRegionState rs = a.rstate[r];
        if (eligible(rs) && rs.n_cand > 4) big += 1ull << rs.n_cand;
    }
    if (big) atomicAdd(&s_big, big);
    __syncthreads();
    const bool small_batch = s_big / 64 < 8ull * (unsigned long long)(sm_count > 0 ? sm_count : 148);
    auto plan = [&](const LcrRegionState &rs, int &bin) -> uint32_t {
        bin = -1;
        if (!eligible(rs)) return 0;
        int shape = lcr_enum_shape_for_dev(rs.n_cand);
        if (shape == 3 && small_batch) shape = 4;
        const uint32_t per_cta = shape == 0 ? 4u : shape == 1 ? 8u : shape == 2 ? 16u : shape == 3 ? 64u : 16u;
        const int cls = rs.n_frag <= 384u ? 0 : rs.n_frag <= 1024u ? 1 : rs.n_frag <= 4096u ? 2 : 3;
        bin = shape * LCR_ENUM_CLASSES + cls;
        return ((1u << rs.n_cand) + per_cta - 1) / per_cta;
    };
    /* pass 1: chunks per region -> es_base (exclusive scan in region order) and the bin sizes */
    for (uint32_t base = 0; base < n_regions; base += EP_THREADS) {
        const uint32_t r = base + tid;
        uint32_t chunks = 0;
        int bin = -1;
        if (r < n_regions) chunks = plan(a.rstate[r], bin);
        if (bin >= 0) atomicAdd(&s_cnt[bin], chunks);
        uint32_t incl = chunks;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += v;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        uint32_t wb = s_carry;
        for (uint32_t w = 0; w < warp; ++w) wb += s_scan[w];
        if (r < n_regions) a.es_base[r] = wb + incl - chunks;
        __syncthreads();
        if (tid == EP_THREADS - 1) s_carry = wb + incl;
        __syncthreads();
    }
    if (tid == 0) {
        a.es_base[n_regions] = s_carry;
        uint32_t off = 0;
        for (int b = 0; b < LCR_ENUM_BINS; ++b) { s_off[b] = off; off += s_cnt[b]; }
        a.ctr->enum_work = off;
        if (off > work_cap) { atomicOr(&a.ctr->overflow, LCR_OVF_ENUM); s_ovf = 1; }
        for (int b = 0; b < LCR_ENUM_BINS; ++b) { a.ctr->enum_off[b] = s_off[b]; a.ctr->enum_cnt[b] = s_ovf ? 0u : s_cnt[b]; a.ctr->enum_ticket[b] = 0; }
    }
    __syncthreads();
    if (s_ovf) return;
    /* pass 2: the (region, chunk) items of every bin */
    for (uint32_t r = tid; r < n_regions; r += EP_THREADS) {
        int bin;
        const uint32_t chunks = plan(a.rstate[r], bin);
        if (bin < 0) continue;
        const uint32_t p0 = s_off[bin] + atomicAdd(&s_cur[bin], chunks);
        for (uint32_t ck = 0; ck < chunks; ++ck) { a.work_region[p0 + ck] = r; a.work_chunk[p0 + ck] = ck; }
    }
}

} // namespace

/* launch shapes by number of configurations: {warps per CTA, configurations per warp} */
static const int SHAPES[LCR_ENUM_SHAPES][2] = {{1, 4}, {2, 4}, {4, 4}, {8, 8}, {8, 2}}; /* the last: 5+ sites when the batch is too small to fill the GPU with 64 configurations per CTA */
static const uint32_t CLASS_ROWS[LCR_ENUM_CLASSES] = {NF_PRE, 1024, 4096, 16384};            /* fragment-count classes: rows staged in shared memory */

static size_t smem_bytes(uint32_t nf_cap, int ew, bool pre) { return (size_t)nf_cap * 12 + (size_t)(ew + 1) * ((nf_cap + 31) / 32) * 4 + 16 + (pre ? (size_t)NF_PRE * 8 * EMAXN + 32 + TILE_BYTES : 0); }

template <int EW, int CPW, bool PRE>
static int launch_shape(const PhaseArgs &a, int bin, uint32_t nf_cap, int grid, cudaStream_t st) {
    const size_t smem = smem_bytes(nf_cap, EW, PRE);
    static size_t smem_set[8] = {0, 0, 0, 0, 0, 0, 0, 0}; /* per device: the attribute is set once per instantiation and size, not per launch */
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 8 || smem_set[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(k_enum_search<EW, CPW, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 8) smem_set[dev] = smem;
    }
    k_enum_search<EW, CPW, PRE><<<grid, EW * 32, smem, st>>>(a, bin, nf_cap);
    return (int)cudaGetLastError();
}

void lcr_launch_enum_plan(const PhaseArgs &a, uint32_t work_cap, int sm_count, cudaStream_t st) {
    k_enum_plan<<<1, EP_THREADS, 0, st>>>(a, work_cap, sm_count);
}

/* one persistent launch per (shape, class) bin; bins without work cost one ticket per CTA.  pre: regions with few fragments
   (class 0) of the 5+ site shapes keep the signed-term table in shared memory */
bool lcr_enum_bin_possible(int bin, uint32_t max_region_reads) {
    const int cls = bin % LCR_ENUM_CLASSES;
    return cls == 0 || max_region_reads > CLASS_ROWS[cls - 1]; /* a region has at most as many fragments as reads */
}

int lcr_launch_enum_search(int bin, const PhaseArgs &a, int sm_count, cudaStream_t st) {
    const int shape = bin / LCR_ENUM_CLASSES, cls = bin % LCR_ENUM_CLASSES;
    const uint32_t nf_cap = CLASS_ROWS[cls];
    const bool pre = cls == 0;
    const int sms = sm_count > 0 ? sm_count : 148;
    switch (shape) {
        case 0: return launch_shape<1, 4, false>(a, bin, nf_cap, sms * 16, st);
        case 1: return launch_shape<2, 4, false>(a, bin, nf_cap, sms * 16, st);
        case 2: return launch_shape<4, 4, false>(a, bin, nf_cap, sms * 8, st);
        case 3: return pre ? launch_shape<8, 8, true>(a, bin, nf_cap, sms * 4, st) : launch_shape<8, 8, false>(a, bin, nf_cap, sms * 4, st);
        default: return pre ? launch_shape<8, 2, true>(a, bin, nf_cap, sms * 4, st) : launch_shape<8, 2, false>(a, bin, nf_cap, sms * 4, st);
    }
}
