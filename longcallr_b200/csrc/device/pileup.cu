/*
 * pileup.cu — read filter, tile work items, fused pileup + site genotyping, candidate lists.
 *
 * Replaces, on the device:
 *   src/util.rs:636-668      read filter and fetch window             (k_slot_prep)
 *   src/util.rs:650-948      Profile::fill_data_into_freq_vec         (k_pileup_tile, phase 1 + 2)
 *   src/util.rs:162-176      BaseFreq::get_two_major_alleles          (site_call)
 *   src/candidate.rs:75-463  filter cascade, genotype likelihood      (site_call)
 *   src/candidate.rs:465-526 dense-cluster filters                    (k_cand_finalize)
 *
 * Layout: reads are decomposed into (read, tile) items on the device; one CTA owns one
 * tile of LCR_TILE reference positions, stages LCR_ROWS reads at a time as one byte per
 * position in shared memory ((q << 3) | code), and every thread accumulates the column of
 * its own position in registers: no atomics on the counters, no per-position record in HBM.
 * Only candidate sites (and, on request, the debug planes) are written out.
 */
#include <cub/cub.cuh>

#include "lcr_device.h"

namespace {

struct PrepArgs {
    lcr_params P;
    uint32_t n_slots;
    const lcr_region *regions;
    const uint32_t *slot_off, *slot_region, *tile_base;
    const int32_t *pos;
    const uint16_t *flag;
    const uint8_t *mapq;
    const float *de;
    const uint64_t *seq_off, *cig_off;
    const uint32_t *cigar;
    LcrRegionState *rstate;
    uint8_t *slot_flags;
    uint32_t *tile_count;  /* COUNT: items per tile; FILL: cursor */
    const uint32_t *tile_off;
    uint32_t *tile_full_n; /* whole-tile intron covers */
    LcrItem *items;
};

__device__ __forceinline__ bool is_ref_consuming(uint32_t opc) { return opc == 0 || opc == 2 || opc == 3 || opc == 7 || opc == 8; }

template <bool FILL>
__global__ void k_slot_prep(PrepArgs a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= a.n_slots) return;
    const uint32_t reg = a.slot_region[slot];
    if (a.rstate[reg].status != 0) return;
    const lcr_region R = a.regions[reg];
    const uint32_t read = R.read_begin + (slot - a.slot_off[reg]);
    const uint64_t c0 = a.cig_off[read], c1 = a.cig_off[read + 1];
    if (!FILL) {
        /* util.rs:652-668 */
        const uint64_t l_seq = a.seq_off[read + 1] - a.seq_off[read];
        const uint16_t fl = a.flag[read];
        bool pass = !((int32_t)a.mapq[read] < a.P.min_mapq || l_seq < (uint64_t)a.P.min_read_length || (fl & 0x4) || (fl & 0x100) || (fl & 0x800));
        const float de = a.de[read];
        if (!(de != de) && de >= a.P.divergence) pass = false;
        /* fetch((chr, start, end)): pos < end && bam_endpos > start on the region's own numbers */
        int64_t rlen = 0;
        for (uint64_t c = c0; c < c1; ++c) {
            const uint32_t op = a.cigar[c];
            if (is_ref_consuming(op & 0xf)) rlen += op >> 4;
        }
        const int64_t p = a.pos[read];
        const bool in_window = p < (int64_t)R.end && p + (rlen ? rlen : 1) > (int64_t)R.start;
        a.slot_flags[slot] = (pass && in_window) ? 1 : 0;
        if (!(pass && in_window)) return;
    } else if (!a.slot_flags[slot]) return;

    const int64_t vec_size = (int64_t)R.end - (int64_t)R.start;
    const int64_t fv_start = (int64_t)R.start - 1;
    const uint32_t tb = a.tile_base[reg];
    int64_t fpos = (int64_t)a.pos[read] - fv_start;
    uint32_t rpos = (c1 > c0 && (a.cigar[c0] & 0xf) == 4) ? (a.cigar[c0] >> 4) : 0; /* leading_softclips */
    int64_t last_tile = -1;
    for (uint64_t c = c0; c < c1; ++c) {
        const uint32_t op = a.cigar[c], opc = op & 0xf, len = op >> 4;
        if (opc == 4 || opc == 5) continue;
        if (opc == 1) {
            if (fpos >= vec_size && fpos >= 1) break;
            rpos += len;
            continue;
        }
        if (!is_ref_consuming(opc)) { /* util.rs:943-945 panics */
            atomicMin(&a.rstate[reg].status, (int32_t)LCR_ERR_BAD_CIGAR);
            return;
        }
        const bool is_m = (opc == 0 || opc == 7 || opc == 8);
        const int64_t lo = fpos, hi = fpos + (int64_t)len;
        if (lo >= vec_size) { /* the walk is over; later ops cannot reach the window */
            fpos = hi;
            continue;
        }
        if (hi > 0) {
            const int64_t a0 = lo < 0 ? 0 : lo, b0 = hi < vec_size ? hi : vec_size;
            for (int64_t t = a0 / LCR_TILE; t * LCR_TILE < b0; ++t) {
                const int64_t ts = a0 > t * LCR_TILE ? a0 : t * LCR_TILE;
                const int64_t tile_end = (t + 1) * LCR_TILE < vec_size ? (t + 1) * LCR_TILE : vec_size;
                const int64_t te = b0 < tile_end ? b0 : tile_end;
                if (opc == 3 && ts == t * LCR_TILE && te == tile_end && t > last_tile) {
                    if (FILL) atomicAdd(&a.tile_full_n[tb + t], 1u);
                    continue;
                }
                if (t > last_tile) {
                    last_tile = t;
                    if (!FILL) atomicAdd(&a.tile_count[tb + t], 1u);
                    else {
                        const uint32_t k = a.tile_off[tb + t] + atomicAdd(&a.tile_count[tb + t], 1u);
                        LcrItem it;
                        it.slot = slot;
                        it.cig = (uint32_t)(c - c0);
                        it.opoff = (uint32_t)(ts - lo);
                        it.rpos = is_m ? rpos + (uint32_t)(ts - lo) : rpos;
                        it.fpos = (int32_t)ts;
                        a.items[k] = it;
                    }
                }
            }
        }
        fpos = hi;
        if (is_m) rpos += len;
    }
}

/* ------------------------------------------------------------------------- */

struct PileArgs {
    lcr_params P;
    const lcr_region *regions;
    const uint32_t *slot_off, *slot_region, *tile_base, *tile_region;
    const uint64_t *pos_off;
    const uint16_t *flag;
    const int8_t *ts;
    const uint64_t *seq_off, *cig_off;
    const uint8_t *seq, *qual;
    const uint32_t *cigar;
    const uint8_t *const *ref_table;
    const uint32_t *tile_off, *tile_full_n;
    const LcrItem *items;
    const LcrDeviceTables *tables;
    LcrRegionState *rstate;
    lcr_candidate *cand;
    uint64_t *cand_key;
    uint32_t cand_cap;
    uint32_t *cand_count;
    lcr_stats *stats;
    uint32_t *pl_acgt, *pl_fwd, *pl_d, *pl_n, *pl_ts; /* debug planes or null */
    struct PreCand *pre;
    uint32_t pre_cap;
    uint32_t *pre_count;
};

__device__ __forceinline__ int base_code_dev(uint8_t b) {
    switch (b) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

/* util.rs:737-789: end trim (ONT) and poly-A / homopolymer mask near the clipped read ends.
   The reference scans the polya windows [ti, ti + polya), ti in [curr - polya, curr + 1], for one made
   of a single letter X in {A,T,C,G} different from the reference base.  Every such window contains
   curr - 1 or curr + 1, so it is enough to grow the homopolymer run around those two anchors inside
   [curr - polya, curr + polya] and compare its length with polya. */
__device__ __forceinline__ bool base_masked(const lcr_params &P, const uint8_t *seq, int64_t curr64, int64_t seq_len64, int64_t lead64, int64_t trail64, uint8_t ref_base) {
    const int32_t curr = (int32_t)curr64, seq_len = (int32_t)seq_len64, lead = (int32_t)lead64, trail = (int32_t)trail64;
    const int32_t dist_end = (int32_t)(P.distance_to_read_end > 0x3fffffffu ? 0x3fffffffu : P.distance_to_read_end);
    const int32_t d0 = curr - lead, d1 = curr - (seq_len - trail);
    const bool near_end = (d0 < 0 ? -d0 : d0) < dist_end || (d1 < 0 ? -d1 : d1) < dist_end;
    if (!near_end) return false;
    if (P.platform == 1) return true;
    const int32_t polya = (int32_t)(P.polya_tail_length > 0x3fffffffu ? 0x3fffffffu : P.polya_tail_length);
    if (polya < 2) { /* literal form for degenerate window lengths */
        for (int32_t ti = curr - polya; ti <= curr + 1; ++ti) {
            if (ti < 0 || ti + polya - 1 >= seq_len) continue;
            int32_t pa = 0, pt = 0, pc = 0, pg = 0;
            for (int32_t tj = 0; tj < polya; ++tj) {
                const uint8_t b = __ldg(seq + ti + tj);
                if (b == 'A' && ref_base != 'A') pa++;
                else if (b == 'T' && ref_base != 'T') pt++;
                else if (b == 'C' && ref_base != 'C') pc++;
                else if (b == 'G' && ref_base != 'G') pg++;
            }
            if (pa >= polya || pt >= polya || pc >= polya || pg >= polya) return true;
        }
        return false;
    }
    const int32_t lo = curr - polya > 0 ? curr - polya : 0;
    const int32_t hi = curr + polya + 1 < seq_len ? curr + polya + 1 : seq_len; /* exclusive */
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const int32_t anchor = side == 0 ? curr - 1 : curr + 1;
        if (anchor < lo || anchor >= hi) continue;
        const uint8_t X = __ldg(seq + anchor);
        if (!(X == 'A' || X == 'T' || X == 'C' || X == 'G') || X == ref_base) continue;
        int32_t s0 = anchor, e0 = anchor + 1;
        while (s0 > lo && __ldg(seq + s0 - 1) == X) --s0;
        while (e0 < hi && e0 - s0 < polya && __ldg(seq + e0) == X) ++e0;
        if (e0 - s0 >= polya) return true;
    }
    return false;
}

struct SiteCounters {
    uint32_t cnt[4], pass[4], fwd[4], ts[2], d, n;
    int64_t ll0, ll2;
    uint32_t q0flags; /* bit0: a reference-matching base of quality 0, bit1: a non-reference one */
};

/* candidate.rs:75-463 for one position; returns true and fills `o` when the site becomes a candidate */
template <bool PRE>
__device__ bool site_call(const lcr_params &P, const LcrDeviceTables &T, const SiteCounters &s, uint8_t ref_base, lcr_candidate &o) {
    const uint32_t total = s.cnt[0] + s.cnt[1] + s.cnt[2] + s.cnt[3];
    if (total < P.min_depth || total > P.max_depth) return false;
    /* get_two_major_alleles: stable descending sort of (A,C,G,T) */
    int ord[4] = {0, 1, 2, 3};
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        const int v = ord[i];
        int j = i - 1;
        while (j >= 0 && s.cnt[ord[j]] < s.cnt[v]) { ord[j + 1] = ord[j]; --j; }
        ord[j + 1] = v;
    }
    const char ACGT[4] = {'A', 'C', 'G', 'T'};
    int i1 = ord[0], i2 = ord[1];
    if ((uint8_t)ACGT[ord[0]] != ref_base && (uint8_t)ACGT[ord[1]] != ref_base) {
        if (s.cnt[ord[2]] == s.cnt[ord[1]] && (uint8_t)ACGT[ord[2]] == ref_base) i2 = ord[2];
        else if (s.cnt[ord[3]] == s.cnt[ord[1]] && (uint8_t)ACGT[ord[3]] == ref_base) i2 = ord[3];
    }
    const uint8_t allele1 = (uint8_t)ACGT[i1], allele2 = (uint8_t)ACGT[i2];
    const uint32_t allele1_cnt = s.cnt[i1], allele2_cnt = s.cnt[i2];
    const float allele1_freq = (float)allele1_cnt / (float)total;
    const float allele2_freq = (float)allele2_cnt / (float)total;
    uint8_t ref_allele_base;
    uint32_t alt_num;
    int alt_i[2] = {0, 0};
    float alt_freq[2] = {0.f, 0.f};
    uint32_t alt_cnt[2] = {0, 0};
    if (allele1 == ref_base) { ref_allele_base = allele1; alt_num = 1; alt_i[0] = i2; alt_freq[0] = allele2_freq; alt_cnt[0] = allele2_cnt; }
    else if (allele2 == ref_base) { ref_allele_base = allele2; alt_num = 1; alt_i[0] = i1; alt_freq[0] = allele1_freq; alt_cnt[0] = allele1_cnt; }
    else { ref_allele_base = ref_base; alt_num = 2; alt_i[0] = i1; alt_freq[0] = allele1_freq; alt_cnt[0] = allele1_cnt; alt_i[1] = i2; alt_freq[1] = allele2_freq; alt_cnt[1] = allele2_cnt; }
    const int ref_code = base_code_dev(ref_allele_base);
    if (ref_code < 0) return false;
    if (alt_num == 1) {
        if (total < 200 && alt_freq[0] < P.low_allele_frac_cutoff) return false;
        if (total >= 200 && alt_cnt[0] < P.low_allele_cnt_cutoff) return false;
    }
    if (s.d >= alt_cnt[0]) return false;
    const uint32_t depth_incl = total + s.d + s.n;
    if ((float)(allele1_cnt + allele2_cnt) / (float)depth_incl < P.min_allele_freq_include_intron) return false;
    if (allele1 != ref_base) { if (allele1_cnt > 0 && s.pass[i1] < 2) return false; }
    else if (allele2 != ref_base) { if (allele2_cnt > 0 && s.pass[i2] < 2) return false; }
    if (P.use_strand_bias) {
        const int32_t rf = (int32_t)s.fwd[ref_code], rr = (int32_t)(s.cnt[ref_code] - s.fwd[ref_code]);
        const int32_t af = (int32_t)s.fwd[alt_i[0]], ar = (int32_t)(s.cnt[alt_i[0]] - s.fwd[alt_i[0]]);
        float sor;
        if (alt_num == 1) sor = lcr_strand_odds_ratio(rf, rr, af, ar);
        else {
            const int32_t bf = (int32_t)s.fwd[alt_i[1]], br = (int32_t)(s.cnt[alt_i[1]] - s.fwd[alt_i[1]]);
            sor = fmaxf(lcr_strand_odds_ratio(rf, rr, af, ar), lcr_strand_odds_ratio(rf, rr, bf, br));
        }
        if (sor > T.sor_threshold) return false;
        if (alt_num == 1) {
            if (af + ar <= 30 && ((T.binom_reject[af + ar] >> af) & 1u)) return false;
            if ((int64_t)af * (int64_t)ar == 0) return false;
        }
    }
    if (!(ref_base == 'A' || ref_base == 'C' || ref_base == 'G' || ref_base == 'T')) return false;
    if (PRE) return true; /* count-based filters passed; the likelihood needs the per-base qualities */
    const double NEG_INF = lcr_u2d(0xfff0000000000000ULL);
    double ll[3];
    ll[0] = (s.q0flags & 2) ? NEG_INF : lcr_fx_to_f64(s.ll0);
    ll[2] = (s.q0flags & 1) ? NEG_INF : lcr_fx_to_f64(s.ll2);
    ll[1] = 0.0;
    ll[1] -= (double)total * T.log10_2;
    double lp[3] = {ll[0] + T.gl_prior_log[0], ll[1] + T.gl_prior_log[1], ll[2] + T.gl_prior_log[2]};
    const double max_lp = fmax(fmax(lp[0], lp[1]), lp[2]);
    lp[0] -= max_lp; lp[1] -= max_lp; lp[2] -= max_lp;
    double vp[3] = {lcr_exp10(lp[0]), lcr_exp10(lp[1]), lcr_exp10(lp[2])};
    const double sum_vp = vp[0] + vp[1] + vp[2];
    vp[0] /= sum_vp; vp[1] /= sum_vp; vp[2] /= sum_vp;
    const double variant_quality = -10.0 * lcr_log10(fmax(10e-301, vp[2]));
    const double max_ll = fmax(fmax(ll[0], ll[1]), ll[2]);
    double l10[3] = {lcr_exp10(ll[0] - max_ll), lcr_exp10(ll[1] - max_ll), lcr_exp10(ll[2] - max_ll)};
    const double sum_l10 = l10[0] + l10[1] + l10[2];
    const double gp[3] = {l10[0] / sum_l10, l10[1] / sum_l10, l10[2] / sum_l10};
    double ph[3] = {-10.0 * lcr_log10(gp[0]), -10.0 * lcr_log10(gp[1]), -10.0 * lcr_log10(gp[2])};
    if (ph[1] < ph[0]) { double t = ph[0]; ph[0] = ph[1]; ph[1] = t; }
    if (ph[2] < ph[1]) { double t = ph[1]; ph[1] = ph[2]; ph[2] = t; if (ph[1] < ph[0]) { double u = ph[0]; ph[0] = ph[1]; ph[1] = u; } }
    const double genotype_quality = ph[1] - ph[0];
    int variant_type, genotype;
    if (gp[0] > gp[1] && gp[0] > gp[2]) { variant_type = 2; genotype = -1; }
    else if (gp[1] > gp[0] && gp[1] > gp[2]) { variant_type = 1; genotype = 0; }
    else { variant_type = 0; genotype = 1; }
    if (variant_quality < (double)P.min_qual) return false;

    uint16_t fl = 0;
    const int32_t fwd_ts = (int32_t)s.ts[0], rev_ts = (int32_t)s.ts[1];
    const uint8_t alt0 = (uint8_t)ACGT[alt_i[0]];
    bool keep = true;
    if (ref_allele_base == 'A' && alt0 == 'G' && (fwd_ts > rev_ts * 2 || (fwd_ts == 0 && rev_ts == 0)) && variant_type != 2) fl = LCR_CF_RNA_EDITING | LCR_CF_EDIT_LIST;
    else if (ref_allele_base == 'T' && alt0 == 'C' && (rev_ts > fwd_ts * 2 || (fwd_ts == 0 && rev_ts == 0)) && variant_type != 2) fl = LCR_CF_RNA_EDITING | LCR_CF_EDIT_LIST;
    else if (alt_num == 1 && alt_freq[0] < P.min_allele_freq) fl = LCR_CF_CAND_SOMATIC | LCR_CF_SOMATIC_LIST;
    else if (variant_type == 2) {
        if (alt_num == 2 && alt_freq[0] >= P.min_allele_freq && alt_freq[1] >= P.min_allele_freq) { variant_type = 3; genotype = -1; }
        fl = LCR_CF_HOM_VAR | LCR_CF_FOR_PHASING;
    } else if (variant_type == 1) {
        if (alt_num == 2) { variant_type = 3; genotype = -1; fl = LCR_CF_HOM_VAR | LCR_CF_FOR_PHASING; }
        else fl = LCR_CF_HET_VAR | LCR_CF_FOR_PHASING;
    } else keep = false;
    if (!keep) return false;
    o.variant_quality = variant_quality;
    o.genotype_quality = genotype_quality;
    o.phase_score = 0.0;
    o.genotype_probability[0] = gp[0]; o.genotype_probability[1] = gp[1]; o.genotype_probability[2] = gp[2];
    o.allele_freqs[0] = allele1_freq; o.allele_freqs[1] = allele2_freq;
    o.depth = total;
    o.phase_set = 0;
    o.reference = ref_base;
    o.alleles[0] = allele1; o.alleles[1] = allele2;
    o.variant_type = (int8_t)variant_type;
    o.genotype = (int8_t)genotype;
    o.haplotype = 0;
    o.flags = fl;
    o.reserved = 0;
    return true;
}

/* shared-memory row byte of the tile kernel:
     bits 0-2  0-3 = A,C,G,T   4 = other read base   5 = deletion   6 = intron   7 = nothing
     bit  3    base quality >= min_baseq
     bit  4    read on the forward strand
     bits 5-6  transcript strand of the read: 1 forward, 2 reverse (util.rs:803-819)              */
#define ROW_NONE 7u
#define SUBTILE 128
#define NSUB (LCR_TILE / SUBTILE)
#define PWARPS (LCR_TILE / 32)

struct PreCand { /* a site that passed every count-based filter; its likelihood is computed by k_site_ll */
    uint32_t tile, col;
    uint32_t cnt[4], pass[4], fwd[4], ts[2], d, n;
};

/* 4 read bases + 4 qualities (little-endian words) -> 4 row bytes */
__device__ __forceinline__ uint32_t codes4(uint32_t w, uint32_t qv, uint32_t minq, uint32_t rconst4) {
    /* A,C,G,T -> 0..3 from bits 1-2 of the letter; anything else -> 4 */
    const uint32_t t = (w >> 1) & 0x03030303u;
    uint32_t c4 = t ^ ((t >> 1) & 0x01010101u);
    const uint32_t canon = __byte_perm(0x54474341u, 0, (c4 & 0x3u) | ((c4 >> 4) & 0x30u) | ((c4 >> 8) & 0x300u) | ((c4 >> 12) & 0x3000u));
    const uint32_t x = (w & 0xdfdfdfdfu) ^ canon;
    const uint32_t nz = ((x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) >> 7) & 0x01010101u;
    c4 = (c4 & ~(nz * 7u)) | (nz * 4u);
    /* quality >= min_baseq per byte */
    const uint32_t ge = (((qv & 0x7f7f7f7fu) | 0x80808080u) - minq * 0x01010101u) | (qv & 0x80808080u);
    const uint32_t pass4 = minq > 30u ? 0u : ((ge >> 4) & 0x08080808u);
    return c4 | pass4 | rconst4;
}

__global__ void __launch_bounds__(LCR_TILE, 2) k_pileup_tile(PileArgs a) {
    extern __shared__ __align__(16) uint8_t rows_raw[]; /* LCR_ROWS x LCR_TILE row bytes */
    uint8_t (*rows)[LCR_TILE] = reinterpret_cast<uint8_t (*)[LCR_TILE]>(rows_raw);
    __shared__ __align__(16) ulonglong2 lut[256];
    __shared__ uint8_t ref_s[LCR_TILE];
    __shared__ int32_t s_col[PWARPS][33];
    __shared__ int32_t s_rp[PWARPS][33];
    __shared__ uint8_t s_typ[PWARPS][32];
    __shared__ uint32_t fill[NSUB];
    __shared__ uint32_t next_item;
    __shared__ unsigned long long s_bases;
    __shared__ int s_err;

    const uint32_t tile = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t reg = a.tile_region[tile];
    const lcr_region R = a.regions[reg];
    const int64_t vec_size = (int64_t)R.end - (int64_t)R.start;
    const int32_t tile_start = (int32_t)((tile - a.tile_base[reg]) * LCR_TILE);
    const int32_t tile_end = (int64_t)tile_start + LCR_TILE < vec_size ? tile_start + LCR_TILE : (int32_t)vec_size;
    const uint32_t npos = (uint32_t)(tile_end - tile_start);
    if (tid == 0) s_err = a.rstate[reg].status; /* one read, so the whole CTA takes the same branch */
    __syncthreads();
    if (s_err != 0) return;
    const uint8_t *ref = a.ref_table[R.tid] + ((int64_t)R.start - 1) + tile_start;

    if (tid < 256) { /* per-code increments of the sixteen 8-bit column counters */
        const uint32_t b = tid & 7u, pass = (tid >> 3) & 1u, fwd = (tid >> 4) & 1u, ts = (tid >> 5) & 3u;
        unsigned long long x = 0, y = 0;
        if (b < 4u) {
            x |= 1ull << (8 * b);
            x |= (unsigned long long)pass << (32 + 8 * b);
            y |= (unsigned long long)fwd << (8 * b);
        }
        if (b <= 4u) {
            if (ts == 1u) y |= 1ull << 32;
            else if (ts == 2u) y |= 1ull << 40;
        }
        if (b == 5u) y |= 1ull << 48;
        if (b == 6u) y |= 1ull << 56;
        lut[tid] = make_ulonglong2(x, y);
    }
    if (tid == 0) s_bases = 0;
    const uint8_t ref_base = tid < npos ? ref[tid] : (uint8_t)'N';
    ref_s[tid] = ref_base;
    const uint32_t minq = (uint32_t)a.P.min_baseq;

    uint32_t cnt[4] = {0, 0, 0, 0}, pas[4] = {0, 0, 0, 0}, fwd[4] = {0, 0, 0, 0}, tsc[2] = {0, 0}, dcnt = 0, ncnt = 0;

    const uint32_t it0 = a.tile_off[tile], it1 = a.tile_off[tile + 1];
    unsigned long long my_bases = 0;
    for (uint32_t base_it = it0; base_it < it1; base_it += LCR_ROWS) {
        const uint32_t nitems = (it1 - base_it) < LCR_ROWS ? (it1 - base_it) : LCR_ROWS;
        {
            uint4 fillv;
            fillv.x = fillv.y = fillv.z = fillv.w = 0x07070707u;
            uint4 *r4 = reinterpret_cast<uint4 *>(rows_raw);
            const uint32_t n16 = nitems * (LCR_TILE / 16);
            for (uint32_t i = tid; i < n16; i += LCR_TILE) r4[i] = fillv;
            if (tid < NSUB) fill[tid] = 0;
            if (tid == 0) next_item = 0;
        }
        __syncthreads();
        /* phase 1: one warp per item.  A batch of up to 32 CIGAR ops is scanned (lane per op) into the first
           column / first read offset of every reference-consuming op; columns are then produced either
           4 per lane from word loads (long runs away from the read ends) or 1 per lane with the op of
           each column found from a ballot over the op starts. */
        for (;;) {
            uint32_t itx = 0;
            if (lane == 0) itx = atomicAdd(&next_item, 1u);
            itx = __shfl_sync(0xffffffffu, itx, 0);
            if (itx >= nitems) break;
            const LcrItem it = a.items[base_it + itx];
            const uint32_t read = R.read_begin + (it.slot - a.slot_off[reg]);
            const uint64_t s0 = a.seq_off[read];
            const int32_t seq_len = (int32_t)(a.seq_off[read + 1] - s0);
            const uint8_t *seq = a.seq + s0, *qual = a.qual + s0;
            const uint64_t c0 = a.cig_off[read];
            const uint32_t ncig = (uint32_t)(a.cig_off[read + 1] - c0);
            const uint32_t *cig = a.cigar + c0;
            const int32_t lead = (ncig && (cig[0] & 0xf) == 4) ? (int32_t)(cig[0] >> 4) : 0;
            const int32_t trail = (ncig && (cig[ncig - 1] & 0xf) == 4) ? (int32_t)(cig[ncig - 1] >> 4) : 0;
            const int32_t rb = seq_len - trail;
            const int32_t dend = (int32_t)(a.P.distance_to_read_end > 0x3fffffffu ? 0x3fffffffu : a.P.distance_to_read_end);
            uint32_t rconst;
            {
                const int strand = (a.flag[read] & 0x10) ? 1 : 0;
                const int8_t ts = a.ts[read];
                uint32_t tcode = 0;
                if (ts == '+') tcode = strand == 0 ? 1u : 2u;
                else if (ts == '-') tcode = strand == 0 ? 2u : 1u;
                rconst = (strand == 0 ? 16u : 0u) | (tcode << 5);
            }
            int32_t fpos = it.fpos;
            int32_t rpos = (int32_t)it.rpos;
            uint32_t ci = it.cig, off = it.opoff;
            uint32_t slots = 0xffffffffu; /* row of this item in each sub-tile, allocated on first touch */
            bool bad = false;
            while (ci < ncig && fpos < tile_end) {
                uint32_t opc = 15, len = 0;
                if (ci + lane < ncig) {
                    const uint32_t op = cig[ci + lane];
                    opc = op & 0xf;
                    len = op >> 4;
                    if (lane == 0) len -= off;
                }
                const bool consuming = opc == 0 || opc == 2 || opc == 3 || opc == 7 || opc == 8;
                const bool is_m = opc == 0 || opc == 7 || opc == 8;
                if (opc != 15 && !consuming && opc != 1 && opc != 4 && opc != 5) bad = true;
                const int32_t rl = consuming ? (int32_t)len : 0, ql = (is_m || opc == 1) ? (int32_t)len : 0;
                int32_t rs = rl, qs = ql;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int32_t r2 = __shfl_up_sync(0xffffffffu, rs, o), q2 = __shfl_up_sync(0xffffffffu, qs, o);
                    if ((int)lane >= o) { rs += r2; qs += q2; }
                }
                const int32_t tot_r = __shfl_sync(0xffffffffu, rs, 31), tot_q = __shfl_sync(0xffffffffu, qs, 31);
                /* compact the reference-consuming ops */
                const bool keep = consuming && rl > 0 && (fpos + rs - rl) < tile_end;
                const uint32_t keepmask = __ballot_sync(0xffffffffu, keep);
                const uint32_t ncomp = __popc(keepmask);
                if (keep) {
                    const uint32_t k = __popc(keepmask & ((1u << lane) - 1u));
                    s_col[warp][k] = fpos + rs - rl;
                    s_rp[warp][k] = rpos + qs - ql;
                    s_typ[warp][k] = (uint8_t)(is_m ? 0 : opc);
                }
                const int32_t batch_end = fpos + tot_r;
                if (lane == 0) s_col[warp][ncomp] = batch_end;
                const int32_t colB = batch_end < tile_end ? batch_end : tile_end;
                if (is_m) { /* aligned bases of this batch inside the tile (n_aligned_bases) + bounds */
                    const int32_t a0 = fpos + rs - rl, b0 = a0 + rl;
                    const int32_t lo = a0 > fpos ? a0 : fpos, hi = b0 < colB ? b0 : colB;
                    if (hi > lo) {
                        my_bases += (unsigned long long)(hi - lo);
                        if (rpos + qs - ql + (hi - a0) > seq_len) bad = true;
                    }
                }
                bad = __any_sync(0xffffffffu, bad);
                if (bad) break;
                __syncwarp();
                /* rows of this item in the sub-tiles this batch reaches */
                if (colB > fpos) {
                    const uint32_t subA = (uint32_t)(fpos - tile_start) / SUBTILE, subB = (uint32_t)(colB - 1 - tile_start) / SUBTILE;
                    for (uint32_t sidx = subA; sidx <= subB; ++sidx) {
                        if (((slots >> (8 * sidx)) & 0xffu) != 0xffu) continue;
                        uint32_t v = 0;
                        if (lane == 0) v = atomicAdd(&fill[sidx], 1u);
                        v = __shfl_sync(0xffffffffu, v, 0);
                        slots = (slots & ~(0xffu << (8 * sidx))) | (v << (8 * sidx));
                    }
                }
                /* every kept op is one run inside [fpos, colB); cut each run into 16-byte aligned chunks of the read
                   (or 16 columns of a D / N run) and let every lane take one chunk */
                int32_t run_a = 0, run_n = 0, run_rp = 0, run_al = 0;
                uint32_t run_typ = 0, nch = 0;
                if (lane < ncomp) {
                    const int32_t st = s_col[warp][lane], en = s_col[warp][lane + 1];
                    run_a = st > fpos ? st : fpos;
                    const int32_t bb = en < colB ? en : colB;
                    run_n = bb - run_a;
                    run_typ = s_typ[warp][lane];
                    run_rp = s_rp[warp][lane] + (run_a - st);
                    if (run_typ == 0) run_al = (int32_t)((uintptr_t)(seq + run_rp) & 15u);
                    nch = (uint32_t)(run_al + run_n + 15) >> 4;
                }
                uint32_t chs = nch; /* inclusive scan of the chunk counts */
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v2 = __shfl_up_sync(0xffffffffu, chs, o);
                    if ((int)lane >= o) chs += v2;
                }
                const uint32_t n_chunks = __shfl_sync(0xffffffffu, chs, 31);
                const int32_t mychoff = lane < ncomp ? (int32_t)(chs - nch) : 0x7fffffff;
                const uint32_t rconst4 = rconst * 0x01010101u;
                for (uint32_t base = 0; base < n_chunks; base += 32) {
                    const int32_t first = (int32_t)__popc(__ballot_sync(0xffffffffu, mychoff <= (int32_t)base)) - 1;
                    const int32_t rel = mychoff - (int32_t)base;
                    const uint32_t starts = __reduce_or_sync(0xffffffffu, (rel > 0 && rel < 32) ? (1u << rel) : 0u);
                    const int32_t k = first + (int32_t)__popc(starts & (0xffffffffu >> (31 - lane)));
                    /* run parameters of my chunk come from the lane that owns run k */
                    const int32_t k_a = __shfl_sync(0xffffffffu, run_a, k), k_n = __shfl_sync(0xffffffffu, run_n, k);
                    const int32_t k_rp = __shfl_sync(0xffffffffu, run_rp, k), k_al = __shfl_sync(0xffffffffu, run_al, k);
                    const uint32_t k_typ = __shfl_sync(0xffffffffu, run_typ, k);
                    const int32_t k_off = __shfl_sync(0xffffffffu, mychoff, k);
                    const uint32_t g = base + lane;
                    if (g < n_chunks) {
                        const int32_t c = (int32_t)g - k_off;              /* chunk index inside the run */
                        const int32_t lo = c == 0 ? k_al : 0;
                        int32_t hi = k_al + k_n - 16 * c;
                        if (hi > 16) hi = 16;
                        uint32_t w0, w1, w2, w3;
                        if (k_typ == 0) {
                            const uintptr_t sa = ((uintptr_t)(seq + k_rp) & ~(uintptr_t)15) + (uintptr_t)(16 * c);
                            const uint4 sv = __ldg(reinterpret_cast<const uint4 *>(sa));
                            const uint4 qv = __ldg(reinterpret_cast<const uint4 *>(sa + (uintptr_t)(qual - seq)));
                            w0 = codes4(sv.x, qv.x, minq, rconst4); w1 = codes4(sv.y, qv.y, minq, rconst4);
                            w2 = codes4(sv.z, qv.z, minq, rconst4); w3 = codes4(sv.w, qv.w, minq, rconst4);
                        } else w0 = w1 = w2 = w3 = (k_typ == 2 ? 5u : 6u) * 0x01010101u;
                        /* byte j of the chunk is column k_a - k_al + 16 c + j */
                        const uint32_t colr0 = (uint32_t)(k_a - k_al + 16 * c - tile_start);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const uint32_t wj = j < 4 ? w0 : (j < 8 ? w1 : (j < 12 ? w2 : w3));
                            if (j >= lo && j < hi) {
                                const uint32_t colr = colr0 + (uint32_t)j;
                                rows[(slots >> (8 * (colr / SUBTILE))) & 0xffu][colr] = (uint8_t)(wj >> (8 * (j & 3)));
                            }
                        }
                    }
                }
                /* read-end zones (util.rs:745-789): bases with |rp - lead| < D or |rp - rb| < D are trimmed (ONT) or tested
                   for poly-A / homopolymer runs; one lane per zone base of this batch */
                if (dend > 0) {
                    const int32_t zlo = lead + dend, zhi = rb - dend; /* rp < zlo or rp > zhi lies in a zone */
                    int32_t z0n = 0, z1n = 0, z1s = 0;                   /* zone bases of my run: [run_rp, run_rp+z0n) and [z1s, z1s+z1n) */
                    if (lane < ncomp && run_typ == 0) {
                        const int32_t e0 = run_rp + run_n < zlo ? run_rp + run_n : zlo;
                        z0n = e0 > run_rp ? e0 - run_rp : 0;
                        z1s = run_rp + z0n > zhi + 1 ? run_rp + z0n : zhi + 1;
                        z1n = run_rp + run_n > z1s ? run_rp + run_n - z1s : 0;
                    }
                    uint32_t zc = (uint32_t)(z0n + z1n), zs = zc;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t v2 = __shfl_up_sync(0xffffffffu, zs, o);
                        if ((int)lane >= o) zs += v2;
                    }
                    const uint32_t n_z = __shfl_sync(0xffffffffu, zs, 31);
                    if (n_z) {
                        __syncwarp();
                        s_rp[warp][lane] = (int32_t)(zs - zc); /* exclusive offsets; s_rp / s_col are free again after the chunk pass */
                        __syncwarp();
                        for (uint32_t base = 0; base < n_z; base += 32) {
                            const uint32_t g = base + lane;
                            /* owner run: last lane whose offset is <= g and that has zone bases */
                            uint32_t lo_k = 0;
#pragma unroll
                            for (int step = 16; step; step >>= 1)
                                if (lo_k + step < 32 && (uint32_t)s_rp[warp][lo_k + step] <= g) lo_k += step;
                            const int32_t k_rp = __shfl_sync(0xffffffffu, run_rp, lo_k), k_a = __shfl_sync(0xffffffffu, run_a, lo_k);
                            const int32_t k_z0n = __shfl_sync(0xffffffffu, z0n, lo_k), k_z1s = __shfl_sync(0xffffffffu, z1s, lo_k);
                            const uint32_t k_off = (uint32_t)__shfl_sync(0xffffffffu, s_rp[warp][lane], lo_k);
                            if (g < n_z) {
                                const int32_t idx = (int32_t)(g - k_off);
                                const int32_t r = idx < k_z0n ? k_rp + idx : k_z1s + (idx - k_z0n);
                                const uint32_t colr = (uint32_t)(k_a + (r - k_rp) - tile_start);
                                if (base_masked(a.P, seq, r, seq_len, lead, trail, ref_s[colr]))
                                    rows[(slots >> (8 * (colr / SUBTILE))) & 0xffu][colr] = (uint8_t)ROW_NONE;
                            }
                        }
                        __syncwarp();
                    }
                }
                fpos = batch_end;
                rpos += tot_q;
                ci += 32;
                off = 0;
                __syncwarp();
            }
            if (bad) s_err = LCR_ERR_BAD_CIGAR;
        }
        __syncthreads();
        /* phase 2: every thread sums the column of its position over the rows of its sub-tile */
        {
            unsigned long long acc0 = 0, acc1 = 0;
            const uint32_t nrow = fill[tid / SUBTILE];
            for (uint32_t row = 0; row < nrow; ++row) {
                const ulonglong2 v = lut[rows[row][tid]];
                acc0 += v.x;
                acc1 += v.y;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cnt[i] += (uint32_t)(acc0 >> (8 * i)) & 0xffu;
                pas[i] += (uint32_t)(acc0 >> (32 + 8 * i)) & 0xffu;
                fwd[i] += (uint32_t)(acc1 >> (8 * i)) & 0xffu;
            }
            tsc[0] += (uint32_t)(acc1 >> 32) & 0xffu;
            tsc[1] += (uint32_t)(acc1 >> 40) & 0xffu;
            dcnt += (uint32_t)(acc1 >> 48) & 0xffu;
            ncnt += (uint32_t)(acc1 >> 56) & 0xffu;
        }
        __syncthreads();
    }
    my_bases = __reduce_add_sync(0xffffffffu, (uint32_t)my_bases);
    if (lane == 0 && my_bases) atomicAdd(&s_bases, my_bases);
    __syncthreads();
    if (tid == 0) {
        if (s_bases) atomicAdd((unsigned long long *)&a.stats->n_aligned_bases, s_bases);
        if (s_err) atomicMin(&a.rstate[reg].status, s_err);
    }
    if (tid >= npos) return;
    ncnt += a.tile_full_n[tile];
    if (a.pl_acgt) {
        const uint64_t g = a.pos_off[reg] + (uint64_t)tile_start + tid;
#pragma unroll
        for (int i = 0; i < 4; ++i) { a.pl_acgt[g * 4 + i] = cnt[i]; a.pl_fwd[g * 4 + i] = fwd[i]; }
        a.pl_d[g] = dcnt; a.pl_n[g] = ncnt; a.pl_ts[g * 2] = tsc[0]; a.pl_ts[g * 2 + 1] = tsc[1];
    }
    SiteCounters sc;
#pragma unroll
    for (int i = 0; i < 4; ++i) { sc.cnt[i] = cnt[i]; sc.pass[i] = pas[i]; sc.fwd[i] = fwd[i]; }
    sc.ts[0] = tsc[0]; sc.ts[1] = tsc[1]; sc.d = dcnt; sc.n = ncnt; sc.ll0 = 0; sc.ll2 = 0; sc.q0flags = 0;
    lcr_candidate dummy;
    if (site_call<true>(a.P, *a.tables, sc, ref_base, dummy)) {
        const uint32_t k = atomicAdd(a.pre_count, 1u);
        if (k < a.pre_cap) {
            PreCand pc;
            pc.tile = tile; pc.col = tid;
#pragma unroll
            for (int i = 0; i < 4; ++i) { pc.cnt[i] = cnt[i]; pc.pass[i] = pas[i]; pc.fwd[i] = fwd[i]; }
            pc.ts[0] = tsc[0]; pc.ts[1] = tsc[1]; pc.d = dcnt; pc.n = ncnt;
            a.pre[k] = pc;
        }
    }
}

/* exact genotype likelihood of the sites that passed the count filters: one warp per site, lanes over the
   reads of the site's tile (candidate.rs:236-282 over the same unmasked bases the pileup counted) */
__global__ void __launch_bounds__(256) k_site_ll(PileArgs a, uint32_t n_pre) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_pre) return;
    const PreCand pc = a.pre[w];
    const uint32_t tile = pc.tile;
    const uint32_t reg = a.tile_region[tile];
    const lcr_region R = a.regions[reg];
    const int32_t tile_start = (int32_t)((tile - a.tile_base[reg]) * LCR_TILE);
    const int32_t col = tile_start + (int32_t)pc.col;
    const uint8_t ref_base = a.ref_table[R.tid][((int64_t)R.start - 1) + col];
    const int refc = (ref_base == 'A') ? 0 : (ref_base == 'C') ? 1 : (ref_base == 'G') ? 2 : (ref_base == 'T') ? 3 : 8;
    long long ll0 = 0, ll2 = 0;
    uint32_t q0flags = 0;
    for (uint32_t idx = a.tile_off[tile] + lane; idx < a.tile_off[tile + 1]; idx += 32) {
        const LcrItem it = a.items[idx];
        const uint32_t read = R.read_begin + (it.slot - a.slot_off[reg]);
        const uint64_t s0 = a.seq_off[read];
        const int64_t seq_len = (int64_t)(a.seq_off[read + 1] - s0);
        const uint8_t *seq = a.seq + s0, *qual = a.qual + s0;
        const uint64_t c0 = a.cig_off[read];
        const uint32_t ncig = (uint32_t)(a.cig_off[read + 1] - c0);
        const uint32_t *cig = a.cigar + c0;
        const int64_t lead = (ncig && (cig[0] & 0xf) == 4) ? (int64_t)(cig[0] >> 4) : 0;
        const int64_t trail = (ncig && (cig[ncig - 1] & 0xf) == 4) ? (int64_t)(cig[ncig - 1] >> 4) : 0;
        int32_t fpos = it.fpos;
        int64_t rpos = it.rpos;
        uint32_t off = it.opoff;
        for (uint32_t ci = it.cig; ci < ncig && fpos <= col; ++ci, off = 0) {
            const uint32_t op = cig[ci], opc = op & 0xf;
            const int32_t len = (int32_t)(op >> 4) - (int32_t)off;
            if (opc == 4 || opc == 5) continue;
            if (opc == 1) { rpos += len; continue; }
            const bool is_m = opc == 0 || opc == 7 || opc == 8;
            if (col < fpos + len) {
                if (is_m) {
                    const int64_t rp = rpos + (col - fpos);
                    if (rp < seq_len) {
                        const uint8_t b = seq[rp];
                        const uint32_t rq = qual[rp];
                        const uint32_t q = rq < LCR_MAX_BASE_QUALITY ? rq : LCR_MAX_BASE_QUALITY;
                        const int bc = base_code_dev(b);
                        if (bc >= 0 && !base_masked(a.P, seq, rp, seq_len, lead, trail, ref_base)) {
                            const bool is_ref = bc == refc;
                            const long long E = a.tables->gl_fx_err[q], K = a.tables->gl_fx_ok[q];
                            ll0 += is_ref ? E : K;
                            ll2 += is_ref ? K : E;
                            if (q == 0) q0flags |= is_ref ? 1u : 2u;
                        }
                    }
                }
                break;
            }
            fpos += len;
            if (is_m) rpos += len;
        }
    }
    for (int o = 16; o; o >>= 1) {
        ll0 += __shfl_xor_sync(0xffffffffu, ll0, o);
        ll2 += __shfl_xor_sync(0xffffffffu, ll2, o);
        q0flags |= __shfl_xor_sync(0xffffffffu, q0flags, o);
    }
    if (lane != 0) return;
    SiteCounters sc;
#pragma unroll
    for (int i = 0; i < 4; ++i) { sc.cnt[i] = pc.cnt[i]; sc.pass[i] = pc.pass[i]; sc.fwd[i] = pc.fwd[i]; }
    sc.ts[0] = pc.ts[0]; sc.ts[1] = pc.ts[1]; sc.d = pc.d; sc.n = pc.n; sc.ll0 = ll0; sc.ll2 = ll2; sc.q0flags = q0flags;
    lcr_candidate o;
    if (site_call<false>(a.P, *a.tables, sc, ref_base, o)) {
        const uint32_t k = atomicAdd(a.cand_count, 1u);
        if (k < a.cand_cap) {
            o.pos = (int64_t)R.start - 1 + col;
            o.region = reg;
            a.cand[k] = o;
            a.cand_key[k] = ((uint64_t)reg << 32) | (uint64_t)(uint32_t)col;
        }
    }
}

/* gather candidates into (region, position) order */
__global__ void k_cand_gather(const lcr_candidate *in, const uint32_t *perm, uint32_t n, lcr_candidate *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}

__global__ void k_iota(uint32_t *p, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

/* per region: candidate range in the sorted array + the dense-cluster filters (candidate.rs:465-526) */
__global__ void k_cand_finalize(lcr_params P, uint32_t n_regions, const uint64_t *keys, uint32_t n_cand, lcr_candidate *cand, LcrRegionState *rstate) {
    const uint32_t reg = blockIdx.x * blockDim.x + threadIdx.x;
    if (reg >= n_regions) return;
    uint32_t lo = 0, hi = n_cand;
    const uint64_t k0 = (uint64_t)reg << 32;
    while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (keys[m] < k0) lo = m + 1; else hi = m; }
    const uint32_t b = lo;
    hi = n_cand;
    const uint64_t k1 = (uint64_t)(reg + 1) << 32;
    while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (keys[m] < k1) lo = m + 1; else hi = m; }
    uint32_t e = lo;
    if (rstate[reg].status != 0) e = b; /* a failed region reports no candidates */
    rstate[reg].cand_begin = b;
    rstate[reg].n_cand = e - b;
    lcr_candidate *c = cand + b;
    const uint32_t n = e - b;
    /* concat_idxes = homo_snps + het_snps, sorted: the candidates carrying HOM_VAR or HET_VAR */
    for (int pass = 0; pass < 2; ++pass) {
        const int64_t win = pass == 0 ? (int64_t)P.dense_win_size : 5;
        const uint32_t min_cnt = pass == 0 ? P.min_dense_cnt : 3u;
        for (uint32_t i = 0; i < n; ++i) {
            if (!(c[i].flags & (LCR_CF_HOM_VAR | LCR_CF_HET_VAR))) continue;
            const int64_t start_pos = c[i].pos;
            uint32_t cnt_between = 0; /* j - i in concat_idxes terms */
            uint32_t last_member = i;
            bool broke = false;
            for (uint32_t j = i; j < n; ++j) {
                if (!(c[j].flags & (LCR_CF_HOM_VAR | LCR_CF_HET_VAR))) continue;
                const int64_t diff = c[j].pos - start_pos;
                const bool over = pass == 0 ? diff > win : diff >= win;
                if (over) {
                    if (cnt_between >= min_cnt)
                        for (uint32_t tk = i; tk < j; ++tk)
                            if (c[tk].flags & (LCR_CF_HOM_VAR | LCR_CF_HET_VAR)) c[tk].flags = (uint16_t)((c[tk].flags | LCR_CF_DENSE) & ~LCR_CF_FOR_PHASING);
                    broke = true;
                    break;
                }
                last_member = j;
                cnt_between++;
            }
            /* reached the last element inside the window: (j - i + 1) >= min_cnt marks i..j exclusive */
            if (!broke && cnt_between >= min_cnt)
                for (uint32_t tk = i; tk < last_member; ++tk)
                    if (c[tk].flags & (LCR_CF_HOM_VAR | LCR_CF_HET_VAR)) c[tk].flags = (uint16_t)((c[tk].flags | LCR_CF_DENSE) & ~LCR_CF_FOR_PHASING);
        }
    }
}

__global__ void k_count_pass(const uint8_t *slot_flags, uint32_t n_slots, lcr_stats *stats) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = (i < n_slots && slot_flags[i]) ? 1u : 0u;
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd((unsigned long long *)&stats->n_reads_pass, (unsigned long long)v);
}

} // namespace

/* scratch kept between the stages of one run (owned by api.cu) */
struct LcrRunScratch;

int lcr_stage_pileup_impl(lcr_ctx *ctx, lcr_device_batch *db, uint8_t *slot_flags);

#define TRY(expr) LCR_CUDA_TRY(ctx, expr)

int lcr_stage_pileup_impl(lcr_ctx *ctx, lcr_device_batch *db, uint8_t *slot_flags) {
    cudaStream_t st = ctx->stream;
    const uint32_t n_tiles = db->n_tiles;
    uint32_t *tile_count = nullptr, *tile_off = nullptr, *tile_full_n = nullptr;
    LcrItem *items = nullptr;
    TRY(cudaMallocAsync(&tile_count, sizeof(uint32_t) * (n_tiles + 1), st));
    TRY(cudaMallocAsync(&tile_off, sizeof(uint32_t) * (n_tiles + 1), st));
    TRY(cudaMallocAsync(&tile_full_n, sizeof(uint32_t) * (n_tiles + 1), st));
    TRY(cudaMemsetAsync(tile_count, 0, sizeof(uint32_t) * (n_tiles + 1), st));
    TRY(cudaMemsetAsync(tile_full_n, 0, sizeof(uint32_t) * (n_tiles + 1), st));

    PrepArgs pa{};
    pa.P = ctx->P;
    pa.n_slots = db->n_slots;
    pa.regions = db->regions;
    pa.slot_off = db->slot_off; pa.slot_region = db->slot_region; pa.tile_base = db->tile_base;
    pa.pos = db->pos; pa.flag = db->flag; pa.mapq = db->mapq; pa.de = db->de;
    pa.seq_off = db->seq_off; pa.cig_off = db->cig_off; pa.cigar = db->cigar;
    pa.rstate = db->rstate;
    pa.slot_flags = slot_flags;
    pa.tile_count = tile_count; pa.tile_off = tile_off; pa.tile_full_n = tile_full_n; pa.items = nullptr;
    const uint32_t pb = 128, pg = (db->n_slots + pb - 1) / pb;
    if (pg) {
        k_slot_prep<false><<<pg, pb, 0, st>>>(pa);
        k_count_pass<<<pg, pb, 0, st>>>(slot_flags, db->n_slots, db->d_stats);
        db->timing.kernel_launches += 2;
    }
    /* exclusive scan of the per-tile item counts */
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, tile_count, tile_off, n_tiles + 1, st);
    void *tmp = nullptr;
    TRY(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, st));
    TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, tile_count, tile_off, n_tiles + 1, st));
    uint32_t n_items = 0;
    TRY(cudaMemcpyAsync(&n_items, tile_off + n_tiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
    TRY(cudaMallocAsync(&items, sizeof(LcrItem) * (size_t)(n_items ? n_items : 1), st));
    TRY(cudaMemsetAsync(tile_count, 0, sizeof(uint32_t) * (n_tiles + 1), st));
    pa.items = items;
    if (pg) {
        k_slot_prep<true><<<pg, pb, 0, st>>>(pa);
        db->timing.kernel_launches += 1;
    }

    /* tile pileup (counts + count-based site filters), then the exact likelihood of the surviving sites */
    uint32_t pre_cap = (uint32_t)std::min<uint64_t>(db->n_pos, db->n_pos / 8 + 4096);
    if (!pre_cap) pre_cap = 1;
    PreCand *pre = nullptr;
    lcr_candidate *cand_raw = nullptr;
    uint64_t *cand_key = nullptr;
    uint32_t *counters = nullptr; /* [0] pre-candidates, [1] candidates */
    TRY(cudaMallocAsync(&counters, 2 * sizeof(uint32_t), st));
    cudaEvent_t ev0, ev1;
    TRY(cudaEventCreate(&ev0));
    TRY(cudaEventCreate(&ev1));
    const size_t tile_smem = (size_t)LCR_ROWS * LCR_TILE;
    TRY(cudaFuncSetAttribute(k_pileup_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
    PileArgs ka{};
    ka.P = ctx->P;
    ka.regions = db->regions;
    ka.slot_off = db->slot_off; ka.slot_region = db->slot_region; ka.tile_base = db->tile_base; ka.tile_region = db->tile_region;
    ka.pos_off = db->pos_off;
    ka.flag = db->flag; ka.ts = db->ts; ka.seq_off = db->seq_off; ka.cig_off = db->cig_off;
    ka.seq = db->seq; ka.qual = db->qual; ka.cigar = db->cigar;
    ka.ref_table = ctx->d_ref_table;
    ka.tile_off = tile_off; ka.tile_full_n = tile_full_n; ka.items = items;
    ka.tables = ctx->d_tables;
    ka.rstate = db->rstate;
    ka.stats = db->d_stats;
    ka.pl_acgt = db->pl_acgt; ka.pl_fwd = db->pl_fwd; ka.pl_d = db->pl_d; ka.pl_n = db->pl_n; ka.pl_ts = db->pl_ts;
    ka.pre_count = counters; ka.cand_count = counters + 1;
    uint32_t n_pre = 0, n_cand = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        TRY(cudaMallocAsync(&pre, sizeof(PreCand) * (size_t)pre_cap, st));
        TRY(cudaMemsetAsync(counters, 0, 2 * sizeof(uint32_t), st));
        ka.pre = pre; ka.pre_cap = pre_cap;
        if (attempt == 1) TRY(cudaMemsetAsync(&db->d_stats->n_aligned_bases, 0, sizeof(uint64_t), st));
        TRY(cudaEventRecord(ev0, st));
        if (n_tiles) {
            k_pileup_tile<<<n_tiles, LCR_TILE, tile_smem, st>>>(ka);
            db->timing.kernel_launches += 1;
        }
        TRY(cudaEventRecord(ev1, st));
        TRY(cudaMemcpyAsync(&n_pre, counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        TRY(cudaStreamSynchronize(st));
        TRY(cudaGetLastError());
        if (n_pre <= pre_cap) break;
        TRY(cudaFreeAsync(pre, st));
        pre_cap = n_pre;
    }
    float ms = 0;
    TRY(cudaEventElapsedTime(&ms, ev0, ev1));
    db->timing.ms_pileup_kernel = ms;
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    const uint32_t cand_cap = n_pre ? n_pre : 1;
    TRY(cudaMallocAsync(&cand_raw, sizeof(lcr_candidate) * (size_t)cand_cap, st));
    TRY(cudaMallocAsync(&cand_key, sizeof(uint64_t) * (size_t)cand_cap, st));
    ka.cand = cand_raw; ka.cand_key = cand_key; ka.cand_cap = cand_cap;
    if (n_pre) {
        k_site_ll<<<(uint32_t)(((uint64_t)n_pre * 32 + 255) / 256), 256, 0, st>>>(ka, n_pre);
        db->timing.kernel_launches += 1;
    }
    /* algorithmic bytes of the tile kernel: base + qual per aligned base, CIGAR, items, reference, surviving sites */
    {
        lcr_stats hs;
        TRY(cudaMemcpyAsync(&n_cand, counters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        TRY(cudaMemcpyAsync(&hs, db->d_stats, sizeof hs, cudaMemcpyDeviceToHost, st));
        TRY(cudaStreamSynchronize(st));
        TRY(cudaGetLastError());
        db->timing.pileup_alg_bytes = 2ull * hs.n_aligned_bases + 4ull * db->n_cigar + sizeof(LcrItem) * (uint64_t)n_items + db->n_pos + sizeof(PreCand) * (uint64_t)n_pre;
    }
    TRY(cudaFreeAsync(pre, st));
    uint32_t *cand_count = counters;

    /* sort candidates by (region, position) */
    db->n_cand = n_cand;
    uint64_t *keys_sorted = nullptr;
    uint32_t *perm_in = nullptr, *perm_out = nullptr;
    TRY(cudaMallocAsync(&db->cand, sizeof(lcr_candidate) * (size_t)(n_cand ? n_cand : 1), st));
    TRY(cudaMallocAsync(&keys_sorted, sizeof(uint64_t) * (size_t)(n_cand ? n_cand : 1), st));
    if (n_cand) {
        TRY(cudaMallocAsync(&perm_in, sizeof(uint32_t) * n_cand, st));
        TRY(cudaMallocAsync(&perm_out, sizeof(uint32_t) * n_cand, st));
        k_iota<<<(n_cand + 255) / 256, 256, 0, st>>>(perm_in, n_cand);
        size_t sb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sb, cand_key, keys_sorted, perm_in, perm_out, (int)n_cand, 0, 64, st);
        void *stmp = nullptr;
        TRY(cudaMallocAsync(&stmp, sb ? sb : 16, st));
        TRY(cub::DeviceRadixSort::SortPairs(stmp, sb, cand_key, keys_sorted, perm_in, perm_out, (int)n_cand, 0, 64, st));
        k_cand_gather<<<(n_cand + 127) / 128, 128, 0, st>>>(cand_raw, perm_out, n_cand, db->cand);
        db->timing.kernel_launches += 2; /* own kernels only; the cub sort is library code */
        TRY(cudaFreeAsync(stmp, st));
        TRY(cudaFreeAsync(perm_in, st));
        TRY(cudaFreeAsync(perm_out, st));
    }
    if (db->n_regions) {
        k_cand_finalize<<<(db->n_regions + 63) / 64, 64, 0, st>>>(ctx->P, db->n_regions, keys_sorted, n_cand, db->cand, db->rstate);
        db->timing.kernel_launches += 1;
    }
    TRY(cudaFreeAsync(keys_sorted, st));
    TRY(cudaFreeAsync(cand_raw, st));
    TRY(cudaFreeAsync(cand_key, st));
    TRY(cudaFreeAsync(cand_count, st));
    TRY(cudaFreeAsync(items, st));
    TRY(cudaFreeAsync(tmp, st));
    TRY(cudaFreeAsync(tile_count, st));
    TRY(cudaFreeAsync(tile_off, st));
    TRY(cudaFreeAsync(tile_full_n, st));
    TRY(cudaGetLastError());
    return LCR_OK;
}
