/*
 * pileup.cu — read filter, tile work items, fused pileup + site genotyping, candidate lists.
 *
 * Replaces, on the device:
 *   src/util.rs:636-668      read filter and fetch window             (k_slot_prep / k_slot_prep_w)
 *   src/util.rs:650-948      Profile::fill_data_into_freq_vec         (k_slot_prep: CIGAR walk + masks; k_pileup_tile: counts)
 *   src/util.rs:162-176      BaseFreq::get_two_major_alleles          (site_call)
 *   src/candidate.rs:75-463  filter cascade, genotype likelihood      (site_call, k_site_ll)
 *   src/candidate.rs:465-526 dense-cluster filters                    (k_cand_ranges, k_cand_dense)
 *
 * Layout: reads are decomposed on the device into (read, tile) items and, per item, segments (runs of unmasked aligned
 * bases / deleted / intron positions on consecutive columns); one CTA owns one tile of LCR_TILE reference positions,
 * stages the reads of the tile as one-hot byte planes in shared memory and sums the columns with carry-save adders:
 * no atomics on the counters, no per-position record in HBM.  Only candidate sites (and, on request, the debug planes)
 * are written out.
 */
#include <cub/cub.cuh>

#include "lcr_device.h"

namespace {

__device__ __forceinline__ bool is_ref_consuming(uint32_t opc) { return opc == 0 || opc == 2 || opc == 3 || opc == 7 || opc == 8; }

/* ------------------------------------------------------------------------- */

__device__ __forceinline__ int base_code_dev(uint8_t b) { /* A,C,G,T in either case -> 0..3, anything else -> -1 */
    const uint32_t u = b & 0xdfu;
    return u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : u == 'T' ? 3 : -1;
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
    /* two rejections of the cascade below, taken first because they settle most positions: the last test of the count
       filters (candidate.rs:132, upper-case A/C/G/T only), and a column holding nothing but the reference base, whose
       alternative allele has count 0 and fails `d >= alt count` (candidate.rs:165) */
    if (!(ref_base == 'A' || ref_base == 'C' || ref_base == 'G' || ref_base == 'T')) return false;
    {
        const uint32_t rc = ref_base == 'A' ? s.cnt[0] : ref_base == 'C' ? s.cnt[1] : ref_base == 'G' ? s.cnt[2] : s.cnt[3];
        if (rc == total) return false;
    }
    /* get_two_major_alleles: stable descending sort of (A,C,G,T); keys (count << 2 | 3 - index) through a 5-exchange network */
    uint32_t k0 = (s.cnt[0] << 2) | 3u, k1 = (s.cnt[1] << 2) | 2u, k2 = (s.cnt[2] << 2) | 1u, k3 = s.cnt[3] << 2;
#define LCR_CX(x, y) do { const uint32_t hi__ = max(x, y), lo__ = min(x, y); x = hi__; y = lo__; } while (0)
    LCR_CX(k0, k1); LCR_CX(k2, k3); LCR_CX(k0, k2); LCR_CX(k1, k3); LCR_CX(k1, k2);
#undef LCR_CX
    const int ord[4] = {3 - (int)(k0 & 3u), 3 - (int)(k1 & 3u), 3 - (int)(k2 & 3u), 3 - (int)(k3 & 3u)};
    const uint32_t oc[4] = {k0 >> 2, k1 >> 2, k2 >> 2, k3 >> 2}; /* counts in sorted order */
    auto letter = [](int c) -> uint8_t { return (uint8_t)(0x54474341u >> (8 * c)); };                  /* "ACGT"[c] */
    auto pick = [](const uint32_t (&v)[4], int i) -> uint32_t { return i == 0 ? v[0] : i == 1 ? v[1] : i == 2 ? v[2] : v[3]; }; /* registers only */
    int i1 = ord[0], i2 = ord[1];
    uint32_t allele1_cnt = oc[0], allele2_cnt = oc[1];
    if (letter(ord[0]) != ref_base && letter(ord[1]) != ref_base) {
        if (oc[2] == oc[1] && letter(ord[2]) == ref_base) { i2 = ord[2]; allele2_cnt = oc[2]; }
        else if (oc[3] == oc[1] && letter(ord[3]) == ref_base) { i2 = ord[3]; allele2_cnt = oc[3]; }
    }
    const uint8_t allele1 = letter(i1), allele2 = letter(i2);
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
    if (allele1 != ref_base) { if (allele1_cnt > 0 && pick(s.pass, i1) < 2) return false; }
    else if (allele2 != ref_base) { if (allele2_cnt > 0 && pick(s.pass, i2) < 2) return false; }
    if (P.use_strand_bias) {
        const int32_t rf = (int32_t)pick(s.fwd, ref_code), rr = (int32_t)(pick(s.cnt, ref_code) - pick(s.fwd, ref_code));
        const int32_t af = (int32_t)pick(s.fwd, alt_i[0]), ar = (int32_t)(pick(s.cnt, alt_i[0]) - pick(s.fwd, alt_i[0]));
        float sor;
        if (alt_num == 1) sor = lcr_strand_odds_ratio(rf, rr, af, ar);
        else {
            const int32_t bf = (int32_t)pick(s.fwd, alt_i[1]), br = (int32_t)(pick(s.cnt, alt_i[1]) - pick(s.fwd, alt_i[1]));
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
    const uint8_t alt0 = letter(alt_i[0]);
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

struct LcrSeg {             /* 16 B: a run of unmasked aligned bases, deleted or intron positions of one read inside one tile */
    uint64_t spos;          /* M: offset of the first base in the seq / qual pools */
    uint32_t row_typ;       /* bits 0-1 type, bit 2 forward strand, bits 3-4 transcript strand code, bits 8-31 row (item index in its tile) */
    uint16_t col;           /* first column inside the tile */
    uint16_t len;           /* 1 .. LCR_TILE */
};
#define SEG_M 0u
#define SEG_D 1u
#define SEG_N 2u
#define LCR_SLOT_RUNS 4     /* homopolymer runs remembered per read for the poly-A mask */

struct PrepArgs {
    lcr_params P;
    uint32_t n_slots;
    const lcr_region *regions;
    const uint32_t *slot_off, *slot_region, *tile_base;
    const int32_t *pos;
    const uint16_t *flag;
    const uint8_t *mapq;
    const int8_t *ts;
    const float *de;
    const uint64_t *seq_off, *cig_off;
    const uint8_t *seq;
    const uint32_t *cigar;
    const uint8_t *const *ref_table;
    LcrRegionState *rstate;
    lcr_stats *stats;
    uint8_t *slot_flags;   /* bit 0: passes the read filter and overlaps the window; bit 1: has homopolymer runs near a read end;
                              bit 2: more than LCR_SLOT_RUNS of them (every zone base takes the exact test) */
    uint64_t *slot_runs;   /* [n_slots][LCR_SLOT_RUNS]: start << 32 | length << 8 | letter, in read coordinates */
    uint32_t *tile_count;  /* COUNT: items per tile; FILL: cursor */
    const uint32_t *tile_off;
    uint32_t *tile_full_n; /* whole-tile intron covers */
    uint32_t *deep_flag;   /* COUNT: set when some tile holds more than 255 items */
    uint32_t *slot_segs;   /* COUNT: upper bound of the read's segments (exact but for poly-A cuts) */
    const uint32_t *slot_seg_off; /* its exclusive scan: the read's segments are written there, in walk order, no atomics */
    uint2 *item_segs;      /* FILL: per item (tile order; an item is the part of one read inside one tile), first segment and segment count */
    LcrSeg *segs;
};

/* Read filter (util.rs:652-668), fetch window, and the decomposition of every passing read into per-tile items
   (a row index in the tile plus the read's segment range there) and segments (what k_pileup_tile and k_site_ll read).  The read-end trim (ONT) and the
   poly-A / homopolymer mask (util.rs:737-789) are applied here by cutting M runs at masked bases.  Two passes
   around an exclusive scan: COUNT sizes the per-tile lists, FILL writes them. */
template <bool FILL>
__global__ void __launch_bounds__(128, 6) k_slot_prep(PrepArgs a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n_bases = 0;
    bool live = slot < a.n_slots;
    uint32_t reg = 0;
    if (live) {
        reg = a.slot_region[slot];
        if (a.rstate[reg].status != 0) live = false;
    }
    if (live) {
        const lcr_region R = a.regions[reg];
        const uint32_t read = R.read_begin + (slot - a.slot_off[reg]);
        const uint64_t c0 = a.cig_off[read], c1 = a.cig_off[read + 1];
        const uint64_t s0 = a.seq_off[read];
        const int64_t seq_len = (int64_t)(a.seq_off[read + 1] - s0);
        const uint8_t *seq = a.seq + s0;
        const int64_t lead = (c1 > c0 && (a.cigar[c0] & 0xf) == 4) ? (int64_t)(a.cigar[c0] >> 4) : 0; /* leading_softclips */
        const int64_t trail = (c1 > c0 && (a.cigar[c1 - 1] & 0xf) == 4) ? (int64_t)(a.cigar[c1 - 1] >> 4) : 0;
        const int64_t rb = seq_len - trail;
        const int64_t dend = (int64_t)(a.P.distance_to_read_end > 0x3fffffffu ? 0x3fffffffu : a.P.distance_to_read_end);
        const bool ont = a.P.platform == 1;
        uint8_t sflags;
        uint64_t runs[LCR_SLOT_RUNS];
#pragma unroll
        for (int i = 0; i < LCR_SLOT_RUNS; ++i) runs[i] = 0;
        if (!FILL) {
            /* util.rs:652-668 */
            const uint16_t fl = a.flag[read];
            bool pass = !((int32_t)a.mapq[read] < a.P.min_mapq || (uint64_t)seq_len < (uint64_t)a.P.min_read_length || (fl & 0x4) || (fl & 0x100) || (fl & 0x800));
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
            sflags = (pass && in_window) ? 1 : 0;
            if (sflags && !ont && dend > 0) {
                /* util.rs:754-789 can only mask a base next to (or inside) a homopolymer run of polya_tail_length letters that
                   lies within polya_tail_length of a read-end zone: remember those runs, the walk below tests only their bases */
                const int64_t polya = (int64_t)(a.P.polya_tail_length > 0x3fffffffu ? 0x3fffffffu : a.P.polya_tail_length);
                if (polya < 2 || rb < lead) sflags |= 6; /* degenerate window or overlapping clips: exact test on every zone base */
                else {
                    const int64_t centre[2] = {lead, rb};
                    uint32_t nrun = 0;
                    int64_t scanned_to = -1; /* runs ending at or before this index are already recorded */
                    for (int z = 0; z < 2; ++z) {
                        int64_t lo = centre[z] - dend + 1 - polya, hi = centre[z] + dend + polya; /* [lo, hi) */
                        if (lo < 0) lo = 0;
                        if (hi > seq_len) hi = seq_len;
                        int64_t run = 0;
                        uint8_t prev = 0;
                        for (int64_t i = lo; i <= hi; ++i) {
                            const uint8_t b = i < hi ? __ldg(seq + i) : (uint8_t)0;
                            const bool letter = b == 'A' || b == 'C' || b == 'G' || b == 'T';
                            if (letter && b == prev) { run++; continue; }
                            if (run >= polya && i > scanned_to) { /* the run [i - run, i) just ended */
                                if (nrun < LCR_SLOT_RUNS) runs[nrun] = ((uint64_t)(i - run) << 32) | ((uint64_t)(run > 0xffffff ? 0xffffff : run) << 8) | (uint64_t)prev;
                                nrun++;
                            }
                            run = letter ? 1 : 0;
                            prev = b;
                        }
                        if (hi > scanned_to) scanned_to = hi;
                    }
                    if (nrun) sflags |= 2;
                    if (nrun > LCR_SLOT_RUNS) sflags |= 4;
                }
            }
            a.slot_flags[slot] = sflags;
            if (sflags & 2) {
#pragma unroll
                for (int i = 0; i < LCR_SLOT_RUNS; ++i) a.slot_runs[(size_t)slot * LCR_SLOT_RUNS + i] = runs[i];
            }
        } else {
            sflags = a.slot_flags[slot];
            if ((sflags & 6) == 2) {
#pragma unroll
                for (int i = 0; i < LCR_SLOT_RUNS; ++i) runs[i] = a.slot_runs[(size_t)slot * LCR_SLOT_RUNS + i];
            }
        }
        if (sflags & 1) {
            const int64_t vec_size = (int64_t)R.end - (int64_t)R.start;
            const int64_t fv_start = (int64_t)R.start - 1;
            const uint32_t tb = a.tile_base[reg];
            const uint8_t *ref = a.ref_table[R.tid] + fv_start;
            uint32_t rowtyp_c;
            {
                const int strand = (a.flag[read] & 0x10) ? 1 : 0;
                const int8_t ts = a.ts[read];
                uint32_t tcode = 0;
                if (ts == '+') tcode = strand == 0 ? 1u : 2u;
                else if (ts == '-') tcode = strand == 0 ? 2u : 1u;
                rowtyp_c = (strand == 0 ? 4u : 0u) | (tcode << 3);
            }
            /* read-end zones in read coordinates, ordered by their first base */
            int64_t zlo[2] = {lead - dend + 1, rb - dend + 1}, zhi[2] = {lead + dend - 1, rb + dend - 1};
            if (zlo[1] < zlo[0]) { int64_t t = zlo[0]; zlo[0] = zlo[1]; zlo[1] = t; t = zhi[0]; zhi[0] = zhi[1]; zhi[1] = t; }
            const int mask_mode = dend <= 0 ? 0 : ont ? 1 : (sflags & 4) ? 2 : (sflags & 2) ? 3 : 0;

            int64_t fpos = (int64_t)a.pos[read] - fv_start;
            int64_t rpos = lead;
            int64_t last_tile = -1;
            uint32_t item_k = 0xffffffffu;                 /* FILL: index of the current item */
            uint32_t seg_w = FILL ? a.slot_seg_off[slot] : 0; /* FILL: next segment slot of this read; COUNT: segments so far */
            uint32_t seg_item0 = seg_w;                    /* first segment of the current item */
            bool bad = false;
            auto leave_tile = [&]() {
                if (FILL && item_k != 0xffffffffu) a.item_segs[item_k] = make_uint2(seg_item0, seg_w - seg_item0);
                seg_item0 = seg_w;
            };
            for (uint64_t c = c0; c < c1 && !bad; ++c) {
                const uint32_t op = a.cigar[c], opc = op & 0xf, len = op >> 4;
                if (opc == 4 || opc == 5) continue;
                if (opc == 1) {
                    if (fpos >= vec_size && fpos >= 1) break;
                    rpos += len;
                    continue;
                }
                if (!is_ref_consuming(opc)) { bad = true; break; } /* util.rs:943-945 panics */
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
                            leave_tile();
                            last_tile = t;
                            /* reads are position-sorted, so the lanes of a warp enter the same few tiles: one atomic per tile and warp */
                            const uint32_t key = tb + (uint32_t)t;
                            const unsigned peers = __match_any_sync(__activemask(), key);
                            const int leader = __ffs(peers) - 1;
                            const uint32_t npeer = (uint32_t)__popc(peers), prank = (uint32_t)__popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
                            uint32_t first = 0;
                            if ((int)(threadIdx.x & 31) == leader) first = atomicAdd(&a.tile_count[key], npeer);
                            first = __shfl_sync(peers, first, leader);
                            if (!FILL) { if (first <= 255u && first + npeer > 255u) *a.deep_flag = 1u; }
                            else {
                                item_k = a.tile_off[key] + first + prank;
                            }
                        }
                        const uint32_t colr = (uint32_t)(ts - t * LCR_TILE);
                        auto put = [&](uint32_t typ, uint64_t spos, uint32_t col, uint32_t n) {
                            if (FILL) {
                                LcrSeg s;
                                s.spos = spos; s.row_typ = rowtyp_c | typ; s.col = (uint16_t)col; s.len = (uint16_t)n;
                                a.segs[seg_w] = s;
                            }
                            seg_w++;
                        };
                        if (!is_m) {
                            put(opc == 2 ? SEG_D : SEG_N, 0, colr, (uint32_t)(te - ts));
                            continue;
                        }
                        const int64_t pa = rpos + (ts - lo), pb = rpos + (te - lo); /* read coordinates of this piece */
                        if (pb > seq_len) { bad = true; break; }
                        n_bases += (uint32_t)(te - ts);
                        int64_t start = pa;
                        auto emit = [&](int64_t x, int64_t y) { /* unmasked stretch [x, y) */
                            if (y > x) put(SEG_M, s0 + (uint64_t)x, colr + (uint32_t)(x - pa), (uint32_t)(y - x));
                        };
                        if (mask_mode == 1) { /* util.rs:745-751: every base of a zone */
                            for (int k = 0; k < 2; ++k) {
                                const int64_t zl = zlo[k] > start ? zlo[k] : start, zh = zhi[k] < pb - 1 ? zhi[k] : pb - 1;
                                if (zl <= zh) { emit(start, zl); start = zh + 1; }
                            }
                        } else if (mask_mode == 2) { /* exact test on every zone base (COUNT: every one may cut) */
                            for (int k = 0; k < 2; ++k) {
                                const int64_t zl = zlo[k] > start ? zlo[k] : start, zh = zhi[k] < pb - 1 ? zhi[k] : pb - 1;
                                if (!FILL) { if (zh >= zl) seg_w += (uint32_t)(zh - zl + 1); continue; }
                                for (int64_t rp = zl; rp <= zh; ++rp)
                                    if (base_masked(a.P, seq, rp, seq_len, lead, trail, ref[ts + (rp - pa)])) { emit(start, rp); start = rp + 1; }
                            }
                        } else if (mask_mode == 3) {
                            /* only bases inside or next to a remembered run can be masked, and the test of util.rs:754-789 needs no
                               sequence loads there: base rp is masked iff, for a run [rs, re) of a letter other than the reference base
                               that holds rp - 1 or rp + 1, at least polya of its bases lie in [rp - polya, rp + polya] */
                            const int64_t polya = (int64_t)a.P.polya_tail_length;
                            int64_t from = pa;
#pragma unroll
                            for (int j = 0; j < LCR_SLOT_RUNS; ++j) {
                                const int64_t rs = (int64_t)(runs[j] >> 32), rn = (int64_t)((runs[j] >> 8) & 0xffffffu);
                                if (rn == 0) continue;
                                const int64_t zl = rs - 1 > from ? rs - 1 : from, zh = rs + rn < pb - 1 ? rs + rn : pb - 1;
                                if (!FILL) { if (zh >= zl) seg_w += (uint32_t)(zh - zl + 1); }
                                else {
                                    for (int64_t rp = zl; rp <= zh; ++rp) {
                                        const int64_t d0 = rp - lead, d1 = rp - rb;
                                        if (!((d0 < 0 ? -d0 : d0) < dend || (d1 < 0 ? -d1 : d1) < dend)) continue;
                                        const uint8_t rbase = ref[ts + (rp - pa)];
                                        const int64_t wlo = rp - polya > 0 ? rp - polya : 0, whi = rp + polya + 1 < seq_len ? rp + polya + 1 : seq_len;
                                        bool m = false;
#pragma unroll
                                        for (int q = 0; q < LCR_SLOT_RUNS; ++q) { /* runs overlap or touch only when the two scans met: test all */
                                            const int64_t qs = (int64_t)(runs[q] >> 32), qn = (int64_t)((runs[q] >> 8) & 0xffffffu), qe = qs + qn;
                                            if (qn == 0 || (uint8_t)(runs[q] & 0xffu) == rbase) continue;
                                            const bool anchored = (rp - 1 >= qs && rp - 1 < qe) || (rp + 1 >= qs && rp + 1 < qe);
                                            const int64_t ov = (qe < whi ? qe : whi) - (qs > wlo ? qs : wlo);
                                            if (anchored && ov >= polya) m = true;
                                        }
                                        if (m) { emit(start, rp); start = rp + 1; }
                                    }
                                }
                                if (zh + 1 > from) from = zh + 1;
                            }
                        }
                        emit(start, pb);
                    }
                }
                fpos = hi;
                if (is_m) rpos += len;
            }
            leave_tile();
            if (bad) { atomicMin(&a.rstate[reg].status, (int32_t)LCR_ERR_BAD_CIGAR); n_bases = 0; }
            if (!FILL) a.slot_segs[slot] = seg_w;
        }
    }
    if (FILL) {
        n_bases = __reduce_add_sync(0xffffffffu, n_bases);
        if ((threadIdx.x & 31) == 0 && n_bases) atomicAdd((unsigned long long *)&a.stats->n_aligned_bases, (unsigned long long)n_bases);
    }
}

/* ---- the same two passes with one warp per read and one lane per CIGAR op ----
   The thread-per-read walk is a serial chain of ~60 instructions per op, so a batch of long-CIGAR reads (ONT: ~100 ops,
   up to several hundred) runs as long as its longest read.  Here the start position of every op comes from warp prefix sums
   over the op lengths; "first op of the read in a tile" (item creation), the segment slots and the open item of the read
   are carried from lane to lane with warp scans.  Results are identical to k_slot_prep (the item and segment orders inside
   a read are the same; rows inside a tile come from the same atomic counter). */
template <bool FILL>
__global__ void __launch_bounds__(128) k_slot_prep_w(PrepArgs a) {
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (slot >= a.n_slots) return;
    const uint32_t reg = a.slot_region[slot];
    if (a.rstate[reg].status != 0) return;
    const lcr_region R = a.regions[reg];
    const uint32_t read = R.read_begin + (slot - a.slot_off[reg]);
    const uint64_t c0 = a.cig_off[read], c1 = a.cig_off[read + 1];
    const uint64_t s0 = a.seq_off[read];
    const int64_t seq_len = (int64_t)(a.seq_off[read + 1] - s0);
    const uint8_t *seq = a.seq + s0;
    const int64_t lead = (c1 > c0 && (a.cigar[c0] & 0xf) == 4) ? (int64_t)(a.cigar[c0] >> 4) : 0; /* leading_softclips */
    const int64_t trail = (c1 > c0 && (a.cigar[c1 - 1] & 0xf) == 4) ? (int64_t)(a.cigar[c1 - 1] >> 4) : 0;
    const int64_t rb = seq_len - trail;
    const int64_t dend = (int64_t)(a.P.distance_to_read_end > 0x3fffffffu ? 0x3fffffffu : a.P.distance_to_read_end);
    const bool ont = a.P.platform == 1;
    uint32_t sflags;
    uint64_t runs[LCR_SLOT_RUNS];
#pragma unroll
    for (int i = 0; i < LCR_SLOT_RUNS; ++i) runs[i] = 0;
    if (!FILL) {
        /* util.rs:652-668 */
        const uint16_t fl = a.flag[read];
        bool pass = !((int32_t)a.mapq[read] < a.P.min_mapq || (uint64_t)seq_len < (uint64_t)a.P.min_read_length || (fl & 0x4) || (fl & 0x100) || (fl & 0x800));
        const float de = a.de[read];
        if (!(de != de) && de >= a.P.divergence) pass = false;
        /* fetch((chr, start, end)): pos < end && bam_endpos > start on the region's own numbers */
        long long rlen = 0;
        for (uint64_t c = c0 + lane; c < c1; c += 32) {
            const uint32_t op = a.cigar[c];
            if (is_ref_consuming(op & 0xf)) rlen += op >> 4;
        }
        for (int o = 16; o; o >>= 1) rlen += __shfl_xor_sync(0xffffffffu, rlen, o);
        const int64_t p = a.pos[read];
        const bool in_window = p < (int64_t)R.end && p + (rlen ? rlen : 1) > (int64_t)R.start;
        sflags = (pass && in_window) ? 1 : 0;
        if (sflags && !ont && dend > 0) { /* the homopolymer-run table of the read ends: lane 0 scans, as k_slot_prep does */
            const int64_t polya = (int64_t)(a.P.polya_tail_length > 0x3fffffffu ? 0x3fffffffu : a.P.polya_tail_length);
            if (polya < 2 || rb < lead) sflags |= 6;
            else {
                uint32_t nrun = 0;
                if (lane == 0) {
                    const int64_t centre[2] = {lead, rb};
                    int64_t scanned_to = -1;
                    for (int z = 0; z < 2; ++z) {
                        int64_t lo = centre[z] - dend + 1 - polya, hi = centre[z] + dend + polya;
                        if (lo < 0) lo = 0;
                        if (hi > seq_len) hi = seq_len;
                        int64_t run = 0;
                        uint8_t prev = 0;
                        for (int64_t i = lo; i <= hi; ++i) {
                            const uint8_t b = i < hi ? __ldg(seq + i) : (uint8_t)0;
                            const bool letter = b == 'A' || b == 'C' || b == 'G' || b == 'T';
                            if (letter && b == prev) { run++; continue; }
                            if (run >= polya && i > scanned_to) {
                                if (nrun < LCR_SLOT_RUNS) runs[nrun] = ((uint64_t)(i - run) << 32) | ((uint64_t)(run > 0xffffff ? 0xffffff : run) << 8) | (uint64_t)prev;
                                nrun++;
                            }
                            run = letter ? 1 : 0;
                            prev = b;
                        }
                        if (hi > scanned_to) scanned_to = hi;
                    }
                }
                nrun = __shfl_sync(0xffffffffu, nrun, 0);
#pragma unroll
                for (int i = 0; i < LCR_SLOT_RUNS; ++i) runs[i] = __shfl_sync(0xffffffffu, runs[i], 0);
                if (nrun) sflags |= 2;
                if (nrun > LCR_SLOT_RUNS) sflags |= 4;
            }
        }
        if (lane == 0) {
            a.slot_flags[slot] = (uint8_t)sflags;
            if (sflags & 2) {
#pragma unroll
                for (int i = 0; i < LCR_SLOT_RUNS; ++i) a.slot_runs[(size_t)slot * LCR_SLOT_RUNS + i] = runs[i];
            }
        }
    } else {
        sflags = a.slot_flags[slot];
        if ((sflags & 6) == 2) {
#pragma unroll
            for (int i = 0; i < LCR_SLOT_RUNS; ++i) runs[i] = a.slot_runs[(size_t)slot * LCR_SLOT_RUNS + i];
        }
    }
    if (!(sflags & 1)) return;

    const int64_t vec_size = (int64_t)R.end - (int64_t)R.start;
    const int64_t fv_start = (int64_t)R.start - 1;
    const uint32_t tb = a.tile_base[reg];
    const uint8_t *ref = a.ref_table[R.tid] + fv_start;
    uint32_t rowtyp_c;
    {
        const int strand = (a.flag[read] & 0x10) ? 1 : 0;
        const int8_t ts = a.ts[read];
        uint32_t tcode = 0;
        if (ts == '+') tcode = strand == 0 ? 1u : 2u;
        else if (ts == '-') tcode = strand == 0 ? 2u : 1u;
        rowtyp_c = (strand == 0 ? 4u : 0u) | (tcode << 3);
    }
    int64_t zlo[2] = {lead - dend + 1, rb - dend + 1}, zhi[2] = {lead + dend - 1, rb + dend - 1};
    if (zlo[1] < zlo[0]) { int64_t t = zlo[0]; zlo[0] = zlo[1]; zlo[1] = t; t = zhi[0]; zhi[0] = zhi[1]; zhi[1] = t; }
    const int mask_mode = dend <= 0 ? 0 : ont ? 1 : (sflags & 4) ? 2 : (sflags & 2) ? 3 : 0;
    const int64_t polya = (int64_t)a.P.polya_tail_length;

    /* the unmasked stretches of the aligned piece [pa, pb) (read coordinates) whose first base sits on region position ts:
       WRITE stores them from segment slot w on; returns how many there are */
    auto piece = [&](bool write, int64_t pa, int64_t pb, int64_t ts, uint32_t colr, uint32_t w) -> uint32_t {
        uint32_t n = 0;
        int64_t start = pa;
        auto emit = [&](int64_t x, int64_t y) {
            if (y <= x) return;
            if (write) {
                LcrSeg s;
                s.spos = s0 + (uint64_t)x; s.row_typ = rowtyp_c | SEG_M; s.col = (uint16_t)(colr + (uint32_t)(x - pa)); s.len = (uint16_t)(y - x);
                a.segs[w + n] = s;
            }
            n++;
        };
        if (mask_mode == 1) {
            for (int k = 0; k < 2; ++k) {
                const int64_t zl = zlo[k] > start ? zlo[k] : start, zh = zhi[k] < pb - 1 ? zhi[k] : pb - 1;
                if (zl <= zh) { emit(start, zl); start = zh + 1; }
            }
        } else if (mask_mode == 2) {
            for (int k = 0; k < 2; ++k) {
                const int64_t zl = zlo[k] > start ? zlo[k] : start, zh = zhi[k] < pb - 1 ? zhi[k] : pb - 1;
                for (int64_t rp = zl; rp <= zh; ++rp)
                    if (base_masked(a.P, seq, rp, seq_len, lead, trail, ref[ts + (rp - pa)])) { emit(start, rp); start = rp + 1; }
            }
        } else if (mask_mode == 3) {
            int64_t from = pa;
#pragma unroll
            for (int j = 0; j < LCR_SLOT_RUNS; ++j) {
                const int64_t rs = (int64_t)(runs[j] >> 32), rn = (int64_t)((runs[j] >> 8) & 0xffffffu);
                if (rn == 0) continue;
                const int64_t zl = rs - 1 > from ? rs - 1 : from, zh = rs + rn < pb - 1 ? rs + rn : pb - 1;
                for (int64_t rp = zl; rp <= zh; ++rp) {
                    const int64_t d0 = rp - lead, d1 = rp - rb;
                    if (!((d0 < 0 ? -d0 : d0) < dend || (d1 < 0 ? -d1 : d1) < dend)) continue;
                    const uint8_t rbase = ref[ts + (rp - pa)];
                    const int64_t wlo = rp - polya > 0 ? rp - polya : 0, whi = rp + polya + 1 < seq_len ? rp + polya + 1 : seq_len;
                    bool m = false;
#pragma unroll
                    for (int q = 0; q < LCR_SLOT_RUNS; ++q) {
                        const int64_t qs = (int64_t)(runs[q] >> 32), qn = (int64_t)((runs[q] >> 8) & 0xffffffu), qe = qs + qn;
                        if (qn == 0 || (uint8_t)(runs[q] & 0xffu) == rbase) continue;
                        const bool anchored = (rp - 1 >= qs && rp - 1 < qe) || (rp + 1 >= qs && rp + 1 < qe);
                        const int64_t ov = (qe < whi ? qe : whi) - (qs > wlo ? qs : wlo);
                        if (anchored && ov >= polya) m = true;
                    }
                    if (m) { emit(start, rp); start = rp + 1; }
                }
                if (zh + 1 > from) from = zh + 1;
            }
        }
        emit(start, pb);
        return n;
    };

    /* carried from batch to batch (the same in every lane) */
    long long fpos_c = (long long)a.pos[read] - fv_start, rpos_c = lead;
    long long last_tile_c = -1;
    uint32_t seg_c = FILL ? a.slot_seg_off[slot] : 0;  /* next segment slot of the read */
    uint32_t open_k = 0xffffffffu, open_begin = 0;      /* FILL: the read's item that later ops may still extend */
    uint32_t n_bases = 0;
    bool bad = false;
    for (uint64_t cbase = c0; cbase < c1 && !bad; cbase += 32) {
        const uint64_t ci = cbase + lane;
        const uint32_t op = ci < c1 ? a.cigar[ci] : 4u; /* padding: a zero-length soft clip */
        const uint32_t opc = op & 0xf;
        const long long len = op >> 4;
        const bool is_m = opc == 0 || opc == 7 || opc == 8, is_dn = opc == 2 || opc == 3, is_i = opc == 1, is_sh = opc == 4 || opc == 5;
        const unsigned badmask = __ballot_sync(0xffffffffu, !(is_m || is_dn || is_i || is_sh)); /* util.rs:943-945 panics */
        const uint32_t nvalid = badmask ? (uint32_t)__ffs(badmask) - 1u : 32u;
        const bool on = lane < nvalid;
        const long long rl = (on && (is_m || is_dn)) ? len : 0, ql = (on && (is_m || is_i)) ? len : 0;
        long long rsum = rl, qsum = ql;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long r2 = __shfl_up_sync(0xffffffffu, rsum, o), q2 = __shfl_up_sync(0xffffffffu, qsum, o);
            if ((int)lane >= o) { rsum += r2; qsum += q2; }
        }
        const long long lo = fpos_c + rsum - rl, hi = lo + rl, rp0 = rpos_c + qsum - ql;
        /* the part of the op inside the region and the tiles where it takes part in an item */
        const long long a0 = lo < 0 ? 0 : lo, b0 = hi < vec_size ? hi : vec_size;
        const bool inside = on && rl > 0 && hi > 0 && lo < vec_size && b0 > a0;
        long long t0 = 0, t1 = -1; /* tiles touched */
        if (inside) { t0 = a0 / LCR_TILE; t1 = (b0 - 1) / LCR_TILE; }
        /* whole-tile intron covers register nothing: for an N op only a partial first / last tile counts */
        auto tile_end_of = [&](long long t) { return (t + 1) * LCR_TILE < vec_size ? (t + 1) * LCR_TILE : vec_size; };
        auto registers = [&](long long t) -> bool {
            if (opc != 3) return true;
            const long long ts = a0 > t * LCR_TILE ? a0 : t * LCR_TILE, te = b0 < tile_end_of(t) ? b0 : tile_end_of(t);
            return !(ts == t * LCR_TILE && te == tile_end_of(t));
        };
        long long tl = -1; /* last tile this op registers in */
        if (inside) {
            if (registers(t1)) tl = t1;
            else if (t1 > t0 && registers(t0)) tl = t0; /* N op: only the first tile is partial */
        }
        /* last registered tile of the ops before this one: exclusive running maximum */
        long long lt = tl;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long v = __shfl_up_sync(0xffffffffu, lt, o);
            if ((int)lane >= o && v > lt) lt = v;
        }
        long long last_before = __shfl_up_sync(0xffffffffu, lt, 1);
        if (lane == 0 || last_before < last_tile_c) last_before = last_tile_c;
        const long long batch_last = __shfl_sync(0xffffffffu, lt, 31);

        /* pass 1: segments of this op (all its registered tiles), new items, aligned bases */
        uint32_t nseg = 0, nnew = 0;
        bool mybad = false;
        if (inside) {
            for (long long t = t0; t <= t1; ++t) {
                if (!registers(t)) continue;
                const long long ts = a0 > t * LCR_TILE ? a0 : t * LCR_TILE, te = b0 < tile_end_of(t) ? b0 : tile_end_of(t);
                if (t > last_before) nnew++;
                if (!is_m) { nseg++; continue; }
                const long long pa = rp0 + (ts - lo), pb = rp0 + (te - lo);
                if (pb > seq_len) { mybad = true; break; }
                n_bases += (uint32_t)(te - ts);
                nseg += piece(false, pa, pb, ts, (uint32_t)(ts - t * LCR_TILE), 0);
            }
        }
        const unsigned mybadmask = __ballot_sync(0xffffffffu, mybad);
        uint32_t sincl = nseg;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, sincl, o);
            if ((int)lane >= o) sincl += v;
        }
        const uint32_t seg_first = seg_c + sincl - nseg; /* first segment slot of this op */
        const uint32_t seg_total = __shfl_sync(0xffffffffu, sincl, 31);

        if (!FILL) {
            if (inside && !mybad)
                for (long long t = t0; t <= t1; ++t)
                    if (registers(t) && t > last_before && atomicAdd(&a.tile_count[tb + (uint32_t)t], 1u) == 255u) *a.deep_flag = 1u;
        } else {
            /* the open item of the ops before this one: last lane before me that created an item, else the carry */
            uint32_t my_open_k = 0xffffffffu, my_open_begin = 0; /* the last item this op creates */
            uint32_t w = seg_first;
            uint32_t prev_k = 0xffffffffu, prev_begin = 0;        /* the item this op is currently extending / has just created */
            const unsigned creators = __ballot_sync(0xffffffffu, nnew > 0 && !mybad);
            /* placeholders filled after the loop below needs them: fetch the inherited open item first */
            const unsigned before = creators & ((1u << lane) - 1u);
            /* every creating lane publishes its last created item after the loop; the inherited one therefore needs a second
               exchange: do the tile loop first for the lane's own items, then close the inherited item with the first new begin */
            uint32_t first_new_begin = 0;
            bool have_first = false;
            if (inside && !mybad) {
                for (long long t = t0; t <= t1; ++t) {
                    const long long ts = a0 > t * LCR_TILE ? a0 : t * LCR_TILE, te = b0 < tile_end_of(t) ? b0 : tile_end_of(t);
                    if (!registers(t)) { atomicAdd(&a.tile_full_n[tb + (uint32_t)t], 1u); continue; }
                    if (t > last_before) {
                        const uint32_t k = a.tile_off[tb + (uint32_t)t] + atomicAdd(&a.tile_count[tb + (uint32_t)t], 1u);
                        if (!have_first) { have_first = true; first_new_begin = w; }
                        else a.item_segs[prev_k] = make_uint2(prev_begin, w - prev_begin); /* my previous new item is complete */
                        prev_k = k; prev_begin = w;
                    }
                    if (!is_m) {
                        LcrSeg s;
                        s.spos = 0; s.row_typ = rowtyp_c | (opc == 2 ? SEG_D : SEG_N); s.col = (uint16_t)(ts - t * LCR_TILE); s.len = (uint16_t)(te - ts);
                        a.segs[w] = s;
                        w++;
                    } else {
                        const long long pa = rp0 + (ts - lo), pb = rp0 + (te - lo);
                        w += piece(true, pa, pb, ts, (uint32_t)(ts - t * LCR_TILE), w);
                    }
                }
                if (have_first) { my_open_k = prev_k; my_open_begin = prev_begin; }
            }
            /* close the inherited open item at the first new item of this op */
            const int src = before ? 31 - __clz(before) : 0;
            const uint32_t inh_k = __shfl_sync(0xffffffffu, my_open_k, src), inh_begin = __shfl_sync(0xffffffffu, my_open_begin, src);
            if (have_first) {
                const uint32_t ok = before ? inh_k : open_k, ob = before ? inh_begin : open_begin;
                if (ok != 0xffffffffu) a.item_segs[ok] = make_uint2(ob, first_new_begin - ob);
            }
            if (creators) {
                const int lastc = 31 - __clz(creators);
                open_k = __shfl_sync(0xffffffffu, my_open_k, lastc);
                open_begin = __shfl_sync(0xffffffffu, my_open_begin, lastc);
            }
        }
        seg_c += seg_total;
        fpos_c += __shfl_sync(0xffffffffu, rsum, 31);
        rpos_c += __shfl_sync(0xffffffffu, qsum, 31);
        if (batch_last > last_tile_c) last_tile_c = batch_last;
        if (mybadmask || badmask) bad = true;
    }
    if (bad) {
        if (lane == 0) atomicMin(&a.rstate[reg].status, (int32_t)LCR_ERR_BAD_CIGAR);
        n_bases = 0;
    }
    if (!FILL) {
        if (lane == 0) a.slot_segs[slot] = seg_c;
    } else {
        if (lane == 0 && open_k != 0xffffffffu) a.item_segs[open_k] = make_uint2(open_begin, seg_c - open_begin);
        n_bases = __reduce_add_sync(0xffffffffu, n_bases);
        if (lane == 0 && n_bases) atomicAdd((unsigned long long *)&a.stats->n_aligned_bases, (unsigned long long)n_bases);
    }
}

/* ------------------------------------------------------------------------- *
 * Tile pileup, version 3: segments -> one-hot row planes -> carry-save column sums.
 *
 * k_pileup_tile (CTA per tile) stages up to ROWS items as two byte planes per (row, column):
 *     plane X   bit 0-3  base is A,C,G,T          bit 4-7  ... and base quality >= min_baseq
 *     plane Y   bit 0-3  A,C,G,T on a forward read; bit 4 / 5 transcript strand forward / reverse
 *               (util.rs:803-819, any base letter); bit 6 deletion; bit 7 intron
 * Whole column words (4 columns) of a segment are produced one lane per 16-byte block of the read: aligned
 * 128-bit loads of seq and qual (issued one block ahead), bytes rotated to the column alignment with PRMT,
 * codes built four columns at a time.  The up to three columns before / after the whole words of a segment
 * are written byte-wise by one lane per segment.  Then every thread sums one 32-bit column word (4 columns x
 * 8 indicators) over the rows with a Harley-Seal carry-save adder tree: ~2.4 logic instructions per row for
 * 32 counters.  Counters are unpacked once per tile (once per 255 rows on deep tiles).
 * ------------------------------------------------------------------------- */
#define PT_THREADS 256
#define PT_WORDS (LCR_TILE / 4)
#define PT_TAB(SEGS) (((SEGS) * (LCR_TILE / 16 + 1)) / 8 + 8) /* one entry per 8 blocks */

struct PreCand { /* a site that passed every count-based filter; its likelihood is computed by k_site_ll */
    uint32_t tile, col;
    uint32_t cnt[4], pass[4], fwd[4], ts[2], d, n;
};

struct __align__(16) LcrTileDesc { /* 48 B, one per tile (k_tile_desc) */
    const uint8_t *ref;     /* reference base of column 0 */
    uint64_t pos_g;         /* index of column 0 in the debug planes */
    uint32_t it0, n_items;  /* items of the tile */
    uint32_t reserved[2];
    uint32_t reg, npos;
    uint32_t full_n;        /* introns covering the whole tile */
    int32_t status;         /* of the region */
};

struct DescArgs {
    uint32_t n_tiles;
    const lcr_region *regions;
    const uint32_t *tile_base, *tile_region, *tile_off, *tile_full_n;
    const uint64_t *pos_off;
    const uint8_t *const *ref_table;
    const LcrRegionState *rstate;
    LcrTileDesc *desc;
};

__global__ void k_tile_desc(DescArgs a) {
    const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= a.n_tiles) return;
    const uint32_t reg = a.tile_region[tile];
    const lcr_region R = a.regions[reg];
    const int64_t vec_size = (int64_t)R.end - (int64_t)R.start;
    const int64_t tile_start = (int64_t)(tile - a.tile_base[reg]) * LCR_TILE;
    const int64_t tile_end = tile_start + LCR_TILE < vec_size ? tile_start + LCR_TILE : vec_size;
    LcrTileDesc d;
    d.status = a.rstate[reg].status;
    d.ref = d.status == 0 ? a.ref_table[R.tid] + ((int64_t)R.start - 1) + tile_start : nullptr;
    d.pos_g = a.pos_off[reg] + (uint64_t)tile_start;
    d.it0 = a.tile_off[tile]; d.n_items = a.tile_off[tile + 1] - d.it0;
    d.reserved[0] = 0; d.reserved[1] = 0;
    d.reg = reg; d.npos = (uint32_t)(tile_end - tile_start);
    d.full_n = a.tile_full_n[tile];
    a.desc[tile] = d;
}

__device__ __forceinline__ uint32_t lop3_xor3(uint32_t x, uint32_t y, uint32_t z) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(x), "r"(y), "r"(z));
    return r;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t x, uint32_t y, uint32_t z) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(r) : "r"(x), "r"(y), "r"(z));
    return r;
}
/* carry-save adder: (hi, lo) = x + y + z per bit */
#define CSA(hi, lo, x, y, z) do { const uint32_t x__ = (x), y__ = (y), z__ = (z); hi = lop3_maj(x__, y__, z__); lo = lop3_xor3(x__, y__, z__); } while (0)

/* PTX prmt in its generic form: bit 3 of a selector nibble replicates the sign bit of the selected byte
   (__byte_perm is specified to ignore that bit) */
__device__ __forceinline__ uint32_t prmt_sign(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(r) : "r"(x));
    return r;
}

/* 4 read bases + 4 qualities (column order) -> plane bytes */
__device__ __forceinline__ void onehot4(uint32_t s, uint32_t q, uint32_t minq4, uint32_t pass_allow, uint32_t fmask, uint32_t tsb, uint32_t &x, uint32_t &y) {
    /* PRMT as an 8-entry table on the low three bits of each letter: A=..001 C=..011 T=..100 G=..111 */
    const uint32_t t = s & 0x07070707u;
    const uint32_t u = t | (t >> 4);
    const uint32_t sel = __byte_perm(u, 0, 0x4420);
    uint32_t oh = __byte_perm(0x02000100u, 0x04000008u, sel);
    const uint32_t canon = __byte_perm(0x43004100u, 0x47000054u, sel);
    const uint32_t d = (s & 0xdfdfdfdfu) ^ canon;                       /* non-zero byte: not exactly that letter (either case) */
    const uint32_t nz = ((d & 0x7f7f7f7fu) + 0x7f7f7f7fu) | d;
    oh &= ~prmt_sign(nz);
    const uint32_t ge = ((((q & 0x7f7f7f7fu) | 0x80808080u) - minq4) | q);
    const uint32_t pm = prmt_sign(ge) & pass_allow;
    x = oh | ((oh << 4) & pm);
    y = (oh & fmask) | tsb;
}

struct PileArgs {
    lcr_params P;
    const lcr_region *regions;
    const uint32_t *slot_off, *slot_region, *tile_base, *tile_region;
    const uint16_t *flag;
    const int8_t *ts;
    const uint64_t *seq_off, *cig_off;
    const uint8_t *seq, *qual;
    const uint32_t *cigar;
    const uint8_t *const *ref_table;
    const uint32_t *tile_off;
    const LcrTileDesc *desc;
    const uint2 *item_segs;
    const LcrSeg *segs;
    const LcrDeviceTables *tables;
    LcrRegionState *rstate;
    lcr_candidate *cand;
    uint64_t *cand_key;
    uint32_t cand_cap;
    uint32_t *cand_count;
    lcr_stats *stats;
    uint32_t *pl_acgt, *pl_fwd, *pl_d, *pl_n, *pl_ts; /* debug planes or null */
    PreCand *pre;
    uint32_t pre_cap;
    uint32_t *pre_count;
};

struct PtBlock { /* one 16-byte block of a segment, loads in flight */
    uint4 sv, qv;
    uint32_t s4w, q4w;
    uint32_t z;        /* row_typ of the segment (type, strand, transcript strand) */
    uint32_t saddr;    /* shared-space byte address of the first column word of the block in plane X */
    uint32_t rel, span, e;
};

/* 32-bit shared-space stores with an immediate offset (the generic-pointer form makes ptxas rebuild the shared window
   base for every store when registers are tight) */
template <int OFF>
__device__ __forceinline__ void sts32(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.b32 [%0+%2], %1;" ::"r"(saddr), "r"(v), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sts8_if(uint32_t saddr, uint32_t v, bool p) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q st.shared.u8 [%0+%2], %1;\n\t}" ::"r"(saddr), "r"(v), "n"(OFF), "r"((uint32_t)p) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sts32_if(uint32_t saddr, uint32_t v, bool p) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q st.shared.b32 [%0+%2], %1;\n\t}" ::"r"(saddr), "r"(v), "n"(OFF), "r"((uint32_t)p) : "memory");
}

template <int ROWS, int SEGS>
constexpr size_t pt_smem_bytes(bool deep) {
    return sizeof(uint32_t) * (2 * ROWS * PT_WORDS + SEGS * 4 + SEGS + 4 + (PT_TAB(SEGS) + 1) / 2 + (deep ? 16 * LCR_TILE : 0));
}

template <bool DEEP, int ROWS, int MINB, int SEGS>
__global__ void __launch_bounds__(PT_THREADS, MINB) k_pileup_tile(PileArgs a) {
    constexpr int PT_SEGS = SEGS;
    static_assert(SEGS % PT_THREADS == 0, "whole segments per thread in the stage");
    static_assert(ROWS % 16 == 0 && ROWS >= 16, "the column sums read whole blocks of 16 rows");
    extern __shared__ __align__(16) uint32_t pt_smem[];
    uint32_t *planes = pt_smem;                                           /* [2][ROWS][PT_WORDS] */
    uint4 *s_seg = reinterpret_cast<uint4 *>(pt_smem + 2 * ROWS * PT_WORDS); /* [PT_SEGS] staged segments */
    uint32_t *s_choff = pt_smem + 2 * ROWS * PT_WORDS + PT_SEGS * 4;      /* [PT_SEGS + 1] first block of every staged segment */
    uint16_t *s_tab = reinterpret_cast<uint16_t *>(s_choff + PT_SEGS + 4); /* [PT_TAB] segment of every 8th block */
    uint32_t *s_out32 = s_choff + PT_SEGS + 4 + (PT_TAB(SEGS) + 1) / 2;   /* DEEP: [16][LCR_TILE] */
    __shared__ uint32_t s_wsum[PT_THREADS / 32];

    const uint32_t tile = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    LcrTileDesc D;
    {
        const uint4 *dp = reinterpret_cast<const uint4 *>(a.desc + tile);
        uint4 *dd = reinterpret_cast<uint4 *>(&D);
        dd[0] = __ldg(dp); dd[1] = __ldg(dp + 1); dd[2] = __ldg(dp + 2);
    }
    if ((D.n_items > 255u) != DEEP || D.status != 0) return;
    const uint32_t n_items = D.n_items, npos = D.npos;
    if (DEEP) {
        for (uint32_t i = tid; i < 16 * LCR_TILE; i += PT_THREADS) s_out32[i] = 0;
        __syncthreads();
    }

    const uint32_t minq = (uint32_t)a.P.min_baseq;
    const uint32_t minq4 = (minq > 30u ? 0u : minq) * 0x01010101u;
    const uint32_t pass_allow = minq > 30u ? 0u : 0xffffffffu;
    const uint8_t *seqp = a.seq, *qualp = a.qual;
    static_assert(ROWS <= 64, "the item scan below handles two items per lane of one warp");
    __shared__ uint32_t s_ioff[65]; /* first staged-segment index of every item of the batch, and the total */
    __shared__ uint32_t s_ibeg[64]; /* first segment of every item of the batch */

    /* carry-save state of this thread's column word: plane (tid / PT_WORDS), word (tid % PT_WORDS) */
    uint32_t ones = 0, twos = 0, fours = 0, eights = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
    uint32_t acc_rows = 0;
    const uint32_t my_plane = tid / PT_WORDS, my_word = tid % PT_WORDS;
    uint32_t cnt8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cnt8[i] = 0;

    auto flush = [&]() { /* bit-sliced counters -> 4 x 8-bit fields per indicator */
        const uint32_t lv[8] = {ones, twos, fours, eights, s4, s5, s6, s7};
#pragma unroll
        for (int i = 0; i < 8; ++i) cnt8[i] = 0;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            if ((acc_rows >> l) != 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) cnt8[i] |= (i >= l ? lv[l] >> (i - l) : lv[l] << (l - i)) & (0x01010101u << l); /* bit l of 4 counters */
            }
        }
        ones = twos = fours = eights = s4 = s5 = s6 = s7 = 0;
        acc_rows = 0;
    };

    for (uint32_t row_base = 0; row_base < n_items; row_base += ROWS) {
        const uint32_t nrow = (n_items - row_base) < (uint32_t)ROWS ? (n_items - row_base) : (uint32_t)ROWS;
        if (DEEP && acc_rows + nrow > 255u) {
            flush();
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s_out32[(my_plane * 8 + i) * LCR_TILE + my_word * 4 + j] += (cnt8[i] >> (8 * j)) & 0xffu;
        }
        if (row_base) __syncthreads(); /* the previous batch's column sums are done with the planes */
        /* the batch's items and where their segments start in the concatenated list (loads issued before the zero fill) */
        uint2 e0 = make_uint2(0, 0), e1 = make_uint2(0, 0);
        if (warp == 0) {
            if (lane < nrow) e0 = a.item_segs[D.it0 + row_base + lane];
            if (lane + 32 < nrow) e1 = a.item_segs[D.it0 + row_base + lane + 32];
        }
        {
            const uint4 z = make_uint4(0, 0, 0, 0);
            uint4 *p4 = reinterpret_cast<uint4 *>(planes);
            const uint32_t nz16 = ((nrow + 15u) & ~15u) * (PT_WORDS / 4); /* whole blocks of 16 rows */
            for (uint32_t i = tid; i < nz16; i += PT_THREADS) {
                p4[i] = z;
                p4[ROWS * (PT_WORDS / 4) + i] = z;
            }
        }
        if (warp == 0) {
            uint32_t i0 = e0.y, i1 = e1.y;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v0 = __shfl_up_sync(0xffffffffu, i0, o), v1 = __shfl_up_sync(0xffffffffu, i1, o);
                if ((int)lane >= o) { i0 += v0; i1 += v1; }
            }
            const uint32_t t0 = __shfl_sync(0xffffffffu, i0, 31), t1 = __shfl_sync(0xffffffffu, i1, 31);
            s_ioff[lane] = i0 - e0.y; s_ibeg[lane] = e0.x;
            s_ioff[lane + 32] = t0 + i1 - e1.y; s_ibeg[lane + 32] = e1.x;
            if (lane == 0) s_ioff[64] = t0 + t1;
        }
        __syncthreads(); /* also: the previous batch's stage is consumed */
        const uint32_t n_seg_batch = s_ioff[64];
        for (uint32_t sb = 0; sb < n_seg_batch; sb += PT_SEGS) {
            const uint32_t ns = (n_seg_batch - sb) < PT_SEGS ? (n_seg_batch - sb) : PT_SEGS;
            if (sb) __syncthreads(); /* previous stage consumed */
            /* stage the segments of this batch's rows and scan their block counts */
            uint32_t nch[PT_SEGS / PT_THREADS], mysum = 0;
#pragma unroll
            for (int qd = 0; qd < PT_SEGS / PT_THREADS; ++qd) {
                const uint32_t i = tid * (PT_SEGS / PT_THREADS) + qd;
                uint32_t n = 0;
                if (i < ns) {
                    const uint32_t idx = sb + i;
                    uint32_t j = 0; /* the item of staged segment idx: last j with s_ioff[j] <= idx */
#pragma unroll
                    for (uint32_t step = 32; step; step >>= 1)
                        if (j + step < nrow && s_ioff[j + step] <= idx) j += step;
                    uint4 raw = __ldg(reinterpret_cast<const uint4 *>(a.segs + s_ibeg[j] + (idx - s_ioff[j])));
                    const uint32_t col = raw.w & 0xffffu, len = raw.w >> 16;
                    const uint32_t wlo = (col + 3u) >> 2, whi = (col + len) >> 2; /* whole column words [wlo, whi) */
                    if (whi > wlo) {
                        const uint32_t nw = whi - wlo;
                        if ((raw.z & 3u) == SEG_M) n = ((((raw.x + 4u * wlo - col) & 15u) + 4u * nw - 4u) >> 4) + 1u;
                        else n = (((wlo & 3u) + nw - 1u) >> 2) + 1u;
                    }
                    raw.z = (raw.z & 0xffu) | (j << 8);
                    s_seg[i] = raw;
                }
                nch[qd] = n;
                mysum += n;
            }
            uint32_t incl = mysum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += v;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads(); /* also: planes zeroed */
            uint32_t wbase = 0, total = 0;
#pragma unroll
            for (int w = 0; w < PT_THREADS / 32; ++w) {
                const uint32_t v = s_wsum[w];
                if (w < (int)warp) wbase += v;
                total += v;
            }
            {
                uint32_t run = wbase + incl - mysum;
#pragma unroll
                for (int qd = 0; qd < PT_SEGS / PT_THREADS; ++qd) {
                    const uint32_t i = tid * (PT_SEGS / PT_THREADS) + qd;
                    s_choff[i] = run;
                    for (uint32_t m = (run + 7u) >> 3; m * 8u < run + nch[qd]; ++m) s_tab[m] = (uint16_t)i;
                    run += nch[qd];
                }
                if (tid == PT_THREADS - 1) s_choff[PT_SEGS] = run;
            }
            __syncthreads();
            /* whole column words: one lane per 16-byte block of a segment, loads issued one block ahead */
            const uint32_t planes_s = (uint32_t)__cvta_generic_to_shared(planes);
            auto fetch = [&](uint32_t g, PtBlock &b) {
                uint32_t k = s_tab[g >> 3];
                while (s_choff[k + 1] <= g) ++k;
                const uint32_t c = g - s_choff[k];
                const uint4 raw = s_seg[k];
                const uint32_t col = raw.w & 0xffffu, len = raw.w >> 16;
                const uint32_t wlo = (col + 3u) >> 2, whi = (col + len) >> 2;
                uint32_t W0; /* first column word of this block */
                b.z = raw.z;
                b.span = whi - wlo;
                if ((raw.z & 3u) == SEG_M) {
                    const uint64_t spos = (((uint64_t)raw.y << 32) | raw.x) + (uint64_t)(4u * wlo - col); /* base of column 4 * wlo */
                    const uint32_t al = (uint32_t)spos & 15u;
                    const uint64_t blk = (spos & ~(uint64_t)15) + 16ull * c;
                    b.e = al & 3u;
                    W0 = wlo + 4u * c - (al >> 2);
                    b.sv = __ldg(reinterpret_cast<const uint4 *>(seqp + blk));
                    b.qv = __ldg(reinterpret_cast<const uint4 *>(qualp + blk));
                    b.s4w = __ldg(reinterpret_cast<const uint32_t *>(seqp + blk + 16));
                    b.q4w = __ldg(reinterpret_cast<const uint32_t *>(qualp + blk + 16));
                } else {
                    b.e = 0;
                    W0 = (wlo & ~3u) + 4u * c;
                }
                b.rel = W0 - wlo; /* word W0 + t is whole iff rel + t < span (unsigned) */
                b.saddr = planes_s + (((raw.z >> 8) * PT_WORDS + W0) << 2);
            };
            auto process = [&](const PtBlock &cur) {
                constexpr int YOFF = ROWS * PT_WORDS * 4;
                const uint32_t typ = cur.z & 3u;
                uint32_t x[4], y[4];
                if (typ == SEG_M) {
                    const uint32_t fmask = (cur.z & 4u) ? 0x0f0f0f0fu : 0u;
                    const uint32_t tsb = ((cur.z >> 3) & 3u) * 0x10101010u; /* code 1 -> bit 4, code 2 -> bit 5 */
                    const uint32_t rot = 0x3210u + 0x1111u * cur.e;
                    onehot4(__byte_perm(cur.sv.x, cur.sv.y, rot), __byte_perm(cur.qv.x, cur.qv.y, rot), minq4, pass_allow, fmask, tsb, x[0], y[0]);
                    onehot4(__byte_perm(cur.sv.y, cur.sv.z, rot), __byte_perm(cur.qv.y, cur.qv.z, rot), minq4, pass_allow, fmask, tsb, x[1], y[1]);
                    onehot4(__byte_perm(cur.sv.z, cur.sv.w, rot), __byte_perm(cur.qv.z, cur.qv.w, rot), minq4, pass_allow, fmask, tsb, x[2], y[2]);
                    onehot4(__byte_perm(cur.sv.w, cur.s4w, rot), __byte_perm(cur.qv.w, cur.q4w, rot), minq4, pass_allow, fmask, tsb, x[3], y[3]);
                } else {
                    const uint32_t v = typ == SEG_D ? 0x40404040u : 0x80808080u;
                    x[0] = x[1] = x[2] = x[3] = 0;
                    y[0] = y[1] = y[2] = y[3] = v;
                }
                const bool isM = typ == SEG_M;
                if (cur.span >= 4u && cur.rel <= cur.span - 4u) { /* all four words whole */
                    if (isM) { sts32<0>(cur.saddr, x[0]); sts32<4>(cur.saddr, x[1]); sts32<8>(cur.saddr, x[2]); sts32<12>(cur.saddr, x[3]); }
                    sts32<YOFF>(cur.saddr, y[0]); sts32<YOFF + 4>(cur.saddr, y[1]); sts32<YOFF + 8>(cur.saddr, y[2]); sts32<YOFF + 12>(cur.saddr, y[3]);
                } else {
                    const bool p0 = cur.rel < cur.span, p1 = cur.rel + 1u < cur.span, p2 = cur.rel + 2u < cur.span, p3 = cur.rel + 3u < cur.span;
                    sts32_if<0>(cur.saddr, x[0], p0 && isM); sts32_if<4>(cur.saddr, x[1], p1 && isM);
                    sts32_if<8>(cur.saddr, x[2], p2 && isM); sts32_if<12>(cur.saddr, x[3], p3 && isM);
                    sts32_if<YOFF>(cur.saddr, y[0], p0); sts32_if<YOFF + 4>(cur.saddr, y[1], p1);
                    sts32_if<YOFF + 8>(cur.saddr, y[2], p2); sts32_if<YOFF + 12>(cur.saddr, y[3], p3);
                }
            };
            { /* two blocks in flight per lane, alternating buffers */
                PtBlock b0, b1;
                uint32_t g = tid;
                if (g < total) fetch(g, b0);
                while (g < total) {
                    if (g + PT_THREADS < total) fetch(g + PT_THREADS, b1);
                    process(b0);
                    g += PT_THREADS;
                    if (g >= total) break;
                    if (g + PT_THREADS < total) fetch(g + PT_THREADS, b0);
                    process(b1);
                    g += PT_THREADS;
                }
            }
            /* the columns before and after the whole words: one lane per segment, 4-byte windows of the read */
            for (uint32_t i = tid; i < ns; i += PT_THREADS) {
                const uint4 raw = s_seg[i];
                const uint32_t typ = raw.z & 3u, col = raw.w & 0xffffu, len = raw.w >> 16;
                if (len == 0) continue;
                const uint32_t end = col + len, wlo = (col + 3u) >> 2, whi = end >> 2;
                constexpr int YOFF = ROWS * PT_WORDS * 4;
                const uint32_t row_s = planes_s + (raw.z >> 8) * (PT_WORDS * 4); /* shared-space address of the row in plane X */
                const uint32_t h1 = (4u * wlo < end) ? 4u * wlo : end;            /* head columns [col, h1) */
                const uint32_t t0 = (whi >= wlo) ? 4u * whi : end;                /* tail columns [t0, end) */
                if (h1 == col && t0 == end) continue;
                const uint32_t ha = row_s + col, ta = row_s + t0;
                const bool h0p = col < h1, h1p = col + 1u < h1, h2p = col + 2u < h1, t0p = t0 < end, t1p = t0 + 1u < end, t2p = t0 + 2u < end;
                if (typ == SEG_M) {
                    const uint64_t spos = ((uint64_t)raw.y << 32) | raw.x;
                    const uint32_t fmask = (raw.z & 4u) ? 0x0f0f0f0fu : 0u, tsb = ((raw.z >> 3) & 3u) * 0x10101010u;
                    const uint64_t ah = spos, at = spos + (t0 - col);
                    const uint32_t *sh = reinterpret_cast<const uint32_t *>(seqp + (ah & ~(uint64_t)3)), *qh = reinterpret_cast<const uint32_t *>(qualp + (ah & ~(uint64_t)3));
                    const uint32_t *st = reinterpret_cast<const uint32_t *>(seqp + (at & ~(uint64_t)3)), *qt = reinterpret_cast<const uint32_t *>(qualp + (at & ~(uint64_t)3));
                    const uint32_t sh0 = __ldg(sh), sh1 = __ldg(sh + 1), qh0 = __ldg(qh), qh1 = __ldg(qh + 1);
                    const uint32_t st0 = __ldg(st), st1 = __ldg(st + 1), qt0 = __ldg(qt), qt1 = __ldg(qt + 1);
                    const uint32_t roth = 0x3210u + 0x1111u * ((uint32_t)ah & 3u), rott = 0x3210u + 0x1111u * ((uint32_t)at & 3u);
                    uint32_t xh, yh, xt, yt;
                    onehot4(__byte_perm(sh0, sh1, roth), __byte_perm(qh0, qh1, roth), minq4, pass_allow, fmask, tsb, xh, yh);
                    onehot4(__byte_perm(st0, st1, rott), __byte_perm(qt0, qt1, rott), minq4, pass_allow, fmask, tsb, xt, yt);
                    sts8_if<0>(ha, xh, h0p); sts8_if<1>(ha, xh >> 8, h1p); sts8_if<2>(ha, xh >> 16, h2p);
                    sts8_if<YOFF>(ha, yh, h0p); sts8_if<YOFF + 1>(ha, yh >> 8, h1p); sts8_if<YOFF + 2>(ha, yh >> 16, h2p);
                    sts8_if<0>(ta, xt, t0p); sts8_if<1>(ta, xt >> 8, t1p); sts8_if<2>(ta, xt >> 16, t2p);
                    sts8_if<YOFF>(ta, yt, t0p); sts8_if<YOFF + 1>(ta, yt >> 8, t1p); sts8_if<YOFF + 2>(ta, yt >> 16, t2p);
                } else {
                    const uint32_t v = typ == SEG_D ? 0x40u : 0x80u;
                    sts8_if<YOFF>(ha, v, h0p); sts8_if<YOFF + 1>(ha, v, h1p); sts8_if<YOFF + 2>(ha, v, h2p);
                    sts8_if<YOFF>(ta, v, t0p); sts8_if<YOFF + 1>(ta, v, t1p); sts8_if<YOFF + 2>(ta, v, t2p);
                }
            }
        }
        __syncthreads();
        /* column sums of this batch: Harley-Seal blocks of 16 rows (rows up to the next multiple of 16 are zero) */
        {
            const uint32_t *pl = planes + my_plane * (ROWS * PT_WORDS) + my_word;
            for (uint32_t r0 = 0; r0 < nrow; r0 += 16) {
                uint32_t w[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) w[j] = pl[(r0 + j) * PT_WORDS];
                uint32_t t2a, t2b, t4a, t4b, t8a, t8b, t16;
                CSA(t2a, ones, ones, w[0], w[1]);
                CSA(t2b, ones, ones, w[2], w[3]);
                CSA(t4a, twos, twos, t2a, t2b);
                CSA(t2a, ones, ones, w[4], w[5]);
                CSA(t2b, ones, ones, w[6], w[7]);
                CSA(t4b, twos, twos, t2a, t2b);
                CSA(t8a, fours, fours, t4a, t4b);
                CSA(t2a, ones, ones, w[8], w[9]);
                CSA(t2b, ones, ones, w[10], w[11]);
                CSA(t4a, twos, twos, t2a, t2b);
                CSA(t2a, ones, ones, w[12], w[13]);
                CSA(t2b, ones, ones, w[14], w[15]);
                CSA(t4b, twos, twos, t2a, t2b);
                CSA(t8b, fours, fours, t4a, t4b);
                CSA(t16, eights, eights, t8a, t8b);
                /* ripple the sixteens into the upper bit slices */
                uint32_t cy = t16, tt;
                tt = s4 & cy; s4 ^= cy; cy = tt;
                tt = s5 & cy; s5 ^= cy; cy = tt;
                tt = s6 & cy; s6 ^= cy; cy = tt;
                s7 ^= cy;
            }
            acc_rows += nrow;
        }
    }
    flush();
    /* counter exchange: 8-bit counters go to rows 0-7 of the thread's own plane, at its own word (only this thread ever
       reads that word during the column sums, so no barrier is needed before the write) */
    if (DEEP) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s_out32[(my_plane * 8 + i) * LCR_TILE + my_word * 4 + j] += (cnt8[i] >> (8 * j)) & 0xffu;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) planes[(my_plane * ROWS + i) * PT_WORDS + my_word] = cnt8[i];
    }
    __syncthreads();
    const uint8_t *ref = D.ref;
    for (uint32_t colr = tid; colr < npos; colr += PT_THREADS) {
        uint32_t v[16];
        if (DEEP) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = s_out32[i * LCR_TILE + colr];
        } else {
            const uint8_t *o8 = reinterpret_cast<const uint8_t *>(planes);
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i] = o8[i * LCR_TILE + colr]; v[8 + i] = o8[(ROWS + i) * LCR_TILE + colr]; }
        }
        SiteCounters sc;
#pragma unroll
        for (int i = 0; i < 4; ++i) { sc.cnt[i] = v[i]; sc.pass[i] = v[4 + i]; sc.fwd[i] = v[8 + i]; }
        sc.ts[0] = v[12]; sc.ts[1] = v[13]; sc.d = v[14]; sc.n = v[15] + D.full_n;
        sc.ll0 = 0; sc.ll2 = 0; sc.q0flags = 0;
        if (a.pl_acgt) {
            const uint64_t g = D.pos_g + colr;
#pragma unroll
            for (int i = 0; i < 4; ++i) { a.pl_acgt[g * 4 + i] = sc.cnt[i]; a.pl_fwd[g * 4 + i] = sc.fwd[i]; }
            a.pl_d[g] = sc.d; a.pl_n[g] = sc.n; a.pl_ts[g * 2] = sc.ts[0]; a.pl_ts[g * 2 + 1] = sc.ts[1];
        }
        lcr_candidate dummy;
        if (site_call<true>(a.P, *a.tables, sc, ref[colr], dummy)) {
            const uint32_t k = atomicAdd(a.pre_count, 1u);
            if (k < a.pre_cap) {
                PreCand pc;
                pc.tile = tile; pc.col = colr;
#pragma unroll
                for (int i = 0; i < 4; ++i) { pc.cnt[i] = sc.cnt[i]; pc.pass[i] = sc.pass[i]; pc.fwd[i] = sc.fwd[i]; }
                pc.ts[0] = sc.ts[0]; pc.ts[1] = sc.ts[1]; pc.d = sc.d; pc.n = sc.n;
                a.pre[k] = pc;
            }
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
    /* one lane per read of the tile: its segments are in column order and hold exactly the unmasked aligned bases */
    const uint32_t colr = pc.col;
    for (uint32_t idx = a.tile_off[tile] + lane; idx < a.tile_off[tile + 1]; idx += 32) {
        const uint2 is = a.item_segs[idx];
        for (uint32_t k = 0; k < is.y; ++k) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(a.segs + is.x + k));
            const uint32_t scol = raw.w & 0xffffu, slen = raw.w >> 16;
            if (scol > colr) break;
            if (colr >= scol + slen) continue;
            if ((raw.z & 3u) == SEG_M) {
                const uint64_t sp = (((uint64_t)raw.y << 32) | raw.x) + (colr - scol);
                const uint8_t b = a.seq[sp];
                const uint32_t rq = a.qual[sp];
                const uint32_t q = rq < LCR_MAX_BASE_QUALITY ? rq : LCR_MAX_BASE_QUALITY;
                const int bc = base_code_dev(b);
                if (bc >= 0) {
                    const bool is_ref = bc == refc;
                    const long long E = a.tables->gl_fx_err[q], K = a.tables->gl_fx_ok[q];
                    ll0 += is_ref ? E : K;
                    ll2 += is_ref ? K : E;
                    if (q == 0) q0flags |= is_ref ? 1u : 2u;
                }
            }
            break;
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

/* per region: candidate range in the sorted array */
__global__ void k_cand_ranges(uint32_t n_regions, const uint64_t *keys, uint32_t n_cand, LcrRegionState *rstate) {
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
}

/* the dense-cluster filters (candidate.rs:465-526), one thread per window start i.  The windows of different starts only
   set the same two flag bits and never read them, so they are independent of each other and of the order of the two passes. */
__global__ void k_cand_dense(lcr_params P, const uint64_t *keys, uint32_t n_cand, lcr_candidate *cand, const LcrRegionState *rstate) {
    const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= n_cand) return;
    const uint32_t reg = (uint32_t)(keys[gi] >> 32);
    const LcrRegionState rs = rstate[reg];
    if (gi < rs.cand_begin || gi >= rs.cand_begin + rs.n_cand) return;
    lcr_candidate *c = cand + rs.cand_begin;
    const uint32_t n = rs.n_cand, i = gi - rs.cand_begin;
    /* concat_idxes = homo_snps + het_snps, sorted: the candidates carrying HOM_VAR or HET_VAR */
    if (!(c[i].flags & (LCR_CF_HOM_VAR | LCR_CF_HET_VAR))) return;
    const int64_t start_pos = c[i].pos;
    for (int pass = 0; pass < 2; ++pass) {
        const int64_t win = pass == 0 ? (int64_t)P.dense_win_size : 5;
        const uint32_t min_cnt = pass == 0 ? P.min_dense_cnt : 3u;
        uint32_t cnt_between = 0; /* j - i in concat_idxes terms */
        uint32_t last_member = i, mark_end = i;
        bool broke = false;
        for (uint32_t j = i; j < n; ++j) {
            if (!(c[j].flags & (LCR_CF_HOM_VAR | LCR_CF_HET_VAR))) continue;
            const int64_t diff = c[j].pos - start_pos;
            const bool over = pass == 0 ? diff > win : diff >= win;
            if (over) {
                if (cnt_between >= min_cnt) mark_end = j;
                broke = true;
                break;
            }
            last_member = j;
            cnt_between++;
        }
        /* reached the last element inside the window: (j - i + 1) >= min_cnt marks i..j exclusive */
        if (!broke && cnt_between >= min_cnt) mark_end = last_member;
        for (uint32_t tk = i; tk < mark_end; ++tk)
            if (c[tk].flags & (LCR_CF_HOM_VAR | LCR_CF_HET_VAR)) c[tk].flags = (uint16_t)((c[tk].flags | LCR_CF_DENSE) & ~LCR_CF_FOR_PHASING);
    }
}

__global__ void k_count_pass(const uint8_t *slot_flags, uint32_t n_slots, lcr_stats *stats) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = (i < n_slots && (slot_flags[i] & 1)) ? 1u : 0u;
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
    /* per-tile counters: [0] items, [1] whole-tile intron covers (+ the deep-tile flag); per-slot segment bounds; their scans */
    uint32_t *tile_cnt = nullptr, *tile_off = nullptr, *slot_segs = nullptr, *slot_seg_off = nullptr;
    uint64_t *slot_runs = nullptr;
    uint2 *item_segs = nullptr;
    LcrSeg *segs = nullptr;
    const size_t tn = (size_t)n_tiles + 1, sn = (size_t)db->n_slots + 1;
    TRY(cudaMallocAsync(&tile_cnt, sizeof(uint32_t) * (2 * tn + 1), st));
    TRY(cudaMallocAsync(&tile_off, sizeof(uint32_t) * tn, st));
    TRY(cudaMallocAsync(&slot_segs, sizeof(uint32_t) * sn, st));
    TRY(cudaMallocAsync(&slot_seg_off, sizeof(uint32_t) * sn, st));
    TRY(cudaMallocAsync(&slot_runs, sizeof(uint64_t) * LCR_SLOT_RUNS * (size_t)(db->n_slots ? db->n_slots : 1), st));
    TRY(cudaMemsetAsync(tile_cnt, 0, sizeof(uint32_t) * (2 * tn + 1), st));
    TRY(cudaMemsetAsync(slot_segs, 0, sizeof(uint32_t) * sn, st));
    uint32_t *tile_count = tile_cnt, *tile_full_n = tile_cnt + tn, *deep_flag = tile_cnt + 2 * tn;

    PrepArgs pa{};
    pa.P = ctx->P;
    pa.n_slots = db->n_slots;
    pa.regions = db->regions;
    pa.slot_off = db->slot_off; pa.slot_region = db->slot_region; pa.tile_base = db->tile_base;
    pa.pos = db->pos; pa.flag = db->flag; pa.mapq = db->mapq; pa.ts = db->ts; pa.de = db->de;
    pa.seq_off = db->seq_off; pa.cig_off = db->cig_off; pa.seq = db->seq; pa.cigar = db->cigar;
    pa.ref_table = ctx->d_ref_table;
    pa.rstate = db->rstate; pa.stats = db->d_stats;
    pa.slot_flags = slot_flags; pa.slot_runs = slot_runs;
    pa.tile_count = tile_count; pa.tile_off = tile_off; pa.tile_full_n = tile_full_n; pa.deep_flag = deep_flag;
    pa.slot_segs = slot_segs; pa.slot_seg_off = slot_seg_off;
    pa.item_segs = nullptr; pa.segs = nullptr;
    /* batches averaging more than 24 CIGAR ops per read (ONT) take the warp-per-read, lane-per-op form (LCR_PREP_WALK: 1 / 2 forces one) */
    static const int prep_mode = [] { const char *e = getenv("LCR_PREP_WALK"); return e && *e ? atoi(e) : 0; }();
    const bool warp_walk = prep_mode == 2 || (prep_mode == 0 && db->n_cigar > 24ull * (db->n_reads ? db->n_reads : 1));
    const uint32_t pb = 128, pg = warp_walk ? (uint32_t)(((uint64_t)db->n_slots * 32 + pb - 1) / pb) : (db->n_slots + pb - 1) / pb;
    if (pg) {
        if (warp_walk) k_slot_prep_w<false><<<pg, pb, 0, st>>>(pa);
        else k_slot_prep<false><<<pg, pb, 0, st>>>(pa);
        k_count_pass<<<(db->n_slots + pb - 1) / pb, pb, 0, st>>>(slot_flags, db->n_slots, db->d_stats);
        db->timing.kernel_launches += 2;
    }
    /* exclusive scans of the per-tile item counts and the per-read segment bounds */
    size_t tmp_bytes = 0, tmp_bytes2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, tile_count, tile_off, n_tiles + 1, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes2, slot_segs, slot_seg_off, db->n_slots + 1, st);
    if (tmp_bytes2 > tmp_bytes) tmp_bytes = tmp_bytes2;
    void *tmp = nullptr;
    TRY(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, st));
    TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, tile_count, tile_off, n_tiles + 1, st));
    TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, slot_segs, slot_seg_off, db->n_slots + 1, st));
    uint32_t totals[3] = {0, 0, 0};
    TRY(cudaMemcpyAsync(&totals[0], tile_off + n_tiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(&totals[1], slot_seg_off + db->n_slots, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(&totals[2], deep_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
    const uint32_t n_items = totals[0], n_segs = totals[1];
    TRY(cudaMallocAsync(&item_segs, sizeof(uint2) * (size_t)(n_items ? n_items : 1), st));
    TRY(cudaMallocAsync(&segs, sizeof(LcrSeg) * (size_t)(n_segs ? n_segs : 1), st));
    TRY(cudaMemsetAsync(tile_count, 0, sizeof(uint32_t) * tn, st));
    TRY(cudaMemsetAsync(item_segs, 0, sizeof(uint2) * (size_t)(n_items ? n_items : 1), st));
    pa.item_segs = item_segs; pa.segs = segs;
    if (pg) {
        if (warp_walk) k_slot_prep_w<true><<<pg, pb, 0, st>>>(pa);
        else k_slot_prep<true><<<pg, pb, 0, st>>>(pa);
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
    LcrTileDesc *desc = nullptr;
    TRY(cudaMallocAsync(&desc, sizeof(LcrTileDesc) * tn, st));
    if (n_tiles) {
        DescArgs da{};
        da.n_tiles = n_tiles; da.regions = db->regions;
        da.tile_base = db->tile_base; da.tile_region = db->tile_region; da.tile_off = tile_off; da.tile_full_n = tile_full_n;
        da.pos_off = db->pos_off; da.ref_table = ctx->d_ref_table; da.rstate = db->rstate; da.desc = desc;
        k_tile_desc<<<(n_tiles + 255) / 256, 256, 0, st>>>(da);
        db->timing.kernel_launches += 1;
    }
    /* launch shape of the tile kernel: rows staged per batch / resident CTAs per SM (LCR_TILE_VARIANT: experiments) */
    static const int variant = [] { const char *e = getenv("LCR_TILE_VARIANT"); return e && *e ? atoi(e) : 0; }();
    /* default: 48 rows, 4 CTAs / SM (64 registers, 56 KB), 256 staged segments; 1: 32 rows; 2: 48 rows, 3 CTAs / SM, 512 segments */
    void (*k_tile)(PileArgs) = variant == 1 ? k_pileup_tile<false, 32, 4, 512> : variant == 2 ? k_pileup_tile<false, 48, 3, 512> : k_pileup_tile<false, 48, 4, 256>;
    void (*k_tile_deep)(PileArgs) = variant == 1 ? k_pileup_tile<true, 32, 4, 512> : variant == 2 ? k_pileup_tile<true, 48, 3, 512> : k_pileup_tile<true, 48, 4, 256>;
    const size_t tile_smem = variant == 1 ? pt_smem_bytes<32, 512>(false) : variant == 2 ? pt_smem_bytes<48, 512>(false) : pt_smem_bytes<48, 256>(false);
    const size_t tile_smem_deep = variant == 1 ? pt_smem_bytes<32, 512>(true) : variant == 2 ? pt_smem_bytes<48, 512>(true) : pt_smem_bytes<48, 256>(true);
    TRY(cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
    TRY(cudaFuncSetAttribute(k_tile_deep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem_deep));
    PileArgs ka{};
    ka.P = ctx->P;
    ka.regions = db->regions;
    ka.slot_off = db->slot_off; ka.slot_region = db->slot_region; ka.tile_base = db->tile_base; ka.tile_region = db->tile_region;
    ka.flag = db->flag; ka.ts = db->ts; ka.seq_off = db->seq_off; ka.cig_off = db->cig_off;
    ka.seq = db->seq; ka.qual = db->qual; ka.cigar = db->cigar;
    ka.ref_table = ctx->d_ref_table;
    ka.tile_off = tile_off;
    ka.desc = desc; ka.item_segs = item_segs; ka.segs = segs;
    ka.tables = ctx->d_tables;
    ka.rstate = db->rstate;
    ka.stats = db->d_stats;
    ka.pl_acgt = db->pl_acgt; ka.pl_fwd = db->pl_fwd; ka.pl_d = db->pl_d; ka.pl_n = db->pl_n; ka.pl_ts = db->pl_ts;
    ka.pre_count = counters; ka.cand_count = counters + 1;
    uint32_t n_pre = 0, n_cand = 0;
    /* tiles with more than 255 items take the variant with 32-bit column counters */
    const bool any_deep = totals[2] != 0;
    if (db->seq_wait_pending) { /* asynchronous upload: seq / qual are first read here */
        TRY(cudaStreamWaitEvent(st, db->ev_seq, 0));
        db->seq_wait_pending = false;
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        TRY(cudaMallocAsync(&pre, sizeof(PreCand) * (size_t)pre_cap, st));
        TRY(cudaMemsetAsync(counters, 0, 2 * sizeof(uint32_t), st));
        ka.pre = pre; ka.pre_cap = pre_cap;
        TRY(cudaEventRecord(ev0, st));
        if (n_tiles) {
            k_tile<<<n_tiles, PT_THREADS, tile_smem, st>>>(ka);
            db->timing.kernel_launches += 1;
            if (any_deep) {
                k_tile_deep<<<n_tiles, PT_THREADS, tile_smem_deep, st>>>(ka);
                db->timing.kernel_launches += 1;
            }
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
        db->timing.pileup_alg_bytes = 2ull * hs.n_aligned_bases + 4ull * db->n_cigar + 32ull * db->n_reads + db->n_pos + sizeof(PreCand) * (uint64_t)n_pre;
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
        k_cand_ranges<<<(db->n_regions + 63) / 64, 64, 0, st>>>(db->n_regions, keys_sorted, n_cand, db->rstate);
        db->timing.kernel_launches += 1;
        if (n_cand) {
            k_cand_dense<<<(n_cand + 63) / 64, 64, 0, st>>>(ctx->P, keys_sorted, n_cand, db->cand, db->rstate);
            db->timing.kernel_launches += 1;
        }
    }
    TRY(cudaFreeAsync(keys_sorted, st));
    TRY(cudaFreeAsync(cand_raw, st));
    TRY(cudaFreeAsync(cand_key, st));
    TRY(cudaFreeAsync(cand_count, st));
    TRY(cudaFreeAsync(segs, st));
    TRY(cudaFreeAsync(desc, st));
    TRY(cudaFreeAsync(slot_runs, st));
    TRY(cudaFreeAsync(tmp, st));
    TRY(cudaFreeAsync(tile_cnt, st));
    TRY(cudaFreeAsync(tile_off, st));
    TRY(cudaFreeAsync(slot_segs, st));
    TRY(cudaFreeAsync(slot_seg_off, st));
    TRY(cudaFreeAsync(item_segs, st));
    TRY(cudaGetLastError());
    return LCR_OK;
}
