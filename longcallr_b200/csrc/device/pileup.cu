/*
 * pileup.cu — read filter, tile work items, fused pileup + site genotyping, candidate lists.
 *
 * Replaces, on the device:
 *   src/util.rs:636-668      read filter and fetch window             (k_read_span)
 *   src/util.rs:650-948      Profile::fill_data_into_freq_vec         (k_read_walk: CIGAR walk + masks -> segments; k_pileup_tile: counts)
 *   src/util.rs:162-176      BaseFreq::get_two_major_alleles          (site_call)
 *   src/candidate.rs:75-463  filter cascade, genotype likelihood      (site_call<true> in k_pileup_tile, k_site_ll + site_call<false>)
 *   src/candidate.rs:465-526 dense-cluster filters                    (k_cand_ranges, k_cand_dense)
 *
 * Layout: the reads of a batch are decomposed on the device into (read, tile) items and, per item, segments (runs of unmasked
 * aligned bases / deleted / intron positions on consecutive columns of one tile of LCR_TILE reference positions).  The tile
 * kernel is persistent and warp-specialised: producer warps stream the rows of a tile into shared memory with asynchronous
 * copies (cp.async completing on mbarriers, a bulk copy for the tile's reference bytes), consumer warps count by difference
 * from the reference (range updates for coverage, a byte-wise XOR against the reference for the bases, shared-memory atomics
 * only for the ~1 % of bytes that differ).  No per-position record goes to HBM: only the sites that pass the count-based
 * filters (and, on request, the debug planes) are written.  The whole stage is enqueued without a host round trip: sizes
 * live in the device counter block (lcr_pipeline.h) and capacities are checked there.
 */
#include <cub/cub.cuh>

#include <algorithm>

#include "lcr_device.h"
#include "lcr_async.h"

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
    /* candidate.rs:177-194 needs base qualities: the tile kernel (PRE) never reads them, k_site_ll counts them at the surviving sites
       (every test of the cascade only rejects, so the order of the tests does not change the outcome) */
    if (!PRE) {
        if (allele1 != ref_base) { if (allele1_cnt > 0 && pick(s.pass, i1) < 2) return false; }
        else if (allele2 != ref_base) { if (allele2_cnt > 0 && pick(s.pass, i2) < 2) return false; }
    } else { /* fewer than two bases of that allele cannot hold two passing ones */
        if (allele1 != ref_base) { if (allele1_cnt == 1) return false; }
        else if (allele2 != ref_base) { if (allele2_cnt == 1) return false; }
    }
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

/* import_external_candidates (candidate.rs:530-613) for one listed position: alleles, frequencies and depth from the pileup counters,
   QUAL and genotype class from the record; 0/0 and unknown classes produce no candidate, a negative QUAL is dropped (min_variant_qual = 0.0) */
__device__ bool site_import(const uint32_t (&cnt)[4], uint8_t ref_base, uint32_t gt, float quality, lcr_candidate &o) {
    if (quality < 0.0f) return false;
    if (gt < 1u || gt > 3u) return false;
    const uint32_t total = cnt[0] + cnt[1] + cnt[2] + cnt[3];
    /* get_two_major_alleles (util.rs:162-176), as in site_call */
    uint32_t k0 = (cnt[0] << 2) | 3u, k1 = (cnt[1] << 2) | 2u, k2 = (cnt[2] << 2) | 1u, k3 = cnt[3] << 2;
#define LCR_CX(x, y) do { const uint32_t hi__ = max(x, y), lo__ = min(x, y); x = hi__; y = lo__; } while (0)
    LCR_CX(k0, k1); LCR_CX(k2, k3); LCR_CX(k0, k2); LCR_CX(k1, k3); LCR_CX(k1, k2);
#undef LCR_CX
    const int ord[4] = {3 - (int)(k0 & 3u), 3 - (int)(k1 & 3u), 3 - (int)(k2 & 3u), 3 - (int)(k3 & 3u)};
    const uint32_t oc[4] = {k0 >> 2, k1 >> 2, k2 >> 2, k3 >> 2};
    auto letter = [](int c) -> uint8_t { return (uint8_t)(0x54474341u >> (8 * c)); };
    int i2 = ord[1];
    uint32_t c2 = oc[1];
    if (letter(ord[0]) != ref_base && letter(ord[1]) != ref_base) {
        if (oc[2] == oc[1] && letter(ord[2]) == ref_base) { i2 = ord[2]; c2 = oc[2]; }
        else if (oc[3] == oc[1] && letter(ord[3]) == ref_base) { i2 = ord[3]; c2 = oc[3]; }
    }
    o.variant_quality = (double)quality;
    o.genotype_quality = (double)quality;
    o.phase_score = 0.0;
    o.genotype_probability[0] = 0.0; o.genotype_probability[1] = 0.0; o.genotype_probability[2] = 0.0;
    o.allele_freqs[0] = (float)oc[0] / (float)total; o.allele_freqs[1] = (float)c2 / (float)total; /* 0 / 0 = NaN on an uncovered position, as in the reference */
    o.depth = total;
    o.phase_set = 0;
    o.reference = ref_base;
    o.alleles[0] = letter(ord[0]); o.alleles[1] = letter(i2);
    o.haplotype = 0;
    o.reserved = 0;
    if (gt == 1u) { o.variant_type = 1; o.genotype = 0; o.flags = LCR_CF_HET_VAR | LCR_CF_FOR_PHASING; }
    else if (gt == 2u) { o.variant_type = 2; o.genotype = -1; o.flags = LCR_CF_HOM_VAR | LCR_CF_FOR_PHASING; }
    else { o.variant_type = 3; o.genotype = -1; o.flags = LCR_CF_HOM_VAR; }
    return true;
}

struct LcrSeg {             /* 16 B: a run of unmasked aligned bases, deleted or intron positions of one read inside one tile */
    uint64_t spos;          /* M: offset of the first base in the seq / qual pools */
    uint32_t typ;           /* bits 0-1 type, bit 2 forward strand, bits 3-4 transcript strand code */
    uint16_t col;           /* first column inside the tile */
    uint16_t len;           /* 1 .. LCR_TILE */
};
#define SEG_M 0u
#define SEG_D 1u
#define SEG_N 2u
#define LCR_SLOT_RUNS 4     /* homopolymer runs remembered per read for the poly-A mask */

struct LcrItem {            /* 16 B: the part of one read inside one tile = one row of the tile's pileup */
    uint64_t spos0_nseg;    /* bits 0-47 pool offset of the item's first aligned base, bits 48-63 number of segments */
    uint32_t seg0;          /* first segment */
    uint32_t span;          /* pool bytes from spos0 to the end of the item's last aligned base (inserted bases included); 0: no aligned base */
};

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
    LcrCounters *ctr;
    uint8_t *slot_flags;   /* bit 0: passes the read filter and overlaps the window; bit 1: has homopolymer runs near a read end;
                              bit 2: more than LCR_SLOT_RUNS of them (every zone base takes the exact test) */
    uint64_t *slot_runs;   /* [n_slots][LCR_SLOT_RUNS]: start << 32 | length << 8 | letter, in read coordinates */
    uint32_t *tile_diff;   /* span pass: +1 at the first tile a read can touch, -1 after the last: its scan bounds the items of a tile */
    const uint32_t *tile_off; /* walk: first item slot of every tile */
    uint32_t *tile_cursor; /* walk: items registered so far */
    uint32_t *tile_full_n; /* whole-tile intron covers */
    uint32_t *slot_seg_ub; /* span pass: upper bound of the read's segments */
    const uint32_t *slot_seg_off; /* its exclusive scan: the read's segments are written there, in walk order, no atomics */
    LcrItem *items;
    LcrSeg *segs;
    uint64_t items_cap, segs_cap;
};

/* Pass 1, one thread per (region, read) slot: the read filter (util.rs:652-668), the fetch window, the homopolymer-run
   table of the read ends (only bases inside / next to such a run can be poly-A masked, util.rs:754-789) and two upper
   bounds the walk needs before it starts: the tiles the read can register in (difference array over its reference span)
   and the number of segments it can emit.  Reads the CIGAR once for its reference length; no per-tile work. */
__global__ void __launch_bounds__(128) k_read_span(PrepArgs a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t pass_cnt = 0;
    if (slot < a.n_slots) {
        const uint32_t reg = a.slot_region[slot];
        uint8_t sflags = 0;
        uint32_t seg_ub = 0;
        if (a.rstate[reg].status == 0) {
            const lcr_region R = a.regions[reg];
            const uint32_t read = R.read_begin + (slot - a.slot_off[reg]);
            const uint64_t c0 = a.cig_off[read], c1 = a.cig_off[read + 1];
            const uint64_t s0 = a.seq_off[read];
            const int64_t seq_len = (int64_t)(a.seq_off[read + 1] - s0);
            const uint8_t *seq = a.seq + s0;
            /* util.rs:652-668 */
            const uint16_t fl = a.flag[read];
            bool pass = !((int32_t)a.mapq[read] < a.P.min_mapq || (uint64_t)seq_len < (uint64_t)a.P.min_read_length || (fl & 0x4) || (fl & 0x100) || (fl & 0x800));
            const float de = a.de[read];
            if (!(de != de) && de >= a.P.divergence) pass = false;
            /* fetch((chr, start, end)): pos < end && bam_endpos > start on the region's own numbers */
            int64_t rlen = 0;
            for (uint64_t c = c0; c < c1; ++c) {
                const uint32_t op = __ldg(a.cigar + c);
                if (is_ref_consuming(op & 0xf)) rlen += op >> 4;
            }
            const int64_t p = a.pos[read];
            const bool in_window = p < (int64_t)R.end && p + (rlen ? rlen : 1) > (int64_t)R.start;
            if (pass && in_window) {
                sflags = 1;
                pass_cnt = 1;
                const int64_t lead = lcr_leading_softclips(a.cigar, c0, c1), trail = lcr_trailing_softclips(a.cigar, c0, c1);
                const int64_t rb = seq_len - trail;
                const int64_t dend = (int64_t)(a.P.distance_to_read_end > 0x3fffffffu ? 0x3fffffffu : a.P.distance_to_read_end);
                const bool ont = a.P.platform == 1;
                uint32_t mask_ub = 0;
                if (dend > 0 && ont) mask_ub = 2; /* each read-end zone removes one stretch */
                if (!ont && dend > 0) {
                    const int64_t polya = (int64_t)(a.P.polya_tail_length > 0x3fffffffu ? 0x3fffffffu : a.P.polya_tail_length);
                    const int64_t zone_ub = 2 * (2 * dend - 1);
                    if (polya < 2 || rb < lead) { sflags |= 6; mask_ub = (uint32_t)(zone_ub > 0x7fffffff ? 0x7fffffff : zone_ub); } /* degenerate window or overlapping clips: exact test on every zone base */
                    else {
                        uint64_t runs[LCR_SLOT_RUNS];
#pragma unroll
                        for (int i = 0; i < LCR_SLOT_RUNS; ++i) runs[i] = 0;
                        const int64_t centre[2] = {lead, rb};
                        uint32_t nrun = 0;
                        int64_t run_ub = 0;
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
                                    run_ub += run + 2;
                                }
                                run = letter ? 1 : 0;
                                prev = b;
                            }
                            if (hi > scanned_to) scanned_to = hi;
                        }
                        if (nrun) sflags |= 2;
                        if (nrun > LCR_SLOT_RUNS) { sflags |= 4; run_ub = zone_ub; }
                        if (run_ub > zone_ub) run_ub = zone_ub;
                        mask_ub = (uint32_t)run_ub;
                        if (sflags & 2) {
#pragma unroll
                            for (int i = 0; i < LCR_SLOT_RUNS; ++i) a.slot_runs[(size_t)slot * LCR_SLOT_RUNS + i] = runs[i];
                        }
                    }
                }
                /* tiles the read can touch inside the region, and the segments it can emit: one per (op, tile) piece plus mask cuts */
                const int64_t vec_size = (int64_t)R.end - (int64_t)R.start;
                const int64_t f0 = p - ((int64_t)R.start - 1);
                const int64_t a0 = f0 < 0 ? 0 : f0, b0 = f0 + rlen < vec_size ? f0 + rlen : vec_size;
                if (b0 > a0) {
                    const uint32_t t0 = (uint32_t)(a0 / LCR_TILE), t1 = (uint32_t)((b0 - 1) / LCR_TILE);
                    const uint32_t tb = a.tile_base[reg];
                    atomicAdd(&a.tile_diff[tb + t0], 1u);
                    atomicAdd(&a.tile_diff[tb + t1 + 1], 0xffffffffu);
                    seg_ub = (uint32_t)(c1 - c0) + (t1 - t0) + mask_ub;
                }
            }
        }
        a.slot_flags[slot] = sflags;
        a.slot_seg_ub[slot] = seg_ub;
    }
    pass_cnt = __reduce_add_sync(0xffffffffu, pass_cnt);
    if ((threadIdx.x & 31) == 0 && pass_cnt) atomicAdd((unsigned long long *)&a.stats->n_reads_pass, (unsigned long long)pass_cnt);
}

/* tile_diff -> per-tile item upper bounds, in place (a running sum over all tiles of the batch: every read's +1 / -1
   pair lies inside its own region's tile range, so the sum is zero again at every region boundary) */
__global__ void k_fix_totals(const uint32_t *tile_off, uint32_t n_tiles, const uint32_t *slot_seg_off, uint32_t n_slots, uint64_t items_cap, uint64_t segs_cap, LcrCounters *ctr) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const uint32_t ni = tile_off[n_tiles], ns = slot_seg_off[n_slots];
        ctr->n_items = ni;
        ctr->n_segs = ns;
        if (ni > items_cap) atomicOr(&ctr->overflow, LCR_OVF_ITEMS);
        if (ns > segs_cap) atomicOr(&ctr->overflow, LCR_OVF_SEGS);
    }
}

/* Pass 2, one thread per slot: the CIGAR walk of util.rs:692-947 as a decomposition of the read into per-tile items
   (rows of the tile's pileup) and segments.  The read-end trim (ONT) and the poly-A / homopolymer mask (util.rs:737-789)
   are applied here by cutting aligned runs at masked bases.  Item and segment slots come from the upper bounds of pass 1,
   so this is the only full walk. */
__global__ void __launch_bounds__(128, 6) k_read_walk(PrepArgs a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n_bases = 0;
    bool live = slot < a.n_slots;
    uint8_t sflags = 0;
    if (live) { sflags = a.slot_flags[slot]; live = sflags & 1; }
    uint32_t reg = 0;
    if (live) {
        reg = a.slot_region[slot];
        if (a.rstate[reg].status != 0) live = false;
    }
    if (live && (a.ctr->overflow & (LCR_OVF_ITEMS | LCR_OVF_SEGS))) live = false; /* the run is repeated with larger buffers */
    if (live) {
        const lcr_region R = a.regions[reg];
        const uint32_t read = R.read_begin + (slot - a.slot_off[reg]);
        const uint64_t c0 = a.cig_off[read], c1 = a.cig_off[read + 1];
        const uint64_t s0 = a.seq_off[read];
        const int64_t seq_len = (int64_t)(a.seq_off[read + 1] - s0);
        const uint8_t *seq = a.seq + s0;
        const int64_t lead = lcr_leading_softclips(a.cigar, c0, c1), trail = lcr_trailing_softclips(a.cigar, c0, c1);
        const int64_t rb = seq_len - trail;
        const int64_t dend = (int64_t)(a.P.distance_to_read_end > 0x3fffffffu ? 0x3fffffffu : a.P.distance_to_read_end);
        const bool ont = a.P.platform == 1;
        uint64_t runs[LCR_SLOT_RUNS];
#pragma unroll
        for (int i = 0; i < LCR_SLOT_RUNS; ++i) runs[i] = 0;
        if ((sflags & 6) == 2) {
#pragma unroll
            for (int i = 0; i < LCR_SLOT_RUNS; ++i) runs[i] = a.slot_runs[(size_t)slot * LCR_SLOT_RUNS + i];
        }
        const int64_t vec_size = (int64_t)R.end - (int64_t)R.start;
        const int64_t fv_start = (int64_t)R.start - 1;
        const uint32_t tb = a.tile_base[reg];
        const uint8_t *ref = a.ref_table[R.tid] + fv_start;
        uint32_t typ_c;
        {
            const int strand = (a.flag[read] & 0x10) ? 1 : 0;
            const int8_t ts = a.ts[read];
            uint32_t tcode = 0;
            if (ts == '+') tcode = strand == 0 ? 1u : 2u;
            else if (ts == '-') tcode = strand == 0 ? 2u : 1u;
            typ_c = (strand == 0 ? 4u : 0u) | (tcode << 3);
        }
        /* read-end zones in read coordinates, ordered by their first base */
        int64_t zlo[2] = {lead - dend + 1, rb - dend + 1}, zhi[2] = {lead + dend - 1, rb + dend - 1};
        if (zlo[1] < zlo[0]) { int64_t t = zlo[0]; zlo[0] = zlo[1]; zlo[1] = t; t = zhi[0]; zhi[0] = zhi[1]; zhi[1] = t; }
        const int mask_mode = dend <= 0 ? 0 : ont ? 1 : (sflags & 4) ? 2 : (sflags & 2) ? 3 : 0;

        int64_t fpos = (int64_t)a.pos[read] - fv_start;
        int64_t rpos = lead;
        int64_t last_tile = -1;
        uint32_t item_k = 0xffffffffu;               /* slot of the current item */
        uint32_t seg_w = a.slot_seg_off[slot];       /* next segment slot of this read */
        const uint32_t seg_end = a.slot_seg_off[slot + 1];
        uint32_t seg_item0 = seg_w;                  /* first segment of the current item */
        uint64_t m_first = ~0ull, m_end = 0;         /* pool range of the current item's aligned bases */
        bool bad = false, ovf = false;
        auto leave_tile = [&]() {
            if (item_k != 0xffffffffu) {
                LcrItem it;
                const uint32_t ns = seg_w - seg_item0;
                it.spos0_nseg = (m_first == ~0ull ? 0ull : m_first) | ((uint64_t)ns << 48);
                it.seg0 = seg_item0;
                it.span = m_first == ~0ull ? 0u : (uint32_t)(m_end - m_first);
                a.items[item_k] = it;
            }
            item_k = 0xffffffffu;
            seg_item0 = seg_w;
            m_first = ~0ull; m_end = 0;
        };
        for (uint64_t c = c0; c < c1 && !bad; ++c) {
            const uint32_t op = __ldg(a.cigar + c), opc = op & 0xf, len = op >> 4;
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
                        atomicAdd(&a.tile_full_n[tb + t], 1u);
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
                        if ((int)(threadIdx.x & 31) == leader) first = atomicAdd(&a.tile_cursor[key], npeer);
                        first = __shfl_sync(peers, first, leader);
                        item_k = a.tile_off[key] + first + prank;
                        if ((uint64_t)item_k >= a.items_cap || item_k >= a.tile_off[key + 1]) { ovf = true; item_k = 0xffffffffu; }
                    }
                    const uint32_t colr = (uint32_t)(ts - t * LCR_TILE);
                    auto put = [&](uint32_t typ, uint64_t spos, uint32_t col, uint32_t n) {
                        if (seg_w < seg_end && (uint64_t)seg_w < a.segs_cap) {
                            LcrSeg s;
                            s.spos = spos; s.typ = typ_c | typ; s.col = (uint16_t)col; s.len = (uint16_t)n;
                            a.segs[seg_w] = s;
                            if (typ == SEG_M) {
                                if (m_first == ~0ull) m_first = spos;
                                m_end = spos + n;
                            }
                            seg_w++;
                        } else ovf = true;
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
                    } else if (mask_mode == 2) { /* exact test on every zone base */
                        for (int k = 0; k < 2; ++k) {
                            const int64_t zl = zlo[k] > start ? zlo[k] : start, zh = zhi[k] < pb - 1 ? zhi[k] : pb - 1;
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
        if (ovf) atomicOr(&a.ctr->overflow, LCR_OVF_SEGS | LCR_OVF_ITEMS);
        if (seg_w > a.slot_seg_off[slot]) atomicAdd(&a.ctr->n_segs_used, seg_w - a.slot_seg_off[slot]);
    }
    n_bases = __reduce_add_sync(0xffffffffu, n_bases);
    if ((threadIdx.x & 31) == 0 && n_bases) atomicAdd((unsigned long long *)&a.stats->n_aligned_bases, (unsigned long long)n_bases);
}

/* ------------------------------------------------------------------------- *
 * Tile pileup: persistent, warp-specialised, asynchronous copies, reference-differential.
 *
 * k_pileup_tile runs CTAs of PT_CONS consumer threads + PT_PROD_WARPS producer warps (4 CTAs per SM by default) over a
 * device-built list of tiles (dynamic tickets).
 *
 * Producers turn the items (rows) of a tile into batches of up to 32 rows: producer warp 0 copies the contiguous seq bytes
 * every row covers, 16 bytes per lane with cp.async.ca.shared.global, producer warp 1 the rows' segment descriptors, and both
 * tie their copies to the stage's "full" mbarrier with cp.async.mbarrier.arrive.noinc; the tile's reference bytes come with
 * the first batch as one 1-D bulk copy (cp.async.bulk.shared::cluster.global, byte count expected on the same mbarrier).
 * Consumers release a stage through its "empty" mbarrier; no consumer ever waits on a global load of its own.
 *
 * Consumers count by difference from the reference.  Almost every aligned base equals the reference base, so a base
 * is not expanded into counters at all:
 *   - coverage of every (strand, transcript-strand) class, deletions and introns are range updates: +1 / -1 on a
 *     per-class difference array at the ends of each segment, prefix-summed once per tile;
 *   - one lane takes a 16-column block of an aligned segment and XORs the 16 read bytes with the 16 compare bytes of its
 *     columns (the reference byte, or 0x80 where the reference is no A/C/G/T) with word-wide logic; blocks holding a byte
 *     that differs are appended (with a 16-bit mask of those bytes) to a short list;
 *   - a second pass takes the listed bytes (mismatches, lower-case and non-ACGT bytes: ~1 % of the bases) to
 *     per-column event counters with shared-memory atomics (two 16-bit counters per word; tiles of more than 65535 rows
 *     take the BIG instantiation with 32-bit counters).
 * Base qualities are not read here: the only count filter that needs them (candidate.rs:177-194) is evaluated by
 * k_site_ll, which reads the qualities of the surviving sites anyway.
 * At the end of the tile the per-column counters of util.rs:100-127 follow exactly from coverage minus events
 * (all integers); columns with any mismatch that pass a cheap depth / fraction prefilter are compacted, the count-based
 * site filters (site_call<true>) run on those, and the surviving sites are appended in column order to the tile's range
 * of the pre-candidate list (slots reserved PT_PRE_CHUNK at a time; unused slots are marked and skipped downstream).
 * ------------------------------------------------------------------------- */
#ifdef LCR_TILE_PROF
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(slot, t0, t1) do { if (lane == 0 && (warp == 0 || warp == PT_CONS / 32)) atomicAdd(&a.ctr->prof[slot], (unsigned long long)((t1) - (t0))); } while (0)
#else
#define PROF_T(var)
#define PROF_ADD(slot, t0, t1)
#endif
#ifndef PT_CONS
#define PT_CONS 256                       /* consumer threads */
#endif
#define PT_PROD_WARPS 2                   /* producer warps: one per copied stream (seq bytes; segment descriptors + reference + row tables) */
#define PT_THREADS (PT_CONS + 32 * PT_PROD_WARPS)
#define PT_ROW_BYTES_MAX 2048u            /* largest item span staged as one row (larger ones are cut into pieces by the producer) */
#define PT_FLAG_FIRST 1u
#define PT_FLAG_LAST 2u
#define PT_FLAG_QUIT 4u
#define PT_PRE_CHUNK 64u                  /* pre-candidate slots a CTA reserves at a time */

struct PreCand { /* a site that passed every count-based filter; its likelihood is computed by k_site_ll */
    uint32_t tile, col;
    uint32_t cnt[4], pass[4], fwd[4], ts[2], d, n;
};

struct __align__(16) LcrTileDesc { /* 48 B, one per tile (k_tile_desc) */
    const uint8_t *ref;     /* reference base of column 0 */
    uint64_t pos_g;         /* index of column 0 in the debug planes */
    uint32_t it0, n_items;  /* items of the tile */
    uint32_t pos1;          /* 1-based reference position of column 0 */
    uint32_t reserved;
    uint32_t reg, npos;
    uint32_t full_n;        /* introns covering the whole tile */
    int32_t status;         /* of the region */
};

struct DescArgs {
    uint32_t n_tiles;
    const lcr_region *regions;
    const uint32_t *tile_base, *tile_region, *tile_off, *tile_cursor, *tile_full_n;
    const uint64_t *pos_off;
    const uint8_t *const *ref_table;
    const LcrRegionState *rstate;
    LcrTileDesc *desc;
    uint32_t *list[2];      /* work lists: tiles of at most / more than 255 items */
    LcrCounters *ctr;
    int all_tiles;          /* debug planes requested or candidates imported: empty tiles are processed too */
    uint32_t big_rows;      /* tiles with more rows than this take the 32-bit event counters */
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
    d.it0 = a.tile_off[tile]; d.n_items = a.tile_cursor[tile];
    d.pos1 = (uint32_t)((int64_t)R.start + tile_start); d.reserved = 0;
    d.reg = reg; d.npos = (uint32_t)(tile_end - tile_start);
    d.full_n = a.tile_full_n[tile];
    a.desc[tile] = d;
    {
        uint32_t ni = d.status == 0 ? d.n_items : 0u;
        ni = __reduce_add_sync(__activemask(), ni);
        if ((threadIdx.x & 31u) == (uint32_t)(__ffs(__activemask()) - 1) && ni) atomicAdd(&a.ctr->n_items_used, ni);
    }
    /* work lists of the tile kernel (tiles of more than 65535 rows take its 32-bit flavour): one atomic per warp */
    const bool want = d.status == 0 && (d.n_items || a.all_tiles);
    if (want && d.n_items > a.big_rows) a.list[1][atomicAdd(&a.ctr->n_list[1], 1u)] = tile;
    const bool want0 = want && d.n_items <= a.big_rows;
    const unsigned m = __ballot_sync(__activemask(), want0);
    if (want0) {
        const unsigned lane = threadIdx.x & 31u;
        const int leader = __ffs(m) - 1;
        uint32_t base = 0;
        if ((int)lane == leader) base = atomicAdd(&a.ctr->n_list[0], (uint32_t)__popc(m));
        base = __shfl_sync(m, base, leader);
        a.list[0][base + __popc(m & ((1u << lane) - 1u))] = tile;
    }
}

struct PileArgs {
    lcr_params P;
    const lcr_region *regions;
    const uint32_t *tile_base, *tile_region;
    const uint8_t *seq, *qual;
    const uint8_t *const *ref_table;
    const LcrTileDesc *desc;
    const uint32_t *tile_list;   /* work list of this launch */
    int list_id;                 /* 0 shallow, 1 deep: which scheduler of the counter block */
    const LcrItem *items;
    const LcrSeg *segs;
    const LcrDeviceTables *tables;
    LcrRegionState *rstate;
    LcrCounters *ctr;
    lcr_stats *stats;
    uint32_t *pl_acgt, *pl_fwd, *pl_d, *pl_n, *pl_ts; /* debug planes or null */
    const uint32_t *exon_off;    /* --exon-only: per region the sorted union of its exon intervals, or null */
    const uint2 *exon_iv;
    const uint32_t *ext_off, *ext_pos; /* -v: imported candidate positions per region (ascending), or null */
    const uint8_t *ext_gt;
    const float *ext_qual;
    PreCand *pre;
    uint32_t pre_cap;
    uint2 *tile_pre;             /* per tile: first pre-candidate and count */
    /* k_site_ll */
    lcr_candidate *cand_raw;
    uint8_t *cand_keep;
};

__device__ __forceinline__ void cons_bar() { asm volatile("bar.sync 1, %0;" ::"n"(PT_CONS) : "memory"); }
__device__ __forceinline__ void prod_bar() { asm volatile("bar.sync 2, %0;" ::"n"(32 * PT_PROD_WARPS) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts32a(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8a(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

#define PT_NDIFF 8                        /* difference arrays: coverage of the 6 (strand, transcript strand) classes, deletions, introns */
#define PT_DIFF_LEN (LCR_TILE + 4)        /* one entry past the last column, padded to whole uint4 */
#define PT_NEV 10                         /* event counters per column */
#define EV_MIS 0                          /* [4] bases of letter L that differ from the reference base */
#define EV_MISFWD 4                       /* [4] ... on forward reads */
#define EV_NONACGT 8                      /* aligned bytes that are no A/C/G/T letter */
#define EV_NONACGT_FWD 9                  /* ... on forward reads */

template <bool BIG, int STAGES, int MINB = 4>
struct PtLayout { /* dynamic shared memory of k_pileup_tile, byte offsets */
    /* three CTAs per SM leave room for batches that hold a whole 30x tile (48 rows, 16 KB of bases); four CTAs per SM take 32 rows / 12 KB */
    static constexpr bool WIDE = STAGES == 1 && MINB <= 3;
    static constexpr uint32_t ROWS = WIDE ? 48u : 32u;                       /* rows per batch */
    static constexpr uint32_t STAGE_BYTES = WIDE ? 16000u : (STAGES == 1 ? 12288u : 8192u); /* seq bytes per stage */
    static constexpr uint32_t SEG_CAP = WIDE ? 320u : 256u;                  /* segments per batch */
    static constexpr uint32_t ROW_SEG_MAX = 128u;                            /* segments of one row */
    static constexpr uint32_t BLK_CAP = ROWS * (LCR_TILE / 16) + SEG_CAP;    /* 16-column blocks per batch */
    static constexpr uint32_t PAD = 32u;                                     /* slack before / after the staged bytes (block loads start up to 15 B early, read 20 B) */
    static constexpr uint32_t diff = 0;                                      /* int32 [PT_NDIFF][PT_DIFF_LEN] */
    static constexpr uint32_t ev = diff + PT_NDIFF * PT_DIFF_LEN * 4u;       /* event counters: uint32 [PT_NEV][LCR_TILE] (BIG) or 16-bit pairs */
    static constexpr uint32_t cnt_end = ev + PT_NEV * LCR_TILE * (BIG ? 4u : 2u); /* everything before is cleared between tiles */
    static constexpr uint32_t ref = cnt_end;                                 /* [LCR_TILE] compare byte of every column */
    static constexpr uint32_t stage0 = ref + LCR_TILE;
    /* one stage */
    static constexpr uint32_t st_seq = 0;
    static constexpr uint32_t st_segs = st_seq + PAD + STAGE_BYTES + PAD;
    static constexpr uint32_t st_ref = st_segs + SEG_CAP * 16u;              /* the tile's reference bytes from the 16-byte boundary below column 0 */
    static constexpr uint32_t st_segrow = st_ref + LCR_TILE + 32u;           /* [SEG_CAP] row of every staged segment */
    static constexpr uint32_t st_rowseg = st_segrow + SEG_CAP;               /* [ROWS + 1] first staged segment of every row */
    static constexpr uint32_t st_rowdelta = st_rowseg + (ROWS + 1u) * 4u;    /* [ROWS] staged byte offset minus pool offset (mod 2^32) */
    static constexpr uint32_t st_hdr = (st_rowdelta + ROWS * 4u + 15u) & ~15u; /* tile, rows, segments, flags */
    static constexpr uint32_t stage_size = (st_hdr + 16u + 127u) & ~127u;
    static constexpr uint32_t blk = stage0 + STAGES * stage_size;            /* [BLK_CAP] u32 per 16-column block: staged byte of its first column | block << 15 | first byte << 20 | last byte << 24 | forward << 28 */
    static constexpr uint32_t evl = blk + BLK_CAP * 4u;                      /* [BLK_CAP] u32: blocks with exceptional bytes: block list index | byte mask << 16 */
    static constexpr uint32_t bytes() { return evl + BLK_CAP * 4u; }
};

/* BIG: tiles of more than 65535 rows (32-bit event counters); the others pack two 16-bit event counters per word */
template <bool BIG, int PT_STAGES, int MINB>
__global__ void __launch_bounds__(PT_THREADS, MINB) k_pileup_tile(PileArgs a) {
    using L = PtLayout<BIG, PT_STAGES, MINB>;
    constexpr uint32_t ROWS = L::ROWS;
    static_assert(ROWS <= 64 && L::SEG_CAP <= 1024, "row ids take 6 bits, segment ids 10");
    static_assert(LCR_TILE % PT_CONS == 0, "whole columns per consumer thread");
    static_assert(L::PAD + L::STAGE_BYTES + 16 <= (1u << 15) && LCR_TILE / 16 <= 32, "block entries: 15 bits of staged offset, 5 bits of block index");
    static_assert(L::BLK_CAP <= 65536, "block list indices take 16 bits");
    static_assert(PT_ROW_BYTES_MAX + 32 <= L::STAGE_BYTES && L::ROW_SEG_MAX <= L::SEG_CAP, "one row always fits an empty batch");
    extern __shared__ __align__(128) unsigned char pt_smem[];
    __shared__ __align__(8) unsigned long long s_bar[2 * PT_STAGES]; /* full[stage], empty[stage] */
    __shared__ uint32_t s_pcnt[2];
    __shared__ uint32_t s_ticket[2];
    __shared__ uint32_t s_nev, s_nblk, s_nsurv;
    __shared__ uint32_t s_bitmap[LCR_TILE / 32];  /* columns that passed the count filters */
    __shared__ uint32_t s_chunk[2];               /* next free slot and slots left of the CTA's chunk of the pre-candidate list */
    __shared__ uint16_t s_surv[LCR_TILE];         /* columns that need the full cascade */

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t smem0 = smem_u32(pt_smem);
    const uint32_t bar_full = smem_u32(&s_bar[0]), bar_empty = smem_u32(&s_bar[PT_STAGES]);
    if (tid < LCR_TILE / 32) s_bitmap[tid] = 0;
    if (tid == 0) {
        s_nblk = 0; s_nsurv = 0; s_chunk[0] = 0; s_chunk[1] = 0;
        /* a batch is full when every producer lane's asynchronous copies have landed (one deferred arrival per lane), the header is
           written and the bulk copy of the reference window has delivered its bytes (one arrival with the expected byte count) */
        for (int s = 0; s < PT_STAGES; ++s) { mbar_init(bar_full + 8 * s, 32 * PT_PROD_WARPS + 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_fence_init();
    }
    { /* the counters start zeroed; every tile's epilogue clears them again */
        uint4 *p4 = reinterpret_cast<uint4 *>(pt_smem + L::diff);
        for (uint32_t i = tid; i < L::cnt_end / 16u; i += PT_THREADS) p4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    if (warp >= PT_CONS / 32) {
        /* ================= producer warps =================
           The two warps run the same control flow over the same tiles and batches.  Rows are a few hundred bytes each, too small
           for one bulk copy apiece (a bulk copy takes uniform operands, so a warp issues its lanes' copies one after the other):
           every lane streams its own row with 16-byte asynchronous copies (LDGSTS), 32 rows in flight per instruction.
           Warp 0 copies the seq bytes; warp 1 the segment descriptors, writes the row tables and the batch header and, with the
           first batch of a tile, fetches the tile's reference window with one bulk copy. */
        const uint32_t role = warp - PT_CONS / 32;
        uint32_t stage = 0, empty_par = (1u << PT_STAGES) - 1u; /* bit s: parity to wait for; a fresh barrier passes a wait on the opposite parity, so both stages start empty */
        uint32_t rows = 0, bytes = 0, nsegs = 0, tx = 0, cur_tile = 0, flags = PT_FLAG_FIRST;
        bool open = false;
        auto stage_base = [&](uint32_t s) { return smem0 + L::stage0 + s * L::stage_size; };
        uint32_t tile_seq = 0;
        auto open_batch = [&]() { /* all lanes; lane 0 waits for the consumers to release the stage */
            PROF_T(pp0);
            if (lane == 0) mbar_wait_relaxed(bar_empty + 8 * stage, (empty_par >> stage) & 1u);
            empty_par ^= 1u << stage;
            __syncwarp();
            PROF_T(pp1);
            PROF_ADD(8, pp0, pp1);
            rows = 0; bytes = 0; nsegs = 0; tx = 0;
            open = true;
        };
        auto close_batch = [&](uint32_t fl) { /* all lanes: tie the copies to the barrier, publish the header */
            cp_async_arrive(bar_full + 8 * stage);
            if (role == 1) {
                __threadfence_block();
                __syncwarp();
                if (lane == 0) {
                    const uint32_t sb = stage_base(stage);
                    sts32a(sb + L::st_rowseg + rows * 4u, nsegs);
                    sts128(sb + L::st_hdr, cur_tile, rows, nsegs, fl);
                    mbar_arrive_expect_tx(bar_full + 8 * stage, tx);
                }
            }
            __syncwarp();
            stage = (stage + 1) % PT_STAGES;
            open = false;
        };
        /* one row by one lane: [src_lo, src_lo + nbytes) of the seq pool, segments [seg_first, seg_first + ns) */
        auto issue_row = [&](uint32_t row, uint32_t byte_off, uint32_t seg_off, uint64_t src_lo, uint32_t nbytes, uint32_t seg_first, uint32_t ns) {
            const uint32_t sb = stage_base(stage);
            if (role == 0) {
                const uint32_t dst = sb + L::st_seq + L::PAD + byte_off;
                const uint8_t *src = a.seq + src_lo;
                for (uint32_t o = 0; o < nbytes; o += 16) cp_async16(dst + o, src + o);
            } else {
                const uint32_t dst = sb + L::st_segs + seg_off * 16u;
                const LcrSeg *src = a.segs + seg_first;
                for (uint32_t o = 0; o < ns; ++o) {
                    cp_async16(dst + o * 16u, src + o);
                    sts8a(sb + L::st_segrow + seg_off + o, row);
                }
                sts32a(sb + L::st_rowseg + row * 4u, seg_off);
                sts32a(sb + L::st_rowdelta + row * 4u, (L::PAD + byte_off) - (uint32_t)src_lo);
            }
        };
        /* the same by the whole warp (rows cut from oversized items) */
        auto issue_row_warp = [&](uint32_t row, uint32_t byte_off, uint32_t seg_off, uint64_t src_lo, uint32_t nbytes, uint32_t seg_first, uint32_t ns) {
            const uint32_t sb = stage_base(stage);
            if (role == 0) {
                for (uint32_t o = lane * 16u; o < nbytes; o += 512u) cp_async16(sb + L::st_seq + L::PAD + byte_off + o, a.seq + src_lo + o);
            } else {
                for (uint32_t o = lane; o < ns; o += 32u) {
                    cp_async16(sb + L::st_segs + (seg_off + o) * 16u, a.segs + seg_first + o);
                    sts8a(sb + L::st_segrow + seg_off + o, row);
                }
                if (lane == 0) {
                    sts32a(sb + L::st_rowseg + row * 4u, seg_off);
                    sts32a(sb + L::st_rowdelta + row * 4u, (L::PAD + byte_off) - (uint32_t)src_lo);
                }
            }
        };
        for (;;) {
            /* warp 0 draws the next tile, the other reads it (slots alternate, so one barrier per tile is enough) */
            PROF_T(pq0);
            if (role == 0 && lane == 0) s_ticket[tile_seq & 1u] = atomicAdd(&a.ctr->ticket[a.list_id], 1u);
            prod_bar();
            const uint32_t t = s_ticket[tile_seq & 1u];
            ++tile_seq;
            if (t >= a.ctr->n_list[a.list_id]) break;
            cur_tile = a.tile_list[t];
            const LcrTileDesc *dp = a.desc + cur_tile;
            const uint32_t it0 = dp->it0, n_items = dp->n_items;
            flags = PT_FLAG_FIRST;
            uint4 nxt = make_uint4(0, 0, 0, 0); /* the items of a round are loaded one round ahead */
            if (lane < n_items) nxt = __ldg(reinterpret_cast<const uint4 *>(a.items + it0 + lane));
#ifdef LCR_TILE_PROF
            if (nxt.x == 0xdeadbeefu) flags |= 8u; /* make the timer below wait for the load */
#endif
            PROF_T(pq1);
            PROF_ADD(9, pq0, pq1);
            /* the tile's reference bytes travel with its first batch */
            if (!open) open_batch();
            if (role == 1) {
                const uint8_t *rp = dp->ref;
                const uint32_t roff = (uint32_t)((uintptr_t)rp & 15u), rbytes = (roff + dp->npos + 15u) & ~15u;
                if (lane == 0) bulk_g2s(stage_base(stage) + L::st_ref, rp - roff, rbytes, bar_full + 8 * stage);
                tx += rbytes;
            }
            for (uint32_t done = 0; done < n_items; done += 32) {
                const uint32_t it = done + lane;
                const bool valid = it < n_items;
                const uint4 raw = nxt;
                if (it + 32 < n_items) nxt = __ldg(reinterpret_cast<const uint4 *>(a.items + it0 + it + 32));
                uint64_t spos0 = 0;
                uint32_t ns = 0, seg0 = 0, span = 0;
                if (valid) {
                    spos0 = (((uint64_t)raw.y << 32) | raw.x) & 0xffffffffffffull;
                    ns = raw.y >> 16; seg0 = raw.z; span = raw.w;
                }
                const uint64_t src_lo = spos0 & ~(uint64_t)15;
                uint32_t nb = span ? (uint32_t)(((spos0 + span + 15) & ~(uint64_t)15) - src_lo) : 0u;
                const bool big = valid && (ns > L::ROW_SEG_MAX || nb > PT_ROW_BYTES_MAX);
                if (!__any_sync(0xffffffffu, big)) {
                    /* fast path: the lanes' rows are appended to the open batch while they fit */
                    uint32_t pb = nb, ps = ns; /* inclusive prefix sums over the lanes */
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t vb = __shfl_up_sync(0xffffffffu, pb, o), vs = __shfl_up_sync(0xffffffffu, ps, o);
                        if ((int)lane >= o) { pb += vb; ps += vs; }
                    }
                    uint32_t first = 0;                   /* first lane not yet placed */
                    uint32_t base_b = 0, base_s = 0;      /* prefix sums up to that lane (exclusive) */
                    const uint32_t nvalid = __popc(__ballot_sync(0xffffffffu, valid));
                    while (first < nvalid) {
                        if (!open) open_batch();
                        const bool pend = valid && lane >= first;
                        const bool fits = pend && rows + (lane - first) < (uint32_t)ROWS && bytes + (pb - base_b) <= L::STAGE_BYTES && nsegs + (ps - base_s) <= L::SEG_CAP;
                        const uint32_t nfit = __popc(__ballot_sync(0xffffffffu, fits)); /* prefix sums are monotone: the fitting lanes are the first nfit pending ones */
                        if (nfit == 0) { close_batch(flags); flags = 0; continue; }
                        if (fits) issue_row(rows + (lane - first), bytes + (pb - base_b) - nb, nsegs + (ps - base_s) - ns, src_lo, nb, seg0, ns);
                        __syncwarp();
                        const uint32_t lastl = first + nfit - 1;
                        const uint32_t eb = __shfl_sync(0xffffffffu, pb, lastl), es = __shfl_sync(0xffffffffu, ps, lastl);
                        rows += nfit; bytes += eb - base_b; nsegs += es - base_s;
                        base_b = eb; base_s = es; first += nfit;
                    }
                } else {
                    /* rare: an item with hundreds of segments or a long insertion inside.  All lanes walk the round's items one by one with
                       the same control flow (the segment loads are warp-uniform); the item is cut into pieces that fit a row */
                    const uint32_t nvalid = __popc(__ballot_sync(0xffffffffu, valid));
                    for (uint32_t l = 0; l < nvalid; ++l) {
                        const uint32_t i_ns = __shfl_sync(0xffffffffu, ns, l), i_seg0 = __shfl_sync(0xffffffffu, seg0, l);
                        uint32_t s = 0;
                        do {
                            uint64_t lo_g = 0, hi_g = 0;
                            bool have_m = false;
                            uint32_t cnt = 0;
                            while (s + cnt < i_ns && cnt < L::ROW_SEG_MAX) {
                                const uint4 sg = __ldg(reinterpret_cast<const uint4 *>(a.segs + i_seg0 + s + cnt));
                                if ((sg.z & 3u) == SEG_M) {
                                    const uint64_t sp = ((uint64_t)sg.y << 32) | sg.x;
                                    const uint64_t nlo = have_m ? lo_g : (sp & ~(uint64_t)15), nhi = (sp + (sg.w >> 16) + 15) & ~(uint64_t)15;
                                    if (nhi - nlo > PT_ROW_BYTES_MAX && cnt) break;
                                    lo_g = nlo; hi_g = nhi; have_m = true;
                                }
                                ++cnt;
                            }
                            const uint32_t pbytes = have_m ? (uint32_t)(hi_g - lo_g) : 0u;
                            if (open && !(rows < (uint32_t)ROWS && bytes + pbytes <= L::STAGE_BYTES && nsegs + cnt <= L::SEG_CAP)) { close_batch(flags); flags = 0; }
                            if (!open) open_batch();
                            issue_row_warp(rows, bytes, nsegs, lo_g, pbytes, i_seg0 + s, cnt);
                            rows += 1; bytes += pbytes; nsegs += cnt;
                            s += cnt;
                        } while (s < i_ns);
                    }
                }
            }
            if (!open) open_batch();
            close_batch(flags | PT_FLAG_LAST);
        }
        /* tell the consumers to leave */
        open_batch();
        close_batch(PT_FLAG_QUIT);
        return;
    }

    /* ================= consumer warps ================= */
    int32_t *s_diff = reinterpret_cast<int32_t *>(pt_smem + L::diff);     /* [PT_NDIFF][PT_DIFF_LEN] */
    uint32_t *s_ev = reinterpret_cast<uint32_t *>(pt_smem + L::ev);       /* [PT_NEV][LCR_TILE] counters (BIG) or [PT_NEV][LCR_TILE / 2] pairs of 16-bit counters */
    uint8_t *s_ref = pt_smem + L::ref;                                    /* compare byte per column: the reference byte if it is an upper-case A/C/G/T, else 0x80 */
    uint32_t *s_blk = reinterpret_cast<uint32_t *>(pt_smem + L::blk);
    uint32_t *s_evl = reinterpret_cast<uint32_t *>(pt_smem + L::evl);
    auto ev_add = [&](uint32_t arr, uint32_t colx) {
        if (BIG) atomicAdd(&s_ev[arr * LCR_TILE + colx], 1u);
        else atomicAdd(&s_ev[arr * (LCR_TILE / 2) + (colx >> 1)], 1u << (16u * (colx & 1u))); /* at most 65535 rows per tile: no carry into the neighbour */
    };
    auto ev_get = [&](uint32_t arr, uint32_t colx) -> uint32_t {
        if (BIG) return s_ev[arr * LCR_TILE + colx];
        return (s_ev[arr * (LCR_TILE / 2) + (colx >> 1)] >> (16u * (colx & 1u))) & 0xffffu;
    };
    constexpr int NCW = PT_CONS / 32;          /* consumer warps */
    constexpr int CPT = LCR_TILE / PT_CONS;    /* columns per consumer thread in the epilogue */

    LcrTileDesc D;
    memset(&D, 0, sizeof D);
    uint32_t stage = 0, full_par = 0; /* bit s: parity to wait for */
    for (;;) {
        PROF_T(tp0);
        mbar_wait(bar_full + 8 * stage, (full_par >> stage) & 1u);
        full_par ^= 1u << stage;
        PROF_T(tp1);
        PROF_ADD(0, tp0, tp1);
        const unsigned char *stg = pt_smem + L::stage0 + stage * L::stage_size;
        const uint32_t stg_s = smem0 + L::stage0 + stage * L::stage_size;
        const uint4 hdr = *reinterpret_cast<const uint4 *>(stg + L::st_hdr);
        const uint32_t tile = hdr.x, nrow = hdr.y, nseg = hdr.z, bflags = hdr.w;
        if (bflags & PT_FLAG_QUIT) {
            for (uint32_t k = tid; k < s_chunk[1]; k += PT_CONS)
                if (s_chunk[0] + k < a.pre_cap) a.pre[s_chunk[0] + k].tile = 0xffffffffu; /* unused tail of the last chunk */
            break;
        }
        if (bflags & PT_FLAG_FIRST) {
            {
                const uint4 *dp = reinterpret_cast<const uint4 *>(a.desc + tile);
                uint4 *dd = reinterpret_cast<uint4 *>(&D);
                dd[0] = __ldg(dp); dd[1] = __ldg(dp + 1); dd[2] = __ldg(dp + 2);
            }
            /* column-aligned compare bytes from the staged reference window */
            const uint32_t roff = (uint32_t)((uintptr_t)D.ref & 15u);
#pragma unroll
            for (int h = 0; h < CPT; ++h) {
                const uint32_t colr = tid + h * PT_CONS;
                uint8_t b = 0x80;
                if (colr < D.npos) {
                    b = stg[L::st_ref + roff + colr];
                    if (!(b == 'A' || b == 'C' || b == 'G' || b == 'T')) b = 0x80;
                }
                s_ref[colr] = b;
            }
        }
        /* the batch's segments: range updates of the coverage / deletion / intron arrays, and the 16-column blocks of the aligned ones
           (listed in any order: a warp reserves its lanes' entries with one atomic) */
        const uint4 *s_seg = reinterpret_cast<const uint4 *>(stg + L::st_segs);
        const uint8_t *s_segrow = stg + L::st_segrow;
        const uint32_t *s_rowdelta = reinterpret_cast<const uint32_t *>(stg + L::st_rowdelta);
        if (tid == 0) s_nev = 0;
        for (uint32_t i0 = 0; i0 < nseg; i0 += PT_CONS) {
            const uint32_t i = i0 + tid;
            uint32_t n = 0, ent0 = 0, lo0 = 0, hi1 = 0;
            if (i < nseg) {
                const uint4 sg = s_seg[i];
                const uint32_t col = sg.w & 0xffffu, len = sg.w >> 16, typ = sg.z & 3u;
                if (len) {
                    uint32_t arr;
                    if (typ == SEG_M) {
                        n = ((col + len + 15u) >> 4) - (col >> 4);
                        arr = ((sg.z & 4u) ? 3u : 0u) + ((sg.z >> 3) & 3u); /* class: forward strand x transcript-strand code */
                        /* staged byte of the first column of the segment's first block (mod 2^32 arithmetic on the pool offset) */
                        const uint32_t src0 = s_rowdelta[s_segrow[i]] + sg.x - (col & 15u);
                        ent0 = src0 | ((col >> 4) << 15) | ((sg.z & 4u) << 26);
                        lo0 = col & 15u; hi1 = (col + len - 1u) & 15u;
                    } else arr = typ == SEG_D ? 6u : 7u;
                    atomicAdd(&s_diff[arr * PT_DIFF_LEN + col], 1);
                    atomicAdd(&s_diff[arr * PT_DIFF_LEN + col + len], -1);
                }
            }
            uint32_t incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += v;
            }
            uint32_t base = 0;
            const uint32_t wtot = __shfl_sync(0xffffffffu, incl, 31);
            if (wtot) {
                if (lane == 31) base = atomicAdd(&s_nblk, wtot);
                base = __shfl_sync(0xffffffffu, base, 31);
                const uint32_t run = base + incl - n;
                for (uint32_t k = 0; k < n; ++k) /* next block: 16 staged bytes and one block further; only the first / last block are partial */
                    s_blk[run + k] = (ent0 + k * (16u + (1u << 15))) | ((k == 0 ? lo0 : 0u) << 20) | ((k + 1 == n ? hi1 : 15u) << 24);
            }
        }
        cons_bar(); /* also: the compare bytes are in place */
        PROF_T(tp2);
        PROF_ADD(1, tp1, tp2);
        const uint32_t total = s_nblk;
        /* aligned bases, pass 1: one lane per 16-column block, read bytes from the stage against the compare bytes of the columns;
           blocks with a differing byte go to the event list with a mask of those bytes */
        const uint32_t seq_s = stg_s + L::st_seq;
        auto detect = [&](uint32_t g) -> uint32_t { /* bit i of the result: byte i of block g differs from its compare byte */
            const uint32_t e = s_blk[g];
            const uint32_t src = e & 0x7fffu, c0 = ((e >> 15) & 31u) << 4, lo = (e >> 20) & 15u, hi = ((e >> 24) & 15u) + 1u;
            const uint32_t al = src & ~3u, rot = 0x3210u + 0x1111u * (src & 3u);
            const uint32_t sa = seq_s + al;
            const uint32_t s0 = lds32(sa), s1 = lds32(sa + 4), s2 = lds32(sa + 8), s3 = lds32(sa + 12), s4w = lds32(sa + 16);
            const uint4 rf = *reinterpret_cast<const uint4 *>(s_ref + c0);
            const uint32_t d0 = __byte_perm(s0, s1, rot) ^ rf.x, d1 = __byte_perm(s1, s2, rot) ^ rf.y, d2 = __byte_perm(s2, s3, rot) ^ rf.z, d3 = __byte_perm(s3, s4w, rot) ^ rf.w;
            if ((d0 | d1 | d2 | d3) == 0u) return 0u;
            auto nzbits = [](uint32_t d) -> uint32_t {
                const uint32_t nz = (((d & 0x7f7f7f7fu) + 0x7f7f7f7fu) | d) & 0x80808080u; /* bit 7 of every non-zero byte */
                return ((nz >> 7) * 0x10204080u) >> 28;                                      /* gathered into 4 bits: byte i -> bit 28 + i */
            };
            const uint32_t m16 = nzbits(d0) | (nzbits(d1) << 4) | (nzbits(d2) << 8) | (nzbits(d3) << 12);
            return m16 & (0xffffu >> (16u - hi)) & (0xffffu << lo); /* only the bytes of the block that belong to the segment */
        };
        auto append = [&](uint32_t g, uint32_t m16) { /* all lanes of the warp */
            const unsigned vote = __ballot_sync(0xffffffffu, m16 != 0u);
            if (vote) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&s_nev, (uint32_t)__popc(vote));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (m16) s_evl[base + __popc(vote & ((1u << lane) - 1u))] = g | (m16 << 16);
            }
        };
        for (uint32_t g0 = 0; g0 < total; g0 += 2 * PT_CONS) { /* warp-uniform trip count; two independent blocks per lane in flight */
            const uint32_t ga = g0 + tid, gb = g0 + PT_CONS + tid;
            const uint32_t ma = ga < total ? detect(ga) : 0u;
            const uint32_t mb = gb < total ? detect(gb) : 0u;
            append(ga, ma);
            append(gb, mb);
        }
        cons_bar();
        PROF_T(tp3);
        PROF_ADD(2, tp2, tp3);
        if (tid == 0) s_nblk = 0; /* everybody has read it; the next batch's reservations come after the barriers below */
        /* pass 2: the exceptional bytes (mismatches, lower-case letters, non-ACGT bytes), one listed block per lane */
        {
            const uint32_t nev = s_nev;
            const unsigned char *seq_b = stg + L::st_seq;
            for (uint32_t x = tid; x < nev; x += PT_CONS) {
                const uint32_t ent = s_evl[x];
                const uint32_t e = s_blk[ent & 0xffffu];
                uint32_t m16 = ent >> 16;
                const uint32_t src = e & 0x7fffu, c0 = ((e >> 15) & 31u) << 4;
                const bool fwd = (e >> 28) & 1u;
                while (m16) {
                    const uint32_t t = (uint32_t)__ffs((int)m16) - 1u;
                    m16 &= m16 - 1u;
                    const uint32_t colx = c0 + t;
                    const uint32_t b = seq_b[src + t], rb = s_ref[colx];
                    const int lc = base_code_dev((uint8_t)b);
                    if (lc < 0) {
                        ev_add(EV_NONACGT, colx);
                        if (fwd) ev_add(EV_NONACGT_FWD, colx);
                    } else if ((b & 0xdfu) != rb) { /* not the reference letter in lower case */
                        ev_add(EV_MIS + lc, colx);
                        if (fwd) ev_add(EV_MISFWD + lc, colx);
                    }
                }
            }
        }
        cons_bar();
        PROF_T(tp4);
        PROF_ADD(3, tp3, tp4);
        if (tid == 0) mbar_arrive(bar_empty + 8 * stage); /* the stage's bytes and segments are consumed */
        stage = (stage + 1) % PT_STAGES;
        if (!(bflags & PT_FLAG_LAST)) continue;

        /* ---- end of the tile: coverage from the difference arrays, exact counters, count-based site filters, pre-candidates in column order ---- */
        for (int ai = warp; ai < PT_NDIFF; ai += NCW) { /* a warp turns a difference array into running sums: 16 consecutive entries per lane, then a warp scan of the lane totals */
            int32_t *arr = s_diff + ai * PT_DIFF_LEN + lane * 16;
            int4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = reinterpret_cast<const int4 *>(arr)[j];
            int32_t x[16] = {v[0].x, v[0].y, v[0].z, v[0].w, v[1].x, v[1].y, v[1].z, v[1].w, v[2].x, v[2].y, v[2].z, v[2].w, v[3].x, v[3].y, v[3].z, v[3].w};
#pragma unroll
            for (int j = 1; j < 16; ++j) x[j] += x[j - 1];
            int32_t tot = x[15], incl2 = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t u = __shfl_up_sync(0xffffffffu, incl2, o);
                if ((int)lane >= o) incl2 += u;
            }
            const int32_t base = incl2 - tot;
#pragma unroll
            for (int j = 0; j < 4; ++j) reinterpret_cast<int4 *>(arr)[j] = make_int4(x[4 * j] + base, x[4 * j + 1] + base, x[4 * j + 2] + base, x[4 * j + 3] + base);
        }
        cons_bar();
        PROF_T(tp5);
        PROF_ADD(4, tp4, tp5);
        const uint32_t npos = D.npos;
        auto load_site = [&](uint32_t colr, SiteCounters &sc, uint8_t &ref_byte) {
            uint32_t cov[6], evv[PT_NEV];
#pragma unroll
            for (int c = 0; c < 6; ++c) cov[c] = (uint32_t)s_diff[c * PT_DIFF_LEN + colr];
#pragma unroll
            for (int i = 0; i < PT_NEV; ++i) evv[i] = ev_get(i, colr);
            const uint8_t rb = s_ref[colr];
            const int refc = rb == 'A' ? 0 : rb == 'C' ? 1 : rb == 'G' ? 2 : rb == 'T' ? 3 : -1;
            const uint32_t cov_all = cov[0] + cov[1] + cov[2] + cov[3] + cov[4] + cov[5], cov_fwd = cov[3] + cov[4] + cov[5];
            const uint32_t acgt = cov_all - evv[EV_NONACGT], acgt_fwd = cov_fwd - evv[EV_NONACGT_FWD];
#pragma unroll
            for (int i = 0; i < 4; ++i) { sc.cnt[i] = evv[EV_MIS + i]; sc.pass[i] = 0; sc.fwd[i] = evv[EV_MISFWD + i]; } /* pass: counted by k_site_ll */
            if (refc >= 0) { /* the bases that are not events are the reference letter */
                const uint32_t mis = evv[EV_MIS] + evv[EV_MIS + 1] + evv[EV_MIS + 2] + evv[EV_MIS + 3];
                const uint32_t misf = evv[EV_MISFWD] + evv[EV_MISFWD + 1] + evv[EV_MISFWD + 2] + evv[EV_MISFWD + 3];
                const uint32_t rc = acgt - mis;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i == refc) { sc.cnt[i] = rc; sc.fwd[i] = acgt_fwd - misf; }
            }
            sc.ts[0] = cov[1] + cov[4]; sc.ts[1] = cov[2] + cov[5]; /* transcript-strand code 1 / 2 on either read strand (util.rs:803-819) */
            sc.d = (uint32_t)s_diff[6 * PT_DIFF_LEN + colr]; sc.n = (uint32_t)s_diff[7 * PT_DIFF_LEN + colr] + D.full_n;
            sc.ll0 = 0; sc.ll2 = 0; sc.q0flags = 0;
            ref_byte = rb == 0x80 ? (uint8_t)'N' : rb; /* any byte that is no upper-case A/C/G/T ends the cascade the same way (candidate.rs:132) */
        };
        /* (a) columns that can still be candidates: a base that differs from an upper-case A/C/G/T reference byte, enough depth,
           and not the everyday case of a stray mismatch under the low-fraction / low-count cut-off (candidate.rs:90-94,132,142-155,165) */
#pragma unroll
        for (int h = 0; h < CPT; ++h) {
            const uint32_t colr = tid + h * PT_CONS;
            bool maybe = colr < npos;
            if (maybe && a.pl_acgt) { /* debug planes: every column is written */
                SiteCounters sc;
                uint8_t rb;
                load_site(colr, sc, rb);
                const uint64_t g = D.pos_g + colr;
#pragma unroll
                for (int i = 0; i < 4; ++i) { a.pl_acgt[g * 4 + i] = sc.cnt[i]; a.pl_fwd[g * 4 + i] = sc.fwd[i]; }
                a.pl_d[g] = sc.d; a.pl_n[g] = sc.n; a.pl_ts[g * 2] = sc.ts[0]; a.pl_ts[g * 2 + 1] = sc.ts[1];
            }
            if (maybe && a.ext_off) { /* -v (candidate.rs:544): the columns the imported records name, whatever was piled on them */
                const uint32_t pos0 = D.pos1 - 1u + colr;
                uint32_t lo = a.ext_off[D.reg], hi = a.ext_off[D.reg + 1];
                const uint32_t end = hi;
                while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a.ext_pos[mid] < pos0) lo = mid + 1; else hi = mid; }
                maybe = lo < end && a.ext_pos[lo] == pos0;
            } else if (maybe) {
                const uint32_t m0 = ev_get(EV_MIS, colr), m1 = ev_get(EV_MIS + 1, colr), m2 = ev_get(EV_MIS + 2, colr), m3 = ev_get(EV_MIS + 3, colr);
                maybe = (m0 | m1 | m2 | m3) != 0u && s_ref[colr] != 0x80;
                if (maybe) {
                    uint32_t cov_all = 0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) cov_all += (uint32_t)s_diff[c * PT_DIFF_LEN + colr];
                    const uint32_t total = cov_all - ev_get(EV_NONACGT, colr);
                    const uint32_t rc = total - (m0 + m1 + m2 + m3), alt = max(max(m0, m1), max(m2, m3));
                    if (total < a.P.min_depth || total > a.P.max_depth) maybe = false;
                    else if (rc > alt) { /* the reference base strictly ahead: it is allele 1 and the largest other count the only alternative allele */
                        if (alt == 1u) maybe = false; /* one base cannot hold the two passing ones candidate.rs:177-194 asks for (site_call<PRE>) */
                        else if (total < 200u) { if ((float)alt / (float)total < a.P.low_allele_frac_cutoff) maybe = false; }
                        else if (alt < a.P.low_allele_cnt_cutoff) maybe = false;
                    }
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, maybe);
            if (m) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&s_nsurv, (uint32_t)__popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (maybe) s_surv[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)colr;
            }
        }
        cons_bar();
        PROF_T(tp6);
        PROF_ADD(5, tp5, tp6);
        /* (b) the full count-based cascade on the survivors, one per thread; the ones that pass set their bit in the column bitmap */
        const uint32_t nsurv = s_nsurv;
        for (uint32_t x = tid; x < nsurv; x += PT_CONS) {
            const uint32_t colr = s_surv[x];
            SiteCounters sc;
            uint8_t rb;
            load_site(colr, sc, rb);
            lcr_candidate dummy;
            bool ok = true;
            if (a.ext_off) { atomicOr(&s_bitmap[colr >> 5], 1u << (colr & 31u)); continue; } /* imported: no filter of the cascade applies */
            if (a.exon_off) { /* candidate.rs:80-89: only positions inside an exon of the region's genes are looked at */
                const uint32_t pos1 = D.pos1 + colr;
                uint32_t lo = a.exon_off[D.reg], hi = a.exon_off[D.reg + 1];
                const uint32_t first = lo;
                while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a.exon_iv[mid].x <= pos1) lo = mid + 1; else hi = mid; }
                ok = lo > first && pos1 < a.exon_iv[lo - 1].y;
            }
            if (ok && site_call<true>(a.P, *a.tables, sc, rb, dummy)) atomicOr(&s_bitmap[colr >> 5], 1u << (colr & 31u));
        }
        cons_bar();
        /* (c) a contiguous range of the pre-candidate list for the tile, sub-allocated from a chunk the CTA reserves with one global atomic */
        if (tid == 0) {
            uint32_t tot = 0;
            for (int i = 0; i < LCR_TILE / 32; ++i) tot += __popc(s_bitmap[i]);
            uint32_t base = 0;
            if (tot) {
                if (tot > s_chunk[1]) { /* the unused tail of the old chunk stays marked invalid (k_site_ll skips it) */
                    for (uint32_t k = 0; k < s_chunk[1]; ++k)
                        if (s_chunk[0] + k < a.pre_cap) a.pre[s_chunk[0] + k].tile = 0xffffffffu;
                    const uint32_t want = tot > PT_PRE_CHUNK ? tot : PT_PRE_CHUNK;
                    s_chunk[0] = atomicAdd(&a.ctr->n_pre, want);
                    s_chunk[1] = want;
                }
                base = s_chunk[0];
                s_chunk[0] += tot; s_chunk[1] -= tot;
            }
            s_pcnt[0] = base;
            a.tile_pre[tile] = make_uint2(base, tot);
            atomicAdd(&a.ctr->n_tiles_done, 1u);
            atomicAdd(&a.ctr->n_pos_done, (unsigned long long)npos);
        }
        cons_bar();
        {
            const uint32_t base = s_pcnt[0];
            for (uint32_t x = tid; x < nsurv; x += PT_CONS) {
                const uint32_t colr = s_surv[x];
                if (!((s_bitmap[colr >> 5] >> (colr & 31u)) & 1u)) continue;
                uint32_t rank = __popc(s_bitmap[colr >> 5] & ((1u << (colr & 31u)) - 1u)); /* passing columns before this one: column order */
                for (uint32_t w = 0; w < (colr >> 5); ++w) rank += __popc(s_bitmap[w]);
                const uint32_t k = base + rank;
                if (k < a.pre_cap) {
                    SiteCounters sc;
                    uint8_t rb;
                    load_site(colr, sc, rb);
                    PreCand pc;
                    pc.tile = tile; pc.col = colr;
#pragma unroll
                    for (int i = 0; i < 4; ++i) { pc.cnt[i] = sc.cnt[i]; pc.pass[i] = sc.pass[i]; pc.fwd[i] = sc.fwd[i]; }
                    pc.ts[0] = sc.ts[0]; pc.ts[1] = sc.ts[1]; pc.d = sc.d; pc.n = sc.n;
                    a.pre[k] = pc;
                }
            }
        }
        cons_bar(); /* the counters are read: clear them for the next tile */
        PROF_T(tp7);
        PROF_ADD(6, tp6, tp7);
        {
            uint4 *p4 = reinterpret_cast<uint4 *>(pt_smem + L::diff);
            for (uint32_t i = tid; i < L::cnt_end / 16u; i += PT_CONS) p4[i] = make_uint4(0, 0, 0, 0);
            if (tid < LCR_TILE / 32) s_bitmap[tid] = 0;
            if (tid == 0) s_nsurv = 0;
        }
        cons_bar();
        PROF_T(tp8);
        PROF_ADD(7, tp7, tp8);
    }
}

/* exact genotype likelihood of the sites that passed the count filters: one warp per site, lanes over the
   reads of the site's tile (candidate.rs:236-282 over the same unmasked bases the pileup counted).  The number of
   sites is a device counter: a fixed grid of warps strides over it. */
__global__ void __launch_bounds__(256) k_site_ll(PileArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n_pre = min(a.ctr->n_pre, a.pre_cap);
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_pre; w += nwarps) {
        const PreCand pc = a.pre[w];
        const uint32_t tile = pc.tile;
        if (tile == 0xffffffffu) { if (lane == 0) a.cand_keep[w] = 0; continue; } /* unused slot at the end of a CTA's chunk */
        const uint32_t reg = a.tile_region[tile];
        const lcr_region R = a.regions[reg];
        const int32_t tile_start = (int32_t)((tile - a.tile_base[reg]) * LCR_TILE);
        const int32_t col = tile_start + (int32_t)pc.col;
        const uint8_t ref_base = a.ref_table[R.tid][((int64_t)R.start - 1) + col];
        if (a.ext_off) { /* -v: the record of this position decides; no base or quality is read */
            if (lane == 0) {
                const uint32_t pos0 = (uint32_t)((int64_t)R.start - 1 + col);
                uint32_t lo = a.ext_off[reg], hi = a.ext_off[reg + 1];
                while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a.ext_pos[mid] < pos0) lo = mid + 1; else hi = mid; }
                lcr_candidate o;
                const bool keep = lo < a.ext_off[reg + 1] && a.ext_pos[lo] == pos0 && site_import(pc.cnt, ref_base, a.ext_gt[lo], a.ext_qual[lo], o);
                if (keep) {
                    o.pos = (int64_t)pos0;
                    o.region = reg;
                    a.cand_raw[w] = o;
                }
                a.cand_keep[w] = keep ? 1 : 0;
            }
            continue;
        }
        const int refc = (ref_base == 'A') ? 0 : (ref_base == 'C') ? 1 : (ref_base == 'G') ? 2 : (ref_base == 'T') ? 3 : 8;
        long long ll0 = 0, ll2 = 0;
        uint32_t q0flags = 0;
        uint32_t npass[4] = {0, 0, 0, 0}; /* bases of each letter with (capped) quality >= min_baseq (candidate.rs:177-194) */
        uint32_t nq = 0;
        /* one lane per read of the tile: its segments are in column order and hold exactly the unmasked aligned bases */
        const uint32_t colr = pc.col;
        const LcrTileDesc *dp = a.desc + tile;
        const uint32_t it0 = dp->it0, n_items = dp->n_items;
        for (uint32_t idx = lane; idx < n_items; idx += 32) {
            const uint4 it = __ldg(reinterpret_cast<const uint4 *>(a.items + it0 + idx));
            const uint32_t ns = it.y >> 16, seg0 = it.z;
            /* the row's segments are in column order and no two cover the same column: the one that can hold the site's column is
               the last that starts at or before it (binary search: ONT rows carry tens of segments) */
            uint32_t lo = 0, hi = ns;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                const uint32_t scol_mid = __ldg(reinterpret_cast<const uint32_t *>(a.segs + seg0 + mid) + 3) & 0xffffu;
                if (scol_mid <= colr) lo = mid + 1; else hi = mid;
            }
            for (uint32_t k = lo - 1; lo != 0;) {
                const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(a.segs + seg0 + k));
                const uint32_t scol = raw.w & 0xffffu, slen = raw.w >> 16;
                if (colr >= scol + slen) break;
                if ((raw.z & 3u) == SEG_M) {
                    const uint64_t sp = (((uint64_t)raw.y << 32) | raw.x) + (colr - scol);
                    const uint8_t b = a.seq[sp];
                    const int bc = base_code_dev(b);
                    if (bc >= 0) {
                        const uint32_t rq = a.qual[sp]; /* with LCR_FLAG_QUAL_ON_DEMAND this is a read of mapped host memory */
                        const uint32_t q = rq < LCR_MAX_BASE_QUALITY ? rq : LCR_MAX_BASE_QUALITY;
                        ++nq;
                        const bool is_ref = bc == refc;
                        const long long E = a.tables->gl_fx_err[q], K = a.tables->gl_fx_ok[q];
                        ll0 += is_ref ? E : K;
                        ll2 += is_ref ? K : E;
                        if (q == 0) q0flags |= is_ref ? 1u : 2u;
                        if ((int32_t)q >= a.P.min_baseq) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) npass[i] += bc == i ? 1u : 0u;
                        }
                    }
                }
                break;
            }
        }
        for (int o = 16; o; o >>= 1) {
            ll0 += __shfl_xor_sync(0xffffffffu, ll0, o);
            ll2 += __shfl_xor_sync(0xffffffffu, ll2, o);
            q0flags |= __shfl_xor_sync(0xffffffffu, q0flags, o);
            nq += __shfl_xor_sync(0xffffffffu, nq, o);
#pragma unroll
            for (int i = 0; i < 4; ++i) npass[i] += __shfl_xor_sync(0xffffffffu, npass[i], o);
        }
        if (lane != 0) continue;
        if (nq) atomicAdd(&a.ctr->qual_reads, (unsigned long long)nq);
        SiteCounters sc;
#pragma unroll
        for (int i = 0; i < 4; ++i) { sc.cnt[i] = pc.cnt[i]; sc.pass[i] = npass[i]; sc.fwd[i] = pc.fwd[i]; }
        sc.ts[0] = pc.ts[0]; sc.ts[1] = pc.ts[1]; sc.d = pc.d; sc.n = pc.n; sc.ll0 = ll0; sc.ll2 = ll2; sc.q0flags = q0flags;
        lcr_candidate o;
        const bool keep = site_call<false>(a.P, *a.tables, sc, ref_base, o);
        if (keep) {
            o.pos = (int64_t)R.start - 1 + col;
            o.region = reg;
            a.cand_raw[w] = o;
        }
        a.cand_keep[w] = keep ? 1 : 0;
    }
}

/* Candidates come out in (region, position) order without a sort: the pre-candidates of a tile are contiguous and in
   column order, tiles are numbered by (region, position), so counting the kept ones per tile, scanning over the tiles
   and copying tile by tile is a stable compaction. */
__global__ void k_tile_cand_count(uint32_t n_tiles, const uint2 *tile_pre, const uint8_t *keep, uint32_t pre_cap, uint32_t *tile_cand_cnt) {
    const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile > n_tiles) return;
    uint32_t c = 0;
    if (tile < n_tiles) {
        const uint2 r = tile_pre[tile];
        for (uint32_t k = r.x; k < r.x + r.y && k < pre_cap; ++k) c += keep[k];
    }
    tile_cand_cnt[tile] = c;
}

__global__ void k_tile_cand_gather(uint32_t n_tiles, const uint2 *tile_pre, const uint8_t *keep, uint32_t pre_cap, const uint32_t *tile_cand_off, const lcr_candidate *raw, lcr_candidate *out,
                                   LcrCounters *ctr) {
    const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= n_tiles) return;
    if (tile == 0) {
        ctr->n_cand = tile_cand_off[n_tiles];
        if (ctr->n_pre > pre_cap) atomicOr(&ctr->overflow, LCR_OVF_PRE);
    }
    const uint2 r = tile_pre[tile];
    uint32_t w = tile_cand_off[tile];
    for (uint32_t k = r.x; k < r.x + r.y && k < pre_cap; ++k)
        if (keep[k]) out[w++] = raw[k];
}

/* per region: candidate range in the ordered array */
__global__ void k_cand_ranges(uint32_t n_regions, const uint32_t *tile_base, const uint32_t *tile_cand_off, LcrRegionState *rstate) {
    const uint32_t reg = blockIdx.x * blockDim.x + threadIdx.x;
    if (reg >= n_regions) return;
    const uint32_t b = tile_cand_off[tile_base[reg]];
    uint32_t e = tile_cand_off[tile_base[reg + 1]];
    if (rstate[reg].status != 0) e = b; /* a failed region reports no candidates */
    rstate[reg].cand_begin = b;
    rstate[reg].n_cand = e - b;
}

/* the dense-cluster filters (candidate.rs:465-526), one thread per window start i.  The windows of different starts only
   set the same two flag bits and never read them, so they are independent of each other and of the order of the two passes. */
__global__ void k_cand_dense(lcr_params P, const LcrCounters *ctr, lcr_candidate *cand, const LcrRegionState *rstate) {
    const uint32_t n_cand = ctr->n_cand;
    for (uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x; gi < n_cand; gi += gridDim.x * blockDim.x) {
        const uint32_t reg = cand[gi].region;
        const LcrRegionState rs = rstate[reg];
        if (gi < rs.cand_begin || gi >= rs.cand_begin + rs.n_cand) continue;
        lcr_candidate *c = cand + rs.cand_begin;
        const uint32_t n = rs.n_cand, i = gi - rs.cand_begin;
        /* concat_idxes = homo_snps + het_snps, sorted: the candidates carrying HOM_VAR or HET_VAR */
        if (!(c[i].flags & (LCR_CF_HOM_VAR | LCR_CF_HET_VAR))) continue;
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
}

} // namespace

#define TRY(expr) LCR_CUDA_TRY(ctx, expr)

template <bool BIG, int STAGES, int MINB>
static cudaError_t launch_tile(const PileArgs &ka, int sms, uint32_t n_tiles, cudaStream_t st) {
    const size_t smem = PtLayout<BIG, STAGES, MINB>::bytes();
    static bool smem_set[8] = {false, false, false, false, false, false, false, false}; /* per device, once per instantiation */
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 8 || !smem_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_pileup_tile<BIG, STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 8) smem_set[dev] = true;
    }
    const int grid = (int)std::min<uint32_t>(n_tiles, (uint32_t)(MINB * sms));
    k_pileup_tile<BIG, STAGES, MINB><<<grid, PT_THREADS, smem, st>>>(ka);
    return cudaGetLastError();
}

/* Carves the stage's scratch from the arena (a dry arena only adds up the sizes) and, with launch set, enqueues the
   whole pileup / genotype stage on the context stream: no host synchronisation inside. */
int lcr_stage_pileup(lcr_ctx *ctx, lcr_device_batch *db, LcrArena &A, LcrCounters *ctr, bool launch) {
    cudaStream_t st = ctx->stream;
    const uint32_t n_tiles = db->n_tiles, n_slots = db->n_slots, n_regions = db->n_regions;
    const LcrCaps &C = db->caps;
    const size_t tn = (size_t)n_tiles + 1, sn = (size_t)n_slots + 1;
    /* one zeroed block: tile_pre [2 tn] | tile_diff [tn] | tile_cursor [tn] | tile_full_n [tn] */
    uint32_t *zero_blk = A.take<uint32_t>(5 * tn);
    uint32_t *tile_off = A.take<uint32_t>(tn);
    uint32_t *slot_seg_ub = A.take<uint32_t>(sn);
    uint32_t *slot_seg_off = A.take<uint32_t>(sn);
    uint64_t *slot_runs = A.take<uint64_t>((size_t)LCR_SLOT_RUNS * n_slots);
    LcrItem *items = A.take<LcrItem>(C.items);
    LcrSeg *segs = A.take<LcrSeg>(C.segs + 2);
    LcrTileDesc *desc = A.take<LcrTileDesc>(tn);
    uint32_t *list0 = A.take<uint32_t>(tn), *list1 = A.take<uint32_t>(tn);
    PreCand *pre = A.take<PreCand>(C.pre);
    lcr_candidate *cand_raw = A.take<lcr_candidate>(C.pre);
    uint8_t *keep = A.take<uint8_t>(C.pre);
    uint32_t *tile_cand_cnt = A.take<uint32_t>(tn), *tile_cand_off = A.take<uint32_t>(tn);
    size_t tmp_bytes = 0;
    {
        size_t b1 = 0, b2 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, b1, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)std::max(tn, sn), st);
        cub::DeviceScan::InclusiveSum(nullptr, b2, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)tn, st);
        tmp_bytes = std::max(b1, b2);
    }
    void *tmp = A.take<char>(tmp_bytes + 256);
    if (!launch) return LCR_OK;

    uint2 *tile_pre = reinterpret_cast<uint2 *>(zero_blk);
    uint32_t *tile_diff = zero_blk + 2 * tn, *tile_cursor = zero_blk + 3 * tn, *tile_full_n = zero_blk + 4 * tn;
    TRY(cudaMemsetAsync(zero_blk, 0, sizeof(uint32_t) * 5 * tn, st));
    TRY(cudaMemsetAsync(slot_seg_ub + n_slots, 0, sizeof(uint32_t), st));

    PrepArgs pa{};
    pa.P = ctx->P;
    pa.n_slots = n_slots;
    pa.regions = db->regions;
    pa.slot_off = db->slot_off; pa.slot_region = db->slot_region; pa.tile_base = db->tile_base;
    pa.pos = db->pos; pa.flag = db->flag; pa.mapq = db->mapq; pa.ts = db->ts; pa.de = db->de;
    pa.seq_off = db->seq_off; pa.cig_off = db->cig_off; pa.seq = db->seq; pa.cigar = db->cigar;
    pa.ref_table = ctx->d_ref_table;
    pa.rstate = db->rstate; pa.stats = db->d_stats; pa.ctr = ctr;
    pa.slot_flags = db->slot_flags; pa.slot_runs = slot_runs;
    pa.tile_diff = tile_diff; pa.tile_off = tile_off; pa.tile_cursor = tile_cursor; pa.tile_full_n = tile_full_n;
    pa.slot_seg_ub = slot_seg_ub; pa.slot_seg_off = slot_seg_off;
    pa.items = items; pa.segs = segs; pa.items_cap = C.items; pa.segs_cap = C.segs;
    const uint32_t pb = 128, pg = (n_slots + pb - 1) / pb;
    /* HiFi presets read the bases of the read ends in the span pass; ONT presets read no base before the tile kernel */
    if (ctx->P.platform != 1) {
        if (db->seq_wait_pending) {
            TRY(cudaStreamWaitEvent(st, db->ev_seq, 0));
            db->seq_wait_pending = false;
        }
        if (db->seq4_pending) { /* bases arrived in the BAM record's 4-bit form */
            const int urc = lcr_unpack_seq4(ctx, db, st);
            if (urc) return urc;
            db->timing.kernel_launches += 1;
        }
    }
    LCR_DEBUG_CHECK(ctx, "pileup stage memsets");
    if (pg) {
        k_read_span<<<pg, pb, 0, st>>>(pa);
        db->timing.kernel_launches += 1;
    }
    LCR_DEBUG_CHECK(ctx, "k_read_span");
    /* item slots per tile (running sum of the difference array, then its exclusive scan) and segment slots per read */
    TRY(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, tile_diff, tile_diff, (int)tn, st));
    TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, tile_diff, tile_off, (int)tn, st));
    TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, slot_seg_ub, slot_seg_off, (int)sn, st));
    k_fix_totals<<<1, 32, 0, st>>>(tile_off, n_tiles, slot_seg_off, n_slots, C.items, C.segs, ctr);
    LCR_DEBUG_CHECK(ctx, "prep scans");
    if (pg) k_read_walk<<<pg, pb, 0, st>>>(pa);
    db->timing.kernel_launches += 2;
    LCR_DEBUG_CHECK(ctx, "k_read_walk");

    DescArgs da{};
    da.n_tiles = n_tiles; da.regions = db->regions;
    da.tile_base = db->tile_base; da.tile_region = db->tile_region; da.tile_off = tile_off; da.tile_cursor = tile_cursor; da.tile_full_n = tile_full_n;
    da.pos_off = db->pos_off; da.ref_table = ctx->d_ref_table; da.rstate = db->rstate; da.desc = desc;
    da.list[0] = list0; da.list[1] = list1; da.ctr = ctr;
    da.all_tiles = db->pl_acgt != nullptr || db->ext_off != nullptr; /* imported candidates may sit on tiles no read row touches (whole-tile introns) */
    const bool force_big = ctx->tile_variant == 3; /* tests: every non-empty tile through the 32-bit flavour */
    da.big_rows = force_big ? 0u : 65535u;
    if (n_tiles) {
        k_tile_desc<<<(n_tiles + 255) / 256, 256, 0, st>>>(da);
        db->timing.kernel_launches += 1;
    }

    PileArgs ka{};
    ka.P = ctx->P;
    ka.regions = db->regions;
    ka.tile_base = db->tile_base; ka.tile_region = db->tile_region;
    ka.seq = db->seq; ka.qual = db->qual;
    ka.exon_off = db->exon_off; ka.exon_iv = db->exon_iv;
    ka.ext_off = db->ext_off; ka.ext_pos = db->ext_pos; ka.ext_gt = db->ext_gt; ka.ext_qual = db->ext_qual;
    ka.ref_table = ctx->d_ref_table;
    ka.desc = desc; ka.items = items; ka.segs = segs;
    ka.tables = ctx->d_tables;
    ka.rstate = db->rstate; ka.ctr = ctr; ka.stats = db->d_stats;
    ka.pl_acgt = db->pl_acgt; ka.pl_fwd = db->pl_fwd; ka.pl_d = db->pl_d; ka.pl_n = db->pl_n; ka.pl_ts = db->pl_ts;
    ka.pre = pre; ka.pre_cap = (uint32_t)std::min<uint64_t>(C.pre, 0xffffffffu); ka.tile_pre = tile_pre;
    ka.cand_raw = cand_raw; ka.cand_keep = keep;
    if (db->seq_wait_pending) { /* asynchronous upload: seq / qual are first read here */
        TRY(cudaStreamWaitEvent(st, db->ev_seq, 0));
        db->seq_wait_pending = false;
    }
    if (db->seq4_pending) {
        const int urc = lcr_unpack_seq4(ctx, db, st);
        if (urc) return urc;
        db->timing.kernel_launches += 1;
    }
    LCR_DEBUG_CHECK(ctx, "k_tile_desc");
    TRY(cudaEventRecord(ctx->ev_t[0], st));
    if (n_tiles) {
        const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
        ka.tile_list = list0; ka.list_id = 0;
        /* LCR_TILE_VARIANT (experiments): 0 = one 12 KB stage, 4 CTAs / SM; 1 = two 8 KB stages, 3 CTAs / SM; 2 = one stage, 3 CTAs / SM */
        if (ctx->tile_variant == 1) TRY((launch_tile<false, 2, 3>(ka, sms, n_tiles, st)));
        else if (ctx->tile_variant == 2) TRY((launch_tile<false, 1, 3>(ka, sms, n_tiles, st)));
        else TRY((launch_tile<false, 1, 4>(ka, sms, n_tiles, st)));
        db->timing.kernel_launches += 1;
        if (db->max_region_slots > 65535u || force_big) { /* only a region with more reads than that can hold a tile whose 16-bit event counters would overflow */
            ka.tile_list = list1; ka.list_id = 1;
            TRY((launch_tile<true, 1, 2>(ka, sms, n_tiles, st)));
            db->timing.kernel_launches += 1;
        }
    }
    TRY(cudaEventRecord(ctx->ev_t[1], st));
    LCR_DEBUG_CHECK(ctx, "k_pileup_tile");
    {
        const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
        k_site_ll<<<sms * 8, 256, 0, st>>>(ka);
        LCR_DEBUG_CHECK(ctx, "k_site_ll");
        k_tile_cand_count<<<(uint32_t)((tn + 127) / 128), 128, 0, st>>>(n_tiles, tile_pre, keep, ka.pre_cap, tile_cand_cnt);
        TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, tile_cand_cnt, tile_cand_off, (int)tn, st));
        if (n_tiles) k_tile_cand_gather<<<(n_tiles + 127) / 128, 128, 0, st>>>(n_tiles, tile_pre, keep, ka.pre_cap, tile_cand_off, cand_raw, db->cand, ctr);
        db->timing.kernel_launches += 3;
        if (n_regions) {
            k_cand_ranges<<<(n_regions + 63) / 64, 64, 0, st>>>(n_regions, db->tile_base, tile_cand_off, db->rstate);
            if (!db->ext_off) k_cand_dense<<<sms * 4, 64, 0, st>>>(ctx->P, ctr, db->cand, db->rstate); /* imported candidates skip the dense filters (candidate.rs:530-613 has none) */
            db->timing.kernel_launches += db->ext_off ? 1 : 2;
        }
    }
    LCR_DEBUG_CHECK(ctx, "candidate compaction");
    TRY(cudaGetLastError());
    return LCR_OK;
}
