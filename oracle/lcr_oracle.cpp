/*
 * lcr_oracle.cpp — CPU restatement of longcallR's per-region worker body.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (longcallr_b200/) may
 * include, link or call this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py load liblcr_oracle.so.
 *
 * PARITY UNPINNED: the reference (huangnengCSU/longcallR v1.12.0, Rust) ships no
 * tests, no golden vectors and no expected outputs, and neither cargo nor rustc
 * exist in this environment, so this restatement could not be checked against
 * a run of the reference binary.  It is pinned only against known-answer values
 * derived by hand from the cited lines (tests/test_oracle_kat.py) and is written
 * to be audited line by line against:
 *
 *   read filter + fetch window      src/util.rs:636-668      src/fragment.rs:19-54
 *   pileup                          src/util.rs:621-949
 *   two major alleles               src/util.rs:162-176
 *   candidate filters + genotype    src/candidate.rs:24-51, 54-528
 *   fragment matrix + pair counts   src/fragment.rs:10-309
 *   LD blocks                       src/candidate.rs:615-747  src/snp.rs:158-195
 *   phasing sweeps                  src/phase.rs:32-49, 77-355, 600-691, 810-1394
 *   read / SNP assignment, rescue   src/snpfrags.rs:191-625
 *   phase sets                      src/snpfrags.rs:628-733
 *   orchestration                   src/thread.rs:78-221
 *
 * Third-party behaviour restated from the published algorithms (not in /root/reference):
 *   petgraph 0.6.4 GraphMap / kosaraju_scc / Dfs / DfsPostOrder / Bfs  (ordering only)
 *   rust-htslib 0.46 CigarStringView::leading_softclips / trailing_softclips
 *       (first / last op, looking through one hard clip), Record::seq_len, htslib bam_endpos + fetch overlap test
 *   statrs 0.16 Binomial::cdf: replaced by the exact integer evaluation
 *       lcr_binom_two_tailed_lt_0p05 (include/lcr_contract.h)
 *   rand 0.8.5 thread_rng: replaced by lcr_uniform (include/lcr_contract.h)
 *
 * Two modes:
 *   mode 0  "contract": order-independent int64 fixed-point sums, lcr_log10 / lcr_exp10;
 *           this is what the CUDA path must reproduce bit-for-bit.
 *   mode 1  "reference-order f64": sequential f64 sums in the reference's loop order
 *           with libm log10 / pow, and the reference's cost profile (per-position qual
 *           vectors, per-position intron increments, hash-map pair counts, row rescans,
 *           the sweep evaluated three times per iteration).  Used as the CPU baseline
 *           ("port") and to check that the contract does not change any discrete output.
 */
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "longcallr_b200.h"
#include "lcr_contract.h"

/* oracle-only test hook (tests/test_oracle_crosscheck.py): stop after phase() and report its state */
#define LCR_ORACLE_FLAG_STOP_AFTER_PHASE 0x40000000u

namespace {

/* ---------------------------------------------------------------- types -- */

struct BaseFreq { /* util.rs:100-127, live fields */
    uint32_t a = 0, c = 0, g = 0, t = 0, n = 0, d = 0;
    uint8_t ref_base = 0;
    std::vector<uint8_t> bq[4]; /* baseq.a/c/g/t */
    int32_t strands[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    int32_t ts[2] = {0, 0};
    /* contract-mode sums (mode 0 does not keep the qual vectors) */
    int64_t ll0 = 0, ll2 = 0; /* homvar / homref log10-likelihood, fixed point */
    uint32_t q0_ref = 0, q0_non = 0;
    uint32_t bq_pass[4] = {0, 0, 0, 0};
};

struct Cand { /* snp.rs:39-90 */
    int64_t pos = 0;
    uint8_t alleles[2] = {0, 0};
    float allele_freqs[2] = {0, 0};
    uint8_t reference = 0;
    uint32_t depth = 0;
    int variant_type = 0;
    double variant_quality = 0;
    double genotype_probability[3] = {0, 0, 0};
    double genotype_quality = 0;
    int genotype = 0;
    int haplotype = 0;
    double phase_score = 0;
    std::vector<uint32_t> cover; /* snp_cover_fragments */
    bool rna_editing = false, dense = false, het_var = false, for_phasing = false, hom_var = false;
    bool single = false, non_selected = false, cand_somatic = false;
    bool in_edit = false, in_somatic = false;
    uint32_t phase_set = 0;
};

struct FragElem { /* snp.rs:197-215 */
    uint32_t snp_idx;
    uint8_t base;
    uint8_t baseq;
    int8_t p;
    bool phase_site;
};

struct Fragment { /* snp.rs:217-239 */
    uint32_t read; /* read index in the batch (stands for read_id) */
    std::vector<FragElem> list;
    int haplotag = 0;
    int assignment = 0;
    double assignment_score = 0;
    uint32_t num_hete_links = 0;
    bool for_phasing = false;
    bool downsampled = false; /* snp.rs:237 */
};

struct PairKeyHash {
    size_t operator()(uint64_t k) const { return (size_t)lcr_mix64(k); }
};

struct LdPair { /* snp.rs:92-104; counts indexed [base1][base2] with A,C,G,T = 0..3 */
    uint32_t cnt[4][4] = {{0}};
    bool valid = false;
    float score = 0;
    int weight = 0;
};

struct Graph { /* petgraph GraphMap<usize,_,Undirected>: insertion-ordered nodes and adjacency */
    std::vector<uint32_t> nodes;
    std::unordered_map<uint32_t, uint32_t> slot;
    std::vector<std::vector<uint32_t>> adj;
    uint32_t add_node(uint32_t v) {
        auto it = slot.find(v);
        if (it != slot.end()) return it->second;
        uint32_t s = (uint32_t)nodes.size();
        slot.emplace(v, s);
        nodes.push_back(v);
        adj.emplace_back();
        return s;
    }
    bool has_node(uint32_t v) const { return slot.find(v) != slot.end(); }
    bool has_edge(uint32_t a, uint32_t b) const {
        auto it = slot.find(a);
        if (it == slot.end()) return false;
        for (uint32_t x : adj[it->second])
            if (x == b) return true;
        return false;
    }
    void add_edge(uint32_t a, uint32_t b) { /* caller checks has_edge first */
        uint32_t sa = add_node(a);
        uint32_t sb = add_node(b);
        adj[sa].push_back(b);
        if (a != b) adj[sb].push_back(a);
    }
};

/* petgraph::algo::kosaraju_scc on an undirected GraphMap: components come out in
   reverse order of their first-inserted node; each lists its nodes in the order of a
   stack DFS (Dfs::next) started at that node. */
std::vector<std::vector<uint32_t>> kosaraju_scc(const Graph &g) {
    size_t n = g.nodes.size();
    std::vector<char> disc(n, 0), fin(n, 0);
    std::vector<uint32_t> finish_order;
    std::vector<uint32_t> stack;
    for (size_t i = 0; i < n; ++i) { /* DfsPostOrder over Reversed(g) == g */
        if (disc[i]) continue;
        stack.clear();
        stack.push_back((uint32_t)i);
        while (!stack.empty()) {
            uint32_t nx = stack.back();
            if (!disc[nx]) {
                disc[nx] = 1;
                for (uint32_t succ : g.adj[nx]) {
                    uint32_t s = g.slot.at(succ);
                    if (!disc[s]) stack.push_back(s);
                }
            } else {
                stack.pop_back();
                if (!fin[nx]) {
                    fin[nx] = 1;
                    finish_order.push_back(nx);
                }
            }
        }
    }
    std::vector<std::vector<uint32_t>> sccs;
    std::fill(disc.begin(), disc.end(), 0);
    for (size_t r = finish_order.size(); r-- > 0;) {
        uint32_t i = finish_order[r];
        if (disc[i]) continue;
        std::vector<uint32_t> scc;
        stack.clear();
        stack.push_back(i);
        while (!stack.empty()) {
            uint32_t node = stack.back();
            stack.pop_back();
            if (disc[node]) continue;
            disc[node] = 1;
            for (uint32_t succ : g.adj[node]) {
                uint32_t s = g.slot.at(succ);
                if (!disc[s]) stack.push_back(s);
            }
            scc.push_back(g.nodes[node]);
        }
        sccs.push_back(std::move(scc));
    }
    return sccs;
}

inline int base_code(uint8_t b) {
    switch (b) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

struct RegionOut {
    int32_t status = 0;
    std::vector<Cand> cands;
    std::vector<Fragment> frags;
    std::vector<std::pair<uint32_t, int>> hp;      /* read -> HP (thread.rs:181) */
    std::vector<std::pair<uint32_t, uint32_t>> ps; /* read -> PS (thread.rs:201) */
    std::vector<uint32_t> fragment_reads;
    std::vector<BaseFreq> pileup; /* kept only with EMIT_PLANES */
    lcr_stats st{};
};

/* ------------------------------------------------------------- worker ---- */

template <bool FX>
struct Worker {
    const lcr_params &P;
    const lcr_batch &B;
    const lcr_luts &T;
    const lcr_region &reg;
    const uint8_t *ref_seq;
    uint64_t ref_len;
    uint64_t region_key;
    uint32_t region_index; /* of reg in the batch (exon intervals are per region) */
    RegionOut &out;

    std::vector<Cand> &cands;
    std::vector<Fragment> &frags;
    std::vector<uint32_t> homo_snps, het_snps, edit_snps, somatic_snps;
    std::unordered_map<uint64_t, LdPair, PairKeyHash> allele_pairs;
    std::vector<std::vector<uint32_t>> ld_blocks;

    Worker(const lcr_params &p, const lcr_batch &b, const lcr_luts &t, const lcr_region &r, const uint8_t *rs,
           uint64_t rl, RegionOut &o)
        : P(p), B(b), T(t), reg(r), ref_seq(rs), ref_len(rl), region_key(lcr_region_key(r.tid, r.start)),
          region_index((uint32_t)(&r - b.regions)), out(o), cands(o.cands), frags(o.frags) {}

    /* ---- P0: read filter (util.rs:652-668) and htslib fetch window (util.rs:636-638) */
    bool read_pass(uint32_t i) const {
        uint64_t l_seq = B.seq_off[i + 1] - B.seq_off[i];
        if ((int32_t)B.mapq[i] < P.min_mapq || l_seq < (uint64_t)P.min_read_length || (B.flag[i] & 0x4) ||
            (B.flag[i] & 0x100) || (B.flag[i] & 0x800))
            return false;
        float de = B.de[i];
        if (!(de != de) && de >= P.divergence) return false;
        return true;
    }
    int64_t ref_span(uint32_t i) const {
        int64_t rlen = 0;
        for (uint64_t c = B.cig_off[i]; c < B.cig_off[i + 1]; ++c) {
            uint32_t op = B.cigar[c] & 0xf, len = B.cigar[c] >> 4;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += len;
        }
        return rlen;
    }
    /* fetch((chr, start, end)) hands the 1-based region numbers to htslib's 0-based
       half-open query: records with pos < end and bam_endpos > start */
    bool in_window(uint32_t i) const {
        int64_t pos = B.pos[i];
        int64_t rlen = ref_span(i);
        int64_t endpos = pos + (rlen ? rlen : 1);
        return pos < (int64_t)reg.end && endpos > (int64_t)reg.start;
    }
    /* rust-htslib CigarStringView::leading_softclips / trailing_softclips: the soft clip at the end of the
       alignment, looking through one hard clip (`5H10S...` has 10 leading soft clips, `...7S3H` 7 trailing ones);
       util.rs:682-690 and fragment.rs:59 start pos_in_read there */
    int64_t leading_softclips(uint32_t i) const {
        uint64_t a = B.cig_off[i], b = B.cig_off[i + 1];
        if (a == b) return 0;
        if ((B.cigar[a] & 0xf) == 4) return (int64_t)(B.cigar[a] >> 4);
        if ((B.cigar[a] & 0xf) == 5 && a + 1 < b && (B.cigar[a + 1] & 0xf) == 4) return (int64_t)(B.cigar[a + 1] >> 4);
        return 0;
    }
    int64_t trailing_softclips(uint32_t i) const {
        uint64_t a = B.cig_off[i], b = B.cig_off[i + 1];
        if (a == b) return 0;
        if ((B.cigar[b - 1] & 0xf) == 4) return (int64_t)(B.cigar[b - 1] >> 4);
        if ((B.cigar[b - 1] & 0xf) == 5 && b - a >= 2 && (B.cigar[b - 2] & 0xf) == 4) return (int64_t)(B.cigar[b - 2] >> 4);
        return 0;
    }

    /* ---- P1: Profile::fill_data_into_freq_vec (util.rs:621-949) */
    int pileup(std::vector<BaseFreq> &fv) {
        const int64_t vec_size = (int64_t)reg.end - (int64_t)reg.start;
        const int64_t fv_start = (int64_t)reg.start - 1;
        fv.assign((size_t)vec_size, BaseFreq());
        for (int64_t i = 0; i < vec_size; ++i) {
            if ((uint64_t)(fv_start + i) >= ref_len) return LCR_ERR_INVALID_ARG;
            fv[i].ref_base = ref_seq[fv_start + i];
        }
        const int64_t polya = (int64_t)P.polya_tail_length;
        const int64_t dist_end = (int64_t)P.distance_to_read_end;
        for (uint32_t r = reg.read_begin; r < reg.read_end; ++r) {
            if (!in_window(r) || !read_pass(r)) continue;
            out.st.n_reads_pass++;
            const uint8_t *seq = B.seq + B.seq_off[r];
            const uint8_t *qual = B.qual + B.seq_off[r];
            const int64_t seq_len = (int64_t)(B.seq_off[r + 1] - B.seq_off[r]);
            const int strand = (B.flag[r] & 0x10) ? 1 : 0;
            const int8_t ts = B.ts[r];
            const int64_t lead = leading_softclips(r), trail = trailing_softclips(r);
            int64_t pos_in_fv = (int64_t)B.pos[r] - fv_start;
            int64_t pos_in_read = lead > 0 ? lead : 0;
            bool stop = false;
            for (uint64_t ci = B.cig_off[r]; ci < B.cig_off[r + 1] && !stop; ++ci) {
                const uint32_t op = B.cigar[ci] & 0xf, len = B.cigar[ci] >> 4;
                switch (op) {
                    case 4: case 5: break; /* S, H */
                    case 0: case 7: case 8: { /* M = X */
                        for (uint32_t cgi = 0; cgi < len; ++cgi) {
                            if (pos_in_fv < 0) { pos_in_fv++; pos_in_read++; continue; }
                            if (pos_in_fv >= vec_size) break;
                            out.st.n_aligned_bases++;
                            if (pos_in_read >= seq_len) return LCR_ERR_BAD_CIGAR;
                            const uint8_t base = seq[pos_in_read];
                            const uint8_t baseq = qual[pos_in_read] < LCR_MAX_BASE_QUALITY ? qual[pos_in_read] : LCR_MAX_BASE_QUALITY;
                            BaseFreq &bf = fv[pos_in_fv];
                            const uint8_t ref_base = bf.ref_base;
                            bool poly_a = false, homopolymer = false, trim = false;
                            const int64_t curr = pos_in_read;
                            const int64_t read_end_boundary = seq_len - trail;
                            const bool near_end = std::llabs(curr - lead) < dist_end || std::llabs(curr - read_end_boundary) < dist_end;
                            if (P.platform == 1 && near_end) trim = true;
                            if (!trim && near_end) {
                                for (int64_t ti = curr - polya; ti <= curr + 1; ++ti) {
                                    if (ti < 0 || ti + polya - 1 >= seq_len) continue;
                                    int64_t pc[4] = {0, 0, 0, 0};
                                    for (int64_t tj = 0; tj < polya; ++tj) {
                                        uint8_t b = seq[ti + tj];
                                        if (b == 'A' && ref_base != 'A') pc[0]++;
                                        else if (b == 'T' && ref_base != 'T') pc[1]++;
                                        else if (b == 'C' && ref_base != 'C') pc[2]++;
                                        else if (b == 'G' && ref_base != 'G') pc[3]++;
                                    }
                                    if (pc[0] >= polya || pc[1] >= polya) poly_a = true;
                                    if (pc[2] >= polya || pc[3] >= polya) homopolymer = true;
                                }
                            }
                            if (!trim && !poly_a && !homopolymer) {
                                if (strand == 0) {
                                    if (ts == '+') bf.ts[0]++;
                                    else if (ts == '-') bf.ts[1]++;
                                } else {
                                    if (ts == '+') bf.ts[1]++;
                                    else if (ts == '-') bf.ts[0]++;
                                }
                                const int bc = base_code(base);
                                if (bc >= 0) {
                                    (bc == 0 ? bf.a : bc == 1 ? bf.c : bc == 2 ? bf.g : bf.t) += 1;
                                    bf.strands[bc][strand] += 1;
                                    if (FX) {
                                        /* contract: what the qual vector is used for, accumulated in place */
                                        if (baseq >= (uint32_t)P.min_baseq) bf.bq_pass[bc]++;
                                        /* candidate.rs:243-254: identical_baseqs is the vector of the allele equal to an upper-case reference byte */
                                        const bool is_ref = (ref_base == 'A' || ref_base == 'C' || ref_base == 'G' || ref_base == 'T') && bc == base_code(ref_base);
                                        if (is_ref) {
                                            bf.ll0 += T.gl_fx_err[baseq];
                                            if (baseq == 0) bf.q0_ref++; else bf.ll2 += T.gl_fx_ok[baseq];
                                        } else {
                                            bf.ll2 += T.gl_fx_err[baseq];
                                            if (baseq == 0) bf.q0_non++; else bf.ll0 += T.gl_fx_ok[baseq];
                                        }
                                    } else {
                                        bf.bq[bc].push_back(baseq);
                                    }
                                }
                            }
                            pos_in_fv++;
                            pos_in_read++;
                        }
                        break;
                    }
                    case 2: /* D */
                        for (uint32_t k = 0; k < len; ++k) {
                            if (pos_in_fv < 0) { pos_in_fv++; continue; }
                            if (pos_in_fv >= vec_size) break;
                            fv[pos_in_fv].d += 1;
                            pos_in_fv++;
                        }
                        break;
                    case 1: /* I */
                        if (pos_in_fv < 1) { pos_in_read += len; break; }
                        if (pos_in_fv >= vec_size) { stop = true; break; }
                        pos_in_read += len; /* ni is never read */
                        break;
                    case 3: /* N */
                        if (FX) { /* same counts, added as a range */
                            const int64_t a = std::max<int64_t>(pos_in_fv, 0), b = std::min<int64_t>(pos_in_fv + len, vec_size);
                            for (int64_t x = a; x < b; ++x) fv[x].n += 1;
                            if (pos_in_fv < vec_size) pos_in_fv = std::min<int64_t>(pos_in_fv + len, vec_size);
                        } else {
                            for (uint32_t k = 0; k < len; ++k) {
                                if (pos_in_fv < 0) { pos_in_fv++; continue; }
                                if (pos_in_fv >= vec_size) break;
                                fv[pos_in_fv].n += 1;
                                pos_in_fv++;
                            }
                        }
                        break;
                    default: return LCR_ERR_BAD_CIGAR; /* util.rs:943-945 panic */
                }
            }
        }
        return 0;
    }

    /* ---- P2: BaseFreq::get_two_major_alleles (util.rs:162-176) */
    static void two_major(const BaseFreq &bf, uint8_t ref_base, uint8_t &a1, uint32_t &c1, uint8_t &a2, uint32_t &c2) {
        std::pair<uint8_t, uint32_t> x[4] = {{'A', bf.a}, {'C', bf.c}, {'G', bf.g}, {'T', bf.t}};
        std::stable_sort(x, x + 4, [](const auto &l, const auto &r) { return l.second > r.second; });
        if (x[0].first != ref_base && x[1].first != ref_base) {
            if (x[2].second == x[1].second && x[2].first == ref_base) { a1 = x[0].first; c1 = x[0].second; a2 = x[2].first; c2 = x[2].second; return; }
            if (x[3].second == x[1].second && x[3].first == ref_base) { a1 = x[0].first; c1 = x[0].second; a2 = x[3].first; c2 = x[3].second; return; }
        }
        a1 = x[0].first; c1 = x[0].second; a2 = x[1].first; c2 = x[1].second;
    }

    static double xlog10(double v) { return FX ? lcr_log10(v) : std::log10(v); }
    static double xexp10(double v) { return FX ? lcr_exp10(v) : std::pow(10.0, v); }

    /* ---- SNPFrag::import_external_candidates (candidate.rs:530-613), called with min_variant_qual = 0.0 (thread.rs:110-115) */
    void import_external_candidates(const std::vector<BaseFreq> &pileup) {
        int64_t position = (int64_t)reg.start - 1;
        uint32_t e = B.ext_off[region_index];
        const uint32_t e1 = B.ext_off[region_index + 1];
        for (size_t bfidx = 0; bfidx < pileup.size(); ++bfidx, ++position) {
            while (e < e1 && (int64_t)B.ext_pos[e] < position) ++e;
            if (e >= e1 || (int64_t)B.ext_pos[e] != position) continue; /* contains_key(position) */
            const BaseFreq &bf = pileup[bfidx];
            uint8_t allele1, allele2;
            uint32_t allele1_cnt, allele2_cnt;
            two_major(bf, bf.ref_base, allele1, allele1_cnt, allele2, allele2_cnt);
            const float quality = B.ext_qual[e];
            if (quality < 0.0f) continue;
            const uint32_t total = bf.a + bf.c + bf.g + bf.t;
            Cand cs;
            cs.pos = position;
            cs.reference = bf.ref_base; /* ref_seq[ref_pos]: the same byte the pileup recorded */
            cs.alleles[0] = allele1; cs.alleles[1] = allele2;
            cs.allele_freqs[0] = (float)allele1_cnt / (float)total; cs.allele_freqs[1] = (float)allele2_cnt / (float)total;
            cs.depth = total;
            cs.variant_quality = (double)quality;
            cs.genotype_quality = (double)quality;
            switch (B.ext_gt[e]) {
                case 1: cs.variant_type = 1; cs.genotype = 0; cs.for_phasing = true; cs.het_var = true; cands.push_back(cs); break;
                case 2: cs.variant_type = 2; cs.genotype = -1; cs.for_phasing = true; cs.hom_var = true; cands.push_back(cs); break;
                case 3: cs.variant_type = 3; cs.genotype = -1; cs.hom_var = true; cands.push_back(cs); break;
                default: break; /* 0/0 builds a record that is never pushed; other genotypes are reported and skipped */
            }
        }
    }

    /* ---- P3-P7: SNPFrag::get_candidate_snps (candidate.rs:54-528) */
    void get_candidate_snps(const std::vector<BaseFreq> &pileup) {
        int64_t position = (int64_t)reg.start - 1;
        for (size_t bfidx = 0; bfidx < pileup.size(); ++bfidx, ++position) {
            const BaseFreq &bf = pileup[bfidx];
            if (B.exon_off) { /* --exon-only (candidate.rs:80-89): Lapper::find(position + 1, position + 2) over the region's exon intervals */
                bool hit = false;
                for (uint32_t e = B.exon_off[region_index]; e < B.exon_off[region_index + 1] && !hit; ++e)
                    hit = (int64_t)B.exon_iv[2 * e] < position + 2 && (int64_t)B.exon_iv[2 * e + 1] > position + 1;
                if (!hit) continue;
            }
            const uint32_t total = bf.a + bf.c + bf.g + bf.t;
            if (total < P.min_depth || total > P.max_depth) continue;
            uint8_t allele1, allele2;
            uint32_t allele1_cnt, allele2_cnt;
            two_major(bf, bf.ref_base, allele1, allele1_cnt, allele2, allele2_cnt);
            const float allele1_freq = (float)allele1_cnt / (float)total;
            const float allele2_freq = (float)allele2_cnt / (float)total;
            uint8_t ref_allele_base;
            uint32_t alt_num;
            uint8_t alt_base[2] = {0, 0};
            float alt_freq[2] = {0, 0};
            uint32_t alt_cnt[2] = {0, 0};
            if (allele1 == bf.ref_base) {
                ref_allele_base = allele1; alt_num = 1;
                alt_base[0] = allele2; alt_freq[0] = allele2_freq; alt_cnt[0] = allele2_cnt;
            } else if (allele2 == bf.ref_base) {
                ref_allele_base = allele2; alt_num = 1;
                alt_base[0] = allele1; alt_freq[0] = allele1_freq; alt_cnt[0] = allele1_cnt;
            } else {
                ref_allele_base = bf.ref_base; alt_num = 2;
                alt_base[0] = allele1; alt_freq[0] = allele1_freq; alt_cnt[0] = allele1_cnt;
                alt_base[1] = allele2; alt_freq[1] = allele2_freq; alt_cnt[1] = allele2_cnt;
            }
            if (base_code(ref_allele_base) < 0) continue; /* VALID_ALLELES, main.rs:23 */
            if (alt_num == 1) {
                if (total < 200 && alt_freq[0] < P.low_allele_frac_cutoff) continue;
                if (total >= 200 && alt_cnt[0] < P.low_allele_cnt_cutoff) continue;
            }
            if (bf.d >= alt_cnt[0]) continue;
            const uint32_t depth_incl = bf.a + bf.c + bf.g + bf.t + bf.d + bf.n;
            if ((float)(allele1_cnt + allele2_cnt) / (float)depth_incl < P.min_allele_freq_include_intron) continue;
            /* alt allele needs >= 2 bases of quality >= min_baseq (candidate.rs:177-194) */
            auto pass_cnt = [&](uint8_t allele) -> uint32_t {
                int bc = base_code(allele);
                if (FX) return bf.bq_pass[bc];
                uint32_t n = 0;
                for (uint8_t q : bf.bq[bc]) if (q >= (uint32_t)P.min_baseq) n++;
                return n;
            };
            if (allele1 != bf.ref_base) {
                if (allele1_cnt > 0 && pass_cnt(allele1) < 2) continue;
            } else if (allele2 != bf.ref_base) {
                if (allele2_cnt > 0 && pass_cnt(allele2) < 2) continue;
            }
            if (P.use_strand_bias) { /* candidate.rs:199-234 */
                const int32_t *rs = bf.strands[base_code(ref_allele_base)];
                float sor;
                if (alt_num == 1) {
                    const int32_t *as = bf.strands[base_code(alt_base[0])];
                    sor = lcr_strand_odds_ratio(rs[0], rs[1], as[0], as[1]);
                } else {
                    const int32_t *a1s = bf.strands[base_code(alt_base[0])];
                    const int32_t *a2s = bf.strands[base_code(alt_base[1])];
                    float s1 = lcr_strand_odds_ratio(rs[0], rs[1], a1s[0], a1s[1]);
                    float s2 = lcr_strand_odds_ratio(rs[0], rs[1], a2s[0], a2s[1]);
                    sor = fmaxf(s1, s2);
                }
                if (sor > T.sor_threshold) continue;
                if (alt_num == 1) {
                    const int32_t *as = bf.strands[base_code(alt_base[0])];
                    if (as[0] + as[1] <= 30) {
                        if (lcr_binom_two_tailed_lt_0p05((uint32_t)as[0], (uint32_t)(as[0] + as[1]))) continue;
                    }
                    if ((int64_t)as[0] * (int64_t)as[1] == 0) continue;
                }
            }
            /* genotype likelihood (candidate.rs:236-335) */
            double loglikelihood[3] = {0.0, 0.0, 0.0};
            if (bf.ref_base == 'A' || bf.ref_base == 'C' || bf.ref_base == 'G' || bf.ref_base == 'T') {
                if (FX) {
                    loglikelihood[0] = bf.q0_non ? -INFINITY : lcr_fx_to_f64(bf.ll0);
                    loglikelihood[2] = bf.q0_ref ? -INFINITY : lcr_fx_to_f64(bf.ll2);
                } else {
                    const int rc = base_code(bf.ref_base);
                    for (uint8_t bq : bf.bq[rc]) {
                        double error_rate = std::pow(0.1, (double)bq / 10.0);
                        loglikelihood[0] += std::log10(error_rate);
                        loglikelihood[2] += std::log10(1.0 - error_rate);
                    }
                    for (int oc = 0; oc < 4; ++oc) {
                        if (oc == rc) continue;
                        for (uint8_t bq : bf.bq[oc]) {
                            double error_rate = std::pow(0.1, (double)bq / 10.0);
                            loglikelihood[0] += std::log10(1.0 - error_rate);
                            loglikelihood[2] += std::log10(error_rate);
                        }
                    }
                }
            } else {
                continue; /* 'N' or any other reference byte (candidate.rs:255-265) */
            }
            const uint32_t num_reads = total;
            loglikelihood[1] -= (double)num_reads * T.log10_2;
            double logprob[3] = {loglikelihood[0] + T.gl_prior_log[0], loglikelihood[1] + T.gl_prior_log[1], loglikelihood[2] + T.gl_prior_log[2]};
            const double max_logprob = fmax(fmax(logprob[0], logprob[1]), logprob[2]);
            for (double &v : logprob) v -= max_logprob;
            double variant_prob[3] = {xexp10(logprob[0]), xexp10(logprob[1]), xexp10(logprob[2])};
            const double sum_vp = variant_prob[0] + variant_prob[1] + variant_prob[2];
            for (double &v : variant_prob) v /= sum_vp;
            const double variant_quality = -10.0 * xlog10(fmax(10e-301, variant_prob[2]));
            const double max_ll = fmax(fmax(loglikelihood[0], loglikelihood[1]), loglikelihood[2]);
            double l10[3] = {xexp10(loglikelihood[0] - max_ll), xexp10(loglikelihood[1] - max_ll), xexp10(loglikelihood[2] - max_ll)};
            const double sum_l10 = l10[0] + l10[1] + l10[2];
            const double genotype_prob[3] = {l10[0] / sum_l10, l10[1] / sum_l10, l10[2] / sum_l10};
            double phred[3] = {-10.0 * xlog10(genotype_prob[0]), -10.0 * xlog10(genotype_prob[1]), -10.0 * xlog10(genotype_prob[2])};
            /* sort_by(cmp_f64) of three values */
            if (phred[1] < phred[0]) std::swap(phred[0], phred[1]);
            if (phred[2] < phred[1]) { std::swap(phred[1], phred[2]); if (phred[1] < phred[0]) std::swap(phred[0], phred[1]); }
            const double genotype_quality = phred[1] - phred[0];

            Cand cs;
            cs.pos = position;
            cs.alleles[0] = allele1; cs.alleles[1] = allele2;
            cs.allele_freqs[0] = allele1_freq; cs.allele_freqs[1] = allele2_freq;
            cs.reference = bf.ref_base;
            cs.depth = total;
            cs.variant_quality = variant_quality;
            memcpy(cs.genotype_probability, genotype_prob, sizeof genotype_prob);
            cs.genotype_quality = genotype_quality;
            if (genotype_prob[0] > genotype_prob[1] && genotype_prob[0] > genotype_prob[2]) { cs.variant_type = 2; cs.genotype = -1; }
            else if (genotype_prob[1] > genotype_prob[0] && genotype_prob[1] > genotype_prob[2]) { cs.variant_type = 1; cs.genotype = 0; }
            else { cs.variant_type = 0; cs.genotype = 1; }
            if (variant_quality < (double)P.min_qual) continue;

            const int32_t fwd_ts = bf.ts[0], rev_ts = bf.ts[1];
            if (ref_allele_base == 'A' && alt_base[0] == 'G' && (fwd_ts > rev_ts * 2 || (fwd_ts == 0 && rev_ts == 0)) && cs.variant_type != 2) {
                cs.rna_editing = true; cs.for_phasing = false; cs.in_edit = true;
                cands.push_back(cs); edit_snps.push_back((uint32_t)cands.size() - 1);
                continue;
            }
            if (ref_allele_base == 'T' && alt_base[0] == 'C' && (rev_ts > fwd_ts * 2 || (fwd_ts == 0 && rev_ts == 0)) && cs.variant_type != 2) {
                cs.rna_editing = true; cs.for_phasing = false; cs.in_edit = true;
                cands.push_back(cs); edit_snps.push_back((uint32_t)cands.size() - 1);
                continue;
            }
            if (alt_num == 1 && alt_freq[0] < P.min_allele_freq) {
                cs.cand_somatic = true; cs.for_phasing = false; cs.in_somatic = true;
                cands.push_back(cs); somatic_snps.push_back((uint32_t)cands.size() - 1);
                continue;
            }
            if (cs.variant_type == 2) {
                if (alt_num == 2 && alt_freq[0] >= P.min_allele_freq && alt_freq[1] >= P.min_allele_freq) { cs.variant_type = 3; cs.genotype = -1; }
                cs.hom_var = true; cs.for_phasing = true;
                cands.push_back(cs); homo_snps.push_back((uint32_t)cands.size() - 1);
                continue;
            }
            if (cs.variant_type == 1) {
                if (alt_num == 2) {
                    cs.variant_type = 3; cs.genotype = -1; cs.hom_var = true; cs.for_phasing = true;
                    cands.push_back(cs); homo_snps.push_back((uint32_t)cands.size() - 1);
                    continue;
                }
                cs.het_var = true; cs.for_phasing = true;
                cands.push_back(cs); het_snps.push_back((uint32_t)cands.size() - 1);
                continue;
            }
        }
        /* dense-cluster filters (candidate.rs:465-526) */
        std::vector<uint32_t> idx(homo_snps);
        idx.insert(idx.end(), het_snps.begin(), het_snps.end());
        std::sort(idx.begin(), idx.end());
        auto dense_pass = [&](int64_t win, bool ge, uint32_t min_cnt) {
            for (size_t i = 0; i < idx.size(); ++i) {
                const int64_t start_pos = cands[idx[i]].pos;
                for (size_t j = i; j < idx.size(); ++j) {
                    const int64_t diff = cands[idx[j]].pos - start_pos;
                    if (ge ? diff >= win : diff > win) {
                        if ((uint32_t)(j - i) >= min_cnt)
                            for (size_t tk = i; tk < j; ++tk) { cands[idx[tk]].dense = true; cands[idx[tk]].for_phasing = false; }
                        break;
                    }
                    if (j == idx.size() - 1 && (uint32_t)(j - i + 1) >= min_cnt)
                        for (size_t tk = i; tk < j; ++tk) { cands[idx[tk]].dense = true; cands[idx[tk]].for_phasing = false; }
                }
            }
        };
        dense_pass((int64_t)P.dense_win_size, false, P.min_dense_cnt);
        dense_pass(5, true, 3);
        auto drop = [&](std::vector<uint32_t> &v) { v.erase(std::remove_if(v.begin(), v.end(), [&](uint32_t i) { return cands[i].dense; }), v.end()); };
        drop(homo_snps);
        drop(het_snps);
    }

    /* ---- F1-F3: SNPFrag::get_fragments (fragment.rs:10-309) */
    int get_fragments() {
        if (cands.empty()) return 0;
        for (uint32_t r = reg.read_begin; r < reg.read_end; ++r) {
            if (!in_window(r) || !read_pass(r)) continue;
            const int64_t pos = B.pos[r];
            if (pos > cands.back().pos) continue;
            const uint8_t *seq = B.seq + B.seq_off[r];
            const uint8_t *qual = B.qual + B.seq_off[r];
            const int64_t seq_len = (int64_t)(B.seq_off[r + 1] - B.seq_off[r]);
            int64_t pos_on_ref = pos;
            int64_t pos_on_query = leading_softclips(r);
            size_t idx = 0;
            if (!(pos <= cands.front().pos))
                while (idx < cands.size() && cands[idx].pos < pos) idx++;
            int64_t snp_pos = cands[idx].pos;
            Fragment fragment;
            fragment.read = r;
            auto advance = [&]() {
                idx++;
                if (idx < cands.size()) snp_pos = cands[idx].pos;
            };
            for (uint64_t ci = B.cig_off[r]; ci < B.cig_off[r + 1]; ++ci) {
                const uint32_t op = B.cigar[ci] & 0xf, len = B.cigar[ci] >> 4;
                switch (op) {
                    case 4: case 5: break;
                    case 0: case 7: case 8:
                        if (!FX) {
                            for (uint32_t k = 0; k < len; ++k) {
                                if (pos_on_ref == snp_pos) {
                                    if (pos_on_query >= seq_len) return LCR_ERR_BAD_CIGAR;
                                    emit_elem(fragment, (uint32_t)idx, seq[pos_on_query], qual[pos_on_query]);
                                    advance();
                                }
                                pos_on_query++;
                                pos_on_ref++;
                            }
                        } else { /* same visits, skipping straight to the candidate positions */
                            const int64_t op_end = pos_on_ref + len;
                            while (idx < cands.size() && snp_pos >= pos_on_ref && snp_pos < op_end) {
                                const int64_t qpos = pos_on_query + (snp_pos - pos_on_ref);
                                if (qpos >= seq_len) return LCR_ERR_BAD_CIGAR;
                                emit_elem(fragment, (uint32_t)idx, seq[qpos], qual[qpos]);
                                advance();
                            }
                            pos_on_query += len;
                            pos_on_ref = op_end;
                        }
                        break;
                    case 1: pos_on_query += len; break;
                    case 2: case 3: {
                        const int64_t op_end = pos_on_ref + len;
                        while (idx < cands.size() && snp_pos >= pos_on_ref && snp_pos < op_end) advance();
                        pos_on_ref = op_end;
                        break;
                    }
                    default: return LCR_ERR_BAD_CIGAR; /* fragment.rs:190-192 panic */
                }
            }
            /* allele-pair counts (fragment.rs:207-240) */
            for (size_t i = 0; i < fragment.list.size(); ++i)
                for (size_t j = i + 1; j < fragment.list.size(); ++j) {
                    const FragElem &x = fragment.list[i], &y = fragment.list[j];
                    const FragElem &lo = x.snp_idx < y.snp_idx ? x : y, &hi = x.snp_idx < y.snp_idx ? y : x;
                    LdPair &lp = allele_pairs[((uint64_t)lo.snp_idx << 32) | hi.snp_idx];
                    int b1 = base_code(lo.base), b2 = base_code(hi.base);
                    if (b1 >= 0 && b2 >= 0) lp.cnt[b1][b2] += 1;
                }
            uint32_t hete_links = 0;
            for (const FragElem &fe : fragment.list) {
                /* a base of quality 0 has prob = 1.0 (fragment.rs:133): at a phase site log10(0) reaches the first sigma
                   sweep and the reference panics on the NaN compare (phase.rs:307) -> status.  At the other sites (edit /
                   low-fraction candidates) the reference carries on: the IEEE outcomes are restated in eval_rescue and
                   assign_reads_haplotype below (mode 1 simply computes them) */
                if (fe.baseq == 0 && fe.phase_site) return LCR_ERR_BASEQ_ZERO;
                if (fe.phase_site) hete_links++;
            }
            fragment.num_hete_links = hete_links;
            fragment.for_phasing = hete_links >= P.min_linkers;
            const uint32_t fidx = (uint32_t)frags.size();
            for (const FragElem &fe : fragment.list) cands[fe.snp_idx].cover.push_back(fidx);
            if (fragment.for_phasing) out.st.nnz_phase += hete_links;
            frags.push_back(std::move(fragment));
        }
        out.st.n_fragments = frags.size();
        return 0;
    }
    void emit_elem(Fragment &fragment, uint32_t idx, uint8_t base, uint8_t rawq) {
        const Cand &c = cands[idx];
        FragElem fe;
        fe.snp_idx = idx;
        fe.base = base;
        fe.baseq = rawq < 30 ? rawq : 30;
        if (base == c.reference) fe.p = 1;
        else if ((base == c.alleles[0] || base == c.alleles[1]) && base != c.reference) fe.p = -1;
        else fe.p = 0;
        fe.phase_site = c.for_phasing;
        if (!c.dense && fe.p != 0) fragment.list.push_back(fe);
    }

    /* ---- S0: aki (phase.rs:32-49) as log10 terms */
    inline double aki(int sigma, int delta, int eta, int p, int q) const {
        const int x = eta == 0 ? sigma * delta : eta;
        return p == x ? 1.0 - T.fr_prob[q] : T.fr_prob[q];
    }
    inline int64_t aki_fx(int sigma, int delta, int eta, int p, int q) const {
        const int x = eta == 0 ? sigma * delta : eta;
        return p == x ? T.fx_ok[q] : T.fx_err[q];
    }
    /* cal_sigma_delta_eta_log (phase.rs:77-96), f64 */
    double cal_sigma_delta_eta_log(int sigma_k, const std::vector<int> &delta, const std::vector<int> &eta, const std::vector<int> &ps, const std::vector<int> &qs) const {
        double log_q1 = 0, log_q2 = 0, log_q3 = 0;
        for (size_t i = 0; i < delta.size(); ++i) log_q1 += std::log10(aki(sigma_k, delta[i], eta[i], ps[i], qs[i]));
        for (size_t i = 0; i < delta.size(); ++i) {
            log_q2 += std::log10(aki(1, delta[i], eta[i], ps[i], qs[i]));
            log_q3 += std::log10(aki(-1, delta[i], eta[i], ps[i], qs[i]));
        }
        return 1.0 - log_q1 / (log_q2 + log_q3);
    }
    /* cal_delta_eta_sigma_log (phase.rs:128-176), f64 */
    double cal_delta_eta_sigma_log(int delta_i, int eta_i, const std::vector<int> &sigma, const std::vector<int> &ps, const std::vector<int> &qs) const {
        double log_q1 = 0, log_q2 = 0, log_q3 = 0, log_q4 = 0, log_q5 = 0;
        const double prior_homref_log = std::log10(1.0 - 1.5 * 0.001);
        const double prior_homvar_log = std::log10(0.5 * 0.001);
        double prior_hetvar_log;
        const uint32_t coverage_i = (uint32_t)sigma.size();
        if (coverage_i == 0) prior_hetvar_log = std::log10(0.001);
        else prior_hetvar_log = std::log10(0.001) - (double)coverage_i * std::log10(2.0);
        for (size_t k = 0; k < sigma.size(); ++k) log_q1 += std::log10(aki(sigma[k], delta_i, eta_i, ps[k], qs[k]));
        if (eta_i == 0) log_q1 += prior_hetvar_log;
        else if (eta_i == 1) log_q1 += prior_homref_log;
        else log_q1 += prior_homvar_log;
        for (size_t k = 0; k < sigma.size(); ++k) {
            log_q2 += std::log10(aki(sigma[k], delta_i, -1, ps[k], qs[k]));
            log_q3 += std::log10(aki(sigma[k], delta_i, 0, ps[k], qs[k]));
            log_q4 += std::log10(aki(sigma[k], delta_i, 1, ps[k], qs[k]));
            log_q5 += std::log10(aki(sigma[k], delta_i * (-1), 0, ps[k], qs[k]));
        }
        log_q2 += prior_homvar_log;
        log_q3 += prior_hetvar_log;
        log_q4 += prior_homref_log;
        log_q5 += prior_hetvar_log;
        return 1.0 - log_q1 / (log_q2 + log_q3 + log_q4 + log_q5);
    }
    /* cal_phase_score_log (phase.rs:238-255), f64 */
    double cal_phase_score_log(int delta_i, const std::vector<int> &sigma, const std::vector<int> &ps, const std::vector<int> &qs) const {
        double log_q1 = 0, log_q2 = 0, log_q3 = 0;
        for (size_t k = 0; k < sigma.size(); ++k) log_q1 += std::log10(aki(sigma[k], delta_i, 0, ps[k], qs[k]));
        for (size_t k = 0; k < sigma.size(); ++k) {
            log_q2 += std::log10(aki(sigma[k], 1, 0, ps[k], qs[k]));
            log_q3 += std::log10(aki(sigma[k], -1, 0, ps[k], qs[k]));
        }
        return 1.0 - log_q1 / (log_q2 + log_q3);
    }

    /* sums of one SNP column in fixed point */
    struct ColFx {
        int64_t het_d = 0, het_nd = 0, homref = 0, homvar = 0; /* het(delta), het(-delta), eta=1, eta=-1 */
        uint32_t cov = 0;
    };
    inline void col_add(ColFx &c, int sigma, int delta, int p, int q) const {
        c.het_d += aki_fx(sigma, delta, 0, p, q);
        c.het_nd += aki_fx(sigma, -delta, 0, p, q);
        c.homref += aki_fx(sigma, delta, 1, p, q);
        c.homvar += aki_fx(sigma, delta, -1, p, q);
        c.cov++;
    }
    inline int64_t prior_het_fx(uint32_t cov) const { return T.fx_prior_het - (int64_t)cov * T.fx_log10_2; }
    /* the four numerators L1..L4 of q1..q4 (phase.rs:905-908) and the common denominator */
    inline void col_L(const ColFx &c, int64_t L[4], int64_t &D) const {
        const int64_t ph = prior_het_fx(c.cov);
        L[0] = c.het_d + ph;
        L[1] = c.het_nd + ph;
        L[2] = c.homref + T.fx_prior_homref;
        L[3] = c.homvar + T.fx_prior_homvar;
        D = L[3] + L[0] + L[2] + L[1];
    }

    const FragElem *find_elem(const Fragment &f, uint32_t snp) const {
        for (const FragElem &fe : f.list) /* the reference's row scan (phase.rs:890-898) */
            if (fe.snp_idx == snp) return &fe;
        return nullptr;
    }
    /* --downsample: while ds_on, a fragment outside the sampled set is skipped wherever the reference tests
       `apply_downsampling && !downsampled` (phase.rs:262,288,327,369,826,885,983,1021,1326; snpfrags.rs:214,306,414,552) */
    bool ds_on = false;
    inline bool ds_ok(const Fragment &f) const { return !ds_on || f.downsampled; }
    inline bool frag_active(const Fragment &f) const { return f.for_phasing && ds_ok(f) && f.haplotag != 0; }
    /* downsample_fragments (phase.rs:693-701) with the seed of thread.rs:149 */
    void downsample_fragments(uint32_t depth, uint64_t seed) {
        lcr_chacha12 g;
        lcr_stdrng_seed_from_u64(&g, seed);
        std::vector<uint32_t> idx(frags.size());
        for (uint32_t i = 0; i < idx.size(); ++i) idx[i] = i;
        lcr_stdrng_shuffle(&g, idx.data(), (uint32_t)idx.size());
        for (uint32_t i = 0; i < depth && i < idx.size(); ++i) frags[idx[i]].downsampled = true;
    }

    /* cal_overall_probability (phase.rs:257-276) */
    double overall_f64() const {
        double logp = 0;
        for (const Fragment &f : frags) {
            if (!frag_active(f)) continue;
            for (const FragElem &fe : f.list) {
                if (!fe.phase_site) continue;
                logp += std::log10(aki(f.haplotag, cands[fe.snp_idx].haplotype, cands[fe.snp_idx].genotype, fe.p, fe.baseq));
            }
        }
        return logp;
    }
    int64_t overall_fx() const {
        int64_t logp = 0;
        for (const Fragment &f : frags) {
            if (!frag_active(f)) continue;
            for (const FragElem &fe : f.list)
                if (fe.phase_site) logp += aki_fx(f.haplotag, cands[fe.snp_idx].haplotype, cands[fe.snp_idx].genotype, fe.p, fe.baseq);
        }
        return logp;
    }
    struct Prob { /* objective value in either representation */
        double d = -INFINITY;
        int64_t fx = INT64_MIN;
        bool set = false;
        bool better_than(const Prob &o) const {
            if (!o.set) return FX ? true : d > o.d;
            return FX ? fx > o.fx : d > o.d;
        }
    };
    Prob overall() const {
        Prob p;
        p.set = true;
        if (FX) p.fx = overall_fx(); else p.d = overall_f64();
        return p;
    }

    /* ---- S1-S4: cross_optimize (phase.rs:810-976) */
    Prob cross_optimize(const std::vector<char> &conserved, bool keep_conserved, bool with_genotype) {
        bool hg_increase = true, ht_increase = true;
        int num_iters = 0;
        out.st.n_cross_optimize++;
        std::vector<int> tmp_tag(frags.size());
        std::vector<char> has_tag(frags.size());
        std::vector<std::pair<int, int>> tmp_hg(cands.size());
        std::vector<char> has_hg(cands.size());
        std::vector<int> delta, eta, ps, qs, sigma;
        while (hg_increase | ht_increase) {
            out.st.n_sweep_iters++;
            /* sigma sweep (phase.rs:823-855) */
            std::fill(has_tag.begin(), has_tag.end(), 0);
            bool any_read_better = false;
            for (size_t k = 0; k < frags.size(); ++k) {
                const Fragment &f = frags[k];
                if (!frag_active(f)) continue;
                const int sigma_k = f.haplotag;
                if (FX) {
                    int64_t A = 0, Bs = 0;
                    uint32_t cnt = 0;
                    for (const FragElem &fe : f.list) {
                        if (!fe.phase_site) continue;
                        const Cand &c = cands[fe.snp_idx];
                        A += aki_fx(sigma_k, c.haplotype, c.genotype, fe.p, fe.baseq);
                        Bs += aki_fx(-sigma_k, c.haplotype, c.genotype, fe.p, fe.baseq);
                        cnt++;
                    }
                    if (!cnt) continue;
                    has_tag[k] = 1;
                    if (A < Bs) { tmp_tag[k] = -sigma_k; any_read_better = true; }
                    else tmp_tag[k] = sigma_k;
                } else {
                    delta.clear(); eta.clear(); ps.clear(); qs.clear();
                    for (const FragElem &fe : f.list) {
                        if (!fe.phase_site) continue;
                        ps.push_back(fe.p); qs.push_back(fe.baseq);
                        delta.push_back(cands[fe.snp_idx].haplotype);
                        eta.push_back(cands[fe.snp_idx].genotype);
                    }
                    if (delta.empty()) continue;
                    const double q = cal_sigma_delta_eta_log(sigma_k, delta, eta, ps, qs);
                    const double qn = cal_sigma_delta_eta_log(-sigma_k, delta, eta, ps, qs);
                    has_tag[k] = 1;
                    tmp_tag[k] = q < qn ? -sigma_k : sigma_k;
                }
            }
            int check_val;
            if (FX) check_val = any_read_better ? 1 : 0;
            else { /* check_new_haplotag (phase.rs:278-314), map visited in index order */
                double logp = 0, pre_logp = 0;
                for (size_t k = 0; k < frags.size(); ++k) {
                    if (!has_tag[k] || frags[k].haplotag == 0) continue;
                    delta.clear(); eta.clear(); ps.clear(); qs.clear();
                    for (const FragElem &fe : frags[k].list) {
                        if (!fe.phase_site) continue;
                        ps.push_back(fe.p); qs.push_back(fe.baseq);
                        delta.push_back(cands[fe.snp_idx].haplotype);
                        eta.push_back(cands[fe.snp_idx].genotype);
                    }
                    if (delta.empty()) continue;
                    logp += cal_sigma_delta_eta_log(tmp_tag[k], delta, eta, ps, qs);
                    pre_logp += cal_sigma_delta_eta_log(frags[k].haplotag, delta, eta, ps, qs);
                }
                check_val = logp > pre_logp ? 1 : (logp == pre_logp ? 0 : -1);
                if (check_val < 0) check_val = 0; /* reference asserts; rounding-level decrease counts as no increase */
            }
            for (size_t k = 0; k < frags.size(); ++k)
                if (has_tag[k]) frags[k].haplotag = tmp_tag[k];
            if (check_val == 0) ht_increase = false;
            else { ht_increase = true; hg_increase = true; }
            if (!FX) { /* check_local_optimal_configuration (phase.rs:978-1006): evaluated, asserts dropped */
                volatile double sink = 0;
                for (size_t k = 0; k < frags.size(); ++k) {
                    const Fragment &f = frags[k];
                    if (!frag_active(f)) continue;
                    delta.clear(); eta.clear(); ps.clear(); qs.clear();
                    for (const FragElem &fe : f.list) {
                        if (!fe.phase_site) continue;
                        ps.push_back(fe.p); qs.push_back(fe.baseq);
                        delta.push_back(cands[fe.snp_idx].haplotype);
                        eta.push_back(cands[fe.snp_idx].genotype);
                    }
                    if (delta.empty()) continue;
                    sink = sink + cal_sigma_delta_eta_log(f.haplotag, delta, eta, ps, qs) - cal_sigma_delta_eta_log(-f.haplotag, delta, eta, ps, qs);
                }
            }
            /* delta / eta sweep (phase.rs:872-940) */
            std::fill(has_hg.begin(), has_hg.end(), 0);
            bool any_snp_better = false;
            for (size_t i = 0; i < cands.size(); ++i) {
                const Cand &c = cands[i];
                if (!c.for_phasing) continue;
                if (keep_conserved && conserved[i]) continue;
                const int delta_i = c.haplotype, eta_i = c.genotype;
                if (FX) {
                    ColFx col;
                    for (uint32_t k : c.cover) {
                        const Fragment &f = frags[k];
                        if (!frag_active(f)) continue;
                        const FragElem *fe = find_elem(f, (uint32_t)i);
                        if (!fe || !fe->phase_site) continue;
                        col_add(col, f.haplotag, delta_i, fe->p, fe->baseq);
                    }
                    if (!col.cov) continue;
                    int64_t L[4], D;
                    col_L(col, L, D);
                    const int64_t L_old = eta_i == 0 ? L[0] : (eta_i == 1 ? L[2] : L[3]);
                    int64_t L_new;
                    if (with_genotype) {
                        const int64_t mx = std::max(std::max(L[0], L[1]), std::max(L[2], L[3]));
                        if (L[0] == mx) { tmp_hg[i] = {delta_i, 0}; L_new = L[0]; }
                        else if (L[1] == mx) { tmp_hg[i] = {-delta_i, 0}; L_new = L[1]; }
                        else if (L[2] == mx) { tmp_hg[i] = {delta_i, 1}; L_new = L[2]; }
                        else { tmp_hg[i] = {delta_i, -1}; L_new = L[3]; }
                    } else if (eta_i == 0) {
                        if (L[0] >= L[1]) { tmp_hg[i] = {delta_i, 0}; L_new = L[0]; }
                        else { tmp_hg[i] = {-delta_i, 0}; L_new = L[1]; }
                    } else {
                        if (L[2] >= L[3]) { tmp_hg[i] = {delta_i, 1}; L_new = L[2]; }
                        else { tmp_hg[i] = {delta_i, -1}; L_new = L[3]; }
                    }
                    has_hg[i] = 1;
                    if (L_new > L_old) any_snp_better = true;
                } else {
                    sigma.clear(); ps.clear(); qs.clear();
                    for (uint32_t k : c.cover) {
                        const Fragment &f = frags[k];
                        if (!frag_active(f)) continue;
                        for (const FragElem &fe : f.list)
                            if (fe.snp_idx == i) {
                                if (!fe.phase_site) continue;
                                ps.push_back(fe.p); qs.push_back(fe.baseq); sigma.push_back(f.haplotag);
                            }
                    }
                    if (sigma.empty()) continue;
                    const double q1 = cal_delta_eta_sigma_log(delta_i, 0, sigma, ps, qs);
                    const double q2 = cal_delta_eta_sigma_log(-delta_i, 0, sigma, ps, qs);
                    const double q3 = cal_delta_eta_sigma_log(delta_i, 1, sigma, ps, qs);
                    const double q4 = cal_delta_eta_sigma_log(delta_i, -1, sigma, ps, qs);
                    if (with_genotype) {
                        const double mx = fmax(q1, fmax(q2, fmax(q3, q4)));
                        has_hg[i] = 1;
                        if (q1 == mx) tmp_hg[i] = {delta_i, 0};
                        else if (q2 == mx) tmp_hg[i] = {-delta_i, 0};
                        else if (q3 == mx) tmp_hg[i] = {delta_i, 1};
                        else tmp_hg[i] = {delta_i, -1};
                    } else if (eta_i == 0) {
                        const double mx = fmax(q1, q2);
                        has_hg[i] = 1;
                        if (q1 == mx) tmp_hg[i] = {delta_i, 0}; else tmp_hg[i] = {-delta_i, 0};
                    } else {
                        const double mx = fmax(q3, q4);
                        has_hg[i] = 1;
                        if (q3 == mx) tmp_hg[i] = {delta_i, 1}; else tmp_hg[i] = {delta_i, -1};
                    }
                }
            }
            if (FX) check_val = any_snp_better ? 1 : 0;
            else { /* check_new_haplotype_genotype (phase.rs:316-355) */
                double logp = 0, pre_logp = 0;
                for (size_t i = 0; i < cands.size(); ++i) {
                    if (!has_hg[i]) continue;
                    sigma.clear(); ps.clear(); qs.clear();
                    for (uint32_t k : cands[i].cover) {
                        const Fragment &f = frags[k];
                        if (!frag_active(f)) continue;
                        for (const FragElem &fe : f.list) {
                            if (fe.snp_idx != i || !fe.phase_site) continue;
                            ps.push_back(fe.p); qs.push_back(fe.baseq); sigma.push_back(f.haplotag);
                        }
                    }
                    if (sigma.empty()) continue;
                    logp += cal_delta_eta_sigma_log(tmp_hg[i].first, tmp_hg[i].second, sigma, ps, qs);
                    pre_logp += cal_delta_eta_sigma_log(cands[i].haplotype, cands[i].genotype, sigma, ps, qs);
                }
                check_val = logp > pre_logp ? 1 : 0;
            }
            for (size_t i = 0; i < cands.size(); ++i)
                if (has_hg[i]) { cands[i].haplotype = tmp_hg[i].first; cands[i].genotype = tmp_hg[i].second; }
            if (check_val == 0) hg_increase = false;
            else { hg_increase = true; ht_increase = true; }
            num_iters++;
            if (num_iters > 20) break;
        }
        return overall();
    }

    /* ---- best configuration bookkeeping (phase.rs:1064-1085) */
    struct Config {
        std::vector<int> haplotype, genotype, haplotag;
    };
    void save_best(Config &c) const {
        c.haplotype.resize(cands.size()); c.genotype.resize(cands.size()); c.haplotag.resize(frags.size());
        for (size_t i = 0; i < cands.size(); ++i) { c.haplotype[i] = cands[i].haplotype; c.genotype[i] = cands[i].genotype; }
        for (size_t k = 0; k < frags.size(); ++k) c.haplotag[k] = frags[k].haplotag;
    }
    void load_best(const Config &c) {
        for (size_t i = 0; i < cands.size(); ++i) { cands[i].haplotype = c.haplotype[i]; cands[i].genotype = c.genotype[i]; }
        for (size_t k = 0; k < frags.size(); ++k) frags[k].haplotag = c.haplotag[k];
    }
    double uniform(uint32_t stream, uint32_t call, uint32_t idx) const { return lcr_uniform(P.seed, region_key, stream, call, idx); }
    void init_assignment(uint32_t call) { /* phase.rs:673-680 */
        for (Fragment &f : frags)
            if (f.for_phasing) f.haplotag = uniform(LCR_RNG_INIT_SIGMA, call, f.read - reg.read_begin) < 0.5 ? -1 : 1;
    }
    void init_genotype() { /* phase.rs:682-691 */
        for (Cand &c : cands) c.genotype = c.variant_type == 0 ? 1 : (c.variant_type == 1 ? 0 : ((c.variant_type == 2 || c.variant_type == 3) ? -1 : c.genotype));
    }

    /* ---- L1: divide_snps_into_blocks (candidate.rs:615-747), LD_Pair::calculate_ld (snp.rs:158-188) */
    Graph divide_snps_into_blocks() {
        std::vector<uint32_t> ld_idxes;
        for (size_t i = 0; i < cands.size(); ++i) if (cands[i].for_phasing) ld_idxes.push_back((uint32_t)i);
        Graph g;
        auto ref_alt = [&](const Cand &s, uint8_t &r, uint8_t &a, float &rf, float &af) -> bool {
            if (s.alleles[0] == s.reference && s.alleles[1] != s.reference) { r = s.alleles[0]; rf = s.allele_freqs[0]; a = s.alleles[1]; af = s.allele_freqs[1]; return true; }
            if (s.alleles[0] != s.reference && s.alleles[1] == s.reference) { r = s.alleles[1]; rf = s.allele_freqs[1]; a = s.alleles[0]; af = s.allele_freqs[0]; return true; }
            return false;
        };
        auto eval_pair = [&](uint32_t idx1, uint32_t idx2, LdPair &lp) {
            uint8_t r1, a1, r2, a2; float rf1, af1, rf2, af2;
            if (!ref_alt(cands[idx1], r1, a1, rf1, af1)) return;
            if (!ref_alt(cands[idx2], r2, a2, rf2, af2)) return;
            if (rf1 == 0.0f || af1 == 0.0f || rf2 == 0.0f || af2 == 0.0f) return;
            const int R1 = base_code(r1), A1 = base_code(a1), R2 = base_code(r2), A2 = base_code(a2);
            int count[4] = {(int)lp.cnt[R1][R2], (int)lp.cnt[R1][A2], (int)lp.cnt[A1][R2], (int)lp.cnt[A1][A2]};
            const int c1 = std::min(count[0] + count[3], count[1] + count[2]);
            const int c2 = std::max(count[0] + count[3], count[1] + count[2]);
            const float score = (float)c1 / (float)c2;
            if (count[0] + count[3] > count[1] + count[2]) { lp.score = score; lp.weight = c2; }
            else { lp.score = -1.0f * score; lp.weight = -c2; }
            lp.valid = true;
            if (lp.score == 0.0f && (uint32_t)std::abs(lp.weight) >= P.ld_weight_threshold) {
                if (!g.has_edge(idx1, idx2)) g.add_edge(idx1, idx2);
            }
        };
        if (!FX) {
            for (size_t i = 0; i < ld_idxes.size(); ++i)
                for (size_t j = i + 1; j < ld_idxes.size(); ++j) {
                    auto it = allele_pairs.find(((uint64_t)ld_idxes[i] << 32) | ld_idxes[j]);
                    if (it == allele_pairs.end()) continue;
                    eval_pair(ld_idxes[i], ld_idxes[j], it->second);
                }
        } else { /* same pairs in the same (i, j) order, found from the sparse side */
            std::vector<uint64_t> keys;
            keys.reserve(allele_pairs.size());
            for (auto &kv : allele_pairs) keys.push_back(kv.first);
            std::sort(keys.begin(), keys.end());
            for (uint64_t k : keys) {
                uint32_t i = (uint32_t)(k >> 32), j = (uint32_t)k;
                if (!cands[i].for_phasing || !cands[j].for_phasing) continue;
                eval_pair(i, j, allele_pairs[k]);
            }
        }
        /* note: an edge whose |weight| < ld_weight_threshold is added and then removed in the
           reference (candidate.rs:705-713); its nodes stay in the graph.  With the hard-coded
           threshold 1 (thread.rs:166) and |weight| = max(cis, trans) >= 1 this never happens. */
        ld_blocks = kosaraju_scc(g);
        return g;
    }

    /* ---- init_haplotypes_LD2 (phase.rs:600-652) */
    std::vector<char> init_haplotypes_ld2(const Graph &g) {
        for (size_t i = 0; i < cands.size(); ++i) cands[i].haplotype = uniform(LCR_RNG_INIT_DELTA, 0, (uint32_t)i) < 0.5 ? 1 : -1;
        std::vector<char> conserved(cands.size(), 0);
        for (const auto &block : ld_blocks) {
            if (block.size() < 2) continue;
            std::vector<char> disc(g.nodes.size(), 0);
            std::deque<uint32_t> queue;
            std::vector<uint32_t> visited_nodes;
            disc[g.slot.at(block[0])] = 1;
            queue.push_back(block[0]);
            cands[block[0]].haplotype = 1;
            visited_nodes.push_back(block[0]);
            while (!queue.empty()) {
                const uint32_t nx = queue.front();
                queue.pop_front();
                for (uint32_t succ : g.adj[g.slot.at(nx)]) {
                    uint32_t s = g.slot.at(succ);
                    if (!disc[s]) { disc[s] = 1; queue.push_back(succ); }
                }
                for (uint32_t v : visited_nodes) {
                    if (v == nx) continue;
                    const uint32_t from = std::min(v, nx), to = std::max(v, nx);
                    auto it = allele_pairs.find(((uint64_t)from << 32) | to);
                    if (it == allele_pairs.end()) continue;
                    const LdPair &lp = it->second;
                    if (!lp.valid || lp.score != 0.0f) continue;
                    if (lp.weight >= (int)P.ld_weight_threshold) { cands[nx].haplotype = cands[v].haplotype; break; }
                    else if (lp.weight <= -(int)P.ld_weight_threshold) { cands[nx].haplotype = -cands[v].haplotype; break; }
                }
                visited_nodes.push_back(nx);
            }
            for (uint32_t idx : block) conserved[idx] = 1;
        }
        return conserved;
    }

    /* ---- S6: cross_optimize_by_block (phase.rs:1298-1394) */
    Prob cross_optimize_by_block() {
        std::vector<int> tmp_haplotype(cands.size());
        std::vector<char> has_haplotype(cands.size(), 0);
        std::vector<int> tmp_haplotag(frags.size());
        bool has_haplotag = false;
        std::vector<int> sigma, sigma_flip, ps, qs;
        for (const auto &block : ld_blocks) {
            std::unordered_set<uint32_t> block_set(block.begin(), block.end());
            std::unordered_map<uint32_t, int> flip_map;
            double q = 0, q_flip = 0;
            int64_t qfx = 0, qfx_flip = 0; /* contract: sums of round(term * 2^40), order-independent */
            for (uint32_t idx : block) {
                const int d = cands[idx].haplotype, e = cands[idx].genotype;
                sigma.clear(); sigma_flip.clear(); ps.clear(); qs.clear();
                for (uint32_t k : cands[idx].cover) {
                    const Fragment &f = frags[k];
                    if (!frag_active(f)) continue;
                    bool flip_read = true;
                    for (const FragElem &fe : f.list) {
                        if (!block_set.count(fe.snp_idx)) flip_read = false;
                        if (fe.snp_idx == idx) {
                            if (!fe.phase_site) continue;
                            ps.push_back(fe.p); qs.push_back(fe.baseq);
                            const int sf = flip_read ? -f.haplotag : f.haplotag;
                            sigma_flip.push_back(sf);
                            flip_map[k] = sf;
                            sigma.push_back(f.haplotag);
                        }
                    }
                }
                /* cal_block_delta_eta_sigma_log (phase.rs:178-236): one term per SNP, summed in block order */
                if (FX) {
                    ColFx c0, c1;
                    for (size_t k = 0; k < sigma.size(); ++k) { col_add(c0, sigma[k], d, ps[k], qs[k]); col_add(c1, sigma_flip[k], -d, ps[k], qs[k]); }
                    auto term = [&](const ColFx &c) {
                        int64_t L[4], D;
                        col_L(c, L, D);
                        const int64_t L1 = e == 0 ? L[0] : (e == 1 ? L[2] : L[3]);
                        return 1.0 - (double)L1 / (double)D;
                    };
                    qfx += (int64_t)llrint(term(c0) * 1099511627776.0);
                    qfx_flip += (int64_t)llrint(term(c1) * 1099511627776.0);
                } else {
                    q += cal_delta_eta_sigma_log(d, e, sigma, ps, qs);
                    q_flip += cal_delta_eta_sigma_log(-d, e, sigma_flip, ps, qs);
                }
            }
            if (FX ? qfx < qfx_flip : q < q_flip) {
                for (uint32_t idx : block) { tmp_haplotype[idx] = -cands[idx].haplotype; has_haplotype[idx] = 1; }
                for (size_t k = 0; k < frags.size(); ++k) {
                    auto it = flip_map.find((uint32_t)k);
                    tmp_haplotag[k] = it != flip_map.end() ? it->second : frags[k].haplotag;
                }
            } else {
                for (uint32_t idx : block) { tmp_haplotype[idx] = cands[idx].haplotype; has_haplotype[idx] = 1; }
                for (size_t k = 0; k < frags.size(); ++k) tmp_haplotag[k] = frags[k].haplotag;
            }
            has_haplotag = true; /* every block rewrites every fragment's entry: the last block wins */
        }
        for (size_t i = 0; i < cands.size(); ++i) if (has_haplotype[i]) cands[i].haplotype = tmp_haplotype[i];
        if (has_haplotag) for (size_t k = 0; k < frags.size(); ++k) frags[k].haplotag = tmp_haplotag[k];
        return overall();
    }

    /* ---- S5: phase (phase.rs:1087-1296) */
    void phase() {
        Prob largest;
        Config best;
        std::vector<char> conserved(cands.size(), 0);
        const size_t n = cands.size();
        if (n <= P.max_enum_snps) {
            if (!FX) divide_snps_into_blocks(); /* result unused on this branch; kept for the reference's cost */
            const uint32_t n_cfg = 1u << n;
            for (uint32_t c = 0; c < n_cfg; ++c) {
                for (size_t i = 0; i < n; ++i) cands[i].haplotype = ((c >> i) & 1) ? -1 : 1;
                init_assignment(c);
                init_genotype();
                Prob prob = cross_optimize(conserved, false, true);
                if (prob.better_than(largest)) { largest = prob; save_best(best); }
            }
            load_best(best);
        } else {
            Graph g = divide_snps_into_blocks();
            conserved = init_haplotypes_ld2(g);
            init_genotype();
            init_assignment(0);
            Prob prob = cross_optimize(conserved, true, false);
            if (prob.better_than(largest)) { largest = prob; save_best(best); }
            load_best(best);
            prob = cross_optimize_by_block();
            if (prob.better_than(largest)) { largest = prob; save_best(best); }
            load_best(best);
            for (uint32_t tidx = 0; tidx <= n / 4; ++tidx) {
                const bool flip = tidx % 2 == 1;
                for (size_t i = 0; i < n; ++i) {
                    const double rg = uniform(LCR_RNG_PERTURB_DELTA, tidx, (uint32_t)i);
                    if (rg < 0.1) cands[i].haplotype = flip ? 1 : -1;
                    else if (rg >= 0.9) cands[i].haplotype = flip ? -1 : 1;
                }
                prob = cross_optimize(conserved, false, false);
                if (prob.better_than(largest)) { largest = prob; save_best(best); }
                load_best(best);
                for (Fragment &f : frags) {
                    if (!f.for_phasing || f.haplotag == 0) continue;
                    if (uniform(LCR_RNG_PERTURB_SIGMA, tidx, f.read - reg.read_begin) < 0.1) f.haplotag *= -1;
                }
                prob = cross_optimize(conserved, false, false);
                if (prob.better_than(largest)) { largest = prob; save_best(best); }
                load_best(best);
            }
            load_best(best);
        }
    }

    /* ---- A1: assign_reads_haplotype (snpfrags.rs:548-625) */
    void assign_reads_haplotype(std::vector<std::pair<uint32_t, int>> *ra) {
        if (ra) ra->clear();
        std::vector<int> delta, eta, ps, qs;
        for (Fragment &f : frags) {
            if (!f.for_phasing || !ds_ok(f)) continue;
            const int sigma_k = f.haplotag;
            delta.clear(); eta.clear(); ps.clear(); qs.clear();
            int64_t A = 0, Bs = 0;
            bool q0_used = false; /* contract: a quality-0 element (only possible at a rescued site) */
            for (FragElem &fe : f.list) {
                const Cand &c = cands[fe.snp_idx];
                if (!fe.phase_site && c.for_phasing) fe.phase_site = true;
                if (!c.for_phasing || c.haplotype == 0 || c.genotype != 0) continue;
                ps.push_back(fe.p); qs.push_back(fe.baseq); delta.push_back(c.haplotype); eta.push_back(c.genotype);
                if (FX) {
                    if (fe.baseq == 0) { q0_used = true; continue; }
                    A += aki_fx(sigma_k, c.haplotype, c.genotype, fe.p, fe.baseq); Bs += aki_fx(-sigma_k, c.haplotype, c.genotype, fe.p, fe.baseq);
                }
            }
            int assign = 0;
            if (sigma_k == 0 || delta.empty()) {
                f.assignment = 0; f.haplotag = 0; f.assignment_score = 0.0;
            } else {
                double q, qn;
                if (FX && q0_used) {
                    /* exactly one of the two sums holds log10(0) = -inf (both when two such elements disagree): one of q, qn
                       is NaN, `(q - qn).abs() >= cutoff` is false and the read goes to the unknown group (snpfrags.rs:611-618) */
                    q = qn = std::nan("");
                } else if (FX) {
                    /* both denominators are the same pair of sums in either order: integer add, one conversion */
                    const double den = lcr_fx_to_f64(A + Bs);
                    q = 1.0 - lcr_fx_to_f64(A) / den;
                    qn = 1.0 - lcr_fx_to_f64(Bs) / den;
                } else {
                    q = cal_sigma_delta_eta_log(sigma_k, delta, eta, ps, qs);
                    qn = cal_sigma_delta_eta_log(-sigma_k, delta, eta, ps, qs);
                }
                if (std::fabs(q - qn) >= P.read_assignment_cutoff) {
                    if (q >= qn) {
                        assign = sigma_k == 1 ? 1 : 2;
                        f.assignment = assign; f.assignment_score = q;
                    } else {
                        assign = sigma_k == 1 ? 2 : 1;
                        f.assignment = assign; f.assignment_score = qn;
                        f.haplotag = sigma_k == 1 ? -1 : 1;
                    }
                } else {
                    f.assignment = 0; f.haplotag = 0; f.assignment_score = 0.0;
                }
            }
            if (ra) ra->emplace_back(f.read, assign);
        }
    }

    /* gather one SNP's column for the assignment / rescue passes */
    struct ColData {
        std::vector<int> sigma, ps, qs;
        int hap1 = 0, hap2 = 0;
    };
    /* phase score from a column (snpfrags.rs:245-246,483): -10 log10(1 - cal_phase_score_log) */
    double phase_score_of(int delta_i, const ColData &cd) const {
        if (FX) {
            /* quality-0 elements: aki(sigma, delta, 0, p, 1.0) is 0 when p == sigma * delta, so log_q2 (delta = +1) or log_q3
               (delta = -1) of phase.rs:238-255 is -inf.  log_q1 / (log_q2 + log_q3) is then NaN when log_q1 is the infinite
               one and +-0 otherwise, i.e. the score is NaN or -10 log10(1 - 1) = +inf */
            bool z1 = false, z2 = false;
            for (size_t k = 0; k < cd.sigma.size(); ++k)
                if (cd.qs[k] == 0) { if (cd.ps[k] == cd.sigma[k]) z1 = true; else z2 = true; }
            if (z1 || z2) {
                const bool mine_inf = delta_i == 1 ? z1 : z2;
                return mine_inf ? std::nan("") : HUGE_VAL;
            }
            int64_t L1 = 0, L2 = 0, L3 = 0;
            for (size_t k = 0; k < cd.sigma.size(); ++k) {
                L1 += aki_fx(cd.sigma[k], delta_i, 0, cd.ps[k], cd.qs[k]);
                L2 += aki_fx(cd.sigma[k], 1, 0, cd.ps[k], cd.qs[k]);
                L3 += aki_fx(cd.sigma[k], -1, 0, cd.ps[k], cd.qs[k]);
            }
            const double v = 1.0 - lcr_fx_to_f64(L1) / lcr_fx_to_f64(L2 + L3);
            return -10.0 * lcr_log10(1.0 - v);
        }
        return -10.0 * std::log10(1.0 - cal_phase_score_log(delta_i, cd.sigma, cd.ps, cd.qs));
    }

    /* ---- A3: eval_rna_edit_var_phase / eval_low_frac_var_phase (snpfrags.rs:191-376) */
    void eval_rescue(const std::vector<uint32_t> &list, bool low_frac, float min_phase_score) {
        for (uint32_t ti : list) {
            Cand &snp = cands[ti];
            if (snp.cover.empty()) { snp.single = true; continue; }
            if (snp.variant_type != 1) { snp.non_selected = true; continue; }
            ColData cd;
            for (uint32_t k : snp.cover) {
                const Fragment &f = frags[k];
                if (!f.for_phasing || f.assignment == 0 || f.num_hete_links < P.min_linkers) continue;
                if (!ds_ok(f)) continue;
                for (const FragElem &fe : f.list)
                    if (fe.snp_idx == ti) {
                        if (f.assignment == 1) cd.hap1++; else if (f.assignment == 2) cd.hap2++;
                        cd.ps.push_back(fe.p); cd.qs.push_back(fe.baseq); cd.sigma.push_back(f.haplotag);
                    }
            }
            if (cd.sigma.empty() || cd.hap1 < 2 || cd.hap2 < 2) { snp.single = true; continue; }
            const double s1 = phase_score_of(1, cd), s2 = phase_score_of(-1, cd);
            snp.single = false;
            if (fmax(s1, s2) >= (double)min_phase_score) {
                snp.non_selected = false;
                if (low_frac) snp.cand_somatic = false;
                snp.rna_editing = false;
                snp.for_phasing = true;
                for (uint32_t k : snp.cover) {
                    Fragment &f = frags[k];
                    f.for_phasing = true;
                    if (f.haplotag == 0 || f.assignment == 0)
                        f.haplotag = uniform(LCR_RNG_RESCUE_SIGMA, ti, f.read - reg.read_begin) < 0.5 ? -1 : 1;
                }
                snp.haplotype = s1 >= s2 ? 1 : -1;
                snp.genotype = 0;
                snp.variant_type = 1;
                snp.phase_score = fmax(s1, s2);
            } else {
                snp.non_selected = true;
                if (low_frac) { snp.cand_somatic = true; snp.for_phasing = false; }
                else snp.rna_editing = true;
            }
        }
    }

    /* ---- A2: assign_snp_haplotype_genotype (snpfrags.rs:378-546) */
    void assign_snp_haplotype_genotype() {
        for (size_t ti = 0; ti < cands.size(); ++ti) {
            Cand &snp = cands[ti];
            if (!snp.for_phasing) { snp.non_selected = true; continue; }
            if (snp.cover.empty()) { snp.single = true; continue; }
            const int delta_i = snp.haplotype;
            ColData cd;
            for (uint32_t k : snp.cover) {
                const Fragment &f = frags[k];
                if (!f.for_phasing || f.num_hete_links < P.min_linkers) continue;
                if (!ds_ok(f)) continue;
                if (snp.variant_type == 1 && f.assignment == 0) continue;
                for (const FragElem &fe : f.list)
                    if (fe.snp_idx == ti) {
                        if (f.assignment == 1) cd.hap1++; else if (f.assignment == 2) cd.hap2++;
                        cd.ps.push_back(fe.p); cd.qs.push_back(fe.baseq); cd.sigma.push_back(f.haplotag);
                    }
            }
            if (cd.sigma.empty()) { snp.non_selected = true; continue; }
            int pick;
            if (FX) {
                ColFx col;
                for (size_t k = 0; k < cd.sigma.size(); ++k) col_add(col, cd.sigma[k], delta_i, cd.ps[k], cd.qs[k]);
                int64_t L[4], D;
                col_L(col, L, D);
                const int64_t mx = std::max(std::max(L[0], L[1]), std::max(L[2], L[3]));
                pick = L[0] == mx ? 0 : (L[1] == mx ? 1 : (L[2] == mx ? 2 : 3));
            } else {
                const double q1 = cal_delta_eta_sigma_log(delta_i, 0, cd.sigma, cd.ps, cd.qs);
                const double q2 = cal_delta_eta_sigma_log(-delta_i, 0, cd.sigma, cd.ps, cd.qs);
                const double q3 = cal_delta_eta_sigma_log(delta_i, 1, cd.sigma, cd.ps, cd.qs);
                const double q4 = cal_delta_eta_sigma_log(delta_i, -1, cd.sigma, cd.ps, cd.qs);
                const double mx = fmax(q1, fmax(q2, fmax(q3, q4)));
                pick = q1 == mx ? 0 : (q2 == mx ? 1 : (q3 == mx ? 2 : 3));
            }
            if (pick == 0) { snp.haplotype = delta_i; snp.genotype = 0; snp.variant_type = 1; }
            else if (pick == 1) { snp.haplotype = -delta_i; snp.genotype = 0; snp.variant_type = 1; }
            else if (pick == 2) { snp.haplotype = delta_i; snp.genotype = 1; snp.variant_type = 0; }
            else { snp.haplotype = delta_i; snp.genotype = -1; if (snp.variant_type != 2 && snp.variant_type != 3) snp.variant_type = 2; }
            if (snp.genotype != 0) { snp.non_selected = true; continue; }
            if (cd.hap1 >= 1 && cd.hap2 >= 1) snp.phase_score = phase_score_of(snp.haplotype, cd);
            else snp.phase_score = 0.19940219;
        }
    }

    /* ---- A4: assign_phase_set (snpfrags.rs:628-733) */
    void assign_phase_set(std::vector<std::pair<uint32_t, uint32_t>> &phase_set) {
        const double mps = (double)P.min_phase_score;
        std::vector<char> is_node(cands.size(), 0);
        std::vector<uint32_t> parent(cands.size());
        for (size_t i = 0; i < cands.size(); ++i) {
            parent[i] = (uint32_t)i;
            const Cand &s = cands[i];
            if (s.genotype != 0 || s.variant_type != 1) continue;
            if (s.dense || s.rna_editing) continue;
            if (s.phase_score < mps) continue;
            is_node[i] = 1;
        }
        /* connected components with the lowest index as representative (== first node petgraph emits) */
        auto find = [&](uint32_t x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
        auto unite = [&](uint32_t a, uint32_t b) { a = find(a); b = find(b); if (a == b) return; if (a < b) parent[b] = a; else parent[a] = b; };
        struct Touch { uint32_t frag, node; };
        std::vector<Touch> touches; /* (fragment, one endpoint of an edge that fragment sits on) */
        for (size_t k = 0; k < frags.size(); ++k) {
            const Fragment &f = frags[k];
            if (!f.for_phasing || f.assignment == 0) continue;
            std::vector<const FragElem *> nodes;
            for (const FragElem &fe : f.list) if (is_node[fe.snp_idx]) nodes.push_back(&fe);
            if (nodes.size() == 1) touches.push_back({(uint32_t)k, nodes[0]->snp_idx});
            if (nodes.size() >= 2)
                for (size_t j0 = 0; j0 < nodes.size(); ++j0)
                    for (size_t j1 = 0; j1 < nodes.size(); ++j1) {
                        if (j0 == j1) continue;
                        const int hp = cands[nodes[j0]->snp_idx].haplotype * cands[nodes[j1]->snp_idx].haplotype;
                        if (hp != nodes[j0]->p * nodes[j1]->p) continue;
                        unite(nodes[j0]->snp_idx, nodes[j1]->snp_idx);
                        touches.push_back({(uint32_t)k, nodes[j0]->snp_idx});
                    }
        }
        for (size_t i = 0; i < cands.size(); ++i)
            if (is_node[i]) cands[i].phase_set = (uint32_t)(cands[find((uint32_t)i)].pos + 1);
        /* components are visited in descending order of their first node; a read keeps the first id it meets */
        std::vector<uint32_t> best_root(frags.size(), UINT32_MAX);
        std::vector<char> has(frags.size(), 0);
        for (const Touch &t : touches) {
            const uint32_t root = find(t.node);
            if (!has[t.frag] || root > best_root[t.frag]) { has[t.frag] = 1; best_root[t.frag] = root; }
        }
        for (size_t k = 0; k < frags.size(); ++k)
            if (has[k]) phase_set.emplace_back(frags[k].read, (uint32_t)(cands[best_root[k]].pos + 1));
    }

    /* ---- orchestration: thread.rs:78-221 */
    void run() {
        std::vector<BaseFreq> pile;
        int st = pileup(pile);
        out.st.n_positions = (uint64_t)(reg.end - reg.start);
        if (st) { /* the reference panics here; the contract reports the status, no candidates, zeroed planes */
            out.status = st;
            if (P.flags & LCR_FLAG_EMIT_PLANES) out.pileup.assign((size_t)(reg.end - reg.start), BaseFreq());
            return;
        }
        if (B.ext_off) import_external_candidates(pile);
        else get_candidate_snps(pile);
        if (P.flags & LCR_FLAG_EMIT_PLANES) out.pileup = std::move(pile);
        else std::vector<BaseFreq>().swap(pile);
        out.st.n_candidates = cands.size();
        if (P.flags & LCR_FLAG_SKIP_PHASING) return;
        st = get_fragments();
        if (st) { out.status = st; frags.clear(); for (Cand &c : cands) c.cover.clear(); return; }
        /* thread.rs:144-151 */
        const bool apply_downsampling = (P.flags & LCR_FLAG_DOWNSAMPLE) && P.downsample_depth > 0 && frags.size() >= (size_t)P.downsample_depth;
        if (apply_downsampling) downsample_fragments(P.downsample_depth, 2025);
        ds_on = apply_downsampling;
        phase();
        if (P.flags & LCR_ORACLE_FLAG_STOP_AFTER_PHASE) { /* test hook: the state phase() leaves (haplotags as HP 1 / 2) */
            for (const Fragment &f : frags) out.hp.emplace_back(f.read, f.haplotag == 1 ? 1 : (f.haplotag == -1 ? 2 : 0));
            return;
        }
        assign_reads_haplotype(nullptr);
        assign_snp_haplotype_genotype();
        assign_reads_haplotype(nullptr);
        assign_snp_haplotype_genotype();
        eval_rescue(edit_snps, false, P.min_phase_score - 3.0f);
        eval_rescue(somatic_snps, true, P.min_phase_score - 3.0f);
        ds_on = false; /* thread.rs:181-182: the last round runs over every fragment */
        assign_reads_haplotype(&out.hp);
        assign_snp_haplotype_genotype();
        assign_phase_set(out.ps);
    }
};

/* ------------------------------------------------------------ packing ---- */

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

} // namespace

extern "C" {

/* mode: 0 = contract (fixed point), 1 = reference-order f64.  ref_seqs[tid] / ref_lens[tid]
   give the contigs.  n_threads regions are processed concurrently (thread.rs:52-77). */
int lcr_oracle_run(const lcr_params *params, const lcr_batch *batch, const uint8_t *const *ref_seqs, const uint64_t *ref_lens,
                   int n_tids, int mode, int n_threads, lcr_result **out) {
    if (!params || !batch || !out) return LCR_ERR_INVALID_ARG;
    lcr_luts T;
    lcr_build_luts(&T);
    std::vector<RegionOut> routs(batch->n_regions);
    std::atomic<uint32_t> next{0};
    auto work = [&]() {
        for (;;) {
            uint32_t r = next.fetch_add(1);
            if (r >= batch->n_regions) break;
            const lcr_region &reg = batch->regions[r];
            RegionOut &ro = routs[r];
            if (reg.tid < 0 || reg.tid >= n_tids || !ref_seqs[reg.tid]) { ro.status = LCR_ERR_NO_REFERENCE; continue; }
            if (reg.end < reg.start || reg.start < 1 || reg.read_end < reg.read_begin || reg.read_end > batch->n_reads) { ro.status = LCR_ERR_INVALID_ARG; continue; }
            if (batch->exon_off && batch->exon_off[r + 1] == batch->exon_off[r]) { ro.status = LCR_REGION_NO_EXON; continue; } /* thread.rs:88-91: nothing is done for the region */
            if (mode == 0) { Worker<true> w(*params, *batch, T, reg, ref_seqs[reg.tid], ref_lens[reg.tid], ro); w.run(); }
            else { Worker<false> w(*params, *batch, T, reg, ref_seqs[reg.tid], ref_lens[reg.tid], ro); w.run(); }
        }
    };
    if (n_threads <= 1) work();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < n_threads; ++i) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
    ResultBox *box = new ResultBox();
    const uint32_t nr = batch->n_regions;
    box->cand_off.assign(nr + 1, 0);
    box->region_status.assign(nr, 0);
    box->hp.assign(batch->n_reads, -1);
    box->ps.assign(batch->n_reads, 0);
    box->is_fragment.assign(batch->n_reads, 0);
    const bool planes = params->flags & LCR_FLAG_EMIT_PLANES, emit_frag = params->flags & LCR_FLAG_EMIT_FRAGMENTS;
    box->pos_off.assign(nr + 1, 0);
    box->frag_off.assign(nr + 1, 0);
    box->elem_off.push_back(0);
    lcr_stats st{};
    for (uint32_t r = 0; r < nr; ++r) {
        RegionOut &ro = routs[r];
        box->region_status[r] = ro.status;
        for (const Cand &c : ro.cands) {
            lcr_candidate o{};
            o.pos = c.pos;
            o.variant_quality = c.variant_quality;
            o.genotype_quality = c.genotype_quality;
            o.phase_score = c.phase_score;
            memcpy(o.genotype_probability, c.genotype_probability, sizeof o.genotype_probability);
            o.allele_freqs[0] = c.allele_freqs[0]; o.allele_freqs[1] = c.allele_freqs[1];
            o.depth = c.depth;
            o.phase_set = c.phase_set;
            o.reference = c.reference;
            o.alleles[0] = c.alleles[0]; o.alleles[1] = c.alleles[1];
            o.variant_type = (int8_t)c.variant_type;
            o.genotype = (int8_t)c.genotype;
            o.haplotype = (int8_t)c.haplotype;
            o.flags = (c.rna_editing ? LCR_CF_RNA_EDITING : 0) | (c.dense ? LCR_CF_DENSE : 0) | (c.het_var ? LCR_CF_HET_VAR : 0) |
                      (c.for_phasing ? LCR_CF_FOR_PHASING : 0) | (c.hom_var ? LCR_CF_HOM_VAR : 0) | (c.single ? LCR_CF_SINGLE : 0) |
                      (c.non_selected ? LCR_CF_NON_SELECTED : 0) | (c.cand_somatic ? LCR_CF_CAND_SOMATIC : 0) |
                      (c.in_edit ? LCR_CF_EDIT_LIST : 0) | (c.in_somatic ? LCR_CF_SOMATIC_LIST : 0);
            o.region = r;
            box->cand.push_back(o);
        }
        box->cand_off[r + 1] = (uint32_t)box->cand.size();
        /* a read shared by several regions keeps the entry of the lowest region (thread.rs:308-325 keeps the first entry
           per qname in queue order, which is completion order there; the contract fixes it to region order) */
        for (auto &h : ro.hp) if (box->hp[h.first] == -1) box->hp[h.first] = (int8_t)h.second;
        for (auto &p : ro.ps) if (box->ps[p.first] == 0) box->ps[p.first] = p.second;
        for (const Fragment &f : ro.frags) box->is_fragment[f.read] = 1;
        if (planes) {
            for (const BaseFreq &bf : ro.pileup) {
                box->acgt.insert(box->acgt.end(), {bf.a, bf.c, bf.g, bf.t});
                box->fwd.insert(box->fwd.end(), {(uint32_t)bf.strands[0][0], (uint32_t)bf.strands[1][0], (uint32_t)bf.strands[2][0], (uint32_t)bf.strands[3][0]});
                box->d.push_back(bf.d);
                box->n.push_back(bf.n);
                box->ts.insert(box->ts.end(), {(uint32_t)bf.ts[0], (uint32_t)bf.ts[1]});
            }
            box->pos_off[r + 1] = box->d.size();
        }
        if (emit_frag) {
            for (const Fragment &f : ro.frags) {
                box->frag_read.push_back(f.read);
                for (const FragElem &fe : f.list) {
                    box->elem_snp.push_back(fe.snp_idx);
                    box->elem_cell.push_back((int8_t)(fe.p * (fe.baseq + 1)));
                    box->elem_base.push_back(fe.base);
                }
                box->elem_off.push_back(box->elem_snp.size());
            }
            box->frag_off[r + 1] = (uint32_t)box->frag_read.size();
        }
        st.n_reads_pass += ro.st.n_reads_pass; st.n_aligned_bases += ro.st.n_aligned_bases; st.n_positions += ro.st.n_positions;
        st.n_candidates += ro.st.n_candidates; st.n_fragments += ro.st.n_fragments; st.nnz_phase += ro.st.nnz_phase;
        st.n_cross_optimize += ro.st.n_cross_optimize; st.n_sweep_iters += ro.st.n_sweep_iters;
        RegionOut().pileup.swap(ro.pileup);
    }
    lcr_result &res = box->res;
    res.n_regions = nr;
    res.n_reads = batch->n_reads;
    res.n_cand = (uint32_t)box->cand.size();
    res.cand_off = box->cand_off.data();
    res.cand = box->cand.data();
    res.region_status = box->region_status.data();
    res.hp = box->hp.data();
    res.ps = box->ps.data();
    res.is_fragment = box->is_fragment.data();
    if (planes) {
        res.planes.n_pos = box->d.size();
        res.planes.pos_off = box->pos_off.data();
        res.planes.acgt = box->acgt.data();
        res.planes.fwd = box->fwd.data();
        res.planes.d = box->d.data();
        res.planes.n = box->n.data();
        res.planes.ts = box->ts.data();
    }
    if (emit_frag) {
        res.fragments.n_frag = box->frag_read.size();
        res.fragments.n_elem = box->elem_snp.size();
        res.fragments.frag_off = box->frag_off.data();
        res.fragments.frag_read = box->frag_read.data();
        res.fragments.elem_off = box->elem_off.data();
        res.fragments.elem_snp = box->elem_snp.data();
        res.fragments.elem_cell = box->elem_cell.data();
        res.fragments.elem_base = box->elem_base.data();
    }
    res.stats = st;
    *out = &box->res;
    return 0;
}

void lcr_oracle_free(lcr_result *res) {
    if (res) delete reinterpret_cast<ResultBox *>(res);
}

/* known-answer hooks for tests/test_oracle_kat.py */
double lcr_oracle_log10(double x) { return lcr_log10(x); }
double lcr_oracle_exp10(double x) { return lcr_exp10(x); }
double lcr_oracle_log(double x) { return lcr_log(x); }
float lcr_oracle_sor(int a, int b, int c, int d) { return lcr_strand_odds_ratio(a, b, c, d); }
int lcr_oracle_binom(uint32_t k, uint32_t n) { return lcr_binom_two_tailed_lt_0p05(k, n); }
double lcr_oracle_uniform(uint64_t seed, int32_t tid, uint32_t start, uint32_t stream, uint32_t call, uint32_t idx) {
    return lcr_uniform(seed, lcr_region_key(tid, start), stream, call, idx);
}
void lcr_oracle_luts(lcr_luts *t) { lcr_build_luts(t); }
int lcr_oracle_f64_as_i32(double v) { return lcr_f64_as_i32(v); }
/* test hooks for the seeded shuffle of --downsample (include/lcr_contract.h) */
void lcr_oracle_chacha_block(const uint32_t *key, uint64_t counter, int rounds, uint32_t *out) { lcr_chacha_block(key, counter, rounds, out); }
void lcr_oracle_shuffle(uint64_t seed, uint32_t n, uint32_t *idx) {
    lcr_chacha12 g;
    lcr_stdrng_seed_from_u64(&g, seed);
    for (uint32_t i = 0; i < n; ++i) idx[i] = i;
    lcr_stdrng_shuffle(&g, idx, n);
}

} /* extern "C" */
