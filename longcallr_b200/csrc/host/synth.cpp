/*
 * synth.cpp — seeded synthetic long-read RNA alignments.
 *
 * Produces the same decoded arrays read_bam() produces, so the synthetic
 * configurations of BASELINE.json feed the C ABI exactly like a BAM does.
 * Deterministic for a given lcr_synth_config regardless of n_threads: every
 * gene and every read draws from its own counter-keyed splitmix64 stream.
 *
 * Model: iid ACGT reference; genes of 1..max_exons exons (150-1500 bp) separated by
 * introns (100..max_intron bp) and by zero-coverage gaps, so that each gene is one
 * isolated region (src/util.rs:287-330); planted het SNPs (>= 6 bp apart) on two
 * haplotypes and A>G editing sites at 30 % read fraction; reads are sub-intervals of the
 * spliced transcript with length ~ LogNormal(ln 1100, 0.45) clipped to [500, 6000];
 * base qualities, substitution errors drawn from the qualities, insertions, deletions,
 * soft clips, strand and the ts / de tags as minimap2 would report them.
 */
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <numeric>
#include <thread>

#include "lcr_contract.h"
#include "lcr_host_impl.h"

namespace lcrhost {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { return lcr_mix64(s += 0x9e3779b97f4a7c15ULL); }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
    uint32_t range(uint32_t lo, uint32_t hi) { return lo + below(hi - lo + 1); } /* inclusive */
    double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

struct Gene {
    std::vector<std::pair<int64_t, int64_t>> exons; /* [start, end) on the contig */
    int64_t tlen = 0;
    int strand = 0; /* transcript strand: 0 '+', 1 '-' */
};

struct GeneReads { /* reads of one gene before the global sort */
    std::vector<int32_t> pos;
    std::vector<uint16_t> flag;
    std::vector<int8_t> ts, hap;
    std::vector<float> de;
    std::vector<uint64_t> seq_off{0}, cig_off{0};
    std::vector<uint8_t> seq, qual;
    std::vector<uint32_t> cigar;
};

struct SynthBox {
    lcr_synth view{};
    Reads reads;
    Fasta fasta;
    std::vector<int32_t> het_tid;
    std::vector<int64_t> het_pos;
    std::vector<uint8_t> het_alt;
    std::vector<int8_t> het_hap, read_hap;
};

static const char ACGT[] = "ACGT";

static inline uint8_t hifi_qual(uint32_t r20) { /* categorical over 2^20 */
    static const struct { double p; uint8_t q; } T[] = {{.45, 40}, {.38, 50}, {.07, 35}, {.03, 27}, {.02, 22}, {.02, 17}, {.02, 10}, {.01, 3}};
    double u = r20 * (1.0 / 1048576.0), acc = 0;
    for (auto &t : T) { acc += t.p; if (u < acc) return t.q; }
    return 3;
}

static void push_op(std::vector<uint32_t> &cig, size_t first, uint32_t op, uint32_t len) {
    if (!len) return;
    if (cig.size() > first && (cig.back() & 0xf) == op) cig.back() += len << 4;
    else cig.push_back((len << 4) | op);
}

/* Per-site truth of one contig: 0 = nothing, else index+1 into the het table, or 0xFFFFFFFF for an editing site */
struct Truth {
    std::vector<int64_t> pos; /* sorted site positions */
    std::vector<uint8_t> alt; /* alt base code */
    std::vector<int8_t> hap;  /* -1 = editing site, 0/1 = haplotype carrying alt */
    int find(int64_t p) const {
        auto it = std::lower_bound(pos.begin(), pos.end(), p);
        return (it != pos.end() && *it == p) ? (int)(it - pos.begin()) : -1;
    }
};

static void make_reads(const lcr_synth_config &cfg, uint32_t contig, uint32_t gene_id, const Gene &g, const std::vector<uint8_t> &ref, const Truth &truth, GeneReads &out) {
    Rng grng(lcr_mix64(cfg.seed ^ 0x5eed5eedULL) ^ lcr_mix64(((uint64_t)contig << 32) | gene_id));
    const double target = (double)cfg.depth * (double)g.tlen;
    const double ins_rate = cfg.platform == 1 ? 0.01 : 0.001, del_rate = ins_rate;
    const uint32_t ins_thr = (uint32_t)(ins_rate * 1048576.0), del_thr = (uint32_t)(del_rate * 1048576.0);
    /* exon prefix sums for transcript -> genome mapping */
    std::vector<int64_t> pre(g.exons.size() + 1, 0);
    for (size_t e = 0; e < g.exons.size(); ++e) pre[e + 1] = pre[e] + (g.exons[e].second - g.exons[e].first);
    double covered = 0;
    uint32_t ridx = 0;
    double eps24[64];
    for (int q = 0; q < 64; ++q) eps24[q] = std::pow(10.0, -(double)q / 10.0) * 16777216.0;
    /* the site positions of this gene, for the fast "is this a site" test */
    while (covered < target) {
        Rng r(lcr_mix64(grng.next() ^ ridx));
        ++ridx;
        double ln = std::exp(std::log(1100.0) + 0.45 * r.normal());
        int64_t len = (int64_t)std::llround(ln);
        len = std::max<int64_t>(500, std::min<int64_t>(6000, len));
        len = std::min<int64_t>(len, g.tlen);
        const int64_t t0 = (g.tlen > len) ? (int64_t)(r.uni() * (double)(g.tlen - len + 1)) : 0;
        covered += (double)len;
        const int hap = (int)(r.next() & 1);
        const int rstrand = cfg.both_strands ? (int)(r.next() & 1) : 0;
        const int8_t ts = (rstrand == 0) ? (g.strand ? '-' : '+') : (g.strand ? '+' : '-');
        const size_t cig_first = out.cigar.size();
        /* leading soft clip */
        uint32_t lead = 0, trail = 0;
        if (r.uni() < 0.2) {
            if (r.next() & 1) lead = r.range(5, 60);
            if (!lead || (r.next() & 1)) trail = r.range(5, 60);
        }
        for (uint32_t i = 0; i < lead; ++i) { out.seq.push_back((uint8_t)ACGT[r.below(4)]); out.qual.push_back(cfg.platform == 1 ? 8 : 20); }
        push_op(out.cigar, cig_first, 4, lead);
        /* locate the first exon */
        size_t e = (size_t)(std::upper_bound(pre.begin(), pre.end(), t0) - pre.begin()) - 1;
        int64_t gpos = g.exons[e].first + (t0 - pre[e]);
        const int64_t read_pos = gpos;
        int64_t remaining = len;
        uint64_t mism = 0, events = 0, mcols = 0;
        bool first_base = true;
        while (remaining > 0) {
            const int64_t exon_end = g.exons[e].second;
            int64_t seg = std::min<int64_t>(remaining, exon_end - gpos);
            int64_t i = 0;
            size_t si_next = (size_t)(std::lower_bound(truth.pos.begin(), truth.pos.end(), gpos) - truth.pos.begin());
            while (i < seg) {
                const uint64_t w = r.next();
                const uint32_t r_indel = (uint32_t)(w & 0xfffff), r_q = (uint32_t)((w >> 20) & 0xfffff), r_err = (uint32_t)(w >> 40);
                const bool edge = first_base || i == 0 || (remaining - i) <= 1;
                if (!edge && r_indel < del_thr && i + 1 < seg) { /* deletion (never at a read or exon edge) */
                    uint32_t dl = 1;
                    while (dl < 8 && (r.next() & 0xff) < 77) ++dl; /* geometric, p = 0.7 */
                    dl = (uint32_t)std::min<int64_t>(dl, seg - i - 1);
                    if (dl) { push_op(out.cigar, cig_first, 2, dl); i += dl; events++; continue; }
                }
                if (!edge && r_indel >= 0x80000 && r_indel - 0x80000 < ins_thr) { /* insertion before this base */
                    uint32_t il = 1;
                    while (il < 8 && (r.next() & 0xff) < 77) ++il;
                    for (uint32_t k = 0; k < il; ++k) { out.seq.push_back((uint8_t)ACGT[r.below(4)]); out.qual.push_back(cfg.platform == 1 ? 6 : 15); }
                    push_op(out.cigar, cig_first, 1, il);
                    events++;
                }
                const int64_t p = gpos + i;
                uint8_t q;
                if (cfg.platform == 1) {
                    /* round(Normal(18, 6)) clipped to [1, 50], from 20 bits through a 12-term sum approximation */
                    uint64_t h = lcr_mix64(w);
                    double z = 0;
                    for (int k = 0; k < 4; ++k) z += (double)((h >> (16 * k)) & 0xffff) * (1.0 / 65536.0);
                    z = (z - 2.0) * 1.7320508075688772; /* var of sum of 4 U(0,1) = 1/3 */
                    int qi = (int)std::lround(18.0 + 6.0 * z);
                    q = (uint8_t)std::max(1, std::min(50, qi));
                    (void)r_q;
                } else q = hifi_qual(r_q);
                uint8_t base = ref[p];
                while (si_next < truth.pos.size() && truth.pos[si_next] < p) ++si_next;
                const int si = (si_next < truth.pos.size() && truth.pos[si_next] == p) ? (int)si_next : -1;
                if (si >= 0) {
                    if (truth.hap[si] < 0) { if ((lcr_mix64(w ^ 0xed17) & 0xffff) < 19661) base = (uint8_t)ACGT[truth.alt[si]]; }
                    else if (truth.hap[si] == hap) base = (uint8_t)ACGT[truth.alt[si]];
                }
                /* substitution error with probability 10^(-q/10), 24-bit resolution */
                if ((double)r_err < eps24[q]) {
                    int bc = base == 'A' ? 0 : base == 'C' ? 1 : base == 'G' ? 2 : 3;
                    base = (uint8_t)ACGT[(bc + 1 + (int)(lcr_mix64(w ^ 0xe440) % 3)) & 3];
                }
                if (base != ref[p]) mism++;
                out.seq.push_back(base);
                out.qual.push_back(q);
                push_op(out.cigar, cig_first, 0, 1);
                mcols++;
                first_base = false;
                ++i;
            }
            remaining -= seg;
            gpos += seg;
            if (remaining > 0) { /* jump the intron */
                const int64_t next_start = g.exons[e + 1].first;
                push_op(out.cigar, cig_first, 3, (uint32_t)(next_start - gpos));
                gpos = next_start;
                ++e;
            }
        }
        for (uint32_t i = 0; i < trail; ++i) { out.seq.push_back((uint8_t)ACGT[r.below(4)]); out.qual.push_back(cfg.platform == 1 ? 8 : 20); }
        push_op(out.cigar, cig_first, 4, trail);
        out.pos.push_back((int32_t)read_pos);
        out.flag.push_back(rstrand ? 16 : 0);
        out.ts.push_back(ts);
        out.hap.push_back((int8_t)hap);
        out.de.push_back((float)((double)(mism + events) / (double)std::max<uint64_t>(1, mcols + events)));
        out.seq_off.push_back(out.seq.size());
        out.cig_off.push_back(out.cigar.size());
    }
}

int synth(const lcr_synth_config &cfg, SynthBox &S) {
    if (!cfg.contig_len || !cfg.n_contigs || cfg.depth <= 0) return LCR_ERR_INVALID_ARG;
    const uint32_t nthreads = std::max<uint32_t>(1, cfg.n_threads);
    const uint32_t max_exons = std::max<uint32_t>(1, cfg.max_exons);
    const uint32_t max_intron = std::max<uint32_t>(100, cfg.max_intron);
    std::vector<std::vector<Gene>> genes(cfg.n_contigs);
    std::vector<Truth> truths(cfg.n_contigs);
    S.fasta.names.resize(cfg.n_contigs);
    S.fasta.seqs.resize(cfg.n_contigs);
    for (uint32_t c = 0; c < cfg.n_contigs; ++c) {
        S.fasta.names[c] = "synth" + std::to_string(c + 1);
        std::vector<uint8_t> &ref = S.fasta.seqs[c];
        ref.resize(cfg.contig_len);
        /* reference bases: 32 per hash */
        {
            std::atomic<uint64_t> next{0};
            const uint64_t chunk = 1 << 20;
            auto work = [&]() {
                for (;;) {
                    uint64_t b = next.fetch_add(chunk);
                    if (b >= cfg.contig_len) break;
                    uint64_t e = std::min<uint64_t>(b + chunk, cfg.contig_len);
                    for (uint64_t i = b; i < e; i += 32) {
                        uint64_t h = lcr_mix64(lcr_mix64(cfg.seed ^ ((uint64_t)c << 48)) ^ (i >> 5));
                        for (uint64_t k = i; k < std::min<uint64_t>(i + 32, e); ++k) { ref[k] = (uint8_t)ACGT[h & 3]; h >>= 2; }
                    }
                }
            };
            std::vector<std::thread> th;
            for (uint32_t t = 0; t < nthreads; ++t) th.emplace_back(work);
            for (auto &t : th) t.join();
        }
        /* gene layout */
        Rng lr(lcr_mix64(cfg.seed ^ 0x6e0e) ^ lcr_mix64(c + 1));
        int64_t x = 1000;
        const int64_t L = (int64_t)cfg.contig_len;
        while (x + 2000 < L) {
            Gene g;
            g.strand = cfg.both_strands ? (int)(lr.next() & 1) : 0;
            if (cfg.single_region) {
                while (x + 200 < L - 1000) {
                    int64_t el = lr.range(150, 1500);
                    el = std::min<int64_t>(el, L - 1000 - x);
                    g.exons.emplace_back(x, x + el);
                    x += el + lr.range(100, max_intron);
                }
            } else {
                uint32_t ne = lr.range(1, max_exons);
                for (uint32_t k = 0; k < ne; ++k) {
                    int64_t el = lr.range(150, 1500);
                    if (x + el + 1000 >= L) break;
                    g.exons.emplace_back(x, x + el);
                    x += el;
                    if (k + 1 < ne) x += lr.range(100, max_intron);
                }
            }
            if (g.exons.empty()) break;
            for (auto &ex : g.exons) g.tlen += ex.second - ex.first;
            if (g.tlen < 700) { /* keep every transcript longer than min_read_length */
                int64_t add = 700 - g.tlen;
                if (g.exons.back().second + add + 1000 >= L) break;
                g.exons.back().second += add;
                g.tlen += add;
                x = std::max(x, g.exons.back().second);
            }
            x = std::max(x, g.exons.back().second) + lr.range(300, std::max<uint32_t>(300, cfg.max_gap));
            genes[c].push_back(std::move(g));
            if (cfg.single_region) break;
        }
        /* planted sites over exonic positions */
        std::vector<std::pair<int64_t, int64_t>> exs;
        for (auto &g : genes[c]) for (auto &ex : g.exons) exs.push_back(ex);
        std::vector<int64_t> pre(exs.size() + 1, 0);
        for (size_t e = 0; e < exs.size(); ++e) pre[e + 1] = pre[e] + (exs[e].second - exs[e].first);
        const int64_t E = pre.back();
        if (E > 0) {
            Rng sr(lcr_mix64(cfg.seed ^ 0x517e) ^ lcr_mix64(c + 1));
            const uint32_t want = cfg.n_het + cfg.n_edit;
            std::vector<int64_t> offs(want);
            for (auto &o : offs) o = (int64_t)(sr.uni() * (double)E);
            std::sort(offs.begin(), offs.end());
            std::vector<int64_t> sites;
            int64_t last = -100;
            for (int64_t o : offs) {
                size_t e = (size_t)(std::upper_bound(pre.begin(), pre.end(), o) - pre.begin()) - 1;
                int64_t p = exs[e].first + (o - pre[e]);
                if (p - last < 6) continue;
                sites.push_back(p);
                last = p;
            }
            /* choose which sites are editing sites: every k-th site whose reference base is A */
            Truth &T = truths[c];
            uint32_t edits_left = cfg.n_edit;
            const double edit_frac = want ? (double)cfg.n_edit / (double)want : 0.0;
            for (int64_t p : sites) {
                const uint8_t rb = ref[p];
                const int rc = rb == 'A' ? 0 : rb == 'C' ? 1 : rb == 'G' ? 2 : 3;
                if (edits_left && sr.uni() < edit_frac * 4.0) { /* only ~1/4 of sites sit on an A */
                    if (rb == 'A') {
                        T.pos.push_back(p); T.alt.push_back(2); T.hap.push_back(-1);
                        edits_left--;
                        continue;
                    }
                }
                T.pos.push_back(p);
                T.alt.push_back((uint8_t)((rc + 1 + sr.below(3)) & 3));
                T.hap.push_back((int8_t)(sr.next() & 1));
            }
            for (size_t i = 0; i < T.pos.size(); ++i)
                if (T.hap[i] >= 0) {
                    S.het_tid.push_back((int32_t)c);
                    S.het_pos.push_back(T.pos[i]);
                    S.het_alt.push_back((uint8_t)ACGT[T.alt[i]]);
                    S.het_hap.push_back(T.hap[i]);
                }
        }
    }
    S.fasta.finish();
    /* reads, gene by gene */
    struct Job { uint32_t contig, gene; };
    std::vector<Job> jobs;
    for (uint32_t c = 0; c < cfg.n_contigs; ++c)
        for (uint32_t g = 0; g < genes[c].size(); ++g) jobs.push_back({c, g});
    std::vector<GeneReads> gr(jobs.size());
    {
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (;;) {
                size_t j = next.fetch_add(1);
                if (j >= jobs.size()) break;
                make_reads(cfg, jobs[j].contig, jobs[j].gene, genes[jobs[j].contig][jobs[j].gene], S.fasta.seqs[jobs[j].contig], truths[jobs[j].contig], gr[j]);
            }
        };
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < nthreads; ++t) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
    /* global coordinate sort (stable in generation order) and assembly */
    struct Key { int32_t tid, pos; uint32_t job, idx; };
    std::vector<Key> keys;
    for (size_t j = 0; j < jobs.size(); ++j)
        for (uint32_t i = 0; i < gr[j].pos.size(); ++i) keys.push_back({(int32_t)jobs[j].contig, gr[j].pos[i], (uint32_t)j, i});
    std::stable_sort(keys.begin(), keys.end(), [](const Key &a, const Key &b) { return a.tid != b.tid ? a.tid < b.tid : a.pos < b.pos; });
    Reads &R = S.reads;
    const size_t n = keys.size();
    for (uint32_t c = 0; c < cfg.n_contigs; ++c) { R.contig_names.push_back(S.fasta.names[c]); R.contig_lens.push_back(cfg.contig_len); }
    R.tid.resize(n); R.pos.resize(n); R.flag.resize(n); R.mapq.assign(n, 60); R.ts.resize(n); R.de.resize(n);
    R.seq_off.assign(n + 1, 0); R.cig_off.assign(n + 1, 0); R.qname_off.assign(n + 1, 0);
    S.read_hap.resize(n);
    for (size_t k = 0; k < n; ++k) {
        const GeneReads &G = gr[keys[k].job];
        const uint32_t i = keys[k].idx;
        R.seq_off[k + 1] = R.seq_off[k] + (G.seq_off[i + 1] - G.seq_off[i]);
        R.cig_off[k + 1] = R.cig_off[k] + (G.cig_off[i + 1] - G.cig_off[i]);
    }
    R.seq.resize(R.seq_off[n]);
    R.qual.resize(R.seq_off[n]);
    R.cigar.resize(R.cig_off[n]);
    {
        std::atomic<size_t> next{0};
        const size_t chunk = 4096;
        auto work = [&]() {
            for (;;) {
                size_t b = next.fetch_add(chunk);
                if (b >= n) break;
                size_t e = std::min(b + chunk, n);
                for (size_t k = b; k < e; ++k) {
                    const GeneReads &G = gr[keys[k].job];
                    const uint32_t i = keys[k].idx;
                    R.tid[k] = keys[k].tid; R.pos[k] = G.pos[i]; R.flag[k] = G.flag[i]; R.ts[k] = G.ts[i]; R.de[k] = G.de[i];
                    S.read_hap[k] = G.hap[i];
                    const size_t sl = G.seq_off[i + 1] - G.seq_off[i], cl = G.cig_off[i + 1] - G.cig_off[i];
                    if (sl) { memcpy(&R.seq[R.seq_off[k]], &G.seq[G.seq_off[i]], sl); memcpy(&R.qual[R.seq_off[k]], &G.qual[G.seq_off[i]], sl); }
                    if (cl) memcpy(&R.cigar[R.cig_off[k]], &G.cigar[G.cig_off[i]], cl * 4);
                }
            }
        };
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < nthreads; ++t) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
    std::vector<GeneReads>().swap(gr);
    /* read names: s<ordinal> */
    for (size_t k = 0; k < n; ++k) {
        char b[24];
        int l = snprintf(b, sizeof b, "s%zu", k);
        R.qnames.insert(R.qnames.end(), b, b + l);
        R.qname_off[k + 1] = R.qnames.size();
    }
    R.finish();
    S.view.reads = &R.view;
    S.view.fasta = &S.fasta.view;
    S.view.n_het_total = (uint32_t)S.het_pos.size();
    S.view.het_tid = S.het_tid.data();
    S.view.het_pos = S.het_pos.data();
    S.view.het_alt = S.het_alt.data();
    S.view.het_hap = S.het_hap.data();
    S.view.read_hap = S.read_hap.data();
    return 0;
}

} // namespace lcrhost

extern "C" {

int lcr_host_synth(const lcr_synth_config *cfg, lcr_synth **out) {
    if (!cfg || !out) return LCR_ERR_INVALID_ARG;
    lcrhost::SynthBox *S = new lcrhost::SynthBox();
    int rc = lcrhost::synth(*cfg, *S);
    if (rc) { delete S; return rc; }
    *out = &S->view;
    return 0;
}
void lcr_host_free_synth(lcr_synth *s) {
    if (s) delete reinterpret_cast<lcrhost::SynthBox *>(reinterpret_cast<char *>(s) - offsetof(lcrhost::SynthBox, view));
}

} /* extern "C" */
