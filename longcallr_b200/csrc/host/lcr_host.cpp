/*
 * lcr_host.cpp — BAM / FASTA decode and isolated-region discovery (host, C++17 + zlib).
 * See lcr_host.h for the reference lines each entry point stands in for.
 */
#include "lcr_host.h"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "lcr_host_impl.h"

namespace lcrhost {

bool read_file(const char *path, std::vector<uint8_t> &buf) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)n);
    size_t got = n ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}

/* BGZF: a series of gzip members, each with a BC extra subfield holding its size */
struct BgzfBlock {
    size_t in_off, in_len;   /* raw deflate payload */
    size_t out_off, out_len; /* position in the inflated stream */
};

static bool bgzf_index(const std::vector<uint8_t> &file, std::vector<BgzfBlock> &blocks, size_t &total) {
    size_t p = 0;
    total = 0;
    while (p + 18 <= file.size()) {
        const uint8_t *h = file.data() + p;
        if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) return false;
        uint16_t xlen = (uint16_t)(h[10] | (h[11] << 8));
        size_t x = p + 12, xend = x + xlen;
        int bsize = -1;
        while (x + 4 <= xend) {
            uint16_t slen = (uint16_t)(file[x + 2] | (file[x + 3] << 8));
            if (file[x] == 'B' && file[x + 1] == 'C' && slen == 2) bsize = file[x + 4] | (file[x + 5] << 8);
            x += 4 + slen;
        }
        if (bsize < 0) return false;
        size_t block_len = (size_t)bsize + 1;
        if (p + block_len > file.size()) return false;
        const uint8_t *tail = file.data() + p + block_len - 8;
        uint32_t isize = (uint32_t)tail[4] | ((uint32_t)tail[5] << 8) | ((uint32_t)tail[6] << 16) | ((uint32_t)tail[7] << 24);
        BgzfBlock b;
        b.in_off = xend;
        b.in_len = p + block_len - 8 - xend;
        b.out_off = total;
        b.out_len = isize;
        blocks.push_back(b);
        total += isize;
        p += block_len;
    }
    return p == file.size();
}

bool bgzf_inflate(const std::vector<uint8_t> &file, std::vector<uint8_t> &out, int n_threads) {
    std::vector<BgzfBlock> blocks;
    size_t total;
    if (!bgzf_index(file, blocks, total)) return false;
    out.resize(total);
    std::atomic<size_t> next{0};
    std::atomic<bool> ok{true};
    auto work = [&]() {
        z_stream zs;
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= blocks.size()) break;
            const BgzfBlock &b = blocks[i];
            if (!b.out_len) continue;
            memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -15) != Z_OK) { ok = false; break; }
            zs.next_in = const_cast<Bytef *>(file.data() + b.in_off);
            zs.avail_in = (uInt)b.in_len;
            zs.next_out = out.data() + b.out_off;
            zs.avail_out = (uInt)b.out_len;
            int rc = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            if (rc != Z_STREAM_END) { ok = false; break; }
        }
    };
    if (n_threads <= 1) work();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < n_threads; ++i) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
    return ok;
}

static inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

/* walk the aux block for ts (kept only when of type A, util.rs:673-679 compares against Aux::Char)
   and de (kept only when of type f, util.rs:661-668) */
static bool parse_aux(const uint8_t *p, const uint8_t *end, int8_t &ts, float &de) {
    ts = '*';
    de = NAN;
    while (p + 3 <= end) {
        const uint8_t t0 = p[0], t1 = p[1], ty = p[2];
        p += 3;
        size_t sz = 0;
        switch (ty) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'Z': case 'H': {
                const uint8_t *q = p;
                while (q < end && *q) ++q;
                if (q >= end) return false;
                sz = (size_t)(q - p) + 1;
                break;
            }
            case 'B': {
                if (p + 5 > end) return false;
                uint8_t sub = p[0];
                uint32_t cnt = rd32(p + 1);
                size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                sz = 5 + es * (size_t)cnt;
                break;
            }
            default: return false;
        }
        if (p + sz > end) return false;
        if (t0 == 't' && t1 == 's' && ty == 'A') ts = (int8_t)p[0];
        if (t0 == 'd' && t1 == 'e' && ty == 'f') { uint32_t u = rd32(p); memcpy(&de, &u, 4); }
        p += sz;
    }
    return true;
}

int read_bam(const char *path, int n_threads, Reads &R) {
    std::vector<uint8_t> file, raw;
    if (!read_file(path, file)) return LCR_ERR_INVALID_ARG;
    if (!bgzf_inflate(file, raw, n_threads)) return LCR_ERR_INVALID_ARG;
    std::vector<uint8_t>().swap(file);
    if (raw.size() < 12 || memcmp(raw.data(), "BAM\1", 4)) return LCR_ERR_INVALID_ARG;
    size_t p = 4;
    uint32_t l_text = rd32(&raw[p]);
    p += 4 + l_text;
    uint32_t n_ref = rd32(&raw[p]);
    p += 4;
    for (uint32_t i = 0; i < n_ref; ++i) {
        uint32_t l_name = rd32(&raw[p]);
        p += 4;
        R.contig_names.emplace_back((const char *)&raw[p], l_name ? l_name - 1 : 0);
        p += l_name;
        R.contig_lens.push_back(rd32(&raw[p]));
        p += 4;
    }
    static const char NIB[] = "=ACMGRSVTWYHKDBN";
    R.seq_off.push_back(0);
    R.cig_off.push_back(0);
    R.qname_off.push_back(0);
    while (p + 4 <= raw.size()) {
        uint32_t bs = rd32(&raw[p]);
        p += 4;
        if (p + bs > raw.size() || bs < 32) return LCR_ERR_INVALID_ARG;
        const uint8_t *b = &raw[p];
        const uint8_t *bend = b + bs;
        p += bs;
        int32_t tid = (int32_t)rd32(b), pos = (int32_t)rd32(b + 4);
        uint8_t l_name = b[8], mapq = b[9];
        uint16_t n_cig = rd16(b + 12), flag = rd16(b + 14);
        uint32_t l_seq = rd32(b + 16);
        const uint8_t *q = b + 32;
        const uint8_t *cig = q + l_name;
        const uint8_t *sq = cig + 4 * (size_t)n_cig;
        const uint8_t *ql = sq + (l_seq + 1) / 2;
        const uint8_t *aux = ql + l_seq;
        if (aux > bend) return LCR_ERR_INVALID_ARG;
        R.tid.push_back(tid);
        R.pos.push_back(pos);
        R.flag.push_back(flag);
        R.mapq.push_back(mapq);
        R.qnames.insert(R.qnames.end(), (const char *)q, (const char *)q + (l_name ? l_name - 1 : 0));
        R.qname_off.push_back(R.qnames.size());
        for (uint16_t c = 0; c < n_cig; ++c) R.cigar.push_back(rd32(cig + 4 * c));
        R.cig_off.push_back(R.cigar.size());
        size_t s0 = R.seq.size();
        R.seq.resize(s0 + l_seq);
        R.qual.resize(s0 + l_seq);
        for (uint32_t i = 0; i < l_seq; ++i) {
            uint8_t byte = sq[i >> 1];
            R.seq[s0 + i] = (uint8_t)NIB[(i & 1) ? (byte & 0xf) : (byte >> 4)];
        }
        if (l_seq) memcpy(&R.qual[s0], ql, l_seq);
        R.seq_off.push_back(R.seq.size());
        int8_t ts;
        float de;
        if (!parse_aux(aux, bend, ts, de)) return LCR_ERR_INVALID_ARG;
        R.ts.push_back(ts);
        R.de.push_back(de);
    }
    R.finish();
    return 0;
}

void Reads::finish() {
    name_ptrs.clear();
    for (auto &s : contig_names) name_ptrs.push_back(s.c_str());
    view.n_reads = (uint32_t)pos.size();
    view.n_contigs = (uint32_t)contig_names.size();
    view.contig_names = name_ptrs.data();
    view.contig_lens = contig_lens.data();
    view.tid = tid.data();
    view.pos = pos.data();
    view.flag = flag.data();
    view.mapq = mapq.data();
    view.ts = ts.data();
    view.de = de.data();
    view.seq_off = seq_off.data();
    view.cig_off = cig_off.data();
    view.seq = seq.data();
    view.qual = qual.data();
    view.cigar = cigar.data();
    view.qname_off = qname_off.data();
    view.qnames = qnames.data();
}

void Fasta::finish() {
    name_ptrs.clear();
    seq_ptrs.clear();
    lens.clear();
    for (size_t i = 0; i < names.size(); ++i) {
        name_ptrs.push_back(names[i].c_str());
        seq_ptrs.push_back(seqs[i].data());
        lens.push_back(seqs[i].size());
    }
    view.n_contigs = (uint32_t)names.size();
    view.names = name_ptrs.data();
    view.lens = lens.data();
    view.seqs = seq_ptrs.data();
}

/* bio::io::fasta::Reader: id = first word of the header, sequence lines concatenated, case preserved */
int read_fasta(const char *path, Fasta &F) {
    std::vector<uint8_t> buf;
    if (!read_file(path, buf)) return LCR_ERR_INVALID_ARG;
    size_t p = 0, n = buf.size();
    while (p < n) {
        size_t e = p;
        while (e < n && buf[e] != '\n') ++e;
        size_t le = e;
        if (le > p && buf[le - 1] == '\r') --le;
        if (le > p && buf[p] == '>') {
            size_t w = p + 1;
            while (w < le && buf[w] != ' ' && buf[w] != '\t') ++w;
            F.names.emplace_back((const char *)&buf[p + 1], w - p - 1);
            F.seqs.emplace_back();
        } else if (le > p && !F.seqs.empty()) {
            F.seqs.back().insert(F.seqs.back().end(), buf.begin() + p, buf.begin() + le);
        }
        p = e + 1;
    }
    F.finish();
    return 0;
}

static inline int64_t ref_span(const lcr_reads &R, uint32_t i) {
    int64_t rlen = 0;
    for (uint64_t c = R.cig_off[i]; c < R.cig_off[i + 1]; ++c) {
        uint32_t op = R.cigar[c] & 0xf, len = R.cigar[c] >> 4;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += len;
    }
    return rlen;
}

/* find_isolated_regions_with_depth (util.rs:236-332): depth over [reference_start, reference_end)
   of every passing read, introns included; maximal runs of depth > 0 become regions.  The depth
   vector is built from a difference array instead of the reference's per-position increments. */
int find_regions(const lcr_reads &R, const lcr_params &P, bool truncation, uint32_t trunc_cov, Regions &out) {
    std::vector<int64_t> endpos(R.n_reads);
    for (uint32_t i = 0; i < R.n_reads; ++i) endpos[i] = (int64_t)R.pos[i] + ref_span(R, i);
    uint32_t lo = 0;
    for (uint32_t t = 0; t < R.n_contigs; ++t) {
        while (lo < R.n_reads && R.tid[lo] < (int32_t)t) ++lo;
        uint32_t hi = lo;
        while (hi < R.n_reads && R.tid[hi] == (int32_t)t) ++hi;
        if (hi == lo) continue;
        const int64_t L = (int64_t)R.contig_lens[t];
        std::vector<int32_t> diff((size_t)L + 1, 0);
        for (uint32_t i = lo; i < hi; ++i) {
            uint64_t l_seq = R.seq_off[i + 1] - R.seq_off[i];
            if ((int32_t)R.mapq[i] < P.min_mapq || l_seq < (uint64_t)P.min_read_length || (R.flag[i] & 0x4) || (R.flag[i] & 0x100) || (R.flag[i] & 0x800)) continue;
            float de = R.de[i];
            if (!(de != de) && de >= P.divergence) continue;
            int64_t a = std::max<int64_t>(R.pos[i], 0), b = std::min<int64_t>(endpos[i], L);
            if (b > a) { diff[a]++; diff[b]--; }
        }
        int64_t region_start = -1, region_end = -1;
        uint32_t max_cov = 0;
        int32_t depth = 0;
        size_t first_region = out.regions.size();
        auto push = [&]() {
            lcr_region rg;
            rg.tid = (int32_t)t;
            rg.start = (uint32_t)(region_start + 1);
            rg.end = (uint32_t)(region_end + 2);
            rg.read_begin = rg.read_end = 0;
            out.regions.push_back(rg);
            out.max_coverage.push_back(max_cov);
        };
        for (int64_t i = 0; i < L; ++i) {
            depth += diff[i];
            const uint32_t dv = (uint32_t)depth;
            if (dv > max_cov) max_cov = dv;
            if (dv == 0 || (truncation && dv > trunc_cov)) {
                if (region_end > region_start) {
                    push();
                    region_start = region_end = -1;
                    max_cov = 0;
                }
            } else {
                if (region_start == -1) region_start = region_end = i;
                else region_end = i;
            }
        }
        if (region_end > region_start) push();
        /* reads of each region: a contiguous superset of what fetch((chr,start,end)) returns */
        std::vector<int64_t> pmax(hi - lo);
        int64_t m = INT64_MIN;
        for (uint32_t i = lo; i < hi; ++i) {
            int64_t e = endpos[i] > R.pos[i] ? endpos[i] : (int64_t)R.pos[i] + 1;
            m = std::max(m, e);
            pmax[i - lo] = m;
        }
        for (size_t r = first_region; r < out.regions.size(); ++r) {
            lcr_region &rg = out.regions[r];
            uint32_t b = (uint32_t)(std::upper_bound(pmax.begin(), pmax.end(), (int64_t)rg.start) - pmax.begin());
            uint32_t e = (uint32_t)(std::lower_bound(R.pos + lo, R.pos + hi, (int32_t)rg.end) - (R.pos + lo));
            if (e < b) e = b;
            rg.read_begin = lo + b;
            rg.read_end = lo + e;
        }
        lo = hi;
    }
    out.view.n_regions = (uint32_t)out.regions.size();
    out.view.regions = out.regions.data();
    out.view.max_coverage = out.max_coverage.data();
    return 0;
}

} // namespace lcrhost

using namespace lcrhost;

extern "C" {

int lcr_host_read_bam(const char *path, int n_threads, lcr_reads **out) {
    if (!path || !out) return LCR_ERR_INVALID_ARG;
    Reads *R = new Reads();
    int rc = read_bam(path, n_threads, *R);
    if (rc) { delete R; return rc; }
    *out = &R->view;
    return 0;
}
void lcr_host_free_reads(lcr_reads *r) {
    if (r) delete reinterpret_cast<Reads *>(reinterpret_cast<char *>(r) - offsetof(Reads, view));
}
int lcr_host_read_fasta(const char *path, lcr_fasta **out) {
    if (!path || !out) return LCR_ERR_INVALID_ARG;
    Fasta *F = new Fasta();
    int rc = read_fasta(path, *F);
    if (rc) { delete F; return rc; }
    *out = &F->view;
    return 0;
}
void lcr_host_free_fasta(lcr_fasta *f) {
    if (f) delete reinterpret_cast<Fasta *>(reinterpret_cast<char *>(f) - offsetof(Fasta, view));
}
int lcr_host_find_regions(const lcr_reads *reads, const lcr_params *p, int truncation, uint32_t truncation_coverage, lcr_region_list **out) {
    if (!reads || !p || !out) return LCR_ERR_INVALID_ARG;
    Regions *G = new Regions();
    int rc = find_regions(*reads, *p, truncation != 0, truncation_coverage, *G);
    if (rc) { delete G; return rc; }
    *out = &G->view;
    return 0;
}
void lcr_host_free_regions(lcr_region_list *r) {
    if (r) delete reinterpret_cast<Regions *>(reinterpret_cast<char *>(r) - offsetof(Regions, view));
}
void lcr_host_free_text(char *t) { free(t); }

} /* extern "C" */
