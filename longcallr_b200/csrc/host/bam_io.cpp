/*
 * bam_io.cpp — BAM output: a plain BAM writer for generated alignments and the phased-BAM emit of
 * reference src/thread.rs:307-361.  Host C++17 + zlib, no htslib.
 */
#include <zlib.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "lcr_host.h"
#include "lcr_host_impl.h"

namespace lcrhost {

static inline void wr32(std::vector<uint8_t> &v, uint32_t x) { for (int k = 0; k < 4; ++k) v.push_back((uint8_t)(x >> (8 * k))); }
static inline void wr16(std::vector<uint8_t> &v, uint16_t x) { v.push_back((uint8_t)x); v.push_back((uint8_t)(x >> 8)); }
static inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

/* BGZF: the stream cut into blocks of at most 0xff00 bytes, each a gzip member with the BC extra field; EOF marker last */
static bool bgzf_write(const char *path, const std::vector<uint8_t> &raw, int n_threads) {
    const size_t BLK = 0xff00;
    const size_t nb = (raw.size() + BLK - 1) / BLK;
    std::vector<std::vector<uint8_t>> out(nb);
    std::atomic<size_t> next{0};
    std::atomic<bool> ok{true};
    auto work = [&]() {
        std::vector<uint8_t> buf(BLK + 1024);
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= nb) break;
            const size_t off = i * BLK, len = std::min(BLK, raw.size() - off);
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { ok = false; break; }
            zs.next_in = const_cast<Bytef *>(raw.data() + off);
            zs.avail_in = (uInt)len;
            zs.next_out = buf.data();
            zs.avail_out = (uInt)buf.size();
            const int rc = deflate(&zs, Z_FINISH);
            const size_t clen = buf.size() - zs.avail_out;
            deflateEnd(&zs);
            if (rc != Z_STREAM_END || clen + 26 > 0x10000) { ok = false; break; }
            std::vector<uint8_t> &o = out[i];
            static const uint8_t head[12] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0};
            o.insert(o.end(), head, head + 12);
            o.push_back('B'); o.push_back('C'); wr16(o, 2); wr16(o, (uint16_t)(clen + 25));
            o.insert(o.end(), buf.begin(), buf.begin() + clen);
            wr32(o, (uint32_t)crc32(crc32(0L, Z_NULL, 0), raw.data() + off, (uInt)len));
            wr32(o, (uint32_t)len);
        }
    };
    if (n_threads <= 1) work();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < n_threads; ++i) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
    if (!ok) return false;
    FILE *f = fopen(path, "wb");
    if (!f) return false;
    bool good = true;
    for (auto &o : out) good = good && fwrite(o.data(), 1, o.size(), f) == o.size();
    static const uint8_t eof[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    good = good && fwrite(eof, 1, 28, f) == 28;
    good = fclose(f) == 0 && good;
    return good;
}

/* reg2bin of the SAM specification, section 5.3 */
static inline uint16_t reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (uint16_t)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (uint16_t)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (uint16_t)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (uint16_t)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (uint16_t)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

static inline int64_t cigar_rlen(const uint8_t *cig, uint32_t n_cig) {
    int64_t r = 0;
    for (uint32_t c = 0; c < n_cig; ++c) {
        const uint32_t v = rd32(cig + 4 * (size_t)c), op = v & 0xf;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) r += v >> 4;
    }
    return r;
}

static int write_bam(const char *path, const lcr_reads &R, const char *header_text, int n_threads) {
    std::vector<uint8_t> raw;
    std::string text;
    if (header_text) text = header_text;
    else {
        text = "@HD\tVN:1.6\tSO:coordinate\n";
        for (uint32_t c = 0; c < R.n_contigs; ++c) text += "@SQ\tSN:" + std::string(R.contig_names[c]) + "\tLN:" + std::to_string(R.contig_lens[c]) + "\n";
    }
    raw.insert(raw.end(), {'B', 'A', 'M', 1});
    wr32(raw, (uint32_t)text.size());
    raw.insert(raw.end(), text.begin(), text.end());
    wr32(raw, R.n_contigs);
    for (uint32_t c = 0; c < R.n_contigs; ++c) {
        const size_t l = strlen(R.contig_names[c]);
        wr32(raw, (uint32_t)l + 1);
        raw.insert(raw.end(), R.contig_names[c], R.contig_names[c] + l + 1);
        wr32(raw, (uint32_t)R.contig_lens[c]);
    }
    static int8_t code[256];
    static bool init = false;
    if (!init) {
        memset(code, 15, sizeof code);
        const char *nib = "=ACMGRSVTWYHKDBN";
        for (int k = 0; k < 16; ++k) code[(uint8_t)nib[k]] = (int8_t)k;
        init = true;
    }
    for (uint32_t i = 0; i < R.n_reads; ++i) {
        const uint64_t q0 = R.qname_off ? R.qname_off[i] : 0, q1 = R.qname_off ? R.qname_off[i + 1] : 0;
        std::string qn = q1 > q0 ? std::string(R.qnames + q0, R.qnames + q1) : "r" + std::to_string(i);
        if (qn.size() > 254) return LCR_ERR_INVALID_ARG;
        const uint32_t n_cig = (uint32_t)(R.cig_off[i + 1] - R.cig_off[i]);
        const uint32_t l_seq = (uint32_t)(R.seq_off[i + 1] - R.seq_off[i]);
        if (n_cig > 0xffff) return LCR_ERR_INVALID_ARG; /* the CG:B overflow form is not produced */
        int64_t rlen = 0;
        for (uint32_t c = 0; c < n_cig; ++c) {
            const uint32_t v = R.cigar[R.cig_off[i] + c], op = v & 0xf;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += v >> 4;
        }
        const bool has_ts = R.ts && R.ts[i] != '*';
        const bool has_de = R.de && !(R.de[i] != R.de[i]);
        const uint32_t bs = 32 + (uint32_t)qn.size() + 1 + 4 * n_cig + (l_seq + 1) / 2 + l_seq + (has_ts ? 4 : 0) + (has_de ? 7 : 0);
        wr32(raw, bs);
        wr32(raw, (uint32_t)R.tid[i]);
        wr32(raw, (uint32_t)R.pos[i]);
        raw.push_back((uint8_t)(qn.size() + 1));
        raw.push_back(R.mapq[i]);
        wr16(raw, reg2bin(R.pos[i] < 0 ? 0 : R.pos[i], (R.pos[i] < 0 ? 0 : R.pos[i]) + (rlen ? rlen : 1)));
        wr16(raw, (uint16_t)n_cig);
        wr16(raw, R.flag[i]);
        wr32(raw, l_seq);
        wr32(raw, 0xffffffffu); /* next refID */
        wr32(raw, 0xffffffffu); /* next pos   */
        wr32(raw, 0);           /* tlen       */
        raw.insert(raw.end(), qn.begin(), qn.end());
        raw.push_back(0);
        for (uint32_t c = 0; c < n_cig; ++c) wr32(raw, R.cigar[R.cig_off[i] + c]);
        const uint8_t *sq = R.seq + R.seq_off[i];
        for (uint32_t k = 0; k < l_seq; k += 2) {
            const uint8_t hi = (uint8_t)code[sq[k]], lo = k + 1 < l_seq ? (uint8_t)code[sq[k + 1]] : 0;
            raw.push_back((uint8_t)(hi << 4 | lo));
        }
        raw.insert(raw.end(), R.qual + R.seq_off[i], R.qual + R.seq_off[i] + l_seq);
        if (has_ts) { raw.push_back('t'); raw.push_back('s'); raw.push_back('A'); raw.push_back((uint8_t)R.ts[i]); }
        if (has_de) { raw.push_back('d'); raw.push_back('e'); raw.push_back('f'); uint32_t u; memcpy(&u, &R.de[i], 4); wr32(raw, u); }
    }
    return bgzf_write(path, raw, n_threads) ? 0 : LCR_ERR_INVALID_ARG;
}

struct RecView {
    size_t off;   /* of the block_size field */
    uint32_t bs;
    int32_t tid, pos;
    int64_t endpos;
    uint16_t flag;
    std::string_view qname;
    bool has_hp, has_ps;
};

/* does the aux block hold tag t0t1 (any type)?  false on a malformed block as well */
static bool aux_scan(const uint8_t *p, const uint8_t *end, bool &has_hp, bool &has_ps) {
    has_hp = has_ps = false;
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
                const uint8_t sub = p[0];
                const uint32_t cnt = rd32(p + 1);
                const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                sz = 5 + es * (size_t)cnt;
                break;
            }
            default: return false;
        }
        if (p + sz > end) return false;
        if (t0 == 'H' && t1 == 'P') has_hp = true;
        if (t0 == 'P' && t1 == 'S') has_ps = true;
        p += sz;
    }
    return true;
}

static int write_phased_bam(const char *in_bam, const char *out_bam, const lcr_region *regions, uint32_t n_regions, const int8_t *hp, const uint32_t *ps, const uint8_t *has_entry,
                            uint32_t n_reads, int n_threads, uint64_t *n_written) {
    std::vector<uint8_t> file, raw;
    if (!read_file(in_bam, file)) return LCR_ERR_INVALID_ARG;
    if (!bgzf_inflate(file, raw, n_threads)) return LCR_ERR_INVALID_ARG;
    std::vector<uint8_t>().swap(file);
    if (raw.size() < 12 || memcmp(raw.data(), "BAM\1", 4)) return LCR_ERR_INVALID_ARG;
    size_t p = 4;
    const uint32_t l_text = rd32(&raw[p]);
    p += 4 + (size_t)l_text;
    if (p + 4 > raw.size()) return LCR_ERR_INVALID_ARG;
    const uint32_t n_ref = rd32(&raw[p]);
    p += 4;
    for (uint32_t i = 0; i < n_ref; ++i) {
        if (p + 4 > raw.size()) return LCR_ERR_INVALID_ARG;
        p += 4 + (size_t)rd32(&raw[p]) + 4;
    }
    if (p > raw.size()) return LCR_ERR_INVALID_ARG;
    const size_t header_end = p;
    std::vector<RecView> recs;
    recs.reserve(n_reads);
    while (p + 4 <= raw.size()) {
        const uint32_t bs = rd32(&raw[p]);
        if (p + 4 + bs > raw.size() || bs < 32) return LCR_ERR_INVALID_ARG;
        const uint8_t *b = &raw[p + 4];
        RecView r;
        r.off = p;
        r.bs = bs;
        r.tid = (int32_t)rd32(b);
        r.pos = (int32_t)rd32(b + 4);
        const uint8_t l_name = b[8];
        const uint16_t n_cig = rd16(b + 12);
        r.flag = rd16(b + 14);
        const uint32_t l_seq = rd32(b + 16);
        const uint8_t *cig = b + 32 + l_name;
        const uint8_t *aux = cig + 4 * (size_t)n_cig + (l_seq + 1) / 2 + l_seq;
        if (aux > b + bs) return LCR_ERR_INVALID_ARG;
        int64_t rlen = (r.flag & 0x4) ? 0 : cigar_rlen(cig, n_cig); /* htslib bam_endpos */
        if (rlen == 0) rlen = 1;
        r.endpos = (int64_t)r.pos + rlen;
        r.qname = std::string_view((const char *)b + 32, l_name ? l_name - 1 : 0);
        if (!aux_scan(aux, b + bs, r.has_hp, r.has_ps)) return LCR_ERR_INVALID_ARG;
        recs.push_back(r);
        p += 4 + (size_t)bs;
    }
    if (recs.size() != n_reads) return LCR_ERR_INVALID_ARG;
    /* QNAME -> first record with an entry (thread.rs:308-325) */
    std::unordered_map<std::string_view, uint32_t> first_hp, first_ps;
    first_hp.reserve(n_reads * 2);
    first_ps.reserve(n_reads * 2);
    for (uint32_t i = 0; i < n_reads; ++i) {
        const bool e_hp = has_entry ? has_entry[i] != 0 : hp[i] != 0;
        if (e_hp) first_hp.emplace(recs[i].qname, i);
        if (ps[i] != 0) first_ps.emplace(recs[i].qname, i);
    }
    std::vector<uint8_t> out(raw.begin(), raw.begin() + header_end);
    out.reserve(raw.size() + 14 * (size_t)n_reads / 2);
    uint64_t written = 0;
    for (uint32_t g = 0; g < n_regions; ++g) {
        const lcr_region &rg = regions[g];
        if (rg.read_end > n_reads || rg.read_begin > rg.read_end) return LCR_ERR_INVALID_ARG;
        for (uint32_t i = rg.read_begin; i < rg.read_end; ++i) {
            const RecView &r = recs[i];
            if (r.tid != rg.tid) continue;
            if (!((int64_t)r.pos < (int64_t)rg.end && r.endpos > (int64_t)rg.start)) continue;             /* fetch((chr, start, end)) */
            if ((r.flag & 0x4) || (r.flag & 0x100) || (r.flag & 0x800)) continue;                            /* thread.rs:336-338 */
            if ((int64_t)r.pos + 1 < (int64_t)rg.start || r.endpos + 1 > (int64_t)rg.end) continue;         /* thread.rs:339-345 */
            int32_t asg = 0;
            bool put_hp = false, put_ps = false;
            uint32_t psv = 0;
            auto ih = first_hp.find(r.qname);
            if (ih != first_hp.end()) { asg = hp[ih->second]; put_hp = asg != 0 && !r.has_hp; }
            auto ip = first_ps.find(r.qname);
            if (ip != first_ps.end()) { psv = ps[ip->second]; put_ps = !r.has_ps; }
            wr32(out, r.bs + (put_hp ? 7 : 0) + (put_ps ? 7 : 0));
            out.insert(out.end(), raw.begin() + r.off + 4, raw.begin() + r.off + 4 + r.bs);
            if (put_hp) { out.push_back('H'); out.push_back('P'); out.push_back('i'); wr32(out, (uint32_t)asg); }
            if (put_ps) { out.push_back('P'); out.push_back('S'); out.push_back('I'); wr32(out, psv); }
            ++written;
        }
    }
    if (n_written) *n_written = written;
    return bgzf_write(out_bam, out, n_threads) ? 0 : LCR_ERR_INVALID_ARG;
}

} // namespace lcrhost

extern "C" {

int lcr_host_pack_seq4(const lcr_reads *R, uint64_t *seq4_off, uint8_t *seq4, int n_threads) {
    if (!R || !seq4_off || (R->n_reads && !R->seq_off)) return LCR_ERR_INVALID_ARG;
    seq4_off[0] = 0;
    for (uint32_t i = 0; i < R->n_reads; ++i) seq4_off[i + 1] = seq4_off[i] + (R->seq_off[i + 1] - R->seq_off[i] + 1) / 2;
    if (!seq4 || !R->n_reads) return 0;
    if (!R->seq) return LCR_ERR_INVALID_ARG;
    uint8_t code[256];
    memset(code, 15, sizeof code);
    const char *nib = "=ACMGRSVTWYHKDBN";
    for (int k = 0; k < 16; ++k) code[(uint8_t)nib[k]] = (uint8_t)k;
    std::atomic<uint32_t> next{0};
    auto work = [&]() {
        for (;;) {
            const uint32_t i0 = next.fetch_add(1024);
            if (i0 >= R->n_reads) break;
            const uint32_t i1 = std::min<uint32_t>(R->n_reads, i0 + 1024);
            for (uint32_t i = i0; i < i1; ++i) {
                const uint8_t *s = R->seq + R->seq_off[i];
                const uint64_t l = R->seq_off[i + 1] - R->seq_off[i];
                uint8_t *o = seq4 + seq4_off[i];
                for (uint64_t k = 0; k + 1 < l; k += 2) o[k >> 1] = (uint8_t)(code[s[k]] << 4 | code[s[k + 1]]);
                if (l & 1) o[l >> 1] = (uint8_t)(code[s[l - 1]] << 4);
            }
        }
    };
    if (n_threads <= 1) work();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < n_threads; ++i) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
    return 0;
}

int lcr_host_write_bam(const char *path, const lcr_reads *reads, const char *header_text, int n_threads) {
    if (!path || !reads) return LCR_ERR_INVALID_ARG;
    return lcrhost::write_bam(path, *reads, header_text, n_threads);
}

int lcr_host_write_phased_bam(const char *in_bam, const char *out_bam, const lcr_region *regions, uint32_t n_regions, const int8_t *hp, const uint32_t *ps, const uint8_t *has_entry,
                              uint32_t n_reads, int n_threads, uint64_t *n_written) {
    if (!in_bam || !out_bam || (n_regions && !regions) || (n_reads && (!hp || !ps))) return LCR_ERR_INVALID_ARG;
    return lcrhost::write_phased_bam(in_bam, out_bam, regions, n_regions, hp, ps, has_entry, n_reads, n_threads, n_written);
}

} /* extern "C" */
