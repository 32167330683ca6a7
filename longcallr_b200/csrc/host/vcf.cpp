/*
 * vcf.cpp — VCF text from lcr_result records.
 *
 * Follows SNPFrag::output_phased_vcf (src/vcf.rs:27-306) for which records are
 * written and what each column holds, and src/thread.rs:225-305 for the header
 * and the line layout.  Integer columns use Rust's saturating `as i32`.
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "lcr_contract.h"
#include "lcr_host.h"

namespace {

struct Rec {
    bool emit = false;
    std::string alt, filter, info, format, genotype;
    int qual = 0;
};

void fmt2(std::string &s, double v) { /* Rust {:.2} */
    char b[64];
    snprintf(b, sizeof b, "%.2f", v);
    s += b;
}

/* first alternative rule used all over vcf.rs: the allele that is not the reference, allele 0 first */
bool single_alt(const lcr_candidate &c, std::string &alt, float &af0) {
    if (c.alleles[0] != c.reference) { alt.assign(1, (char)c.alleles[0]); af0 = c.allele_freqs[0]; return true; }
    if (c.alleles[1] != c.reference) { alt.assign(1, (char)c.alleles[1]); af0 = c.allele_freqs[1]; return true; }
    return false;
}

Rec format_one(const lcr_candidate &c, float min_phase_score) {
    Rec r;
    const bool dense = c.flags & LCR_CF_DENSE, non_selected = c.flags & LCR_CF_NON_SELECTED, rna = c.flags & LCR_CF_RNA_EDITING;
    float af[2] = {0, 0};
    std::string gt = "0/0";
    const int gq = lcr_f64_as_i32(c.genotype_quality);
    r.qual = lcr_f64_as_i32(c.variant_quality);
    auto both_alts = [&]() {
        r.alt.assign(1, (char)c.alleles[0]);
        r.alt += ',';
        r.alt += (char)c.alleles[1];
        af[0] = c.allele_freqs[0];
        af[1] = c.allele_freqs[1];
    };
    auto gt_unphased = [&](bool two) {
        r.genotype = gt + ":" + std::to_string(gq) + ":" + std::to_string(c.depth) + ":";
        fmt2(r.genotype, af[0]);
        if (two) { r.genotype += ','; fmt2(r.genotype, af[1]); }
        r.format = "GT:GQ:DP:AF";
    };
    if (dense) { /* vcf.rs:31-78 */
        if (c.variant_type == 1 || c.variant_type == 2) single_alt(c, r.alt, af[0]);
        else if (c.variant_type == 3) both_alts();
        r.filter = "dn";
        r.info = "RDS=dense_snp";
        if (c.variant_type == 1) gt = "0/1";
        else if (c.variant_type == 2) gt = "1/1";
        else if (c.variant_type == 3) gt = "1/2";
        else return r;
        gt_unphased(c.variant_type == 3);
        r.emit = true;
        return r;
    }
    if (non_selected) { /* vcf.rs:80-174 */
        if (rna) {
            if (c.variant_type == 1 || c.variant_type == 2) single_alt(c, r.alt, af[0]);
            else return r;
            r.filter = "RnaEdit";
            r.info = "RDS=noselect";
            gt = c.variant_type == 1 ? "0/1" : "1/1";
            gt_unphased(false);
            r.emit = true;
            return r;
        }
        if (c.variant_type == 0 || c.variant_type == 1 || c.variant_type == 2) {
            single_alt(c, r.alt, af[0]);
            if (c.variant_type == 0) { gt = "0/0"; r.filter = "HomRef"; }
            else if (c.variant_type == 1) { gt = "0/1"; r.filter = "LowQual"; }
            else { gt = "1/1"; r.filter = "PASS"; }
        } else {
            if (c.genotype == -1 || c.genotype == 1) {
                single_alt(c, r.alt, af[0]);
                if (c.genotype == -1) { gt = "1/1"; r.filter = "PASS"; }
                else { gt = "0/0"; r.filter = "HomRef"; }
            } else if (c.genotype == 0) {
                both_alts();
                gt = "1/2";
                r.filter = "Multiallelic";
            }
        }
        r.info = "RDS=noselect";
        gt_unphased(!(gt == "0/0" || gt == "0/1" || gt == "1/1"));
        r.emit = true;
        return r;
    }
    /* selected: vcf.rs:175-303 */
    if (c.phase_score >= (double)min_phase_score) {
        if (c.variant_type == 1) {
            single_alt(c, r.alt, af[0]);
            gt = c.haplotype == 1 ? "0|1" : "1|0";
            r.filter = "PASS";
        }
    } else {
        if (c.variant_type == 0) { single_alt(c, r.alt, af[0]); gt = "0/0"; r.filter = "HomRef"; }
        else if (c.variant_type == 1) { single_alt(c, r.alt, af[0]); gt = "0/1"; r.filter = "LowQual"; }
        else if (c.variant_type == 2) { single_alt(c, r.alt, af[0]); gt = "1/1"; r.filter = "PASS"; }
        else {
            if (c.genotype == -1 || c.genotype == 1) {
                single_alt(c, r.alt, af[0]);
                if (c.genotype == -1) { gt = "1/1"; r.filter = "PASS"; }
                else { gt = "0/0"; r.filter = "HomRef"; }
            } else if (c.genotype == 0) {
                both_alts();
                gt = "1/2";
                r.filter = "Multiallelic";
            }
        }
    }
    r.info = "RDS=select";
    const bool one = gt == "0/0" || gt == "0/1" || gt == "1/1" || gt == "0|1" || gt == "1|0";
    r.genotype = gt + ":" + std::to_string(gq) + ":" + (c.phase_set ? std::to_string(c.phase_set) : std::string(".")) + ":" + std::to_string(c.depth) + ":";
    fmt2(r.genotype, af[0]);
    if (!one) { r.genotype += ','; fmt2(r.genotype, af[1]); }
    r.genotype += ':';
    fmt2(r.genotype, c.phase_score);
    r.format = "GT:GQ:PS:DP:AF:PQ";
    r.emit = true;
    return r;
}

char *dup_text(const std::string &s, uint64_t *len) {
    char *p = (char *)malloc(s.size() + 1);
    if (!p) return nullptr;
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    if (len) *len = s.size();
    return p;
}

} // namespace

extern "C" {

int lcr_host_format_vcf(const lcr_result *res, const lcr_batch *batch, const char *const *contig_names, float min_phase_score,
                        char **out, uint64_t *out_len) {
    if (!res || !batch || !contig_names || !out) return LCR_ERR_INVALID_ARG;
    std::string text;
    for (uint32_t i = 0; i < res->n_cand; ++i) {
        const lcr_candidate &c = res->cand[i];
        Rec r = format_one(c, min_phase_score);
        /* thread.rs:266-304: only records with one or two ALT alleles are written */
        if (!r.emit || r.alt.empty()) continue;
        text += contig_names[batch->regions[c.region].tid];
        text += '\t';
        text += std::to_string((uint64_t)c.pos + 1);
        text += "\t.\t";
        text += (char)c.reference;
        text += '\t';
        text += r.alt;
        text += '\t';
        text += std::to_string(r.qual);
        text += '\t';
        text += r.filter;
        text += '\t';
        text += r.info;
        text += '\t';
        text += r.format;
        text += '\t';
        text += r.genotype;
        text += '\n';
    }
    *out = dup_text(text, out_len);
    return *out ? 0 : LCR_ERR_OOM;
}

int lcr_host_format_vcf_header(const char *const *contig_names, const uint64_t *contig_lens, uint32_t n_contigs, char **out,
                               uint64_t *out_len) {
    if (!out) return LCR_ERR_INVALID_ARG;
    std::string h;
    h += "##fileformat=VCFv4.3\n";
    for (uint32_t i = 0; i < n_contigs; ++i) {
        h += "##contig=<ID=";
        h += contig_names[i];
        h += ",length=" + std::to_string(contig_lens[i]) + ">\n";
    }
    h += "##FILTER=<ID=PASS,Description=\"All filters passed\">\n";
    h += "##FILTER=<ID=LowQual,Description=\"Low phasing quality\">\n";
    h += "##FILTER=<ID=HomRef,Description=\"Homo reference\">\n";
    h += "##FILTER=<ID=RnaEdit,Description=\"RNA editing\">\n";
    h += "##FILTER=<ID=Multiallelic,Description=\"Multiallelic SNP\">\n";
    h += "##FILTER=<ID=dn,Description=\"Dense cluster of variants\">\n";
    h += "##INFO=<ID=RDS,Number=1,Type=String,Description=\"RNA editing or Dense SNP or Single SNP.\">\n";
    h += "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n";
    h += "##FORMAT=<ID=PS,Number=1,Type=Integer,Description=\"Phase Set\">\n";
    h += "##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"Genotype Quality\">\n";
    h += "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"Read Depth\">\n";
    h += "##FORMAT=<ID=AF,Number=A,Type=Float,Description=\"Allele Frequency\">\n";
    h += "##FORMAT=<ID=PQ,Number=1,Type=Float,Description=\"Phasing Quality\">\n";
    h += "##FORMAT=<ID=AE,Number=A,Type=Integer,Description=\"Haplotype expression of two alleles\">\n";
    h += "##FORMAT=<ID=SQ,Number=1,Type=Float,Description=\"Somatic Score\">\n";
    h += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSample\n";
    *out = dup_text(h, out_len);
    return *out ? 0 : LCR_ERR_OOM;
}

} /* extern "C" */
