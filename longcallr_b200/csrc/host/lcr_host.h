/*
 * lcr_host.h — host-side plumbing above the C ABI (C++17, no CUDA).
 *
 * Stands where longcallR's Rust host stands (cargo/rustc are not available in
 * this image): it decodes BAM/FASTA into the flat arrays of lcr_batch, finds the
 * isolated regions, and formats VCF records from lcr_result.
 *
 *   BAM decode            what rust-htslib hands to src/util.rs:650-683, src/fragment.rs:28-59
 *   FASTA load            src/util.rs:214-222  (load_reference)
 *   isolated regions      src/util.rs:236-332  (find_isolated_regions_with_depth)
 *   VCF text              src/vcf.rs:27-306 + src/thread.rs:224-305
 *   synthetic alignments  SURVEY.md section 8(d) generator (bench / tests input)
 *
 * Everything is exported with C linkage so tests and bench.py can drive it through
 * ctypes; buffers are owned by the returned handle until the matching *_free.
 */
#ifndef LCR_HOST_H
#define LCR_HOST_H

#include <stdint.h>
#include "longcallr_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* decoded alignments of one BAM (or one synthetic data set), coordinate-sorted */
typedef struct lcr_reads {
    uint32_t n_reads;
    uint32_t n_contigs;
    const char *const *contig_names; /* [n_contigs] */
    const uint64_t *contig_lens;     /* [n_contigs] */
    const int32_t *tid;              /* [n_reads] */
    const int32_t *pos;
    const uint16_t *flag;
    const uint8_t *mapq;
    const int8_t *ts;
    const float *de;
    const uint64_t *seq_off; /* [n_reads+1] */
    const uint64_t *cig_off; /* [n_reads+1] */
    const uint8_t *seq;
    const uint8_t *qual;
    const uint32_t *cigar;
    const uint64_t *qname_off; /* [n_reads+1] into qnames (no terminators) */
    const char *qnames;
} lcr_reads;

/* reference contigs */
typedef struct lcr_fasta {
    uint32_t n_contigs;
    const char *const *names;
    const uint64_t *lens;
    const uint8_t *const *seqs; /* bytes as in the file, case preserved */
} lcr_fasta;

int lcr_host_read_bam(const char *path, int n_threads, lcr_reads **out);
void lcr_host_free_reads(lcr_reads *r);
int lcr_host_read_fasta(const char *path, lcr_fasta **out);
void lcr_host_free_fasta(lcr_fasta *f);

/* isolated regions of all contigs (util.rs:236-332, 558-602), sorted by (tid, start);
   read_begin/read_end are filled so that lcr_batch{regions, reads} is ready to submit */
typedef struct lcr_region_list {
    uint32_t n_regions;
    const lcr_region *regions;
    const uint32_t *max_coverage; /* Region.max_coverage */
} lcr_region_list;
int lcr_host_find_regions(const lcr_reads *reads, const lcr_params *p, int truncation, uint32_t truncation_coverage,
                          lcr_region_list **out);
void lcr_host_free_regions(lcr_region_list *r);

/* VCF body lines (no header) of one result, in (region, position) order; contig_names[tid].
   Returns a NUL-terminated buffer to be released with lcr_host_free_text. */
int lcr_host_format_vcf(const lcr_result *res, const lcr_batch *batch, const char *const *contig_names,
                        float min_phase_score, char **out, uint64_t *out_len);
/* the header of src/thread.rs:225-263 */
int lcr_host_format_vcf_header(const char *const *contig_names, const uint64_t *contig_lens, uint32_t n_contigs,
                               char **out, uint64_t *out_len);
void lcr_host_free_text(char *t);

/* Bases in the BAM record's own form (two per byte, high nibble first, every read starting on a byte): fills seq4_off[n_reads+1]
   and, when seq4 is not NULL, the packed bytes (seq4_off[n_reads] of them; call once with NULL to size the buffer).
   Letters outside "=ACMGRSVTWYHKDBN" pack as N (15), as htslib does. */
int lcr_host_pack_seq4(const lcr_reads *reads, uint64_t *seq4_off, uint8_t *seq4, int n_threads);

/* ---- BAM output ---- */

/* A BAM file (BGZF, no index) holding `reads` as records: what tests and the bench use to put synthetic alignments
   through the real decode path.  header_text may be NULL (a minimal @HD/@SQ header is written).  Records carry
   qname, flag, mapq, CIGAR, 4-bit bases, qualities, ts:A (when not '*') and de:f (when not NaN); mate fields are unset. */
int lcr_host_write_bam(const char *path, const lcr_reads *reads, const char *header_text, int n_threads);

/* The phased BAM of thread.rs:307-361.  Records of in_bam are copied byte for byte, region by region in `regions`
   order, with the aux fields the reference pushes:
     - a record is written for a region when htslib's fetch((chr, start, end)) returns it (pos < end && endpos > start),
       it is mapped, primary and not supplementary (thread.rs:336-338), and lies fully inside the region:
       reference_start + 1 >= start && reference_end + 1 <= end (thread.rs:339-345); everything else is dropped;
     - HP:i (int32) when the read's assignment is 1 or 2, PS:I (uint32) when the read has a phase set
       (thread.rs:346-357); rust-htslib's push_aux refuses a tag that is already present and the reference ignores
       that error, so a record that already carries HP / PS keeps its old value;
     - assignments are looked up by QNAME with the first entry winning (thread.rs:308-325).  hp / ps are per record of
       in_bam (record i == read i of lcr_host_read_bam); for records sharing a QNAME the first record in file order that
       has an entry (has_entry[i] != 0; NULL: hp[i] != 0 resp. ps[i] != 0) provides the value for all of them - the
       deterministic stand-in for the reference's thread-order-dependent queue.
   n_written receives the number of records written. */
int lcr_host_write_phased_bam(const char *in_bam, const char *out_bam, const lcr_region *regions, uint32_t n_regions,
                              const int8_t *hp, const uint32_t *ps, const uint8_t *has_entry, uint32_t n_reads,
                              int n_threads, uint64_t *n_written);

/* ---- synthetic long-read RNA alignments ---- */
typedef struct lcr_synth_config {
    uint64_t seed;
    uint64_t contig_len;   /* bases per contig                                     */
    uint32_t n_contigs;
    uint32_t platform;     /* 0 HiFi, 1 ONT: quality and indel model               */
    float depth;           /* target depth over exonic bases                       */
    uint32_t n_het;        /* planted heterozygous SNPs per contig                 */
    uint32_t n_edit;       /* planted A>G editing sites per contig (30% fraction)  */
    uint32_t max_exons;    /* 1..max_exons exons per transcript                    */
    uint32_t max_intron;   /* intron length in [100, max_intron]                   */
    uint32_t max_gap;      /* zero-coverage gap between genes in [300, max_gap]    */
    uint32_t both_strands; /* 1: cDNA / IsoSeq (50/50 strands), 0: all forward     */
    uint32_t single_region;/* 1: one gap-free gene block per contig (phasing stress) */
    uint32_t n_threads;
} lcr_synth_config;

typedef struct lcr_synth {
    lcr_reads *reads;
    lcr_fasta *fasta;
    uint32_t n_het_total;
    const int32_t *het_tid; /* planted het SNP truth */
    const int64_t *het_pos;
    const uint8_t *het_alt;
    const int8_t *het_hap;  /* which haplotype (0/1) carries the alt allele */
    const int8_t *read_hap; /* [n_reads] haplotype each read was drawn from */
} lcr_synth;
int lcr_host_synth(const lcr_synth_config *cfg, lcr_synth **out);
void lcr_host_free_synth(lcr_synth *s);

#ifdef __cplusplus
}
#endif
#endif
