/*
 * longcallr_b200.h — C ABI of the B200-native SNP-calling / read-phasing engine.
 *
 * The reference (huangnengCSU/longcallR v1.12.0) has no FFI: its boundary is the
 * body of the per-region rayon worker, src/thread.rs:78-221, invoked once per
 * isolated region by `isolated_regions.par_iter().for_each` (src/thread.rs:76-77).
 * This header is what that closure would bind with an `extern "C"` block (see
 * INTEGRATION.md): the Rust (or C++) host decodes BAM records into flat arrays,
 * one call processes a batch of regions entirely on the device, and the host
 * formats VCF / tags the BAM from the returned records with the reference's own
 * src/vcf.rs / src/thread.rs:224-361 logic unchanged.
 *
 * Plain C: pointers and sizes only, little-endian, no ownership transfer except
 * `lcr_result` (library-owned until lcr_free_result).
 * All entry points return 0 on success or a negative lcr_status; nothing unwinds
 * across the boundary (reference panics become status codes).
 */
#ifndef LONGCALLR_B200_H
#define LONGCALLR_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCR_ABI_VERSION 3

typedef enum lcr_status {
    LCR_OK = 0,
    LCR_ERR_INVALID_ARG = -1,
    LCR_ERR_CUDA = -2,          /* sticky per context; lcr_last_error() has the CUDA string */
    LCR_ERR_NO_DEVICE = -3,     /* no CUDA device: there is no CPU fallback on this path */
    LCR_ERR_OOM = -4,
    LCR_ERR_BAD_CIGAR = -5,     /* reference: panic "unknown cigar operation" util.rs:944, fragment.rs:191 */
    LCR_ERR_NO_REFERENCE = -6,  /* region on a contig never given to lcr_set_reference (thread.rs:79 unwrap) */
    LCR_ERR_BASEQ_ZERO = -7,    /* base quality 0 at a phase site: reference panics on NaN, phase.rs:307 */
    LCR_ERR_INTERNAL = -8,      /* a device-side invariant did not hold (a bug here, not in the input) */
    LCR_REGION_NO_EXON = 1      /* region_status only, not an error: --exon-only and no exon of the region's genes (thread.rs:88-91 returns early) */
} lcr_status;

/* Scalar parameters of the worker, src/thread.rs:17-51; defaults per preset in
   src/main.rs:272-396 (see lcr_params_preset). */
typedef struct lcr_params {
    int32_t platform;                     /* 0 = Hifi, 1 = Ont            main.rs:33-37            */
    int32_t min_mapq;                     /* thread.rs:28                                          */
    int32_t min_baseq;                    /* thread.rs:29                                          */
    int32_t min_read_length;              /* thread.rs:39                                          */
    float divergence;                     /* thread.rs:30                                          */
    float min_allele_freq;                /* thread.rs:31                                          */
    float min_allele_freq_include_intron; /* thread.rs:33                                          */
    uint32_t min_qual;                    /* thread.rs:32                                          */
    int32_t use_strand_bias;              /* thread.rs:34                                          */
    uint32_t min_depth;                   /* thread.rs:35                                          */
    uint32_t max_depth;                   /* thread.rs:36                                          */
    uint32_t distance_to_read_end;        /* thread.rs:40                                          */
    uint32_t polya_tail_length;           /* thread.rs:41                                          */
    uint32_t dense_win_size;              /* thread.rs:42                                          */
    uint32_t min_dense_cnt;               /* thread.rs:43                                          */
    uint32_t min_linkers;                 /* thread.rs:44                                          */
    float min_phase_score;                /* thread.rs:45                                          */
    uint32_t max_enum_snps;               /* thread.rs:46                                          */
    double read_assignment_cutoff;        /* thread.rs:47                                          */
    float low_allele_frac_cutoff;         /* thread.rs:49                                          */
    uint32_t low_allele_cnt_cutoff;       /* thread.rs:50                                          */
    uint32_t ld_weight_threshold;         /* hard-coded 1 at thread.rs:166                         */
    uint32_t flags;                       /* LCR_FLAG_*                                            */
    uint64_t seed;                        /* seeds lcr_uniform(); replaces thread_rng              */
    uint32_t downsample_depth;            /* thread.rs:38, default 10000; used with LCR_FLAG_DOWNSAMPLE */
    uint32_t reserved0;
} lcr_params;

#define LCR_FLAG_EMIT_PLANES 1u  /* also return the per-position pileup counters (debug / parity) */
#define LCR_FLAG_SKIP_PHASING 2u /* stop after candidate calling (lcr_pileup_genotype semantics)  */
#define LCR_FLAG_EMIT_FRAGMENTS 4u /* also return the read x SNP fragment matrix (debug / parity)  */
/* Base qualities are read at candidate sites only (genotype likelihood, fragment elements): with this flag a `qual` array that
   lies in page-locked, device-mapped host memory (cudaHostAlloc / cudaHostRegister, lcr_pin_host) is not copied; the kernels
   fetch the bytes they need over the bus.  A pageable `qual` is copied as usual.  Results are identical either way. */
#define LCR_FLAG_QUAL_ON_DEMAND 8u
/* --downsample (thread.rs:144-151, phase.rs:693-701): in a region with at least downsample_depth fragments only downsample_depth of
   them, chosen by the reference's seeded shuffle (StdRng::seed_from_u64(2025) = ChaCha12, SliceRandom::shuffle of rand 0.8.5, restated
   in include/lcr_contract.h), take part in phasing, the first two assignment rounds and the rescue passes */
#define LCR_FLAG_DOWNSAMPLE 16u

enum { LCR_PRESET_ONT_CDNA = 0, LCR_PRESET_ONT_DRNA = 1, LCR_PRESET_HIFI_ISOSEQ = 2, LCR_PRESET_HIFI_MASSEQ = 3 };

/* One isolated region, util.rs:21-32: start 1-based inclusive, end 1-based exclusive.
   Its reads are rows [read_begin, read_end) of the batch, in BAM (coordinate) order,
   and must include every record htslib `fetch((chr,start,end))` would return
   (util.rs:636-638); supersets are fine, the device applies the read filter and the
   window test itself. */
typedef struct lcr_region {
    int32_t tid;
    uint32_t start;
    uint32_t end;
    uint32_t read_begin;
    uint32_t read_end;
} lcr_region;

/* Decoded alignments, struct-of-arrays, host-owned. */
typedef struct lcr_batch {
    uint32_t n_regions;
    uint32_t n_reads;
    const lcr_region *regions;
    const int32_t *pos;      /* [n_reads] 0-based leftmost position (record.pos())                 */
    const uint16_t *flag;    /* [n_reads] BAM flag                                                  */
    const uint8_t *mapq;     /* [n_reads]                                                           */
    const int8_t *ts;        /* [n_reads] value of the ts:A tag: '+', '-' or '*' when absent (util.rs:673-679) */
    const float *de;         /* [n_reads] de:f tag, NaN when absent or not of type f (util.rs:661-668) */
    const uint64_t *seq_off; /* [n_reads+1] offsets into seq/qual; l_seq = seq_off[i+1]-seq_off[i]  */
    const uint64_t *cig_off; /* [n_reads+1] offsets into cigar                                      */
    const uint8_t *seq;      /* ASCII as rust-htslib decodes nibbles: "=ACMGRSVTWYHKDBN"            */
    const uint8_t *qual;     /* raw phred, uncapped                                                 */
    const uint32_t *cigar;   /* BAM encoding: len<<4 | op, op in MIDNSHP=X                          */
    /* ABI 3: bases as the BAM record stores them (what record.seq() wraps before rust-htslib decodes it, util.rs:693): two bases
       per byte, high nibble first, read i at seq4[seq4_off[i] .. seq4_off[i] + (l_seq + 1) / 2).  When seq4 is not NULL it is
       used instead of seq (which may then be NULL) and expanded to the same ASCII letters on the device: half the bytes over the bus. */
    const uint8_t *seq4;
    const uint64_t *seq4_off; /* [n_reads+1] */
    /* --exon-only (candidate.rs:80-89, thread.rs:80-91): the exon (CDS) intervals of the genes of every region, as parse_annotation
       keeps them (util.rs:435-439): start 1-based inclusive, stop = end + 1, any order, overlaps allowed.  A position is a candidate only
       if some interval holds it; a region without intervals is skipped (region_status LCR_REGION_NO_EXON).  NULL exon_off: no mask. */
    const uint32_t *exon_off; /* [n_regions+1] offsets into exon_iv, in intervals */
    const uint32_t *exon_iv;  /* [2 * exon_off[n_regions]] start, stop pairs */
    /* -v / --input-vcf (thread.rs:107-116, candidate.rs:530-613): candidates are imported instead of called.  Per region the VCF
       records that fall inside it: position (0-based, ascending, one record per position: the reference keeps them in a map),
       genotype class (vcf.rs:443-449: 0 = 0/0, 1 = 0/1, 2 = 1/1, 3 = 1/2, 4 = anything else) and QUAL (NaN when missing).
       Alleles, frequencies and depth still come from the pileup; no count filter, likelihood, dense filter or exon mask applies.
       NULL ext_off: candidates are called from the pileup. */
    const uint32_t *ext_off;  /* [n_regions+1] offsets into the three arrays below */
    const uint32_t *ext_pos;
    const uint8_t *ext_gt;
    const float *ext_qual;
} lcr_batch;

/* candidate flags (snp.rs:66-84) */
#define LCR_CF_RNA_EDITING 0x0001u
#define LCR_CF_DENSE 0x0002u
#define LCR_CF_HET_VAR 0x0004u
#define LCR_CF_FOR_PHASING 0x0008u
#define LCR_CF_HOM_VAR 0x0010u
#define LCR_CF_SINGLE 0x0020u
#define LCR_CF_NON_SELECTED 0x0040u
#define LCR_CF_CAND_SOMATIC 0x0080u
#define LCR_CF_EDIT_LIST 0x0100u    /* member of SNPFrag.edit_snps (candidate.rs:391,404)    */
#define LCR_CF_SOMATIC_LIST 0x0200u /* member of SNPFrag.somatic_snps (candidate.rs:414)     */

/* One CandidateSNP (snp.rs:39-90) after the whole worker body ran; everything
   src/vcf.rs:27-306 reads. 88 bytes. */
typedef struct lcr_candidate {
    int64_t pos;                    /* 0-based                                  */
    double variant_quality;         /* QUAL before `as i32`                     */
    double genotype_quality;        /* GQ before `as i32`                       */
    double phase_score;             /* PQ                                       */
    double genotype_probability[3]; /* homvar, het, homref (candidate.rs:324)   */
    float allele_freqs[2];
    uint32_t depth;
    uint32_t phase_set;             /* 0 = none                                 */
    uint8_t reference;
    uint8_t alleles[2];
    int8_t variant_type;            /* 0 homref, 1 het, 2 homvar, 3 triallelic  */
    int8_t genotype;                /* eta                                      */
    int8_t haplotype;               /* delta                                    */
    uint16_t flags;                 /* LCR_CF_*                                 */
    uint32_t region;                /* index into batch->regions                */
    uint32_t reserved;
} lcr_candidate;

/* Per-position pileup counters (BaseFreq, util.rs:100-127, live fields only). */
typedef struct lcr_planes {
    uint64_t n_pos;
    const uint64_t *pos_off; /* [n_regions+1] first plane index of each region */
    const uint32_t *acgt;    /* [n_pos][4] a,c,g,t                                   */
    const uint32_t *fwd;     /* [n_pos][4] forward-strand count of a,c,g,t           */
    const uint32_t *d;       /* [n_pos] deletions                                    */
    const uint32_t *n;       /* [n_pos] introns                                      */
    const uint32_t *ts;      /* [n_pos][2] transcript strand forward / reverse       */
} lcr_planes;

/* The fragment matrix as built by SNPFrag::get_fragments (fragment.rs:10-309), CSR by
   fragment; one row per read that became a Fragment, in BAM order within its region. */
typedef struct lcr_fragments {
    uint64_t n_frag;
    uint64_t n_elem;
    const uint32_t *frag_off;  /* [n_regions+1] first fragment of each region                  */
    const uint32_t *frag_read; /* [n_frag] read index in the batch                             */
    const uint64_t *elem_off;  /* [n_frag+1] kept FragElems (fragment.rs:148-152) of each row  */
    const uint32_t *elem_snp;  /* [n_elem] FragElem.snp_idx: candidate index within the region */
    const int8_t *elem_cell;   /* [n_elem] FragElem.p * (FragElem.baseq + 1)                   */
    const uint8_t *elem_base;  /* [n_elem] FragElem.base                                       */
} lcr_fragments;

typedef struct lcr_stats {
    uint64_t n_reads_pass;       /* reads passing the read filter and window test             */
    uint64_t n_aligned_bases;    /* N_al: M/=/X bases inside their region (masked ones count) */
    uint64_t n_positions;        /* sum of region lengths                                     */
    uint64_t n_candidates;
    uint64_t n_fragments;
    uint64_t nnz_phase;          /* phase-site alleles in fragments used for phasing          */
    uint64_t n_cross_optimize;   /* cross_optimize calls (phase.rs:810)                       */
    uint64_t n_sweep_iters;      /* total iterations of its while loop                        */
} lcr_stats;

/* Library-owned result of one lcr_submit. */
typedef struct lcr_result {
    uint32_t n_regions;
    uint32_t n_reads;
    uint32_t n_cand;
    uint32_t reserved;
    const uint32_t *cand_off;      /* [n_regions+1] candidates of region r, position-sorted */
    const lcr_candidate *cand;     /* [n_cand]                                              */
    const int32_t *region_status;  /* [n_regions] 0 or a negative lcr_status                */
    const int8_t *hp;              /* [n_reads] read_assignments value (0/1/2); -1 = read has no entry (thread.rs:181) */
    const uint32_t *ps;            /* [n_reads] phase set (thread.rs:201), 0 = no entry     */
    const uint8_t *is_fragment;    /* [n_reads] read became a Fragment (fragment.rs:82-84)  */
    lcr_planes planes;             /* only with LCR_FLAG_EMIT_PLANES                        */
    lcr_fragments fragments;       /* only with LCR_FLAG_EMIT_FRAGMENTS                     */
    lcr_stats stats;
} lcr_result;

typedef struct lcr_ctx lcr_ctx;

/* fills `p` with the defaults of src/main.rs:272-396 for one of LCR_PRESET_* */
int lcr_params_preset(int preset, lcr_params *p);

/* create / destroy a context bound to one CUDA device */
int lcr_create(const lcr_params *p, int device, lcr_ctx **out);
void lcr_destroy(lcr_ctx *ctx);

/* upload one contig (thread.rs:59,79: load_reference / ref_seqs.get(chr)); bytes as in the FASTA, case preserved */
int lcr_set_reference(lcr_ctx *ctx, int32_t tid, const uint8_t *seq, uint64_t len);

/* the worker body, thread.rs:78-221, for every region of the batch; blocking; host buffers (pinned memory lets the copies run
   asynchronously: large batches are cut into chunks of consecutive regions whose uploads overlap the previous chunk's kernels).
   Every entry point takes the context's mutex: concurrent callers on one context are safe and serialised. */
int lcr_submit(lcr_ctx *ctx, const lcr_batch *batch, lcr_result **out);
void lcr_free_result(lcr_result *res);

/* split form used by bench.py's device-resident timing:
     lcr_upload      copies a batch to HBM (returns a handle),
     lcr_run_device  runs the whole path on the resident batch and leaves results on the device,
     lcr_fetch       copies the results of the last run to the host.
   lcr_submit == upload + run_device + fetch + release, per chunk of regions, with the results merged in batch order. */
typedef struct lcr_device_batch lcr_device_batch;
int lcr_upload(lcr_ctx *ctx, const lcr_batch *batch, lcr_device_batch **out);
int lcr_run_device(lcr_ctx *ctx, lcr_device_batch *db);
int lcr_fetch(lcr_ctx *ctx, lcr_device_batch *db, lcr_result **out);
void lcr_release(lcr_ctx *ctx, lcr_device_batch *db);

/* device-resident results of the last lcr_run_device on this batch (valid until the next run or lcr_release): what a
   caller that keeps results on the GPU reads instead of lcr_fetch, e.g. the per-rank gather of VCF records over NCCL */
typedef struct lcr_device_view {
    const lcr_candidate *cand; /* [n_cand] device pointer, sorted by (region, pos) */
    const int8_t *hp;          /* [n_reads] device pointer                          */
    const uint32_t *ps;        /* [n_reads] device pointer                          */
    uint32_t n_cand;
    uint32_t n_reads;
} lcr_device_view;
int lcr_device_results(lcr_ctx *ctx, lcr_device_batch *db, lcr_device_view *out);

/* Isolated-region discovery on the device: replaces find_isolated_regions_with_depth (util.rs:236-332, called per contig
   from util.rs:558-602), i.e. the reference's first BAM pass.  Input is the per-read columns that pass reads (no bases, no
   qualities); reads in BAM order (tid, pos).  The read filter uses the context's min_mapq / min_read_length / divergence
   (util.rs:262-279).  Output: regions sorted by (tid, start) with read_begin / read_end filled (ready for lcr_batch) and
   Region.max_coverage of each; both arrays are malloc'ed and released with lcr_free_regions.  ms (optional) receives the
   CUDA-event time of the device work, host-to-device copies excluded. */
typedef struct lcr_align_index {
    uint32_t n_reads;
    uint32_t n_contigs;
    const uint64_t *contig_lens; /* [n_contigs] */
    const int32_t *tid;          /* [n_reads] */
    const int32_t *pos;
    const uint16_t *flag;
    const uint8_t *mapq;
    const float *de;
    const uint64_t *seq_off;     /* [n_reads+1] only the differences (l_seq) are used */
    const uint64_t *cig_off;     /* [n_reads+1] */
    const uint32_t *cigar;
} lcr_align_index;
int lcr_discover_regions(lcr_ctx *ctx, const lcr_align_index *in, int truncation, uint32_t truncation_coverage,
                         lcr_region **regions, uint32_t **max_coverage, uint32_t *n_regions, float *ms);
void lcr_free_regions(lcr_region *regions, uint32_t *max_coverage);

/* timing / accounting of the last lcr_run_device on this batch */
typedef struct lcr_timing {
    float ms_total;           /* CUDA-event time of the whole run on the context stream   */
    float ms_pileup;          /* read/segment prep + tile pileup + genotype               */
    float ms_pileup_kernel;   /* the tile pileup kernel alone                             */
    float ms_fragments;       /* fragment-matrix build                                    */
    float ms_phase;           /* phasing sweeps + assignment + phase sets                 */
    uint32_t kernel_launches; /* kernels launched by the run                              */
    uint32_t reserved;
    uint64_t pileup_alg_bytes; /* algorithmic bytes of the tile pileup kernel  */
    uint64_t h2d_bytes;        /* bytes lcr_upload copied                                  */
    uint64_t d2h_bytes;        /* bytes lcr_fetch copied                                   */
    /* ABI 2: finer stage times (CUDA events on the context stream) and the counts behind the roofline figures */
    float ms_prep;            /* read filter / span pass, scans, CIGAR walk, tile descriptors      */
    float ms_enum;            /* enumeration work lists + search kernels (regions of <= 10 sites)  */
    float ms_phase_kernel;    /* k_phase / k_phase_grid: LD path, read / SNP assignment, phase sets */
    uint32_t run_attempts;    /* 1, or more when a capacity overflow made the run repeat           */
    uint64_t n_segments;      /* segments the walk emitted                                         */
    uint64_t n_items;         /* (read, tile) rows                                                 */
    uint64_t n_tiles;         /* tiles the pileup kernel processed                                 */
    uint64_t phase_alg_bytes; /* algorithmic bytes of the phasing sweeps: B_sweep x sweep iterations */
} lcr_timing;
int lcr_get_timing(lcr_ctx *ctx, lcr_device_batch *db, lcr_timing *out);
/* accounting of the last lcr_submit on this context, summed over the chunks it was cut into */
int lcr_last_submit_timing(lcr_ctx *ctx, lcr_timing *out);

/* page-lock (and map for the device) a host buffer the caller will hand over repeatedly: faster copies, and what
   LCR_FLAG_QUAL_ON_DEMAND needs for `qual` */
int lcr_pin_host(void *ptr, size_t bytes);
int lcr_unpin_host(void *ptr);

const char *lcr_strerror(int status);
const char *lcr_last_error(lcr_ctx *ctx);
int lcr_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LONGCALLR_B200_H */
