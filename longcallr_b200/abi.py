"""ctypes mirror of include/longcallr_b200.h and csrc/host/lcr_host.h.

Only layouts live here; no behaviour.  The same structures are used to call the
product library (liblongcallr_b200.so) and, from tests/, the CPU oracle.
"""
import ctypes as C

import numpy as np

LCR_ABI_VERSION = 3

LCR_OK = 0
LCR_ERR_INVALID_ARG = -1
LCR_ERR_CUDA = -2
LCR_ERR_NO_DEVICE = -3
LCR_ERR_OOM = -4
LCR_ERR_BAD_CIGAR = -5
LCR_ERR_NO_REFERENCE = -6
LCR_ERR_BASEQ_ZERO = -7
LCR_ERR_INTERNAL = -8
LCR_REGION_NO_EXON = 1

LCR_FLAG_EMIT_PLANES = 1
LCR_FLAG_SKIP_PHASING = 2
LCR_FLAG_EMIT_FRAGMENTS = 4
LCR_FLAG_QUAL_ON_DEMAND = 8
LCR_FLAG_DOWNSAMPLE = 16

PRESETS = {"ont-cdna": 0, "ont-drna": 1, "hifi-isoseq": 2, "hifi-masseq": 3}

CF_RNA_EDITING = 0x0001
CF_DENSE = 0x0002
CF_HET_VAR = 0x0004
CF_FOR_PHASING = 0x0008
CF_HOM_VAR = 0x0010
CF_SINGLE = 0x0020
CF_NON_SELECTED = 0x0040
CF_CAND_SOMATIC = 0x0080
CF_EDIT_LIST = 0x0100
CF_SOMATIC_LIST = 0x0200


class Params(C.Structure):
    _fields_ = [
        ("platform", C.c_int32),
        ("min_mapq", C.c_int32),
        ("min_baseq", C.c_int32),
        ("min_read_length", C.c_int32),
        ("divergence", C.c_float),
        ("min_allele_freq", C.c_float),
        ("min_allele_freq_include_intron", C.c_float),
        ("min_qual", C.c_uint32),
        ("use_strand_bias", C.c_int32),
        ("min_depth", C.c_uint32),
        ("max_depth", C.c_uint32),
        ("distance_to_read_end", C.c_uint32),
        ("polya_tail_length", C.c_uint32),
        ("dense_win_size", C.c_uint32),
        ("min_dense_cnt", C.c_uint32),
        ("min_linkers", C.c_uint32),
        ("min_phase_score", C.c_float),
        ("max_enum_snps", C.c_uint32),
        ("read_assignment_cutoff", C.c_double),
        ("low_allele_frac_cutoff", C.c_float),
        ("low_allele_cnt_cutoff", C.c_uint32),
        ("ld_weight_threshold", C.c_uint32),
        ("flags", C.c_uint32),
        ("seed", C.c_uint64),
        ("downsample_depth", C.c_uint32),
        ("reserved0", C.c_uint32),
    ]


class Region(C.Structure):
    _fields_ = [
        ("tid", C.c_int32),
        ("start", C.c_uint32),
        ("end", C.c_uint32),
        ("read_begin", C.c_uint32),
        ("read_end", C.c_uint32),
    ]


REGION_DTYPE = np.dtype(
    [("tid", "<i4"), ("start", "<u4"), ("end", "<u4"), ("read_begin", "<u4"), ("read_end", "<u4")]
)


class Batch(C.Structure):
    _fields_ = [
        ("n_regions", C.c_uint32),
        ("n_reads", C.c_uint32),
        ("regions", C.c_void_p),
        ("pos", C.c_void_p),
        ("flag", C.c_void_p),
        ("mapq", C.c_void_p),
        ("ts", C.c_void_p),
        ("de", C.c_void_p),
        ("seq_off", C.c_void_p),
        ("cig_off", C.c_void_p),
        ("seq", C.c_void_p),
        ("qual", C.c_void_p),
        ("cigar", C.c_void_p),
        ("seq4", C.c_void_p),
        ("seq4_off", C.c_void_p),
        ("exon_off", C.c_void_p),
        ("exon_iv", C.c_void_p),
        ("ext_off", C.c_void_p),
        ("ext_pos", C.c_void_p),
        ("ext_gt", C.c_void_p),
        ("ext_qual", C.c_void_p),
    ]


class Candidate(C.Structure):
    _fields_ = [
        ("pos", C.c_int64),
        ("variant_quality", C.c_double),
        ("genotype_quality", C.c_double),
        ("phase_score", C.c_double),
        ("genotype_probability", C.c_double * 3),
        ("allele_freqs", C.c_float * 2),
        ("depth", C.c_uint32),
        ("phase_set", C.c_uint32),
        ("reference", C.c_uint8),
        ("alleles", C.c_uint8 * 2),
        ("variant_type", C.c_int8),
        ("genotype", C.c_int8),
        ("haplotype", C.c_int8),
        ("flags", C.c_uint16),
        ("region", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


CANDIDATE_DTYPE = np.dtype(
    [
        ("pos", "<i8"),
        ("variant_quality", "<f8"),
        ("genotype_quality", "<f8"),
        ("phase_score", "<f8"),
        ("genotype_probability", "<f8", (3,)),
        ("allele_freqs", "<f4", (2,)),
        ("depth", "<u4"),
        ("phase_set", "<u4"),
        ("reference", "u1"),
        ("alleles", "u1", (2,)),
        ("variant_type", "i1"),
        ("genotype", "i1"),
        ("haplotype", "i1"),
        ("flags", "<u2"),
        ("region", "<u4"),
        ("reserved", "<u4"),
    ]
)
assert CANDIDATE_DTYPE.itemsize == C.sizeof(Candidate) == 88


class Planes(C.Structure):
    _fields_ = [
        ("n_pos", C.c_uint64),
        ("pos_off", C.c_void_p),
        ("acgt", C.c_void_p),
        ("fwd", C.c_void_p),
        ("d", C.c_void_p),
        ("n", C.c_void_p),
        ("ts", C.c_void_p),
    ]


class Fragments(C.Structure):
    _fields_ = [
        ("n_frag", C.c_uint64),
        ("n_elem", C.c_uint64),
        ("frag_off", C.c_void_p),
        ("frag_read", C.c_void_p),
        ("elem_off", C.c_void_p),
        ("elem_snp", C.c_void_p),
        ("elem_cell", C.c_void_p),
        ("elem_base", C.c_void_p),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_reads_pass", C.c_uint64),
        ("n_aligned_bases", C.c_uint64),
        ("n_positions", C.c_uint64),
        ("n_candidates", C.c_uint64),
        ("n_fragments", C.c_uint64),
        ("nnz_phase", C.c_uint64),
        ("n_cross_optimize", C.c_uint64),
        ("n_sweep_iters", C.c_uint64),
    ]


class Result(C.Structure):
    _fields_ = [
        ("n_regions", C.c_uint32),
        ("n_reads", C.c_uint32),
        ("n_cand", C.c_uint32),
        ("reserved", C.c_uint32),
        ("cand_off", C.c_void_p),
        ("cand", C.c_void_p),
        ("region_status", C.c_void_p),
        ("hp", C.c_void_p),
        ("ps", C.c_void_p),
        ("is_fragment", C.c_void_p),
        ("planes", Planes),
        ("fragments", Fragments),
        ("stats", Stats),
    ]


class Timing(C.Structure):
    _fields_ = [
        ("ms_total", C.c_float),
        ("ms_pileup", C.c_float),
        ("ms_pileup_kernel", C.c_float),
        ("ms_fragments", C.c_float),
        ("ms_phase", C.c_float),
        ("kernel_launches", C.c_uint32),
        ("reserved", C.c_uint32),
        ("pileup_alg_bytes", C.c_uint64),
        ("h2d_bytes", C.c_uint64),
        ("d2h_bytes", C.c_uint64),
        ("ms_prep", C.c_float),
        ("ms_enum", C.c_float),
        ("ms_phase_kernel", C.c_float),
        ("run_attempts", C.c_uint32),
        ("n_segments", C.c_uint64),
        ("n_items", C.c_uint64),
        ("n_tiles", C.c_uint64),
        ("phase_alg_bytes", C.c_uint64),
    ]


class DeviceView(C.Structure):
    _fields_ = [("cand", C.c_void_p), ("hp", C.c_void_p), ("ps", C.c_void_p), ("n_cand", C.c_uint32), ("n_reads", C.c_uint32)]


class AlignIndex(C.Structure):
    """lcr_align_index: the per-read columns region discovery reads."""

    _fields_ = [("n_reads", C.c_uint32), ("n_contigs", C.c_uint32), ("contig_lens", C.c_void_p), ("tid", C.c_void_p), ("pos", C.c_void_p), ("flag", C.c_void_p),
                ("mapq", C.c_void_p), ("de", C.c_void_p), ("seq_off", C.c_void_p), ("cig_off", C.c_void_p), ("cigar", C.c_void_p)]


# ---- csrc/host/lcr_host.h ----


class Reads(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint32),
        ("n_contigs", C.c_uint32),
        ("contig_names", C.POINTER(C.c_char_p)),
        ("contig_lens", C.c_void_p),
        ("tid", C.c_void_p),
        ("pos", C.c_void_p),
        ("flag", C.c_void_p),
        ("mapq", C.c_void_p),
        ("ts", C.c_void_p),
        ("de", C.c_void_p),
        ("seq_off", C.c_void_p),
        ("cig_off", C.c_void_p),
        ("seq", C.c_void_p),
        ("qual", C.c_void_p),
        ("cigar", C.c_void_p),
        ("qname_off", C.c_void_p),
        ("qnames", C.c_void_p),
    ]


class Fasta(C.Structure):
    _fields_ = [
        ("n_contigs", C.c_uint32),
        ("names", C.POINTER(C.c_char_p)),
        ("lens", C.c_void_p),
        ("seqs", C.POINTER(C.c_void_p)),
    ]


class RegionList(C.Structure):
    _fields_ = [("n_regions", C.c_uint32), ("regions", C.c_void_p), ("max_coverage", C.c_void_p)]


class SynthConfig(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64),
        ("contig_len", C.c_uint64),
        ("n_contigs", C.c_uint32),
        ("platform", C.c_uint32),
        ("depth", C.c_float),
        ("n_het", C.c_uint32),
        ("n_edit", C.c_uint32),
        ("max_exons", C.c_uint32),
        ("max_intron", C.c_uint32),
        ("max_gap", C.c_uint32),
        ("both_strands", C.c_uint32),
        ("single_region", C.c_uint32),
        ("n_threads", C.c_uint32),
    ]


class Synth(C.Structure):
    _fields_ = [
        ("reads", C.POINTER(Reads)),
        ("fasta", C.POINTER(Fasta)),
        ("n_het_total", C.c_uint32),
        ("het_tid", C.c_void_p),
        ("het_pos", C.c_void_p),
        ("het_alt", C.c_void_p),
        ("het_hap", C.c_void_p),
        ("read_hap", C.c_void_p),
    ]


def as_array(ptr, dtype, shape):
    """numpy view (no copy) of library-owned memory; an empty array for NULL / zero length."""
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    n = int(np.prod(shape)) if shape else 1
    dtype = np.dtype(dtype)
    if not ptr or n == 0:
        return np.zeros(shape, dtype=dtype)
    buf = (C.c_uint8 * (n * dtype.itemsize)).from_address(int(ptr))
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
