"""Python face of the host plumbing (csrc/host) and of the C ABI (include/longcallr_b200.h).

Mirrors the reference's per-region worker surface (src/thread.rs:17-51, 78-221):
`Engine.submit(batch)` is the worker body for a batch of regions, with the same
scalar parameters (`Params`, presets of src/main.rs:272-396) and the same outputs
(candidate records for src/vcf.rs, read -> HP, read -> PS).  ctypes only; no
torch types cross the boundary.  There is no CPU fallback: without the CUDA
library or without a device every compute entry point raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
HOST_LIB = os.path.join(LIB_DIR, "liblcr_host.so")
CUDA_LIB = os.path.join(LIB_DIR, "liblongcallr_b200.so")

_host = None
_cuda = None


class LcrError(RuntimeError):
    def __init__(self, status, msg=""):
        super().__init__(f"longcallr_b200 status {status}: {msg}")
        self.status = status


def host_lib():
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB):
            raise FileNotFoundError(f"{HOST_LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(HOST_LIB)
        L.lcr_host_read_bam.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.POINTER(abi.Reads))]
        L.lcr_host_free_reads.argtypes = [C.POINTER(abi.Reads)]
        L.lcr_host_free_reads.restype = None
        L.lcr_host_read_fasta.argtypes = [C.c_char_p, C.POINTER(C.POINTER(abi.Fasta))]
        L.lcr_host_free_fasta.argtypes = [C.POINTER(abi.Fasta)]
        L.lcr_host_free_fasta.restype = None
        L.lcr_host_find_regions.argtypes = [C.POINTER(abi.Reads), C.POINTER(abi.Params), C.c_int, C.c_uint32, C.POINTER(C.POINTER(abi.RegionList))]
        L.lcr_host_free_regions.argtypes = [C.POINTER(abi.RegionList)]
        L.lcr_host_free_regions.restype = None
        L.lcr_host_format_vcf.argtypes = [C.c_void_p, C.POINTER(abi.Batch), C.POINTER(C.c_char_p), C.c_float, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.lcr_host_format_vcf_header.argtypes = [C.POINTER(C.c_char_p), C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.lcr_host_free_text.argtypes = [C.c_void_p]
        L.lcr_host_free_text.restype = None
        L.lcr_host_synth.argtypes = [C.POINTER(abi.SynthConfig), C.POINTER(C.POINTER(abi.Synth))]
        L.lcr_host_free_synth.argtypes = [C.POINTER(abi.Synth)]
        L.lcr_host_free_synth.restype = None
        _host = L
    return _host


def cuda_lib():
    """The product library.  Loading works without a GPU (symbol checks); computing does not."""
    global _cuda
    if _cuda is None:
        if not os.path.exists(CUDA_LIB):
            raise FileNotFoundError(f"{CUDA_LIB} is missing: the CUDA extension was not built; there is no CPU fallback")
        L = C.CDLL(CUDA_LIB)
        L.lcr_params_preset.argtypes = [C.c_int, C.POINTER(abi.Params)]
        L.lcr_create.argtypes = [C.POINTER(abi.Params), C.c_int, C.POINTER(C.c_void_p)]
        L.lcr_destroy.argtypes = [C.c_void_p]
        L.lcr_destroy.restype = None
        L.lcr_set_reference.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64]
        L.lcr_submit.argtypes = [C.c_void_p, C.POINTER(abi.Batch), C.POINTER(C.POINTER(abi.Result))]
        L.lcr_free_result.argtypes = [C.POINTER(abi.Result)]
        L.lcr_free_result.restype = None
        L.lcr_upload.argtypes = [C.c_void_p, C.POINTER(abi.Batch), C.POINTER(C.c_void_p)]
        L.lcr_run_device.argtypes = [C.c_void_p, C.c_void_p]
        L.lcr_fetch.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(abi.Result))]
        L.lcr_release.argtypes = [C.c_void_p, C.c_void_p]
        L.lcr_release.restype = None
        L.lcr_get_timing.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(abi.Timing)]
        L.lcr_device_results.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(abi.DeviceView)]
        L.lcr_discover_regions.argtypes = [C.c_void_p, C.POINTER(abi.AlignIndex), C.c_int, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint32),
                                           C.POINTER(C.c_float)]
        L.lcr_free_regions.argtypes = [C.c_void_p, C.c_void_p]
        L.lcr_free_regions.restype = None
        L.lcr_last_submit_timing.argtypes = [C.c_void_p, C.POINTER(abi.Timing)]
        L.lcr_strerror.argtypes = [C.c_int]
        L.lcr_strerror.restype = C.c_char_p
        L.lcr_last_error.argtypes = [C.c_void_p]
        L.lcr_last_error.restype = C.c_char_p
        L.lcr_abi_version.restype = C.c_int
        _cuda = L
    return _cuda


def params_preset(name, **overrides):
    """Defaults of src/main.rs:272-396 for a preset name, then keyword overrides."""
    p = abi.Params()
    rc = cuda_lib().lcr_params_preset(abi.PRESETS[name], C.byref(p))
    if rc:
        raise LcrError(rc, "lcr_params_preset")
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class ReadSet:
    """Decoded alignments (library-owned arrays viewed as numpy)."""

    def __init__(self, ptr, owner=None, free=None):
        self._ptr = ptr
        self._owner = owner
        self._free = free
        r = ptr.contents
        n = r.n_reads
        self.n_reads = n
        self.contig_names = [r.contig_names[i].decode() for i in range(r.n_contigs)]
        self.contig_lens = abi.as_array(r.contig_lens, "<u8", r.n_contigs)
        self.tid = abi.as_array(r.tid, "<i4", n)
        self.pos = abi.as_array(r.pos, "<i4", n)
        self.flag = abi.as_array(r.flag, "<u2", n)
        self.mapq = abi.as_array(r.mapq, "u1", n)
        self.ts = abi.as_array(r.ts, "i1", n)
        self.de = abi.as_array(r.de, "<f4", n)
        self.seq_off = abi.as_array(r.seq_off, "<u8", n + 1)
        self.cig_off = abi.as_array(r.cig_off, "<u8", n + 1)
        nb = int(self.seq_off[-1]) if n else 0
        nc = int(self.cig_off[-1]) if n else 0
        self.seq = abi.as_array(r.seq, "u1", nb)
        self.qual = abi.as_array(r.qual, "u1", nb)
        self.cigar = abi.as_array(r.cigar, "<u4", nc)
        self.qname_off = abi.as_array(r.qname_off, "<u8", n + 1)
        self.qnames = abi.as_array(r.qnames, "u1", int(self.qname_off[-1]) if n else 0)

    @classmethod
    def from_bam(cls, path, threads=4):
        out = C.POINTER(abi.Reads)()
        rc = host_lib().lcr_host_read_bam(os.fsencode(path), threads, C.byref(out))
        if rc:
            raise LcrError(rc, f"cannot decode {path}")
        return cls(out, free=host_lib().lcr_host_free_reads)

    def qname(self, i):
        return bytes(self.qnames[int(self.qname_off[i]) : int(self.qname_off[i + 1])]).decode()

    def __del__(self):
        if getattr(self, "_free", None) and self._ptr:
            self._free(self._ptr)
            self._ptr = None


class ArrayReadSet:
    """A read set over caller-owned numpy arrays (same attributes as ReadSet); used for pinned staging buffers."""

    FIELDS = (("tid", "<i4"), ("pos", "<i4"), ("flag", "<u2"), ("mapq", "u1"), ("ts", "i1"), ("de", "<f4"),
              ("seq_off", "<u8"), ("cig_off", "<u8"), ("seq", "u1"), ("qual", "u1"), ("cigar", "<u4"))

    def __init__(self, contig_names, contig_lens, **arrays):
        self.contig_names = list(contig_names)
        self.contig_lens = np.asarray(contig_lens, dtype="<u8")
        for name, dt in self.FIELDS:
            setattr(self, name, np.ascontiguousarray(arrays[name], dtype=dt))
        self.n_reads = len(self.pos)

    @classmethod
    def like(cls, reads, alloc):
        """Copy `reads` into buffers obtained from alloc(nbytes) -> writable uint8 numpy array (e.g. pinned memory)."""
        arrays = {}
        for name, dt in cls.FIELDS:
            src = np.ascontiguousarray(getattr(reads, name), dtype=dt)
            buf = alloc(max(src.nbytes, 1))[: src.nbytes].view(dt)
            buf[...] = src
            arrays[name] = buf
        return cls(reads.contig_names, reads.contig_lens, **arrays)


class Reference:
    """FASTA contigs, bytes as in the file (src/util.rs:214-222)."""

    def __init__(self, ptr, owner=None, free=None):
        self._ptr = ptr
        self._owner = owner
        self._free = free
        f = ptr.contents
        self.names = [f.names[i].decode() for i in range(f.n_contigs)]
        self.lens = abi.as_array(f.lens, "<u8", f.n_contigs)
        self.seqs = [abi.as_array(f.seqs[i], "u1", int(self.lens[i])) for i in range(f.n_contigs)]

    @classmethod
    def from_fasta(cls, path):
        out = C.POINTER(abi.Fasta)()
        rc = host_lib().lcr_host_read_fasta(os.fsencode(path), C.byref(out))
        if rc:
            raise LcrError(rc, f"cannot read {path}")
        return cls(out, free=host_lib().lcr_host_free_fasta)

    def for_reads(self, reads):
        """Sequences indexed by the tid numbering of `reads` (None where the FASTA lacks the contig)."""
        idx = {n: i for i, n in enumerate(self.names)}
        return [self.seqs[idx[n]] if n in idx else None for n in reads.contig_names]

    def __del__(self):
        if getattr(self, "_free", None) and self._ptr:
            self._free(self._ptr)
            self._ptr = None


class Synthetic:
    """Seeded synthetic alignments + reference + planted truth (csrc/host/synth.cpp)."""

    def __init__(self, **kw):
        cfg = abi.SynthConfig()
        defaults = dict(seed=20251017, contig_len=1_000_000, n_contigs=1, platform=1, depth=30.0, n_het=1000, n_edit=0,
                        max_exons=6, max_intron=20000, max_gap=3000, both_strands=1, single_region=0, n_threads=os.cpu_count() or 1)
        defaults.update(kw)
        for k, v in defaults.items():
            setattr(cfg, k, v)
        self.config = defaults
        out = C.POINTER(abi.Synth)()
        rc = host_lib().lcr_host_synth(C.byref(cfg), C.byref(out))
        if rc:
            raise LcrError(rc, "lcr_host_synth")
        self._ptr = out
        s = out.contents
        self.reads = ReadSet(s.reads, owner=self)
        self.reference = Reference(s.fasta, owner=self)
        n = s.n_het_total
        self.het_tid = abi.as_array(s.het_tid, "<i4", n)
        self.het_pos = abi.as_array(s.het_pos, "<i8", n)
        self.het_alt = abi.as_array(s.het_alt, "u1", n)
        self.het_hap = abi.as_array(s.het_hap, "i1", n)
        self.read_hap = abi.as_array(s.read_hap, "i1", self.reads.n_reads)

    def __del__(self):
        if getattr(self, "_ptr", None):
            host_lib().lcr_host_free_synth(self._ptr)
            self._ptr = None


def reads_struct(reads):
    """An lcr_reads view of any read set (the returned object keeps the arrays alive)."""
    if isinstance(reads, ReadSet):
        return reads._ptr
    r = abi.Reads()
    names = (C.c_char_p * len(reads.contig_names))(*[n.encode() for n in reads.contig_names])
    lens = np.ascontiguousarray(reads.contig_lens, dtype="<u8")
    r.n_reads = reads.n_reads
    r.n_contigs = len(reads.contig_names)
    r.contig_names = names
    r.contig_lens = lens.ctypes.data
    for name, _ in ArrayReadSet.FIELDS:
        setattr(r, name, getattr(reads, name).ctypes.data)
    r._keep = (names, lens, reads)
    return C.pointer(r)


def write_bam(path, reads, header_text=None, threads=4):
    """Write a read set as a BAM file (lcr_host_write_bam); ReadSet keeps its QNAMEs, an ArrayReadSet gets r<index>."""
    L = host_lib()
    L.lcr_host_write_bam.argtypes = [C.c_char_p, C.POINTER(abi.Reads), C.c_char_p, C.c_int]
    rs = reads_struct(reads)
    rc = L.lcr_host_write_bam(os.fsencode(path), rs, header_text.encode() if header_text is not None else None, threads)
    if rc:
        raise LcrError(rc, f"lcr_host_write_bam({path})")


def write_phased_bam(in_bam, out_bam, regions, hp, ps, has_entry=None, threads=4):
    """The phased BAM of thread.rs:307-361 (lcr_host_write_phased_bam): returns the number of records written."""
    L = host_lib()
    L.lcr_host_write_phased_bam.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]
    regions = np.ascontiguousarray(regions, dtype=abi.REGION_DTYPE)
    hp = np.ascontiguousarray(hp, dtype="i1")
    ps = np.ascontiguousarray(ps, dtype="<u4")
    he = None if has_entry is None else np.ascontiguousarray(has_entry, dtype="u1")
    n = C.c_uint64()
    rc = L.lcr_host_write_phased_bam(os.fsencode(in_bam), os.fsencode(out_bam), regions.ctypes.data, len(regions), hp.ctypes.data, ps.ctypes.data,
                                     he.ctypes.data if he is not None else None, len(hp), threads, C.byref(n))
    if rc:
        raise LcrError(rc, f"lcr_host_write_phased_bam({in_bam})")
    return n.value


def find_regions(reads, params, truncation=False, truncation_coverage=200000):
    """Isolated regions (src/util.rs:236-332) with the read range of each; numpy REGION_DTYPE array."""
    out = C.POINTER(abi.RegionList)()
    rc = host_lib().lcr_host_find_regions(reads_struct(reads), C.byref(params), int(truncation), truncation_coverage, C.byref(out))
    if rc:
        raise LcrError(rc, "lcr_host_find_regions")
    n = out.contents.n_regions
    regions = abi.as_array(out.contents.regions, abi.REGION_DTYPE, n).copy()
    maxcov = abi.as_array(out.contents.max_coverage, "<u4", n).copy()
    host_lib().lcr_host_free_regions(out)
    return regions, maxcov


def pack_seq4(reads, alloc=None, threads=8):
    """(seq4, seq4_off): the bases of a read set in the BAM record's 4-bit form (lcr_host_pack_seq4).  alloc(nbytes) -> writable
    uint8 array lets the caller place the packed bytes in pinned memory."""
    L = host_lib()
    L.lcr_host_pack_seq4.argtypes = [C.POINTER(abi.Reads), C.c_void_p, C.c_void_p, C.c_int]
    rs = reads_struct(reads)
    off = np.zeros(reads.n_reads + 1, dtype="<u8")
    rc = L.lcr_host_pack_seq4(rs, off.ctypes.data, None, threads)
    if rc:
        raise LcrError(rc, "lcr_host_pack_seq4")
    n = int(off[-1])
    buf = (alloc(max(n, 1)) if alloc else np.empty(max(n, 1), dtype="u1"))[:n]
    rc = L.lcr_host_pack_seq4(rs, off.ctypes.data, buf.ctypes.data, threads)
    if rc:
        raise LcrError(rc, "lcr_host_pack_seq4")
    return buf, off


class BatchView:
    """An lcr_batch over numpy arrays (kept alive here).  seq4 = (packed bytes, offsets) from pack_seq4 hands the bases over in the
    BAM record's 4-bit form instead of ASCII (the `seq` pointer is then NULL); exons = per-region lists of (start, stop) turns on the
    --exon-only mask; external = per-region lists of (pos0, genotype class, qual) imports the candidates (-v) instead of calling them."""

    def __init__(self, reads, regions, seq4=None, exons=None, external=None):
        self.reads = reads
        self.regions = np.ascontiguousarray(regions, dtype=abi.REGION_DTYPE)
        names = ["pos", "flag", "mapq", "ts", "de", "seq_off", "cig_off", "qual", "cigar"] + ([] if seq4 is not None else ["seq"])
        self._keep = [np.ascontiguousarray(getattr(reads, n)) for n in names]
        b = abi.Batch()
        b.n_regions = len(self.regions)
        b.n_reads = reads.n_reads
        b.regions = self.regions.ctypes.data
        for name, a in zip(names, self._keep):
            setattr(b, name, a.ctypes.data)
        if seq4 is not None:
            s4, o4 = np.ascontiguousarray(seq4[0], dtype="u1"), np.ascontiguousarray(seq4[1], dtype="<u8")
            self._keep += [s4, o4]
            b.seq4, b.seq4_off = s4.ctypes.data, o4.ctypes.data
        if exons is not None:  # --exon-only: one list of (start, stop) per region, 1-based, stop exclusive (util.rs:435-439)
            off = np.zeros(len(self.regions) + 1, dtype="<u4")
            off[1:] = np.cumsum([len(e) for e in exons])
            iv = np.array([x for e in exons for pair in e for x in pair], dtype="<u4").reshape(-1)
            self._keep += [off, iv]
            b.exon_off, b.exon_iv = off.ctypes.data, (iv.ctypes.data if len(iv) else None)
        if external is not None:  # -v: one list of (pos0, genotype class, qual) per region, ascending positions
            off = np.zeros(len(self.regions) + 1, dtype="<u4")
            off[1:] = np.cumsum([len(e) for e in external])
            flat = [x for e in external for x in e]
            pos = np.array([x[0] for x in flat], dtype="<u4")
            gt = np.array([x[1] for x in flat], dtype="u1")
            ql = np.array([x[2] for x in flat], dtype="<f4")
            self._keep += [off, pos, gt, ql]
            b.ext_off = off.ctypes.data
            if len(flat):
                b.ext_pos, b.ext_gt, b.ext_qual = pos.ctypes.data, gt.ctypes.data, ql.ctypes.data
        self.c = b

    @property
    def nbytes(self):
        return int(sum(a.nbytes for a in self._keep) + self.regions.nbytes)


class ResultView:
    """Copies of everything in an lcr_result (the library buffer is freed by the caller afterwards)."""

    def __init__(self, res_ptr):
        r = res_ptr.contents
        self.n_regions, self.n_reads, self.n_cand = r.n_regions, r.n_reads, r.n_cand
        self.cand_off = abi.as_array(r.cand_off, "<u4", r.n_regions + 1).copy()
        self.cand = abi.as_array(r.cand, abi.CANDIDATE_DTYPE, r.n_cand).copy()
        self.region_status = abi.as_array(r.region_status, "<i4", r.n_regions).copy()
        self.hp = abi.as_array(r.hp, "i1", r.n_reads).copy()
        self.ps = abi.as_array(r.ps, "<u4", r.n_reads).copy()
        self.is_fragment = abi.as_array(r.is_fragment, "u1", r.n_reads).copy()
        self.stats = {k: getattr(r.stats, k) for k, _ in abi.Stats._fields_}
        pl = r.planes
        self.planes = None
        if pl.n_pos:
            n = pl.n_pos
            self.planes = dict(
                pos_off=abi.as_array(pl.pos_off, "<u8", r.n_regions + 1).copy(),
                acgt=abi.as_array(pl.acgt, "<u4", (n, 4)).copy(),
                fwd=abi.as_array(pl.fwd, "<u4", (n, 4)).copy(),
                d=abi.as_array(pl.d, "<u4", n).copy(),
                n=abi.as_array(pl.n, "<u4", n).copy(),
                ts=abi.as_array(pl.ts, "<u4", (n, 2)).copy(),
            )
        fr = r.fragments
        self.fragments = None
        if fr.frag_off:
            self.fragments = dict(
                frag_off=abi.as_array(fr.frag_off, "<u4", r.n_regions + 1).copy(),
                frag_read=abi.as_array(fr.frag_read, "<u4", fr.n_frag).copy(),
                elem_off=abi.as_array(fr.elem_off, "<u8", fr.n_frag + 1).copy(),
                elem_snp=abi.as_array(fr.elem_snp, "<u4", fr.n_elem).copy(),
                elem_cell=abi.as_array(fr.elem_cell, "i1", fr.n_elem).copy(),
                elem_base=abi.as_array(fr.elem_base, "u1", fr.n_elem).copy(),
            )


def format_vcf(result_ptr, batch, contig_names, min_phase_score):
    """VCF body text for a library-owned lcr_result (src/vcf.rs:27-306)."""
    names = (C.c_char_p * len(contig_names))(*[n.encode() for n in contig_names])
    out = C.c_void_p()
    n = C.c_uint64()
    rc = host_lib().lcr_host_format_vcf(C.cast(result_ptr, C.c_void_p), C.byref(batch.c), names, min_phase_score, C.byref(out), C.byref(n))
    if rc:
        raise LcrError(rc, "lcr_host_format_vcf")
    text = C.string_at(out.value, n.value).decode()
    host_lib().lcr_host_free_text(out)
    return text


def vcf_header(contig_names, contig_lens):
    names = (C.c_char_p * len(contig_names))(*[n.encode() for n in contig_names])
    lens = np.ascontiguousarray(contig_lens, dtype="<u8")
    out = C.c_void_p()
    n = C.c_uint64()
    rc = host_lib().lcr_host_format_vcf_header(names, lens.ctypes.data, len(contig_names), C.byref(out), C.byref(n))
    if rc:
        raise LcrError(rc, "lcr_host_format_vcf_header")
    text = C.string_at(out.value, n.value).decode()
    host_lib().lcr_host_free_text(out)
    return text


class Engine:
    """One lcr_ctx bound to a CUDA device: the per-region worker of src/thread.rs:78-221 on a B200."""

    def __init__(self, params, device=0):
        self.L = cuda_lib()
        self.params = params
        self.ctx = C.c_void_p()
        rc = self.L.lcr_create(C.byref(params), device, C.byref(self.ctx))
        if rc:
            raise LcrError(rc, self.L.lcr_strerror(rc).decode())
        self._refs = []

    def _check(self, rc, what):
        if rc:
            raise LcrError(rc, f"{what}: {self.L.lcr_strerror(rc).decode()} / {self.L.lcr_last_error(self.ctx).decode()}")

    def set_reference(self, tid, seq):
        a = np.ascontiguousarray(seq, dtype=np.uint8)
        self._check(self.L.lcr_set_reference(self.ctx, tid, a.ctypes.data, a.size), "lcr_set_reference")

    def set_references(self, seqs):
        for tid, s in enumerate(seqs):
            if s is not None:
                self.set_reference(tid, s)

    def submit_raw(self, batch):
        """lcr_submit: host buffers in, library-owned lcr_result out (free with free_result)."""
        out = C.POINTER(abi.Result)()
        self._check(self.L.lcr_submit(self.ctx, C.byref(batch.c), C.byref(out)), "lcr_submit")
        return out

    def free_result(self, res):
        self.L.lcr_free_result(res)

    def submit(self, batch):
        res = self.submit_raw(batch)
        try:
            return ResultView(res)
        finally:
            self.free_result(res)

    def upload(self, batch):
        h = C.c_void_p()
        self._check(self.L.lcr_upload(self.ctx, C.byref(batch.c), C.byref(h)), "lcr_upload")
        return h

    def run_device(self, handle):
        self._check(self.L.lcr_run_device(self.ctx, handle), "lcr_run_device")

    def fetch(self, handle):
        out = C.POINTER(abi.Result)()
        self._check(self.L.lcr_fetch(self.ctx, handle, C.byref(out)), "lcr_fetch")
        try:
            return ResultView(out)
        finally:
            self.free_result(out)

    def release(self, handle):
        self.L.lcr_release(self.ctx, handle)

    def discover_regions(self, reads, truncation=False, truncation_coverage=200000, with_time=False):
        """Isolated regions found on the device (lcr_discover_regions; reference util.rs:236-332): the same (regions, max_coverage)
        pair find_regions() returns from the host implementation."""
        keep = [np.ascontiguousarray(getattr(reads, f)) for f in ("tid", "pos", "flag", "mapq", "de", "seq_off", "cig_off", "cigar")]
        lens = np.ascontiguousarray(reads.contig_lens, dtype="<u8")
        a = abi.AlignIndex()
        a.n_reads, a.n_contigs, a.contig_lens = reads.n_reads, len(lens), lens.ctypes.data
        for name, arr in zip(("tid", "pos", "flag", "mapq", "de", "seq_off", "cig_off", "cigar"), keep):
            setattr(a, name, arr.ctypes.data)
        pr, pm, n, ms = C.c_void_p(), C.c_void_p(), C.c_uint32(), C.c_float()
        self._check(self.L.lcr_discover_regions(self.ctx, C.byref(a), int(truncation), truncation_coverage, C.byref(pr), C.byref(pm), C.byref(n), C.byref(ms)), "lcr_discover_regions")
        try:
            regions = abi.as_array(pr.value, abi.REGION_DTYPE, n.value).copy() if n.value else np.zeros(0, abi.REGION_DTYPE)
            maxcov = abi.as_array(pm.value, "<u4", n.value).copy() if n.value else np.zeros(0, "<u4")
        finally:
            self.L.lcr_free_regions(pr, pm)
        return (regions, maxcov, ms.value) if with_time else (regions, maxcov)

    def device_view(self, handle):
        """Device pointers of the last run's results on this handle (lcr_device_results): candidates, HP, PS."""
        v = abi.DeviceView()
        self._check(self.L.lcr_device_results(self.ctx, handle, C.byref(v)), "lcr_device_results")
        return v

    def timing(self, handle):
        t = abi.Timing()
        self._check(self.L.lcr_get_timing(self.ctx, handle, C.byref(t)), "lcr_get_timing")
        return {k: getattr(t, k) for k, _ in abi.Timing._fields_}

    def last_submit_timing(self):
        """Accounting of the last submit / submit_raw (summed over the chunks lcr_submit cut the batch into)."""
        t = abi.Timing()
        self._check(self.L.lcr_last_submit_timing(self.ctx, C.byref(t)), "lcr_last_submit_timing")
        return {k: getattr(t, k) for k, _ in abi.Timing._fields_}

    def close(self):
        if self.ctx:
            self.L.lcr_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
