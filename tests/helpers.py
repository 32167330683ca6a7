"""Shared pieces of the parity tests: fixtures, batch construction, result comparison."""
import os

import numpy as np

from longcallr_b200 import abi, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

INT_FIELDS = ("pos", "depth", "phase_set", "reference", "alleles", "variant_type", "genotype", "haplotype", "flags", "region")
FP_FIELDS = ("variant_quality", "genotype_quality", "phase_score", "genotype_probability", "allele_freqs")
FP_TOL = 1e-5  # north_star: "within 1e-5 on FP likelihood/QUAL fields"


def ArrayReads(contig_names, contig_lens, tid, pos, flag, mapq, ts, de, seq_off, cig_off, seq, qual, cigar):
    return host.ArrayReadSet(contig_names, contig_lens, tid=tid, pos=pos, flag=flag, mapq=mapq, ts=ts, de=de, seq_off=seq_off,
                             cig_off=cig_off, seq=seq, qual=qual, cigar=cigar)


def cigar_ops(s):
    """'10S90M' -> uint32 BAM ops."""
    ops = "MIDNSHP=XB"
    out, num = [], ""
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            out.append((int(num) << 4) | ops.index(ch))
            num = ""
    return out


def make_reads(contig_len, recs, contig="c1"):
    """recs: dicts with pos, cigar (str), seq (str), qual (list/int), flag, mapq, ts, de."""
    recs = sorted(recs, key=lambda r: r["pos"])
    seq_off, cig_off, seq, qual, cig = [0], [0], [], [], []
    for r in recs:
        s = r["seq"].encode()
        q = r.get("qual", 30)
        q = [q] * len(s) if np.isscalar(q) else list(q)
        assert len(q) == len(s)
        seq += list(s)
        qual += q
        cig += cigar_ops(r["cigar"])
        seq_off.append(len(seq))
        cig_off.append(len(cig))
    n = len(recs)
    return ArrayReads(
        [contig], [contig_len], [0] * n, [r["pos"] for r in recs], [r.get("flag", 0) for r in recs], [r.get("mapq", 60) for r in recs],
        [ord(r.get("ts", "+")) for r in recs], [r.get("de", 0.01) for r in recs], seq_off, cig_off, seq, qual, cig)


def one_region(start, end, n_reads, tid=0):
    reg = np.zeros(1, dtype=abi.REGION_DTYPE)
    reg[0] = (tid, start, end, 0, n_reads)
    return reg


def load_demo_fixture():
    """cfg1: the decoded demo region (tests/golden/make_demo_fixture.py wrote it from demo/demo.bam + chr20.fa)."""
    z = np.load(os.path.join(GOLDEN, "demo_region.npz"))
    reads = ArrayReads(["chr20"], [int(z["contig_len"])], z["tid"], z["pos"], z["flag"], z["mapq"], z["ts"], z["de"], z["seq_off"], z["cig_off"], z["seq"], z["qual"], z["cigar"])
    ref = np.full(int(z["contig_len"]), ord("N"), dtype=np.uint8)
    lo = int(z["ref_lo"])
    ref[lo : lo + len(z["ref_slice"])] = z["ref_slice"]
    regions = np.zeros(len(z["regions"]), dtype=abi.REGION_DTYPE)
    for i, r in enumerate(z["regions"]):
        regions[i] = tuple(int(v) for v in r)
    return reads, [ref], regions


def compare_results(a, b, what=""):
    """a, b: host.ResultView.  Integers / bytes / indices bit-exact, FP within FP_TOL (and counted when not bit-equal)."""
    assert a.n_regions == b.n_regions and a.n_reads == b.n_reads, what
    np.testing.assert_array_equal(a.region_status, b.region_status, err_msg=what + " region_status")
    np.testing.assert_array_equal(a.cand_off, b.cand_off, err_msg=what + " cand_off")
    assert a.n_cand == b.n_cand, what
    for f in INT_FIELDS:
        np.testing.assert_array_equal(a.cand[f], b.cand[f], err_msg=f"{what} cand.{f}")
    inexact = 0
    for f in FP_FIELDS:
        x, y = a.cand[f].astype(np.float64), b.cand[f].astype(np.float64)
        same_inf = np.isinf(x) & np.isinf(y) & (np.sign(x) == np.sign(y))
        ok = same_inf | (np.abs(x - y) <= FP_TOL) | (np.isnan(x) & np.isnan(y))
        assert ok.all(), f"{what} cand.{f}: max diff {np.nanmax(np.abs(x - y)[~same_inf])}"
        inexact += int((np.ascontiguousarray(a.cand[f]).view(np.uint8) != np.ascontiguousarray(b.cand[f]).view(np.uint8)).any())
    np.testing.assert_array_equal(a.hp, b.hp, err_msg=what + " hp")
    np.testing.assert_array_equal(a.ps, b.ps, err_msg=what + " ps")
    np.testing.assert_array_equal(a.is_fragment, b.is_fragment, err_msg=what + " is_fragment")
    if (a.region_status == 0).all():  # a failed region stops the reference mid-way; its partial counts are not defined
        # the metric's numerator is n_aligned_bases + nnz_phase; the sweep counters pin the iteration-for-iteration equality of the phasing
        for k in ("n_reads_pass", "n_aligned_bases", "n_positions", "n_candidates", "n_fragments", "nnz_phase", "n_cross_optimize", "n_sweep_iters"):
            assert a.stats[k] == b.stats[k], f"{what} stats.{k}: {a.stats[k]} != {b.stats[k]}"
    if a.planes is not None and b.planes is not None:
        for k in ("pos_off", "acgt", "fwd", "d", "n", "ts"):
            np.testing.assert_array_equal(a.planes[k], b.planes[k], err_msg=f"{what} planes.{k}")
    if a.fragments is not None and b.fragments is not None:
        for k in ("frag_off", "frag_read", "elem_off", "elem_snp", "elem_cell", "elem_base"):
            np.testing.assert_array_equal(a.fragments[k], b.fragments[k], err_msg=f"{what} fragments.{k}")
    return inexact
