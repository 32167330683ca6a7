"""C-ABI surface and host plumbing (no GPU needed): symbols, struct layouts, no-device behaviour, regions, VCF, generator."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import helpers
import oracle_binding as ob
from longcallr_b200 import abi, host

ROOT = helpers.ROOT


def _declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lcr_[a-z_0-9]+)\s*\(", text)) - {"lcr_ctx", "lcr_device_batch"})


def test_every_declared_symbol_is_exported():
    L = host.cuda_lib()
    names = _declared_functions("longcallr_b200.h")
    assert {"lcr_create", "lcr_submit", "lcr_upload", "lcr_run_device", "lcr_fetch", "lcr_release", "lcr_set_reference", "lcr_params_preset"} <= set(names)
    for n in names:
        assert hasattr(L, n), n
    assert L.lcr_abi_version() == abi.LCR_ABI_VERSION


def test_struct_layouts_match_the_header():
    """Compile a tiny C program against include/ and compare sizeof/offsetof with the ctypes mirror."""
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "longcallr_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(lcr_params), sizeof(lcr_region), sizeof(lcr_batch), sizeof(lcr_candidate),
         sizeof(lcr_planes), sizeof(lcr_fragments), sizeof(lcr_stats), sizeof(lcr_result), sizeof(lcr_timing));
  printf("%zu %zu %zu %zu\n", offsetof(lcr_params, seed), offsetof(lcr_candidate, flags), offsetof(lcr_result, stats), offsetof(lcr_params, read_assignment_cutoff));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    sizes = [int(x) for x in out]
    want = [C.sizeof(t) for t in (abi.Params, abi.Region, abi.Batch, abi.Candidate, abi.Planes, abi.Fragments, abi.Stats, abi.Result, abi.Timing)]
    assert sizes[:9] == want
    assert sizes[9:] == [abi.Params.seed.offset, abi.Candidate.flags.offset, abi.Result.stats.offset, abi.Params.read_assignment_cutoff.offset]


def test_presets_match_main_rs_defaults():
    rows = {  # SURVEY.md section 5 truth table (src/main.rs:272-396)
        "ont-cdna": (1, 10, 13.0, 0.20, 20, 1), "ont-drna": (1, 10, 13.0, 0.20, 20, 0),
        "hifi-isoseq": (0, 6, 11.0, 0.15, 40, 1), "hifi-masseq": (0, 6, 11.0, 0.15, 40, 0)}
    for name, (plat, depth, mps, maf, dte, sb) in rows.items():
        p = host.params_preset(name)
        assert (p.platform, p.min_depth, p.distance_to_read_end, p.use_strand_bias) == (plat, depth, dte, sb)
        assert abs(p.min_phase_score - mps) < 1e-6 and abs(p.min_allele_freq - maf) < 1e-6
        assert (p.min_mapq, p.min_baseq, p.min_qual, p.polya_tail_length, p.max_depth, p.min_read_length) == (20, 10, 2, 5, 50000, 500)
        assert (p.dense_win_size, p.min_dense_cnt, p.max_enum_snps, p.min_linkers, p.ld_weight_threshold) == (100, 5, 10, 1, 1)
        assert abs(p.divergence - 0.5) < 1e-6 and abs(p.low_allele_frac_cutoff - 0.05) < 1e-6 and p.low_allele_cnt_cutoff == 10


def test_no_device_means_error_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = host.params_preset("hifi-masseq")
    with pytest.raises(host.LcrError) as e:
        host.Engine(p)
    assert e.value.status == abi.LCR_ERR_NO_DEVICE


def _python_regions(reads, p):
    """find_isolated_regions_with_depth (util.rs:236-332) restated naively with per-position increments."""
    out = []
    for t, L in enumerate(reads.contig_lens):
        depth = np.zeros(int(L), dtype=np.int64)
        for i in np.nonzero(reads.tid == t)[0]:
            l_seq = int(reads.seq_off[i + 1] - reads.seq_off[i])
            if reads.mapq[i] < p.min_mapq or l_seq < p.min_read_length or reads.flag[i] & 0x904:
                continue
            if not np.isnan(reads.de[i]) and reads.de[i] >= p.divergence:
                continue
            ops = reads.cigar[int(reads.cig_off[i]):int(reads.cig_off[i + 1])]
            span = int(sum(int(o) >> 4 for o in ops if (int(o) & 0xf) in (0, 2, 3, 7, 8)))
            depth[reads.pos[i]:reads.pos[i] + span] += 1
        start = end = -1
        for i, d in enumerate(depth):
            if d == 0:
                if end > start:
                    out.append((t, start + 1, end + 2))
                start = end = -1
            elif start == -1:
                start = end = i
            else:
                end = i
        if end > start:
            out.append((t, start + 1, end + 2))
    return out


def test_find_regions_matches_naive_restatement():
    syn = host.Synthetic(seed=3, contig_len=40_000, n_contigs=2, platform=1, depth=8.0, n_het=20, n_edit=0, both_strands=1, max_intron=2000, max_gap=1500, n_threads=2)
    p = host.params_preset("ont-cdna")
    regions, maxcov = host.find_regions(syn.reads, p)
    want = _python_regions(syn.reads, p)
    assert [(int(r["tid"]), int(r["start"]), int(r["end"])) for r in regions] == want
    # the read range of a region is a superset of what fetch((chr,start,end)) returns
    for r in regions:
        for i in range(syn.reads.n_reads):
            if syn.reads.tid[i] != r["tid"]:
                continue
            ops = syn.reads.cigar[int(syn.reads.cig_off[i]):int(syn.reads.cig_off[i + 1])]
            span = int(sum(int(o) >> 4 for o in ops if (int(o) & 0xf) in (0, 2, 3, 7, 8)))
            if syn.reads.pos[i] < r["end"] and syn.reads.pos[i] + max(span, 1) > r["start"]:
                assert r["read_begin"] <= i < r["read_end"]


def test_synthetic_generator_is_deterministic_across_thread_counts():
    kw = dict(seed=9, contig_len=60_000, n_contigs=2, platform=1, depth=10.0, n_het=30, n_edit=5, both_strands=1, max_intron=500, max_gap=900)
    a, b = host.Synthetic(n_threads=1, **kw), host.Synthetic(n_threads=4, **kw)
    for f in ("tid", "pos", "flag", "ts", "de", "seq_off", "cig_off", "seq", "qual", "cigar"):
        np.testing.assert_array_equal(getattr(a.reads, f), getattr(b.reads, f))
    np.testing.assert_array_equal(a.het_pos, b.het_pos)
    assert (a.reads.qual >= 1).all()  # quality 0 would make the reference panic (phase.rs:307)
    # reads are coordinate sorted and their CIGARs consume exactly l_seq bases
    assert (np.diff(a.reads.pos[a.reads.tid == 0]) >= 0).all()
    for i in range(0, a.reads.n_reads, 37):
        ops = a.reads.cigar[int(a.reads.cig_off[i]):int(a.reads.cig_off[i + 1])]
        q = sum(int(o) >> 4 for o in ops if (int(o) & 0xf) in (0, 1, 4, 7, 8))
        assert q == int(a.reads.seq_off[i + 1] - a.reads.seq_off[i])


@pytest.mark.skipif(not os.path.exists("/root/reference/demo/demo.bam"), reason="reference demo data not present on this box")
def test_bam_decoder_on_the_demo_file():
    reads = host.ReadSet.from_bam("/root/reference/demo/demo.bam")
    assert reads.n_reads == 1713 and (reads.tid == 11).all() and reads.contig_names[11] == "chr20"
    assert (reads.flag == 0).all() and (reads.ts == ord("+")).all() and float(np.nanmax(reads.de)) < 0.03
    assert set(np.unique(reads.seq)) <= set(b"ACGTN")
    fx, _, _ = helpers.load_demo_fixture()
    for f in ("pos", "mapq", "seq_off", "cig_off", "seq", "qual", "cigar"):
        np.testing.assert_array_equal(getattr(reads, f), getattr(fx, f))


def test_vcf_text_follows_vcf_rs():
    reads, refs, regions = helpers.load_demo_fixture()
    p = host.params_preset("hifi-masseq", seed=1)
    batch = host.BatchView(reads, regions)
    raw = ob.run(p, batch, refs, mode=0, raw=True)
    try:
        text = host.format_vcf(raw, batch, reads.contig_names, p.min_phase_score)
        view = host.ResultView(raw)
    finally:
        ob.lib().lcr_oracle_free(raw)
    lines = text.strip().split("\n")
    assert len(lines) == 19
    by_pos = {int(l.split("\t")[1]): l.split("\t") for l in lines}
    f = by_pos[16730146]
    assert f[0] == "chr20" and f[2] == "." and f[3] == "G" and f[4] == "T" and f[5] == "3000" and f[6] == "PASS" and f[7] == "RDS=select"
    assert f[8] == "GT:GQ:PS:DP:AF:PQ" and f[9].startswith("0|1:2147483647:16730146:626:0.39:") or f[9].startswith("1|0:2147483647:16730146:626:0.39:")
    # every record: QUAL is the truncated variant_quality, AF has two decimals, phased records carry the phase set
    for c in view.cand:
        l = by_pos[int(c["pos"]) + 1]
        assert int(l[5]) == ob.lib().lcr_oracle_f64_as_i32(float(c["variant_quality"]))
        if "|" in l[9].split(":")[0]:
            assert l[9].split(":")[2] == str(int(c["phase_set"])) and c["phase_score"] >= p.min_phase_score
    hdr = host.vcf_header(reads.contig_names, reads.contig_lens)
    assert hdr.startswith("##fileformat=VCFv4.3\n##contig=<ID=chr20,length=") and hdr.endswith("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSample\n")
