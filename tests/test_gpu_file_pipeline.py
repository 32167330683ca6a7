"""The callers either side of the hot path, end to end on the GPU box: a generated BAM + FASTA are decoded, the regions
are found on the device, the batch goes through the C ABI, and the VCF text and phased BAM are written; every product is compared with what the
oracle's result gives through the same writers, and the phased BAM with the restatement of thread.rs:307-361."""
import filecmp

import numpy as np
import pytest

import bam_py
import helpers
import oracle_binding as ob
import region_cases as rc  # noqa: F401
import py_restatement as pr  # noqa: I001
from longcallr_b200 import host

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("preset,platform", [("hifi-masseq", 0), ("ont-cdna", 1)])
def test_bam_in_vcf_and_phased_bam_out(tmp_path, preset, platform):
    syn = host.Synthetic(seed=21 + platform, contig_len=150_000, n_contigs=2, platform=platform, depth=30.0, n_het=150, n_edit=20, both_strands=platform, n_threads=4)
    bam, fa = str(tmp_path / "in.bam"), str(tmp_path / "ref.fa")
    host.write_bam(bam, syn.reads)
    with open(fa, "w") as f:
        for name, seq in zip(syn.reference.names, syn.reference.seqs):
            s = bytes(seq).decode()
            f.write(f">{name} synthetic\n" + "\n".join(s[i : i + 60] for i in range(0, len(s), 60)) + "\n")
    reads = host.ReadSet.from_bam(bam, threads=4)
    ref = host.Reference.from_fasta(fa)
    refs = ref.for_reads(reads)
    p = host.params_preset(preset, seed=5)
    eng = host.Engine(p, device=0)
    regions, maxcov = eng.discover_regions(reads)
    want_regions, want_maxcov = host.find_regions(reads, p)
    assert len(regions) >= 8 and regions.tobytes() == want_regions.tobytes() and (maxcov == want_maxcov).all()
    batch = host.BatchView(reads, regions)
    eng.set_references(refs)
    raw = eng.submit_raw(batch)
    try:
        got = host.ResultView(raw)
        got_vcf = host.format_vcf(raw, batch, reads.contig_names, p.min_phase_score)
    finally:
        eng.free_result(raw)
    eng.close()
    oraw = ob.run(p, batch, refs, mode=0, raw=True)
    try:
        want = host.ResultView(oraw)
        want_vcf = host.format_vcf(oraw, batch, reads.contig_names, p.min_phase_score)
    finally:
        ob.lib().lcr_oracle_free(oraw)
    helpers.compare_results(got, want, preset)
    assert got_vcf == want_vcf and got_vcf.count("\n") > 50
    out_got, out_want = str(tmp_path / "got.bam"), str(tmp_path / "want.bam")
    n = host.write_phased_bam(bam, out_got, regions, got.hp, got.ps, got.is_fragment)
    assert n == host.write_phased_bam(bam, out_want, regions, want.hp, want.ps, want.is_fragment)
    assert filecmp.cmp(out_got, out_want, shallow=False)
    src, out = bam_py.read_bam(bam), bam_py.read_bam(out_got)
    qn = [r["qname"] for r in src["records"]]
    hq = [(qn[i], int(want.hp[i])) for i in range(len(qn)) if want.is_fragment[i]]
    pq = [(qn[i], int(want.ps[i])) for i in range(len(qn)) if want.ps[i]]
    emit = pr.phased_bam(src["records"], [(int(r["tid"]), int(r["start"]), int(r["end"])) for r in regions], hq, pq)
    assert len(emit) == n == len(out["records"])
    tagged = 0
    for (i, w_hp, w_ps), r in zip(emit, out["records"]):
        assert r["core_and_data"] == src["records"][i]["core_and_data"]
        assert r["tags"].get("HP") == (None if w_hp is None else ("i", w_hp)) and r["tags"].get("PS") == (None if w_ps is None else ("I", w_ps))
        tagged += w_hp is not None
    assert tagged > 200
