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


def test_annotation_and_candidate_vcf_files(tmp_path):
    """--exon-only from a GTF file and -v from a VCF file (longcallr_b200/alt_inputs.py) through the CUDA path, against the oracle."""
    from longcallr_b200 import abi, alt_inputs

    syn = host.Synthetic(seed=33, contig_len=150_000, n_contigs=2, platform=0, depth=30.0, n_het=150, n_edit=20, both_strands=0, n_threads=4)
    p = host.params_preset("hifi-masseq", seed=5)
    eng = host.Engine(p, device=0)
    regions, maxcov = eng.discover_regions(syn.reads)
    refs = syn.reference.for_reads(syn.reads)
    eng.set_references(refs)
    names = syn.reads.contig_names
    plain = eng.submit(host.BatchView(syn.reads, regions))
    # a gene over two thirds of the regions (two of them overlapping pairs), CDS records over parts of each gene
    lines = []
    for r, g in enumerate(regions):
        if r % 3 == 2:
            continue
        s, e = int(g["start"]), int(g["end"]) - 1
        lines.append((names[g["tid"]], s, f'{names[g["tid"]]}\tx\tgene\t{s}\t{e}\t.\t+\t.\tgene_id "g{r}";'))
        if r % 6 != 4:  # a gene without CDS: its region is skipped
            lines.append((names[g["tid"]], s, f'{names[g["tid"]]}\tx\tCDS\t{s + (e - s) // 4}\t{s + (e - s) // 2}\t.\t+\t0\tgene_id "g{r}";'))
            lines.append((names[g["tid"]], s, f'{names[g["tid"]]}\tx\tCDS\t{s + (e - s) // 3}\t{e}\t.\t+\t0\tgene_id "g{r}";'))
    (tmp_path / "a.gtf").write_text("\n".join(x[2] for x in lines) + "\n")
    genes, exons = alt_inputs.parse_annotation(str(tmp_path / "a.gtf"))
    cut = alt_inputs.intersect_gene_regions([(names[g["tid"]], int(g["start"]), int(g["end"]), int(m)) for g, m in zip(regions, maxcov)], genes)
    by_start = {(names[g["tid"]], int(g["start"])): g for g in regions}
    sub = np.zeros(len(cut), dtype=abi.REGION_DTYPE)
    for i, (c, s, e, _, gid) in enumerate(cut):
        src = by_start[(c, s)]
        sub[i] = (src["tid"], s, e, src["read_begin"], src["read_end"])
    batch = host.BatchView(syn.reads, sub, exons=alt_inputs.exons_for_regions([c[4] for c in cut], exons))
    got = eng.submit(batch)
    helpers.compare_results(got, ob.run(p, batch, refs, mode=0), "exon-only from a GTF")
    assert (got.region_status == abi.LCR_REGION_NO_EXON).sum() >= 1 and 0 < got.n_cand < plain.n_cand
    # the calls of the plain run written as a VCF (het, hom and a few 0/0 and missing-QUAL records) and imported again
    with open(tmp_path / "c.vcf", "w") as f:
        f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS\n")
        for i, c in enumerate(plain.cand):
            gt = {1: "0/1", 2: "1/1", 3: "1/2"}.get(int(c["variant_type"]), "0/0")
            ql = "." if i % 11 == 0 else str(int(c["variant_quality"]) % 3000)
            f.write(f"{names[regions[c['region']]['tid']]}\t{int(c['pos']) + 1}\t.\t{chr(c['reference'])}\tN\t{ql}\tPASS\t.\tGT:GQ\t{gt}:9\n")
    rec = alt_inputs.read_candidate_vcf(str(tmp_path / "c.vcf"))
    ext = alt_inputs.external_for_regions([(names[g["tid"]], int(g["start"]), int(g["end"])) for g in regions], rec)
    batch = host.BatchView(syn.reads, regions, external=ext)
    got = eng.submit(batch)
    eng.close()
    helpers.compare_results(got, ob.run(p, batch, refs, mode=0), "-v from a VCF")
    assert got.n_cand == int((plain.cand["variant_type"] != 0).sum()) and int((got.hp > 0).sum()) > 200
