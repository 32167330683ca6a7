"""Readers of the alternate entries' files (longcallr_b200/alt_inputs.py): annotation (util.rs:334-556) and candidate VCF (vcf.rs:400-462)."""
import gzip
import math

import numpy as np
import pytest

import oracle_binding as ob
from longcallr_b200 import abi, alt_inputs, host

GTF = """##description: hand-made
chr1\tsrc\tgene\t100\t500\t.\t+\t.\tgene_id "G1"; gene_name "a";
chr1\tsrc\ttranscript\t100\t500\t.\t+\t.\tgene_id "G1"; transcript_id "T1";
chr1\tsrc\tCDS\t120\t200\t.\t+\t0\tgene_id "G1"; transcript_id "T1";
chr1\tsrc\tCDS\t300\t400\t.\t+\t0\tgene_id "G1"; transcript_id "T1";
chr1\tsrc\tgene\t450\t900\t.\t-\t.\tgene_id "G2";
chr1\tsrc\tCDS\t600\t700\t.\t-\t0\tgene_id "G2";
chr1\tsrc\tgene\t480\t520\t.\t+\t.\tgene_id "G3";
chr1\tsrc\tgene\t901\t950\t.\t+\t.\tgene_id "G4";
chr1\tsrc\tCDS\t910\t920\t.\t+\t0\tgene_id "G4";
chr2\tsrc\tgene\t10\t50\t.\t+\t.\tID=gene5;gene_id=G5;biotype=x
chr2\tsrc\tCDS\t20\t30\t.\t+\t0\tParent=t5;gene_id=G5
"""


def test_parse_annotation_and_intersection(tmp_path):
    path = tmp_path / "a.gtf"
    path.write_text(GTF)
    genes, exons = alt_inputs.parse_annotation(str(path))
    # G1 [100,501) overlaps G2 (450..900): merged and extended; G3 is contained; G4 starts exactly where the merged region ends: separate
    assert genes["chr1"] == [[100, 901, "G1,G2,G3"], [901, 951, "G4"]] and genes["chr2"] == [[10, 51, "G5"]]
    assert exons == {"G1": [(120, 201), (300, 401)], "G2": [(600, 701)], "G4": [(910, 921)], "G5": [(20, 31)]}
    regions = [("chr1", 50, 300, 7), ("chr1", 890, 940, 9), ("chr3", 1, 100, 1), ("chr2", 51, 80, 2)]
    cut = alt_inputs.intersect_gene_regions(regions, genes)
    assert cut == [("chr1", 100, 300, 7, "G1,G2,G3"), ("chr1", 890, 901, 9, "G1,G2,G3"), ("chr1", 901, 940, 9, "G4")]
    ex = alt_inputs.exons_for_regions([c[4] for c in cut] + ["G3"], exons)
    assert ex == [[(120, 201), (300, 401), (600, 701)], [(120, 201), (300, 401), (600, 701)], [(910, 921)], []]
    gz = tmp_path / "a.gtf.gz"
    with gzip.open(gz, "wt") as f:
        f.write(GTF)
    assert alt_inputs.parse_annotation(str(gz)) == (genes, exons)
    (tmp_path / "bad.gtf").write_text(GTF.replace("\t480\t520\t", "\t80\t520\t"))
    with pytest.raises(ValueError):
        alt_inputs.parse_annotation(str(tmp_path / "bad.gtf"))


VCF = """##fileformat=VCFv4.2
#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2
chr1\t101\t.\tA\tG\t30\tPASS\t.\tGT:GQ\t0/1:9\t1|1:8
chr1\t150\t.\tA\tG\t.\tPASS\t.\tGT\t0|1\t./.
chr1\t160\t.\tA\tG,T\t12.5\tPASS\t.\tGT\t1/2\t1\t
chr1\t170\t.\tA\tG\t-1\tPASS\t.\tGQ:GT\t5:0/0\t5:1/0
chr2\t7\t.\tC\tT\t3000\tPASS\t.\tGT\t0/2\t2|0
"""


def test_read_candidate_vcf(tmp_path):
    path = tmp_path / "c.vcf"
    path.write_text(VCF)
    rec = alt_inputs.read_candidate_vcf(str(path))
    assert rec["chr1"][100] == (2, 30.0)  # the last sample with two alleles wins
    assert rec["chr1"][149][0] == 4 and math.isnan(rec["chr1"][149][1])  # ./. -> (3, 3) -> other; missing QUAL
    assert rec["chr1"][159] == (3, 12.5)  # the haploid second sample is skipped
    assert rec["chr1"][169] == (1, -1.0) and rec["chr2"][6] == (4, 3000.0)
    ext = alt_inputs.external_for_regions([("chr1", 101, 161), ("chr1", 161, 400), ("chr9", 1, 10)], rec)
    assert [[e[0] for e in r] for r in ext] == [[100, 149, 159], [169], []]


def test_files_drive_the_alternate_entries(tmp_path):
    """GTF + VCF files through the readers into lcr_batch, on the oracle (the GPU twin is tests/test_gpu_file_pipeline.py)."""
    syn = host.Synthetic(seed=9, contig_len=60_000, n_contigs=1, platform=0, depth=25.0, n_het=60, n_edit=10, both_strands=0, n_threads=2)
    p = host.params_preset("hifi-masseq", seed=2)
    regions, maxcov = host.find_regions(syn.reads, p)
    refs = syn.reference.for_reads(syn.reads)
    name = syn.reads.contig_names[0]
    plain = ob.run(p, host.BatchView(syn.reads, regions), refs, mode=0)
    # annotation: one gene over every other alignment region, CDS over its first half
    lines, want_inside = [], []
    for r, g in enumerate(regions):
        if r % 2:
            continue
        s, e = int(g["start"]), int(g["end"]) - 1
        lines.append(f'{name}\tx\tgene\t{s}\t{e}\t.\t+\t.\tgene_id "g{r}";')
        lines.append(f'{name}\tx\tCDS\t{s}\t{(s + e) // 2}\t.\t+\t0\tgene_id "g{r}";')
        want_inside.append((s, (s + e) // 2))
    (tmp_path / "a.gtf").write_text("\n".join(lines) + "\n")
    genes, exons = alt_inputs.parse_annotation(str(tmp_path / "a.gtf"))
    cut = alt_inputs.intersect_gene_regions([(name, int(g["start"]), int(g["end"]), int(m)) for g, m in zip(regions, maxcov)], genes)
    assert len(cut) == (len(regions) + 1) // 2
    sub = np.zeros(len(cut), dtype=abi.REGION_DTYPE)
    for i, (c, s, e, _, gid) in enumerate(cut):
        src = regions[2 * i]
        sub[i] = (0, s, e, src["read_begin"], src["read_end"])
    got = ob.run(p, host.BatchView(syn.reads, sub, exons=alt_inputs.exons_for_regions([c[4] for c in cut], exons)), refs, mode=0)
    assert 0 < got.n_cand < plain.n_cand
    for q in got.cand["pos"]:
        assert any(s <= int(q) + 1 <= e for s, e in want_inside)
    # candidate VCF: the het calls of the plain run, written as a VCF and imported
    het = plain.cand[plain.cand["variant_type"] == 1]
    with open(tmp_path / "c.vcf", "w") as f:
        f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS\n")
        for c in het:
            f.write(f"{name}\t{int(c['pos']) + 1}\t.\t{chr(c['reference'])}\tN\t{int(c['variant_quality'])}\tPASS\t.\tGT\t0/1\n")
    rec = alt_inputs.read_candidate_vcf(str(tmp_path / "c.vcf"))
    ext = alt_inputs.external_for_regions([(name, int(g["start"]), int(g["end"])) for g in regions], rec)
    imp = ob.run(p, host.BatchView(syn.reads, regions, external=ext), refs, mode=0)
    assert imp.n_cand == len(het) and list(imp.cand["pos"]) == list(het["pos"])  # (phasing may still re-genotype a site afterwards)
    assert int((imp.hp > 0).sum()) > 100
