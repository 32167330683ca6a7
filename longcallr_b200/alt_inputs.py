"""Host-side readers of the two alternate entries: the annotation of `--exon-only` and the candidate VCF of
`-v`.  They turn the files into the per-region arrays lcr_batch carries (exon_off / exon_iv, ext_off / ext_pos / ext_gt / ext_qual).

  parse_annotation        src/util.rs:334-452   gene regions per contig (overlapping genes merged, ids joined by ","), CDS intervals per gene
  intersect_gene_regions  src/util.rs:454-556   alignment regions cut to the gene regions they overlap (merge = true, main.rs:223)
  exons_for_regions       src/thread.rs:80-91   the exon intervals of a region's genes (none: the region is skipped)
  read_candidate_vcf      src/vcf.rs:400-462    position -> (genotype class, QUAL) per contig; the last sample of a record wins
  external_for_regions    src/candidate.rs:544  the records whose position falls inside a region

Plain-text and gzip / BGZF-compressed files are read; BCF is not.
"""
import gzip
import math


def _open_text(path):
    with open(path, "rb") as f:
        magic = f.read(2)
    return gzip.open(path, "rt") if magic == b"\x1f\x8b" else open(path, "rt")


def _gene_id(attr):
    for sub in attr.rstrip().split(";"):
        t = sub.strip()
        if t.startswith("gene_id="):  # GFF3
            return t[len("gene_id="):]
        if t.startswith("gene_id "):  # GTF
            return t[len("gene_id "):].strip('"')
    return ""


def parse_annotation(path):
    """(gene_regions, exon_regions): {contig: [[start, end, "id[,id...]"], ...]} with start 1-based inclusive and end exclusive, and
    {gene id: [(start, stop), ...]} from the CDS records (stop = end + 1), exactly as the reference keeps them."""
    gene_regions, exon_regions = {}, {}
    invs, gene_id = [], ""
    with _open_text(path) as f:
        for line in f:
            if line.startswith("#"):
                continue
            parts = line.rstrip("\n").split("\t")
            if len(parts) < 9:
                continue
            seqname, feature, start, end = parts[0], parts[2], int(parts[3]), int(parts[4])
            if feature == "gene":
                if invs:
                    exon_regions[gene_id] = list(invs)  # keyed by the gene that just ended (util.rs:357-360)
                    invs = []
                regions = gene_regions.setdefault(seqname, [])
                gene_id = _gene_id(parts[8])
                if regions:
                    top = regions.pop()
                    if start < top[0]:
                        raise ValueError(f"annotation file is not sorted. {seqname}:{start}-{end}")
                    if top[1] <= start:  # top's end is exclusive: no overlap
                        regions.append(top)
                        regions.append([start, end + 1, gene_id])
                    elif top[1] < end + 1:  # overlap: extend and join the ids
                        regions.append([top[0], end + 1, top[2] + "," + gene_id])
                    else:  # contained
                        regions.append([top[0], top[1], top[2] + "," + gene_id])
                else:
                    regions.append([start, end + 1, gene_id])
            elif feature == "CDS":
                if _gene_id(parts[8]) != gene_id:
                    raise ValueError(f"gene_id in gene and exon are different: gene_id:{gene_id}, exon_gene_id:{_gene_id(parts[8])}")
                invs.append((start, end + 1))
    if invs:
        exon_regions[gene_id] = list(invs)
    return gene_regions, exon_regions


def intersect_gene_regions(regions, gene_regions):
    """regions: [(contig name, start, end, max_coverage)] as region discovery returns them (start 1-based inclusive, end exclusive).
    Returns [(contig, start, end, max_coverage, gene ids)]: every alignment region cut to each gene region it overlaps, in query order and,
    per query, in the order of the gene regions' starts (Lapper::find); contigs without annotation drop out (util.rs:538-549)."""
    out = []
    for chrom, qs, qe, cov in regions:
        if chrom not in gene_regions:
            continue
        for hs, he, gid in sorted(gene_regions[chrom], key=lambda r: r[0]):
            if hs < qe and he > qs:
                out.append((chrom, max(qs, hs), min(qe, he), cov, gid))
    return out


def exons_for_regions(gene_ids, exon_regions):
    """Per region the concatenated exon intervals of its genes (thread.rs:80-87); an empty list means the region is skipped."""
    return [[iv for g in gid.split(",") if g in exon_regions for iv in exon_regions[g]] for gid in gene_ids]


def read_candidate_vcf(path):
    """{contig: {0-based position: (genotype class, qual)}} (vcf.rs:400-462): classes 0 = 0/0, 1 = 0/1 or 1/0, 2 = 1/1, 3 = 1/2 or 2/1,
    4 = anything else ('.' counts as allele 3); records whose GT does not have two alleles are skipped; QUAL '.' is NaN."""
    out = {}
    with _open_text(path) as f:
        for line in f:
            if line.startswith("#"):
                continue
            p = line.rstrip("\n").split("\t")
            if len(p) < 10:
                continue
            chrom, pos0 = p[0], int(p[1]) - 1
            qual = float("nan") if p[5] == "." else float(p[5])
            fmt = p[8].split(":")
            if "GT" not in fmt:
                continue
            gi = fmt.index("GT")
            for sample in p[9:]:
                fields = sample.split(":")
                gt = fields[gi] if gi < len(fields) else "."
                alleles = gt.replace("|", "/").split("/")
                if len(alleles) != 2:
                    continue
                a = [3 if x == "." else int(x) for x in alleles]
                cls = {(0, 0): 0, (0, 1): 1, (1, 0): 1, (1, 1): 2, (1, 2): 3, (2, 1): 3}.get((a[0], a[1]), 4)
                out.setdefault(chrom, {})[pos0] = (cls, qual)
    return out


def external_for_regions(regions, records):
    """regions: [(contig name, start, end)]; per region the ascending list of (pos0, class, qual) with start - 1 <= pos0 < end - 1."""
    out = []
    for chrom, start, end in regions:
        recs = records.get(chrom, {})
        out.append(sorted((pos, cls, q if not (isinstance(q, float) and math.isnan(q)) else float("nan")) for pos, (cls, q) in recs.items() if start - 1 <= pos < end - 1))
    return out
