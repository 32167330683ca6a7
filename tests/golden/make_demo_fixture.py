"""Writes tests/golden/demo_region.npz from the reference's demo data (BASELINE config 1).

Run once in the build container (needs /root/reference/demo, which does not exist on the GPU box):
    python tests/golden/make_demo_fixture.py
It stores the decoded reads of demo.bam, the isolated regions found with the hifi-masseq
defaults, and only the slice of chr20.fa those regions touch (+-1 kb).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from longcallr_b200 import host  # noqa: E402

DEMO = "/root/reference/demo"


def main():
    reads = host.ReadSet.from_bam(os.path.join(DEMO, "demo.bam"))
    ref = host.Reference.from_fasta(os.path.join(DEMO, "chr20.fa"))
    p = host.params_preset("hifi-masseq")
    regions, _ = host.find_regions(reads, p)
    tid = int(regions["tid"][0])
    assert (regions["tid"] == tid).all() and (reads.tid == tid).all()
    seq = ref.for_reads(reads)[tid]
    lo = max(0, int(regions["start"].min()) - 1001)
    hi = min(len(seq), int(regions["end"].max()) + 1000)
    regs = np.array([[0, r["start"], r["end"], r["read_begin"], r["read_end"]] for r in regions], dtype=np.int64)
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "demo_region.npz"),
        contig_len=len(seq), ref_lo=lo, ref_slice=seq[lo:hi], regions=regs,
        tid=np.zeros(reads.n_reads, dtype="<i4"), pos=reads.pos, flag=reads.flag, mapq=reads.mapq, ts=reads.ts, de=reads.de,
        seq_off=reads.seq_off, cig_off=reads.cig_off, seq=reads.seq, qual=reads.qual, cigar=reads.cigar)
    print("regions", regs, "reads", reads.n_reads, "bases", len(reads.seq))


if __name__ == "__main__":
    main()
