"""Hand-made and random alignment sets for isolated-region discovery (reference util.rs:236-332).

A case is (contig_lens, reads) with reads = list of (tid, pos, cigar ops [(op, len)], mapq, flag, de or None); `read_set` turns it
into the struct-of-arrays the host and device entry points take, `restated` into the per-contig input of the Python restatement."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))

from longcallr_b200 import host  # noqa: E402

M, I, D, N, S, H, EQ, X = 0, 1, 2, 3, 4, 5, 7, 8
REF_OPS = (M, D, N, EQ, X)
QRY_OPS = (M, I, S, EQ, X)


def rd(tid, pos, cigar, mapq=60, flag=0, de=None):
    if isinstance(cigar, int):
        cigar = [(M, cigar)]
    return (tid, pos, list(cigar), mapq, flag, de)


def read_set(contig_lens, reads):
    reads = sorted(reads, key=lambda r: (r[0] if r[0] >= 0 else 1 << 30, r[1]))
    n = len(reads)
    cig, cig_off, seq_off = [], [0], [0]
    for r in reads:
        cig += [(ln << 4) | op for op, ln in r[2]]
        cig_off.append(len(cig))
        seq_off.append(seq_off[-1] + sum(ln for op, ln in r[2] if op in QRY_OPS))
    nb = seq_off[-1]
    return host.ArrayReadSet(
        [f"c{i}" for i in range(len(contig_lens))], contig_lens,
        tid=np.array([r[0] for r in reads], "<i4"), pos=np.array([r[1] for r in reads], "<i4"), flag=np.array([r[4] for r in reads], "<u2"),
        mapq=np.array([r[3] for r in reads], "u1"), ts=np.full(n, ord("*"), "i1"), de=np.array([np.nan if r[5] is None else r[5] for r in reads], "<f4"),
        seq_off=np.array(seq_off, "<u8"), cig_off=np.array(cig_off, "<u8"), seq=np.full(nb, ord("A"), "u1"), qual=np.full(nb, 30, "u1"), cigar=np.array(cig, "<u4")), reads


def restated(contig_lens, reads, tid):
    out = []
    for r in reads:
        if r[0] != tid:
            continue
        span = sum(ln for op, ln in r[2] if op in REF_OPS)
        out.append({"mapq": r[3], "l_seq": sum(ln for op, ln in r[2] if op in QRY_OPS), "flag": r[4], "de": r[5], "pos": r[1], "end": r[1] + span})
    return out


def cases():
    c = {}
    c["two_runs"] = ([100], [rd(0, 10, 20), rd(0, 15, 30), rd(0, 70, 5)])
    c["single_position_run_joins_next"] = ([200], [rd(0, 5, 1), rd(0, 100, 50)])  # region 6..151: the one-base run is never reset (util.rs:297-311)
    c["single_position_chain"] = ([300], [rd(0, 5, 1), rd(0, 9, 1), rd(0, 13, 1), rd(0, 20, 1), rd(0, 30, 1), rd(0, 100, 10), rd(0, 150, 1)])
    c["single_position_last_of_contig"] = ([100, 100], [rd(0, 10, 10), rd(0, 50, 1), rd(1, 0, 100)])
    c["single_then_pair_then_single"] = ([100], [rd(0, 3, 1), rd(0, 5, 1), rd(0, 7, 1), rd(0, 9, 2), rd(0, 20, 1), rd(0, 30, 4)])
    c["run_reaches_contig_end"] = ([64, 10], [rd(0, 40, 24), rd(1, 0, 10)])
    c["alignment_past_contig_end"] = ([50], [rd(0, 30, 40)])
    c["introns_count_as_depth"] = ([1000], [rd(0, 10, [(S, 7), (M, 20), (N, 500), (M, 20), (H, 3)]), rd(0, 100, [(M, 10), (D, 5), (I, 3), (EQ, 4), (X, 2)])])
    c["filtered_reads"] = ([400], [rd(0, 10, 30, mapq=5), rd(0, 60, 30, flag=0x100), rd(0, 110, 30, flag=0x800), rd(0, 160, 30, flag=0x4), rd(0, 210, 30, de=0.5), rd(0, 260, 30, de=0.001),
                                   rd(0, 310, [(M, 30)])])
    c["no_passing_read"] = ([100], [rd(0, 10, 30, mapq=0)])
    c["empty_middle_contig"] = ([100, 100, 100], [rd(0, 10, 30), rd(2, 20, 30)])
    c["unmapped_tail"] = ([100], [rd(0, 10, 30), rd(-1, -1, [], flag=0x4), rd(-1, -1, [], flag=0x4)])
    deep = [rd(0, 100 + i, 200) for i in range(40)] + [rd(0, 150, 20) for _ in range(30)]
    c["truncation_splits_a_run"] = ([600], deep)  # with truncation_coverage 50: the pile of 70 at 150..170 is cut out, max_coverage still sees the first cut position
    c["truncation_leaves_single_positions"] = ([300], [rd(0, 10, 50) for _ in range(3)] + [rd(0, 11, 20) for _ in range(9)] + [rd(0, 32, 1) for _ in range(9)] + [rd(0, 34, 5) for _ in range(9)])
    return c


def random_case(rng, n_contigs=3, contig_len=4000, n_reads=300):
    lens = [int(rng.integers(contig_len // 2, contig_len)) for _ in range(n_contigs)]
    reads = []
    for _ in range(n_reads):
        t = int(rng.integers(0, n_contigs))
        kind = rng.integers(0, 10)
        pos = int(rng.integers(0, lens[t]))
        if kind < 3:
            cig = [(M, int(rng.integers(1, 3)))]  # many one- and two-base runs
        elif kind < 6:
            cig = [(M, int(rng.integers(1, 120)))]
        else:
            cig = [(S, int(rng.integers(1, 9))), (M, int(rng.integers(1, 60))), (N, int(rng.integers(1, 400))), (M, int(rng.integers(1, 60))), (D, 2), (M, 3)]
        reads.append(rd(t, pos, cig, mapq=int(rng.choice([0, 20, 60])), flag=int(rng.choice([0, 0, 0, 16, 0x100, 0x800])), de=None if rng.random() < 0.5 else float(rng.random() * 0.06)))
    return lens, reads
