"""BAM output: the BAM writer round-trips through both decoders, and the phased BAM follows
thread.rs:307-361 (restated in oracle/py_restatement.py: phased_bam)."""
import struct

import numpy as np

import bam_py
import helpers  # noqa: F401
import oracle_binding as ob
import region_cases as rc  # noqa: F401  (puts oracle/ on sys.path)
import py_restatement as pr  # noqa: I001
from longcallr_b200 import host


def test_written_bam_round_trips(tmp_path):
    syn = host.Synthetic(seed=2, contig_len=50_000, n_contigs=2, platform=1, depth=12.0, n_het=40, n_edit=8, both_strands=1, n_threads=2)
    path = str(tmp_path / "syn.bam")
    host.write_bam(path, syn.reads)
    back = host.ReadSet.from_bam(path, threads=2)
    assert back.n_reads == syn.reads.n_reads and back.contig_names == syn.reads.contig_names
    for f in ("contig_lens", "tid", "pos", "flag", "mapq", "ts", "seq_off", "cig_off", "seq", "qual", "cigar", "qname_off", "qnames"):
        np.testing.assert_array_equal(getattr(back, f), getattr(syn.reads, f), err_msg=f)
    np.testing.assert_array_equal(np.isnan(back.de), np.isnan(syn.reads.de))
    np.testing.assert_array_equal(back.de[~np.isnan(back.de)], syn.reads.de[~np.isnan(syn.reads.de)])
    # the independent Python decoder sees the same records
    py = bam_py.read_bam(path)
    assert [n for n, _ in py["refs"]] == syn.reads.contig_names and len(py["records"]) == back.n_reads
    for i in (0, 1, back.n_reads // 2, back.n_reads - 1):
        r = py["records"][i]
        assert (r["tid"], r["pos"], r["flag"], r["mapq"], r["qname"]) == (back.tid[i], back.pos[i], back.flag[i], back.mapq[i], back.qname(i))
        assert r["seq"].encode() == bytes(back.seq[int(back.seq_off[i]) : int(back.seq_off[i + 1])])
        assert r["cigar"] == list(back.cigar[int(back.cig_off[i]) : int(back.cig_off[i + 1])])


def emitted(path):
    out = bam_py.read_bam(path)
    return out, [(r["qname"], r["pos"], r["tags"].get("HP"), r["tags"].get("PS")) for r in out["records"]]


def test_phased_bam_hand_made(tmp_path):
    """Fully-inside rule, fetch-window edge, unmapped / secondary / supplementary, existing tags kept, QNAME-level first entry wins."""
    M, S, N = 0, 4, 3
    hp_old = b"HPi" + struct.pack("<i", 9)
    ps_old = b"PSI" + struct.pack("<I", 77)
    recs = [
        ("a", 99, [(M, 50)], 0, b""),            # 0 fully inside region 1 (covered 99..198 -> start 100, end 200)
        ("b", 99, [(M, 1)], 0, b""),             # 1 one base at start-1: endpos 100 is not > start 100, fetch does not return it
        ("c", 120, [(S, 5), (M, 30), (N, 40), (M, 9)], 16, b""),  # 2 inside, spliced
        ("d", 130, [(M, 20)], 0x100, b""),       # 3 secondary
        ("e", 131, [(M, 20)], 0x800, b""),       # 4 supplementary
        ("f", 140, [(M, 20)], 0, hp_old),        # 5 carries HP already: push_aux refuses, PS still pushed
        ("g", 141, [(M, 20)], 0, ps_old),        # 6 carries PS already
        ("h", 150, [(M, 60)], 0, b""),           # 7 ends at 210 > region end - 1: dropped
        ("a", 160, [(M, 20)], 0, b""),           # 8 same QNAME as record 0: gets record 0's values
        ("i", 170, [(M, 29)], 0, b""),           # 9 ends exactly at 199 == end - 1: kept
        ("j", 400, [(M, 50)], 0, b""),           # 10 region 2, no entry at all
        ("k", 410, [(M, 30)], 0, b""),           # 11 region 2, assignment 0 but a phase set
        ("u", -1, [], 4, b""),                   # 12 unmapped
    ]
    records = [bam_py.make_record(0 if pos >= 0 else -1, pos, q, cig, flag=flag, aux=aux) for q, pos, cig, flag, aux in recs]
    path, out_path = str(tmp_path / "in.bam"), str(tmp_path / "out.bam")
    bam_py.write_bam(path, [("chrT", 1000)], records)
    rs = host.ReadSet.from_bam(path)
    assert rs.n_reads == len(recs)
    regions = np.zeros(2, dtype=host.abi.REGION_DTYPE)
    regions[0] = (0, 100, 200, 0, 10)
    regions[1] = (0, 401, 452, 10, 12)
    hp = np.array([1, 2, 2, 1, 1, 1, 2, 1, 2, 0, 0, 0, 0], "i1")
    ps = np.array([100, 100, 100, 100, 100, 100, 100, 100, 555, 0, 0, 405, 0], "<u4")
    has = np.array([1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 1, 0], "u1")
    n = host.write_phased_bam(path, out_path, regions, hp, ps, has)
    out, got = emitted(out_path)
    assert n == len(got)
    src = bam_py.read_bam(path)
    assert out["header_bytes"] == src["header_bytes"]
    # restatement, fed with queues in record order (one entry per read that has one)
    hq = [(recs[i][0], int(hp[i])) for i in range(len(recs)) if has[i]]
    pq = [(recs[i][0], int(ps[i])) for i in range(len(recs)) if ps[i]]
    want = pr.phased_bam(src["records"], [(0, 100, 200), (0, 401, 452)], hq, pq)
    assert [w[0] for w in want] == [0, 2, 5, 6, 8, 9, 10, 11]
    assert len(got) == len(want)
    for (i, w_hp, w_ps), r in zip(want, out["records"]):
        s = src["records"][i]
        assert r["core_and_data"] == s["core_and_data"] and r["aux"][: len(s["aux"])] == s["aux"], i
        new = bam_py.parse_aux(r["aux"][len(s["aux"]) :])[0]
        assert new.get("HP") == (None if w_hp is None else ("i", w_hp)), i
        assert new.get("PS") == (None if w_ps is None else ("I", w_ps)), i
    byq = {(q, p): (h, s) for q, p, h, s in got}
    assert byq[("a", 99)] == (("i", 1), ("I", 100)) and byq[("a", 160)] == (("i", 1), ("I", 100))  # QNAME-level first entry
    assert byq[("f", 140)] == (("i", 9), ("I", 100)) and byq[("g", 141)] == (("i", 2), ("I", 77))  # old tags stay
    assert byq[("i", 170)] == (None, None) and byq[("k", 410)] == (None, ("I", 405)) and byq[("j", 400)] == (None, None)


def test_phased_bam_from_a_run(tmp_path):
    """Synthetic genes through the oracle, then the emit against the restatement on every record."""
    syn = host.Synthetic(seed=4, contig_len=80_000, n_contigs=2, platform=0, depth=25.0, n_het=80, n_edit=10, both_strands=0, n_threads=2)
    p = host.params_preset("hifi-masseq", seed=1)
    regions, _ = host.find_regions(syn.reads, p)
    batch = host.BatchView(syn.reads, regions)
    res = ob.run(p, batch, syn.reference.for_reads(syn.reads), mode=0)
    path, out_path = str(tmp_path / "in.bam"), str(tmp_path / "out.bam")
    host.write_bam(path, syn.reads)
    n = host.write_phased_bam(path, out_path, regions, res.hp, res.ps, res.is_fragment)
    src, out = bam_py.read_bam(path), bam_py.read_bam(out_path)
    qn = [r["qname"] for r in src["records"]]
    hq = [(qn[i], int(res.hp[i])) for i in range(len(qn)) if res.is_fragment[i]]
    pq = [(qn[i], int(res.ps[i])) for i in range(len(qn)) if res.ps[i]]
    want = pr.phased_bam(src["records"], [(int(r["tid"]), int(r["start"]), int(r["end"])) for r in regions], hq, pq)
    assert n == len(want) == len(out["records"]) and n > 300
    n_hp = n_ps = 0
    for (i, w_hp, w_ps), r in zip(want, out["records"]):
        s = src["records"][i]
        assert r["core_and_data"] == s["core_and_data"]
        assert r["tags"].get("HP") == (None if w_hp is None else ("i", w_hp)) and r["tags"].get("PS") == (None if w_ps is None else ("I", w_ps))
        n_hp += w_hp is not None
        n_ps += w_ps is not None
    assert n_hp > 100 and n_ps > 100
