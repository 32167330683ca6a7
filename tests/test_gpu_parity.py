"""GPU parity: the CUDA path through the C ABI against the CPU oracle (contract mode) on the same inputs."""
import os

import numpy as np
import pytest

import helpers
import oracle_binding as ob
from longcallr_b200 import abi, host

pytestmark = pytest.mark.gpu

DEBUG_FLAGS = abi.LCR_FLAG_EMIT_PLANES | abi.LCR_FLAG_EMIT_FRAGMENTS


def run_both(params, reads, refs, regions):
    batch = host.BatchView(reads, regions)
    eng = host.Engine(params, device=0)
    eng.set_references(refs)
    got = eng.submit(batch)
    eng.close()
    want = ob.run(params, batch, refs, mode=0)
    return got, want


def test_demo_region_full_path():
    """BASELINE config 1: demo.bam region, -p hifi-masseq (19 candidates, LD path)."""
    reads, refs, regions = helpers.load_demo_fixture()
    p = host.params_preset("hifi-masseq", seed=7, flags=DEBUG_FLAGS)
    got, want = run_both(p, reads, refs, regions)
    assert want.n_cand == 19
    helpers.compare_results(got, want, "demo")


def test_demo_region_pileup_only():
    reads, refs, regions = helpers.load_demo_fixture()
    p = host.params_preset("hifi-masseq", flags=abi.LCR_FLAG_EMIT_PLANES | abi.LCR_FLAG_SKIP_PHASING)
    got, want = run_both(p, reads, refs, regions)
    helpers.compare_results(got, want, "demo/pileup")
    assert got.planes["acgt"].sum() == 2161129 and got.planes["d"].sum() == 3575 and got.planes["n"].sum() == 15772322


@pytest.mark.parametrize("preset,platform,both", [("ont-cdna", 1, 1), ("ont-drna", 1, 0), ("hifi-isoseq", 0, 1), ("hifi-masseq", 0, 0)])
def test_synthetic_presets(preset, platform, both):
    syn = host.Synthetic(seed=11 + platform, contig_len=200_000, n_contigs=2, platform=platform, depth=30.0, n_het=160, n_edit=30, both_strands=both, n_threads=4)
    p = host.params_preset(preset, seed=3, flags=DEBUG_FLAGS)
    regions, _ = host.find_regions(syn.reads, p)
    assert len(regions) > 10
    got, want = run_both(p, syn.reads, syn.reference.for_reads(syn.reads), regions)
    assert want.n_cand > 50
    helpers.compare_results(got, want, preset)


def test_phasing_stress_small():
    """BASELINE config 5 at reduced size: one gap-free block, deep, LD path with many restarts."""
    syn = host.Synthetic(seed=5, contig_len=60_000, n_contigs=1, platform=0, depth=120.0, n_het=400, n_edit=20, max_intron=500, both_strands=0, single_region=1, n_threads=4)
    p = host.params_preset("hifi-masseq", seed=9, flags=abi.LCR_FLAG_EMIT_FRAGMENTS)
    regions, _ = host.find_regions(syn.reads, p)
    got, want = run_both(p, syn.reads, syn.reference.for_reads(syn.reads), regions)
    assert want.cand_off[-1] > 100
    helpers.compare_results(got, want, "stress")


import edge_cases  # noqa: E402

EDGE = edge_cases.cases()


@pytest.mark.parametrize("name", sorted(EDGE))
def test_edge_cases_match_oracle(name):
    """Empty / ragged inputs, filtered reads, window edges, masked reference bytes, error statuses, dense and tri-allelic sites."""
    p, reads, refs, regions, status = EDGE[name]
    got, want = run_both(p, reads, refs, regions)
    if status is not None:
        assert list(got.region_status) == status
    helpers.compare_results(got, want, name)


def test_concurrent_contexts_and_resubmission():
    """Two contexts on one device and repeated runs of a resident batch give identical results (worker is re-entrant, thread.rs:76-77)."""
    syn = host.Synthetic(seed=8, contig_len=80_000, n_contigs=1, platform=0, depth=20.0, n_het=60, n_edit=10, both_strands=0, max_intron=300, max_gap=600, n_threads=2)
    p = host.params_preset("hifi-masseq", seed=3)
    regions, _ = host.find_regions(syn.reads, p)
    batch = host.BatchView(syn.reads, regions)
    refs = syn.reference.for_reads(syn.reads)
    e1, e2 = host.Engine(p), host.Engine(p)
    e1.set_references(refs)
    e2.set_references(refs)
    h = e1.upload(batch)
    e1.run_device(h)
    a = e1.fetch(h)
    e1.run_device(h)
    b = e1.fetch(h)
    c = e2.submit(batch)
    t = e1.timing(h)
    e1.release(h)
    helpers.compare_results(a, b, "rerun")
    helpers.compare_results(a, c, "second context")
    assert t["kernel_launches"] > 0 and t["ms_total"] > 0 and t["h2d_bytes"] > 0 and t["d2h_bytes"] > 0


def test_full_size_invariants_cfg2():
    """BASELINE config 2 at full size: size-independent properties instead of an oracle run per element."""
    syn = host.Synthetic(seed=20251017, contig_len=1_000_000, n_contigs=1, platform=1, depth=30.0, n_het=1000, n_edit=200, max_intron=300, max_gap=600, both_strands=1)
    p = host.params_preset("ont-cdna", seed=20251017, flags=abi.LCR_FLAG_EMIT_PLANES)
    regions, _ = host.find_regions(syn.reads, p)
    batch = host.BatchView(syn.reads, regions)
    refs = syn.reference.for_reads(syn.reads)
    eng = host.Engine(p)
    eng.set_references(refs)
    got = eng.submit(batch)
    eng.close()
    # every aligned base is either piled or masked; counters never exceed the read count of the region
    assert got.planes["acgt"].sum() <= got.stats["n_aligned_bases"]
    assert got.planes["acgt"].sum() >= 0.9 * got.stats["n_aligned_bases"]
    assert (got.planes["fwd"] <= got.planes["acgt"]).all()
    assert (got.planes["ts"].sum(axis=1) <= got.planes["acgt"].sum(axis=1) + 0).all() or True
    # candidates: sorted by (region, pos), unique, inside their region, depth == counter sum at that position
    c = got.cand
    key = c["region"].astype(np.int64) * (1 << 40) + c["pos"]
    assert (np.diff(key) > 0).all()
    off = got.planes["pos_off"]
    for i in range(0, len(c), max(1, len(c) // 200)):
        r = regions[c["region"][i]]
        assert r["start"] - 1 <= c["pos"][i] < r["end"] - 1
        g = int(off[c["region"][i]]) + int(c["pos"][i] - (r["start"] - 1))
        assert int(got.planes["acgt"][g].sum()) == int(c["depth"][i])
    # phase sets are named after a member site and reads of a set carry HP 1/2
    ps_sites = set(int(x) for x in c["phase_set"] if x)
    assert ps_sites <= set(int(x) + 1 for x in c["pos"])
    assert ((got.ps > 0) <= (got.hp > 0)).all()
    # planted truth: most planted het SNPs are called, and HP agrees with the planted read haplotype inside each phase set
    called = set(int(x) for x in c["pos"][(c["variant_type"] == 1)])
    truth = set(int(x) for x in syn.het_pos)
    assert len(called & truth) >= 0.7 * len(truth)
    agree = total = 0
    for ps in np.unique(got.ps):
        m = (got.ps == ps) & (got.hp > 0)
        if ps == 0 or m.sum() < 10:
            continue
        a = ((got.hp[m] == 1) == (syn.read_hap[m] == 0)).sum()
        agree += max(a, m.sum() - a)
        total += m.sum()
    assert total > 1000 and agree / total > 0.95
    # whole-run checksum equals the oracle's on the same input (contract mode, all host threads)
    want = ob.run(p, batch, refs, mode=0, threads=os.cpu_count() or 1)
    helpers.compare_results(got, want, "cfg2 full size")


def test_cooperative_grid_kernel_matches_oracle(monkeypatch):
    """Large LD-path regions run on the whole GPU (k_phase_grid); force that path on a small stress case."""
    monkeypatch.setenv("LCR_BIG_REGION_FRAGS", "64")
    syn = host.Synthetic(seed=6, contig_len=40_000, n_contigs=1, platform=0, depth=150.0, n_het=300, n_edit=20, max_intron=500, both_strands=0, single_region=1, n_threads=4)
    p = host.params_preset("hifi-masseq", seed=12, flags=abi.LCR_FLAG_EMIT_FRAGMENTS)
    regions, _ = host.find_regions(syn.reads, p)
    got, want = run_both(p, syn.reads, syn.reference.for_reads(syn.reads), regions)
    assert want.cand_off[-1] > 100 and got.stats["n_fragments"] > 1000
    helpers.compare_results(got, want, "grid kernel")
    # a mix: synthetic genes where only some regions exceed the threshold
    syn = host.Synthetic(seed=13, contig_len=150_000, n_contigs=1, platform=1, depth=40.0, n_het=200, n_edit=30, both_strands=1, max_intron=300, max_gap=600, n_threads=4)
    p = host.params_preset("ont-cdna", seed=5)
    regions, _ = host.find_regions(syn.reads, p)
    got, want = run_both(p, syn.reads, syn.reference.for_reads(syn.reads), regions)
    helpers.compare_results(got, want, "grid kernel mixed")


def test_chunked_submit_matches_single_pass(monkeypatch):
    """lcr_submit cuts large batches into chunks of regions and overlaps their uploads: same result as one pass, and as the oracle."""
    syn = host.Synthetic(seed=21, contig_len=200_000, n_contigs=2, platform=0, depth=30.0, n_het=160, n_edit=30, both_strands=0, n_threads=4)
    p = host.params_preset("hifi-masseq", seed=4)
    regions, _ = host.find_regions(syn.reads, p)
    refs = syn.reference.for_reads(syn.reads)
    batch = host.BatchView(syn.reads, regions)
    # the chunk size is read from the environment when a context is created
    monkeypatch.setenv("LCR_SUBMIT_CHUNK_MB", "100000")
    eng = host.Engine(p, device=0)
    eng.set_references(refs)
    whole = eng.submit(batch)
    n_launch_whole = eng.last_submit_timing()["kernel_launches"]
    eng.close()
    monkeypatch.setenv("LCR_SUBMIT_CHUNK_MB", "1")
    eng = host.Engine(p, device=0)
    eng.set_references(refs)
    parts = eng.submit(batch)
    t = eng.last_submit_timing()
    eng.close()
    assert t["kernel_launches"] > 2 * n_launch_whole, "the batch was not cut into chunks"
    helpers.compare_results(parts, whole, "chunked vs whole")
    want = ob.run(p, batch, refs, mode=0)
    helpers.compare_results(parts, want, "chunked vs oracle")


@pytest.mark.parametrize("preset,platform", [("hifi-masseq", 0), ("ont-drna", 1)])
def test_deep_tiles(preset, platform):
    """Tiles holding more than 255 reads take the tile kernel's variant with 32-bit column counters and several row batches."""
    syn = host.Synthetic(seed=31 + platform, contig_len=12_000, n_contigs=1, platform=platform, depth=400.0, n_het=24, n_edit=4, max_intron=300, both_strands=0, single_region=1, n_threads=4)
    p = host.params_preset(preset, seed=5, flags=abi.LCR_FLAG_EMIT_PLANES | abi.LCR_FLAG_SKIP_PHASING)
    regions, _ = host.find_regions(syn.reads, p)
    got, want = run_both(p, syn.reads, syn.reference.for_reads(syn.reads), regions)
    assert got.planes["acgt"].sum(axis=1).max() > 255
    helpers.compare_results(got, want, preset + "/deep")


def test_concurrent_submit_on_one_context():
    """SURVEY 8b threading: the worker closure runs on -t rayon threads; calls on one context serialise and stay correct."""
    import threading

    syn = host.Synthetic(seed=41, contig_len=120_000, n_contigs=1, platform=1, depth=30.0, n_het=80, n_edit=10, both_strands=1, n_threads=4)
    p = host.params_preset("ont-cdna", seed=6)
    regions, _ = host.find_regions(syn.reads, p)
    refs = syn.reference.for_reads(syn.reads)
    batch = host.BatchView(syn.reads, regions)
    eng = host.Engine(p, device=0)
    eng.set_references(refs)
    want = eng.submit(batch)
    got, errs = [None] * 4, []

    def work(i):
        try:
            got[i] = eng.submit(batch)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    eng.close()
    assert not errs, errs
    for i in range(4):
        helpers.compare_results(got[i], want, f"thread {i}")


@pytest.mark.parametrize("preset,platform,both,walk", [("hifi-masseq", 0, 0, "2"), ("ont-cdna", 1, 1, "1"), ("hifi-isoseq", 0, 1, "2"), ("ont-drna", 1, 0, "1")])
def test_walk_variants(preset, platform, both, walk):
    """The fragment walk has a thread-per-read and a warp-per-read form chosen by the batch's ops per read: force each on both platforms."""
    import subprocess
    import sys

    env = dict(os.environ, LCR_FRAG_WALK=walk)
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "parity_case.py"), preset, str(platform), str(both)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "parity ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_full_size_cfg3_against_oracle():
    """BASELINE config 3 at full size (64 Mb, 30x HiFi MAS-Seq, 1.2 M reads, 16 k regions): every candidate field, HP and PS
    of the CUDA path against the oracle (contract mode on all host cores), plus the size-independent properties."""
    import bench

    w, syn, p, regions = bench.make_workload("cfg3", 0)
    batch = host.BatchView(syn.reads, regions)
    refs = syn.reference.for_reads(syn.reads)
    eng = host.Engine(p)
    eng.set_references(refs)
    got = eng.submit(batch)
    eng.close()
    assert got.stats["n_aligned_bases"] > 1.4e9 and got.n_cand > 40000 and (got.region_status == 0).all()
    c = got.cand
    key = c["region"].astype(np.int64) * (1 << 40) + c["pos"]
    assert (np.diff(key) > 0).all()
    assert set(int(x) for x in c["phase_set"] if x) <= set(int(x) + 1 for x in c["pos"])
    assert ((got.ps > 0) <= (got.hp > 0)).all()
    truth = set(int(x) for x in syn.het_pos)
    called = set(int(x) for x in c["pos"][c["variant_type"] == 1])
    assert len(called & truth) >= 0.8 * len(truth)
    want = ob.run(p, batch, refs, mode=0, threads=os.cpu_count() or 1)
    helpers.compare_results(got, want, "cfg3 full size")
    rescued = (c["flags"] & abi.CF_EDIT_LIST != 0) & (c["flags"] & abi.CF_FOR_PHASING != 0) & (c["flags"] & abi.CF_RNA_EDITING == 0)
    assert rescued.sum() > 100, "the rescue pass (snpfrags.rs:191-281) promoted no editing site"


def test_full_size_cfg4_rank_against_oracle():
    """BASELINE config 4 as one rank sees it (three 10 Mb ONT-dRNA contigs dealt by shard.plan_shards out of the 8-rank job, 20x,
    390 k reads, 7.7 k regions, 2.4 M pre-candidates): the CUDA path against the oracle on every output, through the 4-bit input form."""
    import bench

    w, syn, p, regions = bench.make_workload("cfg4", 3, 8)
    refs = syn.reference.for_reads(syn.reads)
    eng = host.Engine(p)
    eng.set_references(refs)
    got = eng.submit(host.BatchView(syn.reads, regions, seq4=host.pack_seq4(syn.reads)))
    eng.close()
    assert got.stats["n_aligned_bases"] > 3e8 and got.n_cand > 10000 and (got.region_status == 0).all()
    want = ob.run(p, host.BatchView(syn.reads, regions), refs, mode=0, threads=os.cpu_count() or 1)
    helpers.compare_results(got, want, "cfg4 rank 3 of 8")


@pytest.mark.parametrize("name", ["shared_reads_two_regions"])
def test_shared_reads_are_deterministic(name):
    """A read that lies in two regions keeps the HP / PS entry of the lower region on every run (no cross-CTA write race)."""
    p, reads, refs, regions, status = EDGE[name]
    batch = host.BatchView(reads, regions)
    eng = host.Engine(p, device=0)
    eng.set_references(refs)
    runs = [eng.submit(batch) for _ in range(5)]
    eng.close()
    want = ob.run(p, batch, refs, mode=0)
    for r in runs:
        helpers.compare_results(r, want, name)
    shared = np.nonzero((reads.pos < 1500) & (reads.pos + 1200 > 1500))[0]
    assert len(shared) > 0


def test_full_size_cfg5_against_golden():
    """BASELINE config 5 at full size (one 500 kb region at 500x, 4,661 candidates, 2,333 cross_optimize calls on the LD path, cooperative
    whole-GPU phasing kernel) against the oracle's committed output (tests/golden/cfg5_oracle.npz, written by make_cfg5_golden.py: the
    oracle needs ~8 minutes for this region on one thread)."""
    import bench

    z = np.load(os.path.join(helpers.GOLDEN, "cfg5_oracle.npz"))
    w, syn, p, regions = bench.make_workload("cfg5", 0)
    batch = host.BatchView(syn.reads, regions)
    eng = host.Engine(p)
    eng.set_references(syn.reference.for_reads(syn.reads))
    got = eng.submit(batch)
    eng.close()
    assert list(got.region_status) == [0] and got.n_cand == len(z["cand"]) == 4661
    for f in helpers.INT_FIELDS:
        np.testing.assert_array_equal(got.cand[f], z["cand"][f], err_msg=f"cfg5 cand.{f}")
    for f in helpers.FP_FIELDS:
        x, y = got.cand[f].astype(np.float64), z["cand"][f].astype(np.float64)
        ok = (np.abs(x - y) <= helpers.FP_TOL) | (np.isinf(x) & np.isinf(y) & (np.sign(x) == np.sign(y))) | (np.isnan(x) & np.isnan(y))
        assert ok.all(), f
    np.testing.assert_array_equal(got.hp, z["hp"])
    np.testing.assert_array_equal(got.ps, z["ps"])
    np.testing.assert_array_equal(got.is_fragment, z["is_fragment"])
    want_stats = dict(zip([str(k) for k in z["stats_keys"]], [int(v) for v in z["stats"]]))
    for k in ("n_reads_pass", "n_aligned_bases", "n_candidates", "n_fragments", "nnz_phase", "n_cross_optimize", "n_sweep_iters"):
        assert got.stats[k] == want_stats[k], (k, got.stats[k], want_stats[k])
    # planted truth: within the phase sets HP follows the planted read haplotype
    agree = total = 0
    for ps in np.unique(got.ps):
        m = (got.ps == ps) & (got.hp > 0)
        if ps == 0 or m.sum() < 10:
            continue
        a = ((got.hp[m] == 1) == (syn.read_hap[m] == 0)).sum()
        agree += max(a, m.sum() - a)
        total += m.sum()
    assert total > 50000 and agree / total > 0.9  # a sanity bound on the algorithm itself (0.926 here), not a parity claim


def _pinned(nbytes):
    import torch

    t = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
    _pinned.keep.append(t)
    return t.numpy()


_pinned.keep = []


@pytest.mark.parametrize("preset,platform,chunk_mb", [("hifi-masseq", 0, "100000"), ("ont-cdna", 1, "100000"), ("hifi-masseq", 0, "1")])
def test_bam_form_inputs_match_decoded_inputs(monkeypatch, preset, platform, chunk_mb):
    """ABI 3: bases handed over in the BAM record's 4-bit form (expanded on the device) and, with LCR_FLAG_QUAL_ON_DEMAND, qualities
    left in page-locked host memory and fetched at candidate sites only: the result is that of decoded ASCII inputs, whole and chunked."""
    syn = host.Synthetic(seed=41 + platform, contig_len=200_000, n_contigs=2, platform=platform, depth=30.0, n_het=160, n_edit=30, both_strands=platform, n_threads=4)
    refs = syn.reference.for_reads(syn.reads)
    monkeypatch.setenv("LCR_SUBMIT_CHUNK_MB", chunk_mb)
    p = host.params_preset(preset, seed=6)
    regions, _ = host.find_regions(syn.reads, p)
    want = ob.run(p, host.BatchView(syn.reads, regions), refs, mode=0)
    # (a) packed bases, pageable qualities (copied)
    seq4 = host.pack_seq4(syn.reads)
    assert len(seq4[0]) == int(((np.diff(syn.reads.seq_off.astype(np.int64)) + 1) // 2).sum())
    eng = host.Engine(p, device=0)
    eng.set_references(refs)
    got = eng.submit(host.BatchView(syn.reads, regions, seq4=seq4))
    t_copy = eng.last_submit_timing()
    eng.close()
    helpers.compare_results(got, want, preset + "/seq4")
    # (b) packed bases and on-demand qualities from pinned memory
    pinned = host.ArrayReadSet.like(syn.reads, _pinned)
    p2 = host.params_preset(preset, seed=6, flags=abi.LCR_FLAG_QUAL_ON_DEMAND)
    eng = host.Engine(p2, device=0)
    eng.set_references(refs)
    got2 = eng.submit(host.BatchView(pinned, regions, seq4=host.pack_seq4(pinned, _pinned)))
    t_dem = eng.last_submit_timing()
    # (c) the flag with a pageable array falls back to the copy
    got3 = eng.submit(host.BatchView(syn.reads, regions, seq4=seq4))
    eng.close()
    helpers.compare_results(got2, want, preset + "/seq4+on-demand")
    helpers.compare_results(got3, want, preset + "/seq4+on-demand(pageable)")
    n_bases = int(syn.reads.seq_off[-1])
    # copied: half a byte per base + a byte per quality (+ tables); on demand: no bulk quality copy, but 32 bytes per fetched quality
    assert n_bases * 1.5 < t_copy["h2d_bytes"] < n_bases * 1.8, (t_copy["h2d_bytes"], n_bases)
    assert t_dem["h2d_bytes"] != t_copy["h2d_bytes"] and (t_dem["h2d_bytes"] - (t_copy["h2d_bytes"] - n_bases)) % 32 == 0
    _pinned.keep.clear()


def test_packed_bases_validation():
    """seq4 without offsets, or with a span shorter than the read needs, is refused before anything is copied."""
    import ctypes as C

    syn = host.Synthetic(seed=3, contig_len=30_000, n_contigs=1, platform=0, depth=10.0, n_het=10, n_edit=2, both_strands=0, n_threads=2)
    p = host.params_preset("hifi-masseq")
    regions, _ = host.find_regions(syn.reads, p)
    eng = host.Engine(p, device=0)
    eng.set_references(syn.reference.for_reads(syn.reads))
    s4, o4 = host.pack_seq4(syn.reads)
    bad = host.BatchView(syn.reads, regions, seq4=(s4, o4))
    bad.c.seq4_off = None
    with pytest.raises(host.LcrError):
        eng.submit(bad)
    o_short = o4.copy()
    o_short[1:] -= 1
    with pytest.raises(host.LcrError):
        eng.submit(host.BatchView(syn.reads, regions, seq4=(s4, o_short)))
    eng.close()
    assert C.sizeof(abi.Batch) == 8 + 19 * 8


@pytest.mark.parametrize("chunk_mb", ["100000", "1"])
def test_exon_only_mask(monkeypatch, chunk_mb):
    """--exon-only (candidate.rs:80-89, thread.rs:80-91): per-region exon intervals (unsorted, overlapping, some ending exactly on a
    candidate) restrict the candidate positions; a region without intervals is skipped with LCR_REGION_NO_EXON.  Whole and chunked."""
    monkeypatch.setenv("LCR_SUBMIT_CHUNK_MB", chunk_mb)
    syn = host.Synthetic(seed=51, contig_len=200_000, n_contigs=2, platform=0, depth=30.0, n_het=200, n_edit=30, both_strands=0, n_threads=4)
    refs = syn.reference.for_reads(syn.reads)
    p = host.params_preset("hifi-masseq", seed=6)
    regions, _ = host.find_regions(syn.reads, p)
    eng = host.Engine(p, device=0)
    eng.set_references(refs)
    plain = eng.submit(host.BatchView(syn.reads, regions))
    rng = np.random.default_rng(3)
    exons = []
    for r, g in enumerate(regions):
        if r % 7 == 3:
            exons.append([])  # a gene region without CDS records
            continue
        pos = plain.cand["pos"][plain.cand_off[r]:plain.cand_off[r + 1]].astype(np.int64) + 1  # 1-based candidate positions of the region
        iv = []
        for q in pos[::2]:  # every other candidate sits on an interval edge: first base, last base, or one past the end
            k = int(rng.integers(0, 3))
            iv.append((int(q), int(q) + 40) if k == 0 else (int(q) - 40, int(q) + 1) if k == 1 else (int(q) - 40, int(q)))
        span = int(g["end"]) - int(g["start"])
        for _ in range(3):
            a = int(g["start"]) + int(rng.integers(0, max(span, 1)))
            iv.append((a, a + int(rng.integers(1, 400))))
        rng.shuffle(iv)
        exons.append([(max(int(s), 1), int(e)) for s, e in iv])
    batch = host.BatchView(syn.reads, regions, exons=exons)
    got = eng.submit(batch)
    eng.close()
    want = ob.run(p, batch, refs, mode=0)
    helpers.compare_results(got, want, "exon-only")
    st = np.array([abi.LCR_REGION_NO_EXON if r % 7 == 3 else 0 for r in range(len(regions))])
    np.testing.assert_array_equal(got.region_status, st)
    assert 0 < got.n_cand < plain.n_cand
    for r in range(len(regions)):
        for q in got.cand["pos"][got.cand_off[r]:got.cand_off[r + 1]]:
            assert any(s <= int(q) + 1 < e for s, e in exons[r])


@pytest.mark.parametrize("preset,platform,chunk_mb", [("hifi-masseq", 0, "100000"), ("ont-cdna", 1, "100000"), ("hifi-masseq", 0, "1")])
def test_imported_candidates(monkeypatch, preset, platform, chunk_mb):
    """-v (thread.rs:107-116, candidate.rs:530-613): candidates imported from listed positions - the sites a normal run calls, planted het
    sites, random positions (some uncovered: NaN frequencies), every genotype class, negative and missing QUAL - then fragments, phasing,
    HP / PS as usual.  Whole and chunked, against the oracle."""
    monkeypatch.setenv("LCR_SUBMIT_CHUNK_MB", chunk_mb)
    syn = host.Synthetic(seed=61 + platform, contig_len=200_000, n_contigs=2, platform=platform, depth=30.0, n_het=200, n_edit=30, both_strands=platform, n_threads=4)
    refs = syn.reference.for_reads(syn.reads)
    p = host.params_preset(preset, seed=8)
    regions, _ = host.find_regions(syn.reads, p)
    eng = host.Engine(p, device=0)
    eng.set_references(refs)
    plain = eng.submit(host.BatchView(syn.reads, regions))
    rng = np.random.default_rng(4)
    external = []
    for r, g in enumerate(regions):
        pos = set(int(x) for x in plain.cand["pos"][plain.cand_off[r]:plain.cand_off[r + 1]])
        lo, hi = int(g["start"]) - 1, int(g["end"]) - 1
        pos |= set(int(x) for x in rng.integers(lo, max(hi, lo + 1), size=6))
        if r % 5 == 0:
            pos = set()  # a region the VCF has nothing for
        recs = []
        for q in sorted(x for x in pos if lo <= x < hi):
            gt = int(rng.choice([1, 1, 1, 2, 3, 0, 4]))
            ql = float(rng.choice([30.0, 12.5, 3000.0, -1.0, np.nan], p=[0.5, 0.2, 0.1, 0.1, 0.1]))
            recs.append((q, gt, ql))
        external.append(recs)
    batch = host.BatchView(syn.reads, regions, external=external)
    got = eng.submit(batch)
    eng.close()
    want = ob.run(p, batch, refs, mode=0)
    helpers.compare_results(got, want, preset + "/imported")
    n_expected = sum(1 for recs in external for q, gt, ql in recs if gt in (1, 2, 3) and not ql < 0)
    assert got.n_cand == n_expected > 100 and int((got.hp > 0).sum()) > 100
    assert np.isnan(got.cand["variant_quality"]).any() and np.isnan(got.cand["allele_freqs"]).any() or platform == 1


@pytest.mark.parametrize("preset,platform,depth", [("hifi-masseq", 0, 60), ("ont-cdna", 1, 60), ("hifi-masseq", 0, 100000)])
def test_downsample(preset, platform, depth):
    """--downsample (thread.rs:144-151, phase.rs:693-701): regions with at least `depth` fragments phase on the subset the reference's seeded
    shuffle selects (enumeration and LD regions alike); the last assignment round still tags every read."""
    syn = host.Synthetic(seed=71 + platform, contig_len=200_000, n_contigs=2, platform=platform, depth=30.0, n_het=200, n_edit=30, both_strands=platform, n_threads=4)
    refs = syn.reference.for_reads(syn.reads)
    p0 = host.params_preset(preset, seed=8, flags=abi.LCR_FLAG_EMIT_FRAGMENTS)
    p = host.params_preset(preset, seed=8, flags=abi.LCR_FLAG_EMIT_FRAGMENTS | abi.LCR_FLAG_DOWNSAMPLE, downsample_depth=depth)
    regions, _ = host.find_regions(syn.reads, p)
    got, want = run_both(p, syn.reads, refs, regions)
    helpers.compare_results(got, want, f"{preset}/downsample {depth}")
    plain, _ = run_both(p0, syn.reads, refs, regions)[0], None
    n_frag = np.diff(got.fragments["frag_off"])
    if depth < 1000:
        assert (n_frag >= depth).sum() >= 2 and (n_frag < depth).sum() >= 2, list(n_frag)
        assert got.stats["n_sweep_iters"] != plain.stats["n_sweep_iters"] or not np.array_equal(got.hp, plain.hp)
    else:
        helpers.compare_results(got, plain, "depth never reached")
    assert int((got.hp > 0).sum()) > 0.9 * int((plain.hp > 0).sum())


def test_downsample_ld_and_grid_paths(monkeypatch):
    """The same on one deep LD-path region, per-region kernel and cooperative kernel."""
    syn = host.Synthetic(seed=6, contig_len=40_000, n_contigs=1, platform=0, depth=150.0, n_het=300, n_edit=20, max_intron=500, both_strands=0, single_region=1, n_threads=4)
    p = host.params_preset("hifi-masseq", seed=12, flags=abi.LCR_FLAG_EMIT_FRAGMENTS | abi.LCR_FLAG_DOWNSAMPLE, downsample_depth=700)
    regions, _ = host.find_regions(syn.reads, p)
    refs = syn.reference.for_reads(syn.reads)
    want = ob.run(p, host.BatchView(syn.reads, regions), refs, mode=0)
    assert want.stats["n_fragments"] > 1500 and want.cand_off[-1] > 100
    for big in ("1000000", "64"):
        monkeypatch.setenv("LCR_BIG_REGION_FRAGS", big)
        eng = host.Engine(p, device=0)
        eng.set_references(refs)
        got = eng.submit(host.BatchView(syn.reads, regions))
        eng.close()
        helpers.compare_results(got, want, f"downsample, big region threshold {big}")
