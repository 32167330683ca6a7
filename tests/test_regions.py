"""Isolated-region discovery: the host implementation against the per-position restatement of util.rs:236-332 on CPU,
and the device implementation (lcr_discover_regions) against the host one on the GPU."""
import numpy as np
import pytest

import region_cases as rc  # puts oracle/ on sys.path
import py_restatement as pr  # noqa: I001
from longcallr_b200 import host

TRUNC = {"truncation_splits_a_run": 50, "truncation_leaves_single_positions": 8}


def want_regions(lens, reads, p, truncation, cov):
    out = []
    for t, L in enumerate(lens):
        for s, e, m in pr.find_isolated_regions(L, rc.restated(lens, reads, t), p.min_mapq, p.min_read_length, p.divergence, truncation, cov):
            out.append((t, s, e, m))
    return out


def check_host(name, lens, reads, truncation, cov):
    p = host.params_preset("hifi-masseq")
    p.min_read_length = 1
    rs, reads = rc.read_set(lens, reads)
    regions, maxcov = host.find_regions(rs, p, truncation, cov)
    got = [(int(r["tid"]), int(r["start"]), int(r["end"]), int(m)) for r, m in zip(regions, maxcov)]
    assert got == want_regions(lens, reads, p, truncation, cov), name
    # read ranges: a superset of every read overlapping [start-1, end-1) on the contig
    for r in regions:
        for i in range(rs.n_reads):
            if rs.tid[i] != r["tid"]:
                continue
            span = sum(ln for op, ln in reads[i][2] if op in rc.REF_OPS)
            end = rs.pos[i] + max(span, 1)
            if rs.pos[i] < r["end"] and end > r["start"]:
                assert r["read_begin"] <= i < r["read_end"], (name, i)
    return rs, p, regions, maxcov


@pytest.mark.parametrize("name", sorted(rc.cases()))
def test_host_regions_match_restatement(name):
    lens, reads = rc.cases()[name]
    check_host(name, lens, reads, name in TRUNC, TRUNC.get(name, 200000))


def test_quirks_are_what_the_reference_loop_does():
    """The two behaviours a 'maximal covered runs' reading would get wrong."""
    p = host.params_preset("hifi-masseq")
    p.min_read_length = 1
    lens, reads = rc.cases()["single_position_run_joins_next"]
    assert want_regions(lens, reads, p, False, 0) == [(0, 6, 151, 1)]
    lens, reads = rc.cases()["truncation_splits_a_run"]
    w = want_regions(lens, reads, p, True, 50)
    assert w[0][:3] == (0, 101, 151) and w[0][3] == 70 > 50  # the push happens at the first cut position, after its depth was folded into max_coverage


def test_host_regions_random():
    rng = np.random.default_rng(5)
    for k in range(40):
        lens, reads = rc.random_case(rng)
        check_host(f"random{k}", lens, reads, bool(k & 1), int(rng.integers(2, 9)))


@pytest.mark.gpu
def test_device_regions_match_host():
    p0 = host.params_preset("hifi-masseq")
    p0.min_read_length = 1
    eng = host.Engine(p0, device=0)
    rng = np.random.default_rng(9)
    todo = [(n, c[0], c[1], n in TRUNC, TRUNC.get(n, 200000)) for n, c in sorted(rc.cases().items())]
    todo += [(f"random{k}", *rc.random_case(rng), bool(k & 1), int(rng.integers(2, 9))) for k in range(60)]
    todo += [("random_big", *rc.random_case(rng, n_contigs=5, contig_len=300_000, n_reads=20_000), True, 6)]
    for name, lens, reads, truncation, cov in todo:
        rs, _ = rc.read_set(lens, reads)
        want_r, want_m = host.find_regions(rs, p0, truncation, cov)
        got_r, got_m = eng.discover_regions(rs, truncation, cov)
        assert len(got_r) == len(want_r), name
        for f in ("tid", "start", "end", "read_begin", "read_end"):
            np.testing.assert_array_equal(got_r[f], want_r[f], err_msg=f"{name}.{f}")
        np.testing.assert_array_equal(got_m, want_m, err_msg=name)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("preset,platform", [("hifi-masseq", 0), ("ont-drna", 1)])
def test_device_regions_synthetic(preset, platform):
    """Synthetic genes at depth 30 (the generator of the bench workloads), with and without truncation."""
    syn = host.Synthetic(seed=3 + platform, contig_len=2_000_000, n_contigs=3, platform=platform, depth=30.0, n_het=500, n_edit=100, both_strands=platform, n_threads=4)
    p = host.params_preset(preset)
    eng = host.Engine(p, device=0)
    for truncation, cov in ((False, 200000), (True, 40)):
        want_r, want_m = host.find_regions(syn.reads, p, truncation, cov)
        got_r, got_m, ms = eng.discover_regions(syn.reads, truncation, cov, with_time=True)
        assert len(want_r) > 100 and ms > 0
        for f in ("tid", "start", "end", "read_begin", "read_end"):
            np.testing.assert_array_equal(got_r[f], want_r[f], err_msg=f)
        np.testing.assert_array_equal(got_m, want_m)
    eng.close()
