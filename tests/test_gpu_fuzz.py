"""GPU differential fuzz: the CUDA path through the C ABI against the oracle (contract mode) on the same random small regions the two
restatements are fuzzed against each other with (tests/test_oracle_crosscheck.py: messy CIGARs with H/S/I/D/N/=/X, N and lower-case
bytes, quality 1..40, filtered reads, poly-A tails), over every output: planes, candidates, fragment matrix, HP, PS, counters."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import helpers
import oracle_binding as ob
import test_oracle_crosscheck as xc
from longcallr_b200 import abi, host

pytestmark = pytest.mark.gpu

FLAGS = abi.LCR_FLAG_EMIT_PLANES | abi.LCR_FLAG_EMIT_FRAGMENTS
ENGINES = {}


def engine(preset, **over):
    key = (preset, tuple(sorted(over.items())))
    if key not in ENGINES:
        p = host.params_preset(preset, seed=5, min_read_length=30, min_depth=4, flags=FLAGS | (abi.LCR_FLAG_DOWNSAMPLE if "downsample_depth" in over else 0), **over)
        ENGINES[key] = (p, host.Engine(p, device=0))
    return ENGINES[key]


@settings(max_examples=int(os.environ.get("LCR_GPU_FUZZ_EXAMPLES", "250")), deadline=None, suppress_health_check=list(HealthCheck), derandomize=not os.environ.get("LCR_FUZZ_RANDOM"), database=None)
@given(xc.regions(), st.booleans(), st.booleans(), st.sampled_from([0, 0, 6, 15]))
def test_random_regions_match_oracle(case, packed, two_regions, ds_depth):
    preset, ref, recs, start, end = case
    if packed:  # a BAM record cannot hold a lower-case base: what does not pack decodes as N (both sides see the decoded record)
        recs = [dict(r, seq="".join(c if c in "=ACMGRSVTWYHKDBN" else "N" for c in r["seq"])) for r in recs]
    reads = helpers.make_reads(len(ref), recs)
    n = reads.n_reads
    if two_regions and end - start > 200:  # the window cut in two regions that share their reads (lowest region wins the HP / PS of a read)
        mid = (start + end) // 2
        regions = np.zeros(2, dtype=abi.REGION_DTYPE)
        regions[0] = (0, start, mid, 0, n)
        regions[1] = (0, mid, end, 0, n)
    else:
        regions = helpers.one_region(start, end, n)
    p, eng = engine(preset, **(dict(downsample_depth=ds_depth) if ds_depth else {}))  # ds_depth > 0: --downsample with that depth
    eng.set_reference(0, ref)
    batch = host.BatchView(reads, regions, seq4=host.pack_seq4(reads) if packed and n else None)
    got = eng.submit(batch)
    want = ob.run(p, host.BatchView(reads, regions), [ref], mode=0)
    helpers.compare_results(got, want, f"fuzz {preset} packed={packed} two={two_regions}")


@settings(max_examples=int(os.environ.get("LCR_GPU_FUZZ_EXAMPLES", "250")) // 3, deadline=None, suppress_health_check=list(HealthCheck), derandomize=not os.environ.get("LCR_FUZZ_RANDOM"), database=None)
@given(xc.regions(), st.lists(st.tuples(st.integers(0, 400), st.integers(1, 60)), min_size=0, max_size=5),
       st.lists(st.tuples(st.integers(0, 300), st.integers(0, 4), st.one_of(st.floats(-5, 60, width=32), st.just(float("nan")))), min_size=0, max_size=12, unique_by=lambda t: t[0]))
def test_random_regions_alternate_entries(case, raw_exons, raw_ext):
    """The same regions through --exon-only (no interval: the region is skipped) and through -v (no record: no candidate)."""
    preset, ref, recs, start, end = case
    reads = helpers.make_reads(len(ref), recs)
    region = helpers.one_region(start, end, reads.n_reads)
    p, eng = engine(preset)
    eng.set_reference(0, ref)
    exons = [[(start + a, start + a + ln) for a, ln in raw_exons]]
    ext = [sorted((start - 1 + a, gt, q) for a, gt, q in raw_ext if start - 1 + a < end - 1)]
    for kw in (dict(exons=exons), dict(external=ext)):
        batch = host.BatchView(reads, region, **kw)
        helpers.compare_results(eng.submit(batch), ob.run(p, batch, [ref], mode=0), f"fuzz {preset} {list(kw)[0]}")
