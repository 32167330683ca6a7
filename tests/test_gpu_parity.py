"""GPU parity: the CUDA path through the C ABI against the CPU oracle (contract mode) on the same inputs."""
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
