"""Writes tests/golden/cfg5_oracle.npz: the oracle's (contract mode) full result for BASELINE config 5 at full size.

The oracle needs ~8 minutes for this single 500x region (2,333 cross_optimize calls on one thread), too long for the GPU
suite, so its output is committed and the GPU test (tests/test_gpu_parity.py::test_full_size_cfg5_against_golden) compares
with it.  Regenerate with:  python tests/golden/make_cfg5_golden.py   (after any change of the oracle or the generator)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
import oracle_binding as ob  # noqa: E402
from longcallr_b200 import host  # noqa: E402


def main():
    w, syn, p, regions = bench.make_workload("cfg5", 0)
    refs = syn.reference.for_reads(syn.reads)
    r = ob.run(p, host.BatchView(syn.reads, regions), refs, mode=0, threads=1)
    out = os.path.join(ROOT, "tests", "golden", "cfg5_oracle.npz")
    np.savez_compressed(out, cand=r.cand, hp=r.hp, ps=r.ps, is_fragment=r.is_fragment, cand_off=r.cand_off, region_status=r.region_status,
                        stats=np.array([r.stats[k] for k in sorted(r.stats)], dtype=np.uint64), stats_keys=np.array(sorted(r.stats)))
    print("wrote", out, os.path.getsize(out), "bytes;", r.n_cand, "candidates", r.stats)


if __name__ == "__main__":
    main()
