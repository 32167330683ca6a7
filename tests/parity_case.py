"""One parity case as a script (the launch-shape knobs are read once per process): python tests/parity_case.py <preset> <platform> <both_strands>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import helpers  # noqa: E402
import oracle_binding as ob  # noqa: E402
from longcallr_b200 import abi, host  # noqa: E402


def main():
    preset, platform, both = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    syn = host.Synthetic(seed=51 + platform, contig_len=150_000, n_contigs=1, platform=platform, depth=30.0, n_het=120, n_edit=20, both_strands=both, n_threads=4)
    p = host.params_preset(preset, seed=8, flags=abi.LCR_FLAG_EMIT_PLANES | abi.LCR_FLAG_EMIT_FRAGMENTS)
    regions, _ = host.find_regions(syn.reads, p)
    refs = syn.reference.for_reads(syn.reads)
    batch = host.BatchView(syn.reads, regions)
    eng = host.Engine(p, device=0)
    eng.set_references(refs)
    got = eng.submit(batch)
    eng.close()
    want = ob.run(p, batch, refs, mode=0)
    assert want.n_cand > 20
    helpers.compare_results(got, want, f"{preset} walk variants")
    print("parity ok", got.n_cand)


if __name__ == "__main__":
    main()
