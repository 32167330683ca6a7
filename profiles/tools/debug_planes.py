import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import helpers, oracle_binding as ob
from longcallr_b200 import abi, host

def run(params, reads, refs, regions):
    batch = host.BatchView(reads, regions)
    eng = host.Engine(params, device=0)
    eng.set_references(refs)
    got = eng.submit(batch)
    eng.close()
    want = ob.run(params, batch, refs, mode=0)
    return got, want

which = sys.argv[1] if len(sys.argv) > 1 else "demo"
if which == "demo":
    reads, refs, regions = helpers.load_demo_fixture()
    p = host.params_preset("hifi-masseq", flags=abi.LCR_FLAG_EMIT_PLANES | abi.LCR_FLAG_SKIP_PHASING)
else:
    platform = 1 if which.startswith("ont") else 0
    syn = host.Synthetic(seed=11 + platform, contig_len=200_000, n_contigs=2, platform=platform, depth=30.0, n_het=160, n_edit=30, both_strands=1, n_threads=4)
    p = host.params_preset(which, seed=3, flags=abi.LCR_FLAG_EMIT_PLANES | abi.LCR_FLAG_SKIP_PHASING)
    regions, _ = host.find_regions(syn.reads, p)
    reads, refs = syn.reads, syn.reference.for_reads(syn.reads)
got, want = run(p, reads, refs, regions)
print("status", got.region_status[:5], want.region_status[:5], "n_cand", got.n_cand, want.n_cand)
print("stats", got.stats, want.stats)
for k in ("acgt", "fwd", "d", "n", "ts"):
    a, b = got.planes[k], want.planes[k]
    bad = np.argwhere(a != b)
    print(k, a.shape, "sum got/want", int(a.sum()), int(b.sum()), "mismatches", len(bad))
    for idx in bad[:12]:
        idx = tuple(idx)
        print("   ", idx, "got", a[idx], "want", b[idx])
    if len(bad):
        pos = np.unique(bad[:, 0])
        print("   first mismatching positions:", pos[:40], " mod 512:", (pos[:40] % 512))
