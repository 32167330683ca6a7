"""Counts behind the roofline figures of one workload: python profiles/tools/counts.py cfg4"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench  # noqa: E402
from longcallr_b200 import host  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
w, syn, p, regions = bench.make_workload(wl, 0)
eng = host.Engine(p)
eng.set_references(syn.reference.for_reads(syn.reads))
h = eng.upload(host.BatchView(syn.reads, regions))
eng.run_device(h)
eng.run_device(h)
t = eng.timing(h)
r = eng.fetch(h)
s = r.stats
n_pre = (t["pileup_alg_bytes"] - s["n_aligned_bases"] - 16 * (t["n_segments"] + t["n_items"]) - 48 * t["n_tiles"] - s["n_positions"]) // 72
print(wl, {k: t[k] for k in ("n_segments", "n_items", "n_tiles", "ms_pileup", "ms_pileup_kernel", "ms_prep", "ms_fragments", "ms_enum", "ms_phase_kernel", "ms_total")})
print({k: s[k] for k in s}, "pre-candidates", n_pre, "candidates", r.n_cand, "regions", len(regions), "reads", syn.reads.n_reads)
