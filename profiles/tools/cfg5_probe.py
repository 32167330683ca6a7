import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from longcallr_b200 import host, abi
t = time.time()
syn = host.Synthetic(seed=20251017, contig_len=500_000, n_contigs=1, platform=0, depth=500.0, n_het=5000, n_edit=0, max_intron=500, max_gap=600, both_strands=0, single_region=1, n_threads=8)
print("synth", round(time.time() - t, 1), "s", flush=True)
for flags in (abi.LCR_FLAG_SKIP_PHASING, 0):
    p = host.params_preset("hifi-masseq", seed=20251017, flags=flags)
    regions, _ = host.find_regions(syn.reads, p)
    eng = host.Engine(p, device=0)
    eng.set_references(syn.reference.for_reads(syn.reads))
    batch = host.BatchView(syn.reads, regions)
    h = eng.upload(batch)
    for it in range(2):
        t = time.time()
        eng.run_device(h)
        tt = eng.timing(h)
        print("flags", flags, "run", it, round(time.time() - t, 2), "s", {k: tt[k] for k in ("ms_total", "ms_pileup", "ms_pileup_kernel", "ms_fragments", "ms_phase", "kernel_launches")}, flush=True)
    r = eng.fetch(h)
    print("regions", len(regions), "cand", r.n_cand, r.stats, flush=True)
    eng.release(h)
    eng.close()
