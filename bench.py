#!/usr/bin/env python
"""bench.py — aligned bases piled-up + het-SNP fragment alleles phased per second (BASELINE.json metric).

One step = one pass of the whole hot path (read filter, CIGAR walk, tile pileup + site genotyping, fragment
matrix, phasing, read/SNP assignment, phase sets) over one synthetic batch.

  value   device-resident: inputs already in HBM, lcr_run_device timed with CUDA events on the
          library's own stream (lcr_get_timing), L2 flushed between steps; at N > 1 every step also
          gathers the candidate records and the per-read (HP, PS) arrays to rank 0 over NCCL
  e2e     the same pass through lcr_submit with pinned HOST buffers: H2D of the decoded reads,
          the run, D2H of candidates / HP / PS (and the NCCL gather at N > 1), wall clock around the call
  --impl reference   the reference's CPU algorithm (the line-faithful C++ port in oracle/, mode 1;
          the Rust crate cannot be built here) on all host cores, same workload and metric

Default workload: cfg3 (BASELINE configs[2], the largest 30x single-GPU configuration).  Multi-GPU (torchrun):
contigs are dealt to ranks by longcallr_b200.shard.plan_shards (longest-processing-time on their aligned-base
estimate; weak scaling: the job holds `contigs_per_rank` x N contigs), rank 0 broadcasts the packed reference once.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: synthetic 1 Mb contig, 30x ONT-cDNA, ~1k het SNPs
    "cfg2": dict(preset="ont-cdna", contigs_per_rank=1, synth=dict(contig_len=1_000_000, platform=1, depth=30.0, n_het=1000, n_edit=200, max_intron=300, max_gap=600, both_strands=1)),
    # configs[2]: synthetic chr20 (64 Mb), 30x HiFi MAS-Seq, ~50k candidate sites
    "cfg3": dict(preset="hifi-masseq", contigs_per_rank=1, synth=dict(contig_len=64_000_000, platform=0, depth=30.0, n_het=40000, n_edit=10000, max_intron=300, max_gap=600, both_strands=0)),
    # configs[3] scaled to what 8 ranks of one box can generate and hold on the host: 3 contigs x 10 Mb per rank at 20x ONT-dRNA, 1 het / 1.3 kb
    # (the full configuration is 300 x 10 Mb, i.e. 37.5 contigs per GPU; the per-contig shape is the same)
    "cfg4": dict(preset="ont-drna", contigs_per_rank=3, synth=dict(contig_len=10_000_000, platform=1, depth=20.0, n_het=7700, n_edit=1500, max_intron=300, max_gap=600, both_strands=0)),
    # configs[4]: phasing stress, 500 kb gene-dense block at 500x, 5k het SNPs
    "cfg5": dict(preset="hifi-masseq", contigs_per_rank=1, synth=dict(contig_len=500_000, platform=0, depth=500.0, n_het=5000, n_edit=0, max_intron=500, max_gap=600, both_strands=0, single_region=1)),
    # small case for quick checks
    "tiny": dict(preset="hifi-masseq", contigs_per_rank=1, synth=dict(contig_len=100_000, platform=0, depth=30.0, n_het=100, n_edit=20, max_intron=300, max_gap=600, both_strands=0)),
}
METRIC = "aligned bases piled-up + het-SNP fragments phased /sec (synthetic 30x)"
UNIT = "bases+alleles/s"
SEED = 20251017


class ClockSampler:
    """SM clocks and throttle reasons of the job's GPUs through NVML, sampled by rank 0 during the timed region."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_indices):
        self.samples, self.stop = [], threading.Event()
        self.thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.handles = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in gpu_indices]
        except Exception:
            self.nv, self.handles = None, []

    def _run(self):
        nv = self.nv
        while not self.stop.is_set():
            for h in self.handles:
                try:
                    sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                    mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                    try:
                        rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    self.samples.append((sm, mx, int(rs)))
                except Exception:
                    pass
            self.stop.wait(0.01)

    def start(self):
        if self.handles:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def finish(self):
        self.stop.set()
        if self.thread:
            self.thread.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(s[0] for s in self.samples)
        bits = 0
        for s in self.samples:
            bits |= s[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": [n for b, n in self.REASONS.items() if bits & b], "samples": len(self.samples), "source": "nvml"}


def deal_contigs(name, world):
    """Contigs of the job and the ranks they go to: LPT on the aligned-base estimate (contig length x depth)."""
    from longcallr_b200 import shard

    w = WORKLOADS[name]
    n = w["contigs_per_rank"] * world
    weights = [int(w["synth"]["contig_len"] * w["synth"]["depth"])] * n
    return shard.plan_shards(weights, world)


def make_workload(name, rank, world=1):
    from longcallr_b200 import host

    w = WORKLOADS[name]
    mine = deal_contigs(name, world)[rank]
    # every contig of the job has its own seed; a rank generates only the contigs it was dealt
    syn = host.Synthetic(seed=SEED + 7919 * int(mine[0]), n_contigs=len(mine), **w["synth"])
    p = host.params_preset(w["preset"], seed=SEED)
    regions, _ = host.find_regions(syn.reads, p)
    return w, syn, p, regions


def bounded_sample(ob, host, p, syn, regions, refs, cores, budget_s):
    """How many leading regions the CPU port can process in about budget_s seconds (probe on 16 regions)."""
    probe_n = max(1, min(len(regions), 16))
    t0 = time.perf_counter()
    ob.run(p, host.BatchView(syn.reads, regions[:probe_n]), refs, mode=1, threads=cores)
    dt = max(time.perf_counter() - t0, 1e-9)
    est_total_s = dt * len(regions) / probe_n
    if est_total_s <= budget_s:
        return len(regions)
    return max(probe_n, int(len(regions) * budget_s / est_total_s))


def run_reference(args, rank, world):
    """CPU arm: the reference algorithm (oracle port, reference-order f64 mode) on all host cores."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from longcallr_b200 import host

    w, syn, p, regions = make_workload(args.workload, 0, 1)
    refs = syn.reference.for_reads(syn.reads)
    cores = os.cpu_count() or 1
    single = len(regions) < 4  # one deep region (cfg5) cannot be sub-sampled by regions: it is timed whole, once
    n_sample = len(regions) if single else bounded_sample(ob, host, p, syn, regions, refs, cores, budget_s=10.0)
    sb = host.BatchView(syn.reads, regions[:n_sample])
    steps, warm = (1, 0) if single else (args.steps, args.warmup)
    for _ in range(warm):
        ob.run(p, sb, refs, mode=1, threads=cores)
    t0 = time.perf_counter()
    units = 0
    for _ in range(steps):
        r = ob.run(p, sb, refs, mode=1, threads=cores)
        units += r.stats["n_aligned_bases"] + r.stats["nnz_phase"]
    dt = time.perf_counter() - t0
    value = units / dt
    sample = f"{n_sample} of {len(regions)} regions of {args.workload} per step ({units // max(steps, 1)} units), C++ port of the reference loops (oracle mode 1, -O3 -march=native), {cores} threads over regions"
    emit_json({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": dt / max(steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64/f64",
        "data": "synthetic", "config": {"workload": args.workload, **WORKLOADS[args.workload]["synth"], "preset": w["preset"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_REAL_STDOUT = None


def guard_stdout():
    """stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) go to stderr instead."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit_json(obj):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(obj), flush=True)
    if _REAL_STDOUT is not None:
        os.dup2(2, 1)  # anything printed during teardown goes to stderr again


class _DevBuf:
    """Device memory owned by the library, exposed to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs (ncu launch lists): skip the end-to-end loop; the e2e keys are then null")
    ap.add_argument("--e2e-input", default="bam4", choices=["bam4", "bam4-ondemand", "ascii"],
                    help="host buffers of the end-to-end path: bam4 = 4-bit bases as the BAM record stores them + byte qualities, bam4-ondemand = qualities fetched from pinned memory by the kernels, ascii = decoded bases")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from longcallr_b200 import abi, host

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one rank per GPU: run (and first-touch the pinned staging buffers) on the CPUs next to that GPU, as `numactl` per rank would
    numa = "unbound"
    all_cpus = os.sched_getaffinity(0)
    if not os.environ.get("LCR_NO_CPU_AFFINITY"):
        try:
            import pynvml

            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            numa = f"{len(os.sched_getaffinity(0))} CPUs next to GPU {local_rank} (nvmlDeviceSetCpuAffinity)"
        except Exception as e:  # not fatal: the run is only slower
            numa = f"unbound ({type(e).__name__})"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    w, syn, p, regions = make_workload(args.workload, rank, world)
    refs = syn.reference.for_reads(syn.reads)

    # reference slices: rank 0 packs every rank's contigs and broadcasts them once over NCCL (north_star: "a single NCCL
    # broadcast of the reference slice"); each rank keeps its own.  Contig sets are equal-sized by construction (weak scaling).
    ref_np = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.uint8) for r in refs]))
    bcast_ms = 0.0
    if world > 1:
        L = int(ref_np.size)
        mine = torch.from_numpy(ref_np).to(dev)
        parts = [torch.empty(L, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, parts, dst=0)
        packed = torch.cat(parts) if rank == 0 else torch.empty(world * L, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(packed, src=0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        got = packed[rank * L:(rank + 1) * L].cpu().numpy()
        assert np.array_equal(got, ref_np), "reference broadcast mismatch"
        ref_np = got
        del packed, parts, mine

    eng = host.Engine(p, device=local_rank)
    off = 0
    for tid, r in enumerate(refs):
        eng.set_reference(tid, ref_np[off:off + len(r)])
        off += len(r)

    # pinned staging buffers for the end-to-end path
    pinned_keep = []

    def alloc_pinned(nbytes):
        t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        pinned_keep.append(t)
        return t.numpy()

    pinned_reads = host.ArrayReadSet.like(syn.reads, alloc_pinned)
    batch = host.BatchView(pinned_reads, regions)
    # end-to-end inputs in the BAM record's own form: 4-bit bases (expanded on the device) next to the byte qualities; bam4-ondemand
    # leaves the qualities in pinned host memory for the kernels to fetch (LCR_FLAG_QUAL_ON_DEMAND: fewer bytes, but 32-byte bus reads
    # are slower than the bulk copy on this box - see DESIGN.md); ascii sends decoded bytes
    eng_e2e, batch_e2e = eng, batch
    if args.e2e_input in ("bam4", "bam4-ondemand"):
        batch_e2e = host.BatchView(pinned_reads, regions, seq4=host.pack_seq4(pinned_reads, alloc_pinned, threads=16))
    if args.e2e_input == "bam4-ondemand":
        p_e2e = host.params_preset(w["preset"], seed=SEED, flags=abi.LCR_FLAG_QUAL_ON_DEMAND)
        eng_e2e = host.Engine(p_e2e, device=local_rank)
        off = 0
        for tid, r in enumerate(refs):
            eng_e2e.set_reference(tid, ref_np[off:off + len(r)])
            off += len(r)

    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def flush_l2():
        flush.add_(1)
        torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from longcallr_b200 import shard

    gbuf = {}
    pg = shard.PackedGather(dist, rank, world, dev) if world > 1 else None

    def gather_results(cand_u8, hp_u8, ps_u8, setup=False):
        """The per-rank VCF records and (read, HP, PS) arrays go to rank 0 (device tensors in): one NCCL gather (shard.PackedGather)."""
        out = pg.gather(cand_u8, hp_u8, ps_u8, setup=setup)
        gbuf["cap"] = pg.cap
        return out

    def device_results(handle):
        v = eng.device_view(handle)
        cand_u8 = torch.as_tensor(_DevBuf(v.cand, v.n_cand * abi.CANDIDATE_DTYPE.itemsize), device=dev) if v.n_cand else torch.zeros(0, dtype=torch.uint8, device=dev)
        hp_u8 = torch.as_tensor(_DevBuf(v.hp, v.n_reads), device=dev) if v.n_reads else torch.zeros(0, dtype=torch.uint8, device=dev)
        ps_u8 = torch.as_tensor(_DevBuf(v.ps, 4 * v.n_reads), device=dev) if v.n_reads else torch.zeros(0, dtype=torch.uint8, device=dev)
        return cand_u8, hp_u8, ps_u8

    # ---- device-resident timing ----
    handle = eng.upload(batch)
    for _ in range(args.warmup):
        eng.run_device(handle)
    if world > 1:  # before the timed region: peer connections of both collectives, and the payload capacity from the real sizes
        for _ in range(2):
            gather_results(*device_results(handle), setup=True)
        torch.cuda.synchronize()
    sampler = ClockSampler(list(range(world)) if rank == 0 else [])
    barrier()
    sampler.start()
    acc = dict(ms_total=0.0, ms_pileup=0.0, ms_pileup_kernel=0.0, ms_fragments=0.0, ms_phase=0.0, ms_prep=0.0, ms_enum=0.0, ms_phase_kernel=0.0)
    pile_bytes = phase_bytes = launches = attempts = 0
    gather_ms = 0.0
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gathered_cands = 0
    for _ in range(args.steps):
        flush_l2()
        eng.run_device(handle)
        t = eng.timing(handle)
        for k in acc:
            acc[k] += t[k]
        pile_bytes += t["pileup_alg_bytes"]
        phase_bytes += t["phase_alg_bytes"]
        launches += t["kernel_launches"]
        attempts += t["run_attempts"]
        if world > 1:
            g0.record()
            outl = gather_results(*device_results(handle))
            g1.record()
            torch.cuda.synchronize()
            gather_ms += g0.elapsed_time(g1)
            if rank == 0:
                gathered_cands = sum(int(o[:24].view(torch.int64)[0].item()) for o in outl) // abi.CANDIDATE_DTYPE.itemsize
    barrier()
    clocks = sampler.finish()
    dev_ms = acc["ms_total"] + gather_ms
    res = eng.fetch(handle)
    eng.release(handle)
    units = res.stats["n_aligned_bases"] + res.stats["nnz_phase"]

    # ---- end to end through lcr_submit with host buffers ----
    for _ in range(0 if args.no_e2e else 2):
        r = eng_e2e.submit_raw(batch_e2e)
        eng_e2e.free_result(r)
    barrier()
    e2e_t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(0 if args.no_e2e else args.steps):
        flush_l2()
        raw = eng_e2e.submit_raw(batch_e2e)  # the reference-facing call: host buffers in, host results (lcr_result) out
        n_cand_last = raw.contents.n_cand  # the result is in host memory: a caller reads it in place
        if world > 1:
            # the call left its results in host memory; they travel to rank 0 over NCCL (staged through the device), and rank 0
            # brings the valid part of every rank's payload back to pinned host memory, where it would write the VCF / tag the BAM
            r0 = raw.contents
            cand_u8 = torch.from_numpy(abi.as_array(r0.cand, "u1", r0.n_cand * abi.CANDIDATE_DTYPE.itemsize)).to(dev, non_blocking=True)
            hp_u8 = torch.from_numpy(abi.as_array(r0.hp, "u1", r0.n_reads)).to(dev, non_blocking=True)
            ps_u8 = torch.from_numpy(abi.as_array(r0.ps, "u1", 4 * r0.n_reads)).to(dev, non_blocking=True)
            outl = gather_results(cand_u8, hp_u8, ps_u8)
            if rank == 0:
                torch.cuda.synchronize()
                if "host" not in gbuf:
                    gbuf["host"] = [torch.empty(gbuf["cap"], dtype=torch.uint8).pin_memory() for _ in range(world)]
                for o, hbuf in zip(outl, gbuf["host"]):
                    nb = 24 + int(o[:24].view(torch.int64).sum().item())
                    hbuf[:nb].copy_(o[:nb], non_blocking=True)
            torch.cuda.synchronize()
        eng_e2e.free_result(raw)
        tt = eng_e2e.last_submit_timing()
        h2d, d2h = tt["h2d_bytes"], tt["d2h_bytes"]
    barrier()
    e2e_ms = (time.perf_counter() - e2e_t0) * 1e3

    # max over ranks
    tvals = torch.tensor([dev_ms, e2e_ms, gather_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tvals, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, gather_ms_max = float(tvals[0]), float(tvals[1]), float(tvals[2])
    uvals = torch.tensor([units, res.n_cand], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(uvals, op=dist.ReduceOp.SUM)
    total_units, total_cands = float(uvals[0]), int(uvals[1])
    if world > 1 and rank == 0:
        assert gathered_cands == total_cands, "the gather lost candidate records"

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        steps = max(args.steps, 1)

        def gbs(nbytes, ms):
            return (nbytes / 1e9) / (ms / 1e3) if ms > 0 else 0.0

        # the kernels / kernel groups of one step with their measured time (CUDA events on the library stream) and algorithmic bytes
        kernels = {
            "k_pileup_tile": dict(ms=acc["ms_pileup_kernel"] / steps, bytes=pile_bytes / steps, what="tile pileup + count filters: 1 B per aligned base (the kernel reads no qualities) + 16 B per segment and item + 48 B and the reference bytes per tile + 72 B per surviving site"),
            "k_read_span+k_read_walk": dict(ms=acc["ms_prep"] / steps, bytes=None, what="read filter, reference spans, CIGAR walk into items / segments (latency-bound walks)"),
            "k_enum_search": dict(ms=acc["ms_enum"] / steps, bytes=None, what="2^n enumeration of regions with <= 10 sites (shared-memory resident)"),
            "k_phase": dict(ms=acc["ms_phase_kernel"] / steps, bytes=phase_bytes / steps, what="LD path, read / SNP assignment, rescue, phase sets (L2 resident): B_sweep x iterations"),
            "fragments+ld": dict(ms=acc["ms_fragments"] / steps, bytes=None, what="fragment matrix (CSR + CSC), LD pair tables and graph"),
        }
        dominant = max(kernels, key=lambda k: kernels[k]["ms"])
        # the roofline object describes the dominant kernel when its bytes are defined, else the HBM-streaming kernel of the path
        rk = dominant if kernels[dominant]["bytes"] else "k_pileup_tile"
        achieved = gbs(kernels[rk]["bytes"], kernels[rk]["ms"])
        traffic = None
        try:  # DRAM bytes of one launch of that kernel from the committed ncu --set full capture of this workload (profiles/)
            t = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"))).get(args.workload, {}).get(rk)
            if t:
                traffic = int(t["dram_read_bytes"] + t["dram_write_bytes"])
        except Exception:
            pass
        stage_bytes = 2 * res.stats["n_aligned_bases"] + 4 * int(syn.reads.cig_off[-1]) + 32 * int(syn.reads.n_reads) + int(res.stats["n_positions"])
        out = {
            "metric": METRIC, "value": total_units * args.steps / (dev_ms_max / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64 fixed-point (f64 for QUAL)",
            "data": "synthetic",
            "config": {"workload": args.workload, "preset": w["preset"], **w["synth"], "contigs_per_gpu": int(w["contigs_per_rank"]), "regions_per_gpu": int(len(regions)),
                       "reads_per_gpu": int(syn.reads.n_reads), "aligned_bases_per_gpu": int(res.stats["n_aligned_bases"]), "phase_alleles_per_gpu": int(res.stats["nnz_phase"]),
                       "candidates": int(total_cands), "cross_optimize_calls": int(res.stats["n_cross_optimize"]), "sweep_iters": int(res.stats["n_sweep_iters"]),
                       "l2": "flushed between steps (512 MiB write)",
                       "sharding": "contigs dealt by LPT (shard.plan_shards); per step one NCCL gather of candidate records + per-read HP/PS to rank 0 inside the timed region" if world > 1 else "single GPU",
                       "reference_broadcast_ms": bcast_ms, "gather_ms_per_step": gather_ms_max / args.steps, "run_attempts_per_step": attempts / steps},
            "e2e": {"value": None, "unit": UNIT, "skipped": "--no-e2e"} if args.no_e2e else {"value": total_units * args.steps / (e2e_ms_max / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms_max / args.steps, "cpu_affinity": numa,
                    "input": {"bam4": "4-bit bases as the BAM record stores them, byte qualities, per-read tables: all copied from pinned memory",
                              "bam4-ondemand": "4-bit bases + per-read tables copied; qualities stay in pinned host memory and the kernels fetch the 32-byte sectors they need (counted in h2d_bytes_per_step)",
                              "ascii": "decoded ASCII bases + qualities copied"}[args.e2e_input]},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": rk, "dominant_kernel": dominant, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "alg_bytes_per_launch": int(kernels[rk]["bytes"]), "ms_per_launch": kernels[rk]["ms"], "what": kernels[rk]["what"]},
            "roofline_stage": {"stage": "pileup + genotype (P0-P7): span pass, walk, tile kernel, site likelihood, candidate compaction", "ms": acc["ms_pileup"] / steps,
                               "alg_bytes": int(stage_bytes), "achieved": gbs(stage_bytes, acc["ms_pileup"] / steps), "frac": gbs(stage_bytes, acc["ms_pileup"] / steps) / peak,
                               "formula": "2 N_al + 4 N_cigar + 32 N_reads + N_pos (SURVEY 8d without the per-position record this build never writes)"},
            "kernel_ms_per_step": {k: v["ms"] for k, v in kernels.items()},
            "stage_ms_per_step": {"pileup": acc["ms_pileup"] / steps, "fragments": acc["ms_fragments"] / steps, "phase": acc["ms_phase"] / steps, "total": acc["ms_total"] / steps,
                                  "gather": gather_ms / steps},
        }
        if not args.no_cpu_baseline and len(regions) < 4:
            # a single deep region cannot be sub-sampled by regions and takes the CPU port minutes (cfg5): timed by --impl reference
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"skipped: {args.workload} is {len(regions)} region(s); bench.py --impl reference --workload {args.workload} times it whole (profiles/)"}
        elif not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_binding as ob

            os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every host core back
            cores = os.cpu_count() or 1
            n_sample = bounded_sample(ob, host, p, syn, regions, refs, cores, budget_s=15.0)
            cb = host.BatchView(syn.reads, regions[:n_sample])
            t0 = time.perf_counter()
            r1 = ob.run(p, cb, refs, mode=1, threads=cores)
            dt = time.perf_counter() - t0
            cu = r1.stats["n_aligned_bases"] + r1.stats["nnz_phase"]
            out["cpu_baseline"] = {"value": cu / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"first {n_sample} of {len(regions)} regions of {args.workload} once ({cu} units, {dt:.2f} s), C++ port of the reference loops (oracle mode 1, -O3 -march=native), {cores} threads over regions"}
        emit_json(out)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
