"""N > 1 path on CPU: world_size-2 gloo processes shard regions, exchange the reference and gather candidates.

The per-rank compute is the CUDA path on a GPU box; here (no GPU) the oracle stands in for it, which is
enough to check the host-side sharding / broadcast / gather logic against the single-process result.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
import oracle_binding as ob
from longcallr_b200 import host, shard


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    syn = host.Synthetic(seed=21, contig_len=150_000, n_contigs=2, platform=0, depth=25.0, n_het=100, n_edit=10, both_strands=0, max_intron=300, max_gap=600, n_threads=2)
    p = host.params_preset("hifi-masseq", seed=4)
    regions, _ = host.find_regions(syn.reads, p)
    # rank 0 owns the FASTA; everyone gets it through one broadcast
    refs = shard.broadcast_reference(dist, syn.reference.for_reads(syn.reads) if rank == 0 else [], rank)
    for a, b in zip(refs, syn.reference.for_reads(syn.reads)):
        assert np.array_equal(a, b)
    plan = shard.plan_shards(shard.region_weights(syn.reads, regions), world)
    sub_reads, sub_regions, _ = shard.shard_batch(syn.reads, regions, plan[rank])
    res = ob.run(p, host.BatchView(sub_reads, sub_regions), refs, mode=0)
    allc = shard.gather_candidates(dist, res.cand, plan[rank], rank, world)
    if rank == 0:
        full = ob.run(p, host.BatchView(syn.reads, regions), refs, mode=0)
        ok = len(allc) == full.n_cand
        for f in helpers.INT_FIELDS + helpers.FP_FIELDS:
            ok = ok and np.array_equal(allc[f], full.cand[f], equal_nan=True)
        covered = sorted(int(i) for pl in plan for i in pl)
        ok = ok and covered == list(range(len(regions)))
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(300)
        assert pr.exitcode == 0
    assert q.get(timeout=10) is True


def _gather_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pg = shard.PackedGather(dist, rank, world)
    rng = np.random.default_rng(100 + rank)
    ok = True
    for step in range(3):  # sizes differ per rank; steps after the first reuse the capacity agreed at setup
        n_c, n_r = 700 + 300 * rank - 50 * step, 5000 + 1000 * rank
        cand = torch.from_numpy(rng.integers(0, 256, size=n_c * 88, dtype=np.uint8))
        hp = torch.from_numpy(rng.integers(0, 3, size=n_r, dtype=np.uint8))
        ps = torch.from_numpy(rng.integers(0, 256, size=4 * n_r, dtype=np.uint8))
        out = pg.gather(cand, hp, ps, setup=(step == 0))
        mine = [t.clone() for t in (cand, hp, ps)]
        allm = [None] * world
        dist.all_gather_object(allm, [m.numpy().tobytes() for m in mine])
        if rank == 0:
            for k in range(world):
                got = [x.numpy().tobytes() for x in shard.PackedGather.unpack(out[k])]
                ok = ok and got == allm[k]
    try:  # a payload above the agreed capacity is refused on the rank that has it
        big = torch.zeros(pg.cap, dtype=torch.uint8)
        if rank == 1:
            pg.gather(big, big[:1], big[:4])
            ok = False
    except RuntimeError:
        pass
    if rank == 0:
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_packed_gather_two_ranks():
    """The one-collective-per-step gather bench.py issues over NCCL, here over gloo: in-band sizes, fixed capacity, exact bytes."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(300)
        assert pr.exitcode == 0
    assert q.get(timeout=10) is True


def test_lpt_plan_is_balanced_and_complete():
    w = np.array([100, 1, 1, 1, 50, 50, 30, 20, 5, 5])
    plan = shard.plan_shards(w, 4)
    assert sorted(int(i) for p in plan for i in p) == list(range(len(w)))
    loads = [int(w[p].sum()) for p in plan]
    assert max(loads) == 100 and max(loads) - min(loads) <= 100 - 50
