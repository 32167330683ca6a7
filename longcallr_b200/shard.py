"""Region sharding across the GPUs of one box.

Regions are independent by construction (zero-coverage gaps, src/util.rs:287-330; the worker shares
nothing mutable, src/thread.rs:77-222), so they are dealt to ranks with no collective on the data
path.  The only exchanges are the two north_star names: one broadcast of the packed reference
slices from rank 0 and one gather of the per-region candidate records to rank 0.
"""
import numpy as np

from . import abi


def region_weights(reads, regions):
    """Work estimate of each region: bases of the reads in its range (what the pileup streams)."""
    so = np.asarray(reads.seq_off, dtype=np.int64)
    return (so[regions["read_end"]] - so[regions["read_begin"]]).astype(np.int64)


def plan_shards(weights, world_size):
    """Longest-processing-time bin packing.  Returns a list of ascending region-index arrays, one per rank."""
    weights = np.asarray(weights, dtype=np.int64)
    order = np.argsort(-weights, kind="stable")
    loads = np.zeros(world_size, dtype=np.int64)
    bins = [[] for _ in range(world_size)]
    for r in order:
        k = int(np.argmin(loads))
        bins[k].append(int(r))
        loads[k] += int(weights[r]) + 1
    return [np.array(sorted(b), dtype=np.int64) for b in bins]


def shard_batch(reads, regions, idx):
    """The sub-batch of one rank: its regions with read ranges re-based onto a compact copy of their reads."""
    from .host import ArrayReadSet

    sub = regions[idx].copy()
    keep = np.zeros(reads.n_reads, dtype=bool)
    for r in sub:
        keep[r["read_begin"]:r["read_end"]] = True
    new_index = np.cumsum(keep) - 1
    for r in sub:
        n = int(r["read_end"] - r["read_begin"])
        b = int(new_index[r["read_begin"]]) if n else 0
        r["read_begin"], r["read_end"] = b, b + n
    sel = np.nonzero(keep)[0]
    so, co = np.asarray(reads.seq_off, dtype=np.int64), np.asarray(reads.cig_off, dtype=np.int64)
    seq_len, cig_len = so[sel + 1] - so[sel], co[sel + 1] - co[sel]
    seq_off = np.concatenate([[0], np.cumsum(seq_len)]).astype("<u8")
    cig_off = np.concatenate([[0], np.cumsum(cig_len)]).astype("<u8")
    seq = np.empty(int(seq_off[-1]), dtype=np.uint8)
    qual = np.empty(int(seq_off[-1]), dtype=np.uint8)
    cigar = np.empty(int(cig_off[-1]), dtype="<u4")
    for j, i in enumerate(sel):
        seq[seq_off[j]:seq_off[j + 1]] = reads.seq[so[i]:so[i + 1]]
        qual[seq_off[j]:seq_off[j + 1]] = reads.qual[so[i]:so[i + 1]]
        cigar[cig_off[j]:cig_off[j + 1]] = reads.cigar[co[i]:co[i + 1]]
    out = ArrayReadSet(reads.contig_names, reads.contig_lens, tid=reads.tid[sel], pos=reads.pos[sel], flag=reads.flag[sel], mapq=reads.mapq[sel],
                       ts=reads.ts[sel], de=reads.de[sel], seq_off=seq_off, cig_off=cig_off, seq=seq, qual=qual, cigar=cigar)
    return out, sub, sel


def broadcast_reference(dist, seqs, rank, device=None):
    """Rank 0 holds the FASTA; every rank receives the packed contigs (one broadcast of lengths, one of bytes)."""
    import torch

    n = torch.tensor([len(seqs) if rank == 0 else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src=0)
    lens = torch.tensor([len(s) if s is not None else 0 for s in seqs] if rank == 0 else [0] * int(n.item()), dtype=torch.int64, device=device)
    dist.broadcast(lens, src=0)
    total = int(lens.sum().item())
    if rank == 0:
        packed = torch.from_numpy(np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs if s is not None and len(s)]) if total else np.zeros(0, np.uint8))
        packed = packed.to(device) if device is not None else packed
    else:
        packed = torch.empty(total, dtype=torch.uint8, device=device)
    if total:
        dist.broadcast(packed, src=0)
    out, off = [], 0
    host_packed = packed.cpu().numpy()
    for ln in lens.tolist():
        out.append(host_packed[off:off + ln] if ln else None)
        off += ln
    return out


def gather_candidates(dist, cand, region_ids, rank, world_size, device=None):
    """Variable-length gather of the fixed 88-byte candidate records (and the global region id of each) to rank 0."""
    import torch

    raw = np.frombuffer(np.ascontiguousarray(cand).tobytes(), dtype=np.uint8)
    gids = np.asarray(region_ids, dtype=np.int64)[cand["region"]] if len(cand) else np.zeros(0, np.int64)
    cnt = torch.tensor([len(cand)], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world_size)]
    dist.all_gather(cnts, cnt)
    mx = max(int(c.item()) for c in cnts)
    rec = torch.zeros(max(mx, 1) * abi.CANDIDATE_DTYPE.itemsize, dtype=torch.uint8, device=device)
    gid = torch.zeros(max(mx, 1), dtype=torch.int64, device=device)
    if len(cand):
        rec[: raw.size] = torch.from_numpy(raw.copy()).to(rec.device)
        gid[: len(cand)] = torch.from_numpy(gids.copy()).to(gid.device)
    recs = [torch.empty_like(rec) for _ in range(world_size)] if rank == 0 else None
    gidl = [torch.empty_like(gid) for _ in range(world_size)] if rank == 0 else None
    dist.gather(rec, recs, dst=0)
    dist.gather(gid, gidl, dst=0)
    if rank != 0:
        return None
    parts = []
    for k in range(world_size):
        n = int(cnts[k].item())
        c = np.frombuffer(recs[k].cpu().numpy().tobytes(), dtype=abi.CANDIDATE_DTYPE)[:n].copy()
        c["region"] = gidl[k].cpu().numpy()[:n]
        parts.append(c)
    allc = np.concatenate(parts) if parts else np.zeros(0, abi.CANDIDATE_DTYPE)
    return allc[np.lexsort((allc["pos"], allc["region"]))]


class PackedGather:
    """One collective per step for the per-rank results: the candidate records and the per-read (HP, PS) arrays of every rank go to
    rank 0 as ONE gather of a fixed-capacity byte payload whose first 24 bytes carry the three sizes ("gather of per-region
    VCF records" and of the (read, HP, PS) tuples the phased BAM needs).  The capacity (1.25 x the largest rank) is agreed by one
    exchange of sizes (`setup=True`, outside a timed region); a later payload that outgrew it is an error, never a silent truncation.
    Works on device tensors over NCCL (bench.py) and on CPU tensors over gloo (tests)."""

    HEADER = 24

    def __init__(self, dist, rank, world, device=None):
        self.dist, self.rank, self.world, self.device = dist, rank, world, device
        self.cap = 0
        self.send = self.recv = self.hdr = None

    def gather(self, cand_u8, hp_u8, ps_u8, setup=False):
        import torch

        sz = [int(cand_u8.numel()), int(hp_u8.numel()), int(ps_u8.numel())]
        need = self.HEADER + sum(sz)
        if setup:
            n = torch.tensor([need], device=self.device, dtype=torch.int64)
            allv = [torch.zeros_like(n) for _ in range(self.world)]
            self.dist.all_gather(allv, n)
            mx = max(int(v.item()) for v in allv)
            cap = (mx + mx // 4 + 4095) // 4096 * 4096
            if self.cap < cap:
                self.cap = cap
                self.send = torch.zeros(cap, dtype=torch.uint8, device=self.device)
                self.recv = [torch.empty(cap, dtype=torch.uint8, device=self.device) for _ in range(self.world)] if self.rank == 0 else None
                self.hdr = torch.zeros(3, dtype=torch.int64)
                if self.device is not None and torch.device(self.device).type == "cuda":
                    self.hdr = self.hdr.pin_memory()
        if need > self.cap:
            raise RuntimeError(f"rank {self.rank}: gather payload {need} B exceeds the capacity {self.cap} B agreed at setup")
        self.hdr[:] = torch.tensor(sz, dtype=torch.int64)
        self.send[: self.HEADER].view(torch.int64).copy_(self.hdr, non_blocking=True)
        a, b = self.HEADER + sz[0], self.HEADER + sz[0] + sz[1]
        self.send[self.HEADER:a].copy_(cand_u8)
        self.send[a:b].copy_(hp_u8)
        self.send[b:b + sz[2]].copy_(ps_u8)
        self.dist.gather(self.send, self.recv, dst=0)
        return self.recv

    @classmethod
    def unpack(cls, payload):
        """(candidate bytes, hp bytes, ps bytes) views of one rank's payload on rank 0."""
        import torch

        sz = [int(x) for x in payload[: cls.HEADER].view(torch.int64).tolist()]
        a, b = cls.HEADER + sz[0], cls.HEADER + sz[0] + sz[1]
        return payload[cls.HEADER:a], payload[a:b], payload[b:b + sz[2]]
