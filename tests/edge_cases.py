"""Edge-case batches shared by the oracle tests (CPU) and the GPU parity tests."""
import numpy as np

import helpers
from longcallr_b200 import abi, host

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def _ref(L, seed=0):
    return np.random.default_rng(seed).choice(BASES, size=L)


def _reads_over(ref, start, n, length, alts=None, qual=30, flag=0, ts="+", cigar=None, **kw):
    """n identical-span reads; odd reads carry the alt alleles in `alts` {pos: base}."""
    recs = []
    for k in range(n):
        seq = bytearray(ref[start:start + length].tobytes())
        if alts and k % 2:
            for pos, b in alts.items():
                seq[pos - start] = b
        recs.append(dict(pos=start, cigar=cigar or f"{length}M", seq=seq.decode(), qual=qual, flag=flag if np.isscalar(flag) else flag[k], ts=ts, **kw))
    return recs


def alt_of(ref, pos):
    return b"ACGT"[(b"ACGT".index(bytes([ref[pos]])) + 2) % 4]


def cases():
    """name -> (params, reads, refs, regions, expected region_status or None)"""
    out = {}
    L = 4000
    ref = _ref(L, 1)
    p = host.params_preset("hifi-masseq", seed=3, flags=abi.LCR_FLAG_EMIT_PLANES | abi.LCR_FLAG_EMIT_FRAGMENTS)

    # empty batch and a region without reads
    out["empty_batch"] = (p, helpers.make_reads(L, []), [ref], np.zeros(0, dtype=abi.REGION_DTYPE), None)
    out["region_without_reads"] = (p, helpers.make_reads(L, []), [ref], helpers.one_region(101, 901, 0), [0])

    # every read filtered: low mapq, short, secondary, supplementary, unmapped, divergent
    recs = _reads_over(ref, 500, 6, 800, alts={900: alt_of(ref, 900)})
    recs[0]["mapq"] = 5
    recs[1]["flag"] = 0x100
    recs[2]["flag"] = 0x800
    recs[3]["flag"] = 0x4
    recs[4]["de"] = 0.9
    recs[5] = dict(pos=500, cigar="300M", seq=ref[500:800].tobytes().decode(), qual=30)
    out["all_reads_filtered"] = (p, helpers.make_reads(L, recs), [ref], helpers.one_region(501, 1301, 6), [0])

    # soft clips, insertions, deletions, introns, =/X ops, reverse strand, reads hanging over both region edges
    alts = {1000: alt_of(ref, 1000), 1300: alt_of(ref, 1300), 1600: alt_of(ref, 1600)}
    recs = []
    for k in range(14):
        body = bytearray(ref[700:1900].tobytes())
        if k % 2:
            for pos, b in alts.items():
                body[pos - 700] = b
        if k % 3 == 0:   # 10S 400M 3I 200M 5D 100N 490M 7S  (ref span 1195)
            seq = b"A" * 10 + bytes(body[0:400]) + b"GGG" + bytes(body[400:600]) + bytes(body[605 + 100:605 + 100 + 490]) + b"C" * 7
            cig = "10S400M3I200M5D100N490M7S"
        elif k % 3 == 1:  # = and X ops
            seq = bytes(body[0:1100])
            cig = "500=1X599M"
        else:
            seq = bytes(body[0:1200])
            cig = "1200M"
        recs.append(dict(pos=700, cigar=cig, seq=seq.decode(), qual=[20 + (i * 7 + k) % 25 for i in range(len(seq))], flag=16 if k % 4 < 2 else 0, ts="+" if k % 5 else "-"))
    out["mixed_cigars_window_edges"] = (host.params_preset("hifi-isoseq", seed=3, flags=p.flags), helpers.make_reads(L, recs), [ref], helpers.one_region(801, 1801, 14), [0])

    # lower-case and N reference bytes never produce calls (candidate.rs:255-264)
    ref2 = ref.copy()
    ref2[1000] = ord("n"); ref2[1300] = ord("N"); ref2[1600] = ord(chr(ref2[1600]).lower())
    out["masked_reference_bytes"] = (p, helpers.make_reads(L, _reads_over(ref, 700, 12, 1200, alts=alts)), [ref2], helpers.one_region(701, 1901, 12), [0])

    # quality 0 bases: allowed in the pileup (log10(0) flows through), fatal at a fragment site
    recs = _reads_over(ref, 700, 12, 1200, alts=alts)
    recs[2]["qual"] = [0 if i == 50 else 30 for i in range(1200)]
    out["baseq_zero_off_site"] = (p, helpers.make_reads(L, recs), [ref], helpers.one_region(701, 1901, 12), [0])
    recs = _reads_over(ref, 700, 12, 1200, alts=alts)
    recs[3]["qual"] = [0 if i == 300 else 30 for i in range(1200)]
    out["baseq_zero_at_site"] = (p, helpers.make_reads(L, recs), [ref], helpers.one_region(701, 1901, 12), [abi.LCR_ERR_BASEQ_ZERO])

    # unknown CIGAR op (P) -> status instead of the reference's panic (util.rs:943-945)
    recs = _reads_over(ref, 700, 12, 1200, alts=alts)
    recs[5]["cigar"] = "600M2P600M"
    out["bad_cigar"] = (p, helpers.make_reads(L, recs), [ref], helpers.one_region(701, 1901, 12), [abi.LCR_ERR_BAD_CIGAR])

    # region on a contig that has no reference; second region fine
    reads = helpers.make_reads(L, _reads_over(ref, 700, 12, 1200, alts=alts))
    regs = np.zeros(2, dtype=abi.REGION_DTYPE)
    regs[0] = (1, 701, 1901, 0, 12)
    regs[1] = (0, 701, 1901, 0, 12)
    out["missing_reference"] = (p, reads, [ref], regs, [abi.LCR_ERR_NO_REFERENCE, 0])

    # dense cluster of SNPs (candidate.rs:465-526) and a tri-allelic site
    dense = {1000 + 7 * i: alt_of(ref, 1000 + 7 * i) for i in range(8)}
    recs = _reads_over(ref, 700, 16, 1200, alts=dense)
    tri = 1500
    others = [b for b in b"ACGT" if b != ref[tri]]
    for k, r in enumerate(recs):
        s = bytearray(r["seq"].encode())
        s[tri - 700] = others[k % 2]
        r["seq"] = s.decode()
    out["dense_and_triallelic"] = (p, helpers.make_reads(L, recs), [ref], helpers.one_region(701, 1901, 16), [0])

    # ONT end trimming + strand bias filter on, alt on one strand only
    po = host.params_preset("ont-cdna", seed=3, flags=p.flags)
    recs = _reads_over(ref, 700, 24, 1200, alts={720: alt_of(ref, 720), 1300: alt_of(ref, 1300), 1890: alt_of(ref, 1890)}, flag=[16 if k % 2 else 0 for k in range(24)])
    out["ont_trim_and_strand_bias"] = (po, helpers.make_reads(L, recs), [ref], helpers.one_region(701, 1901, 24), [0])

    # poly-A tail near the read end over a non-A reference (util.rs:754-789)
    recs = _reads_over(ref, 700, 12, 1200, alts=alts)
    for r in recs:
        s = bytearray(r["seq"].encode())
        s[1170:1200] = b"A" * 30
        r["seq"] = s.decode()
    out["polya_tail"] = (p, helpers.make_reads(L, recs), [ref], helpers.one_region(701, 1901, 12), [0])

    # hard clip then soft clip: rust-htslib's leading/trailing_softclips look through the H (util.rs:682-690, fragment.rs:59)
    recs = []
    for k in range(14):
        body = bytearray(ref[700:1900].tobytes())
        if k % 2:
            for pos, b in alts.items():
                body[pos - 700] = b
        if k % 3 == 0:
            seq, cig = b"T" * 10 + bytes(body) + b"G" * 7, "5H10S1200M7S3H"
        elif k % 3 == 1:
            seq, cig = b"T" * 4 + bytes(body), "2H4S1200M9H"
        else:
            seq, cig = bytes(body) + b"AAAAAA", "8H1200M6S"
        recs.append(dict(pos=700, cigar=cig, seq=seq.decode(), qual=[12 + (i * 5 + k) % 30 for i in range(len(seq))], flag=16 if k % 4 < 2 else 0))
    out["hard_then_soft_clips"] = (p, helpers.make_reads(L, recs), [ref], helpers.one_region(701, 1901, 14), [0])
    po2 = host.params_preset("ont-drna", seed=3, flags=p.flags)
    out["hard_then_soft_clips_ont"] = (po2, helpers.make_reads(L, recs), [ref], helpers.one_region(701, 1901, 14), [0])

    # quality 0 on an RNA-edit candidate (A>G, not a phase site): the reference does not panic there; its IEEE outcomes
    # (NaN / +inf phase scores in the rescue pass, snpfrags.rs:191-281, NaN read scores afterwards, :580-618) are part of the contract
    a_sites = [int(x) for x in np.nonzero(ref[1100:1800] == ord("A"))[0][:3] + 1100]
    hets = {1000: alt_of(ref, 1000), 1300: alt_of(ref, 1300), 1850: alt_of(ref, 1850)}
    for name, zeros in (("baseq_zero_edit_site_one", [(1, 0)]), ("baseq_zero_edit_site_both", [(1, 0), (4, 0)]),
                        ("baseq_zero_edit_site_discordant", [(1, 0), (6, 1)]), ("baseq_zero_edit_site_unused_read", [(3, 0)])):
        recs = _reads_over(ref, 700, 16, 1200, alts=hets)
        for k, r in enumerate(recs):
            sq = bytearray(r["seq"].encode())
            for j, apos in enumerate(a_sites):
                if k % 2:
                    sq[apos - 700] = ord("G")
            r["qual"] = [30] * 1200
            r["seq"] = sq.decode()
        for k, flip in zeros:
            sq = bytearray(recs[k]["seq"].encode())
            if flip:  # this read carries the other haplotype's allele at the edit site
                sq[a_sites[0] - 700] = ord("G") if sq[a_sites[0] - 700] == ord("A") else ord("A")
            recs[k]["seq"] = sq.decode()
            recs[k]["qual"][a_sites[0] - 700] = 0
        out[name] = (p, helpers.make_reads(L, recs), [ref], helpers.one_region(701, 1901, 16), [0])

    # two regions whose read ranges overlap: reads that fall into both windows keep the HP / PS entry of the lower region
    recs = _reads_over(ref, 700, 12, 1200, alts=alts) + _reads_over(ref, 1250, 12, 1400, alts={1300: alt_of(ref, 1300), 1600: alt_of(ref, 1600), 2000: alt_of(ref, 2000), 2400: alt_of(ref, 2400)})
    regs = np.zeros(2, dtype=abi.REGION_DTYPE)
    regs[0] = (0, 701, 1501, 0, 24)
    regs[1] = (0, 1501, 2651, 0, 24)
    out["shared_reads_two_regions"] = (p, helpers.make_reads(L, recs), [ref], regs, [0, 0])

    # maximum depth: sites above max_depth are skipped (candidate.rs:90-94)
    pm = host.params_preset("hifi-masseq", seed=3, max_depth=20, flags=p.flags)
    out["above_max_depth"] = (pm, helpers.make_reads(L, _reads_over(ref, 700, 40, 1200, alts=alts)), [ref], helpers.one_region(701, 1901, 40), [0])
    return out
