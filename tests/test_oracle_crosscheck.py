"""Differential test of the two independent restatements of the reference: oracle/lcr_oracle.cpp (C++, the oracle the
CUDA path is compared with) against oracle/py_restatement.py (Python, written from the Rust sources, libm arithmetic,
no shared header).  Random regions with random CIGARs, clips, qualities, strands and tags (hypothesis); either file
changing its behaviour fails here.  Also pins the contract's own math (include/lcr_contract.h) against libm.
"""
import math
import os
import sys

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import helpers  # noqa: E402
import oracle_binding as ob  # noqa: E402
import py_restatement as py  # noqa: E402
from longcallr_b200 import abi, host  # noqa: E402

STOP_AFTER_PHASE = 0x40000000  # oracle-only test hook
PARAM_KEYS = ("platform", "min_mapq", "min_baseq", "min_read_length", "divergence", "min_allele_freq", "min_allele_freq_include_intron", "min_qual",
              "use_strand_bias", "min_depth", "max_depth", "distance_to_read_end", "polya_tail_length", "dense_win_size", "min_dense_cnt", "min_linkers",
              "max_enum_snps", "low_allele_frac_cutoff", "low_allele_cnt_cutoff", "seed")


def params_dict(p):
    return {k: getattr(p, k) for k in PARAM_KEYS}


@st.composite
def regions(draw):
    """One small region: a reference window, a few planted variant sites, 8-40 reads with messy CIGARs."""
    rng = np.random.default_rng(draw(st.integers(0, 2 ** 32 - 1)))
    L = 900
    ref = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L)
    if draw(st.booleans()):
        ref[rng.integers(100, 800, size=3)] = np.frombuffer(b"Nac", dtype=np.uint8)  # N and soft-masked bytes never call
    n_var = draw(st.integers(0, 9))
    var_pos = sorted(set(int(x) for x in rng.integers(150, 750, size=n_var)))
    alt = {p: int(rng.choice([b for b in b"ACGT" if b != ref[p] and chr(b).upper() != chr(ref[p]).upper()])) for p in var_pos}
    preset = draw(st.sampled_from(["ont-cdna", "ont-drna", "hifi-isoseq", "hifi-masseq"]))
    n_reads = draw(st.integers(25, 70))
    recs = []
    for k in range(n_reads):
        pos = int(rng.integers(60, 300))
        hap = int(rng.integers(0, 2))
        ops, seq, qual = [], bytearray(), []
        lead_h, lead_s = int(rng.integers(0, 4)) == 0, int(rng.integers(0, 3)) == 0
        if lead_h:
            ops.append(("H", int(rng.integers(1, 9))))
        if lead_s:
            n = int(rng.integers(1, 25))
            ops.append(("S", n))
            seq += bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n))
        rp = pos
        n_ops = int(rng.integers(3, 11))
        for j in range(n_ops):
            kind = "M" if j % 2 == 0 else str(rng.choice(list("IDNM=X"), p=[0.25, 0.25, 0.2, 0.1, 0.1, 0.1]))
            n = int(rng.integers(20, 160)) if kind in "M=X" else (int(rng.integers(1, 6)) if kind in "ID" else int(rng.integers(20, 120)))
            if kind in "M=X":
                if rp + n >= L - 1:
                    break
                chunk = bytearray(ref[rp:rp + n].tobytes().upper())
                for p in var_pos:
                    if rp <= p < rp + n and hap == 1:
                        chunk[p - rp] = alt[p]
                for e in rng.integers(0, n, size=int(rng.integers(0, 3))):  # sequencing errors, the odd N and a lower-case letter
                    chunk[int(e)] = int(rng.choice(np.frombuffer(b"ACGTNa", dtype=np.uint8)))
                seq += chunk
                rp += n
            elif kind == "I":
                seq += bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n))
            else:
                if rp + n >= L - 1:
                    break
                rp += n
            ops.append((kind, n))
        if not any(o in "M=X" for o, _ in ops):
            continue
        if int(rng.integers(0, 3)) == 0:
            n = int(rng.integers(1, 25))
            ops.append(("S", n))
            seq += bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n))
        if int(rng.integers(0, 4)) == 0:
            ops.append(("H", int(rng.integers(1, 9))))
        if int(rng.integers(0, 6)) == 0:  # a poly-A tail near the end
            seq[-12:] = b"A" * min(12, len(seq))
        qual = [int(x) for x in rng.choice([1, 3, 9, 10, 11, 20, 29, 30, 31, 40], size=len(seq))]
        merged = []
        for o, n in ops:  # BAM never stores two equal adjacent ops
            if merged and merged[-1][0] == o:
                merged[-1] = (o, merged[-1][1] + n)
            else:
                merged.append((o, n))
        recs.append(dict(pos=pos, cigar="".join(f"{n}{o}" for o, n in merged), seq=bytes(seq).decode(), qual=qual, flag=int(rng.choice([0, 16, 0, 16, 0, 16, 0, 16, 0, 16, 256, 4, 2048])),
                         mapq=int(rng.choice([60] * 9 + [3])), ts=str(rng.choice(["+", "-", "*"])), de=float(rng.choice([0.01, 0.02, 0.01, 0.03, 0.02, 0.9, float("nan"), float("nan")]))))
    start = draw(st.integers(90, 200))
    end = draw(st.integers(500, 820))
    return preset, ref, recs, start, end


def to_py_reads(recs):
    out = []
    for r in sorted(recs, key=lambda r: r["pos"]):
        ops, num = [], ""
        for ch in r["cigar"]:
            if ch.isdigit():
                num += ch
            else:
                ops.append((ch, int(num)))
                num = ""
        out.append(dict(pos=r["pos"], cigar=ops, seq=r["seq"].encode(), qual=r["qual"], flag=r["flag"], mapq=r["mapq"], ts=r["ts"], de=r["de"]))
    return out


FLAG_OF = dict(rna_editing=abi.CF_RNA_EDITING, dense=abi.CF_DENSE, het_var=abi.CF_HET_VAR, hom_var=abi.CF_HOM_VAR, cand_somatic=abi.CF_CAND_SOMATIC)


@settings(max_examples=int(os.environ.get("LCR_FUZZ_EXAMPLES", "300")), deadline=None, suppress_health_check=list(HealthCheck), derandomize=not os.environ.get("LCR_FUZZ_RANDOM"), database=None)
@given(regions(), st.sampled_from([0, 0, 6, 15]))
def test_pileup_candidates_fragments_agree(case, ds_depth):
    preset, ref, recs, start, end = case
    reads = helpers.make_reads(len(ref), recs)
    region = helpers.one_region(start, end, len(recs))
    # ds_depth > 0: --downsample with that depth (regions here hold 10-60 fragments, so the seeded shuffle applies to most of them)
    p = host.params_preset(preset, seed=5, min_read_length=30, min_depth=4, downsample_depth=max(ds_depth, 1),
                           flags=abi.LCR_FLAG_EMIT_PLANES | abi.LCR_FLAG_EMIT_FRAGMENTS | STOP_AFTER_PHASE | (abi.LCR_FLAG_DOWNSAMPLE if ds_depth else 0))
    P = params_dict(p)
    batch = host.BatchView(reads, region)
    got = {m: ob.run(p, batch, [ref], mode=m) for m in (0, 1)}
    reg = dict(tid=0, start=start, end=end)
    pr = to_py_reads(recs)
    fv, n_pass, n_al = py.pileup(P, reg, pr, ref)
    cands, edit, somatic = py.candidates(P, reg, fv)
    for m in (0, 1):
        g = got[m]
        assert list(g.region_status) == [0]
        assert g.stats["n_reads_pass"] == n_pass and g.stats["n_aligned_bases"] == n_al
        np.testing.assert_array_equal(g.planes["acgt"], np.array([[b["a"], b["c"], b["g"], b["t"]] for b in fv], dtype=np.uint32).reshape(-1, 4))
        np.testing.assert_array_equal(g.planes["fwd"], np.array([[b["strands"][x][0] for x in "ACGT"] for b in fv], dtype=np.uint32).reshape(-1, 4))
        np.testing.assert_array_equal(g.planes["d"], [b["d"] for b in fv])
        np.testing.assert_array_equal(g.planes["n"], [b["n"] for b in fv])
        np.testing.assert_array_equal(g.planes["ts"], np.array([b["ts"] for b in fv], dtype=np.uint32).reshape(-1, 2))
        assert g.n_cand == len(cands), (m, [int(x) for x in g.cand["pos"]], [c["pos"] for c in cands])
        for i, c in enumerate(cands):
            r = g.cand[i]
            assert int(r["pos"]) == c["pos"] and chr(r["reference"]) == c["reference"] and (chr(r["alleles"][0]), chr(r["alleles"][1])) == c["alleles"]
            assert int(r["depth"]) == c["depth"] and tuple(float(x) for x in r["allele_freqs"]) == c["freqs"]
            for name, bit in FLAG_OF.items():
                assert bool(r["flags"] & bit) == c[name], (m, i, name)
            for a, b in ((float(r["variant_quality"]), c["variant_quality"]), (float(r["genotype_quality"]), c["genotype_quality"])):
                assert (math.isinf(a) and math.isinf(b)) or abs(a - b) <= 1e-9 * max(1.0, abs(b)), (m, i, a, b)
            assert np.allclose(r["genotype_probability"], c["gp"], rtol=1e-9, atol=1e-300)
    # fragments (the enumeration below also needs them)
    frags = py.fragments(P, reg, pr, cands)
    for m in (0, 1):
        fr = got[m].fragments
        assert int(fr["frag_off"][-1]) == len(frags)
        np.testing.assert_array_equal(fr["frag_read"], [f["read"] for f in frags])
        np.testing.assert_array_equal(fr["elem_off"], np.cumsum([0] + [len(f["list"]) for f in frags]))
        np.testing.assert_array_equal(fr["elem_snp"], [fe["snp"] for f in frags for fe in f["list"]])
        np.testing.assert_array_equal(fr["elem_cell"], [fe["p"] * (fe["baseq"] + 1) for f in frags for fe in f["list"]])
        np.testing.assert_array_equal(fr["elem_base"], [ord(fe["base"]) for f in frags for fe in f["list"]])
    # phase(): the 2^n enumeration with the contract's random source, when no phase site holds a quality-0 base (none here) and n <= 10
    if len(cands) <= P["max_enum_snps"] and len(cands) <= 6:
        if ds_depth and len(frags) >= ds_depth:  # thread.rs:144-151
            chosen = set(py.downsample_fragments(len(frags), ds_depth))
            for k, f in enumerate(frags):
                f["ds_skip"] = k not in chosen
        rel = {i: i for i in range(len(pr))}  # the region's read range starts at read 0
        counters = py.phase_enum(P, reg, cands, frags, rel)
        want_hp = np.full(len(pr), -1, dtype=np.int8)
        for f in frags:
            want_hp[f["read"]] = 1 if f["haplotag"] == 1 else (2 if f["haplotag"] == -1 else 0)
        hap = np.array([c["haplotype"] for c in cands], dtype=np.int8)
        gen = np.array([c["genotype"] for c in cands], dtype=np.int8)
        # reads whose phase sites are all homozygous in the final state keep the random haplotag of the winning start: arbitrary
        informative = np.zeros(len(pr), dtype=bool)
        tied = np.zeros(len(pr), dtype=bool)
        for f in frags:
            informative[f["read"]] = f["for_phasing"] and any(fe["phase_site"] and cands[fe["snp"]]["genotype"] == 0 for fe in f["list"])
            # a read whose heterozygous phase sites split into two sides of identical (capped) qualities: the two sums are the same terms
            # in a different order - an exact tie for the contract's integers, rounding noise for f64 - and the read keeps whatever it had
            side = {1: [], -1: []}
            for fe in f["list"]:
                if fe["phase_site"] and cands[fe["snp"]]["genotype"] == 0:
                    side[fe["p"] * cands[fe["snp"]]["haplotype"]].append(min(fe["baseq"], 30))
            tied[f["read"]] = bool(side[1]) and sorted(side[1]) == sorted(side[-1])
        for m in (0, 1):
            g = got[m]
            assert g.stats["n_cross_optimize"] == counters["calls"], (m, g.stats, counters)
            np.testing.assert_array_equal(g.cand["genotype"], gen, err_msg=f"mode {m}")
            same = np.array_equal(g.cand["haplotype"], hap) and np.array_equal(g.hp[informative], want_hp[informative])
            if m == 1:
                # the same f64 arithmetic in the same order: identical winner, identical state, identical iteration count
                assert same and np.array_equal(g.hp, want_hp) and g.stats["n_sweep_iters"] == counters["iters"], (m, g.stats, counters)
                continue
            # the fixed-point contract breaks exact ties between starts by index where f64 breaks them by rounding noise: a start and its
            # complement reach mirror-image optima of equal objective, so the contract's answer may be the mirror image
            het = gen == 0
            decided = informative & ~tied
            same = np.array_equal(g.cand["haplotype"], hap) and np.array_equal(g.hp[decided], want_hp[decided])
            mirror_hp = np.where(want_hp == 1, 2, np.where(want_hp == 2, 1, want_hp))
            mirrored = np.array_equal(g.cand["haplotype"][het], -hap[het]) and np.array_equal(g.hp[decided], mirror_hp[decided])
            assert same or mirrored, (m, list(g.cand["haplotype"]), list(hap), list(g.hp), list(want_hp))


@settings(max_examples=int(os.environ.get("LCR_FUZZ_EXAMPLES", "300")) // 3, deadline=None, suppress_health_check=list(HealthCheck), derandomize=not os.environ.get("LCR_FUZZ_RANDOM"), database=None)
@given(regions(), st.lists(st.tuples(st.integers(0, 400), st.integers(1, 60)), min_size=1, max_size=5))
def test_exon_only_mask_agrees(case, raw_exons):
    """--exon-only (candidate.rs:80-89): candidates only at positions some exon interval of the region holds; both restatements."""
    preset, ref, recs, start, end = case
    reads = helpers.make_reads(len(ref), recs)
    region = helpers.one_region(start, end, len(recs))
    exons = [(start + a, start + a + ln) for a, ln in raw_exons]  # 1-based start, stop exclusive, unsorted, overlapping
    p = host.params_preset(preset, seed=5, min_read_length=30, min_depth=4, flags=abi.LCR_FLAG_EMIT_PLANES | abi.LCR_FLAG_SKIP_PHASING)
    P = params_dict(p)
    reg = dict(tid=0, start=start, end=end)
    fv, _, _ = py.pileup(P, reg, to_py_reads(recs), ref)
    cands, _, _ = py.candidates(P, reg, fv, exons)
    every, _, _ = py.candidates(P, reg, fv)
    inside = lambda pos: any(s <= pos + 1 < e for s, e in exons)  # noqa: E731
    assert [c["pos"] for c in cands] == [c["pos"] for c in every if inside(c["pos"])]
    for m in (0, 1):
        g = ob.run(p, host.BatchView(reads, region, exons=[exons]), [ref], mode=m)
        assert list(g.region_status) == [0] and [int(x) for x in g.cand["pos"]] == [c["pos"] for c in cands]
        for i, c in enumerate(cands):
            for name, bit in FLAG_OF.items():
                assert bool(g.cand[i]["flags"] & bit) == c[name], (m, i, name)
    # a region whose genes have no exon is skipped altogether (thread.rs:88-91)
    g = ob.run(p, host.BatchView(reads, region, exons=[[]]), [ref], mode=0)
    assert list(g.region_status) == [abi.LCR_REGION_NO_EXON] and g.n_cand == 0 and g.stats["n_reads_pass"] == 0


@settings(max_examples=int(os.environ.get("LCR_FUZZ_EXAMPLES", "300")) // 3, deadline=None, suppress_health_check=list(HealthCheck), derandomize=not os.environ.get("LCR_FUZZ_RANDOM"), database=None)
@given(regions(), st.lists(st.tuples(st.integers(0, 300), st.integers(0, 4), st.one_of(st.floats(-5, 60, width=32), st.just(float("nan")))), min_size=1, max_size=12, unique_by=lambda t: t[0]))
def test_imported_candidates_agree(case, raw):
    """-v (candidate.rs:530-613): candidates at the listed positions with the record's genotype class and QUAL, alleles / frequencies /
    depth from the pileup (also on uncovered positions), no filter, no dense pass; fragments are then built on those candidates."""
    preset, ref, recs, start, end = case
    reads = helpers.make_reads(len(ref), recs)
    region = helpers.one_region(start, end, len(recs))
    ext = sorted((start - 1 + a, gt, q) for a, gt, q in raw if start - 1 + a < end - 1)
    p = host.params_preset(preset, seed=5, min_read_length=30, min_depth=4, flags=abi.LCR_FLAG_EMIT_PLANES | abi.LCR_FLAG_EMIT_FRAGMENTS | STOP_AFTER_PHASE)
    P = params_dict(p)
    reg = dict(tid=0, start=start, end=end)
    prr = to_py_reads(recs)
    fv, _, _ = py.pileup(P, reg, prr, ref)
    cands = py.import_external_candidates(reg, fv, {pos: (gt, q) for pos, gt, q in ext})
    frags = py.fragments(P, reg, prr, cands) if cands else []
    for m in (0, 1):
        g = ob.run(p, host.BatchView(reads, region, external=[ext]), [ref], mode=m)
        assert list(g.region_status) == [0] and [int(x) for x in g.cand["pos"]] == [c["pos"] for c in cands], (m, ext)
        for i, c in enumerate(cands):
            r = g.cand[i]
            assert chr(r["reference"]) == c["reference"] and (chr(r["alleles"][0]), chr(r["alleles"][1])) == c["alleles"] and int(r["depth"]) == c["depth"]
            assert int(r["variant_type"]) == c["variant_type"]
            for a, b in zip([float(x) for x in r["allele_freqs"]] + [float(r["variant_quality"]), float(r["genotype_quality"])], list(c["freqs"]) + [c["variant_quality"], c["genotype_quality"]]):
                assert (math.isnan(a) and math.isnan(b)) or a == b, (m, i, a, b)
            for name, bit in FLAG_OF.items():
                assert bool(r["flags"] & bit) == c[name], (m, i, name)
            assert bool(r["flags"] & abi.CF_FOR_PHASING) == c["for_phasing"]
        fr = g.fragments
        assert int(fr["frag_off"][-1]) == len(frags)
        np.testing.assert_array_equal(fr["elem_snp"], [fe["snp"] for f in frags for fe in f["list"]])
        np.testing.assert_array_equal(fr["elem_cell"], [fe["p"] * (fe["baseq"] + 1) for f in frags for fe in f["list"]])


def test_contract_math_against_libm():
    """lcr_log10 / lcr_exp10 / lcr_log (the deterministic math both the oracle's contract mode and the CUDA path use) against libm
    on the values the path feeds them: the q = 0..30 tables, likelihood sums, posteriors."""
    L = ob.lib()
    xs = [10.0 ** (-q / 10.0) for q in range(0, 31)] + [1.0 - 10.0 ** (-q / 10.0) for q in range(1, 31)] + [0.1 ** (q / 10.0) for q in range(0, 31)]
    xs += [1e-300, 10e-301, 5e-4, 1e-3, 1 - 1.5e-3, 0.5, 2.0, 3.0, 1234.5, 1e300] + list(np.random.default_rng(0).uniform(1e-12, 1.0, 2000))
    for x in xs:
        a, b = L.lcr_oracle_log10(x), math.log10(x)
        assert abs(a - b) <= 4e-16 * max(1.0, abs(b)), (x, a, b)
        a, b = L.lcr_oracle_log(x), math.log(x)
        assert abs(a - b) <= 4e-16 * max(1.0, abs(b)), (x, a, b)
    for y in list(np.random.default_rng(1).uniform(-320.0, 3.0, 3000)) + [0.0, -0.5, -3000.0 / 10, -1.0, -30.0, 2.0]:
        a, b = L.lcr_oracle_exp10(y), math.pow(10.0, y)
        assert (a == b == 0.0) or abs(a - b) <= 4e-15 * abs(b) + 1e-320, (y, a, b)
    # f32 ln of the strand odds ratio (candidate.rs:24-35) against an f32 evaluation with libm
    for args in [(5, 5, 9, 1), (0, 0, 0, 0), (10, 3, 2, 8), (100, 90, 1, 30), (7, 7, 7, 7), (1, 30, 30, 1)]:
        assert L.lcr_oracle_sor(*args) == py.strand_odds_ratio(*args), args
    # the two-tailed binomial test against exact rational arithmetic, every (k, n) the path can ask for
    for n in range(1, 31):
        for k in range(0, n + 1):
            assert bool(L.lcr_oracle_binom(k, n)) == (py.binomial_two_tailed(k, n) < 0.05), (k, n)
    # the random source of the contract as restated from its description
    for args in [(1, 0, 100, 1, 0, 0), (20251017, 3, 123456, 1, 77, 12), (2 ** 63 + 5, 11, 16729961, 4, 3, 900)]:
        seed, tid, start, stream, call, idx = args
        assert L.lcr_oracle_uniform(seed, tid, start, stream, call, idx) == py.uniform(seed, py.region_key(tid, start), stream, call, idx)


@pytest.mark.parametrize("cig,lead,trail", [("5H10S50M7S3H", 10, 7), ("10S50M", 10, 0), ("3H50M4S", 0, 4), ("2H50M9H", 0, 0), ("50M", 0, 0), ("4S", 4, 4), ("1H2S", 2, 2)])
def test_clip_semantics(cig, lead, trail):
    ops, num = [], ""
    for ch in cig:
        if ch.isdigit():
            num += ch
        else:
            ops.append((ch, int(num)))
            num = ""
    assert py.leading_softclips(ops) == lead and py.trailing_softclips(ops) == trail
