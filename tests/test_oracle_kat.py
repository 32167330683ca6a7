"""Known-answer tests that pin the CPU oracle.

The reference ships no tests and cannot be built here (no cargo/rustc), so these values were derived
by hand from the cited reference lines: formula KATs, the contract primitives, and the demo region.
"""
import ctypes as C
import math

import numpy as np
import pytest

import helpers
import oracle_binding as ob
from longcallr_b200 import abi, host


def test_contract_log_exp_match_libm():
    L = ob.lib()
    rng = np.random.default_rng(0)
    xs = np.concatenate([10.0 ** rng.uniform(-300, 300, 2000), rng.uniform(0.5, 2.0, 2000), [1.0, 2.0, 10.0, 1e-310, 5e-324]])
    for x in xs:
        assert abs(L.lcr_oracle_log10(x) - math.log10(x)) <= 4e-16 * max(1.0, abs(math.log10(x)))
        assert abs(L.lcr_oracle_log(x) - math.log(x)) <= 4e-16 * max(1.0, abs(math.log(x)))
    for y in np.concatenate([rng.uniform(-300, 300, 2000), rng.uniform(-1, 1, 2000), [0.0, -0.3, -3000.0 / 10]]):
        ref = 10.0 ** y
        assert abs(L.lcr_oracle_exp10(y) - ref) <= 1e-14 * ref
    assert L.lcr_oracle_log10(0.0) == -math.inf and math.isnan(L.lcr_oracle_log10(-1.0))
    assert L.lcr_oracle_exp10(-400.0) == 0.0 and L.lcr_oracle_exp10(400.0) == math.inf


def test_formula_kats():
    L = ob.lib()
    # cal_strand_odds_ratio(5,5,9,1), candidate.rs:24-35,49-51
    assert abs(L.lcr_oracle_sor(5, 5, 9, 1) - 3.2580965) < 1e-6
    # log10(1 - 10^(-q/10)) for q = 10, 20, 30 (fragment.rs:133, phase.rs:44-47)
    for q, want in ((10, -0.045757490560675115), (20, -0.004364805402450088), (30, -0.0004345117740176917)):
        assert abs(math.log10(1.0 - 10.0 ** (-q / 10)) - want) < 1e-15
    # Rust `as i32`: saturating, NaN -> 0 (vcf.rs:51)
    assert L.lcr_oracle_f64_as_i32(math.inf) == 2147483647 and L.lcr_oracle_f64_as_i32(math.nan) == 0
    assert L.lcr_oracle_f64_as_i32(-1e300) == -2147483648 and L.lcr_oracle_f64_as_i32(3000.99) == 3000


def test_binomial_two_tailed_exact():
    """statrs Binomial::cdf replaced by exact integers: compare with an independent evaluation for every k <= n <= 30."""
    L = ob.lib()
    k_lo = [None, None, None, None, None, 0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 4, 4, 5, 5, 5, 6, 6, 7, 7, 7, 8, 8, 9]  # SURVEY 8c table, n = 1..30
    for n in range(1, 31):
        for k in range(0, n + 1):
            cdf = lambda j: sum(math.comb(n, i) for i in range(0, j + 1)) / 2.0 ** n  # noqa: E731
            if k == 0:
                p = 2.0 * cdf(0)
            elif k == n:
                p = 2.0 * (1.0 - cdf(n - 1))
            else:
                p = 2.0 * min(cdf(k), 1.0 - cdf(k - 1))
            assert bool(L.lcr_oracle_binom(k, n)) == (p < 0.05), (k, n, p)
            lo = k_lo[n - 1]
            assert bool(L.lcr_oracle_binom(k, n)) == (lo is not None and (k <= lo or k >= n - lo))


def test_rng_is_counter_based_and_uniform():
    L = ob.lib()
    u = np.array([L.lcr_oracle_uniform(7, 3, 1000, 1, c, i) for c in range(8) for i in range(500)])
    assert (u >= 0).all() and (u < 1).all() and abs(u.mean() - 0.5) < 0.02
    assert L.lcr_oracle_uniform(7, 3, 1000, 1, 2, 5) == L.lcr_oracle_uniform(7, 3, 1000, 1, 2, 5)
    assert L.lcr_oracle_uniform(7, 3, 1000, 1, 2, 5) != L.lcr_oracle_uniform(8, 3, 1000, 1, 2, 5)


def _phase_kat_batch():
    """Six reads over three het sites, q = 20, two clean haplotypes (SURVEY 8c phase KATs)."""
    L = 3000
    rng = np.random.default_rng(1)
    ref = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L)
    sites = [1000, 1200, 1400]
    alt = {}
    for s in sites:
        alt[s] = b"ACGT"[(b"ACGT".index(bytes([ref[s]])) + 1) % 4]
    recs = []
    for k in range(12):
        seq = bytearray(ref[800:1800].tobytes())
        if k % 2:
            for s in sites:
                seq[s - 800] = alt[s]
        recs.append(dict(pos=800, cigar="1000M", seq=seq.decode(), qual=20))
    return helpers.make_reads(L, recs), [ref], helpers.one_region(801, 1801, 12)


def test_phase_kat_two_haplotypes():
    reads, refs, regions = _phase_kat_batch()
    p = host.params_preset("hifi-masseq", seed=5, min_depth=6)
    for mode in (0, 1):
        r = ob.run(p, host.BatchView(reads, regions), refs, mode=mode)
        assert r.n_cand == 3 and (r.cand["variant_type"] == 1).all() and (r.cand["genotype"] == 0).all()
        # all three sites in one phase set named after the first (1-based) position
        assert (r.cand["phase_set"] == 1001).all()
        hp = r.hp
        assert set(hp[0::2]) | set(hp[1::2]) == {1, 2} and len(set(hp[0::2])) == 1 and len(set(hp[1::2])) == 1
        # the three sites carry the same delta (alt alleles ride together)
        assert len(set(r.cand["haplotype"])) == 1
        # phase score of a clean 6/6 split at q=20: -10 log10(L1 / (L2 + L3)), L1 = 12 log10(1-e), L2+L3 = 12 (log10(1-e) + log10 e)
        ok, err = math.log10(1 - 0.01), math.log10(0.01)
        want = -10.0 * math.log10((12 * ok) / (12 * ok + 12 * err))
        assert np.allclose(r.cand["phase_score"], want, atol=1e-9)


def test_demo_region_candidates_match_survey_table():
    """BASELINE config 1.  The 19-row table of SURVEY.md section 8c (pos1, ref, allele1, allele2, depth, QUAL, GQ)."""
    reads, refs, regions = helpers.load_demo_fixture()
    table = [(16730146, "G", "G", "T", 626, 3000, 2147483647), (16730717, "C", "T", "C", 26, 311, 281), (16733013, "A", "A", "G", 25, 77, 107),
             (16735430, "C", "C", "A", 61, 569, 599), (16735999, "T", "T", "C", 76, 807, 837), (16736400, "A", "A", "G", 85, 941, 971),
             (16736648, "C", "C", "T", 64, 734, 761), (16736735, "G", "A", "G", 64, 632, 662), (16737618, "T", "T", "A", 29, 183, 213),
             (16737747, "A", "A", "G", 29, 182, 212), (16737853, "G", "G", "A", 27, 189, 219), (16738139, "A", "A", "G", 11, 26, 56),
             (16738385, "T", "G", "T", 12, 173, 75), (16739109, "T", "G", "T", 29, 526, 93), (16740157, "G", "A", "G", 30, 493, 179),
             (16741247, "T", "T", "G", 456, 2627, 2657), (16742280, "T", "T", "C", 38, 58, 88), (16742727, "G", "C", "G", 6, 42, 58),
             (16743085, "C", "C", "T", 6, 41, 71)]
    L = ob.lib()
    for mode in (0, 1):
        p = host.params_preset("hifi-masseq", seed=1, flags=abi.LCR_FLAG_EMIT_PLANES | abi.LCR_FLAG_SKIP_PHASING)
        r = ob.run(p, host.BatchView(reads, regions), refs, mode=mode)
        got = [(int(c["pos"]) + 1, chr(c["reference"]), chr(c["alleles"][0]), chr(c["alleles"][1]), int(c["depth"]),
                L.lcr_oracle_f64_as_i32(float(c["variant_quality"])), L.lcr_oracle_f64_as_i32(float(c["genotype_quality"]))) for c in r.cand]
        assert got == table
        assert r.stats["n_reads_pass"] == 1697 and r.stats["n_aligned_bases"] == 2161129 + 1239
        assert r.planes["acgt"].sum() == 2161129 and r.planes["d"].sum() == 3575 and r.planes["n"].sum() == 15772322
        edits = [int(c["pos"]) + 1 for c in r.cand if c["flags"] & abi.CF_RNA_EDITING]
        assert edits == [16733013, 16736400, 16737747, 16738139]


def test_contract_and_reference_order_modes_agree():
    """Fixed-point contract vs sequential-f64 reference order: identical discrete outputs, FP fields within 1e-9."""
    cases = [helpers.load_demo_fixture()]
    for platform, preset, both in ((1, "ont-cdna", 1), (0, "hifi-masseq", 0)):
        syn = host.Synthetic(seed=31 + platform, contig_len=150_000, n_contigs=1, platform=platform, depth=25.0, n_het=120, n_edit=20, both_strands=both, max_intron=400, max_gap=800, n_threads=2)
        p = host.params_preset(preset)
        regions, _ = host.find_regions(syn.reads, p)
        cases.append((syn.reads, syn.reference.for_reads(syn.reads), regions, preset, syn))
    for case in cases:
        reads, refs, regions = case[0], case[1], case[2]
        preset = case[3] if len(case) > 3 else "hifi-masseq"
        p = host.params_preset(preset, seed=11)
        b = host.BatchView(reads, regions)
        a0, a1 = ob.run(p, b, refs, mode=0), ob.run(p, b, refs, mode=1)
        for f in helpers.INT_FIELDS:
            np.testing.assert_array_equal(a0.cand[f], a1.cand[f], err_msg=f)
        for f in helpers.FP_FIELDS:
            x, y = a0.cand[f].astype(float), a1.cand[f].astype(float)
            fin = np.isfinite(x) & np.isfinite(y)
            assert (np.isfinite(x) == np.isfinite(y)).all()
            assert np.allclose(x[fin], y[fin], rtol=1e-9, atol=1e-9), f
        np.testing.assert_array_equal(a0.hp, a1.hp)
        np.testing.assert_array_equal(a0.ps, a1.ps)


def test_planted_haplotypes_are_recovered():
    """Sanity against the statistical behaviour of the reference: planted het SNPs are called and reads split by haplotype."""
    syn = host.Synthetic(seed=77, contig_len=120_000, n_contigs=1, platform=0, depth=40.0, n_het=150, n_edit=0, both_strands=0, max_intron=300, max_gap=600, n_threads=2)
    p = host.params_preset("hifi-masseq", seed=2)
    regions, _ = host.find_regions(syn.reads, p)
    r = ob.run(p, host.BatchView(syn.reads, regions), syn.reference.for_reads(syn.reads), mode=0)
    het = r.cand[(r.cand["variant_type"] == 1) & (r.cand["flags"] & abi.CF_HET_VAR != 0)]
    called = set(int(x) for x in het["pos"])
    truth = set(int(x) for x in syn.het_pos)
    assert len(called & truth) >= 0.8 * len(truth)
    # within each phase set HP must be a relabelling of the planted read haplotype
    ps_ids = [x for x in np.unique(r.ps) if x]
    agree = total = 0
    for ps in ps_ids:
        m = (r.ps == ps) & (r.hp > 0)
        if m.sum() < 10:
            continue
        a = ((r.hp[m] == 1) == (syn.read_hap[m] == 0)).sum()
        agree += max(a, m.sum() - a)
        total += m.sum()
    assert total > 500 and agree / total > 0.97


def test_seeded_shuffle_of_downsample():
    """--downsample draws from StdRng::seed_from_u64(2025) (phase.rs:693-701).  The ChaCha core of the contract's restatement against the
    published zero-key, zero-nonce keystreams (ChaCha20: RFC 7539 appendix A.1 vector 1; ChaCha12 / ChaCha8: the reduced-round vectors of
    the same family); the shuffle against the independent Python restatement.  The rand crates themselves are not available here."""
    import ctypes as C
    import os
    import sys

    import numpy as np

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import py_restatement as py

    L = ob.lib()
    key = (C.c_uint32 * 8)()
    out = (C.c_uint32 * 16)()
    want = {20: "76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7", 12: "9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f",
            8: "3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e"}
    for rounds, hexs in want.items():
        L.lcr_oracle_chacha_block(key, C.c_uint64(0), rounds, out)
        assert bytes(out)[:32].hex() == hexs, rounds
        assert b"".join(int(w).to_bytes(4, "little") for w in py._chacha_block([0] * 8, 0, rounds))[:32].hex() == hexs
    for n, depth in ((1, 1), (2, 2), (10, 10), (1000, 64), (65536, 16), (65537, 16), (100003, 32)):
        idx = np.zeros(n, dtype="<u4")
        L.lcr_oracle_shuffle(C.c_uint64(2025), C.c_uint32(n), C.c_void_p(idx.ctypes.data))
        assert sorted(int(x) for x in idx) == list(range(n)) if n <= 1000 else len(set(int(x) for x in idx)) == n
        assert [int(x) for x in idx[:depth]] == py.downsample_fragments(n, depth)
