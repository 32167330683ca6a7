"""The edge cases the reference handles implicitly (panics, skips, window quirks), pinned on the oracle."""
import numpy as np
import pytest

import edge_cases
import oracle_binding as ob
from longcallr_b200 import abi, host

CASES = edge_cases.cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_edge_case_status_and_mode_agreement(name):
    p, reads, refs, regions, status = CASES[name]
    b = host.BatchView(reads, regions)
    r0, r1 = ob.run(p, b, refs, mode=0), ob.run(p, b, refs, mode=1)
    if status is not None:
        assert list(r0.region_status) == status and list(r1.region_status) == status
    for f in ("pos", "alleles", "variant_type", "genotype", "haplotype", "flags", "phase_set", "depth"):
        np.testing.assert_array_equal(r0.cand[f], r1.cand[f])
    np.testing.assert_array_equal(r0.hp, r1.hp)
    np.testing.assert_array_equal(r0.ps, r1.ps)
    x, y = r0.cand["phase_score"].astype(float), r1.cand["phase_score"].astype(float)
    assert ((np.isnan(x) == np.isnan(y)) & (np.isinf(x) == np.isinf(y))).all()
    fin = np.isfinite(x)
    assert np.allclose(x[fin], y[fin], rtol=1e-9, atol=1e-9)


def test_specific_expectations():
    r = {n: ob.run(c[0], host.BatchView(c[1], c[3]), c[2], mode=0) for n, c in CASES.items()}
    assert r["empty_batch"].n_cand == 0 and r["region_without_reads"].n_cand == 0
    assert r["all_reads_filtered"].stats["n_reads_pass"] == 0 and r["all_reads_filtered"].n_cand == 0
    assert r["masked_reference_bytes"].n_cand == 0
    assert r["baseq_zero_off_site"].n_cand == 3 and (r["baseq_zero_off_site"].hp >= 0).sum() == 12
    z = r["baseq_zero_at_site"]
    assert z.n_cand == 3 and (z.hp == -1).all() and (z.is_fragment == 0).all()
    assert r["bad_cigar"].n_cand == 0
    m = r["missing_reference"]
    assert m.cand_off.tolist() == [0, 0, 3]
    d = r["dense_and_triallelic"]
    assert (d.cand["flags"] & abi.CF_DENSE != 0).sum() >= 5 and (d.cand["variant_type"] == 3).sum() == 1
    o = r["ont_trim_and_strand_bias"]
    assert [int(x) for x in o.cand["pos"]] == []  # ends trimmed, the middle site fails the one-strand test
    a = r["above_max_depth"]
    assert a.n_cand == 0
    # hard clip + soft clip: bases are read from behind the soft clip, so the three planted sites are called with clean counts
    for nm in ("hard_then_soft_clips", "hard_then_soft_clips_ont"):
        h = r[nm]
        assert h.stats["n_aligned_bases"] == 14 * 1200
        if nm == "hard_then_soft_clips":
            assert [int(x) for x in h.cand["pos"]] == [1000, 1300, 1600] and (h.cand["depth"] == 14).all()
            # no spurious mismatch anywhere: off the planted sites every piled base equals the reference base
            refw = CASES[nm][2][0][700:1900]
            onref = h.planes["acgt"][np.arange(1200), np.searchsorted(np.frombuffer(b"ACGT", dtype=np.uint8), refw)]
            off = np.ones(1200, bool)
            off[[300, 600, 900]] = False
            assert (onref[off] == h.planes["acgt"].sum(axis=1)[off]).all() and h.planes["acgt"].sum() > 14 * 1150
    # quality 0 at an RNA-edit candidate: the region runs, the site is rescued with an infinite score (one-sided) or left out-of-phase (both sides)
    one, both = r["baseq_zero_edit_site_one"], r["baseq_zero_edit_site_both"]
    assert list(one.region_status) == [0] and list(both.region_status) == [0]
    e = one.cand[one.cand["pos"] == 1105][0]
    # rescued with +inf / NaN scores (snpfrags.rs:245-251); the final assign_snp_haplotype_genotype re-scores it over the assigned reads
    assert e["flags"] & abi.CF_EDIT_LIST and e["flags"] & abi.CF_FOR_PHASING and not e["flags"] & abi.CF_RNA_EDITING and np.isfinite(e["phase_score"])
    assert one.hp[1] == 0 and (one.hp[[0, 2, 3]] > 0).all()  # the read with the quality-0 base goes to the unknown group (NaN compare)
    assert both.hp[1] == 0 and both.hp[4] == 0
    d = r["baseq_zero_edit_site_discordant"]
    e = d.cand[d.cand["pos"] == 1105][0]  # quality-0 bases on both sides: both scores NaN, the site stays out of phase
    assert e["flags"] & abi.CF_RNA_EDITING and e["flags"] & abi.CF_NON_SELECTED and not e["flags"] & abi.CF_FOR_PHASING
    assert (d.hp > 0).all()
    # a read that falls into two regions keeps the entry of the lower region
    sh = r["shared_reads_two_regions"]
    assert sh.cand_off.tolist()[1] > 0 and sh.cand_off.tolist()[2] > sh.cand_off.tolist()[1] and (sh.hp >= 0).all()
    mixed = r["mixed_cigars_window_edges"]
    assert mixed.n_cand >= 2 and mixed.planes["d"].sum() == 5 * 5 and mixed.planes["n"].sum() == 100 * 5
