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
    mixed = r["mixed_cigars_window_edges"]
    assert mixed.n_cand >= 2 and mixed.planes["d"].sum() == 5 * 5 and mixed.planes["n"].sum() == 100 * 5
