"""The quadrature rules of the Wachspress integration (config_wachspress_integration_type / _order, Registry.xml:603-610;
get_integration_factors, src/shared/mpas_seaice_velocity_solver_wachspress.F:1224-1941): 'dunavant' orders 1-10 and 12,
'fekete' orders 1-6, 8, 9, 'trapezoidal'.

tests/golden/cpu/refexec_quadrature_rules.npz holds every rule as the reference's own routines return it (executed by
tests/golden/fortran_subset.py; generator tests/golden/make_quadrature_tables.py).  Checked here, bit for bit and without a
GPU: the oracle's rules (oracle/evp_precompute_oracle.c) and the rules the shipped library integrates with
(evp_integration_rule, host-only).  Orders the reference does not have are refused by both."""
import os

import numpy as np
import pytest

import oracle
from mpas_seaice_b200 import host

HERE = os.path.dirname(os.path.abspath(__file__))
Z = np.load(os.path.join(HERE, "golden", "cpu", "refexec_quadrature_rules.npz"))
RULES = sorted((k.rsplit("_", 1)[0], int(k.rsplit("_", 1)[1])) for k in Z.files if k != "provenance" and not k.endswith("_norm"))


def test_the_fixture_holds_every_rule_of_the_reference():
    prov = str(Z["provenance"])
    for name in ("get_integration_factors_dunavant", "get_integration_factors_fekete", "get_integration_factors_trapezoidal"):
        assert name in prov
    assert [o for t, o in RULES if t == "dunavant"] == [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12]
    assert [o for t, o in RULES if t == "fekete"] == [1, 2, 3, 4, 5, 6, 8, 9]
    assert [o for t, o in RULES if t == "trapezoidal"] == [1, 2, 3, 4, 5]


@pytest.mark.parametrize("rule", RULES, ids=["%s%d" % r for r in RULES])
def test_oracle_and_library_hold_the_reference_executed_rule(rule):
    kind, order = rule
    want = Z["%s_%d" % rule]
    norm = float(Z["%s_%d_norm" % rule])
    u, v, w, n = oracle.integration_factors(kind, order)
    assert np.array_equal(np.stack([u, v, w]), want) and n == norm
    u, v, w, n = host.integration_rule(kind, order)
    assert np.array_equal(np.stack([u, v, w]), want) and n == norm
    # a rule for the unit triangle: the weights add up to the normalisation's share of its area, the points lie in it
    assert abs(w.sum() / norm - 0.5) < 1e-12
    assert np.all(u >= 0) and np.all(v >= 0) and np.all(u + v <= 1 + 1e-15)
    if kind != "trapezoidal" or order >= 1:
        assert abs((u * w).sum() / norm - 1.0 / 6.0) < 1e-12         # exact for linear integrands


@pytest.mark.parametrize("kind,order", [("dunavant", 11), ("dunavant", 13), ("dunavant", 0), ("fekete", 7), ("fekete", 10),
                                        ("trapezoidal", 0), ("trapezoidal", 11)])
def test_orders_the_reference_does_not_have_are_refused(kind, order):
    """(the reference leaves u, v, weights unallocated for them; trapezoidal order 11 has 78 points, beyond the 64 the
    device's constant tables hold)"""
    with pytest.raises(host.EvpError, match="unsupported integration rule"):
        host.integration_rule(kind, order)
    if not (kind == "trapezoidal" and order == 11):
        with pytest.raises(Exception):
            oracle.integration_factors(kind, order)
