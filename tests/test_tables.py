"""Packed-table construction (native host builder) against the reference's outputs and
against the literal O(M^2) oracle -- integer tables identical, SG float-bit-identical.
CPU only: the host builder lives in libsgm_b200.so but touches no GPU state."""
import numpy as np
import pytest

from conftest import load_golden
from golden_cases import CASES, LEV2KNOTS, knot_family, unpack_plan
from oracle import sg_oracle
from sgmethods_b200 import _lib, multi_index_sets as mis, nodes_1d
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
from sgmethods_b200.tp_inteprolants import (TPLagrangeInterpolator, TPPwCubicInterpolator,
                                            TPPwLinearInterpolator, TPPwQuadraticInterpolator,
                                            resolve_kind)

KIND_CLASS = {"linear": TPPwLinearInterpolator, "quadratic": TPPwQuadraticInterpolator,
              "cubic": TPPwCubicInterpolator, "lagrange": TPLagrangeInterpolator}


def build(name):
    fam, l2k, kind = CASES[name]
    g = load_golden(name)
    it = SGInterpolant(g["mid_set"], knot_family(fam, nodes_1d), LEV2KNOTS[l2k],
                       tp_interpolant=KIND_CLASS[kind], verbose=False)
    return it, g


@pytest.mark.parametrize("name", sorted(CASES))
def test_tables_identical_to_reference(name):
    it, g = build(name)
    dims, shapes, maps = unpack_plan(g)
    assert np.array_equal(np.asarray(it.combination_coeffs, dtype=np.int64), g["coeffs"])
    assert all(isinstance(c, np.int64) for c in it.combination_coeffs)
    assert np.array_equal(np.asarray(it.active_mids).reshape(len(dims), -1), g["active_mids"])
    assert it.num_nodes == int(g["num_nodes"])
    assert it.SG.shape == g["SG"].shape and it.SG.tobytes() == g["SG"].tobytes()
    assert len(it.map_tp_to_SG) == len(maps) == len(it.active_tp_dims)
    for t in range(len(maps)):
        assert np.array_equal(it.active_tp_dims[t], dims[t])
        assert it.map_tp_to_SG[t].shape == shapes[t]
        assert it.map_tp_to_SG[t].dtype == np.int64
        assert np.array_equal(it.map_tp_to_SG[t], maps[t])
        nodes = it.active_tp_nodes_list[t]
        assert len(nodes) == len(shapes[t])
        if dims[t].size:
            assert tuple(len(n) for n in nodes) == shapes[t]


def test_survey_probe_map_and_first_seen_coordinates():
    """smolyak(3,3), cc_nodes, 2^n+1: SG[1] keeps -cos(pi/2) = -6.1e-17 and the
    (3,3) map of term [0,1,1] is [[11,9,12],[0,1,2],[13,10,14]]."""
    it = SGInterpolant(mis.smolyak_mid_set(3, 3), nodes_1d.cc_nodes,
                       lambda n: np.where(n == 0, 1, 2 ** n + 1), verbose=False)
    assert it.SG[1, 2] == -np.cos(np.pi / 2) and it.SG[1, 2] != 0.0
    t = [i for i, m in enumerate(it.active_mids) if m.tolist() == [0, 1, 1]][0]
    assert it.map_tp_to_SG[t].tolist() == [[11, 9, 12], [0, 1, 2], [13, 10, 14]]


@pytest.mark.parametrize("w,N", [(0, 3), (1, 1), (2, 2), (3, 4), (5, 3)])
def test_coefficients_equal_window_algorithm(w, N):
    S = mis.smolyak_mid_set(w, N)
    assert np.array_equal(_lib.combination_coeffs(S), sg_oracle.combination_coefficients(S))


def test_coefficients_on_other_set_shapes():
    rng = np.random.default_rng(3)
    sets = [mis.tensor_product_mid_set(2, 3),
            mis.aniso_smolyak_mid_set(4.5, 4, np.array([0.7, 1.1, 1.3, 2.9])),
            mis.aniso_smolyak_mid_set(3, 12, np.linspace(1, 2, 12))]
    # a random downward-closed set
    S = mis.tensor_product_mid_set(3, 3)
    keep = S[rng.random(S.shape[0]) < 0.5]
    closed = {tuple(r) for r in keep.tolist()}
    for r in list(closed):
        for a in range(r[0] + 1):
            for b in range(r[1] + 1):
                for c in range(r[2] + 1):
                    closed.add((a, b, c))
    sets.append(np.array(sorted(closed)))
    for S in sets:
        assert np.array_equal(_lib.combination_coeffs(S), sg_oracle.combination_coefficients(S))


def test_coefficients_reject_bad_sets():
    with pytest.raises(ValueError):
        _lib.combination_coeffs(np.array([[0, 1], [0, 0]]))       # not sorted
    with pytest.raises(ValueError):
        _lib.combination_coeffs(np.array([[0, 0], [0, 0]]))       # repeated
    with pytest.raises(ValueError):
        _lib.combination_coeffs(np.array([[0, 0], [2, 0]]))       # leading entries not contiguous
    with pytest.raises(ValueError):
        _lib.combination_coeffs(np.array([[1, 0]]))


@pytest.mark.parametrize("fam,l2k", [("cc", "doubling"), ("cc", "n+1"), ("unb", "2^(n+1)-1"),
                                     ("equi", "3n+1"), ("hermite", "n+1")])
def test_grid_equals_literal_oracle(fam, l2k):
    """Hash-based node numbering == the reference's all-pairs L1 search (oracle), also for
    knot families that are NOT nested (cc with n+1, hermite)."""
    S = mis.aniso_smolyak_mid_set(3, 4, [1, 1, 1.5, 2.5])
    knots, lev = knot_family(fam, nodes_1d), LEV2KNOTS[l2k]
    it = SGInterpolant(S, knots, lev, verbose=False)
    plan = sg_oracle.make_plan(S, knots, lev)
    assert it.num_nodes == plan.num_nodes
    assert it.SG.tobytes() == plan.SG.tobytes()
    for a, b in zip(it.map_tp_to_SG, plan.maps):
        assert np.array_equal(a, b)
    for a, b in zip(it.active_tp_dims, plan.dims):
        assert np.array_equal(a, b)


def test_csr_nodes_match_dense():
    it = SGInterpolant(mis.smolyak_mid_set(3, 3), nodes_1d.cc_nodes,
                       lambda n: np.where(n == 0, 1, 2 ** n + 1), verbose=False)
    row_off, cols, vals = it._grid.nodes_csr()
    dense = np.zeros_like(it.SG)
    for r in range(it.num_nodes):
        dense[r, cols[row_off[r]:row_off[r + 1]]] = vals[row_off[r]:row_off[r + 1]]
    assert dense.tobytes() == it.SG.tobytes()


def test_knots_closer_than_tolerance_are_rejected():
    bad = lambda n: np.array([0.0, 4e-11, 8e-11, 1.2e-10])[:n] if n > 1 else 0  # noqa: E731
    with pytest.raises(ValueError):
        SGInterpolant(mis.smolyak_mid_set(3, 2), bad, lambda n: n + 1, verbose=False)


def test_large_set_builds_fast_and_consistently():
    """Config-2-like shape, small enough for CI: the builder is O(sum S_t)."""
    S = mis.smolyak_mid_set(3, 16)
    it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1,
                       verbose=False)
    assert len(it.combination_coeffs) == 969 and it.num_nodes == 7105   # SURVEY 3.1 probe
    assert int(np.sum(it._coeffs)) == 1                                  # partition of unity
    assert it._maps.max() == it.num_nodes - 1


def test_resolve_kind_accepts_factories_and_rejects_user_classes():
    assert resolve_kind(TPPwCubicInterpolator) == _lib.KIND_CUBIC
    assert resolve_kind(lambda nodes, f: TPLagrangeInterpolator(nodes, f)) == _lib.KIND_LAGRANGE

    class Mine:
        def __init__(self, nodes, f):
            pass
    with pytest.raises(NotImplementedError):
        resolve_kind(Mine)


def test_sample_on_SG_recycling_matches_reference():
    g = load_golden("recycle")
    l2k = lambda n: np.where(n == 0, 1, 2 ** n + 1)  # noqa: E731
    fvec = lambda y: np.array([np.sin(np.sum(y)), np.sum(y * y)])  # noqa: E731
    small = SGInterpolant(mis.smolyak_mid_set(2, 2), nodes_1d.cc_nodes, l2k, verbose=False)
    assert small.SG.tobytes() == g["small_SG"].tobytes()
    s_small = small.sample_on_SG(fvec)
    assert np.array_equal(s_small, g["s_small"])
    big = SGInterpolant(mis.smolyak_mid_set(3, 3), nodes_1d.cc_nodes, l2k, verbose=False)
    calls = []

    def f_count(y):
        calls.append(1)
        return fvec(y)
    s_big = big.sample_on_SG(f_count, old_xx=small.SG, old_samples=s_small)
    assert np.array_equal(s_big, g["s_big"]) and len(calls) == int(g["n_new"])
    mid3 = SGInterpolant(mis.smolyak_mid_set(2, 3), nodes_1d.cc_nodes, l2k, verbose=False)
    back = mid3.sample_on_SG(lambda y: 1 / 0, old_xx=big.SG, old_samples=s_big)
    assert np.array_equal(back, g["back"])
    # shrinking the dimension projects nodes whose tail norm truncates to 0 onto each other:
    # the reference's assertion fires (sparse_grid_interpolant.py:224), so does ours
    with pytest.raises(AssertionError):
        small.sample_on_SG(lambda y: 1 / 0, old_xx=big.SG, old_samples=s_big)


def test_sample_on_SG_scalar_function_and_parallel_pool():
    it = SGInterpolant(mis.smolyak_mid_set(2, 2), nodes_1d.cc_nodes,
                       lambda n: np.where(n == 0, 1, 2 ** n + 1), n_parallel=2, verbose=False)
    vals = it.sample_on_SG(_module_level_f)
    assert vals.shape == (it.num_nodes, 1)
    assert np.allclose(vals[:, 0], [_module_level_f(y) for y in it.SG])
    it.n_parallel = 0
    with pytest.raises(ValueError):
        it.sample_on_SG(_module_level_f)


def _module_level_f(y):
    return float(np.exp(y[0]) + y[1])
