"""Pin the CPU oracle (oracle/) against outputs of the unmodified reference (CPU only).

Fixtures: tests/golden/*.npz written by tests/golden/make_golden.py from /root/reference.
Integer tables must be identical; the sparse grid coordinates must be bit-identical; the
interpolated values are required to be bit-identical too (the oracle performs the same
floating-point operations in the same order), with 1e-14 relative slack allowed only where
numpy may legally reassociate (none observed)."""
import numpy as np
import pytest

from conftest import load_golden
from golden_cases import CASES, LEV2KNOTS, knot_family, unpack_plan
from oracle import brownian_oracle, sg_oracle
from sgmethods_b200 import nodes_1d


def build_oracle_plan(name):
    fam, l2k, kind = CASES[name]
    g = load_golden(name)
    plan = sg_oracle.make_plan(g["mid_set"], knot_family(fam, nodes_1d), LEV2KNOTS[l2k])
    return plan, g, kind


@pytest.mark.parametrize("name", sorted(CASES))
def test_tables_match_reference(name):
    if name in ("lag_example0", "lin_example0"):
        pytest.skip("2880-node O(M^2) literal build: covered by the slow marker below")
    plan, g, _ = build_oracle_plan(name)
    dims, shapes, maps = unpack_plan(g)
    assert np.array_equal(np.asarray(plan.coeffs, dtype=np.int64), g["coeffs"])
    assert np.array_equal(np.asarray(plan.mids).reshape(len(plan.coeffs), -1), g["active_mids"])
    assert plan.num_nodes == int(g["num_nodes"])
    assert plan.SG.tobytes() == g["SG"].tobytes()            # float bit-exact
    for t in range(len(dims)):
        assert np.array_equal(plan.dims[t], dims[t])
        assert plan.shapes[t] == shapes[t]
        assert np.array_equal(plan.maps[t], maps[t])


@pytest.mark.parametrize("name", sorted(CASES))
def test_interpolate_matches_reference(name):
    fam, l2k, kind = CASES[name]
    g = load_golden(name)
    dims, shapes, maps = unpack_plan(g)
    knots = knot_family(fam, nodes_1d)
    nodes = [tuple(knots(m) for m in shp) for shp in shapes]
    plan = sg_oracle.plan_from_tables(g["mid_set"].shape[1], g["coeffs"], dims, nodes, maps)
    out = sg_oracle.interpolate(plan, g["x"], g["f_on_SG"], kind)
    ref = g["out"]
    assert out.shape == ref.shape
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.array_equal(out[ok], ref[ok]), np.max(np.abs(out[ok] - ref[ok]))
    if kind == "linear":       # the literal third-party call gives the same bits
        out2 = sg_oracle.interpolate(plan, g["x"], g["f_on_SG"], "linear_scipy")
        assert np.array_equal(out2[ok], ref[ok])


def test_example0_tables_match_reference():
    """The 2880-node example/0 grid through the literal quadratic-cost builder (~20 s)."""
    plan, g, _ = build_oracle_plan("lin_example0")
    dims, shapes, maps = unpack_plan(g)
    assert np.array_equal(np.asarray(plan.coeffs, dtype=np.int64), g["coeffs"])
    assert plan.SG.tobytes() == g["SG"].tobytes()
    assert all(np.array_equal(a, b) for a, b in zip(plan.maps, maps))


def test_linear_brackets_follow_scipy_probe():
    """SURVEY section 8(c) probe of scipy's find_indices."""
    g = np.array([-1, 0, .5, 2.])
    x = np.array([-3, -1, -.5, 0, .25, .5, 2, 5, np.nan])
    i, y = sg_oracle._brackets_linear(g, x)
    assert i[:-1].tolist() == [0, 0, 0, 1, 1, 2, 2, 2]
    assert np.allclose(y[:-1], [-2, 0, .5, 0, .5, 0, 1, 3]) and np.isnan(y[-1])


def test_linear_operator_equals_scipy_on_random_grids():
    rng = np.random.default_rng(5)
    for k in (1, 2, 3, 4):
        nodes = tuple(np.sort(rng.normal(size=rng.integers(2, 7))) for _ in range(k))
        vals = rng.normal(size=tuple(n.size for n in nodes) + (3,))
        x = rng.normal(0, 2, (200, k))
        assert np.array_equal(sg_oracle.tp_linear(nodes, vals, x),
                              sg_oracle.tp_linear_scipy(nodes, vals, x))


def test_quadratic_right_extrapolation_raises_like_reference():
    nodes = (np.linspace(-1, 1, 5),)
    vals = np.arange(5.)[:, None]
    with pytest.raises(IndexError):
        sg_oracle.tp_quadratic(nodes, vals, np.array([[1.0000001]]))
    with pytest.raises(IndexError):
        sg_oracle.tp_quadratic(nodes, vals, np.array([[np.nan]]))
    assert np.isfinite(sg_oracle.tp_quadratic(nodes, vals, np.array([[-1.7], [1.0]]))).all()


def test_cubic_needs_four_knots_like_reference():
    with pytest.raises(UnboundLocalError):
        sg_oracle.tp_cubic((np.linspace(-1, 1, 3),), np.zeros((3, 1)), np.array([[0.1]]))


def test_lagrange_nan_on_knot_like_reference():
    nodes = (np.linspace(-1, 1, 3),)
    out = sg_oracle.tp_lagrange(nodes, np.ones((3, 1)), np.array([[0.0], [0.3]]))
    assert np.isnan(out[0, 0]) and abs(out[1, 0] - 1) < 1e-14


def test_sample_recycling_matches_reference():
    g = load_golden("recycle")
    fvec = lambda y: np.array([np.sin(np.sum(y)), np.sum(y * y)])  # noqa: E731
    calls = []

    def f_count(y):
        calls.append(1)
        return fvec(y)
    s_big, rec = sg_oracle.sample_on_grid(g["big_SG"], f_count, g["small_SG"], g["s_small"])
    assert np.array_equal(s_big, g["s_big"]) and len(calls) == int(g["n_new"])
    back, rec = sg_oracle.sample_on_grid(g["mid3_SG"], lambda y: 1 / 0, g["big_SG"], g["s_big"])
    assert np.array_equal(back, g["back"]) and rec == g["mid3_SG"].shape[0]


def test_brownian_matches_reference():
    g = load_golden("brownian")
    for tag, tt, T in (("lc64", g["lc64_tt"], 1), ("lc5", g["lc5_tt"], 2.0),
                       ("lc1", g["lc5_tt"], 2.0)):
        got = np.array([brownian_oracle.lc_brownian(tt, y, T) for y in g[tag + "_yy"]])
        assert np.array_equal(got, g[tag + "_W"]), tag
    got = np.array([brownian_oracle.kl_brownian(g["kl_tt"], y) for y in g["kl_yy"]])
    assert np.array_equal(got, g["kl_W"])
    assert [brownian_oracle.level_lc(i) for i in range(6)] == \
        [(0, 1), (1, 1), (2, 0), (2, 1), (3, 0), (3, 1)]


# ---------------------------------------------------------------- composite drivers (a12, a14)
# Fixtures ml_*.npz / gg_*.npz are outputs of the reference's OWN driver code
# (multilevel_interpolant.py:89-150, compute_error_indicators.py:78-121) run on alias
# subclasses that carry the camelCase names it calls (tests/golden/make_golden.py).
def ml_oracle_plans(name):
    from golden_cases import ML_CASES, ml_mid_set
    from sgmethods_b200 import multi_index_sets as mis
    fam, l2k, kind, levels, NF, P, seed = ML_CASES[name]
    knots = knot_family(fam, nodes_1d)
    plans = [sg_oracle.make_plan(np.asarray(ml_mid_set(mis, name, k)), knots, LEV2KNOTS[l2k])
             for k in range(levels)]
    return plans, kind, levels


@pytest.mark.parametrize("name", ["ml_lin_nf1", "ml_quad_nf3", "ml_lin_grow"])
def test_multilevel_oracle_matches_reference_driver(name):
    from oracle import drivers_oracle
    g = load_golden(name)
    plans, kind, levels = ml_oracle_plans(name)
    for k, p in enumerate(plans):
        assert p.SG.tobytes() == g[f"SG_{k}"].tobytes()
    ml = [g[f"ml_{k}"] for k in range(levels)]
    out = drivers_oracle.ml_interpolate(plans, kind, g["yy"], ml)
    assert np.array_equal(np.squeeze(out), g["out"])
    for k, term in enumerate(drivers_oracle.ml_terms(plans, kind, g["yy"], ml)):
        assert np.array_equal(np.squeeze(term), g[f"term_{k}"]), k
    sampled = [g[f"sampled_{k}"] for k in range(levels)]
    out_s = drivers_oracle.ml_interpolate(plans, kind, g["yy"], sampled)
    assert np.array_equal(np.squeeze(out_s), g["out_sampled"])


@pytest.mark.parametrize("name", ["gg_lin_unb", "gg_quad_equi"])
def test_indicator_oracle_matches_reference_driver(name):
    from golden_cases import GG_CASES, GG_F, GG_NORM
    from oracle import drivers_oracle
    fam, l2k, kind, script, P, seed = GG_CASES[name]
    g = load_golden(name)
    knots = knot_family(fam, nodes_1d)
    for step in range(int(g["steps"])):
        calls = []

        def f_count(y):
            calls.append(1)
            return GG_F(y)
        est = drivers_oracle.gg_indicators(
            g[f"mid_{step}"], g[f"rm_{step}"], knots, LEV2KNOTS[l2k], f_count, g[f"SG_{step}"],
            g[f"u_on_SG_{step}"], g[f"yy_{step}"], GG_NORM, g[f"u_interp_{step}"], kind,
            old_RM=g[f"old_rm_{step}"], old_estimators=g[f"old_est_{step}"])
        assert np.array_equal(est, g[f"est_{step}"]), (step, est - g[f"est_{step}"])
        assert len(calls) == int(g[f"n_f_calls_{step}"])
