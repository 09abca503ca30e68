"""The contract path of ``SGInterpolant.interpolate``: ``method="auto"`` and the work-reduced
hierarchical-surplus form it selects -- needs a B200.

The north star's bar is max|gpu - ref| / max|ref| <= 1e-12 against the reference's
numpy/scipy path.  The hierarchical form is not bit-identical to the reference (it is the
same function evaluated without the large alternating combination sum), so every test here
states the tolerance: 1e-12, on the reference's own outputs (tests/golden), on the CPU oracle
(pinned to those outputs) and, for point counts the oracle cannot do, on the bit-exact CUDA
path (itself pinned in test_gpu_parity.py).  Shapes: BASELINE configs 2, 3 and 5, random and
smooth nodal values.
"""
import numpy as np
import pytest

from conftest import load_golden
from golden_cases import CASES, LEV2KNOTS, knot_family
from oracle import sg_oracle
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200 import sparse_grid_interpolant as sgi
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
from sgmethods_b200.utils import compute_level_lc

pytestmark = pytest.mark.gpu
TOL = 1.e-12                      # BASELINE.json north_star tolerance
L2K = lambda n: 2 ** (n + 1) - 1  # noqa: E731


def rel_dev(got, ref):
    got, ref = np.asarray(got, dtype=float), np.asarray(ref, dtype=float)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    scale = np.max(np.abs(ref[ok])) if ok.any() else 1.0
    return float(np.max(np.abs(got[ok] - ref[ok])) / (scale if scale > 0 else 1.0)) if ok.any() else 0.0


def oracle_plan_of(it):
    return sg_oracle.plan_from_tables(it.N, it.combination_coeffs, it.active_tp_dims,
                                      it.active_tp_nodes_list, it.map_tp_to_SG)


def smooth_values(it, NF=1):
    """sin(sum_n y_n / n^2) (the reference's example function, example/0 :27-29) and shifts."""
    w = 1.0 / np.arange(1, it.N + 1) ** 2
    s = it.SG @ w
    return np.stack([np.sin(s + 0.3 * c) for c in range(NF)], axis=1)


def config2(w=4):
    return SGInterpolant(mis.smolyak_mid_set(w, 16), nodes_1d.unbounded_nodes_nested, L2K, verbose=False)


def config3(w=8):
    a = np.array([compute_level_lc(n)[0] + 1 for n in range(64)], dtype=float)
    return SGInterpolant(mis.aniso_smolyak_mid_set(w, 64, a),
                         lambda n: nodes_1d.optimal_gaussian_nodes(n, p=2), L2K, verbose=False)


def config5():
    S = mis.aniso_smolyak_mid_set(3, 128, np.array([1.0] * 80 + [3.0] * 48))
    return SGInterpolant(S, nodes_1d.unbounded_nodes_nested, L2K, verbose=False)


# ------------------------------------------------------------ reference outputs (golden)
@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if c[2] == "linear"))
def test_hierarchical_form_on_reference_outputs(name):
    """Every pw-linear fixture: where the hierarchical form exists it reproduces the REFERENCE's
    output to 1e-12 (incl. NaN inputs and extrapolation); where it does not, it says so and
    method='auto' is the exact kernel."""
    fam, l2k, kind = CASES[name]
    g = load_golden(name)
    it = SGInterpolant(g["mid_set"], knot_family(fam, nodes_1d), LEV2KNOTS[l2k], verbose=False)
    try:
        out = it.interpolate(g["x"], g["f_on_SG"], method="hierarchical")
    except NotImplementedError:
        assert name in ("lin_deg1", "lin_exp_cc", "lin_example0")      # not nested / lev2knots(0) != 1
        out = it.interpolate(g["x"], g["f_on_SG"], method="auto")
        assert np.array_equal(out, it.interpolate(g["x"], g["f_on_SG"], method="combination"), equal_nan=True)
    assert rel_dev(out, g["out"]) <= TOL


# ------------------------------------------------------------ configs 2 / 3 / 5 shapes
@pytest.mark.parametrize("data", ["random", "smooth"])
def test_config2_shape_hierarchical_within_contract(data):
    it = config2()
    rng = np.random.default_rng(3)
    f = rng.normal(size=(it.num_nodes, 1)) if data == "random" else smooth_values(it)
    x = rng.normal(size=(60000, 16))
    x[:8] *= 4.0                                         # far outside the knots: extrapolation
    ref = sg_oracle.interpolate(oracle_plan_of(it), x[:48], f)
    fast = it.interpolate(x, f, method="hierarchical")
    exact = it.interpolate(x, f, method="combination")
    assert np.array_equal(exact[:48], ref)               # the exact path is the reference, bit for bit
    print(f"config-2 {data}: hierarchical vs reference arithmetic {rel_dev(fast, exact):.2e}")
    assert rel_dev(fast[:48], ref) <= TOL
    assert rel_dev(fast, exact) <= TOL


@pytest.mark.parametrize("data", ["random", "smooth"])
@pytest.mark.parametrize("NF", [1, 320])
def test_config3_shape_hierarchical_within_contract(data, NF):
    it = config3()
    rng = np.random.default_rng(4)
    f = rng.normal(size=(it.num_nodes, NF)) if data == "random" else smooth_values(it, NF)
    x = rng.normal(size=(9000 if NF == 1 else 4200, 64))
    ref = sg_oracle.interpolate(oracle_plan_of(it), x[:24], f, squeeze=False)
    fast = np.asarray(it.interpolate(x, f, method="hierarchical")).reshape(x.shape[0], NF)
    exact = np.asarray(it.interpolate(x, f, method="combination")).reshape(x.shape[0], NF)
    print(f"config-3 NF={NF} {data}: hierarchical vs reference arithmetic {rel_dev(fast, exact):.2e}")
    assert rel_dev(fast[:24], ref) <= TOL
    assert rel_dev(fast, exact) <= TOL


def test_config5_shape_auto_stays_within_contract_for_random_and_smooth_data():
    """d = 128, 91 929 terms, sum |c| = 6.7e5: where the reference's own cancellation error is
    largest.  With random nodal values the hierarchical form agrees with the reference
    arithmetic to ~1e-13 and method="auto" uses it; with smooth values the REFERENCE's
    rounding error is ~1e-11 (the form is the more accurate of the two, see the long-double
    test in test_gpu_parity.py), so the form is outside the 1e-12 contract, the per-call
    validation rejects it and "auto" returns the exact kernel's result.  The oracle does 4
    points (91 929 scipy calls each); 20 000 points against the bit-exact CUDA path."""
    it = config5()
    assert len(it.combination_coeffs) == 91929
    rng = np.random.default_rng(5)
    x = rng.normal(size=(20000, 128))
    plan = oracle_plan_of(it)
    f = rng.normal(size=(it.num_nodes, 1))
    exact = it.interpolate(x, f, method="combination")
    assert np.array_equal(exact[:4], sg_oracle.interpolate(plan, x[:4], f))
    auto = it.interpolate(x, f)
    print(f"config-5 random: {it.auto_validation}, auto vs reference arithmetic {rel_dev(auto, exact):.2e}")
    assert it.auto_validation["enabled"] and rel_dev(auto, exact) <= TOL
    fs = smooth_values(it)
    exact = it.interpolate(x, fs, method="combination")
    assert np.array_equal(exact[:4], sg_oracle.interpolate(plan, x[:4], fs))
    fast = it.interpolate(x, fs, method="hierarchical")
    auto = it.interpolate(x, fs)
    print(f"config-5 smooth: {it.auto_validation}, hierarchical vs reference arithmetic "
          f"{rel_dev(fast, exact):.2e}")
    assert not it.auto_validation["enabled"]
    assert np.array_equal(auto, exact)                   # the exact kernel answered
    assert rel_dev(fast, exact) <= 1e-10                 # same function; the gap is rounding


# ------------------------------------------------------------ method="auto"
def test_auto_is_the_default_and_validates_before_switching():
    assert sgi.DEFAULT_METHOD == "auto"
    it = config2(3)
    rng = np.random.default_rng(6)
    f = rng.normal(size=(it.num_nodes, 1))
    small = rng.normal(size=(500, 16))
    # small batches: the exact kernel, bit-identical to the reference arithmetic
    assert np.array_equal(it.interpolate(small, f), sg_oracle.interpolate(oracle_plan_of(it), small, f))
    assert getattr(it, "_auto_choice", None) is None
    big = rng.normal(size=(40000, 16))
    got = it.interpolate(big, f)                          # large batch: validated, then used
    assert it.auto_validation["enabled"] and it._auto_choice == "hierarchical"
    assert it.auto_validation["sample_points"] <= sgi.AUTO_SAMPLE
    assert it.auto_validation["max_rel_deviation"] <= sgi.AUTO_TOL
    assert np.array_equal(got, it.interpolate(big, f, method="hierarchical"))
    assert rel_dev(got, it.interpolate(big, f, method="combination")) <= TOL
    # small batches stay exact afterwards as well
    assert np.array_equal(it.interpolate(small, f), it.interpolate(small, f, method="combination"))
    with pytest.raises(ValueError):
        it.interpolate(small, f, method="fast")


def test_auto_keeps_the_exact_kernel_when_the_form_does_not_exist_or_fails_validation(monkeypatch):
    rng = np.random.default_rng(7)
    # knots that are not nested: no hierarchical form
    it = SGInterpolant(mis.smolyak_mid_set(3, 4), nodes_1d.cc_nodes, lambda n: n + 1, verbose=False)
    f = rng.normal(size=(it.num_nodes, 2))
    x = rng.uniform(-1, 1, (20000, 4))
    assert np.array_equal(it.interpolate(x, f), it.interpolate(x, f, method="combination"))
    assert it.auto_validation["enabled"] is False
    # pw-cubic: never (its patches are not nested, hierarchical.py)
    from sgmethods_b200.tp_inteprolants import TPPwCubicInterpolator
    itq = SGInterpolant(mis.smolyak_mid_set(2, 3), nodes_1d.cc_nodes, lambda n: 3 * 2 ** n + 1,
                        tp_interpolant=TPPwCubicInterpolator, verbose=False)
    fq = rng.normal(size=(itq.num_nodes, 1))
    xq = rng.uniform(-1, 1, (20000, 3))
    assert np.array_equal(itq.interpolate(xq, fq), itq.interpolate(xq, fq, method="combination"))
    assert getattr(itq, "_hier", None) is None
    # a validation that fails (tolerance forced to 0) leaves the exact kernel in place
    it2 = config2(2)
    monkeypatch.setattr(sgi, "AUTO_TOL", 0.0)
    f2 = rng.normal(size=(it2.num_nodes, 1))
    x2 = rng.normal(size=(20000, 16))
    out = it2.interpolate(x2, f2)
    assert it2._auto_choice == "combination" and not it2.auto_validation["enabled"]
    assert np.array_equal(out, it2.interpolate(x2, f2, method="combination"))


def test_auto_pipelined_host_inputs_and_torch_tensors():
    import torch
    it = config2(2)
    P = 4 * 113664 + 999                                  # takes the chunked copy/compute pipeline
    x = np.random.default_rng(8).normal(size=(P, 18))     # 2 unused columns
    f = np.random.default_rng(9).normal(size=(it.num_nodes, 1))
    got = it.interpolate(x, f)
    assert it._auto_choice == "hierarchical"
    ref = it.interpolate(torch.from_numpy(x).cuda(), torch.from_numpy(f).cuda(), method="hierarchical")
    assert isinstance(got, np.ndarray) and np.array_equal(got, ref.cpu().numpy())
    pinned = torch.from_numpy(x).pin_memory()
    assert torch.equal(it.interpolate(pinned, torch.from_numpy(f)), ref)
    exact = it.interpolate(torch.from_numpy(x).cuda(), torch.from_numpy(f).cuda(), method="combination")
    assert rel_dev(got, exact.cpu().numpy()) <= TOL


def test_hierarchical_nan_and_ragged_batches():
    it = config2(2)
    rng = np.random.default_rng(10)
    f = rng.normal(size=(it.num_nodes, 3))
    for P in (1, 31, 33, 1000, 4097):
        x = rng.normal(size=(P, 16))
        if P > 1:
            x[P // 2, 5] = np.nan
        fast = np.asarray(it.interpolate(x, f, method="hierarchical")).reshape(P, 3)
        exact = np.asarray(it.interpolate(x, f, method="combination")).reshape(P, 3)
        assert rel_dev(fast, exact) <= TOL


# ------------------------------------------------------------ pw-quadratic and Lagrange
DBL = lambda n: np.where(n == 0, 1, 2 ** n + 1)  # noqa: E731
NO_FORM = {"lag_example0": "lev2knots(0) = 3", "lag_hermite": "Gauss-Hermite knots are not nested"}


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if c[2] in ("quadratic", "lagrange")))
def test_hierarchical_form_of_quadratic_and_lagrange_on_reference_outputs(name):
    """The pw-quadratic and Lagrange fixtures: where the hierarchical form exists it reproduces
    the REFERENCE's output to 1e-12, otherwise it says so and the exact kernel answers."""
    from sgmethods_b200 import tp_inteprolants as tpi
    fam, l2k, kind = CASES[name]
    op = {"quadratic": tpi.TPPwQuadraticInterpolator, "lagrange": tpi.TPLagrangeInterpolator}[kind]
    g = load_golden(name)
    it = SGInterpolant(g["mid_set"], knot_family(fam, nodes_1d), LEV2KNOTS[l2k], tp_interpolant=op,
                       verbose=False)
    if name in NO_FORM:
        with pytest.raises(NotImplementedError):
            it.interpolate(g["x"], g["f_on_SG"], method="hierarchical")
        out = it.interpolate(g["x"], g["f_on_SG"])
    else:
        out = it.interpolate(g["x"], g["f_on_SG"], method="hierarchical")
        print(f"{name}: hierarchical vs reference output {rel_dev(out, g['out']):.2e}")
    assert rel_dev(out, g["out"]) <= TOL


def config4_level(k=3, d=8):
    from sgmethods_b200.tp_inteprolants import TPPwQuadraticInterpolator
    return SGInterpolant(mis.smolyak_mid_set(k, d), nodes_1d.cc_nodes, DBL,
                         tp_interpolant=TPPwQuadraticInterpolator, verbose=False)


@pytest.mark.parametrize("data", ["random", "smooth"])
@pytest.mark.parametrize("NF", [1, 200, 300])      # 300: four points per warp on 256-column tiles + ragged tile
def test_config4_shape_quadratic_hierarchical_within_contract(data, NF):
    """BASELINE config 4's finest level (pw-quadratic, d = 8, Smolyak level 3, Clenshaw-Curtis
    knots 1, 3, 5, 9): 333 basis tuples instead of 165 terms x 3^k nodes."""
    it = config4_level()
    rng = np.random.default_rng(11)
    f = rng.normal(size=(it.num_nodes, NF)) if data == "random" else smooth_values(it, NF)
    x = rng.uniform(-1.2, 1.0, (20000 if NF == 1 else 9000, 8))       # left of the first knot too
    ref = sg_oracle.interpolate(oracle_plan_of(it), x[:64], f, kind="quadratic", squeeze=False)
    fast = np.asarray(it.interpolate(x, f, method="hierarchical")).reshape(x.shape[0], NF)
    exact = np.asarray(it.interpolate(x, f, method="combination")).reshape(x.shape[0], NF)
    assert np.array_equal(exact[:64], ref)
    print(f"config-4 NF={NF} {data}: hierarchical vs reference arithmetic {rel_dev(fast, exact):.2e}")
    assert rel_dev(fast[:64], ref) <= TOL and rel_dev(fast, exact) <= TOL
    auto = np.asarray(it.interpolate(x, f)).reshape(x.shape[0], NF)
    assert it.auto_validation["enabled"] and np.array_equal(auto, fast)


def test_quadratic_hierarchical_raises_like_the_reference_right_of_the_last_knot():
    it = config4_level(2, 3)
    rng = np.random.default_rng(12)
    f = rng.normal(size=(it.num_nodes, 1))
    x = rng.uniform(-1, 1, (9000, 3))
    it.interpolate(x, f, method="hierarchical")
    x[4321, 1] = 1.0 + 1e-9                               # reference: IndexError (tp_inteprolants.py:131)
    for method in ("hierarchical", "auto", "combination"):
        with pytest.raises(IndexError):
            it.interpolate(x, f, method=method)
    x[4321, 1] = 1.0                                      # the last knot itself is fine
    assert rel_dev(it.interpolate(x, f, method="hierarchical"), it.interpolate(x, f, method="combination")) <= TOL


def leja_config2(w=4, d=16):
    from sgmethods_b200.tp_inteprolants import TPLagrangeInterpolator
    return SGInterpolant(mis.smolyak_mid_set(w, d), knot_family("leja", nodes_1d), lambda nu: nu + 1,
                         tp_interpolant=TPLagrangeInterpolator, verbose=False)


@pytest.mark.parametrize("data", ["random", "smooth"])
def test_config2_leja_lagrange_hierarchical_within_contract(data):
    """BASELINE config 2 as worded (Gaussian-Leja knots, d = 16, Lagrange): one Newton-like
    basis polynomial per level, 4845 instead of 58 905 node contributions per point."""
    it = leja_config2()
    rng = np.random.default_rng(13)
    f = rng.normal(size=(it.num_nodes, 1)) if data == "random" else smooth_values(it)
    x = rng.normal(size=(30000, 16))
    ref = sg_oracle.interpolate(oracle_plan_of(it), x[:32], f, kind="lagrange")
    fast = it.interpolate(x, f, method="hierarchical")
    exact = it.interpolate(x, f, method="combination")
    print(f"config-2 Leja {data}: hierarchical vs reference arithmetic {rel_dev(fast, exact):.2e}, "
          f"exact vs oracle {rel_dev(exact[:32], ref):.2e}")
    assert rel_dev(exact[:32], ref) <= TOL
    assert rel_dev(fast[:32], ref) <= TOL and rel_dev(fast, exact) <= TOL
    auto = it.interpolate(x, f)
    # smooth values: the reference's own cancellation error is ~1e-13 here, so the per-call
    # validation (2.5e-13 on its sample) may keep the exact kernel; either way inside 1e-12
    print(f"config-2 Leja {data}: {it.auto_validation}")
    if data == "random":
        assert it.auto_validation["enabled"]
    assert np.array_equal(auto, fast if it.auto_validation["enabled"] else exact)


def test_lagrange_hierarchical_nan_on_knots_like_the_reference():
    """A coordinate equal to a knot of an active direction: NaN in the reference
    (tp_inteprolants.py:346, 0/0), NaN here, and method='auto' still validates (NaN == NaN)."""
    it = leja_config2(3, 6)
    rng = np.random.default_rng(14)
    f = rng.normal(size=(it.num_nodes, 2))
    x = rng.normal(size=(10000, 6))
    knots = knot_family("leja", nodes_1d)(4)
    x[0, 2], x[17, 5], x[9000, 0] = 0.0, knots[-1], knots[1]
    ref = sg_oracle.interpolate(oracle_plan_of(it), x[:32], f, kind="lagrange", squeeze=False)
    assert np.isnan(ref[0]).all() and np.isnan(ref[17]).all() and not np.isnan(ref[1]).any()
    fast = it.interpolate(x, f, method="hierarchical")
    exact = it.interpolate(x, f, method="combination")
    assert np.isnan(fast[9000]).all() and np.isnan(exact[9000]).all()
    assert rel_dev(fast[:32], ref) <= TOL and rel_dev(fast, exact) <= TOL      # rel_dev compares the NaN masks
    xs = x.copy()
    xs[::39, 3] = knots[2]                                # the validation sample sees NaNs as well
    auto = it.interpolate(xs, f)
    assert it.auto_validation["enabled"]
    assert rel_dev(auto, it.interpolate(xs, f, method="combination")) <= TOL


# ------------------------------------------------------------ multilevel (config 4)
@pytest.mark.parametrize("name", ["ml_lin_nf1", "ml_quad_nf3", "ml_lin_grow"])
def test_multilevel_hierarchical_form_on_reference_driver_outputs(name):
    """MLInterpolant.interpolate in hierarchical-surplus form against the outputs of the
    reference's own driver code (tests/golden/ml_*.npz), 1e-12."""
    from golden_cases import ML_CASES, ml_mid_set
    from sgmethods_b200 import tp_inteprolants as tpi
    from sgmethods_b200.multilevel_interpolant import MLInterpolant
    fam, l2k, kind, levels, NF, P, seed = ML_CASES[name]
    op = {"linear": tpi.TPPwLinearInterpolator, "quadratic": tpi.TPPwQuadraticInterpolator}[kind]
    g = load_golden(name)
    seq = [SGInterpolant(np.asarray(ml_mid_set(mis, name, k)), knot_family(fam, nodes_1d),
                         LEV2KNOTS[l2k], tp_interpolant=op, verbose=False) for k in range(levels)]
    mli = MLInterpolant(seq)
    ml = [g[f"ml_{k}"] for k in range(levels)]
    out = mli.interpolate(g["yy"], ml, method="hierarchical")
    assert out.shape == g["out"].shape
    print(f"{name}: hierarchical multilevel vs reference driver output {rel_dev(out, g['out']):.2e}")
    assert rel_dev(out, g["out"]) <= TOL
    assert rel_dev(mli.interpolate(g["yy"], ml), g["out"]) <= TOL          # small batch: exact launch
    assert rel_dev(mli.interpolate(g["yy"], ml, method="combination"), g["out"]) <= TOL


def test_config4_shape_multilevel_auto_within_contract():
    """BASELINE config 4: 4 levels, pw-quadratic, d = 8.  method='auto' evaluates ONE hierarchical
    plan of 705 basis tuples whose surpluses come from the four sample blocks."""
    from oracle import drivers_oracle
    from sgmethods_b200.multilevel_interpolant import MLInterpolant
    seq = [config4_level(k) for k in range(4)]
    mli = MLInterpolant(seq)
    rng = np.random.default_rng(15)
    NF = 40
    w = 1.0 / np.arange(1, 9) ** 2
    ml = [np.stack([np.sin(seq[3 - k].SG @ w + 0.3 * c) + 2.0 ** (-2 * k) * np.cos(seq[3 - k].SG @ w * (c + 1))
                    for c in range(NF)], axis=1) for k in range(4)]
    yy = rng.uniform(-1.1, 1.0, (12000, 8))
    plans = [oracle_plan_of(it) for it in seq]
    for p, it in zip(plans, seq):
        p.SG = it.SG
    ref = drivers_oracle.ml_interpolate(plans, "quadratic", yy[:48], ml)
    exact = mli.interpolate(yy, ml, method="combination")
    auto = mli.interpolate(yy, ml)
    print(f"config-4 multilevel: {mli.auto_validation}, auto vs exact launch {rel_dev(auto, exact):.2e}")
    assert mli.auto_validation["enabled"]
    assert rel_dev(exact[:48], np.squeeze(ref)) <= TOL and rel_dev(auto[:48], np.squeeze(ref)) <= TOL
    assert rel_dev(auto, exact) <= TOL
    assert np.array_equal(auto, mli.interpolate(yy, ml, method="hierarchical"))
    yy[77, 3] = 1.0 + 1e-9
    for method in ("auto", "hierarchical", "combination"):
        with pytest.raises(IndexError):
            mli.interpolate(yy, ml, method=method)


# ------------------------------------------------------------ random downward closed sets
@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("kind", ["linear", "quadratic", "lagrange"])
def test_random_downward_closed_sets_hierarchical_kernels(kind, seed):
    """Random (not Smolyak-shaped) downward closed sets, scalar and vector-valued: grouped
    hierarchical kernel / hierarchical column kernels against the oracle (first rows) and the
    exact CUDA path (all rows), 1e-12."""
    from test_hierarchical_cpu import _random_downward_closed_set
    from sgmethods_b200 import tp_inteprolants as tpi
    rng = np.random.default_rng(200 + seed)
    N = int(rng.integers(2, 7))
    S = _random_downward_closed_set(rng, N, 4 if kind != "lagrange" else 3, int(rng.integers(10, 60)))
    if kind == "linear":
        op, knots, lev = tpi.TPPwLinearInterpolator, nodes_1d.unbounded_nodes_nested, L2K
        x = rng.normal(0, 1.5, (9000, N))
    elif kind == "quadratic":
        op, knots, lev = tpi.TPPwQuadraticInterpolator, nodes_1d.cc_nodes, DBL
        x = rng.uniform(-1.2, 1.0, (9000, N))
    else:
        op, knots, lev = tpi.TPLagrangeInterpolator, nodes_1d.cc_nodes, DBL
        x = rng.uniform(-1.0, 1.0, (9000, N))
    it = SGInterpolant(S, knots, lev, tp_interpolant=op, verbose=False)
    for NF in (1, 130):
        f = rng.normal(size=(it.num_nodes, NF))
        ref = sg_oracle.interpolate(oracle_plan_of(it), x[:40], f, kind=kind, squeeze=False)
        fast = np.asarray(it.interpolate(x, f, method="hierarchical")).reshape(-1, NF)
        exact = np.asarray(it.interpolate(x, f, method="combination")).reshape(-1, NF)
        assert rel_dev(exact[:40], ref) <= TOL
        assert rel_dev(fast[:40], ref) <= TOL and rel_dev(fast, exact) <= TOL


def test_hierarchical_column_kernel_flavours_agree_bit_for_bit(monkeypatch):
    """One, two and four points per warp (128- and 256-column tiles) of the hierarchical column
    kernels and the bulk-copy experiment kernel compute the same bits."""
    import torch
    it = config4_level(3, 6)
    rng = np.random.default_rng(16)
    hp = it.hierarchical()
    plan = hp._dev(0)[0]
    x = torch.from_numpy(rng.uniform(-1.1, 1.0, (5003, 6))).cuda()
    f = torch.from_numpy(rng.normal(size=(it.num_nodes, 610))).cuda()      # 2 full tiles + ragged
    alpha = hp.surpluses(f)
    ref = None
    for env in ({"SGM_WIDE_PTS": "2"}, {"SGM_WIDE_PTS": "4"}, {"SGM_WIDE_PTS": "8"}, {}, {"SGM_WIDE_BULK": "1"}):
        for k in ("SGM_WIDE_PTS", "SGM_WIDE_BULK"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        plan.reload_knobs()
        out = plan.eval(x, alpha)
        ref = out if ref is None else ref
        assert torch.equal(out, ref), env
    for k in ("SGM_WIDE_PTS", "SGM_WIDE_BULK"):
        monkeypatch.delenv(k, raising=False)
    plan.reload_knobs()
    exact = it.interpolate(x, f, method="combination")
    assert rel_dev(ref.cpu().numpy(), exact.cpu().numpy()) <= TOL


def test_max_rel_deviation_kernel_nan_rules():
    """sgm_max_rel_deviation (the validation's comparison): NaN in both arrays is agreement, NaN
    in one makes the result NaN, max |b| = 0 divides by 1; sizes across the block reduction."""
    import torch
    rng = np.random.default_rng(17)
    for n in (1, 31, 1024, 70001):
        a = rng.normal(size=n)
        b = a + 1e-9 * rng.normal(size=n)
        want = np.max(np.abs(a - b)) / np.max(np.abs(b))
        got = float(sgi._max_rel_deviation(torch, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()))
        assert abs(got - want) <= 1e-15 * max(1.0, want) + 1e-25
        if n > 1:
            a2, b2 = a.copy(), b.copy()
            a2[n // 2] = b2[n // 2] = np.nan                      # NaN in both: ignored
            got = float(sgi._max_rel_deviation(torch, torch.from_numpy(a2).cuda(), torch.from_numpy(b2).cuda()))
            keep = np.arange(n) != n // 2
            assert abs(got - np.max(np.abs(a - b)[keep]) / np.max(np.abs(b[keep]))) <= 1e-15
            b2[n // 2] = 1.0                                      # NaN in one: NaN
            assert np.isnan(float(sgi._max_rel_deviation(torch, torch.from_numpy(a2).cuda(),
                                                         torch.from_numpy(b2).cuda())))
    z = torch.zeros(5, dtype=torch.float64, device="cuda")
    assert float(sgi._max_rel_deviation(torch, z + 0.25, z)) == 0.25


@pytest.mark.parametrize("kind", ["linear", "quadratic", "lagrange"])
def test_few_tuple_plans_at_many_points(kind):
    """Regression (found by tools/fuzz_auto.py): a hierarchical plan with fewer than 32 basis
    tuples evaluated at >= 16384 points took the one-thread-per-point kernel while its value
    table had been gathered in the grouped kernel's layout -- wrong values, silently."""
    from sgmethods_b200 import tp_inteprolants as tpi
    rng = np.random.default_rng(18)
    S = mis.smolyak_mid_set(2, 3)                         # 10 multi-indices
    if kind == "linear":
        it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, L2K, verbose=False)
        x = rng.normal(0, 1.2, (20000, 3))
    else:
        op = tpi.TPPwQuadraticInterpolator if kind == "quadratic" else tpi.TPLagrangeInterpolator
        it = SGInterpolant(S, nodes_1d.cc_nodes, DBL, tp_interpolant=op, verbose=False)
        x = rng.uniform(-1.0, 1.0, (20000, 3))
    for NF in (1, 17):
        f = rng.normal(size=(it.num_nodes, NF))
        ref = sg_oracle.interpolate(oracle_plan_of(it), x[:64], f, kind=kind, squeeze=False)
        exact = np.asarray(it.interpolate(x, f, method="combination")).reshape(-1, NF)
        for method in ("auto", "hierarchical"):
            got = np.asarray(it.interpolate(x, f, method=method)).reshape(-1, NF)
            assert rel_dev(got[:64], ref) <= TOL and rel_dev(got, exact) <= TOL, (method, NF)


def test_auto_keeps_exact_kernels_when_the_hierarchical_tables_do_not_fit():
    """Very deep 1-D levels: the knot tables of the hierarchical plan (16 bytes per knot) exceed
    shared memory while the exact plan (8 bytes per knot) still fits -- method='auto' must fall
    back instead of failing; method='hierarchical' reports the limit."""
    S = np.arange(13, dtype=int).reshape(-1, 1)           # levels 0..12: 16 382 knots in all
    it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, L2K, verbose=False)
    rng = np.random.default_rng(19)
    f = rng.normal(size=(it.num_nodes, 1))
    x = rng.normal(0, 1.5, (9000, 1))
    exact = it.interpolate(x, f, method="combination")
    assert np.array_equal(exact[:50], sg_oracle.interpolate(oracle_plan_of(it), x[:50], f))
    auto = it.interpolate(x, f)
    assert it.auto_validation["enabled"] is False and "reason" in it.auto_validation
    assert np.array_equal(auto, exact)
    with pytest.raises(OverflowError):
        it.interpolate(x, f, method="hierarchical")
