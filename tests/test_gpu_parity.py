"""Parity of the CUDA path (through the C ABI) with the reference -- needs a B200.

Sources of truth, in this order:
  1. tests/golden/*.npz: outputs of the unmodified reference on seeded inputs;
  2. the CPU oracle (oracle/), itself pinned to (1), on fresh seeded inputs;
  3. size-independent properties at BASELINE sizes (interpolation at the nodes, partition of
     unity, linearity in the nodal values).
Tolerance (BASELINE.json north_star): max|gpu - ref| / max|ref| <= 1e-12 in fp64.  The
kernels reproduce the reference's operation order, so most results are bit-identical; the
bit-exact fraction is printed for information.
"""
from math import pi

import numpy as np
import pytest

from conftest import load_golden
from golden_cases import CASES, LEV2KNOTS, knot_family
from oracle import brownian_oracle, sg_oracle
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200._device import reload_knobs
from sgmethods_b200.parametric_expansions_brownian_motion import (param_KL_Brownian_motion,
                                                                   param_LC_Brownian_motion)
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
from sgmethods_b200.tp_inteprolants import (TPLagrangeInterpolator, TPPwCubicInterpolator,
                                            TPPwLinearInterpolator, TPPwQuadraticInterpolator)

pytestmark = pytest.mark.gpu
TOL = 1.e-12


@pytest.fixture(autouse=True)
def exact_combination_formula_by_default(monkeypatch):
    """This module pins the EXACT path (bit-identical results, every kernel flavour); the
    validated work-reduced default (method="auto") has its own module, test_gpu_auto.py."""
    import sgmethods_b200.sparse_grid_interpolant as sgi
    monkeypatch.setattr(sgi, "DEFAULT_METHOD", "combination")
KIND_CLASS = {"linear": TPPwLinearInterpolator, "quadratic": TPPwQuadraticInterpolator,
              "cubic": TPPwCubicInterpolator, "lagrange": TPLagrangeInterpolator}


def rel_err(got, ref):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    if not ok.any():
        return 0.0
    scale = np.max(np.abs(ref[ok]))
    return float(np.max(np.abs(got[ok] - ref[ok])) / (scale if scale > 0 else 1.0))


def oracle_plan_of(it):
    return sg_oracle.plan_from_tables(it.N, it.combination_coeffs, it.active_tp_dims,
                                      it.active_tp_nodes_list, it.map_tp_to_SG)


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_reference_outputs(name):
    fam, l2k, kind = CASES[name]
    g = load_golden(name)
    it = SGInterpolant(g["mid_set"], knot_family(fam, nodes_1d), LEV2KNOTS[l2k],
                       tp_interpolant=KIND_CLASS[kind], verbose=False)
    out = it.interpolate(g["x"], g["f_on_SG"])
    err = rel_err(out, g["out"])
    ok = ~np.isnan(g["out"])
    exact = np.mean(out[ok] == g["out"][ok])
    print(f"{name}: rel err {err:.2e}, bit-exact {100 * exact:.1f}%")
    assert err <= TOL


def test_lambda_factory_and_sampling_like_reference_test():
    """tests/test_sparse_grid_interpolant.py:60-106 (operator passed as a lambda factory)."""
    N = 4
    F = lambda x: np.exp(np.sum(x))  # noqa: E731
    np.random.seed(0)
    pts = np.random.uniform(-1, 1, [5000, N])
    f_rnd = np.array(list(map(F, pts)))
    errs, nodes = [], []
    for w in range(7):
        it = SGInterpolant(mis.smolyak_mid_set(w, N=N), nodes_1d.cc_nodes, lambda nu: nu + 1,
                           tp_interpolant=lambda n, f: TPPwLinearInterpolator(n, f),
                           n_parallel=1, verbose=False)
        f_on_SG = it.sample_on_SG(F)
        val = it.interpolate(pts, f_on_SG)
        errs.append(np.amax(np.abs(val - f_rnd)))
        nodes.append(it.num_nodes)
        if w == 3:      # same arithmetic as the CPU path
            ref = sg_oracle.interpolate(oracle_plan_of(it), pts[:512], f_on_SG)
            assert rel_err(val[:512], ref) <= TOL
    errs, nodes = np.array(errs), np.array(nodes, dtype=float)
    assert np.all(np.diff(nodes) > 0) and np.all(np.diff(errs) < 0)
    assert np.all(errs < 31 * nodes ** 0.02 * np.exp(-0.005 * nodes))


def test_exactness_tests_of_the_reference():
    """tests/test_sparse_grid_interpolant.py:15-57: degree-1 (linear) and degree-3 (cubic)."""
    P1 = lambda x: 2. + x[0] + x[0] * x[1]  # noqa: E731
    Lam = np.array([[0, 0], [0, 1], [1, 0], [1, 1]])
    it = SGInterpolant(Lam, nodes_1d.equispaced_nodes, lambda n: n + 1, verbose=False)
    x = np.random.default_rng(0).uniform(-1, 1, [10, 2])
    got = it.interpolate(x, it.sample_on_SG(P1))
    assert np.linalg.norm(got - np.array([P1(r) for r in x])) < 1.e-4
    P3 = lambda x: 2. + x[0] + x[0] * x[1] + x[1] ** 3 + x[0] * x[1] ** 3  # noqa: E731
    Lam = np.array([[0, 0], [0, 1], [0, 2], [0, 3], [1, 0], [1, 1], [1, 2], [1, 3]])
    it = SGInterpolant(Lam, nodes_1d.equispaced_nodes, lambda n: 3 * n + 1,
                       tp_interpolant=TPPwCubicInterpolator, verbose=False)
    got = it.interpolate(x, it.sample_on_SG(P3))
    assert np.linalg.norm(got - np.array([P3(r) for r in x])) < 1.e-4


@pytest.mark.parametrize("kind,NF", [("linear", 1), ("linear", 7), ("linear", 200),
                                     ("quadratic", 1), ("quadratic", 40), ("cubic", 1),
                                     ("cubic", 33), ("lagrange", 1), ("lagrange", 130)])
def test_fresh_inputs_against_oracle(kind, NF):
    rng = np.random.default_rng(100 + NF)
    if kind == "cubic":
        S, knots, lev = mis.smolyak_mid_set(3, 4), nodes_1d.equispaced_nodes, (lambda n: 3 * n + 1)
        x = rng.uniform(-1.4, 1.4, (300, 4))
    elif kind == "lagrange":
        S, knots, lev = mis.smolyak_mid_set(3, 4), nodes_1d.hermite_nodes, (lambda n: 2 * n + 1)
        x = rng.normal(0, 1, (300, 4))
    elif kind == "quadratic":
        S, knots = mis.aniso_smolyak_mid_set(4, 5, [1, 1, 1.5, 2, 2]), nodes_1d.cc_nodes
        lev = lambda n: np.where(n == 0, 1, 2 ** n + 1)  # noqa: E731
        x = rng.uniform(-1, 1, (300, 5))
    else:
        S, knots, lev = mis.smolyak_mid_set(4, 6), nodes_1d.unbounded_nodes_nested, (lambda n: 2 ** (n + 1) - 1)
        x = rng.normal(0, 1.3, (300, 6))
    it = SGInterpolant(S, knots, lev, tp_interpolant=KIND_CLASS[kind], verbose=False)
    f = rng.normal(size=(it.num_nodes, NF))
    ref = sg_oracle.interpolate(oracle_plan_of(it), x, f, kind, squeeze=False)
    got = it.interpolate(x, f).reshape(ref.shape)
    err = rel_err(got, ref)
    print(f"{kind} NF={NF}: T={len(it.combination_coeffs)} M={it.num_nodes} rel err {err:.2e} "
          f"bit-exact {100 * np.mean(got == ref):.1f}%")
    assert err <= TOL


def test_smooth_function_with_large_coefficients():
    """Cancellation check: smooth f, |c| up to 1e3; the GPU result must be
    within 1e-12 of the CPU arithmetic although sum|c| ~ 4e4."""
    S = mis.smolyak_mid_set(4, 16)
    it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1,
                       verbose=False)
    assert len(it.combination_coeffs) == 4845 and it.num_nodes == 69121
    w = 1.0 / np.arange(1, 17) ** 2
    f = np.sin(it.SG @ w).reshape(-1, 1)
    x = np.random.default_rng(0).normal(size=(1024, 16))
    ref = sg_oracle.interpolate(oracle_plan_of(it), x, f)
    got = it.interpolate(x, f)
    err = rel_err(got, ref)
    print(f"C2 shape smooth f: rel err {err:.2e}, bit-exact {100 * np.mean(got == ref):.1f}%")
    assert err <= TOL


def test_properties_at_config2_size():
    """BASELINE config 2 (d=16, smolyak w=4, 10^6 points): properties that need no CPU run."""
    import torch
    S = mis.smolyak_mid_set(4, 16)
    it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1,
                       verbose=False)
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn((10 ** 6, 16), dtype=torch.float64, device=dev, generator=gen)
    f1 = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
    f2 = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
    o1, o2 = it.interpolate(x, f1), it.interpolate(x, f2)
    o12 = it.interpolate(x, 2.0 * f1 - 3.0 * f2)
    scale = float(o12.abs().max())
    assert float((o12 - (2.0 * o1 - 3.0 * o2)).abs().max()) <= 1e-11 * scale   # linearity
    ones = it.interpolate(x, torch.ones_like(f1))
    assert float((ones - 1.0).abs().max()) <= 1e-9                             # sum c_t = 1
    # nested knots: the interpolant reproduces the nodal values at the sparse-grid nodes
    nodes = torch.from_numpy(it.SG).to(dev)
    at_nodes = it.interpolate(nodes, f1)
    assert float((at_nodes - f1[:, 0]).abs().max()) <= 1e-9 * float(f1.abs().max()) * 1e3
    # determinism
    assert torch.equal(o1, it.interpolate(x, f1))


def test_config5_shape_d128_1e5_terms_against_oracle_and_properties():
    """BASELINE config 5 shape: d=128, 91 929 combination terms, 721 697 nodes (coefficients
    between -79127 and 3081), scalar output.  A few points bit for bit against the oracle's
    term loop; partition of unity and reproduction of the nodal values on a larger batch."""
    import torch
    S = mis.aniso_smolyak_mid_set(3, 128, np.array([1.0] * 80 + [3.0] * 48))
    it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1,
                       verbose=False)
    assert len(it.combination_coeffs) == 91929 and it.num_nodes == 721697
    rng = np.random.default_rng(31)
    f = rng.normal(size=(it.num_nodes, 1))
    x = rng.normal(size=(40000, 128))
    got = it.interpolate(x, f)
    ref = sg_oracle.interpolate(oracle_plan_of(it), x[:6], f)
    assert np.array_equal(got[:6], ref)
    assert np.all(np.isfinite(got))
    ones = it.interpolate(x, np.ones_like(f))
    assert np.max(np.abs(ones - 1.0)) <= 1e-8                       # sum_t c_t = 1
    pick = rng.choice(it.num_nodes, 40000, replace=False)
    at_nodes = it.interpolate(it.SG[pick], f)
    assert np.max(np.abs(at_nodes - f[pick, 0])) <= 1e-7 * np.max(np.abs(f))


def test_config3_shape_d64_vector_valued_against_oracle():
    """BASELINE config 3 shape: Levy-Ciesielski anisotropy in d=64 (802 terms, 16 033 nodes),
    NF = 10^4 output columns: a few points against the oracle on every column, partition of
    unity on a larger batch."""
    import torch
    from sgmethods_b200.utils import compute_level_lc
    a = np.array([compute_level_lc(n)[0] + 1 for n in range(64)], dtype=float)
    it = SGInterpolant(mis.aniso_smolyak_mid_set(10, 64, a),
                       lambda n: nodes_1d.optimal_gaussian_nodes(n, p=2),
                       lambda n: 2 ** (n + 1) - 1, verbose=False)
    assert len(it.combination_coeffs) == 802 and it.num_nodes == 16033
    NF = 10 ** 4
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(5)
    f = torch.randn((it.num_nodes, NF), dtype=torch.float64, device=dev, generator=gen)
    x = torch.randn((3000, 64), dtype=torch.float64, device=dev, generator=gen)
    got = it.interpolate(x, f)
    ref = sg_oracle.interpolate(oracle_plan_of(it), x[:12].cpu().numpy(), f.cpu().numpy())
    assert rel_err(got[:12].cpu().numpy(), ref) <= TOL
    print(f"config-3 shape: bit-exact {100 * np.mean(got[:12].cpu().numpy() == ref):.1f}%")
    ones = it.interpolate(x, torch.ones((it.num_nodes, 256), dtype=torch.float64, device=dev))
    assert float((ones - 1.0).abs().max()) <= 1e-9


def test_config4_shape_multilevel_quadratic_nf1000_against_oracle():
    """BASELINE config 4 shape: 4-level multilevel interpolant, piecewise quadratic, d=8,
    NF = 10^3 (levels k = 0..3 of the Smolyak sequence on doubling Clenshaw-Curtis knots)."""
    from oracle import drivers_oracle
    from sgmethods_b200.multilevel_interpolant import MLInterpolant
    dbl = lambda n: np.where(n == 0, 1, 2 ** n + 1)
    seq = [SGInterpolant(mis.smolyak_mid_set(k, 8), nodes_1d.cc_nodes, dbl,
                         tp_interpolant=TPPwQuadraticInterpolator, verbose=False) for k in range(4)]
    rng = np.random.default_rng(41)
    NF = 1000
    ml = [rng.normal(size=(seq[3 - k].num_nodes, NF)) for k in range(4)]
    plans = [oracle_plan_of(it) for it in seq]
    for p, it in zip(plans, seq):
        p.SG = it.SG
    # 2500 points: one point per warp; 4500 points: the accumulating (+I_k, -I_{k-1}) launches go
    # through the two-points-per-warp kernel with its guarded last column tile
    for P in (2500, 4500):
        yy = rng.uniform(-1, 1, (P, 8))
        got = MLInterpolant(seq).interpolate(yy, ml)
        pick = np.concatenate([np.arange(48), [P - 2, P - 1]])
        ref = drivers_oracle.ml_interpolate(plans, "quadratic", yy[pick], ml)
        assert rel_err(got[pick], ref) <= TOL


def test_torch_tensors_in_torch_tensor_out_and_extra_columns():
    import torch
    g = load_golden("lin_example1")
    fam, l2k, kind = CASES["lin_example1"]
    it = SGInterpolant(g["mid_set"], knot_family(fam, nodes_1d), LEV2KNOTS[l2k], verbose=False)
    x = torch.from_numpy(g["x"]).cuda()                       # 1000 columns, 10 used
    out = it.interpolate(x, torch.from_numpy(g["f_on_SG"]).cuda())
    assert out.is_cuda and out.shape == (32,)
    assert rel_err(out.cpu().numpy(), g["out"]) <= TOL
    with pytest.raises(IndexError):
        it.interpolate(g["x"][:, :1], g["f_on_SG"])            # too few columns


def test_quadratic_right_of_last_knot_raises_index_error():
    it = SGInterpolant(mis.smolyak_mid_set(2, 2), nodes_1d.cc_nodes,
                       lambda n: np.where(n == 0, 1, 2 ** n + 1),
                       tp_interpolant=TPPwQuadraticInterpolator, verbose=False)
    f = np.ones((it.num_nodes, 1))
    assert np.allclose(it.interpolate(np.array([[0.2, -1.5], [1.0, 1.0]]), f), 1.0)
    with pytest.raises(IndexError):
        it.interpolate(np.array([[0.2, 1.0000001]]), f)
    with pytest.raises(IndexError):
        it.interpolate(np.array([[np.nan, 0.0]]), f)
    assert np.allclose(it.interpolate(np.array([[0.2, 0.3]]), f), 1.0)   # flag was cleared


def test_cubic_with_three_knots_raises_like_reference():
    it = SGInterpolant(mis.smolyak_mid_set(1, 2), nodes_1d.equispaced_nodes,
                       lambda n: 2 * n + 1, tp_interpolant=TPPwCubicInterpolator, verbose=False)
    with pytest.raises(UnboundLocalError):
        it.interpolate(np.zeros((2, 2)), np.ones((it.num_nodes, 1)))


def test_lagrange_on_a_knot_is_nan_like_reference():
    knots = lambda n: np.atleast_1d(np.asarray(nodes_1d.equispaced_nodes(n), dtype=float))  # noqa
    it = SGInterpolant(mis.smolyak_mid_set(2, 2), knots, lambda n: 2 * n + 1,
                       tp_interpolant=TPLagrangeInterpolator, verbose=False)
    out = it.interpolate(np.array([[0.0, 0.3], [0.1, 0.3]]), np.ones((it.num_nodes, 1)))
    assert np.isnan(out[0]) and abs(out[1] - 1.0) < 1e-12


@pytest.mark.parametrize("cls,op", [(TPPwLinearInterpolator, "linear"),
                                    (TPPwQuadraticInterpolator, "quadratic"),
                                    (TPPwCubicInterpolator, "cubic"),
                                    (TPLagrangeInterpolator, "lagrange")])
def test_standalone_operators_follow_the_plugin_protocol(cls, op):
    """cls(nodes_tuple, f_on_nodes)(x_new) -> (P, dim_f) (tp_inteprolants.py:1-11), with a
    different knot vector per direction and the same node count in two directions."""
    rng = np.random.default_rng(11)
    nodes = (np.linspace(-1, 1, 7), np.sort(rng.uniform(-1, 1, 7)), np.linspace(-1, 1, 4))
    if op == "quadratic":
        nodes = nodes[:2] + (np.linspace(-1, 1, 5),)
    vals = rng.normal(size=tuple(n.size for n in nodes) + (3,))
    x = rng.uniform(-1, 1, (200, 3))
    if op == "quadratic":
        x[:, 1] = np.minimum(x[:, 1], nodes[1][-1])
    got = cls(nodes, vals)(x)
    ref = sg_oracle.OPERATORS[op](nodes, vals, x)
    assert got.shape == (200, 3) and rel_err(got, ref) <= TOL
    scalar = cls(nodes, vals[..., 0])(x)                       # scalar field: (P, 1)
    assert scalar.shape == (200, 1) and rel_err(scalar[:, 0], ref[:, 0]) <= TOL


def test_levy_ciesielski_bit_exact_against_reference_outputs():
    g = load_golden("brownian")
    W = param_LC_Brownian_motion(g["lc64_tt"], g["lc64_yy"], 1)
    assert W.shape == g["lc64_W"].shape and np.array_equal(W, g["lc64_W"])
    W = param_LC_Brownian_motion(g["lc5_tt"], g["lc5_yy"], 2.0)
    assert np.array_equal(W, g["lc5_W"])
    W = param_LC_Brownian_motion(g["lc5_tt"], g["lc1_yy"], 2.0)
    assert np.array_equal(W, g["lc1_W"])
    one = param_LC_Brownian_motion(g["lc64_tt"], g["lc64_yy"][0], 1)      # 1-D in, 1-D out
    assert one.shape == (1001,) and np.array_equal(one, g["lc64_W"][0])
    with pytest.warns(UserWarning):
        param_LC_Brownian_motion(np.array([0.0, 1.5]), g["lc5_yy"][0], 1.0)


def test_levy_ciesielski_large_batch_against_oracle():
    rng = np.random.default_rng(2)
    tt = np.linspace(0, 1, 2000)
    yy = rng.normal(size=(500, 64))
    W = param_LC_Brownian_motion(tt, yy, 1.0)
    for p in (0, 17, 499):
        assert np.array_equal(W[p], brownian_oracle.lc_brownian(tt, yy[p], 1.0))


@pytest.mark.parametrize("d", [3, 5, 16, 32, 48, 64, 100, 150, 200])
def test_levy_ciesielski_ragged_shapes_against_oracle(d):
    """Staging paths of the tiled kernel: coefficient counts that do not divide the block size
    (row/column carry of the cp.async copy), more than 6 levels, shared memory above 48 KB
    (d = 150), the generic kernel (d = 200), several row groups with a ragged last group,
    a strided view of the parameter vectors, a time grid that is not a multiple of the tile."""
    import torch
    rng = np.random.default_rng(300 + d)
    P, nt = 77, 700
    tt = np.sort(rng.uniform(0, 2.0, nt))
    tt[0], tt[-1] = 0.0, 2.0
    big = rng.normal(size=(P, d + 3))
    y_view = torch.from_numpy(big).cuda()[:, 1:d + 1]           # leading dimension d + 3
    from sgmethods_b200.parametric_expansions_brownian_motion import lc_brownian_device
    W = lc_brownian_device(torch.from_numpy(tt).cuda(), y_view, 2.0).cpu().numpy()
    assert W.shape == (P, nt) and not y_view.is_contiguous()
    for p in (0, 31, 32, 63, 64, 76):
        assert np.array_equal(W[p], brownian_oracle.lc_brownian(tt, big[p, 1:d + 1], 2.0)), (d, p)


def test_karhunen_loeve_against_reference_outputs():
    g = load_golden("brownian")
    W = param_KL_Brownian_motion(g["kl_tt"], g["kl_yy"])
    assert rel_err(W, g["kl_W"]) <= TOL


def test_karhunen_loeve_tiled_kernel_equals_per_output_kernel():
    """>= 64 parameter vectors go through the tiled kernel (sines once per block and time,
    rows streamed through shared memory, two rows per thread): bit-identical to the
    one-thread-per-output kernel that small batches use; ragged rows / times, d not a
    divisor of the block size."""
    rng = np.random.default_rng(9)
    for d, P, nt in ((64, 203, 130), (5, 77, 64), (100, 64, 1)):
        tt = np.sort(rng.uniform(0, 1, nt))
        yy = rng.normal(size=(P, d))
        W = param_KL_Brownian_motion(tt, yy)
        assert W.shape == (P, nt)
        for p in (0, 3, 4, 31, 32, 36, P - 1):
            assert np.array_equal(W[p], param_KL_Brownian_motion(tt, yy[p])), (d, p)
        ref = np.array([[np.sum([pi * (n + 0.5) * np.sqrt(2) * np.sin((n + 0.5) * pi * t) * yy[7, n]
                                 for n in range(d)]) for t in tt]])
        assert rel_err(W[7:8], ref) <= 1e-11


# ------------------------------------------------------------------ composite drivers (a12, a14)
def _ml_setup(kind_cls, NF, knots=None):
    rng = np.random.default_rng(7)
    knots = knots or nodes_1d.cc_nodes
    lev = lambda n: np.where(n == 0, 1, 2 ** n + 1)  # noqa: E731
    seq = [SGInterpolant(mis.smolyak_mid_set(k, 4), knots, lev, tp_interpolant=kind_cls,
                         verbose=False) for k in range(4)]
    K = len(seq) - 1
    ml = [rng.normal(size=(seq[K - k].num_nodes, NF)) for k in range(K + 1)]
    yy = rng.uniform(-1, 1, (400, 4))
    return seq, ml, yy


@pytest.mark.parametrize("kind,NF", [("linear", 1), ("quadratic", 24), ("linear", 70)])
def test_multilevel_interpolant_against_oracle(kind, NF):
    from oracle import drivers_oracle
    from sgmethods_b200.multilevel_interpolant import MLInterpolant, MultilevelInterpolant
    assert MultilevelInterpolant is MLInterpolant
    seq, ml, yy = _ml_setup(KIND_CLASS[kind], NF)
    mli = MLInterpolant(seq)
    plans = [oracle_plan_of(it) for it in seq]
    for p, it in zip(plans, seq):
        p.SG = it.SG
    ref = drivers_oracle.ml_interpolate(plans, kind, yy, ml)
    got = mli.interpolate(yy, ml)
    assert rel_err(got, np.squeeze(ref)) <= TOL
    ref_terms = drivers_oracle.ml_terms(plans, kind, yy, ml)
    got_terms = mli.get_ml_terms(yy, ml)
    assert len(got_terms) == 4
    for a, b in zip(got_terms, ref_terms):
        assert rel_err(a, np.squeeze(b)) <= TOL
    assert mli.total_cost(np.array([1., 10., 100., 1000.])) == \
        sum(seq[k].num_nodes * [1., 10., 100., 1000.][3 - k] for k in range(4))


def test_multilevel_sample_and_non_nested_grids():
    from sgmethods_b200.multilevel_interpolant import MLInterpolant
    seq, _, yy = _ml_setup(TPPwLinearInterpolator, 1)
    mli = MLInterpolant(seq)
    ml = mli.sample(lambda y, k: np.array([np.sin(np.sum(y)) + 2.0 ** (-k), float(k)]))
    assert [s.shape for s in ml] == [(seq[3 - k].num_nodes, 2) for k in range(4)]
    assert np.all(ml[2][:, 1] == 2.0)
    val = mli.interpolate(yy, ml)
    assert val.shape == (400, 2) and np.all(np.isfinite(val))
    lev = lambda n: n + 1  # noqa: E731   (n+1 knots: the two families below do not nest)
    bad = [SGInterpolant(mis.smolyak_mid_set(k, 2), fam, lev, verbose=False)
           for k, fam in ((2, nodes_1d.equispaced_interior_nodes), (3, nodes_1d.cc_nodes))]
    with pytest.raises(ZeroDivisionError):
        MLInterpolant(bad).interpolate(yy[:, :2], [np.ones((bad[1].num_nodes, 1)),
                                                   np.ones((bad[0].num_nodes, 1))])


def test_error_indicators_against_oracle():
    from oracle import drivers_oracle
    from sgmethods_b200.compute_error_indicators import compute_GG_indicators_RM
    from sgmethods_b200.mid_set import MidSet
    ms = MidSet(track_reduced_margin=True)
    for op, arg in (("e", 0), ("d", None), ("e", 1), ("e", 0), ("d", None), ("e", 2)):
        ms.enlarge_from_margin(arg) if op == "e" else ms.increase_dim()
    knots = nodes_1d.unbounded_nodes_nested
    lev = lambda n: 2 ** (n + 1) - 1  # noqa: E731
    f = lambda y: np.array([np.exp(-0.3 * np.sum(y * y)), np.sum(y)])  # noqa: E731
    norm = lambda a, b: float(np.sqrt(np.mean(np.sum((a - b) ** 2, axis=1))))  # noqa: E731
    cur = SGInterpolant(ms.mid_set, knots, lev, verbose=False)
    u_on_SG = cur.sample_on_SG(f)
    yy = np.random.default_rng(3).normal(size=(500, ms.get_dim()))
    u_interp = cur.interpolate(yy, u_on_SG)
    ref = drivers_oracle.gg_indicators(ms.mid_set, ms.reduced_margin, knots, lev, f, cur.SG,
                                       u_on_SG, yy, norm, u_interp)
    scale = max(1.0, float(np.max(np.abs(u_interp))))
    for method, tol in (("rebuild", 1e-12), ("surplus", 1e-11)):
        calls = []
        got = compute_GG_indicators_RM(np.zeros((0, 0), dtype=int), np.zeros(0), ms, knots, lev,
                                       lambda y: (calls.append(1), f(y))[1], cur.SG, u_on_SG, yy,
                                       norm, u_interp, method=method)
        assert got.shape == (ms.get_cardinality_reduced_margin(),)
        assert np.max(np.abs(got - ref)) <= tol * scale, (method, np.max(np.abs(got - ref)))
        assert np.all(got > 0)
        print(method, "f evaluations:", len(calls), "max dev", np.max(np.abs(got - ref)))
    # recycling: unchanged reduced-margin entries are copied, not recomputed
    calls = []
    again = compute_GG_indicators_RM(ms.reduced_margin[:-1], got[:-1] + 1.0, ms, knots, lev,
                                     lambda y: (calls.append(1), f(y))[1], cur.SG, u_on_SG, yy,
                                     norm, u_interp)
    assert np.array_equal(again[:-1], got[:-1] + 1.0) and abs(again[-1] - got[-1]) < 1e-14


# Reference-pinned: fixtures ml_*.npz / gg_*.npz are outputs of the reference's own driver code
# (run on alias subclasses, tests/golden/make_golden.py).
@pytest.mark.parametrize("name", ["ml_lin_nf1", "ml_quad_nf3", "ml_lin_grow"])
def test_multilevel_against_reference_driver_outputs(name):
    from golden_cases import ML_CASES, ML_F_APPROX, ml_mid_set
    from sgmethods_b200.multilevel_interpolant import MLInterpolant
    fam, l2k, kind, levels, NF, P, seed = ML_CASES[name]
    g = load_golden(name)
    seq = [SGInterpolant(np.asarray(ml_mid_set(mis, name, k)), knot_family(fam, nodes_1d),
                         LEV2KNOTS[l2k], tp_interpolant=KIND_CLASS[kind], verbose=False)
           for k in range(levels)]
    for k, it in enumerate(seq):
        assert it.SG.tobytes() == g[f"SG_{k}"].tobytes()
    mli = MLInterpolant(seq)
    ml = [g[f"ml_{k}"] for k in range(levels)]
    out = mli.interpolate(g["yy"], ml)
    assert out.shape == g["out"].shape                     # squeezed like the reference
    assert rel_err(out, g["out"]) <= TOL
    terms = mli.get_ml_terms(g["yy"], ml)
    assert len(terms) == levels
    for k in range(levels):
        assert rel_err(terms[k], g[f"term_{k}"]) <= TOL
    sampled = mli.sample(ML_F_APPROX)
    for k in range(levels):
        assert np.array_equal(sampled[k], g[f"sampled_{k}"])
    assert rel_err(mli.interpolate(g["yy"], sampled), g["out_sampled"]) <= TOL
    assert mli.total_cost(np.arange(1., levels + 1.)) == float(g["cost"])


@pytest.mark.parametrize("name", ["gg_lin_unb", "gg_quad_equi"])
@pytest.mark.parametrize("method", ["surplus", "rebuild", "surplus_gpu_norm"])
def test_error_indicators_against_reference_driver_outputs(name, method):
    from golden_cases import GG_CASES, GG_F, GG_NORM
    from sgmethods_b200.compute_error_indicators import MeanSquareNorm, compute_GG_indicators_RM
    norm = GG_NORM
    if method == "surplus_gpu_norm":     # same norm, reduced on the GPU (sgm_surplus_norms)
        method, norm = "surplus", MeanSquareNorm()
    from sgmethods_b200.mid_set import MidSet
    fam, l2k, kind, script, P, seed = GG_CASES[name]
    g = load_golden(name)
    knots, lev = knot_family(fam, nodes_1d), LEV2KNOTS[l2k]
    ms = MidSet(track_reduced_margin=True)
    for step, ops in enumerate(script):
        for op, arg in ops:
            ms.enlarge_from_margin(arg) if op == "e" else ms.increase_dim()
        assert np.array_equal(ms.reduced_margin, g[f"rm_{step}"])
        calls = []

        def f_count(y):
            calls.append(1)
            return GG_F(y)
        est = compute_GG_indicators_RM(
            g[f"old_rm_{step}"], g[f"old_est_{step}"], ms, knots, lev, f_count, g[f"SG_{step}"],
            g[f"u_on_SG_{step}"], g[f"yy_{step}"], norm, g[f"u_interp_{step}"],
            TP_inteporlant=KIND_CLASS[kind], method=method)
        ref = g[f"est_{step}"]
        scale = max(1.0, float(np.max(np.abs(g[f"u_interp_{step}"]))))
        assert est.shape == ref.shape
        assert np.max(np.abs(est - ref)) <= 1e-12 * scale, (step, np.max(np.abs(est - ref)))
        assert len(calls) <= int(g[f"n_f_calls_{step}"])     # never more samples than the reference


def test_adaptive_loop_moves_on_without_reevaluating_or_resampling():
    """The incremental update of the adaptive loop: the indicator launch
    hands back, for every candidate, the extended interpolant's values at the random points and
    the samples taken on its new nodes; after choosing a candidate the next iteration needs no
    evaluation and no call of f.  Checked against the reference's sequence (new SGInterpolant,
    sample_on_SG with recycling, interpolate) over three adaptive steps."""
    from sgmethods_b200.compute_error_indicators import compute_GG_indicators_RM
    from sgmethods_b200.mid_set import MidSet
    knots, lev = nodes_1d.unbounded_nodes_nested, (lambda n: 2 ** (n + 1) - 1)
    calls = []

    def f(y):
        calls.append(1)
        return np.array([np.exp(-0.3 * np.sum(y * y)), np.sin(np.sum(y))])
    norm = lambda a, b: float(np.sqrt(np.mean(np.sum((a - b) ** 2, axis=1))))  # noqa: E731
    ms = MidSet(track_reduced_margin=True)
    for op, arg in (("e", 0), ("d", None), ("e", 1), ("d", None), ("e", 0), ("e", 2)):
        ms.enlarge_from_margin(arg) if op == "e" else ms.increase_dim()
    yy = np.random.default_rng(21).normal(size=(3000, 6))
    cur = SGInterpolant(ms.mid_set, knots, lev, verbose=False)
    SG, u_on_SG = cur.SG, cur.sample_on_SG(f)
    u_interp = cur.interpolate(yy, u_on_SG)
    none_rm, none_est = np.zeros((0, 0), dtype=int), np.zeros(0)   # no recycled indicators: all computed
    for step in range(3):
        ext = {}
        est = compute_GG_indicators_RM(none_rm, none_est, ms, knots, lev, f, SG, u_on_SG, yy, norm,
                                       u_interp, extensions=ext)
        assert sorted(ext) == list(range(ms.get_cardinality_reduced_margin()))
        best = int(np.argmax(est))
        nu = ms.reduced_margin[best]
        ms.enlarge_from_margin(int(np.flatnonzero((ms.margin == nu).all(axis=1))[0]))
        # reference sequence for the next iteration, fed with what the launch handed back
        N = ms.get_dim()
        pad = lambda a: np.hstack((a, np.zeros((a.shape[0], N - a.shape[1]))))  # noqa: E731
        old_xx = np.vstack((pad(SG), pad(ext[best]["nodes"])))
        old_samples = np.vstack((u_on_SG, ext[best]["samples"]))
        nxt = SGInterpolant(ms.mid_set, knots, lev, verbose=False)
        n_before = len(calls)
        u_next = nxt.sample_on_SG(f, old_xx=old_xx, old_samples=old_samples)
        assert len(calls) == n_before                     # every node recycled: f is not called
        full = nxt.interpolate(yy, u_next)                # what the loop no longer has to compute
        assert rel_err(ext[best]["values"], full) <= 1e-12
        SG, u_on_SG, u_interp = nxt.SG, u_next, ext[best]["values"]


def test_pipelined_host_path_equals_device_path():
    """Pinned host points take the chunked copy/compute pipeline; same bits as one launch."""
    import torch
    it = SGInterpolant(mis.smolyak_mid_set(2, 6), nodes_1d.unbounded_nodes_nested,
                       lambda n: 2 ** (n + 1) - 1, verbose=False)
    P = 4 * 131072 + 777
    xh = torch.randn((P, 9), dtype=torch.float64).pin_memory()       # 3 unused columns
    f = torch.randn((it.num_nodes, 1), dtype=torch.float64)
    got = it.interpolate(xh, f)
    ref = it.interpolate(xh.cuda(), f.cuda())
    assert got.is_cuda and torch.equal(got, ref)


# ------------------------------------------------------------------------------ edge cases
def test_empty_ragged_and_strided_inputs():
    import torch
    it = SGInterpolant(mis.smolyak_mid_set(3, 3), nodes_1d.cc_nodes,
                       lambda n: np.where(n == 0, 1, 2 ** n + 1), verbose=False)
    plan = oracle_plan_of(it)
    rng = np.random.default_rng(5)
    f = rng.normal(size=(it.num_nodes, 5))
    # no points
    assert it.interpolate(np.zeros((0, 3)), f).shape == (0, 5)
    assert it.interpolate(np.zeros((0, 3)), f[:, :1]).shape == (0,)
    # one point, and point counts around the 32-point tile / 8-warp block boundaries
    for P in (1, 31, 32, 33, 63, 65, 257):
        x = rng.uniform(-1, 1, (P, 3))
        for cols in (1, 5):
            ref = sg_oracle.interpolate(plan, x, f[:, :cols], squeeze=False)
            got = np.asarray(it.interpolate(x, f[:, :cols])).reshape(ref.shape)
            assert np.array_equal(got, ref), (P, cols)
    # non-contiguous views: strided rows of x, column slice of f, extra rows in f
    xbig = rng.uniform(-1, 1, (40, 7))
    fbig = rng.normal(size=(it.num_nodes + 3, 9))
    ref = sg_oracle.interpolate(plan, xbig[::2, :3], fbig[:it.num_nodes, 2:6], squeeze=False)
    got = it.interpolate(torch.from_numpy(xbig).cuda()[::2], torch.from_numpy(fbig).cuda()[:, 2:6])
    assert np.array_equal(got.cpu().numpy(), ref)
    # 1-D x is one column (reference wrapper, tp_interpolant_wrapper.py:54-55)
    it1 = SGInterpolant(mis.smolyak_mid_set(3, 1), nodes_1d.cc_nodes,
                        lambda n: np.where(n == 0, 1, 2 ** n + 1), verbose=False)
    f1 = rng.normal(size=(it1.num_nodes, 1))
    x1 = rng.uniform(-1, 1, 17)
    assert np.array_equal(it1.interpolate(x1, f1),
                          sg_oracle.interpolate(oracle_plan_of(it1), x1, f1))
    # too few rows in f / wrong rank
    with pytest.raises(IndexError):
        it.interpolate(np.zeros((2, 3)), f[:10])
    with pytest.raises(IndexError):
        it.interpolate(np.zeros((2, 3)), f[:, 0])


def test_vector_outputs_across_column_tile_boundaries():
    """NF around the 32/64/128/256-column tile sizes of the column kernels (wide kernel +
    guarded remainder), odd leading dimensions (unaligned rows fall back to 8-byte loads)."""
    import torch
    it = SGInterpolant(mis.smolyak_mid_set(3, 4), nodes_1d.unbounded_nodes_nested,
                       lambda n: 2 ** (n + 1) - 1, verbose=False)
    plan = oracle_plan_of(it)
    rng = np.random.default_rng(6)
    x = rng.normal(size=(37, 4))
    fall = rng.normal(size=(it.num_nodes, 700))
    for NF in (2, 31, 33, 64, 127, 128, 129, 255, 256, 257, 511, 513, 700):
        ref = sg_oracle.interpolate(plan, x, fall[:, :NF], squeeze=False)
        got = it.interpolate(x, np.ascontiguousarray(fall[:, :NF])).reshape(ref.shape)
        assert np.array_equal(got, ref), NF
    # a view with odd leading dimension and odd column offset: rows are not 16-byte aligned
    fd = torch.from_numpy(fall).cuda()
    ref = sg_oracle.interpolate(plan, x, fall[:, 1:300], squeeze=False)
    got = it.interpolate(torch.from_numpy(x).cuda(), fd[:, 1:300])
    assert np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.parametrize("NF", [523, 457])
@pytest.mark.parametrize("kind", ["linear", "quadratic", "cubic", "lagrange"])
def test_two_points_per_warp_wide_kernel_equals_one_point_kernel_and_oracle(kind, NF, monkeypatch):
    """>= 4096 points and >= 512 columns go through the wide kernel with two points per warp
    (points paired by the locality permutation, shared rows of f loaded once): bit-identical to
    the one-point-per-warp kernel and to the oracle, odd point count, ragged column remainder."""
    rng = np.random.default_rng(77)
    cfg = {"linear": (nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, "normal"),
           "quadratic": (nodes_1d.cc_nodes, lambda n: np.where(n == 0, 1, 2 ** n + 1), "uniform"),
           "cubic": (nodes_1d.equispaced_nodes, lambda n: 3 * n + 1, "normal"),
           "lagrange": (nodes_1d.hermite_nodes, lambda n: 2 * n + 1, "normal")}
    knots, lev, smp = cfg[kind]
    it = SGInterpolant(mis.smolyak_mid_set(3, 5), knots, lev, tp_interpolant=KIND_CLASS[kind], verbose=False)
    P = 4611                                        # NF = 457: 128- and 64-column tiles
    x = rng.normal(size=(P, 5)) if smp == "normal" else rng.uniform(-1, 1, (P, 5))
    f = rng.normal(size=(it.num_nodes, NF))
    got = it.interpolate(x, f)
    monkeypatch.setenv("SGM_WIDE_PAIRS", "0")
    reload_knobs()
    one = it.interpolate(x, f)
    monkeypatch.delenv("SGM_WIDE_PAIRS")
    reload_knobs()
    assert np.array_equal(got, one, equal_nan=True)
    pick = np.concatenate([rng.choice(P, 40, replace=False), [P - 1]])
    ref = sg_oracle.interpolate(oracle_plan_of(it), x[pick], f, kind, squeeze=False)
    assert rel_err(got[pick], ref) <= TOL


def test_term_sharding_plan_slices_on_one_gpu():
    """mode='terms' building block: evaluating contiguous term ranges with separate plans and
    adding the partial results reproduces the full interpolant to rounding."""
    import torch
    from sgmethods_b200._device import DevicePlan
    it = SGInterpolant(mis.smolyak_mid_set(3, 5), nodes_1d.unbounded_nodes_nested,
                       lambda n: 2 ** (n + 1) - 1, verbose=False)
    rng = np.random.default_rng(8)
    x = torch.from_numpy(rng.normal(size=(200, 5))).cuda()
    f = torch.from_numpy(rng.normal(size=(it.num_nodes, 1))).cuda()
    full = it.device_plan().eval(x, f)
    T = len(it._coeffs)
    total = torch.zeros_like(full)
    for t0, t1 in ((0, T // 3), (T // 3, T // 2), (T // 2, T)):
        lo, hi = it._map_off[t0], it._map_off[t1]
        part = DevicePlan(it._kind, it.N, it.num_nodes, it._coeffs[t0:t1].astype(float),
                          it._term_fam[t0:t1], it._map_off[t0:t1 + 1] - lo, it._maps[lo:hi],
                          it._fam_off, it._fam_knots)
        total += part.eval(x, f)
    assert float((total - full).abs().max()) <= 1e-12 * float(full.abs().max())


def test_wrapper_shim_matches_reference_semantics():
    from sgmethods_b200.tp_interpolant_wrapper import TPInterpolatorWrapper
    rng = np.random.default_rng(4)
    nodes = (np.linspace(-1, 1, 5), np.linspace(-1, 1, 3))
    vals = rng.normal(size=(5, 3, 2))
    x = rng.uniform(-1, 1, (20, 6))
    w = TPInterpolatorWrapper(nodes, np.array([1, 4]), vals, TPPwLinearInterpolator)
    assert rel_err(w(x), sg_oracle.tp_linear(nodes, vals, x[:, [1, 4]])) <= TOL
    const = TPInterpolatorWrapper((0,), np.array([], dtype=int), np.array([[3.0, -1.0]]),
                                  TPPwLinearInterpolator)
    assert np.array_equal(const(x), np.tile([[3.0, -1.0]], (20, 1)))


def test_every_scalar_kernel_flavour_against_oracle():
    """Kernel selection paths of the scalar (NF = 1) evaluation: thread-per-point kernel
    (few terms, many points), split kernel with the locality sort (>= 32768 points),
    wide-linear flavour (terms with more than 4 active directions)."""
    rng = np.random.default_rng(12)
    # few terms + many points -> eval_points_kernel
    it = SGInterpolant(mis.smolyak_mid_set(1, 3), nodes_1d.unbounded_nodes_nested,
                       lambda n: 2 ** (n + 1) - 1, verbose=False)
    x = rng.normal(size=(20000, 3))
    f = rng.normal(size=(it.num_nodes, 1))
    ref = sg_oracle.interpolate(oracle_plan_of(it), x, f)
    assert np.array_equal(it.interpolate(x, f), ref)
    # >= 32768 points -> counting sort + permuted tiles; check a subset against the oracle
    it = SGInterpolant(mis.smolyak_mid_set(3, 5), nodes_1d.unbounded_nodes_nested,
                       lambda n: 2 ** (n + 1) - 1, verbose=False)
    x = rng.normal(size=(50001, 5))
    f = rng.normal(size=(it.num_nodes, 1))
    got = it.interpolate(x, f)
    pick = rng.choice(50001, 600, replace=False)
    assert np.array_equal(got[pick], sg_oracle.interpolate(oracle_plan_of(it), x[pick], f))
    # k = 6 active directions -> nested-walk flavour of the pw-linear kernels
    it = SGInterpolant(mis.tensor_product_mid_set(1, 6), nodes_1d.cc_nodes, lambda n: 2 * n + 1,
                       verbose=False)
    x = rng.uniform(-1.2, 1.2, (300, 6))
    f = rng.normal(size=(it.num_nodes, 1))
    assert np.array_equal(it.interpolate(x, f), sg_oracle.interpolate(oracle_plan_of(it), x, f))
    x = rng.uniform(-1.2, 1.2, (40000, 6))
    got = it.interpolate(x, f)
    assert np.array_equal(got[:400], sg_oracle.interpolate(oracle_plan_of(it), x[:400], f))


@pytest.mark.parametrize("run", ["4", "8", "8u", "8p2"])
def test_prefix_kernel_equals_split_kernel_and_oracle(run, monkeypatch):
    """The prefix kernel (consecutive terms per warp, weight / offset prefixes reused, value
    tables first-direction-fastest) performs the same roundings as the split kernel: results
    must be bit-identical to it and to the oracle, for every run partition (uniform runs of 4
    or 8 terms, cost-balanced runs, one or two points per lane), with and without the locality sort, on isotropic, anisotropic and tiny index sets (ragged last chunk,
    constant term, k = 1..4)."""
    rng = np.random.default_rng(21)
    l2k = lambda n: 2 ** (n + 1) - 1
    sets = [mis.smolyak_mid_set(4, 6), mis.smolyak_mid_set(3, 9), mis.smolyak_mid_set(1, 2),
            mis.aniso_smolyak_mid_set(5.0, 4, np.array([1.0, 1.5, 2.0, 1.2])),
            mis.tensor_product_mid_set(2, 4)]
    for mid in sets:
        it = SGInterpolant(mid, nodes_1d.unbounded_nodes_nested, l2k, verbose=False)
        f = rng.normal(size=(it.num_nodes, 1))
        for P in (777, 33000):
            x = rng.normal(size=(P, it.N))
            monkeypatch.setenv("SGM_PREFIX_RUN", run.rstrip("u").split("p")[0])
            monkeypatch.setenv("SGM_PREFIX_PPL", "2" if run.endswith("p2") else "1")
            if run.endswith("u"):
                monkeypatch.setenv("SGM_PREFIX_UNBALANCED", "1")
            monkeypatch.setenv("SGM_FORCE_KERNEL", "prefix")
            reload_knobs()
            got = it.interpolate(x, f)
            monkeypatch.setenv("SGM_FORCE_KERNEL", "split")
            reload_knobs()
            old = it.interpolate(x, f)
            monkeypatch.delenv("SGM_FORCE_KERNEL")
            for name in ("SGM_PREFIX_RUN", "SGM_PREFIX_PPL", "SGM_PREFIX_UNBALANCED"):
                monkeypatch.delenv(name, raising=False)
            reload_knobs()
            default = it.interpolate(x, f)
            assert np.array_equal(got, old, equal_nan=True)
            assert np.array_equal(default, old, equal_nan=True)
            ref = sg_oracle.interpolate(oracle_plan_of(it), x[:500], f)
            assert np.array_equal(got[:500], ref)
    # two output columns go through the per-column scalar path (strided f / out)
    it = SGInterpolant(mis.smolyak_mid_set(3, 5), nodes_1d.unbounded_nodes_nested, l2k, verbose=False)
    f = rng.normal(size=(it.num_nodes, 3))
    x = rng.normal(size=(1000, 5))
    assert np.array_equal(it.interpolate(x, f), sg_oracle.interpolate(oracle_plan_of(it), x, f))


@pytest.mark.parametrize("shape", ["example1", "not_nested", "three_dirs"])
def test_small_plan_kernel_few_terms_many_points(shape, monkeypatch):
    """Kernel A0 (few terms, >= 16384 points: metadata in kernel parameters, coarser brackets of
    a direction derived from the fine one for bitwise-nested families) against the oracle and,
    on every point, against kernel A (thread per point, searches per entry)."""
    rng = np.random.default_rng(77)
    if shape == "example1":            # example/1 as shipped: unbounded nested knots, 7 terms
        S = mis.aniso_smolyak_mid_set(5, 10, 2 ** (np.linspace(0, 9, 10)))
        it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False)
        x = rng.normal(0, 1.5, (40000, 10))
    elif shape == "not_nested":        # cc knots with n + 1 points: every family searches
        it = SGInterpolant(mis.smolyak_mid_set(2, 3), nodes_1d.cc_nodes, lambda n: n + 1, verbose=False)
        x = rng.uniform(-1.2, 1.2, (33333, 3))
    else:                              # doubling cc (nested), 3 active directions per term
        it = SGInterpolant(mis.tensor_product_mid_set(1, 3), nodes_1d.cc_nodes,
                           lambda n: np.where(n == 0, 1, 2 ** n + 1), verbose=False)
        x = rng.uniform(-1.1, 1.1, (20001, 3))
    x[0] = 0.0
    x[1, 0] = np.nan
    x[2] = it.SG[-1][:x.shape[1]] if it.SG.shape[1] >= x.shape[1] else 0.0      # exactly on knots
    assert len(it.combination_coeffs) <= 24
    f = rng.normal(size=(it.num_nodes, 1))
    got = it.interpolate(x, f)
    ref = sg_oracle.interpolate(oracle_plan_of(it), x[:600], f)
    assert np.array_equal(got[:600], ref, equal_nan=True)
    monkeypatch.setenv("SGM_NO_SMALL", "1")
    reload_knobs()
    old = it.interpolate(x, f)
    monkeypatch.delenv("SGM_NO_SMALL")
    reload_knobs()
    assert np.array_equal(got, old, equal_nan=True)
    f3 = rng.normal(size=(it.num_nodes, 3))                # per-column scalar path
    assert np.array_equal(it.interpolate(x[:20000], f3)[:300],
                          sg_oracle.interpolate(oracle_plan_of(it), x[:300], f3), equal_nan=True)


@pytest.mark.parametrize("shape", ["example1", "not_nested", "three_dirs", "nested_not_dyadic"])
def test_plan_specialised_kernel_equals_generic_small_kernel(shape, monkeypatch):
    """The NVRTC-generated kernel of a few-term plan (csrc/sgm_jit.cpp; tables as literals,
    bracket coordinates in registers, next tile's coordinates prefetched) against kernel A0 on
    every point (bit for bit, NaN and on-knot points included) and against the oracle."""
    import torch
    rng = np.random.default_rng(78)
    if shape == "example1":
        S = mis.aniso_smolyak_mid_set(5, 10, 2 ** (np.linspace(0, 9, 10)))
        it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False)
        x = rng.normal(0, 1.5, (300001, 12))                 # ldx > N, ragged last tile
    elif shape == "not_nested":
        it = SGInterpolant(mis.smolyak_mid_set(2, 3), nodes_1d.cc_nodes, lambda n: n + 1, verbose=False)
        x = rng.uniform(-1.2, 1.2, (33333, 3))
    elif shape == "three_dirs":
        it = SGInterpolant(mis.tensor_product_mid_set(1, 3), nodes_1d.cc_nodes,
                           lambda n: np.where(n == 0, 1, 2 ** n + 1), verbose=False)
        x = rng.uniform(-1.1, 1.1, (20001, 3))
    else:      # nested (1, 3, 9, 27 knots: every level keeps the old ones) but not by halving counts
        knots = lambda n: np.atleast_1d(np.linspace(-1.0, 1.0, int(n))) if n > 1 else np.zeros(1)  # noqa: E731
        it = SGInterpolant(mis.smolyak_mid_set(3, 2), knots, lambda n: 3 ** n, verbose=False)
        x = rng.uniform(-1.3, 1.3, (50000, 2))
    x[0] = 0.0
    x[1, 0] = np.nan
    x[2, :it.N] = it.SG[-1]
    f = rng.normal(size=(it.num_nodes, 1))
    xd, fd = torch.from_numpy(x).cuda(), torch.from_numpy(f).cuda()
    monkeypatch.setenv("SGM_NO_JIT", "1")
    plan0 = it.device_plan()
    plan0.reload_knobs()
    generic = plan0.eval(xd, fd).cpu().numpy()
    assert plan0.jit_state()[0] == 0
    monkeypatch.delenv("SGM_NO_JIT")
    monkeypatch.setenv("SGM_JIT_MIN_POINTS", "1")
    monkeypatch.setenv("SGM_JIT_SYNC", "1")
    plan0.reload_knobs()
    got = plan0.eval(xd, fd)
    state, why = plan0.jit_state()
    assert state == 1, why
    assert np.array_equal(got.cpu().numpy(), generic, equal_nan=True)
    ref = sg_oracle.interpolate(oracle_plan_of(it), x[:400, :it.N], f)
    assert np.array_equal(got.cpu().numpy()[:400, 0], ref, equal_nan=True)
    # accumulate / scale path of the same kernel through the multi-column loop (strided f, out)
    f3 = rng.normal(size=(it.num_nodes, 3))
    got3 = plan0.eval(xd[:40000], torch.from_numpy(f3).cuda()).cpu().numpy()
    assert np.array_equal(got3[:300], sg_oracle.interpolate(oracle_plan_of(it), x[:300, :it.N], f3), equal_nan=True)


def test_plan_specialised_kernel_is_built_in_the_background(monkeypatch):
    """Default mode: the call that crosses the point threshold starts the build on a thread and
    is itself served by the generic kernel; after sgm_plan_jit_wait the specialised kernel runs.
    Same bits before and after."""
    import torch
    monkeypatch.setenv("SGM_JIT_MIN_POINTS", "50000")
    monkeypatch.delenv("SGM_JIT_SYNC", raising=False)
    S = mis.aniso_smolyak_mid_set(5, 10, 2 ** (np.linspace(0, 9, 10)))
    it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False)
    plan = it.device_plan()
    plan.reload_knobs()
    x = torch.randn((40000, 10), dtype=torch.float64, device="cuda")
    f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device="cuda")
    a = plan.eval(x, f)
    assert plan.jit_state()[0] == 0                          # 40000 < 50000 points so far
    b = plan.eval(x, f)                                      # crosses the threshold
    assert plan.jit_wait() == 1, plan.jit_state()
    c = plan.eval(x, f)
    assert torch.equal(a, b) and torch.equal(a, c)


def test_levy_ciesielski_many_levels_uses_generic_kernel():
    rng = np.random.default_rng(13)
    tt = np.linspace(0, 1, 29)
    yy = rng.normal(size=(3, 5000))                 # L = 13 > 12
    W = param_LC_Brownian_motion(tt, yy, 1.0)
    assert np.array_equal(W[1], brownian_oracle.lc_brownian(tt, yy[1], 1.0))


# ------------------------------------------------------- opt-in hierarchical-surplus form
@pytest.mark.parametrize("NF", [1, 3, 200])
def test_hierarchical_form_agrees_with_combination_formula(NF):
    """Not bit-identical by design (different but equivalent formula); 1e-12 of max|ref| here
    (the north-star tolerance), and at least as close to a long-double evaluation as the
    float64 combination formula (next test)."""
    rng = np.random.default_rng(21)
    it = SGInterpolant(mis.smolyak_mid_set(4, 6), nodes_1d.unbounded_nodes_nested,
                       lambda n: 2 ** (n + 1) - 1, verbose=False)
    x = rng.normal(0, 1.5, (500, 6))
    f = rng.normal(size=(it.num_nodes, NF))
    ref = sg_oracle.interpolate(oracle_plan_of(it), x, f, squeeze=False)
    got = np.asarray(it.interpolate(x, f, method="hierarchical")).reshape(ref.shape)
    assert np.max(np.abs(got - ref)) <= TOL * np.max(np.abs(ref))
    exact = np.asarray(it.interpolate(x, f)).reshape(ref.shape)
    assert np.array_equal(exact, ref)                        # the default stays bit-identical
    with pytest.raises(ValueError):
        it.interpolate(x, f, method="fast")


def test_hierarchical_form_smooth_function_vs_long_double():
    S = mis.smolyak_mid_set(4, 8)
    it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False)
    w = 1.0 / np.arange(1, 9) ** 2
    f = np.sin(it.SG @ w).reshape(-1, 1)
    x = np.random.default_rng(0).normal(size=(300, 8))
    plan = oracle_plan_of(it)
    # long-double evaluation of the combination formula (same algorithm, 64-bit mantissa)
    truth = np.zeros(300, dtype=np.longdouble)
    for c, dims, nodes, tmap in zip(plan.coeffs, plan.dims, plan.nodes, plan.maps):
        vals = f[tmap, 0].astype(np.longdouble)
        if dims.size == 0:
            truth += np.longdouble(c) * vals.reshape(-1)[0]
            continue
        acc = np.zeros(300, dtype=np.longdouble)
        lo, wy = [], []
        for j, d in enumerate(dims):
            g = np.asarray(nodes[j], dtype=np.longdouble)
            i = np.clip(np.searchsorted(nodes[j], x[:, d], side="right") - 1, 0, g.size - 2)
            lo.append(i)
            wy.append((x[:, d].astype(np.longdouble) - g[i]) / (g[i + 1] - g[i]))
        from itertools import product
        for corner in product((0, 1), repeat=dims.size):
            wgt = np.ones(300, dtype=np.longdouble)
            for j, b in enumerate(corner):
                wgt = wgt * (wy[j] if b else 1 - wy[j])
            acc = acc + vals[tuple(lo[j] + corner[j] for j in range(dims.size))] * wgt
        truth += np.longdouble(c) * acc
    exact = it.interpolate(x, f)
    fast = it.interpolate(x, f, method="hierarchical")
    err_exact = float(np.max(np.abs(exact - truth)))
    err_fast = float(np.max(np.abs(fast - truth)))
    print(f"float64 combination formula err {err_exact:.2e}, hierarchical form err {err_fast:.2e}")
    assert err_fast <= max(err_exact, 1e-15) * 1.5


def test_basis_tables_larger_than_shared_memory_use_the_global_scratch():
    """Lagrange with up to 257 knots per direction: the per-point basis table (~12 KB) does
    not fit shared memory for a 32-point tile, so the thread-per-point kernel keeps it in
    its per-block global scratch (eval_points_kernel<.., GLOBAL_BASIS=true>)."""
    rng = np.random.default_rng(31)
    knots = lambda n: np.atleast_1d(np.asarray(nodes_1d.cc_nodes(n), dtype=float))  # noqa: E731
    S = mis.smolyak_mid_set(8, 3)
    S = S[(S.max(axis=1) == S.sum(axis=1))]          # axis-aligned indices only: 25 cheap terms
    S = S[np.lexsort(np.rot90(S))]
    it = SGInterpolant(S, knots, lambda n: np.where(n == 0, 1, 2 ** n + 1),
                       tp_interpolant=TPLagrangeInterpolator, verbose=False)
    assert it._fam_m.max() == 257
    x = rng.uniform(-1, 1, (700, 3))
    f = rng.normal(size=(it.num_nodes, 1))
    ref = sg_oracle.interpolate(oracle_plan_of(it), x, f, "lagrange")
    got = it.interpolate(x, f)
    assert rel_err(got, ref) <= TOL
    fv = rng.normal(size=(it.num_nodes, 4))
    refv = sg_oracle.interpolate(oracle_plan_of(it), x[:64], fv, "lagrange")
    assert rel_err(it.interpolate(x[:64], fv), refv) <= TOL


def test_pipelined_numpy_path_equals_device_path():
    """Large numpy inputs are staged through two pinned buffers; same bits as one launch."""
    import torch
    it = SGInterpolant(mis.smolyak_mid_set(2, 6), nodes_1d.unbounded_nodes_nested,
                       lambda n: 2 ** (n + 1) - 1, verbose=False)
    P = 5 * 131072 + 333
    x = np.random.default_rng(2).normal(size=(P, 8))                 # 2 unused columns
    f = np.random.default_rng(3).normal(size=(it.num_nodes, 1))
    got = it.interpolate(x, f)
    ref = it.interpolate(torch.from_numpy(x).cuda(), torch.from_numpy(f).cuda()).cpu().numpy()
    assert isinstance(got, np.ndarray) and np.array_equal(got, ref)
    got2 = it.interpolate(x[::-1].copy(), f)                          # staging buffers reused
    assert np.array_equal(got2, ref[::-1])


def test_sharded_interpolant_over_nccl_on_two_gpus():
    """ShardedInterpolant through NCCL + the CUDA kernels (tools/dist_check.py under torchrun):
    points mode bit-identical to one GPU for the contract path and the exact path, terms mode
    within 1e-12.  Needs two GPUs (skipped on a one-GPU box; the CPU suite covers the host logic
    with gloo)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", "29577",
                           os.path.join(root, "tools", "dist_check.py")],
                          capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert proc.stdout.count("root ok True") >= 6
