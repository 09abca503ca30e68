"""Generate the golden fixtures in tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

It imports the reference package from /root/reference (read-only) with a 2-file matplotlib
stub on sys.path (the reference's utils.py imports matplotlib.pyplot at module top and
matplotlib is not installed here).  Everything written below is an OUTPUT of the reference
on seeded inputs; no reference source is copied.  The fixtures pin

  * oracle/ (the numpy restatement) in the `-m "not gpu"` tests, and
  * the CUDA path in the `-m gpu` tests,

on the same inputs.  Reference entry points exercised (file:line under /root/reference):
  sgmethods/sparse_grid_interpolant.py:20-139 (ctor tables), :141-250 (sample_on_SG),
  :252-279 (interpolate); sgmethods/tp_inteprolants.py (all four operators);
  sgmethods/multi_index_sets.py, mid_set.py, nodes_1d.py, utils.py;
  sgmethods/parametric_expansions_brownian_motion.py:17-91;
  sgmethods/multilevel_interpolant.py:89-150 and sgmethods/compute_error_indicators.py:78-121
  -- these two drivers raise AttributeError as shipped (camelCase leftovers: sampleOnSG,
  oldXx, oldSamples, numNodes, reducedMargin, midSet).  Their code is run UNMODIFIED here on
  objects that carry those names as aliases (subclasses defined below; nothing under
  /root/reference is edited), which pins the composite drivers to the reference's own
  arithmetic: `python tests/golden/make_golden.py drivers` regenerates only ml_*/gg_*.npz.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_mpl_stub"))
sys.path.insert(0, "/root/reference")
warnings.simplefilter("ignore")

from sgmethods.sparse_grid_interpolant import SGInterpolant  # noqa: E402
from sgmethods.tp_inteprolants import (TPPwLinearInterpolator,  # noqa: E402
                                       TPPwQuadraticInterpolator,
                                       TPPwCubicInterpolator,
                                       TPLagrangeInterpolator)
from sgmethods import multi_index_sets as mis  # noqa: E402
from sgmethods import nodes_1d as n1d  # noqa: E402
from sgmethods.mid_set import MidSet  # noqa: E402
from sgmethods import utils as rutils  # noqa: E402
from sgmethods.parametric_expansions_brownian_motion import (  # noqa: E402
    param_LC_Brownian_motion, param_KL_Brownian_motion)

KINDS = {"linear": TPPwLinearInterpolator, "quadratic": TPPwQuadraticInterpolator,
         "cubic": TPPwCubicInterpolator, "lagrange": TPLagrangeInterpolator}


def pack_plan(interp):
    """Flatten the reference object's list attributes into arrays an .npz can hold."""
    T = len(interp.combination_coeffs)
    dims_off = np.zeros(T + 1, dtype=np.int64)
    map_off = np.zeros(T + 1, dtype=np.int64)
    dims, shapes, maps = [], [], []
    for t in range(T):
        d = np.asarray(interp.active_tp_dims[t], dtype=np.int64)
        m = np.asarray(interp.map_tp_to_SG[t])
        dims.append(d)
        shapes.append(np.asarray(m.shape, dtype=np.int64))
        maps.append(m.reshape(-1).astype(np.int64))
        dims_off[t + 1] = dims_off[t] + d.size
        map_off[t + 1] = map_off[t] + m.size
    return dict(
        coeffs=np.asarray(interp.combination_coeffs, dtype=np.int64),
        active_mids=np.asarray(interp.active_mids, dtype=np.int64).reshape(T, -1),
        dims_off=dims_off, dims=np.concatenate(dims) if dims else np.zeros(0, np.int64),
        shape_off=np.cumsum([0] + [s.size for s in shapes]).astype(np.int64),
        shapes=np.concatenate(shapes), map_off=map_off, maps=np.concatenate(maps),
        SG=np.asarray(interp.SG, dtype=np.float64), num_nodes=np.int64(interp.num_nodes))


def lc_sq(tt, T):
    return lambda y: np.power(param_LC_Brownian_motion(tt, y, T=T), 2)


def interp_case(name, mid_set, knots, lev2knots, kind, x, f_on_SG=None, f=None,
                extra=None):
    interp = SGInterpolant(mid_set, knots, lev2knots, tp_interpolant=KINDS[kind],
                           verbose=False)
    if f_on_SG is None:
        f_on_SG = interp.sample_on_SG(f)
    out = interp.interpolate(x, f_on_SG)
    d = pack_plan(interp)
    d.update(mid_set=np.asarray(mid_set, dtype=np.int64), x=x, f_on_SG=f_on_SG, out=out,
             kind=np.array(kind))
    if extra:
        d.update(extra)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(f"{name}: T={len(interp.combination_coeffs)} M={interp.num_nodes} "
          f"out{np.shape(out)}")
    return interp, f_on_SG, out


class AliasSGInterpolant(SGInterpolant):
    """The reference class plus the camelCase names its own multilevel driver calls
    (multilevel_interpolant.py:84-85, :109-110, :144-146, :168)."""

    def sampleOnSG(self, f, oldXx=None, oldSamples=None):
        return self.sample_on_SG(f, old_xx=oldXx, old_samples=oldSamples)

    @property
    def numNodes(self):
        return self.num_nodes


class AliasMidSet(MidSet):
    """The reference class plus the camelCase names compute_GG_indicators_RM reads
    (compute_error_indicators.py:84, :89, :92, :101, :110-111)."""

    @property
    def reducedMargin(self):
        return self.reduced_margin

    @property
    def midSet(self):
        return self.mid_set


def drivers():
    """Fixtures of the two composite drivers, produced by the reference's own driver code."""
    from sgmethods.multilevel_interpolant import MLInterpolant
    from sgmethods.compute_error_indicators import compute_GG_indicators_RM
    sys.path.insert(0, os.path.dirname(HERE))
    import golden_cases as gc

    # ---- multilevel (multilevel_interpolant.py:89-150) ------------------------------------
    for name, (fam, l2k, kind, levels, NF, P, seed) in gc.ML_CASES.items():
        rng = np.random.default_rng(seed)
        knots = gc.knot_family(fam, n1d)
        seq = [AliasSGInterpolant(np.asarray(gc.ml_mid_set(mis, name, k)), knots,
                                  gc.LEV2KNOTS[l2k], tp_interpolant=KINDS[kind], verbose=False)
               for k in range(levels)]
        K = levels - 1
        ml = [rng.normal(size=(seq[K - k].num_nodes, NF)) for k in range(levels)]
        n_max = max(it.N for it in seq)
        yy = rng.uniform(-1, 1, (P, n_max)) if fam in ("cc", "equi_arr") else rng.normal(0, 1, (P, n_max))
        mli = MLInterpolant(seq)
        out = np.array(mli.interpolate(yy, ml))
        terms = mli.get_ml_terms(yy, ml)
        cost = mli.total_cost(np.arange(1., levels + 1.))
        # sample(): the closure of the reference binds k late but calls immediately -> fine
        f_approx = gc.ML_F_APPROX
        sampled = mli.sample(f_approx)
        d = dict(yy=yy, out=out, cost=np.float64(cost), levels=np.int64(levels),
                 out_sampled=np.array(mli.interpolate(yy, sampled)))
        for k in range(levels):
            d[f"ml_{k}"] = ml[k]
            d[f"term_{k}"] = np.array(terms[k])
            d[f"sampled_{k}"] = sampled[k]
            d[f"SG_{k}"] = seq[k].SG
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(f"{name}: levels={levels} NF={NF} out{out.shape} M={[it.num_nodes for it in seq]}")

    # ---- Gerstner-Griebel indicators (compute_error_indicators.py:78-121) -------------------
    for name, (fam, l2k, kind, script, P, seed) in gc.GG_CASES.items():
        knots = gc.knot_family(fam, n1d)
        lev = gc.LEV2KNOTS[l2k]
        ms = AliasMidSet(track_reduced_margin=True)
        d = {}
        old_rm, old_est = np.zeros((0, 0), dtype=int), np.zeros(0)
        rng = np.random.default_rng(seed)
        for step, ops in enumerate(script):
            for op, arg in ops:
                ms.enlarge_from_margin(arg) if op == "e" else ms.increase_dim()
            cur = SGInterpolant(ms.mid_set, knots, lev, tp_interpolant=KINDS[kind], verbose=False)
            u_on_SG = cur.sample_on_SG(gc.GG_F)
            yy = rng.normal(0, 1, (P, ms.get_dim())) if fam == "unb" else rng.uniform(-1, 1, (P, ms.get_dim()))
            u_interp = cur.interpolate(yy, u_on_SG)
            calls = []

            def f_count(y):
                calls.append(1)
                return gc.GG_F(y)
            est = compute_GG_indicators_RM(old_rm, old_est, ms, knots, lev, f_count, cur.SG,
                                           u_on_SG, yy, gc.GG_NORM, u_interp,
                                           TP_inteporlant=KINDS[kind])
            d[f"mid_{step}"] = ms.mid_set.copy()
            d[f"rm_{step}"] = ms.reduced_margin.copy()
            d[f"SG_{step}"] = cur.SG
            d[f"u_on_SG_{step}"] = u_on_SG
            d[f"yy_{step}"] = yy
            d[f"u_interp_{step}"] = np.array(u_interp)
            d[f"old_rm_{step}"] = old_rm
            d[f"old_est_{step}"] = old_est
            d[f"est_{step}"] = est
            d[f"n_f_calls_{step}"] = np.int64(len(calls))
            old_rm, old_est = ms.reduced_margin.copy(), est.copy()
        d["steps"] = np.int64(len(script))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(f"{name}: steps={len(script)} last est{est.shape}")


def main():
    rng = np.random.default_rng(20241019)

    # ---- interpolant cases -------------------------------------------------------------
    # (1) tests/test_sparse_grid_interpolant.py:15-33 shape (pw-linear, equispaced, n+1)
    Lam = np.array([[0, 0], [0, 1], [1, 0], [1, 1]])
    x = rng.uniform(-1.3, 1.3, (40, 2))
    interp_case("lin_deg1", Lam, n1d.equispaced_nodes, lambda n: n + 1, "linear", x,
                f=lambda y: 2. + y[0] + y[0] * y[1])

    # (2) test :35-57 shape (pw-cubic, 3n+1), incl. extrapolation on both sides
    Lam = np.array([[0, 0], [0, 1], [0, 2], [0, 3], [1, 0], [1, 1], [1, 2], [1, 3]])
    x = rng.uniform(-1.5, 1.5, (48, 2))
    interp_case("cub_deg3", Lam, n1d.equispaced_nodes, lambda n: 3 * n + 1, "cubic", x,
                f=lambda y: 2. + y[0] + y[0] * y[1] + y[1] ** 3 + y[0] * y[1] ** 3)

    # (3) test :60-106 shape: N=4, cc_nodes, nu+1 (NOT nested), smolyak w=3, exp(sum)
    x = rng.uniform(-1.2, 1.2, (96, 4))
    x[0] = 0.0
    x[1] = [-1.0, 1.0, 0.0, 0.5]            # exactly on knots / the domain corners
    x[2] = [np.nan, 0.1, 0.2, 0.3]          # scipy: f(nan) = nan
    interp_case("lin_exp_cc", mis.smolyak_mid_set(3, 4), n1d.cc_nodes, lambda nu: nu + 1,
                "linear", x, f=lambda y: np.exp(np.sum(y)))

    # (4) test :109-160 shape: N=4, NF=10 via LC^2, optimal gaussian nodes, profit set
    p = 2
    ell = lambda nu: np.where(nu == 0, 0, np.floor(np.log2(nu)).astype(int) + 1)  # noqa
    Profit = lambda nu: np.prod(np.where(nu == 1, 2. ** (-ell(nu) - nu) / 2,  # noqa
                                         (2. ** (-p * nu * (1 + ell(nu)))) / 2))
    Lam = mis.compute_mid_set_fast(Profit, 11. ** (-4), N=4)
    tt = np.linspace(0, 1, 10)
    x = rng.normal(0, 1.5, (64, 4))
    interp_case("lin_gauss_rf", Lam, lambda n: n1d.optimal_gaussian_nodes(n, p=p),
                lambda nu: 2 ** (nu + 1) - 1, "linear", x, f=lc_sq(tt, 1))

    # (5) nested doubling CC grid: the A.2 probe of the survey (map ordering, -6e-17 centre)
    Lam = mis.smolyak_mid_set(3, 3)
    l2k = lambda n: np.where(n == 0, 1, 2 ** n + 1)  # noqa
    x = rng.uniform(-1, 1, (64, 3))
    interp, fsg, _ = interp_case("lin_cc_nested", Lam, n1d.cc_nodes, l2k, "linear", x,
                                 f_on_SG=None, f=lambda y: np.array(
                                     [np.sin(y[0] + 2 * y[1] - y[2]), np.cos(y @ y), y[0]]))

    # (6) pw-quadratic, same nested CC grid, NF=3; ties at even knots, left extrapolation,
    #     no right extrapolation (the reference raises IndexError there)
    x = rng.uniform(-1, 1, (64, 3))
    x[0] = [0.0, -1.0, 1.0]
    x[1] = [-1.4, 0.5, -0.5]
    x[2] = [-np.sqrt(0.5), np.sqrt(0.5), 0.25]
    interp_case("quad_cc_nested", Lam, n1d.cc_nodes, l2k, "quadratic", x, f_on_SG=fsg)

    # (7) pw-quadratic, scalar output, 5-D anisotropic set
    Lam = mis.aniso_smolyak_mid_set(3, 5, [1, 1, 1.5, 2, 3])
    x = rng.uniform(-1, 1, (50, 5))
    # (the zero multi-index is active here; the reference's quadratic ctor needs
    #  len(knots(1)), so the knot callable returns a 1-element array instead of python 0)
    equi_arr = lambda n: np.atleast_1d(np.asarray(n1d.equispaced_nodes(n), dtype=float))  # noqa
    interp_case("quad_aniso5", Lam, equi_arr, l2k, "quadratic", x,
                f=lambda y: 1.0 / (2.5 + y @ np.array([1, .5, .25, .125, .0625])))

    # (8) pw-cubic, NF=2, 3 dims, extrapolation both sides
    Lam = mis.smolyak_mid_set(2, 3)
    x = rng.uniform(-1.6, 1.6, (64, 3))
    x[0] = [-1.0, 1.0, 1.0 / 3.0]
    interp_case("cub_smolyak3", Lam, equi_arr, lambda n: 3 * n + 1, "cubic", x,
                f=lambda y: np.array([np.exp(0.3 * np.sum(y)), y[0] * y[1] - y[2] ** 2]))

    # (9) Lagrange, example/0 as shipped (N=8, every dim has >= 2 nodes, 1000 columns)
    aniso = lambda N: 2 ** (np.linspace(0, N - 1, N))  # noqa
    Lam = mis.aniso_smolyak_mid_set(w=4, N=8, a=aniso(8))
    wfun = lambda N: np.power(np.linspace(1, N, N), -2)  # noqa
    f0 = lambda y: np.sin(np.dot(wfun(y.size), y))  # noqa
    x = rng.uniform(-1, 1, (24, 1000))
    interp_case("lag_example0", Lam, n1d.cc_nodes, lambda n: np.where(0, 1, 2 ** n + 1),
                "lagrange", x, f=f0)
    # same index set with pw-linear (BASELINE config 1 wording)
    interp_case("lin_example0", Lam, n1d.cc_nodes, lambda n: np.where(0, 1, 2 ** n + 1),
                "linear", x, f=f0)

    # (10) Lagrange with inactive dims, NF=2, Gauss-Hermite knots (unbounded)
    Lam = mis.smolyak_mid_set(3, 3)
    x = rng.normal(0, 1, (40, 3))
    # (hermite_nodes(1) is a 1-element array, so the active zero multi-index is fine)
    interp_case("lag_hermite", Lam, n1d.hermite_nodes, lambda n: n + 1, "lagrange", x,
                f=lambda y: np.array([np.exp(-0.1 * y @ y), y[0] + y[1] * y[2]]))

    # (11) example/1 as shipped (N=10, pw-linear, unbounded nested nodes, 1000 columns)
    Lam = mis.aniso_smolyak_mid_set(w=5, N=10, a=aniso(10))
    x = rng.normal(0, 1, (32, 1000))
    interp_case("lin_example1", Lam, n1d.unbounded_nodes_nested,
                lambda n: 2 ** (n + 1) - 1, "linear", x, f=f0)

    # (12) config-2 shape, small: smolyak(2, 16), unbounded nested, random node values
    Lam = mis.smolyak_mid_set(2, 16)
    x = rng.normal(0, 1, (32, 16))
    it = SGInterpolant(Lam, n1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1,
                       verbose=False)
    fsg = np.random.default_rng(1).normal(size=(it.num_nodes, 1))
    interp_case("lin_c2_small", Lam, n1d.unbounded_nodes_nested,
                lambda n: 2 ** (n + 1) - 1, "linear", x, f_on_SG=fsg)

    # (13) Gaussian weighted-Leja knots shipped by the reference as .mat data (config 2
    #      wording), Lagrange operator, lev2knots = nu + 1 (nested for every n)
    from scipy.io import loadmat
    X = loadmat("/root/reference/sgmethods/knots_weighted_leja_guass.mat")["X"].reshape(-1)
    leja = lambda n: np.sort(X[:int(n)])  # noqa: E731
    Lam = mis.smolyak_mid_set(4, 4)
    x = rng.normal(0, 1, (48, 4))
    interp_case("lag_leja_gauss", Lam, leja, lambda nu: nu + 1, "lagrange", x,
                f=lambda y: np.array([np.exp(-0.2 * y @ y) * np.cos(y[0])]),
                extra=dict(leja_X=X[:12]))

    # ---- sample_on_SG recycling (sparse_grid_interpolant.py:141-250) --------------------
    l2k = lambda n: np.where(n == 0, 1, 2 ** n + 1)  # noqa
    fvec = lambda y: np.array([np.sin(np.sum(y)), np.sum(y * y)])  # noqa
    small = SGInterpolant(mis.smolyak_mid_set(2, 2), n1d.cc_nodes, l2k, verbose=False)
    s_small = small.sample_on_SG(fvec)
    big = SGInterpolant(mis.smolyak_mid_set(3, 3), n1d.cc_nodes, l2k, verbose=False)
    calls = []

    def f_count(y):
        calls.append(np.array(y))
        return fvec(y)
    s_big = big.sample_on_SG(f_count, old_xx=small.SG, old_samples=s_small)
    # restriction to a coarser grid of the SAME dimension (what MLInterpolant relies on)
    mid3 = SGInterpolant(mis.smolyak_mid_set(2, 3), n1d.cc_nodes, l2k, verbose=False)
    back = mid3.sample_on_SG(lambda y: 1 / 0, old_xx=big.SG, old_samples=s_big)
    np.savez_compressed(os.path.join(HERE, "recycle.npz"), small_SG=small.SG,
                        s_small=s_small, big_SG=big.SG, s_big=s_big,
                        n_new=np.int64(len(calls)), mid3_SG=mid3.SG, back=back)

    # ---- index sets -----------------------------------------------------------------------
    d = {}
    d["tp_2_3"] = mis.tensor_product_mid_set(2, 3)
    d["tp_0_4"] = mis.tensor_product_mid_set(0, 4)
    d["tp_3_1"] = mis.tensor_product_mid_set(3, 1)
    d["smolyak_4_5"] = mis.smolyak_mid_set(4, 5)
    d["smolyak_0_3"] = mis.smolyak_mid_set(0, 3)
    d["smolyak_3_1"] = mis.smolyak_mid_set(3, 1)
    d["aniso_a"] = mis.aniso_smolyak_mid_set(3, 2, [1, 0.5])
    d["aniso_b"] = mis.aniso_smolyak_mid_set(4.5, 4, np.array([0.7, 1.1, 1.3, 2.9]))
    d["aniso_c"] = mis.aniso_smolyak_mid_set(4, 8, aniso(8))
    d["aniso_d"] = mis.aniso_smolyak_mid_set(0.9, 3, [0.3, 0.3, 0.1])   # float budget order
    d["aniso_free"] = mis.aniso_smolyak_mid_set_free_N(5, lambda N: np.linspace(1, N, N))
    for it_ in (0, 2, 5):
        d[f"profit_fast_{it_}"] = mis.compute_mid_set_fast(Profit, 11. ** (-it_), N=4)
    d["profit_free"] = mis.mid_set_profit_free_dim(
        lambda nu: 2. ** (-np.dot(nu, np.linspace(1, nu.size, nu.size) ** 2))
        if np.ndim(nu) else 2. ** (-nu), 1.e-2)
    d["profit_free_b"] = mis.mid_set_profit_free_dim(
        lambda nu: np.prod(np.power(np.arange(1, np.size(nu) + 1, dtype=float) + 1.0,
                                    -np.atleast_1d(nu).astype(float))), 1.e-2)
    # C3 anisotropy (LC levels) small instance
    a_lc = np.array([rutils.compute_level_lc(n)[0] + 1 for n in range(16)], dtype=float)
    d["aniso_lc16_w4"] = mis.aniso_smolyak_mid_set(4, 16, a_lc)
    d["lexic_in"] = rng.integers(0, 4, (30, 4))
    d["lexic_out"] = rutils.lexic_sort(d["lexic_in"])
    d["level_lc"] = np.array([rutils.compute_level_lc(i) for i in range(40)])
    np.savez_compressed(os.path.join(HERE, "index_sets.npz"), **d)
    print("index_sets:", {k: v.shape for k, v in d.items()})

    # ---- MidSet (mid_set.py) ----------------------------------------------------------------
    d = {}
    ms = MidSet(track_reduced_margin=True)
    script = [("e", 0), ("e", 0), ("d", None), ("e", 1), ("e", 1), ("d", None),   # ref test
              ("e", 3), ("e", 7), ("e", 2), ("d", None), ("e", 10), ("e", 5), ("e", 0)]
    for step, (op, arg) in enumerate(script):
        if op == "e":
            ms.enlarge_from_margin(arg)
        else:
            ms.increase_dim()
        d[f"mid_{step}"] = ms.mid_set.copy()
        d[f"margin_{step}"] = ms.margin.copy()
        d[f"rm_{step}"] = ms.reduced_margin.copy()
    d["script_ops"] = np.array([s[0] for s in script])
    d["script_args"] = np.array([-1 if s[1] is None else s[1] for s in script])
    # recursion path: add a margin index that is NOT in the reduced margin
    ms2 = MidSet(track_reduced_margin=True)
    ms2.enlarge_from_margin(0)
    ms2.increase_dim()
    ms2.enlarge_from_margin(ms2.get_cardinality_margin() - 1)
    idx = int(np.where(np.all(ms2.margin == np.array([2, 1]), axis=1))[0][0])
    d["rec_idx"] = np.int64(idx)
    ms2.enlarge_from_margin(idx)
    d["rec_mid"], d["rec_margin"], d["rec_rm"] = ms2.mid_set, ms2.margin, ms2.reduced_margin
    np.savez_compressed(os.path.join(HERE, "midset.npz"), **d)

    # ---- 1-D knot families (nodes_1d.py) -------------------------------------------------------
    d = {}
    for n in (1, 2, 3, 5, 9, 17, 33):
        d[f"equi_{n}"] = np.asarray(n1d.equispaced_nodes(n), dtype=float)
        d[f"equi_int_{n}"] = np.asarray(n1d.equispaced_interior_nodes(n), dtype=float)
        d[f"cc_{n}"] = np.asarray(n1d.cc_nodes(n), dtype=float)
    for n in (1, 3, 7, 15, 31, 63):
        d[f"optg_{n}"] = n1d.optimal_gaussian_nodes(n)
        d[f"optg3_{n}"] = n1d.optimal_gaussian_nodes(n, p=3)
        d[f"unb_{n}"] = n1d.unbounded_nodes_nested(n)
    for n in (1, 2, 3, 4, 5, 8):
        d[f"herm_{n}"] = n1d.hermite_nodes(n)
    np.savez_compressed(os.path.join(HERE, "nodes_1d.npz"), **d)

    # ---- Brownian-motion expansions ---------------------------------------------------------------
    d = {}
    tt = np.linspace(0, 1, 1001)
    yy = rng.normal(size=(6, 64))
    d["lc64_tt"], d["lc64_yy"] = tt, yy
    d["lc64_W"] = np.array([param_LC_Brownian_motion(tt, y, 1) for y in yy])
    tt2 = np.sort(rng.uniform(0, 2.0, 257))
    tt2[0], tt2[-1], tt2[10], tt2[11] = 0.0, 2.0, 0.5, 1.0     # dyadic ties
    yy2 = rng.normal(size=(5, 5))                          # d=5: zero padding to 8
    d["lc5_tt"], d["lc5_yy"] = tt2, yy2
    d["lc5_W"] = np.array([param_LC_Brownian_motion(tt2, y, 2.0) for y in yy2])
    yy3 = rng.normal(size=(3, 1))                          # d=1: L=0
    d["lc1_yy"] = yy3
    d["lc1_W"] = np.array([param_LC_Brownian_motion(tt2, y, 2.0) for y in yy3])
    yk = rng.normal(size=(4, 12))
    d["kl_tt"], d["kl_yy"] = tt, yk
    d["kl_W"] = np.array([param_KL_Brownian_motion(tt, y) for y in yk])
    np.savez_compressed(os.path.join(HERE, "brownian.npz"), **d)
    print("done")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "drivers":
        drivers()
    else:
        main()
        drivers()
