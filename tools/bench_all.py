"""Dev/measurement tool: device-resident timings of every kernel family on the named
workload shapes, with algorithmic bytes, HBM fraction and the
secondary corner/FP64 rates.  Writes one JSON line per case."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import algorithmic_bytes, measured_peak_gbs
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
from sgmethods_b200.tp_inteprolants import (TPLagrangeInterpolator, TPPwCubicInterpolator,
                                            TPPwQuadraticInterpolator)
from sgmethods_b200.multilevel_interpolant import MLInterpolant
from sgmethods_b200.parametric_expansions_brownian_motion import lc_brownian_device
from sgmethods_b200.utils import compute_level_lc

dev = torch.device("cuda")
peak, _ = measured_peak_gbs()
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def run_interp(name, it, P, NF, sampler="normal", reps=5):
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen) if sampler == "normal" \
        else torch.rand((P, it.N), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    f = torch.randn((it.num_nodes, NF), dtype=torch.float64, device=dev, generator=gen)
    plan = it.device_plan(0)
    out = torch.empty((P, NF), dtype=torch.float64, device=dev)
    ms = timed(lambda: plan.eval(x, f, out), reps)
    b = algorithmic_bytes(it, P, NF)
    corners = plan.corners_per_point
    print(json.dumps({"case": name, "P": P, "NF": NF, "N": it.N, "terms": len(it._coeffs),
                      "nodes": it.num_nodes, "corners_per_point": corners, "ms": ms,
                      "evals_per_s": P * NF / ms * 1e3, "bytes_alg": b,
                      "hbm_gbs": b / ms / 1e6, "hbm_frac": b / ms / 1e6 / peak,
                      "corner_madds_per_s": P * NF * corners / ms * 1e3}), flush=True)


which = sys.argv[1:] or ["c1", "c1lin", "ex1", "c2", "c3", "c4", "lc", "c5"]
l2 = lambda n: 2 ** (n + 1) - 1  # noqa
dbl = lambda n: np.where(n == 0, 1, 2 ** n + 1)  # noqa
for w in which:
    t0 = time.time()
    if w in ("c1", "c1lin"):
        S = mis.aniso_smolyak_mid_set(4, 8, 2 ** (np.linspace(0, 7, 8)))
        it = SGInterpolant(S, nodes_1d.cc_nodes, lambda n: np.where(0, 1, 2 ** n + 1),
                           tp_interpolant=TPLagrangeInterpolator if w == "c1" else
                           SGInterpolant.__init__.__defaults__[0], verbose=False)
        for P in (256, 10 ** 5):
            run_interp(f"{w} example/0 ({'Lagrange' if w == 'c1' else 'pw-linear'})", it, P, 1, "uniform")
    elif w == "ex1":
        S = mis.aniso_smolyak_mid_set(5, 10, 2 ** (np.linspace(0, 9, 10)))
        it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, l2, verbose=False)
        for P in (256, 10 ** 7):
            run_interp("example/1 as shipped (pw-linear, 7 terms)", it, P, 1)
    elif w == "c2":
        it = SGInterpolant(mis.smolyak_mid_set(4, 16), nodes_1d.unbounded_nodes_nested, l2, verbose=False)
        run_interp("c2 d=16 w=4 pw-linear", it, 10 ** 6, 1)
    elif w == "c2hier":
        it = SGInterpolant(mis.smolyak_mid_set(4, 16), nodes_1d.unbounded_nodes_nested, l2, verbose=False)
        t1 = time.time(); hp = it.hierarchical(); t_build = time.time() - t1
        gen = torch.Generator(device=dev).manual_seed(0)
        P = 10 ** 6
        x = torch.randn((P, 16), dtype=torch.float64, device=dev, generator=gen)
        f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
        exact = it.interpolate(x, f)
        fast = hp.interpolate(x, f)
        dev_rel = float((fast - exact).abs().max() / exact.abs().max())
        ms_all = timed(lambda: hp.interpolate(x, f))
        alpha = hp.surpluses(f)
        plan = hp._dev(0)[0]
        out = torch.empty((P, 1), dtype=torch.float64, device=dev)
        ms_eval = timed(lambda: plan.eval(x, alpha, out))
        print(json.dumps({"case": "c2 hierarchical-surplus form (opt-in)", "P": P, "host_build_s": t_build,
                          "ms_total": ms_all, "ms_eval_only": ms_eval, "evals_per_s": P / ms_all * 1e3,
                          "max_rel_dev_from_exact_path": dev_rel}), flush=True)
    elif w == "c3hier":
        a = np.array([compute_level_lc(n)[0] + 1 for n in range(64)], dtype=float)
        it = SGInterpolant(mis.aniso_smolyak_mid_set(10, 64, a),
                           lambda n: nodes_1d.optimal_gaussian_nodes(n, p=2), l2, verbose=False)
        t1 = time.time(); hp = it.hierarchical(); t_build = time.time() - t1
        gen = torch.Generator(device=dev).manual_seed(0)
        P, NF = 10 ** 4, 10 ** 4
        x = torch.randn((P, 64), dtype=torch.float64, device=dev, generator=gen)
        f = torch.randn((it.num_nodes, NF), dtype=torch.float64, device=dev, generator=gen)
        exact = it.interpolate(x, f)
        fast = hp.interpolate(x, f)
        dev_rel = float((fast - exact).abs().max() / exact.abs().max())
        ms_exact = timed(lambda: it.interpolate(x, f), 3)
        ms_all = timed(lambda: hp.interpolate(x, f), 3)
        ms_hier = timed(lambda: hp.surpluses(f), 3)
        print(json.dumps({"case": "c3 (w=10) hierarchical-surplus form (opt-in)", "P": P, "NF": NF,
                          "host_build_s": t_build, "ms_exact": ms_exact, "ms_total": ms_all,
                          "ms_hierarchisation": ms_hier, "evals_per_s": P * NF / ms_all * 1e3,
                          "max_rel_dev_from_exact_path": dev_rel}), flush=True)
    elif w == "c3":
        a = np.array([compute_level_lc(n)[0] + 1 for n in range(64)], dtype=float)
        for wlev, P, NF in ((8, 10 ** 4, 10 ** 4), (10, 10 ** 4, 10 ** 4)):
            it = SGInterpolant(mis.aniso_smolyak_mid_set(wlev, 64, a),
                               lambda n: nodes_1d.optimal_gaussian_nodes(n, p=2), l2, verbose=False)
            run_interp(f"c3 LC-aniso d=64 w={wlev} pw-linear", it, P, NF, reps=3)
    elif w == "c4":
        seq = [SGInterpolant(mis.smolyak_mid_set(k, 8), nodes_1d.cc_nodes, dbl,
                             tp_interpolant=TPPwQuadraticInterpolator, verbose=False) for k in range(4)]
        run_interp("c4 level-3 alone, pw-quadratic d=8", seq[3], 10 ** 5, 10 ** 3, "uniform", reps=3)
        mli = MLInterpolant(seq)
        gen = torch.Generator(device=dev).manual_seed(0)
        P, NF = 10 ** 5, 10 ** 3
        yy = torch.rand((P, 8), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
        ml = [torch.randn((seq[3 - k].num_nodes, NF), dtype=torch.float64, device=dev, generator=gen)
              for k in range(4)]
        ms = timed(lambda: mli.interpolate(yy, ml), 3)
        print(json.dumps({"case": "c4 multilevel 4 levels pw-quadratic NF=1e3 (fused signed sum)",
                          "P": P, "NF": NF, "ms": ms, "evals_per_s": P * NF / ms * 1e3,
                          "hbm_gbs": (8 * P * 8 + 8 * P * NF) / ms / 1e6}), flush=True)
    elif w == "c3full":       # BASELINE config 3 at its full size: 10^5 points x 10^4 outputs
        a = np.array([compute_level_lc(n)[0] + 1 for n in range(64)], dtype=float)
        it = SGInterpolant(mis.aniso_smolyak_mid_set(10, 64, a),
                           lambda n: nodes_1d.optimal_gaussian_nodes(n, p=2), l2, verbose=False)
        run_interp("c3 FULL SIZE LC-aniso d=64 w=10 pw-linear", it, 10 ** 5, 10 ** 4, reps=2)
    elif w == "c4full":       # BASELINE config 4 at its full size: 4 levels, 10^6 points x 10^3 outputs
        seq = [SGInterpolant(mis.smolyak_mid_set(k, 8), nodes_1d.cc_nodes, dbl,
                             tp_interpolant=TPPwQuadraticInterpolator, verbose=False) for k in range(4)]
        mli = MLInterpolant(seq)
        gen = torch.Generator(device=dev).manual_seed(0)
        P, NF = 10 ** 6, 10 ** 3
        yy = torch.rand((P, 8), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
        ml = [torch.randn((seq[3 - k].num_nodes, NF), dtype=torch.float64, device=dev, generator=gen)
              for k in range(4)]
        ms = timed(lambda: mli.interpolate(yy, ml), 2, 1)
        print(json.dumps({"case": "c4 FULL SIZE multilevel 4 levels pw-quadratic NF=1e3 (fused signed sum)",
                          "P": P, "NF": NF, "ms": ms, "evals_per_s": P * NF / ms * 1e3,
                          "hbm_gbs": (8 * P * 8 + 8 * P * NF) / ms / 1e6}), flush=True)
    elif w == "lc":
        for P, d, nt in ((10 ** 5, 64, 10 ** 3), (10 ** 4, 64, 10 ** 4)):
            gen = torch.Generator(device=dev).manual_seed(0)
            yy = torch.randn((P, d), dtype=torch.float64, device=dev, generator=gen)
            tt = torch.linspace(0, 1, nt, dtype=torch.float64, device=dev)
            Wb = torch.empty((P, nt), dtype=torch.float64, device=dev)
            ms = timed(lambda: lc_brownian_device(tt, yy, 1.0, Wb))
            b = 8 * P * d + 8 * nt + 8 * P * nt
            print(json.dumps({"case": f"LC d={d} P={P} nt={nt}", "ms": ms, "bytes_alg": b,
                              "hbm_gbs": b / ms / 1e6, "hbm_frac": b / ms / 1e6 / peak,
                              "outputs_per_s": P * nt / ms * 1e3}), flush=True)
    elif w == "c5":
        S = mis.aniso_smolyak_mid_set(3, 128, np.array([1.0] * 80 + [3.0] * 48))
        print("c5 set", S.shape, time.time() - t0, flush=True)
        it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, l2, verbose=False)
        print("c5 ctor", time.time() - t0, len(it._coeffs), it.num_nodes, it._maps.size, flush=True)
        run_interp("c5 d=128 ~1e5 terms pw-linear", it, 10 ** 5, 1, reps=2)
