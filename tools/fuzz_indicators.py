"""GPU fuzz of compute_GG_indicators_RM: random adaptive states, method='surplus' (callback norm,
GPU norm, extensions) against method='rebuild' (the reference's sequence)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sgmethods_b200 import nodes_1d, tp_inteprolants as tpi
from sgmethods_b200.compute_error_indicators import MeanSquareNorm, compute_GG_indicators_RM
from sgmethods_b200.mid_set import MidSet
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
t_end = time.time() + budget
n = bad = 0
norm = lambda a, b: float(np.sqrt(np.mean(np.sum((np.asarray(a).reshape(len(a), -1) - np.asarray(b).reshape(len(b), -1)) ** 2, axis=1))))  # noqa: E731
while time.time() < t_end:
    rng = np.random.default_rng(500 + n)
    kind = ("linear", "quadratic")[n % 2]
    if kind == "linear":
        op, knots, lev = tpi.TPPwLinearInterpolator, nodes_1d.unbounded_nodes_nested, (lambda q: 2 ** (q + 1) - 1)
    else:
        op, knots, lev = tpi.TPPwQuadraticInterpolator, nodes_1d.cc_nodes, (lambda q: np.where(q == 0, 1, 2 ** q + 1))
    ms = MidSet(track_reduced_margin=True)
    ms.enlarge_from_margin(0)
    for step in range(int(rng.integers(2, 14))):
        if ms.get_dim() < 5 and rng.random() < 0.3:
            ms.increase_dim()
        else:
            ok = np.flatnonzero(ms.margin.max(axis=1) <= 5)        # keep the 1-D levels moderate
            if ok.size:
                ms.enlarge_from_margin(int(rng.choice(ok)))
    N = ms.get_dim()
    NF = int(rng.choice([1, 2, 5]))
    w = rng.uniform(0.3, 1.0, (NF, N))
    f = lambda y: np.sin(w[:, :y.size] @ y) + 0.1 * np.arange(NF)  # noqa: E731
    try:
        cur = SGInterpolant(ms.mid_set, knots, lev, tp_interpolant=op, verbose=False)
        u_on_SG = cur.sample_on_SG(f)
        yy = rng.normal(0, 1.0, (700, N)) if kind == "linear" else rng.uniform(-1, 1, (700, N))
        u_interp = np.asarray(cur.interpolate(yy, u_on_SG)).reshape(700, NF)
        empty_rm, empty_est = np.zeros((0, 0), dtype=int), np.zeros(0)
        args = (empty_rm, empty_est, ms, knots, lev, f, cur.SG, u_on_SG, yy)
        ref = compute_GG_indicators_RM(*args, norm, u_interp, TP_inteporlant=op, method="rebuild")
        ext = {}
        got = compute_GG_indicators_RM(*args, norm, u_interp, TP_inteporlant=op, extensions=ext)
        gpu = compute_GG_indicators_RM(*args, MeanSquareNorm(), u_interp, TP_inteporlant=op)
        scale = max(1.0, float(np.max(np.abs(u_interp))))
        assert np.max(np.abs(got - ref)) <= 1e-11 * scale, ("surplus", float(np.max(np.abs(got - ref))))
        assert np.max(np.abs(gpu - ref)) <= 1e-11 * scale, ("gpu norm", float(np.max(np.abs(gpu - ref))))
        assert sorted(ext) == list(range(ms.get_cardinality_reduced_margin()))
    except Exception as exc:  # noqa: BLE001
        bad += 1
        print(f"CASE {n} kind={kind} N={N} card={ms.get_cardinality_mid_set()}: {type(exc).__name__}: {exc}", flush=True)
    n += 1
print(f"indicator fuzz: {n} cases, {bad} failures", flush=True)
