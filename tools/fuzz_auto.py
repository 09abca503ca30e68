"""GPU fuzz: random operators / downward closed sets / batch shapes; method='auto' and
'hierarchical' against 'combination' (1e-12, NaN masks equal), numpy / CUDA / pinned inputs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_hierarchical_cpu import _random_downward_closed_set
from sgmethods_b200 import nodes_1d, tp_inteprolants as tpi
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t_end = time.time() + budget
n = bad = 0
while time.time() < t_end:
    rng = np.random.default_rng(seed0 + n)
    kind = ("linear", "quadratic", "lagrange")[n % 3]
    N = int(rng.integers(1, 9))
    S = _random_downward_closed_set(rng, N, 4 if kind != "lagrange" else 3, int(rng.integers(2, 80)))
    if kind == "linear":
        op, knots, lev = tpi.TPPwLinearInterpolator, nodes_1d.unbounded_nodes_nested, (lambda q: 2 ** (q + 1) - 1)
    else:
        op = tpi.TPPwQuadraticInterpolator if kind == "quadratic" else tpi.TPLagrangeInterpolator
        knots, lev = nodes_1d.cc_nodes, (lambda q: np.where(q == 0, 1, 2 ** q + 1))
    it = SGInterpolant(S, knots, lev, tp_interpolant=op, verbose=False)
    P = int(rng.choice([1, 33, 4095, 8191, 8192, 8193, 20000, 470000]))
    NF = int(rng.choice([1, 2, 17, 64, 129, 257, 300, 512, 600, 1030]))
    if P * NF > 4e7:
        NF = 1
    extra = int(rng.integers(0, 3))
    x = rng.normal(0, 1.2, (P, N + extra)) if kind == "linear" else rng.uniform(-1.1 if kind == "quadratic" else -1, 1, (P, N + extra))
    f = rng.normal(size=(it.num_nodes, NF))
    mode = int(rng.integers(0, 3))
    xi = x if mode == 0 else (torch.from_numpy(x).cuda() if mode == 1 else torch.from_numpy(x).pin_memory())
    fi = f if mode == 0 else torch.from_numpy(f).cuda() if mode == 1 else torch.from_numpy(f)
    try:
        ref = it.interpolate(xi, fi, method="combination")
        outs = {"auto": it.interpolate(xi, fi)}
        try:
            outs["hierarchical"] = it.interpolate(xi, fi, method="hierarchical")
        except NotImplementedError:
            pass
        ref = ref.cpu().numpy() if isinstance(ref, torch.Tensor) else ref
        for name, o in outs.items():
            o = o.cpu().numpy() if isinstance(o, torch.Tensor) else o
            assert o.shape == ref.shape, (o.shape, ref.shape)
            assert np.array_equal(np.isnan(o), np.isnan(ref))
            ok = ~np.isnan(ref)
            scale = np.max(np.abs(ref[ok])) if ok.any() else 1.0
            dev = np.max(np.abs(o[ok] - ref[ok])) / (scale or 1.0) if ok.any() else 0.0
            assert dev <= 1e-12, (name, dev)
    except Exception as exc:  # noqa: BLE001
        bad += 1
        print(f"CASE {seed0 + n} kind={kind} N={N} |S|={S.shape[0]} P={P} NF={NF} mode={mode}: {type(exc).__name__}: {exc}", flush=True)
    n += 1
print(f"fuzz: {n} cases, {bad} failures", flush=True)

# ---- multilevel interpolants: method='auto' / 'hierarchical' against 'combination'
from sgmethods_b200 import multi_index_sets as mis  # noqa: E402
from sgmethods_b200.multilevel_interpolant import MLInterpolant  # noqa: E402
t_end = time.time() + budget / 4
m = mbad = 0
while time.time() < t_end:
    rng = np.random.default_rng(90000 + seed0 + m)
    kind = ("linear", "quadratic")[m % 2]
    N, levels = int(rng.integers(1, 6)), int(rng.integers(2, 5))
    op = tpi.TPPwLinearInterpolator if kind == "linear" else tpi.TPPwQuadraticInterpolator
    lev = lambda q: np.where(q == 0, 1, 2 ** q + 1)  # noqa: E731
    aniso = rng.uniform(1.0, 2.5, N)
    seq = [SGInterpolant(mis.aniso_smolyak_mid_set(k + int(rng.integers(0, 2)) * 0, N, aniso), nodes_1d.cc_nodes, lev,
                         tp_interpolant=op, verbose=False) for k in range(levels)]
    NF = int(rng.choice([1, 3, 40, 260]))
    P = int(rng.choice([500, 9000, 20000]))
    ml = [rng.normal(size=(seq[levels - 1 - k].num_nodes, NF)) for k in range(levels)]
    yy = rng.uniform(-1.1 if kind == "quadratic" else -1.3, 1.0, (P, N))
    try:
        mli = MLInterpolant(seq)
        ref = np.asarray(mli.interpolate(yy, ml, method="combination"))
        for method in ("auto", "hierarchical"):
            got = np.asarray(mli.interpolate(yy, ml, method=method))
            assert got.shape == ref.shape
            dev = np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300)
            assert dev <= 1e-12, (method, dev)
    except Exception as exc:  # noqa: BLE001
        mbad += 1
        print(f"ML CASE {m} kind={kind} N={N} levels={levels} P={P} NF={NF}: {type(exc).__name__}: {exc}", flush=True)
    m += 1
print(f"multilevel fuzz: {m} cases, {mbad} failures", flush=True)
