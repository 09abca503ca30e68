"""GPU fuzz of the exact path (method='combination'): bit-equal to the CPU oracle on small
batches for all four operators, and independent of the batch a point is evaluated in (large
sorted batches with dynamic tiles against slices of them)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_hierarchical_cpu import _random_downward_closed_set
from oracle import sg_oracle
from sgmethods_b200 import nodes_1d, tp_inteprolants as tpi
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
t_end = time.time() + budget
n = bad = 0
KINDS = {"linear": (tpi.TPPwLinearInterpolator, nodes_1d.unbounded_nodes_nested, lambda q: 2 ** (q + 1) - 1),
         "quadratic": (tpi.TPPwQuadraticInterpolator, nodes_1d.cc_nodes, lambda q: np.where(q == 0, 1, 2 ** q + 1)),
         "cubic": (tpi.TPPwCubicInterpolator, nodes_1d.equispaced_nodes, lambda q: 3 * q + 4),
         "lagrange": (tpi.TPLagrangeInterpolator, nodes_1d.cc_nodes, lambda q: np.where(q == 0, 1, 2 ** q + 1))}
while time.time() < t_end:
    rng = np.random.default_rng(7000 + n)
    kind = list(KINDS)[n % 4]
    op, knots, lev = KINDS[kind]
    N = int(rng.integers(1, 7))
    S = _random_downward_closed_set(rng, N, 3, int(rng.integers(2, 50)))
    try:
        it = SGInterpolant(S, knots, lev, tp_interpolant=op, verbose=False)
        NF = int(rng.choice([1, 3, 40, 130]))
        f = rng.normal(size=(it.num_nodes, NF))
        lo = -1.0 if kind in ("quadratic", "lagrange") else -1.5
        x = rng.uniform(lo, 1.0, (int(rng.choice([1, 31, 97, 300])), N))
        plan = sg_oracle.plan_from_tables(it.N, it.combination_coeffs, it.active_tp_dims,
                                          it.active_tp_nodes_list, it.map_tp_to_SG)
        ref = sg_oracle.interpolate(plan, x, f, kind=kind, squeeze=False)
        got = np.asarray(it.interpolate(x, f, method="combination")).reshape(ref.shape)
        assert np.array_equal(got, ref, equal_nan=True), float(np.nanmax(np.abs(got - ref)))
        P = int(rng.choice([5000, 40000, 70000]))
        xb = torch.from_numpy(rng.uniform(lo, 1.0, (P, N))).cuda()
        fb = torch.from_numpy(f[:, :1].copy()).cuda()
        full = it.interpolate(xb, fb, method="combination").reshape(-1)
        k = int(rng.integers(1, P))
        part = it.interpolate(xb[:k].contiguous(), fb, method="combination").reshape(-1)
        assert torch.equal(full[:k], part), "batch dependence"
    except Exception as exc:  # noqa: BLE001
        bad += 1
        print(f"CASE {n} kind={kind} N={N} |S|={S.shape[0]}: {type(exc).__name__}: {exc}", flush=True)
    n += 1
print(f"exact fuzz: {n} cases, {bad} failures", flush=True)
