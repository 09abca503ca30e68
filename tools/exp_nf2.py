import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
from sgmethods_b200.tp_inteprolants import TPLagrangeInterpolator, TPPwQuadraticInterpolator, TPPwCubicInterpolator
from sgmethods_b200.utils import compute_level_lc
dev = torch.device("cuda")
dbl = lambda n: np.where(n == 0, 1, 2 ** n + 1)
a = np.array([compute_level_lc(n)[0] + 1 for n in range(64)], dtype=float)
cases = {
 "quad d=8 w=3": (SGInterpolant(mis.smolyak_mid_set(3, 8), nodes_1d.cc_nodes, dbl, tp_interpolant=TPPwQuadraticInterpolator, verbose=False), "u", 20000),
 "cubic d=4 w=3": (SGInterpolant(mis.smolyak_mid_set(3, 4), nodes_1d.equispaced_nodes, lambda n: 3 * n + 1, tp_interpolant=TPPwCubicInterpolator, verbose=False), "u", 20000),
 "lagrange ex0": (SGInterpolant(mis.aniso_smolyak_mid_set(4, 8, 2 ** (np.linspace(0, 7, 8))), nodes_1d.cc_nodes, lambda n: np.where(0, 1, 2 ** n + 1), tp_interpolant=TPLagrangeInterpolator, verbose=False), "u", 5000),
 "linear c3 w=8": (SGInterpolant(mis.aniso_smolyak_mid_set(8, 64, a), lambda n: nodes_1d.optimal_gaussian_nodes(n, p=2), lambda n: 2 ** (n + 1) - 1, verbose=False), "n", 20000),
}
gen = torch.Generator(device=dev).manual_seed(0)
for name, (it, smp, P) in cases.items():
    plan = it.device_plan(0)
    x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen) if smp == "n" else torch.rand((P, it.N), dtype=torch.float64, device=dev, generator=gen) * 1.9 - 0.95
    for NF in (1, 8, 32, 100):
        f = torch.randn((it.num_nodes, NF), dtype=torch.float64, device=dev, generator=gen)
        out = torch.empty((P, NF), dtype=torch.float64, device=dev)
        res = []
        for mode in ("0", "100000"):
            os.environ["SGM_COL_LOOP_MAX"] = mode
            for _ in range(2): plan.eval(x, f, out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.eval(x, f, out); plan.eval(x, f, out); e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 2)
        print(f"{name} NF={NF}: column kernels {res[0]:.2f} ms, scalar loop {res[1]:.2f} ms", flush=True)
