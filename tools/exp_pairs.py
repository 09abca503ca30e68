"""Experiment: wide column kernel with one or two points per warp as a function of the number
of points (sharing of rows between the two points of a warp needs well-filled locality bins).
Dev tool, not part of the product."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
from sgmethods_b200.tp_inteprolants import TPPwQuadraticInterpolator
from sgmethods_b200.utils import compute_level_lc
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


a = np.array([compute_level_lc(n)[0] + 1 for n in range(64)], dtype=float)
c3 = SGInterpolant(mis.aniso_smolyak_mid_set(10, 64, a), lambda n: nodes_1d.optimal_gaussian_nodes(n, p=2),
                   lambda n: 2 ** (n + 1) - 1, verbose=False)
c4 = SGInterpolant(mis.smolyak_mid_set(3, 8), nodes_1d.cc_nodes, lambda n: np.where(n == 0, 1, 2 ** n + 1),
                   tp_interpolant=TPPwQuadraticInterpolator, verbose=False)
for name, it, NF, sampler in (("c3 w=10", c3, 1024, "normal"), ("c4 lvl3", c4, 1024, "uniform")):
    plan = it.device_plan(0)
    gen = torch.Generator(device=dev).manual_seed(0)
    f = torch.randn((it.num_nodes, NF), dtype=torch.float64, device=dev, generator=gen)
    for P in (4096, 16384, 65536, 100000, 262144):
        x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen) if sampler == "normal" \
            else torch.rand((P, it.N), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
        out = torch.empty((P, NF), dtype=torch.float64, device=dev)
        res = {}
        for pairs in ("0", "1"):
            os.environ["SGM_WIDE_PAIRS"] = pairs
            res[pairs] = timed(lambda: plan.eval(x, f, out))
        print(f"{name} P={P} NF={NF}: one point/warp {res['0']:.2f} ms, two points/warp {res['1']:.2f} ms "
              f"(x{res['0'] / res['1']:.2f})", flush=True)
