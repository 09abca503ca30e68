import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
from sgmethods_b200.utils import compute_level_lc
a = np.array([compute_level_lc(n)[0] + 1 for n in range(64)], dtype=float)
it = SGInterpolant(mis.aniso_smolyak_mid_set(8, 64, a), lambda n: nodes_1d.optimal_gaussian_nodes(n, p=2),
                   lambda n: 2 ** (n + 1) - 1, verbose=False)
dev = torch.device("cuda"); P, NF = 4000, 10000
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
f = torch.randn((it.num_nodes, NF), dtype=torch.float64, device=dev, generator=gen)
plan = it.device_plan(0); out = torch.empty((P, NF), dtype=torch.float64, device=dev)
for _ in range(3): plan.eval(x, f, out)
torch.cuda.synchronize(); print("ok", plan.corners_per_point, len(it._coeffs), it.num_nodes)
