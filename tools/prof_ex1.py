import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
S = mis.aniso_smolyak_mid_set(5, 10, 2 ** (np.linspace(0, 9, 10)))
it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False)
dev = torch.device("cuda"); P = 10 ** 7
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
plan = it.device_plan(0); out = torch.empty((P, 1), dtype=torch.float64, device=dev)
for _ in range(3): plan.eval(x, f, out)
torch.cuda.synchronize(); print("ok")
