"""Two-level locality key (plans with fewer than 16 keyed directions) on/off: run twice,
python tools/exp_key2.py; SGM_NO_KEY2=1 python tools/exp_key2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
from tools.exp_hier_ops import timed
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(2)
for w, N in ((6, 8), (5, 10), (8, 4)):
    it = SGInterpolant(mis.smolyak_mid_set(w, N), nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False)
    x = torch.randn((10 ** 6, N), dtype=torch.float64, device=dev, generator=gen)
    f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
    th, _ = timed(lambda: it.interpolate(x, f, method="hierarchical"))
    tc, _ = timed(lambda: it.interpolate(x, f, method="combination"), 2)
    print(f"NO_KEY2={os.environ.get('SGM_NO_KEY2', '0')} smolyak({w},{N}) {len(it._coeffs)} terms: hierarchical {th:.3f} ms, combination {tc:.3f} ms", flush=True)
