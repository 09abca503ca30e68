import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
rng = np.random.default_rng(21)
it = SGInterpolant(mis.smolyak_mid_set(4, 6), nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False)
x = rng.normal(0, 1.5, (500, 6))
for NF in (3, 72, 128, 130, 200, 256, 512):
    f = rng.normal(size=(it.num_nodes, NF))
    a = np.asarray(it.interpolate(x, f)).reshape(500, NF)
    b = np.asarray(it.interpolate(x, f, method="hierarchical")).reshape(500, NF)
    d = np.abs(a - b) / np.abs(a).max()
    print(NF, d.max(), "bad cols:", np.flatnonzero(d.max(axis=0) > 1e-10)[:10], "bad rows:", np.flatnonzero(d.max(axis=1) > 1e-10)[:10])
