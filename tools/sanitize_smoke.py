"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
from sgmethods_b200.tp_inteprolants import (TPLagrangeInterpolator, TPPwCubicInterpolator,
                                            TPPwQuadraticInterpolator, TPPwLinearInterpolator)
from sgmethods_b200.multilevel_interpolant import MLInterpolant
from sgmethods_b200.parametric_expansions_brownian_motion import param_LC_Brownian_motion, param_KL_Brownian_motion
rng = np.random.default_rng(0)
dbl = lambda n: np.where(n == 0, 1, 2 ** n + 1)
cases = [(TPPwLinearInterpolator, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, 6, "normal"),
         (TPPwQuadraticInterpolator, nodes_1d.cc_nodes, dbl, 4, "uniform"),
         (TPPwCubicInterpolator, nodes_1d.equispaced_nodes, lambda n: 3 * n + 1, 4, "normal"),
         (TPLagrangeInterpolator, nodes_1d.hermite_nodes, lambda n: 2 * n + 1, 4, "normal")]
for cls, knots, lev, N, smp in cases:
    it = SGInterpolant(mis.smolyak_mid_set(3, N), knots, lev, tp_interpolant=cls, verbose=False)
    for P in (33, 40000):            # split kernel, and (>= 32768) the locality sort
        x = rng.normal(size=(P, N)) if smp == "normal" else rng.uniform(-1, 1, (P, N))
        for NF in (1, 5, 300):
            if P > 1000 and NF > 5:
                continue
            out = it.interpolate(x, rng.normal(size=(it.num_nodes, NF)))
            assert np.all(np.isfinite(out))
    print(cls.__name__, "ok")
# prefix kernel with 4-direction terms, several chunks, ragged last chunk; and the split kernel
it = SGInterpolant(mis.smolyak_mid_set(4, 6), nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False)
for P in (45, 33000):
    a = it.interpolate(rng.normal(size=(P, 6)), rng.normal(size=(it.num_nodes, 1)))
os.environ["SGM_FORCE_KERNEL"] = "split"
it.interpolate(rng.normal(size=(45, 6)), rng.normal(size=(it.num_nodes, 1)))
del os.environ["SGM_FORCE_KERNEL"]
# few-term plan with many points -> thread-per-point kernel
it = SGInterpolant(mis.smolyak_mid_set(1, 3), nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False)
it.interpolate(rng.normal(size=(20000, 3)), rng.normal(size=(it.num_nodes, 1)))
# wide terms (k = 6 > 4) -> wide-linear flavour
it = SGInterpolant(mis.tensor_product_mid_set(1, 6), nodes_1d.cc_nodes, lambda n: 2 * n + 1, verbose=False)
it.interpolate(rng.uniform(-1, 1, (100, 6)), rng.normal(size=(it.num_nodes, 1)))
seq = [SGInterpolant(mis.smolyak_mid_set(k, 3), nodes_1d.cc_nodes, dbl, verbose=False) for k in range(3)]
ml = [rng.normal(size=(seq[2 - k].num_nodes, 130)) for k in range(3)]
MLInterpolant(seq).interpolate(rng.uniform(-1, 1, (50, 3)), ml)
param_LC_Brownian_motion(np.linspace(0, 1, 333), rng.normal(size=(7, 20)), 1.0)
param_LC_Brownian_motion(np.linspace(0, 1, 17), rng.normal(size=(3, 5000)), 1.0)      # L = 13: generic kernel
param_KL_Brownian_motion(np.linspace(0, 1, 50), rng.normal(size=(4, 9)))
torch.cuda.synchronize()
print("all ok")
