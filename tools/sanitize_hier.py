"""Small pass over the round-2 kernels for compute-sanitizer (memcheck / racecheck): grouped
hierarchical kernel (H2) and H1, hierarchical column kernels (1, 2 and 4 points per warp), the
pw-quadratic and Lagrange basis families, cluster hierarchisation, multilevel assembly, segmented
indicator launch, bulk-copy Levy-Ciesielski kernel, plan-specialised kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sgmethods_b200 import multi_index_sets as mis, nodes_1d  # noqa: E402
from sgmethods_b200.multilevel_interpolant import MLInterpolant  # noqa: E402
from sgmethods_b200.parametric_expansions_brownian_motion import lc_brownian_device  # noqa: E402
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant  # noqa: E402
from sgmethods_b200.tp_inteprolants import TPLagrangeInterpolator, TPPwQuadraticInterpolator  # noqa: E402

rng = np.random.default_rng(0)
dbl = lambda n: np.where(n == 0, 1, 2 ** n + 1)  # noqa: E731
l2 = lambda n: 2 ** (n + 1) - 1  # noqa: E731
P = int(os.environ.get("SAN_POINTS", "9000"))

lin = SGInterpolant(mis.smolyak_mid_set(3, 6), nodes_1d.unbounded_nodes_nested, l2, verbose=False)
quad = SGInterpolant(mis.smolyak_mid_set(3, 4), nodes_1d.cc_nodes, dbl, tp_interpolant=TPPwQuadraticInterpolator, verbose=False)
from tests_helper_leja import leja  # noqa: E402
lag = SGInterpolant(mis.smolyak_mid_set(3, 5), leja(), lambda nu: nu + 1, tp_interpolant=TPLagrangeInterpolator, verbose=False)
for name, it, smp in (("linear", lin, "normal"), ("quadratic", quad, "uniform"), ("lagrange", lag, "normal")):
    for NF in (1, 3, 260):
        x = rng.normal(size=(P, it.N)) if smp == "normal" else rng.uniform(-1.1, 1.0, (P, it.N))
        f = rng.normal(size=(it.num_nodes, NF))
        a = np.asarray(it.interpolate(x, f, method="hierarchical")).reshape(P, NF)
        b = np.asarray(it.interpolate(x, f, method="combination")).reshape(P, NF)
        assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b)), name
        it.interpolate(x, f)
        for pts in ("4", "8", "2"):
            os.environ["SGM_WIDE_PTS"] = pts
            it.hierarchical()._dev(0)[0].reload_knobs()
            it.interpolate(x, f, method="hierarchical")
        del os.environ["SGM_WIDE_PTS"]
        it.hierarchical()._dev(0)[0].reload_knobs()
    os.environ["SGM_NO_H2"] = "1"
    it.hierarchical()._dev(0)[0].reload_knobs()
    it.interpolate(rng.normal(size=(100, it.N)) if smp == "normal" else rng.uniform(-1, 1, (100, it.N)),
                   rng.normal(size=(it.num_nodes, 1)), method="hierarchical")
    del os.environ["SGM_NO_H2"]
    print(name, "ok", flush=True)
seq = [SGInterpolant(mis.smolyak_mid_set(k, 3), nodes_1d.cc_nodes, dbl, tp_interpolant=TPPwQuadraticInterpolator, verbose=False)
       for k in range(3)]
mli = MLInterpolant(seq)
ml = [rng.normal(size=(seq[2 - k].num_nodes, 7)) for k in range(3)]
yy = rng.uniform(-1, 1, (P, 3))
a, b = mli.interpolate(yy, ml, method="hierarchical"), mli.interpolate(yy, ml, method="combination")
assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b))
print("multilevel ok", flush=True)
dev = torch.device("cuda")
for d, nt in ((64, 300), (128, 200), (10, 100)):
    yy = torch.randn((700, d), dtype=torch.float64, device=dev)
    lc_brownian_device(torch.linspace(0, 1, nt, dtype=torch.float64, device=dev), yy, 1.0,
                       torch.empty((700, nt), dtype=torch.float64, device=dev))
torch.cuda.synchronize()
print("lc ok", flush=True)
ex1 = SGInterpolant(mis.smolyak_mid_set(1, 5), nodes_1d.unbounded_nodes_nested, l2, verbose=False)
os.environ["SGM_JIT_SYNC"] = "1"
ex1.interpolate(rng.normal(size=(70000, 5)), rng.normal(size=(ex1.num_nodes, 1)), method="combination")
print("small-plan kernels ok", flush=True)
