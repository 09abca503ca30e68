import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload
import sgmethods_b200.sparse_grid_interpolant as sgi
it, desc, P, NF, sampler = build_workload("c2")
x = np.random.default_rng(0).normal(size=(P, it.N)); f = np.random.default_rng(1).normal(size=(it.num_nodes, 1))
for label, chunk in (("pipelined numpy path", 131072), ("single copy", 10 ** 9)):
    sgi._H2D_CHUNK = chunk; it._staging = None
    for _ in range(2): it.interpolate(x, f)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): out = it.interpolate(x, f)
    dt = (time.perf_counter() - t0) / 5
    print(f"{label}: {dt*1e3:.1f} ms per call numpy->numpy, {P/dt:.3e} point-evals/s", flush=True)
