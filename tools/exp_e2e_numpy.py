"""Experiment: chunk size of the pipelined host path (numpy and pinned-tensor inputs).
Dev tool, not part of the product."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload
import sgmethods_b200.sparse_grid_interpolant as sgi
it, desc, P, NF, sampler = build_workload("c2")
x = np.random.default_rng(0).normal(size=(P, it.N)); f = np.random.default_rng(1).normal(size=(it.num_nodes, 1))
xp = torch.from_numpy(x).pin_memory(); fp = torch.from_numpy(f).pin_memory()
outp = torch.empty((P, 1), dtype=torch.float64).pin_memory()
for chunk in (65536, 131072, 14208 * 8, 14208 * 12, 14208 * 16, 14208 * 24, 10 ** 9):
    sgi._H2D_CHUNK = chunk; it._staging = None
    for _ in range(2): it.interpolate(x, f)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): out = it.interpolate(x, f)
    dt = (time.perf_counter() - t0) / 5
    for _ in range(2): outp.copy_(it.interpolate(xp, fp).reshape(P, 1))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): outp.copy_(it.interpolate(xp, fp).reshape(P, 1), non_blocking=True)
    torch.cuda.synchronize()
    dtp = (time.perf_counter() - t0) / 5
    print(f"chunk {chunk}: numpy->numpy {dt*1e3:.1f} ms, pinned->pinned {dtp*1e3:.1f} ms", flush=True)
