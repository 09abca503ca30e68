"""Config 1 (example/0 as shipped: Lagrange, 7 terms, 256 points): per-kernel times (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload
it, desc, P, NF, sampler = build_workload("c1")
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.rand((P, it.N), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
plan = it.device_plan(0)
out = torch.empty((P, 1), dtype=torch.float64, device=dev)
for _ in range(5): plan.eval(x, f, out)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); plan.eval(x, f, out); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(desc, "P", P, "median ms", float(np.median(ts)), "terms", len(it._coeffs), "E", "nodes/term", [int(np.prod(s)) for s in it._num_nodes][:8])
