import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload
it, desc, P0, NF0, sampler = build_workload("c2")
dev = torch.device("cuda"); plan = it.device_plan(0)
gen = torch.Generator(device=dev).manual_seed(0)
P = 100000
x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
for NF in (48, 49, 64, 100, 127, 128, 192):
    f = torch.randn((it.num_nodes, NF), dtype=torch.float64, device=dev, generator=gen)
    out = torch.empty((P, NF), dtype=torch.float64, device=dev)
    for _ in range(2): plan.eval(x, f, out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): plan.eval(x, f, out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"NF={NF}: {ms:.2f} ms, {P*NF/ms*1e3:.3e} evals/s", flush=True)
