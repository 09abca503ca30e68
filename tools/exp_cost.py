"""Scan the cost-model knobs of the prefix kernel's run partition (dev tool)."""
import sys, os, subprocess
code = r'''
import sys, os
sys.path.insert(0, ".")
import numpy as np, torch
from bench import build_workload
it, desc, P, NF, sampler = build_workload("c2")
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
plan = it.device_plan(0)
out = torch.empty((P, 1), dtype=torch.float64, device=dev)
for _ in range(3): plan.eval(x, f, out)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); plan.eval(x, f, out); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"{min(ts):.2f}")
'''
for knobs in sys.argv[1:]:
    env = dict(os.environ)
    for kv in knobs.split(","):
        if kv:
            k, v = kv.split("=")
            env["SGM_COST_" + k] = v
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(knobs or "(default)", "->", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], "ms", flush=True)
