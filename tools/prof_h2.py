"""Minimal driver for ncu: the scalar hierarchical kernel (default shape) on a bench workload."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_workload
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
it, desc, P, NF, sampler = build_workload(name)
if len(sys.argv) > 2:
    P = int(sys.argv[2])
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
hp = it.hierarchical()
plan, _, _ = hp._dev(0)
alpha = hp.surpluses(f)
for _ in range(3):
    plan.eval(x, alpha)
torch.cuda.synchronize()
print("ok")
