"""Dev tool: run the scalar kernel on identical points (all gathers are broadcasts) so that
ncu shows what bounds the instruction stream when L1 divergence is taken away."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_workload
it, desc, P, NF, sampler = build_workload("c2")
P = 400000
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((1, it.N), dtype=torch.float64, device=dev, generator=gen).expand(P, -1).contiguous()
f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
plan = it.device_plan(0)
out = torch.empty((P, 1), dtype=torch.float64, device=dev)
for _ in range(4):
    plan.eval(x, f, out)
torch.cuda.synchronize()
print("ok")
