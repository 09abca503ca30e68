import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_workload
it, desc, P, NF, sampler = build_workload("c2")
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
hp = it.hierarchical()
for _ in range(4): hp.interpolate(x, f)
torch.cuda.synchronize(); print("ok")
