import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgmethods_b200.parametric_expansions_brownian_motion import lc_brownian_device
dev = torch.device("cuda")
P, d, nt = 100000, 64, 1000
gen = torch.Generator(device=dev).manual_seed(0)
yy = torch.randn((P, d), dtype=torch.float64, device=dev, generator=gen)
tt = torch.linspace(0, 1, nt, dtype=torch.float64, device=dev)
W = torch.empty((P, nt), dtype=torch.float64, device=dev)
for _ in range(4):
    lc_brownian_device(tt, yy, 1.0, W)
torch.cuda.synchronize()
print("ok")
