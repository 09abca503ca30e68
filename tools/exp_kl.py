"""Timing of the Karhunen-Loeve expansion kernel (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sgmethods_b200.parametric_expansions_brownian_motion import param_KL_Brownian_motion
dev = torch.device("cuda")
for P, d, nt in ((100000, 64, 1000), (10000, 64, 10000)):
    gen = torch.Generator(device=dev).manual_seed(0)
    yy = torch.randn((P, d), dtype=torch.float64, device=dev, generator=gen)
    tt = torch.linspace(0, 1, nt, dtype=torch.float64, device=dev)
    for _ in range(2): param_KL_Brownian_motion(tt, yy)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); W = param_KL_Brownian_motion(tt, yy); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"KL d={d} P={P} nt={nt}: {np.median(ts):.3f} ms, {P * nt / np.median(ts) / 1e6:.2f} G outputs/s, "
          f"{P * nt * d / np.median(ts) / 1e9:.2f} T madds/s", flush=True)
