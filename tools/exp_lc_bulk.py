"""Levy-Ciesielski kernel: row groups staged by ONE bulk asynchronous copy (cp.async.bulk, 1-D
TMA, mbarrier) against 256 x 8-byte cp.async (SGM_LC_BULK=0).  Run once per flavour (the knob is
read once per process): python tools/exp_lc_bulk.py; SGM_LC_BULK=0 python tools/exp_lc_bulk.py"""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgmethods_b200.parametric_expansions_brownian_motion import lc_brownian_device  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator(device=dev).manual_seed(0)
for P, d, nt in ((100000, 64, 1000), (10000, 64, 10000), (100000, 32, 1000), (200000, 16, 500), (100000, 128, 1000)):
    yy = torch.randn((P, d), dtype=torch.float64, device=dev, generator=gen)
    tt = torch.linspace(0, 1, nt, dtype=torch.float64, device=dev)
    W = torch.empty((P, nt), dtype=torch.float64, device=dev)
    for _ in range(3):
        lc_brownian_device(tt, yy, 1.0, W)
    ts = []
    for _ in range(9):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        lc_brownian_device(tt, yy, 1.0, W)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    nbytes = 8 * P * d + 8 * nt + 8 * P * nt
    digest = hashlib.sha1(W.cpu().numpy().tobytes()).hexdigest()[:12]
    print(f"SGM_LC_BULK={os.environ.get('SGM_LC_BULK', '1')} P={P} d={d} nt={nt}: {ms:.4f} ms "
          f"{nbytes / ms / 1e6:.0f} GB/s sha1 {digest}", flush=True)
