"""Hierarchical-surplus form of the pw-quadratic and Lagrange operators against the exact
kernels: config 2 with Gaussian-Leja knots (Lagrange, 10^6 points) and config 4's finest level
(pw-quadratic d=8, NF=10^3, 10^5 points)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_workload  # noqa: E402


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(n):
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        out = fn()
        t1.record()
        torch.cuda.synchronize()
        best = min(best, t0.elapsed_time(t1))
    return best, out


def main():
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(1)
    it, desc, P, NF, _ = build_workload("c2leja")
    x = torch.randn((P, it.N), generator=gen, device=dev, dtype=torch.float64)
    f = torch.randn((it.num_nodes, 1), generator=gen, device=dev, dtype=torch.float64)
    t0 = time.perf_counter()
    hp = it.hierarchical()
    print(f"{desc}: hierarchical plan {time.perf_counter() - t0:.2f} s, {hp.n_terms} basis tuples")
    tc, oc = timed(lambda: it.interpolate(x, f, method="combination"), 2)
    th, oh = timed(lambda: it.interpolate(x, f, method="hierarchical"))
    ta, oa = timed(lambda: it.interpolate(x, f))
    dev_ = float((oh - oc).abs().max() / oc.abs().max())
    print(f"  combination {tc:.2f} ms, hierarchical {th:.2f} ms, auto {ta:.2f} ms ({it.auto_validation}); rel dev {dev_:.2e}")
    ml, desc, P, NF, _ = build_workload("c4")
    it = ml.interp_seq[-1]
    P = 10 ** 5
    x = torch.rand((P, it.N), generator=gen, device=dev, dtype=torch.float64) * 2 - 1
    f = torch.randn((it.num_nodes, NF), generator=gen, device=dev, dtype=torch.float64)
    hp = it.hierarchical()
    print(f"config 4 finest level: {hp.n_terms} basis tuples, {it.num_nodes} nodes")
    tc, oc = timed(lambda: it.interpolate(x, f, method="combination"), 2)
    th, oh = timed(lambda: it.interpolate(x, f, method="hierarchical"))
    ta, oa = timed(lambda: it.interpolate(x, f))
    dev_ = float((oh - oc).abs().max() / oc.abs().max())
    print(f"  P={P} NF={NF}: combination {tc:.2f} ms, hierarchical {th:.2f} ms, auto {ta:.2f} ms ({it.auto_validation}); rel dev {dev_:.2e}")


if __name__ == "__main__":
    main()
