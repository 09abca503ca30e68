"""numpy in -> numpy out call on config 2 (10^6 points): staging copy threads."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload
from sgmethods_b200 import sparse_grid_interpolant as sgi
it, desc, P, NF, _ = build_workload("c2")
rng = np.random.default_rng(0)
x = rng.normal(size=(P, it.N)); f = rng.normal(size=(it.num_nodes, 1))
for rnd in range(2):
    for th in (1, 2, 4, 8):
        sgi._STAGING_THREADS = th
        it.interpolate(x, f); it.interpolate(x, f)
        ts = []
        for _ in range(7):
            t0 = time.perf_counter(); out = it.interpolate(x, f); ts.append(time.perf_counter() - t0)
        print(f"staging threads {th}: {1e3 * np.median(ts):.2f} ms per call", flush=True)
