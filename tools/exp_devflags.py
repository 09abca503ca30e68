"""A/B of kernel variants selected by SGM_DEV_FLAGS in ONE process (box-to-box variance is ~5 %):
bit 0 = strided entry fill (old), bit 1 = read all 7 member slots of a group (old)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator(device=dev).manual_seed(1)


def timed(fn, reps=7):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


for name, P in (("c2", 10 ** 6), ("c5", 10 ** 6)):
    it, desc, _, NF, _ = build_workload(name)
    x = torch.randn((P, it.N), generator=gen, device=dev, dtype=torch.float64)
    f = torch.randn((it.num_nodes, 1), generator=gen, device=dev, dtype=torch.float64)
    hp = it.hierarchical()
    hplan, eplan = hp._dev(0)[0], it.device_plan(0)
    alpha = hp.surpluses(f)
    for rnd in range(2):
        for flags in (0, 8, 0, 8):
            os.environ["SGM_DEV_FLAGS"] = str(flags)
            hplan.reload_knobs(); eplan.reload_knobs()
            th = timed(lambda: hplan.eval(x, alpha))
            te = timed(lambda: eplan.eval(x[:200000] if name == "c5" else x, f), 3)
            print(f"{name} flags={flags}: hierarchical {th:.3f} ms, exact {te:.3f} ms", flush=True)
