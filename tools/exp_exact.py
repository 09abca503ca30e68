"""Exact (method='combination') path timings on config 2 / config 5 shapes: which scalar kernel
the launcher picks (prefix vs split) and what it costs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_workload
from tools.exp_hier_ops import timed
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(1)
for name, P in (("c2", 10 ** 6), ("c5", 10 ** 5), ("c2small", 10 ** 5)):
    it, desc, _, NF, _ = build_workload(name)
    x = torch.randn((P, it.N), generator=gen, device=dev, dtype=torch.float64)
    f = torch.randn((it.num_nodes, 1), generator=gen, device=dev, dtype=torch.float64)
    plan = it.device_plan(0)
    res = {}
    for force in ("", "prefix", "split"):
        if force:
            os.environ["SGM_FORCE_KERNEL"] = force
        else:
            os.environ.pop("SGM_FORCE_KERNEL", None)
        plan.reload_knobs()
        try:
            res[force or "auto"] = timed(lambda: plan.eval(x, f), 2)[0]
        except Exception as exc:  # noqa: BLE001
            res[force or "auto"] = str(exc)[:40]
    os.environ.pop("SGM_FORCE_KERNEL", None)
    print(name, P, res, flush=True)
