#!/usr/bin/env python
"""Headline benchmark: sparse-grid interpolant evaluation throughput on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2]

Metric (BASELINE.json): interpolant point-evals x outputs / s, fp64.  One "step" is one
pass of the hot path (`SGInterpolant.interpolate`) over one batch of synthetic points.
Default workload = BASELINE config 2 ("c2"): d=16 Smolyak set w=4 (4845 combination terms,
69 121 nodes), unbounded nested Gaussian knots, piecewise linear, scalar output, 10^6
points per GPU (weak scaling over ranks: every rank evaluates its own 10^6 points; the
path has no data-path collective).

One JSON line is printed by rank 0.  Keys follow the driver's contract; additionally
  roofline      HBM roofline of the evaluation kernel: algorithmic bytes / kernel time
                against MEASURED_PEAKS.json (fallback 6650 GB/s), plus the secondary
                FP64-issue and gather figures that actually bound this path;
  cpu_baseline  the CPU oracle (numpy/scipy port of the reference's term loop, the same
                scipy RegularGridInterpolator call per term) on a bounded sample, all cores;
  e2e           the same metric through the public API with HOST (pinned) buffers: H2D of
                points and nodal values, kernel, D2H of the result inside the timed region.
`--impl reference` times the CPU oracle port alone (the reference is pure Python/numpy; its
constructor is O(M^2) so its tables come from the verified builder).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# DRAM bytes per launch of the evaluation kernel from the committed ncu capture
# (136.39 MB read + 4.07 MB written for config 2 at 10^6 points; algorithmic: 138.09 MB;
#  ncu --set full -k regex:eval_points_prefix -s 4 -c 1 python bench.py --steps 2 --warmup 3)
NCU_TRAFFIC = {("c2", 10 ** 6): 140461568}

METRIC = "interpolant point-evals x outputs per second (fp64)"
UNIT = "point-evals*outputs/s"


# ------------------------------------------------------------------------------- workloads
def build_workload(name):
    """Returns (interpolant, description, points-per-rank, NF, point sampler kind)."""
    from sgmethods_b200 import multi_index_sets as mis, nodes_1d
    from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
    from sgmethods_b200.tp_inteprolants import TPLagrangeInterpolator
    if name == "c2":
        it = SGInterpolant(mis.smolyak_mid_set(4, 16), nodes_1d.unbounded_nodes_nested,
                           lambda n: 2 ** (n + 1) - 1, verbose=False)
        return it, "config2: d=16 smolyak(w=4) unbounded-nested knots pw-linear NF=1", 10 ** 6, 1, "normal"
    if name == "c2small":
        it = SGInterpolant(mis.smolyak_mid_set(3, 16), nodes_1d.unbounded_nodes_nested,
                           lambda n: 2 ** (n + 1) - 1, verbose=False)
        return it, "config2-small: d=16 smolyak(w=3) pw-linear NF=1", 10 ** 5, 1, "normal"
    if name == "c1":
        aniso = 2 ** (np.linspace(0, 7, 8))
        it = SGInterpolant(mis.aniso_smolyak_mid_set(4, 8, aniso), nodes_1d.cc_nodes,
                           lambda n: np.where(0, 1, 2 ** n + 1),
                           tp_interpolant=TPLagrangeInterpolator, verbose=False)
        return it, "config1: example/0 as shipped (Lagrange, d=8, 7 terms)", 256, 1, "uniform"
    if name == "c3":
        from sgmethods_b200.utils import compute_level_lc
        a = np.array([compute_level_lc(n)[0] + 1 for n in range(64)], dtype=float)
        it = SGInterpolant(mis.aniso_smolyak_mid_set(8, 64, a),
                           lambda n: nodes_1d.optimal_gaussian_nodes(n, p=2),
                           lambda n: 2 ** (n + 1) - 1, verbose=False)
        return it, "config3 (w=8): LC anisotropy d=64 pw-linear NF=10^4", 10 ** 4, 10 ** 4, "normal"
    raise SystemExit(f"unknown workload {name}")


def algorithmic_bytes(it, P, NF):
    """read the N coordinates of each point, write each output once, read the
    nodal values once, walk the packed tables once."""
    tables = 4 * it._maps.size + 8 * it._fam_knots.size + 16 * len(it._coeffs) + \
        16 * int((it._num_nodes > 1).sum()) + 8 * len(it._coeffs)
    return 8 * P * it.N + 8 * P * NF + 8 * it.num_nodes * NF + tables


def secondary_rooflines(P, corners, terms, kernel_ms, clk):
    """Lower bounds of the exact pw-linear scalar kernel on the on-chip resources that
    actually bind it (148 SMs): L1 data pipe 128 B/clk/SM moves one 8-byte value per corner
    (2 wavefronts per warp-wide 64-bit gather at best), FP64 pipe 64 lanes/clk/SM executes
    ~3 operations per corner with prefix reuse (one weight product, value product, add)."""
    mhz = (clk.get("sm_mhz") or 1965.0) * 1e6
    sms = 148
    t_l1 = P * corners * 8 / (128 * sms * mhz)
    t_fp64 = P * corners * 3 / (64 * sms * mhz)
    t = kernel_ms * 1e-3
    return {"l1_data_pipe_floor_ms": t_l1 * 1e3, "fp64_floor_ms": t_fp64 * 1e3,
            "frac_of_l1_floor": t_l1 / t, "frac_of_fp64_floor": t_fp64 / t}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- CPU baseline
_CPU_STATE = {}


def _cpu_worker(args):
    lo, hi = args
    from oracle import sg_oracle
    st = _CPU_STATE
    return sg_oracle.interpolate(st["plan"], st["x"][lo:hi], st["f"], "linear_scipy"
                                 if st["kind"] == "linear" else st["kind"], squeeze=False)


def cpu_reference_rate(it, kind, NF, sample_points, seed=0, sampler="normal", repeats=1):
    """Throughput of the CPU oracle port on `sample_points` points of the workload, split
    over all host cores (fork pool).  Returns (value, cores, seconds, sample description)."""
    import multiprocessing as mp
    from oracle import sg_oracle
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(sample_points, it.N)) if sampler == "normal" else \
        rng.uniform(-1, 1, (sample_points, it.N))
    f = np.random.default_rng(1).normal(size=(it.num_nodes, NF))
    plan = sg_oracle.plan_from_tables(it.N, it.combination_coeffs, it.active_tp_dims,
                                      it.active_tp_nodes_list, it.map_tp_to_SG)
    _CPU_STATE.update(plan=plan, x=x, f=f, kind=kind)
    bounds = np.linspace(0, sample_points, cores + 1).astype(int)
    chunks = [(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    best = None
    ctx = mp.get_context("fork")
    with ctx.Pool(len(chunks)) as pool:
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, chunks)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return sample_points * NF / best, len(chunks), best, \
        f"{sample_points} points of the workload split over {len(chunks)} processes"


def cpu_single_core_rate(it, kind, NF, sample_points, sampler="normal"):
    from oracle import sg_oracle
    rng = np.random.default_rng(0)
    x = rng.normal(size=(sample_points, it.N)) if sampler == "normal" else \
        rng.uniform(-1, 1, (sample_points, it.N))
    f = np.random.default_rng(1).normal(size=(it.num_nodes, NF))
    plan = sg_oracle.plan_from_tables(it.N, it.combination_coeffs, it.active_tp_dims,
                                      it.active_tp_nodes_list, it.map_tp_to_SG)
    t0 = time.perf_counter()
    sg_oracle.interpolate(plan, x, f, "linear_scipy" if kind == "linear" else kind)
    dt = time.perf_counter() - t0
    return sample_points * NF / dt, dt


KIND_NAME = {0: "linear", 1: "quadratic", 2: "cubic", 3: "lagrange"}


# ------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--points", type=int, default=0, help="points per rank (0 = workload default)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="points of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hierarchical", action="store_true",
                    help="skip the extra (non-headline) measurement of the opt-in hierarchical form")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries the ONE JSON line only: keep a private handle on it and point file
    # descriptor 1 at stderr, so that whatever libraries write to stdout (NCCL prints its
    # version line there) cannot end up next to the result
    sys.stdout.flush()
    try:
        json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    except OSError:                       # no usable descriptors: print where print() goes
        json_out = sys.stdout

    if args.impl == "reference":
        if rank != 0:
            return 0
        it, desc, P, NF, sampler = build_workload(args.workload)
        kind = KIND_NAME[it._kind]
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or cores * 1024
        vals = []
        for _ in range(max(args.warmup, 0) and 1):
            cpu_reference_rate(it, kind, NF, max(cores * 64, 64), sampler=sampler)
        t_all = time.perf_counter()
        for s in range(args.steps):
            v, used, dt, what = cpu_reference_rate(it, kind, NF, sample, seed=s, sampler=sampler)
            vals.append((v, dt))
        total = time.perf_counter() - t_all
        value = sample * NF * len(vals) / sum(dt for _, dt in vals)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * sum(dt for _, dt in vals) / len(vals),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "points_per_step": sample, "NF": NF,
                           "terms": len(it._coeffs), "nodes": it.num_nodes},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                                 "sample": what + "; oracle/ numpy+scipy port of the reference's "
                                 "term loop, tables from the verified builder"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "wall_s": total}
        print(json.dumps(line), file=json_out, flush=True)
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference "
                         "for the CPU baseline")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    it, desc, P_default, NF, sampler = build_workload(args.workload)
    P = args.points or P_default
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    if sampler == "normal":
        x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
    else:
        x = torch.rand((P, it.N), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    f = torch.from_numpy(np.random.default_rng(1).normal(size=(it.num_nodes, NF))).to(dev)
    plan = it.device_plan(local_rank)
    out = torch.empty((P, NF), dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # > 126 MB L2

    def step():
        plan.eval(x, f, out)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- device-resident timing: K steps, each bracketed by CUDA events, L2 flushed between
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as clocks:
        t_wall = time.perf_counter()
        for s in range(args.steps):
            flush.fill_(float(s))
            starts[s].record()
            step()
            ends[s].record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t_wall
    if world > 1:
        dist.barrier()
    step_ms = np.array([a.elapsed_time(b) for a, b in zip(starts, ends)])
    total_ms = torch.tensor([float(step_ms.sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * P * NF * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API with pinned host buffers
    x_host = x.cpu().pin_memory()
    f_host = f.cpu().pin_memory()
    out_host = torch.empty((P, NF), dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        res = it.interpolate(x_host, f_host)          # H2D + kernel, returns a CUDA tensor
        out_host.copy_(res.reshape(P, NF), non_blocking=True)   # D2H of the result
    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * P * NF * e2e_steps / (float(e2e_ms.item()) * 1e-3)
    # the drop-in call exactly as a reference user makes it: numpy in, numpy out (pageable
    # host memory, staged through pinned buffers by the library); wall clock, rank-local
    x_np, f_np = x_host.numpy(), f_host.numpy()
    it.interpolate(x_np, f_np)
    t_np = time.perf_counter()
    for _ in range(3):
        res_np = it.interpolate(x_np, f_np)
    numpy_ms = (time.perf_counter() - t_np) / 3 * 1e3
    assert np.array_equal(res_np.reshape(P, NF), out_host.numpy())
    assert torch.equal(out_host.to(dev), out), "e2e result differs from the device-resident run"

    gather_ms = None
    if world > 1:
        # the optional final exchange of the point-sharded result (north star item 4): NCCL
        # all-gather of the (P, NF) slabs, timed separately from the collective-free step
        from sgmethods_b200.distributed import ShardedInterpolant
        sh = ShardedInterpolant(it, mode="points")
        sh.gather_blocks(out, world * P, "all")
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(5):
            full = sh.gather_blocks(out, world * P, "all")
        g1.record()
        torch.cuda.synchronize()
        gm = torch.tensor([g0.elapsed_time(g1) / 5], dtype=torch.float64, device=dev)
        dist.all_reduce(gm, op=dist.ReduceOp.MAX)
        gather_ms = float(gm.item())
        assert full.shape == (world * P, NF)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        kernel_ms = float(np.mean(step_ms))
        bytes_alg = algorithmic_bytes(it, P, NF)
        achieved = bytes_alg / (kernel_ms * 1e-3) / 1e9
        corners = plan.corners_per_point
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc, "points_per_gpu": P, "NF": NF, "N": it.N,
                       "terms": len(it._coeffs), "nodes": it.num_nodes,
                       "corners_per_point": corners, "l2": "flushed between timed steps "
                       "(256 MB fill)", "sharding": "points, no collective"},
            "gpu_launches": 5 * args.steps,   # pregather, sort hist/scan/scatter, eval
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(x_host.numel() * 8 + f_host.numel() * 8),
                    "d2h_bytes_per_step": int(P * NF * 8), "steps": e2e_steps,
                    "numpy_in_numpy_out": {"value": P * NF / (numpy_ms * 1e-3), "ms_per_call": numpy_ms,
                                           "note": "SGInterpolant.interpolate(ndarray, ndarray) -> "
                                                   "ndarray on rank 0, pageable memory, wall clock"}},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": NCU_TRAFFIC.get((args.workload, P)),
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the "
                         "kernel, ncu --set full capture of this command "
                         "(profiles/r1_v5_prefix_ppl2_ncu_full.txt)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_alg,
                         "kernel": "sgm::eval_points_prefix_kernel<8 warps, 64-term chunks, balanced "
                                   "runs, 2 points per lane> (+ pregather and the point locality "
                                   "sort, <1% of the step; profiles/r1_v5_launches.csv)",
                         "kernel_ms": kernel_ms,
                         "note": "the exact combination form is bound by the SM's L1 data pipe "
                                 "(128 B/clk/SM; 8-byte corner gathers + bracket tables) and "
                                 "FP64 issue, not HBM (DESIGN.md section 5); ncu of this kernel: "
                                 "L1 data pipe 68 %, issue 46 %, FP64 pipe 38 % of peak "
                                 "(profiles/r1_v5_prefix_ppl2_ncu_full.txt)",
                         "corner_gathers_per_s": P * corners / (kernel_ms * 1e-3),
                         "secondary": secondary_rooflines(P, corners, len(it._coeffs),
                                                          kernel_ms, clocks.summary())},
            "clocks": clocks.summary(),
            "wall_s_timed_region": wall,
        }
        if world == 1 and not args.no_hierarchical and it._kind == 0:
            # extra, NOT the headline: the opt-in hierarchical-surplus form of the same
            # interpolant (same function, not bit-identical)
            try:
                hp = it.hierarchical()
                hv = hp.interpolate(x, f)
                dev_rel = float((hv.reshape(out.shape) - out).abs().max() / out.abs().max())
                hs = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
                torch.cuda.synchronize()
                for i in range(5):
                    flush.fill_(0.0)
                    hs[0].record(); hp.interpolate(x, f); hs[1].record()
                    torch.cuda.synchronize()
                    hms = hs[0].elapsed_time(hs[1]) if i == 0 else min(hms, hs[0].elapsed_time(hs[1]))
                line["hierarchical_mode"] = {
                    "value": P * NF / (hms * 1e-3), "unit": UNIT, "ms_per_step": hms,
                    "max_rel_deviation_from_headline_result": dev_rel,
                    "note": "opt-in method='hierarchical' (surplus form, ~10x fewer gathers); "
                            "includes the GPU hierarchisation of the nodal values; best of 5"}
            except NotImplementedError as exc:
                line["hierarchical_mode"] = {"unavailable": str(exc)}
        if world == 1:
            # the HBM-bound kernel of the same path family: batched
            # Levy-Ciesielski map of config 3's parameter vectors, bytes = 8 P d + 8 nt + 8 P nt
            from sgmethods_b200.parametric_expansions_brownian_motion import lc_brownian_device
            Pl, dl, ntl = 10 ** 5, 64, 10 ** 3
            yy = torch.randn((Pl, dl), dtype=torch.float64, device=dev, generator=gen)
            tt = torch.linspace(0, 1, ntl, dtype=torch.float64, device=dev)
            Wl = torch.empty((Pl, ntl), dtype=torch.float64, device=dev)
            for _ in range(3):
                lc_brownian_device(tt, yy, 1.0, Wl)
            lts = []
            for i in range(7):
                flush.fill_(float(i))
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); lc_brownian_device(tt, yy, 1.0, Wl); b.record()
                torch.cuda.synchronize()
                lts.append(a.elapsed_time(b))
            lms = float(np.median(lts))
            lbytes = 8 * Pl * dl + 8 * ntl + 8 * Pl * ntl
            line["hbm_bound_kernel"] = {
                "kernel": "sgm::lc_brownian_tiles_kernel (Levy-Ciesielski expansion, d=64, 10^5 "
                          "parameter vectors x 10^3 times)",
                "ms": lms, "algorithmic_bytes": lbytes, "achieved": lbytes / lms / 1e6, "peak": peak,
                "unit": "GB/s", "frac": lbytes / lms / 1e6 / peak,
                "note": "not the headline workload: shows the HBM fraction of the path's "
                        "HBM-bound kernel next to the gather-bound evaluation kernel"}
        if gather_ms is not None:
            line["final_all_gather"] = {"ms": gather_ms, "bytes_per_rank": int(P * NF * 8),
                                        "backend": "nccl all_gather_into_tensor",
                                        "in_timed_step": False}
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N=1 only
            cores = os.cpu_count() or 1
            sample = args.cpu_sample or cores * 1024
            v, used, dt, what = cpu_reference_rate(it, KIND_NAME[it._kind], NF, sample,
                                                   sampler=sampler)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": used, "kind": "port",
                                    "sample": what + f" ({dt:.1f} s)"}
            # the reference as shipped is single-threaded: the same port on ONE core
            v1, dt1 = cpu_single_core_rate(it, KIND_NAME[it._kind], NF, 1024, sampler)
            line["cpu_baseline"]["single_core"] = {"value": v1, "cores": 1,
                                                   "sample": f"1024 points, 1 process ({dt1:.1f} s)"}
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
