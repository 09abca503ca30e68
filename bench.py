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

# DRAM bytes per launch of the evaluation kernel from the committed ncu captures of this command
# (ncu --set full -k regex:eval_hier_groups -s 5 -c 1 python bench.py --steps 2 --warmup 3
#  --no-extras --no-cpu-baseline -> profiles/r2_h2_groups_c2_ncu_full.txt: 132.63 MB read + 5.47 MB
#  written; exact kernel, round 1: 136.39 + 4.07 MB; algorithmic: 138.09 MB)
NCU_TRAFFIC = {("c2", 10 ** 6, "hierarchical"): 138099200, ("c2", 10 ** 6, "combination"): 140461568}

METRIC = "interpolant point-evals x outputs per second (fp64)"
UNIT = "point-evals*outputs/s"


# ------------------------------------------------------------------------------- workloads
WORKLOADS = ("c1", "c2", "c2small", "c2leja", "c3", "c3full", "c4", "c5")


def build_workload(name):
    """Returns (interpolant, description, points-per-rank, NF, point sampler kind).  For "c4"
    the first entry is an MLInterpolant (4 levels) -- see run_multilevel()."""
    from sgmethods_b200 import multi_index_sets as mis, nodes_1d
    from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
    from sgmethods_b200.tp_inteprolants import TPLagrangeInterpolator, TPPwQuadraticInterpolator
    l2 = lambda n: 2 ** (n + 1) - 1  # noqa: E731
    if name == "c2":
        it = SGInterpolant(mis.smolyak_mid_set(4, 16), nodes_1d.unbounded_nodes_nested, l2, verbose=False)
        return it, "config2: d=16 smolyak(w=4) unbounded-nested knots pw-linear NF=1", 10 ** 6, 1, "normal"
    if name == "c2small":
        it = SGInterpolant(mis.smolyak_mid_set(3, 16), nodes_1d.unbounded_nodes_nested, l2, verbose=False)
        return it, "config2-small: d=16 smolyak(w=3) pw-linear NF=1", 10 ** 5, 1, "normal"
    if name == "c2leja":
        # BASELINE config 2's wording: Gaussian weighted-Leja knots (the sequence the reference
        # ships as .mat data; its first 12 points are a test fixture), Lagrange operator
        from sgmethods_b200.leja import leja_knots
        with np.load(os.path.join(ROOT, "tests", "golden", "lag_leja_gauss.npz")) as z:
            knots = leja_knots(z["leja_X"])
        it = SGInterpolant(mis.smolyak_mid_set(4, 16), knots, lambda nu: nu + 1,
                           tp_interpolant=TPLagrangeInterpolator, verbose=False)
        return it, "config2-leja: d=16 smolyak(w=4) Gaussian-Leja knots Lagrange NF=1", 10 ** 6, 1, "normal"
    if name == "c1":
        aniso = 2 ** (np.linspace(0, 7, 8))
        it = SGInterpolant(mis.aniso_smolyak_mid_set(4, 8, aniso), nodes_1d.cc_nodes,
                           lambda n: np.where(0, 1, 2 ** n + 1),
                           tp_interpolant=TPLagrangeInterpolator, verbose=False)
        return it, "config1: example/0 as shipped (Lagrange, d=8, 7 terms)", 256, 1, "uniform"
    if name in ("c3", "c3full"):
        from sgmethods_b200.utils import compute_level_lc
        a = np.array([compute_level_lc(n)[0] + 1 for n in range(64)], dtype=float)
        w, P = (8, 10 ** 4) if name == "c3" else (10, 10 ** 5)
        it = SGInterpolant(mis.aniso_smolyak_mid_set(w, 64, a),
                           lambda n: nodes_1d.optimal_gaussian_nodes(n, p=2), l2, verbose=False)
        return it, f"config3 (w={w}): LC anisotropy d=64 pw-linear NF=10^4, {P} points", P, 10 ** 4, "normal"
    if name == "c4":
        from sgmethods_b200.multilevel_interpolant import MLInterpolant
        dbl = lambda n: np.where(n == 0, 1, 2 ** n + 1)  # noqa: E731
        seq = [SGInterpolant(mis.smolyak_mid_set(k, 8), nodes_1d.cc_nodes, dbl,
                             tp_interpolant=TPPwQuadraticInterpolator, verbose=False) for k in range(4)]
        return MLInterpolant(seq), "config4: multilevel, 4 levels, pw-quadratic d=8 NF=10^3", 10 ** 6, 10 ** 3, "uniform"
    if name == "c5":
        S = mis.aniso_smolyak_mid_set(3, 128, np.array([1.0] * 80 + [3.0] * 48))
        it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, l2, verbose=False)
        return it, "config5: d=128, 91 929 terms, pw-linear NF=1", 10 ** 6, 1, "normal"
    raise SystemExit(f"unknown workload {name} (choices: {', '.join(WORKLOADS)})")


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


def bind_to_gpu_numa_node(torch, local_rank):
    """One process per GPU: run this rank's host threads (and therefore first-touch its pinned
    staging memory) on the CPU cores of the NUMA node its GPU hangs off, so that the N
    simultaneous host-to-device copies of the e2e step do not cross the socket interconnect.
    Returns a short description (or None when the topology files are not visible)."""
    try:
        if os.environ.get("SGM_BENCH_NO_NUMA"):
            return None
        pr = torch.cuda.get_device_properties(local_rank)
        pci = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{pci}/local_cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        use = cpus & os.sched_getaffinity(0)
        if not use:
            return None
        os.sched_setaffinity(0, use)
        return f"gpu {pci}: {len(use)} local cores ({spec})"
    except Exception:                                   # no sysfs / no permission: leave as is
        return None


def measured_fp64():
    """FP64 pipe rates measured on the box by tools/mb_fp64 (built by __graft_entry__.build();
    run live when the binary is there), else the committed record of the same binary
    (profiles/r2_mb_fp64.json), else None."""
    exe = os.path.join(ROOT, "tools", "mb_fp64")
    try:
        if os.path.exists(exe):
            out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
            rec = json.loads(out.strip().splitlines()[-1])
            rec["source"] = "tools/mb_fp64 run inside this bench"
            return rec
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "r2_mb_fp64.json")) as fh:
            rec = json.load(fh)
        rec["source"] = "profiles/r2_mb_fp64.json (tools/mb_fp64 on a B200 of this pool)"
        return rec
    except Exception:
        return None


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
                 "--format=csv,noheader,nounits", "-lms", "25"],
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
KIND_NAME = {0: "linear", 1: "quadratic", 2: "cubic", 3: "lagrange"}


def reference_tables(name):
    """Everything the CPU arm needs for a workload: the oracle's plan (per-term tables), N,
    number of nodes, operator kind, NF, sampler, description.  From bench_data/<name>.npz
    (tools/make_ref_tables.py) when it exists -- then neither this function nor the arm touches
    the product library --, else built with the product's table builder."""
    from oracle import sg_oracle
    path = os.path.join(ROOT, "bench_data", name + ".npz")
    if os.path.exists(path):
        with np.load(path) as z:
            g = {k: z[k] for k in z.files}
        T = g["coeffs"].size
        knots_of = {int(m): g["fam_knots"][g["fam_off"][i]:g["fam_off"][i + 1]]
                    for i, m in enumerate(g["fam_m"])}
        dims, nodes, maps = [], [], []
        for t in range(T):
            lo, hi = g["dims_off"][t], g["dims_off"][t + 1]
            shape = tuple(int(v) for v in g["shapes"][lo:hi])
            dims.append(g["dims"][lo:hi].astype(np.int64))
            nodes.append(tuple(knots_of[m] for m in shape) if shape else (g["knots_one"],))
            maps.append(g["maps"][g["map_off"][t]:g["map_off"][t + 1]].astype(np.int64)
                        .reshape(shape or (1,)))
        plan = sg_oracle.plan_from_tables(int(g["N"]), list(g["coeffs"]), dims, nodes, maps)
        return dict(plan=plan, N=int(g["N"]), num_nodes=int(g["num_nodes"]), kind=KIND_NAME[int(g["kind"])],
                    NF=int(g["NF"]), sampler=str(g["sampler"]), desc=str(g["desc"]), terms=T,
                    tables="bench_data/%s.npz (tools/make_ref_tables.py)" % name)
    it, desc, P, NF, sampler = build_workload(name)
    plan = sg_oracle.plan_from_tables(it.N, it.combination_coeffs, it.active_tp_dims,
                                      it.active_tp_nodes_list, it.map_tp_to_SG)
    return dict(plan=plan, N=it.N, num_nodes=it.num_nodes, kind=KIND_NAME[it._kind], NF=NF,
                sampler=sampler, desc=desc, terms=len(it._coeffs),
                tables="built by the product's table builder (no cache file for this workload)")


def _cpu_worker(args):
    lo, hi = args
    from oracle import sg_oracle
    st = _CPU_STATE
    return sg_oracle.interpolate(st["plan"], st["x"][lo:hi], st["f"], "linear_scipy"
                                 if st["kind"] == "linear" else st["kind"], squeeze=False)


def cpu_reference_rate(ref, sample_points, seed=0, cores=None):
    """Throughput of the CPU oracle port on `sample_points` points of the workload, split
    over `cores` processes (default: all host cores; fork pool).  Returns (value, cores,
    seconds, sample description)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(sample_points, ref["N"])) if ref["sampler"] == "normal" else \
        rng.uniform(-1, 1, (sample_points, ref["N"]))
    f = np.random.default_rng(1).normal(size=(ref["num_nodes"], ref["NF"]))
    _CPU_STATE.update(plan=ref["plan"], x=x, f=f, kind=ref["kind"])
    if cores == 1:
        t0 = time.perf_counter()
        _cpu_worker((0, sample_points))
        dt = time.perf_counter() - t0
        return sample_points * ref["NF"] / dt, 1, dt, f"{sample_points} points, 1 process"
    bounds = np.linspace(0, sample_points, cores + 1).astype(int)
    chunks = [(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    with mp.get_context("fork").Pool(len(chunks)) as pool:
        t0 = time.perf_counter()
        pool.map(_cpu_worker, chunks)
        dt = time.perf_counter() - t0
    return sample_points * ref["NF"] / dt, len(chunks), dt, \
        f"{sample_points} points of the workload split over {len(chunks)} processes"


# ------------------------------------------------------------------------------- GPU helpers
def timed_steps(torch, fn, steps, flush, dist=None, dev=None):
    """`steps` calls of fn, each bracketed by CUDA events on the current stream with the L2
    flushed (256 MB fill) in between.  Returns the per-step milliseconds."""
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    torch.cuda.synchronize()
    for s in range(steps):
        flush.fill_(float(s))
        starts[s].record()
        fn()
        ends[s].record()
    torch.cuda.synchronize()
    return np.array([a.elapsed_time(b) for a, b in zip(starts, ends)])


def max_over_ranks(torch, dist, dev, value, world):
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sample_points(torch, P, N, sampler, dev, gen):
    if sampler == "normal":
        return torch.randn((P, N), dtype=torch.float64, device=dev, generator=gen)
    return torch.rand((P, N), dtype=torch.float64, device=dev, generator=gen) * 2 - 1


def main_kernel_ms(plan, P):
    """Mean duration of the sgm_eval launches of `plan` on P points recorded by its timing hook
    (pre-gather + locality sort + evaluation kernel; the evaluation kernel is > 95 % of it,
    profiles/r2_launches.csv)."""
    ts = [a.elapsed_time(b) for (p, nf, a, b) in plan.timing if p == P]
    return float(np.mean(ts)) if ts else None


def run_extra(torch, name, dev, flush, peak, steps=2):
    """One of the other BASELINE workloads at its named size on this GPU: throughput through
    the public call on device-resident inputs + HBM roofline fraction of its algorithmic bytes."""
    t0 = time.perf_counter()
    it, desc, P, NF, sampler = build_workload(name)
    gen = torch.Generator(device=dev).manual_seed(7)
    if name == "c4":                                   # multilevel: 4 levels, one fused launch
        mli, seq = it, it.interp_seq
        yy = sample_points(torch, P, 8, sampler, dev, gen)
        ml = [torch.randn((seq[3 - k].num_nodes, NF), dtype=torch.float64, device=dev, generator=gen)
              for k in range(4)]
        fn = lambda: mli.interpolate(yy, ml)            # noqa: E731
        bytes_alg = 8 * P * 8 + 8 * P * NF + sum(8 * s.num_nodes * NF for s in seq) + \
            sum(4 * s._maps.size for s in seq)
        extra = {"levels": 4, "launches_per_step": 1, "terms": sum(len(s._coeffs) for s in seq[:1]) +
                 sum(len(seq[k]._coeffs) + len(seq[k - 1]._coeffs) for k in range(1, 4))}
    else:
        x = sample_points(torch, P, it.N, sampler, dev, gen)
        f = torch.randn((it.num_nodes, NF), dtype=torch.float64, device=dev, generator=gen)
        fn = lambda: it.interpolate(x, f)               # noqa: E731
        bytes_alg = algorithmic_bytes(it, P, NF)
        extra = {"terms": len(it._coeffs), "nodes": it.num_nodes,
                 "corners_per_point": it.device_plan(dev.index).corners_per_point}
    build_s = time.perf_counter() - t0
    fn()
    ms = float(np.min(timed_steps(torch, fn, steps, flush)))
    if name != "c4":
        extra["method"] = getattr(it, "_auto_choice", None) or "combination"
        extra["auto_validation"] = getattr(it, "auto_validation", None)
        if extra["method"] == "hierarchical":           # the bit-exact path next to it
            fx = lambda: it.interpolate(x, f, method="combination")   # noqa: E731
            fx()
            extra["exact_ms"] = float(np.min(timed_steps(torch, fx, steps, flush)))
    else:
        extra["auto_validation"] = mli.auto_validation
        extra["method"] = "hierarchical" if (mli.auto_validation or {}).get("enabled") else "combination"
        if extra["method"] == "hierarchical":
            extra["basis_tuples"] = int(mli._hierarchical()[0][-1].n_terms)
            fx = lambda: mli.interpolate(yy, ml, method="combination")   # noqa: E731
            fx()
            extra["exact_ms"] = float(np.min(timed_steps(torch, fx, min(steps, 2), flush)))
    out = {"workload": desc, "points": P, "NF": NF, "ms_per_step": ms, "value": P * NF / ms * 1e3,
           "unit": UNIT, "algorithmic_bytes": bytes_alg, "hbm_gbs": bytes_alg / ms / 1e6,
           "hbm_frac": bytes_alg / ms / 1e6 / peak, "host_setup_s": build_s, "timing": f"best of {steps}"}
    out.update(extra)
    return out


def run_hbm_kernels(torch, dev, flush, peak, gen):
    """The two HBM-bound kernels of the path family next to the gather-bound headline."""
    from sgmethods_b200 import multi_index_sets as mis, nodes_1d
    from sgmethods_b200.parametric_expansions_brownian_motion import lc_brownian_device
    from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
    res = {}
    # (1) batched Levy-Ciesielski map of config 3's parameter vectors
    Pl, dl, ntl = 10 ** 5, 64, 10 ** 3
    yy = torch.randn((Pl, dl), dtype=torch.float64, device=dev, generator=gen)
    tt = torch.linspace(0, 1, ntl, dtype=torch.float64, device=dev)
    Wl = torch.empty((Pl, ntl), dtype=torch.float64, device=dev)
    fn = lambda: lc_brownian_device(tt, yy, 1.0, Wl)    # noqa: E731
    for _ in range(3):
        fn()
    lms = float(np.median(timed_steps(torch, fn, 7, flush)))
    lbytes = 8 * Pl * dl + 8 * ntl + 8 * Pl * ntl
    res["hbm_bound_kernel"] = {
        "kernel": "sgm::lc_brownian_tiles_kernel (Levy-Ciesielski expansion, d=64, 10^5 "
                  "parameter vectors x 10^3 times)",
        "ms": lms, "algorithmic_bytes": lbytes, "achieved": lbytes / lms / 1e6, "peak": peak,
        "unit": "GB/s", "frac": lbytes / lms / 1e6 / peak}
    # (2) the interpolant itself in its HBM-bound regime: example/1 as shipped (7 terms, 22
    # corners per point), 10^7 points -- bit-identical to the reference
    S = mis.aniso_smolyak_mid_set(5, 10, 2 ** (np.linspace(0, 9, 10)))
    it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False)
    P = 10 ** 7
    x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
    f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
    plan = it.device_plan(dev.index)
    out = torch.empty((P, 1), dtype=torch.float64, device=dev)
    fn = lambda: plan.eval(x, f, out)                   # noqa: E731
    fn()                                                # crosses the plan's JIT threshold (4e6 points)
    generic_ms = None
    if plan.jit_state()[0] in (0, 2):                   # still the generic kernel A0: time it
        os.environ["SGM_NO_JIT"] = "1"
        plan.reload_knobs()
        fn()
        generic_ms = float(np.median(timed_steps(torch, fn, 5, flush)))
        del os.environ["SGM_NO_JIT"]
        plan.reload_knobs()
    plan.jit_wait()                                     # background NVRTC build of the specialised kernel
    state, why = plan.jit_state()
    for _ in range(3):
        fn()
    ems = float(np.median(timed_steps(torch, fn, 9, flush)))
    ebytes = algorithmic_bytes(it, P, 1)
    res["hbm_bound_interpolant"] = {
        "workload": "example/1 as shipped (pw-linear, d=10, 7 terms, 22 corners/point), 10^7 points, "
                    "exact combination formula (bit-identical to the reference)",
        "kernel": "sgm_small_jit (plan-specialised, NVRTC; csrc/sgm_jit.cpp)" if state == 1 else
                  "sgm::eval_points_small_kernel (generic; specialised kernel unavailable: %s)" % why,
        "jit_state": state, "jit": why,
        "ms": ems, "algorithmic_bytes": ebytes, "achieved": ebytes / ems / 1e6, "peak": peak,
        "unit": "GB/s", "frac": ebytes / ems / 1e6 / peak, "value": P / ems * 1e3,
        "generic_kernel_ms": generic_ms,
        "generic_kernel_frac": None if generic_ms is None else ebytes / generic_ms / 1e6 / peak}
    return res


def run_indicators(torch, dev):
    """Batched Gerstner-Griebel indicators (compute_GG_indicators_RM): all reduced-margin
    candidates of an adaptive state in one segmented launch; candidates per second."""
    from sgmethods_b200 import nodes_1d
    from sgmethods_b200.compute_error_indicators import MeanSquareNorm, compute_GG_indicators_RM
    from sgmethods_b200.mid_set import MidSet
    from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
    ms = MidSet(track_reduced_margin=True)
    rng = np.random.default_rng(11)
    ms.enlarge_from_margin(0)
    for step in range(60):                               # a 6-dimensional adaptive state
        if ms.get_dim() < 6 and step % 4 == 0:
            ms.increase_dim()
        else:
            ms.enlarge_from_margin(int(rng.integers(0, ms.get_cardinality_reduced_margin())))
    knots, lev = nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1   # noqa: E731
    f = lambda y: np.array([np.exp(-0.3 * np.sum(y * y)), np.sum(y)])         # noqa: E731
    cur = SGInterpolant(ms.mid_set, knots, lev, verbose=False)
    u_on_SG = cur.sample_on_SG(f)
    yy = rng.normal(size=(20000, ms.get_dim()))
    u_interp = cur.interpolate(yy, u_on_SG)
    empty = (np.zeros((0, 0), dtype=int), np.zeros(0))
    res = {}
    for label, norm in (("gpu_norm", MeanSquareNorm()), ("host_norm_callback", lambda a, b: MeanSquareNorm()(a, b))):
        compute_GG_indicators_RM(*empty, ms, knots, lev, f, cur.SG, u_on_SG, yy, norm, u_interp)
        t0 = time.perf_counter()
        est = compute_GG_indicators_RM(*empty, ms, knots, lev, f, cur.SG, u_on_SG, yy, norm, u_interp)
        dt = time.perf_counter() - t0
        res[label] = {"candidates": int(est.size), "seconds": dt, "candidates_per_s": est.size / dt}
    res["state"] = {"dim": ms.get_dim(), "multi_indices": int(ms.get_cardinality_mid_set()),
                    "reduced_margin": int(ms.get_cardinality_reduced_margin()), "random_points": 20000,
                    "note": "wall clock of compute_GG_indicators_RM incl. host table building and "
                            "sampling f on the new nodes; one segmented launch for all candidates"}
    return res


# ------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=WORKLOADS)
    ap.add_argument("--points", type=int, default=0, help="points per rank (0 = workload default)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="points of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra keys (other workloads, HBM-bound kernels, indicators, scaling modes)")
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
        ref = reference_tables(args.workload)
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or cores * 1024
        if args.warmup > 0:
            cpu_reference_rate(ref, max(cores * 64, 64))
        runs = [cpu_reference_rate(ref, sample, seed=s) for s in range(args.steps)]
        total = sum(r[2] for r in runs)
        value = sample * ref["NF"] * len(runs) / total
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * total / len(runs),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": ref["desc"], "points_per_step": sample, "NF": ref["NF"],
                           "terms": ref["terms"], "nodes": ref["num_nodes"], "tables": ref["tables"]},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": runs[0][1], "kind": "port",
                                 "sample": runs[0][3] + "; oracle/ numpy+scipy port of the reference's "
                                 "term loop (bit-equal to the reference and of the same cost)"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "native_library_loaded": "sgmethods_b200._lib" in sys.modules}
        print(json.dumps(line), file=json_out, flush=True)
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference "
                         "for the CPU baseline")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak, peak_src = measured_peak_gbs()
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # > 126 MB L2
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    warmup = max(args.warmup, 3)

    if args.workload == "c4":            # multilevel driver: its own (simpler) line
        res = run_extra(torch, "c4", dev, flush, peak, steps=max(2, min(args.steps, 3)))
        if rank == 0:
            line = {"metric": METRIC, "value": world * res["value"], "unit": UNIT, "n_gpus": world,
                    "steps": args.steps, "warmup": warmup, "ms_per_step": res["ms_per_step"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                    "data": "synthetic", "config": {"workload": res["workload"]}, "detail": res,
                    "gpu_launches": 1}
            print(json.dumps(line), file=json_out, flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    it, desc, P_default, NF, sampler = build_workload(args.workload)
    P = args.points or P_default
    x = sample_points(torch, P, it.N, sampler, dev, gen)
    f = torch.from_numpy(np.random.default_rng(1).normal(size=(it.num_nodes, NF))).to(dev)
    exact_plan = it.device_plan(local_rank)

    # ---- the step: the public call on device-resident inputs, contract path (method="auto":
    # hierarchisation of the nodal values, per-call validation, locality sort, evaluation)
    out_holder = {}

    def step():
        out_holder["out"] = it.interpolate(x, f)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    method = getattr(it, "_auto_choice", None) or "combination"
    main_plan = it.hierarchical()._dev(local_rank)[0] if method == "hierarchical" else exact_plan
    main_plan.timing = []
    if world > 1:
        dist.barrier()
    with ClockSampler(local_rank) as clocks:
        t_wall = time.perf_counter()
        step_ms = timed_steps(torch, step, args.steps, flush)
        wall = time.perf_counter() - t_wall
    if world > 1:
        dist.barrier()
    kernel_ms = main_kernel_ms(main_plan, P)
    main_plan.timing = None
    total_ms = max_over_ranks(torch, dist, dev, step_ms.sum(), world)
    value = world * P * NF * args.steps / (total_ms * 1e-3)
    out = out_holder["out"].reshape(P, NF).clone()

    # ---- the bit-exact combination formula on the same batch (extra key, not the headline
    # when the validated work-reduced form is in use)
    exact_steps = max(3, min(args.steps, 5))
    exact_out = torch.empty((P, NF), dtype=torch.float64, device=dev)
    exact_plan.eval(x, f, exact_out)
    exact_plan.timing = []
    exact_ms = timed_steps(torch, lambda: exact_plan.eval(x, f, exact_out), exact_steps, flush)
    exact_plan.timing = None
    dev_rel = float((out - exact_out).abs().max() / exact_out.abs().max())

    # ---- end to end through the public API with pinned host buffers; with several ranks the
    # final NCCL gather of the (P, NF) slabs to rank 0 and rank 0's D2H of the whole result are part
    # of it (north star item 4: "points shard ... with a final gather")
    x_host = x.cpu().pin_memory()
    f_host = f.cpu().pin_memory()
    out_host = torch.empty((world * P, NF), dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    sh = None
    if world > 1:
        from sgmethods_b200.distributed import ShardedInterpolant
        sh = ShardedInterpolant(it, mode="points")

    def e2e_step():
        res = it.interpolate(x_host, f_host).reshape(P, NF)      # H2D + kernels, CUDA tensor
        if sh is not None:
            res = sh.gather_blocks(res, world * P, "root")       # NCCL gather of the slabs to rank 0
        if res is not None:
            out_host[:res.shape[0]].copy_(res, non_blocking=True)    # D2H of the (gathered) result, rank 0
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
    e2e_ms = max_over_ranks(torch, dist, dev, e0.elapsed_time(e1), world)
    e2e_value = world * P * NF * e2e_steps / (e2e_ms * 1e-3)
    if rank == 0:
        assert torch.equal(out_host[:P].to(dev), out), "e2e result differs from the device-resident run"
    numpy_ms = None
    if world == 1:
        # the drop-in call exactly as a reference user makes it: numpy in, numpy out (pageable
        # host memory, staged through pinned buffers by the library); wall clock
        x_np, f_np = x_host.numpy(), f_host.numpy()
        it.interpolate(x_np, f_np)
        t_np = time.perf_counter()
        for _ in range(3):
            res_np = it.interpolate(x_np, f_np)
        numpy_ms = (time.perf_counter() - t_np) / 3 * 1e3
        assert np.array_equal(res_np.reshape(P, NF), out_host[:P].numpy())

    gather_ms = None
    if world > 1:
        sh.gather_blocks(out, world * P, "all")
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(5):
            full = sh.gather_blocks(out, world * P, "all")
        g1.record()
        torch.cuda.synchronize()
        gather_ms = max_over_ranks(torch, dist, dev, g0.elapsed_time(g1) / 5, world)
        assert full.shape == (world * P, NF)

    scaling_extras = {}
    if not args.no_extras and args.workload == "c2":
        scaling_extras = run_scaling_modes(torch, dist, dev, rank, world, flush)

    if rank == 0:
        clk = clocks.summary()
        kernel_ms = kernel_ms if kernel_ms is not None else float(np.mean(step_ms))
        bytes_alg = algorithmic_bytes(it, P, NF)
        achieved = bytes_alg / (kernel_ms * 1e-3) / 1e9
        corners = exact_plan.corners_per_point
        hier = method == "hierarchical"
        gathers = it.hierarchical()._term_fam.shape[0] if hier else corners
        fp64 = measured_fp64()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc, "points_per_gpu": P, "NF": NF, "N": it.N,
                       "terms": len(it._coeffs), "nodes": it.num_nodes,
                       "corners_per_point": corners, "l2": "flushed between timed steps "
                       "(256 MB fill)", "sharding": "points (weak scaling: every rank its own batch); "
                       "no data-path collective in the step, final NCCL gather to rank 0 inside e2e",
                       "step": "SGInterpolant.interpolate(x, f) on device-resident tensors, method='auto'",
                       "method_in_use": method, "auto_validation": getattr(it, "auto_validation", None),
                       "max_rel_deviation_from_exact_combination_formula": dev_rel},
            # per step: hier_apply_all, validation sample (2 x (pregather + eval) + deviation kernel),
            # pregather + grouped kernel of the batch (no sort) -- or pregather, sort hist / scan /
            # scatter, eval on the exact path
            "gpu_launches": (8 if hier else 5) * args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(x_host.numel() * 8 + f_host.numel() * 8),
                    "d2h_bytes_per_step": int(world * P * NF * 8), "steps": e2e_steps,
                    "includes_final_gather_to_rank0": world > 1, "host_numa_binding": numa},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": NCU_TRAFFIC.get((args.workload, P, method)),
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the kernel, "
                                           "ncu --set full capture of this command (profiles/)",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_alg,
                         "kernel": ("sgm::eval_hier_groups_kernel<16 warps, 2 blocks/SM> (grouped "
                                    "hierarchical-surplus kernel)" if hier else
                                    "sgm::eval_points_prefix_kernel / split kernel (exact combination formula)")
                                   + " + pre-gather and point locality sort (< 5 % of the launch group)",
                         "kernel_ms": kernel_ms,
                         "note": "gather-bound, not HBM-bound: per point " + str(gathers) +
                                 (" surplus gathers (one per multi-index)" if hier else " corner gathers") +
                                 " of 8 bytes from an L1/L2-resident table against "
                                 f"{bytes_alg / P:.0f} algorithmic HBM bytes (DESIGN.md section 5)",
                         "gathers_per_s": P * gathers / (kernel_ms * 1e-3),
                         "secondary": {"l1_data_pipe_floor_ms": P * gathers * 16 / (128 * 148 * 1.965e9) * 1e3
                                       if hier else secondary_rooflines(P, corners, 0, kernel_ms, clk)["l1_data_pipe_floor_ms"],
                                       "note": "L1 data pipe 128 B/clk/SM; a warp-wide 64-bit gather "
                                               "needs >= 2 wavefronts"}},
            "exact_mode": {"value": P * NF / (float(np.mean(exact_ms)) * 1e-3), "unit": UNIT,
                           "ms_per_step": float(np.mean(exact_ms)), "steps": exact_steps,
                           # separately rounded multiply + add per corner (weight prefix reused: one
                           # more multiply), against the measured DMUL+DADD pair rate
                           "fp64_floor_ms": None if not fp64 else
                           1e3 * P * corners * 1.5 / fp64["dmul_dadd_pairs_per_s"],
                           "hbm_frac": bytes_alg / (float(np.mean(exact_ms)) * 1e-3) / 1e9 / peak,
                           "corner_gathers_per_s": P * corners / (float(np.mean(exact_ms)) * 1e-3),
                           "note": "method='combination': the reference's formula bit for bit "
                                   "(sgm_eval on the same batch)"},
            "clocks": clk,
            "wall_s_timed_region": wall,
            "fp64_pipe_measured": fp64,
        }
        if numpy_ms is not None:
            line["e2e"]["numpy_in_numpy_out"] = {
                "value": P * NF / (numpy_ms * 1e-3), "ms_per_call": numpy_ms,
                "note": "SGInterpolant.interpolate(ndarray, ndarray) -> ndarray, pageable memory, wall clock"}
        if gather_ms is not None:
            line["final_all_gather"] = {"ms": gather_ms, "bytes_per_rank": int(P * NF * 8),
                                        "backend": "nccl all_gather_into_tensor",
                                        "note": "timed alone (the variant for callers that need the result on every rank); "
                                                "the e2e step gathers to rank 0 only"}
        line.update(scaling_extras)
        if world == 1 and not args.no_extras and args.workload == "c2":
            line.update(run_hbm_kernels(torch, dev, flush, peak, gen))
            line["extra_workloads"] = {}
            for name in ("c2leja", "c5", "c3full", "c4"):
                try:
                    line["extra_workloads"][name] = run_extra(torch, name, dev, flush, peak)
                except Exception as exc:                 # keep the headline line whatever happens
                    line["extra_workloads"][name] = {"error": repr(exc)}
                torch.cuda.empty_cache()
            try:
                line["error_indicators"] = run_indicators(torch, dev)
            except Exception as exc:
                line["error_indicators"] = {"error": repr(exc)}
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N=1 only
            ref = reference_tables(args.workload)
            cores = os.cpu_count() or 1
            sample = args.cpu_sample or cores * 1024
            v, used, dt, what = cpu_reference_rate(ref, sample)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": used, "kind": "port",
                                    "sample": what + f" ({dt:.1f} s)"}
            # the reference as shipped is single-threaded: the same port on ONE core
            v1, _, dt1, _ = cpu_reference_rate(ref, 1024, cores=1)
            line["cpu_baseline"]["single_core"] = {"value": v1, "cores": 1,
                                                   "sample": f"1024 points, 1 process ({dt1:.1f} s)"}
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_scaling_modes(torch, dist, dev, rank, world, flush):
    """Two extra measurements that say something about multi-GPU behaviour (also run at N = 1 so
    that the N > 1 lines have their baseline):
      strong_scaling_c5  config 5 (d=128, 91 929 terms), a FIXED 1.6x10^7-point batch split over
                         the ranks, result gathered to rank 0 (NCCL gather), contract path;
      terms_mode_c5      the same interpolant at 4096 points with the TERMS sharded over the ranks
                         and an NCCL all-reduce of the partial sums; deviation from the
                         single-GPU result reported."""
    from sgmethods_b200.distributed import ShardedInterpolant, shard_bounds
    it, desc, _, NF, sampler = build_workload("c5")
    f = torch.from_numpy(np.random.default_rng(1).normal(size=(it.num_nodes, 1))).to(dev)
    res = {}
    # ---- strong scaling: fixed total batch, every rank generates the SAME full point set from
    # one seed and evaluates its block through the contract path (method='auto'; a point's value
    # does not depend on the batch it is in, so the checksum is the same for every N)
    P_total = 16 * 10 ** 6
    b = shard_bounds(P_total, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    g = torch.Generator(device=dev).manual_seed(4242)
    CH = 10 ** 6
    xs = []
    for c0 in range(0, P_total, CH):                       # same stream of random numbers on all ranks
        blk = torch.randn((min(CH, P_total - c0), it.N), dtype=torch.float64, device=dev, generator=g)
        a, e = max(lo, c0), min(hi, c0 + blk.shape[0])
        if e > a:
            xs.append(blk[a - c0:e - c0].clone())
        del blk
    x_loc = torch.cat(xs) if xs else torch.empty((0, it.N), dtype=torch.float64, device=dev)
    del xs
    sh = ShardedInterpolant(it, mode="points")
    plan = it.device_plan(dev.index)
    warm = sh.interpolate_local(x_loc[:16384], f)            # plans, validation, NCCL connections
    sh.gather_blocks(warm, 16384 * world, "root")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    loc = sh.interpolate_local(x_loc, f)
    e1.record()
    full = sh.gather_blocks(loc, P_total, "root")
    e2.record()
    torch.cuda.synchronize()
    t_eval = max_over_ranks(torch, dist, dev, e0.elapsed_time(e1), world)
    t_all = max_over_ranks(torch, dist, dev, e0.elapsed_time(e2), world)
    checksum = float(full.sum().item()) if rank == 0 else 0.0
    res["strong_scaling_c5"] = {
        "workload": desc, "points_total": P_total, "n_gpus": world, "scaling": "strong",
        "method": getattr(it, "_auto_choice", None) or "combination",
        "eval_ms": t_eval, "eval_plus_gather_to_rank0_ms": t_all, "gather_ms": t_all - t_eval,
        "value": P_total / (t_all * 1e-3), "unit": UNIT, "checksum": checksum,
        "limits": "evaluation kernel; the NCCL gather of 8 B per point to rank 0 is the part that "
                  "does not shrink with N"}
    # ---- term sharding: few points, huge term set
    P_small = 4096
    x_small = torch.randn((P_small, it.N), dtype=torch.float64, device=dev,
                          generator=torch.Generator(device=dev).manual_seed(99))
    single = plan.eval(x_small, f)
    st = ShardedInterpolant(it, mode="terms")
    st.interpolate(x_small, f)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(3):
        part = st.interpolate(x_small, f)
    t1.record()
    torch.cuda.synchronize()
    t_terms = max_over_ranks(torch, dist, dev, t0.elapsed_time(t1) / 3, world)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(3):
        plan.eval(x_small, f)
    s1.record()
    torch.cuda.synchronize()
    t_single = max_over_ranks(torch, dist, dev, s0.elapsed_time(s1) / 3, world)
    dev_rel = float((part - single).abs().max() / single.abs().max())
    res["terms_mode_c5"] = {
        "workload": desc, "points": P_small, "n_gpus": world, "terms_per_rank": int(len(it._coeffs) // world),
        "ms_terms_sharded_incl_allreduce": t_terms, "ms_single_gpu_all_terms": t_single,
        "speedup": t_single / t_terms, "max_rel_deviation_from_single_gpu": dev_rel,
        "collective": "nccl all_reduce (sum, fp64) of the (P, NF) partial sums" if world > 1 else "none (1 rank)",
        "limits": "evaluation kernel per rank; the all-reduce moves 32 KB"}
    del x_loc, loc, full
    torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    sys.exit(main())
