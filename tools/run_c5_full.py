"""BASELINE config 5 at full size: d=128, 91 929 combination terms, 10^8 evaluation points
sharded over the ranks (torchrun, one process per GPU, NCCL), scalar output.

Every rank generates its own block of points on the device (the full point set is 102 GB),
evaluates it in chunks through ShardedInterpolant.interpolate_local (no data-path
collective), and the (P/G,) output slabs are assembled with one NCCL all-gather.  Prints one
JSON line on rank 0.  Usage:
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_c5_full.py [P_total]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200.distributed import ShardedInterpolant, shard_bounds
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant

sys.stdout.flush()
json_out = os.fdopen(os.dup(1), "w")      # stdout = the JSON line only (NCCL prints to fd 1)
os.dup2(2, 1)
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
P_total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 8
t0 = time.time()
S = mis.aniso_smolyak_mid_set(3, 128, np.array([1.0] * 80 + [3.0] * 48))
it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False)
t_ctor = time.time() - t0
b = shard_bounds(P_total, world)
P = int(b[rank + 1] - b[rank])
f = torch.from_numpy(np.random.default_rng(1).normal(size=(it.num_nodes, 1))).to(dev)
sh = ShardedInterpolant(it, mode="points")
CH = 10 ** 6
gen = torch.Generator(device=dev).manual_seed(100 + rank)
out = torch.empty((P, 1), dtype=torch.float64, device=dev)
x = torch.empty((CH, 128), dtype=torch.float64, device=dev)
plan = it.device_plan(local)
# warm-up (tables to the device, kernel attributes)
x.normal_(generator=gen); plan.eval(x[:4096], f, out[:4096])
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t_eval = 0.0
wall0 = time.perf_counter()
for lo in range(0, P, CH):
    hi = min(P, lo + CH)
    x.normal_(generator=gen)                      # synthetic points of this chunk (untimed part below)
    e0.record()
    plan.eval(x[:hi - lo], f, out[lo:hi])
    e1.record()
    torch.cuda.synchronize()
    t_eval += e0.elapsed_time(e1) * 1e-3
g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g0.record()
full = sh.gather_blocks(out, P_total, "all")
g1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - wall0
t = torch.tensor([t_eval, g0.elapsed_time(g1) * 1e-3, wall], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ones = plan.eval(x[:65536], torch.ones_like(f))
unity = float((ones - 1.0).abs().max())
assert full.shape[0] == P_total and bool(torch.isfinite(full).all())
if rank == 0:
    print(json.dumps({"case": "config 5 full size", "n_gpus": world, "points_total": P_total,
                      "N": 128, "terms": len(it._coeffs), "nodes": it.num_nodes,
                      "corners_per_point": plan.corners_per_point, "ctor_s": t_ctor,
                      "eval_s_max_over_ranks": float(t[0]), "all_gather_s": float(t[1]),
                      "wall_s_incl_point_generation": float(t[2]),
                      "point_evals_per_s": P_total / float(t[0]),
                      "partition_of_unity_max_dev": unity}), file=json_out, flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
