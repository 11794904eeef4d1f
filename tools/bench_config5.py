#!/usr/bin/env python
"""BASELINE config 5: two-sided interpolative decomposition, c64, 16384 x 16384 low-rank kernel matrix, rank 128,
at 1 and 8 B200 (row-sharded; every rank generates its own rows on device).

    python tools/bench_config5.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29657 \
        tools/bench_config5.py

One JSON line on rank 0: ms per pipeline pass (max over ranks, CUDA events), algorithmic GFLOP/s
(GEMM 8 m n (l + k)) and the ID errors on a Gaussian probe."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from rusty_compression_b200 import api

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--k", type=int, default=128)
ap.add_argument("--p", type=int, default=10)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
ctx = api.Context(device=local)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)
n, k, p = args.n, args.k, args.p
rows = n // world
a = api.helmholtz_kernel_matrix((rows, n), np.complex128, seed=7, row_offset=rank * rows, ctx=ctx)
if world > 1:
    a.set_shard(n, rank * rows)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def step():
    q = api.sample_range_by_rank(a, k, p, seed=42, ctx=ctx, device=True)
    cid = api.QR.compute_from_range_estimate(q, a).compress(api.RANK(k)).column_id()
    return cid, cid.two_sided_id()


for _ in range(args.warmup):
    cid, ts = step()
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(args.steps):
    cid, ts = step()
e1.record(stream)
barrier()
ms = e0.elapsed_time(e1) / args.steps
if world > 1:
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
# probe errors ||(A - C Z) x|| / ||A x|| and the same for C X R, on the local rows, summed over the ranks
x = api.DeviceMatrix.random_gaussian((n, 4), np.complex128, 3, ctx=ctx)
xn = x.to_numpy()
ax = a.matmat(x).to_numpy()
num_c = np.linalg.norm(ax - cid.c.dot(cid.z.dot(xn))) ** 2
num_t = np.linalg.norm(ax - ts.c.dot(ts.x.dot(ts.r.dot(xn)))) ** 2
den = np.linalg.norm(ax) ** 2
v = torch.tensor([num_c, num_t, den], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(v)
l = k + p
flops = 8.0 * n * n * (l + k)
if rank == 0:
    print(json.dumps({"config": "configs[4]: two-sided ID, c64, kernel matrix", "n": n, "k": k, "p": p, "n_gpus": world,
                      "rows_per_gpu": rows, "ms_per_pass": ms, "algorithmic_gflops": flops / ms / 1e6,
                      "column_id_probe_error": float(torch.sqrt(v[0] / v[2])), "two_sided_id_probe_error": float(torch.sqrt(v[1] / v[2])),
                      "skeleton_rows_first8": [int(i) for i in ts.row_ind[:8]], "skeleton_cols_first8": [int(i) for i in ts.col_ind[:8]],
                      "scaling": "strong"}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
