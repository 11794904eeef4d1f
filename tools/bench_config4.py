#!/usr/bin/env python
"""BASELINE config 4: row-sharded tall-skinny range finder + TSQR, f32, 2^23 x 8192, rank 256 (+10), at
2/4/8 B200 (strong scaling: the global matrix is fixed, every rank generates its own 2^23/P rows on device).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29655 \
        tools/bench_config4.py [--log2m 23] [--steps 3] [--warmup 2]
    python tools/bench_config4.py --log2m 20          # one GPU: the per-GPU share of the 8-GPU run

Prints one JSON line on rank 0: ms per pass (max over ranks, CUDA events), algorithmic GFLOP/s
(GEMM 2 m n l + tall QR 2 (2 m l^2 - 2/3 l^3)), and the size-independent checks
(|Q^H Q - I|, range residual on a Gaussian probe)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from rusty_compression_b200 import api

ap = argparse.ArgumentParser()
ap.add_argument("--log2m", type=int, default=23)
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--k", type=int, default=256)
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
m, n, k, p = 1 << args.log2m, args.n, args.k, args.p
l = k + p
rows = m // world
t0 = time.perf_counter()
a = api.tall_shard_matrix(rank * rows, rows, n, np.float32, 9, m, r0=512, decade_every=64.0, ctx=ctx)
if world > 1:
    a.set_shard(m, rank * rows)
ctx.synchronize()
gen_s = time.perf_counter() - t0


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def step():
    return api.sample_range_by_rank(a, k, p, seed=42, ctx=ctx, device=True)


for _ in range(args.warmup):
    q = step()
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ctx.reset_counters()
e0.record(stream)
for _ in range(args.steps):
    q = step()
e1.record(stream)
barrier()
ms = e0.elapsed_time(e1) / args.steps
launches = ctx.counter("kernel_launches") // args.steps
used, fallbacks = ctx.counter("cholqr_used"), ctx.counter("cholqr_fallbacks")
if world > 1:
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
# ---- size-independent checks: orthonormality (Gram matrix summed over the shards) and the range residual
qd = torch.from_numpy(q.to_numpy()).cuda().double()
gram = qd.T @ qd
x = api.DeviceMatrix.random_gaussian((n, 4), np.float32, 3, ctx=ctx)
ax = torch.from_numpy(a.matmat(x).to_numpy()).cuda().double()         # local rows of A x
qtax = qd.T @ ax
nrm = (ax * ax).sum().reshape(1)
if world > 1:
    dist.all_reduce(gram); dist.all_reduce(qtax); dist.all_reduce(nrm)
res = ax - qd @ qtax
rn = (res * res).sum().reshape(1)
if world > 1:
    dist.all_reduce(rn)
orth = float((gram - torch.eye(k, device="cuda", dtype=torch.float64)).abs().max())
probe = float(torch.sqrt(rn / nrm))
flops = 2.0 * m * n * l + 2.0 * (2.0 * m * l * l - 2.0 / 3.0 * l ** 3)
if rank == 0:
    print(json.dumps({"config": "configs[3]: row-sharded tall-skinny range finder + TSQR, f32", "m": m, "n": n, "k": k, "p": p,
                      "n_gpus": world, "rows_per_gpu": rows, "ms_per_pass": ms, "algorithmic_gflops": flops / ms / 1e6,
                      "algorithmic_gflops_per_gpu": flops / ms / 1e6 / world, "hbm_gbs_algorithmic": (m * n * 4.0 + 3.0 * m * l * 4) / ms / 1e6,
                      "kernel_launches_per_pass": int(launches), "cholqr2_panels": int(used), "householder_fallbacks": int(fallbacks),
                      "orthonormality_max_abs": orth, "probe_residual": probe, "generate_s": gen_s, "scaling": "strong"}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
