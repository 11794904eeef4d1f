"""Row-sharded multi-GPU path (one process per GPU, NCCL) against the unsharded oracle.
Needs >= 2 GPUs: skipped on the single-GPU box, run with `gpurun --gpus 2`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_sharded_pipelines_match_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    # torchrun exports OMP_NUM_THREADS=1 unless set: rank 0 runs the CPU oracle (2^19 x 8192 f32 at the config-4 shape)
    env = dict(os.environ, OMP_NUM_THREADS=str(max(1, (os.cpu_count() or 2) // world)))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=2400, env=env)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "multi_gpu_worker.log"), "w") as f:
            f.write(r.stdout + "\n--- stderr ---\n" + r.stderr[-20000:] + f"\nrc={r.returncode}\n")
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout
