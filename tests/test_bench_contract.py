"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the keys the
driver reads, the flop model is the one SURVEY.md 8(d) states, and the CUDA arm refuses to run without a device
(no CPU fallback)."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_module", BENCH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_flop_and_byte_model_match_the_survey():
    b = load_bench()
    m, n, k, p, it = 65536, 8192, 64, 10, 2
    l = k + p
    flops = b.algorithmic_flops(m, n, k, p, it)
    gemm = 2.0 * m * n * (l * (1 + 2 * it) + k)                 # 4.66e11
    assert abs(gemm - 4.66e11) / 4.66e11 < 0.01
    assert gemm < flops < 1.02 * gemm and abs(flops - 4.72e11) / 4.72e11 < 0.01
    assert abs(b.algorithmic_bytes(m, n, k, p, it) - 2.58e10) / 2.58e10 < 0.05   # 6 passes over the 4 GiB matrix
    cfg = b.workload_config(1)
    assert cfg["workload"].startswith("configs[1]") and "model" not in cfg


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "1", "--m", "2048",
                        "--n", "512", "--skip-gemv"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    lines = [x for x in r.stdout.splitlines() if x.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rsvd_f64_algorithmic_gflops" and d["unit"] == "GFLOP/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["m_per_gpu"] == 2048 and d["config"]["n"] == 512


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cuda_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "0", "--skip-cpu", "--skip-e2e"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr


def test_profile_tools_read_the_committed_launch_list():
    """tools/step_breakdown.py and tools/stage_roofline.py over profiles/r1_launches_final.csv (the ncu launch list of
    the final build): the kernel's share of the step and the binding roof quoted in DESIGN.md come from these."""
    csv_path = os.path.join(ROOT, "profiles", "r1_launches_final.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "step_breakdown.py"), csv_path],
                         capture_output=True, text=True, check=True).stdout
    assert "dmma_gemm_kernel" in out.splitlines()[1]          # the dominant kernel leads the list
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stage_roofline.py"), csv_path],
                         capture_output=True, text=True, check=True).stdout
    rows = [l for l in out.splitlines() if l.startswith("Y = A Omega") and "x3" in l]
    assert len(rows) == 1 and "(0.9" in rows[0]               # three passes, ~0.90 of the FP64 pipe
    shares = [float(l.split("%")[0].split()[-1]) for l in out.split("aggregated by stage kind")[1].splitlines() if "%" in l]
    assert abs(sum(shares) - 100.0) < 0.5
