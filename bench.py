#!/usr/bin/env python
"""Benchmark of the randomized low-rank hot path (BASELINE.json metric, configs[1]).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle)

One "step" = one rank-64 (+10) randomized SVD of a 65536 x 8192 f64 matrix with an exponentially
decaying spectrum and 2 power iterations, exactly as the reference executes it
(src/random_sampling.rs:131-160 then src/svd.rs:171-183; every power-iteration trip is executed,
quirk Q1 included).  At N > 1 every rank owns a 65536-row block of a taller matrix (weak scaling,
row-sharded; TSQR R-factors all-gathered, A^H Q partials all-reduced over NCCL).

Prints ONE JSON line on rank 0.  `value` = algorithmic GFLOP/s with A resident in HBM;
`e2e` = the same through the C ABI with pinned host buffers, every step's H2D of A (4 GiB) and D2H of U, s, Vt inside
the timed region, run as a pipeline of depth 2 (rc_matrix_from_host_async: the upload of the next step's operator
overlaps the kernels of the current step); `e2e.one_step_at_a_time` is the same with the blocking upload.

BASELINE's metric is "rSVD/ID ... at 1/2/4/8 B200", so the same line carries the ID half and the sharded configs as
sub-records (each timed like `value`: CUDA events on the launching stream, max over ranks; not part of `value`):
  id_config3      N = 1: configs[2], column ID via pivoted QR on the sketch, f32 32768^2, tol 1e-4 (+ its own cpu_baseline)
  id_config5      every N: configs[4], two-sided ID, c64 16384^2, rank 128, rows sharded over the N GPUs (strong scaling)
  config4_strong  N >= 2: configs[3], row-sharded tall-skinny range finder, f32 2^23 x 8192, rank 256 (strong scaling)
  sharded_parity  N >= 2: the row-sharded result against the unsharded CUDA result on the same (small) matrix
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

if "TORCHELASTIC_RUN_ID" in os.environ and os.environ.get("OMP_NUM_THREADS") == "1":
    # torch.distributed.run exports OMP_NUM_THREADS=1 unless the caller set it; the CPU baseline / reference arm
    # (rank 0 only) must use all host cores, so undo that default before OpenBLAS is loaded
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# rows of the bounded CPU sample: one GPU's full share of the workload (65536 rows, ~2 s per pass on 16 cores,
# best of 3 in the cpu_baseline leg); the n-sized QRs and the SVD do not shrink with the sample, so small
# samples would understate the CPU
CPU_SAMPLE_ROWS = 65536

CFG = dict(m=65536, n=8192, k=64, p=10, it=2, r0=512, decade_every=16.0, seed=1234, omega_seed=42)


def algorithmic_flops(m, n, k, p, it):
    """GEMMs once at 2MNK, tall QR at 2ml^2 - 2/3 l^3 for the factor plus the same
    for forming Q, as executed by the reference (all `it` trips)."""
    l = k + p
    gemm = 2.0 * m * n * (l * (1 + 2 * it) + k)
    qr = lambda rows, w: 2.0 * (2.0 * rows * w * w - (2.0 / 3.0) * w ** 3)
    qrs = (it + 1) * qr(m, l) + it * qr(n, l)
    svd = qr(n, k) + 12.0 * k ** 3
    return gemm + qrs + svd + 2.0 * m * k * k


def algorithmic_bytes(m, n, k, p, it, es=8):
    """One read of A per product with A, plus read/write of the skinny operands."""
    passes = 1 + 2 * it + 1
    l = k + p
    return passes * m * n * es + (2 * passes + 6) * (m + n) * l * es


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t_mark = index, [], None, 0.0

    def start(self):
        """Started BEFORE the warm-up steps: the first nvidia-smi on a fresh box takes seconds to come up and
        would otherwise disturb the timed region; only samples taken after mark() are reported."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark(self):
        self.t_mark = time.monotonic()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.monotonic()
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if self.t_mark <= t <= t_end + 0.05] or [r for _, r in self.rows[-2:]]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, flag in zip(names, r[5:9]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def own_fp64_peaks():
    exe = os.path.join(ROOT, "rusty_compression_b200", "build", "rc_peaks")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception as e:  # pragma: no cover
        return {"error": str(e)}


# ----------------------------------------------------------------------------------- CPU arm
def run_cpu_pipeline(a, omega, k, p, it, route):
    from oracle import reference_path as ref
    q = ref.sample_range_power_iteration(a, k, p, it, ref.OmegaStream(a.dtype, blocks=[omega]), route=route)
    svd = ref.SVD.compute_from_range_estimate(q, a, route=route)
    return svd


_CPU_INPUTS = {}


def cpu_sample(m_sample, route, reps=1):
    """The oracle (the reference's LAPACK path restated) on a row sample of the workload."""
    from oracle.inputs import decaying_spectrum_matrix
    from oracle.philox import random_gaussian
    c = CFG
    key = (m_sample, c["n"])
    if key not in _CPU_INPUTS:      # input generation is untimed; keep it across steps
        a, _ = decaying_spectrum_matrix(m_sample, c["n"], np.float64, c["seed"], r0=c["r0"], decade_every=c["decade_every"])
        omega = random_gaussian((c["n"], c["k"] + c["p"]), np.float64, c["omega_seed"])
        _CPU_INPUTS.clear()
        _CPU_INPUTS[key] = (a, omega)
    a, omega = _CPU_INPUTS[key]
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        run_cpu_pipeline(a, omega, c["k"], c["p"], c["it"], route)
        best = min(best, time.perf_counter() - t0)
    flops = algorithmic_flops(m_sample, c["n"], c["k"], c["p"], c["it"])
    return flops / best / 1e9, best


def reference_arm(args, result_out):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port, the
    Rust crate cannot be built here), all host threads, bounded row sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    m_sample = min(CPU_SAMPLE_ROWS, CFG["m"])
    c = CFG
    vals, faithful = [], None
    for i in range(args.warmup + args.steps):
        gf, sec = cpu_sample(m_sample, "gemm")
        if i >= args.warmup:
            vals.append((gf, sec))
    if not args.skip_gemv:
        faithful, _ = cpu_sample(2048, "gemv")
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([s for _, s in vals])) * 1e3
    sample = (f"{m_sample} of {c['m']} rows (same n, k, p, it, spectrum); one GEMM per product (best-case CPU); "
              f"oracle = scipy LAPACK ?geqp3/?orgqr/?gesdd on OpenBLAS"
              + (f"; at --gpus {args.gpus} this arm still times ONE {m_sample}-row shard on the host cores (a rate, not the "
                 f"{args.gpus}-shard job)" if args.gpus > 1 else ""))
    line = {"impl": "reference", "metric": "rsvd_f64_algorithmic_gflops", "value": value, "unit": "GFLOP/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample,
                             "reference_faithful_gemv_route_gflops": faithful},
            "e2e": {"value": value, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    result_out.write(json.dumps(line) + "\n")
    result_out.flush()
    return 0


def workload_config(n_gpus):
    c = CFG
    return {"workload": f"configs[1]: rank-{c['k']} (+{c['p']}) randomized SVD, f64, {c['m']}x{c['n']} per GPU, "
                        f"sigma_j = 10^(-j/{c['decade_every']:g}), {c['it']} power iterations (as executed by the reference)",
            "m_per_gpu": c["m"], "n": c["n"], "k": c["k"], "p": c["p"], "it_count": c["it"],
            "global_rows": c["m"] * n_gpus, "parallelism": f"row-sharded x{n_gpus}" if n_gpus > 1 else "single GPU",
            "l2_policy": "inputs (4 GiB per pass) exceed the 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------------- CUDA arm
def pin_to_gpu_numa_node(index):
    """Best effort: run the host thread (and so first-touch the pinned staging buffers) on the CPUs of the
    NUMA node the GPU hangs off.  A pinned buffer on the far socket costs the e2e leg a factor ~2.5 in H2D
    bandwidth on two-socket hosts (seen as 271 vs 112 ms per e2e step between two otherwise identical runs)."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(index)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                lo, hi = part.split("-"); cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return allowed, sorted(cpus)
        return allowed, None
    except Exception:
        return None, None


# ----------------------------------------------------------------------------------- sub-records
CFG3 = dict(n=32768, s=64, tol=1e-4, r0=1024, decade_every=64.0, seed=1235, omega_seed=42)
CFG5 = dict(n=16384, k=128, p=10, seed=7, kappa=20.0, shift=1.5, omega_seed=42)
CFG4 = dict(log2m=23, n=8192, k=256, p=10, r0=512, decade_every=64.0, seed=9, omega_seed=42)


def config3_work(n, r_final, s, rank_id):
    """SURVEY 8(d) config 3 as EXECUTED here: A Omega on r + s columns and A^H q on r columns (B = Q^H A of the sampler
    is reused by compute_from_range_estimate, so the reference's third product with A is not executed and not
    counted), re-projections, pivoted QR of the r x n factor, the triangular solves of the column ID."""
    m = n
    gemm = 2.0 * m * n * (2 * r_final + s)
    other = 2.0 * m * r_final ** 2 + n * r_final ** 2 + 2.0 * n * r_final ** 2 + n * rank_id ** 2
    passes = 2 * (r_final // s) + 1
    return gemm + other, passes * m * n * 4.0, passes, gemm


def bench_config3(api, ctx, timed, peaks, hbm_peak, cpu):
    import numpy as np
    c = CFG3
    n, s, tol = c["n"], c["s"], c["tol"]
    a = api.decaying_spectrum_matrix((n, n), np.float32, c["seed"], r0=c["r0"], decade_every=c["decade_every"], ctx=ctx)
    out = {}

    def step():
        q, hist = api.sample_range_adaptive(a, tol, s, seed=c["omega_seed"], ctx=ctx, device=True)
        qrc = api.QR.compute_from_range_estimate(q, a).compress(api.ADAPTIVE(tol))
        cid = qrc.column_id()
        out["hist"], out["rank"] = hist, qrc.rank()
        return cid

    ms = timed(step, 5, 3)
    r_final, rank_id = out["hist"][-1][0], out["rank"]
    flops, bytes_, passes, gemm_flops = config3_work(n, r_final, s, rank_id)
    tf32_peak = peaks.get("tf32_umma_tflops") or 845.0
    rec = {"workload": f"configs[2]: column ID via pivoted QR on the sketch, f32, {n}x{n}, sigma_j = 10^(-j/{c['decade_every']:g}), "
                       f"sample_range_adaptive({tol:g}, {s}) -> QR::compute_from_range_estimate -> compress(ADAPTIVE({tol:g})) -> column_id()",
           "ms": ms, "adaptive_rank": int(r_final), "id_rank": int(rank_id), "passes_over_A": passes,
           "gflops": flops / ms / 1e6, "hbm_gbs": bytes_ / ms / 1e6,
           "frac_hbm": bytes_ / ms / 1e6 / hbm_peak,
           "frac_tf32": 3.0 * gemm_flops / ms / 1e9 / tf32_peak,
           "frac_note": "frac_hbm: algorithmic bytes (one 4 GiB read of A per executed product) / time / measured HBM peak; "
                        "frac_tf32: 3 x algorithmic GEMM flops (the TF32 split executes three products) / time / own-measured "
                        "tcgen05 kind::tf32 issue-rate peak; the binding roof is HBM (each 64-column pass is HBM-bound)",
           "tf32_peak_tflops_own_measured": tf32_peak, "dtype": "f32"}
    if cpu:
        from oracle import reference_path as ref
        a_host = a.to_numpy()
        t0 = time.perf_counter()
        q_ref, hist_ref = ref.sample_range_adaptive(a_host, tol, s, ref.OmegaStream(np.float32, seed=c["omega_seed"]))
        qr_ref = ref.QR.compute_from_range_estimate(q_ref, a_host).compress(ref.ADAPTIVE(tol))
        qr_ref.column_id()
        sec = time.perf_counter() - t0
        # the reference recomputes B = Q^H A in compute_from_range_estimate: its executed work has the third product
        flops_ref = flops + 2.0 * n * n * r_final
        rec["cpu_baseline"] = {"value": flops_ref / sec / 1e9, "unit": "GFLOP/s", "ms": sec * 1e3, "cores": os.cpu_count(), "kind": "port",
                               "sample": f"oracle pipeline on the full {n}x{n} f32 matrix (same A, same Philox Omega stream), one GEMM per "
                                         f"product (best-case CPU), one pass; rank history {[int(r) for r, _ in hist_ref]}"}
        del a_host
    a.free()
    return rec


def bench_config5(api, ctx, timed, peaks, world, rank, cpu):
    import numpy as np
    c = CFG5
    n, k, p = c["n"], c["k"], c["p"]
    rows = n // world
    a = api.helmholtz_kernel_matrix((rows, n), np.complex128, seed=c["seed"], kappa=c["kappa"], shift=c["shift"],
                                    row_offset=rank * rows, ctx=ctx)
    if world > 1:
        a.set_shard(n, rank * rows)

    def step():
        q = api.sample_range_by_rank(a, k, p, seed=c["omega_seed"], ctx=ctx, device=True)
        cid = api.QR.compute_from_range_estimate(q, a).compress(api.RANK(k)).column_id()
        return cid.two_sided_id()

    ms = timed(step, 5, 3)
    l = k + p
    gemm = 8.0 * n * n * (l + k)
    flops = gemm + 4.0 * 2.0 * (2.0 * n * l * l) + 2.0 * 4.0 * 2.0 * n * k * k + 4.0 * 2.0 * n * k * k      # + tall QR, two wide pivoted QRs, TRSMs
    bytes_ = 2.0 * n * n * 16.0
    fp64_peak = peaks.get("dmma_tflops") or 37.0
    rec = {"workload": f"configs[4]: two-sided ID, c64, {n}x{n} Helmholtz kernel matrix (kappa {c['kappa']:g}, box gap 0.5), rank {k} (+{p}): "
                       "sample_range_by_rank -> QR::compute_from_range_estimate -> compress(RANK) -> column_id -> two_sided_id",
           "ms": ms, "n_gpus": world, "rows_per_gpu": rows, "scaling": "strong", "gflops": flops / ms / 1e6, "hbm_gbs": bytes_ / ms / 1e6,
           "frac_fp64": flops / ms / 1e9 / (fp64_peak * world), "fp64_peak_tflops_own_measured": fp64_peak,
           "frac_note": "algorithmic flops (8 m n (l + k) for the two products with A + QRs + solves) / time / (own-measured DMMA peak x N); "
                        "FP64-pipe bound (arithmetic intensity 66 flop/B)", "dtype": "c64"}
    if cpu and world == 1:
        from oracle import reference_path as ref
        from oracle.philox import random_gaussian
        a_host = a.to_numpy()
        omega = random_gaussian((n, l), np.complex128, c["omega_seed"])
        t0 = time.perf_counter()
        q_ref = ref.sample_range_by_rank(a_host, k, p, ref.OmegaStream(np.complex128, blocks=[omega]))
        ref.QR.compute_from_range_estimate(q_ref, a_host).compress(ref.RANK(k)).column_id().two_sided_id()
        sec = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": flops / sec / 1e9, "unit": "GFLOP/s", "ms": sec * 1e3, "cores": os.cpu_count(), "kind": "port",
                               "sample": f"oracle pipeline on the full {n}x{n} c64 matrix (same A, same Omega), one GEMM per product, one pass"}
        del a_host
    a.free()
    return rec


def bench_config4(api, ctx, timed, peaks, hbm_peak, world, rank):
    import numpy as np
    c = CFG4
    m, n, k, p = 1 << c["log2m"], c["n"], c["k"], c["p"]
    l = k + p
    rows = m // world
    a = api.tall_shard_matrix(rank * rows, rows, n, np.float32, c["seed"], m, r0=c["r0"], decade_every=c["decade_every"], ctx=ctx)
    a.set_shard(m, rank * rows)
    ms = timed(lambda: api.sample_range_by_rank(a, k, p, seed=c["omega_seed"], ctx=ctx, device=True), 3, 2)
    flops = 2.0 * m * n * l + 2.0 * (2.0 * m * l * l - 2.0 / 3.0 * l ** 3)
    bytes_ = m * n * 4.0 + 3.0 * m * l * 4.0
    tf32_peak = peaks.get("tf32_umma_tflops") or 845.0
    a.free()
    return {"workload": f"configs[3]: row-sharded tall-skinny range finder + TSQR, f32, 2^{c['log2m']} x {n} (256 GiB, generated per shard on "
                        f"device), rank {k} (+{p}): sample_range_by_rank on the sharded operator",
            "ms": ms, "n_gpus": world, "rows_per_gpu": rows, "scaling": "strong", "tflops": flops / ms / 1e9,
            "hbm_gbs": bytes_ / ms / 1e6, "frac_hbm": bytes_ / ms / 1e6 / (hbm_peak * world),
            "frac_tf32": 3.0 * flops / ms / 1e9 / (tf32_peak * world), "tf32_peak_tflops_own_measured": tf32_peak,
            "frac_note": "algorithmic flops (GEMM 2 m n l + tall QR) x 3 (TF32 split) / time / (own-measured tcgen05 kind::tf32 peak x N); "
                         "strong-scaling efficiency = ms(N = 2) x 2 / (ms(N) x N) across the N = 2, 4, 8 lines", "dtype": "f32"}


def sharded_parity(api, ctx, world, rank, local):
    """N >= 2: the row-sharded pipelines against the UNSHARDED CUDA path on the same matrix (gathered on rank 0), so the
    multi-GPU results are checked on the very box the scaling numbers come from."""
    import numpy as np
    import torch
    import torch.distributed as dist
    rows, n, k, p = 4096, 1024, 48, 10
    m = rows * world
    a_loc = api.decaying_spectrum_matrix((rows, n), np.float64, 77, r0=128, decade_every=10.0, row_offset=rank * rows, ctx=ctx)
    a_loc.set_shard(m, rank * rows)
    q = api.sample_range_power_iteration(a_loc, k, p, 2, seed=5, ctx=ctx, device=True)
    s_sh = api.SVD.compute_from_range_estimate(q, a_loc).s_f64()
    q2 = api.sample_range_by_rank(a_loc, k, p, seed=5, ctx=ctx, device=True)
    cid = api.QR.compute_from_range_estimate(q2, a_loc).compress(api.RANK(k)).column_id()
    col_sh = cid.col_ind[:k].copy()
    row_sh = cid.two_sided_id().row_ind[:k].copy()
    t = torch.from_numpy(a_loc.to_numpy()).cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    rec = None
    if rank == 0:
        ctx1 = api.Context(device=local)            # no communicator: the single-GPU path
        a_full = api.DeviceMatrix.from_numpy(torch.cat(parts, 0).cpu().numpy(), ctx=ctx1)
        q1 = api.sample_range_power_iteration(a_full, k, p, 2, seed=5, ctx=ctx1, device=True)
        s_1 = api.SVD.compute_from_range_estimate(q1, a_full).s_f64()
        q3 = api.sample_range_by_rank(a_full, k, p, seed=5, ctx=ctx1, device=True)
        cid1 = api.QR.compute_from_range_estimate(q3, a_full).compress(api.RANK(k)).column_id()
        rec = {"workload": f"f64 {m}x{n} row-sharded over {world} GPUs vs the unsharded CUDA path on the gathered matrix (rank {k} (+{p}))",
               "max_rel_singular_value_deviation": float(np.max(np.abs(s_sh - s_1) / s_1)),
               "skeleton_columns_identical": bool(np.array_equal(col_sh, cid1.col_ind[:k])),
               "skeleton_rows_identical": None}
        row_1 = cid1.two_sided_id().row_ind[:k].copy()
        rec["skeleton_rows_identical"] = bool(np.array_equal(row_sh, row_1))
        # every handle of ctx1 goes before the context does
        a_full.free()
        for h in (q1, q3, cid1):
            h.free()
        del q1, q3, cid1
        import gc
        gc.collect()
        ctx1.close()
    a_loc.free()
    return rec


def _mark(msg):
    """Progress marks on stderr (never stdout: exactly one JSON line goes there)."""
    if int(os.environ.get("RANK", "0")) == 0:
        sys.stderr.write(f"[bench] {time.strftime('%H:%M:%S')} {msg}\n")
        sys.stderr.flush()


def _claim_stdout():
    """Exactly ONE line may reach stdout (the JSON result): libraries such as NCCL print banners to
    fd 1, so fd 1 is pointed at stderr for the whole run and the result goes to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    result_out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--m", type=int, default=None, help="rows per GPU (default: config)")
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-gemv", action="store_true")
    ap.add_argument("--skip-sub", action="store_true", help="skip the id_config3 / id_config5 / config4_strong sub-records")
    args = ap.parse_args()
    if args.m:
        CFG["m"] = args.m
    if args.n:
        CFG["n"] = args.n
    if args.impl == "reference":
        return reference_arm(args, result_out)

    import torch
    import torch.distributed as dist
    from rusty_compression_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"WORLD_SIZE {world} != --gpus {args.gpus}"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    all_cpus, numa_cpus = pin_to_gpu_numa_node(local)
    ctx = api.Context(device=local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)

    c = CFG
    m, n, k, p, it = c["m"], c["n"], c["k"], c["p"], c["it"]
    l = k + p
    # ---- synthetic input, generated on device (same construction as oracle/inputs.py)
    a_dev = api.decaying_spectrum_matrix((m, n), np.float64, c["seed"], r0=c["r0"], decade_every=c["decade_every"],
                                         row_offset=rank * m, ctx=ctx)
    if world > 1:
        a_dev.set_shard(world * m, rank * m)
    ctx.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        q = api.sample_range_power_iteration(a_dev, k, p, it, seed=c["omega_seed"], ctx=ctx, device=True)
        svd = api.SVD.compute_from_range_estimate(q, a_dev)
        return q, svd

    def timed(fn, steps, warmup, on_start=None):
        for _ in range(warmup):
            fn()
        barrier()
        if on_start:
            on_start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.reset_counters()
    launches_before = ctx.counter("kernel_launches")
    if world > 1:
        # The first multi-process run on a fresh box is ~35 % slow for its first second or so (NCCL channels,
        # peer mappings and the second GPU's clocks come up lazily; measured 28 vs 21 ms per step): absorb it
        # with extra untimed steps on top of the requested warm-up.
        for _ in range(12):
            step_device()
        barrier()
    _mark("timed region: configs[1] step")
    ms_step = timed(step_device, args.steps, args.warmup, on_start=sampler.mark)
    _mark(f"configs[1]: {ms_step:.3f} ms per step")
    launches = (ctx.counter("kernel_launches") - launches_before) // (args.steps + args.warmup) * args.steps
    clocks = sampler.stop() if rank == 0 else None

    flops_rank = algorithmic_flops(m, n, k, p, it)
    value = world * flops_rank / (ms_step * 1e-3) / 1e9
    hbm_gbs = world * algorithmic_bytes(m, n, k, p, it) / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (dmma_gemm_kernel, NN: Y = A Omega), timed alone, live
    omega_dev = api.DeviceMatrix.random_gaussian((n, l), np.float64, c["omega_seed"], ctx=ctx)
    ms_nn = timed(lambda: a_dev.matmat(omega_dev), 10, 3)
    y_dev = a_dev.matmat(omega_dev)
    ms_tn = timed(lambda: a_dev.conj_matmat(y_dev), 10, 3)
    hbm_peak, peak_src = measured_peaks()
    roofline = None
    if rank == 0:
        peaks = own_fp64_peaks()
        gemm_flops = 2.0 * m * n * l
        ach = gemm_flops / (ms_nn * 1e-3) / 1e12
        peak = peaks.get("dmma_tflops") or 37.0
        prof = {}
        pj = os.path.join(ROOT, "profiles", "dominant_kernel.json")
        if os.path.exists(pj):
            prof = json.load(open(pj))
        roofline = {"bound": "tensor", "kernel": "dmma_gemm_kernel<80,false,2> (FP64 tensor pipe, DMMA.8x8x4 + DFMA tail, TMA-fed)",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "peak_source": "own-measured DMMA microbenchmark (rc_peaks); MEASURED_PEAKS.json has no FP64 figure",
                    "own_measured_peaks": peaks, "traffic": prof.get("dram_bytes_per_launch"),
                    "traffic_source": "ncu --set full capture of this kernel committed as profiles/dominant_kernel.json (dram__bytes_read.sum + "
                                      "dram__bytes_write.sum per launch); not re-measured by this run",
                    "algorithmic_bytes_per_launch": m * n * 8 + (n + m) * l * 8,
                    "hbm_gbs_of_this_kernel": (m * n * 8 + (n + m) * l * 8) / (ms_nn * 1e-3) / 1e9,
                    "hbm_frac_of_measured": (m * n * 8 + (n + m) * l * 8) / (ms_nn * 1e-3) / 1e9 / hbm_peak,
                    "hbm_peak_gbs": hbm_peak, "hbm_peak_source": peak_src,
                    "ms_per_launch": ms_nn,
                    "tn_kernel": {"kernel": "dmma_gemm_kernel<80,true,2> + split-K reduce (Z = A^T Y)",
                                  "ms_per_launch": ms_tn, "achieved": gemm_flops / (ms_tn * 1e-3) / 1e12,
                                  "frac": gemm_flops / (ms_tn * 1e-3) / 1e12 / peak}}

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D of A, D2H of U, s, Vt
    e2e = None
    _mark("roofline leg done; e2e leg")
    if not args.skip_e2e:
        a_host = torch.empty((m, n), dtype=torch.float64, pin_memory=True)
        a_np = a_host.numpy()
        ctx.check(ctx.lib.rc_matrix_to_host(ctx.h, a_dev.h, a_np.ctypes.data))
        u_host = torch.empty((m, k), dtype=torch.float64, pin_memory=True).numpy()
        vt_host = torch.empty((k, n), dtype=torch.float64, pin_memory=True).numpy()
        a_dev.free()                                  # the e2e step owns its own upload

        def step_e2e():
            op = api.DeviceMatrix.from_numpy(a_np, ctx=ctx)
            if world > 1:
                op.set_shard(world * m, rank * m)
            q = api.sample_range_power_iteration(op, k, p, it, seed=c["omega_seed"], ctx=ctx, device=True)
            svd = api.SVD.compute_from_range_estimate(q, op)
            ctx.check(ctx.lib.rc_matrix_to_host(ctx.h, ctx.lib.rc_svd_get_u(svd.h), u_host.ctypes.data))
            ctx.check(ctx.lib.rc_matrix_to_host(ctx.h, ctx.lib.rc_svd_get_vt(svd.h), vt_host.ctypes.data))
            s = svd.s_f64()
            op.free()
            return s

        ms_e2e_serial = timed(step_e2e, max(2, min(args.steps, 3)), 2)
        _mark(f"e2e, one step at a time: {ms_e2e_serial:.2f} ms per step; pipelined leg")

        # The same steps as a pipeline of depth 2 (rc_matrix_from_host_async): the upload of step i+1's operator is
        # queued on the context's copy stream before step i's kernels are, so the PCIe transfer runs under the compute.
        # Every step still uploads its own 4 GiB from pinned host memory and downloads its own U, s, Vt inside the
        # timed region (the first upload is exposed, the rest hide the compute).
        def upload_async():
            op = api.DeviceMatrix.from_numpy_async(a_np, ctx=ctx)
            if world > 1:
                op.set_shard(world * m, rank * m)
            return op

        def run_pipelined(steps):
            nxt = upload_async()
            for i in range(steps):
                op, nxt = nxt, (upload_async() if i + 1 < steps else None)
                op.await_upload()
                q = api.sample_range_power_iteration(op, k, p, it, seed=c["omega_seed"], ctx=ctx, device=True)
                svd = api.SVD.compute_from_range_estimate(q, op)
                ctx.check(ctx.lib.rc_matrix_to_host(ctx.h, ctx.lib.rc_svd_get_u(svd.h), u_host.ctypes.data))
                ctx.check(ctx.lib.rc_matrix_to_host(ctx.h, ctx.lib.rc_svd_get_vt(svd.h), vt_host.ctypes.data))
                svd.s_f64()
                op.free()

        e2e_steps = max(4, min(args.steps, 20))
        ms_e2e = timed(lambda: run_pipelined(e2e_steps), 1, 1) / e2e_steps
        e2e = {"value": world * flops_rank / (ms_e2e * 1e-3) / 1e9, "unit": "GFLOP/s", "host_numa_cpus": numa_cpus,
               "h2d_bytes_per_step": m * n * 8, "d2h_bytes_per_step": (m * k + k + k * n) * 8,
               "ms_per_step": ms_e2e, "steps": e2e_steps, "pipeline_depth": 2,
               "api": "rc_matrix_from_host_async / rc_matrix_await -> rc_sample_range_power_iteration -> "
                      "rc_svd_compute_from_range_estimate -> rc_matrix_to_host (pinned host buffers); the upload of the next "
                      "step's operator overlaps the kernels of the current step, every step uploads and downloads its own data",
               "one_step_at_a_time": {"ms_per_step": ms_e2e_serial, "value": world * flops_rank / (ms_e2e_serial * 1e-3) / 1e9,
                                      "api": "rc_matrix_from_host (blocking) -> ... -> rc_matrix_to_host"}}

    if all_cpus:
        try:
            os.sched_setaffinity(0, all_cpus)              # the CPU baselines use every host core
        except Exception:
            pass
    # ---- sub-records: the ID half of the metric and the sharded configs (see the module docstring)
    sub = {}
    if not args.skip_sub:
        try:
            a_dev.free()
        except Exception:
            pass
        peaks_all = roofline["own_measured_peaks"] if roofline else {}

        def guarded(name, fn):
            _mark(f"sub-record {name}")
            try:
                ctx.set_option("release_workspaces", 1)      # every sub-record starts from an empty workspace cache
                r = fn()
                if rank == 0 and r is not None:
                    sub[name] = r
            except Exception as e:            # a sub-record never takes the headline line down with it
                if rank == 0:
                    sub[name] = {"error": f"{type(e).__name__}: {e}"[:300]}

        if world == 1:
            guarded("id_config3", lambda: bench_config3(api, ctx, timed, peaks_all, hbm_peak, not args.skip_cpu))
        guarded("id_config5", lambda: bench_config5(api, ctx, timed, peaks_all, world, rank, not args.skip_cpu))
        if world > 1:
            guarded("config4_strong", lambda: bench_config4(api, ctx, timed, peaks_all, hbm_peak, world, rank))
            guarded("sharded_parity", lambda: sharded_parity(api, ctx, world, rank, local))
    _mark("sub-records done; cpu_baseline leg")
    if rank == 0:
        cpu = None
        if not args.skip_cpu and world == 1:       # the CPU baseline is an N = 1 figure
            m_sample = min(CPU_SAMPLE_ROWS, CFG["m"])
            gf, sec = cpu_sample(m_sample, "gemm", reps=3)
            gv, _ = cpu_sample(2048, "gemv")
            cpu = {"value": gf, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"oracle pipeline on {m_sample} of {m} rows, one GEMM per product (best-case CPU), "
                             f"best of 3 passes, {sec:.1f} s per pass",
                   "reference_faithful_gemv_route_gflops": gv}
        line = {"metric": "rsvd_f64_algorithmic_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu,
                "hbm_gbs_algorithmic": hbm_gbs, "algorithmic_flops_per_step_per_gpu": flops_rank}
        line.update(sub)
        if world > 1:
            line["reference_arm_note"] = ("bench.py --impl reference times ONE 65536-row shard on the host cores at every N (the CPU has "
                                          "no second socket to scale to): the driver's ratio at N > 1 divides an N-GPU aggregate rate by a "
                                          "one-shard CPU rate")
        result_out.write(json.dumps(line) + "\n")
        result_out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
