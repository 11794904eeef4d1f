#!/usr/bin/env python
"""Per-stage roofline table of one configs[1] rSVD step from an ncu launch list of bench.py
(`ncu --metrics gpu__time_duration.sum --csv`, tools/gpu_profile.sh): SURVEY.md 8(d) asks for time, algorithmic
GFLOP/s, algorithmic GB/s and BOTH roofline fractions per timed stage, with the binding roof named.

A stage is a big product with A (a launch longer than 1 ms, plus its split-K reduction) or the run of small
launches between two of them.  Times are ncu's serialised, cold-cache kernel durations: use them for shares and
for which roof binds, not as bench values.  usage: tools/stage_roofline.py launches.csv [fp64_peak_tflops hbm_peak_gbs]"""
import csv
import re
import sys

M, N, K, P, IT = 65536, 8192, 64, 10, 2
L = K + P
FP64 = float(sys.argv[2]) if len(sys.argv) > 2 else 37.0      # own-measured DMMA peak, TFLOP/s (profiles/r1_fp64_peaks.json)
HBM = float(sys.argv[3]) if len(sys.argv) > 3 else 6532.5     # MEASURED_PEAKS.json, GB/s


def qr_flops(rows, w):          # factor + form Q
    return 2.0 * (2.0 * rows * w * w - (2.0 / 3.0) * w ** 3)


# algorithmic (flops, bytes) per stage kind
WORK = {
    "Y = A Omega / A w   (m x n x l)": (2.0 * M * N * L, 8.0 * (M * N + (M + N) * L)),
    "Z = A^H q           (n x m x l)": (2.0 * M * N * L, 8.0 * (M * N + (M + N) * L)),
    "B = Q^H A           (k x m x n)": (2.0 * M * N * K, 8.0 * (M * N + (M + N) * K)),
    "pivoted QR of m x l sketch": (qr_flops(M, L), 4.0 * M * L * 8),
    "pivoted QR of n x l sketch": (qr_flops(N, L), 4.0 * N * L * 8),
    "SVD of b (k x n) + U = Q u_b": (qr_flops(N, K) + 12.0 * K ** 3 + 2.0 * M * K * K, 8.0 * (4.0 * N * K + 2.0 * M * K)),
}


def main():
    lines = open(sys.argv[1]).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = []
    for r in csv.DictReader(lines[start:]):
        if r["Metric Name"] == "gpu__time_duration.sum":
            name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void <unnamed>::", "").replace("<unnamed>::", "").replace("void ", "")
            rows.append((name, float(r["Metric Value"].replace(",", "")) / 1e3))      # us
    jac = [i for i, (n, _) in enumerate(rows) if n.startswith("jacobi")]
    seg = rows[jac[0] + 1:jac[1] + 1]
    # the step starts with the tail of the previous SVD stage (U = Q u_b): rotate it to the end
    first_gauss = [i for i, (n, _) in enumerate(seg) if n.startswith("gaussian")][0]
    seg = seg[first_gauss:] + seg[:first_gauss]
    stages, cur = [], []
    i = 0
    while i < len(seg):
        n, t = seg[i]
        if t > 1000.0:
            if cur:
                stages.append(("small", cur)); cur = []
            big = [(n, t)]
            if i + 1 < len(seg) and "reduce" in seg[i + 1][0]:
                big.append(seg[i + 1]); i += 1
            stages.append(("big", big))
        else:
            cur.append((n, t))
        i += 1
    if cur:
        stages.append(("small", cur))
    # label the stages by their position in the reference's call order (src/random_sampling.rs:138-159, src/svd.rs:175-182)
    labels = []
    big_seen = 0
    for kind, items in stages:
        if kind == "big":
            name = items[0][0]
            if name.startswith("dmma_gemm_kernel<64, 1"):
                labels.append("B = Q^H A           (k x m x n)")
            elif ", 1," in name:
                labels.append("Z = A^H q           (n x m x l)")
            else:
                labels.append("Y = A Omega / A w   (m x n x l)")
            big_seen += 1
        else:
            tot = sum(t for _, t in items)
            has_jac = any(n.startswith("jacobi") for n, _ in items)
            if has_jac:
                labels.append("SVD of b (k x n) + U = Q u_b")
            elif tot < 100.0:
                labels.append("Omega (Philox)")
            else:
                # after an A-product with m rows the sketch is m x l, after Z = A^H q it is n x l
                prev = labels[-1] if labels else ""
                labels.append("pivoted QR of n x l sketch" if prev.startswith("Z") else "pivoted QR of m x l sketch")
    total = sum(t for _, items in stages for _, t in items)
    print(f"one rSVD step of configs[1] ({M} x {N} f64, l = {L}, it = {IT}): {sum(len(it) for _, it in stages)} launches, "
          f"{total / 1e3:.2f} ms of kernel time (ncu, serialised)")
    print(f"peaks: FP64 {FP64:.1f} TFLOP/s (own-measured DMMA), HBM {HBM:.0f} GB/s (MEASURED_PEAKS.json)")
    print(f"{'stage':34s} {'launches':>8s} {'us':>8s} {'share':>6s} {'TFLOP/s':>8s} {'of FP64':>8s} {'GB/s':>7s} {'of HBM':>7s}  bound")
    agg = {}
    for lab, (kind, items) in zip(labels, stages):
        t = sum(x for _, x in items)
        fl, by = WORK.get(lab, (0.0, 0.0))
        tf = fl / (t * 1e-6) / 1e12 if fl else 0.0
        gb = by / (t * 1e-6) / 1e9 if by else 0.0
        ff, hf = tf / FP64, gb / HBM
        bound = "FP64 pipe" if ff > 0.5 else ("HBM" if hf > 0.5 else "latency (dependent one-CTA kernels)")
        print(f"{lab:34s} {len(items):8d} {t:8.0f} {100 * t / total:5.1f}% {tf:8.2f} {ff:8.2f} {gb:7.0f} {hf:7.2f}  {bound}")
        a = agg.setdefault(lab, [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += t; a[2] += fl; a[3] += by
    print("\naggregated by stage kind")
    for lab, (cnt, t, fl, by) in agg.items():
        tf = fl / (t * 1e-6) / 1e12 if fl else 0.0
        gb = by / (t * 1e-6) / 1e9 if by else 0.0
        print(f"{lab:34s} x{cnt:<3d} {t / 1e3:7.2f} ms {100 * t / total:5.1f}%  {tf:6.2f} TFLOP/s ({tf / FP64:4.2f})  {gb:6.0f} GB/s ({gb / HBM:4.2f})")


if __name__ == "__main__":
    main()
