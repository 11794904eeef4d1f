#!/usr/bin/env python
"""Key metrics + hottest source lines of the launches in an ncu report (captured with --set full
--import-source on), as plain text for profiles/.  usage: tools/ncu_summary.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg"]
idx = {k: hdr.index(k) for k in want if k in hdr}
for li, r in enumerate(rows[2:]):
    print(f"== launch {li}")
    for k, i in idx.items():
        print(f"   {k:<78s} {r[i]} {units[i]}")
    dr, dw = idx.get("dram__bytes_read.sum"), idx.get("dram__bytes_write.sum")
    if dr is not None and dw is not None:
        def tobytes(v, u):
            v = float(v.replace(",", "")); return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        print(f"   {'dram bytes read + written per launch':<78s} {tobytes(r[dr], units[dr]) + tobytes(r[dw], units[dw]):.0f} byte")
    out = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "hot_lines.py"), rep, str(li), str(top)],
                         capture_output=True, text=True).stdout
    print("   hottest source lines (share of stall samples / of executed warp instructions):")
    for line in out.splitlines():
        print("     " + line)
