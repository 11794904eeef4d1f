#!/bin/bash
echo "--- nodes"; ls /sys/devices/system/node/ 2>/dev/null | tr '\n' ' '; echo
for n in /sys/devices/system/node/node*; do echo "$n cpus $(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
echo "--- self"; grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
echo "--- gpus"; nvidia-smi --query-gpu=index,pci.bus_id --format=csv,noheader
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class 2>/dev/null | cut -c1-4)" = "0x03" ]; then echo "$d numa_node $(cat $d/numa_node) local_cpulist $(cat $d/local_cpulist)"; fi; done
nvidia-smi topo -m 2>&1 | head -20
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"
python - <<'PY'
import ctypes, os
libc = ctypes.CDLL("libc.so.6", use_errno=True)
# set_mempolicy(MPOL_PREFERRED=1, nodemask, maxnode)
for node in (0, 1):
    mask = ctypes.c_ulong(1 << node)
    r = libc.syscall(238, 1, ctypes.byref(mask), 64)
    print("set_mempolicy preferred node", node, "->", r, os.strerror(ctypes.get_errno()) if r else "ok")
libc.syscall(238, 0, None, 0)
PY
