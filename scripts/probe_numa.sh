#!/bin/bash
# host topology probe for the multi-GPU end-to-end leg (pinned-memory placement): prints what the container may use
echo "== nproc: $(nproc)"; grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
lscpu | grep -E "^CPU\(s\)|Model name|Socket|NUMA|Thread" 
for n in /sys/devices/system/node/node*; do echo "$n cpus=$(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
nvidia-smi topo -m 2>/dev/null | head -20
for d in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader | tr 'A-Z' 'a-z' | sed 's/^0000//'); do echo "gpu $d numa_node=$(cat /sys/bus/pci/devices/$d/numa_node 2>/dev/null)"; done
python - <<'PY'
import ctypes, os
libc = ctypes.CDLL(None, use_errno=True)
import mmap
m = mmap.mmap(-1, 1 << 22)
addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
for node in (0, 1):
    mask = ctypes.c_ulong(1 << node)
    r = libc.syscall(237, ctypes.c_void_p(addr), ctypes.c_ulong(1 << 22), 2, ctypes.byref(mask), ctypes.c_ulong(64), 0)  # mbind MPOL_BIND
    print("mbind node", node, "->", r, os.strerror(ctypes.get_errno()) if r else "ok")
PY
