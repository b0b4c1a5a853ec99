"""Host-side placement for the one-process-per-GPU launch: run each rank on the CPUs of its
GPU's NUMA node, so that the pinned staging buffers it then allocates (first touch) sit on the
memory controller next to that GPU's PCIe root port.  End to end the path is PCIe-bound, and a
staging buffer on the far socket halves the copy rate once several GPUs copy at the same time.
"""
from __future__ import annotations

import os
import subprocess


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(index: int) -> int | None:
    """NUMA node of CUDA device ``index`` (nvidia-smi order), or None when the platform hides it."""
    try:
        out = subprocess.run(["nvidia-smi", f"--id={index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip()
    except (OSError, subprocess.TimeoutExpired):
        return None
    if not out:
        return None
    bus = out.lower()
    if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:        # 00000000:1b:00.0 -> 0000:1b:00.0
        bus = bus[4:]
    try:
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def bind_to_gpu(index: int) -> dict:
    """Restrict this process to the allowed CPUs of the GPU's NUMA node.  Returns what was done."""
    info = {"gpu": index, "numa_node": None, "bound": False}
    node = gpu_numa_node(index)
    info["numa_node"] = node
    if node is None:
        return info
    try:
        local = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = os.sched_getaffinity(0)
        use = local & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["bound"] = True
        info["cpus"] = len(use or allowed)
    except OSError:
        pass
    return info
