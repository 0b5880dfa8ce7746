"""NUMA placement of a rank's host staging memory (multi-GPU end-to-end path).  With one process per GPU on one
host, pinned buffers land on whatever node the process happens to run on; eight ranks pulling 4.6 KB per patch
through one memory controller is what bounded the 8-GPU end-to-end rate.  ``bind_to_gpu_node`` pins the calling
process to the CPUs of the NUMA node its GPU hangs off BEFORE the staging buffers are allocated (first touch /
default local policy then places them there).  Best effort: returns ``None`` when the topology is not exposed."""
from __future__ import annotations

import os


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device):
    """NUMA node of CUDA device ``device`` from sysfs, or None."""
    import torch

    try:
        p = torch.cuda.get_device_properties(device)
        dom, bus, dev = (getattr(p, k, None) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id"))
        if bus is None:
            return None
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom or 0, bus, dev or 0)
        with open(path) as fh:
            node = int(fh.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_node(device):
    """Restrict this process to the CPUs of the GPU's NUMA node.  Returns ``{"node", "cpus"}`` or None."""
    node = gpu_numa_node(device)
    if node is None:
        return None
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            cpus = _cpulist(fh.read()) & set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception:
        return None
