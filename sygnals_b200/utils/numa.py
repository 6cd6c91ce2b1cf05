"""Host-side placement for the host-buffer entry points (``syg_features_host_*``): bind the calling process to the CPUs of the
NUMA node its GPU hangs off, so that pinned staging buffers allocated afterwards live in that node's memory and the chunked
H2D / D2H copies do not cross the socket interconnect.  With one process per GPU on a two-socket box this is what keeps eight
concurrent 54 GB/s uploads from all reading one socket's DRAM.  Linux sysfs only; every failure is a no-op."""
from __future__ import annotations

import os
from typing import Optional


def _pci_address(device_index: int) -> Optional[str]:
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        dom, bus, dev = (getattr(pr, k, None) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id"))
        if bus is not None and dev is not None:
            return f"{int(dom or 0):04x}:{int(bus):02x}:{int(dev):02x}.0"
    except Exception:
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[device_index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else device_index
        info = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx))
        bus_id = info.busId.decode() if isinstance(info.busId, bytes) else info.busId
        dom, rest = bus_id.split(":", 1)
        return f"{int(dom, 16):04x}:{rest.lower()}"
    except Exception:
        return None


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_device_node(device_index: int) -> Optional[dict]:
    """Returns {'node': n, 'cpus': count, 'pci': address} when the affinity was set, else None."""
    addr = _pci_address(device_index)
    if addr is None or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        with open(f"/sys/bus/pci/devices/{addr}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        cpus &= os.sched_getaffinity(0)                       # stay inside the container's cpuset
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus), "pci": addr}
    except Exception:
        return None
