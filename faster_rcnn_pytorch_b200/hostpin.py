"""Host-side placement for callers that feed the region stage from host memory (``region.HostProposalPipeline``):
pin the calling thread -- and with it the pinned staging buffers it allocates afterwards (first touch) -- to the NUMA
node the GPU hangs off.  With one process per GPU on a two-socket host this keeps every rank's 33 MB/step of H2D
traffic off the inter-socket link; it changes no result.  Linux sysfs only; silently a no-op elsewhere."""
from __future__ import annotations

import os


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(device_index: int):
    """NUMA node of the GPU (``/sys/bus/pci/devices/<bus id>/numa_node``) or None."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as fh:
            node = int(fh.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_host_thread_to_gpu_node(device_index: int) -> dict:
    """Restrict the calling thread to the CPUs of the GPU's NUMA node.  Returns what was done (for logs)."""
    info = {"numa_node": None, "cpus": None, "bound": False}
    node = gpu_numa_node(device_index)
    if node is None:
        return info
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = _parse_cpulist(fh.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(numa_node=node, cpus=len(allowed), bound=True)
    except Exception:
        pass
    return info
