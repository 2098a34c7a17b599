"""Host-side placement for the host-buffer batch call: bind the calling process to the CPU cores (and so, through
Linux's local allocation policy, the NUMA node) nearest to its GPU *before* it allocates pinned memory.

On a multi-socket box every rank's pinned frames / masks otherwise land on whichever node the launcher happened to run
the process on, and half the GPUs copy across the socket interconnect (round 1: 8 ranks moved 133 GB/s in aggregate
against 76 GB/s for one).  Nothing here is needed by the device-resident path.

Sources, in order: NVML's ideal CPU set of the device; sysfs (`/sys/bus/pci/devices/<bdf>/numa_node` +
`/sys/devices/system/node/node<N>/cpulist`).  Failure is reported, never fatal."""
from __future__ import annotations

import os
from typing import Dict, List, Optional


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(','):
        if not part:
            continue
        if '-' in part:
            a, b = part.split('-')
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def _nvml_cpus(device: int) -> Optional[List[int]]:
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(device)
            n_words = (os.cpu_count() + 63) // 64
            words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
            cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1]
            return cpus or None
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        return None


def _sysfs_node(device: int) -> Optional[int]:
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def _sysfs_cpus(node: int) -> Optional[List[int]]:
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            return _parse_cpulist(f.read()) or None
    except Exception:
        return None


def visible_device_index(local_rank: int) -> int:
    """Physical index of the `local_rank`-th visible device (CUDA_VISIBLE_DEVICES with plain indices)."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        items = [v.strip() for v in vis.split(',') if v.strip()]
        if local_rank < len(items) and items[local_rank].isdigit():
            return int(items[local_rank])
    return local_rank


def bind_to_gpu(local_rank: int) -> Dict[str, object]:
    """Restrict this process to the cores nearest to its GPU (no-op when they cannot be determined or the set would
    be empty after intersecting with the cores the process may use).  Returns what was done, for the bench line."""
    info: Dict[str, object] = {"bound": False, "source": None, "node": None, "n_cpus": None}
    try:
        allowed = set(os.sched_getaffinity(0))
    except Exception:
        return info
    phys = visible_device_index(local_rank)
    node = _sysfs_node(local_rank)
    info["node"] = node
    cpus = _nvml_cpus(phys)
    src = "nvml"
    if not cpus and node is not None:
        cpus = _sysfs_cpus(node)
        src = "sysfs"
    if not cpus:
        return info
    want = sorted(allowed & set(cpus))
    if not want or len(want) == len(allowed):
        info["source"] = src
        info["n_cpus"] = len(want)
        return info                       # single-node box (or nothing to narrow): leave the mask alone
    try:
        os.sched_setaffinity(0, want)
        info.update(bound=True, source=src, n_cpus=len(want))
    except Exception:
        pass
    return info


def topology_report() -> str:
    """One text block for profiles/: nodes, their cores, and each visible GPU's node (best effort)."""
    lines = []
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except Exception:
        nodes = []
    for n in nodes:
        cp = _sysfs_cpus(n)
        lines.append(f"node{n}: {len(cp) if cp else '?'} cpus")
    try:
        import torch
        for i in range(torch.cuda.device_count()):
            p = torch.cuda.get_device_properties(i)
            lines.append("gpu%d: %04x:%02x:%02x.0 node %s, nvml cpus %s" % (
                i, p.pci_domain_id, p.pci_bus_id, p.pci_device_id, _sysfs_node(i),
                len(_nvml_cpus(visible_device_index(i)) or []) or '?'))
    except Exception as e:            # pragma: no cover
        lines.append(f"gpus: {e}")
    lines.append(f"allowed cpus: {len(os.sched_getaffinity(0))} of {os.cpu_count()}")
    return "\n".join(lines)
