"""Host-side placement for the feature ingest: run this process (and the pinned buffers it allocates afterwards, by
first touch) on the CPU cores of the NUMA node the GPU hangs off.  With several ranks per node each rank then packs and
DMA-reads its own node's memory instead of crossing the socket interconnect.  Pure host plumbing (sysfs + sched_setaffinity);
everything is best effort and reports what it did."""
from __future__ import annotations

import os
from typing import Dict, List, Optional


def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' -> [0,1,2,3,8,10,11] (the format of /sys/devices/system/node/node*/cpulist)."""
    cpus: List[int] = []
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-", 1)
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(pci_bus_id: str) -> Optional[int]:
    """NUMA node of a PCI device ('0000:1b:00.0', any case) from sysfs; None when unknown (-1 / no sysfs)."""
    bid = pci_bus_id.lower()
    if bid.count(":") == 2 and len(bid.split(":")[0]) == 8:      # nvml style 00000000:1B:00.0
        bid = bid[4:]
    try:
        with open(f"/sys/bus/pci/devices/{bid}/numa_node") as fh:
            n = int(fh.read().strip())
        return n if n >= 0 else None
    except (OSError, ValueError):
        return None


def bind_to_gpu_numa(device_index: int, local_world: int = 1, local_rank: int = 0) -> Dict[str, object]:
    """Restrict this process to the cores of the GPU's NUMA node (or, when the node is unknown, to an equal slice of the
    cores it may run on).  Returns {'numa_node', 'cpus', 'bound'}."""
    info: Dict[str, object] = {"numa_node": None, "cpus": None, "bound": False}
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        node = gpu_numa_node(f"{dom:04x}:{bus:02x}:{dev:02x}.0")
    except Exception:  # noqa: BLE001 -- placement is an optimisation, never an error
        node = None
    allowed = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))
    cpus = None
    if node is not None:
        try:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
                cpus = [c for c in parse_cpulist(fh.read()) if c in allowed]
        except OSError:
            cpus = None
        # several ranks on one node share its cores evenly
        if cpus and local_world > 1:
            sharers = max(1, local_world // max(1, _numa_nodes()))
            k = local_rank % sharers
            per = max(1, len(cpus) // sharers)
            cpus = cpus[k * per:(k + 1) * per] or cpus
    if not cpus and local_world > 1:
        per = max(1, len(allowed) // local_world)
        cpus = allowed[local_rank * per:(local_rank + 1) * per] or allowed
    info["numa_node"] = node
    if cpus and hasattr(os, "sched_setaffinity"):
        try:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
            info["cpus"] = len(cpus)
        except OSError:
            pass
    return info


def _numa_nodes() -> int:
    try:
        return max(1, sum(1 for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()))
    except OSError:
        return 1
