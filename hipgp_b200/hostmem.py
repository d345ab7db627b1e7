"""Host-side placement for the host-buffer entry points (`hipgp_pcg_host*`, `hipgp_matvec_host`).

With one process per GPU on a multi-socket node, pinned buffers should live on the NUMA node the GPU hangs off: the H2D / D2H
copies of eight ranks otherwise cross the socket interconnect and share one node's memory controllers.  Linux places new pages
on the node of the allocating thread, so binding the process to the GPU's node BEFORE the pinned allocations is enough.
"""
import os


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(s):
    cpus = set()
    for part in s.split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index):
    """NUMA node of a CUDA device from sysfs, or None when the platform does not say (single node, virtualised PCI)."""
    import torch
    p = torch.cuda.get_device_properties(device_index)
    try:
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except AttributeError:
        return None
    node = _read("/sys/bus/pci/devices/%s/numa_node" % bdf)
    if node is None or int(node) < 0:
        return None
    return int(node)


def bind_to_gpu_numa_node(device_index):
    """Restrict this process to the CPUs of the GPU's NUMA node (so that later pinned allocations are node-local).
    Returns a small dict describing what was done; never raises -- placement is an optimisation."""
    info = {"node": None, "bound": False}
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["nodes"] = len(nodes)
        node = gpu_numa_node(device_index)
        info["node"] = node
        if node is None or len(nodes) < 2:
            return info
        cl = _read("/sys/devices/system/node/node%d/cpulist" % node)
        cpus = _parse_cpulist(cl) & os.sched_getaffinity(0) if cl else set()
        if not cpus:
            return info
        os.sched_setaffinity(0, cpus)
        info["bound"] = True; info["cpus"] = len(cpus)
    except Exception as e:      # pragma: no cover
        info["error"] = repr(e)[:120]
    return info
