"""NUMA placement of a rank's host staging buffers (one process per GPU).

The host-buffer API (`DamPostprocessPlan`, `EncodeTargetsPlan`) moves hundreds of MB per step between pinned host
memory and the GPU.  On a two-socket box a rank whose pinned pages live on the other socket pays the inter-socket
link on every copy, and eight ranks that all allocate on node 0 share one memory controller.  `bind_to_gpu(i)`,
called BEFORE the pinned buffers are allocated, (1) restricts the process to the CPUs of the GPU's NUMA node when
the cpuset allows it and (2) sets the memory policy to prefer that node (set_mempolicy(MPOL_PREFERRED), so that the
pages cudaHostAlloc pins are local to the GPU's PCIe root).  Pure host plumbing: no effect on results.
"""
import ctypes
import os

_MPOL_PREFERRED = 1
_SYS_SET_MEMPOLICY = 238  # x86_64


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(index):
    """NUMA node of CUDA device `index` from sysfs (None when it cannot be determined)"""
    import torch
    try:
        p = torch.cuda.get_device_properties(index)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu(index):
    """Returns a dict describing what was done: {'node', 'cpus_bound', 'mempolicy_rc', 'nodes_online'}."""
    info = {"node": gpu_numa_node(index), "cpus_bound": None, "mempolicy_rc": None, "nodes_online": None}
    try:
        with open("/sys/devices/system/node/online") as f:
            info["nodes_online"] = f.read().strip()
    except Exception:
        pass
    node = info["node"]
    if node is None:
        return info
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        want = cpus & allowed
        if want and want != allowed:
            os.sched_setaffinity(0, want)
        info["cpus_bound"] = len(want) if want else 0
    except Exception:
        pass
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(_SYS_SET_MEMPOLICY, _MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))
        info["mempolicy_rc"] = int(rc) if rc == 0 else -ctypes.get_errno()
    except Exception:
        pass
    return info
