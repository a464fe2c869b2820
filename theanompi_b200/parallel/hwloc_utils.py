"""NUMA / CPU affinity helpers (ref ``theanompi/lib/hwloc_utils.py``).

The reference binds a worker (and, through the exported ``CPULIST_<label>`` variable, its
spawned loader) to the CPU socket closest to its GPU with python-hwloc
(``hwloc_utils.py:40-58``); ``tmlauncher`` derives the core list per GPU from
``nvidia-smi topo -m`` (``bin/tmlauncher:203-248``).  Here the same is done with
``os.sched_setaffinity`` for the cores, the ``set_mempolicy(MPOL_BIND)`` system call for the memory
(the reference's ``topology.set_membind(..., hwloc.MEMBIND_BIND, ...)``, ``hwloc_utils.py:52-56`` — no libnuma /
python-hwloc needed) and ``nvidia-smi topo -m`` parsing in Python.  ``range_to_list`` fixes the
reference's ``ids.extend(int(...))`` bug.
"""
from __future__ import annotations

import ctypes
import os
import platform
import re
import subprocess

MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 0, 1, 2
_SYS_SET_MEMPOLICY = {"x86_64": 238, "aarch64": 237, "ppc64le": 261}


def range_to_list(cpulist):
    """'0-3,8,10-11' → [0,1,2,3,8,10,11]"""
    ids = []
    for part in str(cpulist).split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-", 1)
            ids.extend(range(int(a), int(b) + 1))
        else:
            ids.append(int(part))
    return ids


def range_expand(s):
    return ",".join(str(i) for i in range_to_list(s))


def nodes_of_cpus(cpus):
    """NUMA nodes (from sysfs) that contain any of ``cpus``."""
    nodes = set()
    base = "/sys/devices/system/node"
    if os.path.isdir(base):
        for n in os.listdir(base):
            m = re.match(r"node(\d+)$", n)
            if not m:
                continue
            try:
                with open(os.path.join(base, n, "cpulist")) as f:
                    node_cpus = set(range_to_list(f.read().strip()))
            except OSError:
                continue
            if node_cpus & set(cpus):
                nodes.add(int(m.group(1)))
    return sorted(nodes)


def set_membind(nodes, mode=MPOL_BIND):
    """Bind future page allocations of this process (inherited by children) to the NUMA ``nodes``.  Returns True when the
    kernel accepted the policy; False (and no change) on kernels / containers that refuse it or on unknown architectures."""
    nr = _SYS_SET_MEMPOLICY.get(platform.machine())
    if nr is None:
        return False
    if mode == MPOL_DEFAULT:                                 # back to the system default policy (mask must be empty)
        try:
            libc = ctypes.CDLL(None, use_errno=True)
            return libc.syscall(ctypes.c_long(nr), ctypes.c_int(MPOL_DEFAULT), None, ctypes.c_ulong(0)) == 0
        except (OSError, AttributeError):
            return False
    if not nodes:
        return False
    maxnode = max(nodes) + 2
    nwords = (maxnode + 63) // 64
    mask = (ctypes.c_ulong * nwords)()
    for n in nodes:
        mask[n // 64] |= 1 << (n % 64)
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        rc = libc.syscall(ctypes.c_long(nr), ctypes.c_int(mode), ctypes.byref(mask), ctypes.c_ulong(maxnode))
    except (OSError, AttributeError):
        return False
    return rc == 0


def bind_to_socket_mem(cpulist, label="train", membind=True):
    """Pin this process (and future children/threads) to ``cpulist`` and bind its memory to the NUMA node(s) of those
    cores; export ``CPULIST_<label>`` so the loader inherits it."""
    ids = range_to_list(cpulist)
    avail = os.sched_getaffinity(0)
    use = sorted(set(ids) & set(avail)) or sorted(avail)
    os.sched_setaffinity(0, use)
    os.environ["CPULIST_%s" % label] = ",".join(str(i) for i in use)
    if membind:
        all_nodes = nodes_of_cpus(range(0, 4096))
        nodes = nodes_of_cpus(use)
        if nodes and len(nodes) < len(all_nodes):            # binding to every node is the default policy anyway
            set_membind(nodes)
    return use


def detect_socket_num(debug=True, label="train"):
    """Which NUMA node(s) the current affinity mask covers."""
    cpus = sorted(os.sched_getaffinity(0))
    nodes = nodes_of_cpus(cpus)
    if debug:
        print("[%s] pid %d bound to %d cpus on NUMA node(s) %s" % (label, os.getpid(), len(cpus), sorted(nodes) or "?"))
    return sorted(nodes)


def gpu_cpu_affinity(host=None):
    """Parse ``nvidia-smi topo -m`` → {gpu_index: 'cpu list string'}."""
    cmd = ["nvidia-smi", "topo", "-m"]
    if host:
        cmd = ["ssh", host] + cmd
    try:
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=20).stdout
    except Exception:
        return {}
    return parse_topo(out)


def parse_topo(text):
    aff = {}
    lines = [re.sub(r"\x1b\[[0-9;]*m", "", l) for l in text.splitlines()]
    header = None
    for l in lines:
        if "CPU Affinity" in l:
            header = re.split(r"\t+|\s{2,}", l.strip())
            continue
        m = re.match(r"^GPU(\d+)\s", l)
        if m and header:
            cols = re.split(r"\t+|\s{2,}", l.strip())
            try:
                idx = header.index("CPU Affinity") + 1       # row label shifts columns by one
                aff[int(m.group(1))] = cols[idx]
            except (ValueError, IndexError):
                pass
    return aff
