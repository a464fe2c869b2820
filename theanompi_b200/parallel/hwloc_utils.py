"""NUMA / CPU affinity helpers (ref ``theanompi/lib/hwloc_utils.py``).

The reference binds a worker (and, through the exported ``CPULIST_<label>`` variable, its
spawned loader) to the CPU socket closest to its GPU with python-hwloc
(``hwloc_utils.py:40-58``); ``tmlauncher`` derives the core list per GPU from
``nvidia-smi topo -m`` (``bin/tmlauncher:203-248``).  Here the same is done with
``os.sched_setaffinity`` (+ ``numactl``-free memory locality: first-touch on the bound
cores) and ``nvidia-smi topo -m`` parsing in Python.  ``range_to_list`` fixes the
reference's ``ids.extend(int(...))`` bug.
"""
from __future__ import annotations

import os
import re
import subprocess


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


def bind_to_socket_mem(cpulist, label="train"):
    """Pin this process (and future children/threads) to ``cpulist``; export
    ``CPULIST_<label>`` so the loader inherits it."""
    ids = range_to_list(cpulist)
    avail = os.sched_getaffinity(0)
    use = sorted(set(ids) & set(avail)) or sorted(avail)
    os.sched_setaffinity(0, use)
    os.environ["CPULIST_%s" % label] = ",".join(str(i) for i in use)
    return use


def detect_socket_num(debug=True, label="train"):
    """Which NUMA node(s) the current affinity mask covers."""
    cpus = sorted(os.sched_getaffinity(0))
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
