"""Rule classes — the user API (ref ``theanompi/rules.py``)::

    from theanompi_b200 import BSP
    rule = BSP()
    rule.init(devices=['cuda0', 'cuda1'], modelfile='theanompi_b200.models.alex_net', modelclass='AlexNet')
    rule.wait()

Reference: each rule shells out to ``mpirun`` MPMD with one ``python worker.py …``
program per device and keeps the ``mpirun`` pid for ``wait()`` (``rules.py:12-59,
79-120``).  Here the single child is :mod:`theanompi_b200.launch_agent`, a tiny
supervisor that starts one worker process per device (``ssh`` for remote hosts) with the
``torch.distributed`` rendezvous environment, fails fast (any worker dying tears the
job down, children of the workers included) and returns the job's exit code — so
``Rule.pid`` / ``Rule.wait()`` keep their meaning, including Ctrl-C → SIGTERM → exit 3.
"""
from __future__ import absolute_import

import json
import os
import signal
import socket
import subprocess
import sys

START_INFO = ("theanompi_b200 started %d workers for \n 1.updating %s params through iterations and\n "
              "2.exchange the params with %s\nSee output log.")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def split_device(device):
    """'node1:cuda3' → ('node1', 'cuda3'); 'cuda3' → (None, 'cuda3')."""
    if ":" in device and not device.startswith("cuda:"):
        h, d = device.split(":", 1)
        return h, d
    return None, device


class Rule(object):
    """Base launcher of the synchronisation rules."""

    def __init__(self):
        self.pid = None
        self.proc = None
        self.rulename = "None"
        self.model_config = {}          # forwarded to the model's config dict (TMPI_MODEL_CONFIG)
        self.env = {}

    def init(self, devices, modelfile, modelclass):
        raise NotImplementedError

    # ------------------------------------------------------------------ process plumbing
    def _spec(self, programs, cpulists=None):
        """programs: list of (device, module, argv list) — one worker process each."""
        hosts = [split_device(d)[0] for d, _, _ in programs]
        first = hosts[0]
        master = "127.0.0.1" if all(h in (None, first) for h in hosts) and first in (None, "localhost", socket.gethostname()) \
            else (first or socket.gethostname())
        master = os.environ.get("TMPI_MASTER_ADDR", master)         # e.g. a host name that differs on the fabric network
        port = int(os.environ.get("TMPI_MASTER_PORT", _free_port()))
        workers = []
        for rank, (device, module, argv) in enumerate(programs):
            host, dev = split_device(device)
            a = list(argv)
            if cpulists and cpulists[rank]:
                a.append(cpulists[rank])
            workers.append(dict(rank=rank, host=host, device=dev, module=module, argv=a))
        env = dict(self.env)
        if self.model_config:
            env["TMPI_MODEL_CONFIG"] = json.dumps(self.model_config)
        return dict(master_addr=master, master_port=port, world=len(workers), workers=workers, env=env,
                    python=sys.executable, cwd=os.getcwd())

    def _launch(self, spec, n, modelclass):
        cmd = [sys.executable, "-u", "-m", "theanompi_b200.launch_agent", json.dumps(spec)]
        env = dict(os.environ)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
        self.proc = subprocess.Popen(cmd, env=env)
        print(START_INFO % (n, modelclass, self.rulename))
        self.pid = self.proc.pid

    def wait(self):
        """Reap the job (ref ``rules.py:35-59``); returns its exit code."""
        if self.pid is None:
            print("Rule %s not initialized" % self.rulename)
            return None
        try:
            rcode = self.proc.wait()
            print("\n Rule session {0} terminated with return code: {1}.".format(self.pid, rcode))
            return rcode
        except (RuntimeError, KeyboardInterrupt):
            print("Killing worker processes...")
            os.kill(self.pid, signal.SIGTERM)
            self.proc.wait()
            sys.exit(3)


class BSP(Rule):
    """Bulk Synchronous Parallel: workers iterate in lock step and exchange after every
    iteration (ref ``rules.py:63-120``).  Class-level knobs like the reference
    (``rules.py:71-72``); ``exch_strategy`` additionally accepts the B200-native fused
    strategies and defaults to them."""

    sync_type = "cdd"            # 'avg' or 'cdd'
    exch_strategy = "fused"      # fused|fused16|oneshot|twoshot|nvls|p2p32  |  nccl32|nccl16|ar|asa32|asa16|copper|copper16|swap

    def __init__(self):
        Rule.__init__(self)
        self.rulename = "BSP(%s,%s)" % (BSP.sync_type, BSP.exch_strategy)

    def init(self, devices, modelfile, modelclass, cpulists=None):
        sync, strat = BSP.sync_type, BSP.exch_strategy
        hosts = set(split_device(d)[0] for d in devices)
        if len(devices) == 1:
            sync = "avg"
        if len(hosts) > 1 or any(split_device(d)[1].startswith("cpu") for d in devices):
            if strat not in ("ar", "nccl32", "nccl16"):
                strat = "ar" if any(split_device(d)[1].startswith("cpu") for d in devices) else "nccl32"
        programs = [(d, "theanompi_b200.worker", [split_device(d)[1], sync, strat, modelfile, modelclass]) for d in devices]
        self._launch(self._spec(programs, cpulists), len(devices), modelclass)


class EASGD(Rule):
    """Elastic Averaging SGD: device[0] hosts the center/server, the rest are workers that
    exchange with it every τ iterations (ref ``rules.py:123-179``)."""

    def __init__(self):
        Rule.__init__(self)
        self.rulename = "EASGD"

    def init(self, devices, modelfile, modelclass, cpulists=None):
        programs = []
        for i, d in enumerate(devices):
            mod = "theanompi_b200.easgd_server" if i == 0 else "theanompi_b200.easgd_worker"
            programs.append((d, mod, [split_device(d)[1], modelfile, modelclass]))
        self._launch(self._spec(programs, cpulists), len(devices) - 1, modelclass)


class ASGD(Rule):
    """Asynchronous SGD with delta pushes.  An empty stub in the reference
    (``rules.py:182-186``); here it reuses the EASGD runtimes with the delta exchanger."""

    def __init__(self):
        Rule.__init__(self)
        self.rulename = "ASGD"

    def init(self, devices, modelfile, modelclass, cpulists=None):
        self.env["TMPI_EASGD_EXCHANGER"] = "asgd"
        programs = []
        for i, d in enumerate(devices):
            mod = "theanompi_b200.easgd_server" if i == 0 else "theanompi_b200.easgd_worker"
            programs.append((d, mod, [split_device(d)[1], modelfile, modelclass]))
        self._launch(self._spec(programs, cpulists), len(devices) - 1, modelclass)


class GOSGD(Rule):
    """Gossip SGD: every worker trains alone and, with probability p per iteration, pushes
    its weights to a random peer (ref ``rules.py:189-239``)."""

    def __init__(self):
        Rule.__init__(self)
        self.rulename = "GOSGD"

    def init(self, devices, modelfile, modelclass, cpulists=None):
        programs = [(d, "theanompi_b200.gosgd_worker", [split_device(d)[1], modelfile, modelclass]) for d in devices]
        self._launch(self._spec(programs, cpulists), len(devices), modelclass)
