"""Job supervisor: start one worker process per device, fail fast, clean teardown.

Replaces ``mpirun`` (ref ``rules.py:85-116``, ``bin/tmlauncher:364-385``): the reference
relied on MPI's job control — any worker dying aborts the job.  The agent keeps that
contract: it starts every worker in its own process group with the torch.distributed
rendezvous environment (``RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT``),
waits, and on the first non-zero exit (or SIGTERM/SIGINT) kills all remaining worker
groups (loader children included).  Remote hosts are reached with ``ssh`` like the
reference's ``-host`` MPMD entries (``TMPI_SSH`` overrides the remote-shell command, default
``ssh -o BatchMode=yes``; the multi-host test points it at a local shim).
"""
from __future__ import annotations

import json
import os
import shlex
import signal
import socket
import subprocess
import sys
import time


def _is_local(host):
    return host in (None, "", "localhost", "127.0.0.1", socket.gethostname(), socket.gethostname().split(".")[0])


def main():
    spec = json.loads(sys.argv[1])
    procs = []
    local_index = {}

    def kill_all(sig=signal.SIGTERM):
        for p in procs:
            if p.poll() is None:
                try:
                    os.killpg(p.pid, sig)
                except (ProcessLookupError, PermissionError):
                    pass

    def on_signal(signum, frame):
        kill_all()
        time.sleep(1.0)
        kill_all(signal.SIGKILL)
        sys.exit(3)

    signal.signal(signal.SIGTERM, on_signal)
    signal.signal(signal.SIGINT, on_signal)

    for w in spec["workers"]:
        host = w["host"]
        lr = local_index.get(host, 0)
        local_index[host] = lr + 1
        env = dict(os.environ)
        env.update(spec.get("env", {}))
        env.update(RANK=str(w["rank"]), WORLD_SIZE=str(spec["world"]), LOCAL_RANK=str(lr),
                   MASTER_ADDR=spec["master_addr"], MASTER_PORT=str(spec["master_port"]))
        cmd = [spec["python"], "-u", "-m", w["module"]] + [str(a) for a in w["argv"]]
        if _is_local(host):
            p = subprocess.Popen(cmd, env=env, cwd=spec.get("cwd"), start_new_session=True)
        else:
            exports = " ".join("%s=%s" % (k, shlex.quote(env[k])) for k in
                               ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "PYTHONPATH")
                               if k in env)
            extra = " ".join("%s=%s" % (k, shlex.quote(v)) for k, v in spec.get("env", {}).items())
            remote = "cd %s && env %s %s %s" % (shlex.quote(spec.get("cwd", ".")), exports, extra,
                                                " ".join(shlex.quote(c) for c in cmd))
            rsh = shlex.split(os.environ.get("TMPI_SSH", "ssh -o BatchMode=yes"))
            p = subprocess.Popen(rsh + [host, remote], start_new_session=True)
        procs.append(p)

    rc = 0
    alive = set(range(len(procs)))
    while alive:
        for i in list(alive):
            r = procs[i].poll()
            if r is None:
                continue
            alive.discard(i)
            if r != 0 and rc == 0:
                rc = r
                sys.stderr.write("[launch_agent] worker rank %d exited with %d: tearing the job down\n" % (i, r))
                kill_all()
                deadline = time.time() + 5
                while time.time() < deadline and any(p.poll() is None for p in procs):
                    time.sleep(0.1)
                kill_all(signal.SIGKILL)
        time.sleep(0.05)
    sys.exit(rc if rc >= 0 else 1)


if __name__ == "__main__":
    main()
