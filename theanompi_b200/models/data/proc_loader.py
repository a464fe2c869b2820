"""Loader *process* (B200-native counterpart of ``theanompi/models/data/proc_load_mpi.py:16-133``).

The reference spawns one child per worker with ``MPI.COMM_SELF.Spawn``; the child loads the ``.hkl`` batch, preprocesses it on
the CPU, uploads it through its OWN CUDA context and hands the device buffer to the trainer over CUDA-IPC + ZeroMQ.  Here the
child keeps what must not share the trainer's interpreter — file IO, decompression and the memcpy of a 25 MB batch, i.e. the
GIL-holding part — and nothing else:

* the child process (``multiprocessing`` *spawn* context, no CUDA context, no torch import needed on its hot path) reads each
  requested file straight into a slot of a POSIX shared-memory ring;
* the trainer process maps the same ring, **page-locks it** (``cudaHostRegister``) and its loader thread DMAs the slot to
  the device on the copy stream, followed by the fused normalise / crop / mirror kernel (``ParaLoader``).

So the hand-off costs no extra copy (the file lands in pinned memory the DMA engine reads), there is no second CUDA context
competing for the GPU, and a slow disk / decoder can never stall the trainer's Python thread.  Protocol = the reference's
(filename in → "copy finished" out), over a ``multiprocessing`` pipe instead of MPI tags 40 / 55.
"""
from __future__ import annotations

import os
import sys
from multiprocessing import shared_memory

import numpy as np


def read_batch_file(filename, out, n_channels=3):
    """Fill ``out`` (uint8 NHWC) from one batch file: ``.npy`` (uint8 NHWC) or the reference's ``.hkl`` (c01b)."""
    if filename.endswith(".hkl"):
        try:
            import hickle
        except ImportError as e:
            raise RuntimeError("%s: reading .hkl batches needs the hickle package; convert them to .npy (uint8 NHWC)" % filename) from e
        arr = np.asarray(hickle.load(filename))
        if arr.ndim == 4 and arr.shape[0] == n_channels:           # reference layout c01b → b01c
            arr = np.transpose(arr, (3, 1, 2, 0))
        np.copyto(out, arr.astype(np.uint8, copy=False))
    else:
        from .utils import parallel_copyto
        arr = np.load(filename, mmap_mode="r")
        parallel_copyto(out, arr)


def _synthetic_fill(filename, out, seed):
    idx = int(filename.rsplit("/", 1)[1])
    rs = np.random.RandomState(seed + idx % 4)
    out[...] = rs.randint(0, 256, out.shape, dtype=np.uint8)


def _child_main(shm_names, raw_shape, seed, fin, fout):
    """Loader child: a line ``<slot> <filename>`` in → fill ring slot → a line ``<slot> OK`` (or ``<slot> ERR <why>``) out."""
    try:
        os.sched_setaffinity(0, os.sched_getaffinity(0))           # inherit the parent's NUMA binding explicitly
    except Exception:  # noqa: BLE001
        pass
    # attach without registering with the resource tracker: the parent owns (and unlinks) the segments
    shms = [shared_memory.SharedMemory(name=n, track=False) if sys.version_info >= (3, 13) else shared_memory.SharedMemory(name=n)
            for n in shm_names]
    if sys.version_info < (3, 13):
        from multiprocessing import resource_tracker
        for s_ in shms:                                             # attaching registered them with OUR tracker: undo
            try:
                resource_tracker.unregister(s_._name, "shared_memory")
            except Exception:  # noqa: BLE001
                pass
    views = [np.ndarray(raw_shape, dtype=np.uint8, buffer=s.buf) for s in shms]
    try:
        for line in fin:
            line = line.rstrip("\n")
            if not line or line == "STOP":
                break
            slot_s, filename = line.split(" ", 1)
            slot = int(slot_s)
            try:
                if filename.startswith("synthetic://"):
                    _synthetic_fill(filename, views[slot], seed)
                else:
                    read_batch_file(filename, views[slot])
                fout.write("%d OK\n" % slot)
            except Exception as e:  # noqa: BLE001
                fout.write("%d ERR %s\n" % (slot, repr(e).replace("\n", " ")))
            fout.flush()
    finally:
        del views
        for s in shms:
            try:
                s.close()
            except Exception:  # noqa: BLE001
                pass


class ProcReader(object):
    """Owns the shared-memory ring and the loader child; ``read(item, slot)`` blocks (GIL released while waiting on the pipe)
    until the child has filled the ring slot.  ``tensors`` are the torch views of the ring (page-locked when CUDA is
    available).  The child is a fresh interpreter (``python -m …proc_loader``), not a fork: no CUDA state is inherited."""

    def __init__(self, raw_shape, depth=2, seed=0, pin=True):
        import subprocess
        import torch
        self.raw_shape = tuple(int(v) for v in raw_shape)
        nbytes = int(np.prod(self.raw_shape))
        self.shms = [shared_memory.SharedMemory(create=True, size=nbytes) for _ in range(depth)]
        self.tensors, self._registered = [], []
        for s in self.shms:
            t = torch.frombuffer(s.buf, dtype=torch.uint8, count=nbytes).view(self.raw_shape)
            if pin and torch.cuda.is_available():
                rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), nbytes, 0)
                if int(rc) != 0:
                    raise RuntimeError("cudaHostRegister of the loader ring failed: %s" % (rc,))
                self._registered.append(t.data_ptr())
            self.tensors.append(t)
        root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
        env = dict(os.environ)
        env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
        env["CUDA_VISIBLE_DEVICES"] = ""                                  # the child never touches the GPU
        self.proc = subprocess.Popen([sys.executable, "-u", "-m", "theanompi_b200.models.data.proc_loader",
                                      ",".join(s.name for s in self.shms), ",".join(str(v) for v in self.raw_shape), str(int(seed))],
                                     stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, bufsize=1, env=env)

    def slot_of(self, np_view):
        ptr = np_view.__array_interface__["data"][0]
        for i, t in enumerate(self.tensors):
            if t.data_ptr() == ptr:
                return i
        raise ValueError("buffer is not a slot of the loader ring")

    def read(self, item, out_np):
        """``read_fn`` of :class:`ParaLoader`: delegate the file read to the child; the batch lands in ``out_np`` (a ring slot)."""
        slot = self.slot_of(out_np)
        if " " in item.split(" ", 1)[0] or "\n" in item:
            raise ValueError("bad batch file name %r" % item)
        self.proc.stdin.write("%d %s\n" % (slot, item))
        self.proc.stdin.flush()
        reply = self.proc.stdout.readline()
        if not reply:
            raise RuntimeError("loader process died (exit code %s)" % self.proc.poll())
        parts = reply.rstrip("\n").split(" ", 2)
        if len(parts) < 2 or parts[1] != "OK":
            raise RuntimeError("loader process failed on %s: %s" % (item, reply.strip()))
        assert int(parts[0]) == slot
        return None

    def close(self):
        import torch
        try:
            if self.proc.poll() is None:
                self.proc.stdin.write("STOP\n"); self.proc.stdin.flush()
                self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            try:
                self.proc.kill()
            except Exception:  # noqa: BLE001
                pass
        for p in self._registered:
            try:
                torch.cuda.cudart().cudaHostUnregister(p)
            except Exception:  # noqa: BLE001
                pass
        self._registered = []
        self.tensors = []
        for s in self.shms:
            try:
                s.close(); s.unlink()
            except Exception:  # noqa: BLE001
                pass
        self.shms = []


if __name__ == "__main__":
    _names = sys.argv[1].split(",")
    _shape = tuple(int(v) for v in sys.argv[2].split(","))
    _child_main(_names, _shape, int(sys.argv[3]), sys.stdin, sys.stdout)
