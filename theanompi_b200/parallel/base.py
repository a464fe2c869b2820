"""Process / communicator base (ref ``theanompi/lib/base.py`` ``MPI_GPU_Process``).

Reference: one MPI process per GPU; ``MPI.COMM_WORLD`` is the control plane
(pickled ``send/recv``, ``bcast``, ``allgather``, ``Barrier``, ``Iprobe``), an NCCL-1
clique built by broadcasting the unique id over MPI is the data plane, pair cliques
emulate point-to-point (``base.py:15-150``).

Here: one process per GPU as well, but

* rendezvous + control plane = ``torch.distributed`` (TCPStore + a gloo group for host
  objects; nccl group only for the *baseline* strategies).  ``Comm`` offers the
  mpi4py-flavoured calls the rule runtimes need — ``bcast / allgather / Barrier`` and a
  tagged mailbox ``send / recv(source=ANY) / iprobe`` built on atomic store counters,
  which is how the EASGD server receives from ``ANY_SOURCE`` and GOSGD probes for pushes.
* data plane = the peer-mapped symmetric arena (:mod:`.symmetric`) the fused kernels
  read/write directly over NVLink; pairwise "communicators" are just peer pointers, so
  the reference's C(N,2) NCCL pair cliques (``base.py:124-150``) have no start-up cost.

Devices: ``cudaN`` (GPU) or ``cpuN`` (gloo-only, used by the CPU test-suite).
"""
from __future__ import annotations

import datetime
import os
import pickle
import socket
import time

import torch
import torch.distributed as dist

ANY_SOURCE = -1


def parse_device(device):
    """'cuda3' → ('cuda', 3); 'cpu0' → ('cpu', 0); also accepts 'host:cuda3'."""
    if ":" in device and not device.startswith("cuda:"):
        device = device.split(":", 1)[1]
    device = device.replace("cuda:", "cuda")
    if device.startswith("cuda"):
        return "cuda", int(device[4:] or 0)
    if device.startswith("gpu"):
        return "cuda", int(device[3:] or 0)
    if device.startswith("cpu"):
        return "cpu", int(device[3:] or 0)
    raise ValueError("device must look like cuda0 / cpu0, got %r" % device)


class Comm(object):
    """mpi4py-like facade over a torch.distributed process group + its store."""

    def __init__(self, group=None, store=None, rank=None, size=None, prefix="c0"):
        self.group = group
        self.store = store
        self.rank = dist.get_rank(group) if rank is None else rank
        self.size = dist.get_world_size(group) if size is None else size
        self.prefix = prefix
        self._next = {}          # (tag) -> next ticket to read from my queue
        self._stash = []         # messages popped while looking for a specific source

    # ---- collectives on python objects
    def Barrier(self):
        if self.size > 1:
            dist.barrier(group=self.group)

    barrier = Barrier

    def bcast(self, obj, root=0):
        if self.size == 1:
            return obj
        box = [obj if self.rank == root else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(self.group, root) if self.group else root, group=self.group)
        return box[0]

    def allgather(self, obj):
        if self.size == 1:
            return [obj]
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    # ---- tagged mailbox (point-to-point control messages)
    def _q(self, dest, tag):
        return "%s/mb/%d/%d" % (self.prefix, dest, tag)

    def send(self, obj, dest, tag=0):
        q = self._q(dest, tag)
        ticket = self.store.add(q + "/n", 1)
        self.store.set("%s/%d" % (q, ticket), pickle.dumps((self.rank, obj)))

    isend = send

    def _pop(self, tag, timeout):
        q = self._q(self.rank, tag)
        nxt = self._next.get(tag, 1)
        key = "%s/%d" % (q, nxt)
        self.store.wait([key], datetime.timedelta(seconds=timeout))
        src, obj = pickle.loads(self.store.get(key))
        try:
            self.store.delete_key(key)
        except Exception:
            pass
        self._next[tag] = nxt + 1
        return src, obj

    def recv(self, source=ANY_SOURCE, tag=0, timeout=3600.0, return_source=False):
        for i, (t, s, o) in enumerate(self._stash):
            if t == tag and (source == ANY_SOURCE or s == source):
                self._stash.pop(i)
                return (o, s) if return_source else o
        while True:
            src, obj = self._pop(tag, timeout)
            if source == ANY_SOURCE or src == source:
                return (obj, src) if return_source else obj
            self._stash.append((tag, src, obj))

    def iprobe(self, source=ANY_SOURCE, tag=0):
        for (t, s, o) in self._stash:
            if t == tag and (source == ANY_SOURCE or s == source):
                return True
        q = self._q(self.rank, tag)
        n = self.store.add(q + "/n", 0)
        nxt = self._next.get(tag, 1)
        if source == ANY_SOURCE:
            return n >= nxt
        while n >= self._next.get(tag, 1):              # pull pending messages into the stash and look again
            src, obj = self._pop(tag, 60.0)
            self._stash.append((tag, src, obj))
            if src == source:
                return True
        return False

    Iprobe = iprobe


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class MPI_GPU_Process(object):
    """Per-process runtime: rendezvous, device binding, communicators.

    Attributes mirror the reference (``base.py:7-20``): ``comm`` (world control plane),
    ``rank``, ``size``, ``ctx`` (here: the torch device), ``gpucomm`` (the data-plane
    handle: a :class:`SymmetricComm` on GPUs, ``None`` on CPU)."""

    def __init__(self, device):
        self.device_str = device
        self.kind, self.dev_index = parse_device(device)
        self.comm = None
        self.gpucomm = None
        self.nccl_group = None
        self.get_internode_comm()
        self.init_device()

    # ------------------------------------------------------------------ rendezvous
    def get_internode_comm(self):
        rank = int(os.environ.get("RANK", "0"))
        size = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank, self.size = rank, size
        if size > 1 or os.environ.get("TMPI_FORCE_DIST") == "1":
            if not dist.is_initialized():
                addr = os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                port = os.environ.setdefault("MASTER_PORT", "29533")
                backend = "gloo"
                if self.kind == "cuda":
                    torch.cuda.set_device(self.dev_index)
                    backend = "cpu:gloo,cuda:nccl"
                dist.init_process_group(backend=backend, init_method="tcp://%s:%s" % (addr, port), rank=rank,
                                        world_size=size, timeout=datetime.timedelta(seconds=1800),
                                        device_id=torch.device("cuda", self.dev_index) if self.kind == "cuda" else None)
            store = dist.distributed_c10d._get_default_store()
            self.comm = Comm(None, store, rank, size, prefix="world")
        else:
            self.comm = Comm(None, _LocalStore(), 0, 1, prefix="world")
        return self.comm

    def init_device(self):
        if self.kind == "cuda":
            if not torch.cuda.is_available():
                raise RuntimeError("device %s requested but CUDA is unavailable" % self.device_str)
            torch.cuda.set_device(self.dev_index)
            self.ctx = torch.device("cuda", self.dev_index)
            from ..ops import native
            native.require()                      # fail loudly: no silent eager fallback on a GPU box
        else:
            self.ctx = torch.device("cpu")
        return self.ctx

    # ------------------------------------------------------------------ data-plane communicators
    def get_intranode_comm(self, arena_bytes=None):
        """World data-plane handle (the reference builds an NCCL clique here,
        ``base.py:22-63``).  On GPUs this only records the local rank layout; the
        symmetric arena is created once the model's size is known (``attach_arena``)."""
        hosts = self.comm.allgather("%s,%d" % (socket.gethostname(), self.rank))
        me = socket.gethostname()
        local = [int(h.split(",")[1]) for h in hosts if h.split(",")[0] == me]
        self.local_ranks = local
        self.local_rank, self.local_size = local.index(self.rank), len(local)
        self.gpucomm = None
        return self.gpucomm

    def attach_arena(self, model, strategy="auto", nbytes=None):
        """Create the peer-mapped symmetric arena and hand the model an allocator for it."""
        from .symmetric import SymmetricComm
        self.gpucomm = SymmetricComm(self.comm, self.ctx, nbytes, local_ranks=getattr(self, "local_ranks", None))
        return self.gpucomm

    def get_intranode_pair_comm(self, pair):
        """2-rank view (``base.py:65-122``).  With symmetric peer memory a 'pair comm' is
        just (my rank, peer rank) — no clique creation, no unique-id exchange."""
        a, b = pair
        assert self.rank in (a, b)
        return PairComm(self.gpucomm, self.comm, self.rank, b if self.rank == a else a)

    def get_intranode_pair_comm_dict(self):
        """All C(size,2) pairs this rank is part of, keyed like the reference
        (``'%d%d' % (lo, hi)``, ``base.py:124-150``)."""
        d = {}
        for other in range(self.size):
            if other == self.rank:
                continue
            lo, hi = min(self.rank, other), max(self.rank, other)
            d["%d%d" % (lo, hi)] = self.get_intranode_pair_comm((lo, hi))
        return d

    def nccl(self):
        """NCCL process group for the baseline strategies (nccl32 / nccl16)."""
        if self.size == 1:
            return None
        return dist.group.WORLD

    def finalize(self):
        if dist.is_initialized():
            try:
                dist.barrier()
            except Exception:
                pass
            dist.destroy_process_group()


class PairComm(object):
    def __init__(self, gpucomm, comm, rank, peer):
        self.gpucomm, self.comm, self.rank, self.peer = gpucomm, comm, rank, peer


class _LocalStore(object):
    """In-process stand-in for the c10d store when world size is 1."""

    def __init__(self):
        self.kv = {}

    def add(self, k, v):
        self.kv[k] = int(self.kv.get(k, 0)) + v
        return self.kv[k]

    def set(self, k, v):
        self.kv[k] = v

    def get(self, k):
        return self.kv[k]

    def wait(self, keys, timeout=None):
        t0 = time.time()
        while not all(k in self.kv for k in keys):
            if timeout is not None and time.time() - t0 > timeout.total_seconds():
                raise RuntimeError("store wait timeout: %s" % keys)
            time.sleep(0.001)

    def delete_key(self, k):
        self.kv.pop(k, None)
