"""Allreduce strategies behind ``BSP_Exchanger`` (ref ``theanompi/lib/exchanger_strategy.py``).

Every class keeps the reference's two-call protocol ``prepare(ctx, source_param_list,
dest_param_list)`` / ``exchange()`` and sums ``source[i]`` over ranks into ``dest[i]``
(optionally pre-dividing by the world size, ``avg=True``).

===============  ==========================================================================================
``ar``           host-staged allreduce through gloo — the reference's ``MPI.Allreduce`` path
                 (``:25-72``) and, like there, the multi-node fallback.  Also the CPU test path.
``nccl32``       one ``ncclAllReduce(fp32)`` per tensor (``:75-127``) — **the reference-semantics baseline**.
``nccl16``       fp32→fp16 cast kernel, fp16 NCCL allreduce, fp16→fp32 (``:129-237``; K1 kernels).
``asa32/asa16``  alltoall → ``sumfloats``/``sumhalfs`` (K2/K3) → allgather (``:240-607``), with the
                 reference's kernel bugs fixed.
``copper(16)``   binary-tree reduce + tree broadcast with pairwise send/recv and ``vecadd(half)`` (K4/K5)
                 (``:610-1590``); generalised from the hard-coded sizes {2,4,8,16} to any power of two.
``swap``         random disjoint pairing, partners swap (or winner→loser replace) parameters (``:1593-1773``).
``p2p32``        hand-written one-shot / two-shot / NVLS allreduce kernel over the symmetric arena
                 (no NCCL) producing ``dest`` — the un-fused sibling of the ``fused*`` strategies.
===============  ==========================================================================================

The product path (``fused``, ``fused16``, ``oneshot``, ``twoshot``, ``nvls`` …) is not a
strategy object of this kind: it fuses the reduction WITH the optimizer update in one
kernel and lives in :class:`theanompi_b200.parallel.exchanger.BSP_Exchanger`.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def _native():
    from ..ops import native
    return native.require()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class Exch_strategy(object):
    def __init__(self):
        self.source_param_list = None
        self.dest_param_list = None

    def prepare(self, ctx, source_param_list, dest_param_list=None):
        self.ctx = ctx
        self.source_param_list = list(source_param_list)
        self.dest_param_list = list(dest_param_list) if dest_param_list is not None else self.source_param_list

    def exchange(self):
        raise NotImplementedError


class Exch_allreduce(Exch_strategy):
    """Host allreduce ('ar')."""

    def __init__(self, comm, avg=True, group=None):
        super().__init__()
        self.comm, self.avg, self.group = comm, avg, group
        self.size = comm.size

    def exchange(self):
        if self.size == 1:
            return
        for s, d in zip(self.source_param_list, self.dest_param_list):
            host = s.detach().to("cpu", torch.float32)
            if self.avg:
                host = host / self.size
            elif host.data_ptr() == s.data_ptr():
                host = host.clone()                    # CPU tensors: never reduce in place into the send buffer
            host = host.contiguous()
            dist.all_reduce(host, op=dist.ReduceOp.SUM, group=self.group)
            d.copy_(host.to(d.device), non_blocking=False)


class Exch_nccl32(Exch_strategy):
    """One fp32 ncclAllReduce per tensor — reference default & benchmarked path."""

    def __init__(self, intercomm, intracomm=None, avg=False):
        super().__init__()
        self.intercomm, self.intracomm, self.avg = intercomm, intracomm, avg
        self.size = intercomm.size

    def exchange(self):
        if self.size == 1:
            return
        for s, d in zip(self.source_param_list, self.dest_param_list):
            if d.data_ptr() != s.data_ptr():
                d.copy_(s)
            if self.avg:
                d.div_(self.size)
            dist.all_reduce(d, op=dist.ReduceOp.SUM, group=self.intracomm)


class Exch_nccl16(Exch_strategy):
    """fp16 on the wire, fp32 master (K1 cast kernels + NCCL fp16 allreduce)."""

    def __init__(self, intercomm, intracomm=None, avg=False, wire_dtype=torch.float16):
        super().__init__()
        self.intercomm, self.intracomm, self.avg = intercomm, intracomm, avg
        self.size = intercomm.size
        self.wire_dtype = wire_dtype

    def prepare(self, ctx, source_param_list, dest_param_list=None):
        super().prepare(ctx, source_param_list, dest_param_list)
        self.source_param_list_fp16 = [torch.empty(p.shape, dtype=self.wire_dtype, device=p.device)
                                       for p in self.source_param_list]

    def _cast(self, src, dst, kind):
        if src.is_cuda:
            _native().cast_flat(src.data_ptr(), dst.data_ptr(), src.numel(), kind, _stream(src))
        else:
            dst.copy_(src)

    def exchange(self):
        if self.size == 1:
            return
        to16 = 0 if self.wire_dtype == torch.float16 else 2
        for s, h, d in zip(self.source_param_list, self.source_param_list_fp16, self.dest_param_list):
            src = s / self.size if self.avg else s
            self._cast(src.contiguous(), h, to16)
            dist.all_reduce(h, op=dist.ReduceOp.SUM, group=self.intracomm)
            self._cast(h, d, to16 + 1)


class Exch_asa32(Exch_strategy):
    """alltoall – sum – allgather: a hand-rolled reduce-scatter + all-gather."""

    wire_dtype = torch.float32

    def __init__(self, comm, avg=False, group=None):
        super().__init__()
        self.comm, self.avg, self.group = comm, avg, group
        self.size = comm.size

    def prepare(self, ctx, source_param_list, dest_param_list=None):
        super().prepare(ctx, source_param_list, dest_param_list)
        self.pad = max(8, self.size)
        self.bufs = []
        for p in self.source_param_list:
            n = p.numel()
            npad = (n + self.pad - 1) // self.pad * self.pad
            npad = (npad + self.size - 1) // self.size * self.size
            dev = p.device
            self.bufs.append(dict(n=n, npad=npad,
                                  send=torch.zeros(npad, dtype=self.wire_dtype, device=dev),
                                  tmp=torch.zeros(npad, dtype=self.wire_dtype, device=dev),
                                  red=torch.zeros(npad // self.size, dtype=self.wire_dtype, device=dev),
                                  out=torch.zeros(npad, dtype=self.wire_dtype, device=dev)))

    def _sum_chunks(self, tmp, red, chunk):
        if tmp.is_cuda:
            _native().sum_chunks(tmp.data_ptr(), red.data_ptr(), chunk, self.size,
                                 int(self.wire_dtype == torch.float16), _stream(tmp))
        else:
            red.copy_(tmp.view(self.size, chunk).float().sum(0))

    def exchange(self):
        if self.size == 1:
            return
        for s, d, b in zip(self.source_param_list, self.dest_param_list, self.bufs):
            flat = s.reshape(-1)
            b["send"][:b["n"]].copy_(flat / self.size if self.avg else flat)
            dist.all_to_all_single(b["tmp"], b["send"], group=self.group)
            chunk = b["npad"] // self.size
            self._sum_chunks(b["tmp"], b["red"], chunk)
            dist.all_gather_into_tensor(b["out"], b["red"], group=self.group)
            d.copy_(b["out"][:b["n"]].view(d.shape))


class Exch_asa16(Exch_asa32):
    wire_dtype = torch.float16


class Exch_copper(Exch_strategy):
    """Binary-tree reduce to rank 0 followed by a tree broadcast, pairwise send/recv +
    ``vecadd``.  On the reference's PCIe/QPI box the tree followed the physical
    topology; on an NVSwitch box every pair is equidistant, so this exists for parity
    (and as the latency-friendly log2(N) pattern)."""

    wire_dtype = torch.float32

    def __init__(self, comm, avg=False, group=None):
        super().__init__()
        self.comm, self.avg, self.group = comm, avg, group
        self.size, self.rank = comm.size, comm.rank
        if self.size & (self.size - 1):
            raise ValueError("copper needs a power-of-two world size")

    def prepare(self, ctx, source_param_list, dest_param_list=None):
        super().prepare(ctx, source_param_list, dest_param_list)
        self.cur = [torch.empty(p.numel(), dtype=self.wire_dtype, device=p.device) for p in self.source_param_list]
        self.tmp = [torch.empty_like(c) for c in self.cur]

    def _vecadd(self, cur, tmp):
        if cur.is_cuda:
            _native().vecadd(cur.data_ptr(), tmp.data_ptr(), cur.numel(), int(self.wire_dtype == torch.float16), _stream(cur))
        else:
            cur.add_(tmp)

    def exchange(self):
        if self.size == 1:
            return
        for s, d, cur, tmp in zip(self.source_param_list, self.dest_param_list, self.cur, self.tmp):
            flat = s.reshape(-1)
            cur.copy_(flat / self.size if self.avg else flat)
            step = 1
            while step < self.size:                                # reduce: 1→0, 3→2 …; then 2→0, 6→4 …
                if self.rank % (2 * step) == step:
                    dist.send(cur, self.rank - step, group=self.group)
                elif self.rank % (2 * step) == 0:
                    dist.recv(tmp, self.rank + step, group=self.group)
                    self._vecadd(cur, tmp)
                step *= 2
            step = self.size // 2
            while step >= 1:                                       # broadcast back down the same tree
                if self.rank % (2 * step) == 0:
                    dist.send(cur, self.rank + step, group=self.group)
                elif self.rank % (2 * step) == step:
                    dist.recv(cur, self.rank - step, group=self.group)
                step //= 2
            d.copy_(cur.view(d.shape))


class Exch_copper16(Exch_copper):
    wire_dtype = torch.float16


class Exch_swap(Exch_strategy):
    """Random disjoint pairing drawn by rank 0 and broadcast; partners swap their
    parameters (``exchange``) or the winner overwrites the loser (``replace``) — the
    "GAP" protocol for parallel GAN training (``examples/bsp/session_gap.cfg``)."""

    def __init__(self, comm, group=None, seed=1234):
        super().__init__()
        self.comm, self.group = comm, group
        self.size, self.rank = comm.size, comm.rank
        self.rs = np.random.RandomState(seed)

    def get_pairs(self):
        pairs = None
        if self.rank == 0:
            perm = self.rs.permutation(self.size).tolist()
            pairs = [(perm[i], perm[i + 1]) for i in range(0, self.size - 1, 2)]
        return self.comm.bcast(pairs, root=0)

    def _partner(self, pairs):
        for a, b in pairs:
            if self.rank == a:
                return b
            if self.rank == b:
                return a
        return None

    def exchange(self):
        if self.size == 1:
            return
        other = self._partner(self.get_pairs())
        if other is None:
            return
        for p in self.source_param_list:
            recv = torch.empty_like(p)
            ops = [dist.P2POp(dist.isend, p.contiguous(), other, group=self.group),
                   dist.P2POp(dist.irecv, recv, other, group=self.group)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            p.copy_(recv)

    def replace(self, winner_ranks):
        """Every rank paired with a winner receives the winner's parameters."""
        if self.size == 1:
            return
        other = self._partner(self.get_pairs())
        if other is None:
            return
        i_win, o_win = self.rank in winner_ranks, other in winner_ranks
        if i_win == o_win:
            return
        for p in self.source_param_list:
            if i_win:
                dist.send(p.contiguous(), other, group=self.group)
            else:
                dist.recv(p, other, group=self.group)


class Exch_p2p32(Exch_strategy):
    """Hand-written peer-memory allreduce (no NCCL) of an arena region into another."""

    def __init__(self, gpucomm, arena, src_region, dst_region, avg=False, algo="auto"):
        super().__init__()
        self.gpucomm, self.arena = gpucomm, arena
        self.src, self.dst, self.avg, self.algo = src_region, dst_region, avg, algo

    def exchange(self):
        if self.gpucomm.size == 1:
            return
        scale = 1.0 / self.gpucomm.size if self.avg else 1.0
        self.gpucomm.allreduce(self.arena, self.src, self.dst, scale, algo=self.algo,
                               refresh_shadow=(self.dst == "W"))
