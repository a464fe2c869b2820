"""Multi-process GPU checks of the device-side protocols and of every legacy exchange strategy.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29613 \
        tests/mp_proto_check.py [strategies] [easgd] [gosgd] [barrier]

* ``strategies`` — the reference's own test suite (``test/test-exchanger/test_exchanger.py:39-46``, ``test_nccl32.py:57-79``,
  ``test_nccl16.py:66-88``, ``test_asa32.py``, ``test_copper.py``): every strategy (ar, nccl32, nccl16, asa32, asa16, copper,
  copper16) on a 230,400-float vector (+ odd shapes) against the ground-truth sum, sum and average, ON GPUs; swap / replace.
* ``easgd``  — 1 center + (N−1) workers hammer the center with τ = 1 through the device-side ticket lock (and through the
  lock-free red.add variant): no update may be lost — Σ over workers of the applied deltas equals the center's drift.
* ``gosgd``  — the device-side gossip protocol under heavy traffic (p = 0.5): the push-sum weights still add up to 1, the
  α-weighted mean of the replicas is conserved, every admitted push is merged exactly once.
* ``barrier`` — 10,000 device-side flag barriers with skewed launch order + CUDA-graph replays of a captured barrier chain.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def case_strategies(worker, out):
    from theanompi_b200.parallel import exchanger_strategy as ES
    rank, size, dev = worker.rank, worker.size, worker.ctx
    shapes = [(230400,), (96, 11, 11, 3), (37,), (1000, 129)]
    names = ["ar", "nccl32", "nccl16", "asa32", "asa16"] + (["copper", "copper16"] if size & (size - 1) == 0 else [])
    res = {}
    for name in names:
        for avg in (False, True):
            rs = np.random.RandomState(100 + rank)
            src = [torch.from_numpy(rs.randn(*s).astype(np.float32)).to(dev) for s in shapes]
            dst = [torch.zeros_like(t) for t in src]
            all_src = []
            for r in range(size):
                rr = np.random.RandomState(100 + r)
                all_src.append([rr.randn(*s).astype(np.float32) for s in shapes])
            want = [sum(all_src[r][i] for r in range(size)) / (size if avg else 1.0) for i in range(len(shapes))]
            cls = {"ar": ES.Exch_allreduce, "nccl32": ES.Exch_nccl32, "nccl16": ES.Exch_nccl16, "asa32": ES.Exch_asa32,
                   "asa16": ES.Exch_asa16, "copper": ES.Exch_copper, "copper16": ES.Exch_copper16}[name]
            g = worker.nccl()
            if name == "ar":
                ex = cls(worker.comm, avg=avg)
            elif name in ("nccl32", "nccl16"):
                ex = cls(worker.comm, g, avg=avg)
            else:
                ex = cls(worker.comm, avg=avg, group=g)
            ex.prepare(dev, src, dst)
            ex.exchange()
            torch.cuda.synchronize()
            tol = 2e-2 if name.endswith("16") else 1e-5
            worst = 0.0
            for d, w in zip(dst, want):
                err = float((d.cpu() - torch.from_numpy(w)).abs().max())
                worst = max(worst, err / max(1.0, float(np.abs(w).max())))
            assert worst < tol, (name, avg, worst)
            res["%s/%s" % (name, "avg" if avg else "sum")] = worst
    # swap: partners exchange parameters; replace: the winner overwrites the loser
    ex = ES.Exch_swap(worker.comm, group=worker.nccl())
    mine = [torch.full((1024,), float(rank), device=dev)]
    ex.prepare(dev, mine)
    ex.exchange()
    torch.cuda.synchronize()
    got = worker.comm.allgather(float(mine[0][0]))
    assert sorted(got) == [float(r) for r in range(size)], got                 # a permutation of the ranks
    if size == 2:
        assert got[rank] == float(1 - rank)
    ex.replace(winner_ranks=[0])
    torch.cuda.synchronize()
    worker.comm.Barrier()
    res["swap"] = 0.0
    out["strategies"] = res


def _arena(worker, numel_rows=2048):
    from theanompi_b200.parallel.arena import FlatArena
    alloc = worker.arena_allocator()
    params = [torch.randn(numel_rows, 1024) * 0.1, torch.randn(4096) * 0.1, torch.randn(33) * 0.1]
    arena = FlatArena(params, ["W", "b", "b"], worker.ctx, weight_decay=0.0, allocator=alloc, with_recv=True)
    return arena


def case_easgd(worker, arena, out):
    from theanompi_b200.parallel.exchanger import EASGD_Exchanger
    rank, size, dev = worker.rank, worker.size, worker.ctx
    gc = worker.gpucomm
    res = {}
    for lockfree in (0, 1):
        os.environ["TMPI_EASGD_LOCKFREE"] = str(lockfree)
        torch.manual_seed(77 + rank)
        arena.W.copy_(torch.randn(arena.numel, device=dev) * (1.0 + rank))      # replicas far apart: a lost update is O(1)
        arena.refresh_shadow()
        torch.cuda.synchronize(); dist.barrier()
        w0 = arena.W.double().clone()
        ex = EASGD_Exchanger(0.25, [], "worker" if rank > 0 else "server", comm=worker.comm, gpucomm=gc, arena=arena, server_rank=0)
        served0 = int(gc.proto_words(0)[2].item()) if rank == 0 else 0
        E = 40
        dist.barrier()
        if rank > 0:
            for _ in range(E):
                ex.exchange()                                                # no host sync between exchanges: they queue on the device
        torch.cuda.synchronize(); dist.barrier()
        applied = (w0 - arena.W.double()) if rank > 0 else torch.zeros_like(w0)      # Σ of this worker's deltas
        dist.all_reduce(applied)
        drift = (arena.W.double() - w0) if rank == 0 else torch.zeros_like(w0)
        dist.all_reduce(drift)
        err = float((applied - drift).abs().max())
        scale = float(drift.abs().max())
        res["lockfree" if lockfree else "ticket"] = dict(max_abs_err=err, drift_scale=scale, exchanges=E * (size - 1))
        assert err < 2e-4 * max(1.0, scale), ("lost update?", lockfree, err, scale)
        if rank == 0 and not lockfree:
            served = int(gc.proto_words(0)[2].item()) - served0
            assert served == E * (size - 1), (served, E * (size - 1))
        # the shadow follows the weights
        if rank > 0 and arena.H is not None:
            assert float((arena.H.float() - arena.W).abs().max()) < 0.05 * max(1.0, float(arena.W.abs().max()))
        dist.barrier()
    os.environ["TMPI_EASGD_LOCKFREE"] = "0"
    out["easgd"] = res


class _M(object):
    pass


def case_gosgd(worker, arena, out):
    from theanompi_b200.parallel.exchanger import GOSGD_Exchanger
    rank, size, dev = worker.rank, worker.size, worker.ctx
    torch.manual_seed(5 + rank)
    arena.W.copy_(torch.randn(arena.numel, device=dev) + rank)
    arena.refresh_shadow()
    torch.cuda.synchronize(); dist.barrier()
    wl = [torch.empty_like(arena.W) for _ in range(size)]
    dist.all_gather(wl, arena.W.clone())
    mean0 = torch.stack(wl).double().mean(0)
    m = _M(); m.arena = arena
    ex = GOSGD_Exchanger(worker.comm, worker.gpucomm, m, p=0.5, seed=4242 + rank)
    T = 300
    for it in range(T):
        ex.process_messages(None)
        if ex.draw():
            ex.push_message(ex.choose(), None)
        if it % 7 == rank % 7:
            time.sleep(0.0005)                                               # skew the ranks
    ex.finish(None)
    torch.cuda.synchronize(); dist.barrier()
    alphas = worker.comm.allgather(ex.alpha)
    counts = worker.comm.allgather((ex.n_pushed, ex.n_skipped, ex.n_merged))
    assert abs(sum(alphas) - 1.0) < 1e-5, alphas
    assert sum(c[0] for c in counts) == sum(c[2] for c in counts), counts       # every admitted push merged exactly once
    assert sum(c[0] for c in counts) > 0
    wsum = arena.W.double() * ex.alpha
    dist.all_reduce(wsum)
    err = float((wsum - mean0).abs().max())
    assert err < 1e-3 * max(1.0, float(mean0.abs().max())), err
    out["gosgd"] = dict(alphas=alphas, pushed=sum(c[0] for c in counts), skipped=sum(c[1] for c in counts),
                        merged=sum(c[2] for c in counts), mass_err=err)


def case_barrier(worker, out):
    gc = worker.gpucomm
    rank = worker.rank
    rs = np.random.RandomState(99 + rank)
    t0 = time.time()
    N = 10000
    for i in range(N):
        gc.barrier()
        if i % 997 == (rank * 131) % 997:
            time.sleep(0.002 * rs.rand())                                     # skewed launch order: some rank is always late
        if i % 2000 == 1999:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    t_eager = time.time() - t0
    # captured chain of barriers, replayed (the epochs live in device memory, so a replay continues the count)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        gc.barrier()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10):
                gc.barrier()
    torch.cuda.current_stream().wait_stream(s)
    dist.barrier()
    for i in range(500):
        g.replay()
        if i % 50 == rank:
            time.sleep(0.001)
    torch.cuda.synchronize()
    dist.barrier()
    out["barrier"] = dict(eager=N, eager_s=t_eager, graph_replays=500, barriers_per_replay=10)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    from theanompi_b200.worker import BSP_Worker
    cases = [a for a in sys.argv[1:] if not a.startswith("-")] or ["strategies", "easgd", "gosgd", "barrier"]
    worker = BSP_Worker("cuda%d" % local, "cdd", "fused")
    out = {}
    arena = _arena(worker)
    if "strategies" in cases:
        case_strategies(worker, out)
    if "easgd" in cases:
        case_easgd(worker, arena, out)
    if "gosgd" in cases:
        case_gosgd(worker, arena, out)
    if "barrier" in cases:
        case_barrier(worker, out)
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        print("MP_PROTO_CHECK_OK " + json.dumps(out))
    worker.finalize()


if __name__ == "__main__":
    main()
