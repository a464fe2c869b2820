"""Multi-process GPU check of the peer-memory runtime and the fused exchange kernels.

Launched by ``tests/test_multigpu.py`` (and by hand) as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29611 tests/mp_fused_check.py [--sweep]

Every rank builds the same small arena inside symmetric memory, fills G with
rank-dependent values, runs each fused algorithm and compares W / U / H with a plain
torch reference computed from an all_gather of the gradients.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    from theanompi_b200.worker import BSP_Worker
    from theanompi_b200.parallel.arena import FlatArena
    from theanompi_b200.ops import reference as ref, native

    worker = BSP_Worker("cuda%d" % local, "cdd", "fused")
    dev = worker.ctx
    alloc = worker.arena_allocator()
    torch.manual_seed(1234)
    shapes = [(257, 300), (257,), (96, 11, 11, 3), (96,), (4096, 1024), (4096,), (33,), (2048, 2049)]
    if "--sweep" in sys.argv:
        shapes.append((int(os.environ.get("TMPI_SWEEP_ROWS", "8192")), 16384))     # 512 MiB fp32 for the bandwidth sweep
    params = [torch.randn(s) * 0.1 for s in shapes]
    wtypes = ["W" if len(s) > 1 else "b" for s in shapes]
    arena = FlatArena(params, wtypes, dev, weight_decay=5e-4, allocator=alloc, with_recv=True)
    gc = worker.gpucomm
    info = dict(rank=rank, mode=gc.mode, multicast=gc.has_multicast, multicast_error=gc.multicast_error, numel=arena.numel)
    L = native.require()
    results = {}

    def reset(seed):
        torch.manual_seed(seed)
        w = torch.randn(arena.numel, device=dev) * 0.1
        dist.broadcast(w, 0)
        arena.W.copy_(w); arena.U.zero_(); arena.refresh_shadow()
        torch.manual_seed(seed * 100 + rank)
        arena.G.copy_(torch.randn(arena.numel, device=dev))
        arena.hyper[0] = 0.05
        torch.cuda.synchronize(); dist.barrier()

    def reference_step(wire16):
        gl = [torch.empty_like(arena.G) for _ in range(world)]
        dist.all_gather(gl, arena.G.clone())
        if wire16:
            gl = [g.to(torch.bfloat16).float() for g in gl]
        gs = torch.stack(gl).sum(0)
        w, u = arena.W.clone(), arena.U.clone()
        ref.sgd_flat(w, gs, u, arena.lr_mult_vector(), arena.wd_vector(), 0.05, 0.9, False, 1.0 / world)
        return w, u

    algos = ["oneshot", "twoshot"] + (["nvls"] if gc.has_multicast else [])
    for algo in algos:
        for wire16 in (False, True):
            reset(7)
            w_ref, u_ref = reference_step(wire16)
            for rep in range(3):                              # repeat: exercises flag-epoch reuse
                if rep:
                    reset(7)
                gc.fused_allreduce_sgd(arena, 0, arena.numel, 0.9, False, algo=algo, wire16=wire16, max_blocks=24)
                torch.cuda.synchronize(); dist.barrier()
            ew = float((arena.W - w_ref).abs().max())
            eh = float((arena.H.float() - w_ref).abs().max())
            # every rank must hold bit-identical weights
            wl = [torch.empty_like(arena.W) for _ in range(world)]
            dist.all_gather(wl, arena.W.clone())
            same = all(torch.equal(wl[0], x) for x in wl)
            tol = 2e-3 if wire16 else 1e-5
            results["%s%s" % (algo, "16" if wire16 else "")] = dict(err_w=ew, err_h=eh, identical=same)
            assert ew < tol, (algo, wire16, ew)
            assert eh < 5e-3 + tol, (algo, wire16, eh)
            assert same, (algo, wire16)
            if algo == "oneshot":
                eu = float((arena.U - u_ref).abs().max())
                assert eu < tol * 10, eu

    # owner-keeps-master: the fused kernel ships only the bf16 shadow of an updated slice; the fp32 master copies of non-owners
    # go stale until push_master_slices() re-synchronises them
    for algo in [a for a in algos if a != "oneshot"]:
        reset(8)
        w_ref, u_ref = reference_step(False)
        gc.fused_allreduce_sgd(arena, 0, arena.numel, 0.9, False, algo=algo, max_blocks=24, push_master=False)
        torch.cuda.synchronize(); dist.barrier()
        hl = [torch.empty_like(arena.H) for _ in range(world)]
        dist.all_gather(hl, arena.H.clone())
        assert all(torch.equal(hl[0], x) for x in hl), ("shadows differ", algo)
        assert float((arena.H.float() - w_ref).abs().max()) < 5e-3, algo
        nb = arena.numel // 1024
        per = (nb + world - 1) // world
        lo_, hi_ = rank * per * 1024, min(arena.numel, (rank + 1) * per * 1024)
        assert float((arena.W[lo_:hi_] - w_ref[lo_:hi_]).abs().max()) < 1e-5, ("owner slice", algo)
        gc.push_master_slices(arena, 0, arena.numel)
        torch.cuda.synchronize(); dist.barrier()
        assert float((arena.W - w_ref).abs().max()) < 1e-5, ("after push_master_slices", algo)
        wl = [torch.empty_like(arena.W) for _ in range(world)]
        dist.all_gather(wl, arena.W.clone())
        assert all(torch.equal(wl[0], x) for x in wl), ("masters differ after sync", algo)
        results["owner_keeps_master_" + algo] = True

    # sub-range (bucket) exchange leaves the rest untouched
    reset(9)
    w0 = arena.W.clone()
    b = arena.make_buckets(1 << 20)[0]
    gc.fused_allreduce_sgd(arena, b["lo"], b["hi"], 0.9, False, algo="twoshot", max_blocks=8)
    torch.cuda.synchronize(); dist.barrier()
    assert torch.equal(arena.W[:b["lo"]], w0[:b["lo"]])
    assert not torch.equal(arena.W[b["lo"]:b["hi"]], w0[b["lo"]:b["hi"]])

    # reduce-scatter fused into the wgrad GEMM epilogue: every rank's tcgen05 GEMM red.adds its dW tiles into the OWNER's G over
    # NVLink; the exchange kernel then only updates its slice, pushes W / H and clears G (pre_reduced)
    from theanompi_b200.ops import cuda_impl
    ti = 4                                                       # the (4096, 1024) tensor
    lo = arena.offsets[ti]
    hi = arena.offsets[ti + 1] if ti + 1 < len(arena.offsets) else arena.numel
    O_, I_, B_ = 4096, 1024, 128
    for algo in ["twoshot"] + (["nvls"] if gc.has_multicast else []):
        reset(21)
        arena.G.zero_()
        torch.cuda.synchronize(); dist.barrier()
        torch.manual_seed(500 + rank)
        dym = torch.randn(B_, O_, device=dev).to(torch.bfloat16)
        xin = torch.randn(B_, I_, device=dev).to(torch.bfloat16)
        dws = [torch.empty(O_, I_, device=dev) for _ in range(world)]
        dist.all_gather(dws, dym.float().t() @ xin.float())
        gs = torch.zeros_like(arena.G)
        gs[lo:lo + O_ * I_] = torch.stack(dws).sum(0).reshape(-1)
        w_ref, u_ref = arena.W.clone(), arena.U.clone()
        ref.sgd_flat(w_ref, gs, u_ref, arena.lr_mult_vector(), arena.wd_vector(), 0.05, 0.9, False, 1.0 / world)
        gc.configure_gemm_rs(arena, [(lo, hi)])
        dw_view = arena.G[lo:lo + O_ * I_].view(O_, I_)
        for rep in range(2):                                     # second round: G must have been cleared by the first
            if rep:
                arena.W.copy_(w0_rs); arena.U.zero_(); arena.refresh_shadow()
                torch.cuda.synchronize(); dist.barrier()
            else:
                w0_rs = arena.W.clone()
            cuda_impl.gemm(dym, xin, O_, I_, B_, a_mn=True, b_mn=True, out=dw_view, lda=O_, ldb=I_, ldc=I_)
            gc.fused_allreduce_sgd(arena, lo, hi, 0.9, False, algo=algo, max_blocks=24, pre_reduced=True)
            torch.cuda.synchronize(); dist.barrier()
            ew = float((arena.W[lo:hi] - w_ref[lo:hi]).abs().max())
            assert ew < 2e-4, ("gemm reduce-scatter", algo, rep, ew)
            assert float(arena.G[lo:hi].abs().max()) == 0.0, "G not cleared by the pre_reduced exchange"
            wl = [torch.empty_like(arena.W) for _ in range(world)]
            dist.all_gather(wl, arena.W.clone())
            assert all(torch.equal(wl[0][lo:hi], x[lo:hi]) for x in wl)
        results["gemm_rs_" + algo] = dict(err_w=ew)
        L.gemm_rs_clear()

    # plain allreduce G -> R (sum) and in-place weight averaging
    for algo in algos:
        reset(11)
        gl = [torch.empty_like(arena.G) for _ in range(world)]
        dist.all_gather(gl, arena.G.clone())
        want = torch.stack(gl).sum(0)
        gc.allreduce(arena, "G", "R", 1.0, algo=algo)
        torch.cuda.synchronize(); dist.barrier()
        assert float((arena.R - want).abs().max()) < 1e-4, algo
    reset(12)
    arena.W.add_(rank)                                           # make replicas differ
    torch.cuda.synchronize(); dist.barrier()
    wl = [torch.empty_like(arena.W) for _ in range(world)]
    dist.all_gather(wl, arena.W.clone())
    want = torch.stack(wl).mean(0)
    gc.allreduce(arena, "W", "W", 1.0 / world, refresh_shadow=True)
    torch.cuda.synchronize(); dist.barrier()
    assert float((arena.W - want).abs().max()) < 1e-5
    assert float((arena.H.float() - want).abs().max()) < 2e-2

    # EASGD elastic kernel: rank 1 is the worker, rank 0 holds the center
    reset(13)
    arena.W.add_(0.5 * rank)
    torch.cuda.synchronize(); dist.barrier()
    wl = [torch.empty_like(arena.W) for _ in range(world)]
    dist.all_gather(wl, arena.W.clone())
    if rank == 1:
        center = gc.peer_region(0, arena.layout["W"], arena.numel)
        L.easgd_elastic(arena.W.data_ptr(), arena.H.data_ptr(), center.data_ptr(), 0.5, arena.numel, 64,
                        torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    dist.barrier()
    d = 0.5 * (wl[1] - wl[0])
    if rank == 1:
        assert float((arena.W - (wl[1] - d)).abs().max()) < 1e-6
    if rank == 0:
        assert float((arena.W - (wl[0] + d)).abs().max()) < 1e-6

    # GOSGD pull-merge: rank 0 merges rank 1's snapshot (in R) with weights 0.25 / 0.125
    reset(14)
    arena.R.copy_(arena.W + 1.0 + rank)
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        snap = gc.peer_region(1, arena.layout["R"], arena.numel)
        want = (0.25 * arena.W + 0.125 * snap.clone()) / 0.375
        L.gosgd_merge(arena.W.data_ptr(), arena.H.data_ptr(), snap.data_ptr(), 0.25, 0.125, arena.numel, 64,
                      torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert float((arena.W - want).abs().max()) < 1e-6
    dist.barrier()

    if "--sweep" in sys.argv:
        sweep = bus_sweep(gc, arena, world, algos, dev)
        if rank == 0:
            results["sweep"] = sweep
    if rank == 0:
        print("MP_FUSED_CHECK_OK " + json.dumps(dict(info=info, results=results)))
    worker.finalize()


def bus_sweep(gc, arena, world, algos, dev):
    """Allreduce (G→R) time vs message size: hand-written kernels vs NCCL on the same buffers."""
    out = []
    sizes = [1 << k for k in range(12, 40) if (1 << k) <= arena.numel * 4]
    sizes.append(arena.numel * 4)
    for nbytes in sizes:
        n = (nbytes // 4 // 1024) * 1024
        if n == 0:
            continue
        row = dict(bytes=n * 4)
        for algo in algos + ["nccl"]:
            def run():
                if algo == "nccl":
                    dist.all_reduce(arena.R[:n])
                else:
                    gc.allreduce(arena, "G", "R", 1.0, lo=0, hi=n, algo=algo)
            for _ in range(3):
                run()
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                run()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
            row[algo + "_us"] = ms * 1000
            row[algo + "_busGBs"] = (n * 4) * 2 * (world - 1) / world / (ms / 1000) / 1e9
        out.append(row)
    return out


if __name__ == "__main__":
    main()
