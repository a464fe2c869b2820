"""Convergence evidence (ref README.md:122-123: "1/2/4/8-GPU validation curves coincide"; VERDICT: bf16 vs tf32 curves).

    python scripts/convergence.py --steps 320                       # 1 GPU: bf16 vs tf32 training curves on the separable synthetic set
    python -m torch.distributed.run --nproc-per-node N ... scripts/convergence.py --steps 320 --bsp     # N GPUs, BSP (fused exchange),
                                                                     global batch fixed at 128: per-GPU batch 128 / N

Model: Cifar10_model on the synthetic class-separable CIFAR-shaped set (no dataset in this environment).  Prints one JSON line
with (images seen, smoothed training loss, validation cost, validation error) every 20 steps.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def run_one(dtype, steps, rank, size, batch, worker=None, lr=0.001, overlap=True, comm_blocks=0, graph=False):
    from theanompi_b200.models import layers2
    from theanompi_b200.models.cifar10 import Cifar10_model
    from theanompi_b200.utils.recorder import Recorder
    layers2.reseed(); layers2.Dropout.layers.clear(); layers2.Crop.layers.clear(); layers2.BatchNormal.layers.clear()
    dev = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
    cfg = dict(verbose=False, rank=rank, size=size, device=dev, batch_size=batch, file_batch_size=batch, learning_rate=lr, dtype=dtype,
               data_kwargs=dict(n_synthetic=8192, synthetic=True))
    if worker is not None:
        cfg = worker.model_config("Cifar10_model", **{k: v for k, v in cfg.items() if k not in ("rank", "size", "device", "verbose")})
        cfg["verbose"] = False
        cfg["overlap"] = overlap
        if comm_blocks:
            cfg["comm_blocks"] = comm_blocks
        if graph:
            cfg["cuda_graph"] = True
    m = Cifar10_model(cfg)
    if worker is not None:
        worker.build(m, cfg)
        rec, exch = worker.recorder, worker.exchanger
    else:
        m.compile_iter_fns("avg")
        rec, exch = Recorder(None, 10 ** 6, "c", False, device=dev), None
    curve = []
    verr = None
    for i in range(steps):
        m.train_iter(i, rec)
        if exch is not None:
            exch.exchange(rec)
        if (i + 1) % 20 == 0:
            loss = float(torch.stack([c.float() for c in rec.train_info["cost"][-20:]]).mean())
            # validation cost / error of the current weights (eval mode: no dropout, centre crop) on the first val batches —
            # the smooth curve the reference plots (README.md:122-123)
            n0 = len(rec.val_info["cost"])
            for j in range(min(2, m.data.n_batch_val)):
                m.val_iter(j, rec)
            m.reset_iter("val")
            vc = float(torch.stack([torch.as_tensor(c).float() for c in rec.val_info["cost"][n0:]]).mean())
            verr = float(torch.stack([torch.as_tensor(e).float() for e in rec.val_info["error"][n0:]]).mean())
            curve.append((int((i + 1) * batch * size), round(loss, 4), round(vc, 4), round(verr, 4)))
    m.cleanup()
    return curve, verr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=320)
    ap.add_argument("--bsp", action="store_true")
    ap.add_argument("--global-batch", type=int, default=128)
    ap.add_argument("--strategy", default="fused")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--comm-blocks", type=int, default=0)
    ap.add_argument("--graph", action="store_true")
    a = ap.parse_args()
    rank, size = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    out = {"steps": a.steps, "n_gpus": size, "global_batch": a.global_batch}
    if a.bsp or size > 1:
        from theanompi_b200.worker import BSP_Worker
        w = BSP_Worker("cuda%d" % int(os.environ.get("LOCAL_RANK", "0")), "cdd", a.strategy)
        curve, verr = run_one("bf16", a.steps, rank, size, a.global_batch // size, worker=w if size > 1 else None, overlap=not a.no_overlap,
                              comm_blocks=a.comm_blocks, graph=a.graph)
        out["bsp_bf16"] = {"curve": curve, "val_err": verr, "strategy": a.strategy, "push_master": os.environ.get("TMPI_PUSH_MASTER", "0"),
                           "nvls": os.environ.get("TMPI_NVLS", "1")}
        if rank == 0:
            print("CONVERGENCE " + json.dumps(out), flush=True)
        w.finalize()
        return
    for dt in ("bf16", "tf32"):
        curve, verr = run_one(dt, a.steps, 0, 1, a.global_batch)
        out[dt] = {"curve": curve, "val_err": verr}
    out["max_abs_train_loss_gap_bf16_vs_tf32"] = max(abs(x[1] - y[1]) for x, y in zip(out["bf16"]["curve"], out["tf32"]["curve"]))
    out["max_abs_val_cost_gap_bf16_vs_tf32"] = max(abs(x[2] - y[2]) for x, y in zip(out["bf16"]["curve"], out["tf32"]["curve"]))
    print("CONVERGENCE " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
