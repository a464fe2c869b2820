"""Debug: after a few fused steps in owner-keeps-master mode, compare the replicas' shadows / masters parameter by parameter."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from theanompi_b200.worker import BSP_Worker
from theanompi_b200.models import layers2
from theanompi_b200.models.cifar10 import Cifar10_model

rank, size = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
w = BSP_Worker("cuda%d" % local, "cdd", "fused")
layers2.reseed()
cfg = w.model_config("Cifar10_model", batch_size=64, file_batch_size=64, learning_rate=0.001, data_kwargs=dict(n_synthetic=2048, synthetic=True))
cfg["verbose"] = False
m = Cifar10_model(cfg)
w.build(m, cfg)
rec, ex = w.recorder, w.exchanger
a = m.arena
print(rank, "push_master", ex.push_master, "buckets", [(b["lo"], b["hi"], len(b["params"])) for b in ex.buckets], "numel", a.numel,
      "algos", [w.gpucomm.pick_algo((b["hi"] - b["lo"]) * 4, ex.algo) for b in ex.buckets], flush=True)
for i in range(6):
    m.train_iter(i, rec); ex.exchange(rec)
torch.cuda.synchronize(); dist.barrier()
Hs = [torch.empty_like(a.H) for _ in range(size)]; dist.all_gather(Hs, a.H.clone())
Ws = [torch.empty_like(a.W) for _ in range(size)]; dist.all_gather(Ws, a.W.clone())
Us = [torch.empty_like(a.U) for _ in range(size)]; dist.all_gather(Us, a.U.clone())
if rank == 0:
    nb = a.numel // 1024
    for pi, (o, s_, p) in enumerate(zip(a.offsets, a.sizes, a.params)):
        sl = slice(o, o + s_)
        dh = float((Hs[0][sl].float() - Hs[1][sl].float()).abs().max())
        dw = float((Ws[0][sl] - Ws[1][sl]).abs().max())
        # owner of each element
        newest = torch.where((Us[0][sl].abs() >= Us[1][sl].abs()), Ws[0][sl], Ws[1][sl])
        dhw0 = float((Hs[0][sl].float() - newest).abs().max())
        print("param %d %s shape %s  |H0-H1| %.3g  |W0-W1| %.3g  |H0 - W(owner)| %.3g  |W| %.3g  group %d" % (
            pi, getattr(p, "pname", "?"), tuple(p.shape), dh, dw, dhw0, float(newest.abs().max()), a.group_of[pi]), flush=True)
w.finalize()
