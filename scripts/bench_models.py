"""Device-timed training step of the other published benchmark models on ONE B200, same harness rules as
bench.py (CUDA events, warm-up, public model API ``train_iter`` through the loader), reported as the reference's metric
"seconds per 5120 images".  Random-init weights, synthetic ImageNet-shaped data.

    python scripts/bench_models.py [googlenet vgg16 resnet50 alexnet] [--steps 20]
"""
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

K80 = {"alexnet": 20.50, "googlenet": 63.89, "vgg16": 343.37, "resnet50": 163.15}      # 1-GPU published
MODELS = {
    "alexnet": ("theanompi_b200.models.alex_net", "AlexNet", dict(batch_size=128, file_batch_size=128)),
    "googlenet": ("theanompi_b200.models.googlenet", "GoogLeNet", dict(batch_size=32, file_batch_size=128)),
    "vgg16": ("theanompi_b200.models.lasagne_model_zoo.vgg16", "VGG16", dict(batch_size=32, file_batch_size=128)),
    "resnet50": ("theanompi_b200.models.lasagne_model_zoo.resnet50", "ResNet50", dict(batch_size=64, file_batch_size=64)),
}


def run(name, steps, warmup=5):
    from theanompi_b200.models import layers2
    from theanompi_b200.utils.recorder import Recorder
    modfile, cls, cfg = MODELS[name]
    layers2.reseed(); layers2.Dropout.layers.clear(); layers2.Crop.layers.clear()
    per_file = cfg["file_batch_size"] // cfg["batch_size"]
    n_files = (steps + warmup) // per_file + 3
    base = dict(verbose=False, rank=0, size=1, device="cuda:0", data_kwargs=dict(n_train_files=n_files, n_val_files=1, synthetic=True))
    base.update(cfg)
    m = getattr(importlib.import_module(modfile), cls)(base)
    m.compile_iter_fns("cdd" if name == "resnet50" else "avg")
    rec = Recorder(None, 10 ** 6, cls, False, device="cuda:0")
    c = 0
    for _ in range(warmup):
        m.train_iter(c, rec); c += 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        m.train_iter(c, rec); c += 1
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    bs = cfg["batch_size"]
    sec = ms / 1000.0 * 5120.0 / bs
    out = {"model": cls, "batch": bs, "ms_per_step": ms, "s_per_5120_images": sec, "images_per_s": bs / (ms / 1000.0),
           "k80_published_s": K80[name], "speedup_vs_k80": K80[name] / sec, "params": int(m.arena.numel) if getattr(m, "arena", None) is not None else None,
           "loss": float(rec.train_info["cost"][-1])}
    m.cleanup()
    del m
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 20
    names = [a for a in sys.argv[1:] if a in MODELS] or ["googlenet", "vgg16", "resnet50"]
    for n in names:
        try:
            print(json.dumps(run(n, steps)), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"model": n, "error": repr(e)[:300]}), flush=True)
