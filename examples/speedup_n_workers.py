#!/usr/bin/env python
"""Speed-up vs number of workers — the counterpart of the reference's ``examples/speedup-n_workers.ipynb``.

Reads the round's measured bench lines (``profiles/bench_r2.jsonl``; produce new ones with
``python -m torch.distributed.run --nproc-per-node N … bench.py --gpus N --model M``) and prints, per model, the training
(+communication) time per 5120 images at 1/2/4/8 GPUs, the speed-up and the projected ImageNet epoch time with the reference's
formula ``t · 250.2 / 3600`` hours (1,281,167 images / 5120 = 250.2; ``speedup-n_workers.ipynb``, cell 1), next to the published
K80 numbers.

    python examples/speedup_n_workers.py [--dtype bf16] [--jsonl profiles/bench_r2.jsonl]
"""
import argparse
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K80 = {   # README.md:112-118 of the reference (train + comm seconds per 5120 images)
    "AlexNet": {1: 20.50, 2: 11.13, 4: 5.67, 8: 3.24},
    "GoogLeNet": {1: 63.89, 2: 32.40, 4: 16.22, 8: 8.49},
    "VGG16": {1: 343.37, 2: 176.26, 4: 91.77, 8: 48.70},
    "ResNet50": {1: 163.15, 2: 80.90, 4: 40.81, 8: 20.69},
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jsonl", default=os.path.join(ROOT, "profiles", "bench_r2.jsonl"))
    ap.add_argument("--dtype", default="bf16")
    a = ap.parse_args()
    best = {}
    for l in open(a.jsonl):
        d = json.loads(l)
        cfg = d["config"]
        if d.get("impl", "ours") != "ours" or cfg.get("rule", "BSP") != "BSP" or d.get("dtype") != a.dtype:
            continue
        if cfg.get("exch_strategy") not in ("fused", "local fused SGD") or not cfg.get("overlap", True):
            continue
        key = (cfg["model"], d["n_gpus"])
        if key not in best or d["value"] < best[key]["value"]:
            best[key] = d
    print("| model | GPUs | s / 5120 images | speed-up | efficiency | ImageNet epoch (h) | K80 published s / 5120 (speed-up) |")
    print("|---|---|---|---|---|---|---|")
    for model in ("AlexNet", "GoogLeNet", "VGG16", "ResNet50"):
        t1 = best.get((model, 1))
        for n in (1, 2, 4, 8):
            d = best.get((model, n))
            if d is None or t1 is None:
                continue
            t = d["value"]
            k = K80[model]
            print("| %s | %d | %.4f | %.2f× | %.0f %% | %.4f | %.2f (%.2f×) |" % (
                model, n, t, t1["value"] / t, 100.0 * t1["value"] / t / n, t * 250.2 / 3600.0, k[n], k[1] / k[n]))


if __name__ == "__main__":
    main()
