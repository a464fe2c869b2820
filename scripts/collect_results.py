#!/usr/bin/env python
"""Collect every bench.py JSON line that the GPU calls of this round left under gpurun_out/ into profiles/bench_r2.jsonl and
render profiles/bench_r2.md (one table per family).  Usage: python scripts/collect_results.py"""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")

# file-name pattern → GPU call it came from (chronological; later calls supersede earlier ones in the headline tables)
CALLS = ["c1", "c2", "c3", "c4", "c5", "c6", "diag2b", "mg2", "mg2b", "mg4", "mg8", "c7", "final"]


def load():
    rows = []
    for call in CALLS:
        for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", call + "_*.json")) +
                        glob.glob(os.path.join(ROOT, "gpurun_out", call + "_bench_*.log"))):
            line = None
            for l in open(f, errors="replace"):
                if l.startswith("{") and '"metric"' in l:
                    line = l
            if line is None:
                continue
            try:
                d = json.loads(line)
            except ValueError:
                continue
            d["_file"] = os.path.basename(f)
            d["_call"] = call
            rows.append(d)
    return rows


def model_of(d):
    m = d["config"].get("model", "?")
    return {"alexnet": "AlexNet", "googlenet": "GoogLeNet", "vgg16": "VGG16", "resnet50": "ResNet50", "wrn": "Wide_ResNet"}.get(m, m)


def fmt(x, nd=3):
    return "—" if x is None else ("%." + str(nd) + "f") % x


def main():
    rows = load()
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "bench_r2.jsonl"), "w") as f:
        for d in rows:
            f.write(json.dumps(d) + "\n")
    md = ["# Round-2 bench lines (every `bench.py` JSON line of this round's GPU calls)", "",
          "Produced by `scripts/collect_results.py` from the files the GPU calls wrote (`gpurun_out/<call>_*.json`, copied verbatim into",
          "`bench_r2.jsonl`).  All numbers: CUDA events on the launching stream, ≥ 3 repeats of the K-step region, median reported, max over",
          "ranks, synthetic data, 1965 MHz SM clock unless noted.  `call` = which GPU session (c1…c7 single GPU; mg2/mg2b/mg4/mg8 = 2/2/4/8 GPUs);",
          "later calls contain later code.  `s/5120` = seconds per 5120 images (the reference README's unit).", ""]
    md.append("| call | impl | model | rule | GPUs | dtype | strategy / variant | ms/step | spread % | e2e ms/step | s/5120 | K80 published ÷ ours | file |")
    md.append("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for d in rows:
        impl = "ours" if d.get("impl", "ours") == "ours" else ("torch_best" if "torch_best" in d["impl"] else d["impl"][:14])
        cfg = d["config"]
        var = cfg.get("exch_strategy", "")
        if impl == "ours" and cfg.get("rule", "BSP") == "BSP":
            var += " pm=%d" % int(bool(cfg.get("push_master"))) + ("" if cfg.get("overlap", True) else " no-overlap")
        tag = d["_file"].replace(".json", "").replace(".log", "")
        vb = d.get("vs_baseline")
        md.append("| %s | %s | %s | %s | %d | %s | %s | %s | %s | %s | %s | %s | `%s` |" % (
            d["_call"], impl, model_of(d), cfg.get("rule", "BSP"), d["n_gpus"], d.get("dtype", ""), var, fmt(d["ms_per_step"]),
            fmt(d.get("repeats", {}).get("spread_pct"), 2), fmt(d.get("e2e", {}).get("ms_per_step")), fmt(d["value"], 5),
            ("%.0f×" % (1.0 / vb)) if vb else "—", tag))
    md.append("")
    # protocol extras
    md.append("## EASGD / GOSGD protocol numbers")
    md.append("")
    for d in rows:
        if "easgd" in d:
            e = d["easgd"]
            md.append("* `%s` (%d GPUs): τ=%s, %s workers, %s — exchange alone %s µs (%s GB/s both directions summed, %s bytes over NVLink), "
                      "contended %s µs per exchange (%s GB/s per worker; center link %s GB/s each direction), %s exchanges served."
                      % (d["_file"], d["n_gpus"], e.get("tau"), e.get("workers"), e.get("lock"), fmt(e.get("exchange_us_alone"), 0),
                         fmt(e.get("exchange_GBps_alone"), 0), e.get("bytes_per_exchange_over_nvlink"), fmt(e.get("exchange_us_contended"), 0),
                         fmt(e.get("exchange_GBps_per_worker_contended"), 0), fmt(e.get("center_link_GBps_each_direction"), 0),
                         e.get("center_exchanges_served")))
        if "gosgd" in d:
            g = d["gosgd"]
            md.append("* `%s` (%d GPUs): p=%s, %s pushes (%s skipped: receiver busy), %s merges, Σ push-sum weights = %.8f, merge %s µs = %s GB/s per "
                      "rank (%s bytes over NVLink), %s pushes/s." % (d["_file"], d["n_gpus"], g.get("p"), g.get("pushes"), g.get("pushes_skipped_busy"),
                                                                     g.get("merges"), g.get("sum_push_sum_weights", 0.0), fmt(g.get("merge_us"), 1),
                                                                     fmt(g.get("merge_GBps_per_rank"), 0), g.get("bytes_per_merge_over_nvlink"),
                                                                     fmt(g.get("pushes_per_s"), 0)))
    md.append("")
    open(os.path.join(OUT, "bench_r2.md"), "w").write("\n".join(md) + "\n")
    print("%d lines -> profiles/bench_r2.{jsonl,md}" % len(rows))


if __name__ == "__main__":
    main()
