#!/usr/bin/env python
"""NVLink byte counters: difference of two `nvidia-smi nvlink -gt d` dumps taken around a bench run, per GPU and per training
step, next to the algorithmic bytes of the fused exchange.

    python scripts/nvlink_delta.py before.txt after.txt --steps 210 --world 8 --numel 60965224
"""
import argparse
import re


def parse(path):
    gpus, cur = {}, None
    for l in open(path):
        m = re.match(r"GPU (\d+):", l)
        if m:
            cur = int(m.group(1)); gpus[cur] = {"tx": 0, "rx": 0}
            continue
        m = re.search(r"Data (Tx|Rx): (\d+) KiB", l)
        if m and cur is not None:
            gpus[cur][m.group(1).lower()] += int(m.group(2))
    return gpus


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("before"); ap.add_argument("after")
    ap.add_argument("--steps", type=int, required=True, help="training steps executed between the two dumps")
    ap.add_argument("--world", type=int, required=True)
    ap.add_argument("--numel", type=int, required=True, help="exchanged fp32 elements per step")
    ap.add_argument("--push-bytes", type=float, default=2.0, help="bytes per element of the all-gather half: 2 = owner-keeps-master "
                    "(bf16 shadow only), 6 = fp32 master + bf16 shadow (TMPI_PUSH_MASTER=1)")
    a = ap.parse_args()
    b, c = parse(a.before), parse(a.after)
    n, w = float(a.numel), a.world
    MiB = 2.0 ** 20
    # NVLS path (multimem.ld_reduce + multimem.st), per step per GPU:
    #   reduce-scatter half: the switch pulls every member's copy of every slice (tx 4n — own slice included, the multicast
    #   address does not special-case the issuer) and returns the reduced 1/w slice (rx 4n/w);
    #   all-gather half: the owner stores its updated slice ONCE to the multicast address (tx push·n/w), the switch replicates
    #   it to every member, issuer included (rx push·n).
    tx = 4 * n + a.push_bytes * n / w
    rx = 4 * n / w + a.push_bytes * n
    print("algorithmic (NVLS two-shot, %g B/element pushed), per step per GPU: tx %.1f MiB, rx %.1f MiB" % (a.push_bytes, tx / MiB, rx / MiB))
    # P2P two-shot for comparison: rx = 4n(w-1)/w + push·n(w-1)/w, tx the same
    p2p = (4 * n + a.push_bytes * n) * (w - 1) / w
    print("algorithmic (P2P two-shot, no multicast): tx = rx = %.1f MiB" % (p2p / MiB))
    print("| GPU | tx MiB/step | rx MiB/step |")
    print("|---|---|---|")
    for g in sorted(c):
        tx = (c[g]["tx"] - b[g]["tx"]) / 1024.0 / a.steps
        rx = (c[g]["rx"] - b[g]["rx"]) / 1024.0 / a.steps
        print("| %d | %.1f | %.1f |" % (g, tx, rx))


if __name__ == "__main__":
    main()
