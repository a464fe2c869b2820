#!/usr/bin/env python
"""Host side of the parallel loader with REAL batch files (the part the synthetic bench skips): read 128x256x256x3 uint8 ``.npy``
batches into the staging ring from (a) a loader THREAD in the trainer's interpreter, (b) the loader PROCESS filling the shared-memory
ring — each once with an idle trainer thread and once while the trainer thread executes pure-Python work (holds the GIL the way an
eager training loop does).  Reports files/s and the slowdown of the trainer's Python loop.  No GPU needed.

    python scripts/loader_host_bench.py [--files 12] [--dir /tmp/tmpi_loader_bench]
"""
import argparse
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def busy_python(stop, counter):
    x = 0
    while not stop.is_set():
        for i in range(20000):
            x = (x * 1103515245 + 12345) & 0x7FFFFFFF
        counter[0] += 1


def run(mode, files, shape, busy):
    from theanompi_b200.models.data.proc_loader import ProcReader, read_batch_file
    stop, counter = threading.Event(), [0]
    if mode == "process":
        rd = ProcReader(shape, depth=2, pin=False)
        slots = [t.numpy() for t in rd.tensors]
        read = lambda f, k: rd.read(f, slots[k % 2])                     # noqa: E731
    else:
        rd = None
        slots = [np.empty(shape, dtype=np.uint8) for _ in range(2)]
        read = lambda f, k: read_batch_file(f, slots[k % 2])            # noqa: E731
    for k in range(2):                                                   # child start-up / page cache out of the timing
        read(files[k], k)
    done = [0.0]

    def loader():
        t0 = time.perf_counter()
        for k, f in enumerate(files):
            read(f, k)
        done[0] = time.perf_counter() - t0

    th = threading.Thread(target=loader)
    bt = threading.Thread(target=busy_python, args=(stop, counter)) if busy else None
    if bt:
        bt.start()
        time.sleep(0.2)
        counter[0] = 0
    t0 = time.perf_counter()
    th.start(); th.join()
    wall = time.perf_counter() - t0
    stop.set()
    if bt:
        bt.join()
    if rd is not None:
        rd.close()
    return len(files) / done[0], (counter[0] / wall if busy else None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=12)
    ap.add_argument("--dir", default="/tmp/tmpi_loader_bench")
    a = ap.parse_args()
    shape = (128, 256, 256, 3)
    os.makedirs(a.dir, exist_ok=True)
    files = []
    rs = np.random.RandomState(0)
    for i in range(a.files):
        f = os.path.join(a.dir, "%04d.npy" % i)
        if not os.path.exists(f):
            np.save(f, rs.randint(0, 256, shape, dtype=np.uint8))
        files.append(f)
    mb = np.prod(shape) / 1e6
    # the trainer loop's own speed with no loader at all
    stop, counter = threading.Event(), [0]
    bt = threading.Thread(target=busy_python, args=(stop, counter)); bt.start(); time.sleep(1.0); stop.set(); bt.join()
    base = counter[0] / 1.0
    print("batch file = %.1f MB; trainer-loop baseline %.0f iterations/s; %d cpus" % (mb, base, len(os.sched_getaffinity(0))))
    print("| loader | trainer thread | files/s | MB/s | trainer loop speed vs alone |")
    print("|---|---|---|---|---|")
    for mode in ("thread", "process"):
        for busy in (False, True):
            fps, loop = run(mode, files, shape, busy)
            print("| %s | %s | %.1f | %.0f | %s |" % (mode, "busy (pure Python)" if busy else "idle", fps, fps * mb,
                                                     ("%.0f %%" % (100.0 * loop / base)) if loop else "—"))


if __name__ == "__main__":
    main()
