"""Overlay several runs (ref ``show/show_inforecs.py``): python show/show_inforecs.py a/inforec.pkl b/inforec.pkl"""
import sys

import numpy as np

from theanompi_b200.utils.recorder import Recorder

if __name__ == "__main__":
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:
        plt = None
    recs = []
    for p in sys.argv[1:]:
        r = Recorder(None, 40, p, False, device="cpu"); r.load(p); recs.append((p, r))
        print(p, r.summary())
    if plt and recs:
        fig, axs = plt.subplots(1, 2, figsize=(12, 4))
        for p, r in recs:
            v = np.array(r.info_dict["val_info"]) if r.info_dict["val_info"] else None
            a = np.array(r.info_dict["all_time"]) if r.info_dict["all_time"] else None
            if v is not None:
                axs[0].plot(v[:, 0], v[:, 3], label=p)
            if a is not None:
                axs[1].plot(a[:, 0], a[:, 1], label=p)
        axs[0].set_title("top-5 validation error"); axs[1].set_title("time per 5120 images (s)"); axs[0].legend()
        fig.savefig("inforecs.png")
