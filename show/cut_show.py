"""Truncate a run's ``inforec.pkl`` to a given epoch and show / re-save it
(ref ``test/test-recorder/cut_show.py``: used after resuming from a snapshot so the curves of the
abandoned epochs do not stay in the record).

    python show/cut_show.py ./inforec/inforec.pkl 12 [--save]
"""
import os
import pickle
import sys

from theanompi_b200.utils.recorder import Recorder

if __name__ == "__main__":
    if len(sys.argv) < 3:
        sys.exit(__doc__)
    path, load_epoch = sys.argv[1], int(sys.argv[2])
    r = Recorder(None, 40, "run", True, device="cpu")
    r.load(path)
    r.cut(load_epoch)
    print(r.show(label="%s[:%d]" % (path, load_epoch), show=False, save=path.replace(".pkl", "_cut%d.png" % load_epoch)))
    if "--save" in sys.argv:
        out = os.path.join(os.path.dirname(path) or ".", "inforec_cut%d.pkl" % load_epoch)
        with open(out, "wb") as f:
            pickle.dump(r.info_dict, f, protocol=pickle.HIGHEST_PROTOCOL)
        print("wrote", out)
