"""Offline viewer of one run's ``inforec.pkl`` (ref ``show/show_inforec.py``)."""
import sys

from theanompi_b200.utils.recorder import Recorder

if __name__ == "__main__":
    path = sys.argv[1] if len(sys.argv) > 1 else "./inforec/inforec.pkl"
    r = Recorder(None, 40, "run", True, device="cpu")
    r.load(path)
    print(r.show(label=path, show=False, save=path.replace(".pkl", ".png")))
