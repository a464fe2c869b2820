"""Run a few eager AlexNet training steps; cudaProfilerStart/Stop brackets ONE steady-state step so
``ncu --profile-from-start off`` sees exactly one step's kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theanompi_b200.models.alex_net import AlexNet
from theanompi_b200.utils.recorder import Recorder

torch.cuda.set_device(0)
cfg = dict(verbose=False, rank=0, size=1, device="cuda:0", cuda_graph=False,
           data_kwargs=dict(n_train_files=12, n_val_files=1, synthetic=True))
m = AlexNet(cfg); m.compile_iter_fns("avg")
rec = Recorder(None, 1000, "AlexNet", False, device="cuda:0")
for i in range(4):
    m.train_iter(i, rec)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.train_iter(5, rec)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
m.cleanup()
print("profiled one step")
