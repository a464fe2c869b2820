"""Run a few eager AlexNet training steps; cudaProfilerStart/Stop brackets ONE steady-state step so
``ncu --profile-from-start off`` sees exactly one step's kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import importlib
from theanompi_b200.utils.recorder import Recorder

MODELS = {"alexnet": ("theanompi_b200.models.alex_net", "AlexNet", {}),
          "googlenet": ("theanompi_b200.models.googlenet", "GoogLeNet", dict(batch_size=32, file_batch_size=128)),
          "vgg16": ("theanompi_b200.models.lasagne_model_zoo.vgg16", "VGG16", dict(batch_size=32, file_batch_size=128))}
name = sys.argv[1] if len(sys.argv) > 1 else "alexnet"
modfile, cls, extra = MODELS[name]
torch.cuda.set_device(0)
cfg = dict(verbose=False, rank=0, size=1, device="cuda:0", cuda_graph=False,
           data_kwargs=dict(n_train_files=12, n_val_files=1, synthetic=True))
cfg.update(extra)
m = getattr(importlib.import_module(modfile), cls)(cfg); m.compile_iter_fns("avg")
rec = Recorder(None, 1000, cls, False, device="cuda:0")
for i in range(4):
    m.train_iter(i, rec)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.train_iter(5, rec)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
m.cleanup()
print("profiled one step")
