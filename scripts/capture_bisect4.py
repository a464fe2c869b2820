import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theanompi_b200.models.alex_net import AlexNet
from theanompi_b200.ops import native
torch.cuda.set_device(0)
cfg = dict(verbose=False, rank=0, size=1, device="cuda:0", batch_size=32, file_batch_size=32, no_paraload=True,
           data_kwargs=dict(n_train_files=4, n_val_files=1, synthetic=True))
m = AlexNet(cfg); m.compile_iter_fns("avg")
L = m.layers
Lib = native.require()
S = {}
def status(tag):
    st = S["s"].cuda_stream
    print("   [%s] capture_status(err,status)=%s" % (tag, Lib.capture_status(st)), flush=True)

def f(hooks):
    x = m.x_in
    for i, l in enumerate(L):
        x = l.forward(x)
        if hooks and x.requires_grad:
            x.register_hook(lambda g, i=i: status("grad arrives at output of layer %d %s" % (i, L[i].name.strip())))
    if hooks: status("after forward")
    loss = x.float().sum()
    loss.backward()
    if hooks: status("after backward")

def attempt(name, side_warm):
    print("==== %s" % name, flush=True)
    s = torch.cuda.Stream(); S["s"] = s
    if side_warm:
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3): f(False)
        torch.cuda.current_stream().wait_stream(s)
    else:
        for _ in range(3): f(False)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
                try:
                    f(True)
                except BaseException as e:
                    import traceback; traceback.print_exc(); raise
                status("before capture_end")
        g.replay(); torch.cuda.synchronize()
        print("OK", flush=True)
    except Exception as e:
        print("FAIL", str(e).split("\n")[0][:200], flush=True)
        try: torch.cuda.synchronize()
        except Exception: pass

attempt("warm-up on default stream", False)
