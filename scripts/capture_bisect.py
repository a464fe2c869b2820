"""Find which op breaks CUDA-graph capture: capture small pieces one by one."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theanompi_b200 import ops
from theanompi_b200.ops import cuda_impl as ci

torch.cuda.set_device(0)
dev = "cuda"

def try_capture(name, fn, mode="thread_local", warm=2):
    try:
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s, capture_error_mode=mode):
                out = fn()
        torch.cuda.current_stream().wait_stream(s)
        g.replay(); torch.cuda.synchronize()
        print("OK   %-28s [%s]" % (name, mode), flush=True)
        return True
    except Exception as e:
        print("FAIL %-28s [%s] %s" % (name, mode, str(e).split("\n")[0][:150]), flush=True)
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
        return False

A = torch.randn(256, 512, device=dev).bfloat16(); B = torch.randn(384, 512, device=dev).bfloat16()
try_capture("gemm", lambda: ci.gemm(A, B, 256, 384, 512, lda=512, ldb=512))
At = torch.randn(8192, 96, device=dev).bfloat16(); Bt = torch.randn(8192, 368, device=dev).bfloat16()
out = torch.zeros(96, 363, device=dev)
try_capture("gemm splitk (memset2D)", lambda: ci.gemm(At, Bt, 96, 363, 8192, a_mn=True, b_mn=True, out=out, lda=96, ldb=368, ldc=363))
dy = torch.randn(1024, 256, device=dev).bfloat16(); y = torch.randn(1024, 256, device=dev).bfloat16(); db = torch.zeros(256, device=dev)
try_capture("relu_bias_bwd (memset)", lambda: ci._mask_and_bias_grad(dy, y, True, db, 1024, 256, 256))
x = torch.randn(4, 27, 27, 96, device=dev).bfloat16()
try_capture("lrn", lambda: ci.lrn(x))
try_capture("pool", lambda: ci.pool2d_fwd(x, 3, 2, 0, "max"))
try_capture("dropout", lambda: ci.dropout_fwd(x, 0.5, 0))
lg = torch.randn(128, 1000, device=dev).bfloat16(); lab = torch.randint(0, 1000, (128,), device=dev)
try_capture("softmax", lambda: ci.softmax_xent(lg, lab))
w = torch.randn(64, 3, 3, 96, device=dev).bfloat16(); b = torch.zeros(64, device=dev)
try_capture("conv fwd", lambda: ci.conv2d_bias_act(x, w, b, 1, 1, 1, True))
try_capture("advance_step", lambda: ci.advance_step(torch.device("cuda:0")))

# autograd pieces
xl = torch.randn(128, 512, device=dev).bfloat16().requires_grad_(True)
wl = torch.randn(256, 512, device=dev).bfloat16().requires_grad_(True); bl = torch.zeros(256, device=dev, requires_grad=True)
def lin_fb():
    y = ops.linear_bias_act(xl, wl, bl, True)
    y.float().sum().backward()
for mode in ("thread_local", "global", "relaxed"):
    try_capture("linear fwd+bwd (autograd)", lin_fb, mode)
def torch_fb():
    y = torch.nn.functional.linear(xl.float(), wl.float())
    y.sum().backward()
for mode in ("thread_local", "global"):
    try_capture("pure torch fwd+bwd", torch_fb, mode)

# whole model
from theanompi_b200.models.alex_net import AlexNet
for mode in ("global", "relaxed", "thread_local"):
    cfg = dict(verbose=False, rank=0, size=1, device="cuda:0", batch_size=32, file_batch_size=32, no_paraload=True,
               data_kwargs=dict(n_train_files=4, n_val_files=1, synthetic=True))
    m = AlexNet(cfg); m.compile_iter_fns("avg")
    try_capture("alexnet fwd only", lambda: m.loss(m.x_in, m.y_in), mode)
    try_capture("alexnet step body", m._step_body, mode)
