import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scripts.capture_bisect_util import try_capture
from theanompi_b200.ops import cuda_impl as ci
torch.cuda.set_device(0)
dev = "cuda"
for O in (1000, 1024, 4096):
    B_, I = 32, 4096
    x = torch.randn(B_, I, device=dev).bfloat16(); w = torch.randn(O, I, device=dev).bfloat16()
    y = torch.randn(B_, O, device=dev).bfloat16(); dy = torch.randn(B_, O, device=dev).bfloat16()
    db = torch.zeros(O, device=dev); dw = torch.zeros(O, I, device=dev)
    for relu in (True, False):
        try_capture("maskbias O=%d relu=%s" % (O, relu), lambda: ci._mask_and_bias_grad(dy, y, relu, db, B_, O, O))
        try_capture("dgrad gemm O=%d" % O, lambda: ci.gemm(dy, w, B_, I, O, b_mn=True, lda=O, ldb=I))
        try_capture("wgrad gemm O=%d" % O, lambda: ci.gemm(dy, x, O, I, B_, a_mn=True, b_mn=True, out=dw, lda=O, ldb=I, ldc=I))
        try_capture("linear bwd O=%d relu=%s" % (O, relu), lambda: ci.linear_bias_act_bwd(x, w, y, dy, relu, True, dw, db))
