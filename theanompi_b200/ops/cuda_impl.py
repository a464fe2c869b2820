"""CUDA implementations of the functional ops: thin, checked wrappers that hand raw
device pointers of torch tensors to the hand-written sm_100a kernels in ``csrc/``.

GEMM-shaped work (FC forward/dgrad/wgrad; conv forward/dgrad/wgrad through an NHWC
im2col gather) runs on ONE kernel, ``gemm_bf16`` (TMA → smem → ``tcgen05.mma`` →
TMEM → fused epilogue, see ``csrc/gemm_tcgen05.cu``).  Operand-major flags make
transposed copies unnecessary:

    forward   y  = x · Wᵀ        A = x   (K-major)   B = W   (K-major)   + bias + ReLU → bf16
    dgrad     dx = dy · W        A = dy  (K-major)   B = W   (MN-major)               → bf16
    wgrad     dW = dyᵀ · x       A = dy  (MN-major)  B = x   (MN-major)  split-K      → fp32 (straight into the arena's G)

Everything raises if the extension is missing — there is no eager fallback on a GPU.
"""
from __future__ import annotations

import os

import torch

from . import native, precision

_STEP = {}          # device index -> int64[1] step counter used by the dropout Philox stream
BF16 = torch.bfloat16
F32 = torch.float32


def ADT():
    """Activation dtype of the active precision mode (bf16, or fp32 for the tf32 path)."""
    return precision.act_dtype()


def _is32(t):
    return t.dtype == F32


def _al(t):
    """Elements per 16 bytes (TMA / vector alignment unit): 8 for bf16, 4 for fp32."""
    return 16 // t.element_size()


def L():
    return native.require()


def _st(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def step_counter(device):
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _STEP:
        _STEP[idx] = torch.zeros(1, dtype=torch.int64, device=torch.device("cuda", idx))
    return _STEP[idx]


def advance_step(device):
    s = step_counter(device)
    L().advance_step(s.data_ptr(), _st(s))


def _bf(t):
    """Cast to the activation dtype of the active precision mode (no-op on the hot path: layers already produce it)."""
    d = ADT()
    return t if t.dtype == d else t.to(d)


def _rows8(t2d):
    """Return a 2-D tensor whose row pitch is a multiple of 16 bytes (TMA rule) — the tensor itself when it already is, else
    a zero-padded copy (rare, tiny shapes)."""
    assert t2d.dim() == 2
    al = _al(t2d)
    if t2d.stride(1) == 1 and t2d.stride(0) % al == 0 and t2d.data_ptr() % 16 == 0:
        return t2d, t2d.stride(0)
    R, C = t2d.shape
    ld = (C + al - 1) // al * al
    out = torch.zeros((R, ld), dtype=t2d.dtype, device=t2d.device)
    out[:, :C] = t2d
    return out, ld


# --------------------------------------------------------------------------- GEMM
def gemm(a, b, M, N, K, a_mn=False, b_mn=False, out=None, out_dtype=None, bias=None, bias_mode=0,
         relu=False, alpha=1.0, lda=None, ldb=None, ldc=None, bn=0, splitk=0):
    """``out[M,N] = alpha * op(a) @ op(b) (+bias)(ReLU)``; ``a``/``b`` are bf16 tensors whose
    storage is described by (major flag, leading dimension)."""
    dev = a.device
    tf32 = _is32(a)
    assert a.dtype == b.dtype, "GEMM operands must share a dtype"
    if out is None:
        out = torch.empty((M, N), dtype=(F32 if tf32 else (out_dtype or BF16)), device=dev)
    if ldc is None:
        ldc = out.stride(0) if out.dim() == 2 else N
    if bias is not None:
        assert bias.dtype == torch.float32
    L().gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), _p(bias), int(M), int(N), int(K), int(lda), int(ldb),
                  int(ldc), int(bool(a_mn)), int(bool(b_mn)), int(out.dtype == BF16), int(bias_mode), int(bool(relu)),
                  float(alpha), int(bn), int(splitk), _st(a), int(tf32))
    return out


# --------------------------------------------------------------------------- linear
def linear_bias_act(x, w, b, relu=True):
    x2 = _bf(x).contiguous()
    B_, I = x2.shape
    O = w.shape[0]
    xa, lda = _rows8(x2)
    wa, ldb = _rows8(_bf(w))
    bias = b.float() if b is not None and b.dtype != torch.float32 else b
    if FC_SPLITK and B_ <= 128 and I >= 1024 and O % 8 == 0 and _is32(x2):
        acc = gemm(xa, wa, B_, O, I, lda=lda, ldb=ldb)                    # fp32 split-K accumulation, finished in place
        L().bias_act_f32(acc.data_ptr(), _p(bias), acc.data_ptr(), int(B_), int(O), int(bool(relu)), _st(x2))
        return acc
    if FC_SPLITK and B_ <= 128 and I >= 1024 and O % 8 == 0:
        # small-batch FC forward is a weight stream: one m-tile, so the parallelism comes from n-tiles x split-K (fp32
        # reductions into a scratch tile), followed by a tiny bias + ReLU + bf16 pass.  (The fused-epilogue kernel needs
        # 32-wide tiles to fill the machine and then re-reads the activations 128 times through L2: 37 us vs 13 us for fc6.)
        acc = gemm(xa, wa, B_, O, I, out_dtype=torch.float32, lda=lda, ldb=ldb)
        y = torch.empty((B_, O), dtype=BF16, device=x2.device)
        L().bias_act_cast(acc.data_ptr(), _p(bias), y.data_ptr(), int(B_), int(O), int(bool(relu)), _st(x2))
        return y
    return gemm(xa, wa, B_, O, I, bias=bias, bias_mode=1 if b is not None else 0, relu=relu, lda=lda, ldb=ldb)


def _mask_and_bias_grad(dy, y, relu, db_out, R, C, ld, need_db=True):
    """dym = dy ⊙ (y > 0) (contiguous [R, C]) and db = Σ_rows dym in one pass."""
    dev = dy.device
    if not need_db and not relu and ld == C:
        return dy, None                                    # bias-free linear conv (a BatchNormal follows): nothing to do
    db = db_out if db_out is not None else torch.empty(C, dtype=torch.float32, device=dev)
    contiguous = (ld == C)
    if _is32(dy):
        if relu or not contiguous:
            dym = torch.empty((R, C), dtype=F32, device=dev)
            L().relu_bias_bwd2_f32(dy.data_ptr(), _p(y), dym.data_ptr(), db.data_ptr(), 0, int(C), int(R), int(C), int(ld), int(bool(relu)), _st(dy))
        else:
            dym = dy
            L().relu_bias_bwd2_f32(dy.data_ptr(), 0, 0, db.data_ptr(), 0, int(C), int(R), int(C), int(ld), 0, _st(dy))
        return dym, db
    if relu or not contiguous:
        dym = torch.empty((R, C), dtype=BF16, device=dev)
        L().relu_bias_bwd(dy.data_ptr(), _p(y), dym.data_ptr(), db.data_ptr(), int(R), int(C), int(ld), int(bool(relu)), _st(dy))
    else:
        dym = dy
        L().relu_bias_bwd(dy.data_ptr(), 0, 0, db.data_ptr(), int(R), int(C), int(ld), 0, _st(dy))
    return dym, db


def maxpool_relu_bias_bwd(dyp, arg, y, pool, db0, db1=None):
    """Backward of conv(+ReLU)→max-pool up to the conv's masked output gradient, in one kernel: scatter the pooled
    gradient through the argmax, apply the ReLU mask, accumulate the bias gradient(s).  Returns dym, shaped like y."""
    dyp = _bf(dyp).contiguous()
    N, H, W, C = y.shape
    Ho, Wo = dyp.shape[1], dyp.shape[2]
    k, s_, p_ = int(pool[0]), int(pool[1]), int(pool[2])
    dym = torch.empty((N, H, W, C), dtype=BF16, device=y.device)
    c_split = C if db1 is None else int(db0.numel())
    L().maxpool_relu_bias_bwd(dyp.data_ptr(), arg.data_ptr(), y.data_ptr(), dym.data_ptr(), db0.data_ptr(), _p(db1), c_split,
                              N, H, W, C, Ho, Wo, k, s_, p_, _st(y))
    return dym


def linear_bias_act_bwd(x, w, y, dy, relu, need_dx, dw_out=None, db_out=None):
    x2 = _bf(x).contiguous()
    dy = _bf(dy).contiguous()
    B_, I = x2.shape
    O = w.shape[0]
    al = _al(x2)
    if O % al or I % al:
        return _linear_bwd_padded(x2, w, y, dy, relu, need_dx, dw_out, db_out)
    dym, db = _mask_and_bias_grad(dy, y, relu, db_out.view(-1) if db_out is not None else None, B_, O, O)
    wb = _bf(w)
    dx = None
    if need_dx:
        dx = gemm(dym, wb, B_, I, O, a_mn=False, b_mn=True, lda=O, ldb=I)
    dw = dw_out if dw_out is not None else torch.empty((O, I), dtype=torch.float32, device=x.device)
    gemm(dym, x2, O, I, B_, a_mn=True, b_mn=True, out=dw, lda=O, ldb=I, ldc=I)
    return dx, dw, db


def _linear_bwd_padded(x2, w, y, dy, relu, need_dx, dw_out, db_out):
    """Shapes whose pitches violate the 16-byte TMA rule (e.g. a 10-class test head): pad to 8."""
    B_, I = x2.shape
    O = w.shape[0]
    Op, Ip = (O + 7) // 8 * 8, (I + 7) // 8 * 8
    dev = x2.device
    dt = x2.dtype
    dyf = dy.float()
    if relu:
        dyf = dyf * (y > 0)
    db = dyf.sum(0)
    dyp = torch.zeros((B_, Op), dtype=dt, device=dev); dyp[:, :O] = dyf
    xp = torch.zeros((B_, Ip), dtype=dt, device=dev); xp[:, :I] = x2
    wp = torch.zeros((Op, Ip), dtype=dt, device=dev); wp[:O, :I] = w
    dx = gemm(dyp, wp, B_, Ip, Op, b_mn=True, lda=Op, ldb=Ip)[:, :I].contiguous() if need_dx else None
    dwp = torch.empty((Op, Ip), dtype=torch.float32, device=dev)
    gemm(dyp, xp, Op, Ip, B_, a_mn=True, b_mn=True, out=dwp, lda=Op, ldb=Ip, ldc=Ip)
    dw = dwp[:O, :I]
    if dw_out is not None:
        dw_out.copy_(dw); dw = dw_out
    if db_out is not None:
        db_out.view(-1).copy_(db); db = db_out
    return dx, dw, db


# --------------------------------------------------------------------------- conv (NHWC, im2col + tcgen05 GEMM)
def _out_hw(H, W, KH, KW, s, p):
    return (H + 2 * p - KH) // s + 1, (W + 2 * p - KW) // s + 1


def _im2col(x, c_off, Cg, KH, KW, Ho, Wo, s, p):
    N, H, W, Ct = x.shape
    K = KH * KW * Cg
    al = _al(x)
    if KH == 1 and KW == 1 and s == 1 and p == 0 and c_off == 0 and Cg == Ct and Ct % al == 0:
        return x.view(N * H * W, Ct), Ct, K                     # 1x1 conv: the activation IS the matrix
    Kp = (K + 7) // 8 * 8
    col = torch.empty((N * Ho * Wo, Kp), dtype=x.dtype, device=x.device)
    fn = L().im2col_f32 if _is32(x) else L().im2col
    fn(x.data_ptr(), col.data_ptr(), N, H, W, Ct, int(c_off), int(Cg), KH, KW, Ho, Wo, int(s), int(p), Kp, _st(x))
    return col, Kp, K


def _w2d(w, K, Kp):
    """OHWI bf16 weights as the GEMM's [O, Kp] K-major operand (zero-padded when K % 8 != 0)."""
    O = w.shape[0]
    w2 = _bf(w).reshape(O, K)
    if Kp == K and w2.data_ptr() % 16 == 0:
        return w2
    wp = torch.empty((O, Kp), dtype=w2.dtype, device=w.device)
    (L().pad_rows_f32 if _is32(w2) else L().pad_rows)(w2.data_ptr(), wp.data_ptr(), O, K, K, Kp, _st(w))
    return wp


GROUP2_FUSED = os.environ.get("TMPI_GROUP2_FUSED", "1") != "0"   # both groups of a 2-group conv per launch
FC_SPLITK = os.environ.get("TMPI_FC_SPLITK", "1") != "0"      # small-batch FC forward: n-tiles x split-K + finishing pass
CONV_MODE = os.environ.get("TMPI_CONV", "implicit")      # implicit: TMA-im2col implicit GEMM; explicit: im2col matrix + GEMM


def _implicit_ok(x, w, c_off, Cg, Ot, o_off):
    """TMA im2col needs 16-byte aligned channel slices; C = 3 (first layer) stays on the explicit path."""
    Ct = x.shape[3]
    al = _al(x)
    return (CONV_MODE == "implicit" and Cg % al == 0 and c_off % al == 0 and Ct % al == 0 and Ot % al == 0 and o_off % al == 0
            and w.shape[0] % al == 0 and w.is_contiguous() and w.data_ptr() % 16 == 0)


def _conv_fwd_group(x, w, b, y, o_off, c_off, Cg, s, p, relu):
    N, H, W, Ct = x.shape
    Og, KH, KW, _ = w.shape
    Ho, Wo = y.shape[1], y.shape[2]
    Ot = y.shape[3]
    wb = _bf(w)
    if _implicit_ok(x, wb, c_off, Cg, Ot, o_off):
        # implicit GEMM: the activation tile is gathered by TMA im2col loads inside the kernel — no col matrix
        yp = y.data_ptr() + o_off * y.element_size()
        L().conv_fprop(x.data_ptr(), wb.data_ptr(), yp, _p(b), N, H, W, Ct, int(c_off), int(Cg), KH, KW, Ho, Wo, int(s), int(p),
                       Og, Ot, int(bool(relu)), int(not _is32(x)), 0, _st(x), int(_is32(x)))
        return None
    col, Kp, K = _im2col(x, c_off, Cg, KH, KW, Ho, Wo, s, p)
    w2 = _w2d(w, K, Kp)
    M = N * Ho * Wo
    yv = y.view(M, Ot)[:, o_off:o_off + Og]
    gemm(col, w2, M, Og, K, out=yv, bias=b, bias_mode=1 if b is not None else 0, relu=relu, lda=Kp, ldb=Kp, ldc=Ot)
    return (col, Kp, K)


def _s2d_geom(H, W, C, KH, KW, stride, pad):
    """Geometry of the space-to-depth rewrite of a strided few-channel conv, or None when it does not apply.  The zero padding
    of the original convolution is folded into the space-to-depth image (the kernel reads x[S*i + dy - pad, ...])."""
    if not (CONV_MODE == "implicit" and C < 8 and C % 4 != 0 and stride > 1):
        return None
    S = stride
    Hp, Wp = H + 2 * pad, W + 2 * pad
    Hs, Ws = -(-Hp // S), -(-Wp // S)
    KHs, KWs = -(-KH // S), -(-KW // S)
    Ho, Wo = _out_hw(H, W, KH, KW, stride, pad)
    if Hs - KHs + 1 != Ho or Ws - KWs + 1 != Wo:
        return None
    Cp = (S * S * C + 7) // 8 * 8
    return S, Hs, Ws, KHs, KWs, Cp, Ho, Wo, int(pad)


def _conv_s2d_fwd(x, w, b, relu, g):
    """First-layer conv (e.g. AlexNet 11x11/4 on RGB) as a stride-1 conv on the space-to-depth image (implicit GEMM)."""
    S, Hs, Ws, KHs, KWs, Cp, Ho, Wo, P0 = g
    N, H, W, C = x.shape
    O, KH, KW, _ = w.shape
    dev = x.device
    f32 = _is32(x)
    xs = torch.empty((N, Hs, Ws, Cp), dtype=x.dtype, device=dev)
    ws = torch.empty((O, KHs, KWs, Cp), dtype=x.dtype, device=dev)
    if f32:
        L().space_to_depth_f32(x.data_ptr(), xs.data_ptr(), N, H, W, C, S, Hs, Ws, Cp, P0, _st(x))
        L().s2d_filter_pack_f32(_bf(w).contiguous().data_ptr(), ws.data_ptr(), O, KH, KW, C, S, KHs, KWs, Cp, _st(x))
    else:
        L().space_to_depth(x.data_ptr(), xs.data_ptr(), N, H, W, C, S, Hs, Ws, Cp, P0, _st(x))
        L().s2d_filter(_bf(w).contiguous().data_ptr(), ws.data_ptr(), O, KH, KW, C, S, KHs, KWs, Cp, 0, _st(x))
    y = torch.empty((N, Ho, Wo, O), dtype=x.dtype, device=dev)
    L().conv_fprop(xs.data_ptr(), ws.data_ptr(), y.data_ptr(), _p(b), N, Hs, Ws, Cp, 0, Cp, KHs, KWs, Ho, Wo, 1, 0, O, O,
                   int(bool(relu)), int(not f32), 0, _st(x), int(f32))
    return y, xs


def _conv_s2d_bwd(xs, w, y, dy, relu, g, dw_out, db_out, pre_masked=False):
    S, Hs, Ws, KHs, KWs, Cp, Ho, Wo, P0 = g
    O, KH, KW, C = w.shape
    N = xs.shape[0]
    M = N * Ho * Wo
    dev = xs.device
    if pre_masked:
        dym, db = dy, db_out
    else:
        dym, db = _mask_and_bias_grad(dy.view(M, O), y.view(M, O), relu, db_out.view(-1) if db_out is not None else None, M, O, O)
    dws = torch.empty((O, KHs, KWs, Cp), dtype=torch.float32, device=dev)
    L().conv_wgrad(dym.data_ptr(), xs.data_ptr(), dws.data_ptr(), N, Hs, Ws, Cp, 0, Cp, KHs, KWs, Ho, Wo, 1, 0, O, O, _st(xs), int(_is32(xs)))
    dw = dw_out if dw_out is not None else torch.empty((O, KH, KW, C), dtype=torch.float32, device=dev)
    L().s2d_filter(dws.data_ptr(), dw.data_ptr(), O, KH, KW, C, S, KHs, KWs, Cp, 1, _st(xs))
    return dw, db


def conv2d_bias_act(x, w, b, stride=1, pad=0, groups=1, relu=True, return_cols=False):
    x = _bf(x).contiguous()
    N, H, W, C = x.shape
    O, KH, KW, Cg = w.shape
    assert Cg * groups == C
    g = _s2d_geom(H, W, C, KH, KW, stride, pad) if (groups == 1 and O % 8 == 0) else None
    if g is not None:
        y, xs = _conv_s2d_fwd(x, w, b, relu, g)
        return (y, [("s2d", xs, g)]) if return_cols else y
    Ho, Wo = _out_hw(H, W, KH, KW, stride, pad)
    y = torch.empty((N, Ho, Wo, O), dtype=x.dtype, device=x.device)
    Og = O // groups
    cols = []
    for g in range(groups):
        cols.append(_conv_fwd_group(x, w[g * Og:(g + 1) * Og], None if b is None else b[g * Og:(g + 1) * Og], y, g * Og,
                                    g * Cg, Cg, stride, pad, relu))
    return (y, cols) if return_cols else y


def conv2d_group2_bias_act(x, w0, b0, w1, b1, stride, pad, relu, return_cols=False):
    x = _bf(x).contiguous()
    N, H, W, C = x.shape
    Og, KH, KW, Cg = w0.shape
    Ho, Wo = _out_hw(H, W, KH, KW, stride, pad)
    y = torch.empty((N, Ho, Wo, 2 * Og), dtype=x.dtype, device=x.device)
    wb0, wb1 = _bf(w0), _bf(w1)
    es, f32 = x.element_size(), _is32(x)
    if GROUP2_FUSED and _implicit_ok(x, wb0, 0, Cg, 2 * Og, 0) and _implicit_ok(x, wb1, Cg, Cg, 2 * Og, Og) and (b0 is None) == (b1 is None):
        # both groups in ONE persistent launch: their tiles fill the 148 SMs together instead of two under-filled waves
        L().conv_fprop2(x.data_ptr(), wb0.data_ptr(), wb1.data_ptr(), y.data_ptr(), y.data_ptr() + Og * es, _p(b0), _p(b1), N, H, W, C, 0,
                        int(Cg), int(Cg), KH, KW, Ho, Wo, int(stride), int(pad), Og, 2 * Og, int(bool(relu)), int(not f32), 0, _st(x),
                        int(f32))
        return (y, [None, None]) if return_cols else y
    cols = [_conv_fwd_group(x, w0, b0, y, 0, 0, Cg, stride, pad, relu),
            _conv_fwd_group(x, w1, b1, y, Og, Cg, Cg, stride, pad, relu)]
    return (y, cols) if return_cols else y


def _conv_bwd_group(x, w, y, dy, dx, o_off, c_off, Cg, s, p, relu, need_dx, dw_out, db_out, col=None, pre_masked=False, need_db=True):
    N, H, W, Ct = x.shape
    Og, KH, KW, _ = w.shape
    Ho, Wo, Ot = y.shape[1], y.shape[2], y.shape[3]
    M = N * Ho * Wo
    dev = x.device
    dyv = dy.view(M, Ot)[:, o_off:o_off + Og]
    yv = y.view(M, Ot)[:, o_off:o_off + Og]
    if pre_masked:
        # dy already carries the ReLU mask and db is already accumulated (fused pool backward): the GEMMs read this
        # group's channel slice of the full tensor in place (row pitch Ot)
        dym, db, ldy, dy_coff = dyv, db_out, Ot, o_off
    else:
        dym, db = _mask_and_bias_grad(dyv, yv, relu, db_out.view(-1) if db_out is not None else None, M, Og, Ot, need_db)
        ldy, dy_coff = Og, 0
    wb = _bf(w)
    if col is None and _implicit_ok(x, wb, c_off, Cg, Ot, o_off):
        # ---- implicit GEMM backward: wgrad gathers im2col(x) by TMA; dgrad (stride 1) is a forward conv of dy with the
        # flipped / transposed filter, written straight into dx's channel slice.
        dw = dw_out if dw_out is not None else torch.empty((Og, KH, KW, Cg), dtype=torch.float32, device=dev)
        es, f32 = x.element_size(), _is32(x)
        L().conv_wgrad(dym.data_ptr(), x.data_ptr(), dw.data_ptr(), N, H, W, Ct, int(c_off), int(Cg), KH, KW, Ho, Wo, int(s), int(p),
                       Og, int(ldy), _st(x), int(f32))
        if need_dx:
            if s == 1:
                # dgrad = forward conv of dy with the mirrored, transposed filter — which the kernel's TMA loads read straight
                # out of the forward weights (MN-major boxes of the mirrored tap): no flipped copy
                L().conv_fprop(dym.data_ptr() - dy_coff * es, wb.data_ptr(), dx.data_ptr() + c_off * es, 0, N, Ho, Wo, int(ldy), int(dy_coff),
                               Og, KH, KW, H, W, 1, KH - 1 - int(p), int(Cg), Ct, 0, int(not f32), 1, _st(x), int(f32))
            else:
                colK = KH * KW * Cg
                Kp = (colK + 7) // 8 * 8
                dcol = gemm(dym, _w2d(w, colK, Kp), M, Kp, Og, b_mn=True, lda=ldy, ldb=Kp)
                (L().col2im_f32 if f32 else L().col2im)(dcol.data_ptr(), dx.data_ptr(), N, H, W, Ct, int(c_off), int(Cg), KH, KW, Ho, Wo,
                                                        int(s), int(p), Kp, _st(x))
        return dw, db
    col, Kp, K = col if col is not None else _im2col(x, c_off, Cg, KH, KW, Ho, Wo, s, p)   # forward's matrix is reused
    dw = dw_out if dw_out is not None else torch.empty((Og, KH, KW, Cg), dtype=torch.float32, device=dev)
    # wgrad: dW[Og, K] = dymᵀ[Og, M] · col[M, K]   (both operands MN-major, split-K over M)
    gemm(dym, col, Og, K, M, a_mn=True, b_mn=True, out=dw.view(Og, K), lda=ldy, ldb=Kp, ldc=K)
    if need_dx:
        w2 = _w2d(w, K, Kp)
        one_by_one = (KH == 1 and KW == 1 and s == 1 and p == 0 and c_off == 0 and Cg == Ct)
        if one_by_one:
            gemm(dym, w2, M, Kp, Og, b_mn=True, out=dx.view(M, Ct), lda=ldy, ldb=Kp, ldc=Ct)
        else:
            dcol = gemm(dym, w2, M, Kp, Og, b_mn=True, lda=ldy, ldb=Kp)
            (L().col2im_f32 if _is32(x) else L().col2im)(dcol.data_ptr(), dx.data_ptr(), N, H, W, Ct, int(c_off), int(Cg), KH, KW, Ho, Wo,
                                                         int(s), int(p), Kp, _st(x))
    return dw, db


def conv2d_bias_act_bwd(x, w, y, dy, stride, pad, groups, relu, need_dx, dw_out=None, db_out=None, cols=None, pre_masked=False,
                        need_db=True):
    x = _bf(x).contiguous()
    dy = _bf(dy).contiguous()
    N, H, W, C = x.shape
    O, KH, KW, Cg = w.shape
    if cols and isinstance(cols[0], tuple) and len(cols[0]) == 3 and isinstance(cols[0][0], str) and cols[0][0] == "s2d":
        if need_dx:
            raise RuntimeError("space-to-depth conv path is for the first layer only (no input gradient)")
        dw, db = _conv_s2d_bwd(cols[0][1], w, y, dy, relu, cols[0][2], dw_out, db_out, pre_masked)
        return None, dw, db
    if (Cg % _al(x) or O % _al(x)) and need_dx:
        raise RuntimeError("conv dgrad needs channel counts that are multiples of 16 bytes")
    dx = torch.empty_like(x) if need_dx else None
    Og = O // groups
    if groups == 1:
        dw, db = _conv_bwd_group(x, w, y, dy, dx, 0, 0, Cg, stride, pad, relu, need_dx, dw_out, db_out,
                                 col=cols[0] if cols else None, pre_masked=pre_masked, need_db=need_db)
        return dx, dw, db
    dw = dw_out if dw_out is not None else torch.empty(tuple(w.shape), dtype=torch.float32, device=x.device)
    db = db_out if db_out is not None else torch.empty(O, dtype=torch.float32, device=x.device)
    for g in range(groups):
        _conv_bwd_group(x, w[g * Og:(g + 1) * Og], y, dy, dx, g * Og, g * Cg, Cg, stride, pad, relu, need_dx,
                        dw[g * Og:(g + 1) * Og], db[g * Og:(g + 1) * Og], col=cols[g] if cols else None, pre_masked=pre_masked)
    return dx, dw, db


def conv2d_group2_bias_act_bwd(x, w0, w1, y, dy, stride, pad, relu, need_dx, outs=(None, None, None, None), cols=None,
                               pre_masked=False):
    x = _bf(x).contiguous()
    dy = _bf(dy).contiguous()
    Og, KH, KW, Cg = w0.shape
    dx = torch.empty_like(x) if need_dx else None
    wb0, wb1 = _bf(w0), _bf(w1)
    Ot = 2 * Og
    no_cols = not cols or (cols[0] is None and cols[1] is None)
    if (GROUP2_FUSED and no_cols and _implicit_ok(x, wb0, 0, Cg, Ot, 0) and _implicit_ok(x, wb1, Cg, Cg, Ot, Og)
            and (not need_dx or stride == 1)):
        # ---- both groups per launch: one mask/bias pass over the full width, one wgrad launch, one dgrad launch
        N, H, W, Ct = x.shape
        Ho, Wo = y.shape[1], y.shape[2]
        M = N * Ho * Wo
        dev = x.device
        db0 = outs[1] if outs[1] is not None else torch.empty(Og, dtype=torch.float32, device=dev)
        db1 = outs[3] if outs[3] is not None else torch.empty(Og, dtype=torch.float32, device=dev)
        es, f32 = x.element_size(), _is32(x)
        if pre_masked:
            dym = dy
        else:
            dym = torch.empty((M, Ot), dtype=x.dtype, device=dev)
            (L().relu_bias_bwd2_f32 if f32 else L().relu_bias_bwd2)(dy.data_ptr(), y.data_ptr(), dym.data_ptr(), db0.data_ptr(), db1.data_ptr(),
                                                                     int(Og), int(M), int(Ot), int(Ot), int(bool(relu)), _st(x))
        dw0 = outs[0] if outs[0] is not None else torch.empty((Og, KH, KW, Cg), dtype=torch.float32, device=dev)
        dw1 = outs[2] if outs[2] is not None else torch.empty((Og, KH, KW, Cg), dtype=torch.float32, device=dev)
        L().conv_wgrad2(dym.data_ptr(), dym.data_ptr() + Og * es, x.data_ptr(), dw0.data_ptr(), dw1.data_ptr(), N, H, W, Ct, 0, int(Cg), int(Cg),
                        KH, KW, Ho, Wo, int(stride), int(pad), Og, int(Ot), _st(x), int(f32))
        if need_dx:
            L().conv_fprop2(dym.data_ptr(), wb0.data_ptr(), wb1.data_ptr(), dx.data_ptr(), dx.data_ptr() + Cg * es, 0, 0, N, Ho, Wo, int(Ot), 0,
                            int(Og), int(Og), KH, KW, H, W, 1, KH - 1 - int(pad), int(Cg), Ct, 0, int(not f32), 1, _st(x), int(f32))
        return dx, (dw0, db0, dw1, db1)
    dw0, db0 = _conv_bwd_group(x, w0, y, dy, dx, 0, 0, Cg, stride, pad, relu, need_dx, outs[0], outs[1],
                               col=cols[0] if cols else None, pre_masked=pre_masked)
    dw1, db1 = _conv_bwd_group(x, w1, y, dy, dx, Og, Cg, Cg, stride, pad, relu, need_dx, outs[2], outs[3],
                               col=cols[1] if cols else None, pre_masked=pre_masked)
    return dx, (dw0, db0, dw1, db1)


# --------------------------------------------------------------------------- pool / LRN / dropout / loss
def pool2d_fwd(x, ksize, stride, pad, mode):
    x = _bf(x).contiguous()
    N, H, W, C = x.shape
    Ho, Wo = _out_hw(H, W, ksize, ksize, stride, pad)
    y = torch.empty((N, Ho, Wo, C), dtype=x.dtype, device=x.device)
    is_max = mode == "max"
    arg = torch.empty((N, Ho, Wo, C), dtype=torch.uint8, device=x.device) if is_max else None
    (L().pool_fwd_f32 if _is32(x) else L().pool_fwd)(x.data_ptr(), y.data_ptr(), _p(arg), N, H, W, C, Ho, Wo, int(ksize), int(stride),
                                                     int(pad), int(is_max), _st(x))
    return y, arg


def pool2d_bwd_arg(dy, arg, xshape, ksize, stride, pad, mode):
    dy = _bf(dy).contiguous()
    N, H, W, C = xshape
    Ho, Wo = dy.shape[1], dy.shape[2]
    dx = torch.empty(tuple(xshape), dtype=dy.dtype, device=dy.device)
    (L().pool_bwd_f32 if _is32(dy) else L().pool_bwd)(dy.data_ptr(), _p(arg), dx.data_ptr(), N, H, W, C, Ho, Wo, int(ksize), int(stride),
                                                      int(pad), int(mode == "max"), _st(dy))
    return dx


def lrn(x, n=5, k=2.0, alpha=1e-4, beta=0.75):
    x = _bf(x).contiguous()
    C = x.shape[-1]
    y = torch.empty_like(x)
    (L().lrn_fwd_f32 if _is32(x) else L().lrn_fwd)(x.data_ptr(), y.data_ptr(), x.numel() // C, C, int(n), float(k), float(alpha),
                                                   float(beta), _st(x))
    return y, None


def lrn_bwd(x, dy, n=5, k=2.0, alpha=1e-4, beta=0.75):
    x = _bf(x).contiguous()
    dy = _bf(dy).contiguous()
    C = x.shape[-1]
    dx = torch.empty_like(x)
    (L().lrn_bwd_f32 if _is32(x) else L().lrn_bwd)(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), x.numel() // C, C, int(n), float(k),
                                                   float(alpha), float(beta), _st(x))
    return dx


def dropout_fwd(x, p_drop, layer_id):
    from .functional import rng_state
    x = _bf(x).contiguous()
    y = torch.empty_like(x)
    mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    step = step_counter(x.device)
    (L().dropout_fwd_f32 if _is32(x) else L().dropout_fwd)(x.data_ptr(), y.data_ptr(), mask.data_ptr(), x.numel(), float(p_drop),
                                                           int(rng_state()["seed"]), int(layer_id), step.data_ptr(), _st(x))
    return y, mask


def dropout_bwd(dy, mask):
    dy = _bf(dy).contiguous()
    dx = torch.empty_like(dy)
    (L().dropout_bwd_f32 if _is32(dy) else L().dropout_bwd)(dy.data_ptr(), mask.data_ptr(), dx.data_ptr(), dy.numel(), _st(dy))
    return dx


def softmax_xent(logits, labels, weight=1.0):
    lg = _bf(logits).contiguous()
    B_, C = lg.shape
    labels = labels.contiguous()
    assert labels.dtype == torch.int64
    dl = torch.empty_like(lg)
    rowstat = torch.empty((B_, 3), dtype=torch.float32, device=lg.device)
    out3 = torch.empty(3, dtype=torch.float32, device=lg.device)
    (L().softmax_xent_f32 if _is32(lg) else L().softmax_xent)(lg.data_ptr(), labels.data_ptr(), dl.data_ptr(), rowstat.data_ptr(),
                                                              out3.data_ptr(), B_, C, float(weight), _st(lg))
    return out3[0], out3[1], out3[2], dl


# --------------------------------------------------------------------------- batch norm (+ residual)(+ ReLU), residual add
def batch_norm_fwd(x, gamma, beta, run_mean, run_var, training, momentum, eps, relu, res=None):
    x = _bf(x).contiguous()
    C = x.shape[-1]
    R = x.numel() // C
    dev = x.device
    y = torch.empty_like(x)
    mean = torch.empty(C, dtype=F32, device=dev)
    rstd = torch.empty(C, dtype=F32, device=dev)
    scratch = torch.empty(2 * C, dtype=F32, device=dev)
    if res is not None:
        res = _bf(res).contiguous()
        assert res.shape == x.shape
    assert gamma.dtype == F32 and beta.dtype == F32
    if not training:
        assert run_mean is not None and run_var is not None
    L().bn_forward(x.data_ptr(), _p(res), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                   _p(run_mean), _p(run_var), scratch.data_ptr(), int(R), int(C), float(momentum), float(eps), int(bool(training)),
                   int(bool(relu)), int(_is32(x)), _st(x))
    return y, mean, rstd


def batch_norm_bwd(x, dy, y, gamma, mean, rstd, relu, need_dres, dgamma_out=None, dbeta_out=None):
    x = _bf(x).contiguous()
    dy = _bf(dy).contiguous()
    C = x.shape[-1]
    R = x.numel() // C
    dev = x.device
    dx = torch.empty_like(x)
    # without a ReLU the residual branch receives dy itself: nothing to write
    dres = torch.empty_like(x) if (need_dres and relu) else None
    dgamma = dgamma_out.view(-1) if dgamma_out is not None else torch.empty(C, dtype=F32, device=dev)
    dbeta = dbeta_out.view(-1) if dbeta_out is not None else torch.empty(C, dtype=F32, device=dev)
    scratch = torch.empty(3 * C, dtype=F32, device=dev)
    L().bn_backward(x.data_ptr(), dy.data_ptr(), _p(y) if relu else 0, dx.data_ptr(), _p(dres), gamma.data_ptr(), mean.data_ptr(),
                    rstd.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), scratch.data_ptr(), int(R), int(C), int(bool(relu)),
                    int(_is32(x)), _st(x))
    if need_dres and not relu:
        dres = dy
    return dx, dres, dgamma, dbeta


def add(a, b):
    a = _bf(a).contiguous()
    b = _bf(b).contiguous()
    assert a.shape == b.shape
    y = torch.empty_like(a)
    L().add_tensors(a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), int(_is32(a)), _st(a))
    return y


# --------------------------------------------------------------------------- loader kernel
def crop_mirror_normalize(x, mean, std_scale, crop_hw, offsets, flips, out_dtype=None, out=None, c_out=None):
    out_dtype = out_dtype or ADT()
    x = x.contiguous()
    N, H, W, C = x.shape
    ch, cw = crop_hw
    kind = {torch.uint8: 0, torch.bfloat16: 1, torch.float32: 2}[x.dtype]
    mean = mean.float().contiguous()
    mode = 0 if mean.numel() == 1 else (1 if mean.numel() == C else 2)
    if mode == 2:
        assert mean.numel() == H * W * C
    Cout = c_out or C
    if out is None:
        out = torch.empty((N, ch, cw, Cout), dtype=out_dtype, device=x.device)
    assert out.dtype in (BF16, torch.float32)
    offsets = offsets.to(torch.int32).contiguous()
    flips = flips.to(torch.uint8).contiguous()
    if isinstance(std_scale, torch.Tensor):                   # per-channel scale (1 / 255 / img_std)
        cs = std_scale.to(device=x.device, dtype=torch.float32).contiguous()
        assert cs.numel() == C
        sc, cs_ptr = 1.0, cs.data_ptr()
    else:
        sc, cs_ptr = float(std_scale), 0
    L().crop_mirror_norm(x.data_ptr(), kind, mean.data_ptr(), mode, sc, cs_ptr, out.data_ptr(), int(out.dtype == BF16),
                         offsets.data_ptr(), flips.data_ptr(), N, H, W, C, ch, cw, Cout, _st(x))
    return out


# --------------------------------------------------------------------------- optimizer
def _table(arena):
    if not hasattr(arena, "_tab_cache"):
        arena._tab_cache = (arena.group_lr_mult_np.tolist(), arena.group_wd_np.tolist(),
                            [int(v) for v in arena.group_exch_np.tolist()])
    return arena._tab_cache


def sgd_flat(arena, g, lr, mu, nesterov, inv_k, lo, hi, only_local=False, only_exchanged=False):
    """Fused momentum-SGD over arena elements [lo, hi).  ``lr`` is read from
    ``arena.hyper[0]`` on the device (so a captured CUDA graph follows lr changes)."""
    lrm, wd, ex = _table(arena)
    filt = 1 if only_local else (2 if only_exchanged else 0)
    H = arena.H
    L().sgd_flat(arena.W.data_ptr(), g.data_ptr(), arena.U.data_ptr(), _p(H), arena.block_group.data_ptr(), lrm, wd, ex,
                 arena.hyper.data_ptr(), float(mu), int(bool(nesterov)), float(inv_k), int(lo), int(hi), filt, _st(arena.W))
