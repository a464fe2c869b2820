"""Autograd-aware functional ops.

Each op dispatches on the device of its input:

* CUDA  → :mod:`theanompi_b200.ops.cuda_impl` (hand-written sm_100a kernels from
  ``csrc/``; hard error when the extension is missing),
* CPU   → :mod:`theanompi_b200.ops.reference` (plain torch, fp32).

Weight gradients do not travel through autograd's AccumulateGrad: when a
parameter carries a ``gbuf`` attribute (a view into the flat gradient arena, see
:class:`theanompi_b200.parallel.arena.FlatArena`) the backward kernel writes the
fp32 gradient straight into it and then fires ``on_ready`` — that callback is
what lets the BSP exchanger launch the fused allreduce+SGD kernel for a bucket
on a side stream while backward is still running on the main stream
(the reference is strictly sequential, ``theanompi/worker.py:94-97``).
"""
from __future__ import annotations

import torch

import os

from . import reference as ref

CACHE_COL = os.environ.get("TMPI_CACHE_COL", "1") != "0"


def _impl(x):
    if x.is_cuda:
        from . import cuda_impl
        return cuda_impl
    return ref


def compute_weight(p):
    """bf16 shadow of a master parameter when one exists (GPU), else itself."""
    sh = getattr(p, "shadow", None)
    return sh if sh is not None else p


def _sink(p, grad):
    """Deliver a parameter gradient. Returns what autograd should see."""
    if p is None or grad is None and getattr(p, "gbuf", None) is None:
        return None
    gbuf = getattr(p, "gbuf", None)
    if gbuf is None:
        return grad.to(p.dtype).view_as(p)
    if grad is not None and grad.data_ptr() != gbuf.data_ptr():
        if getattr(p, "gaccum", False):
            gbuf.add_(grad.view_as(gbuf))
        else:
            gbuf.copy_(grad.view_as(gbuf))
    cb = getattr(p, "on_ready", None)
    if cb is not None:
        cb(p)
    return None


def _gout(p):
    """Output buffer the backward kernel may write the fp32 grad into directly."""
    if getattr(p, "gaccum", False):
        return None
    return getattr(p, "gbuf", None)


# --------------------------------------------------------------------------- linear
class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, relu):
        impl = _impl(x)
        wc = compute_weight(w)
        y = impl.linear_bias_act(x, wc, b, relu)
        ctx.save_for_backward(x, y)
        ctx.w, ctx.b, ctx.relu = w, b, relu
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        w, b = ctx.w, ctx.b
        impl = _impl(x)
        wc = compute_weight(w)
        need_dx = ctx.needs_input_grad[0]
        if impl is ref:
            dx, dw, db = ref.linear_bias_act_bwd(x, wc, y, dy, ctx.relu, need_dx)
        else:
            dx, dw, db = impl.linear_bias_act_bwd(x, wc, y, dy, ctx.relu, need_dx,
                                                  dw_out=_gout(w), db_out=_gout(b))
        gb = _sink(b, db)
        gw = _sink(w, dw)
        return dx, gw, gb, None


def linear_bias_act(x, w, b, relu=True):
    return _LinearFn.apply(x, w, b, relu)


# --------------------------------------------------------------------------- conv
def _fused_pool_ok(x, relu, pool):
    """conv(+ReLU) → max-pool blocks run their backward through ONE fused kernel (pool scatter + ReLU mask + bias grad)."""
    # (bf16 path only: the fused kernel works on packed bf16 lanes; the fp32 / tf32 path runs pool-backward and ReLU-mask +
    # bias-gradient as two kernels)
    # TMPI_DETERMINISTIC=1 also takes the two-kernel route: the fused kernel reduces the bias gradient with cross-CTA atomics
    return (pool is not None and x.is_cuda and relu and pool[3] == "max" and x.dtype == torch.bfloat16
            and os.environ.get("TMPI_DETERMINISTIC") != "1")


def _pool_fwd_after(ctx, impl, y, pool):
    """Forward of the pooling half of a fused conv→pool block; returns (pooled, argmax or None)."""
    if impl is ref:
        return ref.pool2d(y, pool[0], pool[1], pool[2], pool[3]), None
    return impl.pool2d_fwd(y, pool[0], pool[1], pool[2], pool[3])


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, pad, groups, relu, pool=None):
        impl = _impl(x)
        wc = compute_weight(w)
        ctx.cols = None
        if impl is ref:
            y = ref.conv2d_bias_act(x, wc, b, stride, pad, groups, relu)
        else:
            # keep the im2col matrix of the forward for wgrad (memory is cheap on a 180 GB part; recomputing
            # it cost ~13 % of the AlexNet step)
            y, ctx.cols = impl.conv2d_bias_act(x, wc, b, stride, pad, groups, relu, return_cols=True)
            if not CACHE_COL:
                ctx.cols = None
        ctx.w, ctx.b = w, b
        ctx.cfg = (stride, pad, groups, relu)
        ctx.pool = pool
        if pool is None:
            ctx.save_for_backward(x, y)
            return y
        yp, arg = _pool_fwd_after(ctx, impl, y, pool)
        ctx.has_arg = arg is not None
        if arg is not None:
            ctx.save_for_backward(x, y, arg)
        else:
            ctx.save_for_backward(x, y, yp)
        return yp

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors[0], ctx.saved_tensors[1]
        w, b = ctx.w, ctx.b
        stride, pad, groups, relu = ctx.cfg
        pool = ctx.pool
        impl = _impl(x)
        wc = compute_weight(w)
        need_dx = ctx.needs_input_grad[0]
        if impl is ref:
            if pool is not None:
                dy = ref.pool2d_bwd(y, ctx.saved_tensors[2], dy.contiguous(), *pool)
            dx, dw, db = ref.conv2d_bias_act_bwd(x, wc, y, dy, stride, pad, groups, relu, need_dx)
        else:
            pre = False
            db_out = _gout(b)
            if pool is not None:
                if _fused_pool_ok(x, relu, pool) and ctx.has_arg:
                    if db_out is None:
                        db_out = torch.empty(b.numel(), dtype=torch.float32, device=x.device)
                    dy = impl.maxpool_relu_bias_bwd(dy, ctx.saved_tensors[2], y, pool, db_out, None)
                    pre = True
                else:
                    dy = impl.pool2d_bwd_arg(dy, ctx.saved_tensors[2] if ctx.has_arg else None, tuple(y.shape), *pool)
            dx, dw, db = impl.conv2d_bias_act_bwd(x, wc, y, dy, stride, pad, groups, relu, need_dx,
                                                  dw_out=_gout(w), db_out=db_out, cols=ctx.cols, pre_masked=pre,
                                                  need_db=b is not None)
            ctx.cols = None
        gb = _sink(b, db)
        gw = _sink(w, dw)
        return dx, gw, gb, None, None, None, None, None


def conv2d_bias_act(x, w, b, stride=1, pad=0, groups=1, relu=True, pool=None):
    """``pool`` = (ksize, stride, pad, mode): run the pooling layer that follows inside the same autograd node, so the
    backward can fuse pool-scatter + ReLU mask + bias gradient into one pass over the conv output."""
    return _ConvFn.apply(x, w, b, stride, pad, groups, relu, pool)


class _ConvG2Fn(torch.autograd.Function):
    """AlexNet-style 2-group conv with two independent parameter sets
    (ref ``layers2.py:504-544``).  On CUDA both groups read channel slices of the
    NHWC input in place (no split/concat copies): the im2col gather takes a channel
    offset and the GEMM epilogue writes each half of the output with ldc = C_out."""

    @staticmethod
    def forward(ctx, x, w0, b0, w1, b1, stride, pad, relu, pool=None):
        impl = _impl(x)
        ws = [compute_weight(w0), compute_weight(w1)]
        if impl is ref:
            C = x.shape[-1]
            y0 = ref.conv2d_bias_act(x[..., :C // 2].contiguous(), ws[0], b0, stride, pad, 1, relu)
            y1 = ref.conv2d_bias_act(x[..., C // 2:].contiguous(), ws[1], b1, stride, pad, 1, relu)
            y = torch.cat([y0, y1], dim=-1)
            ctx.cols = None
        else:
            y, ctx.cols = impl.conv2d_group2_bias_act(x, ws[0], b0, ws[1], b1, stride, pad, relu, return_cols=True)
            if not CACHE_COL:
                ctx.cols = None
        ctx.p = (w0, b0, w1, b1)
        ctx.cfg = (stride, pad, relu)
        ctx.pool = pool
        if pool is None:
            ctx.save_for_backward(x, y)
            return y
        yp, arg = _pool_fwd_after(ctx, impl, y, pool)
        ctx.has_arg = arg is not None
        if arg is not None:
            ctx.save_for_backward(x, y, arg)
        else:
            ctx.save_for_backward(x, y, yp)
        return yp

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors[0], ctx.saved_tensors[1]
        w0, b0, w1, b1 = ctx.p
        stride, pad, relu = ctx.cfg
        pool = ctx.pool
        impl = _impl(x)
        need_dx = ctx.needs_input_grad[0]
        if impl is ref:
            if pool is not None:
                dy = ref.pool2d_bwd(y, ctx.saved_tensors[2], dy.contiguous(), *pool)
            C, O = x.shape[-1], y.shape[-1]
            outs = []
            for g, (w, b) in enumerate(((w0, b0), (w1, b1))):
                xs = x[..., g * C // 2:(g + 1) * C // 2].contiguous()
                ys = y[..., g * O // 2:(g + 1) * O // 2].contiguous()
                dys = dy[..., g * O // 2:(g + 1) * O // 2].contiguous()
                outs.append(ref.conv2d_bias_act_bwd(xs, compute_weight(w), ys, dys, stride, pad, 1,
                                                    relu, need_dx))
            dx = torch.cat([outs[0][0], outs[1][0]], dim=-1) if need_dx else None
            grads = (outs[0][1], outs[0][2], outs[1][1], outs[1][2])
        else:
            pre = False
            db0, db1 = _gout(b0), _gout(b1)
            if pool is not None:
                if _fused_pool_ok(x, relu, pool) and ctx.has_arg:
                    if db0 is None:
                        db0 = torch.empty(b0.numel(), dtype=torch.float32, device=x.device)
                    if db1 is None:
                        db1 = torch.empty(b1.numel(), dtype=torch.float32, device=x.device)
                    dy = impl.maxpool_relu_bias_bwd(dy, ctx.saved_tensors[2], y, pool, db0, db1)
                    pre = True
                else:
                    dy = impl.pool2d_bwd_arg(dy, ctx.saved_tensors[2] if ctx.has_arg else None, tuple(y.shape), *pool)
            dx, grads = impl.conv2d_group2_bias_act_bwd(
                x, compute_weight(w0), compute_weight(w1), y, dy, stride, pad, relu, need_dx,
                outs=(_gout(w0), db0, _gout(w1), db1), cols=ctx.cols, pre_masked=pre)
            ctx.cols = None
        gb1 = _sink(b1, grads[3])
        gw1 = _sink(w1, grads[2])
        gb0 = _sink(b0, grads[1])
        gw0 = _sink(w0, grads[0])
        return dx, gw0, gb0, gw1, gb1, None, None, None, None


def conv2d_group2_bias_act(x, w0, b0, w1, b1, stride=1, pad=0, relu=True, pool=None):
    return _ConvG2Fn.apply(x, w0, b0, w1, b1, stride, pad, relu, pool)


# --------------------------------------------------------------------------- batch norm (+ residual)(+ ReLU)
class _BatchNormFn(torch.autograd.Function):
    """``relu(γ·x̂ + β + residual)`` as one forward pass (statistics + apply) and one backward pass (reduce + apply); the
    parameter gradients are written straight into the arena's G views (``_gout`` / ``_sink``).  Reference: Lasagne
    ``batch_norm`` + ``ElemwiseSumLayer`` + ``rectify`` (``lasagne_model_zoo/resnet50.py:14-77``)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, residual, run_mean, run_var, training, momentum, eps, relu):
        impl = _impl(x)
        y, mean, rstd = impl.batch_norm_fwd(x, gamma.detach(), beta.detach(), run_mean, run_var, training, momentum, eps, relu, residual)
        ctx.gamma, ctx.beta = gamma, beta
        ctx.relu, ctx.has_res = relu, residual is not None
        ctx.save_for_backward(x, y if relu else x.new_empty(0), mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, mean, rstd = ctx.saved_tensors
        gamma, beta = ctx.gamma, ctx.beta
        impl = _impl(x)
        need_dres = ctx.has_res and ctx.needs_input_grad[3]
        if impl is ref:
            dx, dres, dg, db = ref.batch_norm_bwd(x, dy, y, gamma.detach(), mean, rstd, ctx.relu, need_dres)
        else:
            dx, dres, dg, db = impl.batch_norm_bwd(x, dy, y, gamma.detach(), mean, rstd, ctx.relu, need_dres,
                                                   dgamma_out=_gout(gamma), dbeta_out=_gout(beta))
        gb = _sink(beta, db)
        gg = _sink(gamma, dg)
        return dx, gg, gb, dres, None, None, None, None, None, None


def batch_norm(x, gamma, beta, run_mean=None, run_var=None, training=True, momentum=0.1, eps=1e-5, relu=False, residual=None):
    """Batch normalisation over all but the channel (last) axis, optionally fused with a residual add and a ReLU."""
    return _BatchNormFn.apply(x, gamma, beta, residual, run_mean, run_var, training, momentum, eps, relu)


class _AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        if a.is_cuda:
            from . import cuda_impl
            return cuda_impl.add(a, b)
        return a + b

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


class _Fork2Fn(torch.autograd.Function):
    """Fan-out of a tensor to two consumers: the two incoming gradients are summed by the native add kernel (autograd would
    otherwise accumulate them with a library elementwise kernel)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x), x.view_as(x)

    @staticmethod
    def backward(ctx, d1, d2):
        if d1 is None:
            return d2
        if d2 is None:
            return d1
        if d1.is_cuda:
            from . import cuda_impl
            return cuda_impl.add(d1, d2)
        return d1 + d2


def fork2(x):
    """Return two aliases of ``x`` for two consumers (residual shortcut + branch)."""
    if not x.requires_grad:
        return x, x
    return _Fork2Fn.apply(x)


def add(a, b):
    """Residual merge ``a + b`` (native kernel on CUDA; the gradient passes to both branches unchanged)."""
    return _AddFn.apply(a, b)


# --------------------------------------------------------------------------- pool
class _PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ksize, stride, pad, mode):
        ctx.cfg = (ksize, stride, pad, mode)
        if x.is_cuda:
            from . import cuda_impl
            y, arg = cuda_impl.pool2d_fwd(x, ksize, stride, pad, mode)
            ctx.xshape = tuple(x.shape)
            ctx.has_arg = arg is not None
            if arg is not None:
                ctx.save_for_backward(arg)
            return y
        y = ref.pool2d(x, ksize, stride, pad, mode)
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        if dy.is_cuda:
            from . import cuda_impl
            arg = ctx.saved_tensors[0] if ctx.has_arg else None
            return cuda_impl.pool2d_bwd_arg(dy, arg, ctx.xshape, *ctx.cfg), None, None, None, None
        x, y = ctx.saved_tensors
        dx = ref.pool2d_bwd(x, y, dy.contiguous(), *ctx.cfg)
        return dx, None, None, None, None


def pool2d(x, ksize, stride, pad=0, mode="max"):
    return _PoolFn.apply(x, ksize, stride, pad, mode)


# --------------------------------------------------------------------------- LRN
class _LRNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n, k, alpha, beta):
        impl = _impl(x)
        y, _ = impl.lrn(x, n, k, alpha, beta)
        ctx.save_for_backward(x)
        ctx.cfg = (n, k, alpha, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dx = _impl(x).lrn_bwd(x, dy.contiguous(), *ctx.cfg)
        return dx, None, None, None, None


def lrn(x, n=5, k=2.0, alpha=1e-4, beta=0.75):
    return _LRNFn.apply(x, n, k, alpha, beta)


# --------------------------------------------------------------------------- dropout
class _DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p_drop, layer_id):
        if x.is_cuda:
            from . import cuda_impl
            y, mask = cuda_impl.dropout_fwd(x, p_drop, layer_id)
        else:
            st = rng_state()
            mask = ref.dropout_mask(x.shape, p_drop, st["seed"] + layer_id, st["step"], x.device)
            y = ref.dropout(x, p_drop, mask)
        ctx.save_for_backward(mask)
        return y

    @staticmethod
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        if dy.is_cuda:
            from . import cuda_impl
            return cuda_impl.dropout_bwd(dy.contiguous(), mask), None, None
        return dy * mask.to(dy.dtype), None, None


_RNG = {"seed": 0x5EED, "step": 0}


def rng_state():
    return _RNG


def seed_dropout(seed: int):
    _RNG["seed"] = int(seed)
    _RNG["step"] = 0


def advance_rng_step():
    """CPU path only; on CUDA the device step counter advances inside the
    (graph-captured) step so replays draw fresh masks."""
    _RNG["step"] += 1


def dropout(x, p_drop, training, layer_id=0):
    """Reference semantics (``layers2.py:885-891``): train → ``mask*x``;
    eval → ``(1-p)*x`` (no inverted scaling)."""
    if not training:
        return x * (1.0 - p_drop)
    if p_drop <= 0.0:
        return x
    return _DropoutFn.apply(x, p_drop, layer_id)


# --------------------------------------------------------------------------- softmax + NLL
class _SoftmaxXentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        loss, err1, err5, dlogits = _impl(logits).softmax_xent(logits, labels)
        ctx.save_for_backward(dlogits)
        ctx.in_dtype = logits.dtype
        ctx.mark_non_differentiable(err1, err5)
        return loss, err1, err5

    @staticmethod
    def backward(ctx, gl, g1, g5):
        (dlogits,) = ctx.saved_tensors
        # gl is a 0-dim tensor on the same device: no host sync, graph-capturable
        return (dlogits * gl.to(dlogits.dtype)).to(ctx.in_dtype), None


def softmax_xent(logits, labels):
    """Returns (mean NLL, top-1 error, top-5 error) — fused on CUDA."""
    return _SoftmaxXentFn.apply(logits, labels)


# --------------------------------------------------------------------------- data aug
def crop_mirror_normalize(x, mean, std_scale, crop_hw, offsets, flips, out_dtype=None):
    if x.is_cuda:
        from . import cuda_impl
        return cuda_impl.crop_mirror_normalize(x, mean, std_scale, crop_hw, offsets, flips, out_dtype)
    return ref.crop_mirror_normalize(x, mean, std_scale, crop_hw, offsets, flips,
                                     out_dtype or torch.float32)
