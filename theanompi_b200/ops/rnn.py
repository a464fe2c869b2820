"""Recurrent ops: embedding lookup, a whole masked LSTM sequence as ONE autograd node, masked mean pooling over time.

Reference: ``theanompi/models/lstm.py:117-253`` (the Theano tutorial LSTM: ``preact = x_t·W + h_{t-1}·U + b`` sliced
``i | f | o | c̃``, state carried through the padding mask, mean pooling).  CUDA: the matrix products run on the tcgen05 GEMM
(``cuda_impl.gemm``), the gate math / state carry / pooling / embedding scatter are hand-written kernels
(``csrc/rnn_kernels.cu``); the recurrent weight gradient of the whole sequence is ONE GEMM over the stacked time steps.
CPU: plain torch with the same formulas.
"""
from __future__ import annotations

import torch

from .functional import _gout, _sink, compute_weight


def _ci():
    from . import cuda_impl
    return cuda_impl


# --------------------------------------------------------------------------- embedding
class _EmbeddingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, W):
        ctx.W = W
        ctx.save_for_backward(ids)
        wc = compute_weight(W)
        if ids.is_cuda:
            ci = _ci()
            wc = ci._bf(wc).contiguous()
            flat = ids.reshape(-1).contiguous()
            out = torch.empty((flat.numel(), W.shape[1]), dtype=wc.dtype, device=ids.device)
            ci.L().embedding_fwd(flat.data_ptr(), wc.data_ptr(), out.data_ptr(), flat.numel(), int(W.shape[1]), int(ci._is32(wc)), ci._st(out))
            return out.view(tuple(ids.shape) + (W.shape[1],))
        return wc[ids]

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        W = ctx.W
        V, D = W.shape
        if dout.is_cuda:
            ci = _ci()
            dout = ci._bf(dout).contiguous()
            gb = _gout(W)
            dW = gb if gb is not None else torch.empty((V, D), dtype=torch.float32, device=dout.device)
            flat = ids.reshape(-1).contiguous()
            ci.L().embedding_bwd(flat.data_ptr(), dout.data_ptr(), dW.data_ptr(), flat.numel(), int(D), int(V), int(ci._is32(dout)), ci._st(dout))
        else:
            dW = torch.zeros((V, D), dtype=torch.float32, device=dout.device)
            dW.index_add_(0, ids.reshape(-1), dout.reshape(-1, D).float())
        return None, _sink(W, dW)


def embedding(ids, W):
    """``W[ids]`` (ids int64 of any shape) — gather forward, scatter-add backward straight into the arena's gradient view."""
    return _EmbeddingFn.apply(ids, W)


# --------------------------------------------------------------------------- masked LSTM sequence
class _LSTMSeqFn(torch.autograd.Function):
    """h[t] for t < T from the input pre-activations ``gx`` [T, B, 4H] (x_t·W + b), the recurrent matrix ``U`` [4H, H] and the
    padding mask [T, B]; zero initial state."""

    @staticmethod
    def forward(ctx, gx, U, mask):
        Tn, B, H4 = gx.shape
        H = H4 // 4
        uc = compute_weight(U)
        dev = gx.device
        if gx.is_cuda:
            ci = _ci()
            L = ci.L()
            gx = ci._bf(gx).contiguous()
            uc = ci._bf(uc).contiguous()
            f32 = int(ci._is32(gx))
            dt = gx.dtype
            mask = mask.float().contiguous()
            hs = torch.zeros((Tn + 1, B, H), dtype=dt, device=dev)            # hs[t] = h_{t-1}; hs[0] = 0
            cs = torch.zeros((Tn + 1, B, H), dtype=torch.float32, device=dev)
            act = torch.empty((Tn, B, H4), dtype=dt, device=dev)
            gh = torch.empty((B, H4), dtype=dt, device=dev)
            for t in range(Tn):
                ci.gemm(hs[t], uc, B, H4, H, out=gh, lda=H, ldb=H, ldc=H4)        # h_{t-1} · Uᵀ
                L.lstm_cell_fwd(gx[t].data_ptr(), gh.data_ptr(), cs[t].data_ptr(), hs[t].data_ptr(), mask[t].data_ptr(), act[t].data_ptr(),
                                cs[t + 1].data_ptr(), hs[t + 1].data_ptr(), B, H, f32, ci._st(gx))
        else:
            mask = mask.float()
            gxf, ucf = gx.float(), uc.float()
            hs = torch.zeros((Tn + 1, B, H), device=dev)
            cs = torch.zeros((Tn + 1, B, H), device=dev)
            act = torch.empty((Tn, B, H4), device=dev)
            for t in range(Tn):
                pre = gxf[t] + hs[t] @ ucf.t()
                i, f, o = torch.sigmoid(pre[:, :H]), torch.sigmoid(pre[:, H:2 * H]), torch.sigmoid(pre[:, 2 * H:3 * H])
                g = torch.tanh(pre[:, 3 * H:])
                m = mask[t][:, None]
                ct = f * cs[t] + i * g
                cs[t + 1] = m * ct + (1 - m) * cs[t]
                hs[t + 1] = m * (o * torch.tanh(ct)) + (1 - m) * hs[t]
                act[t] = torch.cat([i, f, o, g], 1)
            hs = hs.to(gx.dtype)
        ctx.U = U
        ctx.save_for_backward(hs, cs, act, mask)
        return hs[1:]

    @staticmethod
    def backward(ctx, dh_all):
        hs, cs, act, mask = ctx.saved_tensors
        U = ctx.U
        Tn, B, H4 = act.shape
        H = H4 // 4
        uc = compute_weight(U)
        dev = act.device
        if act.is_cuda:
            ci = _ci()
            L = ci.L()
            dt = act.dtype
            f32 = int(ci._is32(act))
            dh_all = ci._bf(dh_all).contiguous()
            uc = ci._bf(uc).contiguous()
            dG = torch.empty((Tn, B, H4), dtype=dt, device=dev)
            dc = [torch.zeros((B, H), dtype=torch.float32, device=dev) for _ in range(2)]
            dpass = [torch.zeros((B, H), dtype=torch.float32, device=dev) for _ in range(2)]
            drec = torch.empty((B, H), dtype=dt, device=dev)
            for t in range(Tn - 1, -1, -1):
                last = t == Tn - 1
                k = t & 1
                L.lstm_cell_bwd(dh_all[t].data_ptr(), 0 if last else drec.data_ptr(), 0 if last else dpass[1 - k].data_ptr(),
                                0 if last else dc[1 - k].data_ptr(), act[t].data_ptr(), cs[t + 1].data_ptr(), cs[t].data_ptr(),
                                mask[t].data_ptr(), dG[t].data_ptr(), dc[k].data_ptr(), dpass[k].data_ptr(), B, H, f32, ci._st(act))
                if t > 0:
                    ci.gemm(dG[t], uc, B, H, H4, b_mn=True, out=drec, lda=H4, ldb=H, ldc=H)      # dG_t · U  → dh_{t-1}
            # dU = Σ_t dG_tᵀ · h_{t-1}: ONE GEMM over the stacked time steps, fp32, straight into the arena's gradient view
            gb = _gout(U)
            dU = gb if gb is not None else torch.empty((H4, H), dtype=torch.float32, device=dev)
            ci.gemm(dG.view(Tn * B, H4), hs[:Tn].reshape(Tn * B, H), H4, H, Tn * B, a_mn=True, b_mn=True, out=dU, lda=H4, ldb=H, ldc=H)
        else:
            ucf = uc.float()
            dG = torch.empty((Tn, B, H4), device=dev)
            dc = torch.zeros((B, H), device=dev)
            dh_next = torch.zeros((B, H), device=dev)
            for t in range(Tn - 1, -1, -1):
                m = mask[t][:, None]
                dh = dh_all[t].float() + dh_next
                i, f, o, g = act[t][:, :H], act[t][:, H:2 * H], act[t][:, 2 * H:3 * H], act[t][:, 3 * H:]
                tc = torch.tanh(cs[t + 1])
                dht = m * dh
                dct = m * dc + dht * o * (1 - tc * tc)
                dG[t] = torch.cat([dct * g * i * (1 - i), dct * cs[t] * f * (1 - f), dht * tc * o * (1 - o), dct * i * (1 - g * g)], 1)
                dc = dct * f + (1 - m) * dc
                dh_next = (1 - m) * dh + dG[t] @ ucf
            dU = dG.reshape(Tn * B, H4).t() @ hs[:Tn].float().reshape(Tn * B, H)
            dG = dG.to(dh_all.dtype)
        return dG, _sink(U, dU), None


def lstm_sequence(gx, U, mask):
    return _LSTMSeqFn.apply(gx, U, mask)


# --------------------------------------------------------------------------- masked mean over time
class _MaskedMeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, mask):
        Tn, B, H = h.shape
        mask = mask.float().contiguous()
        ctx.save_for_backward(mask)
        ctx.shape = (Tn, B, H)
        if h.is_cuda:
            ci = _ci()
            h = ci._bf(h).contiguous()
            out = torch.empty((B, H), dtype=h.dtype, device=h.device)
            ci.L().masked_mean_fwd(h.data_ptr(), mask.data_ptr(), out.data_ptr(), Tn, B, H, int(ci._is32(h)), ci._st(h))
            return out
        return (h.float() * mask[..., None]).sum(0) / mask.sum(0)[:, None].clamp_min(1)

    @staticmethod
    def backward(ctx, dout):
        (mask,) = ctx.saved_tensors
        Tn, B, H = ctx.shape
        if dout.is_cuda:
            ci = _ci()
            dout = ci._bf(dout).contiguous()
            dh = torch.empty((Tn, B, H), dtype=dout.dtype, device=dout.device)
            ci.L().masked_mean_bwd(dout.data_ptr(), mask.data_ptr(), dh.data_ptr(), Tn, B, H, int(ci._is32(dout)), ci._st(dout))
            return dh, None
        return dout[None].float() * (mask / mask.sum(0).clamp_min(1))[..., None], None


def masked_mean(h, mask):
    """Mean of ``h`` [T, B, H] over the valid time steps of every sequence (ref ``lstm.py:217-253``)."""
    return _MaskedMeanFn.apply(h, mask)
