"""CUDA head op: dense layer + softmax cross-entropy + accuracy in ONE launch on the tensor cores (csrc/head_tc.cu:
TMA-fed tcgen05 tile, logits in TMEM, softmax / NLL / accuracy / dlogits in the epilogue) and the whole backward
(dh, dW, db) in ONE launch that writes dW / db straight into the flat gradient buffer.
Parity: /root/reference/src/rnn.py:214-221 (Dense1), :55-63 (loss), :84-92 (accuracy), :224 (autodiff)."""
from __future__ import annotations

import torch

from .cuda_ext import ext


class _HeadXentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, weights, bias, labels):
        E = ext()
        B = h.shape[0]
        hc = h.detach()
        if hc.dtype not in (torch.bfloat16, torch.float32):
            hc = hc.float()
        if hc.stride(-1) != 1:
            hc = hc.contiguous()
        w = weights.detach().float().contiguous()
        b = bias.detach().float().contiguous()
        lab = labels.long().contiguous()
        logits, dlogits, loss_sum, correct = E.head_fwd(hc, w, b, lab)
        ctx.save_for_backward(hc, w, dlogits)
        ctx.h_dtype = h.dtype
        ctx.addrs = (weights.data_ptr(), bias.data_ptr())
        loss = (loss_sum / B).squeeze(0)
        ctx.mark_non_differentiable(logits, correct)
        return logits, loss, correct.squeeze(0)

    @staticmethod
    def backward(ctx, _dlogits_unused, dloss, _dcorrect_unused):
        from .cuda_lstm import grad_sink
        E = ext()
        hc, w, dlogits = ctx.saved_tensors
        hcc = hc if hc.is_contiguous() else hc.contiguous()
        sw, sb = grad_sink(ctx.addrs[0]), grad_sink(ctx.addrs[1])
        dl = dloss.detach().float().reshape(1).contiguous()
        if sw is not None and sb is not None and sw[1] == sb[1]:
            dh = E.head_bwd(hcc, w, dlogits, dl, sw[0], sb[0], sw[1])
            return dh.to(ctx.h_dtype), None, None, None
        if sw is not None and sb is not None:             # one of the two already holds a gradient: bring both to "accumulate"
            if not sw[1]:
                sw[0].zero_()
            if not sb[1]:
                sb[0].zero_()
            dh = E.head_bwd(hcc, w, dlogits, dl, sw[0], sb[0], True)
            return dh.to(ctx.h_dtype), None, None, None
        dw = torch.empty_like(w)
        db = torch.empty(w.shape[1], dtype=torch.float32, device=w.device)
        dh = E.head_bwd(hcc, w, dlogits, dl, dw, db, False)
        return dh.to(ctx.h_dtype), dw, db, None


def head_xent(h, weights, bias, labels):
    return _HeadXentFn.apply(h, weights, bias, labels)
