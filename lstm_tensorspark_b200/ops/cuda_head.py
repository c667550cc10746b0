"""CUDA head op: dense layer + softmax cross-entropy + accuracy in one launch (csrc/head_xent.cu);
backward = two small library GEMMs on the dlogits the forward kernel already produced.
Parity: /root/reference/src/rnn.py:214-221 (Dense1), :55-63 (loss), :84-92 (accuracy)."""
from __future__ import annotations

import torch

from .cuda_ext import ext


class _HeadXentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, weights, bias, labels):
        E = ext()
        B = h.shape[0]
        hc = h.detach().contiguous()
        if hc.dtype not in (torch.bfloat16, torch.float32):
            hc = hc.float()
        w = weights.detach().float().contiguous()
        b = bias.detach().float().contiguous()
        lab = labels.long().contiguous()
        if w.shape[1] <= 32:
            logits, dlogits, loss_sum, correct = E.head_xent(hc, w, b, lab)
        else:
            logits = torch.addmm(b, hc.float(), w)
            dlogits, loss_sum, correct = E.xent_rows(logits, lab)
        ctx.save_for_backward(hc, w, dlogits)
        ctx.h_dtype = h.dtype
        loss = (loss_sum / B).squeeze(0)
        ctx.mark_non_differentiable(logits, correct)
        return logits, loss, correct.squeeze(0)

    @staticmethod
    def backward(ctx, _dlogits_unused, dloss, _dcorrect_unused):
        hc, w, dlogits = ctx.saved_tensors
        d = dlogits * dloss
        dh = (d @ w.t()).to(ctx.h_dtype)
        dw = hc.float().t() @ d
        db = d.sum(0)
        return dh, dw, db, None


def head_xent(h, weights, bias, labels):
    return _HeadXentFn.apply(h, weights, bias, labels)
