"""CUDA flat optimizer step: one launch over the whole parameter buffer (csrc/multi_tensor_opt.cu)."""
from __future__ import annotations

from .cuda_ext import ext


def flat_step(opt, grad_scale: float = 1.0) -> None:
    E = ext()
    fl = opt.flat
    shadow = fl.shadow
    if opt.kind == "adam":
        E.flat_adam(fl.data, fl.grad, opt.m, opt.v, shadow, opt.lr, opt.beta1, opt.beta2, opt.eps,
                    opt.weight_decay, grad_scale, opt.step_dev, opt.wd_numel)
    else:
        E.flat_sgd(fl.data, fl.grad, shadow, opt.lr, opt.weight_decay, grad_scale, opt.wd_numel)


def cast_shadow(fl) -> None:
    if fl.shadow is not None:
        ext().cast_bf16(fl.data, fl.shadow)
