"""torch.autograd bridge: lets `loss.backward()` on anything computed from the student's outputs
drive the engine's hand-written backward (used by the reference-style API path; the fused training
step in distill.py bypasses autograd entirely)."""
from __future__ import annotations

import torch

from . import engine as E


class _StudentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, source, valid, *params):
        P, W, G = model.engine_state(True)
        c = E.student_forward(P, W, model._geom, source, valid, train=True, heads="all", want_lr=False,
                              drop=model.drop_cfg())
        ctx.model, ctx.c, ctx.names = model, c, [n for n, _ in model.named_parameters()]
        model._last_ctx = c
        return (c.preds, *c.layers)

    @staticmethod
    def backward(ctx, dpreds, *dlayers):
        model, c = ctx.model, ctx.c
        P, W, G = model.engine_state(True)
        G.zero_()
        if dpreds is None:
            dpreds = torch.zeros_like(c.preds)
        dpreds = dpreds.to(torch.bfloat16).contiguous()
        dl = [None if d is None else d.to(torch.bfloat16).contiguous() for d in dlayers]
        E.student_backward(P, W, model._geom, G, c, dpreds, dl)
        grads = G.export()
        return (None, None, None, *[grads.get(n) for n in ctx.names])


def student_apply(model, source, valid):
    params = [p for _, p in model.named_parameters()]
    outs = _StudentFn.apply(model, source, valid, *params)
    return model._last_ctx, outs[0], list(outs[1:])


class _DistillLossFn(torch.autograd.Function):
    """Fused K10: per-layer weighted MSE/L1 between projections and teacher layers + d(loss)/d(pred)
    in one pass (reference train.py:250-293)."""

    @staticmethod
    def forward(ctx, preds, tgt, weights, loss_type):
        from . import kernels as K
        n, B, Tq, D = preds.shape
        Tt = tgt.shape[2]
        layer_loss = torch.zeros(n, device=preds.device, dtype=torch.float32)
        dpred = torch.empty_like(preds)
        K.distill_loss(preds, tgt, weights, layer_loss, dpred, n, B, Tq, Tt, D, loss_type, 1.0)
        ctx.save_for_backward(dpred)
        ctx.args = (preds, tgt, weights, n, B, Tq, Tt, D, loss_type)
        ctx.mark_non_differentiable(layer_loss)
        total = layer_loss.sum()  # 12-element reduction; the heavy lifting is the kernel above
        return total, layer_loss

    @staticmethod
    def backward(ctx, gtotal, _glayers):
        from . import kernels as K
        (dpred,) = ctx.saved_tensors
        scale = float(gtotal)  # one scalar D2H; 1.0 unless the caller rescales the loss (grad accumulation)
        if scale != 1.0:
            preds, tgt, weights, n, B, Tq, Tt, D, loss_type = ctx.args
            scratch = torch.zeros(n, device=preds.device, dtype=torch.float32)
            K.distill_loss(preds, tgt, weights, scratch, dpred, n, B, Tq, Tt, D, loss_type, scale)
        return dpred, None, None, None
