"""torch.autograd bridge: lets `loss.backward()` on anything computed from the student's outputs
drive the engine's hand-written backward (used by the reference-style API path; the fused training
step in distill.py bypasses autograd entirely)."""
from __future__ import annotations

import torch

from . import engine as E

# Loss scale of the fp16 gradient that _DistillLossFn.backward has just handed to autograd: consumed by the next
# _StudentFn.backward (the chain between the two is linear - views, stacks, permutes - so the scale survives it),
# which divides the exported parameter gradients by it.  Gradients of any other loss arrive un-scaled (scale 1):
# callers that feed fp16 outputs into their own loss scale it themselves, as under torch.cuda.amp.
_PENDING_SCALE = [1.0]


class _StudentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, source, valid, *params):
        P, W, G = model.engine_state(True)
        # after _disable_projection_heads (the s3prl expert: fithubert/expert.py:44) only final_proj is left
        heads = "all" if model.proj_head is not None else ("last" if model.final_proj is not None else "none")
        c = E.student_forward(P, W, model._geom, source, valid, train=True, heads=heads, want_lr=False,
                              drop=model.drop_cfg())
        ctx.model, ctx.c, ctx.names = model, c, [n for n, _ in model.named_parameters()]
        ctx.set_materialize_grads(False)  # outputs the loss never touched arrive as None, not as zero tensors
        model._last_ctx = c
        feats = c.cnn_out if c.cnn_out is not None else c.feats  # the returned `features` (features_to_distill)
        outs = (c.preds if c.preds is not None else source.new_zeros(0), feats, *c.layers)
        ctx.n_layers_out = len(c.layers)
        if getattr(model, "_return_attn", False):
            # attention-map recipe: the last layer's (attn_logits, v_rel), fp32 [B*H, T, pitch] (train.py:64-77)
            g = model._geom
            maps = E.attn_maps(c.layer_ctx[-1].qkv, c.valid_s, c.B, c.Ts, g.H, g.d)
            outs = outs + (maps["attn"], maps["vrel"])
        # autograd stamps the RETURNED tensor objects with this node's grad_fn.  The context the backward keeps must not
        # hold those same objects, or node -> ctx -> c -> tensor -> node is a reference cycle no collector sees (every
        # call's saved activations would stay allocated): keep detached aliases instead.
        if c.preds is not None:
            c.preds = c.preds.detach()
        if c.cnn_out is not None:
            c.cnn_out = c.cnn_out.detach()
        else:
            c.feats = c.feats.detach()
        c.layers = [t.detach() for t in c.layers]
        for s_, t in zip(c.layer_ctx, c.layers):
            s_.out = t
        c.x_last = c.x_last.detach()
        return outs

    @staticmethod
    def backward(ctx, dpreds, dfeat, *rest):
        model, c = ctx.model, ctx.c
        dlayers, dmaps = rest[:ctx.n_layers_out], rest[ctx.n_layers_out:]
        P, W, G = model.engine_state(True)
        G.zero_()
        if c.preds is None:
            dpreds = None
        elif dpreds is None:
            dpreds = torch.zeros_like(c.preds)
        scale, _PENDING_SCALE[0] = _PENDING_SCALE[0], 1.0
        if dpreds is not None:
            dpreds = dpreds.to(torch.float16).contiguous()
        dl = [None if d is None else (d.float() * scale).to(torch.float16).contiguous() if scale != 1.0 else
              d.to(torch.float16).contiguous() for d in dlayers]
        if dfeat is not None:
            pre_scaled, _FEAT_GRAD_SCALED[0] = _FEAT_GRAD_SCALED[0], False
            dfeat = ((dfeat.float() * scale) if (scale != 1.0 and not pre_scaled) else dfeat).to(torch.float16).contiguous()
        c.attn_grad = []
        if any(d is not None for d in dmaps):
            # gradients wrt the last layer's maps -> fp16 operands of the head-gradient kernels, centred in fp16's range
            # by a power of two that `alpha` takes out again (one scalar D2H per map: this is the autograd API path)
            pre_scaled, _MAP_GRAD_SCALED[0] = _MAP_GRAD_SCALED[0], False
            sc = 1.0 if pre_scaled else scale
            for kind, d in zip(("qk", "vv"), dmaps):
                if d is None:
                    continue
                d = torch.nan_to_num(d.float(), nan=0.0, posinf=0.0, neginf=0.0)
                amax = float(d.abs().max())
                if amax == 0.0:
                    continue
                boost = E._pow2_near(256.0 / (amax * sc))
                c.attn_grad.append((kind, (d * (sc * boost)).to(torch.float16).contiguous(), model._geom.d ** -0.5 / boost))
        E.student_backward(P, W, model._geom, G, c, dpreds, dl, dfeatures=dfeat)
        grads = G.export(scale=scale)
        return (None, None, None, *[grads.get(n) for n in ctx.names])


def student_apply(model, source, valid):
    params = [p for _, p in model.named_parameters()]
    outs = _StudentFn.apply(model, source, valid, *params)
    c = model._last_ctx
    n = len(c.layers)
    maps = tuple(outs[2 + n:]) if len(outs) > 2 + n else None
    return c, (outs[0] if c.preds is not None else None), list(outs[2:2 + n]), outs[1], maps


class _DistillLossFn(torch.autograd.Function):
    """Fused K10: per-layer weighted MSE/L1 (+ optional cosine, train.py:302-314) between projections and teacher
    layers + d(loss)/d(pred) in one pass (reference train.py:250-314).
    apply(preds, tgt, weights, loss_type[, rec_weight=1, sim_weight=0, loss_scale=auto]) -> (total, per-layer rec
    [, per-layer sim]); total = rec_weight * sum(rec) + sim_weight * sum(sim).  The fp16 gradient handed back to autograd
    is multiplied by loss_scale (see _PENDING_SCALE)."""

    @staticmethod
    def forward(ctx, preds, tgt, weights, loss_type, *extra):
        rec_weight, sim_weight, scale = (tuple(extra) + (1.0, 0.0, 0.0)[len(extra):])[:3]
        n, B, Tq, D = preds.shape
        ctx.scale = float(scale) if scale else E.loss_scale_for(B * Tq * D)
        Tt = tgt.shape[2]
        ctx.n_in = 4 + len(extra)
        rec = torch.zeros(n, device=preds.device, dtype=torch.float32)
        dpred = torch.empty_like(preds)
        ctx.args = (preds, tgt, weights, n, B, Tq, Tt, D, loss_type, float(rec_weight), float(sim_weight))
        sim = _run_loss(ctx.args, rec, dpred, ctx.scale)
        ctx.save_for_backward(dpred)
        total = rec.sum() * rec_weight  # 12-element reductions; the heavy lifting is the kernel above
        if not sim_weight:
            ctx.mark_non_differentiable(rec)
            return total, rec
        ctx.mark_non_differentiable(rec, sim)
        return total + sim.sum() * sim_weight, rec, sim

    @staticmethod
    def backward(ctx, gtotal, *_glayers):
        (dpred,) = ctx.saved_tensors
        scale = float(gtotal)  # one scalar D2H; 1.0 unless the caller rescales the loss (grad accumulation)
        if scale != 1.0:
            scratch = torch.zeros(ctx.args[3], device=dpred.device, dtype=torch.float32)
            _run_loss(ctx.args, scratch, dpred, scale * ctx.scale)
        _PENDING_SCALE[0] = ctx.scale
        return (dpred,) + (None,) * (ctx.n_in - 1)


class _FeatureL1Fn(torch.autograd.Function):
    """CNN-feature loss (train.py:241-246): mean |features - teacher features| and its (loss-scaled) fp16 gradient from
    the same kernel as the layer losses, run with one "layer".  apply(features [B, T, D], teacher_features, loss_scale)."""

    @staticmethod
    def forward(ctx, feats, tgt, scale):
        from . import kernels as K
        B, T, D = feats.shape
        assert tgt.shape == feats.shape, "CNN-feature loss: student and teacher features differ in shape"
        feats, tgt = feats.contiguous(), tgt.contiguous()
        ctx.scale = float(scale)
        rec = torch.zeros(1, device=feats.device, dtype=torch.float32)
        dpred = torch.empty_like(feats)
        ctx.args = (feats, tgt, B, T, D)
        K.distill_loss(feats.view(1, B, T, D), tgt.view(1, B, T, D), torch.ones(1, device=feats.device), rec,
                       dpred.view(1, B, T, D), 1, B, T, T, D, 1, ctx.scale)
        ctx.save_for_backward(dpred)
        return rec[0]

    @staticmethod
    def backward(ctx, g):
        from . import kernels as K
        (dpred,) = ctx.saved_tensors
        w = float(g)  # cnn_loss_weight (x any caller rescaling)
        if w != 1.0:
            feats, tgt, B, T, D = ctx.args
            K.distill_loss(feats.view(1, B, T, D), tgt.view(1, B, T, D), torch.ones(1, device=feats.device),
                           torch.zeros(1, device=feats.device), dpred.view(1, B, T, D), 1, B, T, T, D, 1, ctx.scale * w)
        _PENDING_SCALE[0] = ctx.scale
        _FEAT_GRAD_SCALED[0] = True
        return dpred, None, None


class _AttnMapLossFn(torch.autograd.Function):
    """Attention distribution transfer / value relation transfer loss on [B*H, T, T] maps (train.py:327-368) and its
    gradient wrt the student's map from one kernel.
    apply(pred, target, valid_s, valid_t (device int32 or None), H, mode (0 mse, 1 kldiv), loss_mult, loss_scale)."""

    @staticmethod
    def forward(ctx, pred, target, valid_s, valid_t, H, mode, loss_mult, scale):
        from . import kernels as K
        BH, T = pred.shape[0], pred.shape[1]

        def pitched(t):  # rows of a multiple of 8 floats (what the engine's maps are views of); else a padded copy
            if t.dtype == torch.float32 and t.stride(2) == 1 and t.stride(1) % 8 == 0 and t.stride(0) == T * t.stride(1) \
                    and t.data_ptr() % 16 == 0:
                return t.as_strided((BH, T, t.stride(1)), (T * t.stride(1), t.stride(1), 1))
            p = torch.empty(BH, T, (T + 7) // 8 * 8, device=t.device, dtype=torch.float32)
            p[..., :T] = t
            return p

        s_, t_ = pitched(pred.detach()), pitched(target.detach())
        if s_.shape != t_.shape:
            t2 = torch.empty_like(s_)
            t2[..., :T] = t_[..., :T]
            t_ = t2
        ctx.scale = float(scale)
        loss = torch.zeros(1, device=pred.device, dtype=torch.float32)
        G = torch.empty(s_.shape, device=pred.device, dtype=torch.float16)
        base = ctx.scale * loss_mult
        ctx.boost = E._pow2_near((1.0 if mode == 0 else 0.25 * T) / base)
        K.attn_map_loss(s_, t_, valid_s, valid_t, G, loss, BH // H, T, H, mode, loss_mult, base * ctx.boost)
        ctx.save_for_backward(G)
        ctx.T = T
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (G,) = ctx.saved_tensors
        w = float(g)  # the loss weight (x any caller rescaling)
        _PENDING_SCALE[0] = ctx.scale
        _MAP_GRAD_SCALED[0] = True
        return (G[..., :ctx.T].float() * (w / ctx.boost),) + (None,) * 7


_MAP_GRAD_SCALED = [False]  # the map gradients about to reach _StudentFn.backward already carry the loss scale
_FEAT_GRAD_SCALED = [False]  # the features gradient about to reach _StudentFn.backward already carries the loss scale


def _run_loss(args, rec, dpred, scale):
    """Launch the loss kernel; returns the per-layer cosine term (or None when sim_weight == 0)."""
    from . import kernels as K
    preds, tgt, weights, n, B, Tq, Tt, D, loss_type, rec_w, sim_w = args
    if not sim_w:
        K.distill_loss(preds, tgt, weights, rec, dpred, n, B, Tq, Tt, D, loss_type, scale * rec_w)
        return None
    sim = torch.zeros(n, device=preds.device, dtype=torch.float32)
    K.distill_loss_sim(preds, tgt, weights, rec, sim, dpred, n, B, Tq, Tt, D, loss_type, scale * rec_w, scale * sim_w)
    return sim
