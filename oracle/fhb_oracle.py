"""CPU oracle for the FitHuBERT distillation step.  TEST INFRASTRUCTURE ONLY.

A plain-PyTorch fp32, *functional* restatement (state-dict in, tensors out) of the
reference arithmetic on the hot path.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product
package fithubert_b200/ never does (it fails loudly without its CUDA library).

Parity pin: the reference ships no tests or golden vectors, so
this oracle is pinned against the *reference's own files executed unmodified* through
oracle/fairseq_stub (oracle/gen_golden.py -> tests/golden/*.pt, and the live check
tests/test_oracle_pins.py::test_oracle_matches_the_reference_classes_live when
/root/reference is present), and the teacher additionally against
torchaudio.models.hubert_base (tests/test_oracle_pins.py).  The optimizer (s3prl, source not
available anywhere in this image) is restated from its published algorithm:
"parity unpinned" for that one function.  train.py cannot be IMPORTED (Lightning and
s3prl at module level), but `W2V2Distil.calculate_loss` (train.py:236-405) and
`rtrn_attn_forward` (utils/utils.py:190-258) are self-contained: oracle/ref_extract.py
compiles them straight out of the reference's unmodified source text (ast), and
  * tests/test_oracle_pins.py::test_loss_restatements_match_the_reference_calculate_loss_live
    pins distill_loss, distill_loss_sim, cnn_feature_loss, attn_map_loss and
    value_relation_loss against that very method (live, where /root/reference exists);
  * the attn_* fixtures under tests/golden/ hold what those two reference functions
    produced (loss terms, total, gradients), so they also travel to the GPU box.

Every function cites the reference file:line it follows (paths relative to the
reference root; [EXT] = fairseq @1b61bbad / s3prl @185e4b06, not vendored).
"""
from __future__ import annotations

import ast
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]

FITHUBERT_CONV = "[(128, 10, 5)] + [(256, 1, 1)] + [(256, 3, 2)] * 4 + [(512, 1, 1)] + [(512, 2, 2)] * 2"
HUBERT_CONV = "[(512,10,5)] + [(512,3,2)] * 4 + [(512,2,2)] * 2"


# --------------------------------------------------------------------------- config
def parse_conv_layers(spec) -> List[Tuple[int, int, int]]:
    """The reference eval()s this string (modules/model.py:267,384); we evaluate the
    same list arithmetic ('+' and '*' on list literals) without eval."""
    if not isinstance(spec, str):
        return [tuple(x) for x in spec]

    def ev(node):
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.BinOp) and isinstance(node.op, ast.Add):
            return ev(node.left) + ev(node.right)
        if isinstance(node, ast.BinOp) and isinstance(node.op, ast.Mult):
            return ev(node.left) * ev(node.right)
        return ast.literal_eval(node)

    return [tuple(int(v) for v in t) for t in ev(ast.parse(spec, mode="eval"))]


def student_config(**over) -> dict:
    """Defaults = data/conf/fithubert.yaml:27-89 (the `distiller:` group)."""
    cfg = dict(
        conv_feature_layers=FITHUBERT_CONV, conv_pos=128, conv_pos_groups=16,
        encoder_layers=12, encoder_embed_dim=480, encoder_ffn_embed_dim=480,
        encoder_attention_heads=12, pred_head_final_dim=768, tr_reduce_factor=2,
        enable_tr_layer=True, layerwise_proj=True,
    )
    cfg.update(over)
    return cfg


def teacher_config(**over) -> dict:
    """HuBERT-Base / wav2vec2-Base ([EXT] fairseq HubertConfig / Wav2Vec2Config defaults)."""
    cfg = dict(
        conv_feature_layers=HUBERT_CONV, conv_pos=128, conv_pos_groups=16,
        encoder_layers=12, encoder_embed_dim=768, encoder_ffn_embed_dim=3072,
        encoder_attention_heads=12, kind="hubert",
    )
    cfg.update(over)
    return cfg


# --------------------------------------------------------------------------- masks (integer, bit-exact)
def conv_out_lengths(lengths: Tensor, conv_layers) -> Tensor:
    """modules/model.py:376-391: per layer floor((len - k)/s + 1) in *float* arithmetic,
    then .long().  Restated with the same float ops so ties round identically."""
    x = lengths
    for (_, k, s) in conv_layers:
        x = torch.floor((x - k) / s + 1)
    return x.to(torch.long)


def mask_m1(padding_mask: Optional[Tensor], T: int, conv_layers) -> Optional[Tensor]:
    """Student / wav2vec2 frame mask, modules/model.py:449-472.  Returns None when no
    sample is padded.  Frame j is padding iff j >= conv_out_len(len_i)."""
    if padding_mask is None or not bool(padding_mask.any()):
        return None
    input_lengths = (1 - padding_mask.long()).sum(-1)
    out_len = conv_out_lengths(input_lengths, conv_layers)
    m = torch.zeros(padding_mask.shape[0], T, dtype=torch.float32)
    m[torch.arange(m.shape[0]), out_len - 1] = 1
    return (1 - m.flip([-1]).cumsum(-1).flip([-1])).bool()


def mask_m2(mask: Optional[Tensor], factor: int = 2) -> Optional[Tensor]:
    """Time-reduction of the mask, modules/module.py:324-328: pairs, drop odd tail, any()."""
    if mask is None:
        return None
    sp = mask.split(factor, 1)
    if mask.shape[-1] % factor != 0:
        sp = sp[:-1]
    return torch.stack(sp).any(-1).transpose(0, 1)


def mask_m3(padding_mask: Optional[Tensor], T: int) -> Optional[Tensor]:
    """[EXT] HubertModel.forward_padding_mask: crop to a multiple of T, view
    [B, T, L//T], all(-1).  Applied whenever a mask is passed (even all-False)."""
    if padding_mask is None:
        return None
    extra = padding_mask.size(1) % T
    if extra > 0:
        padding_mask = padding_mask[:, :-extra]
    return padding_mask.view(padding_mask.size(0), T, -1).all(-1)


def valid_lengths(mask: Optional[Tensor], T: int, B: int) -> Tensor:
    """Every mask above is a suffix mask; its per-sample valid length describes it fully."""
    if mask is None:
        return torch.full((B,), T, dtype=torch.long)
    return (~mask).long().sum(-1)


# --------------------------------------------------------------------------- blocks
def conv_extractor(sd: State, prefix: str, x: Tensor, conv_layers) -> Tensor:
    """modules/module.py:94-102 with mode='default', conv_bias=False: layer 0 is
    conv -> Fp32GroupNorm(C, C) -> GELU (:65-71), the others conv -> GELU (:72-73)."""
    x = x.unsqueeze(1)
    for i, (c, k, s) in enumerate(conv_layers):
        x = F.conv1d(x, sd[f"{prefix}conv_layers.{i}.0.weight"], None, stride=s)
        if i == 0:
            x = F.group_norm(x.float(), c, sd[f"{prefix}conv_layers.0.2.weight"],
                             sd[f"{prefix}conv_layers.0.2.bias"], 1e-5)
        x = F.gelu(x)
    return x  # [B, C, T]


def pos_conv(sd: State, prefix: str, x_btc: Tensor, k: int, groups: int) -> Tensor:
    """modules/module.py:186-200,276-278: weight_norm(dim=2) grouped conv, pad k//2,
    SamePad drops the last frame (k even), GELU.  Returns the conv branch [B,T,C]."""
    v, g = sd[f"{prefix}weight_v"], sd[f"{prefix}weight_g"]
    w = g * v / v.norm(dim=(0, 1), keepdim=True)
    y = F.conv1d(x_btc.transpose(1, 2), w, sd[f"{prefix}bias"], padding=k // 2, groups=groups)
    if k % 2 == 0:
        y = y[:, :, :-1]
    return F.gelu(y).transpose(1, 2)


def _mha_heads(sd: State, p: str, x_tbc: Tensor, key_mask: Optional[Tensor], H: int):
    """q (scaled after bias), k, v as [B*H, T, d] and the masked logits [B*H, T, T] of fairseq's manual path."""
    T, B, E = x_tbc.shape
    d = E // H
    q = F.linear(x_tbc, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"]) * d ** -0.5
    k = F.linear(x_tbc, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])
    v = F.linear(x_tbc, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])
    q, k, v = (t.reshape(T, B * H, d).transpose(0, 1) for t in (q, k, v))
    w = torch.bmm(q, k.transpose(1, 2))
    if key_mask is not None:
        w = w.view(B, H, T, T).masked_fill(key_mask[:, None, None, :], float("-inf")).view(B * H, T, T)
    return w, v


def mha(sd: State, p: str, x_tbc: Tensor, key_mask: Optional[Tensor], H: int, return_attn: bool = False):
    """[EXT] fairseq MultiheadAttention manual path: q scaled after
    bias, -inf on padded keys, fp32 softmax, out_proj.
    return_attn: the attention-map recipe (utils/utils.py:190-232, `rtrn_attn_forward` bound over every layer's forward
    when attn_loss_weight > 0, train.py:64-77): fairseq's `before_softmax=True` hands back the masked logits and the
    value heads, the layer finishes the attention itself and also returns (logits, v_rel) with
    v_rel = bmm(v * scaling, v^T) (un-masked, :229)."""
    T, B, E = x_tbc.shape
    d = E // H
    w, v = _mha_heads(sd, p, x_tbc, key_mask, H)
    a = torch.bmm(torch.softmax(w.float(), -1), v)
    a = a.transpose(0, 1).reshape(T, B, E)
    out = F.linear(a, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])
    if return_attn:
        return out, (w, torch.bmm(v * d ** -0.5, v.transpose(1, 2)))
    return out


def encoder_layer(sd: State, p: str, x: Tensor, key_mask, H: int, return_attn: bool = False):
    """modules/module.py:557-580, post-LN branch, dropout = identity.  return_attn: utils/utils.py:233-258 (the same layer
    with the logits / value-relation pair in place of the `None` attention weights)."""
    E = x.shape[-1]
    a = mha(sd, p + "self_attn.", x, key_mask, H, return_attn)
    attn = None
    if return_attn:
        a, attn = a
    x = x + a
    x = F.layer_norm(x, (E,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"], 1e-5)
    h = F.gelu(F.linear(x, sd[p + "fc1.weight"], sd[p + "fc1.bias"]))
    lr = F.linear(h, sd[p + "fc2.weight"], sd[p + "fc2.bias"])
    x = F.layer_norm(x + lr, (E,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], 1e-5)
    if return_attn:
        return x, attn, lr
    return x, lr


def encoder_prologue(sd: State, x: Tensor, mask: Optional[Tensor], cfg: dict) -> Tensor:
    """modules/module.py:273-281: zero padded frames, x + pos_conv(x), LayerNorm."""
    if mask is not None:
        x = x.masked_fill(mask.unsqueeze(-1), 0.0)
    x = x + pos_conv(sd, "encoder.pos_conv.0.", x, cfg["conv_pos"], cfg["conv_pos_groups"])
    E = x.shape[-1]
    return F.layer_norm(x, (E,), sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"], 1e-5)


# --------------------------------------------------------------------------- student
def grad_multiply(x: Tensor, scale: float) -> Tensor:
    """[EXT] fairseq GradMultiply (modules/model.py:428-431): identity forward, gradient x scale."""
    if scale == 1.0:
        return x
    return x * scale + (x * (1.0 - scale)).detach()


def student_forward(sd: State, cfg: dict, source: Tensor, padding_mask: Optional[Tensor] = None,
                    heads: bool = True, return_attn: bool = False) -> dict:
    """CustomStudentModel.forward, modules/model.py:420-552 (n_mels=0, transformer, dropout identity) for the two
    shipped recipes:
      * fithubert.yaml: layerwise_proj=True, conv1d TR layer at index 0 -> 12 LayerWiseProjHeads, x = last projection;
      * ex.yaml: layerwise_proj=False, enable_tr_layer=False, feature_grad_mult < 1 -> DistilHuBERT head
        Linear -> GELU -> SplitLinear on the last layer (:504-518, modules/module.py:585-619), projections a
        [B, N, T, D] tensor, x = the encoder output.
    heads=False mirrors the state after _disable_projection_heads() (:393-399,500-502).
    return_attn=True: every layer runs `rtrn_attn_forward` (train.py:64-77 with attn_loss_weight > 0): layer_results[i] =
    (x, (attn_logits, v_rel), layer_result) (modules/module.py:331-334 appends (x, z, lr) with z the layer's second output)."""
    conv_layers = parse_conv_layers(cfg["conv_feature_layers"])
    H = cfg["encoder_attention_heads"]
    feats = conv_extractor(sd, "feature_extractor.", source, conv_layers)
    feats = grad_multiply(feats, float(cfg.get("feature_grad_mult", 1.0))).transpose(1, 2)
    B, T, C = feats.shape
    feats = F.layer_norm(feats, (C,), sd["layer_norm.weight"], sd["layer_norm.bias"], 1e-5)
    mask = mask_m1(padding_mask, T, conv_layers)
    feats = F.linear(feats, sd["post_extract_proj.weight"], sd["post_extract_proj.bias"])
    # modules/model.py:483,489: `features_to_distill = features` is an ALIAS, and with dropout_input an identity (eval
    # mode, or p = 0: nn.Dropout then returns its input) the encoder's in-place index_put(x, padding_mask, 0)
    # (modules/module.py:273-274) writes through it: the returned `features` have their padded frames zeroed.  (In
    # training mode with p > 0 dropout makes a copy and `features` stay un-zeroed; this oracle restates dropout = identity.)
    features_to_distill = feats if mask is None else feats.masked_fill(mask.unsqueeze(-1), 0.0)
    if "cnn_proj_head.1.weight" in sd:
        # modules/model.py:304-310,486-487: Sequential(GELU, Linear) applied BEFORE the encoder call - a new tensor, so
        # the encoder's in-place zeroing never reaches it (padded frames stay as computed)
        features_to_distill = F.linear(F.gelu(feats), sd["cnn_proj_head.1.weight"], sd["cnn_proj_head.1.bias"])
    x = encoder_prologue(sd, feats, mask, cfg).transpose(0, 1)  # [T,B,C]
    tr = bool(cfg.get("enable_tr_layer", True))
    tr_layer_results = []
    rmask = mask
    if tr:
        # time-reduction conv, modules/module.py:317-321 (drops the last frame when T is odd)
        x = F.conv1d(x.permute(1, 2, 0), sd["encoder.layers.0.weight"], sd["encoder.layers.0.bias"],
                     stride=cfg["tr_reduce_factor"]).permute(2, 0, 1)
        tr_layer_results = [x]
        rmask = mask_m2(mask, cfg["tr_reduce_factor"])
    layer_results = []
    off = 1 if tr else 0
    if return_attn and tr:
        # train.py:70-77 calls `layer.self_attn._set_skip_embed_dim_check()` on every entry of encoder.layers; entry 0 is
        # the time-reduction nn.Conv1d (modules/module.py:230-236)
        raise AttributeError("'Conv1d' object has no attribute 'self_attn'")
    for i in range(cfg["encoder_layers"]):
        if return_attn:
            x, attn, lr = encoder_layer(sd, f"encoder.layers.{i + off}.", x, rmask, H, True)
            layer_results.append((x, attn, lr))
            continue
        x, lr = encoder_layer(sd, f"encoder.layers.{i + off}.", x, rmask, H)
        layer_results.append((x, None, lr))

    if not cfg.get("layerwise_proj", True):
        out = x.transpose(0, 1)  # [B, T, E]
        if tr:
            # shared upsampler (modules/model.py:341-348,402-404,504-505): ConvTranspose1d(k = s = tr_reduce_factor) on the
            # encoder output, in front of the head; `x` of the result is the upsampled tensor
            out = F.conv_transpose1d(out.transpose(1, 2), sd["upsampler.weight"], sd["upsampler.bias"],
                                     stride=cfg["tr_reduce_factor"]).transpose(1, 2)
        projections = None
        if heads:
            n = sd["proj_head.2.weight"].shape[0]
            h = F.gelu(F.linear(out, sd["proj_head.0.weight"], sd["proj_head.0.bias"]))
            h = h.reshape(B, out.shape[1], n, 1, -1)
            pred = torch.einsum("...klm,kmn->...kln", h, sd["proj_head.2.weight"]).squeeze(3) + sd["proj_head.2.bias"]
            projections = pred.reshape(B, out.shape[1], n, -1).permute(0, 2, 1, 3)  # B x N x T x D
        return {"x": out, "padding_mask": mask, "features": features_to_distill,
                "layer_results": layer_results, "tr_layer_results": tr_layer_results, "projections": projections}

    def head(i, h_tbc):  # LayerWiseProjHead.forward, modules/module.py:649-661
        y = h_tbc.transpose(0, 1)
        if tr:  # `self.upsampler` only exists with a time-reduction layer (:633-640)
            y = F.conv_transpose1d(h_tbc.permute(1, 2, 0), sd[f"proj_head.{i}.upsampler.weight"],
                                   sd[f"proj_head.{i}.upsampler.bias"], stride=cfg["tr_reduce_factor"]).transpose(1, 2)
        return F.linear(y, sd[f"proj_head.{i}.lin_proj.weight"], sd[f"proj_head.{i}.lin_proj.bias"])

    if heads:
        projections = [head(i, layer_results[i][0]) for i in range(cfg["encoder_layers"])]
        out = projections[-1]
    else:
        projections = None
        out = head(cfg["encoder_layers"] - 1, x)
    return {"x": out, "padding_mask": mask, "features": features_to_distill,
            "layer_results": layer_results, "tr_layer_results": tr_layer_results,
            "projections": projections}


# --------------------------------------------------------------------------- teacher
def teacher_forward(sd: State, cfg: dict, source: Tensor, padding_mask: Optional[Tensor] = None,
                    return_attn: bool = False) -> dict:
    """TeacherWrapper.extract_features (utils/utils.py:80-99) around [EXT]
    HubertModel / Wav2Vec2Model.extract_features(mask=None), eval mode."""
    conv_layers = parse_conv_layers(cfg["conv_feature_layers"])
    H = cfg["encoder_attention_heads"]
    feats = conv_extractor(sd, "feature_extractor.", source, conv_layers).transpose(1, 2)
    B, T, C = feats.shape
    feats = F.layer_norm(feats, (C,), sd["layer_norm.weight"], sd["layer_norm.bias"], 1e-5)
    if cfg.get("kind", "hubert") == "hubert":
        mask = mask_m3(padding_mask, T)
    else:
        mask = mask_m1(padding_mask, T, conv_layers)
    feats = F.linear(feats, sd["post_extract_proj.weight"], sd["post_extract_proj.bias"])
    x = encoder_prologue(sd, feats, mask, cfg).transpose(0, 1)
    layer_results = []
    for i in range(cfg["encoder_layers"]):
        if return_attn:  # the hook captures (x, ((attn_logits, v_rel), layer_result)), utils/utils.py:65-78,258
            x, attn, lr = encoder_layer(sd, f"encoder.layers.{i}.", x, mask, H, True)
            layer_results.append((x, (attn, lr)))
            continue
        x, lr = encoder_layer(sd, f"encoder.layers.{i}.", x, mask, H)
        layer_results.append((x, (None, lr)))
    return {"layer_results": layer_results, "x": layer_results[-1][0].transpose(0, 1),
            "features": [feats], "padding_mask": mask}


# --------------------------------------------------------------------------- loss
def layer_weights(n_layers: int, random_layer_weight: float = 0.1) -> List[float]:
    """train.py:90-91,290-291 with distil_random_layer = n_layers-1: every lower layer
    is selected (a permutation), each weighted random_layer_weight; last layer 1.0."""
    return [random_layer_weight] * (n_layers - 1) + [1.0]


def distill_loss(projections: Sequence[Tensor], teacher_layer_results, weights: Sequence[float],
                 rec_loss_type: str = "mse") -> Tuple[Tensor, Tensor]:
    """W2V2Distil.calculate_loss rec branch, train.py:250-267,282-293,372-378
    (rec_loss_weight 1, all other weights 0).  Returns (total, per-layer weighted means)
    in *layer order* (the reference logs them in permuted order, train.py:319-321)."""
    pred = torch.stack(list(projections), dim=1)  # B x N x T' x D
    tgt = torch.stack([lr[0].transpose(0, 1) for lr in teacher_layer_results], dim=1)
    tgt = tgt.narrow(2, 0, pred.shape[2])
    if rec_loss_type == "mse":
        e = F.mse_loss(pred.float(), tgt.float(), reduction="none")
    elif rec_loss_type == "l1":
        e = F.l1_loss(pred.float(), tgt.float(), reduction="none")
    else:
        raise NotImplementedError("rec_loss_type must be one of 'l1', 'mse'.")
    w = torch.tensor(list(weights), dtype=e.dtype).view(1, -1, 1, 1)
    per_layer = (e * w).mean((0, 2, 3))
    return per_layer.sum(), per_layer


def distill_loss_sim(projections: Sequence[Tensor], teacher_layer_results, pred_layer_id: Sequence[int],
                     rec_loss_type: str = "l1", rec_loss_weight: float = 1.0,
                     sim_loss_weight: float = 1.0) -> Tuple[Tensor, Tensor, Tensor]:
    """W2V2Distil.calculate_loss with the cosine term, `distil_random_layer == 0` branch:
    train.py:268-281 (pred_layer_id selection), :282-288 (rec), :294-297 (plain means), :302-303,309-312 (cosine),
    :316,322-324 (per-layer log = rec + sim), :372-378 (total).  The random-layer branch (:304-306) calls
    `.mean((0, 2, 3))` on the 3-D cosine loss and cannot run, so it has no restatement.
    Returns (total, per-layer rec means, per-layer sim means), the last two in pred_layer_id order."""
    ids = list(pred_layer_id)
    pred = torch.stack([projections[i] for i in ids], dim=1).float()  # B x N x T' x D
    tgt = torch.stack([teacher_layer_results[i][0].transpose(0, 1) for i in ids], dim=1).float()
    tgt = tgt.narrow(2, 0, pred.shape[2])
    if rec_loss_type == "l1":
        rec = F.l1_loss(pred, tgt, reduction="none")
    elif rec_loss_type == "mse":
        rec = F.mse_loss(pred, tgt, reduction="none")
    else:
        raise NotImplementedError("rec_loss_type must be one of 'l1', 'mse'.")
    rec_layer = rec.mean((0, 2, 3))
    sim = -F.logsigmoid(F.cosine_similarity(pred, tgt, dim=-1))
    sim_layer = sim.mean((0, 2))
    total = rec_loss_weight * rec.mean() + sim_loss_weight * sim.mean()
    return total, rec_layer, sim_layer


# --------------------------------------------------------------------------- optimizer (parity unpinned)
def lr_schedule(step: int, total_steps: int, warmup: float) -> float:
    """[EXT] s3prl warmup_linear: x/w for x < w, else max((x-1)/(w-1), 0)."""
    x = step / total_steps
    return x / warmup if x < warmup else max((x - 1.0) / (warmup - 1.0), 0.0)


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
               betas=(0.9, 0.98), eps=1e-6, weight_decay=1e-6, mode: str = "s3prl") -> None:
    """In-place Adam step.  mode='s3prl': SURVEY App. B.3 ([EXT] Lamb(adam=True,
    correct_bias=True)); eps on the un-corrected sqrt(v), decay inside the corrected
    step.  mode='torch': torch.optim.AdamW semantics.  `step` is 1-based."""
    b1, b2 = betas
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    if mode == "s3prl":
        step_size = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
        p.add_(m / (v.sqrt() + eps) + weight_decay * p, alpha=-step_size)
    else:
        p.mul_(1 - lr * weight_decay)
        denom = (v.sqrt() / math.sqrt(1 - b2 ** step)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))


# --------------------------------------------------------------------------- init
def init_encoder_state(cfg: dict, gen: torch.Generator, n_layers_offset: int) -> State:
    E, Fd, k, G = cfg["encoder_embed_dim"], cfg["encoder_ffn_embed_dim"], cfg["conv_pos"], cfg["conv_pos_groups"]
    sd: State = {}

    def normal(*shape, std):
        return torch.empty(*shape).normal_(0.0, std, generator=gen)

    # modules/module.py:194-199
    wv = normal(E, E // G, k, std=math.sqrt(4.0 / (k * E)))
    sd["encoder.pos_conv.0.bias"] = torch.zeros(E)
    sd["encoder.pos_conv.0.weight_g"] = wv.norm(dim=(0, 1), keepdim=True)
    sd["encoder.pos_conv.0.weight_v"] = wv
    for i in range(cfg["encoder_layers"]):
        p = f"encoder.layers.{i + n_layers_offset}."
        for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):  # init_bert_params
            sd[p + f"self_attn.{nm}.weight"] = normal(E, E, std=0.02)
            sd[p + f"self_attn.{nm}.bias"] = torch.zeros(E)
        sd[p + "self_attn_layer_norm.weight"] = torch.ones(E)
        sd[p + "self_attn_layer_norm.bias"] = torch.zeros(E)
        sd[p + "fc1.weight"] = normal(Fd, E, std=0.02)
        sd[p + "fc1.bias"] = torch.zeros(Fd)
        sd[p + "fc2.weight"] = normal(E, Fd, std=0.02)
        sd[p + "fc2.bias"] = torch.zeros(E)
        sd[p + "final_layer_norm.weight"] = torch.ones(E)
        sd[p + "final_layer_norm.bias"] = torch.zeros(E)
    sd["encoder.layer_norm.weight"] = torch.ones(E)
    sd["encoder.layer_norm.bias"] = torch.zeros(E)
    return sd


def _init_frontend(cfg: dict, gen: torch.Generator) -> State:
    sd: State = {}
    cin = 1
    conv_layers = parse_conv_layers(cfg["conv_feature_layers"])
    for i, (c, k, s) in enumerate(conv_layers):  # kaiming_normal_, modules/module.py:47
        sd[f"feature_extractor.conv_layers.{i}.0.weight"] = torch.empty(c, cin, k).normal_(
            0.0, math.sqrt(2.0 / (cin * k)), generator=gen)
        cin = c
    c0 = conv_layers[0][0]
    sd["feature_extractor.conv_layers.0.2.weight"] = torch.ones(c0)
    sd["feature_extractor.conv_layers.0.2.bias"] = torch.zeros(c0)
    sd["layer_norm.weight"] = torch.ones(cin)
    sd["layer_norm.bias"] = torch.zeros(cin)
    E = cfg["encoder_embed_dim"]
    b = 1 / math.sqrt(cin)  # nn.Linear default init
    sd["post_extract_proj.weight"] = torch.empty(E, cin).uniform_(-b, b, generator=gen)
    sd["post_extract_proj.bias"] = torch.empty(E).uniform_(-b, b, generator=gen)
    return sd


def init_student_state(cfg: dict, seed: int = 0, perturb: bool = False) -> State:
    """Random student weights with the reference's parameter names
    and init distributions.  perturb=True randomises the biases and norm
    affine parameters too (they are 0/1 at init) so parity tests exercise them."""
    gen = torch.Generator().manual_seed(seed)
    sd = _init_frontend(cfg, gen)
    tr = bool(cfg.get("enable_tr_layer", True))
    sd.update(init_encoder_state(cfg, gen, 1 if tr else 0))
    E, D = cfg["encoder_embed_dim"], cfg["pred_head_final_dim"]
    f = cfg["tr_reduce_factor"]

    def uni(*shape, bound):
        return torch.empty(*shape).uniform_(-bound, bound, generator=gen)

    bt = 1 / math.sqrt(E * f)
    if tr:
        sd["encoder.layers.0.weight"] = uni(E, E, f, bound=bt)
        sd["encoder.layers.0.bias"] = uni(E, bound=bt)
        sd["upsampler.weight"] = uni(E, E, f, bound=bt)  # dead when layerwise_proj
        sd["upsampler.bias"] = uni(E, bound=bt)
    if not cfg.get("layerwise_proj", True):
        # DistilHuBERT head (modules/model.py:362-368, modules/module.py:585-604): Linear(E, inter * N), SplitLinear
        n = len(cfg["pred_layer_id"])
        inter = cfg.get("pred_head_inter_dim", 0) or E
        sd["proj_head.0.weight"] = uni(inter * n, E, bound=1 / math.sqrt(E))
        sd["proj_head.0.bias"] = uni(inter * n, bound=1 / math.sqrt(E))
        sd["proj_head.2.weight"] = uni(n, inter, D, bound=inter ** -0.5)
        sd["proj_head.2.bias"] = uni(1, 1, n, D, bound=inter ** -0.5)
    else:
        for i in range(cfg["encoder_layers"]):
            if tr:
                sd[f"proj_head.{i}.upsampler.weight"] = uni(E, E, f, bound=bt)
                sd[f"proj_head.{i}.upsampler.bias"] = uni(E, bound=bt)
            sd[f"proj_head.{i}.lin_proj.weight"] = uni(D, E, bound=1 / math.sqrt(E))
            sd[f"proj_head.{i}.lin_proj.bias"] = uni(D, bound=1 / math.sqrt(E))
    if cfg.get("_cnn_weight", 0) > 0 and D != E:  # cnn_proj_head = Sequential(GELU, Linear(E, D)), modules/model.py:304-310
        sd["cnn_proj_head.1.weight"] = uni(D, E, bound=1 / math.sqrt(E))
        sd["cnn_proj_head.1.bias"] = uni(D, bound=1 / math.sqrt(E))
    if perturb:
        _perturb(sd, gen)
    return sd


def cnn_feature_loss(student_features: Tensor, teacher_features: Tensor) -> Tensor:
    """W2V2Distil.calculate_loss CNN branch, train.py:241-246: plain L1 mean between the student's `features`
    (features_to_distill) and the teacher's post_extract_proj output."""
    return F.l1_loss(student_features.float(), teacher_features.float(), reduction="none").mean()


def attn_map_loss(pred: Tensor, target: Tensor, loss_type: str = "kldiv", nan_like_reference: bool = True) -> Tensor:
    """Attention distribution transfer loss, train.py:327-353.  pred / target: the student's and the teacher's LAST-layer
    attention logits [B*H, T, T], -inf at padded keys (each side's own mask rule).
      mse   (:331-341): squared differences; inf (one side masked) and nan (both masked) entries are zeroed in place and
            left out of the mean - the count is taken per (b*h, key) column (`torch.any(.., 1)` reduces the QUERY axis);
      kldiv (:342-349): F.kl_div(log_softmax(pred), softmax(target), 'none'), inf entries (student-masked key with
            teacher mass) zeroed, sum over keys, mean over all B*H*T query rows (padded queries included).
    Reference quirk: in the kldiv branch a key masked on BOTH sides gives 0 * -inf = nan, and nan is NOT patched there
    (only inf is, :348): with a padded batch the reference's kldiv loss is nan.  nan_like_reference=False restates the
    evidently intended value (such keys contribute nothing), which is what the CUDA path computes."""
    pred, target = pred.float(), target.float()
    if loss_type == "mse":
        loss = F.mse_loss(pred, target, reduction="none")
        inf_count = torch.any(loss.isinf(), 1).count_nonzero() * loss.size(-1)
        nan_count = torch.any(loss.isnan(), 1).count_nonzero() * loss.size(-1)
        loss = loss.masked_fill(loss.isinf() | loss.isnan(), 0.0)
        return loss.sum() / (loss.numel() - inf_count - nan_count)
    if loss_type == "kldiv":
        loss = F.kl_div(F.log_softmax(pred, dim=-1), F.softmax(target, dim=-1), reduction="none")
        loss = loss.masked_fill(loss.isinf(), 0.0)
        if not nan_like_reference:
            loss = loss.masked_fill(loss.isnan(), 0.0)
        return loss.sum(dim=-1).mean()
    raise NotImplementedError("attn_loss_type must be one of 'mse', 'kldiv'.")


def value_relation_loss(pred: Tensor, target: Tensor) -> Tensor:
    """Value relation transfer loss, train.py:355-368: KL(softmax(teacher v_rel) || softmax(student v_rel)) per row of
    the un-masked [B*H, T, T] value-relation maps, summed over the last axis, mean over rows."""
    loss = F.kl_div(F.log_softmax(pred.float(), dim=-1), F.softmax(target.float(), dim=-1), reduction="none")
    return loss.sum(dim=-1).mean()


def init_teacher_state(cfg: dict, seed: int = 1, perturb: bool = False) -> State:
    gen = torch.Generator().manual_seed(seed)
    sd = _init_frontend(cfg, gen)
    sd.update(init_encoder_state(cfg, gen, 0))
    if perturb:
        _perturb(sd, gen)
    return sd


def _perturb(sd: State, gen: torch.Generator) -> None:
    for k, t in sd.items():
        if k.endswith(".bias") or "layer_norm.weight" in k or k.endswith("conv_layers.0.2.weight"):
            t.add_(torch.empty_like(t).normal_(0.0, 0.05, generator=gen))


def synth_batch(B: int, Lmax: int, lengths: Sequence[int], seed: int = 1234) -> Tuple[Tensor, Tensor]:
    """Input contract of utils/dataset.py:63-74: fp32 [B, L] zero-padded waveforms and a
    bool padding mask (True = pad).  Waveform = 0.1*randn."""
    gen = torch.Generator().manual_seed(seed)
    x = 0.1 * torch.randn(B, Lmax, generator=gen)
    ar = torch.arange(Lmax).unsqueeze(0)
    pm = ~(ar < torch.tensor(list(lengths)).unsqueeze(1))
    x = x.masked_fill(pm, 0.0)
    return x, pm
