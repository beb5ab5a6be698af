"""Reference-facing model classes.  Same constructor / forward signatures, attribute names, returned
dict keys and state-dict keys as the reference (modules/model.py:253-588, modules/module.py,
utils/utils.py:51-99), but the modules below are parameter
CONTAINERS only: all arithmetic runs in engine.py on hand-written sm_100a kernels.  Calling them on
CPU tensors raises - there is no fallback path."""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn

from . import engine as E
from . import kernels as K
from . import lib as L
from .config import CustomStudentModelConfig, parse_int_list, parse_layer_spec

bf16 = torch.bfloat16


# --------------------------------------------------------------------------- parameter containers
class _Weight(nn.Module):
    def __init__(self, *shape, bias_shape=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*shape))
        if bias_shape is not None:
            self.bias = nn.Parameter(torch.zeros(*bias_shape))

    def forward(self, *a, **k):  # pragma: no cover
        raise L.FhbError("parameter container: the computation runs in fithubert_b200.engine")


class _Affine(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))


class ConvFeatureExtractionModel(nn.Module):
    """conv_layers.{i}.0.weight [Cout, Cin, k]; conv_layers.0.2.{weight,bias} (GroupNorm).
    Init: kaiming_normal_ (reference modules/module.py:45-48)."""

    def __init__(self, conv_layers, dropout=0.0, mode="default", conv_bias=False):
        super().__init__()
        assert mode in {"default", "layer_norm"}
        self.conv_layers = nn.ModuleList()
        cin = 1
        for i, cl in enumerate(conv_layers):
            assert len(cl) == 3, "invalid conv definition: " + str(cl)
            dim, k, stride = cl
            conv = _Weight(dim, cin, k)
            nn.init.kaiming_normal_(conv.weight)
            mods = [conv, nn.Identity()]
            if i == 0:
                mods.append(_Affine(dim))
            mods.append(nn.Identity())
            self.conv_layers.append(nn.Sequential(*mods))
            cin = dim


class _PosConv(nn.Module):
    """Old-style weight_norm(dim=2) parameters: bias, weight_g [1,1,k], weight_v [E, E/G, k]
    (reference modules/module.py:186-200)."""

    def __init__(self, e, k, g):
        super().__init__()
        v = torch.empty(e, e // g, k)
        nn.init.normal_(v, mean=0, std=math.sqrt(4.0 / (k * e)))
        self.bias = nn.Parameter(torch.zeros(e))
        self.weight_g = nn.Parameter(v.norm(dim=(0, 1), keepdim=True).clone())
        self.weight_v = nn.Parameter(v)


class _SelfAttention(nn.Module):
    def __init__(self, e, heads):
        super().__init__()
        self.embed_dim, self.num_heads = e, heads
        for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
            lin = nn.Linear(e, e, bias=True)
            nn.init.normal_(lin.weight, 0.0, 0.02)  # init_bert_params
            nn.init.zeros_(lin.bias)
            setattr(self, nm, lin)


class TransformerSentenceEncoderLayer(nn.Module):
    def __init__(self, embedding_dim=768, ffn_embedding_dim=3072, num_attention_heads=8, **_):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.self_attn = _SelfAttention(embedding_dim, num_attention_heads)
        self.self_attn_layer_norm = nn.LayerNorm(embedding_dim)
        self.fc1 = nn.Linear(embedding_dim, ffn_embedding_dim)
        self.fc2 = nn.Linear(ffn_embedding_dim, embedding_dim)
        self.final_layer_norm = nn.LayerNorm(embedding_dim)
        for lin in (self.fc1, self.fc2):
            nn.init.normal_(lin.weight, 0.0, 0.02)
            nn.init.zeros_(lin.bias)


class TransformerEncoder(nn.Module):
    """pos_conv.0.*, layers.{i}.* (layers[0] = time-reduction Conv1d when enabled), layer_norm.*"""

    def __init__(self, embed_dim, ffn_dim, heads, n_layers, conv_pos, conv_pos_groups, tr_layer: bool):
        super().__init__()
        self.embedding_dim = embed_dim
        self.pos_conv = nn.Sequential(_PosConv(embed_dim, conv_pos, conv_pos_groups), nn.Identity(), nn.Identity())
        self.layers = nn.ModuleList(
            [TransformerSentenceEncoderLayer(embed_dim, ffn_dim, heads) for _ in range(n_layers)])
        if tr_layer:
            tr = nn.Conv1d(embed_dim, embed_dim, kernel_size=2, stride=2)  # default torch init, as the reference
            self.layers.insert(0, tr)
        self.layer_norm = nn.LayerNorm(embed_dim)


class LayerWiseProjHead(nn.Module):
    def __init__(self, in_dim, out_dim, enable_tr_layer=True, tr_reduce_factor=2):
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        # reference modules/module.py:633-646: the upsampler only exists behind a time-reduction layer
        self.upsampler = nn.ConvTranspose1d(in_dim, in_dim, kernel_size=tr_reduce_factor, stride=tr_reduce_factor) \
            if enable_tr_layer else None
        self.lin_proj = nn.Linear(in_dim, out_dim)


class SplitLinear(nn.Module):
    """Parameter container of reference modules/module.py:585-604: weight [N, Din, Dout], bias [1, 1, N, Dout]."""

    def __init__(self, in_dim, in_split, out_dim):
        super().__init__()
        self.in_dim, self.in_split, self.out_dim = in_dim, in_split, out_dim
        self.weight = nn.Parameter(torch.empty(in_split, in_dim, out_dim).uniform_(-(in_dim ** -0.5), in_dim ** -0.5))
        self.bias = nn.Parameter(torch.empty(1, 1, in_split, out_dim).uniform_(-(in_dim ** -0.5), in_dim ** -0.5))


def _named_param_dict(module: nn.Module):
    # detach() shares the version counter with the Parameter (unlike .data), so in-place updates by any
    # optimizer / load_state_dict are seen by WeightSet.signature().  Cached on the module (walking 264 parameters
    # costs ~0.4 ms per call, four times per training step); `_apply` (.to / .cuda / .float) and structural edits
    # (_disable_projection_heads) drop the cache.
    cached = module.__dict__.get("_fhb_params")
    if cached is None:
        cached = {n: p.detach() for n, p in module.named_parameters()}
        module.__dict__["_fhb_params"] = cached
    return cached


class _ParamCacheMixin:
    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_fhb_params", None)
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.__dict__.pop("_fhb_params", None)
        return super().load_state_dict(*args, **kwargs)


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise L.FhbError(f"{what}: parameters must live on a CUDA device (no CPU fallback for the hot path)")


# --------------------------------------------------------------------------- mask rules (host integers)
def conv_out_lengths(lengths: List[int], conv_layers) -> List[int]:
    """Rule M1 (reference modules/model.py:376-391): per layer floor((n - k)/s + 1)."""
    out = []
    for n in lengths:
        for (_, k, s) in conv_layers:
            n = (n - k) // s + 1
        out.append(int(n))
    return out


def hubert_mask_lengths(lengths: List[int], Lmax: int, T: int) -> List[int]:
    """Rule M3 (fairseq HubertModel.forward_padding_mask): a frame is padding iff all its samples are;
    valid = min(T, ceil(len / chunk)), chunk = (Lmax - Lmax % T) / T."""
    chunk = (Lmax - Lmax % T) // T
    return [min(T, -(-int(n) // chunk)) for n in lengths]


def _lengths_from_mask(padding_mask: Optional[torch.Tensor]) -> Optional[List[int]]:
    """Per-sample un-padded length, or None when there is no padding (the reference then runs mask-free,
    modules/model.py:449,471-472)."""
    if padding_mask is None:
        return None
    if padding_mask.is_cuda:
        out = torch.empty(padding_mask.shape[0], device=padding_mask.device, dtype=torch.int32)
        K.mask_lengths(padding_mask.contiguous().view(torch.uint8), out)
        lengths = out.tolist()
    else:
        # host mask (what utils/dataset.py:63-74 hands over): SIMD byte count row by row, ~0.7 ms for 32 x 250k
        # samples (np.count_nonzero(axis=1) takes a strided path and costs 6-7 ms; torch's bool sum goes through int64)
        import numpy as np
        m = padding_mask.contiguous().view(torch.uint8).numpy()
        n = padding_mask.shape[1]
        lengths = [n - int(np.count_nonzero(row)) for row in m]
    if all(n == padding_mask.shape[1] for n in lengths):
        return None
    return lengths


def _frame_mask(valid: Optional[List[int]], T: int, device) -> Optional[torch.Tensor]:
    if valid is None:
        return None
    m = torch.arange(T).unsqueeze(0) >= torch.tensor(valid).unsqueeze(1)
    return m.to(device)


# --------------------------------------------------------------------------- student
class CustomStudentModel(_ParamCacheMixin, nn.Module):
    def __init__(self, cfg: CustomStudentModelConfig, teacher_model=None, **kwargs):
        super().__init__()
        self.cfg = cfg
        cfg.validate_hot_path()
        self.n_mels = cfg.n_mels
        layers = parse_layer_spec(cfg.conv_feature_layers)
        self.embed = layers[-1][0]
        self.feature_extractor = ConvFeatureExtractionModel(layers, 0.0, cfg.extractor_mode, cfg.conv_bias)
        self.post_extract_proj = nn.Linear(self.embed, cfg.encoder_embed_dim)
        # CNN distillation for encoder dimension mismatch (reference modules/model.py:304-310)
        self.cnn_proj_head = None
        if cfg.pred_head_final_dim != cfg.encoder_embed_dim and cfg._cnn_weight > 0:
            self.cnn_proj_head = nn.Sequential(nn.GELU(), nn.Linear(cfg.encoder_embed_dim, cfg.pred_head_final_dim))
        self.crop_seq_to_multiple = cfg.crop_seq_to_multiple
        self.feature_grad_mult = cfg.feature_grad_mult
        self.encoder = TransformerEncoder(cfg.encoder_embed_dim, cfg.encoder_ffn_embed_dim,
                                          cfg.encoder_attention_heads, cfg.encoder_layers, cfg.conv_pos,
                                          cfg.conv_pos_groups, cfg.enable_tr_layer)
        self.layer_norm = nn.LayerNorm(self.embed)
        self.init_conv_layers = cfg.init_conv_layers
        self.init_encoder_layers = cfg.init_encoder_layers
        self._teacher_task_agnostic = cfg._teacher_task_agnostic
        self.pred_layer_id = parse_int_list(cfg.pred_layer_id)
        self.n_tasks = len(self.pred_layer_id)
        self.enable_tr_layer = cfg.enable_tr_layer
        self.upsampler = None
        if cfg.enable_tr_layer:
            self.upsampler = nn.ConvTranspose1d(cfg.encoder_embed_dim, cfg.encoder_embed_dim,
                                                kernel_size=cfg.tr_reduce_factor, stride=cfg.tr_reduce_factor)
        self.layerwise_proj = cfg.layerwise_proj
        inter = cfg.pred_head_inter_dim if cfg.pred_head_inter_dim > 0 else cfg.encoder_embed_dim
        if cfg.layerwise_proj:
            self.proj_head = nn.ModuleList([
                LayerWiseProjHead(cfg.encoder_embed_dim, cfg.pred_head_final_dim, cfg.enable_tr_layer, cfg.tr_reduce_factor)
                for _ in range(cfg.encoder_layers)])
        else:  # DistilHuBERT style projection (reference modules/model.py:362-368)
            self.proj_head = nn.Sequential(nn.Linear(cfg.encoder_embed_dim, inter * self.n_tasks), nn.GELU(),
                                           SplitLinear(inter, self.n_tasks, cfg.pred_head_final_dim))
        self.final_proj = None
        self.specaug = None
        if self.init_conv_layers:
            assert teacher_model is not None
            self.init_from_teacher_conv(teacher_model)
        if self.init_encoder_layers > 0:
            assert teacher_model is not None
            self.init_from_teacher_enc(teacher_model, self.init_encoder_layers)
        self._geom = E.Geometry(layers, cfg.encoder_embed_dim, cfg.encoder_ffn_embed_dim, cfg.encoder_attention_heads,
                                cfg.conv_pos_groups, cfg.conv_pos, cfg.encoder_layers, cfg.pred_head_final_dim, True,
                                tr=cfg.enable_tr_layer, n_split=0 if cfg.layerwise_proj else self.n_tasks, inter=inter,
                                grad_mult=cfg.feature_grad_mult)
        self._return_attn = False  # attention-map recipe (train.py:64-77): the last layer also returns (logits, v_rel)
        self._weights = None
        self._grads = None
        self._conv_layers = layers
        self._drop_p = dict(p_input=cfg.dropout_input, p_drop=cfg.dropout, p_attn=cfg.attention_dropout,
                            p_act=cfg.activation_dropout)
        self._drop_calls = 0

    def drop_cfg(self):
        """Dropout configuration of the next training forward (None in eval mode or when every p is 0).  nn.Dropout
        semantics: active iff the module is in training mode.  Masks are counter-based (engine.DropCfg): the seed
        mixes torch's initial seed, the data-parallel rank and a per-model call counter."""
        if not self.training or max(self._drop_p.values()) <= 0.0:
            return None
        import torch.distributed as dist
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self._drop_calls += 1
        return E.DropCfg((torch.initial_seed() & 0xFFFFFFFF) ^ (rank * 0x9E3779B1) ^ (self._drop_calls * 0x7F4A7C15),
                         **self._drop_p)

    # ---- reference API -------------------------------------------------------------------
    def add_specaug(self, specaug):
        self.specaug = specaug

    def _get_feat_extract_output_lengths(self, input_lengths: torch.LongTensor):
        return torch.tensor(conv_out_lengths(input_lengths.tolist(), self._conv_layers), dtype=torch.long,
                            device=input_lengths.device)

    def _disable_projection_heads(self):
        if self.layerwise_proj:
            self.final_proj = self.proj_head[-1]
        self.proj_head = None
        self.cnn_proj_head = None
        self._weights = None
        self._grads = None
        self.__dict__.pop("_fhb_params", None)

    def init_from_teacher_conv(self, teacher_model):
        self.feature_extractor.load_state_dict(teacher_model.model.feature_extractor.state_dict())
        try:
            self.post_extract_proj.load_state_dict(teacher_model.model.post_extract_proj.state_dict())
        except Exception:
            pass

    def init_from_teacher_enc(self, teacher_model, n_layers):
        assert n_layers <= self.cfg.encoder_layers
        self.encoder.pos_conv.load_state_dict(teacher_model.model.encoder.pos_conv.state_dict())
        for i in range(n_layers):
            self.encoder.layers[i].load_state_dict(teacher_model.model.encoder.layers[i].state_dict())

    # ---- engine plumbing -----------------------------------------------------------------
    def engine_state(self, train: bool):
        """(parameters, bf16 shadows, flat gradient buffer).  The gradient buffer (and with it the optimizer's moments,
        which are laid out like it) only depends on the parameter tensors: an eval-mode forward between two training
        steps (validation) re-uses the training shadows - they are a superset - and never resets anything."""
        P = _named_param_dict(self)
        first = next(iter(P.values()))
        _require_cuda(first, "CustomStudentModel")
        W = self._weights
        stale = W is None or W.device != first.device or any(
            W.params[k] is not v and W.params[k].data_ptr() != v.data_ptr() for k, v in P.items())
        if stale:
            self._weights = E.WeightSet(P, self._geom, train)
            self._grads = None
        elif train and not W.train:
            self._weights = E.WeightSet(P, self._geom, True)  # adds the dgrad-layout shadows; gradients untouched
        if train and self._grads is None:
            self._grads = E.GradStore(P, self._geom)
        return P, self._weights, self._grads

    def forward(self, source, padding_mask=None, layer=None, lengths: Optional[List[int]] = None):
        """Reference modules/model.py:420-552.  `lengths` (not in the reference signature, optional): the per-sample
        un-padded lengths the caller already knows, i.e. (~padding_mask).sum(-1); when given, the [B, L] mask is
        neither needed nor scanned (UpstreamExpert builds its mask from these very numbers)."""
        dev = self.post_extract_proj.weight.device
        _require_cuda(self.post_extract_proj.weight, "CustomStudentModel")
        source = source.to(dev, non_blocking=True).float().contiguous()
        if lengths is None:
            lengths = _lengths_from_mask(padding_mask)
        else:
            lengths = [int(n) for n in lengths]
            if len(lengths) != source.shape[0] or max(lengths) > source.shape[1] or min(lengths) < 0:
                raise ValueError("lengths must hold one value in [0, L] per sample")
            if all(n == source.shape[1] for n in lengths):
                lengths = None  # no sample is padded: the reference runs mask-free (modules/model.py:449,471-472)
        valid = None if lengths is None else conv_out_lengths(lengths, self._conv_layers)
        heads = "all" if self.proj_head is not None else ("last" if self.final_proj is not None else "none")
        n_run = None
        if layer is not None:
            # modules/module.py:300-340: `layer` indexes encoder.layers (entry 0 is the time-reduction conv when enabled);
            # the loop breaks after that entry, so entries 0..layer run
            n_run = max(0, min(self._geom.n_layers, int(layer) + (0 if self.enable_tr_layer else 1)))
            if n_run < self._geom.n_layers and heads == "all" and self.layerwise_proj:
                # the reference indexes layer_results[i] for every head (modules/model.py:493-499)
                raise IndexError("list index out of range")
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad and n_run is not None and n_run < self._geom.n_layers:
            raise NotImplementedError("the `layer` early exit has no backward on the B200 path (inference-time option)")
        if needs_grad:
            if not self.layerwise_proj and heads != "all":
                raise NotImplementedError("fine-tuning the SplitLinear recipe without its head is not implemented")
            from .autograd import student_apply
            c, preds, layers_out, feats_out, maps = student_apply(self, source, valid)
        else:
            P, W, _ = self.engine_state(False)
            c = E.student_forward(P, W, self._geom, source, valid, train=False, heads=heads, want_lr=True,
                                  drop=self.drop_cfg(), n_run=n_run, keep_qkv_last=self._return_attn)
            preds, layers_out = c.preds, c.layers
            feats_out = c.cnn_out if c.cnn_out is not None else c.feats
            maps = None
            if self._return_attn and len(c.layer_ctx) == self._geom.n_layers:
                maps = E.attn_maps(c.layer_ctx[-1].qkv, c.valid_s, c.B, c.Ts, self._geom.H, self._geom.d)
                maps = (maps["attn"], maps["vrel"])
        B, T, Ts, Em = c.B, c.T, c.Ts, self._geom.E
        mask = _frame_mask(valid, T, dev)
        layer_results = [(lo.view(B, Ts, Em).transpose(0, 1), None,
                          None if lr is None else lr.view(B, Ts, Em).transpose(0, 1))
                         for lo, lr in zip(layers_out, c.lrs)]
        if maps is not None:
            # utils/utils.py:190-258 bound over the layers (train.py:70-77): a layer's second output is (attn_logits, v_rel)
            # instead of None.  The reference returns the pair for every layer and reads the LAST one (train.py:329,357);
            # only that one is materialised here ([B*H, T, T] fp32 each - the other layers keep the flash kernels)
            x_, _, lr_ = layer_results[-1]
            layer_results[-1] = (x_, (maps[0][..., :Ts], maps[1][..., :Ts]), lr_)
        if not self.layerwise_proj:
            # reference modules/model.py:504-518: x stays the encoder output, projections is ONE [B, N, T, D] tensor
            projections = None if preds is None else preds.permute(1, 0, 2, 3)
            # with a TR layer the encoder output goes through the shared upsampler first (modules/model.py:504-505)
            x = c.x_up.view(B, 2 * Ts, Em) if getattr(c, "x_up", None) is not None else c.x_last.view(B, Ts, Em)
        elif heads == "all":
            projections = [preds[i] for i in range(preds.shape[0])]
            x = projections[-1]
        elif heads == "last":
            projections, x = None, preds[0]
        else:
            projections, x = None, c.x_last.view(B, Ts, Em)
        feats = feats_out.view(B, T, -1)
        if c.cnn_out is not None:
            pass  # cnn_proj_head(features) is a new tensor (modules/model.py:486-487): the in-place zeroing never reaches it
        elif mask is not None and not (self.training and self._drop_p["p_input"] > 0.0):
            # `features` aliases the encoder's input in the reference (modules/model.py:483,489): when dropout_input is an
            # identity, the encoder's in-place index_put(x, padding_mask, 0) (modules/module.py:273-274) zeroes the padded
            # frames of the returned tensor too; with dropout live it is a copy and they stay as computed
            feats = feats.masked_fill(mask.unsqueeze(-1), 0.0)
        return {
            "x": x,
            "padding_mask": mask,
            "features": feats,
            "layer_results": layer_results,
            "tr_layer_results": [] if c.tr is None else [c.tr.view(B, Ts, Em).transpose(0, 1)],
            "projections": projections,
            "_valid": valid,  # (not in the reference) per-sample valid frame counts behind `padding_mask`, host ints
        }

    def extract_features(self, source, padding_mask, layer=None):
        return self.forward(source, padding_mask, layer=layer)


# --------------------------------------------------------------------------- teacher
class TeacherModel(_ParamCacheMixin, nn.Module):
    """HuBERT-Base / wav2vec 2.0-Base `features_only` trunk with fairseq's parameter names
    (feature_extractor.*, layer_norm.*, post_extract_proj.*, encoder.*); SURVEY App. B.2, B.4."""

    def __init__(self, kind="hubert", conv_feature_layers="[(512,10,5)] + [(512,3,2)] * 4 + [(512,2,2)] * 2",
                 encoder_embed_dim=768, encoder_ffn_embed_dim=3072, encoder_attention_heads=12, encoder_layers=12,
                 conv_pos=128, conv_pos_groups=16):
        super().__init__()
        assert kind in ("hubert", "wav2vec2")
        self.kind = kind
        layers = parse_layer_spec(conv_feature_layers)
        self._conv_layers = layers
        self.feature_extractor = ConvFeatureExtractionModel(layers)
        self.layer_norm = nn.LayerNorm(layers[-1][0])
        self.post_extract_proj = nn.Linear(layers[-1][0], encoder_embed_dim)
        self.encoder = TransformerEncoder(encoder_embed_dim, encoder_ffn_embed_dim, encoder_attention_heads,
                                          encoder_layers, conv_pos, conv_pos_groups, tr_layer=False)
        self._geom = E.Geometry(layers, encoder_embed_dim, encoder_ffn_embed_dim, encoder_attention_heads,
                                conv_pos_groups, conv_pos, encoder_layers, 0, False)
        self._return_attn = False
        self._weights = None

    def engine_state(self):
        P = _named_param_dict(self)
        first = next(iter(P.values()))
        _require_cuda(first, "TeacherModel")
        if self._weights is None or any(self._weights.params[k].data_ptr() != v.data_ptr() for k, v in P.items()):
            self._weights = E.WeightSet(P, self._geom, False)
        return P, self._weights

    def frame_valid(self, padding_mask, Lmax: int, T: int) -> Optional[List[int]]:
        if padding_mask is None:
            return None
        if self.kind == "hubert":  # applied whenever a mask is passed, even an all-False one
            if padding_mask.is_cuda:
                out = torch.empty(padding_mask.shape[0], device=padding_mask.device, dtype=torch.int32)
                K.mask_lengths(padding_mask.contiguous().view(torch.uint8), out)
                lengths = out.tolist()
            else:
                lengths = (~padding_mask).sum(-1).tolist()
            return hubert_mask_lengths(lengths, Lmax, T)
        lengths = _lengths_from_mask(padding_mask)  # wav2vec2: rule M1, only if mask.any()
        return None if lengths is None else conv_out_lengths(lengths, self._conv_layers)

    @torch.no_grad()
    def extract_features(self, source, padding_mask=None, mask=None, out_buf=None, want_lr=False):
        dev = self.post_extract_proj.weight.device
        source = source.to(dev, non_blocking=True).float().contiguous()
        P, W = self.engine_state()
        T = E.conv_frames(source.shape[1], self._conv_layers)[-1]
        valid = self.frame_valid(padding_mask, source.shape[1], T)
        extras = {} if self._return_attn else None
        res = E.teacher_forward(P, W, self._geom, source, valid, out_buf=out_buf, want_lr=want_lr, extras=extras)
        self._last_maps = None
        if extras is not None:
            g = self._geom
            m = E.attn_maps(extras["qkv"], extras["valid_t"], extras["B"], extras["T"], g.H, g.d)
            self._last_maps = (m["attn"][..., :extras["T"]], m["vrel"][..., :extras["T"]])
        return (*res, valid) if want_lr else (*res, None, valid)


class TeacherWrapper(nn.Module):
    """Same contract as reference utils/utils.py:51-99: extract_features(source, padding_mask) ->
    {'layer_results': [(x [T,B,C], (None, layer_result [T,B,C]))] * n_layers, 'x': [B,T,C],
    'features': [post_extract_proj out]} - each entry is what the forward hook on an encoder layer captures
    (utils/utils.py:65-78: the layer returns (x, (attn, layer_result)), attn None with need_weights=False)."""

    def __init__(self, model: TeacherModel):
        super().__init__()
        self.model = model

    def extract_features(self, source, padding_mask=None, out_buf=None):
        layers, feats, lrs, valid = self.model.extract_features(source, padding_mask, out_buf=out_buf, want_lr=True)
        B, T, C = layers.shape[1:]
        maps = getattr(self.model, "_last_maps", None)
        n = layers.shape[0]
        res = {
            # with the attention-map recipe bound (train.py:64-69) the hook's `attn` slot of the LAST layer holds
            # (attn_logits, v_rel) [B*H, T, T] (utils/utils.py:229-258); see CustomStudentModel.forward
            "layer_results": [(layers[i].transpose(0, 1),
                               (maps if (maps is not None and i == n - 1) else None, lrs[i].view(B, T, C).transpose(0, 1)))
                              for i in range(n)],
            "x": layers[-1],
            "features": [feats],
        }
        res["_stacked"] = layers  # [n_layers, B, T, C] bf16, the fused loss kernel's target operand
        res["_valid"] = valid
        return res


def freeze_model(model: nn.Module):
    for p in model.parameters():
        p.requires_grad = False
