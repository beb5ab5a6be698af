"""CustomStudentModelConfig: same field names / defaults as the reference dataclass
(reference modules/model.py:21-251) so `CustomStudentModelConfig(**yaml['distiller'])`
(reference train.py:40-41, fithubert/expert.py:34) works unchanged."""
from __future__ import annotations

import ast
from dataclasses import dataclass
from typing import List, Optional, Tuple


def parse_layer_spec(spec) -> List[Tuple[int, int, int]]:
    """Evaluate strings like '[(128,10,5)] + [(256,3,2)] * 4' (the reference eval()s them,
    modules/model.py:267,384) with list '+' / '*' only - no eval."""
    if not isinstance(spec, str):
        return [tuple(int(v) for v in t) for t in spec]

    def walk(node):
        if isinstance(node, ast.Expression):
            return walk(node.body)
        if isinstance(node, ast.BinOp) and isinstance(node.op, (ast.Add, ast.Mult)):
            l, r = walk(node.left), walk(node.right)
            return l + r if isinstance(node.op, ast.Add) else l * r
        return ast.literal_eval(node)

    return [tuple(int(v) for v in t) for t in walk(ast.parse(spec.strip(), mode="eval"))]


def parse_int_list(spec) -> List[int]:
    return [int(v) for v in (ast.literal_eval(spec) if isinstance(spec, str) else spec)]


@dataclass
class CustomStudentModelConfig:
    _name: Optional[str] = None  # inherited from FairseqDataclass in the reference
    extractor_mode: str = "default"
    encoder_layers: int = 12
    encoder_embed_dim: int = 768
    encoder_ffn_embed_dim: int = 3072
    encoder_attention_heads: int = 12
    activation_fn: str = "gelu"
    layer_type: str = "transformer"
    n_mels: int = 0
    enable_log_mel: bool = False
    mel_spec_head_conv_layers: str = ""
    dropout: float = 0.1
    attention_dropout: float = 0.1
    activation_dropout: float = 0.0
    encoder_layerdrop: float = 0.0
    dropout_input: float = 0.0
    final_dim: int = 0
    layer_norm_first: bool = False
    conv_feature_layers: str = "[(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512,2,2)] * 2"
    conv_bias: bool = False
    feature_grad_mult: float = 1.0
    conv_pos: int = 128
    conv_pos_groups: int = 16
    pos_conv_depth: int = 1
    max_positions: int = 100000
    checkpoint_activations: bool = False
    required_seq_len_multiple: int = 2
    crop_seq_to_multiple: int = 1
    depthwise_conv_kernel_size: int = 31
    attn_type: str = ""
    pos_enc_type: str = "abs"
    fp16: bool = False
    init_conv_layers: bool = False
    init_encoder_layers: int = 0
    pred_head_inter_dim: int = 0
    pred_head_final_dim: int = 768
    pred_layer_id: str = "[3, 7, 11]"
    layerwise_proj: bool = False
    enable_tr_layer: bool = True
    tr_reduce_factor: int = 2
    tr_layer_type: str = "fc1"
    tr_conv1d_kernel: int = 2
    tr_layer_index: int = 1
    _teacher_task_agnostic: bool = False
    _cnn_weight: float = 0.0

    def validate_hot_path(self) -> None:
        """The B200 path implements exactly the configuration family the shipped FitHuBERT /
        FitW2V2 recipes use (SURVEY 2.1 scope column); anything else raises like the reference's
        own asserts / NotImplementedErrors do for unsupported options."""
        assert self.extractor_mode in {"default", "layer_norm"}
        if self.n_mels > 0:
            raise NotImplementedError("mel front-end (n_mels > 0) is outside the B200 hot path")
        assert self.enable_log_mel is False
        if self.extractor_mode != "default" or self.conv_bias:
            raise NotImplementedError("only extractor_mode='default', conv_bias=False is implemented")
        if self.layer_type != "transformer" or self.layer_norm_first:
            raise NotImplementedError("only post-LN 'transformer' layers are implemented")
        if self.activation_fn != "gelu":
            raise NotImplementedError("only activation_fn='gelu' is implemented")
        if self.pos_conv_depth != 1:
            raise NotImplementedError("pos_conv_depth > 1 is not implemented")
        if self.layerwise_proj:
            # FitHuBERT recipe (data/conf/fithubert.yaml): 12 LayerWiseProjHeads behind a conv1d TR layer at index 0;
            # without a TR layer LayerWiseProjHead is its Linear alone (modules/module.py:633-646), and with equal widths
            # it would be the identity (no parameters at all)
            if not self.enable_tr_layer and self.pred_head_final_dim == self.encoder_embed_dim:
                raise NotImplementedError("layer-wise heads without a TR layer need pred_head_final_dim != encoder_embed_dim")
        else:
            # DistilHuBERT-style recipe (data/conf/ex.yaml): Linear -> GELU -> SplitLinear on the last layer, no TR layer
            if len(parse_int_list(self.pred_layer_id)) < 2:
                raise NotImplementedError("the SplitLinear head needs at least two pred_layer_id entries")
        if self.enable_tr_layer:
            if self.tr_layer_type != "conv1d":
                raise NotImplementedError(
                    "Wrong type of time reduction layer."
                    "Time reduction layers must be one of ['fc1', 'fc2', 'conv1d']."
                    if self.tr_layer_type not in ("fc1", "fc2") else
                    "fc1/fc2 time-reduction layers are broken in the reference (SURVEY 2.1) and not implemented")
            if self.tr_layer_index != 0 or self.tr_reduce_factor != 2:
                raise NotImplementedError("time-reduction layer must be conv1d, index 0, factor 2")
        if self.required_seq_len_multiple != 1 or self.crop_seq_to_multiple != 1:
            raise NotImplementedError("required_seq_len_multiple / crop_seq_to_multiple must be 1")
        if self.encoder_layerdrop != 0.0:
            raise NotImplementedError("encoder_layerdrop must be 0")
        if not (0.0 < self.feature_grad_mult <= 1.0):
            raise NotImplementedError("feature_grad_mult must be in (0, 1] (0 freezes the extractor: not implemented)")
        layers = parse_layer_spec(self.conv_feature_layers)
        assert all(len(cl) == 3 for cl in layers), "invalid conv definition"
        if layers[0][1:] != (10, 5):
            raise NotImplementedError("first conv layer must be (C, 10, 5)")
        for (_, k, s) in layers[1:]:
            if (k, s) not in ((1, 1), (2, 2), (3, 2)):
                raise NotImplementedError(f"conv layer (k={k}, s={s}) is not implemented")
        if layers[-1][0] == self.encoder_embed_dim:
            # modules/model.py:298-302 then builds NO post_extract_proj (state-dict keys and arithmetic differ)
            raise NotImplementedError("conv output width == encoder_embed_dim (the reference drops post_extract_proj) "
                                      "is not implemented")
        assert self.encoder_embed_dim % self.encoder_attention_heads == 0
        assert self.encoder_embed_dim % self.conv_pos_groups == 0
