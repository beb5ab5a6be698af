"""fairseq.modules subset used by reference modules/module.py:10-17,22 and modules/model.py:8."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class GradMultiply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return x.new(x)

    @staticmethod
    def backward(ctx, grad):
        return grad * ctx.scale, None


def LayerNorm(normalized_shape, eps=1e-5, elementwise_affine=True, export=False):
    return torch.nn.LayerNorm(normalized_shape, eps, elementwise_affine)


class Fp32GroupNorm(nn.GroupNorm):
    def forward(self, input):
        output = F.group_norm(
            input.float(),
            self.num_groups,
            self.weight.float() if self.weight is not None else None,
            self.bias.float() if self.bias is not None else None,
            self.eps,
        )
        return output.type_as(input)


class Fp32LayerNorm(nn.LayerNorm):
    def forward(self, input):
        output = F.layer_norm(
            input.float(),
            self.normalized_shape,
            self.weight.float() if self.weight is not None else None,
            self.bias.float() if self.bias is not None else None,
            self.eps,
        )
        return output.type_as(input)


class SamePad(nn.Module):
    def __init__(self, kernel_size, causal=False):
        super().__init__()
        self.remove = 1 if kernel_size % 2 == 0 else 0

    def forward(self, x):
        if self.remove > 0:
            x = x[:, :, : -self.remove]
        return x


class TransposeLast(nn.Module):
    def forward(self, x):
        return x.transpose(-2, -1)


class RelPositionalEncoding(nn.Module):  # import-only on the hot path
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError


class MultiheadAttention(nn.Module):
    """Self-attention, manual (non-fused) path of fairseq's MultiheadAttention."""

    def __init__(self, embed_dim, num_heads, dropout=0.0, self_attention=False, **kw):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.dropout_p = dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim
        self.scaling = self.head_dim ** -0.5
        self.k_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.v_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.q_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        # fairseq's FairseqDropout: the attention-map recipe calls it directly (reference utils/utils.py:224)
        self.dropout_module = nn.Dropout(dropout)
        self.skip_embed_dim_check = False
        self.reset_parameters()

    def _set_skip_embed_dim_check(self):
        self.skip_embed_dim_check = True

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.k_proj.weight, gain=1 / math.sqrt(2))
        nn.init.xavier_uniform_(self.v_proj.weight, gain=1 / math.sqrt(2))
        nn.init.xavier_uniform_(self.q_proj.weight, gain=1 / math.sqrt(2))
        nn.init.xavier_uniform_(self.out_proj.weight)
        nn.init.constant_(self.out_proj.bias, 0.0)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=False,
                attn_mask=None, before_softmax=False, **kw):
        tgt_len, bsz, embed_dim = query.size()
        q = self.q_proj(query) * self.scaling
        k = self.k_proj(query)
        v = self.v_proj(query)
        H, d = self.num_heads, self.head_dim
        q = q.contiguous().view(tgt_len, bsz * H, d).transpose(0, 1)
        k = k.contiguous().view(-1, bsz * H, d).transpose(0, 1)
        v = v.contiguous().view(-1, bsz * H, d).transpose(0, 1)
        w = torch.bmm(q, k.transpose(1, 2))
        if key_padding_mask is not None:
            assert key_padding_mask.shape == (bsz, tgt_len)
            w = w.view(bsz, H, tgt_len, tgt_len)
            w = w.masked_fill(key_padding_mask.unsqueeze(1).unsqueeze(2).to(torch.bool), float("-inf"))
            w = w.view(bsz * H, tgt_len, tgt_len)
        if before_softmax:
            # fairseq multihead_attention.py (commit 1b61bbad): `if before_softmax: return attn_weights, v` right after
            # the key-padding masked_fill - the un-normalised logits [B*H, T, T] and the value heads [B*H, T, d]
            return w, v
        w_float = F.softmax(w, dim=-1, dtype=torch.float32)
        w = w_float.type_as(w)
        p = F.dropout(w, p=self.dropout_p, training=self.training)
        attn = torch.bmm(p, v)
        attn = attn.transpose(0, 1).contiguous().view(tgt_len, bsz, embed_dim)
        attn = self.out_proj(attn)
        return attn, None
