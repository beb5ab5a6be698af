"""Execution engine: orchestrates the hand-written sm_100a kernels (kernels.py -> libfhb_sm100a.so)
into the teacher forward, the student forward and the student backward.

Layout decisions:
  * every 16-bit tensor is fp16 (gradients carry a power-of-two loss scale, include/fhb.h), channel-last [B, T, C]; every contraction is a tcgen05 GEMM over a (possibly
    overlapping-row) strided view - no im2col, no permute copies;
  * fp32 master parameters live in the nn.Module; fp16 GEMM-layout shadows are re-derived by ONE
    multi-tensor kernel whenever the parameters change;
  * parameter gradients accumulate in ONE flat fp32 buffer, in the layout the wgrad GEMM produces
    (the fused AdamW kernel reads them through a 3-D stride; the buffer is what NCCL all-reduces).
There is no autograd inside: backward is written out explicitly (SURVEY App. F backward inventory).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch

from . import kernels as K
from . import lib as L

f16 = torch.float16    # every 16-bit tensor: weights, activations, saved gelu', projections, targets and gradients
bf16 = f16             # (old name, kept as an alias in the backward code: gradients are fp16 WITH a loss scale)


def loss_scale_for(n_elems: int) -> float:
    """Power-of-two loss scale for fp16 gradients (the reference trains under fp16 AMP with a GradScaler,
    data/conf/fithubert.yaml: use_fp16).  d(loss)/d(pred) = 2 w (pred - tgt) / n_elems per element; 2^(ceil(log2 n) + 2)
    puts it at 8 ... 16 x w x (pred - tgt): three decades above fp16's normal minimum, three below its maximum.  The
    scale is applied by the loss kernel and removed by the AdamW kernel (or on export to .grad); conversions saturate."""
    return float(2.0 ** min(24, max(0, math.ceil(math.log2(max(1, n_elems))) + 2)))
f32 = torch.float32


def conv_frames(n: int, layers) -> List[int]:
    out = []
    for (_, k, s) in layers:
        n = (n - k) // s + 1
        out.append(n)
    return out


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _fmix32(x: int) -> int:
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & 0xFFFFFFFF
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & 0xFFFFFFFF
    x ^= x >> 16
    return x


class DropCfg:
    """Dropout probabilities of one training forward/backward (reference modules/model.py:489,
    modules/module.py:294,566,573,578 + fairseq MultiheadAttention) and the seed its counter-based masks derive
    from.  Each dropout site gets its own 32-bit seed; forward and backward kernels regenerate identical masks."""
    SITE_INPUT, SITE_PROLOGUE, SITE_LAYER0 = 0, 1, 16
    ATTN, DROP1, ACT, DROP3 = 0, 1, 2, 3

    def __init__(self, seed: int, p_input=0.0, p_drop=0.0, p_attn=0.0, p_act=0.0):
        self.seed = _fmix32(seed)
        self.p_input, self.p_drop, self.p_attn, self.p_act = float(p_input), float(p_drop), float(p_attn), float(p_act)

    def any(self) -> bool:
        return max(self.p_input, self.p_drop, self.p_attn, self.p_act) > 0.0

    def site(self, site_id: int, p: float):
        """(seed, p) for one dropout site, or None when that site is inactive."""
        if p <= 0.0:
            return None
        return (_fmix32(self.seed + 0x632BE5AB * (site_id + 1)), p)

    def layer(self, l: int, which: int):
        p = (self.p_attn, self.p_drop, self.p_act, self.p_drop)[which]
        return self.site(self.SITE_LAYER0 + 4 * l + which, p)


class Geometry:
    def __init__(self, conv_layers, E, F, H, G, kpos, n_layers, d_out=0, student=True, tr=None, n_split=0, inter=0,
                 grad_mult=1.0):
        self.conv_layers = [tuple(c) for c in conv_layers]
        self.E, self.F, self.H, self.G, self.kpos, self.n_layers = E, F, H, G, kpos, n_layers
        self.d = E // H
        self.cg = E // G
        self.cp = round_up(self.cg, 16)
        # pos-conv time blocking: one GEMM row = pdelta consecutive frames (N = pdelta * cp fills the 128 x N UMMA)
        self.pdelta = 4 if (kpos + 4) * self.cp % 64 == 0 else 1
        self.kpx = kpos + (self.pdelta if self.pdelta > 1 else 0)  # taps of the blocked weight operand
        self.d_out = d_out
        self.student = student
        self.tr = student if tr is None else bool(tr)  # time-reduction conv at encoder.layers[0]
        self.n_split = n_split                          # > 0: DistilHuBERT head (Linear -> GELU -> SplitLinear), N tasks
        self.inter = inter or E
        self.grad_mult = float(grad_mult)               # feature_grad_mult (GradMultiply on the conv features)
        self.c_feat = self.conv_layers[-1][0]


# =============================================================================================
# fp16 / fp32 shadows of the master parameters
# =============================================================================================
class WeightSet:
    """GEMM-layout shadows for one model on one device."""

    def __init__(self, params: Dict[str, torch.Tensor], g: Geometry, train: bool):
        self.params, self.g, self.train = params, g, train
        self.device = next(iter(params.values())).device
        self._sig = None
        self._table = None
        self.views: Dict[str, torch.Tensor] = {}
        self._plan()

    # ---- plan: (shadow name, dtype, dst dims3, [(param, src_off, sstride3, dst_off_elems)])
    def _plan(self):
        g = self.g
        spec = []

        def add(name, dtype, dims, parts):
            spec.append((name, dtype, tuple(dims), parts))

        cin = self.g.conv_layers[0][0]
        for i, (c, k, s) in enumerate(g.conv_layers):
            if i == 0:
                continue
            pn = f"feature_extractor.conv_layers.{i}.0.weight"
            add(f"conv{i}.w", f16, (c, k, cin), [(pn, 0, (cin * k, 1, k), 0)])
            if self.train and (k, s) == (3, 2):
                add(f"conv{i}.wd_even", f16, (cin, 2, c), [(pn, 2, (k, -2, cin * k), 0)])
                add(f"conv{i}.wd_odd", f16, (cin, 1, c), [(pn, 1, (k, 0, cin * k), 0)])
            cin = c
        E, F = g.E, g.F

        def lin(name, pn, n, kd):
            add(name, f16, (n, 1, kd), [(pn, 0, (kd, 0, 1), 0)])

        lin("pp.w", "post_extract_proj.weight", E, g.c_feat)
        if g.student and "cnn_proj_head.1.weight" in self.params:  # CNN-feature head (modules/model.py:304-310)
            lin("cnn.w", "cnn_proj_head.1.weight", self.params["cnn_proj_head.1.weight"].shape[0], E)
        off = 0
        if g.tr:
            add("tr.w", f16, (E, 2, E), [("encoder.layers.0.weight", 0, (2 * E, 1, 2), 0)])
            off = 1
        for l in range(g.n_layers):
            p = f"encoder.layers.{l + off}."
            add(f"l{l}.wqkv", f16, (3 * E, 1, E),
                [(p + f"self_attn.{nm}_proj.weight", 0, (E, 0, 1), j * E * E) for j, nm in enumerate("qkv")])
            add(f"l{l}.bqkv", f32, (3 * E, 1, 1),
                [(p + f"self_attn.{nm}_proj.bias", 0, (1, 0, 0), j * E) for j, nm in enumerate("qkv")])
            lin(f"l{l}.wo", p + "self_attn.out_proj.weight", E, E)
            lin(f"l{l}.w1", p + "fc1.weight", F, E)
            lin(f"l{l}.w2", p + "fc2.weight", E, F)
        if g.student and g.n_split and g.tr and "upsampler.weight" in self.params:
            # shared upsampler in front of the DistilHuBERT head (modules/model.py:341-348,402-404,504-505)
            add("up.wup", f16, (2, E, E), [("upsampler.weight", 0, (1, 2, 2 * E), 0)])
            add("up.bup", f32, (2, E, 1), [("upsampler.bias", 0, (0, 1, 0), 0)])
        if g.student and g.n_split and "proj_head.2.weight" in self.params:
            # DistilHuBERT head (modules/module.py:585-619): Linear(E, N * inter) and the SplitLinear weight
            # [N][inter][D] transposed per task to the K-major [D][inter] the forward GEMM consumes
            lin("sp.w1", "proj_head.0.weight", g.n_split * g.inter, E)
            add("sp.w2", f16, (g.n_split, g.d_out, g.inter), [("proj_head.2.weight", 0, (g.inter * g.d_out, 1, g.d_out), 0)])
        elif g.student:
            for i in range(g.n_layers):
                p = f"proj_head.{i}."
                if p + "lin_proj.weight" not in self.params:
                    if i == g.n_layers - 1 and "final_proj.lin_proj.weight" in self.params:
                        p = "final_proj."
                    else:
                        continue
                if g.tr:  # (no TR layer: LayerWiseProjHead is the Linear alone, modules/module.py:633-646)
                    add(f"h{i}.wup", f16, (2, E, E), [(p + "upsampler.weight", 0, (1, 2, 2 * E), 0)])
                    add(f"h{i}.bup", f32, (2, E, 1), [(p + "upsampler.bias", 0, (0, 1, 0), 0)])
                    add(f"h{i}.bup16", f16, (E, 1, 1), [(p + "upsampler.bias", 0, (1, 0, 0), 0)])  # A operand of the bias fold
                lin(f"h{i}.wlin", p + "lin_proj.weight", g.d_out, E)
                add(f"h{i}.blin", f32, (g.d_out, 1, 1), [(p + "lin_proj.bias", 0, (1, 0, 0), 0)])
        self.spec = spec
        nb = sum(math.prod(d) for (_, t, d, _) in spec if t == f16)
        nf = sum(math.prod(d) for (_, t, d, _) in spec if t == f32)
        pc = g.G * g.pdelta * g.cp * g.kpx * g.cp
        self.buf16 = torch.empty(round_up(nb, 64) + 64 * len(spec) + 2 * round_up(pc, 64) + 128, device=self.device, dtype=f16)
        self.buf32 = torch.empty(nf + 8 * len(spec) + 2 * g.kpos + 64, device=self.device, dtype=f32)
        o16 = o32 = 0
        for (name, t, dims, _) in spec:
            n = math.prod(dims)
            if t == f16:
                self.views[name] = self.buf16[o16:o16 + n]
                o16 += round_up(n, 64)  # keep every shadow 128-byte aligned (TMA base alignment)
            else:
                self.views[name] = self.buf32[o32:o32 + n]
                o32 += round_up(n, 8)
        self.views["pc.w"] = self.buf16[o16:o16 + pc]
        o16 += round_up(pc, 64)
        self.views["pc.wt"] = self.buf16[o16:o16 + pc]
        self.views["pc.w"].zero_()   # the blocked layout keeps structural zeros the prep kernel never touches
        self.views["pc.wt"].zero_()
        self.views["pc.ws"] = self.buf32[o32:o32 + 2 * g.kpos]   # [sum of squares | 1 / norm] per tap
        self.views["pc.inv"] = self.views["pc.ws"][g.kpos:]

    def __getitem__(self, name):
        return self.views[name]

    def head_stride(self, name: str) -> Optional[int]:
        """Element stride between consecutive projection heads' shadow `name` (wup / bup / wlin / blin) if every
        head 0..n-1 has its own parameters at a uniform stride (then the 12 heads run as ONE batched GEMM)."""
        n = self.g.n_layers
        if any(f"h{i}.{name}" not in self.views for i in range(n)) or "proj_head.0.lin_proj.bias" not in self.params \
                or f"proj_head.{n - 1}.lin_proj.bias" not in self.params:
            return None
        es = self.views[f"h0.{name}"].element_size()
        ptrs = [self.views[f"h{i}.{name}"].data_ptr() for i in range(n)]
        if n == 1:
            return self.views[f"h0.{name}"].numel()
        d = ptrs[1] - ptrs[0]
        if d <= 0 or any(ptrs[i + 1] - ptrs[i] != d for i in range(n - 1)):
            return None
        return d // es

    def signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.params.values())

    def mark_stale(self):
        self._sig = None

    def ensure_fresh(self):
        sig = self.signature()
        if sig == self._sig:
            return
        ptrs = tuple(s[0] for s in sig)
        if self._table is None or ptrs != self._ptrs:
            entries = []
            self._max_n = 1
            for (name, t, dims, parts) in self.spec:
                per = math.prod(dims) // len(parts)
                pd = (dims[0] // len(parts), dims[1], dims[2])
                for (pn, soff, sstr, doff) in parts:
                    src = self.params[pn]
                    assert src.dtype == f32 and src.is_contiguous(), pn
                    e = L.PrepTensor()
                    e.src = src.data_ptr() + 4 * soff
                    e.dst = self.views[name].data_ptr() + doff * self.views[name].element_size()
                    e.dim = (L.C.c_int64 * 3)(*pd)
                    e.sstride = (L.C.c_int64 * 3)(*sstr)
                    e.dst_is_f32 = int(t == f32)
                    e.accumulate = 0
                    entries.append(e)
                    self._max_n = max(self._max_n, per)
            self._table = L.table_to_device(entries, self.device)
            self._n_entries = len(entries)
            self._ptrs = ptrs
        K.prep_multi(self._table, self._n_entries, self._max_n)
        g = self.g
        v, gn = self.params["encoder.pos_conv.0.weight_v"], self.params["encoder.pos_conv.0.weight_g"]
        K.posconv_wn_prep(v, gn, self.views["pc.w"], self.views["pc.wt"] if self.train else None, self.views["pc.ws"],
                          g.E, g.G, g.kpos, g.cp, g.pdelta)
        self._sig = sig


# =============================================================================================
# flat gradient buffer (student)
# =============================================================================================
class GradStore:
    """One flat fp32 buffer holding every parameter gradient in wgrad-GEMM layout.
    entry: name -> (offset, numel, param dims3, gstride3)."""

    def __init__(self, params: Dict[str, torch.Tensor], g: Geometry):
        self.params, self.g = params, g
        self.device = next(iter(params.values())).device
        E = g.E
        order: List[Tuple[str, Tuple[int, int, int], Tuple[int, int, int]]] = []

        def plain(pn):
            n = params[pn].numel()
            order.append((pn, (1, 1, n), (0, 0, 1)))

        cin = g.conv_layers[0][0]
        for pn in ("feature_extractor.conv_layers.0.0.weight", "feature_extractor.conv_layers.0.2.weight",
                   "feature_extractor.conv_layers.0.2.bias"):
            plain(pn)
        for i, (c, k, s) in enumerate(g.conv_layers):
            if i == 0:
                continue
            # grad stored as dW2[co][j][ci]; param element (co, ci, j)
            order.append((f"feature_extractor.conv_layers.{i}.0.weight", (c, cin, k), (k * cin, 1, cin)))
            cin = c
        for pn in ("layer_norm.weight", "layer_norm.bias", "post_extract_proj.weight", "post_extract_proj.bias",
                   "encoder.pos_conv.0.bias", "encoder.pos_conv.0.weight_g", "encoder.pos_conv.0.weight_v",
                   "encoder.layer_norm.weight", "encoder.layer_norm.bias"):
            plain(pn)
        if "cnn_proj_head.1.weight" in params:
            plain("cnn_proj_head.1.weight")
            plain("cnn_proj_head.1.bias")
        off = 0
        if g.tr:
            order.append(("encoder.layers.0.weight", (E, E, 2), (2 * E, 1, E)))
            plain("encoder.layers.0.bias")
            off = 1
        for l in range(off, g.n_layers + off):
            p = f"encoder.layers.{l}."
            for nm in "qkv":
                plain(p + f"self_attn.{nm}_proj.weight")
            for nm in "qkv":
                plain(p + f"self_attn.{nm}_proj.bias")
            # out_proj / fc1 / fc2 weights adjacent: when F == E (FitHuBERT) their three wgrads run as ONE batched GEMM
            for pn in ("self_attn.out_proj.weight", "fc1.weight", "fc2.weight", "self_attn.out_proj.bias",
                       "self_attn_layer_norm.weight", "self_attn_layer_norm.bias", "fc1.bias", "fc2.bias",
                       "final_layer_norm.weight", "final_layer_norm.bias"):
                plain(p + pn)
        if g.n_split and "proj_head.2.weight" in params:
            if g.tr and "upsampler.weight" in params:
                # grad stored as dWup[(j, co)][ci]; param element (ci, co, j)
                order.append(("upsampler.weight", (E, E, 2), (1, E, E * E)))
                plain("upsampler.bias")
            for pn in ("proj_head.0.weight", "proj_head.0.bias", "proj_head.2.weight", "proj_head.2.bias"):
                plain(pn)
        for i in range(g.n_layers if not g.n_split else 0):
            p = f"proj_head.{i}."
            if p + "lin_proj.weight" not in params:
                if i == g.n_layers - 1 and "final_proj.lin_proj.weight" in params:
                    p = "final_proj."  # after _disable_projection_heads: the last head under its new name
                else:
                    continue
            if g.tr:
                # grad stored as dWup[(j, co)][ci]; param element (ci, co, j)
                order.append((p + "upsampler.weight", (E, E, 2), (1, E, E * E)))
                plain(p + "upsampler.bias")
            plain(p + "lin_proj.weight")
            plain(p + "lin_proj.bias")
        self.entries = {}
        off = 0
        for (pn, dims, gs) in order:
            n = params[pn].numel()
            assert math.prod(dims) == n, pn
            self.entries[pn] = (off, n, dims, gs)
            off += round_up(n, 4)  # 16-byte alignment for vector / TMA-free fp32 stores
        self.numel = off
        self.loss_scale = 1.0  # scale of the gradients currently in the buffer (set by the fused step)
        self.flat = torch.zeros(off, device=self.device, dtype=f32)
        self.no_grad = [pn for pn in params if pn not in self.entries]  # the dead `upsampler.*`

    def view(self, pn) -> torch.Tensor:
        off, n, _, _ = self.entries[pn]
        return self.flat[off:off + n]

    def span(self, first, last) -> torch.Tensor:
        a = self.entries[first][0]
        b = self.entries[last][0] + self.entries[last][1]
        return self.flat[a:b]

    def zero_(self):
        K.zero_(self.flat)
        self.loss_scale = 1.0

    def head_stride(self) -> Optional[int]:
        """Float stride between consecutive projection heads' gradient blocks, if uniform."""
        n = self.g.n_layers
        names = ("upsampler.weight", "upsampler.bias", "lin_proj.weight", "lin_proj.bias")
        if any(f"proj_head.{i}.{nm}" not in self.entries for i in range(n) for nm in names):
            return None
        if n == 1:
            return self.numel
        d = self.entries["proj_head.1.lin_proj.bias"][0] - self.entries["proj_head.0.lin_proj.bias"][0]
        for nm in names:
            offs = [self.entries[f"proj_head.{i}.{nm}"][0] for i in range(n)]
            if any(offs[i + 1] - offs[i] != d for i in range(n - 1)):
                return None
        return d

    def from_(self, pn) -> torch.Tensor:
        """Flat view from the start of `pn` to the end of the buffer (base pointer of a strided batch)."""
        return self.flat[self.entries[pn][0]:]

    def export(self, accumulate=False, scale: Optional[float] = None) -> Dict[str, torch.Tensor]:
        """Gradients in PARAMETER layout (what autograd would have produced); `scale`: the loss scale the backward ran
        with (default: the one the fused step recorded in .loss_scale); the exported gradients are divided by it."""
        scale = self.loss_scale if scale is None else scale
        assert not (accumulate and scale != 1.0)
        out = {}
        entries = []
        mx = 1
        for pn, (off, n, dims, gs) in self.entries.items():
            dst = torch.empty_like(self.params[pn]) if not accumulate else self.params[pn].grad
            out[pn] = dst
            e = L.PrepTensor()
            e.src = self.flat.data_ptr() + 4 * off
            e.dst = dst.data_ptr()
            e.dim = (L.C.c_int64 * 3)(*dims)
            e.sstride = (L.C.c_int64 * 3)(*gs)
            e.dst_is_f32, e.accumulate = 1, int(accumulate)
            entries.append(e)
            mx = max(mx, n)
        table = L.table_to_device(entries, self.device)
        K.prep_multi(table, len(entries), mx)
        self._keepalive = table
        if scale != 1.0:
            torch._foreach_mul_(list(out.values()), 1.0 / scale)
        return out


# =============================================================================================
# forward building blocks
# =============================================================================================
def _valid_tensor(valid: Optional[List[int]], device) -> Optional[torch.Tensor]:
    if valid is None:
        return None
    return torch.tensor(valid, dtype=torch.int32).to(device, non_blocking=True)


def conv_stack_fwd(P, W: WeightSet, g: Geometry, wave: torch.Tensor, save: bool, wave_chunks=None):
    """Conv feature extractor (reference modules/module.py:94-102).  Returns ctx with the last
    activation [B, T, C] (contiguous view) and, when `save`, everything backward needs.
    When saving for backward, the output of every k=3,s=2 layer is stored as [B, T_i + 2, C_i] with one
    halo row on each side: its gradient buffer shares that layout and the zero halo rows turn the k=3,s=2
    dgrad into two plain overlapped-view GEMMs (even / odd input frames)."""
    B, Ld = wave.shape
    dev = wave.device
    frames = conv_frames(Ld, g.conv_layers)
    C0 = g.conv_layers[0][0]
    T0 = frames[0]
    halos = [1 if (save and i > 0 and (k, s) == (3, 2)) else 0 for i, (_, k, s) in enumerate(g.conv_layers)]
    c = SimpleNamespace(B=B, L=Ld, frames=frames, halos=halos, wave=wave)
    c.stat = torch.empty(B, 65, device=dev, dtype=torch.float64)
    c.mean0 = torch.empty(B, C0, device=dev, dtype=f32)
    c.rstd0 = torch.empty(B, C0, device=dev, dtype=f32)
    y = torch.empty(B, T0, C0, device=dev, dtype=f16)
    gp0 = torch.empty(B, T0, C0, device=dev, dtype=f16) if save else None
    c.y = [y]
    c.u = [gp0]  # per layer: gelu'(pre-activation), saved by the forward epilogue (the backward multiplier)
    for i, (co, k, s) in enumerate(g.conv_layers):
        if i == 0:
            continue
        rows = frames[i] + 2 * halos[i]
        c.y.append(torch.empty(B, rows, co, device=dev, dtype=f16))
        c.u.append(torch.empty(B, rows, co, device=dev, dtype=f16) if save else None)
    # Nothing in the extractor couples samples (GroupNorm statistics are per sample and channel), so when the waveform is
    # still arriving from the host in batch slices (wave_chunks = [(first, last, event)], see h2d_chunked) the WHOLE
    # stack runs slice by slice, each as soon as its copy has landed: the rest of the transfer hides under it.
    for (b0, b1, ev) in (wave_chunks or [(0, B, None)]):
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        nb = b1 - b0
        K.conv0_fwd(wave[b0:b1], P["feature_extractor.conv_layers.0.0.weight"], P["feature_extractor.conv_layers.0.2.weight"],
                    P["feature_extractor.conv_layers.0.2.bias"], T0, c.stat[b0:b1], c.mean0[b0:b1], c.rstd0[b0:b1], y[b0:b1],
                    gp_out=None if gp0 is None else gp0[b0:b1])
        # (buffer, rows allocated per sample, first data row)
        x_buf, x_rows, x_row0, cin, T = y, T0, 0, C0, T0
        for i, (co, k, s) in enumerate(g.conv_layers):
            if i == 0:
                continue
            To = frames[i]
            halo = halos[i]
            rows = To + 2 * halo
            yb, ub = c.y[i], c.u[i]
            a3 = L.tensor3(x_buf, data_ptr=x_buf.data_ptr() + 2 * (b0 * x_rows + x_row0) * cin, dim=(k * cin, To, nb),
                           stride=(s * cin, x_rows * cin))
            b3 = L.tensor3(W[f"conv{i}.w"], data_ptr=W[f"conv{i}.w"].data_ptr(), dim=(k * cin, co, 1), stride=(k * cin, k * cin * co))
            K.gemm_raw(a3, b3, yb, To, co, k * cin, num_ob=nb, a_coord=(0, 1, 0, 0), d_ld=co, d_hi_stride=rows * co,
                       d_offset_elems=(b0 * rows + halo) * co,
                       flags=L.EPI_GELU | ((L.EPI_STORE_PREACT | L.EPI_AUX_DGELU) if save else 0), aux_out=ub)
            x_buf, x_rows, x_row0, cin, T = yb, rows, halo, co, To
    c.T = T
    assert halos[-1] == 0, "the last conv layer must not be k=3,s=2 (its output feeds a dense LayerNorm)"
    c.out = x_buf
    return c


def frontend_fwd(P, W: WeightSet, g: Geometry, wave, valid, save: bool,
                 drop: Optional[DropCfg] = None, wave_chunks=None, want_enc32: bool = False):
    """conv stack -> LayerNorm -> post_extract_proj -> (mask, pos-conv, +x, LayerNorm)
    = reference modules/model.py:428-489 + modules/module.py:273-281."""
    c = conv_stack_fwd(P, W, g, wave, save, wave_chunks)
    B, T, dev = c.B, c.T, wave.device
    # `valid` (per-sample valid frame counts, or None) may be a callable: the conv stack above does not need it, so a
    # caller that still has to derive the lengths from a host padding mask does that scan while the GPU already works
    if callable(valid):
        valid = valid()
    valid_t = _valid_tensor(valid, dev)
    c.valid, c.valid_t = valid, valid_t
    E, Cf = g.E, g.c_feat
    feat = c.out
    f_ln = torch.empty(B, T, Cf, device=dev, dtype=f16)
    c.mean_f = torch.empty(B * T, device=dev, dtype=f32) if save else None
    c.rstd_f = torch.empty(B * T, device=dev, dtype=f32) if save else None
    K.layernorm_fwd(feat, P["layer_norm.weight"], P["layer_norm.bias"], f_ln, c.mean_f, c.rstd_f)
    c.f_ln = f_ln
    c.cnn_out = c.cnn_g = None
    if g.student and "cnn.w" in W.views:
        # CNN-feature head (modules/model.py:304-310,486-487): features_to_distill = Linear(GELU(features)).  ONE GEMM gives
        # both gelu(features) (D) and the features themselves (the stored pre-activation); the backward recomputes gelu'
        feats = torch.empty(B * T, E, device=dev, dtype=f16)
        c.cnn_g = K.linear(f_ln.view(B * T, Cf), W["pp.w"].view(E, Cf), P["post_extract_proj.bias"], gelu=True, preact_out=feats)
        Dc = P["cnn_proj_head.1.weight"].shape[0]
        c.cnn_out = K.linear(c.cnn_g, W["cnn.w"].view(Dc, E), P["cnn_proj_head.1.bias"])  # [B*T, Dc]
    else:
        feats = K.linear(f_ln.view(B * T, Cf), W["pp.w"].view(E, Cf), P["post_extract_proj.bias"])
    c.feats = feats  # [B*T, E], padded frames NOT zeroed (reference `features_to_distill`)
    d_in = drop.site(DropCfg.SITE_INPUT, drop.p_input) if drop is not None else None
    if d_in is not None:  # dropout_input (modules/model.py:489); `features_to_distill` above stays un-dropped
        feats = K.dropout(feats, torch.empty_like(feats), *d_in)
    # positional conv as one batched GEMM per (sample, group)
    G, cp, kp, dl = g.G, g.cp, g.kpos, g.pdelta
    R = (T + dl - 1) // dl                   # GEMM rows per (sample, group): one row = dl consecutive frames
    Tp = dl * R + g.kpx                      # padded time extent every overlapped row window stays inside
    xg = torch.empty(B * G, Tp, cp, device=dev, dtype=f16)
    K.posconv_pack(feats, valid_t, xg, B, T, E, G, cp, kp // 2, Tp)
    conv = torch.empty(B * R, G * dl * cp, device=dev, dtype=f16)  # [b][r][g][dl][cp]
    a3 = L.tensor3(xg, data_ptr=xg.data_ptr(), dim=(g.kpx * cp, R, B * G), stride=(dl * cp, Tp * cp))
    b3 = L.tensor3(W["pc.w"], data_ptr=W["pc.w"].data_ptr(), dim=(g.kpx * cp, dl * cp, G), stride=(g.kpx * cp, dl * cp * g.kpx * cp))
    K.gemm_raw(a3, b3, conv, R, dl * cp, g.kpx * cp, num_ob=B * G, ob_mod=G, a_coord=(0, G, 1, 0), b_coord=(0, 0, 1, 0),
               d_ld=G * dl * cp, d_hi_stride=R * G * dl * cp, d_lo_stride=dl * cp)
    c.xg, c.conv = (xg, conv) if save else (None, conv)
    enc = torch.empty(B * T, E, device=dev, dtype=f16)
    c.h = torch.empty(B * T, E, device=dev, dtype=f16) if save else None
    c.mean_e = torch.empty(B * T, device=dev, dtype=f32) if save else None
    c.rstd_e = torch.empty(B * T, device=dev, dtype=f32) if save else None
    d_pro = drop.site(DropCfg.SITE_PROLOGUE, drop.p_drop) if drop is not None else None
    # fp32 copy of the encoder input: the residual operand of layer 0 when no GEMM sits in between (no TR layer) and no
    # dropout follows (the dropped tensor is fp16 anyway)
    c.enc_in32 = torch.empty(B * T, E, device=dev, dtype=f32) if (want_enc32 and d_pro is None) else None
    K.posconv_finish_fwd(feats, valid_t, conv, P["encoder.pos_conv.0.bias"], P["encoder.layer_norm.weight"],
                         P["encoder.layer_norm.bias"], c.h, enc, c.mean_e, c.rstd_e, B, T, E, G, cp, delta=dl,
                         y32=c.enc_in32)
    c.Tp, c.R = Tp, R
    if d_pro is not None:  # F.dropout after the encoder LayerNorm (modules/module.py:294)
        K.dropout(enc, enc, *d_pro)
    c.enc_in = enc
    return c


def layer_fwd(P, W: WeightSet, g: Geometry, prefix: str, l: int, x, valid_t, B, T, save: bool, want_lr: bool,
              out: Optional[torch.Tensor] = None, drop: Optional[DropCfg] = None, x32: Optional[torch.Tensor] = None,
              out32: Optional[torch.Tensor] = None, stream32: bool = True, keep_qkv: bool = False):
    """One post-LN transformer layer (reference modules/module.py:557-580) on x [B*T, E].
    The residual stream never passes through 16 bits: x32 is the fp32 copy of x (None for the first layer, whose input is
    the fp16 output of a GEMM anyway), the out_proj / fc2 epilogues add it in fp32 and write the sums y1 / y2 in fp32, the
    LayerNorms read those and emit the fp16 GEMM operand together with the next fp32 copy (out32)."""
    E, F, H, d = g.E, g.F, g.H, g.d
    dev = x.device
    s = SimpleNamespace(x=x)
    M = B * T
    if not stream32:
        # forward-only variant with the residual stream in fp16 (the frozen teacher, FHB_TEACHER_STREAM32=0): half the
        # bytes through the out_proj / fc2 epilogues and the LayerNorms; fp16's 11-bit significand keeps the 12-layer
        # deviation at a few 1e-3 (profiles/r02*_teacher_stream16.txt)
        assert not save
        qkv = K.linear(x, W[f"l{l}.wqkv"].view(3 * E, E), W[f"l{l}.bqkv"])
        attn = torch.empty(M, E, device=dev, dtype=f16)
        K.attn_fwd(qkv, valid_t, attn, None, B, T, H, d, d ** -0.5)
        y1 = K.linear(attn, W[f"l{l}.wo"].view(E, E), P[prefix + "self_attn.out_proj.bias"], residual=x)
        x1 = torch.empty(M, E, device=dev, dtype=f16)
        K.layernorm_fwd(y1, P[prefix + "self_attn_layer_norm.weight"], P[prefix + "self_attn_layer_norm.bias"], x1)
        h = K.linear(x1, W[f"l{l}.w1"].view(F, E), P[prefix + "fc1.bias"], gelu=True)
        lr = torch.empty(M, E, device=dev, dtype=f16) if want_lr else None
        y2 = K.linear(h, W[f"l{l}.w2"].view(E, F), P[prefix + "fc2.bias"], residual=x1, preact_out=lr)
        x2 = out if out is not None else torch.empty(M, E, device=dev, dtype=f16)
        K.layernorm_fwd(y2, P[prefix + "final_layer_norm.weight"], P[prefix + "final_layer_norm.bias"], x2)
        s.out, s.lr, s.out32 = x2, lr, None
        if keep_qkv:
            s.qkv = qkv
        return s
    qkv = K.linear(x, W[f"l{l}.wqkv"].view(3 * E, E), W[f"l{l}.bqkv"])
    if keep_qkv:
        s.qkv = qkv
    # training with F == E: the inputs of out_proj / fc1 / fc2 (attn, x1, h) share one [3, M, E] buffer so that their three
    # weight-gradient GEMMs run as one batched launch in the backward (wgrad_batch_enabled)
    xs = torch.empty(3, M, E, device=dev, dtype=f16) if (save and F == E and wgrad_batch_enabled()) else None
    attn = xs[0] if xs is not None else torch.empty(M, E, device=dev, dtype=f16)
    lse = torch.empty(B, H, T, device=dev, dtype=f32) if save else None
    dl = (lambda which: drop.layer(l, which)) if drop is not None else (lambda which: None)
    K.attn_fwd(qkv, valid_t, attn, lse, B, T, H, d, d ** -0.5, drop=dl(DropCfg.ATTN))
    y1 = K.linear(attn, W[f"l{l}.wo"].view(E, E), P[prefix + "self_attn.out_proj.bias"],
                  residual=x32 if x32 is not None else x, drop=dl(DropCfg.DROP1), out_dtype=f32)
    x1 = xs[1] if xs is not None else torch.empty(M, E, device=dev, dtype=f16)
    x1_32 = torch.empty(M, E, device=dev, dtype=f32)
    s.mean1 = torch.empty(M, device=dev, dtype=f32) if save else None
    s.rstd1 = torch.empty(M, device=dev, dtype=f32) if save else None
    K.layernorm_fwd32(y1, P[prefix + "self_attn_layer_norm.weight"], P[prefix + "self_attn_layer_norm.bias"], x1, x1_32,
                      s.mean1, s.rstd1)
    u = torch.empty(M, F, device=dev, dtype=f16) if save else None
    h = K.linear(x1, W[f"l{l}.w1"].view(F, E), P[prefix + "fc1.bias"], gelu=True, dgelu_out=u, drop=dl(DropCfg.ACT),
                 out=None if xs is None else xs[2])
    s.xs = xs
    d3 = dl(DropCfg.DROP3)
    y2 = K.linear(h, W[f"l{l}.w2"].view(E, F), P[prefix + "fc2.bias"], residual=x1_32, drop=d3, out_dtype=f32)
    lr = torch.empty(M, E, device=dev, dtype=f16) if want_lr else None
    if want_lr and d3 is not None:
        # `layer_result` is the fc2 output BEFORE dropout3 (modules/module.py:577-578); y2 - x1 would be the dropped one
        K.linear(h, W[f"l{l}.w2"].view(E, F), P[prefix + "fc2.bias"], out=lr)
    x2 = out if out is not None else torch.empty(M, E, device=dev, dtype=f16)
    s.mean2 = torch.empty(M, device=dev, dtype=f32) if save else None
    s.rstd2 = torch.empty(M, device=dev, dtype=f32) if save else None
    recover = want_lr and d3 is None
    K.layernorm_fwd32(y2, P[prefix + "final_layer_norm.weight"], P[prefix + "final_layer_norm.bias"], x2, out32, s.mean2,
                      s.rstd2, sub32=x1_32 if recover else None, diff_out=lr if recover else None)
    if save:
        s.qkv, s.attn, s.lse, s.y1, s.x1, s.u, s.h, s.y2 = qkv, attn, lse, y1, x1, u, h, y2
    s.out, s.lr, s.out32 = x2, lr, out32
    return s


# =============================================================================================
# attention-map / value-relation distillation
# =============================================================================================
def attn_maps(qkv: torch.Tensor, valid_dev, B: int, T: int, H: int, d: int, want=("attn", "vrel"), out=None):
    """What `rtrn_attn_forward` (reference utils/utils.py:190-232) returns next to a layer's output: the un-normalised
    attention logits bmm(q * scaling, k^T) with -inf at padded keys, and v_rel = bmm(v * scaling, v^T); each fp32
    [B*H, T, pitch] (pitch = T rounded up to 8; the columns beyond T are padding).  qkv: that layer's fused projection
    output [B*T, 3*H*d] fp16.  out: optional dict of buffers to write into."""
    E = H * d
    q, k, v = qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:]
    res = {}
    if "attn" in want:
        res["attn"] = K.attn_scores(q, k, valid_dev, B, T, H, d, d ** -0.5, out=None if out is None else out.get("attn"))
    if "vrel" in want:
        res["vrel"] = K.attn_scores(v, v, None, B, T, H, d, d ** -0.5, out=None if out is None else out.get("vrel"))
    return res


def _pow2_near(x: float) -> float:
    return 2.0 ** max(-40, min(40, round(math.log2(max(x, 1e-30)))))


def attn_transfer_losses(c, g: Geometry, t_extras: dict, gt: Geometry, *, loss_type: str, w_attn: float, w_vrel: float,
                         grad_scale: float, valid_s: Optional[List[int]], valid_t: Optional[List[int]]):
    """Attention distribution transfer + value relation transfer losses of the LAST layer (reference train.py:327-368) and
    their gradients wrt the student's maps, for the fused training step.  Returns fp32 [2] (attn_loss, v_rel_loss,
    un-weighted); stores c.attn_grad = [(kind, G fp16 [B*H, T, pitch], alpha)]: student_backward adds
    alpha * (G k, G^T q) to (dq, dk) for kind 'qk' and alpha * (G + G^T) v to dv for kind 'vv'.
    grad_scale: the loss scale of the fp16 gradients (x 1/accumulation ...).  G additionally carries a power of two that
    centres it in fp16's range; alpha takes it out again in fp32."""
    B, T, H, d = c.B, c.Ts, g.H, g.d
    if g.tr:
        # train.py:70-77 touches `.self_attn` of every encoder.layers entry; entry 0 is the time-reduction nn.Conv1d
        raise AttributeError("'Conv1d' object has no attribute 'self_attn'")
    if t_extras["T"] != T or gt.H != H or t_extras["B"] != B:
        raise RuntimeError(f"The size of tensor a ({B * H}, {T}, {T}) must match the size of tensor b "
                           f"({t_extras['B'] * gt.H}, {t_extras['T']}, {t_extras['T']}): the attention maps of student and "
                           "teacher need the same head count and frame rate")
    if loss_type not in ("mse", "kldiv"):
        raise NotImplementedError("attn_loss_type must be one of 'mse', 'kldiv'.")
    dev = c.lay.device
    s_qkv, t_qkv = c.layer_ctx[-1].qkv, t_extras["qkv"]
    rows = B * H * T
    losses = torch.zeros(2, device=dev, dtype=f32)
    c.attn_grad = []
    bufs = None
    if w_attn > 0:
        S_s = attn_maps(s_qkv, c.valid_s, B, T, H, d, ("attn",))["attn"]
        S_t = attn_maps(t_qkv, t_extras["valid_t"], B, T, gt.H, gt.d, ("attn",))["attn"]
        bufs = (S_s, S_t)
        if loss_type == "mse":
            vs = valid_s if valid_s is not None else [T] * B
            vt = valid_t if valid_t is not None else [T] * B
            loss_mult = 1.0 / (H * T * sum(min(a, b, T) for a, b in zip(vs, vt)))
            boost = _pow2_near(1.0 / (grad_scale * w_attn * loss_mult))
        else:
            loss_mult = 1.0 / rows
            boost = _pow2_near(0.25 * T / (grad_scale * w_attn * loss_mult))
        G = torch.empty(S_s.shape, device=dev, dtype=f16)
        K.attn_map_loss(S_s, S_t, c.valid_s, t_extras["valid_t"], G, losses[0:1], B, T, H, 0 if loss_type == "mse" else 1,
                        loss_mult, grad_scale * w_attn * loss_mult * boost)
        c.attn_grad.append(("qk", G, d ** -0.5 / boost))
    if w_vrel > 0:
        out = None if bufs is None else {"vrel": bufs[0]}
        R_s = attn_maps(s_qkv, None, B, T, H, d, ("vrel",), out)["vrel"]
        out = None if bufs is None else {"vrel": bufs[1]}
        R_t = attn_maps(t_qkv, None, B, T, gt.H, gt.d, ("vrel",), out)["vrel"]
        loss_mult = 1.0 / rows
        boost = _pow2_near(0.25 * T / (grad_scale * w_vrel * loss_mult))
        G = torch.empty(R_s.shape, device=dev, dtype=f16)
        K.attn_map_loss(R_s, R_t, None, None, G, losses[1:2], B, T, H, 1, loss_mult, grad_scale * w_vrel * loss_mult * boost)
        c.attn_grad.append(("vv", G, d ** -0.5 / boost))
    return losses


def attn_transfer_backward(c, g: Geometry, qkv: torch.Tensor, dqkv: torch.Tensor):
    """Adds the gradients of the last layer's attention-map / value-relation terms to dqkv [B*Ts, 3E] (after the flash
    attention backward has written it): logits = scaling * q k^T -> dq += scaling * dS k, dk += scaling * dS^T q;
    v_rel = scaling * v v^T -> dv += scaling * (dR + dR^T) v."""
    B, T, H, d = c.B, c.Ts, g.H, g.d
    E = H * d
    q, k, v = qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:]
    dq, dk, dv = dqkv[:, :E], dqkv[:, E:2 * E], dqkv[:, 2 * E:]
    for kind, G, alpha in c.attn_grad:
        if kind == "qk":
            K.attn_scores_bwd(G, k, dq, B, T, H, d, alpha, trans=0)
            K.attn_scores_bwd(G, q, dk, B, T, H, d, alpha, trans=1)
        else:
            K.attn_scores_bwd(G, v, dv, B, T, H, d, alpha, trans=0)
            K.attn_scores_bwd(G, v, dv, B, T, H, d, alpha, trans=1)


# =============================================================================================
# teacher
# =============================================================================================
def teacher_forward(P, W: WeightSet, g: Geometry, wave, valid: Optional[List[int]], out_buf=None, slots=None,
                    wave_chunks=None, want_lr: bool = False, extras: Optional[dict] = None):
    """Frozen teacher forward (reference utils/utils.py:80-99 around fairseq HubertModel /
    Wav2Vec2Model.extract_features).  Returns (layers [n_layers, B, T, E] fp16, features [B, T, E]).
    slots: optional list mapping teacher layer -> row of out_buf (None = not a distillation target: that layer's
    output goes to a scratch buffer), so the loss kernel finds the pred_layer_id targets stacked without a gather.
    extras: a dict to fill with what the attention-map recipe needs from the LAST layer (utils/utils.py:190-232 bound over the
    teacher's layers, train.py:64-69): 'qkv' [B*T, 3E] fp16, 'valid_t' (device int32 or None), 'B', 'T'."""
    W.ensure_fresh()
    import os
    # the frozen teacher carries its residual stream in fp16 by default (A/B on B200, profiles/r02i_teacher_stream16_ab.txt:
    # 24.0 / 24.6 ms -> 23.7 / 23.6 ms per step; teacher layers deviate 3.1e-3 instead of 1.4e-3 from the fp32 oracle,
    # student gradients unchanged at 2.0e-3); FHB_TEACHER_STREAM32=1 gives it the student's fp32 stream
    stream32 = os.environ.get("FHB_TEACHER_STREAM32", "0") == "1"
    c = frontend_fwd(P, W, g, wave, valid, save=False, wave_chunks=wave_chunks, want_enc32=stream32)
    valid_t = c.valid_t
    B, T, E = c.B, c.T, g.E
    if out_buf is None:
        n_out = g.n_layers if slots is None else 1 + max(s for s in slots if s is not None)
        out_buf = torch.empty(n_out, B, T, E, device=wave.device, dtype=f16)
    x, x32 = c.enc_in, c.enc_in32
    scratch = None
    lrs = []
    ping = [torch.empty(B * T, E, device=wave.device, dtype=f32) for _ in range(2)]  # fp32 copies of the layer outputs
    last = g.n_layers if (slots is None or extras is not None) else 1 + max(l for l, s in enumerate(slots) if s is not None)
    for l in range(last):  # layers above the highest target are never needed
        slot = l if slots is None else slots[l]
        if slot is None:
            scratch = [torch.empty(B * T, E, device=wave.device, dtype=f16) for _ in range(2)] if scratch is None else scratch
            dst = scratch[l & 1]
        else:
            dst = out_buf[slot].view(B * T, E)
        s = layer_fwd(P, W, g, f"encoder.layers.{l}.", l, x, valid_t, B, T, save=False, want_lr=want_lr, out=dst, x32=x32,
                      out32=ping[l & 1] if (l + 1 < last and stream32) else None, stream32=stream32,
                      keep_qkv=extras is not None and l == g.n_layers - 1)
        x, x32 = s.out, s.out32
        lrs.append(s.lr)
        if extras is not None and l == g.n_layers - 1:
            extras.update(qkv=s.qkv, valid_t=valid_t, valid_host=c.valid, B=B, T=T)
    if want_lr:  # the hook output of every layer is (x, (attn, layer_result)), utils/utils.py:65-78
        return out_buf, c.feats.view(B, T, E), lrs
    return out_buf, c.feats.view(B, T, E)


# =============================================================================================
# student
# =============================================================================================
def wgrad_batch_enabled() -> bool:
    """FHB_WGRAD_BATCH=0 keeps one weight-gradient launch per Linear (see layer_fwd / student_backward)."""
    import os
    return os.environ.get("FHB_WGRAD_BATCH", "1") == "1"


def head_compose_enabled() -> bool:
    """Run each LayerWiseProjHead as ONE GEMM against the folded weight (see _compose_heads); FHB_HEAD_COMPOSE=0 keeps
    the two-GEMM form (A/B on B200, profiles/r01z_streams_ab.log: 22.96 / 23.26 ms -> 22.57 / 22.46 ms per step)."""
    import os
    return os.environ.get("FHB_HEAD_COMPOSE", "1") == "1"


def _compose_heads(W: WeightSet, g: Geometry, hs: Dict[str, int], n: int):
    """LayerWiseProjHead (modules/module.py:649-661) is ConvTranspose1d(k=2,s=2) followed by Linear with nothing in
    between: for output phase p,  pred[2t+p] = x[t] (Wup_p Wlin^T) + (bup Wlin^T + blin).  Fold the two weights once per
    step (two tiny batched GEMMs + two 1-row GEMMs for the bias) and the 12 heads run as ONE [B*Ts, E] x [E, 2D] GEMM
    instead of [B*Ts, E] x [E, 2E] followed by [B*2Ts, E] x [E, D]: 38 % fewer head FLOPs forward and backward, the
    [n, B, 2Ts, E] intermediate is never written, and one 16-bit rounding of the activations disappears.
    Returns Wc [n, 2, D, E] fp16 (K-major [2D, E] per head) and bc [n, 2D] fp32."""
    E, D = g.E, g.d_out
    dev = W["h0.wlin"].device
    Wc = torch.empty(n, 2, D, E, device=dev, dtype=f16)
    bc = torch.empty(n, 2 * D, device=dev, dtype=f32)
    wlin3 = L.tensor3(W["h0.wlin"], data_ptr=W["h0.wlin"].data_ptr(), dim=(E, D, n), stride=(E, hs["wlin"]))
    bup3 = L.tensor3(W["h0.bup16"], data_ptr=W["h0.bup16"].data_ptr(), dim=(E, 1, n), stride=(E, hs["bup16"]))
    for p in range(2):
        # Wc[h, p] [D x E_in] = Wlin[h] [D x E_out] (K-major) x Wup[h, p] (rows = E_out, E_in contiguous: MN-major B)
        wup3 = L.tensor3(W["h0.wup"], data_ptr=W["h0.wup"].data_ptr() + 2 * p * E * E, dim=(E, E, n), stride=(E, hs["wup"]))
        K.gemm_raw(wlin3, wup3, Wc, D, E, E, b_major=1, num_ob=n, a_coord=(0, 1, 0, 0), b_coord=(0, 1, 0, 0), d_ld=E,
                   d_hi_stride=2 * D * E, d_offset_elems=p * D * E)
        # bc[h, p*D:(p+1)*D] = bup[h] Wlin[h]^T + blin[h]   (M = 1)
        K.gemm_raw(bup3, wlin3, bc, 1, D, E, num_ob=n, a_coord=(0, 1, 0, 0), b_coord=(0, 1, 0, 0), d_ld=2 * D,
                   d_hi_stride=2 * D, d_offset_elems=p * D, flags=L.EPI_BIAS, bias=W["h0.blin"],
                   bias_hi_stride=hs["blin"])
    return Wc, bc


def student_forward(P, W: WeightSet, g: Geometry, wave, valid: Optional[List[int]], *, train: bool,
                    heads: str = "all", want_lr: bool = False, pred_buf=None, drop: Optional[DropCfg] = None,
                    wave_chunks=None, n_run: Optional[int] = None, keep_qkv_last: bool = False):
    """Student forward (reference modules/model.py:420-552).  heads: 'all' (12 LayerWiseProjHeads),
    'last' (after _disable_projection_heads: final_proj on the last layer), 'none'.
    n_run: transformer layers to execute (the reference's `layer=` early exit, modules/module.py:335-340); the heads then
    see the last executed layer's output.
    Returns ctx: .layers [n_layers][B*Ts, E], .tr [B*Ts, E], .preds [n_layers or 1, B, T', D], .feats."""
    W.ensure_fresh()
    n_run = g.n_layers if n_run is None else max(0, min(int(n_run), g.n_layers))
    assert n_run == g.n_layers or (not train and heads != "all"), "early exit is an inference-time option without heads"
    dev = wave.device
    if drop is not None and not drop.any():
        drop = None
    c = frontend_fwd(P, W, g, wave, valid, save=train, drop=drop, wave_chunks=wave_chunks, want_enc32=not g.tr)
    valid, valid_t = c.valid, c.valid_t  # `valid` may have been a callable (resolved after the conv stack was queued)
    valid_s = _valid_tensor(None if valid is None else [v // 2 for v in valid], dev)
    B, T, E = c.B, c.T, g.E
    if g.tr:
        Ts = T // 2
        # time-reduction Conv1d(k=2, s=2) (modules/module.py:317-321): reshaped-view GEMM, drops an odd tail frame
        tr = torch.empty(B * Ts, E, device=dev, dtype=f16)
        a3 = L.tensor3(c.enc_in, data_ptr=c.enc_in.data_ptr(), dim=(2 * E, Ts, B), stride=(2 * E, T * E))
        b3 = L.tensor3(W["tr.w"], data_ptr=W["tr.w"].data_ptr(), dim=(2 * E, E, 1), stride=(2 * E, 2 * E * E))
        K.gemm_raw(a3, b3, tr, Ts, E, 2 * E, num_ob=B, a_coord=(0, 1, 0, 0), d_ld=E, d_hi_stride=Ts * E,
                   flags=L.EPI_BIAS, bias=P["encoder.layers.0.bias"])
        off = 1
    else:  # ex.yaml: enable_tr_layer False - the layers run at the full frame rate, mask rule M1 un-reduced
        Ts, tr, valid_s, off = T, None, valid_t, 0
    c.Ts, c.valid_t, c.valid_s, c.drop = Ts, valid_t, valid_s, drop
    c.tr = tr
    x, x32 = (tr, None) if tr is not None else (c.enc_in, c.enc_in32)
    c.layer_ctx = []
    lay = torch.empty(g.n_layers, B * Ts, E, device=dev, dtype=f16)  # stacked layer outputs (batched heads)
    ping = [torch.empty(B * Ts, E, device=dev, dtype=f32) for _ in range(2)]  # fp32 copies of the layer outputs
    c.x_last = x
    for l in range(n_run):
        s = layer_fwd(P, W, g, f"encoder.layers.{l + off}.", l, x, valid_s, B, Ts, save=train, want_lr=want_lr, out=lay[l],
                      drop=drop, x32=x32, out32=ping[l & 1] if l + 1 < n_run else None,
                      keep_qkv=keep_qkv_last and l == g.n_layers - 1)
        c.layer_ctx.append(s)
        x, x32 = s.out, s.out32
    c.x_last = x  # output of the last executed layer (the TR conv / prologue output when n_run == 0)
    c.lay = lay
    c.layers = [s.out for s in c.layer_ctx]
    c.lrs = [s.lr for s in c.layer_ctx]
    D = g.d_out
    n = g.n_layers
    if g.n_split:
        # DistilHuBERT head on the last layer (modules/model.py:504-518): Linear(E, N * inter) -> GELU -> SplitLinear
        # (modules/module.py:585-619) = one GEMM + one GEMM batched over the N tasks; preds [N, B, T, D]
        c.head_idx, c.heads_batched, c.preds = [], False, None
        c.x_up = None
        Tq = Ts
        if g.tr and "up.wup" in W.views:
            # shared upsampler (modules/model.py:402-404,504-505): ConvTranspose1d(k=2,s=2) on the encoder output =
            # GEMM to [B*Ts, 2E] = [B*2Ts, E]; `x` of the result dict is this upsampled tensor
            Tq = 2 * Ts
            c.x_up = K.linear(c.x_last, W["up.wup"].view(2 * E, E), W["up.bup"]).view(B * Tq, E)
        c.Tq = Tq
        if heads == "all":
            N, inter = g.n_split, g.inter
            xin = c.x_up if c.x_up is not None else c.x_last
            c.sp_u = torch.empty(B * Tq, N * inter, device=dev, dtype=f16) if train else None
            c.sp_h = K.linear(xin, W["sp.w1"].view(N * inter, E), P["proj_head.0.bias"], gelu=True, dgelu_out=c.sp_u)
            if pred_buf is None:
                pred_buf = torch.empty(N, B, Tq, D, device=dev, dtype=f16)
            a3 = L.tensor3(c.sp_h, data_ptr=c.sp_h.data_ptr(), dim=(inter, B * Tq, N), stride=(N * inter, inter))
            b3 = L.tensor3(W["sp.w2"], data_ptr=W["sp.w2"].data_ptr(), dim=(inter, D, N), stride=(inter, D * inter))
            K.gemm_raw(a3, b3, pred_buf, B * Tq, D, inter, num_ob=N, a_coord=(0, 1, 0, 0), b_coord=(0, 1, 0, 0), d_ld=D,
                       d_hi_stride=B * Tq * D, flags=L.EPI_BIAS, bias=P["proj_head.2.bias"], bias_hi_stride=D)
            c.preds = pred_buf
        return c
    # projection heads (modules/module.py:649-661): ConvTranspose1d(k=2,s=2) = GEMM to [B*Ts, 2E] = [B*2Ts, E], then the
    # Linear; without a TR layer the head is the Linear alone (:633-646) at the encoder's own frame rate
    Tq = 2 * Ts if g.tr else Ts
    idx = list(range(n)) if heads == "all" else ([n - 1] if heads == "last" else [])
    c.head_idx = idx
    hs = {k: W.head_stride(k) for k in ("wup", "bup", "wlin", "blin")} if (heads == "all" and g.tr) else {}
    c.heads_batched = bool(idx) and heads == "all" and g.tr and all(v is not None for v in hs.values())
    if idx:
        if pred_buf is None:
            pred_buf = torch.empty(len(idx), B, Tq, D, device=dev, dtype=f16)
        c.Wc = None
        if c.heads_batched and head_compose_enabled() and W.head_stride("bup16") is not None:
            hs["bup16"] = W.head_stride("bup16")
            c.Wc, bc = _compose_heads(W, g, hs, n)
            a3 = L.tensor3(lay, data_ptr=lay.data_ptr(), dim=(E, B * Ts, n), stride=(E, B * Ts * E))
            b3 = L.tensor3(c.Wc, data_ptr=c.Wc.data_ptr(), dim=(E, 2 * D, n), stride=(E, 2 * D * E))
            K.gemm_raw(a3, b3, pred_buf, B * Ts, 2 * D, E, num_ob=n, a_coord=(0, 1, 0, 0), b_coord=(0, 1, 0, 0),
                       d_ld=2 * D, d_hi_stride=B * Ts * 2 * D, flags=L.EPI_BIAS, bias=bc, bias_hi_stride=2 * D)
            c.z = None
            c.head_strides = hs
        elif c.heads_batched:
            # all n heads as TWO batched GEMMs (ob = head): 12x the tiles per launch, no per-head launch tails
            z = torch.empty(n, B * Tq, E, device=dev, dtype=f16)
            a3 = L.tensor3(lay, data_ptr=lay.data_ptr(), dim=(E, B * Ts, n), stride=(E, B * Ts * E))
            b3 = L.tensor3(W["h0.wup"], data_ptr=W["h0.wup"].data_ptr(), dim=(E, 2 * E, n), stride=(E, hs["wup"]))
            K.gemm_raw(a3, b3, z, B * Ts, 2 * E, E, num_ob=n, a_coord=(0, 1, 0, 0), b_coord=(0, 1, 0, 0), d_ld=2 * E,
                       d_hi_stride=B * Ts * 2 * E, flags=L.EPI_BIAS, bias=W["h0.bup"], bias_hi_stride=hs["bup"])
            a3 = L.tensor3(z, data_ptr=z.data_ptr(), dim=(E, B * Tq, n), stride=(E, B * Tq * E))
            b3 = L.tensor3(W["h0.wlin"], data_ptr=W["h0.wlin"].data_ptr(), dim=(E, D, n), stride=(E, hs["wlin"]))
            K.gemm_raw(a3, b3, pred_buf, B * Tq, D, E, num_ob=n, a_coord=(0, 1, 0, 0), b_coord=(0, 1, 0, 0), d_ld=D,
                       d_hi_stride=B * Tq * D, flags=L.EPI_BIAS, bias=W["h0.blin"], bias_hi_stride=hs["blin"])
            c.z = z if train else None
            c.head_strides = hs
        else:
            c.z = []
            for j, i in enumerate(idx):
                src = c.layers[i] if i < n_run else c.x_last  # `layer=` early exit: final_proj sees the last executed layer
                z = K.linear(src, W[f"h{i}.wup"].view(2 * E, E), W[f"h{i}.bup"]) if g.tr else src
                K.linear(z.view(B * Tq, E), W[f"h{i}.wlin"].view(D, E), P[_head_name(P, i) + "lin_proj.bias"],
                         out=pred_buf[j].view(B * Tq, D))
                c.z.append(z if train else None)
    c.preds = pred_buf if idx else None
    c.Tq = Tq
    return c


def _head_name(P, i):
    return f"proj_head.{i}." if f"proj_head.{i}.lin_proj.bias" in P else "final_proj."


_SIDE_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


def side_stream(dev) -> "torch.cuda.Stream":
    """One extra stream per device for work that is off the step's critical path (FHB_STREAMS)."""
    idx = torch.device(dev).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _SIDE_STREAMS:
        _SIDE_STREAMS[idx] = torch.cuda.Stream(device=idx)
    return _SIDE_STREAMS[idx]


_COPY_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


_H2D_BUFS: Dict[tuple, list] = {}


def h2d_chunked(x_host: torch.Tensor, dev, n_chunks: int = 0):
    """Host -> device copy of a [B, L] fp32 batch on a copy stream, into one of two persistent device buffers per shape
    (a prefetching loader's double buffer).  Returns (x_dev, chunks, done): consumers wait for a chunk's event before
    touching it (conv_stack_fwd does, slice by slice), and record `done` on their stream after the last use of x_dev so
    that the buffer is not overwritten two steps later before they are finished.
      * GPU still busy with the previous step (the caller does not synchronise every step): ONE copy, issued at once - it
        does not wait for the queued work, so it runs under the previous step's kernels and the conv stacks run un-sliced;
      * GPU idle (the caller reads the loss back every step): `n_chunks` batch slices, the conv stacks start on the first
        one while the others are in flight, so only the first slice's transfer is exposed.
    True DMA overlap needs pinned host memory; pageable memory works but is staged by the driver."""
    if n_chunks <= 0:
        import os
        n_chunks = int(os.environ.get("FHB_H2D_CHUNKS", "4"))
    idx = torch.device(dev).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _COPY_STREAMS:
        _COPY_STREAMS[idx] = torch.cuda.Stream(device=idx)
    cs = _COPY_STREAMS[idx]
    key = (idx, tuple(x_host.shape))
    if key not in _H2D_BUFS:
        if len(_H2D_BUFS) > 8:  # shapes change with every bucket of a real loader: keep the cache small
            _H2D_BUFS.clear()
        _H2D_BUFS[key] = [0, [torch.empty(x_host.shape, device=dev, dtype=f32) for _ in range(2)],
                          [torch.cuda.Event(), torch.cuda.Event()]]
    ent = _H2D_BUFS[key]
    ent[0] ^= 1
    x, done = ent[1][ent[0]], ent[2][ent[0]]
    main = torch.cuda.current_stream()
    busy = not main.query()
    cs.wait_event(done)  # the step that used this buffer two calls ago (a never-recorded event is complete)
    B = x_host.shape[0]
    step = B if busy else -(-B // max(1, min(n_chunks, B)))
    chunks = []
    with torch.cuda.stream(cs):
        for b0 in range(0, B, step):
            b1 = min(B, b0 + step)
            x[b0:b1].copy_(x_host[b0:b1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
            chunks.append((b0, b1, ev))
    return x, chunks, done


def stream_mode() -> int:
    """FHB_STREAMS bit mask (default 1).  1 = the frozen teacher's forward runs on the side stream beside the student's
    forward (they only meet in the loss); 2 = the transformer layers' weight-gradient GEMMs and bias column sums run on
    the side stream beside the dgrad chain; 4 = the conv stack's weight-gradient GEMMs too.  Interleaved A/B on B200
    (profiles/r01z_streams_compose_ab.log, r01z_streams2_ab.log): bit 1 saves 0.15 - 0.4 ms of a 23 ms step in every
    pair, bits 2 and 4 are within the noise (every GEMM is a persistent one-CTA-per-SM kernel, so two of them never
    share an SM and only launch tails overlap) and stay off."""
    import os
    return int(os.environ.get("FHB_STREAMS", "1"))


class _OffPath:
    """`with aside(t1, t2, ...):` runs the enclosed launches on the side stream after everything queued so far on the
    current stream.  The tensors named are main-stream temporaries the side-stream kernels read: they are kept alive
    until join() (after which the main stream has waited for the side stream), so the caching allocator never hands
    their blocks out early and - unlike record_stream - the allocation pattern is the same every step.
    Disabled: a no-op context, everything stays in order."""

    def __init__(self, dev, enabled: bool):
        self.side = side_stream(dev) if enabled else None
        self._keep: List[torch.Tensor] = []
        self._ctx = None

    def __call__(self, *tensors):
        if self.side is not None:
            self._keep.extend(t for t in tensors if t is not None)
        return self

    def __enter__(self):
        if self.side is None:
            return self
        self.side.wait_stream(torch.cuda.current_stream())
        self._ctx = torch.cuda.stream(self.side)
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.side is not None:
            self._ctx.__exit__(*exc)
        return False

    def join(self):
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
            self._keep.clear()


def _wgrad(dy3: L.Tensor3, x3: L.Tensor3, out: torch.Tensor, M: int, N: int, Kc: int, num_cb=1, a_cb=0, b_cb=0,
           a_c1_off=0, b_c1_off=0):
    """out[M][N] (fp32, accumulated) += sum over contraction of dy^T x; both MN-major."""
    K.gemm_raw(dy3, x3, out, M, N, Kc, a_major=1, b_major=1, num_cb=num_cb, a_coord=(0, 0, 0, a_cb),
               b_coord=(0, 0, 0, b_cb), d_ld=N, flags=L.EPI_ATOMIC_ADD, a_c1_off=a_c1_off, b_c1_off=b_c1_off)


def student_backward(P, W: WeightSet, g: Geometry, G_: GradStore, c, dpred: torch.Tensor,
                     dlayers: Optional[List[Optional[torch.Tensor]]] = None,
                     dpred_colsum: Optional[torch.Tensor] = None, on_layers_done=None, on_progress=None,
                     dfeatures: Optional[torch.Tensor] = None):
    """Backward of student_forward(train=True, heads='all').  dpred: [n_layers, B, T', D] bf16 gradient of
    the loss wrt every projection (zeros where unused).  Accumulates into the flat gradient buffer.
    dpred_colsum: fp32 [n_layers, D] column sums of dpred over (B, T') if the loss kernel already produced them
    (batched-heads path only); both head bias gradients are derived from them (fhb_head_bias_grads).
    on_progress(offset): called with an offset whenever flat[offset:] of the gradient buffer has become final - after the
    heads and then after every second transformer layer (the buffer is laid out front end, layers 0..n-1, heads, and the
    backward walks it from the end) - so that a data-parallel caller can start reducing that part under the rest of the
    pass; called with None after the layers in between and once the prologue backward is queued (progress ticks).
    dfeatures: fp16 gradient wrt the returned `features` (features_to_distill, train.py:241-246): [B*T, D_cnn] when the
    model has a cnn_proj_head, else [B*T, E]."""
    E, F, H, d, D = g.E, g.F, g.H, g.d, g.d_out
    B, T, Ts, Tq = c.B, c.T, c.Ts, c.Tq
    dev = c.lay.device
    gv = G_.view
    dx = None  # gradient wrt the current layer's output [B*Ts, E]
    n = g.n_layers
    mode = stream_mode()
    aside = _OffPath(dev, bool(mode & 2))       # transformer-layer wgrads / bias column sums
    aside_conv = _OffPath(dev, bool(mode & 4))  # conv-stack wgrads
    gs = G_.head_stride() if getattr(c, "heads_batched", False) else None
    dx_head = None
    if g.n_split:
        # ---- DistilHuBERT head backward: SplitLinear (batched over the N tasks), GELU, Linear.  dpred [N, B, T, D];
        #      the SplitLinear bias gradient (column sums of dpred) was accumulated by the loss kernel or is summed here
        N, inter, rows = g.n_split, g.inter, B * Tq
        if dpred_colsum is None:
            K.colsum_batched(dpred.view(N, rows, D), gv("proj_head.2.bias"), D)
        a3 = L.tensor3(c.sp_h, data_ptr=c.sp_h.data_ptr(), dim=(inter, rows, N), stride=(N * inter, inter))
        b3 = L.tensor3(dpred, data_ptr=dpred.data_ptr(), dim=(D, rows, N), stride=(D, rows * D))
        K.gemm_raw(a3, b3, gv("proj_head.2.weight"), inter, D, rows, a_major=1, b_major=1, num_ob=N, a_coord=(0, 1, 0, 0),
                   b_coord=(0, 1, 0, 0), d_ld=D, d_hi_stride=inter * D, flags=L.EPI_ATOMIC_ADD)
        dh = torch.empty(rows, N * inter, device=dev, dtype=bf16)  # d(pre-GELU): x gelu'(u) in the epilogue
        a3 = L.tensor3(dpred, data_ptr=dpred.data_ptr(), dim=(D, rows, N), stride=(D, rows * D))
        b3 = L.tensor3(W["sp.w2"], data_ptr=W["sp.w2"].data_ptr(), dim=(inter, D, N), stride=(inter, D * inter))
        K.gemm_raw(a3, b3, dh, rows, inter, D, b_major=1, num_ob=N, a_coord=(0, 1, 0, 0), b_coord=(0, 1, 0, 0),
                   d_ld=N * inter, d_hi_stride=inter, flags=L.EPI_MUL_AUX, aux_in=c.sp_u)
        K.colsum(dh, gv("proj_head.0.bias"))
        xin = c.x_up if getattr(c, "x_up", None) is not None else c.layers[-1]
        K.linear_wgrad(dh, xin, out=gv("proj_head.0.weight").view(N * inter, E), accumulate=True)
        dx = K.linear_dgrad(dh, W["sp.w1"].view(N * inter, E))
        if getattr(c, "x_up", None) is not None:
            # shared upsampler backward: dz [B*2Ts, E] == [B*Ts, 2E]
            K.colsum(dx, gv("upsampler.bias"))
            dz2 = dx.view(B * Ts, 2 * E)
            K.linear_wgrad(dz2, c.layers[-1], out=gv("upsampler.weight").view(2 * E, E), accumulate=True)
            dx = K.linear_dgrad(dz2, W["up.wup"].view(2 * E, E))
    if gs is not None:
        # ---- all n projection heads at once (ob = head): 2 wgrad + 2 dgrad batched GEMMs, 1-2 column sums
        hs = c.head_strides
        dp3 = dpred.view(n, B * Tq, D)
        if dpred_colsum is None:
            dpred_colsum = torch.zeros(n, D, device=dev, dtype=f32)
            K.colsum_batched(dp3, dpred_colsum, D)
        # lin_proj.bias += colsum(dpred); upsampler.bias += colsum(dpred) @ Wlin (= colsum of dz below, never re-read)
        K.head_bias_grads(dpred_colsum, W["h0.wlin"], hs["wlin"], G_.from_("proj_head.0.lin_proj.bias"),
                          G_.from_("proj_head.0.upsampler.bias"), gs, n, D, E)
    if gs is not None and getattr(c, "Wc", None) is not None:
        # ---- folded heads (see _compose_heads): dWc = dpred^T x, dx = dpred Wc, then the chain rule back to the two
        #      original weights (four small batched GEMMs) and the rank-1 term of the folded bias
        rows = B * Ts
        # dWc sums (loss-scaled) products over all B*Ts rows: x 2^-ceil(log2 rows) keeps it inside fp16's range (the
        # chain GEMMs below multiply it back while accumulating into the fp32 gradient buffer)
        down = 2.0 ** -math.ceil(math.log2(max(2, rows)))
        dWc = torch.empty(n, D, 2, E, device=dev, dtype=bf16)  # [h][d][p][i]
        a3 = L.tensor3(dpred, data_ptr=dpred.data_ptr(), dim=(2 * D, rows, n), stride=(2 * D, rows * 2 * D))
        x3 = L.tensor3(c.lay, data_ptr=c.lay.data_ptr(), dim=(E, rows, n), stride=(E, rows * E))
        K.gemm_raw(a3, x3, dWc, D, E, rows, a_major=1, b_major=1, num_ob=2 * n, ob_mod=2, a_coord=(D, 1, 0, 0),
                   b_coord=(0, 1, 0, 0), d_ld=2 * E, d_lo_stride=E, d_hi_stride=D * 2 * E, alpha=down)
        dx_head = torch.empty(n, rows, E, device=dev, dtype=bf16)
        b3 = L.tensor3(c.Wc, data_ptr=c.Wc.data_ptr(), dim=(E, 2 * D, n), stride=(E, 2 * D * E))
        K.gemm_raw(a3, b3, dx_head, rows, E, 2 * D, b_major=1, num_ob=n, a_coord=(0, 1, 0, 0), b_coord=(0, 1, 0, 0),
                   d_ld=E, d_hi_stride=rows * E)
        wlin3 = L.tensor3(W["h0.wlin"], data_ptr=W["h0.wlin"].data_ptr(), dim=(E, D, n), stride=(E, hs["wlin"]))
        for ph in range(2):
            dwc3 = L.tensor3(dWc, data_ptr=dWc.data_ptr() + 2 * ph * E, dim=(E, D, n), stride=(2 * E, D * 2 * E))
            wup3 = L.tensor3(W["h0.wup"], data_ptr=W["h0.wup"].data_ptr() + 2 * ph * E * E, dim=(E, E, n), stride=(E, hs["wup"]))
            # dWlin[h] [D x E_out] += dWc[h, :, ph, :] [D x E_in] x Wup[h, ph] ([E_out][E_in]: K-major B)
            K.gemm_raw(dwc3, wup3, G_.from_("proj_head.0.lin_proj.weight"), D, E, E, num_ob=n, a_coord=(0, 1, 0, 0),
                       b_coord=(0, 1, 0, 0), d_ld=E, d_hi_stride=gs, flags=L.EPI_ATOMIC_ADD, alpha=1.0 / down)
            # dWup[h, ph] [E_out x E_in] += Wlin[h]^T [E_out x D] x dWc[h, :, ph, :] [D x E_in]   (both MN-major)
            K.gemm_raw(wlin3, dwc3, G_.from_("proj_head.0.upsampler.weight"), E, E, D, a_major=1, b_major=1, num_ob=n,
                       a_coord=(0, 1, 0, 0), b_coord=(0, 1, 0, 0), d_ld=E, d_hi_stride=gs, d_offset_elems=ph * E * E,
                       flags=L.EPI_ATOMIC_ADD, alpha=1.0 / down)
        # z = x Wup + bup feeds lin_proj: dWlin += colsum(dpred) (x) bup   (n x D x E fp32 elements, one pass)
        glin = G_.from_("proj_head.0.lin_proj.weight").as_strided((n, D, E), (gs, E, 1))
        bup = W["h0.bup"].as_strided((n, 1, E), (hs["bup"], E, 1))
        glin.addcmul_(dpred_colsum.view(n, D, 1), bup)
    elif gs is not None:
        a3 = L.tensor3(dpred, data_ptr=dpred.data_ptr(), dim=(D, B * Tq, n), stride=(D, B * Tq * D))
        z3 = L.tensor3(c.z, data_ptr=c.z.data_ptr(), dim=(E, B * Tq, n), stride=(E, B * Tq * E))
        K.gemm_raw(a3, z3, G_.from_("proj_head.0.lin_proj.weight"), D, E, B * Tq, a_major=1, b_major=1, num_ob=n,
                   a_coord=(0, 1, 0, 0), b_coord=(0, 1, 0, 0), d_ld=E, d_hi_stride=gs, flags=L.EPI_ATOMIC_ADD)
        dz = torch.empty(n, B * Tq, E, device=dev, dtype=bf16)
        b3 = L.tensor3(W["h0.wlin"], data_ptr=W["h0.wlin"].data_ptr(), dim=(E, D, n), stride=(E, hs["wlin"]))
        K.gemm_raw(a3, b3, dz, B * Tq, E, D, b_major=1, num_ob=n, a_coord=(0, 1, 0, 0), b_coord=(0, 1, 0, 0), d_ld=E,
                   d_hi_stride=B * Tq * E)
        a3 = L.tensor3(dz, data_ptr=dz.data_ptr(), dim=(2 * E, B * Ts, n), stride=(2 * E, B * Ts * 2 * E))
        x3 = L.tensor3(c.lay, data_ptr=c.lay.data_ptr(), dim=(E, B * Ts, n), stride=(E, B * Ts * E))
        K.gemm_raw(a3, x3, G_.from_("proj_head.0.upsampler.weight"), 2 * E, E, B * Ts, a_major=1, b_major=1, num_ob=n,
                   a_coord=(0, 1, 0, 0), b_coord=(0, 1, 0, 0), d_ld=E, d_hi_stride=gs, flags=L.EPI_ATOMIC_ADD)
        dx_head = torch.empty(n, B * Ts, E, device=dev, dtype=bf16)
        b3 = L.tensor3(W["h0.wup"], data_ptr=W["h0.wup"].data_ptr(), dim=(E, 2 * E, n), stride=(E, hs["wup"]))
        K.gemm_raw(a3, b3, dx_head, B * Ts, E, 2 * E, b_major=1, num_ob=n, a_coord=(0, 1, 0, 0), b_coord=(0, 1, 0, 0),
                   d_ld=E, d_hi_stride=B * Ts * E)
    off = 1 if g.tr else 0
    # The gradient of the residual stream is carried in fp32 (dx32), like the stream itself in the forward: the dgrad
    # epilogues add an fp32 residual and write fp32, the LayerNorm backward reads that and emits the bf16 GEMM operand
    # next to the fp32 copy that continues down the residual.  dxb: bf16 contributions entering at this layer's output
    # (its projection head, autograd's dlayers), summed inside the LayerNorm backward.
    dx32 = None
    dxb = dx  # the split head's gradient wrt the last layer (bf16) or None
    dx = None
    for l in range(g.n_layers - 1, -1, -1):
        s = c.layer_ctx[l]
        p = f"encoder.layers.{l + off}."
        hp = _head_name(P, l)

        def add_b(t, cur):
            return t if cur is None else K.add_bf16(cur, t, torch.empty_like(t))

        if dx_head is not None:
            dxb = add_b(dx_head[l], dxb)
        elif l in c.head_idx and not g.n_split:
            j = c.head_idx.index(l)
            dp = dpred[j].view(B * Tq, D)
            z = c.z[j].view(B * Tq, E)
            K.colsum(dp, gv(hp + "lin_proj.bias"))
            K.linear_wgrad(dp, z, out=gv(hp + "lin_proj.weight").view(D, E), accumulate=True)
            dz = K.linear_dgrad(dp, W[f"h{l}.wlin"].view(D, E))  # [B*Tq, E] == [B*Ts, 2E]
            if g.tr:
                K.colsum(dz, gv(hp + "upsampler.bias"))
                dz2 = dz.view(B * Ts, 2 * E)
                K.linear_wgrad(dz2, s.out, out=gv(hp + "upsampler.weight").view(2 * E, E), accumulate=True)
                dxb = add_b(K.linear_dgrad(dz2, W[f"h{l}.wup"].view(2 * E, E)), dxb)
            else:  # the head is the Linear alone: z is the layer output itself
                dxb = add_b(dz, dxb)
        if dlayers is not None and dlayers[l] is not None:
            dxb = add_b(dlayers[l], dxb)
        if dx32 is None and dxb is None:
            continue
        M_ = B * Ts
        # final LayerNorm.  With dropout3 active the bf16 copy is dy2 * mask: it feeds the fc2 branch (wgrad, dgrad, bias
        # gradient), the un-masked fp32 dy2 continues down the residual.
        drop = getattr(c, "drop", None)
        dl = (lambda which: drop.layer(l, which)) if drop is not None else (lambda which: None)
        xs = getattr(s, "xs", None)
        dys = torch.empty_like(xs) if xs is not None else None  # [dy1m, du, dy2m]: A operands of the batched wgrad
        d3 = dl(DropCfg.DROP3)
        dy2m = dys[2] if dys is not None else torch.empty(M_, E, device=dev, dtype=bf16)
        dy2_32 = torch.empty(M_, E, device=dev, dtype=f32)
        K.layernorm_bwd32(dx32, s.y2, P[p + "final_layer_norm.weight"], s.mean2, s.rstd2,
                          gv(p + "final_layer_norm.weight"), gv(p + "final_layer_norm.bias"), dy2=dxb,
                          dx=None if d3 is not None else dy2m, dx32=dy2_32, dxsum=gv(p + "fc2.bias"),
                          dx_drop=dy2m if d3 is not None else None, drop=d3)
        dxb = None
        # FFN (fc2 bias gradient = column sums of dy2m: accumulated by the LayerNorm backward above; the saved
        # s.u = gelu'(u) * activation-dropout mask, s.h = dropped activations)
        if dys is None:
            with aside(dy2m):
                K.linear_wgrad(dy2m, s.h, out=gv(p + "fc2.weight").view(E, F), accumulate=True)
        du = K.linear_dgrad(dy2m, W[f"l{l}.w2"].view(E, F), mul_aux=s.u, out=None if dys is None else dys[1])
        with aside(du):
            K.colsum(du, gv(p + "fc1.bias"))
            if dys is None:
                K.linear_wgrad(du, s.x1, out=gv(p + "fc1.weight").view(F, E), accumulate=True)
        dx1_32 = K.linear_dgrad(du, W[f"l{l}.w1"].view(F, E), residual=dy2_32, out_dtype=f32)
        # attention LayerNorm (same scheme for dropout1 on the out_proj branch)
        d1 = dl(DropCfg.DROP1)
        dy1m = dys[0] if dys is not None else torch.empty(M_, E, device=dev, dtype=bf16)
        dy1_32 = dy2_32  # its last reader (the fc1 dgrad epilogue above) is queued before this writer
        K.layernorm_bwd32(dx1_32, s.y1, P[p + "self_attn_layer_norm.weight"], s.mean1, s.rstd1,
                          gv(p + "self_attn_layer_norm.weight"), gv(p + "self_attn_layer_norm.bias"),
                          dx=None if d1 is not None else dy1m, dx32=dy1_32, dxsum=gv(p + "self_attn.out_proj.bias"),
                          dx_drop=dy1m if d1 is not None else None, drop=d1)
        # attention block (out_proj bias gradient: accumulated by the LayerNorm backward above)
        with aside(dys if dys is not None else dy1m):
            if dys is not None:
                # dW[j] += dys[j]^T xs[j] for j = out_proj, fc1, fc2: one batched split-K launch (ob = j); the three
                # gradient segments are adjacent in the flat buffer (GradStore order)
                a3 = L.tensor3(dys, data_ptr=dys.data_ptr(), dim=(E, M_, 3), stride=(E, M_ * E))
                b3 = L.tensor3(xs, data_ptr=xs.data_ptr(), dim=(E, M_, 3), stride=(E, M_ * E))
                assert G_.entries[p + "fc1.weight"][0] - G_.entries[p + "self_attn.out_proj.weight"][0] == E * E and \
                    G_.entries[p + "fc2.weight"][0] - G_.entries[p + "fc1.weight"][0] == E * E
                K.gemm_raw(a3, b3, G_.from_(p + "self_attn.out_proj.weight"), E, E, M_, a_major=1, b_major=1, num_ob=3,
                           a_coord=(0, 1, 0, 0), b_coord=(0, 1, 0, 0), d_ld=E, d_hi_stride=E * E, flags=L.EPI_ATOMIC_ADD)
            else:
                K.linear_wgrad(dy1m, s.attn, out=gv(p + "self_attn.out_proj.weight").view(E, E), accumulate=True)
        dattn = K.linear_dgrad(dy1m, W[f"l{l}.wo"].view(E, E))
        dqkv = torch.empty(M_, 3 * E, device=dev, dtype=bf16)
        delta = torch.empty(B, H, Ts, device=dev, dtype=f32)
        dq_ws = torch.empty(M_, E, device=dev, dtype=f32) if d in (40, 64) else None
        K.attn_bwd(s.qkv, c.valid_s, s.attn, dattn, s.lse, dqkv, delta, B, Ts, H, d, d ** -0.5, drop=dl(DropCfg.ATTN),
                   dq_ws=dq_ws)
        if l == g.n_layers - 1 and getattr(c, "attn_grad", None):
            attn_transfer_backward(c, g, s.qkv, dqkv)  # attention-map / value-relation terms (train.py:327-368)
        with aside(dqkv):
            K.colsum(dqkv, G_.span(p + "self_attn.q_proj.bias", p + "self_attn.v_proj.bias"))
            K.linear_wgrad(dqkv, s.x, out=G_.span(p + "self_attn.q_proj.weight", p + "self_attn.v_proj.weight").view(3 * E, E),
                           accumulate=True)
        if on_progress is not None:
            # every second layer (and after the last one, whose call also covers the heads) a new part of the buffer is
            # handed over; the calls in between only tell the caller that one more layer's worth of kernels has been
            # queued since (it stops reserving SMs for the exchange it started a layer ago)
            issue = l % 2 == 0 or l == g.n_layers - 1
            if issue:
                aside.join()
            on_progress(G_.entries[p + "self_attn.q_proj.weight"][0] if issue else None)
        if l > 0:
            dx32 = K.linear_dgrad(dqkv, W[f"l{l}.wqkv"].view(3 * E, E), residual=dy1_32, out_dtype=f32)
        else:  # what leaves the encoder layers feeds bf16 GEMM operands (TR conv / prologue backward)
            dx = K.linear_dgrad(dqkv, W[f"l{l}.wqkv"].view(3 * E, E), residual=dy1_32)
            dx32 = None
    if on_layers_done is not None:
        # every gradient of the heads and of the transformer layers is final from here on (85 % of the bytes): the
        # data-parallel exchange of that part can start under the front-end / conv-stack backward below
        aside.join()
        on_layers_done()
    if dx is None:
        aside.join()
        return
    if g.tr:
        # ---- time-reduction conv backward
        K.colsum(dx, gv("encoder.layers.0.bias"))
        dy3 = L.tensor3(dx, data_ptr=dx.data_ptr(), dim=(E, Ts, B), stride=(E, Ts * E))
        x3 = L.tensor3(c.enc_in, data_ptr=c.enc_in.data_ptr(), dim=(2 * E, Ts, B), stride=(2 * E, T * E))
        _wgrad(dy3, x3, gv("encoder.layers.0.weight").view(E, 2 * E), E, 2 * E, Ts, num_cb=B, a_cb=1, b_cb=1)
        denc = torch.empty(B * T, E, device=dev, dtype=bf16)
        if T % 2:  # the odd tail frame was dropped by the TR conv: zero gradient
            K.zero_rows(denc, (T - 1) * E, T * E, E, B)
        a3 = L.tensor3(dx, data_ptr=dx.data_ptr(), dim=(E, Ts, B), stride=(E, Ts * E))
        b3 = L.tensor3(W["tr.w"], data_ptr=W["tr.w"].data_ptr(), dim=(2 * E, E, 1), stride=(2 * E, 2 * E * E))
        K.gemm_raw(a3, b3, denc, Ts, 2 * E, E, b_major=1, num_ob=B, a_coord=(0, 1, 0, 0), d_ld=2 * E, d_hi_stride=T * E)
    else:
        denc = dx  # no TR layer: the first transformer layer reads the prologue output directly
    # ---- encoder prologue backward: (dropout,) LayerNorm, + pos-conv residual, GELU, grouped conv, mask
    drop = getattr(c, "drop", None)
    d_pro = drop.site(DropCfg.SITE_PROLOGUE, drop.p_drop) if drop is not None else None
    if d_pro is not None:
        K.dropout(denc, denc, *d_pro)
    G, cp, kp, dl = g.G, g.cp, g.kpos, g.pdelta
    Tp, R = c.Tp, c.R
    dh = torch.empty(B * T, E, device=dev, dtype=bf16)
    pad_b = kp // 2 - 1
    dcg = torch.empty(B * G, Tp, cp, device=dev, dtype=bf16)
    K.zero_rows(dcg, 0, Tp * cp, pad_b * cp, B * G)  # borders only: rows [pad_b, pad_b + T) are written below
    K.zero_rows(dcg, (pad_b + T) * cp, Tp * cp, (Tp - pad_b - T) * cp, B * G)
    K.posconv_finish_bwd(denc, c.h, c.conv, P["encoder.pos_conv.0.bias"], P["encoder.layer_norm.weight"], c.mean_e,
                         c.rstd_e, dh, dcg, gv("encoder.layer_norm.weight"), gv("encoder.layer_norm.bias"),
                         gv("encoder.pos_conv.0.bias"), B, T, E, G, cp, pad_b, Tp, delta=dl)
    if on_progress is not None:
        on_progress(None)
    dxc = torch.empty(B * R, G * dl * cp, device=dev, dtype=bf16)  # [b][r][g][dl][cp], like the forward output
    a3 = L.tensor3(dcg, data_ptr=dcg.data_ptr(), dim=(g.kpx * cp, R, B * G), stride=(dl * cp, Tp * cp))
    b3 = L.tensor3(W["pc.wt"], data_ptr=W["pc.wt"].data_ptr(), dim=(g.kpx * cp, dl * cp, G), stride=(g.kpx * cp, dl * cp * g.kpx * cp))
    K.gemm_raw(a3, b3, dxc, R, dl * cp, g.kpx * cp, num_ob=B * G, ob_mod=G, a_coord=(0, G, 1, 0), b_coord=(0, 0, 1, 0),
               d_ld=G * dl * cp, d_hi_stride=R * G * dl * cp, d_lo_stride=dl * cp)
    # wgrad, time-blocked like the forward: dwt'[g][(j',ci)][(dl,co)] = sum_{b,r} xg[b,g,dl_*r + j',ci] *
    # dcg[b,g,dl_*r + dl + pad_b,co]  (N = dl_*cp instead of cp); posconv_wn_bwd folds the dl_ shifted partials
    dwt = torch.empty(G, g.kpx * cp, dl * cp, device=dev, dtype=f32)
    a3 = L.tensor3(c.xg, data_ptr=c.xg.data_ptr(), dim=(g.kpx * cp, R, B * G), stride=(dl * cp, Tp * cp))
    b3 = L.tensor3(dcg, data_ptr=dcg.data_ptr() + 2 * pad_b * cp, dim=(dl * cp, R, B * G), stride=(dl * cp, Tp * cp))
    K.gemm_raw(a3, b3, dwt, g.kpx * cp, dl * cp, R, a_major=1, b_major=1, num_ob=G, ob_mod=G, num_cb=B,
               a_coord=(0, 0, 1, G), b_coord=(0, 0, 1, G), d_ld=dl * cp, d_lo_stride=g.kpx * cp * dl * cp, split_k=1)
    K.posconv_wn_bwd(dwt, P["encoder.pos_conv.0.weight_v"], P["encoder.pos_conv.0.weight_g"], W["pc.ws"],
                     gv("encoder.pos_conv.0.weight_v"), gv("encoder.pos_conv.0.weight_g"), E, G, kp, cp, True, delta=dl)
    dfeat = torch.empty(B * T, E, device=dev, dtype=bf16)
    K.posconv_unpack_bwd(dh, dxc, c.valid_t, dfeat, B, T, E, G, cp, delta=dl)
    d_in = drop.site(DropCfg.SITE_INPUT, drop.p_input) if drop is not None else None
    if d_in is not None:
        K.dropout(dfeat, dfeat, *d_in)
    if dfeatures is not None:
        # gradient of the CNN-feature loss (train.py:241-246) joins the gradient of `features` AFTER the input dropout
        # (features_to_distill is taken before it, modules/model.py:483-489)
        if getattr(c, "cnn_out", None) is not None:
            Dc = dfeatures.shape[-1]
            dfe = dfeatures.view(B * T, Dc)
            K.colsum(dfe, gv("cnn_proj_head.1.bias"))
            K.linear_wgrad(dfe, c.cnn_g, out=gv("cnn_proj_head.1.weight").view(Dc, E), accumulate=True)
            dfe = K.linear_dgrad(dfe, W["cnn.w"].view(Dc, E), dgelu_of=c.feats)  # x gelu'(features)
        else:
            dfe = dfeatures.view(B * T, E)
        dfeat = K.add_bf16(dfeat, dfe, torch.empty_like(dfeat))
    # ---- post_extract_proj + LayerNorm(C_feat)
    Cf = g.c_feat
    K.colsum(dfeat, gv("post_extract_proj.bias"))
    K.linear_wgrad(dfeat, c.f_ln.view(B * T, Cf), out=gv("post_extract_proj.weight").view(E, Cf), accumulate=True)
    dfl = K.linear_dgrad(dfeat, W["pp.w"].view(E, Cf))
    n_conv = len(g.conv_layers)
    last = n_conv - 1
    dyl = torch.empty(B * T, Cf, device=dev, dtype=bf16)
    K.layernorm_bwd(dfl, c.out, P["layer_norm.weight"], c.mean_f, c.rstd_f, dyl, gv("layer_norm.weight"),
                    gv("layer_norm.bias"))
    if on_progress is not None:
        # everything behind the conv stack in the buffer is final: it goes out under the conv-stack backward
        on_progress(G_.entries["layer_norm.weight"][0])
    # dU_last = dY * gelu'(U_last)   (c.u holds the saved gelu' values)
    du = torch.empty(B, T, Cf, device=dev, dtype=bf16)
    K.mul_bf16(dyl, T * Cf, c.u[last], T * Cf, du, T * Cf, B, T * Cf, alpha=g.grad_mult)  # x feature_grad_mult
    # ---- conv stack backward, layers last .. 1.  du: [B, To + 2*halo_i, C_i], data rows start at halo_i
    for i in range(last, 0, -1):
        co, k, s = g.conv_layers[i]
        cin = g.conv_layers[i - 1][0]
        To, Tin = c.frames[i], c.frames[i - 1]
        halo, in_halo = c.halos[i], c.halos[i - 1]
        rows, in_rows = To + 2 * halo, Tin + 2 * in_halo
        xin = c.y[i - 1]
        du_data = du.data_ptr() + 2 * halo * co
        # wgrad: dW2[co][(j,ci)] += sum_{b,t} dU[b,t,co] * X[b, s*t + j, ci]
        dy3 = L.tensor3(du, data_ptr=du_data, dim=(co, To, B), stride=(co, rows * co))
        x3 = L.tensor3(xin, data_ptr=xin.data_ptr() + 2 * in_halo * cin, dim=(k * cin, To, B), stride=(s * cin, in_rows * cin))
        with aside_conv(du):
            _wgrad(dy3, x3, gv(f"feature_extractor.conv_layers.{i}.0.weight").view(co, k * cin), co, k * cin, To,
                   num_cb=B, a_cb=1, b_cb=1)
        if on_progress is not None and not (mode & 4):
            # the two widest layers (half of the stack's parameters) and then the layers down to 3 are handed over as
            # soon as their weight gradients are queued; only ~1 MB (layers 2, 1, 0) is left for after the backward
            on_progress(G_.entries[f"feature_extractor.conv_layers.{i}.0.weight"][0] if i in (last - 1, 3) else None)
        # dU_{i-1} = dX_{i-1} * gelu'(U_{i-1}) (layer 0 included: its forward saved gelu' of the GroupNorm output)
        dprev = torch.empty(B, in_rows, cin, device=dev, dtype=bf16)
        flags, uprev = L.EPI_MUL_AUX, c.u[i - 1]
        if in_halo:  # halo rows must read as zero in the overlapped-view dgrad of layer i-1
            K.zero_rows(dprev, 0, in_rows * cin, cin, B)
            K.zero_rows(dprev, (in_rows - 1) * cin, in_rows * cin, cin, B)
        covered = Tin if (k, s) != (2, 2) else 2 * To
        if covered < Tin:  # frames no output depends on (odd tail of a k=2,s=2 layer)
            K.zero_rows(dprev, (in_halo + covered) * cin, in_rows * cin, (Tin - covered) * cin, B)
        if (k, s) == (1, 1) or (k, s) == (2, 2):
            # dA[b,t,(j,ci)] = dU[b,t,:] W2[:, (j,ci)]  ==  dX[b, s*t + j, ci]   (W2 consumed MN-major)
            a3 = L.tensor3(du, data_ptr=du_data, dim=(co, To, B), stride=(co, rows * co))
            b3 = L.tensor3(W[f"conv{i}.w"], data_ptr=W[f"conv{i}.w"].data_ptr(), dim=(k * cin, co, 1), stride=(k * cin, k * cin * co))
            K.gemm_raw(a3, b3, dprev, To, k * cin, co, b_major=1, num_ob=B, a_coord=(0, 1, 0, 0), d_ld=k * cin,
                       d_hi_stride=in_rows * cin, d_offset_elems=in_halo * cin, flags=flags, aux_in=uprev)
        else:  # (3, 2): even input frames see taps (2, 0) of outputs (u-1, u); odd frames tap 1 of output u
            n_even, n_odd = (Tin + 1) // 2, Tin // 2
            a3 = L.tensor3(du, data_ptr=du.data_ptr(), dim=(2 * co, n_even, B), stride=(co, rows * co))
            b3 = L.tensor3(W[f"conv{i}.wd_even"], data_ptr=W[f"conv{i}.wd_even"].data_ptr(), dim=(2 * co, cin, 1), stride=(2 * co, 2 * co * cin))
            K.gemm_raw(a3, b3, dprev, n_even, cin, 2 * co, num_ob=B, a_coord=(0, 1, 0, 0), d_ld=2 * cin,
                       d_hi_stride=in_rows * cin, d_offset_elems=in_halo * cin, flags=flags, aux_in=uprev)
            a3 = L.tensor3(du, data_ptr=du_data, dim=(co, n_odd, B), stride=(co, rows * co))
            b3 = L.tensor3(W[f"conv{i}.wd_odd"], data_ptr=W[f"conv{i}.wd_odd"].data_ptr(), dim=(co, cin, 1), stride=(co, co * cin))
            K.gemm_raw(a3, b3, dprev, n_odd, cin, co, num_ob=B, a_coord=(0, 1, 0, 0), d_ld=2 * cin,
                       d_hi_stride=in_rows * cin, d_offset_elems=(in_halo + 1) * cin, flags=flags, aux_in=uprev)
        du = dprev
    # ---- layer 0: conv + GroupNorm + GELU backward (dW0, dgamma, dbeta).  du already carries gelu'(z) (dz): the sums
    # over frames sum_t dz x[5t+j] and sum_t dz are a wgrad-shaped GEMM against a bf16 im2col of the waveform
    # (C a multiple of 64: the MN-major A operand is loaded in 64-channel atoms); other widths take the direct kernel.
    C0 = g.conv_layers[0][0]
    wv, gm, bt = (P["feature_extractor.conv_layers.0.0.weight"], P["feature_extractor.conv_layers.0.2.weight"],
                  P["feature_extractor.conv_layers.0.2.bias"])
    gw, gg, gb = (gv("feature_extractor.conv_layers.0.0.weight"), gv("feature_extractor.conv_layers.0.2.weight"),
                  gv("feature_extractor.conv_layers.0.2.bias"))
    T0 = c.frames[0]
    if C0 % 64 == 0:
        xcol = torch.empty(B, T0, 32, device=dev, dtype=bf16)
        K.conv0_im2col(c.wave, T0, xcol)
        acc32 = torch.zeros(B, C0, 32, device=dev, dtype=f32)
        a3 = L.tensor3(du, data_ptr=du.data_ptr(), dim=(C0, T0, B), stride=(C0, T0 * C0))
        b3 = L.tensor3(xcol, data_ptr=xcol.data_ptr(), dim=(32, T0, B), stride=(32, T0 * 32))
        K.gemm_raw(a3, b3, acc32, C0, 32, T0, a_major=1, b_major=1, num_ob=B, a_coord=(0, 1, 0, 0), b_coord=(0, 1, 0, 0),
                   d_ld=32, d_hi_stride=C0 * 32, flags=L.EPI_ATOMIC_ADD)
        K.conv0_bwd_finalize(acc32, c.wave, wv, gm, bt, T0, c.stat, c.mean0, c.rstd0, gw, gg, gb, True)
    else:
        acc = torch.empty(B, C0, 12, device=dev, dtype=f32)
        K.conv0_bwd(c.wave, wv, gm, bt, T0, c.stat, c.mean0, c.rstd0, du, acc, gw, gg, gb, True, dy_is_dz=True)
    aside.join()
    aside_conv.join()
