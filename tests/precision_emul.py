"""CPU emulation of the CUDA path's bf16 STORAGE roundings over one whole distillation step (teacher forward, student
forward, loss, student backward): every tensor the engine writes to HBM in bf16 is rounded here at the same place, in
the forward value and/or in the gradient, and sites can be switched to fp32 one at a time.  Used to decide which
tensors have to be carried at higher precision to meet north_star's 2e-2 (profiles/r02_precision_emul.txt).

Analysis script (checker side: it runs the oracle's init / loss helpers, hence it lives under tests/; not collected by
pytest).  Usage: python tests/precision_emul.py [tiny|full] [seconds]"""
import sys
import os
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import fhb_oracle as O  # noqa: E402


def q16(t):
    return t.bfloat16().float()


def qh(t):
    return t.half().float()


class RQ(torch.autograd.Function):
    """value rounded to bf16 (f = True) or fp16 (f = "h") in the forward and / or gradient rounded to bf16 in the
    backward (b)"""

    @staticmethod
    def forward(ctx, t, f, b):
        ctx.b = b
        return qh(t) if f == "h" else (q16(t) if f else t.clone())

    @staticmethod
    def backward(ctx, g):
        return (q16(g) if ctx.b else g), None, None


class GeluSaved(torch.autograd.Function):
    """y = gelu(pre); the backward multiplies by gelu'(pre) as SAVED by the forward epilogue (bf16 when `rounded`)"""

    @staticmethod
    def forward(ctx, pre, rounded):
        with torch.enable_grad():
            p = pre.detach().requires_grad_(True)
            y = F.gelu(p)
            (gp,) = torch.autograd.grad(y.sum(), p)
        ctx.save_for_backward(qh(gp) if rounded == "h" else (q16(gp) if rounded else gp))
        return y.detach()

    @staticmethod
    def backward(ctx, g):
        (gp,) = ctx.saved_tensors
        return g * gp, None


class AttnCore(torch.autograd.Function):
    """softmax(q k^T + mask) v per head with the kernels' operand roundings: P (bf16 MMA operand of P V and of dV),
    dS (bf16 operand of dQ / dK)."""

    @staticmethod
    def forward(ctx, q, k, v, bias, rp, rds, ro=False, rdo=False):
        s = torch.bmm(q, k.transpose(1, 2))
        if bias is not None:
            s = s + bias
        p = torch.softmax(s, -1)
        pr = q16(p) if rp else p
        o = torch.bmm(pr, v)
        ctx.save_for_backward(q, k, v, p, pr, q16(o) if ro else o)
        ctx.rds, ctx.rdo = rds, rdo
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, p, pr, o = ctx.saved_tensors
        if ctx.rdo:
            do = q16(do)
        dv = torch.bmm(pr.transpose(1, 2), do)
        dp = torch.bmm(do, v.transpose(1, 2))
        delta = (do * o).sum(-1, keepdim=True)  # the kernels' attn_delta: row sums of dO * O as stored
        ds = p * (dp - delta)
        if ctx.rds:
            ds = q16(ds)
        return torch.bmm(ds, k), torch.bmm(ds.transpose(1, 2), q), dv, None, None, None, None, None


class Sites:
    """fwd / bwd: sites stored in 16 bits; f16: the forward sites among them that use fp16 instead of bf16"""

    def __init__(self, fwd, bwd, f16=()):
        self.fwd, self.bwd, self.f16 = set(fwd), set(bwd), set(f16)

    def fmt(self, f):
        return ("h" if f in self.f16 else True) if f in self.fwd else False

    def r(self, t, f=None, b=None):
        return RQ.apply(t, self.fmt(f), b in self.bwd)

    def w(self, w):
        return RQ.apply(w, self.fmt("w"), False)


# every bf16 storage site of the round-1 CUDA path
FWD_ALL = {"w", "conv_y", "conv_u", "fln", "feats", "pconv", "h", "enc", "tr", "qkv", "p", "attn", "y", "ln", "lnop",
           "ffh", "ffu", "wc", "pred", "tgt"}
BWD_ALL = {"dyop", "dpred", "dwc", "dxhead", "dln", "dlnop", "du", "dx", "dattn", "dqkv", "ds", "denc", "dh", "dcg", "dxc", "dfeat",
           "dfl", "dyl", "dconv"}


def conv_stack(S, sd, x, conv_layers):
    x = x.unsqueeze(1)
    for i, (c, k, s) in enumerate(conv_layers):
        w = sd[f"feature_extractor.conv_layers.{i}.0.weight"]
        if i == 0:
            x = F.conv1d(x, w, None, stride=s)
            x = F.group_norm(x, c, sd["feature_extractor.conv_layers.0.2.weight"], sd["feature_extractor.conv_layers.0.2.bias"], 1e-5)
        else:
            x = F.conv1d(x, S.w(w), None, stride=s)
        x = S.r(x, None, "dconv")  # dU = (dgrad) * gelu' is stored in bf16
        x = GeluSaved.apply(x, S.fmt("conv_u"))
        x = S.r(x, "conv_y", None)
    return x


def layer(S, sd, p, x, bias, H):
    """x: (residual-stream value, GEMM-operand copy).  Sites: 'ln' = the LayerNorm output as read back by the residual
    add, 'lnop' = the bf16 copy the next GEMM consumes, 'y' = the pre-LayerNorm sum (LayerNorm input, saved for bwd)."""
    T, B, E = x.shape
    d = E // H
    a = p + "self_attn."
    xo = S.r(x, "lnop", "dlnop")
    qkv = [F.linear(xo, S.w(sd[a + f"{n}_proj.weight"]), sd[a + f"{n}_proj.bias"]) for n in "qkv"]
    q, k, v = (S.r(t, "qkv", "dqkv").reshape(T, B * H, d).transpose(0, 1) for t in qkv)
    at = AttnCore.apply(q * d ** -0.5, k, v, bias, "p" in S.fwd, "ds" in S.bwd, "attn" in S.fwd, "dattn" in S.bwd)
    at = S.r(at.transpose(0, 1).reshape(T, B, E), "attn", "dattn")
    y1 = S.r(x + S.r(F.linear(at, S.w(sd[a + "out_proj.weight"]), sd[a + "out_proj.bias"]), None, "dyop"), "y", "dx")
    x1 = S.r(F.layer_norm(y1, (E,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"], 1e-5), "ln", "dln")
    x1o = S.r(x1, "lnop", "dlnop")
    u = S.r(F.linear(x1o, S.w(sd[p + "fc1.weight"]), sd[p + "fc1.bias"]), None, "du")
    h = S.r(GeluSaved.apply(u, S.fmt("ffu")), "ffh", None)
    y2 = S.r(x1 + S.r(F.linear(h, S.w(sd[p + "fc2.weight"]), sd[p + "fc2.bias"]), None, "dyop"), "y", "dx")
    return S.r(F.layer_norm(y2, (E,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], 1e-5), "ln", "dln")


def frontend(S, sd, cfg, source, mask, conv_layers):
    feats = conv_stack(S, sd, source, conv_layers).transpose(1, 2)
    C = feats.shape[-1]
    feats = S.r(feats, None, "dyl")
    feats = S.r(F.layer_norm(feats, (C,), sd["layer_norm.weight"], sd["layer_norm.bias"], 1e-5), "fln", "dfl")
    feats = S.r(F.linear(feats, S.w(sd["post_extract_proj.weight"]), sd["post_extract_proj.bias"]), "feats", "dfeat")
    x = feats.masked_fill(mask.unsqueeze(-1), 0.0) if mask is not None else feats
    v, g = sd["encoder.pos_conv.0.weight_v"], sd["encoder.pos_conv.0.weight_g"]
    w = S.w(g * v / v.norm(dim=(0, 1), keepdim=True))
    k = cfg["conv_pos"]
    xin = S.r(x, None, "dxc")
    y = F.conv1d(xin.transpose(1, 2), w, None, padding=k // 2, groups=cfg["conv_pos_groups"])[:, :, :-1]
    y = S.r(y, "pconv", "dcg").transpose(1, 2) + sd["encoder.pos_conv.0.bias"]
    h = S.r(x + F.gelu(y), "h", "dh")  # saved for the LayerNorm backward in bf16; the LayerNorm itself reads registers
    E = h.shape[-1]
    return S.r(F.layer_norm(h, (E,), sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"], 1e-5), "enc", "denc")


def key_bias(mask, B, H, T):
    if mask is None:
        return None
    b = torch.zeros(B, 1, 1, T).masked_fill(mask[:, None, None, :], float("-inf"))
    return b.expand(B, H, T, T).reshape(B * H, T, T)


def student(S, sd, cfg, source, pm):
    conv_layers = O.parse_conv_layers(cfg["conv_feature_layers"])
    H = cfg["encoder_attention_heads"]
    T = O.conv_out_lengths(torch.tensor([source.shape[1]]), conv_layers).item()
    mask = O.mask_m1(pm, T, conv_layers)
    x = frontend(S, sd, cfg, source, mask, conv_layers).transpose(0, 1)
    x = F.conv1d(x.permute(1, 2, 0), S.w(sd["encoder.layers.0.weight"]), sd["encoder.layers.0.bias"], stride=2).permute(2, 0, 1)
    x = S.r(x, "tr", "dx")
    rmask = O.mask_m2(mask, 2)
    Ts, B, E = x.shape
    bias = key_bias(rmask, B, H, Ts)
    outs, preds = [], []
    for i in range(cfg["encoder_layers"]):
        x = layer(S, sd, f"encoder.layers.{i + 1}.", x, bias, H)
        outs.append(x)
        # folded head: Wc_p = Wlin Wup_p (bf16), pred = x Wc + bc
        wup, bup = sd[f"proj_head.{i}.upsampler.weight"], sd[f"proj_head.{i}.upsampler.bias"]
        wlin, blin = sd[f"proj_head.{i}.lin_proj.weight"], sd[f"proj_head.{i}.lin_proj.bias"]
        xo = S.r(S.r(x, "lnop", "dlnop"), None, "dxhead")
        ph = []
        for ph_ in range(2):
            wc = S.r(S.w(wlin) @ S.w(wup[:, :, ph_]).t(), "wc", "dwc")  # [D, E_in]
            ph.append(F.linear(xo, wc, F.linear(q16(bup) if "w" in S.fwd else bup, S.w(wlin), blin)))
        pred = torch.stack(ph, 1).reshape(2 * Ts, B, -1).transpose(0, 1)  # [B, 2Ts, D], frame 2t+p
        preds.append(S.r(pred, "pred", "dpred"))
    return outs, preds


def teacher(S, sd, cfg, source, pm):
    conv_layers = O.parse_conv_layers(cfg["conv_feature_layers"])
    H = cfg["encoder_attention_heads"]
    T = O.conv_out_lengths(torch.tensor([source.shape[1]]), conv_layers).item()
    mask = O.mask_m3(pm, T) if cfg.get("kind", "hubert") == "hubert" else O.mask_m1(pm, T, conv_layers)
    x = frontend(S, sd, cfg, source, mask, conv_layers).transpose(0, 1)
    B = x.shape[1]
    bias = key_bias(mask, B, H, T)
    outs = []
    for i in range(cfg["encoder_layers"]):
        x = layer(S, sd, f"encoder.layers.{i}.", x, bias, H)
        outs.append(S.r(x, "tgt", None))
    return outs


def step(S, ssd, scfg, tsd, tcfg, x, pm, weights, St=None):
    sd = {k: v.clone().requires_grad_(True) for k, v in ssd.items()}
    with torch.no_grad():
        tl = teacher(St or S, tsd, tcfg, x, pm)
    outs, preds = student(S, sd, scfg, x, pm)
    pred = torch.stack(preds, 1)
    tgt = torch.stack([t.transpose(0, 1) for t in tl], 1)[:, :, :pred.shape[2]]
    w = torch.tensor(weights).view(1, -1, 1, 1)
    loss = (F.mse_loss(pred, tgt, reduction="none") * w).mean((0, 2, 3)).sum()
    loss.backward()
    return outs, tl, loss, {k: v.grad for k, v in sd.items() if v.grad is not None}


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def group(name):
    if name.startswith("feature_extractor"):
        return "conv"
    if name.startswith("encoder.layers.") and not name.startswith("encoder.layers.0."):
        return "layers"
    if name.startswith("proj_head"):
        return "heads"
    return "front"


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    if mode == "tiny":
        s_over = dict(conv_feature_layers="[(16, 10, 5)] + [(32, 1, 1)] + [(32, 3, 2)] * 4 + [(64, 1, 1)] + [(64, 2, 2)] * 2",
                      encoder_layers=3, encoder_embed_dim=96, encoder_ffn_embed_dim=96, encoder_attention_heads=4,
                      conv_pos=16, conv_pos_groups=4, pred_head_final_dim=64)
        t_over = dict(conv_feature_layers="[(32,10,5)] + [(32,3,2)] * 4 + [(32,2,2)] * 2", encoder_layers=3,
                      encoder_embed_dim=64, encoder_ffn_embed_dim=128, encoder_attention_heads=4, conv_pos=16, conv_pos_groups=4)
        lens = [9000, 7300, 5200]
        perturb = True
    else:
        s_over, t_over = {}, {}
        sec = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
        lens = [int(16000 * sec), int(16000 * sec * 0.8)]
        perturb = False
    scfg, tcfg = O.student_config(**s_over), O.teacher_config(**t_over)
    ssd, tsd = O.init_student_state(scfg, 0, perturb=perturb), O.init_teacher_state(tcfg, 1, perturb=perturb)
    x, pm = O.synth_batch(len(lens), lens[0], lens, seed=7)
    n = scfg["encoder_layers"]
    w = O.layer_weights(n, 0.1)
    base = step(Sites((), ()), ssd, scfg, tsd, tcfg, x, pm, w)

    def report(name, fwd, bwd, tfwd=None):
        outs, tl, loss, g = step(Sites(fwd, bwd), ssd, scfg, tsd, tcfg, x, pm, w, None if tfwd is None else Sites(tfwd, ()))
        eh = max(rel(a, b) for a, b in zip(outs, base[0]))
        et = max(rel(a, b) for a, b in zip(tl, base[1]))
        worst = {}
        for k in g:
            if base[3][k].abs().max() < 1e-9 or "k_proj.bias" in k:
                continue
            e = rel(g[k], base[3][k])
            gk = group(k)
            if e > worst.get(gk, ("", 0.0))[1]:
                worst[gk] = (k, e)
        print("%-46s hidden %.4f teacher %.4f loss %.1e | grads " % (name, eh, et, abs(float(loss - base[2]) / float(base[2]))) +
              " ".join("%s %.4f" % (gk, worst[gk][1]) for gk in ("conv", "front", "layers", "heads") if gk in worst), flush=True)
        return worst

    torch.manual_seed(0)
    wst = report("round-1 path (all sites bf16)", FWD_ALL - {"lnop"}, BWD_ALL - {"dlnop"})
    print("   worst:", {k: v[0] for k, v in wst.items()})
    stream_f = FWD_ALL - {"ln", "y"}          # LayerNorm outputs keep a high-precision copy for the residual, y fp32
    stream_b = BWD_ALL - {"dln", "dx"}
    report("fwd stream hp (student + teacher)", stream_f, BWD_ALL - {"dlnop"})
    report("fwd + bwd stream hp", stream_f, stream_b)
    for s in sorted(stream_f - {"lnop"}):
        report("  + fwd site fp32: " + s, stream_f - {s}, stream_b)
    for s in sorted(stream_b - {"dlnop"}):
        report("  + bwd site fp32: " + s, stream_f, stream_b - {s})
    report("fwd + bwd stream hp, teacher exact", stream_f, stream_b, set())
    report("fwd + bwd stream hp, teacher round-1", stream_f, stream_b, FWD_ALL - {"lnop"})


if __name__ == "__main__":
    main()
