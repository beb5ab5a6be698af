"""Which bf16 rounding site dominates the deep-layer deviation?  Emulates the CUDA path's storage roundings in the
student's 12 encoder layers (weights bf16, fp32 accumulate) and toggles sites off one at a time.
Analysis script (checker side: it runs the oracle, hence it lives under tests/; not collected by pytest).
Usage: python tests/precision_sites.py   -> profiles/r01final_precision_sites.txt"""
import sys, torch, torch.nn.functional as F
sys.path.insert(0, 'oracle'); sys.path.insert(0, '.')
import fhb_oracle as O
torch.manual_seed(0)
scfg = O.student_config(); sd = O.init_student_state(scfg, 0)
Lm = 160000
x = 0.1 * torch.randn(1, Lm)
with torch.no_grad():
    ref = O.student_forward(sd, scfg, x, None, heads=False)
    tr = ref["tr_layer_results"][0]  # [Ts,1,E] input of layer 0
H = 12
q16 = lambda t: t.bfloat16().float()
def run(sites):
    r = lambda name, t: q16(t) if name in sites else t
    wq = lambda w: q16(w) if "w" in sites else w
    xx = r("x", tr)
    outs = []
    for i in range(12):
        p = f"encoder.layers.{i+1}."
        T, B, E = xx.shape; d = E // H
        a = p + "self_attn."
        q = r("qkv", F.linear(xx, wq(sd[a+"q_proj.weight"]), sd[a+"q_proj.bias"])) * d ** -0.5
        k = r("qkv", F.linear(xx, wq(sd[a+"k_proj.weight"]), sd[a+"k_proj.bias"]))
        v = r("qkv", F.linear(xx, wq(sd[a+"v_proj.weight"]), sd[a+"v_proj.bias"]))
        q, k, v = (t.reshape(T, B*H, d).transpose(0, 1) for t in (q, k, v))
        pr = r("p", torch.softmax(torch.bmm(q, k.transpose(1, 2)), -1))
        at = r("attn", torch.bmm(pr, v).transpose(0, 1).reshape(T, B, E))
        y1 = r("y", xx + F.linear(at, wq(sd[a+"out_proj.weight"]), sd[a+"out_proj.bias"]))
        x1 = r("ln", F.layer_norm(y1, (E,), sd[p+"self_attn_layer_norm.weight"], sd[p+"self_attn_layer_norm.bias"], 1e-5))
        h = r("h", F.gelu(F.linear(x1, wq(sd[p+"fc1.weight"]), sd[p+"fc1.bias"])))
        y2 = r("y", x1 + F.linear(h, wq(sd[p+"fc2.weight"]), sd[p+"fc2.bias"]))
        xx = r("ln", F.layer_norm(y2, (E,), sd[p+"final_layer_norm.weight"], sd[p+"final_layer_norm.bias"], 1e-5))
        outs.append(xx)
    return outs
def rel(a, b): return float((a - b).abs().max() / b.abs().max())
ALL = {"x", "w", "qkv", "p", "attn", "y", "ln", "h"}
with torch.no_grad():
    base = run(set())
    print("fp32 re-run vs oracle l11:", rel(base[11], ref["layer_results"][11][0]))
    full = run(ALL)
    print("all sites   l0 %.4f l5 %.4f l11 %.4f" % tuple(rel(full[i], base[i]) for i in (0, 5, 11)))
    for s in sorted(ALL):
        o = run(ALL - {s})
        print("without %-5s l0 %.4f l5 %.4f l11 %.4f" % ((s,) + tuple(rel(o[i], base[i]) for i in (0, 5, 11))))
    for s in sorted(ALL):
        o = run({s})
        print("only    %-5s l0 %.4f l5 %.4f l11 %.4f" % ((s,) + tuple(rel(o[i], base[i]) for i in (0, 5, 11))))
