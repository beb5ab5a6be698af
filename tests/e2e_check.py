"""Stage-by-stage GPU check of the engine against the golden fixtures (reference outputs).
Usage: python tests/e2e_check.py [fixture-name ...]"""
import glob
import os
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import fhb_oracle as O
import fithubert_b200 as F
from fithubert_b200 import engine as E, kernels as K

dev = "cuda"


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def show(name, a, b, tol):
    e = rel(a, b)
    print(f"{'OK ' if e < tol else 'BAD'} {name}: rel={e:.3e}", flush=True)
    return e < tol


def student_cfg(over):
    import yaml
    d = dict(
        extractor_mode="default", conv_feature_layers=O.FITHUBERT_CONV, feature_grad_mult=1.0, conv_bias=False,
        conv_pos=128, conv_pos_groups=16, pos_conv_depth=1, layer_type="transformer", encoder_layers=12,
        encoder_embed_dim=480, encoder_ffn_embed_dim=480, encoder_attention_heads=12, activation_fn="gelu",
        layer_norm_first=False, dropout=0.1, attention_dropout=0.1, activation_dropout=0.1, encoder_layerdrop=0.0,
        dropout_input=0.05, final_dim=256, pred_head_final_dim=768, pred_head_inter_dim=0, layerwise_proj=True,
        pred_layer_id="[11]", enable_tr_layer=True, tr_conv1d_kernel=2, tr_layer_index=0, tr_reduce_factor=2,
        tr_layer_type="conv1d", required_seq_len_multiple=1, crop_seq_to_multiple=1)
    d.update(over)
    return F.CustomStudentModelConfig(**d)


def run(path):
    print("=====", os.path.basename(path), flush=True)
    g = torch.load(path)
    ok = True
    tc = g["teacher_cfg"]
    teacher = F.TeacherModel(kind=tc["kind"], conv_feature_layers=tc["conv_feature_layers"],
                             encoder_embed_dim=tc["encoder_embed_dim"], encoder_ffn_embed_dim=tc["encoder_ffn_embed_dim"],
                             encoder_attention_heads=tc["encoder_attention_heads"], encoder_layers=tc["encoder_layers"],
                             conv_pos=tc["conv_pos"], conv_pos_groups=tc["conv_pos_groups"])
    teacher.load_state_dict(g["teacher_state"])
    teacher = F.TeacherWrapper(teacher.to(dev))
    x, pm = g["source"], g["padding_mask"]
    tr = teacher.extract_features(x.to(dev), pm)
    torch.cuda.synchronize()
    tv = tr["_valid"]
    ref_tv = None if g["teacher_mask"] is None else (~g["teacher_mask"]).sum(-1).tolist()
    print("teacher valid", tv, ref_tv, "OK" if tv == ref_tv else "BAD")
    ok &= tv == ref_tv
    ok &= show("teacher features", tr["features"][0], g["teacher_features"], 2e-2)
    for i, ref in enumerate(g["teacher_layers"]):
        ok &= show(f"teacher layer {i}", tr["layer_results"][i][0], ref, 2e-2)

    student = F.CustomStudentModel(student_cfg(g["student_cfg"]))
    student.load_state_dict(g["student_state"])
    student = student.to(dev)
    with torch.no_grad():
        sr = student(x.to(dev), pm)
    torch.cuda.synchronize()
    m, rm = sr["padding_mask"], g["student_mask"]
    same = (m is None and rm is None) or (m is not None and rm is not None and torch.equal(m.cpu(), rm))
    print("student mask", "OK" if same else "BAD")
    ok &= same
    ok &= show("student tr", sr["tr_layer_results"][0], g["student_tr"], 2e-2)
    for i, ref in enumerate(g["student_layers"]):
        ok &= show(f"student layer {i}", sr["layer_results"][i][0], ref, 2e-2)
    for i, ref in enumerate(g["projections"]):
        ok &= show(f"projection {i}", sr["projections"][i], ref, 2e-2)

    # fused training path
    P, W, G = student.engine_state(True)
    G.zero_()
    s_valid = None if g["student_mask"] is None else (~g["student_mask"]).sum(-1).tolist()
    c = E.student_forward(P, W, student._geom, x.to(dev), s_valid, train=True, heads="all")
    n, B, Tq, D = c.preds.shape
    tgt = tr["_stacked"]
    ll = torch.zeros(n, device=dev)
    w = torch.tensor(g["layer_weights"], device=dev)
    for i, ref in enumerate(g["projections"]):
        ok &= show(f"train-mode projection {i}", c.preds[i], ref, 2e-2)
    K.distill_loss(c.preds, tgt, w, ll, c.preds, n, B, Tq, tgt.shape[2], D, 0, 1.0)
    torch.cuda.synchronize()
    ok &= show("per-layer loss", ll, g["per_layer"], 2e-2)
    print("loss", float(ll.sum()), float(g["loss"]))
    E.student_backward(P, W, student._geom, G, c, c.preds)
    grads = G.export()
    torch.cuda.synchronize()
    worst = 0.0
    for nme, ref in g["grads"].items():
        if ref.abs().max() < 1e-9:
            continue
        e = rel(grads[nme], ref)
        worst = max(worst, e)
        if e > 5e-2:
            print(f"BAD grad {nme}: rel={e:.3e} |ref|max={float(ref.abs().max()):.3e}", flush=True)
            ok = False
    print(f"worst grad rel err {worst:.3e}; params without grad: {G.no_grad} (ref {g['no_grad_params']})", flush=True)
    print("RESULT", "PASS" if ok else "FAIL", flush=True)
    return ok


if __name__ == "__main__":
    names = sys.argv[1:] or sorted(glob.glob("tests/golden/tiny_*.pt"))
    allok = True
    for nme in names:
        allok &= run(nme if nme.endswith(".pt") else f"tests/golden/{nme}.pt")
    sys.exit(0 if allok else 1)
