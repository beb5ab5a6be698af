"""Generate tests/golden/*.pt by running the UNMODIFIED reference files.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py

It imports reference modules/model.py + modules/module.py through oracle/fairseq_stub,
builds (a) a *tiny* student (same architecture family as data/conf/fithubert.yaml with
small widths so the weights fit in a fixture) and (b) a tiny HuBERT-style teacher
composed from the reference's own ConvFeatureExtractionModel / TransformerEncoder
classes (they are copies of fairseq wav2vec2.py; top-level wiring per SURVEY App. B.2),
runs them on seeded synthetic waveforms with padding, and stores weights, inputs and
outputs.  tests/test_oracle_golden.py replays the fixtures through oracle/fhb_oracle.py
(CPU) and tests/test_gpu_parity.py through the CUDA path.
"""
import os
import sys
import types
import warnings

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FHB_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "fairseq_stub"), REF]
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

from modules.model import CustomStudentModel, CustomStudentModelConfig  # noqa: E402  (reference)
from modules.module import ConvFeatureExtractionModel, TransformerEncoder  # noqa: E402  (reference)
import fhb_oracle as O  # noqa: E402

only_cases = []  # filled from argv by main(): when non-empty, only these fixtures are (re)generated

TINY_STUDENT = dict(
    conv_feature_layers="[(16, 10, 5)] + [(32, 1, 1)] + [(32, 3, 2)] * 4 + [(64, 1, 1)] + [(64, 2, 2)] * 2",
    encoder_layers=3, encoder_embed_dim=96, encoder_ffn_embed_dim=96, encoder_attention_heads=4,
    conv_pos=16, conv_pos_groups=4, pred_head_final_dim=64,
)
TINY_TEACHER = dict(
    conv_feature_layers="[(32,10,5)] + [(32,3,2)] * 4 + [(32,2,2)] * 2",
    encoder_layers=3, encoder_embed_dim=64, encoder_ffn_embed_dim=128, encoder_attention_heads=4,
    conv_pos=16, conv_pos_groups=4,
)


def ref_student_cfg(yaml_distiller: dict, **over):
    d = dict(yaml_distiller)
    d.update(over)
    return CustomStudentModelConfig(**d)


class RefTeacher(nn.Module):
    """HuBERT/wav2vec2 `features_only` top level wired from reference classes."""

    def __init__(self, cfg: dict):
        super().__init__()
        layers = O.parse_conv_layers(cfg["conv_feature_layers"])
        self.kind = cfg.get("kind", "hubert")
        self.conv_layers = layers
        self.feature_extractor = ConvFeatureExtractionModel(layers, 0.0, "default", False)
        self.layer_norm = nn.LayerNorm(layers[-1][0])
        self.post_extract_proj = nn.Linear(layers[-1][0], cfg["encoder_embed_dim"])
        args = types.SimpleNamespace(
            dropout=0.1, encoder_embed_dim=cfg["encoder_embed_dim"], required_seq_len_multiple=1,
            pos_conv_depth=1, conv_pos=cfg["conv_pos"], conv_pos_groups=cfg["conv_pos_groups"],
            enable_tr_layer=False, encoder_layers=cfg["encoder_layers"], layer_type="transformer",
            encoder_ffn_embed_dim=cfg["encoder_ffn_embed_dim"],
            encoder_attention_heads=cfg["encoder_attention_heads"], attention_dropout=0.1,
            activation_dropout=0.0, activation_fn="gelu", layer_norm_first=False,
            checkpoint_activations=False, encoder_layerdrop=0.0)
        self.encoder = TransformerEncoder(args)

    def forward(self, source, padding_mask):
        f = self.feature_extractor(source).transpose(1, 2)
        f = self.layer_norm(f)
        T = f.shape[1]
        if self.kind == "hubert":
            mask = O.mask_m3(padding_mask, T)
        else:
            mask = O.mask_m1(padding_mask, T, self.conv_layers)
        f = self.post_extract_proj(f)
        x, layer_results, _ = self.encoder(f.clone(), padding_mask=mask)
        # the forward hook on every encoder layer captures its output (x, (attn, layer_result)), utils/utils.py:65-78;
        # `attn` is None on the stock layer and (attn_logits, v_rel) once rtrn_attn_forward is bound (train.py:64-69)
        return {"layer_results": [(lr[0], (lr[1], lr[2])) for lr in layer_results],
                "features": [f], "padding_mask": mask}


def perturb_(module, seed):
    """Biases / norm affines are 0/1 at reference init: randomise so they matter."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in module.named_parameters():
            if n.endswith(".bias") or "layer_norm.weight" in n or n.endswith("conv_layers.0.2.weight"):
                p.add_(torch.empty_like(p).normal_(0.0, 0.05, generator=g))


def run_case(name, s_over, t_over, B, Lmax, lengths, yaml_distiller, kind="hubert", grads=True):
    if only_cases and name not in only_cases:
        return
    torch.manual_seed(0)
    student = CustomStudentModel(ref_student_cfg(yaml_distiller, **s_over))
    tcfg = O.teacher_config(**t_over, kind=kind)
    teacher = RefTeacher(tcfg)
    perturb_(student, 11)
    perturb_(teacher, 12)
    student.eval()  # dropout = identity (parity runs use p = 0)
    teacher.eval()
    x, pm = O.synth_batch(B, Lmax, lengths, seed=1234)
    with torch.no_grad():
        t_res = teacher(x, pm)
    s_res = student(source=x, padding_mask=pm)
    n_layers = len(s_res["projections"])
    w = O.layer_weights(n_layers, 0.1)
    # calculate_loss rec branch restated inline from train.py:250-293 (train.py itself
    # cannot be imported: pytorch_lightning / s3prl are absent)
    pred = torch.stack(s_res["projections"], 1)
    tgt = torch.stack([lr[0].transpose(0, 1) for lr in t_res["layer_results"]], 1).narrow(2, 0, pred.shape[2])
    rec = torch.nn.functional.mse_loss(pred, tgt, reduction="none")
    rec[:, :-1] = rec[:, :-1] * 0.1
    per_layer = rec.mean((0, 2, 3))
    loss = per_layer.sum()
    out = {
        "student_cfg": s_over, "teacher_cfg": dict(t_over, kind=kind),
        "student_state": {k: v.detach().clone() for k, v in student.state_dict().items()},
        "teacher_state": {k: v.detach().clone() for k, v in teacher.state_dict().items()},
        "source": x, "padding_mask": pm, "layer_weights": w,
        "student_mask": s_res["padding_mask"], "teacher_mask": t_res["padding_mask"],
        "student_features": s_res["features"].detach(),
        "student_layers": [lr[0].detach() for lr in s_res["layer_results"]],
        "student_tr": s_res["tr_layer_results"][0].detach() if s_res["tr_layer_results"] else None,
        "projections": [p.detach() for p in s_res["projections"]],
        "teacher_layers": [lr[0].detach() for lr in t_res["layer_results"]],
        "teacher_features": t_res["features"][0].detach(),
        "loss": loss.detach(), "per_layer": per_layer.detach(),
    }
    if grads:
        loss.backward()
        out["grads"] = {n: p.grad.detach().clone() for n, p in student.named_parameters() if p.grad is not None}
        out["no_grad_params"] = [n for n, p in student.named_parameters() if p.grad is None]
    path = os.path.join(HERE, "..", "tests", "golden", name + ".pt")
    torch.save(out, path)
    print(name, "loss", float(loss), "bytes", os.path.getsize(path))


TINY_SPLIT_STUDENT = dict(  # data/conf/ex.yaml family: teacher-shaped conv stack, no TR layer, DistilHuBERT head
    conv_feature_layers="[(32,10,5)] + [(32,3,2)] * 4 + [(32,2,2)] * 2",
    encoder_layers=2, encoder_embed_dim=64, encoder_ffn_embed_dim=128, encoder_attention_heads=4,
    conv_pos=16, conv_pos_groups=4, pred_head_final_dim=64, pred_layer_id="[0, 2]",
    init_conv_layers=False, init_encoder_layers=0,
)


def run_case_split(name, s_over, t_over, B, Lmax, lengths, yaml_distiller):
    """ex.yaml recipe: layerwise_proj False, enable_tr_layer False, feature_grad_mult 0.1; loss = L1 + cosine over
    pred_layer_id (train.py:268-314, `distil_random_layer == 0` branch, restated inline like run_case's)."""
    if only_cases and name not in only_cases:
        return
    torch.manual_seed(0)
    cfg = ref_student_cfg(yaml_distiller, **s_over)
    assert not cfg.layerwise_proj and not cfg.enable_tr_layer and cfg.feature_grad_mult == 0.1
    student = CustomStudentModel(cfg)
    teacher = RefTeacher(O.teacher_config(**t_over, kind="hubert"))
    perturb_(student, 21)
    perturb_(teacher, 22)
    student.eval()
    teacher.eval()
    x, pm = O.synth_batch(B, Lmax, lengths, seed=4321)
    with torch.no_grad():
        t_res = teacher(x, pm)
    s_res = student(source=x, padding_mask=pm)
    ids = eval(cfg.pred_layer_id)
    tgt = torch.stack([t_res["layer_results"][i][0].transpose(0, 1) for i in ids], 1)
    pred = s_res["projections"]
    tgt = tgt.narrow(2, 0, pred.shape[2])
    rec = torch.nn.functional.l1_loss(pred, tgt, reduction="none")
    sim = -torch.nn.functional.logsigmoid(torch.nn.functional.cosine_similarity(pred, tgt, dim=-1))
    loss = 1.0 * rec.mean() + 1.0 * sim.mean()
    loss.backward()
    out = {
        "student_cfg": dict(s_over, layerwise_proj=False, enable_tr_layer=False, feature_grad_mult=0.1),
        "teacher_cfg": dict(t_over, kind="hubert"),
        "student_state": {k: v.detach().clone() for k, v in student.state_dict().items()},
        "teacher_state": {k: v.detach().clone() for k, v in teacher.state_dict().items()},
        "source": x, "padding_mask": pm, "pred_layer_id": ids,
        "student_mask": s_res["padding_mask"], "teacher_mask": t_res["padding_mask"],
        "student_features": s_res["features"].detach(),
        "student_layers": [lr[0].detach() for lr in s_res["layer_results"]],
        "x": s_res["x"].detach(), "projections": pred.detach(),
        "teacher_layers": [lr[0].detach() for lr in t_res["layer_results"]],
        "loss": loss.detach(), "rec_layer": rec.mean((0, 2, 3)).detach(), "sim_layer": sim.mean((0, 2)).detach(),
        "grads": {n: p.grad.detach().clone() for n, p in student.named_parameters() if p.grad is not None},
        "no_grad_params": [n for n, p in student.named_parameters() if p.grad is None],
    }
    path = os.path.join(HERE, "..", "tests", "golden", name + ".pt")
    torch.save(out, path)
    print(name, "loss", float(loss), "bytes", os.path.getsize(path))


TINY_UP_STUDENT = dict(  # layerwise_proj False WITH a TR layer (shared upsampler) + CNN-feature head (D != E, _cnn_weight > 0)
    conv_feature_layers="[(16, 10, 5)] + [(32, 3, 2)] * 4 + [(32, 2, 2)] * 2",
    encoder_layers=2, encoder_embed_dim=96, encoder_ffn_embed_dim=96, encoder_attention_heads=4,
    conv_pos=16, conv_pos_groups=4, pred_head_final_dim=64, pred_layer_id="[0, 2]",
    init_conv_layers=False, init_encoder_layers=0, enable_tr_layer=True,
)
CNN_LOSS_WEIGHT = 0.5


def run_case_upsampler_cnn(name, s_over, t_over, B, Lmax, lengths, yaml_distiller):
    """ex.yaml family with enable_tr_layer True: the encoder output goes through the model's shared `upsampler`
    (modules/model.py:341-348,402-404,504-505) before the DistilHuBERT head, and - pred_head_final_dim != encoder_embed_dim,
    _cnn_weight > 0 - `features` come out of cnn_proj_head (modules/model.py:304-310,486-487).  Loss = L1 + cosine over
    pred_layer_id (train.py:268-314) + cnn_loss_weight * L1(features, teacher features[0]) (train.py:241-246,372-378),
    restated inline like the other cases.
    The batch is un-padded on purpose: with a padding mask and dropout_input an identity (eval mode / p = 0) the reference's
    OWN backward raises - the encoder's in-place index_put (modules/module.py:273-274) overwrites the tensor cnn_proj_head's
    GELU saved ("modified by an inplace operation"); it only trains with dropout_input live, whose mask cannot be replayed."""
    if only_cases and name not in only_cases:
        return
    torch.manual_seed(0)
    cfg = ref_student_cfg(yaml_distiller, **s_over)
    cfg._cnn_weight = CNN_LOSS_WEIGHT  # train.py:43
    assert not cfg.layerwise_proj and cfg.enable_tr_layer
    student = CustomStudentModel(cfg)
    assert student.cnn_proj_head is not None and student.upsampler is not None
    teacher = RefTeacher(O.teacher_config(**t_over, kind="hubert"))
    perturb_(student, 31)
    perturb_(teacher, 32)
    student.eval()
    teacher.eval()
    x, pm = O.synth_batch(B, Lmax, lengths, seed=2468)
    with torch.no_grad():
        t_res = teacher(x, pm)
    s_res = student(source=x, padding_mask=pm)
    ids = eval(cfg.pred_layer_id)
    tgt = torch.stack([t_res["layer_results"][i][0].transpose(0, 1) for i in ids], 1)
    pred = s_res["projections"]
    tgt = tgt.narrow(2, 0, pred.shape[2])
    rec = torch.nn.functional.l1_loss(pred, tgt, reduction="none")
    sim = -torch.nn.functional.logsigmoid(torch.nn.functional.cosine_similarity(pred, tgt, dim=-1))
    cnn = torch.nn.functional.l1_loss(s_res["features"], t_res["features"][0], reduction="none").mean()
    loss = 1.0 * rec.mean() + 1.0 * sim.mean() + CNN_LOSS_WEIGHT * cnn
    loss.backward()
    out = {
        "student_cfg": dict(s_over, layerwise_proj=False, feature_grad_mult=cfg.feature_grad_mult),
        "teacher_cfg": dict(t_over, kind="hubert"), "cnn_loss_weight": CNN_LOSS_WEIGHT,
        "student_state": {k: v.detach().clone() for k, v in student.state_dict().items()},
        "teacher_state": {k: v.detach().clone() for k, v in teacher.state_dict().items()},
        "source": x, "padding_mask": pm, "pred_layer_id": ids,
        "student_mask": s_res["padding_mask"], "teacher_mask": t_res["padding_mask"],
        "student_features": s_res["features"].detach(), "teacher_features": t_res["features"][0].detach(),
        "student_layers": [lr[0].detach() for lr in s_res["layer_results"]],
        "student_tr": s_res["tr_layer_results"][0].detach(),
        "x": s_res["x"].detach(), "projections": pred.detach(),
        "teacher_layers": [lr[0].detach() for lr in t_res["layer_results"]],
        "loss": loss.detach(), "rec_layer": rec.mean((0, 2, 3)).detach(), "sim_layer": sim.mean((0, 2)).detach(),
        "cnn_loss": cnn.detach(),
        "grads": {n: p.grad.detach().clone() for n, p in student.named_parameters() if p.grad is not None},
        "no_grad_params": [n for n, p in student.named_parameters() if p.grad is None],
    }
    path = os.path.join(HERE, "..", "tests", "golden", name + ".pt")
    torch.save(out, path)
    print(name, "loss", float(loss), "cnn", float(cnn), "bytes", os.path.getsize(path))


TINY_ATTN_STUDENT = dict(  # ex.yaml family (no TR layer: the attention-map recipe needs matching frame rates and heads)
    conv_feature_layers="[(32,10,5)] + [(32,3,2)] * 4 + [(32,2,2)] * 2",
    encoder_layers=2, encoder_embed_dim=96, encoder_ffn_embed_dim=128, encoder_attention_heads=4,
    conv_pos=16, conv_pos_groups=4, pred_head_final_dim=64, pred_layer_id="[0, 2]",
    init_conv_layers=False, init_encoder_layers=0,
)


def run_case_attn(name, s_over, t_over, B, Lmax, lengths, yaml_distiller, yaml_train, attn_loss_type, attn_w, vrel_w):
    """Attention-map / value-relation distillation: ex.yaml recipe (L1 + cosine over pred_layer_id) with
    attn_loss_weight / v_rel_loss_weight > 0.  As train.py:64-77 does, `rtrn_attn_forward` - compiled from the unmodified
    utils/utils.py (oracle/ref_extract.py) - is bound over every teacher and student encoder layer; the loss is the
    reference's OWN `W2V2Distil.calculate_loss` (train.py:236-405), compiled the same way and called on the two result
    dicts, so nothing in this fixture is builder-restated arithmetic."""
    if only_cases and name not in only_cases:
        return
    import ref_extract as R
    torch.manual_seed(0)
    cfg = ref_student_cfg(yaml_distiller, **s_over)
    assert not cfg.layerwise_proj and not cfg.enable_tr_layer
    student = CustomStudentModel(cfg)
    teacher = RefTeacher(O.teacher_config(**t_over, kind="hubert"))
    perturb_(student, 41)
    perturb_(teacher, 42)
    R.bind_attn_forward(teacher.encoder.layers)
    R.bind_attn_forward(student.encoder.layers)
    student.eval()
    teacher.eval()
    x, pm = O.synth_batch(B, Lmax, lengths, seed=1357)
    with torch.no_grad():
        t_res = teacher(x, pm)
    s_res = student(source=x, padding_mask=pm)
    ids = eval(cfg.pred_layer_id)
    train_cfg = dict(yaml_train, attn_loss_weight=attn_w, attn_loss_type=attn_loss_type, v_rel_loss_weight=vrel_w)
    loss, losses = R.ref_calculate_loss(s_res, t_res, train_cfg=train_cfg, model_cfg=dict(yaml_distiller, **s_over),
                                        pred_layer_id=ids, num_encoders=cfg.encoder_layers)
    loss.backward()
    s_attn, s_vrel = s_res["layer_results"][-1][1]
    t_attn, t_vrel = t_res["layer_results"][-1][1][0]
    out = {
        "student_cfg": dict(s_over, layerwise_proj=False, enable_tr_layer=False, feature_grad_mult=cfg.feature_grad_mult),
        "teacher_cfg": dict(t_over, kind="hubert"),
        "train_cfg": {k: train_cfg[k] for k in ("cnn_loss_weight", "rec_loss_weight", "rec_loss_type", "sim_loss_weight",
                                                "attn_loss_weight", "attn_loss_type", "v_rel_loss_weight",
                                                "distil_random_layer", "random_layer_weight")},
        "student_state": {k: v.detach().clone() for k, v in student.state_dict().items()},
        "teacher_state": {k: v.detach().clone() for k, v in teacher.state_dict().items()},
        "source": x, "padding_mask": pm, "pred_layer_id": ids,
        "student_mask": s_res["padding_mask"], "teacher_mask": t_res["padding_mask"],
        "student_layers": [lr[0].detach() for lr in s_res["layer_results"]],
        "x": s_res["x"].detach(), "projections": s_res["projections"].detach(),
        "teacher_layers": [lr[0].detach() for lr in t_res["layer_results"]],
        "student_attn": s_attn.detach(), "student_vrel": s_vrel.detach(),
        "teacher_attn": t_attn.detach(), "teacher_vrel": t_vrel.detach(),
        "loss": loss.detach(), "losses": {k: v.detach() for k, v in losses.items()},
        "grads": {n: p.grad.detach().clone() for n, p in student.named_parameters() if p.grad is not None},
        "no_grad_params": [n for n, p in student.named_parameters() if p.grad is None],
    }
    path = os.path.join(HERE, "..", "tests", "golden", name + ".pt")
    torch.save(out, path)
    print(name, "loss", float(loss), {k: round(float(v), 6) for k, v in losses.items()}, "bytes", os.path.getsize(path))


def main():
    import yaml
    only_cases.extend(sys.argv[1:])  # optional: names of the cases to (re)generate
    with open(os.path.join(REF, "data/conf/fithubert.yaml")) as f:
        ycfg = yaml.safe_load(f)["distiller"]
    # 1.5 - 2 s utterances (75 - 99 frames): parameter gradients are sums over frames, and with the 0.5 s / 27-frame
    # utterances of round 1 the max-norm deviation of a bf16 run on the smallest tensors (conv layer 0, GroupNorm) was
    # dominated by that handful of frames rather than by the arithmetic under test
    run_case("tiny_hubert_pad", TINY_STUDENT, TINY_TEACHER, 3, 32000, [32000, 27411, 21000], ycfg)
    run_case("tiny_hubert_nopad", TINY_STUDENT, TINY_TEACHER, 2, 24000, [24000, 24000], ycfg)
    run_case("tiny_w2v2_pad_oddT", TINY_STUDENT, TINY_TEACHER, 2, 28100, [28100, 17321], ycfg, kind="wav2vec2")
    # layer-wise heads WITHOUT a time-reduction layer: LayerWiseProjHead is the Linear alone (modules/module.py:633-646),
    # the layers run at the full frame rate with mask rule M1 un-reduced, projections are [B, T, D] (no T' narrowing)
    run_case("notr_layerwise_hubert_pad", dict(TINY_STUDENT, enable_tr_layer=False), TINY_TEACHER, 3, 24000, [24000, 20411, 15000],
             ycfg)
    with open(os.path.join(REF, "data/conf/ex.yaml")) as f:
        ecfg = yaml.safe_load(f)["distiller"]
    run_case_split("split_hubert_pad", TINY_SPLIT_STUDENT, TINY_TEACHER, 3, 32000, [32000, 27411, 21000], ecfg)
    run_case_upsampler_cnn("upsampler_cnn_hubert_nopad", TINY_UP_STUDENT, TINY_TEACHER, 2, 28000, [28000, 28000], ecfg)
    with open(os.path.join(REF, "data/conf/ex.yaml")) as f:
        etrain = yaml.safe_load(f)["train"]
    # mse on a padded batch (HuBERT teacher: the two sides mask different key columns); kldiv + value relation un-padded
    # (with padding the reference's kldiv branch is nan, see oracle attn_map_loss)
    run_case_attn("attn_mse_hubert_pad", TINY_ATTN_STUDENT, TINY_TEACHER, 3, 20000, [20000, 17411, 13000], ecfg, etrain,
                  "mse", 200.0, 300.0)
    run_case_attn("attn_kldiv_hubert_nopad", TINY_ATTN_STUDENT, TINY_TEACHER, 2, 18000, [18000, 18000], ecfg, etrain,
                  "kldiv", 500.0, 300.0)


if __name__ == "__main__":
    main()
