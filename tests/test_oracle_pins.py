"""Independent pins of the oracle (CPU).

1. torchaudio.models.hubert_base: an implementation of HuBERT-Base written by other people, fed the oracle's random
   teacher weights through torchaudio's OWN fairseq-key converter (import_fairseq._convert_state_dict).  This pins the
   teacher's top-level wiring (conv stack -> LayerNorm -> post_extract_proj -> mask -> pos-conv -> 12 post-LN layers),
   which in oracle/gen_golden.py is builder-written around the reference's classes (SURVEY 8c pin 2).
2. The reference's own modules/model.py + modules/module.py, executed unmodified through oracle/fairseq_stub, against the
   oracle on inputs and geometries the committed fixtures do not contain - live, whenever /root/reference exists (it does
   in the build container; the GPU box has only the fixtures)."""
import os
import sys

import pytest
import torch

import fhb_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FHB_REFERENCE", "/root/reference")


def rel(a, b):
    return float((a.detach() - b.detach()).abs().max() / b.detach().abs().max().clamp_min(1e-12))


def test_teacher_matches_torchaudio_hubert_base():
    ta = pytest.importorskip("torchaudio")
    from torchaudio.models.wav2vec2.utils.import_fairseq import _convert_state_dict
    tcfg = O.teacher_config()  # HuBERT-Base geometry: 7 conv layers, 12 x (768, 3072, 12 heads), pos-conv 128 / 16
    sd = O.init_teacher_state(tcfg, 1, perturb=True)
    model = ta.models.hubert_base()
    conv = _convert_state_dict(dict(sd))
    missing, unexpected = model.load_state_dict(conv, strict=False)
    # torchaudio's model has no parameters the fairseq `features_only` trunk lacks (mask_emb / final_proj are dropped)
    assert not [k for k in missing if "mask" not in k], missing
    assert not unexpected, unexpected
    model.eval()
    x, _ = O.synth_batch(2, 24000, [24000, 24000], seed=3)
    with torch.no_grad():
        feats, _ = model.extract_features(x)  # list of the 12 layer outputs, [B, T, 768]
        ours = O.teacher_forward(sd, tcfg, x, None)
    assert len(feats) == 12
    for l in range(12):
        assert rel(ours["layer_results"][l][0].transpose(0, 1), feats[l]) < 2e-5, l
    # with padding: torchaudio masks by conv-output lengths (wav2vec 2.0's rule M1, not HuBERT's M3), so the wav2vec2
    # flavour of the oracle teacher is the one to compare; valid frames only
    lens = [24000, 17000]
    x, pm = O.synth_batch(2, 24000, lens, seed=4)
    w2v = dict(tcfg, kind="wav2vec2")
    with torch.no_grad():
        feats, out_len = model.extract_features(x, torch.tensor(lens))
        ours = O.teacher_forward(sd, w2v, x, pm)
    valid = (~ours["padding_mask"]).sum(-1).tolist()
    assert out_len.tolist() == valid
    for l in (0, 5, 11):
        for b in range(2):
            assert rel(ours["layer_results"][l][0][:valid[b], b], feats[l][b, :valid[b]]) < 2e-5, (l, b)


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "modules", "model.py")), reason="needs /root/reference")
def test_oracle_matches_the_reference_classes_live():
    sys.path[:0] = [os.path.join(ROOT, "oracle", "fairseq_stub"), REF, os.path.join(ROOT, "oracle")]
    import yaml
    import gen_golden as GG  # imports the reference's modules/model.py and modules/module.py unmodified
    with open(os.path.join(REF, "data/conf/fithubert.yaml")) as f:
        ycfg = yaml.safe_load(f)["distiller"]
    # the REAL FitHuBERT geometry (12 layers, D = 480, 30-channel pos-conv groups), which no fixture holds
    torch.manual_seed(0)
    student = GG.CustomStudentModel(GG.ref_student_cfg(ycfg, init_conv_layers=False, init_encoder_layers=0))
    GG.perturb_(student, 5)
    student.eval()
    sd = {k: v.detach().clone() for k, v in student.state_dict().items()}
    scfg = O.student_config()
    x, pm = O.synth_batch(2, 20000, [20000, 13777], seed=9)
    ref = student(source=x, padding_mask=pm)
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ours = O.student_forward(p, scfg, x, pm)
    assert torch.equal(ours["padding_mask"], ref["padding_mask"])
    for l in range(12):
        assert rel(ours["layer_results"][l][0], ref["layer_results"][l][0]) < 1e-5, l
        assert rel(ours["layer_results"][l][2], ref["layer_results"][l][2]) < 1e-5, l
        assert rel(ours["projections"][l], ref["projections"][l]) < 1e-5, l
    assert rel(ours["features"], ref["features"]) < 1e-5 and rel(ours["x"], ref["x"]) < 1e-5
    # gradients through both, same scalar
    ref["x"].pow(2).mean().backward()
    ours["x"].pow(2).mean().backward()
    for n, q in student.named_parameters():
        if q.grad is None:
            assert p[n].grad is None, n
            continue
        if q.grad.abs().max() < 1e-10:
            continue
        assert rel(p[n].grad, q.grad) < 1e-4, n
    # after _disable_projection_heads (the s3prl expert's state) and with the `layer=` early exit
    student._disable_projection_heads()
    with torch.no_grad():
        r2 = student(source=x, padding_mask=pm)
        o2 = O.student_forward(sd, scfg, x, pm, heads=False)
        assert r2["projections"] is None and rel(o2["x"], r2["x"]) < 1e-5
        r3 = student.extract_features(x, pm, layer=3)
        assert len(r3["layer_results"]) == 3  # entry 0 of encoder.layers is the time-reduction conv
        assert rel(o2["layer_results"][2][0], r3["layer_results"][2][0]) < 1e-5
    # teacher: the reference's own ConvFeatureExtractionModel / TransformerEncoder wired as HuBERT (gen_golden.RefTeacher)
    tcfg = O.teacher_config(encoder_layers=3)
    teacher = GG.RefTeacher(tcfg)
    GG.perturb_(teacher, 6)
    teacher.eval()
    with torch.no_grad():
        tr = teacher(x, pm)
        to = O.teacher_forward({k: v for k, v in teacher.state_dict().items()}, tcfg, x, pm)
    assert torch.equal(to["padding_mask"], tr["padding_mask"])
    for l in range(3):
        assert rel(to["layer_results"][l][0], tr["layer_results"][l][0]) < 1e-5
        assert rel(to["layer_results"][l][1][1], tr["layer_results"][l][1][1]) < 1e-5
