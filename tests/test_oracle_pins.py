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


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train.py")), reason="needs /root/reference")
def test_loss_restatements_match_the_reference_calculate_loss_live():
    """The oracle's loss functions against the reference's OWN `W2V2Distil.calculate_loss` (train.py:236-405), compiled from
    the unmodified source by oracle/ref_extract.py (train.py itself cannot be imported: Lightning / s3prl are absent) and
    called on synthetic result dicts: the random-layer MSE recipe (fithubert.yaml), L1 + cosine over pred_layer_id
    (ex.yaml), the CNN-feature L1 term, and the attention-map (mse, kldiv) / value-relation terms."""
    import ref_extract as R
    g = torch.Generator().manual_seed(3)
    B, n, T, D, H = 2, 4, 13, 24, 3
    proj = [torch.randn(B, T - 1, D, generator=g) for _ in range(n)]
    t_layers = [torch.randn(T, B, D, generator=g) for _ in range(n)]

    def attn_pair(valid):
        a = torch.randn(B * H, T, T, generator=g)
        for b, v in enumerate(valid):
            a[b * H:(b + 1) * H, :, v:] = float("-inf")
        return a, torch.randn(B * H, T, T, generator=g)

    s_attn, s_vrel = attn_pair([T, T - 4])
    t_attn, t_vrel = attn_pair([T, T - 3])  # HuBERT's mask rule keeps one more frame than the student's
    feats, t_feats = torch.randn(B, T, D, generator=g), torch.randn(B, T, D, generator=g)
    s_res = {"projections": proj, "features": feats, "layer_results": [(None, None, None)] * (n - 1) + [(None, (s_attn, s_vrel), None)]}
    t_res = {"layer_results": [(t_layers[i], (None, None)) for i in range(n - 1)] + [(t_layers[-1], ((t_attn, t_vrel), None))],
             "features": [t_feats]}
    base = dict(cnn_loss_weight=0, rec_loss_weight=1.0, rec_loss_type="mse", sim_loss_weight=0, attn_loss_weight=0,
                attn_loss_type="kldiv", v_rel_loss_weight=0, distil_random_layer=0, random_layer_weight=0)
    # (1) fithubert.yaml: every lower layer picked at random_layer_weight, the last at 1
    rand_l = [2, 0, 1]
    tc = dict(base, distil_random_layer=n - 1, random_layer_weight=0.1)
    ref, ref_losses = R.ref_calculate_loss(s_res, t_res, train_cfg=tc, model_cfg={"layerwise_proj": True},
                                           pred_layer_id=[n - 1], rand_l=rand_l, num_encoders=n)
    ours, per = O.distill_loss(proj, t_res["layer_results"], O.layer_weights(n, 0.1))
    assert abs(float(ours) - float(ref)) < 1e-6 * float(ref)
    for i, l in enumerate(rand_l):  # logged in the permuted order of rand_l (train.py:319-321)
        assert abs(float(per[l]) - float(ref_losses[f"rand_l{i}"])) < 1e-6
    assert abs(float(per[-1]) - float(ref_losses[f"l{n - 1}"])) < 1e-6
    # (2) ex.yaml: L1 + cosine over pred_layer_id, plus the CNN-feature term
    ids = [0, 2]
    tc = dict(base, rec_loss_type="l1", sim_loss_weight=0.7, rec_loss_weight=1.3, cnn_loss_weight=0.5)
    ref, ref_losses = R.ref_calculate_loss(s_res, t_res, train_cfg=tc, model_cfg={"layerwise_proj": True}, pred_layer_id=ids,
                                           num_encoders=n)
    ours, rec, sim = O.distill_loss_sim(proj, t_res["layer_results"], ids, "l1", 1.3, 0.7)
    cnn = O.cnn_feature_loss(feats, t_feats)
    assert abs(float(ours + 0.5 * cnn) - float(ref)) < 1e-6 * float(ref)
    assert abs(float(cnn) - float(ref_losses["cnn_loss"])) < 1e-6
    for k, i in enumerate(ids):
        assert abs(float(rec[k] + sim[k]) - float(ref_losses[f"layer{i}"])) < 1e-6
    # (3) attention map (mse on padded logits) + value relation
    tc = dict(base, attn_loss_weight=2.0, attn_loss_type="mse", v_rel_loss_weight=3.0)
    ref, ref_losses = R.ref_calculate_loss(s_res, t_res, train_cfg=tc, model_cfg={"layerwise_proj": True}, pred_layer_id=ids,
                                           num_encoders=n)
    rec_only, _, _ = O.distill_loss_sim(proj, t_res["layer_results"], ids, "mse", 1.0, 0.0)
    a, v = O.attn_map_loss(s_attn, t_attn, "mse"), O.value_relation_loss(s_vrel, t_vrel)
    assert abs(float(a) - float(ref_losses["attn_loss"])) < 1e-6 * float(a)
    assert abs(float(v) - float(ref_losses["v_rel_loss"])) < 1e-6 * float(v)
    assert abs(float(rec_only + 2.0 * a + 3.0 * v) - float(ref)) < 1e-6 * float(ref)
    # the mean runs over the keys neither side masks: B*H*T * min(valid) columns
    d2 = (s_attn - t_attn)[..., :T - 4].pow(2)
    expect = (d2[:H].sum() + (s_attn - t_attn)[:H, :, T - 4:].pow(2).sum() + d2[H:].sum()) / (H * T * (T + T - 4))
    assert abs(float(a) - float(expect)) < 1e-6 * float(a)
    # (4) kldiv: nan on a padded batch in the reference (documented quirk), equal on an un-padded one
    tc = dict(base, attn_loss_weight=1.0, attn_loss_type="kldiv")
    ref, ref_losses = R.ref_calculate_loss(s_res, t_res, train_cfg=tc, model_cfg={"layerwise_proj": True}, pred_layer_id=ids,
                                           num_encoders=n)
    assert torch.isnan(ref_losses["attn_loss"]) and torch.isnan(O.attn_map_loss(s_attn, t_attn, "kldiv"))
    s2, t2 = torch.randn(B * H, T, T, generator=g), torch.randn(B * H, T, T, generator=g)
    s_res["layer_results"][-1] = (None, (s2, s_vrel), None)
    t_res["layer_results"][-1] = (t_layers[-1], ((t2, t_vrel), None))
    ref, ref_losses = R.ref_calculate_loss(s_res, t_res, train_cfg=tc, model_cfg={"layerwise_proj": True}, pred_layer_id=ids,
                                           num_encoders=n)
    a = O.attn_map_loss(s2, t2, "kldiv")
    assert abs(float(a) - float(ref_losses["attn_loss"])) < 1e-6 * float(a)
    # v_rel_loss_weight > 0 without attn_loss_weight: the stock layers return None where the pair would be
    s_res["layer_results"][-1] = (None, None, None)
    with pytest.raises(TypeError):
        R.ref_calculate_loss(s_res, t_res, train_cfg=dict(base, v_rel_loss_weight=1.0), model_cfg={"layerwise_proj": True},
                             pred_layer_id=ids, num_encoders=n)
