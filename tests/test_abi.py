"""CPU-only: the C-ABI library builds/loads and exports every symbol include/fhb.h declares;
host-side integer logic (mask rules, config parsing, schedules) matches the oracle."""
import os
import re

import pytest
import torch

import fhb_oracle as O
from fithubert_b200 import lib as L
from fithubert_b200.config import CustomStudentModelConfig, parse_layer_spec
from fithubert_b200.model import conv_out_lengths, hubert_mask_lengths
from fithubert_b200.optim import warmup_linear

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from fithubert_b200 import build
    build.build()
    return L.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "fhb.h")).read()
    declared = set(re.findall(r"\b(fhb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in fhb.h but not exported"
    assert declared == set(L.EXPORTS)
    assert lib.fhb_abi_version() == 1


def test_arg_errors_do_not_need_a_gpu(lib):
    import ctypes as C
    g = L.GemmArgs()
    rc = lib.fhb_gemm(C.byref(g), None)
    assert rc == -1 and b"m,n,k" in lib.fhb_last_error()
    rc = lib.fhb_layernorm_fwd(None, None, None, None, None, None, C.c_int64(4), 480, C.c_float(1e-5), None)
    assert rc == -1 and b"null" in lib.fhb_last_error()
    rc = lib.fhb_distill_loss_fwd_bwd(C.c_void_p(8), C.c_void_p(8), C.c_void_p(8), C.c_void_p(8), None, None, C.c_int64(0),
                                      12, 2, 10, 11, 768, 7, C.c_float(1.0), None)
    assert rc == -1 and b"rec_loss_type must be one of 'l1', 'mse'." in lib.fhb_last_error()


def test_no_cpu_fallback():
    import fithubert_b200 as F
    m = F.CustomStudentModel(CustomStudentModelConfig(
        conv_feature_layers="[(16, 10, 5)] + [(32, 1, 1)] + [(64, 2, 2)] * 2", encoder_layers=1, encoder_embed_dim=32,
        encoder_ffn_embed_dim=32, encoder_attention_heads=2, conv_pos=16, conv_pos_groups=2, pred_head_final_dim=32,
        layerwise_proj=True, tr_layer_type="conv1d", tr_layer_index=0, required_seq_len_multiple=1))
    with pytest.raises(L.FhbError):
        m(torch.zeros(1, 4000))


def test_config_accepts_reference_yaml_keys_and_rejects_out_of_scope():
    import bench
    d = bench.yaml_cfg()["distiller"]
    cfg = CustomStudentModelConfig(**d)
    cfg.validate_hot_path()
    assert parse_layer_spec(cfg.conv_feature_layers) == O.parse_conv_layers(O.FITHUBERT_CONV)
    with pytest.raises(NotImplementedError):
        CustomStudentModelConfig(**dict(d, tr_layer_type="fc3")).validate_hot_path()
    with pytest.raises(NotImplementedError):
        CustomStudentModelConfig(**dict(d, layer_type="conformer")).validate_hot_path()


def test_state_dict_keys_match_reference_contract():
    import bench
    import fithubert_b200 as F
    m = F.CustomStudentModel(CustomStudentModelConfig(**bench.yaml_cfg()["distiller"]))
    ref = O.init_student_state(O.student_config())
    assert set(m.state_dict()) == set(ref)
    assert all(m.state_dict()[k].shape == v.shape for k, v in ref.items())
    assert sum(p.numel() for p in m.parameters()) == 31629632
    m._disable_projection_heads()
    assert sum(p.numel() for p in m.parameters()) == 22492064
    t = F.TeacherModel()
    assert set(t.state_dict()) == set(O.init_teacher_state(O.teacher_config()))


def test_mask_rules_bit_exact_vs_oracle():
    conv = O.parse_conv_layers(O.FITHUBERT_CONV)
    Lm = 249600
    T = conv_out_lengths([Lm], conv)[0]
    assert T == 779
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(400, Lm + 1, (200,), generator=g).tolist() + [Lm, Lm - 1, 400, 719, 720, 721]
    ref = O.conv_out_lengths(torch.tensor(lens), conv).tolist()
    assert conv_out_lengths(lens, conv) == ref
    for n in lens[:50] + lens[-6:]:
        pm = ~(torch.arange(Lm)[None] < torch.tensor([[Lm], [n]]))
        assert hubert_mask_lengths([Lm, n], Lm, T) == O.valid_lengths(O.mask_m3(pm, T), T, 2).tolist()


def test_lr_schedule_matches_oracle():
    for step in (0, 1, 49, 50, 51, 500, 999, 1000):
        assert warmup_linear(step, 1000, 0.05) == pytest.approx(O.lr_schedule(step, 1000, 0.05))
