"""CPU-only: real-checkpoint I/O.  The files the reference reads are pickles that reference
classes of fairseq / omegaconf / pytorch_lightning, none of which exist in this image; the tests build such
pickles from throw-away modules, remove the modules again and read the files back through checkpoint.py."""
import argparse
import os
import sys
import types

import pytest
import torch

import fhb_oracle as O
import fithubert_b200 as F
from fithubert_b200 import checkpoint as CK
from fithubert_b200.config import CustomStudentModelConfig

T_OVER = dict(conv_feature_layers="[(32,10,5)] + [(32,3,2)] * 4 + [(32,2,2)] * 2", encoder_layers=2,
              encoder_embed_dim=64, encoder_ffn_embed_dim=128, encoder_attention_heads=4, conv_pos=16, conv_pos_groups=4)


def _fake_modules():
    """fairseq.data.dictionary.Dictionary, omegaconf-like containers and a Lightning helper: pickled, then deleted."""
    mods = {}
    for name in ("fakeseq", "fakeseq.data", "fakeseq.data.dictionary", "fakeconf", "fake_lightning"):
        mods[name] = types.ModuleType(name)

    class Dictionary:
        def __init__(self):
            self.symbols, self.count = ["<s>", "a", "b"], [1, 2, 3]

    class DictNode:  # keeps children like omegaconf's DictConfig (`_content`) / ValueNode (`_val`)
        def __init__(self, content):
            self.__dict__["_content"] = content

        def __getstate__(self):
            return dict(self.__dict__)

        def __setstate__(self, st):
            self.__dict__.update(st)

    class ValueNode:
        def __init__(self, v):
            self._val = v

    class AttributeDict(dict):
        pass

    for cls, mod in ((Dictionary, "fakeseq.data.dictionary"), (DictNode, "fakeconf"), (ValueNode, "fakeconf"),
                     (AttributeDict, "fake_lightning")):
        cls.__module__ = mod
        cls.__qualname__ = cls.__name__
        setattr(mods[mod], cls.__name__, cls)
    sys.modules.update(mods)
    return mods


def _drop(mods):
    for name in mods:
        sys.modules.pop(name, None)


def _teacher_state(kind):
    tcfg = O.teacher_config(**T_OVER)
    sd = O.init_teacher_state(tcfg, 5, perturb=True)
    extra = {"mask_emb": torch.randn(64), "final_proj.weight": torch.randn(16, 64), "final_proj.bias": torch.zeros(16)}
    if kind == "hubert":
        extra["label_embs_concat"] = torch.randn(10, 16)
    else:
        extra["quantizer.vars"] = torch.randn(1, 8, 4)
        extra["project_q.weight"] = torch.randn(16, 16)
    return sd, dict(sd, **extra)


@pytest.mark.parametrize("style", ["cfg_dict", "cfg_omegaconf", "args_namespace", "bare"])
def test_fairseq_teacher_checkpoint_roundtrip(tmp_path, style):
    mods = _fake_modules()
    kind = "wav2vec2" if style == "args_namespace" else "hubert"
    sd, full = _teacher_state(kind)
    model_cfg = dict(_name=kind, conv_feature_layers=T_OVER["conv_feature_layers"], encoder_attention_heads=4,
                     layer_norm_first=False)
    state = {"model": full, "task_state": {"dictionaries": [mods["fakeseq.data.dictionary"].Dictionary()]},
             "extra_state": {"epoch": 3}, "optimizer_history": [], "args": None, "cfg": None}
    if style == "cfg_dict":
        state["cfg"] = {"model": model_cfg, "task": {"_name": "hubert_pretraining"}}
    elif style == "cfg_omegaconf":
        DN, VN = mods["fakeconf"].DictNode, mods["fakeconf"].ValueNode
        state["cfg"] = DN({"model": DN({k: VN(v) for k, v in model_cfg.items()})})
    elif style == "args_namespace":
        state["args"] = argparse.Namespace(arch="wav2vec2", conv_feature_layers=T_OVER["conv_feature_layers"],
                                           encoder_attention_heads=4, layer_norm_first=False)
    path = os.path.join(tmp_path, "teacher.pt")
    torch.save(state, path)
    _drop(mods)
    if style == "bare":
        # no cfg, no args and a head count that cannot be guessed from E = 64: must say so, not guess
        with pytest.raises(NotImplementedError):
            CK.load_fairseq_teacher(path)
        return
    with pytest.raises(Exception):  # plain torch.load cannot import the pickled classes
        torch.load(path, map_location="cpu", weights_only=False)
    model, got_kind, cfg = CK.load_fairseq_teacher(path)
    assert got_kind == kind and cfg["encoder_layers"] == 2 and cfg["conv_pos_groups"] == 4
    own = model.state_dict()
    assert set(own) == set(sd)
    for k, v in sd.items():
        assert torch.equal(own[k], v), k
    # the reference-facing entry point wraps it (reference utils/utils.py:102-149 return triple)
    wrapper, model_cfg2, task_agnostic = F.load_model_and_config(path)
    assert isinstance(wrapper, F.TeacherWrapper) and task_agnostic is True
    assert torch.equal(wrapper.model.state_dict()["encoder.layers.1.fc2.weight"], sd["encoder.layers.1.fc2.weight"])


def test_teacher_checkpoint_rejects_what_the_path_does_not_implement(tmp_path):
    sd, full = _teacher_state("hubert")
    path = os.path.join(tmp_path, "t.pt")
    torch.save({"model": full, "cfg": {"model": {"_name": "wav2vec_ctc"}}}, path)
    with pytest.raises(NotImplementedError):
        CK.load_fairseq_teacher(path)
    torch.save({"model": full, "cfg": {"model": {"_name": "data2vec_audio"}}}, path)
    with pytest.raises(NotImplementedError, match="is not supported"):  # the reference's own message (utils/utils.py:141)
        CK.load_fairseq_teacher(path)
    partial = {k: v for k, v in full.items() if k != "encoder.layers.1.fc1.bias"}
    torch.save({"model": partial, "cfg": {"model": {"_name": "hubert", "encoder_attention_heads": 4,
                                                    "conv_feature_layers": T_OVER["conv_feature_layers"]}}}, path)
    with pytest.raises(KeyError):
        CK.load_fairseq_teacher(path)
    torch.save({"weights": 1}, path)
    with pytest.raises(ValueError):
        CK.load_fairseq_teacher(path)


def test_lightning_student_checkpoint_into_expert(tmp_path):
    mods = _fake_modules()
    s_over = dict(conv_feature_layers="[(16, 10, 5)] + [(32, 1, 1)] + [(32, 3, 2)] * 4 + [(64, 1, 1)] + [(64, 2, 2)] * 2",
                  encoder_layers=2, encoder_embed_dim=96, encoder_ffn_embed_dim=96, encoder_attention_heads=4,
                  conv_pos=16, conv_pos_groups=4, pred_head_final_dim=64)
    ssd = O.init_student_state(O.student_config(**s_over), 2, perturb=True)
    ckpt = {"state_dict": dict({f"student_model.{k}": v for k, v in ssd.items()},
                               **{"teacher_model.model.mask_emb": torch.zeros(4)}),
            "hyper_parameters": mods["fake_lightning"].AttributeDict(lr=1e-3), "epoch": 7, "global_step": 1234}
    path = os.path.join(tmp_path, "last.ckpt")
    torch.save(ckpt, path)
    _drop(mods)
    got = CK.load_student_state_dict(path)
    assert set(got) == set(ssd) and all(torch.equal(got[k], v) for k, v in ssd.items())
    distiller = dict(extractor_mode="default", layerwise_proj=True, enable_tr_layer=True, tr_layer_index=0,
                     tr_layer_type="conv1d", required_seq_len_multiple=1, pred_layer_id="[1]", init_conv_layers=True,
                     init_encoder_layers=2, **s_over)
    expert = F.UpstreamExpert(path, {"distiller": distiller})  # same call as reference fithubert/expert.py:10
    assert expert.get_downsample_rates("any") == 320
    assert expert.model.proj_head is None and expert.model.final_proj is not None  # _disable_projection_heads ran
    assert torch.equal(expert.model.state_dict()["final_proj.lin_proj.weight"], ssd["proj_head.1.lin_proj.weight"])
    assert torch.equal(expert.model.state_dict()["encoder.layers.0.weight"], ssd["encoder.layers.0.weight"])
    # what we write is what the reference's expert reads
    student = F.CustomStudentModel(CustomStudentModelConfig(**dict(distiller, init_conv_layers=False, init_encoder_layers=0)))
    student.load_state_dict(ssd)
    out = CK.student_checkpoint(student, {"epoch": 1})
    assert all(k.startswith("student_model.") for k in out["state_dict"])
    assert set(k[14:] for k in out["state_dict"]) == set(ssd)


def test_init_student_from_teacher_layers():
    """init_from_teacher_conv / init_from_teacher_enc (reference modules/model.py:560-588)."""
    same = dict(conv_feature_layers=T_OVER["conv_feature_layers"], encoder_layers=2, encoder_embed_dim=64,
                encoder_ffn_embed_dim=128, encoder_attention_heads=4, conv_pos=16, conv_pos_groups=4, pred_head_final_dim=64)
    teacher = F.TeacherModel(kind="hubert", **T_OVER)
    teacher.load_state_dict(O.init_teacher_state(O.teacher_config(**T_OVER), 5, perturb=True))
    wrapper = F.TeacherWrapper(teacher)
    cfg = CustomStudentModelConfig(extractor_mode="default", layerwise_proj=True, enable_tr_layer=True, tr_layer_index=0,
                                   tr_layer_type="conv1d", required_seq_len_multiple=1, pred_layer_id="[1]",
                                   init_conv_layers=True, init_encoder_layers=0, **same)
    student = F.CustomStudentModel(cfg, teacher_model=wrapper)
    tsd, ssd = teacher.state_dict(), student.state_dict()
    for k in tsd:
        if k.startswith("feature_extractor.") or k.startswith("post_extract_proj."):
            assert torch.equal(ssd[k], tsd[k]), k
