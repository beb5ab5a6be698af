"""Real-checkpoint I/O without fairseq / Lightning installed.

The reference reads two kinds of files:
  * a fairseq checkpoint of the teacher (`hubert_base_ls960.pt`, `wav2vec_small.pt`), through
    fairseq's `load_checkpoint_to_cpu` + `build_model` + `load_state_dict(state['model'])`
    (reference utils/utils.py:102-149);
  * a Lightning checkpoint of the distilled student, whose `state_dict` carries the student under the
    `student_model.` prefix (reference fithubert/expert.py:40-43).
Both are `torch.save` pickles that reference classes of packages this image does not have
(`fairseq.data.dictionary.Dictionary`, `omegaconf.dictconfig.DictConfig`, `pytorch_lightning...`).
Only the tensors and a handful of plain config values are needed, so the files are read with an unpickler
that substitutes an inert placeholder for every class it cannot import.  The parameter names of
`TeacherModel` / `CustomStudentModel` are the reference's own, so loading is a key filter
plus a strict completeness check - no renaming.
"""
from __future__ import annotations

import argparse
import pickle
import re
import types
from collections import OrderedDict
from typing import Any, Dict, List, Optional, Tuple

import torch


# --------------------------------------------------------------------------- tolerant unpickling
class _Placeholder:
    """Stands in for an instance of a class that is not importable here; keeps whatever state pickle hands it."""

    def __init__(self, *args, **kwargs):
        self._args, self._kwargs = args, kwargs

    def __setstate__(self, state):
        self._state = state

    def __call__(self, *args, **kwargs):  # some reducers call the reconstructed object
        return self

    # dict / list / set subclasses are rebuilt through SETITEM(S) / APPEND(S) / ADDITEMS opcodes
    def __setitem__(self, key, value):
        self.__dict__.setdefault("_items", {})[key] = value

    def append(self, value):
        self.__dict__.setdefault("_list", []).append(value)

    def extend(self, values):
        self.__dict__.setdefault("_list", []).extend(values)

    def add(self, value):
        self.__dict__.setdefault("_list", []).append(value)

    def state(self) -> Any:
        return getattr(self, "_state", None) or self.__dict__


_PLACEHOLDER_CLASSES: Dict[Tuple[str, str], type] = {}


def _placeholder_class(module: str, name: str) -> type:
    key = (module, name)
    if key not in _PLACEHOLDER_CLASSES:
        _PLACEHOLDER_CLASSES[key] = type(name, (_Placeholder,), {"__module__": f"_missing.{module}"})
    return _PLACEHOLDER_CLASSES[key]


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            return _placeholder_class(module, name)


def _tolerant_pickle_module() -> types.ModuleType:
    m = types.ModuleType("fhb_tolerant_pickle")
    m.Unpickler = _TolerantUnpickler
    m.load = lambda f, **kw: _TolerantUnpickler(f, **kw).load()
    m.loads = pickle.loads
    m.dump, m.dumps, m.Pickler = pickle.dump, pickle.dumps, pickle.Pickler
    m.UnpicklingError, m.PicklingError = pickle.UnpicklingError, pickle.PicklingError
    m.HIGHEST_PROTOCOL, m.DEFAULT_PROTOCOL = pickle.HIGHEST_PROTOCOL, pickle.DEFAULT_PROTOCOL
    return m


def load_checkpoint_to_cpu(path: str) -> Dict[str, Any]:
    """`torch.load(path, map_location='cpu')` that survives pickled objects of absent packages."""
    return torch.load(path, map_location="cpu", pickle_module=_tolerant_pickle_module(), weights_only=False)


# --------------------------------------------------------------------------- config digging
def _get(node: Any, key: str, default=None):
    """Fetch `key` from a dict / Namespace / placeholder-of-omegaconf node."""
    if node is None:
        return default
    if isinstance(node, dict):
        return node.get(key, default)
    if isinstance(node, argparse.Namespace):
        return getattr(node, key, default)
    if isinstance(node, _Placeholder):
        st = node.state()
        if isinstance(st, dict):
            content = st.get("_content", st)  # omegaconf containers keep their children in `_content`
            if isinstance(content, dict) and key in content:
                v = content[key]
                if isinstance(v, _Placeholder):  # omegaconf ValueNode: the value sits in `_val`
                    vs = v.state()
                    if isinstance(vs, dict) and "_val" in vs:
                        return vs["_val"]
                return v
        return default
    return getattr(node, key, default)


def teacher_kind_and_config(state: Dict[str, Any]) -> Tuple[str, Dict[str, Any]]:
    """Model type (`hubert` / `wav2vec2`) and the architecture fields TeacherModel needs, from `state['cfg']`
    (new-style checkpoints), `state['args']` (old argparse checkpoints) or, failing both, the tensor shapes."""
    sd = state["model"]
    model_cfg = _get(state.get("cfg"), "model") if state.get("cfg") is not None else None
    args = state.get("args")
    name = _get(model_cfg, "_name") or _get(args, "arch")
    if name is None:  # infer from what only one of the two families carries
        name = "hubert" if "label_embs_concat" in sd else ("wav2vec2" if any(k.startswith("quantizer.") or k == "project_q.weight" for k in sd) else None)
    name = str(name) if name is not None else None
    if name in ("wav2vec_ctc", "hubert_ctc"):
        raise NotImplementedError(f"model '{name}' (task-specific CTC teacher) is outside the B200 hot path (SURVEY 2.1)")
    if name is None or not (name.startswith("hubert") or name.startswith("wav2vec2")):
        raise NotImplementedError(f"model '{name}' is not supported.")
    kind = "hubert" if name.startswith("hubert") else "wav2vec2"
    src = model_cfg if model_cfg is not None else args
    if any(re.match(r"feature_extractor\.conv_layers\.\d+\.2\.1\.weight", k) for k in sd):
        raise NotImplementedError("extractor_mode='layer_norm' teachers (Large models) are not implemented")
    if "encoder.layers.0.self_attn.q_proj.weight" not in sd:
        raise NotImplementedError("unexpected encoder parameter names (conformer / pre-LN teachers are not implemented)")
    if _get(src, "layer_norm_first", False):
        raise NotImplementedError("layer_norm_first teachers are not implemented")
    n_conv = 1 + max(int(m.group(1)) for m in (re.match(r"feature_extractor\.conv_layers\.(\d+)\.0\.weight", k) for k in sd) if m)
    spec = _get(src, "conv_feature_layers")
    if spec is None:  # shapes give (C, k); strides are wav2vec 2.0's fixed 5,2,2,... (total 320)
        layers = []
        for i in range(n_conv):
            w = sd[f"feature_extractor.conv_layers.{i}.0.weight"]
            layers.append((w.shape[0], w.shape[2], 5 if i == 0 else 2))
        spec = layers
    E = sd["post_extract_proj.weight"].shape[0]
    n_layers = 1 + max(int(m.group(1)) for m in (re.match(r"encoder\.layers\.(\d+)\.", k) for k in sd) if m)
    v = sd["encoder.pos_conv.0.weight_v"]
    heads = _get(src, "encoder_attention_heads")
    if heads is None:
        heads = 12 if E == 768 else (16 if E == 1024 else None)
    if heads is None:
        raise NotImplementedError("cannot infer encoder_attention_heads: the checkpoint carries neither cfg nor args")
    cfg = dict(conv_feature_layers=spec, encoder_embed_dim=E, encoder_ffn_embed_dim=sd["encoder.layers.0.fc1.weight"].shape[0],
               encoder_attention_heads=int(heads), encoder_layers=n_layers, conv_pos=v.shape[2],
               conv_pos_groups=E // v.shape[1])
    return kind, cfg


def load_fairseq_teacher(path: str):
    """Reference utils/utils.py:102-149 without fairseq: returns (TeacherModel with the checkpoint's weights, kind,
    architecture dict).  `state['model']` keys outside the features_only trunk (mask_emb, label_embs_concat,
    final_proj, quantizer, project_q, ...) are ignored exactly as `extract_features(mask=None)` never touches them."""
    from .model import TeacherModel
    state = load_checkpoint_to_cpu(path)
    if not isinstance(state, dict) or "model" not in state:
        raise ValueError(f"{path}: not a fairseq checkpoint (no 'model' entry)")
    kind, cfg = teacher_kind_and_config(state)
    model = TeacherModel(kind=kind, **cfg)
    own = model.state_dict()
    missing = [k for k in own if k not in state["model"]]
    if missing:
        raise KeyError(f"{path}: checkpoint lacks {len(missing)} teacher tensors, e.g. {missing[:4]}")
    model.load_state_dict(OrderedDict((k, state["model"][k].float()) for k in own), strict=True)
    return model, kind, cfg


def load_student_state_dict(ckpt) -> "OrderedDict[str, torch.Tensor]":
    """Lightning checkpoint -> student state dict (reference fithubert/expert.py:40-43: keys under `student_model.`
    with the prefix cut).  A plain state dict (no 'state_dict' entry) is returned as is."""
    state = ckpt if isinstance(ckpt, dict) else load_checkpoint_to_cpu(ckpt)
    sd = state.get("state_dict", state)
    if any("student_model" in k for k in sd):
        return OrderedDict((k[14:], v) for k, v in sd.items() if "student_model" in k)
    return OrderedDict(sd)


def student_checkpoint(student, extra: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
    """The Lightning-shaped dict the reference's UpstreamExpert expects ({'state_dict': {'student_model.*'}})."""
    out = {"state_dict": OrderedDict((f"student_model.{k}", v.detach().cpu()) for k, v in student.state_dict().items())}
    if extra:
        out.update(extra)
    return out


def tensor_names(sd: Dict[str, torch.Tensor]) -> List[str]:
    return sorted(sd)
