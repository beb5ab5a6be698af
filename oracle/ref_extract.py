"""Run single functions of the UNMODIFIED reference whose modules cannot be imported here.

TEST INFRASTRUCTURE ONLY (tests/, oracle/gen_golden.py).

`train.py` needs pytorch_lightning + s3prl, `utils/utils.py` needs pytz + omegaconf + half of fairseq: neither imports in
this container.  The functions this path needs from them are self-contained, though - they only use `torch`, `F` and their
arguments - so they are cut out of the reference's source text with `ast` (the file is read where it lies, nothing is
copied into the repo) and compiled on their own:

  * `W2V2Distil.calculate_loss` (train.py:236-405): called with a plain namespace standing in for `self`;
  * `rtrn_attn_forward` (utils/utils.py:190-280): bound over the encoder layers exactly as train.py:64-77 does.
"""
import ast
import os
import types

import torch
import torch.nn.functional as F

REF = os.environ.get("FHB_REFERENCE", "/root/reference")


def load_ref_function(rel_path: str, name: str, cls: str = None):
    """The function `name` (a method of class `cls` if given) of REF/rel_path, compiled from the reference's own source."""
    path = os.path.join(REF, rel_path)
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    body = tree.body
    if cls is not None:
        body = next(n for n in body if isinstance(n, ast.ClassDef) and n.name == cls).body
    fn = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == name)
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"torch": torch, "F": F, "nn": torch.nn}
    exec(compile(mod, path, "exec"), ns)
    return ns[name]


def ref_calculate_loss(student_results, teacher_results, *, train_cfg: dict, model_cfg: dict, pred_layer_id,
                       rand_l=(), num_encoders: int = 0):
    """W2V2Distil.calculate_loss of the reference on the given result dicts.  `self` is a namespace with exactly the
    attributes the method reads (train.py:55-62,88-91)."""
    fn = load_ref_function("train.py", "calculate_loss", cls="W2V2Distil")
    self = types.SimpleNamespace(
        train_cfg=train_cfg, model_cfg=model_cfg,
        cnn_loss_weight=train_cfg["cnn_loss_weight"], rec_loss_weight=train_cfg["rec_loss_weight"],
        rec_loss_type=train_cfg["rec_loss_type"], sim_loss_weight=train_cfg["sim_loss_weight"],
        attn_loss_weight=train_cfg["attn_loss_weight"], attn_loss_type=train_cfg["attn_loss_type"],
        v_rel_loss_weight=train_cfg["v_rel_loss_weight"], random_layer_weight=train_cfg["random_layer_weight"],
        rand_l=list(rand_l), num_encoders=num_encoders, task_agnostic=True,
        student_model=types.SimpleNamespace(pred_layer_id=list(pred_layer_id)))
    return fn(self, student_results, teacher_results)


def bind_attn_forward(layers):
    """train.py:64-77: every encoder layer's forward is replaced by utils/utils.py `rtrn_attn_forward`."""
    fn = load_ref_function("utils/utils.py", "rtrn_attn_forward")
    for layer in layers:
        layer.self_attn._set_skip_embed_dim_check()
        setattr(layer, "forward", fn.__get__(layer, layer.__class__))
