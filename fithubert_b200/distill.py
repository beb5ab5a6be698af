"""W2V2Distil: the reference's distillation module (train.py:26-446) without Lightning.

Same surface for the hot path: forward(x, padding_mask) -> (student_results, teacher_results),
calculate_loss(student_results, teacher_results, labels=None) -> (total_loss, losses),
training_step(batch, batch_idx) -> loss, configure_optimizers().  training_step is the FUSED path:
teacher forward -> student forward -> loss+gradient kernel -> explicit backward, no autograd.
The CNN-feature L1 loss (train.py:241-246) is built; CTC / attention-map / value-relation losses are out of scope
(SURVEY 2.1): requesting them raises NotImplementedError.
"""
from __future__ import annotations

import os
import random
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import engine as E
from . import kernels as K
from . import lib as L
from .config import CustomStudentModelConfig
from .model import CustomStudentModel, TeacherModel, TeacherWrapper, conv_out_lengths, freeze_model, _lengths_from_mask
from .optim import FusedAdamW, GradAllReduce

f16 = torch.float16


def load_model_and_config(teacher_model: str, device=None):
    """Reference utils/utils.py:102-149: (TeacherWrapper, model_cfg, task_agnostic).  An existing file is read as a
    fairseq checkpoint WITHOUT fairseq (checkpoint.load_fairseq_teacher: tolerant unpickler + the reference's own
    parameter names) and must supply every teacher tensor.  A name that is not a file selects the architecture
    (HuBERT-Base / wav2vec 2.0-Base) with random-init weights - the BASELINE.json benchmark configuration, where
    no checkpoint can be downloaded."""
    import os
    if isinstance(teacher_model, str) and os.path.isfile(teacher_model):
        from .checkpoint import load_fairseq_teacher
        model, _, model_cfg = load_fairseq_teacher(teacher_model)
    else:
        kind = "wav2vec2" if ("wav2vec" in str(teacher_model) or "w2v" in str(teacher_model)) else "hubert"
        model, model_cfg = TeacherModel(kind=kind), None
    if device is not None:
        model = model.to(device)
    return TeacherWrapper(model), model_cfg, True


class W2V2Distil(nn.Module):
    def __init__(self, cfg: dict, teacher_model: Optional[TeacherWrapper] = None, device=None):
        super().__init__()
        self.yaml_cfg = cfg
        self.train_cfg = cfg["train"]
        dev = torch.device(device if device is not None else "cuda")
        if teacher_model is None:
            teacher_model, _, self.task_agnostic = load_model_and_config(cfg["teacher"]["teacher_model"], dev)
        else:
            self.task_agnostic = True
        self.teacher_model = teacher_model.to(dev)
        freeze_model(self.teacher_model)
        self.model_cfg = cfg["distiller"]
        student_config = CustomStudentModelConfig(**self.model_cfg)
        student_config._teacher_task_agnostic = self.task_agnostic
        student_config._cnn_weight = self.train_cfg["cnn_loss_weight"]
        self.student_model = CustomStudentModel(cfg=student_config, teacher_model=self.teacher_model).to(dev)
        t = self.train_cfg
        self.cnn_loss_weight, self.rec_loss_weight = t["cnn_loss_weight"], t["rec_loss_weight"]
        self.rec_loss_type, self.sim_loss_weight = t["rec_loss_type"], t["sim_loss_weight"]
        self.attn_loss_weight, self.v_rel_loss_weight = t["attn_loss_weight"], t["v_rel_loss_weight"]
        self.random_layer_weight = t["random_layer_weight"]
        self.attn_loss_type = t.get("attn_loss_type", "kldiv")
        if self.attn_loss_weight > 0:
            # Attention-map distillation (train.py:64-77): the reference re-binds the forward of EVERY entry of
            # encoder.layers to utils/utils.py `rtrn_attn_forward` after touching its `.self_attn`; with a time-reduction
            # layer entry 0 is an nn.Conv1d (modules/module.py:230-236) and construction fails - same error here
            if self.student_model.enable_tr_layer:
                raise AttributeError("'Conv1d' object has no attribute 'self_attn'")
            if self.attn_loss_type not in ("mse", "kldiv"):
                raise NotImplementedError("attn_loss_type must be one of 'mse', 'kldiv'.")
            self.student_model._return_attn = True
            self.teacher_model.model._return_attn = True
        elif self.v_rel_loss_weight > 0:
            # train.py:357-358 subscripts layer_results[-1][1], which is None unless the attention recipe is bound
            raise TypeError("'NoneType' object is not subscriptable")
        if self.sim_loss_weight and t["distil_random_layer"] > 0:
            # train.py:304-306 reduces the 3-D cosine loss over dim 3 in this branch and cannot execute
            raise NotImplementedError("sim_loss_weight > 0 needs distil_random_layer = 0 (the reference's random-layer "
                                      "branch of the cosine loss, train.py:304-306, indexes a dimension that does not exist)")
        if self.sim_loss_weight and not self.rec_loss_weight:
            raise NotImplementedError("sim_loss_weight > 0 with rec_loss_weight = 0 reads `pred` before assignment "
                                      "in the reference (train.py:249,303)")
        if self.rec_loss_type not in ("mse", "l1"):
            raise NotImplementedError("rec_loss_type must be one of 'l1', 'mse'.")
        if t.get("delete_projections"):
            raise NotImplementedError("delete_projections=True leaves nothing to distil on this path")
        self.num_encoders = self.model_cfg["encoder_layers"]
        n = self.num_encoders
        self.split_head = not self.student_model.layerwise_proj
        self.tgt_slots = None  # teacher layer -> row of the stacked target buffer (split head only)
        if self.split_head:
            # ex.yaml recipe (train.py:268-281 with layerwise_proj False): task k of the SplitLinear head is matched with
            # teacher layer pred_layer_id[k]; plain means over all tasks (train.py:294-297)
            if t["distil_random_layer"] > 0:
                raise NotImplementedError("distil_random_layer > 0 indexes a list of per-layer projections; the "
                                          "SplitLinear head returns one tensor (train.py:258-267)")
            ids = self.student_model.pred_layer_id
            n_teacher = len(self.teacher_model.model.encoder.layers)
            assert max(ids) < n_teacher, "pred_layer_id exceeds the teacher depth"
            self.tgt_slots = [ids.index(l) if l in ids else None for l in range(n_teacher)]
            n = len(ids)
            w = [1.0 / n] * n
            self.rand_l = []
            self.mean_over_layers = True
        elif t["distil_random_layer"] > 0:
            # train.py:88-91: a random subset of the lower layers, each weighted random_layer_weight
            self.all_enc = range(n - 1)
            self.rand_l = random.sample(self.all_enc, t["distil_random_layer"])
            w = [0.0] * n
            for l in self.rand_l:
                w[l] = float(self.random_layer_weight)
            w[n - 1] = 1.0
            self.mean_over_layers = False
        else:
            assert t["random_layer_weight"] == 0
            # train.py:294-297: plain mean over the pred_layer_id layers
            ids = self.student_model.pred_layer_id
            w = [0.0] * n
            for l in ids:
                w[l] = 1.0 / len(ids)
            self.rand_l = []
            self.mean_over_layers = True
        self.n_pred = n  # rows of the stacked prediction / target buffers
        self.layer_weights_host = w
        self.layer_weights = torch.tensor(w, dtype=torch.float32, device=dev)
        self.batch_size = t["batch_size"]
        self.num_gpus = t["gpus"] if not isinstance(t["gpus"], list) else len(t["gpus"])
        self.accumulate = int(t.get("accumulate_grad_batches", 1))
        self.optimizer: Optional[FusedAdamW] = None
        self.reducer: Optional[GradAllReduce] = None
        self._micro = 0
        self._tgt_buf = None
        self._loss_scale: Optional[float] = None

    # ------------------------------------------------------------------ reference-style API (autograd)
    def forward(self, x, padding_mask=None):
        self.teacher_model.eval()
        teacher_results = self.teacher_model.extract_features(source=x, padding_mask=padding_mask)
        student_results = self.student_model(source=x, padding_mask=padding_mask)
        return student_results, teacher_results

    def calculate_loss(self, student_results, teacher_results, labels=None):
        from .autograd import _DistillLossFn
        if labels is not None or not self.task_agnostic:
            raise NotImplementedError("task-specific (CTC) teacher path is dead code in the reference and not built")
        tgt = teacher_results["_stacked"]
        if self.split_head:
            # projections is ONE [B, N, T, D] tensor (a view of the engine's [N, B, T, D] buffer); targets are the
            # teacher layers pred_layer_id (train.py:268-281)
            preds = student_results["projections"].permute(1, 0, 2, 3).contiguous()
            tgt = tgt[self.student_model.pred_layer_id]
        else:
            preds = torch.stack(student_results["projections"], 0) if not isinstance(
                student_results["projections"], torch.Tensor) else student_results["projections"]
            base = student_results["projections"][0]
            if isinstance(student_results["projections"], list) and base._base is not None and \
                    base._base.shape[0] == len(student_results["projections"]):
                preds = base._base  # the engine's stacked [n, B, T', D] buffer: no copy
        lt = 0 if self.rec_loss_type == "mse" else 1
        S = self.loss_scale(preds.shape[1] * preds.shape[2] * preds.shape[3])
        if self.sim_loss_weight:
            total, rec, sim = _DistillLossFn.apply(preds, tgt, self.layer_weights, lt,
                                                   float(self.rec_loss_weight), float(self.sim_loss_weight), S)
            per_layer = rec + sim  # train.py:316 feat_loss = rec_layer_loss + sim_layer_loss (un-weighted sum)
        else:
            total, per_layer = _DistillLossFn.apply(preds, tgt, self.layer_weights, lt, float(self.rec_loss_weight), 0.0, S)
        losses = self._loss_dict(per_layer)
        if self.cnn_loss_weight > 0:
            # CNN post projection loss (train.py:241-246): plain L1 mean between `features` and the teacher's
            from .autograd import _FeatureL1Fn
            cnn_loss = _FeatureL1Fn.apply(student_results["features"], teacher_results["features"][0], S)
            losses = {"cnn_loss": cnn_loss, **losses}
            total = total + self.cnn_loss_weight * cnn_loss
        if self.attn_loss_weight > 0:
            # attention distribution transfer / value relation transfer (train.py:327-378) on the last layer's maps
            from .autograd import _AttnMapLossFn
            pred, v_pred = student_results["layer_results"][-1][1]
            target, v_target = teacher_results["layer_results"][-1][1][0]
            if pred.shape != target.shape:
                raise RuntimeError(f"The size of tensor a {tuple(pred.shape)} must match the size of tensor b "
                                   f"{tuple(target.shape)}")
            BH, T = pred.shape[:2]
            H = self.student_model._geom.H
            B = BH // H
            vs, vt = student_results.get("_valid"), teacher_results.get("_valid")
            dev = pred.device
            vs_d, vt_d = E._valid_tensor(vs, dev), E._valid_tensor(vt, dev)
            if self.attn_loss_type == "mse":
                count = H * T * sum(min(a, b, T) for a, b in zip(vs or [T] * B, vt or [T] * B))
                attn_loss = _AttnMapLossFn.apply(pred, target, vs_d, vt_d, H, 0, 1.0 / count, S)
            else:
                attn_loss = _AttnMapLossFn.apply(pred, target, vs_d, vt_d, H, 1, 1.0 / (BH * T), S)
            losses["attn_loss"] = attn_loss
            total = total + self.attn_loss_weight * attn_loss
            if self.v_rel_loss_weight > 0:
                v_rel_loss = _AttnMapLossFn.apply(v_pred, v_target, None, None, H, 1, 1.0 / (BH * T), S)
                losses["v_rel_loss"] = v_rel_loss
                total = total + self.v_rel_loss_weight * v_rel_loss
        return total, losses

    def _loss_dict(self, per_layer: torch.Tensor) -> Dict[str, torch.Tensor]:
        losses = {}
        n = self.num_encoders
        if self.train_cfg["distil_random_layer"] > 0:
            for i, l in enumerate(self.rand_l):
                losses[f"rand_l{i}"] = per_layer[l]
            losses[f"l{n - 1}"] = per_layer[n - 1]
        else:
            ids = self.student_model.pred_layer_id
            for k, pred_id in enumerate(ids):  # split head: row k of the stacked buffers; layer-wise heads: row pred_id
                losses[f"layer{pred_id}"] = per_layer[k if self.split_head else pred_id] * len(ids)
        return losses

    # ------------------------------------------------------------------ fused training path
    def loss_scale(self, n_elems: int) -> float:
        """The loss scale of this module (train_cfg['loss_scale'] if given, else engine.loss_scale_for of the first
        training batch, the minimum over the ranks): constant afterwards, so that accumulated micro-batches and the
        ranks of a data-parallel job all add gradients of ONE scale into the flat buffer."""
        if self._loss_scale is None:
            S = float(self.train_cfg.get("loss_scale", 0) or E.loss_scale_for(n_elems))
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                t = torch.tensor([S], dtype=torch.float64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                S = float(t[0])
            self._loss_scale = S
            self.student_model._loss_scale = S
        return self._loss_scale

    def configure_optimizers(self, total_steps: int = 0):
        o = self.yaml_cfg["optimizer"]
        lr = float(o["lr"]) if not isinstance(o["lr"], str) else float(eval(o["lr"], {"__builtins__": {}}))
        self.optimizer = FusedAdamW(self.student_model, lr=lr, betas=tuple(o.get("betas", (0.9, 0.999))),
                                    eps=float(o.get("eps", 1e-8)), weight_decay=float(o.get("weight_decay", 0.0)),
                                    total_steps=total_steps, warmup_proportion=float(o.get("warmup_proportion", 0.05)))
        self.reducer = GradAllReduce()
        return {"optimizer": self.optimizer}

    def fused_forward_backward(self, x, padding_mask=None, lengths: Optional[List[int]] = None, grad_scale=1.0):
        """teacher fwd + student fwd + loss + student bwd for one micro-batch.  Gradients accumulate in the
        flat buffer.  Returns the per-layer loss contributions [n] fp32 on the device, already weighted
        (rec_loss_weight * rec + sim_loss_weight * sim per layer): their sum is the step's total loss."""
        sm, tm = self.student_model, self.teacher_model.model
        dev = sm.post_extract_proj.weight.device
        chunks = x_done = None
        if not x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous() and x.shape[0] >= 4:
            # host batch (what a DataLoader hands over): copied on a copy stream into a double buffer - under the previous
            # step when the GPU is still busy with it, else in batch slices that the conv stacks consume as they land
            x, chunks, x_done = E.h2d_chunked(x, dev)
        else:
            x = x.to(dev, non_blocking=True).float().contiguous()
        Ld = x.shape[1]
        T = E.conv_frames(Ld, tm._conv_layers)[-1]
        memo = {}

        def host_lengths():
            # resolved lazily, AFTER the first conv stack has been queued: scanning a host padding mask takes ~0.7 ms
            # for 32 x 250k samples, and nothing before the positional conv needs the result
            if "l" not in memo:
                ls = lengths
                if ls is None:
                    ls = _lengths_from_mask(padding_mask)
                elif all(n == Ld for n in ls):
                    ls = None
                memo["l"] = ls
            return memo["l"]

        def t_valid():
            # teacher mask (M3 for HuBERT: applied whenever a mask exists; M1 for wav2vec2)
            ls = host_lengths()
            if padding_mask is None and ls is None:
                return None
            if tm.kind == "hubert":
                from .model import hubert_mask_lengths
                return hubert_mask_lengths(ls if ls is not None else [Ld] * x.shape[0], Ld, T)
            return None if ls is None else conv_out_lengths(ls, tm._conv_layers)

        def s_valid():
            ls = host_lengths()
            return None if ls is None else conv_out_lengths(ls, sm._conv_layers)

        n, B, D = self.n_pred, x.shape[0], sm._geom.d_out
        Pt, Wt = tm.engine_state()
        t_extras = {} if self.attn_loss_weight > 0 else None  # the teacher's last-layer q | k | v (attention-map recipe)
        if self._tgt_buf is None or self._tgt_buf.shape[1:3] != (B, T):
            self._tgt_buf = torch.empty(n, B, T, tm._geom.E, device=dev, dtype=f16)
        P, W, G = sm.engine_state(True)
        if E.stream_mode() & 1:
            # the frozen teacher's forward and the student's forward only meet in the loss: run them on two streams so
            # that each one's launch tails and HBM-bound kernels fill under the other's GEMMs
            main, side = torch.cuda.current_stream(), E.side_stream(dev)
            side.wait_stream(main)  # the H2D copy of x, the previous step's readers of the target buffer
            with torch.cuda.stream(side):
                tgt, t_feats = E.teacher_forward(Pt, Wt, tm._geom, x, t_valid, out_buf=self._tgt_buf, slots=self.tgt_slots,
                                                 wave_chunks=chunks, extras=t_extras)
            c = E.student_forward(P, W, sm._geom, x, s_valid, train=True, heads="all", drop=sm.drop_cfg(),
                                  wave_chunks=chunks)
            main.wait_stream(side)
        else:
            tgt, t_feats = E.teacher_forward(Pt, Wt, tm._geom, x, t_valid, out_buf=self._tgt_buf, slots=self.tgt_slots,
                                             wave_chunks=chunks, extras=t_extras)
            c = E.student_forward(P, W, sm._geom, x, s_valid, train=True, heads="all", drop=sm.drop_cfg(),
                                  wave_chunks=chunks)
        layer_loss = torch.zeros(n, device=dev, dtype=torch.float32)
        # gradient written in place over the projections (they are not needed again).  fp16 gradients carry the loss
        # scale: fixed at the first training batch (and agreed on by all ranks), removed by the AdamW kernel
        dpred = c.preds
        S = self.loss_scale(B * c.Tq * D)
        grad_scale = grad_scale * S
        # the loss kernel also produces the column sums of the gradient it writes: both head bias gradients follow
        # (batched heads only; the per-head fallback path computes its own column sums)
        fused = getattr(c, "heads_batched", False) and G.head_stride() is not None
        dcs = torch.zeros(n, D, device=dev, dtype=torch.float32) if fused else None
        if self.split_head:  # column sums of dpred = SplitLinear bias gradient [1, 1, N, D]: accumulate it in place
            fused, dcs = True, G.view("proj_head.2.bias")
        lt = 0 if self.rec_loss_type == "mse" else 1
        if self.sim_loss_weight:
            sim_loss = torch.zeros(n, device=dev, dtype=torch.float32)
            K.distill_loss_sim(c.preds, tgt, self.layer_weights, layer_loss, sim_loss, dpred, n, B, c.Tq, T, D, lt,
                               grad_scale * self.rec_loss_weight, grad_scale * self.sim_loss_weight,
                               dbias=dcs, dbias_layer_stride=D if fused else 0)
            layer_loss = layer_loss * self.rec_loss_weight + sim_loss * self.sim_loss_weight
        else:
            K.distill_loss(c.preds, tgt, self.layer_weights, layer_loss, dpred, n, B, c.Tq, T, D, lt,
                           grad_scale * self.rec_loss_weight, dbias=dcs, dbias_layer_stride=D if fused else 0)
            if self.rec_loss_weight != 1.0:
                layer_loss = layer_loss * self.rec_loss_weight
        # data-parallel: the gradient all-reduce runs UNDER the backward by default (Lightning DDP's overlap,
        # train.py:494).  Buckets are issued on a side stream as the backward finalises them - heads, then every second
        # transformer layer, the front end last (optimizer_step) - and the persistent kernels leave a few SMs to NCCL's
        # CTAs (fhb_set_reserved_sms) so that neither side waits for the other's SMs.  FHB_EARLY_REDUCE=0: one exchange
        # after the backward.
        hook = None
        overlap = self.reducer is not None and self.reducer.enabled and self._micro == self.accumulate - 1 and \
            os.environ.get("FHB_EARLY_REDUCE", "1") == "1" and not self.split_head
        prev_reserved = 0
        if overlap:
            self.reducer._full = G.flat.numel()
            comm_sms = int(os.environ.get("FHB_COMM_SMS", "8"))
            windowed = os.environ.get("FHB_COMM_WINDOW", "1") == "1"
            prev_reserved = L.lib().fhb_set_reserved_sms(0 if windowed else comm_sms)

            def hook(off):
                # off: flat[off:] is final -> send it, and leave NCCL its SMs for the next layer's worth of kernels (an
                # exchange of two layers' gradients is over well within that); None: one layer later -> all SMs back
                if off is not None:
                    self.reducer.reduce_tail(G.flat, off)
                    L.lib().fhb_set_reserved_sms(comm_sms)
                elif windowed:
                    L.lib().fhb_set_reserved_sms(0)
        dfeatures = None
        self.last_cnn_loss = None
        if self.cnn_loss_weight > 0:
            # CNN-feature loss (train.py:241-246,372-378): L1 mean between features_to_distill and the teacher's
            # post_extract_proj output; the same loss kernel with one "layer", gradient written over the features
            if c.cnn_out is None:
                raise NotImplementedError("cnn_loss_weight > 0 without a cnn_proj_head (pred_head_final_dim == "
                                          "encoder_embed_dim) is only available through calculate_loss / autograd")
            Dc = c.cnn_out.shape[-1]
            if Dc != t_feats.shape[-1]:
                raise ValueError("CNN-feature loss: student features and teacher features differ in width")
            cnn = torch.zeros(1, device=dev, dtype=torch.float32)
            one = torch.ones(1, device=dev, dtype=torch.float32)
            K.distill_loss(c.cnn_out.view(1, B, T, Dc), t_feats.reshape(1, B, T, Dc), one, cnn, c.cnn_out.view(1, B, T, Dc),
                           1, B, T, T, Dc, 1, grad_scale * self.cnn_loss_weight)
            dfeatures = c.cnn_out
            self.last_cnn_loss = cnn[0]
            layer_loss = torch.cat([layer_loss, cnn * self.cnn_loss_weight])
        self.last_attn_losses = None
        if self.attn_loss_weight > 0:
            # attention distribution transfer + value relation transfer on the LAST layer's maps (train.py:327-378); the
            # gradients wrt the student's maps wait in c.attn_grad for the last layer's attention backward
            al = E.attn_transfer_losses(c, sm._geom, t_extras, tm._geom, loss_type=self.attn_loss_type,
                                        w_attn=float(self.attn_loss_weight), w_vrel=float(self.v_rel_loss_weight),
                                        grad_scale=grad_scale, valid_s=c.valid, valid_t=t_extras["valid_host"])
            self.last_attn_losses = al
            wv = torch.tensor([float(self.attn_loss_weight), float(self.v_rel_loss_weight)], device=dev)
            layer_loss = torch.cat([layer_loss, al * wv])
        try:
            E.student_backward(P, W, sm._geom, G, c, dpred, dpred_colsum=dcs, on_progress=hook, dfeatures=dfeatures)
        finally:
            if overlap:
                L.lib().fhb_set_reserved_sms(prev_reserved)
        G.loss_scale = S
        if x_done is not None:
            x_done.record()  # last reader of the waveform buffer (conv layer 0's backward) is queued
        return layer_loss

    def training_step(self, batch, batch_idx=0):
        """One micro-batch (reference train.py:158-170); every `accumulate_grad_batches`-th call also runs
        the gradient all-reduce and the fused AdamW step (what Lightning's automatic optimisation does)."""
        if self.optimizer is None:
            self.configure_optimizers()
        if self._micro == 0:
            self.optimizer.zero_grad()
        layer_loss = self.fused_forward_backward(batch["x"], batch.get("padding_mask"), batch.get("lengths"),
                                                 grad_scale=1.0 / self.accumulate)
        self._micro += 1
        if self._micro == self.accumulate:
            self._micro = 0
            self.optimizer_step()
        # trailing slots, if present: cnn_loss_weight * cnn_loss, then attn_loss_weight * attn_loss, v_rel_loss_weight * v_rel_loss
        self.last_layer_losses = layer_loss[:self.n_pred]
        return layer_loss.sum()

    def training_epoch_end(self, training_step_outputs=None):
        """train.py:172-178: a fresh random subset of the lower layers every epoch (distil_random_layer > 0)."""
        t = self.train_cfg
        if t["distil_random_layer"] > 0 and not self.split_head:
            n = self.num_encoders
            self.rand_l = random.sample(self.all_enc, t["distil_random_layer"])
            w = [0.0] * n
            for l in self.rand_l:
                w[l] = float(self.random_layer_weight)
            w[n - 1] = 1.0
            self.layer_weights_host = w
            self.layer_weights.copy_(torch.tensor(w, dtype=torch.float32))

    @torch.no_grad()
    def validation_step(self, batch, batch_idx=0):
        """train.py:180-203: teacher + student forward and the loss, no gradient; with random-layer distillation the
        monitored value is the last layer's term only (train.py:198-199)."""
        sm, tm = self.student_model, self.teacher_model.model
        dev = sm.post_extract_proj.weight.device
        x = batch["x"].to(dev, non_blocking=True).float().contiguous()
        lengths = batch.get("lengths")
        if lengths is None:
            lengths = _lengths_from_mask(batch.get("padding_mask"))
        Ld = x.shape[1]
        T = E.conv_frames(Ld, tm._conv_layers)[-1]
        if lengths is None and batch.get("padding_mask") is None:
            t_valid = None
        elif tm.kind == "hubert":
            from .model import hubert_mask_lengths
            t_valid = hubert_mask_lengths(lengths if lengths is not None else [Ld] * x.shape[0], Ld, T)
        else:
            t_valid = None if lengths is None else conv_out_lengths(lengths, tm._conv_layers)
        s_valid = None if lengths is None else conv_out_lengths(lengths, sm._conv_layers)
        n, B, D = self.n_pred, x.shape[0], sm._geom.d_out
        Pt, Wt = tm.engine_state()
        t_extras = {} if self.attn_loss_weight > 0 else None
        tgt, t_feats = E.teacher_forward(Pt, Wt, tm._geom, x, t_valid, slots=self.tgt_slots, extras=t_extras)
        was_training = sm.training
        sm.eval()
        P, W, _ = sm.engine_state(sm._weights.train if sm._weights is not None else False)
        c = E.student_forward(P, W, sm._geom, x, s_valid, train=False, heads="all", keep_qkv_last=t_extras is not None)
        sm.train(was_training)
        rec = torch.zeros(n, device=dev, dtype=torch.float32)
        lt = 0 if self.rec_loss_type == "mse" else 1
        if self.sim_loss_weight:
            sim = torch.zeros(n, device=dev, dtype=torch.float32)
            K.distill_loss_sim(c.preds, tgt, self.layer_weights, rec, sim, None, n, B, c.Tq, T, D, lt, 0.0, 0.0)
            per = rec * self.rec_loss_weight + sim * self.sim_loss_weight
        else:
            K.distill_loss(c.preds, tgt, self.layer_weights, rec, None, n, B, c.Tq, T, D, lt, 0.0)
            per = rec * self.rec_loss_weight
        if self.cnn_loss_weight > 0 and c.cnn_out is not None:
            cnn = torch.zeros(1, device=dev, dtype=torch.float32)
            Dc = c.cnn_out.shape[-1]
            K.distill_loss(c.cnn_out.view(1, B, T, Dc), t_feats.reshape(1, B, T, Dc), torch.ones(1, device=dev), cnn, None,
                           1, B, T, T, Dc, 1, 0.0)
            per = torch.cat([per, cnn * self.cnn_loss_weight])
        if t_extras is not None:
            # v_loss is calculate_loss's total (train.py:183-185): the attention-map / value-relation terms belong to it
            al = E.attn_transfer_losses(c, sm._geom, t_extras, tm._geom, loss_type=self.attn_loss_type,
                                        w_attn=float(self.attn_loss_weight), w_vrel=float(self.v_rel_loss_weight),
                                        grad_scale=1.0, valid_s=c.valid, valid_t=t_extras["valid_host"])
            c.attn_grad = None  # (the gradients the kernel also wrote are not needed here)
            per = torch.cat([per, al * torch.tensor([float(self.attn_loss_weight), float(self.v_rel_loss_weight)], device=dev)])
        # train.py:198-199: with random-layer distillation the monitored value is the last layer's own (un-weighted)
        # feature loss, not the weighted total
        loss = (rec[n - 1] if not self.sim_loss_weight else rec[n - 1] + sim[n - 1]) \
            if (self.train_cfg["distil_random_layer"] > 0 and not self.split_head) else per.sum()
        return {"v_loss": loss}

    comm_events = None  # bench.py: a list collecting (start, end) CUDA events of the exposed all-reduce wait

    def optimizer_step(self):
        _, _, G = self.student_model.engine_state(True)
        ev = None
        if self.comm_events is not None and self.reducer.enabled:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        self.reducer.reduce_all(G.flat)
        self.reducer.wait()
        if ev is not None:
            ev[1].record()
            self.comm_events.append(ev)
        self.optimizer.step(grad_scale=1.0 / (self.reducer.world * (self._loss_scale or 1.0)))
