"""data/conf/ex.yaml recipe (DistilHuBERT-style student: teacher-shaped conv stack, 2 x 768-d layers, 12 heads, no TR layer,
SplitLinear head over teacher layers 3 / 7 / 11, L1 + cosine) at the cfg-2 batch (B x 15.6 s), with and without the
attention-map / value-relation terms (attn_loss_weight, v_rel_loss_weight > 0; train.py:64-77,327-378).
Prints ms per training step (CUDA events) and the per-kernel share of the extra launches.
usage: python tools/attn_recipe_bench.py [B] [steps] [case index 0-3: only that case, 1 warm-up step (for ncu)]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json

import torch

import bench
import fithubert_b200 as F
from fithubert_b200 import lib as L


def ex_cfg(attn_w, vrel_w, attn_type):
    cfg = bench.yaml_cfg()
    cfg["train"].update(batch_size=4, accumulate_grad_batches=1, rec_loss_type="l1", rec_loss_weight=1.0, sim_loss_weight=1.0,
                        attn_loss_weight=attn_w, attn_loss_type=attn_type, v_rel_loss_weight=vrel_w, distil_random_layer=0,
                        random_layer_weight=0)
    cfg["distiller"].update(conv_feature_layers="[(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512,2,2)] * 2", feature_grad_mult=0.1,
                            encoder_layers=2, encoder_embed_dim=768, encoder_ffn_embed_dim=3072, encoder_attention_heads=12,
                            activation_dropout=0.0, dropout_input=0.1, layerwise_proj=False, pred_layer_id="[3, 7, 11]",
                            enable_tr_layer=False, init_conv_layers=False, init_encoder_layers=0)
    cfg["optimizer"].update(lr=2.e-4, warmup_proportion=0.07)
    return cfg


def run(B, steps, attn_w, vrel_w, attn_type, warmup=3):
    torch.manual_seed(0)
    teacher = F.TeacherWrapper(F.TeacherModel(kind="hubert").cuda())
    step = F.W2V2Distil(ex_cfg(attn_w, vrel_w, attn_type), teacher_model=teacher, device="cuda")
    step.configure_optimizers(total_steps=1000)
    Lmax = 249600
    lens = bench.synth_lengths(B, Lmax, 1234)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(B, Lmax, generator=g) * 0.1
    for b, n in enumerate(lens):
        x[b, n:] = 0
    x = x.cuda()
    batch = {"x": x, "padding_mask": None, "lengths": lens}
    for _ in range(warmup):
        loss = step.training_step(batch)
    torch.cuda.synchronize()
    L.reset_counters()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(steps):
        loss = step.training_step(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    al = step.last_attn_losses
    return {"attn_loss_weight": attn_w, "v_rel_loss_weight": vrel_w, "attn_loss_type": attn_type, "ms_per_step": round(ms, 3),
            "audio_s_per_s": round(sum(lens) / 16000 / ms * 1e3, 1), "loss": round(float(loss), 5),
            "attn_loss": None if al is None else round(float(al[0]), 6), "v_rel_loss": None if al is None else round(float(al[1]), 6),
            "launches_per_step": L.launch_count() // steps, "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    cases = ((0, 0, "kldiv"), (1.0, 0, "kldiv"), (1.0, 1.0, "kldiv"), (1.0, 1.0, "mse"))
    only = int(sys.argv[3]) if len(sys.argv) > 3 else None
    for i, (a, v, t) in enumerate(cases):
        if only is not None and i != only:
            continue
        print(json.dumps(dict(run(B, steps, a, v, t, warmup=3 if only is None else 1), B=B)), flush=True)
        torch.cuda.empty_cache()
