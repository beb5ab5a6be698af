#!/usr/bin/env python
"""bench.py - FitHuBERT distillation-step throughput (audio-seconds / second) on B200.

Contract: python bench.py --gpus N --steps K --warmup W [--impl reference]
One "step" = one full distillation step over one batch of synthetic 16 kHz waveforms:
frozen HuBERT-Base teacher forward + student forward/backward + layer-wise loss + gradient
all-reduce (N > 1) + fused AdamW.  Workload at every N: cfg-2 of BASELINE.json per GPU
(32 x 15.6 s, LibriSpeech-bucket lengths, random-init teacher/student), i.e. weak scaling
(global batch 32*N; at N = 8 this is BASELINE configs[2]'s global batch 256).
Prints ONE JSON line on rank 0.  Keys beyond the contract: `roofline` (live CUDA-event time of every tcgen05 GEMM launch
against the measured sustained bf16 peak, per group), `cpu_baseline` (the oracle port on the host cores, bounded
sample), `host_enqueue_ms_per_step` (CPU time to queue one step), `student_fwd` (N = 1: the second half of
BASELINE.json's metric, cfg-4 = the s3prl UpstreamExpert forward on 64 wavs U[5 s, 10 s]).
Other BASELINE configurations: --workload cfg4 (that expert forward alone), --workload cfg5 (FitW2V2: wav2vec 2.0 Base
teacher, 16 mixed-length utterances U[10 s, 30 s]).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 16000
CFG2 = dict(B=32, Lmax=249600)
# algorithmic FLOPs per step / per audio-second: SURVEY 8d (408.0 GFLOP per 15.6 s utterance)
FLOP_PER_UTT = 408.01e9
# the other BASELINE.json configurations: not the default bench line, selected with --workload
CFG4 = dict(B=64, Lmax=160000, Lmin=80000, flop_per_utt=35.23e9)    # s3prl UpstreamExpert forward, heads removed
CFG5 = dict(B=16, Lmax=480000, Lmin=160000, flop_per_utt=843.48e9)  # FitW2V2: wav2vec 2.0 Base teacher, 30 s, mixed


def yaml_cfg():
    """data/conf/fithubert.yaml of the reference, restated (the file itself is not on the GPU box)."""
    return {
        "teacher": {"teacher_model": "hubert_base_ls960.pt"},
        "train": dict(num_epochs=100, gpus=1, batch_size=32, accumulate_grad_batches=1, use_fp16=True,
                      monitor_losses=True, cnn_loss_weight=0, rec_loss_weight=1.0, rec_loss_type="mse",
                      sim_loss_weight=0, attn_loss_weight=0, attn_loss_type="kldiv", v_rel_loss_weight=0,
                      distil_random_layer=11, random_layer_weight=0.1, delete_projections=False, specaug=False),
        "distiller": dict(
            extractor_mode="default",
            conv_feature_layers="[(128, 10, 5)] + [(256, 1, 1)] + [(256, 3, 2)] * 4 + [(512, 1, 1)] + [(512, 2, 2)] * 2",
            feature_grad_mult=1.0, conv_bias=False, n_mels=0, enable_log_mel=False, conv_pos=128, conv_pos_groups=16,
            pos_conv_depth=1, max_positions=100000, layer_type="transformer", encoder_layers=12, encoder_embed_dim=480,
            encoder_ffn_embed_dim=480, encoder_attention_heads=12, activation_fn="gelu", layer_norm_first=False,
            dropout=0.1, attention_dropout=0.1, activation_dropout=0.1, encoder_layerdrop=0.0, dropout_input=0.05,
            final_dim=256, pred_head_final_dim=768, pred_head_inter_dim=0, layerwise_proj=True, pred_layer_id="[11]",
            init_conv_layers=False, init_encoder_layers=0, enable_tr_layer=True, tr_conv1d_kernel=2, tr_layer_index=0,
            tr_reduce_factor=2, tr_layer_type="conv1d", checkpoint_activations=False, required_seq_len_multiple=1,
            crop_seq_to_multiple=1),
        "optimizer": dict(name="AdamW_with_schedule", lr=5.e-4, warmup_proportion=0.05, betas=[0.9, 0.98], eps=1.e-6,
                          weight_decay=1.e-6),
    }


def synth_lengths(B, Lmax, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    lens = Lmax - torch.randint(0, 8001, (B,), generator=g)
    lens[0] = Lmax
    return sorted(lens.tolist(), reverse=True)


def synth_lengths_uniform(B, Lmin, Lmax, seed):
    """cfg-4 / cfg-5: lengths U[Lmin, Lmax] sorted descending, the longest pinned to Lmax."""
    import torch
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(Lmin, Lmax + 1, (B,), generator=g)
    lens[0] = Lmax
    return sorted(lens.tolist(), reverse=True)


def synth_batch(B, Lmax, seed, pin=False, lengths=None):
    import torch
    g = torch.Generator().manual_seed(seed)
    if lengths is None:
        lengths = synth_lengths(B, Lmax, seed)
    x = 0.1 * torch.randn(B, Lmax, generator=g)
    pm = ~(torch.arange(Lmax).unsqueeze(0) < torch.tensor(lengths).unsqueeze(1))
    x.masked_fill_(pm, 0.0)
    if pin:
        x, pm = x.pin_memory(), pm.pin_memory()
    return x, pm, lengths


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [s.strip() for s in r.split(",")]
            if len(f) < 8:
                continue
            try:
                mx = float(f[2])
                if float(f[3]) < 300.0:  # idle sample (before / after the timed region)
                    continue
                sm.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU reference arm
def reference_root():
    """Where the reference's own Python files can be imported from (they run unmodified through oracle/fairseq_stub):
    FHB_REFERENCE, /root/reference (the build container) or baseline/_ref; None on a box that has neither."""
    for r in (os.environ.get("FHB_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if r and os.path.isfile(os.path.join(r, "modules", "model.py")):
            return r
    return None


def cpu_reference_classes_step_fn(ref, n_utts, Lmax, seed=1234):
    """cpu_baseline.kind = "reference": the reference's OWN classes (modules/model.py CustomStudentModel, and its
    ConvFeatureExtractionModel / TransformerEncoder wired as HuBERT-Base, oracle/gen_golden.RefTeacher) imported
    unmodified through the fairseq stub, the restated loss and optimizer step around them (train.py / s3prl cannot be
    imported: Lightning and s3prl are absent).  Same workload and weights-by-key as the port."""
    import torch
    sys.path[:0] = [os.path.join(ROOT, "oracle", "fairseq_stub"), ref, os.path.join(ROOT, "oracle")]
    os.environ["FHB_REFERENCE"] = ref
    import fhb_oracle as O
    import gen_golden as GG  # imports the reference modules
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    ycfg = yaml_cfg()["distiller"]
    student = GG.CustomStudentModel(GG.CustomStudentModelConfig(**ycfg))
    student.load_state_dict(O.init_student_state(O.student_config(), 0))
    tcfg = O.teacher_config()
    teacher = GG.RefTeacher(tcfg)
    teacher.load_state_dict(O.init_teacher_state(tcfg, 1))
    student.eval()  # dropout identity, like the port
    teacher.eval()
    lengths = synth_lengths(CFG2["B"], Lmax, seed)[:n_utts]
    lengths[0] = Lmax
    x, pm = O.synth_batch(n_utts, Lmax, lengths, seed)
    w = O.layer_weights(12, 0.1)
    params = {k: p for k, p in student.named_parameters()}
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    v2 = {k: torch.zeros_like(v) for k, v in params.items()}
    state = {"step": 0}

    def step():
        with torch.no_grad():
            t = teacher(x, pm)
        s_ = student(source=x, padding_mask=pm)
        loss, _ = O.distill_loss(s_["projections"], t["layer_results"], w)
        loss.backward()
        state["step"] += 1
        with torch.no_grad():
            for k, p in params.items():
                if p.grad is None:
                    continue
                O.adamw_step(p, p.grad, m[k], v2[k], state["step"], 5e-4)
                p.grad = None
        return float(loss)

    return step, sum(lengths) / SR, torch.get_num_threads()


def cpu_arm(n_utts, Lmax):
    """(step fn, audio seconds per step, cores, kind)"""
    ref = reference_root()
    if ref is not None:
        try:
            return (*cpu_reference_classes_step_fn(ref, n_utts, Lmax), "reference")
        except Exception as e:  # noqa: BLE001 - the port is always available
            sys.stderr.write(f"bench: reference classes unavailable ({type(e).__name__}: {e}); timing the oracle port\n")
    return (*cpu_reference_step_fn(n_utts, Lmax), "port")


def cpu_reference_step_fn(n_utts, Lmax, seed=1234):
    """The reference's algorithm for this path, restated (oracle/fhb_oracle.py), fp32 on the host cores:
    teacher fwd + student fwd/bwd + loss + AdamW for `n_utts` utterances of the cfg-2 length distribution."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fhb_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    scfg, tcfg = O.student_config(), O.teacher_config()
    ssd = {k: v.requires_grad_(k not in ("upsampler.weight", "upsampler.bias")) for k, v in O.init_student_state(scfg, 0).items()}
    tsd = O.init_teacher_state(tcfg, 1)
    lengths = synth_lengths(CFG2["B"], Lmax, seed)[:n_utts]
    lengths[0] = Lmax
    x, pm = O.synth_batch(n_utts, Lmax, lengths, seed)
    w = O.layer_weights(12, 0.1)
    m = {k: torch.zeros_like(v) for k, v in ssd.items()}
    v2 = {k: torch.zeros_like(v) for k, v in ssd.items()}
    state = {"step": 0}

    def step():
        with torch.no_grad():
            t = O.teacher_forward(tsd, tcfg, x, pm)
        s = O.student_forward(ssd, scfg, x, pm)
        loss, _ = O.distill_loss(s["projections"], t["layer_results"], w)
        loss.backward()
        state["step"] += 1
        with torch.no_grad():
            for k, p in ssd.items():
                if p.grad is None:
                    continue
                O.adamw_step(p, p.grad, m[k], v2[k], state["step"], 5e-4)
                p.grad = None
        return float(loss)

    return step, sum(lengths) / SR, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_utts = 2
    step, audio_s, cores, kind = cpu_arm(n_utts, CFG2["Lmax"])
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = audio_s * args.steps / dt
    sample = f"{n_utts} of the 32 cfg-2 utterances (15.6 s each) per step, fp32, {args.warmup} warm-up + {args.steps} timed full steps"
    print(json.dumps({
        "impl": "reference", "metric": "distill-step audio-sec/sec", "value": val, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg-2: FitHuBERT distillation step, 32 x 15.6 s per GPU (bounded CPU sample)",
                   "per_gpu_batch": 32, "utterance_s": 15.6},
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------- B200 arm
def measure_student_fwd(dev, B, Lmax, Lmin, steps, warmup, seed, world):
    """cfg-4 of BASELINE.json: the s3prl UpstreamExpert forward (fithubert/expert.py:52-75) - list of B variable-length
    wavs in, all hidden states out, projection heads removed, no_grad.  `value`: wavs resident in HBM; `e2e`: the same
    call with pinned HOST wavs (pad + mask on the host, H2D inside the timed region) and a D2H read of the
    last frame of last_hidden_state."""
    import torch
    import torch.distributed as dist
    import fithubert_b200 as F
    from fithubert_b200 import kernels as K
    torch.manual_seed(0)
    ex = F.UpstreamExpert(None, {"distiller": yaml_cfg()["distiller"]}).to(dev).eval()
    for p_ in ex.parameters():  # feature extraction (s3prl's frozen-upstream mode): the forward is differentiable like
        p_.requires_grad_(False)  # the reference's, so gradients are switched off here, at the caller
    torch.set_grad_enabled(False)
    lengths = synth_lengths_uniform(B, Lmin, Lmax, seed)
    g = torch.Generator().manual_seed(seed)
    host = [(0.1 * torch.randn(n, generator=g)).pin_memory() for n in lengths]
    wavs = [w.to(dev) for w in host]
    audio_s = sum(lengths) / SR

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        out = ex(wavs)
    barrier()
    K.reset_counters()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(steps):
        out = ex(wavs)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = K.launch_count()
    assert out["last_hidden_state"].shape[0] == B and len(out["hidden_states"]) == 12
    for _ in range(2):
        float(ex(host)["last_hidden_state"][0, -1, 0])
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        float(ex(host)["last_hidden_state"][0, -1, 0])
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    torch.set_grad_enabled(True)
    flops = CFG4["flop_per_utt"] * (Lmax / CFG4["Lmax"]) * B
    return {
        "metric": "student fwd audio-sec/sec", "value": audio_s * world * steps / (ms / 1e3), "unit": "audio-s/s",
        "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
        "config": {"workload": f"cfg-4: UpstreamExpert.forward (student inference, all hidden states), {B} wavs "
                               f"U[{Lmin / SR:.0f} s, {Lmax / SR:.0f} s] per GPU, random-init weights",
                   "per_gpu_batch": B, "global_batch": B * world, "utterance_s": Lmax / SR,
                   "l2": "working set > 1 GB (>> 126 MB L2); no explicit flush",
                   "step_tflops_algorithmic": flops / 1e12},
        "e2e": {"value": audio_s * world * steps / (e2e_ms / 1e3), "unit": "audio-s/s",
                "h2d_bytes_per_step": 4 * B * Lmax, "d2h_bytes_per_step": 2,
                "api": "UpstreamExpert.forward(list of pinned host wavs)"},
        "gpu_launches": launches,
        "step_tflops_per_s": flops / (ms / steps / 1e3) / 1e12,
    }


def measure_membound(dev, step_obj, peak_gbs):
    """HBM-bound kernels of the step at their cfg-2 shapes, each timed ALONE with CUDA events (L2 flushed before every
    launch by writing a 512 MB buffer): algorithmic bytes, microseconds, GB/s and the fraction of the
    measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)."""
    import torch
    from fithubert_b200 import kernels as K, lib as L
    f16, f32 = torch.float16, torch.float32
    flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)
    rows_t, Et, rows_s, Es = 32 * 779, 768, 32 * 389, 480
    rnd = lambda *sh, dt=f16: torch.randn(*sh, device=dev, dtype=f32).to(dt)
    out = []

    def run(name, nbytes, fn, reps=5):
        fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = sorted(ts)[len(ts) // 2]
        gbs = nbytes / us / 1e3
        out.append({"kernel": name, "bytes": int(nbytes), "us": round(us, 1), "gbs": round(gbs, 1),
                    "frac": round(gbs / peak_gbs, 3)})

    g, b = torch.ones(Et, device=dev), torch.zeros(Et, device=dev)
    x, y = rnd(rows_t, Et), torch.empty(rows_t, Et, device=dev, dtype=f16)
    run("layernorm_fwd (teacher 24928 x 768, fp16 in/out)", 4 * rows_t * Et, lambda: K.layernorm_fwd(x, g, b, y))
    gs, bs = torch.ones(Es, device=dev), torch.zeros(Es, device=dev)
    x32, y16, y32 = rnd(rows_s, Es, dt=f32), torch.empty(rows_s, Es, device=dev, dtype=f16), torch.empty(rows_s, Es, device=dev)
    mean, rstd = torch.empty(rows_s, device=dev), torch.empty(rows_s, device=dev)
    run("layernorm_fwd32 (student 12448 x 480, fp32 in, fp16 + fp32 out)", 10 * rows_s * Es,
        lambda: K.layernorm_fwd32(x32, gs, bs, y16, y32, mean, rstd))
    d32, d16 = rnd(rows_s, Es, dt=f32), rnd(rows_s, Es)
    o16, o32 = torch.empty_like(d16), torch.empty_like(d32)
    dg, db, dsum = (torch.zeros(Es, device=dev) for _ in range(3))
    run("layernorm_bwd32 (student, fp32 + fp16 grads in, fp16 + fp32 out)", 16 * rows_s * Es,
        lambda: K.layernorm_bwd32(d32, x32, gs, mean, rstd, dg, db, dy2=d16, dx=o16, dx32=o32, dxsum=dsum))
    del x, y, x32, y16, y32, d32, d16, o16, o32
    # conv layer 0 (+ GroupNorm + GELU): teacher C = 512, student C = 128 with the saved gelu'
    Bq, Lq = 32, 249600
    T0 = (Lq - 10) // 5 + 1
    wave = 0.1 * torch.randn(Bq, Lq, device=dev)
    for C0, gp, name in ((512, False, "teacher"), (128, True, "student, + saved gelu'")):
        w0, gm, bt = torch.randn(C0, 1, 10, device=dev) * 0.3, torch.ones(C0, device=dev), torch.zeros(C0, device=dev)
        stat = torch.empty(Bq, 65, device=dev, dtype=torch.float64)
        m0, r0 = torch.empty(Bq, C0, device=dev), torch.empty(Bq, C0, device=dev)
        yo = torch.empty(Bq, T0, C0, device=dev, dtype=f16)
        go = torch.empty_like(yo) if gp else None
        nb = 2 * Bq * Lq * 4 + Bq * T0 * C0 * 2 * (2 if gp else 1)
        run(f"conv0_gn_gelu_fwd ({name}: stats + normalise, {C0} ch)", nb,
            lambda: K.conv0_fwd(wave, w0, gm, bt, T0, stat, m0, r0, yo, gp_out=go), reps=3)
        del yo, go
    # loss + gradient over the 12 stacked projections
    n, Tq, D = 12, 778, 768
    pred, tgt = rnd(n, Bq, Tq, D), rnd(n, Bq, 779, D)
    wl, ll = torch.full((n,), 0.1, device=dev), torch.zeros(n, device=dev)
    dcs = torch.zeros(n, D, device=dev)
    run("distill_loss_fwd_bwd (12 x 32 x 778 x 768: read pred + tgt, write dpred)", 3 * n * Bq * Tq * D * 2,
        lambda: K.distill_loss(pred, tgt, wl, ll, pred, n, Bq, Tq, 779, D, 0, 1.0, dbias=dcs, dbias_layer_stride=D), reps=3)
    del pred, tgt
    # AdamW over the student's own flat buffers
    opt = step_obj.optimizer
    opt.step_count += 0
    P_, W_, G_ = step_obj.student_model.engine_state(True)
    if opt._table is None:
        opt._build()
    m_keep, v_keep = opt.m.clone(), opt.v.clone()  # lr = 0, wd = 0: parameters stay, the moments are restored
    run("adamw_multi (31.2 M parameters: p, g, m, v)", 28 * G_.numel,
        lambda: K.adamw_multi(opt._table, opt._n, opt._max_n, 0.0, 0.9, 0.98, 1e-6, 0.0, 1, 0, 1.0), reps=3)
    opt.m.copy_(m_keep)
    opt.v.copy_(v_keep)
    gcol = rnd(rows_s, 3 * Es)
    cs = torch.zeros(3 * Es, device=dev)
    run("colsum (12448 x 1440 fp16 -> fp32 bias gradient)", rows_s * 3 * Es * 2, lambda: K.colsum(gcol, cs))
    return out


def measure_gpu_eager(dev, B, Lmax, lengths_seed):
    """The bar SURVEY 2.2 names: the same distillation step in plain PyTorch eager on this GPU (cuDNN / cuBLAS /
    torch SDPA-free manual attention exactly as the reference computes it), fp16 autocast like the reference's AMP
    recipe.  The restated reference arithmetic (oracle/fhb_oracle.py) runs on CUDA tensors; one warm-up + two timed
    steps, outside every other timed region.  A reported baseline, like cpu_baseline."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fhb_oracle as O
    scfg, tcfg = O.student_config(), O.teacher_config()
    ssd = {k: v.to(dev).requires_grad_(k not in ("upsampler.weight", "upsampler.bias"))
           for k, v in O.init_student_state(scfg, 0).items()}
    tsd = {k: v.to(dev) for k, v in O.init_teacher_state(tcfg, 1).items()}
    lengths = synth_lengths(B, Lmax, lengths_seed)
    x, pm = O.synth_batch(B, Lmax, lengths, lengths_seed)
    x, pm = x.to(dev), pm.to(dev)
    w = O.layer_weights(12, 0.1)
    opt = torch.optim.AdamW([p for p in ssd.values() if p.requires_grad], lr=5e-4, betas=(0.9, 0.98), eps=1e-6,
                            weight_decay=1e-6, fused=True)
    scaler = torch.amp.GradScaler("cuda", init_scale=2.0 ** 16)
    # the oracle's mask helpers build CPU tensors: give them the device through torch's default
    prev = torch.get_default_device() if hasattr(torch, "get_default_device") else None
    torch.set_default_device(dev)

    def step():
        with torch.autocast("cuda", dtype=torch.float16):
            with torch.no_grad():
                t = O.teacher_forward(tsd, tcfg, x, pm)
            s_ = O.student_forward(ssd, scfg, x, pm)
            loss, _ = O.distill_loss(s_["projections"], t["layer_results"], w)
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        return loss

    try:
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        n = 2
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
    finally:
        torch.set_default_device(prev if prev is not None else "cpu")
    del ssd, tsd, opt
    torch.cuda.empty_cache()
    return {"value": sum(lengths) / SR / (ms / 1e3), "unit": "audio-s/s", "ms_per_step": ms, "dtype": "fp16 autocast (torch.amp)",
            "what": "restated reference arithmetic in PyTorch eager (cuDNN conv1d, cuBLAS linear / bmm, fused AdamW) on the "
                    "same GPU and batch; 1 warm-up + 2 timed steps"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fhb")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-student-fwd", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the PyTorch-eager GPU baseline leg")
    ap.add_argument("--no-membound", action="store_true", help="skip the per-kernel HBM table")
    ap.add_argument("--profile", action="store_true", help="2 device steps then exit (for ncu launch lists)")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4", "cfg5"],
                    help="cfg2 (default, the bench line): distillation step 32 x 15.6 s; cfg4: UpstreamExpert "
                         "inference forward 64 x <=10 s; cfg5: FitW2V2 distillation step 16 x <=30 s")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--lmax", type=int, default=None)
    args = ap.parse_args()
    wl = {"cfg2": CFG2, "cfg4": CFG4, "cfg5": CFG5}[args.workload]
    args.batch = args.batch or wl["B"]
    args.lmax = args.lmax or wl["Lmax"]
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the FitHuBERT hot path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the overlapped gradient all-reduce shares the GPU with the backward: a handful of NCCL CTAs is plenty for
        # 125 MB over NVLink, and the persistent kernels leave exactly that many SMs free (FHB_COMM_SMS)
        os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("FHB_COMM_SMS", "8"))
        dist.init_process_group("nccl", device_id=dev)
    import fithubert_b200 as F
    from fithubert_b200 import kernels as K

    torch.manual_seed(0)
    cfg = yaml_cfg()
    cfg["train"]["batch_size"] = args.batch
    B, Lmax = args.batch, args.lmax
    if args.workload == "cfg4":
        out = measure_student_fwd(dev, B, Lmax, CFG4["Lmin"], args.steps, max(args.warmup, 3), 1234 + rank, world)
        if rank == 0:
            out.update({"n_gpus": world, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                        "dtype": "fp16", "data": "synthetic"})
            print(json.dumps(out))
        if world > 1:
            dist.destroy_process_group()
        return
    flop_per_utt = FLOP_PER_UTT * (Lmax / CFG2["Lmax"])
    wl_name = "cfg-2: FitHuBERT distillation step (HuBERT-Base teacher fwd"
    fixed_lengths = None
    if args.workload == "cfg5":
        cfg["teacher"]["teacher_model"] = "wav2vec_small.pt"  # data/conf/fitwav2vec2.yaml: the only recipe difference
        flop_per_utt = CFG5["flop_per_utt"] * (Lmax / CFG5["Lmax"])
        wl_name = "cfg-5: FitW2V2 distillation step, mixed lengths U[10 s, 30 s] (wav2vec 2.0 Base teacher fwd"
        fixed_lengths = synth_lengths_uniform(B, min(CFG5["Lmin"], Lmax), Lmax, 1234 + rank)
    step_obj = F.W2V2Distil(cfg, device=dev)
    step_obj.configure_optimizers(total_steps=1000)
    x_host, pm_host, lengths = synth_batch(B, Lmax, 1234 + rank, pin=True, lengths=fixed_lengths)
    x_dev = x_host.to(dev)
    audio_s = sum(lengths) / SR

    def dev_step():
        step_obj.optimizer.zero_grad()
        ll = step_obj.fused_forward_backward(x_dev, None, lengths, grad_scale=1.0)
        step_obj.optimizer_step()
        return ll

    # D2H read of every step's result: an asynchronous copy of the loss into a pinned slot right behind the step, consumed
    # by the host one step later (what a training loop that logs its loss does) - a blocking float(loss) would drain the
    # launch queue after every step and leave the GPU idle while the CPU queues the next one (FHB_E2E_SYNC=1 does that)
    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"i": 0, "last": None}
    sync_read = os.environ.get("FHB_E2E_SYNC", "0") == "1"

    def e2e_step():
        loss = step_obj.training_step({"x": x_host, "padding_mask": pm_host})
        if sync_read:
            e2e_state["last"] = float(loss.detach())
            return
        i = e2e_state["i"]
        loss_host[i & 1].copy_(loss.detach(), non_blocking=True)
        loss_ev[i & 1].record()
        if i > 0:  # the previous step's loss has had a whole step to arrive
            loss_ev[(i - 1) & 1].synchronize()
            e2e_state["last"] = float(loss_host[(i - 1) & 1])
        e2e_state["i"] = i + 1

    def e2e_drain():
        i = e2e_state["i"]
        if i > 0 and not sync_read:
            loss_ev[(i - 1) & 1].synchronize()
            e2e_state["last"] = float(loss_host[(i - 1) & 1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.profile:
        for _ in range(2):
            dev_step()
        torch.cuda.synchronize()
        if os.environ.get("FHB_GEMM_STEPLOG"):
            # diagnostic (trace build of the library, tools/gemm_steplog.py): start / end of every GEMM launch of 3 more
            # steps queued back to back, in GPU nanoseconds
            import ctypes as C
            from fithubert_b200 import lib as L
            L.lib().fhb_gemm_steplog_read(None, 0, 1)
            for _ in range(3):
                dev_step()
            host = (C.c_longlong * (6 * 8192))()
            n = L.lib().fhb_gemm_steplog_read(host, 8192, 1)
            with open(os.environ["FHB_GEMM_STEPLOG"], "w") as f:
                for i in range(n):
                    e = host[6 * i:6 * i + 6]
                    f.write("%d %d %d %d %d %d %d\n" % (e[0], e[1], e[2] >> 32, e[2] & 0xFFFFFFFF, e[3] >> 32, e[3] & 0xFFFFFFFF, e[5] - e[4]))
        return
    sampler = ClockSampler(local)
    def measure_e2e():
        # ---- end to end through the public API with host buffers
        for _ in range(2):
            e2e_step()
        e2e_drain()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        e2e_drain()  # the last step's loss is on the host before the clock stops
        barrier()
        return time.perf_counter() - t0

    for _ in range(max(args.warmup, 3)):
        dev_step()
    # FHB_BENCH_E2E_FIRST=1: run the end-to-end loop before the device-timed one (diagnostic: on a power-capped part the
    # loop that runs later sees lower clocks)
    e2e_s = measure_e2e() if os.environ.get("FHB_BENCH_E2E_FIRST", "0") == "1" else None
    if rank == 0:
        sampler.start()
        time.sleep(0.3)  # let nvidia-smi come up; the rows kept are those taken under load (see stop())
    barrier()  # after rank 0's sleep: no rank may enter the timed region while another is still outside it
    K.reset_counters()
    step_obj.comm_events = [] if world > 1 else None  # CUDA-event pairs around the main stream's wait for the all-reduce
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(args.steps):
        ll = dev_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    comm_ms = sum(a.elapsed_time(b) for a, b in (step_obj.comm_events or [])) / max(1, args.steps)
    step_obj.comm_events = None
    launches = K.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(ll.sum())

    # ---- same steps again with a CUDA-event pair around every GEMM launch: live kernel time + flops
    # (single stream for this pass: an event pair only brackets its own kernel when nothing else runs beside it)
    streams_env = os.environ.get("FHB_STREAMS")
    os.environ["FHB_STREAMS"] = "0"
    K.enable_gemm_timing(True)
    for _ in range(args.steps):
        dev_step()
    torch.cuda.synchronize()
    if streams_env is None:
        del os.environ["FHB_STREAMS"]
    else:
        os.environ["FHB_STREAMS"] = streams_env
    gemm_ms, gemm_flops, gemm_calls = K.gemm_timing_summary()
    # where the aggregate comes from: the teacher's encoder GEMMs (24 928 rows, K = 768 / 3072) are tensor-bound, the
    # student's 12 448 x 480 GEMMs are launch / tail-bound, the student conv stack (K = 128 .. 768 over 1.6 M rows) and
    # the wgrads are HBM-bound
    Tt = step_obj.teacher_model.model._geom
    rows_t = B * (((Lmax - 400) // 320) + 1)

    def classify(shape):
        rows, n, k = shape
        if rows == rows_t and {n, k} <= {Tt.E, 3 * Tt.E, Tt.F}:
            return "teacher_encoder"
        if rows >= 4 * rows_t:
            return "conv_stacks"
        return "student_and_rest"

    groups = K.gemm_timing_groups(classify)
    K.enable_gemm_timing(False)

    # ---- host side of one step: how long the CPU needs to ENQUEUE a step (no synchronisation inside).  If this is close
    # to the device time, the end-to-end number below - which synchronises every step - is launch-bound, not copy-bound
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev_step()
    host_enqueue_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()

    if e2e_s is None:
        e2e_s = measure_e2e()

    t = torch.tensor([ms, e2e_s * 1e3, comm_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, comm_ms = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak_src = "measured (MEASURED_PEAKS.json, bf16_tflops_sustained)"
        peak = float(peaks["bf16_tflops_sustained"])
    except (OSError, KeyError, ValueError):
        peak, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained)"
    value = audio_s * world * args.steps / (ms / 1e3)
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    # DRAM bytes per fhb_gemm launch (dram__bytes_read.sum + dram__bytes_write.sum averaged over the 225 launches of one
    # cfg-2 step, from the committed ncu capture profiles/gemm_traffic.json; null for other workloads / when absent)
    traffic = traffic_src = None
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "gemm_traffic.json")
    if args.workload == "cfg2" and os.path.isfile(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["traffic_bytes_per_launch"], "profiles/gemm_traffic.json: " + tj["source"]

    out = {
        "metric": "distill-step audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
        "precision": "fp16 storage (weights, activations, loss-scaled gradients), fp32 accumulate, fp32 student residual "
                     "stream / statistics / parameter gradients / optimizer state",
        "config": {"workload": f"{wl_name} + student fwd/bwd + "
                               f"loss + allreduce + AdamW), {B} x {Lmax / SR:.1f} s per GPU, random-init weights",
                   "per_gpu_batch": B, "global_batch": B * world, "utterance_s": Lmax / SR,
                   "l2": "per-step working set is several GB (>> 126 MB L2); no explicit flush",
                   "step_tflops_algorithmic": flop_per_utt * B / 1e12},
        "clocks": clocks,
        "e2e": {"value": audio_s * world * args.steps / (e2e_ms / 1e3), "unit": "audio-s/s",
                "h2d_bytes_per_step": x_host.numel() * 4 + 4 * B, "d2h_bytes_per_step": 4,
                "api": "W2V2Distil.training_step({'x','padding_mask'}) with pinned host tensors",
                "readback": "loss copied to pinned host memory after every step (async), read by the host one step later; "
                            "FHB_E2E_SYNC=1 blocks on it every step"},
        "gpu_launches": launches,
        "comm_exposed_ms": comm_ms if world > 1 else 0.0,  # main-stream time spent waiting for the gradient all-reduce
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "roofline": {"bound": "tensor", "kernel": "fhb_gemm_kernel (tcgen05, all variants)", "achieved": achieved,
                     "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "launches_per_step": gemm_calls / max(1, args.steps),
                     "gemm_ms_per_step": gemm_ms / args.steps,
                     "groups": {g: {"tflops": (f / (ms / 1e3) / 1e12) if ms > 0 else 0.0,
                                    "frac": (f / (ms / 1e3) / 1e12 / peak) if ms > 0 and peak else None,
                                    "ms_per_step": ms / args.steps, "launches_per_step": c / args.steps}
                                for g, (ms, f, c) in sorted(groups.items())},
                     "step_frac": (flop_per_utt * B / (ms / args.steps / 1e3) / 1e12) / peak},
        "loss": loss_val,
    }
    if args.workload == "cfg2" and world == 1 and not args.no_student_fwd:
        # the second half of BASELINE.json's metric: student inference forward (cfg-4, s3prl UpstreamExpert)
        sf = measure_student_fwd(dev, CFG4["B"], CFG4["Lmax"], CFG4["Lmin"], args.steps, 3, 1234, 1)
        out["student_fwd"] = {k: sf[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "gpu_launches")}
        out["student_fwd"]["workload"] = sf["config"]["workload"]
    if args.workload == "cfg2" and world == 1 and not args.no_membound:
        try:
            hbm = float(peaks.get("hbm_gbs", 6545.0))
            out["membound"] = {"peak_gbs": hbm, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback",
                               "kernels": measure_membound(dev, step_obj, hbm)}
        except Exception as e:  # noqa: BLE001 - a diagnostic table must not cost the bench line
            out["membound"] = {"error": f"{type(e).__name__}: {e}"}
    if args.workload == "cfg2" and world == 1 and not args.no_eager:
        try:
            out["gpu_eager_baseline"] = measure_gpu_eager(dev, B, Lmax, 1234)
        except Exception as e:  # noqa: BLE001
            out["gpu_eager_baseline"] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
    if not args.no_cpu_baseline:
        n_utts = 2
        stepf, a_s, cores, kind = cpu_arm(n_utts, Lmax)
        stepf()
        t0 = time.perf_counter()
        n = 3
        for _ in range(n):
            stepf()
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": a_s * n / dt, "unit": "audio-s/s", "cores": cores, "kind": kind,
                               "sample": f"{n_utts} of the {B} utterances ({Lmax / SR:.1f} s each), fp32 "
                                         f"{'reference classes through the fairseq stub' if kind == 'reference' else 'oracle port'}"
                                         f", 1 warm-up + {n} timed full steps"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
