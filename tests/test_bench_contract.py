"""CPU checks of bench.py: the algorithmic-FLOP constants behind roofline.achieved (SURVEY 8d formulas), the synthetic
workload generators, and the JSON contract of the reference arm (`--impl reference` runs the oracle port on the host)."""
import json
import os
import subprocess
import sys

import torch

import fhb_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _fwd_flops(L, conv_spec, D, F, n_layers, student):
    """multiply-add = 2; GroupNorm / LayerNorm / loss ignored."""
    conv = O.parse_conv_layers(conv_spec)
    n, cin, fl = L, 1, 0.0
    for (c, k, s) in conv:
        n = (n - k) // s + 1
        fl += 2.0 * cin * c * k * n
        cin = c
    T = n
    fl += 2.0 * T * 512 * D                 # post_extract_proj
    fl += 2.0 * T * D * (D / 16) * 128      # grouped positional conv
    t = T // 2 if student else T
    fl += n_layers * (8.0 * t * D * D + 4.0 * t * t * D + 4.0 * t * D * F)
    if student:
        fl += 2.0 * t * 960 * 480           # time-reduction conv
        fl += 12 * (2.0 * t * 480 * 960 + 2.0 * (2 * t) * 480 * 768)  # 12 LayerWiseProjHeads
    return fl


def test_flop_constants_follow_the_survey_formulas():
    for L, want in ((bench.CFG2["Lmax"], bench.FLOP_PER_UTT), (bench.CFG5["Lmax"], bench.CFG5["flop_per_utt"])):
        t = _fwd_flops(L, O.HUBERT_CONV, 768, 3072, 12, False)
        s = _fwd_flops(L, O.FITHUBERT_CONV, 480, 480, 12, True)
        assert abs((t + 3 * s) - want) / want < 2e-3, (L, t + 3 * s, want)   # step = teacher fwd + 3 x student fwd
    s4 = _fwd_flops(bench.CFG4["Lmax"], O.FITHUBERT_CONV, 480, 480, 12, True)
    assert abs(s4 - bench.CFG4["flop_per_utt"]) / s4 < 2e-3                  # inference: student forward only


def test_synthetic_workloads_are_deterministic_and_well_formed():
    a = bench.synth_lengths(32, 249600, 1234)
    assert a == bench.synth_lengths(32, 249600, 1234) and a[0] == 249600 and a == sorted(a, reverse=True)
    assert min(a) >= 249600 - 8000
    u = bench.synth_lengths_uniform(64, 80000, 160000, 1234)
    assert u[0] == 160000 and u == sorted(u, reverse=True) and min(u) >= 80000
    x, pm, lens = bench.synth_batch(3, 4000, 7, lengths=[4000, 3000, 1000])
    assert x.shape == (3, 4000) and pm.dtype == torch.bool and lens == [4000, 3000, 1000]
    assert (~pm).sum(-1).tolist() == lens and float(x[2, 1000:].abs().max()) == 0.0 and float(x[2, :1000].abs().max()) > 0
    cfg = bench.yaml_cfg()
    from fithubert_b200.config import CustomStudentModelConfig
    CustomStudentModelConfig(**cfg["distiller"])  # every key of the restated yaml is an accepted field


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "distill-step audio-sec/sec" and line["unit"] == "audio-s/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    # "reference" when the reference's own classes are importable here (/root/reference, through the fairseq stub),
    # "port" (the oracle restatement) on a box without them
    import bench
    assert line["cpu_baseline"]["kind"] == ("reference" if bench.reference_root() else "port")
    assert line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
