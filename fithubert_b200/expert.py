"""s3prl upstream plug-in, same surface as reference fithubert/expert.py:9-75 and hubconf.py:3-13."""
from __future__ import annotations

import torch
import torch.nn as nn
import yaml
from torch.nn.utils.rnn import pad_sequence

from .config import CustomStudentModelConfig
from .model import CustomStudentModel


class UpstreamExpert(nn.Module):
    def __init__(self, ckpt, model_config, **kwargs):
        super().__init__()
        if isinstance(model_config, dict):
            cfg = model_config
        else:
            with open(model_config, "r") as f:
                cfg = yaml.load(f, Loader=yaml.FullLoader)
        model_config = dict(cfg["distiller"])
        model_config["init_conv_layers"] = False
        model_config["init_encoder_layers"] = 0
        self.model_config = CustomStudentModelConfig(**model_config)
        self.model = CustomStudentModel(self.model_config)
        if ckpt is not None:
            # Lightning checkpoint (or an already loaded dict): keys under `student_model.` with the prefix cut,
            # read without pytorch_lightning installed (checkpoint.py)
            from .checkpoint import load_student_state_dict
            self.model.load_state_dict(load_student_state_dict(ckpt))
        self.model._disable_projection_heads()

    def get_downsample_rates(self, key: str):
        return 320

    def forward(self, wavs):
        """fithubert/expert.py:52-75 (differentiable like the reference's: with grad enabled and trainable parameters the
        hand-written backward runs behind torch.autograd, so s3prl's upstream fine-tuning works).  The reference pads the wavs and builds `padding_mask = ~(arange(Lmax) < len)`, from
        which the model recovers `len` again (modules/model.py:449-472); here the lengths go to the model directly and
        the zero-padded batch is assembled in HBM (one async copy per wav - DMA straight from the caller's buffer when
        it is pinned or already on the device), so no [B, Lmax] mask or padded host copy is ever made."""
        wav_lens = [len(wav) for wav in wavs]
        dev = self.model.post_extract_proj.weight.device
        if all(w.is_cuda for w in wavs):
            src = pad_sequence(wavs, batch_first=True)
        else:
            src = torch.zeros(len(wavs), max(wav_lens), device=dev, dtype=torch.float32)
            for i, (w, n) in enumerate(zip(wavs, wav_lens)):
                src[i, :n].copy_(w, non_blocking=True)
        results = self.model(source=src, lengths=wav_lens)
        return {"last_hidden_state": results["x"], "hidden_states": results["layer_results"]}


def fithubert(ckpt, model_config, *args, **kwargs):
    assert ckpt is not None and model_config is not None
    return UpstreamExpert(ckpt, model_config, *args, **kwargs)
