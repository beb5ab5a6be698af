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

    @torch.no_grad()
    def forward(self, wavs):
        wav_lens = torch.LongTensor([len(wav) for wav in wavs])
        src = pad_sequence(wavs, batch_first=True)
        padding_mask = ~torch.lt(torch.arange(int(max(wav_lens))).unsqueeze(0), wav_lens.unsqueeze(1))
        results = self.model(source=src, padding_mask=padding_mask)
        return {"last_hidden_state": results["x"], "hidden_states": results["layer_results"]}


def fithubert(ckpt, model_config, *args, **kwargs):
    assert ckpt is not None and model_config is not None
    return UpstreamExpert(ckpt, model_config, *args, **kwargs)
