import torch.nn as nn


class BaseFairseqModel(nn.Module):
    pass
