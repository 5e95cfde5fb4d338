import torch.nn as nn


class SoftTargetCrossEntropy(nn.Module):
    pass
