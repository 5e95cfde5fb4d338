import torch.nn as nn
from torch.nn.init import trunc_normal_  # noqa: F401


class DropPath(nn.Identity):
    def __init__(self, drop_prob=0.0, *a, **k):
        super().__init__()
        self.drop_prob = drop_prob


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _DropModule:
    DropPath = DropPath


drop = _DropModule
