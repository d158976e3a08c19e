"""timm.models.layers stand-in: exactly the three names the reference imports."""
import torch
from torch import nn


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return x


def to_3tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x, x)
