"""ncu target: dense BatchNorm(+ReLU6) forward/backward on the largest decoder tensor (2x128^3x64 bf16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anatomask_b200 import ops  # noqa: E402

dev = torch.device('cuda:0')
x = torch.randn(2, 128, 128, 128, 64, device=dev).to(torch.bfloat16).requires_grad_(True)
g = torch.ones(64, device=dev, requires_grad=True)
b = torch.zeros(64, device=dev, requires_grad=True)
run = (torch.zeros(64, device=dev), torch.ones(64, device=dev), torch.zeros((), dtype=torch.long, device=dev))
for _ in range(2):
    y = ops.batch_norm_train(x, g, b, 1e-5, 2, run, 0.1)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print('done')
