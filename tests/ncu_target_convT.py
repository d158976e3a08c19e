"""Short ncu target: the decoder's last up-sampling layer (ConvTranspose3d 64 -> 64, 64^3 -> 128^3, batch 2) forward x3 on the
kz-stacked halo-plane kernel (conv_igemm4t.cu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anatomask_b200 import ops, _lib as L  # noqa: E402

dev = torch.device('cuda:0')
N, S, ci, co = 2, 64, 64, 64
x = torch.randn(N, S, S, S, ci, device=dev).to(torch.bfloat16)
w = torch.randn(ci, co, 4, 4, 4, device=dev) / (8 * ci) ** 0.5
wp = ops._pack_conv(w, True, False)
y = torch.empty(N, 2 * S, 2 * S, 2 * S, co, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops._conv_call(L.OP_CONVT, L.IMPL_TCGEN05, (N, S, S, S), ci, co, 4, 2, x, y, wp)
torch.cuda.synchronize()
print('done', (L.load().amb_last_conv_kernel() or b'').decode())
