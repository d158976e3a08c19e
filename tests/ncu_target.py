"""Short ncu target: one conv shape (default the dominant 64->64 @128^3, batch 2; AMB_NT_CI / AMB_NT_CO / AMB_NT_S override)
forward x3, dgrad x3, wgrad x3."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anatomask_b200 import ops, _lib as L  # noqa: E402

dev = torch.device('cuda:0')
N, S, ci, co = 2, int(os.environ.get('AMB_NT_S', 128)), int(os.environ.get('AMB_NT_CI', 64)), int(os.environ.get('AMB_NT_CO', 64))
x = torch.randn(N, S, S, S, ci, device=dev).to(torch.bfloat16)
dy = torch.randn(N, S, S, S, co, device=dev).to(torch.bfloat16)
w = torch.randn(co, ci, 3, 3, 3, device=dev) / (27 * ci) ** 0.5
wf = ops._pack(w, 27, co, ci, 1, ci * 27, 27)
wd = ops._pack(w, 27, ci, co, 1, 27, ci * 27)
y = torch.empty(N, S, S, S, co, dtype=torch.bfloat16, device=dev)
dx = torch.empty_like(x)
dw = torch.zeros(27, co, ci, device=dev)
for _ in range(3):
    ops._conv_call(L.OP_CONV, L.IMPL_TCGEN05, (N, S, S, S), ci, co, 3, 1, x, y, wf)
for _ in range(3):
    ops._conv_call(L.OP_CONV_DGRAD, L.IMPL_TCGEN05, (N, S, S, S), ci, co, 3, 1, dy, dx, wd)
a = L.WgradArgs(L.OP_CONV, L.IMPL_TCGEN05, N, S, S, S, ci, co, 3, 1, x.data_ptr(), dy.data_ptr(), dw.data_ptr(), 1, 1, 1, 0, 0,
                torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    L.call('amb_conv_wgrad', C.byref(a))
torch.cuda.synchronize()
print('done')
