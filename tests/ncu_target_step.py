"""ncu target: ONE eager single-stream AnatoMask step (STUNet-B, 2x128^3) inside a cudaProfilerStart/Stop range, after two
warm-up steps.  Used with `ncu --profile-from-start off` for the launch list (time + DRAM bytes of every kernel of a step) and
for the section captures of the HBM-bound kernels (`-k regex:...`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anatomask_b200 import ops  # noqa: E402
from anatomask_b200.trainer import PretrainEngine, build_model  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else 'B'
dev = torch.device('cuda', 0)
torch.manual_seed(1234)
model = build_model(size, (128, 128, 128), anatomask=True)
eng = PretrainEngine(model, lr=1e-4, epochs=1000, anatomask=True, mask_rng='device')
inp = torch.randn(2, 1, 128, 128, 128, device=dev)
ops.NO_SIDE = True
for _ in range(2):
    eng.device_step(inp, 500)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.device_step(inp, 500)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
