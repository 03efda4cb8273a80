"""GPU-box: one eager Hiera-L train step bracketed by cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sam2_unet_b200.synthetic import synthetic_batch  # noqa: E402
from sam2_unet_b200 import SAM2UNet, TrainStep  # noqa: E402
from sam2_unet_b200.params import fill_deterministic_  # noqa: E402

B = int(os.environ.get("PB", "12"))
dev = torch.device("cuda:0")
m = SAM2UNet(model_cfg=os.environ.get("PCFG", "sam2_hiera_l.yaml"), dtype="bf16")
fill_deterministic_(m, 0)
m = m.to(dev).train()
x, mask = synthetic_batch(B, 352, seed=0)
x, mask = x.to(dev), mask.to(dev)
step = TrainStep(m, use_graph=False)
for _ in range(2):
    step(x, mask)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
loss = step(x, mask)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", loss.sum().item())
