"""One native ResNet-50 + FPN forward on 6 x 928 x 1600 synthetic images (bf16, tcgen05): for `ncu --metrics gpu__time_duration.sum`
launch lists of the backbone alone (argv[1] = number of forwards, default 3)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from occnet_b200.backbone import BackboneEngine      # noqa: E402
from oracle import backbone as OB                     # noqa: E402  (parameter initialiser only)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
p = OB.init_params(seed=5)
img = torch.randn(6, 3, 928, 1600, generator=torch.Generator().manual_seed(3)).cuda()
eng = BackboneEngine(p, 6, (928, 1600), precision='bf16', use_tensor_cores=True)
for i in range(n):
    out = eng.forward(img, channels_last_bf16=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    out = eng.forward(img, channels_last_bf16=True)
e1.record()
torch.cuda.synchronize()
print(f'backbone: {e0.elapsed_time(e1) / n:.3f} ms / 6 images')
