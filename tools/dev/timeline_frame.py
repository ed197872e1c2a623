"""Dev aid: one frame with OCC_GEMM_TIMELINE=1 (set by the caller): every tcgen05 GEMM prints its per-CTA timeline."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from occnet_b200 import fixtures
from occnet_b200.engine import OccEngine
cfg = fixtures.make_cfg('full', num_layers=2)
eng = OccEngine(cfg, fixtures.init_params(cfg, seed=2), precision='bf16', use_tensor_cores=True, device='cuda:0')
eng.set_cameras(fixtures.make_img_metas(cfg))
fr = [f[0].contiguous().cuda() for f in fixtures.make_feats(cfg, bs=1, seed=100)]
for _ in range(2):
    eng.forward(fr, want=('occ_cls',))
torch.cuda.synchronize()
