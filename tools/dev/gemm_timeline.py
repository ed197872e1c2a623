"""In-kernel globaltimer timeline of every tcgen05 GEMM launch of ONE frame (OCC_GEMM_TIMELINE=1 makes every launch synchronous and
prints per-CTA stamps: 0 start, 1 setup done, 2 MMA warp ready, 3/6/9 first A block of tile 0/1/2 landed, 4/7/10 accumulator ready,
5/8/11 epilogue done, 15 end; ns relative to the earliest CTA start).  Usage: python tools/dev/gemm_timeline.py 2> timeline.log"""
import os
import sys

os.environ['OCC_GEMM_TIMELINE'] = '1'
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from occnet_b200 import fixtures                    # noqa: E402
from occnet_b200.engine import OccEngine            # noqa: E402

cfg = fixtures.make_cfg('full', num_layers=2)
params = fixtures.init_params(cfg, seed=2, free_bias=fixtures.FREE_BIAS)
metas = fixtures.make_img_metas(cfg)
frames = [f[0].bfloat16().to('cuda:0').contiguous() for f in fixtures.make_feats(cfg, bs=1, seed=100)]
eng = OccEngine(cfg, params, precision='bf16', use_tensor_cores=True, device='cuda:0')
eng.set_cameras(metas)
eng.set_input_dtype(torch.bfloat16)
for i in range(3):
    sys.stderr.write(f'==== frame {i}\n')
    eng.forward(frames, want=('flow', 'occ_cls'))
torch.cuda.synchronize()
