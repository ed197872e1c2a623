"""One engine configuration (environment variables select kernel variants): total ms/frame + per-category ms/frame."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from occnet_b200 import fixtures                    # noqa: E402
from occnet_b200.engine import OccEngine            # noqa: E402

prec = os.environ.get('AB_PRECISION', 'bf16')
tc = int(os.environ.get('AB_TC', '1'))
cfg = fixtures.make_cfg('full', num_layers=6)
params = fixtures.init_params(cfg, seed=2, free_bias=fixtures.FREE_BIAS)
metas = fixtures.make_img_metas(cfg)
dev = 'cuda:0'
dt = torch.bfloat16 if prec == 'bf16' else torch.float32
frames = [[f[0].to(dt).to(dev).contiguous() for f in fixtures.make_feats(cfg, bs=1, seed=100 + i)] for i in range(3)]
eng = OccEngine(cfg, params, precision=prec, use_tensor_cores=bool(tc), device=dev)
eng.set_cameras(metas)
eng.set_input_dtype(dt)
want = ('flow', 'occ_cls')
n = int(os.environ.get('AB_FRAMES', '200'))
for i in range(40):
    eng.forward(frames[i % 3], want=want)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        eng.forward(frames[i % 3], want=want)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / n)
eng.profile(True)
for i in range(32):
    eng.forward(frames[i % 3], want=want)
prof = eng.profile_read()
print(json.dumps({'ms_per_frame': round(best, 4), 'launches': eng.launches_per_frame,
                  'cats': {k: round(v[0] / 32, 4) for k, v in prof.items() if v[0] > 0}}))
