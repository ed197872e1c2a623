"""Experiment: N independent engine instances ("lanes") on N CUDA streams of ONE GPU, frames dealt round-robin.
Kernels of one frame form a strict chain; frames are independent, so another lane's kernels can fill the ramp-up / tail of
each launch (and co-run where resources allow).  Prints device-resident frames/s for 1, 2, 3 lanes."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from occnet_b200 import fixtures                    # noqa: E402
from occnet_b200.engine import OccEngine            # noqa: E402

cfg = fixtures.make_cfg('full', num_layers=6)
params = fixtures.init_params(cfg, seed=2, free_bias=fixtures.FREE_BIAS)
metas = fixtures.make_img_metas(cfg)
dev = 'cuda:0'
frames = [[f[0].bfloat16().to(dev).contiguous() for f in fixtures.make_feats(cfg, bs=1, seed=100 + i)] for i in range(3)]
want = ('flow', 'occ_cls')
res = {}
for lanes in (1, 2, 3):
    engs = []
    for _ in range(lanes):
        e = OccEngine(cfg, params, precision='bf16', use_tensor_cores=True, device=dev)
        e.set_cameras(metas)
        e.set_input_dtype(torch.bfloat16)
        engs.append(e)
    streams = [torch.cuda.Stream() for _ in range(lanes)]
    outs = [None] * lanes

    def run(n):
        for i in range(n):
            k = i % lanes
            with torch.cuda.stream(streams[k]):
                outs[k] = engs[k].forward(frames[i % 3], want=want)
    run(60)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        run(300)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    res[lanes] = 300 / best
    print(f'lanes {lanes}: {300 / best:.1f} frames/s ({best / 300 * 1e3:.3f} ms/frame)', flush=True)
    del engs, outs
    torch.cuda.empty_cache()
