"""CPU model of the SCA gather's L1 behaviour: which (query -> SM, co-residency) mapping minimises L2->L1 traffic?

The gather is bound by L2->L1 bandwidth (ncu, profiles/r2_sca_pipe_ncu_raw.csv: 1.48 GB per launch over the crossbar = 6.5 TB/s, L1 hit
rate 47 %), so the lever is the L1 hit rate.  This script replays the exact sampling geometry of the bench workload (camera
projection, 4 pillar anchors, 8 heads x 4 levels x 8 points, clamped 2x2 blocks, row-major value layout: one 64-byte segment per
(token, head)) through a per-SM LRU of 64-byte segments and reports the miss traffic for different mappings.
Sampling offsets use the bias grid only (the 0.02-sigma query-dependent perturbation is ignored)."""
import math
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from occnet_b200 import fixtures                    # noqa: E402
from oracle import bevformer_occ as O               # noqa: E402

cfg = fixtures.make_cfg('full', num_layers=1)
metas = fixtures.make_img_metas(cfg, bs=1)
H, W = cfg['bev_h'], cfg['bev_w']
pc = cfg['pc_range']
ref_3d = O.get_reference_points(H, W, pc[5] - pc[2], cfg['num_points_in_pillar'], '3d', 1)
ref_cam, mask = O.point_sampling(ref_3d, pc, metas)          # (cam, B, Nq, D, 2), (cam, B, Nq, D)
ref_cam = ref_cam[:, 0].numpy(); mask = mask[:, 0].numpy()
ncam, Nq, D, _ = ref_cam.shape
vis = mask.any(-1)                                            # (cam, Nq)
print('visible pairs', int(vis.sum()))
lv = list(zip(cfg['level_h'], cfg['level_w'])) if 'level_h' in cfg else None
if lv is None:
    lv = [(116, 200), (58, 100), (29, 50), (15, 25)]
starts = np.cumsum([0] + [h * w for h, w in lv])[:-1]
Nv = sum(h * w for h, w in lv)
M, L, P = 8, 4, 8
thetas = np.arange(M) * (2 * math.pi / M)
gi = np.stack([np.cos(thetas), np.sin(thetas)], -1)
gi = gi / np.abs(gi).max(-1, keepdims=True)
off = gi[:, None, :] * (np.arange(P)[None, :, None] + 1)      # (M, P, 2) pixels, same for every level


def segments(c, q):
    """64-byte segment ids touched by (camera c, query q), in kernel order: sample i = p*4 + l, 4 corners, 8 heads."""
    out = []
    for p in range(P):
        u, v = ref_cam[c, q, p % D]
        for l, (h, w) in enumerate(lv):
            x = u * w + off[:, p, 0] - 0.5                    # (M,)
            y = v * h + off[:, p, 1] - 0.5
            ok = (y > -1) & (x > -1) & (y < h) & (x < w)
            xb = np.clip(np.floor(x).astype(np.int64), 0, w - 2)
            yb = np.clip(np.floor(y).astype(np.int64), 0, h - 2)
            tok = c * Nv + starts[l] + yb * w + xb
            for dy in (0, w):
                for dx in (0, 1):
                    seg = (tok + dy + dx) * 8 + np.arange(M)
                    out.append(seg[ok])
    return out                                                # list of arrays (one warp-level request each)


def simulate(groups, cap_segments):
    """groups: list over time of lists of queries co-resident on ONE SM; warps of a group interleave request by request."""
    lru = OrderedDict()
    hits = misses = 0
    for grp in groups:
        streams = []
        for q in grp:
            for c in range(ncam):
                if vis[c, q]:
                    streams.append(segments(c, q))
        if not streams:
            continue
        n = max(len(s) for s in streams)
        for i in range(n):
            for s in streams:
                if i < len(s):
                    for seg in s[i].tolist():
                        if seg in lru:
                            lru.move_to_end(seg); hits += 1
                        else:
                            misses += 1
                            lru[seg] = 1
                            if len(lru) > cap_segments:
                                lru.popitem(last=False)
    return hits, misses


def mapping_current(sm, nsm=148, resident=6):
    """CTA b = 4 x-consecutive queries; SM sm runs CTAs sm, sm+148, ...; `resident` of them at a time."""
    ctas = list(range(sm, Nq // 4, nsm))
    return [[4 * b + k for b in ctas[i:i + resident] for k in range(4)] for i in range(0, len(ctas), resident)]


def mapping_tiles(sm, tw, th, nsm=148):
    """BEV tiles of tw x th queries = one co-resident group; tiles dealt round-robin to SMs."""
    tiles = [(tx, ty) for ty in range(0, H, th) for tx in range(0, W, tw)]
    return [[(ty + j) * W + tx + i for j in range(th) for i in range(tw) if ty + j < H and tx + i < W] for k, (tx, ty) in enumerate(tiles) if k % nsm == sm]


def mapping_region(sm, tw, th, nsm=148):
    """SM owns a contiguous run of tiles (consecutive groups on one SM are neighbours too)."""
    tiles = [(tx, ty) for ty in range(0, H, th) for tx in range(0, W, tw)]
    per = -(-len(tiles) // nsm)
    mine = tiles[sm * per:(sm + 1) * per]
    return [[(ty + j) * W + tx + i for j in range(th) for i in range(tw) if ty + j < H and tx + i < W] for tx, ty in mine]


if __name__ == '__main__':
    sms = [int(a) for a in sys.argv[1].split(',')] if len(sys.argv) > 1 else [3, 40, 77, 120]
    cap = int(sys.argv[2]) if len(sys.argv) > 2 else 150 * 1024 // 64
    cases = {'current (4x1 CTAs, b mod 148)': lambda s: mapping_current(s),
             'tiles 6x4 round-robin': lambda s: mapping_tiles(s, 6, 4),
             'tiles 4x6 round-robin': lambda s: mapping_tiles(s, 4, 6),
             'tiles 8x3 round-robin': lambda s: mapping_tiles(s, 8, 3),
             'tiles 5x5 round-robin': lambda s: mapping_tiles(s, 5, 5),
             'region of 6x4 tiles': lambda s: mapping_region(s, 6, 4),
             'region of 8x3 tiles': lambda s: mapping_region(s, 8, 3)}
    for name, fn in cases.items():
        h = m = 0
        for s in sms:
            a, b = simulate(fn(s), cap)
            h += a; m += b
        scale = 148 / len(sms)
        print(f'{name:34s} hit rate {h / (h + m):.3f}  L2->L1 {m * 64 * scale / 1e9:.3f} GB/launch (accesses {(h + m) * scale / 1e6:.1f} M)', flush=True)

if len(sys.argv) > 3 and sys.argv[3] == 'half':
    for name, fn in {'tiles 4x3 (12-warp CTA), half cache': lambda s: mapping_tiles(s, 4, 3),
                     'tiles 6x2 (12-warp CTA), half cache': lambda s: mapping_tiles(s, 6, 2),
                     'tiles 4x2 (8-warp CTA), third cache': lambda s: mapping_tiles(s, 4, 2)}.items():
        h = m = 0
        c = cap // 3 if 'third' in name else cap // 2
        for s in sms:
            a, b = simulate(fn(s), c)
            h += a; m += b
        print(f'{name:40s} hit rate {h / (h + m):.3f}  L2->L1 {m * 64 * 148 / len(sms) / 1e9:.3f} GB/launch', flush=True)
