"""Ray-mIoU / mAVE on device + the one collective of the path (section 8e).

Host-side mirror of `projects/mmdet3d_plugin/datasets/ray_metrics.py`:
  generate_lidar_rays :63-86 (constant ray bundle), main :200-257 (per-frame loop + final scores).
The per-frame work (process_one_sample x2 + calc_metrics' counter updates) is one CUDA launch
(`occb200_ray_metric_accumulate`); the 187 counters stay on the device until `finalize`.
"""
import math

import numpy as np
import torch

from . import _lib

NUM_CLASSES = 17
FLOW_CLASSES = 8
NUM_COUNTERS = 11 * NUM_CLASSES          # gt_cnt[17] pred_cnt[17] tp[3][17] ave[3][17] ave_count[3][17]


def generate_lidar_rays():
    """ray_metrics.py:63-86 -> (14040, 3) float32 unit vectors."""
    pitch_angles = []
    for k in range(10):
        pitch_angles.append(-(math.pi / 2 - math.atan(k + 1)))
    while pitch_angles[-1] < 0.21:
        pitch_angles.append(pitch_angles[-1] + (pitch_angles[-1] - pitch_angles[-2]))
    rays = []
    for pitch in pitch_angles:
        for az in np.arange(0, 360, 1):
            az = np.deg2rad(az)
            rays.append((np.cos(pitch) * np.cos(az), np.cos(pitch) * np.sin(az), np.sin(pitch)))
    return np.array(rays, dtype=np.float32)


class RayMetric:
    def __init__(self, device='cuda:0'):
        self.device = torch.device(device)
        self.rays = torch.from_numpy(generate_lidar_rays()).to(self.device)
        self.counters = torch.zeros(NUM_COUNTERS, dtype=torch.float64, device=self.device)
        self.lib = _lib.load()

    def reset(self):
        self.counters.zero_()

    def add_frame(self, sem_pred, flow_pred, sem_gt, flow_gt, origins, return_pcd=False):
        """sem_* (200,200,16) uint8 CUDA, flow_* (200,200,16,2) fp32 CUDA, origins (T,3) or (1,T,3) f32/f64."""
        dev = self.device
        sem_pred = sem_pred.to(dev, torch.uint8).contiguous(); sem_gt = sem_gt.to(dev, torch.uint8).contiguous()
        flow_pred = flow_pred.to(dev, torch.float32).contiguous(); flow_gt = flow_gt.to(dev, torch.float32).contiguous()
        origins = torch.as_tensor(origins).reshape(-1, 3)
        is64 = origins.dtype == torch.float64
        origins = origins.to(dev, torch.float64 if is64 else torch.float32).contiguous()
        T, M = origins.shape[0], self.rays.shape[0]
        pp = pg = None
        if return_pcd:
            pp = torch.empty((T * M, 4), dtype=torch.float32, device=dev)
            pg = torch.empty((T * M, 4), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self.lib.occb200_ray_metric_accumulate(
                _lib.ptr(sem_pred), _lib.ptr(flow_pred), _lib.ptr(sem_gt), _lib.ptr(flow_gt), _lib.ptr(origins),
                int(is64), T, _lib.ptr(self.rays), M, _lib.ptr(self.counters), _lib.ptr(pp), _lib.ptr(pg),
                _lib.stream_ptr()))
        return (pp, pg) if return_pcd else None

    def all_reduce(self):
        """The single collective of the path: SUM of 187 fp64 counters over the ranks (NCCL / NVLink)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.counters, op=dist.ReduceOp.SUM)
        return self.counters

    def finalize(self):
        return finalize_counters(self.counters.detach().cpu().numpy())


def finalize_counters(v):
    """ray_metrics.py:191-195, 248-253 from the 187-vector -> dict(iou (3,16), ave (16,), miou, mave, score)."""
    v = np.asarray(v, np.float64)
    n = NUM_CLASSES
    gt, pred = v[:n], v[n:2 * n]
    tp = v[2 * n:5 * n].reshape(3, n)
    ave = v[5 * n:8 * n].reshape(3, n).copy()
    ave_count = v[8 * n:11 * n].reshape(3, n)
    ave[:, FLOW_CLASSES:] = np.nan                      # non-flow classes are NaN in the reference (:153-156)
    with np.errstate(divide='ignore', invalid='ignore'):
        iou = np.stack([(tp[j] / (gt + pred - tp[j]))[:-1] for j in range(3)])
        ave_l = ave[1][:-1] / ave_count[1][:-1]
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', RuntimeWarning)          # all-NaN slices (classes absent from the fixture)
            miou = np.nanmean(iou)
            mave = np.nanmean(ave_l)
    score = miou * 0.9 + max(1 - mave, 0.0) * 0.1
    return dict(iou=iou, ave=ave_l, miou=float(miou), mave=float(mave), score=float(score))
