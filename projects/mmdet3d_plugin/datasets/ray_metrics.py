"""Ray-mIoU / mAVE under the reference import path (`ray_metrics.main`, `process_one_sample`, `generate_lidar_rays`):
same arguments and return conventions, evaluated on the GPU by libocc_b200 (one launch per frame, counters on device)."""
import numpy as np
import torch

from occnet_b200.metric import RayMetric, generate_lidar_rays   # noqa: F401

occ_class_names = ['car', 'truck', 'trailer', 'bus', 'construction_vehicle', 'bicycle', 'motorcycle', 'pedestrian',
                   'traffic_cone', 'barrier', 'driveable_surface', 'other_flat', 'sidewalk', 'terrain', 'manmade',
                   'vegetation', 'free']
flow_class_names = occ_class_names[:8]


def process_one_sample(sem_pred, lidar_rays, output_origin, flow_pred, device='cuda:0'):
    """-> (T*M, 4) float32 rows [class, distance (m), flow_x, flow_y] for one volume."""
    rm = RayMetric(device)
    sem = torch.from_numpy(np.ascontiguousarray(sem_pred).astype(np.uint8))
    flow = torch.from_numpy(np.ascontiguousarray(flow_pred).astype(np.float32))
    pp, _ = rm.add_frame(sem, flow, sem, flow, torch.as_tensor(output_origin), return_pcd=True)
    return pp.cpu().numpy()


def main(sem_pred_list, sem_gt_list, flow_pred_list, flow_gt_list, lidar_origin_list, device='cuda:0', verbose=True,
         distributed=False, count=None):
    """reference `ray_metrics.main` (:200-257): same five lists, same returned scores.

    The reference evaluates on rank 0 only with the full, de-duplicated result list (`tools/test.py:242`,
    `apis/test.py:130`), so by default NO collective runs here (a rank-0-only call cannot hang an initialised
    process group).  `distributed=True` is the sharded form: every rank passes the frames of its contiguous shard and
    the 187 counters are SUM all-reduced; `count` (list of bool, default all True) marks the frames this rank owns
    uniquely -- pass `occnet_b200.dist.owned_unique(...)`-derived flags so that the sampler's wrap-around padding
    is not counted twice."""
    rm = RayMetric(device)
    for i, (sp, sg, fp, fg, orig) in enumerate(zip(sem_pred_list, sem_gt_list, flow_pred_list, flow_gt_list,
                                                   lidar_origin_list)):
        if count is not None and not count[i]:
            continue
        sp = torch.as_tensor(np.reshape(sp, [200, 200, 16]).astype(np.uint8))
        sg = torch.as_tensor(np.reshape(sg, [200, 200, 16]).astype(np.uint8))
        fp = torch.as_tensor(np.reshape(fp, [200, 200, 16, 2]).astype(np.float32))
        fg = torch.as_tensor(np.reshape(fg, [200, 200, 16, 2]).astype(np.float32))
        rm.add_frame(sp, fp, sg, fg, torch.as_tensor(orig))
    if distributed:
        rm.all_reduce()
    fin = rm.finalize()
    if verbose:
        for i, name in enumerate(occ_class_names[:-1]):
            print(f'{name:22s} IoU@1 {fin["iou"][0][i]:.3f}  IoU@2 {fin["iou"][1][i]:.3f}  IoU@4 {fin["iou"][2][i]:.3f}  '
                  f'AVE {fin["ave"][i]:.3f}')
        print(f'MEAN mIoU {fin["miou"]:.4f}  mAVE {fin["mave"]:.4f}')
        print(' --- Occ score:', fin['score'])
    return fin
