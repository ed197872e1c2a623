"""Python handle of the libocc_b200 frame engine (C ABI: occb200_engine_*).

This is the host-side counterpart of `BEVFormerOccHead.forward` + `get_occ`
(reference: bevformer/dense_heads/bevformer_occ_head.py:99-160, 198-216): parameters come in under
their reference state_dict keys, camera geometry comes from `img_metas` exactly as
`BEVFormerEncoder.point_sampling` reads it (encoder.py:94-101, 133-134).
"""
import ctypes

import numpy as np
import torch

from . import _lib

PRECISIONS = {'fp32': 0, 'bf16': 1}


def _cfg_struct(cfg, precision, use_tensor_cores):
    c = _lib.OccConfig()
    c.bev_h, c.bev_w = cfg['bev_h'], cfg['bev_w']
    c.embed_dims, c.num_heads = cfg['embed_dims'], cfg['num_heads']
    c.num_layers, c.num_cams, c.num_levels = cfg['num_layers'], cfg['num_cams'], cfg['num_levels']
    for i, (h, w) in enumerate(cfg['level_shapes']):
        c.level_h[i], c.level_w[i] = h, w
    c.num_points_in_pillar = cfg['num_points_in_pillar']
    c.sca_points, c.tsa_points = cfg['sca_points'], cfg['tsa_points']
    c.ffn_dim, c.pillar_h, c.out_dim, c.num_classes = cfg['ffn_dim'], cfg['pillar_h'], cfg['out_dim'], cfg['num_classes']
    for i in range(6):
        c.pc_range[i] = cfg['pc_range'][i]
    c.precision = PRECISIONS[precision]
    c.use_tensor_cores = int(use_tensor_cores)
    c.use_cams_embeds = int(cfg.get('use_cams_embeds', True))
    rc = cfg.get('rotate_center', [100, 100])
    c.rotate_center[0], c.rotate_center[1] = int(rc[0]), int(rc[1])
    return c


def camera_params(cfg, img_metas):
    """-> (cam_mat (num_cams,16) f32, zs (D,) f32, img_h, img_w), with the reference's fp32 operation order:
    lidar2img / ego2lidar are cast to fp32 first, then multiplied (encoder.py:100-101, 126)."""
    l2i = torch.from_numpy(np.asarray(img_metas[0]['lidar2img'])).to(torch.float32)
    e2l = torch.from_numpy(np.asarray(img_metas[0]['ego2lidar'])).to(torch.float32)
    cam = torch.matmul(l2i, e2l).reshape(-1, 16).contiguous().numpy()
    Z = cfg['pc_range'][5] - cfg['pc_range'][2]
    D = cfg['num_points_in_pillar']
    zs = (torch.linspace(0.5, Z - 0.5, D, dtype=torch.float32) / Z).contiguous().numpy()
    h, w = img_metas[0]['img_shape'][0][:2]
    return cam, zs, int(h), int(w)


_ROT_CACHE = {}


def rotation_index_map(bev_h, bev_w, angle_deg, center):
    """Index map of the reference's prev_bev rotation (transformer_occ.py:195-205), obtained from the SAME torchvision
    call applied to an image of cell indices -- so ties, centre handling and out-of-image cells are torchvision's, not a
    re-derivation: returns (bev_h*bev_w,) int32, entry q = source cell of output cell q, -1 = outside (zero fill)."""
    key = (bev_h, bev_w, float(angle_deg), tuple(center))
    if key not in _ROT_CACHE:
        from torchvision.transforms.functional import rotate
        idx = torch.arange(1, bev_h * bev_w + 1, dtype=torch.float32).reshape(1, bev_h, bev_w)   # exact in fp32 (< 2^24)
        rot = rotate(idx, float(angle_deg), center=list(center))                                 # nearest, fill 0
        if len(_ROT_CACHE) > 64:
            _ROT_CACHE.clear()
        _ROT_CACHE[key] = (rot.reshape(-1).to(torch.int64) - 1).to(torch.int32).numpy()
    return _ROT_CACHE[key]


class OccEngine:
    def __init__(self, cfg, params, precision='fp32', use_tensor_cores=False, device='cuda:0'):
        if not torch.cuda.is_available():
            raise RuntimeError('OccEngine needs a CUDA device (no CPU fallback)')
        self.cfg = dict(cfg)
        self.precision = precision
        self.device = torch.device(device)
        self.lib = _lib.load()
        self._h = ctypes.c_void_p()
        c = _cfg_struct(cfg, precision, use_tensor_cores)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.occb200_engine_create(ctypes.byref(c), ctypes.byref(self._h)))
            for k, v in params.items():
                if k.endswith('num_batches_tracked'):
                    continue
                a = np.ascontiguousarray(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v, np.float32)
                _lib.check(self.lib.occb200_engine_load_param(self._h, k.encode(), _lib.ptr(a), a.size))
            _lib.check(self.lib.occb200_engine_finalize(self._h))
        self.Nq = cfg['bev_h'] * cfg['bev_w']
        self.vox_shape = (cfg['bev_w'], cfg['bev_h'], cfg['pillar_h'])
        self._pinned = None
        self.feat_dtype, self.feat_channels_last = torch.float32, False

    def set_input_dtype(self, dtype, channels_last=False):
        """Feature levels are handed over as `dtype` from now on (torch.float32, the reference's, or torch.bfloat16);
        `channels_last` (bf16 only): tensors of shape (num_cams, C, h, w) whose MEMORY is (num_cams, h, w, C) -- the
        backbone engine's native output."""
        assert dtype in (torch.float32, torch.bfloat16) and not (channels_last and dtype != torch.bfloat16)
        code = 2 if channels_last else int(dtype == torch.bfloat16)
        _lib.check(self.lib.occb200_engine_set_input_dtype(self._h, code))
        self.feat_dtype, self.feat_channels_last = dtype, bool(channels_last)

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            self.lib.occb200_engine_destroy(h)
            self._h = ctypes.c_void_p()

    def set_cameras(self, img_metas):
        cam, zs, h, w = camera_params(self.cfg, img_metas)
        _lib.check(self.lib.occb200_engine_set_cameras(self._h, _lib.ptr(cam), _lib.ptr(zs), h, w))

    def set_prev_rotation(self, index_map):
        """`index_map` (Nq,) int32 numpy / tensor: source BEV cell of every output cell (-1 = outside), e.g. from
        `rotation_index_map`; None = prev_bev arrives already rotated."""
        if index_map is None:
            _lib.check(self.lib.occb200_engine_set_prev_rotation(self._h, None))
            return
        m = np.ascontiguousarray(index_map.cpu().numpy() if isinstance(index_map, torch.Tensor) else index_map, np.int32)
        assert m.shape == (self.Nq,)
        _lib.check(self.lib.occb200_engine_set_prev_rotation(self._h, _lib.ptr(m)))

    def _check_feats(self, feats, cuda):
        """A mismatched tensor would be an out-of-bounds device read in the pack kernel: fail on the host instead."""
        nc, C = self.cfg['num_cams'], self.cfg['embed_dims']
        if len(feats) != self.cfg['num_levels']:
            raise ValueError(f'expected {self.cfg["num_levels"]} feature levels, got {len(feats)}')
        for l, (f, (h, w)) in enumerate(zip(feats, self.cfg['level_shapes'])):
            if tuple(f.shape) != (nc, C, h, w):
                raise ValueError(f'feature level {l}: shape {tuple(f.shape)} != configured {(nc, C, h, w)}')
            dense = f.permute(0, 2, 3, 1).is_contiguous() if self.feat_channels_last else f.is_contiguous()
            if f.dtype != self.feat_dtype or f.is_cuda != cuda or not dense:
                raise ValueError(f'feature level {l}: need a {"channels-last" if self.feat_channels_last else "contiguous"} '
                                 f'{self.feat_dtype} {"CUDA" if cuda else "CPU (pinned)"} tensor')

    def _feat_ptrs(self, feats):
        arr = (ctypes.c_void_p * 4)()
        for i, f in enumerate(feats):
            arr[i] = f.data_ptr()
        return arr

    def forward(self, feats, prev_bev=None, want=('bev_embed', 'occ', 'flow', 'occ_cls')):
        """feats: 4 CUDA fp32 tensors (num_cams, C, h, w) of one frame.  Returns a dict of CUDA tensors."""
        C = self.cfg['embed_dims']
        X, Y, Z = self.vox_shape
        dev = self.device
        if not self.feat_channels_last:
            feats = [f.contiguous() for f in feats]
        self._check_feats(feats, cuda=True)
        out = {}
        if 'bev_embed' in want:
            out['bev_embed'] = torch.empty((self.Nq, C), dtype=torch.float32, device=dev)
        if 'occ' in want:
            out['occ'] = torch.empty((X, Y, Z, self.cfg['num_classes']), dtype=torch.float32, device=dev)
        if 'flow' in want:
            out['flow'] = torch.empty((X, Y, Z, 2), dtype=torch.float32, device=dev)
        if 'occ_cls' in want:
            out['occ_cls'] = torch.empty((X, Y, Z), dtype=torch.uint8, device=dev)
        if 'occ_cls_i64' in want:
            out['occ_cls_i64'] = torch.empty((X, Y, Z), dtype=torch.int64, device=dev)
        if prev_bev is not None:
            prev_bev = prev_bev.to(device=dev, dtype=torch.float32).reshape(self.Nq, C).contiguous()
        with torch.cuda.device(dev):
            _lib.check(self.lib.occb200_engine_forward(
                self._h, self._feat_ptrs(feats), _lib.ptr(prev_bev), _lib.ptr(out.get('bev_embed')),
                _lib.ptr(out.get('occ')), _lib.ptr(out.get('flow')), _lib.ptr(out.get('occ_cls')),
                _lib.ptr(out.get('occ_cls_i64')), _lib.stream_ptr()))
        return out

    def forward_host(self, feats_host, occ_out=None, flow_out=None):
        """feats_host: 4 pinned CPU fp32 tensors (num_cams, C, h, w).  H2D + frame + D2H + sync inside.
        Returns (occ_cls int64 (X,Y,Z) CPU, flow fp32 (X,Y,Z,2) CPU)."""
        X, Y, Z = self.vox_shape
        if occ_out is None or flow_out is None:
            if self._pinned is None:
                self._pinned = (torch.empty((X, Y, Z), dtype=torch.int64).pin_memory(),
                                torch.empty((X, Y, Z, 2), dtype=torch.float32).pin_memory())
            occ_out, flow_out = self._pinned
        self._check_feats(feats_host, cuda=False)
        arr = (ctypes.c_void_p * 4)()
        for i, f in enumerate(feats_host):
            arr[i] = f.data_ptr()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.occb200_engine_forward_host(self._h, arr, _lib.ptr(occ_out), _lib.ptr(flow_out),
                                                            _lib.stream_ptr()))
        return occ_out, flow_out

    def submit_host(self, slot, feats_host, occ_out, flow_out):
        """Pipelined host-buffer call (slot 0/1): returns immediately; `wait_host(slot)` completes it."""
        self._check_feats(feats_host, cuda=False)
        arr = (ctypes.c_void_p * 4)()
        for i, f in enumerate(feats_host):
            arr[i] = f.data_ptr()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.occb200_engine_submit_host(self._h, slot, arr, _lib.ptr(occ_out), _lib.ptr(flow_out),
                                                           _lib.stream_ptr()))

    def wait_host(self, slot):
        _lib.check(self.lib.occb200_engine_wait_host(self._h, slot))

    def stream_host(self, frames_host):
        """Generator over an iterable of host frames with two frames in flight; yields (occ int64 CPU, flow CPU)
        views of the slot's pinned output buffers (valid until the slot is reused two frames later)."""
        X, Y, Z = self.vox_shape
        if getattr(self, '_stream_outs', None) is None:              # pinned once: cudaHostAlloc costs milliseconds
            self._stream_outs = [(torch.empty((X, Y, Z), dtype=torch.int64).pin_memory(),
                                  torch.empty((X, Y, Z, 2)).pin_memory()) for _ in range(2)]
        outs = self._stream_outs
        pending = []
        for i, fr in enumerate(frames_host):
            slot = i & 1
            if len(pending) == 2:
                s = pending.pop(0)
                self.wait_host(s)
                yield outs[s]
            self.submit_host(slot, fr, *outs[slot])
            pending.append(slot)
        for s in pending:
            self.wait_host(s)
            yield outs[s]

    def enable_taps(self, on=True):
        _lib.check(self.lib.occb200_engine_enable_taps(self._h, int(on)))

    def tap(self, which, layer=0):
        names = {'layer': 0, 'tsa': 1, 'sca': 2, 'voxel': 3, 'tokens': 4}
        w = names[which]
        if w == 4:
            nv = sum(h * w_ for h, w_ in self.cfg['level_shapes'])
            dst = torch.empty((self.cfg['num_cams'], nv, self.cfg['embed_dims']), dtype=torch.float32, device=self.device)
        elif w == 3:
            X, Y, Z = self.vox_shape
            dst = torch.empty((X, Y, Z, self.cfg['out_dim']), dtype=torch.float32, device=self.device)
        else:
            dst = torch.empty((self.Nq, self.cfg['embed_dims']), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.occb200_engine_copy_tap(self._h, w, layer, _lib.ptr(dst), _lib.stream_ptr()))
        return dst

    def project_pillars(self):
        D = self.cfg['num_points_in_pillar']
        nc = self.cfg['num_cams']
        ref = torch.empty((nc, self.Nq, D, 2), dtype=torch.float32, device=self.device)
        mask = torch.empty((nc, self.Nq, D), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.occb200_engine_project_pillars(self._h, _lib.ptr(ref), _lib.ptr(mask), _lib.stream_ptr()))
        return ref, mask

    CATEGORIES = ('pack', 'gemm', 'tsa_gather', 'sca_gather', 'layernorm', 'bev_to_voxel', 'conv3d', 'occ_head')

    def profile(self, on=True):
        _lib.check(self.lib.occb200_engine_profile(self._h, int(on)))

    def profile_read(self):
        """-> {category: (milliseconds, launches)} since the last read (synchronises the device)."""
        ms = np.zeros(8, np.float32)
        n = np.zeros(8, np.int32)
        _lib.check(self.lib.occb200_engine_profile_read(self._h, _lib.ptr(ms), _lib.ptr(n), 8))
        return {c: (float(ms[i]), int(n[i])) for i, c in enumerate(self.CATEGORIES)}

    @property
    def launches_per_frame(self):
        return int(self.lib.occb200_engine_launches_per_frame(self._h))
