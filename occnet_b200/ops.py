"""Operator-level API mirroring what the reference binds from mmcv's compiled extension.

  ext_module.ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_locations,
                                    attention_weights, im2col_step=...)
      reference call site: bevformer/modules/multi_scale_deformable_attn_function.py:118-124
  MultiScaleDeformableAttnFunction_fp32.apply(...)                     :90-128
  dvr.render_forward(sigma, origin, points, tindex, grid, phase)       tools/ray_iou/lib/dvr/dvr.cpp:39-48

Same argument meaning and error behaviour (RuntimeError on non-CUDA / non-contiguous input or a batch
that does not divide im2col_step).  torch is only the tensor container: the arithmetic happens in
libocc_b200 through the C ABI.
"""
import torch

from . import _lib


def _require(t, name, dtype=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f'{name} must be a torch.Tensor')
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor (libocc_b200 has no CPU path)')
    if not t.is_contiguous():
        raise RuntimeError(f'{name} tensor has to be contiguous')
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f'{name} must be {dtype}, got {t.dtype}')


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
                           im2col_step=64):
    _require(value, 'value', torch.float32)
    _require(spatial_shapes, 'spatial_shapes', torch.int64)
    _require(level_start_index, 'level_start_index', torch.int64)
    _require(sampling_locations, 'sampling_loc', torch.float32)
    _require(attention_weights, 'attn_weight', torch.float32)
    B, Nv, M, C = value.shape
    _, Nq, _, L, P, _ = sampling_locations.shape
    out = torch.empty((B, Nq, M * C), dtype=torch.float32, device=value.device)
    lib = _lib.load()
    with torch.cuda.device(value.device):
        _lib.check(lib.occb200_ms_deform_attn_forward(
            _lib.ptr(value), _lib.ptr(spatial_shapes), _lib.ptr(level_start_index), _lib.ptr(sampling_locations),
            _lib.ptr(attention_weights), B, Nv, M, C, Nq, L, P, int(im2col_step), _lib.ptr(out), _lib.stream_ptr()))
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
                            grad_output, grad_value, grad_sampling_loc, grad_attn_weight, im2col_step=64):
    """mmcv `_ext.ms_deform_attn_backward` drop-in: fills the three PRE-ZEROED gradient tensors in place."""
    for t, n in ((value, 'value'), (sampling_locations, 'sampling_loc'), (attention_weights, 'attn_weight'),
                 (grad_output, 'grad_output'), (grad_value, 'grad_value'), (grad_sampling_loc, 'grad_sampling_loc'),
                 (grad_attn_weight, 'grad_attn_weight')):
        _require(t, n, torch.float32)
    _require(spatial_shapes, 'spatial_shapes', torch.int64)
    _require(level_start_index, 'level_start_index', torch.int64)
    B, Nv, M, C = value.shape
    _, Nq, _, L, P, _ = sampling_locations.shape
    lib = _lib.load()
    with torch.cuda.device(value.device):
        _lib.check(lib.occb200_ms_deform_attn_backward(
            _lib.ptr(value), _lib.ptr(spatial_shapes), _lib.ptr(level_start_index), _lib.ptr(sampling_locations),
            _lib.ptr(attention_weights), _lib.ptr(grad_output), B, Nv, M, C, Nq, L, P, int(im2col_step),
            _lib.ptr(grad_value), _lib.ptr(grad_sampling_loc), _lib.ptr(grad_attn_weight), _lib.stream_ptr()))


class MultiScaleDeformableAttnFunction_fp32(torch.autograd.Function):
    """Drop-in for the reference autograd Function, forward and backward (inputs are cast to fp32 like
    `custom_fwd(cast_inputs=torch.float32)`, multi_scale_deformable_attn_function.py:93)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        value = value.float().contiguous()
        sampling_locations = sampling_locations.float().contiguous()
        attention_weights = attention_weights.float().contiguous()
        value_spatial_shapes = value_spatial_shapes.contiguous()
        value_level_start_index = value_level_start_index.contiguous()
        ctx.im2col_step = im2col_step
        out = ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                     attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, aw = ctx.saved_tensors
        grad_value = torch.zeros_like(value)
        grad_loc = torch.zeros_like(loc)
        grad_aw = torch.zeros_like(aw)
        ms_deform_attn_backward(value, shapes, lsi, loc, aw, grad_output.float().contiguous(), grad_value, grad_loc,
                                grad_aw, ctx.im2col_step)
        return grad_value, None, None, grad_loc, grad_aw, None


MultiScaleDeformableAttnFunction_fp16 = MultiScaleDeformableAttnFunction_fp32   # reference: both branches pick fp32


def render_forward(sigma, origin, points, tindex, grid, phase='test'):
    """dvr.render_forward drop-in: -> [pred_dist (N,M), gt_dist (N,M), coord_index (N,M,3)] fp32 CUDA tensors."""
    for t, n in ((sigma, 'sigma'), (origin, 'origin'), (points, 'points'), (tindex, 'tindex')):
        _require(t, n, torch.float32)
    if phase != 'test':
        raise RuntimeError(f'UNKNOWN / unsupported PHASE NAME: {phase} (only "test" is on the metric path)')
    N, M = points.shape[0], points.shape[1]
    T, Z, Y, X = [int(g) for g in grid]
    assert tuple(sigma.shape) == (N, T, Z, Y, X), (sigma.shape, grid)
    pred = torch.empty((N, M), dtype=torch.float32, device=sigma.device)
    gt = torch.empty_like(pred)
    coord = torch.empty((N, M, 3), dtype=torch.float32, device=sigma.device)
    lib = _lib.load()
    with torch.cuda.device(sigma.device):
        _lib.check(lib.occb200_render_forward(_lib.ptr(sigma), _lib.ptr(origin), _lib.ptr(points), _lib.ptr(tindex),
                                              N, T, Z, Y, X, M, _lib.ptr(pred), _lib.ptr(gt), _lib.ptr(coord),
                                              _lib.stream_ptr()))
    torch.cuda.current_stream().synchronize()      # the reference op returns after cudaDeviceSynchronize (dvr.cu:384)
    return [pred, gt, coord]


def _no_autograd(name, *tensors):
    """The module-level mirror is an INFERENCE path: these two building blocks have no backward.  Silently cutting the
    graph would leave projection / FFN / norm weights without gradients, so a training-mode call fails loudly instead
    (only `MultiScaleDeformableAttnFunction_fp32` is differentiable, as the reference's op is)."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise RuntimeError(f'occnet_b200.ops.{name}: no autograd support (inference path); wrap the call in '
                           f'torch.no_grad() or detach the inputs')


def linear(x, weight, bias=None, residual=None, act=0):
    """fp32 y = act(x W^T + b) (+ residual) through the CUDA-core GEMM (module-level API mirror)."""
    _no_autograd('linear', x, weight, bias, residual)
    _require(x, 'x', torch.float32)
    _require(weight, 'weight', torch.float32)
    K = x.shape[-1]
    N = weight.shape[0]
    x2 = x.reshape(-1, K)
    out = torch.empty((x2.shape[0], N), dtype=torch.float32, device=x.device)
    res = None if residual is None else residual.reshape(-1, N).contiguous()
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.occb200_linear_f32(_lib.ptr(x2), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(res), _lib.ptr(out),
                                          x2.shape[0], N, K, act, _lib.stream_ptr()))
    return out.reshape(*x.shape[:-1], N)


def layer_norm(x, gamma, beta):
    _no_autograd('layer_norm', x, gamma, beta)
    _require(x, 'x', torch.float32)
    C = x.shape[-1]
    x2 = x.reshape(-1, C)
    out = torch.empty_like(x2)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.occb200_layernorm_f32(_lib.ptr(x2), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(out), x2.shape[0],
                                             C, _lib.stream_ptr()))
    return out.reshape(x.shape)
