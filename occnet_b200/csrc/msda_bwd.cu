// Backward of multi-scale deformable attention at the operator boundary (mmcv `_ext.ms_deform_attn_backward`
// argument meaning; reference call site bevformer/modules/multi_scale_deformable_attn_function.py:150-160).
// SURVEY section 8f rank 4 ("next" row): completes the operator API for users who train through the plugin.
//
//   grad_value[b, pix, m, c]  += go[b,q,m,c] * w[b,q,m,l,p] * (bilinear corner weight)       (atomic)
//   grad_attn [b,q,m,l,p]      = sum_c go[c] * bilinear(value)[c]
//   grad_loc  [b,q,m,l,p,(x,y)] = (W_l, H_l) * w * sum_c go[c] * d bilinear / d(w_im, h_im)
// One warp per (b, q, head); lane = channel (looping when C > 32); warp-shuffle reductions over the channels.
#include "common.cuh"
#include "kernels.cuh"

namespace occ {

namespace {

__global__ void __launch_bounds__(256)
msda_backward_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ lstart, const float* __restrict__ loc, const float* __restrict__ wts,
                     const float* __restrict__ grad_out, int B, int Nv, int M, int C, int Nq, int L, int P,
                     float* __restrict__ grad_value, float* __restrict__ grad_loc, float* __restrict__ grad_attn)
{
    const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // (b*Nq + q)*M + m
    if (item >= (int64_t)B * Nq * M) return;
    const int lane = threadIdx.x & 31;
    const int m = (int)(item % M);
    const int64_t bq = item / M;
    const int b = (int)(bq / Nq);
    const float* lp = loc + item * (int64_t)L * P * 2;
    const float* wp = wts + item * (int64_t)L * P;
    const float* gop = grad_out + item * C;
    const int64_t st = (int64_t)M * C;                                                       // pixel stride in `value`
    for (int l = 0; l < L; ++l) {
        const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
        const int64_t base = (((int64_t)b * Nv + lstart[l]) * M + m) * C;
        for (int p = 0; p < P; ++p) {
            const float w_im = lp[(l * P + p) * 2] * (float)W - 0.5f;
            const float h_im = lp[(l * P + p) * 2 + 1] * (float)H - 0.5f;
            const float aw = wp[l * P + p];
            float g_attn = 0.f, g_w = 0.f, g_h = 0.f;
            if (h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W) {       // warp-uniform
                const int h_lo = (int)floorf(h_im), w_lo = (int)floorf(w_im);
                const float lh = h_im - h_lo, lw = w_im - w_lo, hh = 1.f - lh, hw = 1.f - lw;
                const bool top = h_lo >= 0, bot = h_lo + 1 <= H - 1, lef = w_lo >= 0, rig = w_lo + 1 <= W - 1;
                const int64_t o1 = base + ((int64_t)h_lo * W + w_lo) * st, o2 = o1 + st, o3 = o1 + (int64_t)W * st, o4 = o3 + st;
                for (int c = lane; c < C; c += 32) {
                    const float go = gop[c];
                    const float v1 = (top && lef) ? value[o1 + c] : 0.f, v2 = (top && rig) ? value[o2 + c] : 0.f;
                    const float v3 = (bot && lef) ? value[o3 + c] : 0.f, v4 = (bot && rig) ? value[o4 + c] : 0.f;
                    const float t = go * aw;
                    if (top && lef) atomicAdd(grad_value + o1 + c, t * hh * hw);
                    if (top && rig) atomicAdd(grad_value + o2 + c, t * hh * lw);
                    if (bot && lef) atomicAdd(grad_value + o3 + c, t * lh * hw);
                    if (bot && rig) atomicAdd(grad_value + o4 + c, t * lh * lw);
                    g_attn += go * (hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4);
                    g_w += t * (hh * (v2 - v1) + lh * (v4 - v3));
                    g_h += t * (hw * (v3 - v1) + lw * (v4 - v2));
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                g_attn += __shfl_xor_sync(0xffffffffu, g_attn, o);
                g_w += __shfl_xor_sync(0xffffffffu, g_w, o);
                g_h += __shfl_xor_sync(0xffffffffu, g_h, o);
            }
            if (lane == 0) {
                grad_attn[item * (int64_t)L * P + l * P + p] = g_attn;
                grad_loc[(item * (int64_t)L * P + l * P + p) * 2] = g_w * (float)W;
                grad_loc[(item * (int64_t)L * P + l * P + p) * 2 + 1] = g_h * (float)H;
            }
        }
    }
}

}  // namespace

int launch_msda_backward(const float* value, const int64_t* shapes, const int64_t* lstart, const float* loc,
                         const float* wts, const float* grad_out, int B, int Nv, int M, int C, int Nq, int L, int P,
                         float* grad_value, float* grad_loc, float* grad_attn, cudaStream_t stream)
{
    const int64_t items = (int64_t)B * Nq * M;
    if (items == 0) return 0;
    msda_backward_kernel<<<ceil_div(items, 8), 256, 0, stream>>>(value, shapes, lstart, loc, wts, grad_out, B, Nv, M, C, Nq,
                                                                L, P, grad_value, grad_loc, grad_attn);
    OCC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace occ
