// Memory-bound glue kernels of the path: camera-feature packing (transformer_occ.py:207-227),
// BEV positional encoding (mmdet LearnedPositionalEncoding), LayerNorm (encoder norms), casts.
#include "common.cuh"
#include "kernels.cuh"

namespace occ {

namespace {

// [cam][C][hw] f32 -> [cam][Nv][C] T (+ cams_embeds[cam][c], then + level_embed[c]; same order as the reference)
template <typename T>
__global__ void pack_level_kernel(const float* __restrict__ feat, const float* __restrict__ cams_embeds,
                                  const float* __restrict__ level_embed, int C, int hw, int Nv, int start,
                                  T* __restrict__ tokens)
{
    __shared__ float tile[32][33];
    const int cam = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;                 // 32 x 8
    const float* src = feat + (int64_t)cam * C * hw;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + i * 8, p = p0 + tx;
        tile[ty + i * 8][tx] = (p < hw) ? __ldg(src + (int64_t)c * hw + p) : 0.f;
    }
    __syncthreads();
    const int c = c0 + tx;
    const float ce = cams_embeds ? cams_embeds[cam * C + c] : 0.f;
    const float le = level_embed[c];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = p0 + ty + i * 8;
        if (p < hw) {
            float v = tile[tx][ty + i * 8];
            if (cams_embeds) v = v + ce;
            v = v + le;
            tokens[((int64_t)cam * Nv + start + p) * C + c] = from_f32<T>(v);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
layernorm256_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ pos, int rows, float* __restrict__ y_f32, T* __restrict__ y_t,
                    T* __restrict__ y_pos_t)
{
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* xr = x + (int64_t)row * 256 + lane * 8;
    float v[8];
    load8(xr, v);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / 256.f);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; ss = fmaf(d, d, ss); }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * (1.f / 256.f) + 1e-5f);
    float g[8], b[8], y[8];
    load8(gamma + lane * 8, g);
    load8(beta + lane * 8, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = (v[i] - mean) * rstd * g[i] + b[i];
    const int64_t o = (int64_t)row * 256 + lane * 8;
    if (y_f32) store8(y_f32 + o, y);
    if (y_t) store8(y_t + o, y);
    if (y_pos_t) {
        float pv[8];
        load8(pos + o, pv);
#pragma unroll
        for (int i = 0; i < 8; ++i) pv[i] += y[i];
        store8(y_pos_t + o, pv);
    }
}

template <typename T>
__global__ void prepare_query_kernel(const float* __restrict__ q, const float* __restrict__ pos, int64_t n8,
                                     float* __restrict__ q_f32, T* __restrict__ q_t, T* __restrict__ q_pos_t)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    float v[8], p[8];
    load8(q + i * 8, v);
    load8(pos + i * 8, p);
    if (q_f32) store8(q_f32 + i * 8, v);
    if (q_t) store8(q_t + i * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) p[k] += v[k];
    if (q_pos_t) store8(q_pos_t + i * 8, p);
}

__global__ void bev_pos_kernel(const float* __restrict__ row_embed, const float* __restrict__ col_embed, int bev_h,
                               int bev_w, int half, float* __restrict__ pos)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int C = 2 * half;
    if (i >= (int64_t)bev_h * bev_w * C) return;
    const int c = (int)(i % C);
    const int q = (int)(i / C);
    const int x = q % bev_w, y = q / bev_w;
    pos[i] = (c < half) ? col_embed[x * half + c] : row_embed[y * half + (c - half)];
}

template <typename T>
__global__ void cast_kernel(const float* __restrict__ s, T* __restrict__ d, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = from_f32<T>(s[i]);
}

}  // namespace

template <typename T>
int launch_pack_level(const float* feat, const float* cams_embeds, const float* level_embed, int num_cams, int C,
                      int hw, int Nv, int start, T* tokens, cudaStream_t stream)
{
    OCC_CHECK(C % 32 == 0, "pack_level: C must be a multiple of 32");
    dim3 grid(ceil_div(hw, 32), C / 32, num_cams), block(32, 8);
    pack_level_kernel<T><<<grid, block, 0, stream>>>(feat, cams_embeds, level_embed, C, hw, Nv, start, tokens);
    OCC_CUDA(cudaGetLastError());
    return 0;
}
template int launch_pack_level<float>(const float*, const float*, const float*, int, int, int, int, int, float*,
                                      cudaStream_t);
template int launch_pack_level<bf16>(const float*, const float*, const float*, int, int, int, int, int, bf16*,
                                     cudaStream_t);

template <typename T>
int launch_layernorm(const float* x, const float* gamma, const float* beta, const float* pos, int rows, int C,
                     float* y_f32, T* y_t, T* y_pos_t, cudaStream_t stream)
{
    OCC_CHECK(C == 256, "layernorm: embed_dims must be 256");
    layernorm256_kernel<T><<<ceil_div(rows, 8), 256, 0, stream>>>(x, gamma, beta, pos, rows, y_f32, y_t, y_pos_t);
    OCC_CUDA(cudaGetLastError());
    return 0;
}
template int launch_layernorm<float>(const float*, const float*, const float*, const float*, int, int, float*,
                                     float*, float*, cudaStream_t);
template int launch_layernorm<bf16>(const float*, const float*, const float*, const float*, int, int, float*, bf16*,
                                    bf16*, cudaStream_t);

template <typename T>
int launch_prepare_query(const float* bev_queries, const float* pos, int64_t n, float* q_f32, T* q_t, T* q_pos_t,
                         cudaStream_t stream)
{
    OCC_CHECK(n % 8 == 0, "prepare_query: size must be a multiple of 8");
    prepare_query_kernel<T><<<ceil_div(n / 8, 256), 256, 0, stream>>>(bev_queries, pos, n / 8, q_f32, q_t, q_pos_t);
    OCC_CUDA(cudaGetLastError());
    return 0;
}
template int launch_prepare_query<float>(const float*, const float*, int64_t, float*, float*, float*, cudaStream_t);
template int launch_prepare_query<bf16>(const float*, const float*, int64_t, float*, bf16*, bf16*, cudaStream_t);

int launch_bev_pos(const float* row_embed, const float* col_embed, int bev_h, int bev_w, int half, float* pos,
                   cudaStream_t stream)
{
    const int64_t n = (int64_t)bev_h * bev_w * 2 * half;
    bev_pos_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(row_embed, col_embed, bev_h, bev_w, half, pos);
    OCC_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
int launch_cast(const float* src, T* dst, int64_t n, cudaStream_t stream)
{
    cast_kernel<T><<<ceil_div(n, 256), 256, 0, stream>>>(src, dst, n);
    OCC_CUDA(cudaGetLastError());
    return 0;
}
template int launch_cast<float>(const float*, float*, int64_t, cudaStream_t);
template int launch_cast<bf16>(const float*, bf16*, int64_t, cudaStream_t);

}  // namespace occ

#include "gemm_tc.cuh"
namespace occ {
namespace {
__global__ void bf16_to_f32_kernel(const bf16* __restrict__ s, float* __restrict__ d, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = __bfloat162float(s[i]);
}
}  // namespace
int launch_bf16_to_f32(const bf16* src, float* dst, int64_t n, cudaStream_t stream)
{
    bf16_to_f32_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(src, dst, n);
    OCC_CUDA(cudaGetLastError());
    return 0;
}
}  // namespace occ
