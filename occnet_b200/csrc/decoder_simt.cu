// CUDA-core implementation of the voxel decoder and the occupancy/flow heads (fp32 parity
// configuration and bring-up path; the bf16 tensor-core path lives in conv3d_tc.cu).
//   TransformerOcc.decoder (use_3d):  transformer_occ.py:106-131, applied :305-308
//   predicter / flow_predicter:       transformer_occ.py:132-141, :318-319
//   get_occ (softmax -> argmax):      bevformer_occ_head.py:211-212
// Voxel tensors are channels-last [X][Y][Z][C]: the order of the reference's final
// `outputs.permute(0,4,3,2,1)` so that occ/flow come out in (X, Y, Z) order without a transpose.
#include "common.cuh"
#include "kernels.cuh"

namespace occ {

namespace {

template <typename T>
__global__ void __launch_bounds__(256)
bev_to_voxel_kernel(const float* __restrict__ bev, int bev_h, int bev_w, int Z, int mid, T* __restrict__ vox)
{
    __shared__ float rows[8][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * 8 + w;
    if (q >= bev_h * bev_w) return;
    float v[8];
    load8(bev + (int64_t)q * 256 + lane * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) rows[w][lane * 8 + i] = v[i];
    __syncwarp();
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int oi = lane * 8 + i, z = oi / mid, cm = oi % mid;   // out index z*mid + cm <- in index cm*Z + z
        o[i] = rows[w][cm * Z + z];
    }
    const int x = q % bev_w, y = q / bev_w;
    store8(vox + ((int64_t)x * bev_h + y) * 256 + lane * 8, o);
}

// The same lift straight from the T32 residual-stream layout of the tensor-core path (Z = mid = 16), bf16 out: one CTA = one
// 32-row block.  Thread (lane = row, warp = (z block of 4, cm block of 8)) reads the 8 float4 {cm, z0..z0+3} of its row -- 512
// contiguous bytes per warp instruction in T32 -- and owns a 4 (z) x 8 (cm) block of the output row; rows are staged in shared
// memory (528-byte pitch) and leave as whole 512-byte voxel columns.  Replaces t32_convert + bev_to_voxel when bev_embed itself
// is not requested.
__global__ void __launch_bounds__(256)
t32_to_voxel_kernel(const float* __restrict__ bev_t32, int bev_h, int bev_w, bf16* __restrict__ vox)
{
    __shared__ __align__(16) uint8_t rows[32 * 528];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int zb = warp & 3, cmb = warp >> 2;                   // z0 = 4 zb, cm0 = 8 cmb
    const int Nq = bev_h * bev_w;
    const int64_t R = blockIdx.x;
    const float4* blk = reinterpret_cast<const float4*>(bev_t32) + R * 8 * 8 * 32 + lane;
    float v[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = (cmb * 8 + k) * 16 + zb * 4;              // input channel of (cm, z0): cm * Z + z
        const float4 t = __ldg(blk + ((c >> 5) * 8 + ((c & 31) >> 2)) * 32);
        v[k][0] = t.x; v[k][1] = t.y; v[k][2] = t.z; v[k][3] = t.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {                               // output channels (z0 + i) * 16 + cm0 .. + 7
        const uint4 o = make_uint4(pack_bf16x2(v[0][i], v[1][i]), pack_bf16x2(v[2][i], v[3][i]), pack_bf16x2(v[4][i], v[5][i]),
                                   pack_bf16x2(v[6][i], v[7][i]));
        *reinterpret_cast<uint4*>(rows + lane * 528 + ((zb * 4 + i) * 16 + cmb * 8) * 2) = o;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {                               // warp w writes rows w, w + 8, ...: 32 lanes x 16 B = one voxel column
        const int r = warp + 8 * i;
        const int64_t q = R * 32 + r;
        if (q < Nq) {
            const int x = (int)(q % bev_w), y = (int)(q / bev_w);
            *reinterpret_cast<uint4*>(vox + ((int64_t)x * bev_h + y) * 256 + lane * 8) = *reinterpret_cast<const uint4*>(rows + r * 528 + lane * 16);
        }
    }
}

// one thread = one voxel x 32 output channels; block = 8 (y) x 16 (z) voxels at one x
template <typename T, int CIN>
__global__ void __launch_bounds__(128)
conv3d_simt_kernel(const T* __restrict__ in, const float* __restrict__ wfold, const float* __restrict__ bfold,
                   int X, int Y, int Z, T* __restrict__ out)
{
    extern __shared__ __align__(16) float wsm[];               // [27][CIN][32]
    for (int i = threadIdx.x; i < 27 * CIN * 32 / 4; i += blockDim.x)
        reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<const float4*>(wfold) + i);
    __syncthreads();
    const int z = threadIdx.x % 16, yl = threadIdx.x / 16;
    const int x = blockIdx.x, y = blockIdx.y * 8 + yl;
    if (y >= Y || z >= Z) return;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = bfold[i];
    for (int dx = 0; dx < 3; ++dx) {
        const int xx = x + dx - 1;
        if (xx < 0 || xx >= X) continue;
        for (int dy = 0; dy < 3; ++dy) {
            const int yy = y + dy - 1;
            if (yy < 0 || yy >= Y) continue;
#pragma unroll
            for (int dz = 0; dz < 3; ++dz) {
                const int zz = z + dz - 1;
                if (zz < 0 || zz >= Z) continue;
                const T* ip = in + (((int64_t)xx * Y + yy) * Z + zz) * CIN;
                const float* wp = wsm + ((dz * 3 + dy) * 3 + dx) * CIN * 32;
#pragma unroll
                for (int c8 = 0; c8 < CIN / 8; ++c8) {
                    float v[8];
                    load8(ip + c8 * 8, v);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4* w4 = reinterpret_cast<const float4*>(wp + (c8 * 8 + k) * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 w = w4[j];
                            acc[4 * j + 0] = fmaf(v[k], w.x, acc[4 * j + 0]);
                            acc[4 * j + 1] = fmaf(v[k], w.y, acc[4 * j + 1]);
                            acc[4 * j + 2] = fmaf(v[k], w.z, acc[4 * j + 2]);
                            acc[4 * j + 3] = fmaf(v[k], w.w, acc[4 * j + 3]);
                        }
                    }
                }
            }
        }
    }
    T* op = out + (((int64_t)x * Y + y) * Z + z) * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaxf(acc[j * 8 + k], 0.f);
        store8(op + j * 8, o);
    }
}

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }

template <typename T>
__global__ void __launch_bounds__(128)
occ_head_kernel(const T* __restrict__ vox, HeadWeights hw, int64_t nvox, float* __restrict__ occ_logits,
                float* __restrict__ flow, uint8_t* __restrict__ cls_u8, int64_t* __restrict__ cls_i64)
{
    __shared__ __align__(16) float s_w1[64 * 32], s_b1[64], s_w2[32 * 64], s_b2[32];
    __shared__ __align__(16) float s_fw1[64 * 32], s_fb1[64], s_fw2[2 * 64], s_fb2[2];
    const int ncls = hw.ncls;
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) { s_w1[i] = hw.w1[i]; s_fw1[i] = hw.fw1[i]; }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) { s_b1[i] = hw.b1[i]; s_fb1[i] = hw.fb1[i]; }
    for (int i = threadIdx.x; i < ncls * 64; i += blockDim.x) s_w2[i] = hw.w2[i];
    for (int i = threadIdx.x; i < ncls; i += blockDim.x) s_b2[i] = hw.b2[i];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_fw2[i] = hw.fw2[i];
    if (threadIdx.x < 2) s_fb2[threadIdx.x] = hw.fb2[threadIdx.x];
    __syncthreads();
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    float f[32];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float t[8];
        load8(vox + v * 32 + j * 8, t);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[j * 8 + k] = t[k];
    }
    float h[64];
    // semantic head: Linear(32,64) -> Softplus -> Linear(64, ncls)
#pragma unroll 4
    for (int o = 0; o < 64; ++o) {
        float a = s_b1[o];
        const float4* w4 = reinterpret_cast<const float4*>(s_w1 + o * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 w = w4[k];
            a = fmaf(f[4 * k], w.x, a); a = fmaf(f[4 * k + 1], w.y, a);
            a = fmaf(f[4 * k + 2], w.z, a); a = fmaf(f[4 * k + 3], w.w, a);
        }
        h[o] = softplus_f(a);
    }
    float best = -INFINITY;
    int arg = 0;
    for (int c = 0; c < ncls; ++c) {
        float a = s_b2[c];
        const float4* w4 = reinterpret_cast<const float4*>(s_w2 + c * 64);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 w = w4[k];
            a = fmaf(h[4 * k], w.x, a); a = fmaf(h[4 * k + 1], w.y, a);
            a = fmaf(h[4 * k + 2], w.z, a); a = fmaf(h[4 * k + 3], w.w, a);
        }
        if (occ_logits) occ_logits[v * ncls + c] = a;
        if (a > best) { best = a; arg = c; }
    }
    if (cls_u8) cls_u8[v] = (uint8_t)arg;
    if (cls_i64) cls_i64[v] = arg;
    // flow head: Linear(32,64) -> ReLU -> Linear(64,2)
#pragma unroll 4
    for (int o = 0; o < 64; ++o) {
        float a = s_fb1[o];
        const float4* w4 = reinterpret_cast<const float4*>(s_fw1 + o * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 w = w4[k];
            a = fmaf(f[4 * k], w.x, a); a = fmaf(f[4 * k + 1], w.y, a);
            a = fmaf(f[4 * k + 2], w.z, a); a = fmaf(f[4 * k + 3], w.w, a);
        }
        h[o] = fmaxf(a, 0.f);
    }
    float fl[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        float a = s_fb2[c];
#pragma unroll
        for (int k = 0; k < 64; ++k) a = fmaf(h[k], s_fw2[c * 64 + k], a);
        fl[c] = a;
    }
    if (flow) *reinterpret_cast<float2*>(flow + v * 2) = make_float2(fl[0], fl[1]);
}

}  // namespace

template <typename T>
int launch_bev_to_voxel(const float* bev, int bev_h, int bev_w, int Z, int mid, T* vox, cudaStream_t stream)
{
    OCC_CHECK(Z * mid == 256, "bev_to_voxel: embed_dims must be 256");
    bev_to_voxel_kernel<T><<<ceil_div((int64_t)bev_h * bev_w, 8), 256, 0, stream>>>(bev, bev_h, bev_w, Z, mid, vox);
    OCC_CUDA(cudaGetLastError());
    return 0;
}
int launch_t32_to_voxel(const float* bev_t32, int bev_h, int bev_w, bf16* vox, cudaStream_t stream)
{
    t32_to_voxel_kernel<<<ceil_div((int64_t)bev_h * bev_w, 32), 256, 0, stream>>>(bev_t32, bev_h, bev_w, vox);
    OCC_CUDA(cudaGetLastError());
    return 0;
}
template int launch_bev_to_voxel<float>(const float*, int, int, int, int, float*, cudaStream_t);
template int launch_bev_to_voxel<bf16>(const float*, int, int, int, int, bf16*, cudaStream_t);

template <typename T>
int launch_conv3d_simt(const T* in, const float* wfold, const float* bfold, int X, int Y, int Z, int Cin, T* out,
                       cudaStream_t stream)
{
    OCC_CHECK(Z <= 16 && (Cin == 16 || Cin == 32), "conv3d_simt: Z <= 16 and Cin in {16, 32}");
    dim3 grid(X, ceil_div(Y, 8));
    const size_t smem = (size_t)27 * Cin * 32 * sizeof(float);
    if (Cin == 16) {
        OCC_CUDA(cudaFuncSetAttribute(conv3d_simt_kernel<T, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv3d_simt_kernel<T, 16><<<grid, 128, smem, stream>>>(in, wfold, bfold, X, Y, Z, out);
    } else {
        OCC_CUDA(cudaFuncSetAttribute(conv3d_simt_kernel<T, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv3d_simt_kernel<T, 32><<<grid, 128, smem, stream>>>(in, wfold, bfold, X, Y, Z, out);
    }
    OCC_CUDA(cudaGetLastError());
    return 0;
}
template int launch_conv3d_simt<float>(const float*, const float*, const float*, int, int, int, int, float*,
                                       cudaStream_t);
template int launch_conv3d_simt<bf16>(const bf16*, const float*, const float*, int, int, int, int, bf16*,
                                      cudaStream_t);

template <typename T>
int launch_occ_head(const T* vox, HeadWeights hw, int64_t nvox, float* occ_logits, float* flow, uint8_t* cls_u8,
                    int64_t* cls_i64, cudaStream_t stream)
{
    OCC_CHECK(hw.ncls <= 32, "occ_head: at most 32 classes");
    occ_head_kernel<T><<<ceil_div(nvox, 128), 128, 0, stream>>>(vox, hw, nvox, occ_logits, flow, cls_u8, cls_i64);
    OCC_CUDA(cudaGetLastError());
    return 0;
}
template int launch_occ_head<float>(const float*, HeadWeights, int64_t, float*, float*, uint8_t*, int64_t*,
                                    cudaStream_t);
template int launch_occ_head<bf16>(const bf16*, HeadWeights, int64_t, float*, float*, uint8_t*, int64_t*,
                                   cudaStream_t);

}  // namespace occ
