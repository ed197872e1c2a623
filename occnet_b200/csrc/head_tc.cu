// Occupancy + flow heads on the tensor cores (sm_100a).  Reference: TransformerOcc.predicter /
// flow_predicter (transformer_occ.py:132-141, applied :318-319) and BEVFormerOccHead.get_occ
// (bevformer_occ_head.py:211-212; argmax of the logits == argmax of their softmax).
//
// Per 128-voxel tile, a chain of two tcgen05 GEMMs with the activations kept on chip:
//   GEMM1  [128 x 32] . [32 x 128]   hidden = [ predicter.0 | flow_predicter.0 ]        -> TMEM cols [0,128)
//   epilogue A: +bias, Softplus (cols 0-63) / ReLU (cols 64-127), bf16, written back to shared memory in the
//               128B-swizzled K-major layout the tensor core reads (generic-proxy stores + fence.proxy.async)
//   GEMM2  [128 x 128] . [128 x 32]  block-diagonal [ predicter.2 (17) ; flow_predicter.2 (2) ; 0-pad ]
//                                                                                        -> TMEM cols [128,160)
//   epilogue B: +bias, argmax over the 17 classes, coalesced stores of logits / class / flow via shared memory
// Two tiles are in flight (TMEM and the hidden buffer are double-buffered): GEMM1 of tile i+1 is issued
// before GEMM2 of tile i, and the 8 epilogue warps run epilogue A of tile i+1 BEFORE epilogue B of tile i, so the
// MUFU-heavy activation pass never waits for GEMM2.  Warp (quarter q, half h) owns TMEM lanes 32q..32q+31 and the
// hidden columns [32h, 32h+32) (Softplus) + [64+32h, 64+32h+32) (ReLU); the half-0 warps also run epilogue B.
#include "common.cuh"
#include "conv3d_tc.cuh"
#include "tc_common.cuh"

namespace occ {

namespace {

constexpr int A_STAGES = 4, BLOCK_M = 128, HID = 128, NOUT = 32;
constexpr int A_BYTES = BLOCK_M * 64;                 // 128 rows x 32 bf16
constexpr int W1_BYTES = HID * 64;                    // 128 rows x 32 bf16
constexpr int W2_CHUNK_BYTES = NOUT * 128;            // 32 rows x 64 bf16
constexpr int H_CHUNK_BYTES = BLOCK_M * 128;          // 128 rows x 64 bf16
constexpr int NUM_THREADS = 320;                     // TMA warp, MMA warp, 8 epilogue warps
constexpr int MAX_CLS = 19;                           // 17 classes + 2 flow channels live in the 32 GEMM2 columns

// Softplus of TWO hidden units with ONE MUFU op each: softplus(x) = max(x, 0) + log1p(t), t = exp(-|x|) in (0, 1], and
// log1p(t) = t * P(t) with a degree-5 minimax polynomial (|P - log1p(t)/t| <= 7.2e-6 on [0,1], so the result is within 7.2e-6
// RELATIVE of softplus everywhere -- the hidden activations are then rounded to bf16, 2^-9).  The polynomial runs on the packed
// fp32x2 FMA of sm_100 (fma.rn.f32x2: two lanes per issue slot).  The previous form log(1 + exp(x)) needed ex2 AND lg2: the kernel
// issued 820 M MUFU ops per frame and was bound by that pipe (profiles/README.md).
__device__ __forceinline__ unsigned long long pack2(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ void softplus2(float x0, float x1, float& y0, float& y1)
{
    constexpr float C0 = 0.9999929070472717f, C1 = -0.4994262754917145f, C2 = 0.32572421431541443f, C3 = -0.211494579911232f,
                    C4 = 0.10287206619977951f, C5 = -0.024528255686163902f;
    float t0, t1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(-fabsf(x0) * 1.4426950408889634f));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(-fabsf(x1) * 1.4426950408889634f));
    const unsigned long long t = pack2(t0, t1);
    unsigned long long p = fma2(pack2(C5, C5), t, pack2(C4, C4));
    p = fma2(p, t, pack2(C3, C3));
    p = fma2(p, t, pack2(C2, C2));
    p = fma2(p, t, pack2(C1, C1));
    p = fma2(p, t, pack2(C0, C0));
    const unsigned long long y = fma2(t, p, pack2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(y0), "=f"(y1) : "l"(y));
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
head_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2, const float* __restrict__ b1cat,
               const float* __restrict__ b2cat, int ncls, int64_t nvox, float* __restrict__ occ_logits,
               float* __restrict__ flow, uint8_t* __restrict__ cls_u8, int64_t* __restrict__ cls_i64)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - tc::smem_u32(smem_raw));
    const uint32_t w1_base = smem_base;                                  // 8 KB
    const uint32_t w2_base = w1_base + W1_BYTES;                         // 2 x 4 KB
    const uint32_t a_base = w2_base + 2 * W2_CHUNK_BYTES;                // 4 x 8 KB
    const uint32_t h_base = a_base + A_STAGES * A_BYTES;                 // 2 stages x 2 chunks x 16 KB
    const uint32_t stage_out = h_base + 4 * H_CHUNK_BYTES;               // 4 warps x 32 x 19 floats
    const uint32_t bias_base = stage_out + 4 * 32 * MAX_CLS * 4;         // 128 + 32 floats: b1cat | b2cat
    const uint32_t bar_base = bias_base + (HID + NOUT) * 4 + 64;
    auto a_full = [&](int s) { return bar_base + s * 8; };
    auto a_empty = [&](int s) { return bar_base + (A_STAGES + s) * 8; };
    auto h1_full = [&](int s) { return bar_base + (2 * A_STAGES + s) * 8; };      // GEMM1 done
    auto h_ready = [&](int s) { return bar_base + (2 * A_STAGES + 2 + s) * 8; };  // hidden tile written
    auto l_full = [&](int s) { return bar_base + (2 * A_STAGES + 4 + s) * 8; };   // GEMM2 done
    auto t_empty = [&](int s) { return bar_base + (2 * A_STAGES + 6 + s) * 8; };  // TMEM stage drained
    const uint32_t w_bar = bar_base + (2 * A_STAGES + 8) * 8;
    const uint32_t tmem_slot = bar_base + (2 * A_STAGES + 9) * 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = (int)((nvox + BLOCK_M - 1) / BLOCK_M);

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmA); tc::tma_prefetch_desc(&tmW1); tc::tma_prefetch_desc(&tmW2);
        for (int s = 0; s < A_STAGES; ++s) { tc::mbar_init(a_full(s), 1); tc::mbar_init(a_empty(s), 1); }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(h1_full(s), 1); tc::mbar_init(h_ready(s), 256);
            tc::mbar_init(l_full(s), 1); tc::mbar_init(t_empty(s), 128);
        }
        tc::mbar_init(w_bar, 1);
        tc::mbar_fence_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
    float* const sbias = reinterpret_cast<float*>(smem_gen + (bias_base - smem_base));
    if (threadIdx.x >= 64 && threadIdx.x < 64 + HID + NOUT) {
        const int t = threadIdx.x - 64;
        sbias[t] = t < HID ? __ldg(b1cat + t) : (t - HID < ncls + 2 ? __ldg(b2cat + t - HID) : 0.f);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(w_bar, W1_BYTES + 2 * W2_CHUNK_BYTES);
            tc::tma_load_2d(w1_base, &tmW1, w_bar, 0, 0);
            tc::tma_load_2d(w2_base, &tmW2, w_bar, 0, 0);
            tc::tma_load_2d(w2_base + W2_CHUNK_BYTES, &tmW2, w_bar, 64, 0);
            int s = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                tc::mbar_wait(a_empty(s), ph ^ 1);
                tc::mbar_arrive_expect_tx(a_full(s), A_BYTES);
                tc::tma_load_2d(a_base + s * A_BYTES, &tmA, a_full(s), 0, tile * BLOCK_M);
                if (++s == A_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc1 = tc::make_idesc_bf16(BLOCK_M, HID), idesc2 = tc::make_idesc_bf16(BLOCK_M, NOUT);
            tc::mbar_wait(w_bar, 0);
            tc::tc_fence_after();
            int s = 0; uint32_t ph = 0;
            int n_my = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) ++n_my;
            for (int i = 0; i <= n_my; ++i) {
                if (i < n_my) {                                           // GEMM1 of tile i
                    const int as = i & 1; const uint32_t aph = (i >> 1) & 1;
                    tc::mbar_wait(t_empty(as), aph ^ 1);
                    tc::mbar_wait(a_full(s), ph);
                    tc::tc_fence_after();
                    const uint64_t da = tc::make_smem_desc(a_base + s * A_BYTES, 64);
                    const uint64_t db = tc::make_smem_desc(w1_base, 64);
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        tc::umma_bf16(tmem_base + as * 256, da + 2 * k, db + 2 * k, idesc1, k != 0);
                    tc::umma_commit(a_empty(s));
                    tc::umma_commit(h1_full(as));
                    if (++s == A_STAGES) { s = 0; ph ^= 1; }
                }
                if (i >= 1) {                                             // GEMM2 of tile i-1
                    const int j = i - 1, as = j & 1; const uint32_t aph = (j >> 1) & 1;
                    tc::mbar_wait(h_ready(as), aph);
                    tc::tc_fence_after();
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const uint64_t da = tc::make_smem_desc(h_base + (as * 2 + c) * H_CHUNK_BYTES, 128);
                        const uint64_t db = tc::make_smem_desc(w2_base + c * W2_CHUNK_BYTES, 128);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc::umma_bf16(tmem_base + as * 256 + 128, da + 2 * k, db + 2 * k, idesc2, (c | k) != 0);
                    }
                    tc::umma_commit(l_full(as));
                }
            }
        }
    } else {
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        float* sout = reinterpret_cast<float*>(smem_gen + (stage_out - smem_base)) + quarter * 32 * MAX_CLS;
        int n_my = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) ++n_my;
        // ---- epilogue A of my tile #i: hidden activations -> swizzled bf16 A operand of GEMM2 in shared memory
        auto epi_a = [&](int i) {
            const int as = i & 1; const uint32_t aph = (i >> 1) & 1;
            const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * 256;
            tc::mbar_wait(h1_full(as), aph);
            tc::tc_fence_after();
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                const int c0 = part * 64 + half * 32;                     // part 0: Softplus columns, part 1: ReLU columns
                uint32_t r[32];
                tc::tmem_ld32(tbase + c0, r);
                tc::tmem_ld_wait();
                const uint32_t chunk_base = h_base + (as * 2 + part) * H_CHUNK_BYTES + row * 128;
#pragma unroll
                for (int p = 0; p < 4; ++p) {                             // 4 x 16 bytes = 8 hidden units each
                    float v[8];
                    const float4 b0 = reinterpret_cast<const float4*>(sbias + c0)[2 * p];
                    const float4 b1 = reinterpret_cast<const float4*>(sbias + c0)[2 * p + 1];
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int k = 0; k < 8; k += 2) {
                        const float x0 = __uint_as_float(r[p * 8 + k]) + bb[k], x1 = __uint_as_float(r[p * 8 + k + 1]) + bb[k + 1];
                        if (part == 0) softplus2(x0, x1, v[k], v[k + 1]);
                        else { v[k] = fmaxf(x0, 0.f); v[k + 1] = fmaxf(x1, 0.f); }
                    }
                    const int piece = half * 4 + p;                       // 16-byte piece index inside the 128 B row
                    const uint32_t addr = chunk_base + ((piece ^ (row & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(v[0], v[1])),
                                 "r"(pack_bf16x2(v[2], v[3])), "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7]))
                                 : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to UMMA
            tc::tc_fence_before();
            tc::mbar_arrive(h_ready(as));
        };
        // ---- epilogue B of my tile #i (half-0 warps): logits / class / flow
        auto epi_b = [&](int i, int tile) {
            const int as = i & 1; const uint32_t aph = (i >> 1) & 1;
            const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * 256;
            tc::mbar_wait(l_full(as), aph);
            tc::tc_fence_after();
            uint32_t r[32];
            tc::tmem_ld32(tbase + 128, r);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            tc::mbar_arrive(t_empty(as));
            const int64_t v0 = (int64_t)tile * BLOCK_M + quarter * 32;    // first voxel of this warp
            float best = -INFINITY, f0 = 0.f, f1 = 0.f; int arg = 0;
#pragma unroll
            for (int c = 0; c < MAX_CLS; ++c) {                           // static register indices (no local memory)
                const float a = __uint_as_float(r[c]) + sbias[HID + c];
                if (c < ncls) {
                    sout[lane * MAX_CLS + c] = a;
                    if (a > best) { best = a; arg = c; }
                } else if (c == ncls) f0 = a;
                else if (c == ncls + 1) f1 = a;
            }
            const int64_t v = v0 + lane;
            if (v < nvox) {
                if (cls_u8) cls_u8[v] = (uint8_t)arg;
                if (cls_i64) cls_i64[v] = arg;
                if (flow) *reinterpret_cast<float2*>(flow + v * 2) = make_float2(f0, f1);
            }
            if (occ_logits) {
                __syncwarp();
                const int64_t nvalid = (nvox - v0 < 32 ? nvox - v0 : 32);
                float* dst = occ_logits + v0 * ncls;
                for (int j = lane; j < nvalid * ncls; j += 32) dst[j] = sout[(j / ncls) * MAX_CLS + (j % ncls)];
                __syncwarp();
            }
        };
        if (n_my > 0) epi_a(0);
        for (int i = 0; i < n_my; ++i) {
            if (i + 1 < n_my) epi_a(i + 1);
            if (half == 0) epi_b(i, blockIdx.x + i * (int)gridDim.x);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

int launch_occ_head_tc(const bf16* vox, const bf16* w1cat, const bf16* w2cat, const float* b1cat, const float* b2cat,
                       int ncls, int64_t nvox, float* occ_logits, float* flow, uint8_t* cls_u8, int64_t* cls_i64,
                       cudaStream_t stream)
{
    OCC_CHECK(ncls + 2 <= NOUT && ncls + 2 <= MAX_CLS, "occ_head_tc: at most 17 classes");
    CUtensorMap tmA, tmW1, tmW2;
    {
        const uint64_t dims[2] = {32, (uint64_t)nvox}, strides[1] = {64};
        const uint32_t box[2] = {32, BLOCK_M};
        if (make_tensor_map_bf16(&tmA, vox, 2, dims, strides, box, 64)) return 1;
    }
    {
        const uint64_t dims[2] = {32, HID}, strides[1] = {64};
        const uint32_t box[2] = {32, HID};
        if (make_tensor_map_bf16(&tmW1, w1cat, 2, dims, strides, box, 64)) return 1;
    }
    {
        const uint64_t dims[2] = {HID, NOUT}, strides[1] = {HID * 2};
        const uint32_t box[2] = {64, NOUT};
        if (make_tensor_map_bf16(&tmW2, w2cat, 2, dims, strides, box, 128)) return 1;
    }
    const int smem = 1024 + W1_BYTES + 2 * W2_CHUNK_BYTES + A_STAGES * A_BYTES + 4 * H_CHUNK_BYTES +
                     4 * 32 * MAX_CLS * 4 + (HID + NOUT) * 4 + 64 + 256;
    const int num_sms = sm_count_current_device();
    const int tiles = (int)((nvox + BLOCK_M - 1) / BLOCK_M);
    const int grid = tiles < num_sms ? tiles : num_sms;
    OCC_CUDA(cudaFuncSetAttribute(head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    head_tc_kernel<<<grid, NUM_THREADS, smem, stream>>>(tmA, tmW1, tmW2, b1cat, b2cat, ncls, nvox, occ_logits, flow,
                                                       cls_u8, cls_i64);
    OCC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace occ
