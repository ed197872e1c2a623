// tcgen05 / TMEM / TMA bf16 GEMM for the dense layers of the encoder (sm_100a):
//     C[M,N] = act(A[M,K] . W[N,K]^T + bias[N]) (+ residual[M,N]),   fp32 accumulation in TMEM.
// Every nn.Linear on the path has this shape (reference: temporal_self_attention.py:99-104,
// spatial_cross_attention.py:67,245-249, mmcv FFN) with K in {256,512}, N in {192,256,512,768}.
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0   TMA producer : A tile 128x64 + W tile BNx64 (bf16, 128B swizzle) per k-block, 4-stage mbarrier ring
//   warp 1   MMA issuer   : tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16 x4 per stage;
//                           accumulators double-buffered in TMEM (2 x 256 columns)
//   warps 2-5 epilogue    : tcgen05.ld 32x32b -> bias / ReLU / residual -> global store (fp32 or bf16),
//                           overlapping the next tile's main loop
// The A operand may be the concatenation of two matrices along K (TSA's cat([value, query+pos]),
// temporal_self_attention.py:197) -- two tensor maps, no materialised concat.
#include <mutex>
#include <map>
#include <tuple>

#include "gemm_tc.cuh"
#include "tc_common.cuh"

namespace occ {

// ------------------------------------------------------------------------------------------------ host helpers
PFN_encodeTiled get_encode_tiled()
{
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

int make_tensor_map_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, int swizzle_bytes)
{
    PFN_encodeTiled enc = get_encode_tiled();
    OCC_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                           bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    OCC_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return 0;
}

namespace {

constexpr int BLOCK_M = 128, BLOCK_K = 64, STAGES = 4, A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_THREADS = 192;

template <typename TC>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmW, const float* __restrict__ bias,
               const float* __restrict__ residual, TC* __restrict__ C, int M, int N, int BN, int nk, int nk1, int act)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_bytes = A_TILE_BYTES + BN * BLOCK_K * 2;
    const uint32_t bar_base = smem_base + STAGES * stage_bytes;
    auto full_bar = [&](int s) { return bar_base + s * 8; };
    auto empty_bar = [&](int s) { return bar_base + (STAGES + s) * 8; };
    auto tfull_bar = [&](int s) { return bar_base + (2 * STAGES + s) * 8; };
    auto tempty_bar = [&](int s) { return bar_base + (2 * STAGES + 2 + s) * 8; };
    const uint32_t tmem_slot = bar_base + (2 * STAGES + 4) * 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M, n_tiles = N / BN;
    const int num_tiles = m_tiles * n_tiles;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmA); tc::tma_prefetch_desc(&tmA2); tc::tma_prefetch_desc(&tmW);
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(full_bar(s), 1); tc::mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(tfull_bar(s), 1); tc::mbar_init(tempty_bar(s), 128); }
        tc::mbar_fence_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
                for (int kb = 0; kb < nk; ++kb) {
                    tc::mbar_wait(empty_bar(s), ph ^ 1);
                    tc::mbar_arrive_expect_tx(full_bar(s), stage_bytes);
                    const uint32_t a_dst = smem_base + s * stage_bytes;
                    if (kb < nk1) tc::tma_load_2d(a_dst, &tmA, full_bar(s), kb * BLOCK_K, m_blk * BLOCK_M);
                    else          tc::tma_load_2d(a_dst, &tmA2, full_bar(s), (kb - nk1) * BLOCK_K, m_blk * BLOCK_M);
                    tc::tma_load_2d(a_dst + A_TILE_BYTES, &tmW, full_bar(s), kb * BLOCK_K, n_blk * BN);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = tc::make_idesc_bf16(BLOCK_M, BN);
        int s = 0; uint32_t ph = 0; int as = 0; uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            if (lane == 0) {
                tc::mbar_wait(tempty_bar(as), aph ^ 1);
                tc::tc_fence_after();
            }
            __syncwarp();
            for (int kb = 0; kb < nk; ++kb) {
                if (lane == 0) {
                    tc::mbar_wait(full_bar(s), ph);
                    tc::tc_fence_after();
                    const uint32_t a_addr = smem_base + s * stage_bytes;
                    const uint64_t da = tc::make_smem_desc(a_addr, 128);
                    const uint64_t db = tc::make_smem_desc(a_addr + A_TILE_BYTES, 128);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k)      // +32 bytes (>>4 = 2) per UMMA_K step inside the swizzle atom
                        tc::umma_bf16(tmem_base + as * 256, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    tc::umma_commit(empty_bar(s));               // frees the smem stage when these MMAs retire
                    if (kb == nk - 1) tc::umma_commit(tfull_bar(as));
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            if (++as == 2) { as = 0; aph ^= 1; }
        }
    } else {
        const int quarter = warp & 3;                            // TMEM lane quarter this warp may access
        int as = 0; uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
            tc::mbar_wait(tfull_bar(as), aph);
            tc::tc_fence_after();
            const int row = m_blk * BLOCK_M + quarter * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * 256;
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tc::tmem_ld32(taddr + c0, r);
                tc::tmem_ld_wait();
                if (row < M) {
                    const int col = n_blk * BN + c0;
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                    if (bias) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col + i));
                            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                        }
                    }
                    if (act == ACT_RELU) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    if (residual) {
                        const float4* rp = reinterpret_cast<const float4*>(residual + (size_t)row * N + col);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 b = __ldg(rp + i);
                            v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
                        }
                    }
                    if constexpr (sizeof(TC) == 4) {
                        float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(C) + (size_t)row * N + col);
#pragma unroll
                        for (int i = 0; i < 8; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    } else {
                        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(C) + (size_t)row * N + col);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint4 u;
                            u.x = pack_bf16x2(v[8 * i], v[8 * i + 1]); u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
                            u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
                            op[i] = u;
                        }
                    }
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(tempty_bar(as));
            if (++as == 2) { as = 0; aph ^= 1; }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

int pick_bn(int N)
{
    if (N <= 256) return N;
    if (N % 256 == 0) return 256;
    if (N % 192 == 0) return 192;
    if (N % 128 == 0) return 128;
    return 0;
}

struct MapKey {
    const void* p; uint64_t d0, d1; uint32_t b0, b1;
    bool operator<(const MapKey& o) const { return std::tie(p, d0, d1, b0, b1) < std::tie(o.p, o.d0, o.d1, o.b0, o.b1); }
};

int cached_map_2d(const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows, CUtensorMap* out)
{
    static std::map<MapKey, CUtensorMap> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    const MapKey k{base, inner, rows, box_inner, box_rows};
    auto it = cache.find(k);
    if (it == cache.end()) {
        CUtensorMap m;
        const uint64_t dims[2] = {inner, rows}, strides[1] = {inner * 2};
        const uint32_t box[2] = {box_inner, box_rows};
        if (make_tensor_map_bf16(&m, base, 2, dims, strides, box, 128)) return 1;
        if (cache.size() > 4096) cache.clear();
        it = cache.emplace(k, m).first;
    }
    *out = it->second;
    return 0;
}

}  // namespace

bool gemm_tc_supported(int M, int N, int K, int K1)
{
    const int bn = pick_bn(N);
    return M > 0 && bn >= 16 && bn % 16 == 0 && K % 64 == 0 && K1 % 64 == 0 && K1 > 0 && K1 <= K;
}

template <typename TC>
int gemm_tc(const bf16* A, const bf16* A2, int K1, const bf16* W, const float* bias, const float* residual, TC* C,
            int M, int N, int K, int act, cudaStream_t stream)
{
    if (A2 == nullptr) K1 = K;
    OCC_CHECK(gemm_tc_supported(M, N, K, K1), "gemm_tc: unsupported shape");
    const int BN = pick_bn(N);
    CUtensorMap tmA, tmA2, tmW;
    if (cached_map_2d(A, (uint64_t)K1, (uint64_t)M, BLOCK_K, BLOCK_M, &tmA)) return 1;
    if (A2) { if (cached_map_2d(A2, (uint64_t)(K - K1), (uint64_t)M, BLOCK_K, BLOCK_M, &tmA2)) return 1; }
    else tmA2 = tmA;
    if (cached_map_2d(W, (uint64_t)K, (uint64_t)N, BLOCK_K, (uint32_t)BN, &tmW)) return 1;
    const int stage_bytes = A_TILE_BYTES + BN * BLOCK_K * 2;
    const int smem = 1024 + STAGES * stage_bytes + 256;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        OCC_CUDA(cudaGetDevice(&dev));
        OCC_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    OCC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M, tiles = m_tiles * (N / BN);
    const int grid = tiles < num_sms ? tiles : num_sms;
    gemm_tc_kernel<TC><<<grid, NUM_THREADS, smem, stream>>>(tmA, tmA2, tmW, bias, residual, C, M, N, BN, K / BLOCK_K,
                                                           K1 / BLOCK_K, act);
    OCC_CUDA(cudaGetLastError());
    return 0;
}

template int gemm_tc<float>(const bf16*, const bf16*, int, const bf16*, const float*, const float*, float*, int, int,
                            int, int, cudaStream_t);
template int gemm_tc<bf16>(const bf16*, const bf16*, int, const bf16*, const float*, const float*, bf16*, int, int, int,
                           int, cudaStream_t);

}  // namespace occ
