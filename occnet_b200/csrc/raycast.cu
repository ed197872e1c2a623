// Ray-casting metric for sm_100a.
//   render_forward_kernel <- tools/ray_iou/lib/dvr/dvr.cu:69-319 ("test" phase) + host wrapper :329-388
//   ray_metric_kernel     <- datasets/ray_metrics.py process_one_sample :89-143 + calc_metrics :146-189
// The traversal arithmetic is double precision with the reference's comparison order, so voxel
// indices are bit-exact.  The reference walks the whole grid recording a path (52 KB of local
// memory per thread) and scans it afterwards; here the first sigma > 0.5 voxel is latched during
// the walk, which needs no path storage.  The fused kernel casts all T origins x M rays through
// the predicted and the ground-truth volume in one launch and reduces the 187 counters on device.
#include <float.h>

#include "common.cuh"
#include "kernels.cuh"

namespace occ {

namespace {

constexpr int MAX_STEP = 1000;

struct Hit { double dist; int x, y, z; bool any; };

// occupied(x,y,z) -> bool.  Returns first-hit (or exit) distance and voxel; any == false when the ray
// never enters the grid (the reference then leaves pred_dist = -1 and coord_index = 0).
template <typename Occ>
__device__ __forceinline__ Hit dda_first_hit(double xo, double yo, double zo, double xe, double ye, double ze,
                                             int vxsize, int vysize, int vzsize, double& gt_d, Occ occupied)
{
    int vx = (int)xo, vy = (int)yo, vz = (int)zo;
    const double rx = xe - xo, ry = ye - yo, rz = ze - zo;
    gt_d = sqrt(rx * rx + ry * ry + rz * rz);
    const double dx = rx / gt_d, dy = ry / gt_d, dz = rz / gt_d;
    const int stepX = (dx >= 0) ? 1 : -1, stepY = (dy >= 0) ? 1 : -1, stepZ = (dz >= 0) ? 1 : -1;
    const double nbx = vx + (stepX < 0 ? 0 : 1), nby = vy + (stepY < 0 ? 0 : 1), nbz = vz + (stepZ < 0 ? 0 : 1);
    double tMaxX = (dx != 0) ? (nbx - xo) / dx : DBL_MAX;
    double tMaxY = (dy != 0) ? (nby - yo) / dy : DBL_MAX;
    double tMaxZ = (dz != 0) ? (nbz - zo) / dz : DBL_MAX;
    const double tDeltaX = (dx != 0) ? stepX / dx : DBL_MAX;
    const double tDeltaY = (dy != 0) ? stepY / dy : DBL_MAX;
    const double tDeltaZ = (dz != 0) ? stepZ / dz : DBL_MAX;
    Hit h; h.dist = 0.0; h.x = h.y = h.z = 0; h.any = false;
    bool was_inside = false, hit = false;
    int step = 0;
    while (true) {
        const bool inside = (0 <= vx && vx < vxsize) && (0 <= vy && vy < vysize) && (0 <= vz && vz < vzsize);
        const int cx = vx, cy = vy, cz = vz;
        if (inside) was_inside = true;
        else if (was_inside) break;
        double d;
        if (tMaxX < tMaxY) {
            if (tMaxX < tMaxZ) { d = tMaxX; vx += stepX; tMaxX += tDeltaX; }
            else               { d = tMaxZ; vz += stepZ; tMaxZ += tDeltaZ; }
        } else {
            if (tMaxY < tMaxZ) { d = tMaxY; vy += stepY; tMaxY += tDeltaY; }
            else               { d = tMaxZ; vz += stepZ; tMaxZ += tDeltaZ; }
        }
        if (inside && !hit) {
            h.any = true; h.dist = d; h.x = cx; h.y = cy; h.z = cz;     // last in-grid voxel so far
            if (occupied(cx, cy, cz)) { hit = true; break; }             // later voxels cannot change the answer
        }
        ++step;
        if (step > MAX_STEP) break;
    }
    return h;
}

__global__ void render_forward_kernel(const float* __restrict__ sigma, const float* __restrict__ origin,
                                      const float* __restrict__ points, const float* __restrict__ tindex, int T,
                                      int vzsize, int vysize, int vxsize, int64_t M, float* __restrict__ pred_dist,
                                      float* __restrict__ gt_dist, float* __restrict__ coord_index)
{
    const int n = blockIdx.y;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M) return;
    const int64_t rc = (int64_t)n * M + c;
    pred_dist[rc] = -1.f; gt_dist[rc] = -1.f;
    coord_index[rc * 3] = 0.f; coord_index[rc * 3 + 1] = 0.f; coord_index[rc * 3 + 2] = 0.f;
    const float tf = tindex[rc];
    if (tf < 0) return;
    const int t = (int)tf;
    const int ts = (T == 1) ? 0 : t;
    const float* o = origin + ((int64_t)n * T + t) * 3;
    const float* p = points + rc * 3;
    const float* sg = sigma + ((int64_t)n * T + ts) * vzsize * vysize * vxsize;
    double gt_d;
    const Hit h = dda_first_hit(o[0], o[1], o[2], p[0], p[1], p[2], vxsize, vysize, vzsize, gt_d,
                                [&](int x, int y, int z) {
                                    return (double)sg[((int64_t)z * vysize + y) * vxsize + x] > 0.5;
                                });
    if (h.any) {
        pred_dist[rc] = (float)h.dist;
        gt_dist[rc] = (float)gt_d;
        coord_index[rc * 3] = (float)h.x; coord_index[rc * 3 + 1] = (float)h.y; coord_index[rc * 3 + 2] = (float)h.z;
    }
}

constexpr int NCLS = 17, FREE = 16, NFLOW = 8, NCNT = 11 * NCLS;
constexpr int GX = 200, GY = 200, GZ = 16;

// one thread = one (origin t, ray m); casts through pred and gt, updates the counters
__global__ void __launch_bounds__(128)
ray_metric_kernel(const uint8_t* __restrict__ sem_pred, const float* __restrict__ flow_pred,
                  const uint8_t* __restrict__ sem_gt, const float* __restrict__ flow_gt,
                  const void* __restrict__ origins, int origin_is_f64, int T, const float* __restrict__ rays,
                  int M, double* __restrict__ counters, float* __restrict__ pcd_pred, float* __restrict__ pcd_gt)
{
    __shared__ double s_cnt[NCNT];
    for (int i = threadIdx.x; i < NCNT; i += blockDim.x) s_cnt[i] = 0.0;
    __syncthreads();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (int64_t)T * M) {
        const int t = (int)(idx / M), m = (int)(idx % M);
        // voxel-unit origin / end point exactly as ray_metrics.py:102-112.  torch type promotion: with the
        // dataset's float64 origins (ego_pose_extractor.py:108-119) the arithmetic is double and rounded
        // once by `.float()`; with float32 origins every step is fp32.
        const float off[3] = {-40.f, -40.f, -1.f};
        float og[3], en[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (origin_is_f64) {
                const double o = reinterpret_cast<const double*>(origins)[t * 3 + k];
                const double e = (double)rays[m * 3 + k] + o;
                og[k] = (float)((o - (double)off[k]) / (double)0.4f);
                en[k] = (float)((e - (double)off[k]) / (double)0.4f);
            } else {
                const float o = reinterpret_cast<const float*>(origins)[t * 3 + k];
                const float e = __fadd_rn(rays[m * 3 + k], o);
                og[k] = __fdiv_rn(__fsub_rn(o, off[k]), 0.4f);
                en[k] = __fdiv_rn(__fsub_rn(e, off[k]), 0.4f);
            }
        }
        float row[2][4];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const uint8_t* sem = v ? sem_gt : sem_pred;
            const float* flow = v ? flow_gt : flow_pred;
            double gt_d;
            const Hit h = dda_first_hit(og[0], og[1], og[2], en[0], en[1], en[2], GX, GY, GZ, gt_d,
                                        [&](int x, int y, int z) { return sem[((int64_t)x * GY + y) * GZ + z] != FREE; });
            const float dist = (h.any ? (float)h.dist : -1.f) * 0.4f;
            const int64_t vi = ((int64_t)h.x * GY + h.y) * GZ + h.z;
            row[v][0] = (float)sem[vi]; row[v][1] = dist; row[v][2] = flow[vi * 2]; row[v][3] = flow[vi * 2 + 1];
        }
        if (pcd_pred) *reinterpret_cast<float4*>(pcd_pred + idx * 4) = make_float4(row[0][0], row[0][1], row[0][2], row[0][3]);
        if (pcd_gt) *reinterpret_cast<float4*>(pcd_gt + idx * 4) = make_float4(row[1][0], row[1][1], row[1][2], row[1][3]);
        const int cp = (int)row[0][0], cg = (int)row[1][0];
        if (cg != FREE) {                                              // ray_metrics.py:218-220
            if (cg < NCLS) atomicAdd(&s_cnt[cg], 1.0);
            if (cp < NCLS) atomicAdd(&s_cnt[NCLS + cp], 1.0);
            if (cg == cp && cg < NCLS) {
                const float l1 = fabsf(row[0][1] - row[1][1]);
                const float fx = row[1][2] - row[0][2], fy = row[1][3] - row[0][3];
                const float err = sqrtf(fx * fx + fy * fy);
                const float thr[3] = {1.f, 2.f, 4.f};
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (l1 < thr[j]) {
                        atomicAdd(&s_cnt[2 * NCLS + j * NCLS + cg], 1.0);
                        if (cg < NFLOW) {
                            atomicAdd(&s_cnt[5 * NCLS + j * NCLS + cg], (double)err);
                            atomicAdd(&s_cnt[8 * NCLS + j * NCLS + cg], 1.0);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NCNT; i += blockDim.x)
        if (s_cnt[i] != 0.0) atomicAdd(&counters[i], s_cnt[i]);
}

}  // namespace

int launch_render_forward(const float* sigma, const float* origin, const float* points, const float* tindex,
                          int N, int T, int Z, int Y, int X, int64_t M, float* pred_dist, float* gt_dist,
                          float* coord_index, cudaStream_t stream)
{
    if (N == 0 || M == 0) return 0;
    dim3 grid(ceil_div(M, 128), N);
    render_forward_kernel<<<grid, 128, 0, stream>>>(sigma, origin, points, tindex, T, Z, Y, X, M, pred_dist, gt_dist,
                                                    coord_index);
    OCC_CUDA(cudaGetLastError());
    return 0;
}

int launch_ray_metric(const uint8_t* sem_pred, const float* flow_pred, const uint8_t* sem_gt, const float* flow_gt,
                      const void* origins, int origin_is_f64, int T, const float* rays, int M, double* counters,
                      float* pcd_pred, float* pcd_gt, cudaStream_t stream)
{
    if (T == 0 || M == 0) return 0;
    ray_metric_kernel<<<ceil_div((int64_t)T * M, 128), 128, 0, stream>>>(sem_pred, flow_pred, sem_gt, flow_gt, origins,
                                                                        origin_is_f64, T, rays, M, counters, pcd_pred,
                                                                        pcd_gt);
    OCC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace occ
