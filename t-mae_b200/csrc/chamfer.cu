// Masked-voxel reconstruction target + Chamfer loss.
//
// Replaces sst_ops_utils.group_inner_inds -> group_inner_inds_kernel / repeat_group_idx_kernel
// (pcdet/ops/sst_ops/sst_ops_utils.py:15-27, src/sst_ops_gpu.cu:22-39), the (M,64) index gather +
// voxel-centre subtraction (SiamWCA_MAE.py:132-139, common_utils.py:130-145) and
// pytorch3d.loss.chamfer_distance (SiamWCA_MAE.py:163; two generic KNN launches).
//
// One warp owns one pillar: its <= 64 ground-truth points come straight out of the voxeliser's
// per-voxel CSR (ascending point order = the canonical "first K arrivals"), the 16 predicted
// points live in registers, all 1024 pair distances are formed in-warp, and pillars with weight 0
// are skipped.  No atomics on the slot assignment, no .max().item() sync, no (M,64) index tensor.
#include "common.cuh"

namespace tmae {

struct GtArgs {
  const float* pts; int stride;          // kept points (n, stride): xyz at columns 1..3
  const int* offset; const int* order;   // CSR
  const int64_t* vcoords;                // (M,4) [b,z,y,x]
  float lo[3], vs[3];
  int K;
};

__device__ __forceinline__ void voxel_center(const GtArgs& g, int64_t v, float c[3]) {
  const int64_t* vc = g.vcoords + v * 4;
  // (coord + 0.5) * voxel_size + range_lo, separate roundings (common_utils.py:141-144)
  c[0] = __fadd_rn(__fmul_rn(__fadd_rn((float)vc[3], 0.5f), g.vs[0]), g.lo[0]);
  c[1] = __fadd_rn(__fmul_rn(__fadd_rn((float)vc[2], 0.5f), g.vs[1]), g.lo[1]);
  c[2] = __fadd_rn(__fmul_rn(__fadd_rn((float)vc[1], 0.5f), g.vs[2]), g.lo[2]);
}

// slot j of pillar v -> kept-point row, with the reference's cyclic padding (sst_ops_gpu.cu:30-39)
__device__ __forceinline__ int gt_point(const GtArgs& g, int beg, int cnt, int j) {
  int kept = cnt < g.K ? cnt : g.K;
  int s = j < kept ? j : j % cnt;
  return g.order[beg + s];
}

__global__ void gt_group_kernel(GtArgs g, int64_t m, float* __restrict__ gt, int64_t* __restrict__ group_inds) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * g.K) return;
  int64_t v = t / g.K;
  int j = (int)(t - v * g.K);
  int beg = g.offset[v], cnt = g.offset[v + 1] - beg;
  if (cnt <= 0) {
    if (group_inds) group_inds[t] = -1;
    return;
  }
  int p = gt_point(g, beg, cnt, j);
  if (group_inds) group_inds[t] = p;
  if (gt) {
    float c[3];
    voxel_center(g, v, c);
    const float* q = g.pts + (int64_t)p * g.stride + 1;
    gt[t * 3 + 0] = __fsub_rn(q[0], c[0]);
    gt[t * 3 + 1] = __fsub_rn(q[1], c[1]);
    gt[t * 3 + 2] = __fsub_rn(q[2], c[2]);
  }
}

struct ChamferArgs {
  const float* pred;     // (M, P1, 3)
  const float* gt;       // (M, P2, 3) or null -> fused gather through GtArgs
  const float* w;        // (M,)
  int64_t m;
  int P1, P2;
  double* acc;           // [3]: sum_x, sum_y, sum_w
  float* loss;           // scalar
  const float* gout;     // upstream gradient (scalar) for backward
  float* dpred;          // (M, P1, 3)
};

__device__ __forceinline__ void load_gt(const ChamferArgs& a, const GtArgs& g, int64_t v, int j, float y[3]) {
  if (a.gt) {
    const float* p = a.gt + (v * a.P2 + j) * 3;
    y[0] = p[0]; y[1] = p[1]; y[2] = p[2];
  } else {
    int beg = g.offset[v], cnt = g.offset[v + 1] - beg;
    float c[3];
    voxel_center(g, v, c);
    const float* q = g.pts + (int64_t)gt_point(g, beg, cnt, j) * g.stride + 1;
    y[0] = __fsub_rn(q[0], c[0]); y[1] = __fsub_rn(q[1], c[1]); y[2] = __fsub_rn(q[2], c[2]);
  }
}

__device__ __forceinline__ float sqdist(const float a[3], const float b[3]) {
  float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  return dx * dx + dy * dy + dz * dz;
}

// warp per pillar.  BWD = false: accumulate the weighted sums.  BWD = true: write dpred.
// The forward is a grid-stride loop: a warp keeps its three partial sums in registers over all its pillars and a block adds ONE set
// of three double atomics (41k warps x 3 atomics on the same three addresses were 2/3 of the forward kernel's 207 us).
template <bool BWD>
__global__ void chamfer_kernel(ChamferArgs a, GtArgs g) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double acc_x = 0.0, acc_y = 0.0, acc_w = 0.0;
  for (int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < a.m; v += warps) {
  float w = a.w[v];
  if (w == 0.f) {
    if (BWD) for (int e = lane; e < a.P1 * 3; e += 32) a.dpred[v * a.P1 * 3 + e] = 0.f;
    continue;
  }
  float x[3] = {0.f, 0.f, 0.f};
  if (lane < a.P1) {
    const float* p = a.pred + (v * a.P1 + lane) * 3;
    x[0] = p[0]; x[1] = p[1]; x[2] = p[2];
  }
  float y0[3], y1[3];
  bool has0 = lane < a.P2, has1 = lane + 32 < a.P2;
  if (has0) load_gt(a, g, v, lane, y0);
  if (has1) load_gt(a, g, v, lane + 32, y1);
  float best0 = INFINITY, best1 = INFINITY;   // per gt: min over preds
  int arg0 = 0, arg1 = 0;
  float sum_x = 0.f;                           // sum_i min_j
  float gx[3] = {0.f, 0.f, 0.f};               // BWD: x-direction gradient of my pred (lane < P1)
  for (int i = 0; i < a.P1; ++i) {
    float xi[3] = {__shfl_sync(0xffffffffu, x[0], i), __shfl_sync(0xffffffffu, x[1], i), __shfl_sync(0xffffffffu, x[2], i)};
    float d0 = has0 ? sqdist(xi, y0) : INFINITY;
    float d1 = has1 ? sqdist(xi, y1) : INFINITY;
    if (d0 < best0) { best0 = d0; arg0 = i; }
    if (d1 < best1) { best1 = d1; arg1 = i; }
    // nearest gt of pred i: min over the 64 candidates, lowest index wins ties
    float dm = fminf(d0, d1);
    int jm = d0 <= d1 ? lane : lane + 32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float od = __shfl_xor_sync(0xffffffffu, dm, o);
      int oj = __shfl_xor_sync(0xffffffffu, jm, o);
      if (od < dm || (od == dm && oj < jm)) { dm = od; jm = oj; }
    }
    sum_x += dm;
    if (BWD) {
      // fetch y_{a(i)} from its owner lane
      int owner = jm & 31;
      float ya[3];
      ya[0] = __shfl_sync(0xffffffffu, jm < 32 ? y0[0] : y1[0], owner);
      ya[1] = __shfl_sync(0xffffffffu, jm < 32 ? y0[1] : y1[1], owner);
      ya[2] = __shfl_sync(0xffffffffu, jm < 32 ? y0[2] : y1[2], owner);
      if (lane == i) { gx[0] = xi[0] - ya[0]; gx[1] = xi[1] - ya[1]; gx[2] = xi[2] - ya[2]; }
    }
  }
  if (!BWD) {
    float sum_y = warp_sum((has0 ? best0 : 0.f) + (has1 ? best1 : 0.f));
    acc_x += (double)(w * sum_x); acc_y += (double)(w * sum_y); acc_w += (double)w;   // identical in every lane
    continue;
  }
  // y-direction: sum over gts whose nearest pred is i of (x_i - y_j)
  float gy[3] = {0.f, 0.f, 0.f};
  for (int i = 0; i < a.P1; ++i) {
    float c = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (has0 && arg0 == i) { c += 1.f; s0 += y0[0]; s1 += y0[1]; s2 += y0[2]; }
    if (has1 && arg1 == i) { c += 1.f; s0 += y1[0]; s1 += y1[1]; s2 += y1[2]; }
    c = warp_sum(c); s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == i) { gy[0] = c * x[0] - s0; gy[1] = c * x[1] - s1; gy[2] = c * x[2] - s2; }
  }
  if (lane < a.P1) {
    float sw = (float)a.acc[2];
    float sc = sw > 0.f ? (*a.gout) * w / sw : 0.f;
    float cx = 2.f / a.P1, cy = 2.f / a.P2;
    float* o = a.dpred + (v * a.P1 + lane) * 3;
    o[0] = sc * (cx * gx[0] + cy * gy[0]);
    o[1] = sc * (cx * gx[1] + cy * gy[1]);
    o[2] = sc * (cx * gx[2] + cy * gy[2]);
  }
  }  // pillar loop
  if (!BWD) {
    __shared__ double part[8][3];
    const int wib = threadIdx.x >> 5;
    if (lane == 0) { part[wib][0] = acc_x; part[wib][1] = acc_y; part[wib][2] = acc_w; }
    __syncthreads();
    if (threadIdx.x < 3) {
      double t = 0.0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i][threadIdx.x];
      if (t != 0.0) atomicAdd(a.acc + threadIdx.x, t);
    }
  }
}

__global__ void chamfer_finalize_kernel(const double* __restrict__ acc, int P1, int P2, float* __restrict__ loss) {
  double sw = acc[2];
  *loss = sw > 0.0 ? (float)((acc[0] / P1 + acc[1] / P2) / sw) : 0.f;
}

static GtArgs make_gt(const float* pts, int stride, const int* offset, const int* order, const int64_t* vcoords, const float* lo,
                      const float* vs, int K) {
  GtArgs g{};
  g.pts = pts; g.stride = stride; g.offset = offset; g.order = order; g.vcoords = vcoords; g.K = K;
  for (int i = 0; i < 3; ++i) { g.lo[i] = lo ? lo[i] : 0.f; g.vs[i] = vs ? vs[i] : 1.f; }
  return g;
}

}  // namespace tmae

using namespace tmae;

extern "C" {

/* gt (M,K,3) = xyz[group] - voxel centre; group_inds (M,K) i64 (either output may be null) */
int tmae_gt_group(const float* points_kept, int32_t point_stride, const int32_t* voxel_offset, const int32_t* pt_order,
                  const int64_t* voxel_coords, const float* range_lo, const float* voxel, int64_t n_voxels, int32_t k, float* gt,
                  int64_t* group_inds, void* stream) {
  TMAE_CHECK_ARG(k >= 1, "k must be positive");
  if (n_voxels <= 0) return 0;
  GtArgs g = make_gt(points_kept, point_stride, voxel_offset, pt_order, voxel_coords, range_lo, voxel, k);
  gt_group_kernel<<<cdiv(n_voxels * k, 256), 256, 0, (cudaStream_t)stream>>>(g, n_voxels, gt, group_inds);
  TMAE_CHECK_LAUNCH();
  return 0;
}

size_t tmae_chamfer_workspace_bytes(void) { return 256; }

/* loss = chamfer_distance(pred, gt, weights=w) with pytorch3d defaults.  gt == null fuses the GT gather
 * (points_kept / CSR / voxel_coords as for tmae_gt_group).  state (3 doubles, caller-owned, = workspace) is
 * kept for the backward call. */
int tmae_chamfer_fwd(const float* pred, const float* gt, const float* w, int64_t n_voxels, int32_t p1, int32_t p2,
                     const float* points_kept, int32_t point_stride, const int32_t* voxel_offset, const int32_t* pt_order,
                     const int64_t* voxel_coords, const float* range_lo, const float* voxel, float* loss, void* state, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CHECK_ARG(p1 >= 1 && p1 <= 32 && p2 >= 1 && p2 <= 64, "need P1 <= 32 and P2 <= 64");
  TMAE_CHECK_ARG(gt != nullptr || (points_kept && voxel_offset && pt_order && voxel_coords), "no ground truth given");
  ChamferArgs a{};
  a.pred = pred; a.gt = gt; a.w = w; a.m = n_voxels; a.P1 = p1; a.P2 = p2; a.acc = (double*)state; a.loss = loss;
  GtArgs g = make_gt(points_kept, point_stride, voxel_offset, pt_order, voxel_coords, range_lo, voxel, p2);
  ProfScope prof("chamfer_fwd", (double)n_voxels * p1 * p2 * 8, (double)n_voxels * (p1 * 12.0 + 4.0 + 24.0 * 5), s);
  TMAE_CUDA(cudaMemsetAsync(state, 0, 3 * sizeof(double), s));
  if (n_voxels > 0) {
    const int64_t blocks = cdiv(n_voxels * 32, 256), cap = (int64_t)kNumSMs * 16;   // 8 warps per block; <= 16 blocks per SM in the grid
    chamfer_kernel<false><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(a, g);
  }
  chamfer_finalize_kernel<<<1, 1, 0, s>>>(a.acc, p1, p2, loss);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_chamfer_bwd(const float* grad_loss, const float* pred, const float* gt, const float* w, int64_t n_voxels, int32_t p1,
                     int32_t p2, const float* points_kept, int32_t point_stride, const int32_t* voxel_offset,
                     const int32_t* pt_order, const int64_t* voxel_coords, const float* range_lo, const float* voxel,
                     const void* state, float* dpred, void* stream) {
  TMAE_CHECK_ARG(p1 >= 1 && p1 <= 32 && p2 >= 1 && p2 <= 64, "need P1 <= 32 and P2 <= 64");
  if (n_voxels <= 0) return 0;
  ChamferArgs a{};
  a.pred = pred; a.gt = gt; a.w = w; a.m = n_voxels; a.P1 = p1; a.P2 = p2; a.acc = (double*)state; a.gout = grad_loss; a.dpred = dpred;
  GtArgs g = make_gt(points_kept, point_stride, voxel_offset, pt_order, voxel_coords, range_lo, voxel, p2);
  ProfScope prof("chamfer_bwd", (double)n_voxels * p1 * p2 * 8, (double)n_voxels * (p1 * 24.0 + 4.0 + 24.0 * 5), (cudaStream_t)stream);
  chamfer_kernel<true><<<cdiv(n_voxels * 32, 256), 256, 0, (cudaStream_t)stream>>>(a, g);
  TMAE_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
