// Dynamic voxelisation and the VFE point features (first half of A3).
//
// Replaces: get_in_range_mask (pcdet/utils/common_utils.py:66-76), the boolean compaction,
// coords.unique(dim=0, return_inverse=True) and torch_scatter.scatter(mean)
// (pcdet/models/backbones_3d/vfe/temporal_dyn_vfe.py:69-85).
//
// B200-first design: no hash, no sort of 4-column rows.  The pillar grid is small (B*Y*X cells),
// so a dense occupancy array + one prefix sum IS the voxel id in lexicographic (b,z,y,x) order,
// which is the order torch.unique(dim=0) returns.  Points are also grouped per voxel (CSR,
// ascending point index inside a voxel) so that the per-voxel mean is a deterministic in-order
// sum (bit-identical to a sequential index_add) and later per-voxel kernels need no atomics.
#include "common.cuh"

namespace tmae {

struct VoxGeom {
  float lo[3];
  float vs[3];
  int g[3];  // gx, gy, gz
  int batch;
};

// cell id in (b, z, y, x) lexicographic order, or -1 when the point is dropped.
__device__ __forceinline__ int point_cell(const float* __restrict__ p, const VoxGeom& G, long long c[3]) {
  // fp32 subtract then IEEE divide, truncation toward zero (torch .to(int64)); common_utils.py:74
  c[0] = (long long)__fdiv_rn(__fsub_rn(p[1], G.lo[0]), G.vs[0]);
  c[1] = (long long)__fdiv_rn(__fsub_rn(p[2], G.lo[1]), G.vs[1]);
  c[2] = (long long)__fdiv_rn(__fsub_rn(p[3], G.lo[2]), G.vs[2]);
  bool keep = c[0] >= 0 && c[0] < G.g[0] && c[1] >= 0 && c[1] < G.g[1] && c[2] >= 0 && c[2] < G.g[2];
  long long b = (long long)p[0];
  if (!keep || b < 0 || b >= G.batch) return -1;
  return (int)(((b * G.g[2] + c[2]) * G.g[1] + c[1]) * G.g[0] + c[0]);
}

__global__ void vox_mark_kernel(const float* __restrict__ pts, int64_t n, int stride, VoxGeom G, int* __restrict__ cell,
                                int* __restrict__ keepflag, int* __restrict__ occ) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long c[3];
  int id = point_cell(pts + i * stride, G, c);
  cell[i] = id;
  keepflag[i] = id >= 0;
  if (id >= 0) occ[id] = 1;
}

__global__ void vox_write_kernel(const float* __restrict__ pts, int64_t n, int stride, VoxGeom G, const int* __restrict__ cell,
                                 const int* __restrict__ ppre, const int* __restrict__ rank, float* __restrict__ pts_out,
                                 int64_t* __restrict__ pcoords, int64_t* __restrict__ inverse, int* __restrict__ npts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int id = cell[i];
  if (id < 0) return;
  int j = ppre[i];
  int v = rank[id];
  const float* p = pts + i * stride;
  for (int f = 0; f < stride; ++f) pts_out[(int64_t)j * stride + f] = p[f];
  int x = id % G.g[0];
  int y = (id / G.g[0]) % G.g[1];
  int z = (id / (G.g[0] * G.g[1])) % G.g[2];
  int b = id / (G.g[0] * G.g[1] * G.g[2]);
  int64_t* pc = pcoords + (int64_t)j * 4;
  pc[0] = b; pc[1] = z; pc[2] = y; pc[3] = x;
  inverse[j] = v;
  atomicAdd(npts + v, 1);
}

__global__ void vox_place_kernel(const int64_t* __restrict__ inverse, const int* __restrict__ n_kept, const int* __restrict__ offset,
                                 int* __restrict__ cursor, int* __restrict__ order_tmp) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= *n_kept) return;
  int v = (int)inverse[j];
  int pos = offset[v] + atomicAdd(cursor + v, 1);
  order_tmp[pos] = (int)j;
}

// one warp per voxel: order the voxel's point ids ascending (canonical, = serial arrival order of
// sst_ops_gpu.cu:22-28) and take the in-order mean.
__global__ void vox_finish_kernel(const int* __restrict__ n_vox, const int* __restrict__ offset, const int* __restrict__ order_tmp,
                                  int* __restrict__ order, const float* __restrict__ pts_out, int stride,
                                  float* __restrict__ mean) {
  int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (v >= *n_vox) return;
  int beg = offset[v], end = offset[v + 1];
  int n = end - beg;
  for (int a = lane; a < n; a += 32) {
    int mine = order_tmp[beg + a];
    int r = 0;
    for (int k = 0; k < n; ++k) r += order_tmp[beg + k] < mine;
    order[beg + r] = mine;
  }
  __syncwarp();
  int nf = stride - 1;
  if (lane < nf) {
    float s = 0.f;
    for (int k = 0; k < n; ++k) s = __fadd_rn(s, pts_out[(int64_t)order[beg + k] * stride + 1 + lane]);
    mean[(int64_t)v * nf + lane] = __fdiv_rn(s, (float)n);
  }
}

__global__ void vox_coords_kernel(const int* __restrict__ occ, const int* __restrict__ rank, int64_t cells, VoxGeom G,
                                  int64_t* __restrict__ vcoords, int64_t* __restrict__ counts) {
  int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= cells) return;
  int per = G.g[0] * G.g[1] * G.g[2];
  if (id % per == 0) counts[2 + id / per] = rank[id];  // first voxel row of each sample
  if (!occ[id]) return;
  int64_t* vc = vcoords + (int64_t)rank[id] * 4;
  vc[3] = id % G.g[0];
  vc[2] = (id / G.g[0]) % G.g[1];
  vc[1] = (id / (G.g[0] * G.g[1])) % G.g[2];
  vc[0] = id / per;
}

__global__ void vox_counts_kernel(const int* __restrict__ n_kept, const int* __restrict__ n_vox, int64_t* __restrict__ counts) {
  counts[0] = *n_kept;
  counts[1] = *n_vox;
}

// X[j] = [f_center(3), x, y, z, feats..., f_cluster(3)]   (temporal_dyn_vfe.py:89-110)
__global__ void vfe_features_kernel(const float* __restrict__ pts, int64_t n, int stride, const int64_t* __restrict__ pcoords,
                                    const int64_t* __restrict__ inverse, const float* __restrict__ mean, VoxGeom G,
                                    float* __restrict__ X) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float* p = pts + j * stride;
  const int64_t* c = pcoords + j * 4;
  int nf = stride - 1;
  int w = nf + 6;
  float* x = X + j * w;
  const float* m = mean + inverse[j] * nf;
  // (c + 0.5) * vs + lo with separate roundings, as the reference's eager ops do
  x[0] = __fsub_rn(p[1], __fadd_rn(__fmul_rn(__fadd_rn((float)c[3], 0.5f), G.vs[0]), G.lo[0]));
  x[1] = __fsub_rn(p[2], __fadd_rn(__fmul_rn(__fadd_rn((float)c[2], 0.5f), G.vs[1]), G.lo[1]));
  x[2] = __fsub_rn(p[3], __fadd_rn(__fmul_rn(__fadd_rn((float)c[1], 0.5f), G.vs[2]), G.lo[2]));
  for (int f = 0; f < nf; ++f) x[3 + f] = p[1 + f];
  x[3 + nf + 0] = __fsub_rn(p[1], m[0]);
  x[3 + nf + 1] = __fsub_rn(p[2], m[1]);
  x[3 + nf + 2] = __fsub_rn(p[3], m[2]);
}

static VoxGeom make_geom(const float* range_lo, const float* voxel, const int32_t* grid, int batch) {
  VoxGeom G;
  for (int i = 0; i < 3; ++i) { G.lo[i] = range_lo[i]; G.vs[i] = voxel[i]; G.g[i] = grid[i]; }
  G.batch = batch;
  return G;
}

}  // namespace tmae

using namespace tmae;

extern "C" {

size_t tmae_voxelize_workspace_bytes(int64_t n_points, int32_t batch, const int32_t* grid) {
  int64_t cells = (int64_t)batch * grid[0] * grid[1] * grid[2];
  int64_t mcap = n_points < cells ? n_points : cells;
  size_t b = 0;
  b += ws_bytes(n_points, 4) * 3;                        // cell, keepflag, ppre
  b += ws_bytes(cells, 4) * 2;                           // occ, rank
  b += ws_bytes(mcap + 1, 4);                            // cursor
  b += ws_bytes(n_points, 4);                            // order_tmp
  b += ws_bytes(scan_scratch_elems(cells > n_points ? cells : n_points), 4);
  b += ws_bytes(8, 4);                                   // totals
  return b + 1024;
}

int tmae_voxelize(const float* points, int64_t n_points, int32_t point_stride, const float* range_lo, const float* voxel,
                  const int32_t* grid, int32_t batch, float* points_out, int64_t* point_coords, int64_t* inverse,
                  int64_t* voxel_coords, float* voxel_mean, int32_t* voxel_npts, int32_t* voxel_offset, int32_t* pt_order,
                  int64_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CHECK_ARG(point_stride >= 4 && point_stride <= 16, "point_stride must be in [4,16]");
  TMAE_CHECK_ARG(batch >= 1 && grid[0] > 0 && grid[1] > 0 && grid[2] > 0, "bad grid/batch");
  int64_t cells = (int64_t)batch * grid[0] * grid[1] * grid[2];
  TMAE_CHECK_ARG(cells < (1ll << 31), "grid too large for int32 cell ids");
  TMAE_CHECK_ARG(workspace_bytes >= tmae_voxelize_workspace_bytes(n_points, batch, grid), "workspace too small");
  int64_t mcap = n_points < cells ? n_points : cells;
  VoxGeom G = make_geom(range_lo, voxel, grid, batch);
  Workspace ws(workspace, workspace_bytes);
  int* cell = ws.take<int>(n_points);
  int* keepflag = ws.take<int>(n_points);
  int* ppre = ws.take<int>(n_points);
  int* occ = ws.take<int>(cells);
  int* rank = ws.take<int>(cells);
  int* cursor = ws.take<int>(mcap + 1);
  int* order_tmp = ws.take<int>(n_points);
  int* scratch = ws.take<int>(scan_scratch_elems(cells > n_points ? cells : n_points));
  int* totals = ws.take<int>(8);
  TMAE_CHECK_ARG(totals != nullptr, "workspace carve failed");

  // algorithmic traffic: read n*stride*4 ; write kept points, coords (32 B), inverse (8 B) per point + 48 B per voxel
  ProfScope prof("voxelize", 0, (double)n_points * (point_stride * 8.0 + 40.0), s);
  TMAE_CUDA(cudaMemsetAsync(occ, 0, cells * sizeof(int), s));
  TMAE_CUDA(cudaMemsetAsync(voxel_npts, 0, mcap * sizeof(int), s));
  TMAE_CUDA(cudaMemsetAsync(cursor, 0, (mcap + 1) * sizeof(int), s));
  TMAE_CUDA(cudaMemsetAsync(counts, 0, (2 + batch) * sizeof(int64_t), s));
  const int T = 256;
  if (n_points > 0) {
    vox_mark_kernel<<<cdiv(n_points, T), T, 0, s>>>(points, n_points, point_stride, G, cell, keepflag, occ);
    TMAE_CHECK_LAUNCH();
  }
  if (scan_exclusive_i32(keepflag, ppre, n_points, totals + 0, scratch, s)) return TMAE_ERR_CUDA;
  if (scan_exclusive_i32(occ, rank, cells, totals + 1, scratch, s)) return TMAE_ERR_CUDA;
  if (n_points > 0) {
    vox_write_kernel<<<cdiv(n_points, T), T, 0, s>>>(points, n_points, point_stride, G, cell, ppre, rank, points_out, point_coords,
                                                     inverse, voxel_npts);
    TMAE_CHECK_LAUNCH();
  }
  // CSR offsets over the voxel capacity (entries past n_vox are zero)
  if (scan_exclusive_i32(voxel_npts, voxel_offset, mcap, nullptr, scratch, s)) return TMAE_ERR_CUDA;
  // offset[mcap] = n_kept closes the last segment whatever n_vox turns out to be: rows >= n_vox are empty
  TMAE_CUDA(cudaMemcpyAsync(voxel_offset + mcap, totals + 0, sizeof(int), cudaMemcpyDeviceToDevice, s));
  if (n_points > 0) {
    vox_place_kernel<<<cdiv(n_points, T), T, 0, s>>>(inverse, totals + 0, voxel_offset, cursor, order_tmp);
    TMAE_CHECK_LAUNCH();
    vox_finish_kernel<<<cdiv(mcap * 32, T), T, 0, s>>>(totals + 1, voxel_offset, order_tmp, pt_order, points_out, point_stride,
                                                       voxel_mean);
    TMAE_CHECK_LAUNCH();
  }
  vox_coords_kernel<<<cdiv(cells, T), T, 0, s>>>(occ, rank, cells, G, voxel_coords, counts);
  TMAE_CHECK_LAUNCH();
  vox_counts_kernel<<<1, 1, 0, s>>>(totals + 0, totals + 1, counts);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_vfe_point_features(const float* points_kept, int64_t n_kept, int32_t point_stride, const int64_t* point_coords,
                            const int64_t* inverse, const float* voxel_mean, const float* range_lo, const float* voxel,
                            float* features, void* stream) {
  TMAE_CHECK_ARG(point_stride >= 4 && point_stride <= 16, "point_stride must be in [4,16]");
  if (n_kept == 0) return 0;
  int32_t g[3] = {1, 1, 1};
  VoxGeom G = make_geom(range_lo, voxel, g, 1);
  vfe_features_kernel<<<cdiv(n_kept, 256), 256, 0, (cudaStream_t)stream>>>(points_kept, n_kept, point_stride, point_coords, inverse,
                                                                           voxel_mean, G, features);
  TMAE_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
