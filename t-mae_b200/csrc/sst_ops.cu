// 1:1 replacements of the reference's only native op on the path, with the reference's own argument shapes so that
// pcdet/ops/sst_ops/sst_ops_utils.py:5-27 can bind them unchanged:
//   ingroup_inds_wrapper(group_inds i64 (N,), out i64 (N,))          pcdet/ops/sst_ops/src/sst_ops.cpp:21-33, sst_ops_gpu.cu:14-20
//   group_inner_inds_wrapper(inverse i64 (N,), group_inds i64 (M,K)) sst_ops.cpp:35-48, sst_ops_gpu.cu:22-39
// The reference hands out slots in atomicAdd arrival order (nondeterministic).  These return the CANONICAL result -- the one a
// serial run of the reference kernel yields: rank by element index inside each group (stable), first K indices per group, cyclic
// padding group_inds[g][i] = group_inds[g][i mod cnt].  A stable radix sort of (group id, element index) gives both: the rank
// is the distance to the group's first sorted position.  (The drop-in modules of this library never call them: they take slots from
// the window occupancy words and the voxeliser's CSR; these exist for callers that keep the reference's Python.)
// Unlike the reference there is no host sync (no .max().item()), no cudaMalloc per call and no exit(): the caller passes the
// workspace and, for ingroup_inds, nothing else -- group ids may be any non-negative int64.
#include <cub/cub.cuh>

#include "common.cuh"

namespace tmae {

__global__ void iota_kernel(int32_t* __restrict__ v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int32_t)i;
}
// first sorted position of every element's group (0 elsewhere), to be max-scanned
__global__ void seg_start_kernel(const int64_t* __restrict__ keys, int32_t* __restrict__ start, int64_t n) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) start[p] = (p == 0 || keys[p] != keys[p - 1]) ? (int32_t)p : 0;
}
__global__ void rank_scatter_kernel(const int32_t* __restrict__ start, const int32_t* __restrict__ idx, int64_t* __restrict__ out, int64_t n) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) out[idx[p]] = p - start[p];
}
// group_inds[g][rank] = element index for rank < K; the group's LAST sorted element also writes the cyclic padding of its row
__global__ void group_fill_kernel(const int64_t* __restrict__ keys, const int32_t* __restrict__ start, const int32_t* __restrict__ idx,
                                  int64_t* __restrict__ group_inds, int64_t n, int64_t m, int k) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int64_t g = keys[p];
  if (g < 0 || g >= m) return;
  const int rank = (int)(p - start[p]);
  if (rank < k) group_inds[g * k + rank] = idx[p];
}
__global__ void group_pad_kernel(const int64_t* __restrict__ keys, const int32_t* __restrict__ start, const int32_t* __restrict__ idx,
                                 int64_t* __restrict__ group_inds, int64_t n, int64_t m, int k) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (p + 1 < n && keys[p + 1] == keys[p]) return;   // not the last element of its group
  const int64_t g = keys[p];
  if (g < 0 || g >= m) return;
  const int s = start[p];
  const int cnt = (int)(p - s) + 1;
  for (int i = cnt; i < k; ++i) group_inds[g * k + i] = idx[s + (i % cnt)];   // sst_ops_gpu.cu:37-38 (cnt < K here)
}

struct MaxOp {
  __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; }
};

struct SortWs {
  int64_t* keys; int32_t* idx_in; int32_t* idx; int32_t* start; void* cub; size_t cub_bytes;
};
static size_t cub_bytes_for(int64_t n) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const int64_t*)nullptr, (int64_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  cub::DeviceScan::InclusiveScan(nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr, MaxOp(), (int)n);
  return a > b ? a : b;
}
static size_t ws_total(int64_t n) {
  return ws_bytes(n, 8) + 3 * ws_bytes(n, 4) + ws_bytes((int64_t)cub_bytes_for(n), 1);
}
static bool carve(SortWs& w, void* ws, size_t bytes, int64_t n) {
  Workspace W(ws, bytes);
  w.keys = W.take<int64_t>(n);
  w.idx_in = W.take<int32_t>(n);
  w.idx = W.take<int32_t>(n);
  w.start = W.take<int32_t>(n);
  w.cub_bytes = cub_bytes_for(n);
  w.cub = W.take<char>((int64_t)w.cub_bytes);
  return w.cub != nullptr;
}
// sorted keys, element index per sorted position, first sorted position of each position's group
static int sort_groups(const int64_t* group, int64_t n, SortWs& w, cudaStream_t s) {
  const int T = 256;
  iota_kernel<<<cdiv(n, T), T, 0, s>>>(w.idx_in, n);
  size_t cb = w.cub_bytes;
  if (cub::DeviceRadixSort::SortPairs(w.cub, cb, group, w.keys, w.idx_in, w.idx, (int)n, 0, 64, s) != cudaSuccess) return TMAE_ERR_CUDA;
  seg_start_kernel<<<cdiv(n, T), T, 0, s>>>(w.keys, w.start, n);
  cb = w.cub_bytes;
  if (cub::DeviceScan::InclusiveScan(w.cub, cb, w.start, w.start, MaxOp(), (int)n, s) != cudaSuccess) return TMAE_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? 0 : TMAE_ERR_CUDA;
}

}  // namespace tmae

using namespace tmae;

extern "C" {

size_t tmae_sst_ops_workspace_bytes(int64_t n) { return n > 0 ? ws_total(n) : 256; }

int tmae_ingroup_inds(const int64_t* group_inds, int64_t* out_inds, int64_t n, void* workspace, size_t workspace_bytes, void* stream) {
  TMAE_CHECK_ARG(n < ((int64_t)1 << 31), "element counts must fit int32");
  if (n <= 0) return 0;
  TMAE_CHECK_ARG(group_inds && out_inds && workspace, "null pointer");
  SortWs w;
  TMAE_CHECK_ARG(workspace_bytes >= ws_total(n) && carve(w, workspace, workspace_bytes, n), "workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  if (sort_groups(group_inds, n, w, s)) { set_error("tmae_ingroup_inds: sort failed"); return TMAE_ERR_CUDA; }
  rank_scatter_kernel<<<cdiv(n, 256), 256, 0, s>>>(w.start, w.idx, out_inds, n);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_group_inner_inds(const int64_t* inverse_inds, int64_t n, int64_t* group_inds, int64_t m, int32_t k, void* workspace, size_t workspace_bytes,
                          void* stream) {
  TMAE_CHECK_ARG(n < ((int64_t)1 << 31) && k > 0, "element counts must fit int32 and K must be positive");
  if (n <= 0 || m <= 0) return 0;
  TMAE_CHECK_ARG(inverse_inds && group_inds && workspace, "null pointer");
  SortWs w;
  TMAE_CHECK_ARG(workspace_bytes >= ws_total(n) && carve(w, workspace, workspace_bytes, n), "workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  if (sort_groups(inverse_inds, n, w, s)) { set_error("tmae_group_inner_inds: sort failed"); return TMAE_ERR_CUDA; }
  group_fill_kernel<<<cdiv(n, 256), 256, 0, s>>>(w.keys, w.start, w.idx, group_inds, n, m, k);
  group_pad_kernel<<<cdiv(n, 256), 256, 0, s>>>(w.keys, w.start, w.idx, group_inds, n, m, k);
  TMAE_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
