// Neighbour tables ("rulebooks") for the 2-D sparse convolutions on the path (row A9):
// SubMConv2d k3 and SparseConv2d k3 s2 p1 (pcdet/utils/spconv_utils.py:37-56; instances at
// spt_backbone.py:282-284,302-304 and SiamWCA.py:312-314).  spconv builds these with a GPU hash
// table; the BEV grid is small, so a dense (B,Y,X) row-index map replaces the hash and a prefix
// sum over output-cell occupancy emits strided-conv outputs in lexicographic (b,y,x) order.
// The contraction itself is the gather-GEMM in gemm.cu / gemm_tc.cu.
#include "common.cuh"

namespace tmae {

__device__ __forceinline__ int64_t active_rows(int64_t cap, const int* dev) {
  if (!dev) return cap;
  int64_t d = *dev;
  return d < cap ? d : cap;
}

__global__ void map_fill_kernel(const int* __restrict__ idx, int64_t cap, const int* __restrict__ ndev, int Y, int X, int* __restrict__ map) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= active_rows(cap, ndev)) return;
  const int* p = idx + i * 3;
  map[((int64_t)p[0] * Y + p[1]) * X + p[2]] = (int)i;
}

// table[i, ky*3+kx] = row of the input at (y + ky - 1, x + kx - 1) or -1
__global__ void subm_table_kernel(const int* __restrict__ idx, int64_t cap, const int* __restrict__ ndev, int Y, int X,
                                  const int* __restrict__ map, int* __restrict__ table) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t n = active_rows(cap, ndev);
  if (t >= cap * 9) return;
  int64_t i = t / 9;
  int tap = (int)(t - i * 9);
  if (i >= n) { table[t] = -1; return; }
  const int* p = idx + i * 3;
  int y = p[1] + tap / 3 - 1, x = p[2] + tap % 3 - 1;
  table[t] = (y >= 0 && y < Y && x >= 0 && x < X) ? map[((int64_t)p[0] * Y + y) * X + x] : -1;
}

// strided conv k3 s2 p1: mark every output cell whose receptive field holds this input
__global__ void strided_mark_kernel(const int* __restrict__ idx, int64_t cap, const int* __restrict__ ndev, int Yo, int Xo,
                                    int* __restrict__ occ) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= active_rows(cap, ndev)) return;
  const int* p = idx + i * 3;
  for (int ky = 0; ky < 3; ++ky) {
    int ty = p[1] + 1 - ky;
    if (ty < 0 || (ty & 1)) continue;
    int oy = ty >> 1;
    if (oy >= Yo) continue;
    for (int kx = 0; kx < 3; ++kx) {
      int tx = p[2] + 1 - kx;
      if (tx < 0 || (tx & 1)) continue;
      int ox = tx >> 1;
      if (ox >= Xo) continue;
      occ[((int64_t)p[0] * Yo + oy) * Xo + ox] = 1;
    }
  }
}

// one thread per output cell: emit its row (index triple + forward table)
__global__ void strided_emit_kernel(const int* __restrict__ occ, const int* __restrict__ rank, int64_t cells, int Yi, int Xi, int Yo,
                                    int Xo, const int* __restrict__ map_in, int64_t out_cap, int* __restrict__ idx_out,
                                    int* __restrict__ table, int* __restrict__ map_out) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cells) return;
  if (!occ[c]) { map_out[c] = -1; return; }
  int r = rank[c];
  if (r >= out_cap) { map_out[c] = -1; return; }
  map_out[c] = r;
  int ox = (int)(c % Xo), oy = (int)((c / Xo) % Yo), b = (int)(c / ((int64_t)Xo * Yo));
  idx_out[r * 3 + 0] = b; idx_out[r * 3 + 1] = oy; idx_out[r * 3 + 2] = ox;
  for (int tap = 0; tap < 9; ++tap) {
    int y = 2 * oy - 1 + tap / 3, x = 2 * ox - 1 + tap % 3;
    table[(int64_t)r * 9 + tap] = (y >= 0 && y < Yi && x >= 0 && x < Xi) ? map_in[((int64_t)b * Yi + y) * Xi + x] : -1;
  }
}

// transposed table for backward-data: tableT[i, tap] = output row o with table[o, tap] == i
__global__ void strided_tableT_kernel(const int* __restrict__ idx, int64_t cap, const int* __restrict__ ndev, int Yo, int Xo,
                                      const int* __restrict__ map_out, int* __restrict__ tableT) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cap * 9) return;
  int64_t i = t / 9;
  int tap = (int)(t - i * 9);
  int v = -1;
  if (i < active_rows(cap, ndev)) {
    const int* p = idx + i * 3;
    int ty = p[1] + 1 - tap / 3, tx = p[2] + 1 - tap % 3;
    if (ty >= 0 && tx >= 0 && !(ty & 1) && !(tx & 1) && (ty >> 1) < Yo && (tx >> 1) < Xo)
      v = map_out[((int64_t)p[0] * Yo + (ty >> 1)) * Xo + (tx >> 1)];
  }
  tableT[t] = v;
}

}  // namespace tmae

using namespace tmae;

extern "C" {

size_t tmae_subm_table_workspace_bytes(int32_t batch, int32_t y, int32_t x) { return ws_bytes((int64_t)batch * y * x, 4) + 256; }

/* SubMConv2d k3 neighbour table.  indices (rows_cap,3) i32 [b,y,x]; rows_dev (nullable) = device row
 * count (rows beyond it get -1 tables).  table (rows_cap, 9) i32.  The backward-data table is the same
 * table with the taps reversed (tmae_transpose_taps flip=1). */
int tmae_subm_table(const int32_t* indices, int64_t rows_cap, const int32_t* rows_dev, int32_t batch, int32_t y, int32_t x,
                    int32_t* table, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CHECK_ARG(workspace_bytes >= tmae_subm_table_workspace_bytes(batch, y, x), "workspace too small");
  if (rows_cap <= 0) return 0;
  int* map = (int*)workspace;
  TMAE_CUDA(cudaMemsetAsync(map, 0xff, (size_t)batch * y * x * 4, s));
  map_fill_kernel<<<cdiv(rows_cap, 256), 256, 0, s>>>(indices, rows_cap, rows_dev, y, x, map);
  subm_table_kernel<<<cdiv(rows_cap * 9, 256), 256, 0, s>>>(indices, rows_cap, rows_dev, y, x, map, table);
  TMAE_CHECK_LAUNCH();
  return 0;
}

size_t tmae_strided_table_workspace_bytes(int32_t batch, int32_t y_in, int32_t x_in) {
  int yo = (y_in + 2 - 3) / 2 + 1, xo = (x_in + 2 - 3) / 2 + 1;
  int64_t ci = (int64_t)batch * y_in * x_in, co = (int64_t)batch * yo * xo;
  return ws_bytes(ci, 4) + 3 * ws_bytes(co, 4) + ws_bytes(scan_scratch_elems(co), 4) + 1024;
}

/* SparseConv2d k3 s2 p1: output site set (lexicographic b,y,x rows), forward table (out_cap, 9) and
 * backward-data table (rows_cap, 9).  n_out (device i32) receives the output row count; out_cap must be
 * >= min(4 * rows, batch * y_out * x_out). */
int tmae_strided_table(const int32_t* indices, int64_t rows_cap, const int32_t* rows_dev, int32_t batch, int32_t y_in, int32_t x_in,
                       int32_t* indices_out, int64_t out_cap, int32_t* n_out, int32_t* table, int32_t* table_t, void* workspace,
                       size_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CHECK_ARG(workspace_bytes >= tmae_strided_table_workspace_bytes(batch, y_in, x_in), "workspace too small");
  int yo = (y_in + 2 - 3) / 2 + 1, xo = (x_in + 2 - 3) / 2 + 1;
  int64_t ci = (int64_t)batch * y_in * x_in, co = (int64_t)batch * yo * xo;
  Workspace ws(workspace, workspace_bytes);
  int* map_in = ws.take<int>(ci);
  int* occ = ws.take<int>(co);
  int* rank = ws.take<int>(co);
  int* map_out = ws.take<int>(co);
  int* scratch = ws.take<int>(scan_scratch_elems(co));
  TMAE_CHECK_ARG(scratch != nullptr, "workspace carve failed");
  TMAE_CUDA(cudaMemsetAsync(map_in, 0xff, (size_t)ci * 4, s));
  TMAE_CUDA(cudaMemsetAsync(occ, 0, (size_t)co * 4, s));
  if (rows_cap > 0) {
    map_fill_kernel<<<cdiv(rows_cap, 256), 256, 0, s>>>(indices, rows_cap, rows_dev, y_in, x_in, map_in);
    strided_mark_kernel<<<cdiv(rows_cap, 256), 256, 0, s>>>(indices, rows_cap, rows_dev, yo, xo, occ);
    TMAE_CHECK_LAUNCH();
  }
  if (scan_exclusive_i32(occ, rank, co, n_out, scratch, s)) return TMAE_ERR_CUDA;
  strided_emit_kernel<<<cdiv(co, 256), 256, 0, s>>>(occ, rank, co, y_in, x_in, yo, xo, map_in, out_cap, indices_out, table, map_out);
  if (rows_cap > 0) strided_tableT_kernel<<<cdiv(rows_cap * 9, 256), 256, 0, s>>>(indices, rows_cap, rows_dev, yo, xo, map_out, table_t);
  TMAE_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
