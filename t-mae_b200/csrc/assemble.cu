// Rows N2 / N3 of SURVEY.md section 8f -- the step right in front of the VFE, on the GPU: from the raw per-sample
// point arrays of one frame set to the collated (N', 1 + F) tensor `points` / `points_prev` the VFE consumes.
//   remove_ego_points(points, 2)          pcdet/datasets/once_temporal/once_eval/once_utils.py:43-45,
//                                          once_temporal_dataset.py:167-168        drop |x| < r AND |y| < r (raw coordinates)
//   convert_prv_frame_to_cur              once_utils.py:4-29                        prev -> global -> current frame, in float64
//   mask_points_by_range                  pcdet/utils/common_utils.py:124-127,
//                                          processor/data_processor.py:81-83        keep x0 <= x <= x1, y0 <= y <= y1 (float64 compare)
//   collate_batch                         pcdet/datasets/dataset.py:203-208         prepend the sample index
// and the final `.float()` of load_data_to_gpu (pcdet/models/__init__.py:16-23).  The reference does all of this per
// sample in numpy on DataLoader workers.  Order inside a sample is preserved (stable compaction), so the output is bit-
// identical to the reference's before its random shuffle_points (data_processor.py:92-102), which only permutes rows.
// HBM-bound: 4 F bytes read and 4 (1 + F) bytes written per kept point.
#include "common.cuh"

namespace tmae {

struct AsmArgs {
  const float* raw; const int64_t* offs; int batch, feats;
  const double* xform; const uint8_t* flags;   // per sample: two 3x4 row-major affine maps (prev->global, global->cur), two "apply" flags
  double ego_r, x0, y0, x1, y1;
  int64_t n;
};

__device__ __forceinline__ int sample_of(const AsmArgs& a, int64_t i) {
  int b = 0;
  while (b + 1 < a.batch && i >= a.offs[b + 1]) ++b;
  return b;
}
// xyz in float64 with the reference's evaluation order: np.dot(p, R.T) + t (dot of length 3, then the translation), and
// np.dot([p, 1], A.T) for the 4x4 inverse (dot of length 4)
__device__ __forceinline__ void transform(const AsmArgs& a, int b, double& x, double& y, double& z) {
  if (!a.xform) return;
  const double* A = a.xform + (int64_t)b * 24;
  if (a.flags[2 * b]) {
    const double nx = fma(z, A[2], fma(y, A[1], x * A[0])) + A[3];
    const double ny = fma(z, A[6], fma(y, A[5], x * A[4])) + A[7];
    const double nz = fma(z, A[10], fma(y, A[9], x * A[8])) + A[11];
    x = nx; y = ny; z = nz;
  }
  if (a.flags[2 * b + 1]) {
    const double* B = A + 12;
    const double nx = fma(1.0, B[3], fma(z, B[2], fma(y, B[1], x * B[0])));
    const double ny = fma(1.0, B[7], fma(z, B[6], fma(y, B[5], x * B[4])));
    const double nz = fma(1.0, B[11], fma(z, B[10], fma(y, B[9], x * B[8])));
    x = nx; y = ny; z = nz;
  }
}
__device__ __forceinline__ bool keep_point(const AsmArgs& a, int64_t i, int& b, double& x, double& y, double& z) {
  const float* p = a.raw + i * a.feats;
  b = sample_of(a, i);
  x = p[0]; y = p[1]; z = p[2];
  if (fabs(x) < a.ego_r && fabs(y) < a.ego_r) return false;   // ego vehicle returns, raw coordinates
  transform(a, b, x, y, z);
  return x >= a.x0 && x <= a.x1 && y >= a.y0 && y <= a.y1;
}

// Stable compaction in two launches, no atomics and no memset: pass 1 leaves one kept-count per 512-point block, pass 2
// sums the counts in front of its block (<= a few hundred integers, L2-resident), ranks its own points in input order
// (ballot + popc per warp, warp totals through shared memory) and writes the collated rows.  Rows [count, n_points) of
// `out` are filled with a far out-of-range sentinel ([0, 1e6, 1e6, 1e6, 0...]) so the whole capacity is defined: a caller
// that wants no host read hands all n_points rows to the voxeliser, whose range test (A1) drops the tail.
constexpr int kAsmThreads = 256, kAsmItems = 2, kAsmTile = kAsmThreads * kAsmItems;
constexpr float kAsmSentinel = 1.0e6f;

__global__ void __launch_bounds__(kAsmThreads) asm_count_kernel(AsmArgs a, int* __restrict__ block_count) {
  __shared__ int s_total;
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kAsmTile;
  int kept = 0;
#pragma unroll
  for (int k = 0; k < kAsmItems; ++k) {   // independent points: the loads of all items are in flight together
    const int64_t i = base + k * kAsmThreads + threadIdx.x;
    int b; double x, y, z;
    kept += (i < a.n && keep_point(a, i, b, x, y, z)) ? 1 : 0;
  }
  kept = __reduce_add_sync(0xffffffffu, kept);
  if ((threadIdx.x & 31) == 0 && kept) atomicAdd(&s_total, kept);
  __syncthreads();
  if (threadIdx.x == 0) block_count[blockIdx.x] = s_total;
}

__global__ void __launch_bounds__(kAsmThreads) asm_write_kernel(AsmArgs a, const int* __restrict__ block_count, float* __restrict__ out,
                                                                int64_t* __restrict__ count) {
  constexpr int kWarps = kAsmThreads / 32;
  __shared__ int s_red[2][kWarps];
  __shared__ int s_warp[kAsmItems][kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * kAsmTile;
  // the points first (independent loads), then the counts in front of this block
  bool valid[kAsmItems], keep[kAsmItems];
  int b[kAsmItems], in_warp[kAsmItems];
  double x[kAsmItems], y[kAsmItems], z[kAsmItems];
#pragma unroll
  for (int k = 0; k < kAsmItems; ++k) {
    const int64_t i = base + k * kAsmThreads + threadIdx.x;
    valid[k] = i < a.n;
    b[k] = 0; x[k] = y[k] = z[k] = 0;
    keep[k] = valid[k] && keep_point(a, i, b[k], x[k], y[k], z[k]);
  }
  int before = 0, total = 0;
  for (int j = threadIdx.x; j < (int)gridDim.x; j += kAsmThreads) {
    const int c = block_count[j];
    total += c;
    if (j < (int)blockIdx.x) before += c;
  }
  before = __reduce_add_sync(0xffffffffu, before);
  total = __reduce_add_sync(0xffffffffu, total);
  if (lane == 0) { s_red[0][warp] = before; s_red[1][warp] = total; }
#pragma unroll
  for (int k = 0; k < kAsmItems; ++k) {
    const unsigned bal = __ballot_sync(0xffffffffu, keep[k]);
    in_warp[k] = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_warp[k][warp] = __popc(bal);
  }
  __syncthreads();
  before = total = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) { before += s_red[0][w]; total += s_red[1][w]; }
  if (blockIdx.x == 0 && threadIdx.x == 0) *count = total;

  const int stride = a.feats + 1;
  int64_t kept_before = before;   // kept points in front of the current 256-point chunk
#pragma unroll
  for (int k = 0; k < kAsmItems; ++k) {
    int in_front = in_warp[k], chunk = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const int c = s_warp[k][w];
      chunk += c;
      if (w < warp) in_front += c;
    }
    if (valid[k]) {
      const int64_t i = base + k * kAsmThreads + threadIdx.x;
      const int64_t rank = kept_before + in_front;                 // kept points in front of point i
      float* o = out + (keep[k] ? rank : (int64_t)total + (i - rank)) * stride;
      if (keep[k]) {
        const float* p = a.raw + i * a.feats;
        o[0] = (float)b[k]; o[1] = (float)x[k]; o[2] = (float)y[k]; o[3] = (float)z[k];
        for (int f = 3; f < a.feats; ++f) o[1 + f] = p[f];
      } else {
        o[0] = 0.f; o[1] = o[2] = o[3] = kAsmSentinel;
        for (int f = 3; f < a.feats; ++f) o[1 + f] = 0.f;
      }
    }
    kept_before += chunk;
  }
}

}  // namespace tmae

using namespace tmae;

extern "C" {

size_t tmae_assemble_frames_workspace_bytes(int64_t n_points) { return ws_bytes(cdiv(n_points > 0 ? n_points : 1, kAsmTile), 4) + 256; }

int tmae_assemble_frames(const float* raw, const int64_t* sample_offsets, int32_t batch, int32_t feats, const double* xform,
                         const uint8_t* xform_flags, float ego_radius, const float* crop_xyxy, float* out, int64_t* count,
                         int64_t n_points, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CHECK_ARG(batch >= 1 && feats >= 3 && crop_xyxy && count && sample_offsets, "bad arguments");
  TMAE_CHECK_ARG(n_points >= 0 && n_points < ((int64_t)1 << 31), "n_points out of range");
  TMAE_CHECK_ARG((xform == nullptr) == (xform_flags == nullptr), "xform and xform_flags go together");
  TMAE_CHECK_ARG(workspace_bytes >= tmae_assemble_frames_workspace_bytes(n_points), "workspace too small");
  if (n_points == 0) { TMAE_CUDA(cudaMemsetAsync(count, 0, sizeof(int64_t), s)); return 0; }
  TMAE_CHECK_ARG(raw && out && workspace, "null buffer");
  Workspace ws(workspace, workspace_bytes);
  const int blocks = cdiv(n_points, kAsmTile);
  int* block_count = ws.take<int>(blocks);
  AsmArgs a{raw, sample_offsets, batch, feats, xform, xform_flags, (double)ego_radius, (double)crop_xyxy[0], (double)crop_xyxy[1],
            (double)crop_xyxy[2], (double)crop_xyxy[3], n_points};
  ProfScope prof("assemble_frames", 0, 4.0 * n_points * (2 * feats + 1), s);
  asm_count_kernel<<<blocks, kAsmThreads, 0, s>>>(a, block_count);
  asm_write_kernel<<<blocks, kAsmThreads, 0, s>>>(a, block_count, out, count);
  TMAE_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
