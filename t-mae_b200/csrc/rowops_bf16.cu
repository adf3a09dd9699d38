// Row-wise kernels of the bf16-storage mode: dtype casts, LayerNorm backward on bf16 rows, column sums of bf16 gradients
// (bias gradients, the binned sum behind the position-table gradient), 16-bit row <-> dense BEV moves, and the tap transpose
// of a sparse-conv weight straight into bf16.  All HBM-bound: one pass, 16-byte accesses, fp32 arithmetic and statistics.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace tmae {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// ------------------------------------------------------------------ casts (n % 8 == 0 fast path + scalar tail)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + i)), b = __ldg(reinterpret_cast<const float4*>(src + i + 4));
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    *reinterpret_cast<uint4*>(dst + i) = pack8(f);
  } else {
    for (int64_t j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
  }
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + i)), f);
    *reinterpret_cast<float4*>(dst + i) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(dst + i + 4) = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    for (int64_t j = i; j < n; ++j) dst[j] = __bfloat162float(src[j]);
  }
}
// many tensors in one launch (the layer weights' bf16 shadows after an optimizer step): segment table on the device
struct CastSeg { const float* src; bf16* dst; int64_t n; };
__global__ void cast_multi_kernel(const CastSeg* __restrict__ segs, int n_seg) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: see tc_common.cuh pdl_trigger
  const CastSeg sg = segs[blockIdx.y];
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < sg.n; i += (int64_t)gridDim.x * blockDim.x * 8) {
    if (i + 8 <= sg.n) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(sg.src + i)), b = __ldg(reinterpret_cast<const float4*>(sg.src + i + 4));
      const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      *reinterpret_cast<uint4*>(sg.dst + i) = pack8(f);
    } else {
      for (int64_t j = i; j < sg.n; ++j) sg.dst[j] = __float2bfloat16_rn(sg.src[j]);
    }
  }
}

// ------------------------------------------------------------------ LayerNorm backward on bf16 rows
// v = the pre-norm sum the forward epilogue saved; dv = rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat)).
// A lane owns one 16-byte chunk (8 channels) of a row: C = 128 -> two rows per warp pass, C = 256 -> one.  dgamma / dbeta
// (and COLSUM: the column sums of dres when it is written, else of dv = the bias gradient of the producing linear layer)
// stay in registers over a grid-stride walk, are reduced across the block in shared memory, 2-3 C atomics per block.
template <int C, bool COLSUM>
__global__ void __launch_bounds__(256) ln_bwd_bf16_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ v, const uint8_t* __restrict__ rowmask,
                                                          const float* __restrict__ gamma, const float* __restrict__ mean_in,
                                                          const float* __restrict__ rstd_in, bf16* __restrict__ dv, bf16* __restrict__ dres,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dcol, int64_t rows) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: see tc_common.cuh pdl_trigger
  constexpr int LPR = C / 8, RPW = 32 / LPR, UNR = 2;
  __shared__ float red[COLSUM ? 3 : 2][8][C];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int sub = lane / LPR, cl = lane % LPR;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float g[8], dg[8], db[8], dc[8];
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(gamma + cl * 8)), b = __ldg(reinterpret_cast<const float4*>(gamma + cl * 8 + 4));
    g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) dg[j] = db[j] = dc[j] = 0.f;
  for (int64_t r0 = warp * (RPW * UNR); r0 < rows; r0 += nwarps * (RPW * UNR)) {
    uint4 ud[UNR], uv[UNR];
    bool ok[UNR], use[UNR];
    float mean[UNR], rstd[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t r = r0 + u * RPW + sub;
      ok[u] = r < rows;
      use[u] = ok[u] && (!rowmask || rowmask[r]);
      mean[u] = ok[u] ? mean_in[r] : 0.f;
      rstd[u] = ok[u] ? rstd_in[r] : 0.f;
      ud[u] = ok[u] ? __ldg(reinterpret_cast<const uint4*>(dy + r * C + cl * 8)) : make_uint4(0, 0, 0, 0);
      uv[u] = ok[u] ? __ldg(reinterpret_cast<const uint4*>(v + r * C + cl * 8)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t r = r0 + u * RPW + sub;
      float d[8], x[8], gg[8];
      unpack8(ud[u], d);
      unpack8(uv[u], x);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        x[j] = (x[j] - mean[u]) * rstd[u];
        gg[j] = d[j] * g[j];
        dg[j] = fmaf(d[j], x[j], dg[j]);
        db[j] += d[j];
        s1 += gg[j];
        s2 = fmaf(gg[j], x[j], s2);
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      s1 *= 1.f / C;
      s2 *= 1.f / C;
      float o8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o8[j] = rstd[u] * (gg[j] - s1 - x[j] * s2);
      if (ok[u]) {
        const uint4 po = pack8(o8);
        *reinterpret_cast<uint4*>(dv + r * C + cl * 8) = po;
        if (dres) *reinterpret_cast<uint4*>(dres + r * C + cl * 8) = use[u] ? po : make_uint4(0, 0, 0, 0);
        if (COLSUM && (use[u] || !dres)) {
#pragma unroll
          for (int j = 0; j < 8; ++j) dc[j] += o8[j];
        }
      }
    }
  }
  if (RPW == 2) {   // the two half-warps hold partial sums of the same columns
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dg[j] += __shfl_xor_sync(0xffffffffu, dg[j], 16);
      db[j] += __shfl_xor_sync(0xffffffffu, db[j], 16);
      if (COLSUM) dc[j] += __shfl_xor_sync(0xffffffffu, dc[j], 16);
    }
  }
  if (sub == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[0][wib][cl * 8 + j] = dg[j];
      red[1][wib][cl * 8 + j] = db[j];
      if (COLSUM) red[2][wib][cl * 8 + j] = dc[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = 0.f, b = 0.f, e = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += red[0][w][c]; b += red[1][w][c]; if (COLSUM) e += red[2][w][c]; }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
    if (COLSUM) atomicAdd(dcol + c, e);
  }
}

// ------------------------------------------------------------------ column sums of a bf16 matrix (bias gradients)
// thread = 16-byte column chunk; a block walks `rpb` rows; partial sums -> atomics.  BINNED: out[bin[row]][col] (64 bins:
// the position-table gradient: rows grouped by their cell in the 8x8 window)
template <bool BINNED>
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* __restrict__ x, const uint8_t* __restrict__ bin, float* __restrict__ out,
                                                          int64_t rows, int n, int rpb) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: see tc_common.cuh pdl_trigger
  // block = 16 column chunks (128 columns) x 16 row lanes
  const int chunks = n / 8;
  const int cx = threadIdx.x & 15, cl = cx + 16 * blockIdx.x;
  const int rl = threadIdx.x >> 4;
  const int64_t r_beg = (int64_t)blockIdx.y * rpb, r_end = r_beg + rpb < rows ? r_beg + rpb : rows;
  __shared__ float sacc[BINNED ? 64 : 16][128 + 1];
  for (int i = threadIdx.x; i < (BINNED ? 64 : 16) * 129; i += 256) (&sacc[0][0])[i] = 0.f;
  __syncthreads();
  if (cl < chunks) {
    if (!BINNED) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      int64_t r = r_beg + rl;
      for (; r + 48 < r_end; r += 64) {     // four independent 16-byte loads in flight per thread
        uint4 u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) u[i] = __ldg(reinterpret_cast<const uint4*>(x + (r + 16 * i) * n + cl * 8));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float f[8];
          unpack8(u[i], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
      }
      for (; r < r_end; r += 16) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + r * n + cl * 8)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sacc[rl][cx * 8 + j] = acc[j];
    } else {
      for (int64_t r = r_beg + rl; r < r_end; r += 16) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + r * n + cl * 8)), f);
        const int b = bin[r] & 63;
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&sacc[b][cx * 8 + j], f[j]);
      }
    }
  }
  __syncthreads();
  const int ncol = min(128, n - (int)blockIdx.x * 128);
  if (!BINNED) {
    for (int c = threadIdx.x; c < ncol; c += 256) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < 16; ++w) a += sacc[w][c];
      atomicAdd(out + blockIdx.x * 128 + c, a);
    }
  } else {
    for (int i = threadIdx.x; i < 64 * ncol; i += 256) {
      const int b = i / ncol, c = i % ncol;
      const float v = sacc[b][c];
      if (v != 0.f) atomicAdd(out + (int64_t)b * n + blockIdx.x * 128 + c, v);
    }
  }
}

// ------------------------------------------------------------------ sparse-conv weight (cout, taps, cin) fp32 -> (cin, taps, cout) bf16
__global__ void transpose_taps_bf16_kernel(const float* __restrict__ w, bf16* __restrict__ wt, int cout, int taps, int cin, int flip) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = (int64_t)cout * taps * cin;
  if (i >= n) return;
  const int co = (int)(i % cout);
  const int t = (int)((i / cout) % taps);
  const int ci = (int)(i / ((int64_t)cout * taps));
  const int ts = flip ? taps - 1 - t : t;
  wt[i] = __float2bfloat16_rn(w[((int64_t)co * taps + ts) * cin + ci]);
}

// ------------------------------------------------------------------ rows (m, C) 16-bit <-> dense (B,Y,X,C) 16-bit, 16-byte chunks
__global__ void rows_dense_b16_kernel(uint4* __restrict__ rows, uint4* __restrict__ dense, const int* __restrict__ idx, int64_t m, int c8, int Y, int X,
                                      int mode) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * c8) return;
  const int64_t i = t / c8;
  const int c = (int)(t - i * c8);
  const int* p = idx + i * 3;
  const int64_t d = (((int64_t)p[0] * Y + p[1]) * X + p[2]) * c8 + c;
  if (mode == 0) dense[d] = rows[t];
  else rows[t] = dense[d];
}

// ------------------------------------------------------------------ one-hot rows of the 64 window cells (bf16): the second B operand
// of the in-projection weight-gradient GEMM (dy^T [x | onehot] = [dW | dtable^T])
__global__ void onehot64_bf16_kernel(const uint8_t* __restrict__ idx, uint4* __restrict__ out, int64_t m) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte chunk (8 cells) per thread
  if (t >= m * 8) return;
  const int64_t r = t >> 3;
  const int c = (int)(t & 7) * 8;
  const int p = idx[r] & 63;
  uint32_t w[4] = {0, 0, 0, 0};
  if (p >= c && p < c + 8) w[(p - c) >> 1] = ((p - c) & 1) ? 0x3f800000u : 0x00003f80u;   // bf16 1.0 = 0x3f80
  out[t] = make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------ attention bridge (bf16 storage <-> the fp32-I/O attention kernels)
// dst[r, c] = src[r, c] * (c < norm_cols ? scale[r, c / hd] : 1): the gradient of a unit-normalised q / k row, back through the
// 1 / |.| the projection epilogue applied (the attention kernels already project out the radial component).
__global__ void scale_cast_kernel(const float* __restrict__ src, const float* __restrict__ scale, bf16* __restrict__ dst, int64_t rows, int n,
                                  int norm_cols, int hd) {
  const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  const int c8 = n / 8;
  if (t >= rows * c8) return;
  const int64_t r = t / c8;
  const int c = (int)(t - r * c8) * 8;
  const float4 a = __ldg(reinterpret_cast<const float4*>(src + r * n + c)), b = __ldg(reinterpret_cast<const float4*>(src + r * n + c + 4));
  float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  if (c < norm_cols) {
    const float sc = scale[r * (norm_cols / hd) + c / hd];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= sc;
  }
  *reinterpret_cast<uint4*>(dst + r * n + c) = pack8(f);
}

}  // namespace tmae

using namespace tmae;

extern "C" {

int tmae_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (n <= 0) return 0;
  TMAE_CHECK_ARG((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "pointers must be 16-byte aligned");
  cast_f32_bf16_kernel<<<cdiv((n + 7) / 8, 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_cast_bf16_f32(const void* src, float* dst, int64_t n, void* stream) {
  if (n <= 0) return 0;
  TMAE_CHECK_ARG((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "pointers must be 16-byte aligned");
  cast_bf16_f32_kernel<<<cdiv((n + 7) / 8, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)src, dst, n);
  TMAE_CHECK_LAUNCH();
  return 0;
}

/* segs: DEVICE array of n_seg {const float* src; void* dst; int64_t n} records (24 bytes each): all of them in one launch */
int tmae_cast_f32_bf16_multi(const void* segs, int32_t n_seg, void* stream) {
  if (n_seg <= 0) return 0;
  dim3 grid(64, (unsigned)n_seg);
  cast_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const CastSeg*)segs, n_seg);
  TMAE_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"

// `zero` = false: the caller has already zero-filled the fp32 accumulation targets (the layer entry points clear all parameter
// gradients of a layer with ONE memset instead of one per target)
namespace tmae {
int bf16_layernorm_bwd_impl(const void* dy, const void* v, const uint8_t* rowmask, const float* gamma, const float* mean, const float* rstd,
                            void* dv, void* dres, float* dgamma, float* dbeta, float* dcolsum, int64_t rows, int32_t c, bool zero, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CHECK_ARG(c == 128 || c == 256, "channels must be 128 or 256");
  if (zero) {
    TMAE_CUDA(cudaMemsetAsync(dgamma, 0, c * sizeof(float), s));
    TMAE_CUDA(cudaMemsetAsync(dbeta, 0, c * sizeof(float), s));
    if (dcolsum) TMAE_CUDA(cudaMemsetAsync(dcolsum, 0, c * sizeof(float), s));
  }
  if (rows <= 0) return 0;
  ProfScope prof("bf16_layernorm_bwd", 0, 2.0 * rows * c * (3 + (dres ? 1 : 0)), s);
  int64_t vb = (rows + 63) / 64;
  int grid = (int)(vb < (int64_t)kNumSMs * 6 ? vb : (int64_t)kNumSMs * 6);
  const bf16 *pdy = (const bf16*)dy, *pv = (const bf16*)v;
  bf16 *pdv = (bf16*)dv, *pdr = (bf16*)dres;
  if (c == 128) {
    if (dcolsum) ln_bwd_bf16_kernel<128, true><<<grid, 256, 0, s>>>(pdy, pv, rowmask, gamma, mean, rstd, pdv, pdr, dgamma, dbeta, dcolsum, rows);
    else ln_bwd_bf16_kernel<128, false><<<grid, 256, 0, s>>>(pdy, pv, rowmask, gamma, mean, rstd, pdv, pdr, dgamma, dbeta, nullptr, rows);
  } else {
    if (dcolsum) ln_bwd_bf16_kernel<256, true><<<grid, 256, 0, s>>>(pdy, pv, rowmask, gamma, mean, rstd, pdv, pdr, dgamma, dbeta, dcolsum, rows);
    else ln_bwd_bf16_kernel<256, false><<<grid, 256, 0, s>>>(pdy, pv, rowmask, gamma, mean, rstd, pdv, pdr, dgamma, dbeta, nullptr, rows);
  }
  TMAE_CHECK_LAUNCH();
  return 0;
}

int bf16_colsum_impl(const void* x, float* out, int64_t rows, int32_t cols, bool zero, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CHECK_ARG(cols % 8 == 0, "cols must be a multiple of 8");
  if (zero) TMAE_CUDA(cudaMemsetAsync(out, 0, (size_t)cols * sizeof(float), s));
  if (rows <= 0) return 0;
  // rows per block: ~4 blocks per SM over the whole grid (134 blocks of 1024 rows left the 148 SMs under-filled: 1.5 TB/s)
  const int64_t col_blocks = cdiv(cols / 8, 16);
  int64_t want = cdiv(rows * col_blocks, (int64_t)kNumSMs * 4);
  const int rpb = (int)(want < 128 ? 128 : (want > 1024 ? 1024 : align_up(want, 64)));
  dim3 grid((unsigned)col_blocks, (unsigned)cdiv(rows, rpb));
  ProfScope prof("bf16_colsum", 0, 2.0 * rows * cols, s);
  colsum_bf16_kernel<false><<<grid, 256, 0, s>>>((const bf16*)x, nullptr, out, rows, cols, rpb);
  TMAE_CHECK_LAUNCH();
  return 0;
}
}  // namespace tmae

extern "C" {

int tmae_bf16_layernorm_bwd(const void* dy, const void* v, const uint8_t* rowmask, const float* gamma, const float* mean, const float* rstd,
                            void* dv, void* dres, float* dgamma, float* dbeta, float* dcolsum, int64_t rows, int32_t c, void* stream) {
  return tmae::bf16_layernorm_bwd_impl(dy, v, rowmask, gamma, mean, rstd, dv, dres, dgamma, dbeta, dcolsum, rows, c, true, stream);
}

int tmae_bf16_colsum(const void* x, float* out, int64_t rows, int32_t cols, void* stream) {
  return tmae::bf16_colsum_impl(x, out, rows, cols, true, stream);
}

/* dtable (64, n) fp32 = sum over rows with rowidx == p of dy[row, :] */
int tmae_bf16_binned_colsum(const void* dy, const uint8_t* rowidx, float* dtable, int64_t rows, int32_t n, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CHECK_ARG(n % 8 == 0, "n must be a multiple of 8");
  TMAE_CUDA(cudaMemsetAsync(dtable, 0, (size_t)64 * n * sizeof(float), s));
  if (rows <= 0) return 0;
  const int rpb = 2048;
  dim3 grid((unsigned)cdiv(n / 8, 16), (unsigned)cdiv(rows, rpb));
  ProfScope prof("bf16_binned_colsum", 0, 2.0 * rows * n, s);
  colsum_bf16_kernel<true><<<grid, 256, 0, s>>>((const bf16*)dy, rowidx, dtable, rows, n, rpb);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_transpose_taps_bf16(const float* w, void* wt, int32_t cout, int32_t taps, int32_t cin, int32_t flip, void* stream) {
  int64_t n = (int64_t)cout * taps * cin;
  if (n <= 0) return 0;
  transpose_taps_bf16_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(w, (bf16*)wt, cout, taps, cin, flip);
  TMAE_CHECK_LAUNCH();
  return 0;
}

/* rows (m, c) <-> dense (B,Y,X,c), both 16-bit elements (bf16): SparseConvTensor.dense() and the BEV gather in the bf16-storage mode */
int tmae_densify_nhwc_b16(const void* rows, const int32_t* indices, int64_t m, int32_t c, int32_t batch, int32_t y, int32_t x, void* dense,
                          int32_t zero_fill, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CHECK_ARG(c % 8 == 0, "channels must be a multiple of 8");
  if (zero_fill) TMAE_CUDA(cudaMemsetAsync(dense, 0, (size_t)batch * y * x * c * 2, s));
  if (m <= 0) return 0;
  ProfScope prof("densify_b16", 0, 4.0 * m * c + (zero_fill ? 2.0 * batch * y * x * c : 0), s);
  rows_dense_b16_kernel<<<cdiv(m * (c / 8), 256), 256, 0, s>>>((uint4*)rows, (uint4*)dense, indices, m, c / 8, y, x, 0);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_gather_nhwc_b16(const void* dense, const int32_t* indices, int64_t m, int32_t c, int32_t y, int32_t x, void* rows, void* stream) {
  TMAE_CHECK_ARG(c % 8 == 0, "channels must be a multiple of 8");
  if (m <= 0) return 0;
  ProfScope prof("gather_b16", 0, 4.0 * m * c, (cudaStream_t)stream);
  rows_dense_b16_kernel<<<cdiv(m * (c / 8), 256), 256, 0, (cudaStream_t)stream>>>((uint4*)rows, (uint4*)dense, indices, m, c / 8, y, x, 1);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_onehot64_bf16(const uint8_t* idx, void* out, int64_t m, void* stream) {
  if (m <= 0) return 0;
  onehot64_bf16_kernel<<<cdiv(m * 8, 256), 256, 0, (cudaStream_t)stream>>>(idx, (uint4*)out, m);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_scale_cast_bf16(const float* src, const float* scale, void* dst, int64_t rows, int32_t n, int32_t norm_cols, int32_t hd, void* stream) {
  TMAE_CHECK_ARG(n % 8 == 0 && hd % 8 == 0 && norm_cols % hd == 0, "n and hd must be multiples of 8");
  if (rows <= 0) return 0;
  scale_cast_kernel<<<cdiv(rows * (n / 8), 256), 256, 0, (cudaStream_t)stream>>>(src, scale, (bf16*)dst, rows, n, norm_cols, hd);
  TMAE_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
