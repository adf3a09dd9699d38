// Row-wise kernels of the path: position-embedding add, residual + LayerNorm, BatchNorm1d(+ReLU)
// over voxel/point rows, per-voxel segment max, and the sparse <-> dense BEV moves.
// All are HBM-bound: one pass over the rows, coalesced along channels, warp per row where a row
// reduction is needed.  Replaces ATen elementwise/reduction kernels, torch_scatter.scatter_max
// (temporal_dyn_vfe.py:113) and SparseConvTensor.dense() (SiamWCA_MAE.py:235).
#include <cuda_bf16.h>

#include "common.cuh"

namespace tmae {

// ------------------------------------------------------------------ x + pos_lut[posidx]
__global__ void add_pos_kernel(const float4* __restrict__ x, const uint8_t* __restrict__ posidx, const float4* __restrict__ lut,
                               float4* __restrict__ y, int64_t rows, int c4) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * c4) return;
  int64_t r = t / c4;
  int c = (int)(t - r * c4);
  float4 a = x[t], b = lut[(int64_t)posidx[r] * c4 + c];
  y[t] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// ------------------------------------------------------------------ LayerNorm(x + res)
// one warp per row; C in {64..1024}, C % 32 == 0
template <int PER>
__global__ void add_ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ res, const uint8_t* __restrict__ rowmask,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y,
                                  float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, float eps) {
  int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= rows) return;
  constexpr int C = PER * 32;
  bool use_res = res && (!rowmask || rowmask[r]);
  float v[PER];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    int c = k * 32 + lane;
    v[k] = x[r * C + c];
    if (use_res) v[k] += res[r * C + c];
    s += v[k];
  }
  float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < PER; ++k) { float d = v[k] - mean; q += d * d; }
  float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    int c = k * 32 + lane;
    y[r * C + c] = (v[k] - mean) * rstd * gamma[c] + beta[c];
  }
  if (lane == 0) {
    if (mean_out) mean_out[r] = mean;
    if (rstd_out) rstd_out[r] = rstd;
  }
}

// backward: dv = rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat)); dgamma/dbeta partial sums per block
template <int PER>
__global__ void add_ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ res,
                                  const uint8_t* __restrict__ rowmask, const float* __restrict__ gamma,
                                  const float* __restrict__ mean_in, const float* __restrict__ rstd_in, float* __restrict__ dv,
                                  float* __restrict__ dres, float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, int rows_per_warp) {
  constexpr int C = PER * 32;
  int lane = threadIdx.x & 31;
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float dg[PER], db[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) dg[k] = db[k] = 0.f;
  for (int it = 0; it < rows_per_warp; ++it) {
    int64_t r = w * rows_per_warp + it;
    if (r >= rows) break;
    bool use_res = res && (!rowmask || rowmask[r]);
    float mean = mean_in[r], rstd = rstd_in[r];
    float xh[PER], g[PER];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      int c = k * 32 + lane;
      float v = x[r * C + c];
      if (use_res) v += res[r * C + c];
      xh[k] = (v - mean) * rstd;
      float d = dy[r * C + c];
      g[k] = d * gamma[c];
      dg[k] += d * xh[k];
      db[k] += d;
      s1 += g[k];
      s2 += g[k] * xh[k];
    }
    s1 = warp_sum(s1) * (1.f / C);
    s2 = warp_sum(s2) * (1.f / C);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      float d = rstd * (g[k] - s1 - xh[k] * s2);
      dv[r * C + k * 32 + lane] = d;
      if (dres) dres[r * C + k * 32 + lane] = use_res ? d : 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    atomicAdd(dgamma + k * 32 + lane, dg[k]);
    atomicAdd(dbeta + k * 32 + lane, db[k]);
  }
}

// ------------------------------------------------------------------ per-voxel max over its points (CSR)
// warp per voxel; lanes stride channels; first maximum in ascending point order wins the argmax.
__global__ void segmax_fwd_kernel(const float* __restrict__ x, const int* __restrict__ offset, const int* __restrict__ order,
                                  int64_t m, int C, float* __restrict__ out, int* __restrict__ arg) {
  int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (v >= m) return;
  int beg = offset[v], end = offset[v + 1];
  for (int c = lane; c < C; c += 32) {
    float best = -INFINITY;
    int bi = -1;
    for (int k = beg; k < end; ++k) {
      int p = order[k];
      float val = x[(int64_t)p * C + c];
      if (val > best || bi < 0) { best = val; bi = p; }
    }
    out[v * C + c] = best;
    arg[v * C + c] = bi;
  }
}

__global__ void segmax_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ arg, int64_t n, int C, float* __restrict__ dx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int p = arg[i];
  if (p >= 0) dx[(int64_t)p * C + (i % C)] = dout[i];
}

// ------------------------------------------------------------------ sparse rows <-> dense NHWC map
// mode 0: dense[b,y,x,:] = rows[i,:]   (scatter; sites are unique)
// mode 1: rows[i,:] = dense[b,y,x,:]   (gather)
__global__ void rows_dense_kernel(float4* __restrict__ rows, float4* __restrict__ dense, const int* __restrict__ idx, int64_t m,
                                  int c4, int Y, int X, int mode) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * c4) return;
  int64_t i = t / c4;
  int c = (int)(t - i * c4);
  const int* p = idx + i * 3;
  int64_t d = (((int64_t)p[0] * Y + p[1]) * X + p[2]) * c4 + c;
  if (mode == 0) dense[d] = rows[t];
  else rows[t] = dense[d];
}

// bf16 dense map variant (the cuDNN decoder runs in bf16 channels-last): rows stay fp32
__global__ void rows_dense_bf16_kernel(float4* __restrict__ rows, uint2* __restrict__ dense, const int* __restrict__ idx, int64_t m,
                                       int c4, int Y, int X, int mode) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * c4) return;
  int64_t i = t / c4;
  int c = (int)(t - i * c4);
  const int* p = idx + i * 3;
  int64_t d = (((int64_t)p[0] * Y + p[1]) * X + p[2]) * c4 + c;
  if (mode == 0) {
    float4 v = rows[t];
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    dense[d] = o;
  } else {
    uint2 o = dense[d];
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&o.x), b = *reinterpret_cast<__nv_bfloat162*>(&o.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    rows[t] = make_float4(fa.x, fa.y, fb.x, fb.y);
  }
}

// rows[i,:] (+)= src[sel[i],:]  /  dst[sel[i],:] = rows[i,:]   (row gather / scatter by index list)
__global__ void rows_index_kernel(float4* __restrict__ a, float4* __restrict__ b, const int* __restrict__ sel, int64_t m, int c4,
                                  int mode) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * c4) return;
  int64_t i = t / c4;
  int c = (int)(t - i * c4);
  int64_t j = sel[i];
  if (mode == 0) a[t] = b[j * c4 + c];   // gather: a = b[sel]
  else b[j * c4 + c] = a[t];             // scatter: b[sel] = a
}

}  // namespace tmae

using namespace tmae;

extern "C" {

int tmae_add_pos(const float* x, const uint8_t* posidx, const float* lut, float* y, int64_t rows, int32_t c, void* stream) {
  TMAE_CHECK_ARG(c % 4 == 0, "channels must be a multiple of 4");
  if (rows <= 0) return 0;
  int c4 = c / 4;
  ProfScope prof("add_pos", 0, 8.0 * rows * c, (cudaStream_t)stream);
  add_pos_kernel<<<cdiv(rows * c4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)x, posidx, (const float4*)lut, (float4*)y, rows, c4);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_add_layernorm_fwd(const float* x, const float* res, const uint8_t* rowmask, const float* gamma, const float* beta, float* y,
                           float* mean, float* rstd, int64_t rows, int32_t c, float eps, void* stream) {
  if (rows <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  int grid = cdiv(rows * 32, 256);
  ProfScope prof("add_layernorm_fwd", 0, 4.0 * rows * c * (res ? 3 : 2), s);
  switch (c) {
    case 64: add_ln_fwd_kernel<2><<<grid, 256, 0, s>>>(x, res, rowmask, gamma, beta, y, mean, rstd, rows, eps); break;
    case 128: add_ln_fwd_kernel<4><<<grid, 256, 0, s>>>(x, res, rowmask, gamma, beta, y, mean, rstd, rows, eps); break;
    case 256: add_ln_fwd_kernel<8><<<grid, 256, 0, s>>>(x, res, rowmask, gamma, beta, y, mean, rstd, rows, eps); break;
    case 512: add_ln_fwd_kernel<16><<<grid, 256, 0, s>>>(x, res, rowmask, gamma, beta, y, mean, rstd, rows, eps); break;
    default: set_error("tmae_add_layernorm_fwd: channels must be 64/128/256/512"); return TMAE_ERR_UNSUPPORTED;
  }
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_add_layernorm_bwd(const float* dy, const float* x, const float* res, const uint8_t* rowmask, const float* gamma,
                           const float* mean, const float* rstd, float* dv, float* dres, float* dgamma, float* dbeta, int64_t rows,
                           int32_t c, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CUDA(cudaMemsetAsync(dgamma, 0, c * sizeof(float), s));
  TMAE_CUDA(cudaMemsetAsync(dbeta, 0, c * sizeof(float), s));
  if (rows <= 0) return 0;
  int rpw = 16;
  int64_t warps = (rows + rpw - 1) / rpw;
  int grid = cdiv(warps * 32, 256);
  ProfScope prof("add_layernorm_bwd", 0, 4.0 * rows * c * (res ? 4 : 3) + (dres ? 4.0 * rows * c : 0), s);
  switch (c) {
    case 64: add_ln_bwd_kernel<2><<<grid, 256, 0, s>>>(dy, x, res, rowmask, gamma, mean, rstd, dv, dres, dgamma, dbeta, rows, rpw); break;
    case 128: add_ln_bwd_kernel<4><<<grid, 256, 0, s>>>(dy, x, res, rowmask, gamma, mean, rstd, dv, dres, dgamma, dbeta, rows, rpw); break;
    case 256: add_ln_bwd_kernel<8><<<grid, 256, 0, s>>>(dy, x, res, rowmask, gamma, mean, rstd, dv, dres, dgamma, dbeta, rows, rpw); break;
    case 512: add_ln_bwd_kernel<16><<<grid, 256, 0, s>>>(dy, x, res, rowmask, gamma, mean, rstd, dv, dres, dgamma, dbeta, rows, rpw); break;
    default: set_error("tmae_add_layernorm_bwd: channels must be 64/128/256/512"); return TMAE_ERR_UNSUPPORTED;
  }
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_segment_max_fwd(const float* x, const int32_t* voxel_offset, const int32_t* pt_order, int64_t n_voxels, int32_t c, float* out,
                         int32_t* argmax, void* stream) {
  if (n_voxels <= 0) return 0;
  segmax_fwd_kernel<<<cdiv(n_voxels * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, voxel_offset, pt_order, n_voxels, c, out, argmax);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_segment_max_bwd(const float* dout, const int32_t* argmax, int64_t n_voxels, int32_t c, float* dx, int64_t n_points,
                         void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CUDA(cudaMemsetAsync(dx, 0, (size_t)n_points * c * sizeof(float), s));
  if (n_voxels <= 0) return 0;
  segmax_bwd_kernel<<<cdiv(n_voxels * c, 256), 256, 0, s>>>(dout, argmax, n_voxels * c, c, dx);
  TMAE_CHECK_LAUNCH();
  return 0;
}

/* dense (B,Y,X,C) channels-last map from sparse rows (zero-fills first) -- SparseConvTensor.dense() */
int tmae_densify_nhwc(const float* rows, const int32_t* indices, int64_t m, int32_t c, int32_t batch, int32_t y, int32_t x,
                      float* dense, int32_t zero_fill, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CHECK_ARG(c % 4 == 0, "channels must be a multiple of 4");
  if (zero_fill) TMAE_CUDA(cudaMemsetAsync(dense, 0, (size_t)batch * y * x * c * sizeof(float), s));
  if (m <= 0) return 0;
  rows_dense_kernel<<<cdiv(m * (c / 4), 256), 256, 0, s>>>((float4*)rows, (float4*)dense, indices, m, c / 4, y, x, 0);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_gather_nhwc(const float* dense, const int32_t* indices, int64_t m, int32_t c, int32_t y, int32_t x, float* rows, void* stream) {
  TMAE_CHECK_ARG(c % 4 == 0, "channels must be a multiple of 4");
  if (m <= 0) return 0;
  rows_dense_kernel<<<cdiv(m * (c / 4), 256), 256, 0, (cudaStream_t)stream>>>((float4*)rows, (float4*)dense, indices, m, c / 4, y, x, 1);
  TMAE_CHECK_LAUNCH();
  return 0;
}

/* bf16 dense map variants (dense is (B,Y,X,C) bf16) */
int tmae_densify_nhwc_bf16(const float* rows, const int32_t* indices, int64_t m, int32_t c, int32_t batch, int32_t y, int32_t x,
                           void* dense, int32_t zero_fill, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CHECK_ARG(c % 4 == 0, "channels must be a multiple of 4");
  if (zero_fill) TMAE_CUDA(cudaMemsetAsync(dense, 0, (size_t)batch * y * x * c * 2, s));
  if (m <= 0) return 0;
  ProfScope prof("densify_bf16", 0, 6.0 * m * c + (zero_fill ? 2.0 * batch * y * x * c : 0), s);
  rows_dense_bf16_kernel<<<cdiv(m * (c / 4), 256), 256, 0, s>>>((float4*)rows, (uint2*)dense, indices, m, c / 4, y, x, 0);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_gather_nhwc_bf16(const void* dense, const int32_t* indices, int64_t m, int32_t c, int32_t y, int32_t x, float* rows, void* stream) {
  TMAE_CHECK_ARG(c % 4 == 0, "channels must be a multiple of 4");
  if (m <= 0) return 0;
  ProfScope prof("gather_bf16", 0, 6.0 * m * c, (cudaStream_t)stream);
  rows_dense_bf16_kernel<<<cdiv(m * (c / 4), 256), 256, 0, (cudaStream_t)stream>>>((float4*)rows, (uint2*)dense, indices, m, c / 4, y, x, 1);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_gather_rows(const float* src, const int32_t* sel, int64_t m, int32_t c, float* out, void* stream) {
  TMAE_CHECK_ARG(c % 4 == 0, "channels must be a multiple of 4");
  if (m <= 0) return 0;
  rows_index_kernel<<<cdiv(m * (c / 4), 256), 256, 0, (cudaStream_t)stream>>>((float4*)out, (float4*)src, sel, m, c / 4, 0);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_scatter_rows(const float* rows, const int32_t* sel, int64_t m, int32_t c, float* dst, void* stream) {
  TMAE_CHECK_ARG(c % 4 == 0, "channels must be a multiple of 4");
  if (m <= 0) return 0;
  rows_index_kernel<<<cdiv(m * (c / 4), 256), 256, 0, (cudaStream_t)stream>>>((float4*)rows, (float4*)dst, sel, m, c / 4, 1);
  TMAE_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
