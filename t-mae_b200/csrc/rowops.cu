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

// ------------------------------------------------------------------ LayerNorm(x + res), vectorised (C = 128 * VPL)
// One warp per row, a lane owns VPL float4 column chunks (coalesced 512-byte row segments), two rows in flight per warp.
// The backward keeps its dgamma / dbeta partial sums in registers over a grid-stride walk of the rows, reduces them
// across the block's warps in shared memory and issues 2C atomics per BLOCK (the scalar version issued them per warp).
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float sum4(const float4& a) { return (a.x + a.y) + (a.z + a.w); }

template <int VPL>
__global__ void __launch_bounds__(256) add_ln_fwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                             const uint8_t* __restrict__ rowmask, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ y,
                                                             float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, float eps) {
  constexpr int C = VPL * 128;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 g[VPL], b[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) { g[k] = ld4(gamma + k * 128 + lane * 4); b[k] = ld4(beta + k * 128 + lane * 4); }
  for (int64_t r0 = warp * 2; r0 < rows; r0 += nwarps * 2) {
    float4 v[2][VPL];
    bool ok[2], use[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t r = r0 + u;
      ok[u] = r < rows;
      use[u] = ok[u] && res && (!rowmask || rowmask[r]);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        v[u][k] = ok[u] ? ld4(x + r * C + k * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (use[u]) {
          const float4 t = ld4(res + r * C + k * 128 + lane * 4);
          v[u][k].x += t.x; v[u][k].y += t.y; v[u][k].z += t.z; v[u][k].w += t.w;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;  // warp-uniform
      const int64_t r = r0 + u;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) s += sum4(v[u][k]);
      const float mean = warp_sum(s) * (1.f / C);
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const float dx = v[u][k].x - mean, dy = v[u][k].y - mean, dz = v[u][k].z - mean, dw = v[u][k].w - mean;
        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
      const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        float4 o;
        o.x = (v[u][k].x - mean) * rstd * g[k].x + b[k].x;
        o.y = (v[u][k].y - mean) * rstd * g[k].y + b[k].y;
        o.z = (v[u][k].z - mean) * rstd * g[k].z + b[k].z;
        o.w = (v[u][k].w - mean) * rstd * g[k].w + b[k].w;
        *reinterpret_cast<float4*>(y + r * C + k * 128 + lane * 4) = o;
      }
      if (lane == 0) {
        if (mean_out) mean_out[r] = mean;
        if (rstd_out) rstd_out[r] = rstd;
      }
    }
  }
}

// COLSUM: also accumulates the column sums of the kernel's own output (of dres when it is written, else of dv) into
// dcol -- the bias gradient of the linear layer that produced `res`, which would otherwise re-read the output from HBM.
int g_ln_bwd_cap = 6;   // blocks per SM of the vectorised LayerNorm backward (each block ends with 2-3 C atomics); tmae_set_option("ln_bwd_cap", n)

template <int VPL, bool COLSUM>
__global__ void __launch_bounds__(256) add_ln_bwd_vec_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                             const float* __restrict__ res, const uint8_t* __restrict__ rowmask,
                                                             const float* __restrict__ gamma, const float* __restrict__ mean_in,
                                                             const float* __restrict__ rstd_in, float* __restrict__ dv,
                                                             float* __restrict__ dres, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                             float* __restrict__ dcol, int64_t rows) {
  constexpr int C = VPL * 128;
  __shared__ float red[COLSUM ? 3 : 2][8][C];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 g[VPL], dg[VPL], db[VPL], dc[COLSUM ? VPL : 1];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    g[k] = ld4(gamma + k * 128 + lane * 4);
    dg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (COLSUM) dc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t r0 = warp * 2; r0 < rows; r0 += nwarps * 2) {
    float4 v[2][VPL], d[2][VPL];
    bool ok[2], use[2];
    float mean[2], rstd[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t r = r0 + u;
      ok[u] = r < rows;
      use[u] = ok[u] && res && (!rowmask || rowmask[r]);
      mean[u] = ok[u] ? mean_in[r] : 0.f;
      rstd[u] = ok[u] ? rstd_in[r] : 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int64_t off = r * C + k * 128 + lane * 4;
        v[u][k] = ok[u] ? ld4(x + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        d[u][k] = ok[u] ? ld4(dy + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (use[u]) {
          const float4 t = ld4(res + off);
          v[u][k].x += t.x; v[u][k].y += t.y; v[u][k].z += t.z; v[u][k].w += t.w;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;  // warp-uniform
      const int64_t r = r0 + u;
      float s1 = 0.f, s2 = 0.f;
      float4 xh[VPL], gg[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        xh[k].x = (v[u][k].x - mean[u]) * rstd[u]; xh[k].y = (v[u][k].y - mean[u]) * rstd[u];
        xh[k].z = (v[u][k].z - mean[u]) * rstd[u]; xh[k].w = (v[u][k].w - mean[u]) * rstd[u];
        gg[k].x = d[u][k].x * g[k].x; gg[k].y = d[u][k].y * g[k].y; gg[k].z = d[u][k].z * g[k].z; gg[k].w = d[u][k].w * g[k].w;
        dg[k].x = fmaf(d[u][k].x, xh[k].x, dg[k].x); dg[k].y = fmaf(d[u][k].y, xh[k].y, dg[k].y);
        dg[k].z = fmaf(d[u][k].z, xh[k].z, dg[k].z); dg[k].w = fmaf(d[u][k].w, xh[k].w, dg[k].w);
        db[k].x += d[u][k].x; db[k].y += d[u][k].y; db[k].z += d[u][k].z; db[k].w += d[u][k].w;
        s1 += sum4(gg[k]);
        s2 += (gg[k].x * xh[k].x + gg[k].y * xh[k].y) + (gg[k].z * xh[k].z + gg[k].w * xh[k].w);
      }
      s1 = warp_sum(s1) * (1.f / C);
      s2 = warp_sum(s2) * (1.f / C);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        float4 o;
        o.x = rstd[u] * (gg[k].x - s1 - xh[k].x * s2); o.y = rstd[u] * (gg[k].y - s1 - xh[k].y * s2);
        o.z = rstd[u] * (gg[k].z - s1 - xh[k].z * s2); o.w = rstd[u] * (gg[k].w - s1 - xh[k].w * s2);
        const int64_t off = r * C + k * 128 + lane * 4;
        *reinterpret_cast<float4*>(dv + off) = o;
        if (dres) *reinterpret_cast<float4*>(dres + off) = use[u] ? o : make_float4(0.f, 0.f, 0.f, 0.f);
        if (COLSUM && (use[u] || !dres)) { dc[k].x += o.x; dc[k].y += o.y; dc[k].z += o.z; dc[k].w += o.w; }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    *reinterpret_cast<float4*>(&red[0][wib][k * 128 + lane * 4]) = dg[k];
    *reinterpret_cast<float4*>(&red[1][wib][k * 128 + lane * 4]) = db[k];
    if (COLSUM) *reinterpret_cast<float4*>(&red[2][wib][k * 128 + lane * 4]) = dc[k];
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

// C % 128 == 0: a lane owns 4 adjacent channels (one 512-byte row = one coalesced warp access), the point loop is
// unrolled by 4 so four independent row loads are in flight; same comparison sequence as the scalar kernel.
__global__ void __launch_bounds__(256) segmax_fwd_vec_kernel(const float* __restrict__ x, const int* __restrict__ offset,
                                                             const int* __restrict__ order, int64_t m, int C, float* __restrict__ out,
                                                             int* __restrict__ arg) {
  const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= m) return;
  const int beg = offset[v], end = offset[v + 1];
  for (int c = lane * 4; c < C; c += 128) {
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int bi[4] = {-1, -1, -1, -1};
    auto take = [&](const float4& t, int p) {
      const float val[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (val[j] > best[j] || bi[j] < 0) { best[j] = val[j]; bi[j] = p; }
    };
    int k = beg;
    for (; k + 4 <= end; k += 4) {
      const int p0 = order[k], p1 = order[k + 1], p2 = order[k + 2], p3 = order[k + 3];
      const float4 t0 = ld4(x + (int64_t)p0 * C + c), t1 = ld4(x + (int64_t)p1 * C + c);
      const float4 t2 = ld4(x + (int64_t)p2 * C + c), t3 = ld4(x + (int64_t)p3 * C + c);
      take(t0, p0); take(t1, p1); take(t2, p2); take(t3, p3);
    }
    for (; k < end; ++k) {
      const int p = order[k];
      take(ld4(x + (int64_t)p * C + c), p);
    }
    *reinterpret_cast<float4*>(out + v * C + c) = make_float4(best[0], best[1], best[2], best[3]);
    *reinterpret_cast<int4*>(arg + v * C + c) = make_int4(bi[0], bi[1], bi[2], bi[3]);
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
  int64_t vb = (rows + 15) / 16;  // vectorised kernels: 8 warps x 2 rows per block iteration
  int vgrid = (int)(vb < (int64_t)kNumSMs * 8 ? vb : (int64_t)kNumSMs * 8);
  ProfScope prof("add_layernorm_fwd", 0, 4.0 * rows * c * (res ? 3 : 2), s);
  switch (c) {
    case 64: add_ln_fwd_kernel<2><<<grid, 256, 0, s>>>(x, res, rowmask, gamma, beta, y, mean, rstd, rows, eps); break;
    case 128: add_ln_fwd_vec_kernel<1><<<vgrid, 256, 0, s>>>(x, res, rowmask, gamma, beta, y, mean, rstd, rows, eps); break;
    case 256: add_ln_fwd_vec_kernel<2><<<vgrid, 256, 0, s>>>(x, res, rowmask, gamma, beta, y, mean, rstd, rows, eps); break;
    case 512: add_ln_fwd_vec_kernel<4><<<vgrid, 256, 0, s>>>(x, res, rowmask, gamma, beta, y, mean, rstd, rows, eps); break;
    default: set_error("tmae_add_layernorm_fwd: channels must be 64/128/256/512"); return TMAE_ERR_UNSUPPORTED;
  }
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_add_layernorm_bwd(const float* dy, const float* x, const float* res, const uint8_t* rowmask, const float* gamma,
                           const float* mean, const float* rstd, float* dv, float* dres, float* dgamma, float* dbeta, int64_t rows,
                           int32_t c, void* stream) {
  return tmae_add_layernorm_bwd_colsum(dy, x, res, rowmask, gamma, mean, rstd, dv, dres, dgamma, dbeta, nullptr, rows, c, stream);
}

int tmae_add_layernorm_bwd_colsum(const float* dy, const float* x, const float* res, const uint8_t* rowmask, const float* gamma,
                                  const float* mean, const float* rstd, float* dv, float* dres, float* dgamma, float* dbeta, float* dcolsum,
                                  int64_t rows, int32_t c, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CUDA(cudaMemsetAsync(dgamma, 0, c * sizeof(float), s));
  TMAE_CUDA(cudaMemsetAsync(dbeta, 0, c * sizeof(float), s));
  if (dcolsum) TMAE_CUDA(cudaMemsetAsync(dcolsum, 0, c * sizeof(float), s));
  if (rows <= 0) return 0;
  if (dcolsum && c != 128 && c != 256) {   // the fused column sum exists for the vectorised 128 / 256-channel kernels
    int rc = tmae_add_layernorm_bwd_colsum(dy, x, res, rowmask, gamma, mean, rstd, dv, dres, dgamma, dbeta, nullptr, rows, c, stream);
    return rc ? rc : tmae_colsum(dres ? dres : dv, dcolsum, rows, c, stream);
  }
  int rpw = 16;
  int64_t warps = (rows + rpw - 1) / rpw;
  int grid = cdiv(warps * 32, 256);
  ProfScope prof("add_layernorm_bwd", 0, 4.0 * rows * c * (res ? 4 : 3) + (dres ? 4.0 * rows * c : 0), s);
  int64_t vb = (rows + 63) / 64;  // vectorised kernels: every warp walks >= 8 rows so the per-block reduction amortises
  int vgrid = (int)(vb < (int64_t)kNumSMs * g_ln_bwd_cap ? vb : (int64_t)kNumSMs * g_ln_bwd_cap);
  switch (c) {
    case 64: add_ln_bwd_kernel<2><<<grid, 256, 0, s>>>(dy, x, res, rowmask, gamma, mean, rstd, dv, dres, dgamma, dbeta, rows, rpw); break;
    case 128:
      if (dcolsum) add_ln_bwd_vec_kernel<1, true><<<vgrid, 256, 0, s>>>(dy, x, res, rowmask, gamma, mean, rstd, dv, dres, dgamma, dbeta, dcolsum, rows);
      else add_ln_bwd_vec_kernel<1, false><<<vgrid, 256, 0, s>>>(dy, x, res, rowmask, gamma, mean, rstd, dv, dres, dgamma, dbeta, nullptr, rows);
      break;
    case 256:
      if (dcolsum) add_ln_bwd_vec_kernel<2, true><<<vgrid, 256, 0, s>>>(dy, x, res, rowmask, gamma, mean, rstd, dv, dres, dgamma, dbeta, dcolsum, rows);
      else add_ln_bwd_vec_kernel<2, false><<<vgrid, 256, 0, s>>>(dy, x, res, rowmask, gamma, mean, rstd, dv, dres, dgamma, dbeta, nullptr, rows);
      break;
    case 512: add_ln_bwd_vec_kernel<4, false><<<vgrid, 256, 0, s>>>(dy, x, res, rowmask, gamma, mean, rstd, dv, dres, dgamma, dbeta, nullptr, rows); break;
    default: set_error("tmae_add_layernorm_bwd: channels must be 64/128/256/512"); return TMAE_ERR_UNSUPPORTED;
  }
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_segment_max_fwd(const float* x, const int32_t* voxel_offset, const int32_t* pt_order, int64_t n_voxels, int32_t c, float* out,
                         int32_t* argmax, void* stream) {
  if (n_voxels <= 0) return 0;
  if (c % 128 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)argmax & 15) == 0)
    segmax_fwd_vec_kernel<<<cdiv(n_voxels * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, voxel_offset, pt_order, n_voxels, c, out, argmax);
  else
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
