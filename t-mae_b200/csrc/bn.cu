// BatchNorm (+ReLU) over the rows of a (rows, C) matrix, training and eval mode, forward and backward:
//   * BatchNorm1d after the VFE linears and the sparse convolutions (network_utils.py:31, spconv_utils.py:50-54),
//     fp32 rows;
//   * BatchNorm2d of the dense decoder (SiamWCA_MAE.py:79-115) on channels-last bf16 maps viewed as
//     (B*Y*X, C) rows, written straight into a column slice of the concatenated (rows, 384) buffer so that the
//     reference's torch.cat never runs.
// One templated implementation: element type fp32 or bf16 (statistics and arithmetic always fp32, column sums in
// double), row pitches given explicitly.  HBM-bound: every thread owns one 16-byte column chunk and walks rows with
// four independent loads in flight; a block reduces its partial column sums in shared memory and issues C double
// atomics.  The backward recomputes the ReLU mask from x (v > 0 <=> y > 0), so y is never read.
#include <cuda_bf16.h>

#include "common.cuh"

namespace tmae {

template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float* v) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static void store(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
  typedef float4 Raw;   // a load kept as it came from memory; unpacked where it is used (register pressure of the unrolled loops)
  __device__ static Raw load_raw(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ static Raw zero_raw() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ static void unpack(const Raw& t, float* v) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float* v) {
    uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ static void store(__nv_bfloat16* p, const float* v) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
  typedef uint4 Raw;
  __device__ static Raw load_raw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static Raw zero_raw() { return make_uint4(0u, 0u, 0u, 0u); }
  __device__ static void unpack(const Raw& t, float* v) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
};

constexpr int BN_THREADS = 256;
constexpr int BN_UNROLL = 4;

// column sums of f0(row) and f1(row) over all rows; MODE 0: (x, x^2) ; MODE 1: (dy', dy' * xhat) with dy' = dy * [v > 0]
template <typename T, int MODE>
__global__ void __launch_bounds__(BN_THREADS) bn_colsum_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t ldd,
                                                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                                                                int64_t rows, int C, double* __restrict__ s0, double* __restrict__ s1) {
  constexpr int V = Vec<T>::N;
  __shared__ float red[2][BN_THREADS][V + 1];
  const int tpr = C / V;               // threads per row
  const int rpb = BN_THREADS / tpr;    // rows per block iteration
  const int ch = (threadIdx.x % tpr) * V, rg = threadIdx.x / tpr;
  float a0[V], a1[V], m[V], sc[V], sh[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
  if (MODE == 1) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      m[j] = mean[ch + j];
      sc[j] = rstd[ch + j] * gamma[ch + j];
      sh[j] = (beta ? beta[ch + j] : 0.f) - m[j] * sc[j];   // same expressions as the forward: identical ReLU mask
    }
  }
  const int64_t stride = (int64_t)gridDim.x * rpb;
  for (int64_t r = (int64_t)blockIdx.x * rpb + rg; r < rows; r += stride * BN_UNROLL) {
    typename Vec<T>::Raw xr[BN_UNROLL], dr[BN_UNROLL];   // all loads of the iteration in flight, unpacked one row at a time
#pragma unroll
    for (int u = 0; u < BN_UNROLL; ++u) {
      const int64_t rr = r + u * stride;
      xr[u] = rr < rows ? Vec<T>::load_raw(x + rr * ldx + ch) : Vec<T>::zero_raw();
      if (MODE == 1) dr[u] = rr < rows ? Vec<T>::load_raw(dy + rr * ldd + ch) : Vec<T>::zero_raw();   // dy' = 0: the row adds nothing
    }
#pragma unroll
    for (int u = 0; u < BN_UNROLL; ++u) {
      float xv[V], dv[V];
      Vec<T>::unpack(xr[u], xv);
      if (MODE == 1) Vec<T>::unpack(dr[u], dv);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        if (MODE == 0) {
          a0[j] += xv[j];
          a1[j] = fmaf(xv[j], xv[j], a1[j]);
        } else {
          float d = dv[j];
          if (relu && !(fmaf(xv[j], sc[j], sh[j]) > 0.f)) d = 0.f;
          a0[j] += d;
          a1[j] = fmaf(d, xv[j] - m[j], a1[j]);   // sum of dy' (x - mean); the common factor rstd is applied once per block below
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) { red[0][threadIdx.x][j] = a0[j]; red[1][threadIdx.x][j] = a1[j]; }
  __syncthreads();
  // thread c sums column c over the row groups
  for (int c = threadIdx.x; c < C; c += BN_THREADS) {
    const int t0 = c / V, j = c % V;
    float u0 = 0.f, u1 = 0.f;
    for (int g = 0; g < rpb; ++g) { u0 += red[0][g * tpr + t0][j]; u1 += red[1][g * tpr + t0][j]; }
    if (MODE == 1) u1 *= rstd[c];
    atomicAdd(s0 + c, (double)u0);
    atomicAdd(s1 + c, (double)u1);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, int64_t rows, int C, float eps,
                                   float momentum, float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double m = sum[c] / (double)rows;
  double var = sumsq[c] / (double)rows - m * m;
  if (var < 0) var = 0;
  mean[c] = (float)m;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    double unb = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
  }
}

// MODE 0: y = [relu](x sc + sh)                       with sc = rstd gamma, sh = beta - mean sc
// MODE 1: dx = sc dy' + x kb + kc                      (also writes dgamma / dbeta once)
//         = gamma rstd (dy' - [train] (mean(dy') + xhat mean(dy' xhat))),  dy' = dy [x sc + sh > 0]
// Four per-channel constants per thread keep the register count low enough for 2+ blocks per SM.
template <typename T, int MODE>
__global__ void __launch_bounds__(BN_THREADS, sizeof(T) == 2 ? 3 : 0) bn_rows_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t ldd,
                                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                                                              int training, const double* __restrict__ s_dy, const double* __restrict__ s_dyx,
                                                              T* __restrict__ out, int64_t ldo, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, int64_t rows, int C) {
  constexpr int V = Vec<T>::N;
  constexpr int U = MODE == 1 ? 2 : BN_UNROLL;
  const int tpr = C / V, rpb = BN_THREADS / tpr;
  const int ch = (threadIdx.x % tpr) * V, rg = threadIdx.x / tpr;
  float sc[V], sh[V], kb[V], kc[V];
  const float inv = 1.f / (float)rows;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const float m = mean[ch + j], rs = rstd[ch + j];
    sc[j] = rs * gamma[ch + j];
    sh[j] = (beta ? beta[ch + j] : 0.f) - m * sc[j];
    if (MODE == 1) {
      const float k0 = training ? (float)s_dy[ch + j] * inv : 0.f;
      const float k1 = training ? (float)s_dyx[ch + j] * inv : 0.f;
      kb[j] = -sc[j] * k1 * rs;
      kc[j] = -sc[j] * k0 - kb[j] * m;
    }
  }
  if (MODE == 1 && blockIdx.x == 0 && rg == 0) {
#pragma unroll
    for (int j = 0; j < V; ++j) { dgamma[ch + j] = (float)s_dyx[ch + j]; dbeta[ch + j] = (float)s_dy[ch + j]; }
  }
  const int64_t stride = (int64_t)gridDim.x * rpb;
  for (int64_t r = (int64_t)blockIdx.x * rpb + rg; r < rows; r += stride * U) {
    typename Vec<T>::Raw xr[U], dr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + u * stride;
      xr[u] = rr < rows ? Vec<T>::load_raw(x + rr * ldx + ch) : Vec<T>::zero_raw();
      if (MODE == 1) dr[u] = rr < rows ? Vec<T>::load_raw(dy + rr * ldd + ch) : Vec<T>::zero_raw();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + u * stride;
      if (rr >= rows) continue;
      float xv[V], dv[V], o[V];
      Vec<T>::unpack(xr[u], xv);
      if (MODE == 1) Vec<T>::unpack(dr[u], dv);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float v = fmaf(xv[j], sc[j], sh[j]);
        if (MODE == 0) {
          o[j] = relu ? fmaxf(v, 0.f) : v;
        } else {
          const float d = (relu && !(v > 0.f)) ? 0.f : dv[j];
          o[j] = fmaf(sc[j], d, fmaf(xv[j], kb[j], kc[j]));
        }
      }
      Vec<T>::store(out + rr * ldo + ch, o);
    }
  }
}

template <typename T>
static bool bn_shape_ok(int c) {
  constexpr int V = Vec<T>::N;
  return c % V == 0 && c / V <= BN_THREADS && BN_THREADS % (c / V) == 0;
}
// Blocks per SM of the column-sum kernels.  Every block ends with 2 C double atomics onto the same 2 C addresses, and with
// 16 blocks per SM that serialised tail cost more than the streaming part on the 36-61 MB sparse-conv tensors (57.7 us at
// 16, 43.4 at 8, 35.3 at 4, 33.0 at 2 for 70k x 128 fp32; tools/kernel_bench.py bn).  0 = automatic: 2, or 4 for maps of
// more than 64 Mi elements (the decoder's BEV maps); tmae_set_option("bn_colsum_cap", n) forces n.
int g_bn_colsum_cap = 0;
static int bn_grid(int64_t rows, int c, int vec, int per_sm = 16) {
  int rpb = BN_THREADS / (c / vec);
  int64_t blocks = (rows + (int64_t)rpb * BN_UNROLL - 1) / ((int64_t)rpb * BN_UNROLL);
  int64_t cap = (int64_t)kNumSMs * per_sm;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

template <typename T>
static int bn_fwd_impl(const T* x, int64_t ldx, const float* gamma, const float* beta, float* running_mean, float* running_var,
                       float momentum, float eps, T* y, int64_t ldy, float* mean, float* rstd, int64_t rows, int c, int relu, bool train,
                       double* ws, cudaStream_t s) {
  const int grid = bn_grid(rows, c, Vec<T>::N);
  if (train) {
    if (cudaMemsetAsync(ws, 0, 2 * c * sizeof(double), s) != cudaSuccess) return TMAE_ERR_CUDA;
    bn_colsum_kernel<T, 0><<<bn_grid(rows, c, Vec<T>::N, g_bn_colsum_cap > 0 ? g_bn_colsum_cap : (rows * c > ((int64_t)64 << 20) ? 4 : 2)), BN_THREADS, 0, s>>>(x, ldx, nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0, rows, c, ws, ws + c);
    bn_finalize_kernel<<<cdiv(c, 128), 128, 0, s>>>(ws, ws + c, rows, c, eps, momentum, mean, rstd, running_mean, running_var);
  }
  bn_rows_kernel<T, 0><<<grid, BN_THREADS, 0, s>>>(x, ldx, nullptr, 0, mean, rstd, gamma, beta, relu, 0, nullptr, nullptr, y, ldy, nullptr,
                                                   nullptr, rows, c);
  return cudaGetLastError() == cudaSuccess ? 0 : TMAE_ERR_CUDA;
}

template <typename T>
static int bn_bwd_impl(const T* dy, int64_t ldd, const T* x, int64_t ldx, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, T* dx, int64_t ldo, float* dgamma, float* dbeta, int64_t rows, int c, int relu, int training,
                       double* ws, cudaStream_t s) {
  const int grid = bn_grid(rows, c, Vec<T>::N);
  if (cudaMemsetAsync(ws, 0, 2 * c * sizeof(double), s) != cudaSuccess) return TMAE_ERR_CUDA;
  bn_colsum_kernel<T, 1><<<bn_grid(rows, c, Vec<T>::N, g_bn_colsum_cap > 0 ? g_bn_colsum_cap : (rows * c > ((int64_t)64 << 20) ? 4 : 2)), BN_THREADS, 0, s>>>(x, ldx, dy, ldd, mean, rstd, gamma, beta, relu, rows, c, ws, ws + c);
  bn_rows_kernel<T, 1><<<grid, BN_THREADS, 0, s>>>(x, ldx, dy, ldd, mean, rstd, gamma, beta, relu, training, ws, ws + c, dx, ldo, dgamma, dbeta,
                                                   rows, c);
  return cudaGetLastError() == cudaSuccess ? 0 : TMAE_ERR_CUDA;
}

}  // namespace tmae

using namespace tmae;

extern "C" {

size_t tmae_bn_workspace_bytes(int32_t c) { return (size_t)2 * c * sizeof(double) + 256; }

int tmae_bn_train_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var, float momentum,
                      float eps, float* y, float* save_mean, float* save_rstd, int64_t rows, int32_t c, int32_t relu,
                      void* workspace, size_t workspace_bytes, void* stream) {
  TMAE_CHECK_ARG(workspace_bytes >= tmae_bn_workspace_bytes(c), "workspace too small");
  TMAE_CHECK_ARG(rows > 0, "BatchNorm needs at least one row");
  TMAE_CHECK_ARG(bn_shape_ok<float>(c), "channels must be a multiple of 4 that divides 1024");
  ProfScope prof("bn_train_fwd", 0, 12.0 * rows * c, (cudaStream_t)stream);
  int r = bn_fwd_impl<float>(x, c, gamma, beta, running_mean, running_var, momentum, eps, y, c, save_mean, save_rstd, rows, c, relu, true,
                             (double*)workspace, (cudaStream_t)stream);
  if (r) set_error("tmae_bn_train_fwd: launch failed");
  return r;
}

int tmae_bn_apply(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, float* y, int64_t rows,
                  int32_t c, int32_t relu, void* stream) {
  if (rows <= 0) return 0;
  TMAE_CHECK_ARG(bn_shape_ok<float>(c), "channels must be a multiple of 4 that divides 1024");
  int r = bn_fwd_impl<float>(x, c, gamma, beta, nullptr, nullptr, 0.f, 0.f, y, c, (float*)mean, (float*)rstd, rows, c, relu, false, nullptr,
                             (cudaStream_t)stream);
  if (r) set_error("tmae_bn_apply: launch failed");
  return r;
}

int tmae_bn_bwd(const float* dy, const float* x, const float* y, const float* mean, const float* rstd, const float* gamma, const float* beta,
                float* dx, float* dgamma, float* dbeta, int64_t rows, int32_t c, int32_t relu, int32_t training, void* workspace,
                size_t workspace_bytes, void* stream) {
  (void)y;  // the ReLU mask is recomputed from x
  TMAE_CHECK_ARG(workspace_bytes >= tmae_bn_workspace_bytes(c), "workspace too small");
  TMAE_CHECK_ARG(rows > 0, "BatchNorm needs at least one row");
  TMAE_CHECK_ARG(bn_shape_ok<float>(c), "channels must be a multiple of 4 that divides 1024");
  TMAE_CHECK_ARG(!relu || beta, "beta is needed to recompute the ReLU mask");
  ProfScope prof("bn_bwd", 0, 20.0 * rows * c, (cudaStream_t)stream);
  int r = bn_bwd_impl<float>(dy, c, x, c, mean, rstd, gamma, beta, dx, c, dgamma, dbeta, rows, c, relu, training, (double*)workspace,
                             (cudaStream_t)stream);
  if (r) set_error("tmae_bn_bwd: launch failed");
  return r;
}

int tmae_bn_bf16_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, float* running_mean, float* running_var,
                     float momentum, float eps, void* y, int64_t ldy, float* mean, float* rstd, int64_t rows, int32_t c, int32_t relu,
                     int32_t training, void* workspace, size_t workspace_bytes, void* stream) {
  TMAE_CHECK_ARG(workspace_bytes >= tmae_bn_workspace_bytes(c), "workspace too small");
  TMAE_CHECK_ARG(rows > 0, "BatchNorm needs at least one row");
  TMAE_CHECK_ARG(bn_shape_ok<__nv_bfloat16>(c) && ldx % 8 == 0 && ldy % 8 == 0, "channels / pitches must be multiples of 8");
  TMAE_CHECK_ARG((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "bf16 maps must be 16-byte aligned");
  ProfScope prof("bn2d_fwd", 0, 6.0 * rows * c, (cudaStream_t)stream);
  int r = bn_fwd_impl<__nv_bfloat16>((const __nv_bfloat16*)x, ldx, gamma, beta, running_mean, running_var, momentum, eps, (__nv_bfloat16*)y, ldy,
                                     mean, rstd, rows, c, relu, training != 0, (double*)workspace, (cudaStream_t)stream);
  if (r) set_error("tmae_bn_bf16_fwd: launch failed");
  return r;
}

int tmae_bn_bf16_bwd(const void* dy, int64_t ldd, const void* x, int64_t ldx, const float* mean, const float* rstd, const float* gamma,
                     const float* beta, void* dx, int64_t ldo, float* dgamma, float* dbeta, int64_t rows, int32_t c, int32_t relu,
                     int32_t training, void* workspace, size_t workspace_bytes, void* stream) {
  TMAE_CHECK_ARG(workspace_bytes >= tmae_bn_workspace_bytes(c), "workspace too small");
  TMAE_CHECK_ARG(rows > 0, "BatchNorm needs at least one row");
  TMAE_CHECK_ARG(bn_shape_ok<__nv_bfloat16>(c) && ldx % 8 == 0 && ldd % 8 == 0 && ldo % 8 == 0, "channels / pitches must be multiples of 8");
  TMAE_CHECK_ARG((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx) & 15) == 0, "bf16 maps must be 16-byte aligned");
  ProfScope prof("bn2d_bwd", 0, 10.0 * rows * c, (cudaStream_t)stream);
  int r = bn_bwd_impl<__nv_bfloat16>((const __nv_bfloat16*)dy, ldd, (const __nv_bfloat16*)x, ldx, mean, rstd, gamma, beta, (__nv_bfloat16*)dx, ldo,
                                     dgamma, dbeta, rows, c, relu, training, (double*)workspace, (cudaStream_t)stream);
  if (r) set_error("tmae_bn_bf16_bwd: launch failed");
  return r;
}

}  // extern "C"
