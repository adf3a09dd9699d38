// fp32 SIMT GEMM family (TMAE_PREC_FP32, the parity mode) for every dense contraction on the path:
// linear layers (cosine_msa.py:57-62,431; sst_basic_block.py:81; network_utils.py:30;
// SiamWCA_MAE.py:117-119), their backward passes, and the gather-GEMM form of the 2-D sparse
// convolutions (utils/spconv_utils.py:37-56).  The bf16 tcgen05 path lives in gemm_tc.cu and is
// validated against this one.
//
// One kernel template computes C[m,n] = sum_k A(m,k) * B(k,n) over 64x64x16 tiles (256 threads,
// 4x4 register micro-tiles, operands staged k-major in shared memory); the operand accessors are
// template modes so forward (NT), backward-data (NN), backward-weight (TN, split over the long
// reduction with fp32 atomics) and the neighbour-table gathers share the inner loop.
#include <string.h>

#include "common.cuh"

namespace tmae {

constexpr int BM = 64, BN = 64, BK = 16, GT = 256;

enum AMode { A_KCONTIG = 0, A_MCONTIG = 1, A_GATHER = 2 };
enum BMode { B_KCONTIG = 0, B_NCONTIG = 1, B_GATHER = 2 };

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  int64_t M, N, K;
  int64_t lda, ldb, ldc;
  const float* bias;      // per n
  const float* lut; const uint8_t* rowidx; int64_t ldlut;  // optional per-row table bias: + lut[rowidx[m]][n]
  const float* residual;  // same layout as C
  float* preact;          // optional copy of the pre-activation (for GELU backward)
  const int* tab;         // neighbour table (rows, taps), -1 = absent
  int taps, cin;          // gather decomposition: k (or n) = tap * cin + c
  const int* row_count;   // optional device row count overriding M (rows beyond are skipped)
  int act;
  int accumulate;         // C += result
  int atomic;             // split-K: atomicAdd into C
  int64_t k_chunk;        // K range per blockIdx.z
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

template <int AM, int BMD>
__global__ void __launch_bounds__(GT) gemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  int64_t M = g.M;
  if (g.row_count) {
    int64_t rc = *g.row_count;
    if (AM != A_MCONTIG) { M = rc < M ? rc : M; }
  }
  int64_t kbeg = (int64_t)blockIdx.z * g.k_chunk;
  int64_t kend = kbeg + g.k_chunk < g.K ? kbeg + g.k_chunk : g.K;
  if (g.row_count && AM == A_MCONTIG) {  // TN: the reduction runs over rows
    int64_t rc = *g.row_count;
    if (kend > rc) kend = rc;
  }
  if (m0 >= M && AM != A_MCONTIG) return;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- stage A tile (BM x BK) into As[k][m]
    if (AM == A_MCONTIG) {
      int k = tid >> 4, m4 = (tid & 15) * 4;
      int64_t kk = k0 + k;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int64_t mm = m0 + m4 + j;
        As[k][m4 + j] = (kk < kend && mm < g.M) ? g.A[kk * g.lda + mm] : 0.f;
      }
    } else {
      int m = tid >> 2, k4 = (tid & 3) * 4;
      int64_t mm = m0 + m;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int64_t kk = k0 + k4 + j;
        float v = 0.f;
        if (mm < M && kk < kend) {
          if (AM == A_KCONTIG) {
            v = g.A[mm * g.lda + kk];
          } else {
            int tap = (int)(kk / g.cin);
            int c = (int)(kk - (int64_t)tap * g.cin);
            int row = g.tab[mm * g.taps + tap];
            v = row >= 0 ? g.A[(int64_t)row * g.lda + c] : 0.f;
          }
        }
        As[k4 + j][m] = v;
      }
    }
    // ---- stage B tile (BK x BN) into Bs[k][n]
    if (BMD == B_KCONTIG) {
      int n = tid >> 2, k4 = (tid & 3) * 4;
      int64_t nn = n0 + n;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int64_t kk = k0 + k4 + j;
        Bs[k4 + j][n] = (nn < g.N && kk < kend) ? g.B[nn * g.ldb + kk] : 0.f;
      }
    } else {
      int k = tid >> 4, n4 = (tid & 15) * 4;
      int64_t kk = k0 + k;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int64_t nn = n0 + n4 + j;
        float v = 0.f;
        if (kk < kend && nn < g.N) {
          if (BMD == B_NCONTIG) {
            v = g.B[kk * g.ldb + nn];
          } else {
            int tap = (int)(nn / g.cin);
            int c = (int)(nn - (int64_t)tap * g.cin);
            int row = g.tab[kk * g.taps + tap];
            v = row >= 0 ? g.B[(int64_t)row * g.ldb + c] : 0.f;
          }
        }
        Bs[k][n4 + j] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  int64_t Mout = (AM == A_MCONTIG) ? g.M : M;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t mm = m0 + ty * 4 + i;
    if (mm >= Mout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t nn = n0 + tx * 4 + j;
      if (nn >= g.N) continue;
      float v = acc[i][j];
      int64_t o = mm * g.ldc + nn;
      if (g.atomic) {
        atomicAdd(g.C + o, v);
        continue;
      }
      if (g.bias) v += g.bias[nn];
      if (g.lut) v += g.lut[(int64_t)g.rowidx[mm] * g.ldlut + nn];
      if (g.preact) g.preact[o] = v;
      if (g.act == TMAE_ACT_GELU) v = gelu_erf(v);
      else if (g.act == TMAE_ACT_RELU) v = fmaxf(v, 0.f);
      if (g.residual) v += g.residual[o];
      if (g.accumulate) v += g.C[o];
      g.C[o] = v;
    }
  }
}

template <int AM, int BMD>
static int launch(GemmArgs& g, int splits, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return 0;
  if (splits < 1) splits = 1;
  g.k_chunk = align_up((g.K + splits - 1) / splits, BK);
  int z = (int)((g.K + g.k_chunk - 1) / g.k_chunk);
  if (z < 1) z = 1;
  g.atomic = z > 1;
  dim3 grid((unsigned)cdiv(g.N, BN), (unsigned)cdiv(g.M, BM), (unsigned)z);
  ProfScope prof("gemm_fp32_simt", 2.0 * g.M * g.N * g.K, 4.0 * ((double)g.M * g.K + (double)g.N * g.K + (double)g.M * g.N), s);
  gemm_kernel<AM, BMD><<<grid, GT, 0, s>>>(g);
  return cudaGetLastError() == cudaSuccess ? 0 : TMAE_ERR_CUDA;
}

// y[m, n] = x[m, k] w[n, k]^T (+ bias) for k <= 16: thread = (row, 4 adjacent outputs), w transposed in shared memory
// (conflict-free float4 reads), the k inputs of a row are read once per thread (16 threads of a row share them through
// L1), 16-byte coalesced stores.  fp32 FMA chain over k in ascending order.
__global__ void __launch_bounds__(256) thin_linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ y, int64_t m, int n, int k) {
  extern __shared__ float wt[];   // [k][n]
  for (int i = threadIdx.x; i < n * k; i += blockDim.x) wt[(i % k) * n + i / k] = w[i];
  __syncthreads();
  const int n4 = n >> 2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * n4) return;
  const int64_t row = t / n4;
  const int c = (int)(t - row * n4) * 4;
  float4 acc = bias ? *reinterpret_cast<const float4*>(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float* xr = x + row * k;
  for (int kk = 0; kk < k; ++kk) {
    const float xv = __ldg(xr + kk);
    const float4 wv = *reinterpret_cast<const float4*>(wt + kk * n + c);
    acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y); acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
  }
  *reinterpret_cast<float4*>(y + row * n + c) = acc;
}

// dw[n][k] += sum_r dy[r][n] x[r][k] for k <= 16 and n in {32, 64, 128, 256} (the VFE's first layer: 240 k point rows, n = 64, k = 10;
// the 64 x 64 x 16 tile kernel ran this shape at 105 us, 0.7 TB/s).  Thread = (row lane, output column): dy is read coalesced, the rows'
// k inputs come from shared memory (one broadcast read per warp), k accumulators per thread; block reduction over the row lanes, one
// atomicAdd per (n, k) per block.
__global__ void __launch_bounds__(256) thin_linear_bwd_weight_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                                                                     int64_t m, int n, int k, int rpb) {
  constexpr int TILE = 64;
  __shared__ float xs[TILE * 16];
  __shared__ float red[256 * 16];
  const int tn = threadIdx.x % n, rl = threadIdx.x / n, RL = 256 / n;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * rpb, r1 = r0 + rpb < m ? r0 + rpb : m;
  for (int64_t base = r0; base < r1; base += TILE) {
    const int rows = (int)(r1 - base < TILE ? r1 - base : TILE);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * k; i += 256) xs[i] = __ldg(x + base * k + i);
    __syncthreads();
    int r = rl;
    for (; r + 3 * RL < rows; r += 4 * RL) {      // four independent dy loads in flight
      float d[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) d[u] = __ldg(dy + (base + r + u * RL) * n + tn);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* xr = xs + (r + u * RL) * k;
#pragma unroll
        for (int kk = 0; kk < 16; ++kk)
          if (kk < k) acc[kk] = fmaf(d[u], xr[kk], acc[kk]);
      }
    }
    for (; r < rows; r += RL) {
      const float d = __ldg(dy + (base + r) * n + tn);
      const float* xr = xs + r * k;
#pragma unroll
      for (int kk = 0; kk < 16; ++kk)
        if (kk < k) acc[kk] = fmaf(d, xr[kk], acc[kk]);
    }
  }
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) red[kk * 256 + threadIdx.x] = acc[kk];
  __syncthreads();
  for (int i = threadIdx.x; i < n * k; i += 256) {
    const int j = i / k, kk = i - j * k;
    float t = 0.f;
    for (int q = 0; q < RL; ++q) t += red[kk * 256 + q * n + j];
    atomicAdd(dw + i, t);
  }
}

__global__ void colsum_kernel(const float* __restrict__ x, int64_t rows, int cols, const int* __restrict__ row_count,
                              float* __restrict__ out, int rows_per_block) {
  // block: 256 threads = 8 row-lanes x 32 column-lanes; grid.x over column groups of 32, grid.y over row chunks
  __shared__ float sm[8][33];
  if (row_count && *row_count < rows) rows = *row_count;
  int c = blockIdx.x * 32 + (threadIdx.x & 31);
  int rl = threadIdx.x >> 5;
  int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s = 0.f;
  if (c < cols)
    for (int64_t r = r0 + rl; r < r1; r += 8) s += x[r * cols + c];
  sm[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x & 31];
    atomicAdd(out + c, t);
  }
}

// table[p][j] = (j < n_pos ? pos_lut[p] . w[j] : 0) + bias[j]: 64 x n dot products of length c (one warp each);
// table_t (optional) receives the transpose (n, 64), the K-major second weight operand of tmae_linear_fwd_dual
__global__ void pos_table_kernel(const float* __restrict__ lut, const float* __restrict__ w, const float* __restrict__ bias,
                                 float* __restrict__ table, float* __restrict__ table_t, int n, int n_pos, int c) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= 64 * n) return;
  const int p = warp / n, j = warp - p * n;
  float s = 0.f;
  if (j < n_pos)
    for (int k = lane; k < c; k += 32) s = fmaf(lut[p * c + k], w[(int64_t)j * c + k], s);
  s = warp_sum(s);
  if (lane == 0) {
    const float v = s + (bias ? bias[j] : 0.f);
    if (table) table[(int64_t)p * n + j] = v;
    if (table_t) table_t[(int64_t)j * 64 + p] = v;
  }
}

// dtable[p][:] = sum of dy rows whose rowidx == p.  Block = 256 threads = 8 row lanes x 32 column lanes x float4 (128
// columns per blockIdx.x), per-block partial bins in shared memory (64 x 128 floats), one atomicAdd per bin entry per block.
__global__ void __launch_bounds__(256) binned_colsum_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ rowidx,
                                                            float* __restrict__ dtable, int64_t rows, int n, int rows_per_block) {
  __shared__ float bins[64][128];
  for (int i = threadIdx.x; i < 64 * 128; i += 256) (&bins[0][0])[i] = 0.f;
  __syncthreads();
  const int c0 = blockIdx.x * 128 + (threadIdx.x & 31) * 4, rl = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  if (c0 < n) {
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(dy + r * n + c0));
      float* b = &bins[rowidx[r] & 63][(threadIdx.x & 31) * 4];
      atomicAdd(b, v.x); atomicAdd(b + 1, v.y); atomicAdd(b + 2, v.z); atomicAdd(b + 3, v.w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 128; i += 256) {
    const int p = i >> 7, c = blockIdx.x * 128 + (i & 127);
    const float v = bins[p][i & 127];
    if (c < n && v != 0.f) atomicAdd(dtable + (int64_t)p * n + c, v);
  }
}

// dbias[j] = sum_p dtable[p][j] ; dw[j][:] += sum_p dtable[p][j] * lut[p][:] for j < n_pos.  One block per j.
__global__ void pos_table_bwd_kernel(const float* __restrict__ dtable, const float* __restrict__ lut, float* __restrict__ dw,
                                     float* __restrict__ dbias, int n, int n_pos, int c, int transposed) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: see tc_common.cuh pdl_trigger
  __shared__ float col[64];
  const int j = blockIdx.x;
  if (threadIdx.x < 64) col[threadIdx.x] = transposed ? dtable[(int64_t)j * 64 + threadIdx.x] : dtable[(int64_t)threadIdx.x * n + j];
  __syncthreads();
  if (threadIdx.x == 0 && dbias) {
    float s = 0.f;
    for (int p = 0; p < 64; ++p) s += col[p];
    dbias[j] = s;
  }
  if (j < n_pos)
    for (int k = threadIdx.x; k < c; k += blockDim.x) {
      float s = 0.f;
#pragma unroll 8
      for (int p = 0; p < 64; ++p) s = fmaf(col[p], lut[p * c + k], s);
      dw[(int64_t)j * c + k] += s;
    }
}

// out[m][p] = (p == idx[m]): the (rows, 64) one-hot form of the window-cell index, the second A operand of the packed projection
__global__ void onehot64_kernel(const uint8_t* __restrict__ idx, float4* __restrict__ out, int64_t m) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one float4 (4 bins) per thread
  if (i >= m * 16) return;
  const int p = idx[i >> 4] & 63, b = (int)(i & 15) * 4;
  out[i] = make_float4(p == b ? 1.f : 0.f, p == b + 1 ? 1.f : 0.f, p == b + 2 ? 1.f : 0.f, p == b + 3 ? 1.f : 0.f);
}

__global__ void gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ pre, float* __restrict__ dx, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = pre[i];
  float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  dx[i] = dy[i] * (cdf + x * pdf);
}

__global__ void transpose_taps_kernel(const float* __restrict__ w, float* __restrict__ wt, int cout, int taps, int cin, int flip) {
  // w (cout, taps, cin) -> wt (cin, taps, cout); flip reverses the tap order
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t n = (int64_t)cout * taps * cin;
  if (i >= n) return;
  int c = (int)(i % cin);
  int t = (int)((i / cin) % taps);
  int o = (int)(i / ((int64_t)cin * taps));
  int tt = flip ? taps - 1 - t : t;
  wt[((int64_t)c * taps + tt) * cout + o] = w[i];
}

// TMA-fed tf32 tensor-core path (gemm_tma.cu)
bool tma_linear_fwd_ok(const float* x, const float* w, const float* y, const float* residual, int64_t m, int64_t n, int64_t k);
bool tma_linear_fwd_dual_ok(const float* x, const float* w, const float* x2, const float* w2, const float* y, int64_t m, int64_t n, int64_t k,
                            int64_t k2);
int tma_linear_fwd_dual(const float* x, const float* w, const float* x2, const float* w2, float* y, int64_t m, int64_t n, int64_t k, int64_t k2,
                        cudaStream_t s);
int tma_linear_fwd(const float* x, const float* w, const float* bias, float* y, float* preact, int64_t m, int64_t n, int64_t k, int act,
                   cudaStream_t s);
bool tma_linear_bwd_data_ok(const float* dy, const float* w, const float* dx, int64_t m, int64_t n, int64_t k);
int tma_linear_bwd_data(const float* dy, const float* w, float* dx, int64_t m, int64_t n, int64_t k, int accumulate, const float* gelu_pre,
                        cudaStream_t s);
bool tma_linear_bwd_weight_ok(const float* dy, const float* x, const float* dw, int64_t m, int64_t n, int64_t k);
int tma_linear_bwd_weight(const float* dy, const float* x, float* dw, int64_t m, int64_t n, int64_t k, cudaStream_t s);
bool tma_sparse_conv_ok(const float* x, const float* w, const float* y, int cin, int cout);
int tma_sparse_conv_fwd(const float* x, const int* table, const float* w, float* y, int64_t rows_out, int taps, int cin, int cout, int accumulate,
                        cudaStream_t s);
bool tma_sparse_conv_bwd_weight_ok(const float* dy, const float* x, const float* dw, int cin, int cout);
int tma_sparse_conv_bwd_weight(const float* dy, const float* x, const int* table, float* dw, int64_t rows_out, int taps, int cin, int cout,
                               cudaStream_t s);

}  // namespace tmae

using namespace tmae;

namespace tmae { extern int g_bf16_wgrad_stream; extern int g_bf16_tn_plain; extern int g_bf16_gemm_pdl; extern int g_bf16_gemm_occ2; extern int g_small_on_warps; extern bool g_wide_st; extern int g_attn_occ; extern int g_attn_occ_fwd; extern int g_bn_colsum_cap; extern int g_ln_bwd_cap; }  // gemm_tma.cu, attention_mma.cu, bn.cu, rowops.cu

#define TMAE_CHECK_PREC(p) TMAE_CHECK_ARG((p) == TMAE_PREC_FP32 || (p) == TMAE_PREC_TF32, "precision must be TMAE_PREC_FP32 or TMAE_PREC_TF32 (the bf16-storage mode has its own entry points: tmae_bf16_*)")
// Two kernels per GEMM shape and no third: the TMA-fed tcgen05 kernel in a tensor-core mode, the fp32 FFMA kernel in
// parity mode.  A tensor-core-mode call whose shape TMA cannot express still runs (fp32 FFMA) but is COUNTED, so tests
// and the bench can assert that the hot path never takes it (tmae_dispatch_counts).
static inline void count_simt(int precision) { count_dispatch(precision == TMAE_PREC_FP32 ? DISP_SIMT_FP32 : DISP_SIMT_IN_TC_MODE); }

extern "C" {

int tmae_set_option(const char* name, int32_t value) {
  if (name && !strcmp(name, "bn_colsum_cap")) { g_bn_colsum_cap = value < 0 ? 0 : value; return 0; }
  if (name && !strcmp(name, "ln_bwd_cap")) { g_ln_bwd_cap = value < 1 ? 1 : value; return 0; }
  if (name && !strcmp(name, "attn_occ_fwd")) { g_attn_occ_fwd = value; return 0; }
  if (name && !strcmp(name, "attn_occ")) { g_attn_occ = value; return 0; }   // 1: mma attention backward at one more CTA per SM
  if (name && !strcmp(name, "wgrad_stream")) { g_bf16_wgrad_stream = value != 0; return 0; }   // layer backward: weight-gradient GEMMs on the auxiliary stream
  if (name && !strcmp(name, "tn_plain_store")) { g_bf16_tn_plain = value != 0; return 0; }  // measurement only: wrong results
  if (name && !strcmp(name, "gemm_pdl")) { g_bf16_gemm_pdl = value != 0; return 0; }     // bf16 GEMMs launched with programmatic stream serialization
  if (name && !strcmp(name, "gemm_occ2")) { g_bf16_gemm_occ2 = value != 0; return 0; }   // bf16 GEMMs with N <= 128: two 3-stage CTAs per SM
  if (name && !strcmp(name, "attn_small_warps")) { g_small_on_warps = value != 0; return 0; }   // bf16 mode: <= 16-token windows on the warp kernels (1) or on tcgen05 tiles (0)
  if (name && !strcmp(name, "wide_st")) { g_wide_st = value != 0; return 0; }   // 0: 128-bit epilogue stores in the TMA GEMM (A/B measurement)
  set_error("tmae_set_option: unknown option");
  return TMAE_ERR_INVALID_ARG;
}

int tmae_linear_fwd(const float* x, const float* w, const float* bias, const float* residual, float* y, float* preact,
                    int64_t m, int64_t n, int64_t k, int32_t act, int32_t precision, void* stream) {
  TMAE_CHECK_PREC(precision);
  if (k <= 16 && n <= 256 && n % 4 == 0 && !residual && !preact && act == TMAE_ACT_NONE && m > 0 && (((uintptr_t)y) & 15) == 0 &&
      (((uintptr_t)bias) & 15) == 0) {
    // thin reduction (the VFE's first layer, k = 10): a tile kernel would run 16-wide k-blocks that are mostly padding
    count_dispatch(DISP_THIN_K);
    ProfScope prof("linear_thin_k", 2.0 * m * n * k, 4.0 * ((double)m * k + (double)n * k + (double)m * n), (cudaStream_t)stream);
    thin_linear_fwd_kernel<<<cdiv(m * (n / 4), 256), 256, (size_t)n * k * sizeof(float), (cudaStream_t)stream>>>(x, w, bias, y, m, (int)n, (int)k);
    TMAE_CHECK_LAUNCH();
    return 0;
  }
  if (precision == TMAE_PREC_TF32 && tma_linear_fwd_ok(x, w, y, residual, m, n, k)) {
    if (tma_linear_fwd(x, w, bias, y, preact, m, n, k, act, (cudaStream_t)stream)) { set_error("tmae_linear_fwd: TMA launch failed"); return TMAE_ERR_CUDA; }
    count_dispatch(DISP_TMA);
    return 0;
  }
  count_simt(precision);
  GemmArgs g{};
  g.A = x; g.B = w; g.C = y; g.M = m; g.N = n; g.K = k; g.lda = k; g.ldb = k; g.ldc = n;
  g.bias = bias; g.residual = residual; g.preact = preact; g.act = act;
  if (launch<A_KCONTIG, B_KCONTIG>(g, 1, (cudaStream_t)stream)) { set_error("tmae_linear_fwd: launch failed"); return TMAE_ERR_CUDA; }
  return 0;
}

int tmae_linear_fwd_lut(const float* x, const float* w, const float* lut, const uint8_t* rowidx, float* y, int64_t m, int64_t n,
                        int64_t k, int32_t precision, void* stream) {
  TMAE_CHECK_PREC(precision);
  TMAE_CHECK_ARG(lut && rowidx, "lut and rowidx are required");
  count_simt(precision);
  GemmArgs g{};  // always the fp32 SIMT kernel: the tensor-core form of the same product is tmae_linear_fwd_dual
  g.A = x; g.B = w; g.C = y; g.M = m; g.N = n; g.K = k; g.lda = k; g.ldb = k; g.ldc = n;
  g.lut = lut; g.rowidx = rowidx; g.ldlut = n;
  if (launch<A_KCONTIG, B_KCONTIG>(g, 1, (cudaStream_t)stream)) { set_error("tmae_linear_fwd_lut: launch failed"); return TMAE_ERR_CUDA; }
  return 0;
}

/* y = x w^T + x2 w2^T: the packed projection with the position term as a second (one-hot, table) source pair */
int tmae_linear_fwd_dual(const float* x, const float* w, const float* x2, const float* w2, float* y, int64_t m, int64_t n, int64_t k, int64_t k2,
                         int32_t precision, void* stream) {
  TMAE_CHECK_PREC(precision);
  if (precision == TMAE_PREC_TF32 && tma_linear_fwd_dual_ok(x, w, x2, w2, y, m, n, k, k2)) {
    if (tma_linear_fwd_dual(x, w, x2, w2, y, m, n, k, k2, (cudaStream_t)stream)) { set_error("tmae_linear_fwd_dual: TMA launch failed"); return TMAE_ERR_CUDA; }
    count_dispatch(DISP_TMA);
    return 0;
  }
  count_simt(precision);
  GemmArgs g{};
  g.A = x; g.B = w; g.C = y; g.M = m; g.N = n; g.K = k; g.lda = k; g.ldb = k; g.ldc = n;
  if (launch<A_KCONTIG, B_KCONTIG>(g, 1, (cudaStream_t)stream)) { set_error("tmae_linear_fwd_dual: launch failed"); return TMAE_ERR_CUDA; }
  GemmArgs h{};
  h.A = x2; h.B = w2; h.C = y; h.M = m; h.N = n; h.K = k2; h.lda = k2; h.ldb = k2; h.ldc = n; h.accumulate = 1;
  if (launch<A_KCONTIG, B_KCONTIG>(h, 1, (cudaStream_t)stream)) { set_error("tmae_linear_fwd_dual: launch failed"); return TMAE_ERR_CUDA; }
  return 0;
}

int tmae_linear_bwd_data(const float* dy, const float* w, float* dx, int64_t m, int64_t n, int64_t k, int32_t accumulate,
                         int32_t precision, void* stream) {
  TMAE_CHECK_PREC(precision);
  if (precision == TMAE_PREC_TF32 && tma_linear_bwd_data_ok(dy, w, dx, m, n, k)) {
    if (tma_linear_bwd_data(dy, w, dx, m, n, k, accumulate, nullptr, (cudaStream_t)stream)) { set_error("tmae_linear_bwd_data: TMA launch failed"); return TMAE_ERR_CUDA; }
    count_dispatch(DISP_TMA);
    return 0;
  }
  // dx[m,k] = sum_n dy[m,n] * w[n,k]
  count_simt(precision);
  GemmArgs g{};
  g.A = dy; g.B = w; g.C = dx; g.M = m; g.N = k; g.K = n; g.lda = n; g.ldb = k; g.ldc = k; g.accumulate = accumulate;
  if (launch<A_KCONTIG, B_NCONTIG>(g, 1, (cudaStream_t)stream)) { set_error("tmae_linear_bwd_data: launch failed"); return TMAE_ERR_CUDA; }
  return 0;
}

}  // extern "C"

extern "C" int tmae_gelu_bwd(const float* dy, const float* preact, float* dx, int64_t n, void* stream);

extern "C" {

/* dx = (dy w) * gelu'(preact): backward through  h = gelu(preact), f = h w2^T  in one pass (sst_basic_block.py:81) */
int tmae_linear_bwd_data_gelu(const float* dy, const float* w, const float* preact, float* dx, int64_t m, int64_t n, int64_t k,
                              int32_t precision, void* stream) {
  TMAE_CHECK_PREC(precision);
  if (precision == TMAE_PREC_TF32 && tma_linear_bwd_data_ok(dy, w, dx, m, n, k) && ((uintptr_t)preact & 15) == 0) {
    if (tma_linear_bwd_data(dy, w, dx, m, n, k, 0, preact, (cudaStream_t)stream)) { set_error("tmae_linear_bwd_data_gelu: TMA launch failed"); return TMAE_ERR_CUDA; }
    count_dispatch(DISP_TMA);
    return 0;
  }
  int r = tmae_linear_bwd_data(dy, w, dx, m, n, k, 0, precision, stream);
  if (r) return r;
  return tmae_gelu_bwd(dx, preact, dx, m * k, stream);
}

static int pick_splits(int64_t out_tiles, int64_t k) {
  int64_t want = (2 * kNumSMs + out_tiles - 1) / out_tiles;
  int64_t maxs = (k + 4 * BK - 1) / (4 * BK);
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  return (int)want;
}

int tmae_linear_bwd_weight(const float* dy, const float* x, float* dw, float* dbias, int64_t m, int64_t n, int64_t k,
                           int32_t precision, void* stream) {
  TMAE_CHECK_PREC(precision);
  cudaStream_t s = (cudaStream_t)stream;
  // dw[n,k] = sum_m dy[m,n] * x[m,k]   (overwrites dw; reduction over m split across CTAs)
  TMAE_CUDA(cudaMemsetAsync(dw, 0, (size_t)n * k * sizeof(float), s));
  if (precision == TMAE_PREC_TF32 && m > 0 && tma_linear_bwd_weight_ok(dy, x, dw, m, n, k)) {
    if (tma_linear_bwd_weight(dy, x, dw, m, n, k, s)) { set_error("tmae_linear_bwd_weight: TMA launch failed"); return TMAE_ERR_CUDA; }
    count_dispatch(DISP_TMA);
  } else if (k <= 16 && (n == 32 || n == 64 || n == 128 || n == 256)) {
    // the first VFE layer (k = 10: a 40-byte row pitch no tensor map can describe): the thin-k family in every mode
    count_dispatch(DISP_THIN_K);
    if (m > 0) {
      int64_t blocks = cdiv(m, 256);
      if (blocks > (int64_t)kNumSMs * 4) blocks = (int64_t)kNumSMs * 4;
      const int rpb = (int)align_up(cdiv(m, blocks), 64);
      ProfScope prof("linear_thin_k_wgrad", 2.0 * m * n * k, 4.0 * ((double)m * k + (double)n * k + (double)m * n), s);
      thin_linear_bwd_weight_kernel<<<(unsigned)cdiv(m, rpb), 256, 0, s>>>(dy, x, dw, m, (int)n, (int)k, rpb);
      TMAE_CHECK_LAUNCH();
    }
  } else {
    if (k <= 16) count_dispatch(DISP_THIN_K); else count_simt(precision);
    GemmArgs g{};
    g.A = dy; g.B = x; g.C = dw; g.M = n; g.N = k; g.K = m; g.lda = n; g.ldb = k; g.ldc = k;
    int splits = pick_splits((int64_t)cdiv(n, BM) * cdiv(k, BN), m);
    if (m > 0 && launch<A_MCONTIG, B_NCONTIG>(g, splits, s)) { set_error("tmae_linear_bwd_weight: launch failed"); return TMAE_ERR_CUDA; }
  }
  if (dbias) {
    TMAE_CUDA(cudaMemsetAsync(dbias, 0, (size_t)n * sizeof(float), s));
    if (m > 0) {
      int rpb = 512;
      dim3 grid((unsigned)cdiv(n, 32), (unsigned)cdiv(m, rpb));
      ProfScope prof("colsum", 0, 4.0 * m * n, s);
      colsum_kernel<<<grid, 256, 0, s>>>(dy, m, (int)n, nullptr, dbias, rpb);
      TMAE_CHECK_LAUNCH();
    }
  }
  return 0;
}

int tmae_colsum(const float* x, float* out, int64_t rows, int32_t cols, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CUDA(cudaMemsetAsync(out, 0, (size_t)cols * sizeof(float), s));
  if (rows > 0) {
    int rpb = 512;
    dim3 grid((unsigned)cdiv(cols, 32), (unsigned)cdiv(rows, rpb));
    colsum_kernel<<<grid, 256, 0, s>>>(x, rows, cols, nullptr, out, rpb);
    TMAE_CHECK_LAUNCH();
  }
  return 0;
}

int tmae_pos_table(const float* pos_lut, const float* w, const float* bias, float* table, float* table_t, int32_t n, int32_t n_pos, int32_t c,
                   void* stream) {
  if (n <= 0) return 0;
  pos_table_kernel<<<cdiv((int64_t)64 * n * 32, 256), 256, 0, (cudaStream_t)stream>>>(pos_lut, w, bias, table, table_t, n, n_pos, c);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_binned_colsum(const float* dy, const uint8_t* rowidx, float* dtable, int64_t rows, int32_t n, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TMAE_CHECK_ARG(n % 4 == 0, "n must be a multiple of 4");
  TMAE_CUDA(cudaMemsetAsync(dtable, 0, (size_t)64 * n * sizeof(float), s));
  if (rows <= 0) return 0;
  int rpb = 1024;
  dim3 grid((unsigned)cdiv(n, 128), (unsigned)cdiv(rows, rpb));
  ProfScope prof("binned_colsum", 0, 4.0 * rows * n, s);
  binned_colsum_kernel<<<grid, 256, 0, s>>>(dy, rowidx, dtable, rows, n, rpb);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_pos_table_bwd(const float* dtable, int32_t transposed, const float* pos_lut, float* dw, float* dbias, int32_t n, int32_t n_pos, int32_t c,
                       void* stream) {
  if (n <= 0) return 0;
  pos_table_bwd_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(dtable, pos_lut, dw, dbias, n, n_pos, c, transposed);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_onehot64(const uint8_t* idx, float* out, int64_t m, void* stream) {
  if (m <= 0) return 0;
  onehot64_kernel<<<cdiv(m * 16, 256), 256, 0, (cudaStream_t)stream>>>(idx, (float4*)out, m);
  TMAE_CHECK_LAUNCH();
  return 0;
}

int tmae_gelu_bwd(const float* dy, const float* preact, float* dx, int64_t n, void* stream) {
  if (n <= 0) return 0;
  ProfScope prof("gelu_bwd", 0, 12.0 * n, (cudaStream_t)stream);
  gelu_bwd_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, preact, dx, n);
  TMAE_CHECK_LAUNCH();
  return 0;
}

/* y[o,:] = sum_tap x[table[o,tap], :] . w[:, tap, :]^T      w is (cout, taps, cin) */
int tmae_sparse_conv_fwd(const float* x, const int32_t* table, const float* w, float* y, int64_t rows_out, int32_t taps,
                         int32_t cin, int32_t cout, int32_t accumulate, int32_t precision, void* stream) {
  TMAE_CHECK_PREC(precision);
  if (precision == TMAE_PREC_TF32 && rows_out > 0 && tma_sparse_conv_ok(x, w, y, cin, cout)) {
    if (tma_sparse_conv_fwd(x, table, w, y, rows_out, taps, cin, cout, accumulate, (cudaStream_t)stream)) { set_error("tmae_sparse_conv_fwd: TMA/cp.async launch failed"); return TMAE_ERR_CUDA; }
    count_dispatch(DISP_TMA);
    return 0;
  }
  count_simt(precision);
  GemmArgs g{};
  g.A = x; g.B = w; g.C = y; g.M = rows_out; g.N = cout; g.K = (int64_t)taps * cin; g.lda = cin; g.ldb = g.K; g.ldc = cout;
  g.tab = table; g.taps = taps; g.cin = cin; g.accumulate = accumulate;
  if (launch<A_GATHER, B_KCONTIG>(g, 1, (cudaStream_t)stream)) { set_error("tmae_sparse_conv_fwd: launch failed"); return TMAE_ERR_CUDA; }
  return 0;
}

/* dw[n, tap, c] = sum_o dy[o, n] * x[table[o, tap], c]   (overwrites dw) */
int tmae_sparse_conv_bwd_weight(const float* dy, const float* x, const int32_t* table, float* dw, int64_t rows_out, int32_t taps,
                                int32_t cin, int32_t cout, int32_t precision, void* stream) {
  TMAE_CHECK_PREC(precision);
  cudaStream_t s = (cudaStream_t)stream;
  int64_t kk = (int64_t)taps * cin;
  TMAE_CUDA(cudaMemsetAsync(dw, 0, (size_t)cout * kk * sizeof(float), s));
  if (rows_out <= 0) return 0;
  if (precision == TMAE_PREC_TF32 && rows_out > 0 && tma_sparse_conv_bwd_weight_ok(dy, x, dw, cin, cout)) {
    if (tma_sparse_conv_bwd_weight(dy, x, table, dw, rows_out, taps, cin, cout, s)) { set_error("tmae_sparse_conv_bwd_weight: TMA/cp.async launch failed"); return TMAE_ERR_CUDA; }
    count_dispatch(DISP_TMA);
    return 0;
  }
  count_simt(precision);
  GemmArgs g{};
  g.A = dy; g.B = x; g.C = dw; g.M = cout; g.N = kk; g.K = rows_out; g.lda = cout; g.ldb = cin; g.ldc = kk;
  g.tab = table; g.taps = taps; g.cin = cin;
  int splits = pick_splits((int64_t)cdiv(cout, BM) * cdiv(kk, BN), rows_out);
  if (launch<A_MCONTIG, B_GATHER>(g, splits, s)) { set_error("tmae_sparse_conv_bwd_weight: launch failed"); return TMAE_ERR_CUDA; }
  return 0;
}

/* w (cout, taps, cin) -> wt (cin, taps, cout): the weight of the backward-data gather GEMM */
int tmae_transpose_taps(const float* w, float* wt, int32_t cout, int32_t taps, int32_t cin, int32_t flip, void* stream) {
  int64_t n = (int64_t)cout * taps * cin;
  if (n <= 0) return 0;
  transpose_taps_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(w, wt, cout, taps, cin, flip);
  TMAE_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
