// Windowed cosine multi-head attention core, self (row A7) and temporal cross (row A8) forms.
//
// Replaces, per window: flat2window gathers, F.normalize, bmm(q,k^T)/clamp(tau), key-padding
// mask, softmax, bmm(p,v) and the window2flat scatter
// (pcdet/models/model_utils/cosine_msa.py:114-176,370-429; sst_basic_block.py:22-54;
//  wca_block.py:26-67; sst_utils.py:118-192).
//
// The reference pads every window to its level's token count and masks the padding; mathematically
// the padded keys contribute exp(-inf) = 0 and padded queries are discarded, so the kernel works
// on the ragged windows directly: a work item is (compact window, head group of 128 channels);
// the window's <= 64 key rows are gathered through the partition's token table into shared
// memory, normalised per head, and each lane owns one query row with an online softmax.  Results
// are written straight to the flat (voxel-major) layout.  The averaged attention map that the
// reference also computes (cosine_msa.py:433-436) is never consumed and is not produced.
#include "common.cuh"

namespace tmae {

constexpr int TW = 128;            // channels per work item
constexpr int TWP = TW + 1;        // padded row stride (conflict-free column access)
constexpr int MAXT = TMAE_WIN_TOKENS;
constexpr int ATT_THREADS = 256;

struct AttnArgs {
  const float* q; const float* k; const float* v;   // (rows, C)
  float* o;                                          // (q rows, C)
  float* lse;                                        // (q rows, H)
  const int* qtok; const int* qcnt;                  // [n_win*64], [n_win]
  const int* ktok; const int* kcnt;
  const int* n_win;                                  // device
  const float* tau; float tau_min;
  int C, H;
  // backward
  const float* dout; float* dq; float* dk; float* dv; float* dtau;
};

__device__ __forceinline__ void load_tile(float* dst, const float* __restrict__ src, const int* __restrict__ tok, int cnt, int C,
                                          int col0) {
  // dst[j][c] for j < cnt, c < TW ; coalesced along c
  for (int e = threadIdx.x; e < cnt * TW; e += ATT_THREADS) {
    int j = e / TW, c = e - j * TW;
    dst[j * TWP + c] = src[(int64_t)tok[j] * C + col0 + c];
  }
}

// scale every (row, head) slice to unit L2 norm (F.normalize, eps 1e-12); optionally keep 1/norm
__device__ __forceinline__ void normalize_tile(float* t, int cnt, int hd, int heads, float* inv_out) {
  for (int e = threadIdx.x; e < cnt * heads; e += ATT_THREADS) {
    int j = e / heads, h = e - j * heads;
    float* p = t + j * TWP + h * hd;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s += p[d] * p[d];
    float inv = 1.f / fmaxf(sqrtf(s), 1e-12f);
    for (int d = 0; d < hd; ++d) p[d] *= inv;
    if (inv_out) inv_out[j * 8 + h] = inv;
  }
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(AttnArgs a) {
  extern __shared__ float sm[];
  float* Ks = sm;                    // [64][TWP]
  float* Vs = sm + MAXT * TWP;       // [64][TWP]
  __shared__ int qt[MAXT], kt[MAXT];
  constexpr int HEADS = TW / HD;     // heads per work item: 8 (hd 16) or 4 (hd 32)
  constexpr int NSUB = 8 / HEADS;    // warps per head
  const int groups = a.C / TW;
  const int n_items = *a.n_win * groups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hl = warp / NSUB, sub = warp % NSUB;
  const float inv_tau = 1.f / fmaxf(*a.tau, a.tau_min);
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    int g = item / groups, col0 = (item - g * groups) * TW;
    int nq = a.qcnt[g], nk = a.kcnt[g];
    __syncthreads();
    if (threadIdx.x < MAXT) {
      qt[threadIdx.x] = threadIdx.x < nq ? a.qtok[g * MAXT + threadIdx.x] : 0;
      kt[threadIdx.x] = threadIdx.x < nk ? a.ktok[g * MAXT + threadIdx.x] : 0;
    }
    __syncthreads();
    load_tile(Ks, a.k, kt, nk, a.C, col0);
    load_tile(Vs, a.v, kt, nk, a.C, col0);
    __syncthreads();
    normalize_tile(Ks, nk, HD, HEADS, nullptr);
    __syncthreads();
    const int hcol = hl * HD;
    for (int i = lane + 32 * sub; i < nq; i += 32 * NSUB) {
      int64_t row = qt[i];
      const float* qp = a.q + row * a.C + col0 + hcol;
      float qv[HD];
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) { qv[d] = qp[d]; s += qv[d] * qv[d]; }
      float inv = inv_tau / fmaxf(sqrtf(s), 1e-12f);
#pragma unroll
      for (int d = 0; d < HD; ++d) qv[d] *= inv;       // q_hat / tau
      float m = -INFINITY, l = 0.f;
      float acc[HD];
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] = 0.f;
      for (int j = 0; j < nk; ++j) {
        const float* kp = Ks + j * TWP + hcol;
        float sc = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) sc = fmaf(qv[d], kp[d], sc);
        if (sc > m) {
          float r = __expf(m - sc);
          l *= r;
#pragma unroll
          for (int d = 0; d < HD; ++d) acc[d] *= r;
          m = sc;
        }
        float p = __expf(sc - m);
        l += p;
        const float* vp = Vs + j * TWP + hcol;
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] = fmaf(p, vp[d], acc[d]);
      }
      float il = 1.f / l;
      float* op = a.o + row * a.C + col0 + hcol;
#pragma unroll
      for (int d = 0; d < HD; ++d) op[d] = acc[d] * il;
      if (a.lse) a.lse[row * a.H + (col0 / HD) + hl] = m + __logf(l);
    }
  }
}

// backward: smem holds Q_hat, K_hat, V, dO tiles (4 x 64 x 129 floats) + per-row stats
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_kernel(AttnArgs a) {
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + MAXT * TWP;
  float* Vs = Ks + MAXT * TWP;
  float* Ds = Vs + MAXT * TWP;          // dO
  float* qinv = Ds + MAXT * TWP;        // [64][8] 1/|q|
  float* kinv = qinv + MAXT * 8;        // [64][8]
  float* lse_s = kinv + MAXT * 8;       // [64][8]
  float* dsum = lse_s + MAXT * 8;       // [64][8]  D_i = dO_i . O_i
  __shared__ int qt[MAXT], kt[MAXT];
  constexpr int HEADS = TW / HD;
  constexpr int NSUB = 8 / HEADS;
  const int groups = a.C / TW;
  const int n_items = *a.n_win * groups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hl = warp / NSUB, sub = warp % NSUB;
  const float tau_raw = *a.tau;
  const float inv_tau = 1.f / fmaxf(tau_raw, a.tau_min);
  float dtau_acc = 0.f;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    int g = item / groups, col0 = (item - g * groups) * TW;
    int nq = a.qcnt[g], nk = a.kcnt[g];
    __syncthreads();
    if (threadIdx.x < MAXT) {
      qt[threadIdx.x] = threadIdx.x < nq ? a.qtok[g * MAXT + threadIdx.x] : 0;
      kt[threadIdx.x] = threadIdx.x < nk ? a.ktok[g * MAXT + threadIdx.x] : 0;
    }
    __syncthreads();
    load_tile(Qs, a.q, qt, nq, a.C, col0);
    load_tile(Ds, a.dout, qt, nq, a.C, col0);
    load_tile(Ks, a.k, kt, nk, a.C, col0);
    load_tile(Vs, a.v, kt, nk, a.C, col0);
    __syncthreads();
    // D_i and lse per (query, head): O is read from global
    for (int e = threadIdx.x; e < nq * HEADS; e += ATT_THREADS) {
      int i = e / HEADS, h = e - i * HEADS;
      const float* op = a.o + (int64_t)qt[i] * a.C + col0 + h * HD;
      const float* dp = Ds + i * TWP + h * HD;
      float s = 0.f;
      for (int d = 0; d < HD; ++d) s += op[d] * dp[d];
      dsum[i * 8 + h] = s;
      lse_s[i * 8 + h] = a.lse[(int64_t)qt[i] * a.H + (col0 / HD) + h];
    }
    normalize_tile(Qs, nq, HD, HEADS, qinv);
    normalize_tile(Ks, nk, HD, HEADS, kinv);
    __syncthreads();
    const int hcol = hl * HD;
    // ---- phase A: lane = query  -> dQ
    for (int i = lane + 32 * sub; i < nq; i += 32 * NSUB) {
      float qv[HD], dov[HD], dqh[HD];
#pragma unroll
      for (int d = 0; d < HD; ++d) { qv[d] = Qs[i * TWP + hcol + d]; dov[d] = Ds[i * TWP + hcol + d]; dqh[d] = 0.f; }
      float Di = dsum[i * 8 + hl], Li = lse_s[i * 8 + hl];
      for (int j = 0; j < nk; ++j) {
        const float* kp = Ks + j * TWP + hcol;
        const float* vp = Vs + j * TWP + hcol;
        float sc = 0.f, dp = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) { sc = fmaf(qv[d], kp[d], sc); dp = fmaf(dov[d], vp[d], dp); }
        sc *= inv_tau;
        float p = __expf(sc - Li);
        float ds = p * (dp - Di);
        dtau_acc -= ds * sc;               // d/dtau of (c / tau) = -(c / tau) / tau ; the 1/tau is applied at the end
        float dsl = ds * inv_tau;
#pragma unroll
        for (int d = 0; d < HD; ++d) dqh[d] = fmaf(dsl, kp[d], dqh[d]);
      }
      // through q_hat = q / max(|q|, eps)
      float dot = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) dot = fmaf(dqh[d], qv[d], dot);
      float inv = qinv[i * 8 + hl];
      float* out = a.dq + (int64_t)qt[i] * a.C + col0 + hcol;
#pragma unroll
      for (int d = 0; d < HD; ++d) out[d] = (dqh[d] - qv[d] * dot) * inv;
    }
    // ---- phase B: lane = key -> dK, dV
    for (int j = lane + 32 * sub; j < nk; j += 32 * NSUB) {
      float kv[HD], vv[HD], dkh[HD], dvv[HD];
#pragma unroll
      for (int d = 0; d < HD; ++d) { kv[d] = Ks[j * TWP + hcol + d]; vv[d] = Vs[j * TWP + hcol + d]; dkh[d] = 0.f; dvv[d] = 0.f; }
      for (int i = 0; i < nq; ++i) {
        const float* qp = Qs + i * TWP + hcol;
        const float* dop = Ds + i * TWP + hcol;
        float sc = 0.f, dp = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) { sc = fmaf(qp[d], kv[d], sc); dp = fmaf(dop[d], vv[d], dp); }
        float p = __expf(sc * inv_tau - lse_s[i * 8 + hl]);
        float dsl = p * (dp - dsum[i * 8 + hl]) * inv_tau;
#pragma unroll
        for (int d = 0; d < HD; ++d) { dkh[d] = fmaf(dsl, qp[d], dkh[d]); dvv[d] = fmaf(p, dop[d], dvv[d]); }
      }
      float dot = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) dot = fmaf(dkh[d], kv[d], dot);
      float inv = kinv[j * 8 + hl];
      float* dko = a.dk + (int64_t)kt[j] * a.C + col0 + hcol;
      float* dvo = a.dv + (int64_t)kt[j] * a.C + col0 + hcol;
#pragma unroll
      for (int d = 0; d < HD; ++d) { dko[d] = (dkh[d] - kv[d] * dot) * inv; dvo[d] = dvv[d]; }
    }
  }
  // tau gradient (clamp passes gradient only where tau > tau_min)
  dtau_acc = warp_sum(dtau_acc);
  if (lane == 0 && a.dtau && tau_raw > a.tau_min && dtau_acc != 0.f) atomicAdd(a.dtau, dtau_acc * inv_tau);
}

static int check(const AttnArgs& a, int hd) {
  if (a.C % TW != 0 || a.H != a.C / hd || (hd != 16 && hd != 32)) return -1;
  return 0;
}

}  // namespace tmae

using namespace tmae;

extern "C" {

/* o[qrow] = softmax(q_hat k_hat^T / max(tau, tau_min)) v per window and head; lse (rows, H) is saved for backward (nullable).
 * Rows of `o` that belong to no window are left untouched (the caller zero-fills for the cross form). */
int tmae_window_attention_fwd(const float* q, const float* k, const float* v, float* o, float* lse, const int32_t* qtok,
                              const int32_t* qcnt, const int32_t* ktok, const int32_t* kcnt, const int32_t* n_win,
                              int64_t max_windows, const float* tau, float tau_min, int32_t channels, int32_t heads, void* stream) {
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.o = o; a.lse = lse; a.qtok = qtok; a.qcnt = qcnt; a.ktok = ktok; a.kcnt = kcnt; a.n_win = n_win;
  a.tau = tau; a.tau_min = tau_min; a.C = channels; a.H = heads;
  int hd = channels / heads;
  TMAE_CHECK_ARG(check(a, hd) == 0, "channels must be a multiple of 128 and head_dim 16 or 32");
  if (max_windows <= 0) return 0;
  size_t smem = (size_t)2 * MAXT * TWP * sizeof(float);
  int64_t items = max_windows * (channels / TW);
  int grid = (int)(items < 4 * kNumSMs ? items : 4 * kNumSMs);
  cudaStream_t s = (cudaStream_t)stream;
  if (hd == 16) {
    TMAE_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_kernel<16><<<grid, ATT_THREADS, smem, s>>>(a);
  } else {
    TMAE_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_kernel<32><<<grid, ATT_THREADS, smem, s>>>(a);
  }
  TMAE_CHECK_LAUNCH();
  return 0;
}

/* dq/dk/dv rows that belong to no window are left untouched (caller zero-fills); dtau (1 float) is accumulated. */
int tmae_window_attention_bwd(const float* dout, const float* q, const float* k, const float* v, const float* o, const float* lse,
                              float* dq, float* dk, float* dv, float* dtau, const int32_t* qtok, const int32_t* qcnt,
                              const int32_t* ktok, const int32_t* kcnt, const int32_t* n_win, int64_t max_windows, const float* tau,
                              float tau_min, int32_t channels, int32_t heads, void* stream) {
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.o = (float*)o; a.lse = (float*)lse; a.qtok = qtok; a.qcnt = qcnt; a.ktok = ktok; a.kcnt = kcnt;
  a.n_win = n_win; a.tau = tau; a.tau_min = tau_min; a.C = channels; a.H = heads;
  a.dout = dout; a.dq = dq; a.dk = dk; a.dv = dv; a.dtau = dtau;
  int hd = channels / heads;
  TMAE_CHECK_ARG(check(a, hd) == 0, "channels must be a multiple of 128 and head_dim 16 or 32");
  if (max_windows <= 0) return 0;
  size_t smem = (size_t)(4 * MAXT * TWP + 4 * MAXT * 8) * sizeof(float);
  int64_t items = max_windows * (channels / TW);
  int grid = (int)(items < 2 * kNumSMs ? items : 2 * kNumSMs);
  cudaStream_t s = (cudaStream_t)stream;
  if (hd == 16) {
    TMAE_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_kernel<16><<<grid, ATT_THREADS, smem, s>>>(a);
  } else {
    TMAE_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_kernel<32><<<grid, ATT_THREADS, smem, s>>>(a);
  }
  TMAE_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
