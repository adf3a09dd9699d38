// Windowed cosine multi-head attention core, self (row A7) and temporal cross (row A8) forms.
//
// Replaces, per window: flat2window gathers, F.normalize, bmm(q,k^T)/clamp(tau), key-padding
// mask, softmax, bmm(p,v) and the window2flat scatter
// (pcdet/models/model_utils/cosine_msa.py:114-176,370-429; sst_basic_block.py:22-54;
//  wca_block.py:26-67; sst_utils.py:118-192).
//
// The reference pads every window to its level's token count and masks the padding; mathematically
// the padded keys contribute exp(-inf) = 0 and padded queries are discarded, so the kernels work
// on the ragged windows directly and write straight to the flat (voxel-major) layout.  The
// averaged attention map the reference also computes (cosine_msa.py:433-436) is never consumed and
// is not produced.
//
// Work decomposition.  An "item" is (row, head) where the row is a query (forward, dQ pass) or a
// key (dK/dV pass); the other side of the window is the "stationary" side it loops over.
//   * small windows (the partition sorts windows by level, so windows [0, small_end) hold <= 16
//     tokens -- 87 % of the windows of a lidar scan): one WARP per window, lanes = items, the
//     stationary rows are read through L1 (every row is re-read by all items of the warp).  No
//     shared memory, no block barriers, full occupancy.
//   * large windows: one CTA per (window, 128-channel group); the stationary rows are gathered
//     once into shared memory (head slices padded to HD+4 floats so that the 8 heads of a quarter
//     warp hit distinct banks with 128-bit loads), normalised there, and threads = items.
// Three passes share the template: MODE 0 forward (saves logsumexp), MODE 1 dQ (+ dtau, saves
// D = dO.O), MODE 2 dK/dV.  Online softmax in fp32; exp via __expf.
#include "common.cuh"

namespace tmae {

constexpr int MAXT = TMAE_WIN_TOKENS;
constexpr int ATT_THREADS = 256;
constexpr int TW = 128;  // channels per large-window work item

struct AttnArgs {
  const float* q; const float* k; const float* v;   // (rows, C) projected
  float* o;                                          // (q rows, C)
  float* lse;                                        // (q rows, H)
  float* dsum;                                       // (q rows, H)  D = dO . O   (written by MODE 1, read by MODE 2)
  const int* qtok; const int* qcnt;                  // [n_win*64], [n_win]
  const int* ktok; const int* kcnt;
  const int* n_win;                                  // device: number of windows
  const int* small_end;                              // device: windows [0, *small_end) hold <= 16 tokens on both sides
  const float* tau; float tau_min;
  int C, H;
  const float* dout; float* dq; float* dk; float* dv; float* dtau;
};

template <int HD>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float* r) {
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p + d));
    r[d] = t.x; r[d + 1] = t.y; r[d + 2] = t.z; r[d + 3] = t.w;
  }
}
template <int HD>
__device__ __forceinline__ void store_row(float* p, const float* r) {
#pragma unroll
  for (int d = 0; d < HD; d += 4) *reinterpret_cast<float4*>(p + d) = make_float4(r[d], r[d + 1], r[d + 2], r[d + 3]);
}
template <int HD>
__device__ __forceinline__ float dot(const float* a, const float* b) {
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < HD; ++d) s = fmaf(a[d], b[d], s);
  return s;
}
// x / max(|x|, 1e-12)  (F.normalize); returns 1 / max(|x|, eps)
template <int HD>
__device__ __forceinline__ float normalize(float* x) {
  float inv = 1.f / fmaxf(sqrtf(dot<HD>(x, x)), 1e-12f);
#pragma unroll
  for (int d = 0; d < HD; ++d) x[d] *= inv;
  return inv;
}

// One item against a stationary side given by an accessor.  ST::row(j, h, out) yields the j-th stationary row's
// head slice; for MODE 0/1 that is (k_hat via .a, v via .b); for MODE 2 it is (q_hat via .a, dO via .b).
template <int HD, int MODE, class ST>
__device__ __forceinline__ void run_item(const AttnArgs& a, const ST& st, int n_other, int64_t row, int h, float inv_tau,
                                         float& dtau_acc) {
  const int64_t off = row * a.C + h * HD;
  if (MODE == 0) {
    float qh[HD];
    load_row<HD>(a.q + off, qh);
    float s = inv_tau / fmaxf(sqrtf(dot<HD>(qh, qh)), 1e-12f);
#pragma unroll
    for (int d = 0; d < HD; ++d) qh[d] *= s;  // q_hat / tau
    float m = -INFINITY, l = 0.f, acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
    for (int j = 0; j < n_other; ++j) {
      float kh[HD], vv[HD];
      st.a(j, h, kh);
      st.b(j, h, vv);
      float sc = dot<HD>(qh, kh);
      if (sc > m) {
        float r = __expf(m - sc);
        l *= r;
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] *= r;
        m = sc;
      }
      float p = __expf(sc - m);
      l += p;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] = fmaf(p, vv[d], acc[d]);
    }
    float il = 1.f / l;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] *= il;
    store_row<HD>(a.o + off, acc);
    if (a.lse) a.lse[row * a.H + h] = m + __logf(l);
  } else if (MODE == 1) {
    float qh[HD], dov[HD], dqh[HD], ov[HD];
    load_row<HD>(a.q + off, qh);
    float inv = normalize<HD>(qh);
    load_row<HD>(a.dout + off, dov);
    load_row<HD>(a.o + off, ov);
    float Di = dot<HD>(dov, ov), Li = a.lse[row * a.H + h];
    a.dsum[row * a.H + h] = Di;
#pragma unroll
    for (int d = 0; d < HD; ++d) dqh[d] = 0.f;
    for (int j = 0; j < n_other; ++j) {
      float kh[HD], vv[HD];
      st.a(j, h, kh);
      st.b(j, h, vv);
      float sc = dot<HD>(qh, kh) * inv_tau;
      float p = __expf(sc - Li);
      float ds = p * (dot<HD>(dov, vv) - Di);
      dtau_acc -= ds * sc;  // d(c/tau)/dtau = -(c/tau)/tau ; the trailing 1/tau is applied once at the end
      float dsl = ds * inv_tau;
#pragma unroll
      for (int d = 0; d < HD; ++d) dqh[d] = fmaf(dsl, kh[d], dqh[d]);
    }
    float dt = dot<HD>(dqh, qh);  // through q_hat = q / max(|q|, eps)
#pragma unroll
    for (int d = 0; d < HD; ++d) dqh[d] = (dqh[d] - qh[d] * dt) * inv;
    store_row<HD>(a.dq + off, dqh);
  } else {
    float kh[HD], vv[HD], dkh[HD], dvv[HD];
    load_row<HD>(a.k + off, kh);
    float inv = normalize<HD>(kh);
    load_row<HD>(a.v + off, vv);
#pragma unroll
    for (int d = 0; d < HD; ++d) { dkh[d] = 0.f; dvv[d] = 0.f; }
    for (int i = 0; i < n_other; ++i) {
      float qh[HD], dov[HD];
      st.a(i, h, qh);
      st.b(i, h, dov);
      float sc = dot<HD>(qh, kh) * inv_tau;
      float p = __expf(sc - st.lse(i, h));
      float dsl = p * (dot<HD>(dov, vv) - st.dsum(i, h)) * inv_tau;
#pragma unroll
      for (int d = 0; d < HD; ++d) { dkh[d] = fmaf(dsl, qh[d], dkh[d]); dvv[d] = fmaf(p, dov[d], dvv[d]); }
    }
    float dt = dot<HD>(dkh, kh);
#pragma unroll
    for (int d = 0; d < HD; ++d) dkh[d] = (dkh[d] - kh[d] * dt) * inv;
    store_row<HD>(a.dk + off, dkh);
    store_row<HD>(a.dv + off, dvv);
  }
}

// ------------------------------------------------------------------ small windows: warp per window, L1-resident rows
template <int HD, int MODE>
struct GlobalSide {
  const AttnArgs& g;
  const int* tok;  // stationary token list of this window
  __device__ __forceinline__ void a(int j, int h, float* out) const {
    const float* src = (MODE == 2 ? g.q : g.k) + (int64_t)tok[j] * g.C + h * HD;
    load_row<HD>(src, out);
    normalize<HD>(out);
  }
  __device__ __forceinline__ void b(int j, int h, float* out) const {
    load_row<HD>((MODE == 2 ? g.dout : g.v) + (int64_t)tok[j] * g.C + h * HD, out);
  }
  __device__ __forceinline__ float lse(int i, int h) const { return g.lse[(int64_t)tok[i] * g.H + h]; }
  __device__ __forceinline__ float dsum(int i, int h) const { return g.dsum[(int64_t)tok[i] * g.H + h]; }
};

template <int HD, int MODE>
__global__ void __launch_bounds__(ATT_THREADS) attn_small_kernel(AttnArgs a) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int n_small = min(*a.small_end, *a.n_win);
  const float tau_raw = *a.tau;
  const float inv_tau = 1.f / fmaxf(tau_raw, a.tau_min);
  float dtau_acc = 0.f;
  for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_small; g += warps) {
    const int nq = a.qcnt[g], nk = a.kcnt[g];
    const int* rows_tok = (MODE == 2 ? a.ktok : a.qtok) + g * MAXT;
    const int n_rows = MODE == 2 ? nk : nq, n_other = MODE == 2 ? nq : nk;
    GlobalSide<HD, MODE> st{a, (MODE == 2 ? a.qtok : a.ktok) + g * MAXT};
    for (int it = lane; it < n_rows * a.H; it += 32) {
      int r = it / a.H, h = it - r * a.H;
      run_item<HD, MODE>(a, st, n_other, rows_tok[r], h, inv_tau, dtau_acc);
    }
  }
  if (MODE == 1) {
    dtau_acc = warp_sum(dtau_acc);
    if (lane == 0 && a.dtau && tau_raw > a.tau_min && dtau_acc != 0.f) atomicAdd(a.dtau, dtau_acc * inv_tau);
  }
}

// ------------------------------------------------------------------ large windows: CTA per (window, 128-channel group)
template <int HD>
struct SmemSide {
  static constexpr int HEADS = TW / HD;
  static constexpr int HS = HD + 4;            // padded head stride: conflict-free LDS.128 across heads
  static constexpr int RS = HEADS * HS;        // row stride
  const float* A; const float* B; const float* L; const float* D;
  int h0;  // first absolute head of this 128-channel group
  __device__ __forceinline__ void a(int j, int h, float* out) const {
    h -= h0;
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      float4 t = *reinterpret_cast<const float4*>(A + j * RS + h * HS + d);
      out[d] = t.x; out[d + 1] = t.y; out[d + 2] = t.z; out[d + 3] = t.w;
    }
  }
  __device__ __forceinline__ void b(int j, int h, float* out) const {
    h -= h0;
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      float4 t = *reinterpret_cast<const float4*>(B + j * RS + h * HS + d);
      out[d] = t.x; out[d + 1] = t.y; out[d + 2] = t.z; out[d + 3] = t.w;
    }
  }
  __device__ __forceinline__ float lse(int i, int h) const { return L[i * HEADS + h - h0]; }
  __device__ __forceinline__ float dsum(int i, int h) const { return D[i * HEADS + h - h0]; }
};

template <int HD, int MODE>
__global__ void __launch_bounds__(ATT_THREADS) attn_large_kernel(AttnArgs a) {
  using S = SmemSide<HD>;
  extern __shared__ __align__(16) float sm[];
  float* As = sm;                       // [64][RS]  k_hat (MODE 0/1) or q_hat (MODE 2)
  float* Bs = sm + MAXT * S::RS;        // [64][RS]  v or dO
  float* Ls = Bs + MAXT * S::RS;        // [64][HEADS] lse   (MODE 2)
  float* Ds = Ls + MAXT * S::HEADS;     // [64][HEADS] D     (MODE 2)
  __shared__ int stat_tok[MAXT], row_tok[MAXT];
  const int groups = a.C / TW;
  const int first = min(*a.small_end, *a.n_win);
  const int n_items = (*a.n_win - first) * groups;
  const float tau_raw = *a.tau;
  const float inv_tau = 1.f / fmaxf(tau_raw, a.tau_min);
  float dtau_acc = 0.f;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int g = first + item / groups, grp = item % groups, col0 = grp * TW;
    const int nq = a.qcnt[g], nk = a.kcnt[g];
    const int n_rows = MODE == 2 ? nk : nq, n_other = MODE == 2 ? nq : nk;
    __syncthreads();
    if (threadIdx.x < MAXT) {
      int t = threadIdx.x;
      stat_tok[t] = t < n_other ? ((MODE == 2 ? a.qtok : a.ktok)[g * MAXT + t]) : 0;
      row_tok[t] = t < n_rows ? ((MODE == 2 ? a.ktok : a.qtok)[g * MAXT + t]) : 0;
    }
    __syncthreads();
    const float* srcA = MODE == 2 ? a.q : a.k;
    const float* srcB = MODE == 2 ? a.dout : a.v;
    for (int e = threadIdx.x; e < n_other * (TW / 4); e += ATT_THREADS) {  // coalesced 512-byte row segments
      int j = e / (TW / 4), c4 = (e - j * (TW / 4)) * 4;
      int h = c4 / HD, d = c4 - h * HD;
      const int64_t src = (int64_t)stat_tok[j] * a.C + col0 + c4;
      *reinterpret_cast<float4*>(As + j * S::RS + h * S::HS + d) = __ldg(reinterpret_cast<const float4*>(srcA + src));
      *reinterpret_cast<float4*>(Bs + j * S::RS + h * S::HS + d) = __ldg(reinterpret_cast<const float4*>(srcB + src));
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n_other * S::HEADS; e += ATT_THREADS) {  // F.normalize per (row, head)
      int j = e / S::HEADS, h = e - j * S::HEADS;
      float* p = As + j * S::RS + h * S::HS;
      float s = 0.f;
      for (int d = 0; d < HD; ++d) s = fmaf(p[d], p[d], s);
      float inv = 1.f / fmaxf(sqrtf(s), 1e-12f);
      for (int d = 0; d < HD; ++d) p[d] *= inv;
      if (MODE == 2) {
        const int64_t r = (int64_t)stat_tok[j] * a.H + grp * S::HEADS + h;
        Ls[e] = a.lse[r];
        Ds[e] = a.dsum[r];
      }
    }
    __syncthreads();
    S st{As, Bs, Ls, Ds, grp * S::HEADS};
    for (int it = threadIdx.x; it < n_rows * S::HEADS; it += ATT_THREADS) {
      int r = it / S::HEADS, h = it - r * S::HEADS;
      run_item<HD, MODE>(a, st, n_other, row_tok[r], grp * S::HEADS + h, inv_tau, dtau_acc);
    }
  }
  if (MODE == 1) {
    dtau_acc = warp_sum(dtau_acc);
    if ((threadIdx.x & 31) == 0 && a.dtau && tau_raw > a.tau_min && dtau_acc != 0.f) atomicAdd(a.dtau, dtau_acc * inv_tau);
  }
}

template <int HD, int MODE>
static int launch_pass(const AttnArgs& a, int64_t max_windows, cudaStream_t s) {
  using S = SmemSide<HD>;
  // small windows: 8 warps per CTA, one window per warp per iteration
  int64_t warps = max_windows < (int64_t)kNumSMs * 32 ? max_windows : (int64_t)kNumSMs * 32;
  attn_small_kernel<HD, MODE><<<cdiv(warps * 32, ATT_THREADS), ATT_THREADS, 0, s>>>(a);
  size_t smem = (size_t)(2 * MAXT * S::RS + 2 * MAXT * S::HEADS) * sizeof(float);
  auto kern = attn_large_kernel<HD, MODE>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TMAE_ERR_CUDA;
  int64_t items = max_windows * (a.C / TW);
  int grid = (int)(items < 2 * kNumSMs ? items : 2 * kNumSMs);
  kern<<<grid, ATT_THREADS, smem, s>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : TMAE_ERR_CUDA;
}

static int check(const AttnArgs& a, int hd) {
  if (a.C % TW != 0 || a.H != a.C / hd || (hd != 16 && hd != 32)) return -1;
  return 0;
}

}  // namespace tmae

using namespace tmae;

extern "C" {

int tmae_window_attention_fwd(const float* q, const float* k, const float* v, float* o, float* lse, const int32_t* qtok,
                              const int32_t* qcnt, const int32_t* ktok, const int32_t* kcnt, const int32_t* n_win,
                              const int32_t* small_end, int64_t max_windows, const float* tau, float tau_min, int32_t channels,
                              int32_t heads, void* stream) {
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.o = o; a.lse = lse; a.qtok = qtok; a.qcnt = qcnt; a.ktok = ktok; a.kcnt = kcnt; a.n_win = n_win;
  a.small_end = small_end; a.tau = tau; a.tau_min = tau_min; a.C = channels; a.H = heads;
  int hd = channels / heads;
  TMAE_CHECK_ARG(check(a, hd) == 0, "channels must be a multiple of 128 and head_dim 16 or 32");
  TMAE_CHECK_ARG(small_end != nullptr && n_win != nullptr, "n_win / small_end must be device pointers");
  if (max_windows <= 0) return 0;
  int r = hd == 16 ? launch_pass<16, 0>(a, max_windows, (cudaStream_t)stream) : launch_pass<32, 0>(a, max_windows, (cudaStream_t)stream);
  if (r) { set_error("tmae_window_attention_fwd: launch failed"); return r; }
  return 0;
}

int tmae_window_attention_bwd(const float* dout, const float* q, const float* k, const float* v, const float* o, const float* lse,
                              float* dsum, float* dq, float* dk, float* dv, float* dtau, const int32_t* qtok, const int32_t* qcnt,
                              const int32_t* ktok, const int32_t* kcnt, const int32_t* n_win, const int32_t* small_end,
                              int64_t max_windows, const float* tau, float tau_min, int32_t channels, int32_t heads, void* stream) {
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.o = (float*)o; a.lse = (float*)lse; a.dsum = dsum; a.qtok = qtok; a.qcnt = qcnt; a.ktok = ktok;
  a.kcnt = kcnt; a.n_win = n_win; a.small_end = small_end; a.tau = tau; a.tau_min = tau_min; a.C = channels; a.H = heads;
  a.dout = dout; a.dq = dq; a.dk = dk; a.dv = dv; a.dtau = dtau;
  int hd = channels / heads;
  TMAE_CHECK_ARG(check(a, hd) == 0, "channels must be a multiple of 128 and head_dim 16 or 32");
  TMAE_CHECK_ARG(small_end != nullptr && n_win != nullptr && dsum != nullptr, "n_win / small_end / dsum must be device pointers");
  if (max_windows <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  int r = hd == 16 ? launch_pass<16, 1>(a, max_windows, s) : launch_pass<32, 1>(a, max_windows, s);
  if (!r) r = hd == 16 ? launch_pass<16, 2>(a, max_windows, s) : launch_pass<32, 2>(a, max_windows, s);
  if (r) { set_error("tmae_window_attention_bwd: launch failed"); return r; }
  return 0;
}

}  // extern "C"
