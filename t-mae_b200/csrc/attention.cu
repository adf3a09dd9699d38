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
// Work decomposition.  An "item" is (row, 16-channel slice) where the row is a query (forward, dQ
// pass) or a key (dK/dV pass); a head of 32 channels is shared by two adjacent lanes that combine
// their partial dot products with one shuffle.  The other side of the window is the "stationary"
// side every item loops over.
//   * small windows (the partition sorts windows by level, so windows [0, small_end) hold <= 16
//     tokens -- 87 % of the windows of a lidar scan): one WARP per window, lanes = items, the
//     stationary rows are read through L1 (every row is re-read by all items of the warp).  No
//     shared memory, no block barriers, full occupancy.
//   * larger windows: one 128-thread CTA per (window, 128-channel group); the stationary rows are
//     gathered once into shared memory (slices padded to 20 floats so the 8 slices read by a quarter
//     warp hit distinct banks with 128-bit loads), normalised there, threads = items.  Shared memory
//     is sized for 32 tokens (windows [small_end, mid_end)) or 64 (the rest).
// Three passes share the template: MODE 0 forward (saves logsumexp), MODE 1 dQ (+ dtau, saves
// D = dO.O), MODE 2 dK/dV.  Online softmax in fp32; exp via __expf.
#include "common.cuh"

namespace tmae {

constexpr int MAXT = TMAE_WIN_TOKENS;
constexpr int ATT_THREADS = 256;    // small-window kernel: 8 windows per CTA
constexpr int LARGE_THREADS = 128;  // large-window kernel
constexpr int TW = 128;             // channels per large-window work item
constexpr int HT = 16;              // channels held by one thread

struct AttnArgs {
  const float* q; const float* k; const float* v;   // (rows, C) projected
  float* o;                                          // (q rows, C)
  float* lse;                                        // (q rows, H)
  float* dsum;                                       // (q rows, H)  D = dO . O   (written by MODE 1, read by MODE 2)
  const int* qtok; const int* qcnt;                  // [n_win*64], [n_win]
  const int* ktok; const int* kcnt;
  const int* n_win;                                  // device: number of windows
  const int* small_end;                              // device: windows [0, *small_end) hold <= 16 tokens on both sides
  const int* mid_end;                                // device: windows [*small_end, *mid_end) hold <= 32 tokens
  const float* tau; float tau_min;
  int C, H;
  const float* dout; float* dq; float* dk; float* dv; float* dtau;
};

__device__ __forceinline__ void load_row(const float* __restrict__ p, float* r) {
#pragma unroll
  for (int d = 0; d < HT; d += 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p + d));
    r[d] = t.x; r[d + 1] = t.y; r[d + 2] = t.z; r[d + 3] = t.w;
  }
}
__device__ __forceinline__ void store_row(float* p, const float* r) {
#pragma unroll
  for (int d = 0; d < HT; d += 4) *reinterpret_cast<float4*>(p + d) = make_float4(r[d], r[d + 1], r[d + 2], r[d + 3]);
}
// dot product over a full head: each of the SPLIT adjacent lanes holds HT channels
template <int SPLIT>
__device__ __forceinline__ float dot(const float* a, const float* b, unsigned mask) {
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < HT; ++d) s = fmaf(a[d], b[d], s);
  if (SPLIT == 2) s += __shfl_xor_sync(mask, s, 1);
  return s;
}
// x / max(|x|, 1e-12)  (F.normalize); returns 1 / max(|x|, eps)
template <int SPLIT>
__device__ __forceinline__ float normalize(float* x, unsigned mask) {
  float inv = 1.f / fmaxf(sqrtf(dot<SPLIT>(x, x, mask)), 1e-12f);
#pragma unroll
  for (int d = 0; d < HT; ++d) x[d] *= inv;
  return inv;
}

// One item against a stationary side given by an accessor.  For MODE 0/1 the accessor yields (k_hat via .a, v via
// .b); for MODE 2 (q_hat via .a, dO via .b).  `col` is the item's first channel.  Every lane of `mask` runs the same
// trip count (same window), so the pair shuffles inside dot() are convergent.
template <int SPLIT, int MODE, class ST>
__device__ __forceinline__ void run_item(const AttnArgs& a, const ST& st, int n_other, int64_t row, int col, int head, float inv_tau,
                                         float& dtau_acc, unsigned mask) {
  const int64_t off = row * a.C + col;
  const bool lead = SPLIT == 1 || (col % (HT * SPLIT)) == 0;  // one lane of a pair owns the per-head scalars
  if (MODE == 0) {
    float qh[HT];
    load_row(a.q + off, qh);
    float s = inv_tau / fmaxf(sqrtf(dot<SPLIT>(qh, qh, mask)), 1e-12f);
#pragma unroll
    for (int d = 0; d < HT; ++d) qh[d] *= s;  // q_hat / tau
    float m = -INFINITY, l = 0.f, acc[HT];
#pragma unroll
    for (int d = 0; d < HT; ++d) acc[d] = 0.f;
    for (int j = 0; j < n_other; ++j) {
      float kh[HT], vv[HT];
      st.a(j, col, kh, mask);
      st.b(j, col, vv);
      float sc = dot<SPLIT>(qh, kh, mask);
      if (sc > m) {
        float r = __expf(m - sc);
        l *= r;
#pragma unroll
        for (int d = 0; d < HT; ++d) acc[d] *= r;
        m = sc;
      }
      float p = __expf(sc - m);
      l += p;
#pragma unroll
      for (int d = 0; d < HT; ++d) acc[d] = fmaf(p, vv[d], acc[d]);
    }
    float il = 1.f / l;
#pragma unroll
    for (int d = 0; d < HT; ++d) acc[d] *= il;
    store_row(a.o + off, acc);
    if (a.lse && lead) a.lse[row * a.H + head] = m + __logf(l);
  } else if (MODE == 1) {
    float qh[HT], dov[HT], dqh[HT];
    load_row(a.q + off, qh);
    float inv = normalize<SPLIT>(qh, mask);
    load_row(a.dout + off, dov);
    float Di;
    {
      float ov[HT];
      load_row(a.o + off, ov);
      Di = dot<SPLIT>(dov, ov, mask);
    }
    const float Li = a.lse[row * a.H + head];
    if (lead) a.dsum[row * a.H + head] = Di;
#pragma unroll
    for (int d = 0; d < HT; ++d) dqh[d] = 0.f;
    for (int j = 0; j < n_other; ++j) {
      float kh[HT], vv[HT];
      st.a(j, col, kh, mask);
      st.b(j, col, vv);
      float sc = dot<SPLIT>(qh, kh, mask) * inv_tau;
      float p = __expf(sc - Li);
      float ds = p * (dot<SPLIT>(dov, vv, mask) - Di);
      if (lead) dtau_acc -= ds * sc;  // d(c/tau)/dtau = -(c/tau)/tau ; the trailing 1/tau is applied once at the end
      float dsl = ds * inv_tau;
#pragma unroll
      for (int d = 0; d < HT; ++d) dqh[d] = fmaf(dsl, kh[d], dqh[d]);
    }
    float dt = dot<SPLIT>(dqh, qh, mask);  // through q_hat = q / max(|q|, eps)
#pragma unroll
    for (int d = 0; d < HT; ++d) dqh[d] = (dqh[d] - qh[d] * dt) * inv;
    store_row(a.dq + off, dqh);
  } else {
    float kh[HT], vv[HT], dkh[HT], dvv[HT];
    load_row(a.k + off, kh);
    float inv = normalize<SPLIT>(kh, mask);
    load_row(a.v + off, vv);
#pragma unroll
    for (int d = 0; d < HT; ++d) { dkh[d] = 0.f; dvv[d] = 0.f; }
    for (int i = 0; i < n_other; ++i) {
      float qh[HT], dov[HT];
      st.a(i, col, qh, mask);
      st.b(i, col, dov);
      float sc = dot<SPLIT>(qh, kh, mask) * inv_tau;
      float p = __expf(sc - st.lse(i, head));
      float dsl = p * (dot<SPLIT>(dov, vv, mask) - st.dsum(i, head)) * inv_tau;
#pragma unroll
      for (int d = 0; d < HT; ++d) { dkh[d] = fmaf(dsl, qh[d], dkh[d]); dvv[d] = fmaf(p, dov[d], dvv[d]); }
    }
    float dt = dot<SPLIT>(dkh, kh, mask);
#pragma unroll
    for (int d = 0; d < HT; ++d) dkh[d] = (dkh[d] - kh[d] * dt) * inv;
    store_row(a.dk + off, dkh);
    store_row(a.dv + off, dvv);
  }
}

// ------------------------------------------------------------------ small windows: warp per window, L1-resident rows
template <int SPLIT, int MODE>
struct GlobalSide {
  const AttnArgs& g;
  const int* tok;  // stationary token list of this window
  __device__ __forceinline__ void a(int j, int col, float* out, unsigned mask) const {
    load_row((MODE == 2 ? g.q : g.k) + (int64_t)tok[j] * g.C + col, out);
    normalize<SPLIT>(out, mask);
  }
  __device__ __forceinline__ void b(int j, int col, float* out) const {
    load_row((MODE == 2 ? g.dout : g.v) + (int64_t)tok[j] * g.C + col, out);
  }
  __device__ __forceinline__ float lse(int i, int h) const { return g.lse[(int64_t)tok[i] * g.H + h]; }
  __device__ __forceinline__ float dsum(int i, int h) const { return g.dsum[(int64_t)tok[i] * g.H + h]; }
};

template <int HD, int MODE>
__global__ void __launch_bounds__(ATT_THREADS) attn_small_kernel(AttnArgs a) {
  constexpr int SPLIT = HD / HT;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int n_small = min(*a.small_end, *a.n_win);
  const float tau_raw = *a.tau;
  const float inv_tau = 1.f / fmaxf(tau_raw, a.tau_min);
  const int slices = a.C / HT;  // items per row
  float dtau_acc = 0.f;
  for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_small; g += warps) {
    const int nq = a.qcnt[g], nk = a.kcnt[g];
    const int* rows_tok = (MODE == 2 ? a.ktok : a.qtok) + g * MAXT;
    const int n_rows = MODE == 2 ? nk : nq, n_other = MODE == 2 ? nq : nk;
    GlobalSide<SPLIT, MODE> st{a, (MODE == 2 ? a.qtok : a.ktok) + g * MAXT};
    const int n_items = n_rows * slices;
    for (int base = 0; base < n_items; base += 32) {
      const int it = base + lane;
      const unsigned mask = __ballot_sync(0xffffffffu, it < n_items);
      if (it < n_items) {
        int r = it / slices, sl = it - r * slices;
        run_item<SPLIT, MODE>(a, st, n_other, rows_tok[r], sl * HT, sl / SPLIT, inv_tau, dtau_acc, mask);
      }
    }
  }
  if (MODE == 1) {
    dtau_acc = warp_sum(dtau_acc);
    if (lane == 0 && a.dtau && tau_raw > a.tau_min && dtau_acc != 0.f) atomicAdd(a.dtau, dtau_acc * inv_tau);
  }
}

// ------------------------------------------------------------------ larger windows: CTA per (window, 128-channel group)
constexpr int SLICES = TW / HT;  // 8
constexpr int SS = HT + 4;       // padded slice stride
constexpr int RS = SLICES * SS;  // row stride (160 floats)

template <int SPLIT>
struct SmemSide {
  const float* A; const float* B; const float* L; const float* D;
  int col0, heads;  // first channel of this group; heads per group
  __device__ __forceinline__ void a(int j, int col, float* out, unsigned) const {
    const float* p = A + j * RS + ((col - col0) / HT) * SS;
#pragma unroll
    for (int d = 0; d < HT; d += 4) {
      float4 t = *reinterpret_cast<const float4*>(p + d);
      out[d] = t.x; out[d + 1] = t.y; out[d + 2] = t.z; out[d + 3] = t.w;
    }
  }
  __device__ __forceinline__ void b(int j, int col, float* out) const {
    const float* p = B + j * RS + ((col - col0) / HT) * SS;
#pragma unroll
    for (int d = 0; d < HT; d += 4) {
      float4 t = *reinterpret_cast<const float4*>(p + d);
      out[d] = t.x; out[d + 1] = t.y; out[d + 2] = t.z; out[d + 3] = t.w;
    }
  }
  __device__ __forceinline__ float lse(int i, int h) const { return L[i * heads + h - col0 / (HT * SPLIT)]; }
  __device__ __forceinline__ float dsum(int i, int h) const { return D[i * heads + h - col0 / (HT * SPLIT)]; }
};

// windows [*begin, *end) of the level-sorted list; TCAP = max tokens per side in that range
template <int HD, int MODE, int TCAP>
__global__ void __launch_bounds__(LARGE_THREADS) attn_large_kernel(AttnArgs a, const int* begin, const int* end) {
  constexpr int SPLIT = HD / HT;
  constexpr int HEADS = TW / HD;
  extern __shared__ __align__(16) float sm[];
  float* As = sm;                 // [TCAP][RS]  k_hat (MODE 0/1) or q_hat (MODE 2)
  float* Bs = sm + TCAP * RS;     // [TCAP][RS]  v or dO
  float* Ls = Bs + TCAP * RS;     // [TCAP][HEADS] lse   (MODE 2)
  float* Ds = Ls + TCAP * HEADS;  // [TCAP][HEADS] D     (MODE 2)
  __shared__ int stat_tok[TCAP], row_tok[TCAP];
  const int groups = a.C / TW;
  const int nw = *a.n_win;
  const int first = min(*begin, nw), last = min(*end, nw);
  const int n_items = (last - first) * groups;
  const float tau_raw = *a.tau;
  const float inv_tau = 1.f / fmaxf(tau_raw, a.tau_min);
  float dtau_acc = 0.f;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int g = first + item / groups, grp = item % groups, col0 = grp * TW;
    const int nq = min(a.qcnt[g], TCAP), nk = min(a.kcnt[g], TCAP);
    const int n_rows = MODE == 2 ? nk : nq, n_other = MODE == 2 ? nq : nk;
    __syncthreads();
    if (threadIdx.x < TCAP) {
      int t = threadIdx.x;
      stat_tok[t] = t < n_other ? ((MODE == 2 ? a.qtok : a.ktok)[g * MAXT + t]) : 0;
      row_tok[t] = t < n_rows ? ((MODE == 2 ? a.ktok : a.qtok)[g * MAXT + t]) : 0;
    }
    __syncthreads();
    const float* srcA = MODE == 2 ? a.q : a.k;
    const float* srcB = MODE == 2 ? a.dout : a.v;
    for (int e = threadIdx.x; e < n_other * (TW / 4); e += LARGE_THREADS) {  // coalesced 512-byte row segments
      int j = e / (TW / 4), c4 = (e - j * (TW / 4)) * 4;
      int sl = c4 / HT, d = c4 - sl * HT;
      const int64_t src = (int64_t)stat_tok[j] * a.C + col0 + c4;
      *reinterpret_cast<float4*>(As + j * RS + sl * SS + d) = __ldg(reinterpret_cast<const float4*>(srcA + src));
      *reinterpret_cast<float4*>(Bs + j * RS + sl * SS + d) = __ldg(reinterpret_cast<const float4*>(srcB + src));
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n_other * HEADS; e += LARGE_THREADS) {  // F.normalize per (row, head)
      int j = e / HEADS, h = e - j * HEADS;
      float s = 0.f;
#pragma unroll
      for (int u = 0; u < SPLIT; ++u) {
        const float* p = As + j * RS + (h * SPLIT + u) * SS;
#pragma unroll
        for (int d = 0; d < HT; ++d) s = fmaf(p[d], p[d], s);
      }
      float inv = 1.f / fmaxf(sqrtf(s), 1e-12f);
#pragma unroll
      for (int u = 0; u < SPLIT; ++u) {
        float* p = As + j * RS + (h * SPLIT + u) * SS;
#pragma unroll
        for (int d = 0; d < HT; ++d) p[d] *= inv;
      }
      if (MODE == 2) {
        const int64_t r = (int64_t)stat_tok[j] * a.H + grp * HEADS + h;
        Ls[e] = a.lse[r];
        Ds[e] = a.dsum[r];
      }
    }
    __syncthreads();
    SmemSide<SPLIT> st{As, Bs, Ls, Ds, col0, HEADS};
    const int n_it = n_rows * SLICES;
    for (int base = 0; base < n_it; base += LARGE_THREADS) {
      const int it = base + threadIdx.x;
      const unsigned mask = __ballot_sync(0xffffffffu, it < n_it);
      if (it < n_it) {
        int r = it / SLICES, sl = it - r * SLICES;
        run_item<SPLIT, MODE>(a, st, n_other, row_tok[r], col0 + sl * HT, (col0 + sl * HT) / HD, inv_tau, dtau_acc, mask);
      }
    }
  }
  if (MODE == 1) {
    dtau_acc = warp_sum(dtau_acc);
    if ((threadIdx.x & 31) == 0 && a.dtau && tau_raw > a.tau_min && dtau_acc != 0.f) atomicAdd(a.dtau, dtau_acc * inv_tau);
  }
}

template <int HD, int MODE, int TCAP>
static int launch_large(const AttnArgs& a, const int* begin, const int* end, int64_t max_windows, cudaStream_t s) {
  constexpr int HEADS = TW / HD;
  size_t smem = (size_t)(2 * TCAP * RS + 2 * TCAP * HEADS) * sizeof(float);
  auto kern = attn_large_kernel<HD, MODE, TCAP>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TMAE_ERR_CUDA;
    attr_set = true;
  }
  int64_t items = max_windows * (a.C / TW);
  int per_sm = TCAP == 32 ? 5 : 2;
  int grid = (int)(items < (int64_t)per_sm * kNumSMs ? items : (int64_t)per_sm * kNumSMs);
  kern<<<grid, LARGE_THREADS, smem, s>>>(a, begin, end);
  return cudaGetLastError() == cudaSuccess ? 0 : TMAE_ERR_CUDA;
}

// tensor-core path for windows with more than 16 tokens (attention_mma.cu)
struct AttnMmaArgs {
  const float* q; const float* k; const float* v; float* o; float* lse;
  const int* qtok; const int* qcnt; const int* ktok; const int* kcnt;
  const int* n_win; const int* begin;
  const float* tau; float tau_min;
  int C, H;
  const float* dout; float* dq; float* dk; float* dv; float* dtau;
};
int attn_mma_fwd(const AttnMmaArgs& a, int hd, int64_t max_windows, cudaStream_t s);
int attn_mma_bwd(const AttnMmaArgs& a, int hd, int64_t max_windows, cudaStream_t s);
bool g_attn_tc = false;  // set by the layer entry points (tensor-core precision mode) or tmae_set_option("attn_tc", 1)

static AttnMmaArgs to_mma(const AttnArgs& a) {
  AttnMmaArgs m{};
  m.q = a.q; m.k = a.k; m.v = a.v; m.o = a.o; m.lse = a.lse; m.qtok = a.qtok; m.qcnt = a.qcnt; m.ktok = a.ktok; m.kcnt = a.kcnt;
  m.n_win = a.n_win; m.begin = a.small_end; m.tau = a.tau; m.tau_min = a.tau_min; m.C = a.C; m.H = a.H;
  m.dout = a.dout; m.dq = a.dq; m.dk = a.dk; m.dv = a.dv; m.dtau = a.dtau;
  return m;
}

template <int HD, int MODE>
static int launch_pass(const AttnArgs& a, int64_t max_windows, cudaStream_t s) {
  static const char* names[3] = {"attn_fwd", "attn_bwd_dq", "attn_bwd_dkv"};
  // algorithmic traffic: fwd reads q,k,v writes o ; dq pass reads q,k,v,o,dO writes dq ; dkv pass reads q,k,v,dO writes dk,dv
  const double rq = g_prof_rows_hint[0] * a.C * 4.0, rk = g_prof_rows_hint[1] * a.C * 4.0;
  const double bytes = MODE == 0 ? 2 * rq + 2 * rk : (MODE == 1 ? 4 * rq + 2 * rk : 2 * rq + 4 * rk);
  ProfScope prof(names[MODE], 0, bytes, s);
  // small windows: 8 warps per CTA, one window per warp per iteration
  int64_t warps = max_windows < (int64_t)kNumSMs * 48 ? max_windows : (int64_t)kNumSMs * 48;
  attn_small_kernel<HD, MODE><<<cdiv(warps * 32, ATT_THREADS), ATT_THREADS, 0, s>>>(a);
  if (g_attn_tc) {  // windows above 16 tokens on mma.sync TF32: forward, and ONE fused backward pass (dQ, dK, dV, dtau)
    if (MODE == 0) return attn_mma_fwd(to_mma(a), HD, max_windows, s);
    if (MODE == 1) return attn_mma_bwd(to_mma(a), HD, max_windows, s);
    return cudaGetLastError() == cudaSuccess ? 0 : TMAE_ERR_CUDA;
  }
  int r = launch_large<HD, MODE, 32>(a, a.small_end, a.mid_end, max_windows, s);
  if (!r) r = launch_large<HD, MODE, 64>(a, a.mid_end, a.n_win, max_windows, s);
  return r;
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
                              const int32_t* small_end, const int32_t* mid_end, int64_t max_windows, const float* tau, float tau_min,
                              int32_t channels, int32_t heads, void* stream) {
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.o = o; a.lse = lse; a.qtok = qtok; a.qcnt = qcnt; a.ktok = ktok; a.kcnt = kcnt; a.n_win = n_win;
  a.small_end = small_end; a.mid_end = mid_end; a.tau = tau; a.tau_min = tau_min; a.C = channels; a.H = heads;
  int hd = channels / heads;
  TMAE_CHECK_ARG(check(a, hd) == 0, "channels must be a multiple of 128 and head_dim 16 or 32");
  TMAE_CHECK_ARG(small_end && mid_end && n_win, "n_win / small_end / mid_end must be device pointers");
  if (max_windows <= 0) return 0;
  int r = hd == 16 ? launch_pass<16, 0>(a, max_windows, (cudaStream_t)stream) : launch_pass<32, 0>(a, max_windows, (cudaStream_t)stream);
  if (r) { set_error("tmae_window_attention_fwd: launch failed"); return r; }
  return 0;
}

int tmae_window_attention_bwd(const float* dout, const float* q, const float* k, const float* v, const float* o, const float* lse,
                              float* dsum, float* dq, float* dk, float* dv, float* dtau, const int32_t* qtok, const int32_t* qcnt,
                              const int32_t* ktok, const int32_t* kcnt, const int32_t* n_win, const int32_t* small_end,
                              const int32_t* mid_end, int64_t max_windows, const float* tau, float tau_min, int32_t channels,
                              int32_t heads, void* stream) {
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.o = (float*)o; a.lse = (float*)lse; a.dsum = dsum; a.qtok = qtok; a.qcnt = qcnt; a.ktok = ktok;
  a.kcnt = kcnt; a.n_win = n_win; a.small_end = small_end; a.mid_end = mid_end; a.tau = tau; a.tau_min = tau_min; a.C = channels;
  a.H = heads; a.dout = dout; a.dq = dq; a.dk = dk; a.dv = dv; a.dtau = dtau;
  int hd = channels / heads;
  TMAE_CHECK_ARG(check(a, hd) == 0, "channels must be a multiple of 128 and head_dim 16 or 32");
  TMAE_CHECK_ARG(small_end && mid_end && n_win && dsum, "n_win / small_end / mid_end / dsum must be device pointers");
  if (max_windows <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  int r = hd == 16 ? launch_pass<16, 1>(a, max_windows, s) : launch_pass<32, 1>(a, max_windows, s);
  if (!r) r = hd == 16 ? launch_pass<16, 2>(a, max_windows, s) : launch_pass<32, 2>(a, max_windows, s);
  if (r) { set_error("tmae_window_attention_bwd: launch failed"); return r; }
  return 0;
}

}  // extern "C"
