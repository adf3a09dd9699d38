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
//     tokens -- 87 % of the windows of a lidar scan): register-resident warp kernels, see below.
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
  int ldq, ldk, ldv;                                 // row pitches (elements) of q / k / v and of dq / dk / dv; o, dO: C
  const float* dout; float* dq; float* dk; float* dv; float* dtau;
  int tc;                                            // host only: 1 = tensor-core precision mode (MUFU math, mma.sync for > 16 tokens)
  double rows_q, rows_kv;                            // host only: row counts for the profiler's algorithmic byte count (0 = unknown)
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
  const int64_t off = row * a.C + col;                       // o, dO
  const int64_t offq = row * a.ldq + col, offk = row * a.ldk + col, offv = row * a.ldv + col;
  const bool lead = SPLIT == 1 || (col % (HT * SPLIT)) == 0;  // one lane of a pair owns the per-head scalars
  if (MODE == 0) {
    float qh[HT];
    load_row(a.q + offq, qh);
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
    load_row(a.q + offq, qh);
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
    store_row(a.dq + offq, dqh);
  } else {
    float kh[HT], vv[HT], dkh[HT], dvv[HT];
    load_row(a.k + offk, kh);
    float inv = normalize<SPLIT>(kh, mask);
    load_row(a.v + offv, vv);
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
    store_row(a.dk + offk, dkh);
    store_row(a.dv + offv, dvv);
  }
}

// ------------------------------------------------------------------ small windows (<= 16 tokens per side): register-resident
// One WARP per (window, 128-channel group).  A lane owns 4 consecutive channels of every row, so a row is one coalesced
// 512-byte access and a head of HD channels spans HL = HD/4 adjacent lanes (per-head sums = HL-lane butterfly).  The
// stationary side (K_hat, V) is loaded ONCE into registers -- all row loads of the window are issued before the first
// use -- and the moving side streams through with one row of prefetch; every row of q/k/v/o/dO is read from memory
// exactly once (twice in the backward when the window has more than 8 keys) and every output row is written once.
constexpr int SW_T = 16;  // tokens per side handled here
constexpr int SW_THREADS = 128;

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ void axpy4(float4& y, float s, const float4& x) {
  y.x = fmaf(s, x.x, y.x); y.y = fmaf(s, x.y, y.y); y.z = fmaf(s, x.z, y.z); y.w = fmaf(s, x.w, y.w);
}
__device__ __forceinline__ void scale4(float4& y, float s) { y.x *= s; y.y *= s; y.z *= s; y.w *= s; }
template <int HL>
__device__ __forceinline__ float head_sum(float v) {
#pragma unroll
  for (int o = 1; o < HL; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// predicated 128-bit read-only load (zero when the predicate is off): keeps the per-key loops branch-free
__device__ __forceinline__ float4 ldg4_if(const float* p, bool on) {
  float4 r;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\tmov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t"
      "@p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
      : "l"(p), "r"((int)on));
  return r;
}
// 1 / max(sqrt(ss), 1e-12)  (F.normalize) on the MUFU path
// EXACT = fp32 parity mode (IEEE sqrt / divide / expf, the reference's own op sequence); otherwise the MUFU path
template <bool EXACT>
__device__ __forceinline__ float inv_norm(float ss) {
  return EXACT ? 1.f / fmaxf(sqrtf(ss), 1e-12f) : rsqrtf(fmaxf(ss, 1e-24f));
}

// The moving side (q rows; q, dO, o rows and lse in the backward) is staged with cp.async into LANE-PRIVATE shared
// memory (every lane reads back only the 16 bytes it copied, so no barrier is needed): all of a window's rows are in
// flight at once, and the second key chunk of the backward re-reads them from shared memory, not from L2.
// The per-window work is instantiated for padded key counts NK in {4, 8, 16} (the count is warp-uniform) so that
// every inner loop is branch-free with NK (x ROWS) independent dependency chains; padded keys are zero rows whose
// probability is forced to zero.
template <int HL, int NK, int ROWS, bool EXACT>
__device__ __forceinline__ void small_fwd_window(const AttnArgs& a, const float4 (*qs)[32], int nq, int nk, int tokv, int col, int lane,
                                                 float scale) {
  constexpr float LN2 = 0.6931471805599453f;
  float4 kr[NK], vr[NK];
#pragma unroll
  for (int j = 0; j < NK; ++j) {
    const int t = __shfl_sync(0xffffffffu, tokv, j);
    kr[j] = ldg4_if(a.k + (int64_t)t * a.ldk + col, j < nk);
    vr[j] = ldg4_if(a.v + (int64_t)t * a.ldv + col, j < nk);
  }
#pragma unroll
  for (int j = 0; j < NK; ++j) scale4(kr[j], inv_norm<EXACT>(head_sum<HL>(dot4(kr[j], kr[j]))));
  cp_async_wait_all();
  for (int i = 0; i < nq; i += ROWS) {
    float4 q[ROWS], acc[ROWS];
    float sc[ROWS][NK], m[ROWS], l[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      q[r] = qs[(i + r) & (SW_T - 1)][lane];
      scale4(q[r], scale * inv_norm<EXACT>(head_sum<HL>(dot4(q[r], q[r]))));
      m[r] = -INFINITY;
    }
#pragma unroll
    for (int j = 0; j < NK; ++j)
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const float s = head_sum<HL>(dot4(q[r], kr[j]));
        sc[r][j] = j < nk ? s : -INFINITY;
        m[r] = fmaxf(m[r], sc[r][j]);
      }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { l[r] = 0.f; acc[r] = make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
    for (int j = 0; j < NK; ++j)
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const float p = EXACT ? exp2f(sc[r][j] - m[r]) : fast_exp2(sc[r][j] - m[r]);
        l[r] += p;
        axpy4(acc[r], p, vr[j]);
      }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      if (i + r < nq) {
        const int row = __shfl_sync(0xffffffffu, tokv, SW_T + ((i + r) & (SW_T - 1)));
        scale4(acc[r], EXACT ? 1.f / l[r] : __fdividef(1.f, l[r]));
        *reinterpret_cast<float4*>(a.o + (int64_t)row * a.C + col) = acc[r];
        if (a.lse && (lane & (HL - 1)) == 0) a.lse[(int64_t)row * a.H + (col / (HL * 4))] = (m[r] + (EXACT ? log2f(l[r]) : __log2f(l[r]))) * LN2;
      }
    }
  }
}

template <int HD, bool EXACT>
__global__ void __launch_bounds__(SW_THREADS, 3) attn_small_fwd_kernel(AttnArgs a) {
  constexpr int HL = HD / 4;
  __shared__ float4 q_s[SW_THREADS / 32][SW_T][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int n_small = min(*a.small_end, *a.n_win);
  const int groups = a.C >> 7;
  const int n_items = n_small * groups;
  constexpr float LOG2E = 1.4426950408889634f;
  const float scale = LOG2E / fmaxf(*a.tau, a.tau_min);  // scores are kept in the log2 domain
  for (int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < n_items; item += warps) {
    const int w = item / groups, col = (item - w * groups) * 128 + lane * 4;
    const int nq = min(a.qcnt[w], SW_T), nk = min(a.kcnt[w], SW_T);
    const int tokv = lane < SW_T ? a.ktok[w * MAXT + lane] : a.qtok[w * MAXT + lane - SW_T];
#pragma unroll
    for (int i = 0; i < SW_T; ++i) {
      const int t = __shfl_sync(0xffffffffu, tokv, SW_T + i);
      if (i < nq) cp_async16(&q_s[wib][i][lane], a.q + (int64_t)t * a.ldq + col);
    }
    if (nk <= 4) small_fwd_window<HL, 4, 2, EXACT>(a, q_s[wib], nq, nk, tokv, col, lane, scale);
    else if (nk <= 8) small_fwd_window<HL, 8, 2, EXACT>(a, q_s[wib], nq, nk, tokv, col, lane, scale);
    else small_fwd_window<HL, 16, 1, EXACT>(a, q_s[wib], nq, nk, tokv, col, lane, scale);
  }
}

// Fused backward: dQ, dK, dV and dtau in one pass.  Keys are processed in register chunks of up to 8 (K_hat, V and the
// dK_hat, dV accumulators = 128 registers); windows with 9..16 keys run the query loop twice and add the second
// chunk's dQ contribution to the row written by the first (the normalisation Jacobian is linear in dQ_hat).
constexpr int SW_KC = 8;
struct SmallBwdSmem {
  float4 q[SW_T][32], g[SW_T][32], o[SW_T][32];
  float lse[SW_T][32];
};
template <int HL, int NC, bool EXACT>
__device__ __forceinline__ void small_bwd_chunk(const AttnArgs& a, const SmallBwdSmem& S, int nq, int kc, int nc, int tokv, int col, int lane,
                                                float inv_tau, float& dtau_acc, bool first) {
  const bool lead = (lane & (HL - 1)) == 0;
  float4 kr[NC], vr[NC], dk[NC], dv[NC];
  float kinv[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int t = __shfl_sync(0xffffffffu, tokv, (kc + j) & (SW_T - 1));
    kr[j] = ldg4_if(a.k + (int64_t)t * a.ldk + col, j < nc);
    vr[j] = ldg4_if(a.v + (int64_t)t * a.ldv + col, j < nc);
    dk[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    dv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    kinv[j] = inv_norm<EXACT>(head_sum<HL>(dot4(kr[j], kr[j])));
    scale4(kr[j], kinv[j]);
  }
  if (first) cp_async_wait_all();
  for (int i = 0; i < nq; ++i) {
    float4 qh = S.q[i][lane];
    const float4 g = S.g[i][lane], o = S.o[i][lane];
    const float L = S.lse[i][lane];
    const float qinv = inv_norm<EXACT>(head_sum<HL>(dot4(qh, qh)));
    scale4(qh, qinv);
    const float D = head_sum<HL>(dot4(g, o));
    float4 dqh = make_float4(0.f, 0.f, 0.f, 0.f);
    float s[NC], dp[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      s[j] = head_sum<HL>(dot4(qh, kr[j])) * inv_tau;
      dp[j] = head_sum<HL>(dot4(g, vr[j]));
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const float p = j < nc ? (EXACT ? expf(s[j] - L) : __expf(s[j] - L)) : 0.f;
      const float ds = p * (dp[j] - D);
      if (lead) dtau_acc = fmaf(-ds, s[j], dtau_acc);
      const float dsl = ds * inv_tau;
      axpy4(dqh, dsl, kr[j]);
      axpy4(dk[j], dsl, qh);
      axpy4(dv[j], p, g);
    }
    const float dt = head_sum<HL>(dot4(dqh, qh));
    float4 dq = make_float4((dqh.x - qh.x * dt) * qinv, (dqh.y - qh.y * dt) * qinv, (dqh.z - qh.z * dt) * qinv, (dqh.w - qh.w * dt) * qinv);
    const int row = __shfl_sync(0xffffffffu, tokv, SW_T + i);
    float4* dst = reinterpret_cast<float4*>(a.dq + (int64_t)row * a.ldq + col);
    if (!first) { const float4 prev = *dst; dq.x += prev.x; dq.y += prev.y; dq.z += prev.z; dq.w += prev.w; }
    *dst = dq;
  }
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int t = __shfl_sync(0xffffffffu, tokv, (kc + j) & (SW_T - 1));
    const float dt = head_sum<HL>(dot4(dk[j], kr[j]));
    if (j < nc) {
      const float4 r = make_float4((dk[j].x - kr[j].x * dt) * kinv[j], (dk[j].y - kr[j].y * dt) * kinv[j],
                                   (dk[j].z - kr[j].z * dt) * kinv[j], (dk[j].w - kr[j].w * dt) * kinv[j]);
      *reinterpret_cast<float4*>(a.dk + (int64_t)t * a.ldk + col) = r;
      *reinterpret_cast<float4*>(a.dv + (int64_t)t * a.ldv + col) = dv[j];
    }
  }
}

template <int HD, bool EXACT>
__global__ void __launch_bounds__(SW_THREADS, 2) attn_small_bwd_kernel(AttnArgs a) {
  constexpr int HL = HD / 4;
  extern __shared__ __align__(16) unsigned char sw_raw[];
  SmallBwdSmem& S = reinterpret_cast<SmallBwdSmem*>(sw_raw)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int n_small = min(*a.small_end, *a.n_win);
  const int groups = a.C >> 7;
  const int n_items = n_small * groups;
  const float tau_raw = *a.tau;
  const float inv_tau = 1.f / fmaxf(tau_raw, a.tau_min);
  float dtau_acc = 0.f;
  for (int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < n_items; item += warps) {
    const int w = item / groups, col = (item - w * groups) * 128 + lane * 4;
    const int head = col / HD;
    const int nq = min(a.qcnt[w], SW_T), nk = min(a.kcnt[w], SW_T);
    const int tokv = lane < SW_T ? a.ktok[w * MAXT + lane] : a.qtok[w * MAXT + lane - SW_T];
#pragma unroll
    for (int i = 0; i < SW_T; ++i) {
      const int t = __shfl_sync(0xffffffffu, tokv, SW_T + i);
      if (i < nq) {
        const int64_t off = (int64_t)t * a.C + col;
        cp_async16(&S.q[i][lane], a.q + (int64_t)t * a.ldq + col);
        cp_async16(&S.g[i][lane], a.dout + off);
        cp_async16(&S.o[i][lane], a.o + off);
        cp_async4(&S.lse[i][lane], a.lse + (int64_t)t * a.H + head);
      }
    }
    for (int kc = 0; kc < nk; kc += SW_KC) {
      const int nc = min(SW_KC, nk - kc);
      if (nc <= 4) small_bwd_chunk<HL, 4, EXACT>(a, S, nq, kc, nc, tokv, col, lane, inv_tau, dtau_acc, kc == 0);
      else small_bwd_chunk<HL, 8, EXACT>(a, S, nq, kc, nc, tokv, col, lane, inv_tau, dtau_acc, kc == 0);
    }
    if (nk == 0) cp_async_wait_all();  // nothing consumed the staged rows of this window
  }
  dtau_acc = warp_sum(dtau_acc);
  if (lane == 0 && a.dtau && tau_raw > a.tau_min && dtau_acc != 0.f) atomicAdd(a.dtau, dtau_acc * inv_tau);
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
      const int64_t srca = (int64_t)stat_tok[j] * (MODE == 2 ? a.ldq : a.ldk) + col0 + c4;
      const int64_t srcb = (int64_t)stat_tok[j] * (MODE == 2 ? a.C : a.ldv) + col0 + c4;
      *reinterpret_cast<float4*>(As + j * RS + sl * SS + d) = __ldg(reinterpret_cast<const float4*>(srcA + srca));
      *reinterpret_cast<float4*>(Bs + j * RS + sl * SS + d) = __ldg(reinterpret_cast<const float4*>(srcB + srcb));
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
  if (smem_attr_once((const void*)kern, (int)smem)) return TMAE_ERR_CUDA;
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
  const int* n_win; const int* begin; const int* mid; const int* end;
  const float* tau; float tau_min;
  int C, H;
  int ldq, ldk, ldv;
  const float* dout; float* dq; float* dk; float* dv; float* dtau;
};
int attn_mma_fwd(const AttnMmaArgs& a, int hd, int64_t max_windows, cudaStream_t s);
int attn_mma_bwd(const AttnMmaArgs& a, int hd, int64_t max_windows, cudaStream_t s);

static AttnMmaArgs to_mma(const AttnArgs& a) {
  AttnMmaArgs m{};
  m.q = a.q; m.k = a.k; m.v = a.v; m.o = a.o; m.lse = a.lse; m.qtok = a.qtok; m.qcnt = a.qcnt; m.ktok = a.ktok; m.kcnt = a.kcnt;
  m.n_win = a.n_win; m.begin = a.small_end; m.mid = a.mid_end; m.end = a.n_win; m.tau = a.tau; m.tau_min = a.tau_min; m.C = a.C; m.H = a.H; m.ldq = a.ldq; m.ldk = a.ldk; m.ldv = a.ldv;
  m.dout = a.dout; m.dq = a.dq; m.dk = a.dk; m.dv = a.dv; m.dtau = a.dtau;
  return m;
}

static void attn_bytes(const AttnArgs& a, double& fwd, double& bwd) {
  // algorithmic traffic: fwd reads q,k,v writes o ; bwd reads q,k,v,o,dO writes dq,dk,dv
  const double rq = a.rows_q * a.C * 4.0, rk = a.rows_kv * a.C * 4.0;
  fwd = 2 * rq + 2 * rk;
  bwd = 4 * rq + 4 * rk;
}

// Forward: small windows on the register-resident warp kernel; larger ones on mma.sync TF32 (tensor-core mode) or the
// fp32 shared-memory kernels (parity mode).  The profiler attributes the whole call's algorithmic bytes to the small
// kernel's scope when it is the only one (share of rows unknown on the host), so the scopes are kept separate and
// the mma scopes carry no byte count.
template <int HD>
static int launch_fwd(const AttnArgs& a, int64_t max_windows, cudaStream_t s) {
  double fb, bb;
  attn_bytes(a, fb, bb);
  {
    ProfScope prof("attn_small_fwd", 0, fb, s);
    int64_t items = max_windows * (a.C / 128);
    int64_t warps = items < (int64_t)kNumSMs * 32 ? items : (int64_t)kNumSMs * 32;
    if (a.tc) attn_small_fwd_kernel<HD, false><<<cdiv(warps * 32, SW_THREADS), SW_THREADS, 0, s>>>(a);
    else attn_small_fwd_kernel<HD, true><<<cdiv(warps * 32, SW_THREADS), SW_THREADS, 0, s>>>(a);
  }
  if (a.tc) return attn_mma_fwd(to_mma(a), HD, max_windows, s);
  ProfScope prof("attn_large_fwd", 0, 0, s);
  int r = launch_large<HD, 0, 32>(a, a.small_end, a.mid_end, max_windows, s);
  if (!r) r = launch_large<HD, 0, 64>(a, a.mid_end, a.n_win, max_windows, s);
  return r;
}

template <int HD>
static int launch_bwd(const AttnArgs& a, int64_t max_windows, cudaStream_t s) {
  double fb, bb;
  attn_bytes(a, fb, bb);
  {
    ProfScope prof("attn_small_bwd", 0, bb, s);
    int64_t items = max_windows * (a.C / 128);
    int64_t warps = items < (int64_t)kNumSMs * 24 ? items : (int64_t)kNumSMs * 24;
    constexpr int smem = (int)sizeof(SmallBwdSmem) * (SW_THREADS / 32);
    if (smem_attr_once((const void*)attn_small_bwd_kernel<HD, false>, smem) || smem_attr_once((const void*)attn_small_bwd_kernel<HD, true>, smem))
      return TMAE_ERR_CUDA;
    if (a.tc) attn_small_bwd_kernel<HD, false><<<cdiv(warps * 32, SW_THREADS), SW_THREADS, smem, s>>>(a);
    else attn_small_bwd_kernel<HD, true><<<cdiv(warps * 32, SW_THREADS), SW_THREADS, smem, s>>>(a);
  }
  if (a.tc) return attn_mma_bwd(to_mma(a), HD, max_windows, s);  // ONE fused pass (dQ, dK, dV, dtau)
  ProfScope prof("attn_large_bwd", 0, 0, s);
  int r = launch_large<HD, 1, 32>(a, a.small_end, a.mid_end, max_windows, s);
  if (!r) r = launch_large<HD, 1, 64>(a, a.mid_end, a.n_win, max_windows, s);
  if (!r) r = launch_large<HD, 2, 32>(a, a.small_end, a.mid_end, max_windows, s);
  if (!r) r = launch_large<HD, 2, 64>(a, a.mid_end, a.n_win, max_windows, s);
  return r;
}

static int check(const AttnArgs& a, int hd) {
  if (a.ldq < a.C || a.ldk < a.C || a.ldv < a.C || (a.ldq | a.ldk | a.ldv) % 4 != 0) return -1;
  if (a.C % TW != 0 || a.H != a.C / hd || (hd != 16 && hd != 32)) return -1;
  return 0;
}

}  // namespace tmae

using namespace tmae;

extern "C" {

int tmae_window_attention_fwd(const float* q, const float* k, const float* v, float* o, float* lse, const int32_t* qtok,
                              const int32_t* qcnt, const int32_t* ktok, const int32_t* kcnt, const int32_t* n_win,
                              const int32_t* small_end, const int32_t* mid_end, int64_t max_windows, const float* tau, float tau_min,
                              int32_t channels, int32_t heads, int32_t ld_q, int32_t ld_k, int32_t ld_v, int64_t rows_q, int64_t rows_kv,
                              int32_t precision, void* stream) {
  TMAE_CHECK_ARG(precision == TMAE_PREC_FP32 || precision == TMAE_PREC_TF32, "precision must be TMAE_PREC_FP32 or TMAE_PREC_TF32");
  AttnArgs a{};
  a.tc = precision != TMAE_PREC_FP32; a.rows_q = (double)rows_q; a.rows_kv = (double)rows_kv;
  a.ldq = ld_q; a.ldk = ld_k; a.ldv = ld_v;
  a.q = q; a.k = k; a.v = v; a.o = o; a.lse = lse; a.qtok = qtok; a.qcnt = qcnt; a.ktok = ktok; a.kcnt = kcnt; a.n_win = n_win;
  a.small_end = small_end; a.mid_end = mid_end; a.tau = tau; a.tau_min = tau_min; a.C = channels; a.H = heads;
  int hd = channels / heads;
  TMAE_CHECK_ARG(check(a, hd) == 0, "channels must be a multiple of 128 and head_dim 16 or 32");
  TMAE_CHECK_ARG(small_end && mid_end && n_win, "n_win / small_end / mid_end must be device pointers");
  if (max_windows <= 0) return 0;
  int r = hd == 16 ? launch_fwd<16>(a, max_windows, (cudaStream_t)stream) : launch_fwd<32>(a, max_windows, (cudaStream_t)stream);
  if (r) { set_error("tmae_window_attention_fwd: launch failed"); return r; }
  return 0;
}

int tmae_window_attention_bwd(const float* dout, const float* q, const float* k, const float* v, const float* o, const float* lse,
                              float* dsum, float* dq, float* dk, float* dv, float* dtau, const int32_t* qtok, const int32_t* qcnt,
                              const int32_t* ktok, const int32_t* kcnt, const int32_t* n_win, const int32_t* small_end,
                              const int32_t* mid_end, int64_t max_windows, const float* tau, float tau_min, int32_t channels,
                              int32_t heads, int32_t ld_q, int32_t ld_k, int32_t ld_v, int64_t rows_q, int64_t rows_kv, int32_t precision,
                              void* stream) {
  TMAE_CHECK_ARG(precision == TMAE_PREC_FP32 || precision == TMAE_PREC_TF32, "precision must be TMAE_PREC_FP32 or TMAE_PREC_TF32");
  AttnArgs a{};
  a.tc = precision != TMAE_PREC_FP32; a.rows_q = (double)rows_q; a.rows_kv = (double)rows_kv;
  a.ldq = ld_q; a.ldk = ld_k; a.ldv = ld_v;
  a.q = q; a.k = k; a.v = v; a.o = (float*)o; a.lse = (float*)lse; a.dsum = dsum; a.qtok = qtok; a.qcnt = qcnt; a.ktok = ktok;
  a.kcnt = kcnt; a.n_win = n_win; a.small_end = small_end; a.mid_end = mid_end; a.tau = tau; a.tau_min = tau_min; a.C = channels;
  a.H = heads; a.dout = dout; a.dq = dq; a.dk = dk; a.dv = dv; a.dtau = dtau;
  int hd = channels / heads;
  TMAE_CHECK_ARG(check(a, hd) == 0, "channels must be a multiple of 128 and head_dim 16 or 32");
  TMAE_CHECK_ARG(small_end && mid_end && n_win && dsum, "n_win / small_end / mid_end / dsum must be device pointers");
  if (max_windows <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  int r = hd == 16 ? launch_bwd<16>(a, max_windows, s) : launch_bwd<32>(a, max_windows, s);
  if (r) { set_error("tmae_window_attention_bwd: launch failed"); return r; }
  return 0;
}

}  // extern "C"
