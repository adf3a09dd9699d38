// Tensor-core attention for the windows with more than 16 tokens (stages 2-3 of a lidar scan are dense: these windows
// hold most of the query-key pairs).  One CTA per (window, head); warp w owns the 16-row strip w of the window: 2-warp
// CTAs for windows of <= 32 tokens, 4-warp CTAs for <= 64 (the partition sorts windows by level), and column tiles
// beyond the window's token count are skipped.  Everything a (window, head) needs -- Q_hat, K_hat, V (and dO) head
// slices, <= 64 x 32 fp32 each -- is gathered once through the partition's token table into shared memory by ALL
// threads with every load in flight before the first use (TF32-rounded, rows padded so that the
// mma.sync fragment loads are bank-conflict free), and S = Q_hat K_hat^T, P = softmax, O = P V run on
// mma.sync.m16n8k8 TF32 with fp32 accumulation; the softmax and the C-fragment -> A-fragment re-layout of P stay in
// registers (quad shuffles).  Backward is ONE pass: each warp first acts on its query strip (S, P, dP = dO V^T,
// dS -> dQ, dtau) and then on its key strip with the roles transposed (S^T, P^T, dP^T -> dV = P^T dO, dK = dS^T Q_hat),
// so no cross-warp reduction is needed.  tcgen05 is not used here on purpose: the MMAs are 16x8x8..64x64x32, far below
// a UMMA tile, and every product feeds a register-resident softmax.
// Reference: cosine_msa.py:114-176 (the per-window bmm / softmax / bmm) and its autograd backward.
#include "common.cuh"

namespace tmae {

constexpr int MT = TMAE_WIN_TOKENS;  // 64 slots

struct AttnMmaArgs {
  const float* q; const float* k; const float* v; float* o; float* lse;
  const int* qtok; const int* qcnt; const int* ktok; const int* kcnt;
  const int* n_win; const int* begin; const int* mid; const int* end;   // windows [*begin, *mid) <= 32 tokens, [*mid, *n_win) <= 64
  const float* tau; float tau_min;
  int C, H;
  int ldq, ldk, ldv;                    // row pitches of q / k / v (and dq / dk / dv); o and dO have pitch C
  const float* dout; float* dq; float* dk; float* dv; float* dtau;
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
// C fragment (rows g / g+8, cols 2t, 2t+1 of an 8-wide tile) -> A fragment (rows g / g+8, cols t, t+4) of the same tile
__device__ __forceinline__ void c_to_a(const float* c, uint32_t* a, int lane) {
  const int t = lane & 3;
  const int src0 = (lane & ~3) | (t >> 1), src1 = src0 + 2;
  const bool odd = t & 1;
  float v00 = __shfl_sync(0xffffffffu, c[0], src0), v01 = __shfl_sync(0xffffffffu, c[1], src0);
  float v10 = __shfl_sync(0xffffffffu, c[2], src0), v11 = __shfl_sync(0xffffffffu, c[3], src0);
  float w00 = __shfl_sync(0xffffffffu, c[0], src1), w01 = __shfl_sync(0xffffffffu, c[1], src1);
  float w10 = __shfl_sync(0xffffffffu, c[2], src1), w11 = __shfl_sync(0xffffffffu, c[3], src1);
  a[0] = to_tf32(odd ? v01 : v00);
  a[1] = to_tf32(odd ? v11 : v10);
  a[2] = to_tf32(odd ? w01 : w00);
  a[3] = to_tf32(odd ? w11 : w10);
}

// ---- staging: ALL threads of the CTA, LPR = HD/4 adjacent lanes per row (one 16-byte piece each, so a row's head
// slice is one coalesced 64/128-byte segment), every load of every operand issued before the first use.
template <int HD, int TCAP, int THREADS>
struct Stage {
  static constexpr int LPR = HD / 4, RPP = THREADS / LPR, PASSES = TCAP / RPP;
  float4 x[PASSES];
  // tok: global token list of the window (n valid entries)
  __device__ __forceinline__ void load(const float* __restrict__ src, const int* __restrict__ tok, int n, int C, int col0) {
    const int sub = threadIdx.x % LPR, r0 = threadIdx.x / LPR;
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      const int r = p * RPP + r0;
      x[p] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < n) x[p] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)__ldg(tok + r) * C + col0 + sub * 4));
    }
  }
  // per-row reduction over the LPR lanes that hold the row
  __device__ __forceinline__ static float row_sum(float v) {
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  // optional L2 normalisation (records 1 / max(|row|, eps)), TF32 rounding, store to dst[row][STRIDE]
  template <int STRIDE>
  __device__ __forceinline__ void store(float* dst, bool normalise, float* inv_out) {
    const int sub = threadIdx.x % LPR, r0 = threadIdx.x / LPR;
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      const int r = p * RPP + r0;
      float4 v = x[p];
      if (normalise) {
        const float ss = row_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
        const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
        v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
        if (inv_out && sub == 0) inv_out[r] = ss > 0.f ? inv : 0.f;
      }
      *reinterpret_cast<uint4*>(dst + r * STRIDE + sub * 4) = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    }
  }
};

// strip (16 rows starting at r0) of  X[rows][HD] (stride XS)  times  Y[cols][HD]^T (stride YS)  -> acc[NT col tiles][4];
// only the first `tiles` column tiles are computed (the rest of the window side is empty)
template <int HD, int XS, int YS, int NT>
__device__ __forceinline__ void strip_xyT(const float* X, const float* Y, int r0, int lane, int tiles, float acc[NT][4]) {
  const int g = lane >> 2, t = lane & 3;
  uint32_t a[HD / 8][4];
#pragma unroll
  for (int ks = 0; ks < HD / 8; ++ks) {
    a[ks][0] = __float_as_uint(X[(r0 + g) * XS + ks * 8 + t]);
    a[ks][1] = __float_as_uint(X[(r0 + g + 8) * XS + ks * 8 + t]);
    a[ks][2] = __float_as_uint(X[(r0 + g) * XS + ks * 8 + t + 4]);
    a[ks][3] = __float_as_uint(X[(r0 + g + 8) * XS + ks * 8 + t + 4]);
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    if (nt < tiles) {
#pragma unroll
      for (int ks = 0; ks < HD / 8; ++ks)
        mma_tf32(acc[nt], a[ks], __float_as_uint(Y[(nt * 8 + g) * YS + ks * 8 + t]), __float_as_uint(Y[(nt * 8 + g) * YS + ks * 8 + t + 4]));
    }
  }
}
// out[HD/8 tiles][4] (16 x HD strip) = P (16 x 8*NT, C-fragment layout per 8-wide tile) times Z[8*NT][HD] (stride ZS)
template <int HD, int ZS, int NT>
__device__ __forceinline__ void strip_pz(const float p[NT][4], const float* Z, int lane, int tiles, float out[HD / 8][4]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int dt = 0; dt < HD / 8; ++dt) out[dt][0] = out[dt][1] = out[dt][2] = out[dt][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < NT; ++kt) {
    if (kt < tiles) {
      uint32_t a[4];
      c_to_a(p[kt], a, lane);
#pragma unroll
      for (int dt = 0; dt < HD / 8; ++dt)
        mma_tf32(out[dt], a, __float_as_uint(Z[(kt * 8 + t) * ZS + dt * 8 + g]), __float_as_uint(Z[(kt * 8 + t + 4) * ZS + dt * 8 + g]));
    }
  }
}

// windows [*a.begin, *a.end) of the level-sorted list hold at most TCAP tokens per side; CTA = TCAP/16 warps
template <int HD, int TCAP, int OCC = 0>
__global__ void __launch_bounds__(TCAP * 2, OCC) attn_mma_fwd_kernel(AttnMmaArgs a) {
  constexpr int KS = HD + 4, VS = HD + 8, NT = TCAP / 8, THREADS = TCAP * 2;
  __shared__ __align__(16) float Qs[TCAP * KS], Ks[TCAP * KS], Vs[TCAP * VS];
  __shared__ int qt[TCAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nw = *a.n_win, first = min(*a.begin, nw), last = min(*a.end, nw);
  const int n_items = (last - first) * a.H;
  const float inv_tau = 1.f / fmaxf(*a.tau, a.tau_min);
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int w = first + item / a.H, h = item % a.H, col0 = h * HD;
    const int nq = min(a.qcnt[w], TCAP), nk = min(a.kcnt[w], TCAP);
    Stage<HD, TCAP, THREADS> sq, sk, sv;
    sq.load(a.q, a.qtok + w * MT, nq, a.ldq, col0);
    sk.load(a.k, a.ktok + w * MT, nk, a.ldk, col0);
    sv.load(a.v, a.ktok + w * MT, nk, a.ldv, col0);
    __syncthreads();  // the previous item's readers are done with shared memory
    if (threadIdx.x < TCAP) qt[threadIdx.x] = threadIdx.x < nq ? a.qtok[w * MT + threadIdx.x] : 0;
    sq.template store<KS>(Qs, true, nullptr);
    sk.template store<KS>(Ks, true, nullptr);
    sv.template store<VS>(Vs, false, nullptr);
    __syncthreads();
    const int r0 = warp * 16;
    const int ktiles = (nk + 7) >> 3;
    if (r0 < nq) {
      float s[NT][4];
      strip_xyT<HD, KS, KS, NT>(Qs, Ks, r0, lane, ktiles, s);
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int col = nt * 8 + 2 * t + (e & 1);
          s[nt][e] = col < nk ? s[nt][e] * inv_tau : -INFINITY;
        }
        m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
        m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
      }
      m0 = quad_max(m0); m1 = quad_max(m1);
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        s[nt][0] = __expf(s[nt][0] - m0); s[nt][1] = __expf(s[nt][1] - m0);
        s[nt][2] = __expf(s[nt][2] - m1); s[nt][3] = __expf(s[nt][3] - m1);
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
      }
      l0 = quad_sum(l0); l1 = quad_sum(l1);
      float o[HD / 8][4];
      strip_pz<HD, VS, NT>(s, Vs, lane, ktiles, o);
      const float i0 = 1.f / l0, i1 = 1.f / l1;
      const int ra = r0 + g, rb = r0 + g + 8;
#pragma unroll
      for (int dt = 0; dt < HD / 8; ++dt) {
        if (ra < nq) *reinterpret_cast<float2*>(a.o + (int64_t)qt[ra] * a.C + col0 + dt * 8 + 2 * t) = make_float2(o[dt][0] * i0, o[dt][1] * i0);
        if (rb < nq) *reinterpret_cast<float2*>(a.o + (int64_t)qt[rb] * a.C + col0 + dt * 8 + 2 * t) = make_float2(o[dt][2] * i1, o[dt][3] * i1);
      }
      if (a.lse && t == 0) {
        if (ra < nq) a.lse[(int64_t)qt[ra] * a.H + h] = m0 + __logf(l0);
        if (rb < nq) a.lse[(int64_t)qt[rb] * a.H + h] = m1 + __logf(l1);
      }
    }
  }
}

// OCC: CTAs per SM the register allocation targets.  The kernel is bound by fixed-latency dependency stalls and CTA
// barriers with 4 warps per scheduler (profiles/r01_ncu_attn_mma_v2.txt), so one more resident CTA per SM buys issue slots.
template <int HD, int TCAP, int OCC>
__global__ void __launch_bounds__(TCAP * 2, OCC) attn_mma_bwd_kernel(AttnMmaArgs a) {
  constexpr int KS = HD + 4, VS = HD + 8, NT = TCAP / 8, THREADS = TCAP * 2;
  __shared__ __align__(16) float Qs[TCAP * KS], Ks[TCAP * KS], Vs[TCAP * VS], Ds[TCAP * KS];  // Ds = dO
  __shared__ float qinv[TCAP], kinv[TCAP], lse_s[TCAP], dsum[TCAP];
  __shared__ int qt[TCAP], kt[TCAP];
  using St = Stage<HD, TCAP, THREADS>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nw = *a.n_win, first = min(*a.begin, nw), last = min(*a.end, nw);
  const int n_items = (last - first) * a.H;
  const float tau_raw = *a.tau;
  const float inv_tau = 1.f / fmaxf(tau_raw, a.tau_min);
  float dtau_acc = 0.f;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int w = first + item / a.H, h = item % a.H, col0 = h * HD;
    const int nq = min(a.qcnt[w], TCAP), nk = min(a.kcnt[w], TCAP);
    St sq, sk, sv, sg, so;
    sq.load(a.q, a.qtok + w * MT, nq, a.ldq, col0);
    sk.load(a.k, a.ktok + w * MT, nk, a.ldk, col0);
    sv.load(a.v, a.ktok + w * MT, nk, a.ldv, col0);
    sg.load(a.dout, a.qtok + w * MT, nq, a.C, col0);
    so.load(a.o, a.qtok + w * MT, nq, a.C, col0);
    __syncthreads();  // the previous item's readers are done with shared memory
    if (threadIdx.x < TCAP) {
      const int r = threadIdx.x;
      qt[r] = r < nq ? a.qtok[w * MT + r] : 0;
      kt[r] = r < nk ? a.ktok[w * MT + r] : 0;
      lse_s[r] = r < nq ? a.lse[(int64_t)qt[r] * a.H + h] : 0.f;
    }
    sq.template store<KS>(Qs, true, qinv);
    sk.template store<KS>(Ks, true, kinv);
    sv.template store<VS>(Vs, false, nullptr);
    {  // D = dO . O per query row, on the un-rounded fp32 values
      const int sub = threadIdx.x % St::LPR, rr = threadIdx.x / St::LPR;
#pragma unroll
      for (int p = 0; p < St::PASSES; ++p) {
        const float d = St::row_sum(sg.x[p].x * so.x[p].x + sg.x[p].y * so.x[p].y + sg.x[p].z * so.x[p].z + sg.x[p].w * so.x[p].w);
        if (sub == 0) dsum[p * St::RPP + rr] = d;
      }
    }
    sg.template store<KS>(Ds, false, nullptr);
    __syncthreads();
    const int r0 = warp * 16, ra = r0 + g, rb = r0 + g + 8;
    const int qtiles = (nq + 7) >> 3, ktiles = (nk + 7) >> 3;
    // ---------------- query strip: dQ (+ dtau)
    if (r0 < nq) {
      float s[NT][4], dp[NT][4];
      strip_xyT<HD, KS, KS, NT>(Qs, Ks, r0, lane, ktiles, s);    // S = Q_hat K_hat^T
      strip_xyT<HD, KS, VS, NT>(Ds, Vs, r0, lane, ktiles, dp);   // dP = dO V^T
      const float La = lse_s[ra], Lb = lse_s[rb], Da = dsum[ra], Db = dsum[rb];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = nt * 8 + 2 * t + (e & 1);
          const bool up = e < 2;
          const bool valid = col < nk && (up ? ra : rb) < nq;
          const float sc = s[nt][e] * inv_tau;
          const float p = valid ? __expf(sc - (up ? La : Lb)) : 0.f;
          const float ds = p * (dp[nt][e] - (up ? Da : Db));
          dtau_acc -= ds * sc;
          s[nt][e] = ds * inv_tau;  // dS / tau
        }
      }
      float dqh[HD / 8][4];
      strip_pz<HD, KS, NT>(s, Ks, lane, ktiles, dqh);  // dQ_hat = dS K_hat / tau
      // through q_hat = q / max(|q|, eps): dq = (dq_hat - q_hat (q_hat . dq_hat)) / |q|
      float da = 0.f, db = 0.f;
#pragma unroll
      for (int dt = 0; dt < HD / 8; ++dt) {
        const float* qa = Qs + ra * KS + dt * 8 + 2 * t;
        const float* qb = Qs + rb * KS + dt * 8 + 2 * t;
        da += dqh[dt][0] * qa[0] + dqh[dt][1] * qa[1];
        db += dqh[dt][2] * qb[0] + dqh[dt][3] * qb[1];
      }
      da = quad_sum(da); db = quad_sum(db);
#pragma unroll
      for (int dt = 0; dt < HD / 8; ++dt) {
        const float* qa = Qs + ra * KS + dt * 8 + 2 * t;
        const float* qb = Qs + rb * KS + dt * 8 + 2 * t;
        if (ra < nq)
          *reinterpret_cast<float2*>(a.dq + (int64_t)qt[ra] * a.ldq + col0 + dt * 8 + 2 * t) =
              make_float2((dqh[dt][0] - qa[0] * da) * qinv[ra], (dqh[dt][1] - qa[1] * da) * qinv[ra]);
        if (rb < nq)
          *reinterpret_cast<float2*>(a.dq + (int64_t)qt[rb] * a.ldq + col0 + dt * 8 + 2 * t) =
              make_float2((dqh[dt][2] - qb[0] * db) * qinv[rb], (dqh[dt][3] - qb[1] * db) * qinv[rb]);
      }
    }
    // ---------------- key strip (roles transposed): dV, dK
    if (r0 < nk) {
      float s[NT][4], dp[NT][4];
      strip_xyT<HD, KS, KS, NT>(Ks, Qs, r0, lane, qtiles, s);    // S^T = K_hat Q_hat^T      (rows = keys, cols = queries)
      strip_xyT<HD, VS, KS, NT>(Vs, Ds, r0, lane, qtiles, dp);   // dP^T = V dO^T
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = nt * 8 + 2 * t + (e & 1);   // query index
          const bool valid = col < nq && (e < 2 ? ra : rb) < nk;
          const float p = valid ? __expf(s[nt][e] * inv_tau - lse_s[col]) : 0.f;
          const float ds = p * (dp[nt][e] - dsum[col]);
          s[nt][e] = p;                // P^T
          dp[nt][e] = ds * inv_tau;    // dS^T / tau
        }
      }
      float dvv[HD / 8][4], dkh[HD / 8][4];
      strip_pz<HD, KS, NT>(s, Ds, lane, qtiles, dvv);    // dV = P^T dO
      strip_pz<HD, KS, NT>(dp, Qs, lane, qtiles, dkh);   // dK_hat = dS^T Q_hat / tau
      float da = 0.f, db = 0.f;
#pragma unroll
      for (int dt = 0; dt < HD / 8; ++dt) {
        const float* ka = Ks + ra * KS + dt * 8 + 2 * t;
        const float* kb = Ks + rb * KS + dt * 8 + 2 * t;
        da += dkh[dt][0] * ka[0] + dkh[dt][1] * ka[1];
        db += dkh[dt][2] * kb[0] + dkh[dt][3] * kb[1];
      }
      da = quad_sum(da); db = quad_sum(db);
#pragma unroll
      for (int dt = 0; dt < HD / 8; ++dt) {
        const float* ka = Ks + ra * KS + dt * 8 + 2 * t;
        const float* kb = Ks + rb * KS + dt * 8 + 2 * t;
        const int c = col0 + dt * 8 + 2 * t;
        if (ra < nk) {
          *reinterpret_cast<float2*>(a.dk + (int64_t)kt[ra] * a.ldk + c) = make_float2((dkh[dt][0] - ka[0] * da) * kinv[ra], (dkh[dt][1] - ka[1] * da) * kinv[ra]);
          *reinterpret_cast<float2*>(a.dv + (int64_t)kt[ra] * a.ldv + c) = make_float2(dvv[dt][0], dvv[dt][1]);
        }
        if (rb < nk) {
          *reinterpret_cast<float2*>(a.dk + (int64_t)kt[rb] * a.ldk + c) = make_float2((dkh[dt][2] - kb[0] * db) * kinv[rb], (dkh[dt][3] - kb[1] * db) * kinv[rb]);
          *reinterpret_cast<float2*>(a.dv + (int64_t)kt[rb] * a.ldv + c) = make_float2(dvv[dt][2], dvv[dt][3]);
        }
      }
    }
  }
  dtau_acc = warp_sum(dtau_acc);
  if (lane == 0 && a.dtau && tau_raw > a.tau_min && dtau_acc != 0.f) atomicAdd(a.dtau, dtau_acc * inv_tau);
}

int g_attn_occ_fwd = 1;   // 0: forward kernels with the compiler's default register allocation (A/B runs)
int g_attn_occ = 1;   // 1: the <= 32-token backward kernels are compiled for 10 CTAs per SM instead of 8 (measured 245 -> 220 us on a
                      // stage-2-like window mix; the 64-token class lost 8 % at 5 instead of 4 and stays at 4); 0 for A/B runs

template <typename K>
static void max_carveout(K kern) {   // static shared memory only: ask for the largest shared-memory carveout once per kernel
  // idempotent and cheap next to a launch; a per-process "done" flag would skip every device after the first
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

template <int HD, int TCAP>
static void launch_mma(bool bwd, const AttnMmaArgs& a, int64_t max_windows, cudaStream_t s) {
  int64_t items = max_windows * a.H;
  constexpr int OCC0 = TCAP == 64 ? 4 : 8, OCC1 = TCAP == 64 ? 4 : 10;
  const int per_sm = bwd ? (g_attn_occ ? OCC1 : (TCAP == 32 ? 10 : 4)) : (TCAP == 32 ? 14 : 7);
  int grid = (int)(items < (int64_t)per_sm * kNumSMs ? items : (int64_t)per_sm * kNumSMs);
  if (bwd && g_attn_occ) {
    max_carveout(attn_mma_bwd_kernel<HD, TCAP, OCC1>);
    attn_mma_bwd_kernel<HD, TCAP, OCC1><<<grid, TCAP * 2, 0, s>>>(a);
  } else if (bwd) {
    attn_mma_bwd_kernel<HD, TCAP, OCC0><<<grid, TCAP * 2, 0, s>>>(a);
  } else if (g_attn_occ_fwd) {
    // 72 registers: 7 resident 4-warp CTAs (64-token class) / 14 resident 2-warp CTAs (32-token class) per SM instead of 5 / 10
    // (measured: 88.4 -> 74.4 us on a stage-3-like and 107.4 -> 92.3 us on a stage-2-like window mix)
    constexpr int O = TCAP == 64 ? 7 : 14;
    max_carveout(attn_mma_fwd_kernel<HD, TCAP, O>);
    attn_mma_fwd_kernel<HD, TCAP, O><<<grid, TCAP * 2, 0, s>>>(a);
  } else {
    attn_mma_fwd_kernel<HD, TCAP><<<grid, TCAP * 2, 0, s>>>(a);
  }
}

static int attn_mma_run(bool bwd, AttnMmaArgs a, int hd, int64_t max_windows, cudaStream_t s) {
  ProfScope prof(bwd ? "attn_mma_bwd" : "attn_mma_fwd", 0, 0, s);
  // windows [small_end, mid_end): <= 32 tokens per side, 2-warp CTAs ; [mid_end, n_win): <= 64 tokens, 4-warp CTAs
  AttnMmaArgs mid = a, big = a;
  mid.end = a.mid;
  big.begin = a.mid; big.end = a.n_win;
  if (hd == 16) { launch_mma<16, 32>(bwd, mid, max_windows, s); launch_mma<16, 64>(bwd, big, max_windows, s); }
  else { launch_mma<32, 32>(bwd, mid, max_windows, s); launch_mma<32, 64>(bwd, big, max_windows, s); }
  return cudaGetLastError() == cudaSuccess ? 0 : TMAE_ERR_CUDA;
}

int attn_mma_fwd(const AttnMmaArgs& a, int hd, int64_t max_windows, cudaStream_t s) { return attn_mma_run(false, a, hd, max_windows, s); }
int attn_mma_bwd(const AttnMmaArgs& a, int hd, int64_t max_windows, cudaStream_t s) { return attn_mma_run(true, a, hd, max_windows, s); }

}  // namespace tmae
