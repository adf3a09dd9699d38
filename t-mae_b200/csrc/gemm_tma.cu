// TMA-fed tcgen05 GEMM for the dense (non-gathered) contractions of the path: operands stay fp32 in HBM and are
// consumed as TF32 (kind::tf32, fp32 accumulate in TMEM), so no thread ever touches an operand byte:
//   warp 0      TMA producer   cp.async.bulk.tensor.2d (SWIZZLE_128B boxes) -> STAGES-deep shared ring, mbarrier tx
//   warp 1      MMA issuer     tcgen05.mma.kind::tf32, M=128, N=BN, K=8 per instruction; tcgen05.commit frees stages
//   warps 2..9  epilogue       tcgen05.ld (lane = row; two warps per TMEM lane quarter take alternate 32-column chunks)
//                              -> bias / GELU / pre-activation copy -> 256-bit stores of the lane's 128-byte row
//                              segment (vector atomics for C += and the split reduction of dW)
// Modes: NT  C[m,n] = sum_k A[m,k] W[n,k]   (A, B K-major)            linear forward
//        NN  C[m,n] = sum_k A[m,k] B[k,n]   (A K-major, B MN-major)   linear backward-data
//        TN  C[m,n] = sum_k A[k,m] B[k,n]   (A, B MN-major)           linear backward-weight, split over k
// Shapes TMA cannot express (row pitch not a multiple of 16 bytes: only the k = 10 first VFE layer on this path) run on
// the fp32 kernels of gemm.cu and are counted (tmae_dispatch_counts).  Every mbarrier wait is bounded and traps
// instead of hanging.
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace tmae {

constexpr int KB = 32;         // fp32 elements per k-block = one 128-byte swizzle span
constexpr int EPI_WARPS = 8;                      // two warps per TMEM lane quarter, alternating 32-column chunks
constexpr int TMA_THREADS = 64 + 32 * EPI_WARPS;  // producer + MMA issuer + epilogue warps
constexpr int GW0 = 2 + EPI_WARPS;                // first gather warp of the GATHER variants

enum TmaMode { T_NT = 0, T_NN = 1, T_TN = 2 };

// Optional timeline of CTA 0 (diagnostics, tmae_debug_set_trace): records (event id << 56 | index << 40 | clock) words.
unsigned long long* g_trace_buf = nullptr;
bool g_wide_st = true;   // tmae_set_option("wide_st", 0): 128-bit epilogue stores (A/B measurement)
long long g_trace_cap = 0;
__device__ __forceinline__ void trace_ev(unsigned long long* buf, int& n, int cap, int ev, int idx) {
  if (buf && n < cap) buf[n++] = ((unsigned long long)ev << 56) | ((unsigned long long)(idx & 0xffff) << 40) | (clock64() & 0xffffffffffull);
}

struct TmaArgs {
  int64_t M, N, K;
  const float* bias;
  int64_t k_split;                   // NT dual source: k-blocks at k >= k_split come from the second pair of tensor maps
  const float* gsrc; const int* gtab; int gtaps, gcin;   // GATHER: A rows come through a neighbour table (sparse conv)
  float* C; float* P; int64_t ldc;   // output, optional pre-activation copy
  const float* gelu_pre;             // optional: multiply the result by gelu'(gelu_pre[m,n])  (fused GELU backward)
  int act, reduce_add, has_preact;
  int wide_st;                       // 1: rows of C (and P) are 32-byte aligned and N % 8 == 0 -> 256-bit epilogue stores
  int64_t k_chunk;
  unsigned long long* trace; int trace_cap;
};

// PERSISTENT kernel: one CTA per SM walks the tile list (n fastest, so CTAs running together share A through L2); the
// smem ring runs across tile boundaries and the accumulator is double-buffered in TMEM (2 x BN columns), so the loads
// and MMAs of tile i+1 overlap the epilogue (TMEM -> registers -> swizzled smem -> TMA store) of tile i.
// One stage of the ring: A then B.
//   K-major operand (rows x 32 fp32): ONE box {32, rows}; 8-row groups 1024 B apart (SBO), k-step = +32 B.
//   MN-major operand (32 k-rows x cols): cols/32 boxes {32, 32} of 4096 B each (LBO between boxes), 4-row K atoms 512 B
//   apart (SBO), k-step (8 rows) = +1024 B.
// 16-byte cp.async with zero fill (src_bytes = 0 writes zeros): the gathered A operand of the sparse convolution
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the executing thread's prior cp.async operations arrive on the mbarrier when they complete (no pending-count change)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
constexpr int GATHER_WARPS = 2;   // extra producer warps of the GATHER variant (warps 6, 7)

template <int MODE, int BN, int STAGES, bool GATHER = false>
__global__ void __launch_bounds__(TMA_THREADS + (GATHER ? GATHER_WARPS * 32 : 0), 1) tma_gemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                  const __grid_constant__ CUtensorMap map_b,
                                                                  const __grid_constant__ CUtensorMap map_a2,
                                                                  const __grid_constant__ CUtensorMap map_b2, TmaArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int A_BYTES = UM * KB * 4, B_BYTES = BN * KB * 4, STAGE = A_BYTES + B_BYTES;
  __shared__ uint64_t bar_full[STAGES], bar_empty[STAGES], bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (int)((g.M + UM - 1) / UM), n_tiles = (int)((g.N + BN - 1) / BN);
  const int z_tiles = (int)((g.K + g.k_chunk - 1) / g.k_chunk);
  const int total = m_tiles * n_tiles * z_tiles;

  if (threadIdx.x == 0) {
    // GATHER: the stage is full when the TMA bytes of B have landed (1 arrival + tx) and every gather thread's cp.asyncs of A have
    for (int s = 0; s < STAGES; ++s) { bar_init(&bar_full[s], GATHER ? 1 + GATHER_WARPS * 32 : 1); bar_init(&bar_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { bar_init(&bar_acc_full[b], 1); bar_init(&bar_acc_empty[b], EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  auto decode = [&](int t, int& m0, int& n0, int64_t& kbeg, int& nkb) {
    int nt = t % n_tiles, rest = t / n_tiles;
    int mt = rest % m_tiles, z = rest / m_tiles;
    m0 = mt * UM; n0 = nt * BN;
    kbeg = (int64_t)z * g.k_chunk;
    int64_t kend = kbeg + g.k_chunk < g.K ? kbeg + g.k_chunk : g.K;
    nkb = (int)((kend - kbeg + KB - 1) / KB);
  };

  if (warp == 0) {
    // ---------------- TMA producer
    if (lane == 0) {
      int it = 0;
      unsigned long long* tb = blockIdx.x == 0 ? g.trace : nullptr;
      int tn = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int m0, n0, nkb; int64_t kbeg;
        decode(t, m0, n0, kbeg, nkb);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES, use = it / STAGES;
          if (use > 0) bar_wait(&bar_empty[s], (use - 1) & 1);
          trace_ev(tb, tn, g.trace_cap / 4, 1, it);
          uint8_t* a = smem + s * STAGE;
          uint8_t* b = a + A_BYTES;
          const int k0 = (int)(kbeg + (int64_t)kb * KB);
          constexpr bool GA = GATHER && MODE == T_NT, GB = GATHER && MODE == T_TN;  // which operand the gather warps write
          bar_expect_tx(&bar_full[s], GA ? B_BYTES : (GB ? A_BYTES : STAGE));
          if (GA) {
            // A comes from the gather warps
          } else if (MODE == T_TN) {
#pragma unroll
            for (int j = 0; j < UM / 32; ++j) tma_load_2d(a + j * 4096, &map_a, m0 + j * 32, k0, &bar_full[s]);
          } else if (MODE == T_NT && k0 >= g.k_split) {   // second source pair: C = A1 B1^T + A2 B2^T
            tma_load_2d(a, &map_a2, k0 - (int)g.k_split, m0, &bar_full[s]);
          } else {
            tma_load_2d(a, &map_a, k0, m0, &bar_full[s]);
          }
          if (MODE == T_NT && k0 >= g.k_split) {
            tma_load_2d(b, &map_b2, k0 - (int)g.k_split, n0, &bar_full[s]);
          } else if (MODE == T_NT) {
            tma_load_2d(b, &map_b, k0, n0, &bar_full[s]);
          } else if (GB) {
            // B comes from the gather warps
          } else {
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) tma_load_2d(b + j * 4096, &map_b, n0 + j * 32, k0, &bar_full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(MODE == T_TN, MODE != T_NT, BN);
      int it = 0, i = 0;
      unsigned long long* tb = blockIdx.x == 0 && g.trace ? g.trace + g.trace_cap / 4 : nullptr;
      int tn = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++i) {
        int m0, n0, nkb; int64_t kbeg;
        decode(t, m0, n0, kbeg, nkb);
        const int buf = i & 1, round = i >> 1;
        if (round > 0) bar_wait(&bar_acc_empty[buf], (round - 1) & 1);  // epilogue drained this accumulator
        trace_ev(tb, tn, g.trace_cap / 4, 2, i);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES, use = it / STAGES;
          bar_wait(&bar_full[s], use & 1);
          trace_ev(tb, tn, g.trace_cap / 4, 3, it);
          if (GATHER) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // cp.async (generic proxy) writes -> UMMA (async proxy) reads
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = s_u32(smem + s * STAGE), b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int kk = 0; kk < KB / 8; ++kk) {
            uint64_t da = MODE == T_TN ? desc_sw128_32b(a_addr + kk * 1024, 4096, 512) : desc_sw128(a_addr + kk * 32, 16, 1024);
            uint64_t db = MODE == T_NT ? desc_sw128(b_addr + kk * 32, 16, 1024) : desc_sw128_32b(b_addr + kk * 1024, 4096, 512);
            umma_tf32(tmem_d, da, db, idesc, (kb | kk) ? 1u : 0u);
          }
          commit_to(&bar_empty[s]);
        }
        commit_to(&bar_acc_full[buf]);
      }
    }
  } else if (GATHER && MODE == T_TN && warp >= GW0) {
    // ---------------- gather producers, weight gradient of the sparse convolution:
    //   dW[co][tap*cin + c] = sum_r dy[r][co] * src[tab[r][tap]][c]        (TN: A = dy by TMA, B gathered, reduction over r)
    // One k-block = 32 reduction rows x BN columns of B in the MN-major SWIZZLE_128B_ATOM_32B layout (boxes of 32 columns,
    // 4096 B each; row k at +128 k; 32-byte chunk j of the row at (j ^ (k % 4)) * 32).  A thread owns one 16-byte piece
    // column of all 32 reduction rows of the k-block.
    static_assert(!(GATHER && MODE == T_TN) || BN == 4 * GATHER_WARPS * 32, "one 16-byte piece column per gather thread");
    const int gt = (warp - GW0) * 32 + lane;          // 0 .. 63 = 16-byte piece of the BN-column row: a warp instruction
                                                      // copies 512 contiguous bytes of ONE source row (4 lines, not 32)
    const uint32_t poff = (uint32_t)((gt >> 3) * 4096 + (gt & 1) * 16);   // box of 32 columns, 16-byte half of the 32-byte chunk
    const int jj = (gt >> 1) & 3;                     // 32-byte chunk inside the box row, XORed with (k % 4) below
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      int m0, n0, nkb; int64_t kbeg;
      decode(t, m0, n0, kbeg, nkb);
      const int64_t kend = kbeg + g.k_chunk < g.K ? kbeg + g.k_chunk : g.K;
      const int col = n0 + gt * 4;                    // first B column (= tap * cin + c) of this thread's piece
      const bool col_ok = col < g.N;
      // cin % 128 == 0 and a warp spans 128 columns: the 32 lanes of a warp share one tap, so lane l looks up the source row
      // of reduction row l (one k-block ahead) and the row loop broadcasts it with a shuffle
      const int tap = col_ok ? col / g.gcin : 0, c0 = col_ok ? col - tap * g.gcin : 0;
      int64_t r = kbeg + lane;
      int srow_l = (col_ok && r < kend) ? __ldg(g.gtab + r * g.gtaps + tap) : -1;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES, use = it / STAGES;
        const int64_t rn = r + KB;
        const int srow_n = (col_ok && kb + 1 < nkb && rn < kend) ? __ldg(g.gtab + rn * g.gtaps + tap) : -1;
        if (use > 0) bar_wait(&bar_empty[s], (use - 1) & 1);
        const uint32_t bbase = s_u32(smem + s * STAGE + A_BYTES) + poff;
        const float* p = g.gsrc + c0;
#pragma unroll 8
        for (int kr = 0; kr < KB; ++kr) {
          const int srow = __shfl_sync(0xffffffffu, srow_l, kr);
          cp_async16_zfill(bbase + (uint32_t)(kr * 128 + ((jj ^ (kr & 3)) * 32)), p + (int64_t)(srow < 0 ? 0 : srow) * g.gcin, srow < 0 ? 0u : 16u);
        }
        cp_async_arrive_noinc(&bar_full[s]);
        r = rn; srow_l = srow_n;
      }
    }
  } else if (GATHER && warp >= GW0) {
    // ---------------- gather producers (sparse convolution): A[m][tap*cin + c] = src[tab[m][tap]][c], absent neighbours
    // are zero-filled.  One k-block (32 fp32 = 128 bytes of one tap) of the 128 tile rows = 1024 16-byte pieces written
    // with cp.async straight into the SWIZZLE_128B K-major layout the UMMA descriptor expects (row r at (r/8)*1024 +
    // (r%8)*128, 16-byte chunk c at ((c ^ (r%8))*16).  Eight lanes copy one row's 128 bytes, so a warp instruction reads 4 rows x
    // 128 contiguous bytes (lane = row would touch 32 lines per instruction); a thread owns one chunk of 16 rows and looks
    // their source rows up one k-block ahead.
    const int gt = (warp - GW0) * 32 + lane;          // 0 .. 63
    const int rsub = gt >> 3, c16 = gt & 7;           // row inside an 8-row group, 16-byte chunk of the 128-byte k-block row
    constexpr int RG = UM / 8;                        // 16 row groups: the thread copies chunk c16 of rows rsub + 8 j
    const uint32_t doff = (uint32_t)(rsub * 128 + ((c16 ^ rsub) * 16));
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      int m0, n0, nkb; int64_t kbeg;
      decode(t, m0, n0, kbeg, nkb);
      int tap = (int)(kbeg / g.gcin);
      int src[RG], src_n[RG];
#pragma unroll
      for (int j = 0; j < RG; ++j) {
        const int64_t row = m0 + j * 8 + rsub;
        src[j] = row < g.M ? __ldg(g.gtab + row * g.gtaps + tap) : -1;
      }
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES, use = it / STAGES;
        const int k0 = (int)(kbeg + (int64_t)kb * KB);
        const int c0 = k0 - tap * g.gcin;
        // look ahead: source rows of the next k-block (same tap unless the k-block crosses into the next one)
        const bool cross = kb + 1 < nkb && c0 + KB >= g.gcin;
        if (cross) {
#pragma unroll
          for (int j = 0; j < RG; ++j) {
            const int64_t row = m0 + j * 8 + rsub;
            src_n[j] = row < g.M ? __ldg(g.gtab + row * g.gtaps + tap + 1) : -1;
          }
        }
        if (use > 0) bar_wait(&bar_empty[s], (use - 1) & 1);
        const uint32_t dst = s_u32(smem + s * STAGE) + doff;
        const float* base = g.gsrc + c0 + c16 * 4;
#pragma unroll
        for (int j = 0; j < RG; ++j)   // one instruction = 4 rows x 128 contiguous bytes (8 lanes per row)
          cp_async16_zfill(dst + j * 1024, base + (int64_t)(src[j] < 0 ? 0 : src[j]) * g.gcin, src[j] < 0 ? 0u : 16u);
        cp_async_arrive_noinc(&bar_full[s]);
        if (cross) {
          ++tap;
#pragma unroll
          for (int j = 0; j < RG; ++j) src[j] = src_n[j];
        }
      }
    }
  } else {
    // ---------------- epilogue warps: TMEM lane quarter q = warp % 4.  tcgen05.ld gives every lane one output row
    // (32 consecutive fp32 = one 128-byte line per chunk), which it writes straight to HBM with 128-bit stores: the
    // L2 merges the 16-byte pieces of a line, and there is no shared-memory staging, proxy fence or bulk-store wait on
    // the critical path (measured: the TMA-store variant spent ~2 us per 32x32 chunk waiting on them).
    const int q = warp & 3;
    constexpr int CHUNKS = BN / 32, CSTEP = EPI_WARPS / 4;
    const int ch0 = (warp - 2) >> 2;   // this warp's first chunk; it takes every CSTEP-th
    int i = 0;
    unsigned long long* tb = blockIdx.x == 0 && g.trace && warp == 2 && lane == 0 ? g.trace + 2 * (g.trace_cap / 4) : nullptr;
    int tn = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++i) {
      int m0, n0, nkb; int64_t kbeg;
      decode(t, m0, n0, kbeg, nkb);
      const int buf = i & 1, round = i >> 1;
      const int64_t row = m0 + q * 32 + lane;
      // fused GELU backward: the pre-activation row segment of chunk c+1 is loaded while chunk c is processed (and the
      // first one while this warp still waits for the accumulator) -- a load issued next to its use costs a DRAM round
      // trip per 32-column chunk in the four warps that pace the whole kernel
      float4 hn[8];
      auto load_pre = [&](int ch) {
        const int col0 = n0 + ch * 32;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          hn[c] = (row < g.M && col0 + 4 * c < g.N) ? __ldg(reinterpret_cast<const float4*>(g.gelu_pre + row * g.ldc + col0 + 4 * c))
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      if (g.gelu_pre) load_pre(ch0);
      trace_ev(tb, tn, g.trace_cap / 4, 4, i);
      bar_wait(&bar_acc_full[buf], round & 1);
      trace_ev(tb, tn, g.trace_cap / 4, 5, i);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_base + buf * BN;
      for (int ch = ch0; ch < CHUNKS; ch += CSTEP) {
        const int col0 = n0 + ch * 32;
        const float bl = (g.bias && col0 + lane < g.N) ? __ldg(g.bias + col0 + lane) : 0.f;  // in flight during the TMEM load
        uint32_t r[32];
        ld_tmem32(tmem_d + ((uint32_t)(q * 32) << 16) + ch * 32, r);
        if (ch + CSTEP >= CHUNKS) {  // this warp has read everything it needs from the accumulator
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) bar_arrive(&bar_acc_empty[buf]);
        }
        if (col0 >= g.N) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bl, j);
        if (row >= g.M) continue;

        float* crow = g.C + row * g.ldc + col0;
        if (g.has_preact) {
          float* prow = g.P + row * g.ldc + col0;
          if (g.wide_st) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (col0 + 8 * c < g.N) st_global_v8(prow + 8 * c, v + 8 * c);
          } else {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (col0 + 4 * c < g.N) *reinterpret_cast<float4*>(prow + 4 * c) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          }
        }
        if (g.gelu_pre) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 h = hn[c];
            v[4 * c] *= gelu_grad_t(h.x); v[4 * c + 1] *= gelu_grad_t(h.y); v[4 * c + 2] *= gelu_grad_t(h.z); v[4 * c + 3] *= gelu_grad_t(h.w);
          }
          if (ch + CSTEP < CHUNKS) load_pre(ch + CSTEP);
        }
        if (g.act == TMAE_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf_t(v[j]);
        } else if (g.act == TMAE_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (g.reduce_add) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (col0 + 4 * c < g.N) atomicAdd(reinterpret_cast<float4*>(crow + 4 * c), make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
        } else if (g.wide_st) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (col0 + 8 * c < g.N) st_global_v8(crow + 8 * c, v + 8 * c);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (col0 + 4 * c < g.N) *reinterpret_cast<float4*>(crow + 4 * c) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        }
      }
      trace_ev(tb, tn, g.trace_cap / 4, 6, i);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
}

// ------------------------------------------------------------------ host side
// cuTensorMapEncodeTiled is fetched through the runtime (cudaGetDriverEntryPoint) so that the library has no link-time
// dependency on libcuda.so.1 and still loads on a machine without a driver (build / ABI checks on CPU boxes).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static bool make_map(CUtensorMap* m, const float* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_inner,
                     uint32_t box_outer, bool tf32, bool mn_major = false) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  CUresult r = enc(m, tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims,
                                      strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

struct DualSrc { const float* A2; const float* B2; int64_t K2; };  // optional second source pair of an NT GEMM (K-major, pitch K2)

template <int MODE, int BN, bool GATHER = false>
static int tma_launch(const float* A, const float* B, float* C, float* preact, int64_t lda, int64_t ldb, int64_t ldc, TmaArgs g, int splits,
                      cudaStream_t s, DualSrc d2 = DualSrc{nullptr, nullptr, 0}) {
  constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  if (g.M <= 0 || g.N <= 0) return 0;
  if (splits < 1) splits = 1;
  const int64_t k_first = g.K;
  if (MODE == T_NT && d2.A2) g.K += d2.K2;  // the reduction runs over K (first pair) then K2 (second pair)
  g.k_chunk = align_up((g.K + splits - 1) / splits, KB);
  int z = (int)((g.K + g.k_chunk - 1) / g.k_chunk);
  if (z < 1) z = 1;
  if (z > 1) g.reduce_add = 1;
  CUtensorMap ma, mb;
  bool ok = true;
  // A: NT/NN K-major (rows = M, inner = K)  box {32, 128} ; TN MN-major (rows = K, inner = M) box {32, 32}
  ok &= MODE == T_TN ? make_map(&ma, A, g.M, g.K, lda, 32, 32, true, true) : make_map(&ma, A, k_first, g.M, lda, KB, UM, true);
  // B: NT K-major (rows = N, inner = K) box {32, BN} ; NN/TN MN-major (rows = K, inner = N) box {32, 32}
  ok &= MODE == T_NT ? make_map(&mb, B, k_first, g.N, ldb, KB, BN, true) : make_map(&mb, B, g.N, g.K, ldb, 32, 32, true, true);
  CUtensorMap ma2 = ma, mb2 = mb;
  g.k_split = g.K;
  if (MODE == T_NT && d2.A2) {  // both reduction lengths are multiples of the 32-wide k-block
    ok &= make_map(&ma2, d2.A2, d2.K2, g.M, d2.K2, KB, UM, true);
    ok &= make_map(&mb2, d2.B2, d2.K2, g.N, d2.K2, KB, BN, true);
    g.k_split = k_first;
  }
  if (!ok) return TMAE_ERR_CUDA;
  g.has_preact = preact != nullptr;
  g.trace = g_trace_buf; g.trace_cap = (int)g_trace_cap;
  g.C = C; g.P = preact; g.ldc = ldc;
  g.wide_st = g_wide_st && ldc % 8 == 0 && g.N % 8 == 0 && ((uintptr_t)C & 31) == 0 && ((uintptr_t)preact & 31) == 0;
  size_t smem = (size_t)STAGES * (UM * KB * 4 + BN * KB * 4) + 1024;
  auto kern = tma_gemm_kernel<MODE, BN, STAGES, GATHER>;
  if (smem_attr_once((const void*)kern, (int)smem)) return TMAE_ERR_CUDA;
  static const char* names[5] = {"tma_gemm_nt", "tma_gemm_nn", "tma_gemm_tn", "tma_gemm_nt_gather", "tma_gemm_tn_gather"};
  double c_el = (double)g.M * g.N * (1.0 + (g.reduce_add && z == 1 ? 1.0 : 0.0) + (preact ? 1.0 : 0.0));
  const double a_el = (GATHER && MODE == T_NT) ? (double)g.M * g.gcin : (double)g.M * g.K;
  const double b_el = (GATHER && MODE == T_TN) ? (double)g.K * g.gcin : (double)g.N * g.K;
  ProfScope prof(names[GATHER ? (MODE == T_TN ? 4 : 3) : MODE], 2.0 * g.M * g.N * g.K, 4.0 * (a_el + b_el + c_el), s);
  int64_t tiles = (int64_t)cdiv(g.N, BN) * cdiv(g.M, UM) * z;
  dim3 grid((unsigned)(tiles < kNumSMs ? tiles : kNumSMs));
  kern<<<grid, TMA_THREADS + (GATHER ? GATHER_WARPS * 32 : 0), smem, s>>>(ma, mb, ma2, mb2, g);
  return cudaGetLastError() == cudaSuccess ? 0 : TMAE_ERR_CUDA;
}

template <int MODE>
static int tma_dispatch(const float* A, const float* B, float* C, float* preact, int64_t lda, int64_t ldb, int64_t ldc, const TmaArgs& g,
                        int splits, cudaStream_t s, DualSrc d2 = DualSrc{nullptr, nullptr, 0}) {
  // N = 384 (packed q/k/v projection at 128 channels): three full 128-wide tiles instead of a full and a half-empty 256
  if (g.N > 128 && !(g.N % 256 != 0 && g.N % 128 == 0 && g.N <= 384))
    return tma_launch<MODE, 256>(A, B, C, preact, lda, ldb, ldc, g, splits, s, d2);
  if (g.N > 64) return tma_launch<MODE, 128>(A, B, C, preact, lda, ldb, ldc, g, splits, s, d2);
  return tma_launch<MODE, 64>(A, B, C, preact, lda, ldb, ldc, g, splits, s, d2);
}

// ---- entry points (gemm.cu dispatches here for TMAE_PREC_BF16 when the shapes allow TMA)
bool tma_linear_fwd_ok(const float* x, const float* w, const float* y, const float* residual, int64_t m, int64_t n, int64_t k) {
  return residual == nullptr && k % 4 == 0 && n % 4 == 0 && aligned16(x) && aligned16(w) && aligned16(y);
}
int tma_linear_fwd(const float* x, const float* w, const float* bias, float* y, float* preact, int64_t m, int64_t n, int64_t k, int act,
                   cudaStream_t s) {
  TmaArgs g{};
  g.M = m; g.N = n; g.K = k; g.bias = bias; g.act = act;
  return tma_dispatch<T_NT>(x, w, y, preact, k, k, n, g, 1, s);
}
// y = x w^T + x2 w2^T in one pass (x2 (m, k2), w2 (n, k2), both K-major; k and k2 multiples of 32)
bool tma_linear_fwd_dual_ok(const float* x, const float* w, const float* x2, const float* w2, const float* y, int64_t m, int64_t n, int64_t k,
                            int64_t k2) {
  return k % KB == 0 && k2 % KB == 0 && n % 4 == 0 && aligned16(x) && aligned16(w) && aligned16(x2) && aligned16(w2) && aligned16(y);
}
int tma_linear_fwd_dual(const float* x, const float* w, const float* x2, const float* w2, float* y, int64_t m, int64_t n, int64_t k, int64_t k2,
                        cudaStream_t s) {
  TmaArgs g{};
  g.M = m; g.N = n; g.K = k;
  return tma_dispatch<T_NT>(x, w, y, nullptr, k, k, n, g, 1, s, DualSrc{x2, w2, k2});
}
bool tma_linear_bwd_data_ok(const float* dy, const float* w, const float* dx, int64_t m, int64_t n, int64_t k) {
  return n % 4 == 0 && k % 4 == 0 && aligned16(dy) && aligned16(w) && aligned16(dx);
}
int tma_linear_bwd_data(const float* dy, const float* w, float* dx, int64_t m, int64_t n, int64_t k, int accumulate, const float* gelu_pre,
                        cudaStream_t s) {
  TmaArgs g{};
  g.M = m; g.N = k; g.K = n; g.reduce_add = accumulate; g.gelu_pre = gelu_pre;
  return tma_dispatch<T_NN>(dy, w, dx, nullptr, n, k, k, g, 1, s);
}
// sparse convolution forward / backward-data as a gathered NT GEMM: y[o, :] = sum_tap x[table[o, tap], :] w[:, tap, :]^T
bool tma_sparse_conv_ok(const float* x, const float* w, const float* y, int cin, int cout) {
  return cin % KB == 0 && cout % 4 == 0 && aligned16(x) && aligned16(w) && aligned16(y);
}
int tma_sparse_conv_fwd(const float* x, const int* table, const float* w, float* y, int64_t rows_out, int taps, int cin, int cout, int accumulate,
                        cudaStream_t s) {
  TmaArgs g{};
  g.M = rows_out; g.N = cout; g.K = (int64_t)taps * cin; g.reduce_add = accumulate;
  g.gsrc = x; g.gtab = table; g.gtaps = taps; g.gcin = cin;
  // the A tensor map is unused by the GATHER variant: describe w twice so that both maps are valid
  if (cout > 128) return tma_launch<T_NT, 256, true>(w, w, y, nullptr, g.K, g.K, cout, g, 1, s);
  if (cout > 64) return tma_launch<T_NT, 128, true>(w, w, y, nullptr, g.K, g.K, cout, g, 1, s);
  return tma_launch<T_NT, 64, true>(w, w, y, nullptr, g.K, g.K, cout, g, 1, s);
}
// dw (cout, taps*cin), zero-filled by the caller: gathered TN GEMM, split over the output rows
bool tma_sparse_conv_bwd_weight_ok(const float* dy, const float* x, const float* dw, int cin, int cout) {
  return cin % 128 == 0 && cout % 4 == 0 && aligned16(dy) && aligned16(x) && aligned16(dw);
}
int tma_sparse_conv_bwd_weight(const float* dy, const float* x, const int* table, float* dw, int64_t rows_out, int taps, int cin, int cout,
                               cudaStream_t s) {
  TmaArgs g{};
  const int64_t kk = (int64_t)taps * cin;
  g.M = cout; g.N = kk; g.K = rows_out; g.reduce_add = 1;
  g.gsrc = x; g.gtab = table; g.gtaps = taps; g.gcin = cin;
  int64_t tiles = (int64_t)cdiv(cout, UM) * cdiv(kk, 256);
  // split the reduction so that tiles x splits fills ONE wave of the persistent grid (rounding up instead -- 5 tiles x 30 =
  // 150 work items on 148 CTAs -- makes two CTAs run two items each and doubles the kernel's makespan)
  int64_t want = kNumSMs / tiles, maxs = (rows_out + 8 * KB - 1) / (8 * KB);
  if (want > maxs) want = maxs;
  // the B tensor map is unused by the gather variant: describe dy twice
  return tma_launch<T_TN, 256, true>(dy, dy, dw, nullptr, cout, kk, kk, g, (int)(want < 1 ? 1 : want), s);
}
bool tma_linear_bwd_weight_ok(const float* dy, const float* x, const float* dw, int64_t m, int64_t n, int64_t k) {
  return n % 4 == 0 && k % 4 == 0 && aligned16(dy) && aligned16(x) && aligned16(dw);
}
// dw must be zero-filled by the caller (split reduction adds into it)
int tma_linear_bwd_weight(const float* dy, const float* x, float* dw, int64_t m, int64_t n, int64_t k, cudaStream_t s) {
  TmaArgs g{};
  g.M = n; g.N = k; g.K = m; g.reduce_add = 1;
  int bn = k > 128 ? 256 : (k > 64 ? 128 : 64);
  int64_t tiles = (int64_t)cdiv(n, UM) * cdiv(k, bn);
  int64_t want = kNumSMs / tiles, maxs = (m + 8 * KB - 1) / (8 * KB);   // one wave: see tma_sparse_conv_bwd_weight
  if (want > maxs) want = maxs;
  return tma_dispatch<T_TN>(dy, x, dw, nullptr, n, k, k, g, (int)(want < 1 ? 1 : want), s);
}

}  // namespace tmae

extern "C" int tmae_debug_set_trace(void* device_u64, int64_t capacity) {
  tmae::g_trace_buf = (unsigned long long*)device_u64;
  tmae::g_trace_cap = device_u64 ? capacity : 0;
  return 0;
}
