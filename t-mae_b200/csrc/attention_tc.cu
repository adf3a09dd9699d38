// Window attention core of the bf16-storage mode on the 5th-generation tensor cores (replaces flat2window +
// _scaled_cosine_attention + window2flat: pcdet/models/model_utils/cosine_msa.py:114-176, sst_basic_block.py:22-54,
// wca_block.py:26-67).  Windows hold <= 64 tokens and heads are 16 / 32 wide, far below a UMMA tile, so windows are PACKED:
// a 128-row tile holds 4 windows of the <= 32-token class or 2 of the <= 64 class (the partition sorts windows by level; the
// <= 16-token class runs on the warp kernels at the end of this file, or as 8 windows per tile with attn_small_warps = 0),
// S = Q K^T is one M = N = 128 tcgen05.mma per head whose off-diagonal (cross-window) blocks are simply never read, and P V runs
// over the tile's 128 keys with P block-diagonal (the off-diagonal part of the P tile is zeroed once: block sizes only grow along a
// CTA's tile list).  The ragged key mask (tokens per window) is applied in the TMEM -> register softmax.
// One work item = (tile, 64-channel group = 4 heads of 16 or 2 heads of 32); 384 threads per CTA, one CTA per SM, persistent:
//   warps 0..7   two softmax warpgroups (thread = tile row = TMEM lane).  Forward: alternate heads, each with its own S slot in TMEM
//                and its own P tile; backward: the two halves of a row's key columns, partial D exchanged through shared memory;
//                then the epilogue (O, or dQ / dK / dV taken back through the L2 normalisation) as 64-byte row pieces
//   warp  8      MMA issuer (one lane): tcgen05.mma.kind::f16, tcgen05.commit -> mbarriers
//   warps 9..11  gather producers: the tile's row indices are looked up once per item into shared memory, then q / k / v (and dO)
//                rows arrive by 16-byte cp.async (zero fill for empty slots) straight into the SWIZZLE_128B K-major layout,
//                double-buffered per item
// q and k arrive L2-normalised per head (the projection epilogue, gemm_bf16.cu E_QKV); logits = q.k / max(tau, tau_min).
// The backward keeps S, dP, dQ, dK, dV accumulators in TMEM and recomputes P from the saved log-sum-exp; D_i = sum_j P_ij dP_ij is
// taken from the registers that hold both, so O is never re-read by the tile kernel.
// Both kernels are launched with programmatic stream serialization: set-up (barriers, zeroed P tiles, TMEM) overlaps the previous
// kernel's tail, pdl_wait() stands in front of the first global access.
#include <cuda_bf16.h>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace tmae {
namespace atc {

typedef __nv_bfloat16 bf16;

constexpr int TILE_BYTES = 128 * 128 * 2;     // one operand tile of a 128-channel group: 128 rows x 256 bytes (two 64-element spans)
constexpr int SPAN_BYTES = 128 * 128;         // one 64-element span of 128 rows

#ifdef TMAE_ATTN_CHECK
__device__ int* g_chk = nullptr;   // host-mapped: [0] = count, then 8 ints per record
__device__ __noinline__ bool chk_fail(int code, long long v0, long long v1, long long v2, long long v3) {
  if (!g_chk) return true;
  const int i = atomicAdd(g_chk, 1);
  if (i < 60) {
    volatile int* p = g_chk + 8 + i * 8;
    p[0] = code; p[1] = blockIdx.x; p[2] = threadIdx.x; p[3] = (int)v0; p[4] = (int)v1; p[5] = (int)v2; p[6] = (int)v3; p[7] = (int)(v3 >> 32);
  }
  __threadfence_system();
  return true;
}
#define CHK(cond, code, v0, v1, v2, v3) ((cond) ? false : chk_fail(code, v0, v1, v2, v3))
#else
#define CHK(cond, code, v0, v1, v2, v3) false
#endif

struct Args {
  const bf16* q; const bf16* k; const bf16* v;
  bf16* o; float* lse;
  const int* qtok; const int* qcnt; const int* ktok; const int* kcnt;
  const int* n_win; const int* small_end; const int* mid_end;
  const float* tau; float tau_min;
  int C, H, hd, ldq, ldk, ldv;
  // backward
  const bf16* dout; const float* inv_q; const float* inv_k; int ld_inv_q, ld_inv_k;
  bf16* dq; bf16* dk; bf16* dv; float* dtau;
  int skip_small;   // 1: the <= 16-token windows run on the warp kernels, the tcgen05 tile list starts at the 32-token class
  int64_t mq, mkv, max_windows;   // bounds (debug checks)
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tile list of one launch: windows are level-sorted, class boundaries live on the device
struct TileInfo { int T, w0, w_end; };   // tokens per window slot block, first window, end of the class
__device__ __forceinline__ void tile_counts(const Args& a, int& se, int& me, int& nw, int& t16, int& t32, int& t64) {
  nw = __ldg(a.n_win);
  se = min(__ldg(a.small_end), nw);
  me = min(max(__ldg(a.mid_end), se), nw);
  if (CHK(nw >= 0 && nw <= a.max_windows && se >= 0, 2, nw, a.max_windows, se, me)) { nw = 0; se = 0; me = 0; }
  t16 = a.skip_small ? 0 : (se + 7) / 8;
  t32 = (me - se + 3) / 4;
  t64 = (nw - me + 1) / 2;
}
__device__ __forceinline__ TileInfo tile_info(int tile, int se, int me, int nw, int t16, int t32) {
  TileInfo ti;
  if (tile < t16) { ti.T = 16; ti.w0 = tile * 8; ti.w_end = se; }
  else if (tile < t16 + t32) { ti.T = 32; ti.w0 = se + (tile - t16) * 4; ti.w_end = me; }
  else { ti.T = 64; ti.w0 = me + (tile - t16 - t32) * 2; ti.w_end = nw; }
  return ti;
}

// shared-memory byte offset of the 16-byte chunk holding elements [8 c, 8 c + 8) of row r in a K-major SWIZZLE_128B tile whose
// 64-element spans are SPAN_BYTES apart
__device__ __forceinline__ uint32_t sw_off(int r, int c16) {
  return (uint32_t)((c16 >> 3) * SPAN_BYTES + (r >> 3) * 1024 + (r & 7) * 128 + (((c16 & 7) ^ (r & 7)) << 4));
}

// gathers one operand tile (128 rows x 128 channels of column block `col0`) through a token table
template <int NT>   // gather threads
__device__ __forceinline__ void gather_tile(uint32_t dst, const bf16* __restrict__ src, int ld, int col0, const int* __restrict__ tok,
                                            const int* __restrict__ cnt, const TileInfo& ti, int gt) {
  const int c16 = gt & 15, rs = gt >> 4;         // 16 lanes copy one row's 256 bytes; NT / 16 rows per pass
  constexpr int RP = NT / 16;
  const int tshift = ti.T == 16 ? 4 : (ti.T == 32 ? 5 : 6);
#pragma unroll 4
  for (int r = rs; r < 128; r += RP) {
    const int w = ti.w0 + (r >> tshift), slot = r & (ti.T - 1);
    const bool ok = w < ti.w_end && slot < __ldg(cnt + w);
    const int row = ok ? __ldg(tok + (int64_t)w * 64 + slot) : 0;
    cp_async16_zfill(dst + sw_off(r, c16), src + (int64_t)row * ld + col0 + c16 * 8, ok ? 16u : 0u);
  }
}

// ============================================================================================== forward
// One item = (tile, 64-channel group = 4 heads of 16 or 2 heads of 32).  Two softmax warpgroups work on alternate heads, each with
// its own S slot in TMEM and its own P tile in shared memory, so the masked softmax of head h + 1 overlaps the P V product of head h.
//   warps 0..3 / 4..7   softmax warpgroup 0 / 1 (thread = tile row = TMEM lane), then each writes half of the O columns
//   warp  8             MMA issuer            warps 9..11   gather producers (cp.async through the token tables)
// shared: [2 x (Q, K, V tiles of 128 rows x 128 bytes)] [2 x P tile (128 x 256 bytes)]      TMEM: S0 | S1 | O0 | O1 (128,128,64,64 columns)
constexpr int BT_BYTES = 128 * 64 * 2;    // one operand tile of a 64-channel group
constexpr int WG = 2, WG_WARPS = 4, MMA_WARP = WG * WG_WARPS, GW0 = MMA_WARP + 1, GATHER_WARPS = 3;
constexpr int THREADS = 32 * (GW0 + GATHER_WARPS);   // 384

__device__ __forceinline__ uint32_t sw_off64(int r, int c) {   // 16-byte chunk c (0..7) of row r in a single-span tile
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void lds_v4(uint32_t addr, uint32_t* u) {
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]) : "r"(addr));
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// voxel row of every tile row (or -1): looked up ONCE per item into shared memory -- the per-copy lookups (two dependent global loads
// in front of every cp.async) made the gather warps the slowest stage of the pipeline
template <int NT>
__device__ __forceinline__ void tile_rows(int* __restrict__ rows, const int* __restrict__ tok, const int* __restrict__ cnt, const TileInfo& ti, int gt,
                                          int64_t nmax = 0, int64_t wmax = 0) {
  const int tshift = ti.T == 16 ? 4 : (ti.T == 32 ? 5 : 6);
  for (int r = gt; r < 128; r += NT) {
    const int w = ti.w0 + (r >> tshift), slot = r & (ti.T - 1);
    const bool ok = w < ti.w_end && slot < __ldg(cnt + w);
    rows[r] = ok ? __ldg(tok + (int64_t)w * 64 + slot) : -1;
    if (ok && CHK(rows[r] >= 0 && rows[r] < nmax && w >= 0 && w < wmax, 1, rows[r], nmax, w, slot)) rows[r] = -1;
  }
}
template <int NT>
__device__ __forceinline__ void gather_tile64(uint32_t dst, const bf16* __restrict__ src, int ld, int col0, const int* __restrict__ rows, int gt) {
  const int c = gt & 7, rs = gt >> 3;            // 8 lanes copy one row's 128 bytes
  constexpr int RP = NT / 8;
#pragma unroll 4
  for (int r = rs; r < 128; r += RP) {
    const int row = rows[r];
    cp_async16_zfill(dst + sw_off64(r, c), src + (int64_t)(row < 0 ? 0 : row) * ld + col0 + c * 8, row < 0 ? 0u : 16u);
  }
}

// masked softmax of one row over its N own key columns x[0..N): returns 1 / sum and the natural-log lse; x becomes exp(. - max)
template <int N>
__device__ __forceinline__ void row_softmax(float* x, int nk, float scale, float& rinv, float& lse) {
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    x[j] = j < nk ? x[j] * scale : -INFINITY;
    mx = fmaxf(mx, x[j]);
  }
  const float mref = mx == -INFINITY ? 0.f : mx;
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < N; ++j) { x[j] = exp2f(x[j] - mref); sum += x[j]; }
  rinv = sum > 0.f ? 1.f / sum : 0.f;
  lse = (mref + log2f(fmaxf(sum, 1e-30f))) * 0.6931471805599453f;
}

template <int HD>
__global__ void __launch_bounds__(THREADS, 1) attn_tc_fwd_kernel(Args a) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: see tc_common.cuh pdl_trigger
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int HG = 64 / HD;                  // heads per 64-channel group (4 or 2): even, so head h uses slot h & 1 = its warpgroup
  constexpr int T_S = 0, T_O = 256;
  __shared__ uint64_t in_full[2], in_empty[2], s_full[2], s_empty[2], p_full[2], p_empty[2], o_full[2], o_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float rs_x[128][HG];              // 1 / rowsum per (row, head of the group): the O epilogue needs both warpgroups' heads
  __shared__ int rows_q[2][128], rows_k[2][128];   // voxel row per tile row (per input buffer), written and read by the gather warps
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* p_tiles = smem + 2 * 3 * BT_BYTES;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      bar_init(&in_full[b], GATHER_WARPS * 32); bar_init(&in_empty[b], 1);
      bar_init(&s_full[b], 1); bar_init(&s_empty[b], WG_WARPS);
      bar_init(&p_full[b], WG_WARPS); bar_init(&p_empty[b], 1);
      bar_init(&o_full[b], 1); bar_init(&o_empty[b], WG * WG_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the P tiles start as zeros: rows only ever write their own (growing) diagonal block
  for (int i = threadIdx.x; i < 2 * TILE_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(p_tiles)[i] = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_wait();   // set-up above (barriers, zeroed P tiles, TMEM) may overlap the previous kernel; nothing below may

  int se, me, nw, t16, t32, t64;
  tile_counts(a, se, me, nw, t16, t32, t64);
  const int G = a.C / 64;
  const int n_items = (t16 + t32 + t64) * G;

  if (warp >= GW0) {
    // ------------------------------------------------------------ gather producers
    const int gt = (warp - GW0) * 32 + lane;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int buf = it & 1, use = it >> 1;
      const TileInfo ti = tile_info(item / G, se, me, nw, t16, t32);
      const int col0 = (item % G) * 64;
      if (use > 0) bar_wait(&in_empty[buf], (use - 1) & 1);
      const uint32_t base = s_u32(smem + buf * 3 * BT_BYTES);
      tile_rows<GATHER_WARPS * 32>(rows_q[buf], a.qtok, a.qcnt, ti, gt, a.mq, a.max_windows);
      tile_rows<GATHER_WARPS * 32>(rows_k[buf], a.ktok, a.kcnt, ti, gt, a.mkv, a.max_windows);
      named_bar(3, GATHER_WARPS * 32);
      gather_tile64<GATHER_WARPS * 32>(base, a.q, a.ldq, col0, rows_q[buf], gt);
      gather_tile64<GATHER_WARPS * 32>(base + BT_BYTES, a.k, a.ldk, col0, rows_k[buf], gt);
      gather_tile64<GATHER_WARPS * 32>(base + 2 * BT_BYTES, a.v, a.ldv, col0, rows_k[buf], gt);
      cp_async_arrive_noinc(&in_full[buf]);
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t id_s = idesc_bf16(0, 0, 128), id_pv = idesc_bf16(0, 1, HD);
      int it = 0;
      int s_uses[2] = {0, 0}, p_uses[2] = {0, 0};
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int buf = it & 1, use = it >> 1;
        const uint32_t q_addr = s_u32(smem + buf * 3 * BT_BYTES), k_addr = q_addr + BT_BYTES, v_addr = q_addr + 2 * BT_BYTES;
        bar_wait(&in_full[buf], use & 1);
        fence_async_smem();          // cp.async (generic proxy) writes -> UMMA (async proxy) reads
        tc_fence_after();
        auto issue_s = [&](int h) {
          const int sb = h & 1;
          if (s_uses[sb] > 0) bar_wait(&s_empty[sb], (s_uses[sb] - 1) & 1);
          tc_fence_after();
          const uint32_t off = (uint32_t)(h * HD * 2);
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk)
            umma_bf16(tmem + T_S + sb * 128, desc_sw128(q_addr + off + kk * 32, 16, 1024), desc_sw128(k_addr + off + kk * 32, 16, 1024), id_s, kk ? 1u : 0u);
          commit_to(&s_full[sb]);
          ++s_uses[sb];
        };
        issue_s(0);
        issue_s(1);
        const int ob = it & 1, ou = it >> 1;
        if (ou > 0) bar_wait(&o_empty[ob], (ou - 1) & 1);   // the epilogue drained this O slot
#pragma unroll
        for (int h = 0; h < HG; ++h) {
          const int pb = h & 1;
          bar_wait(&p_full[pb], p_uses[pb] & 1);
          tc_fence_after();
          const uint32_t p_addr = s_u32(p_tiles + pb * TILE_BYTES);
          const uint32_t voff = (uint32_t)(h * HD * 2);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {   // 128 keys, 16 per instruction
            const uint64_t da = desc_sw128(p_addr + (kk >> 2) * SPAN_BYTES + (kk & 3) * 32, 16, 1024);
            const uint64_t db = desc_sw128(v_addr + voff + kk * 2048, 8192, 1024);
            umma_bf16(tmem + T_O + ob * 64 + h * HD, da, db, id_pv, kk ? 1u : 0u);
          }
          commit_to(&p_empty[pb]);
          ++p_uses[pb];
          if (h + 2 < HG) issue_s(h + 2);
        }
        commit_to(&o_full[ob]);
        commit_to(&in_empty[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups + epilogue: thread = tile row
    const int wg = warp >> 2, wq = warp & 3;
    const int r = wq * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    const float scale = 1.4426950408889634f / fmaxf(__ldg(a.tau), a.tau_min);      // logits in log2 units
    int it = 0, s_use = 0, p_use = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const TileInfo ti = tile_info(item / G, se, me, nw, t16, t32);
      const int col0 = (item % G) * 64;
      const int tshift = ti.T == 16 ? 4 : (ti.T == 32 ? 5 : 6);
      const int w = ti.w0 + (r >> tshift), slot = r & (ti.T - 1);
      const bool w_ok = w < ti.w_end;
      const int nk = w_ok ? __ldg(a.kcnt + w) : 0;
      const bool row_ok = w_ok && slot < __ldg(a.qcnt + w);
      const int64_t vrow = row_ok ? __ldg(a.qtok + (int64_t)w * 64 + slot) : 0;
      const int L = ti.T < 32 ? 32 : ti.T;                 // columns this warp reads (warp-uniform)
      const int cb = (r / L) * L;                          // first of them
      const bool upper = ti.T == 16 && (r & 16);           // <= 16-token class: this row's keys are the upper half of the 32 columns
      const uint32_t p_row = s_u32(p_tiles + wg * TILE_BYTES);
      for (int h = wg; h < HG; h += 2) {
        bar_wait(&s_full[wg], s_use & 1);
        tc_fence_after();
        float rinv, lse;
        if (L == 64) {
          float x[64];
          uint32_t u[32];
          ld_tmem32(tmem + lane_addr + T_S + wg * 128 + cb, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(u[j]);
          ld_tmem32(tmem + lane_addr + T_S + wg * 128 + cb + 32, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) x[32 + j] = __uint_as_float(u[j]);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) bar_arrive(&s_empty[wg]);
          row_softmax<64>(x, nk, scale, rinv, lse);
          if (p_use > 0) bar_wait(&p_empty[wg], (p_use - 1) & 1);   // the previous P V on this tile has consumed it
#pragma unroll
          for (int c = 0; c < 8; ++c)
            sts_v4(p_row + sw_off(r, (cb >> 3) + c), pack_bf16(x[8 * c], x[8 * c + 1]), pack_bf16(x[8 * c + 2], x[8 * c + 3]),
                   pack_bf16(x[8 * c + 4], x[8 * c + 5]), pack_bf16(x[8 * c + 6], x[8 * c + 7]));
        } else {
          uint32_t u[32];
          ld_tmem32(tmem + lane_addr + T_S + wg * 128 + cb, u);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) bar_arrive(&s_empty[wg]);
          if (ti.T == 32) {
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(u[j]);
            row_softmax<32>(x, nk, scale, rinv, lse);
            if (p_use > 0) bar_wait(&p_empty[wg], (p_use - 1) & 1);
#pragma unroll
            for (int c = 0; c < 4; ++c)
              sts_v4(p_row + sw_off(r, (cb >> 3) + c), pack_bf16(x[8 * c], x[8 * c + 1]), pack_bf16(x[8 * c + 2], x[8 * c + 3]),
                     pack_bf16(x[8 * c + 4], x[8 * c + 5]), pack_bf16(x[8 * c + 6], x[8 * c + 7]));
          } else {   // T == 16: own half only; the other half of the 32 columns is never written in this class and stays zero
            float x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(upper ? u[16 + j] : u[j]);
            row_softmax<16>(x, nk, scale, rinv, lse);
            if (p_use > 0) bar_wait(&p_empty[wg], (p_use - 1) & 1);
#pragma unroll
            for (int c = 0; c < 2; ++c)
              sts_v4(p_row + sw_off(r, (cb >> 3) + (upper ? 2 : 0) + c), pack_bf16(x[8 * c], x[8 * c + 1]), pack_bf16(x[8 * c + 2], x[8 * c + 3]),
                     pack_bf16(x[8 * c + 4], x[8 * c + 5]), pack_bf16(x[8 * c + 6], x[8 * c + 7]));
          }
        }
        ++s_use;
        fence_async_smem();
        __syncwarp();
        if (lane == 0) bar_arrive(&p_full[wg]);
        ++p_use;
        rs_x[r][h] = rinv;
        if (row_ok) a.lse[vrow * a.H + col0 / HD + h] = lse;
      }
      // ---- epilogue: O (64 columns = all heads of the group) / rowsum; warpgroup wg writes columns [32 wg, 32 wg + 32)
      named_bar(1, WG * WG_WARPS * 32);     // both warpgroups' 1 / rowsum are in rs_x
      const int ob = it & 1, ou = it >> 1;
      bar_wait(&o_full[ob], ou & 1);
      tc_fence_after();
      {
        uint32_t u[32];
        ld_tmem32(tmem + lane_addr + T_O + ob * 64 + wg * 32, u);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(&o_empty[ob]);
        if (row_ok) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float s0 = rs_x[r][(wg * 32 + 2 * j) / HD];
            pk[j] = pack_bf16(__uint_as_float(u[2 * j]) * s0, __uint_as_float(u[2 * j + 1]) * s0);
          }
          bf16* dst = a.o + vrow * a.C + col0 + wg * 32;
          st_global_v8_u32(dst, pk);
          st_global_v8_u32(dst + 16, pk + 8);
        }
      }
      named_bar(1, WG * WG_WARPS * 32);     // rs_x is free for the next item
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ============================================================================================== backward
// Same items and warp roles.  The two softmax warpgroups split the COLUMNS of a head's S / dP row block (P is recomputed from the
// saved log-sum-exp, so the only row statistic to exchange is D = sum_j P_ij dP_ij); then each takes half of the output columns.
// shared: [2 x (Q, K, V, dO tiles)] [P tile] [dS tile]      TMEM: S | dP | dQ | dK | dV (128,128,64,64,64 columns)
template <int N>   // this thread's N columns of one row
__device__ __forceinline__ void bwd_row_part(float* x, float* dp, int j0, int nk, bool row_ok, float scale, float lse2, float& Dp, float& tsa, float& tsb) {
  // x: raw cosine similarities -> p ; dp: dP (masked) ; partial D, and the two sums behind the temperature gradient
  Dp = 0.f; tsa = 0.f; tsb = 0.f;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const bool valid = row_ok && (unsigned)(j0 + j) < (unsigned)nk;
    const float sj = valid ? x[j] : 0.f;
    const float p = valid ? exp2f(fmaf(sj, scale, -lse2)) : 0.f;
    const float d = valid ? dp[j] : 0.f;
    const float pd_ = p * d;
    Dp += pd_;
    tsa = fmaf(pd_, sj, tsa);
    tsb = fmaf(p, sj, tsb);
    x[j] = p;
    dp[j] = d;
  }
}

template <int HD>
__global__ void __launch_bounds__(THREADS, 1) attn_tc_bwd_kernel(Args a) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: see tc_common.cuh pdl_trigger
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int HG = 64 / HD;
  constexpr int T_S = 0, T_DP = 128, T_DQ = 256, T_DK = 320, T_DV = 384;
  constexpr int NSM = WG * WG_WARPS;           // softmax warps
  __shared__ uint64_t in_full[2], in_empty[2], sp_full, sp_empty, pds_full, pds_empty, acc_full, acc_empty;
  __shared__ uint32_t tmem_slot;
  __shared__ float d_x[2][2][128];             // partial D per (head parity, warpgroup, row): double-buffered, one barrier per head
  __shared__ int rows_q[2][128], rows_k[2][128];   // voxel row per tile row (per input buffer), written and read by the gather warps
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* p_tile = smem + 2 * 4 * BT_BYTES;
  uint8_t* ds_tile = p_tile + TILE_BYTES;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) { bar_init(&in_full[b], GATHER_WARPS * 32); bar_init(&in_empty[b], 1 + NSM); }
    bar_init(&sp_full, 1); bar_init(&sp_empty, NSM);
    bar_init(&pds_full, NSM); bar_init(&pds_empty, 1);
    bar_init(&acc_full, 1); bar_init(&acc_empty, NSM);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * TILE_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(p_tile)[i] = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_wait();   // set-up above (barriers, zeroed P tiles, TMEM) may overlap the previous kernel; nothing below may

  int se, me, nw, t16, t32, t64;
  tile_counts(a, se, me, nw, t16, t32, t64);
  const int G = a.C / 64;
  const int n_items = (t16 + t32 + t64) * G;

  if (warp >= GW0) {
    // ------------------------------------------------------------ gather producers
    const int gt = (warp - GW0) * 32 + lane;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int buf = it & 1, use = it >> 1;
      const TileInfo ti = tile_info(item / G, se, me, nw, t16, t32);
      const int col0 = (item % G) * 64;
      if (use > 0) bar_wait(&in_empty[buf], (use - 1) & 1);
      const uint32_t base = s_u32(smem + buf * 4 * BT_BYTES);
      tile_rows<GATHER_WARPS * 32>(rows_q[buf], a.qtok, a.qcnt, ti, gt, a.mq, a.max_windows);
      tile_rows<GATHER_WARPS * 32>(rows_k[buf], a.ktok, a.kcnt, ti, gt, a.mkv, a.max_windows);
      named_bar(3, GATHER_WARPS * 32);
      gather_tile64<GATHER_WARPS * 32>(base, a.q, a.ldq, col0, rows_q[buf], gt);
      gather_tile64<GATHER_WARPS * 32>(base + BT_BYTES, a.k, a.ldk, col0, rows_k[buf], gt);
      gather_tile64<GATHER_WARPS * 32>(base + 2 * BT_BYTES, a.v, a.ldv, col0, rows_k[buf], gt);
      gather_tile64<GATHER_WARPS * 32>(base + 3 * BT_BYTES, a.dout, a.C, col0, rows_q[buf], gt);
      cp_async_arrive_noinc(&in_full[buf]);
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t id_s = idesc_bf16(0, 0, 128);          // S, dP: A, B K-major
      const uint32_t id_kn = idesc_bf16(0, 1, HD);          // dQ = dS K : A K-major, B MN-major
      const uint32_t id_nn = idesc_bf16(1, 1, HD);          // dV = P^T dO, dK = dS^T Q : A, B MN-major
      const uint32_t p_addr = s_u32(p_tile), ds_addr = s_u32(ds_tile);
      int it = 0, sp_uses = 0, pds_uses = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int buf = it & 1, use = it >> 1;
        const uint32_t q_addr = s_u32(smem + buf * 4 * BT_BYTES), k_addr = q_addr + BT_BYTES, v_addr = q_addr + 2 * BT_BYTES,
                       do_addr = q_addr + 3 * BT_BYTES;
        bar_wait(&in_full[buf], use & 1);
        fence_async_smem();
        tc_fence_after();
        auto issue_sp = [&](int h) {
          const uint32_t hoff = (uint32_t)(h * HD * 2);
          if (sp_uses > 0) bar_wait(&sp_empty, (sp_uses - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            umma_bf16(tmem + T_S, desc_sw128(q_addr + hoff + kk * 32, 16, 1024), desc_sw128(k_addr + hoff + kk * 32, 16, 1024), id_s, kk ? 1u : 0u);
            umma_bf16(tmem + T_DP, desc_sw128(do_addr + hoff + kk * 32, 16, 1024), desc_sw128(v_addr + hoff + kk * 32, 16, 1024), id_s, kk ? 1u : 0u);
          }
          commit_to(&sp_full);
          ++sp_uses;
        };
        issue_sp(0);
        if (it > 0) bar_wait(&acc_empty, (it - 1) & 1);   // the epilogue of the previous item drained dQ / dK / dV
        for (int h = 0; h < HG; ++h) {
          const uint32_t hoff = (uint32_t)(h * HD * 2);
          bar_wait(&pds_full, pds_uses & 1);
          tc_fence_after();
          if (h + 1 < HG) issue_sp(h + 1);     // the softmax threads have S / dP of head h in registers: the next pair overlaps these MMAs
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {   // 128 rows of the contraction, 16 per instruction
            const uint64_t p_mn = desc_sw128(p_addr + kk * 2048, SPAN_BYTES, 1024);
            const uint64_t ds_mn = desc_sw128(ds_addr + kk * 2048, SPAN_BYTES, 1024);
            const uint64_t ds_k = desc_sw128(ds_addr + (kk >> 2) * SPAN_BYTES + (kk & 3) * 32, 16, 1024);
            umma_bf16(tmem + T_DV + h * HD, p_mn, desc_sw128(do_addr + hoff + kk * 2048, 8192, 1024), id_nn, kk ? 1u : 0u);
            umma_bf16(tmem + T_DQ + h * HD, ds_k, desc_sw128(k_addr + hoff + kk * 2048, 8192, 1024), id_kn, kk ? 1u : 0u);
            umma_bf16(tmem + T_DK + h * HD, ds_mn, desc_sw128(q_addr + hoff + kk * 2048, 8192, 1024), id_nn, kk ? 1u : 0u);
          }
          commit_to(&pds_empty);
          ++pds_uses;
        }
        commit_to(&acc_full);
        commit_to(&in_empty[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax / dS warpgroups + epilogue: thread = tile row, half the columns
    const int wg = warp >> 2, wq = warp & 3;
    const int r = wq * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    const float tau_raw = __ldg(a.tau);
    const float tau_c = fmaxf(tau_raw, a.tau_min);
    const float inv_tau = 1.f / tau_c;
    const float scale = 1.4426950408889634f * inv_tau;
    const uint32_t pbase = s_u32(p_tile), dbase = s_u32(ds_tile);
    float dtau_acc = 0.f;
    int it = 0, sp_uses = 0, pds_uses = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      const TileInfo ti = tile_info(item / G, se, me, nw, t16, t32);
      const int col0 = (item % G) * 64;
      const int tshift = ti.T == 16 ? 4 : (ti.T == 32 ? 5 : 6);
      const int w = ti.w0 + (r >> tshift), slot = r & (ti.T - 1);
      const bool w_ok = w < ti.w_end;
      CHK(!w_ok || (w >= 0 && w < a.max_windows), 5, w, a.max_windows, ti.w0, ti.w_end);
      const int nk = w_ok ? __ldg(a.kcnt + w) : 0;
      const bool row_ok = w_ok && slot < __ldg(a.qcnt + w);     // this tile row is a query row
      const bool key_ok = w_ok && slot < nk;                    // ... and / or a key row
      int64_t qrow = row_ok ? __ldg(a.qtok + (int64_t)w * 64 + slot) : 0;
      int64_t krow = key_ok ? __ldg(a.ktok + (int64_t)w * 64 + slot) : 0;
      if (CHK(qrow >= 0 && qrow < a.mq, 3, qrow, a.mq, w, slot)) qrow = 0;
      if (CHK(krow >= 0 && krow < a.mkv, 4, krow, a.mkv, w, slot)) krow = 0;
      const int L = ti.T < 32 ? 32 : ti.T;
      const int cb = (r / L) * L;
      const int koff = (r >> tshift) * ti.T - cb;              // this row's first key column inside its L-block (0, or 16 in the 16-class)
      for (int h = 0; h < HG; ++h) {
        const int head = col0 / HD + h;
        const float lse2 = row_ok ? __ldg(a.lse + qrow * a.H + head) * 1.4426950408889634f : 0.f;
        bar_wait(&sp_full, sp_uses & 1);
        tc_fence_after();
        float Dp, tsa, tsb;
        if (L == 64) {
          // columns [cb + 32 wg, + 32)
          float x[32], dp[32];
          uint32_t u[32];
          ld_tmem32(tmem + lane_addr + T_S + cb + wg * 32, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(u[j]);
          ld_tmem32(tmem + lane_addr + T_DP + cb + wg * 32, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) dp[j] = __uint_as_float(u[j]);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) bar_arrive(&sp_empty);
          bwd_row_part<32>(x, dp, wg * 32 - koff, nk, row_ok, scale, lse2, Dp, tsa, tsb);
          d_x[h & 1][wg][r] = Dp;
          named_bar(1, NSM * 32);
          const float D = Dp + d_x[h & 1][wg ^ 1][r];
          dtau_acc += tsa - D * tsb;
          if (pds_uses > 0) bar_wait(&pds_empty, (pds_uses - 1) & 1);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pp[4], pd[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 8 * c + 2 * e;
              pp[e] = pack_bf16(x[j], x[j + 1]);
              pd[e] = pack_bf16(x[j] * (dp[j] - D) * inv_tau, x[j + 1] * (dp[j + 1] - D) * inv_tau);
            }
            const uint32_t off = sw_off(r, (cb >> 3) + wg * 4 + c);
            sts_v4(pbase + off, pp[0], pp[1], pp[2], pp[3]);
            sts_v4(dbase + off, pd[0], pd[1], pd[2], pd[3]);
          }
        } else {
          // columns [cb + 16 wg, + 16)
          float x[16], dp[16];
          uint32_t u[16];
          ld_tmem16(tmem + lane_addr + T_S + cb + wg * 16, u);
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(u[j]);
          ld_tmem16(tmem + lane_addr + T_DP + cb + wg * 16, u);
#pragma unroll
          for (int j = 0; j < 16; ++j) dp[j] = __uint_as_float(u[j]);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) bar_arrive(&sp_empty);
          bwd_row_part<16>(x, dp, wg * 16 - koff, nk, row_ok, scale, lse2, Dp, tsa, tsb);
          d_x[h & 1][wg][r] = Dp;
          named_bar(1, NSM * 32);
          const float D = Dp + d_x[h & 1][wg ^ 1][r];
          dtau_acc += tsa - D * tsb;
          if (pds_uses > 0) bar_wait(&pds_empty, (pds_uses - 1) & 1);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t pp[4], pd[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 8 * c + 2 * e;
              pp[e] = pack_bf16(x[j], x[j + 1]);
              pd[e] = pack_bf16(x[j] * (dp[j] - D) * inv_tau, x[j + 1] * (dp[j + 1] - D) * inv_tau);
            }
            const uint32_t off = sw_off(r, (cb >> 3) + wg * 2 + c);
            sts_v4(pbase + off, pp[0], pp[1], pp[2], pp[3]);
            sts_v4(dbase + off, pd[0], pd[1], pd[2], pd[3]);
          }
        }
        ++sp_uses;
        fence_async_smem();
        __syncwarp();
        if (lane == 0) bar_arrive(&pds_full);
        ++pds_uses;
      }
      // ---- epilogue: dQ, dK (back through the L2 normalisation) and dV rows; warpgroup wg takes columns [32 wg, 32 wg + 32)
      bar_wait(&acc_full, it & 1);
      tc_fence_after();
      const uint32_t q_s = s_u32(smem + buf * 4 * BT_BYTES), k_s = q_s + BT_BYTES;
      auto out_rows = [&](int tcol, uint32_t unit_tile, const float* inv, int ld_inv, bf16* dst, int ldd, int64_t grow, bool ok, bool normalised) {
        uint32_t u[32];
        ld_tmem32(tmem + lane_addr + tcol + wg * 32, u);
        if (!ok) return;
        float g[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) g[j] = __uint_as_float(u[j]);
        if (normalised) {   // d(x / |x|) : (g - u (u . g)) / |x| per head, u = the unit vector the forward stored
#pragma unroll
          for (int hh = 0; hh < 32 / HD; ++hh) {
            float un[HD];
#pragma unroll
            for (int cc = 0; cc < HD / 8; ++cc) {
              uint32_t w4[4];
              lds_v4(unit_tile + sw_off64(r, (wg * 32 + hh * HD) / 8 + cc), w4);
#pragma unroll
              for (int e = 0; e < 4; ++e) { const float2 f = unpack_bf16(w4[e]); un[cc * 8 + 2 * e] = f.x; un[cc * 8 + 2 * e + 1] = f.y; }
            }
            float dot = 0.f;
#pragma unroll
            for (int e = 0; e < HD; ++e) dot = fmaf(un[e], g[hh * HD + e], dot);
            const float iv = __ldg(inv + grow * ld_inv + (col0 + wg * 32) / HD + hh);
#pragma unroll
            for (int e = 0; e < HD; ++e) g[hh * HD + e] = (g[hh * HD + e] - un[e] * dot) * iv;
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(g[2 * j], g[2 * j + 1]);
        bf16* d = dst + grow * ldd + col0 + wg * 32;
        st_global_v8_u32(d, pk);
        st_global_v8_u32(d + 16, pk + 8);
      };
      out_rows(T_DQ, q_s, a.inv_q, a.ld_inv_q, a.dq, a.ldq, qrow, row_ok, true);
      out_rows(T_DK, k_s, a.inv_k, a.ld_inv_k, a.dk, a.ldk, krow, key_ok, true);
      out_rows(T_DV, 0, nullptr, 0, a.dv, a.ldv, krow, key_ok, false);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { bar_arrive(&acc_empty); bar_arrive(&in_empty[buf]); }
    }
    dtau_acc = warp_sum(dtau_acc);
    if (lane == 0 && a.dtau && tau_raw > a.tau_min && dtau_acc != 0.f) atomicAdd(a.dtau, -dtau_acc * inv_tau * inv_tau);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ============================================================================================== <= 16-token windows
// 72 % of a scan's windows hold fewer than 16 voxels (3 on average): in a 128-row UMMA tile they would be 80-90 % padding, and the
// per-row softmax work is paid for every padded row.  They run one WARP per (window, 128-channel group) instead: a lane owns 4
// channels of every row (a row = one coalesced 256-byte access, a head = 4 or 8 adjacent lanes), the stationary side (K, V) is
// loaded once into registers with every load of the window in flight, the moving side (q; q, dO, o, lse in the backward) is staged
// by cp.async into lane-private shared memory, and the per-window work is instantiated for padded key counts 4 / 8 / 16 so every
// inner loop is branch-free.  Same arithmetic contract as the tcgen05 kernels: q, k are unit vectors, fp32 math, bf16 I/O.
constexpr int SW_T = 16, SW_THREADS = 128, SW_KC = 8, MAXT = 64;

__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }
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
__device__ __forceinline__ float4 bf4(const uint2& u) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint2 pk4(const float4& v) { return make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w)); }
// predicated 64-bit read-only load of 4 bf16 (zeros when off): keeps the per-key loops branch-free
__device__ __forceinline__ float4 ldg_bf4_if(const bf16* p, bool on) {
  uint2 r;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t"
      "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\t"
      "@p ld.global.nc.v2.b32 {%0, %1}, [%2];\n\t}"
      : "=r"(r.x), "=r"(r.y)
      : "l"(p), "r"((int)on));
  return bf4(r);
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int HL, int NK, int ROWS>
__device__ __forceinline__ void small_fwd_window(const Args& a, const uint2 (*qs)[32], int nq, int nk, int tokv, int col, int lane, float scale) {
  constexpr float LN2 = 0.6931471805599453f;
  float4 kr[NK], vr[NK];
#pragma unroll
  for (int j = 0; j < NK; ++j) {
    const int t = __shfl_sync(0xffffffffu, tokv, j);
    kr[j] = ldg_bf4_if(a.k + (int64_t)t * a.ldk + col, j < nk);
    vr[j] = ldg_bf4_if(a.v + (int64_t)t * a.ldv + col, j < nk);
  }
  cp_async_wait_all();
  for (int i = 0; i < nq; i += ROWS) {
    float4 q[ROWS], acc[ROWS];
    float sc[ROWS][NK], m[ROWS], l[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      q[r] = bf4(qs[(i + r) & (SW_T - 1)][lane]);
      scale4(q[r], scale);
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
        const float p = fast_exp2(sc[r][j] - m[r]);
        l[r] += p;
        axpy4(acc[r], p, vr[j]);
      }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      if (i + r < nq) {
        const int row = __shfl_sync(0xffffffffu, tokv, SW_T + ((i + r) & (SW_T - 1)));
        scale4(acc[r], __fdividef(1.f, l[r]));
        *reinterpret_cast<uint2*>(a.o + (int64_t)row * a.C + col) = pk4(acc[r]);
        if ((lane & (HL - 1)) == 0) a.lse[(int64_t)row * a.H + (col / (HL * 4))] = (m[r] + __log2f(l[r])) * LN2;
      }
    }
  }
}

template <int HD>
__global__ void __launch_bounds__(SW_THREADS, 4) attn_small_bf16_fwd_kernel(Args a) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: see tc_common.cuh pdl_trigger
  constexpr int HL = HD / 4;
  __shared__ uint2 q_s[SW_THREADS / 32][SW_T][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int n_small = min(*a.small_end, *a.n_win);
  const int groups = a.C >> 7;
  const int n_items = n_small * groups;
  const float scale = 1.4426950408889634f / fmaxf(*a.tau, a.tau_min);  // scores are kept in the log2 domain
  for (int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < n_items; item += warps) {
    const int w = item / groups, col = (item - w * groups) * 128 + lane * 4;
    const int nq = min(a.qcnt[w], SW_T), nk = min(a.kcnt[w], SW_T);
    const int tokv = lane < SW_T ? a.ktok[w * MAXT + lane] : a.qtok[w * MAXT + lane - SW_T];
#pragma unroll
    for (int i = 0; i < SW_T; ++i) {
      const int t = __shfl_sync(0xffffffffu, tokv, SW_T + i);
      if (i < nq) cp_async8(&q_s[wib][i][lane], a.q + (int64_t)t * a.ldq + col);
    }
    if (nk <= 4) small_fwd_window<HL, 4, 2>(a, q_s[wib], nq, nk, tokv, col, lane, scale);
    else if (nk <= 8) small_fwd_window<HL, 8, 2>(a, q_s[wib], nq, nk, tokv, col, lane, scale);
    else small_fwd_window<HL, 16, 1>(a, q_s[wib], nq, nk, tokv, col, lane, scale);
  }
}

// Fused backward: dQ, dK, dV and dtau in one pass; keys in register chunks of up to 8; windows with 9..16 keys run the query loop
// twice and add the second chunk's dQ contribution to the row written by the first (the normalisation Jacobian is linear).
struct SmallBwdSmem {
  uint2 q[SW_T][32], g[SW_T][32], o[SW_T][32];
  float lse[SW_T][32], qinv[SW_T][32];
};
template <int HL, int NC>
__device__ __forceinline__ void small_bwd_chunk(const Args& a, const SmallBwdSmem& S, int nq, int kc, int nc, int tokv, int col, int head, int lane,
                                                float inv_tau, float& dtau_acc, bool first) {
  const bool lead = (lane & (HL - 1)) == 0;
  float4 kr[NC], vr[NC], dk[NC], dv[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int t = __shfl_sync(0xffffffffu, tokv, (kc + j) & (SW_T - 1));
    kr[j] = ldg_bf4_if(a.k + (int64_t)t * a.ldk + col, j < nc);
    vr[j] = ldg_bf4_if(a.v + (int64_t)t * a.ldv + col, j < nc);
    dk[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    dv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (first) cp_async_wait_all();
  for (int i = 0; i < nq; ++i) {
    const float4 qh = bf4(S.q[i][lane]), g = bf4(S.g[i][lane]), o = bf4(S.o[i][lane]);
    const float L = S.lse[i][lane], qinv = S.qinv[i][lane];
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
      const float p = j < nc ? __expf(s[j] - L) : 0.f;
      const float ds = p * (dp[j] - D);
      if (lead) dtau_acc = fmaf(-ds, s[j], dtau_acc);
      const float dsl = ds * inv_tau;
      axpy4(dqh, dsl, kr[j]);
      axpy4(dk[j], dsl, qh);
      axpy4(dv[j], p, g);
    }
    const float dt = head_sum<HL>(dot4(dqh, qh));   // through q_hat = q / |q| with the 1 / |q| of the projection epilogue
    float4 dq = make_float4((dqh.x - qh.x * dt) * qinv, (dqh.y - qh.y * dt) * qinv, (dqh.z - qh.z * dt) * qinv, (dqh.w - qh.w * dt) * qinv);
    const int row = __shfl_sync(0xffffffffu, tokv, SW_T + i);
    uint2* dst = reinterpret_cast<uint2*>(a.dq + (int64_t)row * a.ldq + col);
    if (!first) { const float4 prev = bf4(*dst); dq.x += prev.x; dq.y += prev.y; dq.z += prev.z; dq.w += prev.w; }
    *dst = pk4(dq);
  }
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int t = __shfl_sync(0xffffffffu, tokv, (kc + j) & (SW_T - 1));
    const float dt = head_sum<HL>(dot4(dk[j], kr[j]));
    if (j < nc) {
      const float kinv = __ldg(a.inv_k + (int64_t)t * a.ld_inv_k + head);
      const float4 r = make_float4((dk[j].x - kr[j].x * dt) * kinv, (dk[j].y - kr[j].y * dt) * kinv, (dk[j].z - kr[j].z * dt) * kinv,
                                   (dk[j].w - kr[j].w * dt) * kinv);
      *reinterpret_cast<uint2*>(a.dk + (int64_t)t * a.ldk + col) = pk4(r);
      *reinterpret_cast<uint2*>(a.dv + (int64_t)t * a.ldv + col) = pk4(dv[j]);
    }
  }
}

template <int HD>
__global__ void __launch_bounds__(SW_THREADS, 2) attn_small_bf16_bwd_kernel(Args a, const bf16* __restrict__ o) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: see tc_common.cuh pdl_trigger
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
    int tokv = lane < SW_T ? a.ktok[w * MAXT + lane] : a.qtok[w * MAXT + lane - SW_T];
    CHK(w >= 0 && w < a.max_windows && nq >= 0 && nk >= 0, 6, w, a.max_windows, nq, nk);
    if (lane < SW_T ? (lane < nk && CHK(tokv >= 0 && tokv < a.mkv, 7, tokv, a.mkv, w, (long long)a.ktok)) : (lane - SW_T < nq && CHK(tokv >= 0 && tokv < a.mq, 8, tokv, a.mq, w, lane))) tokv = 0;
#pragma unroll
    for (int i = 0; i < SW_T; ++i) {
      const int t = __shfl_sync(0xffffffffu, tokv, SW_T + i);
      if (i < nq) {
        const int64_t off = (int64_t)t * a.C + col;
        cp_async8(&S.q[i][lane], a.q + (int64_t)t * a.ldq + col);
        cp_async8(&S.g[i][lane], a.dout + off);
        cp_async8(&S.o[i][lane], o + off);
        cp_async4(&S.lse[i][lane], a.lse + (int64_t)t * a.H + head);
        cp_async4(&S.qinv[i][lane], a.inv_q + (int64_t)t * a.ld_inv_q + head);
      }
    }
    for (int kc = 0; kc < nk; kc += SW_KC) {
      const int nc = min(SW_KC, nk - kc);
      if (nc <= 4) small_bwd_chunk<HL, 4>(a, S, nq, kc, nc, tokv, col, head, lane, inv_tau, dtau_acc, kc == 0);
      else small_bwd_chunk<HL, 8>(a, S, nq, kc, nc, tokv, col, head, lane, inv_tau, dtau_acc, kc == 0);
    }
    if (nk == 0) cp_async_wait_all();  // nothing consumed the staged rows of this window
  }
  dtau_acc = warp_sum(dtau_acc);
  if (lane == 0 && a.dtau && tau_raw > a.tau_min && dtau_acc != 0.f) atomicAdd(a.dtau, dtau_acc * inv_tau);
}

}  // namespace atc
extern int g_small_on_warps;
namespace atc {

static int check(const Args& a) {
  if (a.C % 128 != 0 || (a.hd != 16 && a.hd != 32) || a.H * a.hd != a.C) return -1;
  if (a.ldq % 8 || a.ldk % 8 || a.ldv % 8) return -1;
  return 0;
}

}  // namespace atc

#ifdef TMAE_ATTN_CHECK
static int* g_chk_host = nullptr;
extern "C" __attribute__((visibility("default"))) int* tmae_debug_attn_check_buffer() {
  if (!g_chk_host) {
    cudaHostAlloc((void**)&g_chk_host, 4096, cudaHostAllocMapped);
    memset(g_chk_host, 0, 4096);
    int* dptr = nullptr;
    cudaHostGetDevicePointer((void**)&dptr, g_chk_host, 0);
    cudaMemcpyToSymbol(atc::g_chk, &dptr, sizeof(dptr));
  }
  return g_chk_host;
}
#endif
extern int g_bf16_gemm_pdl;
namespace atc {
// launch with programmatic stream serialization (the kernel's set-up overlaps the previous kernel's tail; see pdl_wait in the kernels)
template <typename K>
static void launch_pdl(K kern, int grid, int threads, size_t smem, cudaStream_t s, const Args& a) {
  if (!g_bf16_gemm_pdl) { kern<<<grid, threads, smem, s>>>(a); return; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, a);
}
}  // namespace atc
int g_small_on_warps = 1;   // measurement switch (tmae_set_option "attn_small_warps"): 0 = every window class on the tcgen05 tiles
using namespace atc;

int attn_tc_fwd(const void* q, const void* k, const void* v, void* o, float* lse, const tmae_layer_tables* T, const float* tau, float tau_min,
                int64_t m_q, int64_t m_kv, int c, int heads, int ldq, int ldk, int ldv, cudaStream_t s) {
  Args a{};
  a.q = (const bf16*)q; a.k = (const bf16*)k; a.v = (const bf16*)v; a.o = (bf16*)o; a.lse = lse;
  a.qtok = T->qtok; a.qcnt = T->qcnt; a.ktok = T->ktok; a.kcnt = T->kcnt; a.n_win = T->n_win; a.small_end = T->small_end; a.mid_end = T->mid_end;
  a.tau = tau; a.tau_min = tau_min; a.C = c; a.H = heads; a.hd = c / heads; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.mq = m_q; a.mkv = m_kv; a.max_windows = T->max_windows;
  if (check(a)) { set_error("attn_tc_fwd: channels must be a multiple of 128 with head_dim 16 or 32, row pitches multiples of 8"); return TMAE_ERR_INVALID_ARG; }
  if (T->max_windows <= 0 || m_q <= 0) return 0;
  const size_t smem = 2 * 3 * BT_BYTES + 2 * TILE_BYTES + 1024;
  // tiles <= windows / 2 + 3 in the worst packing, x channel groups
  int64_t max_items = (T->max_windows / 2 + 3) * (c / 64);
  int grid = (int)(max_items < kNumSMs ? max_items : kNumSMs);
  a.skip_small = g_small_on_warps;
  if (a.skip_small) {
    ProfScope prof("attn_small_bf16_fwd", 0, 0, s);
    int64_t items = T->max_windows * (c / 128);
    int64_t warps = items < (int64_t)kNumSMs * 32 ? items : (int64_t)kNumSMs * 32;
    if (a.hd == 16) attn_small_bf16_fwd_kernel<16><<<cdiv(warps * 32, SW_THREADS), SW_THREADS, 0, s>>>(a);
    else attn_small_bf16_fwd_kernel<32><<<cdiv(warps * 32, SW_THREADS), SW_THREADS, 0, s>>>(a);
  }
  const double bytes = 2.0 * c * (2.0 * m_q + 2.0 * m_kv);
  ProfScope prof("attn_tc_fwd", 0, bytes, s);
  if (a.hd == 16) {
    if (smem_attr_once((const void*)attn_tc_fwd_kernel<16>, (int)smem)) return TMAE_ERR_CUDA;
    launch_pdl(attn_tc_fwd_kernel<16>, grid, THREADS, smem, s, a);
  } else {
    if (smem_attr_once((const void*)attn_tc_fwd_kernel<32>, (int)smem)) return TMAE_ERR_CUDA;
    launch_pdl(attn_tc_fwd_kernel<32>, grid, THREADS, smem, s, a);
  }
  if (cudaGetLastError() != cudaSuccess) { set_error("attn_tc_fwd: launch failed"); return TMAE_ERR_CUDA; }
  return 0;
}

// dq / dk / dv share the pitches of q / k / v (gradients mirror the packed projection layout); inv_q / inv_k: 1 / |.| per (row, head)
// from the projection epilogue with row pitches ld_inv_*; dtau accumulates (atomicAdd) the temperature gradient.
int attn_tc_bwd(const void* dout, const void* q, const void* k, const void* v, const void* o, const float* lse, const float* inv_q, int ld_inv_q,
                const float* inv_k, int ld_inv_k, void* dq, void* dk, void* dv, float* dtau, const tmae_layer_tables* T, const float* tau, float tau_min,
                int64_t m_q, int64_t m_kv, int c, int heads, int ldq, int ldk, int ldv, cudaStream_t s) {
  Args a{};
  a.q = (const bf16*)q; a.k = (const bf16*)k; a.v = (const bf16*)v; a.lse = (float*)lse; a.dout = (const bf16*)dout;
  a.inv_q = inv_q; a.inv_k = inv_k; a.ld_inv_q = ld_inv_q; a.ld_inv_k = ld_inv_k; a.dq = (bf16*)dq; a.dk = (bf16*)dk; a.dv = (bf16*)dv; a.dtau = dtau;
  a.qtok = T->qtok; a.qcnt = T->qcnt; a.ktok = T->ktok; a.kcnt = T->kcnt; a.n_win = T->n_win; a.small_end = T->small_end; a.mid_end = T->mid_end;
  a.tau = tau; a.tau_min = tau_min; a.C = c; a.H = heads; a.hd = c / heads; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.mq = m_q; a.mkv = m_kv; a.max_windows = T->max_windows;
  if (check(a)) { set_error("attn_tc_bwd: channels must be a multiple of 128 with head_dim 16 or 32, row pitches multiples of 8"); return TMAE_ERR_INVALID_ARG; }
  if (T->max_windows <= 0 || m_q <= 0) return 0;
  const size_t smem = 2 * 4 * BT_BYTES + 2 * TILE_BYTES + 1024;
  int64_t max_items = (T->max_windows / 2 + 3) * (c / 64);
  int grid = (int)(max_items < kNumSMs ? max_items : kNumSMs);
  a.skip_small = g_small_on_warps && o != nullptr;   // the warp kernels take D = dO . O from the saved output
  if (a.skip_small) {
    ProfScope prof("attn_small_bf16_bwd", 0, 0, s);
    int64_t items = T->max_windows * (c / 128);
    int64_t warps = items < (int64_t)kNumSMs * 24 ? items : (int64_t)kNumSMs * 24;
    constexpr int ssm = (int)sizeof(SmallBwdSmem) * (SW_THREADS / 32);
    if (smem_attr_once((const void*)attn_small_bf16_bwd_kernel<16>, ssm) || smem_attr_once((const void*)attn_small_bf16_bwd_kernel<32>, ssm)) return TMAE_ERR_CUDA;
    if (a.hd == 16) attn_small_bf16_bwd_kernel<16><<<cdiv(warps * 32, SW_THREADS), SW_THREADS, ssm, s>>>(a, (const bf16*)o);
    else attn_small_bf16_bwd_kernel<32><<<cdiv(warps * 32, SW_THREADS), SW_THREADS, ssm, s>>>(a, (const bf16*)o);
  }
  const double bytes = 2.0 * c * (3.0 * m_q + 4.0 * m_kv);
  ProfScope prof("attn_tc_bwd", 0, bytes, s);
  if (a.hd == 16) {
    if (smem_attr_once((const void*)attn_tc_bwd_kernel<16>, (int)smem)) return TMAE_ERR_CUDA;
    launch_pdl(attn_tc_bwd_kernel<16>, grid, THREADS, smem, s, a);
  } else {
    if (smem_attr_once((const void*)attn_tc_bwd_kernel<32>, (int)smem)) return TMAE_ERR_CUDA;
    launch_pdl(attn_tc_bwd_kernel<32>, grid, THREADS, smem, s, a);
  }
  if (cudaGetLastError() != cudaSuccess) { set_error("attn_tc_bwd: launch failed"); return TMAE_ERR_CUDA; }
  return 0;
}
bool attn_tc_available() { return true; }

}  // namespace tmae

extern "C" int tmae_bf16_window_attention_bwd(const void* dout, const void* q, const void* k, const void* v, const void* o, const float* lse, const float* inv_q,
                                              int32_t ld_inv_q, const float* inv_k, int32_t ld_inv_k, void* dq, void* dk, void* dv, float* dtau,
                                              const int32_t* qtok, const int32_t* qcnt, const int32_t* ktok, const int32_t* kcnt, const int32_t* n_win,
                                              const int32_t* small_end, const int32_t* mid_end, int64_t max_windows, const float* tau, float tau_min,
                                              int32_t channels, int32_t heads, int32_t ld_q, int32_t ld_k, int32_t ld_v, int64_t rows_q, int64_t rows_kv,
                                              void* stream) {
  tmae_layer_tables T{};
  T.qtok = qtok; T.qcnt = qcnt; T.ktok = ktok; T.kcnt = kcnt; T.n_win = n_win; T.small_end = small_end; T.mid_end = mid_end; T.max_windows = max_windows;
  return tmae::attn_tc_bwd(dout, q, k, v, o, lse, inv_q, ld_inv_q, inv_k, ld_inv_k, dq, dk, dv, dtau, &T, tau, tau_min, rows_q, rows_kv, channels,
                           heads, ld_q, ld_k, ld_v, (cudaStream_t)stream);
}

extern "C" int tmae_bf16_window_attention_fwd(const void* q, const void* k, const void* v, void* o, float* lse, const int32_t* qtok, const int32_t* qcnt,
                                              const int32_t* ktok, const int32_t* kcnt, const int32_t* n_win, const int32_t* small_end,
                                              const int32_t* mid_end, int64_t max_windows, const float* tau, float tau_min, int32_t channels,
                                              int32_t heads, int32_t ld_q, int32_t ld_k, int32_t ld_v, int64_t rows_q, int64_t rows_kv, void* stream) {
  tmae_layer_tables T{};
  T.qtok = qtok; T.qcnt = qcnt; T.ktok = ktok; T.kcnt = kcnt; T.n_win = n_win; T.small_end = small_end; T.mid_end = mid_end; T.max_windows = max_windows;
  return tmae::attn_tc_fwd(q, k, v, o, lse, &T, tau, tau_min, rows_q, rows_kv, channels, heads, ld_q, ld_k, ld_v, (cudaStream_t)stream);
}
