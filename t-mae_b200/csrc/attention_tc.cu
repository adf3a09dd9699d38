// Window attention core of the bf16-storage mode on the 5th-generation tensor cores (replaces flat2window +
// _scaled_cosine_attention + window2flat: pcdet/models/model_utils/cosine_msa.py:114-176, sst_basic_block.py:22-54,
// wca_block.py:26-67).  Windows hold <= 64 tokens and heads are 16 / 32 wide, far below a UMMA tile, so windows are PACKED:
// a 128-row tile holds 8 windows of the <= 16-token class, 4 of the <= 32 class or 2 of the <= 64 class (the partition sorts
// windows by level), S = Q K^T is one M = N = 128 tcgen05.mma per head whose off-diagonal (cross-window) blocks are simply
// never read, and P V runs over the tile's 128 keys with P block-diagonal (the off-diagonal part of the P tile is zeroed once:
// block sizes only grow along a CTA's tile list).  The ragged key mask (tokens per window) is applied in the TMEM -> register
// softmax.  One work item = (tile, 128-channel group = 8 heads of 16 or 4 heads of 32):
//   warps 0..3   softmax + epilogue: thread = tile row = TMEM lane; S row block from TMEM, masked softmax (fp32), P (bf16) into
//                the swizzled shared tile; at the end O (all heads, 128 columns) from TMEM, 1 / rowsum, one 256-byte row store
//   warp  4      MMA issuer (one lane): S(h) = Qh Kh^T into a double-buffered TMEM slot, O[:, h] = P(h) Vh; tcgen05.commit -> mbarriers
//   warps 5..7   gather producers: q / k / v rows of the tile's windows through the token tables, 16-byte cp.async (zero fill for
//                empty slots) straight into the SWIZZLE_128B K-major layout, double-buffered per item
// q and k arrive L2-normalised per head (the projection epilogue, gemm_bf16.cu E_QKV); logits = q.k / max(tau, tau_min).
// The backward kernel uses 64-channel groups (TMEM: S, dP, dQ, dK, dV accumulators) and recomputes P from the saved
// log-sum-exp; D_i = sum_j P_ij dP_ij is taken from the registers that hold both, so O is never re-read.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace tmae {
namespace atc {

typedef __nv_bfloat16 bf16;

constexpr int SM_WARPS = 4, GATHER_WARPS = 3;   // 8 warps = 256 threads: the softmax warps may use up to 255 registers
constexpr int THREADS = 32 * (SM_WARPS + 1 + GATHER_WARPS);   // 256
constexpr int GW0 = SM_WARPS + 1;
constexpr int TILE_BYTES = 128 * 128 * 2;     // one operand tile of a 128-channel group: 128 rows x 256 bytes (two 64-element spans)
constexpr int SPAN_BYTES = 128 * 128;         // one 64-element span of 128 rows

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
  t16 = (se + 7) / 8;
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
// shared: [2 x (Q, K, V tiles)] [P tile]   TMEM: S0 | S1 | O0 | O1 (128 columns each)
template <int HD>
__global__ void __launch_bounds__(THREADS, 1) attn_tc_fwd_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int HG = 128 / HD;                 // heads per 128-channel group
  __shared__ uint64_t qkv_full[2], qkv_empty[2], s_full[2], s_empty[2], p_full, p_empty, o_full[2], o_empty[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* p_tile = smem + 2 * 3 * TILE_BYTES;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      bar_init(&qkv_full[b], GATHER_WARPS * 32); bar_init(&qkv_empty[b], 1);
      bar_init(&s_full[b], 1); bar_init(&s_empty[b], SM_WARPS);
      bar_init(&o_full[b], 1); bar_init(&o_empty[b], SM_WARPS);
    }
    bar_init(&p_full, SM_WARPS); bar_init(&p_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the P tile starts as zeros: rows only ever write their own (growing) diagonal block
  for (int i = threadIdx.x; i < TILE_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(p_tile)[i] = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  if (warp == SM_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  int se, me, nw, t16, t32, t64;
  tile_counts(a, se, me, nw, t16, t32, t64);
  const int G = a.C / 128;
  const int n_items = (t16 + t32 + t64) * G;

  if (warp >= GW0) {
    // ------------------------------------------------------------ gather producers
    const int gt = (warp - GW0) * 32 + lane;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int buf = it & 1, use = it >> 1;
      const TileInfo ti = tile_info(item / G, se, me, nw, t16, t32);
      const int col0 = (item % G) * 128;
      if (use > 0) bar_wait(&qkv_empty[buf], (use - 1) & 1);
      const uint32_t base = s_u32(smem + buf * 3 * TILE_BYTES);
      gather_tile<GATHER_WARPS * 32>(base, a.q, a.ldq, col0, a.qtok, a.qcnt, ti, gt);
      gather_tile<GATHER_WARPS * 32>(base + TILE_BYTES, a.k, a.ldk, col0, a.ktok, a.kcnt, ti, gt);
      gather_tile<GATHER_WARPS * 32>(base + 2 * TILE_BYTES, a.v, a.ldv, col0, a.ktok, a.kcnt, ti, gt);
      cp_async_arrive_noinc(&qkv_full[buf]);
    }
  } else if (warp == SM_WARPS) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t id_s = idesc_bf16(0, 0, 128), id_pv = idesc_bf16(0, 1, HD);
      const uint32_t p_addr = s_u32(p_tile);
      int it = 0, s_uses = 0, p_uses = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int buf = it & 1, use = it >> 1;
        const uint32_t q_addr = s_u32(smem + buf * 3 * TILE_BYTES), k_addr = q_addr + TILE_BYTES, v_addr = q_addr + 2 * TILE_BYTES;
        bar_wait(&qkv_full[buf], use & 1);
        fence_async_smem();          // cp.async (generic proxy) writes -> UMMA (async proxy) reads
        tc_fence_after();
        auto issue_s = [&](int h) {
          const int sb = s_uses & 1, su = s_uses >> 1;
          if (su > 0) bar_wait(&s_empty[sb], (su - 1) & 1);
          tc_fence_after();
          const uint32_t off = (uint32_t)(((h * HD) >> 6) * SPAN_BYTES + ((h * HD) & 63) * 2);
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk)
            umma_bf16(tmem + sb * 128, desc_sw128(q_addr + off + kk * 32, 16, 1024), desc_sw128(k_addr + off + kk * 32, 16, 1024), id_s, kk ? 1u : 0u);
          commit_to(&s_full[sb]);
          ++s_uses;
        };
        issue_s(0);
        const int ob = it & 1, ou = it >> 1;
        for (int h = 0; h < HG; ++h) {
          if (h + 1 < HG) issue_s(h + 1);
          if (h == 0 && ou > 0) bar_wait(&o_empty[ob], (ou - 1) & 1);   // the epilogue drained this O slot
          bar_wait(&p_full, p_uses & 1);
          tc_fence_after();
          const uint32_t voff = (uint32_t)(((h * HD) >> 6) * SPAN_BYTES + ((h * HD) & 63) * 2);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {   // 128 keys, 16 per instruction
            const uint64_t da = desc_sw128(p_addr + (kk >> 2) * SPAN_BYTES + (kk & 3) * 32, 16, 1024);
            const uint64_t db = desc_sw128(v_addr + voff + kk * 2048, SPAN_BYTES, 1024);
            umma_bf16(tmem + 256 + ob * 128 + h * HD, da, db, id_pv, kk ? 1u : 0u);
          }
          commit_to(&p_empty);
          ++p_uses;
        }
        commit_to(&o_full[ob]);
        commit_to(&qkv_empty[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax + epilogue: thread = tile row
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const float tau_c = fmaxf(__ldg(a.tau), a.tau_min);
    const float scale = 1.4426950408889634f / tau_c;      // logits in log2 units
    int it = 0, s_uses = 0, p_uses = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const TileInfo ti = tile_info(item / G, se, me, nw, t16, t32);
      const int col0 = (item % G) * 128;
      const int tshift = ti.T == 16 ? 4 : (ti.T == 32 ? 5 : 6);
      const int w = ti.w0 + (r >> tshift), slot = r & (ti.T - 1);
      const bool w_ok = w < ti.w_end;
      const int nk = w_ok ? __ldg(a.kcnt + w) : 0;
      const bool row_ok = w_ok && slot < __ldg(a.qcnt + w);
      const int64_t vrow = row_ok ? __ldg(a.qtok + (int64_t)w * 64 + slot) : 0;
      const int L = ti.T < 32 ? 32 : ti.T;                 // columns this warp reads (warp-uniform)
      const int cb = (r / L) * L;                          // first of them
      const int koff = (r >> tshift) * ti.T - cb;          // this row's keys start here inside the block (0, or 16 for the upper half-warp)
      float rsum[HG], rlse[HG];
      for (int h = 0; h < HG; ++h) {
        const int sb = s_uses & 1, su = s_uses >> 1;
        bar_wait(&s_full[sb], su & 1);
        tc_fence_after();
        float x[64];
        {
          uint32_t u[32];
          ld_tmem32(tmem + lane_addr + sb * 128 + cb, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(u[j]) * scale;
          if (L == 64) {
            ld_tmem32(tmem + lane_addr + sb * 128 + cb + 32, u);
#pragma unroll
            for (int j = 0; j < 32; ++j) x[32 + j] = __uint_as_float(u[j]) * scale;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(&s_empty[sb]);
        ++s_uses;
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const bool valid = j < L && (unsigned)(j - koff) < (unsigned)nk;
          x[j] = valid ? x[j] : -INFINITY;
          mx = fmaxf(mx, x[j]);
        }
        const float mref = mx == -INFINITY ? 0.f : mx;
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) { x[j] = exp2f(x[j] - mref); sum += x[j]; }
        rsum[h] = sum > 0.f ? 1.f / sum : 0.f;
        rlse[h] = (mref + log2f(fmaxf(sum, 1e-30f))) * 0.6931471805599453f;
        // P row block (bf16) into the shared tile; the previous head's P V must have consumed it
        if (p_uses > 0) bar_wait(&p_empty, (p_uses - 1) & 1);
        const uint32_t prow = s_u32(p_tile);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c * 8 < L) {
            const uint4 pk = make_uint4(pack_bf16(x[8 * c], x[8 * c + 1]), pack_bf16(x[8 * c + 2], x[8 * c + 3]),
                                        pack_bf16(x[8 * c + 4], x[8 * c + 5]), pack_bf16(x[8 * c + 6], x[8 * c + 7]));
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(prow + sw_off(r, (cb >> 3) + c)), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) bar_arrive(&p_full);
        ++p_uses;
      }
      // ---- epilogue: O (128 columns = all heads of the group) / rowsum -> one 256-byte row
      const int ob = it & 1, ou = it >> 1;
      bar_wait(&o_full[ob], ou & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t u[32];
        ld_tmem32(tmem + lane_addr + 256 + ob * 128 + c * 32, u);
        if (row_ok) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float s0 = rsum[(c * 32 + 2 * j) / HD];
            pk[j] = pack_bf16(__uint_as_float(u[2 * j]) * s0, __uint_as_float(u[2 * j + 1]) * s0);
          }
          bf16* dst = a.o + vrow * a.C + col0 + c * 32;
          st_global_v8_u32(dst, pk);
          st_global_v8_u32(dst + 16, pk + 8);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) bar_arrive(&o_empty[ob]);
      if (row_ok) {
        float* lp = a.lse + vrow * a.H + (col0 / HD);
#pragma unroll
        for (int h = 0; h < HG; ++h) lp[h] = rlse[h];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == SM_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}


// ============================================================================================== backward
// One item = (tile, 64-channel group = 4 heads of 16 or 2 heads of 32).
// shared: [2 x (Q, K, V, dO tiles of 128 rows x 128 bytes)] [P tile] [dS tile]      TMEM: S | dP | dQ | dK | dV (128,128,64,64,64 columns)
constexpr int BT_BYTES = 128 * 64 * 2;    // one operand tile of a 64-channel group

__device__ __forceinline__ uint32_t sw_off64(int r, int c) {   // 16-byte chunk c (0..7) of row r in a single-span tile
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

template <int NT>
__device__ __forceinline__ void gather_tile64(uint32_t dst, const bf16* __restrict__ src, int ld, int col0, const int* __restrict__ tok,
                                              const int* __restrict__ cnt, const TileInfo& ti, int gt) {
  const int c = gt & 7, rs = gt >> 3;            // 8 lanes copy one row's 128 bytes
  constexpr int RP = NT / 8;
  const int tshift = ti.T == 16 ? 4 : (ti.T == 32 ? 5 : 6);
#pragma unroll 4
  for (int r = rs; r < 128; r += RP) {
    const int w = ti.w0 + (r >> tshift), slot = r & (ti.T - 1);
    const bool ok = w < ti.w_end && slot < __ldg(cnt + w);
    const int row = ok ? __ldg(tok + (int64_t)w * 64 + slot) : 0;
    cp_async16_zfill(dst + sw_off64(r, c), src + (int64_t)row * ld + col0 + c * 8, ok ? 16u : 0u);
  }
}

__device__ __forceinline__ void lds_v4(uint32_t addr, uint32_t* u) {
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]) : "r"(addr));
}

template <int HD>
__global__ void __launch_bounds__(THREADS, 1) attn_tc_bwd_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int HG = 64 / HD;                  // heads per 64-channel group
  constexpr int T_S = 0, T_DP = 128, T_DQ = 256, T_DK = 320, T_DV = 384;
  __shared__ uint64_t in_full[2], in_empty[2], sp_full, sp_empty, pds_full, pds_empty, acc_full, acc_empty;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* p_tile = smem + 2 * 4 * BT_BYTES;
  uint8_t* ds_tile = p_tile + TILE_BYTES;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) { bar_init(&in_full[b], GATHER_WARPS * 32); bar_init(&in_empty[b], 1 + SM_WARPS); }
    bar_init(&sp_full, 1); bar_init(&sp_empty, SM_WARPS);
    bar_init(&pds_full, SM_WARPS); bar_init(&pds_empty, 1);
    bar_init(&acc_full, 1); bar_init(&acc_empty, SM_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * TILE_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(p_tile)[i] = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  if (warp == SM_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

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
      gather_tile64<GATHER_WARPS * 32>(base, a.q, a.ldq, col0, a.qtok, a.qcnt, ti, gt);
      gather_tile64<GATHER_WARPS * 32>(base + BT_BYTES, a.k, a.ldk, col0, a.ktok, a.kcnt, ti, gt);
      gather_tile64<GATHER_WARPS * 32>(base + 2 * BT_BYTES, a.v, a.ldv, col0, a.ktok, a.kcnt, ti, gt);
      gather_tile64<GATHER_WARPS * 32>(base + 3 * BT_BYTES, a.dout, a.C, col0, a.qtok, a.qcnt, ti, gt);
      cp_async_arrive_noinc(&in_full[buf]);
    }
  } else if (warp == SM_WARPS) {
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
        for (int h = 0; h < HG; ++h) {
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
          if (h == 0 && it > 0) bar_wait(&acc_empty, (it - 1) & 1);   // the epilogue of the previous item drained dQ / dK / dV
          bar_wait(&pds_full, pds_uses & 1);
          tc_fence_after();
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
    // ------------------------------------------------------------ softmax / dS + epilogue: thread = tile row
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const float tau_raw = __ldg(a.tau);
    const float tau_c = fmaxf(tau_raw, a.tau_min);
    const float inv_tau = 1.f / tau_c;
    const float scale = 1.4426950408889634f * inv_tau;
    float dtau_acc = 0.f;
    int it = 0, sp_uses = 0, pds_uses = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      const TileInfo ti = tile_info(item / G, se, me, nw, t16, t32);
      const int col0 = (item % G) * 64;
      const int tshift = ti.T == 16 ? 4 : (ti.T == 32 ? 5 : 6);
      const int w = ti.w0 + (r >> tshift), slot = r & (ti.T - 1);
      const bool w_ok = w < ti.w_end;
      const int nk = w_ok ? __ldg(a.kcnt + w) : 0;
      const bool row_ok = w_ok && slot < __ldg(a.qcnt + w);     // this tile row is a query row
      const bool key_ok = w_ok && slot < nk;                    // ... and / or a key row
      const int64_t qrow = row_ok ? __ldg(a.qtok + (int64_t)w * 64 + slot) : 0;
      const int64_t krow = key_ok ? __ldg(a.ktok + (int64_t)w * 64 + slot) : 0;
      const int L = ti.T < 32 ? 32 : ti.T;
      const int cb = (r / L) * L;
      const int koff = (r >> tshift) * ti.T - cb;
      for (int h = 0; h < HG; ++h) {
        const int head = col0 / HD + h;
        const float lse2 = row_ok ? __ldg(a.lse + qrow * a.H + head) * 1.4426950408889634f : 0.f;
        bar_wait(&sp_full, sp_uses & 1);
        tc_fence_after();
        float x[64], dp[64];
        {
          uint32_t u[32];
          ld_tmem32(tmem + lane_addr + T_S + cb, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(u[j]);
          ld_tmem32(tmem + lane_addr + T_DP + cb, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) dp[j] = __uint_as_float(u[j]);
          if (L == 64) {
            ld_tmem32(tmem + lane_addr + T_S + cb + 32, u);
#pragma unroll
            for (int j = 0; j < 32; ++j) x[32 + j] = __uint_as_float(u[j]);
            ld_tmem32(tmem + lane_addr + T_DP + cb + 32, u);
#pragma unroll
            for (int j = 0; j < 32; ++j) dp[32 + j] = __uint_as_float(u[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(&sp_empty);
        ++sp_uses;
        // p_j = softmax probability (recomputed from the saved log-sum-exp), D = sum_j p_j dP_j, dlogit_j = p_j (dP_j - D);
        // temperature: sum_j dlogit_j s_j = sum p dP s - D sum p s, so s is not needed after this pass (x[] then holds p)
        float D = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const bool valid = row_ok && j < L && (unsigned)(j - koff) < (unsigned)nk;
          const float sj = valid ? x[j] : 0.f;
          const float p = valid ? exp2f(fmaf(sj, scale, -lse2)) : 0.f;
          const float d = valid ? dp[j] : 0.f;
          const float pd_ = p * d;
          D += pd_;
          a1 = fmaf(pd_, sj, a1);
          a2 = fmaf(p, sj, a2);
          x[j] = p;
          dp[j] = d;
        }
        const float ts = a1 - D * a2;
        dtau_acc += ts;
        // P and dS (= dlogit / tau) row blocks as bf16 into the shared tiles, once the previous head's MMAs have consumed them
        if (pds_uses > 0) bar_wait(&pds_empty, (pds_uses - 1) & 1);
        const uint32_t pbase = s_u32(p_tile), dbase = s_u32(ds_tile);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c * 8 < L) {
            uint32_t pp[4], pd[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 8 * c + 2 * e;
              pp[e] = pack_bf16(x[j], x[j + 1]);
              pd[e] = pack_bf16(x[j] * (dp[j] - D) * inv_tau, x[j + 1] * (dp[j + 1] - D) * inv_tau);
            }
            const uint32_t off = sw_off(r, (cb >> 3) + c);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(pbase + off), "r"(pp[0]), "r"(pp[1]), "r"(pp[2]), "r"(pp[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dbase + off), "r"(pd[0]), "r"(pd[1]), "r"(pd[2]), "r"(pd[3]) : "memory");
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) bar_arrive(&pds_full);
        ++pds_uses;
      }
      // ---- epilogue: dQ, dK (back through the L2 normalisation) and dV rows of this 64-channel group
      bar_wait(&acc_full, it & 1);
      tc_fence_after();
      const uint32_t q_s = s_u32(smem + buf * 4 * BT_BYTES), k_s = q_s + BT_BYTES;
      auto out_rows = [&](int tcol, uint32_t unit_tile, const float* inv, int ld_inv, bf16* dst, int ldd, int64_t grow, bool ok, bool normalised) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t u[32];
          ld_tmem32(tmem + lane_addr + tcol + c * 32, u);
          if (!ok) continue;
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
                lds_v4(unit_tile + sw_off64(r, (c * 32 + hh * HD) / 8 + cc), w4);
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float2 f = unpack_bf16(w4[e]); un[cc * 8 + 2 * e] = f.x; un[cc * 8 + 2 * e + 1] = f.y; }
              }
              float dot = 0.f;
#pragma unroll
              for (int e = 0; e < HD; ++e) dot = fmaf(un[e], g[hh * HD + e], dot);
              const float iv = __ldg(inv + grow * ld_inv + (col0 + c * 32) / HD + hh);
#pragma unroll
              for (int e = 0; e < HD; ++e) g[hh * HD + e] = (g[hh * HD + e] - un[e] * dot) * iv;
            }
          }
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(g[2 * j], g[2 * j + 1]);
          bf16* d = dst + grow * ldd + col0 + c * 32;
          st_global_v8_u32(d, pk);
          st_global_v8_u32(d + 16, pk + 8);
        }
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
  if (warp == SM_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static int check(const Args& a) {
  if (a.C % 128 != 0 || (a.hd != 16 && a.hd != 32) || a.H * a.hd != a.C) return -1;
  if (a.ldq % 8 || a.ldk % 8 || a.ldv % 8) return -1;
  return 0;
}

}  // namespace atc

using namespace atc;

int attn_tc_fwd(const void* q, const void* k, const void* v, void* o, float* lse, const tmae_layer_tables* T, const float* tau, float tau_min,
                int64_t m_q, int64_t m_kv, int c, int heads, int ldq, int ldk, int ldv, cudaStream_t s) {
  Args a{};
  a.q = (const bf16*)q; a.k = (const bf16*)k; a.v = (const bf16*)v; a.o = (bf16*)o; a.lse = lse;
  a.qtok = T->qtok; a.qcnt = T->qcnt; a.ktok = T->ktok; a.kcnt = T->kcnt; a.n_win = T->n_win; a.small_end = T->small_end; a.mid_end = T->mid_end;
  a.tau = tau; a.tau_min = tau_min; a.C = c; a.H = heads; a.hd = c / heads; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  if (check(a)) { set_error("attn_tc_fwd: channels must be a multiple of 128 with head_dim 16 or 32, row pitches multiples of 8"); return TMAE_ERR_INVALID_ARG; }
  if (T->max_windows <= 0 || m_q <= 0) return 0;
  const size_t smem = 2 * 3 * TILE_BYTES + TILE_BYTES + 1024;
  // tiles <= windows / 2 + 3 in the worst packing, x channel groups
  int64_t max_items = (T->max_windows / 2 + 3) * (c / 128);
  int grid = (int)(max_items < kNumSMs ? max_items : kNumSMs);
  const double bytes = 2.0 * c * (2.0 * m_q + 2.0 * m_kv);
  ProfScope prof("attn_tc_fwd", 0, bytes, s);
  if (a.hd == 16) {
    if (smem_attr_once((const void*)attn_tc_fwd_kernel<16>, (int)smem)) return TMAE_ERR_CUDA;
    attn_tc_fwd_kernel<16><<<grid, THREADS, smem, s>>>(a);
  } else {
    if (smem_attr_once((const void*)attn_tc_fwd_kernel<32>, (int)smem)) return TMAE_ERR_CUDA;
    attn_tc_fwd_kernel<32><<<grid, THREADS, smem, s>>>(a);
  }
  if (cudaGetLastError() != cudaSuccess) { set_error("attn_tc_fwd: launch failed"); return TMAE_ERR_CUDA; }
  return 0;
}

// dq / dk / dv share the pitches of q / k / v (gradients mirror the packed projection layout); inv_q / inv_k: 1 / |.| per (row, head)
// from the projection epilogue with row pitches ld_inv_*; dtau accumulates (atomicAdd) the temperature gradient.
int attn_tc_bwd(const void* dout, const void* q, const void* k, const void* v, const void* o, const float* lse, const float* inv_q, int ld_inv_q,
                const float* inv_k, int ld_inv_k, void* dq, void* dk, void* dv, float* dtau, const tmae_layer_tables* T, const float* tau, float tau_min,
                int64_t m_q, int64_t m_kv, int c, int heads, int ldq, int ldk, int ldv, cudaStream_t s) {
  (void)o;
  Args a{};
  a.q = (const bf16*)q; a.k = (const bf16*)k; a.v = (const bf16*)v; a.lse = (float*)lse; a.dout = (const bf16*)dout;
  a.inv_q = inv_q; a.inv_k = inv_k; a.ld_inv_q = ld_inv_q; a.ld_inv_k = ld_inv_k; a.dq = (bf16*)dq; a.dk = (bf16*)dk; a.dv = (bf16*)dv; a.dtau = dtau;
  a.qtok = T->qtok; a.qcnt = T->qcnt; a.ktok = T->ktok; a.kcnt = T->kcnt; a.n_win = T->n_win; a.small_end = T->small_end; a.mid_end = T->mid_end;
  a.tau = tau; a.tau_min = tau_min; a.C = c; a.H = heads; a.hd = c / heads; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  if (check(a)) { set_error("attn_tc_bwd: channels must be a multiple of 128 with head_dim 16 or 32, row pitches multiples of 8"); return TMAE_ERR_INVALID_ARG; }
  if (T->max_windows <= 0 || m_q <= 0) return 0;
  const size_t smem = 2 * 4 * BT_BYTES + 2 * TILE_BYTES + 1024;
  int64_t max_items = (T->max_windows / 2 + 3) * (c / 64);
  int grid = (int)(max_items < kNumSMs ? max_items : kNumSMs);
  const double bytes = 2.0 * c * (3.0 * m_q + 4.0 * m_kv);
  ProfScope prof("attn_tc_bwd", 0, bytes, s);
  if (a.hd == 16) {
    if (smem_attr_once((const void*)attn_tc_bwd_kernel<16>, (int)smem)) return TMAE_ERR_CUDA;
    attn_tc_bwd_kernel<16><<<grid, THREADS, smem, s>>>(a);
  } else {
    if (smem_attr_once((const void*)attn_tc_bwd_kernel<32>, (int)smem)) return TMAE_ERR_CUDA;
    attn_tc_bwd_kernel<32><<<grid, THREADS, smem, s>>>(a);
  }
  if (cudaGetLastError() != cudaSuccess) { set_error("attn_tc_bwd: launch failed"); return TMAE_ERR_CUDA; }
  return 0;
}
bool attn_tc_available() { return true; }

}  // namespace tmae

extern "C" int tmae_bf16_window_attention_bwd(const void* dout, const void* q, const void* k, const void* v, const float* lse, const float* inv_q,
                                              int32_t ld_inv_q, const float* inv_k, int32_t ld_inv_k, void* dq, void* dk, void* dv, float* dtau,
                                              const int32_t* qtok, const int32_t* qcnt, const int32_t* ktok, const int32_t* kcnt, const int32_t* n_win,
                                              const int32_t* small_end, const int32_t* mid_end, int64_t max_windows, const float* tau, float tau_min,
                                              int32_t channels, int32_t heads, int32_t ld_q, int32_t ld_k, int32_t ld_v, int64_t rows_q, int64_t rows_kv,
                                              void* stream) {
  tmae_layer_tables T{};
  T.qtok = qtok; T.qcnt = qcnt; T.ktok = ktok; T.kcnt = kcnt; T.n_win = n_win; T.small_end = small_end; T.mid_end = mid_end; T.max_windows = max_windows;
  return tmae::attn_tc_bwd(dout, q, k, v, nullptr, lse, inv_q, ld_inv_q, inv_k, ld_inv_k, dq, dk, dv, dtau, &T, tau, tau_min, rows_q, rows_kv, channels,
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
