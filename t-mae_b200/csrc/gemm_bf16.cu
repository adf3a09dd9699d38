// TMA-fed tcgen05 GEMMs of the bf16-STORAGE mode (TMAE_PREC_BF16): encoder activations and their gradients live in HBM
// as bf16, the tensor cores consume them directly (tcgen05.mma.kind::f16, bf16 x bf16 -> fp32 in TMEM), and the TMEM
// epilogue writes bf16 -- every activation byte of the encoder is half of what the fp32-storage modes move.
//   warp 0      TMA producer   cp.async.bulk.tensor.2d (SWIZZLE_128B boxes of 64 bf16 = 128 bytes) -> shared ring, mbarrier tx
//   warp 1      MMA issuer     M = 128, N = BN in {64,128,256}, K = 16 per instruction, 4 instructions per 64-wide k-block
//   warps 2..9  epilogue       tcgen05.ld (lane = output row, 32 columns per chunk; two warps per TMEM lane quarter on
//                              alternate chunks) -> fused row arithmetic -> 2 x 256-bit stores of the lane's 64-byte bf16 segment
//   warps 10,11 (GATHER)       cp.async producers of a gathered operand (sparse convolution)
// Modes: NT  C[m,n] = sum_k A[m,k] W[n,k]   (A, B K-major)            linear forward, gathered: sparse conv fwd / dgrad
//        NN  C[m,n] = sum_k A[m,k] B[k,n]   (A K-major, B MN-major)   linear backward-data
//        TN  C[m,n] = sum_k A[k,m] B[k,n]   (A, B MN-major)           weight gradients (fp32 out, split over k, vector atomics)
// Epilogues (what the reference runs as separate ATen kernels, here folded into the pass that owns the accumulator):
//   E_PLAIN  + bias, GELU / ReLU, optional pre-activation copy (or, TMAE_ACT_GELU_DERIV, GELU'(pre-activation) in its place: the
//            backward then multiplies instead of re-deriving), optional * gelu'(pre) (FFN backward), optional C +=
//   E_QKV    the window position embedding (spt_backbone.py:186-231) folded into a 64-row table (pos_lut W^T + b) that either rides the
//            MMA as a one-hot second A operand, [x | onehot(cell)] [W | table^T]^T with K = C + 64 (tmae_bf16_qkv_fwd_onehot, what the
//            fused layer uses), or is added per row in the epilogue (tmae_bf16_qkv_fwd); then q and k are L2-normalised per head
//            (cosine_msa.py:151-152) -> the attention kernel reads unit vectors; 1 / max(|.|, 1e-12) per (row, head) is kept for the
//            backward of the normalisation
//   E_LN     + bias + residual, LayerNorm over the row (sst_basic_block.py:78,83; a CTA owns whole rows: N <= BN), writes the
//            pre-norm sum v (for backward), y = LN(v), mean, rstd.  Row statistics are exchanged between the two warps that
//            share a TMEM lane quarter through shared memory; v is parked in the accumulator's own TMEM columns between the
//            statistics pass and the normalise pass (tcgen05.st).
//   E_F32    fp32 output: plain store, or vector atomics when the reduction is split (weight gradients); TN takes a second B operand
//            (one-hot cell index): dy^T [x | onehot] = dW and the position-table gradient in one launch
// Every launch carries cudaLaunchAttributeProgrammaticStreamSerialization (tmae_set_option "gemm_pdl"): the kernel triggers its
// dependents at entry and waits (griddepcontrol.wait) after barrier / TMEM set-up, before its first global access.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace tmae {
namespace bfk {
}  // namespace bfk
extern int g_bf16_gemm_pdl;
extern int g_bf16_tn_plain;
namespace bfk {

typedef __nv_bfloat16 bf16;
constexpr int KB = 64;                            // bf16 elements per k-block = one 128-byte swizzle span
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;      // producer + MMA issuer + epilogue warps
constexpr int GW0 = 2 + EPI_WARPS;                // first gather warp of the GATHER variants
constexpr int GATHER_WARPS = 2;

enum Mode { M_NT = 0, M_NN = 1, M_TN = 2 };
enum Epi { E_PLAIN = 0, E_QKV = 1, E_LN = 2, E_F32 = 3 };

struct Args {
  int64_t M, N, K;
  const float* bias;
  const bf16* gsrc; const int* gtab; int gtaps, gcin;   // GATHER: operand rows come through a neighbour table (sparse conv)
  void* C; int64_t ldc;            // bf16 (E_PLAIN / E_QKV / E_LN: y) or fp32 (E_F32)
  bf16* P;                         // E_PLAIN: optional pre-activation copy (pitch ldc)
  const bf16* gelu_pre;            // E_PLAIN: optional, result *= gelu'(gelu_pre[m, n]) (pitch ldc)
  int act, accumulate;             // accumulate: C += (bf16 read-modify-write; fp32 vector atomics for E_F32)
  int dbg_plain;                   // measurement only (tmae_set_option "tn_plain_store"): split-k partials overwrite instead of adding -> WRONG results
  int pre_deriv;                   // gelu_pre already holds gelu'(pre-activation) (written by a TMAE_ACT_GELU_DERIV forward)
  int64_t k_chunk;
  // TN only: B columns at n >= n_split come from a second tensor (map_b2, columns n - n_split) and land in C2 (pitch ldc2):
  // the weight gradient dy^T [x | onehot(posidx)] yields dW and the position-table gradient in one pass over dy
  int64_t n_split; float* C2; int64_t ldc2;
  // E_QKV
  const float* table; const uint8_t* posidx; int norm_cols, hd; float* inv;
  // E_LN
  const bf16* res; const uint8_t* rowmask; const float* gamma; const float* beta; float eps; bf16* V; float* mean; float* rstd;
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// lane's 32 fp32 values -> 32 bf16 = 64 bytes at p (32-byte aligned), two 256-bit stores
__device__ __forceinline__ void store_row32(bf16* p, const float* v) {
  uint32_t u[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) u[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
  st_global_v8_u32(p, u);
  st_global_v8_u32(p + 16, u + 8);
}
__device__ __forceinline__ void load_row32(const bf16* p, float* v) {
  uint32_t u[16];
  ld_global_v8_u32(p, u);
  ld_global_v8_u32(p + 16, u + 8);
#pragma unroll
  for (int j = 0; j < 16; ++j) { const float2 f = unpack_bf16(u[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
}

// One stage of the ring: A then B.
//   K-major operand (rows x 64 bf16): ONE box {64, rows}; row r at (r / 8) * 1024 + (r % 8) * 128, 16-byte chunk c at
//   ((c ^ (r % 8)) * 16); SBO = 1024 between 8-row groups, k-step (16 elements) = +32 B.
//   MN-major operand (64 k-rows x cols): cols / 64 boxes {64, 64} of 8192 B each (LBO between boxes), k-row j of a box at
//   j * 128 (same XOR swizzle), SBO = 1024 between 8-row k groups, k-step (16 rows) = +2048 B.
template <int MODE, int BN, int STAGES, int EPI, bool GATHER, int OCC>
__global__ void __launch_bounds__(THREADS + (GATHER ? GATHER_WARPS * 32 : 0), OCC) bf16_gemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                                                  const __grid_constant__ CUtensorMap map_b,
                                                                                                  const __grid_constant__ CUtensorMap map_b2, Args g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int A_BYTES = UM * KB * 2, B_BYTES = BN * KB * 2, STAGE = A_BYTES + B_BYTES;
  __shared__ uint64_t bar_full[STAGES], bar_empty[STAGES], bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float2 ln_part[EPI == E_LN ? 2 : 1][EPI == E_LN ? UM : 1];
  __shared__ __align__(16) float ln_gb[EPI == E_LN ? 2 : 1][EPI == E_LN ? BN : 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // programmatic dependent launch: let the NEXT kernel of the stream (if it was launched with the attribute) start its set-up now ...
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int m_tiles = (int)((g.M + UM - 1) / UM), n_tiles = (int)((g.N + BN - 1) / BN);
  const int z_tiles = (int)((g.K + g.k_chunk - 1) / g.k_chunk);
  const int total = m_tiles * n_tiles * z_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { bar_init(&bar_full[s], GATHER ? 1 + GATHER_WARPS * 32 : 1); bar_init(&bar_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { bar_init(&bar_acc_full[b], 1); bar_init(&bar_acc_empty[b], EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (EPI == E_LN) {
    for (int c = threadIdx.x; c < BN; c += blockDim.x) {
      ln_gb[0][c] = c < g.N ? g.gamma[c] : 0.f;
      ln_gb[1][c] = c < g.N ? g.beta[c] : 0.f;
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  // ... and wait here, with barriers initialised and TMEM allocated, until the PREVIOUS kernel has completed and flushed (a no-op when
  // this launch did not carry the attribute).  Nothing above reads or writes memory another kernel of the step produces.
  asm volatile("griddepcontrol.wait;" ::: "memory");

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
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int m0, n0, nkb; int64_t kbeg;
        decode(t, m0, n0, kbeg, nkb);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES, use = it / STAGES;
          if (use > 0) bar_wait(&bar_empty[s], (use - 1) & 1);
          uint8_t* a = smem + s * STAGE;
          uint8_t* b = a + A_BYTES;
          const int k0 = (int)(kbeg + (int64_t)kb * KB);
          constexpr bool GA = GATHER && MODE == M_NT, GB = GATHER && MODE == M_TN;  // which operand the gather warps write
          bar_expect_tx(&bar_full[s], GA ? B_BYTES : (GB ? A_BYTES : STAGE));
          if (GA) {
            // A comes from the gather warps
          } else if (MODE == M_TN) {
#pragma unroll
            for (int j = 0; j < UM / 64; ++j) tma_load_2d(a + j * 8192, &map_a, m0 + j * 64, k0, &bar_full[s]);
          } else if (MODE == M_NT && EPI == E_QKV && k0 >= g.n_split) {
            tma_load_2d(a, &map_b2, k0 - (int)g.n_split, m0, &bar_full[s]);   // [x | onehot(cell)]: the position term rides the MMA
          } else {
            tma_load_2d(a, &map_a, k0, m0, &bar_full[s]);
          }
          if (MODE == M_NT) {
            tma_load_2d(b, &map_b, k0, n0, &bar_full[s]);
          } else if (GB) {
            // B comes from the gather warps
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) {
              const int col = n0 + j * 64;
              if (MODE == M_TN && col >= g.n_split) tma_load_2d(b + j * 8192, &map_b2, col - (int)g.n_split, k0, &bar_full[s]);
              else tma_load_2d(b + j * 8192, &map_b, col, k0, &bar_full[s]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(MODE == M_TN, MODE != M_NT, BN);
      int it = 0, i = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++i) {
        int m0, n0, nkb; int64_t kbeg;
        decode(t, m0, n0, kbeg, nkb);
        const int buf = i & 1, round = i >> 1;
        if (round > 0) bar_wait(&bar_acc_empty[buf], (round - 1) & 1);  // epilogue drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES, use = it / STAGES;
          bar_wait(&bar_full[s], use & 1);
          if (GATHER) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // cp.async (generic proxy) writes -> UMMA (async proxy) reads
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = s_u32(smem + s * STAGE), b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int kk = 0; kk < KB / 16; ++kk) {
            uint64_t da = MODE == M_TN ? desc_sw128(a_addr + kk * 2048, 8192, 1024) : desc_sw128(a_addr + kk * 32, 16, 1024);
            uint64_t db = MODE == M_NT ? desc_sw128(b_addr + kk * 32, 16, 1024) : desc_sw128(b_addr + kk * 2048, 8192, 1024);
            umma_bf16(tmem_d, da, db, idesc, (kb | kk) ? 1u : 0u);
          }
          commit_to(&bar_empty[s]);
        }
        commit_to(&bar_acc_full[buf]);
      }
    }
  } else if (GATHER && MODE == M_TN && warp >= GW0) {
    // ---------------- gather producers, weight gradient of the sparse convolution:
    //   dW[co][tap*cin + c] = sum_r dy[r][co] * src[tab[r][tap]][c]        (TN: A = dy by TMA, B gathered, reduction over r)
    // BN = 128 columns = one 256-byte piece of ONE tap (cin % 128 == 0): two boxes of 64 columns.  A thread owns one 16-byte
    // piece (8 columns) of 16 of the 64 reduction rows; a warp instruction copies 2 rows x 256 contiguous bytes.
    static_assert(!(GATHER && MODE == M_TN) || BN == 128, "gathered weight gradient: 128-column tiles");
    const int gt = (warp - GW0) * 32 + lane;          // 0 .. 63
    const int piece = gt & 15, rg = gt >> 4;          // 16-byte piece of the 256-byte row, row group 0..3 (rows rg + 4 i)
    const uint32_t poff = (uint32_t)((piece >> 3) * 8192);   // box of 64 columns
    const int pc = piece & 7;                         // 16-byte chunk inside the box row, XORed with (k % 8) below
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      int m0, n0, nkb; int64_t kbeg;
      decode(t, m0, n0, kbeg, nkb);
      const int64_t kend = kbeg + g.k_chunk < g.K ? kbeg + g.k_chunk : g.K;
      const bool col_ok = n0 + piece * 8 < g.N;
      const int tap = n0 / g.gcin, c0 = n0 - tap * g.gcin + piece * 8;
      // lane l of warp w looks up the source row of the j = l-th row this warp copies: row = 4 (j / 2) + 2 w + (j % 2)
      const int wq = warp - GW0;
      const int myrow = 4 * (lane >> 1) + 2 * wq + (lane & 1);
      int64_t r = kbeg + myrow;
      int srow_l = (r < kend) ? __ldg(g.gtab + r * g.gtaps + tap) : -1;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES, use = it / STAGES;
        const int64_t rn = r + KB;
        const int srow_n = (kb + 1 < nkb && rn < kend) ? __ldg(g.gtab + rn * g.gtaps + tap) : -1;
        if (use > 0) bar_wait(&bar_empty[s], (use - 1) & 1);
        const uint32_t bbase = s_u32(smem + s * STAGE + A_BYTES) + poff;
        const bf16* p = g.gsrc + c0;
#pragma unroll 8
        for (int i = 0; i < 16; ++i) {
          const int kr = rg + 4 * i;                                              // reduction row inside the k-block
          const int srow = __shfl_sync(0xffffffffu, srow_l, 2 * i + (lane >> 4));  // j of (rg, i) in this warp = 2 i + (rg & 1)
          const bool ok = col_ok && srow >= 0;
          cp_async16_zfill(bbase + (uint32_t)(kr * 128 + ((pc ^ (kr & 7)) * 16)), p + (int64_t)(ok ? srow : 0) * g.gcin, ok ? 16u : 0u);
        }
        cp_async_arrive_noinc(&bar_full[s]);
        r = rn; srow_l = srow_n;
      }
    }
  } else if (GATHER && warp >= GW0) {
    // ---------------- gather producers (sparse convolution forward / backward-data): A[m][tap*cin + c] = src[tab[m][tap]][c],
    // absent neighbours zero-filled.  One k-block = 64 bf16 = 128 bytes of one tap for the 128 tile rows = 1024 16-byte pieces,
    // written straight into the SWIZZLE_128B K-major layout.  Eight lanes copy one row's 128 bytes (a warp instruction reads
    // 4 rows x 128 contiguous bytes); a thread owns one chunk of 16 rows and looks their source rows up one k-block ahead.
    const int gt = (warp - GW0) * 32 + lane;          // 0 .. 63
    const int rsub = gt >> 3, c16 = gt & 7;
    constexpr int RG = UM / 8;
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
        const bf16* base = g.gsrc + c0 + c16 * 8;
#pragma unroll
        for (int j = 0; j < RG; ++j)
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
    // ---------------- epilogue warps: TMEM lane quarter q = warp % 4; this warp takes chunks ch0, ch0 + 2, ...
    const int q = warp & 3;
    constexpr int CHUNKS = BN / 32, CSTEP = EPI_WARPS / 4;
    const int ch0 = (warp - 2) >> 2;
    int i = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++i) {
      int m0, n0, nkb; int64_t kbeg;
      decode(t, m0, n0, kbeg, nkb);
      const int buf = i & 1, round = i >> 1;
      const int64_t row = m0 + q * 32 + lane;
      const bool row_ok = row < g.M;
      const uint32_t tmem_d = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
      auto release_acc = [&]() {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) bar_arrive(&bar_acc_empty[buf]);
      };

      if (EPI == E_LN) {
        // ---- bias + residual + LayerNorm over whole rows (n_tiles == 1)
        const bool use_branch = row_ok && (!g.rowmask || __ldg(g.rowmask + row) != 0);
        bar_wait(&bar_acc_full[buf], round & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float s1 = 0.f, s2 = 0.f;
        for (int ch = ch0; ch < CHUNKS; ch += CSTEP) {
          const int col0 = ch * 32;
          if (col0 >= g.N) break;
          float rv[32];
          if (row_ok) load_row32(g.res + row * g.ldc + col0, rv);   // in flight during the TMEM load
          const float bl = g.bias ? __ldg(g.bias + col0 + lane) : 0.f;
          uint32_t r[32];
          ld_tmem32(tmem_d + col0, r);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float acc = __uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bl, j);
            v[j] = row_ok ? (rv[j] + (use_branch ? acc : 0.f)) : 0.f;
            s1 += v[j];
            s2 = fmaf(v[j], v[j], s2);
            r[j] = __float_as_uint(v[j]);
          }
          st_tmem32(tmem_d + col0, r);      // park v in the accumulator's own columns for the normalise pass
          if (row_ok && g.V) store_row32(g.V + row * g.ldc + col0, v);
        }
        ln_part[ch0][q * 32 + lane] = make_float2(s1, s2);
        named_bar_sync(1 + q, 64);          // the two warps of this lane quarter
        const float2 o = ln_part[ch0 ^ 1][q * 32 + lane];
        const float inv_n = 1.f / (float)g.N;
        const float mean = (s1 + o.x) * inv_n;
        const float var = fmaxf((s2 + o.y) * inv_n - mean * mean, 0.f);
        const float rstd = rsqrtf(var + g.eps);
        if (ch0 == 0 && row_ok) { g.mean[row] = mean; g.rstd[row] = rstd; }
        for (int ch = ch0; ch < CHUNKS; ch += CSTEP) {
          const int col0 = ch * 32;
          if (col0 >= g.N) break;
          uint32_t r[32];
          ld_tmem32(tmem_d + col0, r);
          float y[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 ga = *reinterpret_cast<const float4*>(&ln_gb[0][col0 + 4 * j4]);
            const float4 be = *reinterpret_cast<const float4*>(&ln_gb[1][col0 + 4 * j4]);
            y[4 * j4 + 0] = fmaf((__uint_as_float(r[4 * j4 + 0]) - mean) * rstd, ga.x, be.x);
            y[4 * j4 + 1] = fmaf((__uint_as_float(r[4 * j4 + 1]) - mean) * rstd, ga.y, be.y);
            y[4 * j4 + 2] = fmaf((__uint_as_float(r[4 * j4 + 2]) - mean) * rstd, ga.z, be.z);
            y[4 * j4 + 3] = fmaf((__uint_as_float(r[4 * j4 + 3]) - mean) * rstd, ga.w, be.w);
          }
          if (row_ok) store_row32((bf16*)g.C + row * g.ldc + col0, y);
        }
        named_bar_sync(1 + q, 64);          // the partner has read ln_part before the next tile overwrites it
        release_acc();
        continue;
      }

      // ---- E_PLAIN / E_QKV / E_F32
      float hn[32];
      auto load_pre = [&](int ch) {
        const int col0 = n0 + ch * 32;
        if (row_ok && col0 < g.N) load_row32(g.gelu_pre + row * g.ldc + col0, hn);
      };
      if (EPI == E_PLAIN && g.gelu_pre) load_pre(ch0);
      const int pidx = (EPI == E_QKV && g.table && row_ok) ? (int)__ldg(g.posidx + row) : 0;
      bar_wait(&bar_acc_full[buf], round & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int ch = ch0; ch < CHUNKS; ch += CSTEP) {
        const int col0 = n0 + ch * 32;
        const float bl = (EPI != E_QKV && g.bias && col0 + lane < g.N) ? __ldg(g.bias + col0 + lane) : 0.f;
        float4 tb[8];
        if (EPI == E_QKV && g.table && row_ok && col0 < g.N) {
#pragma unroll
          for (int c = 0; c < 8; ++c) tb[c] = __ldg(reinterpret_cast<const float4*>(g.table + (int64_t)pidx * g.N + col0) + c);
        }
        uint32_t r[32];
        ld_tmem32(tmem_d + ch * 32, r);
        if (ch + CSTEP >= CHUNKS) release_acc();   // this warp has read everything it needs from the accumulator
        if (col0 >= g.N) continue;            // warp-uniform
        float v[32];
        if (EPI != E_QKV) {
          if (g.bias) {                       // every lane takes part in the bias broadcast
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bl, j);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          }
        }
        if (!row_ok) continue;
        if (EPI == E_QKV) {
          if (g.table) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              v[4 * c] = __uint_as_float(r[4 * c]) + tb[c].x; v[4 * c + 1] = __uint_as_float(r[4 * c + 1]) + tb[c].y;
              v[4 * c + 2] = __uint_as_float(r[4 * c + 2]) + tb[c].z; v[4 * c + 3] = __uint_as_float(r[4 * c + 3]) + tb[c].w;
            }
          } else {   // the table row came through the MMA (one-hot second A operand)
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          }
          if (col0 < g.norm_cols) {   // q / k columns: unit vectors per head (F.normalize eps 1e-12, cosine_msa.py:151-152)
            const int nh = g.norm_cols / g.hd;
            if (g.hd == 16) {
              float a = 0.f, b = 0.f;
#pragma unroll
              for (int j = 0; j < 16; ++j) { a = fmaf(v[j], v[j], a); b = fmaf(v[16 + j], v[16 + j], b); }
              const float ia = 1.f / fmaxf(sqrtf(a), 1e-12f), ib = 1.f / fmaxf(sqrtf(b), 1e-12f);
#pragma unroll
              for (int j = 0; j < 16; ++j) { v[j] *= ia; v[16 + j] *= ib; }
              *reinterpret_cast<float2*>(g.inv + row * nh + col0 / 16) = make_float2(ia, ib);
            } else {
              float a = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) a = fmaf(v[j], v[j], a);
              const float ia = 1.f / fmaxf(sqrtf(a), 1e-12f);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= ia;
              g.inv[row * nh + col0 / 32] = ia;
            }
          }
          store_row32((bf16*)g.C + row * g.ldc + col0, v);
          continue;
        }
        if (EPI == E_F32) {
          float* crow = (MODE == M_TN && col0 >= g.n_split) ? g.C2 + row * g.ldc2 + (col0 - g.n_split) : (float*)g.C + row * g.ldc + col0;
          if (g.accumulate && !g.dbg_plain) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (col0 + 4 * c < g.N) atomicAdd(reinterpret_cast<float4*>(crow + 4 * c), make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (col0 + 8 * c < g.N) st_global_v8(crow + 8 * c, v + 8 * c);
          }
          continue;
        }
        // E_PLAIN
        if (g.act == TMAE_ACT_GELU_DERIV) {   // y = gelu(v); P = gelu'(v): the backward's epilogue multiplies instead of re-deriving
          float d[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float cdf, pdf;
            gelu_parts(v[j], cdf, pdf);
            d[j] = fmaf(v[j], pdf, cdf);
            v[j] *= cdf;
          }
          if (g.P) store_row32(g.P + row * g.ldc + col0, d);
        } else if (g.P) {
          store_row32(g.P + row * g.ldc + col0, v);
        }
        if (g.gelu_pre) {
          if (g.pre_deriv) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= hn[j];
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= gelu_grad_t(hn[j]);
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
        bf16* crow = (bf16*)g.C + row * g.ldc + col0;
        if (g.accumulate) {
          float old[32];
          load_row32(crow, old);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += old[j];
        }
        store_row32(crow, v);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
}

// ------------------------------------------------------------------ host side
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

static bool make_map(CUtensorMap* m, const bf16* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_inner, uint32_t box_outer) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * sizeof(bf16)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static const char* prof_name(int mode, int epi, bool gather) {
  if (gather) return mode == M_TN ? "bf16_gemm_tn_gather" : "bf16_gemm_nt_gather";
  if (epi == E_QKV) return "bf16_gemm_qkv";
  if (epi == E_LN) return "bf16_gemm_ln";
  return mode == M_NT ? "bf16_gemm_nt" : (mode == M_NN ? "bf16_gemm_nn" : "bf16_gemm_tn");
}

struct SecondB { const bf16* B2; int64_t n2, ldb2; };   // TN: extra B columns [n_split, n_split + n2) from a second (K, n2) tensor; NT (E_QKV): extra A columns k >= n_split from a second (M, n2) tensor

// OCC = 2 (BN <= 128 only: two CTAs need 2 x 2 BN TMEM columns and 2 x STAGES x stage bytes of shared memory): two co-resident CTAs per
// SM with a 3-stage ring each -- one CTA's loads overlap the other's epilogue stores, and a launch has twice as many tiles in flight
template <int MODE, int BN, int EPI, bool GATHER = false, int OCC = 1>
static int launch(const bf16* A, const bf16* B, int64_t lda, int64_t ldb, Args g, int splits, cudaStream_t s, SecondB b2 = SecondB{nullptr, 0, 0}) {
  constexpr int STAGES = OCC == 2 ? (BN == 128 ? 3 : 4) : (BN == 256 ? 4 : (BN == 128 ? 6 : 8));
  if (g.M <= 0 || g.N <= 0) return 0;
  if (splits < 1) splits = 1;
  g.dbg_plain = g_bf16_tn_plain;
  g.k_chunk = align_up((g.K + splits - 1) / splits, KB);
  int z = (int)((g.K + g.k_chunk - 1) / g.k_chunk);
  if (z < 1) z = 1;
  if (z > 1) {
    if (EPI != E_F32) return TMAE_ERR_INVALID_ARG;   // only the fp32 atomics epilogue can split the reduction
    g.accumulate = 1;
  }
  CUtensorMap ma, mb;
  bool ok = true;
  if (GATHER && MODE == M_NT) {          // A comes from the gather warps: only B (the weights, K-major) has a tensor map
    ok &= make_map(&mb, B, g.K, g.N, ldb, KB, BN);
    ma = mb;
  } else if (GATHER && MODE == M_TN) {   // B comes from the gather warps: only A (dy, MN-major) has a tensor map
    ok &= make_map(&ma, A, g.M, g.K, lda, 64, 64);
    mb = ma;
  } else {
    // A: NT/NN K-major (rows = M, inner = K) box {64, 128} ; TN MN-major (rows = K, inner = M) box {64, 64}
    const int64_t ka = (MODE == M_NT && b2.B2) ? g.n_split : g.K;   // NT with a second A source: A covers k < n_split only
    ok &= MODE == M_TN ? make_map(&ma, A, g.M, g.K, lda, 64, 64) : make_map(&ma, A, ka, g.M, lda, KB, UM);
    // B: NT K-major (rows = N, inner = K) box {64, BN} ; NN/TN MN-major (rows = K, inner = N) box {64, 64}
    const int64_t nb = (MODE == M_TN && b2.B2) ? g.n_split : g.N;
    ok &= MODE == M_NT ? make_map(&mb, B, g.K, g.N, ldb, KB, BN) : make_map(&mb, B, nb, g.K, ldb, 64, 64);
  }
  CUtensorMap mb2 = mb;
  if (MODE == M_TN && b2.B2) ok &= make_map(&mb2, b2.B2, b2.n2, g.K, b2.ldb2, 64, 64);
  else if (MODE == M_NT && b2.B2) ok &= make_map(&mb2, b2.B2, b2.n2, g.M, b2.ldb2, KB, UM);   // second A source: (M, n2) K-major, k >= n_split
  else g.n_split = (int64_t)1 << 40;
  if (!ok) return TMAE_ERR_CUDA;
  size_t smem = (size_t)STAGES * (UM * KB * 2 + BN * KB * 2) + 1024;
  auto kern = bf16_gemm_kernel<MODE, BN, STAGES, EPI, GATHER, OCC>;
  if (smem_attr_once((const void*)kern, (int)smem)) return TMAE_ERR_CUDA;
  const double out_el = (double)g.M * g.N;
  const double a_el = (GATHER && MODE == M_NT) ? (double)g.M * g.gcin : (double)g.M * g.K;
  const double b_el = (GATHER && MODE == M_TN) ? (double)g.K * g.gcin : (double)g.N * g.K;
  double bytes = 2.0 * (a_el + b_el) + (EPI == E_F32 ? 4.0 : 2.0) * out_el;
  if (EPI == E_PLAIN) bytes += 2.0 * out_el * ((g.P ? 1 : 0) + (g.gelu_pre ? 1 : 0) + (g.accumulate ? 1 : 0));
  if (EPI == E_LN) bytes += 2.0 * out_el * (1 + (g.V ? 1 : 0));
  if (EPI == E_QKV) bytes += 4.0 * g.M * (g.norm_cols / (g.hd > 0 ? g.hd : 1));
  ProfScope prof(prof_name(MODE, EPI, GATHER), 2.0 * g.M * g.N * g.K, bytes, s);
  count_dispatch(DISP_TMA);
  int64_t tiles = (int64_t)cdiv(g.N, BN) * cdiv(g.M, UM) * z;
  dim3 grid((unsigned)(tiles < OCC * kNumSMs ? tiles : OCC * kNumSMs));
  if (g_bf16_gemm_pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(THREADS + (GATHER ? GATHER_WARPS * 32 : 0)); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, ma, mb, mb2, g) == cudaSuccess ? 0 : TMAE_ERR_CUDA;
  }
  kern<<<grid, THREADS + (GATHER ? GATHER_WARPS * 32 : 0), smem, s>>>(ma, mb, mb2, g);
  return cudaGetLastError() == cudaSuccess ? 0 : TMAE_ERR_CUDA;
}

}  // namespace bfk
int g_bf16_gemm_occ2 = 0;   // measurement switch (tmae_set_option "gemm_occ2")
int g_bf16_tn_plain = 0;
int g_bf16_gemm_pdl = 1;    // tmae_set_option "gemm_pdl": the bf16 GEMMs and the tcgen05 attention kernels are launched with programmatic stream serialization (measured: 20.46 -> 19.78 ms per step)
namespace bfk {
static int pick_bn(int64_t n) {
  // N = 384 (packed q/k/v projection at 128 channels): three full 128-wide tiles instead of a full and a half-empty 256
  if (n > 128 && !(n % 256 != 0 && n % 128 == 0 && n <= 384)) return 256;
  return n > 64 ? 128 : 64;
}

template <int MODE, int EPI>
static int dispatch_bn(const bf16* A, const bf16* B, int64_t lda, int64_t ldb, const Args& g, int splits, cudaStream_t s,
                       SecondB b2 = SecondB{nullptr, 0, 0}) {
  const int bn = pick_bn(g.N);
  if (bn == 256) return launch<MODE, 256, EPI>(A, B, lda, ldb, g, splits, s, b2);
  if (g_bf16_gemm_occ2 && EPI != E_F32) {
    if (bn == 128) return launch<MODE, 128, EPI, false, 2>(A, B, lda, ldb, g, splits, s, b2);
    return launch<MODE, 64, EPI, false, 2>(A, B, lda, ldb, g, splits, s, b2);
  }
  if (bn == 128) return launch<MODE, 128, EPI>(A, B, lda, ldb, g, splits, s, b2);
  return launch<MODE, 64, EPI>(A, B, lda, ldb, g, splits, s, b2);
}

static bool al32(const void* p) { return ((uintptr_t)p & 31) == 0; }

}  // namespace bfk
}  // namespace tmae

using namespace tmae;
using namespace tmae::bfk;

#define BF_CHECK(cond, msg) TMAE_CHECK_ARG(cond, msg)
#define BF_RUN(call, what)                                                   \
  do {                                                                       \
    int rc__ = (call);                                                       \
    if (rc__) { set_error("%s: tcgen05 launch failed", what); return rc__ < 0 ? rc__ : TMAE_ERR_CUDA; } \
  } while (0)

namespace tmae {
int bf16_linear_bwd_weight_impl(const void* dy, const void* x, float* dw, const void* onehot, float* dtab_t, int64_t m, int64_t n, int64_t k,
                                bool zero, void* stream);
}

extern "C" {

/* y = act(x w^T + bias) [+= y]; preact (nullable) receives x w^T + bias.  x (m,k), w (n,k), y (m,n): bf16; bias fp32. */
int tmae_bf16_linear_fwd(const void* x, const void* w, const float* bias, void* y, void* preact, int64_t m, int64_t n, int64_t k, int32_t act,
                         int32_t accumulate, void* stream) {
  BF_CHECK(k % 8 == 0 && n % 32 == 0, "k must be a multiple of 8 and n of 32");
  BF_CHECK(al32(x) && al32(w) && al32(y) && al32(preact), "pointers must be 32-byte aligned");
  if (m <= 0) return 0;
  Args g{};
  g.M = m; g.N = n; g.K = k; g.bias = bias; g.act = act; g.accumulate = accumulate; g.C = y; g.ldc = n; g.P = (bf16*)preact;
  BF_RUN((dispatch_bn<M_NT, E_PLAIN>((const bf16*)x, (const bf16*)w, k, k, g, 1, (cudaStream_t)stream)), "tmae_bf16_linear_fwd");
  return 0;
}

/* packed q/k/v projection: y = x w^T + table[posidx], q/k columns (the first norm_cols) L2-normalised per head of width hd;
 * inv (m, norm_cols / hd) fp32 receives 1 / max(|.|, 1e-12).  table (64, n) fp32 from tmae_pos_table. */
int tmae_bf16_qkv_fwd(const void* x, const void* w, const float* table, const uint8_t* posidx, void* y, float* inv, int64_t m, int64_t n,
                      int64_t k, int32_t norm_cols, int32_t hd, void* stream) {
  BF_CHECK(k % 8 == 0 && n % 32 == 0 && norm_cols % 32 == 0 && norm_cols <= n && (hd == 16 || hd == 32), "shape not supported");
  BF_CHECK(al32(x) && al32(w) && al32(y) && table && posidx && inv, "pointers must be 32-byte aligned and non-null");
  if (m <= 0) return 0;
  Args g{};
  g.M = m; g.N = n; g.K = k; g.C = y; g.ldc = n; g.table = table; g.posidx = posidx; g.norm_cols = norm_cols; g.hd = hd; g.inv = inv;
  BF_RUN((dispatch_bn<M_NT, E_QKV>((const bf16*)x, (const bf16*)w, k, k, g, 1, (cudaStream_t)stream)), "tmae_bf16_qkv_fwd");
  return 0;
}

/* The same projection with the position term inside the MMA: y = [x | onehot(cell)] wcat^T, wcat (n, k + 64) bf16 = [w | table^T] from
 * tmae_bf16_qkv_wcat.  The epilogue is left with the per-head normalisation (the per-row gathers of the 64-row table kept the L1 at 82 %
 * of its throughput: 51 us against 35 us for the plain kernel at m = 68k, n = 384, k = 128). */
int tmae_bf16_qkv_fwd_onehot(const void* x, const void* onehot, const void* wcat, void* y, float* inv, int64_t m, int64_t n, int64_t k,
                             int32_t norm_cols, int32_t hd, void* stream) {
  BF_CHECK(k % 64 == 0 && n % 32 == 0 && norm_cols % 32 == 0 && norm_cols <= n && (hd == 16 || hd == 32), "shape not supported");
  BF_CHECK(al32(x) && al32(onehot) && al32(wcat) && al32(y) && onehot && inv, "pointers must be 32-byte aligned and non-null");
  if (m <= 0) return 0;
  Args g{};
  g.M = m; g.N = n; g.K = k + 64; g.n_split = k; g.C = y; g.ldc = n; g.norm_cols = norm_cols; g.hd = hd; g.inv = inv;
  BF_RUN((dispatch_bn<M_NT, E_QKV>((const bf16*)x, (const bf16*)wcat, k, k + 64, g, 1, (cudaStream_t)stream, SecondB{(const bf16*)onehot, 64, 64})),
         "tmae_bf16_qkv_fwd_onehot");
  return 0;
}

namespace tmae {
namespace bfk {
// wcat[j][0 .. c) = bf16(w[j][:]) ; wcat[j][c + p] = bf16((j < n_pos ? pos_lut[p] . w[j] : 0) + bias[j]).  One warp per (j, p).
__global__ void qkv_wcat_kernel(const float* __restrict__ lut, const float* __restrict__ w, const float* __restrict__ bias, bf16* __restrict__ wcat,
                                int n, int n_pos, int c) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= 64 * n) return;
  const int j = warp >> 6, p = warp & 63;
  const float* wr = w + (int64_t)j * c;
  bf16* out = wcat + (int64_t)j * (c + 64);
  float s = 0.f;
  if (j < n_pos)
    for (int k = lane; k < c; k += 32) s = fmaf(lut[p * c + k], wr[k], s);
  s = warp_sum(s);
  if (lane == 0) out[c + p] = __float2bfloat16_rn(s + (bias ? bias[j] : 0.f));
  if (p == 0)
    for (int k = lane; k < c; k += 32) out[k] = __float2bfloat16_rn(wr[k]);
}
}  // namespace bfk
}  // namespace tmae

namespace tmae {
namespace bfk {
struct WcatSeg { const float* lut; const float* w; const float* bias; bf16* wcat; int64_t n, n_pos, c; };
// every layer's [W | table^T] in ONE launch: blockIdx.y = segment
__global__ void qkv_wcat_multi_kernel(const WcatSeg* __restrict__ segs) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: see tc_common.cuh pdl_trigger
  const WcatSeg sg = segs[blockIdx.y];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n = (int)sg.n, c = (int)sg.c;
  if (warp >= 64 * n) return;
  const int j = warp >> 6, p = warp & 63;
  const float* wr = sg.w + (int64_t)j * c;
  bf16* out = sg.wcat + (int64_t)j * (c + 64);
  float s = 0.f;
  if (j < sg.n_pos)
    for (int k = lane; k < c; k += 32) s = fmaf(sg.lut[p * c + k], wr[k], s);
  s = warp_sum(s);
  if (lane == 0) out[c + p] = __float2bfloat16_rn(s + (sg.bias ? sg.bias[j] : 0.f));
  if (p == 0)
    for (int k = lane; k < c; k += 32) out[k] = __float2bfloat16_rn(wr[k]);
}
}  // namespace bfk
}  // namespace tmae

/* segs: DEVICE array of n_seg records {const float* pos_lut; const float* w; const float* bias; void* wcat; int64 n, n_pos, c} (56 bytes):
 * tmae_bf16_qkv_wcat for every layer of a model in one launch (the operand depends on the weights only: once per optimizer step) */
int tmae_bf16_qkv_wcat_multi(const void* segs, int32_t n_seg, int32_t max_n, void* stream) {
  if (n_seg <= 0 || max_n <= 0) return 0;
  ProfScope prof("qkv_wcat_multi", 0, 0, (cudaStream_t)stream);
  dim3 grid((unsigned)cdiv((int64_t)64 * max_n * 32, 256), (unsigned)n_seg);
  qkv_wcat_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const WcatSeg*)segs);
  TMAE_CHECK_LAUNCH();
  return 0;
}

/* wcat (n, c + 64) bf16 = [w | table^T] of the packed projection: table[p][j] = (j < n_pos ? pos_lut[p] . w[j] : 0) + bias[j]
 * (tmae_pos_table), from the fp32 master weights */
int tmae_bf16_qkv_wcat(const float* pos_lut, const float* w, const float* bias, void* wcat, int32_t n, int32_t n_pos, int32_t c, void* stream) {
  if (n <= 0) return 0;
  BF_CHECK(c % 64 == 0 && al32(wcat), "c must be a multiple of 64, wcat 32-byte aligned");
  ProfScope prof("qkv_wcat", 0, 0, (cudaStream_t)stream);
  qkv_wcat_kernel<<<cdiv((int64_t)64 * n * 32, 256), 256, 0, (cudaStream_t)stream>>>(pos_lut, w, bias, (bf16*)wcat, n, n_pos, c);
  TMAE_CHECK_LAUNCH();
  return 0;
}

/* v = res + (rowmask[row] ? a w^T + bias : 0) ; y = LayerNorm(v) * gamma + beta.  n in {128, 256} (a CTA owns whole rows). */
int tmae_bf16_linear_ln_fwd(const void* a, const void* w, const float* bias, const void* res, const uint8_t* rowmask, const float* gamma,
                            const float* beta, float eps, void* v, void* y, float* mean, float* rstd, int64_t m, int64_t n, int64_t k,
                            void* stream) {
  BF_CHECK(k % 8 == 0 && (n == 128 || n == 256), "n must be 128 or 256, k a multiple of 8");
  BF_CHECK(al32(a) && al32(w) && al32(res) && al32(v) && al32(y) && gamma && beta && mean && rstd, "pointers must be 32-byte aligned and non-null");
  if (m <= 0) return 0;
  Args g{};
  g.M = m; g.N = n; g.K = k; g.bias = bias; g.C = y; g.ldc = n; g.res = (const bf16*)res; g.rowmask = rowmask; g.gamma = gamma; g.beta = beta;
  g.eps = eps; g.V = (bf16*)v; g.mean = mean; g.rstd = rstd;
  if (n == 128 && g_bf16_gemm_occ2) BF_RUN((launch<M_NT, 128, E_LN, false, 2>((const bf16*)a, (const bf16*)w, k, k, g, 1, (cudaStream_t)stream)), "tmae_bf16_linear_ln_fwd");
  else if (n == 128) BF_RUN((launch<M_NT, 128, E_LN>((const bf16*)a, (const bf16*)w, k, k, g, 1, (cudaStream_t)stream)), "tmae_bf16_linear_ln_fwd");
  else BF_RUN((launch<M_NT, 256, E_LN>((const bf16*)a, (const bf16*)w, k, k, g, 1, (cudaStream_t)stream)), "tmae_bf16_linear_ln_fwd");
  return 0;
}

/* dx = dy w  [* gelu'(preact)] [+= dx].  dy (m,n), w (n,k), dx / preact (m,k): bf16. */
int tmae_bf16_linear_bwd_data(const void* dy, const void* w, const void* gelu_pre, void* dx, int64_t m, int64_t n, int64_t k, int32_t accumulate,
                              void* stream) {
  BF_CHECK(n % 8 == 0 && k % 32 == 0, "n must be a multiple of 8 and k of 32");
  BF_CHECK(al32(dy) && al32(w) && al32(dx) && al32(gelu_pre), "pointers must be 32-byte aligned");
  if (m <= 0) return 0;
  Args g{};
  g.M = m; g.N = k; g.K = n; g.accumulate = accumulate & TMAE_BWD_ACCUMULATE; g.pre_deriv = (accumulate & TMAE_BWD_PRE_IS_DERIVATIVE) ? 1 : 0;
  g.gelu_pre = (const bf16*)gelu_pre; g.C = dx; g.ldc = k;
  BF_RUN((dispatch_bn<M_NN, E_PLAIN>((const bf16*)dy, (const bf16*)w, n, k, g, 1, (cudaStream_t)stream)), "tmae_bf16_linear_bwd_data");
  return 0;
}

/* dw (n,k) fp32 = dy^T x  (overwrites dw).  dy (m,n), x (m,k): bf16.
 * onehot (m, 64) bf16, nullable: also dtab_t (n, 64) fp32 = dy^T onehot (overwritten) from the same pass over dy -- the binned column
 * sums behind the position-table gradient (tmae_pos_table_bwd, transposed form); needs k % 64 == 0. */
int tmae_bf16_linear_bwd_weight(const void* dy, const void* x, float* dw, const void* onehot, float* dtab_t, int64_t m, int64_t n, int64_t k,
                                void* stream) {
  return tmae::bf16_linear_bwd_weight_impl(dy, x, dw, onehot, dtab_t, m, n, k, true, stream);
}

}  // extern "C"

namespace tmae {
// zero = false: dw (and dtab_t) were zero-filled by the caller (the split reduction adds into them)
int bf16_linear_bwd_weight_impl(const void* dy, const void* x, float* dw, const void* onehot, float* dtab_t, int64_t m, int64_t n, int64_t k,
                                bool zero, void* stream) {
  BF_CHECK(n % 8 == 0 && k % 8 == 0, "n and k must be multiples of 8");
  BF_CHECK(al32(dy) && al32(x) && al32(dw) && al32(onehot) && al32(dtab_t), "pointers must be 32-byte aligned");
  BF_CHECK(!onehot || (dtab_t && k % 64 == 0), "the one-hot form needs dtab_t and k % 64 == 0");
  cudaStream_t s = (cudaStream_t)stream;
  if (zero) {
    TMAE_CUDA(cudaMemsetAsync(dw, 0, (size_t)n * k * sizeof(float), s));
    if (onehot) TMAE_CUDA(cudaMemsetAsync(dtab_t, 0, (size_t)n * 64 * sizeof(float), s));
  }
  if (m <= 0) return 0;
  Args g{};
  g.M = n; g.N = onehot ? k + 64 : k; g.K = m; g.accumulate = 1; g.C = dw; g.ldc = k;
  g.n_split = k; g.C2 = dtab_t; g.ldc2 = 64;
  int64_t tiles = (int64_t)cdiv(n, UM) * cdiv(g.N, pick_bn(g.N));
  // split the row reduction so that tiles x splits fills ONE wave of the persistent grid (floor: a second partial wave doubles the makespan)
  int64_t want = kNumSMs / tiles, maxs = (m + 4 * KB - 1) / (4 * KB);
  if (want > maxs) want = maxs;
  BF_RUN((dispatch_bn<M_TN, E_F32>((const bf16*)dy, (const bf16*)x, n, k, g, (int)(want < 1 ? 1 : want), s,
                                   SecondB{(const bf16*)onehot, 64, 64})), "tmae_bf16_linear_bwd_weight");
  return 0;
}
}  // namespace tmae

extern "C" {

/* sparse convolution forward / backward-data as a gathered NT GEMM: y[o, :] = sum_tap x[table[o, tap], :] w[:, tap, :]^T
 * x (rows_in, cin), w (cout, taps, cin), y (rows_out, cout): bf16 */
int tmae_bf16_sparse_conv_fwd(const void* x, const int32_t* table, const void* w, void* y, int64_t rows_out, int32_t taps, int32_t cin,
                              int32_t cout, void* stream) {
  BF_CHECK(cin % KB == 0 && cout % 32 == 0, "cin must be a multiple of 64 and cout of 32");
  BF_CHECK(al32(x) && al32(w) && al32(y), "pointers must be 32-byte aligned");
  if (rows_out <= 0) return 0;
  Args g{};
  g.M = rows_out; g.N = cout; g.K = (int64_t)taps * cin; g.C = y; g.ldc = cout;
  g.gsrc = (const bf16*)x; g.gtab = table; g.gtaps = taps; g.gcin = cin;
  cudaStream_t s = (cudaStream_t)stream;
  const bf16* wb = (const bf16*)w;
  if (cout > 128) BF_RUN((launch<M_NT, 256, E_PLAIN, true>(wb, wb, g.K, g.K, g, 1, s)), "tmae_bf16_sparse_conv_fwd");
  else if (cout > 64) BF_RUN((launch<M_NT, 128, E_PLAIN, true>(wb, wb, g.K, g.K, g, 1, s)), "tmae_bf16_sparse_conv_fwd");
  else BF_RUN((launch<M_NT, 64, E_PLAIN, true>(wb, wb, g.K, g.K, g, 1, s)), "tmae_bf16_sparse_conv_fwd");
  return 0;
}

/* dw (cout, taps*cin) fp32 = sum_o dy[o, :]^T x[table[o, tap], :]  (overwrites dw): gathered TN GEMM, split over the output rows */
int tmae_bf16_sparse_conv_bwd_weight(const void* dy, const void* x, const int32_t* table, float* dw, int64_t rows_out, int32_t taps, int32_t cin,
                                     int32_t cout, void* stream) {
  BF_CHECK(cin % 128 == 0 && cout % 8 == 0, "cin must be a multiple of 128 and cout of 8");
  BF_CHECK(al32(dy) && al32(x) && al32(dw), "pointers must be 32-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t kk = (int64_t)taps * cin;
  TMAE_CUDA(cudaMemsetAsync(dw, 0, (size_t)cout * kk * sizeof(float), s));
  if (rows_out <= 0) return 0;
  Args g{};
  g.M = cout; g.N = kk; g.K = rows_out; g.accumulate = 1; g.C = dw; g.ldc = kk;
  g.gsrc = (const bf16*)x; g.gtab = table; g.gtaps = taps; g.gcin = cin;
  int64_t tiles = (int64_t)cdiv(cout, UM) * cdiv(kk, 128);
  int64_t want = kNumSMs / tiles, maxs = (rows_out + 4 * KB - 1) / (4 * KB);
  if (want > maxs) want = maxs;
  BF_RUN((launch<M_TN, 128, E_F32, true>((const bf16*)dy, (const bf16*)dy, cout, kk, g, (int)(want < 1 ? 1 : want), s)), "tmae_bf16_sparse_conv_bwd_weight");
  return 0;
}

}  // extern "C"
