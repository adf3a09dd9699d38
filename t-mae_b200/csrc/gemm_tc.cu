// bf16 tensor-core GEMM family for sm_100a: tcgen05.mma with the accumulator in TMEM (TMAE_PREC_BF16).
//
// Same contractions as gemm.cu (linear forward / backward-data / backward-weight and the gather-GEMM form of
// the sparse convolutions), with fp32 activations and weights in HBM converted to bf16 while they are staged
// into shared memory, fp32 accumulation in tensor memory, and the bias / activation / residual epilogue
// applied on the way out of TMEM.  The A operand is staged by the CTA's threads (not TMA) because its rows
// are converted fp32 -> bf16 and, for the sparse convolutions, gathered through the neighbour table.
//
// Shared-memory operand layout: the no-swizzle canonical UMMA layouts (8 x 16-byte core matrices):
//   K-major  (reduction index contiguous in HBM):  core(mn_group, k_chunk)  8 rows x 8 elements, row = 16 B
//   MN-major (output index contiguous in HBM):     core(mn_group, k_group)  8 k-rows x 8 mn elements
// both stored as [k][mn_group][128 B], so SBO (MN direction) = 128 B and LBO (K direction) = (tile_mn/8)*128 B.
// One tcgen05.mma consumes K = 16 = two cores along K.
//
// Pipeline: STAGES-deep ring of {A,B} stages.  All 256 threads stage k-block kb (global -> regs -> bf16 ->
// st.shared), fence.proxy.async + __syncthreads, then one thread issues BK/16 MMAs and a tcgen05.commit onto the
// stage's mbarrier; staging of the next k-block overlaps the asynchronous MMAs.  Every mbarrier wait is bounded
// and traps instead of hanging.
#include <cuda_bf16.h>

#include "common.cuh"

namespace tmae {

constexpr int TBM = 128;     // UMMA M
constexpr int TBK = 64;      // k-block per stage
constexpr int TC_THREADS = 256;
constexpr int TSTAGES = 2;   // two smem stages + one register-resident k-block in flight

enum TcMode { TC_NT = 0, TC_NN = 1, TC_TN = 2 };

struct TcArgs {
  const float* A; const float* B; float* C;
  int64_t M, N, K;            // output M x N, reduction K
  int64_t lda, ldb, ldc;
  const float* bias; const float* residual; float* preact;
  const int* tab; int taps, cin;   // gather (A rows for NT, B rows for TN)
  int act, accumulate, atomic;
  int64_t k_chunk;            // reduction range per blockIdx.z
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();  // protocol bug: fail loudly, never hang the GPU
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// no-swizzle UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start >> 4, LBO >> 4 at bit 16,
// SBO >> 4 at bit 32, version 1 at bit 46, layout type 0 (SWIZZLE_NONE) at bit 61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, majors, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t umma_idesc(int a_mn_major, int b_mn_major, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(TBM >> 4) << 24);
}

__device__ __forceinline__ uint4 pack8(const float4& a, const float4& b) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
  uint4 r;
  r.x = *reinterpret_cast<uint32_t*>(&p0); r.y = *reinterpret_cast<uint32_t*>(&p1);
  r.z = *reinterpret_cast<uint32_t*>(&p2); r.w = *reinterpret_cast<uint32_t*>(&p3);
  return r;
}

// 32-byte (8 x fp32) global load: one full sector per lane (LDG.E.256 on sm_100)
__device__ __forceinline__ void ldg8(const float* p, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ uint4 pack8v(const float* v) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
  uint4 r;
  r.x = *reinterpret_cast<uint32_t*>(&p0); r.y = *reinterpret_cast<uint32_t*>(&p1);
  r.z = *reinterpret_cast<uint32_t*>(&p2); r.w = *reinterpret_cast<uint32_t*>(&p3);
  return r;
}

// A (ROWS x 64) operand tile is moved in two steps so that the global loads of k-block kb+1 are in flight while
// k-block kb is being consumed: load_* fills a register fragment (one 32-byte sector per lane per chunk),
// store_* converts to bf16 and writes the canonical no-swizzle UMMA layout.
//
// K-major (reduction index contiguous in HBM): element (r, k) = src[row(r) * ld + k0 + k].  8 consecutive lanes take
// 8 consecutive rows of one 16-byte chunk (conflict-free 128 B shared stores); a warp instruction reads 128
// contiguous bytes of each of 8 rows.
template <int ROWS>
struct FragK { float v[ROWS / 32][8]; };

template <int ROWS, bool GATHER>
__device__ __forceinline__ void load_kmajor(FragK<ROWS>& f, const float* __restrict__ src, int64_t ld, int64_t row0, int64_t row_end,
                                            int64_t k0, int64_t k_end, const int* __restrict__ tab, int taps, int cin) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r8 = lane & 7, kcl = lane >> 3;
#pragma unroll
  for (int i = 0; i < ROWS / 32; ++i) {
    int c = warp + i * (TC_THREADS / 32);
    int rg = c >> 1, kc = (c & 1) * 4 + kcl;
    int64_t row = row0 + rg * 8 + r8, k = k0 + kc * 8;
    const float* p = nullptr;
    if (row < row_end && k < k_end) {
      if (GATHER) {
        int tap = (int)(k / cin);
        int srow = tab[row * taps + tap];
        if (srow >= 0) p = src + (int64_t)srow * ld + (k - (int64_t)tap * cin);
      } else {
        p = src + row * ld + k;
      }
    }
    if (p) ldg8(p, f.v[i]);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f.v[i][j] = 0.f;
    }
  }
}
template <int ROWS>
__device__ __forceinline__ void store_kmajor(uint8_t* dst, const FragK<ROWS>& f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r8 = lane & 7, kcl = lane >> 3;
#pragma unroll
  for (int i = 0; i < ROWS / 32; ++i) {
    int c = warp + i * (TC_THREADS / 32);
    int rg = c >> 1, kc = (c & 1) * 4 + kcl;
    *reinterpret_cast<uint4*>(dst + ((kc * (ROWS / 8) + rg) * 128 + r8 * 16)) = pack8v(f.v[i]);
  }
}

// MN-major (output index contiguous in HBM): element (k, c) = src[row(k) * ld + c0 + c].  8 consecutive lanes take the
// 8 k-rows of one core matrix; a warp covers 4 adjacent column groups = 128 contiguous bytes of each of 8 source rows.
template <int COLS, bool GATHER>
__device__ __forceinline__ void load_mnmajor(FragK<COLS>& f, const float* __restrict__ src, int64_t ld, int64_t k0, int64_t k_end,
                                             int64_t c0, int64_t c_end, const int* __restrict__ tab, int taps, int cin) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k8 = lane & 7, cgl = lane >> 3;
#pragma unroll
  for (int i = 0; i < COLS / 32; ++i) {
    int c = warp + i * (TC_THREADS / 32);
    int kg = c / (COLS / 32), cg = (c % (COLS / 32)) * 4 + cgl;
    int64_t k = k0 + kg * 8 + k8, col = c0 + cg * 8;
    const float* p = nullptr;
    if (k < k_end && col < c_end) {
      if (GATHER) {
        int tap = (int)(col / cin);
        int srow = tab[k * taps + tap];
        if (srow >= 0) p = src + (int64_t)srow * ld + (col - (int64_t)tap * cin);
      } else {
        p = src + k * ld + col;
      }
    }
    if (p) ldg8(p, f.v[i]);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f.v[i][j] = 0.f;
    }
  }
}
template <int COLS>
__device__ __forceinline__ void store_mnmajor(uint8_t* dst, const FragK<COLS>& f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k8 = lane & 7, cgl = lane >> 3;
#pragma unroll
  for (int i = 0; i < COLS / 32; ++i) {
    int c = warp + i * (TC_THREADS / 32);
    int kg = c / (COLS / 32), cg = (c % (COLS / 32)) * 4 + cgl;
    *reinterpret_cast<uint4*>(dst + ((kg * (COLS / 8) + cg) * 128 + k8 * 16)) = pack8v(f.v[i]);
  }
}

__device__ __forceinline__ float gelu_erf_tc(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

// MODE NT: C[m,n] = sum_k A[m,k] W[n,k]      A K-major (optionally gathered), B K-major
// MODE NN: C[m,n] = sum_k A[m,k] B[k,n]      A K-major, B MN-major
// MODE TN: C[m,n] = sum_k A[k,m] B[k,n]      A MN-major, B MN-major (optionally gathered), split over k with atomics
template <int MODE, int BN, bool GATHER>
__global__ void __launch_bounds__(TC_THREADS) tc_gemm_kernel(TcArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int A_BYTES = TBM * TBK * 2, B_BYTES = BN * TBK * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + TSTAGES * A_BYTES;
  __shared__ uint64_t bar_free[TSTAGES];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.y * TBM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_chunk;
  const int64_t kend = kbeg + g.k_chunk < g.K ? kbeg + g.k_chunk : g.K;
  const int nkb = (int)((kend - kbeg + TBK - 1) / TBK);

  FragK<TBM> fa;
  FragK<BN> fb;
  auto prefetch = [&](int kb) {
    const int64_t k0 = kbeg + (int64_t)kb * TBK;
    if (MODE == TC_TN) load_mnmajor<TBM, false>(fa, g.A, g.lda, k0, kend, m0, g.M, nullptr, 0, 1);
    else load_kmajor<TBM, GATHER && MODE == TC_NT>(fa, g.A, g.lda, m0, g.M, k0, kend, g.tab, g.taps, g.cin);
    if (MODE == TC_NT) load_kmajor<BN, false>(fb, g.B, g.ldb, n0, g.N, k0, kend, nullptr, 0, 1);
    else load_mnmajor<BN, GATHER && MODE == TC_TN>(fb, g.B, g.ldb, k0, kend, n0, g.N, g.tab, g.taps, g.cin);
  };
  if (nkb > 0) prefetch(0);  // in flight during barrier init / TMEM allocation

  if (threadIdx.x == 0) {
    for (int s = 0; s < TSTAGES; ++s) mbar_init(&bar_free[s], 1);
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, BN < 32 ? 32 : BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  const uint32_t idesc = umma_idesc(MODE == TC_TN, MODE != TC_NT, BN);
  constexpr uint32_t A_LBO = (TBM / 8) * 128, B_LBO = (BN / 8) * 128, SBO = 128;

  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb % TSTAGES, use = kb / TSTAGES;
    if (use > 0) mbar_wait(&bar_free[s], (use - 1) & 1);  // MMAs that read this stage have retired
    uint8_t* a = sA + s * A_BYTES;
    uint8_t* b = sB + s * B_BYTES;
    if (MODE == TC_TN) store_mnmajor<TBM>(a, fa); else store_kmajor<TBM>(a, fa);
    if (MODE == TC_NT) store_kmajor<BN>(b, fb); else store_mnmajor<BN>(b, fb);
    if (kb + 1 < nkb) prefetch(kb + 1);  // overlaps the barrier, the MMA issue and the MMAs themselves
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(a), b_addr = smem_u32(b);
#pragma unroll
      for (int kk = 0; kk < TBK / 16; ++kk) {
        uint64_t da = umma_desc(a_addr + kk * 2 * A_LBO, A_LBO, SBO);
        uint64_t db = umma_desc(b_addr + kk * 2 * B_LBO, B_LBO, SBO);
        umma_bf16(tmem_d, da, db, idesc, (kb | kk) ? 1u : 0u);
      }
      umma_commit(&bar_free[s]);
      if (kb == nkb - 1) umma_commit(&bar_done);
    }
  }
  if (nkb > 0) mbar_wait(&bar_done, 0);
  tc_fence_after();

  // ---- epilogue.  Warp w owns TMEM lanes 32*(w%4).. and the 32-column chunks ch = w/4, w/4+2, ...  Each chunk goes
  // TMEM -> registers (lane = row) -> padded shared tile -> registers (lane = column), so that every global access
  // of the epilogue (C, residual, pre-activation) is a 128-byte row segment.  All MMAs have retired: operand
  // shared memory is reused for the transposition tiles.
  float* tile = reinterpret_cast<float*>(smem) + warp * (32 * 33);
  const int64_t rbase = m0 + (warp & 3) * 32;
  constexpr int CHUNKS = BN / 32;
  for (int ch = (warp >> 2); ch < CHUNKS; ch += 2) {
    uint32_t r[32];
    if (nkb > 0) {
      tmem_ld32(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + ch * 32, r);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = 0u;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = __uint_as_float(r[j]);
    __syncwarp();
    const int64_t col = n0 + ch * 32 + lane;
    if (col < g.N) {
      const float bias = g.bias ? g.bias[col] : 0.f;
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr) {
        const int64_t row = rbase + rr;
        if (row >= g.M) break;
        float v = tile[rr * 33 + lane];
        const int64_t o = row * g.ldc + col;
        if (g.atomic) { atomicAdd(g.C + o, v); continue; }
        v += bias;
        if (g.preact) g.preact[o] = v;
        if (g.act == TMAE_ACT_GELU) v = gelu_erf_tc(v);
        else if (g.act == TMAE_ACT_RELU) v = fmaxf(v, 0.f);
        if (g.residual) v += g.residual[o];
        if (g.accumulate) v += g.C[o];
        g.C[o] = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, BN < 32 ? 32 : BN);
}

template <int MODE, int BN, bool GATHER>
static int tc_launch(TcArgs& g, int splits, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return 0;
  if (splits < 1) splits = 1;
  g.k_chunk = align_up((g.K + splits - 1) / splits, TBK);
  int z = (int)((g.K + g.k_chunk - 1) / g.k_chunk);
  if (z < 1) z = 1;
  g.atomic = z > 1 || g.atomic;
  size_t smem = (size_t)TSTAGES * (TBM * TBK * 2 + BN * TBK * 2);
  auto kern = tc_gemm_kernel<MODE, BN, GATHER>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TMAE_ERR_CUDA;
    attr_set = true;
  }
  dim3 grid((unsigned)cdiv(g.N, BN), (unsigned)cdiv(g.M, TBM), (unsigned)z);
  static const char* names[3][2] = {{"tc_gemm_nt", "tc_gemm_nt_gather"}, {"tc_gemm_nn", "tc_gemm_nn"}, {"tc_gemm_tn", "tc_gemm_tn_gather"}};
  // algorithmic traffic: every operand element once (gathered rows count once, not once per tap), C once (twice if read-modify-write)
  double a_el = GATHER && MODE == TC_NT ? (double)g.M * g.cin : (double)g.M * g.K;
  double b_el = GATHER && MODE == TC_TN ? (double)g.K * g.cin : (double)g.N * g.K;
  double c_el = (double)g.M * g.N * (1.0 + (g.accumulate || g.residual ? 1.0 : 0.0) + (g.preact ? 1.0 : 0.0));
  ProfScope prof(names[MODE][GATHER ? 1 : 0], 2.0 * g.M * g.N * g.K, 4.0 * (a_el + b_el + c_el), s);
  kern<<<grid, TC_THREADS, smem, s>>>(g);
  return cudaGetLastError() == cudaSuccess ? 0 : TMAE_ERR_CUDA;
}

template <int MODE, bool GATHER>
static int tc_dispatch(TcArgs& g, int splits, cudaStream_t s) {
  if (g.N > 128) return tc_launch<MODE, 256, GATHER>(g, splits, s);
  if (g.N > 64) return tc_launch<MODE, 128, GATHER>(g, splits, s);
  return tc_launch<MODE, 64, GATHER>(g, splits, s);
}

// ---- entry points used by gemm.cu's ABI functions when precision == TMAE_PREC_BF16
bool tc_linear_fwd_ok(int64_t m, int64_t n, int64_t k) { return k % 8 == 0; }  // k tails are zero-filled per 16-byte chunk
int tc_linear_fwd(const float* x, const float* w, const float* bias, const float* residual, float* y, float* preact, int64_t m,
                  int64_t n, int64_t k, int act, cudaStream_t s) {
  TcArgs g{};
  g.A = x; g.B = w; g.C = y; g.M = m; g.N = n; g.K = k; g.lda = k; g.ldb = k; g.ldc = n;
  g.bias = bias; g.residual = residual; g.preact = preact; g.act = act;
  return tc_dispatch<TC_NT, false>(g, 1, s);
}

bool tc_linear_bwd_data_ok(int64_t m, int64_t n, int64_t k) { return n % 8 == 0 && k % 8 == 0; }
int tc_linear_bwd_data(const float* dy, const float* w, float* dx, int64_t m, int64_t n, int64_t k, int accumulate, cudaStream_t s) {
  TcArgs g{};
  g.A = dy; g.B = w; g.C = dx; g.M = m; g.N = k; g.K = n; g.lda = n; g.ldb = k; g.ldc = k; g.accumulate = accumulate;
  return tc_dispatch<TC_NN, false>(g, 1, s);
}

static int tn_splits(int64_t out_tiles, int64_t k) {
  int64_t want = (kNumSMs + out_tiles - 1) / out_tiles;
  int64_t maxs = (k + 4 * TBK - 1) / (4 * TBK);
  if (want > maxs) want = maxs;
  return (int)(want < 1 ? 1 : want);
}

bool tc_linear_bwd_weight_ok(int64_t m, int64_t n, int64_t k) { return n % 8 == 0 && k % 8 == 0 && m >= 1; }
// dw must be zero-filled by the caller
int tc_linear_bwd_weight(const float* dy, const float* x, float* dw, int64_t m, int64_t n, int64_t k, cudaStream_t s) {
  TcArgs g{};
  g.A = dy; g.B = x; g.C = dw; g.M = n; g.N = k; g.K = m; g.lda = n; g.ldb = k; g.ldc = k; g.atomic = 1;
  int bn = k > 128 ? 256 : (k > 64 ? 128 : 64);
  return tc_dispatch<TC_TN, false>(g, tn_splits((int64_t)cdiv(n, TBM) * cdiv(k, bn), m), s);
}

bool tc_sparse_conv_ok(int cin, int cout) { return cin % 8 == 0 && cout % 8 == 0; }
int tc_sparse_conv_fwd(const float* x, const int* table, const float* w, float* y, int64_t rows_out, int taps, int cin, int cout,
                       int accumulate, cudaStream_t s) {
  TcArgs g{};
  g.A = x; g.B = w; g.C = y; g.M = rows_out; g.N = cout; g.K = (int64_t)taps * cin; g.lda = cin; g.ldb = g.K; g.ldc = cout;
  g.tab = table; g.taps = taps; g.cin = cin; g.accumulate = accumulate;
  return tc_dispatch<TC_NT, true>(g, 1, s);
}
// dw (cout, taps*cin) must be zero-filled by the caller
int tc_sparse_conv_bwd_weight(const float* dy, const float* x, const int* table, float* dw, int64_t rows_out, int taps, int cin,
                              int cout, cudaStream_t s) {
  TcArgs g{};
  int64_t kk = (int64_t)taps * cin;
  g.A = dy; g.B = x; g.C = dw; g.M = cout; g.N = kk; g.K = rows_out; g.lda = cout; g.ldb = cin; g.ldc = kk;
  g.tab = table; g.taps = taps; g.cin = cin; g.atomic = 1;
  return tc_dispatch<TC_TN, true>(g, tn_splits((int64_t)cdiv(cout, TBM) * cdiv(kk, 256), rows_out), s);
}

}  // namespace tmae
