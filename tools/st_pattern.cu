// Write-path micro-benchmark for the GEMM epilogue (tools/st_pattern.cu; nvcc -arch=sm_100a -O3 -o /tmp/st_pattern tools/st_pattern.cu).
// 148 persistent CTAs write a (M x N) fp32 matrix tile by tile (128 rows x N columns per tile, like the TMA GEMM's
// epilogue after TMEM -> registers), with different store shapes and warp counts.  Prints GB/s per variant: which store
// shape / how many warps the write path of one SM needs to reach the HBM write rate.
//   P0  lane = row, 8 x 128-bit stores per 128-byte line          (the epilogue before round 1's last change)
//   P1  lane = row, 4 x 256-bit stores per line                   (current epilogue)
//   P2  warp-coalesced: 32 lanes x 16 B = 4 full lines of ONE row per instruction (needs a transpose in the real kernel)
//   P3  8 lanes per line: 4 rows x 128 B per instruction          (quarter-warp per row)
//   P4  shared-memory staged 1-D bulk stores (cp.async.bulk.global.shared::cta), 512 B per row, one elected lane
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void st_v8(float* p, float v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int P>
__global__ void __launch_bounds__(1024, 1) st_kernel(float* __restrict__ C, int64_t M, int N, int tiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int q = warp & 3, cgroup = warp >> 2, cgroups = nwarps >> 2;   // row quarter, column group
  const int chunks = N / 32;                                          // 32-column chunks per row
  const float val = (float)threadIdx.x;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t row0 = (int64_t)t * 128 + q * 32;
    for (int ch = cgroup; ch < chunks; ch += cgroups) {
      float* base = C + row0 * N + ch * 32;    // 32 rows x 32 columns handled by this warp
      if (P == 0) {
        float* p = base + (int64_t)lane * N;
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<float4*>(p + 4 * c) = make_float4(val, val, val, val);
      } else if (P == 1) {
        float* p = base + (int64_t)lane * N;
#pragma unroll
        for (int c = 0; c < 4; ++c) st_v8(p + 8 * c, val);
      } else if (P == 3) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
          *reinterpret_cast<float4*>(base + (int64_t)(r * 4 + (lane >> 3)) * N + (lane & 7) * 4) = make_float4(val, val, val, val);
      } else if (P == 4) {
        // stage 32 rows x 128 B (4 KB) in this warp's buffer (double-buffered), then 32 bulk copies of 128 B by lane 0..31
        uint8_t* buf = smem + (warp * 2 + (ch / cgroups & 1)) * 4096;
        if (ch / cgroups >= 2) { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<float4*>(buf + lane * 128 + ((c ^ (lane & 7)) * 16)) = make_float4(val, val, val, val);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(base + (int64_t)lane * N), "r"(s_u32(buf + lane * 128)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (P == 2) {
      // one warp writes whole rows: rows q*32 .. +32 of the tile, all N columns, 512 B per instruction
      for (int r = cgroup; r < 32; r += cgroups) {
        float* p = C + (row0 + r) * N;
        for (int c = lane * 4; c < N; c += 128) *reinterpret_cast<float4*>(p + c) = make_float4(val, val, val, val);
      }
    }
  }
  if (P == 4) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int P>
static void run(const char* name, float* C, int64_t M, int N, int threads, float* flush, size_t flush_bytes) {
  const int tiles = (int)(M / 128);
  size_t smem = P == 4 ? (size_t)(threads / 32) * 2 * 4096 : 0;
  cudaFuncSetAttribute(st_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 6; ++it) {
    cudaMemsetAsync(flush, it, flush_bytes);
    cudaEventRecord(e0);
    st_kernel<P><<<148, threads, smem>>>(C, M, N, tiles);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it > 0 && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  printf("%-28s N=%4d threads=%4d  %7.1f us  %7.0f GB/s %s\n", name, N, threads, best * 1e3, (double)M * N * 4 / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  const int64_t M = 56064;  // 438 tiles of 128 rows
  float *C, *flush;
  const size_t flush_bytes = 256u << 20;
  cudaMalloc(&C, (size_t)M * 768 * 4);
  cudaMalloc(&flush, flush_bytes);
  for (int N : {128, 384, 768}) {
    for (int threads : {128, 256, 512, 1024}) {
      run<0>("P0 lane=row 8x128b", C, M, N, threads, flush, flush_bytes);
      run<1>("P1 lane=row 4x256b", C, M, N, threads, flush, flush_bytes);
      run<3>("P3 8 lanes per line", C, M, N, threads, flush, flush_bytes);
      run<2>("P2 warp-coalesced rows", C, M, N, threads, flush, flush_bytes);
      if (threads <= 512) run<4>("P4 smem + 128B bulk stores", C, M, N, threads, flush, flush_bytes);
    }
  }
  return 0;
}
