// Shared device/host helpers for libtmae_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tmae_sm100.h"

namespace tmae {

void set_error(const char* fmt, ...);

#define TMAE_CHECK_ARG(cond, msg)                                   \
  do {                                                              \
    if (!(cond)) {                                                  \
      tmae::set_error("%s: %s", __func__, msg);                     \
      return TMAE_ERR_INVALID_ARG;                                  \
    }                                                               \
  } while (0)

#define TMAE_CHECK_LAUNCH()                                                         \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      tmae::set_error("%s: CUDA launch failed: %s", __func__, cudaGetErrorString(e__)); \
      return TMAE_ERR_CUDA;                                                         \
    }                                                                               \
  } while (0)

#define TMAE_CUDA(call)                                                             \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      tmae::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__)); \
      return TMAE_ERR_CUDA;                                                         \
    }                                                                               \
  } while (0)

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Bump allocator over the caller-provided workspace (256-byte aligned slices).
struct Workspace {
  char* base;
  size_t cap, off;
  Workspace(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
  template <typename T>
  T* take(int64_t n) {
    size_t bytes = (size_t)align_up(n * (int64_t)sizeof(T), 256);
    if (off + bytes > cap) return nullptr;
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
};
static inline size_t ws_bytes(int64_t n, size_t elem) { return (size_t)align_up(n * (int64_t)elem, 256); }

// Exclusive prefix sum of int32 (n <= 2^31), deterministic three-kernel form.
// scratch: scan_scratch_elems(n) int32.  total (device, may be null) receives the sum.
int64_t scan_scratch_elems(int64_t n);
int scan_exclusive_i32(const int* in, int* out, int64_t n, int* total, int* scratch, cudaStream_t s);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

constexpr int kNumSMs = 148;  // B200

// Optional per-kernel timing with CUDA events on the launching stream (tmae_profile_begin / _end); off by default
// and free when off.  flops / bytes are the ALGORITHMIC work of the launch.
bool prof_enabled();
void count_launch();    // kernel-family launches made by this process (always counted; atomic)
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: set once per (kernel, device), thread-safe
int smem_attr_once(const void* kern, int bytes);
// which kernel family served a GEMM-shaped call (tmae_dispatch_counts): tensor-core TMA path, the fp32 SIMT kernel
// in parity mode, or the fp32 SIMT kernel taken INSIDE a tensor-core mode because TMA cannot express the shape
enum Dispatch { DISP_TMA = 0, DISP_SIMT_FP32 = 1, DISP_SIMT_IN_TC_MODE = 2, DISP_THIN_K = 3, DISP_N = 4 };
void count_dispatch(int which);
void prof_push(const char* name, double flops, double bytes, cudaStream_t s, bool begin);
struct ProfScope {
  const char* name; double flops, bytes; cudaStream_t s; bool on;
  ProfScope(const char* n, double f, double b, cudaStream_t st) : name(n), flops(f), bytes(b), s(st), on(prof_enabled()) {
    count_launch();
    if (on) prof_push(name, flops, bytes, s, true);
  }
  ~ProfScope() { if (on) prof_push(name, flops, bytes, s, false); }
};

}  // namespace tmae
