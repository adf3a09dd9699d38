// Library-level entry points: version, error string, and the shared prefix-sum utility.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "common.cuh"

namespace tmae {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------- per-kernel profiler (CUDA events)
struct ProfRec { const char* name; double flops, bytes; cudaEvent_t e0, e1; };
static bool g_prof = false;
static std::atomic<long long> g_launch_groups{0};
void count_launch() { g_launch_groups.fetch_add(1, std::memory_order_relaxed); }
static std::atomic<long long> g_dispatch[DISP_N];
void count_dispatch(int which) { if (which >= 0 && which < DISP_N) g_dispatch[which].fetch_add(1, std::memory_order_relaxed); }

int smem_attr_once(const void* kern, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return TMAE_ERR_CUDA;
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({kern, dev})) return 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return TMAE_ERR_CUDA;
  done.insert({kern, dev});
  return 0;
}
static std::vector<ProfRec> g_recs;
bool prof_enabled() { return g_prof; }
void prof_push(const char* name, double flops, double bytes, cudaStream_t s, bool begin) {
  if (begin) {
    ProfRec r{name, flops, bytes, nullptr, nullptr};
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, s);
    g_recs.push_back(r);
  } else {
    // scopes do not nest with the same name: the matching record is the most recent one with this name
    for (size_t i = g_recs.size(); i-- > 0;)
      if (g_recs[i].name == name) { cudaEventRecord(g_recs[i].e1, s); break; }
  }
}

// ---------------------------------------------------------------- exclusive scan (int32)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;  // per thread -> 4096 per block
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

int64_t scan_scratch_elems(int64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 8; }

__device__ __forceinline__ int block_exclusive_scan(int v, int* smem, int& block_total) {
  // 256 threads: warp scan + scan of 8 warp totals
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[w] = inc;
  __syncthreads();
  if (w == 0) {
    int t = lane < SCAN_THREADS / 32 ? smem[lane] : 0;
    int ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    if (lane < SCAN_THREADS / 32) smem[lane] = ti - t;
    if (lane == 31) smem[32] = ti;
  }
  __syncthreads();
  block_total = smem[32];
  int r = inc - v + smem[w];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const int* __restrict__ in, int64_t n, int* __restrict__ bsum) {
  __shared__ int sm[33];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
  int tot;
  block_exclusive_scan(s, sm, tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_blocksums_kernel(int* __restrict__ bsum, int nb, int* __restrict__ total) {
  __shared__ int sm[33];
  int carry = 0;
  for (int base = 0; base < nb; base += SCAN_THREADS) {
    int i = base + threadIdx.x;
    int v = i < nb ? bsum[i] : 0;
    int tot;
    int ex = block_exclusive_scan(v, sm, tot);
    if (i < nb) bsum[i] = ex + carry;
    carry += tot;
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const int* __restrict__ in, int* __restrict__ out, int64_t n,
                                                                  const int* __restrict__ bsum) {
  __shared__ int sm[33];
  // blocked arrangement: thread t owns items [t*ITEMS, (t+1)*ITEMS) of the tile
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k;
    v[k] = i < n ? in[i] : 0;
    s += v[k];
  }
  int tot;
  int ex = block_exclusive_scan(s, sm, tot) + bsum[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k;
    if (i < n) out[i] = ex;
    ex += v[k];
  }
}

int scan_exclusive_i32(const int* in, int* out, int64_t n, int* total, int* scratch, cudaStream_t s) {
  if (n <= 0) {
    if (total) cudaMemsetAsync(total, 0, sizeof(int), s);
    return 0;
  }
  int nb = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  scan_reduce_kernel<<<nb, SCAN_THREADS, 0, s>>>(in, n, scratch);
  scan_blocksums_kernel<<<1, SCAN_THREADS, 0, s>>>(scratch, nb, total);
  scan_apply_kernel<<<nb, SCAN_THREADS, 0, s>>>(in, out, n, scratch);
  return cudaGetLastError() == cudaSuccess ? 0 : TMAE_ERR_CUDA;
}

}  // namespace tmae

extern "C" {

const char* tmae_last_error_string(void) { return tmae::g_err; }

int tmae_version(void) { return TMAE_ABI_VERSION; }

int tmae_device_check(void) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    tmae::set_error("tmae_device_check: no CUDA device");
    return TMAE_ERR_CUDA;
  }
  if (p.major != 10) {
    tmae::set_error("tmae_device_check: built for sm_100a, device is sm_%d%d", p.major, p.minor);
    return TMAE_ERR_UNSUPPORTED;
  }
  return 0;
}

int64_t tmae_launch_count(void) { return tmae::g_launch_groups.load(); }

int tmae_dispatch_counts(int64_t* out, int32_t n) {
  for (int i = 0; i < n; ++i) out[i] = i < tmae::DISP_N ? tmae::g_dispatch[i].load() : 0;
  return 0;
}

void tmae_profile_begin(void) {
  cudaDeviceSynchronize();
  for (auto& r : tmae::g_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  tmae::g_recs.clear();
  tmae::g_prof = true;
}

/* Writes one line per kernel family: "name calls total_ms flops bytes\n"; returns the length needed. */
int64_t tmae_profile_end(char* buf, int64_t cap) {
  cudaDeviceSynchronize();
  tmae::g_prof = false;
  struct Agg { double ms = 0, flops = 0, bytes = 0; long calls = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : tmae::g_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      Agg& a = agg[r.name];
      a.ms += ms; a.flops += r.flops; a.bytes += r.bytes; a.calls += 1;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  tmae::g_recs.clear();
  cudaGetLastError();
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s %ld %.6f %.6e %.6e\n", kv.first.c_str(), kv.second.calls, kv.second.ms, kv.second.flops, kv.second.bytes);
    out += line;
  }
  if (buf && cap > 0) {
    size_t n = out.size() < (size_t)cap - 1 ? out.size() : (size_t)cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int64_t)out.size() + 1;
}

int64_t tmae_scan_scratch_elems(int64_t n) { return tmae::scan_scratch_elems(n); }

int tmae_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* total, int32_t* scratch, void* stream) {
  int r = tmae::scan_exclusive_i32(in, out, n, total, scratch, (cudaStream_t)stream);
  if (r) tmae::set_error("tmae_exclusive_scan_i32: launch failed");
  return r;
}

}  // extern "C"
