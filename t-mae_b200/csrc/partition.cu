// Window partition, single-frame and temporal (two-frame) forms.
//
// Replaces, in one device-side pass with no host synchronisation:
//   get_window_coors             pcdet/models/model_utils/sst_utils.py:6-58
//   drop_single_shift/drop_voxel pcdet/models/backbones_3d/spt_backbone.py:47-135
//   drop_single_shift_ref_to_prv pcdet/models/backbones_3d/SiamWCA.py:65-199
//   get_inner_win_inds           pcdet/ops/sst_ops/src/sst_ops_gpu.cu:14-20  (canonical slot order)
//   make_continuous_inds / get_flat2win_inds  sst_utils.py:61-115
//
// B200-first design: a window is 8x8 = 64 cells, so ONE 64-bit occupancy word per window and
// shift holds everything: count = popc(mask), canonical slot = popc(mask & below(bit)).  Voxels
// arrive in lexicographic (b,y,x) order (the voxeliser emits them so), hence "number of occupied
// cells before mine in row-major order" equals the stable rank by element index that a serial run
// of the reference's atomic kernel yields.  atomicOr is commutative, so the result is
// deterministic.  Compact window ids come from one prefix sum over a level-major flag array, which
// reproduces make_continuous_inds' ascending-id order per level and at the same time sorts the
// attention work list by window size class.
#include "common.cuh"

namespace tmae {

constexpr int WIN = 8;

struct Levels {
  int n;
  int lo[TMAE_MAX_LEVELS], hi[TMAE_MAX_LEVELS], tok[TMAE_MAX_LEVELS];
};

struct PartGeom {
  int gx, gy, batch;
  int wx, wy;     // windows per sample along x / y (ceil(g/8)+1, sst_utils.py:23-25)
  int wz;         // 2: the reference's id formula keeps a z factor of ceil(1/1)+1
  int wcap;       // batch * wx * wy
};

__device__ __forceinline__ void voxel_window(const int* __restrict__ c, int shift, const PartGeom& G, int& w, int& bit) {
  int sh = shift ? WIN / 2 : WIN;  // sst_utils.py:29-32
  int x = c[2] + sh, y = c[1] + sh;
  w = (c[0] * G.wx + x / WIN) * G.wy + y / WIN;
  bit = (y % WIN) * WIN + (x % WIN);
}

// status[0] |= 1 : coords not in strictly ascending (b,y,x) order ; |= 2 : out-of-grid coordinate
__global__ void part_mark_kernel(const int* __restrict__ coords, int64_t m, PartGeom G, unsigned long long* __restrict__ mask,
                                 int* __restrict__ status) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * m) return;
  int s = (int)(t / m);
  int64_t i = t - (int64_t)s * m;
  const int* c = coords + i * 3;
  if (s == 0) {
    bool bad = c[0] < 0 || c[0] >= G.batch || c[1] < 0 || c[1] >= G.gy || c[2] < 0 || c[2] >= G.gx;
    if (bad) { atomicOr(status, 2); }
    if (i > 0) {
      const int* p = c - 3;
      bool asc = p[0] < c[0] || (p[0] == c[0] && (p[1] < c[1] || (p[1] == c[1] && p[2] < c[2])));
      if (!asc) atomicOr(status, 1);
    }
    if (bad) return;
  } else if (c[0] < 0 || c[0] >= G.batch || c[1] < 0 || c[1] >= G.gy || c[2] < 0 || c[2] >= G.gx) {
    return;
  }
  int w, bit;
  voxel_window(c, s, G, w, bit);
  atomicOr(mask + (int64_t)s * G.wcap + w, 1ull << bit);
}

__device__ __forceinline__ int level_of(int cnt, const Levels& L) {
  int lvl = -1;
  for (int l = 0; l < L.n; ++l)
    if (cnt >= L.lo[l] && cnt < L.hi[l]) lvl = l;  // later levels overwrite (spt_backbone.py:56-60)
  return lvl;
}

// one thread per (shift, window): level flag for the compaction scan.
// mask_b == nullptr: single frame.  Otherwise temporal: level from max(count_a, count_b), window
// kept only when non-empty in both frames (SiamWCA.py:86-118).
__global__ void part_window_flags_kernel(const unsigned long long* __restrict__ mask_a, const unsigned long long* __restrict__ mask_b,
                                         PartGeom G, Levels L, int* __restrict__ flags, int* __restrict__ status) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * G.wcap) return;
  int s = t / G.wcap, w = t - s * G.wcap;
  int ca = __popcll(mask_a[t]);
  int cb = mask_b ? __popcll(mask_b[t]) : ca;
  if (ca == 0 || cb == 0) return;
  int cnt = max(ca, cb);
  int lvl = level_of(cnt, L);
  if (lvl < 0) { atomicOr(status, 4); return; }      // reference: assert (drop_lvl_per_voxel >= 0).all()
  if (cnt > L.tok[lvl]) atomicOr(status, 8);          // this level would drop voxels: not supported
  flags[((int64_t)s * L.n + lvl) * G.wcap + w] = 1;
}

// one thread per (shift, window): compact ids.  win_gid[s][w] = index in the level-major list of
// kept windows of shift s (or -1); win_conti[s][w] = index inside its level (make_continuous_inds).
__global__ void part_window_ids_kernel(const unsigned long long* __restrict__ mask_a, const unsigned long long* __restrict__ mask_b,
                                       PartGeom G, Levels L, const int* __restrict__ flags, const int* __restrict__ pos,
                                       const int* __restrict__ total, int* __restrict__ win_gid, int* __restrict__ win_conti,
                                       int* __restrict__ win_level, int* __restrict__ cnt_a, int* __restrict__ cnt_b,
                                       int* __restrict__ n_win, int* __restrict__ level_base) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 2 * (L.n + 1)) {  // level_base[s][l] = first compact id of level l ; [s][n] = n_win[s]
    int s = t / (L.n + 1), l = t - s * (L.n + 1);
    int64_t e = ((int64_t)s * L.n + l) * G.wcap;
    int64_t b0 = (int64_t)s * L.n * G.wcap;
    int endv = (e >= 2ll * L.n * G.wcap) ? *total : pos[e];
    int v = endv - pos[b0];
    level_base[t] = v;
    if (l == L.n) n_win[s] = v;
  }
  if (t >= 2 * G.wcap) return;
  int s = t / G.wcap, w = t - s * G.wcap;
  int ca = __popcll(mask_a[t]);
  int cb = mask_b ? __popcll(mask_b[t]) : ca;
  int g = -1, conti = -1;
  if (ca != 0 && cb != 0) {
    int lvl = level_of(max(ca, cb), L);
    if (lvl >= 0) {
      int64_t e = ((int64_t)s * L.n + lvl) * G.wcap;
      if (flags[e + w]) {
        g = pos[e + w] - pos[(int64_t)s * L.n * G.wcap];
        conti = pos[e + w] - pos[e];
        win_level[(int64_t)s * G.wcap + g] = lvl;
        cnt_a[(int64_t)s * G.wcap + g] = ca;
        if (cnt_b) cnt_b[(int64_t)s * G.wcap + g] = cb;
      }
    }
  }
  win_gid[t] = g;
  win_conti[t] = conti;
}

struct VoxelOut {
  int* win;        // [2][m]  compact window id or -1
  int* slot;       // [2][m]
  uint8_t* posidx; // [2][m]  ly*8+lx : row of the position-embedding LUT
  int* tok;        // [2][wcap*64] voxel row per (window, slot)
  int64_t* bwi;    // [2][m] reference batch_win_inds   (nullable)
  int64_t* lvl;    // [2][m] reference voxel_drop_level (nullable; -1 when dropped)
  int64_t* f2w;    // [2][m] reference flat2win index   (nullable; -1 when dropped)
};

__global__ void part_voxels_kernel(const int* __restrict__ coords, int64_t m, PartGeom G, Levels L,
                                   const unsigned long long* __restrict__ mask, const int* __restrict__ win_gid,
                                   const int* __restrict__ win_conti, const int* __restrict__ win_level, VoxelOut o) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * m) return;
  int s = (int)(t / m);
  int64_t i = t - (int64_t)s * m;
  const int* c = coords + i * 3;
  int w, bit;
  voxel_window(c, s, G, w, bit);
  int64_t sw = (int64_t)s * G.wcap + w;
  int slot = __popcll(mask[sw] & ((1ull << bit) - 1ull));
  int g = win_gid[sw];
  o.win[t] = g;
  o.slot[t] = slot;
  o.posidx[t] = (uint8_t)bit;
  if (g >= 0) o.tok[((int64_t)s * G.wcap + g) * TMAE_WIN_TOKENS + slot] = (int)i;
  if (o.bwi) {
    int sh = s ? WIN / 2 : WIN;
    int wxi = (c[2] + sh) / WIN, wyi = (c[1] + sh) / WIN;
    o.bwi[t] = (int64_t)c[0] * (G.wx * G.wy * G.wz) + (int64_t)wxi * (G.wy * G.wz) + (int64_t)wyi * G.wz;  // sst_utils.py:48-51
  }
  if (o.lvl) {
    int lvl = g >= 0 ? win_level[(int64_t)s * G.wcap + g] : -1;
    o.lvl[t] = lvl;
    if (o.f2w) o.f2w[t] = g >= 0 ? (int64_t)win_conti[sw] * L.tok[lvl] + slot : -1;
  }
}

static PartGeom make_geom(int batch, int gx, int gy) {
  PartGeom G;
  G.gx = gx; G.gy = gy; G.batch = batch;
  G.wx = (gx + WIN - 1) / WIN + 1;
  G.wy = (gy + WIN - 1) / WIN + 1;
  G.wz = 2;
  G.wcap = batch * G.wx * G.wy;
  return G;
}

static int make_levels(Levels& L, int n, const int32_t* lo, const int32_t* hi, const int32_t* tok) {
  if (n < 1 || n > TMAE_MAX_LEVELS) return -1;
  L.n = n;
  for (int i = 0; i < n; ++i) {
    L.lo[i] = lo[i]; L.hi[i] = hi[i]; L.tok[i] = tok[i];
    if (tok[i] < 1 || tok[i] > TMAE_WIN_TOKENS) return -1;
  }
  return 0;
}

}  // namespace tmae

using namespace tmae;

extern "C" {

int64_t tmae_partition_window_capacity(int32_t batch, int32_t grid_x, int32_t grid_y) {
  return make_geom(batch, grid_x, grid_y).wcap;
}

size_t tmae_window_partition_workspace_bytes(int32_t batch, int32_t grid_x, int32_t grid_y, int32_t n_levels) {
  PartGeom G = make_geom(batch, grid_x, grid_y);
  int64_t nflag = 2ll * n_levels * G.wcap;
  size_t b = 0;
  b += ws_bytes(2ll * G.wcap, 8) * 2;       // masks (two frames)
  b += ws_bytes(nflag, 4) * 2;              // flags, pos
  b += ws_bytes(2ll * G.wcap, 4) * 2;       // win_gid, win_conti
  b += ws_bytes(scan_scratch_elems(nflag), 4);
  b += ws_bytes(8, 4);
  return b + 1024;
}

/* see include/tmae_sm100.h */
int tmae_window_partition(const int32_t* coords_a, int64_t m_a, const int32_t* coords_b, int64_t m_b, int32_t batch,
                          int32_t grid_x, int32_t grid_y, int32_t n_levels, const int32_t* lvl_lo, const int32_t* lvl_hi,
                          const int32_t* lvl_tokens,
                          int32_t* win_a, int32_t* slot_a, uint8_t* posidx_a, int32_t* tok_a, int32_t* cnt_a,
                          int32_t* win_b, int32_t* slot_b, uint8_t* posidx_b, int32_t* tok_b, int32_t* cnt_b,
                          int32_t* win_level, int32_t* n_win, int32_t* level_base, int32_t* status,
                          int64_t* ref_bwi_a, int64_t* ref_lvl_a, int64_t* ref_f2w_a,
                          int64_t* ref_bwi_b, int64_t* ref_lvl_b, int64_t* ref_f2w_b,
                          void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  Levels L;
  TMAE_CHECK_ARG(make_levels(L, n_levels, lvl_lo, lvl_hi, lvl_tokens) == 0, "bad level table");
  TMAE_CHECK_ARG(batch >= 1 && grid_x >= 1 && grid_y >= 1, "bad grid");
  TMAE_CHECK_ARG(workspace_bytes >= tmae_window_partition_workspace_bytes(batch, grid_x, grid_y, n_levels), "workspace too small");
  bool temporal = coords_b != nullptr;
  PartGeom G = make_geom(batch, grid_x, grid_y);
  int64_t nflag = 2ll * L.n * G.wcap;
  Workspace ws(workspace, workspace_bytes);
  auto* mask_a = ws.take<unsigned long long>(2ll * G.wcap);
  auto* mask_b = ws.take<unsigned long long>(2ll * G.wcap);
  int* flags = ws.take<int>(nflag);
  int* pos = ws.take<int>(nflag);
  int* win_gid = ws.take<int>(2ll * G.wcap);
  int* win_conti = ws.take<int>(2ll * G.wcap);
  int* scratch = ws.take<int>(scan_scratch_elems(nflag));
  int* total = ws.take<int>(8);
  TMAE_CHECK_ARG(total != nullptr, "workspace carve failed");

  ProfScope prof("window_partition", 0, (double)(m_a + m_b) * (12.0 + 2 * 9.0 + (ref_bwi_a ? 48.0 : 0.0)) + 2.0 * G.wcap * 8, s);
  TMAE_CUDA(cudaMemsetAsync(mask_a, 0, 2ll * G.wcap * 8, s));
  if (temporal) TMAE_CUDA(cudaMemsetAsync(mask_b, 0, 2ll * G.wcap * 8, s));
  TMAE_CUDA(cudaMemsetAsync(flags, 0, nflag * 4, s));
  TMAE_CUDA(cudaMemsetAsync(status, 0, 4, s));
  const int T = 256;
  if (m_a > 0) part_mark_kernel<<<cdiv(2 * m_a, T), T, 0, s>>>(coords_a, m_a, G, mask_a, status);
  if (temporal && m_b > 0) part_mark_kernel<<<cdiv(2 * m_b, T), T, 0, s>>>(coords_b, m_b, G, mask_b, status);
  TMAE_CHECK_LAUNCH();
  part_window_flags_kernel<<<cdiv(2 * G.wcap, T), T, 0, s>>>(mask_a, temporal ? mask_b : nullptr, G, L, flags, status);
  TMAE_CHECK_LAUNCH();
  if (scan_exclusive_i32(flags, pos, nflag, total, scratch, s)) return TMAE_ERR_CUDA;
  part_window_ids_kernel<<<cdiv(2 * G.wcap, T), T, 0, s>>>(mask_a, temporal ? mask_b : nullptr, G, L, flags, pos, total, win_gid,
                                                          win_conti, win_level, cnt_a, temporal ? cnt_b : nullptr, n_win,
                                                          level_base);
  TMAE_CHECK_LAUNCH();
  if (m_a > 0) {
    VoxelOut o{win_a, slot_a, posidx_a, tok_a, ref_bwi_a, ref_lvl_a, ref_f2w_a};
    part_voxels_kernel<<<cdiv(2 * m_a, T), T, 0, s>>>(coords_a, m_a, G, L, mask_a, win_gid, win_conti, win_level, o);
  }
  if (temporal && m_b > 0) {
    VoxelOut o{win_b, slot_b, posidx_b, tok_b, ref_bwi_b, ref_lvl_b, ref_f2w_b};
    part_voxels_kernel<<<cdiv(2 * m_b, T), T, 0, s>>>(coords_b, m_b, G, L, mask_b, win_gid, win_conti, win_level, o);
  }
  TMAE_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
