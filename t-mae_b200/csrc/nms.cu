// Row N4: the finetune inference tail behind `spatial_features` -- CenterHead box decoding and rotated BEV NMS
// (pcdet/models/dense_heads/center_head.py:281-347, pcdet/models/model_utils/centernet_utils.py:154-220,
//  pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:106-311, iou3d_nms.cpp:90-135; enabled by tools/cfgs/once_models/t_mae.yaml:241-249).
// What changes against the reference: the decode of a head (7 gathers, exp, atan2, box assembly, range / score mask, boolean-index
// compaction = ~20 ATen kernels per head) is ONE kernel per head with an in-block stable compaction; the suppression sweep, which the
// reference runs on the HOST after copying the N x N/64 bit mask back (a blocking cudaMemcpy + cudaMalloc/cudaFree per call), runs on
// the device, batched over the samples, with the kept count left on the device -- the whole tail needs one host read at the end.
// The overlap arithmetic is the reference's (edge-edge intersections, corners-in-box with its 1e-2 margin, centroid angular sort, fan
// area), written without FMA contraction so that it agrees with the C oracle (oracle/nms_ref.c) to the last bit of everything except
// sinf / cosf / atan2f.
#include "common.cuh"

namespace tmae {

struct P2 { float x, y; };
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float cross2(P2 a, P2 b) { return sub(mul(a.x, b.y), mul(a.y, b.x)); }
__device__ __forceinline__ float cross3(P2 p1, P2 p2, P2 p0) { return sub(mul(sub(p1.x, p0.x), sub(p2.y, p0.y)), mul(sub(p2.x, p0.x), sub(p1.y, p0.y))); }

__device__ __forceinline__ bool rect_cross(P2 p1, P2 p2, P2 q1, P2 q2) {
  return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) && fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) &&
         fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}
__device__ __forceinline__ bool in_box2d(const float* box, P2 p) {
  const float MARGIN = 1e-2f;   // iou3d_nms_kernel.cu:53: part of the reference's semantics
  const float ac = cosf(-box[6]), as = sinf(-box[6]);
  const float dx = sub(p.x, box[0]), dy = sub(p.y, box[1]);
  const float rx = add(mul(dx, ac), mul(dy, -as));
  const float ry = add(mul(dx, as), mul(dy, ac));
  return fabsf(rx) < add(__fdiv_rn(box[3], 2.f), MARGIN) && fabsf(ry) < add(__fdiv_rn(box[4], 2.f), MARGIN);
}
__device__ __forceinline__ bool seg_intersection(P2 p1, P2 p0, P2 q1, P2 q0, P2& ans) {
  if (!rect_cross(p0, p1, q0, q1)) return false;
  const float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
  if (!(mul(s1, s2) > 0.f && mul(s3, s4) > 0.f)) return false;
  const float s5 = cross3(q1, p1, p0);
  if (fabsf(sub(s5, s1)) > 1e-8f) {
    const float d = sub(s5, s1);
    ans.x = __fdiv_rn(sub(mul(s5, q0.x), mul(s1, q1.x)), d);
    ans.y = __fdiv_rn(sub(mul(s5, q0.y), mul(s1, q1.y)), d);
  } else {
    const float a0 = sub(p0.y, p1.y), b0 = sub(p1.x, p0.x), c0 = sub(mul(p0.x, p1.y), mul(p1.x, p0.y));
    const float a1 = sub(q0.y, q1.y), b1 = sub(q1.x, q0.x), c1 = sub(mul(q0.x, q1.y), mul(q1.x, q0.y));
    const float D = sub(mul(a0, b1), mul(a1, b0));
    ans.x = __fdiv_rn(sub(mul(b0, c1), mul(b1, c0)), D);
    ans.y = __fdiv_rn(sub(mul(a1, c0), mul(a0, c1)), D);
  }
  return true;
}
__device__ __forceinline__ void rot_about(P2 c, float ac, float as, P2& p) {
  const float dx = sub(p.x, c.x), dy = sub(p.y, c.y);
  const float nx = add(add(mul(dx, ac), mul(dy, -as)), c.x);
  const float ny = add(add(mul(dx, as), mul(dy, ac)), c.y);
  p.x = nx; p.y = ny;
}

__device__ float box_overlap(const float* a, const float* b) {
  const float adx = __fdiv_rn(a[3], 2.f), bdx = __fdiv_rn(b[3], 2.f), ady = __fdiv_rn(a[4], 2.f), bdy = __fdiv_rn(b[4], 2.f);
  const P2 ca = {a[0], a[1]}, cb = {b[0], b[1]};
  P2 A[5] = {{sub(a[0], adx), sub(a[1], ady)}, {add(a[0], adx), sub(a[1], ady)}, {add(a[0], adx), add(a[1], ady)}, {sub(a[0], adx), add(a[1], ady)}, {0, 0}};
  P2 B[5] = {{sub(b[0], bdx), sub(b[1], bdy)}, {add(b[0], bdx), sub(b[1], bdy)}, {add(b[0], bdx), add(b[1], bdy)}, {sub(b[0], bdx), add(b[1], bdy)}, {0, 0}};
  const float aco = cosf(a[6]), asi = sinf(a[6]), bco = cosf(b[6]), bsi = sinf(b[6]);
#pragma unroll
  for (int k = 0; k < 4; ++k) { rot_about(ca, aco, asi, A[k]); rot_about(cb, bco, bsi, B[k]); }
  A[4] = A[0]; B[4] = B[0];
  P2 pts[16];
  P2 ctr = {0.f, 0.f};
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      P2 x;
      if (seg_intersection(A[i + 1], A[i], B[j + 1], B[j], x)) { pts[cnt++] = x; ctr.x = add(ctr.x, x.x); ctr.y = add(ctr.y, x.y); }
    }
  for (int k = 0; k < 4; ++k) {
    if (in_box2d(a, B[k])) { ctr.x = add(ctr.x, B[k].x); ctr.y = add(ctr.y, B[k].y); pts[cnt++] = B[k]; }
    if (in_box2d(b, A[k])) { ctr.x = add(ctr.x, A[k].x); ctr.y = add(ctr.y, A[k].y); pts[cnt++] = A[k]; }
  }
  ctr.x = __fdiv_rn(ctr.x, (float)cnt); ctr.y = __fdiv_rn(ctr.y, (float)cnt);
  float ang[16];
  for (int i = 0; i < cnt; ++i) ang[i] = atan2f(sub(pts[i].y, ctr.y), sub(pts[i].x, ctr.x));
  for (int j = 0; j < cnt - 1; ++j)          // the reference's bubble sort (same comparator, same tie behaviour)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (ang[i] > ang[i + 1]) { const P2 t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t; const float u = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = u; }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k) {
    const P2 u = {sub(pts[k].x, pts[0].x), sub(pts[k].y, pts[0].y)}, v = {sub(pts[k + 1].x, pts[0].x), sub(pts[k + 1].y, pts[0].y)};
    area = add(area, cross2(u, v));
  }
  return __fdiv_rn(fabsf(area), 2.f);
}
__device__ __forceinline__ float iou_bev(const float* a, const float* b) {
  const float sa = mul(a[3], a[4]), sb = mul(b[3], b[4]), so = box_overlap(a, b);
  return __fdiv_rn(so, fmaxf(sub(add(sa, sb), so), 1e-8f));
}

__global__ void iou_bev_kernel(const float* __restrict__ a, int64_t na, const float* __restrict__ b, int64_t nb, float* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, i = (int64_t)blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= na || j >= nb) return;
  float ba[7], bb[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) { ba[k] = a[i * 7 + k]; bb[k] = b[j * 7 + k]; }
  out[i * nb + j] = iou_bev(ba, bb);
}

// suppression bit mask, upper triangle only: mask[s][i][cb] bit t = IoU(box i, box cb*64 + t) > thresh for cb*64 + t > i
__global__ void __launch_bounds__(64) nms_mask_kernel(const float* __restrict__ boxes, const int32_t* __restrict__ counts, int64_t cap, float thresh,
                                                       unsigned long long* __restrict__ mask, int col_blocks) {
  const int s = blockIdx.z, row_b = blockIdx.y, col_b = blockIdx.x;
  const int n = min((int)cap, counts[s]);
  if (col_b < row_b || row_b * 64 >= n || col_b * 64 >= n) return;
  const float* bx = boxes + (int64_t)s * cap * 7;
  __shared__ float cbx[64 * 7];
  const int ncol = min(64, n - col_b * 64), nrow = min(64, n - row_b * 64);
  if ((int)threadIdx.x < ncol)
    for (int k = 0; k < 7; ++k) cbx[threadIdx.x * 7 + k] = bx[(int64_t)(col_b * 64 + threadIdx.x) * 7 + k];
  __syncthreads();
  if ((int)threadIdx.x < nrow) {
    const int i = row_b * 64 + threadIdx.x;
    float me[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) me[k] = bx[(int64_t)i * 7 + k];
    unsigned long long t = 0;
    for (int c = (row_b == col_b ? threadIdx.x + 1 : 0); c < ncol; ++c)
      if (iou_bev(me, cbx + c * 7) > thresh) t |= 1ull << c;
    mask[((int64_t)s * cap + i) * col_blocks + col_b] = t;
  }
}

// greedy sweep on the device (iou3d_nms.cpp:116-135): block = one sample, thread j owns suppression word j
__global__ void nms_sweep_kernel(const unsigned long long* __restrict__ mask, const int32_t* __restrict__ counts, int64_t cap, int col_blocks,
                                 int32_t post_max, int64_t* __restrict__ keep, int32_t* __restrict__ num_keep) {
  const int s = blockIdx.x, j = threadIdx.x;
  const int n = min((int)cap, counts[s]);
  extern __shared__ unsigned long long sh[];   // [col_blocks] suppression words, then the kept bits of the current 64-box block
  unsigned long long* remv = sh;
  __shared__ unsigned long long kept_bits;
  __shared__ int kept_total;
  if (j < col_blocks) remv[j] = 0;
  if (j == 0) kept_total = 0;
  __syncthreads();
  const unsigned long long* m = mask + (int64_t)s * cap * col_blocks;
  int64_t* kp = keep + (int64_t)s * cap;
  const int nb = (n + 63) / 64;
  for (int rb = 0; rb < nb; ++rb) {
    if (j == 0) {   // boxes of one 64-block suppress each other sequentially: one thread, register bit operations
      unsigned long long cur = remv[rb], kb = 0;
      const int lim = min(64, n - rb * 64);
      for (int t = 0; t < lim; ++t)
        if (!((cur >> t) & 1ull)) {
          kb |= 1ull << t;
          cur |= m[(int64_t)(rb * 64 + t) * col_blocks + rb];
          if (kept_total < post_max) kp[kept_total] = rb * 64 + t;
          ++kept_total;
        }
      kept_bits = kb;
    }
    __syncthreads();
    if (j > rb && j < nb) {   // ... and every later word takes the OR of the kept rows, in parallel
      unsigned long long kb = kept_bits, acc = remv[j];
      while (kb) {
        const int t = __ffsll((long long)kb) - 1;
        kb &= kb - 1;
        acc |= m[(int64_t)(rb * 64 + t) * col_blocks + j];
      }
      remv[j] = acc;
    }
    __syncthreads();
  }
  if (j == 0) num_keep[s] = min(kept_total, post_max);
}

// one block per (sample): decode the K candidates of a head, mask, stable compaction (centernet_utils.py:164-219)
struct DecodeArgs {
  const float* scores; const int64_t* inds;       // (B, K) top-K of sigmoid(hm) flattened over (class, y, x), descending
  const float *center, *center_z, *dim, *rot, *iou;   // (B,2,H,W) (B,1,H,W) (B,3,H,W) log-dims (B,2,H,W) [cos,sin] (B,1,H,W) or NULL
  const int64_t* class_map;                       // (ncls) head-local class -> global class id
  int B, K, H, W, ncls;
  float stride, vx, vy, r0x, r0y, lim[6], score_thresh;
  float* boxes; float* out_scores; int64_t* labels; float* ious; int32_t* counts;   // (B,K,7) (B,K) (B,K) (B,K) (B)
};
__global__ void __launch_bounds__(1024) centerhead_decode_kernel(DecodeArgs a) {
  const int b = blockIdx.x, k = threadIdx.x;
  const int HW = a.H * a.W;
  __shared__ int warp_tot[32];
  float box[7], sc = 0.f, io = 1.f;
  int64_t lab = 0;
  bool ok = false;
  if (k < a.K) {
    sc = a.scores[(int64_t)b * a.K + k];
    const int64_t ind = a.inds[(int64_t)b * a.K + k];
    const int cls = (int)(ind / HW), pos = (int)(ind % HW);
    const int y = pos / a.W, x = pos % a.W;
    const float* c = a.center + (int64_t)b * 2 * HW;
    const float xs = __fadd_rn((float)x, c[pos]), ys = __fadd_rn((float)y, c[HW + pos]);
    box[0] = __fadd_rn(__fmul_rn(__fmul_rn(xs, a.stride), a.vx), a.r0x);
    box[1] = __fadd_rn(__fmul_rn(__fmul_rn(ys, a.stride), a.vy), a.r0y);
    box[2] = a.center_z[(int64_t)b * HW + pos];
    const float* d = a.dim + (int64_t)b * 3 * HW;
    box[3] = expf(d[pos]); box[4] = expf(d[HW + pos]); box[5] = expf(d[2 * HW + pos]);
    const float* r = a.rot + (int64_t)b * 2 * HW;
    box[6] = atan2f(r[HW + pos], r[pos]);
    if (a.iou) io = fminf(fmaxf(__fmul_rn(__fadd_rn(a.iou[(int64_t)b * HW + pos], 1.f), 0.5f), 0.f), 1.f);
    lab = a.class_map ? a.class_map[cls] : cls;
    ok = box[0] >= a.lim[0] && box[1] >= a.lim[1] && box[2] >= a.lim[2] && box[0] <= a.lim[3] && box[1] <= a.lim[4] && box[2] <= a.lim[5] &&
         sc > a.score_thresh;
  }
  // stable compaction over the block (ballot + warp prefix)
  const unsigned bal = __ballot_sync(0xffffffffu, ok);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) warp_tot[w] = __popc(bal);
  __syncthreads();
  int base = 0;
  for (int i = 0; i < w; ++i) base += warp_tot[i];
  const int dst = base + __popc(bal & ((1u << lane) - 1));
  if (ok) {
    float* o = a.boxes + ((int64_t)b * a.K + dst) * 7;
#pragma unroll
    for (int i = 0; i < 7; ++i) o[i] = box[i];
    a.out_scores[(int64_t)b * a.K + dst] = sc;
    a.labels[(int64_t)b * a.K + dst] = lab;
    a.ious[(int64_t)b * a.K + dst] = io;
  }
  if (threadIdx.x == blockDim.x - 1) a.counts[b] = base + __popc(bal);
}

}  // namespace tmae

using namespace tmae;

extern "C" {

/* ans_iou (na, nb) = rotated BEV IoU: iou3d_nms_utils.boxes_iou_bev (iou3d_nms_kernel.cu:251-264) */
int tmae_boxes_iou_bev(const float* boxes_a, int64_t na, const float* boxes_b, int64_t nb, float* ans_iou, void* stream) {
  if (na <= 0 || nb <= 0) return 0;
  dim3 block(16, 16), grid((unsigned)cdiv(nb, 16), (unsigned)cdiv(na, 16));
  iou_bev_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(boxes_a, na, boxes_b, nb, ans_iou);
  TMAE_CHECK_LAUNCH();
  return 0;
}

size_t tmae_nms_bev_workspace_bytes(int32_t samples, int64_t cap) {
  return (size_t)samples * cap * cdiv(cap, 64) * sizeof(unsigned long long) + 256;
}

/* boxes (samples, cap, 7), each sample's first counts[s] rows sorted by descending score -> keep (samples, cap) i64 kept row indices (first
 * num_keep[s], at most post_max), num_keep (samples) i32: iou3d_nms_utils.nms_gpu + the [:NMS_POST_MAXSIZE] cut of class_agnostic_nms. */
int tmae_nms_bev(const float* boxes, const int32_t* counts, int32_t samples, int64_t cap, float thresh, int32_t post_max, int64_t* keep,
                 int32_t* num_keep, void* workspace, size_t workspace_bytes, void* stream) {
  TMAE_CHECK_ARG(cap > 0 && cap <= 65536 && samples > 0, "1 <= cap <= 65536 boxes per sample");
  TMAE_CHECK_ARG(workspace && workspace_bytes >= tmae_nms_bev_workspace_bytes(samples, cap), "workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int cb = cdiv(cap, 64);
  unsigned long long* mask = (unsigned long long*)workspace;
  ProfScope prof("nms_bev", 0, 28.0 * samples * cap, s);
  dim3 grid((unsigned)cb, (unsigned)cb, (unsigned)samples);
  nms_mask_kernel<<<grid, 64, 0, s>>>(boxes, counts, cap, thresh, mask, cb);
  int threads = cb < 32 ? 32 : (cb + 31) / 32 * 32;
  nms_sweep_kernel<<<samples, threads, (size_t)cb * sizeof(unsigned long long), s>>>(mask, counts, cap, cb, post_max, keep, num_keep);
  TMAE_CHECK_LAUNCH();
  return 0;
}

/* decode_bbox_from_heatmap after the top-K selection (centernet_utils.py:165-219) for one head; see DecodeArgs. */
int tmae_centerhead_decode(const float* scores, const int64_t* inds, const float* center, const float* center_z, const float* dim, const float* rot,
                           const float* iou, const int64_t* class_map, int32_t batch, int32_t k, int32_t h, int32_t w, int32_t ncls,
                           float feature_map_stride, const float* voxel_size, const float* range_lo, const float* limit_range, float score_thresh,
                           float* boxes, float* out_scores, int64_t* labels, float* ious, int32_t* counts, void* stream) {
  TMAE_CHECK_ARG(k > 0 && k <= 1024 && batch > 0, "1 <= MAX_OBJ_PER_SAMPLE <= 1024");
  DecodeArgs a;
  a.scores = scores; a.inds = inds; a.center = center; a.center_z = center_z; a.dim = dim; a.rot = rot; a.iou = iou; a.class_map = class_map;
  a.B = batch; a.K = k; a.H = h; a.W = w; a.ncls = ncls; a.stride = feature_map_stride; a.vx = voxel_size[0]; a.vy = voxel_size[1];
  a.r0x = range_lo[0]; a.r0y = range_lo[1];
  for (int i = 0; i < 6; ++i) a.lim[i] = limit_range[i];
  a.score_thresh = score_thresh; a.boxes = boxes; a.out_scores = out_scores; a.labels = labels; a.ious = ious; a.counts = counts;
  const int threads = (k + 31) / 32 * 32;
  ProfScope prof("centerhead_decode", 0, 0, (cudaStream_t)stream);
  centerhead_decode_kernel<<<batch, threads, 0, (cudaStream_t)stream>>>(a);
  TMAE_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
