/* TEST INFRASTRUCTURE ONLY -- the oracle of row N4 (CenterHead post-processing): a plain-C restatement of the reference's rotated
 * BEV IoU and greedy NMS.  Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may link or call this file; the product
 * (t-mae_b200/) never does.
 *
 * Follows pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu (all paths relative to /root/reference):
 *   cross products                      :35-41        check_rect_cross  :43-49      check_in_box2d (MARGIN 1e-2)  :51-61
 *   segment intersection                :63-94        rotate_around_center :96-100  point_cmp (atan2 order)      :102-104
 *   box_overlap                         :106-220      iou_bev           :222-230
 *   nms_kernel (64-wide suppression mask, upper triangle)  :267-311
 * and the host sweep of pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:116-135 (keep box i unless an earlier kept box has IoU > thresh with it).
 * Boxes are (N, 7) float32 [x, y, z, dx, dy, dz, heading], already sorted by descending score (iou3d_nms_utils.py:92-96).
 * Pinned against the reference's own CPU implementation (pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp, compiled from where it lies by
 * oracle/build_ref.py into oracle/_ref/) and against tests/golden/nms.npz made with it: tests/test_oracle_nms.py. */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define EPSF 1e-8f

typedef struct { float x, y; } P2;

static float cross2(P2 a, P2 b) { return a.x * b.y - a.y * b.x; }
static float cross3(P2 p1, P2 p2, P2 p0) { return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y); }
static float fmin2(float a, float b) { return a < b ? a : b; }
static float fmax2(float a, float b) { return a > b ? a : b; }

static int rect_cross(P2 p1, P2 p2, P2 q1, P2 q2) {
  return fmin2(p1.x, p2.x) <= fmax2(q1.x, q2.x) && fmin2(q1.x, q2.x) <= fmax2(p1.x, p2.x) && fmin2(p1.y, p2.y) <= fmax2(q1.y, q2.y) &&
         fmin2(q1.y, q2.y) <= fmax2(p1.y, p2.y);
}

static int in_box2d(const float* box, P2 p) {
  const float MARGIN = 1e-2f;
  float cx = box[0], cy = box[1];
  float ac = cosf(-box[6]), as = sinf(-box[6]);
  float rx = (p.x - cx) * ac + (p.y - cy) * (-as);
  float ry = (p.x - cx) * as + (p.y - cy) * ac;
  return fabsf(rx) < box[3] / 2 + MARGIN && fabsf(ry) < box[4] / 2 + MARGIN;
}

static int seg_intersection(P2 p1, P2 p0, P2 q1, P2 q0, P2* ans) {
  if (!rect_cross(p0, p1, q0, q1)) return 0;
  float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
  float s5 = cross3(q1, p1, p0);
  if (fabsf(s5 - s1) > EPSF) {
    ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
    ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
  } else {
    float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    float D = a0 * b1 - a1 * b0;
    ans->x = (b0 * c1 - b1 * c0) / D;
    ans->y = (a1 * c0 - a0 * c1) / D;
  }
  return 1;
}

static void rot_about(P2 c, float ac, float as, P2* p) {
  float nx = (p->x - c.x) * ac + (p->y - c.y) * (-as) + c.x;
  float ny = (p->x - c.x) * as + (p->y - c.y) * ac + c.y;
  p->x = nx; p->y = ny;
}

float tmae_oracle_box_overlap(const float* a, const float* b) {
  float adx = a[3] / 2, bdx = b[3] / 2, ady = a[4] / 2, bdy = b[4] / 2;
  P2 ca = {a[0], a[1]}, cb = {b[0], b[1]};
  P2 A[5] = {{a[0] - adx, a[1] - ady}, {a[0] + adx, a[1] - ady}, {a[0] + adx, a[1] + ady}, {a[0] - adx, a[1] + ady}};
  P2 B[5] = {{b[0] - bdx, b[1] - bdy}, {b[0] + bdx, b[1] - bdy}, {b[0] + bdx, b[1] + bdy}, {b[0] - bdx, b[1] + bdy}};
  float aco = cosf(a[6]), asi = sinf(a[6]), bco = cosf(b[6]), bsi = sinf(b[6]);
  for (int k = 0; k < 4; ++k) { rot_about(ca, aco, asi, &A[k]); rot_about(cb, bco, bsi, &B[k]); }
  A[4] = A[0]; B[4] = B[0];
  P2 pts[16], ctr = {0, 0};
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (seg_intersection(A[i + 1], A[i], B[j + 1], B[j], &pts[cnt])) { ctr.x += pts[cnt].x; ctr.y += pts[cnt].y; ++cnt; }
  for (int k = 0; k < 4; ++k) {
    if (in_box2d(a, B[k])) { ctr.x += B[k].x; ctr.y += B[k].y; pts[cnt++] = B[k]; }
    if (in_box2d(b, A[k])) { ctr.x += A[k].x; ctr.y += A[k].y; pts[cnt++] = A[k]; }
  }
  ctr.x /= cnt; ctr.y /= cnt;
  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (atan2f(pts[i].y - ctr.y, pts[i].x - ctr.x) > atan2f(pts[i + 1].y - ctr.y, pts[i + 1].x - ctr.x)) { P2 t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t; }
  float area = 0;
  for (int k = 0; k < cnt - 1; ++k) {
    P2 u = {pts[k].x - pts[0].x, pts[k].y - pts[0].y}, v = {pts[k + 1].x - pts[0].x, pts[k + 1].y - pts[0].y};
    area += cross2(u, v);
  }
  return fabsf(area) / 2.0f;
}

float tmae_oracle_iou_bev(const float* a, const float* b) {
  float sa = a[3] * a[4], sb = b[3] * b[4], so = tmae_oracle_box_overlap(a, b);
  return so / fmaxf(sa + sb - so, EPSF);
}

void tmae_oracle_boxes_iou_bev(const float* a, int64_t na, const float* b, int64_t nb, float* out) {
  for (int64_t i = 0; i < na; ++i)
    for (int64_t j = 0; j < nb; ++j) out[i * nb + j] = tmae_oracle_iou_bev(a + i * 7, b + j * 7);
}

/* boxes sorted by descending score; keep (n) receives the kept indices; returns their number */
int64_t tmae_oracle_nms(const float* boxes, int64_t n, float thresh, int64_t* keep) {
  int64_t nk = 0;
  for (int64_t i = 0; i < n; ++i) {
    int sup = 0;
    for (int64_t t = 0; t < nk && !sup; ++t) sup = tmae_oracle_iou_bev(boxes + keep[t] * 7, boxes + i * 7) > thresh;
    if (!sup) keep[nk++] = i;
  }
  return nk;
}
