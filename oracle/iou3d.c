/* ORACLE (test infrastructure, never linked into the product) — BEV IoU of rotated boxes and rotated / axis-aligned NMS
 * on the CPU.  Restates, step by step and in the same fp32 operation order:
 *   box_overlap / iou_bev      efg/operators/src/iou3d_nms/iou3d_cpu.cpp:61-214 (== iou3d_nms_kernel.cu:98-237)
 *   iou_normal                 iou3d_nms_kernel.cu:331-342
 *   the greedy suppression     iou3d_nms.cpp:87-121 (host scan over the 64-wide bit masks of nms_kernel, :283-329)
 * Quirks that are part of the contract and therefore kept: a corner counts as inside the other box with a 1 cm margin
 * (MARGIN = 1e-2), two edges intersect only when both cross-standing products are strictly positive, the polygon is
 * ordered by a bubble sort on atan2 around the mean of its points.
 * Pinned: tests/golden/iou3d_*.npz hold the output of the reference's own iou3d_cpu.cpp (compiled from
 * /root/reference into oracle/_ref by oracle/build_ref.py). */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define EPS 1e-8f

typedef struct { float x, y; } P2;

static inline float cross2(P2 a, P2 b) { return a.x * b.y - a.y * b.x; }
static inline float cross3(P2 p1, P2 p2, P2 p0) { return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y); }

static int rect_cross(P2 p1, P2 p2, P2 q1, P2 q2) {
  return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
         fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}

static int in_box(const float* box, P2 p) {
  const float margin = 1e-2f;
  const float c = cosf(-box[6]), s = sinf(-box[6]);
  const float rx = (p.x - box[0]) * c + (p.y - box[1]) * (-s);
  const float ry = (p.x - box[0]) * s + (p.y - box[1]) * c;
  return fabsf(rx) < box[3] / 2 + margin && fabsf(ry) < box[4] / 2 + margin;
}

static int seg_intersection(P2 p1, P2 p0, P2 q1, P2 q0, P2* ans) {
  if (!rect_cross(p0, p1, q0, q1)) return 0;
  const float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
  const float s5 = cross3(q1, p1, p0);
  if (fabsf(s5 - s1) > EPS) {
    ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
    ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
  } else {
    const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    const float d = a0 * b1 - a1 * b0;
    ans->x = (b0 * c1 - b1 * c0) / d;
    ans->y = (a1 * c0 - a0 * c1) / d;
  }
  return 1;
}

static void corners_of(const float* b, P2* c) {
  const float hx = b[3] / 2, hy = b[4] / 2;
  const float x1 = b[0] - hx, y1 = b[1] - hy, x2 = b[0] + hx, y2 = b[1] + hy;
  const float ca = cosf(b[6]), sa = sinf(b[6]);
  const float xs[4] = {x1, x2, x2, x1}, ys[4] = {y1, y1, y2, y2};
  for (int k = 0; k < 4; ++k) {
    c[k].x = (xs[k] - b[0]) * ca + (ys[k] - b[1]) * (-sa) + b[0];
    c[k].y = (xs[k] - b[0]) * sa + (ys[k] - b[1]) * ca + b[1];
  }
  c[4] = c[0];
}

float oracle_box_overlap(const float* a, const float* b) {
  P2 ca[5], cb[5], pts[16], centre = {0.f, 0.f};
  corners_of(a, ca);
  corners_of(b, cb);
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (seg_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], &pts[cnt])) {
        centre.x += pts[cnt].x;
        centre.y += pts[cnt].y;
        ++cnt;
      }
  for (int k = 0; k < 4; ++k) {
    if (in_box(a, cb[k])) {
      centre.x += cb[k].x;
      centre.y += cb[k].y;
      pts[cnt++] = cb[k];
    }
    if (in_box(b, ca[k])) {
      centre.x += ca[k].x;
      centre.y += ca[k].y;
      pts[cnt++] = ca[k];
    }
  }
  centre.x /= cnt;
  centre.y /= cnt;
  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (atan2f(pts[i].y - centre.y, pts[i].x - centre.x) > atan2f(pts[i + 1].y - centre.y, pts[i + 1].x - centre.x)) {
        P2 t = pts[i];
        pts[i] = pts[i + 1];
        pts[i + 1] = t;
      }
  float area = 0;
  for (int k = 0; k < cnt - 1; ++k) {
    P2 u = {pts[k].x - pts[0].x, pts[k].y - pts[0].y}, v = {pts[k + 1].x - pts[0].x, pts[k + 1].y - pts[0].y};
    area += cross2(u, v);
  }
  return fabsf(area) / 2.0f;
}

float oracle_iou_bev(const float* a, const float* b) {
  const float sa = a[3] * a[4], sb = b[3] * b[4], so = oracle_box_overlap(a, b);
  return so / fmaxf(sa + sb - so, EPS);
}

static float iou_normal(const float* a, const float* b) {
  const float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
  const float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
  const float inter = fmaxf(right - left, 0.f) * fmaxf(bottom - top, 0.f);
  return inter / fmaxf(a[3] * a[4] + b[3] * b[4] - inter, EPS);
}

/* out[i * nb + j]; mode 0 = IoU, 1 = overlap area */
void oracle_boxes_bev(const float* a, int64_t na, const float* b, int64_t nb, int mode, float* out) {
  for (int64_t i = 0; i < na; ++i)
    for (int64_t j = 0; j < nb; ++j)
      out[i * nb + j] = mode ? oracle_box_overlap(a + i * 7, b + j * 7) : oracle_iou_bev(a + i * 7, b + j * 7);
}

/* Greedy suppression over boxes already sorted by descending score: box i is kept unless an earlier kept box j < i has
 * IoU(j, i) > thresh (the mask bit (j, i) of nms_kernel is computed as iou(box j, box i) with j the row).  Returns the
 * number of kept boxes, indices in keep[]. */
int64_t oracle_nms(const float* boxes, int64_t n, float thresh, int normal, int64_t* keep) {
  int64_t nk = 0;
  for (int64_t i = 0; i < n; ++i) {
    int removed = 0;
    for (int64_t k = 0; k < nk && !removed; ++k) {
      const float* bj = boxes + keep[k] * 7;
      const float v = normal ? iou_normal(bj, boxes + i * 7) : oracle_iou_bev(bj, boxes + i * 7);
      if (v > thresh) removed = 1;
    }
    if (!removed) keep[nk++] = i;
  }
  return nk;
}
