// BEV IoU of rotated boxes and rotated / axis-aligned NMS (the CenterPoint evaluation path,
// CP/center_head.py:329-376 -> efg/operators/iou3d_nms.py -> efg/operators/src/iou3d_nms/iou3d_nms_kernel.cu).
//
// Results follow the reference algorithm exactly (the CPU restatement used by the tests is pinned to the reference's own
// iou3d_cpu.cpp): intersection polygon = strict edge crossings + corners inside the other box with a 1 cm margin,
// ordered by atan2 around their mean, shoelace area.  What is different is the machine mapping:
//   * a box is prepared ONCE (four rotated corners, cos / sin of -heading, half sizes, area: 16 floats) instead of
//     re-evaluating four trigonometric functions per PAIR;
//   * NMS computes only the upper-triangular 64 x 64 mask blocks (the reference computes all N^2 / 64 words and
//     copies them to the host) and the greedy scan runs ON THE DEVICE in one CTA: per 64-box block the intra-block
//     decisions are resolved from the diagonal mask word by one thread, then every thread ORs the rows of the kept
//     boxes into its own suppression word with independent loads.  No cudaMalloc, no device-to-host copy of the mask,
//     no host loop: the kept indices and their count stay on the device (the Python `nms_gpu` mirror reads them back
//     because the reference's signature returns a host int).
#include "common.cuh"

namespace efgb {
namespace iou3d {

constexpr float kEps = 1e-8f;
constexpr float kMargin = 1e-2f;
constexpr int kPrep = 16;  // floats per prepared box

struct P2 {
  float x, y;
};

__device__ __forceinline__ float cross3(P2 p1, P2 p2, P2 p0) { return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y); }

// prepared box: [0..7] corners (x0,y0,...,x3,y3), [8] cx, [9] cy, [10] cos(-h), [11] sin(-h), [12] dx/2 + margin,
// [13] dy/2 + margin, [14] dx*dy, [15] unused
__device__ __forceinline__ void prepare(const float* __restrict__ b, float* __restrict__ o) {
  const float hx = b[3] / 2, hy = b[4] / 2;
  const float x1 = b[0] - hx, y1 = b[1] - hy, x2 = b[0] + hx, y2 = b[1] + hy;
  const float ca = cosf(b[6]), sa = sinf(b[6]);
  const float xs[4] = {x1, x2, x2, x1}, ys[4] = {y1, y1, y2, y2};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    o[2 * k] = (xs[k] - b[0]) * ca + (ys[k] - b[1]) * (-sa) + b[0];
    o[2 * k + 1] = (xs[k] - b[0]) * sa + (ys[k] - b[1]) * ca + b[1];
  }
  o[8] = b[0];
  o[9] = b[1];
  o[10] = cosf(-b[6]);
  o[11] = sinf(-b[6]);
  o[12] = b[3] / 2 + kMargin;
  o[13] = b[4] / 2 + kMargin;
  o[14] = b[3] * b[4];
  o[15] = 0.f;
}

__device__ __forceinline__ bool in_box(const float* __restrict__ box, P2 p) {
  const float rx = (p.x - box[8]) * box[10] + (p.y - box[9]) * (-box[11]);
  const float ry = (p.x - box[8]) * box[11] + (p.y - box[9]) * box[10];
  return fabsf(rx) < box[12] && fabsf(ry) < box[13];
}

__device__ __forceinline__ bool seg_intersection(P2 p1, P2 p0, P2 q1, P2 q0, P2* ans) {
  const bool rc = fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) &&
                  fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) && fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y);
  if (!rc) return false;
  const float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return false;
  const float s5 = cross3(q1, p1, p0);
  if (fabsf(s5 - s1) > kEps) {
    ans->x = __fdiv_rn(s5 * q0.x - s1 * q1.x, s5 - s1);
    ans->y = __fdiv_rn(s5 * q0.y - s1 * q1.y, s5 - s1);
  } else {
    const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    const float d = a0 * b1 - a1 * b0;
    ans->x = __fdiv_rn(b0 * c1 - b1 * c0, d);
    ans->y = __fdiv_rn(a1 * c0 - a0 * c1, d);
  }
  return true;
}

// overlap area of two prepared boxes (a, b: kPrep floats each, any address space)
__device__ float overlap(const float* __restrict__ a, const float* __restrict__ b) {
  P2 ca[5], cb[5], pts[16];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    ca[k] = {a[2 * k], a[2 * k + 1]};
    cb[k] = {b[2 * k], b[2 * k + 1]};
  }
  ca[4] = ca[0];
  cb[4] = cb[0];
  int cnt = 0;
  float sx = 0.f, sy = 0.f;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      P2 x;
      if (seg_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], &x)) {
        pts[cnt++] = x;
        sx += x.x;
        sy += x.y;
      }
    }
  for (int k = 0; k < 4; ++k) {
    if (in_box(a, cb[k])) {
      sx += cb[k].x;
      sy += cb[k].y;
      pts[cnt++] = cb[k];
    }
    if (in_box(b, ca[k])) {
      sx += ca[k].x;
      sy += ca[k].y;
      pts[cnt++] = ca[k];
    }
  }
  if (cnt < 3) return 0.f;  // the reference's loops produce |0| / 2 for fewer than three points (NaN centre unused)
  const float cx = __fdiv_rn(sx, static_cast<float>(cnt)), cy = __fdiv_rn(sy, static_cast<float>(cnt));
  float ang[16];
  for (int k = 0; k < cnt; ++k) ang[k] = atan2f(pts[k].y - cy, pts[k].x - cx);
  // the reference's bubble sort (strict >): equal angles keep their order
  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (ang[i] > ang[i + 1]) {
        const float t = ang[i];
        ang[i] = ang[i + 1];
        ang[i + 1] = t;
        const P2 p = pts[i];
        pts[i] = pts[i + 1];
        pts[i + 1] = p;
      }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k)
    area += (pts[k].x - pts[0].x) * (pts[k + 1].y - pts[0].y) - (pts[k].y - pts[0].y) * (pts[k + 1].x - pts[0].x);
  return fabsf(area) / 2.0f;
}

__device__ __forceinline__ float iou_rotated(const float* __restrict__ a, const float* __restrict__ b) {
  const float so = overlap(a, b);
  return __fdiv_rn(so, fmaxf(a[14] + b[14] - so, kEps));
}

__device__ __forceinline__ float iou_normal(const float* __restrict__ a, const float* __restrict__ b) {
  // raw boxes [x, y, z, dx, dy, dz, heading] (iou3d_nms_kernel.cu:331-342): heading ignored
  const float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
  const float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
  const float inter = fmaxf(right - left, 0.f) * fmaxf(bottom - top, 0.f);
  return __fdiv_rn(inter, fmaxf(a[3] * a[4] + b[3] * b[4] - inter, kEps));
}

__global__ void __launch_bounds__(256) prepare_kernel(const float* __restrict__ boxes, int64_t n, float* __restrict__ prep) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) prepare(boxes + i * 7, prep + i * kPrep);
}

// out[i, j] for a 16 x 16 tile per CTA; mode 0 = IoU, 1 = overlap area.  Prepared boxes of the tile in shared memory.
__global__ void __launch_bounds__(256)
pair_kernel(const float* __restrict__ pa, int64_t na, const float* __restrict__ pb, int64_t nb, int mode, float* __restrict__ out) {
  __shared__ float sa[16 * kPrep], sb[16 * kPrep];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t i0 = static_cast<int64_t>(blockIdx.y) * 16, j0 = static_cast<int64_t>(blockIdx.x) * 16;
  {
    const int r = threadIdx.x >> 4, c = threadIdx.x & 15;  // 16 boxes x 16 floats
    sa[threadIdx.x] = (i0 + r < na) ? pa[(i0 + r) * kPrep + c] : 0.f;
    sb[threadIdx.x] = (j0 + r < nb) ? pb[(j0 + r) * kPrep + c] : 0.f;
  }
  __syncthreads();
  const int64_t i = i0 + ty, j = j0 + tx;
  if (i >= na || j >= nb) return;
  const float* a = sa + ty * kPrep;
  const float* b = sb + tx * kPrep;
  out[i * nb + j] = mode ? overlap(a, b) : iou_rotated(a, b);
}

// mask[i, cb] bit k = IoU(box i, box 64 cb + k) > thresh, for cb >= i / 64 only (and k > i % 64 on the diagonal block).
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float* __restrict__ prep, const float* __restrict__ raw, int n, float thresh, int normal, int col_blocks,
                unsigned long long* __restrict__ mask) {
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (cb < rb) return;
  __shared__ float cols[64 * kPrep];
  const int col_n = min(n - cb * 64, 64), row_n = min(n - rb * 64, 64);
  const int stride = normal ? 7 : kPrep;
  const float* src = normal ? raw : prep;
  for (int e = threadIdx.x; e < col_n * stride; e += 64) cols[e] = src[static_cast<int64_t>(cb) * 64 * stride + e];
  __syncthreads();
  if (static_cast<int>(threadIdx.x) >= row_n) return;
  const int i = rb * 64 + threadIdx.x;
  float mine[kPrep];
  for (int e = 0; e < stride; ++e) mine[e] = src[static_cast<int64_t>(i) * stride + e];
  unsigned long long bits = 0;
  const int start = rb == cb ? threadIdx.x + 1 : 0;
  for (int k = start; k < col_n; ++k) {
    const float v = normal ? iou_normal(mine, cols + k * 7) : iou_rotated(mine, cols + k * kPrep);
    if (v > thresh) bits |= 1ull << k;
  }
  mask[static_cast<int64_t>(i) * col_blocks + cb] = bits;
}

// Greedy scan in ONE CTA (blockDim.x >= col_blocks): thread t owns the suppression word of box block t.
__global__ void __launch_bounds__(1024)
nms_scan_kernel(const unsigned long long* __restrict__ mask, int n, int col_blocks, long long* __restrict__ keep,
                int* __restrict__ num_keep) {
  extern __shared__ unsigned long long sm[];
  unsigned long long* remv = sm;                  // [col_blocks]
  unsigned long long* kept = sm + col_blocks;     // [col_blocks] kept bits per block
  unsigned long long* diag = kept + col_blocks;   // [64] diagonal words of the current block
  __shared__ int offsets[1025];
  const int t = threadIdx.x;
  if (t < col_blocks) remv[t] = 0ull;
  __syncthreads();
  for (int b = 0; b < col_blocks; ++b) {
    const int bn = min(n - b * 64, 64);
    if (t < bn) diag[t] = mask[static_cast<int64_t>(b * 64 + t) * col_blocks + b];
    __syncthreads();
    if (t == 0) {  // resolve the block from its diagonal words: pure ALU, 64 steps
      unsigned long long r = remv[b], k = 0ull;
      for (int i = 0; i < bn; ++i)
        if (!((r >> i) & 1ull)) {
          k |= 1ull << i;
          r |= diag[i];
        }
      kept[b] = k;
    }
    __syncthreads();
    if (t > b && t < col_blocks) {  // later blocks: OR in the rows of the kept boxes (independent loads)
      unsigned long long k = kept[b], acc = 0ull;
      while (k) {
        const int i = __ffsll(static_cast<long long>(k)) - 1;
        k &= k - 1;
        acc |= mask[static_cast<int64_t>(b * 64 + i) * col_blocks + t];
      }
      remv[t] |= acc;
    }
    __syncthreads();
  }
  if (t == 0) {
    int off = 0;
    for (int b = 0; b < col_blocks; ++b) {
      offsets[b] = off;
      off += __popcll(kept[b]);
    }
    offsets[col_blocks] = off;
    *num_keep = off;
  }
  __syncthreads();
  if (t < col_blocks) {
    unsigned long long k = kept[t];
    int o = offsets[t];
    while (k) {
      const int i = __ffsll(static_cast<long long>(k)) - 1;
      k &= k - 1;
      keep[o++] = static_cast<long long>(t) * 64 + i;
    }
  }
}

}  // namespace iou3d
}  // namespace efgb

using namespace efgb;

extern "C" size_t efgb_boxes_bev_workspace_bytes(int64_t num_a, int64_t num_b) {
  return align_up(static_cast<size_t>(num_a) * iou3d::kPrep * sizeof(float)) + align_up(static_cast<size_t>(num_b) * iou3d::kPrep * sizeof(float)) + 256;
}

extern "C" int efgb_boxes_bev(const float* boxes_a, int64_t num_a, const float* boxes_b, int64_t num_b, int mode, float* out,
                              void* workspace, size_t workspace_bytes, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(num_a >= 0 && num_b >= 0 && (mode == 0 || mode == 1), EFGB_EINVAL, "boxes_bev: bad argument");
  if (num_a == 0 || num_b == 0) return EFGB_OK;
  EFGB_REQUIRE(boxes_a && boxes_b && out && workspace, EFGB_EINVAL, "boxes_bev: null pointer");
  Workspace ws(workspace, workspace_bytes);
  float* pa = ws.take<float>(static_cast<size_t>(num_a) * iou3d::kPrep);
  float* pb = ws.take<float>(static_cast<size_t>(num_b) * iou3d::kPrep);
  EFGB_REQUIRE(pa && pb, EFGB_EINVAL, "boxes_bev: workspace too small");
  iou3d::prepare_kernel<<<grid_for(num_a, 256), 256, 0, stream>>>(boxes_a, num_a, pa);
  EFGB_LAUNCH_OK("iou3d::prepare_kernel");
  iou3d::prepare_kernel<<<grid_for(num_b, 256), 256, 0, stream>>>(boxes_b, num_b, pb);
  EFGB_LAUNCH_OK("iou3d::prepare_kernel");
  const dim3 grid(static_cast<unsigned>((num_b + 15) / 16), static_cast<unsigned>((num_a + 15) / 16));
  iou3d::pair_kernel<<<grid, 256, 0, stream>>>(pa, num_a, pb, num_b, mode, out);
  EFGB_LAUNCH_OK("iou3d::pair_kernel");
  return EFGB_OK;
}

extern "C" size_t efgb_nms_bev_workspace_bytes(int64_t n) {
  const size_t cb = static_cast<size_t>((n + 63) / 64);
  return align_up(static_cast<size_t>(n) * iou3d::kPrep * sizeof(float)) + align_up(static_cast<size_t>(n) * cb * sizeof(unsigned long long)) + 256;
}

extern "C" int efgb_nms_bev(const float* boxes_sorted, int64_t n, float thresh, int normal, int64_t* keep, int32_t* num_keep,
                            void* workspace, size_t workspace_bytes, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(n >= 0 && n <= 65536, EFGB_EINVAL, "nms_bev: at most 65536 boxes (got %lld)", static_cast<long long>(n));
  EFGB_REQUIRE(num_keep, EFGB_EINVAL, "nms_bev: null pointer");
  if (n == 0) {
    EFGB_CUDA_OK(cudaMemsetAsync(num_keep, 0, sizeof(int32_t), stream));
    return EFGB_OK;
  }
  EFGB_REQUIRE(boxes_sorted && keep && workspace, EFGB_EINVAL, "nms_bev: null pointer");
  const int cb = static_cast<int>((n + 63) / 64);
  Workspace ws(workspace, workspace_bytes);
  float* prep = ws.take<float>(static_cast<size_t>(n) * iou3d::kPrep);
  unsigned long long* mask = ws.take<unsigned long long>(static_cast<size_t>(n) * cb);
  EFGB_REQUIRE(prep && mask, EFGB_EINVAL, "nms_bev: workspace too small");
  if (!normal) {
    iou3d::prepare_kernel<<<grid_for(n, 256), 256, 0, stream>>>(boxes_sorted, n, prep);
    EFGB_LAUNCH_OK("iou3d::prepare_kernel");
  }
  iou3d::nms_mask_kernel<<<dim3(cb, cb), 64, 0, stream>>>(prep, boxes_sorted, static_cast<int>(n), thresh, normal ? 1 : 0, cb, mask);
  EFGB_LAUNCH_OK("iou3d::nms_mask_kernel");
  int threads = 64;
  while (threads < cb) threads <<= 1;
  const size_t smem = (2 * static_cast<size_t>(cb) + 64) * sizeof(unsigned long long);
  iou3d::nms_scan_kernel<<<1, threads, smem, stream>>>(mask, static_cast<int>(n), cb, reinterpret_cast<long long*>(keep), num_keep);
  EFGB_LAUNCH_OK("iou3d::nms_scan_kernel");
  return EFGB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// CenterPoint label assignment, the heatmap part (CP/voxelnet.py:44-192 -> CP/center_utils.py:29-58 draw_gaussian):
// one CTA per object splats its Gaussian (sigma = (2 r + 1) / 6, clipped to the map) into heatmap[scene, class] with
// max() — order independent, so all objects of the batch are drawn concurrently with an integer atomicMax on the
// (non-negative) float bits.  The reference draws them one after another in numpy on the host and uploads the maps.
// ------------------------------------------------------------------------------------------------------------------
namespace efgb {
__global__ void __launch_bounds__(128)
draw_gaussians_kernel(const int32_t* __restrict__ obj, int num_obj, int height, int width, float* __restrict__ heatmaps) {
  // obj[i] = (plane index, x, y, radius): plane = flattened (scene, class) map
  const int i = blockIdx.x;
  if (i >= num_obj) return;
  const int plane = obj[4 * i], cx = obj[4 * i + 1], cy = obj[4 * i + 2], r = obj[4 * i + 3];
  const int d = 2 * r + 1;
  const double sigma = static_cast<double>(d) / 6.0;
  const double inv = 1.0 / (2.0 * sigma * sigma);
  float* hm = heatmaps + static_cast<int64_t>(plane) * height * width;
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
    const int dy = e / d - r, dx = e % d - r;
    const int y = cy + dy, x = cx + dx;
    if (x < 0 || x >= width || y < 0 || y >= height) continue;
    // numpy evaluates the Gaussian in float64 and stores into the float32 map
    const float v = static_cast<float>(exp(-static_cast<double>(dx * dx + dy * dy) * inv));
    atomicMax(reinterpret_cast<int*>(hm + static_cast<int64_t>(y) * width + x), __float_as_int(v));
  }
}
}  // namespace efgb

extern "C" int efgb_draw_gaussians(const int32_t* objects, int num_objects, int height, int width, float* heatmaps,
                                   efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(num_objects >= 0 && height > 0 && width > 0, EFGB_EINVAL, "draw_gaussians: bad argument");
  if (num_objects == 0) return EFGB_OK;
  EFGB_REQUIRE(objects && heatmaps, EFGB_EINVAL, "draw_gaussians: null pointer");
  draw_gaussians_kernel<<<num_objects, 128, 0, stream>>>(objects, num_objects, height, width, heatmaps);
  EFGB_LAUNCH_OK("draw_gaussians_kernel");
  return EFGB_OK;
}
