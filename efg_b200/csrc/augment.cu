// Point-cloud augmentation on the device: the reference's per-sample processor chain
//     RandomFlip3D -> GlobalRotation -> GlobalScaling -> GlobalTranslation -> FilterByRange
// (efg/data/augmentations/extend_3d.py:121-316, efg/geometry/box_ops.py:517-548) applied to the points in two passes
// instead of five numpy / torch round trips over the cloud in a DataLoader worker:
//   flags pass   transform x, y, z in registers, range test -> 0/1 flag per point
//   (device-wide exclusive scan of the flags: order-preserving compaction, as `points[keep]` is)
//   emit pass    transform again, write the kept points to their compacted position; writes the count
// The random draws (flip decisions, angle, scale, translation) are made by the host exactly as the reference makes
// them (np.random, same call order) and passed in; the arithmetic is the reference's, step by step, in fp32.
#include "common.cuh"

namespace efgb {
namespace aug {

struct Xform {
  int flip_x, flip_y;        // RandomFlip3D: y -> -y, then x -> -x
  float cosa, sina;          // GlobalRotation: [x y] <- [x cos - y sin, x sin + y cos]
  float scale;               // GlobalScaling
  float tx, ty, tz;          // GlobalTranslation
  float lo[3], hi[3];        // FilterByRange (inclusive bounds)
  int filter;
};

__device__ __forceinline__ void apply(const Xform& t, float& x, float& y, float& z) {
  if (t.flip_x) y = -y;
  if (t.flip_y) x = -x;
  // torch.matmul(points[:, :3], rot) with rot = [[cos, sin, 0], [-sin, cos, 0], [0, 0, 1]]
  const float xr = __fadd_rn(__fmul_rn(x, t.cosa), __fmul_rn(y, -t.sina));
  const float yr = __fadd_rn(__fmul_rn(x, t.sina), __fmul_rn(y, t.cosa));
  x = __fmul_rn(xr, t.scale) + t.tx;
  y = __fmul_rn(yr, t.scale) + t.ty;
  z = __fmul_rn(z, t.scale) + t.tz;
}

__device__ __forceinline__ bool inside(const Xform& t, float x, float y, float z) {
  return !t.filter || (x >= t.lo[0] && x <= t.hi[0] && y >= t.lo[1] && y <= t.hi[1] && z >= t.lo[2] && z <= t.hi[2]);
}

__global__ void __launch_bounds__(256) flags_kernel(const float* __restrict__ pts, int64_t n, int nfeat, Xform t, uint32_t* __restrict__ flags) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = pts[i * nfeat], y = pts[i * nfeat + 1], z = pts[i * nfeat + 2];
  apply(t, x, y, z);
  flags[i] = inside(t, x, y, z) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
emit_kernel(const float* __restrict__ pts, int64_t n, int nfeat, Xform t, const uint32_t* __restrict__ pos, float* __restrict__ out,
            int32_t* __restrict__ count) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i == 0) *count = static_cast<int32_t>(pos[n]);
  if (i >= n) return;
  if (pos[i + 1] == pos[i]) return;  // dropped
  float x = pts[i * nfeat], y = pts[i * nfeat + 1], z = pts[i * nfeat + 2];
  apply(t, x, y, z);
  float* o = out + static_cast<int64_t>(pos[i]) * nfeat;
  o[0] = x;
  o[1] = y;
  o[2] = z;
  for (int f = 3; f < nfeat; ++f) o[f] = pts[i * nfeat + f];
}


// GT-database paste (DatabaseSampling.__call__, efg/data/augmentations/extend_3d.py:68-92): the output is
// [points of the accepted database objects, translated to their boxes | scene points].  `table[k]` = (first point of
// object k in the resident database, first output row, point count); `planes` (optional, rm_points_after_sample) are the
// inward-pointing face planes of the pasted boxes as the reference builds them (box_ops.py:285-310): a scene point inside
// any box (all six signs < 0, box_ops.py:356-371, evaluated left to right without contraction) is not compacted away but
// moved out of every detection range (kPasteDropped), where the voxelizer drops it — same voxels, no count to read back.
constexpr float kPasteDropped = 1e30f;

__global__ void __launch_bounds__(256)
paste_points_kernel(const float* __restrict__ db, const int32_t* __restrict__ table, const float* __restrict__ centers, int num_obj,
                    int64_t n_paste, const float* __restrict__ scene, int64_t n_scene, int nfeat,
                    const float* __restrict__ planes, int num_rm, float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_paste + n_scene) return;
  float* dst = out + i * nfeat;
  if (i < n_paste) {
    int lo = 0, hi = num_obj - 1;
    while (lo < hi) {   // last object whose first output row is <= i
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(table + mid * 3 + 1) <= i) lo = mid; else hi = mid - 1;
    }
    const float* src = db + (static_cast<int64_t>(__ldg(table + lo * 3)) + (i - __ldg(table + lo * 3 + 1))) * nfeat;
    for (int f = 0; f < nfeat; ++f) dst[f] = f < 3 ? __fadd_rn(__ldg(src + f), __ldg(centers + lo * 3 + f)) : __ldg(src + f);
    return;
  }
  const float* src = scene + (i - n_paste) * nfeat;
  float x = __ldg(src), y = __ldg(src + 1), z = __ldg(src + 2);
  bool inside_any = false;
  for (int m = 0; m < num_rm && !inside_any; ++m) {
    bool inside = true;
    for (int k = 0; k < 6 && inside; ++k) {
      const float4 pl = __ldg(reinterpret_cast<const float4*>(planes) + m * 6 + k);
      const float sign = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, pl.x), __fmul_rn(y, pl.y)), __fmul_rn(z, pl.z)), pl.w);
      inside = !(sign >= 0.f);
    }
    inside_any = inside;
  }
  if (inside_any) x = y = z = kPasteDropped;
  dst[0] = x;
  dst[1] = y;
  dst[2] = z;
  for (int f = 3; f < nfeat; ++f) dst[f] = __ldg(src + f);
}

}  // namespace aug
}  // namespace efgb

using namespace efgb;

extern "C" size_t efgb_augment_workspace_bytes(int64_t n) {
  return align_up(static_cast<size_t>(n + 1) * sizeof(uint32_t)) * 2 + align_up(scan_scratch_elems(n) * sizeof(uint32_t)) + 256;
}

extern "C" int efgb_augment_points(const float* points, int64_t n, int nfeat, int flip_x, int flip_y, float cosa, float sina,
                                   float scale, const float* translation_host3, const float* range_host6, float* out_points,
                                   int32_t* out_count, void* workspace, size_t workspace_bytes, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(n >= 0 && nfeat >= 3 && out_count, EFGB_EINVAL, "augment_points: bad argument");
  if (n == 0) {
    EFGB_CUDA_OK(cudaMemsetAsync(out_count, 0, sizeof(int32_t), stream));
    return EFGB_OK;
  }
  EFGB_REQUIRE(points && out_points && workspace, EFGB_EINVAL, "augment_points: null pointer");
  aug::Xform t;
  t.flip_x = flip_x ? 1 : 0;
  t.flip_y = flip_y ? 1 : 0;
  t.cosa = cosa;
  t.sina = sina;
  t.scale = scale;
  t.tx = translation_host3 ? translation_host3[0] : 0.f;
  t.ty = translation_host3 ? translation_host3[1] : 0.f;
  t.tz = translation_host3 ? translation_host3[2] : 0.f;
  t.filter = range_host6 ? 1 : 0;
  for (int k = 0; k < 3; ++k) {
    t.lo[k] = range_host6 ? range_host6[k] : 0.f;
    t.hi[k] = range_host6 ? range_host6[3 + k] : 0.f;
  }
  Workspace ws(workspace, workspace_bytes);
  uint32_t* flags = ws.take<uint32_t>(static_cast<size_t>(n + 1));
  uint32_t* pos = ws.take<uint32_t>(static_cast<size_t>(n + 1));
  uint32_t* scratch = ws.take<uint32_t>(scan_scratch_elems(n));
  EFGB_REQUIRE(flags && pos && scratch, EFGB_EINVAL, "augment_points: workspace too small");
  aug::flags_kernel<<<grid_for(n, 256, 1 << 30), 256, 0, stream>>>(points, n, nfeat, t, flags);
  EFGB_LAUNCH_OK("aug::flags_kernel");
  const int rc = scan_exclusive_u32(flags, pos, n, scratch, stream);
  if (rc != EFGB_OK) return rc;
  aug::emit_kernel<<<grid_for(n, 256, 1 << 30), 256, 0, stream>>>(points, n, nfeat, t, pos, out_points, out_count);
  EFGB_LAUNCH_OK("aug::emit_kernel");
  return EFGB_OK;
}

extern "C" int efgb_paste_points(const float* db_points, const int32_t* obj_table, const float* obj_centers, int num_obj,
                                 int64_t n_paste, const float* scene_points, int64_t n_scene, int nfeat, const float* planes,
                                 int num_rm_boxes, float* out_points, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(num_obj >= 0 && n_paste >= 0 && n_scene >= 0 && nfeat >= 3 && num_rm_boxes >= 0, EFGB_EINVAL, "paste_points: bad argument");
  const int64_t total = n_paste + n_scene;
  if (total == 0) return EFGB_OK;
  EFGB_REQUIRE(out_points && (n_scene == 0 || scene_points) && (n_paste == 0 || (db_points && obj_table && obj_centers && num_obj > 0)) &&
                   (num_rm_boxes == 0 || (planes && (reinterpret_cast<uintptr_t>(planes) & 15) == 0)),
               EFGB_EINVAL, "paste_points: null or misaligned pointer");
  aug::paste_points_kernel<<<grid_for(total, 256, 1 << 30), 256, 0, stream>>>(db_points, obj_table, obj_centers, num_obj, n_paste,
                                                                             scene_points, n_scene, nfeat, planes, num_rm_boxes,
                                                                             out_points);
  EFGB_LAUNCH_OK("aug::paste_points_kernel");
  return EFGB_OK;
}
