// Hash-and-scatter hard voxelizer with fused mean-VFE (sm_100a).
//
// Semantics follow the reference's serial CPU loop bit for bit
// (efg/geometry/point_cloud_ops.py:31-53, voxelization_cpu.cpp:44-96):
//   voxel id  = rank of the voxel's FIRST point among all first points (input order),
//   slot      = rank of the point among the points of its voxel (input order), kept if < max_points,
//   cut-off   = the first point that would open voxel #max_voxels ends the scene.
// The serial dependency is replaced by: (1) an open-addressing hash on the linear cell id that
// keeps, per voxel, the minimum point index (atomicMin) and a lock-free list of its points
// (atomicExch), (2) a device-wide exclusive scan over "is first point" flags, which IS the
// first-come voxel numbering, (3) one thread per voxel that selects its max_points smallest point
// indices in ascending order and writes the padded voxel row, the count and the mean.
#include "common.cuh"

namespace efgb {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr int kMaxBatch = 64;

struct VoxGeom {
  float vs[3];
  float lo[3];
  int grid[3];  // x, y, z
};

struct SceneTable {
  int batch;
};

__device__ __forceinline__ uint32_t hash_u32(uint32_t k) {
  k ^= k >> 16;
  k *= 0x85ebca6bu;
  k ^= k >> 13;
  k *= 0xc2b2ae35u;
  k ^= k >> 16;
  return k;
}

// fp32 true division + floor, exactly like the reference (no fast-math, no reciprocal).
__device__ __forceinline__ bool point_to_cell(const float* __restrict__ p, const VoxGeom& g, int c[3]) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float f = floorf(__fdiv_rn(__fsub_rn(p[j], g.lo[j]), g.vs[j]));
    if (!(f >= 0.0f) || !(f < static_cast<float>(g.grid[j]))) return false;  // also rejects NaN
    c[j] = static_cast<int>(f);
  }
  return true;
}

// The xyz of the CTA's 256 consecutive points, read as coalesced 128-bit loads: the CTA's slice of the point array is one
// contiguous block of 256 * nfeat floats (16-byte aligned when the array is), copied to shared memory with LDG.128 and
// read back per point (row stride nfeat words).  A 5-float row is not 16-byte aligned, so per-point vector loads are not
// possible; per-point scalar loads at a 20-byte stride touch every sector five times.
constexpr int kStageFeat = 8;   // rows wider than this (or an unaligned array) take the scalar path
__device__ __forceinline__ const float* stage_points(const float* __restrict__ points, int64_t n, int nfeat, float* s_pts,
                                                     bool aligned) {
  const int64_t base = static_cast<int64_t>(blockIdx.x) * blockDim.x;
  if (!aligned || nfeat > kStageFeat) {
    __syncthreads();
    return points + (base + threadIdx.x) * nfeat;
  }
  const int64_t left = n - base;
  const int total = static_cast<int>(left < static_cast<int64_t>(blockDim.x) ? left : blockDim.x) * nfeat;
  const float* src = points + base * nfeat;
  for (int e = threadIdx.x * 4; e < total; e += blockDim.x * 4) {
    if (e + 3 < total) {
      *reinterpret_cast<float4*>(s_pts + e) = __ldg(reinterpret_cast<const float4*>(src + e));
    } else {
      for (int k = e; k < total; ++k) s_pts[k] = __ldg(src + k);
    }
  }
  __syncthreads();
  return s_pts + threadIdx.x * nfeat;
}

__device__ __forceinline__ int scene_of(const int32_t* __restrict__ offs, int batch, int64_t i) {
  int b = 0;
  while (b + 1 < batch && i >= offs[b + 1]) ++b;
  return b;
}

__global__ void __launch_bounds__(256)
vox_insert_kernel(const float* __restrict__ points, int64_t n, int nfeat, const int32_t* __restrict__ offs, int batch,
                  VoxGeom g, uint32_t* __restrict__ pt_slot, uint32_t* tkeys, uint32_t* tfirst, uint32_t* thead,
                  uint32_t* __restrict__ pt_next, uint32_t mask, int aligned) {
  __shared__ int32_t s_offs[kMaxBatch + 1];
  __shared__ __align__(16) float s_pts[256 * kStageFeat];
  for (int t = threadIdx.x; t <= batch; t += blockDim.x) s_offs[t] = offs[t];
  const float* mine = stage_points(points, n, nfeat, s_pts, aligned != 0);   // ends with __syncthreads()
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c[3];
  if (!point_to_cell(mine, g, c)) {
    pt_slot[i] = kEmpty;
    return;
  }
  const int b = scene_of(s_offs, batch, i);
  const uint32_t key =
      ((static_cast<uint32_t>(b) * g.grid[2] + c[2]) * g.grid[1] + c[1]) * static_cast<uint32_t>(g.grid[0]) + c[0];
  uint32_t h = hash_u32(key) & mask;
  while (true) {
    uint32_t old = atomicCAS(&tkeys[h], kEmpty, key);
    if (old == kEmpty || old == key) break;
    h = (h + 1) & mask;
  }
  atomicMin(&tfirst[h], static_cast<uint32_t>(i));
  pt_next[i] = atomicExch(&thead[h], static_cast<uint32_t>(i));
  pt_slot[i] = h;
}

__global__ void __launch_bounds__(256)
vox_flag_kernel(const uint32_t* __restrict__ pt_slot, const uint32_t* __restrict__ tfirst, int64_t n,
                uint32_t* __restrict__ flags) {
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = pt_slot[i];
  flags[i] = (s != kEmpty && tfirst[s] == static_cast<uint32_t>(i)) ? 1u : 0u;
}

// One thread: per-scene bookkeeping.  meta layout: [0..B) rank_base, [B..2B) out_base, [2B..3B) cutoff.
__global__ void vox_scene_kernel(const uint32_t* __restrict__ excl, const int32_t* __restrict__ offs, int batch,
                                 int max_voxels, int32_t* meta, int32_t* voxel_counts) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int out = 0;
  for (int b = 0; b < batch; ++b) {
    int start = static_cast<int>(excl[offs[b]]);
    int end = static_cast<int>(excl[offs[b + 1]]);
    int cnt = end - start;
    int kept = (max_voxels >= 0 && cnt > max_voxels) ? max_voxels : cnt;
    meta[b] = start;
    meta[batch + b] = out;
    meta[2 * batch + b] = 0x7FFFFFFF;
    voxel_counts[b] = kept;
    out += kept;
  }
  voxel_counts[batch] = out;
}

__global__ void __launch_bounds__(256)
vox_cutoff_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ excl,
                  const int32_t* __restrict__ offs, int batch, int max_voxels, int64_t n, int32_t* meta) {
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || !flags[i]) return;
  const int b = scene_of(offs, batch, i);
  if (static_cast<int>(excl[i]) - meta[b] == max_voxels) meta[2 * batch + b] = static_cast<int32_t>(i);
}

__global__ void __launch_bounds__(128)
vox_emit_kernel(const float* __restrict__ points, int64_t n, int nfeat, const int32_t* __restrict__ offs, int batch,
                VoxGeom g, const uint32_t* __restrict__ flags, const uint32_t* __restrict__ excl,
                const uint32_t* __restrict__ pt_slot, const uint32_t* __restrict__ thead,
                const uint32_t* __restrict__ pt_next, const int32_t* __restrict__ meta, int max_points,
                int max_voxels, float* __restrict__ voxels, int32_t* __restrict__ coors, int coors_dim,
                int32_t* __restrict__ npv, float* __restrict__ mean) {
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || !flags[i]) return;
  const int b = scene_of(offs, batch, i);
  const int r = static_cast<int>(excl[i]) - meta[b];
  if (max_voxels >= 0 && r >= max_voxels) return;
  const int64_t vid = meta[batch + b] + r;
  const uint32_t cutoff = static_cast<uint32_t>(meta[2 * batch + b]);

  int c[3];
  point_to_cell(points + i * nfeat, g, c);
  int32_t* co = coors + vid * coors_dim;
  if (coors_dim == 4) *co++ = b;
  co[0] = c[2];
  co[1] = c[1];
  co[2] = c[0];

  constexpr int kMaxFeat = 16;
  float sum[kMaxFeat];
#pragma unroll
  for (int f = 0; f < kMaxFeat; ++f) sum[f] = 0.f;

  const uint32_t head = thead[pt_slot[i]];
  int cnt = 0;
  int64_t last = -1;
  for (int s = 0; s < max_points; ++s) {
    uint32_t best = kEmpty;
    for (uint32_t j = head; j != kEmpty; j = pt_next[j]) {
      if (static_cast<int64_t>(j) > last && j < cutoff && j < best) best = j;
    }
    if (best == kEmpty) break;
    const float* src = points + static_cast<int64_t>(best) * nfeat;
    float* dst = voxels ? voxels + (vid * max_points + s) * nfeat : nullptr;
    for (int f = 0; f < nfeat; ++f) {
      float v = src[f];
      if (dst) dst[f] = v;
      if (f < kMaxFeat) sum[f] += v;
    }
    last = best;
    ++cnt;
  }
  if (voxels) {
    for (int s = cnt; s < max_points; ++s) {
      float* dst = voxels + (vid * max_points + s) * nfeat;
      for (int f = 0; f < nfeat; ++f) dst[f] = 0.f;
    }
  }
  npv[vid] = cnt;
  if (mean) {
    const float denom = static_cast<float>(cnt);
    for (int f = 0; f < nfeat && f < kMaxFeat; ++f) mean[vid * nfeat + f] = __fdiv_rn(sum[f], denom);
  }
}

__global__ void __launch_bounds__(256)
dynamic_voxelize_kernel(const float* __restrict__ points, int64_t n, int nfeat, VoxGeom g, int32_t* __restrict__ coors,
                        int aligned) {
  __shared__ __align__(16) float s_pts[256 * kStageFeat];
  const float* mine = stage_points(points, n, nfeat, s_pts, aligned != 0);
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c[3];
  bool ok = point_to_cell(mine, g, c);
  coors[i * 3 + 0] = ok ? c[2] : -1;
  coors[i * 3 + 1] = ok ? c[1] : -1;
  coors[i * 3 + 2] = ok ? c[0] : -1;
}

static uint32_t table_size_for(int64_t n) {
  uint64_t t = 1024;
  while (t < static_cast<uint64_t>(n) * 2) t <<= 1;
  return static_cast<uint32_t>(t);
}

static int make_geom(const float* vs, const float* range, VoxGeom* g) {
  for (int j = 0; j < 3; ++j) {
    g->vs[j] = vs[j];
    g->lo[j] = range[j];
    // grid = round((hi - lo) / vs) in fp32 (voxelization_cpu.cpp:119-122, point_cloud_ops.py:138-139)
    float q = (range[3 + j] - range[j]) / vs[j];
    g->grid[j] = static_cast<int>(roundf(q));
    if (g->grid[j] <= 0) return -1;
  }
  return 0;
}

}  // namespace efgb

using namespace efgb;

extern "C" size_t efgb_voxelize_workspace_bytes(int64_t num_points, int batch) {
  if (num_points < 0) num_points = 0;
  const size_t t = table_size_for(num_points);
  size_t bytes = 0;
  bytes += 3 * align_up(t * sizeof(uint32_t));                             // keys, first, head
  bytes += 2 * align_up(static_cast<size_t>(num_points) * sizeof(uint32_t));  // pt_slot, pt_next
  bytes += align_up(static_cast<size_t>(num_points) * sizeof(uint32_t));      // flags
  bytes += align_up(static_cast<size_t>(num_points + 1) * sizeof(uint32_t));  // excl
  bytes += align_up(scan_scratch_elems(num_points) * sizeof(uint32_t));
  bytes += align_up(static_cast<size_t>(3 * (batch > 0 ? batch : 1)) * sizeof(int32_t));
  return bytes + 1024;
}

extern "C" int efgb_hard_voxelize(const float* points, int64_t num_points, int num_features,
                                  const int32_t* scene_offsets, int batch, const float* voxel_size,
                                  const float* coors_range, int max_points, int max_voxels, float* voxels,
                                  int32_t* coors, int coors_dim, int32_t* num_points_per_voxel, float* mean_features,
                                  int32_t* voxel_counts, void* workspace, size_t workspace_bytes,
                                  efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(num_points >= 0 && num_points < (1ll << 31) - 1, EFGB_EINVAL, "hard_voxelize: bad num_points %lld",
               (long long)num_points);
  EFGB_REQUIRE(batch >= 1 && batch <= kMaxBatch, EFGB_EINVAL, "hard_voxelize: batch must be in [1,%d]", kMaxBatch);
  EFGB_REQUIRE(num_features >= 3, EFGB_EINVAL, "hard_voxelize: points need >= 3 features");
  EFGB_REQUIRE(!mean_features || num_features <= 16, EFGB_EINVAL, "hard_voxelize: fused mean supports <= 16 features");
  EFGB_REQUIRE(max_points >= 1, EFGB_EINVAL, "hard_voxelize: max_points must be >= 1 (use dynamic_voxelize for -1)");
  EFGB_REQUIRE(coors_dim == 3 || coors_dim == 4, EFGB_EINVAL, "hard_voxelize: coors_dim must be 3 or 4");
  EFGB_REQUIRE(scene_offsets && coors && num_points_per_voxel && voxel_counts && voxel_size && coors_range, EFGB_EINVAL,
               "hard_voxelize: null argument");
  EFGB_REQUIRE(points || num_points == 0, EFGB_EINVAL, "hard_voxelize: null points");
  VoxGeom g;
  EFGB_REQUIRE(make_geom(voxel_size, coors_range, &g) == 0, EFGB_EINVAL, "hard_voxelize: empty grid");
  const double cells = static_cast<double>(batch) * g.grid[0] * g.grid[1] * g.grid[2];
  EFGB_REQUIRE(cells < 4294967295.0, EFGB_ERANGE, "hard_voxelize: batch*grid = %.0f cells exceeds 32-bit cell ids", cells);
  EFGB_REQUIRE(workspace_bytes >= efgb_voxelize_workspace_bytes(num_points, batch), EFGB_EWORKSPACE,
               "hard_voxelize: workspace too small");

  Workspace ws(workspace, workspace_bytes);
  const uint32_t tsize = table_size_for(num_points);
  uint32_t* tkeys = ws.take<uint32_t>(tsize);
  uint32_t* tfirst = ws.take<uint32_t>(tsize);
  uint32_t* thead = ws.take<uint32_t>(tsize);
  uint32_t* pt_slot = ws.take<uint32_t>(num_points);
  uint32_t* pt_next = ws.take<uint32_t>(num_points);
  uint32_t* flags = ws.take<uint32_t>(num_points);
  uint32_t* excl = ws.take<uint32_t>(num_points + 1);
  uint32_t* scratch = ws.take<uint32_t>(scan_scratch_elems(num_points));
  int32_t* meta = ws.take<int32_t>(3 * batch);
  EFGB_REQUIRE(meta != nullptr, EFGB_EWORKSPACE, "hard_voxelize: workspace too small");

  // keys/first/head are contiguous (each aligned_up) -> one 0xFF fill.
  EFGB_CUDA_OK(cudaMemsetAsync(tkeys, 0xFF, 3 * align_up(tsize * sizeof(uint32_t)), stream));
  if (num_points > 0) {
    const unsigned nb = static_cast<unsigned>((num_points + 255) / 256);
    vox_insert_kernel<<<nb, 256, 0, stream>>>(points, num_points, num_features, scene_offsets, batch, g, pt_slot, tkeys,
                                              tfirst, thead, pt_next, tsize - 1,
                                              (reinterpret_cast<uintptr_t>(points) & 15) == 0 ? 1 : 0);
    EFGB_LAUNCH_OK("vox_insert_kernel");
    vox_flag_kernel<<<nb, 256, 0, stream>>>(pt_slot, tfirst, num_points, flags);
    EFGB_LAUNCH_OK("vox_flag_kernel");
  }
  int rc = scan_exclusive_u32(flags, excl, num_points, scratch, stream);
  if (rc != EFGB_OK) return rc;
  vox_scene_kernel<<<1, 32, 0, stream>>>(excl, scene_offsets, batch, max_voxels, meta, voxel_counts);
  EFGB_LAUNCH_OK("vox_scene_kernel");
  if (num_points > 0) {
    const unsigned nb = static_cast<unsigned>((num_points + 255) / 256);
    if (max_voxels >= 0) {
      vox_cutoff_kernel<<<nb, 256, 0, stream>>>(flags, excl, scene_offsets, batch, max_voxels, num_points, meta);
      EFGB_LAUNCH_OK("vox_cutoff_kernel");
    }
    const unsigned nb2 = static_cast<unsigned>((num_points + 127) / 128);
    vox_emit_kernel<<<nb2, 128, 0, stream>>>(points, num_points, num_features, scene_offsets, batch, g, flags, excl,
                                             pt_slot, thead, pt_next, meta, max_points, max_voxels, voxels, coors,
                                             coors_dim, num_points_per_voxel, mean_features);
    EFGB_LAUNCH_OK("vox_emit_kernel");
  }
  return EFGB_OK;
}

extern "C" int efgb_dynamic_voxelize(const float* points, int64_t num_points, int num_features, const float* voxel_size,
                                     const float* coors_range, int32_t* coors, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(num_points >= 0 && num_features >= 3 && voxel_size && coors_range, EFGB_EINVAL,
               "dynamic_voxelize: bad argument");
  EFGB_REQUIRE((points && coors) || num_points == 0, EFGB_EINVAL, "dynamic_voxelize: null pointer");
  VoxGeom g;
  EFGB_REQUIRE(make_geom(voxel_size, coors_range, &g) == 0, EFGB_EINVAL, "dynamic_voxelize: empty grid");
  if (num_points == 0) return EFGB_OK;
  const unsigned nb = static_cast<unsigned>((num_points + 255) / 256);
  dynamic_voxelize_kernel<<<nb, 256, 0, stream>>>(points, num_points, num_features, g, coors,
                                                  (reinterpret_cast<uintptr_t>(points) & 15) == 0 ? 1 : 0);
  EFGB_LAUNCH_OK("dynamic_voxelize_kernel");
  return EFGB_OK;
}
