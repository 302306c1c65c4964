// Dynamic point-to-voxel scatter (sum / mean / max) and its backward.
//
// Reference: efg/operators/src/voxelize/scatter_points_cuda.cu:209-290 (forward: linear id ->
// argsort -> run boundaries -> cumsum -> atomic reduce) and :292-352 (backward).  The sort is
// replaced by the occupancy-bitmap rank used by the rulebook builder: marking the occupied cells
// and scanning their popcounts yields, for every point, the index of its voxel in ascending
// linear-id order — the same output order the reference gets from argsort + cumsum.
#include <math.h>

#include "common.cuh"

namespace efgb {

struct Dims3 {
  int d0, d1, d2;
};

__device__ __forceinline__ bool cell3(const int32_t* __restrict__ c, const Dims3& d, uint32_t* cell) {
  const int a = c[0], b = c[1], e = c[2];
  if (a < 0 || b < 0 || e < 0 || a >= d.d0 || b >= d.d1 || e >= d.d2) return false;
  *cell = (static_cast<uint32_t>(a) * d.d1 + b) * static_cast<uint32_t>(d.d2) + e;
  return true;
}

__global__ void __launch_bounds__(256)
scatter_mark_kernel(const int32_t* __restrict__ coors, int64_t n, Dims3 d, CellWord* cells) {
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t cell;
  if (cell3(coors + i * 3, d, &cell)) atomicOr(&cells[cell >> 5].bits, 1u << (cell & 31));
}

__global__ void __launch_bounds__(256)
scatter_coords_kernel(const CellWord* __restrict__ cells, int64_t num_words, Dims3 d, int32_t* __restrict__ out) {
  int64_t w = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w >= num_words) return;
  CellWord cw = cells[w];
  uint32_t bits = cw.bits, r = cw.prefix;
  while (bits) {
    int bit = __ffs(bits) - 1;
    bits &= bits - 1;
    uint32_t cell = static_cast<uint32_t>(w) * 32u + bit;
    int32_t* o = out + static_cast<int64_t>(r) * 3;
    o[2] = cell % d.d2;
    uint32_t q = cell / d.d2;
    o[1] = q % d.d1;
    o[0] = q / d.d1;
    ++r;
  }
}

__global__ void __launch_bounds__(256) fill_kernel(float* p, int64_t n, float v) {
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v + 0.f));  // +0.f folds -0.0 into +0.0
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// One thread per (point, channel).
__global__ void __launch_bounds__(256)
scatter_reduce_kernel(const float* __restrict__ feats, const int32_t* __restrict__ coors, int64_t n, int channels,
                      Dims3 d, const CellWord* __restrict__ cells, int reduce_type, float* __restrict__ voxel_feats,
                      int32_t* __restrict__ point2voxel, int32_t* __restrict__ count) {
  int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n * channels) return;
  const int64_t i = t / channels;
  const int c = static_cast<int>(t - i * channels);
  uint32_t cell;
  int v = -1;
  if (cell3(coors + i * 3, d, &cell)) v = cell_rank(cells, cell);
  if (c == 0) {
    point2voxel[i] = v;
    if (v >= 0 && reduce_type == 1) atomicAdd(&count[v], 1);
  }
  if (v < 0) return;
  const float f = feats[t];
  float* dst = voxel_feats + static_cast<int64_t>(v) * channels + c;
  if (reduce_type == 2)
    atomic_max_float(dst, f);
  else
    atomicAdd(dst, f);
}

__global__ void __launch_bounds__(256)
scatter_mean_kernel(float* __restrict__ voxel_feats, const int32_t* __restrict__ count, int64_t m, int channels) {
  int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= m * channels) return;
  voxel_feats[t] = __fdiv_rn(voxel_feats[t], static_cast<float>(count[t / channels]));
}

__global__ void __launch_bounds__(256)
scatter_bwd_add_kernel(const float* __restrict__ grad_voxel, const int32_t* __restrict__ point2voxel,
                       const int32_t* __restrict__ count, int64_t n, int channels, int reduce_type,
                       float* __restrict__ grad_feats) {
  int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n * channels) return;
  const int64_t i = t / channels;
  const int c = static_cast<int>(t - i * channels);
  const int v = point2voxel[i];
  float g = 0.f;
  if (v >= 0) {
    g = grad_voxel[static_cast<int64_t>(v) * channels + c];
    if (reduce_type == 1) g = __fdiv_rn(g, static_cast<float>(count[v]));
  }
  grad_feats[t] = g;
}

__global__ void __launch_bounds__(256)
scatter_bwd_argmax_kernel(const float* __restrict__ feats, const float* __restrict__ voxel_feats,
                          const int32_t* __restrict__ point2voxel, int64_t n, int channels, int32_t* reduce_from) {
  int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n * channels) return;
  const int64_t i = t / channels;
  const int c = static_cast<int>(t - i * channels);
  const int v = point2voxel[i];
  if (v < 0) return;
  const int64_t o = static_cast<int64_t>(v) * channels + c;
  if (feats[t] == voxel_feats[o]) atomicMin(&reduce_from[o], static_cast<int32_t>(i));
}

__global__ void __launch_bounds__(256)
scatter_bwd_max_kernel(const float* __restrict__ grad_voxel, const int32_t* __restrict__ reduce_from, int64_t m,
                       int channels, int64_t n, float* __restrict__ grad_feats) {
  int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= m * channels) return;
  const int c = static_cast<int>(t % channels);
  const int32_t from = reduce_from[t];
  if (from >= 0 && from < n) grad_feats[static_cast<int64_t>(from) * channels + c] = grad_voxel[t];
}

static int64_t words3(const int32_t* dims) {
  if (!dims || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return -1;
  double cells = static_cast<double>(dims[0]) * dims[1] * dims[2];
  if (cells >= 4294967295.0) return -1;
  return (static_cast<int64_t>(cells) + 31) / 32;
}

}  // namespace efgb

using namespace efgb;

extern "C" size_t efgb_scatter_workspace_bytes(int64_t num_points, const int32_t* dims) {
  int64_t nw = words3(dims);
  if (nw < 0) return 0;
  (void)num_points;
  return align_up(nw * sizeof(CellWord)) + align_up(scan_scratch_elems(nw) * sizeof(uint32_t)) + 1024;
}

extern "C" int efgb_scatter_phase1(const int32_t* coors, int64_t num_points, const int32_t* dims,
                                   int32_t* num_voxels_dev, void* workspace, size_t workspace_bytes,
                                   efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  const int64_t nw = words3(dims);
  EFGB_REQUIRE(nw > 0, EFGB_ERANGE, "scatter: coordinate space is empty or exceeds 32-bit cell ids");
  EFGB_REQUIRE(num_points >= 0 && num_voxels_dev && (coors || num_points == 0), EFGB_EINVAL, "scatter_phase1: bad argument");
  Workspace ws(workspace, workspace_bytes);
  CellWord* cells = ws.take<CellWord>(nw);
  uint32_t* scratch = ws.take<uint32_t>(scan_scratch_elems(nw));
  EFGB_REQUIRE(scratch != nullptr, EFGB_EWORKSPACE, "scatter_phase1: workspace too small");
  Dims3 d{dims[0], dims[1], dims[2]};
  EFGB_CUDA_OK(cudaMemsetAsync(cells, 0, nw * sizeof(CellWord), stream));
  if (num_points > 0) {
    scatter_mark_kernel<<<static_cast<unsigned>((num_points + 255) / 256), 256, 0, stream>>>(coors, num_points, d, cells);
    EFGB_LAUNCH_OK("scatter_mark_kernel");
  }
  return cells_scan(cells, nw, reinterpret_cast<uint32_t*>(num_voxels_dev), scratch, stream);
}

extern "C" int efgb_scatter_phase2(const float* feats, const int32_t* coors, int64_t num_points, int channels,
                                   const int32_t* dims, int reduce_type, int64_t num_voxels, float* voxel_feats,
                                   int32_t* voxel_coors, int32_t* point2voxel, int32_t* count, void* workspace,
                                   size_t workspace_bytes, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  const int64_t nw = words3(dims);
  EFGB_REQUIRE(nw > 0, EFGB_ERANGE, "scatter: coordinate space is empty or exceeds 32-bit cell ids");
  EFGB_REQUIRE(num_points >= 0 && channels >= 1 && num_voxels >= 0 && reduce_type >= 0 && reduce_type <= 2, EFGB_EINVAL,
               "scatter_phase2: bad argument");
  Workspace ws(workspace, workspace_bytes);
  CellWord* cells = ws.take<CellWord>(nw);
  EFGB_REQUIRE(cells != nullptr, EFGB_EWORKSPACE, "scatter_phase2: workspace too small");
  Dims3 d{dims[0], dims[1], dims[2]};
  if (num_voxels > 0) {
    EFGB_REQUIRE(voxel_feats && voxel_coors && count, EFGB_EINVAL, "scatter_phase2: null output");
    scatter_coords_kernel<<<static_cast<unsigned>((nw + 255) / 256), 256, 0, stream>>>(cells, nw, d, voxel_coors);
    EFGB_LAUNCH_OK("scatter_coords_kernel");
    const int64_t mc = num_voxels * channels;
    fill_kernel<<<static_cast<unsigned>((mc + 255) / 256), 256, 0, stream>>>(voxel_feats, mc,
                                                                             reduce_type == 2 ? -INFINITY : 0.f);
    EFGB_LAUNCH_OK("fill_kernel");
    EFGB_CUDA_OK(cudaMemsetAsync(count, 0, num_voxels * sizeof(int32_t), stream));
  }
  if (num_points > 0) {
    EFGB_REQUIRE(feats && coors && point2voxel, EFGB_EINVAL, "scatter_phase2: null input");
    const int64_t nc = num_points * channels;
    scatter_reduce_kernel<<<static_cast<unsigned>((nc + 255) / 256), 256, 0, stream>>>(
        feats, coors, num_points, channels, d, cells, reduce_type, voxel_feats, point2voxel, count);
    EFGB_LAUNCH_OK("scatter_reduce_kernel");
  }
  if (reduce_type == 1 && num_voxels > 0) {
    const int64_t mc = num_voxels * channels;
    scatter_mean_kernel<<<static_cast<unsigned>((mc + 255) / 256), 256, 0, stream>>>(voxel_feats, count, num_voxels, channels);
    EFGB_LAUNCH_OK("scatter_mean_kernel");
  }
  return EFGB_OK;
}

extern "C" int efgb_scatter_backward(const float* grad_voxel_feats, const float* feats, const float* voxel_feats,
                                     const int32_t* point2voxel, const int32_t* count, int64_t num_points, int channels,
                                     int reduce_type, int64_t num_voxels, float* grad_feats, void* workspace,
                                     size_t workspace_bytes, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(num_points >= 0 && channels >= 1 && num_voxels >= 0 && reduce_type >= 0 && reduce_type <= 2, EFGB_EINVAL,
               "scatter_backward: bad argument");
  if (num_points == 0) return EFGB_OK;
  EFGB_REQUIRE(grad_feats && point2voxel && (grad_voxel_feats || num_voxels == 0), EFGB_EINVAL, "scatter_backward: null pointer");
  const int64_t nc = num_points * channels;
  if (reduce_type != 2) {
    EFGB_REQUIRE(reduce_type == 0 || count, EFGB_EINVAL, "scatter_backward: mean needs count");
    scatter_bwd_add_kernel<<<static_cast<unsigned>((nc + 255) / 256), 256, 0, stream>>>(
        grad_voxel_feats, point2voxel, count, num_points, channels, reduce_type, grad_feats);
    EFGB_LAUNCH_OK("scatter_bwd_add_kernel");
    return EFGB_OK;
  }
  EFGB_REQUIRE(feats && voxel_feats, EFGB_EINVAL, "scatter_backward: max needs feats and voxel_feats");
  const int64_t mc = num_voxels * channels;
  EFGB_REQUIRE(workspace && workspace_bytes >= static_cast<size_t>(mc > 0 ? mc : 1) * sizeof(int32_t), EFGB_EWORKSPACE,
               "scatter_backward: workspace too small (need num_voxels*channels int32)");
  int32_t* reduce_from = static_cast<int32_t*>(workspace);
  EFGB_CUDA_OK(cudaMemsetAsync(grad_feats, 0, nc * sizeof(float), stream));
  if (mc == 0) return EFGB_OK;
  EFGB_CUDA_OK(cudaMemsetAsync(reduce_from, 0x7F, mc * sizeof(int32_t), stream));  // 0x7F7F7F7F > any index
  scatter_bwd_argmax_kernel<<<static_cast<unsigned>((nc + 255) / 256), 256, 0, stream>>>(
      feats, voxel_feats, point2voxel, num_points, channels, reduce_from);
  EFGB_LAUNCH_OK("scatter_bwd_argmax_kernel");
  scatter_bwd_max_kernel<<<static_cast<unsigned>((mc + 255) / 256), 256, 0, stream>>>(
      grad_voxel_feats, reduce_from, num_voxels, channels, num_points, grad_feats);
  EFGB_LAUNCH_OK("scatter_bwd_max_kernel");
  return EFGB_OK;
}
