// BatchNorm1d over the active-voxel rows of a sparse tensor, fused with what follows it in the reference's blocks:
//     y = ReLU( BN(x) [+ residual] )                                   (sparse_net.py:85-95, 135-147, 162-163, 455-469)
// Batch statistics over all M active rows of the local batch (not synced across ranks).
//
// Forward (training): 3 launches, x is read twice and y written once —
//   stats_partial  per-slab column sums of x and x^2 (fixed summation order: deterministic)
//   stats_final    mean / rstd in double, running statistics updated in place (momentum, unbiased variance)
//   apply          normalise + affine + residual + ReLU in one pass; can also emit the bf16 hi / lo operand planes of
//                  y for the next tensor-core convolution (saves the separate split pass over y)
// Backward: the same shape — per-slab sums of g and g * xhat (g = dy masked by y > 0), dgamma / dbeta, then
//   dx = gamma * rstd * (g - dbeta / M - xhat * dgamma / M), and g itself as the residual's gradient.
// The torch path this replaces runs batch_norm_collect_statistics + batch_norm_transform_input + add + clamp forward and
// threshold_backward + batch_norm_backward_reduce + batch_norm_backward_elemt backward: five full passes over [M, C]
// each way instead of three.
#include <cuda_bf16.h>

#include "common.cuh"

namespace efgb {
namespace bn {

constexpr int kThreads = 256;

// Thread t of a CTA owns the float4 column group t % groups and the row lane t / groups.
template <bool kBackward>
__global__ void __launch_bounds__(kThreads)
stats_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y,
                     const float* __restrict__ mean, const float* __restrict__ rstd, int64_t rows, int cols, int relu,
                     int64_t rows_per_slab, float* __restrict__ partial) {
  __shared__ float4 red[2][kThreads];
  const int groups = cols / 4;
  const int lanes = kThreads / groups;
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
  const int c = g * 4;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.x) * rows_per_slab;
  int64_t r_end = r_begin + rows_per_slab;
  if (r_end > rows) r_end = rows;
  float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
  float4 mu = a1, rs = a1;
  if (kBackward && lane < lanes) {
    mu = *reinterpret_cast<const float4*>(mean + c);
    rs = *reinterpret_cast<const float4*>(rstd + c);
  }
  if (lane < lanes) {
    int64_t r = r_begin + lane;
    if (!kBackward) {   // forward statistics: four independent row loads in flight per thread
      for (; r + 3 * lanes < r_end; r += 4 * lanes) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(x + r * cols + c));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(x + (r + lanes) * cols + c));
        const float4 v2 = __ldg(reinterpret_cast<const float4*>(x + (r + 2 * lanes) * cols + c));
        const float4 v3 = __ldg(reinterpret_cast<const float4*>(x + (r + 3 * lanes) * cols + c));
        a1.x += (v0.x + v1.x) + (v2.x + v3.x); a1.y += (v0.y + v1.y) + (v2.y + v3.y);
        a1.z += (v0.z + v1.z) + (v2.z + v3.z); a1.w += (v0.w + v1.w) + (v2.w + v3.w);
        a2.x += (v0.x * v0.x + v1.x * v1.x) + (v2.x * v2.x + v3.x * v3.x); a2.y += (v0.y * v0.y + v1.y * v1.y) + (v2.y * v2.y + v3.y * v3.y);
        a2.z += (v0.z * v0.z + v1.z * v1.z) + (v2.z * v2.z + v3.z * v3.z); a2.w += (v0.w * v0.w + v1.w * v1.w) + (v2.w * v2.w + v3.w * v3.w);
      }
    }
    for (; r < r_end; r += lanes) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * cols + c));
      if (!kBackward) {
        a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w;
        a2.x += v.x * v.x; a2.y += v.y * v.y; a2.z += v.z * v.z; a2.w += v.w * v.w;
      } else {
        float4 gr = __ldg(reinterpret_cast<const float4*>(dy + r * cols + c));
        if (relu) {
          const float4 o = __ldg(reinterpret_cast<const float4*>(y + r * cols + c));
          gr.x = o.x > 0.f ? gr.x : 0.f; gr.y = o.y > 0.f ? gr.y : 0.f; gr.z = o.z > 0.f ? gr.z : 0.f; gr.w = o.w > 0.f ? gr.w : 0.f;
        }
        a1.x += gr.x; a1.y += gr.y; a1.z += gr.z; a1.w += gr.w;
        a2.x += gr.x * ((v.x - mu.x) * rs.x); a2.y += gr.y * ((v.y - mu.y) * rs.y);
        a2.z += gr.z * ((v.z - mu.z) * rs.z); a2.w += gr.w * ((v.w - mu.w) * rs.w);
      }
    }
  }
  red[0][threadIdx.x] = a1;
  red[1][threadIdx.x] = a2;
  __syncthreads();
  if (lane == 0) {
    for (int k = 1; k < lanes; ++k) {
      const float4 o1 = red[0][k * groups + g], o2 = red[1][k * groups + g];
      a1.x += o1.x; a1.y += o1.y; a1.z += o1.z; a1.w += o1.w;
      a2.x += o2.x; a2.y += o2.y; a2.z += o2.z; a2.w += o2.w;
    }
    float* p = partial + static_cast<int64_t>(blockIdx.x) * 2 * cols;
    *reinterpret_cast<float4*>(p + c) = a1;
    *reinterpret_cast<float4*>(p + cols + c) = a2;
  }
}

// The per-slab partial sums are folded by kFoldCols columns x kFoldLanes slab lanes per CTA with independent loads (a single
// thread walking ~600 slabs is a chain of dependent L2 reads: measured 60 us, the whole normalisation pass took less;
// 32 columns x 8 lanes still took 18 us per BatchNorm — 74 slabs per thread — against ~6 us for the other two kernels).
constexpr int kFoldCols = 8;
constexpr int kFoldLanes = 32;
static_assert(kFoldCols * kFoldLanes == 256, "the final kernels run 256 threads");

__device__ __forceinline__ void fold_partials(const float* __restrict__ partial, int slabs, int cols, int c, int ty, double& s1,
                                              double& s2, double (*red)[2][kFoldCols + 1]) {
  s1 = 0.0;
  s2 = 0.0;
  if (c < cols) {
    int k = ty;
    for (; k + 3 * kFoldLanes < slabs; k += 4 * kFoldLanes) {   // four independent pairs of loads in flight
      const float a0 = partial[static_cast<int64_t>(k) * 2 * cols + c], b0 = partial[static_cast<int64_t>(k) * 2 * cols + cols + c];
      const float a1 = partial[static_cast<int64_t>(k + kFoldLanes) * 2 * cols + c], b1 = partial[static_cast<int64_t>(k + kFoldLanes) * 2 * cols + cols + c];
      const float a2 = partial[static_cast<int64_t>(k + 2 * kFoldLanes) * 2 * cols + c], b2 = partial[static_cast<int64_t>(k + 2 * kFoldLanes) * 2 * cols + cols + c];
      const float a3 = partial[static_cast<int64_t>(k + 3 * kFoldLanes) * 2 * cols + c], b3 = partial[static_cast<int64_t>(k + 3 * kFoldLanes) * 2 * cols + cols + c];
      s1 += (static_cast<double>(a0) + a1) + (static_cast<double>(a2) + a3);
      s2 += (static_cast<double>(b0) + b1) + (static_cast<double>(b2) + b3);
    }
    for (; k < slabs; k += kFoldLanes) {
      s1 += static_cast<double>(partial[static_cast<int64_t>(k) * 2 * cols + c]);
      s2 += static_cast<double>(partial[static_cast<int64_t>(k) * 2 * cols + cols + c]);
    }
  }
  const int tx = threadIdx.x % kFoldCols;
  red[ty][0][tx] = s1;
  red[ty][1][tx] = s2;
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int j = 1; j < kFoldLanes; ++j) {
      s1 += red[j][0][tx];
      s2 += red[j][1][tx];
    }
  }
}

__global__ void __launch_bounds__(256)
stats_final_fwd_kernel(const float* __restrict__ partial, int slabs, int cols, int64_t rows, float eps, float momentum,
                       float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ mean,
                       float* __restrict__ rstd) {
  __shared__ double red[kFoldLanes][2][kFoldCols + 1];
  const int c = blockIdx.x * kFoldCols + (threadIdx.x % kFoldCols), ty = threadIdx.x / kFoldCols;
  double s1, s2;
  fold_partials(partial, slabs, cols, c, ty, s1, s2, red);
  if (ty != 0 || c >= cols) return;
  const double m = s1 / static_cast<double>(rows);
  double var = s2 / static_cast<double>(rows) - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = static_cast<float>(m);
  rstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  if (running_mean) {
    const double unbiased = rows > 1 ? var * static_cast<double>(rows) / static_cast<double>(rows - 1) : var;
    running_mean[c] = static_cast<float>((1.0 - momentum) * running_mean[c] + momentum * m);
    running_var[c] = static_cast<float>((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

__global__ void __launch_bounds__(128)
stats_eval_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var, int cols, float eps,
                  float* __restrict__ mean, float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  mean[c] = running_mean[c];
  rstd[c] = rsqrtf(running_var[c] + eps);
}

__global__ void __launch_bounds__(256)
stats_final_bwd_kernel(const float* __restrict__ partial, int slabs, int cols, float* __restrict__ dbeta, float* __restrict__ dgamma) {
  __shared__ double red[kFoldLanes][2][kFoldCols + 1];
  const int c = blockIdx.x * kFoldCols + (threadIdx.x % kFoldCols), ty = threadIdx.x / kFoldCols;
  double s1, s2;
  fold_partials(partial, slabs, cols, c, ty, s1, s2, red);
  if (ty != 0 || c >= cols) return;
  dbeta[c] = static_cast<float>(s1);
  dgamma[c] = static_cast<float>(s2);
}

__device__ __forceinline__ void store_planes(uint8_t* __restrict__ planes, int64_t row, int cols, int c, const float4& v) {
  // planes[row] = [hi cols x bf16 | lo cols x bf16] (spconv_tc.cu: split_bf16_kernel); this thread owns 4 values = 8 bytes
  const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
  const uint32_t w01 = *reinterpret_cast<const uint32_t*>(&h01), w23 = *reinterpret_cast<const uint32_t*>(&h23);
  const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - __uint_as_float(w01 << 16), v.y - __uint_as_float(w01 & 0xFFFF0000u));
  const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - __uint_as_float(w23 << 16), v.w - __uint_as_float(w23 & 0xFFFF0000u));
  uint8_t* dst = planes + row * cols * 4 + c * 2;
  *reinterpret_cast<uint2*>(dst) = make_uint2(w01, w23);
  *reinterpret_cast<uint2*>(dst + cols * 2) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

__global__ void __launch_bounds__(kThreads)
apply_fwd_kernel(const float* __restrict__ x, const float* __restrict__ residual, const float* __restrict__ mean,
                 const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta, int64_t rows,
                 int cols, int relu, float* __restrict__ y, uint8_t* __restrict__ planes) {
  const int groups = cols / 4;
  const int64_t total = rows * groups;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = e / groups;
    const int c = static_cast<int>(e - r * groups) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * cols + c));
    const float4 mu = *reinterpret_cast<const float4*>(mean + c), rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
    float4 o;
    o.x = (v.x - mu.x) * rs.x * ga.x + be.x;
    o.y = (v.y - mu.y) * rs.y * ga.y + be.y;
    o.z = (v.z - mu.z) * rs.z * ga.z + be.z;
    o.w = (v.w - mu.w) * rs.w * ga.w + be.w;
    if (residual) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(residual + r * cols + c));
      o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
    }
    if (relu) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    *reinterpret_cast<float4*>(y + r * cols + c) = o;
    if (planes) store_planes(planes, r, cols, c, o);
  }
}

__global__ void __launch_bounds__(kThreads)
apply_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                 const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                 const float* __restrict__ dbeta, const float* __restrict__ dgamma, int64_t rows, int cols, int relu,
                 float* __restrict__ dx, float* __restrict__ dres) {
  const int groups = cols / 4;
  const int64_t total = rows * groups;
  const float inv_m = 1.0f / static_cast<float>(rows);
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = e / groups;
    const int c = static_cast<int>(e - r * groups) * 4;
    float4 g = __ldg(reinterpret_cast<const float4*>(dy + r * cols + c));
    if (relu) {
      const float4 o = __ldg(reinterpret_cast<const float4*>(y + r * cols + c));
      g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    }
    if (dres) *reinterpret_cast<float4*>(dres + r * cols + c) = g;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * cols + c));
    const float4 mu = *reinterpret_cast<const float4*>(mean + c), rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
    const float4 db = *reinterpret_cast<const float4*>(dbeta + c), dg = *reinterpret_cast<const float4*>(dgamma + c);
    float4 o;
    o.x = ga.x * rs.x * (g.x - db.x * inv_m - (v.x - mu.x) * rs.x * dg.x * inv_m);
    o.y = ga.y * rs.y * (g.y - db.y * inv_m - (v.y - mu.y) * rs.y * dg.y * inv_m);
    o.z = ga.z * rs.z * (g.z - db.z * inv_m - (v.z - mu.z) * rs.z * dg.z * inv_m);
    o.w = ga.w * rs.w * (g.w - db.w * inv_m - (v.w - mu.w) * rs.w * dg.w * inv_m);
    *reinterpret_cast<float4*>(dx + r * cols + c) = o;
  }
}

static bool supported(int cols) { return cols >= 4 && cols % 4 == 0 && cols / 4 <= kThreads && kThreads % (cols / 4) == 0; }

static int num_slabs(int64_t rows, int cols) {
  const int lanes = kThreads / (cols / 4) > 0 ? kThreads / (cols / 4) : 1;
  int64_t slabs = static_cast<int64_t>(kNumSMs) * 2;
  const int64_t max_slabs = (rows + 4 * lanes - 1) / (4 * lanes);
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  return static_cast<int>(slabs);
}

}  // namespace bn
}  // namespace efgb

using namespace efgb;

extern "C" int efgb_bn_supported(int cols) { return bn::supported(cols) ? 1 : 0; }

extern "C" size_t efgb_bn_workspace_bytes(int64_t rows, int cols) {
  if (rows <= 0 || !bn::supported(cols)) return 256;
  return align_up(static_cast<size_t>(bn::num_slabs(rows, cols)) * 2 * cols * sizeof(float));
}

extern "C" int efgb_bn_forward(const float* x, int64_t rows, int cols, const float* gamma, const float* beta,
                               const float* residual, int relu, float eps, float momentum, float* running_mean,
                               float* running_var, int training, float* y, float* save_mean, float* save_rstd, void* planes,
                               void* workspace, size_t workspace_bytes, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(efgb_bn_supported(cols), EFGB_EINVAL, "bn_forward: unsupported channel count %d", cols);
  EFGB_REQUIRE(rows >= 0 && gamma && beta && save_mean && save_rstd, EFGB_EINVAL, "bn_forward: bad argument");
  EFGB_REQUIRE(training || (running_mean && running_var), EFGB_EINVAL, "bn_forward: eval mode needs running statistics");
  if (rows == 0) return EFGB_OK;
  EFGB_REQUIRE(x && y && workspace, EFGB_EINVAL, "bn_forward: null pointer");
  EFGB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(residual) |
                 reinterpret_cast<uintptr_t>(planes)) & 15) == 0, EFGB_EINVAL, "bn_forward: pointers must be 16-byte aligned");
  if (training) {
    const int slabs = bn::num_slabs(rows, cols);
    EFGB_REQUIRE(workspace_bytes >= static_cast<size_t>(slabs) * 2 * cols * sizeof(float), EFGB_EINVAL, "bn_forward: workspace too small");
    const int64_t rows_per_slab = (rows + slabs - 1) / slabs;
    float* partial = static_cast<float*>(workspace);
    bn::stats_partial_kernel<false><<<slabs, bn::kThreads, 0, stream>>>(x, nullptr, nullptr, nullptr, nullptr, rows, cols, 0,
                                                                     rows_per_slab, partial);
    EFGB_LAUNCH_OK("bn::stats_partial_kernel");
    bn::stats_final_fwd_kernel<<<(cols + bn::kFoldCols - 1) / bn::kFoldCols, 256, 0, stream>>>(partial, slabs, cols, rows, eps, momentum, running_mean,
                                                                      running_var, save_mean, save_rstd);
    EFGB_LAUNCH_OK("bn::stats_final_fwd_kernel");
  } else {
    bn::stats_eval_kernel<<<(cols + 127) / 128, 128, 0, stream>>>(running_mean, running_var, cols, eps, save_mean, save_rstd);
    EFGB_LAUNCH_OK("bn::stats_eval_kernel");
  }
  bn::apply_fwd_kernel<<<grid_for(rows * (cols / 4), bn::kThreads, kNumSMs * 8), bn::kThreads, 0, stream>>>(
      x, residual, save_mean, save_rstd, gamma, beta, rows, cols, relu, y, static_cast<uint8_t*>(planes));
  EFGB_LAUNCH_OK("bn::apply_fwd_kernel");
  return EFGB_OK;
}

extern "C" int efgb_bn_backward(const float* dy, const float* x, const float* y, const float* gamma, const float* save_mean,
                                const float* save_rstd, int64_t rows, int cols, int relu, float* dx, float* dres, float* dgamma,
                                float* dbeta, void* workspace, size_t workspace_bytes, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(efgb_bn_supported(cols), EFGB_EINVAL, "bn_backward: unsupported channel count %d", cols);
  EFGB_REQUIRE(rows >= 0 && gamma && save_mean && save_rstd && dgamma && dbeta, EFGB_EINVAL, "bn_backward: bad argument");
  if (rows == 0) {
    EFGB_CUDA_OK(cudaMemsetAsync(dgamma, 0, static_cast<size_t>(cols) * sizeof(float), stream));
    EFGB_CUDA_OK(cudaMemsetAsync(dbeta, 0, static_cast<size_t>(cols) * sizeof(float), stream));
    return EFGB_OK;
  }
  EFGB_REQUIRE(dy && x && dx && workspace && (!relu || y), EFGB_EINVAL, "bn_backward: null pointer");
  const int slabs = bn::num_slabs(rows, cols);
  EFGB_REQUIRE(workspace_bytes >= static_cast<size_t>(slabs) * 2 * cols * sizeof(float), EFGB_EINVAL, "bn_backward: workspace too small");
  const int64_t rows_per_slab = (rows + slabs - 1) / slabs;
  float* partial = static_cast<float*>(workspace);
  bn::stats_partial_kernel<true><<<slabs, bn::kThreads, 0, stream>>>(x, dy, y, save_mean, save_rstd, rows, cols, relu, rows_per_slab,
                                                                  partial);
  EFGB_LAUNCH_OK("bn::stats_partial_kernel");
  bn::stats_final_bwd_kernel<<<(cols + bn::kFoldCols - 1) / bn::kFoldCols, 256, 0, stream>>>(partial, slabs, cols, dbeta, dgamma);
  EFGB_LAUNCH_OK("bn::stats_final_bwd_kernel");
  bn::apply_bwd_kernel<<<grid_for(rows * (cols / 4), bn::kThreads, kNumSMs * 8), bn::kThreads, 0, stream>>>(
      dy, x, y, save_mean, save_rstd, gamma, dbeta, dgamma, rows, cols, relu, dx, dres);
  EFGB_LAUNCH_OK("bn::apply_bwd_kernel");
  return EFGB_OK;
}
