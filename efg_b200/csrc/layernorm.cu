// Fused residual add + LayerNorm over the channel dimension, forward and backward.
//
// Reference: every encoder layer computes  src = norm1(src + dropout(attended))  and  norm2(src + dropout(ffn))
// (VD/transformer.py:60-64 — nn.LayerNorm on [B, 35 344, 256] tokens): an elementwise add kernel followed by
// torch's layer-norm kernels (measured 29 + 121 us forward, 63 + 26 us backward per call at B = 2).
// Here one pass each way, HBM-bound:
//   forward   z = x + r;  y = (z - mean(z)) * rstd(z) * gamma + beta          reads 2, writes 2 (y, z) + 2 floats / row
//   backward  dz = (dy*gamma - mean(dy*gamma) - xhat * mean(dy*gamma*xhat)) * rstd   (= dx = dr)
//             dgamma = sum_rows dy * xhat,  dbeta = sum_rows dy                reads 2 (dy, z), writes 1
// One warp per row, each lane owning kVec float4 column groups (cols = 128 * kVec <= 512); statistics by warp
// shuffles, two-pass variance in registers.  dgamma / dbeta: per-CTA partial rows in a fixed order, summed by a
// second small kernel (deterministic, no atomics).
#include "common.cuh"

namespace efgb {

constexpr int kLnWarps = 8;

__device__ __forceinline__ float ln_warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

template <int kVec>
__global__ void __launch_bounds__(kLnWarps * 32)
add_layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ r, const float* __restrict__ gamma,
                         const float* __restrict__ beta, int64_t rows, float eps, float* __restrict__ y,
                         float* __restrict__ z, float* __restrict__ mean, float* __restrict__ rstd) {
  constexpr int kCols = kVec * 128;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kLnWarps + (threadIdx.x >> 5);
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * kLnWarps;
  float4 g[kVec], b[kVec];
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    g[k] = __ldg(reinterpret_cast<const float4*>(gamma) + k * 32 + lane);
    b[k] = __ldg(reinterpret_cast<const float4*>(beta) + k * 32 + lane);
  }
  for (int64_t row = warp0; row < rows; row += nwarps) {
    const float4* xp = reinterpret_cast<const float4*>(x + row * kCols);
    const float4* rp = r ? reinterpret_cast<const float4*>(r + row * kCols) : nullptr;
    float4 v[kVec];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      v[k] = __ldg(xp + k * 32 + lane);
      if (rp) {
        const float4 t = __ldg(rp + k * 32 + lane);
        v[k].x += t.x;
        v[k].y += t.y;
        v[k].z += t.z;
        v[k].w += t.w;
      }
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    const float mu = ln_warp_sum(s) * (1.f / kCols);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      const float a0 = v[k].x - mu, a1 = v[k].y - mu, a2 = v[k].z - mu, a3 = v[k].w - mu;
      q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
    const float rs = rsqrtf(ln_warp_sum(q) * (1.f / kCols) + eps);
    float4* yp = reinterpret_cast<float4*>(y + row * kCols);
    float4* zp = z ? reinterpret_cast<float4*>(z + row * kCols) : nullptr;
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      float4 o;
      o.x = (v[k].x - mu) * rs * g[k].x + b[k].x;
      o.y = (v[k].y - mu) * rs * g[k].y + b[k].y;
      o.z = (v[k].z - mu) * rs * g[k].z + b[k].z;
      o.w = (v[k].w - mu) * rs * g[k].w + b[k].w;
      yp[k * 32 + lane] = o;
      if (zp) zp[k * 32 + lane] = v[k];
    }
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
  }
}

template <int kVec>
__global__ void __launch_bounds__(kLnWarps * 32)
add_layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z, const float* __restrict__ mean,
                         const float* __restrict__ rstd, const float* __restrict__ gamma, int64_t rows,
                         float* __restrict__ dz, float* __restrict__ partial /* [gridDim.x, 2, cols] */) {
  constexpr int kCols = kVec * 128;
  __shared__ float4 red[kLnWarps][2][kVec][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kLnWarps + warp;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * kLnWarps;
  float4 g[kVec], dg[kVec], db[kVec];
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    g[k] = __ldg(reinterpret_cast<const float4*>(gamma) + k * 32 + lane);
    dg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t row = warp0; row < rows; row += nwarps) {
    const float4* dp = reinterpret_cast<const float4*>(dy + row * kCols);
    const float4* zp = reinterpret_cast<const float4*>(z + row * kCols);
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    float4 d[kVec], xh[kVec];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      d[k] = __ldg(dp + k * 32 + lane);
      const float4 zz = __ldg(zp + k * 32 + lane);
      xh[k] = make_float4((zz.x - mu) * rs, (zz.y - mu) * rs, (zz.z - mu) * rs, (zz.w - mu) * rs);
      db[k].x += d[k].x;
      db[k].y += d[k].y;
      db[k].z += d[k].z;
      db[k].w += d[k].w;
      dg[k].x += d[k].x * xh[k].x;
      dg[k].y += d[k].y * xh[k].y;
      dg[k].z += d[k].z * xh[k].z;
      dg[k].w += d[k].w * xh[k].w;
      d[k].x *= g[k].x;  // dy * gamma
      d[k].y *= g[k].y;
      d[k].z *= g[k].z;
      d[k].w *= g[k].w;
      s1 += (d[k].x + d[k].y) + (d[k].z + d[k].w);
      s2 += (d[k].x * xh[k].x + d[k].y * xh[k].y) + (d[k].z * xh[k].z + d[k].w * xh[k].w);
    }
    const float c1 = ln_warp_sum(s1) * (1.f / kCols), c2 = ln_warp_sum(s2) * (1.f / kCols);
    float4* op = reinterpret_cast<float4*>(dz + row * kCols);
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      float4 o;
      o.x = (d[k].x - c1 - xh[k].x * c2) * rs;
      o.y = (d[k].y - c1 - xh[k].y * c2) * rs;
      o.z = (d[k].z - c1 - xh[k].z * c2) * rs;
      o.w = (d[k].w - c1 - xh[k].w * c2) * rs;
      op[k * 32 + lane] = o;
    }
  }
  // fold the CTA's 8 warps (fixed order) and write one partial row pair
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    red[warp][0][k][lane] = dg[k];
    red[warp][1][k][lane] = db[k];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * kVec * 32; e += kLnWarps * 32) {
    const int which = e / (kVec * 32), rem = e - which * kVec * 32;
    const int k = rem / 32, l = rem - k * 32;
    float4 a = red[0][which][k][l];
#pragma unroll
    for (int w = 1; w < kLnWarps; ++w) {
      const float4 o = red[w][which][k][l];
      a.x += o.x;
      a.y += o.y;
      a.z += o.z;
      a.w += o.w;
    }
    reinterpret_cast<float4*>(partial + (static_cast<int64_t>(blockIdx.x) * 2 + which) * kCols)[k * 32 + l] = a;
  }
}

// out[w, c] = sum over CTAs of partial[cta, w, c]   (w = 0: dgamma, 1: dbeta)
__global__ void __launch_bounds__(256)
layernorm_param_grad_kernel(const float* __restrict__ partial, int ctas, int cols, float* __restrict__ dgamma,
                            float* __restrict__ dbeta) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c2 = blockIdx.x * 32 + tx;  // index into the 2 * cols concatenation
  float s = 0.f;
  if (c2 < 2 * cols) {
    const int which = c2 / cols, c = c2 - which * cols;
    for (int k = ty; k < ctas; k += 8) s += partial[(static_cast<int64_t>(k) * 2 + which) * cols + c];
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c2 < 2 * cols) {
#pragma unroll
    for (int k = 1; k < 8; ++k) s += red[k][tx];
    if (c2 < cols)
      dgamma[c2] = s;
    else
      dbeta[c2 - cols] = s;
  }
}

static int ln_ctas(int64_t rows) {
  int64_t c = (rows + kLnWarps - 1) / kLnWarps;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;  // 8 CTAs of 256 threads per SM
  if (c > cap) c = cap;
  if (c < 1) c = 1;
  return static_cast<int>(c);
}

}  // namespace efgb

using namespace efgb;

extern "C" int efgb_add_layernorm_supported(int cols) { return (cols % 128 == 0 && cols >= 128 && cols <= 512) ? 1 : 0; }

extern "C" size_t efgb_add_layernorm_workspace_bytes(int64_t rows, int cols) {
  return align_up(static_cast<size_t>(ln_ctas(rows)) * 2 * cols * sizeof(float));
}

extern "C" int efgb_add_layernorm_forward(const float* x, const float* residual, const float* gamma, const float* beta,
                                          int64_t rows, int cols, float eps, float* y, float* z, float* mean, float* rstd,
                                          efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(efgb_add_layernorm_supported(cols), EFGB_EINVAL, "add_layernorm: cols must be a multiple of 128 in [128, 512] (got %d)", cols);
  EFGB_REQUIRE(rows >= 0, EFGB_EINVAL, "add_layernorm: negative rows");
  if (rows == 0) return EFGB_OK;
  EFGB_REQUIRE(x && gamma && beta && y && mean && rstd, EFGB_EINVAL, "add_layernorm_forward: null pointer");
  const int ctas = ln_ctas(rows);
#define EFGB_LN_FWD(V) \
  add_layernorm_fwd_kernel<V><<<ctas, kLnWarps * 32, 0, stream>>>(x, residual, gamma, beta, rows, eps, y, z, mean, rstd)
  switch (cols / 128) {
    case 1: EFGB_LN_FWD(1); break;
    case 2: EFGB_LN_FWD(2); break;
    case 3: EFGB_LN_FWD(3); break;
    default: EFGB_LN_FWD(4); break;
  }
#undef EFGB_LN_FWD
  EFGB_LAUNCH_OK("add_layernorm_fwd_kernel");
  return EFGB_OK;
}

extern "C" int efgb_add_layernorm_backward(const float* dy, const float* z, const float* mean, const float* rstd,
                                           const float* gamma, int64_t rows, int cols, float* dz, float* dgamma,
                                           float* dbeta, void* workspace, size_t workspace_bytes, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(efgb_add_layernorm_supported(cols), EFGB_EINVAL, "add_layernorm: cols must be a multiple of 128 in [128, 512] (got %d)", cols);
  EFGB_REQUIRE(rows >= 0 && dgamma && dbeta, EFGB_EINVAL, "add_layernorm_backward: bad argument");
  if (rows == 0) {
    EFGB_CUDA_OK(cudaMemsetAsync(dgamma, 0, static_cast<size_t>(cols) * sizeof(float), stream));
    EFGB_CUDA_OK(cudaMemsetAsync(dbeta, 0, static_cast<size_t>(cols) * sizeof(float), stream));
    return EFGB_OK;
  }
  EFGB_REQUIRE(dy && z && mean && rstd && gamma && dz && workspace, EFGB_EINVAL, "add_layernorm_backward: null pointer");
  const int ctas = ln_ctas(rows);
  EFGB_REQUIRE(workspace_bytes >= static_cast<size_t>(ctas) * 2 * cols * sizeof(float), EFGB_EINVAL,
               "add_layernorm_backward: workspace too small");
  float* partial = static_cast<float*>(workspace);
#define EFGB_LN_BWD(V) \
  add_layernorm_bwd_kernel<V><<<ctas, kLnWarps * 32, 0, stream>>>(dy, z, mean, rstd, gamma, rows, dz, partial)
  switch (cols / 128) {
    case 1: EFGB_LN_BWD(1); break;
    case 2: EFGB_LN_BWD(2); break;
    case 3: EFGB_LN_BWD(3); break;
    default: EFGB_LN_BWD(4); break;
  }
#undef EFGB_LN_BWD
  EFGB_LAUNCH_OK("add_layernorm_bwd_kernel");
  layernorm_param_grad_kernel<<<(2 * cols + 31) / 32, 256, 0, stream>>>(partial, ctas, cols, dgamma, dbeta);
  EFGB_LAUNCH_OK("layernorm_param_grad_kernel");
  return EFGB_OK;
}
