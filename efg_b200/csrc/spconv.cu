// Sparse convolution over a device rulebook: fp32 FFMA output-stationary gather-GEMM (forward,
// dgrad), split-row wgrad, and the SparseConvTensor.dense() scatter / gather.
//
// Math (what spconv computes for the reference, efg/modeling/backbones/sparse_net.py:85-95,
// 125-147, 273-282): out[o,:] = bias + sum_k in[nbr[o,k],:] @ W[k], cross-correlation
// orientation as torch.nn.Conv3d restricted to active sites.  This file is the exact-fp32 path
// (bit-stable accumulation order per output: taps ascending, channels ascending); the tcgen05
// path for C >= 32 lives in spconv_tc.cu.
#include "common.cuh"

namespace efgb {

constexpr int kTK = 16;       // input-channel chunk staged per step
constexpr int kMaxTaps = 128;  // 5x5x5 fits

// 256 threads as 16x16; thread (ty,tx) owns rows ty+16*i, cols tx+16*j.
template <int TM, int TN>
__global__ void __launch_bounds__(256)
spconv_fwd_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                  const int32_t* __restrict__ nbr, float* __restrict__ out, int64_t m_out, int c_in, int c_out,
                  int taps) {
  constexpr int RI = TM / 16, CJ = TN / 16;
  extern __shared__ int32_t s_nbr[];  // [TM * taps]
  __shared__ float As[kTK][TM + 2];
  __shared__ float Bs[kTK][TN];
  __shared__ int s_tap_any[kMaxTaps];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * TM;
  const int c0 = blockIdx.y * TN;

  for (int t = tid; t < taps; t += 256) s_tap_any[t] = 0;
  __syncthreads();
  {
    const int64_t total = static_cast<int64_t>(TM) * taps;
    const int64_t limit = (m_out - r0) * taps;  // entries that exist
    const int32_t* src = nbr + r0 * taps;
    for (int64_t e = tid; e < total; e += 256) {
      int32_t v = e < limit ? src[e] : -1;
      s_nbr[e] = v;
      if (v >= 0) s_tap_any[e % taps] = 1;
    }
  }
  __syncthreads();

  float acc[RI][CJ];
#pragma unroll
  for (int i = 0; i < RI; ++i)
#pragma unroll
    for (int j = 0; j < CJ; ++j) acc[i][j] = 0.f;

  for (int k = 0; k < taps; ++k) {
    if (!s_tap_any[k]) continue;  // uniform across the block
    const float* wk = w + static_cast<int64_t>(k) * c_in * c_out;
    for (int ci0 = 0; ci0 < c_in; ci0 += kTK) {
      // stage A: TM rows x kTK channels (zeros for missing neighbours / channel tail)
      for (int e = tid; e < TM * kTK; e += 256) {
        const int r = e / kTK, kk = e % kTK;
        const int32_t src = s_nbr[r * taps + k];
        const int ci = ci0 + kk;
        float v = 0.f;
        if (src >= 0 && ci < c_in) v = in[static_cast<int64_t>(src) * c_in + ci];
        As[kk][r] = v;
      }
      // stage B: kTK x TN slice of W[k]
      for (int e = tid; e < kTK * TN; e += 256) {
        const int kk = e / TN, n = e % TN;
        const int ci = ci0 + kk, co = c0 + n;
        Bs[kk][n] = (ci < c_in && co < c_out) ? wk[static_cast<int64_t>(ci) * c_out + co] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < kTK; ++kk) {
        float a[RI], b[CJ];
#pragma unroll
        for (int i = 0; i < RI; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
        for (int j = 0; j < CJ; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < RI; ++i)
#pragma unroll
          for (int j = 0; j < CJ; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < RI; ++i) {
    const int64_t r = r0 + ty + 16 * i;
    if (r >= m_out) continue;
#pragma unroll
    for (int j = 0; j < CJ; ++j) {
      const int co = c0 + tx + 16 * j;
      if (co < c_out) out[r * c_out + co] = acc[i][j] + (bias ? bias[co] : 0.f);
    }
  }
}

// wgrad: grid (row_chunks, taps, ci_tiles*co_tiles); 256 threads as 16x16 over a TC x TC tile of dW[k].
template <int TC>
__global__ void __launch_bounds__(256)
spconv_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ gout, const int32_t* __restrict__ nbr,
                    float* __restrict__ dw, int64_t m_out, int c_in, int c_out, int taps, int64_t rows_per_chunk,
                    int co_tiles) {
  constexpr int R = TC / 16;
  constexpr int TR = 16;
  __shared__ float As[TR][TC + 1];
  __shared__ float Gs[TR][TC + 1];
  __shared__ int32_t s_src[TR];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int k = blockIdx.y;
  const int ci0 = (blockIdx.z / co_tiles) * TC;
  const int co0 = (blockIdx.z % co_tiles) * TC;
  const int64_t row_begin = static_cast<int64_t>(blockIdx.x) * rows_per_chunk;
  int64_t row_end = row_begin + rows_per_chunk;
  if (row_end > m_out) row_end = m_out;

  float acc[R][R];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[i][j] = 0.f;

  for (int64_t rb = row_begin; rb < row_end; rb += TR) {
    int valid = 0;
    if (tid < TR) {
      const int64_t r = rb + tid;
      int32_t s = r < row_end ? nbr[r * taps + k] : -1;
      s_src[tid] = s;
      valid = s >= 0;
    }
    if (!__syncthreads_or(valid)) continue;
    for (int e = tid; e < TR * TC; e += 256) {
      const int r = e / TC, c = e % TC;
      const int32_t s = s_src[r];
      const int ci = ci0 + c, co = co0 + c;
      As[r][c] = (s >= 0 && ci < c_in) ? in[static_cast<int64_t>(s) * c_in + ci] : 0.f;
      Gs[r][c] = (s >= 0 && co < c_out) ? gout[(rb + r) * c_out + co] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      float a[R], g[R];
#pragma unroll
      for (int i = 0; i < R; ++i) a[i] = As[r][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < R; ++j) g[j] = Gs[r][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] = fmaf(a[i], g[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* dwk = dw + static_cast<int64_t>(k) * c_in * c_out;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int ci = ci0 + ty + 16 * i;
    if (ci >= c_in) continue;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int co = co0 + tx + 16 * j;
      if (co < c_out && acc[i][j] != 0.f) atomicAdd(&dwk[static_cast<int64_t>(ci) * c_out + co], acc[i][j]);
    }
  }
}

// dense[b, c, z, y, x] = feats[row, c].  Block = 32 rows; a warp writes one channel for the 32
// rows at a time, so x-adjacent (sorted) rows give coalesced stores.
__global__ void __launch_bounds__(256)
sparse_to_dense_kernel(const float* __restrict__ feats, const int4* __restrict__ coords, int64_t m, int channels,
                       int batch, int D, int H, int W, float* __restrict__ dense, bool gather) {
  extern __shared__ float s_tile[];  // [32][channels + 1]
  __shared__ int64_t s_base[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int64_t plane = static_cast<int64_t>(D) * H * W;
  const int cs = channels + 1;
  if (tid < 32) {
    int64_t r = r0 + tid;
    int64_t base = -1;
    if (r < m) {
      int4 c = coords[r];
      if (c.x >= 0 && c.x < batch && c.y >= 0 && c.y < D && c.z >= 0 && c.z < H && c.w >= 0 && c.w < W)
        base = static_cast<int64_t>(c.x) * channels * plane + (static_cast<int64_t>(c.y) * H + c.z) * W + c.w;
    }
    s_base[tid] = base;
  }
  if (!gather) {
    for (int e = tid; e < 32 * channels; e += 256) {
      int r = e / channels, c = e % channels;
      s_tile[r * cs + c] = (r0 + r < m) ? feats[(r0 + r) * channels + c] : 0.f;
    }
  }
  __syncthreads();
  const int64_t base = s_base[lane];
  for (int c = warp; c < channels; c += 8) {
    if (base >= 0) {
      if (gather)
        s_tile[lane * cs + c] = dense[base + c * plane];
      else
        dense[base + c * plane] = s_tile[lane * cs + c];
    }
  }
  if (gather) {
    __syncthreads();
    float* out = const_cast<float*>(feats);
    for (int e = tid; e < 32 * channels; e += 256) {
      int r = e / channels, c = e % channels;
      if (r0 + r < m) out[(r0 + r) * channels + c] = s_base[r] >= 0 ? s_tile[r * cs + c] : 0.f;
    }
  }
}

template <int TM, int TN>
static int launch_fwd(const float* in, const float* w, const float* bias, const int32_t* nbr, float* out,
                      int64_t m_out, int c_in, int c_out, int taps, cudaStream_t stream) {
  dim3 grid(static_cast<unsigned>((m_out + TM - 1) / TM), static_cast<unsigned>((c_out + TN - 1) / TN));
  size_t smem = static_cast<size_t>(TM) * taps * sizeof(int32_t);
  static bool configured = false;
  if (!configured) {
    EFGB_CUDA_OK(cudaFuncSetAttribute(spconv_fwd_kernel<TM, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  EFGB_REQUIRE(smem <= 160 * 1024, EFGB_EINVAL, "spconv_forward: %d taps is too many", taps);
  spconv_fwd_kernel<TM, TN><<<grid, 256, smem, stream>>>(in, w, bias, nbr, out, m_out, c_in, c_out, taps);
  EFGB_LAUNCH_OK("spconv_fwd_kernel");
  return EFGB_OK;
}

}  // namespace efgb

using namespace efgb;

extern "C" int efgb_spconv_forward(const float* in_feats, int64_t num_in, int c_in, const float* w_kio,
                                   const float* bias, const int32_t* nbr, int64_t num_out, int num_taps, int c_out,
                                   float* out_feats, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(num_in >= 0 && num_out >= 0 && c_in >= 1 && c_out >= 1 && num_taps >= 1 && num_taps <= kMaxTaps,
               EFGB_EINVAL, "spconv_forward: bad shape (taps=%d, c_in=%d, c_out=%d)", num_taps, c_in, c_out);
  if (num_out == 0) return EFGB_OK;
  EFGB_REQUIRE(w_kio && nbr && out_feats && (in_feats || num_in == 0), EFGB_EINVAL, "spconv_forward: null pointer");
  if (c_out <= 16) return launch_fwd<256, 16>(in_feats, w_kio, bias, nbr, out_feats, num_out, c_in, c_out, num_taps, stream);
  if (c_out <= 32) return launch_fwd<128, 32>(in_feats, w_kio, bias, nbr, out_feats, num_out, c_in, c_out, num_taps, stream);
  return launch_fwd<64, 64>(in_feats, w_kio, bias, nbr, out_feats, num_out, c_in, c_out, num_taps, stream);
}

extern "C" int efgb_spconv_wgrad(const float* in_feats, int64_t num_in, int c_in, const float* grad_out,
                                 const int32_t* nbr, int64_t num_out, int num_taps, int c_out, float* dw_kio,
                                 efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(num_in >= 0 && num_out >= 0 && c_in >= 1 && c_out >= 1 && num_taps >= 1 && num_taps <= kMaxTaps,
               EFGB_EINVAL, "spconv_wgrad: bad shape");
  EFGB_REQUIRE(dw_kio, EFGB_EINVAL, "spconv_wgrad: null dw");
  EFGB_CUDA_OK(cudaMemsetAsync(dw_kio, 0, static_cast<size_t>(num_taps) * c_in * c_out * sizeof(float), stream));
  if (num_out == 0 || num_in == 0) return EFGB_OK;
  EFGB_REQUIRE(in_feats && grad_out && nbr, EFGB_EINVAL, "spconv_wgrad: null pointer");
  const int cmax = c_in > c_out ? c_in : c_out;
  const int TC = cmax <= 16 ? 16 : (cmax <= 32 ? 32 : 64);
  const int ci_tiles = (c_in + TC - 1) / TC, co_tiles = (c_out + TC - 1) / TC;
  const int64_t ctas_per_chunk = static_cast<int64_t>(num_taps) * ci_tiles * co_tiles;
  int64_t chunks = (kNumSMs * 4 + ctas_per_chunk - 1) / ctas_per_chunk;
  const int64_t max_chunks = (num_out + 255) / 256;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int64_t rows_per_chunk = (num_out + chunks - 1) / chunks;
  rows_per_chunk = (rows_per_chunk + 15) / 16 * 16;
  chunks = (num_out + rows_per_chunk - 1) / rows_per_chunk;
  dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(num_taps), static_cast<unsigned>(ci_tiles * co_tiles));
  if (TC == 16)
    spconv_wgrad_kernel<16><<<grid, 256, 0, stream>>>(in_feats, grad_out, nbr, dw_kio, num_out, c_in, c_out, num_taps,
                                                      rows_per_chunk, co_tiles);
  else if (TC == 32)
    spconv_wgrad_kernel<32><<<grid, 256, 0, stream>>>(in_feats, grad_out, nbr, dw_kio, num_out, c_in, c_out, num_taps,
                                                      rows_per_chunk, co_tiles);
  else
    spconv_wgrad_kernel<64><<<grid, 256, 0, stream>>>(in_feats, grad_out, nbr, dw_kio, num_out, c_in, c_out, num_taps,
                                                      rows_per_chunk, co_tiles);
  EFGB_LAUNCH_OK("spconv_wgrad_kernel");
  return EFGB_OK;
}

static int dense_common(const float* feats, const int32_t* coords, int64_t num_rows, int channels, int batch,
                        const int32_t* dhw, float* dense, bool gather, cudaStream_t stream) {
  EFGB_REQUIRE(num_rows >= 0 && channels >= 1 && batch >= 1 && dhw && dense, EFGB_EINVAL, "dense: bad argument");
  EFGB_REQUIRE(channels <= 2048, EFGB_EINVAL, "dense: too many channels");
  const size_t total = static_cast<size_t>(batch) * channels * dhw[0] * dhw[1] * dhw[2];
  if (!gather) EFGB_CUDA_OK(cudaMemsetAsync(dense, 0, total * sizeof(float), stream));
  if (num_rows == 0) return EFGB_OK;
  EFGB_REQUIRE(feats && coords, EFGB_EINVAL, "dense: null pointer");
  const size_t smem = 32 * static_cast<size_t>(channels + 1) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    EFGB_CUDA_OK(cudaFuncSetAttribute(sparse_to_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  const unsigned nb = static_cast<unsigned>((num_rows + 31) / 32);
  sparse_to_dense_kernel<<<nb, 256, smem, stream>>>(feats, reinterpret_cast<const int4*>(coords), num_rows, channels,
                                                    batch, dhw[0], dhw[1], dhw[2], dense, gather);
  EFGB_LAUNCH_OK("sparse_to_dense_kernel");
  return EFGB_OK;
}

extern "C" int efgb_sparse_to_dense(const float* feats, const int32_t* coords, int64_t num_rows, int channels,
                                    int batch, const int32_t* grid_dhw, float* dense, efgb_stream_t stream) {
  return dense_common(feats, coords, num_rows, channels, batch, grid_dhw, dense, false, as_stream(stream));
}

extern "C" int efgb_dense_to_sparse(const float* dense, const int32_t* coords, int64_t num_rows, int channels,
                                    int batch, const int32_t* grid_dhw, float* feats, efgb_stream_t stream) {
  return dense_common(feats, coords, num_rows, channels, batch, grid_dhw, const_cast<float*>(dense), true,
                      as_stream(stream));
}
