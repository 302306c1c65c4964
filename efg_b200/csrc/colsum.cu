// Column sums of a row-major [rows, cols] fp32 matrix: out[c] = sum_r x[r, c].
//
// This is the bias gradient of every token-wise linear layer of the box-attention encoder
// (db = grad_out.sum(0) over 70 688 rows; the reference gets it from torch's autograd of F.linear,
// VD/transformer.py / VD/modules/box_attention.py).  HBM-bound: rows * cols * 4 bytes read once.
//
// Two deterministic passes (no atomics, the summation order is fixed by the launch geometry):
//   1. grid (slabs, col_tiles): a CTA of 4 x 64 threads walks its row slab, every thread accumulating one
//      float4 column group (a warp reads 512 contiguous bytes of a row); the 4 row lanes are folded through
//      shared memory and the CTA writes one partial row  partial[slab, cols];
//   2. 32 columns x 8 slab lanes per CTA sum the `slabs` partial rows.
#include "common.cuh"

namespace efgb {

constexpr int kColsumTx = 64;   // float4 column groups per CTA (256 columns)
constexpr int kColsumTy = 4;    // row lanes per CTA

__global__ void __launch_bounds__(kColsumTx * kColsumTy)
colsum_partial_kernel(const float* __restrict__ x, int64_t rows, int cols, int64_t rows_per_slab, float* __restrict__ partial) {
  __shared__ float4 red[kColsumTy][kColsumTx];
  const int tx = threadIdx.x % kColsumTx, ty = threadIdx.x / kColsumTx;
  const int c = (blockIdx.y * kColsumTx + tx) * 4;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.x) * rows_per_slab;
  int64_t r_end = r_begin + rows_per_slab;
  if (r_end > rows) r_end = rows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < cols) {
    const float* p = x + c;
    int64_t r = r_begin + ty;
    // 4 independent 16-byte loads in flight per thread
    for (; r + 3 * kColsumTy < r_end; r += 4 * kColsumTy) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(p + r * cols));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(p + (r + kColsumTy) * cols));
      const float4 a2 = __ldg(reinterpret_cast<const float4*>(p + (r + 2 * kColsumTy) * cols));
      const float4 a3 = __ldg(reinterpret_cast<const float4*>(p + (r + 3 * kColsumTy) * cols));
      acc.x += (a0.x + a1.x) + (a2.x + a3.x);
      acc.y += (a0.y + a1.y) + (a2.y + a3.y);
      acc.z += (a0.z + a1.z) + (a2.z + a3.z);
      acc.w += (a0.w + a1.w) + (a2.w + a3.w);
    }
    for (; r < r_end; r += kColsumTy) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p + r * cols));
      acc.x += a.x;
      acc.y += a.y;
      acc.z += a.z;
      acc.w += a.w;
    }
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < cols) {
#pragma unroll
    for (int k = 1; k < kColsumTy; ++k) {
      const float4 o = red[k][tx];
      acc.x += o.x;
      acc.y += o.y;
      acc.z += o.z;
      acc.w += o.w;
    }
    *reinterpret_cast<float4*>(partial + static_cast<int64_t>(blockIdx.x) * cols + c) = acc;
  }
}

__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ partial, int slabs, int cols, float* __restrict__ out) {
  // 32 columns x 8 slab lanes per CTA; fixed summation order
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < cols)
    for (int k = ty; k < slabs; k += 8) s += partial[static_cast<int64_t>(k) * cols + c];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
#pragma unroll
    for (int k = 1; k < 8; ++k) s += red[k][tx];
    out[c] = s;
  }
}

static int colsum_slabs(int64_t rows, int cols) {
  const int col_tiles = (cols / 4 + kColsumTx - 1) / kColsumTx;
  int64_t slabs = (static_cast<int64_t>(kNumSMs) * 4 + col_tiles - 1) / col_tiles;  // ~4 CTAs per SM in total
  const int64_t max_slabs = (rows + 4 * kColsumTy - 1) / (4 * kColsumTy);
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  return static_cast<int>(slabs);
}

}  // namespace efgb

using namespace efgb;

extern "C" size_t efgb_colsum_workspace_bytes(int64_t rows, int cols) {
  if (rows <= 0 || cols <= 0) return 256;
  return align_up(static_cast<size_t>(colsum_slabs(rows, cols)) * cols * sizeof(float));
}

extern "C" int efgb_colsum(const float* x, int64_t rows, int cols, float* out, void* workspace, size_t workspace_bytes,
                           efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(rows >= 0 && cols >= 1 && cols % 4 == 0, EFGB_EINVAL, "colsum: cols must be a positive multiple of 4 (got %d)", cols);
  EFGB_REQUIRE(out, EFGB_EINVAL, "colsum: null output");
  if (rows == 0) {
    EFGB_CUDA_OK(cudaMemsetAsync(out, 0, static_cast<size_t>(cols) * sizeof(float), stream));
    return EFGB_OK;
  }
  EFGB_REQUIRE(x && workspace, EFGB_EINVAL, "colsum: null pointer");
  EFGB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, EFGB_EINVAL, "colsum: input must be 16-byte aligned");
  const int slabs = colsum_slabs(rows, cols);
  EFGB_REQUIRE(workspace_bytes >= static_cast<size_t>(slabs) * cols * sizeof(float), EFGB_EINVAL, "colsum: workspace too small");
  const int col_tiles = (cols / 4 + kColsumTx - 1) / kColsumTx;
  const int64_t rows_per_slab = (rows + slabs - 1) / slabs;
  float* partial = static_cast<float*>(workspace);
  colsum_partial_kernel<<<dim3(slabs, col_tiles), kColsumTx * kColsumTy, 0, stream>>>(x, rows, cols, rows_per_slab, partial);
  EFGB_LAUNCH_OK("colsum_partial_kernel");
  colsum_final_kernel<<<(cols + 31) / 32, 256, 0, stream>>>(partial, slabs, cols, out);
  EFGB_LAUNCH_OK("colsum_final_kernel");
  return EFGB_OK;
}
