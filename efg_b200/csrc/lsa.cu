// Linear sum assignment on the device, for the set-prediction matcher of Voxel-DETR / ConQueR.
//
// Reference: HungarianMatcher3d.forward (VD/modules/matcher.py:86-89) moves every cost matrix to the host and
// calls scipy.optimize.linear_sum_assignment — one full pipeline drain per training step, followed by a
// launch-bound tail while the GPU idles.  This kernel solves all matrices of a step where they are.
//
// The algorithm is scipy's (rectangular_lsap: shortest augmenting paths with dual variables, Crouse 2016), restated
// step for step so that the assignment — including how ties are broken — is the one scipy returns:
//   * the problem is transposed when it has more rows than columns; rows of the solver get assigned in order;
//   * every augmenting-path step scans the `remaining` columns in their array order (initially nc-1 .. 0, removal by
//     swapping the last one in) and picks the minimum reduced cost, preferring — among equal values — the LAST
//     unassigned column, else the FIRST column;
//   * all arithmetic in double, in scipy's association order: ((minVal + c) - u[i]) - v[j].
// One warp per matrix.  The scan is split over the 32 lanes (position it -> lane it % 32); a lane applies the
// sequential rule to its subsequence and a butterfly combines the lanes under the equivalent total order
// (lower value; then unassigned before assigned; larger position among unassigned, smaller among assigned).
// The cost matrix is staged in shared memory (float, solver orientation) when it fits.
// oracle: scipy itself (tests/test_gpu_lsa.py, exact equality, tie-heavy integer matrices included).
#include <math.h>

#include "common.cuh"

namespace efgb {

constexpr int kLsaMaxBatch = 32;

struct LsaProblem {
  const float* cost;
  int rows, cols, ld;
  long long out_off;
};
struct LsaBatch {
  LsaProblem p[kLsaMaxBatch];
};

__device__ __forceinline__ bool lsa_better(double oval, int oit, bool oun, double bval, int bit, bool bun) {
  if (oval < bval) return true;
  if (oval > bval) return false;
  if (oun && bun) return oit > bit;
  if (oun != bun) return oun;
  return oit < bit;
}

__global__ void __launch_bounds__(32)
lsa_kernel(const LsaBatch batch, long long* __restrict__ out_rows, long long* __restrict__ out_cols, int cost_in_smem,
           int* __restrict__ status) {
  const LsaProblem pr = batch.p[blockIdx.x];
  const int lane = threadIdx.x;
  const bool tr = pr.cols < pr.rows;
  const int nr = tr ? pr.cols : pr.rows, nc = tr ? pr.rows : pr.cols;
  if (nr <= 0) return;
  extern __shared__ __align__(16) unsigned char lsa_smem[];
  double* u = reinterpret_cast<double*>(lsa_smem);
  double* v = u + nr;
  double* spc = v + nc;  // shortest path costs
  int* path = reinterpret_cast<int*>(spc + nc);
  int* row4col = path + nc;
  int* remaining = row4col + nc;
  int* col4row = remaining + nc;
  unsigned char* SR = reinterpret_cast<unsigned char*>(col4row + nr);
  unsigned char* SC = SR + nr;
  float* cs = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(SC + nc) + 15) & ~static_cast<uintptr_t>(15));
  const float* __restrict__ cost = pr.cost;
  const int ld = pr.ld;

  if (cost_in_smem) {
    const int total = pr.rows * pr.cols;
    for (int s = lane; s < total; s += 32) {
      const int r = s / pr.cols, c = s - r * pr.cols;
      cs[tr ? c * nc + r : r * nc + c] = __ldg(cost + static_cast<long long>(r) * ld + c);
    }
  }
  for (int i = lane; i < nr; i += 32) {
    u[i] = 0.0;
    col4row[i] = -1;
  }
  for (int j = lane; j < nc; j += 32) {
    v[j] = 0.0;
    row4col[j] = -1;
    path[j] = -1;
  }
  __syncwarp();

  for (int cur = 0; cur < nr; ++cur) {
    for (int j = lane; j < nc; j += 32) {
      remaining[j] = nc - j - 1;
      spc[j] = INFINITY;
      SC[j] = 0;
    }
    for (int i = lane; i < nr; i += 32) SR[i] = 0;
    __syncwarp();
    double min_val = 0.0;
    int num_remaining = nc, sink = -1, i = cur;
    while (sink < 0) {
      if (lane == 0) SR[i] = 1;
      const double ui = u[i];
      double bval = INFINITY;
      int bit = -1;
      bool bun = false;
      for (int it = lane; it < num_remaining; it += 32) {
        const int j = remaining[it];
        const double cij = cost_in_smem ? static_cast<double>(cs[i * nc + j])
                                        : static_cast<double>(tr ? __ldg(cost + static_cast<long long>(j) * ld + i)
                                                                 : __ldg(cost + static_cast<long long>(i) * ld + j));
        const double r = __dsub_rn(__dsub_rn(__dadd_rn(min_val, cij), ui), v[j]);
        double s = spc[j];
        if (r < s) {
          path[j] = i;
          spc[j] = r;
          s = r;
        }
        const bool un = row4col[j] < 0;
        if (s < bval || (s == bval && un)) {
          bval = s;
          bit = it;
          bun = un;
        }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const double oval = __shfl_xor_sync(0xffffffffu, bval, d);
        const int oit = __shfl_xor_sync(0xffffffffu, bit, d);
        const bool oun = __shfl_xor_sync(0xffffffffu, static_cast<int>(bun), d) != 0;
        if (lsa_better(oval, oit, oun, bval, bit, bun)) {
          bval = oval;
          bit = oit;
          bun = oun;
        }
      }
      if (bit < 0 || !(bval < INFINITY)) {  // infeasible (non-finite costs): leave the remaining rows unassigned
        sink = -2;
        break;
      }
      min_val = bval;
      const int j = remaining[bit];
      const int r4 = row4col[j];
      if (r4 < 0)
        sink = j;
      else
        i = r4;
      __syncwarp();
      if (lane == 0) {
        SC[j] = 1;
        remaining[bit] = remaining[num_remaining - 1];
      }
      --num_remaining;
      __syncwarp();
    }
    if (sink < 0) break;
    // dual updates
    for (int i2 = lane; i2 < nr; i2 += 32) {
      if (i2 == cur)
        u[i2] = __dadd_rn(u[i2], min_val);
      else if (SR[i2])
        u[i2] = __dadd_rn(u[i2], __dsub_rn(min_val, spc[col4row[i2]]));
    }
    for (int j2 = lane; j2 < nc; j2 += 32)
      if (SC[j2]) v[j2] = __dsub_rn(v[j2], __dsub_rn(min_val, spc[j2]));
    __syncwarp();
    if (lane == 0) {  // augment along the path
      int j = sink;
      for (int guard = 0; guard <= nr; ++guard) {  // a path visits at most nr rows
        const int i2 = path[j];
        row4col[j] = i2;
        const int nxt = col4row[i2];
        col4row[i2] = j;
        j = nxt;
        if (i2 == cur) break;
      }
    }
    __syncwarp();
  }

  // Infeasible problem (non-finite costs; scipy raises "matrix contains invalid numeric entries" / "cost matrix is
  // infeasible"): flag it for the caller and complete the assignment with the free columns in ascending order, so
  // that every pair written below is a valid index — a diverged step must surface as a clean exception on the host
  // (ops.lsa_status), not as an out-of-bounds gather in the losses.
  if (lane == 0) {
    int next_free = 0, bad = 0;
    for (int r = 0; r < nr; ++r) {
      if (col4row[r] >= 0) continue;
      bad = 1;
      while (next_free < nc && row4col[next_free] >= 0) ++next_free;
      col4row[r] = next_free;
      row4col[next_free] = r;
    }
    if (bad && status) atomicOr(status, 1);
  }
  __syncwarp();

  // pairs sorted by the ORIGINAL row index, as scipy returns them
  long long* orow = out_rows + pr.out_off;
  long long* ocol = out_cols + pr.out_off;
  if (!tr) {
    for (int r = lane; r < nr; r += 32) {
      orow[r] = r;
      ocol[r] = col4row[r];
    }
  } else {
    for (int r = lane; r < nr; r += 32) {
      const int q = col4row[r];
      int rank = 0;
      for (int r2 = 0; r2 < nr; ++r2) rank += (col4row[r2] < q) ? 1 : 0;
      orow[rank] = q;
      ocol[rank] = r;
    }
  }
}

static size_t lsa_smem_bytes(int rows, int cols, bool with_cost) {
  const int nr = rows < cols ? rows : cols, nc = rows < cols ? cols : rows;
  size_t b = sizeof(double) * (static_cast<size_t>(nr) + 2 * nc) + sizeof(int) * (3 * static_cast<size_t>(nc) + nr) + nr + nc + 16;
  if (with_cost) b += sizeof(float) * static_cast<size_t>(nr) * nc;
  return b;
}

}  // namespace efgb

using namespace efgb;

extern "C" int efgb_lsa_batched_status(const void* const* cost_ptrs_host, const int32_t* rows_host, const int32_t* cols_host,
                                       const int32_t* ld_host, const int64_t* out_offsets_host, int count, int64_t* out_rows,
                                       int64_t* out_cols, int32_t* status, efgb_stream_t stream_);

extern "C" int efgb_lsa_batched(const void* const* cost_ptrs_host, const int32_t* rows_host, const int32_t* cols_host,
                                const int32_t* ld_host, const int64_t* out_offsets_host, int count, int64_t* out_rows,
                                int64_t* out_cols, efgb_stream_t stream_) {
  return efgb_lsa_batched_status(cost_ptrs_host, rows_host, cols_host, ld_host, out_offsets_host, count, out_rows, out_cols,
                                 nullptr, stream_);
}

extern "C" int efgb_lsa_batched_status(const void* const* cost_ptrs_host, const int32_t* rows_host, const int32_t* cols_host,
                                       const int32_t* ld_host, const int64_t* out_offsets_host, int count, int64_t* out_rows,
                                       int64_t* out_cols, int32_t* status, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(count >= 0, EFGB_EINVAL, "lsa_batched: negative count");
  if (count == 0) return EFGB_OK;
  EFGB_REQUIRE(cost_ptrs_host && rows_host && cols_host && ld_host && out_offsets_host && out_rows && out_cols, EFGB_EINVAL,
               "lsa_batched: null pointer");
  {
    static unsigned char configured[64] = {};   // per device: the attribute belongs to the device's context
    int dev = 0;
    EFGB_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      EFGB_CUDA_OK(cudaFuncSetAttribute(lsa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      if (dev >= 0 && dev < 64) configured[dev] = 1;
    }
  }
  for (int base = 0; base < count; base += kLsaMaxBatch) {
    const int n = count - base < kLsaMaxBatch ? count - base : kLsaMaxBatch;
    LsaBatch batch;
    memset(&batch, 0, sizeof(batch));
    size_t smem_cost = 0, smem_plain = 0;
    for (int k = 0; k < n; ++k) {
      const int rows = rows_host[base + k], cols = cols_host[base + k];
      EFGB_REQUIRE(rows >= 0 && cols >= 0 && ld_host[base + k] >= cols, EFGB_EINVAL, "lsa_batched: bad shape of problem %d", base + k);
      EFGB_REQUIRE(rows == 0 || cols == 0 || cost_ptrs_host[base + k], EFGB_EINVAL, "lsa_batched: null cost matrix %d", base + k);
      batch.p[k].cost = static_cast<const float*>(cost_ptrs_host[base + k]);
      batch.p[k].rows = rows;
      batch.p[k].cols = cols;
      batch.p[k].ld = ld_host[base + k];
      batch.p[k].out_off = out_offsets_host[base + k];
      const size_t a = lsa_smem_bytes(rows, cols, true), b = lsa_smem_bytes(rows, cols, false);
      if (a > smem_cost) smem_cost = a;
      if (b > smem_plain) smem_plain = b;
    }
    const bool in_smem = smem_cost <= 200 * 1024;
    const size_t smem = in_smem ? smem_cost : smem_plain;
    EFGB_REQUIRE(smem <= 227 * 1024, EFGB_EINVAL, "lsa_batched: a problem is too large for one CTA (%zu bytes of state)", smem);
    lsa_kernel<<<n, 32, smem, stream>>>(batch, reinterpret_cast<long long*>(out_rows), reinterpret_cast<long long*>(out_cols),
                                        in_smem ? 1 : 0, status);
    EFGB_LAUNCH_OK("lsa_kernel");
  }
  return EFGB_OK;
}
