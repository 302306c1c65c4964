// Error plumbing + device-wide scans shared by the voxelizer and the rulebook builder.
#include <stdarg.h>

#include "common.cuh"

namespace efgb {

static thread_local char g_err[512] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- scan ------------------------------------------------------------------------------------
struct LoadU32 {
  const uint32_t* p;
  __device__ __forceinline__ uint32_t operator()(int64_t i) const { return p[i]; }
};
struct StoreU32 {
  uint32_t* p;
  __device__ __forceinline__ void operator()(int64_t i, uint32_t v) const { p[i] = v; }
};
struct LoadCellPop {
  const CellWord* p;
  __device__ __forceinline__ uint32_t operator()(int64_t i) const { return __popc(p[i].bits); }
};
struct StoreCellPrefix {
  CellWord* p;
  __device__ __forceinline__ void operator()(int64_t i, uint32_t v) const { p[i].prefix = v; }
};

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// Block-wide exclusive scan of one value per thread (blockDim.x multiple of 32, <= 1024).
// Returns the exclusive prefix; *block_total gets the block sum (valid in every thread).
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* block_total) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t total_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  uint32_t incl = warp_inclusive_scan(v, lane);
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t ws = lane < nwarps ? warp_sums[lane] : 0u;
    uint32_t wi = warp_inclusive_scan(ws, lane);
    if (lane < nwarps) warp_sums[lane] = wi - ws;  // exclusive warp offsets
    if (lane == 31) total_s = wi;
  }
  __syncthreads();
  uint32_t excl = incl - v + warp_sums[wid];
  *block_total = total_s;
  __syncthreads();  // smem reusable by the next call
  return excl;
}

template <typename Load>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(Load load, int64_t n, uint32_t* tile_sums) {
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kScanTile + static_cast<int64_t>(threadIdx.x) * kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + j;
    if (i < n) s += load(i);
  }
  uint32_t total;
  (void)block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// One block; exclusive scan of tile_sums[0..nb) in place, grand total to tile_sums[nb] and *total_out.
__global__ void __launch_bounds__(1024) scan_tiles_kernel(uint32_t* tile_sums, int64_t nb, uint32_t* total_out) {
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t start = 0; start < nb; start += blockDim.x) {
    int64_t i = start + threadIdx.x;
    uint32_t v = i < nb ? tile_sums[i] : 0u;
    uint32_t total;
    uint32_t excl = block_exclusive_scan(v, &total);
    uint32_t carry = carry_s;
    if (i < nb) tile_sums[i] = excl + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    tile_sums[nb] = carry_s;
    if (total_out) *total_out = carry_s;
  }
}

template <typename Load, typename Store>
__global__ void __launch_bounds__(kScanThreads)
scan_downsweep_kernel(Load load, Store store, int64_t n, const uint32_t* tile_sums, int64_t nb, bool write_total) {
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kScanTile + static_cast<int64_t>(threadIdx.x) * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + j;
    v[j] = i < n ? load(i) : 0u;
    s += v[j];
  }
  uint32_t total;
  uint32_t excl = block_exclusive_scan(s, &total) + tile_sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + j;
    if (i < n) store(i, excl);
    excl += v[j];
  }
  if (write_total && blockIdx.x == 0 && threadIdx.x == 0) store(n, tile_sums[nb]);
}

template <typename Load, typename Store>
static int scan_impl(Load load, Store store, int64_t n, uint32_t* scratch, uint32_t* total_dev, bool write_total,
                     cudaStream_t stream) {
  if (n <= 0) {
    // total = 0; still define out[0] / total
    if (total_dev) EFGB_CUDA_OK(cudaMemsetAsync(total_dev, 0, sizeof(uint32_t), stream));
    return EFGB_OK;
  }
  const int64_t nb = (n + kScanTile - 1) / kScanTile;
  EFGB_REQUIRE(nb < (1ll << 31), EFGB_ERANGE, "scan: %lld elements is too many", (long long)n);
  scan_reduce_kernel<<<static_cast<unsigned>(nb), kScanThreads, 0, stream>>>(load, n, scratch);
  EFGB_LAUNCH_OK("scan_reduce_kernel");
  scan_tiles_kernel<<<1, 1024, 0, stream>>>(scratch, nb, total_dev);
  EFGB_LAUNCH_OK("scan_tiles_kernel");
  scan_downsweep_kernel<<<static_cast<unsigned>(nb), kScanThreads, 0, stream>>>(load, store, n, scratch, nb,
                                                                               write_total);
  EFGB_LAUNCH_OK("scan_downsweep_kernel");
  return EFGB_OK;
}

int scan_exclusive_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* scratch, cudaStream_t stream) {
  if (n <= 0) {
    EFGB_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(uint32_t), stream));
    return EFGB_OK;
  }
  return scan_impl(LoadU32{in}, StoreU32{out}, n, scratch, nullptr, /*write_total=*/true, stream);
}

int cells_scan(CellWord* cells, int64_t num_words, uint32_t* total_dev, uint32_t* scratch, cudaStream_t stream) {
  return scan_impl(LoadCellPop{cells}, StoreCellPrefix{cells}, num_words, scratch, total_dev, /*write_total=*/false,
                   stream);
}

}  // namespace efgb

extern "C" const char* efgb_last_error(void) { return efgb::g_err; }
extern "C" int efgb_version(void) { return 100; }
extern "C" uint64_t efgb_launch_count(void) { return efgb::g_launch_count; }
