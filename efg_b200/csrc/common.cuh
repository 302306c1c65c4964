// Shared device/host helpers for the efgb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/efgb200.h"

namespace efgb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define EFGB_REQUIRE(cond, code, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      ::efgb::set_error(__VA_ARGS__);  \
      return (code);                   \
    }                                  \
  } while (0)

#define EFGB_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::efgb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                      \
      return EFGB_ECUDA;                                                                \
    }                                                                                   \
  } while (0)

extern unsigned long long g_launch_count;  // kernels launched by this library (bench.py reports it)

#define EFGB_LAUNCH_OK(name)                                                                  \
  do {                                                                                        \
    ++::efgb::g_launch_count;                                                                 \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      ::efgb::set_error("launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(_e),      \
                        __FILE__, __LINE__);                                                  \
      return EFGB_ECUDA;                                                                      \
    }                                                                                         \
  } while (0)

inline cudaStream_t as_stream(efgb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Workspace {
  char* base;
  size_t size;
  size_t used;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), used(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    if (used + bytes > size) return nullptr;
    T* r = reinterpret_cast<T*>(base + used);
    used += bytes;
    return r;
  }
};

inline int grid_for(int64_t n, int block, int max_blocks = kNumSMs * 32) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return static_cast<int>(g);
}

// ---- device-wide exclusive scan of u32 (out has n+1 entries; out[n] = total) ------------------
// The element is produced by a functor so callers can scan popcounts / flags without a
// materialised temporary.  Scratch: scan_scratch_elems(n) u32.
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

inline size_t scan_scratch_elems(int64_t n) { return static_cast<size_t>((n + kScanTile - 1) / kScanTile + 2); }

int scan_exclusive_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* scratch, cudaStream_t stream);

// ---- occupancy cells: one (bits, prefix) pair per 32 linear grid cells ------------------------
// rank(cell) = prefix[cell>>5] + popc(bits & lanemask_lt(cell&31)) is the position of an
// occupied cell in ascending linear order — the device rulebook needs no sort and no hash.
struct __align__(8) CellWord {
  uint32_t bits;
  uint32_t prefix;
};

int cells_scan(CellWord* cells, int64_t num_words, uint32_t* total_dev, uint32_t* scratch, cudaStream_t stream);

__device__ __forceinline__ int cell_rank(const CellWord* __restrict__ cells, uint32_t cell) {
  CellWord w = cells[cell >> 5];
  uint32_t bit = 1u << (cell & 31);
  if (!(w.bits & bit)) return -1;
  return static_cast<int>(w.prefix + __popc(w.bits & (bit - 1)));
}

}  // namespace efgb
