// Micro-benchmark: throughput of TMA tile::gather4 row gathers (the candidate A producer of the sparse conv).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/_bin/gather4_bw scripts/micro/gather4_bw.cu
// Each CTA streams `stages` stages; a stage = TAPS sub-tiles of [128 rows][ROWB bytes], fetched by 128 lanes x TAPS
// gather4 instructions (4 rows each).  A consumer warp frees the slot as soon as the bytes have landed (no MMA).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint32_t a, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t phase) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(a), "r"(phase) : "memory");
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

constexpr int kTile = 128;

// idx: [tile][tap][128] int32.  Each producer lane owns 4 consecutive rows of one tap group.
template <int ROWB, int TAPS, int PW>
__global__ void __launch_bounds__(PW * 32 + 64) gather_kernel(const __grid_constant__ CUtensorMap map, const int32_t* __restrict__ idx,
                                                     int num_tiles, int taps_total, int ring, uint8_t* dump) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kStageBytes = TAPS * kTile * ROWB;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + ring * kStageBytes);
  uint64_t* empty = full + ring;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ring; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int groups = taps_total / TAPS;   // stages per tile
  // per-tile rulebook block [taps_total][128] int32, double-buffered in shared memory by bulk copies one tile ahead
  int32_t* idx_s = reinterpret_cast<int32_t*>(smem + ring * kStageBytes + 2 * ring * 8 + 64);
  uint64_t* idx_full = empty + ring;        // [2]
  uint64_t* idx_empty = idx_full + 2;       // [2]
  const int idx_bytes = taps_total * kTile * 4;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&idx_full[i]), 1); mbar_init(smem_u32(&idx_empty[i]), PW * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp < PW) {
    int slot = 0; uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int ib = it & 1;
      const uint32_t iph = (it >> 1) & 1;
      mbar_wait(smem_u32(&idx_full[ib]), iph);
      const int32_t* my = idx_s + ib * taps_total * kTile;
      for (int g = 0; g < groups; ++g) {
        mbar_wait(smem_u32(&empty[slot]), phase ^ 1);
        const uint32_t bar = smem_u32(&full[slot]);
        if (threadIdx.x == 0) mbar_expect(bar, kStageBytes);
        const uint32_t base = smem_u32(smem + slot * kStageBytes);
#pragma unroll
        for (int u0 = 0; u0 < TAPS * 32; u0 += PW * 32) {
          const int u = u0 + threadIdx.x;
          if (u < TAPS * 32) {
            const int t = u >> 5, q = u & 31;
            const int4 r = *reinterpret_cast<const int4*>(my + (g * TAPS + t) * kTile + q * 4);
            if (ROWB <= 128) {
              gather4(base + t * kTile * ROWB + q * 4 * ROWB, &map, 0, r.x, r.y, r.z, r.w, bar);
            } else {
#pragma unroll
              for (int cb = 0; cb < ROWB / 128; ++cb)
                gather4(base + (t * (ROWB / 128) + cb) * kTile * 128 + q * 4 * 128, &map, cb * 64, r.x, r.y, r.z, r.w, bar);
            }
          }
        }
        if (++slot == ring) { slot = 0; phase ^= 1; }
      }
      mbar_arrive(smem_u32(&idx_empty[ib]));
    }
  } else if (warp == PW) {
    int slot = 0; uint32_t phase = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int g = 0; g < groups; ++g) {
        mbar_wait(smem_u32(&full[slot]), phase);
        if (dump && first && blockIdx.x == 74) {
          for (int i = lane; i < kStageBytes / 16; i += 32)
            reinterpret_cast<int4*>(dump)[i] = reinterpret_cast<const int4*>(smem + slot * kStageBytes)[i];
          first = false;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty[slot]));
        if (++slot == ring) { slot = 0; phase ^= 1; }
      }
    }
  } else if (lane == 0) {
    // rulebook loader
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int ib = it & 1;
      const uint32_t iph = (it >> 1) & 1;
      mbar_wait(smem_u32(&idx_empty[ib]), iph ^ 1);
      const uint32_t bar = smem_u32(&idx_full[ib]);
      mbar_expect(bar, idx_bytes);
      asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(idx_s + ib * taps_total * kTile)), "l"(idx + (size_t)tile * taps_total * kTile), "r"(idx_bytes), "r"(bar) : "memory");
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROWB, int TAPS, int PW>
void run(EncodeFn encode, int M, int taps_total, float valid, bool check, bool zero_row = false) {
  const int num_tiles = (M + kTile - 1) / kTile;
  // planes: [M][ROWB bytes] of bf16; value = row * 64 + elem (mod 2^16) so that the landing layout can be checked
  const int cols = ROWB / 2;
  std::vector<uint16_t> h((size_t)(M + 1) * cols, 0);
  for (int r = 0; r < M; ++r) for (int c = 0; c < cols; ++c) h[(size_t)r * cols + c] = (uint16_t)(r * 7 + c);
  uint16_t* d; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  // SubM-like rulebook: neighbours at row + off[tap], a fraction valid
  std::vector<int32_t> hi((size_t)num_tiles * taps_total * kTile);
  const int offs3[3] = {-1, 0, 1};
  srand(1);
  for (int tile = 0; tile < num_tiles; ++tile)
    for (int t = 0; t < taps_total; ++t) {
      const int off = offs3[t % 3] + offs3[(t / 3) % 3] * 61 + offs3[(t / 9) % 3] * 5003;
      for (int r = 0; r < kTile; ++r) {
        const int row = tile * kTile + r, src = row + off;
        const bool ok = row < M && src >= 0 && src < M && (t == taps_total / 2 || (rand() % 1000) < valid * 1000);
        hi[((size_t)tile * taps_total + t) * kTile + r] = ok ? src : (zero_row ? M : -1);
      }
    }
  int32_t* di; CK(cudaMalloc(&di, hi.size() * 4)); CK(cudaMemcpy(di, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap map;
  const int box_cols = ROWB <= 128 ? cols : 64;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)M + 1};
  cuuint64_t gstride[1] = {(cuuint64_t)ROWB};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, 1};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (box_cols * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult rc = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); exit(1); }
  constexpr int kStageBytes = TAPS * kTile * ROWB;
  const int ring = 190 * 1024 / kStageBytes;
  const size_t smem = 1024 + (size_t)ring * kStageBytes + ring * 16 + 128 + 2 * (size_t)taps_total * kTile * 4;
  CK(cudaFuncSetAttribute(gather_kernel<ROWB, TAPS, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  uint8_t* dump = nullptr;
  if (check) CK(cudaMalloc(&dump, kStageBytes));
  const int grid = num_tiles < 148 ? num_tiles : 148;
  gather_kernel<ROWB, TAPS, PW><<<grid, PW * 32 + 64, smem>>>(map, di, num_tiles, taps_total, ring, dump);
  CK(cudaDeviceSynchronize());
  if (check) {
    std::vector<uint16_t> got(kStageBytes / 2);
    CK(cudaMemcpy(got.data(), dump, kStageBytes, cudaMemcpyDeviceToHost));
    // expected: sub-tile t, row r at byte r * rowpitch, 16-byte chunk c at position c ^ f(r)
    const int pitch = ROWB <= 128 ? ROWB : 128;
    const int chunks = pitch / 16;
    long bad = 0, bad_plain = 0;
    for (int t = 0; t < TAPS * (ROWB <= 128 ? 1 : ROWB / 128); ++t)
      for (int r = 0; r < kTile; ++r) {
        const int tap = ROWB <= 128 ? t : t / (ROWB / 128), cb = ROWB <= 128 ? 0 : t % (ROWB / 128);
        const int src = hi[((size_t)74 * taps_total + tap) * kTile + r];
        for (int c = 0; c < chunks; ++c)
          for (int e = 0; e < 8; ++e) {
            const uint16_t want = (src < 0 || src >= M) ? 0 : (uint16_t)(src * 7 + cb * 64 + c * 8 + e);
            // swizzle: chunk ^= (row % 8) >> (3 - log2(chunks))   [128B: r&7, 64B: (r>>1)&3, 32B: (r>>2)&1]
            const int shift = chunks == 8 ? 0 : (chunks == 4 ? 1 : 2);
            const int pc = c ^ ((r & 7) >> shift);
            if (got[((size_t)t * kTile * pitch + r * pitch + pc * 16) / 2 + e] != want) ++bad;
            if (got[((size_t)t * kTile * pitch + r * pitch + c * 16) / 2 + e] != want) ++bad_plain;
          }
      }
    printf("  layout check ROWB=%d: mismatches with swizzle model %ld, without swizzle %ld\n", ROWB, bad, bad_plain);
  }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 20;
  for (int i = 0; i < 3; ++i) gather_kernel<ROWB, TAPS, PW><<<grid, PW * 32 + 64, smem>>>(map, di, num_tiles, taps_total, ring, nullptr);
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) gather_kernel<ROWB, TAPS, PW><<<grid, PW * 32 + 64, smem>>>(map, di, num_tiles, taps_total, ring, nullptr);
  cudaEventRecord(b); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double us = ms * 1e3 / iters;
  const double smem_bytes = (double)num_tiles * taps_total * kTile * ROWB;
  printf("PW=%d zero_row=%d M=%d ROWB=%d taps/stage=%d valid=%.2f ring=%d: %.1f us, %.1f GB/s into smem (%.1f B/clk/SM at 1.9 GHz), %.0f cycles per 128-row tap\n",
         PW, (int)zero_row, M, ROWB, TAPS, valid, ring, us, smem_bytes / us / 1e3, smem_bytes / us / 1e3 / 148 / 1.9, us * 1.9e3 * grid / ((double)num_tiles * taps_total));
  cudaFree(d); cudaFree(di); if (dump) cudaFree(dump);
}

int main() {
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&encode), cudaEnableDefault, &q));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  run<64, 4, 4>(encode, 146575, 28, 1.0f, true);
  run<64, 4, 4>(encode, 146575, 28, 0.5f, false);
  run<64, 4, 4>(encode, 146575, 28, 0.5f, true, true);
  run<64, 8, 8>(encode, 146575, 32, 1.0f, false);
  run<64, 8, 8>(encode, 146575, 32, 0.5f, false, true);
  run<64, 8, 16>(encode, 146575, 32, 1.0f, false);
  run<64, 8, 16>(encode, 146575, 32, 0.5f, false, true);
  run<128, 4, 4>(encode, 146575, 28, 1.0f, true);
  run<128, 4, 8>(encode, 146575, 28, 1.0f, false);
  run<128, 4, 16>(encode, 146575, 28, 1.0f, false);
  run<256, 2, 4>(encode, 67950, 28, 1.0f, true);
  run<256, 2, 8>(encode, 67950, 28, 1.0f, false);
  run<256, 2, 16>(encode, 67950, 28, 1.0f, false);
  run<256, 2, 16>(encode, 67950, 28, 0.55f, false, true);
  return 0;
}
