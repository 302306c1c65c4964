// Sparse convolution forward / dgrad on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Output-stationary implicit GEMM over the device rulebook:
//     D[128 output rows x N] = sum over K = (tap, channel)  A[row, K] * B[n, K]
//   A[row, (tap, c)] = in[nbr[row, tap], c]     gathered on the fly (zeros where nbr = -1)
//   B[n,   (tap, c)] = weight element           pre-packed into the exact shared-memory image
//                                               (K-major, 128-byte swizzle) the tensor core reads
// K is consumed in chunks of one 128-byte swizzled row per operand row: 32 tf32 or 64 bf16 values.
//
// Roles inside one persistent CTA (one CTA per SM, super-tiles strided over the grid):
//   warps 0-15  A producers.  All 16 warps fill the SAME stage (a thread owns two 16-byte pieces of the
//               128 x 128 B tile) and the loop is software-pipelined in registers: the gathers of stage s+1 and
//               the rulebook entries of stage s+2 are in flight while stage s is converted and stored, so
//               neither latency sits on the stage turnaround and a thread spends ~40 instructions per piece
//               (round 1: four warp groups on four stages, one exposed latency per stage, ~70 instructions per
//               piece — producers were issue- and latency-bound, profiles/r1_ncu_source_hotspots_spconv_tc.txt)
//   warp  16    MMA issuer: one elected lane issues tcgen05.mma (M=128, N, 32 bytes of K) into TMEM;
//               tcgen05.commit releases the stage / publishes the accumulator
//   warp  17    B loader: one lane streams the packed weight chunk with cp.async.bulk (bulk-copy engine,
//               mbarrier complete_tx, no tensor map) — weights stay L2-resident
//   warps 18-21 epilogue: tcgen05.ld the fp32 accumulator (double-buffered in TMEM), add bias, store rows
//
// Precision modes (all accumulate in fp32 in TMEM):
//   kBf16x3 (default)  a = a_hi + a_lo, b = b_hi + b_lo in bf16; a_hi*b_hi + a_hi*b_lo + a_lo*b_hi on kind::f16.
//                      Error 2.5e-5 relative (scripts/numerics_split_precision.py), twice the MMA rate and half the
//                      shared-memory operand bytes of 3xTF32.
//   kTf32x3            the same three products with tf32 hi / lo parts (error ~2^-21, "fp32-faithful")
//   kTf32              single pass
// "Concatenated" issue (n_cta <= 128): [B_hi | B_lo] are adjacent in shared memory, so A_hi x [B_hi | B_lo]
// is ONE MMA of N = 2 n_cta whose two column halves the epilogue adds; with A_lo x B_hi that is two reads of
// the A tile per K step instead of three and 2/3 of the MMA instructions for the same tensor-pipe work.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace efgb {
namespace tc {

#if EFGB_TC_TRACE
constexpr int kTraceStages = 2048;
__device__ long long g_trace[kTraceStages * 12];
#define EFGB_TRACE(stage, slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && (stage) < kTraceStages) g_trace[(stage) * 12 + (slot)] = clock64(); } while (0)
#else
#define EFGB_TRACE(stage, slot) do { } while (0)
#endif

constexpr int kTileM = 128;
constexpr int kChunkK = 32;             // tf32 values per K chunk (128 bytes)
constexpr int kBf16ChunkK = 64;         // bf16 values per K chunk (128 bytes)
#ifndef EFGB_TC_PRODUCER_WARPS
#define EFGB_TC_PRODUCER_WARPS 16
#endif
constexpr int kProducerWarps = EFGB_TC_PRODUCER_WARPS;   // gather / split warps (8 or 16)
constexpr int kEpiStageBytes = 32 * 16 * 4;                // per epilogue warp: a 32-row x 16-column fp32 block
constexpr int kMmaWarp = kProducerWarps;
constexpr int kLoaderWarp = kProducerWarps + 1;
constexpr int kThreads = (kProducerWarps + 6) * 32;      // + MMA issuer, weight loader, 4 epilogue warps
static_assert(kProducerWarps == 8 || kProducerWarps == 16, "producer warp groups assume 8 or 16 warps");
static_assert((kProducerWarps + 2) % 4 == 2, "epilogue warps must cover the four TMEM lane quarters");
constexpr int kMaxTaps = 32;

enum : int { kTf32 = 0, kTf32x3 = 1, kBf16x3 = 2 };
// Perf ablations for A/B library variants only (efg_b200/_build.py VARIANTS; results are invalid): 1 = MMAs skipped,
// 2 = planes producers skip the copies, 4 = planes producers skip the rulebook loads.  The product build defines nothing.
// EFGB_TC_TRACE=1 (diagnostic build only): CTA 0 writes clock64() stamps of every pipeline stage into g_trace —
// producer warp 0: [0] before the a_empty wait, [1] after it, [2] copies issued, [3] stage n - depth completed;
// MMA warp: [4] a_full acquired, [5] MMAs + commit issued.  Read back with efgb_debug_trace_read.
#ifndef EFGB_TC_TRACE
#define EFGB_TC_TRACE 0
#endif
#ifndef EFGB_TC_SB_MID
#define EFGB_TC_SB_MID 3    // weight-ring depth for 16..63 KB chunks
#endif
#ifndef EFGB_TC_MAX_T
#define EFGB_TC_MAX_T 8     // row tiles per super-tile (accumulators side by side in TMEM)
#endif
#ifndef EFGB_TC_ABLATE
#define EFGB_TC_ABLATE 0
#endif
// Where the generic -> async proxy fence runs: 0 = in the producers (after their stores), 1 = in the MMA issuer (after its
// acquire on the stage barrier).
#ifndef EFGB_TC_FENCE_AT_MMA
#define EFGB_TC_FENCE_AT_MMA 1
#endif   // `split` argument of the C ABI

template <int kMode>
struct ModeTraits {
  static constexpr int kParts = kMode == kTf32 ? 1 : 2;                    // operand images per stage (hi, lo)
  static constexpr int kChunkVals = kMode == kBf16x3 ? kBf16ChunkK : kChunkK;  // K values per 128-byte row
  static constexpr int kPieceVals = kChunkVals / 8;                        // K values per 16-byte shared-memory piece
  static constexpr int kLoads = kPieceVals / 4;                            // LDG.128 per piece
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// Ampere-style asynchronous 16-byte copy global -> shared, L2 only; src_bytes = 0 writes zeros (missing neighbour).
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }
// One lane of a converged warp (elect.sync): the compiler keeps warp-uniform operands in uniform registers.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
template <int kMode>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if constexpr ((EFGB_TC_ABLATE & 1) != 0) return;
  if constexpr (kMode == kBf16x3)
    tc_mma_bf16(tmem_d, adesc, bdesc, idesc, accum);
  else
    tc_mma_tf32(tmem_d, adesc, bdesc, idesc, accum);
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzle shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp:SmemDescriptor):
// start>>4 [0,14), LBO>>4 [16,30) (unused for swizzled K-major), SBO>>4 [32,46) = 1024 B between 8-row
// groups, version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor (mma_sm100_desc.hpp:InstrDescriptor): D=f32 [4,6)=1, A format [7,10), B format [10,13)
// (kind::tf32: 2 = tf32; kind::f16: 1 = bf16), both K-major, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
template <int kMode>
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return kMode == kBf16x3 ? make_idesc_bf16(m, n) : make_idesc_tf32(m, n);
}

struct Params {
  const float* in;       // [num_in, c_red]
  const uint8_t* planes; // bf16x3 only, nullable: [num_in][hi | lo][c_red] bf16 — the input pre-split once (split_bf16_kernel)
  const uint8_t* packed; // [chunks][parts][n_out][128 B] swizzled image
  const float* bias;     // [n_out] or null
  int relu;              // epilogue applies max(x, 0) after the bias
  const int32_t* nbr;    // [num_out, taps], or null = identity (dense GEMM: taps == 1, src row = out row)
  float* out;            // [num_out, n_out]
  int64_t num_out;
  int c_red, taps, n_out, chunks;
  int num_tiles;         // 128-row tiles
  int n_cta;             // output columns per CTA (n_out / gridDim.y)
  int acc_cols;          // TMEM columns of one tile accumulator: n_cta, or 2 n_cta with concatenated issue
  int concat;            // 1: A_hi x [B_hi | B_lo] as one MMA of N = 2 n_cta (+ A_lo x B_hi); 0: three MMAs of N = n_cta
  int tiles_per_super;   // T: row tiles that share every weight chunk (accumulators side by side in TMEM)
  int num_super;         // ceil(num_tiles / T)
  int wide;              // bf16x3 register path: 8-value units (input 32-byte aligned)
  int epi_staged;        // epilogue through the shared-memory staging buffers (coalesced stores); 0: row-per-thread stores
  int sa, sb;            // A-ring / B-ring depth
  int cred_shift;        // log2(c_red) when c_red is a power of two, else -1
};

// Work of one CTA: super-tiles st = blockIdx.x + i * gridDim.x; a super-tile is T consecutive 128-row tiles
// (the globally last one may hold fewer).  The A-stage stream of the CTA is ordered (super-tile, chunk, tile).
struct CtaWork {
  int n_super;      // super-tiles of this CTA
  int t_last;       // tiles in this CTA's last super-tile
  int total_a;      // A stages of this CTA
  __device__ __forceinline__ CtaWork(const Params& p) {
    n_super = (p.num_super - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    if (n_super < 0) n_super = 0;
    t_last = p.tiles_per_super;
    if (n_super > 0) {
      const int st_last = static_cast<int>(blockIdx.x) + (n_super - 1) * static_cast<int>(gridDim.x);
      const int rem = p.num_tiles - st_last * p.tiles_per_super;
      if (rem < t_last) t_last = rem;
    }
    total_a = n_super > 0 ? ((n_super - 1) * p.tiles_per_super + t_last) * p.chunks : 0;
  }
  __device__ __forceinline__ int tiles_in(const Params& p, int i) const { return i == n_super - 1 ? t_last : p.tiles_per_super; }
};

// fp32 -> operand image conversion of one 4-value unit (one LDG.128 of a source row).
//   kTf32:   the raw words (the tensor core ignores the 13 low mantissa bits)
//   kTf32x3: hi = word & 0xFFFFE000, lo = v - hi (exact)                       16 bytes per image
//   kBf16x3: hi = bf16_rn(v), lo = bf16_rn(v - hi)                              8 bytes per image
template <int kMode>
__device__ __forceinline__ void convert_store(uint8_t* a_hi, uint8_t* a_lo, uint32_t off, const float4& v) {
  if constexpr (kMode == kTf32) {
    *reinterpret_cast<float4*>(a_hi + off) = v;
  } else if constexpr (kMode == kTf32x3) {
    float4 hi, lo;
    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    lo.x = v.x - hi.x;
    lo.y = v.y - hi.y;
    lo.z = v.z - hi.z;
    lo.w = v.w - hi.w;
    *reinterpret_cast<float4*>(a_hi + off) = hi;
    *reinterpret_cast<float4*>(a_lo + off) = lo;
  } else {
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y);
    const __nv_bfloat162 h23 = __floats2bfloat162_rn(v.z, v.w);
    const uint32_t w01 = *reinterpret_cast<const uint32_t*>(&h01), w23 = *reinterpret_cast<const uint32_t*>(&h23);
    const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - __uint_as_float(w01 << 16), v.y - __uint_as_float(w01 & 0xFFFF0000u));
    const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - __uint_as_float(w23 << 16), v.w - __uint_as_float(w23 & 0xFFFF0000u));
    *reinterpret_cast<uint2*>(a_hi + off) = make_uint2(w01, w23);
    *reinterpret_cast<uint2*>(a_lo + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
  }
}

// A producers: see the role description at the top of the file.  Eight consecutive lanes own one row of the stage and
// read it 128 contiguous bytes per request (lane l: the 4-value units l, l + 8, ... of the row's chunk), so a warp-wide
// LDG.128 touches four 128-byte lines — the L1TEX wavefront count, which co-limited the round-1 producers with the
// issue rate, is the minimum the gather allows.  A thread serves two rows (slot and slot + 64); its units have the
// same tap / first channel in both rows.  Row slots are permuted inside a warp (0, 4, 1, 5 | 2, 6, 3, 7) so that the
// two rows of a half-warp fall into different halves of the 128-byte swizzle: the 8-byte bf16 stores are conflict-free.
template <int kMode>
__device__ __forceinline__ void produce_a(const Params& p, const CtaWork& w, uint8_t* a_ring, uint64_t* a_full,
                                          uint64_t* a_empty, const int tid, const int lane) {
  using M = ModeTraits<kMode>;
  static_assert(kProducerWarps == 16, "the row-slot mapping assumes 64 row slots of 8 lanes");
  constexpr int kJ = M::kChunkVals / 32;      // units per row per thread (8 lanes x 4 values = 32 values per pass)
  const int a_part = kTileM * 128;
  const int a_bytes = M::kParts * a_part;
  const int l = tid & 7;
  const int slot_id = tid >> 3;               // 0..63
  const int wv = slot_id >> 2, jj = slot_id & 3;
  const int r_lo = (wv >> 1) * 8 + (wv & 1) * 2 + (jj >> 1) + (jj & 1) * 4;   // row of this thread in the first 64 rows
  uint32_t off[kJ];
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    const int u = l + 8 * j;
    if constexpr (kMode == kBf16x3)
      off[j] = static_cast<uint32_t>(r_lo) * 128u + (static_cast<uint32_t>((u >> 1) ^ (r_lo & 7)) << 4) + static_cast<uint32_t>(u & 1) * 8u;
    else
      off[j] = static_cast<uint32_t>(r_lo) * 128u + (static_cast<uint32_t>(u ^ (r_lo & 7)) << 4);
  }
  const int total = w.total_a;
  if (total == 0) return;
  const uint32_t c_red = static_cast<uint32_t>(p.c_red);
  const bool one_tap = (p.c_red % M::kChunkVals) == 0;   // every unit of a chunk belongs to the same tap

  // cursor of the rulebook prefetch over this CTA's stage stream (super-tile, chunk, tile), no divisions
  int cu_s = 0, cu_c = 0, cu_t = 0, cu_ti = w.tiles_in(p, 0);
  auto load_idx = [&](int32_t (&src)[kJ][2], uint32_t (&ci)[kJ]) {
    const int tile = (static_cast<int>(blockIdx.x) + cu_s * static_cast<int>(gridDim.x)) * p.tiles_per_super + cu_t;
    const int row0 = tile * kTileM + r_lo;
    const bool ok0 = row0 < p.num_out, ok1 = row0 + 64 < p.num_out;
    const int32_t* n0 = p.nbr + static_cast<int64_t>(row0) * p.taps;
    const int32_t* n1 = n0 + 64 * p.taps;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      if (j > 0 && one_tap) {   // same rulebook entries as unit 0 (read from src[0] at gather time: no register copy of a pending load)
        ci[j] = ci[0] + 32u * j;
        continue;
      }
      const int kk = cu_c * M::kChunkVals + 4 * (l + 8 * j);
      const int tap = p.cred_shift >= 0 ? (kk >> p.cred_shift) : kk / p.c_red;
      ci[j] = static_cast<uint32_t>(kk - tap * p.c_red);
      const bool tap_ok = tap < p.taps;
      src[j][0] = (tap_ok && ok0) ? (p.nbr ? __ldg(n0 + tap) : row0) : -1;
      src[j][1] = (tap_ok && ok1) ? (p.nbr ? __ldg(n1 + tap) : row0 + 64) : -1;
    }
    if (++cu_t == cu_ti) {
      cu_t = 0;
      if (++cu_c == p.chunks) {
        cu_c = 0;
        ++cu_s;
        cu_ti = w.tiles_in(p, cu_s);
      }
    }
  };
  auto gather = [&](const int32_t (&src)[kJ][2], const uint32_t (&ci)[kJ], float4 (&v)[kJ][2]) {
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        v[j][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        // 32-bit element offset (the host checks num_in * c_red < 2^32)
        const int32_t e = (j > 0 && one_tap) ? src[0][i] : src[j][i];
        if (e >= 0)
          v[j][i] = __ldg(reinterpret_cast<const float4*>(p.in + static_cast<size_t>(static_cast<uint32_t>(e) * c_red + ci[j])));
      }
    }
  };
  int slot = 0;
  uint32_t phase = 0;
  [[maybe_unused]] int t_fin = 0;
  auto finish = [&](const float4 (&v)[kJ][2]) {
    if (tid == 0) EFGB_TRACE(t_fin, 0);
    mbar_wait(smem_u32(&a_empty[slot]), phase ^ 1);
    if (tid == 0) EFGB_TRACE(t_fin, 1);
    uint8_t* a_hi = a_ring + static_cast<size_t>(slot) * a_bytes;
    uint8_t* a_lo = a_hi + a_part;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      convert_store<kMode>(a_hi, a_lo, off[j], v[j][0]);
      convert_store<kMode>(a_hi, a_lo, off[j] + 64u * 128u, v[j][1]);
    }
    if (tid == 0) EFGB_TRACE(t_fin, 2);
    // No fence.proxy.async here: it lowers to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and the MEMBAR waits for EVERY
    // outstanding memory operation of the thread — including the gathers of the next stage, which serialised the
    // software pipeline on one full memory latency per stage (measured: ~1600 cycles per stage whatever the producer
    // did).  The stores are published by the release of mbarrier.arrive; the MMA issuer, which has no loads in
    // flight, runs the proxy fence after its acquire on a_full.
#if !EFGB_TC_FENCE_AT_MMA
    fence_proxy_async();
#endif
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&a_full[slot]));
    if (tid == 0) {
      EFGB_TRACE(t_fin, 3);
      EFGB_TRACE(t_fin, 9);
      EFGB_TRACE(t_fin, 10);
    }
    ++t_fin;
    if (++slot == p.sa) {
      slot = 0;
      phase ^= 1;
    }
  };

  int32_t src_a[kJ][2], src_b[kJ][2];
  uint32_t ci_a[kJ], ci_b[kJ];
  float4 v_a[kJ][2], v_b[kJ][2];
  load_idx(src_a, ci_a);
  gather(src_a, ci_a, v_a);
  if (total > 1) load_idx(src_b, ci_b);
  for (int s = 0; s < total; s += 2) {
    // stage s is in v_a; the rulebook entries of stage s+1 are in src_b
    if (s + 1 < total) gather(src_b, ci_b, v_b);
    if (s + 2 < total) load_idx(src_a, ci_a);
    finish(v_a);
    if (s + 1 < total) {
      if (s + 2 < total) gather(src_a, ci_a, v_a);
      if (s + 3 < total) load_idx(src_b, ci_b);
      finish(v_b);
    }
  }
}

// bf16x3 register path with 8-value units (dense GEMMs and the channel counts the cp.async path does not take): one
// LDG.256 per unit and one 16-byte shared-memory store per operand image, i.e. half the load / store instructions of the
// 4-value path above.  The pipeline traces (scripts/trace_conv.py) showed that path bound by the load/store unit's queue:
// every LDG / STS of the producers AND of the epilogue warps waited in it (a 4-store epilogue block took ~300 cycles).
// Needs a 32-byte aligned input and c_red % 8 == 0.  Eight lanes own a tile row (lane l: 16-byte piece l of the 128-byte
// hi / lo rows), a warp stores four full rows per instruction: conflict-free under the 128-byte swizzle.
struct Wide8 {
  float4 a, b;
};
__device__ __forceinline__ Wide8 ldg256(const float* ptr) {
  Wide8 v;
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w)
               : "l"(ptr));
  return v;
}
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - __uint_as_float(hi << 16), y - __uint_as_float(hi & 0xFFFF0000u));
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void produce_a_wide(const Params& p, const CtaWork& w, uint8_t* a_ring, uint64_t* a_full,
                                               uint64_t* a_empty, const int tid, const int lane) {
  static_assert(kProducerWarps == 16, "the row-slot mapping assumes 64 row slots of 8 lanes");
  constexpr int a_part = kTileM * 128;
  constexpr int a_bytes = 2 * a_part;
  const int total = w.total_a;
  if (total == 0) return;
  const int l = tid & 7;
  const int r_lo = tid >> 3;   // rows r_lo and r_lo + 64
  const uint32_t off = static_cast<uint32_t>(r_lo) * 128u + (static_cast<uint32_t>(l ^ (r_lo & 7)) << 4);
  const uint32_t c_red = static_cast<uint32_t>(p.c_red);
  struct Src {
    int32_t e[2];     // source rows of tile rows r_lo, r_lo + 64 (-1: nothing)
    uint32_t ci;      // first channel of the unit inside its tap
  };
  int cu_s = 0, cu_c = 0, cu_t = 0, cu_ti = w.tiles_in(p, 0);
  auto load_idx = [&](Src& x) {
    const int tile = (static_cast<int>(blockIdx.x) + cu_s * static_cast<int>(gridDim.x)) * p.tiles_per_super + cu_t;
    const int row0 = tile * kTileM + r_lo;
    const bool ok0 = row0 < p.num_out, ok1 = row0 + 64 < p.num_out;
    const int kk = cu_c * kBf16ChunkK + 8 * l;
    const int tap = p.cred_shift >= 0 ? (kk >> p.cred_shift) : kk / p.c_red;
    x.ci = static_cast<uint32_t>(kk - tap * p.c_red);
    const bool tap_ok = tap < p.taps;
    if (p.nbr) {
      const int32_t* n0 = p.nbr + static_cast<int64_t>(row0) * p.taps + tap;
      x.e[0] = (tap_ok && ok0) ? __ldg(n0) : -1;
      x.e[1] = (tap_ok && ok1) ? __ldg(n0 + 64 * p.taps) : -1;
    } else {
      x.e[0] = (tap_ok && ok0) ? row0 : -1;
      x.e[1] = (tap_ok && ok1) ? row0 + 64 : -1;
    }
    if (++cu_t == cu_ti) {
      cu_t = 0;
      if (++cu_c == p.chunks) {
        cu_c = 0;
        ++cu_s;
        cu_ti = w.tiles_in(p, cu_s);
      }
    }
  };
  auto gather = [&](const Src& x, Wide8 (&v)[2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      v[i].a = make_float4(0.f, 0.f, 0.f, 0.f);
      v[i].b = make_float4(0.f, 0.f, 0.f, 0.f);
      // 32-bit element offset (the host checks num_in * c_red < 2^32)
      if (x.e[i] >= 0) v[i] = ldg256(p.in + static_cast<size_t>(static_cast<uint32_t>(x.e[i]) * c_red + x.ci));
    }
  };
  int slot = 0;
  uint32_t phase = 0;
  [[maybe_unused]] int t_fin = 0;
  auto finish = [&](const Wide8 (&v)[2]) {
    if (tid == 0) EFGB_TRACE(t_fin, 0);
    mbar_wait(smem_u32(&a_empty[slot]), phase ^ 1);
    if (tid == 0) EFGB_TRACE(t_fin, 1);
    uint8_t* a_hi = a_ring + static_cast<size_t>(slot) * a_bytes;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint4 hi, lo;
      split2(v[i].a.x, v[i].a.y, hi.x, lo.x);
      split2(v[i].a.z, v[i].a.w, hi.y, lo.y);
      split2(v[i].b.x, v[i].b.y, hi.z, lo.z);
      split2(v[i].b.z, v[i].b.w, hi.w, lo.w);
      *reinterpret_cast<uint4*>(a_hi + off + static_cast<uint32_t>(i) * (64u * 128u)) = hi;
      *reinterpret_cast<uint4*>(a_hi + a_part + off + static_cast<uint32_t>(i) * (64u * 128u)) = lo;
    }
    if (tid == 0) EFGB_TRACE(t_fin, 2);
#if !EFGB_TC_FENCE_AT_MMA
    fence_proxy_async();
#endif
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&a_full[slot]));
    if (tid == 0) {
      EFGB_TRACE(t_fin, 3);
      EFGB_TRACE(t_fin, 9);
      EFGB_TRACE(t_fin, 10);
    }
    ++t_fin;
    if (++slot == p.sa) {
      slot = 0;
      phase ^= 1;
    }
  };
  Src src_a, src_b;
  Wide8 v_a[2], v_b[2];
  load_idx(src_a);
  gather(src_a, v_a);
  if (total > 1) load_idx(src_b);
  for (int s = 0; s < total; s += 2) {
    // stage s is in v_a; the rulebook entries of stage s+1 are in src_b
    if (s + 1 < total) gather(src_b, v_b);
    if (s + 2 < total) load_idx(src_a);
    finish(v_a);
    if (s + 1 < total) {
      if (s + 2 < total) gather(src_a, v_a);
      if (s + 3 < total) load_idx(src_b);
      finish(v_b);
    }
  }
}

// A producers over a PRE-SPLIT input (bf16x3, sparse convolutions).  A gathered input row is used by ~14 output rows, so
// splitting fp32 -> bf16 hi / lo inside the producers repeats the conversion 14 times and was what bound the kernel
// (issue slots: ~90 instructions per 16-byte piece of the A tile, profiles/r2_ncu_full_spconv_planes_before_mma_fix.txt).  Here the input has
// been split once into `planes[row] = [hi c_red x bf16 | lo c_red x bf16]` (same bytes per row as fp32) and a piece of
// the A tile is ONE cp.async (LDGSTS, zero-fill for a missing neighbour): no registers, no conversion, ~10 instructions
// per piece, kDepth stages of gathers in flight per thread.
//   per row and 64-value chunk the source bytes are, tap segment by tap segment, [hi run | lo run]; eight consecutive
//   lanes copy 128 consecutive source bytes of one row (minimum number of L1TEX wavefronts), a thread serves rows
//   r and r + 64 with two pieces each.
constexpr int kPlaneDepth = 3;   // stages of cp.async in flight, odd (A ring depth must be >= kPlaneDepth + 2)
static_assert(kPlaneDepth % 2 == 1, "the unrolled producer loop assumes an odd depth");

__device__ __forceinline__ void produce_a_planes(const Params& p, const CtaWork& w, uint8_t* a_ring, uint64_t* a_full,
                                                 uint64_t* a_empty, const int tid, const int lane) {
  static_assert(kProducerWarps == 16, "the row-slot mapping assumes 64 row slots of 8 lanes");
  const int a_part = kTileM * 128;
  const int a_bytes = 2 * a_part;
  const int total = w.total_a;
  if (total == 0) return;
  const int C = p.c_red;
  const int seg_vals = C < kBf16ChunkK ? C : kBf16ChunkK;     // values of one tap inside a chunk
  const int wp = seg_vals / 8;                                  // 16-byte pieces of a hi (or lo) run: 2, 4 or 8
  const int tpc = kBf16ChunkK / seg_vals;                       // taps per chunk: 4, 2 or 1
  const int j = tid & 7;
  const int slot_id = tid >> 3;                                 // 0..63
  const int r_lo = slot_id;                                     // rows r_lo and r_lo + 64
  // per-thread constants of its two pieces (k = 0, 1): tap segment, hi/lo, shared-memory offset, source byte offset
  int seg[2];
  uint32_t dst_off[2], src_off[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int g = k * 8 + j;
    seg[k] = g / (2 * wp);
    const int within = g - seg[k] * 2 * wp;
    const int is_lo = within >= wp ? 1 : 0;
    const int piece = within - is_lo * wp;
    const int q = seg[k] * wp + piece;                           // 16-byte piece of the 128-byte tile row
    dst_off[k] = static_cast<uint32_t>(is_lo * a_part + r_lo * 128 + ((q ^ (r_lo & 7)) << 4));
    src_off[k] = static_cast<uint32_t>(is_lo * C * 2 + piece * 16);
  }
  const bool two_taps = seg[1] != seg[0];
  const uint32_t row_bytes = static_cast<uint32_t>(C) * 4u;
  const uint32_t a_base = smem_u32(a_ring);

  // stage cursor (super-tile, chunk, tile) of the rulebook prefetch, which runs two stages ahead of the copies:
  // the entries are consumed a full stage time after their loads were issued, and never copied between registers
  // (a register move of a pending load stalls on its latency: that was 22 % of all stall samples).
  int cu_s = 0, cu_c = 0, cu_t = 0, cu_ti = w.tiles_in(p, 0);
  struct Idx {
    int32_t v[2][2];      // [piece k][row i]; v[1] is unused when both pieces belong to the same tap
    uint32_t ci_bytes;    // byte offset of the chunk inside a tap's hi run (c_red > 64)
  };
  const int num_out = static_cast<int>(p.num_out);
  auto load_idx = [&](Idx& x) {
    const int tile = (static_cast<int>(blockIdx.x) + cu_s * static_cast<int>(gridDim.x)) * p.tiles_per_super + cu_t;
    const int row0 = tile * kTileM + r_lo;
    int tap0;
    if (C <= kBf16ChunkK) {
      tap0 = cu_c * tpc;
      x.ci_bytes = 0;
    } else {
      const int kk = cu_c * kBf16ChunkK;
      tap0 = p.cred_shift >= 0 ? (kk >> p.cred_shift) : kk / C;
      x.ci_bytes = static_cast<uint32_t>(kk - tap0 * C) * 2u;
    }
    const uint32_t e0 = static_cast<uint32_t>(row0) * static_cast<uint32_t>(p.taps);
    const uint32_t e1 = e0 + 64u * static_cast<uint32_t>(p.taps);
    const bool ok0 = row0 < num_out, ok1 = row0 + 64 < num_out;
    {
      const int tap = tap0 + seg[0];
      const bool tap_ok = tap < p.taps;
      if constexpr ((EFGB_TC_ABLATE & 4) != 0) {
        x.v[0][0] = (tap_ok && ok0) ? row0 : -1;
        x.v[0][1] = (tap_ok && ok1) ? row0 + 64 : -1;
      } else {
      x.v[0][0] = (tap_ok && ok0) ? __ldg(p.nbr + (e0 + tap)) : -1;
      x.v[0][1] = (tap_ok && ok1) ? __ldg(p.nbr + (e1 + tap)) : -1;
      }
    }
    if (two_taps && (EFGB_TC_ABLATE & 4) == 0) {
      const int tap = tap0 + seg[1];
      const bool tap_ok = tap < p.taps;
      x.v[1][0] = (tap_ok && ok0) ? __ldg(p.nbr + (e0 + tap)) : -1;
      x.v[1][1] = (tap_ok && ok1) ? __ldg(p.nbr + (e1 + tap)) : -1;
    }
    if constexpr ((EFGB_TC_ABLATE & 8) != 0) x.v[0][0] = x.v[0][1] = x.v[1][0] = x.v[1][1] = -1;
    if (++cu_t == cu_ti) {
      cu_t = 0;
      if (++cu_c == p.chunks) {
        cu_c = 0;
        ++cu_s;
        cu_ti = w.tiles_in(p, cu_s);
      }
    }
  };
  int islot = 0;
  uint32_t iphase = 0;
  [[maybe_unused]] int t_issue = 0, t_done = 0;
  auto issue = [&](const Idx& x) {   // copies of the stage whose rulebook entries are in x
    if (tid == 0) EFGB_TRACE(t_issue, 0);
    mbar_wait(smem_u32(&a_empty[islot]), iphase ^ 1);
    if (tid == 0) EFGB_TRACE(t_issue, 1);
    const uint32_t dst = a_base + static_cast<uint32_t>(islot) * static_cast<uint32_t>(a_bytes);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int32_t e = (k == 1 && two_taps) ? x.v[1][i] : x.v[0][i];
        const uint32_t row = static_cast<uint32_t>(e < 0 ? 0 : e);   // a valid address even when nothing is read
        const uint8_t* g = p.planes + (static_cast<size_t>(row) * row_bytes + (src_off[k] + x.ci_bytes));
        if constexpr ((EFGB_TC_ABLATE & 2) == 0) cp_async_16(dst + dst_off[k] + static_cast<uint32_t>(i * 64 * 128), g, e < 0 ? 0u : 16u);
      }
    }
    if (tid == 0) EFGB_TRACE(t_issue, 2);
    ++t_issue;
    if (++islot == p.sa) {
      islot = 0;
      iphase ^= 1;
    }
  };
  int cslot = 0;
  auto complete = [&]() {
    cp_async_commit();                   // possibly empty: one group per stage keeps the wait count constant
    if (tid == 0) EFGB_TRACE(t_done, 9);
    cp_async_wait<kPlaneDepth>();        // the copies of the oldest stage in flight have landed
    if (tid == 0) EFGB_TRACE(t_done, 3);
#if !EFGB_TC_FENCE_AT_MMA
    fence_proxy_async();                 // -> visible to the tensor core (async proxy)
#endif
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&a_full[cslot]));
    if (tid == 0) EFGB_TRACE(t_done, 10);
    ++t_done;
    if (++cslot == p.sa) cslot = 0;
  };

  // Stage n is issued from xa when n is even, from xb when n is odd; its entries were loaded two issues earlier.
  Idx xa, xb;
  load_idx(xa);
  if (total > 1) load_idx(xb);
  int n = 0;         // next stage to issue
  int loaded = 2;    // stages whose rulebook entries have been requested
  // prologue: kPlaneDepth stages in flight before the first completion (empty groups past the end)
#pragma unroll 1
  for (int d = 0; d < kPlaneDepth; ++d) {
    if (n < total) {
      if ((n & 1) == 0) {
        issue(xa);
        if (loaded < total) load_idx(xa);
      } else {
        issue(xb);
        if (loaded < total) load_idx(xb);
      }
      ++loaded;
      ++n;
    }
    cp_async_commit();
  }
  // steady state, unrolled by two so that xa / xb stay in fixed registers; kPlaneDepth is odd: stage n = s + 3
#pragma unroll 1
  for (int s = 0; s < total; s += 2) {
    if (n < total) {           // n is odd here (kPlaneDepth odd, s even)
      issue(xb);
      if (loaded < total) load_idx(xb);
      ++loaded;
      ++n;
    }
    complete();
    if (s + 1 < total) {
      if (n < total) {
        issue(xa);
        if (loaded < total) load_idx(xa);
        ++loaded;
        ++n;
      }
      complete();
    }
  }
}

template <int kMode>
__global__ void __launch_bounds__(kThreads, 1) spconv_tc_kernel(const Params p) {
  using M = ModeTraits<kMode>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the swizzle atoms
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kParts = M::kParts;
  const int n_cta = p.n_cta;  // output columns owned by this CTA (N split over grid.y)
  const int n0 = static_cast<int>(blockIdx.y) * n_cta;
  const int T = p.tiles_per_super;
  const int a_part = kTileM * 128;
  const int b_part = n_cta * 128;
  const int a_bytes = kParts * a_part;
  const int b_bytes = kParts * b_part;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + static_cast<size_t>(p.sa) * a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + static_cast<size_t>(p.sb) * b_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.sa;
  uint64_t* b_full = a_empty + p.sa;
  uint64_t* b_empty = b_full + p.sb;
  uint64_t* tmem_full = b_empty + p.sb;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>(tmem_slot + 4);   // 4 x kEpiStageBytes, 16-byte aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) EFGB_TRACE(0, 6);

  // TMEM: two sets of T accumulators of acc_cols fp32 columns each, power of two >= 32
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(2 * T * p.acc_cols)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.sa; ++s) {
      mbar_init(smem_u32(&a_full[s]), kProducerWarps);
      mbar_init(smem_u32(&a_empty[s]), 1);
    }
    for (int s = 0; s < p.sb; ++s) {
      mbar_init(smem_u32(&b_full[s]), 1);
      mbar_init(smem_u32(&b_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full[a]), 1);
      mbar_init(smem_u32(&tmem_empty[a]), 4);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const CtaWork w(p);
  if (threadIdx.x == 0) EFGB_TRACE(1, 6);

  if (warp < kProducerWarps) {
    // ================= A producers =================
    if constexpr (kMode == kBf16x3) {
      if (p.planes)
        produce_a_planes(p, w, a_ring, a_full, a_empty, static_cast<int>(threadIdx.x), lane);
      else if (p.wide)
        produce_a_wide(p, w, a_ring, a_full, a_empty, static_cast<int>(threadIdx.x), lane);
      else
        produce_a<kMode>(p, w, a_ring, a_full, a_empty, static_cast<int>(threadIdx.x), lane);
    } else {
      produce_a<kMode>(p, w, a_ring, a_full, a_empty, static_cast<int>(threadIdx.x), lane);
    }
  } else if (warp == kLoaderWarp) {
    // ================= B loader: weight chunks through their own ring (bulk copies) =================
    // A chunk serves all T tiles of the super-tile, and the ring runs ahead of the A stream, so neither the
    // L2 latency nor the L2 bandwidth of the weights sits on the A-stage turnaround.  The hi and lo slabs of the
    // CTA's columns land next to each other: [B_hi | B_lo] is one 2 n_cta-row K-major tile for concatenated issue.
    if (lane == 0) {
      const uint32_t part_bytes = static_cast<uint32_t>(b_part);
      int stage = 0;
      uint32_t phase = 0;
      // Dense GEMM (identity rulebook): the rows of a super-tile are one contiguous block of the input; pull the NEXT
      // one into L2 a whole super-tile ahead of the producers' loads.  The activations of a token-wise linear come
      // from HBM (ncu: 55 % L2 hit rate, 46 % of the stall samples on the first use of a loaded value), and the
      // register pipeline keeps only one stage of loads in flight.
      auto prefetch_super = [&](int i) {
        if (p.nbr != nullptr || p.in == nullptr || blockIdx.y != 0 || i >= w.n_super) return;
        const int64_t st1 = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(i) * gridDim.x;
        const int64_t r0 = st1 * T * kTileM;
        const int64_t r1 = r0 + T * kTileM < p.num_out ? r0 + T * kTileM : p.num_out;
        const int64_t bytes = (r1 - r0) * p.c_red * 4;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.in) + r0 * p.c_red * 4;
        for (int64_t o = 0; o < bytes; o += 65536) {
          const uint32_t n = static_cast<uint32_t>(bytes - o < 65536 ? bytes - o : 65536);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + o), "r"(n) : "memory");
        }
      };
      for (int i = 0; i < w.n_super; ++i) {
        if (i == 0) prefetch_super(0);
        prefetch_super(i + 1);
        for (int c = 0; c < p.chunks; ++c) {
          mbar_wait(smem_u32(&b_empty[stage]), phase ^ 1);
          const uint32_t bar = smem_u32(&b_full[stage]);
          mbar_arrive_expect_tx(bar, static_cast<uint32_t>(b_bytes));
          const uint32_t dst = smem_u32(b_ring + static_cast<size_t>(stage) * b_bytes);
          for (int part = 0; part < kParts; ++part) {
            const uint8_t* src = p.packed + (static_cast<size_t>(c * kParts + part) * p.n_out + n0) * 128u;
            for (uint32_t o = 0; o < part_bytes; o += 16384u) {
              const uint32_t n = part_bytes - o < 16384u ? part_bytes - o : 16384u;
              bulk_g2s(dst + part * part_bytes + o, src + o, n, bar);
            }
          }
          if (++stage == p.sb) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    // The whole warp runs the loop convergently (waits included) and one elected lane issues: descriptors, TMEM
    // addresses and barrier addresses then live in uniform registers.  (Round 1 ran the loop inside `if (lane == 0)`:
    // in divergent code every tcgen05.mma operand went through R2UR moves, ~100 cycles per MMA — with the eight small-N
    // MMAs of a C <= 64 stage that serial issue time, not shared memory, was what bound the kernel.)
    {
      const uint32_t idesc_n = make_idesc<kMode>(kTileM, n_cta);
      const uint32_t idesc_2n = make_idesc<kMode>(kTileM, 2 * n_cta);
      const uint32_t a_ring_u = smem_u32(a_ring), b_ring_u = smem_u32(b_ring);
      const uint32_t a_full_u = smem_u32(a_full), a_empty_u = smem_u32(a_empty);
      const uint32_t b_full_u = smem_u32(b_full), b_empty_u = smem_u32(b_empty);
      int sa_ = 0, sb_ = 0;
      uint32_t pa = 0, pb = 0;
      [[maybe_unused]] int t_mma = 0;
      for (int i = 0; i < w.n_super; ++i) {
        const int acc = i & 1;
        const uint32_t acc_phase = static_cast<uint32_t>((i >> 1) & 1);
        const int ti = w.tiles_in(p, i);
        mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
        tc_fence_after();
        for (int c = 0; c < p.chunks; ++c) {
          mbar_wait(b_full_u + sb_ * 8, pb);
          const uint32_t b_hi = b_ring_u + static_cast<uint32_t>(sb_ * b_bytes);
          const uint64_t db_hi = make_desc_sw128(b_hi);
          const uint64_t db_lo = make_desc_sw128(b_hi + b_part);
          for (int t = 0; t < ti; ++t) {
            mbar_wait(a_full_u + sa_ * 8, pa);
            if (lane == 0) EFGB_TRACE(t_mma, 4);
#if EFGB_TC_FENCE_AT_MMA
            fence_proxy_async();   // the producers' generic-proxy writes (acquired above) -> visible to the async proxy
#endif
            if (lane == 0 && t_mma >= 10) EFGB_TRACE(t_mma, 6);
            tc_fence_after();
            const uint32_t a_hi = a_ring_u + static_cast<uint32_t>(sa_ * a_bytes);
            const uint64_t da_hi = make_desc_sw128(a_hi);
            const uint64_t da_lo = make_desc_sw128(a_hi + a_part);
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>((acc * T + t) * p.acc_cols);
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint64_t adv = static_cast<uint64_t>(j * 2);  // 32 bytes of K = 2 x 16 B
                const uint32_t accum = (c | j) ? 1u : 0u;
                if constexpr (kParts == 2) {
                  if (p.concat) {
                    tc_mma<kMode>(tmem_d, da_hi + adv, db_hi + adv, idesc_2n, accum);   // [a_hi b_hi | a_hi b_lo]
                    tc_mma<kMode>(tmem_d, da_lo + adv, db_hi + adv, idesc_n, 1u);       // first half += a_lo b_hi
                  } else {
                    tc_mma<kMode>(tmem_d, da_lo + adv, db_hi + adv, idesc_n, accum);
                    tc_mma<kMode>(tmem_d, da_hi + adv, db_lo + adv, idesc_n, 1u);
                    tc_mma<kMode>(tmem_d, da_hi + adv, db_hi + adv, idesc_n, 1u);
                  }
                } else {
                  tc_mma<kMode>(tmem_d, da_hi + adv, db_hi + adv, idesc_n, accum);
                }
              }
              if (t_mma >= 10) EFGB_TRACE(t_mma, 7);
              tc_commit(a_empty_u + sa_ * 8);  // A stage reusable once these MMAs have read it
            }
            __syncwarp();
            if (lane == 0) EFGB_TRACE(t_mma, 5);
            ++t_mma;
            if (++sa_ == p.sa) {
              sa_ = 0;
              pa ^= 1;
            }
          }
          if (elect_one()) tc_commit(b_empty_u + sb_ * 8);    // weight chunk consumed by all tiles of the super-tile
          __syncwarp();
          if (++sb_ == p.sb) {
            sb_ = 0;
            pb ^= 1;
          }
        }
        if (elect_one()) tc_commit(smem_u32(&tmem_full[acc]));    // accumulators of the super-tile complete
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue (4 warps; warp w may only touch TMEM lanes 32*(w%4)..+31) =================
    const int quarter = warp & 3;
    uint8_t* stage_w = epi_stage + quarter * kEpiStageBytes;
    for (int i = 0; i < w.n_super; ++i) {
      const int acc = i & 1;
      const int ti = w.tiles_in(p, i);
      const int st = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      mbar_wait(smem_u32(&tmem_full[acc]), static_cast<uint32_t>((i >> 1) & 1));
      tc_fence_after();
      if (quarter == 0 && lane == 0) EFGB_TRACE(2 + 2 * i, 6);
      for (int t = 0; t < ti; ++t) {
        const int64_t row0 = (static_cast<int64_t>(st) * T + t) * kTileM + quarter * 32;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>((acc * T + t) * p.acc_cols);
        for (int c0 = 0; c0 < n_cta; c0 += 16) {
          [[maybe_unused]] const bool tr = quarter == 0 && lane == 0 && i == 1 && t == 0;
          if (tr) EFGB_TRACE(1000 + (c0 >> 4), 0);
          uint32_t r[16];
          tc_ld16(taddr + c0, r);
          if (p.concat) {
            uint32_t r2[16];
            tc_ld16(taddr + n_cta + c0, r2);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
          } else {
            tc_wait_ld();
          }
          if (tr) EFGB_TRACE(1000 + (c0 >> 4), 1);
          // A thread holds 16 columns of ITS row: stored directly, a warp instruction would touch 32 rows (32 L1TEX
          // wavefronts per 512 bytes — measured: 16.6 k cycles per 128 x 256 tile, more than the tile's main loop).
          // The 32 x 16 block goes through a 2 KB staging buffer (XOR-swizzled 16-byte pieces, conflict-free both
          // ways) and leaves as 8 rows x 64 contiguous bytes per instruction.
          if (!p.epi_staged) {
            // no room for the staging buffers (256-column CTAs with two operand images): row-per-thread stores
            const int64_t row = row0 + lane;
            if (row < p.num_out) {
              float* dst = p.out + row * p.n_out + n0 + c0;
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                float4 o;
                o.x = __uint_as_float(r[j + 0]) + (p.bias ? p.bias[n0 + c0 + j + 0] : 0.f);
                o.y = __uint_as_float(r[j + 1]) + (p.bias ? p.bias[n0 + c0 + j + 1] : 0.f);
                o.z = __uint_as_float(r[j + 2]) + (p.bias ? p.bias[n0 + c0 + j + 2] : 0.f);
                o.w = __uint_as_float(r[j + 3]) + (p.bias ? p.bias[n0 + c0 + j + 3] : 0.f);
                if (p.relu) {
                  o.x = fmaxf(o.x, 0.f);
                  o.y = fmaxf(o.y, 0.f);
                  o.z = fmaxf(o.z, 0.f);
                  o.w = fmaxf(o.w, 0.f);
                }
                *reinterpret_cast<float4*>(dst + j) = o;
              }
            }
            continue;
          }
          __syncwarp();   // the previous block's reads of the staging buffer are done
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 o;
            o.x = __uint_as_float(r[j + 0]) + (p.bias ? p.bias[n0 + c0 + j + 0] : 0.f);
            o.y = __uint_as_float(r[j + 1]) + (p.bias ? p.bias[n0 + c0 + j + 1] : 0.f);
            o.z = __uint_as_float(r[j + 2]) + (p.bias ? p.bias[n0 + c0 + j + 2] : 0.f);
            o.w = __uint_as_float(r[j + 3]) + (p.bias ? p.bias[n0 + c0 + j + 3] : 0.f);
            if (p.relu) {
              o.x = fmaxf(o.x, 0.f);
              o.y = fmaxf(o.y, 0.f);
              o.z = fmaxf(o.z, 0.f);
              o.w = fmaxf(o.w, 0.f);
            }
            *reinterpret_cast<float4*>(stage_w + lane * 64 + ((((j >> 2) ^ (lane >> 1)) & 3) << 4)) = o;
          }
          __syncwarp();
          if (tr) EFGB_TRACE(1000 + (c0 >> 4), 2);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int rl = (lane >> 2) + 8 * k;
            const float4 o = *reinterpret_cast<const float4*>(stage_w + rl * 64 + (((lane ^ (rl >> 1)) & 3) << 4));
            const int64_t row = row0 + rl;
            if (row < p.num_out) *reinterpret_cast<float4*>(p.out + row * p.n_out + n0 + c0 + (lane & 3) * 4) = o;
          }
          if (tr) EFGB_TRACE(1000 + (c0 >> 4), 3);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (quarter == 0 && lane == 0) EFGB_TRACE(3 + 2 * i, 6);
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) EFGB_TRACE(0, 7);
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// Pack weights (reference parameter layout [c_out, taps, c_in]) into the swizzled K-major chunk image.
//   mode 0: forward        B[n = co][(tap, c = ci)] = w[co][tap][ci]            (N = c_out, c_red = c_in)
//   mode 1: dgrad          B[n = ci][(tap, c = co)] = w[co][tap][ci]            (N = c_in,  c_red = c_out)
//   mode 2: dgrad, subm    B[n = ci][(tap, c = co)] = w[co][taps-1-tap][ci]     (tap mirrored, see spconv/pytorch.py)
template <bool kSplit>
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ w, int c_out, int taps, int c_in, int mode, int n_out, int c_red, int chunks,
                    float* __restrict__ packed) {
  constexpr int kParts = kSplit ? 2 : 1;
  const int64_t total = static_cast<int64_t>(chunks) * n_out * kChunkK;
  int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int kk = static_cast<int>(t % kChunkK);
  const int n = static_cast<int>((t / kChunkK) % n_out);
  const int chunk = static_cast<int>(t / (static_cast<int64_t>(kChunkK) * n_out));
  const int kidx = chunk * kChunkK + kk;
  const int tap = kidx / c_red;
  const int c = kidx - tap * c_red;
  float v = 0.f;
  if (tap < taps) {
    if (mode == 0) {
      if (n < c_out) v = w[(static_cast<int64_t>(n) * taps + tap) * c_in + c];
    } else {
      const int st = mode == 2 ? taps - 1 - tap : tap;
      if (n < c_in) v = w[(static_cast<int64_t>(c) * taps + st) * c_in + n];
    }
  }
  const int qphys = (kk >> 2) ^ (n & 7);
  const int64_t off = static_cast<int64_t>(n) * kChunkK + qphys * 4 + (kk & 3);
  float* base = packed + static_cast<int64_t>(chunk) * kParts * n_out * kChunkK;
  if (kSplit) {
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    base[off] = hi;
    base[static_cast<int64_t>(n_out) * kChunkK + off] = v - hi;
  } else {
    base[off] = v;
  }
}


// bf16x3 weight image: [chunk of 64 K-values][hi | lo][n][64 bf16], K-major, 128-byte swizzle (16-byte pieces of 8 values).
__device__ __forceinline__ void pack_bf16_element(const float* __restrict__ w, int c_out, int taps, int c_in, int mode, int n_out,
                                                  int c_red, int64_t t, __nv_bfloat16* __restrict__ packed) {
  const int kk = static_cast<int>(t % kBf16ChunkK);
  const int n = static_cast<int>((t / kBf16ChunkK) % n_out);
  const int chunk = static_cast<int>(t / (static_cast<int64_t>(kBf16ChunkK) * n_out));
  const int kidx = chunk * kBf16ChunkK + kk;
  const int tap = kidx / c_red;
  const int c = kidx - tap * c_red;
  float v = 0.f;
  if (tap < taps) {
    if (mode == 0) {
      if (n < c_out) v = w[(static_cast<int64_t>(n) * taps + tap) * c_in + c];
    } else {
      const int st = mode == 2 ? taps - 1 - tap : tap;
      if (n < c_in) v = w[(static_cast<int64_t>(c) * taps + st) * c_in + n];
    }
  }
  const int qphys = (kk >> 3) ^ (n & 7);
  const int64_t off = static_cast<int64_t>(n) * kBf16ChunkK + qphys * 8 + (kk & 7);
  __nv_bfloat16* base = packed + static_cast<int64_t>(chunk) * 2 * n_out * kBf16ChunkK;
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  base[off] = hi;
  base[static_cast<int64_t>(n_out) * kBf16ChunkK + off] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__global__ void __launch_bounds__(256)
pack_weights_bf16_kernel(const float* __restrict__ w, int c_out, int taps, int c_in, int mode, int n_out, int c_red, int chunks,
                         __nv_bfloat16* __restrict__ packed) {
  const int64_t total = static_cast<int64_t>(chunks) * n_out * kBf16ChunkK;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= total) return;
  pack_bf16_element(w, c_out, taps, c_in, mode, n_out, c_red, t, packed);
}

// Every bf16x3 weight image of a training step in ONE launch.  jobs[j] = {weights, image, c_out, taps, c_in, mode,
// first block, blocks} as eight int64; a block finds its job by bisection over the first-block column.
__global__ void __launch_bounds__(256) pack_weights_bf16_batched_kernel(const int64_t* __restrict__ jobs, int n_jobs) {
  int lo = 0, hi = n_jobs - 1;
  const int64_t b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(jobs + mid * 8 + 6) <= b) lo = mid; else hi = mid - 1;
  }
  const int64_t* j = jobs + lo * 8;
  const float* w = reinterpret_cast<const float*>(__ldg(j + 0));
  __nv_bfloat16* packed = reinterpret_cast<__nv_bfloat16*>(__ldg(j + 1));
  const int c_out = static_cast<int>(__ldg(j + 2)), taps = static_cast<int>(__ldg(j + 3)), c_in = static_cast<int>(__ldg(j + 4));
  const int mode = static_cast<int>(__ldg(j + 5));
  const int n_out = mode == 0 ? c_out : c_in;
  const int c_red = mode == 0 ? c_in : c_out;
  const int64_t total = static_cast<int64_t>((taps * c_red + kBf16ChunkK - 1) / kBf16ChunkK) * n_out * kBf16ChunkK;
  const int64_t t = (b - __ldg(j + 6)) * blockDim.x + threadIdx.x;
  if (t >= total) return;
  pack_bf16_element(w, c_out, taps, c_in, mode, n_out, c_red, t, packed);
}

// ================================================================================================
// wgrad on the tensor cores:  dW[(tap, ci), co] = sum over output rows  A[row, (tap, ci)] * G[row, co]
//   A[row, (tap, ci)] = in[nbr[row, tap], ci]   (gathered, zeros where nbr = -1),  G = grad_out.
// The reduction runs over ROWS, so both operands are "MN-major" for the tensor core: a stage holds
// 32 rows; the 128-wide M slice (a group of 128 consecutive flattened (tap, ci) indices) and the N = Cout
// columns are split into 32-float column blocks, each block stored as [32 rows][128 B] with the
// 32-byte-base 128B swizzle that MN-major tf32 requires (atoms of 4 rows x 128 B, LBO = block stride,
// SBO = 512 B between 4-row groups; one K = 8 MMA spans two atoms).
// A work item = (M group, row chunk); persistent CTAs stride over items; the fp32 accumulator
// [128 x Cout] lives in TMEM (double buffered) and is flushed with coalesced red.global.add into the
// reference parameter layout dW[co][tap][ci].
// ================================================================================================
constexpr int kWgRows = 32;  // rows (K) per stage

struct WgradParams {
  const float* in;      // [num_in, c_in]
  const float* gout;    // [num_out, c_out]
  const int32_t* nbr;   // [num_out, taps]
  float* dw;            // [c_out, taps, c_in], zero-initialised
  int64_t num_out;
  int c_in, c_out, taps, n_pad;   // c_out = full output width (row stride of gout / dW); n_pad = columns per CTA slab
  int groups;           // ceil(taps * c_in / 128)
  int row_chunks;
  int64_t rows_per_chunk;  // multiple of kWgRows
  int num_items;
};

__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;   // SBO: 4 K-rows x 128 B per swizzle atom
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(1) << 61;          // SWIZZLE_128B_BASE32B
  return d;
}

// Byte offset of 16-byte piece `pc` (0..7 within a 32-float column block) of K-row r inside a column block.
// MN-major tf32 operands must use the 128B swizzle with a 32-byte base (cutlass sm100_common.inl:89-94,
// Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> over 4 rows x 128 B): the 32-byte chunk index is XORed with r & 3.
__device__ __forceinline__ uint32_t mn_piece_offset(int r, int pc) {
  return static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>(((pc & 7) >> 1) ^ (r & 3)) << 5) +
         (static_cast<uint32_t>(pc & 1) << 4);
}

__device__ __forceinline__ void split_store(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, const float4& v, bool split) {
  if (split) {
    float4 hi, lo;
    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    lo.x = v.x - hi.x;
    lo.y = v.y - hi.y;
    lo.z = v.z - hi.z;
    lo.w = v.w - hi.w;
    *reinterpret_cast<float4*>(hi_base + off) = hi;
    *reinterpret_cast<float4*>(lo_base + off) = lo;
  } else {
    *reinterpret_cast<float4*>(hi_base + off) = v;
  }
}

// wgrad producers: kGroups warp groups, group g owns every kGroups-th 32-row stage of this CTA's stage
// stream (every work item has exactly rows_per_chunk/32 stage slots; slots past the end are zero tiles).
// All per-piece index arithmetic is hoisted: a thread's A pieces share one 16-byte column (tap and channel change
// only with the work item) and walk the rows with a constant stride, likewise its grad_out pieces when the slab
// width is a power of two — the first version spent ~80 instructions per 16-byte piece on divisions and address
// math and was issue-bound.
template <bool kSplit, int kGroups>
__device__ __forceinline__ void wgrad_produce(const WgradParams& p, const int stages, uint8_t* smem, const int stage_bytes,
                                              const int a_part, const int g_part, uint64_t* full_bar, uint64_t* empty_bar,
                                              const int warp, const int lane) {
  constexpr int kParts = kSplit ? 2 : 1;
  constexpr int kWarpsPerGroup = kProducerWarps / kGroups;
  constexpr int kGroupThreads = kWarpsPerGroup * 32;
  constexpr int kPiecesA = (kWgRows * 32) / kGroupThreads;  // 16-byte pieces of the A tile per thread
  constexpr int kRowStepA = kGroupThreads / 32;             // rows between consecutive A pieces of a thread
  constexpr int kBatch = 4;                                  // grad_out pieces per thread per batch (8 spilled: 80-register cap)
  const int gidx = warp / kWarpsPerGroup;
  const int tg = (warp % kWarpsPerGroup) * 32 + lane;
  const int pg = p.n_pad / 4;  // 16-byte pieces per grad_out row slab (padded)
  int pg_shift = -1;
  for (int sft = 3; sft <= 6; ++sft)
    if ((1 << sft) == pg) pg_shift = sft;
  const bool g_fast = pg_shift >= 0 && pg <= kGroupThreads;  // then a thread's pieces all sit in one 16-byte column
  const int co0 = static_cast<int>(blockIdx.y) * p.n_pad;  // first output channel of this CTA's slab
  const int flat_m = p.taps * p.c_in;
  const int spi = static_cast<int>(p.rows_per_chunk / kWgRows);  // stage slots per item
  const int my_items = (p.num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int64_t total = static_cast<int64_t>(my_items) * spi;
  const int g_total = kWgRows * pg;

  // A pieces of this thread: column piece pcA (fixed), rows rA0 + i * kRowStepA
  const int pcA = tg & 31, rA0 = tg >> 5;
  // rows advance by kRowStepA (a multiple of 4) between a thread's pieces, so the swizzle term (r & 3) is constant:
  // piece i sits kRowStepA * 128 bytes after piece i - 1 (no per-piece offset array: registers are the budget here)
  static_assert(kRowStepA % 4 == 0, "constant swizzle phase per thread");
  const uint32_t offA0 = static_cast<uint32_t>(pcA >> 3) * (kWgRows * 128) + mn_piece_offset(rA0, pcA);
  // grad_out pieces (fast path): column piece pcG (fixed), rows rG0 + j * rows_per_pass
  const int pcG = g_fast ? (tg & (pg - 1)) : 0;
  const int rG0 = g_fast ? (tg >> pg_shift) : 0;
  const int rStepG = g_fast ? (kGroupThreads >> pg_shift) : 1;
  const bool colG_ok = co0 + pcG * 4 < p.c_out;
  const uint32_t offG_col = static_cast<uint32_t>(pcG >> 3) * (kWgRows * 128);

  int cur_item = -1, tap = 0, ci = 0, g = 0;
  bool a_col_ok = false;
  int64_t row_begin = 0, row_end = 0;
  // stage cursor without divisions: gs = it * spi + sl
  int it = 0, sl = gidx;
  while (sl >= spi && it < my_items) {
    sl -= spi;
    ++it;
  }
  for (int64_t gs = gidx; gs < total; gs += kGroups) {
    if (it != cur_item) {
      cur_item = it;
      const int item = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      g = item % p.groups;
      row_begin = static_cast<int64_t>(item / p.groups) * p.rows_per_chunk;
      row_end = row_begin + p.rows_per_chunk;
      if (row_end > p.num_out) row_end = p.num_out;
      const int flat = g * 128 + pcA * 4;
      a_col_ok = flat < flat_m;
      tap = flat / p.c_in;
      ci = flat - tap * p.c_in;
    }
    const int64_t rb = row_begin + static_cast<int64_t>(sl) * kWgRows;
    const int stage = static_cast<int>(gs % stages);
    const uint32_t phase = static_cast<uint32_t>((gs / stages) & 1);

    // ---- A: 32 rows x 32 pieces
    int32_t src[kPiecesA];
#pragma unroll
    for (int i = 0; i < kPiecesA; ++i) {
      const int64_t row = rb + rA0 + i * kRowStepA;
      src[i] = -1;
      if (a_col_ok && row < row_end) src[i] = p.nbr ? __ldg(p.nbr + row * p.taps + tap) : static_cast<int32_t>(row);
    }
    float4 va[kPiecesA];
#pragma unroll
    for (int i = 0; i < kPiecesA; ++i) {
      va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src[i] >= 0) va[i] = __ldg(reinterpret_cast<const float4*>(p.in + static_cast<int64_t>(src[i]) * p.c_in + ci));
    }
    // grad_out rows: the first batch of loads is issued together with the A gathers (one memory latency for the
    // whole stage instead of one per batch); wider slabs need further batches
    float4 vg[kBatch];
    auto load_g = [&](int e0) {
      if (g_fast) {
        const int r_first = rG0 + (e0 >> pg_shift);
#pragma unroll
        for (int i = 0; i < kBatch; ++i) {
          const int r = r_first + i * rStepG;
          const int64_t row = rb + r;
          vg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < kWgRows && row < row_end && colG_ok)
            vg[i] = __ldg(reinterpret_cast<const float4*>(p.gout + row * p.c_out + co0 + pcG * 4));
        }
      } else {
#pragma unroll
        for (int i = 0; i < kBatch; ++i) {
          const int e = e0 + i * kGroupThreads + tg;
          vg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e < g_total) {
            const int r = e / pg, pc = e - r * pg;
            const int64_t row = rb + r;
            if (row < row_end && co0 + pc * 4 < p.c_out)
              vg[i] = __ldg(reinterpret_cast<const float4*>(p.gout + row * p.c_out + co0 + pc * 4));
          }
        }
      }
    };
    load_g(0);
    mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
    uint8_t* a_hi = smem + static_cast<size_t>(stage) * stage_bytes;
    uint8_t* a_lo = a_hi + a_part;
    uint8_t* g_hi = a_hi + kParts * a_part;
    uint8_t* g_lo = g_hi + g_part;
#pragma unroll
    for (int i = 0; i < kPiecesA; ++i) split_store(a_hi, a_lo, offA0 + static_cast<uint32_t>(i * kRowStepA * 128), va[i], kSplit);
    for (int e0 = 0; e0 < g_total; e0 += kGroupThreads * kBatch) {
      if (e0 > 0) load_g(e0);
      if (g_fast) {
        const int r_first = rG0 + (e0 >> pg_shift);
#pragma unroll
        for (int i = 0; i < kBatch; ++i) {
          const int r = r_first + i * rStepG;
          if (r < kWgRows) split_store(g_hi, g_lo, offG_col + mn_piece_offset(r, pcG), vg[i], kSplit);
        }
      } else {
#pragma unroll
        for (int i = 0; i < kBatch; ++i) {
          const int e = e0 + i * kGroupThreads + tg;
          if (e < g_total) {
            const int r = e / pg, pc = e - r * pg;
            split_store(g_hi, g_lo, static_cast<uint32_t>(pc >> 3) * (kWgRows * 128) + mn_piece_offset(r, pc), vg[i], kSplit);
          }
        }
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&full_bar[stage]));
    // advance the (item, slot) cursor by kGroups stages
    sl += kGroups;
    while (sl >= spi) {
      sl -= spi;
      ++it;
    }
  }
}

template <bool kSplit>
__global__ void __launch_bounds__(kThreads, 1) spconv_wgrad_tc_kernel(const WgradParams p, const int stages) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kParts = kSplit ? 2 : 1;
  const int a_part = 4 * kWgRows * 128;                 // 4 column blocks of the M slice
  const int g_part = (p.n_pad / 32) * kWgRows * 128;    // n_pad/32 column blocks of grad_out
  const int stage_bytes = kParts * (a_part + g_part);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(stages) * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + stages;
  uint64_t* tmem_full = bars + 2 * stages;
  uint64_t* tmem_empty = bars + 2 * stages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pgroups = stages >= 4 ? 4 : 2;  // producer warp groups (never more than stages)
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(2 * p.n_pad)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), kProducerWarps / pgroups);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full[a]), 1);
      mbar_init(smem_u32(&tmem_empty[a]), 4);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int flat_m = p.taps * p.c_in;

  if (warp < kProducerWarps) {
    // ================= producers: gather A rows and stream grad_out rows =================
    if (pgroups == 4)
      wgrad_produce<kSplit, 4>(p, stages, smem, stage_bytes, a_part, g_part, full_bar, empty_bar, warp, lane);
    else
      wgrad_produce<kSplit, 2>(p, stages, smem, stage_bytes, a_part, g_part, full_bar, empty_bar, warp, lane);
  } else if (warp == kMmaWarp) {
    // the whole warp runs the loop convergently and one elected lane issues (see spconv_tc_kernel's MMA issuer:
    // uniform-register operands instead of ~100 cycles of R2UR moves per tcgen05.mma)
    {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phases = 0;   // bit a = phase of accumulator a (no dynamically indexed local array: that was a stack frame)
      // both operands MN-major: bits 15 and 16
      const uint32_t idesc = make_idesc_tf32(kTileM, p.n_pad) | (1u << 15) | (1u << 16);
      const uint32_t blk = kWgRows * 128;  // column-block stride
      const uint32_t smem_u = smem_u32(smem);
      const uint32_t full_u = smem_u32(full_bar), empty_u = smem_u32(empty_bar);
      const int stages_per_item = static_cast<int>(p.rows_per_chunk / kWgRows);
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        mbar_wait(smem_u32(&tmem_empty[acc]), ((acc_phases >> acc) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * p.n_pad);
        for (int st = 0; st < stages_per_item; ++st) {
          mbar_wait(full_u + stage * 8, phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u + static_cast<uint32_t>(stage * stage_bytes);
          const uint32_t a_lo = a_hi + a_part;
          const uint32_t g_hi = a_hi + kParts * a_part;
          const uint32_t g_lo = g_hi + g_part;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < kWgRows / 8; ++ks) {
              const uint32_t o = ks * 1024;  // 8 rows x 128 B inside every column block
              const uint32_t accum = (st | ks) ? 1u : 0u;
              if (kSplit) {
                tc_mma_tf32(tmem_d, make_desc_mn_sw128(a_lo + o, blk), make_desc_mn_sw128(g_hi + o, blk), idesc, accum);
                tc_mma_tf32(tmem_d, make_desc_mn_sw128(a_hi + o, blk), make_desc_mn_sw128(g_lo + o, blk), idesc, 1u);
                tc_mma_tf32(tmem_d, make_desc_mn_sw128(a_hi + o, blk), make_desc_mn_sw128(g_hi + o, blk), idesc, 1u);
              } else {
                tc_mma_tf32(tmem_d, make_desc_mn_sw128(a_hi + o, blk), make_desc_mn_sw128(g_hi + o, blk), idesc, accum);
              }
            }
            tc_commit(empty_u + stage * 8);
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) tc_commit(smem_u32(&tmem_full[acc]));
        __syncwarp();
        acc_phases ^= 1u << acc;
        acc ^= 1;
      }
    }
  } else if (warp >= kLoaderWarp + 1) {
    // ================= epilogue: TMEM -> red.global.add into dW[co][tap][ci] =================
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phases = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int g = item % p.groups;
      const int64_t row_begin = static_cast<int64_t>(item / p.groups) * p.rows_per_chunk;
      mbar_wait(smem_u32(&tmem_full[acc]), (acc_phases >> acc) & 1u);
      tc_fence_after();
      const int flat = g * 128 + quarter * 32 + lane;
      const int tap = flat / p.c_in;
      const int ci = flat - tap * p.c_in;
      const bool m_ok = flat < flat_m && row_begin < p.num_out;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * p.n_pad);
      for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
        uint32_t r[16];
        tc_ld16(taddr + c0, r);
        tc_wait_ld();
        if (m_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int co = static_cast<int>(blockIdx.y) * p.n_pad + c0 + j;
            const float v = __uint_as_float(r[j]);
            if (co < p.c_out && v != 0.f) atomicAdd(p.dw + (static_cast<int64_t>(co) * p.taps + tap) * p.c_in + ci, v);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
      acc_phases ^= 1u << acc;
      acc ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) function attribute: remember per device
// which of the kernels have been configured (a process may drive several GPUs).
template <typename Kernel>
static cudaError_t ensure_max_smem(Kernel kernel, int slot) {
  static unsigned char done[64][8] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && done[dev][slot]) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev][slot] = 1;
  return e;
}

static bool wgrad_supported(int c_in, int c_out, int taps) {
  const bool cin_ok = c_in >= 16 && c_in % 4 == 0 && ((128 % c_in == 0) || (c_in % 128 == 0));
  const bool cout_ok = c_out >= 16 && c_out % 16 == 0 && (c_out <= 256 || (c_out % 256 == 0 && c_out <= 4096));
  return cin_ok && cout_ok && taps >= 1 && taps <= kMaxTaps;
}

static bool supported(int c_red, int n_out, int taps) {
  const bool n_ok = n_out >= 16 && n_out % 16 == 0 && (n_out <= 256 || (n_out % 256 == 0 && n_out <= 4096));
  return c_red >= 4 && c_red % 4 == 0 && n_ok && taps >= 1 && taps <= kMaxTaps;
}

static int chunks_for(int taps, int c_red) { return (taps * c_red + kChunkK - 1) / kChunkK; }
static int chunks_for_bf16(int taps, int c_red) { return (taps * c_red + kBf16ChunkK - 1) / kBf16ChunkK; }
static bool supported_bf16(int c_red, int n_out, int taps) { return supported(c_red, n_out, taps) && c_red % 8 == 0; }
// Depth of the A ring (128-row stages of `parts` operand images) and of the weight ring for n_cta output columns:
// 227 KB minus barriers / alignment slack and the epilogue staging buffers.
struct RingDepths {
  int sa, sb;
};
static RingDepths ring_depths(int parts, int n_cta, bool staged = true) {
  const int a_bytes = parts * kTileM * 128;
  const int b_bytes = parts * n_cta * 128;
  RingDepths d;
  d.sb = b_bytes >= 65536 ? 2 : (b_bytes >= 16384 ? EFGB_TC_SB_MID : 4);
  const int budget = 227 * 1024 - 2048 - (staged ? 4 * kEpiStageBytes : 0);
  d.sa = (budget - d.sb * b_bytes) / a_bytes;
  if (d.sa > 6) d.sa = 6;
  return d;
}

static bool planes_supported(int c_red) { return c_red == 16 || c_red == 32 || (c_red >= 64 && c_red % 64 == 0); }

// fp32 [rows, C] -> planes [rows][hi C x bf16 | lo C x bf16]: hi = bf16_rn(v), lo = bf16_rn(v - hi).  One thread per 8 values.
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ in, int64_t rows, int C, uint8_t* __restrict__ planes) {
  const int upr = C / 8;
  const int64_t units = rows * upr;
  for (int64_t u = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; u < units; u += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = u / upr;
    const int k = static_cast<int>(u - row * upr);
    const float4* g = reinterpret_cast<const float4*>(in + row * C + k * 8);
    const float4 a = __ldg(g), b = __ldg(g + 1);
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 hh = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      const uint32_t hw = *reinterpret_cast<const uint32_t*>(&hh);
      const __nv_bfloat162 ll = __floats2bfloat162_rn(f[2 * i] - __uint_as_float(hw << 16), f[2 * i + 1] - __uint_as_float(hw & 0xFFFF0000u));
      hi[i] = hw;
      lo[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    uint8_t* dst = planes + row * C * 4 + k * 16;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + C * 2) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

}  // namespace tc
}  // namespace efgb

using namespace efgb;

extern "C" int efgb_spconv_tc_supported(int c_red, int n_out, int taps) { return tc::supported(c_red, n_out, taps) ? 1 : 0; }

extern "C" size_t efgb_spconv_tc_packed_bytes(int taps, int c_red, int n_out, int split) {
  if (!tc::supported(c_red, n_out, taps)) return 0;
  if (split == 2) {  // bf16x3: chunks of 64 K-values, hi + lo, 2 bytes each
    if (!tc::supported_bf16(c_red, n_out, taps)) return 0;
    return static_cast<size_t>(tc::chunks_for_bf16(taps, c_red)) * 2 * n_out * tc::kBf16ChunkK * sizeof(__nv_bfloat16);
  }
  return static_cast<size_t>(tc::chunks_for(taps, c_red)) * (split ? 2 : 1) * n_out * tc::kChunkK * sizeof(float);
}

extern "C" int efgb_spconv_tc_pack(const float* w_param, int c_out, int taps, int c_in, int mode, int split, float* packed,
                                   efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(w_param && packed && mode >= 0 && mode <= 2, EFGB_EINVAL, "spconv_tc_pack: bad argument");
  const int n_out = mode == 0 ? c_out : c_in;
  const int c_red = mode == 0 ? c_in : c_out;
  EFGB_REQUIRE(tc::supported(c_red, n_out, taps), EFGB_EINVAL, "spconv_tc_pack: unsupported shape (c_red=%d n_out=%d taps=%d)",
               c_red, n_out, taps);
  if (split == 2) {
    EFGB_REQUIRE(tc::supported_bf16(c_red, n_out, taps), EFGB_EINVAL, "spconv_tc_pack: bf16x3 needs c_red %% 8 == 0 (c_red=%d)", c_red);
    const int chunks16 = tc::chunks_for_bf16(taps, c_red);
    const int64_t total16 = static_cast<int64_t>(chunks16) * n_out * tc::kBf16ChunkK;
    tc::pack_weights_bf16_kernel<<<static_cast<unsigned>((total16 + 255) / 256), 256, 0, stream>>>(
        w_param, c_out, taps, c_in, mode, n_out, c_red, chunks16, reinterpret_cast<__nv_bfloat16*>(packed));
    EFGB_LAUNCH_OK("pack_weights_bf16_kernel");
    return EFGB_OK;
  }
  const int chunks = tc::chunks_for(taps, c_red);
  const int64_t total = static_cast<int64_t>(chunks) * n_out * tc::kChunkK;
  const unsigned nb = static_cast<unsigned>((total + 255) / 256);
  if (split)
    tc::pack_weights_kernel<true><<<nb, 256, 0, stream>>>(w_param, c_out, taps, c_in, mode, n_out, c_red, chunks, packed);
  else
    tc::pack_weights_kernel<false><<<nb, 256, 0, stream>>>(w_param, c_out, taps, c_in, mode, n_out, c_red, chunks, packed);
  EFGB_LAUNCH_OK("pack_weights_kernel");
  return EFGB_OK;
}

extern "C" int64_t efgb_spconv_tc_pack_blocks(int c_out, int taps, int c_in, int mode) {
  const int n_out = mode == 0 ? c_out : c_in;
  const int c_red = mode == 0 ? c_in : c_out;
  if (mode < 0 || mode > 2 || !tc::supported_bf16(c_red, n_out, taps)) return 0;
  return (static_cast<int64_t>(tc::chunks_for_bf16(taps, c_red)) * n_out * tc::kBf16ChunkK + 255) / 256;
}

extern "C" int efgb_spconv_tc_pack_batched(const int64_t* jobs, int n_jobs, int64_t total_blocks, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(n_jobs >= 0 && total_blocks >= 0 && total_blocks < (1ll << 31), EFGB_EINVAL, "spconv_tc_pack_batched: bad argument");
  if (n_jobs == 0 || total_blocks == 0) return EFGB_OK;
  EFGB_REQUIRE(jobs != nullptr, EFGB_EINVAL, "spconv_tc_pack_batched: null job table");
  tc::pack_weights_bf16_batched_kernel<<<static_cast<unsigned>(total_blocks), 256, 0, stream>>>(jobs, n_jobs);
  EFGB_LAUNCH_OK("pack_weights_bf16_batched_kernel");
  return EFGB_OK;
}

extern "C" int efgb_spconv_tc_forward_ex(const float* in_feats, int64_t num_in, int c_red, const float* packed,
                                         const float* bias, const int32_t* nbr, int64_t num_out, int taps, int n_out,
                                         int split, int relu, float* out_feats, efgb_stream_t stream_);

extern "C" int efgb_spconv_tc_forward(const float* in_feats, int64_t num_in, int c_red, const float* packed,
                                      const float* bias, const int32_t* nbr, int64_t num_out, int taps, int n_out,
                                      int split, float* out_feats, efgb_stream_t stream_) {
  return efgb_spconv_tc_forward_ex(in_feats, num_in, c_red, packed, bias, nbr, num_out, taps, n_out, split, 0, out_feats,
                                   stream_);
}

static int launch_forward(const float* in_feats, const void* planes, int64_t num_in, int c_red, const float* packed,
                          const float* bias, const int32_t* nbr, int64_t num_out, int taps, int n_out, int split, int relu,
                          float* out_feats, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(split >= 0 && split <= 2, EFGB_EINVAL, "spconv_tc_forward: split must be 0 (tf32), 1 (tf32x3) or 2 (bf16x3)");
  EFGB_REQUIRE(tc::supported(c_red, n_out, taps) && (split != 2 || tc::supported_bf16(c_red, n_out, taps)), EFGB_EINVAL,
               "spconv_tc_forward: unsupported shape (c_red=%d n_out=%d taps=%d split=%d)", c_red, n_out, taps, split);
  EFGB_REQUIRE(num_in >= 0 && num_out >= 0, EFGB_EINVAL, "spconv_tc_forward: bad sizes");
  EFGB_REQUIRE(nbr != nullptr || (taps == 1 && num_in >= num_out), EFGB_EINVAL,
               "spconv_tc_forward: a null rulebook means identity and needs taps == 1");
  if (num_out == 0) return EFGB_OK;
  EFGB_REQUIRE(packed && out_feats && (in_feats || planes || num_in == 0), EFGB_EINVAL, "spconv_tc_forward: null pointer");
  EFGB_REQUIRE(planes == nullptr || (split == 2 && nbr != nullptr && tc::planes_supported(c_red) &&
                                     (reinterpret_cast<uintptr_t>(planes) & 15) == 0),
               EFGB_EINVAL, "spconv_tc_forward: pre-split planes need bf16x3, a rulebook and c_red in {16, 32, 64 k} (c_red=%d)", c_red);
  EFGB_REQUIRE((reinterpret_cast<uintptr_t>(in_feats) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_feats) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(packed) & 15) == 0,
               EFGB_EINVAL, "spconv_tc_forward: feature / weight pointers must be 16-byte aligned");
  EFGB_REQUIRE(static_cast<uint64_t>(num_in) * static_cast<uint64_t>(c_red) < (1ull << 32) && num_out < (1ll << 31) - 256,
               EFGB_EINVAL, "spconv_tc_forward: input larger than 2^32 elements (the producers use 32-bit element offsets)");
  tc::Params p;
  p.in = in_feats;
  p.planes = reinterpret_cast<const uint8_t*>(planes);
  p.packed = reinterpret_cast<const uint8_t*>(packed);
  p.bias = bias;
  p.relu = relu ? 1 : 0;
  p.nbr = nbr;
  p.wide = (split == 2 && !planes && (reinterpret_cast<uintptr_t>(in_feats) & 31) == 0) ? 1 : 0;
  p.out = out_feats;
  p.num_out = num_out;
  p.c_red = c_red;
  p.taps = taps;
  p.n_out = n_out;
  p.chunks = split == 2 ? tc::chunks_for_bf16(taps, c_red) : tc::chunks_for(taps, c_red);
  p.num_tiles = static_cast<int>((num_out + tc::kTileM - 1) / tc::kTileM);
  // N split: wide outputs (dense GEMMs) are cut into column slabs over grid.y — as wide as leaves room for at least
  // three A stages next to the weight ring and the epilogue staging (256 columns for one operand image, 128 for two)
  const int parts = split ? 2 : 1;
  const int a_bytes = parts * tc::kTileM * 128;
  int n_split = n_out > 256 ? n_out / 256 : 1;
  int n_cta = n_out / n_split;
  // The epilogue's staging buffers do not fit next to a 256-column weight ring with two operand images.  Short
  // reductions (K <= 512: the tile's epilogue is as long as its main loop) take 128-column CTAs with the coalescing
  // epilogue; long ones keep the 256-column CTA (half the A production per output) and the row-per-thread stores, which
  // their main loop hides.  Measured on 70 k x {256, 1024} token-wise linears, profiles/r2_dense_micro.txt.
  p.epi_staged = 1;
  if (tc::ring_depths(parts, n_cta).sa < 3) {
    if (p.chunks > 8 && tc::ring_depths(parts, n_cta, false).sa >= 3) {
      p.epi_staged = 0;
    } else {
      while (tc::ring_depths(parts, n_cta).sa < 3 && n_cta / 2 >= 32 && (n_cta / 2) % 16 == 0) {
        n_split *= 2;
        n_cta /= 2;
      }
    }
  }
  // super-tile: T row tiles share each weight chunk (T * acc_cols fp32 columns per accumulator set, two sets in TMEM)
  auto acc_cols_of = [&](int n) { return (split != 0 && n <= 128) ? 2 * n : n; };
  int T = 256 / acc_cols_of(n_cta);
  if (T > EFGB_TC_MAX_T) T = EFGB_TC_MAX_T;
  if (T < 1) T = 1;
  // keep at least ~3/4 of the SMs busy; beyond that, sharing weight chunks across more tiles wins (the weight
  // stream from L2 is the bound for C >= 64)
  while (T > 1 && ((p.num_tiles + T - 1) / T) * n_split < (kNumSMs * 3) / 4) T >>= 1;
  if (T == 1) {
    // few tiles (deep levels): split the output channels further so that all SMs get work
    while (p.num_tiles * n_split * 2 <= kNumSMs && n_cta / 2 >= 32 && (n_cta / 2) % 16 == 0) {
      n_split *= 2;
      n_cta /= 2;
    }
  }
  p.n_cta = n_cta;
  p.acc_cols = acc_cols_of(n_cta);
  p.concat = p.acc_cols != n_cta ? 1 : 0;
  p.tiles_per_super = T;
  p.num_super = (p.num_tiles + T - 1) / T;
  const int b_bytes = parts * n_cta * 128;
  if (!p.epi_staged && tc::ring_depths(parts, n_cta).sa >= 3) p.epi_staged = 1;   // n_cta was halved after the choice
  const tc::RingDepths depths = tc::ring_depths(parts, n_cta, p.epi_staged != 0);
  p.sa = depths.sa;
  p.sb = depths.sb;
  EFGB_REQUIRE(p.sa >= 2, EFGB_EINVAL, "spconv_tc_forward: tile does not fit shared memory");
  EFGB_REQUIRE(!planes || p.sa >= tc::kPlaneDepth + 2, EFGB_EINVAL, "spconv_tc_forward: A ring too shallow for the cp.async producers");
  p.cred_shift = -1;
  for (int sft = 2; sft < 16; ++sft)
    if ((1 << sft) == c_red) p.cred_shift = sft;
  const size_t smem = 1024 + static_cast<size_t>(p.sa) * a_bytes + static_cast<size_t>(p.sb) * b_bytes +
                      (2 * p.sa + 2 * p.sb + 4) * 8 + 32 + (p.epi_staged ? 4 * tc::kEpiStageBytes : 0);
  int gx = kNumSMs / n_split;
  if (gx < 1) gx = 1;
  if (gx > p.num_super) gx = p.num_super;
  const dim3 grid(gx, n_split);
  if (split == 2) {
    EFGB_CUDA_OK(tc::ensure_max_smem(tc::spconv_tc_kernel<tc::kBf16x3>, 0));
    tc::spconv_tc_kernel<tc::kBf16x3><<<grid, tc::kThreads, smem, stream>>>(p);
  } else if (split == 1) {
    EFGB_CUDA_OK(tc::ensure_max_smem(tc::spconv_tc_kernel<tc::kTf32x3>, 1));
    tc::spconv_tc_kernel<tc::kTf32x3><<<grid, tc::kThreads, smem, stream>>>(p);
  } else {
    EFGB_CUDA_OK(tc::ensure_max_smem(tc::spconv_tc_kernel<tc::kTf32>, 2));
    tc::spconv_tc_kernel<tc::kTf32><<<grid, tc::kThreads, smem, stream>>>(p);
  }
  EFGB_LAUNCH_OK("spconv_tc_kernel");
  return EFGB_OK;
}

extern "C" int efgb_spconv_tc_forward_ex(const float* in_feats, int64_t num_in, int c_red, const float* packed,
                                         const float* bias, const int32_t* nbr, int64_t num_out, int taps, int n_out,
                                         int split, int relu, float* out_feats, efgb_stream_t stream_) {
  return launch_forward(in_feats, nullptr, num_in, c_red, packed, bias, nbr, num_out, taps, n_out, split, relu, out_feats, stream_);
}

#if EFGB_TC_TRACE
extern "C" int efgb_debug_trace_read(long long* host, int stages) {
  if (stages > tc::kTraceStages) stages = tc::kTraceStages;
  EFGB_CUDA_OK(cudaDeviceSynchronize());
  EFGB_CUDA_OK(cudaMemcpyFromSymbol(host, tc::g_trace, sizeof(long long) * 12 * stages));
  return stages;
}
#endif

extern "C" int efgb_spconv_tc_planes_supported(int c_red, int n_out, int taps) {
  if (!tc::supported_bf16(c_red, n_out, taps) || !tc::planes_supported(c_red)) return 0;
  // the cp.async producers need kPlaneDepth + 2 A stages next to the weight ring
  int n_cta = n_out > 256 ? 256 : n_out;
  while (tc::ring_depths(2, n_cta).sa < 3 && n_cta / 2 >= 32 && (n_cta / 2) % 16 == 0) n_cta /= 2;
  return tc::ring_depths(2, n_cta).sa >= tc::kPlaneDepth + 2 ? 1 : 0;
}

extern "C" int efgb_spconv_tc_forward_planes(const void* in_planes, int64_t num_in, int c_red, const float* packed,
                                             const float* bias, const int32_t* nbr, int64_t num_out, int taps, int n_out,
                                             int relu, float* out_feats, efgb_stream_t stream_) {
  EFGB_REQUIRE(in_planes || num_in == 0 || num_out == 0, EFGB_EINVAL, "spconv_tc_forward_planes: null input");
  EFGB_REQUIRE(efgb_spconv_tc_planes_supported(c_red, n_out, taps), EFGB_EINVAL,
               "spconv_tc_forward_planes: unsupported shape (c_red=%d n_out=%d taps=%d)", c_red, n_out, taps);
  return launch_forward(nullptr, in_planes, num_in, c_red, packed, bias, nbr, num_out, taps, n_out, 2, relu, out_feats, stream_);
}

extern "C" int efgb_split_bf16(const float* in, int64_t rows, int channels, void* planes, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(rows >= 0 && channels >= 8 && channels % 8 == 0, EFGB_EINVAL, "split_bf16: channels must be a multiple of 8");
  if (rows == 0) return EFGB_OK;
  EFGB_REQUIRE(in && planes && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(planes) & 15) == 0,
               EFGB_EINVAL, "split_bf16: null or misaligned pointer");
  const int64_t units = rows * (channels / 8);
  tc::split_bf16_kernel<<<grid_for(units, 256), 256, 0, stream>>>(in, rows, channels, reinterpret_cast<uint8_t*>(planes));
  EFGB_LAUNCH_OK("split_bf16_kernel");
  return EFGB_OK;
}

extern "C" int efgb_spconv_tc_wgrad_supported(int c_in, int c_out, int taps) { return tc::wgrad_supported(c_in, c_out, taps) ? 1 : 0; }

extern "C" int efgb_spconv_tc_wgrad(const float* in_feats, int64_t num_in, int c_in, const float* grad_out,
                                    const int32_t* nbr, int64_t num_out, int taps, int c_out, int split, float* dw_param,
                                    efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(tc::wgrad_supported(c_in, c_out, taps), EFGB_EINVAL, "spconv_tc_wgrad: unsupported shape (c_in=%d c_out=%d taps=%d)",
               c_in, c_out, taps);
  EFGB_REQUIRE(num_in >= 0 && num_out >= 0 && dw_param, EFGB_EINVAL, "spconv_tc_wgrad: bad argument");
  EFGB_CUDA_OK(cudaMemsetAsync(dw_param, 0, static_cast<size_t>(c_out) * taps * c_in * sizeof(float), stream));
  if (num_out == 0 || num_in == 0) return EFGB_OK;
  EFGB_REQUIRE(in_feats && grad_out, EFGB_EINVAL, "spconv_tc_wgrad: null pointer");
  EFGB_REQUIRE(nbr != nullptr || (taps == 1 && num_in >= num_out), EFGB_EINVAL,
               "spconv_tc_wgrad: a null rulebook means identity and needs taps == 1");
  EFGB_REQUIRE((reinterpret_cast<uintptr_t>(in_feats) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad_out) & 15) == 0, EFGB_EINVAL,
               "spconv_tc_wgrad: feature pointers must be 16-byte aligned");
  tc::WgradParams p;
  p.in = in_feats;
  p.gout = grad_out;
  p.nbr = nbr;
  p.dw = dw_param;
  p.num_out = num_out;
  p.c_in = c_in;
  p.c_out = c_out;
  p.taps = taps;
  // output-channel slabs of at most kSlab columns per CTA (128-column slabs were measured slower: more re-reads of A)
  const int kSlab = 256;
  const int n_slabs = c_out > kSlab ? c_out / kSlab : 1;
  p.n_pad = c_out > kSlab ? kSlab : (c_out < 32 ? 32 : (c_out + 31) / 32 * 32);
  p.groups = (taps * c_in + 127) / 128;
  int64_t chunks = (kNumSMs * 3 + p.groups * n_slabs - 1) / (p.groups * n_slabs);
  const int64_t max_chunks = (num_out + 255) / 256;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int64_t rows = (num_out + chunks - 1) / chunks;
  rows = (rows + tc::kWgRows - 1) / tc::kWgRows * tc::kWgRows;
  p.rows_per_chunk = rows;
  p.row_chunks = static_cast<int>((num_out + rows - 1) / rows);
  p.num_items = p.groups * p.row_chunks;
  const int parts = split ? 2 : 1;
  const int stage_bytes = parts * (4 * tc::kWgRows * 128 + (p.n_pad / 32) * tc::kWgRows * 128);
  int stages = (227 * 1024 - 4096) / stage_bytes;
  if (stages > 6) stages = 6;
  EFGB_REQUIRE(stages >= 2, EFGB_EINVAL, "spconv_tc_wgrad: tile does not fit shared memory");
  const size_t smem = 1024 + static_cast<size_t>(stages) * stage_bytes + (2 * stages + 4) * 8 + 16;
  int gx = kNumSMs / n_slabs;
  if (gx < 1) gx = 1;
  if (gx > p.num_items) gx = p.num_items;
  const dim3 grid(gx, n_slabs);
  if (split) {
    EFGB_CUDA_OK(tc::ensure_max_smem(tc::spconv_wgrad_tc_kernel<true>, 3));
    tc::spconv_wgrad_tc_kernel<true><<<grid, tc::kThreads, smem, stream>>>(p, stages);
  } else {
    EFGB_CUDA_OK(tc::ensure_max_smem(tc::spconv_wgrad_tc_kernel<false>, 4));
    tc::spconv_wgrad_tc_kernel<false><<<grid, tc::kThreads, smem, stream>>>(p, stages);
  }
  EFGB_LAUNCH_OK("spconv_wgrad_tc_kernel");
  return EFGB_OK;
}
