// Sparse convolution forward / dgrad on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Output-stationary implicit GEMM over the device rulebook:
//     D[128 output rows x N] = sum over K = (tap, channel)  A[row, K] * B[n, K]
//   A[row, (tap, c)] = in[nbr[row, tap], c]     gathered on the fly (zeros where nbr = -1)
//   B[n,   (tap, c)] = weight element           pre-packed once per call into the exact shared-memory
//                                               image (K-major, 128-byte swizzle) the tensor core reads
// K is consumed in chunks of 32 floats = one 128-byte swizzled row per operand row.
//
// Roles inside one persistent CTA (one CTA per SM, tiles strided over the grid):
//   (warp numbers for the default of 16 producer warps; measured 1.3x faster than 8 on the producer-bound layers)
//   warps 0-15  A producers (four groups of 4 warps, each owning every 4th K chunk): 8 lanes gather one 128-byte
//               row piece with coalesced LDG.128 (16 independent loads in flight per thread), split it
//               into tf32 hi/lo parts, and store it swizzled into the stage's A tiles
//   warp  16    MMA issuer: one elected lane issues tcgen05.mma.kind::tf32 (M=128, N, K=8) into TMEM;
//               tcgen05.commit releases the stage / publishes the accumulator
//   warp  17    B loader: one lane streams the packed weight chunk with cp.async.bulk (TMA engine,
//               mbarrier complete_tx) — weights stay L2-resident
//   warps 18-21 epilogue: tcgen05.ld the fp32 accumulator (double-buffered in TMEM), add bias, store rows
//
// Precision: kSplit = true runs the 3xTF32 scheme (a = a_hi + a_lo, b = b_hi + b_lo;
// a_hi*b_hi + a_lo*b_hi + a_hi*b_lo with fp32 accumulation), error ~2^-21 relative, i.e. fp32-faithful
// (the reference runs this path in fp32).  kSplit = false is single-pass TF32.
#include <stdlib.h>

#include "common.cuh"

namespace efgb {
namespace tc {

constexpr int kTileM = 128;
constexpr int kChunkK = 32;             // floats per K chunk (128 bytes)
#ifndef EFGB_TC_PRODUCER_WARPS
#define EFGB_TC_PRODUCER_WARPS 16
#endif
constexpr int kProducerWarps = EFGB_TC_PRODUCER_WARPS;   // gather / split warps (8 or 16)
constexpr int kMmaWarp = kProducerWarps;
constexpr int kLoaderWarp = kProducerWarps + 1;
constexpr int kThreads = (kProducerWarps + 6) * 32;      // + MMA issuer, weight loader, 4 epilogue warps
static_assert(kProducerWarps == 8 || kProducerWarps == 16, "producer warp groups assume 8 or 16 warps");
static_assert((kProducerWarps + 2) % 4 == 2, "epilogue warps must cover the four TMEM lane quarters");
constexpr int kMaxTaps = 32;

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
// Ampere-style asynchronous 16-byte copy global -> shared; src_bytes = 0 writes zeros (missing neighbour).
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// The mbarrier receives one arrival from this thread once all of its earlier cp.async copies have landed
// (.noinc: the arrival is part of the barrier's initial count).
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
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

// kind::tf32 instruction descriptor (mma_sm100_desc.hpp:InstrDescriptor): D=f32 [4,6)=1, A=tf32 [7,10)=2,
// B=tf32 [10,13)=2, both K-major, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

struct Params {
  const float* in;       // [num_in, c_red]
  const float* packed;   // [chunks][parts][n_out][32] swizzled image
  const float* bias;     // [n_out] or null
  int relu;              // epilogue applies max(x, 0) after the bias
  const int32_t* nbr;    // [num_out, taps], or null = identity (dense GEMM: taps == 1, src row = out row)
  float* out;            // [num_out, n_out]
  int64_t num_out;
  int c_red, taps, n_out, chunks;
  int num_tiles;         // 128-row tiles
  int n_cta;             // output columns per CTA (n_out / gridDim.y)
  int tiles_per_super;   // T: row tiles that share every weight chunk (accumulators side by side in TMEM)
  int num_super;         // ceil(num_tiles / T)
  int sa, sb;            // A-ring / B-ring depth
  int async_gather;      // 1: cp.async producers (produce_a_async), 0: register-staged producers (produce_a)
  int cred_shift;        // log2(c_red) when c_red is a power of two, else -1
  int debug;             // perf experiments only: 1 = no MMAs, 2 = no gather loads, 4 = no index loads,
                         // 64 = async producers leave the raw fp32 words as the hi operand (relies on the tensor core
                         // ignoring the 13 low mantissa bits)
};

template <bool kSplit>
struct Smem {
  static constexpr int kParts = kSplit ? 2 : 1;
  static __host__ __device__ int a_bytes() { return kParts * kTileM * 128; }
  static __host__ __device__ int b_bytes(int n) { return kParts * n * 128; }
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
  // A-stage index -> (tile, chunk)
  __device__ __forceinline__ void locate(const Params& p, int ga, int* tile, int* chunk) const {
    const int full = (n_super - 1) * p.chunks * p.tiles_per_super;
    int i, c, t;
    if (ga < full) {
      const int per = p.chunks * p.tiles_per_super;
      i = ga / per;
      const int rem = ga - i * per;
      c = rem / p.tiles_per_super;
      t = rem - c * p.tiles_per_super;
    } else {
      const int rem = ga - full;
      i = n_super - 1;
      c = rem / t_last;
      t = rem - c * t_last;
    }
    *tile = (static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x)) * p.tiles_per_super + t;
    *chunk = c;
  }
};

// A producers.  kGroups warp groups; group g owns every kGroups-th A stage of this CTA's stream, so kGroups
// stages are gathered concurrently and every thread keeps 1024/(threads per group) independent 16-byte loads
// in flight.  The rulebook entries of a group's NEXT stage are fetched while the gathers of the current one
// are in flight, and the smem slot is only waited for after the loads were issued, so a stage costs one memory
// latency.  kGroups must not exceed the A-ring depth (mbarrier parity would alias).
template <bool kSplit, int kGroups>
__device__ __forceinline__ void produce_a(const Params& p, const CtaWork& w, uint8_t* a_ring, uint64_t* a_full,
                                          uint64_t* a_empty, const int warp, const int lane) {
  constexpr int kWarpsPerGroup = kProducerWarps / kGroups;
  constexpr int kPieces = (kTileM * 8) / (kWarpsPerGroup * 32);  // 16-byte pieces per thread per stage
  const int a_part = kTileM * 128;
  const int a_bytes = Smem<kSplit>::a_bytes();
  const int gidx = warp / kWarpsPerGroup;
  const int gw = warp % kWarpsPerGroup;
  const int row_in_group = lane >> 3;  // 0..3
  const int q = lane & 7;              // 16-byte piece of the 128-byte row

  int32_t src_next[kPieces];
  int tile_n = 0, chunk_n = 0;
  auto load_indices = [&](int ga) {
    w.locate(p, ga, &tile_n, &chunk_n);
    const int64_t r0 = static_cast<int64_t>(tile_n) * kTileM;
    const int kk = chunk_n * kChunkK + q * 4;
    const int tap = kk / p.c_red;
    const bool tap_ok = tap < p.taps;
#pragma unroll
    for (int i = 0; i < kPieces; ++i) {
      const int64_t row = r0 + (i * kWarpsPerGroup + gw) * 4 + row_in_group;
      int32_t v = -1;
      if (tap_ok && row < p.num_out && !(p.debug & 4)) v = p.nbr ? __ldg(p.nbr + row * p.taps + tap) : static_cast<int32_t>(row);
      src_next[i] = v;
    }
  };

  int ga = gidx;
  if (ga < w.total_a) load_indices(ga);
  while (ga < w.total_a) {
    const int kk = chunk_n * kChunkK + q * 4;
    const int tap = kk / p.c_red;
    const int ci = kk - tap * p.c_red;
    float4 v[kPieces];
#pragma unroll
    for (int i = 0; i < kPieces; ++i) {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src_next[i] >= 0 && !(p.debug & 2))
        v[i] = __ldg(reinterpret_cast<const float4*>(p.in + static_cast<int64_t>(src_next[i]) * p.c_red + ci));
    }
    const int round = ga / p.sa;
    const int stage = ga - round * p.sa;
    const uint32_t phase = static_cast<uint32_t>(round & 1);
    const int ga_next = ga + kGroups;
    if (ga_next < w.total_a) load_indices(ga_next);  // overlaps with the gathers above
    mbar_wait(smem_u32(&a_empty[stage]), phase ^ 1);
    uint8_t* a_hi = a_ring + static_cast<size_t>(stage) * a_bytes;
    uint8_t* a_lo = a_hi + a_part;
#pragma unroll
    for (int i = 0; i < kPieces; ++i) {
      if (p.debug & 8) break;
      const int r = (i * kWarpsPerGroup + gw) * 4 + row_in_group;
      const uint32_t off = static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>(q ^ (r & 7)) << 4);
      if (kSplit) {
        float4 hi, lo;
        hi.x = __uint_as_float(__float_as_uint(v[i].x) & 0xFFFFE000u);
        hi.y = __uint_as_float(__float_as_uint(v[i].y) & 0xFFFFE000u);
        hi.z = __uint_as_float(__float_as_uint(v[i].z) & 0xFFFFE000u);
        hi.w = __uint_as_float(__float_as_uint(v[i].w) & 0xFFFFE000u);
        lo.x = v[i].x - hi.x;
        lo.y = v[i].y - hi.y;
        lo.z = v[i].z - hi.z;
        lo.w = v[i].w - hi.w;
        *reinterpret_cast<float4*>(a_hi + off) = hi;
        *reinterpret_cast<float4*>(a_lo + off) = lo;
      } else {
        *reinterpret_cast<float4*>(a_hi + off) = v[i];
      }
    }
    if (!(p.debug & 16)) fence_proxy_async();  // make the generic-proxy stores visible to the tensor core (async proxy)
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&a_full[stage]));
    ga = ga_next;
  }
}

// A producers, asynchronous variant.  All producer warps work on the SAME stage; a thread owns two 16-byte pieces
// of the 128 x 128 B tile.  The gather itself is cp.async (LDGSTS): the row piece lands in its swizzled slot of the
// stage without passing through registers, so `sa - 2` stages of gathers are in flight per CTA while the oldest
// landed stage is split into tf32 hi / lo in place (LDS, 8 ALU ops, 2 STS per piece) — neither the rulebook
// latency nor the gather latency sits on the critical path of a stage any more.
//   a_empty[slot] (MMA commit)  ->  cp.async into slot, cp.async.mbarrier.arrive on raw_full[slot]
//   raw_full[slot] (all copies landed)  ->  split in place, fence.proxy.async, a_full[slot]  ->  MMA
template <bool kSplit>
__device__ __forceinline__ void produce_a_async(const Params& p, const CtaWork& w, uint8_t* a_ring, uint64_t* a_full,
                                                uint64_t* a_empty, uint64_t* raw_full, const int tid, const int lane) {
  constexpr int kThreadsP = kProducerWarps * 32;
  constexpr int kPieces = (kTileM * 8) / kThreadsP;    // 16-byte pieces per thread per stage
  constexpr int kRowsPerPass = kThreadsP / 8;
  const int a_part = kTileM * 128;
  const int a_bytes = Smem<kSplit>::a_bytes();
  const int q = tid & 7;            // 16-byte piece of the 128-byte row
  const int r_base = tid >> 3;      // row of piece 0; piece i is row r_base + i * kRowsPerPass
  uint32_t off[kPieces];
#pragma unroll
  for (int i = 0; i < kPieces; ++i) {
    const int r = r_base + i * kRowsPerPass;
    off[i] = static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>(q ^ (r & 7)) << 4);
  }
  const int total = w.total_a;
  if (total == 0) return;
  const int depth = p.sa - 2;       // stages of gathers in flight ahead of the one being split (host ensures sa >= 3)

  // cursor over this CTA's stage stream (super-tile, chunk, tile-in-super-tile), advanced without divisions
  int ci_s = 0, cc = 0, ct = 0, cti = w.tiles_in(p, 0);
  int32_t src[kPieces];
  int ci = 0;
  auto load_idx = [&]() {
    const int tile = (static_cast<int>(blockIdx.x) + ci_s * static_cast<int>(gridDim.x)) * p.tiles_per_super + ct;
    const int64_t r0 = static_cast<int64_t>(tile) * kTileM;
    const int kk = cc * kChunkK + q * 4;
    const int tap = p.cred_shift >= 0 ? (kk >> p.cred_shift) : kk / p.c_red;
    ci = kk - tap * p.c_red;
    const bool tap_ok = tap < p.taps;
#pragma unroll
    for (int i = 0; i < kPieces; ++i) {
      const int64_t row = r0 + r_base + i * kRowsPerPass;
      int32_t v = -1;
      if (tap_ok && row < p.num_out) v = p.nbr ? __ldg(p.nbr + row * p.taps + tap) : static_cast<int32_t>(row);
      src[i] = v;
    }
    if (++ct == cti) {  // advance the cursor
      ct = 0;
      if (++cc == p.chunks) {
        cc = 0;
        ++ci_s;
        cti = w.tiles_in(p, ci_s);
      }
    }
  };

  int islot = 0;
  uint32_t iphase = 0;
  int issued = 0;
  auto issue = [&]() {  // gather stage `issued` with the indices loaded by the previous call
    mbar_wait(smem_u32(&a_empty[islot]), iphase ^ 1);
    const uint32_t dst = smem_u32(a_ring + static_cast<size_t>(islot) * a_bytes);
#pragma unroll
    for (int i = 0; i < kPieces; ++i) {
      const bool ok = src[i] >= 0;
      const float* g = ok ? p.in + static_cast<int64_t>(src[i]) * p.c_red + ci : p.in;
      cp_async_16(dst + off[i], g, ok ? 16u : 0u);
    }
    cp_async_mbar_arrive_noinc(smem_u32(&raw_full[islot]));
    if (++islot == p.sa) {
      islot = 0;
      iphase ^= 1;
    }
    if (++issued < total) load_idx();  // consumed by the next issue(); its latency hides behind the split below
  };

  load_idx();
  for (int t = 0; t < depth && t < total; ++t) issue();
  int sslot = 0;
  uint32_t sphase = 0;
  for (int s = 0; s < total; ++s) {
    if (issued < total) issue();
    mbar_wait(smem_u32(&raw_full[sslot]), sphase);
    uint8_t* a_hi = a_ring + static_cast<size_t>(sslot) * a_bytes;
    uint8_t* a_lo = a_hi + a_part;
    if (kSplit) {
#pragma unroll
      for (int i = 0; i < kPieces; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(a_hi + off[i]);
        float4 hi, lo;
        hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
        hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
        hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
        lo.x = v.x - hi.x;
        lo.y = v.y - hi.y;
        lo.z = v.z - hi.z;
        lo.w = v.w - hi.w;
        if (!(p.debug & 64)) *reinterpret_cast<float4*>(a_hi + off[i]) = hi;
        *reinterpret_cast<float4*>(a_lo + off[i]) = lo;
      }
    }
    fence_proxy_async();  // cp.async / st.shared writes -> visible to the tensor core (async proxy)
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&a_full[sslot]));
    if (++sslot == p.sa) {
      sslot = 0;
      sphase ^= 1;
    }
  }
}

template <bool kSplit>
__global__ void __launch_bounds__(kThreads, 1) spconv_tc_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the swizzle atoms
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kParts = Smem<kSplit>::kParts;
  const int n_cta = p.n_cta;  // output columns owned by this CTA (N split over grid.y)
  const int n0 = static_cast<int>(blockIdx.y) * n_cta;
  const int T = p.tiles_per_super;
  const int a_bytes = Smem<kSplit>::a_bytes();
  const int b_bytes = Smem<kSplit>::b_bytes(n_cta);
  const int a_part = kTileM * 128;
  const int b_part = n_cta * 128;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + static_cast<size_t>(p.sa) * a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + static_cast<size_t>(p.sb) * b_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.sa;
  uint64_t* b_full = a_empty + p.sa;
  uint64_t* b_empty = b_full + p.sb;
  uint64_t* tmem_full = b_empty + p.sb;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* raw_full = tmem_empty + 2;   // [sa], async producers only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_full + p.sa);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int groups = p.sa >= 4 ? 4 : 2;  // producer warp groups (never more than A stages)

  // TMEM: two sets of T accumulators of n_cta fp32 columns each, power of two >= 32
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(2 * T * n_cta)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.sa; ++s) {
      mbar_init(smem_u32(&a_full[s]), p.async_gather ? kProducerWarps : kProducerWarps / groups);
      mbar_init(smem_u32(&a_empty[s]), 1);
      mbar_init(smem_u32(&raw_full[s]), kProducerWarps * 32);
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

  if (warp < kProducerWarps) {
    // ================= A producers =================
    if (p.async_gather)
      produce_a_async<kSplit>(p, w, a_ring, a_full, a_empty, raw_full, static_cast<int>(threadIdx.x), lane);
    else if (groups == 4)
      produce_a<kSplit, 4>(p, w, a_ring, a_full, a_empty, warp, lane);
    else
      produce_a<kSplit, 2>(p, w, a_ring, a_full, a_empty, warp, lane);
  } else if (warp == kLoaderWarp) {
    // ================= B loader: weight chunks through their own ring (TMA bulk copies) =================
    // A chunk serves all T tiles of the super-tile, and the ring runs ahead of the A stream, so neither the
    // L2 latency nor the L2 bandwidth of the weights sits on the A-stage turnaround.
    if (lane == 0) {
      const uint32_t part_bytes = static_cast<uint32_t>(n_cta) * 128u;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < w.n_super; ++i) {
        for (int c = 0; c < p.chunks; ++c) {
          mbar_wait(smem_u32(&b_empty[stage]), phase ^ 1);
          const uint32_t bar = smem_u32(&b_full[stage]);
          mbar_arrive_expect_tx(bar, static_cast<uint32_t>(b_bytes));
          const uint32_t dst = smem_u32(b_ring + static_cast<size_t>(stage) * b_bytes);
          for (int part = 0; part < kParts; ++part) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.packed) +
                                 (static_cast<size_t>(c * kParts + part) * p.n_out + n0) * 128u;
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
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(kTileM, n_cta);
      int sa_ = 0, sb_ = 0;
      uint32_t pa = 0, pb = 0;
      for (int i = 0; i < w.n_super; ++i) {
        const int acc = i & 1;
        const uint32_t acc_phase = static_cast<uint32_t>((i >> 1) & 1);
        const int ti = w.tiles_in(p, i);
        mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
        tc_fence_after();
        for (int c = 0; c < p.chunks; ++c) {
          mbar_wait(smem_u32(&b_full[sb_]), pb);
          const uint32_t b_hi = smem_u32(b_ring + static_cast<size_t>(sb_) * b_bytes);
          const uint64_t db_hi = make_desc_sw128(b_hi);
          const uint64_t db_lo = make_desc_sw128(b_hi + b_part);
          for (int t = 0; t < ti; ++t) {
            mbar_wait(smem_u32(&a_full[sa_]), pa);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(a_ring + static_cast<size_t>(sa_) * a_bytes);
            const uint64_t da_hi = make_desc_sw128(a_hi);
            const uint64_t da_lo = make_desc_sw128(a_hi + a_part);
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>((acc * T + t) * n_cta);
#pragma unroll
            for (int j = 0; j < kChunkK / 8; ++j) {
              if (p.debug & 1) break;
              const uint64_t adv = static_cast<uint64_t>(j * 2);  // 8 tf32 = 32 bytes = 2 x 16 B
              if (kSplit) {
                tc_mma_tf32(tmem_d, da_lo + adv, db_hi + adv, idesc, (c | j) ? 1u : 0u);
                tc_mma_tf32(tmem_d, da_hi + adv, db_lo + adv, idesc, 1u);
                tc_mma_tf32(tmem_d, da_hi + adv, db_hi + adv, idesc, 1u);
              } else {
                tc_mma_tf32(tmem_d, da_hi + adv, db_hi + adv, idesc, (c | j) ? 1u : 0u);
              }
            }
            tc_commit(smem_u32(&a_empty[sa_]));  // A stage reusable once these MMAs have read it
            if (++sa_ == p.sa) {
              sa_ = 0;
              pa ^= 1;
            }
          }
          tc_commit(smem_u32(&b_empty[sb_]));    // weight chunk consumed by all tiles of the super-tile
          if (++sb_ == p.sb) {
            sb_ = 0;
            pb ^= 1;
          }
        }
        tc_commit(smem_u32(&tmem_full[acc]));    // accumulators of the super-tile complete
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue (4 warps; warp w may only touch TMEM lanes 32*(w%4)..+31) =================
    const int quarter = warp & 3;
    for (int i = 0; i < w.n_super; ++i) {
      const int acc = i & 1;
      const int ti = w.tiles_in(p, i);
      const int st = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      mbar_wait(smem_u32(&tmem_full[acc]), static_cast<uint32_t>((i >> 1) & 1));
      tc_fence_after();
      for (int t = 0; t < ti; ++t) {
        const int64_t row = (static_cast<int64_t>(st) * T + t) * kTileM + quarter * 32 + lane;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>((acc * T + t) * n_cta);
        for (int c0 = 0; c0 < n_cta; c0 += 16) {
          uint32_t r[16];
          tc_ld16(taddr + c0, r);
          tc_wait_ld();
          if (row < p.num_out && !(p.debug & 32)) {
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
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
    }
  }

  tc_fence_before();
  __syncthreads();
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
  constexpr int kBatch = 8;                                  // grad_out pieces per thread per batch
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
  uint32_t offA[kPiecesA];
#pragma unroll
  for (int i = 0; i < kPiecesA; ++i)
    offA[i] = static_cast<uint32_t>(pcA >> 3) * (kWgRows * 128) + mn_piece_offset(rA0 + i * kRowStepA, pcA);
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
    for (int i = 0; i < kPiecesA; ++i) split_store(a_hi, a_lo, offA[i], va[i], kSplit);
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
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase[2] = {0, 0};
      // both operands MN-major: bits 15 and 16
      const uint32_t idesc = make_idesc_tf32(kTileM, p.n_pad) | (1u << 15) | (1u << 16);
      const uint32_t blk = kWgRows * 128;  // column-block stride
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int64_t row_begin = static_cast<int64_t>(item / p.groups) * p.rows_per_chunk;
        int64_t row_end = row_begin + p.rows_per_chunk;
        if (row_end > p.num_out) row_end = p.num_out;
        mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase[acc] ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * p.n_pad);
        uint32_t first = 1;
        (void)row_end;
        for (int64_t rb = row_begin; rb < row_begin + p.rows_per_chunk; rb += kWgRows) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
          const uint32_t a_lo = a_hi + a_part;
          const uint32_t g_hi = a_hi + kParts * a_part;
          const uint32_t g_lo = g_hi + g_part;
#pragma unroll
          for (int ks = 0; ks < kWgRows / 8; ++ks) {
            const uint32_t o = ks * 1024;  // 8 rows x 128 B inside every column block
            if (kSplit) {
              tc_mma_tf32(tmem_d, make_desc_mn_sw128(a_lo + o, blk), make_desc_mn_sw128(g_hi + o, blk), idesc, first ? 0u : 1u);
              tc_mma_tf32(tmem_d, make_desc_mn_sw128(a_hi + o, blk), make_desc_mn_sw128(g_lo + o, blk), idesc, 1u);
              tc_mma_tf32(tmem_d, make_desc_mn_sw128(a_hi + o, blk), make_desc_mn_sw128(g_hi + o, blk), idesc, 1u);
            } else {
              tc_mma_tf32(tmem_d, make_desc_mn_sw128(a_hi + o, blk), make_desc_mn_sw128(g_hi + o, blk), idesc, first ? 0u : 1u);
            }
            first = 0;
          }
          tc_commit(smem_u32(&empty_bar[stage]));
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(smem_u32(&tmem_full[acc]));
        acc_phase[acc] ^= 1;
        acc ^= 1;
      }
    }
    __syncwarp();
  } else if (warp >= kLoaderWarp + 1) {
    // ================= epilogue: TMEM -> red.global.add into dW[co][tap][ci] =================
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase[2] = {0, 0};
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int g = item % p.groups;
      const int64_t row_begin = static_cast<int64_t>(item / p.groups) * p.rows_per_chunk;
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase[acc]);
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
      acc_phase[acc] ^= 1;
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

}  // namespace tc
}  // namespace efgb

using namespace efgb;

extern "C" int efgb_spconv_tc_supported(int c_red, int n_out, int taps) { return tc::supported(c_red, n_out, taps) ? 1 : 0; }

extern "C" size_t efgb_spconv_tc_packed_bytes(int taps, int c_red, int n_out, int split) {
  if (!tc::supported(c_red, n_out, taps)) return 0;
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

extern "C" int efgb_spconv_tc_forward_ex(const float* in_feats, int64_t num_in, int c_red, const float* packed,
                                         const float* bias, const int32_t* nbr, int64_t num_out, int taps, int n_out,
                                         int split, int relu, float* out_feats, efgb_stream_t stream_);

extern "C" int efgb_spconv_tc_forward(const float* in_feats, int64_t num_in, int c_red, const float* packed,
                                      const float* bias, const int32_t* nbr, int64_t num_out, int taps, int n_out,
                                      int split, float* out_feats, efgb_stream_t stream_) {
  return efgb_spconv_tc_forward_ex(in_feats, num_in, c_red, packed, bias, nbr, num_out, taps, n_out, split, 0, out_feats,
                                   stream_);
}

extern "C" int efgb_spconv_tc_forward_ex(const float* in_feats, int64_t num_in, int c_red, const float* packed,
                                         const float* bias, const int32_t* nbr, int64_t num_out, int taps, int n_out,
                                         int split, int relu, float* out_feats, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(tc::supported(c_red, n_out, taps), EFGB_EINVAL, "spconv_tc_forward: unsupported shape (c_red=%d n_out=%d taps=%d)",
               c_red, n_out, taps);
  EFGB_REQUIRE(num_in >= 0 && num_out >= 0, EFGB_EINVAL, "spconv_tc_forward: bad sizes");
  EFGB_REQUIRE(nbr != nullptr || (taps == 1 && num_in >= num_out), EFGB_EINVAL,
               "spconv_tc_forward: a null rulebook means identity and needs taps == 1");
  if (num_out == 0) return EFGB_OK;
  EFGB_REQUIRE(packed && out_feats && (in_feats || num_in == 0), EFGB_EINVAL, "spconv_tc_forward: null pointer");
  EFGB_REQUIRE((reinterpret_cast<uintptr_t>(in_feats) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_feats) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(packed) & 15) == 0,
               EFGB_EINVAL, "spconv_tc_forward: feature / weight pointers must be 16-byte aligned");
  tc::Params p;
  p.in = in_feats;
  p.packed = packed;
  p.bias = bias;
  p.relu = relu ? 1 : 0;
  p.nbr = nbr;
  p.out = out_feats;
  p.num_out = num_out;
  p.c_red = c_red;
  p.taps = taps;
  p.n_out = n_out;
  p.chunks = tc::chunks_for(taps, c_red);
  p.num_tiles = static_cast<int>((num_out + tc::kTileM - 1) / tc::kTileM);
  {
    const char* dbg = getenv("EFGB_TC_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  // N split: wide outputs (dense GEMMs) are cut into 256-column slabs over grid.y
  int n_split = n_out > 256 ? n_out / 256 : 1;
  int n_cta = n_out / n_split;
  // super-tile: T row tiles share each weight chunk (T * n_cta fp32 columns per accumulator set, two sets)
  int T = 256 / n_cta;
  if (T > 8) T = 8;
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
  p.tiles_per_super = T;
  p.num_super = (p.num_tiles + T - 1) / T;
  const int parts = split ? 2 : 1;
  const int a_bytes = parts * tc::kTileM * 128;
  const int b_bytes = parts * n_cta * 128;
  p.sb = b_bytes >= 65536 ? 2 : (b_bytes >= 16384 ? 3 : 4);
  const int budget = 227 * 1024 - 2048;
  p.sa = (budget - p.sb * b_bytes) / a_bytes;
  if (p.sa > 6) p.sa = 6;
  EFGB_REQUIRE(p.sa >= 2, EFGB_EINVAL, "spconv_tc_forward: tile does not fit shared memory");
  {
    // EFGB_TC_PRODUCER=async selects the cp.async producers (need >= 3 A stages).  Measured on B200 they are ~20 %
    // SLOWER than the register-staged ones on the sparse layers (96 vs 77 us at C=16, 184 vs 147 us at C=64) and equal
    // on the dense ones: the stage rate is bound by shared-memory bandwidth (SS-mode tcgen05.mma re-reads A and B
    // from shared memory for each of the three split products), and the in-place split adds a read + a write per piece.
    const char* prod = getenv("EFGB_TC_PRODUCER");
    p.async_gather = (p.sa >= 3 && prod && strcmp(prod, "async") == 0) ? 1 : 0;
    p.cred_shift = -1;
    for (int sft = 2; sft < 16; ++sft)
      if ((1 << sft) == c_red) p.cred_shift = sft;
  }
  const size_t smem = 1024 + static_cast<size_t>(p.sa) * a_bytes + static_cast<size_t>(p.sb) * b_bytes +
                      (3 * p.sa + 2 * p.sb + 4) * 8 + 16;
  int gx = kNumSMs / n_split;
  if (gx < 1) gx = 1;
  if (gx > p.num_super) gx = p.num_super;
  const dim3 grid(gx, n_split);
  if (split) {
    static bool configured = false;
    if (!configured) {
      EFGB_CUDA_OK(cudaFuncSetAttribute(tc::spconv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      configured = true;
    }
    tc::spconv_tc_kernel<true><<<grid, tc::kThreads, smem, stream>>>(p);
  } else {
    static bool configured = false;
    if (!configured) {
      EFGB_CUDA_OK(cudaFuncSetAttribute(tc::spconv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      configured = true;
    }
    tc::spconv_tc_kernel<false><<<grid, tc::kThreads, smem, stream>>>(p);
  }
  EFGB_LAUNCH_OK("spconv_tc_kernel");
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
    static bool configured = false;
    if (!configured) {
      EFGB_CUDA_OK(cudaFuncSetAttribute(tc::spconv_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      configured = true;
    }
    tc::spconv_wgrad_tc_kernel<true><<<grid, tc::kThreads, smem, stream>>>(p, stages);
  } else {
    static bool configured = false;
    if (!configured) {
      EFGB_CUDA_OK(cudaFuncSetAttribute(tc::spconv_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      configured = true;
    }
    tc::spconv_wgrad_tc_kernel<false><<<grid, tc::kThreads, smem, stream>>>(p, stages);
  }
  EFGB_LAUNCH_OK("spconv_wgrad_tc_kernel");
  return EFGB_OK;
}
