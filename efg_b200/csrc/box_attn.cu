// Box attention (multi-scale bilinear sample-and-weight), forward and backward, sm_100a.
//
// Same math as the reference kernels (efg/operators/src/box_attn/box_attn_kernel.cuh:35-98 fwd
// bilinear, :101-190 bwd bilinear, :275-351 fwd loop) but a different mapping: the reference runs
// one thread per output scalar (every thread of a head re-reads the same loc/attn scalars and
// recomputes the same bilinear weights) and, in backward, one 32-thread block per (b,q,h) with a
// thread-0 serial shared-memory reduction.  Here one WARP owns one (b,q,h): lanes are the head's
// channels, so each bilinear corner is one coalesced 128-byte line, loc/attn are loaded once per
// warp (lane p holds point p) and broadcast by shuffle, grad_loc/grad_attn are reduced with
// shuffles, and grad_value is accumulated with one coalesced RED per corner.
#include <stdlib.h>

#include "common.cuh"

namespace efgb {

struct Corner {
  int h_low, w_low;
  float w1, w2, w3, w4;   // bilinear weights: (hl,wl) (hl,wh) (hh,wl) (hh,wh)
  float lh, lw, hh, hw;
  bool ok1, ok2, ok3, ok4;
};

__device__ __forceinline__ bool make_corner(float x, float y, int Hl, int Wl, Corner* c, float* h_im_out, float* w_im_out) {
  // h_im = loc_y * H - 0.5 with the product rounded before the subtraction (box_attn_kernel.cuh:325-326)
  const float h_im = __fsub_rn(__fmul_rn(y, static_cast<float>(Hl)), 0.5f);
  const float w_im = __fsub_rn(__fmul_rn(x, static_cast<float>(Wl)), 0.5f);
  *h_im_out = h_im;
  *w_im_out = w_im;
  if (!(h_im > -1.f && w_im > -1.f && h_im < static_cast<float>(Hl) && w_im < static_cast<float>(Wl))) return false;
  const float hf = floorf(h_im), wf = floorf(w_im);
  c->h_low = static_cast<int>(hf);
  c->w_low = static_cast<int>(wf);
  c->lh = h_im - hf;
  c->lw = w_im - wf;
  c->hh = 1.f - c->lh;
  c->hw = 1.f - c->lw;
  c->w1 = c->hh * c->hw;
  c->w2 = c->hh * c->lw;
  c->w3 = c->lh * c->hw;
  c->w4 = c->lh * c->lw;
  const bool hl = c->h_low >= 0, hh = c->h_low + 1 <= Hl - 1;
  const bool wl = c->w_low >= 0, wh = c->w_low + 1 <= Wl - 1;
  c->ok1 = hl && wl;
  c->ok2 = hl && wh;
  c->ok3 = hh && wl;
  c->ok4 = hh && wh;
  return true;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// NC = ceil(head_dim / 32) channel slots per lane.
template <int NC>
__global__ void __launch_bounds__(256)
box_attn_fwd_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ level_start, const float* __restrict__ loc,
                    const float* __restrict__ attn, int64_t num_warps, int len_value, int num_heads, int head_dim,
                    int num_levels, int len_query, int num_points, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (idx >= num_warps) return;
  const int h = static_cast<int>(idx % num_heads);
  const int64_t b = idx / (static_cast<int64_t>(num_heads) * len_query);
  const int64_t pix_stride = static_cast<int64_t>(num_heads) * head_dim;
  const float* loc_q = loc + idx * num_levels * num_points * 2;
  const float* attn_q = attn + idx * num_levels * num_points;

  float acc[NC];
#pragma unroll
  for (int n = 0; n < NC; ++n) acc[n] = 0.f;

  for (int l = 0; l < num_levels; ++l) {
    const int Hl = static_cast<int>(shapes[2 * l]), Wl = static_cast<int>(shapes[2 * l + 1]);
    const float* vbase = value + (b * len_value + level_start[l]) * pix_stride + static_cast<int64_t>(h) * head_dim;
    for (int p0 = 0; p0 < num_points; p0 += 32) {
      const int p = p0 + lane;
      float lx = 0.f, ly = 0.f, wt = 0.f;
      if (p < num_points) {
        const float2 xy = *reinterpret_cast<const float2*>(loc_q + (l * num_points + p) * 2);
        lx = xy.x;
        ly = xy.y;
        wt = attn_q[l * num_points + p];
      }
      const int np = min(32, num_points - p0);
      for (int pp = 0; pp < np; ++pp) {
        const float x = __shfl_sync(0xffffffffu, lx, pp);
        const float y = __shfl_sync(0xffffffffu, ly, pp);
        const float a = __shfl_sync(0xffffffffu, wt, pp);
        Corner c;
        float h_im, w_im;
        if (!make_corner(x, y, Hl, Wl, &c, &h_im, &w_im)) continue;
        const float* p1 = vbase + (static_cast<int64_t>(c.h_low) * Wl + c.w_low) * pix_stride;
        const float* p2 = p1 + pix_stride;
        const float* p3 = p1 + static_cast<int64_t>(Wl) * pix_stride;
        const float* p4 = p3 + pix_stride;
#pragma unroll
        for (int n = 0; n < NC; ++n) {
          const int ch = lane + 32 * n;
          if (ch < head_dim) {
            const float v1 = c.ok1 ? __ldg(p1 + ch) : 0.f;
            const float v2 = c.ok2 ? __ldg(p2 + ch) : 0.f;
            const float v3 = c.ok3 ? __ldg(p3 + ch) : 0.f;
            const float v4 = c.ok4 ? __ldg(p4 + ch) : 0.f;
            const float val = c.w1 * v1 + c.w2 * v2 + c.w3 * v3 + c.w4 * v4;
            acc[n] += val * a;
          }
        }
      }
    }
  }
  float* o = out + idx * head_dim;
#pragma unroll
  for (int n = 0; n < NC; ++n) {
    const int ch = lane + 32 * n;
    if (ch < head_dim) o[ch] = acc[n];
  }
}

template <int NC>
__global__ void __launch_bounds__(256)
box_attn_bwd_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ level_start, const float* __restrict__ loc,
                    const float* __restrict__ attn, const float* __restrict__ grad_out, int64_t num_warps,
                    int len_value, int num_heads, int head_dim, int num_levels, int len_query, int num_points,
                    float* __restrict__ grad_value, float* __restrict__ grad_loc, float* __restrict__ grad_attn) {
  const int lane = threadIdx.x & 31;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (idx >= num_warps) return;
  const int h = static_cast<int>(idx % num_heads);
  const int64_t b = idx / (static_cast<int64_t>(num_heads) * len_query);
  const int64_t pix_stride = static_cast<int64_t>(num_heads) * head_dim;
  const int64_t lp = static_cast<int64_t>(num_levels) * num_points;
  const float* loc_q = loc + idx * lp * 2;
  const float* attn_q = attn + idx * lp;
  float* gloc_q = grad_loc + idx * lp * 2;
  float* gattn_q = grad_attn + idx * lp;

  float tg[NC];
#pragma unroll
  for (int n = 0; n < NC; ++n) {
    const int ch = lane + 32 * n;
    tg[n] = ch < head_dim ? grad_out[idx * head_dim + ch] : 0.f;
  }

  for (int l = 0; l < num_levels; ++l) {
    const int Hl = static_cast<int>(shapes[2 * l]), Wl = static_cast<int>(shapes[2 * l + 1]);
    const int64_t voff = (b * len_value + level_start[l]) * pix_stride + static_cast<int64_t>(h) * head_dim;
    const float* vbase = value + voff;
    float* gbase = grad_value + voff;
    for (int p0 = 0; p0 < num_points; p0 += 32) {
      const int p = p0 + lane;
      float lx = 0.f, ly = 0.f, wt = 0.f;
      if (p < num_points) {
        const float2 xy = *reinterpret_cast<const float2*>(loc_q + (l * num_points + p) * 2);
        lx = xy.x;
        ly = xy.y;
        wt = attn_q[l * num_points + p];
      }
      // per-point results gathered back to lane p for one coalesced store
      float out_gx = 0.f, out_gy = 0.f, out_ga = 0.f;
      const int np = min(32, num_points - p0);
      for (int pp = 0; pp < np; ++pp) {
        const float x = __shfl_sync(0xffffffffu, lx, pp);
        const float y = __shfl_sync(0xffffffffu, ly, pp);
        const float a = __shfl_sync(0xffffffffu, wt, pp);
        Corner c;
        float h_im, w_im;
        if (!make_corner(x, y, Hl, Wl, &c, &h_im, &w_im)) continue;  // warp-uniform
        const int64_t o1 = (static_cast<int64_t>(c.h_low) * Wl + c.w_low) * pix_stride;
        const int64_t o2 = o1 + pix_stride;
        const int64_t o3 = o1 + static_cast<int64_t>(Wl) * pix_stride;
        const int64_t o4 = o3 + pix_stride;
        float s_val = 0.f, s_gw = 0.f, s_gh = 0.f;
#pragma unroll
        for (int n = 0; n < NC; ++n) {
          const int ch = lane + 32 * n;
          if (ch < head_dim) {
            const float tgv = tg[n] * a;  // top_grad * attn_weight
            float gh = 0.f, gw = 0.f;
            float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
            if (c.ok1) {
              v1 = __ldg(vbase + o1 + ch);
              gh -= c.hw * v1;
              gw -= c.hh * v1;
              atomicAdd(gbase + o1 + ch, c.w1 * tgv);
            }
            if (c.ok2) {
              v2 = __ldg(vbase + o2 + ch);
              gh -= c.lw * v2;
              gw += c.hh * v2;
              atomicAdd(gbase + o2 + ch, c.w2 * tgv);
            }
            if (c.ok3) {
              v3 = __ldg(vbase + o3 + ch);
              gh += c.hw * v3;
              gw -= c.lh * v3;
              atomicAdd(gbase + o3 + ch, c.w3 * tgv);
            }
            if (c.ok4) {
              v4 = __ldg(vbase + o4 + ch);
              gh += c.lw * v4;
              gw += c.lh * v4;
              atomicAdd(gbase + o4 + ch, c.w4 * tgv);
            }
            const float val = c.w1 * v1 + c.w2 * v2 + c.w3 * v3 + c.w4 * v4;
            s_val += tg[n] * val;
            s_gw += static_cast<float>(Wl) * gw * tgv;
            s_gh += static_cast<float>(Hl) * gh * tgv;
          }
        }
        s_val = warp_sum(s_val);
        s_gw = warp_sum(s_gw);
        s_gh = warp_sum(s_gh);
        if (lane == pp) {
          out_ga = s_val;
          out_gx = s_gw;
          out_gy = s_gh;
        }
      }
      if (p < num_points) {
        *reinterpret_cast<float2*>(gloc_q + (l * num_points + p) * 2) = make_float2(out_gx, out_gy);
        gattn_q[l * num_points + p] = out_ga;
      }
    }
  }
}

// =================================================================================================
// head_dim == 32 specialisation ("tile" kernels) — the Voxel-DETR / ConQueR geometry (8 heads x 32 channels).
//
// The generic kernels above are issue-bound (every lane recomputes every tap's corner arithmetic) and, in the
// encoder, L2-gather-bound (the 8 warps of a CTA are the 8 heads of ONE query, so nothing is shared in L1).
// Here:
//   * a CTA owns one head of an 8 x RY tile of queries (warp = query column, RY rows visited in turn).  When the
//     caller says that consecutive queries form a row-major grid of width `qw` (encoder self-attention: the
//     queries ARE the BEV cells) the tile is a 2-D patch whose sampling footprints overlap almost entirely, so
//     the bilinear corner lines are served by L1 instead of L2.  `qw` is a locality hint only: every query is
//     visited exactly once for any value of it.
//   * lane p computes the corner arithmetic of tap p ONCE and parks (offset, attn * bilinear weight) for its four
//     corners in shared memory;
//   * in the tap loop the warp is split into 4 corner groups of 8 lanes, each lane owning 4 channels: ONE
//     LDG.128 per tap fetches the four 128-byte corner lines, and (backward) ONE RED.128 per tap accumulates
//     grad_value;
//   * backward: the per-tap dot products <value_corner, grad_out> are reduced inside each 8-lane group with a
//     transposing butterfly (7 shuffles per 8 taps) and handed back to the tap's owner lane through shared
//     memory, which forms grad_attn and grad_loc from the four corner dots.
// =================================================================================================
constexpr int kTileWarps = 8;

// ---- "where to attend" computed inside the attention kernels (fused variant, one level, <= 32 points) ----------------
// Same arithmetic, in the same order, as box_grid_softmax_kernel below (Box3dAttention._where_to_attend + softmax,
// VD/modules/box_attention.py:62-95,105-108): the sampling locations and attention weights never go through HBM —
// the unfused pair writes and re-reads loc [B,LQ,H,P,2] + attn [B,LQ,H,P] (170 MB per encoder layer and direction).
struct FusedArgs {
  const float* offsets;   // [B*LQ rows, stride ld_offsets] : (h, nv) box offsets
  const float* logits;    // [B*LQ rows, stride ld_logits]  : (h, p) attention logits
  const float* ref;       // [B*LQ, 7]
  const float* kidx;      // [P, 2]
  int64_t ld_logits, ld_offsets;
  int nv;                 // 4, or 5 with rotation
  float* g_offsets;       // backward outputs, same strides
  float* g_logits;
};

struct FusedPoint {
  float a;                // softmax probability of this lane's point
  float kx, ky;           // kernel offset of the point
  float gx, gy;           // kidx * relu(size)
  float cs, sn;
  float w, l, sw, sl;     // reference size, box size before the relu
};

constexpr float kTwoPiF = 6.283185307179586f;

__device__ __forceinline__ float2 fused_where(const FusedArgs& fa, int64_t bq, int h, int lane, int num_points, FusedPoint* fp) {
  const float* r = fa.ref + bq * 7;
  const float cx = __ldg(r + 0), cy = __ldg(r + 1), w = __ldg(r + 3), l = __ldg(r + 4), ref_angle = __ldg(r + 6);
  const float* lg = fa.logits + bq * fa.ld_logits + static_cast<int64_t>(h) * num_points;
  const float x = lane < num_points ? __ldg(lg + lane) : -INFINITY;
  float m = x;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  const float e = lane < num_points ? expf(x - m) : 0.f;
  const float inv = 1.f / warp_sum(e);
  fp->a = e * inv;
  const float* off = fa.offsets + bq * fa.ld_offsets + static_cast<int64_t>(h) * fa.nv;
  const float o0 = __ldg(off + 0), o1 = __ldg(off + 1), o2 = __ldg(off + 2), o3 = __ldg(off + 3);
  const float bx = cx + o0 / 8.f * w, by = cy + o1 / 8.f * l;
  fp->w = w;
  fp->l = l;
  fp->sw = w + o2 / 8.f * w;
  fp->sl = l + o3 / 8.f * l;
  const float angle = fa.nv == 5 ? (ref_angle + __ldg(off + 4) / 16.f) * kTwoPiF : ref_angle;
  sincosf(angle, &fp->sn, &fp->cs);
  const float rw = fmaxf(fp->sw, 0.f), rl = fmaxf(fp->sl, 0.f);
  fp->kx = fp->ky = 0.f;
  if (lane < num_points) {
    fp->kx = __ldg(fa.kidx + 2 * lane);
    fp->ky = __ldg(fa.kidx + 2 * lane + 1);
  }
  fp->gx = fp->kx * rw;
  fp->gy = fp->ky * rl;
  return make_float2(bx + (fp->gx * fp->cs - fp->gy * fp->sn), by + (fp->gx * fp->sn + fp->gy * fp->cs));
}

// Backward of fused_where for one (b, q, h): (g_xy, g_a) of the lane's point -> gradients of the logits and offsets rows.
__device__ __forceinline__ void fused_where_backward(const FusedArgs& fa, int64_t bq, int h, int lane, int num_points,
                                                     const FusedPoint& fp, float gxy_x, float gxy_y, float ga) {
  const float dot = warp_sum(fp.a * ga);
  if (lane < num_points)
    fa.g_logits[bq * fa.ld_logits + static_cast<int64_t>(h) * num_points + lane] = fp.a * (ga - dot);
  const float ggx = gxy_x * fp.cs + gxy_y * fp.sn;    // d loss / d gx
  const float ggy = -gxy_x * fp.sn + gxy_y * fp.cs;   // d loss / d gy
  const float a_cx = warp_sum(gxy_x);
  const float a_cy = warp_sum(gxy_y);
  const float a_sw = warp_sum(ggx * fp.kx);
  const float a_sl = warp_sum(ggy * fp.ky);
  const float a_ang = warp_sum(gxy_x * (-fp.gx * fp.sn - fp.gy * fp.cs) + gxy_y * (fp.gx * fp.cs - fp.gy * fp.sn));
  if (lane == 0) {
    float* go = fa.g_offsets + bq * fa.ld_offsets + static_cast<int64_t>(h) * fa.nv;
    go[0] = a_cx * fp.w / 8.f;
    go[1] = a_cy * fp.l / 8.f;
    go[2] = (fp.sw > 0.f ? a_sw : 0.f) * fp.w / 8.f;
    go[3] = (fp.sl > 0.f ? a_sl : 0.f) * fp.l / 8.f;
    if (fa.nv == 5) go[4] = a_ang * kTwoPiF / 16.f;
  }
}

struct TileGeom {
  int qw, qh;   // query grid (qw = 8, qh = ceil(LQ / 8) when no grid hint was given)
  int tx, ty;   // tiles along x / y
  int ry;       // query rows per tile
};

__device__ __forceinline__ void tile_fill_taps(int2* my, const int lane, const bool active, const float2 xy, const float a,
                                               const int Hl, const int Wl, const int pix_stride, Corner* c, bool* made) {
  int o1 = -1, o2 = -1, o3 = -1, o4 = -1;
  c->w1 = c->w2 = c->w3 = c->w4 = 0.f;
  c->lh = c->lw = c->hh = c->hw = 0.f;
  *made = false;
  if (active) {
    float h_im, w_im;
    if (make_corner(xy.x, xy.y, Hl, Wl, c, &h_im, &w_im)) {
      *made = true;
      const int base = (c->h_low * Wl + c->w_low) * pix_stride;
      if (c->ok1) o1 = base;
      if (c->ok2) o2 = base + pix_stride;
      if (c->ok3) o3 = base + Wl * pix_stride;
      if (c->ok4) o4 = base + (Wl + 1) * pix_stride;
    }
    int4* dst = reinterpret_cast<int4*>(my + lane * 4);
    dst[0] = make_int4(o1, __float_as_int(c->w1 * a), o2, __float_as_int(c->w2 * a));
    dst[1] = make_int4(o3, __float_as_int(c->w3 * a), o4, __float_as_int(c->w4 * a));
  }
}

template <bool kFused>
__global__ void __launch_bounds__(kTileWarps * 32)
box_attn_fwd_tile_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                         const int64_t* __restrict__ level_start, const float* __restrict__ loc,
                         const float* __restrict__ attn, const FusedArgs fa, const TileGeom g, int len_value, int num_heads,
                         int num_levels, int len_query, int num_points, float* __restrict__ out) {
  __shared__ int2 taps[kTileWarps][32 * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = lane >> 3;        // corner group
  const int cq = (lane & 7) * 4;   // first of this lane's 4 channels
  int64_t blk = blockIdx.x;
  const int tx = static_cast<int>(blk % g.tx);
  blk /= g.tx;
  const int ty = static_cast<int>(blk % g.ty);
  blk /= g.ty;
  const int h = static_cast<int>(blk % num_heads);
  const int64_t b = blk / num_heads;
  const int x = tx * kTileWarps + warp;
  if (x >= g.qw) return;  // no block-wide barrier below
  const int pix_stride = num_heads * 32;
  int2* my = taps[warp];

  for (int r = 0; r < g.ry; ++r) {
    const int y = ty * g.ry + r;
    const int64_t q = static_cast<int64_t>(y) * g.qw + x;
    if (y >= g.qh || q >= len_query) break;
    const int64_t idx = (b * len_query + q) * num_heads + h;
    const float* loc_q = loc + idx * num_levels * num_points * 2;
    const float* attn_q = attn + idx * num_levels * num_points;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < num_levels; ++l) {
      const int Hl = static_cast<int>(shapes[2 * l]), Wl = static_cast<int>(shapes[2 * l + 1]);
      const float* vbase = value + (b * len_value + level_start[l]) * pix_stride + h * 32 + cq;
      for (int p0 = 0; p0 < num_points; p0 += 32) {
        const int p = p0 + lane;
        const int np = min(32, num_points - p0);
        float2 xy = make_float2(0.f, 0.f);
        float a = 0.f;
        if constexpr (kFused) {   // one level, num_points <= 32: this loop body runs once per (q, h)
          FusedPoint fp;
          xy = fused_where(fa, b * len_query + q, h, lane, num_points, &fp);
          a = fp.a;
        } else if (p < num_points) {
          xy = __ldg(reinterpret_cast<const float2*>(loc_q + (l * num_points + p) * 2));
          a = __ldg(attn_q + l * num_points + p);
        }
        Corner c;
        bool made;
        __syncwarp();  // readers of the previous chunk are done with `my`
        tile_fill_taps(my, lane, p < num_points, xy, a, Hl, Wl, pix_stride, &c, &made);
        __syncwarp();
        for (int t0 = 0; t0 < np; t0 += 5) {  // 5 independent corner-line loads in flight per lane (P = 25 = 5 x 5)
          int2 e[5];
          float4 v[5];
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            e[j] = make_int2(-1, 0);
            if (t0 + j < np) e[j] = my[(t0 + j) * 4 + cg];  // warp-uniform guard
          }
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e[j].x >= 0) v[j] = __ldg(reinterpret_cast<const float4*>(vbase + e[j].x));
          }
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const float w = e[j].x >= 0 ? __int_as_float(e[j].y) : 0.f;
            acc.x = fmaf(w, v[j].x, acc.x);
            acc.y = fmaf(w, v[j].y, acc.y);
            acc.z = fmaf(w, v[j].z, acc.z);
            acc.w = fmaf(w, v[j].w, acc.w);
          }
        }
      }
    }
#pragma unroll
    for (int d = 8; d <= 16; d <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, d);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, d);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, d);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, d);
    }
    if (lane < 8) *reinterpret_cast<float4*>(out + idx * 32 + cq) = acc;
  }
}

template <bool kFused>
__global__ void __launch_bounds__(kTileWarps * 32, kFused ? 3 : 0)   // fused: three CTAs per SM (104 registers uncapped, occupancy 24 %)
box_attn_bwd_tile_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                         const int64_t* __restrict__ level_start, const float* __restrict__ loc,
                         const float* __restrict__ attn, const float* __restrict__ grad_out, const FusedArgs fa,
                         const TileGeom g, int len_value, int num_heads, int num_levels, int len_query, int num_points,
                         float* __restrict__ grad_value, float* __restrict__ grad_loc, float* __restrict__ grad_attn) {
  __shared__ int2 taps[kTileWarps][32 * 4];
  __shared__ __align__(16) float dots[kTileWarps][32 * 4];  // [tap][corner] = <value_corner, grad_out>
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = lane >> 3;
  const int cq = (lane & 7) * 4;
  int64_t blk = blockIdx.x;
  const int tx = static_cast<int>(blk % g.tx);
  blk /= g.tx;
  const int ty = static_cast<int>(blk % g.ty);
  blk /= g.ty;
  const int h = static_cast<int>(blk % num_heads);
  const int64_t b = blk / num_heads;
  const int x = tx * kTileWarps + warp;
  if (x >= g.qw) return;
  const int pix_stride = num_heads * 32;
  int2* my = taps[warp];
  float* mydots = dots[warp];
  const bool b2 = (lane & 4) != 0, b1 = (lane & 2) != 0, b0 = (lane & 1) != 0;

  for (int r = 0; r < g.ry; ++r) {
    const int y = ty * g.ry + r;
    const int64_t q = static_cast<int64_t>(y) * g.qw + x;
    if (y >= g.qh || q >= len_query) break;
    const int64_t idx = (b * len_query + q) * num_heads + h;
    const int64_t lp = static_cast<int64_t>(num_levels) * num_points;
    const float* loc_q = loc + idx * lp * 2;
    const float* attn_q = attn + idx * lp;
    const float4 tg = __ldg(reinterpret_cast<const float4*>(grad_out + idx * 32 + cq));
    for (int l = 0; l < num_levels; ++l) {
      const int Hl = static_cast<int>(shapes[2 * l]), Wl = static_cast<int>(shapes[2 * l + 1]);
      const int64_t voff = (b * len_value + level_start[l]) * pix_stride + h * 32 + cq;
      const float* vbase = value + voff;
      float* gbase = grad_value + voff;
      for (int p0 = 0; p0 < num_points; p0 += 32) {
        const int p = p0 + lane;
        const int np = min(32, num_points - p0);
        float2 xy = make_float2(0.f, 0.f);
        float a = 0.f;
        if constexpr (kFused) {
          FusedPoint fp;   // not kept across the tap loop (11 live registers: 106 per thread, occupancy 24 % instead of 37 %);
          xy = fused_where(fa, b * len_query + q, h, lane, num_points, &fp);   // recomputed for the backward below
          a = fp.a;
        } else if (p < num_points) {
          xy = __ldg(reinterpret_cast<const float2*>(loc_q + (l * num_points + p) * 2));
          a = __ldg(attn_q + l * num_points + p);
        }
        Corner c;
        bool made;
        __syncwarp();  // owners of the previous chunk have read `mydots`, readers are done with `my`
        tile_fill_taps(my, lane, p < num_points, xy, a, Hl, Wl, pix_stride, &c, &made);
        __syncwarp();
        for (int t0 = 0; t0 < np; t0 += 8) {
          float d[8];
#pragma unroll
          for (int j0 = 0; j0 < 8; j0 += 4) {  // 4 independent corner-line loads in flight per lane
            int2 e[4];
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              e[j] = make_int2(-1, 0);
              if (t0 + j0 + j < np) e[j] = my[(t0 + j0 + j) * 4 + cg];  // warp-uniform guard
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (e[j].x >= 0) v[j] = __ldg(reinterpret_cast<const float4*>(vbase + e[j].x));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              d[j0 + j] = v[j].x * tg.x + v[j].y * tg.y + v[j].z * tg.z + v[j].w * tg.w;
              if (e[j].x >= 0) {
                const float w = __int_as_float(e[j].y);
                atomicAdd(reinterpret_cast<float4*>(gbase + e[j].x), make_float4(w * tg.x, w * tg.y, w * tg.z, w * tg.w));
              }
            }
          }
          // transposing butterfly over the 8 lanes of the corner group: lane j ends with the sum of d[j]
          float e4[4], f2[2];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float send = b2 ? d[i] : d[i + 4];
            const float keep = b2 ? d[i + 4] : d[i];
            e4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float send = b1 ? e4[i] : e4[i + 2];
            const float keep = b1 ? e4[i + 2] : e4[i];
            f2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
          const float send = b0 ? f2[0] : f2[1];
          const float keep = b0 ? f2[1] : f2[0];
          const float tot = keep + __shfl_xor_sync(0xffffffffu, send, 1);
          mydots[(t0 + (lane & 7)) * 4 + cg] = tot;
        }
        __syncwarp();
        {
          float gx = 0.f, gy = 0.f, ga = 0.f;
          if (p < num_points && made) {
            const float4 D = *reinterpret_cast<const float4*>(mydots + lane * 4);
            ga = c.w1 * D.x + c.w2 * D.y + c.w3 * D.z + c.w4 * D.w;
            const float gw = c.hh * (D.y - D.x) + c.lh * (D.w - D.z);
            const float gh = c.hw * (D.z - D.x) + c.lw * (D.w - D.y);
            gx = static_cast<float>(Wl) * gw * a;
            gy = static_cast<float>(Hl) * gh * a;
          }
          if constexpr (kFused) {
            FusedPoint fp;
            fused_where(fa, b * len_query + q, h, lane, num_points, &fp);
            fused_where_backward(fa, b * len_query + q, h, lane, num_points, fp, gx, gy, ga);
          } else if (p < num_points) {
            *reinterpret_cast<float2*>(grad_loc + (idx * lp + l * num_points + p) * 2) = make_float2(gx, gy);
            grad_attn[idx * lp + l * num_points + p] = ga;
          }
        }
      }
    }
  }
}

// Tile geometry for `nw` = batch * len_query * num_heads warps of work.
static TileGeom tile_geom(int len_query, int query_grid_w, int64_t nw) {
  TileGeom g;
  if (query_grid_w > 0 && query_grid_w <= len_query) {
    g.qw = query_grid_w;
  } else {
    g.qw = kTileWarps;
  }
  g.qh = (len_query + g.qw - 1) / g.qw;
  // tall tiles only while at least ~6 CTAs per SM remain
  const int64_t rows = nw / (static_cast<int64_t>(kTileWarps) * kNumSMs * 6);
  g.ry = rows >= 8 ? 8 : rows >= 4 ? 4 : rows >= 2 ? 2 : 1;
  g.tx = (g.qw + kTileWarps - 1) / kTileWarps;
  g.ty = (g.qh + g.ry - 1) / g.ry;
  return g;
}

static bool use_tile_kernels(int len_value, int num_heads, int head_dim) {
  if (head_dim != 32) return false;
  if (static_cast<int64_t>(len_value) * num_heads * 32 >= (1ll << 31)) return false;  // 32-bit tap offsets
  static const bool generic = [] {   // read once per process (A/B measurements against the generic kernels)
    const char* e = getenv("EFGB_BOX_ATTN");
    return e && strcmp(e, "generic") == 0;
  }();
  return !generic;
}

static int check_args(int batch, int len_value, int num_heads, int head_dim, int num_levels, int len_query,
                      int num_points, const char* who) {
  EFGB_REQUIRE(batch >= 0 && len_value >= 0 && num_heads >= 1 && head_dim >= 1 && num_levels >= 1 && len_query >= 0 &&
                   num_points >= 1,
               EFGB_EINVAL, "%s: bad shape", who);
  EFGB_REQUIRE(head_dim <= 128, EFGB_EINVAL, "%s: head_dim %d > 128 is not supported", who, head_dim);
  return EFGB_OK;
}

}  // namespace efgb

using namespace efgb;

extern "C" int efgb_box_attn_forward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start,
                                     const float* loc, const float* attn, int batch, int len_value, int num_heads,
                                     int head_dim, int num_levels, int len_query, int num_points, int query_grid_w,
                                     float* out, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  int rc = check_args(batch, len_value, num_heads, head_dim, num_levels, len_query, num_points, "box_attn_forward");
  if (rc != EFGB_OK) return rc;
  const int64_t nw = static_cast<int64_t>(batch) * len_query * num_heads;
  if (nw == 0) return EFGB_OK;
  EFGB_REQUIRE(value && spatial_shapes && level_start && loc && attn && out, EFGB_EINVAL, "box_attn_forward: null pointer");
  if (use_tile_kernels(len_value, num_heads, head_dim)) {
    const TileGeom g = tile_geom(len_query, query_grid_w, nw);
    const int64_t blocks = static_cast<int64_t>(batch) * num_heads * g.tx * g.ty;
    EFGB_REQUIRE(blocks < (1ll << 31), EFGB_EINVAL, "box_attn_forward: too many query tiles");
    box_attn_fwd_tile_kernel<false><<<static_cast<unsigned>(blocks), kTileWarps * 32, 0, stream>>>(
        value, spatial_shapes, level_start, loc, attn, FusedArgs{}, g, len_value, num_heads, num_levels, len_query, num_points,
        out);
    EFGB_LAUNCH_OK("box_attn_fwd_tile_kernel");
    return EFGB_OK;
  }
  const unsigned nb = static_cast<unsigned>((nw + 7) / 8);
  const int nc = (head_dim + 31) / 32;
#define EFGB_FWD(NC)                                                                                              \
  box_attn_fwd_kernel<NC><<<nb, 256, 0, stream>>>(value, spatial_shapes, level_start, loc, attn, nw, len_value,   \
                                                  num_heads, head_dim, num_levels, len_query, num_points, out)
  if (nc == 1) EFGB_FWD(1);
  else if (nc == 2) EFGB_FWD(2);
  else EFGB_FWD(4);
#undef EFGB_FWD
  EFGB_LAUNCH_OK("box_attn_fwd_kernel");
  return EFGB_OK;
}

extern "C" int efgb_box_attn_backward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start,
                                      const float* loc, const float* attn, const float* grad_out, int batch,
                                      int len_value, int num_heads, int head_dim, int num_levels, int len_query,
                                      int num_points, int query_grid_w, float* grad_value, float* grad_loc,
                                      float* grad_attn, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  int rc = check_args(batch, len_value, num_heads, head_dim, num_levels, len_query, num_points, "box_attn_backward");
  if (rc != EFGB_OK) return rc;
  const size_t vbytes = static_cast<size_t>(batch) * len_value * num_heads * head_dim * sizeof(float);
  if (vbytes) {
    EFGB_REQUIRE(grad_value, EFGB_EINVAL, "box_attn_backward: null grad_value");
    EFGB_CUDA_OK(cudaMemsetAsync(grad_value, 0, vbytes, stream));
  }
  const int64_t nw = static_cast<int64_t>(batch) * len_query * num_heads;
  if (nw == 0) return EFGB_OK;
  EFGB_REQUIRE(value && spatial_shapes && level_start && loc && attn && grad_out && grad_loc && grad_attn, EFGB_EINVAL,
               "box_attn_backward: null pointer");
  if (use_tile_kernels(len_value, num_heads, head_dim)) {
    const TileGeom g = tile_geom(len_query, query_grid_w, nw);
    const int64_t blocks = static_cast<int64_t>(batch) * num_heads * g.tx * g.ty;
    EFGB_REQUIRE(blocks < (1ll << 31), EFGB_EINVAL, "box_attn_backward: too many query tiles");
    box_attn_bwd_tile_kernel<false><<<static_cast<unsigned>(blocks), kTileWarps * 32, 0, stream>>>(
        value, spatial_shapes, level_start, loc, attn, grad_out, FusedArgs{}, g, len_value, num_heads, num_levels, len_query,
        num_points, grad_value, grad_loc, grad_attn);
    EFGB_LAUNCH_OK("box_attn_bwd_tile_kernel");
    return EFGB_OK;
  }
  const unsigned nb = static_cast<unsigned>((nw + 7) / 8);
  const int nc = (head_dim + 31) / 32;
#define EFGB_BWD(NC)                                                                                               \
  box_attn_bwd_kernel<NC><<<nb, 256, 0, stream>>>(value, spatial_shapes, level_start, loc, attn, grad_out, nw,     \
                                                  len_value, num_heads, head_dim, num_levels, len_query, num_points, \
                                                  grad_value, grad_loc, grad_attn)
  if (nc == 1) EFGB_BWD(1);
  else if (nc == 2) EFGB_BWD(2);
  else EFGB_BWD(4);
#undef EFGB_BWD
  EFGB_LAUNCH_OK("box_attn_bwd_kernel");
  return EFGB_OK;
}

// Fused variant: sampling grid + softmax computed inside the tile kernels (one level, head_dim 32, <= 32 points).
extern "C" int efgb_box_attn_fused_supported(int head_dim, int num_levels, int num_points, int num_variables) {
  return (head_dim == 32 && num_levels == 1 && num_points >= 1 && num_points <= 32 && (num_variables == 4 || num_variables == 5)) ? 1 : 0;
}

extern "C" int efgb_box_attn_fused_forward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start,
                                           const float* offsets, const float* logits, const float* ref_windows,
                                           const float* kernel_indices, int batch, int len_value, int num_heads, int len_query,
                                           int num_points, int num_variables, int64_t logits_row_stride,
                                           int64_t offsets_row_stride, int query_grid_w, float* out, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  int rc = check_args(batch, len_value, num_heads, 32, 1, len_query, num_points, "box_attn_fused_forward");
  if (rc != EFGB_OK) return rc;
  EFGB_REQUIRE(efgb_box_attn_fused_supported(32, 1, num_points, num_variables) &&
                   static_cast<int64_t>(len_value) * num_heads * 32 < (1ll << 31),
               EFGB_EINVAL, "box_attn_fused_forward: unsupported shape");
  const int64_t nw = static_cast<int64_t>(batch) * len_query * num_heads;
  if (nw == 0) return EFGB_OK;
  EFGB_REQUIRE(value && spatial_shapes && level_start && offsets && logits && ref_windows && kernel_indices && out, EFGB_EINVAL,
               "box_attn_fused_forward: null pointer");
  FusedArgs fa{};
  fa.offsets = offsets;
  fa.logits = logits;
  fa.ref = ref_windows;
  fa.kidx = kernel_indices;
  fa.ld_logits = logits_row_stride > 0 ? logits_row_stride : static_cast<int64_t>(num_heads) * num_points;
  fa.ld_offsets = offsets_row_stride > 0 ? offsets_row_stride : static_cast<int64_t>(num_heads) * num_variables;
  fa.nv = num_variables;
  const TileGeom g = tile_geom(len_query, query_grid_w, nw);
  const int64_t blocks = static_cast<int64_t>(batch) * num_heads * g.tx * g.ty;
  EFGB_REQUIRE(blocks < (1ll << 31), EFGB_EINVAL, "box_attn_fused_forward: too many query tiles");
  box_attn_fwd_tile_kernel<true><<<static_cast<unsigned>(blocks), kTileWarps * 32, 0, stream>>>(
      value, spatial_shapes, level_start, nullptr, nullptr, fa, g, len_value, num_heads, 1, len_query, num_points, out);
  EFGB_LAUNCH_OK("box_attn_fwd_tile_kernel<fused>");
  return EFGB_OK;
}

extern "C" int efgb_box_attn_fused_backward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start,
                                            const float* offsets, const float* logits, const float* ref_windows,
                                            const float* kernel_indices, const float* grad_out, int batch, int len_value,
                                            int num_heads, int len_query, int num_points, int num_variables,
                                            int64_t logits_row_stride, int64_t offsets_row_stride, int query_grid_w,
                                            float* grad_value, float* grad_offsets, float* grad_logits, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  int rc = check_args(batch, len_value, num_heads, 32, 1, len_query, num_points, "box_attn_fused_backward");
  if (rc != EFGB_OK) return rc;
  EFGB_REQUIRE(efgb_box_attn_fused_supported(32, 1, num_points, num_variables) &&
                   static_cast<int64_t>(len_value) * num_heads * 32 < (1ll << 31),
               EFGB_EINVAL, "box_attn_fused_backward: unsupported shape");
  const size_t vbytes = static_cast<size_t>(batch) * len_value * num_heads * 32 * sizeof(float);
  if (vbytes) {
    EFGB_REQUIRE(grad_value, EFGB_EINVAL, "box_attn_fused_backward: null grad_value");
    EFGB_CUDA_OK(cudaMemsetAsync(grad_value, 0, vbytes, stream));
  }
  const int64_t nw = static_cast<int64_t>(batch) * len_query * num_heads;
  if (nw == 0) return EFGB_OK;
  EFGB_REQUIRE(value && spatial_shapes && level_start && offsets && logits && ref_windows && kernel_indices && grad_out &&
                   grad_offsets && grad_logits,
               EFGB_EINVAL, "box_attn_fused_backward: null pointer");
  FusedArgs fa{};
  fa.offsets = offsets;
  fa.logits = logits;
  fa.ref = ref_windows;
  fa.kidx = kernel_indices;
  fa.ld_logits = logits_row_stride > 0 ? logits_row_stride : static_cast<int64_t>(num_heads) * num_points;
  fa.ld_offsets = offsets_row_stride > 0 ? offsets_row_stride : static_cast<int64_t>(num_heads) * num_variables;
  fa.nv = num_variables;
  fa.g_offsets = grad_offsets;
  fa.g_logits = grad_logits;
  const TileGeom g = tile_geom(len_query, query_grid_w, nw);
  const int64_t blocks = static_cast<int64_t>(batch) * num_heads * g.tx * g.ty;
  EFGB_REQUIRE(blocks < (1ll << 31), EFGB_EINVAL, "box_attn_fused_backward: too many query tiles");
  box_attn_bwd_tile_kernel<true><<<static_cast<unsigned>(blocks), kTileWarps * 32, 0, stream>>>(
      value, spatial_shapes, level_start, nullptr, nullptr, grad_out, fa, g, len_value, num_heads, 1, len_query, num_points,
      grad_value, nullptr, nullptr);
  EFGB_LAUNCH_OK("box_attn_bwd_tile_kernel<fused>");
  return EFGB_OK;
}

// =================================================================================================
// Fused sampling-grid + attention-softmax ("where to attend"), forward and backward.
//
// Reference: Box3dAttention._where_to_attend and the softmax in Box3dAttention.forward
// (VD/modules/box_attention.py:62-95, :105-108), ~20 separate elementwise / reduction kernels over
// [B, LQ, H, L, 25, 2] tensors there.  One warp per (b, q, h): lane p < P owns sampling point p of every level.
//   offsets [B, LQ, H, L, NV]  (NV = 4: dx, dy, dw, dl;  NV = 5: + rotation)     from the box linear layer
//   logits  [B, LQ, H, L*P]                                                      from the attention linear layer
//   ref     [B, LQ, 7]  reference windows (cx, cy, cz, w, l, h, angle), no gradient
//   kidx    [P, 2]      kernel offsets in units of the box size (x, y)
//   box    = (cx + dx/8*w, cy + dy/8*l, w + dw/8*w, l + dl/8*l);  angle = NV == 5 ? (ref_angle + rot/16)*2pi : ref_angle
//   g      = kidx[p] * relu(size);  loc[p] = centre + (g.x*cos - g.y*sin, g.x*sin + g.y*cos)
//   attn   = softmax over all L*P logits of the (b, q, h)
// =================================================================================================
namespace efgb {

constexpr float kTwoPi = 6.283185307179586f;

template <bool kBackward>
__global__ void __launch_bounds__(256)
box_grid_softmax_kernel(const float* __restrict__ offsets, const float* __restrict__ logits, const float* __restrict__ ref,
                        const float* __restrict__ kidx, int64_t num_warps, int num_heads, int num_levels, int num_points,
                        int nv, int64_t ld_logits, int64_t ld_offsets, float* __restrict__ loc, float* __restrict__ attn,
                        const float* __restrict__ g_loc,
                        const float* __restrict__ g_attn, float* __restrict__ g_offsets, float* __restrict__ g_logits) {
  const int lane = threadIdx.x & 31;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (idx >= num_warps) return;
  const int64_t bq = idx / num_heads;
  const int hd = static_cast<int>(idx - bq * num_heads);
  const float* r = ref + bq * 7;
  const float cx = r[0], cy = r[1], w = r[3], l = r[4], ref_angle = r[6];
  const int lp = num_levels * num_points;
  // logits / offsets rows may be slices of a wider projection output (row strides ld_logits / ld_offsets)
  const float* lg = logits + bq * ld_logits + static_cast<int64_t>(hd) * lp;
  float* glg = kBackward ? g_logits + bq * ld_logits + static_cast<int64_t>(hd) * lp : nullptr;

  // ---- softmax over L*P logits (each lane strides over them)
  float m = -INFINITY;
  for (int e = lane; e < lp; e += 32) m = fmaxf(m, lg[e]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  float s = 0.f;
  for (int e = lane; e < lp; e += 32) s += expf(lg[e] - m);
  s = warp_sum(s);
  const float inv = 1.f / s;
  if (!kBackward) {
    for (int e = lane; e < lp; e += 32) attn[idx * lp + e] = expf(lg[e] - m) * inv;
  } else {
    float dot = 0.f;
    for (int e = lane; e < lp; e += 32) dot += expf(lg[e] - m) * inv * g_attn[idx * lp + e];
    dot = warp_sum(dot);
    for (int e = lane; e < lp; e += 32) {
      const float a = expf(lg[e] - m) * inv;
      glg[e] = a * (g_attn[idx * lp + e] - dot);
    }
  }

  // ---- sampling grid
  for (int lvl = 0; lvl < num_levels; ++lvl) {
    const int64_t off_at = bq * ld_offsets + (static_cast<int64_t>(hd) * num_levels + lvl) * nv;
    const float* off = offsets + off_at;
    const float o0 = off[0], o1 = off[1], o2 = off[2], o3 = off[3];
    const float bx = cx + o0 / 8.f * w, by = cy + o1 / 8.f * l;
    const float sw = w + o2 / 8.f * w, sl = l + o3 / 8.f * l;
    const float angle = nv == 5 ? (ref_angle + off[4] / 16.f) * kTwoPi : ref_angle;
    float sn, cs;
    sincosf(angle, &sn, &cs);
    const float rw = fmaxf(sw, 0.f), rl = fmaxf(sl, 0.f);
    float a_cx = 0.f, a_cy = 0.f, a_sw = 0.f, a_sl = 0.f, a_ang = 0.f;
    for (int p0 = 0; p0 < num_points; p0 += 32) {
      const int p = p0 + lane;
      if (p < num_points) {
        const float kx = kidx[2 * p], ky = kidx[2 * p + 1];
        const float gx = kx * rw, gy = ky * rl;
        const int64_t o = ((idx * num_levels + lvl) * num_points + p) * 2;
        if (!kBackward) {
          *reinterpret_cast<float2*>(loc + o) = make_float2(bx + (gx * cs - gy * sn), by + (gx * sn + gy * cs));
        } else {
          const float2 g = *reinterpret_cast<const float2*>(g_loc + o);
          a_cx += g.x;
          a_cy += g.y;
          const float ggx = g.x * cs + g.y * sn;   // d loss / d gx
          const float ggy = -g.x * sn + g.y * cs;  // d loss / d gy
          a_sw += ggx * kx;
          a_sl += ggy * ky;
          a_ang += g.x * (-gx * sn - gy * cs) + g.y * (gx * cs - gy * sn);
        }
      }
    }
    if (kBackward) {
      a_cx = warp_sum(a_cx);
      a_cy = warp_sum(a_cy);
      a_sw = warp_sum(a_sw);
      a_sl = warp_sum(a_sl);
      a_ang = warp_sum(a_ang);
      if (lane == 0) {
        float* go = g_offsets + off_at;
        go[0] = a_cx * w / 8.f;
        go[1] = a_cy * l / 8.f;
        go[2] = (sw > 0.f ? a_sw : 0.f) * w / 8.f;
        go[3] = (sl > 0.f ? a_sl : 0.f) * l / 8.f;
        if (nv == 5) go[4] = a_ang * kTwoPi / 16.f;
      }
    }
  }
}

}  // namespace efgb

extern "C" int efgb_box_grid_softmax_forward(const float* offsets, const float* logits, const float* ref_windows,
                                             const float* kernel_indices, int64_t num_bq, int num_heads, int num_levels,
                                             int num_points, int num_variables, int64_t logits_row_stride,
                                             int64_t offsets_row_stride, float* loc, float* attn, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(num_bq >= 0 && num_heads >= 1 && num_levels >= 1 && num_points >= 1 && (num_variables == 4 || num_variables == 5),
               EFGB_EINVAL, "box_grid_softmax_forward: bad shape");
  const int64_t nw = num_bq * num_heads;
  if (nw == 0) return EFGB_OK;
  EFGB_REQUIRE(offsets && logits && ref_windows && kernel_indices && loc && attn, EFGB_EINVAL, "box_grid_softmax_forward: null pointer");
  const int64_t ldl = logits_row_stride > 0 ? logits_row_stride : static_cast<int64_t>(num_heads) * num_levels * num_points;
  const int64_t ldo = offsets_row_stride > 0 ? offsets_row_stride : static_cast<int64_t>(num_heads) * num_levels * num_variables;
  box_grid_softmax_kernel<false><<<static_cast<unsigned>((nw + 7) / 8), 256, 0, stream>>>(
      offsets, logits, ref_windows, kernel_indices, nw, num_heads, num_levels, num_points, num_variables, ldl, ldo, loc, attn,
      nullptr, nullptr, nullptr, nullptr);
  EFGB_LAUNCH_OK("box_grid_softmax_kernel<fwd>");
  return EFGB_OK;
}

extern "C" int efgb_box_grid_softmax_backward(const float* offsets, const float* logits, const float* ref_windows,
                                              const float* kernel_indices, const float* grad_loc, const float* grad_attn,
                                              int64_t num_bq, int num_heads, int num_levels, int num_points,
                                              int num_variables, int64_t logits_row_stride, int64_t offsets_row_stride,
                                              float* grad_offsets, float* grad_logits, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(num_bq >= 0 && num_heads >= 1 && num_levels >= 1 && num_points >= 1 && (num_variables == 4 || num_variables == 5),
               EFGB_EINVAL, "box_grid_softmax_backward: bad shape");
  const int64_t nw = num_bq * num_heads;
  if (nw == 0) return EFGB_OK;
  EFGB_REQUIRE(offsets && logits && ref_windows && kernel_indices && grad_loc && grad_attn && grad_offsets && grad_logits, EFGB_EINVAL,
               "box_grid_softmax_backward: null pointer");
  const int64_t ldl = logits_row_stride > 0 ? logits_row_stride : static_cast<int64_t>(num_heads) * num_levels * num_points;
  const int64_t ldo = offsets_row_stride > 0 ? offsets_row_stride : static_cast<int64_t>(num_heads) * num_levels * num_variables;
  box_grid_softmax_kernel<true><<<static_cast<unsigned>((nw + 7) / 8), 256, 0, stream>>>(
      offsets, logits, ref_windows, kernel_indices, nw, num_heads, num_levels, num_points, num_variables, ldl, ldo, nullptr,
      nullptr, grad_loc, grad_attn, grad_offsets, grad_logits);
  EFGB_LAUNCH_OK("box_grid_softmax_kernel<bwd>");
  return EFGB_OK;
}
