// Device-side rulebook construction for submanifold and regular sparse 3-D convolution.
//
// The reference leaves this to spconv's indice-pair generation (efg/modeling/backbones/
// sparse_net.py:85-95 build SubMConv3d/SparseConv3d; no spconv source is vendored).  Here the
// active set is represented as an occupancy bitmap over the linear (b,z,y,x) cell id with a
// popcount prefix per 32-cell word (CellWord).  rank(cell) gives the row of a site in ascending
// linear order, so neighbour lookup is one 8-byte load + popc, the output set of a strided conv
// is produced already sorted and duplicate-free, and nothing needs a sort or a hash probe.
#include "common.cuh"

namespace efgb {

struct Grid4 {
  int B, D, H, W;
};

struct Conv3 {
  int k[3], s[3], p[3];
  int K;  // taps
};

__device__ __forceinline__ uint32_t cell_of(const Grid4& g, int b, int z, int y, int x) {
  return ((static_cast<uint32_t>(b) * g.D + z) * g.H + y) * static_cast<uint32_t>(g.W) + x;
}

__global__ void __launch_bounds__(256)
cells_mark_rows_kernel(const int4* __restrict__ coords, int64_t m, Grid4 g, CellWord* cells) {
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= m) return;
  int4 c = coords[i];  // (b,z,y,x)
  if (c.x < 0 || c.x >= g.B || c.y < 0 || c.y >= g.D || c.z < 0 || c.z >= g.H || c.w < 0 || c.w >= g.W) return;
  uint32_t cell = cell_of(g, c.x, c.y, c.z, c.w);
  atomicOr(&cells[cell >> 5].bits, 1u << (cell & 31));
}

__global__ void __launch_bounds__(256)
rank2row_kernel(const int4* __restrict__ coords, int64_t m, Grid4 g, const CellWord* __restrict__ cells,
                int32_t* __restrict__ rank2row) {
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= m) return;
  int4 c = coords[i];
  if (c.x < 0 || c.x >= g.B || c.y < 0 || c.y >= g.D || c.z < 0 || c.z >= g.H || c.w < 0 || c.w >= g.W) return;
  int r = cell_rank(cells, cell_of(g, c.x, c.y, c.z, c.w));
  if (r >= 0) rank2row[r] = static_cast<int32_t>(i);
}

// One thread per (row, tap): nbr[row, tap] = row index of the site at coord + tap - K/2, or -1.
__global__ void __launch_bounds__(256)
subm_nbr_kernel(const int4* __restrict__ coords, int64_t m, Grid4 g, Conv3 cv, const CellWord* __restrict__ cells,
                const int32_t* __restrict__ rank2row, int32_t* __restrict__ nbr) {
  int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= m * cv.K) return;
  const int64_t i = t / cv.K;
  const int k = static_cast<int>(t - i * cv.K);
  const int kx = k % cv.k[2];
  const int ky = (k / cv.k[2]) % cv.k[1];
  const int kz = k / (cv.k[2] * cv.k[1]);
  int4 c = coords[i];
  const int z = c.y + kz - cv.k[0] / 2, y = c.z + ky - cv.k[1] / 2, x = c.w + kx - cv.k[2] / 2;
  int r = -1;
  if (c.x >= 0 && c.x < g.B && z >= 0 && z < g.D && y >= 0 && y < g.H && x >= 0 && x < g.W) {
    r = cell_rank(cells, cell_of(g, c.x, z, y, x));
    if (r >= 0 && rank2row) r = rank2row[r];
  }
  nbr[t] = r;
}

// Output site reached by input (z,y,x) through tap (kz,ky,kx): o = (c + p - k) / s when divisible.
__device__ __forceinline__ bool out_site(const Conv3& cv, const Grid4& og, int z, int y, int x, int kz, int ky, int kx,
                                         int* oz, int* oy, int* ox) {
  int tz = z + cv.p[0] - kz, ty = y + cv.p[1] - ky, tx = x + cv.p[2] - kx;
  if (tz < 0 || ty < 0 || tx < 0) return false;
  if (tz % cv.s[0] || ty % cv.s[1] || tx % cv.s[2]) return false;
  tz /= cv.s[0];
  ty /= cv.s[1];
  tx /= cv.s[2];
  if (tz >= og.D || ty >= og.H || tx >= og.W) return false;
  *oz = tz;
  *oy = ty;
  *ox = tx;
  return true;
}

__global__ void __launch_bounds__(256)
sparse_mark_kernel(const int4* __restrict__ coords, int64_t m, Grid4 ig, Grid4 og, Conv3 cv, CellWord* cells) {
  int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= m * cv.K) return;
  const int64_t j = t / cv.K;
  const int k = static_cast<int>(t - j * cv.K);
  const int kx = k % cv.k[2];
  const int ky = (k / cv.k[2]) % cv.k[1];
  const int kz = k / (cv.k[2] * cv.k[1]);
  int4 c = coords[j];
  if (c.x < 0 || c.x >= ig.B || c.y < 0 || c.y >= ig.D || c.z < 0 || c.z >= ig.H || c.w < 0 || c.w >= ig.W) return;
  int oz, oy, ox;
  if (!out_site(cv, og, c.y, c.z, c.w, kz, ky, kx, &oz, &oy, &ox)) return;
  uint32_t cell = cell_of(og, c.x, oz, oy, ox);
  atomicOr(&cells[cell >> 5].bits, 1u << (cell & 31));
}

__global__ void __launch_bounds__(256)
cells_to_coords_kernel(const CellWord* __restrict__ cells, int64_t num_words, Grid4 og, int4* __restrict__ out_coords) {
  int64_t w = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w >= num_words) return;
  CellWord cw = cells[w];
  uint32_t bits = cw.bits;
  uint32_t r = cw.prefix;
  while (bits) {
    int bit = __ffs(bits) - 1;
    bits &= bits - 1;
    uint32_t cell = static_cast<uint32_t>(w) * 32u + bit;
    int x = cell % og.W;
    uint32_t q = cell / og.W;
    int y = q % og.H;
    q /= og.H;
    int z = q % og.D;
    int b = q / og.D;
    out_coords[r++] = make_int4(b, z, y, x);
  }
}

__global__ void __launch_bounds__(256)
sparse_pairs_kernel(const int4* __restrict__ coords, int64_t m, Grid4 ig, Grid4 og, Conv3 cv,
                    const CellWord* __restrict__ cells, int32_t* __restrict__ nbr, int32_t* __restrict__ nbr_t) {
  int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= m * cv.K) return;
  const int64_t j = t / cv.K;
  const int k = static_cast<int>(t - j * cv.K);
  const int kx = k % cv.k[2];
  const int ky = (k / cv.k[2]) % cv.k[1];
  const int kz = k / (cv.k[2] * cv.k[1]);
  int4 c = coords[j];
  int r = -1;
  if (!(c.x < 0 || c.x >= ig.B || c.y < 0 || c.y >= ig.D || c.z < 0 || c.z >= ig.H || c.w < 0 || c.w >= ig.W)) {
    int oz, oy, ox;
    if (out_site(cv, og, c.y, c.z, c.w, kz, ky, kx, &oz, &oy, &ox)) {
      r = cell_rank(cells, cell_of(og, c.x, oz, oy, ox));
      if (r >= 0) nbr[static_cast<int64_t>(r) * cv.K + k] = static_cast<int32_t>(j);
    }
  }
  if (nbr_t) nbr_t[t] = r;
}

static int make_conv(const int32_t* k, const int32_t* s, const int32_t* p, Conv3* cv) {
  for (int a = 0; a < 3; ++a) {
    cv->k[a] = k[a];
    cv->s[a] = s ? s[a] : 1;
    cv->p[a] = p ? p[a] : 0;
    if (cv->k[a] < 1 || cv->s[a] < 1 || cv->p[a] < 0) return -1;
  }
  cv->K = k[0] * k[1] * k[2];
  return 0;
}

static int64_t words_for(int batch, const int32_t* dhw) {
  double cells = static_cast<double>(batch) * dhw[0] * dhw[1] * dhw[2];
  if (cells <= 0 || cells >= 4294967295.0) return -1;
  return (static_cast<int64_t>(cells) + 31) / 32;
}

struct RulebookWs {
  CellWord* cells;
  uint32_t* scratch;
  int32_t* rank2row;
  uint32_t* total;
};

static bool carve(void* workspace, size_t bytes, int64_t num_words, int64_t num_rows, RulebookWs* out) {
  Workspace ws(workspace, bytes);
  out->cells = ws.take<CellWord>(num_words);
  out->scratch = ws.take<uint32_t>(scan_scratch_elems(num_words));
  out->rank2row = ws.take<int32_t>(num_rows > 0 ? num_rows : 1);
  out->total = ws.take<uint32_t>(1);
  return out->total != nullptr;
}

}  // namespace efgb

using namespace efgb;

extern "C" size_t efgb_rulebook_workspace_bytes(int batch, const int32_t* grid_dhw, int64_t num_rows) {
  int64_t nw = words_for(batch, grid_dhw);
  if (nw < 0) return 0;
  size_t b = align_up(nw * sizeof(CellWord)) + align_up(scan_scratch_elems(nw) * sizeof(uint32_t)) +
             align_up((num_rows > 0 ? num_rows : 1) * sizeof(int32_t)) + align_up(sizeof(uint32_t));
  return b + 1024;
}

extern "C" int efgb_subm_rulebook(const int32_t* coords, int64_t num_rows, int batch, const int32_t* grid_dhw,
                                  const int32_t* ksize, int rows_sorted, int32_t* nbr, void* workspace,
                                  size_t workspace_bytes, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(grid_dhw && ksize && batch >= 1 && num_rows >= 0, EFGB_EINVAL, "subm_rulebook: bad argument");
  EFGB_REQUIRE((coords && nbr) || num_rows == 0, EFGB_EINVAL, "subm_rulebook: null pointer");
  Conv3 cv;
  EFGB_REQUIRE(make_conv(ksize, nullptr, nullptr, &cv) == 0, EFGB_EINVAL, "subm_rulebook: bad kernel size");
  EFGB_REQUIRE((cv.k[0] & 1) && (cv.k[1] & 1) && (cv.k[2] & 1), EFGB_EINVAL,
               "subm_rulebook: submanifold convolution needs odd kernel sizes");
  const int64_t nw = words_for(batch, grid_dhw);
  EFGB_REQUIRE(nw > 0, EFGB_ERANGE, "subm_rulebook: batch*D*H*W does not fit 32-bit cell ids");
  EFGB_REQUIRE(num_rows * cv.K < (1ll << 40), EFGB_ERANGE, "subm_rulebook: too many rows");
  RulebookWs w;
  EFGB_REQUIRE(carve(workspace, workspace_bytes, nw, num_rows, &w), EFGB_EWORKSPACE, "subm_rulebook: workspace too small");
  if (num_rows == 0) return EFGB_OK;
  Grid4 g{batch, grid_dhw[0], grid_dhw[1], grid_dhw[2]};
  const int4* c4 = reinterpret_cast<const int4*>(coords);
  EFGB_CUDA_OK(cudaMemsetAsync(w.cells, 0, nw * sizeof(CellWord), stream));
  const unsigned nb = static_cast<unsigned>((num_rows + 255) / 256);
  cells_mark_rows_kernel<<<nb, 256, 0, stream>>>(c4, num_rows, g, w.cells);
  EFGB_LAUNCH_OK("cells_mark_rows_kernel");
  int rc = cells_scan(w.cells, nw, w.total, w.scratch, stream);
  if (rc != EFGB_OK) return rc;
  const int32_t* r2r = nullptr;
  if (!rows_sorted) {
    rank2row_kernel<<<nb, 256, 0, stream>>>(c4, num_rows, g, w.cells, w.rank2row);
    EFGB_LAUNCH_OK("rank2row_kernel");
    r2r = w.rank2row;
  }
  const int64_t nt = num_rows * cv.K;
  subm_nbr_kernel<<<static_cast<unsigned>((nt + 255) / 256), 256, 0, stream>>>(c4, num_rows, g, cv, w.cells, r2r, nbr);
  EFGB_LAUNCH_OK("subm_nbr_kernel");
  return EFGB_OK;
}

static int sparse_setup(const int32_t* in_dhw, const int32_t* ksize, const int32_t* stride, const int32_t* padding,
                        Conv3* cv, int32_t out_dhw[3]) {
  if (make_conv(ksize, stride, padding, cv) != 0) return -1;
  for (int a = 0; a < 3; ++a) {
    int num = in_dhw[a] + 2 * cv->p[a] - cv->k[a];
    if (num < 0) return -1;
    out_dhw[a] = num / cv->s[a] + 1;
  }
  return 0;
}

extern "C" int efgb_sparse_rulebook_phase1(const int32_t* coords_in, int64_t num_in, int batch, const int32_t* in_dhw,
                                           const int32_t* ksize, const int32_t* stride, const int32_t* padding,
                                           int32_t* out_dhw_host, int32_t* num_out_dev, void* workspace,
                                           size_t workspace_bytes, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(in_dhw && ksize && stride && padding && out_dhw_host && num_out_dev && batch >= 1 && num_in >= 0,
               EFGB_EINVAL, "sparse_rulebook_phase1: bad argument");
  EFGB_REQUIRE(coords_in || num_in == 0, EFGB_EINVAL, "sparse_rulebook_phase1: null coords");
  Conv3 cv;
  int32_t od[3];
  EFGB_REQUIRE(sparse_setup(in_dhw, ksize, stride, padding, &cv, od) == 0, EFGB_EINVAL,
               "sparse_rulebook_phase1: bad conv geometry");
  out_dhw_host[0] = od[0];
  out_dhw_host[1] = od[1];
  out_dhw_host[2] = od[2];
  const int64_t nw = words_for(batch, od);
  EFGB_REQUIRE(nw > 0, EFGB_ERANGE, "sparse_rulebook: batch*outD*outH*outW does not fit 32-bit cell ids");
  RulebookWs w;
  EFGB_REQUIRE(carve(workspace, workspace_bytes, nw, 1, &w), EFGB_EWORKSPACE, "sparse_rulebook_phase1: workspace too small");
  Grid4 ig{batch, in_dhw[0], in_dhw[1], in_dhw[2]};
  Grid4 og{batch, od[0], od[1], od[2]};
  EFGB_CUDA_OK(cudaMemsetAsync(w.cells, 0, nw * sizeof(CellWord), stream));
  if (num_in > 0) {
    const int64_t nt = num_in * cv.K;
    sparse_mark_kernel<<<static_cast<unsigned>((nt + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const int4*>(coords_in), num_in, ig, og, cv, w.cells);
    EFGB_LAUNCH_OK("sparse_mark_kernel");
  }
  int rc = cells_scan(w.cells, nw, reinterpret_cast<uint32_t*>(num_out_dev), w.scratch, stream);
  return rc;
}

extern "C" int efgb_sparse_rulebook_phase2(const int32_t* coords_in, int64_t num_in, int batch, const int32_t* in_dhw,
                                           const int32_t* ksize, const int32_t* stride, const int32_t* padding,
                                           int64_t num_out, int32_t* out_coords, int32_t* nbr, int32_t* nbr_t,
                                           void* workspace, size_t workspace_bytes, efgb_stream_t stream_) {
  cudaStream_t stream = as_stream(stream_);
  EFGB_REQUIRE(in_dhw && ksize && stride && padding && batch >= 1 && num_in >= 0 && num_out >= 0, EFGB_EINVAL,
               "sparse_rulebook_phase2: bad argument");
  Conv3 cv;
  int32_t od[3];
  EFGB_REQUIRE(sparse_setup(in_dhw, ksize, stride, padding, &cv, od) == 0, EFGB_EINVAL,
               "sparse_rulebook_phase2: bad conv geometry");
  const int64_t nw = words_for(batch, od);
  EFGB_REQUIRE(nw > 0, EFGB_ERANGE, "sparse_rulebook: batch*outD*outH*outW does not fit 32-bit cell ids");
  RulebookWs w;
  EFGB_REQUIRE(carve(workspace, workspace_bytes, nw, 1, &w), EFGB_EWORKSPACE, "sparse_rulebook_phase2: workspace too small");
  if (num_out == 0 || num_in == 0) return EFGB_OK;
  EFGB_REQUIRE(out_coords && nbr && coords_in, EFGB_EINVAL, "sparse_rulebook_phase2: null pointer");
  Grid4 ig{batch, in_dhw[0], in_dhw[1], in_dhw[2]};
  Grid4 og{batch, od[0], od[1], od[2]};
  cells_to_coords_kernel<<<static_cast<unsigned>((nw + 255) / 256), 256, 0, stream>>>(
      w.cells, nw, og, reinterpret_cast<int4*>(out_coords));
  EFGB_LAUNCH_OK("cells_to_coords_kernel");
  EFGB_CUDA_OK(cudaMemsetAsync(nbr, 0xFF, static_cast<size_t>(num_out) * cv.K * sizeof(int32_t), stream));
  const int64_t nt = num_in * cv.K;
  sparse_pairs_kernel<<<static_cast<unsigned>((nt + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const int4*>(coords_in), num_in, ig, og, cv, w.cells, nbr, nbr_t);
  EFGB_LAUNCH_OK("sparse_pairs_kernel");
  return EFGB_OK;
}
