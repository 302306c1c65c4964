/*
 * efgb200.h — C ABI of the B200-native (sm_100a) 3D-detection hot path that sits
 * behind EFG's Python operator surface.
 *
 * Conventions (every entry point):
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers on the
 *     caller's current CUDA device unless the parameter is documented as HOST;
 *   - the caller owns every buffer, including the scratch workspace (size it with
 *     the matching *_workspace_bytes query); nothing is allocated or freed here;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no entry
 *     point synchronises the device or touches the legacy default stream;
 *   - return value: 0 on success, a negative EFGB_E* code on failure;
 *     efgb_last_error() returns a thread-local description of the last failure.
 *
 * Reference interfaces replaced (paths relative to the V2AI/EFG tree):
 *   efg/operators/src/voxelize/voxelization.h:51-83   hard_voxelize / dynamic_voxelize
 *   efg/operators/src/voxelize/voxelization.h:96-128  dynamic_point_to_voxel_{forward,backward}
 *   efg/operators/src/box_attn/box_attn.h:29-83       box_attn_{forward,backward}
 *   efg/modeling/backbones/sparse_net.py:6-11         spconv.{SubMConv3d,SparseConv3d,SparseConvTensor.dense}
 *   efg/modeling/readers/voxel_reader.py:14-19        VoxelMeanFeatureExtractor (fused into the voxelizer)
 */
#ifndef EFGB200_H_
#define EFGB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EFGB_OK 0
#define EFGB_EINVAL (-1)    /* bad argument (null pointer, negative size, unsupported shape) */
#define EFGB_EWORKSPACE (-2) /* workspace too small */
#define EFGB_ECUDA (-3)     /* CUDA launch / runtime error */
#define EFGB_ERANGE (-4)    /* index space does not fit the 32-bit cell id used on device */

typedef void* efgb_stream_t; /* cudaStream_t */

const char* efgb_last_error(void);
int efgb_version(void);
/* number of CUDA kernels this library has launched in this process (monotonic; for bench accounting) */
uint64_t efgb_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Voxelizer — first-come "hard" voxelization, bit-exact with the reference's CPU twins
 *   efg/geometry/point_cloud_ops.py:6-53 (numba) and voxelization_cpu.cpp:44-96 (C++):
 *   c_j = floor((p_j - lo_j) / vs_j) in fp32; points outside the grid are dropped; voxel ids
 *   are assigned in order of first occurrence; at most max_points points are kept per voxel
 *   (in input order); the first point that would open voxel number max_voxels+1 ends the
 *   scene (it and every later point are dropped).  max_voxels < 0 means unlimited.
 *
 * A batch of scenes is voxelized in one call: points of scene b are rows
 * [scene_offsets[b], scene_offsets[b+1]) of `points`; voxels of scene b are emitted compactly
 * after those of scene b-1.  coors_dim = 3 writes (z,y,x) (reference layout), coors_dim = 4
 * writes (b,z,y,x) (the layout waymo.py:174-179 `collate` builds).
 *
 * Outputs (caller-allocated for cap = sum_b min(max_voxels, points_b) rows, contents of rows
 * >= the returned count are unspecified):
 *   voxels   [cap, max_points, F] f32, zero padded   (nullable: skip materialising it)
 *   coors    [cap, coors_dim] i32
 *   num_points_per_voxel [cap] i32
 *   mean_features [cap, F] f32 = sum of kept points / count (voxel_reader.py:14-19; nullable)
 *   voxel_counts [batch+1] i32: per-scene voxel count, then the total
 * ------------------------------------------------------------------------------------------ */
size_t efgb_voxelize_workspace_bytes(int64_t num_points, int batch);

int efgb_hard_voxelize(const float* points, int64_t num_points, int num_features,
                       const int32_t* scene_offsets, int batch,
                       const float* voxel_size_host3, const float* coors_range_host6,
                       int max_points, int max_voxels,
                       float* voxels, int32_t* coors, int coors_dim,
                       int32_t* num_points_per_voxel, float* mean_features,
                       int32_t* voxel_counts,
                       void* workspace, size_t workspace_bytes, efgb_stream_t stream);

/* dynamic_voxelize (voxelization_cpu.cpp:8-40): per-point (z,y,x), (-1,-1,-1) when dropped. */
int efgb_dynamic_voxelize(const float* points, int64_t num_points, int num_features,
                          const float* voxel_size_host3, const float* coors_range_host6,
                          int32_t* coors, efgb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dynamic scatter (scatter_points_cuda.cu:209-352): reduce point features by integer
 * coordinate.  Voxels come out in ascending linearised-coordinate order.  Two phases because
 * the voxel count M is data dependent and the caller owns the outputs:
 *   phase 1 marks/ranks the occupied cells and writes M to *num_voxels_dev,
 *   phase 2 fills voxel_feats[M,C], voxel_coors[M,3], point2voxel[N] (-1 = dropped), count[M].
 * dims_host3 = extent of each coordinate (coors.max(0)+1 in the reference); rows with any
 * negative coordinate are dropped.  reduce_type: 0 sum, 1 mean, 2 max.
 * ------------------------------------------------------------------------------------------ */
size_t efgb_scatter_workspace_bytes(int64_t num_points, const int32_t* dims_host3);
int efgb_scatter_phase1(const int32_t* coors, int64_t num_points, const int32_t* dims_host3,
                        int32_t* num_voxels_dev, void* workspace, size_t workspace_bytes,
                        efgb_stream_t stream);
int efgb_scatter_phase2(const float* feats, const int32_t* coors, int64_t num_points, int channels,
                        const int32_t* dims_host3, int reduce_type, int64_t num_voxels,
                        float* voxel_feats, int32_t* voxel_coors, int32_t* point2voxel,
                        int32_t* count, void* workspace, size_t workspace_bytes,
                        efgb_stream_t stream);
int efgb_scatter_backward(const float* grad_voxel_feats, const float* feats, const float* voxel_feats,
                          const int32_t* point2voxel, const int32_t* count, int64_t num_points,
                          int channels, int reduce_type, int64_t num_voxels, float* grad_feats,
                          void* workspace, size_t workspace_bytes, efgb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Rulebooks, built on device.  Active sites are rows of coords[M,4] i32 = (b,z,y,x) inside a
 * [batch, D, H, W] grid.  A rulebook is a dense neighbour table nbr[M_out, K] i32 with
 * K = kd*kh*kw and k = (kz*kh + ky)*kw + kx: nbr[o,k] is the INPUT row that kernel tap k of
 * output site o reads, or -1.  (The reference delegates this to spconv's indice-pair
 * generation, sparse_net.py:85-95; the pair list is the set {(k, nbr[o,k], o) : nbr[o,k] >= 0}.)
 *
 * Submanifold (SubMConv3d): outputs == inputs (same order); tap k reads coord + k - K/2.
 * rows_sorted != 0 promises coords are in ascending linear (b,z,y,x) order (true for the
 * output of efgb_sparse_rulebook_*), which removes one indirection.
 *
 * Regular (SparseConv3d): out = floor((in + 2p - k)/s) + 1 per axis; an output site exists
 * iff at least one input contributes; output rows are emitted in ascending linear order.
 *   phase 1: mark + rank output cells, write M_out to *num_out_dev;
 *   phase 2: out_coords[M_out,4], nbr[M_out,K], nbr_t[M_in,K] (nbr_t[j,k] = output row that
 *            input j feeds through tap k, or -1 — the transposed rulebook used by dgrad).
 * ------------------------------------------------------------------------------------------ */
size_t efgb_rulebook_workspace_bytes(int batch, const int32_t* grid_dhw_host3, int64_t num_rows);

int efgb_subm_rulebook(const int32_t* coords, int64_t num_rows, int batch,
                       const int32_t* grid_dhw_host3, const int32_t* ksize_host3, int rows_sorted,
                       int32_t* nbr, void* workspace, size_t workspace_bytes, efgb_stream_t stream);

int efgb_sparse_rulebook_phase1(const int32_t* coords_in, int64_t num_in, int batch,
                                const int32_t* in_dhw_host3, const int32_t* ksize_host3,
                                const int32_t* stride_host3, const int32_t* padding_host3,
                                int32_t* out_dhw_host3 /* HOST out */, int32_t* num_out_dev,
                                void* workspace, size_t workspace_bytes, efgb_stream_t stream);

int efgb_sparse_rulebook_phase2(const int32_t* coords_in, int64_t num_in, int batch,
                                const int32_t* in_dhw_host3, const int32_t* ksize_host3,
                                const int32_t* stride_host3, const int32_t* padding_host3,
                                int64_t num_out, int32_t* out_coords, int32_t* nbr, int32_t* nbr_t,
                                void* workspace, size_t workspace_bytes, efgb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Sparse convolution as an output-stationary gather-GEMM over a rulebook:
 *   out[o, :] = bias + sum_k  in[nbr[o,k], :] @ w[k]        (rows with nbr = -1 contribute 0)
 * w is [K, Cin, Cout] f32 (tap-major; the host mirror permutes spconv's [Cout,kd,kh,kw,Cin]).
 * dgrad is the same call on (grad_out, w^T per tap, nbr_t); wgrad is below.
 * fp32 FFMA arithmetic, fp32 accumulation.
 * ------------------------------------------------------------------------------------------ */
int efgb_spconv_forward(const float* in_feats, int64_t num_in, int c_in,
                        const float* w_kio, const float* bias /* nullable [c_out] */,
                        const int32_t* nbr, int64_t num_out, int num_taps, int c_out,
                        float* out_feats, efgb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Tensor-core path of the same gather-GEMM (tcgen05.mma kind::tf32 / kind::f16, accumulators in TMEM, weights
 * streamed with cp.async.bulk — the bulk-copy engine without a tensor map).  Weights are given in the reference parameter layout
 * [c_out, taps, c_in] (spconv 2.x, taps = kd*kh*kw flattened) and packed once per call into the
 * shared-memory image the tensor core reads:
 *   mode 0 forward            N = c_out, reduction channels c_red = c_in
 *   mode 1 dgrad (regular)    N = c_in,  c_red = c_out   (use with nbr_t)
 *   mode 2 dgrad (submanifold, tap order mirrored; use with the forward nbr)
 * split: 0 single-pass TF32; 1 3xTF32 (fp32-faithful, error ~2^-21); 2 bf16x3 (three bf16 products on kind::f16,
 * error ~2.5e-5, needs c_red % 8 == 0) — the packed image differs per split.
 * Supported when c_red % 4 == 0, N % 16 == 0, 16 <= N <= 256 or N a multiple of 256 up to 4096 (256-column slabs
 * over grid.y), taps <= 32.
 * ------------------------------------------------------------------------------------------ */
int efgb_spconv_tc_supported(int c_red, int n_out, int taps);
size_t efgb_spconv_tc_packed_bytes(int taps, int c_red, int n_out, int split);
int efgb_spconv_tc_pack(const float* w_param, int c_out, int taps, int c_in, int mode, int split,
                        float* packed, efgb_stream_t stream);
/* All bf16x3 (split = 2) weight images of a training step in one launch.  `jobs` is a DEVICE table of n_jobs rows of
 * eight int64: {w_param pointer, packed pointer, c_out, taps, c_in, mode, first block, blocks}; blocks =
 * efgb_spconv_tc_pack_blocks(...) (0: shape not supported), first block = running sum, total_blocks = the sum. */
int64_t efgb_spconv_tc_pack_blocks(int c_out, int taps, int c_in, int mode);
int efgb_spconv_tc_pack_batched(const int64_t* jobs, int n_jobs, int64_t total_blocks, efgb_stream_t stream);
int efgb_spconv_tc_forward(const float* in_feats, int64_t num_in, int c_red, const float* packed,
                           const float* bias /* nullable [n_out] */, const int32_t* nbr,
                           int64_t num_out, int taps, int n_out, int split, float* out_feats,
                           efgb_stream_t stream);
/* Same, with a fused epilogue: relu != 0 applies max(x, 0) after the bias (F.relu of the FFN,
 * VD/transformer.py:63).  (Fusing ReLU's backward mask into the dgrad epilogue was measured and dropped: the
 * per-row mask reads made the epilogue as long as the MMA phase.) */
int efgb_spconv_tc_forward_ex(const float* in_feats, int64_t num_in, int c_red, const float* packed,
                              const float* bias, const int32_t* nbr, int64_t num_out, int taps, int n_out,
                              int split, int relu, float* out_feats, efgb_stream_t stream);

/* Pre-split operand planes (bf16x3): a gathered input row is read by ~14 output rows of a 3x3x3 convolution, so the
 * fp32 -> bf16 hi / lo split is done ONCE per feature matrix instead of once per gather.
 *   efgb_split_bf16: in [rows, channels] f32 -> planes [rows][hi channels x bf16 | lo channels x bf16]
 *                    (same bytes per row as fp32; channels % 8 == 0)
 *   efgb_spconv_tc_forward_planes: the bf16x3 gather-GEMM over such planes (weights packed with split = 2); the A
 *                    producers are cp.async copies with zero fill for missing neighbours.  Needs a rulebook
 *                    (sparse convolutions; dense GEMMs read every row once and keep the fp32 entry point).
 * Replaces the same spconv call sites as efgb_spconv_tc_forward (sparse_net.py:85-95,125-147). */
int efgb_split_bf16(const float* in, int64_t rows, int channels, void* planes, efgb_stream_t stream);
int efgb_spconv_tc_planes_supported(int c_red, int n_out, int taps);
int efgb_spconv_tc_forward_planes(const void* in_planes, int64_t num_in, int c_red, const float* packed,
                                  const float* bias /* nullable */, const int32_t* nbr, int64_t num_out, int taps,
                                  int n_out, int relu, float* out_feats, efgb_stream_t stream);

/* Tensor-core wgrad: dw_param[co, tap, ci] = sum_o in[nbr[o,tap], ci] * grad_out[o, co], written in the
 * reference parameter layout [c_out, taps, c_in] (zero-filled by the callee, accumulated with
 * red.global.add).  Supported when c_in divides 128 or is a multiple of 128 (>= 16) and c_out % 16 == 0,
 * c_out <= 256 or a multiple of 256 up to 4096 (256-column slabs). */
int efgb_spconv_tc_wgrad_supported(int c_in, int c_out, int taps);
int efgb_spconv_tc_wgrad(const float* in_feats, int64_t num_in, int c_in, const float* grad_out,
                         const int32_t* nbr, int64_t num_out, int taps, int c_out, int split,
                         float* dw_param, efgb_stream_t stream);

/* dw[k, ci, co] = sum_o in[nbr[o,k], ci] * grad_out[o, co]; dw is zero-filled by the callee. */
int efgb_spconv_wgrad(const float* in_feats, int64_t num_in, int c_in,
                      const float* grad_out, const int32_t* nbr, int64_t num_out, int num_taps,
                      int c_out, float* dw_kio, efgb_stream_t stream);

/* SparseConvTensor.dense(): feats[M,C] at coords[M,4] -> out[B,C,D,H,W] (zero filled here);
 * and its adjoint (gather) for the backward pass. */
int efgb_sparse_to_dense(const float* feats, const int32_t* coords, int64_t num_rows, int channels,
                         int batch, const int32_t* grid_dhw_host3, float* dense, efgb_stream_t stream);
int efgb_dense_to_sparse(const float* dense, const int32_t* coords, int64_t num_rows, int channels,
                         int batch, const int32_t* grid_dhw_host3, float* feats, efgb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * BatchNorm1d over the active rows of a sparse tensor fused with the residual add and ReLU that follow it in the
 * reference's blocks: y = ReLU(BN(x) [+ residual])  (efg/modeling/backbones/sparse_net.py:85-95,135-147,162-163,455-469;
 * nn.BatchNorm1d semantics: batch statistics over the rows, biased variance for normalisation, unbiased for the running
 * estimate, momentum update in place).  training == 0 uses the running statistics.  `planes` (nullable) receives the
 * bf16 hi / lo operand planes of y for the next tensor-core convolution (see efgb_split_bf16).
 * Backward: dx, dgamma, dbeta and (dres nullable) the gradient of the residual; y is read for the ReLU mask.
 * Supported when cols % 4 == 0 and 256 % (cols / 4) == 0.
 * ------------------------------------------------------------------------------------------ */
int efgb_bn_supported(int cols);
size_t efgb_bn_workspace_bytes(int64_t rows, int cols);
int efgb_bn_forward(const float* x, int64_t rows, int cols, const float* gamma, const float* beta,
                    const float* residual /* nullable */, int relu, float eps, float momentum,
                    float* running_mean /* nullable in training */, float* running_var, int training, float* y,
                    float* save_mean, float* save_rstd, void* planes /* nullable */, void* workspace,
                    size_t workspace_bytes, efgb_stream_t stream);
int efgb_bn_backward(const float* dy, const float* x, const float* y, const float* gamma, const float* save_mean,
                     const float* save_rstd, int64_t rows, int cols, int relu, float* dx, float* dres /* nullable */,
                     float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes, efgb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * BEV IoU of rotated boxes and rotated NMS — the CenterPoint evaluation path.
 * Replaces efg._C.boxes_iou_bev_gpu / boxes_overlap_bev_gpu / nms_gpu / nms_normal_gpu
 * (efg/operators/src/vision.cpp:100-104, iou3d_nms/iou3d_nms.cpp:22-178, iou3d_nms_kernel.cu:98-342).
 * Boxes are [x, y, z, dx, dy, dz, heading] f32 rows.  Same polygon-clipping algorithm as the reference (1 cm corner
 * margin, strict crossings, atan2 ordering), so IoU values agree to fp32 rounding and NMS keeps the same boxes.
 *   efgb_boxes_bev: out[i * num_b + j] = IoU (mode 0) or overlap area (mode 1) of boxes_a[i] and boxes_b[j].
 *   efgb_nms_bev:   boxes sorted by descending score; keep[0 .. *num_keep) = indices kept by greedy suppression with
 *                   IoU > thresh (normal != 0: axis-aligned IoU, heading ignored).  keep (int64 [n]) and num_keep stay on
 *                   the DEVICE; the reference copies an N x N/64 mask to the host and scans there.
 * ------------------------------------------------------------------------------------------ */
size_t efgb_boxes_bev_workspace_bytes(int64_t num_a, int64_t num_b);
int efgb_boxes_bev(const float* boxes_a, int64_t num_a, const float* boxes_b, int64_t num_b, int mode, float* out,
                   void* workspace, size_t workspace_bytes, efgb_stream_t stream);
size_t efgb_nms_bev_workspace_bytes(int64_t n);
int efgb_nms_bev(const float* boxes_sorted, int64_t n, float thresh, int normal, int64_t* keep, int32_t* num_keep,
                 void* workspace, size_t workspace_bytes, efgb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Point-cloud augmentation on the device: RandomFlip3D -> GlobalRotation -> GlobalScaling ->
 * GlobalTranslation -> FilterByRange of efg/data/augmentations/extend_3d.py:121-316 applied to points [n, nfeat] f32
 * (x, y, z first) with the random draws supplied by the caller; the kept points are written in their original order to
 * out_points (capacity n rows) and their number to out_count (device int32) — no host synchronisation, the count
 * can feed efgb_hard_voxelize's scene offsets directly.  range_host6 == NULL: no filtering.
 * ------------------------------------------------------------------------------------------ */
size_t efgb_augment_workspace_bytes(int64_t n);
int efgb_augment_points(const float* points, int64_t n, int nfeat, int flip_x, int flip_y, float cosa, float sina,
                        float scale, const float* translation_host3 /* nullable */, const float* range_host6 /* nullable */,
                        float* out_points, int32_t* out_count, void* workspace, size_t workspace_bytes,
                        efgb_stream_t stream);

/* GT-database paste on the device (DatabaseSampling.__call__, efg/data/augmentations/extend_3d.py:68-92; the sampling and
 * the collision test stay host logic over <= 100 boxes, efg/data/samplers/gt_database_sampler.py:111-212).
 * out_points [n_paste + n_scene, nfeat] = [database points of the accepted objects, xyz translated by obj_centers |
 * scene points].  obj_table [num_obj, 3] int32 on the device = (first point in db_points, first output row, count), rows
 * in output order.  planes (nullable; rm_points_after_sample): [num_rm_boxes, 6, 4] inward face planes (normal, d) of the
 * pasted boxes (box_ops.py:285-310); scene points inside a box get xyz = 1e30 (dropped by any range check / the
 * voxelizer) instead of being compacted away: no count has to be read back. */
int efgb_paste_points(const float* db_points, const int32_t* obj_table, const float* obj_centers, int num_obj,
                      int64_t n_paste, const float* scene_points, int64_t n_scene, int nfeat,
                      const float* planes /* nullable */, int num_rm_boxes, float* out_points, efgb_stream_t stream);

/* CenterPoint label assignment, heatmap part (CP/voxelnet.py:44-192, CP/center_utils.py:29-58): objects[i] = (plane, x, y,
 * radius) int32 on the device; heatmaps [planes, height, width] f32, zero-initialised by the caller; every object's
 * Gaussian (sigma = (2 r + 1) / 6) is max-ed into its plane.  Replaces the host numpy drawing + upload of the maps. */
int efgb_draw_gaussians(const int32_t* objects, int num_objects, int height, int width, float* heatmaps,
                        efgb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Linear sum assignment for a batch of small dense cost matrices that live on the device — the Hungarian
 * matching of HungarianMatcher3d (VD/modules/matcher.py:54-91), which the reference solves on the host
 * with scipy.optimize.linear_sum_assignment after a blocking copy.  Same algorithm and tie-breaking as
 * scipy (shortest augmenting paths, double precision), so the assignments are the ones scipy returns.
 *   problem k: cost matrix rows_host[k] x cols_host[k] f32 at device address cost_ptrs_host[k], row
 *   stride ld_host[k] floats; writes n_k = min(rows, cols) (row, col) pairs sorted by row to
 *   out_rows / out_cols + out_offsets_host[k] (device int64).  The *_host arrays are host memory and
 *   are consumed before the call returns.  Non-finite costs: see efgb_lsa_batched_status.
 * ------------------------------------------------------------------------------------------ */
int efgb_lsa_batched(const void* const* cost_ptrs_host, const int32_t* rows_host,
                     const int32_t* cols_host, const int32_t* ld_host, const int64_t* out_offsets_host,
                     int count, int64_t* out_rows, int64_t* out_cols, efgb_stream_t stream);
/* Same, with a device status word: bit 0 is OR-ed in when a problem is infeasible (non-finite costs — where scipy
 * raises ValueError).  The pairs written for such a problem are still valid indices (free columns in ascending
 * order), so nothing downstream reads out of bounds; the caller checks the word when it next synchronises. */
int efgb_lsa_batched_status(const void* const* cost_ptrs_host, const int32_t* rows_host,
                            const int32_t* cols_host, const int32_t* ld_host, const int64_t* out_offsets_host,
                            int count, int64_t* out_rows, int64_t* out_cols, int32_t* status /* nullable */,
                            efgb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused residual add + LayerNorm over the last dimension (nn.LayerNorm(d_model) applied to
 * src + dropout(x) in every encoder layer, VD/transformer.py:60-64), forward and backward.
 *   forward:  z = x + residual (residual nullable); y = (z - mean) * rstd * gamma + beta; also writes z (nullable),
 *             mean [rows], rstd [rows] for the backward.  rstd = 1/sqrt(biased var + eps), as torch.
 *   backward: dz (= grad of x = grad of residual), dgamma, dbeta (deterministic two-pass column sums).
 * cols % 128 == 0, 128 <= cols <= 512.  workspace: efgb_add_layernorm_workspace_bytes(rows, cols).
 * ------------------------------------------------------------------------------------------ */
int efgb_add_layernorm_supported(int cols);
size_t efgb_add_layernorm_workspace_bytes(int64_t rows, int cols);
int efgb_add_layernorm_forward(const float* x, const float* residual, const float* gamma,
                               const float* beta, int64_t rows, int cols, float eps, float* y, float* z,
                               float* mean, float* rstd, efgb_stream_t stream);
int efgb_add_layernorm_backward(const float* dy, const float* z, const float* mean, const float* rstd,
                                const float* gamma, int64_t rows, int cols, float* dz, float* dgamma,
                                float* dbeta, void* workspace, size_t workspace_bytes,
                                efgb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Column sums out[c] = sum_r x[r, c] of a row-major [rows, cols] f32 matrix (cols % 4 == 0): the bias
 * gradient of the token-wise linear layers (autograd of F.linear in VD/transformer.py:41-64 and
 * VD/modules/box_attention.py:97-115, i.e. grad_out.sum(0) over B * 35 344 rows).  Deterministic
 * (two passes, no atomics).  workspace: efgb_colsum_workspace_bytes(rows, cols).
 * ------------------------------------------------------------------------------------------ */
size_t efgb_colsum_workspace_bytes(int64_t rows, int cols);
int efgb_colsum(const float* x, int64_t rows, int cols, float* out, void* workspace,
                size_t workspace_bytes, efgb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Box attention (box_attn.h:29-83; kernels box_attn_kernel.cuh:275-351 fwd, :35-190 math):
 *   out[b,q,h,c] = sum_{l,p} attn[b,q,h,l,p] * bilinear(value_l[b,:,h,c], loc[b,q,h,l,p])
 * with h_im = loc_y*H_l - 0.5, w_im = loc_x*W_l - 0.5, zero padding, taps taken only when
 * h_im > -1 && w_im > -1 && h_im < H_l && w_im < W_l.
 *   value [B, LV, H, Ch] f32; spatial_shapes [L,2] i64 (h,w); level_start [L] i64;
 *   loc [B, LQ, H, L, P, 2] f32 (x,y in [0,1]); attn [B, LQ, H, L, P] f32; out [B, LQ, H*Ch].
 * backward zero-fills grad_value then accumulates with red.global.add; grad_loc / grad_attn
 * are written once per element (no atomics).
 * query_grid_w is a LOCALITY HINT with no effect on the result: > 0 says that consecutive queries
 * form a row-major grid of that width (encoder self-attention, where the queries are the BEV cells),
 * so the kernel can give one CTA a 2-D patch of queries whose sampling footprints overlap; 0 = unknown.
 * ------------------------------------------------------------------------------------------ */
int efgb_box_attn_forward(const float* value, const int64_t* spatial_shapes,
                          const int64_t* level_start, const float* loc, const float* attn,
                          int batch, int len_value, int num_heads, int head_dim, int num_levels,
                          int len_query, int num_points, int query_grid_w, float* out,
                          efgb_stream_t stream);

int efgb_box_attn_backward(const float* value, const int64_t* spatial_shapes,
                           const int64_t* level_start, const float* loc, const float* attn,
                           const float* grad_out, int batch, int len_value, int num_heads,
                           int head_dim, int num_levels, int len_query, int num_points,
                           int query_grid_w, float* grad_value, float* grad_loc, float* grad_attn,
                           efgb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused "where to attend" + attention softmax of Box3dAttention
 * (VD/modules/box_attention.py:62-95 _where_to_attend, :105-108 softmax): from the box-offset and
 * attention-logit projections of every (batch*query, head) produce the sampling locations
 * loc [BQ, H, L, P, 2] and the softmax weights attn [BQ, H, L*P] that efgb_box_attn_* consume.
 *   offsets [BQ, H, L, NV] (NV = 4, or 5 with rotation), logits [BQ, H, L*P], ref_windows [BQ, 7]
 *   (cx, cy, cz, w, l, h, angle; no gradient), kernel_indices [P, 2].
 * logits / offsets (and their gradients) may be column slices of one wider projection output
 * [BQ, ld]: logits_row_stride / offsets_row_stride give the floats between consecutive BQ rows
 * (0 = densely packed, H*L*P and H*L*NV).
 * ------------------------------------------------------------------------------------------ */
int efgb_box_grid_softmax_forward(const float* offsets, const float* logits, const float* ref_windows,
                                  const float* kernel_indices, int64_t num_bq, int num_heads,
                                  int num_levels, int num_points, int num_variables,
                                  int64_t logits_row_stride, int64_t offsets_row_stride, float* loc,
                                  float* attn, efgb_stream_t stream);
int efgb_box_grid_softmax_backward(const float* offsets, const float* logits, const float* ref_windows,
                                   const float* kernel_indices, const float* grad_loc,
                                   const float* grad_attn, int64_t num_bq, int num_heads, int num_levels,
                                   int num_points, int num_variables, int64_t logits_row_stride,
                                   int64_t offsets_row_stride, float* grad_offsets, float* grad_logits,
                                   efgb_stream_t stream);

/* Box attention with the sampling grid and the softmax computed INSIDE the attention kernels: Box3dAttention.forward
 * (VD/modules/box_attention.py:97-115: _where_to_attend -> softmax -> BoxAttnFunction) as one operator over the outputs
 * of its two linear layers; `loc` / `attn` never exist in memory.  One value level, head_dim 32, <= 32 sampling points
 * (efgb_box_attn_fused_supported); arguments as efgb_box_grid_softmax_* and efgb_box_attn_*; grad_value is zeroed inside. */
int efgb_box_attn_fused_supported(int head_dim, int num_levels, int num_points, int num_variables);
int efgb_box_attn_fused_forward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start,
                                const float* offsets, const float* logits, const float* ref_windows,
                                const float* kernel_indices, int batch, int len_value, int num_heads, int len_query,
                                int num_points, int num_variables, int64_t logits_row_stride, int64_t offsets_row_stride,
                                int query_grid_w, float* out, efgb_stream_t stream);
int efgb_box_attn_fused_backward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start,
                                 const float* offsets, const float* logits, const float* ref_windows,
                                 const float* kernel_indices, const float* grad_out, int batch, int len_value,
                                 int num_heads, int len_query, int num_points, int num_variables,
                                 int64_t logits_row_stride, int64_t offsets_row_stride, int query_grid_w,
                                 float* grad_value, float* grad_offsets, float* grad_logits, efgb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EFGB200_H_ */
