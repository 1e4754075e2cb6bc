/*
 * bevfusion_b200.h -- C ABI of libbevfusion_b200.so (sm_100a / NVIDIA B200).
 *
 * Drop-in boundary for the BEVFusion view-transform / LiDAR-voxel hot path of
 * mit-han-lab/bevfusion (reference tree mounted at /root/reference in the build
 * container; citations are file:line in that tree).  Every entry point takes plain
 * device pointers, sizes and a CUDA stream; no torch types cross this boundary.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream)
 *   - return value: 0 on success, negative BEVB200_E* on failure; the failing
 *     call's message is available from bevb200_last_error() (thread-local)
 *   - calls are asynchronous on `stream`; none of them synchronises the device
 *   - no global mutable state: re-entrant across host threads / streams as long
 *     as the caller gives each in-flight call its own workspace
 *   - workspaces: query the *_workspace_bytes() function, allocate that many bytes
 *     (256-B aligned) and pass them in; contents are scratch
 *   - there is NO CPU fallback anywhere in this library
 */
#ifndef BEVFUSION_B200_H_
#define BEVFUSION_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BEVB200_API __attribute__((visibility("default")))
#else
#define BEVB200_API
#endif

#define BEVB200_OK 0
#define BEVB200_EINVAL (-1)    /* bad argument (null pointer, bad size, unsupported shape) */
#define BEVB200_ECUDA (-2)     /* CUDA runtime / launch error */
#define BEVB200_EWORKSPACE (-3) /* workspace too small */
#define BEVB200_EUNSUPPORTED (-4)

BEVB200_API int bevb200_version(void);
BEVB200_API const char *bevb200_last_error(void);
/* number of kernels launched by this library on the calling thread since the last
 * reset (bench.py's "gpu_launches" figure comes from here) */
BEVB200_API long long bevb200_launch_count(void);
BEVB200_API void bevb200_reset_launch_count(void);

/* ------------------------------------------------------------------------------------
 * bev_pool  (reference: mmdet3d/ops/bev_pool/src/bev_pool_cuda.cu, bev_pool_cpu.cpp)
 * ---------------------------------------------------------------------------------- */

/* Replaces `void bev_pool(int b,int d,int h,int w,int n,int c,int n_intervals,
 *                         const float* x,const int* geom_feats,const int* interval_starts,
 *                         const int* interval_lengths,float* out)`
 * (bev_pool_cuda.cu:86-90, called from bev_pool_forward bev_pool_cpu.cpp:22-47).
 *   x               [n, c] fp32, rows sorted by rank (the op's contract)
 *   geom_feats      [n, 4] int32 (x, y, z, b) of each sorted row
 *   interval_starts [n_intervals], interval_lengths [n_intervals] int32
 *   out             [b, d, h, w, c] fp32; EVERY element is written (cells without an
 *                   interval get 0), so the caller need not pre-zero it
 * Differences from the reference launcher: runs on `stream` (the reference uses the
 * legacy default stream, bev_pool_cuda.cu:88); sums each interval with a fixed
 * chunked order (deterministic run to run).
 * workspace: bevb200_bev_pool_workspace_bytes(n, c). */
BEVB200_API size_t bevb200_bev_pool_workspace_bytes(int n, int c);
BEVB200_API int bevb200_bev_pool(int b, int d, int h, int w, int n, int c, int n_intervals,
                     const float *x, const int32_t *geom_feats,
                     const int32_t *interval_starts, const int32_t *interval_lengths,
                     float *out, void *workspace, size_t workspace_bytes, void *stream);

/* Replaces `void bev_pool_grad(...)` (bev_pool_cuda.cu:92-96, called from
 * bev_pool_backward bev_pool_cpu.cpp:60-87).
 *   out_grad [b, d, h, w, c] fp32 contiguous -> x_grad [n, c] fp32 (sorted-row order).
 * Every row of x_grad that belongs to an interval is written; rows not covered by any
 * interval (none, for intervals produced by the reference's QuickCumsumCuda) are zeroed. */
BEVB200_API int bevb200_bev_pool_grad(int b, int d, int h, int w, int n, int c, int n_intervals,
                          const float *out_grad, const int32_t *geom_feats,
                          const int32_t *interval_starts, const int32_t *interval_lengths,
                          float *x_grad, void *stream);

/* B200-first variants of the two calls above: rows are read (written) through `perm`,
 * the sorted->original row map produced by bevb200_bev_pool_prepare_*, so that the
 * caller never materialises x[kept][argsort] (base.py:168, bev_pool.py:94).
 *   x / x_grad  [n_total, c] in ORIGINAL (unsorted, unfiltered) order
 *   perm        [n_total] int32: perm[i] for i < n = original row of sorted row i; the tail
 *               perm[n .. n_total) lists the filtered-out rows (only bev_pool_grad_perm reads it)
 * These calls REQUIRE tables produced by bevb200_bev_pool_prepare_* (intervals tile [0, n) in
 * ascending rank order); bev_pool_perm relies on that to zero-fill the empty cells in the same
 * pass.  bev_pool_grad_perm zero-fills the rows of x_grad that were filtered out. */
BEVB200_API int bevb200_bev_pool_perm(int b, int d, int h, int w, int n, int c, int n_intervals,
                          const float *x, const int32_t *perm, const int32_t *geom_feats,
                          const int32_t *interval_starts, const int32_t *interval_lengths,
                          float *out, void *workspace, size_t workspace_bytes, void *stream);
BEVB200_API int bevb200_bev_pool_grad_perm(int b, int d, int h, int w, int n, int n_total, int c,
                               int n_intervals, const float *out_grad, const int32_t *perm,
                               const int32_t *geom_feats, const int32_t *interval_starts,
                               const int32_t *interval_lengths, float *x_grad, void *stream);

/* Op layout [B, Z, X, Y, C] -> module layout [B, Z*C, X, Y] in one tiled transpose: replaces
 * `x.permute(0, 4, 1, 2, 3).contiguous()` (bev_pool.py:97) + `torch.cat(x.unbind(dim=2), 1)`
 * (base.py:174).  rows = X*Y.  out_batch_stride (floats; 0 = nz*rows*c) lets `out` be a channel
 * slice of a wider [B, C_total, X, Y] buffer, i.e. the camera half of the fuser's concatenated
 * input (fusers/conv.py:16 `torch.cat(inputs, dim=1)`) written in place. */
BEVB200_API int bevb200_bev_channels_first(const float *in, float *out, int batch, int nz, int rows, int c,
                                           long long out_batch_stride, void *stream);

/* Fused LSS lift + pool ("next" row (f)1 of SURVEY.md section 8; BEVPoolv2-style): the lifted volume
 * x[p, :] = depth[p] * ctx[pixel(p), :] of LSSTransform / DepthLSSTransform.get_cam_feats
 * (mmdet3d/models/vtransforms/lss.py:68-73, depth_lss.py:92-97) is never materialised; the pooling
 * kernel gathers the (L2-resident) context row and the depth weight of every kept frustum point:
 *     out[cell] = sum_{p in cell} fp32(depth[p] * ctx[pixel(p), :])
 *   depth  [n_total] fp32 = softmax depth volume flattened as [B*N, D, fH*fW]
 *   ctx    [B*N*fH*fW, c] fp32, channels-last context features
 *   original point index i -> pixel row (i / (depth_bins*pixels_per_camera)) * pixels_per_camera
 *                                       + i % pixels_per_camera
 * Tables as for bevb200_bev_pool_perm (from bevb200_bev_pool_prepare_geom).  Forward only. */
BEVB200_API int bevb200_bev_pool_lift(int b, int d, int h, int w, int n, int c, int n_intervals,
                          const float *depth, const float *ctx, int depth_bins,
                          int pixels_per_camera, const int32_t *perm, const int32_t *geom_feats,
                          const int32_t *interval_starts, const int32_t *interval_lengths,
                          float *out, void *workspace, size_t workspace_bytes, void *stream);

/* Column formulation of the fused lift + pool (round 2).  The BEV grid collapses z, so the fH pixels of one image
 * column (cam, w) at one depth bin (almost) always fall into one cell: the kept points are grouped into SEGMENTS
 * (column, depth bin, cell) with a bit mask over the pixel row h (feature_h <= 64), a segment's value
 *     T[seg, :] = sum_{h in mask} fp32(depth[cam, d, h, w] * ctx[cam, h, w, :])
 * is evaluated from ONE shared-memory copy of the column's context rows and depth values, and a cell is the sum of
 * its segments in a fixed order.  Nothing is assumed about the cameras (tilt or the z filter only add segments).
 * Reads depth + ctx once (13.4 MB at C2) instead of one 320-byte context row per kept point (588 MB through L2).
 *   bevb200_bev_pool_lift_prepare   per calibration, from the tables of bevb200_bev_pool_prepare_*: perm [n_kept]
 *       (sorted -> original index over [cameras, depth_bins, feature_h, feature_w], cameras = B*N) and
 *       interval_starts.  Outputs (device): col_begin int32[cameras*feature_w + 1], seg_key / seg_mask
 *       uint64[n_kept] (first n_segments entries valid; key = ((column*depth_bins + d) << 32) | interval),
 *       seg_slot int32[n_kept], interval_slot_begin int32[n_intervals + 1], n_segments int32[1].
 *       Synchronises the stream once (it needs the segment count).
 *   bevb200_bev_pool_lift_columns   per frame: depth [cameras, depth_bins, feature_h, feature_w] fp32,
 *       ctx [cameras, feature_h, feature_w, c] fp32 -> out [b, d, h, w, c] (fully written).  n_segments as read
 *       back from the prepare call; workspace: bevb200_bev_pool_lift_columns_workspace_bytes(). */
BEVB200_API size_t bevb200_bev_pool_lift_prepare_workspace_bytes(int n_kept);
BEVB200_API int bevb200_bev_pool_lift_prepare(const int32_t *perm, const int32_t *interval_starts, int n_kept,
                                  int n_intervals, int cameras, int depth_bins, int feature_h, int feature_w,
                                  int32_t *col_begin, uint64_t *seg_key, uint64_t *seg_mask, int32_t *seg_slot,
                                  int32_t *interval_slot_begin, int32_t *n_segments, void *workspace,
                                  size_t workspace_bytes, void *stream);
BEVB200_API size_t bevb200_bev_pool_lift_columns_workspace_bytes(int n_segments, int n_intervals, int c);
BEVB200_API int bevb200_bev_pool_lift_columns(int b, int d, int h, int w, int n, int c, int n_intervals,
                                  const float *depth, const float *ctx, int cameras, int depth_bins,
                                  int feature_h, int feature_w, const int32_t *geom_feats,
                                  const int32_t *interval_starts, const int32_t *col_begin,
                                  const uint64_t *seg_key, const uint64_t *seg_mask, const int32_t *seg_slot,
                                  const int32_t *interval_slot_begin, int n_segments, float *out,
                                  void *workspace, size_t workspace_bytes, void *stream);


/* bev_pool precompute.  Replaces, on device and in one call, the index glue of
 * BaseTransform.bev_pool (mmdet3d/models/vtransforms/base.py:149-169: quantise, batch
 * index, bounds mask), bev_pool() (ops/bev_pool/bev_pool.py:87-94: rank, argsort,
 * gathers) and QuickCumsumCuda.forward (bev_pool.py:41-46: interval table).
 *
 *   geom_xyz   [n_total, 3] fp32 lidar-frame frustum points (get_geometry output)
 *   n_per_batch  n_total / B (batch index of point i is i / n_per_batch, base.py:151-157)
 *   lower_host[3] = (bx - dx/2) evaluated in fp32 exactly as base.py:149 does,
 *   dx_host[3], nx_host[3] = grid cells along x, y, z (base.py:15-21)
 *   idx = trunc_toward_zero((g - lower) / dx) in fp32; kept iff 0 <= idx_k < nx_k
 *   rank = x*(W*D*B) + y*(D*B) + z*B + b with (B, D, H, W) = (B, nz, nx, ny)
 * Outputs (all sized for the worst case n_total):
 *   ranks_sorted [n_total] int32, perm [n_total] int32 (ascending original index
 *   inside equal ranks -- a stable sort; the reference's argsort is unstable so any
 *   order is within its contract), geom_sorted [n_total, 4] int32 (x, y, z, b),
 *   interval_starts / interval_lengths [n_total] int32,
 *   counts [2] int32 = {n_kept, n_intervals}  (device memory; read it back once)
 */
BEVB200_API size_t bevb200_bev_pool_prepare_workspace_bytes(int n_total);
BEVB200_API int bevb200_bev_pool_prepare_geom(const float *geom_xyz, int n_total, int n_per_batch,
                                  const float *lower_host, const float *dx_host,
                                  const int32_t *nx_host, int B, int32_t *ranks_sorted,
                                  int32_t *perm, int32_t *geom_sorted,
                                  int32_t *interval_starts, int32_t *interval_lengths,
                                  int32_t *counts, void *workspace, size_t workspace_bytes,
                                  void *stream);
/* Same, starting from already quantised int64 coords [n, 4] = (x, y, z, b) as handed to
 * bev_pool() (bev_pool.py:84).  Rows outside [0,H)x[0,W)x[0,D)x[0,B) are dropped
 * (the reference would write out of bounds for them). */
/* bevb200_bev_pool_prepare_geom with BaseTransform.get_geometry (mmdet3d/models/vtransforms/base.py:92-135) fused in:
 * the lidar-frame frustum points are computed inside the rank pass and the 96 MB [B, N, D, fH, fW, 3] tensor is never
 * written (geom_out, nullable, receives it for checks).  Explicit fp32 arithmetic, products summed left to right:
 *   p = frustum - post_trans;  q = inv(post_rot) p;  u = (q.x q.z, q.y q.z, q.z);  v = (R inv(K)) u + t;
 *   optionally v = extra_R v + extra_t (the lidar augmentation, per batch item).
 * frustum [n_frustum, 3] = (u, v, d) of create_frustum (base.py:66-89); cam_params [cameras, 24] =
 * {inv(post_rot) 9, post_trans 3, camera2lidar_rot . inv(intrinsics) 9, camera2lidar_trans 3} -- the 3x3 inverses and
 * the product are the caller's (torch, as the reference computes them); extra_params [B, 12] or NULL; cameras = B *
 * cams_per_batch.  Not bit-identical to torch's batched matmul (whose summation order is unspecified): a frustum point
 * within an ulp of a cell boundary can land in the neighbouring cell. */
BEVB200_API int bevb200_bev_pool_prepare_cameras(const float *frustum, int n_frustum, int cameras, int cams_per_batch,
                                     const float *cam_params, const float *extra_params,
                                     const float *lower_host, const float *dx_host, const int32_t *nx_host, int B,
                                     float *geom_out, int32_t *ranks_sorted, int32_t *perm, int32_t *geom_sorted,
                                     int32_t *interval_starts, int32_t *interval_lengths, int32_t *counts,
                                     void *workspace, size_t workspace_bytes, void *stream);
BEVB200_API int bevb200_bev_pool_prepare_coords(const int64_t *coords, int n, int B, int D, int H, int W,
                                    int32_t *ranks_sorted, int32_t *perm,
                                    int32_t *geom_sorted, int32_t *interval_starts,
                                    int32_t *interval_lengths, int32_t *counts,
                                    void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------
 * voxelization (reference: mmdet3d/ops/voxel/src/voxelization.h, voxelization_cuda.cu)
 * ---------------------------------------------------------------------------------- */

/* Replaces voxelization::hard_voxelize_gpu (voxelization_cuda.cu:231-373; bound as
 * voxel_layer.hard_voxelize, voxelization.h:58-81), deterministic=True semantics:
 *   - c_k = (int)floor((p_k - min_k) / vs_k) in fp32, valid iff 0 <= c_k < grid_k,
 *     grid_k = round((max_k - min_k) / vs_k)                 (:37-58, :256-258)
 *   - voxel ids in order of first appearance by point index; only the first
 *     max_voxels distinct voxels are kept                     (:149-180)
 *   - a point's slot = number of earlier points in its voxel; kept iff < max_points
 *   - voxels[v, slot, :] = points[i, :], coors[v] = (cx, cy, cz) (x,y,z order in this
 *     fork), num_points_per_voxel[v] = min(count, max_points)
 * The caller pre-zeroes voxels / coors / num_points_per_voxel at cap size
 * (voxelize.py:52-54); only used slots are written.  The number of voxels is written to
 * voxel_num (DEVICE int32[1]) -- the reference returns it as a host int after a device
 * sync (:369-370); the host wrapper does that read.
 * No O(N^2) scan, no <<<1,1>>> kernel, no device synchronisation. */
BEVB200_API size_t bevb200_hard_voxelize_workspace_bytes(int num_points, int max_points);
BEVB200_API int bevb200_hard_voxelize(const float *points, int num_points, int num_features,
                          const float *voxel_size_host, const float *coors_range_host,
                          int max_points, int max_voxels, float *voxels, int32_t *coors,
                          int32_t *num_points_per_voxel, int32_t *voxel_num,
                          void *workspace, size_t workspace_bytes, void *stream);

/* Fused voxelize + mean-reduce + batch pad: what BEVFusion.voxelize does around the op when
 * voxelize_reduce is on (mmdet3d/models/fusion_models/bevfusion.py:169-197 -- hard_voxelize, then
 * `feats.sum(dim=1) / sizes` and `F.pad(coords, (1, 0), value=k)`), without materialising the
 * [M, max_points, F] voxel tensor.  Same voxel order / caps as bevb200_hard_voxelize.
 *   feats  [max_voxels, F] mean point row per voxel (slot-order fp32 sum, IEEE divide)
 *   coords4 [max_voxels, 4] int32 (batch_idx, x, y, z), 16-byte aligned
 *   num_points_per_voxel [max_voxels] (nullable), voxel_num device int32
 * Rows >= voxel_num are left untouched.  F <= 8.  Workspace as bevb200_hard_voxelize. */
BEVB200_API int bevb200_hard_voxelize_mean(const float *points, int num_points, int num_features,
                               const float *voxel_size_host, const float *coors_range_host,
                               int max_points, int max_voxels, int batch_idx, float *feats,
                               int32_t *coords4, int32_t *num_points_per_voxel, int32_t *voxel_num,
                               void *workspace, size_t workspace_bytes, void *stream);

/* Replaces voxelization::dynamic_voxelize_gpu (voxelization_cuda.cu:485-528):
 * coors[i] = (cx, cy, cz) or (-1,-1,-1) when the point is out of range.  (The reference
 * kernel only guarantees coors[i][0] == -1 for such points, :38-47.) */
BEVB200_API int bevb200_dynamic_voxelize(const float *points, int num_points, int num_features,
                             const float *voxel_size_host, const float *coors_range_host,
                             int32_t *coors, void *stream);

/* Fused BEVFusion.voxelize glue (mmdet3d/models/fusion_models/bevfusion.py:183,191-195):
 * feats[v, :] = sum_slot voxels[v, slot, :] / num[v]  and coords[v] = (batch_idx, x, y, z). */
BEVB200_API int bevb200_voxel_mean(const float *voxels, const int32_t *coors, const int32_t *num_points,
                       int num_voxels, int max_points, int num_features, int batch_idx,
                       float *feats, int32_t *coords4, void *stream);

/* ---- LiDAR depth images for the depth-aware camera lift ------------------------------------
 * BaseDepthTransform.forward's per-sample loop (mmdet3d/models/vtransforms/base.py:279-329):
 * undo the lidar augmentation, project every point into every camera with lidar2image, clamp z to
 * [1e-5, 1e5], perspective divide, apply the image augmentation, keep points on the image and
 * write their distance at (row, col).  Colliding points: the one with the LARGEST index wins (the
 * sequential meaning of the reference's index_put; its CUDA result is arbitrary).
 *   points [N, F] device fp32 (xyz first); lidar_aug_matrix [4,4], lidar2image [ncam,4,4],
 *   img_aug_matrix [ncam,4,4]: device fp32 row-major, one sample.
 *   one_hot == 0: channel 0 = distance ("scalar", :318-319); one_hot != 0: depth_bins channels,
 *   1.0 at bin (long)min(dist, depth_bins-1) (:320-325).  add_features != 0 appends F channels
 *   holding the winner's point row with xyz minus the lidar-aug translation (:327-329; the
 *   reference has shifted them in place at :290 -- the caller's points are NOT modified here).
 *   depth [ncam, channels, H, W] is fully written.  ncam <= 16. */
BEVB200_API size_t bevb200_depth_rasterize_workspace_bytes(int ncam, int height, int width);
BEVB200_API int bevb200_depth_rasterize(const float *points, int num_points, int num_features,
                            const float *lidar_aug_matrix, const float *lidar2image,
                            const float *img_aug_matrix, int ncam, int height, int width, int one_hot,
                            int depth_bins, int add_features, float *depth, void *workspace,
                            size_t workspace_bytes, void *stream);

/* ---- DynamicScatter (mmdet3d/ops/voxel/src/scatter_points_cuda.cu:187-315; pybind
 * `dynamic_point_to_voxel_forward / _backward`, voxelization.cpp:10-11) -----------------------
 * reduce_type: 0 sum, 1 mean, 2 max (scatter_points_cuda.cu:7).
 * Forward: rows of `coors` [N, ndim] (ndim 1..4; a row with a negative entry is dropped) name
 * voxels; output voxels are the unique rows in lexicographic order (= at::unique_dim, :208).
 *   reduced_feats [>=M, C], out_coors [>=M, ndim], coors_map [N] (voxel row or -1),
 *   reduce_count [>=M], reduce_from [>=M, C] (nullable; arg-max point per (voxel, channel) for
 *   the max backward), meta[2] device int32 = {M, number of rows with an entry too large for the
 *   key (>= 2^20 for ndim <= 3, >= 2^15 for ndim 4; such rows are dropped and the caller should
 *   raise)}.  Size the outputs for M = N.  Sums run in ascending point order: reproducible,
 *   unlike the reference's float atomics. */
BEVB200_API size_t bevb200_dynamic_scatter_workspace_bytes(int num_points);
BEVB200_API int bevb200_dynamic_scatter(const float *feats, const int32_t *coors, int num_points,
                            int num_features, int ndim, int reduce_type, float *reduced_feats,
                            int32_t *out_coors, int32_t *coors_map, int32_t *reduce_count,
                            int32_t *reduce_from, int32_t *meta, void *workspace,
                            size_t workspace_bytes, void *stream);
/* Backward (:243-315): grad_feats [N, C] is fully written.  For max, `reduce_from` [M, C] is
 * used as given when reduce_from_valid != 0, else rebuilt by traceback from feats /
 * reduced_feats (smallest point index attaining the maximum, :145-160). */
BEVB200_API int bevb200_dynamic_scatter_backward(const float *grad_reduced_feats, const float *feats,
                                     const float *reduced_feats, const int32_t *coors_map,
                                     const int32_t *reduce_count, int32_t *reduce_from,
                                     int reduce_from_valid, int num_points, int num_reduced,
                                     int num_features, int reduce_type, float *grad_feats,
                                     void *stream);

/* ------------------------------------------------------------------------------------
 * sparse convolution (reference: mmdet3d/ops/spconv/include/spconv/spconv_ops.h,
 * indice.cu.h, geometry.h, reordering.cu.h)
 * ---------------------------------------------------------------------------------- */

/* Rulebook.  Replaces spconv::getIndicePair<3> (spconv_ops.h:27-141; bound as
 * sparse_conv_ext.get_indice_pairs_3d, all.cc:22-27) for transpose == 0.
 *
 * The rulebook is produced in offset-major neighbour-table form
 *     nbr[k * n_out + o] = input row feeding output row o through kernel offset k, or -1
 * (k = (off_x*K_y + off_y)*K_z + off_z as geometry.h:57-73), which is what the
 * implicit-GEMM kernel consumes; bevb200_rulebook_to_pairs() converts it to the
 * reference's indicePairs[K,2,N] / indiceNum[K] layout.
 *
 * Two calls because the number of outputs of a strided conv is data dependent:
 *   1. bevb200_rulebook_prepare: marks the active output sites in a bitmap over the
 *      dense output grid, ranks them, and writes n_out (device int32[1]).  SubM:
 *      n_out = n_in and the output sites are the input sites in input order
 *      (spconv_ops.h:76-101).  Strided: output rows are ordered by ascending flat index
 *      ((b*X + x)*Y + y)*Z + z -- the order of the reference's GPU path after
 *      torch::_unique (spconv_ops.h:130-136, indice.cu.h:112-127).
 *   2. bevb200_rulebook_fill: writes out_indices [n_out, 4] (b, x, y, z) and
 *      nbr [K, n_out] using the state left in `workspace` by step 1.
 */
BEVB200_API size_t bevb200_rulebook_workspace_bytes(int n_in, int batch_size, const int32_t *out_shape_host);
BEVB200_API int bevb200_rulebook_prepare(const int32_t *indices, int n_in, int batch_size,
                             const int32_t *spatial_shape_host, const int32_t *out_shape_host,
                             const int32_t *ksize_host, const int32_t *stride_host,
                             const int32_t *padding_host, const int32_t *dilation_host,
                             int subm, int32_t *n_out, void *workspace,
                             size_t workspace_bytes, void *stream);
BEVB200_API int bevb200_rulebook_fill(const int32_t *indices, int n_in, int batch_size,
                          const int32_t *spatial_shape_host, const int32_t *out_shape_host,
                          const int32_t *ksize_host, const int32_t *stride_host,
                          const int32_t *padding_host, const int32_t *dilation_host,
                          int subm, int n_out, int32_t *out_indices, int32_t *nbr,
                          void *workspace, size_t workspace_bytes, void *stream);

/* SubM rulebook of a tensor whose rows ARE the outputs of the strided conv whose
 * bevb200_rulebook_prepare() left its state in `workspace` (same batch size and output grid =
 * this tensor's spatial shape): the site bitmap and its ranks are reused as they are (the rows
 * are already in rank order), so no marking / scan / rank table pass is needed. */
BEVB200_API int bevb200_rulebook_fill_subm_sorted(const int32_t *indices, int n, int batch_size,
                                      const int32_t *spatial_shape_host, const int32_t *ksize_host,
                                      const int32_t *dilation_host, int32_t *nbr, void *workspace,
                                      size_t workspace_bytes, void *stream);

/* nbr[K, n_out] -> indicePairs[K, 2, n_in] (-1 padded) + indiceNum[K]
 * (layout of spconv_ops.h:53-57).  Pairs of one offset are emitted in ascending output
 * row (the reference's GPU order is atomic-arrival order, i.e. unspecified). */
BEVB200_API int bevb200_rulebook_to_pairs(const int32_t *nbr, int kernel_volume, int n_out, int n_in,
                              int32_t *indice_pairs, int32_t *indice_num, void *stream);
/* indicePairs[K, 2, n_pairs_dim] + indiceNum[K] -> nbr[K, n_out].  inverse != 0 swaps the
 * roles of the two pair columns (spconv_ops.h:316,345). */
BEVB200_API int bevb200_pairs_to_nbr(const int32_t *indice_pairs, const int32_t *indice_num,
                         int kernel_volume, int pairs_dim, int n_out, int inverse,
                         int32_t *nbr, void *stream);

/* Sparse convolution forward.  Replaces spconv::indiceConv<float> (spconv_ops.h:260-361;
 * bound as sparse_conv_ext.indice_conv_fp32) -- the per-offset gather -> mm_out ->
 * scatter-add loop -- with ONE output-stationary implicit-GEMM kernel:
 *     out[o, :] = epilogue( sum_k features[nbr[k, o], :] @ weight[k] )
 *   features [n_in, c_in] fp32, weight [K, c_in, c_out] fp32 (conv.py:100 layout
 *   [kx,ky,kz,Cin,Cout] flattened), nbr [K, n_out], out [n_out, c_out] fp32.
 * Optional fused epilogue (all may be NULL / 0), applied in this order:
 *     y = acc * scale[c] + shift[c]      (folded eval-mode BatchNorm1d / bias)
 *     y += residual[o, c]                (SparseBasicBlock identity, sparse_block.py:105)
 *     y = max(y, 0) if relu
 * precision: BEVB200_PREC_FP32  exact fp32 FFMA accumulation (SIMT)
 *            BEVB200_PREC_TF32X3 tcgen05 tensor cores, 3xTF32 split (fp32-class accuracy)
 *            BEVB200_PREC_TF32   tcgen05 single-pass TF32 (fast mode, ~1e-3 rel)
 *            BEVB200_PREC_BF16X3 tcgen05 bf16 hi/lo split: half the MMAs and operand bytes of TF32X3,
 *                                per-product error ~2^-17 (measured ~1e-5 relative per layer)
 */
#define BEVB200_PREC_FP32 0
#define BEVB200_PREC_TF32X3 1
#define BEVB200_PREC_TF32 2
#define BEVB200_PREC_BF16X3 3 /* tcgen05 kind::f16: a = bf16 hi + bf16 lo, hi*hi + hi*lo + lo*hi (3 MMAs per 16 K) */
BEVB200_API int bevb200_spconv_forward(const float *features, const float *weight, const int32_t *nbr,
                           int n_in, int n_out, int c_in, int c_out, int kernel_volume,
                           const float *scale, const float *shift, const float *residual,
                           int relu, int precision, float *out, void *stream);

/* Tensor-core path with pre-packed weights.  The tcgen05 kernel consumes the weights as a
 * pre-swizzled shared-memory image (K blocks of 32 over the concatenated (offset, channel)
 * axis, tf32 hi / lo parts); bevb200_spconv_forward() builds it into a stream-ordered temporary
 * on every call, these entry points let a caller with static weights (eval mode) do it once.
 * bevb200_spconv_packed_weight_bytes() returns 0 when (c_in, c_out, kernel_volume, precision)
 * has no tensor-core form (c_out in {16,32,64,128}, c_in <= 128, kernel_volume <= 27, precision
 * TF32X3, TF32 or BF16X3); such shapes go through bevb200_spconv_forward(), which falls back to
 * the fp32 SIMT kernel on the GPU.
 * Input channels are zero-padded to bevb200_spconv_padded_channels(c_in, precision) = 8, 16, 32,
 * 64 or 128 (SparseEncoder.conv_input: 5 -> 8): the packed image already contains the padding,
 * so the image packed for c_in is also the image for the padded count.  A caller that passes
 * features with fewer channels than that gets them copied into a stream-ordered padded temporary
 * on every call; passing rows that are already padded (and the padded count as c_in) avoids it. */
BEVB200_API int bevb200_spconv_padded_channels(int c_in, int precision);
BEVB200_API size_t bevb200_spconv_packed_weight_bytes(int c_in, int c_out, int kernel_volume, int precision);
BEVB200_API int bevb200_spconv_pack_weights(const float *weight, int c_in, int c_out, int kernel_volume,
                                int precision, float *packed, void *stream);
BEVB200_API int bevb200_spconv_forward_packed(const float *features, const float *packed_weight,
                                  const int32_t *nbr, int n_in, int n_out, int c_in, int c_out,
                                  int kernel_volume, const float *scale, const float *shift,
                                  const float *residual, int relu, int precision, float *out,
                                  void *stream);

/* Generation-6 tensor-core path (BEVB200_PREC_BF16X3), split operand images.
 * The tcgen05 kernel consumes a feature row as the bf16 image of its hi/lo split: per group of 16
 * channels 32 B of bf16 hi (= rn(x)) then 32 B of bf16 lo (= rn(x - hi)); c * 4 bytes per row, c a
 * multiple of 16.  A conv writes that image of ITS output from the epilogue (out_split), so a chain of
 * convs (SparseEncoder) splits each row once instead of once per (row, kernel offset) visit, and
 * bevb200_spconv_forward() -- the fp32-in / fp32-out drop-in for indice_conv_fp32 -- is
 * split_rows + forward_split.
 *   bevb200_spconv_split_channels(c)   channel count of the image (c rounded up to 16, <= 128; 0: none)
 *   bevb200_spconv_split_rows          fp32 rows [n, c_in] -> image [n, split_channels(c_in) * 4 B],
 *                                      zero padded; n_dev (nullable, device int32) caps n on the device
 *   bevb200_spconv_pack_split_weights  weight [K, c_in, c_out] fp32 -> the kernel's shared-memory image
 *                                      (bevb200_spconv_split_weight_bytes bytes; c_in is the REAL count)
 *   bevb200_spconv_forward_split       features_split [n_in, c_in * 4 B] (c_in = the padded count),
 *                                      nbr [K][nbr_stride] (rows >= n_out unused), n_out_dev (nullable):
 *                                      device-side row count <= n_out, read by the persistent kernel --
 *                                      no host round trip for the output count of a strided conv;
 *                                      out (fp32 rows) and / or out_split (image) are written, epilogue
 *                                      as bevb200_spconv_forward. */
BEVB200_API int bevb200_spconv_split_channels(int c_in);
BEVB200_API int bevb200_spconv_split_rows(const float *features, int n, const int32_t *n_dev, int c_in,
                              void *split, void *stream);
BEVB200_API size_t bevb200_spconv_split_weight_bytes(int c_in, int c_out, int kernel_volume);
BEVB200_API int bevb200_spconv_pack_split_weights(const float *weight, int c_in, int c_out, int kernel_volume,
                                      void *packed, void *stream);
BEVB200_API int bevb200_spconv_forward_split(const void *features_split, const void *packed_weight,
                                 const int32_t *nbr, long long nbr_stride, int n_in, int n_out,
                                 const int32_t *n_out_dev, int c_in, int c_out, int kernel_volume,
                                 const float *scale, const float *shift, const float *residual, int relu,
                                 float *out, void *out_split, void *stream);

/* Sparse convolution backward.  Replaces spconv::indiceConvBackward<float> (spconv_ops.h:363-456;
 * bound as sparse_conv_ext.indice_conv_backward_fp32):
 *     input_grad[j, :]  = sum_k out_grad[nbr_t[k, j], :] @ weight[k]^T     [n_in, c_in]
 *     weight_grad[k]    = sum_o features[nbr[k, o], :]^T (x) out_grad[o, :] [K, c_in, c_out]
 * nbr_t [K, n_in] is the transposed neighbour table from bevb200_rulebook_transpose()
 * (nbr_t[k, j] = the output row that input row j feeds through offset k, or -1).
 * The input gradient runs the forward implicit-GEMM kernel on (out_grad, W^T, nbr_t) in the
 * requested precision.  The weight gradient sums chunks of rows into per-chunk partials and adds those in
 * ascending chunk order: no atomics, bit-reproducible run to run -- on the tensor cores for BEVB200_PREC_BF16X3 /
 * _TF32 and c_in, c_out in {32, 64, 128} (spconv_wgrad_tc.cu: MN-major tcgen05 MMAs over the bf16 hi/lo
 * images of the gathered feature rows and the out-grad rows; BEVB200_WGRAD_TC=0 keeps the SIMT kernel), else SIMT.
 * workspace: bevb200_spconv_backward_workspace_bytes() (the transposed weights, the operand images, the partials). */
BEVB200_API int bevb200_rulebook_transpose(const int32_t *nbr, int kernel_volume, int n_out, int n_in,
                               int32_t *nbr_t, void *stream);
BEVB200_API size_t bevb200_spconv_backward_workspace_bytes(int n_in, int n_out, int c_in, int c_out, int kernel_volume);
BEVB200_API int bevb200_spconv_backward(const float *features, const float *weight, const float *out_grad,
                            const int32_t *nbr, const int32_t *nbr_t, int n_in, int n_out, int c_in,
                            int c_out, int kernel_volume, int precision, float *input_grad,
                            float *weight_grad, void *workspace, size_t workspace_bytes,
                            void *stream);

/* SparseConvTensor.dense() (structure.py:49-59) fused with SparseEncoder's
 * permute(0,1,4,2,3).view(N, C*D, H, W) (sparse_encoder.py:126-130):
 *   out[b, c*Z + z, x, y] = features[i, c] for indices[i] = (b, x, y, z); zero elsewhere.
 * With z_major == 0 the plain channels-first dense layout out[b, c, x, y, z] is written.
 * out must hold B*C*X*Y*Z floats; it is fully written (zero filled) by the call.
 * out_batch_stride (floats; 0 = C*X*Y*Z) lets `out` be a channel slice of a wider
 * [B, C_total, X, Y] buffer: the LiDAR half of the fuser input (fusers/conv.py:16) in place. */
BEVB200_API int bevb200_sparse_to_dense(const float *features, const int32_t *indices, int n, int c,
                            int batch_size, const int32_t *spatial_shape_host, int z_major,
                            long long out_batch_stride, float *out, void *stream);

/* ------------------------------------------------------------------------------------
 * SparseEncoder as one native, sync-free call (reference: the python loop of
 * mmdet3d/models/backbones/sparse_encoder.py:99-132 over spconv/conv.py:114-223, with the
 * BatchNorm1d / ReLU / residual modules of spconv/modules.py:127-139 and ops/sparse_block.py:94-110).
 *
 * A plan is a chain of convs (conv i reads the output of conv i-1): SubMConv3d or strided
 * SparseConv3d, each followed by y = acc * scale + shift (folded eval-mode BN / bias), an optional
 * residual add of the OUTPUT of an earlier conv (SparseBasicBlock identity) and an optional ReLU.
 * No row count ever returns to the host: buffers are sized by caps (level 0 = max_voxels; a strided
 * conv makes at most prod(ceil(k/s)) outputs per input row and one per output site; level_caps_host,
 * nullable, may tighten the caps of the levels >= 1), kernels read their counts from device memory,
 * so bevb200_encoder_forward() is a fixed sequence of ~45 launches that may be captured in a CUDA graph.
 * BEVB200_PREC_BF16X3 arithmetic (spconv generation 6).  Channel counts: 16 / 32 / 64 / 128 (the first
 * conv's c_in: anything <= 128).
 *   create            host-only object; destroy() frees it (and the CUDA events it may own)
 *   param_bytes       device bytes of the parameter buffer (packed weights, scales, shifts)
 *   set_conv          packs weight [K, c_in, c_out] fp32 and copies scale / shift [c_out] (nullable)
 *                     into the parameter buffer (device pointers, stream ordered)
 *   workspace_bytes   device scratch for (max_voxels, batch_size, caps)
 *   forward           voxel_features [max_voxels, in_channels] fp32, coors [max_voxels, 4] int32
 *                     (b, x, y, z), n_voxels_dev (nullable device int32: valid rows <= max_voxels);
 *                     dense_out [B, C*Z, X, Y] (out_batch_stride floats, 0 = dense) is fully written;
 *                     status_dev int32[1 + levels]: [0] != 0 when a cap truncated a level,
 *                     [1 + l] = rows of level l.  rulebook_stream (nullable): a second stream the
 *                     rulebooks are built on, forked from / joined to `stream` with events. */
typedef struct {
  int32_t c_in, c_out;
  int32_t ksize[3], stride[3], padding[3], dilation[3];
  int32_t subm;          /* 1: SubMConv3d (stride 1, padding k/2 forced, spconv_ops.h:74-83) */
  int32_t relu;
  int32_t residual_from; /* -1, or j < i: add the output of conv j before the ReLU */
} bevb200_encoder_conv_t;
typedef struct bevb200_encoder bevb200_encoder_t;
BEVB200_API int bevb200_encoder_create(int in_channels, const int32_t *sparse_shape_host,
                           const bevb200_encoder_conv_t *convs, int n_convs, bevb200_encoder_t **out);
BEVB200_API void bevb200_encoder_destroy(bevb200_encoder_t *enc);
BEVB200_API size_t bevb200_encoder_param_bytes(const bevb200_encoder_t *enc);
BEVB200_API int bevb200_encoder_num_levels(const bevb200_encoder_t *enc);
BEVB200_API int bevb200_encoder_output_shape(const bevb200_encoder_t *enc, int32_t *shape_out, int32_t *channels_out);
BEVB200_API int bevb200_encoder_set_conv(bevb200_encoder_t *enc, int conv, const float *weight, const float *scale,
                             const float *shift, void *params, size_t params_bytes, void *stream);
BEVB200_API int bevb200_encoder_level_caps(const bevb200_encoder_t *enc, int max_voxels, int batch_size,
                               const int32_t *level_caps_host, int32_t *caps_out);
BEVB200_API size_t bevb200_encoder_workspace_bytes(const bevb200_encoder_t *enc, int max_voxels, int batch_size,
                                       const int32_t *level_caps_host);
BEVB200_API int bevb200_encoder_forward(bevb200_encoder_t *enc, const void *params, const float *voxel_features,
                            const int32_t *coors, int max_voxels, const int32_t *n_voxels_dev,
                            int batch_size, const int32_t *level_caps_host, float *dense_out,
                            long long out_batch_stride, int32_t *status_dev, void *workspace,
                            size_t workspace_bytes, void *stream, void *rulebook_stream);

#ifdef __cplusplus
}
#endif
#endif /* BEVFUSION_B200_H_ */
