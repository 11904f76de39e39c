/*
 * gbops.h -- C ABI of libgbops.so: B200 (sm_100a) kernels for GraspBalance's point-cloud operator hot path.
 *
 * This is the drop-in boundary.  Every entry point takes raw DEVICE pointers, explicit sizes and an explicit
 * cudaStream_t (passed as void*), launches asynchronously on that stream and returns a cudaError_t as int
 * (0 = success; invalid arguments return cudaErrorInvalidValue = 1).  Nothing is retained between calls and no
 * entry point ever calls exit() -- the reference's launchers do (cuda_utils.h:38-47); callers raise instead.
 *
 * Each function names the reference launcher it replaces (paths relative to the GraspBalance tree):
 *   "A" = PointNet/_ext_src           (python module pointnet2._ext,        bindings.cpp:12-27)
 *   "B" = pointnet2_batch/src         (python module pointnet2_batch_cuda,  pointnet2_api.cpp:10-24)
 *   "C" = KNN/Pytorch_CUDA_KNN        (python module KNN._C,                vision.cpp:3-5)
 *
 * Index layouts, dtypes, tie order and floating-point contraction follow the reference kernels exactly
 * (DESIGN.md "Semantics"): indices and masks are bit-exact, fp32 forward values are bit-exact, the scatter-add
 * backward passes agree to fp32 summation-order tolerance.
 */
#ifndef GBOPS_H_
#define GBOPS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GBOPS_ABI_VERSION 1

typedef void *gb_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define GB_API __attribute__((visibility("default")))
#else
#define GB_API
#endif

/* FPS variant: GB_FPS_A = pointnet2._ext semantics (points with |p|^2 <= 1e-3 are skipped, tie order of a
 * 512-thread block: sampling_gpu.cu:74-178); GB_FPS_B = pointnet2_batch_cuda semantics (no skip, tie order of a
 * 1024-thread block: pointnet2_batch/src/sampling_gpu.cu:73-181). */
enum { GB_FPS_A = 0, GB_FPS_B = 1 };

GB_API int gb_abi_version(void);
GB_API const char *gb_error_string(int err);

/* A: furthest_point_sampling_kernel_wrapper (sampling_gpu.cu:180-234); B: furthest_point_sampling_kernel_launcher
 * (pointnet2_batch/src/sampling_gpu.cu:183-220).
 * xyz [b,n,3] f32; idx [b,m] i32 (fully overwritten); temp [b,n] f32 or NULL.  When temp is given it is read as the
 * initial running min-distance (both reference callers fill it with 1e10) and receives the final values, as in the
 * reference; when NULL, 1e10 is used and nothing is written back. */
GB_API int gb_fps(const float *xyz, float *temp, int *idx, int b, int n, int m, int variant, gb_stream_t stream);

/* FPS + the gather_operation every caller runs on its result (pointnet2_modules.py:151-158 `new_xyz = gather_operation(
 * xyz_flipped, inds)`), SURVEY 8f-2: as gb_fps, and new_xyz [b, m, 3] = xyz[b, idx[b, j]] written by the same launch. */
GB_API int gb_fps_xyz(const float *xyz, float *temp, int *idx, float *new_xyz, int b, int n, int m, int variant,
               gb_stream_t stream);

/* gb_fps_xyz (new_xyz may be NULL: gb_fps) with a footprint hint: at most max_cluster (0 = automatic, else 1..16) CTAs per
 * scene, as far as a scene still fits their registers.  The automatic shape minimises the latency of the m - 1 dependent
 * rounds; a caller that samples the NEXT step's clouds beside the current step's kernels prefers fewer, fuller CTAs, which
 * leave whole SMs free.  Same picks whatever the hint. */
GB_API int gb_fps_xyz_hint(const float *xyz, float *temp, int *idx, float *new_xyz, int b, int n, int m, int variant, int max_cluster,
                    gb_stream_t stream);

/* Segmented FPS: nseg independent point sets of different sizes packed in one array, one launch -- the per-object loop of
 * ObjectBalanceSampling (TrainModel/modules.py:186-213 calls furthest_point_sample once per object of every scene).
 * seg [nseg,4] i32 on the DEVICE (16-byte aligned) = (first point, points, samples, first output slot) per segment; idx and
 * the optional new_xyz are indexed by output slot; indices are segment-local and bit-identical to gb_fps on the segment
 * alone.  max_n / max_m = the largest `points` / `samples` entry; segments of more than 10240 points are refused. */
GB_API int gb_fps_segments(const float *xyz, const int *seg, int *idx, float *new_xyz, int nseg, int max_n, int max_m, int variant,
                    gb_stream_t stream);

/* A: gather_points_kernel_wrapper (sampling_gpu.cu:27-35); B: gather_points_kernel_launcher_fast (:21-34).
 * points [b,c,n], idx [b,m] -> out [b,c,m]. */
GB_API int gb_gather_fwd(const float *points, const int *idx, float *out, int b, int c, int n, int m, gb_stream_t stream);

/* A: gather_points_grad_kernel_wrapper (sampling_gpu.cu:54-62); B: gather_points_grad_kernel_launcher_fast (:50-63).
 * grad_out [b,c,m], idx [b,m]; ACCUMULATES into grad_points [b,c,n] (callers pass zeros, as the reference's do). */
GB_API int gb_gather_bwd(const float *grad_out, const int *idx, float *grad_points, int b, int c, int n, int m,
                  gb_stream_t stream);

/* A: query_ball_point_kernel_wrapper (ball_query_gpu.cu:46-54); B: ball_query_kernel_launcher_fast (:45-58).
 * new_xyz [b,m,3], xyz [b,n,3] -> idx [b,m,nsample] i32: the first nsample indices k (ascending) with
 * d2(k) < radius*radius, unfilled slots = first hit, all zero when there is no hit.  Every slot is written (the
 * reference relies on a zero-filled output instead). */
GB_API int gb_ball_query(const float *new_xyz, const float *xyz, int *idx, int b, int n, int m, float radius, int nsample,
                  gb_stream_t stream);

/* A only: query_cylinder_point_kernel_wrapper (cylinder_query_gpu.cu:89-101).  rot [b,m,9] row-major. */
GB_API int gb_cylinder_query(const float *new_xyz, const float *xyz, const float *rot, int *idx, int b, int n, int m,
                      float radius, float hmin, float hmax, int nsample, gb_stream_t stream);

/* The depth loop of GraspWidthGrouping.forward (TrainModel/modules.py:104-113 calls CylinderQueryAndGroup once per hmax of
 * hmax_list with the same seeds, rotations, radius and hmin): ndepth (1..4) nested cylinders per seed in ONE scan.
 * hmax = HOST array of ndepth floats.  idx [b, m, ndepth, nsample] i32: idx[:, :, d, :] is bit-identical to what
 * gb_cylinder_query(..., hmax[d], ...) writes. */
GB_API int gb_cylinder_query_multi(const float *new_xyz, const float *xyz, const float *rot, int *idx, int b, int n, int m,
                            float radius, float hmin, const float *hmax, int ndepth, int nsample, gb_stream_t stream);

/* The four GraspWidthGrouping modules of GraspPoseStage2_seed_features_multi_scale.forward (TrainModel/graspbalance.py:
 * 104-107) see the same seeds, rotations, hmin and depths and differ in the cylinder radius only: nradii x ndepth (1..4
 * each) nested cylinders per seed in ONE scan.  radii, hmax = HOST arrays.  idx [nradii, b, m, ndepth, nsample] i32:
 * idx[k, :, :, d, :] is bit-identical to what gb_cylinder_query(..., radii[k], hmin, hmax[d], ...) writes. */
GB_API int gb_cylinder_query_multi_radius(const float *new_xyz, const float *xyz, const float *rot, int *idx, int b, int n, int m,
                                   const float *radii, int nradii, float hmin, const float *hmax, int ndepth, int nsample,
                                   gb_stream_t stream);

/* A: group_points_kernel_wrapper (group_points_gpu.cu:51-65); B: group_points_kernel_launcher_fast (:58-70).
 * points [b,c,n], idx [b,npoints,nsample] -> out [b,c,npoints,nsample]. */
GB_API int gb_group_fwd(const float *points, const int *idx, float *out, int b, int c, int n, int npoints, int nsample,
                 gb_stream_t stream);

/* A: group_points_grad_kernel_wrapper (group_points_gpu.cu:92-101); B: group_points_grad_kernel_launcher_fast (:24-37).
 * ACCUMULATES into grad_points [b,c,n]. */
GB_API int gb_group_bwd(const float *grad_out, const int *idx, float *grad_points, int b, int c, int n, int npoints,
                 int nsample, gb_stream_t stream);
/* Same sum, but grad_points is fully OVERWRITTEN: what module A's group_points_grad (group_points.cpp:50-75) returns from
 * its own torch::zeros output, without the zero fill and the read-back of an accumulating launch. */
GB_API int gb_group_bwd_set(const float *grad_out, const int *idx, float *grad_points, int b, int c, int n, int npoints,
                     int nsample, gb_stream_t stream);

/* Strided variants for fused grouper modules (pointnet2_utils.py:178-207 builds cat([grouped_xyz, grouped_features])):
 * the C grouped rows of a scene are written to / read from a tensor whose scenes are out_scene_stride floats apart
 * (>= c*npoints*nsample, multiple of 4 for the vector paths), so features can be grouped straight into channels 3.. of
 * the [b, 3+c, npoints, nsample] result and their gradient read from that slice without a copy.  `out` / `grad_out` point
 * at the first of the c rows of scene 0. */
GB_API int gb_group_fwd_strided(const float *points, const int *idx, float *out, int b, int c, int n, int npoints, int nsample,
                         long long out_scene_stride, gb_stream_t stream);
GB_API int gb_group_bwd_strided(const float *grad_out, const int *idx, float *grad_points, int b, int c, int n, int npoints,
                         int nsample, long long grad_out_scene_stride, int overwrite, gb_stream_t stream);

/* The grouped-coordinate part of QueryAndGroup / CylinderQueryAndGroup (pointnet2_utils.py:178-190, 281-291; group.py:167-176)
 * in one pass: out[b, :, j, k] = ((xyz[b, idx[b,j,k]] - new_xyz[b,j]) * scale) . R[b,j]   (3 rows per scene, scenes
 * out_scene_stride floats apart).  xyz [b,n,3] (no transposed copy needed), new_xyz [b,m,3], idx [b,m,nsample],
 * rot [b,m,9] row-major or NULL (no rotation); use_scale = 0 skips the multiply (scale = fp32 reciprocal of the radius
 * when normalize_xyz is set, as ATen evaluates `grouped_xyz /= radius`). */
GB_API int gb_group_xyz(const float *xyz, const float *new_xyz, const int *idx, const float *rot, float *out, int b, int n, int m,
                 int nsample, float scale, int use_scale, long long out_scene_stride, gb_stream_t stream);

/* A whole QueryAndGroup tail in ONE launch: the grouped coordinates of gb_group_xyz (no rotation) and the grouped features of
 * gb_group_fwd_strided over the same idx -- pointnet2_utils.py:178-207 (variant A: out_xyz = rows 0..2 and out_feat = rows 3..
 * of the [b,3+c,npoints,nsample] result, both strides (3+c)*npoints*nsample) and group.py:167-179 (variant B: two tensors,
 * strides 3*npoints*nsample and c*npoints*nsample).  xyz [b,n,3], new_xyz [b,npoints,3], points [b,c,n].  Every CTA of the
 * feature kernel first writes its share of the coordinate rows; shapes the staged feature kernel does not take (unaligned,
 * rows beyond shared memory) run as the two launches.  Results are bit-identical to the two entry points. */
GB_API int gb_group_xyz_feat(const float *xyz, const float *new_xyz, const int *idx, float *out_xyz, long long xyz_scene_stride,
                      float scale, int use_scale, const float *points, float *out_feat, long long feat_scene_stride, int b, int c,
                      int n, int npoints, int nsample, gb_stream_t stream);

/* grouping_operation + max over nsample in one pass: PointnetSAModuleVotes_WOMLP.forward (PointNet/
 * pointnet2_modules.py:324-335) and the 'max' pooling of PointnetSAModuleVotes (:173-175) when no MLP sits between the
 * grouping and the pooling -- group_points_kernel (group_points_gpu.cu:17-36) + F.max_pool2d without the [b,c,npoints,nsample]
 * tensor.  points [b,c,n], idx [b,npoints,nsample] -> out [b,c,npoints] = max_k points[b,c,idx[b,j,k]] (ATen's rule: first
 * maximum in k order, NaN propagates); arg [b,c,npoints] i32 = source index of the maximum, or NULL.  Returns
 * cudaErrorNotSupported (801) for rows that do not fit shared memory four at a time (callers run the two steps). */
GB_API int gb_group_max_fwd(const float *points, const int *idx, float *out, int *arg, int b, int c, int n, int npoints, int nsample,
                     gb_stream_t stream);
/* Its backward: ACCUMULATES grad_out [b,c,npoints] into grad_points [b,c,n] at arg. */
GB_API int gb_group_max_bwd(const float *grad_out, const int *arg, float *grad_points, int b, int c, int n, int npoints,
                     gb_stream_t stream);

/* A: three_nn_kernel_wrapper (interpolate_gpu.cu:66-73); B: three_nn_kernel_launcher_fast (:62-81).
 * unknown [b,n,3], known [b,m,3] -> dist2 [b,n,3] f32 (SQUARED), idx [b,n,3] i32. */
GB_API int gb_three_nn(const float *unknown, const float *known, float *dist2, int *idx, int b, int n, int m,
                gb_stream_t stream);

/* three_nn + the inverse-distance weights every caller derives from it (pointnet2_modules.py:413-416, upsampling.py:69-72,
 * graspbalance.py:37-41), one launch: dist [b,n,3] = sqrt(dist2), idx [b,n,3], weight [b,n,3] =
 * (1/(dist+1e-8)) / sum_k (1/(dist_k+1e-8)), each op rounded as the torch op it replaces. */
GB_API int gb_three_nn_weights(const float *unknown, const float *known, float *dist, int *idx, float *weight, int b, int n, int m,
                        gb_stream_t stream);

/* A: three_interpolate_kernel_wrapper (interpolate_gpu.cu:108-116); B: three_interpolate_kernel_launcher_fast (:106-124).
 * points [b,c,m], idx/weight [b,n,3] -> out [b,c,n]. */
GB_API int gb_three_interp_fwd(const float *points, const int *idx, const float *weight, float *out, int b, int c, int m,
                        int n, gb_stream_t stream);

/* A: three_interpolate_grad_kernel_wrapper (interpolate_gpu.cu:150-159); B: ..._grad_kernel_launcher_fast (:151-168).
 * ACCUMULATES into grad_points [b,c,m]. */
GB_API int gb_three_interp_bwd(const float *grad_out, const int *idx, const float *weight, float *grad_points, int b, int c,
                        int n, int m, gb_stream_t stream);
/* Same sum, grad_points fully OVERWRITTEN (module A's three_interpolate_grad, interpolate.cpp:76-104, zero-fills its own
 * output). */
GB_API int gb_three_interp_bwd_set(const float *grad_out, const int *idx, const float *weight, float *grad_points, int b, int c,
                            int n, int m, gb_stream_t stream);

/* C: knn_device (knn.cu:217-263) looped over the batch as knn.h:31-38 does.
 * ref [b,dim,nref], query [b,dim,nquery] (channel first) -> idx [b,k,nquery] int64, 1-BASED, ascending (dist, index).
 * Needs no distance-matrix scratch (the reference allocates nref*nquery floats, knn.h:29).  Requires 1 <= k <= nref
 * and k <= 1024. */
GB_API int gb_knn(const float *ref, const float *query, int64_t *idx, int b, int dim, int nref, int nquery, int k,
           gb_stream_t stream);

/* three_nn + inverse-distance weights + three_interpolate in ONE launch -- what PointnetFPModule.forward
 * (PointNet/pointnet2_modules.py:413-420), upsampling.three_interpolation (ModifiedNetTools/upsampling.py:67-74) and the seed
 * up-sampling of TrainModel/graspbalance.py:37-41 compute with a neighbour search, five elementwise passes and a gather.
 * unknown [b,n,3], known [b,m,3] (1 <= m <= 4096), feats [b,c,m] -> out [b,c,n] (16-byte aligned).  idx_out / weight_out
 * [b,n,3]: both NULL (nothing but `out` is written) or both given (three_interpolate's backward reads them).  Returns
 * cudaErrorNotSupported (801) for shapes it does not take (callers use the two-launch path).  Bit-identical to
 * gb_three_nn_weights followed by gb_three_interp_fwd. */
GB_API int gb_three_interpolation(const float *unknown, const float *known, const float *feats, float *out, int *idx_out,
                           float *weight_out, int b, int c, int n, int m, gb_stream_t stream);

/* collision_detector.ModelFreeCollisionDetector.detect's grasp x point occupancy test (collision_detector.py:23-41,55),
 * all fp64.  points [np,3]; T [g,3]; R [g,3,3] row-major; thr [g,10] = the per-grasp half-space thresholds
 *   {-h/2, h/2, d-fl, d, -(w/2+fw), -w/2, w/2+fw, w/2, d-fl-fw, d-fl-fw-approach}
 * evaluated by the caller with the reference's numpy expressions; counts [g,6] int64 =
 *   {global, left, right, bottom, shifting, inner} mask sums (fully overwritten). */
GB_API int gb_collision_counts(const double *points, int np, const double *T, const double *R, const double *thr, int g,
                        int64_t *counts, gb_stream_t stream);

/* Host-buffer convenience entry (what a non-torch host language would bind): copies the inputs to the device,
 * runs gb_collision_counts and copies the counts back, synchronising the stream.  All pointers are HOST pointers. */
GB_API int gb_collision_counts_host(const double *points, int np, const double *T, const double *R, const double *thr, int g,
                             int64_t *counts);

/* gb_collision_counts for several scenes in one launch: scene z = rows scene_off[z] .. scene_off[z+1] of the
 * packed points [sum N', 3] f64 (scene_off: nscenes + 1 i64 on the DEVICE), g grasps per scene: T [nscenes,g,3], R
 * [nscenes,g,3,3], thr [nscenes,g,10] -> counts [nscenes,g,6] i64.  max_np = the largest scene. */
GB_API int gb_collision_counts_batched(const double *points, const long long *scene_off, int nscenes, int max_np, const double *T,
                                const double *R, const double *thr, int g, int64_t *counts, gb_stream_t stream);

/* ModelFreeCollisionDetector.detect (collision_detector.py:16-64) without a host round trip: thresholds
 * (:26-35), occupancy counts (:23-41,55), volumes and IoUs (:43-47,55-63) and the masks (:48,56) in one call.
 * points [np,3] f64 = the detector's down-sampled scene.  grasps = g rows of `dtype` (0 = f32, 1 = f64) that lie
 * row_stride elements apart and hold the translation (3), the row-major rotation (9), height, depth and width at column
 * offsets oT, oR, oH, oD, oW: a packed [g,15] array, or the [Ns,17] array pred_decode emits (TrainModel/graspbalance.py:
 * 187-190: score, width, height, depth, rotation, centre, object id) -- so network output feeds the test in place.
 * params = 7 HOST doubles {finger_width, finger_length, max(approach_dist, finger_width), voxel_size**3, 2*finger_width,
 * collision_thresh, empty_thresh}: the Python floats of the reference expressions.  Thresholds and volumes are evaluated in
 * the grasp arrays' dtype with the constants cast to it first, left to right, as numpy does; compares and IoU divisions
 * are fp64.  mask [g] u8 (required) = global_iou > collision_thresh; optional: empty [g] u8 = inner_count / inner_volume <
 * empty_thresh, ious [5,g] f64 = {global, left, right, bottom, shifting}, counts [g,6] i64.  Bit-identical to the reference
 * for f32 and f64 grasp groups. */
GB_API int gb_collision_detect(const double *points, int np, const void *grasps, int dtype, int g, int row_stride, int oT, int oR,
                        int oH, int oD, int oW, const double *params, unsigned char *mask, unsigned char *empty, double *ious,
                        int64_t *counts, gb_stream_t stream);

/* The per-voxel means of the voxel down-sampling in ModelFreeCollisionDetector.__init__ (collision_detector.py:11-14, open3d
 * PointCloud.voxel_down_sample).  points [n,3] f64; order [n] i64 = point indices grouped by voxel, input order
 * kept inside a voxel; seg [v+1] i64 = first position of every voxel in `order`.  out [v,3] f64 = the sequential fp64 sum
 * of a voxel's points divided by their count (open3d's running sum), one thread per voxel. */
GB_API int gb_voxel_means(const double *points, const long long *order, const long long *seg, double *out, int v,
                   gb_stream_t stream);

/* Tuning knobs for benchmarking sweeps (never change results).  Unknown keys return cudaErrorInvalidValue.
 *   "fps_cluster"  0 = auto, else 1/2/4/8/16 CTAs per scene
 *   "fps_threads"  0 = auto, else 256/512/1024
 *   "group_split"  0 = auto, else output splits per (scene, channel chunk)
 *   "group_mode"   bit 0: plain (not streaming) stores in group fwd; bit 1: generic fwd kernel; bit 2: atomic backward;
 *                  bit 3: never the one-launch row kernel of the few-channel backward
 *   "interp_mode"  bit 0: plain stores; bit 1: generic fwd kernel; bit 2: atomic backward
 *   "query_qpw"    0 = auto, else 1/2/4 queries per warp in the full-scan query kernel
 *   "query_mode"   0 = auto, 1 = full scan only, 2 = always build the cell grid
 *   "scatter_cc"   0 = auto, else 1/2/4 channels per CTA in the sorted backward (warps per block in the warp-private one)
 *   "scatter_mode" bit 0: no dense sorted backward; bit 2: targets in index order; bit 3: never the warp-private backward;
 *                  bit 4: the warp-private backward for any number of tasks
 *   "priv_vl" / "priv_cw" / "priv_split"  warp-private backward: positions per lane and load (1/2/4), channels per warp
 *                  (2/4), warps sharing a task (1/2/4); 0 = auto
 *   "priv_rows"    warp-private backward, nsample 8 / 16: 1 = channel planes share a row, 2 = several rows per unit; 0 = auto
 */
GB_API int gb_set_tuning(const char *key, int value);
GB_API int gb_get_tuning(const char *key, int *value);

/* Number of kernel launches issued through this library since load (for bench.py's gpu_launches). */
GB_API uint64_t gb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GBOPS_H_ */
