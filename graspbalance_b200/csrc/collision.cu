// collision.cu -- batched grasp x point gripper-volume occupancy test (fp64).
//
// Replaces the numpy body of ModelFreeCollisionDetector.detect (collision_detector.py:23-41,55), which materialises a
// [G,N,3] fp64 array of transformed points (491 MB at G=1024, N=20000) and ten [G,N] boolean masks on the host.  Here one
// WARP owns a grasp (translation, rotation and the ten half-space thresholds live in registers); the scene is cut into
// packs of 32 consecutive points with precomputed bounds, the warp skips the packs its gripper cannot reach and, for the
// others, every lane transforms one point and keeps six integer counters that are warp-reduced at the end.  Nothing but the
// six counts per grasp ever leaves the SM.
//
// Arithmetic (SURVEY.md A.7 and DESIGN.md "collision rounding"): d = p - T in fp64; t_j = fma(d2,R[2][j], fma(d1,R[1][j],
// d0*R[0][j])) -- the evaluation order of the OpenBLAS dgemm kernel numpy.matmul dispatches to (bit-identical on 9.6e5
// samples); the thresholds are computed on the host with the reference's own numpy expressions, so every compare sees
// the same two doubles as the reference.
#include "common.cuh"

namespace gb {

constexpr int kColWarps = 4;
constexpr int kColPack = 32;  // points per pack: one per lane

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

// Axis-aligned bounds of every pack of 32 consecutive points: bounds [nscenes, pps, 6] = {min x,y,z, max x,y,z}.  One warp per
// pack.  scene_off = nullptr: one scene of np points.  (fmin / fmax skip NaN coordinates: such points satisfy no mask.)
__global__ void collision_bounds_kernel(const double *__restrict__ points, int np, const long long *__restrict__ scene_off, int pps,
                                        double *__restrict__ bounds) {
  const int lane = threadIdx.x & 31;
  const int pack = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (scene_off) {
    const long long o0 = scene_off[blockIdx.y];
    points += 3 * o0;
    np = (int)(scene_off[blockIdx.y + 1] - o0);
  }
  if (pack * kColPack >= np) return;
  const int e = pack * kColPack + lane;
  const bool ok = e < np;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  const double x = ok ? points[(size_t)e * 3] : inf, y = ok ? points[(size_t)e * 3 + 1] : inf, z = ok ? points[(size_t)e * 3 + 2] : inf;
  const double lo0 = warp_min(x), lo1 = warp_min(y), lo2 = warp_min(z);
  const double hi0 = warp_max(ok ? x : -inf), hi1 = warp_max(ok ? y : -inf), hi2 = warp_max(ok ? z : -inf);
  if (lane < 6) {
    const double v = lane == 0 ? lo0 : lane == 1 ? lo1 : lane == 2 ? lo2 : lane == 3 ? hi0 : lane == 4 ? hi1 : hi2;
    bounds[((size_t)blockIdx.y * pps + pack) * 6 + lane] = v;
  }
}

// One warp per grasp.  A point can only count when its gripper-frame coordinates lie inside the box the ten thresholds
// span; for an orthonormal rotation that puts it inside a sphere around the grasp centre, so packs whose bounds miss the
// sphere (with a 2e-4 relative margin, far above the rounding of either side) are skipped without touching their points:
// a gripper reaches ~0.1 m in a scene of ~0.7 m, and the detector's down-sampled cloud is stored in voxel-key order
// (voxel_down_sample_gpu), so consecutive points are neighbours and most packs are skipped.  Rotations that are not
// orthonormal to 1e-6 (or contain NaN) test every pack.  The arithmetic per tested point is unchanged, so the counts are
// those of the exhaustive test.
__global__ void __launch_bounds__(kColWarps * 32) collision_kernel(const double *__restrict__ points, int np, const double *__restrict__ T,
                                                                   const double *__restrict__ R, const double *__restrict__ thr, int g,
                                                                   unsigned long long *__restrict__ counts,
                                                                   const long long *__restrict__ scene_off, const double *__restrict__ bounds,
                                                                   int pps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (scene_off) {  // batched: blockIdx.y = scene; its points are rows scene_off[y] .. scene_off[y+1] of the packed array
    const long long o0 = scene_off[blockIdx.y];
    points += 3 * o0;
    np = (int)(scene_off[blockIdx.y + 1] - o0);
    const size_t go = (size_t)blockIdx.y * g;
    T += go * 3, R += go * 9, thr += go * 10, counts += go * 6;
  }
  bounds += (size_t)blockIdx.y * pps * 6;
  const int gi = blockIdx.x * kColWarps + warp;
  if (gi >= g) return;  // warps are independent
  const double t0 = T[(size_t)gi * 3], t1 = T[(size_t)gi * 3 + 1], t2 = T[(size_t)gi * 3 + 2];
  double r[9], h[10];
#pragma unroll
  for (int e = 0; e < 9; ++e) r[e] = R[(size_t)gi * 9 + e];
#pragma unroll
  for (int e = 0; e < 10; ++e) h[e] = thr[(size_t)gi * 10 + e];
  // reach of the gripper box: x in (h9, h3), |y| < h6 (= w/2 + fw), |z| < h1 (= height / 2)
  const double xr = fmax(fabs(h[9]), fabs(h[3])), yr = fmax(fabs(h[4]), fabs(h[6])), zr = fmax(fabs(h[0]), fabs(h[1]));
  double rho2 = (xr * xr + yr * yr + zr * zr) * 1.0002 + 1e-12;
  bool ortho = true;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = a; b < 3; ++b) {
      const double dot = r[a] * r[b] + r[3 + a] * r[3 + b] + r[6 + a] * r[6 + b];  // (R^T R)_{ab}
      ortho = ortho && fabs(dot - (a == b ? 1.0 : 0.0)) < 1e-6;
    }
  if (!ortho || !(rho2 == rho2)) rho2 = __longlong_as_double(0x7ff0000000000000LL);  // test every pack

  int cg = 0, cl = 0, cr = 0, cb = 0, cs = 0, ci = 0;
  const int npacks = (np + kColPack - 1) / kColPack;
  for (int pbase = 0; pbase < npacks; pbase += 32) {
    const int pk = pbase + lane;
    bool hit = false;
    if (pk < npacks) {
      const double *bb = bounds + (size_t)pk * 6;
      const double dx = fmax(fmax(bb[0] - t0, t0 - bb[3]), 0.0), dy = fmax(fmax(bb[1] - t1, t1 - bb[4]), 0.0),
                   dz = fmax(fmax(bb[2] - t2, t2 - bb[5]), 0.0);
      hit = !(dx * dx + dy * dy + dz * dz > rho2);  // NaN bounds or a NaN centre count as a hit
    }
    unsigned hits = __ballot_sync(0xffffffffu, hit);
    while (hits) {
      const int e = (pbase + __ffs(hits) - 1) * kColPack + lane;
      hits &= hits - 1;
      const bool valid = e < np;
      const size_t es = valid ? e : 0;
      const double d0 = points[es * 3] - t0, d1 = points[es * 3 + 1] - t1, d2 = points[es * 3 + 2] - t2;
      const double tz = __fma_rn(d2, r[8], __fma_rn(d1, r[5], __dmul_rn(d0, r[2])));
      const bool m1 = valid && (tz > h[0]) && (tz < h[1]);
      if (!__any_sync(0xffffffffu, m1)) continue;  // every mask needs m1
      const double tx = __fma_rn(d2, r[6], __fma_rn(d1, r[3], __dmul_rn(d0, r[0])));
      const double ty = __fma_rn(d2, r[7], __fma_rn(d1, r[4], __dmul_rn(d0, r[1])));
      const bool m2 = (tx > h[2]) && (tx < h[3]);
      const bool m3 = ty > h[4];
      const bool m4 = ty < h[5];
      const bool m5 = ty < h[6];
      const bool m6 = ty > h[7];
      const bool m7 = (tx <= h[2]) && (tx > h[8]);
      const bool m8 = (tx <= h[8]) && (tx > h[9]);
      const bool left = m1 && m2 && m3 && m4, right = m1 && m2 && m5 && m6;
      const bool bottom = m1 && m3 && m5 && m7, shifting = m1 && m3 && m5 && m8;
      cg += (left || right || bottom || shifting) ? 1 : 0;
      cl += left ? 1 : 0;
      cr += right ? 1 : 0;
      cb += bottom ? 1 : 0;
      cs += shifting ? 1 : 0;
      ci += (m1 && m2 && !m4 && !m6) ? 1 : 0;
    }
  }
  cg = __reduce_add_sync(0xffffffffu, cg);
  cl = __reduce_add_sync(0xffffffffu, cl);
  cr = __reduce_add_sync(0xffffffffu, cr);
  cb = __reduce_add_sync(0xffffffffu, cb);
  cs = __reduce_add_sync(0xffffffffu, cs);
  ci = __reduce_add_sync(0xffffffffu, ci);
  if (lane < 6) {
    const int v = lane == 0 ? cg : lane == 1 ? cl : lane == 2 ? cr : lane == 3 ? cb : lane == 4 ? cs : ci;
    counts[(size_t)gi * 6 + lane] = (unsigned long long)v;
  }
}

// Per-voxel means for the voxel down-sampling of ModelFreeCollisionDetector.__init__ (collision_detector.py:11-14, open3d's
// PointCloud.voxel_down_sample): the points of a voxel, listed in INPUT order, are summed sequentially in fp64 by one
// thread and divided by their count -- the running sum open3d keeps per voxel, so the means are bit-identical to the host
// restatement.  order [n] = point indices grouped by voxel (stable sort by voxel key), seg [v+1] = first position of
// every voxel in `order`.
__global__ void voxel_means_kernel(const double *__restrict__ points, const long long *__restrict__ order,
                                   const long long *__restrict__ seg, double *__restrict__ out, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v) return;
  const long long s0 = seg[i], s1 = seg[i + 1];
  double sx = 0.0, sy = 0.0, sz = 0.0;
  for (long long k = s0; k < s1; ++k) {
    const double *p = points + 3 * order[k];
    sx = __dadd_rn(sx, p[0]), sy = __dadd_rn(sy, p[1]), sz = __dadd_rn(sz, p[2]);
  }
  const double cnt = (double)(s1 - s0);
  out[3 * (size_t)i] = sx / cnt, out[3 * (size_t)i + 1] = sy / cnt, out[3 * (size_t)i + 2] = sz / cnt;
}


// ---- detect() entirely on the device ------------------------------------------------------------------------
// collision_detector.py:16-64 around the occupancy test is elementwise IEEE arithmetic on the grasp arrays: the ten
// half-space thresholds (:26-35), the gripper volumes in voxels (:43-46,55) and count / (volume + 1e-6) > thresh (:47-48,
// 56-63).  numpy evaluates them in the DTYPE OF THE GRASP ARRAYS -- float32 for a graspnetAPI GraspGroup built from network
// output, float64 otherwise -- with the Python floats cast to that dtype first, left to right, no contraction; only the
// comparisons against the fp64 transformed points and the final integer / float divisions are fp64.  F below is that dtype;
// every operation is an explicit round-to-nearest intrinsic.
template <typename F>
struct ColArith;
template <>
struct ColArith<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <>
struct ColArith<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

struct ColParams {
  double fw, fl, ad, v3, two_fw, collision_thresh, empty_thresh;
};

// grasp rows: `stride` elements apart, translation / rotation (row-major 3x3) / height / depth / width at the given column
// offsets -- [G,15] packed by the host wrapper, or the [Ns,17] array pred_decode emits (graspbalance.py:187-190).
// Writes T64 [g,3], R64 [g,9], thr [g,10] and den [g,5] = {volume + 1e-6, lr + 1e-6, bottom + 1e-6, shifting + 1e-6, inner}.
template <typename F>
__global__ void collision_prepare_kernel(const F *__restrict__ rows, int g, int stride, int oT, int oR, int oH, int oD, int oW,
                                         ColParams p, double *__restrict__ T64, double *__restrict__ R64, double *__restrict__ thr,
                                         double *__restrict__ den) {
  typedef ColArith<F> A;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g) return;
  const F *r = rows + (size_t)i * stride;
#pragma unroll
  for (int e = 0; e < 3; ++e) T64[(size_t)i * 3 + e] = (double)r[oT + e];
#pragma unroll
  for (int e = 0; e < 9; ++e) R64[(size_t)i * 9 + e] = (double)r[oR + e];
  const F h = r[oH], d = r[oD], w = r[oW];
  const F fw = (F)p.fw, fl = (F)p.fl, ad = (F)p.ad, v3 = (F)p.v3, two_fw = (F)p.two_fw, eps = (F)1e-6, two = (F)2;
  const F h2 = A::div(h, two), w2 = A::div(w, two);
  const F w2fw = A::add(w2, fw), dfl = A::sub(d, fl), dflfw = A::sub(dfl, fw);
  double *t = thr + (size_t)i * 10;
  t[0] = (double)(-h2), t[1] = (double)h2, t[2] = (double)dfl, t[3] = (double)d, t[4] = (double)(-w2fw), t[5] = (double)(-w2),
  t[6] = (double)w2fw, t[7] = (double)w2, t[8] = (double)dflfw, t[9] = (double)A::sub(dflfw, ad);
  const F lr = A::div(A::mul(A::mul(h, fl), fw), v3);
  const F wide = A::mul(h, A::add(w, two_fw));
  const F bottom = A::div(A::mul(wide, fw), v3), shifting = A::div(A::mul(wide, ad), v3);
  const F volume = A::add(A::add(A::mul(lr, two), bottom), shifting);
  double *q = den + (size_t)i * 5;
  q[0] = (double)A::add(volume, eps), q[1] = (double)A::add(lr, eps), q[2] = (double)A::add(bottom, eps),
  q[3] = (double)A::add(shifting, eps), q[4] = (double)A::div(A::mul(A::mul(h, fl), w), v3);
}

// counts [g,6] -> collision mask, optional empty mask, optional IoUs [5,g] (global, left, right, bottom, shifting)
__global__ void collision_finish_kernel(const unsigned long long *__restrict__ counts, const double *__restrict__ den, int g, ColParams p,
                                        unsigned char *__restrict__ mask, unsigned char *__restrict__ empty, double *__restrict__ ious) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g) return;
  const unsigned long long *c = counts + (size_t)i * 6;
  const double *q = den + (size_t)i * 5;
  const double giou = __ddiv_rn((double)(long long)c[0], q[0]);
  mask[i] = giou > p.collision_thresh ? 1 : 0;
  if (empty) empty[i] = __ddiv_rn((double)(long long)c[5], q[4]) < p.empty_thresh ? 1 : 0;
  if (ious) {
    ious[i] = giou;
    ious[(size_t)g + i] = __ddiv_rn((double)(long long)c[1], q[1]);
    ious[(size_t)2 * g + i] = __ddiv_rn((double)(long long)c[2], q[1]);
    ious[(size_t)3 * g + i] = __ddiv_rn((double)(long long)c[3], q[2]);
    ious[(size_t)4 * g + i] = __ddiv_rn((double)(long long)c[4], q[3]);
  }
}

// scene_off = nullptr: one scene of np points; else nscenes scenes of at most np points each
static int collision_launch(const double *points, int np, const long long *scene_off, int nscenes, const double *T, const double *R,
                            const double *thr, int g, int64_t *counts, cudaStream_t s) {
  if (np == 0) return (int)cudaMemsetAsync(counts, 0, (size_t)nscenes * g * 6 * sizeof(int64_t), s);
  const int pps = (np + kColPack - 1) / kColPack;
  double *bounds = nullptr;
  cudaError_t e = scratch_alloc((void **)&bounds, (size_t)nscenes * pps * 6 * sizeof(double), s);
  if (e != cudaSuccess) return (int)e;
  collision_bounds_kernel<<<dim3((pps + 7) / 8, nscenes), 256, 0, s>>>(points, np, scene_off, pps, bounds);
  count_launch();
  int rc = finish_launch();
  if (!rc) {
    collision_kernel<<<dim3((g + kColWarps - 1) / kColWarps, nscenes), kColWarps * 32, 0, s>>>(
        points, np, T, R, thr, g, reinterpret_cast<unsigned long long *>(counts), scene_off, bounds, pps);
    count_launch();
    rc = finish_launch();
  }
  cudaFreeAsync(bounds, s);
  return rc;
}

}  // namespace gb

extern "C" int gb_collision_counts(const double *points, int np, const double *T, const double *R, const double *thr, int g,
                                   int64_t *counts, gb_stream_t stream) {
  if (np < 0 || g < 0 || !T || !R || !thr || !counts || (np > 0 && !points)) return (int)cudaErrorInvalidValue;
  if (g == 0) return 0;
  return gb::collision_launch(points, np, nullptr, 1, T, R, thr, g, counts, (cudaStream_t)stream);
}

extern "C" int gb_collision_counts_host(const double *points, int np, const double *T, const double *R, const double *thr, int g,
                                        int64_t *counts) {
  if (np < 0 || g < 0 || !T || !R || !thr || !counts || (np > 0 && !points)) return (int)cudaErrorInvalidValue;
  if (g == 0) return 0;
  cudaStream_t s = nullptr;
  cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  if (e != cudaSuccess) return (int)e;
  const size_t bp = (size_t)np * 3 * sizeof(double), bt = (size_t)g * 3 * sizeof(double), br = (size_t)g * 9 * sizeof(double),
               bh = (size_t)g * 10 * sizeof(double), bc = (size_t)g * 6 * sizeof(int64_t);
  unsigned char *d = nullptr;
  const size_t total = ((bp + 255) & ~(size_t)255) + ((bt + 255) & ~(size_t)255) + ((br + 255) & ~(size_t)255) + ((bh + 255) & ~(size_t)255) + bc;
  e = cudaMalloc((void **)&d, total);
  if (e != cudaSuccess) { cudaStreamDestroy(s); return (int)e; }
  unsigned char *dp = d, *dt = dp + ((bp + 255) & ~(size_t)255), *dr = dt + ((bt + 255) & ~(size_t)255),
                *dh = dr + ((br + 255) & ~(size_t)255), *dc = dh + ((bh + 255) & ~(size_t)255);
  int rc = 0;
  if (bp) rc = (int)cudaMemcpyAsync(dp, points, bp, cudaMemcpyHostToDevice, s);
  if (!rc) rc = (int)cudaMemcpyAsync(dt, T, bt, cudaMemcpyHostToDevice, s);
  if (!rc) rc = (int)cudaMemcpyAsync(dr, R, br, cudaMemcpyHostToDevice, s);
  if (!rc) rc = (int)cudaMemcpyAsync(dh, thr, bh, cudaMemcpyHostToDevice, s);
  if (!rc) rc = gb::collision_launch((const double *)dp, np, nullptr, 1, (const double *)dt, (const double *)dr, (const double *)dh, g, (int64_t *)dc, s);
  if (!rc) rc = (int)cudaMemcpyAsync(counts, dc, bc, cudaMemcpyDeviceToHost, s);
  if (!rc) rc = (int)cudaStreamSynchronize(s);
  cudaFree(d);
  cudaStreamDestroy(s);
  return rc;
}


/* ModelFreeCollisionDetector.detect (collision_detector.py:16-64) without a host round trip.
 * points [np,3] f64 = the detector's down-sampled scene; grasps = g rows of `dtype` (0 = f32, 1 = f64), row_stride elements
 * apart, with the translation (3), row-major rotation (9), height, depth and width at column offsets oT, oR, oH, oD, oW
 * (a packed [g,15] array, or pred_decode's [Ns,17] array: graspbalance.py:187-190).  params (HOST, 7 doubles) = finger_width,
 * finger_length, max(approach_dist, finger_width), voxel_size**3, 2*finger_width, collision_thresh, empty_thresh -- the
 * Python floats of the reference.  mask [g] u8 (required); empty [g] u8, ious [5,g] f64, counts [g,6] i64: optional outputs.
 * Thresholds and volumes are evaluated in the grasp arrays' dtype exactly as numpy does; every output is bit-identical to
 * the reference's for f32 and f64 grasp groups. */
extern "C" int gb_collision_detect(const double *points, int np, const void *grasps, int dtype, int g, int row_stride, int oT, int oR,
                                   int oH, int oD, int oW, const double *params, unsigned char *mask, unsigned char *empty, double *ious,
                                   int64_t *counts, gb_stream_t stream) {
  if (np < 0 || g < 0 || (dtype != 0 && dtype != 1) || !params) return (int)cudaErrorInvalidValue;
  if (g == 0) return 0;
  if (!grasps || !mask || (np > 0 && !points) || row_stride <= 0) return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  gb::ColParams p = {params[0], params[1], params[2], params[3], params[4], params[5], params[6]};
  const size_t nd = (size_t)g * (3 + 9 + 10 + 5);
  double *scratch = nullptr;
  cudaError_t e = gb::scratch_alloc((void **)&scratch, (nd + (counts ? 0 : (size_t)g * 6)) * sizeof(double), s);
  if (e != cudaSuccess) return (int)e;
  double *T64 = scratch, *R64 = T64 + (size_t)g * 3, *thr = R64 + (size_t)g * 9, *den = thr + (size_t)g * 10;
  int64_t *cnt = counts ? counts : reinterpret_cast<int64_t *>(den + (size_t)g * 5);
  const int blocks = (g + 127) / 128;
  if (dtype == 0)
    gb::collision_prepare_kernel<float><<<blocks, 128, 0, s>>>((const float *)grasps, g, row_stride, oT, oR, oH, oD, oW, p, T64, R64, thr, den);
  else
    gb::collision_prepare_kernel<double><<<blocks, 128, 0, s>>>((const double *)grasps, g, row_stride, oT, oR, oH, oD, oW, p, T64, R64, thr, den);
  gb::count_launch();
  int rc = gb::finish_launch();
  if (!rc) rc = gb::collision_launch(points, np, nullptr, 1, T64, R64, thr, g, cnt, s);
  if (!rc) {
    gb::collision_finish_kernel<<<blocks, 128, 0, s>>>(reinterpret_cast<const unsigned long long *>(cnt), den, g, p, mask, empty, ious);
    gb::count_launch();
    rc = gb::finish_launch();
  }
  cudaFreeAsync(scratch, s);
  return rc;
}

/* Voxel means for the down-sampling step of ModelFreeCollisionDetector.__init__ (collision_detector.py:11-14): points [n,3]
 * f64, order [n] i64 = point indices grouped by voxel with the input order kept inside a voxel, seg [v+1] i64 = first
 * position of every voxel in `order`; out [v,3] f64 = sequential fp64 sum of the voxel's points divided by their count. */
extern "C" int gb_voxel_means(const double *points, const long long *order, const long long *seg, double *out, int v,
                              gb_stream_t stream) {
  if (v < 0) return (int)cudaErrorInvalidValue;
  if (v == 0) return 0;
  if (!points || !order || !seg || !out) return (int)cudaErrorInvalidValue;
  gb::voxel_means_kernel<<<(v + 127) / 128, 128, 0, (cudaStream_t)stream>>>(points, order, seg, out, v);
  gb::count_launch();
  return gb::finish_launch();
}

/* The occupancy test of several scenes in one launch: scene z has the rows scene_off[z] .. scene_off[z+1] of the
 * packed `points` [sum N', 3] f64 (scene_off: nscenes + 1 i64 on the DEVICE) and g grasps, T [nscenes, g, 3], R [nscenes, g,
 * 3, 3], thr [nscenes, g, 10]; counts [nscenes, g, 6] i64.  max_np = the largest scene (host knowledge: sizes the grid).
 * Per scene the counts equal gb_collision_counts on that scene alone. */
extern "C" int gb_collision_counts_batched(const double *points, const long long *scene_off, int nscenes, int max_np, const double *T,
                                           const double *R, const double *thr, int g, int64_t *counts, gb_stream_t stream) {
  if (nscenes < 0 || max_np < 0 || g < 0) return (int)cudaErrorInvalidValue;
  if (nscenes == 0 || g == 0) return 0;
  if (!scene_off || !T || !R || !thr || !counts || (max_np > 0 && !points) || nscenes > 65535) return (int)cudaErrorInvalidValue;
  return gb::collision_launch(points, max_np, scene_off, nscenes, T, R, thr, g, counts, (cudaStream_t)stream);
}
