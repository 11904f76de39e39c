/*
 * gb_oracle.c -- CPU restatement of the reference's point-cloud operator hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker / the timed CPU baseline.
 * The product (graspbalance_b200/) never links, imports or calls it.
 *
 * Every function restates one reference CUDA kernel literally (same loop order, same strict / non
 * strict compares, same tie behaviour) in plain C, with the floating-point contraction that nvcc
 * applies to the reference source written out with explicit fmaf() -- the pattern was read from the
 * SASS of the reference kernels compiled for sm_100a (oracle/build_ref.py; see DESIGN.md "FMA
 * patterns"):
 *
 *     a*a + b*b + c*c          ->  fmaf(c,c, fmaf(a,a, b*b))        (the SECOND product is the plain FMUL)
 *     r0*x + r3*y + r6*z       ->  fmaf(r6,z, fmaf(r0,x, r3*y))
 *     y*y + z*z                ->  fmaf(y,y, z*z)
 *     ssd += t*t  (KNN)        ->  ssd = fmaf(t,t,ssd)               (sequential over dim)
 *
 * Compile with -ffp-contract=off so that the compiler adds no contraction of its own.
 *
 * Parity pinning: tests/test_oracle_vs_ref_gpu.py runs the reference's own compiled extensions
 * (oracle/_ref/*.so) beside this file on the GPU box; tests/golden/ holds vectors produced by those
 * extensions (tests/golden/make_golden_gpu.py) and by the reference's numpy collision detector
 * (tests/golden/make_golden_collision.py) that the CPU-only suite checks this file against.
 *
 * Reference citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define GBO_API __attribute__((visibility("default")))

/* PointNet/_ext_src/include/cuda_utils.h:21-27 (cap 512) and pointnet2_batch/src/cuda_utils.h:10-14 (cap 1024). */
GBO_API int gbo_opt_n_threads(int work_size, int cap) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int t = 1 << pow_2;
  if (t > cap) t = cap;
  if (t < 1) t = 1;
  return t;
}

/* ---- tiny pthread parallel-for over independent work items (scenes / queries / grasps).  The split never
 * changes a result: every item is computed by exactly the sequential code of the reference thread it restates. ---- */
static int g_threads = 0;
GBO_API void gbo_set_num_threads(int t) { g_threads = t; }
GBO_API int gbo_num_threads(void) {
  if (g_threads > 0) return g_threads;
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n < 1 ? 1 : (n > 256 ? 256 : (int)n);
}
typedef void (*gbo_body)(void *ctx, long item);
typedef struct { gbo_body fn; void *ctx; long total; long chunk; volatile long next; } gbo_job;
static void *gbo_worker(void *p) {
  gbo_job *j = (gbo_job *)p;
  for (;;) {
    long lo = __sync_fetch_and_add(&j->next, j->chunk);
    if (lo >= j->total) break;
    long hi = lo + j->chunk < j->total ? lo + j->chunk : j->total;
    for (long i = lo; i < hi; ++i) j->fn(j->ctx, i);
  }
  return 0;
}
static void gbo_parallel_for(long total, long chunk, gbo_body fn, void *ctx) {
  int nt = gbo_num_threads();
  if (chunk < 1) chunk = 1;
  if (nt > (total + chunk - 1) / chunk) nt = (int)((total + chunk - 1) / chunk);
  gbo_job job = {fn, ctx, total, chunk, 0};
  if (nt <= 1) { gbo_worker(&job); return; }
  pthread_t th[256];
  int started = 0;
  for (int t = 0; t < nt - 1; ++t)
    if (pthread_create(&th[started], 0, gbo_worker, &job) == 0) ++started;
  gbo_worker(&job);
  for (int t = 0; t < started; ++t) pthread_join(th[t], 0);
}

/* squared distance as nvcc contracts (a-b)*(a-b) + ... in ball_query_gpu.cu:31-32, sampling_gpu.cu:108-109,
 * interpolate_gpu.cu:38 */
static inline float sqdist3(float dx, float dy, float dz) { return fmaf(dz, dz, fmaf(dx, dx, dy * dy)); }

/* one context type for all ops: the work-item bodies below are the per-thread code of the reference kernels */
typedef struct {
  const float *a, *b, *c;   /* float inputs */
  const int *ia;            /* int input */
  float *fo;                /* float output */
  int *io;                  /* int output */
  int64_t *lo;              /* int64 output */
  const double *da, *db, *dc, *dd;
  int B, N, M, C, S, K, V;
  float f0, f1, f2;
} gbo_ctx;

/* ---------------------------------------------------------------------------------------------
 * F1 / F1b  furthest point sampling.
 * PointNet/_ext_src/src/sampling_gpu.cu:64-178 (variant 0: norm skip, block cap 512, temp init 1e10 by
 * sampling.cpp:78-80) and pointnet2_batch/src/sampling_gpu.cu:67-181 (variant 1: no skip, cap 1024).
 * Literal emulation: per-"thread" strided scan, then the shared-memory tree with __update's tie rule.
 * temp: [b,n] scratch, caller-filled (1e10 in both wrappers).
 * ------------------------------------------------------------------------------------------- */
static void fps_item(void *p, long bi) {
  gbo_ctx *x = (gbo_ctx *)p;
  const int n = x->N, m = x->M, variant = x->V, bs = x->K;
  const float *dataset = x->a + (size_t)bi * n * 3;
  float *tmp = x->fo + (size_t)bi * n;
  int *out = x->io + (size_t)bi * m;
  float *dists = (float *)malloc(sizeof(float) * bs);
  int *dists_i = (int *)malloc(sizeof(int) * bs);
  int old = 0;
  out[0] = old;
  for (int j = 1; j < m; ++j) {
    const float x1 = dataset[old * 3 + 0], y1 = dataset[old * 3 + 1], z1 = dataset[old * 3 + 2];
    for (int tid = 0; tid < bs; ++tid) {
      int besti = 0;
      float best = -1.f;
      for (int k = tid; k < n; k += bs) {
        const float x2 = dataset[k * 3 + 0], y2 = dataset[k * 3 + 1], z2 = dataset[k * 3 + 2];
        if (variant == 0) {
          const float mag = fmaf(z2, z2, fmaf(x2, x2, y2 * y2));
          if ((double)mag <= 1e-3) continue; /* sampling_gpu.cu:105-106, double compare */
        }
        const float d = sqdist3(x2 - x1, y2 - y1, z2 - z1);
        const float d2 = fminf(d, tmp[k]);
        tmp[k] = d2;
        besti = d2 > best ? k : besti;
        best = d2 > best ? d2 : best;
      }
      dists[tid] = best;
      dists_i[tid] = besti;
    }
    for (int s = bs / 2; s >= 1; s >>= 1) {
      for (int tid = 0; tid < s; ++tid) { /* __update(dists, dists_i, tid, tid + s) */
        const float v1 = dists[tid], v2 = dists[tid + s];
        const int i1 = dists_i[tid], i2 = dists_i[tid + s];
        dists[tid] = fmaxf(v1, v2);
        dists_i[tid] = v2 > v1 ? i2 : i1;
      }
    }
    old = dists_i[0];
    out[j] = old;
  }
  free(dists);
  free(dists_i);
}
GBO_API void gbo_fps(const float *xyz, float *temp, int *idxs, int b, int n, int m, int variant) {
  if (m <= 0) return;
  gbo_ctx x = {0};
  x.a = xyz; x.fo = temp; x.io = idxs; x.N = n; x.M = m; x.V = variant;
  x.K = gbo_opt_n_threads(n, variant == 0 ? 512 : 1024);
  gbo_parallel_for(b, 1, fps_item, &x);
}

/* F2 gather: sampling_gpu.cu:13-25 / batch :8-19.  points [b,c,n], idx [b,m] -> out [b,c,m] */
static void gather_fwd_item(void *p, long row) {
  gbo_ctx *x = (gbo_ctx *)p;
  const int i = (int)(row / x->C);
  const float *src = x->a + (size_t)row * x->N;
  const int *ii = x->ia + (size_t)i * x->M;
  float *dst = x->fo + (size_t)row * x->M;
  for (int j = 0; j < x->M; ++j) dst[j] = src[ii[j]];
}
GBO_API void gbo_gather_fwd(const float *points, const int *idx, float *out, int b, int c, int n, int m) {
  gbo_ctx x = {0};
  x.a = points; x.ia = idx; x.fo = out; x.C = c; x.N = n; x.M = m;
  gbo_parallel_for((long)b * c, 4, gather_fwd_item, &x);
}

/* F2 gather grad: sampling_gpu.cu:39-52 / batch :36-48.  accumulates (atomicAdd) into grad_points [b,c,n] */
static void gather_bwd_item(void *p, long row) {
  gbo_ctx *x = (gbo_ctx *)p;
  const int i = (int)(row / x->C);
  const float *g = x->a + (size_t)row * x->M;
  const int *ii = x->ia + (size_t)i * x->M;
  float *dst = x->fo + (size_t)row * x->N;
  for (int j = 0; j < x->M; ++j) dst[ii[j]] += g[j];
}
GBO_API void gbo_gather_bwd(const float *grad_out, const int *idx, float *grad_points, int b, int c, int n, int m) {
  gbo_ctx x = {0};
  x.a = grad_out; x.ia = idx; x.fo = grad_points; x.C = c; x.N = n; x.M = m;
  gbo_parallel_for((long)b * c, 4, gather_bwd_item, &x);
}

/* F3 ball query: ball_query_gpu.cu:9-44 (A) == pointnet2_batch/src/ball_query_gpu.cu:10-42 (B).
 * idx [b,m,nsample] must be zero-filled by the caller (A: torch::zeros ball_query.cpp:24-26; B: group.py:136). */
static void ball_item(void *p, long qi) {
  gbo_ctx *x = (gbo_ctx *)p;
  const int n = x->N, nsample = x->S;
  const long bi = qi / x->M;
  const float radius2 = x->f0 * x->f0;
  const float *pts = x->b + (size_t)bi * n * 3;
  const float *q = x->a + (size_t)qi * 3;
  int *o = x->io + (size_t)qi * nsample;
  const float new_x = q[0], new_y = q[1], new_z = q[2];
  for (int k = 0, cnt = 0; k < n && cnt < nsample; ++k) {
    const float d2 = sqdist3(new_x - pts[k * 3 + 0], new_y - pts[k * 3 + 1], new_z - pts[k * 3 + 2]);
    if (d2 < radius2) {
      if (cnt == 0)
        for (int l = 0; l < nsample; ++l) o[l] = k;
      o[cnt] = k;
      ++cnt;
    }
  }
}
GBO_API void gbo_ball_query(const float *new_xyz, const float *xyz, int *idx, int b, int n, int m, float radius,
                            int nsample) {
  gbo_ctx x = {0};
  x.a = new_xyz; x.b = xyz; x.io = idx; x.N = n; x.M = m; x.S = nsample; x.f0 = radius;
  gbo_parallel_for((long)b * m, 16, ball_item, &x);
}

/* F4 cylinder query: cylinder_query_gpu.cu:20-78.  rot [b,m,9] row-major. */
static void cyl_item(void *p, long qi) {
  gbo_ctx *x = (gbo_ctx *)p;
  const int n = x->N, nsample = x->S;
  const long bi = qi / x->M;
  const float radius2 = x->f0 * x->f0, hmin = x->f1, hmax = x->f2;
  const float *pts = x->b + (size_t)bi * n * 3;
  const float *q = x->a + (size_t)qi * 3;
  const float *r = x->c + (size_t)qi * 9;
  int *o = x->io + (size_t)qi * nsample;
  const float new_x = q[0], new_y = q[1], new_z = q[2];
  for (int k = 0, cnt = 0; k < n && cnt < nsample; ++k) {
    const float xx = pts[k * 3 + 0] - new_x, y = pts[k * 3 + 1] - new_y, z = pts[k * 3 + 2] - new_z;
    const float x_rot = fmaf(r[6], z, fmaf(r[0], xx, r[3] * y));
    const float y_rot = fmaf(r[7], z, fmaf(r[1], xx, r[4] * y));
    const float z_rot = fmaf(r[8], z, fmaf(r[2], xx, r[5] * y));
    const float d2 = fmaf(y_rot, y_rot, z_rot * z_rot);
    if (d2 < radius2 && x_rot > hmin && x_rot < hmax) {
      if (cnt == 0)
        for (int l = 0; l < nsample; ++l) o[l] = k;
      o[cnt] = k;
      ++cnt;
    }
  }
}
GBO_API void gbo_cylinder_query(const float *new_xyz, const float *xyz, const float *rot, int *idx, int b, int n,
                                int m, float radius, float hmin, float hmax, int nsample) {
  gbo_ctx x = {0};
  x.a = new_xyz; x.b = xyz; x.c = rot; x.io = idx; x.N = n; x.M = m; x.S = nsample;
  x.f0 = radius; x.f1 = hmin; x.f2 = hmax;
  gbo_parallel_for((long)b * m, 16, cyl_item, &x);
}

/* F5 group: group_points_gpu.cu:17-44 / batch :40-55.  points [b,c,n], idx [b,npoints,nsample] -> out [b,c,npoints,nsample] */
static void group_fwd_item(void *p, long row) {
  gbo_ctx *x = (gbo_ctx *)p;
  const size_t per = (size_t)x->M * x->S;
  const float *src = x->a + (size_t)row * x->N;
  const int *ii = x->ia + (size_t)(row / x->C) * per;
  float *dst = x->fo + (size_t)row * per;
  for (size_t e = 0; e < per; ++e) dst[e] = src[ii[e]];
}
GBO_API void gbo_group_fwd(const float *points, const int *idx, float *out, int b, int c, int n, int npoints,
                           int nsample) {
  gbo_ctx x = {0};
  x.a = points; x.ia = idx; x.fo = out; x.C = c; x.N = n; x.M = npoints; x.S = nsample;
  gbo_parallel_for((long)b * c, 1, group_fwd_item, &x);
}

/* F5 group grad: group_points_gpu.cu:69-90 / batch :9-22.  accumulates into grad_points [b,c,n] */
static void group_bwd_item(void *p, long row) {
  gbo_ctx *x = (gbo_ctx *)p;
  const size_t per = (size_t)x->M * x->S;
  float *dst = x->fo + (size_t)row * x->N;
  const int *ii = x->ia + (size_t)(row / x->C) * per;
  const float *g = x->a + (size_t)row * per;
  for (size_t e = 0; e < per; ++e) dst[ii[e]] += g[e];
}
GBO_API void gbo_group_bwd(const float *grad_out, const int *idx, float *grad_points, int b, int c, int n, int npoints,
                           int nsample) {
  gbo_ctx x = {0};
  x.a = grad_out; x.ia = idx; x.fo = grad_points; x.C = c; x.N = n; x.M = npoints; x.S = nsample;
  gbo_parallel_for((long)b * c, 1, group_bwd_item, &x);
}

/* F6 three_nn: interpolate_gpu.cu:14-64 / batch :16-59.  Returns SQUARED distances (the Python wrapper sqrt()s). */
static void three_nn_item(void *p, long ui) {
  gbo_ctx *x = (gbo_ctx *)p;
  const int m = x->M;
  const long bi = ui / x->N;
  const float *u = x->a + (size_t)ui * 3;
  const float *kn = x->b + (size_t)bi * m * 3;
  const float ux = u[0], uy = u[1], uz = u[2];
  double best1 = 1e40, best2 = 1e40, best3 = 1e40;
  int besti1 = 0, besti2 = 0, besti3 = 0;
  for (int k = 0; k < m; ++k) {
    const float d = sqdist3(ux - kn[k * 3 + 0], uy - kn[k * 3 + 1], uz - kn[k * 3 + 2]);
    if (d < best1) {
      best3 = best2; besti3 = besti2;
      best2 = best1; besti2 = besti1;
      best1 = d; besti1 = k;
    } else if (d < best2) {
      best3 = best2; besti3 = besti2;
      best2 = d; besti2 = k;
    } else if (d < best3) {
      best3 = d; besti3 = k;
    }
  }
  float *od = x->fo + (size_t)ui * 3;
  int *oi = x->io + (size_t)ui * 3;
  od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
  oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
}
GBO_API void gbo_three_nn(const float *unknown, const float *known, float *dist2, int *idx, int b, int n, int m) {
  gbo_ctx x = {0};
  x.a = unknown; x.b = known; x.fo = dist2; x.io = idx; x.N = n; x.M = m;
  gbo_parallel_for((long)b * n, 64, three_nn_item, &x);
}

/* F7 three_interpolate: interpolate_gpu.cu:77-106 / batch :84-104.
 * points [b,c,m], idx/weight [b,n,3] -> out [b,c,n];  p1*w1 + p2*w2 + p3*w3 contracts to fmaf(p3,w3,fmaf(p1,w1,p2*w2)). */
static void interp_fwd_item(void *p, long row) {
  gbo_ctx *x = (gbo_ctx *)p;
  const int n = x->N;
  const long i = row / x->C;
  const float *pt = x->a + (size_t)row * x->M;
  const int *ii = x->ia + (size_t)i * n * 3;
  const float *w = x->b + (size_t)i * n * 3;
  float *o = x->fo + (size_t)row * n;
  for (int j = 0; j < n; ++j)
    o[j] = fmaf(pt[ii[j * 3 + 2]], w[j * 3 + 2], fmaf(pt[ii[j * 3 + 0]], w[j * 3 + 0], pt[ii[j * 3 + 1]] * w[j * 3 + 1]));
}
GBO_API void gbo_three_interp_fwd(const float *points, const int *idx, const float *weight, float *out, int b, int c,
                                  int m, int n) {
  gbo_ctx x = {0};
  x.a = points; x.ia = idx; x.b = weight; x.fo = out; x.C = c; x.M = m; x.N = n;
  gbo_parallel_for((long)b * c, 1, interp_fwd_item, &x);
}

/* F7 grad: interpolate_gpu.cu:121-148 / batch :127-149.  accumulates into grad_points [b,c,m] */
static void interp_bwd_item(void *p, long row) {
  gbo_ctx *x = (gbo_ctx *)p;
  const int n = x->N;
  const long i = row / x->C;
  const float *g = x->a + (size_t)row * n;
  const int *ii = x->ia + (size_t)i * n * 3;
  const float *w = x->b + (size_t)i * n * 3;
  float *o = x->fo + (size_t)row * x->M;
  for (int j = 0; j < n; ++j) {
    o[ii[j * 3 + 0]] += g[j] * w[j * 3 + 0];
    o[ii[j * 3 + 1]] += g[j] * w[j * 3 + 1];
    o[ii[j * 3 + 2]] += g[j] * w[j * 3 + 2];
  }
}
GBO_API void gbo_three_interp_bwd(const float *grad_out, const int *idx, const float *weight, float *grad_points,
                                  int b, int c, int n, int m) {
  gbo_ctx x = {0};
  x.a = grad_out; x.ia = idx; x.b = weight; x.fo = grad_points; x.C = c; x.M = m; x.N = n;
  gbo_parallel_for((long)b * c, 1, interp_bwd_item, &x);
}

/* K1 KNN, CUDA-path semantics: KNN/Pytorch_CUDA_KNN/cuda/knn.cu:36-101 (distance, ssd = fmaf(t,t,ssd) over dim,
 * t = ref - query) and :113-176 (per-query insertion sort, literal).  ref [b,dim,R], query [b,dim,Q] (channel first),
 * idx [b,k,Q] int64, 1-based.  Requires 1 <= k <= R (as the reference does). */
static void knn_item(void *p, long qi) {
  gbo_ctx *x = (gbo_ctx *)p;
  const int dim = x->C, R = x->N, Q = x->M, k = x->K;
  const long bi = qi / Q;
  const int q = (int)(qi % Q);
  const float *A = x->a + (size_t)bi * dim * R;
  const float *Bq = x->b + (size_t)bi * dim * Q;
  float *dist = (float *)malloc(sizeof(float) * (size_t)R);
  int64_t *ind = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
  for (int r = 0; r < R; ++r) {
    float ssd = 0.f;
    for (int d = 0; d < dim; ++d) {
      const float t = A[(size_t)d * R + r] - Bq[(size_t)d * Q + q];
      ssd = fmaf(t, t, ssd);
    }
    dist[r] = ssd;
  }
  /* cuInsertionSort, one column */
  float max_dist = dist[0];
  ind[0] = 1;
  for (int l = 1; l < k; ++l) {
    const float curr_dist = dist[l];
    if (curr_dist < max_dist) {
      int i = l - 1;
      for (int a = 0; a < l - 1; ++a)
        if (dist[a] > curr_dist) { i = a; break; }
      for (int j = l; j > i; --j) { dist[j] = dist[j - 1]; ind[j] = ind[j - 1]; }
      dist[i] = curr_dist;
      ind[i] = l + 1;
    } else {
      ind[l] = l + 1;
    }
    max_dist = dist[l];
  }
  for (int l = k; l < R; ++l) {
    const float curr_dist = dist[l];
    if (curr_dist < max_dist) {
      int i = k - 1;
      for (int a = 0; a < k - 1; ++a)
        if (dist[a] > curr_dist) { i = a; break; }
      for (int j = k - 1; j > i; --j) { dist[j] = dist[j - 1]; ind[j] = ind[j - 1]; }
      dist[i] = curr_dist;
      ind[i] = l + 1;
      max_dist = dist[k - 1];
    }
  }
  for (int l = 0; l < k; ++l) x->lo[((size_t)bi * k + l) * Q + q] = ind[l];
  free(dist);
  free(ind);
}
GBO_API void gbo_knn(const float *ref, const float *query, int64_t *idx, int b, int dim, int R, int Q, int k) {
  gbo_ctx x = {0};
  x.a = ref; x.b = query; x.lo = idx; x.C = dim; x.N = R; x.M = Q; x.K = k;
  gbo_parallel_for((long)b * Q, 8, knn_item, &x);
}

/* C1 collision occupancy counts: collision_detector.py:23-41,55.  All fp64.
 * points [np,3]; T [g,3]; R [g,3,3] row-major; thr [g,10] per-grasp thresholds computed by the caller with the
 * reference's own numpy expressions:
 *   0: -heights/2   1: heights/2   2: depths-fl   3: depths   4: -(widths/2+fw)   5: -widths/2
 *   6: widths/2+fw  7: widths/2    8: depths-fl-fw            9: depths-fl-fw-approach_dist
 * targets = (p - T) @ R: numpy hands this to OpenBLAS dgemm, whose x86 kernels evaluate the K=3 dot product as
 *   t_j = fma(d2,R[2][j], fma(d1,R[1][j], d0*R[0][j]))     (fma_mode 1; measured bit-identical to np.matmul on
 * 9.6e5 samples, tests/golden/make_golden_collision.py); fma_mode 0 is the unfused left-to-right sum.
 * counts [g,6] int64: global, left, right, bottom, shifting, inner. */
static void collision_item(void *p, long gi) {
  gbo_ctx *x = (gbo_ctx *)p;
  const int np = x->N, fma_mode = x->V;
  const double *points = x->da;
  const double *t = x->db + (size_t)gi * 3, *r = x->dc + (size_t)gi * 9, *h = x->dd + (size_t)gi * 10;
  int64_t cg = 0, cl = 0, cr = 0, cb = 0, cs = 0, ci = 0;
  for (int pi = 0; pi < np; ++pi) {
    const double d0 = points[pi * 3 + 0] - t[0], d1 = points[pi * 3 + 1] - t[1], d2 = points[pi * 3 + 2] - t[2];
    double tx, ty, tz;
    if (fma_mode) {
      tx = fma(d2, r[6], fma(d1, r[3], d0 * r[0]));
      ty = fma(d2, r[7], fma(d1, r[4], d0 * r[1]));
      tz = fma(d2, r[8], fma(d1, r[5], d0 * r[2]));
    } else {
      tx = (d0 * r[0] + d1 * r[3]) + d2 * r[6];
      ty = (d0 * r[1] + d1 * r[4]) + d2 * r[7];
      tz = (d0 * r[2] + d1 * r[5]) + d2 * r[8];
    }
    const int m1 = (tz > h[0]) & (tz < h[1]);
    const int m2 = (tx > h[2]) & (tx < h[3]);
    const int m3 = ty > h[4];
    const int m4 = ty < h[5];
    const int m5 = ty < h[6];
    const int m6 = ty > h[7];
    const int m7 = (tx <= h[2]) & (tx > h[8]);
    const int m8 = (tx <= h[8]) & (tx > h[9]);
    const int left = m1 & m2 & m3 & m4, right = m1 & m2 & m5 & m6;
    const int bottom = m1 & m3 & m5 & m7, shifting = m1 & m3 & m5 & m8;
    cg += left | right | bottom | shifting;
    cl += left; cr += right; cb += bottom; cs += shifting;
    ci += m1 & m2 & (!m4) & (!m6);
  }
  int64_t *o = x->lo + (size_t)gi * 6;
  o[0] = cg; o[1] = cl; o[2] = cr; o[3] = cb; o[4] = cs; o[5] = ci;
}
GBO_API void gbo_collision_counts(const double *points, int np, const double *T, const double *R, const double *thr,
                                  int g, int fma_mode, int64_t *counts) {
  gbo_ctx x = {0};
  x.da = points; x.db = T; x.dc = R; x.dd = thr; x.N = np; x.V = fma_mode; x.lo = counts;
  gbo_parallel_for(g, 4, collision_item, &x);
}
