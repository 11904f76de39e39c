// interp.cu -- three_interpolate forward and backward.
//
// Replaces three_interpolate_kernel / three_interpolate_grad_kernel (PointNet/_ext_src/src/interpolate_gpu.cu:77-159; one
// block per scene) and the *_kernel_fast pair (pointnet2_batch/src/interpolate_gpu.cu:84-168; one thread per output,
// idx/weight re-read for every channel, three sector-sized L2 gathers per output float).
//
// Forward, HBM-bound on its output (4*C*n bytes/scene): the m "known" feature rows of a chunk of channels are staged in
// shared memory interleaved four channels per point, so each of the three neighbours of a point costs one LDS.128 for
// four channels; a thread owns four consecutive points (three 128-bit loads each for idx and weight, read once per channel
// CHUNK) and writes one coalesced 128-bit streaming store per channel.  Value = fmaf(p3,w3, fmaf(p1,w1, p2*w2)), the
// contraction nvcc applies to the reference expression (SASS-checked), so the forward is bit-exact.
// Backward: the atomic-free sorted segmented sum of scatter.cu; fallback red.global.add.f32 of g*w_t into the [C,m]
// gradient rows (which stay in L2).
#include "common.cuh"

namespace gb {

// 256 threads, three CTAs per SM (registers capped at 85): B200, 20000 <- 1024, C = 256, 32 scenes: 230 us against 261 us
// with 512 threads (74 registers: one CTA per SM, 16 warps) and 300 us with 512 x 2 (64 registers, spills)
#ifndef GB_INTERP_THREADS
#define GB_INTERP_THREADS 256
#endif
#ifndef GB_INTERP_MINB
#define GB_INTERP_MINB 3
#endif
constexpr int kInterpThreads = GB_INTERP_THREADS;

// points [b,c,m]; idx, weight [b,n,3]; out [b,c,n]; n % 4 == 0.  CH channels per fill (multiple of 4).
__global__ void __launch_bounds__(kInterpThreads, GB_INTERP_MINB) interp_fwd_kernel(const float *__restrict__ points, const int *__restrict__ idx,
                                                                   const float *__restrict__ weight, float *__restrict__ out, int c,
                                                                   int m, int n4, int CH, int chunks, long long total, long long wpc,
                                                                   int streaming) {
  extern __shared__ __align__(16) float s_rows[];  // [CH/4][m] float4
  float4 *srow = reinterpret_cast<float4 *>(s_rows);
  const int tid = threadIdx.x;
  long long w = (long long)blockIdx.x * wpc;
  const long long wend = min(total, w + wpc);
  const int G = CH / 4;
  const size_t n = (size_t)n4 * 4;

  while (w < wend) {
    const long long pair = w / n4;
    const int q0 = (int)(w - pair * n4);
    const int q1 = (int)min((long long)n4, (long long)q0 + (wend - w));
    const int scene = (int)(pair / chunks), chunk = (int)(pair - (long long)scene * chunks);
    const int ch_base = chunk * CH;
    const int gcount = min(G, (c - ch_base + 3) / 4);

    __syncthreads();
    for (int g = 0; g < gcount; ++g) {
      const float *src = points + ((size_t)scene * c + ch_base + g * 4) * m;
      const int nv = min(4, c - (ch_base + g * 4));
      for (int i = tid; i < m; i += kInterpThreads) {
        float4 o;
        o.x = __ldg(src + i);
        o.y = nv > 1 ? __ldg(src + (size_t)m + i) : 0.f;
        o.z = nv > 2 ? __ldg(src + 2 * (size_t)m + i) : 0.f;
        o.w = nv > 3 ? __ldg(src + 3 * (size_t)m + i) : 0.f;
        srow[(size_t)g * m + i] = o;
      }
    }
    __syncthreads();

    const int *ip = idx + (size_t)scene * n * 3;
    const float *wp = weight + (size_t)scene * n * 3;
    for (int q = q0 + tid; q < q1; q += kInterpThreads) {
      // 4 points x 3 neighbours: 12 ints and 12 floats, contiguous
      int id[12];
      float ww[12];
      {
        const int4 a = ld_nc_i4(ip + (size_t)q * 12), b = ld_nc_i4(ip + (size_t)q * 12 + 4), d = ld_nc_i4(ip + (size_t)q * 12 + 8);
        id[0] = a.x, id[1] = a.y, id[2] = a.z, id[3] = a.w, id[4] = b.x, id[5] = b.y, id[6] = b.z, id[7] = b.w;
        id[8] = d.x, id[9] = d.y, id[10] = d.z, id[11] = d.w;
        const float4 u = ld_nc_na_f4(wp + (size_t)q * 12), v = ld_nc_na_f4(wp + (size_t)q * 12 + 4), x = ld_nc_na_f4(wp + (size_t)q * 12 + 8);
        ww[0] = u.x, ww[1] = u.y, ww[2] = u.z, ww[3] = u.w, ww[4] = v.x, ww[5] = v.y, ww[6] = v.z, ww[7] = v.w;
        ww[8] = x.x, ww[9] = x.y, ww[10] = x.z, ww[11] = x.w;
      }
      for (int g = 0; g < gcount; ++g) {
        const float4 *row = srow + (size_t)g * m;
        float o[4][4];  // [channel][point]
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float4 a = row[id[p * 3]], b = row[id[p * 3 + 1]], d = row[id[p * 3 + 2]];
          const float w1 = ww[p * 3], w2 = ww[p * 3 + 1], w3 = ww[p * 3 + 2];
          o[0][p] = __fmaf_rn(d.x, w3, __fmaf_rn(a.x, w1, __fmul_rn(b.x, w2)));
          o[1][p] = __fmaf_rn(d.y, w3, __fmaf_rn(a.y, w1, __fmul_rn(b.y, w2)));
          o[2][p] = __fmaf_rn(d.z, w3, __fmaf_rn(a.z, w1, __fmul_rn(b.z, w2)));
          o[3][p] = __fmaf_rn(d.w, w3, __fmaf_rn(a.w, w1, __fmul_rn(b.w, w2)));
        }
        const int ch0 = ch_base + g * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (ch0 + e < c) {
            float *dst = out + ((size_t)scene * c + ch0 + e) * n + (size_t)q * 4;
            const float4 v = make_float4(o[e][0], o[e][1], o[e][2], o[e][3]);
            if (streaming) st_cs_f4(dst, v);
            else *reinterpret_cast<float4 *>(dst) = v;
          }
        }
      }
    }
    w += (q1 - q0);
  }
}

__global__ void interp_fwd_generic_kernel(const float *__restrict__ points, const int *__restrict__ idx, const float *__restrict__ weight,
                                          float *__restrict__ out, int c, int m, size_t n, size_t total) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t row = e / n, j = e - row * n;
    const size_t scene = row / c;
    const int *ip = idx + (scene * n + j) * 3;
    const float *wp = weight + (scene * n + j) * 3;
    const float *p = points + row * m;
    out[e] = __fmaf_rn(__ldg(p + ip[2]), wp[2], __fmaf_rn(__ldg(p + ip[0]), wp[0], __fmul_rn(__ldg(p + ip[1]), wp[1])));
  }
}

// grad_out [b,c,n]; idx, weight [b,n,3]; grad_points [b,c,m] (+=)
__global__ void interp_bwd_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx, const float *__restrict__ weight,
                                  float *__restrict__ grad_points, int c, int m, size_t n, size_t total) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t row = e / n, j = e - row * n;
    const size_t scene = row / c;
    const int *ip = idx + (scene * n + j) * 3;
    const float *wp = weight + (scene * n + j) * 3;
    const float g = __ldg(grad_out + e);
    float *dst = grad_points + row * m;
    atomicAdd(dst + __ldg(ip), __fmul_rn(g, __ldg(wp)));
    atomicAdd(dst + __ldg(ip + 1), __fmul_rn(g, __ldg(wp + 1)));
    atomicAdd(dst + __ldg(ip + 2), __fmul_rn(g, __ldg(wp + 2)));
  }
}


// ---- fused feature propagation: three_nn -> weights -> three_interpolate in ONE launch ---------------------
// What PointnetFPModule.forward (pointnet2_modules.py:413-420), upsampling.three_interpolation (upsampling.py:67-74) and the
// seed up-sampling of graspbalance.py:37-41 compute with a neighbour search, five elementwise torch passes and a gather:
// dist/idx/weight [B,n,3] never reach global memory unless the caller asks for idx and weight (the backward needs them).
// A CTA owns a range of PR unknown points of one scene:
//   phase 1  every thread finds the three nearest known points of its PR/512 points (known coordinates in shared memory as
//            float4, the arithmetic and tie order of three_nn_kernel<true>: bit-identical indices and weights) and parks
//            idx and weight in shared memory;
//   phase 2  for every chunk of CH channels the m known feature rows are staged interleaved four channels per point (as
//            interp_fwd_kernel does) and the range's outputs are written, one 128-bit streaming store per channel and four
//            points.  The rows come from L2 (C*m*4 bytes per scene, re-read once per range).
// Two CTAs share an SM: one's compute-bound search overlaps the other's store-bound interpolation.
constexpr int kFpThreads = 512;

__global__ void __launch_bounds__(kFpThreads, 2) fp_fused_kernel(const float *__restrict__ unknown, const float *__restrict__ known,
                                                                 const float *__restrict__ feats, float *__restrict__ out,
                                                                 int *__restrict__ idx_out, float *__restrict__ weight_out, int c,
                                                                 int n, int m, int PR, int ranges, int CH) {
  extern __shared__ __align__(16) unsigned char s_fp[];
  int *s_idx = reinterpret_cast<int *>(s_fp);                       // [PR * 3]
  float *s_w = reinterpret_cast<float *>(s_idx + (size_t)PR * 3);   // [PR * 3]
  float4 *s_rows = reinterpret_cast<float4 *>(s_w + (size_t)PR * 3);  // phase 1: known tile; phase 2: [CH/4][m]
  const int tid = threadIdx.x;
  const int scene = blockIdx.x / ranges, range = blockIdx.x - scene * ranges;
  const int p0 = range * PR, pc = min(PR, n - p0);
  const float *kn = known + (size_t)scene * m * 3;
  const float *un = unknown + ((size_t)scene * n + p0) * 3;

  // ---- phase 1: three nearest known points + weights of the range's points (up to PPT points per thread) ----
  constexpr int PPT = 4;
  const int tile_cap = (CH / 4) * m;  // float4 slots of the row region, reused for the known coordinates
  for (int pbase = 0; pbase < pc; pbase += kFpThreads * PPT) {
    float ux[PPT], uy[PPT], uz[PPT], b1[PPT], b2[PPT], b3[PPT];
    int i1[PPT], i2[PPT], i3[PPT];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      const int j = pbase + q * kFpThreads + tid;
      const int js = j < pc ? j : 0;
      ux[q] = __ldg(un + (size_t)js * 3), uy[q] = __ldg(un + (size_t)js * 3 + 1), uz[q] = __ldg(un + (size_t)js * 3 + 2);
      b1[q] = b2[q] = b3[q] = __int_as_float(0x7f800000);
      i1[q] = i2[q] = i3[q] = 0;
    }
    for (int base = 0; base < m; base += tile_cap) {
      const int tc = min(tile_cap, m - base);
      __syncthreads();
      for (int e = tid; e < tc; e += kFpThreads) {
        const float *p = kn + (size_t)(base + e) * 3;
        s_rows[e] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
      }
      __syncthreads();
#pragma unroll 2
      for (int k = 0; k < tc; ++k) {
        const float4 p = s_rows[k];
        const int kk = base + k;
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
          const float d = sqdist3(ux[q] - p.x, uy[q] - p.y, uz[q] - p.z);
          if (d < b3[q]) {  // strict `<` cascade in ascending index: the lowest index wins ties (interpolate_gpu.cu:38-56)
            if (d < b1[q]) {
              b3[q] = b2[q], i3[q] = i2[q], b2[q] = b1[q], i2[q] = i1[q], b1[q] = d, i1[q] = kk;
            } else if (d < b2[q]) {
              b3[q] = b2[q], i3[q] = i2[q], b2[q] = d, i2[q] = kk;
            } else {
              b3[q] = d, i3[q] = kk;
            }
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      const int j = pbase + q * kFpThreads + tid;
      if (j < pc) {
        const float d1 = __fsqrt_rn(b1[q]), d2 = __fsqrt_rn(b2[q]), d3 = __fsqrt_rn(b3[q]);
        const float r1 = __frcp_rn(__fadd_rn(d1, 1e-8f)), r2 = __frcp_rn(__fadd_rn(d2, 1e-8f)), r3 = __frcp_rn(__fadd_rn(d3, 1e-8f));
        const float norm = __fadd_rn(__fadd_rn(r1, r3), r2);  // torch.sum over three contiguous elements: (r0 + r2) + r1
        const float w1 = __fdiv_rn(r1, norm), w2 = __fdiv_rn(r2, norm), w3 = __fdiv_rn(r3, norm);
        s_idx[j * 3] = i1[q], s_idx[j * 3 + 1] = i2[q], s_idx[j * 3 + 2] = i3[q];
        s_w[j * 3] = w1, s_w[j * 3 + 1] = w2, s_w[j * 3 + 2] = w3;
        if (idx_out) {
          const size_t o = ((size_t)scene * n + p0 + j) * 3;
          idx_out[o] = i1[q], idx_out[o + 1] = i2[q], idx_out[o + 2] = i3[q];
          weight_out[o] = w1, weight_out[o + 1] = w2, weight_out[o + 2] = w3;
        }
      }
    }
  }
  // points past the end of a ragged last quad read slot 0 with weight 0
  for (int j = pc + tid; j < ((pc + 3) & ~3); j += kFpThreads) {
    s_idx[j * 3] = s_idx[j * 3 + 1] = s_idx[j * 3 + 2] = 0;
    s_w[j * 3] = s_w[j * 3 + 1] = s_w[j * 3 + 2] = 0.f;
  }

  // ---- phase 2: interpolate the range for every channel chunk ----
  const int G = CH / 4;
  const int quads = (pc + 3) / 4;
  const bool vec_ok = (n % 4 == 0) && (p0 % 4 == 0);
  for (int ch_base = 0; ch_base < c; ch_base += CH) {
    const int gcount = min(G, (c - ch_base + 3) / 4);
    __syncthreads();
    for (int g = 0; g < gcount; ++g) {
      const float *src = feats + ((size_t)scene * c + ch_base + g * 4) * m;
      const int nv = min(4, c - (ch_base + g * 4));
      for (int i = tid; i < m; i += kFpThreads) {
        float4 o;
        o.x = __ldg(src + i);
        o.y = nv > 1 ? __ldg(src + (size_t)m + i) : 0.f;
        o.z = nv > 2 ? __ldg(src + 2 * (size_t)m + i) : 0.f;
        o.w = nv > 3 ? __ldg(src + 3 * (size_t)m + i) : 0.f;
        s_rows[(size_t)g * m + i] = o;
      }
    }
    __syncthreads();
    for (int qd = tid; qd < quads; qd += kFpThreads) {
      int id[12];
      float ww[12];
      {
        const int4 a = *reinterpret_cast<const int4 *>(s_idx + qd * 12), b = *reinterpret_cast<const int4 *>(s_idx + qd * 12 + 4),
                   d = *reinterpret_cast<const int4 *>(s_idx + qd * 12 + 8);
        id[0] = a.x, id[1] = a.y, id[2] = a.z, id[3] = a.w, id[4] = b.x, id[5] = b.y, id[6] = b.z, id[7] = b.w;
        id[8] = d.x, id[9] = d.y, id[10] = d.z, id[11] = d.w;
        const float4 u = *reinterpret_cast<const float4 *>(s_w + qd * 12), v = *reinterpret_cast<const float4 *>(s_w + qd * 12 + 4),
                     x = *reinterpret_cast<const float4 *>(s_w + qd * 12 + 8);
        ww[0] = u.x, ww[1] = u.y, ww[2] = u.z, ww[3] = u.w, ww[4] = v.x, ww[5] = v.y, ww[6] = v.z, ww[7] = v.w;
        ww[8] = x.x, ww[9] = x.y, ww[10] = x.z, ww[11] = x.w;
      }
      const int valid = min(4, pc - qd * 4);
      for (int g = 0; g < gcount; ++g) {
        const float4 *row = s_rows + (size_t)g * m;
        float o[4][4];  // [channel][point]
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float4 a = row[id[p * 3]], b = row[id[p * 3 + 1]], d = row[id[p * 3 + 2]];
          const float w1 = ww[p * 3], w2 = ww[p * 3 + 1], w3 = ww[p * 3 + 2];
          o[0][p] = __fmaf_rn(d.x, w3, __fmaf_rn(a.x, w1, __fmul_rn(b.x, w2)));
          o[1][p] = __fmaf_rn(d.y, w3, __fmaf_rn(a.y, w1, __fmul_rn(b.y, w2)));
          o[2][p] = __fmaf_rn(d.z, w3, __fmaf_rn(a.z, w1, __fmul_rn(b.z, w2)));
          o[3][p] = __fmaf_rn(d.w, w3, __fmaf_rn(a.w, w1, __fmul_rn(b.w, w2)));
        }
        const int ch0 = ch_base + g * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (ch0 + e < c) {
            float *dst = out + ((size_t)scene * c + ch0 + e) * n + p0 + (size_t)qd * 4;
            if (vec_ok && valid == 4) {
              st_cs_f4(dst, make_float4(o[e][0], o[e][1], o[e][2], o[e][3]));
            } else {
              for (int p = 0; p < valid; ++p) st_cs_f1(dst + p, o[e][p]);
            }
          }
        }
      }
    }
  }
}

}  // namespace gb

using namespace gb;

extern "C" int gb_three_interp_fwd(const float *points, const int *idx, const float *weight, float *out, int b, int c, int m, int n,
                                   gb_stream_t stream) {
  if (b < 0 || c < 0 || m <= 0 || n < 0) return (int)cudaErrorInvalidValue;
  if (b == 0 || c == 0 || n == 0) return 0;
  if (!points || !idx || !weight || !out) return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  const bool aligned = (n % 4 == 0) && ((((uintptr_t)idx | (uintptr_t)weight | (uintptr_t)out) & 15u) == 0);
  const size_t row_bytes = (size_t)m * sizeof(float);
  if (aligned && 4 * row_bytes <= 200u * 1024u && !(g_tuning.interp_mode & 2)) {
    int CH = (int)((64u * 1024u) / row_bytes);  // ~64 KB of rows per CTA: three CTAs per SM
    CH -= CH % 4;
    if (CH < 4) CH = 4;
    if (CH > ((c + 3) / 4) * 4) CH = ((c + 3) / 4) * 4;
    if (CH > 64) CH = 64;
    const size_t smem = (size_t)CH * row_bytes;
    if (int rc_ = raise_smem_limit(interp_fwd_kernel, smem)) return rc_;
    const int chunks = (c + CH - 1) / CH;
    const int n4 = n / 4;
    const long long total = (long long)b * chunks * n4;
    int ctas_per_sm = (int)((220u * 1024u) / (smem + 1024));
    ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 3 ? 3 : ctas_per_sm);
    long long ctas = (long long)num_sms() * ctas_per_sm;
    const long long min_w = kInterpThreads;
    if (ctas * min_w > total) ctas = (total + min_w - 1) / min_w;
    if (ctas < 1) ctas = 1;
    const long long wpc = (total + ctas - 1) / ctas;
    ctas = (total + wpc - 1) / wpc;
    interp_fwd_kernel<<<(unsigned)ctas, kInterpThreads, smem, s>>>(points, idx, weight, out, c, m, n4, CH, chunks, total, wpc,
                                                                 (g_tuning.interp_mode & 1) ? 0 : 1);
    count_launch();
    return finish_launch();
  }
  const size_t total = (size_t)b * c * n;
  size_t grid = (total + 255) / 256;
  if (grid > (size_t)num_sms() * 32) grid = (size_t)num_sms() * 32;
  interp_fwd_generic_kernel<<<(unsigned)grid, 256, 0, s>>>(points, idx, weight, out, c, m, (size_t)n, total);
  count_launch();
  return finish_launch();
}


/* three_nn + inverse-distance weights + three_interpolate in one launch: what PointnetFPModule.forward
 * (pointnet2_modules.py:413-420), upsampling.three_interpolation (upsampling.py:67-74) and graspbalance.py:37-41 compute.
 * unknown [b,n,3], known [b,m,3] (m >= 1), feats [b,c,m] -> out [b,c,n].  idx_out / weight_out [b,n,3]: both NULL (inference:
 * nothing but `out` is written) or both given (training: three_interpolate's backward reads them).  Indices, weights and
 * values are bit-identical to gb_three_nn_weights followed by gb_three_interp_fwd. */
extern "C" int gb_three_interpolation(const float *unknown, const float *known, const float *feats, float *out, int *idx_out,
                                      float *weight_out, int b, int c, int n, int m, gb_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || m <= 0 || ((idx_out == nullptr) != (weight_out == nullptr))) return (int)cudaErrorInvalidValue;
  if (b == 0 || n == 0) return 0;
  if (!unknown || !known || (c > 0 && (!feats || !out))) return (int)cudaErrorInvalidValue;
  if ((((uintptr_t)out) & 15u) != 0 || (size_t)m * 16 > 64u * 1024u) return (int)cudaErrorNotSupported;  // callers fall back to the two-launch path
  cudaStream_t s = (cudaStream_t)stream;
  // points per CTA: enough CTAs to fill the GPU twice, at most 2048 points (48 KB of parked idx / weight)
  long long PR = ((long long)b * n + 2LL * 2 * num_sms() - 1) / (2LL * 2 * num_sms());
  PR = ((PR + 511) / 512) * 512;
  PR = PR < 512 ? 512 : (PR > 2048 ? 2048 : PR);
  const int ranges = (int)((n + PR - 1) / PR);
  int CH = (int)((64u * 1024u) / ((size_t)m * 4));
  CH -= CH % 4;
  CH = CH < 4 ? 4 : (CH > 64 ? 64 : CH);
  if (CH > ((c + 3) / 4) * 4 && c > 0) CH = ((c + 3) / 4) * 4;
  const size_t smem = (size_t)PR * 3 * 8 + (size_t)CH * m * 4;
  if (int rc_ = raise_smem_limit(fp_fused_kernel, smem)) return rc_;
  fp_fused_kernel<<<(unsigned)((long long)b * ranges), kFpThreads, smem, s>>>(unknown, known, feats, out, idx_out, weight_out, c, n, m, (int)PR,
                                                                             ranges, CH);
  count_launch();
  return finish_launch();
}

static int interp_bwd_impl(const float *grad_out, const int *idx, const float *weight, float *grad_points, int b, int c, int n, int m,
                           int overwrite, gb_stream_t stream) {
  if (b < 0 || c < 0 || m <= 0 || n < 0) return (int)cudaErrorInvalidValue;
  if (b == 0 || c == 0) return 0;
  if (!grad_points || (n > 0 && (!grad_out || !idx || !weight))) return (int)cudaErrorInvalidValue;
  if (n == 0) return overwrite ? (int)cudaMemsetAsync(grad_points, 0, (size_t)b * c * m * sizeof(float), (cudaStream_t)stream) : 0;
  // atomic-free sorted segmented sum (scatter.cu): entries e = 3*j + t, source g[c][e / 3], weight w[e]
  if (!(g_tuning.interp_mode & 4) && seg_scatter_supported(b, c, m, (size_t)n * 3, 3))
    return seg_scatter_add(grad_out, (size_t)c * n, idx, weight, grad_points, b, c, m, (size_t)n * 3, 3, overwrite, (cudaStream_t)stream);
  if (overwrite) {
    cudaError_t e = cudaMemsetAsync(grad_points, 0, (size_t)b * c * m * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
  }
  const size_t total = (size_t)b * c * n;
  size_t grid = (total + 255) / 256;
  if (grid > (size_t)num_sms() * 32) grid = (size_t)num_sms() * 32;
  interp_bwd_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(grad_out, idx, weight, grad_points, c, m, (size_t)n, total);
  count_launch();
  return finish_launch();
}

extern "C" int gb_three_interp_bwd(const float *grad_out, const int *idx, const float *weight, float *grad_points, int b, int c, int n,
                                   int m, gb_stream_t stream) {
  return interp_bwd_impl(grad_out, idx, weight, grad_points, b, c, n, m, 0, stream);
}

extern "C" int gb_three_interp_bwd_set(const float *grad_out, const int *idx, const float *weight, float *grad_points, int b, int c,
                                       int n, int m, gb_stream_t stream) {
  return interp_bwd_impl(grad_out, idx, weight, grad_points, b, c, n, m, 1, stream);
}
