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

constexpr int kInterpThreads = 512;

// points [b,c,m]; idx, weight [b,n,3]; out [b,c,n]; n % 4 == 0.  CH channels per fill (multiple of 4).
__global__ void __launch_bounds__(kInterpThreads) interp_fwd_kernel(const float *__restrict__ points, const int *__restrict__ idx,
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
