// group_max.cu -- grouping_operation followed by the max over the nsample axis, in one pass.
//
// What PointnetSAModuleVotes_WOMLP.forward (PointNet/pointnet2_modules.py:324-335) and the 'max' pooling of
// PointnetSAModuleVotes (:173-175) compute with group_points_kernel (group_points_gpu.cu:17-36) + F.max_pool2d when no MLP
// sits between the two: out[b,c,j] = max_k f[b,c,idx[b,j,k]].  The [B,C,m,nsample] tensor (nsample times the output) is
// never written.  Rows of a channel chunk are staged in shared memory interleaved four channels per point, as the
// grouping forward does; a thread owns one query: nsample/4 128-bit index loads (shared by all channel groups of the
// chunk), one LDS.128 per index and group, a running maximum with ATen's max-pool rule (`val > max || isnan(val)`, scan in
// k order: the first maximum wins, a NaN sticks), and the winning SOURCE index for the backward.
// Backward: grad[b,c,arg[b,c,j]] += gout[b,c,j] -- m adds per row into an L2-resident [C,N] tensor (red.global.add.f32).
#include "common.cuh"

namespace gb {

constexpr int kGroupMaxThreads = 256;

// points [b,c,n]; idx [b,m,ns]; out [b,c,m]; arg [b,c,m] (source index of the maximum) or nullptr.  CH % 4 == 0.
__global__ void __launch_bounds__(kGroupMaxThreads) group_max_fwd_kernel(const float *__restrict__ points, const int *__restrict__ idx,
                                                                        float *__restrict__ out, int *__restrict__ arg, int c, int n,
                                                                        int m, int ns, int CH, int chunks, int ranges) {
  extern __shared__ __align__(16) float s_gm[];  // [CH/4][n] float4
  float4 *srow = reinterpret_cast<float4 *>(s_gm);
  const int tid = threadIdx.x;
  const int pair = blockIdx.x / ranges, r = blockIdx.x - pair * ranges;
  const int scene = pair / chunks, chunk = pair - scene * chunks;
  const int ch_base = chunk * CH;
  const int gcount = min(CH / 4, (c - ch_base + 3) / 4);
  for (int g = 0; g < gcount; ++g) {
    const float *src = points + ((size_t)scene * c + ch_base + g * 4) * n;
    const int nv = min(4, c - (ch_base + g * 4));
    for (int i = tid; i < n; i += kGroupMaxThreads) {
      float4 o;
      o.x = __ldg(src + i);
      o.y = nv > 1 ? __ldg(src + (size_t)n + i) : 0.f;
      o.z = nv > 2 ? __ldg(src + 2 * (size_t)n + i) : 0.f;
      o.w = nv > 3 ? __ldg(src + 3 * (size_t)n + i) : 0.f;
      srow[(size_t)g * n + i] = o;
    }
  }
  __syncthreads();
  const int j0 = (int)(((long long)m * r) / ranges), j1 = (int)(((long long)m * (r + 1)) / ranges);
  const bool vec = (ns % 4 == 0);
  for (int j = j0 + tid; j < j1; j += kGroupMaxThreads) {
    const int *ip = idx + ((size_t)scene * m + j) * ns;
    for (int g = 0; g < gcount; ++g) {
      const float4 *row = srow + (size_t)g * n;
      float best[4];
      int bi[4];
      auto take = [&](int id, bool first) {
        const float4 v = row[id];
        const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (first || x[e] > best[e] || x[e] != x[e]) best[e] = x[e], bi[e] = id;  // ATen max-pool: val > max || isnan(val)
      };
      if (vec) {
        for (int k = 0; k < ns; k += 4) {
          const int4 id = ld_nc_i4(ip + k);
          take(id.x, k == 0), take(id.y, false), take(id.z, false), take(id.w, false);
        }
      } else {
        for (int k = 0; k < ns; ++k) take(__ldg(ip + k), k == 0);
      }
      const int ch0 = ch_base + g * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (ch0 + e < c) {
          const size_t o = ((size_t)scene * c + ch0 + e) * m + j;
          out[o] = best[e];
          if (arg) arg[o] = bi[e];
        }
      }
    }
  }
}

__global__ void group_max_bwd_kernel(const float *__restrict__ grad_out, const int *__restrict__ arg, float *__restrict__ grad_points, int n,
                                     int m, size_t total) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t row = e / m;
    atomicAdd(grad_points + row * n + __ldg(arg + e), __ldg(grad_out + e));
  }
}

}  // namespace gb

using namespace gb;

/* grouping_operation + max over nsample in one pass (PointNet/pointnet2_modules.py:324-335, :173-175).
 * points [b,c,n], idx [b,npoints,nsample] (nsample >= 1) -> out [b,c,npoints] = max_k points[b,c,idx[b,j,k]] with ATen's
 * max-pool rule (first maximum in k order, NaN propagates); arg [b,c,npoints] i32 = the source index of the maximum (NULL
 * when no backward follows).  Equals gb_group_fwd followed by F.max_pool2d(kernel_size=[1, nsample]). */
extern "C" int gb_group_max_fwd(const float *points, const int *idx, float *out, int *arg, int b, int c, int n, int npoints, int nsample,
                                gb_stream_t stream) {
  if (b < 0 || c < 0 || n <= 0 || npoints < 0 || nsample <= 0) return (int)cudaErrorInvalidValue;
  if (b == 0 || c == 0 || npoints == 0) return 0;
  if (!points || !idx || !out) return (int)cudaErrorInvalidValue;
  const size_t row_bytes = (size_t)n * sizeof(float);
  if (4 * row_bytes > 200u * 1024u || (nsample % 4 == 0 && ((uintptr_t)idx & 15u) != 0)) return (int)cudaErrorNotSupported;
  int CH = (int)((64u * 1024u) / row_bytes);  // ~64 KB of rows per CTA
  CH -= CH % 4;
  CH = CH < 4 ? 4 : (CH > 64 ? 64 : CH);
  if (CH > ((c + 3) / 4) * 4) CH = ((c + 3) / 4) * 4;
  const int chunks = (c + CH - 1) / CH;
  const size_t smem = (size_t)CH * row_bytes;
  if (int rc_ = raise_smem_limit(group_max_fwd_kernel, smem)) return rc_;
  // position ranges per (scene, chunk): enough CTAs for ~3 per SM, at least one sweep of the block each
  long long ranges = (3LL * num_sms() + (long long)b * chunks - 1) / ((long long)b * chunks);
  const long long max_r = (npoints + kGroupMaxThreads - 1) / kGroupMaxThreads;
  ranges = ranges < 1 ? 1 : (ranges > max_r ? max_r : ranges);
  group_max_fwd_kernel<<<(unsigned)((long long)b * chunks * ranges), kGroupMaxThreads, smem, (cudaStream_t)stream>>>(
      points, idx, out, arg, c, n, npoints, nsample, CH, chunks, (int)ranges);
  count_launch();
  return finish_launch();
}

/* Backward of gb_group_max_fwd: grad_out [b,c,npoints], arg [b,c,npoints] -> ACCUMULATES into grad_points [b,c,n] (what
 * max_pool2d's backward followed by group_points_grad_kernel, group_points_gpu.cu:69-90, adds up). */
extern "C" int gb_group_max_bwd(const float *grad_out, const int *arg, float *grad_points, int b, int c, int n, int npoints,
                                gb_stream_t stream) {
  if (b < 0 || c < 0 || n <= 0 || npoints < 0) return (int)cudaErrorInvalidValue;
  const size_t total = (size_t)b * c * npoints;
  if (total == 0) return 0;
  if (!grad_out || !arg || !grad_points) return (int)cudaErrorInvalidValue;
  size_t grid = (total + 255) / 256;
  if (grid > (size_t)num_sms() * 32) grid = (size_t)num_sms() * 32;
  group_max_bwd_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(grad_out, arg, grad_points, n, npoints, total);
  count_launch();
  return finish_launch();
}
