// three_nn.cu -- three nearest "known" points for every "unknown" point.
//
// Replaces three_nn_kernel (PointNet/_ext_src/src/interpolate_gpu.cu:14-64, one block per scene) and three_nn_kernel_fast
// (pointnet2_batch/src/interpolate_gpu.cu:16-59).  One thread per unknown; the known set is staged in shared memory as
// float4 so that every candidate costs one broadcast LDS.128 instead of three dependent global loads.
//
// Semantics: candidates in ascending index, strict `<` cascade so the lowest index wins ties; the
// reference keeps its bests as double initialised to 1e40 -- a float compared as double is exact, and (float)1e40 is
// +inf, so float bests initialised to +inf give identical outputs (including +inf distances and index 0 when m < 3).
// Outputs SQUARED distances; the Python wrappers take the sqrt (pointnet2_utils.py:84, upsampling.py:26).
#include "common.cuh"

namespace gb {

constexpr int kNNThreads = 256;
constexpr int kNNTile = 2048;  // known points per shared-memory tile (32 KB as float4)

// WEIGHTS: additionally the inverse-distance interpolation weights every caller derives from the result
// (pointnet2_modules.py:413-416, upsampling.py:69-72, graspbalance.py:37-41):
//   dist = sqrt(dist2); r = 1 / (dist + 1e-8); weight = r / (r0 + r1 + r2)
// with the roundings of the torch ops they replace (IEEE sqrt, add, reciprocal, divide; the sum in the order torch's
// reduce kernel adds three contiguous elements, (r0 + r2) + r1 -- checked bit for bit on a B200), and dist2 then
// holds dist (the square root), which is what the Python-level three_nn returns.
template <bool WEIGHTS>
__global__ void __launch_bounds__(kNNThreads) three_nn_kernel(const float *__restrict__ unknown, const float *__restrict__ known,
                                                              float *__restrict__ dist2, int *__restrict__ idx,
                                                              float *__restrict__ weight, int n, int m) {
  __shared__ float4 tile[kNNTile];
  const int scene = blockIdx.y;
  const int j = blockIdx.x * kNNThreads + threadIdx.x;
  known += (size_t)scene * m * 3;
  const bool ok = j < n;
  const size_t uj = (size_t)scene * n + (ok ? j : 0);
  const float ux = __ldg(unknown + uj * 3), uy = __ldg(unknown + uj * 3 + 1), uz = __ldg(unknown + uj * 3 + 2);

  float b1 = __int_as_float(0x7f800000), b2 = b1, b3 = b1;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int base = 0; base < m; base += kNNTile) {
    const int tc = min(kNNTile, m - base);
    __syncthreads();
    for (int e = threadIdx.x; e < tc; e += kNNThreads) {
      const float *p = known + (size_t)(base + e) * 3;
      tile[e] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < tc; ++k) {
      const float4 p = tile[k];
      const float d = sqdist3(ux - p.x, uy - p.y, uz - p.z);
      if (d < b3) {  // b1 <= b2 <= b3 always, so this guard is equivalent to the reference's three-way cascade
        const int kk = base + k;
        if (d < b1) {
          b3 = b2, i3 = i2, b2 = b1, i2 = i1, b1 = d, i1 = kk;
        } else if (d < b2) {
          b3 = b2, i3 = i2, b2 = d, i2 = kk;
        } else {
          b3 = d, i3 = kk;
        }
      }
    }
  }
  if (ok) {
    if (WEIGHTS) {
      b1 = __fsqrt_rn(b1), b2 = __fsqrt_rn(b2), b3 = __fsqrt_rn(b3);
      const float r1 = __frcp_rn(__fadd_rn(b1, 1e-8f)), r2 = __frcp_rn(__fadd_rn(b2, 1e-8f)), r3 = __frcp_rn(__fadd_rn(b3, 1e-8f));
      const float norm = __fadd_rn(__fadd_rn(r1, r3), r2);  // torch.sum over a 3-element inner dimension: (r0 + r2) + r1
      weight[uj * 3] = __fdiv_rn(r1, norm), weight[uj * 3 + 1] = __fdiv_rn(r2, norm), weight[uj * 3 + 2] = __fdiv_rn(r3, norm);
    }
    dist2[uj * 3] = b1, dist2[uj * 3 + 1] = b2, dist2[uj * 3 + 2] = b3;
    idx[uj * 3] = i1, idx[uj * 3 + 1] = i2, idx[uj * 3 + 2] = i3;
  }
}

}  // namespace gb

extern "C" int gb_three_nn(const float *unknown, const float *known, float *dist2, int *idx, int b, int n, int m,
                           gb_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return (int)cudaErrorInvalidValue;
  if (b == 0 || n == 0) return 0;  // nothing to do (empty tensors have null data pointers)
  if (!unknown || (m > 0 && !known) || !dist2 || !idx) return (int)cudaErrorInvalidValue;
  if (gb::three_nn_grid_worth(b, n, m)) return gb::three_nn_grid(unknown, known, dist2, idx, nullptr, b, n, m, (cudaStream_t)stream);
  for (int b0 = 0; b0 < b; b0 += 65535) {  // the batch rides on gridDim.y: slabs of 65535 scenes
    const int bb = b - b0 < 65535 ? b - b0 : 65535;
    dim3 grid((n + gb::kNNThreads - 1) / gb::kNNThreads, bb);
    gb::three_nn_kernel<false><<<grid, gb::kNNThreads, 0, (cudaStream_t)stream>>>(unknown + (size_t)b0 * n * 3, known + (size_t)b0 * m * 3,
                                                                                 dist2 + (size_t)b0 * n * 3, idx + (size_t)b0 * n * 3, nullptr, n, m);
    gb::count_launch();
    if (int rc = gb::finish_launch()) return rc;
  }
  return 0;
}

/* three_nn followed by the inverse-distance weights of its callers, one launch:
 * dist [b,n,3] = sqrt of the squared distances, idx [b,n,3], weight [b,n,3] = (1/(dist+1e-8)) / sum_k (1/(dist_k+1e-8)). */
extern "C" int gb_three_nn_weights(const float *unknown, const float *known, float *dist, int *idx, float *weight, int b, int n, int m,
                                   gb_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return (int)cudaErrorInvalidValue;
  if (b == 0 || n == 0) return 0;
  if (!unknown || (m > 0 && !known) || !dist || !idx || !weight) return (int)cudaErrorInvalidValue;
  if (gb::three_nn_grid_worth(b, n, m)) return gb::three_nn_grid(unknown, known, dist, idx, weight, b, n, m, (cudaStream_t)stream);
  for (int b0 = 0; b0 < b; b0 += 65535) {
    const int bb = b - b0 < 65535 ? b - b0 : 65535;
    dim3 grid((n + gb::kNNThreads - 1) / gb::kNNThreads, bb);
    gb::three_nn_kernel<true><<<grid, gb::kNNThreads, 0, (cudaStream_t)stream>>>(unknown + (size_t)b0 * n * 3, known + (size_t)b0 * m * 3,
                                                                                dist + (size_t)b0 * n * 3, idx + (size_t)b0 * n * 3,
                                                                                weight + (size_t)b0 * n * 3, n, m);
    gb::count_launch();
    if (int rc = gb::finish_launch()) return rc;
  }
  return 0;
}
