// knn.cu -- brute-force k nearest neighbours without the distance matrix.
//
// Replaces cuComputeDistanceGlobal + cuInsertionSort (KNN/Pytorch_CUDA_KNN/cuda/knn.cu:36-176) and the host loop over the
// batch in knn.h:31-38.  The reference writes the full nref x nquery distance matrix to a scratch buffer (82 MB at
// 20000 x 1024) and then walks each column with one thread.  Here one WARP owns a query: reference points are staged
// tile by tile in shared memory, every lane evaluates one candidate per step, and the k best are kept in a sorted
// per-warp list in shared memory (registers for k = 1).  No scratch, one launch for the whole batch.
//
// Semantics: distance ssd = fmaf(t,t,ssd) over the dim rows in order with t = ref - query (the 16-wide
// zero padding of the reference's tiles adds exact zeros); result = the k smallest by (distance, index) ascending --
// a candidate equal to the current k-th distance is NOT inserted, equal distances keep index order; indices are
// 1-BASED int64 at idx[b, rank, q].
#include "common.cuh"

namespace gb {

constexpr int kKnnWarps = 8;
constexpr int kKnnTileFloats = 6144;  // 24 KB of reference coordinates per tile

// DT = compile-time dim (3) or 0 for a runtime dim.  K1 = true: k == 1 fast path (registers only).
template <int DT, bool K1>
__global__ void __launch_bounds__(kKnnWarps * 32) knn_kernel(const float *__restrict__ ref, const float *__restrict__ query,
                                                             long long *__restrict__ idx, int dim_rt, int R, int Q, int k, int ts) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int dim = DT ? DT : dim_rt;
  float *tile = reinterpret_cast<float *>(s_raw);                       // [dim][ts]
  float *qs = tile + kKnnTileFloats;                                     // [warps][dim]
  float *ldist = qs + kKnnWarps * dim;                                   // [warps][k]   (unused for K1)
  int *lidx = reinterpret_cast<int *>(ldist + (K1 ? 0 : kKnnWarps * k));  // [warps][k]

  const int scene = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = blockIdx.x * kKnnWarps + warp;
  const bool qok = q < Q;
  ref += (size_t)scene * dim * R;
  query += (size_t)scene * dim * Q;

  float qreg[DT ? DT : 1];
  if (DT) {
#pragma unroll
    for (int d = 0; d < DT; ++d) qreg[d] = qok ? __ldg(query + (size_t)d * Q + q) : 0.f;
  } else {
    for (int d = lane; d < dim; d += 32) qs[warp * dim + d] = qok ? __ldg(query + (size_t)d * Q + q) : 0.f;
  }
  float *md = ldist + warp * k;
  int *mi = lidx + warp * k;

  // k == 1 state: per-lane best;  general: warp-uniform count and threshold
  float best = __int_as_float(0x7f800000);
  int besti = -1;
  int count = 0;
  float tau = __int_as_float(0x7f800000);

  for (int base = 0; base < R; base += ts) {
    const int tc = min(ts, R - base);
    __syncthreads();
    for (int d = 0; d < dim; ++d)
      for (int e = tid; e < tc; e += kKnnWarps * 32) tile[d * ts + e] = __ldg(ref + (size_t)d * R + base + e);
    __syncthreads();
    if (!qok) continue;
    for (int off = 0; off < tc; off += 32) {
      const int e = off + lane;
      const bool valid = e < tc;
      float ssd = 0.f;
      if (valid) {
        if (DT) {
#pragma unroll
          for (int d = 0; d < DT; ++d) {
            const float t = tile[d * ts + e] - qreg[d];
            ssd = __fmaf_rn(t, t, ssd);
          }
        } else {
          for (int d = 0; d < dim; ++d) {
            const float t = tile[d * ts + e] - qs[warp * dim + d];
            ssd = __fmaf_rn(t, t, ssd);
          }
        }
      }
      const int r = base + e;
      if (K1) {
        if (valid && ssd < best) best = ssd, besti = r;  // ascending r per lane: first minimum kept
        // rows 0..k-1 are always taken by the reference (part 1 of cuInsertionSort): with k == 1 that is row 0
        if (valid && r == 0 && besti < 0) best = ssd, besti = 0;
      } else {
        unsigned mask = __ballot_sync(0xffffffffu, valid && (count < k || ssd < tau));
        while (mask) {
          const int l = __ffs(mask) - 1;
          mask &= mask - 1;
          const float d = __shfl_sync(0xffffffffu, ssd, l);
          const int rr = base + off + l;
          if (count < k || d < tau) {
            // position = number of list entries that are not greater than d (stable: after equal distances)
            int pos = 0;
            for (int j0 = 0; j0 < count; j0 += 32) {
              const int j = j0 + lane;
              pos += __popc(__ballot_sync(0xffffffffu, j < count && !(md[j] > d)));
            }
            const int newcount = min(count + 1, k);
            // shift [pos, newcount-2] up by one, highest chunk first
            for (int j0 = ((newcount - 1) / 32) * 32; j0 >= 0; j0 -= 32) {
              const int j = j0 + lane;
              const bool mv = j > pos && j < newcount;
              float vd = 0.f;
              int vi = 0;
              if (mv) vd = md[j - 1], vi = mi[j - 1];
              __syncwarp();
              if (mv) md[j] = vd, mi[j] = vi;
              __syncwarp();
            }
            if (lane == 0 && pos < newcount) md[pos] = d, mi[pos] = rr;
            __syncwarp();
            count = newcount;
            if (count == k) tau = md[k - 1];
          }
        }
      }
    }
  }

  if (!qok) return;
  long long *o = idx + (size_t)scene * k * Q + q;
  if (K1) {
    // warp argmin on (distance, index)
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, d);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, d);
      const bool take = (oi >= 0) && (besti < 0 || ob < best || (ob == best && oi < besti));
      if (take) best = ob, besti = oi;
    }
    if (lane == 0) o[0] = (long long)besti + 1;
  } else {
    __syncwarp();
    for (int j = lane; j < count; j += 32) o[(size_t)j * Q] = (long long)mi[j] + 1;
  }
}

}  // namespace gb

using namespace gb;

extern "C" int gb_knn(const float *ref, const float *query, int64_t *idx, int b, int dim, int nref, int nquery, int k,
                      gb_stream_t stream) {
  if (b < 0 || nquery < 0) return (int)cudaErrorInvalidValue;
  if (b == 0 || nquery == 0) return 0;  // nothing to do (empty tensors have null data pointers)
  if (dim <= 0 || nref <= 0 || k <= 0 || k > nref || k > 1024 || dim > 256 || !ref || !query || !idx) return (int)cudaErrorInvalidValue;
  if (b > 65535) {  // the batch rides on gridDim.y: slabs of 65535 scenes
    for (int b0 = 0; b0 < b; b0 += 65535) {
      const int bb = b - b0 < 65535 ? b - b0 : 65535;
      const int rc = gb_knn(ref + (size_t)b0 * dim * nref, query + (size_t)b0 * dim * nquery, idx + (size_t)b0 * k * nquery, bb, dim, nref,
                            nquery, k, stream);
      if (rc) return rc;
    }
    return 0;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int ts = (kKnnTileFloats / dim) & ~31;
  if (ts > 2048) ts = 2048;
  if (ts < 32) return (int)cudaErrorInvalidValue;
  const bool k1 = (k == 1);
  const size_t smem = sizeof(float) * ((size_t)kKnnTileFloats + (size_t)kKnnWarps * dim) + (k1 ? 0 : (size_t)kKnnWarps * k * 8);
  dim3 grid((nquery + kKnnWarps - 1) / kKnnWarps, b);
  long long *o = reinterpret_cast<long long *>(idx);
#define GB_KNN_LAUNCH(DT, K1)                                                                                   \
  do {                                                                                                          \
    if (int rc_ = raise_smem_limit(knn_kernel<DT, K1>, smem)) return rc_;                                                                        \
    knn_kernel<DT, K1><<<grid, kKnnWarps * 32, smem, s>>>(ref, query, o, dim, nref, nquery, k, ts);             \
  } while (0)
  if (dim == 3) {
    if (k1) GB_KNN_LAUNCH(3, true); else GB_KNN_LAUNCH(3, false);
  } else {
    if (k1) GB_KNN_LAUNCH(0, true); else GB_KNN_LAUNCH(0, false);
  }
#undef GB_KNN_LAUNCH
  count_launch();
  return finish_launch();
}
