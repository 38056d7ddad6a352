// InvertibleBatchNorm, TRAINING mode (src/bijectors/normalise.jl:51-60): batch statistics over the N columns
// (over all ranks when a communicator is given: the path's second, 2·D+1-double all-reduce), moving-average
// update of the layer's m / v in place (with the n/(n-1) correction of :60), then the forward map and the
// log-Jacobian with the BATCH statistics (:66-67).
//
//   pass 1  bn_stats_kernel     reads x once: per-row Σx and Σx² in fp64, per-CTA partials (deterministic)
//           bn_reduce_kernel    fixed-order sum of the partials -> acc[2D+1] = {Σx, Σx², n}
//           (ncclAllReduce of acc when sharded)
//           bn_finalize_kernel  m = Σx/n, v = Σx²/n − m² (= sum((x−m)²)/n, :55), moving update, batch m/v as float
//   pass 2  the eval-mode BatchNorm op of the chain kernels with the batch statistics
// Algorithmic traffic: 3 passes over D x N floats (read, read, write).
#include <cuda_runtime.h>

#include <cstring>

#include "b2b_internal.h"

namespace b2b {

constexpr int BNT_THREADS = 256;

// Each warp walks columns; lane l owns the float4 chunks {l + 32 v}.  Rows are accumulated in fp64 registers.
template <int V, bool VEC>
__global__ void __launch_bounds__(BNT_THREADS) bn_stats_kernel(const float* __restrict__ x, int D, long long N,
                                                               long long ldx, double* __restrict__ partials) {
  extern __shared__ double sred[];  // [2][Dp] per CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = BNT_THREADS / 32;
  const int Dp = 128 * V;
  for (int i = threadIdx.x; i < 2 * Dp; i += BNT_THREADS) sred[i] = 0.0;
  __syncthreads();
  double s1[V][4], s2[V][4];
#pragma unroll
  for (int v = 0; v < V; ++v)
#pragma unroll
    for (int e = 0; e < 4; ++e) s1[v][e] = s2[v][e] = 0.0;
  const long long gw = (long long)blockIdx.x * nwarps + warp, stride = (long long)gridDim.x * nwarps;
  for (long long col = gw; col < N; col += stride) {
    const float* xc = x + col * ldx;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int r0 = 4 * (lane + 32 * v);
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (VEC) {
        if (r0 < D) q = __ldcs(reinterpret_cast<const float4*>(xc) + lane + 32 * v);
      } else {
        if (r0 + 0 < D) q.x = xc[r0 + 0];
        if (r0 + 1 < D) q.y = xc[r0 + 1];
        if (r0 + 2 < D) q.z = xc[r0 + 2];
        if (r0 + 3 < D) q.w = xc[r0 + 3];
      }
      const double a = q.x, b = q.y, c = q.z, d = q.w;
      s1[v][0] += a; s2[v][0] = fma(a, a, s2[v][0]);
      s1[v][1] += b; s2[v][1] = fma(b, b, s2[v][1]);
      s1[v][2] += c; s2[v][2] = fma(c, c, s2[v][2]);
      s1[v][3] += d; s2[v][3] = fma(d, d, s2[v][3]);
    }
  }
  // combine the warps of the CTA in a fixed order (warp 0 first, ...) for determinism
  for (int w = 0; w < nwarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int v = 0; v < V; ++v)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = 4 * (lane + 32 * v) + e;
          sred[r] += s1[v][e];
          sred[Dp + r] += s2[v][e];
        }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < D; i += BNT_THREADS) {
    partials[(size_t)blockIdx.x * 2 * D + i] = sred[i];
    partials[(size_t)blockIdx.x * 2 * D + D + i] = sred[Dp + i];
  }
}

__global__ void __launch_bounds__(256) bn_reduce_kernel(const double* __restrict__ partials, int nblk, int D,
                                                        long long N, double* __restrict__ acc) {
  for (int i = blockIdx.x * 256 + threadIdx.x; i < 2 * D; i += gridDim.x * 256) {
    double t = 0.0;
    for (int b = 0; b < nblk; ++b) t += partials[(size_t)b * 2 * D + i];
    acc[i] = t;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) acc[2 * D] = (double)N;
}

__global__ void __launch_bounds__(256) bn_finalize_kernel(const double* __restrict__ acc, int D, float mtm,
                                                          float* __restrict__ mov_m, float* __restrict__ mov_v,
                                                          float* __restrict__ batch_m, float* __restrict__ batch_v) {
  const double n = acc[2 * D];
  for (int i = blockIdx.x * 256 + threadIdx.x; i < D; i += gridDim.x * 256) {
    const double mean = acc[i] / n;                       // mean(x; dims), normalise.jl:54
    double var = acc[D + i] / n - mean * mean;            // sum((x .- m).^2) ./ n, :55
    var = var > 0.0 ? var : 0.0;
    batch_m[i] = (float)mean;
    batch_v[i] = (float)var;
    // moving statistics, :59-60  (T.(…) rounds the batch statistic to the parameter eltype first)
    const float mf = (float)mean, vf = (float)var;
    mov_m[i] = (1.0f - mtm) * mov_m[i] + mtm * mf;
    mov_v[i] = (1.0f - mtm) * mov_v[i] + (float)((double)mtm * n / (n - 1.0)) * vf;
  }
}

}  // namespace b2b

extern "C" size_t b2b_batchnorm_train_workspace_bytes(int32_t D) {
  // per-CTA partials (<= 1184 CTAs) + acc[2D+1] + batch m / v
  return (size_t)1184 * 2 * D * sizeof(double) + (size_t)(2 * D + 2) * sizeof(double) + (size_t)2 * D * sizeof(float) + 256;
}

extern "C" int b2b_batchnorm_train_fwd_f32(const float* x, float* y, float* logjac, const float* b, const float* logs,
                                           float* m, float* v, float eps, float mtm, int32_t D, int64_t N,
                                           int64_t ldx, int64_t ldy, int accumulate_logjac, b2b_comm* comm,
                                           void* workspace, size_t workspace_bytes, void* stream_) {
  using namespace b2b;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !b || !logs || !m || !v || D < 1 || N < 2 || ldx < D || (y && ldy < D)) return B2B_EINVAL;
  if (D > 1024) return B2B_EUNSUPPORTED;
  if (!workspace || workspace_bytes < b2b_batchnorm_train_workspace_bytes(D)) return B2B_EWORKSPACE;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = sms * 4;
  if (grid > 1184) grid = 1184;
  const long long want = (N + 7) / 8;
  if (grid > want) grid = (int)want;
  if (grid < 1) grid = 1;
  char* ws = static_cast<char*>(workspace);
  ws += (256 - (reinterpret_cast<uintptr_t>(ws) & 255)) & 255;
  double* partials = reinterpret_cast<double*>(ws);
  double* acc = partials + (size_t)1184 * 2 * D;
  float* batch_m = reinterpret_cast<float*>(acc + 2 * D + 2);
  float* batch_v = batch_m + D;
  const bool vec = (D % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const int V = (D + 127) / 128 <= 1 ? 1 : ((D + 127) / 128 <= 2 ? 2 : ((D + 127) / 128 <= 4 ? 4 : 8));
  const size_t smem = (size_t)2 * 128 * V * sizeof(double);
#define B2B_BNT_LAUNCH(VV)                                                                      \
  if (vec) bn_stats_kernel<VV, true><<<grid, BNT_THREADS, smem, stream>>>(x, D, N, ldx, partials); \
  else bn_stats_kernel<VV, false><<<grid, BNT_THREADS, smem, stream>>>(x, D, N, ldx, partials);
  if (V == 1) { B2B_BNT_LAUNCH(1) } else if (V == 2) { B2B_BNT_LAUNCH(2) } else if (V == 4) { B2B_BNT_LAUNCH(4) } else { B2B_BNT_LAUNCH(8) }
#undef B2B_BNT_LAUNCH
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  bn_reduce_kernel<<<(2 * D + 255) / 256, 256, 0, stream>>>(partials, grid, D, N, acc);
  if (comm) {  // sharded batch: one all-reduce of {Σx, Σx², n}
    const int rc = b2b_allreduce_sum_f64(comm, acc, 2 * D + 1, stream);
    if (rc != B2B_OK) return rc;
  }
  bn_finalize_kernel<<<(D + 255) / 256, 256, 0, stream>>>(acc, D, mtm, m, v, batch_m, batch_v);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (!y && !logjac) return B2B_OK;
  b2b_layer_desc d;
  memset(&d, 0, sizeof(d));
  d.kind = B2B_BATCHNORM;
  d.p0 = b;
  d.p1 = logs;
  d.p2 = batch_m;
  d.p3 = batch_v;
  d.f0 = eps;
  return b2b_chain_run_f32(&d, 1, x, y, logjac, nullptr, D, N, ldx, ldy, accumulate_logjac, nullptr, 0, stream_);
}
