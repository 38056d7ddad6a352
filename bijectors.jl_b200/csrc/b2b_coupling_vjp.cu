// Reverse mode (vector-Jacobian product) of the affine coupling layer and of the eval-mode InvertibleBatchNorm -- the two
// layer kinds of a RealNVP flow (BASELINE config 5) -- so that such a flow can be TRAINED on the device path.
//
// Reference: what the reference's reverse-mode AD computes for Coupling (src/bijectors/coupling.jl:206-228) with the law
// Shift(t) ∘ Scale(exp.(s)), [s; t] = W·x₂ + c; the pullback of `combine` (ext/BijectorsChainRulesCoreExt.jl:48-62) is
// the row scatter of the three cotangent blocks.  Restated and finite-difference-checked in oracle/oracle_np.py
// (coupling_affine_vjp, batchnorm_eval_vjp).
//   forward : y₁ = e^s x₁ + t,  lj = Σ s        x̄₁ = e^s ȳ₁,   s̄ = ȳ₁ e^s x₁ + l̄,   t̄ = ȳ₁
//   inverse : x₁ = (y₁ − t) e^−s, lj = −Σ s     ȳ₁ = e^−s x̄₁,  s̄ = −x₁ x̄₁ − l̄,    t̄ = −e^−s x̄₁
//   both    : x̄₂ = ȳ₂ + Wᵀ[s̄; t̄],  W̄ = Σ_n [s̄; t̄]_n x₂ₙᵀ,  c̄ = Σ_n [s̄; t̄]_n,  pass-through rows x̄₃ = ȳ₃
//
// Three GEMMs of the forward's size per tile (recompute [s; t], the x̄₂ product, the W̄ outer-product accumulation) in
// exact fp32 on the CUDA cores: a persistent CTA per SM, tiles of 32 columns transposed into shared memory, the CTA's
// partial W̄ (2·n1 x n2 <= 256 x 128 floats) lives in REGISTERS for the whole launch (128 accumulators per thread) and is
// written once; a second kernel sums the per-CTA partials in a fixed order (deterministic).  First version: the tensor-core
// (fp16-split tcgen05) form of the forward kernel has not been carried over to the reverse mode.
#include <cuda_runtime.h>

#include <cstring>

#include "b2b_internal.h"

namespace b2b {

constexpr int CV_TC = 32;            // columns per tile
constexpr int CV_LD = CV_TC + 1;     // padded row stride of the shared-memory tiles
constexpr int CV_THREADS = 256;

struct CvParams {
  const float* x;
  const float* ybar;
  const float* ljbar;
  float* xbar;
  const float* W;
  const float* c;
  const int32_t* idx1;
  const int32_t* idx2;
  float* part;  // per-CTA partials: [grid][2n1*n2 + 2n1]
  long long N, ldx, ldyb, ldxb;
  int D, n1, n2, row1, row2;
};

template <bool INV>
__global__ void __launch_bounds__(CV_THREADS, 1) coupling_vjp_kernel(const __grid_constant__ CvParams P) {
  extern __shared__ float smem[];
  const int D = P.D, n1 = P.n1, n2 = P.n2, m2 = 2 * n1;
  float* X = smem;                         // [D][CV_LD]   the layer's input tile
  float* YB = X + (size_t)D * CV_LD;       // [D][CV_LD]   cotangent tile, turned into the input cotangent in place
  float* SB = YB + (size_t)D * CV_LD;      // [2n1][CV_LD] s̄ | t̄
  float* LB = SB + (size_t)m2 * CV_LD;     // [CV_TC]      l̄ of the tile's columns
  int* s1 = reinterpret_cast<int*>(LB + CV_TC);  // [n1] rows of x₁
  int* s2 = s1 + n1;                              // [n2] rows of x₂
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < n1; k += CV_THREADS) s1[k] = P.idx1 ? P.idx1[k] : P.row1 + k;
  for (int k = threadIdx.x; k < n2; k += CV_THREADS) s2[k] = P.idx2 ? P.idx2[k] : P.row2 + k;
  const long long tiles = (P.N + CV_TC - 1) / CV_TC;
  const int ldw = m2;
  // this thread's block of the W̄ partial: rows 32·warp .. +31 (of [s̄; t̄]), columns lane + 32·q (of x₂)
  float acc[32][4];
#pragma unroll
  for (int i = 0; i < 32; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
  float cacc = 0.f;  // c̄ partial of row threadIdx.x

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long col0 = tile * CV_TC;
    __syncthreads();  // previous tile fully consumed / index lists visible
    // ---- load + transpose x and ȳ (zero-filled beyond N: such columns contribute exactly 0 to every cotangent) ----
    for (int cidx = warp; cidx < CV_TC; cidx += CV_THREADS / 32) {
      const long long col = col0 + cidx;
      const bool ok = col < P.N;
      for (int r = lane; r < D; r += 32) {
        X[r * CV_LD + cidx] = ok ? __ldcs(P.x + col * P.ldx + r) : 0.f;
        YB[r * CV_LD + cidx] = ok ? __ldcs(P.ybar + col * P.ldyb + r) : 0.f;
      }
      if (lane == 0) LB[cidx] = (ok && P.ljbar) ? P.ljbar[col] : 0.f;
    }
    __syncthreads();
    // ---- [s; t] = W·x₂ + c, then the elementwise cotangents (thread: 4 rows j, column `lane`) -------------------
    for (int jb = 4 * warp; jb < n1; jb += 4 * (CV_THREADS / 32)) {
      float sv[4] = {0.f, 0.f, 0.f, 0.f}, tv[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < n2; ++k) {
        const float xk = X[s2[k] * CV_LD + lane];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (jb + i < n1) {
            sv[i] = fmaf(__ldg(P.W + (size_t)k * ldw + jb + i), xk, sv[i]);
            tv[i] = fmaf(__ldg(P.W + (size_t)k * ldw + n1 + jb + i), xk, tv[i]);
          }
        }
      }
      const float lb = LB[lane];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = jb + i;
        if (j >= n1) break;
        const float s_ = sv[i] + (P.c ? P.c[j] : 0.f), t_ = tv[i] + (P.c ? P.c[n1 + j] : 0.f);
        const int r1 = s1[j];
        const float in1 = X[r1 * CV_LD + lane], cb1 = YB[r1 * CV_LD + lane];
        float sbar, tbar, out1;
        if (!INV) {
          const float e = expf(s_);
          out1 = e * cb1;                 // x̄₁ = e^s ȳ₁
          sbar = fmaf(cb1 * e, in1, lb);  // ȳ₁ e^s x₁ + l̄
          tbar = cb1;
        } else {
          const float em = expf(-s_);
          const float x1 = (in1 - t_) * em;  // the recovered x₁
          out1 = em * cb1;                   // ȳ₁ = e^−s x̄₁
          sbar = -fmaf(x1, cb1, lb);         // −x₁ x̄₁ − l̄
          tbar = -out1;
        }
        YB[r1 * CV_LD + lane] = out1;
        SB[j * CV_LD + lane] = sbar;
        SB[(n1 + j) * CV_LD + lane] = tbar;
      }
    }
    __syncthreads();
    // ---- x̄₂ = ȳ₂ + Wᵀ[s̄; t̄]  (thread: 4 rows k of x₂, column `lane`) ----------------------------------------------
    for (int kb = 4 * warp; kb < n2; kb += 4 * (CV_THREADS / 32)) {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < m2; ++j) {
        const float sb = SB[j * CV_LD + lane];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (kb + i < n2) a[i] = fmaf(__ldg(P.W + (size_t)(kb + i) * ldw + j), sb, a[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (kb + i < n2) YB[s2[kb + i] * CV_LD + lane] += a[i];
    }
    // ---- W̄ += [s̄; t̄]·x₂ᵀ over the tile's columns; c̄ += Σ columns ---------------------------------------------------
    {
      const int rbase = 32 * warp;
      if (rbase < m2) {
#pragma unroll 4
        for (int cidx = 0; cidx < CV_TC; ++cidx) {
          float b[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) b[q] = (lane + 32 * q < n2) ? X[s2[lane + 32 * q] * CV_LD + cidx] : 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float av = (rbase + i < m2) ? SB[(rbase + i) * CV_LD + cidx] : 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][q] = fmaf(av, b[q], acc[i][q]);
          }
        }
      }
      if ((int)threadIdx.x < m2) {
        float t = 0.f;
        for (int cidx = 0; cidx < CV_TC; ++cidx) t += SB[threadIdx.x * CV_LD + cidx];
        cacc += t;
      }
    }
    __syncthreads();  // x̄₂ complete
    // ---- write the input cotangent tile back (coalesced) ---------------------------------------------------------
    for (int cidx = warp; cidx < CV_TC; cidx += CV_THREADS / 32) {
      const long long col = col0 + cidx;
      if (col < P.N)
        for (int r = lane; r < D; r += 32) __stcs(P.xbar + col * P.ldxb + r, YB[r * CV_LD + cidx]);
    }
  }
  // ---- this CTA's partial of W̄ (column-major like W) and c̄ -----------------------------------------------------------
  float* part = P.part + (size_t)blockIdx.x * ((size_t)m2 * n2 + m2);
  const int rbase = 32 * warp;
#pragma unroll
  for (int i = 0; i < 32; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (rbase + i < m2 && lane + 32 * q < n2) part[(size_t)(lane + 32 * q) * m2 + rbase + i] = acc[i][q];
  if ((int)threadIdx.x < m2) part[(size_t)m2 * n2 + threadIdx.x] = cacc;
}

// out[e] = Σ_cta part[cta][e], fixed order
__global__ void __launch_bounds__(256) partial_sum_kernel(const float* __restrict__ part, int nparts, int len, float* __restrict__ out0,
                                                          int len0, float* __restrict__ out1) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < len; e += gridDim.x * blockDim.x) {
    float t = 0.f;
    for (int p = 0; p < nparts; ++p) t += part[(size_t)p * len + e];
    if (e < len0) out0[e] = t;
    else out1[e - len0] = t;
  }
}

// ---- eval-mode InvertibleBatchNorm -------------------------------------------------------------------------------------
// y = A(x − m) + b, A = e^logs / sqrt(v + eps) (normalise.jl:61-67).  forward: x̄ = A ȳ, b̄ = Σ ȳ, l̄ogs = Σ ȳ⊙(y − b) + Σ l̄;
// inverse: ȳ = x̄ / A, b̄ = −Σ x̄/A, l̄ogs = −Σ x̄⊙(x − m) − Σ l̄.  Warp per column, lane r owns rows r, r+32, ...
struct BvParams {
  const float* x;
  const float* ybar;
  const float* ljbar;
  float* xbar;
  const float *b, *logs, *m, *v;
  float eps;
  float* part;  // [grid][2D + 1]
  long long N, ldx, ldyb, ldxb;
  int D, inverse;
};

constexpr int BV_MAXR = 32;  // rows per lane: D <= 1024

__global__ void __launch_bounds__(256) bn_eval_vjp_kernel(const __grid_constant__ BvParams P) {
  extern __shared__ float sm[];  // [8 warps][2D + 1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, D = P.D;
  const int nr = (D + 31) / 32;
  float A[BV_MAXR], sh[BV_MAXR], gb[BV_MAXR], gl[BV_MAXR];
#pragma unroll
  for (int i = 0; i < BV_MAXR; ++i) {
    const int r = lane + 32 * i;
    gb[i] = gl[i] = 0.f;
    if (i < nr && r < D) {
      A[i] = expf(P.logs[r]) / sqrtf(P.v[r] + P.eps);
      sh[i] = P.inverse ? P.b[r] : P.m[r];  // the shift removed before scaling: y − b (inverse) / x − m (forward)
    } else {
      A[i] = 1.f;
      sh[i] = 0.f;
    }
  }
  float lsum = 0.f;
  for (long long n = (long long)blockIdx.x * 8 + warp; n < P.N; n += (long long)gridDim.x * 8) {
#pragma unroll
    for (int i = 0; i < BV_MAXR; ++i) {
      const int r = lane + 32 * i;
      if (i < nr && r < D) {
        const float xv = P.x[n * P.ldx + r], cb = P.ybar[n * P.ldyb + r];
        if (!P.inverse) {
          P.xbar[n * P.ldxb + r] = A[i] * cb;
          gb[i] += cb;
          gl[i] = fmaf(cb, A[i] * (xv - sh[i]), gl[i]);  // ȳ ⊙ (y − b)
        } else {
          const float o = cb / A[i];
          P.xbar[n * P.ldxb + r] = o;
          gb[i] -= o;
          gl[i] = fmaf(-cb, (xv - sh[i]) / A[i], gl[i]);  // −x̄ ⊙ (x − m)
        }
      }
    }
    if (lane == 0 && P.ljbar) lsum += P.ljbar[n];
  }
  float* mine = sm + (size_t)warp * (2 * D + 1);
#pragma unroll
  for (int i = 0; i < BV_MAXR; ++i) {
    const int r = lane + 32 * i;
    if (i < nr && r < D) {
      mine[r] = gb[i];
      mine[D + r] = gl[i];
    }
  }
  if (lane == 0) mine[2 * D] = lsum;
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * D + 1; e += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sm[(size_t)w * (2 * D + 1) + e];
    P.part[(size_t)blockIdx.x * (2 * D + 1) + e] = t;
  }
}

// bbar[r] = Σ parts, logsbar[r] = Σ parts ± Σ l̄
__global__ void bn_vjp_finalize_kernel(const float* __restrict__ part, int nparts, int D, int inverse, float* __restrict__ bbar,
                                       float* __restrict__ logsbar) {
  __shared__ float ls;
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int p = 0; p < nparts; ++p) t += part[(size_t)p * (2 * D + 1) + 2 * D];
    ls = t;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < D; r += blockDim.x) {
    float tb = 0.f, tl = 0.f;
    for (int p = 0; p < nparts; ++p) {
      tb += part[(size_t)p * (2 * D + 1) + r];
      tl += part[(size_t)p * (2 * D + 1) + D + r];
    }
    bbar[r] = tb;
    logsbar[r] = tl + (inverse ? -ls : ls);
  }
}

static int sm_count() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}

}  // namespace b2b

extern "C" size_t b2b_coupling_affine_vjp_workspace_bytes(int32_t n1, int32_t n2) {
  if (n1 < 1 || n1 > 128 || n2 < 1 || n2 > 128) return 0;
  return (size_t)b2b::sm_count() * ((size_t)2 * n1 * n2 + 2 * n1) * sizeof(float) + 256;
}

extern "C" int b2b_coupling_affine_vjp_f32(const b2b_layer_desc* layer, const float* x, const float* ybar, const float* ljbar,
                                           float* xbar, float* Wbar, float* cbar, int32_t D, int64_t N, int64_t ldx,
                                           int64_t ldybar, int64_t ldxbar, void* workspace, size_t workspace_bytes,
                                           void* stream_) {
  using namespace b2b;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!layer || layer->kind != B2B_COUPLING_AFFINE || D < 1 || N < 0 || !Wbar || !cbar) return B2B_EINVAL;
  const b2b_layer_desc& d = *layer;
  const int n1 = d.n0, n2 = d.n1;
  if (!d.p0 || n1 < 1 || n2 < 1 || n1 + n2 > D || (!d.i0 && d.n2 < 0) || (!d.i1 && d.n3 < 0)) return B2B_EINVAL;
  if (n1 > 128 || n2 > 128) return B2B_EUNSUPPORTED;
  if (N == 0) {
    cudaMemsetAsync(Wbar, 0, sizeof(float) * (size_t)2 * n1 * n2, stream);
    return (int)cudaMemsetAsync(cbar, 0, sizeof(float) * 2 * n1, stream);
  }
  if (!x || !ybar || !xbar || ldx < D || ldybar < D || ldxbar < D) return B2B_EINVAL;
  const size_t need = b2b_coupling_affine_vjp_workspace_bytes(n1, n2);
  if (!workspace || workspace_bytes < need) return B2B_EWORKSPACE;
  char* wsb = static_cast<char*>(workspace);
  wsb += (256 - (reinterpret_cast<uintptr_t>(wsb) & 255)) & 255;
  CvParams P;
  P.x = x;
  P.ybar = ybar;
  P.ljbar = ljbar;
  P.xbar = xbar;
  P.W = d.p0;
  P.c = d.p1;
  P.idx1 = d.i0;
  P.idx2 = d.i1;
  P.part = reinterpret_cast<float*>(wsb);
  P.N = N;
  P.ldx = ldx;
  P.ldyb = ldybar;
  P.ldxb = ldxbar;
  P.D = D;
  P.n1 = n1;
  P.n2 = n2;
  P.row1 = d.n2;
  P.row2 = d.n3;
  const long long tiles = (N + CV_TC - 1) / CV_TC;
  long long grid = sm_count();
  if (grid > tiles) grid = tiles;
  const size_t smem = ((size_t)2 * D * CV_LD + (size_t)2 * n1 * CV_LD + CV_TC) * sizeof(float) + (size_t)(n1 + n2) * sizeof(int);
  if (smem > 220 * 1024) return B2B_EUNSUPPORTED;
  auto kernel = d.inverse ? coupling_vjp_kernel<true> : coupling_vjp_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kernel<<<(int)grid, CV_THREADS, smem, stream>>>(P);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  const int len0 = 2 * n1 * n2, len = len0 + 2 * n1;
  partial_sum_kernel<<<(len + 255) / 256, 256, 0, stream>>>(P.part, (int)grid, len, Wbar, len0, cbar);
  return (int)cudaGetLastError();
}

extern "C" size_t b2b_batchnorm_eval_vjp_workspace_bytes(int32_t D) {
  if (D < 1 || D > 1024) return 0;
  return (size_t)b2b::sm_count() * 2 * (size_t)(2 * D + 1) * sizeof(float) + 256;
}

extern "C" int b2b_batchnorm_eval_vjp_f32(const b2b_layer_desc* layer, const float* x, const float* ybar, const float* ljbar,
                                          float* xbar, float* bbar, float* logsbar, int32_t D, int64_t N, int64_t ldx,
                                          int64_t ldybar, int64_t ldxbar, void* workspace, size_t workspace_bytes,
                                          void* stream_) {
  using namespace b2b;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!layer || layer->kind != B2B_BATCHNORM || D < 1 || N < 0 || !bbar || !logsbar) return B2B_EINVAL;
  const b2b_layer_desc& d = *layer;
  if (!d.p0 || !d.p1 || !d.p2 || !d.p3) return B2B_EINVAL;
  if (D > 1024) return B2B_EUNSUPPORTED;
  if (N == 0) {
    cudaMemsetAsync(bbar, 0, sizeof(float) * D, stream);
    return (int)cudaMemsetAsync(logsbar, 0, sizeof(float) * D, stream);
  }
  if (!x || !ybar || !xbar || ldx < D || ldybar < D || ldxbar < D) return B2B_EINVAL;
  const size_t need = b2b_batchnorm_eval_vjp_workspace_bytes(D);
  if (!workspace || workspace_bytes < need) return B2B_EWORKSPACE;
  char* wsb = static_cast<char*>(workspace);
  wsb += (256 - (reinterpret_cast<uintptr_t>(wsb) & 255)) & 255;
  BvParams P;
  P.x = x;
  P.ybar = ybar;
  P.ljbar = ljbar;
  P.xbar = xbar;
  P.b = d.p0;
  P.logs = d.p1;
  P.m = d.p2;
  P.v = d.p3;
  P.eps = d.f0;
  P.part = reinterpret_cast<float*>(wsb);
  P.N = N;
  P.ldx = ldx;
  P.ldyb = ldybar;
  P.ldxb = ldxbar;
  P.D = D;
  P.inverse = d.inverse ? 1 : 0;
  long long grid = (long long)sm_count() * 2;
  const long long want = (N + 7) / 8;
  if (grid > want) grid = want;
  const size_t smem = (size_t)8 * (2 * D + 1) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(bn_eval_vjp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  bn_eval_vjp_kernel<<<(int)grid, 256, smem, stream>>>(P);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  bn_vjp_finalize_kernel<<<1, 256, 0, stream>>>(P.part, (int)grid, D, P.inverse, bbar, logsbar);
  return (int)cudaGetLastError();
}
