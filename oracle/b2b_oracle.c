/*
 * CPU ORACLE in C (test infrastructure, NOT product code): float32 restatement of the reference's CPU
 * path for the hot-path layers, keeping the reference's PASS STRUCTURE (per layer: a BLAS-style wᵀz pass,
 * then a broadcast pass into a freshly allocated output -- SURVEY.md §3.1), so that it can stand in as
 * the timed CPU baseline ("restatement of the reference CPU path").  OpenMP over columns
 * models BLAS threads / all host cores; nthreads = 1 models Julia's single-threaded broadcast.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline, --impl reference) may load this.
 * It is cross-checked against oracle/oracle_np.py (which is pinned by the reference's golden vectors)
 * in tests/test_oracle_c.py.
 *
 * Layout: Julia column-major D x N: column n at x + n*D.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float log1pexpf_(float x) { /* LogExpFunctions.log1pexp */
  return x > 0.f ? x + log1pf(expf(-x)) : log1pf(expf(x));
}

int oracle_num_procs(void) { return omp_get_num_procs(); }

/* with_logabsdet_jacobian(::PlanarLayer, z::Matrix)  -- src/bijectors/planar_layer.jl:65-80,102-110 */
void oracle_planar_fwd_f32(const float* w, const float* u, float b, const float* z, float* y, float* logjac,
                           int D, int64_t N, int nthreads) {
  /* get_u_hat, :65-70 */
  float wTu = 0.f, q = 0.f;
  for (int i = 0; i < D; ++i) { wTu += w[i] * u[i]; q += w[i] * w[i]; }
  float* uh = (float*)malloc(sizeof(float) * D);
  const float k = (log1pexpf_(-wTu) - 1.f) / q;
  for (int i = 0; i < D; ++i) uh[i] = u[i] + k * w[i];
  const float wTuh = log1pexpf_(wTu) - 1.f;
  float* wTz = (float*)malloc(sizeof(float) * (size_t)N); /* aT_b(w, z): 1 x N, pass #1 (utils.jl:2) */
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    const float* zc = z + (size_t)n * D;
    float s = 0.f;
    for (int i = 0; i < D; ++i) s += w[i] * zc[i];
    wTz[n] = s;
  }
  /* z .+ û .* tanh.(wT_z .+ b): pass #2 into a fresh output (:78) */
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    const float* zc = z + (size_t)n * D;
    float* yc = y + (size_t)n * D;
    const float t = tanhf(wTz[n] + b);
    for (int i = 0; i < D; ++i) yc[i] = zc[i] + uh[i] * t;
  }
  /* log1p.(wT_û .* abs2.(sech.(wT_z .+ b))): pass #3 (:107) */
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    const float sech = 1.f / coshf(wTz[n] + b);
    logjac[n] = log1pf(wTuh * sech * sech);
  }
  free(wTz);
  free(uh);
}

/* with_logabsdet_jacobian(::RadialLayer, z::Matrix) -- src/bijectors/radial_layer.jl:43-53,58-72 */
void oracle_radial_fwd_f32(float alpha_raw, float beta, const float* z0, const float* z, float* y,
                           float* logjac, int D, int64_t N, int nthreads) {
  const float alpha = log1pexpf_(alpha_raw);
  const float bhat = -alpha + log1pexpf_(beta);
  float* r = (float*)malloc(sizeof(float) * (size_t)N);
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t n = 0; n < N; ++n) { /* sqrt.(sum(abs2, z .- z_0; dims=1)), :49 */
    const float* zc = z + (size_t)n * D;
    float s = 0.f;
    for (int i = 0; i < D; ++i) { const float d = zc[i] - z0[i]; s += d * d; }
    r[n] = sqrtf(s);
  }
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t n = 0; n < N; ++n) { /* z .+ β̂ ./ (α .+ r') .* (z .- z_0), :51 */
    const float* zc = z + (size_t)n * D;
    float* yc = y + (size_t)n * D;
    const float g = bhat / (alpha + r[n]);
    for (int i = 0; i < D; ++i) yc[i] = zc[i] + g * (zc[i] - z0[i]);
  }
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t n = 0; n < N; ++n) { /* :68-70 */
    const float h = 1.f / (alpha + r[n]);
    logjac[n] = (float)(D - 1) * logf(1.f + bhat * h) + logf(1.f + bhat * h + bhat * (-(h * h)) * r[n]);
  }
  free(r);
}

/* rqs_forward per element -- src/bijectors/rational_quadratic_spline.jl:317-357; tables are D x K1
 * column-major (the struct fields), mapped over the columns of x.
 * The reference evaluates transform and logabsdetjac separately (:363-367), i.e. two bin searches. */
static void rqs_elem(const float* W, const float* H, const float* Dv, int D, int K1, int i, float x, float* y,
                     float* lj) {
  const float Bw = W[(size_t)(K1 - 1) * D + i];
  if (x <= -Bw || x >= Bw) { *y = x; *lj = 0.f; return; }
  int k = 0;
  for (int q = 0; q < K1; ++q) k += W[(size_t)q * D + i] < x; /* searchsortedfirst - 1 */
  const float w_k = k == 0 ? -Bw : W[(size_t)(k - 1) * D + i];
  const float w = W[(size_t)k * D + i] - w_k;
  const float h_k = k == 0 ? -H[(size_t)(K1 - 1) * D + i] : H[(size_t)(k - 1) * D + i];
  const float dy = H[(size_t)k * D + i] - h_k;
  const float s = dy / w, xi = (x - w_k) / w;
  const float d_k = k == 0 ? 1.f : Dv[(size_t)(k - 1) * D + i];
  const float d_k1 = k == K1 - 1 ? 1.f : Dv[(size_t)k * D + i];
  const float den = s + (d_k1 + d_k - 2.f * s) * xi * (1.f - xi);
  const float num = s * s * (d_k1 * xi * xi + 2.f * s * xi * (1.f - xi) + d_k * (1.f - xi) * (1.f - xi));
  *lj = logf(num) - 2.f * logf(den);
  *y = h_k + dy * (s * xi * xi + d_k * xi * (1.f - xi)) / den;
}

void oracle_rqs_fwd_f32(const float* W, const float* H, const float* Dv, int K1, const float* x, float* y,
                        float* logjac, int D, int64_t N, int nthreads) {
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    const float* xc = x + (size_t)n * D;
    float* yc = y + (size_t)n * D;
    float tot = 0.f;
    for (int i = 0; i < D; ++i) {
      float l;
      rqs_elem(W, H, Dv, D, K1, i, xc[i], &yc[i], &l);
      tot += l;
    }
    logjac[n] = tot;
  }
}

/* InvertibleBatchNorm eval forward -- src/bijectors/normalise.jl:61-67 */
void oracle_batchnorm_fwd_f32(const float* b, const float* logs, const float* m, const float* v, float eps,
                              const float* x, float* y, float* logjac, int D, int64_t N, int nthreads) {
  float lj = 0.f;
  for (int i = 0; i < D; ++i) lj += logs[i] - logf(v[i] + eps) / 2.f;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    const float* xc = x + (size_t)n * D;
    float* yc = y + (size_t)n * D;
    for (int i = 0; i < D; ++i) yc[i] = expf(logs[i]) * (xc[i] - m[i]) / sqrtf(v[i] + eps) + b[i];
    logjac[n] = lj;
  }
}

/* inverse of the above -- normalise.jl:74-86 */
void oracle_batchnorm_inv_f32(const float* b, const float* logs, const float* m, const float* v, float eps,
                              const float* y, float* x, float* logjac, int D, int64_t N, int nthreads) {
  float lj = 0.f;
  for (int i = 0; i < D; ++i) lj += logs[i] - logf(v[i] + eps) / 2.f;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    const float* yc = y + (size_t)n * D;
    float* xc = x + (size_t)n * D;
    for (int i = 0; i < D; ++i) xc[i] = (yc[i] - b[i]) / expf(logs[i]) * sqrtf(v[i] + eps) + m[i];
    logjac[n] = -lj;
  }
}

/* Affine coupling, forward (inverse = 0) or inverse -- src/bijectors/coupling.jl:206-228 with the law
 * Shift(t) ∘ Scale(exp.(s)), [s;t] = W x₂ + c; W is (2 n1 x n2) column-major; idx 0-based. */
void oracle_coupling_affine_f32(const int32_t* idx1, int n1, const int32_t* idx2, int n2, const float* W,
                                const float* c, int inverse, const float* x, float* y, float* logjac, int D,
                                int64_t N, int nthreads) {
#pragma omp parallel num_threads(nthreads)
  {
    float* st = (float*)malloc(sizeof(float) * 2 * (size_t)n1);
#pragma omp for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
      const float* xc = x + (size_t)n * D;
      float* yc = y + (size_t)n * D;
      for (int r = 0; r < 2 * n1; ++r) st[r] = c ? c[r] : 0.f;
      for (int k = 0; k < n2; ++k) {
        const float x2 = xc[idx2[k]];
        const float* wk = W + (size_t)k * 2 * n1;
        for (int r = 0; r < 2 * n1; ++r) st[r] += wk[r] * x2;
      }
      if (y != x) memcpy(yc, xc, sizeof(float) * D);
      float tot = 0.f;
      for (int j = 0; j < n1; ++j) {
        const float s = st[j], t = st[n1 + j];
        if (!inverse) yc[idx1[j]] = expf(s) * xc[idx1[j]] + t;
        else yc[idx1[j]] = (1.f / expf(s)) * (xc[idx1[j]] + (-t));
        tot += s;
      }
      logjac[n] = inverse ? -tot : tot;
    }
    free(st);
  }
}

/* logpdf(MvNormal(mu, Diagonal(sigma^2)), x) per column (Distributions/PDMats) */
void oracle_mvnormal_diag_logpdf_f32(const float* mu, const float* sigma, const float* x, float* out, int D,
                                     int64_t N, int nthreads) {
  float ls = 0.f;
  for (int i = 0; i < D; ++i) ls += sigma ? logf(sigma[i] * sigma[i]) : 0.f;
  const float cst = -((float)D * 1.8378770664093453f + ls) / 2.f;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    const float* xc = x + (size_t)n * D;
    float q = 0.f;
    for (int i = 0; i < D; ++i) {
      const float z = (xc[i] - (mu ? mu[i] : 0.f)) / (sigma ? sigma[i] : 1.f);
      q += z * z;
    }
    out[n] = cst - q / 2.f;
  }
}

/* y[n] += a[n] : the `logjac + logjac` of chained layers (vector + vector) */
void oracle_add_f32(float* y, const float* a, int64_t N, int nthreads) {
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t n = 0; n < N; ++n) y[n] += a[n];
}
