// Chains of PlanarLayers with their (derived) parameters in the CONSTANT BANK.
//
// Why: in the layer interpreter (b2b_chain_v1.cu) every parameter element is a warp-uniform LDS broadcast and the
// 8-layer D = 128 headline chain is bound by LSU wavefronts (~70 % of the HBM roofline).  A constant-bank operand
// reaches FFMA2 through a uniform register (LDCU -> UR) and costs no LSU work, but the first-level constant cache
// only holds ~4 KB.  So: chains whose w and û fit in 4 KB read both from the constant bank (MODE 0); the 8-layer
// D = 128 chain (8 KB) keeps w in shared memory and û in the constant bank (MODE 2), splitting the operand traffic
// over both paths -- 86 % of the roofline instead of 72 %.
//
// Two parameter sources share the per-tile program:
//   ArgSrc  parameters in HOST memory (b2b_planar_chain_hostparams_f32): derived on the host, passed BY VALUE as
//           kernel arguments (bank 0).  No device-side preparation at all.
//   SymSrc  parameters in DEVICE memory (b2b_chain_run_f32 segments made of PlanarLayers only): a one-CTA kernel
//           derives û / wᵀû into a staging buffer, cudaMemcpyToSymbolAsync moves it into a __constant__ array
//           (stream-ordered, so the constant cache is coherent), then the main kernel runs.  The symbol is per-device
//           library state: launches that use it are serialised by an event (same stream: free).
//
// Reference semantics: planar_layer.jl:65-80 (get_u_hat, forward), :102-110 (logabsdetjac), :112-127 + :160-185
// (inverse through find_alpha).
#include "b2b_planar_common.cuh"

namespace b2b {

template <int D, int L>
struct PlanarHP {
  float v[2 * L * D + 2 * L];
};

template <int D, int L>
struct ArgSrc {
  const PlanarHP<D, L>& H;
  int invmask;
  __device__ __forceinline__ float w(int l, int i) const { return H.v[l * D + i]; }
  __device__ __forceinline__ float uh(int l, int i) const { return H.v[L * D + l * D + i]; }
  __device__ __forceinline__ float c(int l) const { return H.v[2 * L * D + l]; }
  __device__ __forceinline__ float b(int l) const { return H.v[2 * L * D + L + l]; }
  __device__ __forceinline__ bool inv(int l) const { return (invmask >> l) & 1; }
  __device__ __forceinline__ float raw(int i) const { return H.v[i]; }  // dynamic index: staging only
};

// MODE 0: w and û from the constant bank; MODE 1: û staged in shared memory (LDS broadcast), w from the constant
// bank; MODE 2: w in shared memory, û from the constant bank.
// DIR 0: every layer forward, 1: every layer inverse, 2: per-layer direction from the mask.  The unrolled program of
// 8 layers is large; carrying the (unused) root-finder of the other direction in the hot path costs a quarter of the
// forward throughput in instruction-cache misses, so the pure directions get their own kernels.
// MVN: the chain ends in the base MvNormal log-density (P.layers[0] holds its descriptor): logpdf(td, y).
template <int D, int L, int MODE, int DIR, bool MVN, class Src>
struct PlanarConstProg {
  using State = V1NoState;
  const Src src;
  const B2BChainParams& P;
  static constexpr int MVN_OFF = MODE ? L * D : 0;
  __device__ __forceinline__ void stage(float* params, int warp, int lane, int nw) const {
    if (MVN && warp == nw - 1) stage_layer(P.layers[0], params + MVN_OFF, D, D, lane);
    if constexpr (MODE != 0) {
      for (int i = warp * 32 + lane; i < L * D; i += nw * 32) params[i] = src.raw((MODE == 1 ? L * D : 0) + i);
    }
  }
  __device__ __forceinline__ void apply(float2 (&x)[1][D / 2], const ColCtx<D, 1>& ctx, const float* params,
                                        float (&lj)[1]) const {
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const float4* sp4 = reinterpret_cast<const float4*>(params + l * D);
      float2 acc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < D / 4; ++i) {
        float4 w;
        if (MODE == 2) w = sp4[i];
        else w = make_float4(src.w(l, 4 * i), src.w(l, 4 * i + 1), src.w(l, 4 * i + 2), src.w(l, 4 * i + 3));
        acc[(i & 1) * 2 + 0] = __ffma2_rn(make_float2(w.x, w.y), x[0][2 * i], acc[(i & 1) * 2 + 0]);
        acc[(i & 1) * 2 + 1] = __ffma2_rn(make_float2(w.z, w.w), x[0][2 * i + 1], acc[(i & 1) * 2 + 1]);
      }
      const float2 s = __fadd2_rn(__fadd2_rn(acc[0], acc[1]), __fadd2_rn(acc[2], acc[3]));
      const float wz = s.x + s.y;  // aT_b(w, z), utils.jl:2
      const float cc_ = src.c(l), bb = src.b(l);
      float t, s2;
      if (DIR == 0 || (DIR == 2 && !src.inv(l))) {
        tanh_sech2(wz + bb, t, s2);
        lj[0] += log1pf(cc_ * s2);  // planar_layer.jl:107
      } else {
        find_alpha_ts(wz, cc_, bb, t, s2);  // planar_layer.jl:121; t = tanh(α+b), s2 = sech²(α+b)
        lj[0] -= log1pf(cc_ * s2);
        t = -t;
      }
      const float2 t2 = make_float2(t, t);
#pragma unroll
      for (int i = 0; i < D / 4; ++i) {
        float4 u;
        if (MODE == 1) u = sp4[i];
        else u = make_float4(src.uh(l, 4 * i), src.uh(l, 4 * i + 1), src.uh(l, 4 * i + 2), src.uh(l, 4 * i + 3));
        x[0][2 * i] = __ffma2_rn(make_float2(u.x, u.y), t2, x[0][2 * i]);  // planar_layer.jl:78 / :124
        x[0][2 * i + 1] = __ffma2_rn(make_float2(u.z, u.w), t2, x[0][2 * i + 1]);
      }
    }
    if (MVN) mvnormal_apply<D, 1, 1>(x, ctx, params + MVN_OFF, lj);
  }
};

template <int D, int L, int NW, int MODE, int DIR>
__global__ void __launch_bounds__(NW * 32, 1)
    planar_arg_kernel(const __grid_constant__ B2BChainParams P, const __grid_constant__ V1Extra E,
                      const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                      const __grid_constant__ PlanarHP<D, L> H, const int invmask) {
  const PlanarConstProg<D, L, MODE, DIR, false, ArgSrc<D, L>> prog{{H, invmask}, P};
  v1_run<D, 1, 1, NW>(P, E, map_x, map_y, prog);
}

template <int D, int L, int NW, int MODE, int DIR, bool MVN>
__global__ void __launch_bounds__(NW * 32, 1)
    planar_sym_kernel(const __grid_constant__ B2BChainParams P, const __grid_constant__ V1Extra E,
                      const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                      const float* stage, const int invmask) {
  const PlanarConstProg<D, L, MODE, DIR, MVN, SymSrc<D, L>> prog{{stage, invmask}, P};
  v1_run<D, 1, 1, NW>(P, E, map_x, map_y, prog);
}

// ---- host side -----------------------------------------------------------------------------------------
template <int D, int L, int NW, int MODE, int DIR>
static int launch_arg(const B2BChainParams& q, const V1Geom& g, const CUtensorMap& mx, const CUtensorMap& my,
                      const float* packed, int invmask, cudaStream_t stream) {
  static PlanarHP<D, L> H;  // copied into the launch's argument buffer by <<<>>>
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  memcpy(H.v, packed, sizeof(H.v));
  auto kernel = planar_arg_kernel<D, L, NW, MODE, DIR>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return (int)e;
  kernel<<<g.grid, NW * 32, g.smem, stream>>>(q, g.extra, mx, my, H, invmask);
  return (int)cudaGetLastError();
}

template <int D, int L, int NW, int MODE, int DIR, bool MVN = false>
static int launch_sym(const B2BChainParams& q, const V1Geom& g, const CUtensorMap& mx, const CUtensorMap& my,
                      const float* stage, int invmask, cudaStream_t stream) {
  auto kernel = planar_sym_kernel<D, L, NW, MODE, DIR, MVN>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return (int)e;
  kernel<<<g.grid, NW * 32, g.smem, stream>>>(q, g.extra, mx, my, stage, invmask);
  return (int)cudaGetLastError();
}

// dispatch over (D, L, MODE, DIR) for either source
template <bool SYM, int D, int NW, int LL, int MM>
static int dispatch_dir(const B2BChainParams& q, const V1Geom& g, const CUtensorMap& mx, const CUtensorMap& my,
                        const float* params, int invmask, cudaStream_t stream) {
  const int all = (1 << LL) - 1;
  const int dir = (invmask & all) == 0 ? 0 : ((invmask & all) == all ? 1 : 2);
  if (q.L == 1) {  // terminal MvNormal (device-resident parameters, all-inverse chains only)
    if constexpr (SYM) {
      if (dir == 1) return launch_sym<D, LL, NW, MM, 1, true>(q, g, mx, my, params, invmask, stream);
    }
    return B2B_EUNSUPPORTED;
  }
#define B2B_HP_DIR(DD)                                                                                   \
  if (dir == DD)                                                                                         \
    return SYM ? launch_sym<D, LL, NW, MM, DD>(q, g, mx, my, params, invmask, stream)                    \
               : launch_arg<D, LL, NW, MM, DD>(q, g, mx, my, params, invmask, stream);
  B2B_HP_DIR(0)
  B2B_HP_DIR(1)
  B2B_HP_DIR(2)
#undef B2B_HP_DIR
  return B2B_EUNSUPPORTED;
}

template <bool SYM, int D, int NW>
static int dispatch_L(int L, int mode, const B2BChainParams& q, const V1Geom& g, const CUtensorMap& mx,
                      const CUtensorMap& my, const float* params, int invmask, cudaStream_t stream) {
  if (L == 1 && mode == 0) return dispatch_dir<SYM, D, NW, 1, 0>(q, g, mx, my, params, invmask, stream);
  if (L == 2 && mode == 0) return dispatch_dir<SYM, D, NW, 2, 0>(q, g, mx, my, params, invmask, stream);
  if (L == 4 && mode == 0) return dispatch_dir<SYM, D, NW, 4, 0>(q, g, mx, my, params, invmask, stream);
  if constexpr (2 * D * 8 * 4 > 4096) {
    if (L == 8 && mode == 2) return dispatch_dir<SYM, D, NW, 8, 2>(q, g, mx, my, params, invmask, stream);
  } else {
    if (L == 8 && mode == 0) return dispatch_dir<SYM, D, NW, 8, 0>(q, g, mx, my, params, invmask, stream);
  }
  return B2B_EUNSUPPORTED;
}

template <bool SYM>
static int launch_planar_const(const B2BChainParams& p, int L, const float* params, int invmask, cudaStream_t stream) {
  B2BChainParams q = p;  // q.L: 0, or 1 when q.layers[0] is the terminal MvNormal
  q.scratch_off = -1;
  if (!(q.D == 32 || q.D == 64 || q.D == 128)) return B2B_EUNSUPPORTED;
  if (v1_check_io(q) != 0) return B2B_EUNSUPPORTED;
  const HPShape sh = hp_shape(q.D, L);
  V1Geom g;
  const int rc = v1_geometry(q.D, q.N, sh.nw, 32, (sh.mode ? (size_t)L * q.D : 0) + (q.L ? 2 * q.D + 4 : 0), g);
  if (rc != 0) return rc;
  CUtensorMap mx, my;
  if (!make_maps(q, g.cols, &mx, &my, &g.extra.tma3d)) return B2B_EUNSUPPORTED;
  if (q.D == 128) return dispatch_L<SYM, 128, 8>(L, sh.mode, q, g, mx, my, params, invmask, stream);
  if (q.D == 64) return dispatch_L<SYM, 64, 12>(L, sh.mode, q, g, mx, my, params, invmask, stream);
  return dispatch_L<SYM, 32, 16>(L, sh.mode, q, g, mx, my, params, invmask, stream);
}

}  // namespace b2b

int b2b_planar_const_grid_size(const B2BChainParams& p) {
  using namespace b2b;
  if (!(p.D == 32 || p.D == 64 || p.D == 128)) return 0;
  V1Geom g;
  const HPShape sh = hp_shape(p.D, 8);
  if (v1_geometry(p.D, p.N, sh.nw, 32, 0, g) != 0) return 0;
  return g.grid;
}

// `L` (1, 2, 4 or 8) planar layers, derived parameters packed for (D, L) in HOST memory -> kernel arguments
int b2b_launch_planar_hostparams(const B2BChainParams& p, int L, const float* packed, int invmask,
                                 cudaStream_t stream) {
  B2BChainParams q = p;
  q.L = 0;
  return b2b::launch_planar_const<false>(q, L, packed, invmask, stream);
}

// Applicability of the constant-bank path to a fusable segment: 1..8 PlanarLayers, optionally followed by the
// terminal MvNormal when every planar layer is inverse (= logpdf(td, y)); D in {32,64,128}; 16-byte aligned batches.
// Returns the number of planar layers, 0 when not applicable.
int b2b_planar_const_layers(const B2BChainParams& p) {
  using namespace b2b;
  if (!(p.D == 32 || p.D == 64 || p.D == 128) || p.L < 1) return 0;
  const bool mvn = p.layers[p.L - 1].kind == B2B_MVNORMAL_DIAG;
  const int n = p.L - (mvn ? 1 : 0);
  if (n < 1 || n > HP_MAX_L) return 0;
  for (int l = 0; l < n; ++l) {
    if (p.layers[l].kind != B2B_PLANAR) return 0;
    if (mvn && !p.layers[l].inverse) return 0;
  }
  if (v1_check_io(p) != 0) return 0;
  return n;
}

// prep kernel -> staging buffer -> __constant__ symbol -> main kernel.  B2B_EUNSUPPORTED when not applicable.
int b2b_launch_planar_chain_const(const B2BChainParams& p, cudaStream_t stream) {
  using namespace b2b;
  const int n = b2b_planar_const_layers(p);
  if (n == 0) return B2B_EUNSUPPORTED;
  int invmask = 0;
  for (int l = 0; l < n; ++l)
    if (p.layers[l].inverse) invmask |= 1 << l;
  int Lp = 1;
  while (Lp < n) Lp <<= 1;
  // identity padding is its own inverse: an all-inverse chain stays all-inverse (the single-direction kernel)
  if (invmask == (1 << n) - 1) invmask = (1 << Lp) - 1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return B2B_EUNSUPPORTED;
  SlotState& st = g_slots[dev];
  std::lock_guard<std::mutex> lock(st.mu);
  cudaError_t e;
  {
    const int rcp = planar_slot_prepare(st, p, n, Lp, stream);
    if (rcp != 0) return rcp;
  }
  B2BChainParams q = p;
  q.L = 0;
  if (p.L > n) {
    q.L = 1;
    q.layers[0] = p.layers[p.L - 1];
  }
  const int rc = launch_planar_const<true>(q, Lp, st.stage, invmask, stream);
  if (rc != B2B_OK) return rc;
  if ((e = cudaEventRecord(st.free_ev, stream)) != cudaSuccess) return (int)e;
  return B2B_OK;
}

