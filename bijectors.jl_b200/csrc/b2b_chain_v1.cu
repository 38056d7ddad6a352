// v1 fused column-local chain kernel: TMA-staged, THREAD-PER-COLUMN.
//
// Why: in the v0 (lane-group) kernel an 8-layer chain costs ~250 warp-instructions per column (cross-lane
// reductions, selects, broadcasts, redundant transcendental work) and the kernel is issue-bound at ~45 % of
// the HBM roofline.  Here one thread owns one whole column in registers, so a row reduction is a private
// FFMA chain (no shuffles), per-column scalars (tanh, log1p, find_alpha) are computed exactly once, and
// every layer parameter read is a warp-uniform shared-memory broadcast.
//
// (For D >= 128 a column is split over TPC = 2 or 4 adjacent lanes -- 64 rows each -- so that 16 warps fit
// in the register file; the only cross-lane traffic is one or two shuffles per row reduction.)
//
// Data movement: the D x N batch is described by a 2-D TMA tensor map {D, N}.  A tile is 32/TPC columns;
// it is fetched as D/32 boxes of {32 floats, 32/TPC columns} with the 128-byte swizzle, i.e. box q holds rows
// [32q, 32q+32) of the tile's columns, one column per 128-byte line, and the 16-byte chunk k of line t lives
// at chunk slot k ^ (t & 7).  A thread therefore reads its rows with conflict-free LDS.128
// (the 8 lanes of a quarter-warp hit 8 different 16-byte bank groups) while the register <-> row mapping
// stays identical in all lanes (so parameters are warp-uniform).  Results go back through a per-warp
// staging buffer in the same layout and a TMA store.
//
// Pipeline (persistent CTA, one per SM): tiles j = 0,1,2,... of a CTA are consumed round-robin by its NW
// warps; tile j lands in input buffer j % P.  The warp that consumed tile j copies it to registers and
// IMMEDIATELY re-arms the same buffer with the load of tile j + P (its lane 0 issues the TMA), so P tiles
// (P*16 KB at D=128) are always in flight per SM with no producer warp and no empty-barriers: refills of a
// buffer are ordered by the consumption of its previous tile.
//
// Reference semantics per layer: see b2b_chain_v0.cu / b2b_device.cuh (file:line cited there).
#include "b2b_v1_pipeline.cuh"

namespace b2b {

template <int D, int TPC, int CPT>
__device__ __forceinline__ void planar_apply(float2 (&x)[CPT][D / TPC / 2], const ColCtx<D, TPC>& c, const float* sp,
                                             bool inverse, float (&lj)[CPT]) {
  using C = ColCtx<D, TPC>;
  const float4* w4 = reinterpret_cast<const float4*>(sp);
  const float4* u4 = reinterpret_cast<const float4*>(sp + D);
  const float cc_ = sp[2 * D], bb = sp[2 * D + 1];
  float2 acc[CPT][4];
  B2B_FOR_COLS {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[cc][i] = make_float2(0.f, 0.f);
  }
  B2B_FOR_SLOTS {
    const float4 w = w4[c.prm(ql, r)];
    const int i = (ql * 8 + r) * 2;
    B2B_FOR_COLS {
      acc[cc][(r & 1) * 2 + 0] = __ffma2_rn(make_float2(w.x, w.y), x[cc][i], acc[cc][(r & 1) * 2 + 0]);
      acc[cc][(r & 1) * 2 + 1] = __ffma2_rn(make_float2(w.z, w.w), x[cc][i + 1], acc[cc][(r & 1) * 2 + 1]);
    }
  }
  float2 t2[CPT];
  B2B_FOR_COLS {
    const float2 s = __fadd2_rn(__fadd2_rn(acc[cc][0], acc[cc][1]), __fadd2_rn(acc[cc][2], acc[cc][3]));
    const float wz = part_sum<TPC>(s.x + s.y);  // aT_b(w, z), utils.jl:2
    float t, s2;
    if (!inverse) {
      tanh_sech2(wz + bb, t, s2);
      lj[cc] += log1pf(cc_ * s2);  // planar_layer.jl:107
    } else {
      find_alpha_ts(wz, cc_, bb, t, s2);  // planar_layer.jl:121; t = tanh(α+b), s2 = sech²(α+b)
      lj[cc] -= log1pf(cc_ * s2);
      t = -t;
    }
    t2[cc] = make_float2(t, t);
  }
  B2B_FOR_SLOTS {
    const float4 u = u4[c.prm(ql, r)];
    const int i = (ql * 8 + r) * 2;
    B2B_FOR_COLS {
      x[cc][i] = __ffma2_rn(make_float2(u.x, u.y), t2[cc], x[cc][i]);  // planar_layer.jl:78 / :124
      x[cc][i + 1] = __ffma2_rn(make_float2(u.z, u.w), t2[cc], x[cc][i + 1]);
    }
  }
}

template <int D, int TPC, int CPT>
__device__ __forceinline__ void batchnorm_apply(float2 (&x)[CPT][D / TPC / 2], const ColCtx<D, TPC>& c,
                                                const float* sp, bool inverse, float (&lj)[CPT]) {
  using C = ColCtx<D, TPC>;
  // staged as y = A·x + C (fwd) / x = iA·y + iC (inverse): normalise.jl:66 / :84 with the constants folded
  const float4* A4 = reinterpret_cast<const float4*>(sp + (inverse ? 2 * D : 0));
  const float4* C4 = reinterpret_cast<const float4*>(sp + (inverse ? 3 * D : D));
  B2B_FOR_SLOTS {
    const float4 a = A4[c.prm(ql, r)], k = C4[c.prm(ql, r)];
    const int i = (ql * 8 + r) * 2;
    B2B_FOR_COLS {
      x[cc][i] = __ffma2_rn(x[cc][i], make_float2(a.x, a.y), make_float2(k.x, k.y));
      x[cc][i + 1] = __ffma2_rn(x[cc][i + 1], make_float2(a.z, a.w), make_float2(k.z, k.w));
    }
  }
  const float ljc = sp[4 * D];
  B2B_FOR_COLS lj[cc] += inverse ? -ljc : ljc;
}

template <int D, int TPC, int CPT>
__device__ __forceinline__ void rqs_apply(float2 (&x)[CPT][D / TPC / 2], const ColCtx<D, TPC>& c, const float* sp,
                                          int K1, bool inverse, float (&lj)[CPT]) {
  using C = ColCtx<D, TPC>;
  const int KP = rqs_kp(K1);
  float acc[CPT];
  B2B_FOR_COLS acc[cc] = 0.f;
  B2B_FOR_SLOTS {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int row = c.row(ql, r, e);
      B2B_FOR_COLS {
        float& xe = (e & 1) ? x[cc][(ql * 8 + r) * 2 + (e >> 1)].y : x[cc][(ql * 8 + r) * 2 + (e >> 1)].x;
        float o = xe, l1 = 0.f;
        if (inverse) rqs_element<true>(sp, K1, KP, D, row, xe, o, l1);
        else rqs_element<false>(sp, K1, KP, D, row, xe, o, l1);
        xe = o;
        acc[cc] += l1;
      }
    }
  }
  B2B_FOR_COLS lj[cc] += part_sum<TPC>(acc[cc]);  // sum over dimensions, rational_quadratic_spline.jl:304-309
}

template <int D, int TPC, int CPT>
__device__ __forceinline__ void stacked_apply(float2 (&x)[CPT][D / TPC / 2], const ColCtx<D, TPC>& c, const float* sp,
                                              bool inverse, float (&lj)[CPT]) {
  using C = ColCtx<D, TPC>;
  const int* code = reinterpret_cast<const int*>(sp);
  const float* av = sp + D;
  float acc[CPT];
  B2B_FOR_COLS acc[cc] = 0.f;
  B2B_FOR_SLOTS {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int row = c.row(ql, r, e);
      const int op = code[row];
      const float a = av[row], b = av[D + row];
      B2B_FOR_COLS {
        float& xe = (e & 1) ? x[cc][(ql * 8 + r) * 2 + (e >> 1)].y : x[cc][(ql * 8 + r) * 2 + (e >> 1)].x;
        xe = ew_apply(op, inverse, a, b, xe, acc[cc]);
      }
    }
  }
  B2B_FOR_COLS lj[cc] += part_sum<TPC>(acc[cc]);
}

// Program 1: the layer-descriptor interpreter (parameters staged in shared memory from DEVICE pointers).
template <int D, int TPC, int CPT>
struct InterpProg {
  using State = V1NoState;
  const B2BChainParams& P;
  __device__ __forceinline__ void stage(float* params, int warp, int lane, int nw) const {
    // RQS tables are large: all threads of the CTA build them; every other layer is staged by one warp
    for (int l = 0; l < P.L; ++l)
      if (P.layers[l].kind == B2B_RQS) stage_rqs_cta(P.layers[l], params + P.soff[l], D, D, warp * 32 + lane, nw * 32);
    for (int l = warp; l < P.L; l += nw)
      if (P.layers[l].kind != B2B_RQS) stage_layer(P.layers[l], params + P.soff[l], D, D, lane);
  }
  __device__ __forceinline__ void apply(float2 (&x)[CPT][D / TPC / 2], const ColCtx<D, TPC>& ctx, const float* params,
                                        float (&lj)[CPT]) const {
    using C = ColCtx<D, TPC>;
#pragma unroll 1
    for (int l = 0; l < P.L; ++l) {
      const b2b_layer_desc& d = P.layers[l];
      const float* sp = params + P.soff[l];
      switch (d.kind) {
        case B2B_PLANAR: planar_apply<D, TPC, CPT>(x, ctx, sp, d.inverse != 0, lj); break;
        case B2B_RADIAL: radial_apply<D, TPC, CPT>(x, ctx, sp, d.inverse != 0, lj); break;
        case B2B_BATCHNORM: batchnorm_apply<D, TPC, CPT>(x, ctx, sp, d.inverse != 0, lj); break;
        case B2B_RQS:
          if constexpr (C::EPT * CPT <= 64) rqs_apply<D, TPC, CPT>(x, ctx, sp, d.n0, d.inverse != 0, lj);
          break;
        case B2B_STACKED_EW:
          if constexpr (C::EPT * CPT <= 64) stacked_apply<D, TPC, CPT>(x, ctx, sp, d.inverse != 0, lj);
          break;
        case B2B_MVNORMAL_DIAG: mvnormal_apply<D, TPC, CPT>(x, ctx, sp, lj); break;
        default: break;
      }
    }
  }
};

template <int D, int TPC, int CPT, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
    chain_v1_kernel(const __grid_constant__ B2BChainParams P, const __grid_constant__ V1Extra E,
                    const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y) {
  const InterpProg<D, TPC, CPT> prog{P};
  v1_run<D, TPC, CPT, NW>(P, E, map_x, map_y, prog);
}

typedef void (*v1_kernel_t)(const B2BChainParams, const V1Extra, const CUtensorMap, const CUtensorMap);

struct V1Plan : V1Geom {
  v1_kernel_t kernel;
};

static int plan_v1(B2BChainParams& p, V1Plan& plan) {
  const int D = p.D;
  if (!(D == 32 || D == 64 || D == 128 || D == 256)) return B2B_EUNSUPPORTED;
  if (v1_check_io(p) != 0) return B2B_EUNSUPPORTED;
  int off = 0;
  for (int l = 0; l < p.L; ++l) {
    const int k = p.layers[l].kind;
    if (k == B2B_PERMUTE || k == B2B_COUPLING_AFFINE) return B2B_EUNSUPPORTED;
    p.soff[l] = off;
    off += (b2b_layer_smem_floats(p.layers[l], D) + 3) & ~3;
  }
  p.scratch_off = -1;
  bool per_row = false;  // RQS / Stacked are unrolled per row: only built for <= 64 rows per thread
  for (int l = 0; l < p.L; ++l) per_row |= p.layers[l].kind == B2B_RQS || p.layers[l].kind == B2B_STACKED_EW;
  // <D, lanes per column, columns per thread, warps>: warps are chosen so that the per-thread register budget
  // (65536 / threads) holds the fragment without spilling: 64 data registers need ~170 (12 warps), 128 need 255
  // (8 warps), 32 fit in 128 (16 warps).  B2B_V1_CFG selects alternative builds for experiments.
  static const int cfg = getenv("B2B_V1_CFG") ? atoi(getenv("B2B_V1_CFG")) : 0;
  int nw, tpc, cpt = 1;
  if (D == 256) { plan.kernel = chain_v1_kernel<256, 4, 1, 12>; nw = 12; tpc = 4; }
  else if (D == 128) {
    // default: one thread per column (measured fastest: 4.4-4.5 G samples/s on C2).  Chains with per-row layers
    // (RQS, Stacked) use two lanes per column.  B2B_V1_CFG=2218 selects the build with two lanes per column and
    // TWO columns per thread (every parameter load serves two columns: LDS wavefronts -36 %, but more scalar work
    // per warp; measured 4.06 G samples/s), 2112 the 2-lane / 12-warp build (3.7 G samples/s).
    if (cfg == 2218 && !per_row) { plan.kernel = chain_v1_kernel<128, 2, 2, 8>; nw = 8; tpc = 2; cpt = 2; }
    else if (cfg == 2112 || per_row) { plan.kernel = chain_v1_kernel<128, 2, 1, 12>; nw = 12; tpc = 2; }
    else { plan.kernel = chain_v1_kernel<128, 1, 1, 8>; nw = 8; tpc = 1; }
  }
  else if (D == 64) {
    if (cfg == 1128 && !per_row) { plan.kernel = chain_v1_kernel<64, 1, 2, 8>; nw = 8; tpc = 1; cpt = 2; }
    else { plan.kernel = chain_v1_kernel<64, 1, 1, 12>; nw = 12; tpc = 1; }
  }
  else { plan.kernel = chain_v1_kernel<32, 1, 1, 16>; nw = 16; tpc = 1; }
  return v1_geometry(D, p.N, nw, (32 / tpc) * cpt, (size_t)off, plan);
}

}  // namespace b2b

int b2b_chain_grid_size_v1(const B2BChainParams& p) {
  B2BChainParams q = p;
  b2b::V1Plan plan;
  if (b2b::plan_v1(q, plan) != 0) return 0;
  return plan.grid;
}

int b2b_launch_chain_v1(const B2BChainParams& p, cudaStream_t stream) {
  using namespace b2b;
  B2BChainParams q = p;
  V1Plan plan;
  const int rc = plan_v1(q, plan);
  if (rc != 0) return rc;
  CUtensorMap mx, my;
  if (!make_maps(q, plan.cols, &mx, &my, &plan.extra.tma3d)) return B2B_EUNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(plan.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
  if (e != cudaSuccess) return (int)e;
  plan.kernel<<<plan.grid, plan.nw * 32, plan.smem, stream>>>(q, plan.extra, mx, my);
  return (int)cudaGetLastError();
}

