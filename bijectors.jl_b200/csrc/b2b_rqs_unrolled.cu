// A single RationalQuadraticSpline layer with K = 8 bins (K1 = 9 knots, the BASELINE configuration) as a program
// specialised on (D, K1, direction) on the TMA pipeline: the bin search is four unrolled compare/select steps with
// immediate offsets and every table address is a compile-time constant.  The layer interpreter needs ~98 instructions
// per element (runtime knot count, generic addressing); this program ~60: C4 forward 27 % -> 40 % of the HBM roofline
// (the layer stays compute-bound per element).  Tables are staged by the whole CTA (stage_rqs_cta).
//
// Reference semantics: rational_quadratic_spline.jl:317-357 (forward), :183-220 (inverse), see rqs_element.
#include "b2b_v1_pipeline.cuh"

namespace b2b {

template <int D, int K1, bool INV>
struct RqsProg {
  using State = V1NoState;
  const B2BChainParams& P;
  __device__ __forceinline__ void stage(float* params, int warp, int lane, int nw) const {
    stage_rqs_cta(P.layers[0], params, D, D, warp * 32 + lane, nw * 32);
  }
  __device__ __forceinline__ void apply(float2 (&x)[1][D / 2], const ColCtx<D, 1>&, const float* params,
                                        float (&lj)[1]) const {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < D / 2; ++i) {
      float o, l1;
      rqs_element<INV, K1>(params, K1, rqs_kp(K1), D, 2 * i, x[0][i].x, o, l1);
      x[0][i].x = o;
      acc += l1;
      rqs_element<INV, K1>(params, K1, rqs_kp(K1), D, 2 * i + 1, x[0][i].y, o, l1);
      x[0][i].y = o;
      acc += l1;
    }
    lj[0] += acc;  // sum over dimensions, rational_quadratic_spline.jl:304-309
  }
};

template <int D, int K1, int NW, bool INV>
__global__ void __launch_bounds__(NW * 32, 1)
    rqs_unrolled_kernel(const __grid_constant__ B2BChainParams P, const __grid_constant__ V1Extra E,
                        const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y) {
  const RqsProg<D, K1, INV> prog{P};
  v1_run<D, 1, 1, NW>(P, E, map_x, map_y, prog);
}

template <int D, int NW>
static int launch_rqs(const B2BChainParams& q, cudaStream_t stream) {
  constexpr int K1 = 9;
  V1Geom g;
  const int rc = v1_geometry(D, q.N, NW, 32, (size_t)b2b_layer_smem_floats(q.layers[0], D), g);
  if (rc != 0) return rc;
  CUtensorMap mx, my;
  if (!make_maps(q, g.cols, &mx, &my, &g.extra.tma3d)) return B2B_EUNSUPPORTED;
  auto kernel = q.layers[0].inverse ? rqs_unrolled_kernel<D, K1, NW, true> : rqs_unrolled_kernel<D, K1, NW, false>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return (int)e;
  kernel<<<g.grid, NW * 32, g.smem, stream>>>(q, g.extra, mx, my);
  return (int)cudaGetLastError();
}

}  // namespace b2b

// p: a segment that is exactly one RQS layer with 9 knots, D in {32, 64}; B2B_EUNSUPPORTED otherwise
int b2b_rqs_unrolled_applicable(const B2BChainParams& p) {
  return p.L == 1 && p.layers[0].kind == B2B_RQS && p.layers[0].n0 == 9 && (p.D == 32 || p.D == 64) &&
         b2b::v1_check_io(p) == 0;
}

int b2b_launch_rqs_unrolled(const B2BChainParams& p, cudaStream_t stream) {
  using namespace b2b;
  if (!b2b_rqs_unrolled_applicable(p)) return B2B_EUNSUPPORTED;
  B2BChainParams q = p;
  q.scratch_off = -1;
  if (q.D == 64) return launch_rqs<64, 12>(q, stream);
  return launch_rqs<32, 16>(q, stream);
}
