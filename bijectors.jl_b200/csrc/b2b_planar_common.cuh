// Shared by b2b_planar_const.cu (forward / inverse evaluation: launch shapes only) and b2b_planar_vjp.cu (reverse mode:
// the constant-bank parameter slot, its staging buffer and the device-side get_u_hat preparation).  Every translation
// unit that includes this header owns ITS OWN copy of the __constant__ slot (and of the event that serialises it).
#pragma once
#include <cstring>

#include "b2b_v1_pipeline.cuh"

namespace b2b {

constexpr int HP_MAX_L = 8;
constexpr int HP_MAX_D = 128;
// packed layout for (D, L): w[L][D] | û[L][D] | c[L] | b[L]
constexpr int HP_MAX_FLOATS = 2 * HP_MAX_L * HP_MAX_D + 2 * HP_MAX_L;

static __constant__ float c_planar[HP_MAX_FLOATS];
static __device__ float g_planar_stage[HP_MAX_FLOATS];

template <int D, int L>
struct SymSrc {
  const float* stage;  // the same packed parameters in global memory (source of the shared-memory half)
  int invmask;
  __device__ __forceinline__ float w(int l, int i) const { return c_planar[l * D + i]; }
  __device__ __forceinline__ float uh(int l, int i) const { return c_planar[L * D + l * D + i]; }
  __device__ __forceinline__ float c(int l) const { return c_planar[2 * L * D + l]; }
  __device__ __forceinline__ float b(int l) const { return c_planar[2 * L * D + L + l]; }
  __device__ __forceinline__ bool inv(int l) const { return (invmask >> l) & 1; }
  __device__ __forceinline__ float raw(int i) const { return stage[i]; }
};

// get_u_hat (planar_layer.jl:65-70) for Lp layers (the last Lp - L are identity padding), one warp per layer,
// packed for (D, Lp) into `out`.
static __global__ void __launch_bounds__(HP_MAX_L * 32)
    planar_prep_kernel(const __grid_constant__ B2BChainParams P, int L, int Lp, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, l = threadIdx.x >> 5, D = P.D;
  if (l >= Lp) return;
  auto put = [&](int idx, float v) { out[idx] = v; };
  const int wo = l * D, uo = Lp * D + l * D;
  if (l >= L) {
    for (int i = lane; i < D; i += 32) {
      put(wo + i, 0.f);
      put(uo + i, 0.f);
    }
    if (lane == 0) {
      put(2 * Lp * D + l, 0.f);
      put(2 * Lp * D + Lp + l, 0.f);
    }
    return;
  }
  const b2b_layer_desc& d = P.layers[l];
  float s = 0.f, q = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float w = d.p0[i], u = d.p1[i];
    s = fmaf(w, u, s);
    q = fmaf(w, w, q);
  }
  s = warp_sum(s);
  q = warp_sum(q);
  const float k = (softplus(-s) - 1.0f) / q;  // planar_layer.jl:67
  for (int i = lane; i < D; i += 32) {
    const float w = d.p0[i];
    put(wo + i, w);
    put(uo + i, fmaf(k, w, d.p1[i]));
  }
  if (lane == 0) {
    put(2 * Lp * D + l, softplus(s) - 1.0f);  // wᵀû, planar_layer.jl:68
    put(2 * Lp * D + Lp + l, d.p2[0]);        // first(flow.b), :75
  }
}

// ---- host side: launch shapes, slot state ------------------------------------------------------------------
struct HPShape {
  int nw, mode;
};

// warps per CTA as in the interpreter (register budget); MODE 2 when w and û exceed ~4 KB of constants
static HPShape hp_shape(int D, int L) {
  HPShape s;
  s.nw = D == 128 ? 8 : (D == 64 ? 12 : 16);
  s.mode = (2 * D * L * 4 > 4096) ? 2 : 0;
  return s;
}

// per-device state of the __constant__ slot
struct SlotState {
  std::mutex mu;
  cudaEvent_t free_ev = nullptr;
  float* stage = nullptr;
};
static SlotState g_slots[64];


// Under the slot's mutex: wait for the previous user, derive the parameters of p.layers[0..n) (padded to Lp layers)
// into the staging buffer and copy them into the constant bank.  The caller launches its kernels and then records
// st.free_ev.
static inline int planar_slot_prepare(SlotState& st, const B2BChainParams& p, int n, int Lp, cudaStream_t stream) {
  cudaError_t e;
  if (!st.free_ev) {
    if ((e = cudaEventCreateWithFlags(&st.free_ev, cudaEventDisableTiming)) != cudaSuccess) return (int)e;
    if ((e = cudaGetSymbolAddress(reinterpret_cast<void**>(&st.stage), g_planar_stage)) != cudaSuccess) return (int)e;
    if ((e = cudaEventRecord(st.free_ev, stream)) != cudaSuccess) return (int)e;
  }
  // the previous user of the slot (possibly on another stream) must have finished
  if ((e = cudaStreamWaitEvent(stream, st.free_ev, 0)) != cudaSuccess) return (int)e;
  planar_prep_kernel<<<1, HP_MAX_L * 32, 0, stream>>>(p, n, Lp, st.stage);
  if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
  const size_t bytes = sizeof(float) * (size_t)(2 * Lp * p.D + 2 * Lp);
  return (int)cudaMemcpyToSymbolAsync(c_planar, st.stage, bytes, 0, cudaMemcpyDeviceToDevice, stream);
}

}  // namespace b2b
