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
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>

#include "b2b_device.cuh"

namespace b2b {

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "B2B_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra B2B_DONE;\n"
      "bra B2B_WAIT;\n"
      "B2B_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(src)
               : "memory");
}
// 3-D forms: the tensor map views a batch as {32 floats, N columns, D/32 row-blocks} so that ONE instruction moves
// a whole [row-block][column][32 floats] tile (the 128-byte swizzle limits the innermost box extent to 32 floats)
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2), "r"(src)
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void flag_store_release(int* p, int v) {
  asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int flag_load_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}

struct V1Extra {
  int n_in;        // input ring depth P
  int param_off;   // byte offset of the staged parameters in dynamic smem
  int bar_off;     // byte offset of the mbarriers
  int nwarps;
  int tma3d;       // 1: the tensor maps are 3-D (one TMA instruction per tile), 0: 2-D (one per 32-row block)
  long long tiles;
};

// Register layout of one thread: it owns rows [h*EPT, (h+1)*EPT) (h = part index, TPC parts per column) of CPT
// columns, as float2 pairs so that the packed sm_100 FP32 pipe (FFMA2: two fp32 results per issue slot) does
// the per-row work.  Every layer parameter that is loaded from shared memory is used for all CPT columns of the
// thread: at D = 128 (TPC = 2, CPT = 2) that halves the LDS wavefronts per column, which is what bounds the
// one-column-per-thread mapping.  Box ql (of NQT = EPT/32 boxes) slot r holds the LOGICAL 16-byte chunk r ^ rot
// (rot = h * 8/TPC keeps the TPC parts of a quarter-warp on different bank groups); pair index =
// (ql*8 + r)*2 + {0,1}.  prm() returns the float4 index of the parameters matching slot (ql, r).
template <int D, int TPC>
struct ColCtx {
  static constexpr int EPT = D / TPC;
  static constexpr int NQT = EPT / 32;
  int h, rot;
  __device__ __forceinline__ int prm(int ql, int r) const { return (h * NQT + ql) * 8 + (r ^ rot); }
  __device__ __forceinline__ int row(int ql, int r, int e) const { return prm(ql, r) * 4 + e; }
};

template <int TPC>
__device__ __forceinline__ float part_sum(float v) {
#pragma unroll
  for (int o = 1; o < TPC; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#define B2B_FOR_SLOTS                        \
  _Pragma("unroll") for (int ql = 0; ql < C::NQT; ++ql) _Pragma("unroll") for (int r = 0; r < 8; ++r)
#define B2B_FOR_COLS _Pragma("unroll") for (int cc = 0; cc < CPT; ++cc)

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
      const float alpha = find_alpha(wz, cc_, bb);  // planar_layer.jl:121
      tanh_sech2(alpha + bb, t, s2);
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
__device__ __forceinline__ void radial_apply(float2 (&x)[CPT][D / TPC / 2], const ColCtx<D, TPC>& c, const float* sp,
                                             bool inverse, float (&lj)[CPT]) {
  using C = ColCtx<D, TPC>;
  const float4* z4 = reinterpret_cast<const float4*>(sp);
  const float alpha = sp[D], bhat = sp[D + 1], apb = sp[D + 2];
  const float2 m1 = make_float2(-1.f, -1.f);
  float2 acc[CPT][4];
  B2B_FOR_COLS {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[cc][i] = make_float2(0.f, 0.f);
  }
  B2B_FOR_SLOTS {
    const float4 z0 = z4[c.prm(ql, r)];
    const int i = (ql * 8 + r) * 2;
    B2B_FOR_COLS {
      const float2 d0 = __ffma2_rn(make_float2(z0.x, z0.y), m1, x[cc][i]);
      const float2 d1 = __ffma2_rn(make_float2(z0.z, z0.w), m1, x[cc][i + 1]);
      acc[cc][(r & 1) * 2 + 0] = __ffma2_rn(d0, d0, acc[cc][(r & 1) * 2 + 0]);
      acc[cc][(r & 1) * 2 + 1] = __ffma2_rn(d1, d1, acc[cc][(r & 1) * 2 + 1]);
    }
  }
  float2 g2[CPT];
  B2B_FOR_COLS {
    const float2 s = __fadd2_rn(__fadd2_rn(acc[cc][0], acc[cc][1]), __fadd2_rn(acc[cc][2], acc[cc][3]));
    const float nrm = sqrtf(part_sum<TPC>(s.x + s.y));  // radial_layer.jl:49 / :125
    float r_ = nrm;
    if (inverse) {
      const float a = apb - nrm;  // radial_layer.jl:126-127
      const float sq = sqrtf(fmaf(a, a, 4.0f * alpha * nrm));
      r_ = a > 0.f ? (2.0f * alpha * nrm) / (sq + a) : 0.5f * (sq - a);
    }
    const float hh = 1.0f / (alpha + r_);
    const float bh = bhat * hh;
    const float ljf = (float)(D - 1) * log1pf(bh) + log1pf(bh * alpha * hh);  // radial_layer.jl:68-70
    float g;
    if (!inverse) {
      g = bh;
      lj[cc] += ljf;
    } else {
      g = -bhat / (apb + r_);  // (α+r)/(α+β̂+r) − 1, radial_layer.jl:96
      lj[cc] -= ljf;
    }
    g2[cc] = make_float2(g, g);
  }
  B2B_FOR_SLOTS {
    const float4 z0 = z4[c.prm(ql, r)];
    const int i = (ql * 8 + r) * 2;
    B2B_FOR_COLS {
      const float2 d0 = __ffma2_rn(make_float2(z0.x, z0.y), m1, x[cc][i]);
      const float2 d1 = __ffma2_rn(make_float2(z0.z, z0.w), m1, x[cc][i + 1]);
      x[cc][i] = __ffma2_rn(g2[cc], d0, x[cc][i]);
      x[cc][i + 1] = __ffma2_rn(g2[cc], d1, x[cc][i + 1]);
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
      int op = code[row];
      const float a = av[row];
      if (inverse) op = op == B2B_EW_EXP ? B2B_EW_LOG : (op == B2B_EW_LOG ? B2B_EW_EXP : op);
      B2B_FOR_COLS {
        float& xe = (e & 1) ? x[cc][(ql * 8 + r) * 2 + (e >> 1)].y : x[cc][(ql * 8 + r) * 2 + (e >> 1)].x;
        const float xv = xe;
        if (op == B2B_EW_EXP) {
          xe = expf(xv);
          acc[cc] += xv;
        } else if (op == B2B_EW_LOG) {
          const float lg = logf(xv);
          xe = lg;
          acc[cc] -= lg;
        } else if (op == B2B_EW_SHIFT) {
          xe = inverse ? xv - a : a + xv;
        } else if (op == B2B_EW_SCALE) {
          xe = inverse ? xv / a : a * xv;
          const float la = logf(fabsf(a));
          acc[cc] += inverse ? -la : la;
        } else if (op == B2B_EW_LEAKY_RELU) {
          const float al = inverse ? 1.0f / a : a;  // leaky_relu.jl:16,18-22
          if (xv < 0.f) {
            xe = al * xv;
            acc[cc] += logf(fabsf(al));
          }
        }
      }
    }
  }
  B2B_FOR_COLS lj[cc] += part_sum<TPC>(acc[cc]);
}

template <int D, int TPC, int CPT>
__device__ __forceinline__ void mvnormal_apply(const float2 (&x)[CPT][D / TPC / 2], const ColCtx<D, TPC>& c,
                                               const float* sp, float (&lj)[CPT]) {
  using C = ColCtx<D, TPC>;
  const float4* mu4 = reinterpret_cast<const float4*>(sp);
  const float4* is4 = reinterpret_cast<const float4*>(sp + D);
  const float2 m1 = make_float2(-1.f, -1.f);
  float2 acc[CPT][2];
  B2B_FOR_COLS acc[cc][0] = acc[cc][1] = make_float2(0.f, 0.f);
  B2B_FOR_SLOTS {
    const float4 mu = mu4[c.prm(ql, r)], is = is4[c.prm(ql, r)];
    const int i = (ql * 8 + r) * 2;
    B2B_FOR_COLS {
      const float2 z0 = __fmul2_rn(__ffma2_rn(make_float2(mu.x, mu.y), m1, x[cc][i]), make_float2(is.x, is.y));
      const float2 z1 = __fmul2_rn(__ffma2_rn(make_float2(mu.z, mu.w), m1, x[cc][i + 1]), make_float2(is.z, is.w));
      acc[cc][0] = __ffma2_rn(z0, z0, acc[cc][0]);
      acc[cc][1] = __ffma2_rn(z1, z1, acc[cc][1]);
    }
  }
  B2B_FOR_COLS {
    const float2 s = __fadd2_rn(acc[cc][0], acc[cc][1]);
    lj[cc] += sp[2 * D] - 0.5f * part_sum<TPC>(s.x + s.y);
  }
}

// The pipeline (TMA tile ring, register-resident fragments, per-warp TMA store) is independent of WHAT is applied
// to the fragments: `prog.stage()` prepares per-CTA state, `prog.apply()` maps the fragments and accumulates logjac.
template <int D, int TPC, int CPT, int NW, class Prog>
__device__ __forceinline__ void v1_run(const B2BChainParams& P, const V1Extra& E, const CUtensorMap& map_x,
                                       const CUtensorMap& map_y, const Prog& prog) {
  using C = ColCtx<D, TPC>;
  constexpr int NQ = D / 32;                 // boxes per tile
  constexpr int LPC = 32 / TPC;              // lane groups per warp
  constexpr int COLS = LPC * CPT;            // columns per tile (= per warp)
  constexpr int BOX_BYTES = COLS * 128;      // COLS lines of 128 B
  constexpr int TILE_BYTES = NQ * BOX_BYTES;
  extern __shared__ unsigned char smem_dyn[];
  // the 128-byte swizzle pattern repeats every 1024 B: align the tile area by hand (1 KB of slack is allocated)
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* in_base = smem_raw;                                    // n_in tiles
  unsigned char* out_base = smem_raw + (size_t)E.n_in * TILE_BYTES;     // NW tiles
  float* params = reinterpret_cast<float*>(smem_raw + E.param_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + E.bar_off);
  // armed[b] = index j of the tile whose load has been issued into input buffer b.  A warp may only wait on
  // bars[b] for tile j once armed[b] == j: an mbarrier parity wait is only meaningful one phase ahead, and
  // with P < NW a warp could otherwise be two phases ahead of the buffer it shares with another warp.
  // (flag hand-off between warps: st.release / ld.acquire at CTA scope; compute-sanitizer's racecheck reports the
  // polling load against the releasing store -- that pairing is the synchronisation itself)
  int* armed = reinterpret_cast<int*>(bars + 8);

  // the warp index is made provably warp-uniform: tile / buffer / barrier addresses then live in uniform registers
  // and the TMA instructions take them directly (no per-instruction R2UR + BRA.U.ANY uniformisation loop)
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  auto load_tile = [&](uint32_t dst, int col0, uint32_t bar) {
    if (E.tma3d) {
      tma_load_3d(dst, &map_x, 0, col0, 0, bar);
    } else {
#pragma unroll
      for (int q = 0; q < NQ; ++q) tma_load_2d(dst + q * BOX_BYTES, &map_x, q * 32, col0, bar);
    }
  };
  prog.stage(params, warp, lane, NW);
  if (threadIdx.x == 0) {
    for (int i = 0; i < E.n_in; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_mbar_init();
  }
  __syncthreads();

  // tiles of this CTA: global tile id = blockIdx.x + j * gridDim.x
  const long long my_tiles = (E.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  if (threadIdx.x == 0) {
    for (int j = 0; j < E.n_in && j < my_tiles; ++j) {
      const uint32_t bar = smem_u32(&bars[j]);
      mbar_expect_tx(bar, TILE_BYTES);
      const long long tile = blockIdx.x + (long long)j * gridDim.x;
      load_tile(smem_u32(in_base + (size_t)j * TILE_BYTES), (int)(tile * COLS), bar);
      flag_store_release(&armed[j], j);
    }
  }
  __syncthreads();

  C ctx;
  const int t = lane / TPC;  // lane group: columns t, t + LPC, ... of the tile
  ctx.h = lane % TPC;
  ctx.rot = ctx.h * (8 / TPC);
  unsigned char* my_out = out_base + (size_t)warp * TILE_BYTES;
  // byte XOR: physical slot = r ^ rot ^ (line & 7); LPC is a multiple of 8, so it is the same for all CPT columns
  const int sw = (((t & 7) ^ ctx.rot) & 7) * 16;
  const int line = t * 128 + ctx.h * C::NQT * BOX_BYTES;  // this thread's first line inside its first box
  double dsum = 0.0;
  bool store_pending = false;

  for (long long j = warp; j < my_tiles; j += NW) {
    const int buf = (int)(j % E.n_in);
    const uint32_t parity = (uint32_t)((j / E.n_in) & 1);
    const long long tile = blockIdx.x + j * gridDim.x;
    const long long col = tile * COLS + t;
    while (flag_load_acquire(&armed[buf]) != (int)j) __nanosleep(20);
    mbar_wait(smem_u32(&bars[buf]), parity);

    float2 x[CPT][C::EPT / 2];
    {
      const unsigned char* src = in_base + (size_t)buf * TILE_BYTES + line;
      B2B_FOR_COLS {
        B2B_FOR_SLOTS {
          const float4 v =
              *reinterpret_cast<const float4*>(src + cc * (LPC * 128) + ql * BOX_BYTES + ((r * 16) ^ sw));
          x[cc][(ql * 8 + r) * 2] = make_float2(v.x, v.y);
          x[cc][(ql * 8 + r) * 2 + 1] = make_float2(v.z, v.w);
        }
      }
    }
    __syncwarp();
    // re-arm this input buffer with the tile P steps ahead
    if (lane == 0 && j + E.n_in < my_tiles) {
      const uint32_t bar = smem_u32(&bars[buf]);
      mbar_expect_tx(bar, TILE_BYTES);
      const long long nt = blockIdx.x + (j + E.n_in) * gridDim.x;
      load_tile(smem_u32(in_base + (size_t)buf * TILE_BYTES), (int)(nt * COLS), bar);
      flag_store_release(&armed[buf], (int)(j + E.n_in));
    }

    float lj[CPT];
    B2B_FOR_COLS {
      const long long cl = col + cc * LPC;
      lj[cc] = (P.accumulate && P.logjac && cl < P.N) ? P.logjac[cl] : 0.0f;
    }
    prog.apply(x, ctx, params, lj);

    if (P.y) {
      if (store_pending) {
        if (lane == 0) tma_wait_read0();  // previous store of this warp has finished reading my_out
        __syncwarp();
      }
      unsigned char* dst = my_out + line;
      B2B_FOR_COLS {
        B2B_FOR_SLOTS {
          const float2 a = x[cc][(ql * 8 + r) * 2], b = x[cc][(ql * 8 + r) * 2 + 1];
          *reinterpret_cast<float4*>(dst + cc * (LPC * 128) + ql * BOX_BYTES + ((r * 16) ^ sw)) =
              make_float4(a.x, a.y, b.x, b.y);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (E.tma3d) {
          tma_store_3d(&map_y, 0, (int)(tile * COLS), 0, smem_u32(my_out));
        } else {
#pragma unroll
          for (int q = 0; q < NQ; ++q)
            tma_store_2d(&map_y, q * 32, (int)(tile * COLS), smem_u32(my_out + q * BOX_BYTES));
        }
        tma_commit();
      }
      store_pending = true;
    }
    if (ctx.h == 0) {
      B2B_FOR_COLS {
        const long long cl = col + cc * LPC;
        if (cl < P.N) {
          if (P.logjac) P.logjac[cl] = lj[cc];
          dsum += (double)lj[cc];
        }
      }
    }
  }
  if (lane == 0 && store_pending) tma_wait_all0();  // smem must stay valid until the stores have drained

  if (P.partials) {
    __shared__ double red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    if (lane == 0) red[warp] = dsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tt = 0.0;
      for (int w = 0; w < NW; ++w) tt += red[w];
      P.partials[blockIdx.x] = tt;
    }
  }
}

// Program 1: the layer-descriptor interpreter (parameters staged in shared memory from DEVICE pointers).
template <int D, int TPC, int CPT>
struct InterpProg {
  const B2BChainParams& P;
  __device__ __forceinline__ void stage(float* params, int warp, int lane, int nw) const {
    for (int l = warp; l < P.L; l += nw) stage_layer(P.layers[l], params + P.soff[l], D, D, lane);
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

// Program 2: a chain of L PlanarLayers whose (derived) parameters arrive BY VALUE as kernel arguments, i.e. in the
// constant bank: every w / û element is read with a uniform LDCU into a uniform register and used directly as an
// FFMA2 operand -- no shared-memory traffic for parameters at all (the LSU broadcast of parameters is what bounds the
// interpreter at ~70 % of the roofline).  This is the reference's own parameter residency: PlanarLayer fields are
// host Arrays (planar_layer.jl:13-18); û and wᵀû (get_u_hat, :65-70) are computed on the host at launch time.
template <int D, int L>
struct PlanarHP {
  float w[L][D];
  float uh[L][D];
  float c[L];
  float b[L];
  int inverse;
};

template <int D, int L, int U>
struct PlanarHPProg {
  const PlanarHP<D, L>& H;
  __device__ __forceinline__ void stage(float*, int, int, int) const {}
  __device__ __forceinline__ void apply(float2 (&x)[1][D / 2], const ColCtx<D, 1>&, const float*, float (&lj)[1]) const {
    // U layers are unrolled; the outer loop is kept rolled (constant-bank addresses indexed by a uniform register)
#pragma unroll 1
    for (int l0 = 0; l0 < L; l0 += U)
#pragma unroll
    for (int lu = 0; lu < U; ++lu) {
      const int l = l0 + lu;
      float2 acc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < D / 4; ++i) {
        acc[(i & 1) * 2 + 0] = __ffma2_rn(make_float2(H.w[l][4 * i], H.w[l][4 * i + 1]), x[0][2 * i], acc[(i & 1) * 2 + 0]);
        acc[(i & 1) * 2 + 1] =
            __ffma2_rn(make_float2(H.w[l][4 * i + 2], H.w[l][4 * i + 3]), x[0][2 * i + 1], acc[(i & 1) * 2 + 1]);
      }
      const float2 s = __fadd2_rn(__fadd2_rn(acc[0], acc[1]), __fadd2_rn(acc[2], acc[3]));
      const float wz = s.x + s.y;  // aT_b(w, z), utils.jl:2
      const float cc_ = H.c[l], bb = H.b[l];
      float t, s2;
      if (!H.inverse) {
        tanh_sech2(wz + bb, t, s2);
        lj[0] += log1pf(cc_ * s2);  // planar_layer.jl:107
      } else {
        const float alpha = find_alpha(wz, cc_, bb);  // planar_layer.jl:121
        tanh_sech2(alpha + bb, t, s2);
        lj[0] -= log1pf(cc_ * s2);
        t = -t;
      }
      const float2 t2 = make_float2(t, t);
#pragma unroll
      for (int i = 0; i < D / 4; ++i) {
        x[0][2 * i] = __ffma2_rn(make_float2(H.uh[l][4 * i], H.uh[l][4 * i + 1]), t2, x[0][2 * i]);  // :78 / :124
        x[0][2 * i + 1] = __ffma2_rn(make_float2(H.uh[l][4 * i + 2], H.uh[l][4 * i + 3]), t2, x[0][2 * i + 1]);
      }
    }
  }
};

template <int D, int L, int U, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
    planar_hp_kernel(const __grid_constant__ B2BChainParams P, const __grid_constant__ V1Extra E,
                     const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                     const __grid_constant__ PlanarHP<D, L> H) {
  const PlanarHPProg<D, L, U> prog{H};
  v1_run<D, 1, 1, NW>(P, E, map_x, map_y, prog);
}

// ---- host side -----------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
  static encode_tiled_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
  });
  return fn;
}

static bool make_map(CUtensorMap* m, const float* base, int D, long long N, long long ld, int cols, bool three_d) {
  encode_tiled_fn enc = get_encode();
  if (!enc) return false;
  const cuuint32_t estr[3] = {1, 1, 1};
  if (three_d) {
    // {32 floats, N columns, D/32 row-blocks}: the row-block stride (128 B) is SMALLER than the column stride
    const cuuint64_t dims[3] = {32, (cuuint64_t)N, (cuuint64_t)(D / 32)};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(float), 128};
    const cuuint32_t box[3] = {32, (cuuint32_t)cols, (cuuint32_t)(D / 32)};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)N};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {32, (cuuint32_t)cols};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Both maps of a launch; 3-D when the driver accepts them (B2B_V1_TMA=2 forces the 2-D form)
static bool make_maps(const B2BChainParams& q, int cols, CUtensorMap* mx, CUtensorMap* my, int* tma3d) {
  static const int force2d = getenv("B2B_V1_TMA") && atoi(getenv("B2B_V1_TMA")) == 2;
  for (int three_d = force2d ? 0 : 1; three_d >= 0; --three_d) {
    if (three_d && q.D == 32) continue;  // one row-block: the 2-D form already is one instruction
    bool ok = make_map(mx, q.x, q.D, q.N, q.ldx, cols, three_d != 0);
    if (ok && q.y) ok = make_map(my, q.y, q.D, q.N, q.ldy, cols, three_d != 0);
    if (ok) {
      if (!q.y) *my = *mx;
      *tma3d = three_d;
      return true;
    }
  }
  return false;
}

typedef void (*v1_kernel_t)(const B2BChainParams, const V1Extra, const CUtensorMap, const CUtensorMap);

struct V1Plan {
  v1_kernel_t kernel;
  int nw, grid, cols;
  size_t smem;
  V1Extra extra;
};

static int plan_v1(B2BChainParams& p, V1Plan& plan, int nw_override = 0) {
  const int D = p.D;
  if (!(D == 32 || D == 64 || D == 128 || D == 256)) return B2B_EUNSUPPORTED;
  if (p.N >= (1ll << 31) - 64) return B2B_EUNSUPPORTED;
  if ((p.ldx % 4) || (reinterpret_cast<uintptr_t>(p.x) & 15)) return B2B_EUNSUPPORTED;
  if (p.y && ((p.ldy % 4) || (reinterpret_cast<uintptr_t>(p.y) & 15))) return B2B_EUNSUPPORTED;
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
  if (nw_override) nw = nw_override;
  plan.cols = (32 / tpc) * cpt;
  const int tile_bytes = D * 4 * plan.cols;
  const size_t param_bytes = (size_t)off * sizeof(float);
  const size_t budget = 225 * 1024;
  const size_t fixed = (size_t)nw * tile_bytes + ((param_bytes + 15) & ~(size_t)15) + 16 * sizeof(uint64_t) + 1024;
  if (fixed + 2 * (size_t)tile_bytes > budget) return B2B_EUNSUPPORTED;
  int n_in = (int)((budget - fixed) / tile_bytes);
  if (n_in > 8) n_in = 8;
  plan.nw = nw;
  plan.extra.n_in = n_in;
  plan.extra.nwarps = nw;
  plan.extra.param_off = (n_in + nw) * tile_bytes;
  plan.extra.bar_off = plan.extra.param_off + (int)((param_bytes + 15) & ~(size_t)15);
  plan.extra.tiles = (p.N + plan.cols - 1) / plan.cols;
  plan.smem = (size_t)plan.extra.bar_off + 16 * sizeof(uint64_t) + 1024;  // +1024: base alignment slack
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = sms;
  const long long want = (plan.extra.tiles + nw - 1) / nw;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  plan.grid = (int)grid;
  return 0;
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

// ---- host-parameter planar chains ------------------------------------------------------------------------
namespace b2b {

template <int D, int L, int U, int NW>
static int launch_planar_hp(const B2BChainParams& q, const V1Plan& plan, const CUtensorMap& mx, const CUtensorMap& my,
                            const float* w, const float* uh, const float* c, const float* b, int inverse,
                            cudaStream_t stream) {
  static PlanarHP<D, L> H;  // 2·L·D+2·L+1 floats; filled and copied into the launch's parameter buffer
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  memcpy(H.w, w, sizeof(float) * L * D);
  memcpy(H.uh, uh, sizeof(float) * L * D);
  memcpy(H.c, c, sizeof(float) * L);
  memcpy(H.b, b, sizeof(float) * L);
  H.inverse = inverse;
  auto kernel = planar_hp_kernel<D, L, U, NW>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
  if (e != cudaSuccess) return (int)e;
  kernel<<<plan.grid, NW * 32, plan.smem, stream>>>(q, plan.extra, mx, my, H);  // arguments are copied at launch
  return (int)cudaGetLastError();
}

template <int D, int NW>
static int launch_planar_hp_L(int L, const B2BChainParams& q, const V1Plan& plan, const CUtensorMap& mx,
                              const CUtensorMap& my, const float* w, const float* uh, const float* c, const float* b,
                              int inverse, cudaStream_t stream) {
  static const int u_env = getenv("B2B_HP_U") ? atoi(getenv("B2B_HP_U")) : 0;
  if (D == 128 && L == 8) {
    if (u_env == 1) return launch_planar_hp<D, 8, 1, NW>(q, plan, mx, my, w, uh, c, b, inverse, stream);
    if (u_env == 2) return launch_planar_hp<D, 8, 2, NW>(q, plan, mx, my, w, uh, c, b, inverse, stream);
    if (u_env == 4) return launch_planar_hp<D, 8, 4, NW>(q, plan, mx, my, w, uh, c, b, inverse, stream);
  }
  switch (L) {
    case 8: return launch_planar_hp<D, 8, 8, NW>(q, plan, mx, my, w, uh, c, b, inverse, stream);
    case 4: return launch_planar_hp<D, 4, 4, NW>(q, plan, mx, my, w, uh, c, b, inverse, stream);
    case 2: return launch_planar_hp<D, 2, 2, NW>(q, plan, mx, my, w, uh, c, b, inverse, stream);
    case 1: return launch_planar_hp<D, 1, 1, NW>(q, plan, mx, my, w, uh, c, b, inverse, stream);
    default: return B2B_EINVAL;
  }
}

}  // namespace b2b

// One launch of `L` (1, 2, 4 or 8) planar layers with derived parameters (w, û, c = wᵀû, b) in HOST memory.
int b2b_launch_planar_hostparams(const B2BChainParams& p, int L, const float* w, const float* uh, const float* c,
                                 const float* b, int inverse, cudaStream_t stream) {
  using namespace b2b;
  B2BChainParams q = p;
  q.L = 0;
  V1Plan plan;
  // warps per CTA of the host-parameter kernels (registers: 185 at D = 128 -> 10 warps fit)
  static const int nw128 = getenv("B2B_HP_NW") ? atoi(getenv("B2B_HP_NW")) : 8;
  const int rc = plan_v1(q, plan, q.D == 128 ? nw128 : 0);
  if (rc != 0) return rc;
  if (q.D > 128) return B2B_EUNSUPPORTED;
  CUtensorMap mx, my;
  if (!make_maps(q, plan.cols, &mx, &my, &plan.extra.tma3d)) return B2B_EUNSUPPORTED;
  if (q.D == 128 && nw128 == 10) return launch_planar_hp_L<128, 10>(L, q, plan, mx, my, w, uh, c, b, inverse, stream);
  if (q.D == 128 && nw128 == 12) return launch_planar_hp_L<128, 12>(L, q, plan, mx, my, w, uh, c, b, inverse, stream);
  if (q.D == 128) return launch_planar_hp_L<128, 8>(L, q, plan, mx, my, w, uh, c, b, inverse, stream);
  if (q.D == 64) return launch_planar_hp_L<64, 12>(L, q, plan, mx, my, w, uh, c, b, inverse, stream);
  return launch_planar_hp_L<32, 16>(L, q, plan, mx, my, w, uh, c, b, inverse, stream);
}
