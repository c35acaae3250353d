// Heisenberg-picture engine kernel for the HS loss on layered templates (the Synthesize.static()
// hot path: cost_HST of matrix_utils.py:35-42 over build_unitary of main.py:106-146).
//
// The adjoint sweep of engine_impl.cuh propagates two N x N complex states (phi, lambda) backwards:
// 22 FMA per amplitude and fused gate plus a cross-lane reduction for every gradient entry.  Here
// the backward sweep runs in the Heisenberg picture instead (tools/heisenberg_model.py is the numpy
// statement of the same math, checked against the oracle by tests/test_heisenberg_model.py):
//
//   forward   Y = U V^dag, started from V^dag (staged in shared memory by TMA) instead of the identity:
//             t = Tr(V^dag U) = Tr(Y), loss = 1 - |t|^2/N^2.  All n row bits of a column live in
//             registers (CPT = 2 columns packed in float2 -> FFMA2), so the forward sweep has no
//             shuffles at all: lane = column group.
//   pivot     dL/dtheta_k = Tr(H_k sigma) with H_k = Herm(s Z_k), s = i conj(t)/N^2,
//             Z_k = G_k..G_1 V^dag G_M..G_{k+1}, Z_M = Y.  H is Hermitian, so its Pauli coefficients
//             h[x, z] = Re(i^{|x&z|} s W[x,z]),  W[x,z] = sum_r (-1)^{|z&r|} Y[r, r^x]
//             are N^2 REAL numbers: one xor-shuffle all-to-all gathers the x-diagonals (lane = x >> PB)
//             and a register-local Walsh-Hadamard transform over r produces W.
//   backward  H_{k-1} = G_k^dag H_k G_k acts on h by real linear maps.  Every fused gate is used as
//             Rz(phi_out + pi/2) Rx(theta) Rz(phi_in - pi/2): Rx mixes (h_Y, h_Z) (2 FFMA2 + 1 SHFL per packed pair
//             when the qubit's x bit is a lane bit, the exchanged value is the raw register), Rz mixes
//             (h_X, h_Y) inside a lane (4 FFMA2); the Rz factors of consecutive gates on a qubit, the Rz halves of
//             CP(a) ~ Rz Rz exp(i a/4 ZZ) and the pending phases commute and are undone as ONE rotation per gate;
//             the ZZ part is a pair rotation by a/2 in lifting form (three shears).
//             Every gradient entry is ONE coefficient of h (X: h[b,0], Y: h[b,b], Z: h[0,b]; CP:
//             -(h[0,0]-h[0,b1]-h[0,b2]+h[0,b1|b2])/2): no reductions.
//
// Shared memory per sample: 8 words per fused gate + 3 per entangler + a staging area for the coefficients of
// ONE layer of the backward sweep (produced just in time from the forward data), see HEIS_SU2_WORDS below.
// Per sample and eval for C3 (n = 4, K = 40): ~5.6 k warp instructions (x 1/4 warp) instead of ~22 k.
#pragma once
#include <atomic>
#include <cmath>

#include "engine_impl.cuh"

namespace cpf {

// Shared-memory slot of a fused one-qubit gate, 8 words: ty, sy (lifting form of Ry: ty = -sy / (1 + cy)) | u_out |
// u_in (surface) / A (lower-qubit gate) / A B e^{ia} (higher-qubit gate) | B (lower-qubit gate) / r * penalty
// slope of the block's entangler (higher-qubit gate, word 6).  The backward sweep overwrites words 0, 1, 4 with the
// gradient sums (S_X, S_Y, S_Z); u_out stays (the parameter phase needs it to bring the sums back to the gate's
// output frame).  The coefficients of the backward sweep are NOT kept per gate: they are produced one layer at a
// time in a staging area (HEIS_STAGE_WORDS per gate of a layer), which keeps the per-sample footprint small enough
// for 64 resident samples (16 warps) per SM on C3.
constexpr int HEIS_SU2_WORDS = 8;
// Staged coefficients of a gate of the backward sweep (Z-X-Z form, see backward()), 8 words: two rows
// (cos theta, -+sin theta, cos zeta | 1, sin zeta | 0): row 0 for lanes that hold (I, Z) of the gate's qubit, row 1
// for lanes that hold (X, Y).  A lane picks its row by address: no selects.
constexpr int HEIS_STAGE_WORDS = 8;
constexpr int HEIS_CP_WORDS = 3;     // cos(a/2) >= 0, sin(a/2), -tan(a/4); word 0 <- dL/da after the backward sweep

inline int heis_coef_stride(int n_su2, int n_cp, int n_stage) {
  int w = (HEIS_SU2_WORDS * n_su2 + HEIS_CP_WORDS * n_cp + HEIS_STAGE_WORDS * n_stage + 3) & ~3;
  if (w == 0) w = 4;
  if (((w / 4) & 1) == 0) w += 4;   // odd number of 16-byte groups: samples of a warp hit distinct banks
  return w;
}
template <typename R> inline int heis_target_words(int n, int cpt) {
  const int N = 1 << n;
  return (N / cpt) * (N + 1) * 2 * cpt;
}

// Gate metadata as the kernel reads it, staged in shared memory by the prologue (the global-memory records of
// program.hpp sat on the critical path of every gate update: a dependent L2 round trip before the load of the
// packed state).  Constant angles (pidx < 0) stay in the global records.
struct __align__(8) HSu2 { int16_t pidx[3]; uint16_t axes; };                   // axes: 4 bits per rotation, 15 = unused
struct __align__(8) HCp { int16_t pidx; uint16_t flags; int16_t prev_lo, prev_hi; };   // flags: 1 penalised, 2 CZ, bits [4,7) lower qubit, [8,11) higher qubit
__host__ __device__ inline int heis_meta_bytes(int n_su2, int n_cp) { return (8 * (n_su2 + n_cp) + 15) & ~15; }

template <typename R, int NQ, int CPT>
struct HCfg {
  static constexpr int N = 1 << NQ;
  static constexpr int PB = CPT == 2 ? 1 : 0;   // x bits of a thread held in registers
  static constexpr int XR = 1 << PB;
  static constexpr int TPS = N / CPT;           // threads per sample
  // forward state registers: 2 * N * CPT words (x2 for double) -> register cap 128 (512 threads) or 255 (256)
#ifndef CPF_F64_MAXT
#define CPF_F64_MAXT 512
#endif
  static constexpr int MAXT = 2 * N * CPT * (int)(sizeof(R) / 4) <= 64 ? (sizeof(R) == 8 && NQ >= 4 ? CPF_F64_MAXT : 512) : 256;
  static_assert(TPS >= 1 && TPS <= 32, "a sample must fit in one warp");
};

// A CTA barrier at the start of each sweep keeps the warps of a CTA on the same instruction-cache lines (a barrier
// per layer, the first design, cost 7 %; none at all 20 %).  fwd / bwd: barrier before the forward / backward sweep.
struct LayerBar {
  bool fwd, bwd;
};

// Store to shared memory under a predicate, without a branch (the compiler turns `if (lane == k) s[i] = v`
// into BSSY / BRA / BSYNC sequences inside the sweeps).
__device__ __forceinline__ void sts_if(bool on, float* q, float v) {
  asm volatile("{ .reg .pred p; setp.ne.s32 p, %0, 0; @p st.shared.f32 [%1], %2; }" ::"r"((int)on),
               "r"((unsigned)__cvta_generic_to_shared(q)), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_if(bool on, double* q, double v) {
  asm volatile("{ .reg .pred p; setp.ne.s32 p, %0, 0; @p st.shared.f64 [%1], %2; }" ::"r"((int)on),
               "r"((unsigned)__cvta_generic_to_shared(q)), "d"(v) : "memory");
}

template <typename V> struct AddV;
template <> struct AddV<float> { static __device__ __forceinline__ float add(float a, float b) { return a + b; }
                                 static __device__ __forceinline__ float sub(float a, float b) { return a - b; } };
template <> struct AddV<double> { static __device__ __forceinline__ double add(double a, double b) { return a + b; }
                                  static __device__ __forceinline__ double sub(double a, double b) { return a - b; } };
template <> struct AddV<float2> {
  static __device__ __forceinline__ float2 add(float2 a, float2 b) { return __fadd2_rn(a, b); }
  static __device__ __forceinline__ float2 sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
};

// Layered template (same slot order as LayerSweep, program.cpp: detect_layered): surface SU2 of qubit q
// in slot q; block k = phase gate k on the qubit pair of position k % NBL, SU2 of the pair's lower
// qubit in slot NQ + 2k, of its higher qubit in slot NQ + 2k + 1.
template <typename R, int NQ, int CPT, int NBL, unsigned long long LOQ, unsigned long long HIQ>
struct HeisSweep {
  using CO = Cols<R, NQ, CPT>;
  using T = VT<R, CPT>;
  using V = typename T::V;
  static constexpr int N = 1 << NQ, PB = CPT == 2 ? 1 : 0, XR = 1 << PB, TPS = N / CPT;
  static constexpr int SW = HEIS_SU2_WORDS, CW = HEIS_CP_WORDS, STW = HEIS_STAGE_WORDS;
  static constexpr int NSTAGE = 2 * NBL > NQ ? 2 * NBL : NQ;   // gates staged at a time (one layer / the surface)
  static __host__ __device__ constexpr int lo_q(int j) { return (int)((LOQ >> (4 * j)) & 15); }
  static __host__ __device__ constexpr int hi_q(int j) { return (int)((HIQ >> (4 * j)) & 15); }

  // ------------------------------- forward: row operations on Y -------------------------------
  // Every fused gate is used in ZYZ form G ~ diag(1, u_out) Ry diag(1, u_in) (global phase dropped: the HS
  // loss and the Hermitian part that seeds the backward sweep do not see it).  u_in merges with the block's
  // CP phase and with the u_out still pending on the same qubits into one diagonal (1, B, A, A B e^{ia}),
  // prepared by the parameter phase: a block costs 3 + 4 + 4 = 11 FMA per amplitude instead of 1 + 8 + 8.
  // Slot words during the forward sweep: 0 ty = -sy / (1 + cy); 1 sy; [2,4) u_out; lower-qubit slot [4,8) A, B;
  // higher-qubit slot [4,6) A B e^{ia}; surface slots [4,6) u_in.
  template <int BP>
  static __device__ __forceinline__ void ry_fwd(V (&yr)[N], V (&yi)[N], const R* cf) {
    CO::template ry_lift<BP>(yr, yi, cf[0], cf[1]);
  }
  template <int Q>
  static __device__ __forceinline__ void surface_fwd(V (&yr)[N], V (&yi)[N], const R* coef) {
    if constexpr (Q < NQ) {
      constexpr int BP = NQ - 1 - Q;
      CO::template phase_mask<(1 << BP), 0>(yr, yi, coef[SW * Q + 4], coef[SW * Q + 5]);
      ry_fwd<BP>(yr, yi, coef + SW * Q);
      surface_fwd<Q + 1>(yr, yi, coef);
    }
  }
  template <int Q>
  static __device__ __forceinline__ void tail_fwd(const KParams<R>& p, V (&yr)[N], V (&yi)[N], const R* coef) {
    if constexpr (Q < NQ) {
      const R* cf = coef + SW * p.last_slot[Q];
      CO::template phase_mask<(1 << (NQ - 1 - Q)), 0>(yr, yi, cf[2], cf[3]);
      tail_fwd<Q + 1>(p, yr, yi, coef);
    }
  }
  // one block on amplitude bits PA (lower qubit) / PC (higher qubit); cl: slot of the lower-qubit gate
  template <int PA, int PC>
  static __device__ __forceinline__ void block_fwd(const R* cl, V (&yr)[N], V (&yi)[N]) {
    const R* ch = cl + SW;
    R ar, ai, br, bi;
    Vec4Load<R>::ld(cl + 4, ar, ai, br, bi);
    CO::template phase_mask<(1 << PA), (1 << PC)>(yr, yi, ar, ai);
    CO::template phase_mask<(1 << PC), (1 << PA)>(yr, yi, br, bi);
    CO::template phase_mask<(1 << PA) | (1 << PC), 0>(yr, yi, ch[4], ch[5]);
    ry_fwd<PA>(yr, yi, cl);
    ry_fwd<PC>(yr, yi, ch);
  }
  template <int J>
  static __device__ __forceinline__ void blocks_fwd(int k0, int K, const R* cs, V (&yr)[N], V (&yi)[N]) {
    if constexpr (J < NBL) {
      if (k0 + J >= K) return;
      block_fwd<NQ - 1 - lo_q(J), NQ - 1 - hi_q(J)>(cs + 2 * SW * J, yr, yi);
      blocks_fwd<J + 1>(k0, K, cs, yr, yi);
    }
  }
  static __device__ __forceinline__ void forward(const KParams<R>& p, const LayerBar lb, const HCp*, const R* coef,
                                                 V (&yr)[N], V (&yi)[N]) {
    surface_fwd<0>(yr, yi, coef);
    const int K = p.n_cp;
    const R* cs = coef + SW * NQ;
#pragma unroll 1
    if (lb.fwd) __syncthreads();
    for (int k0 = 0; k0 < K; k0 += NBL) {
      blocks_fwd<0>(k0, K, cs, yr, yi);
      cs += 2 * SW * NBL;
    }
    tail_fwd<0>(p, yr, yi, coef);
  }

  // ------------------ pivot: gather the x-diagonals, Walsh-Hadamard transform over r ------------------
  // Before: lane l holds Y[r, l*CPT + k] (k = packed component).  Round d: every lane exchanges its rows
  // r >> PB == d with lane l ^ d, in place.  After: lane m holds, in slot r, Y[r, ((m ^ (r>>PB)) << PB) | k],
  // i.e. the element of diagonal x = (m << PB) | (k ^ (r & PB-bit)) at row r.
  static __device__ __forceinline__ void gather_wht(V (&yr)[N], V (&yi)[N]) {
#pragma unroll
    for (int d = 1; d < TPS; ++d) {
#pragma unroll
      for (int r0 = 0; r0 < CPT; ++r0) {
        const int r = d * CPT + r0;
        yr[r] = ShflV<V>::x(yr[r], d);
        yi[r] = ShflV<V>::x(yi[r], d);
      }
    }
    if constexpr (CPT == 2) {
      // butterfly over row bit 0: slot r0 = 0 holds (x0 = 0, x0 = 1), slot r0 = 1 holds (x0 = 1, x0 = 0)
#pragma unroll
      for (int r = 0; r < N; r += 2) {
        const V ur = yr[r], wr = yr[r + 1], ui = yi[r], wi = yi[r + 1];
        yr[r] = T::make(ur.x + wr.y, ur.y + wr.x); yr[r + 1] = T::make(ur.x - wr.y, ur.y - wr.x);
        yi[r] = T::make(ui.x + wi.y, ui.y + wi.x); yi[r + 1] = T::make(ui.x - wi.y, ui.y - wi.x);
      }
    }
#pragma unroll
    for (int j = PB; j < NQ; ++j) {
#pragma unroll
      for (int r = 0; r < N; ++r) {
        if (r & (1 << j)) continue;
        const V ar = yr[r], br = yr[r | (1 << j)], ai = yi[r], bi = yi[r | (1 << j)];
        yr[r] = AddV<V>::add(ar, br); yr[r | (1 << j)] = AddV<V>::sub(ar, br);
        yi[r] = AddV<V>::add(ai, bi); yi[r | (1 << j)] = AddV<V>::sub(ai, bi);
      }
    }
  }

  // ------------------------------- backward: real maps on h -------------------------------
  // h[xr][z] is held packed over xr: hv[z] = (h[0][z], h[1][z]) for CPT = 2 (scalar for CPT = 1), so every map
  // whose coefficients do not depend on xr issues as FFMA2; only gates on amplitude bit 0 (x bit = xr) are scalar.
  //
  // Z-X-Z form with merged Rz factors (tools/heisenberg_model.py: grad_hs_zxz).  G ~ Rz(phi_out) Ry(theta) Rz(phi_in)
  // = Rz(phi_out + pi/2) Rx(theta) Rz(phi_in - pi/2).  Rx mixes (Y, Z): for a lane-bit qubit both sit in register
  // slot z | BM of the two paired lanes, so the exchange is the raw register (no send computation): 2 FMA + SHFL per
  // packed pair; Rz mixes (X, Y), which one lane holds: 4 FMA.  Everything between the Rx of consecutive gates on
  // a qubit is Z-type and commutes, so a gate's outgoing Rz is undone together with the incoming Rz of the NEXT
  // gate on that qubit: per gate, undo Rx(theta), then Rz(zeta) with e^{i zeta} = (A or B of the forward sweep's
  // merged diagonal) e^{i a/2}.  6 instead of 8 packed instructions per pair, and nothing to select.
  // cf: the gate's staged rows; words 0, 1, 4 of the gate's slot sl receive (S_X, S_Y, S_Z) = entries of h in the
  // frame where the gate's own Rz(phi_out + pi/2) is already undone (the parameter phase rotates them back).
  template <int B>
  static __device__ __forceinline__ void su2_bwd(V (&hv)[N], const R* cf, R* sl, int m) {
    constexpr int BM = 1 << B;
    if constexpr (B < PB) {
      const bool own = m == 0;
      sts_if(own, sl, T::get(hv[0], 1)); sts_if(own, sl + 1, T::get(hv[BM], 1)); sts_if(own, sl + 4, T::get(hv[BM], 0));
      // x bit = xr: I = h[0][z], Z = h[0][z|1], X = h[1][z], Y = h[1][z|1]
      R ct, st, cz, sz;
      Vec4Load<R>::ld(cf + 4, ct, st, cz, sz);
#pragma unroll
      for (int z = 0; z < N; ++z) {
        if (z & BM) continue;
        const R X = T::get(hv[z], 1), Y = T::get(hv[z | BM], 1), Z = T::get(hv[z | BM], 0);
        const R Y1 = ct * Y + st * Z;
        hv[z] = T::make(T::get(hv[z], 0), cz * X + sz * Y1);
        hv[z | BM] = T::make(ct * Z - st * Y, cz * Y1 - sz * X);
      }
    } else {
      // x bit is lane bit J: lanes with the bit clear hold (I, Z), lanes with it set hold (X, Y)
      constexpr int J = B - PB;
      const bool mb = ((m >> J) & 1) != 0;
      const bool own = m == (1 << J);
      sts_if(own, sl, T::get(hv[0], 0)); sts_if(own, sl + 1, T::get(hv[BM], 0));
      sts_if(m == 0, sl + 4, T::get(hv[BM], 0));
      R ct, st, cz, sz;
      Vec4Load<R>::ld(cf + (mb ? 4 : 0), ct, st, cz, sz);
      const V kct = T::bc(ct), kst = T::bc(st), kcz = T::bc(cz), ksz = T::bc(sz), knz = T::bc(-sz);
#pragma unroll
      for (int z = 0; z < N; ++z) {
        if (z & BM) continue;
        const V recv = ShflV<V>::x(hv[z | BM], 1 << J);
        const V e1 = T::fma(kst, recv, T::mul(kct, hv[z | BM]));   // (X, Y) lanes: Y' = ct Y + st Z; (I, Z): Z' = ct Z - st Y
        const V e0 = hv[z];
        hv[z] = T::fma(ksz, e1, T::mul(kcz, e0));                  // X' = cz X + sz Y'      (identity on (I, Z) lanes)
        hv[z | BM] = T::fma(knz, e0, T::mul(kcz, e1));             // Y'' = cz Y' - sz X
      }
    }
  }
  // Rz alone (the outgoing Rz of the last gate on a qubit, undone before the sweep starts): (X, Y) -> (c X + s Y, c Y - s X)
  template <int B>
  static __device__ __forceinline__ void rz_bwd(V (&hv)[N], R c, R s, int m) {
    constexpr int BM = 1 << B;
    if constexpr (B < PB) {
#pragma unroll
      for (int z = 0; z < N; ++z) {
        if (z & BM) continue;
        const R X = T::get(hv[z], 1), Y = T::get(hv[z | BM], 1);
        hv[z] = T::make(T::get(hv[z], 0), c * X + s * Y);
        hv[z | BM] = T::make(T::get(hv[z | BM], 0), c * Y - s * X);
      }
    } else {
      const bool mb = ((m >> (B - PB)) & 1) != 0;
      const V kc = T::bc(mb ? c : R(1)), ks = T::bc(mb ? s : R(0)), kn = T::bc(mb ? -s : R(0));
#pragma unroll
      for (int z = 0; z < N; ++z) {
        if (z & BM) continue;
        const V e0 = hv[z], e1 = hv[z | BM];
        hv[z] = T::fma(ks, e1, T::mul(kc, e0));
        hv[z | BM] = T::fma(kn, e0, T::mul(kc, e1));
      }
    }
  }

  // x bit of amplitude bit B (B >= PB) for this lane
  template <int B>
  static __device__ __forceinline__ bool xlane(int m) { return ((m >> (B - PB)) & 1) != 0; }

  // Entangler of a block on amplitude bits (B1, B2).  CP(a) ~ Rz_1(a/2) Rz_2(a/2) exp(i (a/4) Z1 Z2): the two Rz
  // join the merged Rz rotation of the block's fused gates (stage_layer: zeta), what remains rotates,
  // for coefficients whose x bits on the two qubits differ, the pairs (e00, e11) and (e01, e10) of every
  // (z1, z2) quad by a/2.  cf: cos(a/2), sin(a/2); word 0 receives dL/da = -(h_II - h_ZI - h_IZ + h_ZZ)/2
  // (Z-type coefficients: they do not see the Rz shuffle).
  // Both pair rotations in lifting form (three shears each: 6 instead of 8 FMA-pipe instructions per quad): the
  // parameter phase keeps cos(a/2) >= 0 (CP(a) = CP(a - 2 pi)), so t = -tan(a/4) = -s / (1 + c) lies in [-1, 1].
  template <typename E>
  static __device__ __forceinline__ void zz_quad(E& e00, E& e01, E& e10, E& e11, E tl, E sl, E t2, E s2) {
    using O = VT<R, sizeof(E) == sizeof(R) ? 1 : 2>;
    E a = O::fma(tl, e11, e00), b = O::fma(t2, e10, e01);
    e11 = O::fma(sl, a, e11); e10 = O::fma(s2, b, e10);
    e00 = O::fma(tl, e11, a); e01 = O::fma(t2, e10, b);
  }
  template <int B1, int B2>
  static __device__ __forceinline__ void phase_bwd(V (&hv)[N], R* cf, int m) {
    constexpr int M1 = 1 << B1, M2 = 1 << B2;
    const R s = cf[1], t = cf[2];
    const R g11 = R(-0.5) * ((T::get(hv[0], 0) - T::get(hv[M1], 0)) - (T::get(hv[M2], 0) - T::get(hv[M1 | M2], 0)));
    __syncwarp();
    sts_if(m == 0, cf, g11);
    if constexpr (B1 >= PB && B2 >= PB) {
      // both x bits are lane bits: one case per lane, packed arithmetic over xr
      const bool x1 = xlane<B1>(m), x2 = xlane<B2>(m), on = x1 != x2;
      const V tl = T::bc(on ? t : R(0)), sl = T::bc(on ? s : R(0));
      const V t2 = T::bc(on ? (x1 ? t : -t) : R(0)), s2 = T::bc(on ? (x1 ? s : -s) : R(0));
#pragma unroll
      for (int z = 0; z < N; ++z) {
        if (z & (M1 | M2)) continue;
        zz_quad<V>(hv[z], hv[z | M2], hv[z | M1], hv[z | M1 | M2], tl, sl, t2, s2);
      }
    } else {
      // one of the bits is amplitude bit 0, whose x bit is the packed component xr: scalar per component
      constexpr int BL = B1 >= PB ? B1 : B2;          // the lane bit
      const bool xl = xlane<BL>(m);
#pragma unroll
      for (int xr = 0; xr < XR; ++xr) {
        const bool x1 = B1 >= PB ? xl : xr != 0, x2 = B2 >= PB ? xl : xr != 0, on = x1 != x2;
        const R tl = on ? t : R(0), sl = on ? s : R(0), t2 = on ? (x1 ? t : -t) : R(0), s2 = on ? (x1 ? s : -s) : R(0);
#pragma unroll
        for (int z = 0; z < N; ++z) {
          if (z & (M1 | M2)) continue;
          R e00 = T::get(hv[z], xr), e01 = T::get(hv[z | M2], xr), e10 = T::get(hv[z | M1], xr),
            e11 = T::get(hv[z | M1 | M2], xr);
          zz_quad<R>(e00, e01, e10, e11, tl, sl, t2, s2);
          setc(hv[z], xr, e00); setc(hv[z | M2], xr, e01); setc(hv[z | M1], xr, e10); setc(hv[z | M1 | M2], xr, e11);
        }
      }
    }
  }
  static __device__ __forceinline__ void setc(V& v, int k, R x) {
    if constexpr (CPT == 2) { if (k) v.y = x; else v.x = x; }
    else v = x;
  }

  // one block of the backward sweep; st: staged rows of the block's two gates (lower, higher), cl: slot of the
  // lower-qubit gate, cph: the entangler's words
  template <int PA, int PC>
  static __device__ __forceinline__ void block_bwd(V (&h)[N], const R* st, R* cl, R* cph, int m) {
    su2_bwd<PC>(h, st + STW, cl + SW, m);
    su2_bwd<PA>(h, st, cl, m);
    phase_bwd<PA, PC>(h, cph, m);
  }
  template <int J>
  static __device__ __forceinline__ void blocks_bwd(int k0, int K, R* cs, const R* stage, R* cph, int m, V (&h)[N]) {
    if constexpr (J >= 0) {
      if (k0 + J < K)
        block_bwd<NQ - 1 - lo_q(J), NQ - 1 - hi_q(J)>(h, stage + STW * (2 * J), cs + 2 * SW * J, cph + CW * J, m);
      blocks_bwd<J - 1>(k0, K, cs, stage, cph, m, h);
    }
  }
  template <int Q>
  static __device__ __forceinline__ void surface_bwd(R* coef, const R* stage, int m, V (&h)[N]) {
    if constexpr (Q >= 0) {
      su2_bwd<NQ - 1 - Q>(h, stage + STW * Q, coef + SW * Q, m);
      surface_bwd<Q - 1>(coef, stage, m, h);
    }
  }
  // Staged rows of one gate: Rx angle from the forward data (cos phi = 1 + ty sy, theta = 2 phi), Rz angle zeta.
  static __device__ __forceinline__ void stage_gate(R* st, R ty, R sy, R cz, R sz) {
    const R cy = R(1) + ty * sy;   // 1 - tan(phi/2) sin phi = cos phi
    const R ct = cy * cy - sy * sy, sth = R(2) * cy * sy;
    Vec4Load<R>::st(st, ct, -sth, R(1), R(0));
    Vec4Load<R>::st(st + 4, ct, sth, cz, sz);
  }
  // Staging of one layer (blocks k0 .. k0 + NBL - 1): the sample's lanes split the layer's 2 NBL fused gates.
  // e^{i zeta} = d e^{i a/2}, d = A (lower-qubit gate) / B (higher-qubit gate) of the forward sweep's merged diagonal.
  static __device__ __forceinline__ void stage_layer(int k0, int K, const R* coef, const R* cph0, R* stage, int m,
                                                     int group = NBL) {
    const int nb = K - k0 < group ? K - k0 : group;
#pragma unroll 1
    for (int j = m; j < 2 * nb; j += TPS) {
      const int k = k0 + (j >> 1), hi = j & 1;
      const R* cl = coef + SW * (NQ + 2 * k);
      const R* cf = cl + SW * hi;
      const R* cc = cph0 + CW * k;
      const R dr = cl[4 + 2 * hi], di = cl[5 + 2 * hi];
      const R ch = cc[0], sh = cc[1];
      stage_gate(stage + STW * j, cf[0], cf[1], dr * ch - di * sh, dr * sh + di * ch);
    }
  }
  // surface gates: zeta = phi_in - pi/2
  static __device__ __forceinline__ void stage_surface(const R* coef, R* stage, int m) {
#pragma unroll 1
    for (int q = m; q < NQ; q += TPS) {
      const R* cf = coef + SW * q;
      stage_gate(stage + STW * q, cf[0], cf[1], cf[5], -cf[4]);
    }
  }
  template <int Q>
  static __device__ __forceinline__ void tail_bwd(const KParams<R>& p, const R* coef, int m, V (&h)[N]) {
    if constexpr (Q < NQ) {
      const R* cf = coef + SW * p.last_slot[Q];
      rz_bwd<NQ - 1 - Q>(h, -cf[3], cf[2], m);     // Rz(phi_out + pi/2): e^{i w} = i u_out
      tail_bwd<Q + 1>(p, coef, m, h);
    }
  }
  static __device__ __forceinline__ void backward(const KParams<R>& p, const LayerBar lb, const HCp*, R* coef, R* stage,
                                                  R* cph0, int m, V (&h)[N]) {
    const int K = p.n_cp;
    if (lb.bwd) __syncthreads();
    tail_bwd<0>(p, coef, m, h);
#pragma unroll 1
    for (int k0 = K > 0 ? ((K - 1) / NBL) * NBL : -1; k0 >= 0; k0 -= NBL) {
      stage_layer(k0, K, coef, cph0, stage, m);
      __syncwarp();
      blocks_bwd<NBL - 1>(k0, K, coef + SW * NQ + 2 * SW * k0, stage, cph0 + CW * k0, m, h);
      __syncwarp();
    }
    stage_surface(coef, stage, m);
    __syncwarp();
    surface_bwd<NQ - 1>(coef, stage, m, h);
  }
};

// The same sweeps for ANY block-structured template (topology.py:7-20 allows every `layer`; the paper's kite and
// square Toffoli-4 layers, twisted placements, non-periodic free blocks): the qubit pair of a block is read from the
// shared-memory gate metadata and dispatched by one uniform switch per block into the compile-time-pair code above
// (n (n - 1) / 2 variants).  All warps of a CTA take the same case, so the instruction-cache footprint per layer is
// what the compile-time-layer kernels have; the cost is the branch and the lost scheduling across block boundaries.
#define CPF_PAIR_LIST(X) X(0, 1) X(0, 2) X(0, 3) X(0, 4) X(1, 2) X(1, 3) X(1, 4) X(2, 3) X(2, 4) X(3, 4)
template <typename R, int NQ, int CPT>
struct HeisSweepAny : HeisSweep<R, NQ, CPT, 1, 0x0ull, 0x1ull> {
  using Base = HeisSweep<R, NQ, CPT, 1, 0x0ull, 0x1ull>;
  using V = typename Base::V;
  static constexpr int N = Base::N, SW = Base::SW, CW = Base::CW, STW = Base::STW, TPS = Base::TPS;
  static constexpr int GROUP = 4;                                   // blocks staged at a time (backward sweep)
  static constexpr int NSTAGE = 2 * GROUP > NQ ? 2 * GROUP : NQ;
  static __device__ __forceinline__ int pair_code(const HCp& md) { return ((md.flags >> 4) & 7) * 8 + ((md.flags >> 8) & 7); }

  static __device__ __forceinline__ void forward(const KParams<R>& p, const LayerBar lb, const HCp* s_cp, const R* coef,
                                                 V (&yr)[N], V (&yi)[N]) {
    Base::template surface_fwd<0>(yr, yi, coef);
    const int K = p.n_cp;
    const R* cs = coef + SW * NQ;
    if (lb.fwd) __syncthreads();
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      switch (pair_code(s_cp[k])) {
#define CPF_X(LO, HI)                                                                                        \
        case LO * 8 + HI:                                                                                    \
          if constexpr (HI < NQ) Base::template block_fwd<NQ - 1 - LO, NQ - 1 - HI>(cs, yr, yi);             \
          break;
        CPF_PAIR_LIST(CPF_X)
#undef CPF_X
        default: break;
      }
      cs += 2 * SW;
    }
    Base::template tail_fwd<0>(p, yr, yi, coef);
  }

  static __device__ __forceinline__ void backward(const KParams<R>& p, const LayerBar lb, const HCp* s_cp, R* coef,
                                                  R* stage, R* cph0, int m, V (&h)[N]) {
    const int K = p.n_cp;
    if (lb.bwd) __syncthreads();
    Base::template tail_bwd<0>(p, coef, m, h);
#pragma unroll 1
    for (int k0 = K > 0 ? ((K - 1) / GROUP) * GROUP : -1; k0 >= 0; k0 -= GROUP) {
      Base::stage_layer(k0, K, coef, cph0, stage, m, GROUP);
      __syncwarp();
      const int k1 = k0 + GROUP < K ? k0 + GROUP : K;
#pragma unroll 1
      for (int k = k1 - 1; k >= k0; --k) {
        const R* st = stage + STW * 2 * (k - k0);
        R* cl = coef + SW * (NQ + 2 * k);
        R* cph = cph0 + CW * k;
        switch (pair_code(s_cp[k])) {
#define CPF_X(LO, HI)                                                                                        \
          case LO * 8 + HI:                                                                                  \
            if constexpr (HI < NQ) Base::template block_bwd<NQ - 1 - LO, NQ - 1 - HI>(h, st, cl, cph, m);    \
            break;
          CPF_PAIR_LIST(CPF_X)
#undef CPF_X
          default: break;
        }
      }
      __syncwarp();
    }
    Base::stage_surface(coef, stage, m);
    __syncwarp();
    Base::template surface_bwd<NQ - 1>(coef, stage, m, h);
  }
};

// ------------------------------------------------------------------------------------------
// update / coefficient phase helpers (per gate, executed by the gate's owner thread)
// ------------------------------------------------------------------------------------------
// Per-sample arrays are addressed as array[off + index] with a 32-bit element offset off = b * P (the
// host checks B * P < 2^32): one IMAD.WIDE per access instead of a 64-bit multiply chain.
template <typename R>
struct UpdCtx {
  int phase;            // PH_COEF / PH_ADAM / PH_GRAD
  bool active;          // this thread's sample exists
  bool store_best;      // the step being finished improved on the best regloss: keep its parameters
  bool skip_coef;       // last pass: no new coefficients
  bool hist;            // parameter history is recorded for this step
  unsigned off;         // b * P
  size_t hist_off;      // (b * hist_len + gu + 1) * P
  R bc1, bc2, ibc1, ibc2;
};

// Optimiser state of one parameter, packed so that the update phase moves it with one 16-byte (float) or two
// 16-byte (double) accesses: theta, Adam moments, best-regloss parameter.  The kernel packs the caller's
// separate arrays (include/cpflow_b200.h: cpf_adam_buffers) into this scratch at launch and unpacks at the end.
template <typename R> struct __align__(4 * sizeof(R) > 16 ? 16 : 4 * sizeof(R)) Pk4 { R th, mu, nu, best; };
__device__ __forceinline__ Pk4<float> pk_load(const Pk4<float>* q) {
  const float4 t = *reinterpret_cast<const float4*>(q);
  return {t.x, t.y, t.z, t.w};
}
__device__ __forceinline__ void pk_store(Pk4<float>* q, const Pk4<float>& v) {
  *reinterpret_cast<float4*>(q) = make_float4(v.th, v.mu, v.nu, v.best);
}
__device__ __forceinline__ Pk4<double> pk_load(const Pk4<double>* q) {
  const double2 a = reinterpret_cast<const double2*>(q)[0], b = reinterpret_cast<const double2*>(q)[1];
  return {a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void pk_store(Pk4<double>* q, const Pk4<double>& v) {
  reinterpret_cast<double2*>(q)[0] = make_double2(v.th, v.mu);
  reinterpret_cast<double2*>(q)[1] = make_double2(v.nu, v.best);
}

// MUFU approximations (rsqrt: 2^-22.4, rcp: 1 ulp) refined by one Newton step: the unit phases of the ZYZ data are
// built from these, and a modulus error of 2e-7 per phase accumulates over the ~3 K diagonal factors of a sweep
// (measured: worst gradient error of 4096 samples on C3 1.46e-5 with the raw approximations; profiles/grad_accuracy_r2.txt).
__device__ __forceinline__ float rsqrt_fast(float a) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  const float h = 0.5f * a * r;
  return fmaf(r, fmaf(-h, r, 0.5f), r);          // r (1.5 - 0.5 a r^2)
}
__device__ __forceinline__ double rsqrt_fast(double a) { return rsqrt_r(a); }
__device__ __forceinline__ float rcp_fast(float a) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return fmaf(r, fmaf(-a, r, 1.0f), r);          // r (2 - a r)
}
__device__ __forceinline__ double rcp_fast(double a) { return 1.0 / a; }
static __device__ __noinline__ SinCos<float> sincos_slow_v(float x) { float s, c; sincosf(x, &s, &c); return {s, c}; }
// sin/cos for the parameter phase, inlined so the three evaluations of a fused gate interleave
__device__ __forceinline__ void sincos_core(float x, float& s, float& c) {
  const float j = rintf(x * 0.636619747f);
  float r = fmaf(j, -1.57079601e+00f, x);
  r = fmaf(j, -3.13916473e-07f, r);
  r = fmaf(j, -5.39030253e-15f, r);
  const int q = __float2int_rn(j);
  const float r2 = r * r;
  float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = fmaf(sp, r2, -1.6666654611e-1f);
  sp = fmaf(sp * r2, r, r);
  float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = fmaf(cp, r2, 4.166664568298827e-2f);
  cp = fmaf(cp * r2, r2, fmaf(r2, -0.5f, 1.0f));
  const float ss = (q & 1) ? cp : sp;
  const float cc = (q & 1) ? sp : cp;
  s = (q & 2) ? -ss : ss;
  c = ((q + 1) & 2) ? -cc : cc;
}
__device__ __forceinline__ void sincos_inl(float x, float& s, float& c) {
  sincos_core(x, s, c);
  if (fabsf(x) > 48000.f) { const SinCos<float> t = sincos_slow_v(x); s = t.s; c = t.c; }
}
__device__ __forceinline__ void sincos_inl(double x, double& s, double& c) { sincos_r(x, s, c); }
// the three half angles of a fused gate: one (never taken in practice) large-argument test for all of them
__device__ __forceinline__ void sincos3(bool on0, bool on1, bool on2, float x0, float x1, float x2, float& s0, float& c0,
                                        float& s1, float& c1, float& s2, float& c2) {
  if (on0) sincos_core(x0, s0, c0);
  if (on1) sincos_core(x1, s1, c1);
  if (on2) sincos_core(x2, s2, c2);
  if (fmaxf(fmaxf(on0 ? fabsf(x0) : 0.f, on1 ? fabsf(x1) : 0.f), on2 ? fabsf(x2) : 0.f) > 48000.f) {
    if (on0) sincos_inl(x0, s0, c0);
    if (on1) sincos_inl(x1, s1, c1);
    if (on2) sincos_inl(x2, s2, c2);
  }
}
__device__ __forceinline__ void sincos3(bool on0, bool on1, bool on2, double x0, double x1, double x2, double& s0,
                                        double& c0, double& s1, double& c1, double& s2, double& c2) {
  if (on0) sincos_r(x0, s0, c0);
  if (on1) sincos_r(x1, s1, c1);
  if (on2) sincos_r(x2, s2, c2);
}

// optax scale_by_adam + scale(-lr) (optimization.py:22-23).  double: exact IEEE sequence of the oracle.
// float: reciprocal bias corrections, approximate sqrt and division (<= 2 ulp each; the reference's XLA
// arithmetic is not bit-reproducible either, the Adam parity tests bound the drift).
__device__ __forceinline__ void adam_inl(const KParams<double>& p, const UpdCtx<double>& u, double g, double& th,
                                         double& mu, double& nu) {
  mu = add_rn(mul_rn(p.omb1, g), mul_rn(p.b1, mu));
  nu = add_rn(mul_rn(p.omb2, mul_rn(g, g)), mul_rn(p.b2, nu));
  const double mu_hat = mu / u.bc1, nu_hat = nu / u.bc2;
  th = add_rn(th, mul_rn(-p.lr, mu_hat / add_rn(sqrt(nu_hat), p.eps)));
}
__device__ __forceinline__ void adam_inl(const KParams<float>& p, const UpdCtx<float>& u, float g, float& th,
                                         float& mu, float& nu) {
  mu = add_rn(mul_rn(p.omb1, g), mul_rn(p.b1, mu));
  nu = add_rn(mul_rn(p.omb2, mul_rn(g, g)), mul_rn(p.b2, nu));
  const float mu_hat = mu * u.ibc1, nu_hat = nu * u.ibc2;
  // .ftz forms: without them ptxas wraps each MUFU in a denormal rescue (4 extra instructions); a denormal
  // nu_hat is far below eps^2 either way
  float rt, iv;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(nu_hat));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iv) : "f"(add_rn(rt, p.eps)));
  th = add_rn(th, mul_rn(-p.lr, mul_rn(mu_hat, iv)));
}

// one parameter: gradient sink (loss_grad mode) or best-parameter bookkeeping + Adam step on the packed state
// PLAIN: Adam pass of a launch without freeze mask and parameter history (the stage-1 runs of Synthesize.static()):
// the per-parameter tests on those pointers disappear from the gate loops.
template <typename R, bool PLAIN>
__device__ __forceinline__ void heis_apply(const KParams<R>& p, const UpdCtx<R>& u, Pk4<R>* pk, int pi, R g, Pk4<R>& v) {
  if (!PLAIN && u.phase == PH_GRAD) {
    if (u.active) p.grad_out[u.off + (unsigned)pi] = g;
    return;
  }
  // v.th is still the pre-update parameter of the step being finished (optimization.py:70-73)
  if (u.store_best) v.best = v.th;
  if (PLAIN || !(p.freeze && p.freeze[u.off + (unsigned)pi])) adam_inl(p, u, g, v.th, v.mu, v.nu);
  if (u.active) {
    pk_store(pk + pi, v);
    if (!PLAIN && u.hist) p.hist_params[u.hist_off + pi] = v.th;
  }
}

// (alpha, beta) <- R_a(c, s) * (alpha, beta) for a rotation about a compile-time-foldable axis: 8 FMA
template <typename R>
__device__ __forceinline__ void su2_lmul_axis(int a, R c, R s, R& ar, R& ai, R& br, R& bi) {
  if (a == 0) {        // Rx = [[c, -is], [-is, c]]
    const R nar = c * ar + s * bi, nai = c * ai - s * br, nbr = c * br + s * ai, nbi = c * bi - s * ar;
    ar = nar; ai = nai; br = nbr; bi = nbi;
  } else if (a == 1) { // Ry = [[c, -s], [s, c]]
    const R nar = c * ar - s * br, nai = c * ai - s * bi, nbr = c * br + s * ar, nbi = c * bi + s * ai;
    ar = nar; ai = nai; br = nbr; bi = nbi;
  } else if (a == 2) { // Rz = diag(c - is, c + is)
    const R nar = c * ar + s * ai, nai = c * ai - s * ar, nbr = c * br - s * bi, nbi = c * bi + s * br;
    ar = nar; ai = nai; br = nbr; bi = nbi;
  }
}

// Everything the update of one fused gate reads from global / shared memory.  The gate loops below are
// software pipelined: the loads of the next gate are issued before the current gate is processed, so the
// L2 round trips (theta, Adam moments, half-angle cos/sin of the fused rotations) overlap the arithmetic.
template <typename R>
struct GateIn {
  int pi0, pi1, pi2, axes;
  Pk4<R> v0, v1, v2;
  R sx, sy, sz, c2, s2, c3, s3;
};
template <typename R>
__device__ __forceinline__ GateIn<R> heis_gate_load(const KParams<R>& p, const UpdCtx<R>& u, const Pk4<R>* pk, bool valid,
                                                    const Su2Meta* md, const HSu2* ms, const R* cf, const R* ax) {
  GateIn<R> in;
  in.pi0 = in.pi1 = in.pi2 = -1; in.axes = 0xfff;
  in.v0 = in.v1 = in.v2 = Pk4<R>{R(0), R(0), R(0), R(0)};
  in.sx = in.sy = in.sz = in.c2 = in.s2 = in.c3 = in.s3 = R(0);
  if (!valid) return in;
  const HSu2 hm = *ms;
  in.pi0 = hm.pidx[0]; in.pi1 = hm.pidx[1]; in.pi2 = hm.pidx[2];
  in.axes = hm.axes;
  if (in.pi0 >= 0) in.v0 = pk_load(pk + in.pi0); else in.v0.th = R(md->cangle[0]);
  if (in.pi1 >= 0) in.v1 = pk_load(pk + in.pi1); else in.v1.th = R(md->cangle[1]);
  if (in.pi2 >= 0) in.v2 = pk_load(pk + in.pi2); else in.v2.th = R(md->cangle[2]);
  if (u.phase != PH_COEF) {
    // gradient sums of the backward sweep, back to the gate's output frame: the sweep reads them after the gate's
    // own Rz(phi_out + pi/2) is undone, e^{i w} = i u_out (u_out is still in words 2, 3 from the last update)
    const R px = cf[0], py = cf[1], wr = -cf[3], wi = cf[2];
    in.sx = wr * px - wi * py; in.sy = wi * px + wr * py; in.sz = cf[4];
    Vec4Load<R>::ld(ax, in.c2, in.s2, in.c3, in.s3);
  }
  return in;
}

// Fused one-qubit gate: finish the step (chain rule through the fusion, Adam), then the new ZYZ data of the gate.
// AX* >= 0: compile-time rotation axes (the selects fold away); AX0 == -2: axes from the gate metadata.
template <typename R, int AX0, int AX1, int AX2, bool PLAIN>
__device__ __forceinline__ void heis_su2_update(const KParams<R>& p, const UpdCtx<R>& u, Pk4<R>* pk, const Su2Meta* md,
                                                GateIn<R> in, R* cf, R* ax) {
  const int a0 = in.axes & 15, a1 = (in.axes >> 4) & 15, a2 = (in.axes >> 8) & 15;   // 15 = unused slot
  const int ax0 = AX0 == -2 ? (a0 == 15 ? -1 : a0) : AX0;
  const int ax1 = AX0 == -2 ? (a1 == 15 ? -1 : a1) : AX1;
  const int ax2 = AX0 == -2 ? (a2 == 15 ? -1 : a2) : AX2;
  if (u.phase != PH_COEF) {
    // chain rule through G = R_2 R_1 R_0: g_2 = S . e_2, g_1 = S . (R_2 e_1) = (R_2^T S) . e_1,
    // g_0 = (R_1^T R_2^T S) . e_0: rotate S backwards (no products with the zero entries of unit vectors)
    const R C2 = in.c2 * in.c2 - in.s2 * in.s2, S2 = R(2) * in.c2 * in.s2;
    const R C3 = in.c3 * in.c3 - in.s3 * in.s3, S3 = R(2) * in.c3 * in.s3;
    R sx = in.sx, sy = in.sy, sz = in.sz;
    const R g2 = sel3(ax2, sx, sy, sz);
    rot_axis(ax2, C3, -S3, sx, sy, sz);
    const R g1 = sel3(ax1, sx, sy, sz);
    rot_axis(ax1, C2, -S2, sx, sy, sz);
    const R g0 = sel3(ax0, sx, sy, sz);
    if (in.pi2 >= 0) heis_apply<R, PLAIN>(p, u, pk, in.pi2, g2, in.v2);
    if (in.pi1 >= 0) heis_apply<R, PLAIN>(p, u, pk, in.pi1, g1, in.v1);
    if (in.pi0 >= 0) heis_apply<R, PLAIN>(p, u, pk, in.pi0, g0, in.v0);
  }
  if (!u.skip_coef) {
    R c0 = R(1), s0 = R(0), c1 = R(1), s1 = R(0), c2 = R(1), s2 = R(0);
    sincos3(ax0 >= 0, ax1 >= 0, ax2 >= 0, in.v0.th * R(0.5), in.v1.th * R(0.5), in.v2.th * R(0.5), s0, c0, s1, c1, s2, c2);
    R ar, ai, br, bi;
    su2_of(ax0, c0, s0, ar, ai, br, bi);
    su2_lmul_axis(ax1, c1, s1, ar, ai, br, bi);
    su2_lmul_axis(ax2, c2, s2, ar, ai, br, bi);
    if (u.active) { ax[0] = c1; ax[1] = s1; ax[2] = c2; ax[3] = s2; }
    // ZYZ form for the forward sweep: alpha = cy p_a, beta = sy p_b, u_out = p_b conj(p_a), u_in = conj(p_a p_b)
    const R na = ar * ar + ai * ai, nb = br * br + bi * bi;
    const bool oka = na > R(1e-30), okb = nb > R(1e-30);
    const R ia = oka ? rsqrt_fast(na) : R(0), ib = okb ? rsqrt_fast(nb) : R(0);
    const R par = oka ? ar * ia : R(1), pai = ai * ia, pbr = okb ? br * ib : R(1), pbi = bi * ib;
    // Ry by phi with cos phi = cy = |alpha| >= 0, sin phi = sy = |beta| >= 0, in lifting form (Cols::ry_lift)
    const R cyv = na * ia, syv = nb * ib;
    cf[0] = -syv * rcp_fast(R(1) + cyv); cf[1] = syv;
    cf[2] = pbr * par + pbi * pai; cf[3] = pbi * par - pbr * pai;
    cf[4] = par * pbr - pai * pbi; cf[5] = -(par * pbi + pai * pbr);
  }
}
// packed axes of a gate class: a0 | a1 << 4 | a2 << 8 (15 = unused slot); 0xffff = not uniform
constexpr int AXP_ZXZ = 2 | (0 << 4) | (2 << 8);
constexpr int AXP_XYZ = 0 | (1 << 4) | (2 << 8);
constexpr int AXP_XZ = 0 | (2 << 4) | (15 << 8);

// gates g0, g0 + stride, ... < g_end of one class (compile-time axes), software pipelined.  (Issuing the first loads of
// all three gate classes together at the top of the parameter phase was measured: no gain in float - 121.6 vs 121.3 M
// evals/s on C3 - and 1.5 KB of spills in the double kernels; each loop keeps its own prologue.)
template <typename R, int AX0, int AX1, int AX2, bool PLAIN>
__device__ __forceinline__ void heis_su2_loop(const KParams<R>& p, const HSu2* ms, const UpdCtx<R>& u, Pk4<R>* pk, int g0,
                                              int g_end, int stride, R* coef, R* aux) {
  constexpr int SW = HEIS_SU2_WORDS;
  // Two gates in flight, loop unrolled by two so that the buffers keep their registers (no copies at the back
  // edge): the state of gate g + 2 stride is requested as soon as gate g is done and has the whole update of
  // gate g + stride to arrive (the L2 round trip is about as long as one gate's update).
  GateIn<R> ga = heis_gate_load(p, u, pk, g0 < g_end, p.su2 + g0, ms + g0, coef + SW * g0, aux + 4 * g0);
  GateIn<R> gb = heis_gate_load(p, u, pk, g0 + stride < g_end, p.su2 + g0 + stride, ms + g0 + stride,
                                coef + SW * (g0 + stride), aux + 4 * (g0 + stride));
#pragma unroll 1
  for (int g = g0; g < g_end; g += 2 * stride) {
    heis_su2_update<R, AX0, AX1, AX2, PLAIN>(p, u, pk, p.su2 + g, ga, coef + SW * g, aux + 4 * g);
    const int g2 = g + 2 * stride;
    ga = heis_gate_load(p, u, pk, g2 < g_end, p.su2 + g2, ms + g2, coef + SW * g2, aux + 4 * g2);
    const int g1 = g + stride;
    if (g1 < g_end) heis_su2_update<R, AX0, AX1, AX2, PLAIN>(p, u, pk, p.su2 + g1, gb, coef + SW * g1, aux + 4 * g1);
    const int g3 = g + 3 * stride;
    gb = heis_gate_load(p, u, pk, g3 < g_end, p.su2 + g3, ms + g3, coef + SW * g3, aux + 4 * g3);
  }
}
template <typename R>
__device__ __forceinline__ void heis_su2_loop_any(int axp, bool plain, const KParams<R>& p, const HSu2* ms,
                                                  const UpdCtx<R>& u, Pk4<R>* pk, int g0, int g_end, int stride, R* coef,
                                                  R* aux) {
  if (axp == AXP_XYZ) {
    if (plain) heis_su2_loop<R, 0, 1, 2, true>(p, ms, u, pk, g0, g_end, stride, coef, aux);
    else heis_su2_loop<R, 0, 1, 2, false>(p, ms, u, pk, g0, g_end, stride, coef, aux);
  } else if (axp == AXP_ZXZ) {
    if (plain) heis_su2_loop<R, 2, 0, 2, true>(p, ms, u, pk, g0, g_end, stride, coef, aux);
    else heis_su2_loop<R, 2, 0, 2, false>(p, ms, u, pk, g0, g_end, stride, coef, aux);
  } else if (axp == AXP_XZ) heis_su2_loop<R, 0, 2, -1, false>(p, ms, u, pk, g0, g_end, stride, coef, aux);
  else heis_su2_loop<R, -2, -2, -2, false>(p, ms, u, pk, g0, g_end, stride, coef, aux);
}

// The block size is a launch parameter (a multiple of 32 up to HCfg::MAXT); p.spb of its blockDim.x / TPS
// sample slots are used (heis_geometry spreads the batch evenly over SMs and rounds).
template <typename R, int NQ, int CPT, typename SWP>
__global__ void __launch_bounds__(HCfg<R, NQ, CPT>::MAXT, 1)
heis_kernel(const KParams<R> p) {
  using C = HCfg<R, NQ, CPT>;
  using T = VT<R, CPT>;
  using V = typename T::V;
  constexpr int N = C::N, TPS = C::TPS, PB = C::PB, XR = C::XR;
  constexpr int SW = HEIS_SU2_WORDS, CW = HEIS_CP_WORDS;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  R* s_target = reinterpret_cast<R*>(smem_raw);
  HSu2* s_su2 = reinterpret_cast<HSu2*>(smem_raw + p.target_bytes);
  HCp* s_cp = reinterpret_cast<HCp*>(s_su2 + p.n_su2);
  R* s_coef = reinterpret_cast<R*>(smem_raw + p.target_bytes + heis_meta_bytes(p.n_su2, p.n_cp));

  const int tid = threadIdx.x;
  // ---- prologue: TMA-stage V^dag (packed by pack_target_heis_kernel) ----
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&s_bar, (uint32_t)p.target_bytes);
    tma_bulk_g2s(s_target, p.target_packed, (uint32_t)p.target_bytes, &s_bar);
  }
  // gate metadata -> shared memory (compact records)
  for (int g = tid; g < p.n_su2; g += blockDim.x) {
    const Su2Meta* md = p.su2 + g;
    HSu2 hm;
    unsigned axes = 0;
    for (int j = 0; j < 3; ++j) {
      hm.pidx[j] = (int16_t)md->pidx[j];
      axes |= (unsigned)(md->axis[j] < 0 ? 15 : md->axis[j]) << (4 * j);
    }
    hm.axes = (uint16_t)axes;
    s_su2[g] = hm;
  }
  for (int k = tid; k < p.n_cp; k += blockDim.x) {
    const CpMeta* md = p.cp + k;
    HCp hm;
    hm.pidx = (int16_t)md->pidx;
    hm.flags = (uint16_t)(((p.cp_pen ? p.cp_pen[k] != 0 : md->penalised != 0) ? 1 : 0) | (md->is_cz ? 2 : 0) |
                          ((md->lo_q & 7) << 4) | ((md->hi_q & 7) << 8));
    hm.prev_lo = md->prev_lo; hm.prev_hi = md->prev_hi;
    s_cp[k] = hm;
  }
  __syncthreads();
  mbar_wait(&s_bar, 0);

  const int sl = tid / TPS;   // sample within the block
  const int m = tid % TPS;    // lane within the sample: column group (forward), x >> PB (backward)
  // ring position -> (sample, visit); an unsliced launch has ring_start = 0, ring_end = B (visit 0 for everybody)
  const long long c_raw = p.ring_start + (long long)blockIdx.x * p.spb + sl;
  const bool active = sl < p.spb && c_raw < p.ring_end;
  const long long c_pos = active ? c_raw : p.ring_end - 1;
  const long long b = c_pos % p.B;
  const long long step0 = p.step0 + (c_pos / p.B) * (long long)p.nsteps;
  const int P = p.P;
  // idle sample slots (block size rounded up to whole warps) replay the launch's last sample in one spare store
  R* coef = s_coef + (size_t)(sl < p.spb ? sl : p.spb) * p.coef_stride;
  R* stage = coef + SW * p.n_su2;                          // staged rows of one layer of the backward sweep
  R* coef_cp = stage + HEIS_STAGE_WORDS * SWP::NSTAGE;
  const V* tv = reinterpret_cast<const V*>(s_target) + 2 * (size_t)m * (N + 1);

  const unsigned off = (unsigned)(b * P);   // host: B * P < 2^32
  R* aux = p.aux + (size_t)b * p.n_su2 * 4;
  Pk4<R>* pk = reinterpret_cast<Pk4<R>*>(p.pk) + off;
  {
    // pack this sample's optimiser state (a resumed run, step0 > 0, carries its moments and best parameters)
    const bool resume = p.mode == M_ADAM && step0 > 0;
    for (int i = m; i < P; i += TPS) {
      Pk4<R> v;
      v.th = p.angles[off + i];
      v.mu = resume ? p.m[off + i] : R(0);
      v.nu = resume ? p.v[off + i] : R(0);
      v.best = resume ? p.best_params[off + i] : v.th;
      if (active) pk_store(pk + i, v);
    }
    __syncwarp();
  }
  const R NN = R(N) * R(N);

  LayerBar lb;
  lb.fwd = (p.sync_sweeps & 1) != 0; lb.bwd = (p.sync_sweeps & 2) != 0;

  R best = R(0), best_reg_v = R(0);
  bool improved_prev = false;
  if (p.mode == M_ADAM && step0 > 0) { best = p.best_regloss[b]; best_reg_v = p.best_reg[b]; }

  for (int it = 0; it <= p.nsteps; ++it) {
    const long long gi = step0 + it;
    const int phase = it == 0 ? PH_COEF : (p.mode == M_ADAM ? PH_ADAM : PH_GRAD);
    // ---------------- parameter phase (a sample's threads split the gates) ----------------
    R reg_part = R(0);
    {
      UpdCtx<R> u;
      u.phase = phase; u.active = active; u.off = off;
      const long long gu = gi - 1;
      u.store_best = improved_prev && p.mode == M_ADAM && phase == PH_ADAM;
      u.skip_coef = it == p.nsteps;
      u.hist = p.hist_params != nullptr && gu + 1 < p.hist_len;
      u.hist_off = u.hist ? (size_t)(b * p.hist_len + gu + 1) * (size_t)P : 0;
      u.bc1 = u.bc2 = u.ibc1 = u.ibc2 = R(1);
      if (phase == PH_ADAM) {
        u.bc1 = bias_corr(p.b1, R(gu + 1));
        u.bc2 = bias_corr(p.b2, R(gu + 1));
        u.ibc1 = R(1) / u.bc1; u.ibc2 = R(1) / u.bc2;
      }
      // surface gates (slots < NQ) and block gates (the rest) each share one axis pattern in the templates
      // plain: an Adam RUN without freeze mask and parameter history.  Every pass of such a run, the first
      // (coefficient-only) one of a launch included, goes through the same PLAIN instantiation of the gate loops:
      // split and time-sliced runs are bit-identical to one launch by construction, not by the compiler's choice of
      // identical FMA contractions in two instantiations
      const bool plain = p.mode == M_ADAM && p.freeze == nullptr && p.hist_params == nullptr;
      heis_su2_loop_any(p.axp_surface, plain, p, s_su2, u, pk, m, NQ < p.n_su2 ? NQ : p.n_su2, TPS, coef, aux);
      heis_su2_loop_any(p.axp_block, plain, p, s_su2, u, pk, NQ + m, p.n_su2, TPS, coef, aux);
      // entangler angles, software pipelined like the fused-gate loops: the packed state of the next gate is
      // requested before the current one is processed
      {
        int k = m;
        int pi_n = -1;
        Pk4<R> v_n = Pk4<R>{R(0), R(0), R(0), R(0)};
        HCp md_n = HCp{-1, 0, 0, 0};
        if (k < p.n_cp) { md_n = s_cp[k]; pi_n = md_n.pidx; if (pi_n >= 0) v_n = pk_load(pk + pi_n); }
#pragma unroll 1
        for (; k < p.n_cp; k += TPS) {
          const HCp md = md_n;
          R* cf = coef_cp + CW * k;
          const int pi = pi_n;
          Pk4<R> v = v_n;
          const int kn = k + TPS;
          if (kn < p.n_cp) { md_n = s_cp[kn]; pi_n = md_n.pidx; if (pi_n >= 0) v_n = pk_load(pk + pi_n); }
          const bool pen_on = p.pen.kind != CPF_PEN_NONE && pi >= 0 && (md.flags & 1) != 0;
          if (pi < 0) v.th = R(p.cp[k].cangle);
          // cf[0]: dL/da from the sweep; word 6 of the block's higher-qubit gate slot: r * penalty slope at this angle
          R* rsw = coef + SW * (NQ + 2 * k + 1) + 6;
          if (phase != PH_COEF && pi >= 0) {
            const R g = add_rn(cf[0], *rsw);
            if (plain) heis_apply<R, true>(p, u, pk, pi, g, v);
            else heis_apply<R, false>(p, u, pk, pi, g, v);
          }
          const R th = v.th;
          if (!u.skip_coef) {
            R s = R(1), c = R(0);                  // half angle; CZ = CP(pi): cos(pi/2) = 0 exactly
            if (!(md.flags & 2)) sincos_inl(th * R(0.5), s, c);
            R rs = R(0);
            if (pen_on) {
              R val, slope;
              penalty_eval_fast(p.pen, th, val, slope);
              reg_part += val;
              rs = mul_rn(p.pen.r, slope);
            }
            // CP(a) = CP(a - 2 pi): keep cos(a/2) >= 0, so the ZZ pair rotation of the backward sweep is a rotation
            // by at most pi/2 and its lifting coefficient t = -tan(a/4) is bounded (phase_bwd)
            if (c < R(0)) { c = -c; s = -s; }
            cf[0] = c; cf[1] = s; cf[2] = -s * rcp_fast(R(1) + c); *rsw = rs;
          }
        }
      }
    }
    __syncwarp();
    if (it == p.nsteps) break;
    // merged diagonals of the forward sweep: A = pending(lo) u_in(lo), B = pending(hi) u_in(hi), A B e^{ia}
    for (int k = m; k < p.n_cp; k += TPS) {
      const HCp md = s_cp[k];
      R* cl = coef + SW * (NQ + 2 * k);
      R* ch = cl + SW;
      const R* pl = coef + SW * md.prev_lo;
      const R* ph = coef + SW * md.prev_hi;
      const R* cc = coef_cp + CW * k;
      const R plr = pl[2], pli = pl[3], phr = ph[2], phi = ph[3];
      const R ar = plr * cl[4] - pli * cl[5], ai = plr * cl[5] + pli * cl[4];
      const R br = phr * ch[4] - phi * ch[5], bi = phr * ch[5] + phi * ch[4];
      const R abr = ar * br - ai * bi, abi = ar * bi + ai * br;
      const R c = cc[0] * cc[0] - cc[1] * cc[1], s = R(2) * cc[0] * cc[1];   // e^{ia} from the half angle
      cl[4] = ar; cl[5] = ai; cl[6] = br; cl[7] = bi;
      ch[4] = abr * c - abi * s; ch[5] = abr * s + abi * c;
    }
    __syncwarp();

    // ---------------- forward sweep: Y = U V^dag ----------------
    V yr[N], yi[N];
#pragma unroll
    for (int r = 0; r < N; ++r) { yr[r] = tv[2 * r]; yi[r] = tv[2 * r + 1]; }
    SWP::forward(p, lb, s_cp, coef, yr, yi);

    // ---------------- pivot to the Pauli basis ----------------
    SWP::gather_wht(yr, yi);
    const R tr = __shfl_sync(0xffffffffu, T::get(yr[0], 0), 0, TPS);
    const R ti = __shfl_sync(0xffffffffu, T::get(yi[0], 0), 0, TPS);
    R reg = sample_sum<TPS>(reg_part);
    const R ab = sqrt_r(tr * tr + ti * ti);
    const R loss = R(1) - mul_rn(ab, ab) / NN;
    reg = mul_rn(p.pen.r, reg);

    if (p.mode == M_LOSSGRAD) {
      if (active && m == 0) {
        p.loss_out[b] = loss;
        if (p.reg_out) p.reg_out[b] = reg;
      }
      if (!p.grad_out) return;
    } else {
      const R regloss = add_rn(loss, reg);
      bool improved;
      if (gi == 0) {
        improved = true;
        if (active && m == 0) { p.init_regloss[b] = regloss; p.init_reg[b] = reg; }
      } else {
        improved = regloss < best;
      }
      // the parameters of an improving step are saved by the next update phase, which has them in registers
      improved_prev = improved;
      if (improved) { best = regloss; best_reg_v = reg; }
      if (active && p.hist_regloss && m == 0 && gi < p.hist_len)
        p.hist_regloss[b * p.hist_len + gi] = regloss;
      if (active && p.hist_params && gi == 0)
        for (int i = m; i < P; i += TPS) p.hist_params[b * p.hist_len * P + i] = p.angles[off + i];
    }

    // h[xr][z] = Re(i^{|x&z|} s W[x,z]),  s = i conj(t)/N^2 = (ti + i tr)/N^2,  x = (m << PB) | xr
    V h[N];
    {
      const R sr = ti / NN, si = tr / NN;
#pragma unroll
      for (int z = 0; z < N; ++z) {
        R hc[XR];
#pragma unroll
        for (int xr = 0; xr < XR; ++xr) {
          const R wr = T::get(yr[z], xr), wi = T::get(yi[z], xr);
          const R pq = sr * wr - si * wi, qq = sr * wi + si * wr;
          const int k = __popc(m & (z >> PB)) + (PB ? (xr & z & 1) : 0);
          const R val = (k & 1) ? -qq : pq;
          hc[xr] = (k & 2) ? -val : val;
        }
        h[z] = T::make(hc[0], hc[XR - 1]);
      }
    }
    __syncwarp();

    // ---------------- Heisenberg sweep ----------------
    SWP::backward(p, lb, s_cp, coef, stage, coef_cp, m, h);
    __syncwarp();
  }

  if (p.mode == M_ADAM && active) {
    // unpack the optimiser state into the caller's arrays
    for (int i = m; i < P; i += TPS) {
      const Pk4<R> v = pk_load(pk + i);
      p.angles[off + i] = v.th; p.m[off + i] = v.mu; p.v[off + i] = v.nu; p.best_params[off + i] = v.best;
    }
    if (m == 0) { p.best_regloss[b] = best; p.best_reg[b] = best_reg_v; }
  }
}

// ---- target packing for heis_kernel: Y0 = V^dag, lane l = column group, all rows per lane ----
// dst index ((l * (N + 1) + r) * 2 + part) * CPT + k :  part 0 = Re, 1 = Im of conj(V[l*CPT + k][r])
template <typename R>
__global__ void pack_target_heis_kernel(const R* __restrict__ src, R* __restrict__ dst, int N, int cpt) {
  const int total = (N / cpt) * (N + 1) * 2 * cpt;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int k = i % cpt, t = i / cpt;
    int part = t % 2; t /= 2;
    int r = t % (N + 1), l = t / (N + 1);
    R val = R(0);
    if (r < N) {
      const R v = src[((size_t)(l * cpt + k) * N + r) * 2 + part];
      val = part ? -v : v;
    }
    dst[i] = val;
  }
}

// Launch geometry: spread the batch evenly over SMs and rounds so that the last wave is as full as the
// first (every CTA runs all the Adam steps of its samples, so a ragged last wave costs a whole wave).
//   ctas  resident CTAs per SM.  Each CTA is one synchronised instruction stream (its warps share the instruction
//         cache); two streams per SM overlap their phases (+4 % on C3 when both hold 8 warps) as long as the SM
//         keeps as many samples resident as with one: chosen automatically, env CPF_HEIS_CTAS overrides.
//   cap   samples per CTA allowed by shared memory, the register file, the thread limit and CPF_HEIS_WARPS
struct HeisGeometry { int block, spb, ctas; long long grid; size_t smem; };
inline long long heis_cap(int ctas, size_t fixed_bytes, size_t per_sample, int tps, int maxt, int regs, int warps_env) {
  // registers: allocated per warp in units of 256, 64 K per SM
  const int regs_warp = ((regs > 0 ? regs : 128) * 32 + 255) / 256 * 256;
  int warps = 65536 / regs_warp / ctas;
  if (warps > maxt / 32) warps = maxt / 32;
  if (warps_env > 0 && warps > warps_env) warps = warps_env;
  const long long cap_thr = (long long)warps * 32 / tps;
  const long long smem_cta = (long long)(227 * 1024) / ctas - 1024 - (long long)fixed_bytes;
  long long cap = smem_cta > 0 ? smem_cta / (long long)per_sample : 0;
  // a block whose last warp is only partly used parks the idle lanes on one spare slot
  if (cap > 0 && (cap * tps) % 32 != 0 && cap <= cap_thr) cap -= 1;
  if (cap > cap_thr) cap = cap_thr;
  return cap < 1 ? 1 : cap;
}
inline HeisGeometry heis_geometry(long long B, size_t fixed_bytes, size_t per_sample, int tps, int maxt, int regs,
                                  int n_sm = 0) {
  if (n_sm <= 0) {          // the current device (cpf_launch_plan passes a number to plan without one)
    int dev = 0;
    n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  int ctas = 0, warps_env = 0;
  if (const char* e = getenv("CPF_HEIS_CTAS")) { int v = atoi(e); if (v >= 1 && v <= 16) ctas = v; }
  if (const char* e = getenv("CPF_HEIS_WARPS")) { int v = atoi(e); if (v >= 1) warps_env = v; }
  const long long slots1 = n_sm;
  auto spread = [&](int nc, long long& spb_out) {
    const long long cap = heis_cap(nc, fixed_bytes, per_sample, tps, maxt, regs, warps_env);
    const long long slots = slots1 * nc;
    long long rounds = (B + slots * cap - 1) / (slots * cap);
    if (rounds < 1) rounds = 1;               // empty batch
    long long spb = (B + slots * rounds - 1) / (slots * rounds);
    if (spb > cap) spb = cap;
    if (spb < 1) spb = 1;
    spb_out = spb;
    return cap;
  };
  long long spb = 1;
  if (ctas == 0) {
    // measured on B200: two streams of 8 warps beat one of 16 (+4 % C3, +29 % 5 qubits), two of 6 lose to one of 11
    long long spb1, spb2;
    const long long c1 = spread(1, spb1), c2 = spread(2, spb2);
    ctas = 2 * c2 >= c1 && (spb2 * tps + 31) / 32 >= 8 ? 2 : 1;
    spb = ctas == 2 ? spb2 : spb1;
  } else {
    spread(ctas, spb);
  }
  HeisGeometry g;
  g.spb = (int)spb;
  g.ctas = ctas;
  g.block = (int)((spb * tps + 31) / 32 * 32);
  g.grid = (B + spb - 1) / spb;
  g.smem = fixed_bytes + (size_t)(spb + (g.block > spb * tps ? 1 : 0)) * per_sample;
  return g;
}

// Time slicing of an Adam run whose batch is not a whole number of full waves ("ring slicing").  Every CTA
// keeps its samples for all the steps of a launch, so a batch of 1.3 waves either runs as 2 rounds of 2/3-full CTAs
// (the even spread of heis_geometry) or, sliced, as a ring: the T steps are cut into k chunks of T / k, the B k
// (sample, chunk) items are laid on a ring in sample-major order, and consecutive launches of `slots` items each walk
// along it.  All launches but the last run at full residency; stream order guarantees that chunk v + 1 of a sample
// starts after its chunk v has been written back (the run is resumable by construction: split runs are bit-identical
// to one run).
// The choice is an empirical throughput model fitted to B200 measurements of the C3 kernel (profiles/r2_*perf*):
//   one launch      relative rate (s / s_full)^0.7 for s resident samples per SM, times a tail factor
//                   1 - 0.05 min(1, 3 / rounds) (a CTA runs all T steps, so with few rounds the SMs that finish first
//                   idle: 10^5 samples in 11 rounds 121 M evals/s, 5 x 10^4 in 6 rounds 111 M, 2.5 x 10^4 in 3: 109 M,
//                   1.25 x 10^4 in 2 rounds of 43: 102 M)
//   sliced          1 / (1.035 + 3 / C) for chunks of C steps (drift between CTAs exposed at every launch boundary,
//                   state pack / unpack and the coefficient pass), times the fill of the last launch
//                   (116-118 M evals/s at every batch size measured).
struct HeisSlicing { int k; long long slots; };
inline HeisSlicing heis_slicing(long long B, int nsteps, size_t fixed_bytes, size_t per_sample, int tps, int maxt, int regs,
                                int n_sm, bool allowed) {
  HeisSlicing best{1, 0};
  const HeisGeometry full = heis_geometry((long long)1 << 40, fixed_bytes, per_sample, tps, maxt, regs, n_sm);
  const long long slots = (long long)full.spb * full.ctas * n_sm;
  best.slots = slots;
  int forced = -1;
  if (const char* e = getenv("CPF_HEIS_SLICES")) forced = atoi(e);
  if (!allowed || forced == 0 || forced == 1 || B <= 0 || nsteps < 2) return best;
  // a ring shorter than one launch would put two visits of a sample into the same launch: never sliced
  if (B <= slots) return best;
  if (forced > 1) {
    if (nsteps % forced == 0) best.k = forced;
    return best;
  }
  const HeisGeometry g1 = heis_geometry(B, fixed_bytes, per_sample, tps, maxt, regs, n_sm);
  const long long rounds = (g1.grid + (long long)g1.ctas * n_sm - 1) / ((long long)g1.ctas * n_sm);
  const double occ = (double)g1.spb * g1.ctas / ((double)full.spb * full.ctas);
  double rate_best = pow(occ < 1.0 ? occ : 1.0, 0.7) * (1.0 - 0.05 * (rounds >= 3 ? 3.0 / (double)rounds : 1.0));
  for (int k = 2; k <= 64 && nsteps / k >= 20; ++k) {
    if (nsteps % k) continue;
    const double launches = (double)(B * k) / (double)slots;
    const double rate = 1.0 / (1.035 + 3.0 / (nsteps / k)) * launches / ceil(launches);
    if (rate > rate_best * 1.01) { rate_best = rate; best.k = k; }
  }
  return best;
}

template <typename R, int NQ, int CPT, typename SWP>
int launch_heis_sized(KParams<R> p, cudaStream_t st, std::string& err) {
  using C = HCfg<R, NQ, CPT>;
  p.n_sched = 0; p.n_red = 0;
  p.coef_stride = heis_coef_stride(p.n_su2, p.n_cp, SWP::NSTAGE);
  if (p.P > 32767) { err = "heis kernel: more than 32767 parameters"; return CPF_ERR_UNSUPPORTED; }
  auto kern = heis_kernel<R, NQ, CPT, SWP>;
  static std::atomic<int> regs_cached{0};      // per instantiation
  int regs = regs_cached.load(std::memory_order_relaxed);
  if (regs == 0) {
    cudaFuncAttributes fa;
    regs = cudaFuncGetAttributes(&fa, kern) == cudaSuccess ? fa.numRegs : 128;
    regs_cached.store(regs, std::memory_order_relaxed);
  }
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const size_t fixed = (size_t)p.target_bytes + heis_meta_bytes(p.n_su2, p.n_cp);
  const size_t per_sample = (size_t)p.coef_stride * sizeof(R);
  // CTA barrier at the start of the forward (bit 0) / backward (bit 1) sweep; env CPF_HEIS_SYNC overrides (tests)
  p.sync_sweeps = 3;
  if (const char* e = getenv("CPF_HEIS_SYNC")) { int v = atoi(e); if (v >= 0 && v <= 3) p.sync_sweeps = v; }
  const bool sliceable = p.mode == M_ADAM && p.hist_params == nullptr && p.hist_regloss == nullptr;
  const HeisSlicing sl = heis_slicing(p.B, p.nsteps, fixed, per_sample, C::TPS, C::MAXT, regs, n_sm, sliceable);
  const long long ring_total = p.B * sl.k;
  p.nsteps /= sl.k;
  long long smem_attr = 0;      // per call: cudaFuncSetAttribute is cheap, the limit is re-asserted for this device
  const long long per_launch = sl.k > 1 ? sl.slots : ring_total;      // unsliced: one launch over the whole batch
  for (long long c0 = 0; c0 < ring_total; c0 += per_launch) {
    const long long count = ring_total - c0 < per_launch ? ring_total - c0 : per_launch;
    const HeisGeometry g = heis_geometry(count, fixed, per_sample, C::TPS, C::MAXT, regs, n_sm);
    p.spb = g.spb;
    p.ring_start = c0; p.ring_end = c0 + count;
    if (g.smem > 227 * 1024) {
      err = "program too large for the shared-memory coefficient store (" + std::to_string(g.smem) + " bytes)";
      return CPF_ERR_UNSUPPORTED;
    }
    if ((long long)g.smem > smem_attr) {
      // the opt-in limit only ever grows (per instantiation and device; a smaller launch runs under a larger limit)
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
      if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return CPF_ERR_CUDA; }
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      smem_attr = (long long)g.smem;
    }
    if (g.grid <= 0 || count <= 0) return CPF_OK;
    if (g.grid > 2147483647LL) { err = "batch too large for one launch"; return CPF_ERR_UNSUPPORTED; }
    kern<<<(unsigned)g.grid, g.block, g.smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("kernel launch: ") + cudaGetErrorString(e); return CPF_ERR_CUDA; }
  }
  return CPF_OK;
}

// Returns true and sets `rc` when a Heisenberg kernel compiled for this layered program exists.
// `dry` only answers the question (used before the target is staged in the heis layout).
template <typename R> bool launch_heis(const KParams<R>& p, const Program& prog, cudaStream_t st,
                                       std::string& err, int& rc, bool dry, int* n_stage = nullptr);
// columns per thread of the heis kernels for (dtype, n): selects the target packing
template <typename R> int heis_cpt(int n_qubits);
template <typename R> int launch_pack_target_heis(const R* src, R* dst, int n_qubits, int cpt, cudaStream_t st);

}  // namespace cpf
