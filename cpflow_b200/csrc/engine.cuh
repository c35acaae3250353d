// sm_100a device engine for the cpflow hot path.
//
// One launch runs, for every sample of the batch and `nsteps` times:
//   parameter phase : sin/cos of the angles, fuse up to three rotations into one SU(2)
//                     (gates.py:22-35, main.py:77-80, main.py:122-124), CP phases (gates.py:51-58),
//                     penalty value (penalty.py:44-71)
//   forward sweep   : apply the gate schedule to the identity -> U      (main.py:106-146)
//   loss            : HS / state-prep / relative-phase trace reduction  (matrix_utils.py:35-42)
//   best tracking   : strict <, pre-update parameters                   (optimization.py:70-73)
//   adjoint sweep   : walk the schedule backwards with inverse gates, O(1) memory; per gate
//                     accumulate Im<lambda|sigma_a|phi>
//   update phase    : gradients from the Pauli sums, penalty slope, optax Adam
//                     (optimization.py:14-25), then the next parameter phase.
//
// Data layout.  A thread owns CPT whole columns of the 2^n x 2^n unitary in registers
// (amplitude index = compile-time register index, so every single-qubit gate is a register-
// local butterfly on mask 1<<(n-1-q) and a CP gate is a phase FMA on a register subset — no
// shuffles on the gate path).  With CPT = 2 the two columns are packed in float2 registers and
// all gate arithmetic issues as packed FFMA2/FMUL2 (fma.rn.f32x2, sm_100+): the FP32 pipe
// saturates at half the issue slots, which leaves room for the LDS/SHFL/integer traffic.
// The TPS = 2^n / CPT threads of a sample sit in one warp; cross-column sums (the loss trace,
// the per-gate Pauli sums) are xor-butterflies over those lanes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "program.hpp"

namespace cpf {

enum Mode : int { M_ADAM = 0, M_LOSSGRAD = 1, M_UNITARY = 2, M_COTANGENT = 3 };

template <typename R>
struct PenaltyT {
  int kind, nseg;
  R r, period;
  R lo[CPF_MAX_SEGMENTS], hi[CPF_MAX_SEGMENTS], slope[CPF_MAX_SEGMENTS], icpt[CPF_MAX_SEGMENTS];
};

template <typename R>
struct KParams {
  const uint32_t* sched; int n_sched;
  const Su2Meta* su2; int n_su2;
  const CpMeta* cp; int n_cp;
  const uint8_t* cp_pen;  // device [n_cp] or null: which CP params are penalised
  int P; long long B;
  int mode, loss_kind;
  const R* target_packed; int target_bytes;
  R lr, b1, b2, eps, omb1, omb2;
  long long step0; int nsteps;
  R* angles; R* m; R* v; const uint8_t* freeze;
  R* best_params; R* best_regloss; R* best_reg; R* init_regloss; R* init_reg;
  R* hist_params; R* hist_regloss; long long hist_len;
  R* loss_out; R* reg_out; R* grad_out;
  R* u_out; const R* cot;
  int coef_stride;  // R words per sample in shared memory
  PenaltyT<R> pen;
};

// ------------------------------------------------------------------------------------------
// packed-vector abstraction: V holds CPT lanes of R
// ------------------------------------------------------------------------------------------
template <typename R, int CPT> struct VT;
template <> struct VT<float, 1> {
  using V = float;
  static __device__ __forceinline__ V bc(float a) { return a; }
  static __device__ __forceinline__ V mul(V a, V b) { return a * b; }
  static __device__ __forceinline__ V fma(V a, V b, V c) { return fmaf(a, b, c); }
  static __device__ __forceinline__ float hsum(V a) { return a; }
  static __device__ __forceinline__ float get(V a, int) { return a; }
  static __device__ __forceinline__ V onehot(int j, int col0) { return j == col0 ? 1.f : 0.f; }
  static __device__ __forceinline__ V make(float a, float) { return a; }
};
template <> struct VT<double, 1> {
  using V = double;
  static __device__ __forceinline__ V bc(double a) { return a; }
  static __device__ __forceinline__ V mul(V a, V b) { return a * b; }
  static __device__ __forceinline__ V fma(V a, V b, V c) { return ::fma(a, b, c); }
  static __device__ __forceinline__ double hsum(V a) { return a; }
  static __device__ __forceinline__ double get(V a, int) { return a; }
  static __device__ __forceinline__ V onehot(int j, int col0) { return j == col0 ? 1.0 : 0.0; }
  static __device__ __forceinline__ V make(double a, double) { return a; }
};
template <> struct VT<float, 2> {
  using V = float2;
  static __device__ __forceinline__ V bc(float a) { return make_float2(a, a); }
  static __device__ __forceinline__ V mul(V a, V b) { return __fmul2_rn(a, b); }
  static __device__ __forceinline__ V fma(V a, V b, V c) { return __ffma2_rn(a, b, c); }
  static __device__ __forceinline__ float hsum(V a) { return a.x + a.y; }
  static __device__ __forceinline__ float get(V a, int k) { return k ? a.y : a.x; }
  static __device__ __forceinline__ V onehot(int j, int col0) {
    return make_float2(j == col0 ? 1.f : 0.f, j == col0 + 1 ? 1.f : 0.f);
  }
  static __device__ __forceinline__ V make(float a, float b) { return make_float2(a, b); }
};

// ------------------------------------------------------------------------------------------
// scalar helpers
// ------------------------------------------------------------------------------------------
static __device__ __noinline__ void sincos_slow_f(float x, float* s, float* c) { sincosf(x, s, c); }

// sin/cos with Cody-Waite reduction and the Cephes single-precision kernels (~1 ulp), no
// local-memory slow path inlined.  XLA-class accuracy is required for 1e-5 parity.
__device__ __forceinline__ void sincos_r(float x, float& s, float& c) {
  if (fabsf(x) > 48000.f) { sincos_slow_f(x, &s, &c); return; }
  float j = rintf(x * 0.636619747f);
  float r = fmaf(j, -1.57079601e+00f, x);
  r = fmaf(j, -3.13916473e-07f, r);
  r = fmaf(j, -5.39030253e-15f, r);
  int q = __float2int_rn(j);
  float r2 = r * r;
  float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = fmaf(sp, r2, -1.6666654611e-1f);
  sp = fmaf(sp * r2, r, r);
  float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = fmaf(cp, r2, 4.166664568298827e-2f);
  cp = fmaf(cp * r2, r2, fmaf(r2, -0.5f, 1.0f));
  float ss = (q & 1) ? cp : sp;
  float cc = (q & 1) ? sp : cp;
  s = (q & 2) ? -ss : ss;
  c = ((q + 1) & 2) ? -cc : cc;
}
__device__ __forceinline__ void sincos_r(double x, double& s, double& c) { sincos(x, &s, &c); }

// IEEE operations that must NOT be contracted into FMAs (optax/XLA evaluate the Adam update and
// the penalty line as separate multiplies and adds; the oracle does the same).
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sqrt_r(float a) { return sqrtf(a); }
__device__ __forceinline__ double sqrt_r(double a) { return sqrt(a); }
__device__ __forceinline__ float pow_r(float a, float b) { return powf(a, b); }
__device__ __forceinline__ double pow_r(double a, double b) { return pow(a, b); }
__device__ __forceinline__ float fmod_r(float a, float b) { return fmodf(a, b); }
__device__ __forceinline__ double fmod_r(double a, double b) { return fmod(a, b); }
__device__ __forceinline__ float abs_r(float a) { return fabsf(a); }
__device__ __forceinline__ double abs_r(double a) { return fabs(a); }

template <typename R>
__device__ __forceinline__ R shfl_xor_r(R x, int m) { return __shfl_xor_sync(0xffffffffu, x, m); }

template <int TPS, typename R>
__device__ __forceinline__ R sample_sum(R x) {
#pragma unroll
  for (int m = TPS / 2; m >= 1; m >>= 1) x += shfl_xor_r(x, m);
  return x;
}

// jnp.mod semantics (result has the sign of the divisor; period > 0)
template <typename R>
__device__ __forceinline__ R pymod(R a, R period) {
  R r = fmod_r(a, period);
  if (r != R(0) && r < R(0)) r = add_rn(r, period);
  return r;
}

// penalty value and slope at angle a (penalty.py:44-76)
template <typename R>
__device__ __forceinline__ void penalty_eval(const PenaltyT<R>& pen, R a, R& val, R& slope) {
  val = R(0); slope = R(0);
  if (pen.kind == CPF_PEN_PIECEWISE) {
    R am = pymod(a, pen.period);
    for (int s = 0; s < pen.nseg; ++s) {
      if (pen.lo[s] < am && am <= pen.hi[s]) {
        val = add_rn(mul_rn(pen.slope[s], am), pen.icpt[s]);
        slope = pen.slope[s];
        break;
      }
    }
  } else if (pen.kind == CPF_PEN_L1) {
    val = abs_r(a);
    slope = a > R(0) ? R(1) : (a < R(0) ? R(-1) : R(0));
  }
}

// rotate vector (x,y,z) by angle with cos C, sin S about coordinate axis `a` (right-handed):
// R_a(theta) sigma_b R_a(theta)^dagger = sum_c [Rot_a(theta)]_{cb} sigma_c
template <typename R>
__device__ __forceinline__ void rot_axis(int a, R C, R S, R& x, R& y, R& z) {
  if (a == 0) { R ny = C * y - S * z, nz = S * y + C * z; y = ny; z = nz; }
  else if (a == 1) { R nz = C * z - S * x, nx = S * z + C * x; z = nz; x = nx; }
  else if (a == 2) { R nx = C * x - S * y, ny = S * x + C * y; x = nx; y = ny; }
}

// SU(2) element (alpha, beta) of a rotation about axis a with half-angle cos c / sin s:
// [[alpha, -conj(beta)], [beta, conj(alpha)]]  (gates.py:22-35)
template <typename R>
__device__ __forceinline__ void su2_of(int a, R c, R s, R& ar, R& ai, R& br, R& bi) {
  ar = c; ai = R(0); br = R(0); bi = R(0);
  if (a == 0) bi = -s;        // Rx: beta = -i s
  else if (a == 1) br = s;    // Ry: beta = s
  else if (a == 2) ai = -s;   // Rz: alpha = c - i s
}

// ------------------------------------------------------------------------------------------
// gate application on register-resident columns
// ------------------------------------------------------------------------------------------
template <typename R, int NQ, int CPT>
struct Cols {
  using T = VT<R, CPT>;
  using V = typename T::V;
  static constexpr int N = 1 << NQ;

  // x' = alpha x - conj(beta) y ; y' = beta x + conj(alpha) y on qubit Q (big-endian mask)
  template <int Q>
  static __device__ __forceinline__ void su2(V (&re)[N], V (&im)[N], R ar, R ai, R br, R bi) {
    constexpr int M = 1 << (NQ - 1 - Q);
    const V Ar = T::bc(ar), Ai = T::bc(ai), nAi = T::bc(-ai);
    const V Br = T::bc(br), nBr = T::bc(-br), Bi = T::bc(bi), nBi = T::bc(-bi);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (j & M) continue;
      const V xr = re[j], xi = im[j], yr = re[j | M], yi = im[j | M];
      re[j] = T::fma(nBi, yi, T::fma(nBr, yr, T::fma(nAi, xi, T::mul(Ar, xr))));
      im[j] = T::fma(Bi, yr, T::fma(nBr, yi, T::fma(Ai, xr, T::mul(Ar, xi))));
      re[j | M] = T::fma(Ai, yi, T::fma(Ar, yr, T::fma(nBi, xi, T::mul(Br, xr))));
      im[j | M] = T::fma(nAi, yr, T::fma(Ar, yi, T::fma(Bi, xr, T::mul(Br, xi))));
    }
  }

  // Pauli sums S_a = Im <lam| sigma_a |phi> on qubit Q, split in positive / negative parts so
  // that every term is a plain FMA (no operand negation on the packed path).
  template <int Q>
  static __device__ __forceinline__ void pauli_sums(const V (&pr)[N], const V (&pi)[N],
                                                    const V (&lr)[N], const V (&li)[N],
                                                    R& sx, R& sy, R& sz) {
    constexpr int M = 1 << (NQ - 1 - Q);
    V xp = T::bc(R(0)), xn = xp, yp = xp, yn = xp, zp = xp, zn = xp;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (j & M) continue;
      const int k = j | M;
      zp = T::fma(lr[j], pi[j], zp); zn = T::fma(li[j], pr[j], zn);
      zp = T::fma(li[k], pr[k], zp); zn = T::fma(lr[k], pi[k], zn);
      xp = T::fma(lr[j], pi[k], xp); xn = T::fma(li[j], pr[k], xn);
      xp = T::fma(lr[k], pi[j], xp); xn = T::fma(li[k], pr[j], xn);
      yp = T::fma(lr[k], pr[j], yp); yn = T::fma(lr[j], pr[k], yn);
      yp = T::fma(li[k], pi[j], yp); yn = T::fma(li[j], pi[k], yn);
    }
    sx = T::hsum(xp) - T::hsum(xn);
    sy = T::hsum(yp) - T::hsum(yn);
    sz = T::hsum(zp) - T::hsum(zn);
  }

  // multiply amplitudes with both bits set by (c + i s)
  template <int QA, int QB>
  static __device__ __forceinline__ void phase(V (&re)[N], V (&im)[N], R c, R s) {
    constexpr int M = (1 << (NQ - 1 - QA)) | (1 << (NQ - 1 - QB));
    const V C = T::bc(c), S = T::bc(s), nS = T::bc(-s);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if ((j & M) != M) continue;
      const V xr = re[j], xi = im[j];
      re[j] = T::fma(nS, xi, T::mul(C, xr));
      im[j] = T::fma(S, xr, T::mul(C, xi));
    }
  }
  template <int QA, int QB>
  static __device__ __forceinline__ void negate(V (&re)[N], V (&im)[N]) {
    constexpr int M = (1 << (NQ - 1 - QA)) | (1 << (NQ - 1 - QB));
    const V m1 = T::bc(R(-1));
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if ((j & M) != M) continue;
      re[j] = T::mul(m1, re[j]); im[j] = T::mul(m1, im[j]);
    }
  }
  // sum over |11> amplitudes of Im(conj(lam) phi)
  template <int QA, int QB>
  static __device__ __forceinline__ R phase_sum(const V (&pr)[N], const V (&pi)[N],
                                                const V (&lr)[N], const V (&li)[N]) {
    constexpr int M = (1 << (NQ - 1 - QA)) | (1 << (NQ - 1 - QB));
    V p = T::bc(R(0)), n = p;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if ((j & M) != M) continue;
      p = T::fma(lr[j], pi[j], p); n = T::fma(li[j], pr[j], n);
    }
    return T::hsum(p) - T::hsum(n);
  }
  // CNOT: swap target bit where control bit is set (register renaming, no arithmetic)
  template <int QC, int QT>
  static __device__ __forceinline__ void cnot(V (&re)[N], V (&im)[N]) {
    constexpr int MC = 1 << (NQ - 1 - QC), MT = 1 << (NQ - 1 - QT);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (!(j & MC) || (j & MT)) continue;
      V t = re[j]; re[j] = re[j | MT]; re[j | MT] = t;
      t = im[j]; im[j] = im[j | MT]; im[j | MT] = t;
    }
  }
};

// ---- compile-time qubit dispatch ----------------------------------------------------------
#define CPF_Q_SWITCH(NQ, q, ...)                                               \
  switch (q) {                                                                 \
    case 0: { constexpr int Q = 0; __VA_ARGS__; } break;                              \
    case 1: { constexpr int Q = 1; __VA_ARGS__; } break;                              \
    case 2: if constexpr (NQ > 2) { constexpr int Q = 2; __VA_ARGS__; } break;        \
    case 3: if constexpr (NQ > 3) { constexpr int Q = 3; __VA_ARGS__; } break;        \
    case 4: if constexpr (NQ > 4) { constexpr int Q = 4; __VA_ARGS__; } break;        \
    default: break;                                                            \
  }

// pair index -> (QA < QB), lexicographic, must match cpf::pair_index
template <int NQ, int IDX> struct PairOf {
  static __host__ __device__ constexpr int a() { int i = IDX, a = 0; while (i >= NQ - 1 - a) { i -= NQ - 1 - a; ++a; } return a; }
  static __host__ __device__ constexpr int b() { int i = IDX, a = 0; while (i >= NQ - 1 - a) { i -= NQ - 1 - a; ++a; } return a + 1 + i; }
};
#define CPF_PAIR_CASE(NQ, I, ...)                                              \
  case I: if constexpr (I < NQ * (NQ - 1) / 2) {                               \
    constexpr int QA = PairOf<NQ, I>::a(); constexpr int QB = PairOf<NQ, I>::b(); __VA_ARGS__; } break;
#define CPF_PAIR_SWITCH(NQ, p, ...)                                            \
  switch (p) {                                                                 \
    CPF_PAIR_CASE(NQ, 0, __VA_ARGS__) CPF_PAIR_CASE(NQ, 1, __VA_ARGS__) CPF_PAIR_CASE(NQ, 2, __VA_ARGS__)  \
    CPF_PAIR_CASE(NQ, 3, __VA_ARGS__) CPF_PAIR_CASE(NQ, 4, __VA_ARGS__) CPF_PAIR_CASE(NQ, 5, __VA_ARGS__)  \
    CPF_PAIR_CASE(NQ, 6, __VA_ARGS__) CPF_PAIR_CASE(NQ, 7, __VA_ARGS__) CPF_PAIR_CASE(NQ, 8, __VA_ARGS__)  \
    CPF_PAIR_CASE(NQ, 9, __VA_ARGS__)                                                 \
    default: break;                                                            \
  }
// ordered (control, target) dispatch for CX
#define CPF_QQ_SWITCH(NQ, qc, qt, ...)                                         \
  CPF_Q_SWITCH(NQ, qc, { constexpr int QC = Q; switch (qt) {                   \
    case 0: if constexpr (QC != 0) { constexpr int QT = 0; __VA_ARGS__; } break;      \
    case 1: if constexpr (QC != 1) { constexpr int QT = 1; __VA_ARGS__; } break;      \
    case 2: if constexpr (NQ > 2 && QC != 2) { constexpr int QT = 2; __VA_ARGS__; } break; \
    case 3: if constexpr (NQ > 3 && QC != 3) { constexpr int QT = 3; __VA_ARGS__; } break; \
    case 4: if constexpr (NQ > 4 && QC != 4) { constexpr int QT = 4; __VA_ARGS__; } break; \
    default: break; } })

// ------------------------------------------------------------------------------------------
// kernel configuration
// ------------------------------------------------------------------------------------------
template <typename R, int NQ, int CPT, bool SINGLE>
struct Cfg {
  static constexpr int N = 1 << NQ;
  static constexpr int COLS = SINGLE ? 1 : N;
  static constexpr int TPS = COLS / CPT;             // threads per sample (<= 32)
  static constexpr int BLOCK = SINGLE ? 32 : 128;
  static constexpr int SPB = BLOCK / TPS;            // samples per block
  static_assert(TPS >= 1 && TPS <= 32, "a sample must fit in one warp");
  // packed target in shared memory: per (column group, amplitude): V re, V im
  static constexpr int TARGET_WORDS = (SINGLE ? N : N * N) * 2;
};

// per-sample coefficient storage (R words): su2 gate g at [8g, 8g+8) =
//   {alpha_r, alpha_i, beta_r, beta_i, c2, s2, c3, s3}; the adjoint sweep overwrites the first
//   three words with the Pauli sums (Sx, Sy, Sz).  CP gate k at [8*n_su2 + 2k, +2) =
//   {cos a, sin a}; overwritten with the |11> sum.
__host__ __device__ inline int coef_words(int n_su2, int n_cp) { return 8 * n_su2 + 2 * n_cp; }

}  // namespace cpf
