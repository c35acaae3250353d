// Specialised engine kernels for the standard layered templates (float): straight-line sweeps with
// compile-time qubit pairs (engine_impl.cuh: LayerSweep).  Other layers run on the interpreter kernel.
// Generated list: n, layer name, pairs packed 4 bits per block (lower qubits, higher qubits).
#include "launch.cuh"

namespace cpf {

template <>
bool launch_layered<float>(const KParams<float>& p, const Program& prog, bool single, cudaStream_t st,
                           std::string& err, int& rc) {
  const int n = prog.n_qubits, nb = prog.period;
  const unsigned long long lo = prog.lo_pack, hi = prog.hi_pack;
  if (single) {
    // state preparation (one column per sample) on the standard 4- and 5-qubit layers (+17 % / +20 % over the
    // interpreter kernel); register bits as in inst_f32.cu (single_rb_f32).  Other layers: interpreter kernel.
#define CPF_SINGLE_LAYER(N_, RB_, NB_, LO_, HI_)                                                                  \
    if (n == N_ && nb == NB_ && lo == LO_ && hi == HI_) {                                                          \
      rc = launch_one<float, N_, RB_, 1, true, LayerSweep<float, N_, RB_, 1, true, NB_, LO_, HI_>>(p, st, err);    \
      return true;                                                                                                 \
    }
    CPF_SINGLE_LAYER(4, 1, 3, 0x210ull, 0x321ull)                    // 4q chain
    CPF_SINGLE_LAYER(4, 1, 3, 0x0ull, 0x321ull)                      // 4q star
    CPF_SINGLE_LAYER(4, 1, 6, 0x211000ull, 0x332321ull)              // 4q connected
    CPF_SINGLE_LAYER(5, 1, 4, 0x3210ull, 0x4321ull)                  // 5q chain
    CPF_SINGLE_LAYER(5, 1, 10, 0x3221110000ull, 0x4434324321ull)     // 5q connected
    // (6 and 7 qubits: the straight-line sweeps were measured slower than the interpreter, 31.2 vs 34.7 and 18.3 vs
    // 20.5 M evals/s on the chain templates with K = 60 - profiles/r2_exp_single_rb.txt)
#undef CPF_SINGLE_LAYER
    return false;
  }
  // 2q chain [(0, 1)]
  if (n == 2 && nb == 1 && lo == 0x0ull && hi == 0x1ull) {
    rc = launch_one<float, 2, 2, 2, false, LayerSweep<float, 2, 2, 2, false, 1, 0x0ull, 0x1ull>>(p, st, err);
    return true;
  }
  // 3q chain [(0, 1), (1, 2)]
  if (n == 3 && nb == 2 && lo == 0x10ull && hi == 0x21ull) {
    rc = launch_one<float, 3, 2, 2, false, LayerSweep<float, 3, 2, 2, false, 2, 0x10ull, 0x21ull>>(p, st, err);
    return true;
  }
  // 3q connected [(0, 1), (0, 2), (1, 2)]
  if (n == 3 && nb == 3 && lo == 0x100ull && hi == 0x221ull) {
    rc = launch_one<float, 3, 2, 2, false, LayerSweep<float, 3, 2, 2, false, 3, 0x100ull, 0x221ull>>(p, st, err);
    return true;
  }
  // 4q chain [(0, 1), (1, 2), (2, 3)]
  if (n == 4 && nb == 3 && lo == 0x210ull && hi == 0x321ull) {
    rc = launch_one<float, 4, 2, 2, false, LayerSweep<float, 4, 2, 2, false, 3, 0x210ull, 0x321ull>>(p, st, err);
    return true;
  }
  // 4q star [(0, 1), (0, 2), (0, 3)]
  if (n == 4 && nb == 3 && lo == 0x0ull && hi == 0x321ull) {
    rc = launch_one<float, 4, 2, 2, false, LayerSweep<float, 4, 2, 2, false, 3, 0x0ull, 0x321ull>>(p, st, err);
    return true;
  }
  // 4q connected [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
  if (n == 4 && nb == 6 && lo == 0x211000ull && hi == 0x332321ull) {
    rc = launch_one<float, 4, 2, 2, false, LayerSweep<float, 4, 2, 2, false, 6, 0x211000ull, 0x332321ull>>(p, st, err);
    return true;
  }
  // 5q chain [(0, 1), (1, 2), (2, 3), (3, 4)]
  if (n == 5 && nb == 4 && lo == 0x3210ull && hi == 0x4321ull) {
    rc = launch_one<float, 5, 4, 2, false, LayerSweep<float, 5, 4, 2, false, 4, 0x3210ull, 0x4321ull>>(p, st, err);
    return true;
  }
  // 5q connected [(0, 1), (0, 2), (0, 3), (0, 4), (1, 2), (1, 3), (1, 4), (2, 3), (2, 4), (3, 4)]
  if (n == 5 && nb == 10 && lo == 0x3221110000ull && hi == 0x4434324321ull) {
    rc = launch_one<float, 5, 4, 2, false, LayerSweep<float, 5, 4, 2, false, 10, 0x3221110000ull, 0x4434324321ull>>(p, st, err);
    return true;
  }
  return false;
}

}  // namespace cpf
