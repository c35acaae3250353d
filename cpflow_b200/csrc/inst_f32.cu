// float32 / complex64 instantiations of the engine (the reference's default x32 mode).
// launch_one<R, NQ, RB, CPT, SINGLE>: RB register bits, NQ-RB lane bits, CPT packed columns.
#include "launch.cuh"

namespace cpf {

template <>
int launch_engine<float>(const KParams<float>& p, int n, bool single, cudaStream_t st, std::string& err) {
  if (single) {
    switch (n) {
      case 2: return launch_one<float, 2, 2, 1, true>(p, st, err);
      case 3: return launch_one<float, 3, 2, 1, true>(p, st, err);
      case 4: return launch_one<float, 4, 2, 1, true>(p, st, err);
      case 5: return launch_one<float, 5, 3, 1, true>(p, st, err);
      case 6: return launch_one<float, 6, 3, 1, true>(p, st, err);
      case 7: return launch_one<float, 7, 4, 1, true>(p, st, err);
    }
  } else {
    switch (n) {
      case 2: return launch_one<float, 2, 2, 2, false>(p, st, err);
      case 3: return launch_one<float, 3, 2, 2, false>(p, st, err);
      case 4: return launch_one<float, 4, 2, 2, false>(p, st, err);
      case 5: return launch_one<float, 5, 4, 2, false>(p, st, err);
    }
  }
  err = single ? "unsupported number of qubits"
               : "full-unitary losses need n <= 5 qubits (6-7 qubits: state preparation and cpf_unitary only)";
  return CPF_ERR_UNSUPPORTED;
}

template <>
int engine_rb<float>(int n, bool single) {
  if (single) {
    switch (n) { case 2: return 2; case 3: return 2; case 4: return 2; case 5: return 3; case 6: return 3; case 7: return 4; }
  } else {
    switch (n) { case 2: return 2; case 3: return 2; case 4: return 2; case 5: return 4; }
  }
  return -1;
}

template <>
int launch_pack_target<float>(const float* src, float* dst, int n, int cpt, bool single, cudaStream_t st) {
  pack_target_kernel<float><<<8, 256, 0, st>>>(src, dst, 1 << n, cpt, single ? 1 : 0);
  return cudaGetLastError() == cudaSuccess ? CPF_OK : CPF_ERR_CUDA;
}

}  // namespace cpf
