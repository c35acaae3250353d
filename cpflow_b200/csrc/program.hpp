// Host-side gate program: validation, single-qubit gate fusion and the device schedule.
//
// The primitive program (include/cpflow_b200.h: cpf_op) follows the reference's gate order
// (cpflow/main.py:119-146, cpflow/main.py:69-82).  For the device it is compiled into a
// schedule of
//   * fused SU(2) gates: up to 3 consecutive rotations on one qubit (e.g. the surface round
//     Rz Rx Rz, main.py:122-124, or the Rx Ry Rz tail of an entangling block, main.py:77-80)
//     become one 2x2 unitary [[alpha, -conj(beta)], [beta, conj(alpha)]];
//   * diagonal two-qubit phases (CP with a parameter or a constant angle, CZ);
//   * CX permutations.
// Rotations on different qubits commute, so pending rotations are only flushed when a
// two-qubit gate touches their qubit (or three have accumulated).
#pragma once
#include <cstdint>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "cpflow_b200.h"

namespace cpf {

enum SchedKind : uint32_t { S_SU2 = 0, S_CP = 1, S_CZ = 2, S_CX = 3 };

// Packed schedule word: kind[0:4) | q0[4:8) | q1[8:12) | flags[12:16) | slot[16:32)
// q0/q1 are qubit indices (CP/CZ: q0 < q1; CX: q0 control, q1 target).
constexpr uint32_t FLAG_HAS_PARAM = 1u;
inline uint32_t pack_op(uint32_t kind, uint32_t q0, uint32_t q1, uint32_t flags, uint32_t slot) {
  return kind | (q0 << 4) | (q1 << 8) | (flags << 12) | (slot << 16);
}

// Metadata of one fused SU(2) gate, consumed by the parameter phase of the kernels.
struct Su2Meta {
  int32_t pidx[3];   // parameter index per rotation, -1 = constant / unused
  int8_t axis[3];    // 0,1,2 = x,y,z ; -1 = unused slot (identity)
  int8_t nrot;
  double cangle[3];  // constant angle when pidx < 0
};
struct CpMeta {
  int32_t pidx;      // -1 = constant angle
  int32_t penalised; // 1 if the default penalty mask covers this parameter
  double cangle;
};

struct DeviceProgram {
  uint32_t* sched = nullptr;
  Su2Meta* su2 = nullptr;
  CpMeta* cp = nullptr;
};

struct Program {
  int n_qubits = 0, n_params = 0;
  std::vector<cpf_op> ops;
  std::vector<uint32_t> sched;
  std::vector<Su2Meta> su2;
  std::vector<CpMeta> cp;
  std::vector<uint8_t> is_cp_param;  // [P]
  int n_rot = 0, n_phase = 0;
  // lazily created per-device copies
  mutable std::mutex mu;
  mutable std::unordered_map<int, DeviceProgram> dev;
};

// Returns empty string on success, else an error message.
std::string compile_program(Program& p);

}  // namespace cpf
