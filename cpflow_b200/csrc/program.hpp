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
  int16_t penalised; // 1 if the default penalty mask covers this parameter
  int16_t is_cz;     // 1: CZ = diag(1,1,1,-1) exactly (no angle)
  // layered programs: slot of the previous fused gate on the block's lower / higher qubit (its outgoing
  // Rz phase is still pending when this block starts: heis_impl.cuh, forward with merged diagonals)
  int16_t prev_lo, prev_hi;
  int16_t lo_q, hi_q;   // the block's qubit pair, lower / higher qubit (block-structured programs)
  double cangle;
};

// ---- decoded schedule (what the kernels execute) -------------------------------------------
// One 64-bit word per scheduled op, specialised for a kernel configuration (NQ qubits, RB register
// bits): the kernel's inner loops do one jump-table dispatch per op and no index arithmetic.
//   x: case[0:8) | lane mask[8:16) (SU2 on a lane bit: xor mask; phase: required lane bits;
//      CX: control bit position [8:12), target bit position [12:16)) | flags[16:20) | acc base[20:24)
//   y: coefficient offset in words [0:16) | reduce-table index [16:32)
// Cases: 0..4 SU2 on register bit, 5 SU2 on a lane bit, 6 CX, 8 + rm: diagonal phase on the
// amplitudes whose register bits `rm` are set.
enum DecCase : uint32_t { DC_SU2_REG0 = 0, DC_SU2_LANE = 5, DC_CX = 6, DC_PHASE0 = 8 };
constexpr uint32_t DF_PARAM = 1u;   // the adjoint sweep accumulates this op's gradient sums
constexpr uint32_t DF_REDUCE = 2u;  // reduce and store the accumulated sums after this op
struct DecodedSchedule {
  std::vector<uint32_t> ops;   // 2 words per op (x, y)
  std::vector<uint16_t> red;   // 8 destination offsets (words, 0xffff = unused) per reduce group
};
struct DeviceDecoded {
  uint32_t* ops = nullptr;
  uint16_t* red = nullptr;
  int n_red = 0;
};

struct DeviceProgram {
  uint32_t* sched = nullptr;
  Su2Meta* su2 = nullptr;
  CpMeta* cp = nullptr;
  DeviceDecoded dec[8];   // indexed by RB (register bits of the kernel configuration)
  bool has_dec[8] = {false, false, false, false, false, false, false, false};
};

struct Program {
  int n_qubits = 0, n_params = 0;
  std::vector<cpf_op> ops;
  std::vector<uint32_t> sched;
  std::vector<Su2Meta> su2;
  std::vector<CpMeta> cp;
  std::vector<uint8_t> is_cp_param;  // [P]
  int n_rot = 0, n_phase = 0;
  // block structure (detect_layered): surface SU2 per qubit, then blocks [phase(lo,hi), SU2(lo), SU2(hi)].
  // `layered`: the program has this structure (Heisenberg kernels, heis_impl.cuh); `period` > 0: the qubit pairs
  // repeat with that period (compile-time-layer kernels), 0: no period of at most 16 blocks
  bool layered = false;
  int period = 0;
  unsigned long long lo_pack = 0, hi_pack = 0;   // 4 bits per block of the layer
  int last_slot[16] = {0};                        // slot of the last fused gate on each qubit
  // lazily created per-device copies
  mutable std::mutex mu;
  mutable std::unordered_map<int, DeviceProgram> dev;
};

// Returns empty string on success, else an error message.
std::string compile_program(Program& p);

// words of per-sample coefficient storage: 8 per fused SU(2) gate, 4 per phase gate
inline int coef_words(int n_su2, int n_cp) { return 8 * n_su2 + 4 * n_cp; }

// Recognise the layered template structure and renumber the SU2 slots into its canonical order
// (slot q: surface gate of qubit q; slots n+2k, n+2k+1: block k's lower / higher qubit).
void detect_layered(Program& p);

// Specialise the schedule for a kernel with `rb` register bits.
DecodedSchedule decode_schedule(const Program& p, int rb);

}  // namespace cpf
