#include "program.hpp"

#include <sstream>

namespace cpf {

std::string compile_program(Program& p) {
  const int n = p.n_qubits;
  std::ostringstream err;
  if (n < 2 || n > CPF_MAX_QUBITS) {
    err << "n_qubits=" << n << " outside supported range [2," << CPF_MAX_QUBITS << "]";
    return err.str();
  }
  std::vector<int> use(p.n_params, 0);
  p.is_cp_param.assign(p.n_params, 0);
  struct Pending { int axis[3]; int pidx[3]; double c[3]; int n = 0; };
  std::vector<Pending> pend(n);
  p.sched.clear(); p.su2.clear(); p.cp.clear();
  p.n_rot = p.n_phase = 0;

  auto flush = [&](int q) {
    Pending& pd = pend[q];
    if (pd.n == 0) return;
    Su2Meta m{};
    bool has_param = false;
    for (int k = 0; k < 3; ++k) {
      if (k < pd.n) {
        m.axis[k] = (int8_t)pd.axis[k]; m.pidx[k] = pd.pidx[k]; m.cangle[k] = pd.c[k];
        has_param |= pd.pidx[k] >= 0;
      } else {
        m.axis[k] = -1; m.pidx[k] = -1; m.cangle[k] = 0.0;
      }
    }
    m.nrot = (int8_t)pd.n;
    uint32_t slot = (uint32_t)p.su2.size();
    p.su2.push_back(m);
    p.sched.push_back(pack_op(S_SU2, q, 0, has_param ? FLAG_HAS_PARAM : 0, slot));
    pd.n = 0;
  };

  for (size_t i = 0; i < p.ops.size(); ++i) {
    const cpf_op& op = p.ops[i];
    if (op.param >= p.n_params || op.param < -1) {
      err << "op " << i << ": param index " << op.param << " out of range";
      return err.str();
    }
    if (op.param >= 0 && use[op.param]++) {
      err << "op " << i << ": parameter " << op.param << " feeds more than one gate (unsupported)";
      return err.str();
    }
    if (op.q0 < 0 || op.q0 >= n) { err << "op " << i << ": q0 out of range"; return err.str(); }
    switch (op.kind) {
      case CPF_RX: case CPF_RY: case CPF_RZ: {
        Pending& pd = pend[op.q0];
        pd.axis[pd.n] = op.kind;  // CPF_RX..RZ == 0..2 == x,y,z
        pd.pidx[pd.n] = op.param; pd.c[pd.n] = op.const_angle; pd.n++;
        p.n_rot++;
        if (pd.n == 3) flush(op.q0);
        break;
      }
      case CPF_CP: case CPF_CZ: case CPF_CX: {
        if (op.q1 < 0 || op.q1 >= n || op.q1 == op.q0) {
          err << "op " << i << ": bad qubit pair (" << op.q0 << "," << op.q1 << ")";
          return err.str();
        }
        flush(op.q0); flush(op.q1);
        int lo = op.q0 < op.q1 ? op.q0 : op.q1, hi = op.q0 < op.q1 ? op.q1 : op.q0;
        if (op.kind == CPF_CP) {
          CpMeta m{}; m.pidx = op.param; m.cangle = op.const_angle; m.penalised = op.param >= 0;
          if (op.param >= 0) p.is_cp_param[op.param] = 1;
          uint32_t slot = (uint32_t)p.cp.size();
          p.cp.push_back(m);
          p.sched.push_back(pack_op(S_CP, lo, hi,
                                    op.param >= 0 ? FLAG_HAS_PARAM : 0, slot));
        } else if (op.kind == CPF_CZ) {
          if (op.param >= 0) { err << "op " << i << ": CZ takes no parameter"; return err.str(); }
          p.sched.push_back(pack_op(S_CZ, lo, hi, 0, 0));
        } else {
          if (op.param >= 0) { err << "op " << i << ": CX takes no parameter"; return err.str(); }
          p.sched.push_back(pack_op(S_CX, op.q0, op.q1, 0, 0));
        }
        p.n_phase++;
        break;
      }
      default:
        err << "op " << i << ": unknown gate kind " << op.kind;
        return err.str();
    }
  }
  for (int q = 0; q < n; ++q) flush(q);
  if (p.sched.size() > 60000 || p.su2.size() > 60000 || p.cp.size() > 60000)
    return "program too long for the 16-bit slot field";
  return "";
}

}  // namespace cpf
