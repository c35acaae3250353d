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
          m.is_cz = 0;
          if (op.param >= 0) p.is_cp_param[op.param] = 1;
          uint32_t slot = (uint32_t)p.cp.size();
          p.cp.push_back(m);
          p.sched.push_back(pack_op(S_CP, lo, hi,
                                    op.param >= 0 ? FLAG_HAS_PARAM : 0, slot));
        } else if (op.kind == CPF_CZ) {
          if (op.param >= 0) { err << "op " << i << ": CZ takes no parameter"; return err.str(); }
          CpMeta m{}; m.pidx = -1; m.cangle = 0.0; m.penalised = 0; m.is_cz = 1;
          uint32_t slot = (uint32_t)p.cp.size();
          p.cp.push_back(m);
          p.sched.push_back(pack_op(S_CZ, lo, hi, 0, slot));
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
  if (p.sched.size() > 60000 || p.su2.size() > 60000 || p.cp.size() > 60000 ||
      coef_words((int)p.su2.size(), (int)p.cp.size()) > 65000)
    return "program too long for the 16-bit slot field";
  detect_layered(p);
  return "";
}

void detect_layered(Program& p) {
  p.layered = false;
  const int n = p.n_qubits;
  const int K = (int)p.cp.size();
  if ((int)p.su2.size() != n + 2 * K || n > 15) return;
  // state[q]: slot the next SU2 on qubit q must take (-1: none expected)
  std::vector<int> state(n), new_slot(p.su2.size(), -1);
  std::vector<int> lo(K), hi(K);
  for (int q = 0; q < n; ++q) state[q] = q;
  int k = 0;
  for (uint32_t op : p.sched) {
    const uint32_t kind = op & 15, q0 = (op >> 4) & 15, q1 = (op >> 8) & 15, slot = op >> 16;
    if (kind == S_SU2) {
      if (state[q0] < 0) return;
      new_slot[slot] = state[q0];
      state[q0] = -1;
    } else if (kind == S_CP || kind == S_CZ) {
      if (state[q0] >= 0 || state[q1] >= 0 || (int)slot != k) return;
      lo[k] = (int)q0; hi[k] = (int)q1;
      state[q0] = n + 2 * k; state[q1] = n + 2 * k + 1;
      ++k;
    } else {
      return;
    }
  }
  for (int q = 0; q < n; ++q) if (state[q] >= 0) return;
  int period = 0;
  for (int c = 1; c <= 16 && c <= (K > 0 ? K : 1); ++c) {
    bool ok = true;
    for (int i = c; i < K && ok; ++i) ok = lo[i] == lo[i - c] && hi[i] == hi[i - c];
    if (ok) { period = c; break; }
  }
  if (K == 0) period = 1;
  // renumber
  std::vector<Su2Meta> su2(p.su2.size());
  for (size_t s = 0; s < p.su2.size(); ++s) su2[new_slot[s]] = p.su2[s];
  p.su2.swap(su2);
  for (uint32_t& op : p.sched)
    if ((op & 15) == S_SU2) op = (op & 0xffffu) | ((uint32_t)new_slot[op >> 16] << 16);
  p.lo_pack = p.hi_pack = 0;
  for (int j = 0; j < period && j < K; ++j) {
    p.lo_pack |= (unsigned long long)lo[j] << (4 * j);
    p.hi_pack |= (unsigned long long)hi[j] << (4 * j);
  }
  p.period = period;
  p.layered = true;
  // pending-phase bookkeeping of the merged-diagonal forward sweep
  for (int q = 0; q < n; ++q) p.last_slot[q] = q;
  for (int j = 0; j < K; ++j) {
    p.cp[j].prev_lo = (int16_t)p.last_slot[lo[j]];
    p.cp[j].prev_hi = (int16_t)p.last_slot[hi[j]];
    p.cp[j].lo_q = (int16_t)lo[j];
    p.cp[j].hi_q = (int16_t)hi[j];
    p.last_slot[lo[j]] = n + 2 * j;
    p.last_slot[hi[j]] = n + 2 * j + 1;
  }
}

DecodedSchedule decode_schedule(const Program& p, int rb) {
  DecodedSchedule d;
  const int n = p.n_qubits;
  const int n_su2 = (int)p.su2.size();
  d.ops.resize(2 * p.sched.size());
  for (size_t i = 0; i < p.sched.size(); ++i) {
    const uint32_t op = p.sched[i];
    const uint32_t kind = op & 15, q0 = (op >> 4) & 15, q1 = (op >> 8) & 15, slot = op >> 16;
    const bool hp = ((op >> 12) & FLAG_HAS_PARAM) != 0;
    uint32_t x = 0, y = 0;
    if (kind == S_SU2) {
      const int bp = n - 1 - (int)q0;
      if (bp < rb) x = DC_SU2_REG0 + bp;
      else x = DC_SU2_LANE | ((1u << (bp - rb)) << 8);
      y = 8 * slot;
    } else if (kind == S_CP || kind == S_CZ) {
      const int pa = n - 1 - (int)q0, pb = n - 1 - (int)q1;
      const uint32_t rm = (pa < rb ? 1u << pa : 0u) | (pb < rb ? 1u << pb : 0u);
      const uint32_t lm = (pa >= rb ? 1u << (pa - rb) : 0u) | (pb >= rb ? 1u << (pb - rb) : 0u);
      x = (DC_PHASE0 + rm) | (lm << 8);
      y = 8 * n_su2 + 4 * slot;
    } else {
      x = DC_CX | ((uint32_t)(n - 1 - (int)q0) << 8) | ((uint32_t)(n - 1 - (int)q1) << 12);
    }
    if (hp) x |= DF_PARAM << 16;
    d.ops[2 * i] = x; d.ops[2 * i + 1] = y;
  }
  // accumulator slots in the order of the adjoint sweep: SU2 sums use slots {0,1,2} or {3,4,5},
  // phase sums slot 6 or 7; a group is reduced when the next op does not fit.
  int last = -1;
  bool su2_used[2] = {false, false}, ph_used[2] = {false, false};
  uint16_t cur[8];
  auto reset = [&]() { su2_used[0] = su2_used[1] = ph_used[0] = ph_used[1] = false;
                       for (int k = 0; k < 8; ++k) cur[k] = 0xffff; };
  auto close = [&]() {
    if (last < 0) return;
    d.ops[2 * last] |= DF_REDUCE << 16;
    d.ops[2 * last + 1] |= (uint32_t)(d.red.size() / 8) << 16;
    d.red.insert(d.red.end(), cur, cur + 8);
    last = -1;
    reset();
  };
  reset();
  for (int i = (int)p.sched.size() - 1; i >= 0; --i) {
    uint32_t& x = d.ops[2 * i];
    if (!((x >> 16) & DF_PARAM)) continue;
    const uint32_t cs = x & 0xff;
    const uint16_t off = (uint16_t)(d.ops[2 * i + 1] & 0xffff);
    if (cs <= DC_SU2_LANE) {
      if (su2_used[0] && su2_used[1]) close();
      const int g = su2_used[0] ? 1 : 0;
      su2_used[g] = true;
      x |= (uint32_t)(3 * g) << 20;
      cur[3 * g] = off; cur[3 * g + 1] = off + 1; cur[3 * g + 2] = off + 2;
    } else {
      if (ph_used[0] && ph_used[1]) close();
      const int g = ph_used[0] ? 1 : 0;
      ph_used[g] = true;
      x |= (uint32_t)(6 + g) << 20;
      cur[6 + g] = off;
    }
    last = i;
  }
  close();
  if (d.red.empty()) d.red.assign(8, 0xffff);
  return d;
}

}  // namespace cpf
