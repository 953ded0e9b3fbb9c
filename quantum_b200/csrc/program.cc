// See program.h.
#include "program.h"

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <set>
#include <tuple>

namespace tfqb {
namespace {

// absl::SimpleAtoi for int32: optional sign, decimal digits, surrounding
// ASCII whitespace allowed, no overflow.
bool simple_atoi(const std::string& s, int* out) {
  size_t b = 0, e = s.size();
  while (b < e && isspace((unsigned char)s[b])) ++b;
  while (e > b && isspace((unsigned char)s[e - 1])) --e;
  if (b == e) return false;
  bool neg = false;
  if (s[b] == '+' || s[b] == '-') { neg = s[b] == '-'; ++b; }
  if (b == e) return false;
  long long v = 0;
  for (size_t i = b; i < e; ++i) {
    if (!isdigit((unsigned char)s[i])) return false;
    v = v * 10 + (s[i] - '0');
    if (v > (long long)INT_MAX + 1) return false;
  }
  v = neg ? -v : v;
  if (v > INT_MAX || v < INT_MIN) return false;
  *out = int(v);
  return true;
}

std::vector<std::string> split(const std::string& s, char sep) {
  std::vector<std::string> out;
  size_t start = 0;
  for (;;) {
    const size_t pos = s.find(sep, start);
    if (pos == std::string::npos) { out.push_back(s.substr(start)); break; }
    out.push_back(s.substr(start, pos - start));
    start = pos + 1;
  }
  return out;
}

using QubitKey = std::tuple<int, int, std::string>;  // (row, col, id string)

// RegisterQubits, program_resolution.cc:49-88.
Status register_qubits(const std::string& qb_string, std::set<QubitKey>* ids) {
  if (qb_string.empty()) return Status::OK();
  for (const std::string& qb : split(qb_string, ',')) {
    std::vector<std::string> parts = split(qb, '_');
    if (parts.size() == 1) parts.insert(parts.begin(), "2147483647");
    int r, c;
    if (parts.size() != 2 || !simple_atoi(parts[0], &r) ||
        !simple_atoi(parts[1], &c))
      return Status::Error("Unable to parse qubit: " + qb);
    ids->insert(QubitKey(r, c, qb));
  }
  return Status::OK();
}

struct GateSpec {
  const char* id;
  int kind;
  int nq;
  int nparams;
  const char* params[5];
  // which params can carry a gradient symbol, in GateMetaData order
  int nsym;
  int sym_param[2];
};

// Arg names per gate id: circuit_parser_qsim.cc:204-562.
const GateSpec kSpecs[] = {
    {"I", kI, 1, 0, {}, 0, {}},
    {"I2", kI2, 2, 0, {}, 0, {}},
    {"XP", kXP, 1, 3, {"exponent", "exponent_scalar", "global_shift"}, 1, {0}},
    {"YP", kYP, 1, 3, {"exponent", "exponent_scalar", "global_shift"}, 1, {0}},
    {"ZP", kZP, 1, 3, {"exponent", "exponent_scalar", "global_shift"}, 1, {0}},
    {"HP", kHP, 1, 3, {"exponent", "exponent_scalar", "global_shift"}, 1, {0}},
    {"XXP", kXXP, 2, 3, {"exponent", "exponent_scalar", "global_shift"}, 1, {0}},
    {"YYP", kYYP, 2, 3, {"exponent", "exponent_scalar", "global_shift"}, 1, {0}},
    {"ZZP", kZZP, 2, 3, {"exponent", "exponent_scalar", "global_shift"}, 1, {0}},
    {"CZP", kCZP, 2, 3, {"exponent", "exponent_scalar", "global_shift"}, 1, {0}},
    {"CNP", kCNP, 2, 3, {"exponent", "exponent_scalar", "global_shift"}, 1, {0}},
    {"SP", kSP, 2, 3, {"exponent", "exponent_scalar", "global_shift"}, 1, {0}},
    {"ISP", kISP, 2, 3, {"exponent", "exponent_scalar", "global_shift"}, 1, {0}},
    {"PXP", kPXP, 1, 5,
     {"phase_exponent", "phase_exponent_scalar", "exponent", "exponent_scalar",
      "global_shift"}, 2, {0, 2}},
    {"FSIM", kFSIM, 2, 4, {"theta", "theta_scalar", "phi", "phi_scalar"}, 2,
     {0, 2}},
    {"PISP", kPISP, 2, 4,
     {"phase_exponent", "phase_exponent_scalar", "exponent", "exponent_scalar"},
     2, {0, 2}},
};

const GateSpec* find_spec(const std::string& id) {
  for (const auto& s : kSpecs)
    if (id == s.id) return &s;
  return nullptr;
}

}  // namespace

namespace {

// QsimCircuitFromProgram (circuit_parser_qsim.cc:828-861) on an already
// resolved qubit map (out->qubit_index, out->n).
struct ChannelSpec {
  const char* id;
  int type;
  int nargs;
  const char* args[3];
};
// ParseAppendChannel (circuit_parser_qsim.cc:744-771); arg names :598-742
const ChannelSpec kChannels[] = {
    {"DP", kChDP, 1, {"p"}},        {"ADP", kChADP, 3, {"p_x", "p_y", "p_z"}},
    {"GAD", kChGAD, 2, {"p", "gamma"}}, {"AD", kChAD, 1, {"gamma"}},
    {"RST", kChRST, 0, {}},         {"PD", kChPD, 1, {"gamma"}},
    {"PF", kChPF, 1, {"p"}},        {"BF", kChBF, 1, {"p"}},
};

Status lower_gates(const ProgramPB& pb, const SymbolTable& symbols, CircuitT* out,
                   bool allow_channels = false) {
  const int n = out->n;
  out->n_symbols = symbols.size;
  out->n_channels = 0;
  out->n_nonunitary = 0;
  for (const auto& m : pb.moments) {
    for (const auto& op : m.operations) {
      const GateSpec* spec = find_spec(op.gate_id);
      if (!spec && allow_channels) {
        const ChannelSpec* ch = nullptr;
        for (const auto& c : kChannels)
          if (op.gate_id == c.id) ch = &c;
        if (!ch) return Status::Error("Could not parse channel id: " + op.gate_id);
        if (op.qubits.empty())
          return Status::Error("Gate " + op.gate_id + " has too few qubits in op.");
        auto qit = out->qubit_index.find(op.qubits[0]);
        if (qit == out->qubit_index.end())
          return Status::Error("Unable to parse qubit: " + op.qubits[0]);
        GateT g;
        g.kind = kCH;
        g.nq = 1;
        g.bit[0] = n - qit->second - 1;
        g.nparams = 5;
        g.p[0].value = float(ch->type);
        for (int k = 0; k < ch->nargs; ++k) {
          const ArgPB* a = op.find(ch->args[k]);
          if (!a)
            return Status::Error(std::string("Could not find arg: ") + ch->args[k] + " in op.");
          g.p[1 + k].value = a->float_value;     // channels cannot hold symbols
        }
        g.p[4].sym = symbols.size + out->n_channels++;
        if (!ChannelIsMixture(ch->type)) g.aux_sym = out->n_nonunitary++;   // rebased below
        out->gates.push_back(g);
        continue;
      }
      if (!spec)
        return Status::Error(
            "Could not parse gate id: " + op.gate_id +
            ". This is likely because a cirq.Channel was used in an op that "
            "does not support them.");
      if (int(op.qubits.size()) < spec->nq)
        return Status::Error("Gate " + op.gate_id +
                             " has too few qubits in op.");
      GateT g;
      g.kind = spec->kind;
      g.nq = spec->nq;
      uint64_t used = 0;
      for (int k = 0; k < spec->nq; ++k) {
        // an empty id, or one holding ',' (registered as two ids), is not in
        // the map: the reference's absl::SimpleAtoi fails on both
        auto qit = out->qubit_index.find(op.qubits[k]);
        if (qit == out->qubit_index.end())
          return Status::Error("Unable to parse qubit: " + op.qubits[k]);
        g.bit[k] = n - qit->second - 1;
        used |= 1ull << g.bit[k];
      }
      if (spec->nq == 2 && g.bit[0] == g.bit[1])
        return Status::Error("Two-qubit gate acts on one qubit twice.");
      g.nparams = spec->nparams;
      // ParseProtoArg (circuit_parser_qsim.cc:53-82)
      for (int k = 0; k < spec->nparams; ++k) {
        const ArgPB* a = op.find(spec->params[k]);
        if (!a)
          return Status::Error(std::string("Could not find arg: ") +
                               spec->params[k] + " in op.");
        g.p[k].value = a->float_value;
        if (!a->symbol.empty()) {
          auto it = symbols.col.find(a->symbol);
          if (it == symbols.col.end())
            return Status::Error(
                "Could not find symbol in parameter map: " + a->symbol);
          g.p[k].sym = it->second;
        }
      }
      for (int k = 0; k < spec->nsym; ++k) {
        const int pi = spec->sym_param[k];
        if (g.p[pi].sym >= 0) {
          g.sym_param[g.nsym] = pi;
          g.sym_col[g.nsym] = g.p[pi].sym;
          ++g.nsym;
        }
      }
      // ParseProtoControls (circuit_parser_qsim.cc:84-129)
      const ArgPB* cq = op.find("control_qubits");
      const ArgPB* cv = op.find("control_values");
      if (!cv)
        return Status::Error("Operation is missing the control_values arg.");
      if (!(cq->string_value.empty() && cv->string_value.empty())) {
        const auto ctoks = split(cq->string_value, ',');
        const auto vtoks = split(cv->string_value, ',');
        if (ctoks.size() != vtoks.size())
          return Status::Error(
              "Mistmatched number of control qubits and control values.");
        for (size_t k = 0; k < ctoks.size(); ++k) {
          auto it = out->qubit_index.find(ctoks[k]);
          if (it == out->qubit_index.end())
            return Status::Error("Unable to parse qubit: " + ctoks[k]);
          int v;
          if (!simple_atoi(vtoks[k], &v) || v < 0)
            return Status::Error("Unparseable control value: " + vtoks[k]);
          const int b = n - it->second - 1;
          if ((used >> b) & 1)
            return Status::Error(
                "Control qubit overlaps a target or another control.");
          used |= 1ull << b;
          g.cmask |= 1ull << b;
          // qsim's cmask keeps one bit per control: any nonzero value is
          // truncated to its low bit there; 0/1 are the only values the
          // serializer writes (serializer.py:127-131).
          if (v & 1) g.cbits |= 1ull << b;
        }
      }
      out->gates.push_back(g);
    }
  }
  // the population columns follow the uniform columns
  for (GateT& g : out->gates)
    if (g.kind == kCH && g.aux_sym >= 0) g.aux_sym += symbols.size + out->n_channels;
  return Status::OK();
}

}  // namespace

SymbolTable MakeSymbolTable(const char* const* names, const size_t* lens,
                            int count) {
  SymbolTable t;
  t.size = count;
  for (int j = 0; j < count; ++j) t.col[std::string(names[j], lens[j])] = j;
  return t;
}

Status LowerProgram(const ProgramPB& pb, const SymbolTable& symbols,
                    CircuitT* out, bool allow_channels) {
  out->n = 0;
  out->gates.clear();
  out->qubit_index.clear();
  if (pb.moments.empty()) return Status::OK();  // (#679) empty program

  // ---- ResolveQubitIds (program_resolution.cc:90-186)
  std::set<QubitKey> ids;
  for (const auto& m : pb.moments) {
    for (const auto& op : m.operations) {
      for (const auto& q : op.qubits) {
        Status s = register_qubits(q, &ids);
        if (!s.ok) return s;
      }
      const ArgPB* cq = op.find("control_qubits");
      if (!cq)
        return Status::Error(
            "Operation is missing the control_qubits arg (serializer.py "
            "always writes it).");
      Status s = register_qubits(cq->string_value, &ids);
      if (!s.ok) return s;
    }
  }
  const int n = int(ids.size());
  if (n > 62)
    return Status::Error("Circuits with more than 62 qubits are unsupported.");
  int idx = 0;
  for (const auto& k : ids) out->qubit_index[std::get<2>(k)] = idx++;
  out->n = n;
  if (n <= 0) return Status::OK();

  return lower_gates(pb, symbols, out, allow_channels);
}

Status LowerPairedProgram(const ProgramPB& pb, const CircuitT& reference,
                          CircuitT* out) {
  out->n = reference.n;
  out->gates.clear();
  out->qubit_index = reference.qubit_index;
  std::set<std::string> unvisited;
  for (const auto& kv : reference.qubit_index) unvisited.insert(kv.first);
  for (const auto& m : pb.moments) {
    for (const auto& op : m.operations) {
      for (const auto& q : op.qubits) {
        unvisited.erase(q);
        if (!reference.qubit_index.count(q))
          return Status::Error(
              "A paired circuit contains qubits not found in reference circuit.");
      }
      const ArgPB* cq = op.find("control_qubits");
      if (!cq)
        return Status::Error("Operation is missing the control_qubits arg.");
      if (!cq->string_value.empty()) {
        for (const std::string& id : split(cq->string_value, ',')) {
          unvisited.erase(id);
          if (!reference.qubit_index.count(id))
            return Status::Error(
                "A paired circuit contains qubits not found in reference circuit.");
        }
      }
      for (const auto& a : op.args)
        if (!a.symbol.empty())
          return Status::Error(
              "Found symbols in other_programs.No symbols are allowed in these "
              "circuits.");
    }
  }
  if (!unvisited.empty())
    return Status::Error(
        "A reference circuit contains qubits not found in paired circuit.");
  SymbolTable none;
  return lower_gates(pb, none, out);
}

Status LowerPauliSum(const PauliSumPB& pb, const CircuitT& circuit,
                     PauliSumT* out) {
  out->terms.clear();
  const int n = circuit.n;
  for (const auto& t : pb.terms) {
    PauliTermT term;
    term.coeff = t.coefficient_real;
    term.identity = t.paulis.empty();
    for (const auto& p : t.paulis) {
      auto it = circuit.qubit_index.find(p.qubit_id);
      if (it == circuit.qubit_index.end()) {
        if (circuit.n == 0) continue;  // empty program: resolution is skipped
        return Status::Error(
            "Found a Pauli sum operating on qubits not found in circuit.");
      }
      const int b = n - it->second - 1;
      uint64_t x2 = 0, z2 = 0;
      int ny = 0, rot = 0;
      if (p.pauli_type == "X") { x2 = 1ull << b; rot = 1; }
      else if (p.pauli_type == "Y") { x2 = z2 = 1ull << b; ny = 1; rot = 2; }
      else if (p.pauli_type == "Z") { z2 = 1ull << b; }
      else
        return Status::Error(
            "Could not parse gate id: " + p.pauli_type +
            "P. This is likely because a cirq.Channel was used in an op that "
            "does not support them.");
      // left-multiply: W(x2,z2) W(x,z) = (-1)^{|z2 & x|} W(x^x2, z^z2)
      term.phase = (term.phase + ny + 2 * int(__builtin_popcountll(z2 & term.x) & 1)) & 3;
      term.x ^= x2;
      term.z ^= z2;
      term.parity_mask |= 1ull << b;
      if (rot) term.rot.emplace_back(b, rot);
    }
    out->terms.push_back(std::move(term));
  }
  return Status::OK();
}

}  // namespace tfqb
