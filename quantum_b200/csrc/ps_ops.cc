// Parameter-shift helper ops: host-side rewrites of
// serialized tfq.proto.Program strings.  No device work: the reference runs
// them on the TF CPU thread pool too.
//
//   TfqPsDecompose           core/ops/tfq_ps_decompose_op.cc:43-328
//   TfqPsSymbolReplace       core/ops/tfq_ps_symbol_replace_op.cc:41-182
//   TfqPsWeightsFromSymbols  core/ops/tfq_ps_weights_from_symbols_op.cc:43-171
//
// Programs are decoded with the library's own wire decoder (wire.cc; there is
// no protobuf in this image) and re-encoded by the small encoder below, which
// writes what TFQ's serializer writes: language.gate_set = "tfq_gate_set",
// circuit.scheduling_strategy = MOMENT_BY_MOMENT, per operation the gate id,
// the args map (float / string / symbol / bool_values) and the qubit ids.
// What TFQ's serializer never writes (arg_function_language, ArgFunction,
// Schedule) is not carried over; map entries keep the order they were read in.
#include "ps_ops.h"

#include <cstring>

namespace tfqb {
namespace {

void PutVarint(std::string* o, uint64_t v) {
  while (v >= 0x80) {
    o->push_back(char((v & 0x7f) | 0x80));
    v >>= 7;
  }
  o->push_back(char(v));
}
void PutLen(std::string* o, int field, const std::string& payload) {
  PutVarint(o, (uint64_t(field) << 3) | 2u);
  PutVarint(o, payload.size());
  o->append(payload);
}
void PutFloat(std::string* o, int field, float v) {
  PutVarint(o, (uint64_t(field) << 3) | 5u);
  char b[4];
  memcpy(b, &v, 4);      // little-endian hosts only (x86-64, aarch64)
  o->append(b, 4);
}

// The decoder folds the Arg oneof into (float_value, string_value, symbol).
// TFQ's serializer writes strings only for the control metadata
// (serializer.py:738-755), so the kind is recovered from that.
bool IsStringArg(const ArgPB& a) {
  return !a.string_value.empty() || a.key == "control_qubits" || a.key == "control_values";
}

std::string EncodeArg(const ArgPB& a) {
  std::string arg;
  if (!a.symbol.empty()) {
    PutLen(&arg, 2, a.symbol);                 // Arg.symbol
    return arg;
  }
  std::string value;                           // ArgValue
  if (a.has_bools) PutLen(&value, 2, a.bools_wire);          // RepeatedBoolean, as read
  else if (IsStringArg(a)) PutLen(&value, 3, a.string_value);
  else PutFloat(&value, 1, a.float_value);
  PutLen(&arg, 1, value);                      // Arg.arg_value
  return arg;
}

std::string EncodeOperation(const OperationPB& op) {
  std::string o, gate;
  PutLen(&gate, 1, op.gate_id);
  PutLen(&o, 1, gate);
  for (const ArgPB& a : op.args) {
    std::string entry;
    PutLen(&entry, 1, a.key);
    PutLen(&entry, 2, EncodeArg(a));
    PutLen(&o, 2, entry);
  }
  for (const std::string& q : op.qubits) {
    std::string qb;
    PutLen(&qb, 2, q);                         // Qubit.id = 2
    PutLen(&o, 3, qb);
  }
  return o;
}

ArgPB FloatArg(const std::string& key, float v) {
  ArgPB a;
  a.key = key;
  a.float_value = v;
  return a;
}
ArgPB StringArg(const std::string& key, const std::string& v) {
  ArgPB a;
  a.key = key;
  a.string_value = v;
  return a;
}
ArgPB SymbolArg(const std::string& key, const std::string& s) {
  ArgPB a;
  a.key = key;
  a.symbol = s;
  return a;
}
float FloatOf(const OperationPB& op, const std::string& key) {
  const ArgPB* a = op.find(key);
  return a ? a->float_value : 0.f;        // map operator[] default in the reference
}
std::string StringOf(const OperationPB& op, const std::string& key) {
  const ArgPB* a = op.find(key);
  return a ? a->string_value : std::string();
}
void CopyControls(const OperationPB& from, OperationPB* to) {
  to->args.push_back(StringArg("control_qubits", StringOf(from, "control_qubits")));
  to->args.push_back(StringArg("control_values", StringOf(from, "control_values")));
}

// getOpForISP (tfq_ps_decompose_op.cc:156-179)
OperationPB OpForISP(const OperationPB& cur, const std::string& id, const std::string& symbol) {
  OperationPB n;
  n.gate_id = id;
  n.args.push_back(FloatArg("global_shift", -0.5f));
  n.args.push_back(FloatArg("exponent_scalar", FloatOf(cur, "exponent_scalar") * -0.5f));
  n.args.push_back(SymbolArg("exponent", symbol));
  CopyControls(cur, &n);
  n.qubits = cur.qubits;
  return n;
}

// exponent / exponent_scalar of a decomposed factor: symbol -> scaled scalar,
// literal -> scaled value with scalar 1 (:199-216, :243-260, :291-309)
void PutExponent(const ArgPB* target, float scalar, float factor, OperationPB* n) {
  if (target && !target->symbol.empty()) {
    n->args.push_back(FloatArg("exponent_scalar", factor * scalar));
    n->args.push_back(SymbolArg("exponent", target->symbol));
  } else {
    n->args.push_back(FloatArg("exponent_scalar", 1.0f));
    n->args.push_back(FloatArg("exponent", factor * (target ? target->float_value : 0.f)));
  }
}

// getOpForPXP (:181-228)
OperationPB OpForPXP(const OperationPB& cur, const std::string& id, const std::string& key,
                     bool sign_flip) {
  OperationPB n;
  n.gate_id = id;
  n.args.push_back(FloatArg("global_shift", 0.0f));
  PutExponent(cur.find(key), FloatOf(cur, key + "_scalar"), sign_flip ? -1.0f : 1.0f, &n);
  n.qubits = cur.qubits;
  CopyControls(cur, &n);
  return n;
}

// getOpForPISP (:230-273): Z^{+-phase_exponent} on the first or second qubit
OperationPB OpForPISP(const OperationPB& cur, bool sign_flip, bool use_target) {
  OperationPB n;
  n.gate_id = "ZP";
  n.args.push_back(FloatArg("global_shift", 0.0f));
  PutExponent(cur.find("phase_exponent"), FloatOf(cur, "phase_exponent_scalar"),
              sign_flip ? -1.0f : 1.0f, &n);
  if (cur.qubits.size() >= 2) n.qubits.push_back(cur.qubits[use_target ? 1 : 0]);
  CopyControls(cur, &n);
  return n;
}

// getOpForFSIM (:275-322)
OperationPB OpForFSIM(const OperationPB& cur, const std::string& id, const std::string& key,
                      bool use_global_shift) {
  OperationPB n;
  n.gate_id = id;
  n.args.push_back(FloatArg("global_shift", use_global_shift ? -0.5f : 0.0f));
  const float sign = key == "theta" ? 1.0f : -1.0f;
  const ArgPB* target = cur.find(key);
  const float scalar = FloatOf(cur, key + "_scalar");
  // the reference divides in float by the literal 3.14159265359
  if (target && !target->symbol.empty()) {
    n.args.push_back(FloatArg("exponent_scalar", float(sign * scalar / 3.14159265359)));
    n.args.push_back(SymbolArg("exponent", target->symbol));
  } else {
    n.args.push_back(FloatArg("exponent_scalar", 1.0f));
    n.args.push_back(
        FloatArg("exponent", float(sign * (target ? target->float_value : 0.f) / 3.14159265359)));
  }
  n.qubits = cur.qubits;
  CopyControls(cur, &n);
  return n;
}

bool IsSymbol(const OperationPB& op, const std::string& key) {
  const ArgPB* a = op.find(key);
  return a && !a->symbol.empty();
}

}  // namespace

std::string EncodeProgram(const ProgramPB& p) {
  std::string circuit;
  PutVarint(&circuit, (1u << 3) | 0u);   // scheduling_strategy = MOMENT_BY_MOMENT
  PutVarint(&circuit, 1);
  for (const MomentPB& m : p.moments) {
    std::string mo;
    for (const OperationPB& op : m.operations) PutLen(&mo, 1, EncodeOperation(op));
    PutLen(&circuit, 2, mo);
  }
  std::string out, lang;
  PutLen(&lang, 1, "tfq_gate_set");
  PutLen(&out, 1, lang);
  PutLen(&out, 2, circuit);
  return out;
}

std::string EmptyProgram() {
  // Program{language{gate_set}, circuit{}} (tfq_ps_symbol_replace_op.cc:134-139)
  std::string out, lang;
  PutLen(&lang, 1, "tfq_gate_set");
  PutLen(&out, 1, lang);
  PutLen(&out, 2, std::string());
  return out;
}

Status PsDecompose(const ProgramPB& in, ProgramPB* out) {
  out->moments.clear();
  for (const MomentPB& src : in.moments) {
    MomentPB cur = src;
    MomentPB extra[5];
    int n_extra = 0;
    for (size_t k = 0; k < cur.operations.size(); ++k) {
      const OperationPB op = cur.operations[k];   // copy: the slot is overwritten
      const std::string& id = op.gate_id;
      if (id == "PISP") {
        if (!op.find("exponent") || !op.find("phase_exponent"))
          return Status::Error("PISP operation without exponent / phase_exponent");
        if (IsSymbol(op, "exponent") || IsSymbol(op, "phase_exponent")) {
          n_extra = 5;
          const std::string sym = op.find("exponent")->symbol;
          cur.operations[k] = OpForPISP(op, false, false);
          extra[0].operations.push_back(OpForPISP(op, true, true));
          extra[1].operations.push_back(OpForISP(op, "XXP", sym));
          extra[2].operations.push_back(OpForISP(op, "YYP", sym));
          extra[3].operations.push_back(OpForPISP(op, true, false));
          extra[4].operations.push_back(OpForPISP(op, false, true));
        }
      } else if (id == "ISP") {
        if (!op.find("exponent")) return Status::Error("ISP operation without exponent");
        if (IsSymbol(op, "exponent")) {
          if (n_extra == 0) n_extra = 1;
          const std::string sym = op.find("exponent")->symbol;
          cur.operations[k] = OpForISP(op, "XXP", sym);
          extra[0].operations.push_back(OpForISP(op, "YYP", sym));
        }
      } else if (id == "PXP") {
        if (!op.find("exponent") || !op.find("phase_exponent"))
          return Status::Error("PXP operation without exponent / phase_exponent");
        if (IsSymbol(op, "exponent") || IsSymbol(op, "phase_exponent")) {
          n_extra = 2;
          cur.operations[k] = OpForPXP(op, "ZP", "phase_exponent", true);
          extra[0].operations.push_back(OpForPXP(op, "XP", "exponent", false));
          extra[1].operations.push_back(OpForPXP(op, "ZP", "phase_exponent", false));
        }
      } else if (id == "FSIM") {
        if (!op.find("theta") || !op.find("phi"))
          return Status::Error("FSIM operation without theta / phi");
        if (IsSymbol(op, "theta") || IsSymbol(op, "phi")) {
          n_extra = 2;
          cur.operations[k] = OpForFSIM(op, "XXP", "theta", true);
          extra[0].operations.push_back(OpForFSIM(op, "YYP", "theta", true));
          extra[1].operations.push_back(OpForFSIM(op, "CZP", "phi", false));
        }
      }
    }
    out->moments.push_back(std::move(cur));
    for (int l = 0; l < n_extra; ++l) out->moments.push_back(std::move(extra[l]));
  }
  return Status::OK();
}

void PsSymbolReplace(const ProgramPB& in, const std::string& symbol,
                     const std::string& replacement, std::vector<std::string>* out) {
  for (size_t j = 0; j < in.moments.size(); ++j)
    for (size_t k = 0; k < in.moments[j].operations.size(); ++k) {
      const OperationPB& op = in.moments[j].operations[k];
      for (size_t l = 0; l < op.args.size(); ++l) {
        if (op.args[l].symbol.empty() || op.args[l].symbol != symbol) continue;
        ProgramPB copy = in;
        copy.moments[j].operations[k].args[l].symbol = replacement;
        out->push_back(EncodeProgram(copy));
      }
    }
}

Status PsWeightsFromSymbols(const ProgramPB& in, const std::vector<std::string>& symbols,
                            std::vector<std::vector<float>>* out) {
  static const char* const kIgnored[] = {"I",  "ISP", "PXP", "FSIM", "PISP", "AD", "ADP",
                                         "DP", "GAD", "BF",  "PF",   "PD",   "RST"};
  out->assign(symbols.size(), std::vector<float>());
  for (const MomentPB& m : in.moments)
    for (const OperationPB& op : m.operations) {
      bool skip = false;
      for (const char* g : kIgnored) skip = skip || op.gate_id == g;
      if (skip) continue;
      const ArgPB* e = op.find("exponent");
      if (!e) return Status::Error("operation " + op.gate_id + " has no exponent");
      if (e->symbol.empty()) continue;
      size_t col = symbols.size();
      for (size_t s = 0; s < symbols.size(); ++s)
        if (symbols[s] == e->symbol) col = s;      // map semantics: last index wins
      if (col == symbols.size())
        return Status::Error("A circuit contains a sympy.Symbol not found in symbols!");
      (*out)[col].push_back(FloatOf(op, "exponent_scalar"));
    }
  return Status::OK();
}

}  // namespace tfqb
