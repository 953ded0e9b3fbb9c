// Host-side circuit preparation: qubit-id resolution, proto Operation ->
// gate templates, PauliSum -> mask form.  Restates the *host* half of the
// reference:
//   ResolveQubitIds           core/src/program_resolution.cc:49-186
//   QsimCircuitFromProgram    core/src/circuit_parser_qsim.cc:53-596,828-861
//   GetSymbolMaps             core/ops/parse_context.cc:291-347
//   QsimCircuitFromPauliTerm  core/src/circuit_parser_qsim.cc:863-945
// B200-first differences: a program is lowered ONCE per distinct serialized
// string into *templates* whose parameters are (symbol column | literal)
// references; the per-row float32 matrices are evaluated on the GPU
// (gates.cuh) instead of per row on the host.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "wire.h"

namespace tfqb {

struct Status {
  bool ok = true;
  std::string msg;
  static Status OK() { return Status(); }
  static Status Error(const std::string& m) {
    Status s;
    s.ok = false;
    s.msg = m;
    return s;
  }
};

// Gate ids of circuit_parser_qsim.cc:577-584. Keep in sync with gates.cuh.
enum GateKind : int {
  kI = 0, kI2, kXP, kYP, kZP, kHP, kXXP, kYYP, kZZP, kCZP, kCNP, kSP, kISP,
  kPXP, kFSIM, kPISP, kNumGateKinds
};

struct ParamRef {
  int32_t sym = -1;   // column in symbol_names, or -1 for a literal
  float value = 0.f;  // literal value when sym < 0
};

struct GateT {
  int kind = kI;
  int nq = 1;
  int bit[2] = {0, 0};        // amplitude-index bit of each target, in the
                              // operation's own qubit order (bit = n-1-id)
  uint64_t cmask = 0;         // control bits (amplitude-index positions)
  uint64_t cbits = 0;         // required values on those bits
  int nparams = 0;
  ParamRef p[5];              // reference order, e.g. (exp, exp_s, gs)
  // symbols that produce gradient gates (GateMetaData, circuit_parser_qsim.h
  // :35-65): index into p[] of the shifted parameter, and the symbol column.
  int nsym = 0;
  int sym_param[2] = {0, 0};
  int sym_col[2] = {0, 0};

  uint64_t target_mask() const {
    uint64_t m = 1ull << bit[0];
    if (nq == 2) m |= 1ull << bit[1];
    return m;
  }
  bool is_identity() const { return kind == kI || kind == kI2; }
  // diagonal in the computational basis for every parameter value
  bool is_diagonal() const {
    return kind == kZP || kind == kZZP || kind == kCZP || is_identity();
  }
};

struct CircuitT {
  int n = 0;                  // number of qubits (0 = empty program)
  std::vector<GateT> gates;   // moment order
  std::unordered_map<std::string, int> qubit_index;  // id string -> 0..n-1
};

// PauliTerm in mask form over amplitude-index bits:
//   P = i^phase * prod_b X_b^{x_b} Z_b^{z_b};  P|k> = i^phase (-1)^{|k&z|}|k^x>
struct PauliTermT {
  float coeff = 0.f;          // coefficient_real (imag is ignored upstream)
  uint64_t x = 0, z = 0;
  int phase = 0;              // power of i, mod 4
  bool identity = false;      // term with no paulis (util_qsim.h:154-158)
  // Z-basis change for sampling (circuit_parser_qsim.cc:897-945), term order:
  // rot kind per pauli: 0 = none (Z), 1 = Y^-0.5 (for X), 2 = X^+0.5 (for Y)
  std::vector<std::pair<int, int>> rot;  // (bit, rot kind)
  uint64_t parity_mask = 0;   // util_qsim.h:241-256
};

struct PauliSumT {
  std::vector<PauliTermT> terms;
};

struct SymbolTable {
  std::unordered_map<std::string, int> col;  // later duplicates win
  int size = 0;
};

SymbolTable MakeSymbolTable(const char* const* names, const size_t* lens,
                            int count);

// Parse + resolve + lower one program. `n == 0` for an empty program.
Status LowerProgram(const ProgramPB& pb, const SymbolTable& symbols,
                    CircuitT* out);

// Lower a symbol-free "paired" program against the qubit ids of a reference
// circuit (ResolveQubitIds(Program*, unsigned*, vector<Program>*),
// program_resolution.cc:188-311): every qubit of the paired circuit must exist
// in the reference and every reference qubit must be touched by it.
Status LowerPairedProgram(const ProgramPB& pb, const CircuitT& reference,
                          CircuitT* out);

// Resolve one PauliSum against a lowered circuit's qubit ids.
Status LowerPauliSum(const PauliSumPB& pb, const CircuitT& circuit,
                     PauliSumT* out);

}  // namespace tfqb
