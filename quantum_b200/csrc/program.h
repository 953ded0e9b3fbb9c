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
  kPXP, kFSIM, kPISP,
  // a 1-qubit noise channel of the noisy trajectory ops (next-row N2;
  // circuit_parser_qsim.cc:598-771): the "gate" is the Kraus operator the
  // trajectory draws, see gates.cuh channel_matrix.  Parameters: p[0] = the
  // ChannelType, p[1..3] = its literal arguments, p[4] = the trajectory's
  // uniform for this channel (a parameter column past the symbols)
  kCH,
  kNumGateKinds
};

// Channel ids of circuit_parser_qsim.cc:752-756, Kraus operators in qsim's
// order (lib/channels_cirq.h).  Mixtures of unitaries draw without looking at
// the state; the others need the population of |1> on their qubit first.
enum ChannelType : int {
  kChADP = 0,   // (p_x, p_y, p_z): I X Y Z
  kChDP,        // (p): I X Y Z with p/3 each
  kChBF,        // (p): I X
  kChPF,        // (p): I Z
  kChAD,        // (gamma)
  kChPD,        // (gamma)
  kChRST,       // ()
  kChGAD,       // (p, gamma)
};
inline bool ChannelIsMixture(int type) { return type <= kChPF; }

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
  // kCH, non-unitary channels only: parameter column that holds the measured
  // population of |1> on the channel's qubit (written on the device just
  // before the channel), else -1
  int aux_sym = -1;

  uint64_t target_mask() const {
    uint64_t m = 1ull << bit[0];
    if (nq == 2) m |= 1ull << bit[1];
    return m;
  }
  bool is_identity() const { return kind == kI || kind == kI2; }
  bool is_channel() const { return kind == kCH; }
  bool needs_population() const { return kind == kCH && aux_sym >= 0; }
  // diagonal in the computational basis for every parameter value
  bool is_diagonal() const {
    return kind == kZP || kind == kZZP || kind == kCZP || is_identity();
  }
};

struct CircuitT {
  int n = 0;                  // number of qubits (0 = empty program)
  std::vector<GateT> gates;   // moment order
  std::unordered_map<std::string, int> qubit_index;  // id string -> 0..n-1
  // noisy programs (LowerProgram with allow_channels): a trajectory's
  // parameter row is [symbols | one uniform per channel | one measured
  // population per non-unitary channel]
  int n_symbols = 0;
  int n_channels = 0;
  int n_nonunitary = 0;
  int param_cols() const { return n_symbols + n_channels + n_nonunitary; }
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
// `allow_channels`: NoisyQsimCircuitFromProgram (circuit_parser_qsim.cc:773-826)
// instead of QsimCircuitFromProgram: noise channels become kCH gates; without
// it they are the reference's "Could not parse gate id" error.
Status LowerProgram(const ProgramPB& pb, const SymbolTable& symbols,
                    CircuitT* out, bool allow_channels = false);

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
