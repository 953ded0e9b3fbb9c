// Protobuf wire-format decoder for tfq.proto.Program / tfq.proto.PauliSum.
//
// Replaces `ParseProto<T>` (reference tensorflow_quantum/core/ops/
// parse_context.cc:41-56), which relies on protoc-generated classes; this
// image has no protoc / libprotobuf headers, and the hot path only needs a
// handful of fields, so the messages are decoded straight into the plain
// structs below (schema: core/proto/program.proto:20-161,
// core/proto/pauli_sum.proto:20-35).  Binary first, text format second, like
// the reference.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace tfqb {

struct ArgPB {
  std::string key;
  // oneof arg { arg_value | symbol | func }; func is ignored like
  // ParseProtoArg does (circuit_parser_qsim.cc:53-82).
  float float_value = 0.f;       // arg_value.float_value (0 when absent)
  std::string string_value;      // arg_value.string_value
  std::string symbol;            // non-empty => symbolic
  // arg_value.bool_values (the "unused" arg of identity gates): never read by
  // the circuit parser, kept so that the program rewrites (ps_ops.cc) write
  // back what they read.  The RepeatedBoolean payload in wire form.
  bool has_bools = false;
  std::string bools_wire;
};

struct OperationPB {
  std::string gate_id;
  std::vector<ArgPB> args;
  std::vector<std::string> qubits;
  const ArgPB* find(const std::string& key) const {
    // map semantics: the last entry with a given key wins.
    const ArgPB* r = nullptr;
    for (const auto& a : args)
      if (a.key == key) r = &a;
    return r;
  }
};

struct MomentPB {
  std::vector<OperationPB> operations;
};

struct ProgramPB {
  std::vector<MomentPB> moments;
};

struct PauliPairPB {
  std::string qubit_id;
  std::string pauli_type;
};

struct PauliTermPB {
  float coefficient_real = 0.f;
  std::vector<PauliPairPB> paulis;
};

struct PauliSumPB {
  std::vector<PauliTermPB> terms;
};

// Return false on malformed input ("Unparseable proto").
bool ParseProgram(const char* data, size_t len, ProgramPB* out);
bool ParsePauliSum(const char* data, size_t len, PauliSumPB* out);

}  // namespace tfqb
