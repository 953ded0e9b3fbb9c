// Parameter-shift helper ops on decoded programs (ps_ops.cc).
#pragma once
#include <string>
#include <vector>

#include "program.h"
#include "wire.h"

namespace tfqb {

// tfq.proto.Program bytes of a decoded program (what TFQ's serializer writes).
std::string EncodeProgram(const ProgramPB& p);
// Program{language{gate_set: "tfq_gate_set"}, circuit{}}: the padding entry of
// TfqPsSymbolReplace (tfq_ps_symbol_replace_op.cc:134-139).
std::string EmptyProgram();

// TfqPsDecompose: parameterised ISP / PXP / FSIM / PISP operations become
// XXP / YYP / ZP / XP / CZP operations in extra moments.
Status PsDecompose(const ProgramPB& in, ProgramPB* out);
// TfqPsSymbolReplace: one serialized copy of the program per occurrence of
// `symbol`, with that occurrence renamed to `replacement`.
void PsSymbolReplace(const ProgramPB& in, const std::string& symbol,
                     const std::string& replacement, std::vector<std::string>* out);
// TfqPsWeightsFromSymbols: per symbol, the exponent_scalar of every operation
// whose exponent is that symbol.
Status PsWeightsFromSymbols(const ProgramPB& in, const std::vector<std::string>& symbols,
                            std::vector<std::vector<float>>* out);

}  // namespace tfqb
