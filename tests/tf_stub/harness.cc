// Runs the TF shim's DEVICE_GPU kernels (quantum_b200/csrc/tf_ops/
// tfq_b200_ops.cc) on host tensors through the stub op-kernel API of this
// directory.  Input file (written by tests/test_tf_shim.py):
//   line 1: B P M            (batch, symbols, PauliSums per row)
//   then B programs, P symbol names, B*P floats, B*M PauliSums, B*M floats
//   (downstream grads); every string as "<len>\n<bytes>\n".
// Output: the registry, then "TfqSimulateExpectation" values, then
// "TfqAdjointGradient" values, one line each.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>

#include "tensorflow/core/framework/op_kernel.h"

using namespace tensorflow;

static std::string ReadString(std::istream& in) {
  size_t n = 0;
  in >> n;
  in.get();
  std::string s(n, '\0');
  in.read(&s[0], std::streamsize(n));
  in.get();
  return s;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  int B, P, M;
  in >> B >> P >> M;
  std::vector<tstring> progs(B), names(P), sums(size_t(B) * M);
  for (auto& s : progs) s = ReadString(in);
  for (auto& s : names) s = ReadString(in);
  Tensor vals = Tensor::Make<float>({B, P});
  for (int i = 0; i < B * P; ++i) in >> vals.flat<float>()(i);
  for (auto& s : sums) s = ReadString(in);
  Tensor down = Tensor::Make<float>({B, M});
  for (int i = 0; i < B * M; ++i) in >> down.flat<float>()(i);

  std::cout << "registered";
  for (auto& kv : KernelRegistry())
    std::cout << " " << kv.first << ":" << kv.second.device << ":" << kv.second.host_memory.size();
  std::cout << "\n";

  OpKernelConstruction cons;
  for (const char* op : {"TfqSimulateExpectation", "TfqAdjointGradient"}) {
    auto it = KernelRegistry().find(op);
    if (it == KernelRegistry().end()) return 3;
    std::unique_ptr<OpKernel> k(it->second.factory(&cons));
    OpKernelContext c;
    c.inputs = {Tensor::Strings({B}, progs), Tensor::Strings({P}, names), vals,
                Tensor::Strings({B, M}, sums)};
    if (std::string(op) == "TfqAdjointGradient") c.inputs.push_back(down);
    k->Compute(&c);
    if (!c.status.ok()) {
      std::cout << op << " ERROR " << c.status.code() << " " << c.status.message() << "\n";
      continue;
    }
    std::cout << op;
    const Tensor& o = c.outputs[0];
    const auto f = o.flat<float>();
    for (int64_t i = 0; i < o.NumElements(); ++i) printf(" %.9g", double(f(i)));
    std::cout << "\n";
  }
  // an error path: symbol_values of the wrong rank
  {
    std::unique_ptr<OpKernel> k(KernelRegistry()["TfqSimulateExpectation"].factory(&cons));
    OpKernelContext c;
    c.inputs = {Tensor::Strings({B}, progs), Tensor::Strings({P}, names),
                Tensor::Make<float>({B * P}), Tensor::Strings({B, M}, sums)};
    k->Compute(&c);
    std::cout << "rank_error " << c.status.code() << " " << c.status.message() << "\n";
  }
  return 0;
}
