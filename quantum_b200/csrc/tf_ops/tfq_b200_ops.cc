// TensorFlow shim: DEVICE_GPU kernels for TFQ's five circuit-execution ops
// (and the two inner-product math ops), forwarding 1:1 to the C ABI in
// include/tfqb.h.
//
// TensorFlow 2.18 is not in this repository's image (INTEGRATION.md has the
// bazel target); tests/test_tf_shim.py compiles this file against a minimal
// stand-in of the op-kernel API (tests/tf_stub/), links it with libtfqb.so and
// runs its Compute() methods.  The op *registrations* (names,
// inputs, outputs, shape functions) stay where they are in the reference:
//   tensorflow_quantum/core/ops/tfq_simulate_expectation_op.cc:257-283
//   tensorflow_quantum/core/ops/tfq_simulate_sampled_expectation_op.cc:312-342
//   tensorflow_quantum/core/ops/tfq_simulate_samples_op.cc:258-285
//   tensorflow_quantum/core/ops/tfq_simulate_state_op.cc:221-242
//   tensorflow_quantum/core/ops/tfq_adj_grad_op.cc:397-427
// This file only adds REGISTER_KERNEL_BUILDER(... DEVICE_GPU ...) entries, so
// tfq.layers.* and tfq.differentiators.Adjoint pick the GPU kernel whenever
// the op is placed on a GPU.  Every input lives in host memory (string
// tensors always do; the float tensors are a few KB), outputs are produced in
// host memory too: the device boundary is inside libtfqb.so.
//
// No simulation logic lives here: rank checks (the reference raises them from
// parse_context.cc:70-73,263-266,301-303,313-316,359-362,396-407,426-429),
// output allocation, and status translation only.
#include <cstdlib>
#include <string>
#include <vector>

#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/lib/random/random.h"
#include "tfqb.h"

namespace tfq_b200 {

using ::tensorflow::OpKernel;
using ::tensorflow::OpKernelConstruction;
using ::tensorflow::OpKernelContext;
using ::tensorflow::Tensor;
using ::tensorflow::tstring;

namespace {

// One context per (process, GPU ordinal), created on first use.  With
// TFQB_TF_ALL_GPUS=<n> the op fans its rows over GPUs 0..n-1 from one Compute
// (tfqb_create_multi), as the reference fans one Compute over all host cores
// (tfq_simulate_expectation_op.cc:245-248).
tfqb_context* ContextFor(OpKernelContext* c) {
  static tensorflow::mutex mu;
  static std::vector<tfqb_context*> by_device(64, nullptr);
  static tfqb_context* all_gpus = nullptr;
  int ordinal = 0;
  if (const auto* info = c->device()->tensorflow_accelerator_device_info())
    ordinal = info->gpu_id;
  tensorflow::mutex_lock l(mu);
  if (const char* n = getenv("TFQB_TF_ALL_GPUS")) {
    const int count = atoi(n);
    if (count > 1) {
      if (!all_gpus) {
        std::vector<int> ids(count);
        for (int i = 0; i < count; ++i) ids[i] = i;
        tfqb_create_multi(ids.data(), count, &all_gpus);
      }
      if (all_gpus) return all_gpus;
    }
  }
  if (!by_device[ordinal]) tfqb_create(ordinal, &by_device[ordinal]);
  return by_device[ordinal];
}

struct Strings {
  std::vector<const char*> data;
  std::vector<size_t> size;
  tfqb_strings c{nullptr, nullptr};
  explicit Strings(const Tensor& t) {
    const auto flat = t.flat<tstring>();
    data.resize(flat.size());
    size.resize(flat.size());
    for (int64_t i = 0; i < flat.size(); ++i) {
      data[i] = flat(i).data();
      size[i] = flat(i).size();
    }
    c.data = data.data();
    c.size = size.data();
  }
};

tensorflow::Status ToStatus(int rc) {
  if (rc == TFQB_OK) return tensorflow::Status();
  const std::string msg = tfqb_last_error();
  switch (rc) {
    case TFQB_INVALID_ARGUMENT: return tensorflow::errors::InvalidArgument(msg);
    case TFQB_RESOURCE_EXHAUSTED: return tensorflow::errors::ResourceExhausted(msg);
    case TFQB_UNAVAILABLE: return tensorflow::errors::Unavailable(msg);
    default: return tensorflow::errors::Internal(msg);
  }
}

#define TFQB_RANK(ctx, idx, want, name)                                        \
  OP_REQUIRES(ctx, ctx->input(idx).dims() == want,                             \
              tensorflow::errors::InvalidArgument(                             \
                  name " must be rank " #want ". Got rank ",                   \
                  ctx->input(idx).dims(), "."))

struct Common {
  Strings programs, names;
  tfqb_circuit_inputs in;
  Common(OpKernelContext* c)
      : programs(c->input(0)), names(c->input(1)) {
    in.programs = programs.c;
    in.batch = static_cast<int>(c->input(0).NumElements());
    in.symbol_names = names.c;
    in.n_symbols = static_cast<int>(c->input(1).NumElements());
    in.symbol_values = c->input(2).flat<float>().data();
    in.symbol_rows = static_cast<int>(c->input(2).dim_size(0));
  }
};

bool CheckCommon(OpKernelContext* c) {
  if (c->input(0).dims() != 1) {
    c->SetStatus(tensorflow::errors::InvalidArgument(
        "programs must be rank 1. Got rank ", c->input(0).dims(), "."));
    return false;
  }
  if (c->input(1).dims() != 1) {
    c->SetStatus(tensorflow::errors::InvalidArgument(
        "symbol_names must be rank 1. Got rank ", c->input(1).dims(), "."));
    return false;
  }
  if (c->input(2).dims() != 2) {
    c->SetStatus(tensorflow::errors::InvalidArgument(
        "symbol_values must be rank 2. Got rank ", c->input(2).dims(), "."));
    return false;
  }
  if (c->input(2).dim_size(1) != c->input(1).dim_size(0)) {
    c->SetStatus(tensorflow::errors::InvalidArgument(
        "Input symbol names and value sizes do not match."));
    return false;
  }
  return true;
}

}  // namespace

class ExpectationGpuOp : public OpKernel {
 public:
  explicit ExpectationGpuOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    if (!CheckCommon(c)) return;
    TFQB_RANK(c, 3, 2, "pauli_sums");
    Common in(c);
    Strings sums(c->input(3));
    const int rows = c->input(3).dim_size(0), cols = c->input(3).dim_size(1);
    Tensor* out = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, {in.in.batch, cols}, &out));
    OP_REQUIRES_OK(c, ToStatus(tfqb_simulate_expectation(
                          ContextFor(c), &in.in, sums.c, rows, cols,
                          out->flat<float>().data())));
  }
};

class SampledExpectationGpuOp : public OpKernel {
 public:
  explicit SampledExpectationGpuOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    if (!CheckCommon(c)) return;
    TFQB_RANK(c, 3, 2, "pauli_sums");
    TFQB_RANK(c, 4, 2, "num_samples");
    Common in(c);
    Strings sums(c->input(3));
    const int rows = c->input(3).dim_size(0), cols = c->input(3).dim_size(1);
    Tensor* out = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, {in.in.batch, cols}, &out));
    // the reference seeds qsim from a non-deterministic Philox
    // (tfq_simulate_sampled_expectation_op.cc:168-169)
    OP_REQUIRES_OK(
        c, ToStatus(tfqb_simulate_sampled_expectation(
               ContextFor(c), &in.in, sums.c, rows, cols,
               c->input(4).flat<int32_t>().data(), c->input(4).dim_size(0),
               c->input(4).dim_size(1), tensorflow::random::New64(), nullptr, 0,
               0, out->flat<float>().data())));
  }
};

class SamplesGpuOp : public OpKernel {
 public:
  explicit SamplesGpuOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    if (!CheckCommon(c)) return;
    TFQB_RANK(c, 3, 1, "num_samples");
    OP_REQUIRES(c, c->input(3).dim_size(0) == 1,
                tensorflow::errors::InvalidArgument(
                    "num_samples must contain 1 element. Got ",
                    c->input(3).dim_size(0), "."));
    Common in(c);
    const int shots = c->input(3).flat<int32_t>()(0);
    tfqb_job* job = nullptr;
    int nmax = 0;
    OP_REQUIRES_OK(c, ToStatus(tfqb_simulate_samples_prepare(
                          ContextFor(c), &in.in, shots, &job, &nmax)));
    Tensor* out = nullptr;
    tensorflow::Status s = c->allocate_output(0, {in.in.batch, shots, nmax}, &out);
    if (s.ok())
      s = ToStatus(tfqb_simulate_samples_run(job, tensorflow::random::New64(),
                                             nullptr, out->flat<int8_t>().data()));
    tfqb_job_free(job);
    OP_REQUIRES_OK(c, s);
  }
};

class StateGpuOp : public OpKernel {
 public:
  explicit StateGpuOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    if (!CheckCommon(c)) return;
    Common in(c);
    tfqb_job* job = nullptr;
    int nmax = 0;
    OP_REQUIRES_OK(c, ToStatus(tfqb_simulate_state_prepare(ContextFor(c), &in.in,
                                                           &job, &nmax)));
    Tensor* out = nullptr;
    tensorflow::Status s =
        c->allocate_output(0, {in.in.batch, int64_t(1) << nmax}, &out);
    if (s.ok())
      s = ToStatus(tfqb_simulate_state_run(
          job, reinterpret_cast<float*>(out->flat<std::complex<float>>().data())));
    tfqb_job_free(job);
    OP_REQUIRES_OK(c, s);
  }
};

class AdjointGradientGpuOp : public OpKernel {
 public:
  explicit AdjointGradientGpuOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    if (!CheckCommon(c)) return;
    TFQB_RANK(c, 3, 2, "pauli_sums");
    TFQB_RANK(c, 4, 2, "downstream_grads");
    Common in(c);
    Strings sums(c->input(3));
    const int rows = c->input(3).dim_size(0), cols = c->input(3).dim_size(1);
    Tensor* out = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, {in.in.batch, in.in.n_symbols}, &out));
    OP_REQUIRES_OK(c, ToStatus(tfqb_adjoint_gradient(
                          ContextFor(c), &in.in, sums.c, rows, cols,
                          c->input(4).flat<float>().data(),
                          c->input(4).dim_size(0), c->input(4).dim_size(1),
                          out->flat<float>().data())));
  }
};

// next-row N1: registrations stay in
//   tensorflow_quantum/core/ops/math_ops/tfq_inner_product.cc:298-325
//   tensorflow_quantum/core/ops/math_ops/tfq_inner_product_grad.cc:460-501
class InnerProductGpuOp : public OpKernel {
 public:
  explicit InnerProductGpuOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    if (!CheckCommon(c)) return;
    TFQB_RANK(c, 3, 2, "other_programs");
    Common in(c);
    Strings others(c->input(3));
    const int rows = c->input(3).dim_size(0), cols = c->input(3).dim_size(1);
    Tensor* out = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, {in.in.batch, cols}, &out));
    OP_REQUIRES_OK(c, ToStatus(tfqb_inner_product(
                          ContextFor(c), &in.in, others.c, rows, cols,
                          reinterpret_cast<float*>(
                              out->flat<std::complex<float>>().data()))));
  }
};

class InnerProductGradGpuOp : public OpKernel {
 public:
  explicit InnerProductGradGpuOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    if (!CheckCommon(c)) return;
    TFQB_RANK(c, 3, 2, "other_programs");
    TFQB_RANK(c, 4, 2, "downstream_grads");
    Common in(c);
    Strings others(c->input(3));
    const int rows = c->input(3).dim_size(0), cols = c->input(3).dim_size(1);
    Tensor* out = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, {in.in.batch, in.in.n_symbols}, &out));
    OP_REQUIRES_OK(c, ToStatus(tfqb_inner_product_grad(
                          ContextFor(c), &in.in, others.c, rows, cols,
                          c->input(4).flat<float>().data(), c->input(4).dim_size(0),
                          c->input(4).dim_size(1),
                          reinterpret_cast<float*>(
                              out->flat<std::complex<float>>().data()))));
  }
};

// next-row N2: registrations stay in
//   tensorflow_quantum/core/ops/noise/tfq_noisy_expectation.cc:394-426
//   tensorflow_quantum/core/ops/noise/tfq_noisy_sampled_expectation.cc:407-440
//   tensorflow_quantum/core/ops/noise/tfq_noisy_samples.cc:324-352
template <bool kSampled>
class NoisyExpectationGpuOp : public OpKernel {
 public:
  explicit NoisyExpectationGpuOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    if (!CheckCommon(c)) return;
    OP_REQUIRES(c, c->input(3).dims() == 2,
                tensorflow::errors::InvalidArgument("pauli_sums must be rank 2. Got ",
                                                    c->input(3).dims()));
    TFQB_RANK(c, 4, 2, "num_samples");
    Common in(c);
    Strings sums(c->input(3));
    const int rows = c->input(3).dim_size(0), cols = c->input(3).dim_size(1);
    Tensor* out = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, {in.in.batch, cols}, &out));
    auto* fn = kSampled ? tfqb_noisy_sampled_expectation : tfqb_noisy_expectation;
    OP_REQUIRES_OK(c, ToStatus(fn(ContextFor(c), &in.in, sums.c, rows, cols,
                                  c->input(4).flat<int32_t>().data(), c->input(4).dim_size(0),
                                  c->input(4).dim_size(1), tensorflow::random::New64(), nullptr,
                                  0, 0, out->flat<float>().data())));
  }
};

class NoisySamplesGpuOp : public OpKernel {
 public:
  explicit NoisySamplesGpuOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    if (!CheckCommon(c)) return;
    TFQB_RANK(c, 3, 1, "num_samples");
    OP_REQUIRES(c, c->input(3).dim_size(0) == 1,
                tensorflow::errors::InvalidArgument(
                    "num_samples must contain 1 element. Got ", c->input(3).dim_size(0), "."));
    Common in(c);
    const int shots = c->input(3).flat<int32_t>()(0);
    tfqb_job* job = nullptr;
    int nmax = 0;
    OP_REQUIRES_OK(c, ToStatus(tfqb_noisy_samples_prepare(ContextFor(c), &in.in, shots, &job, &nmax)));
    Tensor* out = nullptr;
    tensorflow::Status s = c->allocate_output(0, {in.in.batch, shots, nmax}, &out);
    if (s.ok())
      s = ToStatus(tfqb_noisy_samples_run(job, tensorflow::random::New64(), nullptr, 0, nullptr,
                                          out->flat<int8_t>().data()));
    tfqb_job_free(job);
    OP_REQUIRES_OK(c, s);
  }
};

// next-row N4: registration stays in
//   tensorflow_quantum/core/ops/tfq_calculate_unitary_op.cc:147-164
class CalculateUnitaryGpuOp : public OpKernel {
 public:
  explicit CalculateUnitaryGpuOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    if (!CheckCommon(c)) return;
    Common in(c);
    tfqb_job* job = nullptr;
    int nmax = 0;
    OP_REQUIRES_OK(c, ToStatus(tfqb_calculate_unitary_prepare(ContextFor(c), &in.in, &job, &nmax)));
    Tensor* out = nullptr;
    tensorflow::Status s =
        c->allocate_output(0, {in.in.batch, int64_t(1) << nmax, int64_t(1) << nmax}, &out);
    if (s.ok())
      s = ToStatus(tfqb_calculate_unitary_run(
          job, reinterpret_cast<float*>(out->flat<std::complex<float>>().data())));
    tfqb_job_free(job);
    OP_REQUIRES_OK(c, s);
  }
};

#define TFQB_GPU_KERNEL(NAME, CLS, ...)                                    \
  REGISTER_KERNEL_BUILDER(Name(NAME).Device(tensorflow::DEVICE_GPU)       \
                              __VA_ARGS__,                                 \
                          CLS)

TFQB_GPU_KERNEL("TfqSimulateExpectation", ExpectationGpuOp,
                .HostMemory("programs").HostMemory("symbol_names")
                .HostMemory("symbol_values").HostMemory("pauli_sums")
                .HostMemory("expectations"));
TFQB_GPU_KERNEL("TfqSimulateSampledExpectation", SampledExpectationGpuOp,
                .HostMemory("programs").HostMemory("symbol_names")
                .HostMemory("symbol_values").HostMemory("pauli_sums")
                .HostMemory("num_samples").HostMemory("expectations"));
TFQB_GPU_KERNEL("TfqSimulateSamples", SamplesGpuOp,
                .HostMemory("programs").HostMemory("symbol_names")
                .HostMemory("symbol_values").HostMemory("num_samples")
                .HostMemory("samples"));
TFQB_GPU_KERNEL("TfqSimulateState", StateGpuOp,
                .HostMemory("programs").HostMemory("symbol_names")
                .HostMemory("symbol_values").HostMemory("state_vector"));
TFQB_GPU_KERNEL("TfqAdjointGradient", AdjointGradientGpuOp,
                .HostMemory("programs").HostMemory("symbol_names")
                .HostMemory("symbol_values").HostMemory("pauli_sums")
                .HostMemory("downstream_grads").HostMemory("grads"));

TFQB_GPU_KERNEL("TfqNoisyExpectation", NoisyExpectationGpuOp<false>,
                .HostMemory("programs").HostMemory("symbol_names")
                .HostMemory("symbol_values").HostMemory("pauli_sums")
                .HostMemory("num_samples").HostMemory("expectations"));
TFQB_GPU_KERNEL("TfqNoisySampledExpectation", NoisyExpectationGpuOp<true>,
                .HostMemory("programs").HostMemory("symbol_names")
                .HostMemory("symbol_values").HostMemory("pauli_sums")
                .HostMemory("num_samples").HostMemory("expectations"));
TFQB_GPU_KERNEL("TfqNoisySamples", NoisySamplesGpuOp,
                .HostMemory("programs").HostMemory("symbol_names")
                .HostMemory("symbol_values").HostMemory("num_samples")
                .HostMemory("samples"));

TFQB_GPU_KERNEL("TfqCalculateUnitary", CalculateUnitaryGpuOp,
                .HostMemory("programs").HostMemory("symbol_names")
                .HostMemory("symbol_values").HostMemory("unitary"));

TFQB_GPU_KERNEL("TfqInnerProduct", InnerProductGpuOp,
                .HostMemory("programs").HostMemory("symbol_names")
                .HostMemory("symbol_values").HostMemory("other_programs")
                .HostMemory("inner_products"));
TFQB_GPU_KERNEL("TfqInnerProductGrad", InnerProductGradGpuOp,
                .HostMemory("programs").HostMemory("symbol_names")
                .HostMemory("symbol_values").HostMemory("other_programs")
                .HostMemory("downstream_grads").HostMemory("inner_products_grad"));

}  // namespace tfq_b200
