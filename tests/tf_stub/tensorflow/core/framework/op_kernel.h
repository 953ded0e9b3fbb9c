// Minimal stand-in for the slice of TensorFlow's C++ op-kernel API that
// quantum_b200/csrc/tf_ops/tfq_b200_ops.cc uses (TensorFlow is not in this
// image).  TEST INFRASTRUCTURE: it lets the shim compile, link against
// libtfqb.so and actually run its Compute() methods on host tensors
// (tests/tf_stub/harness.cc), so the DEVICE_GPU registrations are exercised
// end to end; it is not a TensorFlow re-implementation.
#pragma once
#include <complex>
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

namespace tensorflow {

using tstring = std::string;
using mutex = std::mutex;
using mutex_lock = std::lock_guard<std::mutex>;

class Status {
 public:
  Status() = default;
  Status(int code, std::string msg) : code_(code), msg_(std::move(msg)) {}
  bool ok() const { return code_ == 0; }
  int code() const { return code_; }
  const std::string& message() const { return msg_; }

 private:
  int code_ = 0;
  std::string msg_;
};

namespace errors {
template <typename... A>
std::string Cat(const A&... a) {
  std::ostringstream o;
  (o << ... << a);
  return o.str();
}
template <typename... A> Status InvalidArgument(const A&... a) { return Status(3, Cat(a...)); }
template <typename... A> Status ResourceExhausted(const A&... a) { return Status(8, Cat(a...)); }
template <typename... A> Status Internal(const A&... a) { return Status(13, Cat(a...)); }
template <typename... A> Status Unavailable(const A&... a) { return Status(14, Cat(a...)); }
}  // namespace errors

template <typename T>
struct Flat {
  T* p;
  int64_t n;
  T* data() const { return p; }
  int64_t size() const { return n; }
  T& operator()(int64_t i) const { return p[i]; }
};

class Tensor {
 public:
  Tensor() = default;
  template <typename T>
  static Tensor Make(std::vector<int64_t> shape) {
    Tensor t;
    t.shape_ = std::move(shape);
    t.bytes_.resize(size_t(t.NumElements()) * sizeof(T));
    return t;
  }
  static Tensor Strings(std::vector<int64_t> shape, std::vector<tstring> v) {
    Tensor t;
    t.shape_ = std::move(shape);
    t.strings_ = std::move(v);
    return t;
  }
  int dims() const { return int(shape_.size()); }
  int64_t dim_size(int i) const { return shape_[size_t(i)]; }
  int64_t NumElements() const {
    int64_t n = 1;
    for (int64_t d : shape_) n *= d;
    return n;
  }
  template <typename T>
  Flat<T> flat() { return Flat<T>{reinterpret_cast<T*>(bytes_.data()), NumElements()}; }
  template <typename T>
  Flat<const T> flat() const {
    return Flat<const T>{reinterpret_cast<const T*>(bytes_.data()), NumElements()};
  }

 private:
  std::vector<int64_t> shape_;
  std::vector<char> bytes_;
  std::vector<tstring> strings_;
  friend struct StringAccess;
};
struct StringAccess {
  static const std::vector<tstring>& Get(const Tensor& t) { return t.strings_; }
};
template <>
inline Flat<const tstring> Tensor::flat<tstring>() const {
  return Flat<const tstring>{strings_.data(), int64_t(strings_.size())};
}

struct AcceleratorDeviceInfo { int gpu_id = 0; };
class DeviceBase {
 public:
  const AcceleratorDeviceInfo* tensorflow_accelerator_device_info() const { return &info_; }
  AcceleratorDeviceInfo info_;
};

class OpKernelConstruction {};

class OpKernelContext {
 public:
  std::vector<Tensor> inputs;
  std::vector<Tensor> outputs;
  Status status;
  DeviceBase dev;
  const Tensor& input(int i) const { return inputs[size_t(i)]; }
  DeviceBase* device() { return &dev; }
  void SetStatus(const Status& s) { status = s; }
  Status allocate_output(int idx, std::initializer_list<int64_t> shape, Tensor** out) {
    if (size_t(idx) >= outputs.size()) outputs.resize(size_t(idx) + 1);
    std::vector<int64_t> sh(shape);
    int64_t n = 1;
    for (int64_t d : sh) n *= d;
    // the widest element type any output of this shim has is complex64
    outputs[size_t(idx)] = Tensor::Make<std::complex<float>>(sh);
    *out = &outputs[size_t(idx)];
    return Status();
  }
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() = default;
  virtual void Compute(OpKernelContext* c) = 0;
};

constexpr const char* DEVICE_GPU = "GPU";

// registry: op name -> (device, host-memory args, factory)
struct KernelDef {
  std::string name, device;
  std::vector<std::string> host_memory;
  std::function<OpKernel*(OpKernelConstruction*)> factory;
};
inline std::map<std::string, KernelDef>& KernelRegistry() {
  static std::map<std::string, KernelDef> r;
  return r;
}
struct Name {
  KernelDef def;
  explicit Name(const char* n) { def.name = n; }
  Name& Device(const char* d) { def.device = d; return *this; }
  Name& HostMemory(const char* a) { def.host_memory.push_back(a); return *this; }
};
struct KernelRegistrar {
  KernelRegistrar(Name n, std::function<OpKernel*(OpKernelConstruction*)> f) {
    n.def.factory = std::move(f);
    KernelRegistry()[n.def.name] = n.def;
  }
};

}  // namespace tensorflow

using tensorflow::Name;

#define TF_STUB_CAT2(a, b) a##b
#define TF_STUB_CAT(a, b) TF_STUB_CAT2(a, b)
#define REGISTER_KERNEL_BUILDER(builder, cls)                                  \
  static ::tensorflow::KernelRegistrar TF_STUB_CAT(tf_stub_registrar_, __COUNTER__)( \
      builder, [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { \
        return new cls(c);                                                      \
      })
#define OP_REQUIRES(ctx, cond, status) \
  do {                                 \
    if (!(cond)) {                     \
      (ctx)->SetStatus(status);        \
      return;                          \
    }                                  \
  } while (0)
#define OP_REQUIRES_OK(ctx, expr)          \
  do {                                     \
    ::tensorflow::Status s__ = (expr);     \
    if (!s__.ok()) {                       \
      (ctx)->SetStatus(s__);               \
      return;                              \
    }                                      \
  } while (0)
