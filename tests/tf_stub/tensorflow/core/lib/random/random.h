// stand-in for tensorflow/core/lib/random/random.h (see op_kernel.h here)
#pragma once
#include <cstdint>
#include <random>
namespace tensorflow {
namespace random {
inline uint64_t New64() {
  static std::mt19937_64 g{std::random_device{}()};
  return g();
}
}  // namespace random
}  // namespace tensorflow
