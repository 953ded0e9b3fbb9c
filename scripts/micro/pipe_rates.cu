// Micro-benchmark: issue rates of the instructions the gate pass could be
// built from, on the GPU it runs on (B200): scalar FFMA, packed FFMA2, and the
// legacy warp-level tensor-core path (mma.sync m16n8k8 tf32, m16n8k16 bf16).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a pipe_rates.cu -o pipe_rates
#include <cuda_runtime.h>
#include <cstdio>

constexpr int kIters = 4096;

__global__ void k_ffma(float* out, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float2* out, float2 a, float2 b) {
  float2 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = make_float2(threadIdx.x + i, i);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = __ffma2_rn(x[i], a, b);
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int i = 0; i < 16; ++i) { s.x += x[i].x; s.y += x[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mma_tf32(float* out, unsigned a0, unsigned b0) {
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {b0, b0 + 1};
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile(
          "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 "
          "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
          : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
          : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mma_bf16(float* out, unsigned a0, unsigned b0) {
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {b0, b0 + 1};
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile(
          "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 "
          "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
          : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
          : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) launch();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / 5;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, ctas = sms * 8, thr = 256;
  void* out; cudaMalloc(&out, size_t(ctas) * thr * 16);
  const double warps = double(ctas) * thr / 32;
  float ms = time_ms([&] { k_ffma<<<ctas, thr>>>((float*)out, 1.0001f, 0.5f); });
  printf("{\"kernel\":\"FFMA scalar\",\"tflops\":%.1f,\"warp_instr_per_clk_per_smsp\":%.3f}\n",
         2.0 * 32 * 16 * kIters * warps / ms / 1e9,
         16.0 * kIters * warps / (ms * 1e-3 * p.clockRate * 1e3) / (sms * 4));
  ms = time_ms([&] { k_ffma2<<<ctas, thr>>>((float2*)out, make_float2(1.0001f, 0.999f), make_float2(0.5f, 0.25f)); });
  printf("{\"kernel\":\"FFMA2 packed\",\"tflops\":%.1f,\"warp_instr_per_clk_per_smsp\":%.3f}\n",
         4.0 * 32 * 16 * kIters * warps / ms / 1e9,
         16.0 * kIters * warps / (ms * 1e-3 * p.clockRate * 1e3) / (sms * 4));
  ms = time_ms([&] { k_mma_tf32<<<ctas, thr>>>((float*)out, 0x3f800000u, 0x3f000000u); });
  printf("{\"kernel\":\"mma.sync m16n8k8 tf32\",\"tflops\":%.1f,\"warp_instr_per_clk_per_smsp\":%.3f}\n",
         2.0 * 16 * 8 * 8 * 8 * kIters * warps / ms / 1e9,
         8.0 * kIters * warps / (ms * 1e-3 * p.clockRate * 1e3) / (sms * 4));
  ms = time_ms([&] { k_mma_bf16<<<ctas, thr>>>((float*)out, 0x3f803f80u, 0x3f003f00u); });
  printf("{\"kernel\":\"mma.sync m16n8k16 bf16\",\"tflops\":%.1f,\"warp_instr_per_clk_per_smsp\":%.3f}\n",
         2.0 * 16 * 8 * 16 * 8 * kIters * warps / ms / 1e9,
         8.0 * kIters * warps / (ms * 1e-3 * p.clockRate * 1e3) / (sms * 4));
  printf("{\"sms\":%d,\"clock_mhz\":%d}\n", sms, p.clockRate / 1000);
  return 0;
}
