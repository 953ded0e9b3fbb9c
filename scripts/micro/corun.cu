// Micro-benchmark: can an HBM-bound pass run UNDER an FP32-issue-bound gate pass?
// Kernel A models the dense gate passes of C2 (one 32 KiB tile per CTA, 128
// threads, 16 amplitudes per thread, ~84 packed FMAs per amplitude, read + write,
// 5 CTAs / SM).  Kernel B models the sparse last gate pass and the expectation
// passes (tile in, a few FMAs, optional tile out), written as a PERSISTENT grid
// of `bcta` CTAs per SM on a high-priority stream, so its CTAs stay resident
// next to A's.  Reported: each alone, then together (wall = max).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a corun.cu -o corun
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kTile = 4096;           // amplitudes (float2) per tile

template <int FMAS>
__global__ void __launch_bounds__(128, 5)
kernel_a(float2* __restrict__ psi, float2 m0, float2 m1) {
  extern __shared__ float2 s[];
  float2* g = psi + size_t(blockIdx.x) * kTile;
  for (int i = threadIdx.x; i < kTile / 2; i += 128) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    s[2 * i] = make_float2(v.x, v.y);
    s[2 * i + 1] = make_float2(v.z, v.w);
  }
  __syncthreads();
  float2 a[16];
#pragma unroll 1
  for (int r = 0; r < 8; ++r) {       // 4 rounds x 2 register groups, as the real pass
    const int g = (r & 1) * 2048 + ((threadIdx.x + 17 * r) & 127);   // conflict-free
#pragma unroll
    for (int e = 0; e < 16; ++e) a[e] = s[g + e * 128];
#pragma unroll 1
    for (int k = 0; k < FMAS / 8; ++k) {
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        const float2 x = a[e], y = a[e + 1];
        a[e] = __ffma2_rn(m1, y, __fmul2_rn(m0, x));
        a[e + 1] = __ffma2_rn(m0, y, __fmul2_rn(m1, x));
      }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) s[g + e * 128] = a[e];
    if (r & 1) __syncthreads();
  }
  for (int i = threadIdx.x; i < kTile / 2; i += 128) {
    const float2 p = s[2 * i], q = s[2 * i + 1];
    reinterpret_cast<float4*>(g)[i] = make_float4(p.x, p.y, q.x, q.y);
  }
}

// persistent streaming kernel: grid-stride over tiles
template <bool WRITE>
__global__ void __launch_bounds__(256, 2)
kernel_b(float2* __restrict__ psi, size_t n_tiles, float* __restrict__ out) {
  extern __shared__ float2 s[];
  float acc = 0.f;
  for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    float2* g = psi + t * kTile;
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(g) + u * 256 + threadIdx.x);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc = fmaf(v[u].x, v[u].x, fmaf(v[u].y, v[u].y, fmaf(v[u].z, v[u].z, fmaf(v[u].w, v[u].w, acc))));
      if (WRITE) {
        v[u].x = -v[u].x;
        __stcs(reinterpret_cast<float4*>(g) + u * 256 + threadIdx.x, v[u]);
      }
    }
  }
  s[threadIdx.x].x = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 256; ++i) t += s[i].x;
    atomicAdd(out, t);
  }
}

int main(int argc, char** argv) {
  const int rows = argc > 1 ? atoi(argv[1]) : 1024;       // 20-qubit states
  const size_t n_tiles = size_t(rows) * 256;
  const size_t bytes = n_tiles * kTile * sizeof(float2);
  float2 *pa, *pb;
  float* out;
  CK(cudaMalloc(&pa, bytes));
  CK(cudaMalloc(&pb, bytes));
  CK(cudaMalloc(&out, 4));
  CK(cudaMemset(pa, 0, bytes));
  CK(cudaMemset(pb, 0, bytes));
  int lo, hi;
  CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  cudaStream_t sa, sb;
  CK(cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, lo));
  CK(cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, hi));
  constexpr int F = 84;
  CK(cudaFuncSetAttribute(kernel_a<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024));
  cudaEvent_t e0, e1, e2, e3;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2)); CK(cudaEventCreate(&e3));
  const float2 m0 = make_float2(0.8f, 0.8f), m1 = make_float2(0.6f, -0.6f);
  auto run_a = [&]() { kernel_a<F><<<unsigned(n_tiles), 128, 34 * 1024, sa>>>(pa, m0, m1); };
  auto run_b = [&](int bcta, bool write, int reps) {
    for (int r = 0; r < reps; ++r) {
      if (write) kernel_b<true><<<148 * bcta, 256, 33 * 1024, sb>>>(pb, n_tiles, out);
      else kernel_b<false><<<148 * bcta, 256, 33 * 1024, sb>>>(pb, n_tiles, out);
    }
  };
  float ms;
  // A alone (2 passes of the batch, like C2's passes 0/1)
  for (int w = 0; w < 2; ++w) {
    CK(cudaEventRecord(e0, sa)); run_a(); run_a(); CK(cudaEventRecord(e1, sa));
    CK(cudaStreamSynchronize(sa));
  }
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double gb = double(bytes) / 1e9;
  printf("{\"what\":\"A alone (2 dense passes)\",\"rows\":%d,\"ms\":%.3f,\"GBps\":%.0f}\n", rows, ms,
         4 * gb / ms * 1e3);
  const float ms_a = ms;
  for (int bcta : {1, 2, 3, 4, 32}) {
    // B = one read+write pass and two read passes of the same amount of state
    for (int w = 0; w < 2; ++w) {
      CK(cudaEventRecord(e2, sb)); run_b(bcta, true, 1); run_b(bcta, false, 2);
      CK(cudaEventRecord(e3, sb)); CK(cudaStreamSynchronize(sb));
    }
    CK(cudaEventElapsedTime(&ms, e2, e3));
    const float ms_b = ms;
    // together
    float ms_a2 = 0, ms_b2 = 0, wall = 0;
    for (int w = 0; w < 2; ++w) {
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0, sa)); CK(cudaEventRecord(e2, sb));
      run_a(); run_b(bcta, true, 1); run_a(); run_b(bcta, false, 2);
      CK(cudaEventRecord(e1, sa)); CK(cudaEventRecord(e3, sb));
      CK(cudaDeviceSynchronize());
      CK(cudaEventElapsedTime(&ms_a2, e0, e1));
      CK(cudaEventElapsedTime(&ms_b2, e2, e3));
      float x, y;
      CK(cudaEventElapsedTime(&x, e0, e3));
      CK(cudaEventElapsedTime(&y, e2, e1));
      wall = fmaxf(fmaxf(ms_a2, ms_b2), fmaxf(x, y));
    }
    printf("{\"what\":\"B persistent, %d CTA/SM x 256 thr\",\"B_alone_ms\":%.3f,\"B_alone_GBps\":%.0f,"
           "\"together_A_ms\":%.3f,\"together_B_ms\":%.3f,\"together_wall_ms\":%.3f,"
           "\"serial_sum_ms\":%.3f,\"speedup_vs_serial\":%.3f}\n",
           bcta, ms_b, 4 * gb / ms_b * 1e3, ms_a2, ms_b2, wall, ms_a + ms_b, (ms_a + ms_b) / wall);
  }
  return 0;
}
