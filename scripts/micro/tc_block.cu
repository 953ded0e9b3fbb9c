// Stand-alone check of the tcgen05 building block for fused 4-qubit gate
// blocks: D[128 x 32] = A[128 x 32] * B[32 x 32]^T in 3xTF32, A (the threads'
// amplitude groups) written to TMEM with tcgen05.st, B (the fused real 32x32
// gate matrix) in shared memory, canonical K-major SWIZZLE_128B layout.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tc_block.cu -o tc_block
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4);        // start address
  d |= uint64_t(0) << 16;                       // LBO (unused, swizzled K-major)
  d |= uint64_t(1024 >> 4) << 32;               // SBO: 8 rows * 128 B
  d |= uint64_t(1) << 46;                       // descriptor version (sm_100)
  d |= uint64_t(2) << 61;                       // SWIZZLE_128B
  return d;
}
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) |
                            ((128u >> 4) << 24);

__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
      :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(kIdesc), "r"(accumulate),
         "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}

#define TMEM_ST32(addr, v)                                                         \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
               :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), \
                  "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), \
                  "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), \
                  "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory")
#define TMEM_LD32(addr, v)                                                         \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
               : "r"(addr) : "memory")

template <int MODE>
__global__ void __launch_bounds__(128) tc_test(const float* X, const float* R, float* out,
                                              int reps) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sB_hi = reinterpret_cast<float*>(smem);
  float* sB_lo = sB_hi + 32 * 32;
  float* sB_l2 = sB_lo + 32 * 32;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < 1024; idx += 128) {
    const int n = idx >> 5, k = idx & 31;
    const float v = R[n * 32 + k];
    float hi, lo, l2 = 0.f;
    if (MODE == 1) {
      uint32_t t; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v)); hi = __uint_as_float(t);
      lo = v - hi;
    } else {
      hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
      lo = v - hi;
      if (MODE == 2) { const float m = __uint_as_float(__float_as_uint(lo) & 0xffffe000u); l2 = lo - m; lo = m; }
    }
    const int off = n * 32 + ((((k >> 2) ^ (n & 7)) << 2) | (k & 3));
    sB_hi[off] = hi;
    sB_lo[off] = lo;
    sB_l2[off] = l2;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&tmem_base_s)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s;
  const uint32_t taddr = base + (uint32_t(warp * 32) << 16);
  uint32_t phase = 0;
  uint32_t r[32];
  for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(X[tid * 32 + k]);
  for (int rep = 0; rep < reps; ++rep) {
    uint32_t hi[32], lo[32], l2[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (MODE == 1) {
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi[k]) : "f"(__uint_as_float(r[k])));
        lo[k] = __float_as_uint(__uint_as_float(r[k]) - __uint_as_float(hi[k]));
      } else {
        hi[k] = r[k] & 0xffffe000u;
        const float l = __uint_as_float(r[k]) - __uint_as_float(hi[k]);
        lo[k] = __float_as_uint(l);
        if (MODE == 2) {
          lo[k] = __float_as_uint(l) & 0xffffe000u;
          l2[k] = __float_as_uint(l - __uint_as_float(lo[k]));
        }
      }
    }
    TMEM_ST32(taddr + 0, hi);
    TMEM_ST32(taddr + 32, lo);
    if (MODE == 2) TMEM_ST32(taddr + 96, l2);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0 && lane == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t bh = smem_u32(sB_hi), bl = smem_u32(sB_lo), b2 = smem_u32(sB_l2);
      // small terms first so the fp32 accumulator adds them before the big one
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t dh = make_b_desc(bh + kk * 32), dl = make_b_desc(bl + kk * 32);
        const uint64_t d2 = make_b_desc(b2 + kk * 32);
        uint32_t acc = kk > 0 ? 1u : 0u;
        if (MODE == 2) {
          mma_ts(base + 64, base + 96 + kk * 8, dh, acc); acc = 1u;   // l*h
          mma_ts(base + 64, base + 0 + kk * 8, d2, 1u);               // h*l
          mma_ts(base + 64, base + 32 + kk * 8, dl, 1u);              // m*m
        }
        mma_ts(base + 64, base + 0 + kk * 8, dl, acc);
        mma_ts(base + 64, base + 32 + kk * 8, dh, 1u);
        mma_ts(base + 64, base + 0 + kk * 8, dh, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                   :: "r"(smem_u32(&bar)) : "memory");
    }
    {   // wait for the MMAs
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
      }
      phase ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    TMEM_LD32(taddr + 64, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  for (int k = 0; k < 32; ++k) out[(blockIdx.x * 128 + tid) * 32 + k] = __uint_as_float(r[k]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(128u) : "memory");
}

int main() {
  std::vector<float> X(128 * 32), R(32 * 32);
  srand(1);
  for (auto& v : X) v = float(rand()) / RAND_MAX - 0.5f;
  for (auto& v : R) v = float(rand()) / RAND_MAX - 0.5f;
  float *dX, *dR, *dO;
  const int blocks = 148 * 4;
  cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dR, R.size() * 4);
  cudaMalloc(&dO, size_t(blocks) * 128 * 32 * 4);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dR, R.data(), R.size() * 4, cudaMemcpyHostToDevice);
  for (int mode = 0; mode < 3; ++mode) {
    auto launch = [&](int nb, int reps) {
      if (mode == 0) tc_test<0><<<nb, 128, 12288 + 1024>>>(dX, dR, dO, reps);
      else if (mode == 1) tc_test<1><<<nb, 128, 12288 + 1024>>>(dX, dR, dO, reps);
      else tc_test<2><<<nb, 128, 12288 + 1024>>>(dX, dR, dO, reps);
    };
    launch(1, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"mode\":%d,\"launch\":\"%s\"}\n", mode, cudaGetErrorString(e)); return 1; }
    std::vector<float> O(128 * 32);
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0, fp32err = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 32; ++n) {
        double ref = 0; float f = 0.f;
        for (int k = 0; k < 32; ++k) { ref += double(X[m * 32 + k]) * double(R[n * 32 + k]); f = fmaf(X[m * 32 + k], R[n * 32 + k], f); }
        maxerr = fmax(maxerr, fabs(ref - O[m * 32 + n]));
        fp32err = fmax(fp32err, fabs(ref - f));
        maxref = fmax(maxref, fabs(ref));
      }
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int reps = 2000;
    launch(blocks, 10);
    cudaEventRecord(a);
    launch(blocks, reps);
    cudaEventRecord(b); e = cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("{\"mode\":%d,\"max_abs_err\":%.3e,\"fp32_fma_err\":%.3e,\"max_ref\":%.3f,\"us_per_block_per_sm\":%.3f,\"status\":\"%s\"}\n",
           mode, maxerr, fp32err, maxref, ms * 1e3 / reps / (blocks / 148.0), cudaGetErrorString(e));
  }
  return 0;
}
