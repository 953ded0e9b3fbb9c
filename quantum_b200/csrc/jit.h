// Run-time specialised gate-pass kernels.
//
// The interpreted pass kernel (kernels.cu: pass_kernel) spends about a third
// of its issue slots on op fetch / dispatch / address bookkeeping (ncu source
// pages under profiles/).  When a plan is going to be applied to many rows,
// each of its passes is instead emitted as straight-line CUDA C++ — the SAME
// register-level primitives (pass_device.cuh: g1_packed, g2_packed, diag*,
// sign*, adj*), called with every register index, shared-memory offset, sign
// mask and tile position as a literal — compiled for sm_100a with NVRTC and
// launched through the driver API.  Nothing but the unrolling is generated:
// the arithmetic is the hand-written code of pass_device.cuh.
//
// Four generators: gate passes (forward and adjoint), PauliSum expectation
// passes, operator-accumulation passes.  On top of the plain unrolling the
// generator uses what only it can see: gates that are a phase times a real /
// real-diagonal-imaginary-off-diagonal matrix (Y^t, Y^t Z^t, X^t, H) are
// applied with 2-3 packed FMAs per amplitude when the caller cannot observe a
// global phase (GeneratePassSource phase_free), and runs of diagonal adjoint
// steps share their conj(lambda) psi products.
//
// libnvrtc / libcuda are opened with dlopen; when either is missing, or
// TFQB_JIT=0, the interpreted kernel runs (same device code, same results).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "plan.h"

namespace tfqb {

struct JitKernel {
  void* module = nullptr;    // CUmodule
  void* func = nullptr;      // CUfunction
  int threads = 0;
  int tiles = 1;             // tiles per CTA (gate passes)
  size_t smem = 0;
  bool adjoint = false;
};

// True when every op of the pass has a specialised emission (no controlled
// "slow" ops) and the pass uses the full 2^12 tile.
bool PassIsJitable(const DevicePlan& plan, int pass, bool adjoint);

// CUDA C++ source of the specialised kernel for one pass of `plan`
// (forward plans: reg_bits 4, two register groups; adjoint plans: reg_bits 3).
// `device_src` is the text of pass_device.cuh.
// `phase_free`: the caller's result cannot depend on a global phase of the
// state (expectation, sampling, adjoint gradient): "phased real" gates
// (Y^t with Z^t before or after) drop theirs and cost 3 packed FMAs per
// amplitude instead of 4.  Forward plans only.
std::string GeneratePassSource(const DevicePlan& plan, int pass, bool adjoint,
                               bool phase_free = false);

// Packed FP32 instructions (FFMA2 / FMUL2: one per 2 cycles per SM
// sub-partition) the specialised forward kernel of `pass` spends per
// amplitude on gate arithmetic; -1 when the pass is not specialisable.  The
// FP32-pipe floor of a pass is this x amplitudes / the measured FFMA2 rate.
double PassPackedFp32PerAmplitude(const DevicePlan& plan, int pass, bool phase_free);

// Text of pass_device.cuh embedded at build time.
const char* PassDeviceSource();

// NVRTC -> cubin -> module.  Returns false and fills *err on any failure.
// Compilation is asynchronous and memoised by source text (process-wide map,
// plus cubins on disk under TFQB_JIT_CACHE_DIR): JitPrefetch starts compiling
// on a host thread and returns; JitCompile waits for that result (or compiles
// now) and loads the module into the CURRENT device's context.
void JitPrefetch(const std::string& src);
double JitCompileSeconds();   // NVRTC time spent by this process so far
int JitPending();             // compilations running on background threads right now
bool JitAvailable(std::string* why);
bool JitCompile(const std::string& src, const char* entry, bool adjoint, int threads,
                size_t smem, JitKernel* out, std::string* err);
void JitRelease(JitKernel* k);

// grid = (tiles, rows).  Kernel signature (both kinds):
//   (float2* psi, float2* lam, size_t row_stride, const float* mats,
//    size_t mat_row_stride, double* grad_out, int n_slots, int init_mode,
//    unsigned long long rank_base, const float2* const* peer_tab,
//    int peer_shift, unsigned long long peer_self)
// init_mode 3 + peer_*: the tiles are gathered from the peers' shards of a
// sharded state (kernels.cuh PassLaunch).
bool JitLaunch(const JitKernel& k, unsigned tiles, unsigned rows, float2* psi,
               float2* lam, size_t row_stride, const float* mats,
               size_t mat_row_stride, double* grad_out, int n_slots,
               int init_mode, unsigned long long rank_base, cudaStream_t s,
               std::string* err, const float2* const* peer_tab = nullptr,
               int peer_shift = 0, unsigned long long peer_self = 0);

// ---- PauliSum expectation passes (ExpectationPlan), entry "tfqb_jit_expect":
//   (const float2* psi, size_t row_stride, unsigned long long n_tiles,
//    unsigned long long rank_base, double* per_term, int n_terms), grid (ctas, rows)
bool ExpectPassIsJitable(const ExpectationPlan& plan, int pass);
std::string GenerateExpectSource(const ExpectationPlan& plan, int pass);
size_t JitExpectSmem(const ExpectationPlan& plan, int pass);
int JitExpectThreads();
int JitAccumThreads();
bool JitLaunchExpect(const JitKernel& k, unsigned ctas, unsigned rows, const float2* psi,
                     size_t row_stride, unsigned long long n_tiles,
                     unsigned long long rank_base, double* per_term, int n_terms,
                     cudaStream_t s, std::string* err);

// ---- operator accumulation passes (K3) over the same plans, entry
// "tfqb_jit_accum": (const float2* psi, float2* lam, size_t row_stride,
//    const DevTerm* terms, int n_terms, const float* downstream, int n_ops,
//    int accumulate, unsigned long long n_tiles), grid (ctas, rows)
std::string GenerateAccumSource(const ExpectationPlan& plan, int pass);
size_t JitAccumSmem(const ExpectationPlan& plan, int pass, int n_terms);
bool JitLaunchAccum(const JitKernel& k, unsigned ctas, unsigned rows, const float2* psi,
                    float2* lam, size_t row_stride, const void* terms, int n_terms,
                    const float* downstream, int n_ops, int accumulate,
                    unsigned long long n_tiles, cudaStream_t s, std::string* err);

// launch geometry / shared memory of the specialised kernel of a pass
int JitPassThreads(const DevicePlan& plan, bool adjoint);   // per CTA
int JitPassTiles(const DevicePlan& plan, bool adjoint, int pass);     // tiles per CTA
size_t JitPassSmem(const DevicePlan& plan, int pass, bool adjoint);

}  // namespace tfqb
