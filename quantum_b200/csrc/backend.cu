// C ABI of the B200 state-vector backend (include/tfqb.h).
//
// Each entry point is the body of one reference OpKernel::Compute:
//   parse + resolve + lower (wire.cc, program.cc)  -- once per DISTINCT
//   program string, not once per row as QsimCircuitFromProgram does
//   (tfq_simulate_expectation_op.cc:94-108);
//   plan passes (plan.cc); then per group of rows that share a program (and
//   PauliSums) stream memory-sized chunks of states through the kernels of
//   kernels.cu on the context stream.
// There is no CPU fallback: without a CUDA device tfqb_create fails.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/tfqb.h"
#include "gates.cuh"
#include "jit.h"
#include "kernels.cuh"
#include "plan.h"
#include "ps_ops.h"
#include "program.h"
#include "wire.h"

using namespace tfqb;

namespace {

thread_local std::string g_last_error;

int Fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define TFQB_CUDA(expr)                                                     \
  do {                                                                      \
    cudaError_t e__ = (expr);                                               \
    if (e__ != cudaSuccess)                                                 \
      return Fail(e__ == cudaErrorMemoryAllocation ? TFQB_RESOURCE_EXHAUSTED \
                                                   : TFQB_INTERNAL,         \
                  std::string("CUDA error: ") + cudaGetErrorString(e__) +   \
                      " at " #expr);                                        \
  } while (0)

#define TFQB_RETURN_IF(expr)   \
  do {                         \
    int rc__ = (expr);         \
    if (rc__ != TFQB_OK) return rc__; \
  } while (0)

constexpr int kMaxDeviceQubits = 40;   // 8 TiB of amplitudes: beyond any device

// NVTX ranges (the reference has TF's profiler
// annotations; here Nsight Systems / ncu --nvtx see the phases of an op call).
// Header-only NVTX3: a no-op unless a profiler injects itself.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

struct DevPlan {           // device copy of a DevicePlan
  void* blob = nullptr;
  PassRec* passes = nullptr;
  RoundRec* rounds = nullptr;
  OpRec* ops = nullptr;
  MatRec* mats = nullptr;
  FactorRec* factors = nullptr;
  ~DevPlan() { if (blob) cudaFree(blob); }
};

struct CompiledPlan {
  DevicePlan host;
  DevPlan dev;
  // run-time specialised pass kernels (jit.h), built once the plan has been
  // applied to enough amplitudes to pay for the compilation
  mutable std::mutex jit_mu;
  mutable std::vector<JitKernel> jit;
  mutable std::vector<std::string> jit_src;   // generated once, compiled asynchronously
  mutable std::vector<int> jit_state;     // 0 untried, 1 ready, -1 not possible
  mutable double jit_work = 0.0;          // amplitudes this plan was run over
  mutable int jit_calls = 0;
  ~CompiledPlan() {
    for (JitKernel& k : jit) JitRelease(&k);
  }
};

// run-time specialised expectation kernels of one ExpectationPlan (jit.h);
// plans are rebuilt per call, so these live in the context keyed by the
// plan's content
struct ExpJitEntry {
  std::vector<JitKernel> k;
  std::vector<std::string> src, src_acc;   // generated once, compiled asynchronously
  std::vector<int> state;      // 0 untried, 1 ready, -1 not possible
  double work = 0.0;           // amplitudes the plan was evaluated over
  int calls = 0, calls_acc = 0;
  std::vector<JitKernel> k_acc;   // operator-accumulation kernels of the same plan
  std::vector<int> state_acc;
  double work_acc = 0.0;
  ~ExpJitEntry() {
    for (JitKernel& x : k) JitRelease(&x);
    for (JitKernel& x : k_acc) JitRelease(&x);
  }
};

struct CompiledExpPlan {     // device copy of an ExpectationPlan
  std::shared_ptr<ExpJitEntry> jit;
  ExpectationPlan host;
  void* blob = nullptr;
  PassRec* passes = nullptr;
  RoundRec* rounds = nullptr;
  ExpXOp* xops = nullptr;
  ExpZTerm* zterms = nullptr;
  int32_t* generic = nullptr;
  // the blob comes from the context's caching allocator (plans are rebuilt
  // per call: cudaMalloc / cudaFree would synchronise the device every time)
  void (*release)(void* owner, void* p) = nullptr;
  void* owner = nullptr;
  ~CompiledExpPlan() {
    if (blob && release) release(owner, blob);
    else if (blob) cudaFree(blob);
  }
};

// A noisy program cut at its non-unitary channels (next-row N2): segment k
// starts with the channel whose Kraus operator depends on the population
// measured right before it.
struct NoisyPlan {
  std::vector<std::unique_ptr<CompiledPlan>> segs;
  struct Measure { int bit = -1, col = -1; };
  std::vector<Measure> before;       // before segment k (bit < 0: nothing)
  size_t mat_floats = 64;            // widest matrix block of any segment
};

struct CompiledProgram {
  CircuitT circuit;
  std::unique_ptr<CompiledPlan> fwd, adj;
  std::unique_ptr<CompiledPlan> fwd_any;   // forward plan for an arbitrary input state
  std::unique_ptr<NoisyPlan> noisy;
  // gate segments of sharded-state plans, keyed by (world, rank, Pauli terms):
  // kept so that a repeated sharded evaluation re-uses its specialised kernels
  std::map<std::string, std::vector<std::shared_ptr<CompiledPlan>>> sharded_gates;
};

struct TimedLaunch {
  cudaEvent_t a, b;
  int kind;      // 0 forward pass, 1 adjoint pass, 2 expectation
  double bytes;
};

}  // namespace

struct tfqb_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  size_t budget = 0;
  std::mutex mu;
  // multi-device context (tfqb_create_multi): one child per GPU; the parent
  // owns no device state and fans the rows of every op call over the children
  std::vector<tfqb_context*> children;
  // global index of this context's row 0: the Philox streams of the sampling
  // ops are keyed by GLOBAL row, so a batch split over devices (or ranks)
  // draws the same uniforms as the unsplit batch
  int64_t row_offset = 0;
  // lower bound for a job's max_qubits (the [batch, ..., max_qubits] outputs
  // of a split batch share one padded width)
  int force_nmax = 0;
  // caching allocator for big device buffers
  struct Block { void* p; size_t cap; };
  std::vector<Block> free_blocks;
  std::unordered_map<void*, size_t> live;
  // compiled-program cache (key: symbol names + program bytes)
  std::unordered_map<std::string, std::shared_ptr<CompiledProgram>> cache;
  size_t cache_bytes = 0;
  std::unordered_map<std::string, std::shared_ptr<ExpJitEntry>> exp_jit;
  // profile
  tfqb_profile prof{};
  bool prof_timing = false;
  std::vector<TimedLaunch> timed;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> event_pool;

  int Alloc(size_t bytes, void** out) {
    if (bytes == 0) bytes = 256;
    int best = -1;
    for (size_t i = 0; i < free_blocks.size(); ++i)
      if (free_blocks[i].cap >= bytes && free_blocks[i].cap <= 2 * bytes + (1 << 20) &&
          (best < 0 || free_blocks[i].cap < free_blocks[best].cap))
        best = int(i);
    if (best >= 0) {
      *out = free_blocks[best].p;
      live[*out] = free_blocks[best].cap;
      free_blocks.erase(free_blocks.begin() + best);
      return TFQB_OK;
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      Trim();
      e = cudaMalloc(out, bytes);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      return Fail(TFQB_RESOURCE_EXHAUSTED,
                  "Out of device memory allocating " + std::to_string(bytes) +
                      " bytes for state vectors.");
    }
    live[*out] = bytes;
    return TFQB_OK;
  }
  void Release(void* p) {
    if (!p) return;
    auto it = live.find(p);
    if (it == live.end()) return;
    free_blocks.push_back(Block{p, it->second});
    live.erase(it);
  }
  void Trim() {
    cudaStreamSynchronize(stream);
    for (auto& b : free_blocks) cudaFree(b.p);
    free_blocks.clear();
  }
};

namespace {

template <typename T>
int AllocT(tfqb_context* ctx, size_t count, T** out) {
  void* p = nullptr;
  TFQB_RETURN_IF(ctx->Alloc(count * sizeof(T), &p));
  *out = static_cast<T*>(p);
  return TFQB_OK;
}

int UploadPlan(tfqb_context* ctx, const DevicePlan& hp, DevPlan* dp) {
  auto al = [](size_t v) { return (v + 255) & ~size_t(255); };
  const size_t b0 = al(hp.passes.size() * sizeof(PassRec));
  const size_t b1 = al(hp.rounds.size() * sizeof(RoundRec));
  const size_t b2 = al(hp.ops.size() * sizeof(OpRec));
  const size_t b3 = al(hp.mats.size() * sizeof(MatRec));
  const size_t b4 = al(hp.factors.size() * sizeof(FactorRec));
  std::vector<char> host(b0 + b1 + b2 + b3 + b4 + 256, 0);
  if (!hp.passes.empty()) memcpy(host.data(), hp.passes.data(), hp.passes.size() * sizeof(PassRec));
  if (!hp.rounds.empty()) memcpy(host.data() + b0, hp.rounds.data(), hp.rounds.size() * sizeof(RoundRec));
  if (!hp.ops.empty()) memcpy(host.data() + b0 + b1, hp.ops.data(), hp.ops.size() * sizeof(OpRec));
  if (!hp.mats.empty()) memcpy(host.data() + b0 + b1 + b2, hp.mats.data(), hp.mats.size() * sizeof(MatRec));
  if (!hp.factors.empty()) memcpy(host.data() + b0 + b1 + b2 + b3, hp.factors.data(), hp.factors.size() * sizeof(FactorRec));
  TFQB_CUDA(cudaMalloc(&dp->blob, host.size()));
  TFQB_CUDA(cudaMemcpyAsync(dp->blob, host.data(), host.size(),
                            cudaMemcpyHostToDevice, ctx->stream));
  TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
  char* base = static_cast<char*>(dp->blob);
  dp->passes = reinterpret_cast<PassRec*>(base);
  dp->rounds = reinterpret_cast<RoundRec*>(base + b0);
  dp->ops = reinterpret_cast<OpRec*>(base + b0 + b1);
  dp->mats = reinterpret_cast<MatRec*>(base + b0 + b1 + b2);
  dp->factors = reinterpret_cast<FactorRec*>(base + b0 + b1 + b2 + b3);
  ctx->prof.h2d_bytes += int64_t(host.size());
  return TFQB_OK;
}

int BeginTimed(tfqb_context* ctx, int kind, double bytes);
void EndTimed(tfqb_context* ctx, int h);

int CompileExpPlan(tfqb_context* ctx, ExpectationPlan&& hp,
                   std::unique_ptr<CompiledExpPlan>* out) {
  auto cp = std::make_unique<CompiledExpPlan>();
  cp->host = std::move(hp);
  const ExpectationPlan& h = cp->host;
  auto al = [](size_t v) { return (v + 255) & ~size_t(255); };
  const size_t b0 = al(h.passes.size() * sizeof(PassRec));
  const size_t b1 = al(h.rounds.size() * sizeof(RoundRec));
  const size_t b2 = al(h.xops.size() * sizeof(ExpXOp));
  const size_t b3 = al(h.zterms.size() * sizeof(ExpZTerm));
  const size_t b4 = al(h.generic_terms.size() * sizeof(int32_t));
  std::vector<char> host(b0 + b1 + b2 + b3 + b4 + 256, 0);
  if (!h.passes.empty()) memcpy(host.data(), h.passes.data(), h.passes.size() * sizeof(PassRec));
  if (!h.rounds.empty()) memcpy(host.data() + b0, h.rounds.data(), h.rounds.size() * sizeof(RoundRec));
  if (!h.xops.empty()) memcpy(host.data() + b0 + b1, h.xops.data(), h.xops.size() * sizeof(ExpXOp));
  if (!h.zterms.empty()) memcpy(host.data() + b0 + b1 + b2, h.zterms.data(), h.zterms.size() * sizeof(ExpZTerm));
  if (!h.generic_terms.empty()) memcpy(host.data() + b0 + b1 + b2 + b3, h.generic_terms.data(), h.generic_terms.size() * sizeof(int32_t));
  {
    std::string key(host.data(), host.size());
    key += "|" + std::to_string(h.n_alloc);
    if (ctx->exp_jit.size() > 256) ctx->exp_jit.clear();
    auto& slot = ctx->exp_jit[key];
    if (!slot) slot = std::make_shared<ExpJitEntry>();
    cp->jit = slot;
  }
  TFQB_RETURN_IF(ctx->Alloc(host.size(), &cp->blob));
  cp->owner = ctx;
  cp->release = [](void* owner, void* p) {
    tfqb_context* c = static_cast<tfqb_context*>(owner);
    cudaStreamSynchronize(c->stream);     // kernels may still read the plan
    c->Release(p);
  };
  TFQB_CUDA(cudaMemcpyAsync(cp->blob, host.data(), host.size(),
                            cudaMemcpyHostToDevice, ctx->stream));
  TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
  char* base = static_cast<char*>(cp->blob);
  cp->passes = reinterpret_cast<PassRec*>(base);
  cp->rounds = reinterpret_cast<RoundRec*>(base + b0);
  cp->xops = reinterpret_cast<ExpXOp*>(base + b0 + b1);
  cp->zterms = reinterpret_cast<ExpZTerm*>(base + b0 + b1 + b2);
  cp->generic = reinterpret_cast<int32_t*>(base + b0 + b1 + b2 + b3);
  ctx->prof.h2d_bytes += int64_t(host.size());
  *out = std::move(cp);
  return TFQB_OK;
}

// When to pay for a run-time compilation (a few seconds per pass): the plan
// has covered TFQB_JIT_MIN_AMPS amplitudes (default 2^27) AND it is being
// re-used -- a batch of >= 64 rows, or a second call with the same plan.  A
// one-off evaluation of a single large state stays on the interpreted kernel.
// TFQB_JIT_MIN_AMPS=0 compiles unconditionally (tests).
static bool JitWorthIt(double work, int rows, int calls) {
  const char* env_min = getenv("TFQB_JIT_MIN_AMPS");
  if (env_min && *env_min) {
    const double v = atof(env_min);
    if (v <= 0.0) return true;
    return work >= v && (rows >= 64 || calls >= 2);
  }
  return work >= double(1ull << 27) && (rows >= 64 || calls >= 2);
}

// Generate the source of every pass of an expectation (or accumulation) plan
// and start compiling all of them on host threads (jit.h JitPrefetch).
static void PrefetchExpJit(const CompiledExpPlan& ep, bool accum) {
  if (!ep.jit) return;
  std::string why;
  if (!JitAvailable(&why)) return;
  ExpJitEntry& e = *ep.jit;
  const size_t np = ep.host.passes.size();
  std::vector<std::string>& src = accum ? e.src_acc : e.src;
  if (src.size() == np) return;
  src.assign(np, std::string());
  for (size_t q = 0; q < np; ++q) {
    if (!ExpectPassIsJitable(ep.host, int(q))) continue;
    src[q] = accum ? GenerateAccumSource(ep.host, int(q)) : GenerateExpectSource(ep.host, int(q));
    JitPrefetch(src[q]);
  }
}

static const JitKernel* ExpJitKernelFor(tfqb_context* ctx, const CompiledExpPlan& ep,
                                        int p, double amps, int rows, bool account = true) {
  if (!ep.jit) return nullptr;
  ExpJitEntry& e = *ep.jit;
  const size_t np = ep.host.passes.size();
  if (e.state.size() != np) {
    e.state.assign(np, 0);
    e.k.assign(np, JitKernel());
  }
  if (p == 0 && account) {
    e.work += amps;
    e.calls++;
  }
  if (e.state[p] == 1) return &e.k[p];
  // !account: a look-up only; a sharded job compiles before its first wait
  if (e.state[p] < 0 || !account || !JitWorthIt(e.work, rows, e.calls)) return nullptr;
  std::string why;
  if (!JitAvailable(&why)) return nullptr;
  PrefetchExpJit(ep, false);
  e.state[p] = -1;
  if (!ExpectPassIsJitable(ep.host, p)) return nullptr;
  const std::string& src = e.src[p];
  std::string err;
  if (!JitCompile(src, "tfqb_jit_expect", false, JitExpectThreads(),
                  JitExpectSmem(ep.host, p), &e.k[p], &err)) {
    if (getenv("TFQB_JIT_VERBOSE")) fprintf(stderr, "tfqb jit: %s\n", err.c_str());
    return nullptr;
  }
  ctx->prof.jit_kernels++;
  e.state[p] = 1;
  return &e.k[p];
}

static const JitKernel* AccumJitKernelFor(tfqb_context* ctx, const CompiledExpPlan& ep,
                                          int p, double amps, int n_terms, int rows) {
  if (!ep.jit) return nullptr;
  ExpJitEntry& e = *ep.jit;
  const size_t np = ep.host.passes.size();
  if (e.state_acc.size() != np) {
    e.state_acc.assign(np, 0);
    e.k_acc.assign(np, JitKernel());
  }
  if (p == 0) {
    e.work_acc += amps;
    e.calls_acc++;
  }
  const size_t smem = JitAccumSmem(ep.host, p, n_terms);
  if (e.state_acc[p] == 1) return smem <= e.k_acc[p].smem ? &e.k_acc[p] : nullptr;
  if (e.state_acc[p] < 0 || !JitWorthIt(e.work_acc, rows, e.calls_acc)) return nullptr;
  std::string why;
  if (!JitAvailable(&why)) return nullptr;
  PrefetchExpJit(ep, true);
  e.state_acc[p] = -1;
  if (!ExpectPassIsJitable(ep.host, p)) return nullptr;
  const std::string& src = e.src_acc[p];
  std::string err;
  if (!JitCompile(src, "tfqb_jit_accum", false, JitAccumThreads(), smem, &e.k_acc[p], &err)) {
    if (getenv("TFQB_JIT_VERBOSE")) fprintf(stderr, "tfqb jit: %s\n", err.c_str());
    return nullptr;
  }
  ctx->prof.jit_kernels++;
  e.state_acc[p] = 1;
  return &e.k_acc[p];
}

// All terms of a group's PauliSums: tile passes + generic leftovers.
int RunExpectationTerms(tfqb_context* ctx, const CompiledExpPlan& ep,
                        const float2* psi, int rows, const DevTerm* d_terms,
                        int n_terms, int n_ops, double* per_term,
                        unsigned long long rank_base = 0, bool account = true) {
  NvtxRange nvtx("tfqb:expectation_passes");
  const ExpectationPlan& h = ep.host;
  const size_t row_stride = size_t(1) << h.n_alloc;
  const double ebytes = 8.0 * double(row_stride) * rows;
  for (size_t p = 0; p < h.passes.size(); ++p) {
    const PassRec& pr = h.passes[p];
    ExpectLaunch el;
    el.passes = ep.passes;
    el.rounds = ep.rounds;
    el.xops = ep.xops;
    el.zterms = ep.zterms;
    el.n_zterms = p == 0 ? int(h.zterms.size()) : 0;
    el.pass_index = int(p);
    el.tile_bits = pr.tile_bits;
    el.low_bits = pr.low_bits;
    el.n_alloc = h.n_alloc;
    el.n_rounds = pr.round_end - pr.round_begin;
    el.n_xops = el.n_rounds ? h.rounds[pr.round_end - 1].op_end -
                                  h.rounds[pr.round_begin].op_begin
                            : 0;
    el.n_terms = n_terms;
    el.rank_base = rank_base;
    const JitKernel* jk =
        ExpJitKernelFor(ctx, ep, int(p), double(row_stride) * rows, rows, account);
    const int hnd = BeginTimed(ctx, 2, ebytes);
    if (jk) {
      const unsigned long long n_tiles = 1ull << (h.n_alloc - pr.tile_bits);
      // one CTA keeps its terms in registers over many tiles: a few CTAs per
      // row are enough once there are rows to fill the machine
      unsigned long long ctas = (2368 + rows - 1) / rows;
      if (ctas * 1024 < n_tiles) ctas = (n_tiles + 1023) / 1024;
      if (ctas > n_tiles) ctas = n_tiles;
      if (Deterministic()) ctas = 1;   // per-term sums of a row in one CTA, one order
      std::string jerr;
      ctx->prof.jit_pass_launches++;
      if (!JitLaunchExpect(*jk, unsigned(ctas), unsigned(rows), psi, row_stride, n_tiles,
                           rank_base, per_term, n_terms, ctx->stream, &jerr))
        return Fail(TFQB_INTERNAL, jerr);
    } else {
      LaunchExpectPass(el, psi, row_stride, rows, per_term, ctx->stream);
    }
    EndTimed(ctx, hnd);
    ctx->prof.kernel_launches++;
    ctx->prof.expectation_launches++;
    ctx->prof.expectation_bytes += ebytes;
  }
  if (!h.generic_terms.empty()) {
    const double gbytes = ebytes * double(h.generic_terms.size());
    const int hnd = BeginTimed(ctx, 2, gbytes);
    LaunchExpectationTerms(psi, row_stride, h.n_alloc, d_terms, n_terms,
                           ep.generic, int(h.generic_terms.size()), rows,
                           per_term, ctx->stream);
    EndTimed(ctx, hnd);
    ctx->prof.kernel_launches++;
    ctx->prof.expectation_launches++;
    ctx->prof.expectation_bytes += gbytes;
  }
  (void)n_ops;
  return TFQB_OK;
}

int ExpLowBits() {
  static const int v = [] {
    const char* e = getenv("TFQB_EXP_LOW_BITS");
    const int r = e && *e ? atoi(e) : 3;
    return r < 1 ? 1 : (r > kLowBits ? kLowBits : r);
  }();
  return v;
}

int GateLowBits() {
  static const int v = [] {
    const char* e = getenv("TFQB_GATE_LOW_BITS");
    const int r = e && *e ? atoi(e) : kLowBits;
    return r < 1 ? 1 : (r > 7 ? 7 : r);
  }();
  return v;
}

int AdjRegBits() {
  static const int v = [] {
    const char* e = getenv("TFQB_ADJ_REGBITS");
    const int r = e && *e ? atoi(e) : kRegBitsAdj;
    return r == 4 ? 4 : 3;
  }();
  return v;
}

int CompilePlan(tfqb_context* ctx, DevicePlan&& hp,
                std::unique_ptr<CompiledPlan>* out) {
  auto cp = std::make_unique<CompiledPlan>();
  cp->host = std::move(hp);
  TFQB_RETURN_IF(UploadPlan(ctx, cp->host, &cp->dev));
  *out = std::move(cp);
  return TFQB_OK;
}

// ---- pass execution -------------------------------------------------------
int BeginTimed(tfqb_context* ctx, int kind, double bytes) {
  if (!ctx->prof_timing) return -1;
  std::pair<cudaEvent_t, cudaEvent_t> ev;
  if (!ctx->event_pool.empty()) {
    ev = ctx->event_pool.back();
    ctx->event_pool.pop_back();
  } else {
    cudaEventCreate(&ev.first);
    cudaEventCreate(&ev.second);
  }
  cudaEventRecord(ev.first, ctx->stream);
  ctx->timed.push_back(TimedLaunch{ev.first, ev.second, kind, bytes});
  return int(ctx->timed.size()) - 1;
}
void EndTimed(tfqb_context* ctx, int h) {
  if (h >= 0) cudaEventRecord(ctx->timed[h].b, ctx->stream);
}

// Generate the source of every pass of a plan variant and start compiling all
// of them at once on host threads; cp.jit_mu must be held.
static void PrefetchPlanJitLocked(const CompiledPlan& cp, bool adjoint, bool pf) {
  const size_t np = cp.host.passes.size();
  if (cp.jit_state.size() != 2 * np) {
    cp.jit_state.assign(2 * np, 0);
    cp.jit.assign(2 * np, JitKernel());
  }
  if (cp.jit_src.size() != 2 * np) cp.jit_src.assign(2 * np, std::string());
  for (size_t q = 0; q < np; ++q) {
    const size_t idx = q + (pf ? np : 0);
    if (cp.jit_state[idx] != 0 || !cp.jit_src[idx].empty()) continue;
    if (!PassIsJitable(cp.host, int(q), adjoint)) continue;
    cp.jit_src[idx] = GeneratePassSource(cp.host, int(q), adjoint, pf);
    JitPrefetch(cp.jit_src[idx]);
  }
}

static bool JitPhaseFreeEnabled() {
  static const bool v = [] {
    const char* e = getenv("TFQB_JIT_PHASE_FREE");
    return !(e && *e == '0');
  }();
  return v;
}

// Start compiling everything a job will need -- forward, expectation /
// accumulation and adjoint kernels -- before its first pass runs, so that the
// NVRTC work of all of them overlaps (cold start of a new circuit structure).
static void PrefetchJobJit(const CompiledPlan* fwd, const CompiledPlan* adj,
                           const CompiledExpPlan* ep, bool accum, double amps, int rows,
                           bool phase_free) {
  std::string why;
  if (!JitAvailable(&why)) return;
  if (fwd) {
    std::lock_guard<std::mutex> lock(fwd->jit_mu);
    if (JitWorthIt(fwd->jit_work + amps, rows, fwd->jit_calls + 1))
      PrefetchPlanJitLocked(*fwd, false, JitPhaseFreeEnabled() && phase_free);
  }
  if (adj) {
    std::lock_guard<std::mutex> lock(adj->jit_mu);
    if (JitWorthIt(adj->jit_work + amps, rows, adj->jit_calls + 1))
      PrefetchPlanJitLocked(*adj, true, JitPhaseFreeEnabled());
  }
  if (ep && ep->jit) {
    const ExpJitEntry& e = *ep->jit;
    if (accum ? JitWorthIt(e.work_acc + amps, rows, e.calls_acc + 1)
              : JitWorthIt(e.work + amps, rows, e.calls + 1))
      PrefetchExpJit(*ep, accum);
  }
}

// Specialised kernel of pass `p`, compiled on first use once the plan has seen
// TFQB_JIT_MIN_AMPS amplitudes (default 2^27); nullptr -> interpreted kernel.
static const JitKernel* JitKernelFor(tfqb_context* ctx, const CompiledPlan& cp,
                                     int pass, bool adjoint, int rows, bool phase_free,
                                     bool may_compile = true) {
  // jobs that cannot see a global phase (expectation, sampling, adjoint) get
  // their own variant of a forward pass: see jit.h GeneratePassSource
  // (the adjoint variant keeps psi and lambda in one frame and is always valid)
  const bool pf = JitPhaseFreeEnabled() && (phase_free || adjoint);
  std::lock_guard<std::mutex> lock(cp.jit_mu);
  const size_t np = cp.host.passes.size();
  if (cp.jit_state.size() != 2 * np) {
    cp.jit_state.assign(2 * np, 0);
    cp.jit.assign(2 * np, JitKernel());
  }
  const int p = pass + (pf ? int(np) : 0);
  if (cp.jit_state[p] == 1) return &cp.jit[p];
  if (cp.jit_state[p] < 0 || !may_compile || !JitWorthIt(cp.jit_work, rows, cp.jit_calls))
    return nullptr;
  std::string why;
  if (!JitAvailable(&why)) return nullptr;
  PrefetchPlanJitLocked(cp, adjoint, pf);
  cp.jit_state[p] = -1;
  if (!PassIsJitable(cp.host, pass, adjoint)) return nullptr;
  const std::string& src = cp.jit_src[p];
  if (src.empty()) return nullptr;
  std::string err;
  cp.jit[p].tiles = JitPassTiles(cp.host, adjoint, pass);
  if (!JitCompile(src, "tfqb_jit_pass", adjoint, JitPassThreads(cp.host, adjoint),
                  JitPassSmem(cp.host, pass, adjoint), &cp.jit[p], &err)) {
    if (getenv("TFQB_JIT_VERBOSE")) fprintf(stderr, "tfqb jit: %s\n", err.c_str());
    return nullptr;
  }
  ctx->prof.jit_kernels++;
  cp.jit_state[p] = 1;
  return &cp.jit[p];
}

// Evaluate the matrices of `cp` for `rows` rows (params: [rows, n_params]) and
// run every pass over psi (and lam for adjoint plans).
int RunPlan(tfqb_context* ctx, const CompiledPlan& cp, float2* psi, float2* lam,
            int rows, const float* d_params, int n_params, float* d_mats,
            bool init_zero, double* grad_out,
            unsigned long long rank_base = 0,
            bool phase_free = false, bool account = true,
            const float2* const* peer_tab = nullptr, int peer_shift = 0,
            unsigned long long peer_self = 0, int pass_select = 0) {
  // pass_select: 0 every pass; 1 only pass 0 (it builds the matrices);
  // -1 every pass but pass 0 (matrices are already built)
  NvtxRange nvtx(lam ? "tfqb:adjoint_passes" : "tfqb:gate_passes");
  // !account: the caller has done the use accounting and the compilation of
  // this plan already (sharded jobs: nothing may load a module behind a wait)
  const DevicePlan& hp = cp.host;
  const size_t row_stride = size_t(1) << hp.n_alloc;
  const bool adjoint = lam != nullptr;
  const int mat_rows = hp.row_dependent ? rows : 1;
  if (!hp.mats.empty() && pass_select >= 0) {
    LaunchBuildMatrices(cp.dev.mats, cp.dev.factors, int(hp.mats.size()), d_params, n_params,
                        mat_rows, d_mats, size_t(hp.mat_floats), ctx->stream);
    ctx->prof.kernel_launches++;
  }
  if (hp.passes.empty() && init_zero) {
    LaunchSetZeroState(psi, row_stride, rows, ctx->stream);
    ctx->prof.kernel_launches++;
  }
  for (size_t p = 0; p < hp.passes.size(); ++p) {
    if ((pass_select > 0 && p > 0) || (pass_select < 0 && p == 0)) continue;
    const PassRec& pr = hp.passes[p];
    PassLaunch pl;
    pl.passes = cp.dev.passes;
    pl.rounds = cp.dev.rounds;
    pl.ops = cp.dev.ops;
    pl.mats = d_mats;
    pl.mat_row_stride = hp.row_dependent ? size_t(hp.mat_floats) : 0;
    pl.pass_index = int(p);
    pl.tile_bits = pr.tile_bits;
    pl.low_bits = pr.low_bits;
    pl.n_alloc = hp.n_alloc;
    const bool has_rounds = pr.round_end > pr.round_begin;
    pl.first_op = has_rounds ? hp.rounds[pr.round_begin].op_begin : 0;
    pl.n_ops_in_pass =
        has_rounds ? hp.rounds[pr.round_end - 1].op_end - pl.first_op : 0;
    pl.mat_len = pr.mat_len;
    pl.n_rounds = pr.round_end - pr.round_begin;
    pl.reg_bits = hp.reg_bits;
    pl.rank_base = rank_base;
    // pass 0 of a segment that follows a qubit swap gathers from the peers
    const bool gather = peer_tab != nullptr && p == 0 && !adjoint;
    if (gather) {
      pl.peer_tab = peer_tab;
      pl.peer_shift = peer_shift;
      pl.peer_self = peer_self;
    }
    const double amps = double(row_stride) * rows;
    if (p == 0 && account) {
      std::lock_guard<std::mutex> lock(cp.jit_mu);
      cp.jit_work += amps;
      cp.jit_calls++;
    }
    const JitKernel* jk = JitKernelFor(ctx, cp, int(p), adjoint, rows, phase_free, account);
    std::string jerr;
    if (adjoint) {
      const int h = BeginTimed(ctx, 1, 32.0 * amps);
      if (jk) {
        ctx->prof.jit_pass_launches++;
        if (!JitLaunch(*jk, (1u << (hp.n_alloc - pr.tile_bits)) / unsigned(jk->tiles), unsigned(rows), psi, lam,
                       row_stride, d_mats, pl.mat_row_stride, grad_out,
                       int(hp.grad_slots.size()), 0, rank_base, ctx->stream, &jerr))
          return Fail(TFQB_INTERNAL, jerr);
      } else {
        LaunchAdjointPass(pl, psi, lam, row_stride, rows, grad_out,
                          int(hp.grad_slots.size()), ctx->stream);
      }
      EndTimed(ctx, h);
      ctx->prof.adjoint_pass_launches++;
      ctx->prof.adjoint_pass_bytes += 32.0 * amps;
    } else {
      const bool zero = init_zero && p == 0;
      // a pass that synthesises |0..0> only writes: 8 B/amplitude
      const double bytes = (zero ? 8.0 : 16.0) * amps;
      const int h = BeginTimed(ctx, 0, bytes);
      const int init_mode = gather ? 3 : zero ? (hp.product_init ? 2 : 1) : 0;
      if (jk) {
        ctx->prof.jit_pass_launches++;
        if (!JitLaunch(*jk, (1u << (hp.n_alloc - pr.tile_bits)) / unsigned(jk->tiles), unsigned(rows), psi,
                       nullptr, row_stride, d_mats, pl.mat_row_stride, nullptr, 0,
                       init_mode, rank_base, ctx->stream, &jerr, pl.peer_tab, pl.peer_shift,
                       pl.peer_self))
          return Fail(TFQB_INTERNAL, jerr);
      } else {
        LaunchForwardPass(pl, psi, row_stride, rows, init_mode, ctx->stream);
      }
      EndTimed(ctx, h);
      ctx->prof.gate_pass_launches++;
      ctx->prof.gate_pass_bytes += bytes;
    }
    ctx->prof.kernel_launches++;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return Fail(TFQB_INTERNAL, std::string("kernel launch failed: ") +
                                   cudaGetErrorString(e));
  return TFQB_OK;
}

// ---- job ------------------------------------------------------------------
enum JobKind { kJobExpectation, kJobAdjoint, kJobSamples, kJobState,
               kJobSampledExpectation, kJobSharded, kJobNoisy, kJobNoisySamples,
               kJobUnitary };

constexpr int kShardedMaxShots = 1 << 16;   // per tfqb_sharded_sample call

struct ShardedState {
  ShardedPlan plan;
  std::vector<std::shared_ptr<CompiledPlan>> gates;   // cached in the CompiledProgram
  std::vector<std::unique_ptr<CompiledExpPlan>> exps;
  float2* buf[2] = {nullptr, nullptr};
  int cur = 0;
  int rank = 0, world = 1;
  bool state_ready = false;   // a gate segment has initialised the shard
  double* d_per_term = nullptr;
  int n_terms = 0;
  // ---- peer-memory link (tfqb_sharded_export / _connect / _enqueue): every
  // rank has the shard buffers and the flag block of every other rank mapped
  // (CUDA IPC between processes, raw pointers inside one process)
  bool connected = false;
  unsigned char* flag_block = nullptr;   // [0] ready, [1] done (u32), +64: partials
  size_t flag_bytes = 0;
  std::vector<void*> ipc_opened;         // mappings to close with the job
  const float2** d_peer_buf[2] = {nullptr, nullptr};   // device tables [world]
  unsigned** d_peer_ready = nullptr;
  unsigned** d_peer_done = nullptr;
  double** d_peer_parts = nullptr;
  int* d_error = nullptr;
  double* d_total = nullptr;
  unsigned epoch = 0;                    // last epoch this rank signalled
  bool enqueued = false;
  std::vector<cudaEvent_t> events;       // 3 per exchange of the last run
  int exchanges_run = 0;
  int fused_run = 0;                     // exchanges fused into the next pass
  // sampling buffers (jobs prepared without PauliSums): allocated at prepare,
  // because a device allocation is an implicit synchronisation point and may
  // not happen while any rank of this process spins in a peer wait
  double* d_tree = nullptr;
  double* d_norms = nullptr;             // [world + 1]
  double* d_u = nullptr;                 // [kShardedMaxShots]
  uint64_t* d_idx = nullptr;
  int32_t* d_rowid = nullptr;
  ~ShardedState() {
    if (flag_block) cudaFree(flag_block);
    for (void* p : ipc_opened) cudaIpcCloseMemHandle(p);
    for (cudaEvent_t e : events) cudaEventDestroy(e);
  }
};

struct Group {
  std::shared_ptr<CompiledProgram> prog;
  std::vector<int> rows;       // global row indices, ascending
  int begin = 0;               // offset in group order
  int chunk = 1;               // rows per launch (memory budget)
  // PauliSums of the group (identical strings for every row of the group)
  std::vector<PauliSumT> sums;           // [n_ops]
  std::vector<DevTerm> terms;            // flattened, op-major
  DevTerm* d_terms = nullptr;
  std::unique_ptr<CompiledExpPlan> exp;  // tile-based expectation plan
};

}  // namespace

struct tfqb_job {
  tfqb_context* ctx = nullptr;
  JobKind kind = kJobExpectation;
  int batch = 0, n_symbols = 0, n_ops = 0, nmax = 0, num_samples = 0;
  std::vector<Group> groups;
  std::vector<int> perm;            // group order -> global row
  std::vector<void*> owned;         // device buffers returned on free
  float* d_params = nullptr;        // [batch, n_symbols] group order
  float* d_down = nullptr;          // [batch, n_ops]
  float* d_out = nullptr;           // [batch, out_cols] group order
  int out_cols = 0;
  float2* d_psi = nullptr;
  float2* d_lam = nullptr;
  float* d_mats = nullptr;
  double* d_scratch64 = nullptr;    // per-term partials / gradient slots
  size_t scratch64_count = 0;
  int chunk_cap = 0;                // rows per chunk (upper bound)
  bool ran = false;
  bool noisy = false;               // programs may hold noise channels
  std::unique_ptr<ShardedState> sharded;
  // job of a multi-device context: one sub-job per child, rows [lo, hi)
  struct Sub { tfqb_job* job; int lo, hi; };
  std::vector<Sub> sub;

  ~tfqb_job() {
    if (!ctx || !sub.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (void* p : owned) ctx->Release(p);
    for (auto& g : groups)
      if (g.d_terms) ctx->Release(g.d_terms);
  }
  template <typename T>
  int Own(size_t count, T** out) {
    TFQB_RETURN_IF(AllocT(ctx, count, out));
    owned.push_back(*out);
    return TFQB_OK;
  }
};

namespace {

int CheckContext(tfqb_context* ctx) {
  if (!ctx)
    return Fail(TFQB_UNAVAILABLE,
                "No CUDA context: the B200 backend has no CPU fallback.");
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess)
    return Fail(TFQB_UNAVAILABLE, std::string("cudaSetDevice failed: ") +
                                      cudaGetErrorString(e));
  return TFQB_OK;
}

// Common prologue of all five ops: parse / resolve / lower each DISTINCT
// program, group the rows.  `sum_key(row)` distinguishes rows whose PauliSums
// differ (empty when the op has no PauliSums).
int BuildGroups(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                const tfqb_strings* pauli_sums, int sum_rows, int n_ops,
                tfqb_job* job) {
  NvtxRange nvtx("tfqb:parse_resolve_lower");
  if (in->batch < 0 || in->n_symbols < 0)
    return Fail(TFQB_INVALID_ARGUMENT, "negative tensor dimension");
  if (pauli_sums && sum_rows != in->batch)
    return Fail(TFQB_INVALID_ARGUMENT,
                "Number of circuits and PauliSums do not match. Got " +
                    std::to_string(in->batch) + " circuits and " +
                    std::to_string(sum_rows) + " paulisums.");
  if (in->symbol_rows != in->batch)
    return Fail(TFQB_INVALID_ARGUMENT,
                "Number of circuits and symbol_values do not match. Got " +
                    std::to_string(in->batch) + " circuits and " +
                    std::to_string(in->symbol_rows) + " symbol values.");
  job->batch = in->batch;
  job->n_symbols = in->n_symbols;
  job->n_ops = n_ops;

  std::string names_key = job->noisy ? "N\x1d" : "";
  for (int j = 0; j < in->n_symbols; ++j) {
    names_key.append(in->symbol_names.data[j], in->symbol_names.size[j]);
    names_key.push_back('\x1f');
  }
  names_key.push_back('\x1e');
  SymbolTable symbols =
      MakeSymbolTable(in->symbol_names.data, in->symbol_names.size, in->n_symbols);

  // row -> compiled program (dedupe by bytes; consecutive equal rows are the
  // common case: PQC tiles one circuit over the batch, pqc.py:336-339)
  std::vector<std::shared_ptr<CompiledProgram>> row_prog(in->batch);
  std::unordered_map<std::string, std::shared_ptr<CompiledProgram>> local;
  for (int i = 0; i < in->batch; ++i) {
    const char* d = in->programs.data[i];
    const size_t n = in->programs.size[i];
    if (i > 0 && in->programs.size[i - 1] == n &&
        (in->programs.data[i - 1] == d ||
         memcmp(in->programs.data[i - 1], d, n) == 0)) {
      row_prog[i] = row_prog[i - 1];
      continue;
    }
    std::string key = names_key;
    key.append(d, n);
    auto it = ctx->cache.find(key);
    if (it != ctx->cache.end()) {
      row_prog[i] = it->second;
      continue;
    }
    ProgramPB pb;
    if (!ParseProgram(d, n, &pb))
      return Fail(TFQB_INVALID_ARGUMENT, "Unparseable proto: " + std::string(d, std::min<size_t>(n, 64)));
    auto cp = std::make_shared<CompiledProgram>();
    Status s = LowerProgram(pb, symbols, &cp->circuit, job->noisy);
    if (!s.ok) return Fail(TFQB_INVALID_ARGUMENT, s.msg);
    if (ctx->cache_bytes > (size_t(256) << 20)) {
      ctx->cache.clear();
      ctx->cache_bytes = 0;
    }
    ctx->cache_bytes += key.size();
    ctx->cache.emplace(std::move(key), cp);
    row_prog[i] = cp;
  }

  // group rows by (program, pauli-sum strings)
  std::map<std::pair<CompiledProgram*, std::string>, int> index;
  int prev_group = -1;
  for (int i = 0; i < in->batch; ++i) {
    // same program and the very same PauliSum buffers as the previous row
    // (a tiled batch): same group, no key to build
    if (i > 0 && prev_group >= 0 && row_prog[i] == row_prog[i - 1]) {
      bool same = true;
      if (pauli_sums)
        for (int j = 0; j < n_ops && same; ++j) {
          const size_t k = size_t(i) * n_ops + j;
          same = pauli_sums->data[k] == pauli_sums->data[k - n_ops] &&
                 pauli_sums->size[k] == pauli_sums->size[k - n_ops];
        }
      if (same) {
        job->groups[prev_group].rows.push_back(i);
        continue;
      }
    }
    std::string skey;
    if (pauli_sums) {
      for (int j = 0; j < n_ops; ++j) {
        const size_t k = size_t(i) * n_ops + j;
        const uint64_t len = pauli_sums->size[k];
        skey.append(reinterpret_cast<const char*>(&len), 8);
        skey.append(pauli_sums->data[k], len);
      }
    }
    auto key = std::make_pair(row_prog[i].get(), std::move(skey));
    auto it = index.find(key);
    if (it == index.end()) {
      Group g;
      g.prog = row_prog[i];
      if (pauli_sums) {
        g.sums.resize(n_ops);
        for (int j = 0; j < n_ops; ++j) {
          const size_t k = size_t(i) * n_ops + j;
          PauliSumPB pb;
          if (!ParsePauliSum(pauli_sums->data[k], pauli_sums->size[k], &pb))
            return Fail(TFQB_INVALID_ARGUMENT,
                        "Unparseable proto: " +
                            std::string(pauli_sums->data[k],
                                        std::min<size_t>(pauli_sums->size[k], 64)));
          Status s = LowerPauliSum(pb, g.prog->circuit, &g.sums[j]);
          if (!s.ok) return Fail(TFQB_INVALID_ARGUMENT, s.msg);
          for (const auto& t : g.sums[j].terms) {
            DevTerm dt{};
            dt.x = t.x;
            dt.z = t.z;
            dt.coeff = t.coeff;
            dt.phase = t.phase;
            dt.op = j;
            dt.identity = t.identity ? 1 : 0;
            g.terms.push_back(dt);
          }
        }
      }
      it = index.emplace(std::move(key), int(job->groups.size())).first;
      job->groups.push_back(std::move(g));
    }
    job->groups[it->second].rows.push_back(i);
    prev_group = it->second;
  }
  job->nmax = 0;
  int off = 0;
  job->perm.clear();
  for (auto& g : job->groups) {
    g.begin = off;
    off += int(g.rows.size());
    job->perm.insert(job->perm.end(), g.rows.begin(), g.rows.end());
    job->nmax = std::max(job->nmax, g.prog->circuit.n);
  }
  job->nmax = std::max(job->nmax, ctx->force_nmax);
  // shifts by n below would wrap; sharded jobs check their local size instead
  if (job->kind != kJobSharded && job->nmax > kMaxDeviceQubits)
    return Fail(TFQB_RESOURCE_EXHAUSTED,
                "A " + std::to_string(job->nmax) +
                    "-qubit state does not fit in the device memory budget.");
  return TFQB_OK;
}

int UploadTerms(tfqb_job* job) {
  tfqb_context* ctx = job->ctx;
  for (auto& g : job->groups) {
    if (g.terms.empty() || g.prog->circuit.n == 0) continue;
    TFQB_RETURN_IF(AllocT(ctx, g.terms.size(), &g.d_terms));
    TFQB_CUDA(cudaMemcpyAsync(g.d_terms, g.terms.data(),
                              g.terms.size() * sizeof(DevTerm),
                              cudaMemcpyHostToDevice, ctx->stream));
    ctx->prof.h2d_bytes += int64_t(g.terms.size() * sizeof(DevTerm));
  }
  TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
  return TFQB_OK;
}

// Gather a [batch, cols] host tensor into group order and upload it.
template <typename T>
int UploadPermuted(tfqb_job* job, const T* host, int cols, T** dev) {
  tfqb_context* ctx = job->ctx;
  const size_t count = size_t(job->batch) * cols;
  TFQB_RETURN_IF(job->Own(std::max<size_t>(count, 1), dev));
  if (count == 0) return TFQB_OK;
  std::vector<T> tmp(count);
  for (int r = 0; r < job->batch; ++r)
    memcpy(tmp.data() + size_t(r) * cols, host + size_t(job->perm[r]) * cols,
           sizeof(T) * cols);
  TFQB_CUDA(cudaMemcpyAsync(*dev, tmp.data(), count * sizeof(T),
                            cudaMemcpyHostToDevice, ctx->stream));
  TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->prof.h2d_bytes += int64_t(count * sizeof(T));
  return TFQB_OK;
}

size_t Budget(tfqb_context* ctx) {
  if (ctx->budget) return ctx->budget;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  size_t cached = 0;
  for (auto& b : ctx->free_blocks) cached += b.cap;
  return size_t(double(free_b + cached) * 0.8);
}

// Ensure forward (and adjoint) plans exist for every non-empty group and size
// the shared chunk buffers.  state_bufs = number of full state vectors per
// row; extra_row_bytes = other per-row device bytes.
int PlanAndSize(tfqb_job* job, bool need_adj, int state_bufs,
                size_t extra_row_bytes, size_t scratch64_per_row_fn(const Group&)) {
  NvtxRange nvtx("tfqb:plan_passes");
  tfqb_context* ctx = job->ctx;
  size_t max_state = 0, max_mats = 0, max_s64 = 0;
  const size_t budget = Budget(ctx);
  const int cap = 65535;
  int max_chunk = 1;
  for (auto& g : job->groups) {
    CompiledProgram& cp = *g.prog;
    if (cp.circuit.n == 0) continue;
    // 8 << n wraps for n >= 61, and no device holds 2^41 amplitudes
    if (cp.circuit.n > kMaxDeviceQubits)
      return Fail(TFQB_RESOURCE_EXHAUSTED,
                  "A " + std::to_string(cp.circuit.n) +
                      "-qubit state does not fit in the device memory budget.");
    if (!cp.fwd) TFQB_RETURN_IF(CompilePlan(ctx, PlanForward(cp.circuit, kTileMax, GateLowBits(), true), &cp.fwd));
    if (need_adj && !cp.adj)
      TFQB_RETURN_IF(CompilePlan(ctx, PlanAdjoint(cp.circuit, kTileMax, GateLowBits(), AdjRegBits()), &cp.adj));
    const size_t sb = size_t(8) << cp.fwd->host.n_alloc;
    size_t mat_f = size_t(cp.fwd->host.mat_floats);
    if (need_adj) mat_f = std::max(mat_f, size_t(cp.adj->host.mat_floats));
    const size_t s64 = scratch64_per_row_fn ? scratch64_per_row_fn(g) : 0;
    const size_t per_row = sb * state_bufs + mat_f * 4 + s64 * 8 + extra_row_bytes;
    size_t rows_fit = budget / per_row;
    if (rows_fit == 0)
      return Fail(TFQB_RESOURCE_EXHAUSTED,
                  "A " + std::to_string(cp.circuit.n) +
                      "-qubit state does not fit in the device memory budget.");
    const int rows = int(std::min<size_t>({rows_fit, g.rows.size(), size_t(cap)}));
    g.chunk = rows;
    max_chunk = std::max(max_chunk, rows);
    max_state = std::max(max_state, sb * rows);
    max_mats = std::max(max_mats, mat_f * rows);
    max_s64 = std::max(max_s64, s64 * rows);
  }
  job->chunk_cap = max_chunk;
  if (max_state) {
    TFQB_RETURN_IF(job->Own(max_state / sizeof(float2), &job->d_psi));
    if (state_bufs >= 2) TFQB_RETURN_IF(job->Own(max_state / sizeof(float2), &job->d_lam));
  }
  TFQB_RETURN_IF(job->Own(std::max<size_t>(max_mats, 64), &job->d_mats));
  TFQB_RETURN_IF(job->Own(std::max<size_t>(max_s64, 8), &job->d_scratch64));
  job->scratch64_count = std::max<size_t>(max_s64, 8);
  return TFQB_OK;
}

size_t ExpScratch(const Group& g) { return g.terms.size(); }
size_t AdjScratch(const Group& g) {
  return g.prog->adj ? g.prog->adj->host.grad_slots.size() : 0;
}

// lambda = sum_j g_j sum_t c_t P_t psi (K3) for one chunk of a group.
int RunAccumulate(tfqb_context* ctx, const Group& g, const float2* psi,
                  float2* lam, int rows, const float* d_down, int n_ops) {
  NvtxRange nvtx("tfqb:accumulate_operators");
  const int nt = int(g.terms.size());
  const CompiledExpPlan* ep = g.exp.get();
  const int n_alloc = g.prog->fwd->host.n_alloc;
  const size_t row_stride = size_t(1) << n_alloc;
  bool written = false;
  if (ep) {
    const ExpectationPlan& h = ep->host;
    for (size_t p = 0; p < h.passes.size(); ++p) {
      const PassRec& pr = h.passes[p];
      ExpectLaunch el;
      el.passes = ep->passes;
      el.rounds = ep->rounds;
      el.xops = ep->xops;
      el.zterms = ep->zterms;
      el.n_zterms = p == 0 ? int(h.zterms.size()) : 0;
      el.pass_index = int(p);
      el.tile_bits = pr.tile_bits;
    el.low_bits = pr.low_bits;
      el.n_alloc = h.n_alloc;
      el.n_rounds = pr.round_end - pr.round_begin;
      el.n_xops = el.n_rounds ? h.rounds[pr.round_end - 1].op_end -
                                    h.rounds[pr.round_begin].op_begin
                              : 0;
      el.n_terms = nt;
      const JitKernel* jk = AccumJitKernelFor(ctx, *ep, int(p), double(row_stride) * rows, nt, rows);
      if (jk) {
        const unsigned long long n_tiles = 1ull << (h.n_alloc - pr.tile_bits);
        std::string jerr;
        ctx->prof.jit_pass_launches++;
        if (!JitLaunchAccum(*jk, unsigned(n_tiles < 65535 ? n_tiles : 65535), unsigned(rows),
                            psi, lam, row_stride, g.d_terms, nt, d_down, n_ops,
                            written ? 1 : 0, n_tiles, ctx->stream, &jerr))
          return Fail(TFQB_INTERNAL, jerr);
      } else {
        LaunchAccumPass(el, psi, lam, row_stride, rows, g.d_terms, d_down, n_ops,
                        written, ctx->stream);
      }
      ctx->prof.kernel_launches++;
      written = true;
    }
    if (!h.generic_terms.empty()) {
      LaunchAccumulateOperators(psi, lam, row_stride, n_alloc, g.d_terms, nt,
                                ep->generic, int(h.generic_terms.size()), written,
                                d_down, n_ops, rows, ctx->stream);
      ctx->prof.kernel_launches++;
      written = true;
    }
  }
  if (!written) {
    TFQB_CUDA(cudaMemsetAsync(lam, 0, size_t(rows) * row_stride * sizeof(float2),
                              ctx->stream));
  }
  return TFQB_OK;
}

// ---- the device work of the expectation / adjoint jobs --------------------
int RunExpectationDevice(tfqb_job* job) {
  tfqb_context* ctx = job->ctx;
  const int M = job->n_ops, P = job->n_symbols;
  for (auto& g : job->groups) {
    if (g.prog->circuit.n == 0) continue;
    const CompiledPlan& fwd = *g.prog->fwd;
    const int nt = int(g.terms.size());
    const int per = g.chunk;
    {
      const int rows0 = std::min(per, int(g.rows.size()));
      PrefetchJobJit(&fwd, nullptr, g.exp.get(), false,
                     double(size_t(1) << fwd.host.n_alloc) * rows0, rows0, true);
    }
    for (int c0 = 0; c0 < int(g.rows.size()); c0 += per) {
      const int rows = std::min(per, int(g.rows.size()) - c0);
      const int r0 = g.begin + c0;
      TFQB_RETURN_IF(RunPlan(ctx, fwd, job->d_psi, nullptr, rows,
                             job->d_params + size_t(r0) * P, P, job->d_mats,
                             true, nullptr, 0, true));
      if (nt > 0) {
        TFQB_CUDA(cudaMemsetAsync(job->d_scratch64, 0,
                                  size_t(rows) * nt * sizeof(double), ctx->stream));
        TFQB_RETURN_IF(RunExpectationTerms(ctx, *g.exp, job->d_psi, rows,
                                           g.d_terms, nt, M, job->d_scratch64));
      }
      LaunchCombineTerms(job->d_scratch64, g.d_terms, nt, M, rows,
                         job->d_out + size_t(r0) * M, size_t(M), ctx->stream);
      ctx->prof.kernel_launches++;
    }
  }
  TFQB_CUDA(cudaGetLastError());
  return TFQB_OK;
}

int RunAdjointDevice(tfqb_job* job) {
  tfqb_context* ctx = job->ctx;
  const int M = job->n_ops, P = job->n_symbols;
  for (auto& g : job->groups) {
    if (g.prog->circuit.n == 0) continue;
    const CompiledPlan& fwd = *g.prog->fwd;
    const CompiledPlan& adj = *g.prog->adj;
    const size_t row_stride = size_t(1) << fwd.host.n_alloc;
    const int nt = int(g.terms.size());
    const int ns = int(adj.host.grad_slots.size());
    const int per = g.chunk;
    {
      const int rows0 = std::min(per, int(g.rows.size()));
      PrefetchJobJit(&fwd, &adj, g.exp.get(), true, double(row_stride) * rows0, rows0, true);
    }
    for (int c0 = 0; c0 < int(g.rows.size()); c0 += per) {
      const int rows = std::min(per, int(g.rows.size()) - c0);
      const int r0 = g.begin + c0;
      const float* params = job->d_params + size_t(r0) * P;
      TFQB_RETURN_IF(RunPlan(ctx, fwd, job->d_psi, nullptr, rows, params, P,
                             job->d_mats, true, nullptr, 0, true));
      TFQB_RETURN_IF(RunAccumulate(ctx, g, job->d_psi, job->d_lam, rows,
                                   job->d_down + size_t(r0) * M, M));
      TFQB_CUDA(cudaMemsetAsync(job->d_scratch64, 0,
                                std::max<size_t>(size_t(rows) * ns, 1) * sizeof(double),
                                ctx->stream));
      TFQB_RETURN_IF(RunPlan(ctx, adj, job->d_psi, job->d_lam, rows, params, P,
                             job->d_mats, false, job->d_scratch64));
      // d_terms block is followed by the slot->column table (see prepare)
      const int32_t* d_slot_col =
          reinterpret_cast<const int32_t*>(g.d_terms + std::max(nt, 1));
      LaunchReduceGradSlots(job->d_scratch64, d_slot_col, ns, rows,
                            job->d_out + size_t(r0) * P, P, ctx->stream);
      ctx->prof.kernel_launches++;
    }
  }
  TFQB_CUDA(cudaGetLastError());
  return TFQB_OK;
}

int FetchOut(tfqb_job* job, float* out, float empty_fill, bool fill_empty) {
  NvtxRange nvtx("tfqb:fetch_result");
  tfqb_context* ctx = job->ctx;
  const int cols = job->out_cols;
  const size_t count = size_t(job->batch) * cols;
  std::vector<float> tmp(count);
  if (count) {
    TFQB_CUDA(cudaMemcpyAsync(tmp.data(), job->d_out, count * sizeof(float),
                              cudaMemcpyDeviceToHost, ctx->stream));
  }
  TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->prof.d2h_bytes += int64_t(count * sizeof(float));
  for (auto& g : job->groups) {
    const bool empty = g.prog->circuit.n == 0;
    for (size_t k = 0; k < g.rows.size(); ++k) {
      float* dst = out + size_t(g.rows[k]) * cols;
      if (empty) {
        if (fill_empty)
          for (int j = 0; j < cols; ++j) dst[j] = empty_fill;
      } else {
        memcpy(dst, tmp.data() + size_t(g.begin + k) * cols, sizeof(float) * cols);
      }
    }
  }
  return TFQB_OK;
}

int PrepareExpectation(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                       tfqb_strings pauli_sums, int sum_rows, int n_ops,
                       tfqb_job* job) {
  job->ctx = ctx;
  job->kind = kJobExpectation;
  TFQB_RETURN_IF(BuildGroups(ctx, in, &pauli_sums, sum_rows, n_ops, job));
  job->out_cols = n_ops;
  TFQB_RETURN_IF(UploadTerms(job));
  for (auto& g : job->groups) {
    if (g.prog->circuit.n == 0 || g.terms.empty()) continue;
    std::vector<TermMask> tm(g.terms.size());
    for (size_t k = 0; k < g.terms.size(); ++k)
      tm[k] = TermMask{g.terms[k].x, g.terms[k].z, g.terms[k].phase,
                       g.terms[k].identity != 0};
    TFQB_RETURN_IF(CompileExpPlan(ctx, PlanExpectation(g.prog->circuit.n, tm, false, kTileMax, ExpLowBits()), &g.exp));
  }
  TFQB_RETURN_IF(UploadPermuted(job, in->symbol_values, in->n_symbols, &job->d_params));
  TFQB_RETURN_IF(job->Own(std::max<size_t>(size_t(job->batch) * n_ops, 1), &job->d_out));
  TFQB_RETURN_IF(PlanAndSize(job, false, 1, 0, ExpScratch));
  return TFQB_OK;
}

int PrepareAdjoint(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                   tfqb_strings pauli_sums, int sum_rows, int n_ops,
                   const float* downstream, int grad_rows, int grad_cols,
                   tfqb_job* job) {
  job->ctx = ctx;
  job->kind = kJobAdjoint;
  TFQB_RETURN_IF(BuildGroups(ctx, in, &pauli_sums, sum_rows, n_ops, job));
  if (grad_rows != in->batch)
    return Fail(TFQB_INVALID_ARGUMENT,
                "Number of gradients and circuits do not match. Got " +
                    std::to_string(grad_rows) + " gradients and " +
                    std::to_string(in->batch) + " circuits.");
  if (grad_cols != n_ops)
    return Fail(TFQB_INVALID_ARGUMENT,
                "Number of gradients and pauli sum dimension do not match. Got " +
                    std::to_string(grad_cols) + " gradient entries and " +
                    std::to_string(n_ops) + " paulis per circuit.");
  job->out_cols = in->n_symbols;
  // plans first: the slot->column table is uploaded next to the terms
  for (auto& g : job->groups) {
    CompiledProgram& cp = *g.prog;
    if (cp.circuit.n == 0) continue;
    if (!cp.fwd) TFQB_RETURN_IF(CompilePlan(ctx, PlanForward(cp.circuit, kTileMax, GateLowBits(), true), &cp.fwd));
    if (!cp.adj) TFQB_RETURN_IF(CompilePlan(ctx, PlanAdjoint(cp.circuit, kTileMax, GateLowBits(), AdjRegBits()), &cp.adj));
    const int nt = int(g.terms.size());
    const auto& slots = cp.adj->host.grad_slots;
    const size_t bytes = size_t(std::max(nt, 1)) * sizeof(DevTerm) +
                         std::max<size_t>(slots.size(), 1) * sizeof(int32_t);
    void* p = nullptr;
    TFQB_RETURN_IF(ctx->Alloc(bytes, &p));
    g.d_terms = static_cast<DevTerm*>(p);
    std::vector<char> host(bytes, 0);
    if (nt) memcpy(host.data(), g.terms.data(), size_t(nt) * sizeof(DevTerm));
    int32_t* sc = reinterpret_cast<int32_t*>(host.data() + size_t(std::max(nt, 1)) * sizeof(DevTerm));
    for (size_t s = 0; s < slots.size(); ++s) sc[s] = slots[s].symbol_col;
    TFQB_CUDA(cudaMemcpyAsync(p, host.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->prof.h2d_bytes += int64_t(bytes);
  }
  for (auto& g : job->groups) {
    if (g.prog->circuit.n == 0 || g.terms.empty()) continue;
    std::vector<TermMask> tm(g.terms.size());
    for (size_t k = 0; k < g.terms.size(); ++k)
      tm[k] = TermMask{g.terms[k].x, g.terms[k].z, g.terms[k].phase,
                       g.terms[k].identity != 0};
    TFQB_RETURN_IF(CompileExpPlan(ctx, PlanExpectation(g.prog->circuit.n, tm, true), &g.exp));
  }
  TFQB_RETURN_IF(UploadPermuted(job, in->symbol_values, in->n_symbols, &job->d_params));
  TFQB_RETURN_IF(UploadPermuted(job, downstream, n_ops, &job->d_down));
  TFQB_RETURN_IF(job->Own(std::max<size_t>(size_t(job->batch) * in->n_symbols, 1), &job->d_out));
  TFQB_RETURN_IF(PlanAndSize(job, true, 2, 0, AdjScratch));
  return TFQB_OK;
}

uint32_t NextPow2(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================

namespace {

// Every int-returning entry point runs inside this barrier: std::bad_alloc and
// any other exception becomes a status code instead of unwinding through C.
template <typename F>
int GuardAbi(F&& f) {
  try {
    return f();
  } catch (const std::bad_alloc&) {
    return Fail(TFQB_RESOURCE_EXHAUSTED, "Out of host memory.");
  } catch (const std::exception& e) {
    return Fail(TFQB_INTERNAL, std::string("internal error: ") + e.what());
  } catch (...) {
    return Fail(TFQB_INTERNAL, "internal error: unknown exception");
  }
}
}  // namespace

static bool IsMultiCtx(const tfqb_context* ctx) { return ctx && !ctx->children.empty(); }

extern "C" {

int tfqb_abi_version(void) { return 4; }   // 3: tfqb_create_multi, tfqb_set_row_offset, in-library sharded exchange; 4: tfqb_sharded_sample, tfqb_ps_*, tfqb_jit_pending (additions only)

const char* tfqb_last_error(void) { return g_last_error.c_str(); }

static int impl_tfqb_create(int device, tfqb_context** out) {
  if (!out) return Fail(TFQB_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return Fail(TFQB_UNAVAILABLE,
                "No CUDA device available: the B200 backend has no CPU fallback.");
  }
  if (device < 0 || device >= count)
    return Fail(TFQB_INVALID_ARGUMENT, "CUDA device ordinal out of range.");
  TFQB_CUDA(cudaSetDevice(device));
  auto ctx = std::make_unique<tfqb_context>();
  ctx->device = device;
  TFQB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  *out = ctx.release();
  return TFQB_OK;
}

void tfqb_destroy(tfqb_context* ctx) {
  if (!ctx) return;
  if (!ctx->children.empty()) {
    for (tfqb_context* c : ctx->children) tfqb_destroy(c);
    delete ctx;
    return;
  }
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->cache.clear();
  ctx->Trim();
  for (auto& kv : ctx->live) cudaFree(kv.first);
  for (auto& t : ctx->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  for (auto& ev : ctx->event_pool) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

static int impl_tfqb_set_memory_budget(tfqb_context* ctx, size_t bytes) {
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->budget = bytes;
  return TFQB_OK;
}

static int impl_tfqb_expectation_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                             tfqb_strings pauli_sums, int sum_rows, int n_ops,
                             tfqb_job** job) {
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  auto j = std::make_unique<tfqb_job>();
  TFQB_RETURN_IF(PrepareExpectation(ctx, in, pauli_sums, sum_rows, n_ops, j.get()));
  *job = j.release();
  return TFQB_OK;
}

static int impl_tfqb_adjoint_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                         tfqb_strings pauli_sums, int sum_rows, int n_ops,
                         const float* downstream_grads, int grad_rows,
                         int grad_cols, tfqb_job** job) {
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  auto j = std::make_unique<tfqb_job>();
  TFQB_RETURN_IF(PrepareAdjoint(ctx, in, pauli_sums, sum_rows, n_ops,
                                downstream_grads, grad_rows, grad_cols, j.get()));
  *job = j.release();
  return TFQB_OK;
}

static int impl_tfqb_job_run_device(tfqb_job* job) {
  if (!job) return Fail(TFQB_INVALID_ARGUMENT, "job is null");
  TFQB_RETURN_IF(CheckContext(job->ctx));
  std::lock_guard<std::mutex> lock(job->ctx->mu);
  job->ran = true;
  if (job->kind == kJobExpectation) return RunExpectationDevice(job);
  if (job->kind == kJobAdjoint) return RunAdjointDevice(job);
  return Fail(TFQB_INVALID_ARGUMENT, "job kind has no device-resident run");
}

static int impl_tfqb_job_fetch(tfqb_job* job, float* out) {
  if (!job) return Fail(TFQB_INVALID_ARGUMENT, "job is null");
  TFQB_RETURN_IF(CheckContext(job->ctx));
  std::lock_guard<std::mutex> lock(job->ctx->mu);
  if (job->kind == kJobExpectation)
    return FetchOut(job, out, -2.0f, true);   // tfq_simulate_expectation_op.cc:213-216
  if (job->kind == kJobAdjoint)
    return FetchOut(job, out, 0.0f, true);    // tfq_adj_grad_op.cc:152,209-212
  return Fail(TFQB_INVALID_ARGUMENT, "job kind has no float result");
}

void tfqb_job_free(tfqb_job* job) {
  if (!job) return;
  tfqb_context* ctx = job->ctx;
  if (!job->sub.empty()) {
    for (auto& sj : job->sub) tfqb_job_free(sj.job);
    delete job;
    return;
  }
  if (ctx) {
    cudaSetDevice(ctx->device);
    std::lock_guard<std::mutex> lock(ctx->mu);
    delete job;
  } else {
    delete job;
  }
}

static int impl_tfqb_simulate_expectation(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                              tfqb_strings pauli_sums, int sum_rows, int n_ops,
                              float* expectations) {
  tfqb_job* job = nullptr;
  TFQB_RETURN_IF(tfqb_expectation_prepare(ctx, in, pauli_sums, sum_rows, n_ops, &job));
  int rc = tfqb_job_run_device(job);
  if (rc == TFQB_OK) rc = tfqb_job_fetch(job, expectations);
  tfqb_job_free(job);
  return rc;
}

static int impl_tfqb_adjoint_gradient(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                          tfqb_strings pauli_sums, int sum_rows, int n_ops,
                          const float* downstream_grads, int grad_rows,
                          int grad_cols, float* grads) {
  tfqb_job* job = nullptr;
  TFQB_RETURN_IF(tfqb_adjoint_prepare(ctx, in, pauli_sums, sum_rows, n_ops,
                                      downstream_grads, grad_rows, grad_cols, &job));
  int rc = tfqb_job_run_device(job);
  if (rc == TFQB_OK) rc = tfqb_job_fetch(job, grads);
  tfqb_job_free(job);
  return rc;
}

// ---- state ----------------------------------------------------------------
static int impl_tfqb_simulate_state_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                                tfqb_job** job, int* max_qubits) {
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  auto j = std::make_unique<tfqb_job>();
  j->ctx = ctx;
  j->kind = kJobState;
  TFQB_RETURN_IF(BuildGroups(ctx, in, nullptr, 0, 0, j.get()));
  TFQB_RETURN_IF(UploadPermuted(j.get(), in->symbol_values, in->n_symbols, &j->d_params));
  // second "state buffer" is the padded export tile [rows, 2^nmax]
  const size_t out_row = size_t(8) << j->nmax;
  TFQB_RETURN_IF(PlanAndSize(j.get(), false, 1, out_row, nullptr));
  if (max_qubits) *max_qubits = j->nmax;
  *job = j.release();
  return TFQB_OK;
}

static int impl_tfqb_simulate_state_run(tfqb_job* job, float* state_vector) {
  if (!job || job->kind != kJobState) return Fail(TFQB_INVALID_ARGUMENT, "not a state job");
  tfqb_context* ctx = job->ctx;
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  const int P = job->n_symbols;
  const size_t out_cols = size_t(1) << job->nmax;
  float2* out = reinterpret_cast<float2*>(state_vector);
  float2* d_export = nullptr;
  for (auto& g : job->groups) {
    if (g.prog->circuit.n == 0) {
      // tfq_simulate_state_op.cc:154-167: a 1-amplitude |0> then (-2, 0)
      for (int r : g.rows) {
        float2* dst = out + size_t(r) * out_cols;
        dst[0] = make_float2(1.f, 0.f);
        for (size_t k = 1; k < out_cols; ++k) dst[k] = make_float2(-2.f, 0.f);
      }
      continue;
    }
    const CompiledPlan& fwd = *g.prog->fwd;
    const size_t row_stride = size_t(1) << fwd.host.n_alloc;
    const int per = g.chunk;
    if (!d_export)
      TFQB_RETURN_IF(job->Own(size_t(job->chunk_cap) * out_cols, &d_export));
    for (int c0 = 0; c0 < int(g.rows.size()); c0 += per) {
      const int rows = std::min(per, int(g.rows.size()) - c0);
      const int r0 = g.begin + c0;
      TFQB_RETURN_IF(RunPlan(ctx, fwd, job->d_psi, nullptr, rows,
                             job->d_params + size_t(r0) * P, P, job->d_mats,
                             true, nullptr, 0));
      LaunchExportState(job->d_psi, row_stride, g.prog->circuit.n, d_export,
                        out_cols, rows, ctx->stream);
      ctx->prof.kernel_launches++;
      for (int k = 0; k < rows; ++k) {
        TFQB_CUDA(cudaMemcpyAsync(out + size_t(g.rows[c0 + k]) * out_cols,
                                  d_export + size_t(k) * out_cols,
                                  out_cols * sizeof(float2),
                                  cudaMemcpyDeviceToHost, ctx->stream));
      }
      TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
      ctx->prof.d2h_bytes += int64_t(size_t(rows) * out_cols * sizeof(float2));
    }
  }
  return TFQB_OK;
}

// ---- samples --------------------------------------------------------------
static int impl_tfqb_simulate_samples_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                                  int num_samples, tfqb_job** job, int* max_qubits) {
  TFQB_RETURN_IF(CheckContext(ctx));
  if (num_samples < 0) return Fail(TFQB_INVALID_ARGUMENT, "num_samples must be >= 0");
  std::lock_guard<std::mutex> lock(ctx->mu);
  auto j = std::make_unique<tfqb_job>();
  j->ctx = ctx;
  j->kind = kJobSamples;
  j->num_samples = num_samples;
  TFQB_RETURN_IF(BuildGroups(ctx, in, nullptr, 0, 0, j.get()));
  TFQB_RETURN_IF(UploadPermuted(j.get(), in->symbol_values, in->n_symbols, &j->d_params));
  const size_t padded = NextPow2(std::max(num_samples, 1));
  // per row: uniforms (8B * padded), indices (8B * S), int8 out (S * nmax), tree
  size_t extra = padded * 8 + size_t(num_samples) * 8 + size_t(num_samples) * std::max(j->nmax, 1);
  extra += TreeDoublesPerRow(std::max(j->nmax, kMinStateBits)) * 8;
  TFQB_RETURN_IF(PlanAndSize(j.get(), false, 1, extra, nullptr));
  if (max_qubits) *max_qubits = j->nmax;
  *job = j.release();
  return TFQB_OK;
}

static int impl_tfqb_simulate_samples_run(tfqb_job* job, uint64_t seed, const double* uniforms,
                              int8_t* samples) {
  if (!job || job->kind != kJobSamples) return Fail(TFQB_INVALID_ARGUMENT, "not a samples job");
  tfqb_context* ctx = job->ctx;
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  const int S = job->num_samples, P = job->n_symbols, nmax = job->nmax;
  if (S == 0 || job->batch == 0) return TFQB_OK;   // tfq_simulate_samples_op.cc:113-115
  const size_t padded = NextPow2(S);
  const int cap = job->chunk_cap;
  double* d_u = nullptr;
  uint64_t* d_idx = nullptr;
  int8_t* d_out8 = nullptr;
  double* d_tree = nullptr;
  int32_t* d_rowids = nullptr;
  TFQB_RETURN_IF(job->Own(size_t(cap) * padded, &d_u));
  TFQB_RETURN_IF(job->Own(size_t(cap) * S, &d_idx));
  TFQB_RETURN_IF(job->Own(size_t(cap) * S * std::max(nmax, 1), &d_out8));
  TFQB_RETURN_IF(job->Own(size_t(cap) * TreeDoublesPerRow(std::max(nmax, kMinStateBits)), &d_tree));
  TFQB_RETURN_IF(job->Own(std::max(job->batch, 1), &d_rowids));
  std::vector<int32_t> row_ids(job->perm.begin(), job->perm.end());
  for (int32_t& r : row_ids) r += int32_t(ctx->row_offset);
  TFQB_CUDA(cudaMemcpyAsync(d_rowids, row_ids.data(), sizeof(int32_t) * job->batch,
                            cudaMemcpyHostToDevice, ctx->stream));
  TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<double> hu;
  std::vector<int8_t> hout;
  for (auto& g : job->groups) {
    if (g.prog->circuit.n == 0) {
      for (int r : g.rows) memset(samples + size_t(r) * S * nmax, 0xFE, size_t(S) * nmax);  // -2
      continue;
    }
    const CompiledPlan& fwd = *g.prog->fwd;
    const int na = fwd.host.n_alloc;
    const size_t row_stride = size_t(1) << na;
    const int per = g.chunk;
    for (int c0 = 0; c0 < int(g.rows.size()); c0 += per) {
      const int rows = std::min(per, int(g.rows.size()) - c0);
      const int r0 = g.begin + c0;
      TFQB_RETURN_IF(RunPlan(ctx, fwd, job->d_psi, nullptr, rows,
                             job->d_params + size_t(r0) * P, P, job->d_mats,
                             true, nullptr, 0, true));
      LaunchBuildTree(job->d_psi, row_stride, na, d_tree, rows, ctx->stream);
      if (uniforms) {
        hu.assign(size_t(rows) * padded, 2.0);
        for (int k = 0; k < rows; ++k)
          memcpy(hu.data() + size_t(k) * padded,
                 uniforms + size_t(g.rows[c0 + k]) * S, sizeof(double) * S);
        TFQB_CUDA(cudaMemcpyAsync(d_u, hu.data(), hu.size() * sizeof(double),
                                  cudaMemcpyHostToDevice, ctx->stream));
        TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->prof.h2d_bytes += int64_t(hu.size() * sizeof(double));
      } else {
        LaunchFillUniforms(d_u, padded, seed, d_rowids + r0, 0, 0, S, rows, ctx->stream);
      }
      LaunchSortRows(d_u, padded, rows, ctx->stream);
      LaunchSample(job->d_psi, row_stride, na, d_tree, d_u, padded, nullptr, S,
                   rows, d_idx, size_t(S), ctx->stream);
      LaunchUnpackSamples(d_idx, size_t(S), g.prog->circuit.n, nmax, S, rows,
                          d_out8, ctx->stream);
      ctx->prof.kernel_launches += 5;
      const size_t row_bytes = size_t(S) * nmax;
      hout.resize(size_t(rows) * row_bytes);
      if (row_bytes) {
        TFQB_CUDA(cudaMemcpyAsync(hout.data(), d_out8, hout.size(),
                                  cudaMemcpyDeviceToHost, ctx->stream));
      }
      TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
      ctx->prof.d2h_bytes += int64_t(hout.size());
      for (int k = 0; k < rows; ++k)
        memcpy(samples + size_t(g.rows[c0 + k]) * row_bytes,
               hout.data() + size_t(k) * row_bytes, row_bytes);
    }
  }
  TFQB_CUDA(cudaGetLastError());
  return TFQB_OK;
}

// ---- sampled expectation ----------------------------------------------------
static int impl_tfqb_simulate_sampled_expectation(
    tfqb_context* ctx, const tfqb_circuit_inputs* in, tfqb_strings pauli_sums,
    int sum_rows, int n_ops, const int32_t* num_samples, int ns_rows,
    int ns_cols, uint64_t seed, const double* uniforms, int uniform_terms,
    int uniform_shots, float* expectations) {
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  auto jp = std::make_unique<tfqb_job>();
  tfqb_job* job = jp.get();
  job->ctx = ctx;
  job->kind = kJobSampledExpectation;
  TFQB_RETURN_IF(BuildGroups(ctx, in, &pauli_sums, sum_rows, n_ops, job));
  if (ns_rows != sum_rows)
    return Fail(TFQB_INVALID_ARGUMENT,
                "Dimension 0 of num_samples and pauli_sums do not match.Got " +
                    std::to_string(ns_rows) + " lists of sample sizes and " +
                    std::to_string(sum_rows) + " lists of pauli sums.");
  if (ns_cols != n_ops)
    return Fail(TFQB_INVALID_ARGUMENT,
                "Dimension 1 of num_samples and pauli_sums do not match.Got " +
                    std::to_string(ns_cols) + " lists of sample sizes and " +
                    std::to_string(n_ops) + " lists of pauli sums.");
  const int B = job->batch, M = n_ops, P = job->n_symbols;
  int max_shots = 0;
  for (size_t k = 0; k < size_t(B) * M; ++k) {
    if (num_samples[k] < 1)
      return Fail(TFQB_INVALID_ARGUMENT,
                  "Each element of num_samples must be greater than 0.");
    max_shots = std::max(max_shots, num_samples[k]);
  }
  if (uniforms && uniform_shots < max_shots)
    return Fail(TFQB_INVALID_ARGUMENT, "uniforms tensor holds too few shots");
  job->out_cols = M;
  TFQB_RETURN_IF(UploadPermuted(job, in->symbol_values, P, &job->d_params));
  TFQB_RETURN_IF(job->Own(std::max<size_t>(size_t(B) * M, 1), &job->d_out));
  TFQB_CUDA(cudaMemsetAsync(job->d_out, 0, std::max<size_t>(size_t(B) * M, 1) * sizeof(float), ctx->stream));
  // num_samples transposed to [M][B] in group order
  int32_t* d_ns = nullptr;
  {
    std::vector<int32_t> t(std::max<size_t>(size_t(B) * M, 1));
    for (int r = 0; r < B; ++r)
      for (int j = 0; j < M; ++j)
        t[size_t(j) * B + r] = num_samples[size_t(job->perm[r]) * M + j];
    TFQB_RETURN_IF(job->Own(t.size(), &d_ns));
    TFQB_CUDA(cudaMemcpyAsync(d_ns, t.data(), t.size() * sizeof(int32_t),
                              cudaMemcpyHostToDevice, ctx->stream));
    TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  int32_t* d_rowids = nullptr;
  TFQB_RETURN_IF(job->Own(std::max(B, 1), &d_rowids));
  {
    std::vector<int32_t> row_ids(job->perm.begin(), job->perm.end());
    for (int32_t& r : row_ids) r += int32_t(ctx->row_offset);
    TFQB_CUDA(cudaMemcpyAsync(d_rowids, row_ids.data(), sizeof(int32_t) * B,
                              cudaMemcpyHostToDevice, ctx->stream));
    TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  size_t extra = size_t(max_shots) * 16 +
                 TreeDoublesPerRow(std::max(job->nmax, kMinStateBits)) * 8;
  TFQB_RETURN_IF(PlanAndSize(job, false, 2, extra, nullptr));
  const int cap = job->chunk_cap;
  double* d_u = nullptr;
  uint64_t* d_idx = nullptr;
  double* d_tree = nullptr;
  float* d_rot_mats = nullptr;
  TFQB_RETURN_IF(job->Own(size_t(cap) * std::max(max_shots, 1), &d_u));
  TFQB_RETURN_IF(job->Own(size_t(cap) * std::max(max_shots, 1), &d_idx));
  TFQB_RETURN_IF(job->Own(size_t(cap) * TreeDoublesPerRow(std::max(job->nmax, kMinStateBits)), &d_tree));
  TFQB_RETURN_IF(job->Own(64 * 8 + 64, &d_rot_mats));
  std::vector<double> hu;
  for (auto& g : job->groups) {
    if (g.prog->circuit.n == 0) continue;
    const CompiledPlan& fwd = *g.prog->fwd;
    const int n = g.prog->circuit.n;
    const int na = fwd.host.n_alloc;
    const size_t row_stride = size_t(1) << na;
    const int per = g.chunk;
    // Z-basis rotation plans, one per non-identity term with X/Y factors
    std::vector<std::vector<std::unique_ptr<CompiledPlan>>> rot(M);
    for (int j = 0; j < M; ++j) {
      rot[j].resize(g.sums[j].terms.size());
      for (size_t t = 0; t < g.sums[j].terms.size(); ++t) {
        const PauliTermT& term = g.sums[j].terms[t];
        if (term.identity || term.rot.empty()) continue;
        TFQB_RETURN_IF(CompilePlan(ctx, PlanRotations(n, term.rot), &rot[j][t]));
      }
    }
    for (int c0 = 0; c0 < int(g.rows.size()); c0 += per) {
      const int rows = std::min(per, int(g.rows.size()) - c0);
      const int r0 = g.begin + c0;
      TFQB_RETURN_IF(RunPlan(ctx, fwd, job->d_psi, nullptr, rows,
                             job->d_params + size_t(r0) * P, P, job->d_mats,
                             true, nullptr, 0, true));
      bool psi_tree_valid = false;   // d_tree holds the tree of d_psi
      for (int j = 0; j < M; ++j) {
        float* acc = job->d_out + size_t(r0) * M + j;
        const int32_t* shots_row = d_ns + size_t(j) * B + r0;
        int chunk_shots = 0;
        for (int k = 0; k < rows; ++k)
          chunk_shots = std::max(chunk_shots, num_samples[size_t(g.rows[c0 + k]) * M + j]);
        for (size_t t = 0; t < g.sums[j].terms.size(); ++t) {
          const PauliTermT& term = g.sums[j].terms[t];
          if (term.identity) {   // util_qsim.h:213-217
            LaunchAddConstant(term.coeff, rows, acc, size_t(M), ctx->stream);
            ctx->prof.kernel_launches++;
            continue;
          }
          const float2* src = job->d_psi;
          // Z-type terms sample the circuit's own state: its probability tree
          // is built once per chunk, not once per term
          const bool reuse_tree = !rot[j][t] && psi_tree_valid;
          if (rot[j][t]) {
            TFQB_CUDA(cudaMemcpyAsync(job->d_lam, job->d_psi,
                                      size_t(rows) * row_stride * sizeof(float2),
                                      cudaMemcpyDeviceToDevice, ctx->stream));
            TFQB_RETURN_IF(RunPlan(ctx, *rot[j][t], job->d_lam, nullptr, rows,
                                   nullptr, 0, d_rot_mats, false, nullptr));
            src = job->d_lam;
          }
          if (!reuse_tree) LaunchBuildTree(src, row_stride, na, d_tree, rows, ctx->stream);
          psi_tree_valid = !rot[j][t];
          if (uniforms) {
            hu.assign(size_t(rows) * chunk_shots, 0.0);
            for (int k = 0; k < rows; ++k) {
              if (int(t) >= uniform_terms)
                return Fail(TFQB_INVALID_ARGUMENT, "uniforms tensor holds too few terms");
              const size_t off =
                  ((size_t(g.rows[c0 + k]) * M + j) * uniform_terms + t) * uniform_shots;
              memcpy(hu.data() + size_t(k) * chunk_shots, uniforms + off,
                     sizeof(double) * chunk_shots);
            }
            TFQB_CUDA(cudaMemcpyAsync(d_u, hu.data(), hu.size() * sizeof(double),
                                      cudaMemcpyHostToDevice, ctx->stream));
            TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
          } else {
            LaunchFillUniforms(d_u, size_t(chunk_shots), seed, d_rowids + r0,
                               uint32_t(j), uint32_t(t), chunk_shots, rows, ctx->stream);
          }
          LaunchSample(src, row_stride, na, d_tree, d_u, size_t(chunk_shots),
                       shots_row, chunk_shots, rows, d_idx, size_t(chunk_shots),
                       ctx->stream);
          LaunchParityExpectation(d_idx, size_t(chunk_shots), term.parity_mask,
                                  term.coeff, shots_row, chunk_shots, rows, acc,
                                  size_t(M), ctx->stream);
          ctx->prof.kernel_launches += 4;
        }
      }
      // rotation plans are destroyed after the group: wait for the chunk
      TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
  }
  TFQB_CUDA(cudaGetLastError());
  return FetchOut(job, expectations, -2.0f, true);
}



// ---- noisy trajectory ops (next-row N2) ---------------------------------------
// TfqNoisyExpectation / TfqNoisySampledExpectation / TfqNoisySamples
// (core/ops/noise/*.cc).  The reference runs one qsim trajectory at a time per
// host thread; here the trajectories ARE the batch: every (circuit,
// trajectory) pair is a row of the same pass kernels, the Kraus operator a row
// draws at a channel is just that row's 2x2 matrix (gates.cuh channel_matrix,
// evaluated on the device from the row's uniform), and only the non-unitary
// channels cut the circuit, because their draw needs a population of the
// state as it is at that point (one read sweep each).
}  // extern "C"

namespace {

constexpr uint32_t kSampleStream = 0x73616D70u;    // "samp", oracle SAMPLE_STREAM

int EnsureNoisyPlan(tfqb_context* ctx, CompiledProgram& cp) {
  if (cp.noisy) return TFQB_OK;
  auto np = std::make_unique<NoisyPlan>();
  const CircuitT& c = cp.circuit;
  std::vector<CircuitT> subs(1);
  auto fresh = [&]() {
    CircuitT x;
    x.n = c.n;
    x.n_symbols = c.n_symbols;
    x.n_channels = c.n_channels;
    x.n_nonunitary = c.n_nonunitary;
    return x;
  };
  subs[0] = fresh();
  np->before.push_back(NoisyPlan::Measure{});
  for (const GateT& g : c.gates) {
    if (g.needs_population()) {
      subs.push_back(fresh());
      NoisyPlan::Measure m;
      m.bit = g.bit[0];
      m.col = g.aux_sym;
      np->before.push_back(m);
    }
    subs.back().gates.push_back(g);
  }
  for (size_t k = 0; k < subs.size(); ++k) {
    std::unique_ptr<CompiledPlan> plan;
    TFQB_RETURN_IF(CompilePlan(
        ctx, PlanForward(subs[k], kTileMax, GateLowBits(), true, k == 0), &plan));
    np->mat_floats = std::max(np->mat_floats, size_t(plan->host.mat_floats));
    np->segs.push_back(std::move(plan));
  }
  cp.noisy = std::move(np);
  return TFQB_OK;
}

// One (circuit, trajectory) row of a group.
struct TrajRow { int32_t group_pos, circuit, traj; };

struct NoisyBuffers {
  float* d_params = nullptr;      // [chunk, cols]
  int32_t* d_sym_row = nullptr;   // [chunk] row of the job's symbol values
  int32_t* d_circuit = nullptr;   // [chunk] global circuit index (Philox)
  int32_t* d_traj = nullptr;      // [chunk]
  long long* d_given_off = nullptr;
  double* d_pop = nullptr;        // [chunk, 2]
  float* d_given = nullptr;       // caller's uniforms (whole tensor)
  int chunk = 0;
};

// Simulate rows [r0, r0 + rows) of `list`: psi holds the final states.
int RunTrajectories(tfqb_job* job, const Group& g, const std::vector<TrajRow>& list, int r0,
                    int rows, NoisyBuffers& nb, uint64_t seed, int uniform_traj,
                    int uniform_chan) {
  tfqb_context* ctx = job->ctx;
  const CircuitT& c = g.prog->circuit;
  const NoisyPlan& np = *g.prog->noisy;
  const int P = job->n_symbols, C = c.n_channels, cols = c.param_cols();
  std::vector<int32_t> h(size_t(rows) * 3);
  std::vector<long long> off(rows, 0);
  for (int k = 0; k < rows; ++k) {
    const TrajRow& tr = list[r0 + k];
    h[k] = tr.group_pos;
    h[rows + k] = int32_t(tr.circuit + ctx->row_offset);
    h[2 * rows + k] = tr.traj;
    if (nb.d_given)
      off[k] = ((long long)tr.circuit * uniform_traj + tr.traj) * uniform_chan;
  }
  TFQB_CUDA(cudaMemcpyAsync(nb.d_sym_row, h.data(), sizeof(int32_t) * rows, cudaMemcpyHostToDevice, ctx->stream));
  TFQB_CUDA(cudaMemcpyAsync(nb.d_circuit, h.data() + rows, sizeof(int32_t) * rows, cudaMemcpyHostToDevice, ctx->stream));
  TFQB_CUDA(cudaMemcpyAsync(nb.d_traj, h.data() + 2 * rows, sizeof(int32_t) * rows, cudaMemcpyHostToDevice, ctx->stream));
  if (nb.d_given)
    TFQB_CUDA(cudaMemcpyAsync(nb.d_given_off, off.data(), sizeof(long long) * rows, cudaMemcpyHostToDevice, ctx->stream));
  TFQB_CUDA(cudaStreamSynchronize(ctx->stream));     // h / off are locals
  ctx->prof.h2d_bytes += int64_t(rows) * (12 + (nb.d_given ? 8 : 0));
  LaunchNoisyFillParams(nb.d_params, cols, P, C, job->d_params, nb.d_sym_row, nb.d_circuit,
                        nb.d_traj, nb.d_given, nb.d_given_off, seed, rows, ctx->stream);
  ctx->prof.kernel_launches++;
  const int n_alloc = np.segs[0]->host.n_alloc;
  const size_t row_stride = size_t(1) << n_alloc;
  for (size_t k = 0; k < np.segs.size(); ++k) {
    if (np.before[k].bit >= 0) {
      LaunchPopulation(job->d_psi, row_stride, n_alloc, np.before[k].bit, nb.d_pop, nb.d_params,
                       cols, np.before[k].col, rows, ctx->stream);
      ctx->prof.kernel_launches += 2;
    }
    TFQB_RETURN_IF(RunPlan(ctx, *np.segs[k], job->d_psi, nullptr, rows, nb.d_params, cols,
                           job->d_mats, k == 0, nullptr, 0, true));
  }
  return TFQB_OK;
}

int AllocNoisy(tfqb_job* job, NoisyBuffers* nb, int chunk, int max_cols, size_t max_state_amps,
               size_t max_mats, const float* uniforms, size_t given_count) {
  tfqb_context* ctx = job->ctx;
  nb->chunk = chunk;
  TFQB_RETURN_IF(job->Own(max_state_amps * size_t(chunk), &job->d_psi));
  TFQB_RETURN_IF(job->Own(max_mats * size_t(chunk), &job->d_mats));
  TFQB_RETURN_IF(job->Own(size_t(chunk) * max_cols, &nb->d_params));
  TFQB_RETURN_IF(job->Own(size_t(chunk), &nb->d_sym_row));
  TFQB_RETURN_IF(job->Own(size_t(chunk), &nb->d_circuit));
  TFQB_RETURN_IF(job->Own(size_t(chunk), &nb->d_traj));
  TFQB_RETURN_IF(job->Own(size_t(chunk), &nb->d_given_off));
  TFQB_RETURN_IF(job->Own(size_t(chunk) * 2, &nb->d_pop));
  if (uniforms && given_count) {
    TFQB_RETURN_IF(job->Own(given_count, &nb->d_given));
    TFQB_CUDA(cudaMemcpyAsync(nb->d_given, uniforms, given_count * sizeof(float),
                              cudaMemcpyHostToDevice, ctx->stream));
    TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->prof.h2d_bytes += int64_t(given_count * sizeof(float));
  }
  return TFQB_OK;
}

int CheckNumSamples(const int32_t* num_samples, int ns_rows, int ns_cols, int sum_rows,
                    int n_ops, int batch) {
  if (ns_rows != sum_rows)
    return Fail(TFQB_INVALID_ARGUMENT,
                "Dimension 0 of num_samples and pauli_sums do not match.Got " +
                    std::to_string(ns_rows) + " lists of sample sizes and " +
                    std::to_string(sum_rows) + " lists of pauli sums.");
  if (ns_cols != n_ops)
    return Fail(TFQB_INVALID_ARGUMENT,
                "Dimension 1 of num_samples and pauli_sums do not match.Got " +
                    std::to_string(ns_cols) + " lists of sample sizes and " +
                    std::to_string(n_ops) + " lists of pauli sums.");
  for (size_t k = 0; k < size_t(batch) * n_ops; ++k)
    if (num_samples[k] < 1)
      return Fail(TFQB_INVALID_ARGUMENT, "Each element of num_samples must be greater than 0.");
  return TFQB_OK;
}

// TfqNoisyExpectation (sampled = false) and TfqNoisySampledExpectation.
int NoisyExpectationImpl(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                         tfqb_strings pauli_sums, int sum_rows, int n_ops,
                         const int32_t* num_samples, int ns_rows, int ns_cols, uint64_t seed,
                         const float* uniforms, int uniform_traj, int uniform_chan,
                         bool sampled, float* out) {
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  auto jp = std::make_unique<tfqb_job>();
  tfqb_job* job = jp.get();
  job->kind = kJobNoisy;
  NoisyBuffers nb;
  // sampled: a rotated copy of the state, a probability tree, one index per row
  job->ctx = ctx;
  job->noisy = true;
  TFQB_RETURN_IF(BuildGroups(ctx, in, &pauli_sums, sum_rows, n_ops, job));
  TFQB_RETURN_IF(CheckNumSamples(num_samples, ns_rows, ns_cols, sum_rows, n_ops, job->batch));
  const int B = job->batch, M = n_ops, P = job->n_symbols;
  job->out_cols = M;
  TFQB_RETURN_IF(UploadPermuted(job, in->symbol_values, P, &job->d_params));
  TFQB_RETURN_IF(UploadTerms(job));
  // plans and sizes
  size_t max_row = 0, max_mats = 64, max_amps = 0, max_terms = 1;
  int max_cols = 1, max_traj = 0;
  std::vector<int> T(B, 0);
  for (int i = 0; i < B; ++i)
    for (int j = 0; j < M; ++j) T[i] = std::max(T[i], num_samples[size_t(i) * M + j]);
  for (auto& g : job->groups) {
    CompiledProgram& cp = *g.prog;
    if (cp.circuit.n == 0) continue;
    if (cp.circuit.n > kMaxDeviceQubits)
      return Fail(TFQB_RESOURCE_EXHAUSTED, "A " + std::to_string(cp.circuit.n) +
                                               "-qubit state does not fit in the device memory budget.");
    TFQB_RETURN_IF(EnsureNoisyPlan(ctx, cp));
    if (uniforms && cp.circuit.n_channels > uniform_chan)
      return Fail(TFQB_INVALID_ARGUMENT, "uniforms tensor holds too few channels");
    if (!sampled && !g.terms.empty()) {
      std::vector<TermMask> tm(g.terms.size());
      for (size_t k = 0; k < g.terms.size(); ++k)
        tm[k] = TermMask{g.terms[k].x, g.terms[k].z, g.terms[k].phase, g.terms[k].identity != 0};
      TFQB_RETURN_IF(CompileExpPlan(
          ctx, PlanExpectation(cp.circuit.n, tm, false, kTileMax, ExpLowBits()), &g.exp));
    }
    const int na = cp.noisy->segs[0]->host.n_alloc;
    max_amps = std::max(max_amps, size_t(1) << na);
    max_mats = std::max(max_mats, cp.noisy->mat_floats);
    max_cols = std::max(max_cols, cp.circuit.param_cols());
    max_terms = std::max(max_terms, g.terms.size());
    size_t per_row = (size_t(8) << na) * (sampled ? 2 : 1) + cp.noisy->mat_floats * 4 +
                     size_t(cp.circuit.param_cols()) * 4 + g.terms.size() * 16 + size_t(M) * 4 + 96;
    if (sampled) per_row += TreeDoublesPerRow(std::max(na, kMinStateBits)) * 8;
    max_row = std::max(max_row, per_row);
    for (int r : g.rows) {
      max_traj = std::max(max_traj, T[r]);
      if (uniforms && T[r] > uniform_traj)
        return Fail(TFQB_INVALID_ARGUMENT, "uniforms tensor holds too few trajectories");
    }
  }
  std::vector<double> acc(size_t(B) * M, 0.0);
  if (max_row > 0) {
    const size_t given = uniforms ? size_t(B) * uniform_traj * uniform_chan : 0;
    const size_t budget = Budget(ctx);
    if (budget <= given * 4 + max_row)
      return Fail(TFQB_RESOURCE_EXHAUSTED, "A " + std::to_string(job->nmax) +
                                               "-qubit state does not fit in the device memory budget.");
    size_t total_rows = 0;
    for (int i = 0; i < B; ++i) total_rows += size_t(T[i]);
    const int chunk = int(std::max<size_t>(
        1, std::min<size_t>({(budget - given * 4) / max_row, size_t(65535), total_rows})));
    TFQB_RETURN_IF(AllocNoisy(job, &nb, chunk, max_cols, max_amps, max_mats, uniforms, given));
    double* d_terms64 = nullptr;
    float* d_rowvals = nullptr;
    TFQB_RETURN_IF(job->Own(size_t(chunk) * max_terms, &d_terms64));
    TFQB_RETURN_IF(job->Own(size_t(chunk) * std::max(M, 1), &d_rowvals));
    double* d_u = nullptr;
    uint64_t* d_idx = nullptr;
    double* d_tree = nullptr;
    float* d_rot_mats = nullptr;
    if (sampled) {
      TFQB_RETURN_IF(job->Own(max_amps * size_t(chunk), &job->d_lam));
      TFQB_RETURN_IF(job->Own(size_t(chunk) * max_terms, &d_u));
      TFQB_RETURN_IF(job->Own(size_t(chunk), &d_idx));
      size_t tree = 0;
      for (auto& g : job->groups)
        if (g.prog->circuit.n)
          tree = std::max(tree, TreeDoublesPerRow(std::max(g.prog->noisy->segs[0]->host.n_alloc,
                                                           kMinStateBits)));
      TFQB_RETURN_IF(job->Own(size_t(chunk) * tree, &d_tree));
      TFQB_RETURN_IF(job->Own(64 * 8 + 64, &d_rot_mats));
    }
    std::vector<float> hvals;
    for (auto& g : job->groups) {
      const CircuitT& c = g.prog->circuit;
      if (c.n == 0) continue;
      const int nt = int(g.terms.size());
      const int na = g.prog->noisy->segs[0]->host.n_alloc;
      const size_t row_stride = size_t(1) << na;
      std::vector<TrajRow> list;
      for (size_t k = 0; k < g.rows.size(); ++k)
        for (int t = 0; t < T[g.rows[k]]; ++t)
          list.push_back(TrajRow{int32_t(g.begin + int(k)), int32_t(g.rows[k]), int32_t(t)});
      // Z-basis rotation plans of the sampled variant (util_qsim.h:230-239)
      std::vector<std::vector<std::unique_ptr<CompiledPlan>>> rot(M);
      if (sampled)
        for (int j = 0; j < M; ++j) {
          rot[j].resize(g.sums[j].terms.size());
          for (size_t t = 0; t < g.sums[j].terms.size(); ++t) {
            const PauliTermT& term = g.sums[j].terms[t];
            if (term.identity || term.rot.empty()) continue;
            TFQB_RETURN_IF(CompilePlan(ctx, PlanRotations(c.n, term.rot), &rot[j][t]));
          }
        }
      for (int r0 = 0; r0 < int(list.size()); r0 += chunk) {
        const int rows = std::min(chunk, int(list.size()) - r0);
        TFQB_RETURN_IF(RunTrajectories(job, g, list, r0, rows, nb, seed, uniform_traj, uniform_chan));
        if (!sampled) {
          if (nt > 0) {
            TFQB_CUDA(cudaMemsetAsync(d_terms64, 0, size_t(rows) * nt * sizeof(double), ctx->stream));
            TFQB_RETURN_IF(RunExpectationTerms(ctx, *g.exp, job->d_psi, rows, g.d_terms, nt, M, d_terms64));
          }
          LaunchCombineTerms(d_terms64, g.d_terms, nt, M, rows, d_rowvals, size_t(M), ctx->stream);
          ctx->prof.kernel_launches++;
        } else {
          TFQB_CUDA(cudaMemsetAsync(d_rowvals, 0, size_t(rows) * M * sizeof(float), ctx->stream));
          for (int j = 0; j < M; ++j) {
            const int ntj = int(g.sums[j].terms.size());
            if (ntj == 0) continue;
            // one shot per term: uniform (row, term k) = Philox counter
            // (k, circuit, trajectory, SAMPLE_STREAM + 1 + j)
            LaunchNoisyFillUniforms(d_u, size_t(ntj), ntj, nb.d_circuit, nb.d_traj,
                                    kSampleStream + 1u + uint32_t(j), seed, rows, ctx->stream);
            ctx->prof.kernel_launches++;
            for (int t = 0; t < ntj; ++t) {
              const PauliTermT& term = g.sums[j].terms[t];
              float* accp = d_rowvals + j;
              if (term.identity) {
                LaunchAddConstant(term.coeff, rows, accp, size_t(M), ctx->stream);
                ctx->prof.kernel_launches++;
                continue;
              }
              const float2* src = job->d_psi;
              if (rot[j][t]) {
                TFQB_CUDA(cudaMemcpyAsync(job->d_lam, job->d_psi, size_t(rows) * row_stride * sizeof(float2),
                                          cudaMemcpyDeviceToDevice, ctx->stream));
                TFQB_RETURN_IF(RunPlan(ctx, *rot[j][t], job->d_lam, nullptr, rows, nullptr, 0,
                                       d_rot_mats, false, nullptr));
                src = job->d_lam;
              }
              LaunchBuildTree(src, row_stride, na, d_tree, rows, ctx->stream);
              LaunchSample(src, row_stride, na, d_tree, d_u + t, size_t(ntj), nullptr, 1, rows,
                           d_idx, size_t(1), ctx->stream);
              LaunchParityExpectation(d_idx, size_t(1), term.parity_mask, term.coeff, nullptr, 1,
                                      rows, accp, size_t(M), ctx->stream);
              ctx->prof.kernel_launches += 3;
            }
          }
        }
        hvals.resize(size_t(rows) * M);
        if (M > 0)
          TFQB_CUDA(cudaMemcpyAsync(hvals.data(), d_rowvals, hvals.size() * sizeof(float),
                                    cudaMemcpyDeviceToHost, ctx->stream));
        TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->prof.d2h_bytes += int64_t(hvals.size() * sizeof(float));
        // rolling sums in double, trajectory order (tfq_noisy_expectation.cc:232-243)
        for (int k = 0; k < rows; ++k) {
          const TrajRow& tr = list[r0 + k];
          for (int j = 0; j < M; ++j)
            if (tr.traj < num_samples[size_t(tr.circuit) * M + j])
              acc[size_t(tr.circuit) * M + j] += double(hvals[size_t(k) * M + j]);
        }
      }
    }
  }
  for (auto& g : job->groups) {
    const bool empty = g.prog->circuit.n == 0;
    for (int r : g.rows)
      for (int j = 0; j < M; ++j)
        out[size_t(r) * M + j] =
            empty ? -2.0f   // (#679) tfq_noisy_expectation.cc:196-201
                  : float(acc[size_t(r) * M + j] / double(num_samples[size_t(r) * M + j]));
  }
  TFQB_CUDA(cudaGetLastError());
  return TFQB_OK;
}

}  // namespace

extern "C" {

static int impl_tfqb_noisy_expectation(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                                       tfqb_strings pauli_sums, int sum_rows, int n_ops,
                                       const int32_t* num_samples, int ns_rows, int ns_cols,
                                       uint64_t seed, const float* uniforms,
                                       int uniform_trajectories, int uniform_channels,
                                       float* expectations) {
  return NoisyExpectationImpl(ctx, in, pauli_sums, sum_rows, n_ops, num_samples, ns_rows, ns_cols,
                              seed, uniforms, uniform_trajectories, uniform_channels, false,
                              expectations);
}

static int impl_tfqb_noisy_sampled_expectation(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                                               tfqb_strings pauli_sums, int sum_rows, int n_ops,
                                               const int32_t* num_samples, int ns_rows,
                                               int ns_cols, uint64_t seed, const float* uniforms,
                                               int uniform_trajectories, int uniform_channels,
                                               float* expectations) {
  return NoisyExpectationImpl(ctx, in, pauli_sums, sum_rows, n_ops, num_samples, ns_rows, ns_cols,
                              seed, uniforms, uniform_trajectories, uniform_channels, true,
                              expectations);
}

static int impl_tfqb_noisy_samples_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                                           int num_samples, tfqb_job** job, int* max_qubits) {
  TFQB_RETURN_IF(CheckContext(ctx));
  if (num_samples < 0) return Fail(TFQB_INVALID_ARGUMENT, "num_samples must be >= 0");
  std::lock_guard<std::mutex> lock(ctx->mu);
  auto j = std::make_unique<tfqb_job>();
  j->ctx = ctx;
  j->kind = kJobNoisySamples;
  j->noisy = true;
  j->num_samples = num_samples;
  TFQB_RETURN_IF(BuildGroups(ctx, in, nullptr, 0, 0, j.get()));
  TFQB_RETURN_IF(UploadPermuted(j.get(), in->symbol_values, in->n_symbols, &j->d_params));
  for (auto& g : j->groups)
    if (g.prog->circuit.n) TFQB_RETURN_IF(EnsureNoisyPlan(ctx, *g.prog));
  if (max_qubits) *max_qubits = j->nmax;
  *job = j.release();
  return TFQB_OK;
}

static int impl_tfqb_noisy_samples_run(tfqb_job* job, uint64_t seed, const float* uniforms,
                                       int uniform_channels, const double* measure_uniforms,
                                       int8_t* samples) {
  if (!job || job->kind != kJobNoisySamples)
    return Fail(TFQB_INVALID_ARGUMENT, "not a noisy samples job");
  tfqb_context* ctx = job->ctx;
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  const int S = job->num_samples, B = job->batch, nmax = job->nmax;
  if (S == 0 || B == 0) return TFQB_OK;
  size_t max_row = 0, max_mats = 64, max_amps = 0, tree = 0;
  int max_cols = 1;
  for (auto& g : job->groups) {
    const CircuitT& c = g.prog->circuit;
    if (c.n == 0) continue;
    if (uniforms && c.n_channels > uniform_channels)
      return Fail(TFQB_INVALID_ARGUMENT, "uniforms tensor holds too few channels");
    const int na = g.prog->noisy->segs[0]->host.n_alloc;
    max_amps = std::max(max_amps, size_t(1) << na);
    max_mats = std::max(max_mats, g.prog->noisy->mat_floats);
    max_cols = std::max(max_cols, c.param_cols());
    tree = std::max(tree, TreeDoublesPerRow(std::max(na, kMinStateBits)));
    max_row = std::max(max_row, (size_t(8) << na) + g.prog->noisy->mat_floats * 4 +
                                    size_t(c.param_cols()) * 4 + tree * 8 + size_t(nmax) + 128);
  }
  for (auto& g : job->groups)
    if (g.prog->circuit.n == 0)
      for (int r : g.rows) memset(samples + size_t(r) * S * nmax, 0xFE, size_t(S) * nmax);  // -2
  if (max_row == 0) return TFQB_OK;
  const size_t given = uniforms ? size_t(B) * S * uniform_channels : 0;
  const size_t budget = Budget(ctx);
  if (budget <= given * 4 + max_row)
    return Fail(TFQB_RESOURCE_EXHAUSTED, "A " + std::to_string(nmax) +
                                             "-qubit state does not fit in the device memory budget.");
  const int chunk = int(std::max<size_t>(
      1, std::min<size_t>({(budget - given * 4) / max_row, size_t(65535), size_t(B) * S})));
  NoisyBuffers nb;
  TFQB_RETURN_IF(AllocNoisy(job, &nb, chunk, max_cols, max_amps, max_mats, uniforms, given));
  double* d_u = nullptr;
  uint64_t* d_idx = nullptr;
  double* d_tree = nullptr;
  int8_t* d_out8 = nullptr;
  TFQB_RETURN_IF(job->Own(size_t(chunk), &d_u));
  TFQB_RETURN_IF(job->Own(size_t(chunk), &d_idx));
  TFQB_RETURN_IF(job->Own(size_t(chunk) * tree, &d_tree));
  TFQB_RETURN_IF(job->Own(size_t(chunk) * std::max(nmax, 1), &d_out8));
  std::vector<double> hu;
  std::vector<int8_t> hout;
  for (auto& g : job->groups) {
    const CircuitT& c = g.prog->circuit;
    if (c.n == 0) continue;
    const int na = g.prog->noisy->segs[0]->host.n_alloc;
    const size_t row_stride = size_t(1) << na;
    std::vector<TrajRow> list;
    for (size_t k = 0; k < g.rows.size(); ++k)
      for (int t = 0; t < S; ++t)
        list.push_back(TrajRow{int32_t(g.begin + int(k)), int32_t(g.rows[k]), int32_t(t)});
    for (int r0 = 0; r0 < int(list.size()); r0 += chunk) {
      const int rows = std::min(chunk, int(list.size()) - r0);
      TFQB_RETURN_IF(RunTrajectories(job, g, list, r0, rows, nb, seed, S, uniform_channels));
      // the terminal measurement of every qubit: one shot per trajectory
      if (measure_uniforms) {
        hu.resize(rows);
        for (int k = 0; k < rows; ++k)
          hu[k] = measure_uniforms[size_t(list[r0 + k].circuit) * S + list[r0 + k].traj];
        TFQB_CUDA(cudaMemcpyAsync(d_u, hu.data(), sizeof(double) * rows, cudaMemcpyHostToDevice, ctx->stream));
        TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
      } else {
        LaunchNoisyFillUniforms(d_u, 1, 1, nb.d_circuit, nb.d_traj, kSampleStream, seed, rows, ctx->stream);
      }
      LaunchBuildTree(job->d_psi, row_stride, na, d_tree, rows, ctx->stream);
      LaunchSample(job->d_psi, row_stride, na, d_tree, d_u, size_t(1), nullptr, 1, rows, d_idx,
                   size_t(1), ctx->stream);
      LaunchUnpackSamples(d_idx, size_t(1), c.n, nmax, 1, rows, d_out8, ctx->stream);
      ctx->prof.kernel_launches += 4;
      hout.resize(size_t(rows) * nmax);
      if (nmax)
        TFQB_CUDA(cudaMemcpyAsync(hout.data(), d_out8, hout.size(), cudaMemcpyDeviceToHost, ctx->stream));
      TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
      ctx->prof.d2h_bytes += int64_t(hout.size());
      for (int k = 0; k < rows; ++k)
        memcpy(samples + (size_t(list[r0 + k].circuit) * S + list[r0 + k].traj) * nmax,
               hout.data() + size_t(k) * nmax, size_t(nmax));
    }
  }
  TFQB_CUDA(cudaGetLastError());
  return TFQB_OK;
}


// ---- unitary (next-row N4) -----------------------------------------------------
// TfqCalculateUnitaryOp::Compute (tfq_calculate_unitary_op.cc:47-164):
// unitary[i, j, k] = <j| U_i |k>, padded with (-2, 0) to 2^max_qubits.  Column k
// is the circuit applied to |k>: the 2^n basis states are the rows of one batch
// of the ordinary gate passes.
static int impl_tfqb_calculate_unitary_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                                               tfqb_job** job, int* max_qubits) {
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (in->symbol_rows != in->batch)       // this op's own wording (:60-64)
    return Fail(TFQB_INVALID_ARGUMENT,
                "Number of circuits and values do not match. Got " + std::to_string(in->batch) +
                    " circuits and " + std::to_string(in->symbol_rows) + " values.");
  auto j = std::make_unique<tfqb_job>();
  j->ctx = ctx;
  j->kind = kJobUnitary;
  TFQB_RETURN_IF(BuildGroups(ctx, in, nullptr, 0, 0, j.get()));
  if (j->nmax > 15)
    return Fail(TFQB_RESOURCE_EXHAUSTED, "A " + std::to_string(j->nmax) +
                                             "-qubit unitary does not fit in the device memory budget.");
  TFQB_RETURN_IF(UploadPermuted(j.get(), in->symbol_values, in->n_symbols, &j->d_params));
  for (auto& g : j->groups) {
    CompiledProgram& cp = *g.prog;
    if (cp.circuit.n == 0 || cp.fwd_any) continue;
    TFQB_RETURN_IF(CompilePlan(
        ctx, PlanForward(cp.circuit, kTileMax, GateLowBits(), true, false), &cp.fwd_any));
  }
  if (max_qubits) *max_qubits = j->nmax;
  *job = j.release();
  return TFQB_OK;
}

static int impl_tfqb_calculate_unitary_run(tfqb_job* job, float* unitary) {
  if (!job || job->kind != kJobUnitary) return Fail(TFQB_INVALID_ARGUMENT, "not a unitary job");
  tfqb_context* ctx = job->ctx;
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  const int P = job->n_symbols;
  const size_t D = size_t(1) << job->nmax;
  float2* out = reinterpret_cast<float2*>(unitary);
  if (job->batch == 0) return TFQB_OK;
  float2* d_out = nullptr;
  TFQB_RETURN_IF(job->Own(D * D, &d_out));
  size_t max_mats = 64;
  for (auto& g : job->groups)
    if (g.prog->circuit.n) max_mats = std::max(max_mats, size_t(g.prog->fwd_any->host.mat_floats));
  const size_t budget = Budget(ctx);
  for (auto& g : job->groups) {
    const CircuitT& c = g.prog->circuit;
    const size_t dim = c.n == 0 ? 1 : size_t(1) << c.n;
    for (size_t k = 0; k < g.rows.size(); ++k) {
      const int row = g.rows[k];
      float2* dst = out + size_t(row) * D * D;
      LaunchFillPad(d_out, D * D, ctx->stream);
      ctx->prof.kernel_launches++;
      if (c.n == 0) {
        // an empty program: the 1x1 identity (SetIdentity on a fresh unitary)
        const float2 one = make_float2(1.f, 0.f);
        TFQB_CUDA(cudaMemcpyAsync(d_out, &one, sizeof(one), cudaMemcpyHostToDevice, ctx->stream));
      } else {
        const CompiledPlan& plan = *g.prog->fwd_any;
        const size_t row_stride = size_t(1) << plan.host.n_alloc;
        const size_t per_col = row_stride * sizeof(float2) + max_mats * 4;
        const int chunk = int(std::max<size_t>(1, std::min<size_t>({budget / per_col, dim, size_t(65535)})));
        if (!job->d_psi) {
          size_t max_stride = 0;
          for (auto& gg : job->groups)
            if (gg.prog->circuit.n)
              max_stride = std::max(max_stride, size_t(1) << gg.prog->fwd_any->host.n_alloc);
          const int cap = int(std::max<size_t>(1, std::min<size_t>({budget / (max_stride * 8 + max_mats * 4), D, size_t(65535)})));
          TFQB_RETURN_IF(job->Own(max_stride * size_t(cap), &job->d_psi));
          TFQB_RETURN_IF(job->Own(max_mats * size_t(cap), &job->d_mats));
          TFQB_RETURN_IF(job->Own(size_t(cap) * std::max(P, 1), &job->d_down));   // parameter rows
          job->chunk_cap = cap;
        }
        const int per = std::min(chunk, job->chunk_cap);
        // every column of this row shares the row's symbol values: one
        // parameter row, repeated (row-dependent matrices index params by row)
        float* d_prow = nullptr;
        if (plan.host.row_dependent && P > 0) {
          d_prow = job->d_down;
          for (int r = 0; r < per; ++r)
            TFQB_CUDA(cudaMemcpyAsync(d_prow + size_t(r) * P, job->d_params + size_t(g.begin + k) * P,
                                      sizeof(float) * P, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        for (size_t k0 = 0; k0 < dim; k0 += size_t(per)) {
          const int cols = int(std::min<size_t>(size_t(per), dim - k0));
          LaunchBasisStates(job->d_psi, row_stride, k0, cols, ctx->stream);
          TFQB_RETURN_IF(RunPlan(ctx, plan, job->d_psi, nullptr, cols,
                                 d_prow ? d_prow : job->d_params + size_t(g.begin + k) * P, P,
                                 job->d_mats, false, nullptr, 0));
          LaunchExportUnitary(job->d_psi, row_stride, dim, k0, cols, d_out, D, ctx->stream);
          ctx->prof.kernel_launches += 2;
        }
      }
      TFQB_CUDA(cudaMemcpyAsync(dst, d_out, D * D * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
      TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
      ctx->prof.d2h_bytes += int64_t(D * D * sizeof(float2));
    }
  }
  TFQB_CUDA(cudaGetLastError());
  return TFQB_OK;
}

// ---- inner product (N1) -----------------------------------------------------
static int impl_tfqb_inner_product(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                       tfqb_strings other_programs, int other_rows,
                       int n_other, float* inner_products) {
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  auto jp = std::make_unique<tfqb_job>();
  tfqb_job* job = jp.get();
  job->ctx = ctx;
  job->kind = kJobState;
  TFQB_RETURN_IF(BuildGroups(ctx, in, nullptr, 0, 0, job));
  if (other_rows != in->batch)
    return Fail(TFQB_INVALID_ARGUMENT,
                "programs and other_programs batch dimension do not match. Foud: " +
                    std::to_string(in->batch) + " and " + std::to_string(other_rows));
  const int B = in->batch, K = n_other, P = job->n_symbols;
  TFQB_RETURN_IF(UploadPermuted(job, in->symbol_values, P, &job->d_params));
  // one extra single-row buffer for phi: account it as extra bytes per row of
  // the smallest chunk (conservative)
  TFQB_RETURN_IF(PlanAndSize(job, false, 1, 0, nullptr));
  // lower every paired program against its row's qubit map (cached by bytes)
  struct Paired { CircuitT circuit; std::unique_ptr<CompiledPlan> plan; };
  std::map<std::pair<CompiledProgram*, std::string>, std::unique_ptr<Paired>> paired;
  std::vector<Paired*> of(size_t(B) * K, nullptr);
  size_t max_state = 0, max_mats = 64;
  for (auto& g : job->groups) {
    if (g.prog->circuit.n == 0) continue;
    max_state = std::max(max_state, size_t(1) << g.prog->fwd->host.n_alloc);
    for (int r : g.rows) {
      for (int j = 0; j < K; ++j) {
        const size_t k = size_t(r) * K + j;
        auto key = std::make_pair(g.prog.get(),
                                  std::string(other_programs.data[k], other_programs.size[k]));
        auto it = paired.find(key);
        if (it == paired.end()) {
          ProgramPB pb;
          if (!ParseProgram(other_programs.data[k], other_programs.size[k], &pb))
            return Fail(TFQB_INVALID_ARGUMENT, "Unparseable proto: " + key.second.substr(0, 64));
          auto pp = std::make_unique<Paired>();
          Status st = LowerPairedProgram(pb, g.prog->circuit, &pp->circuit);
          if (!st.ok) return Fail(TFQB_INVALID_ARGUMENT, st.msg);
          TFQB_RETURN_IF(CompilePlan(
              ctx, PlanForward(pp->circuit, kTileMax, kLowBits, true),
              &pp->plan));
          max_mats = std::max(max_mats, size_t(pp->plan->host.mat_floats));
          it = paired.emplace(std::move(key), std::move(pp)).first;
        }
        of[k] = it->second.get();
      }
    }
  }
  float2* d_phi = nullptr;
  float* d_pmats = nullptr;
  double* d_ip = nullptr;
  TFQB_RETURN_IF(job->Own(std::max<size_t>(max_state, 32), &d_phi));
  TFQB_RETURN_IF(job->Own(max_mats, &d_pmats));
  TFQB_RETURN_IF(job->Own(std::max<size_t>(size_t(job->chunk_cap) * 2, 2), &d_ip));
  std::vector<double> hip;
  for (auto& g : job->groups) {
    if (g.prog->circuit.n == 0) {   // (#679): <empty|anything> = 1
      for (int r : g.rows)
        for (int j = 0; j < K; ++j) {
          inner_products[(size_t(r) * K + j) * 2] = 1.f;
          inner_products[(size_t(r) * K + j) * 2 + 1] = 0.f;
        }
      continue;
    }
    const CompiledPlan& fwd = *g.prog->fwd;
    const int na = fwd.host.n_alloc;
    const size_t row_stride = size_t(1) << na;
    const int per = g.chunk;
    for (int c0 = 0; c0 < int(g.rows.size()); c0 += per) {
      const int rows = std::min(per, int(g.rows.size()) - c0);
      const int r0 = g.begin + c0;
      TFQB_RETURN_IF(RunPlan(ctx, fwd, job->d_psi, nullptr, rows,
                             job->d_params + size_t(r0) * P, P, job->d_mats,
                             true, nullptr, 0));
      for (int j = 0; j < K; ++j) {
        // consecutive rows that pair with the same circuit share one phi
        int k0 = 0;
        while (k0 < rows) {
          Paired* pp = of[size_t(g.rows[c0 + k0]) * K + j];
          int k1 = k0 + 1;
          while (k1 < rows && of[size_t(g.rows[c0 + k1]) * K + j] == pp) ++k1;
          TFQB_RETURN_IF(RunPlan(ctx, *pp->plan, d_phi, nullptr, 1, nullptr, 0,
                                 d_pmats, true, nullptr, 0));
          TFQB_CUDA(cudaMemsetAsync(d_ip, 0, size_t(k1 - k0) * 2 * sizeof(double),
                                    ctx->stream));
          LaunchInnerProduct(job->d_psi + size_t(k0) * row_stride, row_stride, d_phi,
                             na, k1 - k0, d_ip, ctx->stream);
          ctx->prof.kernel_launches++;
          hip.resize(size_t(k1 - k0) * 2);
          TFQB_CUDA(cudaMemcpyAsync(hip.data(), d_ip, hip.size() * sizeof(double),
                                    cudaMemcpyDeviceToHost, ctx->stream));
          TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
          for (int k = k0; k < k1; ++k) {
            const size_t o = (size_t(g.rows[c0 + k]) * K + j) * 2;
            inner_products[o] = float(hip[size_t(k - k0) * 2]);
            inner_products[o + 1] = float(hip[size_t(k - k0) * 2 + 1]);
          }
          k0 = k1;
        }
      }
    }
  }
  TFQB_CUDA(cudaGetLastError());
  return TFQB_OK;
}

// TfqInnerProductGrad (math_ops/tfq_inner_product_grad.cc:46-501):
//   grads[i, p] = sum over the gradient gates of symbol p of <dG psi' | lam>,
//   lam = sum_j downstream[i, j] |phi_ij> rewound together with psi.
// The reverse sweep of the adjoint op yields 2 Re<lam| dG |psi'> per gate; the
// complex inner product is recovered from two sweeps, one with lam and one
// with i * lam (Re<i lam| x> = Im<lam| x>): same kernels, twice the work.
static int impl_tfqb_inner_product_grad(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                            tfqb_strings other_programs, int other_rows,
                            int n_other, const float* downstream, int grad_rows,
                            int grad_cols, float* grads) {
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (in->n_symbols <= 0)
    return Fail(TFQB_INVALID_ARGUMENT,
                "The number of symbols must be a positive integer, got 0 symbols.");
  auto jp = std::make_unique<tfqb_job>();
  tfqb_job* job = jp.get();
  job->ctx = ctx;
  job->kind = kJobAdjoint;
  TFQB_RETURN_IF(BuildGroups(ctx, in, nullptr, 0, 0, job));
  if (other_rows != in->batch)
    return Fail(TFQB_INVALID_ARGUMENT,
                "programs and other_programs batch dimension do not match. Foud: " +
                    std::to_string(in->batch) + " and " + std::to_string(other_rows));
  if (grad_rows != in->batch)
    return Fail(TFQB_INVALID_ARGUMENT,
                "Number of gradients and circuits do not match. Got " +
                    std::to_string(grad_rows) + " gradients and " +
                    std::to_string(in->batch) + " circuits.");
  if (grad_cols != n_other)
    return Fail(TFQB_INVALID_ARGUMENT,
                "Number of gradients and other_programs do not match. Got " +
                    std::to_string(grad_cols) + " gradient entries and " +
                    std::to_string(n_other) + " other programs.");
  const int B = in->batch, K = n_other, P = job->n_symbols;
  for (size_t i = 0; i < size_t(B) * P * 2; ++i) grads[i] = 0.f;
  TFQB_RETURN_IF(UploadPermuted(job, in->symbol_values, P, &job->d_params));
  TFQB_RETURN_IF(UploadPermuted(job, downstream, K, &job->d_down));
  TFQB_RETURN_IF(PlanAndSize(job, true, 2, 0, AdjScratch));
  struct Paired { CircuitT circuit; std::unique_ptr<CompiledPlan> plan; };
  std::map<std::pair<CompiledProgram*, std::string>, std::unique_ptr<Paired>> paired;
  std::vector<Paired*> of(size_t(B) * K, nullptr);
  size_t max_state = 0, max_mats = 64;
  for (auto& g : job->groups) {
    if (g.prog->circuit.n == 0) continue;
    max_state = std::max(max_state, size_t(1) << g.prog->fwd->host.n_alloc);
    for (int r : g.rows) {
      for (int j = 0; j < K; ++j) {
        const size_t k = size_t(r) * K + j;
        auto key = std::make_pair(g.prog.get(),
                                  std::string(other_programs.data[k], other_programs.size[k]));
        auto it = paired.find(key);
        if (it == paired.end()) {
          ProgramPB pb;
          if (!ParseProgram(other_programs.data[k], other_programs.size[k], &pb))
            return Fail(TFQB_INVALID_ARGUMENT, "Unparseable proto: " + key.second.substr(0, 64));
          auto pp = std::make_unique<Paired>();
          Status st = LowerPairedProgram(pb, g.prog->circuit, &pp->circuit);
          if (!st.ok) return Fail(TFQB_INVALID_ARGUMENT, st.msg);
          TFQB_RETURN_IF(CompilePlan(
              ctx, PlanForward(pp->circuit, kTileMax, kLowBits, true),
              &pp->plan));
          max_mats = std::max(max_mats, size_t(pp->plan->host.mat_floats));
          it = paired.emplace(std::move(key), std::move(pp)).first;
        }
        of[k] = it->second.get();
      }
    }
  }
  float2* d_phi = nullptr;
  float* d_pmats = nullptr;
  TFQB_RETURN_IF(job->Own(std::max<size_t>(max_state, 32), &d_phi));
  TFQB_RETURN_IF(job->Own(max_mats, &d_pmats));
  std::vector<double> part[2];
  for (auto& g : job->groups) {
    if (g.prog->circuit.n == 0) continue;     // empty circuit: the row stays 0
    const CompiledPlan& fwd = *g.prog->fwd;
    const CompiledPlan& adj = *g.prog->adj;
    const int na = fwd.host.n_alloc;
    const size_t row_stride = size_t(1) << na;
    const int ns = int(adj.host.grad_slots.size());
    if (ns == 0) continue;
    const int per = g.chunk;
    for (int c0 = 0; c0 < int(g.rows.size()); c0 += per) {
      const int rows = std::min(per, int(g.rows.size()) - c0);
      const int r0 = g.begin + c0;
      const float* params = job->d_params + size_t(r0) * P;
      for (int im = 0; im < 2; ++im) {
        TFQB_RETURN_IF(RunPlan(ctx, fwd, job->d_psi, nullptr, rows, params, P,
                               job->d_mats, true, nullptr, 0));
        if (K == 0)
          TFQB_CUDA(cudaMemsetAsync(job->d_lam, 0, size_t(rows) * row_stride * sizeof(float2),
                                    ctx->stream));
        for (int j = 0; j < K; ++j) {
          int k0 = 0;
          while (k0 < rows) {     // consecutive rows pairing with the same circuit share phi
            Paired* pp = of[size_t(g.rows[c0 + k0]) * K + j];
            int k1 = k0 + 1;
            while (k1 < rows && of[size_t(g.rows[c0 + k1]) * K + j] == pp) ++k1;
            TFQB_RETURN_IF(RunPlan(ctx, *pp->plan, d_phi, nullptr, 1, nullptr, 0,
                                   d_pmats, true, nullptr, 0));
            LaunchAxpyRows(job->d_lam + size_t(k0) * row_stride, row_stride, d_phi, na,
                           job->d_down + size_t(r0 + k0) * K + j, K, im == 1, j == 0,
                           k1 - k0, ctx->stream);
            ctx->prof.kernel_launches++;
            k0 = k1;
          }
        }
        TFQB_CUDA(cudaMemsetAsync(job->d_scratch64, 0,
                                  size_t(rows) * ns * sizeof(double), ctx->stream));
        TFQB_RETURN_IF(RunPlan(ctx, adj, job->d_psi, job->d_lam, rows, params, P,
                               job->d_mats, false, job->d_scratch64));
        part[im].resize(size_t(rows) * ns);
        TFQB_CUDA(cudaMemcpyAsync(part[im].data(), job->d_scratch64,
                                  part[im].size() * sizeof(double),
                                  cudaMemcpyDeviceToHost, ctx->stream));
        TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
      }
      // slot s = 2 Re<lam| dG |psi'> (and 2 Im for the i*lam sweep);
      // <dG psi'| lam> is its conjugate.  Slots are in the reference's
      // accumulation order (last gate first), summed in complex<float>.
      for (int k = 0; k < rows; ++k) {
        float* dst = grads + size_t(g.rows[c0 + k]) * P * 2;
        for (int s2 = 0; s2 < ns; ++s2) {
          const int col = adj.host.grad_slots[s2].symbol_col;
          if (col < 0 || col >= P) continue;
          dst[2 * col] += float(0.5 * part[0][size_t(k) * ns + s2]);
          dst[2 * col + 1] += float(-0.5 * part[1][size_t(k) * ns + s2]);
        }
      }
    }
  }
  TFQB_CUDA(cudaGetLastError());
  return TFQB_OK;
}

// ---- sharded single state ---------------------------------------------------
static int impl_tfqb_sharded_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                         tfqb_strings pauli_sums, int n_ops, int world,
                         int rank, tfqb_job** job, int* n_stages, int* n_terms) {
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (in->batch != 1)
    return Fail(TFQB_INVALID_ARGUMENT, "sharded simulation takes exactly one program");
  int g = 0;
  while ((1 << g) < world) ++g;
  if (world < 1 || (1 << g) != world || rank < 0 || rank >= world)
    return Fail(TFQB_INVALID_ARGUMENT, "world must be a power of two and 0 <= rank < world");
  auto jp = std::make_unique<tfqb_job>();
  tfqb_job* j = jp.get();
  j->ctx = ctx;
  j->kind = kJobSharded;
  TFQB_RETURN_IF(BuildGroups(ctx, in, &pauli_sums, 1, n_ops, j));
  j->out_cols = n_ops;
  Group& grp = j->groups[0];
  const CircuitT& c = grp.prog->circuit;
  if (c.n == 0)
    return Fail(TFQB_INVALID_ARGUMENT, "sharded simulation of an empty program");
 if (c.n - g > kMaxDeviceQubits)
    return Fail(TFQB_RESOURCE_EXHAUSTED,
                "A " + std::to_string(c.n) + "-qubit state sharded over " +
                    std::to_string(world) + " ranks does not fit in device memory.");
  if (c.n - g < std::max(kMinStateBits, 2 * g + 2))
    return Fail(TFQB_INVALID_ARGUMENT,
                "too few qubits (" + std::to_string(c.n) + ") to shard over " +
                    std::to_string(world) + " ranks");
  auto st = std::make_unique<ShardedState>();
  st->rank = rank;
  st->world = world;
  std::vector<TermMask> tm(grp.terms.size());
  for (size_t k = 0; k < grp.terms.size(); ++k)
    tm[k] = TermMask{grp.terms[k].x, grp.terms[k].z, grp.terms[k].phase,
                     grp.terms[k].identity != 0};
  st->plan = PlanSharded(c, g, tm);
  st->n_terms = int(grp.terms.size());
  size_t mat_floats = 64;
  for (auto& gp : st->plan.gate_plans) mat_floats = std::max(mat_floats, size_t(gp.mat_floats));
  {
    std::string key = std::to_string(world) + "|" + std::to_string(rank) + "|";
    for (const TermMask& t : tm)
      key += std::to_string(t.x) + "," + std::to_string(t.z) + "," + std::to_string(t.phase) +
             (t.identity ? "i;" : ";");
    auto& cached = grp.prog->sharded_gates[key];
    if (cached.size() != st->plan.gate_plans.size()) {
      cached.clear();
      for (auto& gp : st->plan.gate_plans) {
        std::unique_ptr<CompiledPlan> cp;
        DevicePlan copy = gp;
        TFQB_RETURN_IF(CompilePlan(ctx, std::move(copy), &cp));
        cached.push_back(std::shared_ptr<CompiledPlan>(std::move(cp)));
      }
    }
    st->gates = cached;
  }
  for (auto& ep : st->plan.exp_plans) {
    if (!ep.generic_terms.empty())
      return Fail(TFQB_INVALID_ARGUMENT,
                  "sharded expectation supports Pauli terms with at most 4 X/Y "
                  "factors");
    std::unique_ptr<CompiledExpPlan> cp;
    ExpectationPlan copy = ep;
    TFQB_RETURN_IF(CompileExpPlan(ctx, std::move(copy), &cp));
    st->exps.push_back(std::move(cp));
  }
  TFQB_RETURN_IF(UploadTerms(j));
  TFQB_RETURN_IF(UploadPermuted(j, in->symbol_values, in->n_symbols, &j->d_params));
  const size_t amps = size_t(1) << st->plan.n_local;
  TFQB_RETURN_IF(j->Own(amps, &st->buf[0]));
  if (st->plan.n_exchanges > 0) TFQB_RETURN_IF(j->Own(amps, &st->buf[1]));
  TFQB_RETURN_IF(j->Own(mat_floats, &j->d_mats));
  TFQB_RETURN_IF(j->Own(std::max<size_t>(st->n_terms, 1), &st->d_per_term));
  TFQB_CUDA(cudaMemsetAsync(st->d_per_term, 0,
                            std::max<size_t>(st->n_terms, 1) * sizeof(double), ctx->stream));
  TFQB_RETURN_IF(j->Own(std::max<size_t>(st->n_terms, 1), &st->d_total));
  TFQB_RETURN_IF(j->Own(16, &st->d_error));
  TFQB_CUDA(cudaMemsetAsync(st->d_error, 0, 16 * sizeof(int), ctx->stream));
  if (n_ops == 0) {     // a sampling job
    TFQB_RETURN_IF(j->Own(TreeDoublesPerRow(std::max(st->plan.n_local, kMinStateBits)), &st->d_tree));
    TFQB_RETURN_IF(j->Own(size_t(world) + 2, &st->d_norms));
    TFQB_RETURN_IF(j->Own(size_t(kShardedMaxShots), &st->d_u));
    TFQB_RETURN_IF(j->Own(size_t(kShardedMaxShots), &st->d_idx));
    TFQB_RETURN_IF(j->Own(4, &st->d_rowid));
  }
  // the flag block is its own cudaMalloc: peers map it whole (CUDA IPC)
  st->flag_bytes = 64 + std::max<size_t>(st->n_terms, 1) * sizeof(double);
  {
    void* fb = nullptr;
    TFQB_CUDA(cudaMalloc(&fb, st->flag_bytes));
    st->flag_block = static_cast<unsigned char*>(fb);
    TFQB_CUDA(cudaMemsetAsync(fb, 0, st->flag_bytes, ctx->stream));
    TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (n_stages) *n_stages = int(st->plan.stages.size());
  if (n_terms) *n_terms = st->n_terms;
  j->sharded = std::move(st);
  *job = jp.release();
  return TFQB_OK;
}

static int impl_tfqb_sharded_stage_kind(tfqb_job* job, int stage) {
  if (!job || !job->sharded || stage < 0 ||
      stage >= int(job->sharded->plan.stages.size()))
    return -1;
  return job->sharded->plan.stages[stage].kind;
}

static int impl_tfqb_sharded_buffers(tfqb_job* job, void** send, void** recv, size_t* bytes) {
  if (!job || !job->sharded) return Fail(TFQB_INVALID_ARGUMENT, "not a sharded job");
  ShardedState& st = *job->sharded;
  if (send) *send = st.buf[st.cur];
  if (recv) *recv = st.buf[st.cur ^ 1];
  if (bytes) *bytes = (size_t(1) << st.plan.n_local) * sizeof(float2);
  return TFQB_OK;
}

static int impl_tfqb_sharded_run_stage(tfqb_job* job, int stage) {
  if (!job || !job->sharded) return Fail(TFQB_INVALID_ARGUMENT, "not a sharded job");
  tfqb_context* ctx = job->ctx;
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  ShardedState& st = *job->sharded;
  if (stage < 0 || stage >= int(st.plan.stages.size()))
    return Fail(TFQB_INVALID_ARGUMENT, "stage out of range");
  const ShardedStage sg = st.plan.stages[stage];
  const unsigned long long rank_base =
      (unsigned long long)st.rank << st.plan.n_local;
  const size_t amps = size_t(1) << st.plan.n_local;
  auto ensure_state = [&]() -> int {
    if (st.state_ready) return TFQB_OK;
    // |0...0>: amplitude 1 at global index 0, which lives on rank 0
    TFQB_CUDA(cudaMemsetAsync(st.buf[st.cur], 0, amps * sizeof(float2), ctx->stream));
    if (st.rank == 0) {
      const float2 one = make_float2(1.f, 0.f);
      TFQB_CUDA(cudaMemcpyAsync(st.buf[st.cur], &one, sizeof(one),
                                cudaMemcpyHostToDevice, ctx->stream));
      TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    st.state_ready = true;
    return TFQB_OK;
  };
  if (sg.kind == 0) {
    const CompiledPlan& cp = *st.gates[sg.index];
    const bool init = !st.state_ready && !cp.host.passes.empty();
    if (!init) TFQB_RETURN_IF(ensure_state());
    TFQB_RETURN_IF(RunPlan(ctx, cp, st.buf[st.cur], nullptr, 1, job->d_params,
                           job->n_symbols, job->d_mats, init, nullptr, rank_base));
    st.state_ready = true;
  } else if (sg.kind == 1) {
    st.cur ^= 1;   // the host filled the alternate buffer by all-to-all
  } else {
    TFQB_RETURN_IF(ensure_state());
    TFQB_RETURN_IF(RunExpectationTerms(ctx, *st.exps[sg.index], st.buf[st.cur], 1,
                                       job->groups[0].d_terms, st.n_terms,
                                       job->n_ops, st.d_per_term, rank_base));
  }
  TFQB_CUDA(cudaGetLastError());
  return TFQB_OK;
}

static int impl_tfqb_sharded_partials(tfqb_job* job, double* per_term) {
  if (!job || !job->sharded) return Fail(TFQB_INVALID_ARGUMENT, "not a sharded job");
  tfqb_context* ctx = job->ctx;
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  ShardedState& st = *job->sharded;
  if (st.n_terms)
    TFQB_CUDA(cudaMemcpyAsync(per_term, st.d_per_term, sizeof(double) * st.n_terms,
                              cudaMemcpyDeviceToHost, ctx->stream));
  TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
  return TFQB_OK;
}

static int impl_tfqb_sharded_finish(tfqb_job* job, const double* per_term_total,
                        float* expectations) {
  if (!job || !job->sharded) return Fail(TFQB_INVALID_ARGUMENT, "not a sharded job");
  const Group& g = job->groups[0];
  // same float accumulation as combine_terms_kernel (util_qsim.h:154-185)
  for (int j = 0; j < job->n_ops; ++j) {
    float e = 0.f;
    for (size_t t = 0; t < g.terms.size(); ++t) {
      const DevTerm& term = g.terms[t];
      if (term.op != j) continue;
      if (term.identity) e = e + term.coeff;
      else e = float(double(e) + double(term.coeff) * per_term_total[t]);
    }
    expectations[j] = e;
  }
  return TFQB_OK;
}


// ---- sharded single state: exchange through peer memory ----------------------
// What a rank publishes about itself (TFQB_PEER_HANDLE_BYTES = 256).
struct PeerBlob {
  cudaIpcMemHandle_t h_buf[2];     // 2 x 64 bytes
  cudaIpcMemHandle_t h_flags;      // 64 bytes
  uint64_t pid;
  uint64_t raw_buf[2];             // valid inside the publishing process
  uint64_t raw_flags;
  int32_t device;
  int32_t has_buf1;
  uint64_t shard_bytes;
};
static_assert(sizeof(PeerBlob) <= 256, "PeerBlob must fit TFQB_PEER_HANDLE_BYTES");

// TFQB_FUSED_EXCHANGE=0: always the stand-alone pull kernel (A/B measurement)
static bool FusedExchangeEnabled() {
  static const bool v = [] {
    const char* e = getenv("TFQB_FUSED_EXCHANGE");
    return !(e && *e == '0');
  }();
  return v;
}

static unsigned long long PeerTimeoutNs() {
  const char* e = getenv("TFQB_PEER_TIMEOUT_S");
  const double sec = e && *e ? atof(e) : 30.0;
  return (unsigned long long)(sec * 1e9);
}

static int impl_tfqb_sharded_export(tfqb_job* job, unsigned char* handle) {
  if (!job || !job->sharded || !handle) return Fail(TFQB_INVALID_ARGUMENT, "not a sharded job");
  tfqb_context* ctx = job->ctx;
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  ShardedState& st = *job->sharded;
  PeerBlob b;
  memset(&b, 0, sizeof b);
  b.pid = uint64_t(getpid());
  b.device = ctx->device;
  b.shard_bytes = (uint64_t(1) << st.plan.n_local) * sizeof(float2);
  TFQB_CUDA(cudaIpcGetMemHandle(&b.h_buf[0], st.buf[0]));
  b.raw_buf[0] = uint64_t(reinterpret_cast<uintptr_t>(st.buf[0]));
  if (st.buf[1]) {
    TFQB_CUDA(cudaIpcGetMemHandle(&b.h_buf[1], st.buf[1]));
    b.raw_buf[1] = uint64_t(reinterpret_cast<uintptr_t>(st.buf[1]));
    b.has_buf1 = 1;
  }
  TFQB_CUDA(cudaIpcGetMemHandle(&b.h_flags, st.flag_block));
  b.raw_flags = uint64_t(reinterpret_cast<uintptr_t>(st.flag_block));
  memset(handle, 0, 256);
  memcpy(handle, &b, sizeof b);
  return TFQB_OK;
}

static int impl_tfqb_sharded_connect(tfqb_job* job, const unsigned char* handles, int world) {
  if (!job || !job->sharded || !handles) return Fail(TFQB_INVALID_ARGUMENT, "not a sharded job");
  tfqb_context* ctx = job->ctx;
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  ShardedState& st = *job->sharded;
  if (world != st.world) return Fail(TFQB_INVALID_ARGUMENT, "world does not match the job");
  if (st.connected) return Fail(TFQB_INVALID_ARGUMENT, "sharded job is already connected");
  std::vector<const float2*> pb[2];
  std::vector<unsigned*> ready(world), done(world);
  std::vector<double*> parts(world);
  pb[0].resize(world, nullptr);
  pb[1].resize(world, nullptr);
  const uint64_t my_bytes = (uint64_t(1) << st.plan.n_local) * sizeof(float2);
  auto open = [&](const cudaIpcMemHandle_t& h, void** out) -> int {
    cudaError_t e = cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return Fail(TFQB_INTERNAL, std::string("cudaIpcOpenMemHandle failed: ") +
                                     cudaGetErrorString(e) +
                                     " (ranks must be on NVLink/P2P-capable GPUs of one node)");
    }
    st.ipc_opened.push_back(*out);
    return TFQB_OK;
  };
  for (int r = 0; r < world; ++r) {
    PeerBlob b;
    memcpy(&b, handles + size_t(r) * 256, sizeof b);
    if (b.shard_bytes != my_bytes)
      return Fail(TFQB_INVALID_ARGUMENT, "rank " + std::to_string(r) + " holds a different shard size");
    unsigned char* flags = nullptr;
    if (r == st.rank) {
      pb[0][r] = st.buf[0];
      pb[1][r] = st.buf[1];
      flags = st.flag_block;
    } else if (b.pid == uint64_t(getpid())) {
      // another rank of this process: its pointers are ours too
      if (b.device != ctx->device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, ctx->device, b.device);
        if (!can) return Fail(TFQB_UNAVAILABLE, "no peer access between the GPUs of the ranks");
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return Fail(TFQB_INTERNAL, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
      }
      pb[0][r] = reinterpret_cast<const float2*>(uintptr_t(b.raw_buf[0]));
      pb[1][r] = reinterpret_cast<const float2*>(uintptr_t(b.raw_buf[1]));
      flags = reinterpret_cast<unsigned char*>(uintptr_t(b.raw_flags));
    } else {
      void* p = nullptr;
      TFQB_RETURN_IF(open(b.h_buf[0], &p));
      pb[0][r] = static_cast<const float2*>(p);
      if (b.has_buf1) {
        TFQB_RETURN_IF(open(b.h_buf[1], &p));
        pb[1][r] = static_cast<const float2*>(p);
      }
      TFQB_RETURN_IF(open(b.h_flags, &p));
      flags = static_cast<unsigned char*>(p);
    }
    ready[r] = reinterpret_cast<unsigned*>(flags);
    done[r] = reinterpret_cast<unsigned*>(flags) + 1;
    parts[r] = reinterpret_cast<double*>(flags + 64);
  }
  // device tables
  void* tab = nullptr;
  const size_t tb = size_t(world) * 8;
  TFQB_RETURN_IF(job->Own(5 * tb + 64, reinterpret_cast<unsigned char**>(&tab)));
  unsigned char* base = static_cast<unsigned char*>(tab);
  std::vector<unsigned char> host(5 * tb);
  memcpy(host.data(), pb[0].data(), tb);
  memcpy(host.data() + tb, pb[1].data(), tb);
  memcpy(host.data() + 2 * tb, ready.data(), tb);
  memcpy(host.data() + 3 * tb, done.data(), tb);
  memcpy(host.data() + 4 * tb, parts.data(), tb);
  TFQB_CUDA(cudaMemcpyAsync(base, host.data(), host.size(), cudaMemcpyHostToDevice, ctx->stream));
  TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
  st.d_peer_buf[0] = reinterpret_cast<const float2**>(base);
  st.d_peer_buf[1] = reinterpret_cast<const float2**>(base + tb);
  st.d_peer_ready = reinterpret_cast<unsigned**>(base + 2 * tb);
  st.d_peer_done = reinterpret_cast<unsigned**>(base + 3 * tb);
  st.d_peer_parts = reinterpret_cast<double**>(base + 4 * tb);
  st.connected = true;
  return TFQB_OK;
}

// Enqueue every stage of the job on the context stream, exchanges included;
// returns without waiting.  All ranks must call it the same number of times.
static int impl_tfqb_sharded_enqueue(tfqb_job* job) {
  if (!job || !job->sharded) return Fail(TFQB_INVALID_ARGUMENT, "not a sharded job");
  tfqb_context* ctx = job->ctx;
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  ShardedState& st = *job->sharded;
  if (st.world > 1 && !st.connected)
    return Fail(TFQB_INVALID_ARGUMENT, "tfqb_sharded_connect has not been called");
  const unsigned long long rank_base = (unsigned long long)st.rank << st.plan.n_local;
  const size_t amps = size_t(1) << st.plan.n_local;
  const unsigned long long timeout = PeerTimeoutNs();
  const int nt = st.n_terms;
  // Everything that may load code does so now, before the first wait is on
  // the stream: resident kernels, and the specialised kernels of every gate
  // segment and expectation pass (compiled once the job is being re-used).
  PreloadShardedKernels();
  for (auto& gp : st.gates) {
    const CompiledPlan& cp = *gp;
    {
      std::lock_guard<std::mutex> lock(cp.jit_mu);
      cp.jit_work += double(amps);
      cp.jit_calls++;
    }
    for (size_t p = 0; p < cp.host.passes.size(); ++p)
      JitKernelFor(ctx, cp, int(p), false, 1, false);
  }
  for (auto& ep : st.exps)
    for (size_t p = 0; p < ep->host.passes.size(); ++p)
      ExpJitKernelFor(ctx, *ep, int(p), double(amps), 1);
  // a new evaluation: fresh |0..0>, fresh partial sums
  st.state_ready = false;
  TFQB_CUDA(cudaMemsetAsync(st.d_per_term, 0, std::max<size_t>(nt, 1) * sizeof(double), ctx->stream));
  const unsigned last_reduce = st.epoch;     // peers read our partials at this epoch
  st.exchanges_run = 0;
  st.fused_run = 0;
  auto event = [&](size_t k) -> cudaEvent_t {
    while (st.events.size() <= k) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      st.events.push_back(e);
    }
    return st.events[k];
  };
  for (size_t i = 0; i < st.plan.stages.size(); ++i) {
    const ShardedStage sg = st.plan.stages[i];
    if (sg.kind == 0) {
      const CompiledPlan& cp = *st.gates[sg.index];
      const bool init = !st.state_ready && !cp.host.passes.empty();
      if (!init && !st.state_ready) {
        TFQB_CUDA(cudaMemsetAsync(st.buf[st.cur], 0, amps * sizeof(float2), ctx->stream));
        if (st.rank == 0) {
          const float2 one = make_float2(1.f, 0.f);
          TFQB_CUDA(cudaMemcpyAsync(st.buf[st.cur], &one, sizeof(one), cudaMemcpyHostToDevice, ctx->stream));
          TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
      }
      TFQB_RETURN_IF(RunPlan(ctx, cp, st.buf[st.cur], nullptr, 1, job->d_params, job->n_symbols,
                             job->d_mats, init, nullptr, rank_base, false, false));
      st.state_ready = true;
    } else if (sg.kind == 1) {
      if (!st.state_ready) {
        TFQB_CUDA(cudaMemsetAsync(st.buf[st.cur], 0, amps * sizeof(float2), ctx->stream));
        if (st.rank == 0) {
          const float2 one = make_float2(1.f, 0.f);
          TFQB_CUDA(cudaMemcpyAsync(st.buf[st.cur], &one, sizeof(one), cudaMemcpyHostToDevice, ctx->stream));
          TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        st.state_ready = true;
      }
      NvtxRange nvtx_x("tfqb:qubit_exchange_peer_memory");
      const unsigned e = ++st.epoch;
      const size_t k = size_t(st.exchanges_run) * 3;
      cudaEventRecord(event(k), ctx->stream);
      // our shard is complete; the peers' shards are complete; nobody still
      // reads the buffer we are about to overwrite (it was the source of the
      // previous exchange)
      LaunchPeerSignal(reinterpret_cast<unsigned*>(st.flag_block), e, ctx->stream);
      LaunchPeerWait(st.d_peer_ready, st.world, st.rank, e, timeout, st.d_error, ctx->stream);
      LaunchPeerWait(st.d_peer_done, st.world, st.rank, e - 1, timeout, st.d_error, ctx->stream);
      cudaEventRecord(event(k + 1), ctx->stream);
      // Fused: when a gate segment with at least one pass follows, ITS first
      // pass loads the tiles straight from the peers (kernels.cuh init_mode 3)
      // and writes them, gates applied, into the alternate buffer: the
      // all-to-all costs no HBM write + read of its own and its NVLink time
      // overlaps the gate arithmetic tile by tile.
      const bool next_is_gates =
          FusedExchangeEnabled() && i + 1 < st.plan.stages.size() &&
          st.plan.stages[i + 1].kind == 0 &&
          !st.gates[st.plan.stages[i + 1].index]->host.passes.empty();
      if (next_is_gates) {
        const CompiledPlan& cp = *st.gates[st.plan.stages[i + 1].index];
        TFQB_RETURN_IF(RunPlan(ctx, cp, st.buf[st.cur ^ 1], nullptr, 1, job->d_params,
                               job->n_symbols, job->d_mats, false, nullptr, rank_base,
                               false, false, st.d_peer_buf[st.cur], st.plan.n_local - st.plan.g,
                               (unsigned long long)st.rank << (st.plan.n_local - st.plan.g),
                               /*first_pass_only=*/1));
        st.fused_run++;
      } else {
        LaunchPeerPull(st.buf[st.cur ^ 1], st.d_peer_buf[st.cur], st.world, st.rank,
                       amps >> st.plan.g, ctx->stream);
      }
      cudaEventRecord(event(k + 2), ctx->stream);
      LaunchPeerSignal(reinterpret_cast<unsigned*>(st.flag_block) + 1, e, ctx->stream);
      ctx->prof.kernel_launches += 5;
      st.cur ^= 1;
      st.exchanges_run++;
      if (next_is_gates) {
        // the rest of that segment, in place
        const CompiledPlan& cp = *st.gates[st.plan.stages[i + 1].index];
        TFQB_RETURN_IF(RunPlan(ctx, cp, st.buf[st.cur], nullptr, 1, job->d_params,
                               job->n_symbols, job->d_mats, false, nullptr, rank_base,
                               false, false, nullptr, 0, 0, /*first_pass_only=*/-1));
        ++i;            // the segment is done
      }
    } else {
      if (!st.state_ready) return Fail(TFQB_INTERNAL, "expectation stage before any gate segment");
      TFQB_RETURN_IF(RunExpectationTerms(ctx, *st.exps[sg.index], st.buf[st.cur], 1,
                                         job->groups[0].d_terms, nt, job->n_ops, st.d_per_term,
                                         rank_base, false));
    }
  }
  // per-term partial sums: publish, then every rank adds all of them in rank
  // order (identical bits everywhere, no collective)
  if (st.world > 1) {
    LaunchPeerWait(st.d_peer_done, st.world, st.rank, last_reduce, timeout, st.d_error, ctx->stream);
    LaunchPeerPublishPartials(st.d_per_term, reinterpret_cast<double*>(st.flag_block + 64), nt, ctx->stream);
    const unsigned e = ++st.epoch;
    LaunchPeerSignal(reinterpret_cast<unsigned*>(st.flag_block), e, ctx->stream);
    LaunchPeerWait(st.d_peer_ready, st.world, st.rank, e, timeout, st.d_error, ctx->stream);
    LaunchPeerReducePartials(st.d_peer_parts, st.world, nt, st.d_total, ctx->stream);
    LaunchPeerSignal(reinterpret_cast<unsigned*>(st.flag_block) + 1, e, ctx->stream);
    ctx->prof.kernel_launches += 6;
  } else if (nt > 0) {
    TFQB_CUDA(cudaMemcpyAsync(st.d_total, st.d_per_term, size_t(nt) * sizeof(double),
                              cudaMemcpyDeviceToDevice, ctx->stream));
  }
  st.enqueued = true;
  TFQB_CUDA(cudaGetLastError());
  return TFQB_OK;
}

static int impl_tfqb_sharded_result(tfqb_job* job, float* expectations) {
  if (!job || !job->sharded) return Fail(TFQB_INVALID_ARGUMENT, "not a sharded job");
  tfqb_context* ctx = job->ctx;
  TFQB_RETURN_IF(CheckContext(ctx));
  std::vector<double> tot;
  {
    std::lock_guard<std::mutex> lock(ctx->mu);
    ShardedState& st = *job->sharded;
    if (!st.enqueued) return Fail(TFQB_INVALID_ARGUMENT, "tfqb_sharded_enqueue has not been called");
    tot.assign(std::max(st.n_terms, 1), 0.0);
    int err = 0;
    if (st.n_terms)
      TFQB_CUDA(cudaMemcpyAsync(tot.data(), st.d_total, sizeof(double) * st.n_terms,
                                cudaMemcpyDeviceToHost, ctx->stream));
    TFQB_CUDA(cudaMemcpyAsync(&err, st.d_error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (err != 0)
      return Fail(TFQB_INTERNAL, "sharded exchange: timed out waiting for rank " +
                                     std::to_string(err - 1) + " (TFQB_PEER_TIMEOUT_S)");
  }
  return tfqb_sharded_finish(job, tot.data(), expectations);
}


// Sampling from a sharded state ("per-rank norms ->
// all-gather -> each shot assigned to a rank by the coarse CDF"; the
// single-circuit path it covers: tfq_simulate_samples_op.cc:122-180).  After
// tfqb_sharded_enqueue: every rank builds the probability tree of its shard,
// publishes the shard's norm through the flag block, reads all norms, takes
// the shots whose (sorted) uniforms fall into its slice of the coarse CDF,
// and samples them from its own tree with the uniform rescaled to the slice.
// `samples`: int8[num_samples, n_qubits], only this rank's rows are written,
// `owned[s]` = 1 for them: the host merges the ranks (every shot is owned by
// exactly one).  Outcomes are ordered by PHYSICAL index along the CDF (rank
// bits on top, qubits permuted by the swaps), so for given uniforms the
// bitstrings differ from the unsharded op's; their distribution is the same.
static int impl_tfqb_sharded_sample(tfqb_job* job, int num_samples, uint64_t seed,
                                    const double* uniforms, int8_t* samples, int32_t* owned) {
  if (!job || !job->sharded) return Fail(TFQB_INVALID_ARGUMENT, "not a sharded job");
  tfqb_context* ctx = job->ctx;
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  ShardedState& st = *job->sharded;
  if (!st.enqueued || !st.state_ready)
    return Fail(TFQB_INVALID_ARGUMENT, "tfqb_sharded_enqueue has not been called");
  if (num_samples < 0) return Fail(TFQB_INVALID_ARGUMENT, "num_samples must be >= 0");
  const int S = num_samples, n = st.plan.n, nl = st.plan.n_local, W = st.world;
  for (int s2 = 0; s2 < S; ++s2) owned[s2] = 0;
  if (S == 0) return TFQB_OK;
  const size_t amps = size_t(1) << nl;
  const unsigned long long timeout = PeerTimeoutNs();
  if (!st.d_tree)
    return Fail(TFQB_INVALID_ARGUMENT,
                "tfqb_sharded_sample needs a job prepared without PauliSums (n_ops = 0)");
  if (S > kShardedMaxShots)
    return Fail(TFQB_INVALID_ARGUMENT, "at most " + std::to_string(kShardedMaxShots) +
                                           " shots per tfqb_sharded_sample call");
  double* d_tree = st.d_tree;
  double* d_norms = st.d_norms;
  double* d_u = st.d_u;
  uint64_t* d_idx = st.d_idx;
  int32_t* d_rowid = st.d_rowid;
  const size_t padded = NextPow2(uint32_t(S));
  PreloadShardedKernels();
  // the shots' uniforms, sorted ascending as the unsharded op sorts them
  std::vector<double> u(S);
  if (uniforms) {
    u.assign(uniforms, uniforms + S);
    std::sort(u.begin(), u.end());
  } else {
    const int32_t row0 = int32_t(ctx->row_offset);
    TFQB_CUDA(cudaMemcpyAsync(d_rowid, &row0, sizeof(row0), cudaMemcpyHostToDevice, ctx->stream));
    LaunchFillUniforms(d_u, padded, seed, d_rowid, 0, 0, S, 1, ctx->stream);
    LaunchSortRows(d_u, padded, 1, ctx->stream);
    TFQB_CUDA(cudaMemcpyAsync(u.data(), d_u, sizeof(double) * S, cudaMemcpyDeviceToHost, ctx->stream));
  }
  LaunchBuildTree(st.buf[st.cur], amps, nl, d_tree, 1, ctx->stream);
  LaunchTreeTotal(d_tree, nl, d_norms + W, ctx->stream);
  ctx->prof.kernel_launches += 3;
  if (W > 1) {
    // the flag block's scalar area is read by the peers at the previous epoch
    LaunchPeerWait(st.d_peer_done, W, st.rank, st.epoch, timeout, st.d_error, ctx->stream);
    LaunchPeerPublishPartials(d_norms + W, reinterpret_cast<double*>(st.flag_block + 64), 1, ctx->stream);
    const unsigned e = ++st.epoch;
    LaunchPeerSignal(reinterpret_cast<unsigned*>(st.flag_block), e, ctx->stream);
    LaunchPeerWait(st.d_peer_ready, W, st.rank, e, timeout, st.d_error, ctx->stream);
    LaunchPeerGatherScalars(st.d_peer_parts, W, d_norms, ctx->stream);
    LaunchPeerSignal(reinterpret_cast<unsigned*>(st.flag_block) + 1, e, ctx->stream);
    ctx->prof.kernel_launches += 6;
  } else {
    TFQB_CUDA(cudaMemcpyAsync(d_norms, d_norms + W, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  std::vector<double> norms(W);
  int err = 0;
  TFQB_CUDA(cudaMemcpyAsync(norms.data(), d_norms, sizeof(double) * W, cudaMemcpyDeviceToHost, ctx->stream));
  TFQB_CUDA(cudaMemcpyAsync(&err, st.d_error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (err != 0)
    return Fail(TFQB_INTERNAL, "sharded exchange: timed out waiting for rank " +
                                   std::to_string(err - 1) + " (TFQB_PEER_TIMEOUT_S)");
  double total = 0.0;
  std::vector<double> cum(W + 1, 0.0);
  for (int r = 0; r < W; ++r) {
    total += norms[r];
    cum[r + 1] = total;
  }
  std::vector<double> mine;
  std::vector<int> shot_of;
  for (int s2 = 0; s2 < S; ++s2) {
    const double x = u[s2] * total;
    int r = 0;
    while (r + 1 < W && !(x < cum[r + 1])) ++r;
    if (r != st.rank) continue;
    double v = norms[r] > 0.0 ? (x - cum[r]) / norms[r] : 0.0;
    if (v < 0.0) v = 0.0;
    if (!(v < 1.0)) v = std::nextafter(1.0, 0.0);
    mine.push_back(v);
    shot_of.push_back(s2);
  }
  if (mine.empty()) return TFQB_OK;
  const int cnt = int(mine.size());
  TFQB_CUDA(cudaMemcpyAsync(d_u, mine.data(), sizeof(double) * cnt, cudaMemcpyHostToDevice, ctx->stream));
  LaunchSample(st.buf[st.cur], amps, nl, d_tree, d_u, size_t(cnt), nullptr, cnt, 1, d_idx, size_t(cnt),
               ctx->stream);
  ctx->prof.kernel_launches++;
  std::vector<uint64_t> idx(cnt);
  TFQB_CUDA(cudaMemcpyAsync(idx.data(), d_idx, sizeof(uint64_t) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
  TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < cnt; ++k) {
    const uint64_t phys = (uint64_t(st.rank) << nl) | idx[k];
    int8_t* row = samples + size_t(shot_of[k]) * n;
    for (int b = 0; b < n; ++b)     // logical index bit b sits at physical position final_phys[b]
      row[n - 1 - b] = int8_t((phys >> st.plan.final_phys[b]) & 1ull);
    owned[shot_of[k]] = 1;
  }
  return TFQB_OK;
}

static int impl_tfqb_sharded_stats(tfqb_job* job, tfqb_exchange_stats* out) {
  if (!job || !job->sharded || !out) return Fail(TFQB_INVALID_ARGUMENT, "not a sharded job");
  tfqb_context* ctx = job->ctx;
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  ShardedState& st = *job->sharded;
  TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
  memset(out, 0, sizeof *out);
  out->n_qubits = st.plan.n;
  out->n_local = st.plan.n_local;
  out->exchanges = st.exchanges_run;
  out->fused_exchanges = st.fused_run;
  out->shard_bytes = double((size_t(1) << st.plan.n_local) * sizeof(float2));
  // what one exchange moves over NVLink per GPU: the shard minus its own chunk
  out->bytes_received_per_exchange = out->shard_bytes * (1.0 - 1.0 / double(st.world));
  for (int k = 0; k < st.exchanges_run; ++k) {
    float w = 0.f, p = 0.f;
    cudaEventElapsedTime(&w, st.events[3 * k], st.events[3 * k + 1]);
    cudaEventElapsedTime(&p, st.events[3 * k + 1], st.events[3 * k + 2]);
    out->wait_ms += w;
    out->pull_ms += p;
  }
  for (auto& gp : st.gates) out->gate_passes += int(gp->host.passes.size());
  for (auto& ep : st.exps) out->expectation_passes += int(ep->host.passes.size());
  return TFQB_OK;
}

// ---- instrumentation --------------------------------------------------------
static int impl_tfqb_sync(tfqb_context* ctx) {
  TFQB_RETURN_IF(CheckContext(ctx));
  TFQB_CUDA(cudaStreamSynchronize(ctx->stream));
  return TFQB_OK;
}

void* tfqb_stream(tfqb_context* ctx) {
  if (IsMultiCtx(ctx)) ctx = ctx->children[0];
  return ctx ? (void*)ctx->stream : nullptr;
}

static int impl_tfqb_profile_enable(tfqb_context* ctx, int enable) {
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->prof_timing = enable != 0;
  return TFQB_OK;
}

static void DrainTimed(tfqb_context* ctx, bool accumulate) {
  cudaStreamSynchronize(ctx->stream);
  for (auto& t : ctx->timed) {
    if (accumulate) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
        if (t.kind == 0) ctx->prof.gate_pass_ms += ms;
        else if (t.kind == 1) ctx->prof.adjoint_pass_ms += ms;
        else ctx->prof.expectation_ms += ms;
      }
    }
    ctx->event_pool.emplace_back(t.a, t.b);
  }
  ctx->timed.clear();
}

static int impl_tfqb_profile_reset(tfqb_context* ctx) {
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  DrainTimed(ctx, false);
  ctx->prof = tfqb_profile{};
  return TFQB_OK;
}

static int impl_tfqb_profile_read(tfqb_context* ctx, tfqb_profile* out) {
  TFQB_RETURN_IF(CheckContext(ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  DrainTimed(ctx, true);
  *out = ctx->prof;
  return TFQB_OK;
}

// ---- host-only helpers ------------------------------------------------------
static int impl_tfqb_host_gate_matrix(int kind, const float* params, int n_params,
                          int grad_param, float* out) {
  if (kind < 0 || kind >= kNumGateKinds)
    return Fail(TFQB_INVALID_ARGUMENT, "unknown gate kind");
  float p[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < n_params && i < 5; ++i) p[i] = params[i];
  const bool two = kind == kI2 || (kind >= kXXP && kind <= kISP) ||
                   kind == kFSIM || kind == kPISP;
  const int dim = two ? 4 : 2;
  cf m[16];
  if (grad_param >= 0) gradient_matrix(kind, p, grad_param, dim, m);
  else gate_matrix(kind, p, -1, 0.f, m);
  for (int i = 0; i < dim * dim; ++i) {
    out[2 * i] = m[i].re;
    out[2 * i + 1] = m[i].im;
  }
  return TFQB_OK;
}

static char* DupString(const std::string& s) {
  char* r = static_cast<char*>(malloc(s.size() + 1));
  memcpy(r, s.c_str(), s.size() + 1);
  return r;
}

static int impl_tfqb_host_describe_plan(const char* program, size_t program_size,
                            tfqb_strings symbol_names, int n_symbols,
                            int adjoint, char** json_out) {
  ProgramPB pb;
  if (!ParseProgram(program, program_size, &pb))
    return Fail(TFQB_INVALID_ARGUMENT, "Unparseable proto: " + std::string(program, std::min<size_t>(program_size, 64)));
  SymbolTable symbols = MakeSymbolTable(symbol_names.data, symbol_names.size, n_symbols);
  CircuitT c;
  Status s = LowerProgram(pb, symbols, &c);
  if (!s.ok) return Fail(TFQB_INVALID_ARGUMENT, s.msg);
  std::ostringstream o;
  o << "{\"n\":" << c.n << ",\"gates\":[";
  for (size_t i = 0; i < c.gates.size(); ++i) {
    const GateT& g = c.gates[i];
    if (i) o << ",";
    o << "{\"kind\":" << g.kind << ",\"bits\":[" << g.bit[0];
    if (g.nq == 2) o << "," << g.bit[1];
    o << "],\"cmask\":" << g.cmask << ",\"cbits\":" << g.cbits << ",\"syms\":[";
    for (int k = 0; k < g.nsym; ++k) o << (k ? "," : "") << g.sym_col[k];
    o << "]}";
  }
  o << "]";
  if (c.n > 0) {
    DevicePlan p = adjoint ? PlanAdjoint(c, kTileMax, GateLowBits(), AdjRegBits())
                           : PlanForward(c, kTileMax, GateLowBits(), true);
    o << ",\"n_alloc\":" << p.n_alloc << ",\"passes\":[";
    for (size_t i = 0; i < p.passes.size(); ++i) {
      const PassRec& pr = p.passes[i];
      if (i) o << ",";
      o << "{\"tile\":[";
      for (int k = 0; k < pr.tile_bits; ++k) o << (k ? "," : "") << pr.tile_pos[k];
      const int nops = pr.round_end > pr.round_begin
                           ? p.rounds[pr.round_end - 1].op_end -
                                 p.rounds[pr.round_begin].op_begin
                           : 0;
      o << "],\"rounds\":" << (pr.round_end - pr.round_begin) << ",\"ops\":"
        << nops;
      if (!adjoint)
        o << ",\"packed_fp32_per_amp\":" << PassPackedFp32PerAmplitude(p, int(i), false)
          << ",\"packed_fp32_per_amp_phase_free\":" << PassPackedFp32PerAmplitude(p, int(i), true);
      o << ",\"round_ops\":[";
      for (int r = pr.round_begin; r < pr.round_end; ++r) {
        const RoundRec& rr = p.rounds[r];
        o << (r > pr.round_begin ? "," : "") << "{\"pos\":[" << rr.pos[0] << ","
          << rr.pos[1] << "," << rr.pos[2] << "," << rr.pos[3] << "],\"codes\":[";
        for (int k = rr.op_begin; k < rr.op_end; ++k)
          o << (k > rr.op_begin ? "," : "") << p.ops[k].code;
        o << "]}";
      }
      o << "]}";
    }
    o << "],\"n_ops\":" << p.ops.size() << ",\"n_factors\":" << p.factors.size()
      << ",\"mat_floats\":" << p.mat_floats
      << ",\"grad_slots\":[";
    for (size_t i = 0; i < p.grad_slots.size(); ++i)
      o << (i ? "," : "") << p.grad_slots[i].symbol_col;
    int init_identity = 0;
    if (p.product_init)
      for (const MatRec& mr : p.mats)
        if (mr.layout == 4 && mr.factor_end - mr.factor_begin == 1 &&
            p.factors[mr.factor_begin].gate_kind == kI)
          ++init_identity;
    o << "],\"row_dependent\":" << (p.row_dependent ? "true" : "false")
      << ",\"product_init\":" << (p.product_init ? "true" : "false")
      << ",\"init_identity_bits\":" << init_identity
      << ",\"macro_merged\":" << p.macro_merged;
  }
  o << "}";
  *json_out = DupString(o.str());
  return TFQB_OK;
}

static int impl_tfqb_host_jit_source(const char* program, size_t program_size,
                         tfqb_strings symbol_names, int n_symbols,
                         int adjoint, int pass, char** source_out) {
  ProgramPB pb;
  if (!ParseProgram(program, program_size, &pb))
    return Fail(TFQB_INVALID_ARGUMENT, "Unparseable proto");
  SymbolTable symbols = MakeSymbolTable(symbol_names.data, symbol_names.size, n_symbols);
  CircuitT c;
  Status s = LowerProgram(pb, symbols, &c);
  if (!s.ok) return Fail(TFQB_INVALID_ARGUMENT, s.msg);
  std::string src;
  if (c.n > 0) {
    DevicePlan p = adjoint == 1 ? PlanAdjoint(c, kTileMax, GateLowBits(), AdjRegBits())
                                : PlanForward(c, kTileMax, GateLowBits(), true);
    if (pass >= 0 && pass < int(p.passes.size()) && PassIsJitable(p, pass, adjoint == 1))
      src = GeneratePassSource(p, pass, adjoint == 1, adjoint != 0);
  }
  *source_out = DupString(src);
  return TFQB_OK;
}

static int impl_tfqb_host_describe_pauli_sum(const char* program, size_t program_size,
                                 const char* pauli_sum, size_t pauli_sum_size,
                                 char** json_out) {
  ProgramPB pb;
  if (!ParseProgram(program, program_size, &pb))
    return Fail(TFQB_INVALID_ARGUMENT, "Unparseable proto: " + std::string(program, std::min<size_t>(program_size, 64)));
  SymbolTable symbols;
  CircuitT c;
  // symbols are irrelevant for the qubit map: lower with an "accept all" table
  for (auto& m : pb.moments)
    for (auto& op : m.operations)
      for (auto& a : op.args)
        if (!a.symbol.empty() && !symbols.col.count(a.symbol)) {
          const int next = int(symbols.col.size());
          symbols.col[a.symbol] = next;
        }
  symbols.size = int(symbols.col.size());
  Status s = LowerProgram(pb, symbols, &c);
  if (!s.ok) return Fail(TFQB_INVALID_ARGUMENT, s.msg);
  PauliSumPB ps;
  if (!ParsePauliSum(pauli_sum, pauli_sum_size, &ps))
    return Fail(TFQB_INVALID_ARGUMENT, "Unparseable proto: " + std::string(pauli_sum, std::min<size_t>(pauli_sum_size, 64)));
  PauliSumT t;
  s = LowerPauliSum(ps, c, &t);
  if (!s.ok) return Fail(TFQB_INVALID_ARGUMENT, s.msg);
  std::ostringstream o;
  o << "{\"n\":" << c.n << ",\"terms\":[";
  for (size_t i = 0; i < t.terms.size(); ++i) {
    const auto& tt = t.terms[i];
    if (i) o << ",";
    char buf[64];
    snprintf(buf, sizeof buf, "%.9g", double(tt.coeff));
    o << "{\"coeff\":" << buf << ",\"x\":" << tt.x << ",\"z\":" << tt.z
      << ",\"phase\":" << tt.phase << ",\"identity\":" << (tt.identity ? 1 : 0)
      << ",\"parity_mask\":" << tt.parity_mask << "}";
  }
  o << "]}";
  *json_out = DupString(o.str());
  return TFQB_OK;
}

static int impl_tfqb_host_jit_expect_source(const char* program, size_t program_size,
                                tfqb_strings pauli_sums, int n_ops, int pass,
                                char** source_out) {
  ProgramPB pb;
  if (!ParseProgram(program, program_size, &pb))
    return Fail(TFQB_INVALID_ARGUMENT, "Unparseable proto");
  SymbolTable symbols;
  for (auto& m : pb.moments)
    for (auto& op : m.operations)
      for (auto& a : op.args)
        if (!a.symbol.empty() && !symbols.col.count(a.symbol)) {
          const int next = int(symbols.col.size());
          symbols.col[a.symbol] = next;
        }
  symbols.size = int(symbols.col.size());
  CircuitT c;
  Status s = LowerProgram(pb, symbols, &c);
  if (!s.ok) return Fail(TFQB_INVALID_ARGUMENT, s.msg);
  std::vector<TermMask> tm;
  for (int j = 0; j < n_ops; ++j) {
    PauliSumPB ps;
    if (!ParsePauliSum(pauli_sums.data[j], pauli_sums.size[j], &ps))
      return Fail(TFQB_INVALID_ARGUMENT, "Unparseable proto: pauli sum");
    PauliSumT t;
    s = LowerPauliSum(ps, c, &t);
    if (!s.ok) return Fail(TFQB_INVALID_ARGUMENT, s.msg);
    for (const auto& tt : t.terms) tm.push_back(TermMask{tt.x, tt.z, tt.phase, tt.identity});
  }
  std::string src;
  if (c.n > 0 && !tm.empty()) {
    const bool accum = pass >= 1000;     // the K3 kernel of pass - 1000
    if (accum) pass -= 1000;
    ExpectationPlan ep = accum ? PlanExpectation(c.n, tm, true)
                               : PlanExpectation(c.n, tm, false, kTileMax, ExpLowBits());
    if (pass >= 0 && pass < int(ep.passes.size()) && ExpectPassIsJitable(ep, pass))
      src = accum ? GenerateAccumSource(ep, pass) : GenerateExpectSource(ep, pass);
  }
  *source_out = DupString(src);
  return TFQB_OK;
}

static int impl_tfqb_host_describe_sharded(const char* program, size_t program_size,
                               tfqb_strings symbol_names, int n_symbols,
                               tfqb_strings pauli_sums, int n_ops, int world,
                               char** json_out) {
  ProgramPB pb;
  if (!ParseProgram(program, program_size, &pb))
    return Fail(TFQB_INVALID_ARGUMENT, "Unparseable proto: " + std::string(program, std::min<size_t>(program_size, 64)));
  SymbolTable symbols = MakeSymbolTable(symbol_names.data, symbol_names.size, n_symbols);
  CircuitT c;
  Status s = LowerProgram(pb, symbols, &c);
  if (!s.ok) return Fail(TFQB_INVALID_ARGUMENT, s.msg);
  int g = 0;
  while ((1 << g) < world) ++g;
  if ((1 << g) != world || c.n - g < std::max(kMinStateBits, 2 * g + 2))
    return Fail(TFQB_INVALID_ARGUMENT, "bad world size for this circuit");
  std::vector<TermMask> tm;
  for (int j = 0; j < n_ops; ++j) {
    PauliSumPB ps;
    if (!ParsePauliSum(pauli_sums.data[j], pauli_sums.size[j], &ps))
      return Fail(TFQB_INVALID_ARGUMENT, "Unparseable proto: pauli sum");
    PauliSumT t;
    s = LowerPauliSum(ps, c, &t);
    if (!s.ok) return Fail(TFQB_INVALID_ARGUMENT, s.msg);
    for (const auto& tt : t.terms) tm.push_back(TermMask{tt.x, tt.z, tt.phase, tt.identity});
  }
  ShardedPlan sp = PlanSharded(c, g, tm);
  std::ostringstream o;
  o << "{\"n\":" << sp.n << ",\"n_local\":" << sp.n_local << ",\"g\":" << sp.g
    << ",\"n_exchanges\":" << sp.n_exchanges << ",\"stages\":[";
  for (size_t i = 0; i < sp.stages.size(); ++i) {
    const auto& st = sp.stages[i];
    if (i) o << ",";
    o << "{\"kind\":" << st.kind;
    if (st.kind == 0) {
      const DevicePlan& p = sp.gate_plans[st.index];
      o << ",\"passes\":" << p.passes.size() << ",\"ops\":" << p.ops.size()
        << ",\"factors\":" << p.factors.size() << ",\"after_exchange\":" << (p.after_exchange ? 1 : 0)
        << ",\"tiles_per_cta\":[";     // of the specialised kernels (jit.h JitPassTiles)
      for (size_t q = 0; q < p.passes.size(); ++q)
        o << (q ? "," : "") << JitPassTiles(p, false, int(q));
      o << "]";
    } else if (st.kind == 2) {
      const ExpectationPlan& e = sp.exp_plans[st.index];
      o << ",\"passes\":" << e.passes.size() << ",\"zterms\":" << e.zterms.size()
        << ",\"xops\":" << e.xops.size() << ",\"deferred\":" << e.deferred_terms.size();
    }
    o << "}";
  }
  o << "],\"final_phys\":[";
  for (size_t i = 0; i < sp.final_phys.size(); ++i) o << (i ? "," : "") << sp.final_phys[i];
  o << "]}";
  *json_out = DupString(o.str());
  return TFQB_OK;
}

void tfqb_free_string(char* s) { free(s); }

}  // extern "C"

// ===========================================================================
// One context over several GPUs (tfqb_create_multi).
//
// The reference spreads ONE OpKernel::Compute over every host core, rows
// first (tfq_simulate_expectation_op.cc:245-248, tfq_adj_grad_op.cc:282-283:
// context->device()->tensorflow_cpu_worker_threads()->workers->ParallelFor).
// Here one call spreads its rows over every GPU of the context: contiguous row
// blocks, one host thread per device, each block running the single-device
// path on its own child context and writing its slice of the caller's output
// tensor.  Rows are independent: there is no collective.
// ===========================================================================
namespace {

bool IsMulti(const tfqb_context* ctx) { return ctx && !ctx->children.empty(); }

struct RowBlock { int lo, hi; };

// contiguous, balanced; never more blocks than rows (an empty batch is one
// empty block so that the child validates the call and shapes the output)
std::vector<RowBlock> SplitRows(int batch, int n_dev) {
  std::vector<RowBlock> out;
  if (batch <= 0) {
    out.push_back(RowBlock{0, 0});
    return out;
  }
  const int n = std::min(batch, n_dev);
  for (int k = 0; k < n; ++k)
    out.push_back(RowBlock{int((long long)batch * k / n), int((long long)batch * (k + 1) / n)});
  return out;
}

tfqb_strings Shift(tfqb_strings s, size_t off) {
  tfqb_strings r = s;
  if (r.data) r.data += off;
  if (r.size) r.size += off;
  return r;
}

tfqb_circuit_inputs SubInputs(const tfqb_circuit_inputs* in, RowBlock b) {
  tfqb_circuit_inputs r = *in;
  r.programs = Shift(in->programs, size_t(b.lo));
  r.batch = b.hi - b.lo;
  if (in->symbol_values) r.symbol_values = in->symbol_values + size_t(b.lo) * in->n_symbols;
  r.symbol_rows = b.hi - b.lo;
  return r;
}

// fn(child, block index, block) on one host thread per block; the first
// failing block (lowest rows) decides the status and the message.
template <typename Fn>
int FanOut(tfqb_context* ctx, const std::vector<RowBlock>& blocks, Fn fn) {
  const size_t nb = blocks.size();
  std::vector<int> rc(nb, TFQB_OK);
  std::vector<std::string> msg(nb);
  auto run = [&](size_t k) {
    tfqb_context* child = ctx->children[k];
    child->row_offset = ctx->row_offset + blocks[k].lo;
    try {
      rc[k] = fn(child, int(k), blocks[k]);
    } catch (const std::bad_alloc&) {
      rc[k] = Fail(TFQB_RESOURCE_EXHAUSTED, "Out of host memory.");
    } catch (const std::exception& e) {
      rc[k] = Fail(TFQB_INTERNAL, std::string("internal error: ") + e.what());
    }
    if (rc[k] != TFQB_OK) msg[k] = g_last_error;   // thread-local of the worker
  };
  if (nb == 1) {
    run(0);
  } else {
    std::vector<std::thread> th;
    for (size_t k = 0; k < nb; ++k) th.emplace_back(run, k);
    for (auto& t : th) t.join();
  }
  for (size_t k = 0; k < nb; ++k)
    if (rc[k] != TFQB_OK) return Fail(rc[k], msg[k]);
  return TFQB_OK;
}

// Row counts that do not match are an error of the whole call: let one child
// report it with the reference's message.
bool RowsConsistent(const tfqb_circuit_inputs* in, int other_rows) {
  return in->batch >= 0 && in->symbol_rows == in->batch && other_rows == in->batch;
}

int MultiExpectation(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                     tfqb_strings pauli_sums, int sum_rows, int n_ops, float* out) {
  if (!RowsConsistent(in, sum_rows))
    return tfqb_simulate_expectation(ctx->children[0], in, pauli_sums, sum_rows, n_ops, out);
  return FanOut(ctx, SplitRows(in->batch, int(ctx->children.size())),
                [&](tfqb_context* c, int, RowBlock b) {
                  const tfqb_circuit_inputs sub = SubInputs(in, b);
                  return tfqb_simulate_expectation(
                      c, &sub, Shift(pauli_sums, size_t(b.lo) * n_ops), b.hi - b.lo, n_ops,
                      out + size_t(b.lo) * n_ops);
                });
}

int MultiAdjoint(tfqb_context* ctx, const tfqb_circuit_inputs* in, tfqb_strings pauli_sums,
                 int sum_rows, int n_ops, const float* down, int grad_rows, int grad_cols,
                 float* grads) {
  if (!RowsConsistent(in, sum_rows) || grad_rows != in->batch)
    return tfqb_adjoint_gradient(ctx->children[0], in, pauli_sums, sum_rows, n_ops, down,
                                 grad_rows, grad_cols, grads);
  return FanOut(ctx, SplitRows(in->batch, int(ctx->children.size())),
                [&](tfqb_context* c, int, RowBlock b) {
                  const tfqb_circuit_inputs sub = SubInputs(in, b);
                  return tfqb_adjoint_gradient(
                      c, &sub, Shift(pauli_sums, size_t(b.lo) * n_ops), b.hi - b.lo, n_ops,
                      down + size_t(b.lo) * grad_cols, b.hi - b.lo, grad_cols,
                      grads + size_t(b.lo) * in->n_symbols);
                });
}

int MultiSampledExpectation(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                            tfqb_strings pauli_sums, int sum_rows, int n_ops,
                            const int32_t* num_samples, int ns_rows, int ns_cols,
                            uint64_t seed, const double* uniforms, int uniform_terms,
                            int uniform_shots, float* out) {
  if (!RowsConsistent(in, sum_rows) || ns_rows != in->batch || ns_cols != n_ops)
    return tfqb_simulate_sampled_expectation(ctx->children[0], in, pauli_sums, sum_rows,
                                             n_ops, num_samples, ns_rows, ns_cols, seed,
                                             uniforms, uniform_terms, uniform_shots, out);
  return FanOut(ctx, SplitRows(in->batch, int(ctx->children.size())),
                [&](tfqb_context* c, int, RowBlock b) {
                  const tfqb_circuit_inputs sub = SubInputs(in, b);
                  const size_t urow = size_t(n_ops) * uniform_terms * uniform_shots;
                  return tfqb_simulate_sampled_expectation(
                      c, &sub, Shift(pauli_sums, size_t(b.lo) * n_ops), b.hi - b.lo, n_ops,
                      num_samples + size_t(b.lo) * ns_cols, b.hi - b.lo, ns_cols, seed,
                      uniforms ? uniforms + size_t(b.lo) * urow : nullptr, uniform_terms,
                      uniform_shots, out + size_t(b.lo) * n_ops);
                });
}

int MultiNoisyExpectation(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                          tfqb_strings pauli_sums, int sum_rows, int n_ops,
                          const int32_t* num_samples, int ns_rows, int ns_cols, uint64_t seed,
                          const float* uniforms, int uniform_traj, int uniform_chan, bool sampled,
                          float* out) {
  auto call = [&](tfqb_context* c, const tfqb_circuit_inputs* sub, tfqb_strings sums, int rows,
                  const int32_t* ns, int nsr, const float* u, float* o) {
    return sampled ? tfqb_noisy_sampled_expectation(c, sub, sums, rows, n_ops, ns, nsr, ns_cols, seed,
                                                    u, uniform_traj, uniform_chan, o)
                   : tfqb_noisy_expectation(c, sub, sums, rows, n_ops, ns, nsr, ns_cols, seed, u,
                                            uniform_traj, uniform_chan, o);
  };
  if (!RowsConsistent(in, sum_rows) || ns_rows != in->batch || ns_cols != n_ops)
    return call(ctx->children[0], in, pauli_sums, sum_rows, num_samples, ns_rows, uniforms, out);
  return FanOut(ctx, SplitRows(in->batch, int(ctx->children.size())),
                [&](tfqb_context* c, int, RowBlock b) {
                  const tfqb_circuit_inputs sub = SubInputs(in, b);
                  const size_t urow = size_t(uniform_traj) * uniform_chan;
                  return call(c, &sub, Shift(pauli_sums, size_t(b.lo) * n_ops), b.hi - b.lo,
                              num_samples + size_t(b.lo) * ns_cols, b.hi - b.lo,
                              uniforms ? uniforms + size_t(b.lo) * urow : nullptr,
                              out + size_t(b.lo) * n_ops);
                });
}

int MultiInnerProduct(tfqb_context* ctx, const tfqb_circuit_inputs* in, tfqb_strings others,
                      int other_rows, int n_other, float* out) {
  if (!RowsConsistent(in, other_rows))
    return tfqb_inner_product(ctx->children[0], in, others, other_rows, n_other, out);
  return FanOut(ctx, SplitRows(in->batch, int(ctx->children.size())),
                [&](tfqb_context* c, int, RowBlock b) {
                  const tfqb_circuit_inputs sub = SubInputs(in, b);
                  return tfqb_inner_product(c, &sub, Shift(others, size_t(b.lo) * n_other),
                                            b.hi - b.lo, n_other,
                                            out + size_t(b.lo) * n_other * 2);
                });
}

int MultiInnerProductGrad(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                          tfqb_strings others, int other_rows, int n_other, const float* down,
                          int grad_rows, int grad_cols, float* grads) {
  if (!RowsConsistent(in, other_rows) || grad_rows != in->batch || in->n_symbols <= 0)
    return tfqb_inner_product_grad(ctx->children[0], in, others, other_rows, n_other, down,
                                   grad_rows, grad_cols, grads);
  return FanOut(ctx, SplitRows(in->batch, int(ctx->children.size())),
                [&](tfqb_context* c, int, RowBlock b) {
                  const tfqb_circuit_inputs sub = SubInputs(in, b);
                  return tfqb_inner_product_grad(
                      c, &sub, Shift(others, size_t(b.lo) * n_other), b.hi - b.lo, n_other,
                      down + size_t(b.lo) * grad_cols, b.hi - b.lo, grad_cols,
                      grads + size_t(b.lo) * in->n_symbols * 2);
                });
}

// Jobs of a multi context: one sub-job per block.  `prepare(child, sub inputs,
// block, &job)` builds them; all sub-jobs share the widest max_qubits.
template <typename Prep>
int MultiPrepare(tfqb_context* ctx, const tfqb_circuit_inputs* in, JobKind kind, Prep prepare,
                 tfqb_job** job, int* max_qubits) {
  const std::vector<RowBlock> blocks = SplitRows(in->batch, int(ctx->children.size()));
  auto mj = std::make_unique<tfqb_job>();
  mj->ctx = ctx;
  mj->kind = kind;
  mj->batch = std::max(in->batch, 0);
  mj->n_symbols = in->n_symbols;
  mj->sub.resize(blocks.size());
  for (size_t k = 0; k < blocks.size(); ++k) mj->sub[k] = tfqb_job::Sub{nullptr, blocks[k].lo, blocks[k].hi};
  auto prep_all = [&](int force) {
    return FanOut(ctx, blocks, [&](tfqb_context* c, int k, RowBlock b) {
      if (force > 0 && mj->sub[k].job && mj->sub[k].job->nmax == force) return int(TFQB_OK);
      if (mj->sub[k].job) {
        tfqb_job_free(mj->sub[k].job);
        mj->sub[k].job = nullptr;
      }
      const tfqb_circuit_inputs sub = SubInputs(in, b);
      c->force_nmax = force;
      const int rc = prepare(c, &sub, b, &mj->sub[k].job);
      c->force_nmax = 0;
      return rc;
    });
  };
  int rc = prep_all(0);
  if (rc == TFQB_OK) {
    int nmax = 0;
    bool uneven = false;
    for (auto& sj : mj->sub) nmax = std::max(nmax, sj.job->nmax);
    for (auto& sj : mj->sub) uneven = uneven || sj.job->nmax != nmax;
    if (uneven) rc = prep_all(nmax);
    mj->nmax = nmax;
  }
  if (rc != TFQB_OK) {
    const std::string keep = g_last_error;
    for (auto& sj : mj->sub)
      if (sj.job) tfqb_job_free(sj.job);
    mj->sub.clear();
    return Fail(rc, keep);
  }
  if (max_qubits) *max_qubits = mj->nmax;
  *job = mj.release();
  return TFQB_OK;
}

template <typename Fn>
int ForEachSub(tfqb_job* job, Fn fn) {
  std::vector<RowBlock> blocks;
  for (auto& sj : job->sub) blocks.push_back(RowBlock{sj.lo, sj.hi});
  return FanOut(job->ctx, blocks,
                [&](tfqb_context*, int k, RowBlock b) { return fn(job->sub[k].job, b); });
}

}  // namespace

extern "C" {

int tfqb_create_multi(const int* device_ids, int n_devices, tfqb_context** out) {
  return GuardAbi([&]() -> int {
    if (!out) return Fail(TFQB_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    if (!device_ids || n_devices < 1)
      return Fail(TFQB_INVALID_ARGUMENT, "tfqb_create_multi needs at least one device");
    auto ctx = std::make_unique<tfqb_context>();
    // the same ordinal twice only for tests on a one-GPU box (every child
    // sizes its chunks as if the device were its own)
    const char* dup = getenv("TFQB_MULTI_ALLOW_DUPLICATES");
    const bool allow_dup = dup && *dup == '1';
    for (int k = 0; k < n_devices; ++k) {
      for (int j = 0; j < k && !allow_dup; ++j)
        if (device_ids[j] == device_ids[k]) {
          for (tfqb_context* c : ctx->children) tfqb_destroy(c);
          return Fail(TFQB_INVALID_ARGUMENT, "tfqb_create_multi: duplicate device ordinal");
        }
      tfqb_context* child = nullptr;
      const int rc = tfqb_create(device_ids[k], &child);
      if (rc != TFQB_OK) {
        for (tfqb_context* c : ctx->children) tfqb_destroy(c);
        return rc;
      }
      ctx->children.push_back(child);
    }
    ctx->device = device_ids[0];
    *out = ctx.release();
    return TFQB_OK;
  });
}

int tfqb_trim(tfqb_context* ctx) {
  return GuardAbi([&]() -> int {
    if (!ctx) return Fail(TFQB_UNAVAILABLE, "No CUDA context: the B200 backend has no CPU fallback.");
    if (IsMulti(ctx)) {
      for (tfqb_context* c : ctx->children) TFQB_RETURN_IF(tfqb_trim(c));
      return TFQB_OK;
    }
    TFQB_RETURN_IF(CheckContext(ctx));
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->Trim();
    return TFQB_OK;
  });
}

double tfqb_jit_compile_seconds(void) { return JitCompileSeconds(); }
int tfqb_jit_pending(void) { return JitPending(); }

int tfqb_device_count(tfqb_context* ctx) {
  if (!ctx) return 0;
  return ctx->children.empty() ? 1 : int(ctx->children.size());
}

int tfqb_set_row_offset(tfqb_context* ctx, int64_t first_row) {
  if (!ctx) return Fail(TFQB_UNAVAILABLE, "No CUDA context: the B200 backend has no CPU fallback.");
  if (first_row < 0) return Fail(TFQB_INVALID_ARGUMENT, "row offset must be >= 0");
  ctx->row_offset = first_row;
  return TFQB_OK;
}

// ---- exception barrier: nothing may unwind across the C ABI ----------------
int tfqb_create(int device, tfqb_context** out) {
  return GuardAbi([&]() -> int { return impl_tfqb_create(device, out); });
}

int tfqb_set_memory_budget(tfqb_context* ctx, size_t bytes) {
  return GuardAbi([&]() -> int { 
    if (IsMulti(ctx)) {
      for (tfqb_context* c : ctx->children) TFQB_RETURN_IF(impl_tfqb_set_memory_budget(c, bytes));
      return TFQB_OK;
    }
    return impl_tfqb_set_memory_budget(ctx, bytes); });
}

int tfqb_expectation_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                             tfqb_strings pauli_sums, int sum_rows, int n_ops,
                             tfqb_job** job) {
  return GuardAbi([&]() -> int { 
    if (IsMulti(ctx) && RowsConsistent(in, sum_rows))
      return MultiPrepare(ctx, in, kJobExpectation,
                          [&](tfqb_context* c, const tfqb_circuit_inputs* sub, RowBlock b, tfqb_job** j) {
                            return impl_tfqb_expectation_prepare(
                                c, sub, Shift(pauli_sums, size_t(b.lo) * n_ops), b.hi - b.lo, n_ops, j);
                          }, job, nullptr);
    if (IsMulti(ctx)) ctx = ctx->children[0];
    return impl_tfqb_expectation_prepare(ctx, in, pauli_sums, sum_rows, n_ops, job); });
}

int tfqb_adjoint_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                         tfqb_strings pauli_sums, int sum_rows, int n_ops,
                         const float* downstream_grads, int grad_rows,
                         int grad_cols, tfqb_job** job) {
  return GuardAbi([&]() -> int { 
    if (IsMulti(ctx) && RowsConsistent(in, sum_rows) && grad_rows == in->batch)
      return MultiPrepare(ctx, in, kJobAdjoint,
                          [&](tfqb_context* c, const tfqb_circuit_inputs* sub, RowBlock b, tfqb_job** j) {
                            return impl_tfqb_adjoint_prepare(
                                c, sub, Shift(pauli_sums, size_t(b.lo) * n_ops), b.hi - b.lo, n_ops,
                                downstream_grads + size_t(b.lo) * grad_cols, b.hi - b.lo, grad_cols, j);
                          }, job, nullptr);
    if (IsMulti(ctx)) ctx = ctx->children[0];
    return impl_tfqb_adjoint_prepare(ctx, in, pauli_sums, sum_rows, n_ops, downstream_grads, grad_rows, grad_cols, job); });
}

int tfqb_job_run_device(tfqb_job* job) {
  NvtxRange nvtx("tfqb_job_run_device");
  return GuardAbi([&]() -> int { 
    if (job && !job->sub.empty())
      return ForEachSub(job, [&](tfqb_job* sj, RowBlock) { return impl_tfqb_job_run_device(sj); });
    return impl_tfqb_job_run_device(job); });
}

int tfqb_job_fetch(tfqb_job* job, float* out) {
  return GuardAbi([&]() -> int { 
    if (job && !job->sub.empty())
      return ForEachSub(job, [&](tfqb_job* sj, RowBlock b) {
        return impl_tfqb_job_fetch(sj, out + size_t(b.lo) * sj->out_cols);
      });
    return impl_tfqb_job_fetch(job, out); });
}

int tfqb_simulate_expectation(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                              tfqb_strings pauli_sums, int sum_rows, int n_ops,
                              float* expectations) {
  NvtxRange nvtx("tfqb_simulate_expectation");
  return GuardAbi([&]() -> int { 
    if (IsMulti(ctx)) return MultiExpectation(ctx, in, pauli_sums, sum_rows, n_ops, expectations);
    return impl_tfqb_simulate_expectation(ctx, in, pauli_sums, sum_rows, n_ops, expectations); });
}

int tfqb_adjoint_gradient(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                          tfqb_strings pauli_sums, int sum_rows, int n_ops,
                          const float* downstream_grads, int grad_rows,
                          int grad_cols, float* grads) {
  NvtxRange nvtx("tfqb_adjoint_gradient");
  return GuardAbi([&]() -> int { 
    if (IsMulti(ctx))
      return MultiAdjoint(ctx, in, pauli_sums, sum_rows, n_ops, downstream_grads, grad_rows, grad_cols, grads);
    return impl_tfqb_adjoint_gradient(ctx, in, pauli_sums, sum_rows, n_ops, downstream_grads, grad_rows, grad_cols, grads); });
}

int tfqb_simulate_state_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                                tfqb_job** job, int* max_qubits) {
  return GuardAbi([&]() -> int { 
    if (IsMulti(ctx) && in->batch >= 0 && in->symbol_rows == in->batch)
      return MultiPrepare(ctx, in, kJobState,
                          [&](tfqb_context* c, const tfqb_circuit_inputs* sub, RowBlock, tfqb_job** j) {
                            return impl_tfqb_simulate_state_prepare(c, sub, j, nullptr);
                          }, job, max_qubits);
    if (IsMulti(ctx)) ctx = ctx->children[0];
    return impl_tfqb_simulate_state_prepare(ctx, in, job, max_qubits); });
}

int tfqb_simulate_state_run(tfqb_job* job, float* state_vector) {
  NvtxRange nvtx("tfqb_simulate_state_run");
  return GuardAbi([&]() -> int { 
    if (job && !job->sub.empty()) {
      const size_t row = size_t(2) << job->nmax;    // floats per output row
      return ForEachSub(job, [&](tfqb_job* sj, RowBlock b) {
        return impl_tfqb_simulate_state_run(sj, state_vector + size_t(b.lo) * row);
      });
    }
    return impl_tfqb_simulate_state_run(job, state_vector); });
}

int tfqb_simulate_samples_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                                  int num_samples, tfqb_job** job, int* max_qubits) {
  return GuardAbi([&]() -> int { 
    if (IsMulti(ctx) && in->batch >= 0 && in->symbol_rows == in->batch && num_samples >= 0) {
      const int rc = MultiPrepare(ctx, in, kJobSamples,
                          [&](tfqb_context* c, const tfqb_circuit_inputs* sub, RowBlock, tfqb_job** j) {
                            return impl_tfqb_simulate_samples_prepare(c, sub, num_samples, j, nullptr);
                          }, job, max_qubits);
      if (rc == TFQB_OK) (*job)->num_samples = num_samples;
      return rc;
    }
    if (IsMulti(ctx)) ctx = ctx->children[0];
    return impl_tfqb_simulate_samples_prepare(ctx, in, num_samples, job, max_qubits); });
}

int tfqb_simulate_samples_run(tfqb_job* job, uint64_t seed, const double* uniforms,
                              int8_t* samples) {
  NvtxRange nvtx("tfqb_simulate_samples_run");
  return GuardAbi([&]() -> int { 
    if (job && !job->sub.empty()) {
      const size_t S = size_t(job->num_samples);
      const size_t row = S * size_t(job->nmax);
      return ForEachSub(job, [&](tfqb_job* sj, RowBlock b) {
        sj->ctx->row_offset = job->ctx->row_offset + b.lo;
        return impl_tfqb_simulate_samples_run(sj, seed, uniforms ? uniforms + size_t(b.lo) * S : nullptr,
                                              samples + size_t(b.lo) * row);
      });
    }
    return impl_tfqb_simulate_samples_run(job, seed, uniforms, samples); });
}

int tfqb_simulate_sampled_expectation(
    tfqb_context* ctx, const tfqb_circuit_inputs* in, tfqb_strings pauli_sums,
    int sum_rows, int n_ops, const int32_t* num_samples, int ns_rows,
    int ns_cols, uint64_t seed, const double* uniforms, int uniform_terms,
    int uniform_shots, float* expectations) {
  NvtxRange nvtx("tfqb_simulate_sampled_expectation");
  return GuardAbi([&]() -> int { 
    if (IsMulti(ctx))
      return MultiSampledExpectation(ctx, in, pauli_sums, sum_rows, n_ops, num_samples, ns_rows, ns_cols, seed,
                                     uniforms, uniform_terms, uniform_shots, expectations);
    return impl_tfqb_simulate_sampled_expectation(ctx, in, pauli_sums, sum_rows, n_ops, num_samples, ns_rows, ns_cols, seed, uniforms, uniform_terms, uniform_shots, expectations); });
}

int tfqb_calculate_unitary_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                                   tfqb_job** job, int* max_qubits) {
  return GuardAbi([&]() -> int {
    if (IsMulti(ctx) && in->batch >= 0 && in->symbol_rows == in->batch)
      return MultiPrepare(ctx, in, kJobUnitary,
                          [&](tfqb_context* c, const tfqb_circuit_inputs* sub, RowBlock, tfqb_job** j) {
                            return impl_tfqb_calculate_unitary_prepare(c, sub, j, nullptr);
                          }, job, max_qubits);
    if (IsMulti(ctx)) ctx = ctx->children[0];
    return impl_tfqb_calculate_unitary_prepare(ctx, in, job, max_qubits);
  });
}

int tfqb_calculate_unitary_run(tfqb_job* job, float* unitary) {
  NvtxRange nvtx("tfqb_calculate_unitary_run");
  return GuardAbi([&]() -> int {
    if (job && !job->sub.empty()) {
      const size_t row = (size_t(2) << job->nmax) << job->nmax;    // floats per output row
      return ForEachSub(job, [&](tfqb_job* sj, RowBlock b) {
        return impl_tfqb_calculate_unitary_run(sj, unitary + size_t(b.lo) * row);
      });
    }
    return impl_tfqb_calculate_unitary_run(job, unitary);
  });
}

int tfqb_noisy_expectation(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                           tfqb_strings pauli_sums, int sum_rows, int n_ops,
                           const int32_t* num_samples, int ns_rows, int ns_cols, uint64_t seed,
                           const float* uniforms, int uniform_trajectories, int uniform_channels,
                           float* expectations) {
  NvtxRange nvtx("tfqb_noisy_expectation");
  return GuardAbi([&]() -> int {
    if (IsMulti(ctx))
      return MultiNoisyExpectation(ctx, in, pauli_sums, sum_rows, n_ops, num_samples, ns_rows, ns_cols,
                                   seed, uniforms, uniform_trajectories, uniform_channels, false,
                                   expectations);
    return impl_tfqb_noisy_expectation(ctx, in, pauli_sums, sum_rows, n_ops, num_samples, ns_rows,
                                       ns_cols, seed, uniforms, uniform_trajectories,
                                       uniform_channels, expectations);
  });
}

int tfqb_noisy_sampled_expectation(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                                   tfqb_strings pauli_sums, int sum_rows, int n_ops,
                                   const int32_t* num_samples, int ns_rows, int ns_cols,
                                   uint64_t seed, const float* uniforms, int uniform_trajectories,
                                   int uniform_channels, float* expectations) {
  NvtxRange nvtx("tfqb_noisy_sampled_expectation");
  return GuardAbi([&]() -> int {
    if (IsMulti(ctx))
      return MultiNoisyExpectation(ctx, in, pauli_sums, sum_rows, n_ops, num_samples, ns_rows, ns_cols,
                                   seed, uniforms, uniform_trajectories, uniform_channels, true,
                                   expectations);
    return impl_tfqb_noisy_sampled_expectation(ctx, in, pauli_sums, sum_rows, n_ops, num_samples,
                                               ns_rows, ns_cols, seed, uniforms,
                                               uniform_trajectories, uniform_channels, expectations);
  });
}

int tfqb_noisy_samples_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in, int num_samples,
                               tfqb_job** job, int* max_qubits) {
  return GuardAbi([&]() -> int {
    if (IsMulti(ctx) && in->batch >= 0 && in->symbol_rows == in->batch && num_samples >= 0) {
      const int rc = MultiPrepare(ctx, in, kJobNoisySamples,
                          [&](tfqb_context* c, const tfqb_circuit_inputs* sub, RowBlock, tfqb_job** j) {
                            return impl_tfqb_noisy_samples_prepare(c, sub, num_samples, j, nullptr);
                          }, job, max_qubits);
      if (rc == TFQB_OK) (*job)->num_samples = num_samples;
      return rc;
    }
    if (IsMulti(ctx)) ctx = ctx->children[0];
    return impl_tfqb_noisy_samples_prepare(ctx, in, num_samples, job, max_qubits);
  });
}

int tfqb_noisy_samples_run(tfqb_job* job, uint64_t seed, const float* uniforms, int uniform_channels,
                           const double* measure_uniforms, int8_t* samples) {
  NvtxRange nvtx("tfqb_noisy_samples_run");
  return GuardAbi([&]() -> int {
    if (job && !job->sub.empty()) {
      const size_t S = size_t(job->num_samples);
      const size_t row = S * size_t(job->nmax);
      return ForEachSub(job, [&](tfqb_job* sj, RowBlock b) {
        sj->ctx->row_offset = job->ctx->row_offset + b.lo;
        return impl_tfqb_noisy_samples_run(
            sj, seed, uniforms ? uniforms + size_t(b.lo) * S * uniform_channels : nullptr,
            uniform_channels, measure_uniforms ? measure_uniforms + size_t(b.lo) * S : nullptr,
            samples + size_t(b.lo) * row);
      });
    }
    return impl_tfqb_noisy_samples_run(job, seed, uniforms, uniform_channels, measure_uniforms, samples);
  });
}

int tfqb_inner_product(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                       tfqb_strings other_programs, int other_rows,
                       int n_other, float* inner_products) {
  NvtxRange nvtx("tfqb_inner_product");
  return GuardAbi([&]() -> int { 
    if (IsMulti(ctx)) return MultiInnerProduct(ctx, in, other_programs, other_rows, n_other, inner_products);
    return impl_tfqb_inner_product(ctx, in, other_programs, other_rows, n_other, inner_products); });
}

int tfqb_inner_product_grad(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                            tfqb_strings other_programs, int other_rows,
                            int n_other, const float* downstream, int grad_rows,
                            int grad_cols, float* grads) {
  NvtxRange nvtx("tfqb_inner_product_grad");
  return GuardAbi([&]() -> int { 
    if (IsMulti(ctx))
      return MultiInnerProductGrad(ctx, in, other_programs, other_rows, n_other, downstream, grad_rows, grad_cols, grads);
    return impl_tfqb_inner_product_grad(ctx, in, other_programs, other_rows, n_other, downstream, grad_rows, grad_cols, grads); });
}

int tfqb_sharded_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                         tfqb_strings pauli_sums, int n_ops, int world,
                         int rank, tfqb_job** job, int* n_stages, int* n_terms) {
  return GuardAbi([&]() -> int { 
    if (IsMulti(ctx))
      return Fail(TFQB_INVALID_ARGUMENT,
                  "tfqb_sharded_* takes one single-device context per rank (tfqb_create)");
    return impl_tfqb_sharded_prepare(ctx, in, pauli_sums, n_ops, world, rank, job, n_stages, n_terms); });
}

int tfqb_sharded_stage_kind(tfqb_job* job, int stage) {
  return GuardAbi([&]() -> int { return impl_tfqb_sharded_stage_kind(job, stage); });
}

int tfqb_sharded_buffers(tfqb_job* job, void** send, void** recv, size_t* bytes) {
  return GuardAbi([&]() -> int { return impl_tfqb_sharded_buffers(job, send, recv, bytes); });
}

int tfqb_sharded_run_stage(tfqb_job* job, int stage) {
  return GuardAbi([&]() -> int { return impl_tfqb_sharded_run_stage(job, stage); });
}

int tfqb_sharded_partials(tfqb_job* job, double* per_term) {
  return GuardAbi([&]() -> int { return impl_tfqb_sharded_partials(job, per_term); });
}

int tfqb_sharded_finish(tfqb_job* job, const double* per_term_total,
                        float* expectations) {
  return GuardAbi([&]() -> int { return impl_tfqb_sharded_finish(job, per_term_total, expectations); });
}

int tfqb_sharded_export(tfqb_job* job, unsigned char* handle) {
  return GuardAbi([&]() -> int { return impl_tfqb_sharded_export(job, handle); });
}

int tfqb_sharded_connect(tfqb_job* job, const unsigned char* handles, int world) {
  return GuardAbi([&]() -> int { return impl_tfqb_sharded_connect(job, handles, world); });
}

int tfqb_sharded_enqueue(tfqb_job* job) {
  NvtxRange nvtx("tfqb_sharded_enqueue");
  return GuardAbi([&]() -> int { return impl_tfqb_sharded_enqueue(job); });
}

int tfqb_sharded_result(tfqb_job* job, float* expectations) {
  return GuardAbi([&]() -> int { return impl_tfqb_sharded_result(job, expectations); });
}

int tfqb_sharded_sample(tfqb_job* job, int num_samples, uint64_t seed, const double* uniforms,
                        int8_t* samples, int32_t* owned) {
  return GuardAbi([&]() -> int {
    return impl_tfqb_sharded_sample(job, num_samples, seed, uniforms, samples, owned);
  });
}

int tfqb_sharded_stats(tfqb_job* job, tfqb_exchange_stats* out) {
  return GuardAbi([&]() -> int { return impl_tfqb_sharded_stats(job, out); });
}

int tfqb_sync(tfqb_context* ctx) {
  return GuardAbi([&]() -> int { 
    if (IsMulti(ctx)) {
      for (tfqb_context* c : ctx->children) TFQB_RETURN_IF(impl_tfqb_sync(c));
      return TFQB_OK;
    }
    return impl_tfqb_sync(ctx); });
}

int tfqb_profile_enable(tfqb_context* ctx, int enable) {
  return GuardAbi([&]() -> int { 
    if (IsMulti(ctx)) {
      for (tfqb_context* c : ctx->children) TFQB_RETURN_IF(impl_tfqb_profile_enable(c, enable));
      return TFQB_OK;
    }
    return impl_tfqb_profile_enable(ctx, enable); });
}

int tfqb_profile_reset(tfqb_context* ctx) {
  return GuardAbi([&]() -> int { 
    if (IsMulti(ctx)) {
      for (tfqb_context* c : ctx->children) TFQB_RETURN_IF(impl_tfqb_profile_reset(c));
      return TFQB_OK;
    }
    return impl_tfqb_profile_reset(ctx); });
}

int tfqb_profile_read(tfqb_context* ctx, tfqb_profile* out) {
  return GuardAbi([&]() -> int { 
    if (IsMulti(ctx)) {       // counters and times summed over the devices
      tfqb_profile tot{};
      for (tfqb_context* c : ctx->children) {
        tfqb_profile p{};
        TFQB_RETURN_IF(impl_tfqb_profile_read(c, &p));
        tot.kernel_launches += p.kernel_launches;
        tot.gate_pass_launches += p.gate_pass_launches;
        tot.adjoint_pass_launches += p.adjoint_pass_launches;
        tot.gate_pass_ms += p.gate_pass_ms;
        tot.adjoint_pass_ms += p.adjoint_pass_ms;
        tot.gate_pass_bytes += p.gate_pass_bytes;
        tot.adjoint_pass_bytes += p.adjoint_pass_bytes;
        tot.h2d_bytes += p.h2d_bytes;
        tot.d2h_bytes += p.d2h_bytes;
        tot.expectation_launches += p.expectation_launches;
        tot.expectation_ms += p.expectation_ms;
        tot.expectation_bytes += p.expectation_bytes;
        tot.jit_kernels += p.jit_kernels;
        tot.jit_pass_launches += p.jit_pass_launches;
      }
      *out = tot;
      return TFQB_OK;
    }
    return impl_tfqb_profile_read(ctx, out); });
}

int tfqb_host_gate_matrix(int kind, const float* params, int n_params,
                          int grad_param, float* out) {
  return GuardAbi([&]() -> int { return impl_tfqb_host_gate_matrix(kind, params, n_params, grad_param, out); });
}

int tfqb_host_describe_plan(const char* program, size_t program_size,
                            tfqb_strings symbol_names, int n_symbols,
                            int adjoint, char** json_out) {
  return GuardAbi([&]() -> int { return impl_tfqb_host_describe_plan(program, program_size, symbol_names, n_symbols, adjoint, json_out); });
}

int tfqb_host_jit_source(const char* program, size_t program_size,
                         tfqb_strings symbol_names, int n_symbols,
                         int adjoint, int pass, char** source_out) {
  return GuardAbi([&]() -> int { return impl_tfqb_host_jit_source(program, program_size, symbol_names, n_symbols, adjoint, pass, source_out); });
}

int tfqb_host_describe_pauli_sum(const char* program, size_t program_size,
                                 const char* pauli_sum, size_t pauli_sum_size,
                                 char** json_out) {
  return GuardAbi([&]() -> int { return impl_tfqb_host_describe_pauli_sum(program, program_size, pauli_sum, pauli_sum_size, json_out); });
}

int tfqb_host_jit_expect_source(const char* program, size_t program_size,
                                tfqb_strings pauli_sums, int n_ops, int pass,
                                char** source_out) {
  return GuardAbi([&]() -> int { return impl_tfqb_host_jit_expect_source(program, program_size, pauli_sums, n_ops, pass, source_out); });
}

int tfqb_host_describe_sharded(const char* program, size_t program_size,
                               tfqb_strings symbol_names, int n_symbols,
                               tfqb_strings pauli_sums, int n_ops, int world,
                               char** json_out) {
  return GuardAbi([&]() -> int { return impl_tfqb_host_describe_sharded(program, program_size, symbol_names, n_symbols, pauli_sums, n_ops, world, json_out); });
}


// ---- parameter-shift helper ops (next-row N4; host only) ---------------------
namespace {
int ParseAll(tfqb_strings programs, int batch, std::vector<tfqb::ProgramPB>* out) {
  if (batch < 0) return Fail(TFQB_INVALID_ARGUMENT, "negative tensor dimension");
  out->resize(size_t(batch));
  for (int i = 0; i < batch; ++i) {
    const char* d = programs.data[i];
    const size_t n = programs.size[i];
    if (!tfqb::ParseProgram(d, n, &(*out)[i]))
      return Fail(TFQB_INVALID_ARGUMENT,
                        "Unparseable proto: " + std::string(d, std::min<size_t>(n, 64)));
  }
  return TFQB_OK;
}
int FillList(const std::vector<std::string>& v, tfqb_string_list* out) {
  out->count = v.size();
  out->data = static_cast<char**>(calloc(std::max<size_t>(v.size(), 1), sizeof(char*)));
  out->size = static_cast<size_t*>(calloc(std::max<size_t>(v.size(), 1), sizeof(size_t)));
  if (!out->data || !out->size) return Fail(TFQB_RESOURCE_EXHAUSTED, "Out of host memory.");
  for (size_t i = 0; i < v.size(); ++i) {
    out->data[i] = static_cast<char*>(malloc(std::max<size_t>(v[i].size(), 1)));
    if (!out->data[i]) return Fail(TFQB_RESOURCE_EXHAUSTED, "Out of host memory.");
    memcpy(out->data[i], v[i].data(), v[i].size());
    out->size[i] = v[i].size();
  }
  return TFQB_OK;
}
}  // namespace

int tfqb_ps_decompose(tfqb_strings programs, int batch, tfqb_string_list* out) {
  return GuardAbi([&]() -> int {
    std::vector<tfqb::ProgramPB> progs;
    TFQB_RETURN_IF(ParseAll(programs, batch, &progs));
    std::vector<std::string> res(progs.size());
    for (size_t i = 0; i < progs.size(); ++i) {
      tfqb::ProgramPB dec;
      tfqb::Status st = tfqb::PsDecompose(progs[i], &dec);
      if (!st.ok) return Fail(TFQB_INVALID_ARGUMENT, st.msg);
      res[i] = tfqb::EncodeProgram(dec);
    }
    return FillList(res, out);
  });
}

int tfqb_ps_symbol_replace(tfqb_strings programs, int batch, tfqb_strings symbols, int n_symbols,
                           tfqb_strings replacement_symbols, int n_replacements,
                           tfqb_string_list* out, int* pad) {
  return GuardAbi([&]() -> int {
    if (n_symbols != n_replacements)
      return Fail(TFQB_INVALID_ARGUMENT,
                        "symbols.shape is not equal to replacement_symbols.shape: " +
                            std::to_string(n_symbols) + " != " + std::to_string(n_replacements));
    std::vector<tfqb::ProgramPB> progs;
    TFQB_RETURN_IF(ParseAll(programs, batch, &progs));
    // (i, j, k) = the kth replaced program for symbols(j) in programs(i)
    std::vector<std::vector<std::string>> found(progs.size() * size_t(n_symbols));
    size_t biggest = 0;
    for (size_t i = 0; i < progs.size(); ++i)
      for (int j = 0; j < n_symbols; ++j) {
        auto& v = found[i * n_symbols + j];
        tfqb::PsSymbolReplace(progs[i], std::string(symbols.data[j], symbols.size[j]),
                              std::string(replacement_symbols.data[j], replacement_symbols.size[j]),
                              &v);
        biggest = std::max(biggest, v.size());
      }
    const std::string empty = tfqb::EmptyProgram();
    std::vector<std::string> flat;
    flat.reserve(found.size() * biggest);
    for (auto& v : found)
      for (size_t k = 0; k < biggest; ++k) flat.push_back(k < v.size() ? v[k] : empty);
    *pad = int(biggest);
    return FillList(flat, out);
  });
}

int tfqb_ps_weights_from_symbols(tfqb_strings programs, int batch, tfqb_strings symbols,
                                 int n_symbols, float** weights, int* pad) {
  return GuardAbi([&]() -> int {
    std::vector<tfqb::ProgramPB> progs;
    TFQB_RETURN_IF(ParseAll(programs, batch, &progs));
    std::vector<std::string> names;
    for (int j = 0; j < n_symbols; ++j) names.emplace_back(symbols.data[j], symbols.size[j]);
    std::vector<std::vector<std::vector<float>>> all(progs.size());
    size_t biggest = 0;
    for (size_t i = 0; i < progs.size(); ++i) {
      tfqb::Status st = tfqb::PsWeightsFromSymbols(progs[i], names, &all[i]);
      if (!st.ok) return Fail(TFQB_INVALID_ARGUMENT, st.msg);
      for (auto& v : all[i]) biggest = std::max(biggest, v.size());
    }
    const size_t total = progs.size() * size_t(n_symbols) * biggest;
    float* w = static_cast<float*>(calloc(std::max<size_t>(total, 1), sizeof(float)));
    if (!w) return Fail(TFQB_RESOURCE_EXHAUSTED, "Out of host memory.");
    for (size_t i = 0; i < progs.size(); ++i)
      for (int j = 0; j < n_symbols; ++j)
        for (size_t k = 0; k < all[i][j].size(); ++k)
          w[(i * n_symbols + j) * biggest + k] = all[i][j][k];
    *weights = w;
    *pad = int(biggest);
    return TFQB_OK;
  });
}

void tfqb_free_string_list(tfqb_string_list* l) {
  if (!l) return;
  for (size_t i = 0; i < l->count; ++i) free(l->data ? l->data[i] : nullptr);
  free(l->data);
  free(l->size);
  l->data = nullptr;
  l->size = nullptr;
  l->count = 0;
}
void tfqb_free_floats(float* p) { free(p); }

}  // extern "C"
