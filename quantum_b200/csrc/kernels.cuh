// Hand-written sm_100a kernels for the TFQ state-vector hot path
// (SURVEY.md §2b table / §8a rows Q1-Q3, K1-K3).  Launch wrappers are
// declared here and defined in kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "plan.h"

namespace tfqb {

// Mask-form Pauli term as the device sees it.
struct DevTerm {
  uint64_t x, z;
  float coeff;
  int32_t phase;     // power of i
  int32_t op;        // which PauliSum (column j) the term belongs to
  int32_t identity;  // 1: adds coeff (util_qsim.h:154-158)
};

struct PassLaunch {
  const PassRec* passes;    // device arrays of the plan
  const RoundRec* rounds;
  const OpRec* ops;
  const float* mats;        // [rows][mat_floats] or [mat_floats]
  size_t mat_row_stride;    // 0 when matrices are row independent
  int pass_index;
  int tile_bits;            // host copies of pass fields used for launch dims
  int low_bits = kLowBits;
  int n_alloc;
  int n_ops_in_pass;
  int first_op;
  int mat_len;
  int n_rounds;
  int reg_bits;      // register bits per round the plan was built for
  unsigned long long rank_base = 0;  // index bits above the local shard
  // init_mode 3 (sharded states): the pass that follows a global<->local qubit
  // swap loads its tiles straight from the peers' shards -- amplitude g of the
  // swapped shard is amplitude (rank << peer_shift | g & mask) of the shard of
  // rank g >> peer_shift -- and stores them into this rank's own buffer: the
  // all-to-all IS the pass's load phase (peer memory over NVLink)
  const float2* const* peer_tab = nullptr;   // device array of `world` shard bases
  int peer_shift = 0;
  unsigned long long peer_self = 0;          // rank << peer_shift
};

// TFQB_DETERMINISTIC=1: reductions across CTAs run in one CTA per row (fixed
// order, bit-reproducible results); see kernels.cu
bool Deterministic();

// --- gate passes (Q1): one read+write sweep of `rows` states -------------
// init_mode: 0 load the state, 1 synthesise |0..0>, 2 synthesise the plan's
// product state (pass 0 of a forward plan), 3 gather the tiles from the peers
// of a sharded state (PassLaunch::peer_tab)
void LaunchForwardPass(const PassLaunch& pl, float2* psi, size_t row_stride,
                       int rows, int init_mode, cudaStream_t s);
void LaunchAdjointPass(const PassLaunch& pl, float2* psi, float2* lam,
                       size_t row_stride, int rows, double* grad_out,
                       int n_slots, cudaStream_t s);
size_t ForwardPassSmem(int tile_bits, int mat_len, int n_ops, int n_rounds);
size_t AdjointPassSmem(int tile_bits, int mat_len, int n_ops, int n_rounds);

// --- per-row matrix evaluation (H5/H8 on device) --------------------------
void LaunchBuildMatrices(const MatRec* recs, const FactorRec* factors,
                         int n_recs, const float* params, int n_params,
                         int rows, float* out, size_t out_row_stride,
                         cudaStream_t s);

// --- Q2 state-space primitives -------------------------------------------
void LaunchSetZeroState(float2* psi, size_t row_stride, int rows,
                        cudaStream_t s);
// out[row, 0:2^n] = psi ; out[row, 2^n:out_cols] = (-2, 0)
void LaunchExportState(const float2* psi, size_t row_stride, int n,
                       float2* out, size_t out_cols, int rows, cudaStream_t s);

// --- K1: PauliSum expectation --------------------------------------------
// partial[row, term] (fp64) = Re<psi| P_term |psi>, reduced over the state.
// generic path: one read of the state (plus partner gather) per term; `subset`
// (device, may be null) restricts the launch to those term indices.
void LaunchExpectationTerms(const float2* psi, size_t row_stride, int n_alloc,
                            const DevTerm* terms, int n_terms,
                            const int32_t* subset, int n_subset, int rows,
                            double* per_term, cudaStream_t s);
// fast path: one tile-staged pass of an ExpectationPlan (plan.h)
struct ExpectLaunch {
  const PassRec* passes;
  const RoundRec* rounds;
  const ExpXOp* xops;
  const ExpZTerm* zterms;
  int n_zterms;      // > 0 only for pass 0
  int pass_index;
  int tile_bits;
  int low_bits = kLowBits;
  int n_alloc;
  int n_xops;        // in this pass
  int n_rounds;      // in this pass
  int n_terms;       // size of a per_term row
  unsigned long long rank_base = 0;  // index bits above the local shard
};
void LaunchExpectPass(const ExpectLaunch& el, const float2* psi, size_t row_stride,
                      int rows, double* per_term, cudaStream_t s);
// fast path of K3 over the same plan (built with identity terms as z = 0)
void LaunchAccumPass(const ExpectLaunch& el, const float2* psi, float2* lam,
                     size_t row_stride, int rows, const DevTerm* terms,
                     const float* downstream, int n_ops, bool accumulate,
                     cudaStream_t s);
// out[row, j] = float( sum_t coeff_t * per_term[row, t] ) (+ identities)
void LaunchCombineTerms(const double* per_term, const DevTerm* terms,
                        int n_terms, int n_ops, int rows, float* out,
                        size_t out_stride, cudaStream_t s);

// out[row] (+)= <psi_row | phi> as (re, im) doubles; `out` must be zeroed
void LaunchInnerProduct(const float2* psi, size_t row_stride, const float2* phi,
                        int n_alloc, int rows, double* out, cudaStream_t s);

// lam[row] (first ? = : +=) coeff[row * coeff_stride] * (times_i ? i : 1) * phi
void LaunchAxpyRows(float2* lam, size_t row_stride, const float2* phi, int n_alloc,
                    const float* coeff, int coeff_stride, bool times_i, bool first,
                    int rows, cudaStream_t s);

// --- K3: lambda = sum_j g_j sum_t c_t P_t psi -----------------------------
// generic path (global partner gather); `subset` restricts the terms,
// `accumulate` adds to lambda instead of overwriting it
void LaunchAccumulateOperators(const float2* psi, float2* lam,
                               size_t row_stride, int n_alloc,
                               const DevTerm* terms, int n_terms,
                               const int32_t* subset, int n_subset,
                               bool accumulate,
                               const float* downstream, int n_ops, int rows,
                               cudaStream_t s);

// --- O5 epilogue: grads[row, col] = float(sum of slots mapped to col) -----
void LaunchReduceGradSlots(const double* slot_vals, const int32_t* slot_col,
                           int n_slots, int rows, float* grads, int n_cols,
                           cudaStream_t s);

// --- Q3: sampling ---------------------------------------------------------
// Canonical fp64 pairwise probability tree (see oracle sample_tree): levels
// >= kTreeChunkBits are stored, lower levels are recomputed per shot.
constexpr int kTreeChunkBits = 8;
size_t TreeDoublesPerRow(int n_alloc);
void LaunchBuildTree(const float2* psi, size_t row_stride, int n_alloc,
                     double* tree, int rows, cudaStream_t s);
// uniforms: [rows, uniform_row_stride] in [0,1); shots_per_row (device, may
// be null) limits the shots of each row. indices out: [rows, shots].
void LaunchSample(const float2* psi, size_t row_stride, int n_alloc,
                  const double* tree, const double* uniforms,
                  size_t uniform_row_stride, const int32_t* shots_per_row,
                  int shots, int rows, uint64_t* indices,
                  size_t index_row_stride, cudaStream_t s);
// Philox4x32-10(seed), counter (shot, row_id, stream_a, stream_b); entries
// s >= shots of the padded row are set to 2.0 so they sort to the end.
void LaunchFillUniforms(double* u, size_t row_stride, uint64_t seed,
                        const int32_t* row_ids, uint32_t stream_a,
                        uint32_t stream_b, int shots, int rows, cudaStream_t s);
// ascending in-place sort of each row; row_stride must be a power of two
void LaunchSortRows(double* u, size_t row_stride, int rows, cudaStream_t s);
// out[row, shot, nmax-1-q] = bit q of index (q < n) else -2
void LaunchUnpackSamples(const uint64_t* indices, size_t index_row_stride,
                         int n, int nmax, int shots, int rows, int8_t* out,
                         cudaStream_t s);
// acc[row] += coeff * (sum_s parity sign) / shots  (util_qsim.h:241-267)
void LaunchParityExpectation(const uint64_t* indices, size_t index_row_stride,
                             uint64_t mask, float coeff,
                             const int32_t* shots_per_row, int shots, int rows,
                             float* acc, size_t acc_stride, cudaStream_t s);
// acc[row] += c  (identity terms, util_qsim.h:154-158)
void LaunchAddConstant(float c, int rows, float* acc, size_t acc_stride,
                       cudaStream_t s);

// --- peer-memory exchange of a sharded state (kernels.cu, last section) -----
// flag = value with release.sys semantics, in stream order
void LaunchPeerSignal(unsigned* flag, unsigned value, cudaStream_t s);
// wait (bounded) until *flags[r] >= value for every rank r != self
void LaunchPeerWait(const unsigned* const* flags, int world, int self, unsigned value,
                    unsigned long long timeout_ns, int* error, cudaStream_t s);
// dst chunk r <- chunk `rank` of peers[r] (chunk_amps amplitudes each)
void LaunchPeerPull(float2* dst, const float2* const* peers, int world, int rank,
                    size_t chunk_amps, cudaStream_t s);
void LaunchPeerPublishPartials(const double* src, double* dst, int n, cudaStream_t s);
void LaunchPeerReducePartials(const double* const* parts, int world, int n, double* out,
                              cudaStream_t s);

// --- noisy trajectory ops (next-row N2): rows are (circuit, trajectory) -----
// params[row] = [symbol values of the row's circuit | one uniform per channel]
void LaunchNoisyFillParams(float* params, int cols, int P, int C, const float* symbol_values,
                           const int32_t* sym_row, const int32_t* circuit_id,
                           const int32_t* trajectory, const float* given_uniforms,
                           const long long* given_offset, uint64_t seed, int rows,
                           cudaStream_t s);
// u[row, k < count] = Philox(seed; (k, circuit, trajectory, stream)) as doubles
void LaunchNoisyFillUniforms(double* u, size_t stride, int count, const int32_t* circuit_id,
                             const int32_t* trajectory, uint32_t stream, uint64_t seed, int rows,
                             cudaStream_t s);
// params[row, col] = normalised population of |1> on index bit `bit`;
// acc: scratch double[rows, 2]
void LaunchPopulation(const float2* psi, size_t row_stride, int n_alloc, int bit, double* acc,
                      float* params, int cols, int col, int rows, cudaStream_t s);

// --- sampling from a sharded state: shard norm, all ranks' norms ------------
void LaunchTreeTotal(const double* tree, int n_alloc, double* out, cudaStream_t s);
void LaunchPeerGatherScalars(const double* const* parts, int world, double* out, cudaStream_t s);

// --- TfqCalculateUnitary (next-row N4): basis states in, columns out ---------
void LaunchBasisStates(float2* psi, size_t row_stride, size_t first, int rows, cudaStream_t s);
void LaunchExportUnitary(const float2* psi, size_t row_stride, size_t dim, size_t k0, int cols,
                         float2* out, size_t out_dim, cudaStream_t s);
void LaunchFillPad(float2* out, size_t count, cudaStream_t s);

// load every kernel a sharded job launches (see kernels.cu)
void PreloadShardedKernels();

}  // namespace tfqb
