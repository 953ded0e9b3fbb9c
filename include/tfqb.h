/* tfqb.h — C ABI of the B200 state-vector backend for TensorFlow Quantum's
 * circuit-execution ops.
 *
 * Every `tfqb_simulate_*` / `tfqb_adjoint_gradient` entry point is the body of
 * one reference OpKernel::Compute and takes exactly that op's inputs as plain
 * host pointers (serialized protos as (pointer, length) pairs, dense tensors
 * row-major).  A TF `OpKernel` registered for DEVICE_GPU with
 * `.HostMemory(...)` on every input forwards its tensors 1:1
 * (quantum_b200/csrc/tf_ops/tfq_b200_ops.cc).
 *
 *   tfqb_simulate_expectation          TfqSimulateExpectationOp::Compute
 *       tensorflow_quantum/core/ops/tfq_simulate_expectation_op.cc:50-250
 *   tfqb_simulate_sampled_expectation  TfqSimulateSampledExpectationOp::Compute
 *       tensorflow_quantum/core/ops/tfq_simulate_sampled_expectation_op.cc:54-306
 *   tfqb_simulate_samples              TfqSimulateSamplesOp::Compute
 *       tensorflow_quantum/core/ops/tfq_simulate_samples_op.cc:53-252
 *   tfqb_simulate_state                TfqSimulateStateOp::Compute
 *       tensorflow_quantum/core/ops/tfq_simulate_state_op.cc:48-216
 *   tfqb_adjoint_gradient              TfqAdjointGradientOp::Compute
 *       tensorflow_quantum/core/ops/tfq_adj_grad_op.cc:51-390
 *
 * All functions return 0 on success or a TF-style status code
 * (TFQB_INVALID_ARGUMENT = 3 mirrors tf.errors.InvalidArgumentError); the
 * message — same substrings the reference tests assert — is available from
 * tfqb_last_error().  There is NO CPU fallback: without a CUDA device
 * tfqb_create fails and every compute entry point returns TFQB_UNAVAILABLE.
 */
#ifndef TFQB_H_
#define TFQB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFQB_OK 0
#define TFQB_INVALID_ARGUMENT 3
#define TFQB_RESOURCE_EXHAUSTED 8
#define TFQB_INTERNAL 13
#define TFQB_UNAVAILABLE 14

typedef struct tfqb_context tfqb_context;
typedef struct tfqb_job tfqb_job;

/* A batch of serialized strings (a tf.string tensor's contents). */
typedef struct {
  const char* const* data;
  const size_t* size;
} tfqb_strings;

/* The inputs shared by all five ops (REGISTER_OP signatures,
 * tfq_simulate_expectation_op.cc:257-283 etc.). */
typedef struct {
  tfqb_strings programs;        /* string[batch] */
  int batch;
  tfqb_strings symbol_names;    /* string[n_symbols] */
  int n_symbols;
  const float* symbol_values;   /* float[symbol_rows, n_symbols] */
  int symbol_rows;              /* must equal batch ("... do not match") */
} tfqb_circuit_inputs;

int tfqb_abi_version(void);

/* One context per (process, GPU). `device` is a CUDA ordinal. */
int tfqb_create(int device, tfqb_context** out);
/* One context over SEVERAL GPUs of this process.  The reference spreads one
 * OpKernel::Compute over every host core, rows first
 * (tfq_simulate_expectation_op.cc:245-248, tfq_adj_grad_op.cc:282-283); a
 * multi-device context spreads one tfqb_simulate_* / tfqb_adjoint_gradient /
 * tfqb_inner_product* call over every listed GPU the same way: contiguous row
 * blocks, one host thread per device, each block written into its slice of
 * the caller's output tensor, no collective.  Results are identical to the
 * single-device call (the sampling ops key their Philox streams by global
 * row).  tfqb_sharded_* takes single-device contexts. */
int tfqb_create_multi(const int* device_ids, int n_devices, tfqb_context** out);
int tfqb_device_count(tfqb_context* ctx);
/* Global index of the first row this context is given (default 0): a batch
 * that the CALLER splits over processes (one rank per GPU) draws, with the
 * same seed, the same uniforms as the unsplit batch. */
int tfqb_set_row_offset(tfqb_context* ctx, int64_t first_row);
void tfqb_destroy(tfqb_context* ctx);
/* Last error of the calling thread (valid until the next failing call). */
const char* tfqb_last_error(void);
/* Limit device memory used for state vectors (bytes; 0 = 80% of free). */
int tfqb_set_memory_budget(tfqb_context* ctx, size_t bytes);
/* Give the device memory the context caches between calls back to CUDA. */
int tfqb_trim(tfqb_context* ctx);
/* Host seconds this process has spent in NVRTC for the run-time specialised
 * kernels (the cold-start cost of a new circuit structure; compilations run
 * on parallel host threads, so wall time is lower). */
double tfqb_jit_compile_seconds(void);
/* Kernel compilations running on background host threads right now (a job
 * that first qualifies hands ALL its passes to the compiler; the ones it does
 * not launch yet keep host cores busy for a few seconds). */
int tfqb_jit_pending(void);

/* expectations: float[batch, n_ops]; pauli_sums: string[sum_rows, n_ops]. */
int tfqb_simulate_expectation(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                              tfqb_strings pauli_sums, int sum_rows, int n_ops,
                              float* expectations);

/* num_samples: int32[ns_rows, ns_cols] (every entry >= 1). `seed` drives the
 * device Philox stream (the TF shim passes random::New64()).  `uniforms`, if
 * non-NULL, replaces the stream for parity tests: double[batch, n_ops,
 * uniform_terms, uniform_shots] in [0,1), term-major as in the PauliSum. */
int tfqb_simulate_sampled_expectation(
    tfqb_context* ctx, const tfqb_circuit_inputs* in, tfqb_strings pauli_sums,
    int sum_rows, int n_ops, const int32_t* num_samples, int ns_rows,
    int ns_cols, uint64_t seed, const double* uniforms, int uniform_terms,
    int uniform_shots, float* expectations);

/* Two-step because the output shape [batch, num_samples, max_qubits] depends
 * on the parsed programs: _prepare parses and returns max_qubits, _run fills
 * the caller-allocated int8 tensor, tfqb_job_free releases the job.
 * `uniforms` (optional): double[batch, num_samples] in [0,1); they are sorted
 * ascending per row before use, as qsim sorts its draws. */
int tfqb_simulate_samples_prepare(tfqb_context* ctx,
                                  const tfqb_circuit_inputs* in,
                                  int num_samples, tfqb_job** job,
                                  int* max_qubits);
int tfqb_simulate_samples_run(tfqb_job* job, uint64_t seed,
                              const double* uniforms, int8_t* samples);

/* state_vector: complex64[batch, 2^max_qubits] as interleaved floats. */
int tfqb_simulate_state_prepare(tfqb_context* ctx,
                                const tfqb_circuit_inputs* in, tfqb_job** job,
                                int* max_qubits);
int tfqb_simulate_state_run(tfqb_job* job, float* state_vector);

/* grads: float[batch, n_symbols]; downstream_grads: float[grad_rows,
 * grad_cols]. */
int tfqb_adjoint_gradient(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                          tfqb_strings pauli_sums, int sum_rows, int n_ops,
                          const float* downstream_grads, int grad_rows,
                          int grad_cols, float* grads);

/* ---- "next" row N1 of SURVEY.md 8(f): TfqInnerProductOp::Compute
 * (tensorflow_quantum/core/ops/math_ops/tfq_inner_product.cc:45-292).
 * other_programs: string[other_rows, n_other], symbol free, on the same
 * qubits as programs[i]; inner_products: complex64[batch, n_other] as
 * interleaved floats, <psi_i | phi_ij>; (1, 0) where programs[i] is empty. */
int tfqb_inner_product(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                       tfqb_strings other_programs, int other_rows,
                       int n_other, float* inner_products);
/* TfqInnerProductGradOp::Compute (math_ops/tfq_inner_product_grad.cc:46-501):
 * downstream_grads: float[grad_rows, grad_cols] (= [batch, n_other]);
 * grads: complex64[batch, n_symbols] as interleaved floats,
 * grads[i, p] = sum over the gradient gates of symbol p of
 * <dG psi' | sum_j downstream[i, j] phi_ij> (what the op returns; the Python
 * wrapper inner_product_op.py:66-70 conjugates it).  Rows whose program is
 * empty stay 0.  n_symbols must be positive (:64-67). */
int tfqb_inner_product_grad(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                            tfqb_strings other_programs, int other_rows,
                            int n_other, const float* downstream_grads,
                            int grad_rows, int grad_cols, float* grads);

/* ---- "next" row N2 of SURVEY.md 8(f): the noisy trajectory ops
 *   tfqb_noisy_expectation          TfqNoisyExpectationOp::Compute
 *       tensorflow_quantum/core/ops/noise/tfq_noisy_expectation.cc:57-391
 *   tfqb_noisy_sampled_expectation  TfqNoisySampledExpectationOp::Compute
 *       tensorflow_quantum/core/ops/noise/tfq_noisy_sampled_expectation.cc:57-404
 *   tfqb_noisy_samples_*            TfqNoisySamplesOp::Compute
 *       tensorflow_quantum/core/ops/noise/tfq_noisy_samples.cc:54-321
 * Programs may hold the channels of circuit_parser_qsim.cc:752-756 (DP ADP
 * GAD AD RST PD PF BF).  expectations[i, j] = mean over the first
 * num_samples[i, j] trajectories of row i of <psi_t|O_j|psi_t> (exact, or from
 * one measured shot per term per trajectory for the sampled variant); -2
 * where the program is empty.  The trajectories are rows of the same batched
 * kernels.  `seed` keys the Philox streams: uniform of (row i, trajectory t,
 * channel c) = counter (c, i, t, "nois"); `uniforms`, if non-NULL, replaces
 * them for parity tests: float[batch, uniform_trajectories, uniform_channels],
 * channels in program order. */
int tfqb_noisy_expectation(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                           tfqb_strings pauli_sums, int sum_rows, int n_ops,
                           const int32_t* num_samples, int ns_rows, int ns_cols,
                           uint64_t seed, const float* uniforms,
                           int uniform_trajectories, int uniform_channels,
                           float* expectations);
int tfqb_noisy_sampled_expectation(tfqb_context* ctx,
                                   const tfqb_circuit_inputs* in,
                                   tfqb_strings pauli_sums, int sum_rows,
                                   int n_ops, const int32_t* num_samples,
                                   int ns_rows, int ns_cols, uint64_t seed,
                                   const float* uniforms,
                                   int uniform_trajectories,
                                   int uniform_channels, float* expectations);
/* samples: int8[batch, num_samples, max_qubits]; shot s of row i is the
 * terminal measurement of trajectory s.  uniforms: float[batch, num_samples,
 * uniform_channels]; measure_uniforms: double[batch, num_samples] (optional). */
int tfqb_noisy_samples_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                               int num_samples, tfqb_job** job, int* max_qubits);
int tfqb_noisy_samples_run(tfqb_job* job, uint64_t seed, const float* uniforms,
                           int uniform_channels, const double* measure_uniforms,
                           int8_t* samples);

/* ---- "next" row N4 of SURVEY.md 8(f): TfqCalculateUnitaryOp::Compute
 * (tensorflow_quantum/core/ops/tfq_calculate_unitary_op.cc:47-164).
 * unitary: complex64[batch, 2^max_qubits, 2^max_qubits] as interleaved
 * floats, unitary[i, j, k] = <j| U_i |k>, (-2, 0) outside a smaller circuit's
 * block; the columns are computed as one batch of 2^n basis states.  (The
 * parameter-shift helper ops of N4, tfq_ps_*_op.cc, rewrite protos on the
 * host and are not part of this library.) */
int tfqb_calculate_unitary_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                                   tfqb_job** job, int* max_qubits);
int tfqb_calculate_unitary_run(tfqb_job* job, float* unitary);

/* ---- device-resident variants (parse/plan/upload once, then run on data
 * already in HBM; used by bench.py for the kernel-only `value`). ---------- */
int tfqb_expectation_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                             tfqb_strings pauli_sums, int sum_rows, int n_ops,
                             tfqb_job** job);
int tfqb_adjoint_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                         tfqb_strings pauli_sums, int sum_rows, int n_ops,
                         const float* downstream_grads, int grad_rows,
                         int grad_cols, tfqb_job** job);
/* Enqueue the device work of a prepared expectation/adjoint job on the
 * context stream (no host<->device copies, no host sync). */
int tfqb_job_run_device(tfqb_job* job);
/* Wait for the stream and copy the result tensor to `out`
 * (float[batch, n_ops] or float[batch, n_symbols]). */
int tfqb_job_fetch(tfqb_job* job, float* out);
void tfqb_job_free(tfqb_job* job);

/* ---- ONE state sharded over world = 2^g ranks by its top ("global") qubits
 * (the reference's single-circuit path, ComputeLarge,
 * tfq_simulate_expectation_op.cc:130-180, has no multi-device form; this is
 * new).  Every rank calls with identical inputs (batch must be 1) and its own
 * `rank`.  The job is a list of stages:
 *   kind 0  gate segment   : tfqb_sharded_run_stage enqueues local passes
 *   kind 1  qubit exchange : the HOST performs one all-to-all over the ranks
 *           (NCCL / torch.distributed all_to_all_single with equal splits)
 *           from `send` to `recv` of tfqb_sharded_buffers, then calls
 *           tfqb_sharded_run_stage, which adopts `recv` as the shard
 *   kind 2  expectation    : tfqb_sharded_run_stage accumulates this rank's
 *           per-term partial sums
 * After the last stage: tfqb_sharded_partials -> sum the n_terms doubles over
 * ranks (all-reduce) -> tfqb_sharded_finish gives float[n_ops]. ------------- */
int tfqb_sharded_prepare(tfqb_context* ctx, const tfqb_circuit_inputs* in,
                         tfqb_strings pauli_sums, int n_ops, int world,
                         int rank, tfqb_job** job, int* n_stages,
                         int* n_terms);
int tfqb_sharded_stage_kind(tfqb_job* job, int stage);
int tfqb_sharded_run_stage(tfqb_job* job, int stage);
int tfqb_sharded_buffers(tfqb_job* job, void** send, void** recv,
                         size_t* bytes);
int tfqb_sharded_partials(tfqb_job* job, double* per_term);
int tfqb_sharded_finish(tfqb_job* job, const double* per_term_total,
                        float* expectations);

/* The same job with the exchange done INSIDE the library, through peer
 * memory: every rank maps the shard buffers of all ranks (CUDA IPC between
 * processes, raw pointers inside one process; NVLink / NVSwitch P2P) and each
 * exchange is this rank's kernels loading its incoming chunks straight from
 * the peers' HBM, ordered by epoch flags in device memory -- no host
 * collective, no NCCL, no host synchronisation between stages.
 *   1. tfqb_sharded_prepare on every rank;
 *   2. tfqb_sharded_export -> TFQB_PEER_HANDLE_BYTES bytes; the host
 *      all-gathers them over the ranks (any transport: torch.distributed,
 *      MPI, a TF collective) in rank order;
 *   3. tfqb_sharded_connect with the world x TFQB_PEER_HANDLE_BYTES table;
 *   4. tfqb_sharded_enqueue (asynchronous: all stages, exchanges and the
 *      cross-rank sum of the partials) any number of times, each followed by
 *      tfqb_sharded_result -> float[n_ops], identical on every rank.
 * A rank that does not show up makes its peers fail with TFQB_INTERNAL after
 * TFQB_PEER_TIMEOUT_S seconds (default 30) instead of hanging. */
#define TFQB_PEER_HANDLE_BYTES 256
int tfqb_sharded_export(tfqb_job* job, unsigned char* handle);
int tfqb_sharded_connect(tfqb_job* job, const unsigned char* handles, int world);
int tfqb_sharded_enqueue(tfqb_job* job);
int tfqb_sharded_result(tfqb_job* job, float* expectations);
/* Sampling from the sharded state (the single-circuit path of
 * TfqSimulateSamples, tfq_simulate_samples_op.cc:122-180, for a state too
 * large for one GPU; prepare the job with n_ops = 0).  After
 * tfqb_sharded_enqueue: shard norms are exchanged through the flag blocks,
 * every shot is assigned to ONE rank by the coarse CDF over the ranks and
 * sampled from that rank's shard.  samples: int8[num_samples, n_qubits] (same
 * bit order as tfqb_simulate_samples); only this rank's rows are written and
 * owned[s] = 1 for them -- the host merges the ranks.  `uniforms` (optional):
 * double[num_samples], identical on every rank; else Philox(seed) as in the
 * unsharded op.  The CDF runs over the PHYSICAL amplitude order, so for given
 * uniforms the bitstrings differ from the unsharded op's, not their
 * distribution. */
int tfqb_sharded_sample(tfqb_job* job, int num_samples, uint64_t seed,
                        const double* uniforms, int8_t* samples, int32_t* owned);
typedef struct {
  int n_qubits, n_local;
  int exchanges;                      /* in the last enqueue */
  int fused_exchanges;                /* of those: done by the load phase of the
                                         next gate pass (pull_ms then holds that
                                         pass, gate arithmetic included) */
  int gate_passes, expectation_passes;
  double shard_bytes;
  double bytes_received_per_exchange; /* over NVLink, per GPU */
  double wait_ms;                     /* waiting for the peers' shards */
  double pull_ms;                     /* the peer-memory loads themselves */
} tfqb_exchange_stats;
int tfqb_sharded_stats(tfqb_job* job, tfqb_exchange_stats* out);

/* ---- instrumentation ------------------------------------------------- */
int tfqb_sync(tfqb_context* ctx);
/* The context's CUDA stream as a cudaStream_t handle (for event timing). */
void* tfqb_stream(tfqb_context* ctx);

typedef struct {
  /* counters since the last tfqb_profile_reset */
  int64_t kernel_launches;        /* all kernels of this library */
  int64_t gate_pass_launches;     /* forward (Q1) pass launches */
  int64_t adjoint_pass_launches;  /* reverse-sweep pass launches */
  double gate_pass_ms;            /* CUDA-event time inside those launches */
  double adjoint_pass_ms;         /*   (only when event timing is enabled) */
  double gate_pass_bytes;         /* algorithmic bytes: 16 * 2^n * rows each */
  double adjoint_pass_bytes;      /* 32 * 2^n * rows each */
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  int64_t expectation_launches;   /* K1 PauliSum expectation launches */
  double expectation_ms;
  double expectation_bytes;       /* 8 * 2^n * rows per (state, sum) */
  int64_t jit_kernels;            /* pass kernels specialised at run time (csrc/jit.h) */
  int64_t jit_pass_launches;      /* pass launches that used a specialised kernel */
} tfqb_profile;
/* enable != 0 brackets every pass launch with CUDA events on the stream. */
int tfqb_profile_enable(tfqb_context* ctx, int enable);
int tfqb_profile_reset(tfqb_context* ctx);
int tfqb_profile_read(tfqb_context* ctx, tfqb_profile* out);

/* ---- host-only helpers (no GPU needed; used by the CPU test-suite) ---- */
/* Gate matrix exactly as the device builder computes it. kind = GateKind of
 * csrc/program.h; out = 2*dim*dim floats. grad_param < 0: the gate; else the
 * finite-difference gradient gate w.r.t. params[grad_param]. */
int tfqb_host_gate_matrix(int kind, const float* params, int n_params,
                          int grad_param, float* out);
/* Parse + lower one program and describe the forward/adjoint plan as JSON
 * text (caller frees with tfqb_free_string). */
int tfqb_host_describe_plan(const char* program, size_t program_size,
                            tfqb_strings symbol_names, int n_symbols,
                            int adjoint, char** json_out);
/* Parse + lower one PauliSum against a program; JSON of the mask form. */
int tfqb_host_describe_pauli_sum(const char* program, size_t program_size,
                                 const char* pauli_sum, size_t pauli_sum_size,
                                 char** json_out);
/* Stage list of the sharded-state plan for `world` ranks as JSON. */
int tfqb_host_describe_sharded(const char* program, size_t program_size,
                               tfqb_strings symbol_names, int n_symbols,
                               tfqb_strings pauli_sums, int n_ops, int world,
                               char** json_out);
/* CUDA C++ source of the run-time specialised kernel (csrc/jit.h) for pass
 * `pass` of the program's forward (adjoint = 0; adjoint = 2: the variant for
 * jobs that cannot see a global phase) or adjoint (1) plan; an empty string
 * when that pass is not specialisable. Host only: no GPU needed. */
int tfqb_host_jit_source(const char* program, size_t program_size,
                         tfqb_strings symbol_names, int n_symbols,
                         int adjoint, int pass, char** source_out);
/* Same for pass `pass` of the tile-based expectation plan of one row's
 * PauliSums (n_ops strings) against the program; pass + 1000 selects the
 * operator-accumulation kernel (lambda = sum g O psi) of that pass. */
int tfqb_host_jit_expect_source(const char* program, size_t program_size,
                                tfqb_strings pauli_sums, int n_ops, int pass,
                                char** source_out);
void tfqb_free_string(char* s);

/* ---- parameter-shift helper ops (host only) -------
 * Rewrites of serialized tfq.proto.Program strings; no device work and no
 * context.  Output strings are what TFQ's serializer would write
 * (language.gate_set "tfq_gate_set", MOMENT_BY_MOMENT circuit, every arg
 * value kind the serializer writes); ArgFunction / Schedule are not carried
 * over. */
typedef struct {
  char** data;        /* count strings, owned by the library */
  size_t* size;
  size_t count;
} tfqb_string_list;
void tfqb_free_string_list(tfqb_string_list* list);
void tfqb_free_floats(float* p);
/* TfqPsDecompose (core/ops/tfq_ps_decompose_op.cc:43-328): parameterised ISP /
 * PXP / FSIM / PISP operations become XXP / YYP / ZP / XP / CZP operations in
 * extra moments.  out: string[batch]. */
int tfqb_ps_decompose(tfqb_strings programs, int batch, tfqb_string_list* out);
/* TfqPsSymbolReplace (core/ops/tfq_ps_symbol_replace_op.cc:41-209): for every
 * (program i, symbol j) one copy of the program per occurrence of symbols[j],
 * that occurrence renamed to replacement_symbols[j].  out: string[batch,
 * n_symbols, *pad] row-major, padded with empty programs.  Error
 * "symbols.shape is not equal to replacement_symbols.shape". */
int tfqb_ps_symbol_replace(tfqb_strings programs, int batch, tfqb_strings symbols, int n_symbols,
                           tfqb_strings replacement_symbols, int n_replacements,
                           tfqb_string_list* out, int* pad);
/* TfqPsWeightsFromSymbols (core/ops/tfq_ps_weights_from_symbols_op.cc:43-171):
 * weights float[batch, n_symbols, *pad] (zero padded; free with
 * tfqb_free_floats): the exponent_scalar of every operation whose exponent is
 * that symbol.  Error "A circuit contains a sympy.Symbol not found in symbols!". */
int tfqb_ps_weights_from_symbols(tfqb_strings programs, int batch, tfqb_strings symbols,
                                 int n_symbols, float** weights, int* pad);

#ifdef __cplusplus
}
#endif
#endif /* TFQB_H_ */
