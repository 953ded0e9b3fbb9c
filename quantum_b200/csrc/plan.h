// Pass planner: turns a gate-template list into the device program the
// cache-blocked kernels interpret.
//
// Replaces `qsim::BasicGateFuser::FuseGates` (call sites
// circuit_parser_qsim.cc:857-859, adj_util.cc:155-172) and the reference's
// one-sweep-per-fused-gate loop (tfq_simulate_expectation_op.cc:164-166): the
// unit of HBM traffic here is a *pass* — one read + one write of every
// amplitude — during which a tile of 2^t amplitudes sits in shared memory and
// many gates are applied to it; inside a pass, gates are grouped into
// *rounds*, each of which holds 2^R amplitudes per thread in registers.
//
//  pass  : tile = t amplitude-index bits (always including the low L bits, so
//          global accesses are >= 2^L*8-byte contiguous runs)
//  round : R tile-local bit positions live in registers; dense 1q/2q gates
//          need their targets among them; diagonal gates and controls may
//          touch any bit (they become predicates / phase selects on the index)
#pragma once
#include <cstdint>
#include <vector>

#include "program.h"

namespace tfqb {

constexpr int kTileMax = 12;   // 2^12 amplitudes = 32 KiB of smem per state
constexpr int kLowBits = 4;    // low bits always in the tile (128 B runs)
constexpr int kRegBits = 4;    // forward kernel: 16 amplitudes per thread
constexpr int kRegBitsAdj = 3; // adjoint kernel: 8 + 8 amplitudes per thread
constexpr int kMinStateBits = 5;  // states are padded to >= 2^5 amplitudes

enum OpKind : int {
  kOpG1 = 0,    // dense 2x2 on register bit b0           (8 floats)
  kOpG2 = 1,    // dense 4x4 on register bits (b0=msb,b1) (32 floats)
  kOpD = 2,     // diagonal on 1..2 arbitrary bits        (8 floats, d[sel])
  kOpGrad1 = 3, // adjoint: 2 Re<lam| D |psi>, D dense 2x2
  kOpGrad2 = 4, // adjoint: D dense 4x4
  kOpGradD = 5, // adjoint: D diagonal
  // fused adjoint step of one parameterised gate with one symbol: psi <- G'psi,
  // grad += 2 Re<lam| D |psi>, lam <- G'lam (two matrices: dagger, then D)
  kOpAdj1 = 6,
  kOpAdj2 = 7,
  kOpAdjD = 8,
};

enum OpTarget : int { kTgtPsi = 1, kTgtLam = 2, kTgtBoth = 3 };

// Pre-decoded dispatch code of an op (one switch in the kernel). Register
// indices are folded into the code so every case is straight-line code.
enum OpCode : int {
  kCodeG1 = 0,       // +J            (4)   dense 2x2, no controls
  kCodeG2 = 4,       // +pair(JH,JL)  (6)   dense 4x4, no controls
  kCodeD0 = 10,      //                     diagonal, thread-constant selector
  kCodeD1 = 11,      // +J            (4)   diagonal, one register bit
  kCodeD2 = 15,      // +pair         (6)   diagonal, two register bits
  kCodeSlow = 21,    //                     anything with controls
  kCodeGrad1 = 22,   // +J            (4)
  kCodeGrad2 = 26,   // +pair         (6)
  kCodeGradD0 = 32,
  kCodeGradD1 = 33,  // +J            (4)
  kCodeGradD2 = 37,  // +pair         (6)
  kCodeAdj1 = 43,    // +J            (4)   fused adjoint step, dense 2x2
  kCodeAdj2 = 47,    // +pair         (6)   dense 4x4
  kCodeAdjD0 = 53,   //                     diagonal, thread-constant selector
  kCodeAdjD1 = 54,   // +J            (4)
  kCodeAdjD2 = 58,   // +pair         (6)
  // literal diagonal gates whose entries are all +-1 (CZ, Z, ZZ at exponent 1):
  // pure sign flips, no FP32-pipe work; ident_mask holds the entries to negate
  kCodeS0 = 64,
  kCodeS1 = 65,      // +J            (4)
  kCodeS2 = 69,      // +pair         (6)
  // (75 was the tensor-core block op of rounds 1-2: removed)
  // macro-ops formed after scheduling (plan.cc merge_macro_ops): one dispatch
  // for several commuting ops of a round
  // G1 on several distinct register bits: ident_mask = mask of the bits, the
  // 2x2 matrices are consecutive in descending bit order
  kCodeG1Run = 76,
  // every thread-constant sign op of the round as one GF(2) polynomial of the
  // group's base index g: parity = (ident_mask & 1) ^ |g & crest_mask|
  //   ^ |g & (g >> dpos0) & crest_bits| ^ |g & (g >> dpos1) & pad_|
  kCodeS0Run = 77,
};

// One interpreted op (device-visible POD, 80 bytes, 16-byte aligned rows).
struct OpRec {
  // --- hot word 0
  int32_t code;          // OpCode
  int32_t mat_off;       // float offset inside the row's matrix block
  int32_t target;        // OpTarget (adjoint kernel only)
  int32_t grad_slot;     // gradient ops: output slot, else -1
  // --- hot word 1 (diagonal ops): each of the two selector bits is either a
  // register bit (dreg >= 0) or a bit of the global base index (dpos);
  // dpos1 < 0 => 1 qubit
  int32_t dreg0, dreg1;
  int32_t dpos0, dpos1;
  // --- word 2
  uint32_t ident_mask;   // diagonal: bit s set => entry s is exactly 1
                         // (sign ops: bit s set => entry s is -1)
  int32_t b0, b1;        // register-bit indices for dense ops, else -1
  int32_t kind;          // OpKind
  // --- controls (slow path)
  uint32_t creg_mask;    // controls that are register bits (mask over R bits)
  uint32_t creg_bits;
  uint64_t crest_mask;   // controls elsewhere: predicate on the group's
  uint64_t crest_bits;   //   global base index
  // kCodeS0Run: second quadratic mask.  kCodeG1 / kCodeG1Run: 4 bits per
  // register bit j at [4j, 4j+3]: 1 = the gate is D R (real rotation, then
  // diagonal), 2 = R D, 3 = R alone, 4 = X^t alone, 0 = general (plan.cc
  // phased_real_flag); kCodeAdj1: 3 / 4 = a pure Y / X rotation times a phase
  uint64_t pad_;
};
static_assert(sizeof(OpRec) == 80, "OpRec layout");

struct RoundRec {
  int32_t pos[4];        // tile-local positions held in registers, ascending
  int32_t op_begin, op_end;
};

struct PassRec {
  int32_t tile_bits;                 // t
  int32_t low_bits;                  // L (local bits 0..L-1 == global 0..L-1)
  int32_t tile_pos[kTileMax];        // global position of tile-local bit i
  int32_t comp_pos[64];              // global positions NOT in the tile
  int32_t n_comp;
  int32_t round_begin, round_end;
  int32_t mat_begin, mat_len;        // floats: slice of the row matrix block
  // pass 0 of a forward plan only: the state starts as a PRODUCT state
  // prod_b u_b[i_b]; the 2 complex entries of u_b sit at float offset
  // init_off + 4*b of this pass's matrix slice, for b < init_bits.
  int32_t init_bits;                 // 0: no product init
  int32_t init_off;
};

// Recipe for one op matrix, evaluated per row by the builder kernel.
enum MatMode : int {
  kMatGate = 0,     // the gate itself
  kMatDagger = 1,   // conjugate transpose
  kMatGrad = 2,     // finite-difference gradient gate w.r.t. p[shift_idx]
};

// One gate of a fused op: the op's matrix is the ordered product of its
// factors, each embedded on the op's qubits (replaces the on-the-fly matrix
// product of qsim's ApplyFusedGate, call site
// tfq_simulate_expectation_op.cc:164-166).
struct FactorRec {
  int32_t gate_kind;
  int32_t nparams;
  int32_t slot;          // 0: 1q gate on matrix qubit 0 (msb); 1: on qubit 1;
                         // 2: 2q gate in the op's qubit order; 3: reversed
  int32_t aux_sym;       // kCH (non-unitary noise channel): parameter column of
                         // the measured |1> population, else -1
  int32_t sym[5];        // symbol column or -1
  float value[5];        // literal when sym < 0
};

struct MatRec {
  int32_t mode;          // MatMode
  int32_t shift_idx;     // kMatGrad: index into p[] of the (single) factor
  int32_t layout;        // 0: dense 2x2, 1: dense 4x4, 2: diag(2), 3: diag(4),
                         // 4: first column of a 2x2 (product-state init)
  int32_t swap;          // dense 4x4: exchange the two qubits (b0<->b1)
  int32_t out_off;       // float offset inside the row matrix block
  int32_t factor_begin, factor_end;
  int32_t pad_;
};

struct GradSlot {
  int32_t symbol_col;    // output column in grads[B,P]
};

struct DevicePlan {
  int n = 0;             // circuit qubits
  int n_alloc = 0;       // state bits actually stored (>= kMinStateBits)
  int reg_bits = kRegBits;
  std::vector<PassRec> passes;
  std::vector<RoundRec> rounds;
  std::vector<OpRec> ops;
  std::vector<MatRec> mats;
  std::vector<FactorRec> factors;
  std::vector<GradSlot> grad_slots;
  int mat_floats = 0;    // floats per row in the matrix block
  bool row_dependent = false;  // any matrix depends on a symbol
  bool product_init = false;   // pass 0 synthesises a product state
  int macro_merged = 0;  // dispatches saved by merge_macro_ops
  // segment of a sharded state that follows a qubit swap: its first pass
  // loads the tiles from the peers' shards (kernels.cuh init_mode 3)
  bool after_exchange = false;
};

// ---- PauliSum expectation plan (K1, util_qsim.h:142-188) ------------------
// Terms are evaluated from tiles staged in shared memory instead of one
// copy/apply/inner-product sweep per term:
//  * Z-type terms (x == 0): one Walsh-Hadamard transform of the tile's
//    probabilities gives every Z-string character over the tile bits; a term
//    is then one table lookup per tile (sign of the non-tile bits from the
//    tile base).  All of them ride on pass 0.
//  * X/Y-type terms (x != 0, |x| <= kRegBits): the x bits must be register
//    bits of a round; the partner amplitude i^x is then in the same thread.
//  * anything else falls back to the generic global-gather kernel.
struct ExpXOp {          // one X/Y-type term inside a round
  uint32_t xreg;         // x mask over the round's register bits (1..15)
  uint32_t sign16;       // bit e: parity(e & z restricted to register bits)
  uint64_t zrest;        // z bits that are not register bits (global positions)
  int32_t use_im;        // odd phase: Im(conj(a_i) a_k), else Re
  int32_t negate;        // overall -1 (phase 1 or 2)
  int32_t term;          // index into the per-term partial array
  // expectation kernel dispatch: 0 = generic (runtime sign16), else
  // 1 + (use_im * 5 + zs) * 15 + (xreg - 1) with zs = 0 (no register z bit)
  // or 1 + j (the only register z bit is j): signs become compile-time
  int32_t code;
};

struct ExpZTerm {        // Z-type term (pass 0)
  uint32_t ztile;        // z mask in tile-local bit positions (table index)
  int32_t negate;
  uint64_t zrest;        // z bits outside the tile (global positions)
  int32_t term;
  int32_t pad_;
};

struct ExpectationPlan {
  int n_alloc = 0;
  std::vector<PassRec> passes;     // tile layout + [round_begin, round_end)
  std::vector<RoundRec> rounds;    // register bits + [op_begin, op_end) in xops
  std::vector<ExpXOp> xops;
  std::vector<ExpZTerm> zterms;    // all evaluated in pass 0
  std::vector<int32_t> generic_terms;  // indices for the fallback kernel
  // sharded states only: X/Y-type terms whose x mask touches a non-local
  // (rank) bit in the current qubit layout; they need a qubit swap first
  std::vector<int32_t> deferred_terms;
};

struct TermMask {        // a PauliTerm in mask form (program.h PauliTermT)
  uint64_t x, z;
  int phase;
  bool identity;
};

// `identity_as_z`: identity terms become Z-type terms with z = 0 (operator
// accumulation needs them; the expectation adds their coefficient on the host
// side of the combine kernel).
// `n_local` < n: the state is sharded, only index bits < n_local are
// addressable on this rank (the others are rank bits: signs only).
ExpectationPlan PlanExpectation(int n, const std::vector<TermMask>& terms,
                                bool identity_as_z = false,
                                int tile_max = kTileMax, int low_bits = kLowBits,
                                int n_local = -1);

// ---- one state sharded over 2^g ranks by its top index bits ----
// The circuit is cut into segments whose dense gates only touch local bits;
// between segments a global<->local qubit swap exchanges the g rank bits with
// the top g local bits (one all-to-all of the whole shard).  Which logical
// qubits become global is chosen by farthest next dense use; they are moved
// to the top local positions with SWAP gates that ride in the segment.
struct ShardedStage {
  int32_t kind;    // 0: gate segment, 1: exchange, 2: expectation
  int32_t index;   // into gate_plans / exp_plans
};
struct ShardedPlan {
  int n = 0, n_local = 0, g = 0;
  std::vector<ShardedStage> stages;
  std::vector<DevicePlan> gate_plans;
  std::vector<ExpectationPlan> exp_plans;
  std::vector<int> final_phys;     // physical position of each logical bit
  int n_exchanges = 0;
};
ShardedPlan PlanSharded(const CircuitT& c, int g,
                        const std::vector<TermMask>& terms);

// Forward plan: applies the circuit. With `fuse`, runs of 1-qubit gates on a
// qubit collapse into one 2x2 and 1-qubit gates are absorbed into adjacent
// dense 2-qubit gates (the per-row products are evaluated on the device).
// `from_zero_state` = false: the plan continues a state that is already in
// memory (a later segment of a noisy trajectory), so the leading 1-qubit
// gates are NOT folded into a synthesised product state.
DevicePlan PlanForward(const CircuitT& c, int tile_max = kTileMax,
                       int low_bits = kLowBits, bool fuse = true,
                       bool from_zero_state = true);
// Reverse plan for the adjoint sweep (tfq_adj_grad_op.cc:225-276): gates in
// reverse, daggered, on psi and lambda, with gradient ops at parameterised
// gates.
DevicePlan PlanAdjoint(const CircuitT& c, int tile_max = kTileMax,
                       int low_bits = kLowBits, int reg_bits = kRegBitsAdj);
// Plan for a list of 1-qubit basis rotations (sampled expectation).
DevicePlan PlanRotations(int n, const std::vector<std::pair<int, int>>& rot,
                         int tile_max = kTileMax, int low_bits = kLowBits);

}  // namespace tfqb
