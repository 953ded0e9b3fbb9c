// See plan.h.
#include "plan.h"

#include <cstdlib>

#include <algorithm>
#include <cassert>
#include <cmath>

#include "gates.cuh"

namespace tfqb {
namespace {

constexpr int kMaxPassMatFloats = 3072;  // 24 KiB of smem once expanded to float4

struct PFactor {
  int kind = 0;
  int nparams = 0;
  ParamRef p[5];
  int aux_sym = -1;      // see FactorRec::aux_sym
  int slot = 0;          // see FactorRec::slot
  bool diagonal = false;
  uint32_t ident = 0;    // diagonal entries that are exactly 1 for every row
};

struct PItem {
  int gate = -1;         // index into the circuit's gate list (or -1)
  int opkind = kOpG1;
  int target = kTgtBoth;
  int mode = kMatGate;
  int shift_idx = -1;
  int grad_slot = -1;
  int nt = 1;
  int t[2] = {0, 0};     // global target bits, t[0] = matrix msb
  uint64_t cmask = 0, cbits = 0;
  bool dense = true;     // targets must sit in registers
  uint64_t qmask = 0;    // every bit the item touches (dependencies)
  int mat_floats = 8;
  bool fused_adj = false;  // adjoint triple in one op: matrices [dagger][grad]
  bool sign_only = false;  // literal diagonal with +-1 entries
  uint32_t sign_mask = 0;  // entries equal to -1
  // matrix recipe: ordered product of factors (first applied first)
  std::vector<PFactor> factors;
  void retype() {        // after the factor list changed
    bool all_diag = true;
    for (const PFactor& f : factors) all_diag = all_diag && f.diagonal;
    dense = !all_diag;
    if (all_diag) {
      opkind = kOpD;
      mat_floats = 8;
    } else {
      opkind = nt == 1 ? kOpG1 : kOpG2;
      mat_floats = nt == 1 ? 8 : 32;
    }
  }
  uint32_t ident_mask() const {
    if (dense || mode == kMatGrad) return 0;
    uint32_t m = nt == 1 ? 0x3u : 0xFu;
    for (const PFactor& f : factors) m &= f.ident;
    return m;
  }
};

struct Group {
  std::vector<int> pos;    // sorted positions
  std::vector<int> items;  // indices into the item vector, execution order
};

inline int popc(uint64_t v) { return __builtin_popcountll(v); }

// Greedy list scheduling: walk the not-yet-done items in order, keep a
// position set S (|S| <= cap); an item is taken if nothing earlier on its
// qubits was skipped and its dense targets fit in S.
std::vector<Group> schedule(const std::vector<PItem>& items,
                            const std::vector<int>& subset, int cap,
                            uint64_t mandatory, uint64_t universe,
                            uint64_t dep_universe, int mat_budget,
                            uint64_t (*dense_mask)(const PItem&, const void*),
                            const void* ctx) {
  std::vector<Group> out;
  std::vector<char> done(subset.size(), 0);
  size_t remaining = subset.size();
  while (remaining) {
    uint64_t S = mandatory, blocked = 0;
    int mat_used = 0;
    Group g;
    for (size_t k = 0; k < subset.size(); ++k) {
      if (done[k]) continue;
      const PItem& it = items[subset[k]];
      if (it.qmask & blocked) { blocked |= it.qmask; continue; }
      const uint64_t need = it.dense ? dense_mask(it, ctx) : 0;
      const int it_floats = it.mat_floats * (it.fused_adj ? 2 : 1);
      if (popc(S | need) <= cap && mat_used + it_floats <= mat_budget) {
        S |= need;
        mat_used += it_floats;
        g.items.push_back(subset[k]);
        done[k] = 1;
        --remaining;
      } else {
        blocked |= it.qmask;
        if (mat_used + it_floats > mat_budget) blocked = ~0ull;
      }
      if ((blocked & dep_universe) == dep_universe) break;
    }
    assert(!g.items.empty());
    // pad S with the lowest free positions of the universe
    for (int b = 0; b < 64 && popc(S) < cap; ++b)
      if (((universe >> b) & 1) && !((S >> b) & 1)) S |= 1ull << b;
    for (int b = 0; b < 64; ++b)
      if ((S >> b) & 1) g.pos.push_back(b);
    out.push_back(std::move(g));
  }
  return out;
}

uint64_t dense_global(const PItem& it, const void*) {
  uint64_t m = 1ull << it.t[0];
  if (it.nt == 2) m |= 1ull << it.t[1];
  return m;
}

struct LocalCtx { const int* local_of; };  // global bit -> tile-local bit

uint64_t dense_local(const PItem& it, const void* c) {
  const int* lo = static_cast<const LocalCtx*>(c)->local_of;
  uint64_t m = 1ull << lo[it.t[0]];
  if (it.nt == 2) m |= 1ull << lo[it.t[1]];
  return m;
}

// Entries of a diagonal gate that are exactly (1, 0) whatever the exponent:
// with a literal global_shift of 0 the eigenphase e^{i*pi*t*0} is cs(0).
uint32_t identity_entries(const GateT& g) {
  if (!g.is_diagonal() || g.is_identity()) return 0;
  if (g.p[2].sym >= 0 || g.p[2].value != 0.f) return 0;
  switch (g.kind) {
    case kZP: return 0x1;    // diag(1, w)
    case kCZP: return 0x7;   // diag(1, 1, 1, w)
    case kZZP: return 0x9;   // diag(1, w, w, 1)
    default: return 0;
  }
}

PFactor factor_from_gate(const GateT& g, int slot) {
  PFactor f;
  f.kind = g.kind;
  f.nparams = g.nparams;
  for (int k = 0; k < 5; ++k) f.p[k] = g.p[k];
  f.aux_sym = g.aux_sym;
  f.slot = slot;
  f.diagonal = g.is_diagonal();
  f.ident = identity_entries(g);
  return f;
}

// A diagonal item whose parameters are all literals and whose entries are all
// +-1 (to float32 round-off, e.g. CZ = diag(1,1,1,-1-8.7e-8i)) becomes a sign
// flip: no matrix, no FP32-pipe work.  Returns false when it is the identity.
bool classify_sign(PItem* it) {
  it->sign_only = false;
  if (it->dense || it->mode == kMatGrad || it->fused_adj || it->cmask) return true;
  const int dim = it->nt == 1 ? 2 : 4;
  cf d[4] = {mk(1.f, 0.f), mk(1.f, 0.f), mk(1.f, 0.f), mk(1.f, 0.f)};
  for (const PFactor& f : it->factors) {
    float p[5];
    for (int k = 0; k < 5; ++k) {
      if (k < f.nparams && f.p[k].sym >= 0) return true;   // row dependent
      p[k] = k < f.nparams ? f.p[k].value : 0.f;
    }
    cf m[16];
    gate_matrix(f.kind, p, -1, 0.f, m);
    for (int i = 0; i < dim; ++i) {
      // diagonal items never mix 1q factors into 2q ones: same dimension
      const cf e = m[i * dim + i];
      d[i] = mk(d[i].re * e.re - d[i].im * e.im, d[i].re * e.im + d[i].im * e.re);
    }
  }
  uint32_t mask = 0;
  for (int i = 0; i < dim; ++i) {
    if (std::fabs(d[i].im) > 4e-7f) return true;
    if (std::fabs(d[i].re - 1.f) <= 4e-7f) continue;
    if (std::fabs(d[i].re + 1.f) <= 4e-7f) { mask |= 1u << i; continue; }
    return true;
  }
  it->sign_only = true;
  it->sign_mask = mask;
  it->mat_floats = 0;
  return mask != 0;
}

PItem item_from_gate(const GateT& g, int gate_index, int mode) {
  PItem it;
  it.gate = gate_index;
  it.mode = mode;
  it.nt = g.nq;
  it.t[0] = g.bit[0];
  it.t[1] = g.nq == 2 ? g.bit[1] : 0;
  it.cmask = g.cmask;
  it.cbits = g.cbits;
  it.qmask = g.target_mask() | g.cmask;
  it.factors.push_back(factor_from_gate(g, g.nq == 2 ? 2 : 0));
  it.retype();
  return it;
}

// Greedy gate fusion for forward plans (the role qsim::BasicGateFuser plays
// at circuit_parser_qsim.cc:857-859; the grouping rule here is our own):
//  * a 1-qubit gate joins the latest item on its qubit when that item is an
//    uncontrolled 1-qubit item or an uncontrolled dense 2-qubit item;
//  * a dense 2-qubit gate absorbs the latest 1-qubit items on its qubits.
// Diagonal 2-qubit gates and controlled gates stay alone (they are cheap /
// predicated), and act as barriers on the qubits they touch.
std::vector<PItem> fuse_forward(const CircuitT& c) {
  std::vector<PItem> out;
  std::vector<char> dead;
  int last[64];
  for (int b = 0; b < 64; ++b) last[b] = -1;
  auto touch = [&](uint64_t mask, int idx) {
    for (int b = 0; b < 64; ++b)
      if ((mask >> b) & 1) last[b] = idx;
  };
  for (size_t gi = 0; gi < c.gates.size(); ++gi) {
    const GateT& g = c.gates[gi];
    if (g.is_identity()) continue;
    if (g.cmask) {
      out.push_back(item_from_gate(g, int(gi), kMatGate));
      dead.push_back(0);
      touch(g.target_mask() | g.cmask, int(out.size()) - 1);
      continue;
    }
    if (g.nq == 1) {
      const int q = g.bit[0];
      const int L = last[q];
      if (L >= 0 && !dead[L] && out[L].cmask == 0) {
        PItem& F = out[L];
        if (F.nt == 1) {
          F.factors.push_back(factor_from_gate(g, 0));
          F.retype();
          continue;
        }
        if (F.nt == 2 && F.dense) {
          F.factors.push_back(factor_from_gate(g, F.t[0] == q ? 0 : 1));
          continue;
        }
      }
      out.push_back(item_from_gate(g, int(gi), kMatGate));
      dead.push_back(0);
      last[q] = int(out.size()) - 1;
      continue;
    }
    PItem N = item_from_gate(g, int(gi), kMatGate);
    if (N.dense) {
      std::vector<PFactor> pre;
      for (int k = 0; k < 2; ++k) {
        const int q = g.bit[k];
        const int L = last[q];
        if (L >= 0 && !dead[L] && out[L].nt == 1 && out[L].cmask == 0) {
          for (PFactor f : out[L].factors) {
            f.slot = k;   // N.t[k] == q
            pre.push_back(f);
          }
          dead[L] = 1;
        }
      }
      N.factors.insert(N.factors.begin(), pre.begin(), pre.end());
    }
    out.push_back(std::move(N));
    dead.push_back(0);
    touch(g.target_mask(), int(out.size()) - 1);
  }
  std::vector<PItem> live;
  for (size_t i = 0; i < out.size(); ++i) {
    if (dead[i]) continue;
    if (!classify_sign(&out[i])) continue;   // literal identity diagonal
    live.push_back(std::move(out[i]));
  }
  return live;
}

// A dense 1-qubit op whose factors are real rotations (Y^t, any exponent: a
// real matrix times a phase) followed OR preceded by diagonal gates is
// U = D R (1: rows carry one phase each) or U = R D (2: columns).  Kernels that
// may drop a global phase (run-time specialised passes of the expectation,
// sampling and adjoint jobs, jit.cc) then apply it with 3 packed FMAs per
// amplitude instead of 4 (3: no diagonal factor at all, 2 instead of 4;
// 4: X^t alone, a phase times [[c, -i s], [-i s, c]], also 2).
// 0: no such structure.
int phased_real_flag(const std::vector<PFactor>& factors) {
  int lead = 0, real = 0, trail = 0, reflections = 0;
  {     // X^t alone: phase times [[c, -i s], [-i s, c]] (real diagonal,
        // imaginary off-diagonal): 4
    int nx = 0, other = 0;
    for (const PFactor& f : factors) {
      if (f.kind == kI) continue;
      if (f.kind == kXP) ++nx; else ++other;
    }
    if (nx && !other) return 4;
  }
  for (const PFactor& f : factors) {
    if (f.kind == kI) continue;
    // H at a literal exponent of 1 is a real matrix times e^{i pi shift} too
    const bool h1 = f.kind == kHP && f.nparams >= 2 && f.p[0].sym < 0 && f.p[1].sym < 0 &&
                    f.p[0].value * f.p[1].value == 1.f;
    if (f.kind == kYP || h1) {
      if (trail) return 0;        // D R D R ...: general
      ++real;
      if (h1) ++reflections;
    } else if (f.diagonal) {
      if (real) ++trail; else ++lead;
    } else {
      return 0;
    }
  }
  if (!real || (lead && trail)) return 0;
  // + 8: every real factor is a Y^t, so R is a proper rotation (no H): the
  // specialised kernels may apply it as three shears (pass_device.cuh `lift`)
  const int lift = reflections == 0 ? 8 : 0;
  if (!lead && !trail) return 3 | lift;     // R times a phase: nothing but the rotation
  return (lead ? 2 : 1) | lift;
}

// Macro-ops: fewer dispatches in the interpreted pass kernel.
//  * all thread-constant sign ops (kCodeS0) of a round commute with every
//    other op of the round (they touch no register bit): they are hoisted to
//    the front and merged into kCodeS0Run ops (at most two distinct bit
//    distances of quadratic terms per op);
//  * consecutive kCodeG1 ops on distinct, descending register bits with
//    adjacent matrices become one kCodeG1Run.
void merge_macro_ops(DevicePlan* plan, int op_begin) {
  std::vector<OpRec> ops(plan->ops.begin() + op_begin, plan->ops.end());
  plan->ops.resize(op_begin);
  std::vector<OpRec> s0, rest;
  for (const OpRec& op : ops)
    (op.code == kCodeS0 ? s0 : rest).push_back(op);
  if (s0.size() < 2) {
    s0.clear();
    rest = ops;
  }
  // ---- S0 runs (same target only)
  size_t i = 0;
  while (i < s0.size()) {
    OpRec m{};
    m.code = kCodeS0Run;
    m.kind = kOpD;
    m.target = s0[i].target;
    m.grad_slot = -1;
    m.b0 = m.b1 = m.dreg0 = m.dreg1 = -1;
    m.dpos0 = m.dpos1 = 0;
    m.mat_off = s0[i].mat_off;
    uint32_t c = 0;
    uint64_t lin = 0, q[2] = {0, 0};
    int d[2] = {0, 0};
    size_t j = i;
    for (; j < s0.size(); ++j) {
      const OpRec& op = s0[j];
      if (op.target != m.target) break;
      const uint32_t e = op.ident_mask;
      if (op.dpos1 < 0) {           // 1 qubit: entries (e0, e1)
        c ^= e & 1u;
        if (((e >> 1) ^ e) & 1u) lin ^= 1ull << op.dpos0;
        continue;
      }
      // 2 qubits: entry index = 2 * bit(dpos0) + bit(dpos1)
      const uint32_t e0 = e & 1u, e1 = (e >> 1) & 1u, e2 = (e >> 2) & 1u,
                     e3 = (e >> 3) & 1u;
      if (e0 ^ e1 ^ e2 ^ e3) {
        const int lo = std::min(op.dpos0, op.dpos1);
        const int dist = std::abs(op.dpos0 - op.dpos1);
        int k = -1;
        for (int u = 0; u < 2 && k < 0; ++u)
          if (d[u] == dist || d[u] == 0) k = u;
        if (k < 0) break;           // a third distance: start a new run
        d[k] = dist;
        q[k] ^= 1ull << lo;
      }
      c ^= e0;
      if (e0 ^ e2) lin ^= 1ull << op.dpos0;
      if (e0 ^ e1) lin ^= 1ull << op.dpos1;
    }
    m.ident_mask = c;
    m.crest_mask = lin;
    m.crest_bits = q[0];
    m.pad_ = q[1];
    m.dpos0 = d[0];
    m.dpos1 = d[1];
    plan->ops.push_back(m);
    i = j;
  }
  // ---- G1 runs
  for (size_t k = 0; k < rest.size();) {
    const OpRec& op = rest[k];
    if (!(op.code >= kCodeG1 && op.code < kCodeG1 + 4)) {
      plan->ops.push_back(op);
      ++k;
      continue;
    }
    size_t e = k + 1;
    uint32_t mask = 1u << (op.code - kCodeG1);
    while (e < rest.size() && rest[e].code >= kCodeG1 &&
           rest[e].code < rest[e - 1].code &&
           rest[e].mat_off == rest[e - 1].mat_off + 8 &&
           rest[e].target == op.target) {
      mask |= 1u << (rest[e].code - kCodeG1);
      ++e;
    }
    if (e - k >= 2) {
      OpRec m = op;
      m.code = kCodeG1Run;
      m.ident_mask = mask;
      for (size_t u = k + 1; u < e; ++u) m.pad_ |= rest[u].pad_;   // phased-real flags
      plan->ops.push_back(m);
    } else {
      plan->ops.push_back(op);
    }
    k = e;
  }
  plan->macro_merged += int(ops.size()) - (int(plan->ops.size()) - op_begin);
}

// `first_low_bits` > low_bits: pass 0 alone keeps that many low bits in its tile
// (longer contiguous runs for a pass that loads from peer memory).
DevicePlan build(const std::vector<PItem>& items, int n, int reg_bits,
                 int tile_max, int low_bits, int n_local = -1,
                 const std::vector<PItem>* init = nullptr,
                 int first_low_bits = 0) {
  static const bool macro_ops = [] {     // TFQB_MACRO_OPS=0 keeps one op per gate
    const char* v = getenv("TFQB_MACRO_OPS");
    return !(v && *v == '0');
  }();
  DevicePlan plan;
  plan.n = n;
  // sharded states: only bits < n_local are addressable on this rank
  plan.n_alloc = n_local >= 0 ? n_local : std::max(n, kMinStateBits);
  plan.reg_bits = reg_bits;
  const int na = plan.n_alloc;
  const int ntot = std::max(na, n);
  const int t = std::min(tile_max, na);
  const int L = std::min(low_bits, t);
  const uint64_t universe = na >= 64 ? ~0ull : ((1ull << na) - 1);

  std::vector<int> all(items.size());
  for (size_t i = 0; i < items.size(); ++i) all[i] = int(i);
  // When the whole state is one tile the low-bit constraint is moot.
  const uint64_t mandatory = (1ull << L) - 1;
  uint64_t dep_all = 0;
  for (const PItem& it : items) dep_all |= it.qmask;
  std::vector<Group> passes =
      schedule(items, all, t, mandatory, universe, dep_all, kMaxPassMatFloats,
               dense_global, nullptr);
  const int L0 = std::min(std::max(first_low_bits, L), t);
  if (L0 > L && !passes.empty()) {
    // pass 0 with the wider mandatory set, everything else as usual
    std::vector<Group> wide = schedule(items, all, t, (1ull << L0) - 1, universe, dep_all,
                                       kMaxPassMatFloats, dense_global, nullptr);
    if (!wide.empty() && !wide[0].items.empty()) {
      std::vector<char> taken(items.size(), 0);
      for (int idx : wide[0].items) taken[idx] = 1;
      std::vector<int> rest;
      for (size_t i = 0; i < items.size(); ++i)
        if (!taken[i]) rest.push_back(int(i));
      uint64_t dep_rest = 0;
      for (int idx : rest) dep_rest |= items[idx].qmask;
      std::vector<Group> tail = schedule(items, rest, t, mandatory, universe, dep_rest,
                                         kMaxPassMatFloats, dense_global, nullptr);
      passes.clear();
      passes.push_back(wide[0]);
      passes.insert(passes.end(), tail.begin(), tail.end());
    }
  }

  if (init && passes.empty()) {   // only 1-qubit gates: one pass writes the state
    Group g0;
    for (int b = 0; b < t; ++b) g0.pos.push_back(b);
    passes.push_back(g0);
  }
  for (const Group& pg : passes) {
    PassRec pr{};
    pr.tile_bits = t;
    pr.low_bits = (L0 > L && &pg == &passes[0]) ? L0 : L;
    int local_of[64];
    for (int b = 0; b < 64; ++b) local_of[b] = -1;
    for (int i = 0; i < t; ++i) {
      pr.tile_pos[i] = pg.pos[i];
      local_of[pg.pos[i]] = i;
    }
    pr.n_comp = 0;
    for (int b = 0; b < na; ++b)
      if (local_of[b] < 0) pr.comp_pos[pr.n_comp++] = b;
    pr.round_begin = int(plan.rounds.size());
    pr.mat_begin = plan.mat_floats;
    if (init && plan.passes.empty()) {
      // product-state init vectors: one first-column record per index bit
      plan.product_init = true;
      pr.init_bits = ntot;
      pr.init_off = 0;
      for (int b = 0; b < ntot; ++b) {
        MatRec mr{};
        mr.mode = kMatGate;
        mr.layout = 4;
        mr.out_off = plan.mat_floats;
        mr.factor_begin = int(plan.factors.size());
        const PItem* src = nullptr;
        for (const PItem& it : *init)
          if (it.t[0] == b) src = &it;
        if (src) {
          for (const PFactor& f : src->factors) {
            FactorRec fr{};
            fr.gate_kind = f.kind;
            fr.nparams = f.nparams;
            fr.slot = 0;
            fr.aux_sym = f.aux_sym;
            for (int k = 0; k < 5; ++k) {
              fr.sym[k] = f.p[k].sym;
              fr.value[k] = f.p[k].value;
              if (k < f.nparams && f.p[k].sym >= 0) plan.row_dependent = true;
            }
            plan.factors.push_back(fr);
          }
        } else {
          FactorRec fr{};
          fr.gate_kind = kI;
          fr.aux_sym = -1;
          for (int k = 0; k < 5; ++k) fr.sym[k] = -1;
          plan.factors.push_back(fr);
        }
        mr.factor_end = int(plan.factors.size());
        plan.mats.push_back(mr);
        plan.mat_floats += 4;
      }
    }

    LocalCtx ctx{local_of};
    const uint64_t tile_universe = (1ull << t) - 1;
    // items are in local coordinates for dependency purposes only through
    // qmask (global), which is fine: blocked/qmask comparisons stay global.
    uint64_t dep_pass = 0;
    for (int idx : pg.items) dep_pass |= items[idx].qmask;
    std::vector<Group> rounds =
        schedule(items, pg.items, std::min(reg_bits, t), 0, tile_universe,
                 dep_pass, 1 << 30, dense_local, &ctx);
    for (const Group& rg : rounds) {
      RoundRec rr{};
      int reg_of_local[64];
      for (int b = 0; b < 64; ++b) reg_of_local[b] = -1;
      const int R = int(rg.pos.size());
      for (int j = 0; j < 4; ++j) rr.pos[j] = j < R ? rg.pos[j] : -1;
      for (int j = 0; j < R; ++j) reg_of_local[rg.pos[j]] = j;
      rr.op_begin = int(plan.ops.size());
      for (int idx : rg.items) {
        const PItem& it = items[idx];
        OpRec op{};
        op.kind = it.opkind;
        op.target = it.target;
        op.b0 = op.b1 = -1;
        op.grad_slot = it.grad_slot;
        op.dreg0 = op.dreg1 = -1;
        op.dpos0 = op.dpos1 = -1;
        op.ident_mask = it.ident_mask();
        MatRec mr{};
        mr.mode = it.mode;
        mr.shift_idx = it.shift_idx;
        mr.factor_begin = int(plan.factors.size());
        for (const PFactor& f : it.factors) {
          if (it.sign_only) break;   // no matrix: nothing to evaluate
          FactorRec fr{};
          fr.gate_kind = f.kind;
          fr.nparams = f.nparams;
          fr.slot = f.slot;
          fr.aux_sym = f.aux_sym;
          for (int k = 0; k < 5; ++k) {
            fr.sym[k] = f.p[k].sym;
            fr.value[k] = f.p[k].value;
            if (k < f.nparams && f.p[k].sym >= 0) plan.row_dependent = true;
          }
          plan.factors.push_back(fr);
        }
        mr.factor_end = int(plan.factors.size());
        mr.out_off = plan.mat_floats;
        op.mat_off = plan.mat_floats - pr.mat_begin;
        auto reg_of = [&](int gbit) {
          const int l = local_of[gbit];
          return l < 0 ? -1 : reg_of_local[l];
        };
        if (it.dense) {
          op.b0 = reg_of(it.t[0]);
          assert(op.b0 >= 0);
          if (it.nt == 2) {
            op.b1 = reg_of(it.t[1]);
            assert(op.b1 >= 0 && op.b1 != op.b0);
            mr.layout = 1;
            if (op.b0 < op.b1) {  // canonical: b0 (matrix msb) > b1
              std::swap(op.b0, op.b1);
              mr.swap = 1;
            }
          } else {
            mr.layout = 0;
          }
        } else {
          mr.layout = it.nt == 2 ? 3 : 2;
          op.dreg0 = reg_of(it.t[0]);
          op.dpos0 = it.t[0];
          if (it.nt == 2) {
            op.dreg1 = reg_of(it.t[1]);
            op.dpos1 = it.t[1];
          }
        }
        for (int b = 0; b < ntot; ++b) {
          if (!((it.cmask >> b) & 1)) continue;
          const int r = reg_of(b);
          const uint64_t v = (it.cbits >> b) & 1;
          if (r >= 0) {
            op.creg_mask |= 1u << r;
            op.creg_bits |= uint32_t(v) << r;
          } else {
            op.crest_mask |= 1ull << b;
            op.crest_bits |= v << b;
          }
        }
        // canonical order for two-register-bit diagonals: selector msb on the
        // higher register (the builder exchanges entries 1 and 2)
        if (!it.dense && it.nt == 2 && op.dreg0 >= 0 && op.dreg1 >= 0 &&
            op.dreg0 < op.dreg1) {
          std::swap(op.dreg0, op.dreg1);
          std::swap(op.dpos0, op.dpos1);
          mr.swap = 1;
          op.ident_mask = (op.ident_mask & 9u) | ((op.ident_mask & 2u) << 1) |
                          ((op.ident_mask & 4u) >> 1);
        }
        {
          const bool ctrl = op.creg_mask != 0 || op.crest_mask != 0;
          const bool grad = it.mode == kMatGrad;
          auto pair = [](int hi, int lo) { return hi * (hi - 1) / 2 + lo; };
          if (it.sign_only && !ctrl) {
            uint32_t m = it.sign_mask;
            if (mr.swap && it.nt == 2)   // selector bits were exchanged
              m = (m & 9u) | ((m & 2u) << 1) | ((m & 4u) >> 1);
            op.ident_mask = m;
            const int nreg = (op.dreg0 >= 0) + (it.nt == 2 && op.dreg1 >= 0);
            if (nreg == 0) op.code = kCodeS0;
            else if (nreg == 2) op.code = kCodeS2 + pair(op.dreg0, op.dreg1);
            else op.code = kCodeS1 + (op.dreg0 >= 0 ? op.dreg0 : op.dreg1);
          } else if (it.fused_adj) {
            assert(!ctrl);
            if (it.dense) {
              op.kind = it.nt == 1 ? kOpAdj1 : kOpAdj2;
              op.code = it.nt == 1 ? kCodeAdj1 + op.b0
                                   : kCodeAdj2 + pair(op.b0, op.b1);
            } else {
              op.kind = kOpAdjD;
              // entries that are exactly 1 (and whose gradient entries are
              // exactly 0): only the specialised kernels use them, through
              // creg_bits (fused adjoint steps have no controls); the
              // interpreted kernel multiplies every entry
              op.creg_bits = op.ident_mask;
              op.ident_mask = 0;
              const int nreg = (op.dreg0 >= 0) + (it.nt == 2 && op.dreg1 >= 0);
              if (nreg == 0) op.code = kCodeAdjD0;
              else if (nreg == 2) op.code = kCodeAdjD2 + pair(op.dreg0, op.dreg1);
              else op.code = kCodeAdjD1 + (op.dreg0 >= 0 ? op.dreg0 : op.dreg1);
            }
          } else if (ctrl) {
            op.code = kCodeSlow;
          } else if (it.dense) {
            if (it.nt == 1) op.code = (grad ? kCodeGrad1 : kCodeG1) + op.b0;
            else op.code = (grad ? kCodeGrad2 : kCodeG2) + pair(op.b0, op.b1);
          } else {
            const int nreg = (op.dreg0 >= 0) + (it.nt == 2 && op.dreg1 >= 0);
            if (nreg == 0) op.code = grad ? kCodeGradD0 : kCodeD0;
            else if (nreg == 2)
              op.code = (grad ? kCodeGradD2 : kCodeD2) + pair(op.dreg0, op.dreg1);
            else
              op.code = (grad ? kCodeGradD1 : kCodeD1) +
                        (op.dreg0 >= 0 ? op.dreg0 : op.dreg1);
          }
        }
        if (op.code >= kCodeG1 && op.code < kCodeG1 + 4 && it.mode == kMatGate)
          op.pad_ = uint64_t(phased_real_flag(it.factors)) << (4 * op.b0);
        // fused adjoint step of a pure rotation (Y^t): the dagger is R' times a
        // phase as well; the specialised kernel moves that phase into the
        // gradient gate (pass_device.cuh adj1_real)
        if (op.code >= kCodeAdj1 && op.code < kCodeAdj1 + 4 &&
            (phased_real_flag(it.factors) & 7) >= 3)
          op.pad_ = uint64_t(phased_real_flag(it.factors)) << (4 * op.b0);
        plan.mat_floats += it.mat_floats;
        plan.ops.push_back(op);
        if (!it.sign_only) plan.mats.push_back(mr);
        if (it.fused_adj) {   // second matrix: the gradient gate, same layout
          MatRec gr = mr;
          gr.mode = kMatGrad;
          gr.out_off = plan.mat_floats;
          plan.mats.push_back(gr);
          plan.mat_floats += it.mat_floats;
        }
      }
      if (macro_ops) merge_macro_ops(&plan, rr.op_begin);
      rr.op_end = int(plan.ops.size());
      plan.rounds.push_back(rr);
    }
    pr.round_end = int(plan.rounds.size());
    pr.mat_len = plan.mat_floats - pr.mat_begin;
    plan.passes.push_back(pr);
  }
  return plan;
}

}  // namespace

// The circuit acts on |0...0>: an uncontrolled 1-qubit item that is the first
// item on its qubit commutes to the front, so the state after all of them is
// the product state prod_b (U_b|0>)[i_b].  They are removed from the item
// list and synthesised by pass 0 instead of being applied as gates.
std::vector<PItem> extract_product_init(std::vector<PItem>* items) {
  std::vector<PItem> init, rest;
  uint64_t touched = 0;
  for (PItem& it : *items) {
    const bool first = !(it.qmask & touched);
    touched |= it.qmask;
    if (first && it.nt == 1 && it.cmask == 0 && it.mode == kMatGate &&
        !it.fused_adj) {
      init.push_back(it);
    } else {
      rest.push_back(it);
    }
  }
  *items = std::move(rest);
  return init;
}

DevicePlan PlanForward(const CircuitT& c, int tile_max, int low_bits,
                       bool fuse, bool from_zero_state) {
  std::vector<PItem> items;
  if (fuse) {
    items = fuse_forward(c);
  } else {
    for (size_t i = 0; i < c.gates.size(); ++i) {
      if (c.gates[i].is_identity()) continue;
      PItem it = item_from_gate(c.gates[i], int(i), kMatGate);
      if (!classify_sign(&it)) continue;
      items.push_back(it);
    }
  }
  if (fuse && from_zero_state) {
    std::vector<PItem> init = extract_product_init(&items);
    if (!init.empty())
      return build(items, c.n, kRegBits, tile_max, low_bits, -1, &init);
  }
  return build(items, c.n, kRegBits, tile_max, low_bits, -1, nullptr);
}

DevicePlan PlanAdjoint(const CircuitT& c, int tile_max, int low_bits,
                       int reg_bits) {
  std::vector<PItem> items;
  std::vector<GradSlot> slots;
  for (int i = int(c.gates.size()) - 1; i >= 0; --i) {
    const GateT& g = c.gates[i];
    if (g.is_identity()) continue;   // identity: no sweep, zero gradient
    PItem dag = item_from_gate(g, i, kMatDagger);
    if (g.nsym == 0) {
      dag.target = kTgtBoth;
      if (!classify_sign(&dag)) continue;    // literal identity diagonal
      items.push_back(dag);
      continue;
    }
    if (g.nsym == 1 && g.cmask == 0) {
      // one op does psi <- G'psi, grad, lam <- G'lam
      dag.fused_adj = true;
      dag.target = kTgtBoth;
      dag.shift_idx = g.sym_param[0];
      dag.grad_slot = int(slots.size());
      slots.push_back(GradSlot{g.sym_col[0]});
      items.push_back(dag);
      continue;
    }
    dag.target = kTgtPsi;
    items.push_back(dag);
    for (int k = 0; k < g.nsym; ++k) {
      PItem gr = item_from_gate(g, i, kMatGrad);
      gr.shift_idx = g.sym_param[k];
      gr.opkind = gr.dense ? (g.nq == 1 ? kOpGrad1 : kOpGrad2) : kOpGradD;
      gr.target = kTgtBoth;
      gr.grad_slot = int(slots.size());
      slots.push_back(GradSlot{g.sym_col[k]});
      items.push_back(gr);
    }
    dag.target = kTgtLam;
    items.push_back(dag);
  }
  DevicePlan p = build(items, c.n, reg_bits, tile_max, low_bits);
  p.grad_slots = slots;
  return p;
}

DevicePlan PlanRotations(int n, const std::vector<std::pair<int, int>>& rot,
                         int tile_max, int low_bits) {
  std::vector<PItem> items;
  for (const auto& r : rot) {
    GateT g;
    g.nq = 1;
    g.bit[0] = r.first;
    g.nparams = 3;
    // X -> Y^-0.5 ; Y -> X^+0.5 (circuit_parser_qsim.cc:907-920)
    g.kind = r.second == 1 ? kYP : kXP;
    g.p[0].value = r.second == 1 ? -0.5f : 0.5f;
    g.p[1].value = 1.f;
    g.p[2].value = 0.f;
    items.push_back(item_from_gate(g, -1, kMatGate));
  }
  return build(items, n, kRegBits, tile_max, low_bits);
}

ExpectationPlan PlanExpectation(int n, const std::vector<TermMask>& terms,
                                bool identity_as_z, int tile_max, int low_bits,
                                int n_local) {
  ExpectationPlan plan;
  plan.n_alloc = n_local >= 0 ? n_local : std::max(n, kMinStateBits);
  const int na = plan.n_alloc;
  const int ntot = std::max(na, n);
  const int t = std::min(tile_max, na);
  const int L = std::min(low_bits, t);
  const int R = std::min(kRegBits, t);
  const uint64_t universe = na >= 64 ? ~0ull : ((1ull << na) - 1);

  // X/Y-type terms become scheduling items whose dense targets are the x
  // bits; they only read the state, so there are no ordering constraints.
  std::vector<PItem> items;
  std::vector<int> item_term;
  std::vector<int> zlist;
  for (size_t k = 0; k < terms.size(); ++k) {
    const TermMask& tm = terms[k];
    if (tm.identity && !identity_as_z) continue;
    if (tm.identity || tm.x == 0) { zlist.push_back(int(k)); continue; }
    if (tm.x & ~universe) { plan.deferred_terms.push_back(int(k)); continue; }
    if (__builtin_popcountll(tm.x) > R) { plan.generic_terms.push_back(int(k)); continue; }
    PItem it;
    it.dense = true;
    it.qmask = 0;
    it.mat_floats = 0;
    it.cmask = tm.x;       // reused below as "x mask" by dense_mask_x
    items.push_back(it);
    item_term.push_back(int(k));
  }
  auto dense_x = [](const PItem& it, const void*) -> uint64_t { return it.cmask; };
  std::vector<Group> passes;
  if (!items.empty()) {
    std::vector<int> all(items.size());
    for (size_t i = 0; i < items.size(); ++i) all[i] = int(i);
    passes = schedule(items, all, t, (1ull << L) - 1, universe, ~0ull, 1 << 30,
                      dense_x, nullptr);
  } else {
    Group g;           // a single pass over the low tile for the Z-type terms
    for (int b = 0; b < t; ++b) g.pos.push_back(b);
    if (!zlist.empty()) passes.push_back(g);
  }
  for (size_t pi = 0; pi < passes.size(); ++pi) {
    const Group& pg = passes[pi];
    PassRec pr{};
    pr.tile_bits = t;
    pr.low_bits = L;
    int local_of[64];
    for (int b = 0; b < 64; ++b) local_of[b] = -1;
    for (int i = 0; i < t; ++i) {
      pr.tile_pos[i] = pg.pos[i];
      local_of[pg.pos[i]] = i;
    }
    pr.n_comp = 0;
    for (int b = 0; b < na; ++b)
      if (local_of[b] < 0) pr.comp_pos[pr.n_comp++] = b;
    pr.round_begin = int(plan.rounds.size());
    if (pi == 0) {
      for (int k : zlist) {
        const TermMask& tm = terms[k];
        ExpZTerm zt{};
        zt.term = k;
        zt.negate = (!tm.identity && (tm.phase & 2)) ? 1 : 0;
        for (int b = 0; b < ntot; ++b) {
          if (tm.identity || !((tm.z >> b) & 1)) continue;
          if (local_of[b] >= 0) zt.ztile |= 1u << local_of[b];
          else zt.zrest |= 1ull << b;
        }
        plan.zterms.push_back(zt);
      }
    }
    if (!pg.items.empty()) {
      struct LC { const int* local_of; } lc{local_of};
      auto dense_x_local = [](const PItem& it, const void* c) -> uint64_t {
        const int* lo = static_cast<const LC*>(c)->local_of;
        uint64_t m = 0;
        for (int b = 0; b < 64; ++b)
          if ((it.cmask >> b) & 1) m |= 1ull << lo[b];
        return m;
      };
      std::vector<Group> rounds =
          schedule(items, pg.items, R, 0, (1ull << t) - 1, ~0ull, 1 << 30,
                   dense_x_local, &lc);
      for (const Group& rg : rounds) {
        RoundRec rr{};
        int reg_of_global[64];
        for (int b = 0; b < 64; ++b) reg_of_global[b] = -1;
        const int nr = int(rg.pos.size());
        for (int j = 0; j < 4; ++j) rr.pos[j] = j < nr ? rg.pos[j] : -1;
        for (int j = 0; j < nr; ++j) reg_of_global[pg.pos[rg.pos[j]]] = j;
        rr.op_begin = int(plan.xops.size());
        for (int idx : rg.items) {
          const TermMask& tm = terms[item_term[idx]];
          ExpXOp op{};
          op.term = item_term[idx];
          uint32_t zreg = 0;
          for (int b = 0; b < ntot; ++b) {
            const int r = reg_of_global[b];
            if ((tm.x >> b) & 1) {
              assert(r >= 0);
              op.xreg |= 1u << r;
            }
            if ((tm.z >> b) & 1) {
              if (r >= 0) zreg |= 1u << r;
              else op.zrest |= 1ull << b;
            }
          }
          for (int e = 0; e < 16; ++e)
            if (__builtin_popcount(e & zreg) & 1) op.sign16 |= 1u << e;
          op.use_im = tm.phase & 1;
          op.negate = ((tm.phase & 3) == 1 || (tm.phase & 3) == 2) ? 1 : 0;
          if (__builtin_popcount(zreg) <= 1) {
            const int zs = zreg ? 1 + __builtin_ctz(zreg) : 0;
            op.code = 1 + (op.use_im * 5 + zs) * 15 + (int(op.xreg) - 1);
          }
          plan.xops.push_back(op);
        }
        rr.op_end = int(plan.xops.size());
        plan.rounds.push_back(rr);
      }
    }
    pr.round_end = int(plan.rounds.size());
    plan.passes.push_back(pr);
  }
  return plan;
}

// Low index bits that every tile of a sharded gate pass keeps (runs of
// 2^L * 8 bytes): the pass after a qubit swap loads its tiles from the peers
// over NVLink, where longer runs pay (TFQB_SHARDED_LOW_BITS, default kLowBits).
static int ShardedLowBits() {
  static const int v = [] {
    const char* e = getenv("TFQB_SHARDED_LOW_BITS");
    const int r = e && *e ? atoi(e) : kLowBits;
    return r < kLowBits ? kLowBits : (r > 7 ? 7 : r);
  }();
  return v;
}

static int GatherLowBits() {
  static const int v = [] {
    const char* e = getenv("TFQB_GATHER_LOW_BITS");
    const int r = e && *e ? atoi(e) : kLowBits;
    return r < kLowBits ? kLowBits : (r > 7 ? 7 : r);
  }();
  return v;
}

ShardedPlan PlanSharded(const CircuitT& c, int g,
                        const std::vector<TermMask>& terms) {
  ShardedPlan sp;
  sp.n = c.n;
  sp.g = g;
  sp.n_local = c.n - g;
  const int n = c.n, nl = sp.n_local;
  std::vector<int> phys(n), inv(n);
  for (int b = 0; b < n; ++b) phys[b] = inv[b] = b;
  const std::vector<PItem> items = fuse_forward(c);

  auto map_mask = [&](uint64_t m) {
    uint64_t r = 0;
    for (int b = 0; b < n; ++b)
      if ((m >> b) & 1) r |= 1ull << phys[b];
    return r;
  };
  auto to_physical = [&](const PItem& it) {
    PItem p = it;
    p.t[0] = phys[it.t[0]];
    if (it.nt == 2) p.t[1] = phys[it.t[1]];
    p.cmask = map_mask(it.cmask);
    p.cbits = map_mask(it.cbits);
    p.qmask = map_mask(it.qmask);
    return p;
  };
  auto swap_item = [&](int pa, int pb) {   // exact-enough SWAP = SP^1
    GateT gt;
    gt.kind = kSP;
    gt.nq = 2;
    gt.bit[0] = pa;
    gt.bit[1] = pb;
    gt.nparams = 3;
    gt.p[0].value = 1.f;
    gt.p[1].value = 1.f;
    gt.p[2].value = 0.f;
    return item_from_gate(gt, -1, kMatGate);
  };

  std::vector<PItem> seg;   // current segment, physical coordinates
  auto close_segment = [&]() {
    if (seg.empty()) return;
    sp.stages.push_back(ShardedStage{0, int(sp.gate_plans.size())});
    // A segment that follows a qubit swap loads its first pass from the peers
    // over NVLink.  TFQB_GATHER_LOW_BITS=6 gives that pass 512-byte runs, which
    // reach the NVLink rate from a single peer (2 GPUs, 34 qubits: 642 GB/s per
    // GPU against 281 with 128-byte runs) but cost a pass; with 7 peers in
    // flight the 128-byte runs already reach 574 GB/s and the narrow default
    // is faster end to end (8 GPUs, 36 qubits: 0.643 s against 0.664 s;
    // profiles/r02k_sharded_36q_8gpu_*.jsonl), so the default stays kLowBits.
    sp.gate_plans.push_back(build(seg, n, kRegBits, kTileMax, ShardedLowBits(), nl, nullptr,
                                  sp.n_exchanges > 0 ? GatherLowBits() : 0));
    sp.gate_plans.back().after_exchange = sp.n_exchanges > 0;
    seg.clear();
  };
  // make the logical qubits in `keep` local: evict g others, exchange
  auto exchange = [&](uint64_t keep_logical, const std::vector<long>& next_use) {
    // candidates: local logical qubits not in keep, farthest next use first
    std::vector<int> cand;
    for (int q = 0; q < n; ++q)
      if (phys[q] < nl && !((keep_logical >> q) & 1)) cand.push_back(q);
    std::stable_sort(cand.begin(), cand.end(),
                     [&](int a, int b) { return next_use[a] > next_use[b]; });
    assert(int(cand.size()) >= g);
    uint64_t evict = 0;
    for (int k = 0; k < g; ++k) evict |= 1ull << cand[k];
    // bring the evicted qubits to the top g local positions
    for (int k = 0; k < g; ++k) {
      const int e = cand[k];
      if (phys[e] >= nl - g) continue;
      int top = -1;
      for (int p = nl - g; p < nl; ++p)
        if (!((evict >> inv[p]) & 1)) { top = p; break; }
      assert(top >= 0);
      const int f = inv[top], pe = phys[e];
      seg.push_back(swap_item(pe, top));
      phys[e] = top; inv[top] = e;
      phys[f] = pe; inv[pe] = f;
    }
    close_segment();
    sp.stages.push_back(ShardedStage{1, sp.n_exchanges++});
    for (int k = 0; k < g; ++k) {
      const int pl = nl - g + k, pg = nl + k;
      const int a = inv[pl], b = inv[pg];
      phys[a] = pg; inv[pg] = a;
      phys[b] = pl; inv[pl] = b;
    }
  };

  // List scheduling: within a segment take every item whose dense targets
  // are local and whose qubits are not blocked by an earlier skipped item;
  // when nothing more fits, swap qubits for the first skipped item.
  std::vector<char> done(items.size(), 0);
  size_t remaining = items.size();
  while (remaining) {
    uint64_t blocked = 0;
    long first_skipped = -1;
    for (size_t i = 0; i < items.size(); ++i) {
      if (done[i]) continue;
      const PItem& it = items[i];
      bool ok = !(it.qmask & blocked);
      if (ok && it.dense) {
        if (phys[it.t[0]] >= nl) ok = false;
        if (it.nt == 2 && phys[it.t[1]] >= nl) ok = false;
      }
      if (ok) {
        seg.push_back(to_physical(it));
        done[i] = 1;
        --remaining;
      } else {
        blocked |= it.qmask;
        if (first_skipped < 0) first_skipped = long(i);
      }
    }
    if (!remaining) break;
    // make room for the first skipped item (and the skipped ones after it,
    // as long as their dense qubits fit)
    uint64_t keep = 0;
    for (size_t i = size_t(first_skipped); i < items.size(); ++i) {
      if (done[i] || !items[i].dense) continue;
      uint64_t need = 1ull << items[i].t[0];
      if (items[i].nt == 2) need |= 1ull << items[i].t[1];
      if (__builtin_popcountll(keep | need) > nl - g) break;
      keep |= need;
      if (__builtin_popcountll(keep) >= nl - g) break;
    }
    // next dense use among the not-yet-done items
    std::vector<long> nu(n, 1L << 40);
    for (size_t j = items.size(); j-- > 0;) {
      if (done[j] || !items[j].dense) continue;
      nu[items[j].t[0]] = long(j);
      if (items[j].nt == 2) nu[items[j].t[1]] = long(j);
    }
    exchange(keep, nu);
  }
  close_segment();

  // expectation: evaluate what the layout allows, swap, repeat
  std::vector<TermMask> todo = terms;   // identity=true marks "done / skip"
  for (int round = 0; round < 64; ++round) {
    std::vector<TermMask> ph(todo.size());
    bool any = false;
    for (size_t k = 0; k < todo.size(); ++k) {
      ph[k] = todo[k];
      if (todo[k].identity) continue;
      any = true;
      ph[k].x = map_mask(todo[k].x);
      ph[k].z = map_mask(todo[k].z);
    }
    if (!any) break;
    ExpectationPlan ep = PlanExpectation(n, ph, false, kTileMax, kLowBits, nl);
    std::vector<char> deferred(todo.size(), 0);
    for (int k : ep.deferred_terms) deferred[k] = 1;
    for (size_t k = 0; k < todo.size(); ++k)
      if (!deferred[k]) todo[k].identity = true;
    const bool more = !ep.deferred_terms.empty();
    uint64_t keep = 0;
    if (more) {
      // make the qubits of the first deferred terms local (as many as fit)
      for (int k : ep.deferred_terms) {
        const uint64_t want = keep | terms[k].x;
        if (__builtin_popcountll(want) > nl - g) break;
        keep = want;
      }
    }
    sp.stages.push_back(ShardedStage{2, int(sp.exp_plans.size())});
    sp.exp_plans.push_back(std::move(ep));
    if (!more) break;
    std::vector<long> nu(n, 0);
    for (int q = 0; q < n; ++q) nu[q] = ((keep >> q) & 1) ? 0 : 1;
    exchange(keep, nu);
  }
  sp.final_phys = phys;
  return sp;
}

}  // namespace tfqb
