// See jit.h.
#include "jit.h"

#include <dlfcn.h>

#include <nvtx3/nvToolsExt.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <thread>
#include <vector>

namespace tfqb {

namespace {

constexpr int kT = kTileMax;            // specialised kernels use the full tile

int EnvInt(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}
// Geometry of the specialised kernels.  Straight-line code makes instruction
// fetch the first limiter (ncu: stall_no_inst), so the body is kept small: one
// register group per thread, and `iters` loop trips per round re-use a round's
// code from the instruction cache.
struct Geometry {
  int groups, threads, min_blocks;
  int tiles;   // tiles per CTA: `tiles` sub-groups of `threads` threads run the
               // same code in step, so a fetched instruction serves all of them
};
Geometry FwdGeometry() {
  static const Geometry g = [] {
    Geometry x;
    x.groups = EnvInt("TFQB_JIT_FWD_GROUPS", 1);
    x.threads = EnvInt("TFQB_JIT_FWD_THREADS", 128);
    x.min_blocks = EnvInt("TFQB_JIT_FWD_MINB", 5);
    x.tiles = EnvInt("TFQB_JIT_FWD_TILES", 1);
    return x;
  }();
  return g;
}
Geometry AdjGeometry(int reg_bits) {
  static const Geometry g = [] {
    Geometry x;
    x.groups = 1;
    x.threads = EnvInt("TFQB_JIT_ADJ_THREADS", 256);
    x.min_blocks = EnvInt("TFQB_JIT_ADJ_MINB", 2);
    x.tiles = EnvInt("TFQB_JIT_ADJ_TILES", 1);
    return x;
  }();
  Geometry r = g;
  if (reg_bits == 4 && !getenv("TFQB_JIT_ADJ_MINB")) {   // 32 amplitudes per thread
    r.threads = EnvInt("TFQB_JIT_ADJ_THREADS", 128);
    r.min_blocks = 2;
  }
  return r;
}

// TFQB_JIT_LIFT=0: rotations as 2x2 real matrices again (A/B switch)
bool LiftEnabled() {
  static const bool v = EnvInt("TFQB_JIT_LIFT", 1) != 0;
  return v;
}

uint32_t swz_host(uint32_t i) { return i ^ (((i >> 4) ^ (i >> 8)) & 15u); }

// expression scattering bit k of `var` to position pos[k] (runs of
// consecutive positions become one shift)
std::string Scatter(const std::string& var, const std::vector<int>& pos) {
  if (pos.empty()) return "0ull";
  std::ostringstream o;
  size_t k = 0;
  bool first = true;
  while (k < pos.size()) {
    size_t len = 1;
    while (k + len < pos.size() && pos[k + len] == pos[k] + int(len)) ++len;
    if (!first) o << " | ";
    first = false;
    o << "((unsigned long long)((" << var << " >> " << k << ") & "
      << ((1u << len) - 1u) << "u) << " << pos[k] << ")";
    k += len;
  }
  return o.str();
}

int SeqTilesOf(const DevicePlan& plan, bool adjoint, int tpc, int pass);

struct Gen {
  const DevicePlan& plan;
  const PassRec& pr;
  bool adj;
  int R, G, nthr, iters, minb, tpc;
  std::ostringstream o;
  std::vector<int> grad_slots;   // grad op ordinal -> output slot
  bool pf;                       // the job cannot see a global phase
  std::vector<std::pair<int, int>> real_mats;   // (float4 offset, 1 row / 2 col phased)
  // packed FP32 instructions (FFMA2 / FMUL2) per amplitude of the forward ops
  // emitted so far: the FP32-pipe floor of the pass (bench.py roofline.fp32)
  double packed_per_amp = 0.0;

  int pass_index;
  Gen(const DevicePlan& p, int pass, bool adjoint, bool phase_free)
      : plan(p), pr(p.passes[pass]), adj(adjoint), pf(phase_free), pass_index(pass) {
    R = plan.reg_bits;
    const Geometry geo = adj ? AdjGeometry(R) : FwdGeometry();
    G = geo.groups;
    nthr = geo.threads;
    minb = geo.min_blocks;
    tpc = geo.tiles;
    while (tpc > 1 && (1 << (plan.n_alloc - kT)) < tpc) tpc >>= 1;
    iters = (1 << (kT - R)) / (nthr * G);
  }

  static void pair_of(int idx, int* hi, int* lo) {
    static const int H[6] = {1, 2, 2, 3, 3, 3}, L[6] = {0, 0, 1, 0, 1, 2};
    *hi = H[idx];
    *lo = L[idx];
  }
  std::string A(int g) const { return "a" + std::to_string(g); }
  std::string Lm(int g) const { return "l" + std::to_string(g); }
  std::string GB(int g) const { return "gb" + std::to_string(g); }
  std::string Bit(int g, int pos) const {
    return "int((" + GB(g) + " >> " + std::to_string(pos) + ") & 1ull)";
  }
  std::string Sm(const OpRec& op, int extra = 0) const {
    return "(s_mat + " + std::to_string((op.mat_off >> 1) + extra) + ")";
  }
  // F<...>(x, args) on psi and / or lambda of every group, as the op targets
  void Apply(const OpRec& op, const std::string& fn, const std::string& args) {
    const int tgt = adj ? op.target : kTgtPsi;
    for (int g = 0; g < G; ++g) {
      if (tgt & kTgtPsi) o << "      " << fn << "(" << A(g) << args << ");\n";
      if (adj && (tgt & kTgtLam)) o << "      " << fn << "(" << Lm(g) << args << ");\n";
    }
  }
  // Gradient partials of consecutive grad ops wait in gq0..gq3 and are reduced
  // four at a time: a transposing butterfly (levels 16 and 8 halve the number
  // of values a lane carries, level 4 finishes) needs 4 shuffles for 4 gates
  // instead of 12, and one fp64 shared-memory update instead of four.
  std::vector<int> pending;      // grad ordinals waiting in gq0..
  static bool GradBatch() {
    static const bool v = EnvInt("TFQB_GRAD_BATCH", 1) != 0;
    return v;
  }
  void EmitSingleReduce(int k, const std::string& var) {
    o << "      { float gs = " << var << ";\n"
         "      gs += __shfl_xor_sync(0xffffffffu, gs, 16);\n"
         "      gs += __shfl_xor_sync(0xffffffffu, gs, 8);\n"
         "      gs += __shfl_xor_sync(0xffffffffu, gs, 4);\n"
         "      if ((tid & 31) < 4) s_grad["
      << k << " * kGradSlots + (threadIdx.x >> 5) * 4 + (tid & 3)] += 2.0 * double(gs); }\n";
  }
  void FlushGrad() {
    for (size_t i = 0; i < pending.size(); ++i)
      EmitSingleReduce(pending[i], "gq" + std::to_string(i));
    pending.clear();
  }
  void GradReduce(const OpRec& op) {
    const int k = int(grad_slots.size());
    grad_slots.push_back(op.grad_slot);
    // The reference sums in double (tfq_adj_grad_op.cc:272-273).  Here: float
    // per thread (<= 16 amplitudes) and over the three shuffle levels (8
    // lanes), fp64 from the shared-memory accumulator on (iterations, warps,
    // tiles).  TFQB_GRAD_SHUFFLE=double makes the shuffles fp64 too: measured
    // (profiles/r02_float32_floor.jsonl) it changes no digit of the error
    // against the double-state yardstick and costs 7% of the adjoint
    // throughput, so it is not the default.
    static const bool float_shuffle = [] {
      const char* e = getenv("TFQB_GRAD_SHUFFLE");
      return !(e && *e == 'd');
    }();
    if (float_shuffle && GradBatch()) {
      o << "      gq" << pending.size() << " = gv;\n";
      pending.push_back(k);
      if (pending.size() < 4) return;
      const int k0 = pending[0];
      pending.clear();
      o << "      {  // gradient slots " << k0 << ".." << k0 + 3 << "\n"
           "        const bool h16 = (tid & 16u) != 0u, h8 = (tid & 8u) != 0u;\n"
           "        const float gt0 = h16 ? gq0 : gq2, gt1 = h16 ? gq1 : gq3;\n"
           "        float gk0 = h16 ? gq2 : gq0, gk1 = h16 ? gq3 : gq1;\n"
           "        gk0 += __shfl_xor_sync(0xffffffffu, gt0, 16);\n"
           "        gk1 += __shfl_xor_sync(0xffffffffu, gt1, 16);\n"
           "        const float gt2 = h8 ? gk0 : gk1;\n"
           "        float r = h8 ? gk1 : gk0;\n"
           "        r += __shfl_xor_sync(0xffffffffu, gt2, 8);\n"
           "        r += __shfl_xor_sync(0xffffffffu, r, 4);\n"
           "        if (!(tid & 4u)) s_grad[("
        << k0 << " + 2 * int(h16) + int(h8)) * kGradSlots + (threadIdx.x >> 5) * 4 + (tid & 3)] += "
           "2.0 * double(r);\n      }\n";
      return;
    }
    if (float_shuffle) {
      EmitSingleReduce(k, "gv");
      return;
    }
    o << "      { double gd = double(gv);\n"
         "      gd += __shfl_xor_sync(0xffffffffu, gd, 16);\n"
         "      gd += __shfl_xor_sync(0xffffffffu, gd, 8);\n"
         "      gd += __shfl_xor_sync(0xffffffffu, gd, 4);\n"
         "      if ((tid & 31) < 4) s_grad["
      << k << " * kGradSlots + (threadIdx.x >> 5) * 4 + (tid & 3)] += 2.0 * gd; }\n";
  }
  // selector of a thread-constant diagonal (D0 / S0 / AdjD0)
  std::string Sel(const OpRec& op, int g) const {
    std::string s = Bit(g, op.dpos0);
    if (op.dpos1 >= 0) s = "(2 * " + s + " + " + Bit(g, op.dpos1) + ")";
    return s;
  }
  // (s0, s1) of a one-register-bit diagonal
  void Sel1(const OpRec& op, int g, std::string* s0, std::string* s1) const {
    if (op.dpos1 < 0) {
      *s0 = "0";
      *s1 = "1";
    } else if (op.dreg0 >= 0) {
      const std::string c = Bit(g, op.dpos1);
      *s0 = c;
      *s1 = "(2 + " + c + ")";
    } else {
      const std::string c = Bit(g, op.dpos0);
      *s0 = "(2 * " + c + ")";
      *s1 = "(2 * " + c + " + 1)";
    }
  }

  bool EmitOp(const OpRec& op, bool* has_ph, bool* has_neg) {
    const int c = op.code;
    const std::string Rs = std::to_string(R);
    auto tmpl1 = [&](const char* name, int j) {
      return std::string(name) + "<" + Rs + ", " + std::to_string(j) + ">";
    };
    auto tmpl2 = [&](const char* name, int idx) {
      int hi, lo;
      pair_of(idx, &hi, &lo);
      return std::string(name) + "<" + Rs + ", " + std::to_string(hi) + ", " +
             std::to_string(lo) + ">";
    };
    const int tgt = adj ? op.target : kTgtPsi;
    o << "    {  // op code " << c << "\n";
    // dense 2x2 on register bit j, matrix at float4 offset `extra` of the op
    auto g1 = [&](int j, int extra) {
      const int raw = pf ? int((op.pad_ >> (4 * j)) & 15u) : 0;
      const int flag = raw & 7;
      // proper rotations as three shears (pass_device.cuh `lift`): kRealCol + 8
      const bool lift = LiftEnabled() && (flag == 4 || (raw & 8));
      if (flag == 4 && !adj) {
        real_mats.emplace_back((op.mat_off >> 1) + extra, 4 + (lift ? 8 : 0));   // X^t, no gradient gate
        Apply(op, tmpl1(lift ? "g1_ximag_lift" : "g1_ximag", j), ", " + Sm(op, extra));
        packed_per_amp += lift ? 1.5 : 2;
      } else if (flag && !adj) {
        // setup modes: 0 = D R, 1 = R D, 2 = R alone
        real_mats.emplace_back((op.mat_off >> 1) + extra,
                               (flag == 1 ? 0 : flag == 2 ? 1 : 2) + (lift ? 8 : 0));
        const std::string fn =
            std::string(flag == 1 ? "g1_rowreal" : flag == 2 ? "g1_colreal" : "g1_real") +
            (lift ? "_lift" : "");
        Apply(op, tmpl1(fn.c_str(), j), ", " + Sm(op, extra));
        packed_per_amp += (flag == 3 ? 2 : 3) - (lift ? 0.5 : 0);
      } else {
        Apply(op, tmpl1("g1_packed", j), ", " + Sm(op, extra));
        packed_per_amp += 4;      // 2 complex multiply-adds of 2 packed ops
      }
    };
    if (c >= kCodeG1 && c < kCodeG1 + 4) {
      g1(c - kCodeG1, 0);
    } else if (c >= kCodeG2 && c < kCodeG2 + 6) {
      Apply(op, tmpl2("g2_packed", c - kCodeG2), ", " + Sm(op));
      packed_per_amp += 8;
    } else if (c == kCodeG1Run) {
      int extra = 0;
      for (int j = 3; j >= 0; --j) {
        if (!((op.ident_mask >> j) & 1u)) continue;
        g1(j, extra);
        extra += 4;
      }
    } else if (c == kCodeD0) {
      for (int g = 0; g < G; ++g) {
        const std::string f = Sm(op) + "[" + Sel(op, g) + "]";
        if (adj) {
          if (tgt & kTgtPsi) o << "      scale_all<" << Rs << ">(" << A(g) << ", " << f << ");\n";
          if (tgt & kTgtLam) o << "      scale_all<" << Rs << ">(" << Lm(g) << ", " << f << ");\n";
        } else {
          o << "      ph" << g << " = cmulf(ph" << g << ", plain(" << f << "));\n";
          *has_ph = true;
        }
      }
    } else if (c >= kCodeD1 && c < kCodeD1 + 4) {
      const bool own = op.dpos1 < 0;
      const std::string d0 = own && (op.ident_mask & 1u) ? "false" : "true";
      const std::string d1 = own && (op.ident_mask & 2u) ? "false" : "true";
      packed_per_amp += (d0 == "true" ? 1 : 0) + (d1 == "true" ? 1 : 0);
      for (int g = 0; g < G; ++g) {
        std::string s0, s1;
        Sel1(op, g, &s0, &s1);
        const std::string args = ", " + Sm(op) + "[" + s0 + "], " + Sm(op) + "[" + s1 +
                                 "], " + d0 + ", " + d1;
        const std::string fn = tmpl1("diag1", c - kCodeD1);
        if (tgt & kTgtPsi) o << "      " << fn << "(" << A(g) << args << ");\n";
        if (adj && (tgt & kTgtLam)) o << "      " << fn << "(" << Lm(g) << args << ");\n";
      }
    } else if (c >= kCodeD2 && c < kCodeD2 + 6) {
      Apply(op, tmpl2("diag2", c - kCodeD2),
            ", " + Sm(op) + ", " + std::to_string(op.ident_mask) + "u");
      packed_per_amp += 2.0 * (4 - __builtin_popcount(op.ident_mask & 15u)) / 4.0;
    } else if (c == kCodeS0 || c == kCodeS0Run) {
      for (int g = 0; g < G; ++g) {
        std::string cond;
        if (c == kCodeS0) {
          cond = "((" + std::to_string(op.ident_mask) + "u >> " + Sel(op, g) + ") & 1u)";
        } else {
          std::ostringstream p;
          p << "((" << (op.ident_mask & 1u) << " + __popcll(" << GB(g) << " & "
            << op.crest_mask << "ull)";
          if (op.crest_bits)
            p << " + __popcll(" << GB(g) << " & (" << GB(g) << " >> " << op.dpos0
              << ") & " << op.crest_bits << "ull)";
          if (op.pad_)
            p << " + __popcll(" << GB(g) << " & (" << GB(g) << " >> " << op.dpos1
              << ") & " << op.pad_ << "ull)";
          p << ") & 1)";
          cond = p.str();
        }
        if (adj) {
          o << "      if (" << cond << ") {\n";
          if (tgt & kTgtPsi) o << "        sign_all<" << Rs << ">(" << A(g) << ");\n";
          if (tgt & kTgtLam) o << "        sign_all<" << Rs << ">(" << Lm(g) << ");\n";
          o << "      }\n";
        } else {
          o << "      ng" << g << " ^= uint32_t(" << cond << ");\n";
          *has_neg = true;
        }
      }
    } else if (c >= kCodeS1 && c < kCodeS1 + 4) {
      for (int g = 0; g < G; ++g) {
        std::string s0, s1;
        Sel1(op, g, &s0, &s1);
        const std::string m = std::to_string(op.ident_mask) + "u";
        const std::string args = ", ((" + m + " >> " + s0 + ") & 1u) != 0, ((" + m +
                                 " >> " + s1 + ") & 1u) != 0";
        const std::string fn = tmpl1("sign1", c - kCodeS1);
        if (tgt & kTgtPsi) o << "      " << fn << "(" << A(g) << args << ");\n";
        if (adj && (tgt & kTgtLam)) o << "      " << fn << "(" << Lm(g) << args << ");\n";
      }
    } else if (c >= kCodeS2 && c < kCodeS2 + 6) {
      Apply(op, tmpl2("sign2", c - kCodeS2), ", " + std::to_string(op.ident_mask) + "u");
    } else if (!adj) {
      return false;
    } else {
      // ---- gradient ops (adjoint plans)
      o << "      float gv = 0.f;\n";
      const std::string al = "(a0, l0, ";
      if (c >= kCodeGrad1 && c < kCodeGrad1 + 4) {
        o << "      gv += " << tmpl1("grad1_packed", c - kCodeGrad1) << al << Sm(op) << ");\n";
      } else if (c >= kCodeGrad2 && c < kCodeGrad2 + 6) {
        o << "      gv += " << tmpl2("grad2_packed", c - kCodeGrad2) << al << Sm(op) << ");\n";
      } else if (c == kCodeGradD0) {
        o << "      gv += gdiag0<" << Rs << ">" << al << Sm(op) << "[" << Sel(op, 0) << "]);\n";
      } else if (c >= kCodeGradD1 && c < kCodeGradD1 + 4) {
        std::string s0, s1;
        Sel1(op, 0, &s0, &s1);
        o << "      gv += " << tmpl1("gdiag1", c - kCodeGradD1) << al << Sm(op) << "[" << s0
          << "], " << Sm(op) << "[" << s1 << "]);\n";
      } else if (c >= kCodeGradD2 && c < kCodeGradD2 + 6) {
        o << "      gv += " << tmpl2("gdiag2", c - kCodeGradD2) << al << Sm(op) << ");\n";
      } else if (c >= kCodeAdj1 && c < kCodeAdj1 + 4) {
        const int j = c - kCodeAdj1;
        static const bool no_adj_real = getenv("TFQB_JIT_NO_ADJ_REAL") != nullptr;
        const int raw = pf && !no_adj_real ? int((op.pad_ >> (4 * j)) & 15u) : 0;
        const int flag = raw & 7;
        const bool lift = LiftEnabled() && (flag == 4 || (raw & 8));
        if (flag == 3) {
          real_mats.emplace_back(op.mat_off >> 1, 3 + (lift ? 8 : 0));
          o << "      gv += " << tmpl1(lift ? "adj1_real_lift" : "adj1_real", j) << al << Sm(op)
            << ");\n";
        } else if (flag == 4) {
          real_mats.emplace_back(op.mat_off >> 1, 5 + (lift ? 8 : 0));    // X^t with its gradient gate
          o << "      gv += " << tmpl1(lift ? "adj1_ximag_lift" : "adj1_ximag", j) << al << Sm(op)
            << ");\n";
        } else {
          o << "      gv += " << tmpl1("adj1_packed", j) << al << Sm(op) << ");\n";
        }
      } else if (c >= kCodeAdj2 && c < kCodeAdj2 + 6) {
        o << "      gv += " << tmpl2("adj2_packed", c - kCodeAdj2) << al << Sm(op) << ");\n";
      } else if (c == kCodeAdjD0) {
        const std::string sel = Sel(op, 0);
        o << "      { const int sel = " << sel << ";\n        gv += adjd0<" << Rs << ">" << al
          << Sm(op) << "[sel], " << Sm(op) << "[4 + sel]); }\n";
      } else if (c >= kCodeAdjD1 && c < kCodeAdjD1 + 4) {
        std::string s0, s1;
        Sel1(op, 0, &s0, &s1);
        o << "      { const int s0 = " << s0 << ", s1 = " << s1 << ";\n        gv += "
          << tmpl1("adjd1", c - kCodeAdjD1) << al << Sm(op) << "[s0], " << Sm(op) << "[s1], "
          << Sm(op) << "[4 + s0], " << Sm(op) << "[4 + s1]); }\n";
      } else if (c >= kCodeAdjD2 && c < kCodeAdjD2 + 6) {
        o << "      gv += " << tmpl2("adjd2", c - kCodeAdjD2) << al << Sm(op) << ");\n";
      } else {
        return false;
      }
      GradReduce(op);
    }
    o << "    }\n";
    return true;
  }

  // ops [k0, k1): consecutive fused adjoint steps of diagonal gates (thread-
  // constant selector, or one register bit): see pass_device.cuh conj_products
  void EmitDiagAdjRun(int k0, int k1) {
    const std::string Rs = std::to_string(R);
    o << "    {  // run of " << (k1 - k0) << " diagonal adjoint steps\n"
      << "      float2 cj[" << (1 << R) << "];\n"
      << "      conj_products<" << Rs << ">(a0, l0, cj);\n"
      << "      const float2 ctot = csum_all<" << Rs << ">(cj);\n"
      << "      float2 phr = make_float2(1.f, 0.f);\n";
    bool any_d0 = false;
    for (int k = k0; k < k1; ++k) {
      const OpRec& op = plan.ops[k];
      o << "      {  // op code " << op.code << "\n";
      if (op.code == kCodeAdjD0) {
        any_d0 = true;
        o << "        const int sel = " << Sel(op, 0) << ";\n"
          << "        float gv = re_hs(" << Sm(op) << "[4 + sel], " << Sm(op) << "[sel], ctot);\n";
        GradReduce(op);
        o << "        phr = cmulf(phr, plain(" << Sm(op) << "[sel]));\n";
      } else if (op.code >= kCodeAdjD2) {
        int jh, jl;
        pair_of(op.code - kCodeAdjD2, &jh, &jl);
        const std::string t = "<" + Rs + ", " + std::to_string(jh) + ", " + std::to_string(jl) + ">";
        // entries that are exactly 1 for every row (ZZ^t = diag(1, w, w, 1)):
        // their gradient entry is exactly 0 and psi, lambda need no multiply
        const uint32_t ident = op.creg_bits & 15u;
        o << "        float2 S4[4];\n        csum_2bit" << t << "(cj, S4);\n"
          << "        float gv = 0.f;\n";
        for (int e4 = 0; e4 < 4; ++e4) {
          if ((ident >> e4) & 1u) continue;
          o << "        gv += re_hs(" << Sm(op) << "[" << 4 + e4 << "], " << Sm(op) << "[" << e4
            << "], S4[" << e4 << "]);\n";
        }
        GradReduce(op);
        o << "        diag2" << t << "(a0, " << Sm(op) << ", " << ident << "u);\n"
          << "        diag2" << t << "(l0, " << Sm(op) << ", " << ident << "u);\n";
      } else {
        const int j = op.code - kCodeAdjD1;
        std::string s0, s1;
        Sel1(op, 0, &s0, &s1);
        const bool own = op.dpos1 < 0;       // a 1-qubit diagonal on register bit j
        const bool id0 = own && (op.creg_bits & 1u), id1 = own && (op.creg_bits & 2u);
        o << "        const int s0 = " << s0 << ", s1 = " << s1 << ";\n"
          << "        float2 S0, S1;\n        csum_bit<" << Rs << ", " << j << ">(cj, S0, S1);\n"
          << "        float gv = 0.f;\n";
        if (!id0)
          o << "        gv += re_hs(" << Sm(op) << "[4 + s0], " << Sm(op) << "[s0], S0);\n";
        if (!id1)
          o << "        gv += re_hs(" << Sm(op) << "[4 + s1], " << Sm(op) << "[s1], S1);\n";
        GradReduce(op);
        o << "        diag1<" << Rs << ", " << j << ">(a0, " << Sm(op) << "[s0], " << Sm(op)
          << "[s1], " << (id0 ? "false" : "true") << ", " << (id1 ? "false" : "true") << ");\n"
          << "        diag1<" << Rs << ", " << j << ">(l0, " << Sm(op) << "[s0], " << Sm(op)
          << "[s1], " << (id0 ? "false" : "true") << ", " << (id1 ? "false" : "true") << ");\n";
      }
      o << "      }\n";
    }
    if (any_d0)
      o << "      scale_all_c<" << Rs << ">(a0, phr);\n      scale_all_c<" << Rs << ">(l0, phr);\n";
    o << "    }\n";
  }

  bool EmitRound(const RoundRec& rr) {
    const int first_op = plan.rounds[pr.round_begin].op_begin;
    (void)first_op;
    uint32_t so[4] = {0, 0, 0, 0};
    for (int j = 0; j < R; ++j) so[j] = swz_host(1u << rr.pos[j]);
    o << "  {  // round on tile bits";
    for (int j = 0; j < R; ++j) o << " " << rr.pos[j];
    o << "\n";
    if (iters > 1) o << "#pragma unroll 1\n  for (uint32_t it = 0; it < " << iters << "u; ++it) {\n";
    else o << "  { const uint32_t it = 0;\n";
    for (int g = 0; g < G; ++g) {
      o << "    uint32_t b" << g << " = it * " << nthr * G << "u + " << g * nthr << "u + tid;\n";
      for (int j = 0; j < R; ++j) {
        const uint32_t lo = (1u << rr.pos[j]) - 1u;
        o << "    b" << g << " = ((b" << g << " & ~" << lo << "u) << 1) | (b" << g << " & " << lo
          << "u);\n";
      }
      o << "    const uint32_t sb" << g << " = swz(b" << g << ");\n";
      o << "    float2 " << A(g) << "[" << (1 << R) << "];\n";
      if (adj) o << "    float2 " << Lm(g) << "[" << (1 << R) << "];\n";
      for (int e = 0; e < (1 << R); ++e) {
        uint32_t x = 0;
        for (int j = 0; j < R; ++j)
          if (e & (1 << j)) x ^= so[j];
        o << "    " << A(g) << "[" << e << "] = s_psi[sb" << g << " ^ " << x << "u];\n";
        if (adj) o << "    " << Lm(g) << "[" << e << "] = s_lam[sb" << g << " ^ " << x << "u];\n";
      }
      o << "    const unsigned long long " << GB(g) << " = rank_base | base | (b" << g << " & "
        << ((1u << pr.low_bits) - 1u) << "u) | hi_of(b" << g << " >> " << pr.low_bits << ");\n";
      if (!adj) {
        o << "    float2 ph" << g << " = make_float2(1.f, 0.f);\n    uint32_t ng" << g
          << " = 0u;\n";
      }
    }
    if (adj) o << "    float gq0 = 0.f, gq1 = 0.f, gq2 = 0.f, gq3 = 0.f;\n"
                  "    (void)gq0; (void)gq1; (void)gq2; (void)gq3;\n";
    bool has_ph = false, has_neg = false;
    static const bool no_diag_run = getenv("TFQB_JIT_NO_DIAG_RUN") != nullptr;
    auto diag_adj = [&](const OpRec& op) {
      return adj && pf && !no_diag_run && (op.code == kCodeAdjD0 ||
                           (op.code >= kCodeAdjD1 && op.code < kCodeAdjD1 + 4) ||
                           (op.code >= kCodeAdjD2 && op.code < kCodeAdjD2 + 6));
    };
    for (int k = rr.op_begin; k < rr.op_end;) {
      int e = k;
      while (e < rr.op_end && diag_adj(plan.ops[e])) ++e;
      if (e - k >= 2) {
        EmitDiagAdjRun(k, e);
        k = e;
        continue;
      }
      if (!EmitOp(plan.ops[k], &has_ph, &has_neg)) return false;
      ++k;
    }
    if (adj) FlushGrad();
    for (int g = 0; g < G; ++g) {
      if (!adj && has_ph) {
        if (g == 0) packed_per_amp += 2;      // one complex scale per amplitude
        if (has_neg) o << "    if (ng" << g << " & 1u) ph" << g << " = cneg2(ph" << g << ");\n";
        o << "    scale_all_c<" << R << ">(" << A(g) << ", ph" << g << ");\n";
      } else if (!adj && has_neg) {
        o << "    if (ng" << g << " & 1u) sign_all<" << R << ">(" << A(g) << ");\n";
      }
      for (int e = 0; e < (1 << R); ++e) {
        uint32_t x = 0;
        for (int j = 0; j < R; ++j)
          if (e & (1 << j)) x ^= so[j];
        o << "    s_psi[sb" << g << " ^ " << x << "u] = " << A(g) << "[" << e << "];\n";
        if (adj) o << "    s_lam[sb" << g << " ^ " << x << "u] = " << Lm(g) << "[" << e << "];\n";
      }
    }
    o << "  }\n  __syncthreads();\n  }\n";
    return true;
  }

  bool Run(std::string* out) {
    const int L = pr.low_bits;
    const int n_entries = (pr.mat_len + 1) / 2;
    std::vector<int> hi_pos, comp_pos;
    for (int k = L; k < kT; ++k) hi_pos.push_back(pr.tile_pos[k]);
    for (int k = 0; k < pr.n_comp; ++k) comp_pos.push_back(pr.comp_pos[k]);

    // rounds first (they fill grad_slots), header afterwards
    for (int r = pr.round_begin; r < pr.round_end; ++r) {
      if (!EmitRound(plan.rounds[r])) return false;
    }
    const std::string rounds_src = o.str();
    o.str("");

    const int n_grad = int(grad_slots.size());
    const int seq = SeqTilesOf(plan, adj, tpc, pass_index);
    const int grad_sl = (nthr * tpc / 32) * 4;
    const int cta = nthr * tpc;
    o << "// generated by quantum_b200/csrc/jit.cc: one gate pass, specialised\n";
    // experiment switch (1: adjoint passes, 2: forward passes too); measured neutral, off
    if (EnvInt("TFQB_JIT_SIGN_XOR", 0) >= (adj ? 1 : 2)) o << "#define TFQB_SIGN_XOR 1\n";
    o << PassDeviceSource() << "\n";
    o << "constexpr int kGradSlots = " << grad_sl << ";\n";
    if (n_grad > 0) {
      o << "__device__ const int kSlotOf[" << n_grad << "] = {";
      for (int i = 0; i < n_grad; ++i) o << (i ? ", " : "") << grad_slots[i];
      o << "};\n";
    }
    if (!real_mats.empty()) {
      o << "__device__ const int kRealOff[" << real_mats.size() << "] = {";
      for (size_t i = 0; i < real_mats.size(); ++i) o << (i ? ", " : "") << real_mats[i].first;
      o << "};\n__device__ const int kRealCol[" << real_mats.size() << "] = {";
      for (size_t i = 0; i < real_mats.size(); ++i) o << (i ? ", " : "") << real_mats[i].second;
      o << "};\n";
    }
    o << "__device__ __forceinline__ unsigned long long hi_of(uint32_t h) {\n  return "
      << Scatter("h", hi_pos) << ";\n}\n";
    o << "__device__ __forceinline__ unsigned long long base_of(unsigned long long v) {\n"
         "  return "
      << Scatter("v", comp_pos) << ";\n}\n";
    o << "extern \"C\" __global__ void __launch_bounds__(" << cta << ", " << minb << ")\n"
      << "tfqb_jit_pass(float2* __restrict__ psi, float2* __restrict__ lam, size_t row_stride,\n"
         "              const float* __restrict__ mats, size_t mat_row_stride,\n"
         "              double* __restrict__ grad_out, int n_slots, int init_mode,\n"
         "              unsigned long long rank_base,\n"
         "              const float2* const* __restrict__ peer_tab, int peer_shift,\n"
         "              unsigned long long peer_self) {\n"
         "  extern __shared__ __align__(16) unsigned char smem_raw[];\n"
         "  // sub-group `sub` of the CTA owns tile blockIdx.x * tiles + sub\n"
         "  const uint32_t tid = threadIdx.x & "
      << (nthr - 1) << "u;\n"
      << "  const uint32_t sub = threadIdx.x / " << nthr << "u;\n"
      << "  const size_t row = blockIdx.y;\n"
         "  float2* s_psi = reinterpret_cast<float2*>(smem_raw) + sub * 4096u;\n";
    if (adj) o << "  float2* s_lam = s_psi + " << tpc * 4096 << ";\n";
    o << "  float4* s_mat = reinterpret_cast<float4*>(reinterpret_cast<float2*>(smem_raw) + "
      << (adj ? 8192 : 4096) * tpc << ");\n";
    if (adj) o << "  double* s_grad = reinterpret_cast<double*>(s_mat + " << n_entries << ");\n";
    o << "  {\n"
         "    const float2* src = reinterpret_cast<const float2*>(mats + row * mat_row_stride + "
      << pr.mat_begin << ");\n"
      << "    for (uint32_t i = threadIdx.x; i < " << n_entries << "u; i += " << cta << "u) {\n"
      << "      const float2 m = src[i];\n"
         "      s_mat[i] = make_float4(m.x, m.x, -m.y, m.y);\n"
         "    }\n"
         "  }\n";
    if (!real_mats.empty()) {
      // phased-real gates: rewrite their staged matrices once per CTA
      o << "  __syncthreads();\n  for (uint32_t i = threadIdx.x; i < " << real_mats.size()
        << "u; i += " << cta << "u) {\n"
        << "    const int rc = kRealCol[i] & 7;\n    const bool lift = kRealCol[i] >= 8;\n"
           "    if (rc >= 4) phased_ximag_setup(s_mat + kRealOff[i], rc == 5, lift);\n"
           "    else phased_real_setup(s_mat + kRealOff[i], rc, lift);\n  }\n";
    }
    if (adj && n_grad > 0)
      o << "  for (uint32_t i = threadIdx.x; i < " << n_grad * grad_sl << "u; i += " << cta
        << "u) s_grad[i] = 0.0;\n";
    o << "  float2* g_psi = psi + row * row_stride;\n";
    if (adj) o << "  float2* g_lam = lam + row * row_stride;\n";
    o << "#pragma unroll 1\n  for (uint32_t ti = 0; ti < " << seq << "u; ++ti) {\n"
      << "  const unsigned long long base = base_of((blockIdx.x * " << seq << "u + ti) * " << tpc
      << "u + sub);\n";
    const bool product = !adj && pr.init_bits > 0;
    if (product) {
      o << "  if (init_mode == 2) {\n"
           "    __syncthreads();   // init vectors live in s_mat\n"
           "    const float4* iv = s_mat + "
        << (pr.init_off >> 1) << ";\n"
        << "    const unsigned long long fb = base | rank_base;\n"
           "    float2 C = make_float2(1.f, 0.f);\n";
      for (int k = 0; k < pr.n_comp; ++k) {
        const int b = pr.comp_pos[k];
        o << "    C = cmulf(C, plain(iv[" << 2 * b << " + int((fb >> " << b << ") & 1ull)]));\n";
      }
      for (int b = pr.n_comp + kT; b < pr.init_bits; ++b)
        o << "    C = cmulf(C, plain(iv[" << 2 * b << " + int((fb >> " << b << ") & 1ull)]));\n";
      o << "    for (uint32_t blk = tid; blk < 256u; blk += " << nthr << "u) {\n"
        << "      float2 T[16];\n      T[0] = C;\n";
      for (int k = 4; k < kT; ++k)
        o << "      T[0] = cmulf(T[0], plain(iv[" << 2 * pr.tile_pos[k] << " + int((blk >> "
          << (k - 4) << ") & 1u)]));\n";
      for (int k = 0; k < 4; ++k) {
        o << "      { const float2 u0 = plain(iv[" << 2 * pr.tile_pos[k] << "]), u1 = plain(iv["
          << 2 * pr.tile_pos[k] + 1 << "]);\n";
        for (int e = 0; e < (1 << k); ++e)
          o << "        { const float2 lo = T[" << e << "]; T[" << e << "] = cmulf(lo, u0); T["
            << (e | (1 << k)) << "] = cmulf(lo, u1); }\n";
        o << "      }\n";
      }
      o << "#pragma unroll\n      for (int e = 0; e < 16; ++e) s_psi[swz(blk * 16 + e)] = T[e];\n"
           "    }\n  } else {\n";
    } else {
      o << "  {\n";
    }
    o << "    for (uint32_t c0 = 0; c0 < 2048u; c0 += " << nthr * 4 << "u) {\n"
      << "      float4 v[4];\n";
    if (adj) o << "      float4 w[4];\n";
    o << "#pragma unroll\n      for (int u = 0; u < 4; ++u) {\n"
      << "        const uint32_t i = 2u * (c0 + u * " << nthr << "u + tid);\n"
      << "        const unsigned long long g = base | (i & " << ((1u << L) - 1u)
      << "u) | hi_of(i >> " << L << ");\n"
      << "        if (init_mode == 3)   // qubit swap fused into the load (peer shards over NVLink)\n"
         "          v[u] = __ldcs(reinterpret_cast<const float4*>(peer_tab[g >> peer_shift] +\n"
         "                        (peer_self | (g & ((1ull << peer_shift) - 1ull)))));\n"
         "        else if (init_mode) v[u] = make_float4((g | rank_base) == 0 ? 1.f : 0.f, 0.f, 0.f, 0.f);\n"
         "        else v[u] = *reinterpret_cast<const float4*>(g_psi + g);\n";
    if (adj) o << "        w[u] = *reinterpret_cast<const float4*>(g_lam + g);\n";
    o << "      }\n#pragma unroll\n      for (int u = 0; u < 4; ++u) {\n"
      << "        const uint32_t i = 2u * (c0 + u * " << nthr << "u + tid);\n"
      << "        const uint32_t x0 = swz(i), x1 = x0 ^ 1u;\n"
         "        s_psi[x0] = make_float2(v[u].x, v[u].y);\n"
         "        s_psi[x1] = make_float2(v[u].z, v[u].w);\n";
    if (adj)
      o << "        s_lam[x0] = make_float2(w[u].x, w[u].y);\n"
           "        s_lam[x1] = make_float2(w[u].z, w[u].w);\n";
    o << "      }\n    }\n  }\n  __syncthreads();\n";
    o << rounds_src;
    o << "  for (uint32_t c = tid; c < 2048u; c += " << nthr << "u) {\n"
      << "    const uint32_t i = 2u * c;\n"
      << "    const unsigned long long g = base | (i & " << ((1u << L) - 1u) << "u) | hi_of(i >> "
      << L << ");\n"
      << "    const uint32_t x0 = swz(i), x1 = x0 ^ 1u;\n"
         "    const float2 p0 = s_psi[x0], p1 = s_psi[x1];\n"
         "    *reinterpret_cast<float4*>(g_psi + g) = make_float4(p0.x, p0.y, p1.x, p1.y);\n";
    if (adj)
      o << "    const float2 q0 = s_lam[x0], q1 = s_lam[x1];\n"
           "    *reinterpret_cast<float4*>(g_lam + g) = make_float4(q0.x, q0.y, q1.x, q1.y);\n";
    o << "  }\n";
    if (seq > 1) o << "  __syncthreads();   // the tile buffers are reused by the next tile\n";
    o << "  }\n";
    if (adj && n_grad > 0) {
      o << "  __syncthreads();\n  for (uint32_t i = threadIdx.x; i < " << n_grad << "u; i += " << cta << "u) {\n"
        << "    double v = 0.0;\n"
           "    for (int k = 0; k < kGradSlots; ++k) v += s_grad[i * kGradSlots + k];\n"
           "    const int slot = kSlotOf[i];\n"
           "    if (slot >= 0 && v != 0.0)\n"
           "      atomicAdd(&grad_out[row * size_t(n_slots) + slot], v);\n"
           "  }\n";
    }
    o << "}\n";
    *out = o.str();
    return true;
  }
};

bool OpJitable(const OpRec& op, bool adj) {
  const int c = op.code;
  if (c == kCodeSlow) return false;
  if (!adj && c >= kCodeGrad1 && c < kCodeS0) return false;
  return c >= 0 && c <= kCodeS0Run;
}

}  // namespace

// Tiles one CTA works through one after the other (same row): the per-CTA
// prologue (matrix staging, the fp64 phase-free rewrites, gradient-slot
// zeroing and the final slot reduction) is paid once per `seq` tiles.
static int SeqTiles(const DevicePlan& plan, bool adjoint, int tpc, int pass) {
  static const int fwd = EnvInt("TFQB_JIT_FWD_SEQ", 0), adj = EnvInt("TFQB_JIT_ADJ_SEQ", 8);
  // forward default: 8 tiles per CTA; 32 for states of 2^28 amplitudes and
  // more, where the tiles a CTA walks share their 2 MB pages (34 qubits on one
  // GPU: 0.800 s with 8, 0.759 s with 32, 0.760 s with 128)
  int k = adjoint ? adj : fwd > 0 ? fwd : plan.n_alloc >= 28 ? 32 : 8;
  // shards of a state spread over several GPUs (n_alloc = local bits < n): the
  // first pass of a segment that follows a qubit swap loads its tiles from the
  // peers over NVLink, and two tiles per CTA moved 592 GB/s per GPU where
  // eight moved 545 (36 qubits on 8 GPUs,
  // profiles/r03_sharded_36q_seq_ab.jsonl); the local passes keep eight (34
  // qubits on one GPU: 0.77 s against 0.84 s)
  static const int gather = EnvInt("TFQB_JIT_GATHER_SEQ", 2);
  if (!adjoint && plan.after_exchange && pass == 0 && k > gather) k = gather;
  // TFQB_DETERMINISTIC=1: one CTA walks every tile of its row, so the gradient
  // slots of a row are summed in one place and one order (no fp64 atomics
  // between CTAs)
  if (adjoint && EnvInt("TFQB_DETERMINISTIC", 0) != 0) k = 1 << 30;
  if (k < 1) k = 1;
  while (k & (k - 1)) k &= k - 1;           // power of two
  const long tiles = (1l << (plan.n_alloc - kT)) / tpc;
  while (k > 1 && k > tiles) k >>= 1;
  return k;
}

namespace {
int SeqTilesOf(const DevicePlan& plan, bool adjoint, int tpc, int pass) {
  return SeqTiles(plan, adjoint, tpc, pass);
}
}  // namespace

static int TilesPerCta(const DevicePlan& plan, bool adjoint) {
  int tpc = adjoint ? AdjGeometry(plan.reg_bits).tiles : FwdGeometry().tiles;
  while (tpc > 1 && (1 << (plan.n_alloc - kT)) < tpc) tpc >>= 1;
  return tpc;
}
int JitPassTiles(const DevicePlan& plan, bool adjoint, int pass) {
  const int tpc = TilesPerCta(plan, adjoint);
  return tpc * SeqTiles(plan, adjoint, tpc, pass);
}
int JitPassThreads(const DevicePlan& plan, bool adjoint) {
  return (adjoint ? AdjGeometry(plan.reg_bits).threads : FwdGeometry().threads) *
         TilesPerCta(plan, adjoint);
}

size_t JitPassSmem(const DevicePlan& plan, int pass, bool adjoint) {
  const PassRec& pr = plan.passes[pass];
  int n_grad = 0;
  if (adjoint && pr.round_end > pr.round_begin)
    for (int k = plan.rounds[pr.round_begin].op_begin;
         k < plan.rounds[pr.round_end - 1].op_end; ++k)
      if (plan.ops[k].code >= kCodeGrad1 && plan.ops[k].code < kCodeS0) ++n_grad;
  return (size_t(adjoint ? 16 : 8) << kT) * TilesPerCta(plan, adjoint) +
         size_t((pr.mat_len + 1) / 2) * 16 +
         size_t(n_grad) * (JitPassThreads(plan, adjoint) / 32) * 4 * 8 + 16;
}

bool PassIsJitable(const DevicePlan& plan, int pass, bool adj) {
  const PassRec& pr = plan.passes[pass];
  if (pr.tile_bits != kT) return false;
  if (plan.reg_bits != 3 && plan.reg_bits != 4) return false;
  if (!adj && plan.reg_bits != 4) return false;
  long cost = 0;
  for (int r = pr.round_begin; r < pr.round_end; ++r) {
    const RoundRec& rr = plan.rounds[r];
    for (int j = 0; j < plan.reg_bits; ++j)
      if (rr.pos[j] < 0) return false;
    for (int k = rr.op_begin; k < rr.op_end; ++k) {
      const OpRec& op = plan.ops[k];
      if (!OpJitable(op, adj)) return false;
      const int c = op.code;
      const bool two = (c >= kCodeG2 && c < kCodeG2 + 6) || (c >= kCodeGrad2 && c < kCodeGrad2 + 6) ||
                       (c >= kCodeAdj2 && c < kCodeAdj2 + 6);
      cost += two ? 8 : 2;
    }
  }
  return cost <= 1200;    // keeps NVRTC + ptxas time to a few seconds
}

std::string GeneratePassSource(const DevicePlan& plan, int pass, bool adjoint,
                               bool phase_free) {
  Gen g(plan, pass, adjoint, phase_free);
  std::string out;
  if (!g.Run(&out)) return std::string();
  return out;
}

double PassPackedFp32PerAmplitude(const DevicePlan& plan, int pass, bool phase_free) {
  if (!PassIsJitable(plan, pass, false)) return -1.0;
  Gen g(plan, pass, false, phase_free);
  std::string out;
  if (!g.Run(&out)) return -1.0;
  return g.packed_per_amp;
}


// ---------------------------------------------------------------------------
// PauliSum expectation pass (kernels.cu: expect_pass_kernel), specialised:
// every X/Y-type term is a straight-line FMA chain into its own register
// accumulator (reduced once per CTA), Z-type terms are owned by threads.
// ---------------------------------------------------------------------------
namespace {
constexpr int kExpThreads = 256;      // operator-accumulation kernel
constexpr int kExpMaxXops = 40;
// expectation kernel: 256 threads x 2 CTAs per SM.  TFQB_JIT_EXP_THREADS=128
// (4 CTAs per SM, a tile's load latency hidden behind three other tiles) was
// measured and is 7% slower on C2's two passes (profiles/r03h_*)
int ExpThreads() {
  static const int v = EnvInt("TFQB_JIT_EXP_THREADS", 256) == 128 ? 128 : 256;
  return v;
}
}

bool ExpectPassIsJitable(const ExpectationPlan& plan, int pass) {
  const PassRec& pr = plan.passes[pass];
  if (pr.tile_bits != kT) return false;
  const int nz = pass == 0 ? int(plan.zterms.size()) : 0;
  if (nz > ExpThreads()) return false;
  int nx = 0;
  for (int r = pr.round_begin; r < pr.round_end; ++r) {
    const RoundRec& rr = plan.rounds[r];
    for (int j = 0; j < 4; ++j)
      if (rr.pos[j] < 0) return false;
    nx += rr.op_end - rr.op_begin;
  }
  return nx <= kExpMaxXops && nx + nz > 0;
}

size_t JitExpectSmem(const ExpectationPlan& plan, int pass) {
  const PassRec& pr = plan.passes[pass];
  const bool with_z = pass == 0 && !plan.zterms.empty();
  int nx = 0;
  if (pr.round_end > pr.round_begin)
    nx = plan.rounds[pr.round_end - 1].op_end - plan.rounds[pr.round_begin].op_begin;
  return (size_t(8) << kT) + (with_z ? (size_t(4) << kT) : 0) +
         size_t(nx) * (ExpThreads() / 32) * 4 + 16;
}

int JitExpectThreads() { return ExpThreads(); }
int JitAccumThreads() { return kExpThreads; }

std::string GenerateExpectSource(const ExpectationPlan& plan, int pass) {
  const PassRec& pr = plan.passes[pass];
  const int L = pr.low_bits;
  const int nz = pass == 0 ? int(plan.zterms.size()) : 0;
  const int first_op = pr.round_end > pr.round_begin ? plan.rounds[pr.round_begin].op_begin : 0;
  const int nx = pr.round_end > pr.round_begin
                     ? plan.rounds[pr.round_end - 1].op_end - first_op
                     : 0;
  std::vector<int> hi_pos, comp_pos;
  for (int k = L; k < kT; ++k) hi_pos.push_back(pr.tile_pos[k]);
  for (int k = 0; k < pr.n_comp; ++k) comp_pos.push_back(pr.comp_pos[k]);
  const uint32_t lowmask = (1u << L) - 1u;
  std::ostringstream o;
  o << "// generated by quantum_b200/csrc/jit.cc: one PauliSum expectation pass\n"
    << PassDeviceSource() << "\n";
  if (nz > 0) {
    o << "__device__ const uint32_t kZtile[" << nz << "] = {";
    for (int k = 0; k < nz; ++k) o << (k ? ", " : "") << plan.zterms[k].ztile << "u";
    o << "};\n__device__ const unsigned long long kZrest[" << nz << "] = {";
    for (int k = 0; k < nz; ++k) o << (k ? ", " : "") << plan.zterms[k].zrest << "ull";
    o << "};\n__device__ const int kZneg[" << nz << "] = {";
    for (int k = 0; k < nz; ++k) o << (k ? ", " : "") << plan.zterms[k].negate;
    o << "};\n__device__ const int kZterm[" << nz << "] = {";
    for (int k = 0; k < nz; ++k) o << (k ? ", " : "") << plan.zterms[k].term;
    o << "};\n";
  }
  if (nx > 0) {
    o << "__device__ const int kXterm[" << nx << "] = {";
    for (int k = 0; k < nx; ++k) o << (k ? ", " : "") << plan.xops[first_op + k].term;
    o << "};\n";
  }
  o << "__device__ __forceinline__ unsigned long long hi_of(uint32_t h) {\n  return "
    << Scatter("h", hi_pos) << ";\n}\n"
    << "__device__ __forceinline__ unsigned long long base_of(unsigned long long v) {\n"
       "  return "
    << Scatter("v", comp_pos) << ";\n}\n";
  o << "extern \"C\" __global__ void __launch_bounds__(" << ExpThreads() << ", " << 512 / ExpThreads() << ")\n"
       "tfqb_jit_expect(const float2* __restrict__ psi, size_t row_stride,\n"
       "                unsigned long long n_tiles, unsigned long long rank_base,\n"
       "                double* __restrict__ per_term, int n_terms) {\n"
       "  extern __shared__ __align__(16) unsigned char smem_raw[];\n"
       "  const uint32_t tid = threadIdx.x;\n"
       "  const size_t row = blockIdx.y;\n"
       "  float2* s_psi = reinterpret_cast<float2*>(smem_raw);\n"
       "  float* s_p = reinterpret_cast<float*>(s_psi + 4096);\n"
       "  float* s_red = s_p + "
    << (nz > 0 ? 4096 : 0) << ";\n"
    << "  const float2* g_psi = psi + row * row_stride;\n";
  for (int k = 0; k < nx; ++k) o << "  float x" << k << " = 0.f;\n";
  if (nz > 0) o << "  double zacc = 0.0;\n";
  o << "  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {\n"
       "    const unsigned long long base = base_of(tile);\n"
       "    for (uint32_t h = 0; h < " << 2048 / (8 * ExpThreads()) << "u; ++h) {\n"
       "      float4 v[8];\n#pragma unroll\n      for (int u = 0; u < 8; ++u) {\n"
       "        const uint32_t i = 2u * ((h * 8u + u) * "
    << ExpThreads() << "u + tid);\n"
    << "        v[u] = *reinterpret_cast<const float4*>(g_psi + (base | (i & " << lowmask
    << "u) | hi_of(i >> " << L << ")));\n      }\n"
    << "#pragma unroll\n      for (int u = 0; u < 8; ++u) {\n"
       "        const uint32_t i = 2u * ((h * 8u + u) * "
    << ExpThreads() << "u + tid);\n"
    << "        const uint32_t x0 = swz(i), x1 = x0 ^ 1u;\n"
       "        s_psi[x0] = make_float2(v[u].x, v[u].y);\n"
       "        s_psi[x1] = make_float2(v[u].z, v[u].w);\n";
  if (nz > 0)
    o << "        s_p[x0] = fmaf(v[u].x, v[u].x, v[u].y * v[u].y);\n"
         "        s_p[x1] = fmaf(v[u].z, v[u].z, v[u].w * v[u].w);\n";
  o << "      }\n    }\n    __syncthreads();\n";
  if (nz > 0) {
    for (int lvl = 0; lvl < kT; lvl += 4)
      o << "    wht_level<4>(s_p, 4096u, " << lvl << ", int(tid), " << ExpThreads()
        << ");\n    __syncthreads();\n";
    o << "    if (tid < " << nz << "u) {\n"
      << "      const float v = s_p[swz(kZtile[tid])];\n"
         "      const int neg = (__popcll((base | rank_base) & kZrest[tid]) & 1) ^ kZneg[tid];\n"
         "      zacc += double(neg ? -v : v);\n    }\n";
  }
  for (int r = pr.round_begin; r < pr.round_end; ++r) {
    const RoundRec& rr = plan.rounds[r];
    uint32_t so[4];
    for (int j = 0; j < 4; ++j) so[j] = swz_host(1u << rr.pos[j]);
    o << "    {  // round on tile bits " << rr.pos[0] << " " << rr.pos[1] << " " << rr.pos[2]
      << " " << rr.pos[3] << "\n#pragma unroll 1\n      for (uint32_t it = 0; it < "
      << 256 / ExpThreads() << "u; ++it) {\n      uint32_t b = it * " << ExpThreads()
      << "u + tid;\n";
    for (int j = 0; j < 4; ++j) {
      const uint32_t lo = (1u << rr.pos[j]) - 1u;
      o << "      b = ((b & ~" << lo << "u) << 1) | (b & " << lo << "u);\n";
    }
    o << "      const uint32_t sb = swz(b);\n      float2 a[16];\n";
    for (int e = 0; e < 16; ++e) {
      uint32_t x = 0;
      for (int j = 0; j < 4; ++j)
        if (e & (1 << j)) x ^= so[j];
      o << "      a[" << e << "] = s_psi[sb ^ " << x << "u];\n";
    }
    o << "      const unsigned long long gb = rank_base | base | (b & " << lowmask
      << "u) | hi_of(b >> " << L << ");\n      (void)gb;\n";
    for (int k = rr.op_begin; k < rr.op_end; ++k) {
      const ExpXOp& op = plan.xops[k];
      const int idx = k - first_op;
      o << "      {\n        float v = ";
      if (op.code != 0) {
        const int cc = op.code - 1;
        const int xr = cc % 15 + 1, zs = (cc / 15) % 5, im = cc / 75;
        o << "xterm_fixed<" << xr << ", " << zs << ", " << (im ? "true" : "false") << ">(a);\n";
      } else {
        o << "xterm_pairs<" << op.xreg << ">(a, " << op.sign16 << "u)." << (op.use_im ? "y" : "x")
          << ";\n";
      }
      if (op.zrest)
        o << "        if ((__popcll(gb & " << op.zrest << "ull) + " << op.negate
          << ") & 1) v = -v;\n        x" << idx << " += v;\n";
      else
        o << "        x" << idx << (op.negate ? " -= v;\n" : " += v;\n");
      o << "      }\n";
    }
    o << "      }\n    }\n";
  }
  o << "    __syncthreads();   // tile buffers are reused by the next tile\n  }\n";
  // ---- one reduction per CTA
  if (nx > 0) {
    for (int k = 0; k < nx; ++k) {
      o << "  { float v = x" << k << ";\n"
        << "    v += __shfl_xor_sync(0xffffffffu, v, 16); v += __shfl_xor_sync(0xffffffffu, v, 8);\n"
           "    v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 2);\n"
           "    v += __shfl_xor_sync(0xffffffffu, v, 1);\n"
           "    if ((tid & 31) == 0) s_red["
        << k * (ExpThreads() / 32) << " + (tid >> 5)] = v; }\n";
    }
    o << "  __syncthreads();\n  if (tid < " << nx << "u) {\n    double v = 0.0;\n"
      << "    for (int w = 0; w < " << ExpThreads() / 32 << "; ++w) v += double(s_red[tid * "
      << ExpThreads() / 32 << " + w]);\n"
      << "    // pairs are counted once: the mirrored half contributes the same\n"
         "    if (v != 0.0) atomicAdd(&per_term[row * size_t(n_terms) + kXterm[tid]], 2.0 * v);\n"
         "  }\n";
  }
  if (nz > 0)
    o << "  if (tid < " << nz << "u && zacc != 0.0)\n"
         "    atomicAdd(&per_term[row * size_t(n_terms) + kZterm[tid]], zacc);\n";
  o << "}\n";
  return o.str();
}


// ---------------------------------------------------------------------------
// lambda = sum_j g_j sum_t c_t P_t psi (kernels.cu: accum_pass_kernel), specialised
// ---------------------------------------------------------------------------
size_t JitAccumSmem(const ExpectationPlan& plan, int pass, int n_terms) {
  const bool with_z = pass == 0 && !plan.zterms.empty();
  return (size_t(16) << kT) + (with_z ? (size_t(4) << kT) : 0) + size_t(n_terms) * 4 + 16;
}

std::string GenerateAccumSource(const ExpectationPlan& plan, int pass) {
  const PassRec& pr = plan.passes[pass];
  const int L = pr.low_bits;
  const int nz = pass == 0 ? int(plan.zterms.size()) : 0;
  std::vector<int> hi_pos, comp_pos;
  for (int k = L; k < kT; ++k) hi_pos.push_back(pr.tile_pos[k]);
  for (int k = 0; k < pr.n_comp; ++k) comp_pos.push_back(pr.comp_pos[k]);
  const uint32_t lowmask = (1u << L) - 1u;
  std::ostringstream o;
  o << "// generated by quantum_b200/csrc/jit.cc: one operator-accumulation pass\n"
    << PassDeviceSource() << "\n"
    << "struct DevTermJ { unsigned long long x, z; float coeff; int phase, op, identity; };\n";
  if (nz > 0) {
    o << "__device__ const uint32_t kZtile[" << nz << "] = {";
    for (int k = 0; k < nz; ++k) o << (k ? ", " : "") << plan.zterms[k].ztile << "u";
    o << "};\n__device__ const unsigned long long kZrest[" << nz << "] = {";
    for (int k = 0; k < nz; ++k) o << (k ? ", " : "") << plan.zterms[k].zrest << "ull";
    o << "};\n__device__ const int kZneg[" << nz << "] = {";
    for (int k = 0; k < nz; ++k) o << (k ? ", " : "") << plan.zterms[k].negate;
    o << "};\n__device__ const int kZterm[" << nz << "] = {";
    for (int k = 0; k < nz; ++k) o << (k ? ", " : "") << plan.zterms[k].term;
    o << "};\n";
  }
  o << "__device__ __forceinline__ unsigned long long hi_of(uint32_t h) {\n  return "
    << Scatter("h", hi_pos) << ";\n}\n"
    << "__device__ __forceinline__ unsigned long long base_of(unsigned long long v) {\n"
       "  return "
    << Scatter("v", comp_pos) << ";\n}\n";
  o << "extern \"C\" __global__ void __launch_bounds__(" << kExpThreads << ", 2)\n"
       "tfqb_jit_accum(const float2* __restrict__ psi, float2* __restrict__ lam,\n"
       "               size_t row_stride, const DevTermJ* __restrict__ terms, int n_terms,\n"
       "               const float* __restrict__ downstream, int n_ops, int accumulate,\n"
       "               unsigned long long n_tiles) {\n"
       "  extern __shared__ __align__(16) unsigned char smem_raw[];\n"
       "  const uint32_t tid = threadIdx.x;\n"
       "  const size_t row = blockIdx.y;\n"
       "  float2* s_psi = reinterpret_cast<float2*>(smem_raw);\n"
       "  float2* s_out = s_psi + 4096;\n"
       "  float* s_p = reinterpret_cast<float*>(s_out + 4096);\n"
       "  float* s_lead = s_p + "
    << (nz > 0 ? 4096 : 0) << ";\n"
    << "  for (int i = tid; i < n_terms; i += " << kExpThreads << ") {\n"
       "    // `leading = downstream * coefficient`, terms below 1e-5 are skipped\n"
       "    // (util_qsim.h:378-383)\n"
       "    const float lead = __fmul_rn(downstream[row * n_ops + terms[i].op], terms[i].coeff);\n"
       "    s_lead[i] = fabsf(lead) < 1e-5f ? 0.f : lead;\n  }\n  __syncthreads();\n"
       "  const float2* g_psi = psi + row * row_stride;\n"
       "  float2* g_lam = lam + row * row_stride;\n"
       "  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {\n"
       "    const unsigned long long base = base_of(tile);\n"
       "    for (uint32_t c0 = 0; c0 < 2048u; c0 += "
    << kExpThreads * 4 << "u) {\n      float4 v[4], w[4];\n#pragma unroll\n"
    << "      for (int u = 0; u < 4; ++u) {\n        const uint32_t i = 2u * (c0 + u * "
    << kExpThreads << "u + tid);\n"
    << "        const unsigned long long g = base | (i & " << lowmask << "u) | hi_of(i >> " << L
    << ");\n        v[u] = *reinterpret_cast<const float4*>(g_psi + g);\n"
       "        w[u] = accumulate ? *reinterpret_cast<const float4*>(g_lam + g)\n"
       "                          : make_float4(0.f, 0.f, 0.f, 0.f);\n      }\n"
       "#pragma unroll\n      for (int u = 0; u < 4; ++u) {\n        const uint32_t i = 2u * (c0 + u * "
    << kExpThreads << "u + tid);\n"
    << "        const uint32_t x0 = swz(i), x1 = x0 ^ 1u;\n"
       "        s_psi[x0] = make_float2(v[u].x, v[u].y);\n"
       "        s_psi[x1] = make_float2(v[u].z, v[u].w);\n"
       "        s_out[x0] = make_float2(w[u].x, w[u].y);\n"
       "        s_out[x1] = make_float2(w[u].z, w[u].w);\n";
  if (nz > 0) o << "        s_p[x0] = 0.f;\n        s_p[x1] = 0.f;\n";
  o << "      }\n    }\n    __syncthreads();\n";
  if (nz > 0) {
    o << "    if (tid < " << nz << "u) {   // sparse coefficient vector over the tile's Z characters\n"
      << "      float v = s_lead[kZterm[tid]];\n"
         "      if ((__popcll(base & kZrest[tid]) & 1) ^ kZneg[tid]) v = -v;\n"
         "      if (v != 0.f) atomicAdd(&s_p[swz(kZtile[tid])], v);\n    }\n    __syncthreads();\n";
    for (int lvl = 0; lvl < kT; lvl += 4)
      o << "    wht_level<4>(s_p, 4096u, " << lvl << ", int(tid), " << kExpThreads
        << ");\n    __syncthreads();\n";
    o << "    for (uint32_t i = tid; i < 4096u; i += " << kExpThreads << "u) {\n"
      << "      const float c = s_p[i];\n      const float2 a = s_psi[i];\n"
         "      float2 q = s_out[i];\n      q.x = fmaf(c, a.x, q.x);\n      q.y = fmaf(c, a.y, q.y);\n"
         "      s_out[i] = q;\n    }\n    __syncthreads();\n";
  }
  for (int r = pr.round_begin; r < pr.round_end; ++r) {
    const RoundRec& rr = plan.rounds[r];
    uint32_t so[4];
    for (int j = 0; j < 4; ++j) so[j] = swz_host(1u << rr.pos[j]);
    o << "    {  // round on tile bits " << rr.pos[0] << " " << rr.pos[1] << " " << rr.pos[2]
      << " " << rr.pos[3] << "\n      uint32_t b = tid;\n";
    for (int j = 0; j < 4; ++j) {
      const uint32_t lo = (1u << rr.pos[j]) - 1u;
      o << "      b = ((b & ~" << lo << "u) << 1) | (b & " << lo << "u);\n";
    }
    o << "      const uint32_t sb = swz(b);\n      float2 a[16], acc[16];\n";
    for (int e = 0; e < 16; ++e) {
      uint32_t x = 0;
      for (int j = 0; j < 4; ++j)
        if (e & (1 << j)) x ^= so[j];
      o << "      a[" << e << "] = s_psi[sb ^ " << x << "u]; acc[" << e << "] = s_out[sb ^ " << x
        << "u];\n";
    }
    o << "      const unsigned long long gb = base | (b & " << lowmask << "u) | hi_of(b >> " << L
      << ");\n      (void)gb;\n";
    for (int k = rr.op_begin; k < rr.op_end; ++k) {
      const ExpXOp& op = plan.xops[k];
      o << "      {\n        float lead = s_lead[" << op.term << "];\n"
        << "        if (lead != 0.f) {\n";
      if (op.zrest) o << "          if (__popcll(gb & " << op.zrest << "ull) & 1) lead = -lead;\n";
      // coefficient lead * i^phase, as accum_pass_kernel
      if (!op.use_im)
        o << "          const float cx = " << (op.negate ? "-lead" : "lead") << ", cy = 0.f;\n";
      else
        o << "          const float cx = 0.f, cy = " << (op.negate ? "lead" : "-lead") << ";\n";
      o << "          xterm_accumulate<" << op.xreg << ">(acc, a, make_float4(cx, cx, -cy, cy), "
        << op.sign16 << "u);\n        }\n      }\n";
    }
    for (int e = 0; e < 16; ++e) {
      uint32_t x = 0;
      for (int j = 0; j < 4; ++j)
        if (e & (1 << j)) x ^= so[j];
      o << "      s_out[sb ^ " << x << "u] = acc[" << e << "];\n";
    }
    o << "    }\n    __syncthreads();\n";
  }
  o << "    for (uint32_t c = tid; c < 2048u; c += " << kExpThreads << "u) {\n"
    << "      const uint32_t i = 2u * c;\n"
    << "      const unsigned long long g = base | (i & " << lowmask << "u) | hi_of(i >> " << L
    << ");\n      const uint32_t x0 = swz(i), x1 = x0 ^ 1u;\n"
       "      const float2 q0 = s_out[x0], q1 = s_out[x1];\n"
       "      *reinterpret_cast<float4*>(g_lam + g) = make_float4(q0.x, q0.y, q1.x, q1.y);\n"
       "    }\n    __syncthreads();\n  }\n}\n";
  return o.str();
}

const char* PassDeviceSource() {
  static const char kSrc[] =
#include "pass_device_src.inc"
      ;
  return kSrc;
}

// ---------------------------------------------------------------------------
// NVRTC + driver API through dlopen (no link-time dependency)
// ---------------------------------------------------------------------------
namespace {

struct Api {
  bool ok = false;
  std::string why;
  int nvrtc_version = 0;
  // nvrtc
  int (*nvrtcCreateProgram)(void**, const char*, const char*, int, const char* const*,
                            const char* const*) = nullptr;
  int (*nvrtcCompileProgram)(void*, int, const char* const*) = nullptr;
  int (*nvrtcGetProgramLogSize)(void*, size_t*) = nullptr;
  int (*nvrtcGetProgramLog)(void*, char*) = nullptr;
  int (*nvrtcGetCUBINSize)(void*, size_t*) = nullptr;
  int (*nvrtcGetCUBIN)(void*, char*) = nullptr;
  int (*nvrtcDestroyProgram)(void**) = nullptr;
  // driver
  int (*cuModuleLoadData)(void**, const void*) = nullptr;
  int (*cuModuleUnload)(void*) = nullptr;
  int (*cuModuleGetFunction)(void**, void*, const char*) = nullptr;
  int (*cuFuncSetAttribute)(void*, int, int) = nullptr;
  int (*cuLaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                        unsigned, void*, void**, void**) = nullptr;
  int (*cuGetErrorString)(int, const char**) = nullptr;
};

template <typename F>
bool Sym(void* h, const char* name, F* f) {
  *f = reinterpret_cast<F>(dlsym(h, name));
  return *f != nullptr;
}

Api& GetApi() {
  static Api api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* off = getenv("TFQB_JIT");
    if (off && *off == '0') {
      api.why = "disabled by TFQB_JIT=0";
      return;
    }
    // Newest NVRTC wins: the 12.9 toolkit's ptxas folds the (im, re) operand
    // swap of the packed FMAs into FFMA2's .F32x2.LO_HI operand swizzle, the
    // 12.8 one bundled with PyTorch (already loaded in a Python process under
    // the same soname) spends a MOV pair on every swap: 1.7x the instructions.
    void* rtc = nullptr;
    int best = -1;
    for (const char* n : {"/usr/local/cuda/lib64/libnvrtc.so.12",
                          "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so.12",
                          "libnvrtc.so"}) {
      void* h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (!h) continue;
      int (*ver)(int*, int*) = nullptr;
      int major = 0, minor = 0;
      if (Sym(h, "nvrtcVersion", &ver) && ver(&major, &minor) == 0 &&
          major * 100 + minor > best) {
        best = major * 100 + minor;
        rtc = h;
      }
    }
    if (!rtc) {
      api.why = "libnvrtc not found";
      return;
    }
    api.nvrtc_version = best;
    void* drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!drv) {
      api.why = "libcuda.so.1 not found";
      return;
    }
    bool ok = Sym(rtc, "nvrtcCreateProgram", &api.nvrtcCreateProgram) &&
              Sym(rtc, "nvrtcCompileProgram", &api.nvrtcCompileProgram) &&
              Sym(rtc, "nvrtcGetProgramLogSize", &api.nvrtcGetProgramLogSize) &&
              Sym(rtc, "nvrtcGetProgramLog", &api.nvrtcGetProgramLog) &&
              Sym(rtc, "nvrtcGetCUBINSize", &api.nvrtcGetCUBINSize) &&
              Sym(rtc, "nvrtcGetCUBIN", &api.nvrtcGetCUBIN) &&
              Sym(rtc, "nvrtcDestroyProgram", &api.nvrtcDestroyProgram) &&
              Sym(drv, "cuModuleLoadData", &api.cuModuleLoadData) &&
              Sym(drv, "cuModuleUnload", &api.cuModuleUnload) &&
              Sym(drv, "cuModuleGetFunction", &api.cuModuleGetFunction) &&
              Sym(drv, "cuFuncSetAttribute", &api.cuFuncSetAttribute) &&
              Sym(drv, "cuLaunchKernel", &api.cuLaunchKernel) &&
              Sym(drv, "cuGetErrorString", &api.cuGetErrorString);
    if (!ok) {
      api.why = "missing NVRTC / driver symbols";
      return;
    }
    api.ok = true;
  });
  return api;
}

std::string DrvErr(Api& api, int rc) {
  const char* s = nullptr;
  api.cuGetErrorString(rc, &s);
  return s ? s : ("CUresult " + std::to_string(rc));
}

}  // namespace

bool JitAvailable(std::string* why) {
  Api& api = GetApi();
  if (!api.ok && why) *why = api.why;
  return api.ok;
}

// ---- source -> cubin: asynchronous, memoised, cached on disk ---------------
// Compilation needs no CUDA context, so every pass of a plan (and the
// expectation / accumulation kernels next to it) compiles on its own host
// thread while the caller goes on; a source that was compiled before -- by
// this process (another device of a multi-GPU context, another program with
// the same pass structure: matrices are data, not text) or by an earlier one
// (TFQB_JIT_CACHE_DIR, default ~/.cache/tfqb_jit; "off" disables) -- is not
// compiled again.
namespace {

struct Cubin {
  std::vector<char> data;
  std::string err;
  double compile_ms = 0.0;   // 0: memo / disk hit
};

struct CubinKey {
  uint64_t a, b;
  bool operator<(const CubinKey& o) const { return a != o.a ? a < o.a : b < o.b; }
};

CubinKey HashSource(const std::string& s, int nvrtc_version) {
  uint64_t a = 1469598103934665603ull ^ uint64_t(nvrtc_version), b = 0x9e3779b97f4a7c15ull;
  for (unsigned char c : s) {
    a = (a ^ c) * 1099511628211ull;
    b = (b + c) * 0xff51afd7ed558ccdull;
    b ^= b >> 29;
  }
  return CubinKey{a, b ^ uint64_t(s.size())};
}

std::string CacheDir() {
  const char* e = getenv("TFQB_JIT_CACHE_DIR");
  if (e && *e) return std::string(e) == "off" ? std::string() : std::string(e);
  const char* h = getenv("HOME");
  if (!h || !*h) return std::string();
  return std::string(h) + "/.cache/tfqb_jit";
}

std::string CachePath(const CubinKey& k) {
  const std::string d = CacheDir();
  if (d.empty()) return d;
  char name[64];
  snprintf(name, sizeof name, "/%016llx%016llx.cubin", (unsigned long long)k.a,
           (unsigned long long)k.b);
  return d + name;
}

std::shared_ptr<Cubin> CompileNow(const std::string& src, const CubinKey& key) {
  auto out = std::make_shared<Cubin>();
  Api& api = GetApi();
  const std::string path = CachePath(key);
  if (!path.empty()) {
    if (FILE* f = fopen(path.c_str(), "rb")) {
      fseek(f, 0, SEEK_END);
      const long n = ftell(f);
      fseek(f, 0, SEEK_SET);
      if (n > 0) {
        out->data.resize(size_t(n));
        if (fread(out->data.data(), 1, size_t(n), f) != size_t(n)) out->data.clear();
      }
      fclose(f);
      if (!out->data.empty()) return out;
    }
  }
  const auto t0 = std::chrono::steady_clock::now();
  nvtxRangePushA("tfqb:nvrtc_compile");
  struct Pop { ~Pop() { nvtxRangePop(); } } pop;
  void* prog = nullptr;
  if (api.nvrtcCreateProgram(&prog, src.c_str(), "tfqb_jit_pass.cu", 0, nullptr, nullptr) != 0) {
    out->err = "nvrtcCreateProgram failed";
    return out;
  }
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
  const int rc = api.nvrtcCompileProgram(prog, 3, opts);
  if (rc != 0) {
    size_t n = 0;
    api.nvrtcGetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) api.nvrtcGetProgramLog(prog, &log[0]);
    out->err = "NVRTC compile failed: " + log.substr(0, 2000);
    api.nvrtcDestroyProgram(&prog);
    return out;
  }
  size_t n = 0;
  api.nvrtcGetCUBINSize(prog, &n);
  out->data.resize(n);
  api.nvrtcGetCUBIN(prog, out->data.data());
  api.nvrtcDestroyProgram(&prog);
  out->compile_ms =
      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (!path.empty()) {      // best effort, atomic: write aside, then rename
    const std::string dir = CacheDir();
    mkdir(dir.substr(0, dir.rfind('/')).c_str(), 0755);
    mkdir(dir.c_str(), 0755);
    const std::string tmp = path + "." + std::to_string(getpid()) + ".tmp";
    if (FILE* f = fopen(tmp.c_str(), "wb")) {
      const bool ok = fwrite(out->data.data(), 1, out->data.size(), f) == out->data.size();
      fclose(f);
      if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
    }
  }
  return out;
}

std::mutex g_cubin_mu;
std::map<CubinKey, std::shared_future<std::shared_ptr<Cubin>>> g_cubins;
std::atomic<int> g_compiling{0};
std::atomic<long long> g_compile_us{0};

std::shared_future<std::shared_ptr<Cubin>> CompileAsync(const std::string& src) {
  Api& api = GetApi();
  const CubinKey key = HashSource(src, api.nvrtc_version);
  std::lock_guard<std::mutex> lock(g_cubin_mu);
  auto it = g_cubins.find(key);
  if (it != g_cubins.end()) return it->second;
  if (g_cubins.size() > 4096) g_cubins.clear();      // cubins are ~100 KiB each
  // no more compiler threads than cores: further requests compile when awaited
  const unsigned cores = std::max(1u, std::thread::hardware_concurrency());
  const bool spawn = unsigned(g_compiling.load()) < cores;
  auto job = [src, key]() {
    g_compiling++;
    std::shared_ptr<Cubin> r = CompileNow(src, key);
    g_compiling--;
    g_compile_us += (long long)(r->compile_ms * 1e3);
    return r;
  };
  std::shared_future<std::shared_ptr<Cubin>> f =
      std::async(spawn ? std::launch::async : std::launch::deferred, job).share();
  g_cubins.emplace(key, f);
  return f;
}

}  // namespace

void JitPrefetch(const std::string& src) {
  if (src.empty() || !GetApi().ok) return;
  CompileAsync(src);
}

double JitCompileSeconds() { return double(g_compile_us.load()) * 1e-6; }
int JitPending() { return g_compiling.load(); }

bool JitCompile(const std::string& src, const char* entry, bool adjoint, int threads,
                size_t smem, JitKernel* out, std::string* err) {
  Api& api = GetApi();
  if (!api.ok) {
    *err = api.why;
    return false;
  }
  const std::shared_ptr<Cubin> cubin = CompileAsync(src).get();
  if (!cubin->err.empty() || cubin->data.empty()) {
    *err = cubin->err.empty() ? "empty cubin" : cubin->err;
    return false;
  }
  // the runtime's primary context must be current on this thread
  cudaFree(nullptr);
  void* mod = nullptr;
  int drc = api.cuModuleLoadData(&mod, cubin->data.data());
  if (drc != 0) {
    *err = "cuModuleLoadData: " + DrvErr(api, drc);
    return false;
  }
  void* fn = nullptr;
  drc = api.cuModuleGetFunction(&fn, mod, entry);
  if (drc == 0)   // CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES = 8
    drc = api.cuFuncSetAttribute(fn, 8, int(smem));
  if (drc != 0) {
    *err = "cuModuleGetFunction / cuFuncSetAttribute: " + DrvErr(api, drc);
    api.cuModuleUnload(mod);
    return false;
  }
  out->module = mod;
  out->func = fn;
  out->threads = threads;
  out->smem = smem;
  out->adjoint = adjoint;
  return true;
}

void JitRelease(JitKernel* k) {
  Api& api = GetApi();
  if (api.ok && k->module) api.cuModuleUnload(k->module);
  k->module = k->func = nullptr;
}

bool JitLaunch(const JitKernel& k, unsigned tiles, unsigned rows, float2* psi,
               float2* lam, size_t row_stride, const float* mats,
               size_t mat_row_stride, double* grad_out, int n_slots,
               int init_mode, unsigned long long rank_base, cudaStream_t s,
               std::string* err, const float2* const* peer_tab, int peer_shift,
               unsigned long long peer_self) {
  Api& api = GetApi();
  void* args[] = {&psi, &lam, &row_stride, &mats, &mat_row_stride,
                  &grad_out, &n_slots, &init_mode, &rank_base,
                  &peer_tab, &peer_shift, &peer_self};
  const int rc = api.cuLaunchKernel(k.func, tiles, rows, 1, unsigned(k.threads), 1, 1,
                                    unsigned(k.smem), s, args, nullptr);
  if (rc != 0) {
    *err = "cuLaunchKernel: " + DrvErr(api, rc);
    return false;
  }
  return true;
}

bool JitLaunchExpect(const JitKernel& k, unsigned ctas, unsigned rows, const float2* psi,
                     size_t row_stride, unsigned long long n_tiles,
                     unsigned long long rank_base, double* per_term, int n_terms,
                     cudaStream_t s, std::string* err) {
  Api& api = GetApi();
  void* args[] = {&psi, &row_stride, &n_tiles, &rank_base, &per_term, &n_terms};
  const int rc = api.cuLaunchKernel(k.func, ctas, rows, 1, unsigned(k.threads), 1, 1,
                                    unsigned(k.smem), s, args, nullptr);
  if (rc != 0) {
    *err = "cuLaunchKernel: " + DrvErr(api, rc);
    return false;
  }
  return true;
}

bool JitLaunchAccum(const JitKernel& k, unsigned ctas, unsigned rows, const float2* psi,
                    float2* lam, size_t row_stride, const void* terms, int n_terms,
                    const float* downstream, int n_ops, int accumulate,
                    unsigned long long n_tiles, cudaStream_t s, std::string* err) {
  Api& api = GetApi();
  void* args[] = {&psi, &lam, &row_stride, &terms, &n_terms, &downstream, &n_ops,
                  &accumulate, &n_tiles};
  const int rc = api.cuLaunchKernel(k.func, ctas, rows, 1, unsigned(k.threads), 1, 1,
                                    unsigned(k.smem), s, args, nullptr);
  if (rc != 0) {
    *err = "cuLaunchKernel: " + DrvErr(api, rc);
    return false;
  }
  return true;
}

}  // namespace tfqb
