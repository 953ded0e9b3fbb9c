// Float32 gate matrices from Cirq's closed forms, evaluated per batch row on
// the GPU (and on the host for unit tests).  Replaces the calls
// `qsim::Cirq::<Gate><float>::Create(...)` made per row by the reference
// (circuit_parser_qsim.cc:204-562) and the finite-difference gradient gates of
// adj_util.cc:175-302 (eps = 5e-3, adj_util.cc:32).
//
// Canonical float32 recipe (kept bit-identical with oracle/tfq_oracle.py so
// that finite-difference noise is common-mode):
//   ang = f32(pi32 * t); c,s = f32(cos/sin(f64(ang)*0.5));
//   g = f32(cos/sin(f64(ang)*(0.5 + f64(shift)))); entries are single float32
//   products / sums with NO fma contraction (hence the *_rn intrinsics).
// A 2-qubit matrix is over the operation's own qubit order (a,b):
// index = 2*x_a + x_b.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define TFQB_HD __host__ __device__ __forceinline__
#else
#define TFQB_HD inline
#endif

namespace tfqb {

struct cf { float re, im; };

TFQB_HD float fmul_(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
TFQB_HD float fadd_(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
TFQB_HD float fsub_(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b;
  return r;
#endif
}
TFQB_HD double dmul_(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}
TFQB_HD double dadd_(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}

TFQB_HD cf mk(float re, float im) { cf r; r.re = re; r.im = im; return r; }
TFQB_HD cf cneg(cf a) { return mk(-a.re, -a.im); }
TFQB_HD cf cmul_rn(cf a, cf b) {
  return mk(fsub_(fmul_(a.re, b.re), fmul_(a.im, b.im)),
            fadd_(fmul_(a.re, b.im), fmul_(a.im, b.re)));
}
TFQB_HD cf cs(double arg) { return mk(float(cos(arg)), float(sin(arg))); }

#define TFQB_PI32 3.14159265358979323846f
#define TFQB_IS2 0.70710678118654752440f

struct EigenParts { float c, s; cf g, g0; double ang; };

TFQB_HD EigenParts eigen_parts(float t, float shift) {
  EigenParts e;
  e.ang = double(fmul_(TFQB_PI32, t));
  const cf h = cs(dmul_(e.ang, 0.5));
  e.c = h.re;
  e.s = h.im;
  e.g = cs(dmul_(e.ang, dadd_(0.5, double(shift))));
  e.g0 = cs(dmul_(e.ang, double(shift)));
  return e;
}

TFQB_HD void zero16(cf* m, int dim) {
  for (int i = 0; i < dim * dim; ++i) m[i] = mk(0.f, 0.f);
}

// z-pow diagonal entries
TFQB_HD void zpow_diag(float t, float shift, cf* z0, cf* z1) {
  const double ang = double(fmul_(TFQB_PI32, t));
  *z0 = cs(dmul_(ang, double(shift)));
  *z1 = cs(dmul_(ang, dadd_(1.0, double(shift))));
}

TFQB_HD void xpow_entries(float t, float shift, cf* d, cf* o) {
  const EigenParts e = eigen_parts(t, shift);
  *d = mk(fmul_(e.c, e.g.re), fmul_(e.c, e.g.im));          // c g
  *o = mk(fmul_(e.s, e.g.im), -fmul_(e.s, e.g.re));         // -i s g
}

// ---- noise channels (next-row N2) ------------------------------------------
// The Kraus operator a trajectory applies at a 1-qubit channel, as a 2x2
// matrix, selected exactly as qsim's QuantumTrajectorySimulator does
// (lib/qtrajectory.h, restated in oracle/tfq_oracle.py run_trajectory): walk
// the operators with the cumulative probabilities (unitary operators) or
// lower bounds min eig(K'K) (non-unitary ones); if the uniform r is beyond
// them, walk the non-unitary operators again with their true probabilities
// <psi|K'K|psi> = kd0 * P0 + kd1 * P1 on the normalised state, falling back to
// the most probable one.  A non-unitary operator is returned divided by the
// square root of its probability, so the state stays normalised.
//   p[0] ChannelType (program.h), p[1..3] arguments, p[4] uniform r,
//   pop1 = population of |1> on the qubit (only read by non-unitary channels)
TFQB_HD void pauli_matrix(int k, cf* m) {      // 0 I, 1 X, 2 Y, 3 Z
  m[0] = m[3] = mk(1.f, 0.f);
  m[1] = m[2] = mk(0.f, 0.f);
  if (k == 1) { m[0] = m[3] = mk(0.f, 0.f); m[1] = m[2] = mk(1.f, 0.f); }
  if (k == 2) { m[0] = m[3] = mk(0.f, 0.f); m[1] = mk(0.f, -1.f); m[2] = mk(0.f, 1.f); }
  if (k == 3) { m[3] = mk(-1.f, 0.f); }
}
// <psi| K'K |psi> for a real 2x2 K = (k00, k01, k10, k11):
// K'K = diag(k00^2 + k10^2, k01^2 + k11^2) for every TFQ channel
TFQB_HD double kraus_prob(const float* K, double P0, double P1) {
  const double d0 = double(K[0]) * double(K[0]) + double(K[2]) * double(K[2]);
  const double d1 = double(K[1]) * double(K[1]) + double(K[3]) * double(K[3]);
  return d0 * P0 + d1 * P1;
}
TFQB_HD void channel_matrix(const float* p, float pop1, cf* m) {
  const int type = int(p[0]);
  const double a = double(p[1]), b = double(p[2]), c = double(p[3]);
  const double r = double(p[4]);
  const double P1 = double(pop1), P0 = 1.0 - P1;
  m[0] = m[3] = mk(1.f, 0.f);
  m[1] = m[2] = mk(0.f, 0.f);
  if (type <= 3) {               // mixtures of unitaries
    double pr[4] = {0.0, 0.0, 0.0, 0.0};
    int which[4] = {0, 1, 2, 3}, count = 4;
    if (type == 0) { pr[0] = 1.0 - a - b - c; pr[1] = a; pr[2] = b; pr[3] = c; }
    if (type == 1) { pr[0] = 1.0 - a; pr[1] = pr[2] = pr[3] = a / 3.0; }
    if (type == 2) { pr[0] = 1.0 - a; pr[1] = a; count = 2; }
    if (type == 3) { pr[0] = 1.0 - a; pr[1] = a; which[1] = 3; count = 2; }
    double cp = 0.0;
    for (int k = 0; k < count; ++k) {
      cp += pr[k];
      if (r < cp) { pauli_matrix(which[k], m); return; }
    }
    return;                      // r beyond the sum (round-off): no operator
  }
  // non-unitary: entries (k00, k01, k10, k11) real, lower bound, kd = diag(K'K)
  float K[4][4];
  double lb[4] = {0.0, 0.0, 0.0, 0.0};
  int count = 2;
  for (int k = 0; k < 4; ++k)
    for (int e = 0; e < 4; ++e) K[k][e] = 0.f;
  if (type == 4 || type == 5) {  // AD / PD (gamma = a)
    lb[0] = 1.0 - a;
    K[0][0] = 1.f; K[0][3] = float(sqrt(1.0 - a));
    if (type == 4) K[1][1] = float(sqrt(a)); else K[1][3] = float(sqrt(a));
  } else if (type == 6) {        // RST
    K[0][0] = 1.f;
    K[1][1] = 1.f;
  } else {                       // GAD (p = a, gamma = b)
    count = 4;
    lb[0] = a * (1.0 - b);
    lb[1] = (1.0 - a) * (1.0 - b);
    K[0][0] = float(sqrt(a)); K[0][3] = float(sqrt(a * (1.0 - b)));
    K[1][0] = float(sqrt((1.0 - a) * (1.0 - b))); K[1][3] = float(sqrt(1.0 - a));
    K[2][1] = float(sqrt(a * b));
    K[3][2] = float(sqrt((1.0 - a) * b));
  }
  int chosen = -1;
  double cp = 0.0;
  for (int k = 0; k < count; ++k) {
    cp += lb[k];
    if (r < cp) { chosen = k; break; }
  }
  if (chosen < 0) {
    int best = 0;
    double best_p = -1.0;
    for (int k = 0; k < count; ++k) {
      const double pk = kraus_prob(K[k], P0, P1);
      if (pk > best_p) { best = k; best_p = pk; }
      cp += pk - lb[k];
      if (r < cp || k == count - 1) { chosen = r < cp ? k : best; break; }
    }
  }
  const double pk = kraus_prob(K[chosen], P0, P1);
  const float scale = pk > 0.0 ? float(1.0 / sqrt(pk)) : 0.f;
  for (int e = 0; e < 4; ++e) m[e] = mk(fmul_(K[chosen][e], scale), 0.f);
}

// Gate kinds: keep in sync with program.h (GateKind).
// p[] are the resolved float parameters in reference order; the parameter
// `shift_idx` (unscaled symbol value) is displaced by `delta` before it is
// multiplied by its scalar: (v + delta) * scalar  (adj_util.cc:182-183).
TFQB_HD void gate_matrix(int kind, const float* p, int shift_idx, float delta,
                         cf* m) {
  float q[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < 5; ++i) q[i] = p[i];
  if (shift_idx >= 0) q[shift_idx] = fadd_(q[shift_idx], delta);
  switch (kind) {
    case 0:  // I
      zero16(m, 2);
      m[0] = m[3] = mk(1.f, 0.f);
      return;
    case 1:  // I2
      zero16(m, 4);
      m[0] = m[5] = m[10] = m[15] = mk(1.f, 0.f);
      return;
    case 2: {  // XP
      cf d, o;
      xpow_entries(fmul_(q[0], q[1]), q[2], &d, &o);
      m[0] = d; m[1] = o; m[2] = o; m[3] = d;
      return;
    }
    case 3: {  // YP
      const EigenParts e = eigen_parts(fmul_(q[0], q[1]), q[2]);
      const cf d = mk(fmul_(e.c, e.g.re), fmul_(e.c, e.g.im));
      const cf s = mk(fmul_(e.s, e.g.re), fmul_(e.s, e.g.im));   // s g
      m[0] = d; m[1] = cneg(s); m[2] = s; m[3] = d;
      return;
    }
    case 4: {  // ZP
      cf z0, z1;
      zpow_diag(fmul_(q[0], q[1]), q[2], &z0, &z1);
      m[0] = z0; m[1] = mk(0.f, 0.f); m[2] = mk(0.f, 0.f); m[3] = z1;
      return;
    }
    case 5: {  // HP: g (c I - i s H)
      const EigenParts e = eigen_parts(fmul_(q[0], q[1]), q[2]);
      const float ar = fmul_(fmul_(e.s, e.g.im), TFQB_IS2);
      const float ai = -fmul_(fmul_(e.s, e.g.re), TFQB_IS2);
      const float dr = fmul_(e.c, e.g.re), di = fmul_(e.c, e.g.im);
      m[0] = mk(fadd_(dr, ar), fadd_(di, ai));
      m[1] = mk(ar, ai);
      m[2] = mk(ar, ai);
      m[3] = mk(fsub_(dr, ar), fsub_(di, ai));
      return;
    }
    case 6: {  // XXP
      cf d, o;
      xpow_entries(fmul_(q[0], q[1]), q[2], &d, &o);
      zero16(m, 4);
      for (int i = 0; i < 4; ++i) { m[5 * i] = d; m[4 * i + 3 - i] = o; }
      return;
    }
    case 7: {  // YYP
      cf d, o;
      xpow_entries(fmul_(q[0], q[1]), q[2], &d, &o);
      zero16(m, 4);
      for (int i = 0; i < 4; ++i) m[5 * i] = d;
      m[3] = cneg(o); m[12] = cneg(o); m[6] = o; m[9] = o;
      return;
    }
    case 8: {  // ZZP
      cf z0, z1;
      zpow_diag(fmul_(q[0], q[1]), q[2], &z0, &z1);
      zero16(m, 4);
      m[0] = z0; m[5] = z1; m[10] = z1; m[15] = z0;
      return;
    }
    case 9: {  // CZP
      cf z0, z1;
      zpow_diag(fmul_(q[0], q[1]), q[2], &z0, &z1);
      zero16(m, 4);
      m[0] = z0; m[5] = z0; m[10] = z0; m[15] = z1;
      return;
    }
    case 10: {  // CNP (first qubit = control)
      const float t = fmul_(q[0], q[1]);
      cf d, o, z0, z1;
      xpow_entries(t, q[2], &d, &o);
      zpow_diag(t, q[2], &z0, &z1);
      zero16(m, 4);
      m[0] = z0; m[5] = z0; m[10] = d; m[11] = o; m[14] = o; m[15] = d;
      return;
    }
    case 11: {  // SP
      const float t = fmul_(q[0], q[1]);
      cf d, o, z0, z1;
      xpow_entries(t, q[2], &d, &o);
      zpow_diag(t, q[2], &z0, &z1);
      zero16(m, 4);
      m[0] = z0; m[15] = z0; m[5] = d; m[10] = d; m[6] = o; m[9] = o;
      return;
    }
    case 12: {  // ISP
      const EigenParts e = eigen_parts(fmul_(q[0], q[1]), q[2]);
      const cf d = mk(fmul_(e.c, e.g0.re), fmul_(e.c, e.g0.im));
      const cf o = mk(-fmul_(e.s, e.g0.im), fmul_(e.s, e.g0.re));  // i s g0
      zero16(m, 4);
      m[0] = e.g0; m[15] = e.g0; m[5] = d; m[10] = d; m[6] = o; m[9] = o;
      return;
    }
    case 13: {  // PXP: (pexp, pexp_s, exp, exp_s, gs)
      cf d, o;
      xpow_entries(fmul_(q[2], q[3]), q[4], &d, &o);
      const cf ph = cs(double(fmul_(TFQB_PI32, fmul_(q[0], q[1]))));
      m[0] = d;
      m[1] = cmul_rn(o, mk(ph.re, -ph.im));
      m[2] = cmul_rn(o, ph);
      m[3] = d;
      return;
    }
    case 14: {  // FSIM: (theta, theta_s, phi, phi_s)
      const cf t = cs(double(fmul_(q[0], q[1])));
      const cf f = cs(double(fmul_(q[2], q[3])));
      zero16(m, 4);
      m[0] = mk(1.f, 0.f);
      m[5] = mk(t.re, 0.f); m[10] = mk(t.re, 0.f);
      m[6] = mk(0.f, -t.im); m[9] = mk(0.f, -t.im);
      m[15] = mk(f.re, -f.im);
      return;
    }
    case 15: {  // PISP: (pexp, pexp_s, exp, exp_s)
      const double ang = double(fmul_(TFQB_PI32, fmul_(q[2], q[3])));
      const cf h = cs(dmul_(ang, 0.5));
      const cf f =
          cs(dmul_(double(fmul_(TFQB_PI32, fmul_(q[0], q[1]))), 2.0));
      zero16(m, 4);
      m[0] = mk(1.f, 0.f); m[15] = mk(1.f, 0.f);
      m[5] = mk(h.re, 0.f); m[10] = mk(h.re, 0.f);
      m[6] = mk(-fmul_(h.im, f.im), fmul_(h.im, f.re));   // i s f
      m[9] = mk(fmul_(h.im, f.im), fmul_(h.im, f.re));    // i s conj(f)
      return;
    }
    case 16:   // CH with no measured population: mixtures only (build_matrices
               // calls channel_matrix directly with the row's population)
      channel_matrix(q, 0.f, m);
      return;
    default:
      zero16(m, 4);
  }
}

#define TFQB_GRAD_EPS 5e-3f

// (G(p+eps) - G(p-eps)) * (0.5/eps) in float32 (adj_util.h:104-117).
TFQB_HD void gradient_matrix(int kind, const float* p, int shift_idx, int dim,
                             cf* m) {
  cf r[16];
  gate_matrix(kind, p, shift_idx, TFQB_GRAD_EPS, m);
  gate_matrix(kind, p, shift_idx, -TFQB_GRAD_EPS, r);
  const float scale = float(0.5 / double(TFQB_GRAD_EPS));
  for (int i = 0; i < dim * dim; ++i) {
    m[i].re = fmul_(fsub_(m[i].re, r[i].re), scale);
    m[i].im = fmul_(fsub_(m[i].im, r[i].im), scale);
  }
}

}  // namespace tfqb
