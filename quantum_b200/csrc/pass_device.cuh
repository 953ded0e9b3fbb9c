// Device helpers of the gate-pass kernels: packed complex arithmetic and the
// register-level gate / diagonal / sign / adjoint primitives.  Included by
// kernels.cu (inside namespace tfqb::{anonymous}) and, as text, by the
// run-time specialised pass kernels that jit.cc generates (NVRTC), so it must
// not include any host header.
#pragma once
#ifdef __CUDACC_RTC__
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
#endif

// Shared-memory slot of tile-local amplitude i (8-byte slots). Folding the
// higher nibbles onto the low one spreads the 16 lanes of a half-warp over
// the 16 distinct 8-byte columns whichever tile bits a round keeps in
// registers.
__device__ __forceinline__ uint32_t swz(uint32_t i) {
  return i ^ (((i >> 4) ^ (i >> 8)) & 15u);
}

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// acc + m * a
__device__ __forceinline__ float2 cfma(float2 m, float2 a, float2 acc) {
  acc.x = fmaf(m.x, a.x, acc.x);
  acc.x = fmaf(-m.y, a.y, acc.x);
  acc.y = fmaf(m.x, a.y, acc.y);
  acc.y = fmaf(m.y, a.x, acc.y);
  return acc;
}
// Re(conj(l) * p)
__device__ __forceinline__ float redot(float2 l, float2 p) {
  return fmaf(l.x, p.x, l.y * p.y);
}

// ------------------------------------------------------------------------
// Packed complex arithmetic.  sm_100a has 2-wide fp32 FMA (FFMA2 / FMUL2 in
// SASS): a complex multiply-accumulate acc += m*a is two of them,
//   acc = fma2((m.re, m.re), (a.re, a.im), acc)
//   acc = fma2((-m.im, m.im), (a.im, a.re), acc)
// so matrices sit in shared memory pre-expanded as float4
// (m.re, m.re, -m.im, m.im) and each amplitude is used with its swap
// s = (a.im, a.re).
// ------------------------------------------------------------------------
__device__ __forceinline__ float2 swp(float2 a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ float2 pmul(float4 m, float2 a, float2 s) {
  return __ffma2_rn(make_float2(m.z, m.w), s, __fmul2_rn(make_float2(m.x, m.y), a));
}
__device__ __forceinline__ float2 pmac(float4 m, float2 a, float2 s, float2 acc) {
  acc = __ffma2_rn(make_float2(m.x, m.y), a, acc);
  return __ffma2_rn(make_float2(m.z, m.w), s, acc);
}
// the plain complex value of an expanded entry
__device__ __forceinline__ float2 plain(float4 m) { return make_float2(m.x, m.w); }

// ------------------------------------------------------------------------
// register-level gate application. `a` holds 2^R amplitudes; bit j of the
// array index is register bit j of the round.
// ------------------------------------------------------------------------
template <int R, int J>
__device__ __forceinline__ void g1_packed(float2 (&a)[1 << R], const float4* __restrict__ sm) {
  const float4 m0 = sm[0], m1 = sm[1], m2 = sm[2], m3 = sm[3];
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    const float2 a0 = a[e], a1 = a[e | (1 << J)];
    const float2 s0 = swp(a0), s1 = swp(a1);
    a[e] = pmac(m1, a1, s1, pmul(m0, a0, s0));
    a[e | (1 << J)] = pmac(m3, a1, s1, pmul(m2, a0, s0));
  }
}

// ---- "phased real" 2x2 gates (plan.cc phased_real_flag), global phase dropped:
// U = diag(p0, p1) R or R diag(p0, p1) with R real; what is applied is
// diag(1, q) R resp. R diag(1, q), q = p1 conj(p0): 3 packed FMAs per
// amplitude instead of 4.  Only the kernels of jobs that cannot see a global
// phase use these (expectation, sampling, adjoint; never the state or
// inner-product ops).
// In-place rewrite of one staged complex matrix (4 expanded entries) into
//   sm[0] = (r00, r00, r01, r01), sm[1] = (r10, r10, r11, r11), sm[2] = q expanded
// mode 0: D R, 1: R D, 2: R times a phase (q = +-1 is folded into row 1),
// 3: the same for the dagger matrix of a fused adjoint step, whose dropped
// phase p0 moves into the gradient gate at sm[4..7] (psi' and lam then live in
// the frame conj(p0): <lam| p0 dG |psi'> is the reference's value)
// The rewrite runs once per CTA and matrix, in fp64 with ONE rounding per
// constant: these constants are shared by every amplitude, so a float32
// round-off here is a systematic error of the whole gate (it showed up as 3x
// the gradient error of the exact-gate kernels at 22 qubits,
// profiles/r02_float32_floor.jsonl).
// `lift` (pure rotations only: every real factor is a Y^t, so the real matrix
// is a proper rotation up to the signs absorbed by q and by the dropped global
// phase): the rotation [[c, -s], [s, c]], c >= 0, is applied as three shears
//   a0 -= t a1 ; a1 += s a0 ; a0 -= t a1,   t = s / (1 + c)  (|t| <= 1)
// i.e. 3 packed FMAs per PAIR instead of 4 (g1_real_lift & co. below);
//   sm[0] = (-t, -t, s, s), sm[2] = q as before
__device__ __forceinline__ void phased_real_setup(float4* sm, int mode, bool lift = false) {
  const int col_phased = mode == 1;
  double mx[4], my[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {        // m00 m01 m10 m11
    const float2 t = plain(sm[k]);
    mx[k] = double(t.x);
    my[k] = double(t.y);
  }
  // entries sharing the phase p0 / p1: rows (D R) or columns (R D)
  const int iu1 = col_phased ? 2 : 1, iv0 = col_phased ? 1 : 2;
  const double nu0 = mx[0] * mx[0] + my[0] * my[0], nu1 = mx[iu1] * mx[iu1] + my[iu1] * my[iu1];
  const double nv0 = mx[iv0] * mx[iv0] + my[iv0] * my[iv0], nv1 = mx[3] * mx[3] + my[3] * my[3];
  const int wu = nu0 >= nu1 ? 0 : iu1, wv = nv0 >= nv1 ? iv0 : 3;
  const double iu = rsqrt(fmax(nu0, nu1)), iv = rsqrt(fmax(nv0, nv1));
  const double p0x = mx[wu] * iu, p0y = my[wu] * iu, p1x = mx[wv] * iv, p1y = my[wv] * iv;
  const double ru0 = mx[0] * p0x + my[0] * p0y, ru1 = mx[iu1] * p0x + my[iu1] * p0y;
  const double rv0 = mx[iv0] * p1x + my[iv0] * p1y, rv1 = mx[3] * p1x + my[3] * p1y;
  const double qx = p1x * p0x + p1y * p0y, qy = p1y * p0x - p1x * p0y;
  // r00 r01 / r10 r11 in matrix positions
  const double r00 = ru0, r11 = rv1;
  const double r01 = col_phased ? rv0 : ru1;
  double r10 = col_phased ? ru1 : rv0, r11b = r11;
  if (mode >= 2 && qx < 0.0) {
    r10 = -r10;
    r11b = -r11b;
  }
  double q_x = qx, q_y = qy, gsign = 1.0;
  double lt = 0.0, ls = 0.0;
  if (lift) {
    double a00 = r00, a01 = r01, a10 = r10, a11 = r11b;
    if (a00 * a11 - a01 * a10 < 0.0) {   // independent row / column signs: move one into q
      if (col_phased) a01 = -a01; else a10 = -a10;
      a11 = -a11;
      q_x = -q_x;
      q_y = -q_y;
    }
    double c = 0.5 * (a00 + a11);
    ls = 0.5 * (a10 - a01);
    if (c < 0.0) {                       // -R(c, s) = R(-c, -s): a global sign
      c = -c;
      ls = -ls;
      gsign = -1.0;
    }
    const double nrm = rsqrt(c * c + ls * ls);
    c *= nrm;
    ls *= nrm;
    lt = ls / (1.0 + c);
  }
  if (mode == 3) {
    // (the dropped sign of a lifted dagger flips psi' but not the lambda the
    // gradient is taken with: it moves into the gradient gate as well)
#pragma unroll
    for (int k = 4; k < 8; ++k) {
      const float2 g = plain(sm[k]);
      const float dx = float(gsign * (p0x * double(g.x) - p0y * double(g.y)));
      const float dy = float(gsign * (p0x * double(g.y) + p0y * double(g.x)));
      sm[k] = make_float4(dx, dx, -dy, dy);
    }
  }
  if (lift) {
    sm[0] = make_float4(-float(lt), -float(lt), float(ls), float(ls));
  } else {
    sm[0] = make_float4(float(r00), float(r00), float(r01), float(r01));
    sm[1] = make_float4(float(r10), float(r10), float(r11b), float(r11b));
  }
  sm[2] = make_float4(float(q_x), float(q_x), -float(q_y), float(q_y));
}
// X^t = p0 [[c, -i s], [-i s, c]]: rewrite the staged dagger / gate matrix into
//   sm[0] = (r00, r00, -x01, x01), sm[1] = (-x10, x10, r11, r11)
// (r: real parts, x: imaginary parts after dividing by p0); with_grad: the
// gradient gate at sm[4..7] takes the dropped phase (adjoint steps)
// `lift`: [[c, i x], [i x, c]] (c >= 0 after dropping a global sign) as three
// shears a0 += i t a1 ; a1 += i x a0 ; a0 += i t a1, t = x / (1 + c):
//   sm[0] = (-t, t, -x, x)
__device__ __forceinline__ void phased_ximag_setup(float4* sm, int with_grad, bool lift = false) {
  const float2 f00 = plain(sm[0]), f01 = plain(sm[1]), f10 = plain(sm[2]), f11 = plain(sm[3]);
  const double m00x = f00.x, m00y = f00.y, m01x = f01.x, m01y = f01.y;
  const double m10x = f10.x, m10y = f10.y, m11x = f11.x, m11y = f11.y;
  const double n0 = m00x * m00x + m00y * m00y, n1 = m01x * m01x + m01y * m01y;
  // the diagonal entry is real after the division, the off-diagonal imaginary
  const double wx = n0 >= n1 ? m00x : -m01y, wy = n0 >= n1 ? m00y : m01x;    // i * m01
  const double inv = rsqrt(fmax(n0, n1));
  const double p0x = wx * inv, p0y = wy * inv;
  const double r00 = m00x * p0x + m00y * p0y, r11 = m11x * p0x + m11y * p0y;
  const double x01 = m01y * p0x - m01x * p0y, x10 = m10y * p0x - m10x * p0y;
  double gsign = 1.0, lt = 0.0, lx = 0.0;
  if (lift) {
    double c = 0.5 * (r00 + r11);
    lx = 0.5 * (x01 + x10);
    if (c < 0.0) {
      c = -c;
      lx = -lx;
      gsign = -1.0;
    }
    const double nrm = rsqrt(c * c + lx * lx);
    c *= nrm;
    lx *= nrm;
    lt = lx / (1.0 + c);
  }
  if (with_grad) {
#pragma unroll
    for (int k = 4; k < 8; ++k) {
      const float2 g = plain(sm[k]);
      const float dx = float(gsign * (p0x * double(g.x) - p0y * double(g.y)));
      const float dy = float(gsign * (p0x * double(g.y) + p0y * double(g.x)));
      sm[k] = make_float4(dx, dx, -dy, dy);
    }
  }
  if (lift) {
    sm[0] = make_float4(-float(lt), float(lt), -float(lx), float(lx));
    return;
  }
  sm[0] = make_float4(float(r00), float(r00), -float(x01), float(x01));
  sm[1] = make_float4(-float(x10), float(x10), float(r11), float(r11));
}
// [[r00, i x01], [i x10, r11]] on register bit J: 2 packed FMAs per amplitude
template <int R, int J>
__device__ __forceinline__ void g1_ximag(float2 (&a)[1 << R], const float4* __restrict__ sm) {
  const float4 r0 = sm[0], r1 = sm[1];
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    const float2 a0 = a[e], a1 = a[e | (1 << J)];
    a[e] = __ffma2_rn(make_float2(r0.z, r0.w), swp(a1), __fmul2_rn(make_float2(r0.x, r0.y), a0));
    a[e | (1 << J)] =
        __ffma2_rn(make_float2(r1.x, r1.y), swp(a0), __fmul2_rn(make_float2(r1.z, r1.w), a1));
  }
}
// R alone (phased_real_setup mode 2)
template <int R, int J>
__device__ __forceinline__ void g1_real(float2 (&a)[1 << R], const float4* __restrict__ sm) {
  const float4 r0 = sm[0], r1 = sm[1];
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    const float2 a0 = a[e], a1 = a[e | (1 << J)];
    a[e] = __ffma2_rn(make_float2(r0.z, r0.w), a1, __fmul2_rn(make_float2(r0.x, r0.y), a0));
    a[e | (1 << J)] =
        __ffma2_rn(make_float2(r1.z, r1.w), a1, __fmul2_rn(make_float2(r1.x, r1.y), a0));
  }
}
// diag(1, q) R
template <int R, int J>
__device__ __forceinline__ void g1_rowreal(float2 (&a)[1 << R], const float4* __restrict__ sm) {
  const float4 r0 = sm[0], r1 = sm[1], q = sm[2];
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    const float2 a0 = a[e], a1 = a[e | (1 << J)];
    a[e] = __ffma2_rn(make_float2(r0.z, r0.w), a1, __fmul2_rn(make_float2(r0.x, r0.y), a0));
    const float2 t =
        __ffma2_rn(make_float2(r1.z, r1.w), a1, __fmul2_rn(make_float2(r1.x, r1.y), a0));
    a[e | (1 << J)] = pmul(q, t, swp(t));
  }
}
// R diag(1, q)
template <int R, int J>
__device__ __forceinline__ void g1_colreal(float2 (&a)[1 << R], const float4* __restrict__ sm) {
  const float4 r0 = sm[0], r1 = sm[1], q = sm[2];
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    const float2 a0 = a[e], a1 = a[e | (1 << J)];
    const float2 b1 = pmul(q, a1, swp(a1));
    a[e] = __ffma2_rn(make_float2(r0.z, r0.w), b1, __fmul2_rn(make_float2(r0.x, r0.y), a0));
    a[e | (1 << J)] =
        __ffma2_rn(make_float2(r1.z, r1.w), b1, __fmul2_rn(make_float2(r1.x, r1.y), a0));
  }
}

// ---- the same four with the rotation applied as three shears (setup `lift`):
// 1.5 packed FMAs per amplitude for the rotation instead of 2
template <int R, int J>
__device__ __forceinline__ void g1_real_lift(float2 (&a)[1 << R], const float4* __restrict__ sm) {
  const float4 k = sm[0];
  const float2 mt = make_float2(k.x, k.y), s = make_float2(k.z, k.w);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    float2 a0 = a[e], a1 = a[e | (1 << J)];
    a0 = __ffma2_rn(mt, a1, a0);
    a1 = __ffma2_rn(s, a0, a1);
    a[e] = __ffma2_rn(mt, a1, a0);
    a[e | (1 << J)] = a1;
  }
}
template <int R, int J>
__device__ __forceinline__ void g1_rowreal_lift(float2 (&a)[1 << R], const float4* __restrict__ sm) {
  const float4 k = sm[0], q = sm[2];
  const float2 mt = make_float2(k.x, k.y), s = make_float2(k.z, k.w);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    float2 a0 = a[e], a1 = a[e | (1 << J)];
    a0 = __ffma2_rn(mt, a1, a0);
    a1 = __ffma2_rn(s, a0, a1);
    a[e] = __ffma2_rn(mt, a1, a0);
    a[e | (1 << J)] = pmul(q, a1, swp(a1));
  }
}
template <int R, int J>
__device__ __forceinline__ void g1_colreal_lift(float2 (&a)[1 << R], const float4* __restrict__ sm) {
  const float4 k = sm[0], q = sm[2];
  const float2 mt = make_float2(k.x, k.y), s = make_float2(k.z, k.w);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    float2 a0 = a[e];
    const float2 b1 = a[e | (1 << J)];
    float2 a1 = pmul(q, b1, swp(b1));
    a0 = __ffma2_rn(mt, a1, a0);
    a1 = __ffma2_rn(s, a0, a1);
    a[e] = __ffma2_rn(mt, a1, a0);
    a[e | (1 << J)] = a1;
  }
}
// [[c, i x], [i x, c]]: i t a = (-t a.im, t a.re), the swap is free in FFMA2
template <int R, int J>
__device__ __forceinline__ void g1_ximag_lift(float2 (&a)[1 << R], const float4* __restrict__ sm) {
  const float4 k = sm[0];
  const float2 it = make_float2(k.x, k.y), ix = make_float2(k.z, k.w);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    float2 a0 = a[e], a1 = a[e | (1 << J)];
    a0 = __ffma2_rn(it, swp(a1), a0);
    a1 = __ffma2_rn(ix, swp(a0), a1);
    a[e] = __ffma2_rn(it, swp(a1), a0);
    a[e | (1 << J)] = a1;
  }
}

// dense 4x4, matrix rows streamed from shared memory (B0 = register of the
// matrix msb, B0 > B1)
template <int R, int B0, int B1>
__device__ __forceinline__ void g2_packed(float2 (&a)[1 << R], const float4* __restrict__ sm) {
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & ((1 << B0) | (1 << B1))) continue;
    const int i1 = e | (1 << B1), i2 = e | (1 << B0), i3 = i1 | i2;
    const float2 a0 = a[e], a1 = a[i1], a2 = a[i2], a3 = a[i3];
    const float2 s0 = swp(a0), s1 = swp(a1), s2 = swp(a2), s3 = swp(a3);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float4 c0 = sm[4 * r], c1 = sm[4 * r + 1], c2 = sm[4 * r + 2],
                   c3 = sm[4 * r + 3];
      const float2 v = pmac(c3, a3, s3, pmac(c2, a2, s2, pmac(c1, a1, s1, pmul(c0, a0, s0))));
      a[r == 0 ? e : r == 1 ? i1 : r == 2 ? i2 : i3] = v;
    }
  }
}

// 2 Re<l| D |a> / 2 over this thread's amplitudes, D dense 2x2
template <int R, int J>
__device__ __forceinline__ float grad1_packed(const float2 (&a)[1 << R],
                                              const float2 (&l)[1 << R],
                                              const float4* __restrict__ sm) {
  const float4 m0 = sm[0], m1 = sm[1], m2 = sm[2], m3 = sm[3];
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    const float2 a0 = a[e], a1 = a[e | (1 << J)];
    const float2 s0 = swp(a0), s1 = swp(a1);
    acc = __ffma2_rn(l[e], pmac(m1, a1, s1, pmul(m0, a0, s0)), acc);
    acc = __ffma2_rn(l[e | (1 << J)], pmac(m3, a1, s1, pmul(m2, a0, s0)), acc);
  }
  return acc.x + acc.y;
}

template <int R, int B0, int B1>
__device__ __forceinline__ float grad2_packed(const float2 (&a)[1 << R],
                                              const float2 (&l)[1 << R],
                                              const float4* __restrict__ sm) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & ((1 << B0) | (1 << B1))) continue;
    const int i1 = e | (1 << B1), i2 = e | (1 << B0), i3 = i1 | i2;
    const float2 a0 = a[e], a1 = a[i1], a2 = a[i2], a3 = a[i3];
    const float2 s0 = swp(a0), s1 = swp(a1), s2 = swp(a2), s3 = swp(a3);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float4 c0 = sm[4 * r], c1 = sm[4 * r + 1], c2 = sm[4 * r + 2],
                   c3 = sm[4 * r + 3];
      const float2 v = pmac(c3, a3, s3, pmac(c2, a2, s2, pmac(c1, a1, s1, pmul(c0, a0, s0))));
      acc = __ffma2_rn(l[r == 0 ? e : r == 1 ? i1 : r == 2 ? i2 : i3], v, acc);
    }
  }
  return acc.x + acc.y;
}

// ---- diagonal ops ---------------------------------------------------------
template <int R>
__device__ __forceinline__ void scale_all(float2 (&a)[1 << R], float4 f) {
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) a[e] = pmul(f, a[e], swp(a[e]));
}
template <int R>
__device__ __forceinline__ void scale_all_c(float2 (&a)[1 << R], float2 f) {
  scale_all<R>(a, make_float4(f.x, f.x, -f.y, f.y));
}

// one selector bit is register bit J: entries f0 (bit clear) / f1 (bit set)
template <int R, int J>
__device__ __forceinline__ void diag1(float2 (&a)[1 << R], float4 f0, float4 f1,
                                      bool do0, bool do1) {
  if (do0) {
#pragma unroll
    for (int e = 0; e < (1 << R); ++e)
      if (!(e & (1 << J))) a[e] = pmul(f0, a[e], swp(a[e]));
  }
  if (do1) {
#pragma unroll
    for (int e = 0; e < (1 << R); ++e)
      if (e & (1 << J)) a[e] = pmul(f1, a[e], swp(a[e]));
  }
}

// both selector bits are register bits: JH = register of the selector msb
template <int R, int JH, int JL>
__device__ __forceinline__ void diag2(float2 (&a)[1 << R], const float4* __restrict__ sm,
                                      uint32_t skip) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    if ((skip >> s) & 1u) continue;     // uniform
    const float4 d = sm[s];
#pragma unroll
    for (int e = 0; e < (1 << R); ++e)
      if ((((e >> JH) & 1) * 2 + ((e >> JL) & 1)) == s) a[e] = pmul(d, a[e], swp(a[e]));
  }
}

// gradient of a diagonal gate: sum_e Re(conj(l_e) * d[sel(e)] * a_e)
template <int R>
__device__ __forceinline__ float gdiag0(const float2 (&a)[1 << R], const float2 (&l)[1 << R],
                                        float4 f) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e)
    acc = __ffma2_rn(l[e], pmul(f, a[e], swp(a[e])), acc);
  return acc.x + acc.y;
}
template <int R, int J>
__device__ __forceinline__ float gdiag1(const float2 (&a)[1 << R], const float2 (&l)[1 << R],
                                        float4 f0, float4 f1) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e)
    acc = __ffma2_rn(l[e], pmul((e & (1 << J)) ? f1 : f0, a[e], swp(a[e])), acc);
  return acc.x + acc.y;
}
template <int R, int JH, int JL>
__device__ __forceinline__ float gdiag2(const float2 (&a)[1 << R], const float2 (&l)[1 << R],
                                        const float4* __restrict__ sm) {
  const float4 d0 = sm[0], d1 = sm[1], d2 = sm[2], d3 = sm[3];
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    const int s = ((e >> JH) & 1) * 2 + ((e >> JL) & 1);
    acc = __ffma2_rn(l[e], pmul(s == 0 ? d0 : s == 1 ? d1 : s == 2 ? d2 : d3, a[e], swp(a[e])), acc);
  }
  return acc.x + acc.y;
}

// ---- sign ops: literal +-1 diagonals (CZ, Z, ZZ at exponent 1) ---------------
// TFQB_SIGN_XOR (defined by the generated adjoint passes): flip the sign bit on
// the integer pipe.  There the negations cannot fold into FMA operands (psi
// and lambda are both flipped and stored) and as FADDs they were 12% of the
// instructions of a kernel whose limiter is the FMA pipe.
__device__ __forceinline__ float2 cneg2(float2 a) {
#ifdef TFQB_SIGN_XOR
  uint32_t x = __float_as_uint(a.x), y = __float_as_uint(a.y);
  asm("xor.b32 %0, %0, 0x80000000;" : "+r"(x));
  asm("xor.b32 %0, %0, 0x80000000;" : "+r"(y));
  return make_float2(__uint_as_float(x), __uint_as_float(y));
#else
  return make_float2(-a.x, -a.y);
#endif
}
template <int R>
__device__ __forceinline__ void sign_all(float2 (&a)[1 << R]) {
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) a[e] = cneg2(a[e]);
}
template <int R, int J>
__device__ __forceinline__ void sign1(float2 (&a)[1 << R], bool n0, bool n1) {
  if (n0) {
#pragma unroll
    for (int e = 0; e < (1 << R); ++e)
      if (!(e & (1 << J))) a[e] = cneg2(a[e]);
  }
  if (n1) {
#pragma unroll
    for (int e = 0; e < (1 << R); ++e)
      if (e & (1 << J)) a[e] = cneg2(a[e]);
  }
}
template <int R, int JH, int JL>
__device__ __forceinline__ void sign2(float2 (&a)[1 << R], uint32_t mask) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    if (!((mask >> s) & 1u)) continue;     // uniform
#pragma unroll
    for (int e = 0; e < (1 << R); ++e)
      if ((((e >> JH) & 1) * 2 + ((e >> JL) & 1)) == s) a[e] = cneg2(a[e]);
  }
}

// ---- fused adjoint step of one parameterised gate ---------------------------
// sm holds two matrices back to back: G' (dagger) then the gradient gate D.
//   psi <- G' psi ; acc += Re(conj(lam) . D psi) ; lam <- G' lam
template <int R, int J>
__device__ __forceinline__ float adj1_packed(float2 (&a)[1 << R], float2 (&l)[1 << R],
                                             const float4* __restrict__ sm) {
  const float4 m0 = sm[0], m1 = sm[1], m2 = sm[2], m3 = sm[3];
  const float4 d0 = sm[4], d1 = sm[5], d2 = sm[6], d3 = sm[7];
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    const int f = e | (1 << J);
    const float2 a0 = a[e], a1 = a[f];
    const float2 s0 = swp(a0), s1 = swp(a1);
    const float2 n0 = pmac(m1, a1, s1, pmul(m0, a0, s0));
    const float2 n1 = pmac(m3, a1, s1, pmul(m2, a0, s0));
    const float2 t0 = swp(n0), t1 = swp(n1);
    acc = __ffma2_rn(l[e], pmac(d1, n1, t1, pmul(d0, n0, t0)), acc);
    acc = __ffma2_rn(l[f], pmac(d3, n1, t1, pmul(d2, n0, t0)), acc);
    a[e] = n0;
    a[f] = n1;
    const float2 l0 = l[e], l1 = l[f];
    const float2 u0 = swp(l0), u1 = swp(l1);
    l[e] = pmac(m1, l1, u1, pmul(m0, l0, u0));
    l[f] = pmac(m3, l1, u1, pmul(m2, l0, u0));
  }
  return acc.x + acc.y;
}

// the same with a real dagger matrix (phased_real_setup mode 3): 18 packed
// FMAs per pair instead of 26
template <int R, int J>
__device__ __forceinline__ float adj1_real(float2 (&a)[1 << R], float2 (&l)[1 << R],
                                           const float4* __restrict__ sm) {
  const float4 r0 = sm[0], r1 = sm[1];
  const float4 d0 = sm[4], d1 = sm[5], d2 = sm[6], d3 = sm[7];
  const float2 r00 = make_float2(r0.x, r0.y), r01 = make_float2(r0.z, r0.w);
  const float2 r10 = make_float2(r1.x, r1.y), r11 = make_float2(r1.z, r1.w);
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    const int f = e | (1 << J);
    const float2 a0 = a[e], a1 = a[f];
    const float2 n0 = __ffma2_rn(r01, a1, __fmul2_rn(r00, a0));
    const float2 n1 = __ffma2_rn(r11, a1, __fmul2_rn(r10, a0));
    const float2 t0 = swp(n0), t1 = swp(n1);
    acc = __ffma2_rn(l[e], pmac(d1, n1, t1, pmul(d0, n0, t0)), acc);
    acc = __ffma2_rn(l[f], pmac(d3, n1, t1, pmul(d2, n0, t0)), acc);
    a[e] = n0;
    a[f] = n1;
    const float2 l0 = l[e], l1 = l[f];
    l[e] = __ffma2_rn(r01, l1, __fmul2_rn(r00, l0));
    l[f] = __ffma2_rn(r11, l1, __fmul2_rn(r10, l0));
  }
  return acc.x + acc.y;
}

// and with the dagger of X^t (phased_ximag_setup)
template <int R, int J>
__device__ __forceinline__ float adj1_ximag(float2 (&a)[1 << R], float2 (&l)[1 << R],
                                            const float4* __restrict__ sm) {
  const float4 r0 = sm[0], r1 = sm[1];
  const float4 d0 = sm[4], d1 = sm[5], d2 = sm[6], d3 = sm[7];
  const float2 c00 = make_float2(r0.x, r0.y), x01 = make_float2(r0.z, r0.w);
  const float2 x10 = make_float2(r1.x, r1.y), c11 = make_float2(r1.z, r1.w);
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    const int f = e | (1 << J);
    const float2 a0 = a[e], a1 = a[f];
    const float2 n0 = __ffma2_rn(x01, swp(a1), __fmul2_rn(c00, a0));
    const float2 n1 = __ffma2_rn(x10, swp(a0), __fmul2_rn(c11, a1));
    const float2 t0 = swp(n0), t1 = swp(n1);
    acc = __ffma2_rn(l[e], pmac(d1, n1, t1, pmul(d0, n0, t0)), acc);
    acc = __ffma2_rn(l[f], pmac(d3, n1, t1, pmul(d2, n0, t0)), acc);
    a[e] = n0;
    a[f] = n1;
    const float2 l0 = l[e], l1 = l[f];
    l[e] = __ffma2_rn(x01, swp(l1), __fmul2_rn(c00, l0));
    l[f] = __ffma2_rn(x10, swp(l0), __fmul2_rn(c11, l1));
  }
  return acc.x + acc.y;
}

// adj1_real / adj1_ximag with the dagger applied as three shears (setup `lift`):
// 16 packed FMAs per pair instead of 18
template <int R, int J>
__device__ __forceinline__ float adj1_real_lift(float2 (&a)[1 << R], float2 (&l)[1 << R],
                                                const float4* __restrict__ sm) {
  const float4 k = sm[0];
  const float4 d0 = sm[4], d1 = sm[5], d2 = sm[6], d3 = sm[7];
  const float2 mt = make_float2(k.x, k.y), s = make_float2(k.z, k.w);
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    const int f = e | (1 << J);
    float2 n0 = a[e], n1 = a[f];
    n0 = __ffma2_rn(mt, n1, n0);
    n1 = __ffma2_rn(s, n0, n1);
    n0 = __ffma2_rn(mt, n1, n0);
    const float2 t0 = swp(n0), t1 = swp(n1);
    acc = __ffma2_rn(l[e], pmac(d1, n1, t1, pmul(d0, n0, t0)), acc);
    acc = __ffma2_rn(l[f], pmac(d3, n1, t1, pmul(d2, n0, t0)), acc);
    a[e] = n0;
    a[f] = n1;
    float2 l0 = l[e], l1 = l[f];
    l0 = __ffma2_rn(mt, l1, l0);
    l1 = __ffma2_rn(s, l0, l1);
    l[e] = __ffma2_rn(mt, l1, l0);
    l[f] = l1;
  }
  return acc.x + acc.y;
}
template <int R, int J>
__device__ __forceinline__ float adj1_ximag_lift(float2 (&a)[1 << R], float2 (&l)[1 << R],
                                                 const float4* __restrict__ sm) {
  const float4 k = sm[0];
  const float4 d0 = sm[4], d1 = sm[5], d2 = sm[6], d3 = sm[7];
  const float2 it = make_float2(k.x, k.y), ix = make_float2(k.z, k.w);
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    const int f = e | (1 << J);
    float2 n0 = a[e], n1 = a[f];
    n0 = __ffma2_rn(it, swp(n1), n0);
    n1 = __ffma2_rn(ix, swp(n0), n1);
    n0 = __ffma2_rn(it, swp(n1), n0);
    const float2 t0 = swp(n0), t1 = swp(n1);
    acc = __ffma2_rn(l[e], pmac(d1, n1, t1, pmul(d0, n0, t0)), acc);
    acc = __ffma2_rn(l[f], pmac(d3, n1, t1, pmul(d2, n0, t0)), acc);
    a[e] = n0;
    a[f] = n1;
    float2 l0 = l[e], l1 = l[f];
    l0 = __ffma2_rn(it, swp(l1), l0);
    l1 = __ffma2_rn(ix, swp(l0), l1);
    l[e] = __ffma2_rn(it, swp(l1), l0);
    l[f] = l1;
  }
  return acc.x + acc.y;
}

template <int R, int B0, int B1>
__device__ __forceinline__ float adj2_packed(float2 (&a)[1 << R], float2 (&l)[1 << R],
                                             const float4* __restrict__ sm) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & ((1 << B0) | (1 << B1))) continue;
    const int idx[4] = {e, e | (1 << B1), e | (1 << B0), e | (1 << B0) | (1 << B1)};
    float2 x[4], sx[4], n[4], sn[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { x[k] = a[idx[k]]; sx[k] = swp(x[k]); }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      n[r] = pmac(sm[4 * r + 3], x[3], sx[3],
                  pmac(sm[4 * r + 2], x[2], sx[2],
                       pmac(sm[4 * r + 1], x[1], sx[1], pmul(sm[4 * r], x[0], sx[0]))));
      sn[r] = swp(n[r]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float2 p = pmac(sm[16 + 4 * r + 3], n[3], sn[3],
                            pmac(sm[16 + 4 * r + 2], n[2], sn[2],
                                 pmac(sm[16 + 4 * r + 1], n[1], sn[1],
                                      pmul(sm[16 + 4 * r], n[0], sn[0]))));
      acc = __ffma2_rn(l[idx[r]], p, acc);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { a[idx[k]] = n[k]; x[k] = l[idx[k]]; sx[k] = swp(x[k]); }
#pragma unroll
    for (int r = 0; r < 4; ++r)
      l[idx[r]] = pmac(sm[4 * r + 3], x[3], sx[3],
                       pmac(sm[4 * r + 2], x[2], sx[2],
                            pmac(sm[4 * r + 1], x[1], sx[1], pmul(sm[4 * r], x[0], sx[0]))));
  }
  return acc.x + acc.y;
}

// diagonal: f = dagger entry, g = gradient entry of this amplitude
__device__ __forceinline__ void adjd_elem(float2& a, float2& l, float4 f, float4 g,
                                          float2& acc) {
  const float2 n = pmul(f, a, swp(a));
  acc = __ffma2_rn(l, pmul(g, n, swp(n)), acc);
  a = n;
  l = pmul(f, l, swp(l));
}
template <int R>
__device__ __forceinline__ float adjd0(float2 (&a)[1 << R], float2 (&l)[1 << R],
                                       float4 f, float4 g) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) adjd_elem(a[e], l[e], f, g, acc);
  return acc.x + acc.y;
}
template <int R, int J>
__device__ __forceinline__ float adjd1(float2 (&a)[1 << R], float2 (&l)[1 << R],
                                       float4 f0, float4 f1, float4 g0, float4 g1) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e)
    adjd_elem(a[e], l[e], (e & (1 << J)) ? f1 : f0, (e & (1 << J)) ? g1 : g0, acc);
  return acc.x + acc.y;
}
template <int R, int JH, int JL>
__device__ __forceinline__ float adjd2(float2 (&a)[1 << R], float2 (&l)[1 << R],
                                       const float4* __restrict__ sm) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const float4 f = sm[s], g = sm[4 + s];
#pragma unroll
    for (int e = 0; e < (1 << R); ++e)
      if ((((e >> JH) & 1) * 2 + ((e >> JL) & 1)) == s) adjd_elem(a[e], l[e], f, g, acc);
  }
  return acc.x + acc.y;
}

// ---- runs of diagonal fused adjoint steps (specialised kernels) -----------------
// A diagonal step multiplies psi_e and lam_e by the SAME unit-modulus entry, so
// c_e = conj(lam_e) psi_e does not change along a run of such steps: it is
// computed once, every step's gradient 2 Re<lam| dD D' |psi> is Re(h . sum c_e)
// over the elements of each entry (h = gradient entry times dagger entry), and
// thread-constant entries are accumulated into one phase applied at the end.
template <int R>
__device__ __forceinline__ void conj_products(const float2 (&a)[1 << R],
                                              const float2 (&l)[1 << R],
                                              float2 (&c)[1 << R]) {
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    c[e].x = fmaf(l[e].x, a[e].x, l[e].y * a[e].y);
    c[e].y = fmaf(l[e].x, a[e].y, -(l[e].y * a[e].x));
  }
}
template <int R>
__device__ __forceinline__ float2 csum_all(const float2 (&c)[1 << R]) {
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    s.x += c[e].x;
    s.y += c[e].y;
  }
  return s;
}
template <int R, int J>
__device__ __forceinline__ void csum_bit(const float2 (&c)[1 << R], float2& s0, float2& s1) {
  s0 = make_float2(0.f, 0.f);
  s1 = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) {
      s1.x += c[e].x;
      s1.y += c[e].y;
    } else {
      s0.x += c[e].x;
      s0.y += c[e].y;
    }
  }
}
// sums over the elements of each entry of a two-register-bit diagonal
// (entry index = 2 * bit JH + bit JL, as diag2)
template <int R, int JH, int JL>
__device__ __forceinline__ void csum_2bit(const float2 (&c)[1 << R], float2 (&s)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) s[k] = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    const int k = ((e >> JH) & 1) * 2 + ((e >> JL) & 1);
    s[k].x += c[e].x;
    s[k].y += c[e].y;
  }
}
// Re(g f s), associated as g (f s): the product g f would be the same rounded
// constant in every thread -- a systematic error of the whole gate's gradient --
// while f s rounds differently per thread and averages out
__device__ __forceinline__ float re_hs(float4 g, float4 f, float2 s) {
  const float2 t = cmulf(plain(f), s);
  const float2 gp = plain(g);
  return fmaf(gp.x, t.x, -(gp.y * t.y));
}

// ---- PauliSum expectation primitives (K1) -------------------------------------
// sum over the pairs (e, k = e ^ XR) held by this thread of
// (-1)^{parity(k & zreg)} * conj(a_e) * a_k  -> (real part, imaginary part)
template <int XR>
__device__ __forceinline__ float2 xterm_pairs(const float2 (&a)[16], uint32_t sign16) {
  constexpr int LSB = XR & (-XR);
  float accr = 0.f, acci = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    if (e & LSB) continue;
    const int k = e ^ XR;
    const uint32_t flip = ((sign16 >> k) & 1u) << 31;
    const float ux = __uint_as_float(__float_as_uint(a[e].x) ^ flip);
    const float uy = __uint_as_float(__float_as_uint(a[e].y) ^ flip);
    const float2 v = a[k];
    accr = fmaf(ux, v.x, accr);
    accr = fmaf(uy, v.y, accr);
    acci = fmaf(ux, v.y, acci);
    acci = fmaf(-uy, v.x, acci);
  }
  return make_float2(accr, acci);
}

// Same sum with the sign pattern known at compile time (ZS = 0: no register z
// bit; ZS = 1 + j: the only register z bit is j) and only the part the term
// needs (IM: imaginary, else real): 1 FFMA2 per pair, negations folded into
// the operands.
template <int XR, int ZS, bool IM>
__device__ __forceinline__ float xterm_fixed(const float2 (&a)[16]) {
  constexpr int LSB = XR & (-XR);
  // the opaque zero pins the FMA chains inside their switch case (otherwise
  // the compiler evaluates every case speculatively and selects afterwards)
  float z = 0.f;
  asm volatile("" : "+f"(z));
  // one packed FMA per pair: lanes (ux vx, uy vy) for the real part,
  // (ux vy, uy vx) for the imaginary one (the swap is free in FFMA2)
  float2 acc[2] = {make_float2(z, z), make_float2(z, z)};
  int n = 0;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    if (e & LSB) continue;
    const int k = e ^ XR;
    const bool neg = ZS > 0 && ((k >> (ZS > 0 ? ZS - 1 : 0)) & 1);
    const float2 u = neg ? make_float2(-a[e].x, -a[e].y) : a[e];
    acc[n & 1] = __ffma2_rn(u, IM ? swp(a[k]) : a[k], acc[n & 1]);
    ++n;
  }
  const float sx = acc[0].x + acc[1].x, sy = acc[0].y + acc[1].y;
  return IM ? sx - sy : sx + sy;
}

// NB butterfly stages of the Walsh-Hadamard transform on bits [lvl, lvl+NB)
template <int NB>
__device__ __forceinline__ void wht_level(float* __restrict__ s_p, uint32_t tile_size,
                                          int lvl, int tid, int nthr) {
  const uint32_t groups = tile_size >> NB;
  const uint32_t lo = (1u << lvl) - 1u;
  uint32_t so[NB];       // swz is GF(2)-linear: swz(b|off) = swz(b)^swz(off)
#pragma unroll
  for (int j = 0; j < NB; ++j) so[j] = swz(1u << (lvl + j));
  for (uint32_t gi = tid; gi < groups; gi += nthr) {
    const uint32_t sb = swz(((gi & ~lo) << NB) | (gi & lo));
    float w[1 << NB];
#pragma unroll
    for (int e = 0; e < (1 << NB); ++e) {
      uint32_t x = sb;
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (e & (1 << j)) x ^= so[j];
      w[e] = s_p[x];
    }
    // level 0 on scalars, the others two butterflies per packed add (FADD2)
#pragma unroll
    for (int e = 0; e < (1 << NB); e += 2) {
      const float x = w[e], y = w[e + 1];
      w[e] = x + y;
      w[e + 1] = x - y;
    }
#pragma unroll
    for (int j = 1; j < NB; ++j) {
#pragma unroll
      for (int e = 0; e < (1 << NB); e += 2) {
        if (e & (1 << j)) continue;
        const float2 x = make_float2(w[e], w[e + 1]);
        const float2 y = make_float2(w[e | (1 << j)], w[(e | (1 << j)) + 1]);
        const float2 u = __fadd2_rn(x, y), v = __fadd2_rn(x, make_float2(-y.x, -y.y));
        w[e] = u.x;
        w[e + 1] = u.y;
        w[e | (1 << j)] = v.x;
        w[(e | (1 << j)) + 1] = v.y;
      }
    }
#pragma unroll
    for (int e = 0; e < (1 << NB); ++e) {
      uint32_t x = sb;
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (e & (1 << j)) x ^= so[j];
      s_p[x] = w[e];
    }
  }
}

// acc_e += c * (+-a_{e ^ XR}): one X/Y-type term of lambda = sum g O psi (K3)
template <int XR>
__device__ __forceinline__ void xterm_accumulate(float2 (&acc)[16], const float2 (&a)[16],
                                                 float4 c4, uint32_t sign16) {
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int k = e ^ XR;
    const uint32_t flip = ((sign16 >> k) & 1u) << 31;
    const float2 b = make_float2(__uint_as_float(__float_as_uint(a[k].x) ^ flip),
                                 __uint_as_float(__float_as_uint(a[k].y) ^ flip));
    acc[e] = pmac(c4, b, swp(b), acc[e]);
  }
}

