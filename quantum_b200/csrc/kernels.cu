// Hand-written sm_100a kernels for the TFQ state-vector hot path.
// See kernels.cuh for the reference call sites each launch replaces.
#include "kernels.cuh"

#include <cassert>
#include <cstdio>

#include "gates.cuh"

namespace tfqb {
namespace {

constexpr int kThreads = 256;
constexpr unsigned kFull = 0xffffffffu;

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
// register-level gate application. `a` holds 2^R amplitudes; bit j of the
// array index is register bit j of the round.
// ------------------------------------------------------------------------
template <int R, int J, bool CTRL>
__device__ __forceinline__ void apply_g1(float2 (&a)[1 << R], const float2 (&m)[4],
                                         uint32_t cm, uint32_t cb) {
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    if (CTRL && ((e & cm) != cb)) continue;
    const float2 a0 = a[e], a1 = a[e | (1 << J)];
    a[e] = cfma(m[1], a1, cmulf(m[0], a0));
    a[e | (1 << J)] = cfma(m[3], a1, cmulf(m[2], a0));
  }
}

template <int R, int B0, int B1, bool CTRL>
__device__ __forceinline__ void apply_g2(float2 (&a)[1 << R], const float2 (&m)[16],
                                         uint32_t cm, uint32_t cb) {
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & ((1 << B0) | (1 << B1))) continue;
    if (CTRL && ((e & cm) != cb)) continue;
    const float2 a0 = a[e], a1 = a[e | (1 << B1)], a2 = a[e | (1 << B0)],
                 a3 = a[e | (1 << B0) | (1 << B1)];
    a[e] = cfma(m[3], a3, cfma(m[2], a2, cfma(m[1], a1, cmulf(m[0], a0))));
    a[e | (1 << B1)] =
        cfma(m[7], a3, cfma(m[6], a2, cfma(m[5], a1, cmulf(m[4], a0))));
    a[e | (1 << B0)] =
        cfma(m[11], a3, cfma(m[10], a2, cfma(m[9], a1, cmulf(m[8], a0))));
    a[e | (1 << B0) | (1 << B1)] =
        cfma(m[15], a3, cfma(m[14], a2, cfma(m[13], a1, cmulf(m[12], a0))));
  }
}

// 2 Re<l| D |a> restricted to this thread's amplitudes, D dense 2x2
template <int R, int J, bool CTRL>
__device__ __forceinline__ float grad_g1(const float2 (&a)[1 << R],
                                         const float2 (&l)[1 << R],
                                         const float2 (&m)[4], uint32_t cm,
                                         uint32_t cb) {
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    if (CTRL && ((e & cm) != cb)) continue;
    const float2 a0 = a[e], a1 = a[e | (1 << J)];
    const float2 p0 = cfma(m[1], a1, cmulf(m[0], a0));
    const float2 p1 = cfma(m[3], a1, cmulf(m[2], a0));
    acc += redot(l[e], p0) + redot(l[e | (1 << J)], p1);
  }
  return acc;
}

template <int R, int B0, int B1, bool CTRL>
__device__ __forceinline__ float grad_g2(const float2 (&a)[1 << R],
                                         const float2 (&l)[1 << R],
                                         const float2 (&m)[16], uint32_t cm,
                                         uint32_t cb) {
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & ((1 << B0) | (1 << B1))) continue;
    if (CTRL && ((e & cm) != cb)) continue;
    const int i1 = e | (1 << B1), i2 = e | (1 << B0), i3 = i1 | i2;
    const float2 a0 = a[e], a1 = a[i1], a2 = a[i2], a3 = a[i3];
    acc += redot(l[e], cfma(m[3], a3, cfma(m[2], a2, cfma(m[1], a1, cmulf(m[0], a0)))));
    acc += redot(l[i1], cfma(m[7], a3, cfma(m[6], a2, cfma(m[5], a1, cmulf(m[4], a0)))));
    acc += redot(l[i2], cfma(m[11], a3, cfma(m[10], a2, cfma(m[9], a1, cmulf(m[8], a0)))));
    acc += redot(l[i3], cfma(m[15], a3, cfma(m[14], a2, cfma(m[13], a1, cmulf(m[12], a0)))));
  }
  return acc;
}

__device__ __forceinline__ void load_m4(const float* sm, float2 (&m)[4]) {
  const float4 u = *reinterpret_cast<const float4*>(sm);
  const float4 v = *reinterpret_cast<const float4*>(sm + 4);
  m[0] = make_float2(u.x, u.y); m[1] = make_float2(u.z, u.w);
  m[2] = make_float2(v.x, v.y); m[3] = make_float2(v.z, v.w);
}
__device__ __forceinline__ void load_m16(const float* sm, float2 (&m)[16]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 u = *reinterpret_cast<const float4*>(sm + 4 * k);
    m[2 * k] = make_float2(u.x, u.y);
    m[2 * k + 1] = make_float2(u.z, u.w);
  }
}

template <int R, bool CTRL>
__device__ __forceinline__ void dispatch_g1(float2 (&a)[1 << R], const float2 (&m)[4],
                                            int b0, uint32_t cm, uint32_t cb) {
  switch (b0) {
    case 0: apply_g1<R, 0, CTRL>(a, m, cm, cb); break;
    case 1: apply_g1<R, 1, CTRL>(a, m, cm, cb); break;
    case 2: apply_g1<R, 2, CTRL>(a, m, cm, cb); break;
    default:
      if constexpr (R > 3) apply_g1<R, 3, CTRL>(a, m, cm, cb);
      break;
  }
}

template <int R, bool CTRL>
__device__ __forceinline__ void dispatch_g2(float2 (&a)[1 << R], const float2 (&m)[16],
                                            int b0, int b1, uint32_t cm, uint32_t cb) {
  switch (b0 * (b0 - 1) / 2 + b1) {  // b0 > b1
    case 0: apply_g2<R, 1, 0, CTRL>(a, m, cm, cb); break;
    case 1: apply_g2<R, 2, 0, CTRL>(a, m, cm, cb); break;
    case 2: apply_g2<R, 2, 1, CTRL>(a, m, cm, cb); break;
    case 3: if constexpr (R > 3) apply_g2<R, 3, 0, CTRL>(a, m, cm, cb); break;
    case 4: if constexpr (R > 3) apply_g2<R, 3, 1, CTRL>(a, m, cm, cb); break;
    default: if constexpr (R > 3) apply_g2<R, 3, 2, CTRL>(a, m, cm, cb); break;
  }
}

template <int R, bool CTRL>
__device__ __forceinline__ float dispatch_grad1(const float2 (&a)[1 << R],
                                                const float2 (&l)[1 << R],
                                                const float2 (&m)[4], int b0,
                                                uint32_t cm, uint32_t cb) {
  switch (b0) {
    case 0: return grad_g1<R, 0, CTRL>(a, l, m, cm, cb);
    case 1: return grad_g1<R, 1, CTRL>(a, l, m, cm, cb);
    case 2: return grad_g1<R, 2, CTRL>(a, l, m, cm, cb);
    default:
      if constexpr (R > 3) return grad_g1<R, 3, CTRL>(a, l, m, cm, cb);
      return 0.f;
  }
}

template <int R, bool CTRL>
__device__ __forceinline__ float dispatch_grad2(const float2 (&a)[1 << R],
                                                const float2 (&l)[1 << R],
                                                const float2 (&m)[16], int b0,
                                                int b1, uint32_t cm, uint32_t cb) {
  switch (b0 * (b0 - 1) / 2 + b1) {
    case 0: return grad_g2<R, 1, 0, CTRL>(a, l, m, cm, cb);
    case 1: return grad_g2<R, 2, 0, CTRL>(a, l, m, cm, cb);
    case 2: return grad_g2<R, 2, 1, CTRL>(a, l, m, cm, cb);
    case 3: if constexpr (R > 3) return grad_g2<R, 3, 0, CTRL>(a, l, m, cm, cb); return 0.f;
    case 4: if constexpr (R > 3) return grad_g2<R, 3, 1, CTRL>(a, l, m, cm, cb); return 0.f;
    default: if constexpr (R > 3) return grad_g2<R, 3, 2, CTRL>(a, l, m, cm, cb); return 0.f;
  }
}

// ------------------------------------------------------------------------
// diagonal ops.  A diagonal gate multiplies amplitude i by d[sel(i)], sel
// formed from 1..2 index bits.  Per round each selector bit is either a
// register bit (compile-time after dispatch) or constant for the thread.
// ------------------------------------------------------------------------
// generic (runtime masks): only for controlled diagonal gates
template <int R>
__device__ __forceinline__ void apply_diag_generic(float2 (&a)[1 << R], const float* sm,
                                                   uint32_t rm0, uint32_t rm1, int w0,
                                                   int w1, int selbase, uint32_t cm,
                                                   uint32_t cb) {
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if ((e & cm) != cb) continue;
    const int sel = selbase + ((e & rm0) ? w0 : 0) + ((e & rm1) ? w1 : 0);
    const float2 d = *reinterpret_cast<const float2*>(sm + 2 * sel);
    a[e] = cmulf(a[e], d);
  }
}

template <int R>
__device__ __forceinline__ float grad_diag_generic(const float2 (&a)[1 << R],
                                                   const float2 (&l)[1 << R],
                                                   const float* sm, uint32_t rm0,
                                                   uint32_t rm1, int w0, int w1,
                                                   int selbase, uint32_t cm, uint32_t cb) {
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if ((e & cm) != cb) continue;
    const int sel = selbase + ((e & rm0) ? w0 : 0) + ((e & rm1) ? w1 : 0);
    const float2 d = *reinterpret_cast<const float2*>(sm + 2 * sel);
    acc += redot(l[e], cmulf(a[e], d));
  }
  return acc;
}

template <int R>
__device__ __forceinline__ void scale_all(float2 (&a)[1 << R], float2 f) {
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) a[e] = cmulf(a[e], f);
}

// one selector bit is register bit J: entries f0 (bit clear) / f1 (bit set)
template <int R, int J>
__device__ __forceinline__ void diag1(float2 (&a)[1 << R], float2 f0, float2 f1,
                                      bool do0, bool do1) {
  if (do0) {
#pragma unroll
    for (int e = 0; e < (1 << R); ++e)
      if (!(e & (1 << J))) a[e] = cmulf(a[e], f0);
  }
  if (do1) {
#pragma unroll
    for (int e = 0; e < (1 << R); ++e)
      if (e & (1 << J)) a[e] = cmulf(a[e], f1);
  }
}
template <int R>
__device__ __forceinline__ void dispatch_diag1(float2 (&a)[1 << R], int j, float2 f0,
                                               float2 f1, bool do0, bool do1) {
  switch (j) {
    case 0: diag1<R, 0>(a, f0, f1, do0, do1); break;
    case 1: diag1<R, 1>(a, f0, f1, do0, do1); break;
    case 2: diag1<R, 2>(a, f0, f1, do0, do1); break;
    default: if constexpr (R > 3) diag1<R, 3>(a, f0, f1, do0, do1); break;
  }
}

// both selector bits are register bits: JH = register of the selector msb
template <int R, int JH, int JL>
__device__ __forceinline__ void diag2(float2 (&a)[1 << R], const float2 (&d)[4],
                                      uint32_t skip) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    if ((skip >> s) & 1u) continue;     // uniform
#pragma unroll
    for (int e = 0; e < (1 << R); ++e)
      if ((((e >> JH) & 1) * 2 + ((e >> JL) & 1)) == s) a[e] = cmulf(a[e], d[s]);
  }
}
template <int R>
__device__ __forceinline__ void dispatch_diag2(float2 (&a)[1 << R], int jh, int jl,
                                               float2 (&d)[4], uint32_t skip) {
  if (jh < jl) {   // canonical JH > JL: exchange selector bits
    const float2 t = d[1]; d[1] = d[2]; d[2] = t;
    skip = (skip & 9u) | ((skip & 2u) << 1) | ((skip & 4u) >> 1);
    const int t2 = jh; jh = jl; jl = t2;
  }
  switch (jh * (jh - 1) / 2 + jl) {
    case 0: diag2<R, 1, 0>(a, d, skip); break;
    case 1: diag2<R, 2, 0>(a, d, skip); break;
    case 2: diag2<R, 2, 1>(a, d, skip); break;
    case 3: if constexpr (R > 3) diag2<R, 3, 0>(a, d, skip); break;
    case 4: if constexpr (R > 3) diag2<R, 3, 1>(a, d, skip); break;
    default: if constexpr (R > 3) diag2<R, 3, 2>(a, d, skip); break;
  }
}

// gradient of a diagonal gate: sum_e Re(conj(l_e) * d[sel(e)] * a_e)
template <int R, int J>
__device__ __forceinline__ float gdiag1(const float2 (&a)[1 << R], const float2 (&l)[1 << R],
                                        float2 f0, float2 f1) {
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < (1 << R); ++e)
    acc += redot(l[e], cmulf(a[e], (e & (1 << J)) ? f1 : f0));
  return acc;
}
template <int R>
__device__ __forceinline__ float dispatch_gdiag1(const float2 (&a)[1 << R],
                                                 const float2 (&l)[1 << R], int j,
                                                 float2 f0, float2 f1) {
  switch (j) {
    case 0: return gdiag1<R, 0>(a, l, f0, f1);
    case 1: return gdiag1<R, 1>(a, l, f0, f1);
    case 2: return gdiag1<R, 2>(a, l, f0, f1);
    default:
      if constexpr (R > 3) return gdiag1<R, 3>(a, l, f0, f1);
      return 0.f;
  }
}
template <int R, int JH, int JL>
__device__ __forceinline__ float gdiag2(const float2 (&a)[1 << R], const float2 (&l)[1 << R],
                                        const float2 (&d)[4]) {
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < (1 << R); ++e)
    acc += redot(l[e], cmulf(a[e], d[((e >> JH) & 1) * 2 + ((e >> JL) & 1)]));
  return acc;
}
template <int R>
__device__ __forceinline__ float dispatch_gdiag2(const float2 (&a)[1 << R],
                                                 const float2 (&l)[1 << R], int jh,
                                                 int jl, float2 (&d)[4]) {
  if (jh < jl) {
    const float2 t = d[1]; d[1] = d[2]; d[2] = t;
    const int t2 = jh; jh = jl; jl = t2;
  }
  switch (jh * (jh - 1) / 2 + jl) {
    case 0: return gdiag2<R, 1, 0>(a, l, d);
    case 1: return gdiag2<R, 2, 0>(a, l, d);
    case 2: return gdiag2<R, 2, 1>(a, l, d);
    case 3: if constexpr (R > 3) return gdiag2<R, 3, 0>(a, l, d); return 0.f;
    case 4: if constexpr (R > 3) return gdiag2<R, 3, 1>(a, l, d); return 0.f;
    default: if constexpr (R > 3) return gdiag2<R, 3, 2>(a, l, d); return 0.f;
  }
}

__device__ __forceinline__ float2 ld_c(const float* sm, int idx) {
  return *reinterpret_cast<const float2*>(sm + 2 * idx);
}

// ------------------------------------------------------------------------
// The cache-blocked pass kernel (Q1). One CTA = one tile of one row.
//   smem: [psi tile][lam tile (ADJ)][pass matrices][hi table][ops][grad acc]
// ------------------------------------------------------------------------
template <int R, bool ADJ>
__global__ void __launch_bounds__(kThreads, 2)
pass_kernel(float2* __restrict__ psi, float2* __restrict__ lam,
            size_t row_stride, const PassRec* __restrict__ passes,
            const RoundRec* __restrict__ rounds, const OpRec* __restrict__ ops,
            const float* __restrict__ mats, size_t mat_row_stride,
            int pass_index, int first_op, int n_ops_in_pass,
            double* __restrict__ grad_out, int n_slots, int init_zero_state) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PassRec& P = passes[pass_index];
  const int t = P.tile_bits;
  const int L = P.low_bits;
  const uint32_t tile_size = 1u << t;
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  const size_t row = blockIdx.y;

  float2* s_psi = reinterpret_cast<float2*>(smem_raw);
  float2* s_lam = s_psi + (ADJ ? tile_size : 0);
  float* s_mat = reinterpret_cast<float*>(s_lam + tile_size);
  const int mat_len = P.mat_len;
  unsigned long long* s_hi =
      reinterpret_cast<unsigned long long*>(s_mat + ((mat_len + 3) & ~3));
  OpRec* s_ops = reinterpret_cast<OpRec*>(s_hi + (1u << (t - L)));
  float* s_grad = reinterpret_cast<float*>(s_ops + n_ops_in_pass);

  // tile base: scatter the tile id over the non-tile bit positions
  unsigned long long base = 0;
  {
    const unsigned long long tile = blockIdx.x;
    const int nc = P.n_comp;
    for (int k = 0; k < nc; ++k)
      base |= ((tile >> k) & 1ull) << P.comp_pos[k];
  }
  for (uint32_t h = tid; h < (1u << (t - L)); h += nthr) {
    unsigned long long v = 0;
    for (int k = 0; k < t - L; ++k)
      v |= (unsigned long long)((h >> k) & 1u) << P.tile_pos[L + k];
    s_hi[h] = v;
  }
  {
    const float* src = mats + row * mat_row_stride + P.mat_begin;
    for (int i = tid; i < mat_len; i += nthr) s_mat[i] = src[i];
    const uint32_t* osrc = reinterpret_cast<const uint32_t*>(ops + first_op);
    uint32_t* odst = reinterpret_cast<uint32_t*>(s_ops);
    const int nw = n_ops_in_pass * int(sizeof(OpRec) / 4);
    for (int i = tid; i < nw; i += nthr) odst[i] = osrc[i];
  }
  if (ADJ)
    for (int i = tid; i < n_ops_in_pass; i += nthr) s_grad[i] = 0.f;
  __syncthreads();

  const uint32_t lowmask = (1u << L) - 1u;
  float2* g_psi = psi + row * row_stride;
  float2* g_lam = ADJ ? lam + row * row_stride : nullptr;

  // ---- load the tile: 16-byte vectors, 2^L*8-byte contiguous runs
  for (uint32_t c0 = 0; c0 < tile_size / 2; c0 += nthr * 4) {
    float4 v[4], w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t c = c0 + u * nthr + tid;
      if (c < tile_size / 2) {
        const uint32_t i = 2 * c;
        const unsigned long long g = base | (i & lowmask) | s_hi[i >> L];
        if (init_zero_state) {
          v[u] = make_float4(g == 0 ? 1.f : 0.f, 0.f, 0.f, 0.f);
        } else {
          v[u] = *reinterpret_cast<const float4*>(g_psi + g);
        }
        if (ADJ) w[u] = *reinterpret_cast<const float4*>(g_lam + g);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t c = c0 + u * nthr + tid;
      if (c < tile_size / 2) {
        const uint32_t i = 2 * c;
        s_psi[swz(i)] = make_float2(v[u].x, v[u].y);
        s_psi[swz(i + 1)] = make_float2(v[u].z, v[u].w);
        if (ADJ) {
          s_lam[swz(i)] = make_float2(w[u].x, w[u].y);
          s_lam[swz(i + 1)] = make_float2(w[u].z, w[u].w);
        }
      }
    }
  }
  __syncthreads();

  // ---- rounds
  const uint32_t ngroups = tile_size >> R;
  const uint32_t iters = (ngroups + nthr - 1) / nthr;
  for (int r = P.round_begin; r < P.round_end; ++r) {
    const RoundRec rr = rounds[r];
    uint32_t o[R], so[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      o[j] = 1u << rr.pos[j];
      so[j] = swz(o[j]);     // swz is GF(2)-linear: swz(b|off) = swz(b)^swz(off)
    }
    for (uint32_t it = 0; it < iters; ++it) {
      const uint32_t gi = it * nthr + tid;
      const bool active = gi < ngroups;
      uint32_t b = active ? gi : 0;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const uint32_t lo = o[j] - 1u;
        b = ((b & ~lo) << 1) | (b & lo);
      }
      const uint32_t sb = swz(b);
      float2 a[1 << R];
      float2 l[ADJ ? (1 << R) : 1];
#pragma unroll
      for (int e = 0; e < (1 << R); ++e) {
        uint32_t x = sb;
#pragma unroll
        for (int j = 0; j < R; ++j)
          if (e & (1 << j)) x ^= so[j];
        a[e] = s_psi[x];
        if constexpr (ADJ) l[e] = s_lam[x];
      }
      const unsigned long long gbase = base | (b & lowmask) | s_hi[b >> L];
      // forward kernel: scalar phase of the diagonal ops that touch no
      // register bit of this round; applied once at the end of the round
      float2 ph = make_float2(1.f, 0.f);
      bool ph_dirty = false;

      for (int oi = rr.op_begin; oi < rr.op_end; ++oi) {
        const OpRec& op = s_ops[oi - first_op];
        const int kind = op.kind;
        const uint32_t cm = op.creg_mask, cb = op.creg_bits;
        const bool rest_ok =
            active && ((gbase & op.crest_mask) == op.crest_bits);
        const float* sm = s_mat + op.mat_off;
        const int tgt = ADJ ? op.target : kTgtPsi;
        if (kind == kOpG1) {
          if (rest_ok) {
            float2 m[4];
            load_m4(sm, m);
            if (cm == 0) {
              if (tgt & kTgtPsi) dispatch_g1<R, false>(a, m, op.b0, 0, 0);
              if constexpr (ADJ) { if (tgt & kTgtLam) dispatch_g1<R, false>(l, m, op.b0, 0, 0); }
            } else {
              if (tgt & kTgtPsi) dispatch_g1<R, true>(a, m, op.b0, cm, cb);
              if constexpr (ADJ) { if (tgt & kTgtLam) dispatch_g1<R, true>(l, m, op.b0, cm, cb); }
            }
          }
        } else if (kind == kOpG2) {
          if (rest_ok) {
            float2 m[16];
            load_m16(sm, m);
            if (cm == 0) {
              if (tgt & kTgtPsi) dispatch_g2<R, false>(a, m, op.b0, op.b1, 0, 0);
              if constexpr (ADJ) { if (tgt & kTgtLam) dispatch_g2<R, false>(l, m, op.b0, op.b1, 0, 0); }
            } else {
              if (tgt & kTgtPsi) dispatch_g2<R, true>(a, m, op.b0, op.b1, cm, cb);
              if constexpr (ADJ) { if (tgt & kTgtLam) dispatch_g2<R, true>(l, m, op.b0, op.b1, cm, cb); }
            }
          }
        } else if (kind == kOpD || kind == kOpGradD) {
          const bool two = op.dpos1 >= 0;
          const int r0 = op.dreg0, r1 = two ? op.dreg1 : -1;
          // selector contribution of the thread-constant bits
          const int c0 = r0 < 0 ? int((gbase >> op.dpos0) & 1ull) : 0;
          const int c1 = (two && r1 < 0) ? int((gbase >> op.dpos1) & 1ull) : 0;
          const int w0 = two ? 2 : 1;
          if (cm != 0) {   // controlled diagonal gate: generic path
            const uint32_t rm0 = r0 >= 0 ? (1u << r0) : 0u;
            const uint32_t rm1 = r1 >= 0 ? (1u << r1) : 0u;
            const int selbase = c0 * w0 + c1;
            if (kind == kOpD) {
              if (rest_ok) {
                if (tgt & kTgtPsi) apply_diag_generic<R>(a, sm, rm0, rm1, w0, 1, selbase, cm, cb);
                if constexpr (ADJ) { if (tgt & kTgtLam) apply_diag_generic<R>(l, sm, rm0, rm1, w0, 1, selbase, cm, cb); }
              }
            } else if constexpr (ADJ) {
              if (ph_dirty) { scale_all<R>(a, ph); ph = make_float2(1.f, 0.f); ph_dirty = false; }
              float v = 0.f;
              if (rest_ok) v = grad_diag_generic<R>(a, l, sm, rm0, rm1, w0, 1, selbase, cm, cb);
#pragma unroll
              for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
              if ((tid & 31) == 0 && v != 0.f) atomicAdd(&s_grad[oi - first_op], 2.f * v);
            }
          } else if (r0 < 0 && r1 < 0) {          // D0: constant for the thread
            const float2 f = ld_c(sm, c0 * w0 + c1);
            if (kind == kOpD) {
              if constexpr (ADJ) {
                if (rest_ok) {
                  if (tgt & kTgtPsi) scale_all<R>(a, f);
                  if (tgt & kTgtLam) scale_all<R>(l, f);
                }
              } else {
                if (rest_ok) ph = cmulf(ph, f);
                ph_dirty = true;
              }
            } else if constexpr (ADJ) {
              float v = 0.f;
              if (rest_ok) {
#pragma unroll
                for (int e = 0; e < (1 << R); ++e) v += redot(l[e], cmulf(a[e], f));
              }
#pragma unroll
              for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
              if ((tid & 31) == 0 && v != 0.f) atomicAdd(&s_grad[oi - first_op], 2.f * v);
            }
          } else if (r0 >= 0 && r1 >= 0) {        // D2: both register bits
            float2 d[4] = {ld_c(sm, 0), ld_c(sm, 1), ld_c(sm, 2), ld_c(sm, 3)};
            if (kind == kOpD) {
              if (rest_ok) {
                if (tgt & kTgtPsi) dispatch_diag2<R>(a, r0, r1, d, op.ident_mask);
                if constexpr (ADJ) {
                  if (tgt & kTgtLam) {
                    float2 d2[4] = {ld_c(sm, 0), ld_c(sm, 1), ld_c(sm, 2), ld_c(sm, 3)};
                    dispatch_diag2<R>(l, r0, r1, d2, op.ident_mask);
                  }
                }
              }
            } else if constexpr (ADJ) {
              float v = 0.f;
              if (rest_ok) v = dispatch_gdiag2<R>(a, l, r0, r1, d);
#pragma unroll
              for (int dd = 16; dd > 0; dd >>= 1) v += __shfl_xor_sync(kFull, v, dd);
              if ((tid & 31) == 0 && v != 0.f) atomicAdd(&s_grad[oi - first_op], 2.f * v);
            }
          } else {                                 // D1: one register bit
            int j, s0, s1;
            bool do0 = true, do1 = true;
            if (!two) {
              j = r0; s0 = 0; s1 = 1;
              do0 = !(op.ident_mask & 1u);
              do1 = !(op.ident_mask & 2u);
            } else if (r0 >= 0) {   // register bit is the selector msb
              j = r0; s0 = c1; s1 = 2 + c1;
            } else {                // register bit is the selector lsb
              j = r1; s0 = 2 * c0; s1 = 2 * c0 + 1;
            }
            const float2 f0 = ld_c(sm, s0), f1 = ld_c(sm, s1);
            if (kind == kOpD) {
              if (rest_ok) {
                if (tgt & kTgtPsi) dispatch_diag1<R>(a, j, f0, f1, do0, do1);
                if constexpr (ADJ) { if (tgt & kTgtLam) dispatch_diag1<R>(l, j, f0, f1, do0, do1); }
              }
            } else if constexpr (ADJ) {
              float v = 0.f;
              if (rest_ok) v = dispatch_gdiag1<R>(a, l, j, f0, f1);
#pragma unroll
              for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
              if ((tid & 31) == 0 && v != 0.f) atomicAdd(&s_grad[oi - first_op], 2.f * v);
            }
          }
        } else if (kind == kOpGrad1) {
         if constexpr (ADJ) {
          float v = 0.f;
          if (rest_ok) {
            float2 m[4];
            load_m4(sm, m);
            v = cm == 0 ? dispatch_grad1<R, false>(a, l, m, op.b0, 0, 0)
                        : dispatch_grad1<R, true>(a, l, m, op.b0, cm, cb);
          }
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
          if ((tid & 31) == 0 && v != 0.f) atomicAdd(&s_grad[oi - first_op], 2.f * v);
         }
        } else if (kind == kOpGrad2) {
         if constexpr (ADJ) {
          float v = 0.f;
          if (rest_ok) {
            float2 m[16];
            load_m16(sm, m);
            v = cm == 0 ? dispatch_grad2<R, false>(a, l, m, op.b0, op.b1, 0, 0)
                        : dispatch_grad2<R, true>(a, l, m, op.b0, op.b1, cm, cb);
          }
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
          if ((tid & 31) == 0 && v != 0.f) atomicAdd(&s_grad[oi - first_op], 2.f * v);
         }
        }
      }

      if (active) {
        if (!ADJ && ph_dirty) scale_all<R>(a, ph);
#pragma unroll
        for (int e = 0; e < (1 << R); ++e) {
          uint32_t x = sb;
#pragma unroll
          for (int j = 0; j < R; ++j)
            if (e & (1 << j)) x ^= so[j];
          s_psi[x] = a[e];
          if constexpr (ADJ) s_lam[x] = l[e];
        }
      }
    }
    __syncthreads();
  }

  // ---- store the tile
  for (uint32_t c = tid; c < tile_size / 2; c += nthr) {
    const uint32_t i = 2 * c;
    const unsigned long long g = base | (i & lowmask) | s_hi[i >> L];
    const float2 p0 = s_psi[swz(i)], p1 = s_psi[swz(i + 1)];
    *reinterpret_cast<float4*>(g_psi + g) = make_float4(p0.x, p0.y, p1.x, p1.y);
    if (ADJ) {
      const float2 q0 = s_lam[swz(i)], q1 = s_lam[swz(i + 1)];
      *reinterpret_cast<float4*>(g_lam + g) = make_float4(q0.x, q0.y, q1.x, q1.y);
    }
  }
  if (ADJ) {
    for (int i = tid; i < n_ops_in_pass; i += nthr) {
      const int slot = s_ops[i].grad_slot;
      const float v = s_grad[i];
      if (slot >= 0 && v != 0.f)
        atomicAdd(&grad_out[row * size_t(n_slots) + slot], double(v));
    }
  }
}

// ------------------------------------------------------------------------
// per-row matrix builder: product of the op's factors (first applied first)
// ------------------------------------------------------------------------
__device__ __forceinline__ cf cmul_f(cf a, cf b) {
  return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}

__global__ void build_matrices_kernel(const MatRec* __restrict__ recs,
                                      const FactorRec* __restrict__ factors,
                                      int n_recs,
                                      const float* __restrict__ params,
                                      int n_params, int rows,
                                      float* __restrict__ out,
                                      size_t out_row_stride) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * n_recs) return;
  const int row = int(idx / n_recs);
  const MatRec rec = recs[idx % n_recs];
  const int dim = (rec.layout == 0 || rec.layout == 2) ? 2 : 4;
  cf m[16];
  for (int fi = rec.factor_begin; fi < rec.factor_end; ++fi) {
    const FactorRec f = factors[fi];
    float p[5];
#pragma unroll
    for (int k = 0; k < 5; ++k)
      p[k] = (k < f.nparams)
                 ? (f.sym[k] >= 0 ? params[size_t(row) * n_params + f.sym[k]]
                                  : f.value[k])
                 : 0.f;
    cf g[16];
    const int gdim = (dim == 4 && f.slot <= 1) ? 2 : dim;
    if (rec.mode == kMatGrad)
      gradient_matrix(f.gate_kind, p, rec.shift_idx, gdim, g);
    else
      gate_matrix(f.gate_kind, p, -1, 0.f, g);
    if (rec.factor_end - rec.factor_begin == 1 && gdim == dim && f.slot != 3) {
      for (int i = 0; i < dim * dim; ++i) m[i] = g[i];   // exact single gate
      break;
    }
    cf e[16];
    if (gdim == dim) {
      for (int r = 0; r < dim; ++r)
        for (int c = 0; c < dim; ++c) {
          int rr = r, cc = c;
          if (f.slot == 3) {
            rr = ((r & 1) << 1) | (r >> 1);
            cc = ((c & 1) << 1) | (c >> 1);
          }
          e[r * dim + c] = g[rr * dim + cc];
        }
    } else {   // 1-qubit gate embedded in the 4x4: slot 0 = msb, 1 = lsb
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
          const int rh = r >> 1, rl = r & 1, ch = c >> 1, cl = c & 1;
          cf v = mk(0.f, 0.f);
          if (f.slot == 0) { if (rl == cl) v = g[rh * 2 + ch]; }
          else { if (rh == ch) v = g[rl * 2 + cl]; }
          e[r * 4 + c] = v;
        }
    }
    if (fi == rec.factor_begin) {
      for (int i = 0; i < dim * dim; ++i) m[i] = e[i];
    } else {
      cf t[16];
      for (int r = 0; r < dim; ++r)
        for (int c = 0; c < dim; ++c) {
          cf acc = mk(0.f, 0.f);
          for (int k = 0; k < dim; ++k) {
            const cf pr = cmul_f(e[r * dim + k], m[k * dim + c]);
            acc.re += pr.re;
            acc.im += pr.im;
          }
          t[r * dim + c] = acc;
        }
      for (int i = 0; i < dim * dim; ++i) m[i] = t[i];
    }
  }
  float* o = out + size_t(row) * out_row_stride + rec.out_off;
  const bool dag = rec.mode == kMatDagger;
  if (rec.layout >= 2) {  // diagonal: d[0..dim)
    for (int i = 0; i < 4; ++i) {
      cf v = i < dim ? m[i * dim + i] : mk(0.f, 0.f);
      if (dag) v.im = -v.im;
      o[2 * i] = v.re;
      o[2 * i + 1] = v.im;
    }
    return;
  }
  for (int r = 0; r < dim; ++r) {
    for (int c = 0; c < dim; ++c) {
      int rr = r, cc = c;
      if (rec.swap) {  // exchange the two qubits
        rr = ((r & 1) << 1) | (r >> 1);
        cc = ((c & 1) << 1) | (c >> 1);
      }
      cf v = dag ? m[cc * dim + rr] : m[rr * dim + cc];
      if (dag) v.im = -v.im;
      o[2 * (r * dim + c)] = v.re;
      o[2 * (r * dim + c) + 1] = v.im;
    }
  }
}

// ------------------------------------------------------------------------
// Q2 primitives
// ------------------------------------------------------------------------
__global__ void set_zero_state_kernel(float2* psi, size_t row_stride) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= row_stride) return;
  psi[blockIdx.y * row_stride + i] = make_float2(i == 0 ? 1.f : 0.f, 0.f);
}

__global__ void export_state_kernel(const float2* __restrict__ psi,
                                    size_t row_stride, size_t n_amps,
                                    float2* __restrict__ out, size_t out_cols) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= out_cols) return;
  const size_t row = blockIdx.y;
  out[row * out_cols + i] =
      i < n_amps ? psi[row * row_stride + i] : make_float2(-2.f, 0.f);
}

// ------------------------------------------------------------------------
// K1: per-term expectation, generic masks (global partner gather)
// ------------------------------------------------------------------------
__device__ __forceinline__ double block_reduce_sum(double v, double* s_red) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) s_red[w] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) r += __shfl_xor_sync(kFull, r, d);
  }
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kThreads)
expectation_terms_kernel(const float2* __restrict__ psi, size_t row_stride,
                         unsigned long long n_amps,
                         const DevTerm* __restrict__ terms, int n_terms,
                         double* __restrict__ per_term) {
  __shared__ double s_red[32];
  const int t = blockIdx.y;
  const size_t row = blockIdx.z;
  const DevTerm term = terms[t];
  if (term.identity) return;
  const float2* st = psi + row * row_stride;
  double acc = 0.0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
       i < n_amps; i += stride) {
    const float2 a = st[i];
    float v;
    if (term.x == 0) {
      v = a.x * a.x + a.y * a.y;
      if (term.phase & 1) v = 0.f;           // Re(+-i |a|^2) = 0
      else if (term.phase & 2) v = -v;
      if (__popcll(i & term.z) & 1) v = -v;
    } else {
      const unsigned long long k = i ^ term.x;
      const float2 b = st[k];
      const float wr = a.x * b.x + a.y * b.y;   // Re(conj(a) b)
      const float wi = a.x * b.y - a.y * b.x;   // Im(conj(a) b)
      switch (term.phase & 3) {
        case 0: v = wr; break;
        case 1: v = -wi; break;
        case 2: v = -wr; break;
        default: v = wi; break;
      }
      if (__popcll(k & term.z) & 1) v = -v;
    }
    acc += double(v);
  }
  const double tot = block_reduce_sum(acc, s_red);
  if (threadIdx.x == 0 && tot != 0.0)
    atomicAdd(&per_term[row * size_t(n_terms) + t], tot);
}

__global__ void combine_terms_kernel(const double* __restrict__ per_term,
                                     const DevTerm* __restrict__ terms,
                                     int n_terms, int n_ops, int rows,
                                     float* __restrict__ out, size_t out_stride) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * n_ops) return;
  const int row = idx / n_ops, j = idx % n_ops;
  // `*expectation_value += coeff * RealInnerProduct(...)`: float += double
  float e = 0.f;
  for (int t = 0; t < n_terms; ++t) {
    const DevTerm term = terms[t];
    if (term.op != j) continue;
    if (term.identity) {
      e = __fadd_rn(e, term.coeff);
    } else {
      e = float(double(e) + double(term.coeff) * per_term[size_t(row) * n_terms + t]);
    }
  }
  out[size_t(row) * out_stride + j] = e;
}

// ------------------------------------------------------------------------
// K3: lambda = sum_j g_j sum_t c_t P_t psi   (util_qsim.h:362-414)
// ------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
accumulate_operators_kernel(const float2* __restrict__ psi,
                            float2* __restrict__ lam, size_t row_stride,
                            unsigned long long n_amps,
                            const DevTerm* __restrict__ terms, int n_terms,
                            const float* __restrict__ downstream, int n_ops) {
  const size_t row = blockIdx.y;
  const float2* st = psi + row * row_stride;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
       i < n_amps; i += stride) {
    float2 acc = make_float2(0.f, 0.f);
    for (int t = 0; t < n_terms; ++t) {
      const DevTerm term = terms[t];
      const float lead = __fmul_rn(downstream[row * n_ops + term.op], term.coeff);
      if (fabsf(lead) < 1e-5f) continue;   // util_qsim.h:378-383
      float2 v;
      if (term.identity) {
        v = st[i];
      } else {
        const unsigned long long k = i ^ term.x;
        const float2 b = st[k];
        switch (term.phase & 3) {
          case 0: v = b; break;
          case 1: v = make_float2(-b.y, b.x); break;
          case 2: v = make_float2(-b.x, -b.y); break;
          default: v = make_float2(b.y, -b.x); break;
        }
        if (__popcll(k & term.z) & 1) v = make_float2(-v.x, -v.y);
      }
      acc.x = __fadd_rn(acc.x, __fmul_rn(lead, v.x));
      acc.y = __fadd_rn(acc.y, __fmul_rn(lead, v.y));
    }
    lam[row * row_stride + i] = acc;
  }
}

__global__ void reduce_grad_slots_kernel(const double* __restrict__ slot_vals,
                                         const int32_t* __restrict__ slot_col,
                                         int n_slots, int rows,
                                         float* __restrict__ grads, int n_cols) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * n_cols) return;
  const int row = idx / n_cols, col = idx % n_cols;
  // reference accumulates `out(i, loc) += float(2 Re<..>)` per gradient gate
  float g = 0.f;
  for (int s = 0; s < n_slots; ++s)
    if (slot_col[s] == col)
      g = float(double(g) + slot_vals[size_t(row) * n_slots + s]);
  grads[size_t(row) * n_cols + col] = g;
}

// ------------------------------------------------------------------------
// Q3: sampling on a canonical fp64 pairwise tree
// ------------------------------------------------------------------------
__device__ __forceinline__ size_t tree_level_offset(int nc, int level) {
  // levels kTreeChunkBits..nc stored back to back, level l has 2^(nc-l) nodes
  return (size_t(1) << (nc - kTreeChunkBits + 1)) - (size_t(1) << (nc - level + 1));
}

struct ChunkTree {
  double p[8], s2[4], s4[2], w[6];
};

// all 32 lanes: lane holds amplitudes [8*lane, 8*lane+8) of the chunk
__device__ __forceinline__ void chunk_tree(const float2* __restrict__ st,
                                           unsigned long long chunk,
                                           unsigned long long n_amps,
                                           ChunkTree& c) {
  const int lane = threadIdx.x & 31;
  const unsigned long long i0 = (chunk << kTreeChunkBits) + 8ull * lane;
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i0 + k < n_amps) v = *reinterpret_cast<const float4*>(st + i0 + k);
    c.p[k] = double(v.x) * double(v.x) + double(v.y) * double(v.y);
    c.p[k + 1] = double(v.z) * double(v.z) + double(v.w) * double(v.w);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) c.s2[k] = c.p[2 * k] + c.p[2 * k + 1];
  c.s4[0] = c.s2[0] + c.s2[1];
  c.s4[1] = c.s2[2] + c.s2[3];
  c.w[0] = c.s4[0] + c.s4[1];
#pragma unroll
  for (int k = 0; k < 5; ++k)
    c.w[k + 1] = c.w[k] + __shfl_xor_sync(kFull, c.w[k], 1 << k);
}

__global__ void __launch_bounds__(kThreads)
tree_leaves_kernel(const float2* __restrict__ psi, size_t row_stride,
                   unsigned long long n_amps, int nc, double* __restrict__ tree,
                   size_t tree_row_stride) {
  const size_t row = blockIdx.y;
  const unsigned long long chunk =
      blockIdx.x * (unsigned long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chunk >= (1ull << (nc - kTreeChunkBits))) return;
  ChunkTree c;
  chunk_tree(psi + row * row_stride, chunk, n_amps, c);
  if ((threadIdx.x & 31) == 0) tree[row * tree_row_stride + chunk] = c.w[5];
}

// builds up to 9 further levels from level `lvl` inside one CTA
__global__ void __launch_bounds__(kThreads)
tree_upper_kernel(double* __restrict__ tree, size_t tree_row_stride, int nc,
                  int lvl, int n_levels) {
  __shared__ double s[2 * kThreads];
  double* tr = tree + blockIdx.y * tree_row_stride;
  const size_t n_in = size_t(1) << (nc - lvl);
  const size_t in0 = blockIdx.x * size_t(2 * kThreads);
  for (int k = threadIdx.x; k < 2 * kThreads; k += kThreads)
    s[k] = (in0 + k < n_in) ? tr[tree_level_offset(nc, lvl) + in0 + k] : 0.0;
  __syncthreads();
  int width = 2 * kThreads;
  for (int d = 1; d <= n_levels; ++d) {
    width >>= 1;
    double v = 0.0;
    const bool on = int(threadIdx.x) < width;
    if (on) v = s[2 * threadIdx.x] + s[2 * threadIdx.x + 1];
    __syncthreads();
    if (on) {
      s[threadIdx.x] = v;
      const size_t o = blockIdx.x * size_t(width) + threadIdx.x;
      if (o < (size_t(1) << (nc - lvl - d)))
        tr[tree_level_offset(nc, lvl + d) + o] = v;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads)
sample_kernel(const float2* __restrict__ psi, size_t row_stride,
              unsigned long long n_amps, int nc, const double* __restrict__ tree,
              size_t tree_row_stride, const double* __restrict__ uniforms,
              size_t uniform_row_stride, const int32_t* __restrict__ shots_per_row,
              int shots, unsigned long long* __restrict__ indices,
              size_t index_row_stride) {
  const size_t row = blockIdx.y;
  const int shot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int my_shots = shots_per_row ? shots_per_row[row] : shots;
  if (shot >= my_shots) return;
  const double* tr = tree + row * tree_row_stride;
  const double u = uniforms[row * uniform_row_stride + shot];
  double r = u * tr[tree_level_offset(nc, nc)];
  unsigned long long j = 0;
  for (int l = nc - 1; l >= kTreeChunkBits; --l) {
    const double left = tr[tree_level_offset(nc, l) + 2 * j];
    if (r < left) { j = 2 * j; } else { r -= left; j = 2 * j + 1; }
  }
  ChunkTree c;
  chunk_tree(psi + row * row_stride, j, n_amps, c);
  int pos = 0;
#pragma unroll
  for (int wi = 4; wi >= 0; --wi) {   // child level = 3 + wi
    const double left = __shfl_sync(kFull, c.w[wi], (2 * pos) << wi);
    if (r < left) { pos = 2 * pos; } else { r -= left; pos = 2 * pos + 1; }
  }
  // pos = lane owning the level-3 node
  int sub = 0;
  {
    const double left = __shfl_sync(kFull, c.s4[0], pos);
    if (!(r < left)) { r -= left; sub = 1; }
  }
  {
    const double cand = sub ? c.s2[2] : c.s2[0];
    const double left = __shfl_sync(kFull, cand, pos);
    sub = 2 * sub;
    if (!(r < left)) { r -= left; sub += 1; }
  }
  {
    const double cand = sub == 0 ? c.p[0] : sub == 1 ? c.p[2] : sub == 2 ? c.p[4] : c.p[6];
    const double left = __shfl_sync(kFull, cand, pos);
    sub = 2 * sub;
    if (!(r < left)) { sub += 1; }
  }
  if ((threadIdx.x & 31) == 0)
    indices[row * index_row_stride + shot] =
        (j << kTreeChunkBits) | (unsigned long long)(pos << 3) | (unsigned long long)sub;
}

// Philox4x32-10: key = seed, counter = (shot, row, stream_a, stream_b)
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const unsigned long long p0 = 0xD2511F53ull * c[0];
  const unsigned long long p1 = 0xCD9E8D57ull * c[2];
  const uint32_t n0 = uint32_t(p1 >> 32) ^ c[1] ^ k[0];
  const uint32_t n1 = uint32_t(p1);
  const uint32_t n2 = uint32_t(p0 >> 32) ^ c[3] ^ k[1];
  const uint32_t n3 = uint32_t(p0);
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u;
  k[1] += 0xBB67AE85u;
}

__global__ void fill_uniforms_kernel(double* __restrict__ u, size_t row_stride,
                                     unsigned long long seed,
                                     const int32_t* __restrict__ row_ids,
                                     uint32_t stream_a, uint32_t stream_b,
                                     int shots, size_t padded) {
  const size_t s = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (s >= padded) return;
  const size_t row = blockIdx.y;
  double v = 2.0;  // padding sorts to the end
  if (s < size_t(shots)) {
    uint32_t c[4] = {uint32_t(s), uint32_t(row_ids ? row_ids[row] : int(row)),
                     stream_a, stream_b};
    uint32_t k[2] = {uint32_t(seed), uint32_t(seed >> 32)};
#pragma unroll
    for (int i = 0; i < 10; ++i) philox_round(c, k);
    const unsigned long long x = ((unsigned long long)c[0] << 32) | c[1];
    v = double(x >> 11) * 0x1.0p-53;
  }
  u[row * row_stride + s] = v;
}

// in-place bitonic sort of each row (row_stride = padded power of two)
__global__ void __launch_bounds__(1024)
sort_rows_kernel(double* __restrict__ u, size_t row_stride, size_t padded) {
  double* a = u + blockIdx.x * row_stride;
  for (size_t k = 2; k <= padded; k <<= 1) {
    for (size_t j = k >> 1; j > 0; j >>= 1) {
      for (size_t i = threadIdx.x; i < padded; i += blockDim.x) {
        const size_t p = i ^ j;
        if (p > i) {
          const double x = a[i], y = a[p];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[p] = x; }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void add_constant_kernel(float c, int rows, float* acc, size_t acc_stride) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) acc[r * acc_stride] = __fadd_rn(acc[r * acc_stride], c);
}

__global__ void unpack_samples_kernel(const unsigned long long* __restrict__ indices,
                                      size_t index_row_stride, int n, int nmax,
                                      int shots, int8_t* __restrict__ out) {
  const size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  const size_t per_row = size_t(shots) * nmax;
  if (idx >= per_row) return;
  const size_t row = blockIdx.y;
  const int s = int(idx / nmax), col = int(idx % nmax);
  const int q = nmax - 1 - col;
  int8_t v = -2;
  if (q < n) v = int8_t((indices[row * index_row_stride + s] >> q) & 1ull);
  out[row * per_row + idx] = v;
}

__global__ void __launch_bounds__(kThreads)
parity_expectation_kernel(const unsigned long long* __restrict__ indices,
                          size_t index_row_stride, unsigned long long mask,
                          float coeff, const int32_t* __restrict__ shots_per_row,
                          int shots, float* __restrict__ acc, size_t acc_stride) {
  __shared__ double s_red[32];
  const size_t row = blockIdx.x;
  const int my_shots = shots_per_row ? shots_per_row[row] : shots;
  long long tot = 0;
  for (int s = threadIdx.x; s < my_shots; s += blockDim.x)
    tot += (__popcll(indices[row * index_row_stride + s] & mask) & 1) ? -1 : 1;
  const double t = block_reduce_sum(double(tot), s_red);
  if (threadIdx.x == 0) {
    const float term = __fdiv_rn(__fmul_rn(float(int(t)), coeff), float(my_shots));
    acc[row * acc_stride] = __fadd_rn(acc[row * acc_stride], term);
  }
}

inline unsigned cdiv(size_t a, size_t b) { return unsigned((a + b - 1) / b); }

}  // namespace

// ==========================================================================
// launch wrappers
// ==========================================================================
size_t ForwardPassSmem(int tile_bits, int mat_len, int n_ops) {
  const int L = tile_bits < kLowBits ? tile_bits : kLowBits;
  return (size_t(8) << tile_bits) + size_t((mat_len + 3) & ~3) * 4 +
         (size_t(8) << (tile_bits - L)) + size_t(n_ops) * sizeof(OpRec) + 16;
}
size_t AdjointPassSmem(int tile_bits, int mat_len, int n_ops) {
  const int L = tile_bits < kLowBits ? tile_bits : kLowBits;
  return (size_t(16) << tile_bits) + size_t((mat_len + 3) & ~3) * 4 +
         (size_t(8) << (tile_bits - L)) + size_t(n_ops) * (sizeof(OpRec) + 4) + 16;
}

static int pass_threads(int tile_bits, int reg_bits) {
  int g = 1 << (tile_bits - reg_bits);
  if (g < 32) g = 32;
  if (g > kThreads) g = kThreads;
  return g;
}

void LaunchForwardPass(const PassLaunch& pl, float2* psi, size_t row_stride,
                       int rows, bool init_zero_state, cudaStream_t s) {
  cudaFuncSetAttribute(pass_kernel<kRegBits, false>,
                       cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  const size_t smem = ForwardPassSmem(pl.tile_bits, pl.mat_len, pl.n_ops_in_pass);
  const dim3 grid(1u << (pl.n_alloc - pl.tile_bits), rows);
  pass_kernel<kRegBits, false><<<grid, pass_threads(pl.tile_bits, kRegBits), smem, s>>>(
      psi, nullptr, row_stride, pl.passes, pl.rounds, pl.ops, pl.mats,
      pl.mat_row_stride, pl.pass_index, pl.first_op, pl.n_ops_in_pass, nullptr,
      0, init_zero_state ? 1 : 0);
}

void LaunchAdjointPass(const PassLaunch& pl, float2* psi, float2* lam,
                       size_t row_stride, int rows, double* grad_out,
                       int n_slots, cudaStream_t s) {
  cudaFuncSetAttribute(pass_kernel<kRegBitsAdj, true>,
                       cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  const size_t smem = AdjointPassSmem(pl.tile_bits, pl.mat_len, pl.n_ops_in_pass);
  const dim3 grid(1u << (pl.n_alloc - pl.tile_bits), rows);
  pass_kernel<kRegBitsAdj, true>
      <<<grid, pass_threads(pl.tile_bits, kRegBitsAdj), smem, s>>>(
          psi, lam, row_stride, pl.passes, pl.rounds, pl.ops, pl.mats,
          pl.mat_row_stride, pl.pass_index, pl.first_op, pl.n_ops_in_pass,
          grad_out, n_slots, 0);
}

void LaunchBuildMatrices(const MatRec* recs, const FactorRec* factors,
                         int n_recs, const float* params, int n_params,
                         int rows, float* out, size_t out_row_stride,
                         cudaStream_t s) {
  if (n_recs == 0 || rows == 0) return;
  const size_t total = size_t(rows) * n_recs;
  build_matrices_kernel<<<cdiv(total, 128), 128, 0, s>>>(
      recs, factors, n_recs, params, n_params, rows, out, out_row_stride);
}

void LaunchSetZeroState(float2* psi, size_t row_stride, int rows, cudaStream_t s) {
  const dim3 grid(cdiv(row_stride, 256), rows);
  set_zero_state_kernel<<<grid, 256, 0, s>>>(psi, row_stride);
}

void LaunchExportState(const float2* psi, size_t row_stride, int n, float2* out,
                       size_t out_cols, int rows, cudaStream_t s) {
  const dim3 grid(cdiv(out_cols, 256), rows);
  export_state_kernel<<<grid, 256, 0, s>>>(psi, row_stride, size_t(1) << n, out,
                                           out_cols);
}

void LaunchExpectationTerms(const float2* psi, size_t row_stride, int n_alloc,
                            const DevTerm* terms, int n_terms, int rows,
                            double* per_term, cudaStream_t s) {
  if (n_terms == 0 || rows == 0) return;
  const size_t n_amps = size_t(1) << n_alloc;
  unsigned chunks = cdiv(n_amps, size_t(kThreads) * 16);
  if (chunks > 2048) chunks = 2048;
  const dim3 grid(chunks, n_terms, rows);
  expectation_terms_kernel<<<grid, kThreads, 0, s>>>(psi, row_stride, n_amps,
                                                     terms, n_terms, per_term);
}

void LaunchCombineTerms(const double* per_term, const DevTerm* terms,
                        int n_terms, int n_ops, int rows, float* out,
                        size_t out_stride, cudaStream_t s) {
  if (rows * n_ops == 0) return;
  combine_terms_kernel<<<cdiv(size_t(rows) * n_ops, 128), 128, 0, s>>>(
      per_term, terms, n_terms, n_ops, rows, out, out_stride);
}

void LaunchAccumulateOperators(const float2* psi, float2* lam,
                               size_t row_stride, int n_alloc,
                               const DevTerm* terms, int n_terms,
                               const float* downstream, int n_ops, int rows,
                               cudaStream_t s) {
  const size_t n_amps = size_t(1) << n_alloc;
  unsigned chunks = cdiv(n_amps, size_t(kThreads) * 4);
  if (chunks > 65535) chunks = 65535;
  const dim3 grid(chunks, rows);
  accumulate_operators_kernel<<<grid, kThreads, 0, s>>>(
      psi, lam, row_stride, n_amps, terms, n_terms, downstream, n_ops);
}

void LaunchReduceGradSlots(const double* slot_vals, const int32_t* slot_col,
                           int n_slots, int rows, float* grads, int n_cols,
                           cudaStream_t s) {
  if (rows * n_cols == 0) return;
  reduce_grad_slots_kernel<<<cdiv(size_t(rows) * n_cols, 128), 128, 0, s>>>(
      slot_vals, slot_col, n_slots, rows, grads, n_cols);
}

static int tree_bits(int n_alloc) {
  return n_alloc < kTreeChunkBits ? kTreeChunkBits : n_alloc;
}

size_t TreeDoublesPerRow(int n_alloc) {
  const int nc = tree_bits(n_alloc);
  return (size_t(1) << (nc - kTreeChunkBits + 1));
}

void LaunchBuildTree(const float2* psi, size_t row_stride, int n_alloc,
                     double* tree, int rows, cudaStream_t s) {
  const int nc = tree_bits(n_alloc);
  const size_t stride = TreeDoublesPerRow(n_alloc);
  const size_t chunks = size_t(1) << (nc - kTreeChunkBits);
  {
    const dim3 grid(cdiv(chunks, kThreads / 32), rows);
    tree_leaves_kernel<<<grid, kThreads, 0, s>>>(
        psi, row_stride, size_t(1) << n_alloc, nc, tree, stride);
  }
  int lvl = kTreeChunkBits;
  while (lvl < nc) {
    const int n_levels = (nc - lvl) < 9 ? (nc - lvl) : 9;
    const size_t n_in = size_t(1) << (nc - lvl);
    const dim3 grid(cdiv(n_in, 2 * kThreads), rows);
    tree_upper_kernel<<<grid, kThreads, 0, s>>>(tree, stride, nc, lvl, n_levels);
    lvl += n_levels;
  }
}

void LaunchSample(const float2* psi, size_t row_stride, int n_alloc,
                  const double* tree, const double* uniforms,
                  size_t uniform_row_stride, const int32_t* shots_per_row,
                  int shots, int rows, uint64_t* indices,
                  size_t index_row_stride, cudaStream_t s) {
  if (shots == 0 || rows == 0) return;
  const int nc = tree_bits(n_alloc);
  const dim3 grid(cdiv(size_t(shots), kThreads / 32), rows);
  sample_kernel<<<grid, kThreads, 0, s>>>(
      psi, row_stride, size_t(1) << n_alloc, nc, tree,
      TreeDoublesPerRow(n_alloc), uniforms, uniform_row_stride, shots_per_row,
      shots, reinterpret_cast<unsigned long long*>(indices), index_row_stride);
}

void LaunchFillUniforms(double* u, size_t row_stride, uint64_t seed,
                        const int32_t* row_ids, uint32_t stream_a,
                        uint32_t stream_b, int shots, int rows, cudaStream_t s) {
  if (rows == 0 || row_stride == 0) return;
  const dim3 grid(cdiv(row_stride, 256), rows);
  fill_uniforms_kernel<<<grid, 256, 0, s>>>(u, row_stride, seed, row_ids,
                                            stream_a, stream_b, shots, row_stride);
}

void LaunchSortRows(double* u, size_t row_stride, int rows, cudaStream_t s) {
  if (rows == 0 || row_stride < 2) return;
  sort_rows_kernel<<<rows, 1024, 0, s>>>(u, row_stride, row_stride);
}

void LaunchUnpackSamples(const uint64_t* indices, size_t index_row_stride,
                         int n, int nmax, int shots, int rows, int8_t* out,
                         cudaStream_t s) {
  if (rows == 0 || shots == 0 || nmax == 0) return;
  const dim3 grid(cdiv(size_t(shots) * nmax, 256), rows);
  unpack_samples_kernel<<<grid, 256, 0, s>>>(
      reinterpret_cast<const unsigned long long*>(indices), index_row_stride, n,
      nmax, shots, out);
}

void LaunchParityExpectation(const uint64_t* indices, size_t index_row_stride,
                             uint64_t mask, float coeff,
                             const int32_t* shots_per_row, int shots, int rows,
                             float* acc, size_t acc_stride, cudaStream_t s) {
  if (rows == 0) return;
  parity_expectation_kernel<<<rows, kThreads, 0, s>>>(
      reinterpret_cast<const unsigned long long*>(indices), index_row_stride,
      mask, coeff, shots_per_row, shots, acc, acc_stride);
}

void LaunchAddConstant(float c, int rows, float* acc, size_t acc_stride,
                       cudaStream_t s) {
  if (rows == 0) return;
  add_constant_kernel<<<cdiv(size_t(rows), 128), 128, 0, s>>>(c, rows, acc, acc_stride);
}

}  // namespace tfqb
