// Hand-written sm_100a kernels for the TFQ state-vector hot path.
// See kernels.cuh for the reference call sites each launch replaces.
#include "kernels.cuh"

#include <cassert>
#include <cstdio>
#include <cstdlib>

#include "gates.cuh"

namespace tfqb {
namespace {

constexpr int kThreads = 256;
constexpr int kFwdGroups = 2;   // register groups per thread, forward pass
constexpr int kAdjGroups = 1;   // adjoint pass (psi and lambda groups)
constexpr unsigned kFull = 0xffffffffu;

#include "pass_device.cuh"

// ---- slow path (controlled gates): scalar arithmetic, runtime masks ----------
template <int R, int J>
__device__ __forceinline__ void apply_g1_ctrl(float2 (&a)[1 << R], const float2 (&m)[4],
                                              uint32_t cm, uint32_t cb) {
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    if ((e & cm) != cb) continue;
    const float2 a0 = a[e], a1 = a[e | (1 << J)];
    a[e] = cfma(m[1], a1, cmulf(m[0], a0));
    a[e | (1 << J)] = cfma(m[3], a1, cmulf(m[2], a0));
  }
}
template <int R, int B0, int B1>
__device__ __forceinline__ void apply_g2_ctrl(float2 (&a)[1 << R], const float2 (&m)[16],
                                              uint32_t cm, uint32_t cb) {
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & ((1 << B0) | (1 << B1))) continue;
    if ((e & cm) != cb) continue;
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
template <int R, int J>
__device__ __forceinline__ float grad_g1_ctrl(const float2 (&a)[1 << R],
                                              const float2 (&l)[1 << R],
                                              const float2 (&m)[4], uint32_t cm,
                                              uint32_t cb) {
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & (1 << J)) continue;
    if ((e & cm) != cb) continue;
    const float2 a0 = a[e], a1 = a[e | (1 << J)];
    const float2 p0 = cfma(m[1], a1, cmulf(m[0], a0));
    const float2 p1 = cfma(m[3], a1, cmulf(m[2], a0));
    acc += redot(l[e], p0) + redot(l[e | (1 << J)], p1);
  }
  return acc;
}
template <int R, int B0, int B1>
__device__ __forceinline__ float grad_g2_ctrl(const float2 (&a)[1 << R],
                                              const float2 (&l)[1 << R],
                                              const float2 (&m)[16], uint32_t cm,
                                              uint32_t cb) {
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if (e & ((1 << B0) | (1 << B1))) continue;
    if ((e & cm) != cb) continue;
    const int i1 = e | (1 << B1), i2 = e | (1 << B0), i3 = i1 | i2;
    const float2 a0 = a[e], a1 = a[i1], a2 = a[i2], a3 = a[i3];
    acc += redot(l[e], cfma(m[3], a3, cfma(m[2], a2, cfma(m[1], a1, cmulf(m[0], a0)))));
    acc += redot(l[i1], cfma(m[7], a3, cfma(m[6], a2, cfma(m[5], a1, cmulf(m[4], a0)))));
    acc += redot(l[i2], cfma(m[11], a3, cfma(m[10], a2, cfma(m[9], a1, cmulf(m[8], a0)))));
    acc += redot(l[i3], cfma(m[15], a3, cfma(m[14], a2, cfma(m[13], a1, cmulf(m[12], a0)))));
  }
  return acc;
}
template <int R>
__device__ __forceinline__ void diag_generic(float2 (&a)[1 << R], const float4* __restrict__ sm,
                                             uint32_t rm0, uint32_t rm1, int w0,
                                             int selbase, uint32_t cm, uint32_t cb) {
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if ((e & cm) != cb) continue;
    const int sel = selbase + ((e & rm0) ? w0 : 0) + ((e & rm1) ? 1 : 0);
    a[e] = cmulf(a[e], plain(sm[sel]));
  }
}
template <int R>
__device__ __forceinline__ float gdiag_generic(const float2 (&a)[1 << R],
                                               const float2 (&l)[1 << R],
                                               const float4* __restrict__ sm, uint32_t rm0,
                                               uint32_t rm1, int w0, int selbase,
                                               uint32_t cm, uint32_t cb) {
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < (1 << R); ++e) {
    if ((e & cm) != cb) continue;
    const int sel = selbase + ((e & rm0) ? w0 : 0) + ((e & rm1) ? 1 : 0);
    acc += redot(l[e], cmulf(a[e], plain(sm[sel])));
  }
  return acc;
}

// Everything with controls: full OpRec decode, runtime masks.
template <int R, bool ADJ>
__device__ __noinline__ float slow_op(float2 (&a)[1 << R], float2 (&l)[ADJ ? (1 << R) : 1],
                                      const OpRec& op, const float4* __restrict__ sm,
                                      unsigned long long gbase, bool active) {
  const uint32_t cm = op.creg_mask, cb = op.creg_bits;
  const bool rest_ok = active && ((gbase & op.crest_mask) == op.crest_bits);
  const int tgt = ADJ ? op.target : kTgtPsi;
  const int kind = op.kind;
  float v = 0.f;
  if (!rest_ok) return 0.f;
  if (kind == kOpG1 || kind == kOpGrad1) {
    float2 m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = plain(sm[k]);
#define TFQB_G1_CASE(J)                                                      \
  if (kind == kOpG1) {                                                       \
    if (tgt & kTgtPsi) apply_g1_ctrl<R, J>(a, m, cm, cb);                    \
    if constexpr (ADJ) { if (tgt & kTgtLam) apply_g1_ctrl<R, J>(l, m, cm, cb); } \
  } else if constexpr (ADJ) {                                                \
    v = grad_g1_ctrl<R, J>(a, l, m, cm, cb);                                 \
  }
    switch (op.b0) {
      case 0: TFQB_G1_CASE(0) break;
      case 1: TFQB_G1_CASE(1) break;
      case 2: TFQB_G1_CASE(2) break;
      default: if constexpr (R > 3) { TFQB_G1_CASE(3) } break;
    }
#undef TFQB_G1_CASE
  } else if (kind == kOpG2 || kind == kOpGrad2) {
    float2 m[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m[k] = plain(sm[k]);
#define TFQB_G2_CASE(B0, B1)                                                 \
  if (kind == kOpG2) {                                                       \
    if (tgt & kTgtPsi) apply_g2_ctrl<R, B0, B1>(a, m, cm, cb);               \
    if constexpr (ADJ) { if (tgt & kTgtLam) apply_g2_ctrl<R, B0, B1>(l, m, cm, cb); } \
  } else if constexpr (ADJ) {                                                \
    v = grad_g2_ctrl<R, B0, B1>(a, l, m, cm, cb);                            \
  }
    switch (op.b0 * (op.b0 - 1) / 2 + op.b1) {
      case 0: TFQB_G2_CASE(1, 0) break;
      case 1: TFQB_G2_CASE(2, 0) break;
      case 2: TFQB_G2_CASE(2, 1) break;
      case 3: if constexpr (R > 3) { TFQB_G2_CASE(3, 0) } break;
      case 4: if constexpr (R > 3) { TFQB_G2_CASE(3, 1) } break;
      default: if constexpr (R > 3) { TFQB_G2_CASE(3, 2) } break;
    }
#undef TFQB_G2_CASE
  } else {   // diagonal / diagonal gradient
    const bool two = op.dpos1 >= 0;
    const int r0 = op.dreg0, r1 = two ? op.dreg1 : -1;
    const int c0 = r0 < 0 ? int((gbase >> op.dpos0) & 1ull) : 0;
    const int c1 = (two && r1 < 0) ? int((gbase >> op.dpos1) & 1ull) : 0;
    const int w0 = two ? 2 : 1;
    const uint32_t rm0 = r0 >= 0 ? (1u << r0) : 0u;
    const uint32_t rm1 = r1 >= 0 ? (1u << r1) : 0u;
    const int selbase = c0 * w0 + c1;
    if (kind == kOpD) {
      if (tgt & kTgtPsi) diag_generic<R>(a, sm, rm0, rm1, w0, selbase, cm, cb);
      if constexpr (ADJ) { if (tgt & kTgtLam) diag_generic<R>(l, sm, rm0, rm1, w0, selbase, cm, cb); }
    } else if constexpr (ADJ) {
      v = gdiag_generic<R>(a, l, sm, rm0, rm1, w0, selbase, cm, cb);
    }
  }
  return v;
}

// ------------------------------------------------------------------------
// The cache-blocked pass kernel (Q1). One CTA = one tile of one row.
//   smem: [psi tile][lam tile (ADJ)][expanded matrices][hi table][ops]
//         [rounds][grad acc]
// ------------------------------------------------------------------------
template <int R, int G, bool ADJ>
__global__ void __launch_bounds__(kThreads / G,
                                  (ADJ && R == 4) ? 1 : ((ADJ || G == 1) ? 2 : 3))
pass_kernel(float2* __restrict__ psi, float2* __restrict__ lam,
            size_t row_stride, const PassRec* __restrict__ passes,
            const RoundRec* __restrict__ rounds, const OpRec* __restrict__ ops,
            const float* __restrict__ mats, size_t mat_row_stride,
            int pass_index, int first_op, int n_ops_in_pass,
            double* __restrict__ grad_out, int n_slots, int init_zero_state,
            unsigned long long rank_base, const float2* const* __restrict__ peer_tab,
            int peer_shift, unsigned long long peer_self) {
  // rank_base: index bits above the local shard (state sharded over ranks by
  // its top qubits); they feed predicates and phases, never addresses.
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PassRec& P = passes[pass_index];
  const int t = P.tile_bits;
  const int L = P.low_bits;
  const uint32_t tile_size = 1u << t;
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  const size_t row = blockIdx.y;
  const int n_rounds = P.round_end - P.round_begin;

  float2* s_psi = reinterpret_cast<float2*>(smem_raw);
  float2* s_lam = s_psi + (ADJ ? tile_size : 0);
  float4* s_mat = reinterpret_cast<float4*>(s_lam + tile_size);
  const int n_entries = (P.mat_len + 1) / 2;          // complex entries
  OpRec* s_ops = reinterpret_cast<OpRec*>(s_mat + n_entries);
  unsigned long long* s_hi =
      reinterpret_cast<unsigned long long*>(s_ops + n_ops_in_pass);
  RoundRec* s_rounds = reinterpret_cast<RoundRec*>(s_hi + (1u << (t - L)));
  // gradient partials are fp64 from the warp shuffle on (the reference's
  // RealInnerProduct is double throughout, tfq_adj_grad_op.cc:272-273)
  // (s_rounds is 8-byte aligned and sizeof(RoundRec) == 24)
  double* s_grad = reinterpret_cast<double*>(s_rounds + n_rounds);

  for (uint32_t h = tid; h < (1u << (t - L)); h += nthr) {
    unsigned long long v = 0;
    for (int k = 0; k < t - L; ++k)
      v |= (unsigned long long)((h >> k) & 1u) << P.tile_pos[L + k];
    s_hi[h] = v;
  }
  {
    const float2* src = reinterpret_cast<const float2*>(
        mats + row * mat_row_stride + P.mat_begin);
    for (int i = tid; i < n_entries; i += nthr) {
      const float2 m = src[i];
      s_mat[i] = make_float4(m.x, m.x, -m.y, m.y);
    }
    const uint32_t* osrc = reinterpret_cast<const uint32_t*>(ops + first_op);
    uint32_t* odst = reinterpret_cast<uint32_t*>(s_ops);
    const int nw = n_ops_in_pass * int(sizeof(OpRec) / 4);
    for (int i = tid; i < nw; i += nthr) odst[i] = osrc[i];
    const uint32_t* rsrc = reinterpret_cast<const uint32_t*>(rounds + P.round_begin);
    uint32_t* rdst = reinterpret_cast<uint32_t*>(s_rounds);
    const int nr = n_rounds * int(sizeof(RoundRec) / 4);
    for (int i = tid; i < nr; i += nthr) rdst[i] = rsrc[i];
  }
  // gradient partials: one fp64 slot per (op, warp), written without atomics
  const int grad_slots = nthr >> 5;
  if (ADJ)
    for (int i = tid; i < n_ops_in_pass * grad_slots; i += nthr) s_grad[i] = 0.0;
  __syncthreads();

  const uint32_t lowmask = (1u << L) - 1u;
  float2* g_psi = psi + row * row_stride;
  float2* g_lam = ADJ ? lam + row * row_stride : nullptr;

  // A CTA works through several tiles of its row (grid-stride): the prologue
  // above -- op / round / matrix staging, the scatter table -- is paid once.
  const unsigned long long n_tiles = 1ull << P.n_comp;
  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
  // tile base: scatter the tile id over the non-tile bit positions
  unsigned long long base = 0;
  for (int k = 0; k < P.n_comp; ++k) base |= ((tile >> k) & 1ull) << P.comp_pos[k];

  if (!ADJ && init_zero_state == 2) {
    // ---- pass 0 of a forward plan: synthesise the product state
    // prod_b u_b[i_b] left by the leading 1-qubit gates (plan.cc
    // extract_product_init) instead of loading |0..0> and applying them
    const float4* iv = s_mat + (P.init_off >> 1);
    const unsigned long long fb = base | rank_base;
    float2 C = make_float2(1.f, 0.f);
    for (int k = 0; k < P.n_comp; ++k) {
      const int b = P.comp_pos[k];
      C = cmulf(C, plain(iv[2 * b + int((fb >> b) & 1ull)]));
    }
    for (int b = P.n_comp + t; b < P.init_bits; ++b)   // rank bits
      C = cmulf(C, plain(iv[2 * b + int((fb >> b) & 1ull)]));
    for (uint32_t blk = tid; blk < tile_size / 16; blk += nthr) {
      float2 T[16];
      T[0] = C;
      for (int k = 4; k < t; ++k)
        T[0] = cmulf(T[0], plain(iv[2 * P.tile_pos[k] + int((blk >> (k - 4)) & 1u)]));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 u0 = plain(iv[2 * P.tile_pos[k]]);
        const float2 u1 = plain(iv[2 * P.tile_pos[k] + 1]);
#pragma unroll
        for (int e = 0; e < (1 << k); ++e) {
          const float2 lo = T[e];
          T[e] = cmulf(lo, u0);
          T[e | (1 << k)] = cmulf(lo, u1);
        }
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) s_psi[swz(blk * 16 + e)] = T[e];
    }
  } else {
    // ---- load the tile: 16-byte vectors, 2^L*8-byte contiguous runs
    for (uint32_t c0 = 0; c0 < tile_size / 2; c0 += nthr * 4) {
      float4 v[4], w[4];
  #pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t c = c0 + u * nthr + tid;
        if (c < tile_size / 2) {
          const uint32_t i = 2 * c;
          const unsigned long long g = base | (i & lowmask) | s_hi[i >> L];
          if (init_zero_state == 3) {
            // the qubit swap fused into this load: the tile comes from the
            // shard of rank (g >> peer_shift), over NVLink
            const float2* src = peer_tab[g >> peer_shift] +
                                (peer_self | (g & ((1ull << peer_shift) - 1ull)));
            v[u] = __ldcs(reinterpret_cast<const float4*>(src));
          } else if (init_zero_state) {
            v[u] = make_float4((g | rank_base) == 0 ? 1.f : 0.f, 0.f, 0.f, 0.f);
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
  }
  __syncthreads();

  // ---- rounds.  Each thread owns G register groups of 2^R amplitudes and
  // decodes every op once for all of them.
  const uint32_t ngroups = tile_size >> R;
  const uint32_t per_iter = uint32_t(nthr) * G;
  const uint32_t iters = (ngroups + per_iter - 1) / per_iter;
  for (int r = 0; r < n_rounds; ++r) {
    const RoundRec rr = s_rounds[r];
    uint32_t o[R], so[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      o[j] = 1u << rr.pos[j];
      so[j] = swz(o[j]);     // swz is GF(2)-linear: swz(b|off) = swz(b)^swz(off)
    }
    for (uint32_t it = 0; it < iters; ++it) {
      float2 a[G][1 << R];
      float2 l[G][ADJ ? (1 << R) : 1];
      uint32_t sb[G];
      unsigned long long gbase[G];
      bool active[G];
      // forward kernel: scalar phase of the diagonal ops that touch no
      // register bit of this round; applied once at the end of the round
      float2 ph[G];
      bool ph_dirty = false;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const uint32_t gi = it * per_iter + uint32_t(g) * nthr + tid;
        active[g] = gi < ngroups;
        uint32_t b = active[g] ? gi : 0;
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const uint32_t lo = o[j] - 1u;
          b = ((b & ~lo) << 1) | (b & lo);
        }
        sb[g] = swz(b);
#pragma unroll
        for (int e = 0; e < (1 << R); ++e) {
          uint32_t x = sb[g];
#pragma unroll
          for (int j = 0; j < R; ++j)
            if (e & (1 << j)) x ^= so[j];
          a[g][e] = s_psi[x];
          if constexpr (ADJ) l[g][e] = s_lam[x];
        }
        gbase[g] = rank_base | base | (b & lowmask) | s_hi[b >> L];
        ph[g] = make_float2(1.f, 0.f);
      }

      // hot word of the next op is fetched while the current one executes
      int4 w_next = *reinterpret_cast<const int4*>(&s_ops[rr.op_begin - first_op]);
      for (int oi = rr.op_begin - first_op; oi < rr.op_end - first_op; ++oi) {
        const int4 w0 = w_next;
        if (oi + 1 < rr.op_end - first_op)
          w_next = *reinterpret_cast<const int4*>(&s_ops[oi + 1]);
        const int code = w0.x;
        const float4* sm = s_mat + (w0.y >> 1);
        const int tgt = ADJ ? w0.z : kTgtPsi;
        float gv = 0.f;       // gradient contribution of this thread
        bool is_grad = false;
// apply F(args...) to psi and/or lambda of every group
#define TFQB_APPLY(F, ...)                                        \
  do {                                                            \
    _Pragma("unroll") for (int g = 0; g < G; ++g) {               \
      if (tgt & kTgtPsi) F(a[g], __VA_ARGS__);                    \
      if constexpr (ADJ) { if (tgt & kTgtLam) F(l[g], __VA_ARGS__); } \
    }                                                             \
  } while (0)
#define TFQB_GRAD(F, ...)                                         \
  do {                                                            \
    if constexpr (ADJ) {                                          \
      _Pragma("unroll") for (int g = 0; g < G; ++g)               \
        if (active[g]) gv += F(a[g], l[g], __VA_ARGS__);          \
      is_grad = true;                                             \
    }                                                             \
  } while (0)
        switch (code) {
          case kCodeG1 + 0: TFQB_APPLY((g1_packed<R, 0>), sm); break;
          case kCodeG1 + 1: TFQB_APPLY((g1_packed<R, 1>), sm); break;
          case kCodeG1 + 2: TFQB_APPLY((g1_packed<R, 2>), sm); break;
          case kCodeG1 + 3: if constexpr (R > 3) TFQB_APPLY((g1_packed<R, 3>), sm); break;
          case kCodeG2 + 0: TFQB_APPLY((g2_packed<R, 1, 0>), sm); break;
          case kCodeG2 + 1: TFQB_APPLY((g2_packed<R, 2, 0>), sm); break;
          case kCodeG2 + 2: TFQB_APPLY((g2_packed<R, 2, 1>), sm); break;
          case kCodeG2 + 3: if constexpr (R > 3) TFQB_APPLY((g2_packed<R, 3, 0>), sm); break;
          case kCodeG2 + 4: if constexpr (R > 3) TFQB_APPLY((g2_packed<R, 3, 1>), sm); break;
          case kCodeG2 + 5: if constexpr (R > 3) TFQB_APPLY((g2_packed<R, 3, 2>), sm); break;
          case kCodeD0:
          case kCodeGradD0: {
            const int4 w1 = *(reinterpret_cast<const int4*>(&s_ops[oi]) + 1);
#pragma unroll
            for (int g = 0; g < G; ++g) {
              int sel = int((gbase[g] >> w1.z) & 1ull);
              if (w1.w >= 0) sel = 2 * sel + int((gbase[g] >> w1.w) & 1ull);
              const float4 f = sm[sel];
              if (code == kCodeD0) {
                if constexpr (ADJ) {
                  if (tgt & kTgtPsi) scale_all<R>(a[g], f);
                  if (tgt & kTgtLam) scale_all<R>(l[g], f);
                } else {
                  ph[g] = cmulf(ph[g], plain(f));
                }
              } else if constexpr (ADJ) {
                if (active[g]) gv += gdiag0<R>(a[g], l[g], f);
              }
            }
            if (code == kCodeD0) ph_dirty = true;
            else is_grad = true;
            break;
          }
          case kCodeD1 + 0: case kCodeD1 + 1: case kCodeD1 + 2: case kCodeD1 + 3:
          case kCodeGradD1 + 0: case kCodeGradD1 + 1: case kCodeGradD1 + 2:
          case kCodeGradD1 + 3: {
            const int4 w1 = *(reinterpret_cast<const int4*>(&s_ops[oi]) + 1);
            const uint32_t ident = *reinterpret_cast<const uint32_t*>(
                reinterpret_cast<const int4*>(&s_ops[oi]) + 2);
            const bool grad = code >= kCodeGradD1;
            const int j = grad ? code - kCodeGradD1 : code - kCodeD1;
            bool do0 = true, do1 = true;
            if (w1.w < 0) {            // 1-qubit diagonal on the register bit
              do0 = !(ident & 1u);
              do1 = !(ident & 2u);
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
              int s0, s1;
              if (w1.w < 0) {
                s0 = 0; s1 = 1;
              } else if (w1.x >= 0) {    // register bit is the selector msb
                const int c1 = int((gbase[g] >> w1.w) & 1ull);
                s0 = c1; s1 = 2 + c1;
              } else {                   // register bit is the selector lsb
                const int c0 = int((gbase[g] >> w1.z) & 1ull);
                s0 = 2 * c0; s1 = 2 * c0 + 1;
              }
              const float4 f0 = sm[s0], f1 = sm[s1];
              if (!grad) {
#define TFQB_D1(J)                                                          \
  if (tgt & kTgtPsi) diag1<R, J>(a[g], f0, f1, do0, do1);                  \
  if constexpr (ADJ) { if (tgt & kTgtLam) diag1<R, J>(l[g], f0, f1, do0, do1); }
                switch (j) {
                  case 0: TFQB_D1(0) break;
                  case 1: TFQB_D1(1) break;
                  case 2: TFQB_D1(2) break;
                  default: if constexpr (R > 3) { TFQB_D1(3) } break;
                }
#undef TFQB_D1
              } else if constexpr (ADJ) {
                float v = 0.f;
                switch (j) {
                  case 0: v = gdiag1<R, 0>(a[g], l[g], f0, f1); break;
                  case 1: v = gdiag1<R, 1>(a[g], l[g], f0, f1); break;
                  case 2: v = gdiag1<R, 2>(a[g], l[g], f0, f1); break;
                  default: if constexpr (R > 3) v = gdiag1<R, 3>(a[g], l[g], f0, f1); break;
                }
                if (active[g]) gv += v;
              }
            }
            if (grad) is_grad = true;
            break;
          }
#define TFQB_D2_CASE(IDX, JH, JL)                                                  \
  case kCodeD2 + IDX: {                                                            \
    if constexpr (R > JH) {                                                        \
      const uint32_t ident = *reinterpret_cast<const uint32_t*>(                   \
          reinterpret_cast<const int4*>(&s_ops[oi]) + 2);                          \
      TFQB_APPLY((diag2<R, JH, JL>), sm, ident);                                   \
    }                                                                              \
    break;                                                                         \
  }                                                                                \
  case kCodeGradD2 + IDX:                                                          \
    if constexpr (R > JH) TFQB_GRAD((gdiag2<R, JH, JL>), sm);                      \
    break;
          TFQB_D2_CASE(0, 1, 0)
          TFQB_D2_CASE(1, 2, 0)
          TFQB_D2_CASE(2, 2, 1)
          TFQB_D2_CASE(3, 3, 0)
          TFQB_D2_CASE(4, 3, 1)
          TFQB_D2_CASE(5, 3, 2)
#undef TFQB_D2_CASE
          case kCodeGrad1 + 0: TFQB_GRAD((grad1_packed<R, 0>), sm); break;
          case kCodeGrad1 + 1: TFQB_GRAD((grad1_packed<R, 1>), sm); break;
          case kCodeGrad1 + 2: TFQB_GRAD((grad1_packed<R, 2>), sm); break;
          case kCodeGrad1 + 3: if constexpr (R > 3) TFQB_GRAD((grad1_packed<R, 3>), sm); break;
          case kCodeGrad2 + 0: TFQB_GRAD((grad2_packed<R, 1, 0>), sm); break;
          case kCodeGrad2 + 1: TFQB_GRAD((grad2_packed<R, 2, 0>), sm); break;
          case kCodeGrad2 + 2: TFQB_GRAD((grad2_packed<R, 2, 1>), sm); break;
          case kCodeGrad2 + 3: if constexpr (R > 3) TFQB_GRAD((grad2_packed<R, 3, 0>), sm); break;
          case kCodeGrad2 + 4: if constexpr (R > 3) TFQB_GRAD((grad2_packed<R, 3, 1>), sm); break;
          case kCodeGrad2 + 5: if constexpr (R > 3) TFQB_GRAD((grad2_packed<R, 3, 2>), sm); break;
          case kCodeS0: {
            const int4 w1 = *(reinterpret_cast<const int4*>(&s_ops[oi]) + 1);
            const uint32_t mask = *reinterpret_cast<const uint32_t*>(
                reinterpret_cast<const int4*>(&s_ops[oi]) + 2);
#pragma unroll
            for (int g = 0; g < G; ++g) {
              int sel = int((gbase[g] >> w1.z) & 1ull);
              if (w1.w >= 0) sel = 2 * sel + int((gbase[g] >> w1.w) & 1ull);
              if ((mask >> sel) & 1u) {
                if constexpr (ADJ) {
                  if (tgt & kTgtPsi) sign_all<R>(a[g]);
                  if (tgt & kTgtLam) sign_all<R>(l[g]);
                } else {
                  ph[g] = cneg2(ph[g]);
                }
              }
            }
            if (!ADJ) ph_dirty = true;
            break;
          }
          case kCodeS1 + 0: case kCodeS1 + 1: case kCodeS1 + 2: case kCodeS1 + 3: {
            const int4 w1 = *(reinterpret_cast<const int4*>(&s_ops[oi]) + 1);
            const uint32_t mask = *reinterpret_cast<const uint32_t*>(
                reinterpret_cast<const int4*>(&s_ops[oi]) + 2);
            const int j = code - kCodeS1;
#pragma unroll
            for (int g = 0; g < G; ++g) {
              int s0, s1;
              if (w1.w < 0) {
                s0 = 0; s1 = 1;
              } else if (w1.x >= 0) {
                const int c1 = int((gbase[g] >> w1.w) & 1ull);
                s0 = c1; s1 = 2 + c1;
              } else {
                const int c0 = int((gbase[g] >> w1.z) & 1ull);
                s0 = 2 * c0; s1 = 2 * c0 + 1;
              }
              const bool n0 = (mask >> s0) & 1u, n1 = (mask >> s1) & 1u;
#define TFQB_S1(J)                                                  \
  if (tgt & kTgtPsi) sign1<R, J>(a[g], n0, n1);                     \
  if constexpr (ADJ) { if (tgt & kTgtLam) sign1<R, J>(l[g], n0, n1); }
              switch (j) {
                case 0: TFQB_S1(0) break;
                case 1: TFQB_S1(1) break;
                case 2: TFQB_S1(2) break;
                default: if constexpr (R > 3) { TFQB_S1(3) } break;
              }
#undef TFQB_S1
            }
            break;
          }
#define TFQB_S2_CASE(IDX, JH, JL)                                                  \
  case kCodeS2 + IDX: {                                                            \
    if constexpr (R > JH) {                                                        \
      const uint32_t mask = *reinterpret_cast<const uint32_t*>(                    \
          reinterpret_cast<const int4*>(&s_ops[oi]) + 2);                          \
      TFQB_APPLY((sign2<R, JH, JL>), mask);                                        \
    }                                                                              \
    break;                                                                         \
  }
          TFQB_S2_CASE(0, 1, 0)
          TFQB_S2_CASE(1, 2, 0)
          TFQB_S2_CASE(2, 2, 1)
          TFQB_S2_CASE(3, 3, 0)
          TFQB_S2_CASE(4, 3, 1)
          TFQB_S2_CASE(5, 3, 2)
#undef TFQB_S2_CASE
#define TFQB_ADJ(F, ...)                                          \
  do {                                                            \
    if constexpr (ADJ) {                                          \
      _Pragma("unroll") for (int g = 0; g < G; ++g) {             \
        const float v_ = F(a[g], l[g], __VA_ARGS__);              \
        if (active[g]) gv += v_;                                  \
      }                                                           \
      is_grad = true;                                             \
    }                                                             \
  } while (0)
          case kCodeAdj1 + 0: TFQB_ADJ((adj1_packed<R, 0>), sm); break;
          case kCodeAdj1 + 1: TFQB_ADJ((adj1_packed<R, 1>), sm); break;
          case kCodeAdj1 + 2: TFQB_ADJ((adj1_packed<R, 2>), sm); break;
          case kCodeAdj1 + 3: if constexpr (R > 3) TFQB_ADJ((adj1_packed<R, 3>), sm); break;
          case kCodeAdj2 + 0: TFQB_ADJ((adj2_packed<R, 1, 0>), sm); break;
          case kCodeAdj2 + 1: TFQB_ADJ((adj2_packed<R, 2, 0>), sm); break;
          case kCodeAdj2 + 2: TFQB_ADJ((adj2_packed<R, 2, 1>), sm); break;
          case kCodeAdj2 + 3: if constexpr (R > 3) TFQB_ADJ((adj2_packed<R, 3, 0>), sm); break;
          case kCodeAdj2 + 4: if constexpr (R > 3) TFQB_ADJ((adj2_packed<R, 3, 1>), sm); break;
          case kCodeAdj2 + 5: if constexpr (R > 3) TFQB_ADJ((adj2_packed<R, 3, 2>), sm); break;
          case kCodeAdjD0: {
            if constexpr (ADJ) {
              const int4 w1 = *(reinterpret_cast<const int4*>(&s_ops[oi]) + 1);
#pragma unroll
              for (int g = 0; g < G; ++g) {
                int sel = int((gbase[g] >> w1.z) & 1ull);
                if (w1.w >= 0) sel = 2 * sel + int((gbase[g] >> w1.w) & 1ull);
                const float v_ = adjd0<R>(a[g], l[g], sm[sel], sm[4 + sel]);
                if (active[g]) gv += v_;
              }
              is_grad = true;
            }
            break;
          }
          case kCodeAdjD1 + 0: case kCodeAdjD1 + 1: case kCodeAdjD1 + 2:
          case kCodeAdjD1 + 3: {
            if constexpr (ADJ) {
              const int4 w1 = *(reinterpret_cast<const int4*>(&s_ops[oi]) + 1);
              const int j = code - kCodeAdjD1;
#pragma unroll
              for (int g = 0; g < G; ++g) {
                int s0, s1;
                if (w1.w < 0) {
                  s0 = 0; s1 = 1;
                } else if (w1.x >= 0) {
                  const int c1 = int((gbase[g] >> w1.w) & 1ull);
                  s0 = c1; s1 = 2 + c1;
                } else {
                  const int c0 = int((gbase[g] >> w1.z) & 1ull);
                  s0 = 2 * c0; s1 = 2 * c0 + 1;
                }
                const float4 f0 = sm[s0], f1 = sm[s1], g0 = sm[4 + s0], g1 = sm[4 + s1];
                float v_ = 0.f;
                switch (j) {
                  case 0: v_ = adjd1<R, 0>(a[g], l[g], f0, f1, g0, g1); break;
                  case 1: v_ = adjd1<R, 1>(a[g], l[g], f0, f1, g0, g1); break;
                  case 2: v_ = adjd1<R, 2>(a[g], l[g], f0, f1, g0, g1); break;
                  default: if constexpr (R > 3) v_ = adjd1<R, 3>(a[g], l[g], f0, f1, g0, g1); break;
                }
                if (active[g]) gv += v_;
              }
              is_grad = true;
            }
            break;
          }
          case kCodeAdjD2 + 0: TFQB_ADJ((adjd2<R, 1, 0>), sm); break;
          case kCodeAdjD2 + 1: TFQB_ADJ((adjd2<R, 2, 0>), sm); break;
          case kCodeAdjD2 + 2: TFQB_ADJ((adjd2<R, 2, 1>), sm); break;
          case kCodeAdjD2 + 3: if constexpr (R > 3) TFQB_ADJ((adjd2<R, 3, 0>), sm); break;
          case kCodeAdjD2 + 4: if constexpr (R > 3) TFQB_ADJ((adjd2<R, 3, 1>), sm); break;
          case kCodeAdjD2 + 5: if constexpr (R > 3) TFQB_ADJ((adjd2<R, 3, 2>), sm); break;
#undef TFQB_ADJ
          case kCodeG1Run: {
            const uint32_t bits = *reinterpret_cast<const uint32_t*>(
                reinterpret_cast<const int4*>(&s_ops[oi]) + 2);
            const float4* m = sm;
            if constexpr (R > 3) {
              if (bits & 8u) { TFQB_APPLY((g1_packed<R, 3>), m); m += 4; }
            }
            if (bits & 4u) { TFQB_APPLY((g1_packed<R, 2>), m); m += 4; }
            if (bits & 2u) { TFQB_APPLY((g1_packed<R, 1>), m); m += 4; }
            if (bits & 1u) { TFQB_APPLY((g1_packed<R, 0>), m); }
            break;
          }
          case kCodeS0Run: {
            const int4 w1 = *(reinterpret_cast<const int4*>(&s_ops[oi]) + 1);
            const uint32_t c0 = *reinterpret_cast<const uint32_t*>(
                reinterpret_cast<const int4*>(&s_ops[oi]) + 2);
            const ulonglong2 w3 = *reinterpret_cast<const ulonglong2*>(
                reinterpret_cast<const int4*>(&s_ops[oi]) + 3);
            const ulonglong2 w4 = *reinterpret_cast<const ulonglong2*>(
                reinterpret_cast<const int4*>(&s_ops[oi]) + 4);
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const unsigned long long gb = gbase[g];
              const int par = int(c0) + __popcll(gb & w3.y) +
                              __popcll(gb & (gb >> w1.z) & w4.x) +
                              __popcll(gb & (gb >> w1.w) & w4.y);
              if (par & 1) {
                if constexpr (ADJ) {
                  if (tgt & kTgtPsi) sign_all<R>(a[g]);
                  if (tgt & kTgtLam) sign_all<R>(l[g]);
                } else {
                  ph[g] = cneg2(ph[g]);
                }
              }
            }
            if (!ADJ) ph_dirty = true;
            break;
          }
          default: {   // kCodeSlow
            const int kind = s_ops[oi].kind;
            is_grad = kind == kOpGrad1 || kind == kOpGrad2 || kind == kOpGradD;
#pragma unroll
            for (int g = 0; g < G; ++g) {
              if (!ADJ && ph_dirty) scale_all_c<R>(a[g], ph[g]);
              ph[g] = make_float2(1.f, 0.f);
              // copies keep a[] / l[] in registers outside this rare path
              float2 ta[1 << R];
              float2 tl[ADJ ? (1 << R) : 1];
#pragma unroll
              for (int e = 0; e < (1 << R); ++e) {
                ta[e] = a[g][e];
                if constexpr (ADJ) tl[e] = l[g][e];
              }
              gv += slow_op<R, ADJ>(ta, tl, s_ops[oi], sm, gbase[g], active[g]);
#pragma unroll
              for (int e = 0; e < (1 << R); ++e) {
                a[g][e] = ta[e];
                if constexpr (ADJ) l[g][e] = tl[e];
              }
            }
            ph_dirty = false;
            break;
          }
        }
#undef TFQB_APPLY
#undef TFQB_GRAD
        if constexpr (ADJ) {
          if (is_grad) {      // uniform across the CTA
            double gd = double(gv);   // per-thread float over <= 16 amplitudes
            gd += __shfl_xor_sync(kFull, gd, 16);
            gd += __shfl_xor_sync(kFull, gd, 8);
            gd += __shfl_xor_sync(kFull, gd, 4);
            gd += __shfl_xor_sync(kFull, gd, 2);
            gd += __shfl_xor_sync(kFull, gd, 1);
            if ((tid & 31) == 0)    // this warp owns the slot: no atomics
              s_grad[oi * grad_slots + (tid >> 5)] += 2.0 * gd;
          }
        }
      }

#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (!active[g]) continue;
        if (!ADJ && ph_dirty) scale_all_c<R>(a[g], ph[g]);
#pragma unroll
        for (int e = 0; e < (1 << R); ++e) {
          uint32_t x = sb[g];
#pragma unroll
          for (int j = 0; j < R; ++j)
            if (e & (1 << j)) x ^= so[j];
          s_psi[x] = a[g][e];
          if constexpr (ADJ) s_lam[x] = l[g][e];
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
  __syncthreads();   // the tile buffers are reused by the next tile
  }
  if (ADJ) {
    for (int i = tid; i < n_ops_in_pass; i += nthr) {
      const int slot = s_ops[i].grad_slot;
      double v = 0.0;
      for (int k = 0; k < grad_slots; ++k) v += s_grad[i * grad_slots + k];
      if (slot >= 0 && v != 0.0)
        atomicAdd(&grad_out[row * size_t(n_slots) + slot], v);
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
  const int dim = (rec.layout == 0 || rec.layout == 2 || rec.layout == 4) ? 2 : 4;
  cf m[16];
  // products of several factors are carried in fp64 and rounded ONCE: a
  // float32 product has a systematic error per fused gate that accumulates
  // over a deep circuit (each factor itself is the float32 recipe, exactly)
  double mr[16], mi[16];
  bool wide = false;
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
    if (f.gate_kind == kCH)       // noise channel: the Kraus operator this row drew
      channel_matrix(p, f.aux_sym >= 0 ? params[size_t(row) * n_params + f.aux_sym] : 0.f, g);
    else if (rec.mode == kMatGrad)
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
    wide = true;
    if (fi == rec.factor_begin) {
      for (int i = 0; i < dim * dim; ++i) {
        mr[i] = double(e[i].re);
        mi[i] = double(e[i].im);
      }
    } else {
      double tr[16], ti[16];
      for (int r = 0; r < dim; ++r)
        for (int c = 0; c < dim; ++c) {
          double ar = 0.0, ai = 0.0;
          for (int k = 0; k < dim; ++k) {
            const double er = double(e[r * dim + k].re), ei = double(e[r * dim + k].im);
            ar += er * mr[k * dim + c] - ei * mi[k * dim + c];
            ai += er * mi[k * dim + c] + ei * mr[k * dim + c];
          }
          tr[r * dim + c] = ar;
          ti[r * dim + c] = ai;
        }
      for (int i = 0; i < dim * dim; ++i) {
        mr[i] = tr[i];
        mi[i] = ti[i];
      }
    }
  }
  if (wide)
    for (int i = 0; i < dim * dim; ++i) m[i] = mk(float(mr[i]), float(mi[i]));
  float* o = out + size_t(row) * out_row_stride + rec.out_off;
  const bool dag = rec.mode == kMatDagger;
  if (rec.layout == 4) {  // first column: U|0>
    o[0] = m[0].re; o[1] = m[0].im; o[2] = m[2].re; o[3] = m[2].im;
    return;
  }
  if (rec.layout >= 2) {  // diagonal: d[0..dim)
    for (int i = 0; i < 4; ++i) {
      // swap: exchange the two selector bits (entries 1 <-> 2)
      const int src = (rec.layout == 3 && rec.swap) ? (((i & 1) << 1) | (i >> 1)) : i;
      cf v = i < dim ? m[src * dim + src] : mk(0.f, 0.f);
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
// K1 fast path: tile-based PauliSum expectation (see plan.h ExpectationPlan)
//   smem: [psi tile][probability / WHT table (pass 0)][hi table][xops]
//         [rounds][zterms][per-term float accumulators]
// ------------------------------------------------------------------------
// X/Y-type partial sums are kept per thread (or per 2^part_shift lanes) in
// shared memory as plain floats: no warp reduction and no atomics in the tile
// loop.  They are folded into the fp64 per-term accumulators every
// kExpFlushTiles tiles.
constexpr int kExpFlushTiles = 32;

__global__ void __launch_bounds__(kThreads, 2)
expect_pass_kernel(const float2* __restrict__ psi, size_t row_stride,
                   const PassRec* __restrict__ passes,
                   const RoundRec* __restrict__ rounds,
                   const ExpXOp* __restrict__ xops,
                   const ExpZTerm* __restrict__ zterms, int n_zterms,
                   int pass_index, int n_terms, unsigned long long n_tiles,
                   unsigned long long rank_base, int part_shift,
                   double* __restrict__ per_term) {
  constexpr int R = 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PassRec& P = passes[pass_index];
  const int t = P.tile_bits;
  const int L = P.low_bits;
  const uint32_t tile_size = 1u << t;
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  const size_t row = blockIdx.y;
  const int n_rounds = P.round_end - P.round_begin;
  const int first_op = n_rounds ? rounds[P.round_begin].op_begin : 0;
  const int n_xops = n_rounds ? rounds[P.round_end - 1].op_end - first_op : 0;
  const bool do_z = n_zterms > 0;
  const int slots = nthr >> part_shift;

  float2* s_psi = reinterpret_cast<float2*>(smem_raw);
  float* s_p = reinterpret_cast<float*>(s_psi + tile_size);
  unsigned long long* s_hi =
      reinterpret_cast<unsigned long long*>(s_p + (do_z ? tile_size : 0));
  ExpXOp* s_x = reinterpret_cast<ExpXOp*>(s_hi + (1u << (t - L)));
  RoundRec* s_rounds = reinterpret_cast<RoundRec*>(s_x + n_xops);
  ExpZTerm* s_z = reinterpret_cast<ExpZTerm*>(s_rounds + n_rounds);
  // per-term accumulators over all tiles of this CTA: fp64, so that large
  // states (2^18+ tiles per CTA loop) do not lose the small terms
  double* s_acc = reinterpret_cast<double*>(s_z + n_zterms);
  float* s_part = reinterpret_cast<float*>(s_acc + n_terms);

  for (uint32_t h = tid; h < (1u << (t - L)); h += nthr) {
    unsigned long long v = 0;
    for (int k = 0; k < t - L; ++k)
      v |= (unsigned long long)((h >> k) & 1u) << P.tile_pos[L + k];
    s_hi[h] = v;
  }
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(xops + first_op);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_x);
    for (int i = tid; i < n_xops * int(sizeof(ExpXOp) / 4); i += nthr) dst[i] = src[i];
    src = reinterpret_cast<const uint32_t*>(rounds + P.round_begin);
    dst = reinterpret_cast<uint32_t*>(s_rounds);
    for (int i = tid; i < n_rounds * int(sizeof(RoundRec) / 4); i += nthr) dst[i] = src[i];
    src = reinterpret_cast<const uint32_t*>(zterms);
    dst = reinterpret_cast<uint32_t*>(s_z);
    for (int i = tid; i < n_zterms * int(sizeof(ExpZTerm) / 4); i += nthr) dst[i] = src[i];
  }
  for (int i = tid; i < n_terms; i += nthr) s_acc[i] = 0.0;
  for (int i = tid; i < n_xops * slots; i += nthr) s_part[i] = 0.f;
  __syncthreads();

  const uint32_t lowmask = (1u << L) - 1u;
  const float2* g_psi = psi + row * row_stride;
  const uint32_t part_mask = (1u << part_shift) - 1u;

  // fold the float partials of every X/Y op into its term (fp64)
  auto flush_partials = [&]() {
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    for (int oi = warp; oi < n_xops; oi += nwarps) {
      float* p = s_part + oi * slots;
      double v = 0.0;
      for (int k = lane; k < slots; k += 32) {
        v += double(p[k]);
        p[k] = 0.f;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
      // pairs are counted once: the mirrored half contributes the same
      if (lane == 0) atomicAdd(&s_acc[s_x[oi].term], 2.0 * v);
    }
  };

  int since_flush = 0;
  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    unsigned long long base = 0;
    {
      const int nc = P.n_comp;
      for (int k = 0; k < nc; ++k)
        base |= ((tile >> k) & 1ull) << P.comp_pos[k];
    }
    // ---- load the tile (and its probabilities): all loads of a thread are
    // in flight before the first use
    for (uint32_t c0 = 0; c0 < tile_size / 2; c0 += nthr * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t c = c0 + u * nthr + tid;
        if (c < tile_size / 2) {
          const uint32_t i = 2 * c;
          const unsigned long long g = base | (i & lowmask) | s_hi[i >> L];
          v[u] = *reinterpret_cast<const float4*>(g_psi + g);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t c = c0 + u * nthr + tid;
        if (c < tile_size / 2) {
          const uint32_t i = 2 * c;
          const uint32_t x0 = swz(i), x1 = x0 ^ 1u;    // i is even
          s_psi[x0] = make_float2(v[u].x, v[u].y);
          s_psi[x1] = make_float2(v[u].z, v[u].w);
          if (do_z) {
            s_p[x0] = fmaf(v[u].x, v[u].x, v[u].y * v[u].y);
            s_p[x1] = fmaf(v[u].z, v[u].z, v[u].w * v[u].w);
          }
        }
      }
    }
    __syncthreads();

    // ---- Z-type terms: Walsh-Hadamard transform of the probabilities
    if (do_z) {
      for (int lvl = 0; lvl < t; lvl += 4) {
        switch (min(4, t - lvl)) {
          case 4: wht_level<4>(s_p, tile_size, lvl, tid, nthr); break;
          case 3: wht_level<3>(s_p, tile_size, lvl, tid, nthr); break;
          case 2: wht_level<2>(s_p, tile_size, lvl, tid, nthr); break;
          default: wht_level<1>(s_p, tile_size, lvl, tid, nthr); break;
        }
        __syncthreads();
      }
      for (int k = tid; k < n_zterms; k += nthr) {
        const ExpZTerm zt = s_z[k];
        float v = s_p[swz(zt.ztile)];
        const int neg = (__popcll((base | rank_base) & zt.zrest) & 1) ^ zt.negate;
        s_acc[zt.term] += double(neg ? -v : v);     // one thread owns the term
      }
    }

    // ---- X/Y-type terms: partner amplitude inside the thread's registers
    const uint32_t ngroups = tile_size >> R;
    for (int r = 0; r < n_rounds; ++r) {
      const RoundRec rr = s_rounds[r];
      uint32_t o[R], so[R];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        o[j] = 1u << rr.pos[j];
        so[j] = swz(o[j]);
      }
      const uint32_t iters = (ngroups + nthr - 1) / nthr;
      for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t gi = it * nthr + tid;
        const bool active = gi < ngroups;     // whole warps stay in the loop
        uint32_t b = active ? gi : 0;
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const uint32_t lo = o[j] - 1u;
          b = ((b & ~lo) << 1) | (b & lo);
        }
        const uint32_t sb = swz(b);
        float2 a[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          uint32_t x = sb;
#pragma unroll
          for (int j = 0; j < R; ++j)
            if (e & (1 << j)) x ^= so[j];
          a[e] = s_psi[x];
        }
        const unsigned long long gbase =
            rank_base | base | (b & lowmask) | s_hi[b >> L];
        int4 w = *reinterpret_cast<const int4*>(&s_x[rr.op_begin - first_op]);
        for (int oi = rr.op_begin - first_op; oi < rr.op_end - first_op; ++oi) {
          // {xreg, sign16, zrest lo, zrest hi} of this op; prefetch the next
          const int4 w0 = w;
          const int4 w1 = *(reinterpret_cast<const int4*>(&s_x[oi]) + 1);
          if (oi + 1 < rr.op_end - first_op)
            w = *reinterpret_cast<const int4*>(&s_x[oi + 1]);
          float v;
          switch (w1.w) {
#define TFQB_X1(ZS, IM, XR) \
  case 1 + (IM * 5 + ZS) * 15 + (XR - 1): v = xterm_fixed<XR, ZS, (IM != 0)>(a); break;
#define TFQB_X15(ZS, IM)                                                       \
  TFQB_X1(ZS, IM, 1) TFQB_X1(ZS, IM, 2) TFQB_X1(ZS, IM, 3) TFQB_X1(ZS, IM, 4)   \
  TFQB_X1(ZS, IM, 5) TFQB_X1(ZS, IM, 6) TFQB_X1(ZS, IM, 7) TFQB_X1(ZS, IM, 8)   \
  TFQB_X1(ZS, IM, 9) TFQB_X1(ZS, IM, 10) TFQB_X1(ZS, IM, 11)                   \
  TFQB_X1(ZS, IM, 12) TFQB_X1(ZS, IM, 13) TFQB_X1(ZS, IM, 14) TFQB_X1(ZS, IM, 15)
            TFQB_X15(0, 0) TFQB_X15(1, 0) TFQB_X15(2, 0) TFQB_X15(3, 0) TFQB_X15(4, 0)
            TFQB_X15(0, 1) TFQB_X15(1, 1) TFQB_X15(2, 1) TFQB_X15(3, 1) TFQB_X15(4, 1)
#undef TFQB_X15
#undef TFQB_X1
            default: {         // several register z bits: runtime signs
              float2 ri;
              const uint32_t sg = uint32_t(w0.y);
              switch (w0.x) {
                case 1: ri = xterm_pairs<1>(a, sg); break;
                case 2: ri = xterm_pairs<2>(a, sg); break;
                case 3: ri = xterm_pairs<3>(a, sg); break;
                case 4: ri = xterm_pairs<4>(a, sg); break;
                case 5: ri = xterm_pairs<5>(a, sg); break;
                case 6: ri = xterm_pairs<6>(a, sg); break;
                case 7: ri = xterm_pairs<7>(a, sg); break;
                case 8: ri = xterm_pairs<8>(a, sg); break;
                case 9: ri = xterm_pairs<9>(a, sg); break;
                case 10: ri = xterm_pairs<10>(a, sg); break;
                case 11: ri = xterm_pairs<11>(a, sg); break;
                case 12: ri = xterm_pairs<12>(a, sg); break;
                case 13: ri = xterm_pairs<13>(a, sg); break;
                case 14: ri = xterm_pairs<14>(a, sg); break;
                default: ri = xterm_pairs<15>(a, sg); break;
              }
              v = w1.x ? ri.y : ri.x;
              break;
            }
          }
          const unsigned long long zrest =
              (unsigned long long)(uint32_t(w0.z)) | ((unsigned long long)(uint32_t(w0.w)) << 32);
          if ((__popcll(gbase & zrest) & 1) ^ w1.y) v = -v;
          if (!active) v = 0.f;
          for (int d = 1; d <= int(part_mask); d <<= 1) v += __shfl_xor_sync(kFull, v, d);
          if ((uint32_t(tid) & part_mask) == 0) s_part[oi * slots + (tid >> part_shift)] += v;
        }
      }
    }
    __syncthreads();   // tile buffers are reused by the next tile
    if (++since_flush == kExpFlushTiles) {
      flush_partials();
      since_flush = 0;
      __syncthreads();
    }
  }
  if (since_flush) {
    flush_partials();
    __syncthreads();
  }

  for (int i = tid; i < n_terms; i += nthr) {
    const double v = s_acc[i];
    if (v != 0.0) atomicAdd(&per_term[row * size_t(n_terms) + i], v);
  }
}

// ------------------------------------------------------------------------
// K3 fast path: lambda = sum_j g_j sum_t c_t P_t psi from staged tiles
// (util_qsim.h:362-414).  Same plan as the expectation: Z-type and identity
// terms collapse into ONE real multiplier per amplitude, synthesised by a
// Walsh-Hadamard transform of the sparse coefficient vector; X/Y-type terms
// add the rotated partner amplitude from the thread's registers.
//   smem: [psi tile][lambda tile][coefficient table][hi][xops][rounds]
//         [zterms][per-term lead coefficients]
// ------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2)
accum_pass_kernel(const float2* __restrict__ psi, float2* __restrict__ lam,
                  size_t row_stride, const PassRec* __restrict__ passes,
                  const RoundRec* __restrict__ rounds,
                  const ExpXOp* __restrict__ xops,
                  const ExpZTerm* __restrict__ zterms, int n_zterms,
                  const DevTerm* __restrict__ terms, int n_terms,
                  const float* __restrict__ downstream, int n_ops,
                  int pass_index, int accumulate, unsigned long long n_tiles) {
  constexpr int R = 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PassRec& P = passes[pass_index];
  const int t = P.tile_bits;
  const int L = P.low_bits;
  const uint32_t tile_size = 1u << t;
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  const size_t row = blockIdx.y;
  const int n_rounds = P.round_end - P.round_begin;
  const int first_op = n_rounds ? rounds[P.round_begin].op_begin : 0;
  const int n_xops = n_rounds ? rounds[P.round_end - 1].op_end - first_op : 0;
  const bool do_z = n_zterms > 0;

  float2* s_psi = reinterpret_cast<float2*>(smem_raw);
  float2* s_out = s_psi + tile_size;
  float* s_p = reinterpret_cast<float*>(s_out + tile_size);
  unsigned long long* s_hi =
      reinterpret_cast<unsigned long long*>(s_p + (do_z ? tile_size : 0));
  ExpXOp* s_x = reinterpret_cast<ExpXOp*>(s_hi + (1u << (t - L)));
  RoundRec* s_rounds = reinterpret_cast<RoundRec*>(s_x + n_xops);
  ExpZTerm* s_z = reinterpret_cast<ExpZTerm*>(s_rounds + n_rounds);
  float* s_lead = reinterpret_cast<float*>(s_z + n_zterms);

  for (uint32_t h = tid; h < (1u << (t - L)); h += nthr) {
    unsigned long long v = 0;
    for (int k = 0; k < t - L; ++k)
      v |= (unsigned long long)((h >> k) & 1u) << P.tile_pos[L + k];
    s_hi[h] = v;
  }
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(xops + first_op);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_x);
    for (int i = tid; i < n_xops * int(sizeof(ExpXOp) / 4); i += nthr) dst[i] = src[i];
    src = reinterpret_cast<const uint32_t*>(rounds + P.round_begin);
    dst = reinterpret_cast<uint32_t*>(s_rounds);
    for (int i = tid; i < n_rounds * int(sizeof(RoundRec) / 4); i += nthr) dst[i] = src[i];
    src = reinterpret_cast<const uint32_t*>(zterms);
    dst = reinterpret_cast<uint32_t*>(s_z);
    for (int i = tid; i < n_zterms * int(sizeof(ExpZTerm) / 4); i += nthr) dst[i] = src[i];
  }
  for (int i = tid; i < n_terms; i += nthr) {
    // `leading = downstream * coefficient`, terms below 1e-5 are skipped
    // (util_qsim.h:378-383)
    const float lead = __fmul_rn(downstream[row * n_ops + terms[i].op], terms[i].coeff);
    s_lead[i] = fabsf(lead) < 1e-5f ? 0.f : lead;
  }
  __syncthreads();

  const uint32_t lowmask = (1u << L) - 1u;
  const float2* g_psi = psi + row * row_stride;
  float2* g_lam = lam + row * row_stride;

  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    unsigned long long base = 0;
    {
      const int nc = P.n_comp;
      for (int k = 0; k < nc; ++k)
        base |= ((tile >> k) & 1ull) << P.comp_pos[k];
    }
    for (uint32_t c = tid; c < tile_size / 2; c += nthr) {
      const uint32_t i = 2 * c;
      const unsigned long long g = base | (i & lowmask) | s_hi[i >> L];
      const float4 v = *reinterpret_cast<const float4*>(g_psi + g);
      s_psi[swz(i)] = make_float2(v.x, v.y);
      s_psi[swz(i + 1)] = make_float2(v.z, v.w);
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (accumulate) w = *reinterpret_cast<const float4*>(g_lam + g);
      s_out[swz(i)] = make_float2(w.x, w.y);
      s_out[swz(i + 1)] = make_float2(w.z, w.w);
      if (do_z) {
        s_p[swz(i)] = 0.f;
        s_p[swz(i + 1)] = 0.f;
      }
    }
    __syncthreads();

    if (do_z) {
      // sparse coefficient vector over the tile's Z characters
      for (int k = tid; k < n_zterms; k += nthr) {
        const ExpZTerm zt = s_z[k];
        float v = s_lead[zt.term];
        if ((__popcll(base & zt.zrest) & 1) ^ zt.negate) v = -v;
        if (v != 0.f) atomicAdd(&s_p[swz(zt.ztile)], v);
      }
      __syncthreads();
      for (int lvl = 0; lvl < t; lvl += 4) {
        switch (min(4, t - lvl)) {
          case 4: wht_level<4>(s_p, tile_size, lvl, tid, nthr); break;
          case 3: wht_level<3>(s_p, tile_size, lvl, tid, nthr); break;
          case 2: wht_level<2>(s_p, tile_size, lvl, tid, nthr); break;
          default: wht_level<1>(s_p, tile_size, lvl, tid, nthr); break;
        }
        __syncthreads();
      }
      // out_i += C_i * psi_i (same swizzled slot in all three arrays)
      for (uint32_t i = tid; i < tile_size; i += nthr) {
        const float c = s_p[i];
        const float2 a = s_psi[i];
        float2 o = s_out[i];
        o.x = fmaf(c, a.x, o.x);
        o.y = fmaf(c, a.y, o.y);
        s_out[i] = o;
      }
      __syncthreads();
    }

    const uint32_t ngroups = tile_size >> R;
    const uint32_t iters = (ngroups + nthr - 1) / nthr;
    for (int r = 0; r < n_rounds; ++r) {
      const RoundRec rr = s_rounds[r];
      uint32_t o[R], so[R];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        o[j] = 1u << rr.pos[j];
        so[j] = swz(o[j]);
      }
      for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t gi = it * nthr + tid;
        if (gi >= ngroups) continue;
        uint32_t b = gi;
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const uint32_t lo = o[j] - 1u;
          b = ((b & ~lo) << 1) | (b & lo);
        }
        const uint32_t sb = swz(b);
        float2 a[16], acc[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          uint32_t x = sb;
#pragma unroll
          for (int j = 0; j < R; ++j)
            if (e & (1 << j)) x ^= so[j];
          a[e] = s_psi[x];
          acc[e] = s_out[x];
        }
        const unsigned long long gbase = base | (b & lowmask) | s_hi[b >> L];
        for (int oi = rr.op_begin - first_op; oi < rr.op_end - first_op; ++oi) {
          const ExpXOp op = s_x[oi];
          float lead = s_lead[op.term];
          if (lead == 0.f) continue;      // uniform
          if (__popcll(gbase & op.zrest) & 1) lead = -lead;
          // coefficient lead * i^phase: (use_im, negate) encode the phase as
          // phase 0: (0,0)  1: (1,1)  2: (0,1)  3: (1,0); P psi uses +i^phase
          float2 c;
          if (!op.use_im) c = make_float2(op.negate ? -lead : lead, 0.f);
          else c = make_float2(0.f, op.negate ? lead : -lead);
          const float4 c4 = make_float4(c.x, c.x, -c.y, c.y);
          const uint32_t sg = op.sign16;
          switch (op.xreg) {
            case 1: xterm_accumulate<1>(acc, a, c4, sg); break;
            case 2: xterm_accumulate<2>(acc, a, c4, sg); break;
            case 3: xterm_accumulate<3>(acc, a, c4, sg); break;
            case 4: xterm_accumulate<4>(acc, a, c4, sg); break;
            case 5: xterm_accumulate<5>(acc, a, c4, sg); break;
            case 6: xterm_accumulate<6>(acc, a, c4, sg); break;
            case 7: xterm_accumulate<7>(acc, a, c4, sg); break;
            case 8: xterm_accumulate<8>(acc, a, c4, sg); break;
            case 9: xterm_accumulate<9>(acc, a, c4, sg); break;
            case 10: xterm_accumulate<10>(acc, a, c4, sg); break;
            case 11: xterm_accumulate<11>(acc, a, c4, sg); break;
            case 12: xterm_accumulate<12>(acc, a, c4, sg); break;
            case 13: xterm_accumulate<13>(acc, a, c4, sg); break;
            case 14: xterm_accumulate<14>(acc, a, c4, sg); break;
            default: xterm_accumulate<15>(acc, a, c4, sg); break;
          }
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          uint32_t x = sb;
#pragma unroll
          for (int j = 0; j < R; ++j)
            if (e & (1 << j)) x ^= so[j];
          s_out[x] = acc[e];
        }
      }
      __syncthreads();
    }

    for (uint32_t c = tid; c < tile_size / 2; c += nthr) {
      const uint32_t i = 2 * c;
      const unsigned long long g = base | (i & lowmask) | s_hi[i >> L];
      const float2 q0 = s_out[swz(i)], q1 = s_out[swz(i + 1)];
      *reinterpret_cast<float4*>(g_lam + g) = make_float4(q0.x, q0.y, q1.x, q1.y);
    }
    __syncthreads();
  }
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
                         const int32_t* __restrict__ subset,
                         double* __restrict__ per_term) {
  __shared__ double s_red[32];
  const int t = subset ? subset[blockIdx.y] : blockIdx.y;
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

// <psi_row | phi> = sum_k conj(psi_k) phi_k for every row against one phi
// (StateSpace::InnerProduct, call site math_ops/tfq_inner_product.cc:203,275)
__global__ void __launch_bounds__(kThreads)
inner_product_kernel(const float2* __restrict__ psi, size_t row_stride,
                     const float2* __restrict__ phi, unsigned long long n_amps,
                     double* __restrict__ out /* [rows][2] */) {
  __shared__ double s_red[32];
  const size_t row = blockIdx.y;
  const float2* st = psi + row * row_stride;
  double re = 0.0, im = 0.0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
       i < n_amps; i += stride) {
    const float2 a = st[i], b = phi[i];
    re += double(a.x) * double(b.x) + double(a.y) * double(b.y);
    im += double(a.x) * double(b.y) - double(a.y) * double(b.x);
  }
  const double tr = block_reduce_sum(re, s_red);
  const double ti = block_reduce_sum(im, s_red);
  if (threadIdx.x == 0) {
    atomicAdd(&out[2 * row], tr);
    atomicAdd(&out[2 * row + 1], ti);
  }
}

// lam_row (+)= c_row * u * phi with u = 1 or i: the weighted sum of paired
// states the inner-product gradient sweeps back (AccumulateFusedCircuits,
// util_qsim.h:422-442: Multiply by the float coefficient, then Add)
__global__ void __launch_bounds__(kThreads)
axpy_rows_kernel(float2* __restrict__ lam, size_t row_stride,
                 const float2* __restrict__ phi, unsigned long long n_amps,
                 const float* __restrict__ coeff, int coeff_stride, int times_i,
                 int first) {
  const size_t row = blockIdx.y;
  float2* dst = lam + row * row_stride;
  const float c = coeff[row * size_t(coeff_stride)];
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
       i < n_amps; i += stride) {
    const float2 b = phi[i];
    float2 v = make_float2(__fmul_rn(c, b.x), __fmul_rn(c, b.y));
    if (times_i) v = make_float2(-v.y, v.x);
    if (!first) {
      const float2 o = dst[i];
      v.x = __fadd_rn(o.x, v.x);
      v.y = __fadd_rn(o.y, v.y);
    }
    dst[i] = v;
  }
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
                            const int32_t* __restrict__ subset, int n_subset,
                            int accumulate,
                            const float* __restrict__ downstream, int n_ops) {
  const size_t row = blockIdx.y;
  const float2* st = psi + row * row_stride;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
       i < n_amps; i += stride) {
    float2 acc = accumulate ? lam[row * row_stride + i] : make_float2(0.f, 0.f);
    const int count = subset ? n_subset : n_terms;
    for (int tt = 0; tt < count; ++tt) {
      const DevTerm term = terms[subset ? subset[tt] : tt];
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
constexpr int kMaxDevices = 64;
static int CurrentDevice() {
  int d = 0;
  cudaGetDevice(&d);
  return d >= 0 && d < kMaxDevices ? d : 0;
}

static size_t PassSmem(int tile_bits, int mat_len, int n_ops, int n_rounds,
                       bool adj, int low_bits = kLowBits) {
  const int L = tile_bits < low_bits ? tile_bits : low_bits;
  return (size_t(adj ? 16 : 8) << tile_bits) + size_t((mat_len + 1) / 2) * 16 +
         size_t(n_ops) * sizeof(OpRec) + (size_t(8) << (tile_bits - L)) +
         size_t(n_rounds) * sizeof(RoundRec) +
         (adj ? size_t(n_ops) * 8 * (kThreads / 32) + 8 : 0) + 32;
}
size_t ForwardPassSmem(int tile_bits, int mat_len, int n_ops, int n_rounds) {
  return PassSmem(tile_bits, mat_len, n_ops, n_rounds, false);
}
size_t AdjointPassSmem(int tile_bits, int mat_len, int n_ops, int n_rounds) {
  return PassSmem(tile_bits, mat_len, n_ops, n_rounds, true);
}

static int pass_threads(int tile_bits, int reg_bits, int groups) {
  int g = (1 << (tile_bits - reg_bits)) / groups;
  if (g < 32) g = 32;
  if (g > kThreads / groups) g = kThreads / groups;
  return g;
}

static int EnvInt(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

// TFQB_DETERMINISTIC=1: every reduction that crosses CTAs with fp64 atomics
// (gradient slots, per-term partial sums) is done by ONE CTA per row instead,
// in a fixed order: results are bit-reproducible from run to run.  Meant for
// batches (rows >= a few hundred fill the GPU with one CTA each); a single
// large state becomes slow.
bool Deterministic() {
  static const bool v = EnvInt("TFQB_DETERMINISTIC", 0) != 0;
  return v;
}

template <int R, int G, bool ADJ>
static void LaunchPassT(const PassLaunch& pl, float2* psi, float2* lam,
                        size_t row_stride, int rows, double* grad_out,
                        int n_slots, int init_mode, cudaStream_t s) {
  // the attribute is per DEVICE: one process may drive several GPUs
  static bool configured[kMaxDevices] = {};  // per template instance
  const int dev = CurrentDevice();
  if (!configured[dev]) {
    cudaFuncSetAttribute(pass_kernel<R, G, ADJ>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    configured[dev] = true;
  }
  const size_t smem = PassSmem(pl.tile_bits, pl.mat_len, pl.n_ops_in_pass,
                               pl.n_rounds, ADJ, pl.low_bits);
  // tiles per CTA: up to 8 once there are enough CTAs left to fill the GPU
  unsigned n_tiles = 1u << (pl.n_alloc - pl.tile_bits);
  static const int max_seq = EnvInt("TFQB_PASS_SEQ", 8);
  unsigned gx = n_tiles;
  for (int k = 1; k < max_seq && gx > 1 && size_t(gx / 2) * rows >= 148u * 16u; k *= 2) gx /= 2;
  if (ADJ && Deterministic()) gx = 1;   // gradient slots: one CTA per row, fixed order
  const dim3 grid(gx, rows);
  const int threads = pass_threads(pl.tile_bits, R, G);
  pass_kernel<R, G, ADJ><<<grid, threads, smem, s>>>(
      psi, lam, row_stride, pl.passes, pl.rounds, pl.ops, pl.mats,
      pl.mat_row_stride, pl.pass_index, pl.first_op, pl.n_ops_in_pass, grad_out,
      n_slots, init_mode, pl.rank_base, pl.peer_tab, pl.peer_shift, pl.peer_self);
}

void LaunchForwardPass(const PassLaunch& pl, float2* psi, size_t row_stride,
                       int rows, int init_mode, cudaStream_t s) {
  static const int groups = EnvInt("TFQB_FWD_GROUPS", kFwdGroups);
  if (groups == 1)
    LaunchPassT<kRegBits, 1, false>(pl, psi, nullptr, row_stride, rows, nullptr, 0,
                                    init_mode, s);
  else
    LaunchPassT<kRegBits, 2, false>(pl, psi, nullptr, row_stride, rows, nullptr, 0,
                                    init_mode, s);
}

void LaunchAdjointPass(const PassLaunch& pl, float2* psi, float2* lam,
                       size_t row_stride, int rows, double* grad_out,
                       int n_slots, cudaStream_t s) {
  static const int groups = EnvInt("TFQB_ADJ_GROUPS", kAdjGroups);
  if (pl.reg_bits == 4)
    LaunchPassT<4, 1, true>(pl, psi, lam, row_stride, rows, grad_out, n_slots,
                            0, s);
  else if (groups == 1)
    LaunchPassT<kRegBitsAdj, 1, true>(pl, psi, lam, row_stride, rows, grad_out,
                                      n_slots, 0, s);
  else
    LaunchPassT<kRegBitsAdj, 2, true>(pl, psi, lam, row_stride, rows, grad_out,
                                      n_slots, 0, s);
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
                            const DevTerm* terms, int n_terms,
                            const int32_t* subset, int n_subset, int rows,
                            double* per_term, cudaStream_t s) {
  const int count = subset ? n_subset : n_terms;
  if (count == 0 || rows == 0) return;
  const size_t n_amps = size_t(1) << n_alloc;
  unsigned chunks = cdiv(n_amps, size_t(kThreads) * 16);
  if (chunks > 2048) chunks = 2048;
  const dim3 grid(chunks, count, rows);
  expectation_terms_kernel<<<grid, kThreads, 0, s>>>(
      psi, row_stride, n_amps, terms, n_terms, subset, per_term);
}

size_t ExpectPassSmem(int tile_bits, int low_bits, bool with_z, int n_xops,
                      int n_rounds, int n_zterms, int n_terms) {
  const int L = tile_bits < low_bits ? tile_bits : low_bits;
  return (size_t(8) << tile_bits) + (with_z ? (size_t(4) << tile_bits) : 0) +
         (size_t(8) << (tile_bits - L)) + size_t(n_xops) * sizeof(ExpXOp) +
         size_t(n_rounds) * sizeof(RoundRec) + size_t(n_zterms) * sizeof(ExpZTerm) +
         size_t(n_terms) * 8 + 64;
}

void LaunchExpectPass(const ExpectLaunch& el, const float2* psi, size_t row_stride,
                      int rows, double* per_term, cudaStream_t s) {
  if (rows == 0) return;
  cudaFuncSetAttribute(expect_pass_kernel,
                       cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  const unsigned long long n_tiles = 1ull << (el.n_alloc - el.tile_bits);
  // several tiles per CTA amortise the per-term global atomics
  unsigned ctas = unsigned(n_tiles < 64 ? n_tiles : 64);
  if (Deterministic()) ctas = 1;   // one CTA per row: no cross-CTA fp64 atomics
  int threads = 1 << (el.tile_bits > 4 ? el.tile_bits - 4 : 0);
  if (threads < 32) threads = 32;
  if (threads > kThreads) threads = kThreads;
  // per-op float partials: one slot per thread while they fit in 48 KiB,
  // else one per 4 or 32 lanes (shuffle-reduced first)
  int part_shift = 0;
  while (part_shift < 5 &&
         (size_t(el.n_xops) * size_t(threads) * 4 >> part_shift) > 48 * 1024)
    part_shift += part_shift == 0 ? 2 : 3;
  const size_t smem = ExpectPassSmem(el.tile_bits, el.low_bits, el.n_zterms > 0,
                                     el.n_xops, el.n_rounds, el.n_zterms, el.n_terms) +
                      (size_t(el.n_xops) * size_t(threads) * 4 >> part_shift);
  const dim3 grid(ctas, rows);
  expect_pass_kernel<<<grid, threads, smem, s>>>(
      psi, row_stride, el.passes, el.rounds, el.xops, el.zterms, el.n_zterms,
      el.pass_index, el.n_terms, n_tiles, el.rank_base, part_shift, per_term);
}

void LaunchInnerProduct(const float2* psi, size_t row_stride, const float2* phi,
                        int n_alloc, int rows, double* out, cudaStream_t s) {
  if (rows == 0) return;
  const size_t n_amps = size_t(1) << n_alloc;
  unsigned chunks = cdiv(n_amps, size_t(kThreads) * 8);
  if (chunks > 1024) chunks = 1024;
  const dim3 grid(chunks, rows);
  inner_product_kernel<<<grid, kThreads, 0, s>>>(psi, row_stride, phi, n_amps, out);
}

void LaunchAxpyRows(float2* lam, size_t row_stride, const float2* phi, int n_alloc,
                    const float* coeff, int coeff_stride, bool times_i, bool first,
                    int rows, cudaStream_t s) {
  if (rows == 0) return;
  const size_t n_amps = size_t(1) << n_alloc;
  unsigned chunks = cdiv(n_amps, size_t(kThreads) * 8);
  if (chunks > 1024) chunks = 1024;
  const dim3 grid(chunks, rows);
  axpy_rows_kernel<<<grid, kThreads, 0, s>>>(lam, row_stride, phi, n_amps, coeff,
                                             coeff_stride, times_i ? 1 : 0, first ? 1 : 0);
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
                               const int32_t* subset, int n_subset,
                               bool accumulate,
                               const float* downstream, int n_ops, int rows,
                               cudaStream_t s) {
  const size_t n_amps = size_t(1) << n_alloc;
  unsigned chunks = cdiv(n_amps, size_t(kThreads) * 4);
  if (chunks > 65535) chunks = 65535;
  const dim3 grid(chunks, rows);
  accumulate_operators_kernel<<<grid, kThreads, 0, s>>>(
      psi, lam, row_stride, n_amps, terms, n_terms, subset, n_subset,
      accumulate ? 1 : 0, downstream, n_ops);
}

void LaunchAccumPass(const ExpectLaunch& el, const float2* psi, float2* lam,
                     size_t row_stride, int rows, const DevTerm* terms,
                     const float* downstream, int n_ops, bool accumulate,
                     cudaStream_t s) {
  if (rows == 0) return;
  static bool configured[kMaxDevices] = {};
  const int dev = CurrentDevice();
  if (!configured[dev]) {
    cudaFuncSetAttribute(accum_pass_kernel,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    configured[dev] = true;
  }
  const size_t smem = ExpectPassSmem(el.tile_bits, el.low_bits, el.n_zterms > 0,
                                     el.n_xops, el.n_rounds, el.n_zterms, el.n_terms) +
                      (size_t(8) << el.tile_bits);
  const unsigned long long n_tiles = 1ull << (el.n_alloc - el.tile_bits);
  int threads = 1 << (el.tile_bits > 4 ? el.tile_bits - 4 : 0);
  if (threads < 32) threads = 32;
  if (threads > kThreads) threads = kThreads;
  const dim3 grid(unsigned(n_tiles < 65535 ? n_tiles : 65535), rows);
  accum_pass_kernel<<<grid, threads, smem, s>>>(
      psi, lam, row_stride, el.passes, el.rounds, el.xops, el.zterms,
      el.n_zterms, terms, el.n_terms, downstream, n_ops, el.pass_index,
      accumulate ? 1 : 0, n_tiles);
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

// ==========================================================================
// Peer-memory exchange of a state sharded over ranks.
//
// Every rank maps the shard buffers and the flag block of every other rank
// (CUDA IPC between processes, plain pointers inside one process) and the
// global<->local qubit swap is done by THIS rank's kernels loading its
// incoming chunks straight from the peers' HBM over NVLink: no send side, no
// staging, no NCCL.  Ordering is a pair of monotonically increasing epoch
// counters per rank in its flag block:
//   ready = e : every store of the segment before exchange e has completed
//   done  = e : this rank has finished reading its peers for exchange e
// written by a one-thread kernel in stream order (release.sys) and polled by
// the readers over NVLink (acquire.sys) with a bounded spin.
// ==========================================================================
__global__ void peer_signal_kernel(unsigned* flag, unsigned value) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

// thread s waits until *flags[s] >= value (wrap-safe); gives up after
// `timeout_ns` and raises *error so that the host reports it instead of hanging
__global__ void peer_wait_kernel(const unsigned* const* __restrict__ flags, int world,
                                 int self, unsigned value, unsigned long long timeout_ns,
                                 int* error) {
  const int s = threadIdx.x;
  if (s >= world || s == self) return;
  const unsigned* f = flags[s];
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (int(v - value) >= 0) break;
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) {
      atomicExch(error, 1 + s);
      break;
    }
    __nanosleep(500);
  }
}

// dst chunk s  <-  chunk `rank` of peer s's shard, for every s (chunk = the
// top g local index bits): the all-to-all of the qubit swap as seen from the
// receiving rank.  16-byte loads over NVLink, 16-byte local stores.
// The work is cut into 16 KiB pieces that rotate over the peers, starting at
// rank + 1: at any moment this GPU has loads in flight to EVERY peer and every
// peer is read by all others evenly.  (Walking the chunks in order made all
// ranks read from the same source GPU at once: 309 GB/s per GPU at 8 GPUs
// instead of 596 at 2, profiles/r02f_bench_n8.json.)
constexpr unsigned kPullPiece = 1024;     // float4 per piece: 256 threads x 4
__global__ void __launch_bounds__(256)
peer_pull_kernel(float4* __restrict__ dst, const float4* const* __restrict__ peers,
                 int world, int rank, unsigned long long chunk_vec) {
  const unsigned long long pieces_per_chunk = (chunk_vec + kPullPiece - 1) / kPullPiece;
  const unsigned long long n_pieces = pieces_per_chunk * (unsigned long long)world;
  for (unsigned long long p = blockIdx.x; p < n_pieces; p += gridDim.x) {
    const int k = int(p % (unsigned long long)world);
    const unsigned long long j0 = (p / (unsigned long long)world) * kPullPiece;
    const int s = (rank + 1 + k) % world;
    const float4* __restrict__ src = peers[s] + (unsigned long long)rank * chunk_vec;
    float4* __restrict__ out = dst + (unsigned long long)s * chunk_vec;
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned long long j = j0 + (unsigned long long)u * 256u + threadIdx.x;
      if (j < chunk_vec) v[u] = __ldcs(src + j);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned long long j = j0 + (unsigned long long)u * 256u + threadIdx.x;
      if (j < chunk_vec) out[j] = v[u];
    }
  }
}

__global__ void peer_publish_partials_kernel(const double* __restrict__ src, double* dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// out[t] = sum over ranks, in rank order, of parts[r][t]: every rank gets the
// same bits without a collective
__global__ void peer_reduce_partials_kernel(const double* const* __restrict__ parts, int world,
                                            int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = 0.0;
  for (int r = 0; r < world; ++r) acc += parts[r][i];
  out[i] = acc;
}

void LaunchPeerSignal(unsigned* flag, unsigned value, cudaStream_t s) {
  peer_signal_kernel<<<1, 1, 0, s>>>(flag, value);
}
void LaunchPeerWait(const unsigned* const* flags, int world, int self, unsigned value,
                    unsigned long long timeout_ns, int* error, cudaStream_t s) {
  peer_wait_kernel<<<1, world < 32 ? 32 : world, 0, s>>>(flags, world, self, value,
                                                          timeout_ns, error);
}
void LaunchPeerPull(float2* dst, const float2* const* peers, int world, int rank,
                    size_t chunk_amps, cudaStream_t s) {
  const unsigned long long chunk_vec = chunk_amps / 2;
  unsigned long long blocks = ((chunk_vec + kPullPiece - 1) / kPullPiece) * (unsigned long long)world;
  if (blocks > 148ull * 16) blocks = 148ull * 16;
  if (blocks == 0) blocks = 1;
  peer_pull_kernel<<<unsigned(blocks), 256, 0, s>>>(
      reinterpret_cast<float4*>(dst), reinterpret_cast<const float4* const*>(peers), world,
      rank, chunk_vec);
}
void LaunchPeerPublishPartials(const double* src, double* dst, int n, cudaStream_t s) {
  if (n <= 0) return;
  peer_publish_partials_kernel<<<(n + 127) / 128, 128, 0, s>>>(src, dst, n);
}
void LaunchPeerReducePartials(const double* const* parts, int world, int n, double* out,
                              cudaStream_t s) {
  if (n <= 0) return;
  peer_reduce_partials_kernel<<<(n + 127) / 128, 128, 0, s>>>(parts, world, n, out);
}


// ==========================================================================
// Noisy trajectory ops (next-row N2): rows are (circuit, trajectory) pairs.
// ==========================================================================
// Parameter rows [rows, cols]: columns [0, P) = the circuit's symbol values,
// [P, P + C) = one uniform per noise channel: Philox4x32-10(seed), counter
// (channel, circuit, trajectory, kNoiseStream), rounded to float32 (the same
// contract as oracle/tfq_oracle.py channel_uniforms), or the caller's.
constexpr uint32_t kNoiseStream = 0x6E6F6973u;    // "nois"
__global__ void noisy_fill_params_kernel(float* __restrict__ params, int cols, int P, int C,
                                         const float* __restrict__ symbol_values,
                                         const int32_t* __restrict__ sym_row,
                                         const int32_t* __restrict__ circuit_id,
                                         const int32_t* __restrict__ trajectory,
                                         const float* __restrict__ given_uniforms,
                                         const long long* __restrict__ given_offset,
                                         uint64_t seed, int rows) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * cols) return;
  const int row = int(idx / cols), c = int(idx % cols);
  float v = 0.f;
  if (c < P) {
    v = symbol_values[size_t(sym_row[row]) * P + c];
  } else if (c < P + C) {
    const int k = c - P;
    if (given_uniforms) {
      v = given_uniforms[given_offset[row] + k];
    } else {
      uint32_t ctr[4] = {uint32_t(k), uint32_t(circuit_id[row]), uint32_t(trajectory[row]),
                         kNoiseStream};
      uint32_t key[2] = {uint32_t(seed), uint32_t(seed >> 32)};
      for (int i = 0; i < 10; ++i) philox_round(ctr, key);
      const unsigned long long x = ((unsigned long long)ctr[0] << 32) | ctr[1];
      v = float(double(x >> 11) * (1.0 / 9007199254740992.0));
    }
  }
  params[idx] = v;
}

// u[row, k] (double) = Philox(seed; counter (k, circuit, trajectory, stream)), k < count
__global__ void noisy_fill_uniforms_kernel(double* __restrict__ u, size_t stride, int count,
                                           const int32_t* __restrict__ circuit_id,
                                           const int32_t* __restrict__ trajectory,
                                           uint32_t stream, uint64_t seed, int rows) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * count) return;
  const int row = int(idx / count), k = int(idx % count);
  uint32_t ctr[4] = {uint32_t(k), uint32_t(circuit_id[row]), uint32_t(trajectory[row]), stream};
  uint32_t key[2] = {uint32_t(seed), uint32_t(seed >> 32)};
  for (int i = 0; i < 10; ++i) philox_round(ctr, key);
  const unsigned long long x = ((unsigned long long)ctr[0] << 32) | ctr[1];
  u[size_t(row) * stride + k] = double(x >> 11) * (1.0 / 9007199254740992.0);
}

// params[row, col] = P(bit = 1) / (P(bit = 0) + P(bit = 1)) of psi_row: the
// population a non-unitary channel needs (one read of the state, fp64 sums)
__global__ void __launch_bounds__(256)
population_kernel(const float2* __restrict__ psi, size_t row_stride, int n_alloc, int bit,
                  double* __restrict__ acc /* [rows, 2] zeroed */) {
  const size_t row = blockIdx.y;
  const float2* p = psi + row * row_stride;
  const size_t N = size_t(1) << n_alloc;
  double s0 = 0.0, s1 = 0.0;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < N;
       i += size_t(gridDim.x) * blockDim.x) {
    const float2 a = p[i];
    const double w = double(a.x) * double(a.x) + double(a.y) * double(a.y);
    if ((i >> bit) & 1) s1 += w; else s0 += w;
  }
  __shared__ double red[2][8];
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s0;
    red[1][threadIdx.x >> 5] = s1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) {
      t0 += red[0][w];
      t1 += red[1][w];
    }
    atomicAdd(&acc[2 * row], t0);
    atomicAdd(&acc[2 * row + 1], t1);
  }
}
__global__ void population_store_kernel(const double* __restrict__ acc, float* __restrict__ params,
                                        int cols, int col, int rows) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const double t = acc[2 * row] + acc[2 * row + 1];
  params[size_t(row) * cols + col] = t > 0.0 ? float(acc[2 * row + 1] / t) : 0.f;
}

void LaunchNoisyFillParams(float* params, int cols, int P, int C, const float* symbol_values,
                           const int32_t* sym_row, const int32_t* circuit_id,
                           const int32_t* trajectory, const float* given_uniforms,
                           const long long* given_offset, uint64_t seed, int rows,
                           cudaStream_t s) {
  const long long total = (long long)rows * cols;
  if (total <= 0) return;
  noisy_fill_params_kernel<<<cdiv(size_t(total), 256), 256, 0, s>>>(
      params, cols, P, C, symbol_values, sym_row, circuit_id, trajectory, given_uniforms,
      given_offset, seed, rows);
}
void LaunchNoisyFillUniforms(double* u, size_t stride, int count, const int32_t* circuit_id,
                             const int32_t* trajectory, uint32_t stream, uint64_t seed, int rows,
                             cudaStream_t s) {
  const long long total = (long long)rows * count;
  if (total <= 0) return;
  noisy_fill_uniforms_kernel<<<cdiv(size_t(total), 256), 256, 0, s>>>(
      u, stride, count, circuit_id, trajectory, stream, seed, rows);
}
void LaunchPopulation(const float2* psi, size_t row_stride, int n_alloc, int bit, double* acc,
                      float* params, int cols, int col, int rows, cudaStream_t s) {
  if (rows <= 0) return;
  cudaMemsetAsync(acc, 0, size_t(rows) * 2 * sizeof(double), s);
  const size_t N = size_t(1) << n_alloc;
  unsigned bx = unsigned(std::min<size_t>((N + 255) / 256, std::max<size_t>(1, 2368 / size_t(rows))));
  if (bx == 0) bx = 1;
  population_kernel<<<dim3(bx, unsigned(rows)), 256, 0, s>>>(psi, row_stride, n_alloc, bit, acc);
  population_store_kernel<<<(rows + 127) / 128, 128, 0, s>>>(acc, params, cols, col, rows);
}

// ---- TfqCalculateUnitary (next-row N4, tfq_calculate_unitary_op.cc:47-164):
// the unitary is the circuit applied to all 2^n basis states at once (they are
// the rows of one batch of the ordinary gate passes)
__global__ void basis_states_kernel(float2* __restrict__ psi, size_t row_stride, size_t first) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= row_stride) return;
  const size_t row = blockIdx.y;
  psi[row * row_stride + i] = make_float2(i == first + row ? 1.f : 0.f, 0.f);
}
// out[j, k0 + r] = psi_r[j] for j < dim (column k0 + r of the unitary), rows = columns here
__global__ void export_unitary_kernel(const float2* __restrict__ psi, size_t row_stride,
                                      size_t dim, size_t k0, int cols,
                                      float2* __restrict__ out, size_t out_dim) {
  const size_t j = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  const size_t r = blockIdx.y;
  if (j >= dim || r >= size_t(cols)) return;
  out[j * out_dim + k0 + r] = psi[r * row_stride + j];
}
__global__ void fill_pad_kernel(float2* __restrict__ out, size_t count) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < count) out[i] = make_float2(-2.f, 0.f);
}
void LaunchBasisStates(float2* psi, size_t row_stride, size_t first, int rows, cudaStream_t s) {
  if (rows <= 0) return;
  basis_states_kernel<<<dim3(cdiv(row_stride, 256), unsigned(rows)), 256, 0, s>>>(psi, row_stride, first);
}
void LaunchExportUnitary(const float2* psi, size_t row_stride, size_t dim, size_t k0, int cols,
                         float2* out, size_t out_dim, cudaStream_t s) {
  if (cols <= 0) return;
  export_unitary_kernel<<<dim3(cdiv(dim, 256), unsigned(cols)), 256, 0, s>>>(psi, row_stride, dim, k0,
                                                                             cols, out, out_dim);
}
void LaunchFillPad(float2* out, size_t count, cudaStream_t s) {
  if (count == 0) return;
  fill_pad_kernel<<<cdiv(count, 256), 256, 0, s>>>(out, count);
}

// ---- sampling from a sharded state ---------------------------------------------
// out[0] = total of a probability tree (the norm of this rank's shard)
__global__ void tree_total_kernel(const double* __restrict__ tree, int nc, double* out) {
  out[0] = tree[tree_level_offset(nc, nc)];
}
// out[r] = parts[r][0]: every rank's published scalar, in rank order
__global__ void peer_gather_scalars_kernel(const double* const* __restrict__ parts, int world,
                                           double* __restrict__ out) {
  const int r = threadIdx.x;
  if (r < world) out[r] = parts[r][0];
}
void LaunchTreeTotal(const double* tree, int n_alloc, double* out, cudaStream_t s) {
  tree_total_kernel<<<1, 1, 0, s>>>(tree, tree_bits(n_alloc), out);
}
void LaunchPeerGatherScalars(const double* const* parts, int world, double* out, cudaStream_t s) {
  peer_gather_scalars_kernel<<<1, world < 32 ? 32 : world, 0, s>>>(parts, world, out);
}

// CUDA loads a kernel lazily at its first launch, and that load may wait for
// the device to go idle.  A sharded job launches kernels behind a spinning
// peer_wait_kernel, so everything it can launch is loaded up front (once per
// device), before the first wait is enqueued.
void PreloadShardedKernels() {
  static bool done[kMaxDevices] = {};
  const int dev = CurrentDevice();
  if (done[dev]) return;
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, pass_kernel<kRegBits, 1, false>);
  cudaFuncGetAttributes(&a, pass_kernel<kRegBits, 2, false>);
  cudaFuncGetAttributes(&a, expect_pass_kernel);
  cudaFuncGetAttributes(&a, expectation_terms_kernel);
  cudaFuncGetAttributes(&a, build_matrices_kernel);
  cudaFuncGetAttributes(&a, set_zero_state_kernel);
  cudaFuncGetAttributes(&a, peer_signal_kernel);
  cudaFuncGetAttributes(&a, peer_wait_kernel);
  cudaFuncGetAttributes(&a, peer_pull_kernel);
  cudaFuncGetAttributes(&a, peer_publish_partials_kernel);
  cudaFuncGetAttributes(&a, peer_reduce_partials_kernel);
  cudaFuncGetAttributes(&a, peer_gather_scalars_kernel);
  cudaFuncGetAttributes(&a, tree_total_kernel);
  cudaFuncGetAttributes(&a, tree_leaves_kernel);
  cudaFuncGetAttributes(&a, tree_upper_kernel);
  cudaFuncGetAttributes(&a, sample_kernel);
  cudaFuncGetAttributes(&a, fill_uniforms_kernel);
  cudaFuncGetAttributes(&a, sort_rows_kernel);
  // the attributes the launch wrappers set lazily
  cudaFuncSetAttribute(pass_kernel<kRegBits, 1, false>,
                       cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  cudaFuncSetAttribute(pass_kernel<kRegBits, 2, false>,
                       cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  cudaFuncSetAttribute(expect_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       112 * 1024);
  cudaGetLastError();
  done[dev] = true;
}

}  // namespace tfqb
