/* CPU ORACLE (test infrastructure only; never linked into the product).
 *
 * C restatement of the state-space primitives TFQ takes from qsim v0.21.0
 * (not vendored under /root/reference; WORKSPACE:74-83): gate application on
 * a complex64 state (Simulator::ApplyGate / ApplyControlledGate),
 * SetStateZero, SetAllZeros, Copy, Multiply, Add, RealInnerProduct (fp64
 * accumulation), BulkSetAmpl(exclude) -- the use counts and call sites are
 * listed in SURVEY.md 8(a) rows Q1-Q2.  It executes the step programs
 * produced by oracle/tfq_oracle.py (which restates the reference
 * orchestration sweep for sweep).
 *
 * Parallelism mirrors the reference's two modes:
 *   - one thread per circuit (ComputeSmall,
 *     tfq_simulate_expectation_op.cc:182-250): `threads` pthread workers,
 *   - all threads inside one state (ComputeLarge, :130-180; qsim's
 *     ParallelFor over the amplitude range): `inner_threads` OpenMP threads
 *     per sweep.
 * Un-controlled 1- and 2-qubit gates whose target strides are >= 4
 * amplitudes use AVX2 (8 floats per vector, like qsim's SimulatorAVX); the
 * complex product is mul, mul, addsub -- the SAME operation order as the
 * scalar code, with no FMA contraction, so scalar and vector paths give
 * bit-identical states.
 *
 * State layout is plain interleaved (re, im) float; amplitude index bit k
 * <-> qsim qubit k (pinned by util_qsim_test.cc:510-517).
 * Build: gcc -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp -shared -fPIC
 *        -pthread.
 */
#include <immintrin.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float re, im; } c32;

static inline c32 cmul(c32 a, c32 b) {
  c32 r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
  return r;
}
static inline c32 cadd(c32 a, c32 b) {
  c32 r = {a.re + b.re, a.im + b.im};
  return r;
}

/* insert a zero bit at position b of idx */
static inline uint64_t ins0(uint64_t idx, int b) {
  const uint64_t lo = idx & ((1ull << b) - 1ull);
  return ((idx >> b) << (b + 1)) | lo;
}

#ifdef __AVX2__
/* m * a for 4 packed complex numbers: (mr*ar - mi*ai, mr*ai + mi*ar), as
 * mul, mul, addsub (no FMA): bit-identical to cmul(). */
static inline __m256 vcmul(__m256 mr, __m256 mi, __m256 a) {
  const __m256 sw = _mm256_permute_ps(a, 0xB1); /* (ai, ar) */
  return _mm256_addsub_ps(_mm256_mul_ps(mr, a), _mm256_mul_ps(mi, sw));
}
#endif

/* 1-qubit gate on bit b, no controls. */
static void apply1(c32* s, int n, int b, const c32* m, int T) {
  const uint64_t N = 1ull << n, st = 1ull << b;
  const uint64_t pairs = N >> 1;
#ifdef __AVX2__
  if (st >= 4) {
    __m256 mr[4], mi[4];
    for (int k = 0; k < 4; ++k) {
      mr[k] = _mm256_set1_ps(m[k].re);
      mi[k] = _mm256_set1_ps(m[k].im);
    }
    const int64_t nv = (int64_t)(pairs >> 2);
#pragma omp parallel for schedule(static) num_threads(T) if (T > 1)
    for (int64_t v = 0; v < nv; ++v) {
      float* lo = (float*)(s + ins0((uint64_t)v << 2, b));
      float* hi = lo + 2 * st;
      const __m256 a0 = _mm256_loadu_ps(lo), a1 = _mm256_loadu_ps(hi);
      _mm256_storeu_ps(lo, _mm256_add_ps(vcmul(mr[0], mi[0], a0), vcmul(mr[1], mi[1], a1)));
      _mm256_storeu_ps(hi, _mm256_add_ps(vcmul(mr[2], mi[2], a0), vcmul(mr[3], mi[3], a1)));
    }
    return;
  }
#endif
#pragma omp parallel for schedule(static) num_threads(T) if (T > 1)
  for (int64_t p = 0; p < (int64_t)pairs; ++p) {
    c32* lo = s + ins0((uint64_t)p, b);
    c32* hi = lo + st;
    const c32 a0 = *lo, a1 = *hi;
    *lo = cadd(cmul(m[0], a0), cmul(m[1], a1));
    *hi = cadd(cmul(m[2], a0), cmul(m[3], a1));
  }
}

/* 2-qubit gate, matrix index = 2*x_{b0} + x_{b1}, no controls. */
static void apply2(c32* s, int n, int b0, int b1, const c32* m, int T) {
  const uint64_t N = 1ull << n;
  const uint64_t s0 = 1ull << b0, s1 = 1ull << b1;
  const int lo = b0 < b1 ? b0 : b1, hi = b0 < b1 ? b1 : b0;
  const uint64_t quads = N >> 2;
#ifdef __AVX2__
  if (lo >= 2) {
    __m256 mr[16], mi[16];
    for (int k = 0; k < 16; ++k) {
      mr[k] = _mm256_set1_ps(m[k].re);
      mi[k] = _mm256_set1_ps(m[k].im);
    }
    const int64_t nv = (int64_t)(quads >> 2);
#pragma omp parallel for schedule(static) num_threads(T) if (T > 1)
    for (int64_t v = 0; v < nv; ++v) {
      float* base = (float*)(s + ins0(ins0((uint64_t)v << 2, lo), hi));
      float* p[4] = {base, base + 2 * s1, base + 2 * s0, base + 2 * (s0 + s1)};
      __m256 a[4], r[4];
      for (int x = 0; x < 4; ++x) a[x] = _mm256_loadu_ps(p[x]);
      for (int x = 0; x < 4; ++x) {
        __m256 acc = vcmul(mr[4 * x], mi[4 * x], a[0]);
        acc = _mm256_add_ps(acc, vcmul(mr[4 * x + 1], mi[4 * x + 1], a[1]));
        acc = _mm256_add_ps(acc, vcmul(mr[4 * x + 2], mi[4 * x + 2], a[2]));
        acc = _mm256_add_ps(acc, vcmul(mr[4 * x + 3], mi[4 * x + 3], a[3]));
        r[x] = acc;
      }
      for (int x = 0; x < 4; ++x) _mm256_storeu_ps(p[x], r[x]);
    }
    return;
  }
#endif
#pragma omp parallel for schedule(static) num_threads(T) if (T > 1)
  for (int64_t q = 0; q < (int64_t)quads; ++q) {
    c32* base = s + ins0(ins0((uint64_t)q, lo), hi);
    c32 a[4] = {base[0], base[s1], base[s0], base[s0 + s1]};
    c32 r[4];
    for (int x = 0; x < 4; ++x) {
      c32 acc = cmul(m[4 * x], a[0]);
      acc = cadd(acc, cmul(m[4 * x + 1], a[1]));
      acc = cadd(acc, cmul(m[4 * x + 2], a[2]));
      acc = cadd(acc, cmul(m[4 * x + 3], a[3]));
      r[x] = acc;
    }
    base[0] = r[0];
    base[s1] = r[1];
    base[s0] = r[2];
    base[s0 + s1] = r[3];
  }
}

/* generic k<=4 gate with a control predicate (idx & cmask) == cbits. */
static void applyk_ctrl(c32* s, int n, int k, const int* bits, const c32* m,
                        uint64_t cmask, uint64_t cbits, int T) {
  const uint64_t N = 1ull << n;
  uint64_t tmask = 0;
  for (int j = 0; j < k; ++j) tmask |= 1ull << bits[j];
  const int d = 1 << k;
  uint64_t off[16];
  for (int x = 0; x < d; ++x) {
    uint64_t o = 0;
    for (int j = 0; j < k; ++j)
      if ((x >> (k - 1 - j)) & 1) o |= 1ull << bits[j];
    off[x] = o;
  }
#pragma omp parallel for schedule(static) num_threads(T) if (T > 1)
  for (int64_t ii = 0; ii < (int64_t)N; ++ii) {
    const uint64_t i = (uint64_t)ii;
    if (i & tmask) continue;
    if ((i & cmask) != cbits) continue;
    c32 a[16], r[16];
    for (int x = 0; x < d; ++x) a[x] = s[i | off[x]];
    for (int x = 0; x < d; ++x) {
      c32 acc = {0.f, 0.f};
      for (int y = 0; y < d; ++y) acc = cadd(acc, cmul(m[d * x + y], a[y]));
      r[x] = acc;
    }
    for (int x = 0; x < d; ++x) s[i | off[x]] = r[x];
  }
}

static double real_inner(const c32* a, const c32* b, uint64_t N, int T) {
  double acc = 0.0;
  if (T <= 1) {
    for (uint64_t i = 0; i < N; ++i)
      acc += (double)a[i].re * (double)b[i].re + (double)a[i].im * (double)b[i].im;
    return acc;
  }
#pragma omp parallel for schedule(static) num_threads(T) reduction(+ : acc)
  for (int64_t i = 0; i < (int64_t)N; ++i)
    acc += (double)a[i].re * (double)b[i].re + (double)a[i].im * (double)b[i].im;
  return acc;
}

static int run_one(int n, const int64_t* code, const float* data, double* out,
                   c32* final_state, int T) {
  const uint64_t N = 1ull << n;
  c32* buf[3];
  for (int i = 0; i < 3; ++i) {
    buf[i] = (c32*)aligned_alloc(64, N * sizeof(c32) < 64 ? 64 : N * sizeof(c32));
    if (!buf[i]) return 1;
    memset(buf[i], 0, N * sizeof(c32));
  }
  const int64_t* p = code;
  int rc = 0;
  while (*p >= 0 && rc == 0) {
    switch (*p) {
      case 0: {  /* SetStateZero */
        c32* b = buf[p[1]];
        memset(b, 0, N * sizeof(c32));
        b[0].re = 1.f;
        p += 2;
      } break;
      case 1:
        memset(buf[p[1]], 0, N * sizeof(c32));
        p += 2;
        break;
      case 2: {
        const c32* a = buf[p[1]];
        c32* b = buf[p[2]];
        if (T <= 1) {
          memcpy(b, a, N * sizeof(c32));
        } else {
#pragma omp parallel for schedule(static) num_threads(T)
          for (int64_t i = 0; i < (int64_t)N; ++i) b[i] = a[i];
        }
        p += 3;
      } break;
      case 3: {  /* Multiply */
        c32* b = buf[p[1]];
        const float c = data[p[2]];
#pragma omp parallel for schedule(static) num_threads(T) if (T > 1)
        for (int64_t i = 0; i < (int64_t)N; ++i) { b[i].re *= c; b[i].im *= c; }
        p += 3;
      } break;
      case 4: {  /* Add */
        const c32* a = buf[p[1]];
        c32* b = buf[p[2]];
#pragma omp parallel for schedule(static) num_threads(T) if (T > 1)
        for (int64_t i = 0; i < (int64_t)N; ++i) { b[i].re += a[i].re; b[i].im += a[i].im; }
        p += 3;
      } break;
      case 5: {  /* ApplyGate */
        c32* b = buf[p[1]];
        const int k = (int)p[2];
        int bits[4];
        if (k > 4) { rc = 2; break; }
        for (int j = 0; j < k; ++j) bits[j] = (int)p[3 + j];
        const uint64_t cmask = (uint64_t)p[3 + k], cbits = (uint64_t)p[4 + k];
        const c32* m = (const c32*)(data + p[5 + k]);
        if (cmask == 0 && k == 1) apply1(b, n, bits[0], m, T);
        else if (cmask == 0 && k == 2) apply2(b, n, bits[0], bits[1], m, T);
        else applyk_ctrl(b, n, k, bits, m, cmask, cbits, T);
        p += 6 + k;
      } break;
      case 6: {  /* BulkSetAmpl(mask, bits, 0, 0, exclude=true) */
        c32* b = buf[p[1]];
        const uint64_t cmask = (uint64_t)p[2], cbits = (uint64_t)p[3];
#pragma omp parallel for schedule(static) num_threads(T) if (T > 1)
        for (int64_t i = 0; i < (int64_t)N; ++i)
          if (((uint64_t)i & cmask) != cbits) { b[i].re = 0.f; b[i].im = 0.f; }
        p += 4;
      } break;
      case 7: {  /* out[slot] += coeff * RealInnerProduct(a, b) */
        const double v = real_inner(buf[p[1]], buf[p[2]], N, T);
        out[p[3]] += p[5] ? (double)data[p[4]] * v : v;
        p += 6;
      } break;
      default:
        rc = 3;
    }
  }
  if (final_state) memcpy(final_state, buf[0], N * sizeof(c32));
  for (int i = 0; i < 3; ++i) free(buf[i]);
  return rc;
}

typedef struct {
  int nb;
  const int32_t* ns;
  void* const* codes;
  void* const* datas;
  void* const* outs;
  void* const* states;
  int next;
  int rc;
  int inner;
  pthread_mutex_t mu;
} batch_t;

static void* worker(void* arg) {
  batch_t* b = (batch_t*)arg;
  for (;;) {
    pthread_mutex_lock(&b->mu);
    const int i = b->next++;
    pthread_mutex_unlock(&b->mu);
    if (i >= b->nb) break;
    const int rc = run_one(b->ns[i], (const int64_t*)b->codes[i],
                           (const float*)b->datas[i], (double*)b->outs[i],
                           b->states ? (c32*)b->states[i] : NULL, b->inner);
    if (rc) {
      pthread_mutex_lock(&b->mu);
      b->rc = rc;
      pthread_mutex_unlock(&b->mu);
    }
  }
  return NULL;
}

/* `threads` circuits at a time, `inner_threads` OpenMP threads inside each
 * sweep of a circuit. */
int qvm_run_batch2(int nb, const int32_t* ns, void* const* codes,
                   void* const* datas, void* const* outs, void* const* states,
                   int threads, int inner_threads) {
  batch_t b = {nb, ns, codes, datas, outs, states, 0, 0,
               inner_threads < 1 ? 1 : inner_threads, PTHREAD_MUTEX_INITIALIZER};
  if (threads < 1) threads = 1;
  if (threads > nb) threads = nb;
  if (threads <= 1) {
    worker(&b);
    return b.rc;
  }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, worker, &b);
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  free(th);
  return b.rc;
}

int qvm_run_batch(int nb, const int32_t* ns, void* const* codes,
                  void* const* datas, void* const* outs, void* const* states,
                  int threads) {
  return qvm_run_batch2(nb, ns, codes, datas, outs, states, threads, 1);
}

/* 1: the AVX2 gate loops are compiled in */
int qvm_has_avx2(void) {
#ifdef __AVX2__
  return 1;
#else
  return 0;
#endif
}
