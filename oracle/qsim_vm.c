/* CPU ORACLE (test infrastructure only; never linked into the product).
 *
 * Plain-C restatement of the state-space primitives TFQ takes from qsim
 * v0.21.0 (not vendored under /root/reference; WORKSPACE:74-83): gate
 * application on a complex64 state (Simulator::ApplyGate /
 * ApplyControlledGate), SetStateZero, SetAllZeros, Copy, Multiply, Add,
 * RealInnerProduct (fp64 accumulation), BulkSetAmpl(exclude) -- the use
 * counts and call sites are listed in SURVEY.md 8(a) rows Q1-Q2.  It executes
 * the step programs produced by oracle/tfq_oracle.py (which restates the
 * reference orchestration sweep for sweep), one thread per circuit like
 * ComputeSmall (tfq_simulate_expectation_op.cc:182-250).
 *
 * State layout here is plain interleaved (re, im) float; amplitude index
 * bit k <-> qsim qubit k (pinned by util_qsim_test.cc:510-517).
 * Build: gcc -O3 -march=native -ffp-contract=off -shared -fPIC -pthread.
 */
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

/* 1-qubit gate on bit b, no controls: contiguous inner loop (vectorises). */
static void apply1(c32* s, int n, int b, const c32* m) {
  const uint64_t N = 1ull << n, st = 1ull << b;
  for (uint64_t i = 0; i < N; i += 2 * st) {
    c32* lo = s + i;
    c32* hi = s + i + st;
    for (uint64_t j = 0; j < st; ++j) {
      c32 a0 = lo[j], a1 = hi[j];
      lo[j] = cadd(cmul(m[0], a0), cmul(m[1], a1));
      hi[j] = cadd(cmul(m[2], a0), cmul(m[3], a1));
    }
  }
}

/* 2-qubit gate, matrix index = 2*x_{b0} + x_{b1}, no controls. */
static void apply2(c32* s, int n, int b0, int b1, const c32* m) {
  const uint64_t N = 1ull << n;
  const uint64_t s0 = 1ull << b0, s1 = 1ull << b1;
  const int lo = b0 < b1 ? b0 : b1, hi = b0 < b1 ? b1 : b0;
  const uint64_t slo = 1ull << lo, shi = 1ull << hi;
  for (uint64_t i = 0; i < N; i += 2 * shi) {
    for (uint64_t j = 0; j < shi; j += 2 * slo) {
      c32* base = s + i + j;
      for (uint64_t k = 0; k < slo; ++k) {
        c32 a[4] = {base[k], base[k + s1], base[k + s0], base[k + s0 + s1]};
        c32 r[4];
        for (int x = 0; x < 4; ++x) {
          c32 acc = cmul(m[4 * x], a[0]);
          acc = cadd(acc, cmul(m[4 * x + 1], a[1]));
          acc = cadd(acc, cmul(m[4 * x + 2], a[2]));
          acc = cadd(acc, cmul(m[4 * x + 3], a[3]));
          r[x] = acc;
        }
        base[k] = r[0];
        base[k + s1] = r[1];
        base[k + s0] = r[2];
        base[k + s0 + s1] = r[3];
      }
    }
  }
}

/* generic k<=4 gate with a control predicate (idx & cmask) == cbits. */
static void applyk_ctrl(c32* s, int n, int k, const int* bits, const c32* m,
                        uint64_t cmask, uint64_t cbits) {
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
  for (uint64_t i = 0; i < N; ++i) {
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

static double real_inner(const c32* a, const c32* b, uint64_t N) {
  double acc = 0.0;
  for (uint64_t i = 0; i < N; ++i)
    acc += (double)a[i].re * (double)b[i].re + (double)a[i].im * (double)b[i].im;
  return acc;
}

static int run_one(int n, const int64_t* code, const float* data, double* out,
                   c32* final_state) {
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
      case 2:
        memcpy(buf[p[2]], buf[p[1]], N * sizeof(c32));
        p += 3;
        break;
      case 3: {  /* Multiply */
        c32* b = buf[p[1]];
        const float c = data[p[2]];
        for (uint64_t i = 0; i < N; ++i) { b[i].re *= c; b[i].im *= c; }
        p += 3;
      } break;
      case 4: {  /* Add */
        const c32* a = buf[p[1]];
        c32* b = buf[p[2]];
        for (uint64_t i = 0; i < N; ++i) { b[i].re += a[i].re; b[i].im += a[i].im; }
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
        if (cmask == 0 && k == 1) apply1(b, n, bits[0], m);
        else if (cmask == 0 && k == 2) apply2(b, n, bits[0], bits[1], m);
        else applyk_ctrl(b, n, k, bits, m, cmask, cbits);
        p += 6 + k;
      } break;
      case 6: {  /* BulkSetAmpl(mask, bits, 0, 0, exclude=true) */
        c32* b = buf[p[1]];
        const uint64_t cmask = (uint64_t)p[2], cbits = (uint64_t)p[3];
        for (uint64_t i = 0; i < N; ++i)
          if ((i & cmask) != cbits) { b[i].re = 0.f; b[i].im = 0.f; }
        p += 4;
      } break;
      case 7: {  /* out[slot] += coeff * RealInnerProduct(a, b) */
        const double v = real_inner(buf[p[1]], buf[p[2]], N);
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
                           b->states ? (c32*)b->states[i] : NULL);
    if (rc) {
      pthread_mutex_lock(&b->mu);
      b->rc = rc;
      pthread_mutex_unlock(&b->mu);
    }
  }
  return NULL;
}

int qvm_run_batch(int nb, const int32_t* ns, void* const* codes,
                  void* const* datas, void* const* outs, void* const* states,
                  int threads) {
  batch_t b = {nb, ns, codes, datas, outs, states, 0, 0,
               PTHREAD_MUTEX_INITIALIZER};
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
