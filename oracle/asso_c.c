/*
 * TEST INFRASTRUCTURE ONLY -- plain-C, bit-packed CPU restatement of the arithmetic of PyBMF's Asso hot path.
 *
 * This file is the *checker* (and the CPU arm of bench.py), never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it (through
 * oracle/asso_oracle_c.py).  Nothing under pybmf_b200/ links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks the fit built on these functions against
 * the numpy restatement (oracle/asso_oracle.py), the known-answer table of examples/ex01_6_logs.ipynb and the
 * golden outputs of the genuine reference (the npz files under tests/golden).
 *
 * It exists because the numpy restatement needs dense m x n float32 operands (34 GB at the Netflix-shaped config)
 * and the genuine reference needs ~43 s per candidate there: this version keeps X, the covered mask and the candidate
 * basis as bit rows and evaluates every (data row, candidate) pair with AND + POPCNT, OpenMP over row blocks, so a
 * whole greedy step of BASELINE config c4 (8.5e9 pairs) finishes in about a minute on a few host cores.
 *
 * Layout: bit matrices are row-major uint64 words, bit c of a row in word c>>6 at position c&63, pad bits 0
 * (the layout of include/pybmf_b200.h, so tests can hand the same arrays to both sides).
 * Each function cites the reference lines (relative to /root/reference) whose arithmetic it restates.
 * Floating point: compiled with -ffp-contract=off; every expression is written in the reference's order.
 *
 *   gcc -O3 -fopenmp -fPIC -shared -ffp-contract=off -mpopcnt oracle/asso_c.c -o oracle/libasso_oracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;
typedef uint64_t u64;

/* ---- |a & b| over w words ----------------------------------------------------------------- */
static inline i64 and_popc_scalar(const u64* a, const u64* b, i64 w) {
  i64 s = 0;
  for (i64 k = 0; k < w; ++k) s += __builtin_popcountll(a[k] & b[k]);
  return s;
}
#if defined(__x86_64__)
__attribute__((target("avx512f,avx512vpopcntdq"))) static inline i64 and_popc_avx512(const u64* a, const u64* b, i64 w) {
  __m512i acc = _mm512_setzero_si512();
  i64 k = 0;
  for (; k + 8 <= w; k += 8)
    acc = _mm512_add_epi64(acc, _mm512_popcnt_epi64(_mm512_and_si512(_mm512_loadu_si512(a + k), _mm512_loadu_si512(b + k))));
  i64 s = _mm512_reduce_add_epi64(acc);
  for (; k < w; ++k) s += __builtin_popcountll(a[k] & b[k]);
  return s;
}
static int have_avx512(void) {
  static int cached = -1;
  if (cached < 0) cached = (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vpopcntdq")) ? 1 : 0;
  return cached;
}
#else
static int have_avx512(void) { return 0; }
#endif
typedef i64 (*and_popc_fn)(const u64*, const u64*, i64);
static and_popc_fn pick_and_popc(void) {
#if defined(__x86_64__)
  if (have_avx512()) return and_popc_avx512;
#endif
  return and_popc_scalar;
}

int bmfo_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void bmfo_set_threads(int t) {
#ifdef _OPENMP
  if (t > 0) omp_set_num_threads(t);
#else
  (void)t;
#endif
}
int bmfo_has_avx512(void) { return have_avx512(); }

/* ---- build_assoc, PyBMF/models/Asso.py:207: cnt = X^T X (co-occurrence counts) --------------
 * xt = bit rows of X^T: [n][wm], wm = words covering the m data rows.  cnt is n x n int32 (symmetric). */
void bmfo_assoc_counts(const u64* xt, i64 n, i64 wm, int32_t* cnt) {
  and_popc_fn f = pick_and_popc();
#pragma omp parallel for schedule(dynamic, 8)
  for (i64 i = 0; i < n; ++i)
    for (i64 j = i; j < n; ++j) {
      const int32_t v = (int32_t)f(xt + i * wm, xt + j * wm, wm);
      cnt[i * n + j] = v;
      cnt[j * n + i] = v;
    }
}

/* ---- build_assoc row normalisation Asso.py:208-212 + build_basis Asso.py:231-234 (binarize, common.py:75) ----
 * bit (i, j) = (double)cnt[i][j] / (double)cnt[i][i] > tau (IEEE division, strict >); the association row of an empty
 * column is 0 (`... if s[i] > 0 else 0`, Asso.py:211), so its bits are 0 > tau (set only for a negative tau).
 * alive[i] = row has any bit (all-zero rows are dropped by the reference, order kept); pop[i] = |b_i|. */
void bmfo_basis(const int32_t* cnt, i64 n, double tau, u64* basis, i64 words, uint8_t* alive, int32_t* pop) {
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; ++i) {
    u64* row = basis + i * words;
    memset(row, 0, sizeof(u64) * (size_t)words);
    const int32_t si = cnt[i * n + i];
    int p = 0;
    const double s = (double)si;
    for (i64 j = 0; j < n; ++j) {
      const double a = si > 0 ? (double)cnt[i * n + j] / s : 0.0;
      if (a > tau) { row[j >> 6] |= 1ull << (j & 63); ++p; }
    }
    alive[i] = p > 0;
    pop[i] = p;
  }
}

/* The per-row decision of get_vector, Asso.py:179-181 with coverage_score metrics.py:201:
 *   s = (-w_fp) * FP + w_fn * TP  (two products, one sum, no FMA);  use = s_new > s_old  (strict). */
static inline int row_uses(double neg_w_fp, double w_fn, i64 tpo, i64 fpo, i64 P, i64 N) {
  const double s_old = neg_w_fp * (double)fpo + w_fn * (double)tpo;
  const double s_new = neg_w_fp * (double)(fpo + N) + w_fn * (double)(tpo + P);
  return s_new > s_old;
}

/* ---- hot loop Asso.py:83-95 -> get_vector Asso.py:144-188, all candidates of one greedy step -----------------
 * x, c: [m][words] data rows and covered mask; basis [n][words] with alive[] flags and pop[j] = |b_j|;
 * tpo/fpo [m]: TP / FP of the current cover per data row.
 * For candidate j and data row i:  P = |x_i & ~c_i & b_j|,  N = |~x_i & ~c_i & b_j| = |b_j| - |b_j & c_i| - P.
 * Outputs per candidate (overwritten; dead candidates get 0):
 *   gain_p[j] = sum_{use} P, gain_n[j] = sum_{use} N, and when wa|wb (weights wa/2^s, wb/2^s, all arithmetic exact)
 *   gain_d[j] = sum_i relu(wb*P - wa*N)  (the quantity the tensor-core path accumulates). */
void bmfo_score_all(const u64* x, const u64* c, i64 m, i64 words, const u64* basis, const uint8_t* alive,
                    const int32_t* pop, i64 n, const int32_t* tpo, const int32_t* fpo, double w_fp, double w_fn,
                    int wa, int wb, i64* gain_p, i64* gain_n, i64* gain_d) {
  and_popc_fn f = pick_and_popc();
  const double neg_w_fp = -w_fp;
  memset(gain_p, 0, sizeof(i64) * (size_t)n);
  memset(gain_n, 0, sizeof(i64) * (size_t)n);
  if (gain_d) memset(gain_d, 0, sizeof(i64) * (size_t)n);
  const i64 RB = 64;   /* data rows per block: their words stay in L2 while all candidates stream past */
#pragma omp parallel
  {
    i64* lp = (i64*)calloc((size_t)n * 3, sizeof(i64));
    i64 *ln = lp + n, *ld = lp + 2 * n;
    u64* xu = (u64*)malloc(sizeof(u64) * (size_t)(RB * words));
    uint8_t* covered = (uint8_t*)malloc((size_t)RB);
#pragma omp for schedule(dynamic, 1)
    for (i64 i0 = 0; i0 < m; i0 += RB) {
      const i64 rows = (m - i0) < RB ? (m - i0) : RB;
      for (i64 r = 0; r < rows; ++r) {
        const u64 *xr = x + (i0 + r) * words, *cr = c + (i0 + r) * words;
        u64 any = 0;
        for (i64 k = 0; k < words; ++k) { xu[r * words + k] = xr[k] & ~cr[k]; any |= cr[k]; }
        covered[r] = any != 0;
      }
      for (i64 j = 0; j < n; ++j) {
        if (!alive[j]) continue;
        const u64* b = basis + j * words;
        i64 sp = 0, sn = 0, sd = 0;
        for (i64 r = 0; r < rows; ++r) {
          const i64 P = f(xu + r * words, b, words);
          const i64 Q = covered[r] ? f(c + (i0 + r) * words, b, words) : 0;
          const i64 N = (i64)pop[j] - Q - P;
          if (wa | wb) {
            const i64 d = (i64)wb * P - (i64)wa * N;
            if (d > 0) { sd += d; sp += P; sn += N; }
          } else if (row_uses(neg_w_fp, w_fn, tpo[i0 + r], fpo[i0 + r], P, N)) {
            sp += P; sn += N;
          }
        }
        lp[j] += sp; ln[j] += sn; ld[j] += sd;
      }
    }
#pragma omp critical
    {
      for (i64 j = 0; j < n; ++j) { gain_p[j] += lp[j]; gain_n[j] += ln[j]; if (gain_d) gain_d[j] += ld[j]; }
    }
    free(lp); free(xu); free(covered);
  }
}

/* ---- set_factors + cover update for the chosen row b (Asso.py:103-110): u_i = use(i), c_i |= b where used,
 * tpo/fpo updated; used[i] (bytes) written; totals = (#used, sum P, sum N). */
void bmfo_apply(const u64* x, u64* c, i64 m, i64 words, const u64* b, i64 b_pop, int32_t* tpo, int32_t* fpo,
                double w_fp, double w_fn, int wa, int wb, uint8_t* used, i64* totals) {
  and_popc_fn f = pick_and_popc();
  const double neg_w_fp = -w_fp;
  i64 t_used = 0, t_p = 0, t_n = 0;
#pragma omp parallel for schedule(static) reduction(+ : t_used, t_p, t_n)
  for (i64 i = 0; i < m; ++i) {
    const u64* xr = x + i * words;
    u64* cr = c + i * words;
    i64 P = 0, Q = 0;
    for (i64 k = 0; k < words; ++k) {
      P += __builtin_popcountll(xr[k] & ~cr[k] & b[k]);
      Q += __builtin_popcountll(cr[k] & b[k]);
    }
    (void)f;
    const i64 N = b_pop - Q - P;
    int use;
    if (wa | wb) use = ((i64)wb * P - (i64)wa * N) > 0;
    else use = row_uses(neg_w_fp, w_fn, tpo[i], fpo[i], P, N);
    used[i] = (uint8_t)use;
    if (use) {
      for (i64 k = 0; k < words; ++k) cr[k] |= b[k];
      tpo[i] += (int32_t)P;
      fpo[i] += (int32_t)N;
      t_used += 1; t_p += P; t_n += N;
    }
  }
  totals[0] = t_used; totals[1] = t_p; totals[2] = t_n;
}

/* ---- TP / FP / FN of PyBMF/utils/metrics.py:56-76 on bit rows (totals and, optionally, per row) ---- */
void bmfo_confusion(const u64* gt, const u64* pd, i64 m, i64 words, i64* counts, int32_t* row_tp, int32_t* row_fp) {
  i64 tp = 0, fp = 0, fn = 0;
#pragma omp parallel for schedule(static) reduction(+ : tp, fp, fn)
  for (i64 i = 0; i < m; ++i) {
    i64 a = 0, b = 0, g = 0;
    for (i64 k = 0; k < words; ++k) {
      const u64 x = gt[i * words + k], p = pd[i * words + k];
      a += __builtin_popcountll(x & p);
      b += __builtin_popcountll(p & ~x);
      g += __builtin_popcountll(x & ~p);
    }
    if (row_tp) row_tp[i] = (int32_t)a;
    if (row_fp) row_fp[i] = (int32_t)b;
    tp += a; fp += b; fn += g;
  }
  counts[0] = tp; counts[1] = fp; counts[2] = fn;
}

/* ---- get_prediction / matmul(boolean=True), PyBMF/utils/boolean_utils.py:71-78: pd_i = OR_{l in U_i} vt_l.
 * u_words [m][kw] (bit l of row i = U[i][l]); vt [k][words]; skip >= 0 leaves factor `skip` out (AssoIter.py:85-86). */
void bmfo_bool_product(const u64* u_words, i64 m, i64 kw, const u64* vt, i64 k, i64 words, i64 skip, u64* pd) {
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < m; ++i) {
    u64* out = pd + i * words;
    memset(out, 0, sizeof(u64) * (size_t)words);
    for (i64 l = 0; l < k; ++l) {
      if (l == skip) continue;
      if ((u_words[i * kw + (l >> 6)] >> (l & 63)) & 1ull) {
        const u64* v = vt + l * words;
        for (i64 q = 0; q < words; ++q) out[q] |= v[q];
      }
    }
  }
}

/* ---- csr pattern -> bit rows (the containers of BaseModel.load_dataset, PyBMF/models/BaseModel.py:146).
 * transposed = 1 writes X^T ([n][words] with words covering m).  bits must be zero-filled. */
void bmfo_pack_csr(const i64* indptr, const int32_t* indices, i64 m, int transposed, u64* bits, i64 words) {
  if (!transposed) {
#pragma omp parallel for schedule(static)
    for (i64 r = 0; r < m; ++r)
      for (i64 e = indptr[r]; e < indptr[r + 1]; ++e) bits[r * words + (indices[e] >> 6)] |= 1ull << (indices[e] & 63);
  } else {
    for (i64 r = 0; r < m; ++r)
      for (i64 e = indptr[r]; e < indptr[r + 1]; ++e) bits[(i64)indices[e] * words + (r >> 6)] |= 1ull << (r & 63);
  }
}
