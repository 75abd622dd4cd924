/* CPU restatement of the Hamming kNN hot path in plain C.  TEST / BASELINE INFRASTRUCTURE ONLY:
 * nothing under prograph_b200/ links or loads this file; tests/ pins it against the numpy oracle
 * and bench.py times it as the "best effort" CPU baseline next to the reference's own algorithm.
 *
 * What it restates (acmater/prograph, read-only reference):
 *   hamming.py:34         d[m,n] = sum_l [X[n,l] != Y[m,l]]
 *   prograph.py:757-762   sort each row of d, drop sorted position 0, keep the next k
 *                         (ties by ascending index: the stable order)
 * How: residues (tokens 0..31) are spread over 5 bit planes of 64-bit words; a mismatch mask is the
 * OR over the planes of the XORs, the distance its popcount; every query row keeps the k+1 smallest
 * (distance, index) keys seen so far in a small sorted array.  Query rows are dealt to POSIX threads.
 *
 * Build: make -C oracle   (gcc -O3 -mpopcnt -pthread -shared -fPIC -> oracle/_build/libpg_oracle_c.so)
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PLANES 5

/* tokens [n][L] uint8 -> planes [n][PLANES][W] uint64, W = ceil(L / 64); returns 0, or -1 if a
 * token does not fit in 5 bits */
int pgo_pack(const uint8_t* tokens, int64_t n, int L, uint64_t* planes) {
  const int W = (L + 63) / 64;
  int bad = 0;
  for (int64_t r = 0; r < n; ++r) {
    uint64_t* out = planes + (size_t)r * PLANES * W;
    memset(out, 0, sizeof(uint64_t) * PLANES * W);
    for (int l = 0; l < L; ++l) {
      const unsigned t = tokens[(size_t)r * L + l];
      if (t > 31) bad = 1;
      for (int p = 0; p < PLANES; ++p)
        if ((t >> p) & 1u) out[p * W + (l >> 6)] |= 1ull << (l & 63);
    }
  }
  return bad ? -1 : 0;
}

static inline int distance(const uint64_t* a, const uint64_t* b, int W) {
  int d = 0;
  for (int w = 0; w < W; ++w) {
    uint64_t m = a[w] ^ b[w];
    for (int p = 1; p < PLANES; ++p) m |= a[p * W + w] ^ b[p * W + w];
    d += __builtin_popcountll(m);
  }
  return d;
}

struct knn_job {
  const uint64_t* planes;
  int64_t n, q0, nq;
  int W, k, tid, threads;
  int64_t *out_idx, *out_d;
  const int64_t* rows;     /* optional: the query rows (nq indices); NULL = rows q0 .. q0+nq-1 */
  int32_t* out_all;        /* pgo_hamming_rows: [nq][n] distances */
};

static inline int64_t query_row(const struct knn_job* jb, int64_t i) { return jb->rows ? jb->rows[i] : jb->q0 + i; }

#define QB 8   /* query rows a thread sweeps together: the table is streamed once per block, not per row */

static void* knn_worker(void* arg) {
  const struct knn_job* jb = (const struct knn_job*)arg;
  const int W = jb->W, k = jb->k, k1 = jb->k + 1;
  const int64_t n_blocks = (jb->nq + QB - 1) / QB;
  uint64_t (*keys)[256] = malloc(sizeof(uint64_t[QB][256]));
  if (!keys) return NULL;
  for (int64_t b = jb->tid; b < n_blocks; b += jb->threads) {       /* blocks dealt round robin */
    const int64_t qb0 = b * QB;
    const int nb = (int)(jb->nq - qb0 < QB ? jb->nq - qb0 : QB);
    int have[QB];
    uint64_t last[QB];                          /* key of the k1-th entry once a list is full */
    for (int i = 0; i < nb; ++i) { have[i] = 0; last[i] = ~0ull; }
    for (int64_t j = 0; j < jb->n; ++j) {
      const uint64_t* x = jb->planes + (size_t)j * PLANES * W;
      for (int qi = 0; qi < nb; ++qi) {
        const uint64_t* q = jb->planes + (size_t)query_row(jb, qb0 + qi) * PLANES * W;
        const uint64_t key = ((uint64_t)distance(q, x, W) << 32) | (uint64_t)j;
        if (have[qi] == k1 && key >= last[qi]) continue;
        int i = have[qi] < k1 ? have[qi]++ : k1 - 1;      /* insertion into the ascending array */
        while (i > 0 && keys[qi][i - 1] > key) { keys[qi][i] = keys[qi][i - 1]; --i; }
        keys[qi][i] = key;
        if (have[qi] == k1) last[qi] = keys[qi][k1 - 1];
      }
    }
    for (int qi = 0; qi < nb; ++qi) {
      int64_t* oi = jb->out_idx + (qb0 + qi) * k;
      int64_t* od = jb->out_d + (qb0 + qi) * k;
      for (int j = 0; j < k; ++j) {
        if (j + 1 < have[qi]) {
          oi[j] = (int64_t)(keys[qi][j + 1] & 0xffffffffull);
          od[j] = (int64_t)(keys[qi][j + 1] >> 32);
        } else {
          oi[j] = -1;
          od[j] = 0;
        }
      }
    }
  }
  free(keys);
  return NULL;
}

static int run_jobs(void* (*worker)(void*), struct knn_job proto, int threads) {
  pthread_t tids[1024];
  struct knn_job jobs[1024];
  for (int t = 0; t < threads; ++t) {
    jobs[t] = proto;
    jobs[t].tid = t;
    jobs[t].threads = threads;
    if (t > 0 && pthread_create(&tids[t], NULL, worker, &jobs[t]) != 0) {
      for (int u = 1; u < t; ++u) pthread_join(tids[u], NULL);
      return -2;
    }
  }
  worker(&jobs[0]);
  for (int t = 1; t < threads; ++t) pthread_join(tids[t], NULL);
  return 0;
}

/* kNN of query rows [q0, q0+nq) of `planes` against all n rows on `threads` threads: out_idx /
 * out_d [nq][k] hold sorted positions 1..k of every row in (distance, index) order; missing
 * entries (n < k+1) get -1 / 0 */
int pgo_hamming_knn(const uint64_t* planes, int64_t n, int L, int64_t q0, int64_t nq, int k, int threads,
                    int64_t* out_idx, int64_t* out_d) {
  if (k < 1 || k + 1 > 256 || q0 < 0 || nq < 0 || q0 + nq > n || threads < 1 || threads > 1024) return -1;
  struct knn_job proto = {planes, n, q0, nq, (L + 63) / 64, k, 0, threads, out_idx, out_d, NULL, NULL};
  return run_jobs(knn_worker, proto, threads);
}

/* the same for an arbitrary list of query rows (bench.py samples band edges and random rows) */
int pgo_hamming_knn_rows(const uint64_t* planes, int64_t n, int L, const int64_t* rows, int64_t nq, int k, int threads,
                         int64_t* out_idx, int64_t* out_d) {
  if (k < 1 || k + 1 > 256 || nq < 0 || !rows || threads < 1 || threads > 1024) return -1;
  for (int64_t i = 0; i < nq; ++i)
    if (rows[i] < 0 || rows[i] >= n) return -1;
  struct knn_job proto = {planes, n, 0, nq, (L + 63) / 64, k, 0, threads, out_idx, out_d, rows, NULL};
  return run_jobs(knn_worker, proto, threads);
}

static void* rows_worker(void* arg) {
  const struct knn_job* jb = (const struct knn_job*)arg;
  const int W = jb->W;
  for (int64_t i = jb->tid; i < jb->nq; i += jb->threads) {
    const uint64_t* q = jb->planes + (size_t)query_row(jb, i) * PLANES * W;
    int32_t* out = jb->out_all + (size_t)i * jb->n;
    for (int64_t j = 0; j < jb->n; ++j) out[j] = distance(q, jb->planes + (size_t)j * PLANES * W, W);
  }
  return NULL;
}

/* hamming.py:34 for a list of query rows against all n rows: out [nq][n] int32 */
int pgo_hamming_rows(const uint64_t* planes, int64_t n, int L, const int64_t* rows, int64_t nq, int threads,
                     int32_t* out) {
  if (nq < 0 || !rows || !out || threads < 1 || threads > 1024) return -1;
  for (int64_t i = 0; i < nq; ++i)
    if (rows[i] < 0 || rows[i] >= n) return -1;
  struct knn_job proto = {planes, n, 0, nq, (L + 63) / 64, 0, 0, threads, NULL, NULL, rows, out};
  return run_jobs(rows_worker, proto, threads);
}
