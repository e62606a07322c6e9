/*
 * dbl_refsweep.c -- reference-STYLE CPU link update, for timing only (bench.py cpu_baseline / --impl reference).
 * TEST/BENCH INFRASTRUCTURE ONLY; never linked into the product.
 *
 * It keeps the reference's algorithmic choices for the dominant phase of a sweep:
 *   PCG-II  : updateEntityIdCollapsed (GU:363-395): weights over ALL entities of the block, literal
 *             per-attribute formula, then DiscreteDist(weights).sample() = alias-table build + one draw
 *             (random/AliasSampler.scala:49-118).
 *   PCG-I   : updateEntityId (GU:399-430) with the inverted index (GU:41-76) rebuilt per block per sweep
 *             (GU:178-184) and getPossibleEntities' smallest-set-first intersection (GU:473-530).
 * and parallelises over records with all host threads (the reference runs one Spark task per block,
 * GU:137; a record-level split balances better, i.e. this baseline is at least as fast as that schedule).
 * Look-ups that are hash maps in the JVM (distProbs(attrId, fileId), expSimOf) are plain array accesses
 * here, again in the baseline's favour.  The RNG is a 64-bit LCG-free xorshift (speed only).
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "dbl_oracle.h"

#include "dbl_oracle_priv.h"
#define orc_index_pub orc_index
#define orc_model_pub orc_model
#define orc_state_pub orc_state

typedef struct {
  const struct orc_state_pub *s;
  int sampler, P;
  const int64_t *bptr;   /* block -> entity range in bent */
  const int32_t *bent;
  /* inverted index per block: for (block, attr): values sorted + postings */
  const int64_t *inv_ptr;   /* (P*A + 1) offsets into inv_val/inv_off */
  const int32_t *inv_val;   /* distinct values, ascending, per (block, attr) */
  const int64_t *inv_off;   /* postings offsets per distinct value (+1 sentinel per (block, attr)) */
  const int32_t *inv_post;  /* entity ids */
  const int64_t *sample;    /* record ids */
  int64_t n_sample;
  int tid, nthreads;
  int64_t pairs;
  int32_t *out;
} work_t;

static inline uint64_t xs64(uint64_t *st) {
  uint64_t x = *st;
  x ^= x << 13; x ^= x >> 7; x ^= x << 17;
  *st = x;
  return x;
}
static inline double xs_unit(uint64_t *st) { return (double)(xs64(st) >> 11) * (1.0 / 9007199254740992.0); }

static double exp_sim(const struct orc_index_pub *ix, int v1, int v2) {
  int lo = ix->rowptr[v1], hi = ix->rowptr[v1 + 1] - 1;
  while (lo <= hi) {
    int mid = (lo + hi) >> 1;
    int c = ix->col[mid];
    if (c == v2) return ix->expsim[mid];
    if (c < v2) lo = mid + 1; else hi = mid - 1;
  }
  return 1.0;
}

static void *worker(void *arg) {
  work_t *w = (work_t *)arg;
  const struct orc_state_pub *s = w->s;
  const struct orc_model_pub *m = s->m;
  const int A = m->A, F = m->F;
  int64_t maxn = 0;
  for (int b = 0; b < w->P; ++b) if (w->bptr[b + 1] - w->bptr[b] > maxn) maxn = w->bptr[b + 1] - w->bptr[b];
  double *wt = (double *)malloc(sizeof(double) * (size_t)(maxn + 1));
  double *prob = (double *)malloc(sizeof(double) * (size_t)(maxn + 1));
  int32_t *alias = (int32_t *)malloc(sizeof(int32_t) * (size_t)(maxn + 1));
  int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (size_t)(maxn + 1));
  uint64_t rng = 0x9E3779B97F4A7C15ull ^ (uint64_t)(w->tid + 1) * 0xD1B54A32D192ED03ull;
  int64_t lo = w->n_sample * w->tid / w->nthreads, hi = w->n_sample * (w->tid + 1) / w->nthreads;
  for (int64_t i = lo; i < hi; ++i) {
    const int64_t r = w->sample[i];
    const int b = s->blk[s->link[r]];
    const int32_t *ents = w->bent + w->bptr[b];
    const int64_t n = w->bptr[b + 1] - w->bptr[b];
    const int f = s->file[r];
    int64_t nc = 0;
    if (w->sampler == ORC_PCG_II) {
      for (int64_t j = 0; j < n; ++j) { /* GU:370-393 */
        const int32_t *ye = s->y + (int64_t)ents[j] * A;
        double weight = 1.0;
        for (int a = 0; a < A; ++a) {
          const int32_t xv = s->x[r * A + a];
          if (xv < 0) continue;
          const struct orc_index_pub *ix = m->idx[a];
          const double th = s->theta[a * F + f];
          const double px = ix->phi[xv];
          if (ix->is_const) weight *= ((xv == ye[a]) ? 1.0 - th : 0.0) + th * px;
          else weight *= ((xv == ye[a]) ? 1.0 - th : 0.0) + th * px * ix->norm[ye[a]] * exp_sim(ix, xv, ye[a]);
        }
        wt[j] = weight;
      }
      nc = n;
      w->pairs += n;
      for (int64_t j = 0; j < n; ++j) cand[j] = ents[j];
    } else {
      /* getPossibleEntities GU:473-530: posting lists of the observed non-distorted attributes, smallest first */
      int obs_nd[64], n_nd = 0, obs_d[64], n_d = 0;
      const int32_t *plist[64]; int64_t plen[64];
      for (int a = 0; a < A; ++a) {
        const int32_t xv = s->x[r * A + a];
        if (xv < 0) continue;
        if (s->z[r * A + a]) { obs_d[n_d++] = a; continue; }
        const int64_t base = w->inv_ptr[(int64_t)b * A + a], end = w->inv_ptr[(int64_t)b * A + a + 1];
        int64_t l = base, h = end - 1, pos = -1;
        while (l <= h) { int64_t mid = (l + h) >> 1; int32_t v = w->inv_val[mid]; if (v == xv) { pos = mid; break; } if (v < xv) l = mid + 1; else h = mid - 1; }
        plist[n_nd] = pos >= 0 ? w->inv_post + w->inv_off[pos + ((int64_t)b * A + a)] : NULL;
        plen[n_nd] = pos >= 0 ? w->inv_off[pos + ((int64_t)b * A + a) + 1] - w->inv_off[pos + ((int64_t)b * A + a)] : 0;
        obs_nd[n_nd++] = a;
      }
      if (n_nd == 0) { for (int64_t j = 0; j < n; ++j) cand[j] = ents[j]; nc = n; }
      else {
        int best = 0;
        for (int q = 1; q < n_nd; ++q) if (plen[q] < plen[best]) best = q;
        for (int64_t j = 0; j < plen[best]; ++j) {
          const int32_t e = plist[best][j];
          int ok = 1;
          for (int q = 0; q < n_nd && ok; ++q) if (q != best) ok = (s->y[(int64_t)e * A + obs_nd[q]] == s->x[r * A + obs_nd[q]]);
          if (ok) cand[nc++] = e;
        }
      }
      w->pairs += nc;
      if (n_d == 0) { /* GU:408-411 */
        w->out[i] = nc ? cand[(int64_t)(xs_unit(&rng) * (double)nc)] : -1;
        continue;
      }
      for (int64_t j = 0; j < nc; ++j) { /* GU:414-425 */
        const int32_t *ye = s->y + (int64_t)cand[j] * A;
        double weight = 1.0;
        for (int q = 0; q < n_d; ++q) {
          const int a = obs_d[q];
          const struct orc_index_pub *ix = m->idx[a];
          const int32_t xv = s->x[r * A + a];
          if (ix->is_const) weight *= ix->phi[xv];
          else weight *= ix->norm[ye[a]] * exp_sim(ix, xv, ye[a]) * ix->phi[xv];
        }
        wt[j] = weight;
      }
    }
    /* DiscreteDist(weights).sample(): normalise + alias table + one uniform */
    if (nc > 0 && orc_alias_build(wt, (int)nc, prob, alias) == 0)
      w->out[i] = cand[orc_alias_sample(prob, alias, (int)nc, xs_unit(&rng))];
    else
      w->out[i] = -1;
  }
  free(wt); free(prob); free(alias); free(cand);
  return NULL;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Times the reference-style link update of `n_sample` evenly spaced records on `nthreads` threads.
   Returns seconds (inverted-index build for PCG-I included); pairs_out[0] = candidate pairs scored,
   pairs_out[1] = sum over ALL records of their block's entity count (pairs of one dense sweep), pairs_out[2] =
   nanoseconds of the per-sweep fixed work inside that time (the index build: it does not grow with the sample). */
double orc_refsweep_run(orc_state *state, int sampler, int P, int n_sample_i, int nthreads, uint64_t seed,
                        int64_t *pairs_out) {
  (void)seed;
  const struct orc_state_pub *s = (const struct orc_state_pub *)state;
  const int A = s->m->A;
  int64_t n_sample = n_sample_i;
  if (n_sample > s->R) n_sample = s->R;
  if (nthreads < 1) nthreads = 1;
  int64_t *bptr = (int64_t *)calloc((size_t)P + 1, sizeof(int64_t));
  for (int64_t e = 0; e < s->E; ++e) bptr[s->blk[e] + 1]++;
  for (int b = 0; b < P; ++b) bptr[b + 1] += bptr[b];
  int32_t *bent = (int32_t *)malloc(sizeof(int32_t) * (size_t)(s->E + 1));
  int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)(P + 1));
  memcpy(fill, bptr, sizeof(int64_t) * (size_t)(P + 1));
  for (int64_t e = 0; e < s->E; ++e) bent[fill[s->blk[e]]++] = (int32_t)e;
  free(fill);
  int64_t dense = 0;
  for (int64_t r = 0; r < s->R; ++r) { int b = s->blk[s->link[r]]; dense += bptr[b + 1] - bptr[b]; }
  int64_t *sample = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_sample + 1));
  for (int64_t i = 0; i < n_sample; ++i) sample[i] = i * s->R / n_sample;
  int32_t *out = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_sample + 1));

  double t0 = now_s();
  int64_t *inv_ptr = NULL, *inv_off = NULL;
  int32_t *inv_val = NULL, *inv_post = NULL;
  if (sampler != ORC_PCG_II) {
    /* EntityInvertedIndex (GU:41-76) rebuilt every sweep (GU:178-184): per (block, attr) value -> entities */
    inv_ptr = (int64_t *)calloc((size_t)P * A + 1, sizeof(int64_t));
    inv_val = (int32_t *)malloc(sizeof(int32_t) * (size_t)(s->E * A + 1));
    inv_off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(s->E * A + (int64_t)P * A + 2));
    inv_post = (int32_t *)malloc(sizeof(int32_t) * (size_t)(s->E * A + 1));
    int64_t nv = 0, np = 0;
    int64_t maxn = 0;
    for (int b = 0; b < P; ++b) if (bptr[b + 1] - bptr[b] > maxn) maxn = bptr[b + 1] - bptr[b];
    int64_t *key = (int64_t *)malloc(sizeof(int64_t) * (size_t)(maxn + 1));
    for (int b = 0; b < P; ++b)
      for (int a = 0; a < A; ++a) {
        const int64_t n = bptr[b + 1] - bptr[b];
        inv_ptr[(int64_t)b * A + a] = nv;
        for (int64_t j = 0; j < n; ++j) key[j] = ((int64_t)s->y[(int64_t)bent[bptr[b] + j] * A + a] << 32) | (uint32_t)bent[bptr[b] + j];
        /* sort by (value, entity) */
        for (int64_t gap = n / 2; gap > 0; gap /= 2)
          for (int64_t i = gap; i < n; ++i) { int64_t t = key[i], j = i; while (j >= gap && key[j - gap] > t) { key[j] = key[j - gap]; j -= gap; } key[j] = t; }
        for (int64_t j = 0; j < n; ++j) {
          int32_t v = (int32_t)(key[j] >> 32);
          if (j == 0 || v != (int32_t)(key[j - 1] >> 32)) { inv_val[nv] = v; inv_off[nv + ((int64_t)b * A + a)] = np; ++nv; }
          inv_post[np++] = (int32_t)(key[j] & 0xffffffff);
        }
        inv_off[nv + ((int64_t)b * A + a)] = np; /* sentinel for this (block, attr) */
      }
    inv_ptr[(int64_t)P * A] = nv;
    free(key);
  }
  const double t_built = now_s();  /* everything above inside the timed region is per-sweep fixed work */
  work_t *ws = (work_t *)calloc((size_t)nthreads, sizeof(work_t));
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
  for (int t = 0; t < nthreads; ++t) {
    ws[t].s = s; ws[t].sampler = sampler; ws[t].P = P; ws[t].bptr = bptr; ws[t].bent = bent;
    ws[t].inv_ptr = inv_ptr; ws[t].inv_val = inv_val; ws[t].inv_off = inv_off; ws[t].inv_post = inv_post;
    ws[t].sample = sample; ws[t].n_sample = n_sample; ws[t].tid = t; ws[t].nthreads = nthreads; ws[t].out = out;
    pthread_create(&th[t], NULL, worker, &ws[t]);
  }
  int64_t pairs = 0;
  for (int t = 0; t < nthreads; ++t) { pthread_join(th[t], NULL); pairs += ws[t].pairs; }
  double dt = now_s() - t0;
  if (pairs_out) { pairs_out[0] = pairs; pairs_out[1] = dense; pairs_out[2] = (int64_t)((t_built - t0) * 1e9); }
  free(ws); free(th); free(bptr); free(bent); free(sample); free(out);
  free(inv_ptr); free(inv_val); free(inv_off); free(inv_post);
  return dt;
}
