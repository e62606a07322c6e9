// Device half of libdblink_b200: the Gibbs sweep of dblink's record-linkage model on one B200.
//
// Reference (cleanzr/dblink @ dc3dd0d, src/main/scala/com/github/cleanzr/dblink/, GU = GibbsUpdates.scala):
//   State.nextState State.scala:78-99 -> updatePartition GU:156-211
//     link draw per record        GU:363-395 (PCG-II), GU:399-466 (PCG-I / Gibbs, dense form)   -> k_link
//     entity values per (e, attr) GU:534-599, 605-646, 702-755                                   -> k_values
//     distortions per (r, attr)   GU:324-359                                                     -> k_dist
//     new partition id            GU:206, partitioning/MutableBST.scala:61-79                    -> k_entity_post
//   updateSummaryVariables GU:219-301                                   -> k_entity_post / k_dist accumulators
//   the shuffle GU:144                                                  -> relayout() (sort by block id)
//
// Numerics: all likelihood arithmetic is IEEE binary64 with no FMA contraction (-fmad=false) so that the
// CPU oracle (oracle/dbl_oracle.c, -ffp-contract=off) reproduces every draw bit for bit.  Draw protocol:
// DESIGN.md section 4.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/cub.cuh>
#include <string>
#include <vector>

#include "dbl_internal.h"
#define DBL_ENGINE_TU 1
#include "dbl_link.cuh"
#include "dbl_link_pcg2.cuh"

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ctx->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                       \
      return DBL_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// device-side model
// ---------------------------------------------------------------------------------------------------
struct TreeDev {
  int n_nodes;
  const int *attr, *kind, *split, *set_ptr, *set_val, *leaf_no;
};

// ---------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int invcdf(const double *cdf, int V, double u) {
  int lo = 0, hi = V;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] > u) hi = mid; else lo = mid + 1;
  }
  return lo < V ? lo : V - 1;
}

__device__ __forceinline__ int tree_leaf(const TreeDev &t, const int *yrow) {
  int node = 0;
  while (node < t.n_nodes && t.attr[node] >= 0) {
    const int v = yrow[t.attr[node]];
    bool right;
    if (t.kind[node]) {
      int lo = t.set_ptr[node], hi = t.set_ptr[node + 1] - 1;
      right = false;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int s = t.set_val[mid];
        if (s == v) { right = true; break; }
        if (s < v) lo = mid + 1; else hi = mid - 1;
      }
    } else {
      right = v > t.split[node];
    }
    node = right ? 2 * node + 2 : 2 * node + 1;
  }
  return t.leaf_no[node];
}

// ---------------------------------------------------------------------------------------------------
// k_entity_post: per entity N(e) = prod_{non-const a} n_a(y_a); new block id (GU:206); entity part of the
// summary: isolates (GU:267-269) and sum_a log phi_a(y_a) (GU:234-237, 271-274).
// ---------------------------------------------------------------------------------------------------
__global__ void k_entity_post(int64_t E, int A, const int *__restrict__ y, const AttrDev *__restrict__ attrs,
                              TreeDev tree, double *__restrict__ entN, int *__restrict__ blk,
                              const int *__restrict__ ent_rec_ptr, long long *__restrict__ counts, int iso_slot,
                              double *__restrict__ loglik, const unsigned char *__restrict__ ent_owned) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double ll = 0.0;
  int iso = 0;
  if (e < E && ent_owned[e]) {
    const int *ye = y + e * A;
    double n = 1.0;
    for (int a = 0; a < A; ++a) {
      const AttrDev &at = attrs[a];
      const int v = ye[a];
      if (!at.is_const) n = n * at.norm[v];
      ll += at.logphi[v];
    }
    entN[e] = n;
    blk[e] = tree.n_nodes > 0 ? tree_leaf(tree, ye) : 0;
    if (ent_rec_ptr) iso = (ent_rec_ptr[e] == ent_rec_ptr[e + 1]);
  }
  if (counts) {
    typedef cub::BlockReduce<double, 256> BR;
    typedef cub::BlockReduce<int, 256> BRI;
    __shared__ typename BR::TempStorage t1;
    __shared__ typename BRI::TempStorage t2;
    const double s = BR(t1).Sum(ll);
    const int c = BRI(t2).Sum(iso);
    if (threadIdx.x == 0) {
      atomicAdd(loglik, s);
      if (c) atomicAdd((unsigned long long *)&counts[iso_slot], (unsigned long long)c);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// layout ("re-partitioning", replaces the shuffle GU:144): entities and records grouped by block
// ---------------------------------------------------------------------------------------------------
__global__ void k_iota(int64_t n, int *out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)i;
}
__global__ void k_hist(int64_t n, const int *__restrict__ key, int *__restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&cnt[key[i]], 1);
}
// sort keys; entities / records this rank does not own go to the dummy slot (block P, entity E)
// Records are grouped by block and, within a block, by a static cost class, so that the records sharing a CTA of
// the link kernel (and its tile ring) advance at the same pace.  The order of records inside a block has no effect
// on the draws (every record has its own counter-based stream).
constexpr int REC_CLASS_BITS = 8;
__global__ void k_rec_block_keys(int64_t R, const int *__restrict__ link, const int *__restrict__ blk,
                                 const unsigned char *__restrict__ rec_owned,
                                 const unsigned char *__restrict__ rec_class, int P, int *__restrict__ key) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) key[r] = rec_owned[r] ? ((blk[link[r]] << REC_CLASS_BITS) | rec_class[r]) : (P << REC_CLASS_BITS);
}
// cost class of a record (x is static): bits 7..6 = number of missing non-constant attributes (each one adds a
// gather per candidate), bits 5..0 = expected number of similar-but-different candidate values per 32-candidate
// step (how often the warp takes the similarity multiply), from the empirical value frequencies.
__global__ void k_rec_class(int64_t R, int A, const AttrDev *__restrict__ attrs, const int *__restrict__ x,
                            unsigned char *__restrict__ cls) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  int miss = 0;
  double h = 0.0;
  for (int a = 0; a < A; ++a) {
    const AttrDev &at = attrs[a];
    if (at.is_const) continue;
    const int xv = x[r * A + a];
    if (xv < 0) { ++miss; continue; }
    for (int i = at.rowptr[xv]; i < at.rowptr[xv + 1]; ++i)
      if (at.col[i] != xv) h += at.probs[at.col[i]];
  }
  const int hq = min(63, (int)(h * 32.0 * 8.0));
  cls[r] = (unsigned char)((min(miss, 3) << 6) | hq);
}
__global__ void k_ent_block_keys(int64_t E, const int *__restrict__ blk, const unsigned char *__restrict__ ent_owned,
                                 int P, int *__restrict__ key) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) key[e] = ent_owned[e] ? blk[e] : P;
}
__global__ void k_rec_link_keys(int64_t R, const int *__restrict__ link, const unsigned char *__restrict__ rec_owned,
                                int E, int *__restrict__ key) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) key[r] = rec_owned[r] ? link[r] : E;
}
// offsets from sorted keys: ptr[k] = first position whose key is >= k, for k = 0..n_keys (keys beyond the data
// point at n).  One pass over the sorted array; replaces an atomic histogram + scan.
__global__ void k_segment_ptr(int64_t n, int n_keys, const int *__restrict__ sorted_key, int *__restrict__ ptr,
                              int shift = 0) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const int cur = (i < n) ? min(sorted_key[i] >> shift, n_keys) : n_keys;
  const int prev = (i > 0) ? min(sorted_key[i - 1] >> shift, n_keys) : -1;
  for (int k = prev + 1; k <= cur; ++k) ptr[k] = (int)i;
}
// prefix sums over the P blocks (P is 2^numLevels: small) -- one CTA, Hillis-Steele over chunks
__global__ void k_block_scan(int P, const int *__restrict__ ent_ptr, const int *__restrict__ rec_ptr,
                             int *__restrict__ tile_ptr, int *__restrict__ cta_ptr, int warps_per_cta,
                             int *__restrict__ cta_ptr2, int warps_per_cta2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int t = 0, c = 0, c2 = 0;
    for (int b = 0; b < P; ++b) {
      tile_ptr[b] = t; cta_ptr[b] = c; cta_ptr2[b] = c2;
      t += (ent_ptr[b + 1] - ent_ptr[b] + TE - 1) / TE;
      const int nr = rec_ptr[b + 1] - rec_ptr[b];
      c += (nr + warps_per_cta - 1) / warps_per_cta;
      c2 += (nr + warps_per_cta2 - 1) / warps_per_cta2;
    }
    tile_ptr[P] = t; cta_ptr[P] = c; cta_ptr2[P] = c2;
  }
}
// tiled, block-sorted copy of the entity table: tile = { int32 y[A][TE]; double N[TE]; uint32 packed_consts[TE] }
__global__ void k_build_tiles(int64_t E, int A, const int *__restrict__ y, const double *__restrict__ entN,
                              const int *__restrict__ blk_sorted, const int *__restrict__ ent_sorted,
                              const int *__restrict__ ent_ptr, const int *__restrict__ tile_ptr,
                              int *__restrict__ tiles, const int *__restrict__ perm, int P, int npack) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E) return;
  const int b = blk_sorted[i];
  if (b >= P) return;  // not owned by this rank
  const int e = ent_sorted[i];
  const int j = (int)(i - ent_ptr[b]);
  int *tile = tiles + (size_t)(tile_ptr[b] + j / TE) * tile_words(A);
  const int slot = j % TE;
  for (int k = 0; k < A; ++k) tile[k * TE + slot] = y[(int64_t)e * A + perm[k]];  // kernel order
  reinterpret_cast<double *>(tile + (size_t)A * TE)[slot] = entN[e];
  unsigned pk = 0;  // constant attributes (kernel positions 0..npack-1), one byte each, for k_link_pcg2
  for (int k = 0; k < npack; ++k) pk |= ((unsigned)y[(int64_t)e * A + perm[k]] & 0xFFu) << (8 * k);
  tile[(size_t)(A + 2) * TE + slot] = (int)pk;
}

// ---------------------------------------------------------------------------------------------------
// k_values: one thread per (entity, attribute).  updateEntityValueCollapsed GU:576-599 +
// perturbedDistYCollapsed GU:534-570 (PCG-I/II); updateEntityValue GU:605-646 + perturbedDistY GU:702-727.
// ---------------------------------------------------------------------------------------------------
struct ValParams {
  int A, F, sampler;
  uint64_t seed;
  uint32_t iter;
  int64_t E;
  const AttrDev *attrs;
  const int *x, *file;
  const unsigned *zmask;
  const double *theta;
  const int *ent_rec_ptr, *rec_by_ent;
  const unsigned char *ent_owned;
  int *y;
};

struct BaseDist {
  const AttrDev *at;
  int k;
  const double *pk, *cdf;
  double zk;
};
__device__ __forceinline__ double base_weight(const AttrDev &at, int k, int v) {
  double w = at.probs[v];
  if (!at.is_const)
    for (int i = 0; i < k; ++i) w = w * at.norm[v];
  return w;
}
__device__ void base_init(BaseDist &b, const AttrDev &at, int k) {
  b.at = &at;
  if (at.is_const) k = 0;
  b.k = k;
  if (k <= at.kmax) {
    b.pk = at.pk + (size_t)k * at.V;
    b.cdf = at.cdf + (size_t)k * at.V;
    b.zk = 0.0;
  } else {  // beyond the cached powers (getSimNormDist cache miss, AttributeIndex.scala:199-205)
    b.pk = nullptr; b.cdf = nullptr;
    double z = 0.0;
    for (int v = 0; v < at.V; ++v) z += base_weight(at, k, v);
    b.zk = z;
  }
}
__device__ __forceinline__ double base_prob(const BaseDist &b, int v) {
  return b.pk ? b.pk[v] : base_weight(*b.at, b.k, v) / b.zk;
}
__device__ int base_draw(const BaseDist &b, double u) {
  if (b.cdf) return invcdf(b.cdf, b.at->V, u);
  double c = 0.0;
  for (int v = 0; v < b.at->V; ++v) {
    c += base_weight(*b.at, b.k, v) / b.zk;
    if (c > u) return v;
  }
  return b.at->V - 1;
}

// factor contributed by record r to candidate value v; false when v is outside the record's support
__device__ __forceinline__ bool g_factor(const ValParams &p, const AttrDev &at, int a, int r, bool collapsed, int v,
                                         double &g) {
  const int xr = p.x[(int64_t)r * p.A + a];
  if (at.is_const) {
    if (v != xr || !collapsed) return false;
    const double th = p.theta[a * p.F + p.file[r]];
    g = 1.0 + (1.0 / th - 1.0) / at.phi[xr];  // GU:553
    return true;
  }
  double e;
  if (!row_find(at, xr, v, e)) return false;
  if (collapsed && v == xr) {
    const double th = p.theta[a * p.F + p.file[r]];
    g = e + (1.0 / th - 1.0) / (at.phi[xr] * at.norm[xr]);  // GU:557,560
  } else {
    g = e;
  }
  return true;
}
__device__ __forceinline__ bool in_support(const ValParams &p, const AttrDev &at, int a, int r, int v) {
  const int xr = p.x[(int64_t)r * p.A + a];
  if (at.is_const) return v == xr;
  double e;
  return row_find(at, xr, v, e);
}

__global__ void k_values(ValParams p) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= p.E * p.A) return;
  const int64_t e = tid / p.A;
  if (!p.ent_owned[e]) return;
  const int a = (int)(tid % p.A);
  const AttrDev &at = p.attrs[a];
  const bool collapsed = (p.sampler == DBL_PCG_I || p.sampler == DBL_PCG_II);
  const U2 u = uniform2(p.seed, PH_VALUE, p.iter, (uint32_t)e, (uint32_t)a);
  const int lo = p.ent_rec_ptr[e], hi = p.ent_rec_ptr[e + 1];
  const int *rec = p.rec_by_ent;

  int k = 0;
  for (int i = lo; i < hi; ++i) k += (p.x[(int64_t)rec[i] * p.A + a] >= 0);
  BaseDist b;
  if (k == 0) {  // GU:588-589
    base_init(b, at, 0);
    p.y[tid] = base_draw(b, u.u1);
    return;
  }
  if (!collapsed) {
    for (int i = lo; i < hi; ++i) {  // GU:619-630
      const int r = rec[i];
      const int xr = p.x[(int64_t)r * p.A + a];
      if (xr >= 0 && !((p.zmask[r] >> a) & 1u)) { p.y[tid] = xr; return; }
    }
    if (at.is_const) {  // GU:633-634
      base_init(b, at, 0);
      p.y[tid] = base_draw(b, u.u1);
      return;
    }
  }
  base_init(b, at, k);  // GU:584-586
  double total = 0.0, target = 0.0, cum = 0.0;
  int picked = -1, last_pos = -1;
  if (hi - lo == 1) {
    // one linked record (most entities): the candidate values are the record's own similarity row, so the factors
    // come straight from that row (no searches, no first-appearance check); same operations as the general path
    const int r = rec[lo];
    const int xr = p.x[(int64_t)r * p.A + a];  // observed (k == 1)
    const int q0 = at.is_const ? 0 : at.rowptr[xr];
    const int nv = at.is_const ? 1 : (at.rowptr[xr + 1] - q0);
    double extra = 0.0;  // what the collapsed update adds to the factor of v == x (GU:553,557,560)
    if (collapsed) {
      const double th = p.theta[a * p.F + p.file[r]];
      extra = at.is_const ? (1.0 / th - 1.0) / at.phi[xr] : (1.0 / th - 1.0) / (at.phi[xr] * at.norm[xr]);
    }
    auto weight = [&](int q, int &v) -> double {
      double G = 1.0;
      if (at.is_const) {
        v = xr;
        if (collapsed) G = G * (1.0 + extra);
      } else {
        v = at.col[q0 + q];
        const double e = at.expsim[q0 + q];
        G = G * ((collapsed && v == xr) ? e + extra : e);
      }
      return base_prob(b, v) * (G - 1.0);  // GU:567 / 724
    };
    for (int q = 0; q < nv; ++q) {
      int v;
      total += weight(q, v);
    }
    if (u.u0 < 1.0 / (1.0 + total)) {  // GU:593-594
      p.y[tid] = base_draw(b, u.u1);
      return;
    }
    target = u.u1 * total;
    for (int q = 0; q < nv && picked < 0; ++q) {
      int v;
      const double W = weight(q, v);
      cum += W;
      if (W > 0.0) last_pos = v;
      if (cum > target) picked = v;
    }
    if (picked < 0) picked = last_pos;
    if (picked < 0) picked = base_draw(b, u.u1);
    p.y[tid] = picked;
    return;
  }
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = lo; i < hi; ++i) {
      const int r = rec[i];
      const int xr = p.x[(int64_t)r * p.A + a];
      if (xr < 0) continue;
      const int nv = at.is_const ? 1 : (at.rowptr[xr + 1] - at.rowptr[xr]);
      const int *vals = at.is_const ? nullptr : (at.col + at.rowptr[xr]);
      for (int q = 0; q < nv; ++q) {
        const int v = at.is_const ? xr : vals[q];
        bool seen = false;
        for (int j = lo; j < i && !seen; ++j) {
          const int rj = rec[j];
          if (p.x[(int64_t)rj * p.A + a] < 0) continue;
          seen = in_support(p, at, a, rj, v);
        }
        if (seen) continue;
        double G = 1.0;
        for (int j = i; j < hi; ++j) {
          const int rj = rec[j];
          if (p.x[(int64_t)rj * p.A + a] < 0) continue;
          double g;
          if (g_factor(p, at, a, rj, collapsed, v, g)) G = G * g;
        }
        const double W = base_prob(b, v) * (G - 1.0);  // GU:567 / 724
        if (pass == 0) {
          total += W;
        } else {
          cum += W;
          if (W > 0.0) last_pos = v;
          if (picked < 0 && cum > target) picked = v;
        }
      }
    }
    if (pass == 0) {
      if (u.u0 < 1.0 / (1.0 + total)) {  // GU:593-594
        p.y[tid] = base_draw(b, u.u1);
        return;
      }
      target = u.u1 * total;
    }
  }
  if (picked < 0) picked = last_pos;
  if (picked < 0) picked = base_draw(b, u.u1);
  p.y[tid] = picked;
}

// ---------------------------------------------------------------------------------------------------
// k_dist: one thread per record.  updateDistortions GU:324-359 with the new y, fused with the record part
// of updateSummaryVariables GU:239-266.  draw = 0 only accumulates the summary of the current state.
// ---------------------------------------------------------------------------------------------------
struct DistParams {
  int A, F, draw;
  uint64_t seed;
  uint32_t iter;
  int64_t R;
  const AttrDev *attrs;
  const int *x, *file, *link, *y;
  unsigned *zmask;
  const double *theta;
  long long *counts;  // [A*F] aggDist, then [A+1] recDist
  double *loglik;
  const unsigned char *rec_owned;
};

__global__ void k_dist(DistParams p) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double ll = 0.0;
  if (r < p.R && p.rec_owned[r]) {
    const int f = p.file[r];
    const int *ye = p.y + (int64_t)p.link[r] * p.A;
    unsigned zm = p.zmask[r];
    int nd = 0;
    for (int a = 0; a < p.A; ++a) {
      const AttrDev &at = p.attrs[a];
      const int xv = p.x[r * p.A + a];
      const int yv = ye[a];
      bool z;
      if (p.draw) {
        const double th = p.theta[a * p.F + f];
        if (xv < 0) {
          const U2 u = uniform2(p.seed, PH_DIST, p.iter, (uint32_t)r, (uint32_t)a);
          z = u.u0 < th;  // GU:331-334
        } else if (xv != yv) {
          z = true;  // GU:352-354
        } else {
          const U2 u = uniform2(p.seed, PH_DIST, p.iter, (uint32_t)r, (uint32_t)a);
          double pr1 = th * at.phi[xv];
          if (!at.is_const) {
            double ediag = 1.0;
            row_find(at, xv, xv, ediag);
            pr1 = pr1 * at.norm[xv];
            pr1 = pr1 * ediag;
          }
          const double pr0 = 1.0 - th;
          const double den = pr1 + pr0;
          const double pz = (den != 0.0) ? pr1 / den : 0.0;  // GU:349-350
          z = u.u0 < pz;
        }
        zm = z ? (zm | (1u << a)) : (zm & ~(1u << a));
      } else {
        z = (zm >> a) & 1u;
      }
      if (z) {
        ++nd;
        atomicAdd((unsigned long long *)&p.counts[a * p.F + f], 1ull);  // GU:246
        if (xv >= 0) {                                                   // GU:248-258
          ll += at.logphi[xv];
          if (!at.is_const) {
            ll += at.lognorm[yv];
            double e;
            if (row_find(at, xv, yv, e)) ll += log(e);
          }
        }
      }
    }
    if (p.draw) p.zmask[r] = zm;
    // GU:265 -- warp-aggregated: one atomic per distinct count in the warp
    const unsigned peers = __match_any_sync(__activemask(), nd);
    if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31))
      atomicAdd((unsigned long long *)&p.counts[p.A * p.F + nd], (unsigned long long)__popc(peers));
  }
  typedef cub::BlockReduce<double, 256> BR;
  __shared__ typename BR::TempStorage tmp;
  const double s = BR(tmp).Sum(ll);
  if (threadIdx.x == 0 && s != 0.0) atomicAdd(p.loglik, s);
}

// ---------------------------------------------------------------------------------------------------
// k_init_state: State.deterministic (State.scala:253-301) -- entity e copies record e (if any), missing
// values drawn from phi; record r links to entity r mod E; z = (x>=0 && x!=y).
// ---------------------------------------------------------------------------------------------------
__global__ void k_init_entities(int64_t E, int64_t R, int A, uint64_t seed, const AttrDev *__restrict__ attrs,
                                const int *__restrict__ x, int *__restrict__ y) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= E * A) return;
  const int64_t e = tid / A;
  const int a = (int)(tid % A);
  int v = (e < R) ? x[e * A + a] : -1;
  if (v < 0) {
    const U2 u = uniform2(seed, PH_INIT, 0u, (uint32_t)e, (uint32_t)a);
    v = invcdf(attrs[a].cdf, attrs[a].V, u.u1);
  }
  y[tid] = v;
}
__global__ void k_init_records(int64_t E, int64_t R, int A, const int *__restrict__ x, const int *__restrict__ y,
                               int *__restrict__ link, unsigned *__restrict__ zmask) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int64_t e = r % E;
  link[r] = (int)e;
  unsigned zm = 0;
  for (int a = 0; a < A; ++a) {
    const int xv = x[r * A + a];
    if (xv >= 0 && xv != y[e * A + a]) zm |= 1u << a;
  }
  zmask[r] = zm;
}
// range checks of an uploaded state (the reference would fail with ArrayIndexOutOfBounds / require):
// bit 0 record value id, bit 1 file id, bit 2 link, bit 3 entity value id
__global__ void k_validate(int64_t R, int64_t E, int A, int F, const AttrDev *__restrict__ attrs,
                           const int *__restrict__ x, const int *__restrict__ file, const int *__restrict__ link,
                           const int *__restrict__ y, int *__restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int bad = 0;
  if (i < R) {
    for (int a = 0; a < A; ++a) {
      const int v = x[i * A + a];
      if (v < -1 || v >= attrs[a].V) bad |= 1;
    }
    if (file[i] < 0 || file[i] >= F) bad |= 2;
    if (link && (link[i] < 0 || link[i] >= E)) bad |= 4;
  }
  if (y && i < E)
    for (int a = 0; a < A; ++a) {
      const int v = y[i * A + a];
      if (v < 0 || v >= attrs[a].V) bad |= 8;
    }
  if (bad) atomicOr(flag, bad);
}
__global__ void k_pack_z(int64_t R, int A, const uint8_t *__restrict__ z, unsigned *__restrict__ zmask) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  unsigned zm = 0;
  for (int a = 0; a < A; ++a) zm |= (z[r * A + a] ? 1u : 0u) << a;
  zmask[r] = zm;
}
__global__ void k_unpack_z(int64_t R, int A, const unsigned *__restrict__ zmask, uint8_t *__restrict__ z) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const unsigned zm = zmask[r];
  for (int a = 0; a < A; ++a) z[r * A + a] = (zm >> a) & 1u;
}

// ---------------------------------------------------------------------------------------------------
// multi-GPU: ownership and the exchange of clusters whose new block belongs to another rank (replaces the
// shuffle `.partitionBy(partitioner)`, GU:144).  Messages: entity = [e, y_0..y_{A-1}], record = [r, e, zmask].
// ---------------------------------------------------------------------------------------------------
__global__ void k_mark_ent_owned(int64_t E, const int *__restrict__ blk, const int *__restrict__ owner, int rank,
                                 unsigned char *__restrict__ ent_owned) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) ent_owned[e] = (owner[blk[e]] == rank);
}
__global__ void k_mark_ent_block(int64_t E, const int *__restrict__ blk, int block, unsigned char *__restrict__ ent_owned) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) ent_owned[e] = (blk[e] == block);
}
__global__ void k_mark_rec_owned(int64_t R, const int *__restrict__ link, const unsigned char *__restrict__ ent_owned,
                                 unsigned char *__restrict__ rec_owned) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) rec_owned[r] = ent_owned[link[r]];
}
__global__ void k_move_count_ent(int64_t E, const int *__restrict__ blk, const int *__restrict__ owner, int rank,
                                 const unsigned char *__restrict__ ent_owned, int *__restrict__ ent_dest,
                                 unsigned long long *__restrict__ cnt) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int d = -1;
  if (ent_owned[e]) {
    const int o = owner[blk[e]];
    if (o != rank) { d = o; atomicAdd(&cnt[o], 1ull); }
  }
  ent_dest[e] = d;
}
__global__ void k_move_count_rec(int64_t R, const int *__restrict__ link, const unsigned char *__restrict__ rec_owned,
                                 const int *__restrict__ ent_dest, unsigned long long *__restrict__ cnt) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R || !rec_owned[r]) return;
  const int d = ent_dest[link[r]];
  if (d >= 0) atomicAdd(&cnt[d], 1ull);
}
__global__ void k_move_pack_ent(int64_t E, int A, const int *__restrict__ y, const int *__restrict__ ent_dest,
                                unsigned char *__restrict__ ent_owned, unsigned long long *__restrict__ cursor,
                                int *__restrict__ buf) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int d = ent_dest[e];
  if (d < 0) return;
  const unsigned long long slot = atomicAdd(&cursor[d], 1ull);
  int *m = buf + slot * (A + 1);
  m[0] = (int)e;
  for (int a = 0; a < A; ++a) m[1 + a] = y[e * A + a];
  ent_owned[e] = 0;
}
__global__ void k_move_pack_rec(int64_t R, const int *__restrict__ link, const unsigned *__restrict__ zmask,
                                const int *__restrict__ ent_dest, unsigned char *__restrict__ rec_owned,
                                unsigned long long *__restrict__ cursor, int *__restrict__ buf) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R || !rec_owned[r]) return;
  const int d = ent_dest[link[r]];
  if (d < 0) return;
  const unsigned long long slot = atomicAdd(&cursor[d], 1ull);
  int *m = buf + slot * 3;
  m[0] = (int)r; m[1] = link[r]; m[2] = (int)zmask[r];
  rec_owned[r] = 0;
}
__global__ void k_merge_links(int64_t R, const unsigned char *__restrict__ rec_owned, const int *__restrict__ old_link,
                              int *__restrict__ new_link) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R && !rec_owned[r]) new_link[r] = old_link[r];
}
__global__ void k_unpack_ent(int64_t n, int A, const int *__restrict__ buf, const AttrDev *__restrict__ attrs,
                             TreeDev tree, int *__restrict__ y, double *__restrict__ entN, int *__restrict__ blk,
                             unsigned char *__restrict__ ent_owned) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int *m = buf + i * (A + 1);
  const int64_t e = m[0];
  double nn = 1.0;
  for (int a = 0; a < A; ++a) {
    const int v = m[1 + a];
    y[e * A + a] = v;
    if (!attrs[a].is_const) nn = nn * attrs[a].norm[v];
  }
  entN[e] = nn;
  blk[e] = tree.n_nodes > 0 ? tree_leaf(tree, m + 1) : 0;
  ent_owned[e] = 1;
}
__global__ void k_unpack_rec(int64_t n, const int *__restrict__ buf, int *__restrict__ link,
                             unsigned *__restrict__ zmask, unsigned char *__restrict__ rec_owned) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int *m = buf + i * 3;
  const int64_t r = m[0];
  link[r] = m[1];
  zmask[r] = (unsigned)m[2];
  rec_owned[r] = 1;
}

// owned rows of the state into caller-provided device buffers, zeros elsewhere (the host sums them over ranks)
__global__ void k_export_ent(int64_t E, int A, const unsigned char *__restrict__ owned, const int *__restrict__ y,
                             const int *__restrict__ blk, int *__restrict__ y_out, int *__restrict__ blk_out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const bool o = owned[e];
  for (int a = 0; a < A; ++a) y_out[e * A + a] = o ? y[e * A + a] : 0;
  blk_out[e] = o ? blk[e] : 0;
}
__global__ void k_export_rec(int64_t R, int A, const unsigned char *__restrict__ owned, const int *__restrict__ link,
                             const unsigned *__restrict__ zmask, int *__restrict__ link_out,
                             unsigned char *__restrict__ z_out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const bool o = owned[r];
  link_out[r] = o ? link[r] : 0;
  const unsigned zm = o ? zmask[r] : 0u;
  for (int a = 0; a < A; ++a) z_out[r * A + a] = (zm >> a) & 1u;
}

// ---------------------------------------------------------------------------------------------------
// host-side context
// ---------------------------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    return cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T));
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  ~DevBuf() { release(); }
};

struct dbl_ctx {
  int A = 0, F = 0, P = 1, device = 0;
  uint64_t seed = 0;
  int rank = 0, world = 1;
  std::vector<double> alpha, beta;
  std::string err;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  // model (device)
  std::vector<DevBuf<double>> dtab;  // per-attr double tables
  std::vector<DevBuf<int>> itab;
  DevBuf<AttrDev> attrs;
  DevBuf<int> perm_dev;
  int perm[DBL_MAX_ATTRS] = {0};
  int n_str = 0;       // non-constant attributes
  int pack_consts = 0; // constant attributes byte-packed into the tiles (0 = none)
  int hslots = 32, hshift = 27;  // common hash-table size of the non-constant attributes; hslots = 0: none
  std::vector<AttrDev> h_attrs;
  DevBuf<int> tree_buf;
  TreeDev tree{};

  // state
  int64_t R = 0, E = 0, iteration = 0;
  bool has_state = false;
  DevBuf<int> x, file, link, newlink, y, blk;
  DevBuf<uint8_t> zbytes;  // staging for the byte-per-flag host format of z
  DevBuf<int> vflag, file_cnt;
  DevBuf<unsigned> zmask;
  DevBuf<double> entN, theta;
  std::vector<double> h_theta;
  std::vector<int64_t> file_sizes;

  // ownership (multi-GPU sharding by block)
  std::vector<int> owner_h;  // P entries; default: everything owned by this rank
  DevBuf<int> owner, ent_dest, ent_key, link_key;
  DevBuf<unsigned char> ent_owned, rec_owned;
  DevBuf<unsigned long long> move_cnt;  // [2*world] counts then [2*world] cursors
  std::vector<int64_t> h_move_ent, h_move_rec;
  bool in_sweep = false, in_block_sweep = false;
  int block_sampler = 0;
  std::vector<char> block_done;
  DevBuf<int> blk_frozen;

  // layout
  DevBuf<int> iota, blk_sorted, ent_sorted, rec_key, rec_key_sorted, rec_sorted, ent_cnt, rec_cnt;
  DevBuf<int> ent_ptr, tile_ptr, rec_ptr, cta_ptr, cta_ptr2, tiles;
  // inverted index of the block tables for the pruned PCG-I link kernel (built on demand, once per sweep)
  DevBuf<unsigned long long> inv_key_in, inv_key;
  DevBuf<int> inv_pos_in, inv_pos, inv_seg, inv_vptr;
  InvDense inv_dense;
  bool inv_use_dense = false;
  DevBuf<unsigned char> inv_tmp;
  size_t inv_tmp_bytes = 0;
  bool inv_valid = false;
  int inv_vbits = 32;
  DevBuf<int> link_sorted, rec_by_ent, ent_rec_cnt, ent_rec_ptr;
  DevBuf<unsigned char> rec_class;  // static cost class of a record (k_rec_class)
  DevBuf<unsigned char> cub_tmp;
  size_t cub_bytes = 0;
  int max_ctas = 0;

  // summary
  DevBuf<long long> counts;  // A*F + (A+1) + 2
  DevBuf<double> loglik;
  DevBuf<int> status;
  DevBuf<unsigned long long> pairs;
  std::vector<long long> h_counts;
  double h_loglik_part = 0.0;
  int64_t h_pairs = 0;
  int h_owned_ent = -1;  // entities in the blocks this rank owns (= ent_ptr[P]) after the last relayout; -1 = unknown

  int64_t launches = 0;
  double link_ms = 0.0;
  int link_mode = 0;  // 0 auto, 1 generic kernel everywhere, 2 dense TMA kernels everywhere (no pruning)
  double last_sweep_ms = 0.0;
  int64_t link_launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending_events;

  void set_error(const std::string &s) { err = s; }
  int n_counts() const { return A * F + (A + 1) + 2; }
  int iso_slot() const { return A * F + (A + 1); }
};

static inline int grid_for(int64_t n, int bs) { return (int)std::max<int64_t>(1, (n + bs - 1) / bs); }

static int upload_tree(dbl_ctx *ctx, const dbl_kdtree *t) {
  if (t) {
    const int n = t->n_nodes;
    std::vector<int> pack;
    pack.insert(pack.end(), t->attr.begin(), t->attr.end());
    pack.insert(pack.end(), t->kind.begin(), t->kind.end());
    pack.insert(pack.end(), t->split.begin(), t->split.end());
    pack.insert(pack.end(), t->set_ptr.begin(), t->set_ptr.end());
    pack.insert(pack.end(), t->leaf_no.begin(), t->leaf_no.end());
    pack.insert(pack.end(), t->set_val.begin(), t->set_val.end());
    for (int i = 0; i < n; ++i)
      if (t->attr[i] >= ctx->A) { ctx->set_error("partitioner attribute id out of range"); return DBL_ERR_INVALID; }
    CUDA_TRY(ctx->tree_buf.alloc(pack.size()));
    CUDA_TRY(cudaMemcpy(ctx->tree_buf.p, pack.data(), pack.size() * sizeof(int), cudaMemcpyHostToDevice));
    int *base = ctx->tree_buf.p;
    ctx->tree.n_nodes = n;
    ctx->tree.attr = base;
    ctx->tree.kind = base + n;
    ctx->tree.split = base + 2 * n;
    ctx->tree.set_ptr = base + 3 * n;
    ctx->tree.leaf_no = base + 3 * n + (n + 1);
    ctx->tree.set_val = base + 4 * n + (n + 1);
    ctx->P = t->n_leaves;
  } else {
    ctx->tree.n_nodes = 0;
    ctx->P = 1;
  }
  return DBL_OK;
}

static int upload_model(dbl_ctx *ctx, const dbl_model_desc *d) {
  const int A = d->num_attrs;
  ctx->h_attrs.resize(A);
  ctx->dtab.resize((size_t)A * 10);
  ctx->itab.resize((size_t)A * 4);
  auto up_d = [&](DevBuf<double> &b, const std::vector<double> &v) -> cudaError_t {
    cudaError_t e = b.alloc(v.size());
    if (e != cudaSuccess) return e;
    if (v.empty()) return cudaSuccess;
    return cudaMemcpy(b.p, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice);
  };
  auto up_i = [&](DevBuf<int> &b, const std::vector<int32_t> &v) -> cudaError_t {
    cudaError_t e = b.alloc(v.size());
    if (e != cudaSuccess) return e;
    if (v.empty()) return cudaSuccess;
    return cudaMemcpy(b.p, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice);
  };
  // one hash-table size for the whole model: re-hash the smaller tables to the largest
  int Hmax = 32;
  bool hash_ok = true;
  for (int a = 0; a < A; ++a) {
    if (d->indexes[a]->is_const) continue;
    if (d->indexes[a]->hsize <= 0) hash_ok = false;
    Hmax = std::max(Hmax, d->indexes[a]->hsize);
  }
  std::vector<dbl_index> rehashed(A);
  for (int a = 0; a < A; ++a) {
    const dbl_index *ix = d->indexes[a];
    if (hash_ok && !ix->is_const && ix->hsize != Hmax) {
      rehashed[a] = *ix;
      rehashed[a].build_hash(Hmax);
      if (rehashed[a].hsize != Hmax) hash_ok = false;
      ix = &rehashed[a];
    }
    DevBuf<double> *t = &ctx->dtab[(size_t)a * 10];
    DevBuf<int> *ti = &ctx->itab[(size_t)a * 4];
    CUDA_TRY(up_d(t[0], ix->phi));
    CUDA_TRY(up_d(t[1], ix->probs));
    CUDA_TRY(up_d(t[2], ix->norm));
    CUDA_TRY(up_d(t[3], ix->invnorm));
    CUDA_TRY(up_d(t[4], ix->pk));
    CUDA_TRY(up_d(t[5], ix->cdf));
    CUDA_TRY(up_d(t[6], ix->logphi));
    CUDA_TRY(up_d(t[7], ix->lognorm));
    CUDA_TRY(up_d(t[8], ix->expsim));
    CUDA_TRY(up_i(ti[0], ix->rowptr));
    CUDA_TRY(up_i(ti[1], ix->col));
    CUDA_TRY(up_d(t[9], ix->hvals));
    CUDA_TRY(up_i(ti[2], ix->hkeys));
    {
      std::vector<int32_t> hm(ix->hmult.begin(), ix->hmult.end());
      CUDA_TRY(up_i(ti[3], hm));
    }
    AttrDev &h = ctx->h_attrs[a];
    h.V = ix->V; h.is_const = ix->is_const ? 1 : 0; h.kmax = ix->kmax; h.hsize = ix->hsize;
    h.hshift = ix->hshift; h.pad0 = h.pad1 = h.pad2 = 0;
    h.hvals = t[9].p; h.hkeys = ti[2].p; h.hmult = reinterpret_cast<const unsigned *>(ti[3].p);
    h.phi = t[0].p; h.probs = t[1].p; h.norm = t[2].p; h.invnorm = t[3].p; h.pk = t[4].p; h.cdf = t[5].p;
    h.logphi = t[6].p; h.lognorm = t[7].p; h.expsim = t[8].p;
    h.rowptr = ti[0].p; h.col = ti[1].p;
  }
  CUDA_TRY(ctx->attrs.alloc(A));
  CUDA_TRY(cudaMemcpy(ctx->attrs.p, ctx->h_attrs.data(), sizeof(AttrDev) * A, cudaMemcpyHostToDevice));
  {
    int k = 0;
    for (int a = 0; a < A; ++a) if (ctx->h_attrs[a].is_const) ctx->perm[k++] = a;
    ctx->n_str = A - k;
    // byte-packed copy of the constant attributes in the tiles (k_link_pcg2): 1..4 of them, every vocabulary <= 255
    ctx->pack_consts = (k >= 1 && k <= 4) ? k : 0;
    for (int q = 0; q < k; ++q) if (ctx->h_attrs[ctx->perm[q]].V > 255) ctx->pack_consts = 0;
    if (getenv("DBL_NO_PACK")) ctx->pack_consts = 0;  // tests: the unpacked kernels on a packable model
    for (int a = 0; a < A; ++a) if (!ctx->h_attrs[a].is_const) ctx->perm[k++] = a;
    ctx->hslots = hash_ok ? Hmax : 0;
    ctx->hshift = 32;
    for (int h2 = 1; h2 < Hmax; h2 <<= 1) ctx->hshift -= 1;
    CUDA_TRY(ctx->perm_dev.alloc(A));
    CUDA_TRY(cudaMemcpy(ctx->perm_dev.p, ctx->perm, sizeof(int) * A, cudaMemcpyHostToDevice));
  }
  return upload_tree(ctx, d->tree);
}

extern "C" int dbl_ctx_create(dbl_ctx **out, const dbl_model_desc *d) {
  if (!out || !d || d->num_attrs <= 0 || d->num_attrs > DBL_MAX_ATTRS || d->num_files <= 0 || !d->indexes ||
      !d->alpha || !d->beta)
    return DBL_ERR_INVALID;
  for (int a = 0; a < d->num_attrs; ++a)
    if (!d->indexes[a] || !(d->alpha[a] > 0.0) || !(d->beta[a] > 0.0)) return DBL_ERR_INVALID;  // package.scala:165
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return DBL_ERR_CUDA;  // no CPU fallback
  auto *ctx = new dbl_ctx();
  ctx->A = d->num_attrs;
  ctx->F = d->num_files;
  ctx->seed = d->seed;
  ctx->rank = d->rank;
  ctx->world = d->world_size > 0 ? d->world_size : 1;
  ctx->alpha.assign(d->alpha, d->alpha + ctx->A);
  ctx->beta.assign(d->beta, d->beta + ctx->A);
  cudaGetDevice(&ctx->device);
  *out = ctx;
  CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&ctx->ev0));
  CUDA_TRY(cudaEventCreate(&ctx->ev1));
  int rc = upload_model(ctx, d);
  if (rc != DBL_OK) return rc;
  CUDA_TRY(ctx->counts.alloc(ctx->n_counts()));
  CUDA_TRY(ctx->loglik.alloc(1));
  CUDA_TRY(ctx->status.alloc(1));
  CUDA_TRY(ctx->pairs.alloc(1));
  CUDA_TRY(cudaMemset(ctx->status.p, 0, sizeof(int)));
  CUDA_TRY(cudaMemset(ctx->pairs.p, 0, sizeof(unsigned long long)));
  CUDA_TRY(ctx->theta.alloc((size_t)ctx->A * ctx->F));
  ctx->h_theta.assign((size_t)ctx->A * ctx->F, 0.0);
  ctx->h_counts.assign(ctx->n_counts(), 0);
  return DBL_OK;
}

extern "C" void dbl_ctx_destroy(dbl_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto &pe : ctx->pending_events) { cudaEventDestroy(pe.first); cudaEventDestroy(pe.second); }
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}
extern "C" const char *dbl_last_error(const dbl_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" int64_t dbl_num_records(const dbl_ctx *ctx) { return ctx ? ctx->R : 0; }
extern "C" int64_t dbl_num_entities(const dbl_ctx *ctx) { return ctx ? ctx->E : 0; }
extern "C" int64_t dbl_iteration(const dbl_ctx *ctx) { return ctx ? ctx->iteration : 0; }
extern "C" int64_t dbl_kernel_launches(const dbl_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" const char *dbl_version(void) { return "dblink_b200 0.1 (sm_100a)"; }

static int alloc_blocks(dbl_ctx *ctx) {
  const int P = ctx->P;
  if ((int)ctx->owner_h.size() != P) ctx->owner_h.assign(P, ctx->rank);
  CUDA_TRY(ctx->owner.alloc(P));
  CUDA_TRY(cudaMemcpy(ctx->owner.p, ctx->owner_h.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
  CUDA_TRY(ctx->ent_cnt.alloc(P + 1));
  CUDA_TRY(ctx->rec_cnt.alloc(P + 1));
  CUDA_TRY(ctx->ent_ptr.alloc(P + 1));
  CUDA_TRY(ctx->tile_ptr.alloc(P + 1));
  CUDA_TRY(ctx->rec_ptr.alloc(P + 1));
  CUDA_TRY(ctx->cta_ptr.alloc(P + 1));
  CUDA_TRY(ctx->cta_ptr2.alloc(P + 1));
  const size_t max_tiles = (size_t)(ctx->E / TE) + (size_t)P + 1;
  CUDA_TRY(ctx->tiles.alloc(max_tiles * tile_words(ctx->A)));
  ctx->max_ctas = (int)((ctx->R + LINK_WARPS - 1) / LINK_WARPS) + P;
  return DBL_OK;
}

static int alloc_state(dbl_ctx *ctx, int64_t R, int64_t E) {
  const int A = ctx->A;
  if (R <= 0 || E <= 0 || R > 0x7fffffff || E > 0x7fffffff) { ctx->set_error("bad R/E"); return DBL_ERR_INVALID; }
  if (ctx->R == R && ctx->E == E && ctx->x.p && ctx->tiles.p) {  // same shape as the previous state: reuse buffers
    CUDA_TRY(cudaMemsetAsync(ctx->ent_owned.p, 1, E, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(ctx->rec_owned.p, 1, R, ctx->stream));
    return DBL_OK;
  }
  ctx->R = R; ctx->E = E;
  CUDA_TRY(ctx->zbytes.alloc((size_t)R * A));
  CUDA_TRY(ctx->vflag.alloc(1));
  CUDA_TRY(ctx->file_cnt.alloc(ctx->F));
  CUDA_TRY(ctx->x.alloc((size_t)R * A));
  CUDA_TRY(ctx->file.alloc(R));
  CUDA_TRY(ctx->link.alloc(R));
  CUDA_TRY(ctx->newlink.alloc(R));
  CUDA_TRY(ctx->zmask.alloc(R));
  CUDA_TRY(ctx->y.alloc((size_t)E * A));
  CUDA_TRY(ctx->blk.alloc(E));
  CUDA_TRY(ctx->entN.alloc(E));
  CUDA_TRY(ctx->ent_owned.alloc(E));
  CUDA_TRY(ctx->rec_owned.alloc(R));
  CUDA_TRY(ctx->rec_class.alloc(R));
  CUDA_TRY(cudaMemsetAsync(ctx->ent_owned.p, 1, E, ctx->stream));
  CUDA_TRY(cudaMemsetAsync(ctx->rec_owned.p, 1, R, ctx->stream));
  CUDA_TRY(ctx->ent_dest.alloc(E));
  CUDA_TRY(ctx->ent_key.alloc(E));
  CUDA_TRY(ctx->link_key.alloc(R));
  CUDA_TRY(ctx->move_cnt.alloc(4 * (size_t)ctx->world));
  const int64_t M = std::max(R, E);
  CUDA_TRY(ctx->iota.alloc(M));
  k_iota<<<grid_for(M, 256), 256, 0, ctx->stream>>>(M, ctx->iota.p);
  CUDA_TRY(ctx->blk_sorted.alloc(E));
  CUDA_TRY(ctx->ent_sorted.alloc(E));
  CUDA_TRY(ctx->rec_key.alloc(R));
  CUDA_TRY(ctx->rec_key_sorted.alloc(R));
  CUDA_TRY(ctx->rec_sorted.alloc(R));
  { int rc = alloc_blocks(ctx); if (rc) return rc; }
  CUDA_TRY(ctx->link_sorted.alloc(R));
  CUDA_TRY(ctx->rec_by_ent.alloc(R));
  CUDA_TRY(ctx->ent_rec_cnt.alloc(E + 1));
  CUDA_TRY(ctx->ent_rec_ptr.alloc(E + 1));
  size_t b1 = 0, b2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b1, (const int *)nullptr, (int *)nullptr, (const int *)nullptr,
                                  (int *)nullptr, (int)M, 0, 32, ctx->stream);
  cub::DeviceScan::ExclusiveSum(nullptr, b2, (const int *)nullptr, (int *)nullptr, (int)(E + 1), ctx->stream);
  ctx->cub_bytes = std::max(b1, b2) + 256;
  CUDA_TRY(ctx->cub_tmp.alloc(ctx->cub_bytes));
  return DBL_OK;
}

static int bits_for(int64_t n) {
  int b = 1;
  while (((int64_t)1 << b) < n) ++b;
  return b;
}

// CSR entity -> linked records in ascending record id (LinksIndex, GU:84-119)
static int build_links_csr(dbl_ctx *ctx) {
  const int64_t R = ctx->R, E = ctx->E;
  size_t tb = ctx->cub_bytes;
  k_rec_link_keys<<<grid_for(R, 256), 256, 0, ctx->stream>>>(R, ctx->link.p, ctx->rec_owned.p, (int)E, ctx->link_key.p);
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tb, (const int *)ctx->link_key.p, ctx->link_sorted.p,
                                           (const int *)ctx->iota.p, ctx->rec_by_ent.p, (int)R, 0, bits_for(E + 1),
                                           ctx->stream));
  k_segment_ptr<<<grid_for(R + 1, 256), 256, 0, ctx->stream>>>(R, (int)E, ctx->link_sorted.p, ctx->ent_rec_ptr.p);
  ctx->launches += 3;
  return DBL_OK;
}

// group entities and records by block, build the tiled entity table (replaces the shuffle, GU:144)
static int relayout(dbl_ctx *ctx) {
  const int64_t R = ctx->R, E = ctx->E;
  const int A = ctx->A, P = ctx->P;
  const int pb = bits_for(P + 1);
  size_t tb = ctx->cub_bytes;
  k_ent_block_keys<<<grid_for(E, 256), 256, 0, ctx->stream>>>(E, ctx->blk.p, ctx->ent_owned.p, P, ctx->ent_key.p);
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tb, (const int *)ctx->ent_key.p, ctx->blk_sorted.p,
                                           (const int *)ctx->iota.p, ctx->ent_sorted.p, (int)E, 0, pb, ctx->stream));
  k_rec_block_keys<<<grid_for(R, 256), 256, 0, ctx->stream>>>(R, ctx->link.p, ctx->blk.p, ctx->rec_owned.p,
                                                              ctx->rec_class.p, P, ctx->rec_key.p);
  tb = ctx->cub_bytes;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tb, (const int *)ctx->rec_key.p, ctx->rec_key_sorted.p,
                                           (const int *)ctx->iota.p, ctx->rec_sorted.p, (int)R, 0,
                                           pb + REC_CLASS_BITS, ctx->stream));
  k_segment_ptr<<<grid_for(E + 1, 256), 256, 0, ctx->stream>>>(E, P, ctx->blk_sorted.p, ctx->ent_ptr.p);
  k_segment_ptr<<<grid_for(R + 1, 256), 256, 0, ctx->stream>>>(R, P, ctx->rec_key_sorted.p, ctx->rec_ptr.p,
                                                                REC_CLASS_BITS);
  k_block_scan<<<1, 32, 0, ctx->stream>>>(P, ctx->ent_ptr.p, ctx->rec_ptr.p, ctx->tile_ptr.p, ctx->cta_ptr.p,
                                          LINK_WARPS, ctx->cta_ptr2.p, MATCH_WARPS);
  CUDA_TRY(cudaMemsetAsync(ctx->tiles.p, 0, ctx->tiles.n * sizeof(int), ctx->stream));
  k_build_tiles<<<grid_for(E, 256), 256, 0, ctx->stream>>>(E, A, ctx->y.p, ctx->entN.p, ctx->blk_sorted.p,
                                                           ctx->ent_sorted.p, ctx->ent_ptr.p, ctx->tile_ptr.p,
                                                           ctx->tiles.p, ctx->perm_dev.p, P, ctx->pack_consts);
  ctx->launches += 11;
  ctx->inv_valid = false;
  ctx->h_owned_ent = -1;  // ent_ptr[P] changed; fetch_summary (or the index build) reads it back
  CUDA_TRY(cudaGetLastError());
  return DBL_OK;
}

// entity N / block ids / entity+record summary of the current state
static int refresh_summary(dbl_ctx *ctx, bool draw_z, int sampler) {
  const int A = ctx->A, F = ctx->F;
  (void)sampler;
  CUDA_TRY(cudaMemsetAsync(ctx->counts.p, 0, sizeof(long long) * ctx->n_counts(), ctx->stream));
  CUDA_TRY(cudaMemsetAsync(ctx->loglik.p, 0, sizeof(double), ctx->stream));
  k_entity_post<<<grid_for(ctx->E, 256), 256, 0, ctx->stream>>>(ctx->E, A, ctx->y.p, ctx->attrs.p, ctx->tree,
                                                               ctx->entN.p, ctx->blk.p, ctx->ent_rec_ptr.p,
                                                               ctx->counts.p, ctx->iso_slot(), ctx->loglik.p, ctx->ent_owned.p);
  DistParams dp;
  dp.A = A; dp.F = F; dp.draw = draw_z ? 1 : 0; dp.seed = ctx->seed; dp.iter = (uint32_t)(ctx->iteration + 1);
  dp.R = ctx->R; dp.attrs = ctx->attrs.p; dp.x = ctx->x.p; dp.file = ctx->file.p; dp.link = ctx->link.p;
  dp.y = ctx->y.p; dp.zmask = ctx->zmask.p; dp.theta = ctx->theta.p; dp.counts = ctx->counts.p;
  dp.loglik = ctx->loglik.p;
  dp.rec_owned = ctx->rec_owned.p;
  k_dist<<<grid_for(ctx->R, 256), 256, 0, ctx->stream>>>(dp);
  ctx->launches += 4;
  CUDA_TRY(cudaGetLastError());
  return DBL_OK;
}

static int fetch_summary(dbl_ctx *ctx) {
  CUDA_TRY(cudaMemcpyAsync(ctx->h_counts.data(), ctx->counts.p, sizeof(long long) * ctx->n_counts(),
                           cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(&ctx->h_loglik_part, ctx->loglik.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  int st = 0;
  CUDA_TRY(cudaMemcpyAsync(&st, ctx->status.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  unsigned long long pr = 0;
  CUDA_TRY(cudaMemcpyAsync(&pr, ctx->pairs.p, sizeof(pr), cudaMemcpyDeviceToHost, ctx->stream));
  int owned = -1;
  if (ctx->ent_ptr.p) CUDA_TRY(cudaMemcpyAsync(&owned, ctx->ent_ptr.p + ctx->P, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->h_pairs = (int64_t)pr;
  ctx->h_owned_ent = owned;
  if (st) {
    ctx->set_error("zero probability mass in a link draw");
    cudaMemsetAsync(ctx->status.p, 0, sizeof(int), ctx->stream);
    return DBL_ERR_ZERO_MASS;
  }
  return DBL_OK;
}

static int finish_new_state(dbl_ctx *ctx, bool check_state) {
  // range checks on the device, then file sizes (RecordsCache.fileSizes)
  CUDA_TRY(cudaMemsetAsync(ctx->vflag.p, 0, sizeof(int), ctx->stream));
  k_validate<<<grid_for(std::max(ctx->R, ctx->E), 256), 256, 0, ctx->stream>>>(
      ctx->R, ctx->E, ctx->A, ctx->F, ctx->attrs.p, ctx->x.p, ctx->file.p, check_state ? ctx->link.p : nullptr,
      check_state ? ctx->y.p : nullptr, ctx->vflag.p);
  int bad = 0;
  CUDA_TRY(cudaMemcpyAsync(&bad, ctx->vflag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (bad) {
    ctx->has_state = false;
    ctx->set_error(bad & 1 ? "record value id out of range" : bad & 2 ? "file id out of range"
                   : bad & 4 ? "link out of range" : "entity value id out of range");
    return DBL_ERR_INVALID;
  }
  CUDA_TRY(cudaMemsetAsync(ctx->file_cnt.p, 0, sizeof(int) * ctx->F, ctx->stream));
  k_hist<<<grid_for(ctx->R, 256), 256, 0, ctx->stream>>>(ctx->R, ctx->file.p, ctx->file_cnt.p);
  std::vector<int> hc(ctx->F);
  CUDA_TRY(cudaMemcpyAsync(hc.data(), ctx->file_cnt.p, sizeof(int) * ctx->F, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->file_sizes.assign(hc.begin(), hc.end());
  ctx->launches += 2;
  int rc = build_links_csr(ctx);
  if (rc) return rc;
  rc = refresh_summary(ctx, false, 0);
  if (rc) return rc;
  rc = relayout(ctx);
  if (rc) return rc;
  rc = fetch_summary(ctx);
  if (rc) return rc;
  ctx->has_state = true;
  return DBL_OK;
}

extern "C" int dbl_set_partitioner(dbl_ctx *ctx, const dbl_kdtree *tree) {
  if (!ctx) return DBL_ERR_INVALID;
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  int rc = upload_tree(ctx, tree);
  if (rc) return rc;
  if (!ctx->has_state) return DBL_OK;
  rc = alloc_blocks(ctx);
  if (rc) return rc;
  rc = refresh_summary(ctx, false, 0);  // recomputes block ids (and the unchanged summary)
  if (rc) return rc;
  rc = relayout(ctx);
  if (rc) return rc;
  return fetch_summary(ctx);
}
extern "C" int32_t dbl_num_partitions(const dbl_ctx *ctx) { return ctx ? ctx->P : 0; }

extern "C" int dbl_state_init(dbl_ctx *ctx, int64_t R, const int32_t *x, const int32_t *file, int64_t pop) {
  if (!ctx || !x || !file) return DBL_ERR_INVALID;
  CUDA_TRY(cudaSetDevice(ctx->device));
  const int64_t E = pop > 0 ? pop : R;
  int rc = alloc_state(ctx, R, E);
  if (rc) return rc;
  const int A = ctx->A;
  CUDA_TRY(cudaMemcpyAsync(ctx->x.p, x, sizeof(int) * R * A, cudaMemcpyHostToDevice, ctx->stream));
  k_rec_class<<<grid_for(R, 256), 256, 0, ctx->stream>>>(R, A, ctx->attrs.p, ctx->x.p, ctx->rec_class.p);
  CUDA_TRY(cudaMemcpyAsync(ctx->file.p, file, sizeof(int) * R, cudaMemcpyHostToDevice, ctx->stream));
  k_init_entities<<<grid_for(E * A, 256), 256, 0, ctx->stream>>>(E, R, A, ctx->seed, ctx->attrs.p, ctx->x.p, ctx->y.p);
  k_init_records<<<grid_for(R, 256), 256, 0, ctx->stream>>>(E, R, A, ctx->x.p, ctx->y.p, ctx->link.p, ctx->zmask.p);
  ctx->launches += 3;
  for (int a = 0; a < A; ++a)
    for (int f = 0; f < ctx->F; ++f)
      ctx->h_theta[a * ctx->F + f] = ctx->alpha[a] / (ctx->alpha[a] + ctx->beta[a]);  // DistortionProbs.scala:38-40
  CUDA_TRY(cudaMemcpyAsync(ctx->theta.p, ctx->h_theta.data(), sizeof(double) * A * ctx->F, cudaMemcpyHostToDevice,
                           ctx->stream));
  ctx->iteration = 0;
  return finish_new_state(ctx, false);
}

extern "C" int dbl_state_upload(dbl_ctx *ctx, int64_t R, int64_t E, const int32_t *x, const int32_t *file,
                                const uint8_t *z, const int32_t *link, const int32_t *y, const double *theta,
                                int64_t iteration) {
  if (!ctx || !x || !file || !z || !link || !y || !theta) return DBL_ERR_INVALID;
  CUDA_TRY(cudaSetDevice(ctx->device));
  int rc = alloc_state(ctx, R, E);
  if (rc) return rc;
  const int A = ctx->A;
  DevBuf<uint8_t> &zb = ctx->zbytes;
  CUDA_TRY(cudaMemcpyAsync(ctx->x.p, x, sizeof(int) * R * A, cudaMemcpyDefault, ctx->stream));
  k_rec_class<<<grid_for(R, 256), 256, 0, ctx->stream>>>(R, A, ctx->attrs.p, ctx->x.p, ctx->rec_class.p);
  CUDA_TRY(cudaMemcpyAsync(ctx->file.p, file, sizeof(int) * R, cudaMemcpyDefault, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(zb.p, z, (size_t)R * A, cudaMemcpyDefault, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->link.p, link, sizeof(int) * R, cudaMemcpyDefault, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->y.p, y, sizeof(int) * E * A, cudaMemcpyDefault, ctx->stream));
  k_pack_z<<<grid_for(R, 256), 256, 0, ctx->stream>>>(R, A, zb.p, ctx->zmask.p);
  ctx->launches += 2;
  std::copy(theta, theta + (size_t)A * ctx->F, ctx->h_theta.begin());
  CUDA_TRY(cudaMemcpyAsync(ctx->theta.p, ctx->h_theta.data(), sizeof(double) * A * ctx->F, cudaMemcpyHostToDevice,
                           ctx->stream));
  ctx->iteration = iteration;
  rc = finish_new_state(ctx, true);
  return rc;
}

extern "C" int dbl_state_download(dbl_ctx *ctx, uint8_t *z, int32_t *link, int32_t *y, double *theta,
                                  int32_t *block_of_entity) {
  if (!ctx) return DBL_ERR_INVALID;
  if (!ctx->has_state) { ctx->set_error("no state"); return DBL_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  const int A = ctx->A;
  DevBuf<uint8_t> &zb = ctx->zbytes;
  if (z) {
    k_unpack_z<<<grid_for(ctx->R, 256), 256, 0, ctx->stream>>>(ctx->R, A, ctx->zmask.p, zb.p);
    ctx->launches += 1;
    CUDA_TRY(cudaMemcpyAsync(z, zb.p, (size_t)ctx->R * A, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (link) CUDA_TRY(cudaMemcpyAsync(link, ctx->link.p, sizeof(int) * ctx->R, cudaMemcpyDeviceToHost, ctx->stream));
  if (y) CUDA_TRY(cudaMemcpyAsync(y, ctx->y.p, sizeof(int) * ctx->E * A, cudaMemcpyDeviceToHost, ctx->stream));
  if (block_of_entity)
    CUDA_TRY(cudaMemcpyAsync(block_of_entity, ctx->blk.p, sizeof(int) * ctx->E, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (theta) std::copy(ctx->h_theta.begin(), ctx->h_theta.end(), theta);
  return DBL_OK;
}

extern "C" int dbl_links_download(dbl_ctx *ctx, int32_t *link_out, int32_t *block_out) {
  return dbl_state_download(ctx, nullptr, link_out, nullptr, nullptr, block_out);
}

static void drain_link_events(dbl_ctx *ctx) {
  for (auto &pe : ctx->pending_events) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pe.first, pe.second) == cudaSuccess) {
      ctx->link_ms += ms;
      ctx->link_launches += 1;
    }
    cudaEventDestroy(pe.first);
    cudaEventDestroy(pe.second);
  }
  ctx->pending_events.clear();
}

// ---------------------------------------------------------------------------------------------------
// link kernel dispatch
// ---------------------------------------------------------------------------------------------------
#define DBL_DECL(N) int dbl_launch_pcg2_a##N(int ns, int grid, cudaStream_t stream, const LinkParams &lp);
DBL_DECL(1) DBL_DECL(2) DBL_DECL(3) DBL_DECL(4) DBL_DECL(5) DBL_DECL(6) DBL_DECL(7) DBL_DECL(8)
DBL_DECL(9) DBL_DECL(10) DBL_DECL(11) DBL_DECL(12) DBL_DECL(13) DBL_DECL(14) DBL_DECL(15) DBL_DECL(16)
#undef DBL_DECL

// (block, attribute, value) -> candidate positions, for k_link_pruned
static int ensure_inverted_index(dbl_ctx *ctx) {
  if (ctx->inv_valid) return DBL_OK;
  const int64_t cap = ctx->E * ctx->A;
  if (cap > 0x7fffffff) { ctx->set_error("inverted index too large"); return DBL_ERR_INVALID; }
  if (ctx->h_owned_ent < 0) {  // relayout without a summary fetch since (block-level sweeps)
    CUDA_TRY(cudaMemcpyAsync(&ctx->h_owned_ent, ctx->ent_ptr.p + ctx->P, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  // only the entities of owned blocks are indexed: they come first in ent_sorted
  const int64_t n = (int64_t)ctx->h_owned_ent * ctx->A;
  if (ctx->inv_key.n != (size_t)cap) {
    CUDA_TRY(ctx->inv_key_in.alloc(cap));
    CUDA_TRY(ctx->inv_key.alloc(cap));
    CUDA_TRY(ctx->inv_pos_in.alloc(cap));
    CUDA_TRY(ctx->inv_pos.alloc(cap));
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const int *)nullptr, (int *)nullptr, (int)cap, 0, 64, ctx->stream);
    ctx->inv_tmp_bytes = tb + 256;
    CUDA_TRY(ctx->inv_tmp.alloc(ctx->inv_tmp_bytes));
  }
  int vmax = 1;
  for (int a = 0; a < ctx->A; ++a) vmax = std::max(vmax, ctx->h_attrs[a].V);
  ctx->inv_vbits = bits_for(vmax + 1);
  const int nbits = ctx->inv_vbits + bits_for((int64_t)(ctx->P + 1) * ctx->A);
  if (n > 0) {
    k_inv_keys<<<grid_for(n, 256), 256, 0, ctx->stream>>>((int64_t)ctx->h_owned_ent, ctx->A, ctx->P, ctx->inv_vbits,
                                                          ctx->y.p, ctx->blk_sorted.p, ctx->ent_sorted.p,
                                                          ctx->ent_ptr.p, ctx->perm_dev.p, ctx->inv_key_in.p,
                                                          ctx->inv_pos_in.p);
    size_t tb = ctx->inv_tmp_bytes;
    // stable radix sort on the significant bits only: positions stay ascending inside a key
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(ctx->inv_tmp.p, tb, (const unsigned long long *)ctx->inv_key_in.p,
                                             ctx->inv_key.p, (const int *)ctx->inv_pos_in.p, ctx->inv_pos.p, (int)n,
                                             0, std::min(64, nbits), ctx->stream));
  }
  // dense (block, attribute, value) -> posting pointers when the table is small enough (P * sum of vocabulary
  // sizes entries); otherwise (block, attribute) segment pointers + a binary search per record
  InvDense &dn = ctx->inv_dense;
  dn.A = ctx->A; dn.vbits = ctx->inv_vbits; dn.sumV = 0;
  for (int k = 0; k < ctx->A; ++k) { dn.voff[k] = dn.sumV; dn.sumV += ctx->h_attrs[ctx->perm[k]].V; }
  const long long n_ids = (long long)ctx->P * dn.sumV;
  long long dense_max = 1ll << 25;  // entries; DBL_INV_DENSE_MAX overrides (tests force the binary-search path with 0)
  if (const char *ev = getenv("DBL_INV_DENSE_MAX")) dense_max = atoll(ev);
  ctx->inv_use_dense = n_ids <= dense_max;
  if (ctx->inv_use_dense) {
    if (ctx->inv_vptr.n != (size_t)n_ids + 1) CUDA_TRY(ctx->inv_vptr.alloc((size_t)n_ids + 1));
    k_inv_value_ptr<<<grid_for(n + 1, 256), 256, 0, ctx->stream>>>(n, n_ids, dn, ctx->inv_key.p, ctx->inv_vptr.p);
  } else {
    const int n_groups = (ctx->P + 1) * ctx->A;
    if (ctx->inv_seg.n != (size_t)n_groups + 1) CUDA_TRY(ctx->inv_seg.alloc((size_t)n_groups + 1));
    k_inv_segments<<<grid_for(n + 1, 256), 256, 0, ctx->stream>>>(n, n_groups, ctx->inv_vbits, ctx->inv_key.p,
                                                                  ctx->inv_seg.p);
  }
  ctx->launches += 5;
  ctx->inv_valid = true;
  CUDA_TRY(cudaGetLastError());
  return DBL_OK;
}

static int launch_link(dbl_ctx *ctx, int sampler, uint32_t it) {
  const int A = ctx->A;
  LinkParams lp;
  memset(&lp, 0, sizeof(lp));
  lp.A = A; lp.F = ctx->F; lp.P = ctx->P; lp.sampler = sampler; lp.seed = ctx->seed; lp.iter = it;
  lp.attrs = ctx->attrs.p; lp.x = ctx->x.p; lp.file = ctx->file.p; lp.link = ctx->link.p; lp.zmask = ctx->zmask.p;
  lp.theta = ctx->theta.p; lp.ent_ptr = ctx->ent_ptr.p; lp.tile_ptr = ctx->tile_ptr.p; lp.rec_ptr = ctx->rec_ptr.p;
  lp.cta_ptr = ctx->cta_ptr.p; lp.ent_sorted = ctx->ent_sorted.p; lp.rec_sorted = ctx->rec_sorted.p;
  lp.tiles = ctx->tiles.p; lp.newlink = ctx->newlink.p; lp.status = ctx->status.p; lp.pairs = ctx->pairs.p;
  for (int k = 0; k < A; ++k) lp.perm[k] = ctx->perm[k];
  lp.blk_of_link = ctx->blk.p;
  lp.pack_consts = ctx->pack_consts;
  const size_t ring = (size_t)LINK_STAGES * tile_words(A) * 4 + 128;
  const int mode = ctx->link_mode;  // 0 auto, 1 force generic
  lp.hslots = ctx->hslots; lp.hshift = ctx->hshift;
  if (mode != 1 && sampler == DBL_PCG_II && ctx->hslots > 0 && A <= LINK_MAX_UNROLL_A &&
      pcg2_smem_bytes(A, ctx->n_str, ctx->hslots) <= 100 * 1024) {
    int rc = -1;
    switch (A) {
#define DBL_CASE(N) case N: rc = dbl_launch_pcg2_a##N(ctx->n_str, ctx->max_ctas, ctx->stream, lp); break;
      DBL_CASE(1) DBL_CASE(2) DBL_CASE(3) DBL_CASE(4) DBL_CASE(5) DBL_CASE(6) DBL_CASE(7) DBL_CASE(8)
      DBL_CASE(9) DBL_CASE(10) DBL_CASE(11) DBL_CASE(12) DBL_CASE(13) DBL_CASE(14) DBL_CASE(15) DBL_CASE(16)
#undef DBL_CASE
    }
    if (rc != 0) { ctx->set_error(std::string("k_link_pcg2 launch: ") + cudaGetErrorString((cudaError_t)rc)); return DBL_ERR_CUDA; }
    return DBL_OK;
  }
  if (mode == 0 && sampler != DBL_PCG_II) {  // pruned scoring through the inverted index
    int rc = ensure_inverted_index(ctx);
    if (rc) return rc;
    PrunedParams pp;
    pp.lp = lp;
    pp.inv_key = ctx->inv_key.p;
    pp.inv_pos = ctx->inv_pos.p;
    pp.inv_n = (long long)ctx->h_owned_ent * ctx->A;
    pp.R = ctx->R;
    pp.vbits = ctx->inv_vbits;
    pp.inv_seg = ctx->inv_seg.p;
    pp.rec_key_sorted = ctx->rec_key_sorted.p;
    pp.rec_key_shift = REC_CLASS_BITS;
    pp.inv_vptr = ctx->inv_use_dense ? ctx->inv_vptr.p : nullptr;
    pp.sumV = ctx->inv_dense.sumV;
    for (int k = 0; k < A; ++k) pp.voff[k] = ctx->inv_dense.voff[k];
    k_link_pruned<<<grid_for(ctx->R, LINK_WARPS), LINK_WARPS * 32, 0, ctx->stream>>>(pp);
    return DBL_OK;
  }
  if (mode != 1 && sampler != DBL_PCG_II && ring <= 160 * 1024) {
    static size_t configured = 0;
    if (configured < ring) {
      CUDA_TRY(cudaFuncSetAttribute(k_link_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring));
      configured = ring;
    }
    lp.cta_ptr = ctx->cta_ptr2.p;  // MATCH_WARPS records per CTA
    const int grid = (int)((ctx->R + MATCH_WARPS - 1) / MATCH_WARPS) + ctx->P;
    k_link_match<<<grid, (MATCH_WARPS + 1) * 32, ring, ctx->stream>>>(lp);
    return DBL_OK;
  }
  k_link_generic<<<ctx->max_ctas, LINK_WARPS * 32, 0, ctx->stream>>>(lp);
  return DBL_OK;
}

// theta | summary -> links -> entity values -> N/blocks/distortions/summary of the shard owned by this rank
// (1) theta | summary of the previous state (State.scala:83, GU:305-320) -- A*F scalars on the host
static int draw_theta(dbl_ctx *ctx) {
  const int A = ctx->A, F = ctx->F;
  const uint32_t it = (uint32_t)(ctx->iteration + 1);
  std::vector<int64_t> agg((size_t)A * F);
  for (int i = 0; i < A * F; ++i) agg[i] = ctx->h_counts[i];
  host_draw_theta(A, F, ctx->alpha.data(), ctx->beta.data(), ctx->seed, agg.data(), ctx->file_sizes.data(), it,
                  ctx->h_theta.data());
  CUDA_TRY(cudaMemcpyAsync(ctx->theta.p, ctx->h_theta.data(), sizeof(double) * A * F, cudaMemcpyHostToDevice,
                           ctx->stream));
  return DBL_OK;
}

// (2)-(4) for the owned entities / records (all of them on an unsharded context): updatePartition, GU:156-211
static int update_owned(dbl_ctx *ctx, int sampler, bool masked) {
  const int A = ctx->A, F = ctx->F;
  const uint32_t it = (uint32_t)(ctx->iteration + 1);
  // (2) links
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  CUDA_TRY(cudaEventRecord(e0, ctx->stream));
  {
    int rc = launch_link(ctx, sampler, it);
    if (rc) return rc;
  }
  CUDA_TRY(cudaEventRecord(e1, ctx->stream));
  ctx->pending_events.emplace_back(e0, e1);
  ctx->launches += 1;
  CUDA_TRY(cudaGetLastError());
  if (masked) {  // the link kernel only writes records of owned blocks: keep the others as they were
    k_merge_links<<<grid_for(ctx->R, 256), 256, 0, ctx->stream>>>(ctx->R, ctx->rec_owned.p, ctx->link.p, ctx->newlink.p);
    ctx->launches += 1;
  }
  std::swap(ctx->link.p, ctx->newlink.p);
  // (3) entity values
  int rc = build_links_csr(ctx);
  if (rc) return rc;
  ValParams vp;
  vp.A = A; vp.F = F; vp.sampler = sampler; vp.seed = ctx->seed; vp.iter = it; vp.E = ctx->E;
  vp.attrs = ctx->attrs.p; vp.x = ctx->x.p; vp.file = ctx->file.p; vp.zmask = ctx->zmask.p; vp.theta = ctx->theta.p;
  vp.ent_rec_ptr = ctx->ent_rec_ptr.p; vp.rec_by_ent = ctx->rec_by_ent.p; vp.y = ctx->y.p;
  vp.ent_owned = ctx->ent_owned.p;
  k_values<<<grid_for(ctx->E * A, 128), 128, 0, ctx->stream>>>(vp);
  ctx->launches += 1;
  // (4) N(e), new block ids, distortions, summary
  return refresh_summary(ctx, true, sampler);
}

static int sweep_local(dbl_ctx *ctx, int sampler) {
  int rc = draw_theta(ctx);
  if (rc) return rc;
  return update_owned(ctx, sampler, ctx->world > 1);
}

// (5) re-partition + summary fetch (also the sync point that bounds the sweep)
static int sweep_finish(dbl_ctx *ctx) {
  int rc = relayout(ctx);
  if (rc) return rc;
  ctx->iteration += 1;
  return fetch_summary(ctx);
}

extern "C" int dbl_sweep(dbl_ctx *ctx, int sampler, int32_t n_sweeps) {
  if (!ctx) return DBL_ERR_INVALID;
  if (sampler < 0 || sampler > 3 || n_sweeps < 0) { ctx->set_error("bad sampler / n_sweeps"); return DBL_ERR_INVALID; }
  if (!ctx->has_state) { ctx->set_error("dbl_sweep before dbl_state_init/upload"); return DBL_ERR_STATE; }
  if (ctx->world > 1) { ctx->set_error("dbl_sweep on a sharded context: use dbl_sweep_begin/exchange/end"); return DBL_ERR_STATE; }
  if (ctx->in_sweep) { ctx->set_error("dbl_sweep inside an open sweep"); return DBL_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
  for (int s = 0; s < n_sweeps; ++s) {
    int rc = sweep_local(ctx, sampler);
    if (rc) return rc;
    rc = sweep_finish(ctx);
    if (rc) return rc;
  }
  CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
  CUDA_TRY(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->last_sweep_ms = ms;
  drain_link_events(ctx);
  return DBL_OK;
}

// ---------------------------------------------------------------------------------------------------
// block-level entry points: one call per k-d-tree block, mirroring GibbsUpdates.updatePartition (GU:156-211)
// ---------------------------------------------------------------------------------------------------
extern "C" int dbl_block_sweep_begin(dbl_ctx *ctx, int sampler) {
  if (!ctx) return DBL_ERR_INVALID;
  if (sampler < 0 || sampler > 3) { ctx->set_error("bad sampler"); return DBL_ERR_INVALID; }
  if (!ctx->has_state) { ctx->set_error("no state"); return DBL_ERR_STATE; }
  if (ctx->world > 1) { ctx->set_error("block-level sweeps need an unsharded context"); return DBL_ERR_STATE; }
  if (ctx->in_sweep) { ctx->set_error("a sweep is already open"); return DBL_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = draw_theta(ctx);
  if (rc) return rc;
  // block membership is fixed for the whole sweep (the shuffle happens after every partition was updated, GU:144)
  CUDA_TRY(ctx->blk_frozen.alloc(ctx->E));
  CUDA_TRY(cudaMemcpyAsync(ctx->blk_frozen.p, ctx->blk.p, sizeof(int) * ctx->E, cudaMemcpyDeviceToDevice, ctx->stream));
  ctx->block_done.assign(ctx->P, 0);
  ctx->block_sampler = sampler;
  ctx->in_sweep = true;
  ctx->in_block_sweep = true;
  return DBL_OK;
}

extern "C" int dbl_update_block(dbl_ctx *ctx, int32_t block_id) {
  if (!ctx) return DBL_ERR_INVALID;
  if (!ctx->in_block_sweep) { ctx->set_error("dbl_update_block outside dbl_block_sweep_begin/end"); return DBL_ERR_STATE; }
  if (block_id < 0 || block_id >= ctx->P) { ctx->set_error("block id out of range"); return DBL_ERR_INVALID; }
  if (ctx->block_done[block_id]) { ctx->set_error("block already updated in this sweep"); return DBL_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_mark_ent_block<<<grid_for(ctx->E, 256), 256, 0, ctx->stream>>>(ctx->E, ctx->blk_frozen.p, block_id, ctx->ent_owned.p);
  k_mark_rec_owned<<<grid_for(ctx->R, 256), 256, 0, ctx->stream>>>(ctx->R, ctx->link.p, ctx->ent_owned.p,
                                                                  ctx->rec_owned.p);
  ctx->launches += 2;
  int rc = relayout(ctx);  // tiles of this block only (everything else sorts into the dummy block)
  if (rc) return rc;
  rc = update_owned(ctx, ctx->block_sampler, true);
  if (rc) return rc;
  ctx->block_done[block_id] = 1;
  return DBL_OK;
}

extern "C" int dbl_block_sweep_end(dbl_ctx *ctx) {
  if (!ctx) return DBL_ERR_INVALID;
  if (!ctx->in_block_sweep) { ctx->set_error("dbl_block_sweep_end without dbl_block_sweep_begin"); return DBL_ERR_STATE; }
  for (int b = 0; b < ctx->P; ++b)
    if (!ctx->block_done[b]) { ctx->set_error("dbl_block_sweep_end: not every block was updated"); return DBL_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  ctx->in_sweep = ctx->in_block_sweep = false;
  CUDA_TRY(cudaMemsetAsync(ctx->ent_owned.p, 1, ctx->E, ctx->stream));
  CUDA_TRY(cudaMemsetAsync(ctx->rec_owned.p, 1, ctx->R, ctx->stream));
  int rc = build_links_csr(ctx);
  if (rc) return rc;
  rc = refresh_summary(ctx, false, 0);  // summary of the whole state (updateSummaryVariables, GU:219-301)
  if (rc) return rc;
  rc = sweep_finish(ctx);               // the shuffle: regroup by the new block ids
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
  CUDA_TRY(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->last_sweep_ms = ms;
  drain_link_events(ctx);
  return DBL_OK;
}

// ---------------------------------------------------------------------------------------------------
// multi-GPU entry points: blocks are sharded over ranks; one exchange of moved clusters per sweep
// ---------------------------------------------------------------------------------------------------
static int apply_ownership(dbl_ctx *ctx) {
  k_mark_ent_owned<<<grid_for(ctx->E, 256), 256, 0, ctx->stream>>>(ctx->E, ctx->blk.p, ctx->owner.p, ctx->rank,
                                                                  ctx->ent_owned.p);
  k_mark_rec_owned<<<grid_for(ctx->R, 256), 256, 0, ctx->stream>>>(ctx->R, ctx->link.p, ctx->ent_owned.p,
                                                                  ctx->rec_owned.p);
  ctx->launches += 2;
  CUDA_TRY(cudaGetLastError());
  return DBL_OK;
}

extern "C" int dbl_set_block_owners(dbl_ctx *ctx, const int32_t *owner_of_block) {
  if (!ctx || !owner_of_block) return DBL_ERR_INVALID;
  if (!ctx->has_state) { ctx->set_error("set_block_owners needs a (replicated) state"); return DBL_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  for (int b = 0; b < ctx->P; ++b)
    if (owner_of_block[b] < 0 || owner_of_block[b] >= ctx->world) { ctx->set_error("owner out of range"); return DBL_ERR_INVALID; }
  ctx->owner_h.assign(owner_of_block, owner_of_block + ctx->P);
  CUDA_TRY(cudaMemcpyAsync(ctx->owner.p, ctx->owner_h.data(), sizeof(int) * ctx->P, cudaMemcpyHostToDevice, ctx->stream));
  int rc = apply_ownership(ctx);
  if (rc) return rc;
  rc = build_links_csr(ctx);
  if (rc) return rc;
  rc = refresh_summary(ctx, false, 0);
  if (rc) return rc;
  rc = relayout(ctx);
  if (rc) return rc;
  return fetch_summary(ctx);
}

extern "C" int dbl_sweep_begin(dbl_ctx *ctx, int sampler, int64_t *ent_counts, int64_t *rec_counts) {
  if (!ctx || !ent_counts || !rec_counts) return DBL_ERR_INVALID;
  if (sampler < 0 || sampler > 3) { ctx->set_error("bad sampler"); return DBL_ERR_INVALID; }
  if (!ctx->has_state) { ctx->set_error("no state"); return DBL_ERR_STATE; }
  if (ctx->in_sweep) { ctx->set_error("dbl_sweep_begin twice"); return DBL_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = sweep_local(ctx, sampler);
  if (rc) return rc;
  const int W = ctx->world;
  CUDA_TRY(cudaMemsetAsync(ctx->move_cnt.p, 0, sizeof(unsigned long long) * 4 * W, ctx->stream));
  k_move_count_ent<<<grid_for(ctx->E, 256), 256, 0, ctx->stream>>>(ctx->E, ctx->blk.p, ctx->owner.p, ctx->rank,
                                                                  ctx->ent_owned.p, ctx->ent_dest.p, ctx->move_cnt.p);
  k_move_count_rec<<<grid_for(ctx->R, 256), 256, 0, ctx->stream>>>(ctx->R, ctx->link.p, ctx->rec_owned.p,
                                                                  ctx->ent_dest.p, ctx->move_cnt.p + W);
  ctx->launches += 2;
  std::vector<unsigned long long> h(2 * W);
  CUDA_TRY(cudaMemcpyAsync(h.data(), ctx->move_cnt.p, sizeof(unsigned long long) * 2 * W, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->h_move_ent.assign(W, 0);
  ctx->h_move_rec.assign(W, 0);
  for (int d = 0; d < W; ++d) {
    ent_counts[d] = ctx->h_move_ent[d] = (int64_t)h[d];
    rec_counts[d] = ctx->h_move_rec[d] = (int64_t)h[W + d];
  }
  ctx->in_sweep = true;
  return DBL_OK;
}

extern "C" int dbl_exchange_pack(dbl_ctx *ctx, void *ent_buf_dev, void *rec_buf_dev) {
  if (!ctx) return DBL_ERR_INVALID;
  if (!ctx->in_sweep) { ctx->set_error("dbl_exchange_pack outside a sweep"); return DBL_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  const int W = ctx->world;
  std::vector<unsigned long long> cur(2 * W);
  unsigned long long oe = 0, orc = 0;
  for (int d = 0; d < W; ++d) { cur[d] = oe; oe += ctx->h_move_ent[d]; cur[W + d] = orc; orc += ctx->h_move_rec[d]; }
  CUDA_TRY(cudaMemcpyAsync(ctx->move_cnt.p + 2 * W, cur.data(), sizeof(unsigned long long) * 2 * W, cudaMemcpyHostToDevice,
                           ctx->stream));
  if (orc > 0) {
    if (!rec_buf_dev) return DBL_ERR_INVALID;
    k_move_pack_rec<<<grid_for(ctx->R, 256), 256, 0, ctx->stream>>>(ctx->R, ctx->link.p, ctx->zmask.p, ctx->ent_dest.p,
                                                                   ctx->rec_owned.p, ctx->move_cnt.p + 3 * W,
                                                                   (int *)rec_buf_dev);
  }
  if (oe > 0) {
    if (!ent_buf_dev) return DBL_ERR_INVALID;
    k_move_pack_ent<<<grid_for(ctx->E, 256), 256, 0, ctx->stream>>>(ctx->E, ctx->A, ctx->y.p, ctx->ent_dest.p,
                                                                   ctx->ent_owned.p, ctx->move_cnt.p + 2 * W,
                                                                   (int *)ent_buf_dev);
  }
  ctx->launches += 2;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // the host moves the buffers with NCCL on its own stream
  return DBL_OK;
}

extern "C" int dbl_exchange_unpack(dbl_ctx *ctx, const void *ent_buf_dev, int64_t n_ent, const void *rec_buf_dev,
                                   int64_t n_rec) {
  if (!ctx || n_ent < 0 || n_rec < 0) return DBL_ERR_INVALID;
  if (!ctx->in_sweep) { ctx->set_error("dbl_exchange_unpack outside a sweep"); return DBL_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (n_ent > 0)
    k_unpack_ent<<<grid_for(n_ent, 256), 256, 0, ctx->stream>>>(n_ent, ctx->A, (const int *)ent_buf_dev, ctx->attrs.p,
                                                                ctx->tree, ctx->y.p, ctx->entN.p, ctx->blk.p,
                                                                ctx->ent_owned.p);
  if (n_rec > 0)
    k_unpack_rec<<<grid_for(n_rec, 256), 256, 0, ctx->stream>>>(n_rec, (const int *)rec_buf_dev, ctx->link.p,
                                                                ctx->zmask.p, ctx->rec_owned.p);
  ctx->launches += 2;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return DBL_OK;
}

extern "C" int dbl_sweep_end(dbl_ctx *ctx) {
  if (!ctx) return DBL_ERR_INVALID;
  if (!ctx->in_sweep) { ctx->set_error("dbl_sweep_end without dbl_sweep_begin"); return DBL_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  ctx->in_sweep = false;
  int rc = sweep_finish(ctx);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
  CUDA_TRY(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->last_sweep_ms = ms;
  drain_link_events(ctx);
  return DBL_OK;
}

extern "C" int dbl_partial_summary(dbl_ctx *ctx, int64_t *counts, double *loglik) {
  if (!ctx || !counts || !loglik) return DBL_ERR_INVALID;
  for (int i = 0; i < ctx->n_counts(); ++i) counts[i] = ctx->h_counts[i];
  *loglik = ctx->h_loglik_part;
  return DBL_OK;
}
extern "C" int dbl_set_global_summary(dbl_ctx *ctx, const int64_t *counts, double loglik) {
  if (!ctx || !counts) return DBL_ERR_INVALID;
  for (int i = 0; i < ctx->n_counts(); ++i) ctx->h_counts[i] = counts[i];
  ctx->h_loglik_part = loglik;
  return DBL_OK;
}
extern "C" int32_t dbl_summary_words(const dbl_ctx *ctx) { return ctx ? ctx->n_counts() : 0; }
extern "C" int dbl_export_owned_dev(dbl_ctx *ctx, void *y_dev, void *blk_dev, void *link_dev, void *z_dev) {
  if (!ctx || !y_dev || !blk_dev || !link_dev || !z_dev) return DBL_ERR_INVALID;
  if (!ctx->has_state) { ctx->set_error("no state"); return DBL_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_export_ent<<<grid_for(ctx->E, 256), 256, 0, ctx->stream>>>(ctx->E, ctx->A, ctx->ent_owned.p, ctx->y.p, ctx->blk.p,
                                                              (int *)y_dev, (int *)blk_dev);
  k_export_rec<<<grid_for(ctx->R, 256), 256, 0, ctx->stream>>>(ctx->R, ctx->A, ctx->rec_owned.p, ctx->link.p,
                                                              ctx->zmask.p, (int *)link_dev, (unsigned char *)z_dev);
  ctx->launches += 2;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return DBL_OK;
}

extern "C" int dbl_owned_masks(dbl_ctx *ctx, uint8_t *ent_owned, uint8_t *rec_owned) {
  if (!ctx) return DBL_ERR_INVALID;
  if (!ctx->has_state) { ctx->set_error("no state"); return DBL_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (ent_owned) CUDA_TRY(cudaMemcpyAsync(ent_owned, ctx->ent_owned.p, ctx->E, cudaMemcpyDeviceToHost, ctx->stream));
  if (rec_owned) CUDA_TRY(cudaMemcpyAsync(rec_owned, ctx->rec_owned.p, ctx->R, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return DBL_OK;
}

extern "C" int dbl_set_link_mode(dbl_ctx *ctx, int mode) {
  if (!ctx || mode < 0 || mode > 2) return DBL_ERR_INVALID;
  ctx->link_mode = mode;
  return DBL_OK;
}

extern "C" double dbl_last_sweep_ms(const dbl_ctx *ctx) { return ctx ? ctx->last_sweep_ms : 0.0; }

extern "C" double dbl_link_kernel_ms(dbl_ctx *ctx, int64_t *launches) {
  if (!ctx) return 0.0;
  const double ms = ctx->link_ms;
  if (launches) *launches = ctx->link_launches;
  ctx->link_ms = 0.0;
  ctx->link_launches = 0;
  return ms;
}

extern "C" int dbl_summary(dbl_ctx *ctx, dbl_summary_head *head, int64_t *agg_dist, int64_t *rec_dist, double *theta) {
  if (!ctx) return DBL_ERR_INVALID;
  if (!ctx->has_state) { ctx->set_error("no state"); return DBL_ERR_STATE; }
  const int A = ctx->A, F = ctx->F;
  if (head) {
    head->iteration = ctx->iteration;
    head->num_isolates = ctx->h_counts[ctx->iso_slot()];
    double ll = ctx->h_loglik_part;
    for (int a = 0; a < A; ++a)
      for (int f = 0; f < F; ++f) {  // GU:286-293
        const double th = ctx->h_theta[a * F + f];
        const double nd = (double)ctx->h_counts[a * F + f];
        ll += (ctx->alpha[a] + nd - 1.0) * std::log(th) +
              (ctx->beta[a] + (double)ctx->file_sizes[f] - nd - 1.0) * std::log(1.0 - th);
      }
    head->log_likelihood = ll;
    head->pairs_scored = ctx->h_pairs;
  }
  if (agg_dist) for (int i = 0; i < A * F; ++i) agg_dist[i] = ctx->h_counts[i];
  if (rec_dist) for (int i = 0; i <= A; ++i) rec_dist[i] = ctx->h_counts[A * F + i];
  if (theta) std::copy(ctx->h_theta.begin(), ctx->h_theta.end(), theta);
  return DBL_OK;
}
