// k_link_pcg2<A, NS>: the PCG-II link update (updateEntityIdCollapsed, GU:363-395) with everything about the
// model shape known at compile time: A attributes in kernel order, the last NS of them non-constant.
//
//  * the block's entity table streams through shared memory in TE-entity tiles moved by TMA bulk copies
//    (cp.async.bulk.shared::cluster.global + mbarrier ring, one producer warp per CTA);
//  * each consumer warp owns one record; its constants (value ids, exact-match multipliers, hash multipliers)
//    are registers; the sparse similarity row of each non-constant record attribute is a 32-slot perfect-hash
//    table in shared memory: one key word per bank, so a probe is one conflict-free wavefront;
//  * lane l scores candidate 32*step + l; lane sums / chunk totals / draw as in DESIGN.md section 4.
#pragma once
#include "dbl_link.cuh"

// Per (record, non-constant attribute): H key words then H f64 values, H = p.hslots (a power of two >= 32, the same
// for every attribute of the model; 32 = one key per bank = conflict-free probes).
__device__ __host__ __forceinline__ int pcg2_tab_bytes(int H) { return H * 12; }

// w *= r when y == x (ptxas turns any predicated form into DMUL + 2 FSEL; plain C avoids extra moves)
__device__ __forceinline__ void mul_if_eq(double &w, int y, int x, double r) {
  if (y == x) w = w * r;
}

template <int A, int NS>
struct Pcg2Rec {
  int x[A];                      // record value id; -1 = missing (never equals an entity value)
  double rm[A];                  // multiplier on an exact match
  unsigned hm[NS > 0 ? NS : 1];  // hash multipliers of the non-constant attributes
  unsigned mmask;                // missing non-constant attributes (bit = kernel position)
};

// CONVERGED: every lane of the warp executes the call (main loop), so the rare similar-value multiply is skipped
// warp-wide with a vote; pass 2 calls it under divergence and must not vote.
template <int A, int NS, int HC, bool CONVERGED>
__device__ __forceinline__ double pcg2_weight(const Pcg2Rec<A, NS> &rc, const LinkParams &p, const char *tab,
                                              const int *y, double N) {
  const int hslots = HC ? HC : p.hslots;
  const int hshift = HC ? 27 : p.hshift;
  const int tabb = pcg2_tab_bytes(hslots);
  double w = N;
#pragma unroll
  for (int k = 0; k < A; ++k) mul_if_eq(w, y[k], rc.x[k], rc.rm[k]);
  if (CONVERGED) {
    // probe all NS tables first (no control flow), then ONE vote: the multiply by a similarity is rare
    bool hit[NS > 0 ? NS : 1];
    bool any = false;
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const int yv = y[A - NS + q];
      const unsigned slot = ((unsigned)yv * rc.hm[q]) >> hshift;
      hit[q] = (reinterpret_cast<const int *>(tab + q * tabb)[slot] == yv);
      any = any || hit[q];
    }
    if (NS > 0 && __any_sync(FULL, any)) {
#pragma unroll
      for (int q = 0; q < NS; ++q) {
        if (hit[q]) {
          const unsigned slot = ((unsigned)y[A - NS + q] * rc.hm[q]) >> hshift;
          w = w * reinterpret_cast<const double *>(tab + q * tabb + hslots * 4)[slot];
        }
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const int yv = y[A - NS + q];
      const unsigned slot = ((unsigned)yv * rc.hm[q]) >> hshift;
      if (reinterpret_cast<const int *>(tab + q * tabb)[slot] == yv)
        w = w * reinterpret_cast<const double *>(tab + q * tabb + hslots * 4)[slot];
    }
  }
  if (rc.mmask) {
#pragma unroll
    for (int q = 0; q < NS; ++q)
      if ((rc.mmask >> (A - NS + q)) & 1u) w = w * p.attrs[p.perm[A - NS + q]].invnorm[y[A - NS + q]];
  }
  return w;
}

template <int A, int NS, int HC>
__global__ void __launch_bounds__((LINK_WARPS + 1) * 32, 2) k_link_pcg2(LinkParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int cta = blockIdx.x;
  if (cta >= p.cta_ptr[p.P]) return;
  const int b = find_block(p, cta);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = p.ent_ptr[b + 1] - p.ent_ptr[b];
  const int ntiles = p.tile_ptr[b + 1] - p.tile_ptr[b];
  constexpr int TW = A * TE + 2 * TE;
  TileRing rg;
  rg.tiles = reinterpret_cast<int *>(smem);
  rg.full = reinterpret_cast<uint64_t *>(smem + (size_t)LINK_STAGES * TW * 4);
  rg.empty = rg.full + LINK_STAGES;
  rg.tw = TW;
  static_assert(2 * LINK_STAGES * 8 <= 128, "barrier area");
  char *tab = reinterpret_cast<char *>(smem) + (size_t)LINK_STAGES * TW * 4 + 128 +
              (size_t)warp * (NS > 0 ? NS : 1) * pcg2_tab_bytes(HC ? HC : p.hslots);
  const int *gtiles = p.tiles + (size_t)p.tile_ptr[b] * TW;
  ring_init(rg, LINK_WARPS);

  if (warp == LINK_WARPS) {  // producer warp
    if (lane == 0) ring_produce<true>(rg, gtiles, ntiles);
    return;
  }
  const int ridx = p.rec_ptr[b] + (cta - p.cta_ptr[b]) * LINK_WARPS + warp;
  const bool active = ridx < p.rec_ptr[b + 1];
  const int r = active ? p.rec_sorted[ridx] : -1;

  // ---- per-record constants: lane k prepares kernel-order attribute k, then everything is broadcast
  Pcg2Rec<A, NS> rc;
  {
    int xv = -1;
    double rmv = 1.0;
    unsigned hmv = 0;
    bool is_m = false;
    if (active && lane < A) {
      const int a = p.perm[lane];
      const AttrDev &at = p.attrs[a];
      xv = p.x[(int64_t)r * A + a];
      if (xv < 0) {
        is_m = !at.is_const;
      } else {
        const double th = p.theta[a * p.F + p.file[r]];
        double d = th * at.phi[xv];
        if (at.is_const) {
          rmv = 1.0 + (1.0 - th) / d;
        } else {
          d = d * at.norm[xv];
          double ediag = 1.0;
          row_find(at, xv, xv, ediag);
          rmv = ediag + (1.0 - th) / d;
          hmv = at.hmult[xv];
        }
      }
    }
    rmv = (rmv - 1.0) + 1.0;  // protocol: the multiplier is defined through (r - 1) (identity below ~2^53)
    rc.mmask = __ballot_sync(FULL, is_m);
#pragma unroll
    for (int k = 0; k < A; ++k) {
      rc.x[k] = __shfl_sync(FULL, xv, k);
      rc.rm[k] = shfl_d(rmv, k);
    }
#pragma unroll
    for (int q = 0; q < NS; ++q) rc.hm[q] = __shfl_sync(FULL, hmv, A - NS + q);
    // hash tables of the record's similarity rows -> shared memory (all-empty table when the value is missing)
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const AttrDev &at = p.attrs[p.perm[A - NS + q]];
      const int H = HC ? HC : p.hslots;
      int *kd = reinterpret_cast<int *>(tab + q * pcg2_tab_bytes(H));
      double *vd = reinterpret_cast<double *>(tab + q * pcg2_tab_bytes(H) + H * 4);
      const int xq = rc.x[A - NS + q];
      for (int i = lane; i < H; i += 32) {
        kd[i] = (xq >= 0) ? at.hkeys[(size_t)xq * H + i] : -1;
        vd[i] = (xq >= 0) ? at.hvals[(size_t)xq * H + i] : 1.0;
      }
    }
    __syncwarp();
  }

  const int nsteps = ntiles * (TE / 32);          // steps beyond the last candidate add zeros
  const int tpc = max(1, (ntiles + 31) >> 5);     // a chunk is a whole number of tiles
  const int spc = (TE / 32) * tpc;
  const int nchunks = (nsteps + spc - 1) / spc;

  // ---- pass 1 over the TMA-staged tiles
  double run = 0.0, Q = 0.0, acc = 0.0;
  int chunk = 0, tile_in_chunk = 0;
#ifdef DBL_EXP_WAITCLK
  long long wclk_ = 0;
  const long long cstart_ = clock64();
#endif
  for (int t = 0; t < ntiles; ++t) {
    const int s = t % LINK_STAGES;
#ifdef DBL_EXP_WAITCLK
    const long long c0_ = clock64();
#endif
    mbar_wait(&rg.full[s], (t / LINK_STAGES) & 1);
#ifdef DBL_EXP_WAITCLK
    wclk_ += clock64() - c0_;
#endif
    if (active) {
      const int *tile = rg.tiles + (size_t)s * TW;
      const double *tileN = reinterpret_cast<const double *>(tile + A * TE);
#pragma unroll
      for (int q = 0; q < TE / 32; ++q) {
        const int slot = q * 32 + lane;
        int y[A];
#pragma unroll
        for (int k = 0; k < A; ++k) y[k] = tile[k * TE + slot];
        acc = acc + pcg2_weight<A, NS, HC, true>(rc, p, tab, y, tileN[slot]);
      }
      if (++tile_in_chunk == tpc || t + 1 == ntiles) {
        run = run + butterfly_sum(acc);
        if (lane == chunk) Q = run;
        ++chunk;
        acc = 0.0;
        tile_in_chunk = 0;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&rg.empty[s]);
  }
#ifdef DBL_EXP_WAITCLK
  const long long cloop_ = clock64() - cstart_;
#endif
  if (!active) return;
  if (!(run > 0.0) || isinf(run)) { fail_link(p, lane, r); return; }

  // ---- pass 2 from the L2-resident copy of the tiles
  auto wf = [&](int j) -> double {
    if (j >= n) return 0.0;
    const int *tile = gtiles + (size_t)(j / TE) * TW;
    const int slot = j % TE;
    int y[A];
#pragma unroll
    for (int k = 0; k < A; ++k) y[k] = tile[k * TE + slot];
    return pcg2_weight<A, NS, HC, false>(rc, p, tab, y, reinterpret_cast<const double *>(tile + A * TE)[slot]);
  };
  const U2 u = uniform2(p.seed, PH_LINK, p.iter, (uint32_t)r, 0u);
  const int j = finish_draw(lane, n, nsteps, spc, nchunks, Q, run, u.u0, wf);
  store_link(p, lane, r, b, n, j);
#ifdef DBL_EXP_WAITCLK
  if (lane == 0 && (cta % 997) == 0 && p.iter == 2)
    printf("cta %d warp %d ntiles %d loop %lld wait %lld total %lld\n", cta, warp, ntiles, cloop_, wclk_, clock64() - cstart_);
#endif
}

inline size_t pcg2_smem_bytes(int A, int NS, int H) {
  return (size_t)LINK_STAGES * tile_words(A) * 4 + 128 + (size_t)LINK_WARPS * (NS > 0 ? NS : 1) * pcg2_tab_bytes(H);
}

// launch k_link_pcg2<A, NS, HC> for a runtime NS in [0, A]; HC = 32 (compile-time table size) when the model's
// tables have 32 slots, else 0 (size read from the parameters); returns cudaError_t as int
template <int A, int NS, int HC>
int pcg2_launch_one(int grid, cudaStream_t stream, const LinkParams &lp) {
  const size_t smem = pcg2_smem_bytes(A, NS, lp.hslots);
  static size_t configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(k_link_pcg2<A, NS, HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  k_link_pcg2<A, NS, HC><<<grid, (LINK_WARPS + 1) * 32, smem, stream>>>(lp);
  return (int)cudaGetLastError();
}

template <int A, int NS>
struct Pcg2Launch {
  static int go(int ns, int grid, cudaStream_t stream, const LinkParams &lp) {
    if (ns == NS)
      return lp.hslots == 32 ? pcg2_launch_one<A, NS, 32>(grid, stream, lp) : pcg2_launch_one<A, NS, 0>(grid, stream, lp);
    if constexpr (NS > 0) return Pcg2Launch<A, NS - 1>::go(ns, grid, stream, lp);
    return (int)cudaErrorInvalidValue;
  }
};
