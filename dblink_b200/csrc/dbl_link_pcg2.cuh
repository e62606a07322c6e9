// k_link_pcg2<A, NS, HC, PK>: the PCG-II link update (updateEntityIdCollapsed, GU:363-395) with everything about
// the model shape known at compile time: A attributes in kernel order, the last NS of them non-constant; HC = 32
// when the hash tables have 32 slots (else the size is a run-time parameter); PK = the constant attributes arrive
// byte-packed.
//
//  * the block's entity table streams through shared memory in TE-entity tiles moved by TMA bulk copies
//    (cp.async.bulk.shared::cluster.global + mbarrier ring, one producer warp per CTA);
//  * each consumer warp owns one record; its constants (value ids, hash multipliers) are registers;
//  * exact matches: the multipliers of the matching constant attributes and of the matching non-constant
//    attributes are two products, each fetched from a per-record table in shared memory indexed by
//    the match mask (16 entries for PK, 2^NS entries for NS <= 8);
//  * similar-but-different values: the sparse similarity row of each non-constant record attribute is a 32-slot
//    perfect-hash table in shared memory, one key word per bank, so a probe is one conflict-free wavefront; one
//    warp vote per step decides whether anybody needs the multiply;
//  * lane l scores candidate 32*step + l; lane sums / chunk totals / draw as in DESIGN.md section 4.
#pragma once
#include "dbl_link.cuh"

// Per (record, non-constant attribute): H key words then H f64 values, H = p.hslots (a power of two >= 32, the same
// for every attribute of the model; 32 = one key per bank = conflict-free probes).
__device__ __host__ __forceinline__ int pcg2_tab_bytes(int H) { return H * 12; }

// w *= r when y == x (shapes without a product table; ptxas turns any predicated form into DMUL + 2 FSEL)
__device__ __forceinline__ void mul_if_eq(double &w, int y, int x, double r) {
  if (y == x) w = w * r;
}

template <int A, int NS>
struct Pcg2Rec {
  int x[A];                      // record value id; -1 = missing (never equals an entity value)
  double rm[A];                  // multiplier on an exact match
  unsigned hm[NS > 0 ? NS : 1];  // hash multipliers of the non-constant attributes
  unsigned mmask;                // missing non-constant attributes (bit = kernel position)
  unsigned xpack;                // PK: the record's constant-attribute values, one byte each (0xFF = cannot match)
};

// up to 8 non-constant attributes: the product of the exact-match multipliers comes from a per-record table indexed
// by the match mask (2^NS doubles per warp in shared memory)
__host__ __device__ constexpr bool pcg2_dtab(int NS) { return NS >= 1 && NS <= 8; }
// Position (in entries) of attribute q's bit in that table.  Not 1 << q: the weights are chosen so that every subset
// sum is distinct AND the single-attribute entries (by far the most used after entry 0) fall in different
// shared-memory banks from entry 0 and from each other (position mod 16 all distinct and non-zero); with powers of
// two, attributes 4 and 5 would sit 128 and 256 bytes from entry 0, i.e. in its banks.
__host__ __device__ constexpr int pcg2_dtab_weight(int q) {
  constexpr int w[8] = {1, 2, 4, 8, 19, 38, 75, 149};
  return w[q];
}
__host__ __device__ constexpr int pcg2_dtab_entries(int NS) {
  int n = 1;
  for (int q = 0; q < NS; ++q) n += pcg2_dtab_weight(q);
  return pcg2_dtab(NS) ? n : 0;
}

// PK kernels: index into the record's table of constant-attribute products from the byte-packed values of a
// candidate: bit k of the index = (byte k of ypack == byte k of xpack)
__device__ __forceinline__ unsigned pcg2_const_index(unsigned ypack, unsigned xpack) {
  const unsigned d = ypack ^ xpack;
  const unsigned nz = ((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d;  // bit 7 of a byte set <=> the byte of d is non-zero
  const unsigned eq = (~nz & 0x80808080u) >> 7;               // bit 8k set <=> byte k equal
  return (eq * 0x01020408u) >> 24;                            // gathers bits 0, 8, 16, 24 into bits 0..3
}

// CONVERGED: every lane of the warp executes the call (main loop), so the rare similar-value multiply is skipped
// warp-wide with a vote; pass 2 calls it under divergence and must not vote.
// PK: the A - NS constant attributes (1..4 of them, vocabularies <= 255) travel as one byte-packed word per
// candidate (ypack) and their product comes from the record's 16-entry table ctab; otherwise y[0..A-NS) are used.
template <int A, int NS, int HC, bool CONVERGED, bool PK>
__device__ __forceinline__ double pcg2_weight(const Pcg2Rec<A, NS> &rc, const LinkParams &p, const char *tab,
                                              const double *ctab, const double *dtab, const int *y, unsigned ypack,
                                              double N) {
  const int hslots = HC ? HC : p.hslots;
  const int hshift = HC ? 27 : p.hshift;
  const int tabb = pcg2_tab_bytes(hslots);
  double w = N;
  if constexpr (NS < A) {  // protocol 4.1: the constant attributes form their own product c; w = N * c
    double c = 1.0;
    if constexpr (PK) {
      c = ctab[pcg2_const_index(ypack, rc.xpack)];
    } else {
#pragma unroll
      for (int k = 0; k < A - NS; ++k) mul_if_eq(c, y[k], rc.x[k], rc.rm[k]);
    }
    w = w * c;
  }
  if constexpr (NS >= 1) {  // protocol 4.1: so do the exact matches of the non-constant attributes (product d)
    double d = 1.0;
    if constexpr (pcg2_dtab(NS)) {
      unsigned di = 0;  // byte offset into the table: bit q of the index = attribute q matches
#pragma unroll
      for (int q = 0; q < NS; ++q) di += (y[A - NS + q] == rc.x[A - NS + q]) ? 8u * pcg2_dtab_weight(q) : 0u;
      d = *reinterpret_cast<const double *>(reinterpret_cast<const char *>(dtab) + di);
    } else {
#pragma unroll
      for (int k = A - NS; k < A; ++k) mul_if_eq(d, y[k], rc.x[k], rc.rm[k]);
    }
    w = w * d;
  }
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

template <int A, int NS, int HC, bool PK>
__global__ void __launch_bounds__((LINK_WARPS + 1) * 32, 2) k_link_pcg2(LinkParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int cta = blockIdx.x;
  if (sweep_dead(p.ctl) || cta >= p.cta_ptr[p.P]) return;
  const int b = find_block(p, cta);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = p.ent_ptr[b + 1] - p.ent_ptr[b];
  const int ntiles = p.tile_ptr[b + 1] - p.tile_ptr[b];
  constexpr int TW = A * TE + 3 * TE;  // tile_words(A)
  constexpr int NC = A - NS;
  TileRing rg;
  rg.tiles = reinterpret_cast<int *>(smem);
  rg.full = reinterpret_cast<uint64_t *>(smem + (size_t)LINK_STAGES * TW * 4);
  rg.empty = rg.full + LINK_STAGES;
  rg.tw = TW;
  static_assert(2 * LINK_STAGES * 8 <= 128, "barrier area");
  char *tab = reinterpret_cast<char *>(smem) + (size_t)LINK_STAGES * TW * 4 + 128 +
              (size_t)warp * (NS > 0 ? NS : 1) * pcg2_tab_bytes(HC ? HC : p.hslots);
  double *ctab = reinterpret_cast<double *>(reinterpret_cast<char *>(smem) + (size_t)LINK_STAGES * TW * 4 + 128 +
                                           (size_t)LINK_WARPS * (NS > 0 ? NS : 1) * pcg2_tab_bytes(HC ? HC : p.hslots)) +
                 warp * 16;  // PK: products of the matching constant attributes, by match mask
  double *dtab = ctab + (LINK_WARPS - warp) * 16 + warp * pcg2_dtab_entries(NS);  // same for the others
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
    rc.xpack = 0xFFFFFFFFu;
    if constexpr (PK) {
      static_assert(!PK || (NC >= 1 && NC <= 4), "PK packs 1..4 constant attributes");
#pragma unroll
      for (int k = 0; k < NC; ++k)
        rc.xpack = (rc.xpack & ~(0xFFu << (8 * k))) | ((unsigned)(rc.x[k] < 0 ? 0xFF : rc.x[k]) << (8 * k));
      if (lane < 16) {
        double c = 1.0;
#pragma unroll
        for (int k = 0; k < NC; ++k)
          if ((lane >> k) & 1) c = c * rc.rm[k];
        ctab[lane] = c;
      }
    }
    if constexpr (pcg2_dtab(NS)) {
      for (int idx = lane; idx < (1 << NS); idx += 32) {  // idx = match mask; stored at its weighted position
        double d = 1.0;
        int pos = 0;
#pragma unroll
        for (int q = 0; q < NS; ++q)
          if ((idx >> q) & 1) { d = d * rc.rm[NC + q]; pos += pcg2_dtab_weight(q); }
        dtab[pos] = d;
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
        for (int k = PK ? NC : 0; k < A; ++k) y[k] = tile[k * TE + slot];
        const unsigned ypack = PK ? (unsigned)tile[(A + 2) * TE + slot] : 0u;
        acc = acc + pcg2_weight<A, NS, HC, true, PK>(rc, p, tab, ctab, dtab, y, ypack, tileN[slot]);
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
    for (int k = PK ? NC : 0; k < A; ++k) y[k] = tile[k * TE + slot];
    const unsigned ypack = PK ? (unsigned)tile[(A + 2) * TE + slot] : 0u;
    return pcg2_weight<A, NS, HC, false, PK>(rc, p, tab, ctab, dtab, y, ypack,
                                             reinterpret_cast<const double *>(tile + A * TE)[slot]);
  };
  const U2 u = uniform2(p.seed, PH_LINK, link_iter(p), (uint32_t)r, 0u);
  const int j = finish_draw(lane, n, nsteps, spc, nchunks, Q, run, u.u0, wf);
  store_link(p, lane, r, b, n, j);
#ifdef DBL_EXP_WAITCLK
  if (lane == 0 && (cta % 997) == 0 && link_iter(p) == 2)
    printf("cta %d warp %d ntiles %d loop %lld wait %lld total %lld\n", cta, warp, ntiles, cloop_, wclk_, clock64() - cstart_);
#endif
}

inline size_t pcg2_smem_bytes(int A, int NS, int H) {
  return (size_t)LINK_STAGES * tile_words(A) * 4 + 128 + (size_t)LINK_WARPS * (NS > 0 ? NS : 1) * pcg2_tab_bytes(H) +
         (size_t)LINK_WARPS * (16 + pcg2_dtab_entries(NS)) * sizeof(double);
}

// launch k_link_pcg2<A, NS, HC> for a runtime NS in [0, A]; HC = 32 (compile-time table size) when the model's
// tables have 32 slots, else 0 (size read from the parameters); returns cudaError_t as int
template <int A, int NS, int HC, bool PK>
int pcg2_launch_one(int grid, cudaStream_t stream, const LinkParams &lp, size_t *configured) {
  const size_t smem = pcg2_smem_bytes(A, NS, lp.hslots);
  // the opt-in is per device: the cache belongs to the context (one model shape = one instantiation per context)
  if (*configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(k_link_pcg2<A, NS, HC, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    *configured = smem;
  }
  k_link_pcg2<A, NS, HC, PK><<<grid, (LINK_WARPS + 1) * 32, smem, stream>>>(lp);
  return (int)cudaGetLastError();
}

template <int A, int NS>
struct Pcg2Launch {
  static int go(int ns, int grid, cudaStream_t stream, const LinkParams &lp, size_t *cfg) {
    if (ns == NS) {
      // byte-packed constant attributes: 1..4 of them, every vocabulary <= 255 (lp.pack_consts), 32-slot tables
      if constexpr (A - NS >= 1 && A - NS <= 4) {
        if (lp.pack_consts && lp.hslots == 32) return pcg2_launch_one<A, NS, 32, true>(grid, stream, lp, cfg);
      }
      return lp.hslots == 32 ? pcg2_launch_one<A, NS, 32, false>(grid, stream, lp, cfg)
                             : pcg2_launch_one<A, NS, 0, false>(grid, stream, lp, cfg);
    }
    if constexpr (NS > 0) return Pcg2Launch<A, NS - 1>::go(ns, grid, stream, lp, cfg);
    return (int)cudaErrorInvalidValue;
  }
};
