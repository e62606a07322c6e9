// Internal declarations shared by the host (.cpp) and device (.cu) halves of libdblink_b200.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "dblink_b200.h"

// ---- RNG protocol (DESIGN.md "Draw protocol"): Philox4x32-10, counter = (id, sub, iteration, phase),
// key = 64-bit seed; the same function cuRAND ships as curand_Philox4x32_10.
#if defined(__CUDACC__)
#define DBL_HD __host__ __device__ __forceinline__
#else
#define DBL_HD inline
#endif

enum : uint32_t { PH_INIT = 0, PH_THETA = 1, PH_LINK = 2, PH_VALUE = 3, PH_DIST = 4 };

struct Philox4 {
  uint32_t v[4];
};

DBL_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox4 r;
  r.v[0] = c0; r.v[1] = c1; r.v[2] = c2; r.v[3] = c3;
  return r;
}

// uniform in the open interval (0,1) from 64 random bits: ((bits >> 12) + 0.5) * 2^-52 (exact in binary64)
DBL_HD double unit_from_bits(uint32_t lo, uint32_t hi) {
  const uint64_t x = ((uint64_t)hi << 32) | lo;
  return ((double)(x >> 12) + 0.5) * 2.220446049250313080847263336181640625e-16;
}

struct U2 {
  double u0, u1;
};
DBL_HD U2 uniform2(uint64_t seed, uint32_t phase, uint32_t iter, uint32_t id, uint32_t sub) {
  const Philox4 p = philox4x32_10(id, sub, iter, phase, (uint32_t)seed, (uint32_t)(seed >> 32));
  U2 r;
  r.u0 = unit_from_bits(p.v[0], p.v[1]);
  r.u1 = unit_from_bits(p.v[2], p.v[3]);
  return r;
}

// ---- host-side model objects ------------------------------------------------------------------------
struct dbl_index {
  int32_t V = 0;
  bool is_const = true;
  int32_t kmax = 0;
  std::vector<std::string> values;  // sorted
  std::vector<double> probs;        // weight / total
  std::vector<double> phi;          // probabilityOf (renormalised)
  std::vector<double> norm, invnorm;
  std::vector<int32_t> rowptr, col;
  std::vector<double> expsim;
  std::vector<double> pk, cdf;      // (kmax+1) x V
  std::vector<double> logphi, lognorm;
  // per-row perfect hash of the off-diagonal similar values (link kernel): slot = (uint32(v) * hmult[x]) >> hshift
  int32_t hsize = 0, hshift = 32;  // hsize == 0: no table (constant attribute, or rows too long)
  std::vector<uint32_t> hmult;     // V
  std::vector<int32_t> hkeys;      // V x hsize, -1 = empty
  std::vector<double> hvals;       // V x hsize
  void finish();
  void build_hash(int min_slots = 32);
};

struct dbl_kdtree {
  int32_t n_nodes = 1, n_leaves = 1;
  std::vector<int32_t> attr, kind, split, set_ptr, set_val, leaf_no;
  int32_t leaf_node(const int32_t *yrow) const;
};

// theta draw (GU:305-320) -- host, libm
void host_draw_theta(int A, int F, const double *alpha, const double *beta, uint64_t seed, const int64_t *agg_dist,
                     const int64_t *file_sizes, uint32_t iter, double *theta_out);
int host_levenshtein(const char *a, int la, const char *b, int lb);
// GPU all-pairs candidate distances for the attribute index (dbl_index_gpu.cu); false = not applicable
bool gpu_levenshtein_candidates(const std::vector<std::string> &values, double threshold, double max_sim,
                                std::vector<int> &oi, std::vector<int> &oj, std::vector<int> &od);
double host_similarity_from_distance(int dist, int la, int lb, double threshold, double max_sim);
