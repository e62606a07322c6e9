// All-pairs thresholded Levenshtein on the GPU for AttributeIndex construction
// (computeSimValueIndex, AttributeIndex.scala:219-231: V^2 pairs, keep exp(sim) > 1).
//
// The device computes only the INTEGER edit distance of the pairs that can possibly have a positive truncated
// similarity (length filter); the host turns (distance, lengths) into sim and exp(sim) with the same double
// arithmetic as the host-only path, so both paths produce identical tables.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "dbl_internal.h"

namespace {
constexpr int MAXL = 64;

__global__ void k_lev_pairs(int V, const char *__restrict__ strs, const int *__restrict__ lens, double ratio,
                            unsigned long long cap, unsigned long long *__restrict__ count, int4 *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= V || j <= i) return;
  const int la = lens[i], lb = lens[j];
  // sim > 0  <=>  1 - 2d/(L+d) > thr/max  <=>  d < L*(1-t)/(1+t);  +1 keeps the filter conservative
  const int maxd = (int)((double)(la + lb) * ratio) + 1;
  const int diff = la > lb ? la - lb : lb - la;
  if (diff > maxd) return;
  unsigned char a[MAXL], b[MAXL];
  for (int k = 0; k < la; ++k) a[k] = (unsigned char)strs[(size_t)i * MAXL + k];
  for (int k = 0; k < lb; ++k) b[k] = (unsigned char)strs[(size_t)j * MAXL + k];
  unsigned char row[MAXL + 1];
  for (int k = 0; k <= lb; ++k) row[k] = (unsigned char)k;
  for (int p = 1; p <= la; ++p) {
    int diag = row[0];
    row[0] = (unsigned char)p;
    const unsigned char ca = a[p - 1];
    for (int q = 1; q <= lb; ++q) {
      const int up = row[q];
      int best = diag + (ca != b[q - 1]);
      best = min(best, up + 1);
      best = min(best, (int)row[q - 1] + 1);
      row[q] = (unsigned char)best;
      diag = up;
    }
  }
  const int d = row[lb];
  if (d > maxd) return;
  const unsigned long long slot = atomicAdd(count, 1ull);
  if (slot < cap) out[slot] = make_int4(i, j, d, 0);
}
}  // namespace

// Fills `out` with (i, j>i, distance) for every pair that may have a positive similarity.  Returns false when the
// GPU path is not applicable (no device, strings too long) -- the caller then uses the host loop.
bool gpu_levenshtein_candidates(const std::vector<std::string> &values, double threshold, double max_sim,
                                std::vector<int> &oi, std::vector<int> &oj, std::vector<int> &od) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return false; }
  const int V = (int)values.size();
  for (auto &s : values)
    if ((int)s.size() > MAXL) return false;
  std::vector<char> h_strs((size_t)V * MAXL, 0);
  std::vector<int> h_lens(V);
  for (int v = 0; v < V; ++v) {
    h_lens[v] = (int)values[v].size();
    std::memcpy(&h_strs[(size_t)v * MAXL], values[v].data(), values[v].size());
  }
  const double t = threshold / max_sim;
  const double ratio = (1.0 - t) / (1.0 + t);
  char *d_strs = nullptr;
  int *d_lens = nullptr;
  unsigned long long *d_count = nullptr;
  int4 *d_out = nullptr;
  bool ok = true;
  unsigned long long cap = std::max<unsigned long long>((unsigned long long)V * 64ull, 1ull << 20);
  auto fail = [&]() { ok = false; };
  if (cudaMalloc(&d_strs, h_strs.size()) != cudaSuccess) fail();
  if (ok && cudaMalloc(&d_lens, sizeof(int) * V) != cudaSuccess) fail();
  if (ok && cudaMalloc(&d_count, sizeof(unsigned long long)) != cudaSuccess) fail();
  if (ok) {
    cudaMemcpy(d_strs, h_strs.data(), h_strs.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_lens, h_lens.data(), sizeof(int) * V, cudaMemcpyHostToDevice);
  }
  for (int attempt = 0; ok && attempt < 3; ++attempt) {
    if (cudaMalloc(&d_out, sizeof(int4) * cap) != cudaSuccess) { fail(); break; }
    cudaMemset(d_count, 0, sizeof(unsigned long long));
    dim3 grid((V + 127) / 128, V);
    k_lev_pairs<<<grid, 128>>>(V, d_strs, d_lens, ratio, cap, d_count, d_out);
    unsigned long long n = 0;
    if (cudaMemcpy(&n, d_count, sizeof(n), cudaMemcpyDeviceToHost) != cudaSuccess) { fail(); break; }
    if (n <= cap) {
      std::vector<int4> h((size_t)n);
      if (n) cudaMemcpy(h.data(), d_out, sizeof(int4) * n, cudaMemcpyDeviceToHost);
      oi.resize(n); oj.resize(n); od.resize(n);
      for (size_t k = 0; k < n; ++k) { oi[k] = h[k].x; oj[k] = h[k].y; od[k] = h[k].z; }
      cudaFree(d_out);
      d_out = nullptr;
      break;
    }
    cudaFree(d_out);
    d_out = nullptr;
    cap = n + 1024;
    if (attempt == 2) fail();
  }
  if (d_out) cudaFree(d_out);
  if (d_strs) cudaFree(d_strs);
  if (d_lens) cudaFree(d_lens);
  if (d_count) cudaFree(d_count);
  if (!ok) cudaGetLastError();
  return ok;
}
