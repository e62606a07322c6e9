// One translation unit per attribute count: compiled with -DDBL_INST_A=<A>; instantiates k_link_pcg2<A, 0..A, HC, PK>.
#include "dbl_link_pcg2.cuh"

#ifndef DBL_INST_A
#error "compile with -DDBL_INST_A=<number of attributes>"
#endif

#define DBL_CAT2(a, b) a##b
#define DBL_CAT(a, b) DBL_CAT2(a, b)

int DBL_CAT(dbl_launch_pcg2_a, DBL_INST_A)(int ns, int grid, cudaStream_t stream, const LinkParams &lp, size_t *smem_configured) {
  return Pcg2Launch<DBL_INST_A, DBL_INST_A>::go(ns, grid, stream, lp, smem_configured);
}
