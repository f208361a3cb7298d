// pattern.cuh — block-CSR pattern construction on the device: replaces lhsa_ns::lhsa (Code/Source/solver/lhsa.cpp:153-380)
// for meshes without undeformed-Neumann rewiring (idMap = identity) and without shells.  SURVEY.md par. 8(f) row 3: the
// reference builds the pattern serially with per-row insertion lists (add_col, lhsa.cpp:46-95), minutes at 80 M tets.
//
// Here: every (element, a, b) emits the key (row << bits | col), bits = ceil(log2 nNo); the keys are radix-sorted on
// their 2*bits significant bits, duplicates removed, and the CSR read off the unique keys: colPtr = low words (sorted inside a
// row, diagonal present because a == b is emitted), rowPtr by a lower bound per row.  Same rowPtr / colPtr as the
// reference's, integer for integer.  Sort / unique are CUB device primitives (library sorts, like cuBLAS for a GEMM).
#pragma once

#include <cub/cub.cuh>

#include "kernels.cuh"

namespace svb200 {

__global__ void k_pattern_keys(size_t nEl, int eNoN, int bits, const int* __restrict__ ien, unsigned long long* __restrict__ keys)
{
  const size_t tot = nEl*size_t(eNoN)*eNoN;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t t = size_t(blockIdx.x)*blockDim.x + threadIdx.x; t < tot; t += nth) {
    const size_t e = t/(size_t(eNoN)*eNoN);
    const int r = int(t % (size_t(eNoN)*eNoN));
    const unsigned long long row = (unsigned long long)ien[e*eNoN + r/eNoN];
    const unsigned long long col = (unsigned long long)ien[e*eNoN + r % eNoN];
    keys[t] = (row << bits) | col;
  }
}

// rowPtr[r] = first position whose key has row >= r (r = 0..nNo); colPtr[i] = low word
__global__ void k_pattern_csr(int nNo, int bits, size_t nnz, const unsigned long long* __restrict__ ukeys, int* __restrict__ rowPtr, int* __restrict__ colPtr)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t t = size_t(blockIdx.x)*blockDim.x + threadIdx.x; t < nnz + size_t(nNo) + 1; t += nth) {
    if (t < nnz) {
      colPtr[t] = int(ukeys[t] & ((1ull << bits) - 1ull));
    } else {
      const unsigned long long r = (unsigned long long)(t - nnz);
      const unsigned long long target = r << bits;
      size_t lo = 0, hi = nnz;
      while (lo < hi) {
        const size_t mid = (lo + hi) >> 1;
        if (ukeys[mid] < target) lo = mid + 1; else hi = mid;
      }
      rowPtr[r] = int(lo);
    }
  }
}

} // namespace svb200
