// kernels.cuh — hand-written FP64 kernels for sm_100a: dof-blocked CSR SpMV in the four FSILS
// shapes, deterministic multi-dot / Gram-Schmidt kernels, Jacobi scaling, depart, halo pack/add,
// resistance-face rank-1 update.  All kernels are HBM-bound streaming kernels:
// 256-bit / 128-bit vector loads, one contiguous block row per lane group, streaming cache hints on
// the matrix (read once per product) and default/evict-last on the gathered vector (re-used from L2).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "peer_comm.cuh"
#include "krylov.hpp"

namespace svb200 {

constexpr int kSmCount = 148;           // B200: 2 dies x 74 SMs; grids are sized in multiples of this
constexpr int kRedBlocks = kSmCount*4;  // CTAs of every two-stage reduction
constexpr int kRedThreads = 256;
constexpr int kDotJB = 8;               // basis vectors per register batch inside the multi-dot kernel
constexpr int kMaxDots = 256;           // basis vectors per multi-dot launch (partials: kRedBlocks x kMaxDots)
constexpr int kMaxComb = 32;            // vectors per lin_comb launch

// ---- vector memory helpers -------------------------------------------------------------------
struct d4 { double x, y, z, w; };

// streaming 256-bit load: matrix data is touched once per product -> do not pollute L1, evict first from L2
__device__ __forceinline__ d4 ld256_stream(const double* p)
{
  d4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
// gathered vector entries are re-used by ~15 rows: keep them
__device__ __forceinline__ d4 ld256_keep(const double* p)
{
  d4 v;
  asm volatile("ld.global.nc.L2::evict_last.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ d4 ld256(const double* p)
{
  d4 v;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st256(double* p, const d4& v)
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
// staging data is written once and read once by another kernel: keep it out of L1, evict first from L2
__device__ __forceinline__ void st256_stream(double* p, const d4& v)
{
  asm volatile("st.global.L1::no_allocate.L2::evict_first.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ double ld_stream(const double* p)
{
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
// scalar forms of the two cache policies of the dof-4 kernel: matrix entries are read once per product (evict first),
// gathered vector entries are re-used by ~15 rows and by the next kernel of the Krylov step (evict last).  Below 256 bits the
// L2 eviction priority cannot be named in the instruction: it is passed as a cache-policy operand (createpolicy).
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double ld_hint(const double* p, uint64_t pol)
{
  double v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ double ld_hint_na(const double* p, uint64_t pol)      // + no L1 allocation (streamed once)
{
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int ld_stream_i(const int* p)
{
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// =================================================================================================
// K1  block SpMV  KU = K U   (fsils_spar_mul_vv, liner_solver/spar_mul.cpp:191-260)
// One group of 4 lanes per row; lane i owns component i of the row: it streams block-row i of every
// 4x4 block (one 256-bit load), gathers the 4-vector of the column node (one 256-bit load, same
// address for the 4 lanes -> one sector) and accumulates in the reference's order (blocks left to
// right, j = 0..3 inside).  A warp therefore reads 8 full 128-byte lines per step and writes 8
// consecutive 32-byte results.  rowPtr is the standard (nNo+1) CSR pointer in solver ordering.
// =================================================================================================
__global__ void __launch_bounds__(256)
k_spmv_vv4(const int* __restrict__ skip, int nNo, const int* __restrict__ rowPtr, const int* __restrict__ col,
           const double* __restrict__ K, const double* __restrict__ U, double* __restrict__ KU)
{
  if (skip && *skip) return;
  const int lane4 = threadIdx.x & 3;
  const int group = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int ngroups = (gridDim.x*blockDim.x) >> 2;
  for (int row = group; row < nNo; row += ngroups) {
    const int s = __ldg(rowPtr + row);
    const int e = __ldg(rowPtr + row + 1);
    double acc = 0.0;
    int p = s;
    // 4 blocks in flight per lane (8 x 256-bit loads)
    for (; p + 4 <= e; p += 4) {
      const int c0 = ld_stream_i(col + p), c1 = ld_stream_i(col + p + 1), c2 = ld_stream_i(col + p + 2), c3 = ld_stream_i(col + p + 3);
      const d4 k0 = ld256_stream(K + (size_t(p)*16 + lane4*4));
      const d4 k1 = ld256_stream(K + (size_t(p+1)*16 + lane4*4));
      const d4 k2 = ld256_stream(K + (size_t(p+2)*16 + lane4*4));
      const d4 k3 = ld256_stream(K + (size_t(p+3)*16 + lane4*4));
      const d4 u0 = ld256_keep(U + size_t(c0)*4);
      const d4 u1 = ld256_keep(U + size_t(c1)*4);
      const d4 u2 = ld256_keep(U + size_t(c2)*4);
      const d4 u3 = ld256_keep(U + size_t(c3)*4);
      acc = acc + k0.x*u0.x + k0.y*u0.y + k0.z*u0.z + k0.w*u0.w;
      acc = acc + k1.x*u1.x + k1.y*u1.y + k1.z*u1.z + k1.w*u1.w;
      acc = acc + k2.x*u2.x + k2.y*u2.y + k2.z*u2.z + k2.w*u2.w;
      acc = acc + k3.x*u3.x + k3.y*u3.y + k3.z*u3.z + k3.w*u3.w;
    }
    for (; p < e; p++) {
      const int c0 = ld_stream_i(col + p);
      const d4 k0 = ld256_stream(K + (size_t(p)*16 + lane4*4));
      const d4 u0 = ld256_keep(U + size_t(c0)*4);
      acc = acc + k0.x*u0.x + k0.y*u0.y + k0.z*u0.z + k0.w*u0.w;
    }
    KU[size_t(row)*4 + lane4] = acc;
  }
}

// Generic small-dof variant (dof 1..3; mK of the NS solver is dof 3: 72-byte blocks).  4 lanes per
// row, lane i < DOF owns component i.
template <int DOF, bool HINT = false>
__global__ void __launch_bounds__(256)
k_spmv_vv(const int* __restrict__ skip, int nNo, const int* __restrict__ rowPtr, const int* __restrict__ col,
          const double* __restrict__ K, const double* __restrict__ U, double* __restrict__ KU)
{
  if (skip && *skip) return;
  const int lane4 = threadIdx.x & 3;
  const int group = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int ngroups = (gridDim.x*blockDim.x) >> 2;
  const uint64_t pol_k = HINT ? l2_policy_evict_first() : 0, pol_u = HINT ? l2_policy_evict_last() : 0;
  for (int row = group; row < nNo; row += ngroups) {
    const int s = __ldg(rowPtr + row);
    const int e = __ldg(rowPtr + row + 1);
    if (lane4 < DOF) {
      double acc = 0.0;
#pragma unroll 4
      for (int p = s; p < e; p++) {
        const int c = __ldg(col + p);
        const double* k = K + (size_t(p)*DOF*DOF + lane4*DOF);
        const double* u = U + size_t(c)*DOF;
        double t = acc;
#pragma unroll
        for (int j = 0; j < DOF; j++) t = t + (HINT ? __ldg(k + j)*ld_hint(u + j, pol_u) : __ldg(k + j)*__ldg(u + j));
        acc = t;
      }
      KU[size_t(row)*DOF + lane4] = acc;
    }
  }
}

// dof-3 variant with all four lanes busy (mK of the NS solver, 72-byte blocks): lane l takes the blocks
// p = s+l, s+l+4, ... of the row (whole 3x3 block + the 3-vector of its column), the quad then adds
// its partial 3-vectors in a fixed order.  Per step a quad streams 288 contiguous bytes.
__global__ void __launch_bounds__(256)
k_spmv_vv3s(const int* __restrict__ skip, int nNo, const int* __restrict__ rowPtr, const int* __restrict__ col,
            const double* __restrict__ K, const double* __restrict__ U, double* __restrict__ KU)
{
  if (skip && *skip) return;
  const int lane4 = threadIdx.x & 3;
  const int group = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int ngroups = (gridDim.x*blockDim.x) >> 2;
  const int nrounds = (nNo + ngroups - 1)/ngroups;
  for (int r = 0; r < nrounds; r++) {
    const int row = group + r*ngroups;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    if (row < nNo) {
      const int s = __ldg(rowPtr + row);
      const int e = __ldg(rowPtr + row + 1);
#pragma unroll 2
      for (int p = s + lane4; p < e; p += 4) {
        const int c = __ldg(col + p);
        const double* k = K + size_t(p)*9;
        const double* u = U + size_t(c)*3;
        const double u0 = __ldg(u), u1 = __ldg(u + 1), u2 = __ldg(u + 2);
        a0 = a0 + (__ldg(k)*u0 + __ldg(k + 1)*u1 + __ldg(k + 2)*u2);
        a1 = a1 + (__ldg(k + 3)*u0 + __ldg(k + 4)*u1 + __ldg(k + 5)*u2);
        a2 = a2 + (__ldg(k + 6)*u0 + __ldg(k + 7)*u1 + __ldg(k + 8)*u2);
      }
    }
    a0 += __shfl_xor_sync(0xffffffffu, a0, 1); a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
    a0 += __shfl_xor_sync(0xffffffffu, a0, 2); a1 += __shfl_xor_sync(0xffffffffu, a1, 2); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
    if (row < nNo && lane4 < 3) KU[size_t(row)*3 + lane4] = (lane4 == 0) ? a0 : (lane4 == 1) ? a1 : a2;
  }
}

// dof-3 variant with COLUMN-owner lanes: lane i < 3 of the quad loads the three words 3*jj + i (jj = 0..2) of every 3x3 block,
// i.e. each load instruction of a quad reads 24 CONSECUTIVE bytes (one or two sectors) instead of three words 24 bytes apart, and
// the lane needs only ONE gathered vector entry u_i per block (the quad's gather is again 24 consecutive bytes).  Lane i
// accumulates the column-i contributions to the three outputs over the whole row; the quad adds the three partial 3-vectors once
// per row in a fixed order.  L1 sector requests per block drop from ~11 (profiles/r01_tour_c_ncu_raw.csv) to ~6; the sum is
// re-associated (columns outer, blocks inner), a rounding-level difference like the strided variants.
template <bool HINT, int MINB = 6>
__global__ void __launch_bounds__(256, MINB)
k_spmv_vv3c(const int* __restrict__ skip, int nNo, const int* __restrict__ rowPtr, const int* __restrict__ col,
            const double* __restrict__ K, const double* __restrict__ U, double* __restrict__ KU)
{
  if (skip && *skip) return;
  const int lane4 = threadIdx.x & 3;
  const int group = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int ngroups = (gridDim.x*blockDim.x) >> 2;
  const int nrounds = (nNo + ngroups - 1)/ngroups;
  const int li = lane4 < 3 ? lane4 : 0;          // lane 3 idles on a duplicate of lane 0's addresses (its sums are discarded)
  const uint64_t pol_u = HINT ? l2_policy_evict_last() : 0;
  for (int r = 0; r < nrounds; r++) {
    const int row = group + r*ngroups;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    if (row < nNo) {
      const int s = __ldg(rowPtr + row);
      const int e = __ldg(rowPtr + row + 1);
#pragma unroll 4
      for (int p = s; p < e; p++) {
        // (the three words of a lane share sectors with its neighbours' words of the NEXT instruction: they must allocate in L1 -
        // with L1::no_allocate every instruction refetches its sectors from L2 and the kernel drops to 0.60 of the HBM peak)
        const int c = __ldg(col + p);
        const double* k = K + size_t(p)*9 + li;
        const double u = HINT ? ld_hint(U + size_t(c)*3 + li, pol_u) : __ldg(U + size_t(c)*3 + li);
        a0 = fma(__ldg(k), u, a0);
        a1 = fma(__ldg(k + 3), u, a1);
        a2 = fma(__ldg(k + 6), u, a2);
      }
    }
    if (lane4 == 3) { a0 = 0.0; a1 = 0.0; a2 = 0.0; }
    a0 += __shfl_xor_sync(0xffffffffu, a0, 1); a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
    a0 += __shfl_xor_sync(0xffffffffu, a0, 2); a1 += __shfl_xor_sync(0xffffffffu, a1, 2); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
    if (row < nNo && lane4 < 3) KU[size_t(row)*3 + lane4] = (lane4 == 0) ? a0 : (lane4 == 1) ? a1 : a2;
  }
}

// K2a  KU(i) = sum_j K(j) U(col_j)                     (fsils_spar_mul_ss, spar_mul.cpp:46-61)
// 4 lanes per row striding over the row's entries; fixed-order quad reduction.
__global__ void __launch_bounds__(256)
k_spmv_ss(const int* __restrict__ skip, int nNo, const int* __restrict__ rowPtr, const int* __restrict__ col,
          const double* __restrict__ K, const double* __restrict__ U, double* __restrict__ KU)
{
  if (skip && *skip) return;
  const int lane4 = threadIdx.x & 3;
  const int group = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int ngroups = (gridDim.x*blockDim.x) >> 2;
  const int nrounds = (nNo + ngroups - 1)/ngroups;
  for (int r = 0; r < nrounds; r++) {
    const int row = group + r*ngroups;
    double acc = 0.0;
    if (row < nNo) {
      const int s = __ldg(rowPtr + row);
      const int e = __ldg(rowPtr + row + 1);
      for (int p = s + lane4; p < e; p += 4) acc = acc + ld_stream(K + p)*__ldg(U + __ldg(col + p));
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (row < nNo && lane4 == 0) KU[row] = acc;
  }
}

// K2b  KU(m,i) = sum_j K(m,j) U(col_j)                 (fsils_spar_mul_sv, spar_mul.cpp:63-127)
template <int DOF>
__global__ void __launch_bounds__(256)
k_spmv_sv(const int* __restrict__ skip, int nNo, const int* __restrict__ rowPtr, const int* __restrict__ col,
          const double* __restrict__ K, const double* __restrict__ U, double* __restrict__ KU)
{
  if (skip && *skip) return;
  const int lane4 = threadIdx.x & 3;
  const int group = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int ngroups = (gridDim.x*blockDim.x) >> 2;
  for (int row = group; row < nNo; row += ngroups) {
    const int s = __ldg(rowPtr + row);
    const int e = __ldg(rowPtr + row + 1);
    if (lane4 < DOF) {
      double acc = 0.0;
#pragma unroll 4
      for (int p = s; p < e; p++) acc = acc + ld_stream(K + size_t(p)*DOF + lane4)*__ldg(U + __ldg(col + p));
      KU[size_t(row)*DOF + lane4] = acc;
    }
  }
}

// K2c  KU(i) = sum_j K(:,j) . U(:,col_j)                (fsils_spar_mul_vs, spar_mul.cpp:129-189)
template <int DOF>
__global__ void __launch_bounds__(256)
k_spmv_vs(const int* __restrict__ skip, int nNo, const int* __restrict__ rowPtr, const int* __restrict__ col,
          const double* __restrict__ K, const double* __restrict__ U, double* __restrict__ KU)
{
  if (skip && *skip) return;
  const int lane4 = threadIdx.x & 3;
  const int group = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int ngroups = (gridDim.x*blockDim.x) >> 2;
  const int nrounds = (nNo + ngroups - 1)/ngroups;
  for (int r = 0; r < nrounds; r++) {
    const int row = group + r*ngroups;
    double acc = 0.0;
    if (row < nNo) {
      const int s = __ldg(rowPtr + row);
      const int e = __ldg(rowPtr + row + 1);
      for (int p = s + lane4; p < e; p += 4) {
        const int c = __ldg(col + p);
        double t = 0.0;
#pragma unroll
        for (int m = 0; m < DOF; m++) t = t + ld_stream(K + size_t(p)*DOF + m)*__ldg(U + size_t(c)*DOF + m);
        acc = acc + t;
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (row < nNo && lane4 == 0) KU[row] = acc;
  }
}

// K2d/K2e  the pressure Schur operator of the NS solver, SP = L P - Gt (G P)  (cgrad.cpp:96-113), as
// two passes over 32-byte-friendly layouts instead of three SpMVs + an axpy:
//   pass 1  V4(i) = [ sum_j G(:,j) P(col_j) , P(i) ]        G(3,nnz); P gathered as scalars
//   pass 2  SP(i) = sum_j L(j) V4(3,col_j) - sum_j Gt(:,j).V4(0:2,col_j)   GtL(4,nnz) = [Gt, L]
// Pass 2 streams one aligned 256-bit matrix entry and gathers one aligned 256-bit (one-sector) vector
// entry per non-zero, like the dof-4 block SpMV.  Between the passes the caller halo-adds V4(0:2,:)
// and applies the resistance-face preconditioner to it (ld = 4).
// 4 lanes per row striding over the row's entries; fixed-order quad reduction.
__global__ void __launch_bounds__(256)
k_schur_gp(const int* __restrict__ skip, int nNo, const int* __restrict__ rowPtr, const int* __restrict__ col, const double* __restrict__ G,
           const double* __restrict__ P, const double* __restrict__ Pown, double* __restrict__ V4)
{
  if (skip && *skip) return;
  const int lane4 = threadIdx.x & 3;
  const int group = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int ngroups = (gridDim.x*blockDim.x) >> 2;
  const int nrounds = (nNo + ngroups - 1)/ngroups;
  for (int r = 0; r < nrounds; r++) {
    const int row = group + r*ngroups;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    if (row < nNo) {
      const int s = __ldg(rowPtr + row);
      const int e = __ldg(rowPtr + row + 1);
#pragma unroll 2
      for (int p = s + lane4; p < e; p += 4) {
        const double u = __ldg(P + __ldg(col + p));
        const double* g = G + size_t(p)*3;
        a0 = fma(__ldg(g), u, a0);
        a1 = fma(__ldg(g + 1), u, a1);
        a2 = fma(__ldg(g + 2), u, a2);
      }
    }
    a0 += __shfl_xor_sync(0xffffffffu, a0, 1); a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
    a0 += __shfl_xor_sync(0xffffffffu, a0, 2); a1 += __shfl_xor_sync(0xffffffffu, a1, 2); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
    if (row < nNo && lane4 == 0) {
      d4 o; o.x = a0; o.y = a1; o.z = a2; o.w = __ldg(Pown + row);
      st256(V4 + size_t(row)*4, o);
    }
  }
}

// Pass 1 on a component-wise (SoA) copy of G: the four lanes of a quad read four CONSECUTIVE doubles per component and instruction
// (one or two sectors) instead of four doubles 24 bytes apart (three or four sectors): ~2.7 instead of ~3.4 L1 sector requests per
// entry for the kernel that ncu shows at 96 % l1tex throughput.
__global__ void __launch_bounds__(256)
k_schur_gp_soa(const int* __restrict__ skip, int nNo, const int* __restrict__ rowPtr, const int* __restrict__ col, const double* __restrict__ Gs,
               size_t nnz, const double* __restrict__ P, const double* __restrict__ Pown, double* __restrict__ V4)
{
  if (skip && *skip) return;
  const int lane4 = threadIdx.x & 3;
  const int group = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int ngroups = (gridDim.x*blockDim.x) >> 2;
  const int nrounds = (nNo + ngroups - 1)/ngroups;
  const double* G0 = Gs; const double* G1 = Gs + nnz; const double* G2 = Gs + 2*nnz;
  for (int r = 0; r < nrounds; r++) {
    const int row = group + r*ngroups;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    if (row < nNo) {
      const int s = __ldg(rowPtr + row);
      const int e = __ldg(rowPtr + row + 1);
#pragma unroll 2
      for (int p = s + lane4; p < e; p += 4) {
        const double u = __ldg(P + __ldg(col + p));
        a0 = fma(__ldg(G0 + p), u, a0);
        a1 = fma(__ldg(G1 + p), u, a1);
        a2 = fma(__ldg(G2 + p), u, a2);
      }
    }
    a0 += __shfl_xor_sync(0xffffffffu, a0, 1); a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
    a0 += __shfl_xor_sync(0xffffffffu, a0, 2); a1 += __shfl_xor_sync(0xffffffffu, a1, 2); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
    if (row < nNo && lane4 == 0) {
      d4 o; o.x = a0; o.y = a1; o.z = a2; o.w = __ldg(Pown + row);
      st256(V4 + size_t(row)*4, o);
    }
  }
}

__global__ void __launch_bounds__(256)
k_schur_sp(const int* __restrict__ skip, int nNo, const int* __restrict__ rowPtr, const int* __restrict__ col, const double* __restrict__ GtL,
           const double* __restrict__ V4, double* __restrict__ SP)
{
  if (skip && *skip) return;
  const int lane4 = threadIdx.x & 3;
  const int group = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int ngroups = (gridDim.x*blockDim.x) >> 2;
  const int nrounds = (nNo + ngroups - 1)/ngroups;
  for (int r = 0; r < nrounds; r++) {
    const int row = group + r*ngroups;
    double aL = 0.0, aD = 0.0;
    if (row < nNo) {
      const int s = __ldg(rowPtr + row);
      const int e = __ldg(rowPtr + row + 1);
#pragma unroll 2
      for (int p = s + lane4; p < e; p += 4) {
        const int c = ld_stream_i(col + p);
        const d4 k = ld256_stream(GtL + size_t(p)*4);
        const d4 v = ld256_keep(V4 + size_t(c)*4);
        aL = fma(k.w, v.w, aL);
        aD = aD + (k.x*v.x + k.y*v.y + k.z*v.z);
      }
    }
    aL += __shfl_xor_sync(0xffffffffu, aL, 1); aD += __shfl_xor_sync(0xffffffffu, aD, 1);
    aL += __shfl_xor_sync(0xffffffffu, aL, 2); aD += __shfl_xor_sync(0xffffffffu, aD, 2);
    if (row < nNo && lane4 == 0) SP[row] = aL - aD;
  }
}

// Variants of the two Schur passes with four entries in flight per lane (rows of <= 16 entries take one
// trip): 4 x (index -> matrix entry + gathered vector entry) independent chains per thread instead of two.
__global__ void __launch_bounds__(256)
k_schur_gp4(const int* __restrict__ skip, int nNo, const int* __restrict__ rowPtr, const int* __restrict__ col, const double* __restrict__ G,
            const double* __restrict__ P, const double* __restrict__ Pown, double* __restrict__ V4)
{
  if (skip && *skip) return;
  const int lane4 = threadIdx.x & 3;
  const int group = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int ngroups = (gridDim.x*blockDim.x) >> 2;
  const int nrounds = (nNo + ngroups - 1)/ngroups;
  for (int r = 0; r < nrounds; r++) {
    const int row = group + r*ngroups;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    if (row < nNo) {
      const int s = __ldg(rowPtr + row);
      const int e = __ldg(rowPtr + row + 1);
      for (int base = s + lane4; base < e; base += 16) {
        int c[4];
        double u[4], g0[4], g1[4], g2[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { const int p = base + 4*q; c[q] = (p < e) ? ld_stream_i(col + p) : -1; }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int p = base + 4*q;
          const bool on = c[q] >= 0;
          const double* g = G + size_t(on ? p : s)*3;
          g0[q] = on ? ld_stream(g) : 0.0; g1[q] = on ? ld_stream(g + 1) : 0.0; g2[q] = on ? ld_stream(g + 2) : 0.0;
          u[q] = on ? __ldg(P + c[q]) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) { a0 = fma(g0[q], u[q], a0); a1 = fma(g1[q], u[q], a1); a2 = fma(g2[q], u[q], a2); }
      }
    }
    a0 += __shfl_xor_sync(0xffffffffu, a0, 1); a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
    a0 += __shfl_xor_sync(0xffffffffu, a0, 2); a1 += __shfl_xor_sync(0xffffffffu, a1, 2); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
    if (row < nNo && lane4 == 0) {
      d4 o; o.x = a0; o.y = a1; o.z = a2; o.w = __ldg(Pown + row);
      st256(V4 + size_t(row)*4, o);
    }
  }
}

__global__ void __launch_bounds__(256)
k_schur_sp4(const int* __restrict__ skip, int nNo, const int* __restrict__ rowPtr, const int* __restrict__ col, const double* __restrict__ GtL,
            const double* __restrict__ V4, double* __restrict__ SP)
{
  if (skip && *skip) return;
  const int lane4 = threadIdx.x & 3;
  const int group = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int ngroups = (gridDim.x*blockDim.x) >> 2;
  const int nrounds = (nNo + ngroups - 1)/ngroups;
  for (int r = 0; r < nrounds; r++) {
    const int row = group + r*ngroups;
    double aL = 0.0, aD = 0.0;
    if (row < nNo) {
      const int s = __ldg(rowPtr + row);
      const int e = __ldg(rowPtr + row + 1);
      for (int base = s + lane4; base < e; base += 16) {
        int c[4];
        d4 k[4], v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { const int p = base + 4*q; c[q] = (p < e) ? ld_stream_i(col + p) : -1; }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int p = base + 4*q;
          if (c[q] >= 0) { k[q] = ld256_stream(GtL + size_t(p)*4); v[q] = ld256_keep(V4 + size_t(c[q])*4); }
          else { k[q].x = k[q].y = k[q].z = k[q].w = 0.0; v[q] = k[q]; }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          aL = fma(k[q].w, v[q].w, aL);
          aD = aD + (k[q].x*v[q].x + k[q].y*v[q].y + k[q].z*v[q].z);
        }
      }
    }
    aL += __shfl_xor_sync(0xffffffffu, aL, 1); aD += __shfl_xor_sync(0xffffffffu, aD, 1);
    aL += __shfl_xor_sync(0xffffffffu, aL, 2); aD += __shfl_xor_sync(0xffffffffu, aD, 2);
    if (row < nNo && lane4 == 0) SP[row] = aL - aD;
  }
}

// =================================================================================================
// K3  multi-dot: red[slot0 + j] = sum_{idx < n} V_j[idx] * w[idx],  j = 0..cnt-1, V_j = base + j*stride.
// (fsils_nc_dot_v inside the Arnoldi loop, liner_solver/gmres.cpp:550-555; dot.cpp:134-175.)
// ONE launch serves every basis vector of an Arnoldi step: each CTA owns a contiguous chunk of the
// index range and walks the vectors in batches of kDotJB, so every V_j is streamed from HBM exactly
// once and the CTA's chunk of w (<= 70 KB at P10) is re-read from L1/L2, not from HBM.
// Deterministic: fixed grid and chunking, fixed-shape tree inside the CTA, per-CTA partials added in
// CTA order by the last CTA to finish.  partial must hold gridDim.x * cnt doubles.
// =================================================================================================
__global__ void __launch_bounds__(kRedThreads)
k_multi_dot(const int* __restrict__ skip, size_t n, const double* __restrict__ base, size_t stride, const double* __restrict__ w, int cnt,
            double* __restrict__ partial, unsigned int* __restrict__ counter, double* __restrict__ red, int slot0,
            PeerRedArgs pa = PeerRedArgs(), PeerState* ps = nullptr,      // ps != null: the last CTA also all-reduces red[slot0 .. slot0+cnt)
            int defer_tail = 0)        // 1: only write the per-CTA partials, TRANSPOSED (partial[j*gridDim + cta]); k_multi_dot_final adds them
{
  if (skip && *skip) return;
  __shared__ double sm[kRedThreads/32][kDotJB];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t chunk = ((n + gridDim.x - 1)/gridDim.x + 3) & ~size_t(3);
  const size_t beg = size_t(blockIdx.x)*chunk;
  const size_t end = (beg + chunk < n) ? beg + chunk : n;

  for (int j0 = 0; j0 < cnt; j0 += kDotJB) {
    const int m = (cnt - j0 < kDotJB) ? cnt - j0 : kDotJB;
    double acc[kDotJB];
#pragma unroll
    for (int j = 0; j < kDotJB; j++) acc[j] = 0.0;
    const double* vb = base + size_t(j0)*stride;
    if (m == kDotJB) {
      for (size_t idx = beg + threadIdx.x; idx < end; idx += kRedThreads) {
        const double wv = w[idx];
        double v[kDotJB];
#pragma unroll
        for (int j = 0; j < kDotJB; j++) v[j] = ld_stream(vb + size_t(j)*stride + idx);
#pragma unroll
        for (int j = 0; j < kDotJB; j++) acc[j] = fma(v[j], wv, acc[j]);
      }
    } else {
      for (size_t idx = beg + threadIdx.x; idx < end; idx += kRedThreads) {
        const double wv = w[idx];
#pragma unroll
        for (int j = 0; j < kDotJB; j++)
          if (j < m) acc[j] = fma(ld_stream(vb + size_t(j)*stride + idx), wv, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < kDotJB; j++) {
      double v = acc[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) sm[wid][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < m) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < kRedThreads/32; k++) v += sm[k][threadIdx.x];
      if (defer_tail) partial[size_t(j0 + threadIdx.x)*gridDim.x + blockIdx.x] = v;
      else partial[size_t(blockIdx.x)*cnt + j0 + threadIdx.x] = v;
    }
    __syncthreads();
  }
  if (defer_tail) return;
  __threadfence();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(counter, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    // warp `wid` sums partial[:, j] for j = wid, wid+8, ...: lane-strided serial sums in CTA order,
    // then a fixed shuffle tree
    for (int j = wid; j < cnt; j += kRedThreads/32) {
      double v = 0.0;
      for (unsigned int b = lane; b < gridDim.x; b += 32) v += __ldcg(partial + size_t(b)*cnt + j);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) red[slot0 + j] = v;
    }
    if (threadIdx.x == 0) *counter = 0u;
    if (ps) peer_allreduce_body<0>(pa, ps, red + slot0, cnt);        // (starts with a __syncthreads: red[] is complete)
  }
}

// Second stage of k_multi_dot for many dots: with one CTA doing the ordered sum of G partials for cnt columns (the in-kernel tail)
// the tail costs ~0.3 us per column - 30+ us at a basis depth of 100, more than the first stage itself once the vectors are split
// over 8 GPUs.  Here one WARP per column, spread over ceil(cnt/8) CTAs, reads the transposed partials with coalesced loads and adds
// them in exactly the order of the in-kernel tail (lane-strided serial sums in CTA order, fixed shuffle tree): bit-identical results.
// The last CTA to finish runs the all-reduce epilogue.
__global__ void __launch_bounds__(kRedThreads)
k_multi_dot_final(const int* __restrict__ skip, int G, int cnt, const double* __restrict__ partialT, unsigned int* __restrict__ counter,
                  double* __restrict__ red, int slot0, PeerRedArgs pa = PeerRedArgs(), PeerState* ps = nullptr)
{
  if (skip && *skip) return;
  __shared__ bool last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int j = blockIdx.x*(kRedThreads/32) + wid;
  if (j < cnt) {
    const double* col = partialT + size_t(j)*G;
    double v = 0.0;
    int b = lane;
    for (; b + 96 < G; b += 128) {
      const double p0 = __ldcg(col + b), p1 = __ldcg(col + b + 32), p2 = __ldcg(col + b + 64), p3 = __ldcg(col + b + 96);
      v += p0; v += p1; v += p2; v += p3;
    }
    for (; b < G; b += 32) v += __ldcg(col + b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[slot0 + j] = v;
  }
  if (!ps) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *counter = 0u;
    __threadfence();
    peer_allreduce_body<0>(pa, ps, red + slot0, cnt);
  }
}

// ---- device-resident Arnoldi bookkeeping (Hessenberg::finish_column of krylov.hpp, liner_solver/gmres.cpp:556-612) -------------
// One thread: column i of the Hessenberg matrix from the reduced dots, the Pythagorean norm, the Givens rotations and the residual
// estimate; sets `done` when |err(i+1)| < eps.  Every product and sum is rounded separately (__dmul_rn / __dadd_rn: no FMA
// contraction), so the numbers are bit-identical to the host version the serial test policy runs against the compiled reference.
// `done` must stay the FIRST int after the doubles: the skip pointer of the heavy kernels points at it.
struct GmresState { double eps, err0; int done, suc, last_i, pad; };
constexpr int kGivensMax = 1024;          // Krylov dimensions up to this keep the column in shared memory
__device__ __forceinline__ void gmres_givens_body(GmresState* st, int i, int sD, const double* __restrict__ red, double* __restrict__ h,
                                                  double* __restrict__ c, double* __restrict__ s, double* __restrict__ err)
{
  if (st->done) return;
  // the column and the rotations are staged in shared memory by the whole CTA (coalesced), the inherently sequential recurrences
  // run on one thread out of shared memory (a single thread walking global memory costs ~1 us per dependent access)
  __shared__ double sc[kGivensMax + 2], cc[kGivensMax], ss[kGivensMax];
  for (int j = threadIdx.x; j <= i + 1; j += blockDim.x) sc[j] = red[j];
  for (int j = threadIdx.x; j < i; j += blockDim.x) { cc[j] = c[j]; ss[j] = s[j]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    // the recurrences themselves: krylov.hpp's givens_finish_column, the function the host loop runs (one source for both)
    double e_i = (i == 0) ? st->err0 : err[i], e_i1 = 0.0;
    const double e1 = givens_finish_column(sc, cc, ss, e_i, e_i1, i);
    c[i] = cc[i];
    s[i] = ss[i];
    err[i] = e_i;
    err[i+1] = e_i1;
    st->last_i = i;
    if (e1 < st->eps) { st->done = 1; st->suc = 1; }
  }
  __syncthreads();
  double* col = h + size_t(i)*(sD + 1);
  for (int j = threadIdx.x; j <= i + 1; j += blockDim.x) col[j] = sc[j];
}
__global__ void __launch_bounds__(256)
k_gmres_givens(GmresState* st, int i, int sD, const double* __restrict__ red, double* __restrict__ h,
               double* __restrict__ c, double* __restrict__ s, double* __restrict__ err)
{
  gmres_givens_body(st, i, sD, red, h, c, s, err);
}


// K4  classical Gram-Schmidt update + normalisation in one pass
//   w <- (w - sum_{j<k} h_j u_j) * 1/sqrt|h_k - sum_j h_j^2|,  h = red[slot0 ..]  (already reduced)
// (omp_sum_v / omp_mul_v calls at liner_solver/gmres.cpp:561-569; same left-to-right order.)
__global__ void __launch_bounds__(256)
k_cgs_update_scale(size_t n, int k, const double* __restrict__ base, size_t stride, double* __restrict__ w,
                   const double* __restrict__ red, int slot0, const int* __restrict__ skip = nullptr)
{
  if (skip && *skip) return;
  // (the Givens bookkeeping of the device-resident Arnoldi loop once rode on this kernel as an extra CTA: its 24 KB of static shared
  // memory cost the streaming update 50 % of its bandwidth - it is a kernel of its own again, k_gmres_givens)
  extern __shared__ double hs[];     // k+1 coefficients
  for (int j = threadIdx.x; j <= k; j += blockDim.x) hs[j] = red[slot0 + j];
  __syncthreads();
  __shared__ double inv;
  if (threadIdx.x == 0) {
    double hh = hs[k];
    for (int j = 0; j < k; j++) hh = __dsub_rn(hh, __dmul_rn(hs[j], hs[j]));
    inv = 1.0 / sqrt(fabs(hh));
  }
  __syncthreads();
  const double sc = inv;
  const size_t tid = size_t(blockIdx.x)*blockDim.x + threadIdx.x;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t idx = tid; idx < n; idx += nth) {
    double v = w[idx];
    int j = 0;
    // 8 independent streaming loads in flight, then the updates in the reference's order (j ascending)
    for (; j + 8 <= k; j += 8) {
      double b[8];
#pragma unroll
      for (int u = 0; u < 8; u++) b[u] = ld_stream(base + size_t(j + u)*stride + idx);
#pragma unroll
      for (int u = 0; u < 8; u++) v = fma(-hs[j + u], b[u], v);
    }
    for (; j < k; j++) v = fma(-hs[j], ld_stream(base + size_t(j)*stride + idx), v);
    w[idx] = sc*v;
  }
}

// out = base + sum_j coef[j] * V[j]   (sequential in j; base may be null = 0; out may alias base)
struct CombArgs { double coef[kMaxComb]; };
__global__ void __launch_bounds__(256)
k_lin_comb(size_t n, double* __restrict__ out, const double* base, int k, const double* __restrict__ V, size_t stride, CombArgs a)
{
  const size_t tid = size_t(blockIdx.x)*blockDim.x + threadIdx.x;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t idx = tid; idx < n; idx += nth) {
    double v = base ? base[idx] : 0.0;
    for (int j = 0; j < k; j++) v = fma(a.coef[j], V[size_t(j)*stride + idx], v);
    out[idx] = v;
  }
}

// ---- BLAS-1 (omp_la.cpp:39-148) ------------------------------------------------------------------
__global__ void k_axpy(size_t n, double a, const double* __restrict__ x, double* __restrict__ y)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) y[i] = fma(a, x[i], y[i]);
}
__global__ void k_scal(size_t n, double a, double* __restrict__ x)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) x[i] = a*x[i];
}
__global__ void k_divs(size_t n, double d, double* __restrict__ x)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) x[i] = x[i]/d;
}
__global__ void k_fill(size_t n, double a, double* __restrict__ x)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) x[i] = a;
}
__global__ void k_sub(size_t n, const double* a, const double* b, double* out)     // out = a - b (may alias)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) out[i] = a[i] - b[i];
}
__global__ void k_mul(size_t n, const double* __restrict__ w, double* __restrict__ x)   // x = w (.) x
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) x[i] = w[i]*x[i];
}
// out = a*x + b*y    (S = R - alpha V, R = S - omega T of bicgs.cpp:96,103)
__global__ void k_lin2(size_t n, double* out, double a, const double* x, double b, const double* y)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) out[i] = fma(b, y[i], a*x[i]);
}
// X = X + a P + b S   (bicgs.cpp:102)
__global__ void k_axpy2(size_t n, double* __restrict__ X, double a, const double* __restrict__ P, double b, const double* __restrict__ S)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) X[i] = fma(b, S[i], fma(a, P[i], X[i]));
}
// P = R + beta (P - omega V)   (bicgs.cpp:113)
__global__ void k_bicg_p(size_t n, double* __restrict__ P, const double* __restrict__ R, const double* __restrict__ V, double beta, double omega)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) P[i] = fma(beta, fma(-omega, V[i], P[i]), R[i]);
}

// ---- device-resident CG iteration of the Schur / vector CG solvers (cgrad.cpp:50-160, 166-246) ----------
// The scalars of the iteration (err, errO, alpha) live on the device, so that the host can enqueue
// iterations back to back and only polls the state every few iterations; once `done` is set every
// kernel of the remaining (already enqueued) iterations returns at once.
struct CgState { double err, errO, eps; int done, suc, last_i, pad; };



// top of iteration i:  last_i = i; if (err < eps) { suc; break; }  errO = err;
__global__ void k_cg_head(CgState* st, int i)
{
  if (st->done) return;
  st->last_i = i;
  if (st->err < st->eps) { st->done = 1; st->suc = 1; }
  else st->errO = st->err;
}
// alpha = errO / <P,SP>;  X += alpha P;  R -= alpha SP;  red[slot] = sum_{idx < nOwn} R^2  (local part)
// Same two-stage deterministic reduction as k_multi_dot.
__global__ void __launch_bounds__(kRedThreads)
k_cg_update(size_t n, size_t nOwn, const CgState* __restrict__ st, const double* __restrict__ red_dot,
            const double* __restrict__ P, const double* __restrict__ SP, double* __restrict__ X, double* __restrict__ R,
            double* __restrict__ partial, unsigned int* __restrict__ counter, double* __restrict__ red_out,
            PeerRedArgs pa = PeerRedArgs(), PeerState* ps = nullptr)      // ps != null: the last CTA also all-reduces red_out[0]
{
  if (st->done) return;
  __shared__ double sm[kRedThreads/32];
  __shared__ bool last;
  const double alpha = st->errO / red_dot[0];
  const size_t chunk = ((n + gridDim.x - 1)/gridDim.x + 3) & ~size_t(3);
  const size_t beg = size_t(blockIdx.x)*chunk;
  const size_t end = (beg + chunk < n) ? beg + chunk : n;
  double acc = 0.0;
  for (size_t idx = beg + threadIdx.x; idx < end; idx += kRedThreads) {
    X[idx] = fma(alpha, P[idx], X[idx]);
    const double r = fma(-alpha, SP[idx], R[idx]);
    R[idx] = r;
    if (idx < nOwn) acc = fma(r, r, acc);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) sm[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < kRedThreads/32; k++) v += sm[k];
    partial[blockIdx.x] = v;
    __threadfence();
    last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && wid == 0) {
    __threadfence();
    double v = 0.0;
    for (unsigned int b = lane; b < gridDim.x; b += 32) v += __ldcg(partial + b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) { red_out[0] = v; *counter = 0u; }
  }
  if (last && ps) peer_allreduce_body<0>(pa, ps, red_out, 1);
}
// err = (sqrt(red))^2;  P = (err/errO) * (P + (errO/err) R);  state.err = err   (cgrad.cpp:222-227)
__global__ void __launch_bounds__(256)
k_cg_pupdate(size_t n, CgState* st, const double* __restrict__ red_rr, const double* __restrict__ R, double* __restrict__ P)
{
  if (st->done) return;
  const double errO = st->errO;
  double err = sqrt(red_rr[0]);
  err = err*err;
  const double a = errO/err, b = err/errO;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) P[i] = b*fma(a, R[i], P[i]);
  // every CTA has read st->errO / red_rr before this store can matter: the only reader of st->err is
  // the next iteration's k_cg_head, ordered after this kernel by the stream
  if (blockIdx.x == 0 && threadIdx.x == 0) st->err = err;
}

// ---- permutation by lhs.map (solve.cpp:116-120,188-192) ------------------------------------------
// out(:, map[a]) = in(:, a)
__global__ void k_permute_fwd(int nNo, int dof, const int* __restrict__ map, const double* __restrict__ in, double* __restrict__ out)
{
  const size_t n = size_t(nNo)*dof;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) {
    const int a = int(i / dof), l = int(i % dof);
    out[size_t(map[a])*dof + l] = in[i];
  }
}
// out(:, a) = in(:, map[a])
__global__ void k_permute_bwd(int nNo, int dof, const int* __restrict__ map, const double* __restrict__ in, double* __restrict__ out)
{
  const size_t n = size_t(nNo)*dof;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) {
    const int a = int(i / dof), l = int(i % dof);
    out[i] = in[size_t(map[a])*dof + l];
  }
}
// Val rows between assembly layout (rows in assembly order) and solver layout (rows in solver order):
// row a of the assembly CSR [rowPtrA[a], rowPtrA[a+1]) <-> row map[a] of the solver CSR.
__global__ void k_val_rows(int nNo, int bs, const int* __restrict__ map, const int* __restrict__ rowPtrA,
                           const int* __restrict__ rowPtrS, const double* __restrict__ src, double* __restrict__ dst, int to_solver)
{
  // one warp per row
  const int warp = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x*blockDim.x) >> 5;
  for (int a = warp; a < nNo; a += nwarps) {
    const int sa = rowPtrA[a], len = rowPtrA[a+1] - sa;
    const int ss = rowPtrS[map[a]];
    const size_t cnt = size_t(len)*bs;
    const double* s = to_solver ? src + size_t(sa)*bs : src + size_t(ss)*bs;
    double* d = to_solver ? dst + size_t(ss)*bs : dst + size_t(sa)*bs;
    for (size_t i = lane; i < cnt; i += 32) d[i] = s[i];
  }
}

// ---- K6 Jacobi (diagonal) preconditioner (liner_solver/precond.cpp:122-256) -----------------------
__global__ void k_diag_extract(int nNo, int dof, const int* __restrict__ diagPtr, const double* __restrict__ Val, double* __restrict__ W)
{
  const size_t n = size_t(nNo)*dof;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) {
    const int a = int(i / dof), l = int(i % dof);
    W[i] = Val[size_t(diagPtr[a])*dof*dof + l*dof + l];
  }
}
__global__ void k_w_invsqrt(size_t n, double* __restrict__ W)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) {
    double w = W[i];
    if (w == 0.0) w = 1.0;
    W[i] = 1.0 / sqrt(fabs(w));
  }
}
// W(i, glob[a]) *= val(i,a), i < m   (Dirichlet masking, precond.cpp:204-224)
__global__ void k_face_mask(int fnNo, int m, int fdof, int dof, const int* __restrict__ glob, const double* __restrict__ val, double* __restrict__ W)
{
  const int n = fnNo*m;
  for (int t = blockIdx.x*blockDim.x + threadIdx.x; t < n; t += gridDim.x*blockDim.x) {
    const int a = t / m, i = t % m;
    W[size_t(glob[a])*dof + i] *= val[size_t(a)*fdof + i];
  }
}
// valM(i,a) = val(i,a) * W(i, glob[a])   (precond.cpp:245-255)
__global__ void k_face_valM(int fnNo, int m, int fdof, int dof, const int* __restrict__ glob, const double* __restrict__ val,
                            const double* __restrict__ W, double* __restrict__ valM)
{
  const int n = fnNo*m;
  for (int t = blockIdx.x*blockDim.x + threadIdx.x; t < n; t += gridDim.x*blockDim.x) {
    const int a = t / m, i = t % m;
    valM[size_t(a)*fdof + i] = val[size_t(a)*fdof + i] * W[size_t(glob[a])*dof + i];
  }
}
// Val <- (W_row Val) W_col in ONE read-modify-write pass (the reference makes two: pre_mul :549 then
// pos_mul :46); each entry is rounded as (v*Wr)*Wc like the reference.  4 lanes per row, lane i owns
// block-row i (dof 4: one 256-bit load + store).
template <int DOF>
__global__ void __launch_bounds__(256)
k_scale_val(int nNo, const int* __restrict__ rowPtr, const int* __restrict__ col, const double* __restrict__ Wr,
            const double* __restrict__ Wc, double* __restrict__ Val)
{
  const int lane4 = threadIdx.x & 3;
  const int group = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int ngroups = (gridDim.x*blockDim.x) >> 2;
  for (int row = group; row < nNo; row += ngroups) {
    const int s = rowPtr[row], e = rowPtr[row+1];
    if (lane4 < DOF) {
      const double wr = Wr[size_t(row)*DOF + lane4];
      for (int p = s; p < e; p++) {
        const int c = col[p];
        double* v = Val + (size_t(p)*DOF*DOF + lane4*DOF);
        if constexpr (DOF == 4) {
          d4 k = ld256(v);
          const d4 wc = ld256(Wc + size_t(c)*4);
          k.x = (k.x*wr)*wc.x; k.y = (k.y*wr)*wc.y; k.z = (k.z*wr)*wc.z; k.w = (k.w*wr)*wc.w;
          st256(v, k);
        } else {
#pragma unroll
          for (int j = 0; j < DOF; j++) v[j] = (v[j]*wr)*Wc[size_t(c)*DOF + j];
        }
      }
    }
  }
}

// ---- K7 row-and-column-scaling preconditioner (liner_solver/precond.cpp:266-540) ------------------
// Wr <- 1 where Wr > 0.5, 0 elsewhere, with the reference's arithmetic (:305-307): after the overlap
// add a Dirichlet mask shared by several ranks can exceed 1.
__global__ void k_rcs_renorm(size_t n, double* __restrict__ W)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) {
    double w = W[i] - 0.5;
    w = w / fabs(w);
    W[i] = (w + fabs(w)) * 0.5;
  }
}
// unit diagonal on the killed rows: Val(ii,d) = Wr(i)*(Val(ii,d) - 1) + 1   (:321-366)
__global__ void k_rcs_unit_diag(int nNo, int dof, const int* __restrict__ diagPtr, const double* __restrict__ Wr, double* __restrict__ Val)
{
  const size_t n = size_t(nNo)*dof;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t t = size_t(blockIdx.x)*blockDim.x + threadIdx.x; t < n; t += nth) {
    const int a = int(t / dof), i = int(t % dof);
    double* v = Val + (size_t(diagPtr[a])*dof*dof + i*dof + i);
    *v = Wr[t]*(*v - 1.0) + 1.0;
  }
}
// max is exact and order-independent, so an atomic max on the bit pattern of a non-negative double
// is deterministic
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v)
{
  atomicMax(reinterpret_cast<unsigned long long*>(addr), static_cast<unsigned long long>(__double_as_longlong(v)));
}
// One sweep's norms (:384-510): Wr(i,row) = max_j,p |Val(i,j,p)| over the row's blocks,
// Wc(j,col_p) = max_i |Val(i,j,p)|.  Quad per row, lane i owns block-row i; the column maxima of a
// block are first reduced across the quad (lane j ends up with column j), then ONE atomic per lane.
// Wc must be zero on entry.
template <int DOF>
__global__ void __launch_bounds__(256)
k_rcs_norms(int nNo, const int* __restrict__ rowPtr, const int* __restrict__ col, const double* __restrict__ Val,
            double* __restrict__ Wr, double* __restrict__ Wc)
{
  const int lane4 = threadIdx.x & 3;
  const int group = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int ngroups = (gridDim.x*blockDim.x) >> 2;
  const int nrounds = (nNo + ngroups - 1)/ngroups;
  for (int r = 0; r < nrounds; r++) {
    const int row = group + r*ngroups;
    const bool live = row < nNo;
    const int s = live ? rowPtr[row] : 0, e = live ? rowPtr[row+1] : 0;
    // all four lanes of a quad walk the same row, so the shuffles below are uniform inside the quad;
    // quads of one warp may have different lengths -> pad to the warp's longest row
    int len = e - s;
    int wl = len;
#pragma unroll
    for (int o = 16; o >= 4; o >>= 1) wl = max(wl, __shfl_xor_sync(0xffffffffu, wl, o));
    double rmax = 0.0;
    for (int k = 0; k < wl; k++) {
      const int p = s + k;
      const bool on = k < len;
      double v[4] = {0.0, 0.0, 0.0, 0.0};
      if (on && lane4 < DOF) {
#pragma unroll
        for (int j = 0; j < DOF; j++) v[j] = fabs(Val[size_t(p)*DOF*DOF + lane4*DOF + j]);
      }
#pragma unroll
      for (int j = 0; j < DOF; j++) rmax = fmax(rmax, v[j]);
      // column maxima across the quad
      double mine = 0.0;
#pragma unroll
      for (int j = 0; j < DOF; j++) {
        double m = v[j];
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
        if (lane4 == j) mine = m;
      }
      if (on && lane4 < DOF) atomic_max_nonneg(Wc + size_t(col[p])*DOF + lane4, mine);
    }
    if (live && lane4 < DOF) Wr[size_t(row)*DOF + lane4] = rmax;
  }
}
// out[0] = max(out[0], max_i |1 - W[i]|)  (:514); out must be zero on entry
__global__ void __launch_bounds__(256)
k_rcs_dev1(size_t n, const double* __restrict__ W, double* __restrict__ out)
{
  double m = 0.0;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) m = fmax(m, fabs(1.0 - W[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomic_max_nonneg(out, m);
}
// W = 1/sqrt(W); Wacc *= W   (:518-527)
__global__ void k_rcs_invsqrt_accum(size_t n, double* __restrict__ W, double* __restrict__ Wacc)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) {
    const double w = 1.0 / sqrt(W[i]);
    W[i] = w;
    Wacc[i] = Wacc[i] * w;
  }
}

// ---- K8 depart (liner_solver/ns_solver.cpp:91-163), nsd = 3 --------------------------------------
// Splits the scaled 4x4 blocks into mK(9), mG(3), mD(3), mL(1) and builds Gt(:,l) = -mG(:,tpos[l])
// with the transpose position precomputed once (the reference searches the row each time).
__global__ void __launch_bounds__(256)
k_depart3(size_t nnz, const int* __restrict__ tpos, const double* __restrict__ Val, double* __restrict__ Gt,
          double* __restrict__ mK, double* __restrict__ mG, double* __restrict__ mD, double* __restrict__ mL,
          double* __restrict__ GtL, double* __restrict__ Gs = nullptr)       // Gs: component-wise copy of mG, Gs[m*nnz + p]
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t p = size_t(blockIdx.x)*blockDim.x + threadIdx.x; p < nnz; p += nth) {
    const double* v = Val + p*16;
    const d4 r0 = ld256(v), r1 = ld256(v + 4), r2 = ld256(v + 8), r3 = ld256(v + 12);
    double* k = mK + p*9;
    k[0] = r0.x; k[1] = r0.y; k[2] = r0.z;
    k[3] = r1.x; k[4] = r1.y; k[5] = r1.z;
    k[6] = r2.x; k[7] = r2.y; k[8] = r2.z;
    double* g = mG + p*3;
    g[0] = r0.w; g[1] = r1.w; g[2] = r2.w;
    if (Gs) { Gs[p] = r0.w; Gs[nnz + p] = r1.w; Gs[2*nnz + p] = r2.w; }
    double* d = mD + p*3;
    d[0] = r3.x; d[1] = r3.y; d[2] = r3.z;
    mL[p] = r3.w;
    const size_t t = size_t(tpos[p]);
    const double* vt = Val + t*16;
    double* gt = Gt + p*3;
    const double g0 = -vt[3], g1 = -vt[7], g2 = -vt[11];
    gt[0] = g0; gt[1] = g1; gt[2] = g2;
    if (GtL) { d4 o; o.x = g0; o.y = g1; o.z = g2; o.w = r3.w; st256(GtL + p*4, o); }   // packed [Gt, L] for k_schur_sp
  }
}
template <int NSD>
__global__ void k_depart_generic(size_t nnz, const int* __restrict__ tpos, const double* __restrict__ Val, double* __restrict__ Gt,
                                 double* __restrict__ mK, double* __restrict__ mG, double* __restrict__ mD, double* __restrict__ mL)
{
  constexpr int D = NSD + 1;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t p = size_t(blockIdx.x)*blockDim.x + threadIdx.x; p < nnz; p += nth) {
    const double* v = Val + p*D*D;
    for (int i = 0; i < NSD; i++) {
      for (int j = 0; j < NSD; j++) mK[p*NSD*NSD + i*NSD + j] = v[i*D + j];
      mG[p*NSD + i] = v[i*D + NSD];
      mD[p*NSD + i] = v[NSD*D + i];
    }
    mL[p] = v[D*D - 1];
    const double* vt = Val + size_t(tpos[p])*D*D;
    for (int i = 0; i < NSD; i++) Gt[p*NSD + i] = -vt[i*D + NSD];
  }
}

// Ri(dof,nNo) <-> Rm(nsd,nNo), Rc(nNo)   (ns_solver.cpp:201-213, 481-488)
__global__ void k_split_mc(int nNo, int dof, const double* __restrict__ Ri, double* __restrict__ Rm, double* __restrict__ Rc)
{
  const size_t n = size_t(nNo)*dof;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) {
    const size_t a = i / dof; const int l = int(i % dof);
    if (l < dof-1) Rm[a*(dof-1) + l] = Ri[i]; else Rc[a] = Ri[i];
  }
}
__global__ void k_join_mc(int nNo, int dof, const double* __restrict__ Rm, const double* __restrict__ Rc, double* __restrict__ Ri)
{
  const size_t n = size_t(nNo)*dof;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += nth) {
    const size_t a = i / dof; const int l = int(i % dof);
    Ri[i] = (l < dof-1) ? Rm[a*(dof-1) + l] : Rc[a];
  }
}

// ---- K9 resistance-face rank-1 update (liner_solver/add_bc_mul.cpp:53-121) ------------------------
// stage 1: out[0] = sum_{a, i<m} valM(i,a) * X(i, glob[a])  over face nodes with glob[a] < lim
// (lim = nNo for a face owned by one rank, mynNo for a shared face whose dot is completed by an
// all-reduce).  Up to kFaceBlocks CTAs over contiguous chunks of the face, fixed-shape tree inside a
// CTA, partials added in CTA order by the last CTA -> deterministic.  ld = leading dimension of X.
constexpr int kFaceBlocks = 64;
__global__ void __launch_bounds__(256)
k_face_dot(int fnNo, int m, int fdof, int ld, int lim, const int* __restrict__ glob, const double* __restrict__ valM,
           const double* X, double* __restrict__ partial, unsigned int* __restrict__ counter, double* __restrict__ out)
{
  const int n = fnNo*m;
  const int chunk = (n + gridDim.x - 1)/gridDim.x;
  const int beg = blockIdx.x*chunk;
  const int end = min(n, beg + chunk);
  double acc = 0.0;
  for (int t = beg + threadIdx.x; t < end; t += blockDim.x) {
    const int a = t / m, i = t % m;
    const int Ac = glob[a];
    if (Ac < lim) {
      const double vm = valM[size_t(a)*fdof + i];
      const double xv = X ? X[size_t(Ac)*ld + i] : vm;
      acc = fma(vm, xv, acc);
    }
  }
  __shared__ double sm[8];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) sm[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int k = 0; k < 8; k++) v += sm[k];
    if (gridDim.x == 1) { out[0] = v; last = false; }
    else {
      partial[blockIdx.x] = v;
      __threadfence();
      last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double v = 0.0;
    for (unsigned int b = 0; b < gridDim.x; b++) v += __ldcg(partial + b);
    out[0] = v;
    *counter = 0u;
  }
}
// stage 2: Y(i, glob[a]) += valM(i,a) * (coef * S)
__global__ void k_face_axpy(int fnNo, int m, int fdof, int dof, const int* __restrict__ glob, const double* __restrict__ valM,
                            double coef, const double* __restrict__ S, double* Y)
{
  const double s = coef * S[0];
  const int n = fnNo*m;
  for (int t = blockIdx.x*blockDim.x + threadIdx.x; t < n; t += gridDim.x*blockDim.x) {
    const int a = t / m, i = t % m;
    Y[size_t(glob[a])*dof + i] += valM[size_t(a)*fdof + i]*s;
  }
}

// both stages in one MULTI-CTA launch for a face that lives on one rank: the chunked partial sums of k_face_dot, a grid barrier (at
// most kFaceBlocks CTAs: always resident), every CTA adds the partials in CTA order (the same S as k_face_dot's last CTA, bit for
// bit) and updates its share of Y.  X may alias Y: nobody writes before everybody has passed the barrier.
// counter[0]: arrivals, counter[1]: finished CTAs (the last one resets both).
__global__ void __launch_bounds__(256)
k_face_rank1_grid(int fnNo, int m, int fdof, int ld, int lim, const int* __restrict__ glob, const double* __restrict__ valM,
                  const double* X, double coef, double* Y, double* __restrict__ partial, unsigned int* __restrict__ counter)
{
  const int n = fnNo*m;
  const int chunk = (n + gridDim.x - 1)/gridDim.x;
  const int beg = blockIdx.x*chunk;
  const int end = min(n, beg + chunk);
  double acc = 0.0;
  for (int t = beg + threadIdx.x; t < end; t += blockDim.x) {
    const int a = t / m, i = t % m;
    const int Ac = glob[a];
    if (Ac < lim) acc = fma(valM[size_t(a)*fdof + i], X[size_t(Ac)*ld + i], acc);
  }
  __shared__ double sm[8];
  __shared__ double S;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) sm[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int k = 0; k < 8; k++) v += sm[k];
    partial[blockIdx.x] = v;
    __threadfence();
    atomicAdd(&counter[0], 1u);
    { const long long t0 = clock64(); while (ld_acquire_gpu(&counter[0]) < gridDim.x) { if (clock64() - t0 > kSpinLimit) break; } }
    double sum = 0.0;
    for (unsigned int b = 0; b < gridDim.x; b++) sum += __ldcg(partial + b);
    S = coef*sum;
  }
  __syncthreads();
  const double s = S;
  for (int t = beg + threadIdx.x; t < end; t += blockDim.x) {
    const int a = t / m, i = t % m;
    Y[size_t(glob[a])*ld + i] += valM[size_t(a)*fdof + i]*s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(&counter[1], 1u) == gridDim.x - 1) { counter[0] = 0u; counter[1] = 0u; }
  }
}

// both stages in ONE single-CTA launch for a face that lives on one rank (no all-reduce between the stages): S = v^T X over the
// face, then Y += coef S v.  X may alias Y: every thread finishes reading before anyone writes.  Fixed-shape reduction
// (thread-strided serial sums, shuffle tree, warp sums in order) -> deterministic.
__global__ void __launch_bounds__(1024)
k_face_rank1(int fnNo, int m, int fdof, int ld, int lim, const int* __restrict__ glob, const double* __restrict__ valM,
             const double* X, double coef, double* Y)
{
  const int n = fnNo*m;
  double acc = 0.0;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const int a = t / m, i = t % m;
    const int Ac = glob[a];
    if (Ac < lim) acc = fma(valM[size_t(a)*fdof + i], X[size_t(Ac)*ld + i], acc);
  }
  __shared__ double sm[32];
  __shared__ double S;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) sm[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int k = 0; k < int(blockDim.x >> 5); k++) v += sm[k];
    S = coef*v;
  }
  __syncthreads();
  const double s = S;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const int a = t / m, i = t % m;
    Y[size_t(glob[a])*ld + i] += valM[size_t(a)*fdof + i]*s;
  }
}

// ---- halo exchange (fsils_commuv/commus, liner_solver/in_commu.cpp:49-170) ------------------------
// ld = leading dimension of V (ld >= dof; V4 of the Schur operator carries 3 components with ld 4)
__global__ void k_halo_pack(int n, int dof, int ld, const int* __restrict__ ptr, const double* __restrict__ V, double* __restrict__ buf)
{
  const int tot = n*dof;
  for (int t = blockIdx.x*blockDim.x + threadIdx.x; t < tot; t += gridDim.x*blockDim.x) {
    const int j = t / dof, l = t % dof;
    buf[t] = V[size_t(ptr[j])*ld + l];
  }
}
__global__ void k_halo_add(int n, int dof, int ld, const int* __restrict__ ptr, const double* __restrict__ buf, double* __restrict__ V)
{
  const int tot = n*dof;
  for (int t = blockIdx.x*blockDim.x + threadIdx.x; t < tot; t += gridDim.x*blockDim.x) {
    const int j = t / dof, l = t % dof;
    V[size_t(ptr[j])*ld + l] += buf[t];
  }
}

// ---- FP64 pipe peak: the roofline denominator of the element kernels (BASELINE.md par. 2) ---------------------------------
// 8 independent DFMA chains per thread, `iters` x 8 x 2 flops each; the result is stored so that nothing is optimised away.
__global__ void __launch_bounds__(256) k_fma_peak(int iters, double a, double b, double* __restrict__ out)
{
  double x0 = threadIdx.x*1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
#pragma unroll 4
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[size_t(blockIdx.x)*blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// ---- staged element scatter (LinearAlgebra::assemble path; lhsa.cpp:97-142) -----------------------
// One CTA walks the staged elements in submission order (deterministic); threads cover the entries
// of one element.  rows are SOLVER ids; the column search runs on the solver CSR (columns of a row
// keep the assembly order, i.e. they are sorted by assembly id, so we search linearly).
__global__ void __launch_bounds__(256)
k_scatter_staged(int nElem, int d, int dof, const int* __restrict__ rows, const int* __restrict__ pos,
                 const double* __restrict__ lK, const double* __restrict__ lR, double* __restrict__ R, double* __restrict__ Val)
{
  const int bs = dof*dof;
  for (int e = 0; e < nElem; e++) {
    const int* er = rows + size_t(e)*d;
    const int* ep = pos + size_t(e)*d*d;
    const double* eK = lK + size_t(e)*bs*d*d;
    const double* eR = lR + size_t(e)*dof*d;
    for (int t = threadIdx.x; t < d*dof; t += blockDim.x) {
      const int a = t / dof, i = t % dof;
      if (er[a] >= 0) R[size_t(er[a])*dof + i] += eR[t];
    }
    for (int t = threadIdx.x; t < d*d*bs; t += blockDim.x) {
      const int ab = t / bs, i = t % bs;      // lK(i,a,b) = eK[i + bs*(a + d*b)]
      const int a = ab % d, b = ab / d;
      const int p = ep[a*d + b];
      if (p >= 0) Val[size_t(p)*bs + i] += eK[t];
    }
    __syncthreads();
  }
}

} // namespace svb200
