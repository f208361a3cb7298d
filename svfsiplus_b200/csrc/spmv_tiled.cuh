// spmv_tiled.cuh — the narrow FSILS SpMV shapes (dof-3 block rows of 72 bytes, 24-byte G / Gt entries, scalar L) with the
// matrix stream staged through shared memory by TMA bulk copies (cp.async.bulk, SASS UBLKCP) instead of per-lane loads.
//
// Why: a quad of lanes per row makes every 8-byte load instruction of a warp touch eight strided row segments; ncu showed
// k_spmv_vv<3> / k_schur_gp at 83-96 % l1tex throughput with DRAM at ~60 % (profiles/r01_tour_c_ncu_raw.csv): ~13 L1 sector
// requests per 72-byte block - an L1-wavefront wall, not an HBM wall.  Here a CTA owns a TILE of consecutive rows (CSR rows are
// contiguous in memory, so a tile's matrix entries and column ids are ONE contiguous byte range each); an elected thread issues two
// bulk copies per tile (entries, column ids) that land in shared memory without passing through L1 or the register file, double
// buffered on mbarriers so the next tile streams in while the current one is consumed.  The lanes then read the entries from shared
// memory (conflict-free: consecutive lanes read consecutive 72-byte blocks) and only the gathered vector still goes through L1/L2,
// exactly as in the dof-4 kernel.  Bytes moved = algorithmic bytes (no padding, no re-layout of the matrix).
//
// Tiles are cut on the host at b200_lhs_create (CudaOps::build_tiles): at most kTileRows rows and kTileCap entries, never across
// the boundary-row / interior-row split of rows_then_halo.  Lane l of a quad takes entries s+l, s+l+4, ... of its row and the quad
// adds its partial sums in a fixed order (the scheme the Schur passes already use): deterministic, rounding-level different from the
// strictly sequential sum.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace svb200 {

constexpr int kTileThreads = 512;
constexpr int kTileRows = kTileThreads/4;      // one quad of lanes per row
constexpr int kTileCap = 1152;                 // entries per tile (72-byte blocks: 81 KB + 4.5 KB of column ids per stage)
constexpr int kTileStages = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// global -> shared bulk copy (TMA, 1-D): 16-byte aligned addresses and size, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- shapes --------------------------------------------------------------------------------------------------------------
// EW: doubles per matrix entry; NA: accumulators per lane; entry(): one non-zero; store(): after the quad reduction (lane4 valid
// for all four lanes, acc identical on all of them).
struct TileVV3 {            // KU(3,i) = sum_j K(3x3,j) U(3,col_j)        fsils_spar_mul_vv dof 3, spar_mul.cpp:191
  static constexpr int ID = 0, EW = 9, NA = 3;
  const double* U; double* KU;
  __device__ __forceinline__ void entry(const double* k, int c, double (&a)[NA]) const
  {
    const double* u = U + size_t(c)*3;
    const double u0 = __ldg(u), u1 = __ldg(u + 1), u2 = __ldg(u + 2);
    a[0] = a[0] + (k[0]*u0 + k[1]*u1 + k[2]*u2);
    a[1] = a[1] + (k[3]*u0 + k[4]*u1 + k[5]*u2);
    a[2] = a[2] + (k[6]*u0 + k[7]*u1 + k[8]*u2);
  }
  __device__ __forceinline__ void store(int row, int lane4, const double (&a)[NA]) const
  {
    if (lane4 < 3) KU[size_t(row)*3 + lane4] = (lane4 == 0) ? a[0] : (lane4 == 1) ? a[1] : a[2];
  }
};
struct TileGP {             // pass 1 of the Schur operator: V4(i) = [sum_j G(:,j) P(col_j), P(i)]      cgrad.cpp:96-100
  static constexpr int ID = 1, EW = 3, NA = 3;
  const double* P; const double* Pown; double* V4;
  __device__ __forceinline__ void entry(const double* g, int c, double (&a)[NA]) const
  {
    const double u = __ldg(P + c);
    a[0] = fma(g[0], u, a[0]); a[1] = fma(g[1], u, a[1]); a[2] = fma(g[2], u, a[2]);
  }
  __device__ __forceinline__ void store(int row, int lane4, const double (&a)[NA]) const
  {
    if (lane4 == 0) { d4 o; o.x = a[0]; o.y = a[1]; o.z = a[2]; o.w = __ldg(Pown + row); st256(V4 + size_t(row)*4, o); }
  }
};
struct TileSV3 {            // KU(3,i) = sum_j K(3,j) U(col_j)            fsils_spar_mul_sv, spar_mul.cpp:63
  static constexpr int ID = 2, EW = 3, NA = 3;
  const double* U; double* KU;
  __device__ __forceinline__ void entry(const double* g, int c, double (&a)[NA]) const
  {
    const double u = __ldg(U + c);
    a[0] = fma(g[0], u, a[0]); a[1] = fma(g[1], u, a[1]); a[2] = fma(g[2], u, a[2]);
  }
  __device__ __forceinline__ void store(int row, int lane4, const double (&a)[NA]) const
  {
    if (lane4 < 3) KU[size_t(row)*3 + lane4] = (lane4 == 0) ? a[0] : (lane4 == 1) ? a[1] : a[2];
  }
};
struct TileVS3 {            // KU(i) = sum_j K(:,j) . U(:,col_j)          fsils_spar_mul_vs, spar_mul.cpp:129
  static constexpr int ID = 3, EW = 3, NA = 1;
  const double* U; double* KU;
  __device__ __forceinline__ void entry(const double* k, int c, double (&a)[NA]) const
  {
    const double* u = U + size_t(c)*3;
    a[0] = a[0] + (k[0]*__ldg(u) + k[1]*__ldg(u + 1) + k[2]*__ldg(u + 2));
  }
  __device__ __forceinline__ void store(int row, int lane4, const double (&a)[NA]) const { if (lane4 == 0) KU[row] = a[0]; }
};
struct TileSS {             // KU(i) = sum_j K(j) U(col_j)                fsils_spar_mul_ss, spar_mul.cpp:46
  static constexpr int ID = 4, EW = 1, NA = 1;
  const double* U; double* KU;
  __device__ __forceinline__ void entry(const double* k, int c, double (&a)[NA]) const { a[0] = fma(k[0], __ldg(U + c), a[0]); }
  __device__ __forceinline__ void store(int row, int lane4, const double (&a)[NA]) const { if (lane4 == 0) KU[row] = a[0]; }
};
struct TileSP {             // pass 2 of the Schur operator: SP(i) = sum_j L(j) V4(3,col_j) - sum_j Gt(:,j).V4(0:2,col_j)
  static constexpr int ID = 5, EW = 4, NA = 2;
  const double* V4; double* SP;
  __device__ __forceinline__ void entry(const double* k, int c, double (&a)[NA]) const
  {
    const d4 v = ld256_keep(V4 + size_t(c)*4);
    a[0] = fma(k[3], v.w, a[0]);
    a[1] = a[1] + (k[0]*v.x + k[1]*v.y + k[2]*v.z);
  }
  __device__ __forceinline__ void store(int row, int lane4, const double (&a)[NA]) const { if (lane4 == 0) SP[row] = a[0] - a[1]; }
};

template <class S> constexpr size_t tile_smem_bytes()
{
  return size_t(kTileStages)*(size_t(kTileCap)*S::EW*8 + size_t(kTileCap)*4) + 64;
}

// tiles [t0, t1) of tile_row (tile t = rows [tile_row[t], tile_row[t+1])); rowPtr / col / K are the whole arrays
template <class S>
__global__ void __launch_bounds__(kTileThreads, 1)
k_spmv_tiled(const int* __restrict__ skip, int t0, int t1, const int* __restrict__ tile_row, const int* __restrict__ rowPtr,
             const int* __restrict__ col, const double* __restrict__ K, S shape)
{
  if (skip && *skip) return;
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr size_t kBytesK = size_t(kTileCap)*S::EW*8, kBytesC = size_t(kTileCap)*4;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);                                   // kTileStages barriers
  unsigned char* base = smem + 64;
  auto sK = [&](int s) { return reinterpret_cast<double*>(base + size_t(s)*(kBytesK + kBytesC)); };
  auto sC = [&](int s) { return reinterpret_cast<int*>(base + size_t(s)*(kBytesK + kBytesC) + kBytesK); };

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < kTileStages; s++) mbar_init(bar + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int tile, int s) {          // one thread: both bulk copies of a tile into stage s
    const int r0 = __ldg(tile_row + tile), r1 = __ldg(tile_row + tile + 1);
    const int p0 = __ldg(rowPtr + r0) & ~3, p1 = __ldg(rowPtr + r1);
    const uint32_t bk = (uint32_t(p1 - p0)*S::EW*8 + 15u) & ~15u;
    const uint32_t bc = (uint32_t(p1 - p0)*4 + 15u) & ~15u;
    mbar_expect_tx(bar + s, bk + bc);
    bulk_g2s(sK(s), K + size_t(p0)*S::EW, bk, bar + s);
    bulk_g2s(sC(s), col + p0, bc, bar + s);
  };

  int tile = t0 + blockIdx.x;
  if (tile >= t1) return;
  if (tid == 0) issue(tile, 0);
  const int lane4 = tid & 3, quad = tid >> 2;
  uint32_t phases = 0;                                   // bit s: parity the next wait on stage s expects
  int s = 0;
  for (; tile < t1; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    if (tid == 0 && next < t1) issue(next, s ^ 1);       // stage s^1 was released by the __syncthreads that ended the previous tile
    const int r0 = __ldg(tile_row + tile), r1 = __ldg(tile_row + tile + 1);
    const int row = r0 + quad;
    int b = 0, e = 0;
    if (row < r1) { b = __ldg(rowPtr + row); e = __ldg(rowPtr + row + 1); }
    const int p0 = __ldg(rowPtr + r0) & ~3;
    mbar_wait(bar + s, (phases >> s) & 1u);
    phases ^= (1u << s);
    const double* k = sK(s);
    const int* c = sC(s);
    double acc[S::NA];
#pragma unroll
    for (int i = 0; i < S::NA; i++) acc[i] = 0.0;
#pragma unroll 2
    for (int p = b + lane4; p < e; p += 4) shape.entry(k + size_t(p - p0)*S::EW, c[p - p0], acc);
#pragma unroll
    for (int i = 0; i < S::NA; i++) {
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 1);
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 2);
    }
    if (row < r1) shape.store(row, lane4, acc);
    __syncthreads();                                      // everybody is done with stage s: it may be refilled
    s ^= 1;
  }
}

} // namespace svb200
