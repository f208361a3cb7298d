// assembly_fluid_gen.cuh — K10 for element types with non-constant gradients: whole-mesh Navier-Stokes (VMS,
// equal-order) assembly for HEX8 (8 Gauss points, gnn per point) and TET10 (15 points, gnn + gn_nxx per point).
//
// Replaces construct_fluid + fluid_3d_m + fluid_3d_c + gnn + gn_nxx + do_assem
// (Code/Source/solver/fluid.cpp:464-708, 1697-2139, 1389-1689; nn.cpp:455-541, 809-924; lhsa.cpp:97-142).
// The Gauss-point arithmetic lives in fluid_elem.hpp (host/device shared, checked on the CPU against the compiled
// reference); this file is the device orchestration:
//
//   a CTA owns EPB consecutive elements, blockDim.x = EPB*NG, tables (w, N, dN/dxi, d2N/dxi2) staged in shared
//   memory at kernel start;
//   phase A  one thread per (element, Gauss point): fluid_geom (Jacobian, metric, dN/dx, second derivatives)
//   phase B  same thread: fluid_point (both preambles; the continuity one reads the second derivatives of the
//            element's LAST Gauss point, hence the barrier between A and B for TET10)
//   phase R  one thread per (element, a): residual row, Gauss points summed in the reference's order
//   phase 2  one thread per (element, a, b): the 4x4 tangent block, Gauss points summed in order, whole 128-byte
//            block streamed to its staging slot
// Scatter: the destination-sorted staging of assembly.cuh (k_sum_segments adds each destination's run in element
// order = do_assem's order; no atomics, bitwise reproducible).
#pragma once

#include "assembly.cuh"

namespace svb200 {

template <int ENON, int NG, bool NXX>
__host__ __device__ constexpr int fluid_gen_tabn() { return NG + NG*ENON + NG*ENON*3 + (NXX ? NG*ENON*6 : 0); }

template <int ENON, int NG, int EPB, bool NXX>
__global__ void __launch_bounds__(EPB*NG)
k_assemble_fluid_gen(int nEl, const int* __restrict__ elist, const double* __restrict__ Dmesh, FluidConsts c,
                     const double* __restrict__ tab,      // packed: w[NG], N[NG][ENON], Nxi[NG][ENON][3], Nxi2[NG][ENON][6]
                     const int* __restrict__ ien, const int* __restrict__ rslot, const int* __restrict__ kslot,
                     const double* __restrict__ x, const double* __restrict__ Ag, const double* __restrict__ Yg,
                     const double* __restrict__ Bf, double* __restrict__ stageR, double* __restrict__ stageK,
                     int* __restrict__ err_flag)
{
  typedef FluidRec<ENON, NXX> L;
  constexpr int REC = L::SIZE;
  constexpr int NT = EPB*NG;
  constexpr int TABN = fluid_gen_tabn<ENON, NG, NXX>();
  extern __shared__ double sm[];
  const double* s_w = sm;
  const double* s_N = sm + NG;                      // [g][a]
  const double* s_Nxi = s_N + NG*ENON;              // [g][a][3]
  const double* s_Nxi2 = s_Nxi + NG*ENON*3;         // [g][a][6] (NXX only)
  double* s_rec = sm + ((TABN + 3) & ~3);
  for (int i = threadIdx.x; i < TABN; i += NT) sm[i] = tab[i];
  __syncthreads();

  const int e0 = blockIdx.x*EPB;

  // ---------------- phases A + B: one thread per (element, Gauss point) ------------------------------
  {
    const int el = threadIdx.x / NG, g = threadIdx.x % NG;
    const bool live = (e0 + el) < nEl;
    const int e = live ? (elist ? elist[e0 + el] : e0 + el) : 0;
    double* rec = s_rec + size_t(threadIdx.x)*REC;
    int nd[ENON];
    if (live) {
#pragma unroll
      for (int a = 0; a < ENON; a++) nd[a] = ien[size_t(e)*ENON + a];
      const bool ok = fluid_geom<ENON, NXX>(nd, x, Dmesh, c.tDof, s_w[g], s_Nxi + g*ENON*3, NXX ? s_Nxi2 + g*ENON*6 : nullptr, rec);
      if (!ok) atomicExch(err_flag, e + 1);
    }
    if (NXX) __syncthreads();
    if (live) {
      const double* recLast = s_rec + size_t(el*NG + NG - 1)*REC;
      fluid_point<ENON, NXX>(c, nd, Ag, Yg, Bf, s_N + g*ENON, rec, recLast);
    }
  }
  __syncthreads();

  // ---------------- phase R: residual rows, one thread per (element, a) ------------------------------
  for (int item = threadIdx.x; item < EPB*ENON; item += NT) {
    const int el = item / ENON, a = item % ENON;
    if (e0 + el >= nEl) continue;
    const int e = elist ? elist[e0 + el] : e0 + el;
    double r[4];
    fluid_res_row<ENON, NG, NXX>(c, s_rec + size_t(el*NG)*REC, s_N, a, r);
    d4 v; v.x = r[0]; v.y = r[1]; v.z = r[2]; v.w = r[3];
    st256_stream(stageR + size_t(rslot[size_t(e)*ENON + a])*4, v);
  }

  // ---------------- phase 2: tangent blocks, one thread per (element, a, b) --------------------------
  // b runs fastest over the lanes: the 16 x ENON slots of one (element, a) are read as a contiguous piece of kslot.
  for (int item = threadIdx.x; item < EPB*ENON*ENON; item += NT) {
    const int el = item / (ENON*ENON), r = item % (ENON*ENON);
    const int a = r / ENON, b = r % ENON;
    if (e0 + el >= nEl) continue;
    const int e = elist ? elist[e0 + el] : e0 + el;
    double kb[16];
    fluid_tan_block<ENON, NG, NXX>(c, s_rec + size_t(el*NG)*REC, s_N, a, b, kb);
    double* out = stageK + size_t(kslot[(size_t(e)*ENON + a)*ENON + b])*16;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      d4 t; t.x = kb[i*4]; t.y = kb[i*4 + 1]; t.z = kb[i*4 + 2]; t.w = kb[i*4 + 3];
      st256_stream(out + 4*i, t);
    }
  }
}

} // namespace svb200
