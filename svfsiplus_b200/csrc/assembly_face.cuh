// assembly_face.cuh — boundary-face (Neumann) assembly on the device: b_assem_neu_bc + gnnb + b_fluid / b_l_elas
// (Code/Source/solver/eq_assem.cpp:58-170, nn.cpp:552-755, fluid.cpp:46-133, l_elas.cpp:48-59) for TRI3, QUD4 and TRI6
// faces.  SURVEY.md par. 8(f) row 1: replaces the per-element LinearAlgebra::assemble calls of the boundary code (host
// staging list + upload every Newton iteration).  The arithmetic is face_elem.hpp (host/device shared).
//
// Faces are small (1e3..1e5 elements): one thread per face element; the element's rows / diagonal tangent scalars go
// to a face-private staging buffer whose slots were sorted at b200_face_mesh_set by destination and, inside a
// destination, by face element; a second kernel walks every touched destination once and adds its run onto R / Val in
// that order -- the order in which do_assem adds the face elements after the volume assembly.  No atomics.
#pragma once

#include "kernels.cuh"
#include "face_elem.hpp"

namespace svb200 {

template <int NB, int NG>
__global__ void __launch_bounds__(128)
k_bneu_elem(int nElb, BneuConsts c, const double* __restrict__ tab,     // packed: w[NG], N[NG][NB], Nx[NG][NB][2]
            const int* __restrict__ ienb, const int* __restrict__ inode, const int* __restrict__ rslot, const int* __restrict__ kslot,
            const double* __restrict__ x, const double* __restrict__ Do, const double* __restrict__ hg, const double* __restrict__ Yg,
            double* __restrict__ stageR, double* __restrict__ stageT)
{
  __shared__ double s_tab[NG + NG*NB + NG*NB*2];
  for (int i = threadIdx.x; i < NG + NG*NB + NG*NB*2; i += blockDim.x) s_tab[i] = tab[i];
  __syncthreads();
  const int e = blockIdx.x*blockDim.x + threadIdx.x;
  if (e >= nElb) return;
  int nd[NB];
#pragma unroll
  for (int a = 0; a < NB; a++) nd[a] = ienb[size_t(e)*NB + a];
  double lR[NB*3], lKd[NB*NB];
  face_element<NB, NG>(c, nd, inode[e], x, Do, hg, Yg, s_tab, s_tab + NG, s_tab + NG + NG*NB, lR, lKd);
#pragma unroll
  for (int a = 0; a < NB; a++) {
    double* o = stageR + size_t(rslot[size_t(e)*NB + a])*3;
    o[0] = lR[a*3]; o[1] = lR[a*3 + 1]; o[2] = lR[a*3 + 2];
  }
  if (c.kind == 0) {
#pragma unroll
    for (int q = 0; q < NB*NB; q++) stageT[kslot[size_t(e)*NB*NB + q]] = lKd[q];
  }
}

// R(0:2, row) += run of staged rows, in face-element order.  One thread per touched row.
__global__ void k_bneu_sum_R(int nU, int dof, const int* __restrict__ udest, const int* __restrict__ useg, const double* __restrict__ stageR,
                             double* __restrict__ R)
{
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= nU) return;
  double* r = R + size_t(udest[t])*dof;
  double a0 = r[0], a1 = r[1], a2 = r[2];
  for (int q = useg[t]; q < useg[t + 1]; q++) {
    a0 += stageR[size_t(q)*3]; a1 += stageR[size_t(q)*3 + 1]; a2 += stageR[size_t(q)*3 + 2];
  }
  r[0] = a0; r[1] = a1; r[2] = a2;
}

// Val(0, 5, 10; block) += run of staged scalars (the three equal diagonal entries b_fluid adds), in face-element order.
__global__ void k_bneu_sum_K(int nU, const int* __restrict__ udest, const int* __restrict__ useg, const double* __restrict__ stageT,
                             double* __restrict__ Val)
{
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= nU) return;
  double* v = Val + size_t(udest[t])*16;
  double a0 = v[0], a1 = v[5], a2 = v[10];
  for (int q = useg[t]; q < useg[t + 1]; q++) {
    const double s = stageT[q];
    a0 += s; a1 += s; a2 += s;
  }
  v[0] = a0; v[5] = a1; v[10] = a2;
}

// all_fun::integ: one thread per face element writes its NG terms; k_face_integ_sum adds them serially in (element,
// Gauss point) order -- the reference's running sum, bit for bit (faces are small: <= 1e5 elements).
template <int NB, int NG>
__global__ void __launch_bounds__(128)
k_face_integ_terms(int nElb, const double* __restrict__ tab, const int* __restrict__ ienb, const int* __restrict__ inode,
                   const double* __restrict__ x, const double* __restrict__ geo, int gtD, int goff, const double* __restrict__ s,
                   int stD, int l, int nrow, double* __restrict__ terms)
{
  __shared__ double s_tab[NG + NG*NB + NG*NB*2];
  for (int i = threadIdx.x; i < NG + NG*NB + NG*NB*2; i += blockDim.x) s_tab[i] = tab[i];
  __syncthreads();
  const int e = blockIdx.x*blockDim.x + threadIdx.x;
  if (e >= nElb) return;
  int nd[NB];
#pragma unroll
  for (int a = 0; a < NB; a++) nd[a] = ienb[size_t(e)*NB + a];
  double t[NG];
  face_integ_terms<NB, NG>(nd, inode[e], x, geo, gtD, goff, s, stD, l, nrow, s_tab, s_tab + NG, s_tab + NG + NG*NB, t);
#pragma unroll
  for (int g = 0; g < NG; g++) terms[size_t(e)*NG + g] = t[g];
}

__global__ void k_face_integ_sum(size_t n, const double* __restrict__ terms, double* __restrict__ out)
{
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double r = 0.0;
  for (size_t i = 0; i < n; i++) r = r + terms[i];
  *out = r;
}

// fsi_ls_upd: sV(i, node) = int N_a n_i dGamma.  Per face element the NB x NG x 3 terms go to the element's row slots;
// k_face_nrm_sum adds every node's run in (element, Gauss point) order into a nodal buffer (solver ordering).
template <int NB, int NG>
__global__ void __launch_bounds__(128)
k_face_nrm_elem(int nElb, const double* __restrict__ tab, const int* __restrict__ ienb, const int* __restrict__ inode,
                const int* __restrict__ rslot, const double* __restrict__ x, const double* __restrict__ geo, int gtD, int goff,
                double* __restrict__ stage)
{
  __shared__ double s_tab[NG + NG*NB + NG*NB*2];
  for (int i = threadIdx.x; i < NG + NG*NB + NG*NB*2; i += blockDim.x) s_tab[i] = tab[i];
  __syncthreads();
  const int e = blockIdx.x*blockDim.x + threadIdx.x;
  if (e >= nElb) return;
  int nd[NB];
#pragma unroll
  for (int a = 0; a < NB; a++) nd[a] = ienb[size_t(e)*NB + a];
  double t[NB*NG*3];
  face_normal_terms<NB, NG>(nd, inode[e], x, geo, gtD, goff, s_tab, s_tab + NG, s_tab + NG + NG*NB, t);
#pragma unroll
  for (int a = 0; a < NB; a++) {
    double* o = stage + size_t(rslot[size_t(e)*NB + a])*(NG*3);
#pragma unroll
    for (int q = 0; q < NG*3; q++) o[q] = t[a*NG*3 + q];
  }
}

__global__ void k_face_nrm_sum(int nU, int ng, const int* __restrict__ udest, const int* __restrict__ useg, const double* __restrict__ stage,
                               double* __restrict__ buf)
{
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= nU) return;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int q = useg[t]; q < useg[t + 1]; q++)
    for (int g = 0; g < ng; g++) {
      const double* s = stage + (size_t(q)*ng + g)*3;
      a0 = a0 + s[0]; a1 = a1 + s[1]; a2 = a2 + s[2];
    }
  double* o = buf + size_t(udest[t])*3;
  o[0] = a0; o[1] = a1; o[2] = a2;
}

// lhs.face val(0:2, a) = buf(0:2, glob[a])
__global__ void k_face_gather3(int fnNo, const int* __restrict__ glob, const double* __restrict__ buf, double* __restrict__ val)
{
  const int n = fnNo*3;
  for (int t = blockIdx.x*blockDim.x + threadIdx.x; t < n; t += gridDim.x*blockDim.x) val[t] = buf[size_t(glob[t/3])*3 + t % 3];
}

// Follower pressure load on a struct face (b_neu_folw_p + b_struct_3d; face_follower_element): one thread per face element.
// The element's NP rows and NP x NP x 6 off-diagonal tangent entries go to the face's parent-pair slots (sorted by
// destination and face element); k_bfolw_sum_K adds every block's run onto the dof-3 Val in that order.
template <int NP, int NB, int NG>
__global__ void __launch_bounds__(64)
k_bfolw_elem(int nElb, FolwConsts c, const double* __restrict__ tab, const int* __restrict__ ienb, const int* __restrict__ parent,
             const int* __restrict__ inode, const int* __restrict__ rslot, const int* __restrict__ kslot, const double* __restrict__ x,
             const double* __restrict__ Dg, const double* __restrict__ hg, double* __restrict__ stageR, double* __restrict__ stageK,
             double* __restrict__ stageKm,      // ustruct: the afm-scaled copy for the velocity block of Val; null for struct
             int* __restrict__ err_flag)
{
  __shared__ double s_tab[NG + NG*NB + NG*NB*2];
  for (int i = threadIdx.x; i < NG + NG*NB + NG*NB*2; i += blockDim.x) s_tab[i] = tab[i];
  __syncthreads();
  const int e = blockIdx.x*blockDim.x + threadIdx.x;
  if (e >= nElb) return;
  int nd[NB], pn[NP];
#pragma unroll
  for (int a = 0; a < NB; a++) nd[a] = ienb[size_t(e)*NB + a];
#pragma unroll
  for (int a = 0; a < NP; a++) pn[a] = parent[size_t(e)*NP + a];
  double lR[NP*3], lK6[NP*NP*6], lK6m[NP*NP*6];
  if (face_follower_element<NP, NB, NG>(c, pn, nd, inode[e], x, Dg, hg, s_tab, s_tab + NG, s_tab + NG + NG*NB, lR, lK6,
                                        stageKm ? lK6m : nullptr) != 0)
    atomicExch(err_flag, e + 1);
  for (int a = 0; a < NP; a++) {
    double* o = stageR + size_t(rslot[size_t(e)*NP + a])*3;
    o[0] = lR[a*3]; o[1] = lR[a*3 + 1]; o[2] = lR[a*3 + 2];
  }
  for (int q = 0; q < NP*NP; q++) {
    const size_t sl = size_t(kslot[size_t(e)*NP*NP + q])*6;
    for (int i = 0; i < 6; i++) stageK[sl + i] = lK6[q*6 + i];
    if (stageKm) for (int i = 0; i < 6; i++) stageKm[sl + i] = lK6m[q*6 + i];
  }
}

// out(block of bs doubles) += run of the six staged off-diagonal entries (0,1), (1,0), (0,2), (2,0), (1,2), (2,1), whose positions
// inside the block are idx: {1,3,2,6,5,7} in a 3x3 block (struct Val, ustruct Kd), {1,4,2,8,6,9} in a 4x4 block (ustruct Val).
struct Idx6 { int i[6]; };
__global__ void k_bfolw_sum_K(int nU, int bs, Idx6 idx, const int* __restrict__ udest, const int* __restrict__ useg,
                              const double* __restrict__ stageK, double* __restrict__ out)
{
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= nU) return;
  double* v = out + size_t(udest[t])*bs;
  double a0 = v[idx.i[0]], a1 = v[idx.i[1]], a2 = v[idx.i[2]], a3 = v[idx.i[3]], a4 = v[idx.i[4]], a5 = v[idx.i[5]];
  for (int q = useg[t]; q < useg[t + 1]; q++) {
    const double* s = stageK + size_t(q)*6;
    a0 += s[0]; a1 += s[1]; a2 += s[2]; a3 += s[3]; a4 += s[4]; a5 += s[5];
  }
  v[idx.i[0]] = a0; v[idx.i[1]] = a1; v[idx.i[2]] = a2; v[idx.i[3]] = a3; v[idx.i[4]] = a4; v[idx.i[5]] = a5;
}

// rows of IEN for a list of elements (face parents), to find the interior node on the host
__global__ void k_gather_ien(int n, int eNoN, const int* __restrict__ gE, const int* __restrict__ ien, int* __restrict__ out)
{
  const int tot = n*eNoN;
  for (int t = blockIdx.x*blockDim.x + threadIdx.x; t < tot; t += gridDim.x*blockDim.x)
    out[t] = ien[size_t(gE[t / eNoN])*eNoN + (t % eNoN)];
}

} // namespace svb200
