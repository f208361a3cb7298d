// pic.cuh — generalised-alpha time integrator on the device: predictor, initiator, corrector
// (pic::picp / pici / picc, Code/Source/solver/pic.cpp:591-715, 486-574, 74-330).  SURVEY.md par. 8(f) row 2: with
// Ao/Yo/Do/An/Yn/Dn/Ad resident in HBM the Newton loop needs no per-iteration upload of Ag/Yg/Dg and no download of
// the solution: pici writes the state the assembly kernels read, picc consumes the solver's device R.
//
// All arrays are (tDof, nNo) in ASSEMBLY order like com_mod.An (the dof values of a node are contiguous); the solver's
// R is (dof, nNo) in solver ordering and is reached through lhs.map.  Every expression is evaluated with explicit
// round-to-nearest multiplies and adds in the reference's order (no FMA contraction): the results are bit-identical
// to the reference's.
#pragma once

#include "kernels.cuh"

namespace svb200 {

// picp for one equation (pic.cpp:671-713).  dmode 0: Dn = Do; 1: Dn = Do + Yn*dt + An*coefD; 2: Dn left alone.
__global__ void k_picp(int nNo, int tDof, int s, int e, double coefA, int dmode, double dt, double coefD,
                       const double* __restrict__ Ao, const double* __restrict__ Yo, const double* __restrict__ Do,
                       double* __restrict__ An, double* __restrict__ Yn, double* __restrict__ Dn)
{
  const int nr = e - s + 1;
  const size_t n = size_t(nNo)*nr;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t t = size_t(blockIdx.x)*blockDim.x + threadIdx.x; t < n; t += nth) {
    const size_t q = (t / nr)*tDof + s + (t % nr);
    const double an = __dmul_rn(Ao[q], coefA);
    const double yn = Yo[q];
    An[q] = an;
    Yn[q] = yn;
    if (dmode == 0) Dn[q] = Do[q];
    else if (dmode == 1) Dn[q] = __dadd_rn(__dadd_rn(Do[q], __dmul_rn(yn, dt)), __dmul_rn(an, coefD));
  }
}

// pici for one equation (pic.cpp:556-566): Ag = Ao (1-am) + An am; Yg = Yo (1-af) + Yn af; Dg = Do (1-af) + Dn af
__global__ void k_pici(int nNo, int tDof, int s, int e, double c0, double c1, double c2, double c3,
                       const double* __restrict__ Ao, const double* __restrict__ An, const double* __restrict__ Yo,
                       const double* __restrict__ Yn, const double* __restrict__ Do, const double* __restrict__ Dn,
                       double* __restrict__ Ag, double* __restrict__ Yg, double* __restrict__ Dg)
{
  const int nr = e - s + 1;
  const size_t n = size_t(nNo)*nr;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t t = size_t(blockIdx.x)*blockDim.x + threadIdx.x; t < n; t += nth) {
    const size_t q = (t / nr)*tDof + s + (t % nr);
    Ag[q] = __dadd_rn(__dmul_rn(Ao[q], c0), __dmul_rn(An[q], c1));
    Yg[q] = __dadd_rn(__dmul_rn(Yo[q], c2), __dmul_rn(Yn[q], c3));
    Dg[q] = __dadd_rn(__dmul_rn(Do[q], c2), __dmul_rn(Dn[q], c3));
  }
}

// picc, general branch (pic.cpp:148-160, and the mesh equation :136-144): An -= R; Yn -= R gam dt; Dn -= R beta dt^2.
// R(dof,nNo) in solver ordering; one thread per node.
__global__ void k_picc(int nNo, int tDof, int s, int dof, double cY, double cD, const int* __restrict__ map,
                       const double* __restrict__ R, double* __restrict__ An, double* __restrict__ Yn, double* __restrict__ Dn)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t a = size_t(blockIdx.x)*blockDim.x + threadIdx.x; a < size_t(nNo); a += nth) {
    const double* r = R + size_t(map[a])*dof;
    for (int i = 0; i < dof; i++) {
      const size_t q = a*tDof + s + i;
      const double ri = r[i];
      An[q] = __dadd_rn(An[q], -ri);
      Yn[q] = __dadd_rn(Yn[q], -__dmul_rn(ri, cY));
      Dn[q] = __dadd_rn(Dn[q], -__dmul_rn(ri, cD));
    }
  }
}

// picc, ustruct / FSI under sstEq (pic.cpp:118-134): An, Yn from R(0..dof-1); Ad, Dn from Rd and R(0..dof-2).
// Rd is what ustruct_r left in com_mod.Rd (ustruct.cpp:1753-1764): amg Ad - Yg(s..) on the first Newton iteration
// of the time step, zero afterwards.
__global__ void k_picc_ustruct(int nNo, int tDof, int s, int dof, double c0, double c2, double c3, int first_itr, double amg,
                               const int* __restrict__ map, const double* __restrict__ R, const double* __restrict__ Yg,
                               double* __restrict__ An, double* __restrict__ Yn, double* __restrict__ Dn, double* __restrict__ Ad)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t a = size_t(blockIdx.x)*blockDim.x + threadIdx.x; a < size_t(nNo); a += nth) {
    const double* r = R + size_t(map[a])*dof;
    for (int i = 0; i < dof; i++) {
      const size_t q = a*tDof + s + i;
      An[q] = __dadd_rn(An[q], -r[i]);
      Yn[q] = __dadd_rn(Yn[q], -__dmul_rn(r[i], c0));
    }
    for (int i = 0; i < dof - 1; i++) {
      const double ad = Ad[a*3 + i];
      const double rd = first_itr ? __dadd_rn(__dmul_rn(amg, ad), -Yg[a*tDof + s + i]) : 0.0;
      const double dUl = __dadd_rn(__dmul_rn(rd, c2), __dmul_rn(r[i], c3));
      Ad[a*3 + i] = __dadd_rn(ad, -dUl);
      const size_t q = a*tDof + s + i;
      Dn[q] = __dadd_rn(Dn[q], -__dmul_rn(dUl, c0));
    }
  }
}

// picc of the FSI equation (pic.cpp:166-181): on the listed (solid-domain) nodes rows [0,cnt) are copied to [s2, s2+cnt)
__global__ void k_pic_copy_rows(int n, int tDof, int s2, int cnt, const int* __restrict__ nodes, double* __restrict__ An,
                                double* __restrict__ Yn, double* __restrict__ Dn)
{
  const size_t tot = size_t(n)*cnt;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t t = size_t(blockIdx.x)*blockDim.x + threadIdx.x; t < tot; t += nth) {
    const size_t a = size_t(nodes[t / cnt]);
    const int i = int(t % cnt);
    An[a*tDof + s2 + i] = An[a*tDof + i];
    Yn[a*tDof + s2 + i] = Yn[a*tDof + i];
    Dn[a*tDof + s2 + i] = Dn[a*tDof + i];
  }
}

// set_bc_dir writes (set_bc.cpp:794): arr[idx[k]] = val[k]
__global__ void k_pic_scatter(int n, const int* __restrict__ idx, const double* __restrict__ val, double* __restrict__ arr)
{
  const int nth = gridDim.x*blockDim.x;
  for (int t = blockIdx.x*blockDim.x + threadIdx.x; t < n; t += nth) arr[idx[t]] = val[t];
}

} // namespace svb200
