// assembly_solid.cuh — K11: whole-mesh assembly of the displacement-based solid equations
//   struct   construct_dsolid + struct_3d_carray + get_pk2cc<3> + gnn + do_assem
//            (Code/Source/solver/sv_struct.cpp:213-362, 552-846; mat_models_carray.h:182-1359;
//             mat_models.cpp:1626-1645; nn.cpp:455-541; lhsa.cpp:97-142)
//   lElas    construct_l_elas + l_elas_3d                     (l_elas.cpp:58-170, 274-390)
//   mesh     construct_mesh (ALE mesh motion: l_elas_3d on the step-start configuration, weights
//            WITHOUT the Jacobian)                            (mesh.cpp:42-160)
// for TET4 (4 Gauss points, constant gradients) and HEX8 (8 Gauss points, gnn per point), dof = 3.
//
// Design: a CTA owns EPB consecutive elements and works in three phases on shared memory.
//   phase 1  one thread per (element, Gauss point): gather, Jacobian, shape gradients (gnn), F, the
//            constitutive law (S, 6x6 Voigt Dm), P = F S, inertia/body-force vector; the per-point
//            record (<= 73 doubles) goes to shared memory.
//   phase R  one thread per (element, a): residual rows, Gauss points summed in the reference's order.
//   phase 2  the eNoN x eNoN tangent blocks: a warp covers whole elements (HEX8: 32 lanes = 8 b x 4
//            pairs of a; TET4: 16 lanes per element), so F, S, Dm of a Gauss point are shared-memory
//            broadcasts; each lane forms Dm*Bm_b once per Gauss point and reuses it for its a's.
// The Gauss tables (w, N, dN/dxi) are staged in shared memory at kernel start.
// Scatter: the same destination-sorted staging as the fluid kernel (assembly.cuh): blocks / rows are
// written to precomputed slots and k_sum_run adds each destination's run in element order, i.e. in
// do_assem's order.  No atomics, bitwise reproducible.
#pragma once

#include "assembly.cuh"
#include "elem_tables.hpp"
#include "solid_law.hpp"

namespace svb200 {

enum { SREC_W = 0, SREC_F = 1, SREC_S = 10, SREC_DM = 16, SREC_UD = 37, SREC_P = 40, SREC_NX = 49 };
__host__ __device__ constexpr int solid_rec(int eNoN) { return SREC_NX + 3*eNoN; }

// ENON nodes, NG Gauss points, EPB elements per CTA, APT a-indices per lane in phase 2.
// blockDim.x = EPB*NG.  Dynamic shared memory: tables + EPB*NG records.
// ODOF: block size of the system the element is scattered into (3: struct/lElas/mesh equations; 4: the FSI
// equation, where struct_3d fills the 3x3 corner of lK(dof*dof,a,b) and leaves the pressure row/column zero,
// fsi.cpp:225).  elist != nullptr: the kernel covers the nEl elements elist[0..nEl) (one FSI domain).
// VISC: the extended struct element - solid viscosity (dmn.solid_visc: the record grows by VISC_REC doubles per Gauss point,
// visc_point's matrices, and phase 2 adds afu*Kvis_u + afv*Kvis_v; sv_struct.cpp:666-675, 771-842) and prestress (pS0 != null:
// S += S0 interpolated from the nodal prestress; stageP != null: the pstEq accumulations pSn += w N_a pSl, pSa += w N_a of
// construct_dsolid, sv_struct.cpp:646-700, 333-343; 6 more doubles per record for pSl).  A separate instantiation, so that the
// plain kernel keeps its record size and occupancy; inside it c.viscType / pS0 / stageP select at run time.
template <int ENON, int NG, int EPB, int APT, int ODOF, bool VISC = false>
__global__ void __launch_bounds__(EPB*NG)
k_assemble_solid(int nEl, const int* __restrict__ elist, SolidConsts c, const double* __restrict__ tab,      // packed: w[NG], N[NG][ENON], Nxi[NG][ENON][3]
                 const int* __restrict__ ien, const int* __restrict__ rslot, const int* __restrict__ kslot,
                 const double* __restrict__ x, const double* __restrict__ Ag, const double* __restrict__ Yg,
                 const double* __restrict__ Dg, const double* __restrict__ Do, const double* __restrict__ Bf,
                 const double* __restrict__ fN,      // 6 x nEl fibre + sheet directions (Holzapfel-Ogden) or null
                 double* __restrict__ stageR, double* __restrict__ stageK, int* __restrict__ err_flag,
                 const double* __restrict__ pS0 = nullptr,      // 6 x nNo nodal prestress (VISC instantiation only) or null
                 double* __restrict__ stageP = nullptr)         // 7 doubles per (element, a) slot: pSn contribution + pSa, or null
{
  constexpr int REC = solid_rec(ENON) + (VISC ? VISC_REC + 6 : 0);
  constexpr int SREC_V = solid_rec(ENON);      // visc_point's record (VISC only)
  constexpr int SREC_PS = SREC_V + VISC_REC;   // pSl: the stress before the prestress is added (VISC only)
  constexpr int NT = EPB*NG;
  constexpr int TABN = NG + NG*ENON + NG*ENON*3;
  extern __shared__ double sm[];
  double* s_w = sm;
  double* s_N = sm + NG;                 // [g][a]
  double* s_Nxi = s_N + NG*ENON;         // [g][a][3]
  double* s_rec = sm + ((TABN + 3) & ~3);
  for (int i = threadIdx.x; i < TABN; i += NT) sm[i] = tab[i];
  __syncthreads();

  const int e0 = blockIdx.x*EPB;
  const int tD = c.tDof, s0 = c.s;

  // ---------------- phase 1: one thread per (element, Gauss point) --------------------------------
  {
    const int el = threadIdx.x / NG, g = threadIdx.x % NG;
    const int live = (e0 + el) < nEl;
    const int e = live ? (elist ? elist[e0 + el] : e0 + el) : 0;
    double* rec = s_rec + size_t(threadIdx.x)*REC;
    if (live) {
      int nd[ENON];
#pragma unroll
      for (int a = 0; a < ENON; a++) nd[a] = ien[size_t(e)*ENON + a];
      // nn::gnn (nn.cpp:505-540).  For lShpF elements (TET4) the reference evaluates it at g = 0 only;
      // the gradients are constant, so evaluating it per point gives the same numbers.
      double xXi[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
      for (int a = 0; a < ENON; a++) {
        double xa[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
          xa[i] = x[size_t(nd[a])*3 + i];
          if (c.kind == 2) xa[i] = xa[i] + Do[size_t(nd[a])*tD + s0 + i];     // mesh.cpp:117-122
        }
        const double* nxi = s_Nxi + (g*ENON + a)*3;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          xXi[i][0] = xXi[i][0] + xa[i]*nxi[0];
          xXi[i][1] = xXi[i][1] + xa[i]*nxi[1];
          xXi[i][2] = xXi[i][2] + xa[i]*nxi[2];
        }
      }
      const double Jac = xXi[0][0]*xXi[1][1]*xXi[2][2] + xXi[0][1]*xXi[1][2]*xXi[2][0] + xXi[0][2]*xXi[1][0]*xXi[2][1]
                       - xXi[0][0]*xXi[1][2]*xXi[2][1] - xXi[0][1]*xXi[1][0]*xXi[2][2] - xXi[0][2]*xXi[1][1]*xXi[2][0];
      if (is_zero_d(Jac)) atomicExch(err_flag, e + 1);
      double xiX[3][3];
      xiX[0][0] = (xXi[1][1]*xXi[2][2] - xXi[1][2]*xXi[2][1])/Jac;
      xiX[0][1] = (xXi[2][1]*xXi[0][2] - xXi[2][2]*xXi[0][1])/Jac;
      xiX[0][2] = (xXi[0][1]*xXi[1][2] - xXi[0][2]*xXi[1][1])/Jac;
      xiX[1][0] = (xXi[1][2]*xXi[2][0] - xXi[1][0]*xXi[2][2])/Jac;
      xiX[1][1] = (xXi[2][2]*xXi[0][0] - xXi[2][0]*xXi[0][2])/Jac;
      xiX[1][2] = (xXi[0][2]*xXi[1][0] - xXi[0][0]*xXi[1][2])/Jac;
      xiX[2][0] = (xXi[1][0]*xXi[2][1] - xXi[1][1]*xXi[2][0])/Jac;
      xiX[2][1] = (xXi[2][0]*xXi[0][1] - xXi[2][1]*xXi[0][0])/Jac;
      xiX[2][2] = (xXi[0][0]*xXi[1][1] - xXi[0][1]*xXi[1][0])/Jac;

      // struct / lElas integrate with w*Jac, the mesh equation with w alone (mesh.cpp:141)
      rec[SREC_W] = (c.kind == 2) ? s_w[g] : s_w[g]*Jac;

      double F[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
      double vx[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};      // dv/dX (sv_struct.cpp:617-626), VISC only
      double ed[6] = {0, 0, 0, 0, 0, 0};
      double ud[3];
      if (c.kind == 0) { ud[0] = -c.rho*c.f[0]; ud[1] = -c.rho*c.f[1]; ud[2] = -c.rho*c.f[2]; }
      else { ud[0] = -c.f[0]; ud[1] = -c.f[1]; ud[2] = -c.f[2]; }
#pragma unroll
      for (int a = 0; a < ENON; a++) {
        const double* nxi = s_Nxi + (g*ENON + a)*3;
        double nx[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
          nx[i] = ((0.0 + nxi[0]*xiX[0][i]) + nxi[1]*xiX[1][i]) + nxi[2]*xiX[2][i];
          rec[SREC_NX + a*3 + i] = nx[i];
        }
        const double Na = s_N[g*ENON + a];
        const size_t A = size_t(nd[a]);
        double al[3], dl[3], yl[3], bl[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
          al[i] = Ag[A*tD + s0 + i];
          dl[i] = Dg[A*tD + s0 + i];
          if (c.kind == 2) { dl[i] = dl[i] - Do[A*tD + s0 + i]; bl[i] = 0.0; yl[i] = 0.0; }
          else { bl[i] = Bf[A*3 + i]; yl[i] = (c.kind == 0) ? Yg[A*tD + s0 + i] : 0.0; }
        }
        if (c.kind == 0) {
#pragma unroll
          for (int i = 0; i < 3; i++) {
            ud[i] += Na*(c.rho*(al[i] - bl[i]) + c.dmp*yl[i]);
            F[i][0] += nx[0]*dl[i];
            F[i][1] += nx[1]*dl[i];
            F[i][2] += nx[2]*dl[i];
            if (VISC && c.viscType != 0) { vx[i][0] += nx[0]*yl[i]; vx[i][1] += nx[1]*yl[i]; vx[i][2] += nx[2]*yl[i]; }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 3; i++) ud[i] = ud[i] + Na*(al[i] - bl[i]);
          ed[0] = ed[0] + nx[0]*dl[0];
          ed[1] = ed[1] + nx[1]*dl[1];
          ed[2] = ed[2] + nx[2]*dl[2];
          ed[3] = ed[3] + nx[1]*dl[0] + nx[0]*dl[1];
          ed[4] = ed[4] + nx[2]*dl[1] + nx[1]*dl[2];
          ed[5] = ed[5] + nx[0]*dl[2] + nx[2]*dl[0];
        }
      }
      rec[SREC_UD] = ud[0]; rec[SREC_UD + 1] = ud[1]; rec[SREC_UD + 2] = ud[2];
      if (c.kind == 0) {
        double S6[6];
        double fl[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (fN) {
#pragma unroll
          for (int i = 0; i < 6; i++) fl[i] = fN[size_t(e)*6 + i];
        }
        pk2cc_iso(c, F, fl, S6, rec + SREC_DM);
        if (VISC && c.viscType != 0) {
          // elastic + viscous stress (sv_struct.cpp:666-675); the record keeps the six entries 00 11 22 01 12 20 the reference
          // copies into pSl - its Newtonian Svis is symmetric up to rounding only
          double Sv[3][3];
          visc_point(c.viscType, c.visc_mu, F, vx, Sv, rec + SREC_V);
          S6[0] += Sv[0][0]; S6[1] += Sv[1][1]; S6[2] += Sv[2][2]; S6[3] += Sv[0][1]; S6[4] += Sv[1][2]; S6[5] += Sv[2][0];
        }
        if (VISC) {
          // prestress (sv_struct.cpp:683-700): pSl = S before S0 is added; S0 = sum_a N_a pS0(:,a), rows 00 11 22 01 12 20
#pragma unroll
          for (int i = 0; i < 6; i++) rec[SREC_PS + i] = S6[i];
          if (pS0) {
            double S0[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int a = 0; a < ENON; a++) {
              const double Na = s_N[g*ENON + a];
#pragma unroll
              for (int i = 0; i < 6; i++) S0[i] += Na*pS0[size_t(nd[a])*6 + i];
            }
#pragma unroll
            for (int i = 0; i < 6; i++) S6[i] += S0[i];
          }
        }
#pragma unroll
        for (int i = 0; i < 6; i++) rec[SREC_S + i] = S6[i];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) rec[SREC_F + i*3 + j] = F[i][j];
        // P = F S (mat_mul<3>)
        const double S[3][3] = {{S6[0], S6[3], S6[5]}, {S6[3], S6[1], S6[4]}, {S6[5], S6[4], S6[2]}};
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) rec[SREC_P + i*3 + j] = ((0.0 + F[i][0]*S[0][j]) + F[i][1]*S[1][j]) + F[i][2]*S[2][j];
      } else {
        // l_elas_3d (l_elas.cpp:300-360)
        const double lambda = c.elM*c.nu / (1.0 + c.nu) / (1.0 - 2.0*c.nu);
        const double mu = c.elM*0.5 / (1.0 + c.nu);
        const double divD = lambda*(ed[0] + ed[1] + ed[2]);
        rec[SREC_S + 0] = divD + 2.0*mu*ed[0];
        rec[SREC_S + 1] = divD + 2.0*mu*ed[1];
        rec[SREC_S + 2] = divD + 2.0*mu*ed[2];
        rec[SREC_S + 3] = mu*ed[3];
        rec[SREC_S + 4] = mu*ed[4];
        rec[SREC_S + 5] = mu*ed[5];
      }
    }
  }
  __syncthreads();

  // ---------------- phase R: residual rows, one thread per (element, a) ------------------------------
  for (int item = threadIdx.x; item < EPB*ENON; item += NT) {
    const int el = item / ENON, a = item % ENON;
    if (e0 + el >= nEl) continue;
    const int e = elist ? elist[e0 + el] : e0 + el;
    double r0 = 0.0, r1 = 0.0, r2 = 0.0;
    for (int g = 0; g < NG; g++) {
      const double* rec = s_rec + size_t(el*NG + g)*REC;
      const double w = rec[SREC_W];
      const double Na = s_N[g*ENON + a];
      const double n0 = rec[SREC_NX + a*3], n1 = rec[SREC_NX + a*3 + 1], n2 = rec[SREC_NX + a*3 + 2];
      if (c.kind == 0) {
        const double* P = rec + SREC_P;
        r0 = r0 + w*(Na*rec[SREC_UD]     + n0*P[0] + n1*P[1] + n2*P[2]);
        r1 = r1 + w*(Na*rec[SREC_UD + 1] + n0*P[3] + n1*P[4] + n2*P[5]);
        r2 = r2 + w*(Na*rec[SREC_UD + 2] + n0*P[6] + n1*P[7] + n2*P[8]);
      } else {
        const double* S = rec + SREC_S;
        r0 = r0 + w*(c.rho*Na*rec[SREC_UD]     + n0*S[0] + n1*S[3] + n2*S[5]);
        r1 = r1 + w*(c.rho*Na*rec[SREC_UD + 1] + n0*S[3] + n1*S[1] + n2*S[4]);
        r2 = r2 + w*(c.rho*Na*rec[SREC_UD + 2] + n0*S[5] + n1*S[4] + n2*S[2]);
      }
    }
    double* out = stageR + size_t(rslot[size_t(e)*ENON + a])*ODOF;
    out[0] = r0; out[1] = r1; out[2] = r2;
    if (ODOF == 4) out[3] = 0.0;
    if (VISC && stageP && c.kind == 0) {
      // pstEq (sv_struct.cpp:333-343): this element's share of pSn(:,Ac) and pSa(Ac)
      double ps[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      for (int g = 0; g < NG; g++) {
        const double* rec = s_rec + size_t(el*NG + g)*REC;
        const double wN = rec[SREC_W]*s_N[g*ENON + a];
#pragma unroll
        for (int i = 0; i < 6; i++) ps[i] = ps[i] + wN*rec[SREC_PS + i];
        ps[6] = ps[6] + wN;
      }
      double* po = stageP + size_t(rslot[size_t(e)*ENON + a])*7;
#pragma unroll
      for (int i = 0; i < 7; i++) po[i] = ps[i];
    }
  }

  // ---------------- phase 2: tangent blocks ------------------------------------------------------------
  constexpr int AGN = ENON/APT;            // a-groups per b
  constexpr int IPE = ENON*AGN;            // lanes per element
  const double afu = c.af*c.beta*c.dt*c.dt;
  const double afv = c.af*c.gam*c.dt;
  for (int item = threadIdx.x; item < EPB*IPE; item += NT) {
    const int el = item / IPE, r = item % IPE;
    const int b = r % ENON, a0 = (r / ENON)*APT;
    if (e0 + el >= nEl) continue;
    const int e = elist ? elist[e0 + el] : e0 + el;
    double acc[APT][9];
#pragma unroll
    for (int q = 0; q < APT; q++)
#pragma unroll
      for (int i = 0; i < 9; i++) acc[q][i] = 0.0;

    if (c.kind == 0) {
      const double amd = c.am*c.rho + c.af*c.gam*c.dt*c.dmp;
      for (int g = 0; g < NG; g++) {
        const double* rec = s_rec + size_t(el*NG + g)*REC;
        const double w = rec[SREC_W];
        const double* F = rec + SREC_F;
        const double* S = rec + SREC_S;
        const double* Dm = rec + SREC_DM;
        const double nb0 = rec[SREC_NX + b*3], nb1 = rec[SREC_NX + b*3 + 1], nb2 = rec[SREC_NX + b*3 + 2];
        const double Nb = s_N[g*ENON + b];
        // Bm(:,:,b) (sv_struct.cpp:716-742) and DBm = Dm Bm_b (mat_mul6x3)
        double Bb[6][3], DB[6][3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
          Bb[0][j] = nb0*F[j*3 + 0];
          Bb[1][j] = nb1*F[j*3 + 1];
          Bb[2][j] = nb2*F[j*3 + 2];
          Bb[3][j] = nb0*F[j*3 + 1] + F[j*3 + 0]*nb1;
          Bb[4][j] = nb1*F[j*3 + 2] + F[j*3 + 1]*nb2;
          Bb[5][j] = nb2*F[j*3 + 0] + F[j*3 + 2]*nb0;
        }
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) {
            double sum = 0.0;
#pragma unroll
            for (int k = 0; k < 6; k++) sum += Dm[(i <= k) ? dm_idx(i, k) : dm_idx(k, i)]*Bb[k][j];
            DB[i][j] = sum;
          }
#pragma unroll
        for (int q = 0; q < APT; q++) {
          const int a = a0 + q;
          const double na0 = rec[SREC_NX + a*3], na1 = rec[SREC_NX + a*3 + 1], na2 = rec[SREC_NX + a*3 + 2];
          const double Na = s_N[g*ENON + a];
          // geometric stiffness (sv_struct.cpp:753-757): S is exactly symmetric
          const double NxSNx = na0*S[0]*nb0 + na1*S[3]*nb0 + na2*S[5]*nb0 + na0*S[3]*nb1 + na1*S[1]*nb1
                             + na2*S[4]*nb1 + na0*S[5]*nb2 + na1*S[4]*nb2 + na2*S[2]*nb2;
          const double T1 = amd*Na*Nb + afu*NxSNx;
          double Ku[9], Kv[9];
          if (VISC) {
            if (c.viscType != 0) {
              const double na[3] = {na0, na1, na2}, nb[3] = {nb0, nb1, nb2};
              visc_pair(c.viscType, c.visc_mu, rec + SREC_V, F, na, nb, Ku, Kv);
            } else {
#pragma unroll
              for (int i = 0; i < 9; i++) { Ku[i] = 0.0; Kv[i] = 0.0; }
            }
          }
          double Ba[6][3];
#pragma unroll
          for (int j = 0; j < 3; j++) {
            Ba[0][j] = na0*F[j*3 + 0];
            Ba[1][j] = na1*F[j*3 + 1];
            Ba[2][j] = na2*F[j*3 + 2];
            Ba[3][j] = na0*F[j*3 + 1] + F[j*3 + 0]*na1;
            Ba[4][j] = na1*F[j*3 + 2] + F[j*3 + 1]*na2;
            Ba[5][j] = na2*F[j*3 + 0] + F[j*3 + 2]*na0;
          }
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
              const double BDB = Ba[0][i]*DB[0][j] + Ba[1][i]*DB[1][j] + Ba[2][i]*DB[2][j]
                               + Ba[3][i]*DB[3][j] + Ba[4][i]*DB[4][j] + Ba[5][i]*DB[5][j];
              if (VISC) acc[q][i*3 + j] = acc[q][i*3 + j] + w*(((i == j) ? T1 : 0.0) + afu*(BDB + Ku[i*3 + j]) + afv*Kv[i*3 + j]);
              else acc[q][i*3 + j] = acc[q][i*3 + j] + w*(((i == j) ? T1 : 0.0) + afu*BDB);
            }
        }
      }
    } else {
      // l_elas_3d tangent (l_elas.cpp:362-386)
      const double lambda = c.elM*c.nu / (1.0 + c.nu) / (1.0 - 2.0*c.nu);
      const double mu = c.elM*0.5 / (1.0 + c.nu);
      const double lDm = lambda/mu;
      const double amd = c.am/afu*c.rho;
      for (int g = 0; g < NG; g++) {
        const double* rec = s_rec + size_t(el*NG + g)*REC;
        const double wl = rec[SREC_W]*afu*mu;
        const double nb[3] = {rec[SREC_NX + b*3], rec[SREC_NX + b*3 + 1], rec[SREC_NX + b*3 + 2]};
        const double Nb = s_N[g*ENON + b];
#pragma unroll
        for (int q = 0; q < APT; q++) {
          const int a = a0 + q;
          const double na[3] = {rec[SREC_NX + a*3], rec[SREC_NX + a*3 + 1], rec[SREC_NX + a*3 + 2]};
          const double Na = s_N[g*ENON + a];
          const double NxdNx = na[0]*nb[0] + na[1]*nb[1] + na[2]*nb[2];
          const double T1 = amd*Na*Nb/mu + NxdNx;
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
              const double t = (i == j) ? (T1 + (1.0 + lDm)*na[i]*nb[i]) : (lDm*na[i]*nb[j] + na[j]*nb[i]);
              acc[q][i*3 + j] = acc[q][i*3 + j] + wl*t;
            }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < APT; q++) {
      double* out = stageK + size_t(kslot[(size_t(e)*ENON + (a0 + q))*ENON + b])*(ODOF*ODOF);
      if (ODOF == 3) {
#pragma unroll
        for (int i = 0; i < 9; i++) out[i] = acc[q][i];
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          d4 t;
          t.x = (i < 3) ? acc[q][i*3] : 0.0; t.y = (i < 3) ? acc[q][i*3 + 1] : 0.0; t.z = (i < 3) ? acc[q][i*3 + 2] : 0.0; t.w = 0.0;
          st256_stream(out + 4*i, t);
        }
      }
    }
  }
}

// Ordered run sum for arbitrary block sizes (dof 3: 9-double blocks, 3-double rows): one thread per
// destination double, runs walked left to right (element order) -> do_assem's order.
template <bool ASSIGN>
__global__ void __launch_bounds__(256)
k_sum_run(size_t nDest, int bs, const int* __restrict__ seg, const double* __restrict__ stage, double* __restrict__ out)
{
  const size_t t = size_t(blockIdx.x)*blockDim.x + threadIdx.x;
  const size_t d = t / bs;
  const int l = int(t % bs);
  if (d >= nDest) return;
  const int s = __ldg(seg + d), e = __ldg(seg + d + 1);
  double acc = ASSIGN ? 0.0 : out[t];
  for (int q = s; q < e; q++) acc += ld_stream(stage + size_t(q)*bs + l);
  out[t] = acc;
}

} // namespace svb200
