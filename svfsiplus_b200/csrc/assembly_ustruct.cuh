// assembly_ustruct.cuh — K11: whole-mesh assembly of the mixed velocity-pressure (VMS-stabilised) solid
//   construct_usolid + ustruct_3d_m + ustruct_3d_c + ustruct_do_assem
//            (Code/Source/solver/ustruct.cpp:216-430, 1158-1575, 632-876, 1579-1721)
//   get_pk2cc_dev (neo-Hookean), g_vol_pen, get_tau      (mat_models.cpp:630-700, 1696-1747, 1655-1676)
//   ustruct_r: R -= (1/am) Kd Rd on the first Newton iteration (ustruct.cpp:1726-1828)
// for equal-order elements (one function space: P1-P1 TET4, Q1-Q1 HEX8), dof = 4, with the extra
// displacement tangent Kd((nsd+1)*nsd, nnz).  idMap is the identity (no undeformed-Neumann faces), so
// ustruct_do_assem's four sub-block scatters address the same (row, column) entry.
//
// Same three-phase shared-memory design and ordered staging scatter as assembly_solid.cuh; the momentum
// (m) and continuity (c) Gauss loops of the reference use the same rule and basis here and share the
// kinematics.
#pragma once

#include "assembly_solid.cuh"

namespace svb200 {

struct UstructConsts {
  double dt, am, af, gam;        // eq.am, eq.af, eq.gam
  double rho0, f[3];
  double elM, nu, ctM, ctC;      // get_tau inputs
  int iso, vol;                  // iso as in SolidConsts (isochoric laws: 0, 3, 4, 5, 6); vol: 0 none, 1 Quad, 2 ST91, 3 M94
  double C10, Kpen;
  SolidConsts law;               // the isochoric law for pk2cc_iso: Kpen = 0, vol = 0 (get_pk2cc_dev has no volumetric part)
  int tDof, s;
};

// record layout per (element, Gauss point)
enum { UR_W = 0, UR_J = 1, UR_F = 2, UR_S = 11, UR_DM = 17, UR_P = 38, UR_VD = 47, UR_VXFI = 50, UR_PXFI = 59, UR_RM = 62,
       UR_RHO = 65, UR_BETA = 66, UR_DRHO = 67, UR_DBETA = 68, UR_TAUM = 69, UR_TAUC = 70, UR_RC = 71, UR_RCL = 72, UR_PD = 73,
       UR_NX = 74 };
__host__ __device__ constexpr int ustruct_rec(int eNoN) { return UR_NX + 6*eNoN; }   // + Nx[a][3], NxFi[a][3]

// VISC: solid viscosity (dmn.solid_visc; ustruct.cpp:1275-1302, 1406-1550): Siso += Svis, Ku += Kvis_u, Tv = af Kvis_v - a separate
// instantiation with VISC_REC more doubles per Gauss-point record (visc_point's matrices), like k_assemble_solid.
template <int ENON, int NG, int EPB, int APT, bool VISC = false>
__global__ void __launch_bounds__(EPB*NG)
k_assemble_ustruct(int nEl, UstructConsts c, const double* __restrict__ tab, const int* __restrict__ ien,
                   const int* __restrict__ rslot, const int* __restrict__ kslot, const double* __restrict__ x,
                   const double* __restrict__ Ag, const double* __restrict__ Yg, const double* __restrict__ Dg,
                   const double* __restrict__ Bf, const double* __restrict__ fN, double* __restrict__ stageR,
                   double* __restrict__ stageK, double* __restrict__ stageKd, int* __restrict__ err_flag)
{
  constexpr int REC = ustruct_rec(ENON) + (VISC ? VISC_REC : 0);
  constexpr int UR_V = ustruct_rec(ENON);      // visc_point's record (VISC only)
  constexpr int NT = EPB*NG;
  constexpr int TABN = NG + NG*ENON + NG*ENON*3;
  extern __shared__ double sm[];
  double* s_w = sm;
  double* s_N = sm + NG;
  double* s_Nxi = s_N + NG*ENON;
  double* s_rec = sm + ((TABN + 3) & ~3);
  for (int i = threadIdx.x; i < TABN; i += NT) sm[i] = tab[i];
  __syncthreads();

  const int e0 = blockIdx.x*EPB;
  const int tD = c.tDof, s0 = c.s;
  const double am = c.am;
  const double af = c.af*c.gam*c.dt;
  const double afm = af/am;

  // ---------------- phase 1: one thread per (element, Gauss point) --------------------------------
  {
    const int el = threadIdx.x / NG, g = threadIdx.x % NG;
    const int e = e0 + el;
    double* rec = s_rec + size_t(threadIdx.x)*REC;
    if (e < nEl) {
      int nd[ENON];
#pragma unroll
      for (int a = 0; a < ENON; a++) nd[a] = ien[size_t(e)*ENON + a];
      double xXi[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
      for (int a = 0; a < ENON; a++) {
        const double* nxi = s_Nxi + (g*ENON + a)*3;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const double xa = x[size_t(nd[a])*3 + i];
          xXi[i][0] = xXi[i][0] + xa*nxi[0];
          xXi[i][1] = xXi[i][1] + xa*nxi[1];
          xXi[i][2] = xXi[i][2] + xa*nxi[2];
        }
      }
      const double Je = xXi[0][0]*xXi[1][1]*xXi[2][2] + xXi[0][1]*xXi[1][2]*xXi[2][0] + xXi[0][2]*xXi[1][0]*xXi[2][1]
                      - xXi[0][0]*xXi[1][2]*xXi[2][1] - xXi[0][1]*xXi[1][0]*xXi[2][2] - xXi[0][2]*xXi[1][1]*xXi[2][0];
      if (is_zero_d(Je)) atomicExch(err_flag, e + 1);
      double xiX[3][3];
      xiX[0][0] = (xXi[1][1]*xXi[2][2] - xXi[1][2]*xXi[2][1])/Je;
      xiX[0][1] = (xXi[2][1]*xXi[0][2] - xXi[2][2]*xXi[0][1])/Je;
      xiX[0][2] = (xXi[0][1]*xXi[1][2] - xXi[0][2]*xXi[1][1])/Je;
      xiX[1][0] = (xXi[1][2]*xXi[2][0] - xXi[1][0]*xXi[2][2])/Je;
      xiX[1][1] = (xXi[2][2]*xXi[0][0] - xXi[2][0]*xXi[0][2])/Je;
      xiX[1][2] = (xXi[0][2]*xXi[1][0] - xXi[0][0]*xXi[1][2])/Je;
      xiX[2][0] = (xXi[1][0]*xXi[2][1] - xXi[1][1]*xXi[2][0])/Je;
      xiX[2][1] = (xXi[2][0]*xXi[0][1] - xXi[2][1]*xXi[0][0])/Je;
      xiX[2][2] = (xXi[0][0]*xXi[1][1] - xXi[0][1]*xXi[1][0])/Je;
      rec[UR_W] = s_w[g]*Je;

      // kinematics (ustruct.cpp:1208-1256 / 676-727)
      double F[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
      double vx[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      double vd[3] = {-c.f[0], -c.f[1], -c.f[2]};
      double p = 0.0, pd = 0.0, px[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int a = 0; a < ENON; a++) {
        const double* nxi = s_Nxi + (g*ENON + a)*3;
        double nx[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
          nx[i] = ((0.0 + nxi[0]*xiX[0][i]) + nxi[1]*xiX[1][i]) + nxi[2]*xiX[2][i];
          rec[UR_NX + a*6 + i] = nx[i];
        }
        const double Na = s_N[g*ENON + a];
        const size_t A = size_t(nd[a]);
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const double al = Ag[A*tD + s0 + i], yl = Yg[A*tD + s0 + i], dl = Dg[A*tD + s0 + i], bl = Bf[A*3 + i];
          vd[i] = vd[i] + Na*(al - bl);
          vx[i][0] = vx[i][0] + nx[0]*yl; vx[i][1] = vx[i][1] + nx[1]*yl; vx[i][2] = vx[i][2] + nx[2]*yl;
          F[i][0] = F[i][0] + nx[0]*dl;   F[i][1] = F[i][1] + nx[1]*dl;   F[i][2] = F[i][2] + nx[2]*dl;
        }
        const double yp = Yg[A*tD + s0 + 3], ap = Ag[A*tD + s0 + 3];
        p = p + Na*yp;
        pd = pd + Na*ap;
        px[0] = px[0] + nx[0]*yp; px[1] = px[1] + nx[1]*yp; px[2] = px[2] + nx[2]*yp;
      }
      // mat_det / mat_inv (mat_fun.cpp:140-196)
      const double J = F[0][0]*F[1][1]*F[2][2] + F[0][1]*F[1][2]*F[2][0] + F[0][2]*F[1][0]*F[2][1]
                     - F[0][0]*F[1][2]*F[2][1] - F[0][1]*F[1][0]*F[2][2] - F[0][2]*F[1][1]*F[2][0];
      double Fi[3][3];
      Fi[0][0] = (F[1][1]*F[2][2] - F[1][2]*F[2][1]) / J;
      Fi[0][1] = (F[0][2]*F[2][1] - F[0][1]*F[2][2]) / J;
      Fi[0][2] = (F[0][1]*F[1][2] - F[0][2]*F[1][1]) / J;
      Fi[1][0] = (F[1][2]*F[2][0] - F[1][0]*F[2][2]) / J;
      Fi[1][1] = (F[0][0]*F[2][2] - F[0][2]*F[2][0]) / J;
      Fi[1][2] = (F[0][2]*F[1][0] - F[0][0]*F[1][2]) / J;
      Fi[2][0] = (F[1][0]*F[2][1] - F[1][1]*F[2][0]) / J;
      Fi[2][1] = (F[0][1]*F[2][0] - F[0][0]*F[2][1]) / J;
      Fi[2][2] = (F[0][0]*F[1][1] - F[0][1]*F[1][0]) / J;
      rec[UR_J] = J;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) rec[UR_F + i*3 + j] = F[i][j];

      // get_pk2cc_dev (mat_models.cpp:630-1000): deviatoric S and isochoric CC = the law of solid_law.hpp without its
      // volumetric terms (c.law has Kpen = 0, so p = pl = 0 there: c2 = 2 r1, c3 = -2 r1 / nd)
      double S6[6];
      {
        double fl[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (fN) {
#pragma unroll
          for (int i = 0; i < 6; i++) fl[i] = fN[size_t(e)*6 + i];
        }
        pk2cc_iso(c.law, F, fl, S6, rec + UR_DM);
        if (VISC) {
          // total isochoric stress = elastic + viscous (ustruct.cpp:1279, 1302); six entries kept, as in k_assemble_solid
          double Sv[3][3];
          visc_point(c.law.viscType, c.law.visc_mu, F, vx, Sv, rec + UR_V);
          S6[0] += Sv[0][0]; S6[1] += Sv[1][1]; S6[2] += Sv[2][2]; S6[3] += Sv[0][1]; S6[4] += Sv[1][2]; S6[5] += Sv[2][0];
        }
#pragma unroll
        for (int i = 0; i < 6; i++) rec[UR_S + i] = S6[i];
        // Pdev = F Siso
        const double S[3][3] = {{S6[0], S6[3], S6[5]}, {S6[3], S6[1], S6[4]}, {S6[5], S6[4], S6[2]}};
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) rec[UR_P + i*3 + j] = ((0.0 + F[i][0]*S[0][j]) + F[i][1]*S[1][j]) + F[i][2]*S[2][j];
      }

      // g_vol_pen (mat_models.cpp:1696-1747), Ja = 1
      double rho = c.rho0, beta = 0.0, drho = 0.0, dbeta = 0.0;
      if (!is_zero_d(c.Kpen)) {
        const double Kp = c.Kpen;
        if (c.vol == 1) { const double r1 = 1.0/(Kp - p); rho = rho*Kp*r1; beta = r1; drho = rho*r1; dbeta = r1*r1; }
        else if (c.vol == 2) { const double r1 = rho/Kp; const double r2 = sqrt(p*p + Kp*Kp); rho = r1*(p + r2); beta = 1.0/r2; drho = rho*beta; dbeta = -beta*p/(p*p + Kp*Kp); }
        else if (c.vol == 3) { const double r1 = rho/Kp; const double r2 = Kp + p; rho = r1*r2; beta = 1.0/r2; drho = r1; dbeta = -beta*beta; }
      }
      // get_tau (mat_models.cpp:1655-1676)
      double tauM, tauC;
      {
        const double he = 0.5*pow(Je, 1.0/3.0);
        const double mu = 0.5*c.elM / (1.0 + c.nu);
        double cw;
        if (is_zero_d(c.nu - 0.5)) cw = sqrt(mu / c.rho0);
        else { const double lam = 2.0*mu*c.nu / (1.0 - 2.0*c.nu); cw = sqrt((lam + 2.0*mu)/c.rho0); }
        tauM = c.ctM*(he/cw)*(J/c.rho0);
        tauC = c.ctC*(he*cw)*(c.rho0/J);
      }
      // VxFi = vx Fi, PxFi = px Fi (ustruct.cpp:1308, 762-776)
      double VxFi[3][3], PxFi[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
          VxFi[i][j] = vx[i][0]*Fi[0][j] + vx[i][1]*Fi[1][j] + vx[i][2]*Fi[2][j];
          rec[UR_VXFI + i*3 + j] = VxFi[i][j];
        }
        PxFi[i] = px[0]*Fi[0][i] + px[1]*Fi[1][i] + px[2]*Fi[2][i];
        rec[UR_PXFI + i] = PxFi[i];
      }
      const double rC = beta*pd + VxFi[0][0] + VxFi[1][1] + VxFi[2][2];
      const double rCl = -p + tauC*rC;
#pragma unroll
      for (int i = 0; i < 3; i++) { rec[UR_VD + i] = vd[i]; rec[UR_RM + i] = rho*vd[i] + PxFi[i]; }
      rec[UR_RHO] = rho; rec[UR_BETA] = beta; rec[UR_DRHO] = drho; rec[UR_DBETA] = dbeta;
      rec[UR_TAUM] = tauM; rec[UR_TAUC] = tauC; rec[UR_RC] = rC; rec[UR_RCL] = rCl; rec[UR_PD] = pd;
      // NxFi(:,a) (ustruct.cpp:1299-1304)
#pragma unroll
      for (int a = 0; a < ENON; a++) {
        const double n0 = rec[UR_NX + a*6], n1 = rec[UR_NX + a*6 + 1], n2 = rec[UR_NX + a*6 + 2];
        rec[UR_NX + a*6 + 3] = n0*Fi[0][0] + n1*Fi[1][0] + n2*Fi[2][0];
        rec[UR_NX + a*6 + 4] = n0*Fi[0][1] + n1*Fi[1][1] + n2*Fi[2][1];
        rec[UR_NX + a*6 + 5] = n0*Fi[0][2] + n1*Fi[1][2] + n2*Fi[2][2];
      }
    }
  }
  __syncthreads();

  // ---------------- phase R: residual rows, one thread per (element, a) ------------------------------
  for (int item = threadIdx.x; item < EPB*ENON; item += NT) {
    const int el = item / ENON, a = item % ENON;
    const int e = e0 + el;
    if (e >= nEl) continue;
    double r[4] = {0.0, 0.0, 0.0, 0.0};
    for (int g = 0; g < NG; g++) {
      const double* rec = s_rec + size_t(el*NG + g)*REC;
      const double w = rec[UR_W], J = rec[UR_J];
      const double Na = s_N[g*ENON + a];
      const double* nx = rec + UR_NX + a*6;
      const double* nf = nx + 3;
      const double* P = rec + UR_P;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double T1 = J*rec[UR_RHO]*rec[UR_VD + i]*Na;
        const double T2 = P[i*3]*nx[0] + P[i*3 + 1]*nx[1] + P[i*3 + 2]*nx[2];
        const double T3 = J*rec[UR_RCL]*nf[i];
        r[i] = r[i] + w*(T1 + T2 + T3);
      }
      const double rMNqx = rec[UR_RM]*nf[0] + rec[UR_RM + 1]*nf[1] + rec[UR_RM + 2]*nf[2];
      r[3] = r[3] + w*J*(Na*rec[UR_RC] + rec[UR_TAUM]*rMNqx);
    }
    d4 o; o.x = r[0]; o.y = r[1]; o.z = r[2]; o.w = r[3];
    st256_stream(stageR + size_t(rslot[size_t(e)*ENON + a])*4, o);
  }

  // ---------------- phase 2: tangent blocks lK(16) and lKd(12) ---------------------------------------------
  constexpr int AGN = ENON/APT;
  constexpr int IPE = ENON*AGN;
  for (int item = threadIdx.x; item < EPB*IPE; item += NT) {
    const int el = item / IPE, rr = item % IPE;
    const int b = rr % ENON, a0 = (rr / ENON)*APT;
    const int e = e0 + el;
    if (e >= nEl) continue;
    double K[APT][16], Kd[APT][12];
#pragma unroll
    for (int q = 0; q < APT; q++) {
#pragma unroll
      for (int i = 0; i < 16; i++) K[q][i] = 0.0;
#pragma unroll
      for (int i = 0; i < 12; i++) Kd[q][i] = 0.0;
    }
    for (int g = 0; g < NG; g++) {
      const double* rec = s_rec + size_t(el*NG + g)*REC;
      const double w = rec[UR_W], J = rec[UR_J];
      const double* F = rec + UR_F;
      const double* S = rec + UR_S;
      const double* Dm = rec + UR_DM;
      const double* vd = rec + UR_VD;
      const double* VxFi = rec + UR_VXFI;
      const double* PxFi = rec + UR_PXFI;
      const double* rM = rec + UR_RM;
      const double rho = rec[UR_RHO], beta = rec[UR_BETA], drho = rec[UR_DRHO], dbeta = rec[UR_DBETA];
      const double tauM = rec[UR_TAUM], tauC = rec[UR_TAUC], rC = rec[UR_RC], rCl = rec[UR_RCL], pd = rec[UR_PD];
      const double* nxb = rec + UR_NX + b*6;
      const double* nfb = nxb + 3;
      const double Nb = s_N[g*ENON + b];
      double Bb[6][3], DB[6][3];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        Bb[0][j] = nxb[0]*F[j*3 + 0];
        Bb[1][j] = nxb[1]*F[j*3 + 1];
        Bb[2][j] = nxb[2]*F[j*3 + 2];
        Bb[3][j] = nxb[0]*F[j*3 + 1] + F[j*3 + 0]*nxb[1];
        Bb[4][j] = nxb[1]*F[j*3 + 2] + F[j*3 + 1]*nxb[2];
        Bb[5][j] = nxb[2]*F[j*3 + 0] + F[j*3 + 2]*nxb[0];
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
      // VxNx(:,b), rMNwx(b)
      double VxNxb[3];
#pragma unroll
      for (int j = 0; j < 3; j++) VxNxb[j] = VxFi[0*3 + j]*nfb[0] + VxFi[1*3 + j]*nfb[1] + VxFi[2*3 + j]*nfb[2];
      const double rMNwxb = rM[0]*nfb[0] + rM[1]*nfb[1] + rM[2]*nfb[2];
      const double T0p = am*tauC*beta + af*(tauC*dbeta*pd - 1.0);

#pragma unroll
      for (int q = 0; q < APT; q++) {
        const int a = a0 + q;
        const double* nxa = rec + UR_NX + a*6;
        const double* nfa = nxa + 3;
        const double Na = s_N[g*ENON + a];
        const double NxSNx = nxa[0]*S[0]*nxb[0] + nxa[0]*S[3]*nxb[1] + nxa[0]*S[5]*nxb[2]
                           + nxa[1]*S[3]*nxb[0] + nxa[1]*S[1]*nxb[1] + nxa[1]*S[4]*nxb[2]
                           + nxa[2]*S[5]*nxb[0] + nxa[2]*S[4]*nxb[1] + nxa[2]*S[2]*nxb[2];
        double Ba[6][3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
          Ba[0][j] = nxa[0]*F[j*3 + 0];
          Ba[1][j] = nxa[1]*F[j*3 + 1];
          Ba[2][j] = nxa[2]*F[j*3 + 2];
          Ba[3][j] = nxa[0]*F[j*3 + 1] + F[j*3 + 0]*nxa[1];
          Ba[4][j] = nxa[1]*F[j*3 + 2] + F[j*3 + 1]*nxa[2];
          Ba[5][j] = nxa[2]*F[j*3 + 0] + F[j*3 + 2]*nxa[0];
        }
        double Kvu[9], Kvv[9];
        if (VISC) visc_pair(c.law.viscType, c.law.visc_mu, rec + UR_V, F, nxa, nxb, Kvu, Kvv);
        // A block (ustruct.cpp:1398-1538)
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) {
            const double BtDB = Ba[0][i]*DB[0][j] + Ba[1][i]*DB[1][j] + Ba[2][i]*DB[2][j]
                              + Ba[3][i]*DB[3][j] + Ba[4][i]*DB[4][j] + Ba[5][i]*DB[5][j];
            const double T1 = J*rho*vd[i]*Na*nfb[j];
            const double T2 = -tauC*J*nfa[i]*VxNxb[j];
            double Ku, T2k;
            if (i == j) {
              Ku = w*af*(T1 + T2 + BtDB + NxSNx + (VISC ? Kvu[i*3 + j] : 0.0));
              const double T1k = am*J*rho*Na*Nb;
              T2k = T1k + af*J*tauC*rho*nfa[i]*nfb[i];
            } else {
              const double T3 = J*rCl*(nfa[i]*nfb[j] - nfa[j]*nfb[i]);
              Ku = w*af*(T1 + T2 + T3 + BtDB + (VISC ? Kvu[i*3 + j] : 0.0));
              T2k = af*J*tauC*rho*nfa[i]*nfb[j];
            }
            Kd[q][i*3 + j] = Kd[q][i*3 + j] + Ku;
            K[q][i*4 + j] = K[q][i*4 + j] + w*(T2k + (VISC ? af*Kvv[i*3 + j] : 0.0)) + afm*Ku;
          }
        // B block (ustruct.cpp:1555-1573)
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const double T1 = T0p*nfa[i]*Nb + af*drho*vd[i]*Na*Nb;
          K[q][i*4 + 3] = K[q][i*4 + 3] + w*J*T1;
        }
        // C block (ustruct.cpp:816-857)
        const double rMNqxa = rM[0]*nfa[0] + rM[1]*nfa[1] + rM[2]*nfa[2];
        const double NxNx = nfa[0]*nfb[0] + nfa[1]*nfb[1] + nfa[2]*nfb[2];
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const double T0 = Na*(rC*nfb[j] - VxNxb[j]);
          const double T1 = tauM*(rMNqxa*nfb[j] - rMNwxb*nfa[j]);
          const double T2 = -tauM*NxNx*PxFi[j];
          const double Ku = w*af*J*(T0 + T1 + T2);
          Kd[q][9 + j] = Kd[q][9 + j] + Ku;
          const double T1k = (am*tauM*rho)*nfa[j]*Nb + af*Na*nfb[j];
          K[q][12 + j] = K[q][12 + j] + w*J*T1k + afm*Ku;
        }
        // D block (ustruct.cpp:859-871)
        {
          const double T0 = (am*beta + af*dbeta*pd)*Na*Nb;
          const double T1 = nfa[0]*vd[0] + nfa[1]*vd[1] + nfa[2]*vd[2];
          const double T2 = T0 + af*tauM*(NxNx + drho*T1*Nb);
          K[q][15] = K[q][15] + w*J*T2;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < APT; q++) {
      const size_t slot = size_t(kslot[(size_t(e)*ENON + (a0 + q))*ENON + b]);
      double* o = stageK + slot*16;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        d4 t; t.x = K[q][i*4]; t.y = K[q][i*4 + 1]; t.z = K[q][i*4 + 2]; t.w = K[q][i*4 + 3];
        st256_stream(o + 4*i, t);
      }
      double* od = stageKd + slot*12;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        d4 t; t.x = Kd[q][i*4]; t.y = Kd[q][i*4 + 1]; t.z = Kd[q][i*4 + 2]; t.w = Kd[q][i*4 + 3];
        st256_stream(od + 4*i, t);
      }
    }
  }
}

// ---- ustruct_r (ustruct.cpp:1726-1828): Rd = amg Ad - Yg(s:s+2);  R -= (1/am) Kd Rd  -------------------------
// Rd in solver ordering: Rd(:, map[a]) from the assembly-ordered Ad / Yg.
__global__ void k_ustruct_rd(int nNo, int tDof, int s, double amg, const int* __restrict__ map, const double* __restrict__ Ad,
                             const double* __restrict__ Yg, double* __restrict__ Rd)
{
  const size_t n = size_t(nNo)*3;
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t t = size_t(blockIdx.x)*blockDim.x + threadIdx.x; t < n; t += nth) {
    const size_t a = t / 3; const int i = int(t % 3);
    Rd[size_t(map[a])*3 + i] = amg*Ad[a*3 + i] - Yg[a*tDof + s + i];
  }
}
// KU(i,row) = sum_p Kd(i*3 + j, p) Rd(j, col_p), i = 0..3: quad per row, lane i owns component i
__global__ void __launch_bounds__(256)
k_spmv_kd(int nNo, const int* __restrict__ rowPtr, const int* __restrict__ col, const double* __restrict__ Kd,
          const double* __restrict__ Rd, double* __restrict__ KU)
{
  const int lane4 = threadIdx.x & 3;
  const int group = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int ngroups = (gridDim.x*blockDim.x) >> 2;
  for (int row = group; row < nNo; row += ngroups) {
    const int s = __ldg(rowPtr + row), e = __ldg(rowPtr + row + 1);
    double acc = 0.0;
#pragma unroll 4
    for (int p = s; p < e; p++) {
      const int cc = __ldg(col + p);
      const double* k = Kd + (size_t(p)*12 + lane4*3);
      const double* u = Rd + size_t(cc)*3;
      acc = acc + __ldg(k)*__ldg(u) + __ldg(k + 1)*__ldg(u + 1) + __ldg(k + 2)*__ldg(u + 2);
    }
    KU[size_t(row)*4 + lane4] = acc;
  }
}

} // namespace svb200
