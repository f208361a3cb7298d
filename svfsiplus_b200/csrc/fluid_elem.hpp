// fluid_elem.hpp — Gauss-point routines of the Navier-Stokes (VMS, equal-order) element for element types whose
// shape-function gradients vary inside the element (HEX8, TET10).  Replaces, per Gauss point,
//   nn::gnn              Code/Source/solver/nn.cpp:455-541      (Jacobian, metric, dN/dx)
//   nn::gn_nxx           nn.cpp:809-924                         (second derivatives, 6x6 solve via dgesv_)
//   fluid::fluid_3d_m    fluid.cpp:1697-2139                    (momentum residual + K, G blocks)
//   fluid::fluid_3d_c    fluid.cpp:1389-1689                    (continuity residual + D, L blocks)
//   fluid::get_viscosity fluid.cpp:2142-2200
// as called from construct_fluid (fluid.cpp:464-708).
//
// The functions are plain inline code that compiles for the device (included by assembly_fluid_gen.cuh) and for the
// host: tests/hostlogic/fluid_elem_host.cpp instantiates the SAME source on the CPU, in the test tree only, so the
// element arithmetic is checked against the compiled reference without a GPU.  Nothing in the product calls the host
// instantiation.
//
// Reference behaviours that are reproduced on purpose:
//  * construct_fluid evaluates gn_nxx only in its first Gauss loop (momentum, fluid.cpp:617-618); the second loop
//    (continuity, :652-690) re-evaluates gnn per point but keeps Nwxx of the LAST point of loop one.  fluid_point
//    therefore takes the second derivatives of the last Gauss point for its continuity half.
//  * HEX8 has no second-derivative table (nn::get_gn_nxx returns early for HEX8, nn.cpp:166-171), so Nxi2 = 0,
//    the right-hand side of gn_nxx is 0 and Nwxx = 0: NXX = false drops those exact zeros.
#pragma once

#include <cmath>

#if defined(__CUDACC__)
#define SVB_HD __host__ __device__ __forceinline__
#define SVB_HD_NOINL __host__ __device__
#else
#define SVB_HD inline
#define SVB_HD_NOINL inline
#endif

namespace svb200 {

struct FluidConsts {
  double dt, am, af, gam;
  double rho, f[3], Kinv;
  int viscType;
  double mu_i, mu_o, lam, a, n;
  int tDof, mvMsh;
  double w[4];          // TET4 kernel only: Gauss weights (nn_elem_gip.h:501-517)
  double N[4][4];       // TET4 kernel only: N[g][a] (nn_elem_gnn.h:1232-1238)
};

// utils::is_zero(a) with b = 0 (solver/utils.cpp:170-190): relative test against eps.
SVB_HD bool is_zero_d(double v)
{
  const double eps = 2.220446049250313e-16;
  const double a = fabs(v);
  const double nrm = fmax(a, eps);
  return (a/nrm) < 10.0*eps;
}

// fluid::get_viscosity (solver/fluid.cpp:2142-2200)
SVB_HD void viscosity(const FluidConsts& c, double& gamma, double& mu, double& mu_g)
{
  if (c.viscType == 0) {
    mu = c.mu_i; mu_g = 0.0;
  } else if (c.viscType == 1) {
    double T1 = 1.0 + pow(c.lam*gamma, c.a);
    double T2 = pow(T1, (c.n - 1.0)/c.a);
    mu = c.mu_i + (c.mu_o - c.mu_i)*T2;
    T1 = T2/T1;
    T2 = pow(c.lam, c.a) * pow(gamma, c.a - 1.0) * T1;
    mu_g = (c.mu_o - c.mu_i)*(c.n - 1.0)*T2;
  } else {
    double mu_o = c.mu_o;
    if (gamma < c.lam) { mu_o = mu_o/sqrt(c.lam); gamma = c.lam; }
    else               { mu_o = mu_o/sqrt(gamma); }
    mu = (c.mu_i + mu_o)*(c.mu_i + mu_o);
    mu_g = 2.0*mu_o*(mu_o + c.mu_i)/gamma;
  }
}

// Per-(element, Gauss point) record in shared memory (doubles).  Odd size: records of consecutive threads fall
// into different banks.
template <int N, bool NXX> struct FluidRec {
  enum {
    W = 0, MU = 1, MUG = 2, TAUM = 3, TAUC = 4, TAUB = 5, DIVU = 6,
    RV = 7,          // 3   inertia + convection with the full velocity (fluid.cpp:2009-2011)
    RM = 10,         // 9   rM[i][j] (fluid.cpp:1997-2007)
    UUP = 19,        // 3   u + u' (Brinkman residual term, fluid.cpp:2134-2138)
    UPC = 22,        // 3   u' of the continuity form (differs from the momentum one through Nwxx only)
    MUX = 25, D2U = 28,       // 3 + 3  mu_x, d2u2 of the momentum form
    MUXC = 31, D2UC = 34,     // 3 + 3  ... of the continuity form
    KS = 37,         // 9   metric ks = xiX^T xiX
    JAC = 46,
    NX = 48,                  // 3N  dN_a/dx_i at [a*3 + i]
    UNX = NX + 3*N,           // N   u . grad N_a
    UPNX = UNX + N,           // N   u' . grad N_a (momentum form)
    ESNX = UPNX + N,          // 3N  esNx[i][a] at [a*3 + i]
    T1U = ESNX + 3*N,         // N   T1 of updu (momentum form)
    T1UC = T1U + N,           // N   ... (continuity form)
    NXX_ = T1UC + N,          // 6N  second derivatives at [a*6 + k] (NXX only)
    SIZE = (NXX_ + (NXX ? 6*N : 0)) | 1
  };
};

// nn::gnn for insd = 3 (+ the weight w*Jac construct_fluid forms, fluid.cpp:621) and, for NXX, nn::gn_nxx.
// nd[N]: assembly node ids; x(3,nNo); Dmesh != null: ALE configuration x + Dg(4:6) (fsi.cpp:157-163).
// Nxi[a*3+k] = dN_a/dxi_k at this Gauss point, Nxi2[a*6+k] second parametric derivatives (NXX only).
// Returns false when the Jacobian is (relatively) zero (construct_fluid throws, fluid.cpp:612-614).
template <int N, bool NXX>
SVB_HD_NOINL bool fluid_geom(const int* nd, const double* x, const double* Dmesh, int tDof, double wg,
                             const double* Nxi, const double* Nxi2, double* rec)
{
  typedef FluidRec<N, NXX> L;
  double xXi[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  double xXi2[3][6];
  if (NXX) {
    for (int i = 0; i < 3; i++)
      for (int k = 0; k < 6; k++) xXi2[i][k] = 0.0;
  }
  for (int a = 0; a < N; a++) {
    const size_t A = size_t(nd[a]);
    double xa[3];
    for (int i = 0; i < 3; i++) {
      xa[i] = x[A*3 + i];
      if (Dmesh) xa[i] = xa[i] + Dmesh[A*tDof + 4 + i];
    }
    for (int i = 0; i < 3; i++) {
      xXi[i][0] = xXi[i][0] + xa[i]*Nxi[a*3 + 0];
      xXi[i][1] = xXi[i][1] + xa[i]*Nxi[a*3 + 1];
      xXi[i][2] = xXi[i][2] + xa[i]*Nxi[a*3 + 2];
      if (NXX) {
        for (int k = 0; k < 6; k++) xXi2[i][k] = xXi2[i][k] + xa[i]*Nxi2[a*6 + k];
      }
    }
  }
  const double Jac = xXi[0][0]*xXi[1][1]*xXi[2][2] + xXi[0][1]*xXi[1][2]*xXi[2][0] + xXi[0][2]*xXi[1][0]*xXi[2][1]
                   - xXi[0][0]*xXi[1][2]*xXi[2][1] - xXi[0][1]*xXi[1][0]*xXi[2][2] - xXi[0][2]*xXi[1][1]*xXi[2][0];
  const bool ok = !is_zero_d(Jac);
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

  double* ks = rec + L::KS;       // ks[i*3 + j]
  ks[0] = xiX[0][0]*xiX[0][0] + xiX[1][0]*xiX[1][0] + xiX[2][0]*xiX[2][0];
  ks[1] = xiX[0][1]*xiX[0][0] + xiX[1][1]*xiX[1][0] + xiX[2][1]*xiX[2][0];
  ks[2] = xiX[0][2]*xiX[0][0] + xiX[1][2]*xiX[1][0] + xiX[2][2]*xiX[2][0];
  ks[4] = xiX[0][1]*xiX[0][1] + xiX[1][1]*xiX[1][1] + xiX[2][1]*xiX[2][1];
  ks[5] = xiX[0][1]*xiX[0][2] + xiX[1][1]*xiX[1][2] + xiX[2][1]*xiX[2][2];
  ks[8] = xiX[0][2]*xiX[0][2] + xiX[1][2]*xiX[1][2] + xiX[2][2]*xiX[2][2];
  ks[3] = ks[1]; ks[6] = ks[2]; ks[7] = ks[5];
  rec[L::JAC] = Jac;
  rec[L::W] = wg*Jac;

  for (int a = 0; a < N; a++) {
    const double n0 = Nxi[a*3], n1 = Nxi[a*3 + 1], n2 = Nxi[a*3 + 2];
    for (int i = 0; i < 3; i++)
      rec[L::NX + a*3 + i] = ((0.0 + n0*xiX[0][i]) + n1*xiX[1][i]) + n2*xiX[2][i];
  }

  if (NXX) {
    // K X = B (nn.cpp:866-922): rows of K as set there, K(i,j) = Kmat[i][j]
    double Km[6][6];
    for (int i = 0; i < 3; i++) {
      Km[i][0] = xXi[0][i]*xXi[0][i]; Km[i][1] = xXi[1][i]*xXi[1][i]; Km[i][2] = xXi[2][i]*xXi[2][i];
      Km[i][3] = 2.0*xXi[0][i]*xXi[1][i]; Km[i][4] = 2.0*xXi[1][i]*xXi[2][i]; Km[i][5] = 2.0*xXi[0][i]*xXi[2][i];
    }
    const int pi[3] = {0, 1, 0}, pj[3] = {1, 2, 2};
    for (int r = 0; r < 3; r++) {
      const int i = pi[r], j = pj[r];
      Km[3 + r][0] = xXi[0][i]*xXi[0][j];
      Km[3 + r][1] = xXi[1][i]*xXi[1][j];
      Km[3 + r][2] = xXi[2][i]*xXi[2][j];
      Km[3 + r][3] = xXi[0][i]*xXi[1][j] + xXi[0][j]*xXi[1][i];
      Km[3 + r][4] = xXi[1][i]*xXi[2][j] + xXi[1][j]*xXi[2][i];
      Km[3 + r][5] = xXi[0][i]*xXi[2][j] + xXi[0][j]*xXi[2][i];
    }
    // dgesv_: LU with partial pivoting (right-looking, unblocked), then the two triangular solves per column.
    // Row exchanges are written as predicated swaps so that nothing is indexed dynamically.
    int piv[6];
    for (int k = 0; k < 6; k++) {
      int p = k;
      double mx = fabs(Km[k][k]);
      for (int i = k + 1; i < 6; i++) {
        const double v = fabs(Km[i][k]);
        if (v > mx) { mx = v; p = i; }
      }
      piv[k] = p;
      for (int i = k + 1; i < 6; i++) {
        if (p == i) {
          for (int j = 0; j < 6; j++) { const double t = Km[k][j]; Km[k][j] = Km[i][j]; Km[i][j] = t; }
        }
      }
      for (int i = k + 1; i < 6; i++) Km[i][k] = Km[i][k]/Km[k][k];
      for (int j = k + 1; j < 6; j++) {
        const double akj = Km[k][j];
        for (int i = k + 1; i < 6; i++) Km[i][j] = Km[i][j] - Km[i][k]*akj;
      }
    }
    for (int a = 0; a < N; a++) {
      double b[6];
      const double n0 = rec[L::NX + a*3], n1 = rec[L::NX + a*3 + 1], n2 = rec[L::NX + a*3 + 2];
      for (int i = 0; i < 6; i++) b[i] = Nxi2[a*6 + i] - n0*xXi2[0][i] - n1*xXi2[1][i] - n2*xXi2[2][i];
      for (int k = 0; k < 6; k++) {
        for (int i = k + 1; i < 6; i++) {
          if (piv[k] == i) { const double t = b[k]; b[k] = b[i]; b[i] = t; }
        }
      }
      for (int k = 0; k < 6; k++)
        for (int i = k + 1; i < 6; i++) b[i] = b[i] - Km[i][k]*b[k];
      for (int k = 5; k >= 0; k--) {
        b[k] = b[k]/Km[k][k];
        for (int i = 0; i < k; i++) b[i] = b[i] - Km[i][k]*b[k];
      }
      for (int i = 0; i < 6; i++) rec[L::NXX_ + a*6 + i] = b[i];
    }
  }
  return ok;
}

// The second-derivative dependent tail of the preamble shared by fluid_3d_m and fluid_3d_c: from uxx (through d2u2,
// es_x, mu_x) to rS and the fine-scale velocity u' (fluid.cpp:1778-1800, 1850-1880, 1945-1956).
SVB_HD void fluid_fine_scale(const double uxxs[3][6], const double es[3][3], double mu, double mu_g, double rho, double tauM,
                             double muK, const double u[3], const double rVm[3], const double px[3],
                             double d2u2[3], double mu_x[3], double up[3])
{
  // uxx[i][j][k]: i,k spatial directions of the second derivative, j the velocity component; the six
  // independent (i,k) pairs are filled from Nwxx rows 0..5 = (00, 11, 22, 10, 21, 02), the others mirrored.
  double uxx[3][3][3];
  for (int j = 0; j < 3; j++) {
    uxx[0][j][0] = uxxs[j][0];
    uxx[1][j][1] = uxxs[j][1];
    uxx[2][j][2] = uxxs[j][2];
    uxx[1][j][0] = uxxs[j][3];
    uxx[2][j][1] = uxxs[j][4];
    uxx[0][j][2] = uxxs[j][5];
    uxx[0][j][1] = uxx[1][j][0];
    uxx[1][j][2] = uxx[2][j][1];
    uxx[2][j][0] = uxx[0][j][2];
  }
  d2u2[0] = uxx[0][0][0] + uxx[1][0][1] + uxx[2][0][2];
  d2u2[1] = uxx[0][1][0] + uxx[1][1][1] + uxx[2][1][2];
  d2u2[2] = uxx[0][2][0] + uxx[1][2][1] + uxx[2][2][2];
  double es_x[3][3][3];
  for (int k = 0; k < 3; k++) {
    es_x[0][0][k] = uxx[0][0][k] + uxx[0][0][k];
    es_x[1][1][k] = uxx[1][1][k] + uxx[1][1][k];
    es_x[2][2][k] = uxx[2][2][k] + uxx[2][2][k];
    es_x[1][0][k] = uxx[1][0][k] + uxx[0][1][k];
    es_x[2][1][k] = uxx[2][1][k] + uxx[1][2][k];
    es_x[0][2][k] = uxx[0][2][k] + uxx[2][0][k];
  }
  for (int k = 0; k < 3; k++) {
    mu_x[k] = (es_x[0][0][k]*es[0][0] + es_x[1][1][k]*es[1][1] + es_x[2][2][k]*es[2][2])*0.5
            + es_x[1][0][k]*es[1][0] + es_x[2][1][k]*es[2][1] + es_x[0][2][k]*es[0][2];
  }
  for (int k = 0; k < 3; k++) mu_x[k] = mu_g*mu_x[k];
  double rS[3];
  for (int j = 0; j < 3; j++) rS[j] = mu_x[0]*es[0][j] + mu_x[1]*es[1][j] + mu_x[2]*es[2][j] + mu*d2u2[j];
  for (int j = 0; j < 3; j++) up[j] = -tauM*(rho*rVm[j] + px[j] - rS[j] + muK*u[j]);
}

// fluid_3d_m + fluid_3d_c preambles at one Gauss point.  rec holds the geometry of this point (fluid_geom);
// recLast the record of the element's LAST Gauss point (its NXX_ field feeds the continuity form, see header).
// Ng[a] = N_a at this point.  Nodal fields are gathered on the fly (no per-thread element arrays).
template <int N, bool NXX>
SVB_HD_NOINL void fluid_point(const FluidConsts& c, const int* nd, const double* Ag, const double* Yg, const double* Bf,
                              const double* Ng, double* rec, const double* recLast)
{
  typedef FluidRec<N, NXX> L;
  const int tD = c.tDof;
  const double rho = c.rho;
  const double* Nx = rec + L::NX;
  const double* ks = rec + L::KS;

  double ud[3] = {-c.f[0], -c.f[1], -c.f[2]};
  double u[3] = {0.0, 0.0, 0.0};
  double ux[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};      // ux[i][j] = d u_j / d x_i
  double uxxm[3][6], uxxc[3][6];                            // [component][Nwxx row]
  double p = 0.0, px[3] = {0.0, 0.0, 0.0};
  if (NXX) {
    for (int j = 0; j < 3; j++)
      for (int k = 0; k < 6; k++) { uxxm[j][k] = 0.0; uxxc[j][k] = 0.0; }
  }
  for (int a = 0; a < N; a++) {
    const size_t A = size_t(nd[a]);
    const double Na = Ng[a];
    double yl[4];
    for (int i = 0; i < 3; i++) {
      const double al = Ag[A*tD + i], bl = Bf[A*3 + i];
      yl[i] = Yg[A*tD + i];
      ud[i] = ud[i] + Na*(al - bl);
      u[i] = u[i] + Na*yl[i];
    }
    yl[3] = Yg[A*tD + 3];
    for (int j = 0; j < 3; j++) {
      ux[0][j] += Nx[a*3 + 0]*yl[j];
      ux[1][j] += Nx[a*3 + 1]*yl[j];
      ux[2][j] += Nx[a*3 + 2]*yl[j];
    }
    if (NXX) {
      for (int j = 0; j < 3; j++)
        for (int k = 0; k < 6; k++) {
          uxxm[j][k] += rec[L::NXX_ + a*6 + k]*yl[j];
          uxxc[j][k] += recLast[L::NXX_ + a*6 + k]*yl[j];
        }
    }
    p = p + Na*yl[3];
    px[0] = px[0] + Nx[a*3 + 0]*yl[3];
    px[1] = px[1] + Nx[a*3 + 1]*yl[3];
    px[2] = px[2] + Nx[a*3 + 2]*yl[3];
  }
  const double divU = ux[0][0] + ux[1][1] + ux[2][2];
  if (c.mvMsh) {        // convection velocity relative to the mesh velocity (fluid.cpp:1837-1843)
    for (int a = 0; a < N; a++) {
      const size_t A = size_t(nd[a]);
      for (int i = 0; i < 3; i++) u[i] = u[i] - Ng[a]*Yg[A*tD + 4 + i];
    }
  }

  double es[3][3];
  es[0][0] = ux[0][0] + ux[0][0];
  es[1][1] = ux[1][1] + ux[1][1];
  es[2][2] = ux[2][2] + ux[2][2];
  es[1][0] = ux[1][0] + ux[0][1];
  es[2][1] = ux[2][1] + ux[1][2];
  es[0][2] = ux[0][2] + ux[2][0];
  es[0][1] = es[1][0]; es[1][2] = es[2][1]; es[2][0] = es[0][2];

  double gam = es[0][0]*es[0][0] + es[1][0]*es[1][0] + es[2][0]*es[2][0]
             + es[0][1]*es[0][1] + es[1][1]*es[1][1] + es[2][1]*es[2][1]
             + es[0][2]*es[0][2] + es[1][2]*es[1][2] + es[2][2]*es[2][2];
  gam = sqrt(0.5*gam);
  double mu, mu_g;
  viscosity(c, gam, mu, mu_g);
  if (is_zero_d(gam)) mu_g = 0.0; else mu_g = mu_g/gam;

  const double muK = mu*c.Kinv;
  double kT = 4.0*((1.0/c.dt)*(1.0/c.dt));
  {
    const double t = c.Kinv*mu/rho;
    kT = kT + t*t;
  }
  const double kU = u[0]*u[0]*ks[0] + u[1]*u[0]*ks[3] + u[2]*u[0]*ks[6]
                  + u[0]*u[1]*ks[1] + u[1]*u[1]*ks[4] + u[2]*u[1]*ks[7]
                  + u[0]*u[2]*ks[2] + u[1]*u[2]*ks[5] + u[2]*u[2]*ks[8];
  double kS = ks[0]*ks[0] + ks[3]*ks[3] + ks[6]*ks[6]
            + ks[1]*ks[1] + ks[4]*ks[4] + ks[7]*ks[7]
            + ks[2]*ks[2] + ks[5]*ks[5] + ks[8]*ks[8];
  {
    const double t = mu/rho;
    kS = 36.0*kS*(t*t);
  }
  const double tauM = 1.0/(rho*sqrt(kT + kU + kS));

  double rV[3];
  for (int j = 0; j < 3; j++) rV[j] = ud[j] + u[0]*ux[0][j] + u[1]*ux[1][j] + u[2]*ux[2][j];

  double up[3], upc[3];
  double d2u2[3] = {0.0, 0.0, 0.0}, mu_x[3] = {0.0, 0.0, 0.0};
  double d2u2c[3] = {0.0, 0.0, 0.0}, mu_xc[3] = {0.0, 0.0, 0.0};
  if (NXX) {
    fluid_fine_scale(uxxm, es, mu, mu_g, rho, tauM, muK, u, rV, px, d2u2, mu_x, up);
    fluid_fine_scale(uxxc, es, mu, mu_g, rho, tauM, muK, u, rV, px, d2u2c, mu_xc, upc);
  } else {
    for (int j = 0; j < 3; j++) up[j] = -tauM*(rho*rV[j] + px[j] - 0.0 + muK*u[j]);
    for (int j = 0; j < 3; j++) upc[j] = up[j];
  }

  const double tauC = 1.0/(tauM*(ks[0] + ks[4] + ks[8]));
  double tauB = up[0]*up[0]*ks[0] + up[1]*up[0]*ks[3] + up[2]*up[0]*ks[6]
              + up[0]*up[1]*ks[1] + up[1]*up[1]*ks[4] + up[2]*up[1]*ks[7]
              + up[0]*up[2]*ks[2] + up[1]*up[2]*ks[5] + up[2]*up[2]*ks[8];
  if (is_zero_d(tauB)) tauB = 2.220446049250313e-16;
  tauB = rho/sqrt(tauB);
  const double ua[3] = {u[0] + up[0], u[1] + up[1], u[2] + up[2]};
  const double pa = p - tauC*divU;

  for (int j = 0; j < 3; j++) rV[j] = tauB*(up[0]*ux[0][j] + up[1]*ux[1][j] + up[2]*ux[2][j]);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double t = mu*es[i][j] - rho*up[j]*ua[i] + rV[j]*up[i];
      if (i == j) t = t - pa;
      rec[L::RM + i*3 + j] = t;
    }
  for (int j = 0; j < 3; j++) rec[L::RV + j] = ud[j] + ua[0]*ux[0][j] + ua[1]*ux[1][j] + ua[2]*ux[2][j];

  rec[L::MU] = mu; rec[L::MUG] = mu_g; rec[L::TAUM] = tauM; rec[L::TAUC] = tauC; rec[L::TAUB] = tauB; rec[L::DIVU] = divU;
  for (int j = 0; j < 3; j++) {
    rec[L::UUP + j] = u[j] + up[j];
    rec[L::UPC + j] = upc[j];
    rec[L::MUX + j] = mu_x[j]; rec[L::D2U + j] = d2u2[j];
    rec[L::MUXC + j] = mu_xc[j]; rec[L::D2UC + j] = d2u2c[j];
  }

  for (int a = 0; a < N; a++) {
    const double n0 = Nx[a*3], n1 = Nx[a*3 + 1], n2 = Nx[a*3 + 2];
    rec[L::ESNX + a*3 + 0] = es[0][0]*n0 + es[1][0]*n1 + es[2][0]*n2;
    rec[L::ESNX + a*3 + 1] = es[0][1]*n0 + es[1][1]*n1 + es[2][1]*n2;
    rec[L::ESNX + a*3 + 2] = es[0][2]*n0 + es[1][2]*n1 + es[2][2]*n2;
    const double uNx = u[0]*n0 + u[1]*n1 + u[2]*n2;
    rec[L::UNX + a] = uNx;
    rec[L::UPNX + a] = up[0]*n0 + up[1]*n1 + up[2]*n2;
    if (NXX) {
      const double* q = rec + L::NXX_ + a*6;
      const double* ql = recLast + L::NXX_ + a*6;
      rec[L::T1U + a] = -rho*uNx + mu*(q[0] + q[1] + q[2]) + mu_x[0]*n0 + mu_x[1]*n1 + mu_x[2]*n2 - muK*Ng[a];
      rec[L::T1UC + a] = -rho*uNx + mu*(ql[0] + ql[1] + ql[2]) + mu_xc[0]*n0 + mu_xc[1]*n1 + mu_xc[2]*n2 - muK*Ng[a];
    } else {
      const double t = -rho*uNx + mu*(0.0) - muK*Ng[a];
      rec[L::T1U + a] = t;
      rec[L::T1UC + a] = t;
    }
  }
}

// lR(:,a): Gauss points summed in the reference's order (per point: momentum residual, then the Brinkman term,
// fluid.cpp:2025-2028, 2134-2138; continuity fluid.cpp:1655-1658).  recs: the element's NG records, Ntab[g*N + a].
template <int N, int NG, bool NXX>
SVB_HD void fluid_res_row(const FluidConsts& c, const double* recs, const double* Ntab, int a, double out[4])
{
  typedef FluidRec<N, NXX> L;
  double r[4] = {0.0, 0.0, 0.0, 0.0};
  for (int g = 0; g < NG; g++) {
    const double* rec = recs + size_t(g)*L::SIZE;
    const double w = rec[L::W], wr = w*c.rho;
    const double Na = Ntab[g*N + a];
    const double n0 = rec[L::NX + a*3], n1 = rec[L::NX + a*3 + 1], n2 = rec[L::NX + a*3 + 2];
    const double* rM = rec + L::RM;
    const double muK = rec[L::MU]*c.Kinv;
    for (int j = 0; j < 3; j++) {
      r[j] = r[j] + wr*Na*rec[L::RV + j] + w*(n0*rM[j] + n1*rM[3 + j] + n2*rM[6 + j]);
      r[j] = r[j] + muK*w*Na*rec[L::UUP + j];
    }
    const double upn = NXX ? (rec[L::UPC]*n0 + rec[L::UPC + 1]*n1 + rec[L::UPC + 2]*n2) : rec[L::UPNX + a];
    r[3] = r[3] + w*(Na*rec[L::DIVU] - upn);
  }
  out[0] = r[0]; out[1] = r[1]; out[2] = r[2]; out[3] = r[3];
}

// lK(:,a,b) (row-major 4x4): fluid_3d_m K and G blocks (fluid.cpp:2058-2130), fluid_3d_c D and L blocks
// (fluid.cpp:1662-1688), Gauss points summed in order.
template <int N, int NG, bool NXX>
SVB_HD void fluid_tan_block(const FluidConsts& c, const double* recs, const double* Ntab, int a, int b, double kb[16])
{
  typedef FluidRec<N, NXX> L;
  const double rho = c.rho;
  const double T1c = c.af*c.gam*c.dt;
  const double amd = c.am/T1c;
  for (int i = 0; i < 16; i++) kb[i] = 0.0;
  for (int g = 0; g < NG; g++) {
    const double* rec = recs + size_t(g)*L::SIZE;
    const double wl = rec[L::W]*T1c;
    const double mu = rec[L::MU], mu_g = rec[L::MUG], tauM = rec[L::TAUM], tauC = rec[L::TAUC], tauB = rec[L::TAUB];
    const double muK = mu*c.Kinv;
    const double Na = Ntab[g*N + a], Nb = Ntab[g*N + b];
    const double Nxa[3] = {rec[L::NX + a*3], rec[L::NX + a*3 + 1], rec[L::NX + a*3 + 2]};
    const double Nxb[3] = {rec[L::NX + b*3], rec[L::NX + b*3 + 1], rec[L::NX + b*3 + 2]};
    const double esa[3] = {rec[L::ESNX + a*3], rec[L::ESNX + a*3 + 1], rec[L::ESNX + a*3 + 2]};
    const double esb[3] = {rec[L::ESNX + b*3], rec[L::ESNX + b*3 + 1], rec[L::ESNX + b*3 + 2]};
    const double uNxb = rec[L::UNX + b], upNxa = rec[L::UPNX + a], upNxb = rec[L::UPNX + b];
    const double uaNxa = rec[L::UNX + a] + upNxa;
    const double rtu = rho*tauM*uaNxa;
    const double T1ub = rec[L::T1U + b], T1ucb = rec[L::T1UC + b];

    const double NxNx = Nxa[0]*Nxb[0] + Nxa[1]*Nxb[1] + Nxa[2]*Nxb[2];
    const double T1 = mu*NxNx + rho*amd*Nb*(Na + rho*tauM*uaNxa) + rho*Na*(uNxb + upNxb) + tauB*upNxa*upNxb;
    // momentum-velocity block.  updu[i][j][b] = mu_x[i]*Nwx(j,b) + d2u2[i]*mu_g*esNx[j][b] (+ T1u[b] if i == j)
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) {
        if (i == j) {
          double updu;
          if (NXX) updu = rec[L::MUX + i]*Nxb[i] + rec[L::D2U + i]*mu_g*esb[i] + T1ub;
          else updu = T1ub;
          const double T2 = (mu + tauC)*(Nxa[i]*Nxb[i]) + esa[i]*mu_g*esb[i] - rtu*updu;
          kb[i*4 + i] = kb[i*4 + i] + wl*(T2 + T1);
          kb[i*4 + i] = kb[i*4 + i] + muK*wl*Nb*Na;
        } else {
          double T2 = mu*(Nxa[j]*Nxb[i]) + tauC*(Nxa[i]*Nxb[j]) + esa[i]*mu_g*esb[j];
          if (NXX) T2 = T2 - rtu*(rec[L::MUX + j]*Nxb[i] + rec[L::D2U + j]*mu_g*esb[i]);
          kb[i*4 + j] = kb[i*4 + j] + wl*T2;
        }
      }
    }
    // momentum-pressure block
    for (int i = 0; i < 3; i++) kb[i*4 + 3] = kb[i*4 + 3] - wl*(Nxa[i]*Nb - Nxb[i]*rtu);
    // continuity-velocity block: lK(12+i): T2 = sum_j Nqx(j,a)*(updu_c[i][j][b] - delta_ij T1cc)
    {
      const double T1cc = rho*amd*Nb;
      for (int i = 0; i < 3; i++) {
        double T2;
        if (NXX) {
          double t[3];
          for (int j = 0; j < 3; j++) {
            double v = rec[L::MUXC + i]*Nxb[j] + rec[L::D2UC + i]*mu_g*esb[j];
            if (i == j) v = (v + T1ucb) - T1cc;
            t[j] = Nxa[j]*v;
          }
          T2 = t[0] + t[1] + t[2];
        } else {
          T2 = Nxa[i]*(T1ucb - T1cc);
        }
        kb[12 + i] = kb[12 + i] + wl*(Na*Nxb[i] - tauM*T2);
      }
    }
    // continuity-pressure block
    kb[15] = kb[15] + wl*tauM*NxNx;
  }
}

} // namespace svb200
