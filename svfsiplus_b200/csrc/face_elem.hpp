// face_elem.hpp — boundary-face (Neumann) element routines, host/device shared like fluid_elem.hpp:
//   nn::gnnb               Code/Source/solver/nn.cpp:552-755    area-weighted outward normal at a face Gauss point
//   b_assem_neu_bc         eq_assem.cpp:58-170                  interpolation of h and y, weights
//   fluid::b_fluid         fluid.cpp:46-133                     traction + backflow stabilisation (residual and tangent)
//   l_elas::b_l_elas       l_elas.cpp:48-59                     traction on struct / lElas / ustruct / mesh (residual only)
// and the face Gauss tables of nn::select_eleb (nn_elem_gip.h:664,747,767; nn_elem_gnn.h:1536,1561,1578).
// SURVEY.md par. 8(f) row 1.  tests/hostlogic/fluid_elem_host.cpp instantiates the same source on the CPU (test tree only).
#pragma once

#include <cmath>
#include <cstring>

#include "fluid_elem.hpp"

namespace svb200 {

struct FaceTables {            // lFa.w, lFa.N, lFa.Nx
  enum { MAXN = 6, MAXG = 7 };
  int eNoNb, nG;
  double w[MAXG];
  double N[MAXG][MAXN];        // [g][a]
  double Nx[MAXG][MAXN][2];    // [g][a][i]
};

inline bool face_supported(int eNoNb) { return eNoNb == 3 || eNoNb == 4 || eNoNb == 6; }

inline void fill_face_tables(FaceTables& t, int eNoNb, double qmTRI3)
{
  std::memset(&t, 0, sizeof(t));
  t.eNoNb = eNoNb;
  if (eNoNb == 3) {                      // TRI3
    t.nG = 3;
    const double s = qmTRI3, q = -0.5*s + 0.5;
    const double xi[3][2] = {{q, q}, {s, q}, {q, s}};
    for (int g = 0; g < 3; g++) {
      t.w[g] = 1.0/6.0;
      t.N[g][0] = xi[g][0]; t.N[g][1] = xi[g][1]; t.N[g][2] = 1.0 - xi[g][0] - xi[g][1];
      const double d[3][2] = {{1.0, 0.0}, {0.0, 1.0}, {-1.0, -1.0}};
      for (int a = 0; a < 3; a++) { t.Nx[g][a][0] = d[a][0]; t.Nx[g][a][1] = d[a][1]; }
    }
  } else if (eNoNb == 4) {               // QUD4
    t.nG = 4;
    const double s = 1.0/std::sqrt(3.0);
    const double xi[4][2] = {{-s, -s}, {s, -s}, {s, s}, {-s, s}};
    for (int g = 0; g < 4; g++) {
      t.w[g] = 1.0;
      const double lx = 1.0 - xi[g][0], ly = 1.0 - xi[g][1], ux = 1.0 + xi[g][0], uy = 1.0 + xi[g][1];
      t.N[g][0] = lx*ly/4.0; t.N[g][1] = ux*ly/4.0; t.N[g][2] = ux*uy/4.0; t.N[g][3] = lx*uy/4.0;
      const double d[4][2] = {{-ly/4.0, -lx/4.0}, {ly/4.0, -ux/4.0}, {uy/4.0, ux/4.0}, {-uy/4.0, lx/4.0}};
      for (int a = 0; a < 4; a++) { t.Nx[g][a][0] = d[a][0]; t.Nx[g][a][1] = d[a][1]; }
    }
  } else {                               // TRI6
    t.nG = 7;
    const double wt[7] = {0.225000000000000*5.0e-1, 0.125939180544827*5.0e-1, 0.125939180544827*5.0e-1, 0.125939180544827*5.0e-1,
                          0.132394152788506*5.0e-1, 0.132394152788506*5.0e-1, 0.132394152788506*5.0e-1};
    double xi[7][2];
    {
      double s = 0.333333333333333;
      xi[0][0] = s; xi[0][1] = s;
      s = 0.797426985353087;
      double q = 0.101286507323456;
      xi[1][0] = s; xi[1][1] = q;
      xi[2][0] = q; xi[2][1] = s;
      xi[3][0] = q; xi[3][1] = q;
      s = 0.059715871789770; q = 0.470142064105115;
      xi[4][0] = s; xi[4][1] = q;
      xi[5][0] = q; xi[5][1] = s;
      xi[6][0] = q; xi[6][1] = q;
    }
    for (int g = 0; g < 7; g++) {
      t.w[g] = wt[g];
      const double x0 = xi[g][0], x1 = xi[g][1], s = 1.0 - x0 - x1;
      t.N[g][0] = x0*(2.0*x0 - 1.0); t.N[g][1] = x1*(2.0*x1 - 1.0); t.N[g][2] = s*(2.0*s - 1.0);
      t.N[g][3] = 4.0*x0*x1; t.N[g][4] = 4.0*x1*s; t.N[g][5] = 4.0*x0*s;
      const double d[6][2] = {{4.0*x0 - 1.0, 0.0}, {0.0, 4.0*x1 - 1.0}, {1.0 - 4.0*s, 1.0 - 4.0*s}, {4.0*x1, 4.0*x0},
                              {-4.0*x1, 4.0*(s - x1)}, {4.0*(s - x0), -4.0*x0}};
      for (int a = 0; a < 6; a++) { t.Nx[g][a][0] = d[a][0]; t.Nx[g][a][1] = d[a][1]; }
    }
  }
}

struct BneuConsts {
  double dt, af, gam;          // com_mod.dt, eq.af, eq.gam
  int tDof, mvMsh;
  double rho, bfs;             // fluid_density, backflow_stab of the domain (kind 0)
  int kind;                    // 0 b_fluid, 1 b_l_elas
  int dof;                     // block size of the system (4 fluid / FSI / ustruct, 3 struct / lElas / mesh)
};

// One face element of NB nodes with NG Gauss points.  nd[NB]: face nodes (assembly ids), inode: a node of the parent
// volume element that is not on the face (gnnb's ptr(eNoNb), orientation only).  x(3,nNo); Do != null: moving mesh,
// geometry x + Do(4:6) (nn.cpp:615-620).  hg(nNo) nodal Neumann values, Yg(tDof,nNo).
// Outputs: lR[a*3 + i] (rows 0..2; the pressure row of a dof-4 system gets nothing) and, for kind 0, lKd[a*NB + b] =
// the common value of lK(0,a,b) = lK(5,a,b) = lK(10,a,b).
template <int NB, int NG>
SVB_HD_NOINL void face_element(const BneuConsts& c, const int* nd, int inode, const double* x, const double* Do, const double* hg,
                               const double* Yg, const double* wtab, const double* Ntab /*[g][a]*/, const double* Nxtab /*[g][a][2]*/,
                               double* lR, double* lKd)
{
  const int tD = c.tDof;
  double lX[NB][3], hl[NB];
  for (int a = 0; a < NB; a++) {
    const size_t A = size_t(nd[a]);
    for (int i = 0; i < 3; i++) {
      lX[a][i] = x[A*3 + i];
      if (c.mvMsh) lX[a][i] = lX[a][i] + Do[A*tD + 4 + i];
    }
    hl[a] = hg[A];
  }
  double xin[3];
  for (int i = 0; i < 3; i++) {
    xin[i] = x[size_t(inode)*3 + i];
    if (c.mvMsh) xin[i] = xin[i] + Do[size_t(inode)*tD + 4 + i];
  }
  for (int a = 0; a < NB; a++) {
    lR[a*3] = 0.0; lR[a*3 + 1] = 0.0; lR[a*3 + 2] = 0.0;
    if (c.kind == 0) for (int b = 0; b < NB; b++) lKd[a*NB + b] = 0.0;
  }
  for (int g = 0; g < NG; g++) {
    // gnnb: xXi(j,i) = sum_a Nx(i,a) lX(j,a); n = cross(xXi); outward: n.(x_a0 - x_interior) >= 0
    double xXi[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    for (int a = 0; a < NB; a++)
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 3; j++) xXi[j][i] = xXi[j][i] + Nxtab[(g*NB + a)*2 + i]*lX[a][j];
    double nV[3];
    nV[0] = xXi[1][0]*xXi[2][1] - xXi[2][0]*xXi[1][1];
    nV[1] = xXi[2][0]*xXi[0][1] - xXi[0][0]*xXi[2][1];
    nV[2] = xXi[0][0]*xXi[1][1] - xXi[1][0]*xXi[0][1];
    {
      double dotv = 0.0;
      for (int i = 0; i < 3; i++) dotv += nV[i]*(lX[0][i] - xin[i]);
      if (dotv < 0.0) { nV[0] = -nV[0]; nV[1] = -nV[1]; nV[2] = -nV[2]; }
    }
    double nn = 0.0;
    for (int i = 0; i < 3; i++) nn += nV[i]*nV[i];
    const double Jac = sqrt(nn);
    for (int i = 0; i < 3; i++) nV[i] = nV[i]/Jac;
    const double w = wtab[g]*Jac;
    const double* N = Ntab + g*NB;
    double h = 0.0;
    for (int a = 0; a < NB; a++) h = h + N[a]*hl[a];
    if (c.kind == 0) {
      // b_fluid
      double y[7] = {0, 0, 0, 0, 0, 0, 0};
      const int ny = c.mvMsh ? 7 : 3;
      for (int a = 0; a < NB; a++) {
        const size_t A = size_t(nd[a]);
        for (int i = 0; i < ny; i++) y[i] = y[i] + N[a]*Yg[A*tD + i];
      }
      const double wl = w*c.af*c.gam*c.dt;
      double udn = 0.0, u[3];
      for (int i = 0; i < 3; i++) {
        u[i] = c.mvMsh ? (y[i] - y[i + 4]) : y[i];
        udn = udn + u[i]*nV[i];
      }
      udn = 0.50*c.bfs*c.rho*(udn - fabs(udn));
      double hc[3];
      for (int i = 0; i < 3; i++) hc[i] = h*nV[i] + udn*u[i];
      for (int a = 0; a < NB; a++) {
        for (int i = 0; i < 3; i++) lR[a*3 + i] = lR[a*3 + i] - w*N[a]*hc[i];
        for (int b = 0; b < NB; b++) {
          const double T1 = wl*N[a]*N[b]*udn;
          lKd[a*NB + b] = lKd[a*NB + b] - T1;
        }
      }
    } else {
      // b_l_elas
      double hc[3];
      for (int i = 0; i < 3; i++) hc[i] = h*nV[i];
      for (int a = 0; a < NB; a++)
        for (int i = 0; i < 3; i++) lR[a*3 + i] = lR[a*3 + i] - w*N[a]*hc[i];
    }
  }
}

// all_fun::integ over a face (Code/Source/solver/all_fun.cpp:561-722 scalar, :724-856 vector flux): the terms one face
// element adds to the running sum, one per Gauss point, in the reference's order.
//   vector (nrow == 3): w(g) * sum_a sum_i N(a,g) s(l+i, Ac) n(i)        n = gnnb's area-weighted normal (not normalised)
//   scalar (nrow == 1): |n| * w(g) * sum_a s(l, Ac) N(a,g)               s == null: integrand 1 (the face area)
// geo != null: geometry x + geo(goff : goff+3, node) (gnnb: Do(4:6) when the mesh moves, Do / Dn(0:2) for the old / new
// configuration, nn.cpp:609-640).
template <int NB, int NG>
SVB_HD_NOINL void face_integ_terms(const int* nd, int inode, const double* x, const double* geo, int gtD, int goff,
                                   const double* s, int stD, int l, int nrow, const double* wtab, const double* Ntab,
                                   const double* Nxtab, double* terms)
{
  double lX[NB][3], xin[3];
  for (int a = 0; a < NB; a++) {
    const size_t A = size_t(nd[a]);
    for (int i = 0; i < 3; i++) {
      lX[a][i] = x[A*3 + i];
      if (geo) lX[a][i] = lX[a][i] + geo[A*gtD + goff + i];
    }
  }
  for (int i = 0; i < 3; i++) {
    xin[i] = x[size_t(inode)*3 + i];
    if (geo) xin[i] = xin[i] + geo[size_t(inode)*gtD + goff + i];
  }
  for (int g = 0; g < NG; g++) {
    double xXi[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    for (int a = 0; a < NB; a++)
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 3; j++) xXi[j][i] = xXi[j][i] + Nxtab[(g*NB + a)*2 + i]*lX[a][j];
    double n[3];
    n[0] = xXi[1][0]*xXi[2][1] - xXi[2][0]*xXi[1][1];
    n[1] = xXi[2][0]*xXi[0][1] - xXi[0][0]*xXi[2][1];
    n[2] = xXi[0][0]*xXi[1][1] - xXi[1][0]*xXi[0][1];
    double dotv = 0.0;
    for (int i = 0; i < 3; i++) dotv += n[i]*(lX[0][i] - xin[i]);
    if (dotv < 0.0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
    const double* N = Ntab + g*NB;
    double sHat = 0.0;
    if (nrow == 3) {
      for (int a = 0; a < NB; a++) {
        const size_t A = size_t(nd[a]);
        for (int i = 0; i < 3; i++) sHat = sHat + N[a]*s[A*stD + l + i]*n[i];
      }
      terms[g] = wtab[g]*sHat;
    } else {
      double nn = 0.0;
      for (int i = 0; i < 3; i++) nn += n[i]*n[i];
      const double Jac = sqrt(nn);
      for (int a = 0; a < NB; a++) sHat = sHat + (s ? s[size_t(nd[a])*stD + l] : 1.0)*N[a];
      terms[g] = Jac*wtab[g]*sHat;
    }
  }
}

// eq_assem::fsi_ls_upd (Code/Source/solver/eq_assem.cpp:316-371): the terms N(a,g) w(g) n(i) one face element adds to
// sV(i, node a) = int N_a n_i dGamma on the configuration `geo` (new time step there; gnnb's area-weighted normal).
// out[(a*NG + g)*3 + i]; the caller adds them per node in (element, Gauss point) order.
template <int NB, int NG>
SVB_HD_NOINL void face_normal_terms(const int* nd, int inode, const double* x, const double* geo, int gtD, int goff,
                                    const double* wtab, const double* Ntab, const double* Nxtab, double* out)
{
  double lX[NB][3], xin[3];
  for (int a = 0; a < NB; a++) {
    const size_t A = size_t(nd[a]);
    for (int i = 0; i < 3; i++) {
      lX[a][i] = x[A*3 + i];
      if (geo) lX[a][i] = lX[a][i] + geo[A*gtD + goff + i];
    }
  }
  for (int i = 0; i < 3; i++) {
    xin[i] = x[size_t(inode)*3 + i];
    if (geo) xin[i] = xin[i] + geo[size_t(inode)*gtD + goff + i];
  }
  for (int g = 0; g < NG; g++) {
    double xXi[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    for (int a = 0; a < NB; a++)
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 3; j++) xXi[j][i] = xXi[j][i] + Nxtab[(g*NB + a)*2 + i]*lX[a][j];
    double n[3];
    n[0] = xXi[1][0]*xXi[2][1] - xXi[2][0]*xXi[1][1];
    n[1] = xXi[2][0]*xXi[0][1] - xXi[0][0]*xXi[2][1];
    n[2] = xXi[0][0]*xXi[1][1] - xXi[1][0]*xXi[0][1];
    double dotv = 0.0;
    for (int i = 0; i < 3; i++) dotv += n[i]*(lX[0][i] - xin[i]);
    if (dotv < 0.0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
    for (int a = 0; a < NB; a++)
      for (int i = 0; i < 3; i++) out[(a*NG + g)*3 + i] = Ntab[g*NB + a]*wtab[g]*n[i];
  }
}

} // namespace svb200
