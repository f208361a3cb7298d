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
#include "elem_tables.hpp"

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

// ---- follower pressure load on a struct face: eq_assem::b_neu_folw_p + nn::get_nnx / get_xi + struct_ns::b_struct_3d -----
// (Code/Source/solver/eq_assem.cpp:186-303, nn.cpp:314-440, sv_struct.cpp:116-210).  The pressure h acts along the CURRENT
// normal: Nanson's formula da n = J dA F^-T N with F from the parent element's displacement, so the face integral needs
// the parent's shape functions at the face Gauss point (found by a Newton inverse map) and has a tangent.

// parent-element shape functions and their parametric derivatives at xi (nn_elem_gnn.h:40, 547, 572)
template <int NP>
SVB_HD void parent_shape(const double xi[3], double* N, double* Nxi /*[a*3 + k]*/)
{
  if (NP == 4) {
    N[0] = xi[0]; N[1] = xi[1]; N[2] = xi[2]; N[3] = 1.0 - xi[0] - xi[1] - xi[2];
    const double d[4][3] = {{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}, {-1.0, -1.0, -1.0}};
    for (int a = 0; a < 4; a++) for (int k = 0; k < 3; k++) Nxi[a*3 + k] = d[a][k];
  } else if (NP == 8) {
    const double lx = 1.0 - xi[0], ly = 1.0 - xi[1], lz = 1.0 - xi[2], ux = 1.0 + xi[0], uy = 1.0 + xi[1], uz = 1.0 + xi[2];
    N[0] = lx*ly*lz/8.0; N[1] = ux*ly*lz/8.0; N[2] = ux*uy*lz/8.0; N[3] = lx*uy*lz/8.0;
    N[4] = lx*ly*uz/8.0; N[5] = ux*ly*uz/8.0; N[6] = ux*uy*uz/8.0; N[7] = lx*uy*uz/8.0;
    const double d[8][3] = {{-ly*lz/8.0, -lx*lz/8.0, -lx*ly/8.0}, { ly*lz/8.0, -ux*lz/8.0, -ux*ly/8.0},
                            { uy*lz/8.0,  ux*lz/8.0, -ux*uy/8.0}, {-uy*lz/8.0,  lx*lz/8.0, -lx*uy/8.0},
                            {-ly*uz/8.0, -lx*uz/8.0,  lx*ly/8.0}, { ly*uz/8.0, -ux*uz/8.0,  ux*ly/8.0},
                            { uy*uz/8.0,  ux*uz/8.0,  ux*uy/8.0}, {-uy*uz/8.0,  lx*uz/8.0,  lx*uy/8.0}};
    for (int a = 0; a < 8; a++) for (int k = 0; k < 3; k++) Nxi[a*3 + k] = d[a][k];
  } else {
    const double x0 = xi[0], x1 = xi[1], x2 = xi[2], s = 1.0 - x0 - x1 - x2;
    N[0] = x0*(2.0*x0 - 1.0); N[1] = x1*(2.0*x1 - 1.0); N[2] = x2*(2.0*x2 - 1.0); N[3] = s*(2.0*s - 1.0);
    N[4] = 4.0*x0*x1; N[5] = 4.0*x1*x2; N[6] = 4.0*x0*x2; N[7] = 4.0*x0*s; N[8] = 4.0*x1*s; N[9] = 4.0*x2*s;
    const double d[10][3] = {{4.0*x0 - 1.0, 0.0, 0.0}, {0.0, 4.0*x1 - 1.0, 0.0}, {0.0, 0.0, 4.0*x2 - 1.0},
                             {1.0 - 4.0*s, 1.0 - 4.0*s, 1.0 - 4.0*s}, {4.0*x1, 4.0*x0, 0.0}, {0.0, 4.0*x2, 4.0*x1},
                             {4.0*x2, 0.0, 4.0*x0}, {4.0*(s - x0), -4.0*x0, -4.0*x0}, {-4.0*x1, 4.0*(s - x1), -4.0*x1},
                             {-4.0*x2, -4.0*x2, 4.0*(s - x2)}};
    for (int a = 0; a < 10; a++) for (int k = 0; k < 3; k++) Nxi[a*3 + k] = d[a][k];
  }
}

// mat_fun::mat_det / mat_inv for nd = 3 (mat_fun.cpp:65-95, 158-174)
SVB_HD double det3(const double A[3][3])
{
  double D = 0.0;
  D = D + 1.0*A[0][0]*(A[1][1]*A[2][2] - A[1][2]*A[2][1]);
  D = D + (-1.0)*A[0][1]*(A[1][0]*A[2][2] - A[1][2]*A[2][0]);
  D = D + 1.0*A[0][2]*(A[1][0]*A[2][1] - A[1][1]*A[2][0]);
  return D;
}
SVB_HD void inv3(const double A[3][3], double Ai[3][3])
{
  const double d = det3(A);
  Ai[0][0] = (A[1][1]*A[2][2] - A[1][2]*A[2][1])/d;
  Ai[0][1] = (A[0][2]*A[2][1] - A[0][1]*A[2][2])/d;
  Ai[0][2] = (A[0][1]*A[1][2] - A[0][2]*A[1][1])/d;
  Ai[1][0] = (A[1][2]*A[2][0] - A[1][0]*A[2][2])/d;
  Ai[1][1] = (A[0][0]*A[2][2] - A[0][2]*A[2][0])/d;
  Ai[1][2] = (A[0][2]*A[1][0] - A[0][0]*A[1][2])/d;
  Ai[2][0] = (A[1][0]*A[2][1] - A[1][1]*A[2][0])/d;
  Ai[2][1] = (A[0][1]*A[2][0] - A[0][0]*A[2][1])/d;
  Ai[2][2] = (A[0][0]*A[1][1] - A[0][1]*A[1][0])/d;
}

struct FolwConsts {
  double afl;               // tangent factor: struct eq.af*eq.beta*dt*dt (sv_struct.cpp:131), ustruct eq.af*eq.gam*dt (ustruct.cpp:140)
  double afm;               // ustruct only: afl / eq.am, the factor of the velocity-block copy of the tangent (ustruct.cpp:141)
  int tDof, s;              // state width and eq.s (rows of the displacement in Dg)
  double xi0[3];            // mean of the parent's Gauss points: start of the Newton inverse map (eq_assem.cpp:249-253)
  double xib[2][3];         // bounds of the parent's parametric coordinates (+- 1e-4, nn.cpp:186-293)
  double Nb_lo_c, Nb_hi_c;  // bounds of the corner / (TET10) mid-edge shape functions
  double Nb_lo_m, Nb_hi_m;
};

// parent-element data of FolwConsts: what select_ele / get_nn_bnds leave in lM.xi, lM.xib, lM.Nb (host only)
inline void fill_folw_parent(FolwConsts& c, const ElemTables& t)
{
  double xi0[3] = {0.0, 0.0, 0.0};
  for (int g = 0; g < t.nG; g++) for (int i = 0; i < 3; i++) xi0[i] = xi0[i] + t.xi[g][i];
  for (int i = 0; i < 3; i++) c.xi0[i] = xi0[i]/static_cast<double>(t.nG);
  const double tol = 1.0E-4;
  const bool tet = (t.eNoN == 4 || t.eNoN == 10);
  for (int i = 0; i < 3; i++) { c.xib[0][i] = (tet ? 0.0 : -1.0) - tol; c.xib[1][i] = 1.0 + tol; }
  c.Nb_lo_c = ((t.eNoN == 10) ? -0.125 : 0.0) - tol; c.Nb_hi_c = 1.0 + tol;
  c.Nb_lo_m = 0.0 - tol; c.Nb_hi_m = 4.0 + tol;
}

// One face element with NB nodes / NG Gauss points of a parent with NP nodes.  pn[NP]: parent nodes (assembly ids), nd[NB]: face
// nodes, inode: a parent node off the face.  Outputs lR[a*3 + i] (a over the PARENT nodes) and lK6[(a*NP + b)*6 + q] = the six
// off-diagonal entries (0,1), (1,0), (0,2), (2,0), (1,2), (2,1) of the 3x3 block (a,b); the diagonal stays zero.  For the
// struct equation that is lK (b_struct_3d); for ustruct it is lKd and lK6m != null receives the velocity-block entries
// afm*Ku of lK (b_ustruct_3d, ustruct.cpp:132-211).
// Returns 0, or 1 when the inverse map fails (the reference throws "Error in computing shape functions", nn.cpp:362).
template <int NP, int NB, int NG>
SVB_HD_NOINL int face_follower_element(const FolwConsts& c, const int* pn, const int* nd, int inode, const double* x, const double* Dg,
                                       const double* hg, const double* wtab, const double* Ntab, const double* Nxtab,
                                       double* lR, double* lK6, double* lK6m)
{
  const int tD = c.tDof;
  double xl[NP][3], dl[NP][3], hl[NP];
  for (int a = 0; a < NP; a++) {
    const size_t A = size_t(pn[a]);
    hl[a] = hg[A];
    for (int i = 0; i < 3; i++) { xl[a][i] = x[A*3 + i]; dl[a][i] = Dg[A*tD + c.s + i]; }
  }
  double xf[NB][3], xin[3];
  for (int a = 0; a < NB; a++) for (int i = 0; i < 3; i++) xf[a][i] = x[size_t(nd[a])*3 + i];
  for (int i = 0; i < 3; i++) xin[i] = x[size_t(inode)*3 + i];
  for (int q = 0; q < NP*3; q++) lR[q] = 0.0;
  for (int q = 0; q < NP*NP*6; q++) lK6[q] = 0.0;
  if (lK6m) for (int q = 0; q < NP*NP*6; q++) lK6m[q] = 0.0;
  const double afl = c.afl, afm = c.afm;
  int fail = 0;
  for (int g = 0; g < NG; g++) {
    // physical position of the face Gauss point
    double xp[3] = {0.0, 0.0, 0.0};
    for (int a = 0; a < NB; a++)
      for (int i = 0; i < 3; i++) xp[i] = xp[i] + xf[a][i]*Ntab[g*NB + a];
    // nn::get_xi: Newton inverse map into the parent (nn.cpp:369-440)
    double xi[3] = {c.xi0[0], c.xi0[1], c.xi0[2]};
    double N[NP], Nxi[NP*3];
    bool conv = false;
    int itr = 0;
    while (true) {
      itr = itr + 1;
      parent_shape<NP>(xi, N, Nxi);
      double xK[3] = {0.0, 0.0, 0.0}, rK[3];
      for (int i = 0; i < 3; i++) {
        for (int a = 0; a < NP; a++) xK[i] = xK[i] + N[a]*xl[a][i];
        rK[i] = xK[i] - xp[i];
      }
      double rmsA = 0.0, rmsR = 0.0;
      for (int i = 0; i < 3; i++) {
        rmsA = rmsA + rK[i]*rK[i];
        const double q = rK[i]/(xK[i] + 2.220446049250313e-16);
        rmsR = rmsR + q*q;
      }
      rmsA = sqrt(rmsA/3.0);
      rmsR = sqrt(rmsR/3.0);
      const bool l1 = itr > 5, l2 = rmsA <= 1.0e-12, l3 = rmsR <= 1.0e-6;
      if (l1 || l2 || l3) { conv = l2 || l3; break; }
      double Am[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, Ai[3][3];
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          for (int a = 0; a < NP; a++) Am[i][j] = Am[i][j] + xl[a][i]*Nxi[a*3 + j];
      inv3(Am, Ai);
      double dx[3];
      for (int i = 0; i < 3; i++) dx[i] = ((0.0 + Ai[i][0]*rK[0]) + Ai[i][1]*rK[1]) + Ai[i][2]*rK[2];
      for (int i = 0; i < 3; i++) xi[i] = xi[i] - dx[i];
    }
    // nn::get_nnx checks (nn.cpp:333-364)
    {
      int j = 0;
      for (int i = 0; i < 3; i++) if (xi[i] >= c.xib[0][i] && xi[i] <= c.xib[1][i]) j++;
      const bool l2 = (j == 3);
      parent_shape<NP>(xi, N, Nxi);
      j = 0;
      double rt = 0.0;
      for (int a = 0; a < NP; a++) {
        rt = rt + N[a];
        const double lo = (NP == 10 && a >= 4) ? c.Nb_lo_m : c.Nb_lo_c, hi = (NP == 10 && a >= 4) ? c.Nb_hi_m : c.Nb_hi_c;
        if (N[a] > lo && N[a] < hi) j++;
      }
      const bool l3 = (j == NP), l4 = (rt >= 0.9999) && (rt <= 1.0001);
      if (!(conv && l2 && l3 && l4)) fail = 1;
    }
    // nn::gnn on the parent at xi: dN/dX in the reference configuration
    double Nx[NP][3];
    {
      double xXi[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (int a = 0; a < NP; a++)
        for (int i = 0; i < 3; i++) {
          xXi[i][0] = xXi[i][0] + xl[a][i]*Nxi[a*3 + 0];
          xXi[i][1] = xXi[i][1] + xl[a][i]*Nxi[a*3 + 1];
          xXi[i][2] = xXi[i][2] + xl[a][i]*Nxi[a*3 + 2];
        }
      const double Jp = xXi[0][0]*xXi[1][1]*xXi[2][2] + xXi[0][1]*xXi[1][2]*xXi[2][0] + xXi[0][2]*xXi[1][0]*xXi[2][1]
                      - xXi[0][0]*xXi[1][2]*xXi[2][1] - xXi[0][1]*xXi[1][0]*xXi[2][2] - xXi[0][2]*xXi[1][1]*xXi[2][0];
      double xiX[3][3];
      xiX[0][0] = (xXi[1][1]*xXi[2][2] - xXi[1][2]*xXi[2][1])/Jp;
      xiX[0][1] = (xXi[2][1]*xXi[0][2] - xXi[2][2]*xXi[0][1])/Jp;
      xiX[0][2] = (xXi[0][1]*xXi[1][2] - xXi[0][2]*xXi[1][1])/Jp;
      xiX[1][0] = (xXi[1][2]*xXi[2][0] - xXi[1][0]*xXi[2][2])/Jp;
      xiX[1][1] = (xXi[2][2]*xXi[0][0] - xXi[2][0]*xXi[0][2])/Jp;
      xiX[1][2] = (xXi[0][2]*xXi[1][0] - xXi[0][0]*xXi[1][2])/Jp;
      xiX[2][0] = (xXi[1][0]*xXi[2][1] - xXi[1][1]*xXi[2][0])/Jp;
      xiX[2][1] = (xXi[2][0]*xXi[0][1] - xXi[2][1]*xXi[0][0])/Jp;
      xiX[2][2] = (xXi[0][0]*xXi[1][1] - xXi[0][1]*xXi[1][0])/Jp;
      for (int a = 0; a < NP; a++)
        for (int i = 0; i < 3; i++)
          Nx[a][i] = ((0.0 + Nxi[a*3 + 0]*xiX[0][i]) + Nxi[a*3 + 1]*xiX[1][i]) + Nxi[a*3 + 2]*xiX[2][i];
    }
    // nn::gnnb: reference-configuration normal of the face
    double nV[3];
    {
      double xXi[3][2] = {{0, 0}, {0, 0}, {0, 0}};
      for (int a = 0; a < NB; a++)
        for (int i = 0; i < 2; i++)
          for (int j = 0; j < 3; j++) xXi[j][i] = xXi[j][i] + Nxtab[(g*NB + a)*2 + i]*xf[a][j];
      nV[0] = xXi[1][0]*xXi[2][1] - xXi[2][0]*xXi[1][1];
      nV[1] = xXi[2][0]*xXi[0][1] - xXi[0][0]*xXi[2][1];
      nV[2] = xXi[0][0]*xXi[1][1] - xXi[1][0]*xXi[0][1];
      double dotv = 0.0;
      for (int i = 0; i < 3; i++) dotv += nV[i]*(xf[0][i] - xin[i]);
      if (dotv < 0.0) { nV[0] = -nV[0]; nV[1] = -nV[1]; nV[2] = -nV[2]; }
    }
    double nn = 0.0;
    for (int i = 0; i < 3; i++) nn += nV[i]*nV[i];
    const double Jac = sqrt(nn);
    for (int i = 0; i < 3; i++) nV[i] = nV[i]/Jac;
    const double w = wtab[g]*Jac;
    // struct_ns::b_struct_3d (sv_struct.cpp:116-210)
    double F[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double h = 0.0;
    for (int a = 0; a < NP; a++) {
      h = h + N[a]*hl[a];
      for (int i = 0; i < 3; i++) {
        F[i][0] = F[i][0] + Nx[a][0]*dl[a][i];
        F[i][1] = F[i][1] + Nx[a][1]*dl[a][i];
        F[i][2] = F[i][2] + Nx[a][2]*dl[a][i];
      }
    }
    const double JF = det3(F);
    double Fi[3][3];
    inv3(F, Fi);
    double NxFi[NP][3], nFi[3];
    for (int a = 0; a < NP; a++)
      for (int i = 0; i < 3; i++) NxFi[a][i] = Nx[a][0]*Fi[0][i] + Nx[a][1]*Fi[1][i] + Nx[a][2]*Fi[2][i];
    for (int i = 0; i < 3; i++) nFi[i] = nV[0]*Fi[0][i] + nV[1]*Fi[1][i] + nV[2]*Fi[2][i];
    const double wl = w*JF*h;
    for (int a = 0; a < NP; a++) {
      for (int i = 0; i < 3; i++) lR[a*3 + i] = lR[a*3 + i] - wl*N[a]*nFi[i];
      for (int b = 0; b < NP; b++) {
        double* k6 = lK6 + (a*NP + b)*6;
        double* m6 = lK6m ? lK6m + (a*NP + b)*6 : nullptr;
        double Ku = wl*afl*N[a]*(nFi[1]*NxFi[b][0] - nFi[0]*NxFi[b][1]);
        k6[0] = k6[0] + Ku; k6[1] = k6[1] - Ku;
        if (m6) { m6[0] = m6[0] + afm*Ku; m6[1] = m6[1] - afm*Ku; }
        Ku = wl*afl*N[a]*(nFi[2]*NxFi[b][0] - nFi[0]*NxFi[b][2]);
        k6[2] = k6[2] + Ku; k6[3] = k6[3] - Ku;
        if (m6) { m6[2] = m6[2] + afm*Ku; m6[3] = m6[3] - afm*Ku; }
        Ku = wl*afl*N[a]*(nFi[2]*NxFi[b][1] - nFi[1]*NxFi[b][2]);
        k6[4] = k6[4] + Ku; k6[5] = k6[5] - Ku;
        if (m6) { m6[4] = m6[4] + afm*Ku; m6[5] = m6[5] - afm*Ku; }
      }
    }
  }
  return fail;
}

} // namespace svb200
