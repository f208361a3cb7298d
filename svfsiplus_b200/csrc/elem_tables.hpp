// elem_tables.hpp — Gauss rules and shape-function tables of the element types the device kernels cover:
// what nn::select_ele + get_gip + get_gnn (+ get_gn_nxx via fs::init_fs_msh) leave in lM.w / lM.N / lM.Nx /
// lM.fs[0].Nxx of the reference.
//   TET4   nn_elem_gip.h:501-517, nn_elem_gnn.h:1232-1250   4 points, w = 1/24, s = qmTET4
//   HEX8   nn_elem_gip.h:298-330 (+-1/sqrt(3), w = 1), nn_elem_gnn.h:732-786; no second-derivative table
//          (nn::get_gn_nxx returns early, nn.cpp:166-171) => Nxi2 = 0
//   TET10  nn_elem_gip.h:520-565 (15 points), nn_elem_gnn.h:1256-1310, nn_elem_gnnxx.h:137-150
// Host code only (no device types): shared by csrc/api.cu and the CPU element test.
#pragma once

#include <cmath>
#include <cstring>
#include <vector>

namespace svb200 {

struct ElemTables {            // lM.w, lM.N, lM.Nx, lM.fs[0].Nxx
  enum { MAXN = 10, MAXG = 15 };
  int eNoN, nG;
  bool has_nxx;                // false: the reference's second parametric derivatives are identically zero
  double w[MAXG];
  double N[MAXG][MAXN];        // [g][a]
  double Nxi[MAXG][MAXN][3];   // [g][a][i]
  double Nxi2[MAXG][MAXN][6];  // [g][a][k], k = (00, 11, 22, 01, 12, 02)
  double xi[MAXG][3];          // lM.xi: parametric coordinates of the Gauss points
};

inline bool elem_supported(int eNoN) { return eNoN == 4 || eNoN == 8 || eNoN == 10; }

inline void fill_tables(ElemTables& t, int eNoN, double qmTET4)
{
  std::memset(&t, 0, sizeof(t));
  t.eNoN = eNoN;
  t.has_nxx = false;
  if (eNoN == 4) {
    t.nG = 4;
    const double s = qmTET4, r = (1.0 - s)/3.0;
    const double xi[4][3] = {{s, r, r}, {r, s, r}, {r, r, s}, {r, r, r}};
    for (int g = 0; g < 4; g++) {
      t.w[g] = 1.0/24.0;
      for (int i = 0; i < 3; i++) t.xi[g][i] = xi[g][i];
      t.N[g][0] = xi[g][0]; t.N[g][1] = xi[g][1]; t.N[g][2] = xi[g][2];
      t.N[g][3] = 1.0 - xi[g][0] - xi[g][1] - xi[g][2];
      const double d[4][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, -1, -1}};
      for (int a = 0; a < 4; a++) for (int i = 0; i < 3; i++) t.Nxi[g][a][i] = d[a][i];
    }
  } else if (eNoN == 8) {
    t.nG = 8;
    const double s = 1.0/std::sqrt(3.0), m = -1.0/std::sqrt(3.0);
    const double xi[8][3] = {{m, m, m}, {s, m, m}, {s, s, m}, {m, s, m}, {m, m, s}, {s, m, s}, {s, s, s}, {m, s, s}};
    for (int g = 0; g < 8; g++) {
      t.w[g] = 1.0;
      for (int i = 0; i < 3; i++) t.xi[g][i] = xi[g][i];
      const double lx = 1.0 - xi[g][0], ly = 1.0 - xi[g][1], lz = 1.0 - xi[g][2];
      const double ux = 1.0 + xi[g][0], uy = 1.0 + xi[g][1], uz = 1.0 + xi[g][2];
      const double N[8] = {lx*ly*lz/8.0, ux*ly*lz/8.0, ux*uy*lz/8.0, lx*uy*lz/8.0, lx*ly*uz/8.0, ux*ly*uz/8.0, ux*uy*uz/8.0, lx*uy*uz/8.0};
      const double D[8][3] = {{-ly*lz/8.0, -lx*lz/8.0, -lx*ly/8.0}, { ly*lz/8.0, -ux*lz/8.0, -ux*ly/8.0},
                              { uy*lz/8.0,  ux*lz/8.0, -ux*uy/8.0}, {-uy*lz/8.0,  lx*lz/8.0, -lx*uy/8.0},
                              {-ly*uz/8.0, -lx*uz/8.0,  lx*ly/8.0}, { ly*uz/8.0, -ux*uz/8.0,  ux*ly/8.0},
                              { uy*uz/8.0,  ux*uz/8.0,  ux*uy/8.0}, {-uy*uz/8.0,  lx*uz/8.0,  lx*uy/8.0}};
      for (int a = 0; a < 8; a++) { t.N[g][a] = N[a]; for (int i = 0; i < 3; i++) t.Nxi[g][a][i] = D[a][i]; }
    }
  } else {
    t.nG = 15;
    t.has_nxx = true;
    const double w0 = 0.0302836780970890, w1 = 0.0060267857142860, w2 = 0.0116452490860290, w3 = 0.0109491415613860;
    const double wt[15] = {w0, w1, w1, w1, w1, w2, w2, w2, w2, w3, w3, w3, w3, w3, w3};
    double xi[15][3];
    {
      double s = 0.250;
      xi[0][0] = s; xi[0][1] = s; xi[0][2] = s;
      s = 0.3333333333333330;
      double q = 0.0;
      xi[1][0] = q; xi[1][1] = s; xi[1][2] = s;
      xi[2][0] = s; xi[2][1] = q; xi[2][2] = s;
      xi[3][0] = s; xi[3][1] = s; xi[3][2] = q;
      xi[4][0] = s; xi[4][1] = s; xi[4][2] = s;
      s = 0.0909090909090910; q = 0.7272727272727270;
      xi[5][0] = q; xi[5][1] = s; xi[5][2] = s;
      xi[6][0] = s; xi[6][1] = q; xi[6][2] = s;
      xi[7][0] = s; xi[7][1] = s; xi[7][2] = q;
      xi[8][0] = s; xi[8][1] = s; xi[8][2] = s;
      s = 0.0665501535736640; q = 0.4334498464263360;
      xi[9][0]  = s; xi[9][1]  = s; xi[9][2]  = q;
      xi[10][0] = s; xi[10][1] = q; xi[10][2] = s;
      xi[11][0] = s; xi[11][1] = q; xi[11][2] = q;
      xi[12][0] = q; xi[12][1] = q; xi[12][2] = s;
      xi[13][0] = q; xi[13][1] = s; xi[13][2] = q;
      xi[14][0] = q; xi[14][1] = s; xi[14][2] = s;
    }
    const double fp = 4.0, fn = -4.0, en = -8.0, ze = 0.0;
    const double X2[10][6] = {{fp, ze, ze, ze, ze, ze}, {ze, fp, ze, ze, ze, ze}, {ze, ze, fp, ze, ze, ze}, {fp, fp, fp, fp, fp, fp},
                              {ze, ze, ze, fp, ze, ze}, {ze, ze, ze, ze, fp, ze}, {ze, ze, ze, ze, ze, fp}, {en, ze, ze, fn, ze, fn},
                              {ze, en, ze, fn, fn, ze}, {ze, ze, en, ze, fn, fn}};
    for (int g = 0; g < 15; g++) {
      t.w[g] = wt[g];
      for (int i = 0; i < 3; i++) t.xi[g][i] = xi[g][i];
      const double x0 = xi[g][0], x1 = xi[g][1], x2 = xi[g][2];
      const double s = 1.0 - x0 - x1 - x2;
      double* N = t.N[g];
      N[0] = x0*(2.0*x0 - 1.0); N[1] = x1*(2.0*x1 - 1.0); N[2] = x2*(2.0*x2 - 1.0); N[3] = s*(2.0*s - 1.0);
      N[4] = 4.0*x0*x1; N[5] = 4.0*x1*x2; N[6] = 4.0*x0*x2; N[7] = 4.0*x0*s; N[8] = 4.0*x1*s; N[9] = 4.0*x2*s;
      const double D[10][3] = {{4.0*x0 - 1.0, 0.0, 0.0}, {0.0, 4.0*x1 - 1.0, 0.0}, {0.0, 0.0, 4.0*x2 - 1.0},
                               {1.0 - 4.0*s, 1.0 - 4.0*s, 1.0 - 4.0*s}, {4.0*x1, 4.0*x0, 0.0}, {0.0, 4.0*x2, 4.0*x1},
                               {4.0*x2, 0.0, 4.0*x0}, {4.0*(s - x0), -4.0*x0, -4.0*x0}, {-4.0*x1, 4.0*(s - x1), -4.0*x1},
                               {-4.0*x2, -4.0*x2, 4.0*(s - x2)}};
      for (int a = 0; a < 10; a++) {
        for (int i = 0; i < 3; i++) t.Nxi[g][a][i] = D[a][i];
        for (int k = 0; k < 6; k++) t.Nxi2[g][a][k] = X2[a][k];
      }
    }
  }
}

// packed copy the kernels stage in shared memory: w[nG], N[nG][eNoN], Nxi[nG][eNoN][3], Nxi2[nG][eNoN][6]
inline std::vector<double> pack_tables(const ElemTables& t)
{
  std::vector<double> pk;
  for (int g = 0; g < t.nG; g++) pk.push_back(t.w[g]);
  for (int g = 0; g < t.nG; g++) for (int a = 0; a < t.eNoN; a++) pk.push_back(t.N[g][a]);
  for (int g = 0; g < t.nG; g++) for (int a = 0; a < t.eNoN; a++) for (int i = 0; i < 3; i++) pk.push_back(t.Nxi[g][a][i]);
  for (int g = 0; g < t.nG; g++) for (int a = 0; a < t.eNoN; a++) for (int k = 0; k < 6; k++) pk.push_back(t.Nxi2[g][a][k]);
  return pk;
}

} // namespace svb200
