// tests/hostlogic/fluid_elem_host.cpp — TEST-ONLY host instantiation of svfsiplus_b200/csrc/fluid_elem.hpp.
// It lets the CPU test-suite (`-m "not gpu"`) check the Gauss-point arithmetic of the generic fluid element (HEX8,
// TET10: gnn, gn_nxx, fluid_3d_m, fluid_3d_c) against the compiled reference without a GPU.  It walks the elements
// serially and adds lR / lK straight into R / Val in element order (what do_assem does, lhsa.cpp:97-142).  It is
// never built into, loaded by, or reachable from the product library (libsvb200.so runs the same header on the
// device through assembly_fluid_gen.cuh).
#include <algorithm>
#include <cstring>
#include <vector>

#include "elem_tables.hpp"
#include "fluid_elem.hpp"
#include "face_elem.hpp"
#include "solid_law.hpp"

using namespace svb200;

namespace {

template <int N, int NG, bool NXX>
int assemble(int nEl, const int* ien, const double* x, const double* Dmesh, const FluidConsts& c, const ElemTables& t,
             const double* Ag, const double* Yg, const double* Bf, const int* rowPtr, const int* colPtr, double* R, double* Val)
{
  typedef FluidRec<N, NXX> L;
  std::vector<double> Ntab(NG*N), Nxi(NG*N*3), Nxi2(NG*N*6);
  for (int g = 0; g < NG; g++)
    for (int a = 0; a < N; a++) {
      Ntab[g*N + a] = t.N[g][a];
      for (int i = 0; i < 3; i++) Nxi[(g*N + a)*3 + i] = t.Nxi[g][a][i];
      for (int k = 0; k < 6; k++) Nxi2[(g*N + a)*6 + k] = t.Nxi2[g][a][k];
    }
  std::vector<double> recs(static_cast<size_t>(NG)*L::SIZE);
  for (int e = 0; e < nEl; e++) {
    const int* nd = ien + size_t(e)*N;
    for (int g = 0; g < NG; g++)
      if (!fluid_geom<N, NXX>(nd, x, Dmesh, c.tDof, t.w[g], &Nxi[g*N*3], NXX ? &Nxi2[g*N*6] : nullptr, &recs[size_t(g)*L::SIZE]))
        return e + 1;
    for (int g = 0; g < NG; g++)
      fluid_point<N, NXX>(c, nd, Ag, Yg, Bf, &Ntab[g*N], &recs[size_t(g)*L::SIZE], &recs[size_t(NG - 1)*L::SIZE]);
    for (int a = 0; a < N; a++) {
      double r[4];
      fluid_res_row<N, NG, NXX>(c, recs.data(), Ntab.data(), a, r);
      for (int i = 0; i < 4; i++) R[size_t(nd[a])*4 + i] += r[i];
      const int* beg = colPtr + rowPtr[nd[a]];
      const int* end = colPtr + rowPtr[nd[a] + 1];
      for (int b = 0; b < N; b++) {
        double kb[16];
        fluid_tan_block<N, NG, NXX>(c, recs.data(), Ntab.data(), a, b, kb);
        const int* it = std::lower_bound(beg, end, nd[b]);
        if (it == end || *it != nd[b]) return -(e + 1);
        double* v = Val + size_t(it - colPtr)*16;
        for (int i = 0; i < 16; i++) v[i] += kb[i];
      }
    }
  }
  return 0;
}

} // namespace

extern "C" {

// par = {dt, am, af, gam, rho, f0, f1, f2, Kinv, viscType, mu_i, mu_o, lam, a, n, tDof, mvMsh}.
// Dmesh: Dg(tDof,nNo) for the ALE configuration or NULL.  R(4,nNo), Val(16,nnz) must be zero on entry.
// Returns 0, e+1 when element e has a zero Jacobian, -(e+1) when a column is missing from the pattern, -1000000 for
// an unsupported element.
int host_fluid_assemble(int eNoN, int nEl, const int* ien, const double* x, const double* Dmesh, const double* par, double qmTET4,
                        const double* Ag, const double* Yg, const double* Bf, const int* rowPtr, const int* colPtr,
                        double* R, double* Val)
{
  FluidConsts c;
  std::memset(&c, 0, sizeof(c));
  c.dt = par[0]; c.am = par[1]; c.af = par[2]; c.gam = par[3]; c.rho = par[4];
  c.f[0] = par[5]; c.f[1] = par[6]; c.f[2] = par[7]; c.Kinv = par[8];
  c.viscType = int(par[9]); c.mu_i = par[10]; c.mu_o = par[11]; c.lam = par[12]; c.a = par[13]; c.n = par[14];
  c.tDof = int(par[15]); c.mvMsh = int(par[16]);
  ElemTables t;
  if (!elem_supported(eNoN)) return -1000000;
  fill_tables(t, eNoN, qmTET4 > 0.0 ? qmTET4 : (5.0 + 3.0*std::sqrt(5.0))/20.0);
  if (eNoN == 4) return assemble<4, 4, false>(nEl, ien, x, Dmesh, c, t, Ag, Yg, Bf, rowPtr, colPtr, R, Val);
  if (eNoN == 8) return assemble<8, 8, false>(nEl, ien, x, Dmesh, c, t, Ag, Yg, Bf, rowPtr, colPtr, R, Val);
  return assemble<10, 15, true>(nEl, ien, x, Dmesh, c, t, Ag, Yg, Bf, rowPtr, colPtr, R, Val);
}

// Boundary-face assembly (face_elem.hpp): b_assem_neu_bc on one face, serial, face elements in order.
// par = {dt, af, gam, rho, bfs, tDof, mvMsh}; kind 0 b_fluid (dof 4), 1 b_l_elas (dof 3).  ien: the volume mesh (eNoN x nEl).
// R(dof,nNo), Val(dof*dof,nnz) must be zero on entry.
int host_bneu_assemble(int kind, int eNoN, const int* ien, int eNoNb, int nElb, const int* IENb, const int* gE, const double* par,
                       const double* x, const double* Do, const double* hg, const double* Yg, const int* rowPtr, const int* colPtr,
                       double* R, double* Val)
{
  if (!face_supported(eNoNb)) return -1000000;
  BneuConsts c;
  c.dt = par[0]; c.af = par[1]; c.gam = par[2]; c.rho = par[3]; c.bfs = par[4]; c.tDof = int(par[5]); c.mvMsh = int(par[6]);
  c.kind = kind; c.dof = (kind == 0) ? 4 : 3;
  FaceTables t;
  fill_face_tables(t, eNoNb, 2.0/3.0);
  std::vector<double> N(t.nG*eNoNb), Nx(t.nG*eNoNb*2);
  for (int g = 0; g < t.nG; g++)
    for (int a = 0; a < eNoNb; a++) {
      N[g*eNoNb + a] = t.N[g][a];
      Nx[(g*eNoNb + a)*2] = t.Nx[g][a][0]; Nx[(g*eNoNb + a)*2 + 1] = t.Nx[g][a][1];
    }
  const int dof = c.dof;
  for (int e = 0; e < nElb; e++) {
    const int* nd = IENb + size_t(e)*eNoNb;
    const int* pn = ien + size_t(gE[e])*eNoN;
    int inode = -1;
    for (int b = 0; b < eNoN && inode < 0; b++)
      if (std::find(nd, nd + eNoNb, pn[b]) == nd + eNoNb) inode = pn[b];
    if (inode < 0) return e + 1;
    double lR[18], lKd[36];
    if (eNoNb == 3) face_element<3, 3>(c, nd, inode, x, Do, hg, Yg, t.w, N.data(), Nx.data(), lR, lKd);
    else if (eNoNb == 4) face_element<4, 4>(c, nd, inode, x, Do, hg, Yg, t.w, N.data(), Nx.data(), lR, lKd);
    else face_element<6, 7>(c, nd, inode, x, Do, hg, Yg, t.w, N.data(), Nx.data(), lR, lKd);
    for (int a = 0; a < eNoNb; a++) {
      for (int i = 0; i < 3; i++) R[size_t(nd[a])*dof + i] += lR[a*3 + i];
      if (kind != 0) continue;
      const int* beg = colPtr + rowPtr[nd[a]];
      const int* end = colPtr + rowPtr[nd[a] + 1];
      for (int b = 0; b < eNoNb; b++) {
        const int* it = std::lower_bound(beg, end, nd[b]);
        if (it == end || *it != nd[b]) return -(e + 1);
        double* v = Val + size_t(it - colPtr)*16;
        v[0] += lKd[a*eNoNb + b]; v[5] += lKd[a*eNoNb + b]; v[10] += lKd[a*eNoNb + b];
      }
    }
  }
  return 0;
}

// all_fun::integ over one face (face_integ_terms), the running sum in (element, Gauss point) order.
double host_face_integ(int eNoN, const int* ien, int eNoNb, int nElb, const int* IENb, const int* gE, const double* x,
                       const double* geo, int gtD, int goff, const double* s, int stD, int l, int nrow)
{
  FaceTables t;
  fill_face_tables(t, eNoNb, 2.0/3.0);
  std::vector<double> N(t.nG*eNoNb), Nx(t.nG*eNoNb*2);
  for (int g = 0; g < t.nG; g++)
    for (int a = 0; a < eNoNb; a++) {
      N[g*eNoNb + a] = t.N[g][a];
      Nx[(g*eNoNb + a)*2] = t.Nx[g][a][0]; Nx[(g*eNoNb + a)*2 + 1] = t.Nx[g][a][1];
    }
  double result = 0.0;
  for (int e = 0; e < nElb; e++) {
    const int* nd = IENb + size_t(e)*eNoNb;
    const int* pn = ien + size_t(gE[e])*eNoN;
    int inode = -1;
    for (int b = 0; b < eNoN && inode < 0; b++)
      if (std::find(nd, nd + eNoNb, pn[b]) == nd + eNoNb) inode = pn[b];
    double terms[7];
    if (eNoNb == 3) face_integ_terms<3, 3>(nd, inode, x, geo, gtD, goff, s, stD, l, nrow, t.w, N.data(), Nx.data(), terms);
    else if (eNoNb == 4) face_integ_terms<4, 4>(nd, inode, x, geo, gtD, goff, s, stD, l, nrow, t.w, N.data(), Nx.data(), terms);
    else face_integ_terms<6, 7>(nd, inode, x, geo, gtD, goff, s, stD, l, nrow, t.w, N.data(), Nx.data(), terms);
    for (int g = 0; g < t.nG; g++) result = result + terms[g];
  }
  return result;
}

// fsi_ls_upd (face_normal_terms): sV(3,nNo) accumulated per node in (element, Gauss point) order; sV must be zero on entry.
int host_face_normals(int eNoN, const int* ien, int eNoNb, int nElb, const int* IENb, const int* gE, const double* x,
                      const double* geo, int gtD, int goff, double* sV)
{
  FaceTables t;
  fill_face_tables(t, eNoNb, 2.0/3.0);
  std::vector<double> N(t.nG*eNoNb), Nx(t.nG*eNoNb*2);
  for (int g = 0; g < t.nG; g++)
    for (int a = 0; a < eNoNb; a++) {
      N[g*eNoNb + a] = t.N[g][a];
      Nx[(g*eNoNb + a)*2] = t.Nx[g][a][0]; Nx[(g*eNoNb + a)*2 + 1] = t.Nx[g][a][1];
    }
  for (int e = 0; e < nElb; e++) {
    const int* nd = IENb + size_t(e)*eNoNb;
    const int* pn = ien + size_t(gE[e])*eNoN;
    int inode = -1;
    for (int b = 0; b < eNoN && inode < 0; b++)
      if (std::find(nd, nd + eNoNb, pn[b]) == nd + eNoNb) inode = pn[b];
    if (inode < 0) return e + 1;
    double out[6*7*3];
    if (eNoNb == 3) face_normal_terms<3, 3>(nd, inode, x, geo, gtD, goff, t.w, N.data(), Nx.data(), out);
    else if (eNoNb == 4) face_normal_terms<4, 4>(nd, inode, x, geo, gtD, goff, t.w, N.data(), Nx.data(), out);
    else face_normal_terms<6, 7>(nd, inode, x, geo, gtD, goff, t.w, N.data(), Nx.data(), out);
    for (int g = 0; g < t.nG; g++)
      for (int a = 0; a < eNoNb; a++)
        for (int i = 0; i < 3; i++) sV[size_t(nd[a])*3 + i] = sV[size_t(nd[a])*3 + i] + out[(a*t.nG + g)*3 + i];
  }
  return 0;
}

// get_pk2cc<3> (solid_law.hpp): par = {iso, vol, C10, C01, Kpen, a, b, aff, bff, ass, bss, afs, bfs, khs, Tfa, Tsa, kap};
// F row-major 3x3, fl = fibre + sheet directions; outputs S (00 11 22 01 12 20) and the 21 upper-triangle entries of Dm.
void host_pk2cc(const double* par, const double* F9, const double* fl6, double* S6, double* Dm21)
{
  SolidConsts c;
  std::memset(&c, 0, sizeof(c));
  c.iso = int(par[0]); c.vol = int(par[1]); c.C10 = par[2]; c.C01 = par[3]; c.Kpen = par[4];
  c.ho_a = par[5]; c.ho_b = par[6]; c.ho_aff = par[7]; c.ho_bff = par[8]; c.ho_ass = par[9]; c.ho_bss = par[10];
  c.ho_afs = par[11]; c.ho_bfs = par[12]; c.ho_khs = par[13]; c.Tfa = par[14]; c.Tsa = par[15]; c.kap = par[16];
  double F[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) F[i][j] = F9[i*3 + j];
  pk2cc_iso(c, F, fl6, S6, Dm21);
}

// Solid viscosity (solid_law.hpp visc_point + visc_pair): model 1 Newtonian, 2 potential; Nx (eNoN x 3, node-major); outputs
// Svis (3x3 row-major) and Ku / Kv as [a][b][9].
void host_visc(int model, double mu, int eNoN, const double* Nx, const double* vx9, const double* F9, double* Svis9, double* Ku, double* Kv)
{
  double F[3][3], vx[3][3], S[3][3], v[VISC_REC];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { F[i][j] = F9[i*3 + j]; vx[i][j] = vx9[i*3 + j]; }
  visc_point(model, mu, F, vx, S, v);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Svis9[i*3 + j] = S[i][j];
  for (int a = 0; a < eNoN; a++)
    for (int b = 0; b < eNoN; b++)
      visc_pair(model, mu, v, F9, Nx + a*3, Nx + b*3, Ku + (size_t(a)*eNoN + b)*9, Kv + (size_t(a)*eNoN + b)*9);
}

// Follower pressure load (face_follower_element): b_neu_folw_p on one face, serial.  par = {afl, afm, tDof, s, ustruct}.
// struct (ustruct = 0): R(3,nNo), Val(9,nnz); ustruct: R(4,nNo), Val(16,nnz), Kd(12,nnz).  All zero on entry.
extern "C++" {
namespace {
template <int NP, int NB, int NG>
int bfolw(const FolwConsts& c, bool ustruct, const FaceTables& ft, const double* N, const double* Nx, const int* ien, int nElb, const int* IENb,
          const int* gE, const double* x, const double* Dg, const double* hg, const int* rowPtr, const int* colPtr, double* R, double* Val,
          double* Kd)
{
  std::vector<double> lR(NP*3), lK6(NP*NP*6), lK6m(NP*NP*6);
  const int dof = ustruct ? 4 : 3;
  for (int e = 0; e < nElb; e++) {
    const int* nd = IENb + size_t(e)*NB;
    const int* pn = ien + size_t(gE[e])*NP;
    int inode = -1;
    for (int b = 0; b < NP && inode < 0; b++)
      if (std::find(nd, nd + NB, pn[b]) == nd + NB) inode = pn[b];
    if (inode < 0) return e + 1;
    if (face_follower_element<NP, NB, NG>(c, pn, nd, inode, x, Dg, hg, ft.w, N, Nx, lR.data(), lK6.data(), ustruct ? lK6m.data() : nullptr) != 0)
      return -(e + 1);
    for (int a = 0; a < NP; a++) {
      for (int i = 0; i < 3; i++) R[size_t(pn[a])*dof + i] += lR[a*3 + i];
      const int* beg = colPtr + rowPtr[pn[a]];
      const int* end = colPtr + rowPtr[pn[a] + 1];
      for (int b = 0; b < NP; b++) {
        const int* it = std::lower_bound(beg, end, pn[b]);
        if (it == end || *it != pn[b]) return -1000000;
        const double* k6 = &lK6[(a*NP + b)*6];
        if (!ustruct) {
          double* v = Val + size_t(it - colPtr)*9;
          v[1] += k6[0]; v[3] += k6[1]; v[2] += k6[2]; v[6] += k6[3]; v[5] += k6[4]; v[7] += k6[5];
        } else {
          double* kd = Kd + size_t(it - colPtr)*12;
          kd[1] += k6[0]; kd[3] += k6[1]; kd[2] += k6[2]; kd[6] += k6[3]; kd[5] += k6[4]; kd[7] += k6[5];
          const double* m6 = &lK6m[(a*NP + b)*6];
          double* v = Val + size_t(it - colPtr)*16;
          v[1] += m6[0]; v[4] += m6[1]; v[2] += m6[2]; v[8] += m6[3]; v[6] += m6[4]; v[9] += m6[5];
        }
      }
    }
  }
  return 0;
}
} // namespace
} // extern "C++"

int host_bfolw_assemble(int eNoN, const int* ien, int eNoNb, int nElb, const int* IENb, const int* gE, const double* par,
                        const double* x, const double* Dg, const double* hg, const int* rowPtr, const int* colPtr, double* R, double* Val,
                        double* Kd)
{
  if (!face_supported(eNoNb) || !elem_supported(eNoN)) return -1000001;
  FolwConsts c;
  c.afl = par[0]; c.afm = par[1]; c.tDof = int(par[2]); c.s = int(par[3]);
  const bool us = par[4] != 0.0;
  ElemTables et;
  fill_tables(et, eNoN, (5.0 + 3.0*std::sqrt(5.0))/20.0);
  fill_folw_parent(c, et);
  FaceTables t;
  fill_face_tables(t, eNoNb, 2.0/3.0);
  std::vector<double> N(t.nG*eNoNb), Nx(t.nG*eNoNb*2);
  for (int g = 0; g < t.nG; g++)
    for (int a = 0; a < eNoNb; a++) {
      N[g*eNoNb + a] = t.N[g][a];
      Nx[(g*eNoNb + a)*2] = t.Nx[g][a][0]; Nx[(g*eNoNb + a)*2 + 1] = t.Nx[g][a][1];
    }
  if (eNoN == 4 && eNoNb == 3) return bfolw<4, 3, 3>(c, us, t, N.data(), Nx.data(), ien, nElb, IENb, gE, x, Dg, hg, rowPtr, colPtr, R, Val, Kd);
  if (eNoN == 8 && eNoNb == 4) return bfolw<8, 4, 4>(c, us, t, N.data(), Nx.data(), ien, nElb, IENb, gE, x, Dg, hg, rowPtr, colPtr, R, Val, Kd);
  if (eNoN == 10 && eNoNb == 6) return bfolw<10, 6, 7>(c, us, t, N.data(), Nx.data(), ien, nElb, IENb, gE, x, Dg, hg, rowPtr, colPtr, R, Val, Kd);
  return -1000002;
}

// the tables themselves (checked against what the reference's select_ele leaves in lM): returns nG
int host_elem_tables(int eNoN, double qmTET4, double* w, double* N, double* Nxi, double* Nxi2)
{
  if (!elem_supported(eNoN)) return -1;
  ElemTables t;
  fill_tables(t, eNoN, qmTET4 > 0.0 ? qmTET4 : (5.0 + 3.0*std::sqrt(5.0))/20.0);
  for (int g = 0; g < t.nG; g++) {
    if (w) w[g] = t.w[g];
    for (int a = 0; a < eNoN; a++) {
      if (N) N[g*eNoN + a] = t.N[g][a];
      if (Nxi) for (int i = 0; i < 3; i++) Nxi[(g*eNoN + a)*3 + i] = t.Nxi[g][a][i];
      if (Nxi2) for (int k = 0; k < 6; k++) Nxi2[(g*eNoN + a)*6 + k] = t.Nxi2[g][a][k];
    }
  }
  return t.nG;
}

}
