// tests/hostlogic/host_ops.cpp — TEST-ONLY serial host policy for svfsiplus_b200/csrc/krylov.hpp.
// It lets the CPU test-suite (`-m "not gpu"`) run the product's solver CONTROL FLOW (restarts,
// Givens, Gram system, convergence tests, iteration counters, face flags) against the compiled
// reference without a GPU.  It is never built into, loaded by, or reachable from the product library
// (svfsiplus_b200/libsvb200.so instantiates krylov.hpp with CudaOps only).
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "krylov.hpp"
#include "lhs_layout.hpp"

namespace {

struct HFace {
  int nNo = 0, dof = 0, bGrp = 0;
  bool inc = true, coupled = false;
  double res = 0, nS = 0;
  std::vector<int> glob;
  std::vector<double> val, valM;
};

struct HostOps {
  int nNo_ = 0, nnz_ = 0;
  std::vector<int> rowPtr, col, diag, tpos;
  std::vector<HFace> faces;
  std::vector<std::vector<double>> pool;
  double red[1024];
  double phase_ms[4] = {0, 0, 0, 0};

  int nNo() const { return nNo_; }
  size_t nnz() const { return size_t(nnz_); }
  bool is_master() const { return true; }
  size_t mark() const { return pool.size(); }
  void release(size_t m) { pool.resize(m); }
  double* vec(size_t n) { pool.emplace_back(n, 0.0); return pool.back().data(); }

  void zero(size_t n, double* x) { std::memset(x, 0, n*sizeof(double)); }
  void copy(size_t n, const double* x, double* y) { std::memmove(y, x, n*sizeof(double)); }
  void fill(size_t n, double a, double* x) { for (size_t i = 0; i < n; i++) x[i] = a; }
  void axpy(size_t n, double a, const double* x, double* y) { for (size_t i = 0; i < n; i++) y[i] = y[i] + a*x[i]; }
  void scal(size_t n, double a, double* x) { for (size_t i = 0; i < n; i++) x[i] = a*x[i]; }
  void divs(size_t n, double d, double* x) { for (size_t i = 0; i < n; i++) x[i] = x[i]/d; }
  void sub(size_t n, const double* a, const double* b, double* o) { for (size_t i = 0; i < n; i++) o[i] = a[i] - b[i]; }
  void mul_inplace(size_t n, const double* w, double* x) { for (size_t i = 0; i < n; i++) x[i] = w[i]*x[i]; }
  void lin2(size_t n, double* o, double a, const double* x, double b, const double* y) { for (size_t i = 0; i < n; i++) o[i] = a*x[i] + b*y[i]; }
  void axpy2(size_t n, double* X, double a, const double* P, double b, const double* S) { for (size_t i = 0; i < n; i++) X[i] = X[i] + a*P[i] + b*S[i]; }
  void bicg_p_update(size_t n, double* P, const double* R, const double* V, double beta, double omega) { for (size_t i = 0; i < n; i++) P[i] = R[i] + beta*(P[i] - omega*V[i]); }
  void lin_comb(size_t n, double* out, const double* base, int k, const double* V, size_t stride, int j0, const double* coef)
  {
    for (size_t i = 0; i < n; i++) {
      double v = base ? base[i] : 0.0;
      for (int j = 0; j < k; j++) v = v + coef[j]*V[size_t(j0 + j)*stride + i];
      out[i] = v;
    }
  }
  void dots_local(int dof, int count, const double* base, size_t stride, const double* w, int slot0)
  {
    const size_t n = size_t(dof)*nNo_;
    for (int j = 0; j < count; j++) {
      double s = 0.0;
      const double* v = base + size_t(j)*stride;
      for (size_t i = 0; i < n; i++) s = s + v[i]*w[i];
      red[slot0 + j] = s;
    }
  }
  void reduce_begin(int) {}
  void reduce_fetch(int n, double* out) { std::memcpy(out, red, sizeof(double)*n); }
  double dot(int dof, const double* a, const double* b) { dots_local(dof, 1, a, 0, b, 0); return red[0]; }
  double norm(int dof, const double* a) { return std::sqrt(dot(dof, a, a)); }
  void cgs_update_scale(int dof, int k, const double* base, size_t stride, double* w, int slot0)
  {
    const size_t n = size_t(dof)*nNo_;
    double hh = red[slot0 + k];
    for (int j = 0; j < k; j++) hh = hh - red[slot0 + j]*red[slot0 + j];
    const double inv = 1.0/std::sqrt(std::fabs(hh));
    for (size_t i = 0; i < n; i++) {
      double v = w[i];
      for (int j = 0; j < k; j++) v = v - red[slot0 + j]*base[size_t(j)*stride + i];
      w[i] = inv*v;
    }
  }
  void spmv_vv(int dof, const double* K, const double* U, double* KU)
  {
    for (int r = 0; r < nNo_; r++)
      for (int i = 0; i < dof; i++) {
        double acc = 0.0;
        for (int p = rowPtr[r]; p < rowPtr[r+1]; p++)
          for (int j = 0; j < dof; j++) acc = acc + K[size_t(p)*dof*dof + i*dof + j]*U[size_t(col[p])*dof + j];
        KU[size_t(r)*dof + i] = acc;
      }
  }
  void spmv_ss(const double* K, const double* U, double* KU)
  {
    for (int r = 0; r < nNo_; r++) { double a = 0; for (int p = rowPtr[r]; p < rowPtr[r+1]; p++) a = a + K[p]*U[col[p]]; KU[r] = a; }
  }
  void spmv_sv(int dof, const double* K, const double* U, double* KU)
  {
    for (int r = 0; r < nNo_; r++)
      for (int m = 0; m < dof; m++) { double a = 0; for (int p = rowPtr[r]; p < rowPtr[r+1]; p++) a = a + K[size_t(p)*dof + m]*U[col[p]]; KU[size_t(r)*dof + m] = a; }
  }
  void spmv_vs(int dof, const double* K, const double* U, double* KU)
  {
    for (int r = 0; r < nNo_; r++) {
      double a = 0;
      for (int p = rowPtr[r]; p < rowPtr[r+1]; p++) { double t = 0; for (int m = 0; m < dof; m++) t = t + K[size_t(p)*dof + m]*U[size_t(col[p])*dof + m]; a = a + t; }
      KU[r] = a;
    }
  }
  void schur_op(int nsd, const double* Gt, const double* G, const double* L, const double* P, double* GP, double* DGP,
                double* SP, bool coupled)
  {
    spmv_sv(nsd, G, P, GP);
    if (coupled) add_bc_mul(svb200::BCOP_PRE, nsd, GP, GP);
    spmv_vs(nsd, Gt, GP, DGP);
    spmv_ss(L, P, SP);
    axpy(size_t(nNo_), -1.0, DGP, SP);
  }
  int n_faces() const { return int(faces.size()); }
  bool face_coupled(int f) const { return faces[f].coupled; }
  bool face_inc(int f) const { return faces[f].inc; }
  int face_bgrp(int f) const { return faces[f].bGrp; }
  void face_set_inc(int f, bool v) { faces[f].inc = v; }
  void face_set_coupled(int f, bool c, double r) { faces[f].coupled = c; if (c) faces[f].res = r; }
  void bc_pre(int nsd)
  {
    for (auto& f : faces) {
      if (!f.coupled) continue;
      f.nS = 0.0;
      for (int a = 0; a < f.nNo; a++) for (int i = 0; i < std::min(f.dof, nsd); i++) f.nS += f.valM[size_t(a)*f.dof + i]*f.valM[size_t(a)*f.dof + i];
    }
  }
  void add_bc_mul(int op, int dof, const double* X, double* Y)
  {
    for (auto& f : faces) {
      if (!f.coupled) continue;
      const double coef = (op == svb200::BCOP_ADD) ? f.res : -f.res/(1.0 + f.res*f.nS);
      const int m = std::min(f.dof, dof);
      double S = 0.0;
      for (int a = 0; a < f.nNo; a++) for (int i = 0; i < m; i++) S += f.valM[size_t(a)*f.dof + i]*X[size_t(f.glob[a])*dof + i];
      S = coef*S;
      for (int a = 0; a < f.nNo; a++) for (int i = 0; i < m; i++) Y[size_t(f.glob[a])*dof + i] += f.valM[size_t(a)*f.dof + i]*S;
    }
  }
  void precond_diag(int dof, double* Val, double* R, double* W)
  {
    for (int a = 0; a < nNo_; a++) for (int i = 0; i < dof; i++) W[size_t(a)*dof + i] = Val[size_t(diag[a])*dof*dof + i*dof + i];
    for (size_t i = 0; i < size_t(nNo_)*dof; i++) { if (W[i] == 0.0) W[i] = 1.0; W[i] = 1.0/std::sqrt(std::fabs(W[i])); }
    for (auto& f : faces) {
      if (!f.inc || f.bGrp != B200_BC_DIR) continue;
      for (int a = 0; a < f.nNo; a++) for (int i = 0; i < std::min(f.dof, dof); i++) W[size_t(f.glob[a])*dof + i] *= f.val[size_t(a)*f.dof + i];
    }
    for (int r = 0; r < nNo_; r++)
      for (int p = rowPtr[r]; p < rowPtr[r+1]; p++)
        for (int i = 0; i < dof; i++) for (int j = 0; j < dof; j++) {
          double& v = Val[size_t(p)*dof*dof + i*dof + j];
          v = (v*W[size_t(r)*dof + i])*W[size_t(col[p])*dof + j];
        }
    for (size_t i = 0; i < size_t(nNo_)*dof; i++) R[i] = W[i]*R[i];
    for (auto& f : faces) {
      if (!f.coupled) continue;
      for (int a = 0; a < f.nNo; a++) for (int i = 0; i < std::min(f.dof, dof); i++) f.valM[size_t(a)*f.dof + i] = f.val[size_t(a)*f.dof + i]*W[size_t(f.glob[a])*dof + i];
    }
  }
  void precond_rcs(int, double*, double*, double*, double*) { throw std::runtime_error("rcs not in host policy"); }
  void depart(int nsd, const double* Val, double* Gt, double* mK, double* mG, double* mD, double* mL)
  {
    const int D = nsd + 1;
    for (int p = 0; p < nnz_; p++) {
      const double* v = Val + size_t(p)*D*D;
      for (int i = 0; i < nsd; i++) {
        for (int j = 0; j < nsd; j++) mK[size_t(p)*nsd*nsd + i*nsd + j] = v[i*D + j];
        mG[size_t(p)*nsd + i] = v[i*D + nsd];
        mD[size_t(p)*nsd + i] = v[nsd*D + i];
      }
      mL[p] = v[D*D - 1];
      const double* vt = Val + size_t(tpos[p])*D*D;
      for (int i = 0; i < nsd; i++) Gt[size_t(p)*nsd + i] = -vt[i*D + nsd];
    }
  }
  void split_mc(int dof, const double* Ri, double* Rm, double* Rc)
  {
    for (int a = 0; a < nNo_; a++) { for (int l = 0; l < dof-1; l++) Rm[size_t(a)*(dof-1) + l] = Ri[size_t(a)*dof + l]; Rc[a] = Ri[size_t(a)*dof + dof-1]; }
  }
  void join_mc(int dof, const double* Rm, const double* Rc, double* Ri)
  {
    for (int a = 0; a < nNo_; a++) { for (int l = 0; l < dof-1; l++) Ri[size_t(a)*dof + l] = Rm[size_t(a)*(dof-1) + l]; Ri[size_t(a)*dof + dof-1] = Rc[a]; }
  }
  void phase_mark(int, double) {}
};

std::string g_err;

} // namespace

extern "C" {

const char* hl_last_error() { return g_err.c_str(); }

// svfsiplus_b200/csrc/lhs_layout.hpp: the host-side tables of the device transport and of the row-tile kernels
// lists: nReq lists concatenated (req_n[i] entries each).  Outputs sized by the caller: node / ptr (<= nNo + 1), src_req / src_pos (total).
int hl_halo_sources(int nNo, int nReq, const int* req_n, const int* req_ptr, int* nh, int* node, int* ptr, int* src_req, int* src_pos)
{
  try {
    std::vector<std::vector<int>> lists(nReq);
    size_t o = 0;
    for (int i = 0; i < nReq; i++) { lists[i].assign(req_ptr + o, req_ptr + o + req_n[i]); o += size_t(req_n[i]); }
    const svb200::HaloSources h = svb200::halo_source_lists(nNo, lists);
    *nh = int(h.node.size());
    std::copy(h.node.begin(), h.node.end(), node);
    std::copy(h.ptr.begin(), h.ptr.end(), ptr);
    std::copy(h.src_req.begin(), h.src_req.end(), src_req);
    std::copy(h.src_pos.begin(), h.src_pos.end(), src_pos);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
// returns the number of tiles (0: a row does not fit); tile_row must hold nNo + 1 ints
int hl_row_tiles(int nNo, const int* rowPtr, int ovA, int ovB, int max_rows, int cap, int* tile_row, int* tile_at)
{
  const int cuts[4] = {0, ovA, ovB, nNo};
  const std::vector<int> rp(rowPtr, rowPtr + nNo + 1);
  const std::vector<int> tr = svb200::row_tiles(rp, cuts, max_rows, cap, tile_at);
  std::copy(tr.begin(), tr.end(), tile_row);
  return tr.empty() ? 0 : int(tr.size()) - 1;
}

// Single-rank solve with the product's krylov.hpp driven by the serial host policy.
// faces: arrays of length nFaces; glob/val concatenated.  ls = 12 doubles as in oracle/ref.py.
// out = {RI.suc, RI.itr, RI.iNorm, RI.fNorm, RI.dB, GM.itr, CG.itr, Resm, Resc}
int hl_solve(int nNo, int nnz, const int* rowPtr, const int* col, int dof, double* R, double* Val,
             const double* ls, int prec, int nFaces, const int* f_nNo, const int* f_dof, const int* f_bGrp,
             const int* f_glob, const double* f_val, const int* incL, const double* res, double* out)
{
  try {
    HostOps ops;
    ops.nNo_ = nNo; ops.nnz_ = nnz;
    ops.rowPtr.assign(rowPtr, rowPtr + nNo + 1);
    ops.col.assign(col, col + nnz);
    ops.diag.assign(nNo, -1);
    ops.tpos.assign(nnz, -1);
    for (int r = 0; r < nNo; r++)
      for (int p = rowPtr[r]; p < rowPtr[r+1]; p++) {
        if (col[p] == r) ops.diag[r] = p;
        const int c = col[p];
        for (int q = rowPtr[c]; q < rowPtr[c+1]; q++) if (col[q] == r) { ops.tpos[p] = q; break; }
      }
    size_t og = 0, ov = 0;
    for (int f = 0; f < nFaces; f++) {
      HFace hf;
      hf.nNo = f_nNo[f]; hf.dof = f_dof[f]; hf.bGrp = f_bGrp[f];
      hf.glob.assign(f_glob + og, f_glob + og + hf.nNo);
      hf.val.assign(f_val + ov, f_val + ov + size_t(hf.nNo)*hf.dof);
      hf.valM.assign(size_t(hf.nNo)*hf.dof, 0.0);
      og += hf.nNo; ov += size_t(hf.nNo)*hf.dof;
      ops.faces.push_back(std::move(hf));
    }
    svb200::Ls L;
    L.LS_type = int(ls[0]);
    L.RI.relTol = ls[1]; L.RI.absTol = ls[2]; L.RI.mItr = int(ls[3]); L.RI.sD = int(ls[4]);
    L.GM.relTol = ls[5]; L.GM.absTol = ls[6]; L.GM.mItr = int(ls[7]); L.GM.sD = int(ls[8]);
    L.CG.relTol = ls[9]; L.CG.absTol = ls[10]; L.CG.mItr = int(ls[11]);
    svb200::solve(ops, L, dof, prec, R, Val, incL, res);
    out[0] = L.RI.suc; out[1] = L.RI.itr; out[2] = L.RI.iNorm; out[3] = L.RI.fNorm; out[4] = L.RI.dB;
    out[5] = L.GM.itr; out[6] = L.CG.itr; out[7] = L.Resm; out[8] = L.Resc;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

}
