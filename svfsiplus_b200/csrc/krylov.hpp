// krylov.hpp — host-side control flow of the FSILS-equivalent solvers, written once against an
// `Ops` policy that owns the vectors and launches the kernels.  The product instantiates it with
// CudaOps (ops_cuda.cuh) and nothing else; tests/hostlogic instantiates it with a serial host policy
// so that iteration logic (Givens, restarts, Gram system, convergence tests, iteration counters) can
// be checked against the compiled reference without a GPU.
//
// Algorithms follow the reference (file:line cited at each routine); data never leaves the device in
// the product: the only host<->device traffic inside a solve is the handful of reduced scalars each
// iteration needs for its convergence test.
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <limits>
#include <stdexcept>
#include <type_traits>
#include <vector>

#include "svb200.h"

namespace svb200 {

struct SubLs {                 // FSILS_subLsType, liner_solver/fils_struct.hpp:211-257
  bool suc = false;
  int mItr = 0, sD = 0, itr = 0;
  double absTol = 0, relTol = 0, iNorm = 0, fNorm = 0, dB = 0, callD = 0;
};

struct Ls {                    // FSILS_lsType, fils_struct.hpp:259-279
  int LS_type = B200_LS_GMRES;
  int Resm = 0, Resc = 0;
  SubLs GM, CG, RI;
};

inline double wall_s()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

enum { BCOP_ADD = 0, BCOP_PRE = 1 };   // fils_struct.hpp:64 BcopType

// ---------------------------------------------------------------------------------------------
// Small dense solve used by the NS solver's Gram system: Jacobi-scaled Gaussian elimination with
// partial pivoting (liner_solver/ge.cpp:40-140).  A is nV x nV column-major, only the leading N x N
// part is used.  Returns false when the system is (numerically) singular; B is then zeroed.
// ---------------------------------------------------------------------------------------------
inline bool ge_solve(int nV, int N, const std::vector<double>& A, std::vector<double>& B)
{
  const double tol = std::numeric_limits<double>::denorm_min();
  const double eps = std::numeric_limits<double>::epsilon();
  auto a = [&](int i, int j) { return A[i + size_t(j)*nV]; };
  if (N <= 0) return false;
  std::vector<double> W(N);
  for (int i = 0; i < N; i++) {
    if (std::fabs(a(i,i)) < tol) { std::fill(B.begin(), B.end(), 0.0); return false; }
    W[i] = 1.0 / std::sqrt(std::fabs(a(i,i)));
  }
  const int ld = N;
  std::vector<double> Cm(size_t(N)*(N+1));
  auto c = [&](int i, int j) -> double& { return Cm[i + size_t(j)*ld]; };
  for (int i = 0; i < N; i++) {
    for (int j = 0; j < N; j++) c(i,j) = W[i]*W[j]*a(i,j);
    c(i,N) = W[i]*B[i];
  }
  if (N == 1) {
    B[0] = c(0,1) / c(0,0);
    B[0] = B[0]*W[0];
    return true;
  }
  if (N == 2) {
    double pivot = c(0,0)*c(1,1) - c(1,0)*c(0,1);
    if (std::fabs(pivot) < eps) { std::fill(B.begin(), B.end(), 0.0); return false; }
    double b0 = (c(0,2)*c(1,1) - c(1,2)*c(0,1)) / pivot;
    double b1 = (c(1,2)*c(0,0) - c(0,2)*c(1,0)) / pivot;
    B[0] = W[0]*b0;
    B[1] = W[1]*b1;
    return true;
  }
  for (int m = 0; m < N-1; m++) {
    int ipv = m;
    double pivot = std::fabs(c(m,m));
    for (int i = m+1; i < N; i++) {
      if (std::fabs(c(i,m)) > pivot) { ipv = i; pivot = std::fabs(c(i,m)); }
    }
    if (std::fabs(pivot) < eps) { std::fill(B.begin(), B.end(), 0.0); return false; }
    if (ipv != m) {
      for (int j = m; j < N+1; j++) std::swap(c(m,j), c(ipv,j));
    }
    for (int i = m+1; i < N; i++) {
      double s = c(i,m) / c(m,m);
      c(i,m) = 0.0;
      for (int j = m+1; j < N+1; j++) c(i,j) = c(i,j) - s*c(m,j);
    }
  }
  for (int j = N-1; j >= 0; j--) {
    for (int i = j+1; i < N; i++) c(j,N) = c(j,N) - c(j,i)*c(i,N);
    c(j,N) = c(j,N) / c(j,j);
  }
  for (int i = 0; i < N; i++) B[i] = W[i]*c(i,N);
  return true;
}

// ---------------------------------------------------------------------------------------------
// One column of the Arnoldi bookkeeping, written ONCE for the host (Hessenberg::finish_column below, which the serial test policy
// runs against the compiled reference) and for the device (k_gmres_givens of the device-resident loop): `col` = column i of the
// Hessenberg matrix holding the reduced dots <u_j, w>, j = 0..i+1 (entry i+1 = <w,w>); c, s: the rotations so far; err_i, err_i1:
// the residual estimates err(i), err(i+1).  Every product and sum is rounded separately on both sides (the device spells it __dmul_rn / __dadd_rn so
// that nvcc cannot contract them into FMAs; the host build has no FMA contraction), so the two give the same bits.
// Returns |err(i+1)|.
// ---------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
#define SVB_KR_HD __host__ __device__ __forceinline__
#else
#define SVB_KR_HD inline
#endif
SVB_KR_HD double kr_mul(double a, double b)
{
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a*b;
#endif
}
SVB_KR_HD double kr_add(double a, double b)
{
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
SVB_KR_HD double givens_finish_column(double* col, double* c, double* s, double& err_i, double& err_i1, int i)
{
  for (int j = 0; j <= i; j++) col[i+1] = kr_add(col[i+1], -kr_mul(col[j], col[j]));
  col[i+1] = sqrt(fabs(col[i+1]));
  for (int j = 0; j <= i-1; j++) {
    const double tmp = kr_add(kr_mul(c[j], col[j]), kr_mul(s[j], col[j+1]));
    col[j+1] = kr_add(kr_mul(-s[j], col[j]), kr_mul(c[j], col[j+1]));
    col[j] = tmp;
  }
  const double tmp = sqrt(kr_add(kr_mul(col[i], col[i]), kr_mul(col[i+1], col[i+1])));
  c[i] = col[i] / tmp;
  s[i] = col[i+1] / tmp;
  col[i] = tmp;
  col[i+1] = 0.0;
  err_i1 = kr_mul(-s[i], err_i);
  err_i = kr_mul(c[i], err_i);
  return fabs(err_i1);
}

// ---------------------------------------------------------------------------------------------
// Arnoldi / Givens bookkeeping shared by gmres_v and the NS solver's inner gmres
// (liner_solver/gmres.cpp:550-612 and :202-262).  h is (sD+1) x sD column-major.
// ---------------------------------------------------------------------------------------------
struct Hessenberg {
  int sD;
  std::vector<double> h, c, s, err, y;
  explicit Hessenberg(int sD_) : sD(sD_), h(size_t(sD_+1)*sD_, 0.0), c(sD_, 0.0), s(sD_, 0.0), err(sD_+1, 0.0), y(sD_, 0.0) {}
  double& H(int i, int j) { return h[i + size_t(j)*(sD+1)]; }

  // Column i has been filled with the reduced dots <u_j, w>, j = 0..i+1 (entry i+1 = <w,w>).
  // Performs the Pythagorean norm, Givens rotations and the residual-estimate update; returns
  // |err(i+1)|.
  double finish_column(int i) { return givens_finish_column(&H(0,i), c.data(), s.data(), err[i], err[i+1], i); }

  void back_substitute(int last_i)
  {
    for (int i = 0; i <= last_i; i++) y[i] = err[i];
    for (int j = last_i; j >= 0; j--) {
      for (int k = j+1; k <= last_i; k++) y[j] = y[j] - H(j,k)*y[k];
      y[j] = y[j] / H(j,j);
    }
  }
};

// Policies that keep the CG scalars on the device (CudaOps) expose cg_device(); the serial test policy does not.
template <class T, class = void> struct has_cg_device : std::false_type {};
template <class T> struct has_cg_device<T, std::void_t<decltype(&T::cg_batch)>> : std::true_type {};

template <class T, class = void> struct has_gmres_device : std::false_type {};
template <class T> struct has_gmres_device<T, std::void_t<decltype(&T::gmres_device_ok)>> : std::true_type {};
template <class T, class = void> struct has_dots_reduce : std::false_type {};
template <class T> struct has_dots_reduce<T, std::void_t<decltype(&T::dots_reduce)>> : std::true_type {};

template <class Ops> inline bool any_coupled(Ops& ops)
{
  for (int f = 0; f < ops.n_faces(); f++) if (ops.face_coupled(f)) return true;
  return false;
}

// One Arnoldi step shared by both GMRES flavours: w = u[i+1] already holds K u[i] (+bc terms).
// Queues the (i+2) local dots, reduces them, queues the Gram-Schmidt update + normalisation (which
// reads the reduced dots on the device) and returns the reduced column on the host.
// the device part of an Arnoldi step alone (dots + all-reduce + Gram-Schmidt update), for the device-resident loop
template <class Ops>
inline void arnoldi_enqueue(Ops& ops, int dof, double* u, size_t stride, int i)
{
  double* w = u + size_t(i+1)*stride;
  if constexpr (has_dots_reduce<Ops>::value) {
    ops.dots_reduce(dof, i+2, u, stride, w, 0);
  } else {
    ops.dots_local(dof, i+2, u, stride, w, 0);
    ops.reduce_begin(i+2);
  }
  ops.cgs_update_scale(dof, i+1, u, stride, w, 0);
}

template <class Ops>
inline void arnoldi_orthogonalise(Ops& ops, int dof, double* u, size_t stride, int i, Hessenberg& hs)
{
  double* w = u + size_t(i+1)*stride;
  if constexpr (has_dots_reduce<Ops>::value) {
    ops.dots_reduce(dof, i+2, u, stride, w, 0);        // local dots + all-reduce in one launch where the policy can
  } else {
    ops.dots_local(dof, i+2, u, stride, w, 0);
    ops.reduce_begin(i+2);
  }
  ops.cgs_update_scale(dof, i+1, u, stride, w, 0);     // consumes the reduced slots on the device
  ops.reduce_fetch(i+2, &hs.H(0,i));
}

// ---------------------------------------------------------------------------------------------
// gmres_v: restarted GMRES for vector problems, solution returned in R
// (liner_solver/gmres.cpp:450-631).  `u` is caller-provided basis storage of (sD+1) vectors.
// ---------------------------------------------------------------------------------------------
template <class Ops>
void gmres_v(Ops& ops, SubLs& ls, int dof, const double* Val, double* R)
{
  const size_t n = size_t(dof)*ops.nNo();
  auto mk = ops.mark();
  double* X = ops.vec(n);
  double* u = ops.vec(n*(ls.sD+1));
  Hessenberg hs(ls.sD);

  ls.callD = wall_s();
  ls.suc = false;
  double eps = ops.norm(dof, R);
  ls.iNorm = eps;
  ls.fNorm = eps;
  eps = std::max(ls.absTol, ls.relTol*eps);
  ls.itr = 0;
  int last_i = 0;

  ops.bc_pre(dof);

  if (ls.iNorm <= ls.absTol) {
    ls.callD = std::numeric_limits<double>::epsilon();
    ls.dB = 0.0;
    ops.release(mk);
    return;
  }
  ops.zero(n, X);

  for (int l = 0; l < ls.mItr; l++) {
    ls.dB = ls.fNorm;
    ls.itr++;
    double* u0 = u;
    if (l == 0) {
      ops.copy(n, R, u0);                       // X = 0  =>  K X + bc = 0 exactly
    } else {
      ops.spmv_vv(dof, Val, X, u0);
      ops.add_bc_mul(BCOP_ADD, dof, X, u0);
      ops.sub(n, R, u0, u0);                    // u0 = R - u0
    }
    // (the reference's `flag` is hard-wired false here: no BCOP_PRE in gmres_v, gmres.cpp:462)
    hs.err[0] = ops.norm(dof, u0);
    ops.divs(n, hs.err[0], u0);                 // u0 = u0 / err0 (a true division, like the reference)

    bool on_device = false;
    if constexpr (has_gmres_device<Ops>::value) on_device = ops.gmres_device_ok() && ls.sD <= Ops::kGivensMaxHost;
    if constexpr (has_gmres_device<Ops>::value) {
      if (on_device) {
        bool suc = false;
        last_i = ops.gmres_device_cycle(ls.sD, eps, hs.err[0], [&](int i) {
          double* ui = u + size_t(i)*n;
          double* ui1 = u + size_t(i+1)*n;
          ops.spmv_vv(dof, Val, ui, ui1);
          ops.add_bc_mul(BCOP_ADD, dof, ui, ui1);
          arnoldi_enqueue(ops, dof, u, n, i);
        }, hs, suc);
        ls.itr += last_i + 1;
        if (suc) ls.suc = true;
      }
    }
    if (!on_device)
    for (int i = 0; i < ls.sD; i++) {
      ls.itr++;
      last_i = i;
      double* ui = u + size_t(i)*n;
      double* ui1 = u + size_t(i+1)*n;
      ops.spmv_vv(dof, Val, ui, ui1);
      ops.add_bc_mul(BCOP_ADD, dof, ui, ui1);
      arnoldi_orthogonalise(ops, dof, u, n, i, hs);
      if (hs.finish_column(i) < eps) {
        ls.suc = true;
        break;
      }
    }
    if (last_i >= ls.sD) last_i = ls.sD - 1;
    hs.back_substitute(last_i);
    ops.lin_comb(n, X, X, last_i+1, u, n, 0, hs.y.data());
    ls.fNorm = std::fabs(hs.err[last_i+1]);
    if (ls.suc) break;
  }

  ops.copy(n, X, R);
  ls.callD = wall_s() - ls.callD;
  ls.dB = 10.0 * std::log(ls.fNorm / ls.dB);
  ops.release(mk);
}

// ---------------------------------------------------------------------------------------------
// gmres: inner GMRES of the NS solver, X = K^-1 R with zero initial guess and the resistance
// preconditioner BCOP_PRE applied when a face is coupled (liner_solver/gmres.cpp:91-281).
// ls.itr and ls.callD accumulate across calls (the NS solver resets them).
// ---------------------------------------------------------------------------------------------
template <class Ops>
void gmres_inner(Ops& ops, SubLs& ls, int dof, const double* Val, const double* R, double* X)
{
  const size_t n = size_t(dof)*ops.nNo();
  auto mk = ops.mark();
  double* u = ops.vec(n*(ls.sD+1));
  Hessenberg hs(ls.sD);
  const bool coupled = any_coupled(ops);

  double time = wall_s();
  ls.suc = false;
  double eps = 0.0;
  int last_i = 0;
  ops.zero(n, X);

  for (int l = 0; l < ls.mItr; l++) {
    double* u0 = u;
    if (l == 0) {
      ops.copy(n, R, u0);
    } else {
      ops.spmv_vv(dof, Val, X, u0);
      ops.add_bc_mul(BCOP_ADD, dof, X, u0);
      ls.itr++;
      ops.sub(n, R, u0, u0);
    }
    if (coupled) ops.add_bc_mul(BCOP_PRE, dof, u0, u0);

    hs.err[0] = ops.norm(dof, u0);
    if (l == 0) {
      eps = hs.err[0];
      if (eps <= ls.absTol) {
        ls.callD = std::numeric_limits<double>::epsilon();
        ls.dB = 0.0;
        ops.release(mk);
        return;
      }
      ls.iNorm = eps;
      ls.fNorm = eps;
      eps = std::max(ls.absTol, ls.relTol*eps);
    }
    ls.dB = ls.fNorm;
    ops.divs(n, hs.err[0], u0);

    bool on_device = false;
    if constexpr (has_gmres_device<Ops>::value) on_device = ops.gmres_device_ok() && ls.sD <= Ops::kGivensMaxHost;
    if constexpr (has_gmres_device<Ops>::value) {
      if (on_device) {
        bool suc = false;
        last_i = ops.gmres_device_cycle(ls.sD, eps, hs.err[0], [&](int i) {
          double* ui = u + size_t(i)*n;
          double* ui1 = u + size_t(i+1)*n;
          ops.spmv_vv(dof, Val, ui, ui1);
          ops.add_bc_mul(BCOP_ADD, dof, ui, ui1);
          if (coupled) ops.add_bc_mul(BCOP_PRE, dof, ui1, ui1);
          arnoldi_enqueue(ops, dof, u, n, i);
        }, hs, suc);
        ls.itr += last_i + 1;
        if (suc) ls.suc = true;
      }
    }
    if (!on_device)
    for (int i = 0; i < ls.sD; i++) {
      last_i = i;
      double* ui = u + size_t(i)*n;
      double* ui1 = u + size_t(i+1)*n;
      ops.spmv_vv(dof, Val, ui, ui1);
      ops.add_bc_mul(BCOP_ADD, dof, ui, ui1);
      ls.itr++;
      if (coupled) ops.add_bc_mul(BCOP_PRE, dof, ui1, ui1);
      arnoldi_orthogonalise(ops, dof, u, n, i, hs);
      if (hs.finish_column(i) < eps) {
        ls.suc = true;
        break;
      }
    }
    if (last_i >= ls.sD) last_i = ls.sD - 1;
    hs.back_substitute(last_i);
    ops.lin_comb(n, X, X, last_i+1, u, n, 0, hs.y.data());
    ls.fNorm = std::fabs(hs.err[last_i+1]);
    if (ls.suc) break;
  }
  ls.callD = wall_s() - time + ls.callD;
  ls.dB = 10.0 * std::log(ls.fNorm / ls.dB);
  ops.release(mk);
}

// ---------------------------------------------------------------------------------------------
// cgrad_v: conjugate gradients on the (Jacobi-scaled) vector system (liner_solver/cgrad.cpp:166-246).
// ---------------------------------------------------------------------------------------------
template <class Ops>
void cgrad_v(Ops& ops, SubLs& ls, int dof, const double* K, double* R)
{
  const size_t n = size_t(dof)*ops.nNo();
  auto mk = ops.mark();
  double* P = ops.vec(n);
  double* KP = ops.vec(n);
  double* X = ops.vec(n);

  ls.callD = wall_s();
  ls.suc = false;
  ls.iNorm = ops.norm(dof, R);
  double eps = std::pow(std::max(ls.absTol, ls.relTol*ls.iNorm), 2.0);
  double errO = ls.iNorm*ls.iNorm;
  double err = errO;
  ops.zero(n, X);
  ops.copy(n, R, P);
  int last_i = 0;

  if constexpr (has_cg_device<Ops>::value) {
    ops.cg_device(ls, dof, R, X, P, KP, [&](const double* p, double* kp) { ops.spmv_vv(dof, K, p, kp); }, last_i, err, errO);
  } else {
  for (int i = 0; i < ls.mItr; i++) {
    last_i = i;
    if (err < eps) { ls.suc = true; break; }
    errO = err;
    ops.spmv_vv(dof, K, P, KP);
    double alpha = errO / ops.dot(dof, P, KP);
    ops.axpy(n, alpha, P, X);
    ops.axpy(n, -alpha, KP, R);
    err = ops.norm(dof, R);
    err = err*err;
    ops.axpy(n, errO/err, R, P);      // P = P + (errO/err) R
    ops.scal(n, err/errO, P);         // P = (err/errO) P
  }
  }
  ops.copy(n, X, R);
  ls.itr = last_i;
  ls.fNorm = std::sqrt(err);
  ls.callD = wall_s() - ls.callD;
  ls.dB = (errO < std::numeric_limits<double>::epsilon()) ? 0.0 : 5.0*std::log(err/errO);
  ops.release(mk);
}

// ---------------------------------------------------------------------------------------------
// schur: CG on the pressure Schur complement  [L + Gt G]  (liner_solver/cgrad.cpp:50-160).
// D = Gt (nsd x nnz), G = mG (nsd x nnz), L = mL (nnz); R (nNo) in: rhs, out: solution.
// ---------------------------------------------------------------------------------------------
template <class Ops>
void schur(Ops& ops, SubLs& ls, int nsd, const double* D, const double* G, const double* L, double* R)
{
  const size_t nn = ops.nNo();
  auto mk = ops.mark();
  double* X = ops.vec(nn);
  double* P = ops.vec(nn);
  double* SP = ops.vec(nn);
  double* DGP = ops.vec(nn);
  double* GP = ops.vec(nn*nsd);
  const bool coupled = any_coupled(ops);

  double time = wall_s();
  ls.suc = false;
  ls.iNorm = ops.norm(1, R);
  double eps = std::pow(std::max(ls.absTol, ls.relTol*ls.iNorm), 2.0);
  double errO = ls.iNorm*ls.iNorm;
  double err = errO;
  ops.zero(nn, X);
  ops.copy(nn, R, P);
  int last_i = 0;

  if constexpr (has_cg_device<Ops>::value) {
    ops.cg_device(ls, 1, R, X, P, SP, [&](const double* p, double* sp) { ops.schur_op(nsd, D, G, L, p, GP, DGP, sp, coupled); },
                  last_i, err, errO);
  } else {
  for (int i = 0; i < ls.mItr; i++) {
    last_i = i;
    if (err < eps) { ls.suc = true; break; }
    errO = err;
    ops.schur_op(nsd, D, G, L, P, GP, DGP, SP, coupled);     // SP = L P - D (G P [+ PRE])
    double alpha = errO / ops.dot(1, P, SP);
    ops.axpy(nn, alpha, P, X);
    ops.axpy(nn, -alpha, SP, R);
    err = ops.norm(1, R);
    err = err*err;
    ops.axpy(nn, errO/err, R, P);
    ops.scal(nn, err/errO, P);
  }
  }
  ops.copy(nn, X, R);
  ls.fNorm = std::sqrt(err);
  ls.callD = wall_s() - time + ls.callD;
  ls.itr = ls.itr + last_i;
  ls.dB = (errO < std::numeric_limits<double>::epsilon()) ? 0.0 : 5.0*std::log(err/errO);
  ops.release(mk);
}

// ---------------------------------------------------------------------------------------------
// bicgsv: BiCGStab (liner_solver/bicgs.cpp:49-144).
// ---------------------------------------------------------------------------------------------
template <class Ops>
void bicgs_v(Ops& ops, SubLs& ls, int dof, const double* K, double* R)
{
  const size_t n = size_t(dof)*ops.nNo();
  auto mk = ops.mark();
  double* P = ops.vec(n);
  double* Rh = ops.vec(n);
  double* X = ops.vec(n);
  double* V = ops.vec(n);
  double* S = ops.vec(n);
  double* T = ops.vec(n);

  ls.callD = wall_s();
  ls.suc = false;
  double err = ops.norm(dof, R);
  double errO = err;
  ls.iNorm = err;
  double eps = std::max(ls.absTol, ls.relTol*err);
  double rho = err*err;
  double beta = rho;
  (void)beta;
  ops.zero(n, X);
  ops.copy(n, R, P);
  ops.copy(n, R, Rh);
  int i_itr = 1;

  for (int i = 0; i < ls.mItr; i++) {
    if (err < eps) { ls.suc = true; break; }
    ops.spmv_vv(dof, K, P, V);
    double alpha = rho / ops.dot(dof, Rh, V);
    ops.lin2(n, S, 1.0, R, -alpha, V);                 // S = R - alpha V
    ops.spmv_vv(dof, K, S, T);
    double omega = ops.norm(dof, T);
    omega = ops.dot(dof, T, S) / (omega*omega);
    ops.axpy2(n, X, alpha, P, omega, S);               // X = X + alpha P + omega S
    ops.lin2(n, R, 1.0, S, -omega, T);                 // R = S - omega T
    errO = err;
    err = ops.norm(dof, R);
    double rhoO = rho;
    rho = ops.dot(dof, R, Rh);
    beta = rho*alpha / (rhoO*omega);
    ops.bicg_p_update(n, P, R, V, beta, omega);        // P = R + beta (P - omega V)
    i_itr++;
  }
  ops.copy(n, X, R);
  ls.itr = i_itr - 1;
  ls.fNorm = err;
  ls.callD = wall_s() - ls.callD;
  ls.dB = (errO < std::numeric_limits<double>::epsilon()) ? 0.0 : 10.0*std::log(err/errO);
  ops.release(mk);
}

// ---------------------------------------------------------------------------------------------
// ns_solver: bi-partitioned solver for A = [K D; -G L] (liner_solver/ns_solver.cpp:167-500).
// Val is the scaled dof x dof block matrix, Ri (dof x nNo) in: rhs, out: solution.
// ---------------------------------------------------------------------------------------------
template <class Ops>
void ns_solver(Ops& ops, Ls& ls, int dof, const double* Val, double* Ri)
{
  const int nNo = ops.nNo();
  const size_t nnz = ops.nnz();
  const int nsd = dof - 1;
  const int iBmax = ls.RI.mItr;
  const int nB = 2*iBmax;
  const size_t nm = size_t(nsd)*nNo;   // momentum vector length
  const size_t nc = size_t(nNo);       // continuity vector length

  auto mk = ops.mark();
  double* Rm = ops.vec(nm);
  double* Rmi = ops.vec(nm);
  double* Rc = ops.vec(nc);
  double* Rci = ops.vec(nc);
  double* U = ops.vec(nm*iBmax);
  double* MU = ops.vec(nm*nB);
  double* P = ops.vec(nc*iBmax);
  double* MP = ops.vec(nc*nB);
  double* Gt = ops.vec(size_t(nsd)*nnz);
  double* mK = ops.vec(size_t(nsd)*nsd*nnz);
  double* mG = ops.vec(size_t(nsd)*nnz);
  double* mD = ops.vec(size_t(nsd)*nnz);
  double* mL = ops.vec(nnz);

  std::vector<double> A(size_t(nB)*nB, 0.0), B(nB, 0.0), xB(nB, 0.0), oldxB(nB, 0.0), tmp(size_t(nB)*nB + nB, 0.0);

  ops.split_mc(dof, Ri, Rmi, Rci);           // Rmi = Ri(0:nsd-1,:), Rci = Ri(dof-1,:)
  ops.copy(nm, Rmi, Rm);
  ops.copy(nc, Rci, Rc);

  {
    // eps = sqrt(|Rm|^2 + |Rc|^2): both reductions in one round trip
    ops.dots_local(nsd, 1, Rm, 0, Rm, 0);
    ops.dots_local(1, 1, Rc, 0, Rc, 1);
    ops.reduce_begin(2);
    double t2[2];
    ops.reduce_fetch(2, t2);
    // the reference squares the square roots: pow(sqrt(a),2) + pow(sqrt(b),2)
    double a = std::sqrt(t2[0]), b = std::sqrt(t2[1]);
    double eps0 = std::sqrt(a*a + b*b);
    ls.RI.iNorm = eps0;
    ls.RI.fNorm = eps0*eps0;
  }
  double eps = ls.RI.iNorm;
  ls.CG.callD = 0.0;
  ls.GM.callD = 0.0;
  ls.RI.callD = wall_s();
  ls.CG.itr = 0;
  ls.GM.itr = 0;
  ls.RI.suc = false;
  eps = std::max(ls.RI.absTol, ls.RI.relTol*eps);

  ops.depart(nsd, Val, Gt, mK, mG, mD, mL);
  ops.bc_pre(nsd);

  int iB = 0, iBB = 0, i_count = 0;
  for (int i = 0; i < ls.RI.mItr; i++) {
    iB = 2*i;
    iBB = 2*i + 1;
    ls.RI.dB = ls.RI.fNorm;
    i_count = i;

    double* Ui = U + size_t(i)*nm;
    double* Pi = P + size_t(i)*nc;
    double* MU_iB = MU + size_t(iB)*nm;
    double* MU_iBB = MU + size_t(iBB)*nm;
    double* MP_iB = MP + size_t(iB)*nc;
    double* MP_iBB = MP + size_t(iBB)*nc;

    // U = K^-1 Rm
    gmres_inner(ops, ls.GM, nsd, mK, Rm, Ui);
    // P = Rc - D U
    ops.spmv_vs(nsd, mD, Ui, Pi);
    ops.sub(nc, Rc, Pi, Pi);
    // P = [L + Gt G]^-1 P
    schur(ops, ls.CG, nsd, Gt, mG, mL, Pi);
    // MU1 = G P ; MU2 = Rm - G P
    ops.spmv_sv(nsd, mG, Pi, MU_iB);
    ops.sub(nm, Rm, MU_iB, MU_iBB);
    // U = K^-1 [Rm - G P]
    gmres_inner(ops, ls.GM, nsd, mK, MU_iBB, Ui);
    // MU2 = K U (+bc)
    ops.spmv_vv(nsd, mK, Ui, MU_iBB);
    ops.add_bc_mul(BCOP_ADD, nsd, Ui, MU_iBB);
    // MP1 = L P ; MP2 = D U
    ops.spmv_ss(mL, Pi, MP_iB);
    ops.spmv_vs(nsd, mD, Ui, MP_iBB);

    // Gram rows iB, iBB (+ right-hand sides): local dots, ONE reduction.
    // slot layout: for k in {iB,iBB}: [ (MU_j.MU_k) j=0..k | MU_k.Rmi | (MP_j.MP_k) j=0..k | MP_k.Rci ]
    int c = 0;
    for (int k = iB; k <= iBB; k++) {
      ops.dots_local(nsd, k+1, MU, nm, MU + size_t(k)*nm, c);         c += k+1;
      ops.dots_local(nsd, 1, MU + size_t(k)*nm, 0, Rmi, c);           c += 1;
      ops.dots_local(1, k+1, MP, nc, MP + size_t(k)*nc, c);           c += k+1;
      ops.dots_local(1, 1, MP + size_t(k)*nc, 0, Rci, c);             c += 1;
    }
    ops.reduce_begin(c);
    std::vector<double> red(c);
    ops.reduce_fetch(c, red.data());

    c = 0;
    int t = 0;
    for (int k = iB; k <= iBB; k++) {
      const double* mu = &red[c];
      const double* mp = &red[c + k + 2];
      for (int j = 0; j <= k; j++) tmp[t++] = mu[j] + mp[j];
      tmp[t++] = mu[k+1] + mp[k+1];
      c += 2*(k+2);
    }
    t = 0;
    for (int k = iB; k <= iBB; k++) {
      for (int j = 0; j <= k; j++) {
        A[j + size_t(k)*nB] = tmp[t];
        A[k + size_t(j)*nB] = tmp[t];
        t++;
      }
      B[k] = tmp[t++];
    }
    xB = B;

    if (ge_solve(nB, iBB+1, A, xB)) {
      oldxB = xB;
    } else {
      if (ops.is_master()) throw std::runtime_error("FSILS: Singular matrix detected");
      xB = oldxB;
      if (i > 0) { iB -= 2; iBB -= 2; }
      break;
    }

    double sum = 0.0;
    for (int j = 0; j <= iBB; j++) sum += xB[j]*B[j];
    ls.RI.fNorm = std::pow(ls.RI.iNorm, 2.0) - sum;

    if (ls.RI.fNorm < eps*eps) {
      ls.RI.suc = true;
      break;
    }
    // Rm = Rmi - sum_j xB_j MU_j ; Rc = Rci - sum_j xB_j MP_j
    std::vector<double> neg(iBB+1);
    for (int j = 0; j <= iBB; j++) neg[j] = -xB[j];
    ops.lin_comb(nm, Rm, Rmi, iBB+1, MU, nm, 0, neg.data());
    ops.lin_comb(nc, Rc, Rci, iBB+1, MP, nc, 0, neg.data());
  }

  if (i_count >= ls.RI.mItr) {
    ls.RI.itr = ls.RI.mItr;
  } else {
    ls.RI.itr = i_count;
    std::vector<double> neg(iBB+1);
    for (int j = 0; j <= iBB; j++) neg[j] = -xB[j];
    ops.lin_comb(nc, Rc, Rci, iBB+1, MP, nc, 0, neg.data());
  }

  {
    double nrc = ops.norm(1, Rc);
    ls.Resc = static_cast<int>(100.0 * std::pow(nrc, 2.0) / ls.RI.fNorm);
    ls.Resm = 100 - ls.Resc;
  }

  // Rmi = sum_i xB(2i+1) U_i ; Rci = sum_i xB(2i) P_i
  {
    std::vector<double> cu(ls.RI.itr+1), cp(ls.RI.itr+1);
    for (int i = 0; i <= ls.RI.itr; i++) { cu[i] = xB[2*i+1]; cp[i] = xB[2*i]; }
    ops.lin_comb(nm, Rmi, nullptr, ls.RI.itr+1, U, nm, 0, cu.data());
    ops.lin_comb(nc, Rci, nullptr, ls.RI.itr+1, P, nc, 0, cp.data());
  }

  ls.RI.callD = wall_s() - ls.RI.callD;
  ls.RI.dB = 5.0 * std::log(ls.RI.fNorm / ls.RI.dB);

  if (ls.Resc < 0 || ls.Resm < 0) {
    ls.Resc = 0;
    ls.Resm = 0;
    ls.RI.dB = 0;
    ls.RI.fNorm = 0.0;
    if (ops.is_master()) {
      ops.release(mk);
      throw std::runtime_error("FSILS: unexpected behavior in FSILS (likely due to the ill-conditioned LHS matrix)");
    }
  }
  ls.RI.fNorm = std::sqrt(ls.RI.fNorm);

  ops.join_mc(dof, Rmi, Rci, Ri);
  ops.release(mk);
}

// ---------------------------------------------------------------------------------------------
// solve: fsils_solve (liner_solver/solve.cpp:50-193).  R (dof x nNo) and Val are in SOLVER ordering
// already (the permutation by lhs.map is done by the caller on upload / download).
// incL/res may be null.  On return R holds the solution and Val the scaled matrix.
// ---------------------------------------------------------------------------------------------
template <class Ops>
void solve(Ops& ops, Ls& ls, int dof, int prec, double* R, double* Val, const int* incL, const double* res)
{
  const int nFaces = ops.n_faces();
  if (nFaces != 0) {
    bool any_neu = false;
    for (int f = 0; f < nFaces; f++) {
      ops.face_set_inc(f, incL ? (incL[f] != 0) : true);
      if (ops.face_bgrp(f) == B200_BC_NEU) any_neu = true;
    }
    if (res == nullptr && any_neu) throw std::runtime_error("[fsils_solve] res is required for Neu surfaces");
    for (int f = 0; f < nFaces; f++) {
      bool coupled = false;
      double r = 0.0;
      if (ops.face_inc(f) && ops.face_bgrp(f) == B200_BC_NEU && res[f] != 0.0) { coupled = true; r = res[f]; }
      ops.face_set_coupled(f, coupled, r);
    }
  }

  const size_t n = size_t(dof)*ops.nNo();
  auto mk = ops.mark();
  double* Wr = ops.vec(n);
  double* Wc = ops.vec(n);

  double t0 = wall_s();
  if (prec == B200_PREC_FSILS) {
    ops.precond_diag(dof, Val, R, Wc);
  } else if (prec == B200_PREC_RCS) {
    ops.precond_rcs(dof, Val, R, Wr, Wc);
  } else {
    ops.fill(n, 1.0, Wc);      // the reference leaves Wc = 0 here (solution zeroed); we refuse instead
    throw std::runtime_error("This linear solver and preconditioner combination is not supported.");
  }
  ops.phase_mark(1, t0);

  t0 = wall_s();
  switch (ls.LS_type) {
    case B200_LS_NS:    ns_solver(ops, ls, dof, Val, R); break;
    case B200_LS_GMRES: gmres_v(ops, ls.RI, dof, Val, R); break;
    case B200_LS_CG:    cgrad_v(ops, ls.RI, dof, Val, R); break;
    case B200_LS_BICGS: bicgs_v(ops, ls.RI, dof, Val, R); break;
    default: throw std::runtime_error("FSILS: LS_type not defined");
  }
  ops.mul_inplace(n, Wc, R);    // R = Wc (.) R
  ops.phase_mark(2, t0);
  ops.release(mk);
}

} // namespace svb200
