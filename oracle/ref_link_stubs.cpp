// TEST INFRASTRUCTURE ONLY (oracle).  Link-time stand-ins for symbols the reference's hot-path
// sources reference but which live outside the path (cardiac electrophysiology) or in an
// external library that is not installed here (LAPACK).  The LAPACK routines are small textbook
// partial-pivoting LU implementations with the Fortran calling convention; the reference calls
// dgesv_ only for second derivatives of shape functions (Code/Source/solver/nn.cpp:845,916), whose
// right-hand side is identically zero for affine TET4 elements, and dgetrf_/dgetri_ only from
// mat_fun::mat_inv_lp (Code/Source/solver/mat_fun.cpp:434,443).
#include "ComMod.h"
#include "CepMod.h"
#include "Simulation.h"

#include <cmath>
#include <stdexcept>
#include <vector>

#ifndef ORACLE_FULL_REFERENCE      // the full-reference build (libsvfull.so) links the real cep.cpp / cep_ion.cpp
namespace cep {
void b_cep(ComMod&, const int, const double, const Vector<double>&, const double, Array<double>&)
{ throw std::runtime_error("[oracle] cep::b_cep is outside the hot path"); }
void construct_cep(ComMod&, CepMod&, const mshType&, const Array<double>&, const Array<double>&, const Array<double>&)
{ throw std::runtime_error("[oracle] cep::construct_cep is outside the hot path"); }
}

namespace cep_ion {
void cep_integ(Simulation*, const int, const int, const Array<double>&)
{ throw std::runtime_error("[oracle] cep_ion::cep_integ is outside the hot path"); }
}

#endif

extern "C" {

// A (n x n, column-major, lda) = P L U in place; ipiv 1-based.
void dgetrf_(int* m, int* n, double* A, int* lda, int* ipiv, int* info)
{
  const int N = *n, M = *m, ld = *lda;
  *info = 0;
  for (int k = 0; k < std::min(M, N); k++) {
    int p = k;
    double mx = std::fabs(A[k + k*ld]);
    for (int i = k+1; i < M; i++) if (std::fabs(A[i + k*ld]) > mx) { mx = std::fabs(A[i + k*ld]); p = i; }
    ipiv[k] = p + 1;
    if (mx == 0.0) { if (*info == 0) *info = k + 1; continue; }
    if (p != k) for (int j = 0; j < N; j++) std::swap(A[k + j*ld], A[p + j*ld]);
    for (int i = k+1; i < M; i++) A[i + k*ld] /= A[k + k*ld];
    for (int j = k+1; j < N; j++) {
      double akj = A[k + j*ld];
      for (int i = k+1; i < M; i++) A[i + j*ld] -= A[i + k*ld]*akj;
    }
  }
}

void dgetrs_n(int N, int nrhs, const double* A, int ld, const int* ipiv, double* B, int ldb)
{
  for (int r = 0; r < nrhs; r++) {
    double* b = B + size_t(r)*ldb;
    for (int k = 0; k < N; k++) { int p = ipiv[k]-1; if (p != k) std::swap(b[k], b[p]); }
    for (int k = 0; k < N; k++) for (int i = k+1; i < N; i++) b[i] -= A[i + k*ld]*b[k];
    for (int k = N-1; k >= 0; k--) { b[k] /= A[k + k*ld]; for (int i = 0; i < k; i++) b[i] -= A[i + k*ld]*b[k]; }
  }
}

void dgesv_(int* n, int* nrhs, double* A, int* lda, int* ipiv, double* B, int* ldb, int* info)
{
  dgetrf_(n, n, A, lda, ipiv, info);
  if (*info != 0) return;
  dgetrs_n(*n, *nrhs, A, *lda, ipiv, B, *ldb);
}

void dgetri_(int* n, double* A, int* lda, int* ipiv, double* work, int* lwork, int* info)
{
  const int N = *n, ld = *lda;
  *info = 0;
  if (*lwork == -1) { work[0] = double(N)*N; return; }
  std::vector<double> inv(size_t(N)*N, 0.0);
  for (int i = 0; i < N; i++) inv[i + size_t(i)*N] = 1.0;
  dgetrs_n(N, N, A, ld, ipiv, inv.data(), N);
  for (int j = 0; j < N; j++) for (int i = 0; i < N; i++) A[i + j*ld] = inv[i + size_t(j)*N];
}

}
