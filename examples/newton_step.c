/* One Newton iteration of the Navier-Stokes hot path through the C ABI (include/svb200.h), in plain C99: what a binding from
 * any language does.  A two-tet channel is far too small to be interesting - the point is the call sequence:
 *
 *   b200_create -> b200_pattern_* (lhsa) -> b200_lhs_create (fsils_lhs_create's result; one rank: identity map) -> b200_face_set
 *   (fsils_bc_create) -> b200_mesh_set -> [ b200_zero (ls_alloc) -> b200_state_set -> b200_assemble_fluid (construct_fluid)
 *   -> b200_solve (fsils_solve) ] per Newton iteration -> b200_destroy
 *
 *   gcc -std=c99 -Iinclude examples/newton_step.c -Lsvfsiplus_b200 -lsvb200 -Wl,-rpath,$PWD/svfsiplus_b200 -lm -o /tmp/newton_step
 *
 * Without a CUDA device b200_create fails (there is no CPU fallback) and the program says so and exits with status 3. */
#include "svb200.h"

#include <stdio.h>
#include <stdlib.h>

#define CK(call) do { if ((call) != 0) { fprintf(stderr, "%s failed: %s\n", #call, b200_last_error(h)); b200_destroy(h); return 1; } } while (0)

int main(void)
{
  b200_handle* h = NULL;
  if (b200_device_count() < 1) {
    fprintf(stderr, "no CUDA device (b200_device_count() = %d): the hot path has no CPU fallback\n", b200_device_count());
    return 3;
  }
  if (b200_create(&h, 0) != 0) {
    fprintf(stderr, "b200_create failed: %s\n", b200_last_error(NULL));
    return 3;
  }

  /* a unit cube cut into 6 tets around the diagonal 0-7 (Kuhn), 8 nodes */
  enum { nNo = 8, nEl = 6, tDof = 4 };
  const double x[nNo*3] = {0,0,0, 1,0,0, 0,1,0, 1,1,0, 0,0,1, 1,0,1, 0,1,1, 1,1,1};
  const int ien[nEl*4] = {0,1,3,7, 0,3,2,7, 0,2,6,7, 0,6,4,7, 0,4,5,7, 0,5,1,7};

  /* sparsity pattern on the device = lhsa */
  int nnz = 0;
  CK(b200_pattern_begin(h, nNo));
  CK(b200_pattern_add_mesh(h, 4, nEl, ien));
  CK(b200_pattern_finish(h, &nnz));
  int* rowPtr = malloc(sizeof(int)*(nNo + 1));
  int* colPtr = malloc(sizeof(int)*nnz);
  CK(b200_pattern_get(h, rowPtr, colPtr));

  /* one rank: no map, every row counted, no overlap lists; one Dirichlet face (the bottom nodes, all three velocity components) */
  CK(b200_lhs_create(h, nNo, nNo, nNo, nnz, rowPtr, colPtr, NULL, 0, NULL, NULL, NULL, 1));
  const int wall[4] = {0, 1, 2, 3};
  const double wall_val[4*3] = {0};            /* val = 0: the component is constrained (fsils_bc_create's Dirichlet convention) */
  CK(b200_face_set(h, 0, 4, 3, B200_BC_DIR, wall, wall_val, 0));
  CK(b200_mesh_set(h, 4, nEl, ien, x, -1.0));

  /* state at the generalised-alpha points: a shear flow, no acceleration, no body force */
  double Ag[nNo*tDof] = {0}, Yg[nNo*tDof] = {0}, Bf[nNo*3] = {0};
  for (int a = 0; a < nNo; a++) Yg[a*tDof + 0] = x[a*3 + 2];

  b200_fluid_props p;
  p.dt = 0.005; p.am = 5.0/6.0; p.af = 2.0/3.0; p.gam = 2.0/3.0;      /* rho_inf = 0.5 */
  p.tDof = tDof; p.mvMsh = 0; p.rho = 1.06; p.f[0] = p.f[1] = p.f[2] = 0.0; p.Kinv = 0.0;
  p.viscType = 0; p.mu_i = 0.04; p.mu_o = 0.0; p.lam = 0.0; p.a = 0.0; p.n = 0.0;

  CK(b200_zero(h, 4));
  CK(b200_state_set(h, tDof, Ag, Yg, Bf));
  CK(b200_assemble_fluid(h, &p));

  const b200_tol RI = {1e-6, 1e-14, 10, 50};
  const int incL[1] = {1};
  const double res[1] = {0.0};
  double X[nNo*4];
  b200_ls_out out;
  CK(b200_solve(h, B200_LS_GMRES, B200_PREC_FSILS, &RI, NULL, NULL, incL, res, X, &out));
  printf("GMRES: suc %d, %d iterations, |R| %.3e -> %.3e\n", out.RI.suc, out.RI.itr, out.RI.iNorm, out.RI.fNorm);
  for (int a = 0; a < nNo; a++) printf("  node %d: du = (% .3e % .3e % .3e), dp = % .3e\n", a, X[a*4], X[a*4 + 1], X[a*4 + 2], X[a*4 + 3]);

  free(rowPtr); free(colPtr);
  b200_destroy(h);
  return 0;
}
