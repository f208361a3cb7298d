// assembly.cuh — K10: whole-mesh Navier-Stokes (VMS, P1-P1) element assembly for linear tetrahedra.
//
// Replaces construct_fluid + fluid_3d_m + fluid_3d_c + gnn + do_assem
// (Code/Source/solver/fluid.cpp:464-708, 1697-2139, 1389-1689; nn.cpp:455-541; lhsa.cpp:97-142).
//
// Design: one thread per element.  The reference walks the Gauss points twice (momentum, then
// continuity) and recomputes the whole kinematic preamble in both; here the preamble is computed
// once per Gauss point and shared.  For an affine TET4 the shape-function gradients are constant in
// the element and the second derivatives vanish identically (Nxx = 0 in nn_elem_gnnxx.h, so the
// dgesv_ right-hand side of gn_nxx, nn.cpp:845, is 0): uxx = 0, d2u2 = 0, mu_x = 0, rS = 0 and
// updu(i,j,a) = delta_ij * T1(a); strain rate, viscosity and the gradient terms are element
// constants and only N_a(g), u, tauM, tauC, tauB, u' vary with the Gauss point.  Those exact zeros
// are dropped, every other term is evaluated in the reference's order.
//
// Scatter is deterministic AND in the reference's order: no colours, no atomics, no read-modify-write.
// Every element writes its 16 tangent blocks and 4 residual rows into a staging buffer at slots that
// were sorted once (mesh_set) by destination and, inside a destination, by ascending element number;
// a second streaming kernel then sums each destination's contiguous run of contributions left to
// right, i.e. in exactly the order in which do_assem (lhsa.cpp:97-142) adds them while construct_fluid
// walks e = 0..nEl-1, and writes every Val block / R row once.  Traffic: 2 KB written + 2 KB read per
// element (HBM streaming, full sectors) instead of 16 scattered 128-byte read-modify-writes.
#pragma once

#include "kernels.cuh"
#include "fluid_elem.hpp"      // FluidConsts, is_zero_d, viscosity (host/device shared)

namespace svb200 {

// one thread per element, elements in mesh order
// elist != nullptr: the kernel covers the nEl elements elist[0..nEl) (the fluid domain of an FSI equation);
// Dmesh != nullptr: the element lives on the ALE-displaced configuration x + Dg(4:6) (fsi.cpp:157-163).
__global__ void __launch_bounds__(128)
k_assemble_fluid_tet4(int nEl, const int* __restrict__ elist, const double* __restrict__ Dmesh, FluidConsts c,
                      const int* __restrict__ ien,      // 4 x nEl, assembly node ids
                      const int* __restrict__ rslot,    // 4 x nEl staging slot (32-byte rows) of lR(:,a)
                      const int* __restrict__ kslot,    // 16 x nEl staging slot (128-byte blocks) of lK(:,a,b)
                      const double* __restrict__ x,     // 3 x nNo
                      const double* __restrict__ Ag, const double* __restrict__ Yg, const double* __restrict__ Bf,
                      double* __restrict__ stageR, double* __restrict__ stageK, int* __restrict__ err_flag)
{
  const int ei = blockIdx.x*blockDim.x + threadIdx.x;
  if (ei >= nEl) return;
  const int e = elist ? elist[ei] : ei;

  int nd[4];
  {
    const int4 v = *reinterpret_cast<const int4*>(ien + size_t(e)*4);
    nd[0] = v.x; nd[1] = v.y; nd[2] = v.z; nd[3] = v.w;
  }
  const int tD = c.tDof;

  // ---- gather (fluid.cpp:546-558) ---------------------------------------------------------------
  double xl[4][3], al[4][3], yl[4][4], bl[4][3], ym[4][3];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const size_t A = size_t(nd[a]);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      xl[a][i] = x[A*3 + i];
      if (Dmesh) xl[a][i] = xl[a][i] + Dmesh[A*tD + 4 + i];
      bl[a][i] = Bf[A*3 + i];
      al[a][i] = Ag[A*tD + i];
      yl[a][i] = Yg[A*tD + i];
      ym[a][i] = c.mvMsh ? Yg[A*tD + 4 + i] : 0.0;
    }
    yl[a][3] = Yg[A*tD + 3];
  }

  // ---- nn::gnn for TET4 (nn.cpp:505-540): Nxi = [e1 e2 e3 -1] ---------------------------------------
  double xXi[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    // xXi(i,k) accumulated over a = 0..3 exactly like the reference: x(i,k)*1 + ... - x(i,3)
    xXi[i][0] = xl[0][i] - xl[3][i];
    xXi[i][1] = xl[1][i] - xl[3][i];
    xXi[i][2] = xl[2][i] - xl[3][i];
  }
  const double Jac = xXi[0][0]*xXi[1][1]*xXi[2][2] + xXi[0][1]*xXi[1][2]*xXi[2][0] + xXi[0][2]*xXi[1][0]*xXi[2][1]
                   - xXi[0][0]*xXi[1][2]*xXi[2][1] - xXi[0][1]*xXi[1][0]*xXi[2][2] - xXi[0][2]*xXi[1][1]*xXi[2][0];
  if (is_zero_d(Jac)) { atomicExch(err_flag, e + 1); return; }      // fluid.cpp:612-614 throws
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

  double ks[3][3];
  ks[0][0] = xiX[0][0]*xiX[0][0] + xiX[1][0]*xiX[1][0] + xiX[2][0]*xiX[2][0];
  ks[0][1] = xiX[0][1]*xiX[0][0] + xiX[1][1]*xiX[1][0] + xiX[2][1]*xiX[2][0];
  ks[0][2] = xiX[0][2]*xiX[0][0] + xiX[1][2]*xiX[1][0] + xiX[2][2]*xiX[2][0];
  ks[1][1] = xiX[0][1]*xiX[0][1] + xiX[1][1]*xiX[1][1] + xiX[2][1]*xiX[2][1];
  ks[1][2] = xiX[0][1]*xiX[0][2] + xiX[1][1]*xiX[1][2] + xiX[2][1]*xiX[2][2];
  ks[2][2] = xiX[0][2]*xiX[0][2] + xiX[1][2]*xiX[1][2] + xiX[2][2]*xiX[2][2];
  ks[1][0] = ks[0][1]; ks[2][0] = ks[0][2]; ks[2][1] = ks[1][2];

  // Nx(i,a) = sum_k Nxi(k,a) xiX(k,i)
  double Nx[4][3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    Nx[0][i] = xiX[0][i];
    Nx[1][i] = xiX[1][i];
    Nx[2][i] = xiX[2][i];
    Nx[3][i] = -xiX[0][i] - xiX[1][i] - xiX[2][i];
  }

  // ---- element constants ------------------------------------------------------------------------
  const double rho = c.rho;
  const double T1c = c.af*c.gam*c.dt;
  const double amd = c.am/T1c;

  double ux[3][3];          // ux[i][j] = d u_j / d x_i
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < 4; a++) s += Nx[a][i]*yl[a][j];
      ux[i][j] = s;
    }
  const double divU = ux[0][0] + ux[1][1] + ux[2][2];
  double px[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < 4; a++) s += Nx[a][i]*yl[a][3];
    px[i] = s;
  }
  double es[3][3];
  es[0][0] = ux[0][0] + ux[0][0];
  es[1][1] = ux[1][1] + ux[1][1];
  es[2][2] = ux[2][2] + ux[2][2];
  es[1][0] = ux[1][0] + ux[0][1];
  es[2][1] = ux[2][1] + ux[1][2];
  es[0][2] = ux[0][2] + ux[2][0];
  es[0][1] = es[1][0]; es[1][2] = es[2][1]; es[2][0] = es[0][2];

  double esNx[3][4];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    esNx[0][a] = es[0][0]*Nx[a][0] + es[1][0]*Nx[a][1] + es[2][0]*Nx[a][2];
    esNx[1][a] = es[0][1]*Nx[a][0] + es[1][1]*Nx[a][1] + es[2][1]*Nx[a][2];
    esNx[2][a] = es[0][2]*Nx[a][0] + es[1][2]*Nx[a][1] + es[2][2]*Nx[a][2];
  }
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
  double kS = ks[0][0]*ks[0][0] + ks[1][0]*ks[1][0] + ks[2][0]*ks[2][0]
            + ks[0][1]*ks[0][1] + ks[1][1]*ks[1][1] + ks[2][1]*ks[2][1]
            + ks[0][2]*ks[0][2] + ks[1][2]*ks[1][2] + ks[2][2]*ks[2][2];
  {
    const double t = mu/rho;
    kS = 36.0*kS*(t*t);
  }
  const double trks = ks[0][0] + ks[1][1] + ks[2][2];

  // ---- Gauss loop 1: residual + per-point scalars kept for the tangent ------------------------------
  double tauM_g[4], tauC_g[4], tauB_g[4];
  double uNx[4][4], upNx[4][4];        // [g][a]
  double lR[4][4];                     // [a][i]
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int i = 0; i < 4; i++) lR[a][i] = 0.0;

#pragma unroll
  for (int g = 0; g < 4; g++) {
    const double w = c.w[g]*Jac;
    const double wr = w*rho;
    double ud[3] = {-c.f[0], -c.f[1], -c.f[2]};
    double u[3] = {0.0, 0.0, 0.0};
    double p = 0.0;
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double Na = c.N[g][a];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        ud[i] = ud[i] + Na*(al[a][i] - bl[a][i]);
        u[i] = u[i] + Na*yl[a][i];
      }
      p = p + Na*yl[a][3];
    }
    if (c.mvMsh) {
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int i = 0; i < 3; i++) u[i] = u[i] - c.N[g][a]*ym[a][i];
    }
    const double kU = u[0]*u[0]*ks[0][0] + u[1]*u[0]*ks[1][0] + u[2]*u[0]*ks[2][0]
                    + u[0]*u[1]*ks[0][1] + u[1]*u[1]*ks[1][1] + u[2]*u[1]*ks[2][1]
                    + u[0]*u[2]*ks[0][2] + u[1]*u[2]*ks[1][2] + u[2]*u[2]*ks[2][2];
    const double tauM = 1.0/(rho*sqrt(kT + kU + kS));

    double rV[3];
#pragma unroll
    for (int j = 0; j < 3; j++) rV[j] = ud[j] + u[0]*ux[0][j] + u[1]*ux[1][j] + u[2]*ux[2][j];
    double up[3];
#pragma unroll
    for (int j = 0; j < 3; j++) up[j] = -tauM*(rho*rV[j] + px[j] - 0.0 + muK*u[j]);

    const double tauC = 1.0/(tauM*trks);
    double tauB = up[0]*up[0]*ks[0][0] + up[1]*up[0]*ks[1][0] + up[2]*up[0]*ks[2][0]
                + up[0]*up[1]*ks[0][1] + up[1]*up[1]*ks[1][1] + up[2]*up[1]*ks[2][1]
                + up[0]*up[2]*ks[0][2] + up[1]*up[2]*ks[1][2] + up[2]*up[2]*ks[2][2];
    if (is_zero_d(tauB)) tauB = 2.220446049250313e-16;
    tauB = rho/sqrt(tauB);
    double ua[3] = {u[0] + up[0], u[1] + up[1], u[2] + up[2]};
    const double pa = p - tauC*divU;

#pragma unroll
    for (int j = 0; j < 3; j++) rV[j] = tauB*(up[0]*ux[0][j] + up[1]*ux[1][j] + up[2]*ux[2][j]);
    double rM[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        double t = mu*es[i][j] - rho*up[j]*ua[i] + rV[j]*up[i];
        if (i == j) t = t - pa;
        rM[i][j] = t;
      }
#pragma unroll
    for (int j = 0; j < 3; j++) rV[j] = ud[j] + ua[0]*ux[0][j] + ua[1]*ux[1][j] + ua[2]*ux[2][j];

#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double Na = c.N[g][a];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        lR[a][j] = lR[a][j] + wr*Na*rV[j] + w*(Nx[a][0]*rM[0][j] + Nx[a][1]*rM[1][j] + Nx[a][2]*rM[2][j]);
      }
      const double un = u[0]*Nx[a][0] + u[1]*Nx[a][1] + u[2]*Nx[a][2];
      const double upn = up[0]*Nx[a][0] + up[1]*Nx[a][1] + up[2]*Nx[a][2];
      uNx[g][a] = un;
      upNx[g][a] = upn;
      // continuity residual (fluid_3d_c, fluid.cpp:1655-1658)
      lR[a][3] = lR[a][3] + w*(Na*divU - upn);
    }
    // Brinkman residual term (fluid.cpp:2134-2138) is added after the tangent in the reference; the
    // accumulation order per Gauss point is kept: it comes after this point's momentum residual.
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int j = 0; j < 3; j++) lR[a][j] = lR[a][j] + muK*w*c.N[g][a]*(u[j] + up[j]);

    tauM_g[g] = tauM; tauC_g[g] = tauC; tauB_g[g] = tauB;
  }

  // ---- residual rows to their staging slots (summed in element order by k_sum_segments) -------------
  {
    const int4 rd = *reinterpret_cast<const int4*>(rslot + size_t(e)*4);
    const int rr[4] = {rd.x, rd.y, rd.z, rd.w};
#pragma unroll
    for (int a = 0; a < 4; a++) {
      d4 v; v.x = lR[a][0]; v.y = lR[a][1]; v.z = lR[a][2]; v.w = lR[a][3];
      st256_stream(stageR + size_t(rr[a])*4, v);
    }
  }

  // ---- tangent: (a,b) outer, Gauss points inner, 4x4 block kept in registers -------------------------
  // `a` stays a run-time loop (code size); everything indexed by it is first selected into scalars
  // so that no array is indexed dynamically (which would push it to local memory).
#pragma unroll 1
  for (int a = 0; a < 4; a++) {
    const int4 ed = *reinterpret_cast<const int4*>(kslot + size_t(e)*16 + a*4);
    const int pos[4] = {ed.x, ed.y, ed.z, ed.w};
    double Nxa[3], esNxa[3], uNxa[4], upNxa[4], Na_g[4];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      Nxa[i] = (a == 0) ? Nx[0][i] : (a == 1) ? Nx[1][i] : (a == 2) ? Nx[2][i] : Nx[3][i];
      esNxa[i] = (a == 0) ? esNx[i][0] : (a == 1) ? esNx[i][1] : (a == 2) ? esNx[i][2] : esNx[i][3];
    }
#pragma unroll
    for (int g = 0; g < 4; g++) {
      uNxa[g] = (a == 0) ? uNx[g][0] : (a == 1) ? uNx[g][1] : (a == 2) ? uNx[g][2] : uNx[g][3];
      upNxa[g] = (a == 0) ? upNx[g][0] : (a == 1) ? upNx[g][1] : (a == 2) ? upNx[g][2] : upNx[g][3];
      Na_g[g] = c.N[g][a];
    }
#pragma unroll
    for (int b = 0; b < 4; b++) {
      double kb[4][4];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) kb[i][j] = 0.0;

      const double NxNx = Nxa[0]*Nx[b][0] + Nxa[1]*Nx[b][1] + Nxa[2]*Nx[b][2];
#pragma unroll
      for (int g = 0; g < 4; g++) {
        const double w = c.w[g]*Jac;
        const double wl = w*T1c;
        const double Na = Na_g[g], Nb = c.N[g][b];
        const double tauM = tauM_g[g], tauC = tauC_g[g], tauB = tauB_g[g];
        const double uaNx_a = uNxa[g] + upNxa[g];
        // updu(i,i,b) = T1u(b) for the affine tet (fluid.cpp:2038-2050 with Nwxx = mu_x = d2u2 = 0)
        const double T1u_b = -rho*uNx[g][b] + mu*(0.0) - muK*Nb;
        const double rtu = rho*tauM*uaNx_a;

        // momentum-velocity block (fluid.cpp:2058-2113)
        const double T1 = mu*NxNx + rho*amd*Nb*(Na + rho*tauM*uaNx_a) + rho*Na*(uNx[g][b] + upNx[g][b]) + tauB*upNxa[g]*upNx[g][b];
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
          for (int j = 0; j < 3; j++) {
            if (i == j) {
              const double T2 = (mu + tauC)*(Nxa[i]*Nx[b][i]) + esNxa[i]*mu_g*esNx[i][b] - rtu*T1u_b;
              kb[i][i] = kb[i][i] + wl*(T2 + T1);
              kb[i][i] = kb[i][i] + muK*wl*Nb*Na;
            } else {
              const double T2 = mu*(Nxa[j]*Nx[b][i]) + tauC*(Nxa[i]*Nx[b][j]) + esNxa[i]*mu_g*esNx[j][b];
              kb[i][j] = kb[i][j] + wl*T2;
            }
          }
        }
        // momentum-pressure block (fluid.cpp:2117-2130)
#pragma unroll
        for (int i = 0; i < 3; i++) kb[i][3] = kb[i][3] - wl*(Nxa[i]*Nb - Nx[b][i]*rtu);
        // continuity-velocity block (fluid_3d_c, fluid.cpp:1662-1678): updu diagonal => single term
        {
          const double T1cc = rho*amd*Nb;
#pragma unroll
          for (int j = 0; j < 3; j++) {
            const double T2 = Nxa[j]*(T1u_b - T1cc);
            kb[3][j] = kb[3][j] + wl*(Na*Nx[b][j] - tauM*T2);
          }
        }
        // continuity-pressure block (fluid.cpp:1680-1688)
        kb[3][3] = kb[3][3] + wl*tauM*NxNx;
      }

      // ---- lK(:,a,b) to its staging slot, whole 128-byte block ----------------------------------------
      double* v = stageK + size_t(pos[b])*16;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        d4 t; t.x = kb[i][0]; t.y = kb[i][1]; t.z = kb[i][2]; t.w = kb[i][3];
        st256_stream(v + 4*i, t);
      }
    }
  }
}

// do_assem (lhsa.cpp:97-142) as an ordered segmented sum: destination d (a Val block when W = 4 lanes
// of 32 bytes, an R row when W = 1) owns the contiguous staging run [seg[d], seg[d+1]), ordered by
// ascending element; out(:,d) (+)= sum of the run, left to right.  ASSIGN: out is known to be zero
// (ls_alloc just ran), so it is neither read nor was it memset.  W lanes per destination, one 256-bit
// row each: a warp streams 32/W consecutive runs, i.e. one contiguous piece of the staging buffer.
template <int W, bool ASSIGN>
__global__ void __launch_bounds__(256)
k_sum_segments(size_t nDest, const int* __restrict__ seg, const double* __restrict__ stage, double* __restrict__ out)
{
  const size_t t = size_t(blockIdx.x)*blockDim.x + threadIdx.x;
  const size_t d = t / W;
  const int l = int(t % W);
  if (d >= nDest) return;
  const int s = __ldg(seg + d), e = __ldg(seg + d + 1);
  d4 acc;
  if (ASSIGN) { acc.x = acc.y = acc.z = acc.w = 0.0; }
  else acc = ld256(out + (d*W + l)*4);
  int q = s;
  for (; q + 4 <= e; q += 4) {
    const d4 v0 = ld256_stream(stage + (size_t(q)*W + l)*4);
    const d4 v1 = ld256_stream(stage + (size_t(q+1)*W + l)*4);
    const d4 v2 = ld256_stream(stage + (size_t(q+2)*W + l)*4);
    const d4 v3 = ld256_stream(stage + (size_t(q+3)*W + l)*4);
    acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
    acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
    acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
    acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
  }
  for (; q < e; q++) {
    const d4 v = ld256_stream(stage + (size_t(q)*W + l)*4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  st256(out + (d*W + l)*4, acc);
}

// ---- one-time slot construction (mesh_set) --------------------------------------------------------
// key[i] = destination of item i (i = e*16 + a*4 + b for blocks, e*4 + a for rows); items of one
// destination are ranked by ascending i, so that the run of a destination is in element order.
__global__ void k_slot_count(size_t nItems, const int* __restrict__ key, int* __restrict__ cnt, int* __restrict__ bad)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < nItems; i += nth) {
    const int k = key[i];
    if (k < 0) { atomicExch(bad, 1); continue; }
    atomicAdd(cnt + k, 1);
  }
}
__global__ void k_slot_fill(size_t nItems, const int* __restrict__ key, const int* __restrict__ seg, int* __restrict__ cursor,
                            int* __restrict__ items)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < nItems; i += nth) {
    const int k = key[i];
    if (k < 0) continue;
    items[seg[k] + atomicAdd(cursor + k, 1)] = int(i);
  }
}
// one thread per destination: insertion-sort its (short) item list, then publish the slots
__global__ void k_slot_rank(size_t nDest, const int* __restrict__ seg, int* __restrict__ items, int* __restrict__ slot)
{
  const size_t nth = size_t(gridDim.x)*blockDim.x;
  for (size_t d = size_t(blockIdx.x)*blockDim.x + threadIdx.x; d < nDest; d += nth) {
    const int s = seg[d], e = seg[d+1];
    for (int i = s + 1; i < e; i++) {
      const int v = items[i];
      int j = i - 1;
      while (j >= s && items[j] > v) { items[j+1] = items[j]; j--; }
      items[j+1] = v;
    }
    for (int i = s; i < e; i++) slot[items[i]] = i;
  }
}

// positions of the 16 (a,b) pairs of each element in the solver-layout Val, and solver rows.
// rowPtrA/colA: assembly CSR (sorted columns); map: assembly -> solver id; rowPtrS: solver CSR.
__global__ void k_elem_dest(int nEl, int eNoN, const int* __restrict__ ien, const int* __restrict__ rowPtrA,
                            const int* __restrict__ colA, const int* __restrict__ map, const int* __restrict__ rowPtrS,
                            int* __restrict__ rdest, int* __restrict__ edest)
{
  const int tot = nEl*eNoN;
  for (int t = blockIdx.x*blockDim.x + threadIdx.x; t < tot; t += gridDim.x*blockDim.x) {
    const int e = t / eNoN, a = t % eNoN;
    const int A = ien[size_t(e)*eNoN + a];
    const int s = rowPtrA[A], len = rowPtrA[A+1] - s;
    const int rowS = map[A];
    rdest[t] = rowS;
    for (int b = 0; b < eNoN; b++) {
      const int B = ien[size_t(e)*eNoN + b];
      int lo = 0, hi = len - 1, pos = -1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int cv = colA[s + mid];
        if (cv == B) { pos = mid; break; }
        if (cv < B) lo = mid + 1; else hi = mid - 1;
      }
      edest[(size_t(e)*eNoN + a)*eNoN + b] = (pos < 0) ? -1 : rowPtrS[rowS] + pos;
    }
  }
}

} // namespace svb200
