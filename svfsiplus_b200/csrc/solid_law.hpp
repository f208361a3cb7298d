// solid_law.hpp — constitutive laws of the displacement-based and mixed solid elements, host/device shared like
// fluid_elem.hpp: 2nd Piola-Kirchhoff stress S and the Voigt elasticity matrix Dm from the deformation gradient.
// Replaces mat_models_carray::get_pk2cc<3> (Code/Source/solver/mat_models_carray.h:182-1380; neo-Hookean :370-434,
// Mooney-Rivlin :438-540, Holzapfel-Gasser-Ogden :544-688, Guccione :692-903, St.Venant-Kirchhoff :302-322, modified StVK :326-356, Holzapfel-Ogden :905-1135, HO-ma :1137-1353) with
// get_svol_p (mat_models.cpp:1626-1645) and the fibre reinforcement stress (mat_models_carray.h:222-225).
// tests/hostlogic/fluid_elem_host.cpp instantiates the same source on the CPU (test tree only) and
// tests/test_solid_laws.py compares it with the compiled reference's get_pk2cc on random deformation gradients.
#pragma once

#include "fluid_elem.hpp"      // SVB_HD, is_zero_d

namespace svb200 {

struct SolidConsts {
  double dt, am, af, gam, beta;
  double rho, dmp, f[3];
  int iso, vol;                // iso: 0 nHook, 1 StVK, 2 mStVK, 3 Holzapfel-Ogden, 4 Mooney-Rivlin, 5 HGO, 6 Guccione, 7 Holzapfel-Ogden modified anisotropy (HO-ma); vol: 0 none, 1 Quad, 2 ST91, 3 M94
  double C10, C01, Kpen;
  double ho_a, ho_b, ho_aff, ho_bff, ho_ass, ho_bss, ho_afs, ho_bfs, ho_khs;   // stModelType a..bfs, khs
  double Tfa, Tsa;             // fibre / sheet reinforcement stress (get_fib_stress, mat_models_carray.h:222-225)
  double kap;                  // HGO fibre dispersion (stM.kap)
  double elM, nu;              // lElas / mesh
  int tDof, s;                 // row offset of this equation's unknowns in Ag/Yg/Dg (eq.s)
  int kind;                    // 0 struct, 1 lElas, 2 mesh
  int viscType;                // solid viscosity: 0 none, 1 Newtonian, 2 pseudo-potential (dmn.solid_visc)
  double visc_mu;
};

// index of (I,J), I <= J, in the packed upper triangle of the 6x6 Voigt matrix
SVB_HD int dm_idx(int I, int J) { return I*6 - (I*(I-1))/2 + (J - I); }

// x^3 of the smoothed-Heaviside derivatives: the reference calls pow(x, 3) (correctly rounded in glibc), and the expression
// -x + 3x^2 - 2x^3 cancels to ~1 - x for compressed fibres (x -> 1), so the host instantiation calls pow as well to stay
// within rounding of the reference there; the device multiplies (CUDA's pow is not correctly rounded either way).
SVB_HD double cube_d(double x)
{
#ifdef __CUDA_ARCH__
  return x*x*x;
#else
  return pow(x, 3.0);
#endif
}

struct HoParams { double a, b, aff, bff, ass, bss, afs, bfs, khs, Tfa, Tsa; };

// Isochoric part of the Holzapfel-Ogden law (mat_models_carray.h:905-1060, mat_models.cpp:866-935): isochoric
// stress S = J2d Sb - r1 Ci, r1, and the projected rank-one factors of the isochoric tangent
//   PP : (sum_k g_k H_k (x) H_k) : PP^T = sum_k g_k Hd_k (x) Hd_k,   Hd_k = H_k - (1/3)(C : H_k) Ci
// (the reference forms CCb and contracts it with PP = Ids - (1/3) Ci (x) C from both sides; for symmetric H_k
// this is the same tensor up to rounding).  H_0 = I, H_1 = sym(f (x) s), H_2 = f (x) f, H_3 = s (x) s.
SVB_HD_NOINL void ho_isochoric(const HoParams& h, const double C[3][3], const double Ci[3][3], double J2d, double Inv1,
                             const double* fl, double S[3][3], double& r1, double gk[4], double H[4][3][3])
{
  const double nd = 3.0;
  const double J4d = J2d*J2d;
  const double f0[3] = {fl[0], fl[1], fl[2]}, s0[3] = {fl[3], fl[4], fl[5]};
  double Cf[3], Cs[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    Cf[i] = C[i][0]*f0[0] + C[i][1]*f0[1] + C[i][2]*f0[2];
    Cs[i] = C[i][0]*s0[0] + C[i][1]*s0[1] + C[i][2]*s0[2];
  }
  const double Inv4 = J2d*(f0[0]*Cf[0] + f0[1]*Cf[1] + f0[2]*Cf[2]);
  const double Inv6 = J2d*(s0[0]*Cs[0] + s0[1]*Cs[1] + s0[2]*Cs[2]);
  const double Inv8 = J2d*(f0[0]*Cs[0] + f0[1]*Cs[1] + f0[2]*Cs[2]);
  const double Eff = Inv4 - 1.0, Ess = Inv6 - 1.0, Efs = Inv8;
  const double k = h.khs;
  const double of = 1.0/(exp(k*Eff) + 1.0), os = 1.0/(exp(k*Ess) + 1.0);
  const double c4f = 1.0 - of, c4s = 1.0 - os;
  const double dc4f = k*(of - of*of), dc4s = k*(os - os*os);
  const double ddc4f = k*k*(-of + 3.0*(of*of) - 2.0*cube_d(of)), ddc4s = k*k*(-os + 3.0*(os*os) - 2.0*cube_d(os));
  // stress coefficients
  const double g1 = h.a*exp(h.b*(Inv1 - 3.0));
  const double g2 = 2.0*h.afs*exp(h.bfs*Efs*Efs);
  const double rexpf = exp(h.bff*Eff*Eff), rexps = exp(h.bss*Ess*Ess);
  double gff = c4f*Eff*rexpf; gff = gff + (0.5*dc4f/h.bff)*(rexpf - 1.0); gff = 2.0*h.aff*gff + h.Tfa;
  double gss = c4s*Ess*rexps; gss = gss + (0.5*dc4s/h.bss)*(rexps - 1.0); gss = 2.0*h.ass*gss + h.Tsa;
  double Sb[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      H[0][i][j] = (i == j) ? 1.0 : 0.0;
      H[1][i][j] = 0.5*(f0[i]*s0[j] + f0[j]*s0[i]);
      H[2][i][j] = f0[i]*f0[j];
      H[3][i][j] = s0[i]*s0[j];
      Sb[i][j] = g1*H[0][i][j] + g2*Efs*H[1][i][j];
      Sb[i][j] += gff*H[2][i][j];
      Sb[i][j] += gss*H[3][i][j];
    }
  // stiffness coefficients
  gk[0] = g1*2.0*J4d*h.b;
  gk[1] = g2*2.0*J4d*(1.0 + 2.0*h.bfs*Efs*Efs);
  {
    double t = c4f*(1.0 + 2.0*h.bff*Eff*Eff); t = (t + 2.0*dc4f*Eff)*rexpf; t = t + (0.5*ddc4f/h.bff)*(rexpf - 1.0);
    gk[2] = 4.0*J4d*h.aff*t;
    double u = c4s*(1.0 + 2.0*h.bss*Ess*Ess); u = (u + 2.0*dc4s*Ess)*rexps; u = u + (0.5*ddc4s/h.bss)*(rexps - 1.0);
    gk[3] = 4.0*J4d*h.ass*u;
  }
  double CSb = 0.0;
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++) CSb = CSb + C[i][j]*Sb[i][j];
  r1 = J2d*CSb/nd;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) S[i][j] = J2d*Sb[i][j] - r1*Ci[i][j];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    double ch = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) ch += C[i][j]*H[q][i][j];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) H[q][i][j] = H[q][i][j] - (1.0/nd)*ch*Ci[i][j];
  }
}

// get_pk2cc<3> for the isotropic laws without fibres / active stress, + get_svol_p.
// Outputs S (sym: 00 11 22 01 12 20) and the upper triangle of Dm (Voigt order 00 11 22 01 12 20).
// fl: the element's fibre (fl[0..2]) and sheet (fl[3..5]) directions, read by the Holzapfel-Ogden law only.
SVB_HD_NOINL void pk2cc_iso(const SolidConsts& c, const double F[3][3], const double* fl,
                          double* S6, double* Dm21)
{
  // mat_det<3> (mat_fun_carray.h:92-122): cofactor expansion along the first row
  const double J = ((0.0 + 1.0*F[0][0]*(F[1][1]*F[2][2] - F[1][2]*F[2][1]))
                    + (-1.0)*F[0][1]*(F[1][0]*F[2][2] - F[1][2]*F[2][0]))
                    + 1.0*F[0][2]*(F[1][0]*F[2][1] - F[1][1]*F[2][0]);
  const double nd = 3.0;
  const double J2d = pow(J, -2.0/nd);
  double C[3][3], Ci[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[i][j] = (0.0 + F[0][i]*F[0][j]) + F[1][i]*F[1][j] + F[2][i]*F[2][j];
  {
    const double d = ((0.0 + C[0][0]*(C[1][1]*C[2][2] - C[1][2]*C[2][1]))
                      - C[0][1]*(C[1][0]*C[2][2] - C[1][2]*C[2][0]))
                      + C[0][2]*(C[1][0]*C[2][1] - C[1][1]*C[2][0]);
    Ci[0][0] = (C[1][1]*C[2][2] - C[1][2]*C[2][1]) / d;
    Ci[0][1] = (C[0][2]*C[2][1] - C[0][1]*C[2][2]) / d;
    Ci[0][2] = (C[0][1]*C[1][2] - C[0][2]*C[1][1]) / d;
    Ci[1][0] = (C[1][2]*C[2][0] - C[1][0]*C[2][2]) / d;
    Ci[1][1] = (C[0][0]*C[2][2] - C[0][2]*C[2][0]) / d;
    Ci[1][2] = (C[0][2]*C[1][0] - C[0][0]*C[1][2]) / d;
    Ci[2][0] = (C[1][0]*C[2][1] - C[1][1]*C[2][0]) / d;
    Ci[2][1] = (C[0][1]*C[2][0] - C[0][0]*C[2][1]) / d;
    Ci[2][2] = (C[0][0]*C[1][1] - C[0][1]*C[1][0]) / d;
  }
  const double trC = C[0][0] + C[1][1] + C[2][2];
  const double Inv1 = J2d*trC;
  double p = 0.0, pl = 0.0;
  if (!is_zero_d(c.Kpen)) {                     // get_svol_p (mat_models.cpp:1626-1645)
    if (c.vol == 1)      { p = c.Kpen*(J - 1.0);        pl = c.Kpen*(2.0*J - 1.0); }
    else if (c.vol == 2) { p = 0.5*c.Kpen*(J - 1.0/J);  pl = c.Kpen*J; }
    else if (c.vol == 3) { p = c.Kpen*(1.0 - 1.0/J);    pl = c.Kpen; }
  }
  const int vi[6] = {0, 1, 2, 0, 1, 2}, vj[6] = {0, 1, 2, 1, 2, 0};
  double S[3][3];
  if (c.iso == 0) {
    // neo-Hookean (mat_models_carray.h:370-434)
    const double g1 = 2.0*c.C10;
    const double r1 = g1*Inv1/nd;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        // Sb = g1 I + Tfa f (x) f (mat_models_carray.h:371-388); without fibres f = 0
        const double Sb = ((i == j) ? g1 : 0.0) + c.Tfa*(fl[i]*fl[j]);
        S[i][j] = J2d*Sb - r1*Ci[i][j];
      }
    const double c2 = 2.0*(r1 - p*J), c3 = pl*J - 2.0*r1/nd;
#pragma unroll
    for (int I = 0; I < 6; I++)
#pragma unroll
      for (int Jv = I; Jv < 6; Jv++) {
        const int i = vi[I], j = vj[I], k = vi[Jv], l = vj[Jv];
        double cc = (-2.0/nd)*(Ci[i][j]*S[k][l] + S[i][j]*Ci[k][l]);
        cc += c2*(0.5*(Ci[i][k]*Ci[j][l] + Ci[i][l]*Ci[j][k])) + c3*(Ci[i][j]*Ci[k][l]);
        Dm21[dm_idx(I, Jv)] = cc;
      }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) S[i][j] += p*J*Ci[i][j];
  } else if (c.iso == 1) {
    // St. Venant-Kirchhoff (:302-322): C10 = lambda, C01 = mu
    const double g1 = c.C10, g2 = c.C01*2.0;
    double E[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) E[i][j] = 0.5*(C[i][j] - ((i == j) ? 1.0 : 0.0));
    const double trE = E[0][0] + E[1][1] + E[2][2];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) S[i][j] = g1*trE*((i == j) ? 1.0 : 0.0) + g2*E[i][j];
#pragma unroll
    for (int I = 0; I < 6; I++)
#pragma unroll
      for (int Jv = I; Jv < 6; Jv++) {
        const int i = vi[I], j = vj[I], k = vi[Jv], l = vj[Jv];
        const double idp = ((i == j) && (k == l)) ? 1.0 : 0.0;
        const double ids = (((i == k) && (j == l)) ? 0.5 : 0.0) + (((i == l) && (j == k)) ? 0.5 : 0.0);
        Dm21[dm_idx(I, Jv)] = g1*idp + g2*ids;
      }
  } else if (c.iso == 3) {
    // Holzapfel-Ogden (mat_models_carray.h:905-1135), see ho_isochoric
    const HoParams hp = {c.ho_a, c.ho_b, c.ho_aff, c.ho_bff, c.ho_ass, c.ho_bss, c.ho_afs, c.ho_bfs, c.ho_khs, c.Tfa, c.Tsa};
    double r1, gk[4], H[4][3][3];
    ho_isochoric(hp, C, Ci, J2d, Inv1, fl, S, r1, gk, H);
    const double c2 = 2.0*(r1 - p*J), c3 = pl*J - 2.0*r1/nd;
#pragma unroll
    for (int I = 0; I < 6; I++)
#pragma unroll
      for (int Jv = I; Jv < 6; Jv++) {
        const int i = vi[I], j = vj[I], kk = vi[Jv], l = vj[Jv];
        double cc = gk[0]*H[0][i][j]*H[0][kk][l] + gk[1]*H[1][i][j]*H[1][kk][l] + gk[2]*H[2][i][j]*H[2][kk][l] + gk[3]*H[3][i][j]*H[3][kk][l];
        cc -= (2.0/nd)*(Ci[i][j]*S[kk][l] + S[i][j]*Ci[kk][l]);
        cc += c2*(0.5*(Ci[i][kk]*Ci[j][l] + Ci[i][l]*Ci[j][kk])) + c3*(Ci[i][j]*Ci[kk][l]);
        Dm21[dm_idx(I, Jv)] = cc;
      }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) S[i][j] += p*J*Ci[i][j];
  } else if (c.iso == 4) {
    // Mooney-Rivlin (:438-540): Sb = g1 I + g2 J2d C (+ Tfa f (x) f), CCb = gk (I (x) I - Ids); the isochoric tangent
    // PP : CCb : PP^T in closed form with PP = Ids - (1/3) Ci (x) C:
    //   PP : (I (x) I) : PP^T = Hd (x) Hd,  Hd = I - (1/3) tr(C) Ci
    //   PP : Ids : PP^T       = Ids - (1/3)(C (x) Ci + Ci (x) C) + (1/9)(C : C) Ci (x) Ci
    const double J4d = J2d*J2d;
    const double g1 = 2.0*(c.C10 + Inv1*c.C01), g2 = -2.0*c.C01;
    double Sb[3][3], Hd[3][3];
    double CSb = 0.0, CC2 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        Sb[i][j] = (((i == j) ? g1 : 0.0) + g2*J2d*C[i][j]) + c.Tfa*(fl[i]*fl[j]);
        Hd[i][j] = ((i == j) ? 1.0 : 0.0) - (1.0/nd)*trC*Ci[i][j];
      }
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int i = 0; i < 3; i++) { CSb = CSb + C[i][j]*Sb[i][j]; CC2 = CC2 + C[i][j]*C[i][j]; }
    const double gk = 4.0*J4d*c.C01;
    const double r1 = J2d*CSb/nd;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) S[i][j] = J2d*Sb[i][j] - r1*Ci[i][j];
    const double c2 = 2.0*(r1 - p*J), c3 = pl*J - 2.0*r1/nd;
#pragma unroll
    for (int I = 0; I < 6; I++)
#pragma unroll
      for (int Jv = I; Jv < 6; Jv++) {
        const int i = vi[I], j = vj[I], k = vi[Jv], l = vj[Jv];
        const double ids = (((i == k) && (j == l)) ? 0.5 : 0.0) + (((i == l) && (j == k)) ? 0.5 : 0.0);
        double cc = gk*(Hd[i][j]*Hd[k][l] - (ids - (1.0/nd)*(C[i][j]*Ci[k][l] + Ci[i][j]*C[k][l]) + (1.0/(nd*nd))*CC2*(Ci[i][j]*Ci[k][l])));
        cc -= (2.0/nd)*(Ci[i][j]*S[k][l] + S[i][j]*Ci[k][l]);
        cc += c2*(0.5*(Ci[i][k]*Ci[j][l] + Ci[i][l]*Ci[j][k])) + c3*(Ci[i][j]*Ci[k][l]);
        Dm21[dm_idx(I, Jv)] = cc;
      }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) S[i][j] += p*J*Ci[i][j];
  } else if (c.iso == 5) {
    // Holzapfel-Gasser-Ogden with fibre dispersion kap (:544-688): two families H = kap I + (1 - 3 kap) f (x) f, exponential
    // terms in E = kap Inv1 + (1 - 3 kap) Inv4 - 1; C10 = the isotropic modulus, aff / bff / ass / bss the fibre parameters.
    // Isochoric tangent PP : (k1 Hf (x) Hf + k2 Hs (x) Hs) : PP^T through the projected factors Hd = H - (1/3)(C : H) Ci.
    const double J4d = J2d*J2d, kap = c.kap;
    const double f0[3] = {fl[0], fl[1], fl[2]}, s0[3] = {fl[3], fl[4], fl[5]};
    double Cf[3], Cs[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      Cf[i] = C[i][0]*f0[0] + C[i][1]*f0[1] + C[i][2]*f0[2];
      Cs[i] = C[i][0]*s0[0] + C[i][1]*s0[1] + C[i][2]*s0[2];
    }
    const double Inv4 = J2d*(f0[0]*Cf[0] + f0[1]*Cf[1] + f0[2]*Cf[2]);
    const double Inv6 = J2d*(s0[0]*Cs[0] + s0[1]*Cs[1] + s0[2]*Cs[2]);
    const double Eff = kap*Inv1 + (1.0 - 3.0*kap)*Inv4 - 1.0;
    const double Ess = kap*Inv1 + (1.0 - 3.0*kap)*Inv6 - 1.0;
    const double ef = exp(c.ho_bff*Eff*Eff), es = exp(c.ho_bss*Ess*Ess);
    const double g1 = c.C10, g2 = c.ho_aff*Eff*ef, g3 = c.ho_ass*Ess*es;
    const double k1 = 4.0*J4d*(c.ho_aff*(1.0 + 2.0*c.ho_bff*Eff*Eff)*ef);
    const double k2 = 4.0*J4d*(c.ho_ass*(1.0 + 2.0*c.ho_bss*Ess*Ess)*es);
    double Hf[3][3], Hs[3][3], Sb[3][3];
    double CSb = 0.0, CHf = 0.0, CHs = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const double d = (i == j) ? 1.0 : 0.0;
        Hf[i][j] = kap*d + (1.0 - 3.0*kap)*(f0[i]*f0[j]);
        Hs[i][j] = kap*d + (1.0 - 3.0*kap)*(s0[i]*s0[j]);
        Sb[i][j] = 2.0*(g1*d + g2*Hf[i][j] + g3*Hs[i][j]) + c.Tfa*(f0[i]*f0[j]);
      }
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int i = 0; i < 3; i++) { CSb = CSb + C[i][j]*Sb[i][j]; CHf = CHf + C[i][j]*Hf[i][j]; CHs = CHs + C[i][j]*Hs[i][j]; }
    const double r1 = J2d*CSb/nd;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        S[i][j] = J2d*Sb[i][j] - r1*Ci[i][j];
        Hf[i][j] = Hf[i][j] - (1.0/nd)*CHf*Ci[i][j];
        Hs[i][j] = Hs[i][j] - (1.0/nd)*CHs*Ci[i][j];
      }
    const double c2 = 2.0*(r1 - p*J), c3 = pl*J - 2.0*r1/nd;
#pragma unroll
    for (int I = 0; I < 6; I++)
#pragma unroll
      for (int Jv = I; Jv < 6; Jv++) {
        const int i = vi[I], j = vj[I], k = vi[Jv], l = vj[Jv];
        double cc = k1*Hf[i][j]*Hf[k][l] + k2*Hs[i][j]*Hs[k][l];
        cc -= (2.0/nd)*(Ci[i][j]*S[k][l] + S[i][j]*Ci[k][l]);
        cc += c2*(0.5*(Ci[i][k]*Ci[j][l] + Ci[i][l]*Ci[j][k])) + c3*(Ci[i][j]*Ci[k][l]);
        Dm21[dm_idx(I, Jv)] = cc;
      }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) S[i][j] += p*J*Ci[i][j];
  } else if (c.iso == 6) {
    // Guccione (1995), transversely isotropic in the local frame (f, s, f x s) (:692-903): Q quadratic in the isochoric
    // Green strain E* of that frame, W = C10/2 (exp Q - 1); parameters C10, bff, bss, bfs.  Isochoric tangent
    //   CCb = r2 J4d (2 Sq (x) Sq + bff M0 (x) M0 + bss (M1 (x) M1 + M2 (x) M2 + 2 M4 (x) M4) + 2 bfs (M3 (x) M3 + M5 (x) M5)),
    // seven rank-one terms, projected through Hd = H - (1/3)(C : H) Ci like the Holzapfel-Ogden law.
    const double J4d = J2d*J2d;
    const double g1 = c.ho_bff, g2 = c.ho_bss, g3 = c.ho_bfs;
    const double f0[3] = {fl[0], fl[1], fl[2]}, s0[3] = {fl[3], fl[4], fl[5]};
    const double n0[3] = {f0[1]*s0[2] - f0[2]*s0[1], f0[2]*s0[0] - f0[0]*s0[2], f0[0]*s0[1] - f0[1]*s0[0]};
    double Rm[3][3], Eb[3][3], E1[3][3], Es[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++) { Rm[i][0] = f0[i]; Rm[i][1] = s0[i]; Rm[i][2] = n0[i]; }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Eb[i][j] = 0.50*(J2d*C[i][j] - ((i == j) ? 1.0 : 0.0));
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) E1[i][j] = ((0.0 + Eb[i][0]*Rm[0][j]) + Eb[i][1]*Rm[1][j]) + Eb[i][2]*Rm[2][j];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Es[i][j] = ((0.0 + Rm[0][i]*E1[0][j]) + Rm[1][i]*E1[1][j]) + Rm[2][i]*E1[2][j];
    const double QQ = g1*Es[0][0]*Es[0][0]
                    + g2*(Es[1][1]*Es[1][1] + Es[2][2]*Es[2][2] + Es[1][2]*Es[1][2] + Es[2][1]*Es[2][1])
                    + g3*(Es[0][1]*Es[0][1] + Es[1][0]*Es[1][0] + Es[0][2]*Es[0][2] + Es[2][0]*Es[2][0]);
    double r2 = c.C10*exp(QQ);
    // H[0] = Sq (the stress direction before the exp factor), H[1..6] = M0..M5
    double H[7][3][3];
    const int pa[6] = {0, 1, 2, 0, 1, 2}, pb[6] = {0, 1, 2, 1, 2, 0};
#pragma unroll
    for (int q = 0; q < 6; q++)
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
          H[1 + q][i][j] = (q < 3) ? Rm[i][pa[q]]*Rm[j][pa[q]] : 0.5*(Rm[i][pa[q]]*Rm[j][pb[q]] + Rm[j][pa[q]]*Rm[i][pb[q]]);
    double Sb[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        H[0][i][j] = g1*Es[0][0]*H[1][i][j]
                   + g2*(Es[1][1]*H[2][i][j] + Es[2][2]*H[3][i][j] + 2.0*Es[1][2]*H[5][i][j])
                   + 2.0*g3*(Es[0][1]*H[4][i][j] + Es[0][2]*H[6][i][j]);
        Sb[i][j] = H[0][i][j]*r2 + c.Tfa*(f0[i]*f0[j]);
      }
    double CSb = 0.0;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int i = 0; i < 3; i++) CSb = CSb + C[i][j]*Sb[i][j];
    const double r1 = J2d*CSb/nd;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) S[i][j] = J2d*Sb[i][j] - r1*Ci[i][j];
    r2 = r2*J4d;
    const double wk[7] = {2.0*r2, g1*r2, g2*r2, g2*r2, 2.0*g3*r2, 2.0*g2*r2, 2.0*g3*r2};
#pragma unroll
    for (int q = 0; q < 7; q++) {
      double ch = 0.0;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) ch += C[i][j]*H[q][i][j];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) H[q][i][j] = H[q][i][j] - (1.0/nd)*ch*Ci[i][j];
    }
    const double c2 = 2.0*(r1 - p*J), c3 = pl*J - 2.0*r1/nd;
#pragma unroll
    for (int I = 0; I < 6; I++)
#pragma unroll
      for (int Jv = I; Jv < 6; Jv++) {
        const int i = vi[I], j = vj[I], k = vi[Jv], l = vj[Jv];
        double cc = 0.0;
#pragma unroll
        for (int q = 0; q < 7; q++) cc += wk[q]*H[q][i][j]*H[q][k][l];
        cc -= (2.0/nd)*(Ci[i][j]*S[k][l] + S[i][j]*Ci[k][l]);
        cc += c2*(0.5*(Ci[i][k]*Ci[j][l] + Ci[i][l]*Ci[j][k])) + c3*(Ci[i][j]*Ci[k][l]);
        Dm21[dm_idx(I, Jv)] = cc;
      }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) S[i][j] += p*J*Ci[i][j];
  } else if (c.iso == 7) {
    // Holzapfel-Ogden with modified anisotropy (HO-ma; mat_models_carray.h:1137-1353, mat_models.cpp:513-610 / :963-1054 for
    // the mixed form): the isotropic exponential term is split isochorically (projected rank-one factor Hd = I - (1/3) tr(C) Ci,
    // like the neo-Hookean part of the HO law), the fibre / sheet / fibre-sheet terms use the FULL invariants
    // Inv4 = f.Cf, Inv6 = s.Cs, Inv8 = f.Cs and enter S and CC unprojected, after the volumetric terms.
    const double J4d = J2d*J2d;
    const double f0[3] = {fl[0], fl[1], fl[2]}, s0[3] = {fl[3], fl[4], fl[5]};
    double Cf[3], Cs[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      Cf[i] = C[i][0]*f0[0] + C[i][1]*f0[1] + C[i][2]*f0[2];
      Cs[i] = C[i][0]*s0[0] + C[i][1]*s0[1] + C[i][2]*s0[2];
    }
    const double Eff = (f0[0]*Cf[0] + f0[1]*Cf[1] + f0[2]*Cf[2]) - 1.0;
    const double Ess = (s0[0]*Cs[0] + s0[1]*Cs[1] + s0[2]*Cs[2]) - 1.0;
    const double Efs = f0[0]*Cs[0] + f0[1]*Cs[1] + f0[2]*Cs[2];
    const double k = c.ho_khs;
    const double of = 1.0/(exp(k*Eff) + 1.0), os = 1.0/(exp(k*Ess) + 1.0);
    const double c4f = 1.0 - of, c4s = 1.0 - os;
    const double dc4f = k*(of - of*of), dc4s = k*(os - os*os);
    const double ddc4f = k*k*(-of + 3.0*(of*of) - 2.0*cube_d(of)), ddc4s = k*k*(-os + 3.0*(os*os) - 2.0*cube_d(os));
    const double g1 = c.ho_a*exp(c.ho_b*(Inv1 - 3.0));
    const double r1 = J2d/nd*(g1*trC);
    const double gk0 = g1*2.0*J4d*c.ho_b;
    double Hd[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const double d = (i == j) ? 1.0 : 0.0;
        S[i][j] = J2d*(g1*d) - r1*Ci[i][j];
        Hd[i][j] = d - (1.0/nd)*trC*Ci[i][j];
      }
    // anisotropic coefficients (stress: s_fs, s_ff, s_ss; stiffness: k_fs, k_ff, k_ss)
    const double efs = 2.0*c.ho_afs*exp(c.ho_bfs*Efs*Efs);
    const double s_fs = efs*Efs, k_fs = efs*2.0*(1.0 + 2.0*c.ho_bfs*Efs*Efs);
    const double rexpf = exp(c.ho_bff*Eff*Eff), rexps = exp(c.ho_bss*Ess*Ess);
    double s_ff = c4f*Eff*rexpf; s_ff = s_ff + (0.5*dc4f/c.ho_bff)*(rexpf - 1.0); s_ff = (2.0*c.ho_aff*s_ff) + c.Tfa;
    double k_ff = c4f*(1.0 + (2.0*c.ho_bff*Eff*Eff)); k_ff = (k_ff + (2.0*dc4f*Eff))*rexpf; k_ff = k_ff + (0.5*ddc4f/c.ho_bff)*(rexpf - 1.0); k_ff = 4.0*c.ho_aff*k_ff;
    double s_ss = c4s*Ess*rexps; s_ss = s_ss + (0.5*dc4s/c.ho_bss)*(rexps - 1.0); s_ss = 2.0*c.ho_ass*s_ss + c.Tsa;
    double k_ss = c4s*(1.0 + (2.0*c.ho_bss*Ess*Ess)); k_ss = (k_ss + (2.0*dc4s*Ess))*rexps; k_ss = k_ss + (0.5*ddc4s/c.ho_bss)*(rexps - 1.0); k_ss = 4.0*c.ho_ass*k_ss;
    const double c2 = 2.0*(r1 - p*J), c3 = pl*J - 2.0*r1/nd;
#pragma unroll
    for (int I = 0; I < 6; I++)
#pragma unroll
      for (int Jv = I; Jv < 6; Jv++) {
        const int i = vi[I], j = vj[I], kk = vi[Jv], l = vj[Jv];
        double cc = gk0*Hd[i][j]*Hd[kk][l];
        cc -= (2.0/nd)*(Ci[i][j]*S[kk][l] + S[i][j]*Ci[kk][l]);                 // S: the isotropic isochoric stress only
        cc += c2*(0.5*(Ci[i][kk]*Ci[j][l] + Ci[i][l]*Ci[j][kk])) + c3*(Ci[i][j]*Ci[kk][l]);
        cc += k_fs*((0.5*(f0[i]*s0[j] + f0[j]*s0[i]))*(0.5*(f0[kk]*s0[l] + f0[l]*s0[kk])));
        cc += k_ff*((f0[i]*f0[j])*(f0[kk]*f0[l]));
        cc += k_ss*((s0[i]*s0[j])*(s0[kk]*s0[l]));
        Dm21[dm_idx(I, Jv)] = cc;
      }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        S[i][j] += p*J*Ci[i][j];
        S[i][j] += s_fs*(0.5*(f0[i]*s0[j] + f0[j]*s0[i]));
        S[i][j] += s_ff*(f0[i]*f0[j]);
        S[i][j] += s_ss*(s0[i]*s0[j]);
      }
  } else {
    // modified St. Venant-Kirchhoff (:326-356): C10 = kappa, C01 = mu
    const double g1 = c.C10, g2 = c.C01;
    const double lJ = log(J);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) S[i][j] = g1*lJ*Ci[i][j] + g2*(C[i][j] - ((i == j) ? 1.0 : 0.0));
#pragma unroll
    for (int I = 0; I < 6; I++)
#pragma unroll
      for (int Jv = I; Jv < 6; Jv++) {
        const int i = vi[I], j = vj[I], k = vi[Jv], l = vj[Jv];
        const double ids = (((i == k) && (j == l)) ? 0.5 : 0.0) + (((i == l) && (j == k)) ? 0.5 : 0.0);
        const double sym = 0.5*(Ci[i][k]*Ci[j][l] + Ci[i][l]*Ci[j][k]);
        Dm21[dm_idx(I, Jv)] = g1*(-2.0*lJ*sym + Ci[i][j]*Ci[k][l]) + 2.0*g2*ids;
      }
  }
  S6[0] = S[0][0]; S6[1] = S[1][1]; S6[2] = S[2][2]; S6[3] = S[0][1]; S6[4] = S[1][2]; S6[5] = S[2][0];
}

// ---------------------------------------------------------------------------------------------------------------------------
// Solid viscosity (mat_models_carray.h:1383-1590, get_visc_stress_and_tangent<3>): viscous 2nd Piola-Kirchhoff stress of a
// Gauss point and its tangent contributions per node pair.  model 1: Newtonian (viscType_Newtonian, :1470-1560: pull-back of
// the deviatoric Cauchy stress 2 mu d_dev), model 2: pseudo-potential (viscType_Potential, :1383-1450: S = mu sym(F^T dv/dX)).
// visc_point fills Svis and a record v[VISC_REC] of the point's matrices; visc_pair forms Kvis_u / Kvis_v (row-major 3x3,
// entry i*3+j) of the pair (a, b) from the record and the two nodes' reference gradients.  Sums run in the reference's order.
// ---------------------------------------------------------------------------------------------------------------------------
enum { VISC_REC = 28 };

SVB_HD void mm3(const double* A, const double* B, double* R)          // mat_mul<3>, row-major 3x3
{
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double sum = 0.0;
#pragma unroll
      for (int k = 0; k < 3; k++) sum += A[i*3 + k]*B[k*3 + j];
      R[i*3 + j] = sum;
    }
}

SVB_HD_NOINL void visc_point(int model, double mu, const double F[3][3], const double vx[3][3], double Svis[3][3], double* v)
{
  double Fm[9], Vm[9], Ft[9], Vt[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) { Fm[i*3 + j] = F[i][j]; Vm[i*3 + j] = vx[i][j]; Ft[j*3 + i] = F[i][j]; Vt[j*3 + i] = vx[i][j]; }
  if (model == 2) {
    // v[0..8] = F vx^T, v[9..17] = F F^T, v[18..26] = vx
    double FtV[9];
    mm3(Fm, Ft, v + 9);
    mm3(Ft, Vm, FtV);
    mm3(Fm, Vt, v);
#pragma unroll
    for (int i = 0; i < 9; i++) v[18 + i] = Vm[i];
    v[27] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Svis[i][j] = mu*(0.5*(FtV[i*3 + j] + FtV[j*3 + i]));
  } else {
    // v[0..8] = F^-1, v[9..17] = d_dev, v[18..26] = vx F^-1, v[27] = J
    const double J = ((0.0 + 1.0*F[0][0]*(F[1][1]*F[2][2] - F[1][2]*F[2][1]))
                      + (-1.0)*F[0][1]*(F[1][0]*F[2][2] - F[1][2]*F[2][0]))
                      + 1.0*F[0][2]*(F[1][0]*F[2][1] - F[1][1]*F[2][0]);
    double* Fi = v;
    Fi[0] = (F[1][1]*F[2][2] - F[1][2]*F[2][1]) / J;
    Fi[1] = (F[0][2]*F[2][1] - F[0][1]*F[2][2]) / J;
    Fi[2] = (F[0][1]*F[1][2] - F[0][2]*F[1][1]) / J;
    Fi[3] = (F[1][2]*F[2][0] - F[1][0]*F[2][2]) / J;
    Fi[4] = (F[0][0]*F[2][2] - F[0][2]*F[2][0]) / J;
    Fi[5] = (F[0][2]*F[1][0] - F[0][0]*F[1][2]) / J;
    Fi[6] = (F[1][0]*F[2][1] - F[1][1]*F[2][0]) / J;
    Fi[7] = (F[0][1]*F[2][0] - F[0][0]*F[2][1]) / J;
    Fi[8] = (F[0][0]*F[1][1] - F[0][1]*F[1][0]) / J;
    double* vF = v + 18;
    mm3(Vm, Fi, vF);
    double sy[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) sy[i*3 + j] = 0.5*(vF[i*3 + j] + vF[j*3 + i]);
    const double tr = ((0.0 + sy[0]) + sy[4]) + sy[8];
    double* dd = v + 9;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) dd[i*3 + j] = sy[i*3 + j] - (tr/3.0)*((i == j) ? 1.0 : 0.0);
    v[27] = J;
    double Fit[9], dFit[9], X[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Fit[j*3 + i] = Fi[i*3 + j];
    mm3(dd, Fit, dFit);
    mm3(Fi, dFit, X);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Svis[i][j] = 2.0*mu*J*X[i*3 + j];
  }
}

SVB_HD void visc_pair(int model, double mu, const double* v, const double* Fm /* row-major F */, const double* na, const double* nb,
                      double* Ku, double* Kv)
{
  if (model == 2) {
    const double* FVt = v; const double* FFt = v + 9; const double* Vm = v + 18;
    double FNa[3], FNb[3], VNa[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      FNa[i] = ((0.0 + Fm[i*3]*na[0]) + Fm[i*3 + 1]*na[1]) + Fm[i*3 + 2]*na[2];
      FNb[i] = ((0.0 + Fm[i*3]*nb[0]) + Fm[i*3 + 1]*nb[1]) + Fm[i*3 + 2]*nb[2];
      VNa[i] = ((0.0 + Vm[i*3]*na[0]) + Vm[i*3 + 1]*na[1]) + Vm[i*3 + 2]*na[2];
    }
    const double NN = ((0.0 + na[0]*nb[0]) + na[1]*nb[1]) + na[2]*nb[2];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        Ku[i*3 + j] = 0.5*mu*(FNb[i]*VNa[j] + NN*FVt[i*3 + j]);
        Kv[i*3 + j] = 0.5*mu*(NN*FFt[i*3 + j] + FNb[i]*FNa[j]);
      }
  } else {
    const double* Fi = v; const double* dd = v + 9; const double* vF = v + 18;
    const double J = v[27];
    double NFa[3], NFb[3], dNa[3], dNb[3], vNa[3], vNb[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      NFa[i] = ((0.0 + na[0]*Fi[i]) + na[1]*Fi[3 + i]) + na[2]*Fi[6 + i];
      NFb[i] = ((0.0 + nb[0]*Fi[i]) + nb[1]*Fi[3 + i]) + nb[2]*Fi[6 + i];
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      dNa[i] = ((0.0 + dd[i*3]*NFa[0]) + dd[i*3 + 1]*NFa[1]) + dd[i*3 + 2]*NFa[2];
      dNb[i] = ((0.0 + dd[i*3]*NFb[0]) + dd[i*3 + 1]*NFb[1]) + dd[i*3 + 2]*NFb[2];
      vNa[i] = ((0.0 + vF[i]*NFa[0]) + vF[3 + i]*NFa[1]) + vF[6 + i]*NFa[2];
      vNb[i] = ((0.0 + vF[i]*NFb[0]) + vF[3 + i]*NFb[1]) + vF[6 + i]*NFb[2];
    }
    const double NN = ((0.0 + NFa[0]*NFb[0]) + NFa[1]*NFb[1]) + NFa[2]*NFb[2];
    const double r2d = 2.0/3.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        Ku[i*3 + j] = mu*J*(2.0*(dNa[i]*NFb[j] - dNb[i]*NFa[j]) - (NN*vF[i*3 + j] + NFb[i]*vNa[j] - r2d*NFa[i]*vNb[j]));
        Kv[i*3 + j] = mu*J*(NN*((i == j) ? 1.0 : 0.0) + NFb[i]*NFa[j] - r2d*NFa[i]*NFb[j]);
      }
  }
}

} // namespace svb200
