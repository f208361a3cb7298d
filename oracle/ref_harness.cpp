// TEST INFRASTRUCTURE ONLY (oracle).  C-ABI harness around the UNMODIFIED reference sources
// (compiled from where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libsvref.so).  It fills the reference's own ComMod / mshType / eqType / FSILS_lhsType
// objects programmatically from flat arrays (the reference's VTK readers are not available here)
// and then calls the reference's own functions:
//
//   lhsa_ns::lhsa                 Code/Source/solver/lhsa.cpp:153
//   nn::select_ele                Code/Source/solver/nn.cpp:943
//   fs::init_fs_msh               Code/Source/solver/fs.cpp:258
//   fluid::construct_fluid        Code/Source/solver/fluid.cpp:464   (-> do_assem lhsa.cpp:97)
//   fsils_commu_create            Code/Source/liner_solver/commu.cpp:44
//   fsils_lhs_create              Code/Source/liner_solver/lhs.cpp:57
//   fsils_bc_create               Code/Source/liner_solver/bc.cpp:45
//   fsils_solve                   Code/Source/liner_solver/solve.cpp:50
//   fsils_spar_mul_vv / commuv    Code/Source/liner_solver/spar_mul.cpp:191, in_commu.cpp:111
//   pic::picp / pici / picc       Code/Source/solver/pic.cpp:591, 486, 74
//   nn::select_eleb, eq_assem::b_assem_neu_bc   Code/Source/solver/nn.cpp:997, eq_assem.cpp:58
//
// Nothing in the product (svfsiplus_b200/) links or loads this file; only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs do.
#include "Simulation.h"
#include "ComMod.h"
#include "FsilsLinearAlgebra.h"
#include "all_fun.h"
#include "consts.h"
#include "fluid.h"
#include "fs.h"
#include "mat_fun.h"
#include "mat_models.h"
#include "mat_models_carray.h"
#include "fsi.h"
#include "l_elas.h"
#include "mesh.h"
#include "sv_struct.h"
#include "ustruct.h"
#include "fsils_api.hpp"
#include "lhsa.h"
#include "nn.h"
#include "spar_mul.h"
#include "commu.h"
#include "lhs.h"
#include "ls.h"
#include "pic.h"
#include "eq_assem.h"
#include "output.h"
#ifdef WITH_B200_DROPIN
#include "B200LinearAlgebra.h"
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
// test infrastructure: a native backtrace on SIGSEGV / SIGABRT (the Python fault handler only shows the ctypes call)
static void dropin_crash_handler(int sig)
{
  void* frames[64];
  const int n = backtrace(frames, 64);
  const char msg[] = "[dropin harness] fatal signal, native backtrace:\n";
  (void)!write(2, msg, sizeof(msg) - 1);
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}
static void dropin_install_crash_handler()
{
  static bool done = false;
  if (done) return;
  done = true;
  signal(SIGSEGV, dropin_crash_handler);
  signal(SIGABRT, dropin_crash_handler);
}
#endif

#include "mpi.h"

#include <chrono>
#include <optional>
#include <functional>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

using namespace fsi_linear_solver;

namespace {

thread_local std::string t_err;
std::string g_err;

double now_s()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ----------------------------------------------------------------------------------------------
// Assembly context: one Simulation (ComMod) with one mesh and one equation.
// ----------------------------------------------------------------------------------------------
struct AsmCtx {
  std::unique_ptr<Simulation> sim;
  int nnz = 0;
  int visc_model = 0;        // solid viscosity applied by ref_asm_solid: 0 none, 1 Newtonian, 2 potential (ref_asm_set_visc)
  double visc_mu = 0.0;
  std::vector<double> pS0;   // nodal prestress (6 x nNo) applied by ref_asm_solid, empty = none (ref_asm_set_prestress)
  bool pstEq = false;
};

} // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// Create a ComMod with one 3-D mesh of nEl elements with eNoN nodes each (4 = TET4, 8 = HEX8,
// 10 = TET10), call the reference's select_ele / init_fs_msh / lhsa.  Returns an opaque context.
void* ref_asm_create(int nNo, int nEl, int eNoN, const int* IEN, const double* x, int nFs, double qmTET4)
{
  try {
    mpistub_set_world(1);
    mpistub_bind_rank(0);
    auto ctx = new AsmCtx;
    ctx->sim.reset(new Simulation());
    auto& com_mod = ctx->sim->com_mod;
    com_mod.nsd = 3;
    com_mod.nsymd = 6;
    com_mod.tnNo = nNo;
    com_mod.gtnNo = nNo;
    com_mod.nMsh = 1;
    com_mod.msh.resize(1);
    auto& msh = com_mod.msh[0];
    msh.name = "msh";
    msh.nNo = nNo;
    msh.gnNo = nNo;
    msh.nEl = nEl;
    msh.gnEl = nEl;
    msh.eNoN = eNoN;
    msh.nFs = nFs;
    if (qmTET4 > 0.0) msh.qmTET4 = qmTET4;
    msh.IEN.resize(eNoN, nEl);
    std::memcpy(msh.IEN.data(), IEN, sizeof(int)*size_t(eNoN)*nEl);
    msh.gN.resize(nNo);
    for (int a = 0; a < nNo; a++) msh.gN(a) = a;
    com_mod.x.resize(3, nNo);
    std::memcpy(com_mod.x.data(), x, sizeof(double)*3*size_t(nNo));

    nn::select_ele(com_mod, msh);
    fs::init_fs_msh(com_mod, msh);
    mat_fun::ten_init(3);            // what initialize() does once (S/initialize.cpp:547)

    int nnz = 0;
    lhsa_ns::lhsa(ctx->sim.get(), nnz);
    ctx->nnz = nnz;
    com_mod.lhs.nnz = nnz;
    com_mod.lhs.nNo = nNo;
    return ctx;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

void ref_asm_destroy(void* h) { delete static_cast<AsmCtx*>(h); }

// lM.fN(nsd*nFn, nEl) / lM.nFn: fibre directions per element (fN == NULL clears them).
void ref_asm_set_fibers(void* h, int nFn, const double* fN)
{
  auto& msh = static_cast<AsmCtx*>(h)->sim->com_mod.msh[0];
  if (!fN) { msh.nFn = 0; msh.fN.resize(0, 0); return; }
  msh.nFn = nFn;
  msh.fN.resize(3*nFn, msh.nEl);
  std::memcpy(msh.fN.data(), fN, sizeof(double)*size_t(3*nFn)*msh.nEl);
}

int ref_asm_nnz(void* h) { return static_cast<AsmCtx*>(h)->nnz; }

// rowPtr: nNo+1 ints, colPtr: nnz ints (the reference's lhsa output, S/lhsa.cpp:153).
void ref_asm_get_csr(void* h, int* rowPtr, int* colPtr)
{
  auto& com_mod = static_cast<AsmCtx*>(h)->sim->com_mod;
  std::memcpy(rowPtr, com_mod.rowPtr.data(), sizeof(int)*com_mod.rowPtr.size());
  std::memcpy(colPtr, com_mod.colPtr.data(), sizeof(int)*com_mod.colPtr.size());
}

// Gauss tables chosen by the reference for the mesh (for checking the product's own tables).
// w: nG, N: eNoN*nG (col-major N(a,g)), Nx: 3*eNoN*nG (Nx(i,a,g)).  Returns nG.
int ref_asm_get_tables(void* h, double* w, double* N, double* Nx)
{
  auto& msh = static_cast<AsmCtx*>(h)->sim->com_mod.msh[0];
  if (w) std::memcpy(w, msh.w.data(), sizeof(double)*msh.w.size());
  if (N) std::memcpy(N, msh.N.data(), sizeof(double)*msh.N.size());
  if (Nx) std::memcpy(Nx, msh.Nx.data(), sizeof(double)*msh.Nx.size());
  return msh.nG;
}

} // extern "C"

namespace {
// Fill ComMod / eqType / dmnType for one Navier-Stokes equation on the context's mesh.
// visc = {type(0 const,1 Carreau-Yasuda,2 Casson), mu_i, mu_o, lam, a, n}
void configure_fluid(AsmCtx* ctx, int tDof, int mvMsh, double dt, double am, double af, double gam,
                     double rho, const double* f, double Kinv_darcy, const double* visc, const double* Bf)
{
  using namespace consts;
  auto& com_mod = ctx->sim->com_mod;
  const int nNo = com_mod.tnNo;
  const int dof = 4;
  com_mod.tDof = tDof;
  com_mod.dof = dof;
  com_mod.dt = dt;
  com_mod.mvMsh = (mvMsh != 0);
  com_mod.cEq = 0;
  com_mod.nEq = 1;
  if (com_mod.eq.size() != 1) com_mod.eq.resize(1);
  auto& eq = com_mod.eq[0];
  eq.phys = EquationType::phys_fluid;
  eq.dof = dof;
  eq.s = 0;
  eq.e = dof - 1;
  eq.am = am;
  eq.af = af;
  eq.gam = gam;
  eq.nDmn = 1;
  if (eq.dmn.size() != 1) eq.dmn.resize(1);
  auto& dmn = eq.dmn[0];
  dmn.Id = -1;
  dmn.phys = EquationType::phys_fluid;
  dmn.prop[PhysicalProperyType::fluid_density] = rho;
  dmn.prop[PhysicalProperyType::f_x] = f[0];
  dmn.prop[PhysicalProperyType::f_y] = f[1];
  dmn.prop[PhysicalProperyType::f_z] = f[2];
  dmn.prop[PhysicalProperyType::inverse_darcy_permeability] = Kinv_darcy;
  int vt = int(visc[0]);
  dmn.fluid_visc.viscType = (vt == 0) ? FluidViscosityModelType::viscType_Const
                          : (vt == 1) ? FluidViscosityModelType::viscType_CY
                                      : FluidViscosityModelType::viscType_Cass;
  dmn.fluid_visc.mu_i = visc[1];
  dmn.fluid_visc.mu_o = visc[2];
  dmn.fluid_visc.lam = visc[3];
  dmn.fluid_visc.a = visc[4];
  dmn.fluid_visc.n = visc[5];
  com_mod.Bf.resize(3, nNo);
  std::memcpy(com_mod.Bf.data(), Bf, sizeof(double)*3*size_t(nNo));
}
} // namespace

extern "C" {

// Fluid (Navier-Stokes VMS) assembly through the reference's construct_fluid.
// visc = {type(0 const,1 Carreau-Yasuda,2 Casson), mu_i, mu_o, lam, a, n}
// Ag, Yg: tDof x nNo (column-major, i.e. node-contiguous), Bf: 3 x nNo.
// Outputs R (dof x nNo), Val (dof*dof x nnz).  Returns seconds spent inside construct_fluid, <0 on error.
double ref_asm_fluid(void* h, int tDof, int mvMsh, double dt, double am, double af, double gam,
                     double rho, const double* f, double Kinv_darcy, const double* visc,
                     const double* Ag, const double* Yg, const double* Bf, double* R, double* Val)
{
  try {
    using namespace consts;
    auto ctx = static_cast<AsmCtx*>(h);
    auto& com_mod = ctx->sim->com_mod;
    const int nNo = com_mod.tnNo;
    const int dof = 4;
    configure_fluid(ctx, tDof, mvMsh, dt, am, af, gam, rho, f, Kinv_darcy, visc, Bf);
    auto& eq = com_mod.eq[0];
    if (!eq.linear_algebra) eq.linear_algebra = new FsilsLinearAlgebra();

    Array<double> Ag_a(tDof, nNo), Yg_a(tDof, nNo);
    std::memcpy(Ag_a.data(), Ag, sizeof(double)*size_t(tDof)*nNo);
    std::memcpy(Yg_a.data(), Yg, sizeof(double)*size_t(tDof)*nNo);

    // ls_alloc contract (S/ls.cpp:51): R and Val zeroed by resize.
    com_mod.R.resize(dof, nNo);
    eq.linear_algebra->alloc(com_mod, eq);

    double t0 = now_s();
    fluid::construct_fluid(com_mod, com_mod.msh[0], Ag_a, Yg_a);
    double t1 = now_s();

    std::memcpy(R, com_mod.R.data(), sizeof(double)*size_t(dof)*nNo);
    std::memcpy(Val, com_mod.Val.data(), sizeof(double)*size_t(dof)*dof*ctx->nnz);
    return t1 - t0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1.0;
  }
}

} // extern "C"

namespace {
// Fill ComMod / eqType / dmnType for one struct (kind 0), lElas (1) or mesh (2) equation; par as in ref_asm_solid.
void configure_solid(AsmCtx* ctx, int kind, int tDof, int s, const double* par, const double* Do, const double* Bf)
{
  using namespace consts;
  auto& com_mod = ctx->sim->com_mod;
  const int nNo = com_mod.tnNo;
    const int dof = 3;
    com_mod.tDof = tDof;
    com_mod.dof = dof;
    com_mod.dt = par[0];
    com_mod.mvMsh = false;
    com_mod.cEq = 0;
    com_mod.nEq = 1;
    if (com_mod.eq.size() != 1) com_mod.eq.resize(1);
    auto& eq = com_mod.eq[0];
    eq.phys = (kind == 0) ? EquationType::phys_struct : (kind == 1) ? EquationType::phys_lElas : EquationType::phys_mesh;
    eq.dof = dof;
    eq.s = s;
    eq.e = s + dof - 1;
    if (kind == 2) {
      com_mod.Do.resize(tDof, nNo);
      std::memcpy(com_mod.Do.data(), Do, sizeof(double)*size_t(tDof)*nNo);
    }
    eq.am = par[1]; eq.af = par[2]; eq.gam = par[3]; eq.beta = par[4];
    eq.nDmn = 1;
    if (eq.dmn.size() != 1) eq.dmn.resize(1);
    auto& dmn = eq.dmn[0];
    dmn.Id = -1;
    dmn.phys = eq.phys;
    dmn.prop[PhysicalProperyType::solid_density] = par[5];
    dmn.prop[PhysicalProperyType::damping] = par[6];
    dmn.prop[PhysicalProperyType::f_x] = par[7];
    dmn.prop[PhysicalProperyType::f_y] = par[8];
    dmn.prop[PhysicalProperyType::f_z] = par[9];
    dmn.prop[PhysicalProperyType::elasticity_modulus] = par[15];
    dmn.prop[PhysicalProperyType::poisson_ratio] = par[16];
    const int iso = int(par[10]), vol = int(par[11]);
    dmn.stM.isoType = (iso == 0) ? ConstitutiveModelType::stIso_nHook
                    : (iso == 1) ? ConstitutiveModelType::stIso_StVK
                    : (iso == 2) ? ConstitutiveModelType::stIso_mStVK
                    : (iso == 4) ? ConstitutiveModelType::stIso_MR
                    : (iso == 5) ? ConstitutiveModelType::stIso_HGO
                    : (iso == 6) ? ConstitutiveModelType::stIso_Gucci
                    : (iso == 7) ? ConstitutiveModelType::stIso_HO_ma : ConstitutiveModelType::stIso_HO;
    dmn.stM.volType = (vol == 1) ? ConstitutiveModelType::stVol_Quad
                    : (vol == 2) ? ConstitutiveModelType::stVol_ST91
                    : (vol == 3) ? ConstitutiveModelType::stVol_M94 : ConstitutiveModelType::stIso_NA;
    dmn.stM.C10 = par[12];
    dmn.stM.C01 = par[13];
    dmn.stM.Kpen = par[14];
    // Holzapfel-Ogden parameters (par[17..25]) and the mesh's fibre / sheet directions
    dmn.stM.a = par[17]; dmn.stM.b = par[18]; dmn.stM.aff = par[19]; dmn.stM.bff = par[20];
    dmn.stM.ass = par[21]; dmn.stM.bss = par[22]; dmn.stM.afs = par[23]; dmn.stM.bfs = par[24]; dmn.stM.khs = par[25];
    // steady fibre-reinforcement stress (get_fib_stress, S/mat_models.cpp:126): par[26] = Tf.g, par[27] = Tf.eta_s
    dmn.stM.Tf.fType = (par[26] != 0.0) ? utils::ibset(0, iBC_std) : 0;
    dmn.stM.Tf.g = par[26];
    dmn.stM.Tf.eta_s = par[27];
    dmn.stM.kap = par[28];                        // HGO fibre dispersion
    com_mod.Bf.resize(3, nNo);
    std::memcpy(com_mod.Bf.data(), Bf, sizeof(double)*3*size_t(nNo));
}
} // namespace

extern "C" {

// Solid viscosity of the next ref_asm_solid calls (dmn.solid_visc): model 0 none, 1 Newtonian, 2 potential.
void ref_asm_set_visc(void* h, int model, double mu)
{
  auto ctx = static_cast<AsmCtx*>(h);
  ctx->visc_model = model;
  ctx->visc_mu = mu;
}

// Prestress of the next ref_asm_solid calls: pS0 (6 x nNo, or null for none), pstEq switches the pSn / pSa accumulations on.
void ref_asm_set_prestress(void* h, const double* pS0, int pstEq)
{
  auto ctx = static_cast<AsmCtx*>(h);
  const int nNo = ctx->sim->com_mod.tnNo;
  if (pS0) ctx->pS0.assign(pS0, pS0 + size_t(6)*nNo); else ctx->pS0.clear();
  ctx->pstEq = pstEq != 0;
}

// com_mod.pSn (6 x nNo) and com_mod.pSa (nNo) after a ref_asm_solid with pstEq.
void ref_asm_get_prestress(void* h, double* pSn, double* pSa)
{
  auto& com_mod = static_cast<AsmCtx*>(h)->sim->com_mod;
  std::memcpy(pSn, com_mod.pSn.data(), sizeof(double)*6*size_t(com_mod.tnNo));
  std::memcpy(pSa, com_mod.pSa.data(), sizeof(double)*size_t(com_mod.tnNo));
}

// Solid assembly through the reference's construct_dsolid (S/sv_struct.cpp:213 -> struct_3d_carray :552 ->
// get_pk2cc<3> S/mat_models_carray.h:182) or construct_l_elas (S/l_elas.cpp:58 -> l_elas_3d :274).
// kind 0: struct, 1: lElas, 2: mesh (construct_mesh S/mesh.cpp:42; needs Do and eq.s = s).  par = {dt, am, af, gam, beta, rho, dmp, fx, fy, fz,
//   iso (0 nHook, 1 StVK, 2 mStVK), vol (0 none, 1 Quad, 2 ST91, 3 M94), C10, C01, Kpen, elM, nu}
// Ag, Yg, Dg: tDof x nNo (eq.s = 0, dof = 3).  Outputs R (3 x nNo), Val (9 x nnz).
double ref_asm_solid(void* h, int kind, int tDof, int s, const double* par, const double* Ag, const double* Yg,
                     const double* Dg, const double* Do, const double* Bf, double* R, double* Val)
{
  try {
    using namespace consts;
    auto ctx = static_cast<AsmCtx*>(h);
    auto& com_mod = ctx->sim->com_mod;
    const int nNo = com_mod.tnNo;
    const int dof = 3;
    configure_solid(ctx, kind, tDof, s, par, Do, Bf);
    auto& eq = com_mod.eq[0];
    eq.dmn[0].solid_visc.viscType = (ctx->visc_model == 1) ? SolidViscosityModelType::viscType_Newtonian
                                  : (ctx->visc_model == 2) ? SolidViscosityModelType::viscType_Potential : SolidViscosityModelType::viscType_NA;
    eq.dmn[0].solid_visc.mu = ctx->visc_mu;
    // prestress (S/sv_struct.cpp:232-235, 333-343): com_mod.pS0 read by struct_3d, pSn / pSa accumulated when pstEq
    com_mod.pstEq = ctx->pstEq;
    if (ctx->pS0.empty()) com_mod.pS0.clear();
    else { com_mod.pS0.resize(6, nNo); std::memcpy(com_mod.pS0.data(), ctx->pS0.data(), sizeof(double)*6*size_t(nNo)); }
    if (ctx->pstEq) { com_mod.pSn.resize(6, nNo); com_mod.pSa.resize(nNo); com_mod.pSn = 0.0; com_mod.pSa = 0.0; }
    if (!eq.linear_algebra) eq.linear_algebra = new FsilsLinearAlgebra();

    Array<double> Ag_a(tDof, nNo), Yg_a(tDof, nNo), Dg_a(tDof, nNo);
    std::memcpy(Ag_a.data(), Ag, sizeof(double)*size_t(tDof)*nNo);
    std::memcpy(Yg_a.data(), Yg, sizeof(double)*size_t(tDof)*nNo);
    std::memcpy(Dg_a.data(), Dg, sizeof(double)*size_t(tDof)*nNo);

    com_mod.R.resize(dof, nNo);
    eq.linear_algebra->alloc(com_mod, eq);

    double t0 = now_s();
    if (kind == 0) struct_ns::construct_dsolid(com_mod, ctx->sim->cep_mod, com_mod.msh[0], Ag_a, Yg_a, Dg_a);
    else if (kind == 1) l_elas::construct_l_elas(com_mod, com_mod.msh[0], Ag_a, Dg_a);
    else mesh::construct_mesh(com_mod, ctx->sim->cep_mod, com_mod.msh[0], Ag_a, Dg_a);
    double t1 = now_s();

    std::memcpy(R, com_mod.R.data(), sizeof(double)*size_t(dof)*nNo);
    std::memcpy(Val, com_mod.Val.data(), sizeof(double)*size_t(dof)*dof*ctx->nnz);
    return t1 - t0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1.0;
  }
}

// Multi-domain struct (kind 0, dof 3) or fluid (kind 3, dof 4) equation: nDmn domains of the same physics with their own
// properties (eq.dmn[d], Id = d), elem_dmn[e] = domain of element e (lM.eId = 1 << d, S/all_fun.cpp:149).
// struct: par (nDmn x 26) rows as in ref_asm_solid; fluid: par (nDmn x 17) rows = {dt, am, af, gam, rho, fx, fy, fz, Kinv,
// viscType, mu_i, mu_o, lam, a, n, 0, 0}.
//
// only_dmn >= 0 (struct): construct_dsolid keeps a COPY of com_mod.cDmn (S/sv_struct.cpp:229 `auto cDmn`, where
// construct_fluid takes a reference, S/fluid.cpp:491), so struct_3d_carray reads the properties of whatever domain
// com_mod.cDmn was left at, for every element.  To obtain what the element loop evidently intends, the harness assembles
// one domain at a time: domain only_dmn keeps phys_struct (the others are skipped by the phys test :265) and
// com_mod.cDmn = only_dmn; the caller adds the per-domain results.
double ref_asm_domains(void* h, int kind, int tDof, int nDmn, int npar, const double* par, const int* elem_dmn, int only_dmn,
                       const double* Ag, const double* Yg, const double* Dg, const double* Bf, double* R, double* Val)
{
  try {
    using namespace consts;
    auto ctx = static_cast<AsmCtx*>(h);
    auto& com_mod = ctx->sim->com_mod;
    auto& msh = com_mod.msh[0];
    const int nNo = com_mod.tnNo;
    const int dof = (kind == 0) ? 3 : 4;
    std::vector<dmnType> dm;
    for (int d = 0; d < nDmn; d++) {
      const double* q = par + size_t(d)*npar;
      if (kind == 0) configure_solid(ctx, 0, tDof, 0, q, nullptr, Bf);
      else configure_fluid(ctx, tDof, 0, q[0], q[1], q[2], q[3], q[4], q + 5, q[8], q + 9, Bf);
      dm.push_back(com_mod.eq[0].dmn[0]);
      dm.back().Id = d;
    }
    auto& eq = com_mod.eq[0];
    eq.nDmn = nDmn;
    eq.dmn = dm;
    if (only_dmn >= 0) {
      for (int d = 0; d < nDmn; d++) if (d != only_dmn) eq.dmn[d].phys = EquationType::phys_lElas;
      com_mod.cDmn = only_dmn;
    }
    msh.eId.resize(msh.nEl);
    for (int e = 0; e < msh.nEl; e++) msh.eId(e) = 1 << elem_dmn[e];
    if (!eq.linear_algebra) eq.linear_algebra = new FsilsLinearAlgebra();
    Array<double> Ag_a(tDof, nNo), Yg_a(tDof, nNo), Dg_a(tDof, nNo);
    std::memcpy(Ag_a.data(), Ag, sizeof(double)*size_t(tDof)*nNo);
    std::memcpy(Yg_a.data(), Yg, sizeof(double)*size_t(tDof)*nNo);
    if (Dg) std::memcpy(Dg_a.data(), Dg, sizeof(double)*size_t(tDof)*nNo);
    com_mod.R.resize(dof, nNo);
    eq.linear_algebra->alloc(com_mod, eq);
    double t0 = now_s();
    if (kind == 0) struct_ns::construct_dsolid(com_mod, ctx->sim->cep_mod, msh, Ag_a, Yg_a, Dg_a);
    else fluid::construct_fluid(com_mod, msh, Ag_a, Yg_a);
    double t1 = now_s();
    std::memcpy(R, com_mod.R.data(), sizeof(double)*size_t(dof)*nNo);
    std::memcpy(Val, com_mod.Val.data(), sizeof(double)*size_t(dof)*dof*ctx->nnz);
    // leave a single-domain equation behind for the calls that follow on this context
    eq.nDmn = 1; eq.dmn.resize(1); eq.dmn[0].Id = -1; msh.eId.resize(0); com_mod.cDmn = 0;
    return t1 - t0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1.0;
  }
}

// FSI equation through the reference's construct_fsi (S/fsi.cpp:42): domain 0 = fluid (Id 0), domain 1 =
// struct (Id 1); elem_dmn[e] in {0,1} becomes the bit mask lM.eId.  tDof = 7 (FSI unknowns 0..3, mesh 4..6).
// fpar = {rho, fx, fy, fz, visc type, mu_i, mu_o, lam, a, n};  spar as in ref_asm_solid (struct).
double ref_asm_fsi(void* h, int tDof, double dt, double am, double af, double gam, double beta, const double* fpar,
                   const double* spar, const int* elem_dmn, const double* Ag, const double* Yg, const double* Dg,
                   const double* Bf, double* R, double* Val)
{
  try {
    using namespace consts;
    auto ctx = static_cast<AsmCtx*>(h);
    auto& com_mod = ctx->sim->com_mod;
    auto& msh = com_mod.msh[0];
    const int nNo = com_mod.tnNo;
    const int dof = 4;
    com_mod.tDof = tDof;
    com_mod.dof = dof;
    com_mod.dt = dt;
    com_mod.mvMsh = true;
    com_mod.cEq = 0;
    com_mod.nEq = 1;
    if (com_mod.eq.size() != 1) com_mod.eq.resize(1);
    auto& eq = com_mod.eq[0];
    eq.phys = EquationType::phys_FSI;
    eq.dof = dof; eq.s = 0; eq.e = dof - 1;
    eq.am = am; eq.af = af; eq.gam = gam; eq.beta = beta;
    eq.nDmn = 2;
    if (eq.dmn.size() != 2) eq.dmn.resize(2);
    {
      auto& dmn = eq.dmn[0];
      dmn.Id = 0;
      dmn.phys = EquationType::phys_fluid;
      dmn.prop[PhysicalProperyType::fluid_density] = fpar[0];
      dmn.prop[PhysicalProperyType::f_x] = fpar[1];
      dmn.prop[PhysicalProperyType::f_y] = fpar[2];
      dmn.prop[PhysicalProperyType::f_z] = fpar[3];
      dmn.prop[PhysicalProperyType::inverse_darcy_permeability] = 0.0;
      const int vt = int(fpar[4]);
      dmn.fluid_visc.viscType = (vt == 0) ? FluidViscosityModelType::viscType_Const
                              : (vt == 1) ? FluidViscosityModelType::viscType_CY : FluidViscosityModelType::viscType_Cass;
      dmn.fluid_visc.mu_i = fpar[5]; dmn.fluid_visc.mu_o = fpar[6]; dmn.fluid_visc.lam = fpar[7];
      dmn.fluid_visc.a = fpar[8]; dmn.fluid_visc.n = fpar[9];
    }
    {
      auto& dmn = eq.dmn[1];
      dmn.Id = 1;
      dmn.phys = EquationType::phys_struct;
      dmn.prop[PhysicalProperyType::solid_density] = spar[5];
      dmn.prop[PhysicalProperyType::damping] = spar[6];
      dmn.prop[PhysicalProperyType::f_x] = spar[7];
      dmn.prop[PhysicalProperyType::f_y] = spar[8];
      dmn.prop[PhysicalProperyType::f_z] = spar[9];
      const int iso = int(spar[10]), vol = int(spar[11]);
      dmn.stM.isoType = (iso == 0) ? ConstitutiveModelType::stIso_nHook
                      : (iso == 1) ? ConstitutiveModelType::stIso_StVK : ConstitutiveModelType::stIso_mStVK;
      dmn.stM.volType = (vol == 1) ? ConstitutiveModelType::stVol_Quad
                      : (vol == 2) ? ConstitutiveModelType::stVol_ST91
                      : (vol == 3) ? ConstitutiveModelType::stVol_M94 : ConstitutiveModelType::stIso_NA;
      dmn.stM.C10 = spar[12]; dmn.stM.C01 = spar[13]; dmn.stM.Kpen = spar[14];
      // wall viscosity / prestress of the next call (ref_asm_set_visc / ref_asm_set_prestress); construct_fsi reads com_mod.pS0 only
      dmn.solid_visc.viscType = (ctx->visc_model == 1) ? SolidViscosityModelType::viscType_Newtonian
                              : (ctx->visc_model == 2) ? SolidViscosityModelType::viscType_Potential : SolidViscosityModelType::viscType_NA;
      dmn.solid_visc.mu = ctx->visc_mu;
      if (ctx->pS0.empty()) com_mod.pS0.clear();
      else { com_mod.pS0.resize(6, com_mod.tnNo); std::memcpy(com_mod.pS0.data(), ctx->pS0.data(), sizeof(double)*6*size_t(com_mod.tnNo)); }
    }
    msh.eId.resize(msh.nEl);
    for (int e = 0; e < msh.nEl; e++) msh.eId(e) = 1 << elem_dmn[e];
    com_mod.Bf.resize(3, nNo);
    std::memcpy(com_mod.Bf.data(), Bf, sizeof(double)*3*size_t(nNo));
    if (!eq.linear_algebra) eq.linear_algebra = new FsilsLinearAlgebra();

    Array<double> Ag_a(tDof, nNo), Yg_a(tDof, nNo), Dg_a(tDof, nNo);
    std::memcpy(Ag_a.data(), Ag, sizeof(double)*size_t(tDof)*nNo);
    std::memcpy(Yg_a.data(), Yg, sizeof(double)*size_t(tDof)*nNo);
    std::memcpy(Dg_a.data(), Dg, sizeof(double)*size_t(tDof)*nNo);

    com_mod.R.resize(dof, nNo);
    eq.linear_algebra->alloc(com_mod, eq);
    double t0 = now_s();
    fsi::construct_fsi(com_mod, ctx->sim->cep_mod, msh, Ag_a, Yg_a, Dg_a);
    double t1 = now_s();
    std::memcpy(R, com_mod.R.data(), sizeof(double)*size_t(dof)*nNo);
    std::memcpy(Val, com_mod.Val.data(), sizeof(double)*size_t(dof)*dof*ctx->nnz);
    return t1 - t0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1.0;
  }
}

// ustruct equation through the reference's construct_usolid (S/ustruct.cpp:216) and, when Ad != NULL,
// ustruct_r (S/ustruct.cpp:1726, first Newton iteration).  par = {dt, am, af, gam, rho, fx, fy, fz, elM, nu,
// ctM, ctC, vol (0 none,1 Quad,2 ST91,3 M94), C10, Kpen, iso (0 nHook, 3 HO), a, b, aff, bff, ass, bss, afs, bfs, khs}.
// tDof = 4, eq.s = 0.  Fibres: ref_asm_set_fibers.
// Outputs R (4 x nNo), Val (16 x nnz), Kd (12 x nnz).
double ref_asm_ustruct(void* h, int tDof, const double* par, const double* Ag, const double* Yg, const double* Dg,
                       const double* Bf, const double* Ad, double* R, double* Val, double* Kd)
{
  try {
    using namespace consts;
    auto ctx = static_cast<AsmCtx*>(h);
    auto& com_mod = ctx->sim->com_mod;
    const int nNo = com_mod.tnNo;
    const int dof = 4;
    com_mod.tDof = tDof;
    com_mod.dof = dof;
    com_mod.dt = par[0];
    com_mod.mvMsh = false;
    com_mod.cEq = 0;
    com_mod.nEq = 1;
    if (com_mod.eq.size() != 1) com_mod.eq.resize(1);
    auto& eq = com_mod.eq[0];
    eq.phys = EquationType::phys_ustruct;
    eq.dof = dof; eq.s = 0; eq.e = dof - 1;
    eq.am = par[1]; eq.af = par[2]; eq.gam = par[3];
    eq.itr = 1;
    eq.nDmn = 1;
    if (eq.dmn.size() != 1) eq.dmn.resize(1);
    auto& dmn = eq.dmn[0];
    dmn.Id = -1;
    dmn.phys = EquationType::phys_ustruct;
    dmn.solid_visc.viscType = (ctx->visc_model == 1) ? SolidViscosityModelType::viscType_Newtonian
                            : (ctx->visc_model == 2) ? SolidViscosityModelType::viscType_Potential : SolidViscosityModelType::viscType_NA;
    dmn.solid_visc.mu = ctx->visc_mu;
    dmn.prop[PhysicalProperyType::solid_density] = par[4];
    dmn.prop[PhysicalProperyType::f_x] = par[5];
    dmn.prop[PhysicalProperyType::f_y] = par[6];
    dmn.prop[PhysicalProperyType::f_z] = par[7];
    dmn.prop[PhysicalProperyType::elasticity_modulus] = par[8];
    dmn.prop[PhysicalProperyType::poisson_ratio] = par[9];
    dmn.prop[PhysicalProperyType::ctau_M] = par[10];
    dmn.prop[PhysicalProperyType::ctau_C] = par[11];
    const int vol = int(par[12]);
    {
      const int iso = int(par[15]);
      dmn.stM.isoType = (iso == 3) ? ConstitutiveModelType::stIso_HO : (iso == 4) ? ConstitutiveModelType::stIso_MR
                      : (iso == 5) ? ConstitutiveModelType::stIso_HGO : (iso == 6) ? ConstitutiveModelType::stIso_Gucci
                      : (iso == 7) ? ConstitutiveModelType::stIso_HO_ma : ConstitutiveModelType::stIso_nHook;
    }
    dmn.stM.volType = (vol == 1) ? ConstitutiveModelType::stVol_Quad
                    : (vol == 2) ? ConstitutiveModelType::stVol_ST91
                    : (vol == 3) ? ConstitutiveModelType::stVol_M94 : ConstitutiveModelType::stIso_NA;
    dmn.stM.C10 = par[13];
    dmn.stM.Kpen = par[14];
    dmn.stM.a = par[16]; dmn.stM.b = par[17]; dmn.stM.aff = par[18]; dmn.stM.bff = par[19];
    dmn.stM.ass = par[20]; dmn.stM.bss = par[21]; dmn.stM.afs = par[22]; dmn.stM.bfs = par[23]; dmn.stM.khs = par[24];
    dmn.stM.Tf.fType = (par[25] != 0.0) ? utils::ibset(0, iBC_std) : 0;        // par[25] = Tf.g, par[26] = Tf.eta_s
    dmn.stM.Tf.g = par[25];
    dmn.stM.Tf.eta_s = par[26];
    dmn.stM.C01 = par[27];                        // Mooney-Rivlin
    dmn.stM.kap = par[28];                        // HGO
    com_mod.Bf.resize(3, nNo);
    std::memcpy(com_mod.Bf.data(), Bf, sizeof(double)*3*size_t(nNo));
    if (!eq.linear_algebra) eq.linear_algebra = new FsilsLinearAlgebra();

    Array<double> Ag_a(tDof, nNo), Yg_a(tDof, nNo), Dg_a(tDof, nNo);
    std::memcpy(Ag_a.data(), Ag, sizeof(double)*size_t(tDof)*nNo);
    std::memcpy(Yg_a.data(), Yg, sizeof(double)*size_t(tDof)*nNo);
    std::memcpy(Dg_a.data(), Dg, sizeof(double)*size_t(tDof)*nNo);

    com_mod.R.resize(dof, nNo);
    eq.linear_algebra->alloc(com_mod, eq);
    com_mod.Kd.resize(12, ctx->nnz);        // S/initialize.cpp:595; zeroed every Newton iteration (S/main.cpp:412)
    double t0 = now_s();
    ustruct::construct_usolid(com_mod, ctx->sim->cep_mod, com_mod.msh[0], Ag_a, Yg_a, Dg_a);
    double t1 = now_s();
    if (Ad) {
      com_mod.Ad.resize(3, nNo);
      std::memcpy(com_mod.Ad.data(), Ad, sizeof(double)*3*size_t(nNo));
      com_mod.Rd.resize(3, nNo);
      ustruct::ustruct_r(com_mod, Yg_a);
    }
    std::memcpy(R, com_mod.R.data(), sizeof(double)*size_t(dof)*nNo);
    std::memcpy(Val, com_mod.Val.data(), sizeof(double)*size_t(dof)*dof*ctx->nnz);
    std::memcpy(Kd, com_mod.Kd.data(), sizeof(double)*size_t(12)*ctx->nnz);
    return t1 - t0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1.0;
  }
}

// ----------------------------------------------------------------------------------------------
// FSILS: multi-"rank" (thread per rank) lhs_create + bc_create + solve / spmv.
// ----------------------------------------------------------------------------------------------
struct RefRank {
  // inputs
  int gnNo = 0, nNo = 0, nnz = 0;
  std::vector<int> gNodes, rowPtr, colPtr;
  struct Face { int nNo, dof, bGrp; std::vector<int> gNodes; std::vector<double> val; };
  std::vector<Face> faces;
  // built
  FSILS_lhsType lhs;
  FSILS_commuType commu;
  bool built = false;
  std::string err;
};

void* ref_rank_create(int gnNo, int nNo, int nnz, const int* gNodes, const int* rowPtr, const int* colPtr)
{
  auto r = new RefRank;
  r->gnNo = gnNo; r->nNo = nNo; r->nnz = nnz;
  r->gNodes.assign(gNodes, gNodes + nNo);
  r->rowPtr.assign(rowPtr, rowPtr + nNo + 1);
  r->colPtr.assign(colPtr, colPtr + nnz);
  return r;
}

void ref_rank_destroy(void* h) { delete static_cast<RefRank*>(h); }

// bGrp: 0 = Dirichlet, 1 = Neumann (L/fils_struct.hpp:58).  val: dof x nNo or NULL.
void ref_rank_add_face(void* h, int nNo, int dof, int bGrp, const int* gNodes, const double* val)
{
  auto r = static_cast<RefRank*>(h);
  RefRank::Face f;
  f.nNo = nNo; f.dof = dof; f.bGrp = bGrp;
  f.gNodes.assign(gNodes, gNodes + nNo);
  if (val) f.val.assign(val, val + size_t(dof)*nNo);
  r->faces.push_back(std::move(f));
}

static void run_ranks(int nranks, void** hs, const std::function<void(int, RefRank&)>& fn)
{
  mpistub_set_world(nranks);
  std::vector<std::thread> th;
  for (int p = 0; p < nranks; p++) {
    th.emplace_back([&, p]() {
      mpistub_bind_rank(p);
      auto& r = *static_cast<RefRank*>(hs[p]);
      try { fn(p, r); } catch (const std::exception& e) { r.err = e.what(); }
    });
  }
  for (auto& t : th) t.join();
}

static void build_rank(RefRank& r)
{
  if (r.built) return;
  fsils_commu_create(r.commu, MPI_COMM_WORLD);
  Vector<int> gNodes(r.nNo), rowPtr(r.nNo+1), colPtr(r.nnz);
  std::memcpy(gNodes.data(), r.gNodes.data(), sizeof(int)*r.nNo);
  std::memcpy(rowPtr.data(), r.rowPtr.data(), sizeof(int)*(r.nNo+1));
  std::memcpy(colPtr.data(), r.colPtr.data(), sizeof(int)*r.nnz);
  int nFaces = int(r.faces.size());
  fsils_lhs_create(r.lhs, r.commu, r.gnNo, r.nNo, r.nnz, gNodes, rowPtr, colPtr, nFaces);
  for (int i = 0; i < nFaces; i++) {
    auto& f = r.faces[i];
    Vector<int> gN(f.nNo);
    std::memcpy(gN.data(), f.gNodes.data(), sizeof(int)*f.nNo);
    BcType bt = f.bGrp == 0 ? BcType::BC_TYPE_Dir : BcType::BC_TYPE_Neu;
    if (!f.val.empty()) {
      Array<double> v(f.dof, f.nNo);
      std::memcpy(v.data(), f.val.data(), sizeof(double)*f.val.size());
      fsils_bc_create(r.lhs, i, f.nNo, f.dof, bt, gN, v);
    } else {
      fsils_bc_create(r.lhs, i, f.nNo, f.dof, bt, gN);
    }
  }
  r.built = true;
}

// Build lhs (+faces) on all ranks.  Returns 0 on success.
int ref_ranks_build(int nranks, void** hs)
{
  run_ranks(nranks, hs, [](int, RefRank& r) { build_rank(r); });
  for (int p = 0; p < nranks; p++) {
    auto& r = *static_cast<RefRank*>(hs[p]);
    if (!r.err.empty()) { g_err = r.err; return 1; }
  }
  return 0;
}

// Introspection of the reference's lhs (to pin the product's own lhs construction).
// info = {mynNo, shnNo, nReq}
void ref_rank_get_info(void* h, int* info, int* map, int* rowPtr2, int* colPtr, int* diagPtr)
{
  auto& lhs = static_cast<RefRank*>(h)->lhs;
  info[0] = lhs.mynNo; info[1] = lhs.shnNo; info[2] = lhs.nReq;
  if (map) std::memcpy(map, lhs.map.data(), sizeof(int)*lhs.nNo);
  if (rowPtr2) std::memcpy(rowPtr2, lhs.rowPtr.data(), sizeof(int)*2*lhs.nNo);
  if (colPtr) std::memcpy(colPtr, lhs.colPtr.data(), sizeof(int)*lhs.nnz);
  if (diagPtr) std::memcpy(diagPtr, lhs.diagPtr.data(), sizeof(int)*lhs.nNo);
}

// i-th communication request: returns n, fills iP; ptr may be NULL to query n first.
int ref_rank_get_req(void* h, int i, int* iP, int* ptr)
{
  auto& cs = static_cast<RefRank*>(h)->lhs.cS[i];
  *iP = cs.iP;
  if (ptr) std::memcpy(ptr, cs.ptr.data(), sizeof(int)*cs.n);
  return cs.n;
}

// ls = {LS_type(795..798), RI.relTol, RI.absTol, RI.mItr, RI.sD, GM.relTol, GM.absTol, GM.mItr, GM.sD,
//       CG.relTol, CG.absTol, CG.mItr}   (doubles)
// prec: 701 = PREC_FSILS, 709 = PREC_RCS (S/consts.h:426)
// Per rank p: Ri[p] (dof x nNo, in: residual, out: solution), Val[p] (dof*dof x nnz, destroyed).
// incL/res: nFaces (same on all ranks), may be NULL when no faces.
// out[p*OUTN ..]: {RI.suc, RI.itr, RI.iNorm, RI.fNorm, RI.dB, RI.callD, GM.itr, CG.itr, Resm, Resc, GM.callD, CG.callD, wall_s}
int ref_ranks_solve(int nranks, void** hs, int dof, const double* ls, int prec, double** Ri, double** Val,
                    const int* incL, const double* res, double* out)
{
  run_ranks(nranks, hs, [&](int p, RefRank& r) {
    build_rank(r);
    FSILS_lsType L{};
    fsils_ls_create(L, static_cast<LinearSolverType>(int(ls[0])));
    L.RI.relTol = ls[1]; L.RI.absTol = ls[2]; L.RI.mItr = int(ls[3]); L.RI.sD = int(ls[4]);
    L.GM.relTol = ls[5]; L.GM.absTol = ls[6]; L.GM.mItr = int(ls[7]); L.GM.sD = int(ls[8]);
    L.CG.relTol = ls[9]; L.CG.absTol = ls[10]; L.CG.mItr = int(ls[11]);
    L.RI.itr = 0; L.GM.itr = 0; L.CG.itr = 0;
    L.RI.callD = 0; L.GM.callD = 0; L.CG.callD = 0; L.Resm = 0; L.Resc = 0;
    L.RI.iNorm = 0; L.RI.fNorm = 0; L.RI.dB = 0; L.RI.suc = false;
    int nFaces = int(r.faces.size());
    Array<double> Ri_a(dof, r.nNo), Val_a(dof*dof, r.nnz);
    std::memcpy(Ri_a.data(), Ri[p], sizeof(double)*size_t(dof)*r.nNo);
    std::memcpy(Val_a.data(), Val[p], sizeof(double)*size_t(dof)*dof*r.nnz);
    Vector<int> incL_v(nFaces);
    Vector<double> res_v(nFaces);
    for (int i = 0; i < nFaces; i++) { incL_v(i) = incL ? incL[i] : 1; res_v(i) = res ? res[i] : 0.0; }
    auto pt = static_cast<consts::PreconditionerType>(prec);
    double t0 = now_s();
    fsils_solve(r.lhs, L, dof, Ri_a, Val_a, pt, incL_v, res_v);
    double t1 = now_s();
    std::memcpy(Ri[p], Ri_a.data(), sizeof(double)*size_t(dof)*r.nNo);
    std::memcpy(Val[p], Val_a.data(), sizeof(double)*size_t(dof)*dof*r.nnz);
    double* o = out + 13*p;
    o[0] = L.RI.suc; o[1] = L.RI.itr; o[2] = L.RI.iNorm; o[3] = L.RI.fNorm; o[4] = L.RI.dB; o[5] = L.RI.callD;
    o[6] = L.GM.itr; o[7] = L.CG.itr; o[8] = L.Resm; o[9] = L.Resc; o[10] = L.GM.callD; o[11] = L.CG.callD;
    o[12] = t1 - t0;
  });
  for (int p = 0; p < nranks; p++) {
    auto& r = *static_cast<RefRank*>(hs[p]);
    if (!r.err.empty()) { g_err = r.err; r.err.clear(); return 1; }
  }
  return 0;
}

// y = K x (+ halo add) through fsils_spar_mul_vv on vectors given in the CALLER's node order
// (they are permuted by lhs.map on the way in and out, like fsils_solve does).  reps>1 repeats the
// product for timing; returns seconds per product on rank 0 in *secs.
int ref_ranks_spmv(int nranks, void** hs, int dof, double** Val, double** X, double** Y, int reps, double* secs)
{
  run_ranks(nranks, hs, [&](int p, RefRank& r) {
    build_rank(r);
    Array<double> K(dof*dof, r.nnz), U(dof, r.nNo), KU(dof, r.nNo);
    std::memcpy(K.data(), Val[p], sizeof(double)*size_t(dof)*dof*r.nnz);
    for (int a = 0; a < r.nNo; a++)
      for (int i = 0; i < dof; i++) U(i, r.lhs.map(a)) = X[p][i + size_t(a)*dof];
    double t0 = now_s();
    for (int k = 0; k < reps; k++)
      spar_mul::fsils_spar_mul_vv(r.lhs, r.lhs.rowPtr, r.lhs.colPtr, dof, K, U, KU);
    double t1 = now_s();
    for (int a = 0; a < r.nNo; a++)
      for (int i = 0; i < dof; i++) Y[p][i + size_t(a)*dof] = KU(i, r.lhs.map(a));
    if (p == 0 && secs) *secs = (t1 - t0) / reps;
  });
  for (int p = 0; p < nranks; p++) {
    auto& r = *static_cast<RefRank*>(hs[p]);
    if (!r.err.empty()) { g_err = r.err; r.err.clear(); return 1; }
  }
  return 0;
}

// In-place halo add of a dof x nNo vector given in the caller's node order (fsils_commuv).
int ref_ranks_commuv(int nranks, void** hs, int dof, double** V)
{
  run_ranks(nranks, hs, [&](int p, RefRank& r) {
    build_rank(r);
    Array<double> U(dof, r.nNo);
    for (int a = 0; a < r.nNo; a++)
      for (int i = 0; i < dof; i++) U(i, r.lhs.map(a)) = V[p][i + size_t(a)*dof];
    fsils_commuv(r.lhs, dof, U);
    for (int a = 0; a < r.nNo; a++)
      for (int i = 0; i < dof; i++) V[p][i + size_t(a)*dof] = U(i, r.lhs.map(a));
  });
  return 0;
}

#ifdef WITH_B200_DROPIN
// ----------------------------------------------------------------------------------------------
// Drop-in test: one Newton-iteration hot path driven the way the reference drives a LinearAlgebra
// plug-in (main.cpp:68-77, 480-600): the reference's own ComMod / eqType / FSILS_lhsType objects,
// ls_ns::ls_alloc -> global assembly -> ls_ns::ls_solve, with eq.linear_algebra = B200LinearAlgebra
// (svfsiplus_b200/host).  mode 0: host assembly by the reference's construct_fluid (do_assem) +
// device solve; mode 1: device assembly (assemble_mesh) + device solve.
// faces: nFaces x {nNo, dof, bGrp}; nodes/vals concatenated.  ls as in ref_ranks_solve.
// out = {RI.suc, RI.itr, RI.iNorm, RI.fNorm, GM.itr, CG.itr, Resm, Resc, used_device_assembly}
// ----------------------------------------------------------------------------------------------
} // extern "C"

namespace {
// Shared tail of the drop-in steps: what initialize() / fsi_ls_ini leave behind (com_mod.lhs + faces), the <LS> block,
// eq.linear_algebra = B200LinearAlgebra, then one Newton iteration (ls_alloc, global_eq_assem hook, ls_solve).
template <class HostAssembly>
void dropin_newton_iteration(AsmCtx* ctx, int dof, int mode, int tDof, const double* Ag, const double* Yg, const double* Dg,
                             int nFaces, const int* f_info, const int* f_nodes, const double* f_val,
                             const double* ls, const int* incL, const double* res, double* X, double* out,
                             HostAssembly&& host_assembly,
                             const std::function<void(B200LinearAlgebra*, Array<double>&)>& after_assembly = nullptr)
{
  using namespace consts;
  dropin_install_crash_handler();
  auto& com_mod = ctx->sim->com_mod;
  const int nNo = com_mod.tnNo;
  auto& eq = com_mod.eq[0];
  auto& lhs = com_mod.lhs;
  lhs = FSILS_lhsType();
  fsils_commu_create(lhs.commu, MPI_COMM_WORLD);
  Vector<int> gNodes(nNo);
  for (int a = 0; a < nNo; a++) gNodes(a) = a;
  fsils_lhs_create(lhs, lhs.commu, nNo, nNo, ctx->nnz, gNodes, com_mod.rowPtr, com_mod.colPtr, nFaces);
  size_t on = 0, ov = 0;
  for (int i = 0; i < nFaces; i++) {
    const int fn = f_info[3*i], fd = f_info[3*i+1], fb = f_info[3*i+2];
    Vector<int> gN(fn);
    std::memcpy(gN.data(), f_nodes + on, sizeof(int)*fn);
    Array<double> v(fd, fn);
    std::memcpy(v.data(), f_val + ov, sizeof(double)*size_t(fd)*fn);
    fsils_bc_create(lhs, i, fn, fd, fb == 0 ? BcType::BC_TYPE_Dir : BcType::BC_TYPE_Neu, gN, v);
    on += fn; ov += size_t(fd)*fn;
  }
  // read_files.cpp:2046-2075 + add_eq_linear_algebra (main.cpp:68-77)
  fsils_ls_create(eq.FSILS, static_cast<LinearSolverType>(int(ls[0])));
  eq.FSILS.RI.relTol = ls[1]; eq.FSILS.RI.absTol = ls[2]; eq.FSILS.RI.mItr = int(ls[3]); eq.FSILS.RI.sD = int(ls[4]);
  eq.FSILS.GM.relTol = ls[5]; eq.FSILS.GM.absTol = ls[6]; eq.FSILS.GM.mItr = int(ls[7]); eq.FSILS.GM.sD = int(ls[8]);
  eq.FSILS.CG.relTol = ls[9]; eq.FSILS.CG.absTol = ls[10]; eq.FSILS.CG.mItr = int(ls[11]);
  eq.linear_algebra_preconditioner = PreconditionerType::PREC_FSILS;
  delete eq.linear_algebra;
  auto* la = new B200LinearAlgebra();
  eq.linear_algebra = la;
  la->check_options(PreconditionerType::PREC_FSILS, mode == 1 ? B200_LINEAR_ALGEBRA_TYPE : LinearAlgebraType::fsils);
  la->set_preconditioner(eq.linear_algebra_preconditioner);
  la->initialize(com_mod, eq);
  la->set_assembly(mode == 1 ? B200_LINEAR_ALGEBRA_TYPE : LinearAlgebraType::fsils);

  Array<double> Ag_a(tDof, nNo), Yg_a(tDof, nNo), Dg_a(tDof, nNo);
  std::memcpy(Ag_a.data(), Ag, sizeof(double)*size_t(tDof)*nNo);
  std::memcpy(Yg_a.data(), Yg, sizeof(double)*size_t(tDof)*nNo);
  if (Dg) std::memcpy(Dg_a.data(), Dg, sizeof(double)*size_t(tDof)*nNo);

  // one Newton iteration (main.cpp iterate_solution): ls_alloc, global_eq_assem, ls_solve
  ls_ns::ls_alloc(com_mod, eq);
  bool on_device = la->assemble_mesh(com_mod, com_mod.msh[0], Ag_a, Yg_a, Dg_a, &ctx->sim->cep_mod);   // the global_eq_assem hook
  if (!on_device) host_assembly(Ag_a, Yg_a, Dg_a);
  if (after_assembly) after_assembly(la, Yg_a);        // set_bc_neu (main.cpp:478) with the set_bc_neu_l hook
  Vector<int> incL_v(nFaces);
  Vector<double> res_v(nFaces);
  for (int i = 0; i < nFaces; i++) { incL_v(i) = incL ? incL[i] : 1; res_v(i) = res ? res[i] : 0.0; }
  ls_ns::ls_solve(com_mod, eq, incL_v, res_v);

  std::memcpy(X, com_mod.R.data(), sizeof(double)*size_t(dof)*nNo);
  auto& L = eq.FSILS;
  out[0] = L.RI.suc; out[1] = L.RI.itr; out[2] = L.RI.iNorm; out[3] = L.RI.fNorm;
  out[4] = L.GM.itr; out[5] = L.CG.itr; out[6] = L.Resm; out[7] = L.Resc; out[8] = on_device ? 1.0 : 0.0;
  delete eq.linear_algebra;
  eq.linear_algebra = nullptr;
}
} // namespace

extern "C" {

int ref_dropin_fluid_step(void* h, int mode, int tDof, double dt, double am, double af, double gam, double rho,
                          const double* f, double Kinv_darcy, const double* visc,
                          const double* Ag, const double* Yg, const double* Bf,
                          int nFaces, const int* f_info, const int* f_nodes, const double* f_val,
                          const double* ls, const int* incL, const double* res, double* X, double* out)
{
  try {
    mpistub_set_world(1);
    mpistub_bind_rank(0);
    auto ctx = static_cast<AsmCtx*>(h);
    auto& com_mod = ctx->sim->com_mod;
    configure_fluid(ctx, tDof, 0, dt, am, af, gam, rho, f, Kinv_darcy, visc, Bf);
    dropin_newton_iteration(ctx, 4, mode, tDof, Ag, Yg, nullptr, nFaces, f_info, f_nodes, f_val, ls, incL, res, X, out,
                            [&](Array<double>& A, Array<double>& Y, Array<double>&) { fluid::construct_fluid(com_mod, com_mod.msh[0], A, Y); });
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Fluid step with one Neumann face (lFa.IEN, lFa.gE, nodal values hg) assembled after the volume: through
// B200LinearAlgebra::assemble_face when it takes the face, else through the reference's b_assem_neu_bc + assemble().
// out[8]: device volume assembly, out[9] (when out has 10 entries): 1 if the face went to the device.
int ref_dropin_fluid_face_step(void* h, int mode, int tDof, double dt, double am, double af, double gam, double rho, double bfs,
                               const double* visc, const double* Ag, const double* Yg, const double* Bf,
                               int nFaces, const int* f_info, const int* f_nodes, const double* f_val,
                               int eNoNb, int nElb, const int* IENb, const int* gE, const double* hg,
                               const double* ls, const int* incL, const double* res, double* X, double* out)
{
  try {
    mpistub_set_world(1);
    mpistub_bind_rank(0);
    auto ctx = static_cast<AsmCtx*>(h);
    auto& com_mod = ctx->sim->com_mod;
    const double f[3] = {0.0, 0.0, 0.0};
    configure_fluid(ctx, tDof, 0, dt, am, af, gam, rho, f, 0.0, visc, Bf);
    com_mod.eq[0].dmn[0].prop[consts::PhysicalProperyType::backflow_stab] = bfs;
    auto& msh = com_mod.msh[0];
    msh.nFa = 1;
    msh.fa.resize(1);
    auto& fa = msh.fa[0];
    fa.name = "face"; fa.iM = 0; fa.eNoN = eNoNb; fa.nEl = nElb;
    fa.IEN.resize(eNoNb, nElb);
    std::memcpy(fa.IEN.data(), IENb, sizeof(int)*size_t(eNoNb)*nElb);
    fa.gE.resize(nElb);
    std::memcpy(fa.gE.data(), gE, sizeof(int)*size_t(nElb));
    nn::select_eleb(ctx->sim.get(), msh, fa);
    Vector<double> hg_v(com_mod.tnNo);
    std::memcpy(hg_v.data(), hg, sizeof(double)*size_t(com_mod.tnNo));
    double face_on_device = 0.0;
    dropin_newton_iteration(ctx, 4, mode, tDof, Ag, Yg, nullptr, nFaces, f_info, f_nodes, f_val, ls, incL, res, X, out,
                            [&](Array<double>& A, Array<double>& Y, Array<double>&) { fluid::construct_fluid(com_mod, com_mod.msh[0], A, Y); },
                            [&](B200LinearAlgebra* la, Array<double>& Y) {
                              if (la->assemble_face(com_mod, fa, hg_v, Y)) face_on_device = 1.0;
                              else eq_assem::b_assem_neu_bc(com_mod, fa, hg_v, Y);
                            });
    out[9] = face_on_device;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Same for a struct (kind 0) / lElas (1) / mesh (2) equation; par, s, Do as in ref_asm_solid.
int ref_dropin_solid_step(void* h, int mode, int kind, int tDof, int s, const double* par,
                          const double* Ag, const double* Yg, const double* Dg, const double* Do, const double* Bf,
                          int nFaces, const int* f_info, const int* f_nodes, const double* f_val,
                          const double* ls, const int* incL, const double* res, double* X, double* out)
{
  try {
    mpistub_set_world(1);
    mpistub_bind_rank(0);
    auto ctx = static_cast<AsmCtx*>(h);
    auto& com_mod = ctx->sim->com_mod;
    configure_solid(ctx, kind, tDof, s, par, Do, Bf);
    dropin_newton_iteration(ctx, 3, mode, tDof, Ag, Yg, Dg, nFaces, f_info, f_nodes, f_val, ls, incL, res, X, out,
                            [&](Array<double>& A, Array<double>& Y, Array<double>& D) {
                              if (kind == 0) struct_ns::construct_dsolid(com_mod, ctx->sim->cep_mod, com_mod.msh[0], A, Y, D);
                              else if (kind == 1) l_elas::construct_l_elas(com_mod, com_mod.msh[0], A, D);
                              else mesh::construct_mesh(com_mod, ctx->sim->cep_mod, com_mod.msh[0], A, D);
                            });
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}
#endif

} // extern "C"

// ----------------------------------------------------------------------------------------------
// Generalised-alpha time integrator: the reference's own pic::picp / pici / picc on caller arrays.
// ----------------------------------------------------------------------------------------------
extern "C" {

// op: 0 picp, 1 pici, 2 picc (for equation cEq).  eqpar[8*i..] = {s, e, am, af, gam, beta, phys, itr} of equation i with
// phys 0 fluid, 1 struct, 2 lElas, 3 ustruct, 5 mesh.  State arrays are (tDof, tnNo), Ad and Rd (3, tnNo), R (dof, tnNo).
int ref_pic(int op, int tnNo, int tDof, int nEq, int cEq, const double* eqpar, int dFlag, int sstEq, double dt,
            double* Ao, double* Yo, double* Do, double* An, double* Yn, double* Dn, double* Ad, double* Ag, double* Yg, double* Dg,
            const double* R, const double* Rd)
{
  try {
    using namespace consts;
    mpistub_set_world(1);
    mpistub_bind_rank(0);
    Simulation sim;
    auto& com_mod = sim.com_mod;
    com_mod.nsd = 3;
    com_mod.tnNo = tnNo;
    com_mod.tDof = tDof;
    com_mod.dt = dt;
    com_mod.nMsh = 0;
    com_mod.nEq = nEq;
    com_mod.cEq = cEq;
    com_mod.dFlag = (dFlag != 0);
    com_mod.sstEq = (sstEq != 0);
    com_mod.pstEq = false;
    com_mod.eq.resize(nEq);
    const EquationType phys_of[6] = {EquationType::phys_fluid, EquationType::phys_struct, EquationType::phys_lElas,
                                     EquationType::phys_ustruct, EquationType::phys_FSI, EquationType::phys_mesh};
    for (int i = 0; i < nEq; i++) {
      auto& eq = com_mod.eq[i];
      const double* q = eqpar + 8*i;
      eq.s = int(q[0]); eq.e = int(q[1]); eq.am = q[2]; eq.af = q[3]; eq.gam = q[4]; eq.beta = q[5];
      eq.phys = phys_of[int(q[6])];
      eq.itr = int(q[7]);
      eq.dof = eq.e - eq.s + 1;
      eq.coupled = true; eq.ok = false; eq.minItr = 1; eq.maxItr = 100; eq.tol = 1e-30; eq.iNorm = 1.0; eq.pNorm = 1.0;
      eq.FSILS.RI.iNorm = 1.0;
    }
    auto load = [&](Array<double>& A, const double* src, int nr) { A.resize(nr, tnNo); if (src) std::memcpy(A.data(), src, sizeof(double)*size_t(nr)*tnNo); };
    load(com_mod.Ao, Ao, tDof); load(com_mod.Yo, Yo, tDof); load(com_mod.Do, Do, tDof);
    load(com_mod.An, An, tDof); load(com_mod.Yn, Yn, tDof); load(com_mod.Dn, Dn, tDof);
    load(com_mod.Ad, Ad, 3);
    Array<double> Ag_a, Yg_a, Dg_a;
    load(Ag_a, Ag, tDof); load(Yg_a, Yg, tDof); load(Dg_a, Dg, tDof);
    if (op == 2) {
      const int dof = com_mod.eq[cEq].dof;
      com_mod.dof = dof;
      load(com_mod.R, R, dof);
      load(com_mod.Rd, Rd, 3);
    }
    if (op == 0) pic::picp(&sim);
    else if (op == 1) pic::pici(&sim, Ag_a, Yg_a, Dg_a);
    else pic::picc(&sim);
    auto store = [&](double* dst, const Array<double>& A) { if (dst) std::memcpy(dst, A.data(), sizeof(double)*A.size()); };
    store(An, com_mod.An); store(Yn, com_mod.Yn); store(Dn, com_mod.Dn); store(Ad, com_mod.Ad);
    store(Ag, Ag_a); store(Yg, Yg_a); store(Dg, Dg_a);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

} // extern "C"

// ----------------------------------------------------------------------------------------------
// Boundary-face (Neumann) assembly: the reference's b_assem_neu_bc (+ gnnb, b_fluid / b_l_elas) on one face.
// ----------------------------------------------------------------------------------------------
extern "C" {

// all_fun::integ over one face (S/all_fun.cpp:561,724,858): rows l..u of s(nrows,nNo).  s == NULL: integrand 1 (area).
// geo: 0 reference, 1 old time step (x + Do(0:2)), 2 new (x + Dn(0:2)), 3 moving mesh (x + Do(4:6)); D(tDofD,nNo).
int ref_face_integ(void* h, int eNoNb, int nElb, const int* IENb, const int* gE, int nrows, const double* s, int l, int u,
                   int geo, int tDofD, const double* D, double* result)
{
  try {
    using namespace consts;
    auto ctx = static_cast<AsmCtx*>(h);
    auto& com_mod = ctx->sim->com_mod;
    const int nNo = com_mod.tnNo;
    auto& msh = com_mod.msh[0];
    msh.nFa = 1;
    msh.fa.resize(1);
    auto& fa = msh.fa[0];
    fa.name = "face"; fa.iM = 0; fa.eNoN = eNoNb; fa.nEl = nElb;
    fa.IEN.resize(eNoNb, nElb);
    std::memcpy(fa.IEN.data(), IENb, sizeof(int)*size_t(eNoNb)*nElb);
    fa.gE.resize(nElb);
    std::memcpy(fa.gE.data(), gE, sizeof(int)*size_t(nElb));
    nn::select_eleb(ctx->sim.get(), msh, fa);
    fs::init_fs_face(com_mod, msh, fa);
    com_mod.mvMsh = (geo == 3);
    auto cfg = MechanicalConfigurationType::reference;
    if (geo == 1 || geo == 3) {
      com_mod.Do.resize(tDofD, nNo);
      std::memcpy(com_mod.Do.data(), D, sizeof(double)*size_t(tDofD)*nNo);
      if (geo == 1) cfg = MechanicalConfigurationType::old_timestep;
    } else if (geo == 2) {
      com_mod.Dn.resize(tDofD, nNo);
      std::memcpy(com_mod.Dn.data(), D, sizeof(double)*size_t(tDofD)*nNo);
      cfg = MechanicalConfigurationType::new_timestep;
    }
    if (!s) {
      Vector<double> one(nNo);
      one = 1.0;
      *result = all_fun::integ(com_mod, ctx->sim->cm_mod, fa, one, false, cfg);
    } else {
      Array<double> sa(nrows, nNo);
      std::memcpy(sa.data(), s, sizeof(double)*size_t(nrows)*nNo);
      *result = all_fun::integ(com_mod, ctx->sim->cm_mod, fa, sa, l, std::optional<int>(u), false, cfg);
    }
    com_mod.mvMsh = false;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// eq_assem::fsi_ls_upd (S/eq_assem.cpp:316) + fsils_bc_update: the face vector val(3,nNoFace) = int N_a n dGamma of a Neumann
// face on the new-time-step configuration x + Dn(0:2) (mvMsh = 0) or the moving mesh x + Do(4:6) (mvMsh = 1).
// gN(nNoFace): the face's node list (lFa.gN).  The lhs is built here (1 rank) with that one face.
int ref_fsi_ls_upd(void* h, int eNoNb, int nElb, const int* IENb, const int* gE, int nNoFace, const int* gN, int mvMsh, int tDofD,
                   const double* D, double* val)
{
  try {
    using namespace consts;
    mpistub_set_world(1);
    mpistub_bind_rank(0);
    auto ctx = static_cast<AsmCtx*>(h);
    auto& com_mod = ctx->sim->com_mod;
    const int nNo = com_mod.tnNo;
    auto& msh = com_mod.msh[0];
    msh.nFa = 1;
    msh.fa.resize(1);
    auto& fa = msh.fa[0];
    fa.name = "face"; fa.iM = 0; fa.eNoN = eNoNb; fa.nEl = nElb; fa.nNo = nNoFace;
    fa.IEN.resize(eNoNb, nElb);
    std::memcpy(fa.IEN.data(), IENb, sizeof(int)*size_t(eNoNb)*nElb);
    fa.gE.resize(nElb);
    std::memcpy(fa.gE.data(), gE, sizeof(int)*size_t(nElb));
    fa.gN.resize(nNoFace);
    std::memcpy(fa.gN.data(), gN, sizeof(int)*size_t(nNoFace));
    nn::select_eleb(ctx->sim.get(), msh, fa);
    com_mod.mvMsh = (mvMsh != 0);
    if (mvMsh) { com_mod.Do.resize(tDofD, nNo); std::memcpy(com_mod.Do.data(), D, sizeof(double)*size_t(tDofD)*nNo); }
    else { com_mod.Dn.resize(tDofD, nNo); std::memcpy(com_mod.Dn.data(), D, sizeof(double)*size_t(tDofD)*nNo); }
    auto& lhs = com_mod.lhs;
    lhs = FSILS_lhsType();
    fsils_commu_create(lhs.commu, MPI_COMM_WORLD);
    Vector<int> gNodes(nNo);
    for (int a = 0; a < nNo; a++) gNodes(a) = a;
    fsils_lhs_create(lhs, lhs.commu, nNo, nNo, ctx->nnz, gNodes, com_mod.rowPtr, com_mod.colPtr, 1);
    Array<double> v0(3, nNoFace);
    fsils_bc_create(lhs, 0, nNoFace, 3, BcType::BC_TYPE_Neu, fa.gN, v0);
    bcType lBc;
    lBc.lsPtr = 0;
    eq_assem::fsi_ls_upd(com_mod, lBc, fa);
    std::memcpy(val, lhs.face[0].val.data(), sizeof(double)*size_t(3)*nNoFace);
    com_mod.mvMsh = false;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// mat_models_carray::get_pk2cc<3> (S/mat_models_carray.h:182) at one deformation gradient.  par as in ref_asm_solid (28
// entries); F(3,3) row-major; fl = fibre and sheet directions (6).  Outputs S(3,3) and Dm(6,6) row-major.
int ref_pk2cc(void* h, const double* par, const double* F9, const double* fl6, double* S9, double* Dm36)
{
  try {
    auto ctx = static_cast<AsmCtx*>(h);
    auto& com_mod = ctx->sim->com_mod;
    std::vector<double> Bf(size_t(3)*com_mod.tnNo, 0.0);
    configure_solid(ctx, 0, 3, 0, par, nullptr, Bf.data());
    double F[3][3], S[3][3], Dm[6][6];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) F[i][j] = F9[i*3 + j];
    Array<double> fl(3, 2);
    for (int i = 0; i < 3; i++) { fl(i, 0) = fl6[i]; fl(i, 1) = fl6[3 + i]; }
    mat_fun_carray::ten_init(3);                 // struct_3d_carray does this before every call (S/sv_struct.cpp:654)
    mat_models_carray::get_pk2cc<3>(com_mod, ctx->sim->cep_mod, com_mod.eq[0].dmn[0], F, 2, fl, 0.0, S, Dm);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) S9[i*3 + j] = S[i][j];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Dm36[i*6 + j] = Dm[i][j];
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// mat_models::get_pk2cc_dev (S/mat_models.cpp:630; the ustruct form: deviatoric S, isochoric Dm) at one deformation gradient.
// Arguments as ref_pk2cc.
int ref_pk2cc_dev(void* h, const double* par, const double* F9, const double* fl6, double* S9, double* Dm36)
{
  try {
    auto ctx = static_cast<AsmCtx*>(h);
    auto& com_mod = ctx->sim->com_mod;
    std::vector<double> Bf(size_t(3)*com_mod.tnNo, 0.0);
    configure_solid(ctx, 0, 3, 0, par, nullptr, Bf.data());
    Array<double> F(3, 3), S(3, 3), Dm(6, 6), fl(3, 2);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) F(i, j) = F9[i*3 + j];
    for (int i = 0; i < 3; i++) { fl(i, 0) = fl6[i]; fl(i, 1) = fl6[3 + i]; }
    double Ja = 0.0;
    mat_fun::ten_init(3);
    mat_models::get_pk2cc_dev(com_mod, ctx->sim->cep_mod, com_mod.eq[0].dmn[0], F, 2, fl, 0.0, S, Dm, Ja);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) S9[i*3 + j] = S(i, j);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Dm36[i*6 + j] = Dm(i, j);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Follower pressure load on one face of a struct or ustruct equation: eq_assem::b_neu_folw_p (S/eq_assem.cpp:186; get_nnx,
// gnn, gnnb, struct_ns::b_struct_3d / ustruct::b_ustruct_3d + ustruct_do_assem).  par = {dt, af, beta, tDof, ustruct, am, gam}.
// Outputs the face's contribution alone: struct R(3,nNo), Val(9,nnz); ustruct R(4,nNo), Val(16,nnz), Kd(12,nnz).
int ref_asm_bfolw(void* h, int eNoNb, int nElb, const int* IENb, const int* gE, const double* par, const double* hg,
                  const double* Dg, double* R, double* Val, double* Kd)
{
  try {
    using namespace consts;
    auto ctx = static_cast<AsmCtx*>(h);
    auto& com_mod = ctx->sim->com_mod;
    const int nNo = com_mod.tnNo;
    const int tDof = int(par[3]);
    const bool us = par[4] != 0.0;
    const int dof = us ? 4 : 3;
    const auto phys = us ? EquationType::phys_ustruct : EquationType::phys_struct;
    com_mod.tDof = tDof; com_mod.dof = dof; com_mod.dt = par[0]; com_mod.mvMsh = false;
    com_mod.cEq = 0; com_mod.nEq = 1;
    if (com_mod.eq.size() != 1) com_mod.eq.resize(1);
    auto& eq = com_mod.eq[0];
    eq.phys = phys; eq.dof = dof; eq.s = 0; eq.e = dof - 1; eq.af = par[1]; eq.beta = par[2]; eq.am = par[5]; eq.gam = par[6];
    eq.nDmn = 1;
    if (eq.dmn.size() != 1) eq.dmn.resize(1);
    eq.dmn[0].Id = -1;
    eq.dmn[0].phys = phys;
    if (us) { com_mod.Kd.resize(12, ctx->nnz); com_mod.Kd = 0.0; }
    if (!eq.linear_algebra) eq.linear_algebra = new FsilsLinearAlgebra();
    auto& msh = com_mod.msh[0];
    msh.nFa = 1;
    msh.fa.resize(1);
    auto& fa = msh.fa[0];
    fa.name = "face"; fa.iM = 0; fa.eNoN = eNoNb; fa.nEl = nElb;
    fa.IEN.resize(eNoNb, nElb);
    std::memcpy(fa.IEN.data(), IENb, sizeof(int)*size_t(eNoNb)*nElb);
    fa.gE.resize(nElb);
    std::memcpy(fa.gE.data(), gE, sizeof(int)*size_t(nElb));
    nn::select_eleb(ctx->sim.get(), msh, fa);
    Vector<double> hg_v(nNo);
    std::memcpy(hg_v.data(), hg, sizeof(double)*size_t(nNo));
    Array<double> Dg_a(tDof, nNo);
    std::memcpy(Dg_a.data(), Dg, sizeof(double)*size_t(tDof)*nNo);
    com_mod.R.resize(dof, nNo);
    eq.linear_algebra->alloc(com_mod, eq);
    bcType lBc;
    lBc.flwP = true;
    eq_assem::b_neu_folw_p(com_mod, lBc, fa, hg_v, Dg_a);
    std::memcpy(R, com_mod.R.data(), sizeof(double)*size_t(dof)*nNo);
    std::memcpy(Val, com_mod.Val.data(), sizeof(double)*size_t(dof)*dof*ctx->nnz);
    if (us && Kd) std::memcpy(Kd, com_mod.Kd.data(), sizeof(double)*size_t(12)*ctx->nnz);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// kind 0: fluid equation (dof 4, b_fluid), 1: struct equation (dof 3, b_l_elas).
// par = {dt, af, gam, rho, bfs, tDof, mvMsh}.  IENb(eNoNb,nElb), gE(nElb); hg(nNo); Yg, Do (tDof,nNo; Do may be NULL).
// Outputs the face's contribution alone: R (dof,nNo), Val (dof*dof,nnz).
int ref_asm_bneu(void* h, int kind, int eNoNb, int nElb, const int* IENb, const int* gE, const double* par, const double* hg,
                 const double* Yg, const double* Do, double* R, double* Val)
{
  try {
    using namespace consts;
    auto ctx = static_cast<AsmCtx*>(h);
    auto& com_mod = ctx->sim->com_mod;
    const int nNo = com_mod.tnNo;
    const int tDof = int(par[5]);
    const int dof = (kind == 0) ? 4 : 3;
    if (kind == 0) {
      const double f[3] = {0.0, 0.0, 0.0};
      const double visc[6] = {0.0, 0.04, 0.0, 0.0, 0.0, 0.0};
      std::vector<double> Bf(size_t(3)*nNo, 0.0);
      configure_fluid(ctx, tDof, int(par[6]), par[0], 1.0, par[1], par[2], par[3], f, 0.0, visc, Bf.data());
      com_mod.eq[0].dmn[0].prop[PhysicalProperyType::backflow_stab] = par[4];
    } else {
      com_mod.tDof = tDof; com_mod.dof = dof; com_mod.dt = par[0]; com_mod.mvMsh = false;
      com_mod.cEq = 0; com_mod.nEq = 1;
      if (com_mod.eq.size() != 1) com_mod.eq.resize(1);
      auto& eq = com_mod.eq[0];
      eq.phys = EquationType::phys_struct; eq.dof = dof; eq.s = 0; eq.e = dof - 1; eq.af = par[1]; eq.gam = par[2];
      eq.nDmn = 1;
      if (eq.dmn.size() != 1) eq.dmn.resize(1);
      eq.dmn[0].Id = -1;
      eq.dmn[0].phys = EquationType::phys_struct;
    }
    auto& eq = com_mod.eq[0];
    if (!eq.linear_algebra) eq.linear_algebra = new FsilsLinearAlgebra();
    auto& msh = com_mod.msh[0];
    msh.nFa = 1;
    msh.fa.resize(1);
    auto& fa = msh.fa[0];
    fa.name = "face";
    fa.iM = 0;
    fa.eNoN = eNoNb;
    fa.nEl = nElb;
    fa.IEN.resize(eNoNb, nElb);
    std::memcpy(fa.IEN.data(), IENb, sizeof(int)*size_t(eNoNb)*nElb);
    fa.gE.resize(nElb);
    std::memcpy(fa.gE.data(), gE, sizeof(int)*size_t(nElb));
    nn::select_eleb(ctx->sim.get(), msh, fa);
    if (Do) {
      com_mod.Do.resize(tDof, nNo);
      std::memcpy(com_mod.Do.data(), Do, sizeof(double)*size_t(tDof)*nNo);
    }
    Vector<double> hg_v(nNo);
    std::memcpy(hg_v.data(), hg, sizeof(double)*size_t(nNo));
    Array<double> Yg_a(tDof, nNo);
    std::memcpy(Yg_a.data(), Yg, sizeof(double)*size_t(tDof)*nNo);
    com_mod.R.resize(dof, nNo);
    eq.linear_algebra->alloc(com_mod, eq);
    eq_assem::b_assem_neu_bc(com_mod, fa, hg_v, Yg_a);
    std::memcpy(R, com_mod.R.data(), sizeof(double)*size_t(dof)*nNo);
    std::memcpy(Val, com_mod.Val.data(), sizeof(double)*size_t(dof)*dof*ctx->nnz);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// output::write_restart (S/output.cpp:202-345) on a bare Simulation: one rank, file <stem>_NNN.bin (+ the hard link
// <stem>_last.bin the reference makes).  Arrays as the solver holds them (column-major (tDof, tnNo)).  Dn / Ad / pS0 may be null.
int ref_io_write_restart(const char* stem, const int* stamp7, int cTS, double time, int nEq, const double* iNorm, int nXn, const double* xn,
                         int tDof, int tnNo, const double* Yn, const double* An, const double* Dn, int nsd, const double* Ad,
                         int nsymd, const double* pS0, long long* recLn_out)
{
  try {
    mpistub_set_world(1);
    mpistub_bind_rank(0);
    Simulation sim;
    auto& com_mod = sim.com_mod;
    com_mod.stFileName = stem;
    com_mod.stFileRepl = false;
    for (int i = 0; i < 7; i++) com_mod.stamp[i] = stamp7[i];
    com_mod.cTS = cTS;
    com_mod.time = time;
    com_mod.nEq = nEq;
    com_mod.eq.resize(nEq);
    for (int i = 0; i < nEq; i++) com_mod.eq[i].iNorm = iNorm[i];
    com_mod.cplBC.nX = nXn;
    com_mod.cplBC.xn.resize(nXn);
    std::memcpy(com_mod.cplBC.xn.data(), xn, sizeof(double)*size_t(nXn));
    com_mod.tDof = tDof;
    com_mod.tnNo = tnNo;
    com_mod.nsd = 3;
    com_mod.nsymd = 6;
    com_mod.Yn.resize(tDof, tnNo);
    com_mod.An.resize(tDof, tnNo);
    std::memcpy(com_mod.Yn.data(), Yn, sizeof(double)*size_t(tDof)*tnNo);
    std::memcpy(com_mod.An.data(), An, sizeof(double)*size_t(tDof)*tnNo);
    com_mod.ibFlag = false;
    com_mod.dFlag = Dn != nullptr;
    com_mod.sstEq = Ad != nullptr;
    com_mod.pstEq = pS0 != nullptr;
    sim.cep_mod.cepEq = false;
    if (Dn) { com_mod.Dn.resize(tDof, tnNo); std::memcpy(com_mod.Dn.data(), Dn, sizeof(double)*size_t(tDof)*tnNo); }
    if (Ad) { com_mod.Ad.resize(nsd, tnNo); std::memcpy(com_mod.Ad.data(), Ad, sizeof(double)*size_t(nsd)*tnNo); }
    if (pS0) { com_mod.pS0.resize(nsymd, tnNo); std::memcpy(com_mod.pS0.data(), pS0, sizeof(double)*size_t(nsymd)*tnNo); }
    // the record length exactly as S/initialize.cpp:505-513 computes it
    int i = 2*tDof;
    if (com_mod.dFlag) i = 3*tDof;
    if (com_mod.pstEq) i = i + com_mod.nsymd;
    if (com_mod.sstEq) i = i + com_mod.nsd;
    i = sizeof(int)*(1 + com_mod.stamp.size()) + sizeof(double)*(2 + com_mod.nEq + com_mod.cplBC.nX + i*com_mod.tnNo);
    com_mod.recLn = i;
    *recLn_out = i;
    std::array<double,3> timeP = {utils::cput(), 0.0, 0.0};
    output::write_restart(&sim, timeP);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// output::output_result (S/output.cpp:46-180) into the history file `path`: the header block (co = 1) followed by one line
// (co = 2, or 3 when saved) for an equation with the given norms.  The reference takes the elapsed time from the wall clock:
// timeP[0] is set so that the line is written `elapsed` seconds "after the start" (the millisecond jitter stays below the four
// printed digits for the values the test uses).
int ref_io_history(const char* path, int nEq, const char* sym, int cTS, int itr, int saved, double elapsed, double eq_iNorm, double eq_pNorm,
                   double ri_iNorm, double ri_fNorm, double ri_dB, double ri_callD, int ri_itr, int ri_suc)
{
  try {
    mpistub_set_world(1);
    mpistub_bind_rank(0);
    Simulation sim;
    sim.logger.initialize(path, false);
    auto& com_mod = sim.com_mod;
    com_mod.nEq = nEq;
    com_mod.eq.resize(nEq);
    com_mod.cTS = cTS;
    auto& eq = com_mod.eq[0];
    eq.sym = sym;
    eq.itr = itr;
    eq.maxItr = 1000;
    eq.iNorm = eq_iNorm;
    eq.pNorm = eq_pNorm;
    eq.FSILS.RI.iNorm = ri_iNorm;
    eq.FSILS.RI.fNorm = ri_fNorm;
    eq.FSILS.RI.dB = ri_dB;
    eq.FSILS.RI.callD = ri_callD;
    eq.FSILS.RI.itr = ri_itr;
    eq.FSILS.RI.suc = ri_suc != 0;
    std::array<double,3> timeP = {utils::cput(), 0.0, 0.0};
    output::output_result(&sim, timeP, 1, 0);          // header; leaves timeP[0] = cput() - timeP[0] ~ 0
    timeP[0] = utils::cput() - elapsed;
    timeP[1] = 0.0;
    output::output_result(&sim, timeP, saved ? 3 : 2, 0);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// mat_models_carray::get_visc_stress_and_tangent<3> (S/mat_models_carray.h:1578) at one Gauss point: model 1 Newtonian, 2 potential.
// Nx (3 x eNoN, column-major), vx / F row-major 3x3.  Outputs Svis (3x3 row-major), Kvis_u / Kvis_v (9 x eNoN x eNoN, Array3 layout).
int ref_visc(int model, double mu, int eNoN, const double* Nx, const double* vx9, const double* F9, double* Svis9, double* Ku, double* Kv)
{
  try {
    dmnType dmn;
    dmn.solid_visc.viscType = (model == 1) ? consts::SolidViscosityModelType::viscType_Newtonian : consts::SolidViscosityModelType::viscType_Potential;
    dmn.solid_visc.mu = mu;
    Array<double> Nx_a(3, eNoN);
    std::memcpy(Nx_a.data(), Nx, sizeof(double)*3*size_t(eNoN));
    double vx[3][3], F[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { vx[i][j] = vx9[i*3 + j]; F[i][j] = F9[i*3 + j]; }
    Array<double> Svis(3, 3);
    Array3<double> Kvis_u(9, eNoN, eNoN), Kvis_v(9, eNoN, eNoN);
    mat_models_carray::get_visc_stress_and_tangent<3>(dmn, eNoN, Nx_a, vx, F, Svis, Kvis_u, Kvis_v);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Svis9[i*3 + j] = Svis(i, j);
    std::memcpy(Ku, Kvis_u.data(), sizeof(double)*9*size_t(eNoN)*eNoN);
    std::memcpy(Kv, Kvis_v.data(), sizeof(double)*9*size_t(eNoN)*eNoN);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

} // extern "C"
