// B200LinearAlgebra.cpp — see B200LinearAlgebra.h.  Thin: it only translates the reference's
// containers (ComMod, eqType, FSILS_lhsType; all column-major, Array.h:379) into the flat arrays of
// the C ABI (include/svb200.h) and converts status codes into the std::runtime_error exceptions the
// reference's virtuals use.  No numerical work happens here.
#include "B200LinearAlgebra.h"

#include "svb200.h"
#include "lhsa.h"
#include "all_fun.h"
#include "fft.h"
#include "utils.h"

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

std::set<consts::LinearAlgebraType> B200LinearAlgebra::valid_assemblers = {
  consts::LinearAlgebraType::none,
  consts::LinearAlgebraType::fsils,
  B200_LINEAR_ALGEBRA_TYPE,
};

B200LinearAlgebra::B200LinearAlgebra()
{
  interface_type = B200_LINEAR_ALGEBRA_TYPE;
  assembly_type = consts::LinearAlgebraType::fsils;
  preconditioner_type = consts::PreconditionerType::PREC_FSILS;
}

// The reference never deletes eq.linear_algebra, so the totals are printed when the process exits (or by the destructor when an
// owner does delete the object).
namespace {
std::vector<B200LinearAlgebra*>& live_instances() { static std::vector<B200LinearAlgebra*> v; return v; }
void report_at_exit() { for (auto* p : live_instances()) if (p) p->report(); }
}

B200LinearAlgebra::~B200LinearAlgebra()
{
  for (auto& p : live_instances()) if (p == this) p = nullptr;
  if (h_) { report(); b200_destroy(h_); }
}

void B200LinearAlgebra::check(int rc, const char* what)
{
  if (rc != 0) {
    throw std::runtime_error(std::string("[B200LinearAlgebra] ") + what + ": " + b200_last_error(h_));
  }
}

void B200LinearAlgebra::check_options(const consts::PreconditionerType prec_cond_type,
    const consts::LinearAlgebraType assembly_type)
{
  std::string error_msg;
  if (valid_assemblers.count(assembly_type) == 0) {
    error_msg = "b200 linear algebra can't use '" + LinearAlgebra::type_to_name.at(assembly_type) + "' for assembly.";
  }
  // the backend implements the FSILS preconditioners (diagonal 'fsils' and row-column scaling 'rcs')
  if (consts::fsils_preconditioners.count(prec_cond_type) == 0) {
    error_msg = "b200 linear algebra can't use '" + consts::preconditioner_type_to_name.at(prec_cond_type) +
        "' for a preconditioner.";
  }
  if (error_msg != "") {
    throw std::runtime_error("[svFSIplus] ERROR: " + error_msg);
  }
}

void B200LinearAlgebra::set_assembly(consts::LinearAlgebraType atype)
{
  if (atype == consts::LinearAlgebraType::none) {
    return;
  }
  if (valid_assemblers.count(atype) == 0) {
    throw std::runtime_error("[B200LinearAlgebra] ERROR: Can't set b200 linear algebra to use '" +
        LinearAlgebra::type_to_name.at(atype) + "' for assembly.");
  }
  assembly_type = atype;
  device_assembly_ = (atype == B200_LINEAR_ALGEBRA_TYPE);
}

void B200LinearAlgebra::set_preconditioner(consts::PreconditionerType prec_type)
{
  if (consts::fsils_preconditioners.count(prec_type) == 0) {
    throw std::runtime_error("[B200LinearAlgebra] ERROR: b200 linear algebra can't use '" +
        consts::preconditioner_type_to_name.at(prec_type) + "' for a preconditioner.");
  }
  preconditioner_type = prec_type;
}

/// Called once per equation by add_eq_linear_algebra (main.cpp:68-77), after com_mod.lhs, rowPtr,
/// colPtr and the faces exist.  The device handle is created here, not in the constructor.
void B200LinearAlgebra::initialize(ComMod& com_mod, eqType& lEq)
{
  if (h_) return;
  // one rank per GPU: rank r of the solver's communicator drives device r mod (devices on this node)
  if (device_ < 0) {
    const int n = b200_device_count();
    device_ = (n > 0) ? com_mod.cm.idcm() % n : 0;
  }
  if (b200_create(&h_, device_) != 0) {
    throw std::runtime_error(std::string("[B200LinearAlgebra] ") + b200_last_error(nullptr));
  }
  if (live_instances().empty()) std::atexit(report_at_exit);
  live_instances().push_back(this);
  // multi-rank: rank 0 creates the NCCL id, the solver's own communicator distributes it
  const int nranks = com_mod.cm.np();
  if (nranks > 1) {
    char uid[128];
    if (com_mod.cm.idcm() == 0) check(b200_comm_unique_id(uid), "b200_comm_unique_id");
    MPI_Bcast(uid, 128, MPI_CHAR, 0, com_mod.cm.com());
    check(b200_comm_init(h_, com_mod.cm.idcm(), nranks, uid), "b200_comm_init");
  }
  upload_structure(com_mod);
}

/// CSR pattern (lhsa, lhsa.cpp:153), node map, halo lists and faces (fsils_lhs_create lhs.cpp:57,
/// fsils_bc_create bc.cpp:45) -> device, once.
void B200LinearAlgebra::upload_structure(ComMod& com_mod)
{
  auto& lhs = com_mod.lhs;
  const int nNo = lhs.nNo;
  std::vector<int> req_rank, req_n, req_ptr;
  for (int i = 0; i < lhs.nReq; i++) {
    req_rank.push_back(lhs.cS[i].iP);
    req_n.push_back(lhs.cS[i].n);
    for (int j = 0; j < lhs.cS[i].n; j++) req_ptr.push_back(lhs.cS[i].ptr(j));
  }
  check(b200_lhs_create(h_, lhs.gnNo, nNo, lhs.mynNo, lhs.nnz, com_mod.rowPtr.data(), com_mod.colPtr.data(),
                        lhs.map.data(), lhs.nReq, req_rank.data(), req_n.data(), req_ptr.data(), lhs.nFaces),
        "b200_lhs_create");
  structure_uploaded_ = true;
  update_faces(com_mod);
}

void B200LinearAlgebra::update_faces(ComMod& com_mod)
{
  auto& lhs = com_mod.lhs;
  for (int f = 0; f < lhs.nFaces; f++) {
    auto& face = lhs.face[f];      // (fsils_bc_create never raises face.foC in the C++ reference: upload every slot)
    const int bGrp = (face.bGrp == fsi_linear_solver::BcType::BC_TYPE_Dir) ? B200_BC_DIR : B200_BC_NEU;
    check(b200_face_set(h_, f, face.nNo, face.dof, bGrp, face.glob.data(), face.val.data(), face.sharedFlag ? 1 : 0),
          "b200_face_set");
  }
}

/// ls_alloc contract (ls.cpp:51-60): afterwards the system is zero.
void B200LinearAlgebra::alloc(ComMod& com_mod, eqType& lEq)
{
  const int dof = com_mod.dof;
  if (!device_assembly_) {
    com_mod.Val.resize(dof*dof, com_mod.lhs.nnz);      // host assembly target, exactly like FsilsLinearAlgebra::alloc
  }
  check(b200_zero(h_, dof), "b200_zero");
  any_device_contribution_ = false;
  ustruct_on_device_ = false;
}

/// Per-element entry (boundary faces and any physics without a device kernel).
void B200LinearAlgebra::assemble(ComMod& com_mod, const int num_elem_nodes, const Vector<int>& eqN,
    const Array3<double>& lK, const Array<double>& lR)
{
  if (!device_assembly_) {
    lhsa_ns::do_assem(com_mod, num_elem_nodes, eqN, lK, lR);
    return;
  }
  check(b200_assemble_elem(h_, num_elem_nodes, eqN.data(), lK.data(), lR.data()), "b200_assemble_elem");
  any_device_contribution_ = true;
}

void B200LinearAlgebra::upload_mesh(ComMod& com_mod, const mshType& lM)
{
  // IEN holds assembly (local) node ids already (ComMod.h:893); x is com_mod.x (3 x tnNo)
  check(b200_mesh_set(h_, lM.eNoN, lM.nEl, lM.IEN.data(), com_mod.x.data(), lM.qmTET4), "b200_mesh_set");
  if (lM.nFn == 2 && lM.fN.size() != 0) {
    check(b200_mesh_fibers(h_, 2, lM.fN.data()), "b200_mesh_fibers");      // fN(nsd*nFn, nEl), ComMod.h:975
  }
  mesh_uploaded_ = &lM;
}

// Where the work ran: every routing decision of the three assembly hooks is counted, the first host fall-back of each kind is
// announced once on stdout (a user must be able to tell a device-assembly run from a host-assembly run), and the destructor
// prints the totals.
void B200LinearAlgebra::note(int which, bool on_device, const std::string& what)
{
  (on_device ? stats_.device : stats_.host)[which]++;
  if (!on_device && device_assembly_ && !(stats_.announced & (1u << which))) {
    stats_.announced |= (1u << which);
    // (stdio, not iostream: in a build that links libstdc++ statically into a dlopen'ed library - the test harness does - that
    // library's own std::cout has no locale facets and formatted output of numbers crashes)
    std::printf("[B200LinearAlgebra] %s: no device kernel for this physics / element type / option -> the reference's host path "
                "runs for it (assembly on the host, solve on the GPU)\n", what.c_str());
    std::fflush(stdout);
  }
}

void B200LinearAlgebra::report() const
{
  static const char* names[3] = {"whole-mesh assemblies", "Neumann faces", "follower-load faces"};
  std::printf("[B200LinearAlgebra] solves on the GPU: %ld", stats_.solves);
  for (int i = 0; i < 3; i++)
    if (stats_.device[i] + stats_.host[i] > 0)
      std::printf("; %s: %ld on the GPU, %ld on the host", names[i], stats_.device[i], stats_.host[i]);
  std::printf("\n");
  std::fflush(stdout);
}

bool B200LinearAlgebra::assemble_mesh(ComMod& com_mod, const mshType& lM, const Array<double>& Ag,
    const Array<double>& Yg, const Array<double>& Dg, const CepMod* cep_mod)
{
  const bool ok = assemble_mesh_impl(com_mod, lM, Ag, Yg, Dg, cep_mod);
  if (device_assembly_) note(0, ok, "equation " + com_mod.eq[com_mod.cEq].sym + ", mesh " + lM.name);
  return ok;
}

bool B200LinearAlgebra::assemble_mesh_impl(ComMod& com_mod, const mshType& lM, const Array<double>& Ag,
    const Array<double>& Yg, const Array<double>& Dg, const CepMod* cep_mod)
{
  using namespace consts;
  if (!device_assembly_) return false;
  auto& eq = com_mod.eq[com_mod.cEq];
  // one function space, 3-D; anything else falls back to the reference's construct_*
  if (lM.nFs != 1 || com_mod.nsd != 3) return false;
  if (eq.phys == EquationType::phys_FSI) return assemble_fsi_mesh(com_mod, lM, Ag, Yg, Dg, cep_mod);
  if (eq.nDmn != 1) return assemble_domains_mesh(com_mod, lM, Ag, Yg, Dg, cep_mod);
  switch (eq.phys) {
    case EquationType::phys_fluid:
      return assemble_fluid_mesh(com_mod, lM, Ag, Yg);
    case EquationType::phys_struct:
    case EquationType::phys_lElas:
    case EquationType::phys_mesh:
      return assemble_solid_mesh(com_mod, lM, Ag, Yg, Dg, cep_mod);
    case EquationType::phys_ustruct:
      return assemble_ustruct_mesh(com_mod, lM, Ag, Yg, Dg, cep_mod);
    default:
      return false;
  }
}

bool B200LinearAlgebra::assemble_fluid_mesh(ComMod& com_mod, const mshType& lM, const Array<double>& Ag,
    const Array<double>& Yg)
{
  using namespace consts;
  auto& eq = com_mod.eq[com_mod.cEq];
  // device kernels: Navier-Stokes VMS (one function space) on TET4, HEX8 and TET10
  if ((lM.eType != ElementType::TET4 && lM.eType != ElementType::HEX8 && lM.eType != ElementType::TET10) || com_mod.dof != 4) return false;
  if (mesh_uploaded_ != &lM) upload_mesh(com_mod, lM);

  b200_fluid_props p;
  if (!fill_fluid_props(com_mod, eq, eq.dmn[0], p)) return false;

  check(b200_state_set(h_, com_mod.tDof, Ag.data(), Yg.data(), com_mod.Bf.data()), "b200_state_set");
  check(b200_assemble_fluid(h_, &p), "b200_assemble_fluid");
  any_device_contribution_ = true;
  return true;
}

bool B200LinearAlgebra::fill_fluid_props(ComMod& com_mod, const eqType& eq, const dmnType& dmn, b200_fluid_props& p)
{
  using namespace consts;
  p.dt = com_mod.dt; p.am = eq.am; p.af = eq.af; p.gam = eq.gam;
  p.tDof = com_mod.tDof; p.mvMsh = com_mod.mvMsh ? 1 : 0;
  p.rho = dmn.prop.at(PhysicalProperyType::fluid_density);
  p.f[0] = dmn.prop.at(PhysicalProperyType::f_x);
  p.f[1] = dmn.prop.at(PhysicalProperyType::f_y);
  p.f[2] = dmn.prop.at(PhysicalProperyType::f_z);
  p.Kinv = dmn.prop.at(PhysicalProperyType::inverse_darcy_permeability);
  switch (dmn.fluid_visc.viscType) {
    case FluidViscosityModelType::viscType_Const: p.viscType = 0; break;
    case FluidViscosityModelType::viscType_CY:    p.viscType = 1; break;
    case FluidViscosityModelType::viscType_Cass:  p.viscType = 2; break;
    default: return false;
  }
  p.mu_i = dmn.fluid_visc.mu_i; p.mu_o = dmn.fluid_visc.mu_o; p.lam = dmn.fluid_visc.lam;
  p.a = dmn.fluid_visc.a; p.n = dmn.fluid_visc.n;
  return true;
}

/// mat_models::get_fib_stress (mat_models.cpp:126-139) + the cross-fibre factor (mat_models_carray.h:225): the fibre
/// reinforcement stress of this time step, steady (Tf.g) or interpolated from the Fourier series (Tf.gt).
void B200LinearAlgebra::fibre_stress(const ComMod& com_mod, const fibStrsType& Tf, double& Tfa, double& Tsa)
{
  using namespace consts;
  Tfa = 0.0;
  if (utils::btest(Tf.fType, iBC_std)) {
    Tfa = Tf.g;
  } else if (utils::btest(Tf.fType, iBC_ustd)) {
    Vector<double> gv(1), tv(1);
    ifft(com_mod, Tf.gt, gv, tv);
    Tfa = gv[0];
  }
  Tsa = Tfa*Tf.eta_s;
}

bool B200LinearAlgebra::fill_struct_props(ComMod& com_mod, const eqType& eq, const dmnType& dmn, b200_struct_props& sp)
{
  using namespace consts;
  const auto& stM = dmn.stM;
  // solid viscosity (get_visc_stress_and_tangent, mat_models_carray.h:1578): device kernel for struct equations (any number of domains)
  switch (dmn.solid_visc.viscType) {
    case SolidViscosityModelType::viscType_NA:        sp.viscType = 0; break;
    case SolidViscosityModelType::viscType_Newtonian: sp.viscType = 1; break;
    case SolidViscosityModelType::viscType_Potential: sp.viscType = 2; break;
    default: return false;
  }
  sp.visc_mu = dmn.solid_visc.mu;
  if (sp.viscType != 0 && eq.phys != EquationType::phys_struct && eq.phys != EquationType::phys_FSI) return false;
  fibre_stress(com_mod, stM.Tf, sp.Tfa, sp.Tsa);
  if (sp.Tfa != 0.0 && stM.isoType != ConstitutiveModelType::stIso_nHook && stM.isoType != ConstitutiveModelType::stIso_HO &&
      stM.isoType != ConstitutiveModelType::stIso_MR && stM.isoType != ConstitutiveModelType::stIso_HGO &&
      stM.isoType != ConstitutiveModelType::stIso_Gucci && stM.isoType != ConstitutiveModelType::stIso_HO_ma) return false;
  switch (stM.isoType) {
    case ConstitutiveModelType::stIso_nHook: sp.isoType = 0; break;
    case ConstitutiveModelType::stIso_StVK:  sp.isoType = 1; break;
    case ConstitutiveModelType::stIso_mStVK: sp.isoType = 2; break;
    case ConstitutiveModelType::stIso_HO:    sp.isoType = 3; break;      // needs lM.fN with two families (upload_mesh)
    case ConstitutiveModelType::stIso_MR:    sp.isoType = 4; break;
    case ConstitutiveModelType::stIso_HGO:   sp.isoType = 5; sp.kap = stM.kap; break;   // two fibre families (upload_mesh)
    case ConstitutiveModelType::stIso_Gucci: sp.isoType = 6; break;                     // fibre + sheet frame (upload_mesh)
    case ConstitutiveModelType::stIso_HO_ma: sp.isoType = 7; break;                     // HO with full fibre invariants (upload_mesh)
    default: return false;
  }
  sp.a = stM.a; sp.b = stM.b; sp.aff = stM.aff; sp.bff = stM.bff; sp.ass = stM.ass; sp.bss = stM.bss;
  sp.afs = stM.afs; sp.bfs = stM.bfs; sp.khs = stM.khs;
  switch (stM.volType) {
    case ConstitutiveModelType::stVol_Quad: sp.volType = 1; break;
    case ConstitutiveModelType::stVol_ST91: sp.volType = 2; break;
    case ConstitutiveModelType::stVol_M94:  sp.volType = 3; break;
    default: sp.volType = 0; break;
  }
  sp.dt = com_mod.dt; sp.am = eq.am; sp.af = eq.af; sp.gam = eq.gam; sp.beta = eq.beta;
  sp.tDof = com_mod.tDof; sp.s = eq.s;
  sp.rho = dmn.prop.at(PhysicalProperyType::solid_density);
  sp.dmp = dmn.prop.at(PhysicalProperyType::damping);
  sp.f[0] = dmn.prop.at(PhysicalProperyType::f_x);
  sp.f[1] = dmn.prop.at(PhysicalProperyType::f_y);
  sp.f[2] = dmn.prop.at(PhysicalProperyType::f_z);
  sp.C10 = stM.C10; sp.C01 = stM.C01; sp.Kpen = stM.Kpen;
  return true;
}

/// FSI equation (construct_fsi, fsi.cpp:42): fluid and struct domains of one TET4 / HEX8 / TET10 mesh in one dof-4 system.
bool B200LinearAlgebra::assemble_fsi_mesh(ComMod& com_mod, const mshType& lM, const Array<double>& Ag,
    const Array<double>& Yg, const Array<double>& Dg, const CepMod* cep_mod)
{
  using namespace consts;
  auto& eq = com_mod.eq[com_mod.cEq];
  if ((lM.eType != ElementType::TET4 && lM.eType != ElementType::HEX8 && lM.eType != ElementType::TET10) || com_mod.dof != 4 || !com_mod.mvMsh) return false;
  // the wall's prestress is read by construct_fsi (fsi.cpp:147-148) and never accumulated there; a prestress EQUATION is a struct run
  if (com_mod.pstEq) return false;
  if (com_mod.pS0.size() != 0 && (com_mod.pS0.nrows() != 6 || com_mod.pS0.ncols() != com_mod.tnNo)) return false;
  if (cep_mod && (cep_mod->cem.cpld || cep_mod->cem.aStress || cep_mod->cem.aStrain)) return false;
  const int nDmn = eq.nDmn;
  std::vector<int> kinds(nDmn, -1);
  std::vector<b200_fluid_props> fl(nDmn);
  std::vector<b200_struct_props> st(nDmn);
  for (int d = 0; d < nDmn; d++) {
    if (eq.dmn[d].phys == EquationType::phys_fluid) {
      kinds[d] = 0;
      if (!fill_fluid_props(com_mod, eq, eq.dmn[d], fl[d])) return false;
    } else if (eq.dmn[d].phys == EquationType::phys_struct) {
      kinds[d] = 1;
      if (!fill_struct_props(com_mod, eq, eq.dmn[d], st[d])) return false;
      if ((st[d].isoType == 3 || st[d].isoType == 5 || st[d].isoType == 6 || st[d].isoType == 7 || st[d].Tfa != 0.0) && (lM.nFn != 2 || lM.fN.size() == 0)) return false;
    } else {
      return false;
    }
  }
  if (mesh_uploaded_ != &lM) upload_mesh(com_mod, lM);
  if (domains_uploaded_ != &lM) {
    std::vector<int> elem_dmn(lM.nEl);
    for (int e = 0; e < lM.nEl; e++) elem_dmn[e] = all_fun::domain(com_mod, lM, com_mod.cEq, e);
    check(b200_mesh_domains(h_, nDmn, elem_dmn.data()), "b200_mesh_domains");
    domains_uploaded_ = &lM;
  }
  check(b200_state_set(h_, com_mod.tDof, Ag.data(), Yg.data(), com_mod.Bf.data()), "b200_state_set");
  // Do: the old mesh displacement the moving-mesh Neumann faces of this equation are evaluated on (assemble_face)
  const bool have_do = com_mod.Do.size() != 0 && com_mod.Do.nrows() == com_mod.tDof;
  check(b200_disp_set(h_, com_mod.tDof, Dg.data(), have_do ? com_mod.Do.data() : nullptr), "b200_disp_set");
  do_uploaded_ = have_do;
  if (com_mod.pS0.size() != 0 || prestress_on_device_) {
    check(b200_prestress_set(h_, com_mod.pS0.size() != 0 ? com_mod.pS0.data() : nullptr, 0), "b200_prestress_set");
    prestress_on_device_ = com_mod.pS0.size() != 0;
  }
  check(b200_assemble_fsi(h_, nDmn, kinds.data(), fl.data(), st.data()), "b200_assemble_fsi");
  any_device_contribution_ = true;
  return true;
}

/// Neumann face on the device (b_assem_neu_bc, eq_assem.cpp:58): b_fluid for fluid domains, b_l_elas for the solid-type ones.
/// Equations with several domains (FSI: fluid lumen + struct wall) are taken when every element of the face lies in ONE domain
/// (all_fun::domain of the parent element, all_fun.cpp:149-175) - the reference switches physics per face element, a face
/// that straddles domains stays on the host path.  Moving meshes (com_mod.mvMsh, gnnb on x + Do(nsd+1..), nn.cpp:609-640)
/// use the Do that assemble_mesh uploaded with this iteration's state.
bool B200LinearAlgebra::assemble_face(ComMod& com_mod, const faceType& lFa, const Vector<double>& hg, const Array<double>& Yg)
{
  const bool ok = assemble_face_impl(com_mod, lFa, hg, Yg);
  if (device_assembly_) note(1, ok, "equation " + com_mod.eq[com_mod.cEq].sym + ", Neumann face " + lFa.name);
  return ok;
}

bool B200LinearAlgebra::assemble_face_impl(ComMod& com_mod, const faceType& lFa, const Vector<double>& hg, const Array<double>& Yg)
{
  using namespace consts;
  if (!device_assembly_ || !any_device_contribution_ || com_mod.nsd != 3) return false;
  auto& eq = com_mod.eq[com_mod.cEq];
  const auto& msh = com_mod.msh[lFa.iM];
  if (msh.lShl || mesh_uploaded_ != &msh) return false;
  if (lFa.eType != ElementType::TRI3 && lFa.eType != ElementType::QUD4 && lFa.eType != ElementType::TRI6) return false;
  if (lFa.eType == ElementType::TRI3 && lFa.qmTRI3 != 2.0/3.0) return false;       // device tables use the default rule
  int iDmn = 0;
  if (eq.nDmn != 1) {
    if (lFa.nEl == 0) return false;
    iDmn = all_fun::domain(com_mod, msh, com_mod.cEq, lFa.gE(0));
    for (int e = 1; e < lFa.nEl; e++) if (all_fun::domain(com_mod, msh, com_mod.cEq, lFa.gE(e)) != iDmn) return false;
  }
  int kind;
  switch (eq.dmn[iDmn].phys) {
    case EquationType::phys_fluid: kind = 0; break;
    case EquationType::phys_lElas: case EquationType::phys_struct: case EquationType::phys_ustruct:
    case EquationType::phys_mesh: case EquationType::phys_stokes: kind = 1; break;
    default: return false;
  }
  if (kind == 0 && com_mod.dof != 4) return false;
  if (com_mod.mvMsh && (!do_uploaded_ || com_mod.tDof < 2*com_mod.nsd + 1)) return false;
  auto it = face_meshes_.find(&lFa);
  if (it == face_meshes_.end()) {
    const int slot = int(face_meshes_.size());
    check(b200_face_mesh_set(h_, slot, lFa.eNoN, lFa.nEl, lFa.IEN.data(), lFa.gE.data()), "b200_face_mesh_set");
    it = face_meshes_.emplace(&lFa, slot).first;
  }
  b200_bneu_props p;
  p.dt = com_mod.dt; p.af = eq.af; p.gam = eq.gam; p.tDof = com_mod.tDof; p.mvMsh = com_mod.mvMsh ? 1 : 0;
  p.rho = 0.0; p.bfs = 0.0;
  if (kind == 0) {
    p.rho = eq.dmn[iDmn].prop.at(PhysicalProperyType::fluid_density);
    p.bfs = eq.dmn[iDmn].prop.at(PhysicalProperyType::backflow_stab);
  }
  (void)Yg;      // the device holds the Yg uploaded for the volume assembly of this iteration
  check(b200_assemble_bneu(h_, it->second, kind, &p, hg.data()), "b200_assemble_bneu");
  return true;
}

/// Follower pressure load on the device (b_neu_folw_p, eq_assem.cpp:186) for a single-domain struct or ustruct equation.
bool B200LinearAlgebra::assemble_follower_face(ComMod& com_mod, const faceType& lFa, const Vector<double>& hg, const Array<double>& Dg)
{
  const bool ok = assemble_follower_face_impl(com_mod, lFa, hg, Dg);
  if (device_assembly_) note(2, ok, "equation " + com_mod.eq[com_mod.cEq].sym + ", follower-load face " + lFa.name);
  return ok;
}

bool B200LinearAlgebra::assemble_follower_face_impl(ComMod& com_mod, const faceType& lFa, const Vector<double>& hg, const Array<double>& Dg)
{
  using namespace consts;
  if (!device_assembly_ || !any_device_contribution_ || com_mod.nsd != 3 || com_mod.mvMsh) return false;
  auto& eq = com_mod.eq[com_mod.cEq];
  const auto& msh = com_mod.msh[lFa.iM];
  if (eq.nDmn != 1 || msh.lShl || mesh_uploaded_ != &msh) return false;
  const bool us = (eq.dmn[0].phys == EquationType::phys_ustruct);
  if (!us && eq.dmn[0].phys != EquationType::phys_struct) return false;
  if (com_mod.dof != (us ? 4 : 3) || (us && !ustruct_on_device_)) return false;
  const bool pair_ok = (msh.eType == ElementType::TET4 && lFa.eType == ElementType::TRI3) ||
                       (msh.eType == ElementType::HEX8 && lFa.eType == ElementType::QUD4) ||
                       (msh.eType == ElementType::TET10 && lFa.eType == ElementType::TRI6);
  if (!pair_ok || (lFa.eType == ElementType::TRI3 && lFa.qmTRI3 != 2.0/3.0)) return false;
  auto it = face_meshes_.find(&lFa);
  if (it == face_meshes_.end()) {
    const int slot = int(face_meshes_.size());
    check(b200_face_mesh_set(h_, slot, lFa.eNoN, lFa.nEl, lFa.IEN.data(), lFa.gE.data()), "b200_face_mesh_set");
    it = face_meshes_.emplace(&lFa, slot).first;
  }
  b200_bfolw_props p;
  p.dt = com_mod.dt; p.af = eq.af; p.beta = eq.beta; p.tDof = com_mod.tDof; p.s = eq.s;
  p.ustruct = us ? 1 : 0; p.am = eq.am; p.gam = eq.gam;
  (void)Dg;      // the device holds the Dg uploaded for the volume assembly of this iteration
  check(b200_assemble_bfolw(h_, it->second, &p, hg.data()), "b200_assemble_bfolw");
  return true;
}

/// ustruct (construct_usolid, ustruct.cpp:216) on equal-order TET4 / HEX8 / TET10 with idMap = identity.
bool B200LinearAlgebra::assemble_ustruct_mesh(ComMod& com_mod, const mshType& lM, const Array<double>& Ag,
    const Array<double>& Yg, const Array<double>& Dg, const CepMod* cep_mod)
{
  using namespace consts;
  auto& eq = com_mod.eq[com_mod.cEq];
  if ((lM.eType != ElementType::TET4 && lM.eType != ElementType::HEX8 && lM.eType != ElementType::TET10) || com_mod.dof != 4) return false;
  if (cep_mod && (cep_mod->cem.cpld || cep_mod->cem.aStress || cep_mod->cem.aStrain)) return false;
  for (int a = 0; a < com_mod.tnNo; a++) if (com_mod.idMap(a) != a) return false;      // undeformed-Neumann faces
  const auto& dmn = eq.dmn[0];
  const auto& stM = dmn.stM;
  int iso;
  switch (stM.isoType) {                       // the laws of get_pk2cc_dev with a device kernel
    case ConstitutiveModelType::stIso_nHook: iso = 0; break;
    case ConstitutiveModelType::stIso_HO:    iso = 3; break;
    case ConstitutiveModelType::stIso_MR:    iso = 4; break;
    case ConstitutiveModelType::stIso_HGO:   iso = 5; break;
    case ConstitutiveModelType::stIso_Gucci: iso = 6; break;
    case ConstitutiveModelType::stIso_HO_ma: iso = 7; break;
    default: return false;
  }
  if ((iso == 3 || iso == 5 || iso == 6 || iso == 7) && (lM.nFn != 2 || lM.fN.size() == 0)) return false;
  b200_ustruct_props p{};
  switch (dmn.solid_visc.viscType) {             // solid viscosity (ustruct.cpp:1275-1302)
    case SolidViscosityModelType::viscType_NA:        p.viscType = 0; break;
    case SolidViscosityModelType::viscType_Newtonian: p.viscType = 1; break;
    case SolidViscosityModelType::viscType_Potential: p.viscType = 2; break;
    default: return false;
  }
  p.visc_mu = dmn.solid_visc.mu;
  fibre_stress(com_mod, stM.Tf, p.Tfa, p.Tsa);
  if (p.Tfa != 0.0 && (lM.nFn != 2 || lM.fN.size() == 0)) return false;
  p.dt = com_mod.dt; p.am = eq.am; p.af = eq.af; p.gam = eq.gam;
  p.tDof = com_mod.tDof; p.s = eq.s;
  p.rho = dmn.prop.at(PhysicalProperyType::solid_density);
  p.f[0] = dmn.prop.at(PhysicalProperyType::f_x);
  p.f[1] = dmn.prop.at(PhysicalProperyType::f_y);
  p.f[2] = dmn.prop.at(PhysicalProperyType::f_z);
  p.elM = dmn.prop.at(PhysicalProperyType::elasticity_modulus);
  p.nu = dmn.prop.at(PhysicalProperyType::poisson_ratio);
  p.ctM = dmn.prop.at(PhysicalProperyType::ctau_M);
  p.ctC = dmn.prop.at(PhysicalProperyType::ctau_C);
  p.isoType = iso;
  p.C01 = stM.C01; p.kap = stM.kap;
  p.a = stM.a; p.b = stM.b; p.aff = stM.aff; p.bff = stM.bff; p.ass = stM.ass; p.bss = stM.bss;
  p.afs = stM.afs; p.bfs = stM.bfs; p.khs = stM.khs;
  switch (stM.volType) {
    case ConstitutiveModelType::stVol_Quad: p.volType = 1; break;
    case ConstitutiveModelType::stVol_ST91: p.volType = 2; break;
    case ConstitutiveModelType::stVol_M94:  p.volType = 3; break;
    default: p.volType = 0; break;
  }
  p.C10 = stM.C10; p.Kpen = stM.Kpen;
  if (mesh_uploaded_ != &lM) upload_mesh(com_mod, lM);
  check(b200_state_set(h_, com_mod.tDof, Ag.data(), Yg.data(), com_mod.Bf.data()), "b200_state_set");
  check(b200_disp_set(h_, com_mod.tDof, Dg.data(), nullptr), "b200_disp_set");
  check(b200_assemble_ustruct(h_, &p), "b200_assemble_ustruct");
  any_device_contribution_ = true;
  ustruct_on_device_ = true;
  return true;
}

bool B200LinearAlgebra::ustruct_r(ComMod& com_mod, const Array<double>& Yg)
{
  if (!device_assembly_ || !ustruct_on_device_) return false;
  const auto& eq = com_mod.eq[com_mod.cEq];
  if (eq.itr > 1) { com_mod.Rd = 0.0; return true; }
  const double amg = (eq.gam - eq.am) / (eq.gam - 1.0);
  const double ami = 1.0 / eq.am;
  check(b200_ustruct_r(h_, amg, ami, eq.s, com_mod.Ad.data()), "b200_ustruct_r");
  // The host time integrator (pic::picc, pic.cpp:92,139: dUl = Rd*coef[2] + R*coef[3]) reads com_mod.Rd, so the first iteration's
  // Rd = amg*Ad - Yg(s..) has to exist on the host as well (ustruct.cpp:1768-1775); the device copy only feeds Kd*Rd.
  // Rd keeps whatever it held on rows outside a ustruct domain, exactly as the reference loop does.
  const int nsd = com_mod.nsd;
  for (int a = 0; a < com_mod.tnNo; a++) {
    if (!all_fun::is_domain(com_mod, eq, a, consts::EquationType::phys_ustruct)) continue;
    for (int i = 0; i < nsd; i++) com_mod.Rd(i, a) = amg*com_mod.Ad(i, a) - Yg(eq.s + i, a);
  }
  return true;
}

/// struct (construct_dsolid, sv_struct.cpp:213), lElas (construct_l_elas, l_elas.cpp:58) and the ALE mesh
/// equation (construct_mesh, mesh.cpp:42) on TET4 / HEX8.  Features without a device kernel (fibre
/// stress, solid viscosity, prestress, electromechanics, anisotropic laws) return false.
bool B200LinearAlgebra::assemble_solid_mesh(ComMod& com_mod, const mshType& lM, const Array<double>& Ag,
    const Array<double>& Yg, const Array<double>& Dg, const CepMod* cep_mod)
{
  using namespace consts;
  auto& eq = com_mod.eq[com_mod.cEq];
  if ((lM.eType != ElementType::TET4 && lM.eType != ElementType::HEX8 && lM.eType != ElementType::TET10) || com_mod.dof != 3) return false;
  // prestress (com_mod.pS0 / pstEq, sv_struct.cpp:232-235): device kernel for the struct equation; lElas / mesh never read it
  const bool prestress = (eq.phys == EquationType::phys_struct) && (com_mod.pS0.size() != 0 || com_mod.pstEq);
  if (prestress && com_mod.pS0.size() != 0 && (com_mod.pS0.nrows() != 6 || com_mod.pS0.ncols() != com_mod.tnNo)) return false;
  if (cep_mod && (cep_mod->cem.cpld || cep_mod->cem.aStress || cep_mod->cem.aStrain)) return false;
  const auto& dmn = eq.dmn[0];

  b200_struct_props sp{};
  b200_lelas_props lp{};
  const bool is_struct = (eq.phys == EquationType::phys_struct);
  if (is_struct) {
    if (!fill_struct_props(com_mod, eq, dmn, sp)) return false;
    if ((sp.isoType == 3 || sp.isoType == 5 || sp.isoType == 6 || sp.isoType == 7 || sp.Tfa != 0.0) && (lM.nFn != 2 || lM.fN.size() == 0)) return false;
  } else {
    lp.dt = com_mod.dt; lp.am = eq.am; lp.af = eq.af; lp.beta = eq.beta;
    lp.tDof = com_mod.tDof; lp.s = eq.s;
    lp.mesh_mode = (eq.phys == EquationType::phys_mesh) ? 1 : 0;
    lp.rho = dmn.prop.at(PhysicalProperyType::solid_density);
    lp.elM = dmn.prop.at(PhysicalProperyType::elasticity_modulus);
    lp.nu = dmn.prop.at(PhysicalProperyType::poisson_ratio);
    lp.f[0] = dmn.prop.at(PhysicalProperyType::f_x);
    lp.f[1] = dmn.prop.at(PhysicalProperyType::f_y);
    lp.f[2] = dmn.prop.at(PhysicalProperyType::f_z);
  }
  if (mesh_uploaded_ != &lM) upload_mesh(com_mod, lM);
  check(b200_state_set(h_, com_mod.tDof, Ag.data(), Yg.data(), com_mod.Bf.data()), "b200_state_set");
  check(b200_disp_set(h_, com_mod.tDof, Dg.data(), lp.mesh_mode && !is_struct ? com_mod.Do.data() : nullptr), "b200_disp_set");
  if (is_struct) {
    if (prestress || prestress_on_device_) {
      check(b200_prestress_set(h_, com_mod.pS0.size() != 0 ? com_mod.pS0.data() : nullptr, com_mod.pstEq ? 1 : 0), "b200_prestress_set");
      prestress_on_device_ = prestress;
    }
    check(b200_assemble_struct(h_, &sp), "b200_assemble_struct");
    if (prestress && com_mod.pstEq) {
      // construct_dsolid adds this mesh's share to com_mod.pSn / pSa (sv_struct.cpp:333-343); pic::picc normalises them
      Array<double> pSn(6, com_mod.tnNo);
      Vector<double> pSa(com_mod.tnNo);
      check(b200_prestress_get(h_, pSn.data(), pSa.data()), "b200_prestress_get");
      for (int a = 0; a < com_mod.tnNo; a++) {
        com_mod.pSa(a) += pSa(a);
        for (int i = 0; i < 6; i++) com_mod.pSn(i,a) += pSn(i,a);
      }
    }
  }
  else check(b200_assemble_lelas(h_, &lp), "b200_assemble_lelas");
  any_device_contribution_ = true;
  return true;
}

/// Fluid or struct equation with several domains of that one physics (eq.nDmn > 1): one launch per domain over the
/// domain's element list.  Every element takes the properties of its own domain (all_fun::domain), which is what
/// construct_fluid does; construct_dsolid reads them through a stale copy of com_mod.cDmn (sv_struct.cpp:229) -- that
/// defect is not reproduced.
bool B200LinearAlgebra::assemble_domains_mesh(ComMod& com_mod, const mshType& lM, const Array<double>& Ag,
    const Array<double>& Yg, const Array<double>& Dg, const CepMod* cep_mod)
{
  using namespace consts;
  auto& eq = com_mod.eq[com_mod.cEq];
  if (lM.eType != ElementType::TET4 && lM.eType != ElementType::HEX8 && lM.eType != ElementType::TET10) return false;
  const int nDmn = eq.nDmn;
  const bool fluid = (eq.phys == EquationType::phys_fluid);
  if (!fluid && eq.phys != EquationType::phys_struct) return false;
  if (com_mod.dof != (fluid ? 4 : 3)) return false;
  std::vector<b200_fluid_props> fl(nDmn);
  std::vector<b200_struct_props> st(nDmn);
  for (int d = 0; d < nDmn; d++) {
    if (eq.dmn[d].phys != eq.phys) return false;
    if (fluid) {
      if (!fill_fluid_props(com_mod, eq, eq.dmn[d], fl[d])) return false;
    } else {
      if (!fill_struct_props(com_mod, eq, eq.dmn[d], st[d])) return false;
      if ((st[d].isoType == 3 || st[d].isoType == 5 || st[d].isoType == 6 || st[d].isoType == 7 || st[d].Tfa != 0.0) && (lM.nFn != 2 || lM.fN.size() == 0)) return false;
    }
  }
  if (!fluid) {
    if (com_mod.pS0.size() != 0 || com_mod.pstEq) return false;
    if (cep_mod && (cep_mod->cem.cpld || cep_mod->cem.aStress || cep_mod->cem.aStrain)) return false;
  }
  if (mesh_uploaded_ != &lM) upload_mesh(com_mod, lM);
  if (domains_uploaded_ != &lM) {
    std::vector<int> elem_dmn(lM.nEl);
    for (int e = 0; e < lM.nEl; e++) elem_dmn[e] = all_fun::domain(com_mod, lM, com_mod.cEq, e);
    check(b200_mesh_domains(h_, nDmn, elem_dmn.data()), "b200_mesh_domains");
    domains_uploaded_ = &lM;
  }
  check(b200_state_set(h_, com_mod.tDof, Ag.data(), Yg.data(), com_mod.Bf.data()), "b200_state_set");
  if (fluid) {
    check(b200_assemble_fluid_dmn(h_, nDmn, fl.data()), "b200_assemble_fluid_dmn");
  } else {
    check(b200_disp_set(h_, com_mod.tDof, Dg.data(), nullptr), "b200_disp_set");
    check(b200_assemble_struct_dmn(h_, nDmn, st.data()), "b200_assemble_struct_dmn");
  }
  any_device_contribution_ = true;
  return true;
}

/// ls_solve (ls.cpp:69-82) -> fsils_solve (solve.cpp:50) on the device.  On return com_mod.R holds
/// the solution, like FsilsLinearAlgebra::solve.
void B200LinearAlgebra::solve(ComMod& com_mod, eqType& lEq, const Vector<int>& incL, const Vector<double>& res)
{
  stats_.solves++;
  const int dof = com_mod.dof;
  auto& ls = lEq.FSILS;
  if (device_assembly_) {
    // host code may have added to com_mod.R since ls_alloc (it starts from zero): fold it in
    check(b200_add_R(h_, dof, com_mod.R.data()), "b200_add_R");
  } else {
    check(b200_set_R(h_, dof, com_mod.R.data()), "b200_set_R");
    check(b200_set_Val(h_, dof, com_mod.Val.data()), "b200_set_Val");
  }
  b200_tol RI{ls.RI.relTol, ls.RI.absTol, ls.RI.mItr, ls.RI.sD};
  b200_tol GM{ls.GM.relTol, ls.GM.absTol, ls.GM.mItr, ls.GM.sD};
  b200_tol CG{ls.CG.relTol, ls.CG.absTol, ls.CG.mItr, ls.CG.sD};
  const int prec = (lEq.linear_algebra_preconditioner == consts::PreconditionerType::PREC_RCS) ? B200_PREC_RCS : B200_PREC_FSILS;
  b200_ls_out out;
  check(b200_solve(h_, static_cast<int>(ls.LS_type), prec, &RI, &GM, &CG,
                   incL.size() ? incL.data() : nullptr, res.size() ? res.data() : nullptr, com_mod.R.data(), &out),
        "b200_solve");
  auto put = [](fsi_linear_solver::FSILS_subLsType& s, const b200_sub_out& o) {
    s.suc = o.suc != 0; s.itr = o.itr; s.iNorm = o.iNorm; s.fNorm = o.fNorm; s.dB = o.dB; s.callD = o.callD;
  };
  put(ls.RI, out.RI); put(ls.GM, out.GM); put(ls.CG, out.CG);
  ls.Resm = out.Resm; ls.Resc = out.Resc;
}
