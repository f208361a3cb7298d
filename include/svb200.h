/* svb200.h — C ABI of the B200 (sm_100a) backend for svMultiPhysics' nonlinear-step hot path:
 * element assembly into the block-CSR system followed by the FSILS-equivalent Krylov solve.
 *
 * This is the drop-in boundary.  The reference's plug-in interface is the C++ class
 * `LinearAlgebra` (Code/Source/solver/LinearAlgebra.h:39-63: alloc / assemble / check_options /
 * initialize / set_assembly / set_preconditioner / solve); the thin host class that implements it on
 * top of this header is svfsiplus_b200/host/B200LinearAlgebra.{h,cpp}, and INTEGRATION.md shows the
 * four registration lines a maintainer adds.  The shape of the layer follows the reference's own
 * extern "C" shim for Trilinos (Code/Source/solver/trilinos_impl.h:185-224).
 *
 * Conventions: plain pointers and sizes only; every array is HOST memory unless the name says
 * `_dev`; all indices are 0-based 32-bit ints; all reals are IEEE double.  2-D arrays use the
 * reference's column-major containers seen from C: R(dof,nNo) = R[i + dof*a] (the dof values of a
 * node are contiguous), Val(dof*dof,nnz) = Val[i*dof + j + dof*dof*p] = dR_i/du_j of non-zero p
 * (Code/Source/solver/Array.h:379; block row-major inside, liner_solver/spar_mul.cpp:229-236).
 * Every function returns 0 on success, non-zero on failure with a message in b200_last_error().
 * There is no CPU fallback: a call made without a usable CUDA device fails.
 */
#ifndef SVB200_H
#define SVB200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_handle b200_handle;

/* Linear solver / preconditioner codes: identical to the reference's enums
 * (liner_solver/fils_struct.hpp:70-76, solver/consts.h:426-440). */
enum { B200_LS_BICGS = 795, B200_LS_NS = 796, B200_LS_GMRES = 797, B200_LS_CG = 798 };
enum { B200_PREC_FSILS = 701, B200_PREC_RCS = 709 };
enum { B200_BC_DIR = 0, B200_BC_NEU = 1 };                  /* fils_struct.hpp:58 BcType */

/* Inputs of one FSILS_subLsType (fils_struct.hpp:211-257). */
typedef struct { double relTol, absTol; int mItr, sD; } b200_tol;
/* Outputs of one FSILS_subLsType. */
typedef struct { int suc, itr; double iNorm, fNorm, dB, callD; } b200_sub_out;
/* Outputs of FSILS_lsType (fils_struct.hpp:259-279). */
typedef struct { b200_sub_out RI, GM, CG; int Resm, Resc; } b200_ls_out;

/* Per-(equation, domain) constants of the fluid element, flattened from eqType/dmnType
 * (solver/ComMod.h:1021,431; look-ups at solver/fluid.cpp:1727-1737,2142-2200). */
typedef struct {
  double dt, am, af, gam;       /* com_mod.dt, eq.am, eq.af, eq.gam */
  int tDof, mvMsh;              /* com_mod.tDof, com_mod.mvMsh */
  double rho, f[3], Kinv;       /* fluid_density, f_x..f_z, inverse_darcy_permeability */
  int viscType;                 /* 0 constant, 1 Carreau-Yasuda, 2 Casson */
  double mu_i, mu_o, lam, a, n; /* dmn.fluid_visc */
} b200_fluid_props;

/* Per-(equation, domain) constants of the displacement-based solid element, flattened from eqType /
 * dmnType / stModelType (solver/sv_struct.cpp:576-594, solver/ComMod.h:345-388).  s = eq.s, the row of
 * the equation's first unknown inside Ag/Yg/Dg(tDof,nNo).  isoType: 0 neo-Hookean (C10 = mu/2),
 * 1 St.Venant-Kirchhoff (C10 = lambda, C01 = mu), 2 modified StVK (C10 = kappa, C01 = mu), 3 Holzapfel-Ogden
 * (solver/mat_models_carray.h:905-1135; needs b200_mesh_fibers), 4 Mooney-Rivlin (C10, C01; :438-540),
 * 5 Holzapfel-Gasser-Ogden (:544-688; needs b200_mesh_fibers), 6 Guccione (C10, bff, bss, bfs; :692-903; needs b200_mesh_fibers),
 * 7 Holzapfel-Ogden with modified anisotropy (HO-ma, full fibre invariants; :1137-1353; needs b200_mesh_fibers);
 * volType: 0 none, 1 Quad, 2 ST91, 3 M94 (solver/mat_models.cpp:1626-1645). */
typedef struct {
  double dt, am, af, gam, beta;
  int tDof, s;
  double rho, dmp, f[3];
  int isoType, volType;
  double C10, C01, Kpen;
  double a, b, aff, bff, ass, bss, afs, bfs, khs;   /* isoType 3 (Holzapfel-Ogden): stModelType a..bfs, khs */
  double Tfa, Tsa;   /* fibre-reinforcement / active stress along the fibre and sheet directions: what get_fib_stress returns
                        for stM.Tf at this time, and Tfa*Tf.eta_s (mat_models_carray.h:222-225); laws 0, 3, 4, 5, needs fibres */
  double kap;        /* isoType 5 (Holzapfel-Gasser-Ogden): fibre dispersion stM.kap; C10, aff, bff, ass, bss as in the XML */
  int viscType;      /* solid viscosity (dmn.solid_visc, mat_models_carray.h:1383-1590): 0 none, 1 Newtonian, 2 pseudo-potential;
                        struct equations (b200_assemble_struct, b200_assemble_struct_dmn) and the struct domains of the FSI equation */
  double visc_mu;
} b200_struct_props;

/* Linear elasticity (solver/l_elas.cpp:274-390).  mesh_mode != 0: the ALE mesh-motion equation
 * (solver/mesh.cpp:42-160): reference configuration x + Do(s..s+2), displacement Dg - Do, Gauss weights
 * without the Jacobian, no body force. */
typedef struct {
  double dt, am, af, beta;
  int tDof, s, mesh_mode;
  double rho, elM, nu, f[3];
} b200_lelas_props;

/* Mixed velocity-pressure solid (ustruct; solver/ustruct.cpp:1158-1575, 632-876).  elM, nu, ctM, ctC feed
 * get_tau (solver/mat_models.cpp:1655); Kpen, volType feed g_vol_pen (:1696).  isoType as in the struct properties, the
 * laws with an isochoric split (get_pk2cc_dev, mat_models.cpp:630): 0 neo-Hookean, 3 Holzapfel-Ogden, 4 Mooney-Rivlin,
 * 5 Holzapfel-Gasser-Ogden, 6 Guccione, 7 HO-ma (3, 5, 6, 7 need b200_mesh_fibers). */
typedef struct {
  double dt, am, af, gam;
  int tDof, s;
  double rho, f[3];
  double elM, nu, ctM, ctC;
  int isoType, volType;
  double C10, Kpen;
  double a, b, aff, bff, ass, bss, afs, bfs, khs;   /* isoType 3 (Holzapfel-Ogden) */
  double Tfa, Tsa;   /* fibre / sheet reinforcement stress as in the struct properties above; mat_models.cpp:682-684 */
  double C01, kap;   /* isoType 4 (Mooney-Rivlin) second modulus; isoType 5 (HGO) fibre dispersion */
  int viscType;      /* solid viscosity as in the struct properties (ustruct.cpp:1275-1302): 0 none, 1 Newtonian, 2 potential */
  double visc_mu;
} b200_ustruct_props;

/* ---- life cycle ------------------------------------------------------------------------- */
int  b200_create(b200_handle** h, int device);
void b200_destroy(b200_handle* h);
const char* b200_last_error(b200_handle* h);                 /* h may be NULL: last create error */
int  b200_device_count(void);
/* Gauss rule and shape functions the element kernels use for eNoN = 4 (TET4) / 8 (HEX8) / 10 (TET10): w[nG],
 * N[nG][eNoN], Nxi[nG][eNoN][3]; returns nG = 4 / 8 / 15 (what nn::select_ele leaves in lM.w / lM.N / lM.Nx,
 * solver/nn_elem_gip.h:40,501,520; nn_elem_gnn.h:732,1232,1256).  Host-only, needs no device. */
int  b200_elem_tables(int eNoN, double qmTET4, double* w, double* N, double* Nxi);

/* ---- communicator (replaces FSILS_commuType + MPI, liner_solver/commu.cpp:44) ------------- */
/* uid: 128 bytes produced on rank 0 and distributed by the caller (MPI_Bcast / torch.distributed). */
int b200_comm_unique_id(void* uid128);
int b200_comm_init(b200_handle* h, int rank, int nranks, const void* uid128);
/* Which transport carries the overlap-node adds (fsils_commuv, liner_solver/in_commu.cpp:111) and the Krylov all-reduces
 * (liner_solver/dot.cpp, norm.cpp, bcast.cpp:51) after b200_lhs_create: "p2p: ..." = the library's own kernels over
 * peer-mapped windows (CUDA IPC, NVLink stores + epoch flags), "nccl: <why not p2p>" = ncclSend/Recv + ncclAllReduce,
 * "none: single rank".  With more than one rank b200_lhs_create is COLLECTIVE (every rank calls it, also ranks without
 * neighbours).  SVB200_P2P=0 in the environment selects the NCCL path.  The string is owned by the handle. */
const char* b200_comm_transport(b200_handle* h);

/* ---- structure (replaces fsils_lhs_create liner_solver/lhs.cpp:57; consumes its result) ---- */
/* rowPtr(nNo+1)/colPtr(nnz): the assembly-order CSR of lhsa (solver/lhsa.cpp:153).  map(nNo):
 * assembly id -> solver id (lhs.map).  mynNo: rows counted in dots (lhs.mynNo).  Halo lists as in
 * lhs.cS[i]: req_rank[i] = iP, req_n[i] = n, req_ptr = the ptr arrays concatenated (solver ids). */
int b200_lhs_create(b200_handle* h, int gnNo, int nNo, int mynNo, int nnz,
                    const int* rowPtr, const int* colPtr, const int* map,
                    int nReq, const int* req_rank, const int* req_n, const int* req_ptr, int nFaces);
/* replaces lhs.face[faIn] as left by fsils_bc_create / fsils_bc_update (liner_solver/bc.cpp:45,171):
 * glob = solver ids, val(dof,nNo) already halo-summed when the face is shared. */
int b200_face_set(b200_handle* h, int faIn, int nNo, int dof, int bGrp, const int* glob,
                  const double* val, int shared);

/* ---- partition layout (replaces the renumbering of fsils_lhs_create, liner_solver/lhs.cpp:57-376) -------------------- */
/* Host-side, needs no device and no handle.  Input: every rank's global node ids in that rank's local (assembly) order --
 * what the MPI_Allgatherv at lhs.cpp:156 collects.  Output for `rank`: map (local assembly id -> solver id: nodes shared
 * with lower ranks first, interior, nodes shared with higher ranks last), mynNo, shnNo and one overlap list per neighbour
 * (solver ids, ordered as the higher rank of the pair walks its renumbered nodes) -- the map / mynNo / request arguments of
 * b200_lhs_create, integer for integer what the reference leaves in lhs.map / lhs.mynNo / lhs.shnNo / lhs.cS[].ptr.
 * Errors: non-zero return, text from b200_last_error(NULL). */
typedef struct b200_layout b200_layout;
int b200_lhs_layout_create(int rank, int nRanks, int gnNo, const int* counts, const int* const* gnodes, b200_layout** out);
int b200_lhs_layout_sizes(const b200_layout* lay, int* nNo, int* mynNo, int* shnNo, int* nReq);
int b200_lhs_layout_map(const b200_layout* lay, int* map /* nNo */);
/* i-th neighbour (ascending rank): peer, list length, and -- when ptr is not NULL -- the list */
int b200_lhs_layout_req(const b200_layout* lay, int i, int* peer, int* n, int* ptr);
void b200_lhs_layout_free(b200_layout* lay);
/* Element -> rank map by recursive coordinate bisection of the element centroids (3 x nEl): the stand-in for the reference's
 * ParMETIS call (solver/distribute.cpp:1683-1700) on meshes that are not generated slab by slab.  Balanced to one element,
 * deterministic, parts numbered along the cuts.  Host-side, no device. */
int b200_partition_rcb(int nEl, const double* centroids, int nParts, int* part /* nEl */);
/* The reference's own partition criterion: k-way partition of the mesh's DUAL graph, two elements being neighbours when they share
 * `ncommon` nodes (solver/distribute.cpp:1683-1706 passes eNoNb, the node count of a boundary element, to ParMETIS_V3_PartMeshKway
 * through split_).  IEN(eNoN,nEl) 0-based, element-major.  Serial METIS 5 (libmetis_static.a of the CUDA toolkit), host side, set-up
 * only; edgecut (may be NULL) = number of element faces between parts.  Fails when the library was not present at build time. */
int b200_partition_metis(int nEl, int eNoN, int nNo, const int* IEN, int ncommon, int nParts, int* part /* nEl */, long long* edgecut);

/* ---- prestress (com_mod.pS0 / pSn / pSa of the struct equation; solver/sv_struct.cpp:262-345, 646-700) ------------------ */
/* b200_prestress_set uploads the nodal prestress pS0(6,nNo) (rows 00 11 22 01 12 20, assembly node order; NULL: none) that
 * the next struct assemblies add to the 2nd Piola-Kirchhoff stress, and switches the pstEq accumulations on or off: with
 * pstEq != 0 every struct assembly also forms pSn(:,A) = sum w N_a pSl and pSa(A) = sum w N_a (pSl = the stress before the
 * prestress is added), which b200_prestress_get copies out - what construct_dsolid leaves in com_mod.pSn / pSa for pic::picc
 * (pic.cpp:208-219) to normalise.  Both start from zero at every assembly, as pic::pici zeroes them (pic.cpp:571). */
int b200_prestress_set(b200_handle* h, const double* pS0, int pstEq);
int b200_prestress_get(b200_handle* h, double* pSn /* 6 x nNo */, double* pSa /* nNo */);

/* ---- pattern (replaces lhsa_ns::lhsa, solver/lhsa.cpp:153, for idMap = identity, no shells) ------------------------- */
/* Device-side construction of the block-CSR pattern from the connectivity of every mesh of the equation system:
 * b200_pattern_begin(h, tnNo); b200_pattern_add_mesh(...) once per mesh (IEN(eNoN,nEl), assembly node ids);
 * b200_pattern_finish returns nnz; b200_pattern_get copies rowPtr(tnNo+1) / colPtr(nnz) to the host: sorted columns,
 * diagonal present, 0-based -- integer for integer what lhsa leaves in com_mod.rowPtr / colPtr.  Independent of
 * b200_lhs_create (it produces that call's inputs). */
int b200_pattern_begin(b200_handle* h, int tnNo);
int b200_pattern_add_mesh(b200_handle* h, int eNoN, int nEl, const int* IEN);
int b200_pattern_finish(b200_handle* h, int* nnz);
int b200_pattern_get(b200_handle* h, int* rowPtr, int* colPtr);

/* ---- assembly (replaces construct_fluid + do_assem, solver/fluid.cpp:464, lhsa.cpp:97) ------- */
/* IEN(eNoN,nEl) with assembly node ids, x(3,nNo).  eNoN = 4 (TET4), 8 (HEX8, the reference's node
 * order, nn_elem_gnn.h:732) or 10 (TET10, nn_elem_gnn.h:1256).  qmTET4 <= 0 selects the
 * default (5+3*sqrt(5))/20 (solver/ComMod.h:1011). */
int b200_mesh_set(b200_handle* h, int eNoN, int nEl, const int* IEN, const double* x, double qmTET4);
/* ls_alloc contract (solver/ls.cpp:51-60): after it R(dof,nNo) and Val(dof*dof,nnz) are zero. */
int b200_zero(b200_handle* h, int dof);
/* Upload Ag, Yg (tDof,nNo) and Bf (3,nNo; NULL = zero) for the next b200_assemble_fluid.  Ag = Yg = NULL keeps the
 * device copies (written by b200_pici) and uploads Bf only. */
int b200_state_set(b200_handle* h, int tDof, const double* Ag, const double* Yg, const double* Bf);
/* Whole-mesh fluid assembly on the device into R/Val, using the state uploaded last: TET4 (constant gradients),
 * HEX8 (nn::gnn per Gauss point) and TET10 (gnn + gn_nxx second derivatives, solver/nn.cpp:455,809), one function
 * space (VMS-stabilised equal order, lM.nFs = 1). */
int b200_assemble_fluid(b200_handle* h, const b200_fluid_props* p);
/* Upload Dg (tDof,nNo) and, for the mesh equation, Do (tDof,nNo; NULL otherwise). */
int b200_disp_set(b200_handle* h, int tDof, const double* Dg, const double* Do);
/* Whole-mesh solid assembly into R/Val (dof 3; b200_zero(h,3) first): replaces construct_dsolid +
 * struct_3d_carray + get_pk2cc (solver/sv_struct.cpp:213,552; mat_models_carray.h:182) ... */
int b200_assemble_struct(b200_handle* h, const b200_struct_props* p);
/* ... and construct_l_elas / construct_mesh + l_elas_3d (solver/l_elas.cpp:58,274; mesh.cpp:42). */
int b200_assemble_lelas(b200_handle* h, const b200_lelas_props* p);
/* ustruct equation (dof 4; b200_zero(h,4) first): replaces construct_usolid + ustruct_3d_m/c + ustruct_do_assem
 * (solver/ustruct.cpp:216,1158,632,1579) for equal-order TET4/HEX8/TET10 with idMap = identity; also fills the device
 * copy of com_mod.Kd(12,nnz). */
int b200_assemble_ustruct(b200_handle* h, const b200_ustruct_props* p);
/* ustruct_r (solver/ustruct.cpp:1726): R -= ami * Kd * (amg*Ad - Yg(s:s+2)) + overlap add; the caller invokes it on
 * the first Newton iteration only (eq.itr <= 1), like the reference.  Ad(3,nNo) host, assembly order. */
int b200_ustruct_r(b200_handle* h, double amg, double ami, int s, const double* Ad);
int b200_get_Kd(b200_handle* h, double* Kd);      /* parity tap: Kd(12,nnz), assembly layout */
/* lM.fN for nFn = 2: fN(6,nEl) = fibre and sheet direction of every element (solver/ComMod.h:975). */
int b200_mesh_fibers(b200_handle* h, int nFn, const double* fN);
/* FSI equation (solver/fsi.cpp:42-334: one element loop with a per-element domain switch).  elem_dmn[e] =
 * index of the equation domain element e belongs to (all_fun::domain, solver/all_fun.cpp:149), uploaded once;
 * the device keeps one element list per domain, so each domain is one divergence-free launch. */
int b200_mesh_domains(b200_handle* h, int nDmn, const int* elem_dmn);
/* Equations with several domains of ONE physics (eq.nDmn > 1, each eq.dmn[d] with its own properties; the element loop
 * of construct_fluid / construct_dsolid picks them per element, solver/fluid.cpp:531, sv_struct.cpp:261): p[d] are the
 * properties of domain d, the element lists come from b200_mesh_domains.  dof 4 / dof 3 systems as in the single-domain
 * calls. */
int b200_assemble_fluid_dmn(b200_handle* h, int nDmn, const b200_fluid_props* p);
int b200_assemble_struct_dmn(b200_handle* h, int nDmn, const b200_struct_props* p);
/* dmn_kind[d]: 0 fluid (fluid_3d_m/c on the ALE configuration x + Dg(4:6), mvMsh), 1 struct (struct_3d into the
 * 3x3 corner of the dof-4 blocks).  fluid[d] / solid[d] are read for the domains of that kind (dof 4; TET4, HEX8, TET10). */
int b200_assemble_fsi(b200_handle* h, int nDmn, const int* dmn_kind, const b200_fluid_props* fluid, const b200_struct_props* solid);
/* Boundary-face (Neumann) assembly on the device: replaces b_assem_neu_bc + gnnb + b_fluid / b_l_elas
 * (solver/eq_assem.cpp:58-170, nn.cpp:552-755, fluid.cpp:46-133, l_elas.cpp:48-59) and with them the per-element
 * LinearAlgebra::assemble calls of set_bc_neu_l.  b200_face_mesh_set: lFa.IEN(eNoNb,nElb) with assembly node ids and
 * lFa.gE(nElb) (parent element of every face element), eNoNb = 3 (TRI3), 4 (QUD4) or 6 (TRI6), uploaded once per face
 * (index faIn as in b200_face_set).  b200_assemble_bneu adds the face's contribution to the device R / Val AFTER the
 * volume assembly, face elements in order (do_assem's order): kind 0 = b_fluid (traction h n + backflow stabilisation,
 * residual + tangent; dof 4), kind 1 = b_l_elas (traction, residual only; dof 3 or 4).  hg(nNo): the nodal Neumann
 * values set_bc_neu_l passes (read on the face nodes only); the velocity comes from the device Yg (b200_state_set /
 * b200_pici), the moving-mesh geometry from Do (b200_disp_set). */
typedef struct { double dt, af, gam; int tDof, mvMsh; double rho, bfs; } b200_bneu_props;
int b200_face_mesh_set(b200_handle* h, int faIn, int eNoNb, int nElb, const int* IENb, const int* gE);
/* all_fun::integ over face faIn (solver/all_fun.cpp:561,724,858) of rows l..u of a device-resident array: `which` is a
 * B200_PIC_* id (Yo, Yn, ... of the time integrator; Ag/Yg/Dg state) or -1 for the integrand 1 (the face area).
 * u - l + 1 == 3: flux  sum w N s.n  with the area-weighted normal; l == u: scalar  sum |n| w N s.  geo: configuration of the
 * normals as in gnnb (nn.cpp:609-640): 0 reference x, 1 x + Do(0:2) (old time step), 2 x + Dn(0:2) (new), 3 moving mesh
 * x + Do(4:6).  The sum runs serially over (element, Gauss point), the reference's order.  In a multi-rank
 * run this is the rank's part; the caller reduces it like cm.reduce does. */
int b200_face_integ(b200_handle* h, int faIn, int which, int l, int u, int geo, double* result);
int b200_assemble_bneu(b200_handle* h, int faIn, int kind, const b200_bneu_props* p, const double* hg);
/* Follower pressure load on a struct or ustruct face (lBc.flwP): replaces eq_assem::b_neu_folw_p (solver/eq_assem.cpp:186-303)
 * with nn::get_nnx / get_xi (Newton inverse map of the face Gauss points into the parent element, nn.cpp:314-440), gnn, gnnb
 * and struct_ns::b_struct_3d (Nanson's formula; residual and tangent, sv_struct.cpp:116-210; dof-3 system, tangent factor
 * af*beta*dt^2) or, ustruct != 0, ustruct::b_ustruct_3d + ustruct_do_assem (ustruct.cpp:132-211, 1579: dof-4 system, the
 * tangent goes to Kd and, scaled by af*gam*dt/am, to the velocity block of Val).  Displacement from the device Dg
 * (b200_disp_set / b200_pici), rows s..s+2; parents TET4 / HEX8 / TET10.  Fails like the reference when the inverse map
 * does not converge ("Error in computing shape functions"). */
typedef struct { double dt, af, beta; int tDof, s; int ustruct; double am, gam; } b200_bfolw_props;
int b200_assemble_bfolw(b200_handle* h, int faIn, const b200_bfolw_props* p, const double* hg);
/* eq_assem::fsi_ls_upd (solver/eq_assem.cpp:316-371) + fsils_bc_update (liner_solver/bc.cpp:171): recompute the vector
 * val(i,a) = int N_a n_i dGamma of the coupled Neumann face lsFace (index of b200_face_set, dof 3) from the face mesh
 * faIn on the configuration geo (as in b200_face_integ; the reference uses 2 = new time step), halo-summed when the face
 * is shared between ranks.  The face vectors of the linear solve then follow the moving mesh without a host round trip. */
int b200_face_normal_update(b200_handle* h, int faIn, int lsFace, int geo);
int b200_face_get_val(b200_handle* h, int lsFace, double* val);            /* parity tap: lhs.face[lsFace].val(dof,nNo) */
/* LinearAlgebra::assemble for the few boundary-face elements: staged on the host, flushed by one
 * scatter kernel before the next get/solve.  eqN(d), lK(dof*dof,d,d), lR(dof,d). */
int b200_assemble_elem(b200_handle* h, int d, const int* eqN, const double* lK, const double* lR);
/* parity taps / host boundary-condition code (assembly ordering, like com_mod.R / com_mod.Val). */
int b200_get_R(b200_handle* h, double* R);
int b200_set_R(b200_handle* h, int dof, const double* R);
/* device R += host R (what host boundary-condition code added to com_mod.R since ls_alloc). */
int b200_add_R(b200_handle* h, int dof, const double* R);
int b200_get_Val(b200_handle* h, double* Val);
int b200_set_Val(b200_handle* h, int dof, const double* Val);
/* all_fun::commu(R) (solver/all_fun.cpp:122): overlap-node add of the device R. */
int b200_commu_R(b200_handle* h);

/* ---- time integrator on the device (replaces pic::picp / pici / picc, solver/pic.cpp:591,486,74) ---------- */
/* With the generalised-alpha state resident on the device a Newton iteration needs no upload of Ag/Yg/Dg and no
 * download of the solution: b200_pici writes the state the assembly kernels read (what b200_state_set / b200_disp_set
 * upload otherwise), b200_picc consumes the device R left by b200_solve.  Arrays are (tDof,nNo) in assembly order like
 * com_mod.Ao..Dn; Ad is com_mod.Ad(3,nNo).  kind = what picc does for the equation (pic.cpp:116-160): 0 the general
 * update of An, Yn, Dn (every equation when !sstEq; the mesh equation under sstEq), 1 ustruct / FSI under sstEq (An,
 * Yn from R; Ad, Dn through Rd), 2 nothing (any other equation under sstEq). */
typedef struct { int s, e; double am, af, gam, beta; int kind; } b200_pic_eq;
enum { B200_PIC_AO = 0, B200_PIC_YO, B200_PIC_DO, B200_PIC_AN, B200_PIC_YN, B200_PIC_DN, B200_PIC_AD,
       B200_PIC_AG, B200_PIC_YG, B200_PIC_DG };
/* dFlag, sstEq: com_mod.dFlag / com_mod.sstEq (they select the displacement predictor, pic.cpp:690-712). */
int b200_pic_init(b200_handle* h, int tDof, int nEq, const b200_pic_eq* eqs, int dFlag, int sstEq);
int b200_pic_set(b200_handle* h, int which, const double* a);        /* upload one array */
int b200_pic_get(b200_handle* h, int which, double* a);              /* download one array */
/* set_bc_dir (solver/set_bc.cpp:794) on the device copy: arr[idx[k]] = val[k], idx = i + tDof*a. */
int b200_pic_scatter(b200_handle* h, int which, int n, const int* idx, const double* val);
int b200_picp(b200_handle* h, double dt);                            /* predictor, every equation */
int b200_pici(b200_handle* h);                                       /* Ag, Yg, Dg of every equation */
/* corrector for equation iEq from the device solution; first_itr: eq.itr == 1 (ustruct_r's Rd, kind 1 only). */
int b200_picc(b200_handle* h, int iEq, double dt, int first_itr);
/* FSI tail of picc (pic.cpp:166-181): on the listed solid-domain nodes copy rows [0,cnt) of An/Yn/Dn to [s2,s2+cnt). */
int b200_pic_copy_rows(b200_handle* h, int n, const int* nodes, int s2, int cnt);
int b200_pic_advance(b200_handle* h);                                /* end of time step: Ao = An, Yo = Yn, Do = Dn */

/* ---- solve (replaces fsils_solve, liner_solver/solve.cpp:50) ------------------------------- */
/* Consumes the device R/Val (Val is scaled in place like the reference), leaves the solution in the
 * device R and, when R_out != NULL, copies it to the host (dof,nNo, assembly order).  GM/CG are
 * read for ls_type == B200_LS_NS only.  incL/res: nFaces entries or NULL. */
int b200_solve(b200_handle* h, int ls_type, int prec, const b200_tol* RI, const b200_tol* GM,
               const b200_tol* CG, const int* incL, const double* res, double* R_out, b200_ls_out* out);

/* ---- single-kernel taps used by the parity tests and the roofline bench --------------------- */
/* y = K x (+ overlap add) with the device Val; x, y host (dof,nNo) in assembly order. */
int b200_spmv(b200_handle* h, int dof, const double* x, double* y);
/* Stand-alone kernel bench: times `reps` back-to-back launches of ONE kernel class (ids as in
 * b200_profile_read) on device-resident data with CUDA events on the launch stream, after 3 warm-up
 * launches (none when reps == 1, for profiler captures); returns mean milliseconds and the ALGORITHMIC bytes per launch.  Needs an assembled dof-4
 * system.  NS shapes (1-4, 9) run on the departed matrix; 3 / 4 are the two passes of the fused Schur
 * operator; k = number of basis vectors for multi_dot (5) and cgs_update_scale (6) on dof-3 vectors; for the SpMV shapes k selects
 * the kernel variant (see b200_tune).  op 100 = the FP64 FMA peak micro-kernel (needs no system): k x 1024 dependent DFMAs in 8
 * independent chains per thread, `bytes_per_launch` then returns the FLOPs of one launch - the roofline denominator of the element
 * kernels. */
int b200_op_bench(b200_handle* h, int op, int k, int reps, double* ms_per_launch, double* bytes_per_launch);
/* Kernel-launch counter (all kernels this handle launched since creation). */
long long b200_launch_count(b200_handle* h);
/* Kernel-variant knobs for A/B measurements and the parity tests of every variant (all variants compute the same products; they
 * differ in how the matrix stream reaches the lanes): "vv3" 0 lane = component / 1 lanes stride over the row's blocks / 2 TMA-staged
 * row tiles; "schur_gp", "schur_sp" 0 / 1 (two / four entries in flight) / 2 TMA-staged; "narrow" (spmv_ss/sv/vs) 0 / 2;
 * "cg_batch" iterations enqueued per host poll of the device-resident CG loops; "fused" 1 (default) / 0: with the peer transport, the
 * partitioned block products run as ONE cooperative kernel each (rows + overlap exchange + ordered add, csrc/fused_halo.cuh) or as the
 * four launches boundary rows / push / interior rows / wait-add; "gmres_device" 1 (default) / 0: the Arnoldi loops keep their Givens
 * bookkeeping and convergence test on the device (the host polls one batch behind) or fetch the reduced dots every iteration. */
int b200_tune(b200_handle* h, const char* name, int value);
/* Live kernel timing inside a step.  b200_profile(h,1) resets the counters and makes every kernel
 * class record CUDA-event pairs on the launch stream; b200_profile_read synchronises and returns,
 * per class, the summed device milliseconds, the ALGORITHMIC bytes (SURVEY.md par. 8d formulas) and the
 * number of launches.  Classes: 0 spmv_vv dof4, 1 spmv_vv dof<4, 2 spmv_ss, 3 spmv_sv, 4 spmv_vs,
 * 5 multi_dot, 6 cgs_update_scale, 7 blas1 (axpy/scale/lin_comb/...), 8 scale_val (Jacobi), 9 depart,
 * 10 assembly (all colours), 11 halo pack/add.  enable: 0 off, 1 every class, 2 + c only class c. */
#define B200_NUM_KERNEL_CLASSES 12
int b200_profile(b200_handle* h, int enable);
int b200_profile_read(b200_handle* h, int max_classes, double* ms, double* bytes, long long* launches);
/* Device-side stopwatch on the launch stream: b200_timer(h,0,NULL) records the start event,
 * b200_timer(h,1,&ms) records + synchronises the stop event and returns the elapsed milliseconds. */
int b200_timer(b200_handle* h, int stop, double* ms);
/* Milliseconds spent (CUDA events) in the phases of the last b200_assemble_fluid / b200_solve:
 * t[0] assembly, t[1] preconditioning, t[2] Krylov, t[3] SpMV share of t[2] (0 unless profiling on). */
int b200_last_timings(b200_handle* h, double* t4);

#ifdef __cplusplus
}
#endif
#endif /* SVB200_H */
