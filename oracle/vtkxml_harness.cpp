// TEST INFRASTRUCTURE ONLY.  Runs the reference's OWN, UNMODIFIED vtk_xml.cpp (Code/Source/solver/vtk_xml.cpp: read_vtu :568,
// read_vtp :438, read_vtu_pdata :667, write_vtu :855, write_vtp :827) with its two VTK-bound translation units replaced by the
// product's VTK-free ones (svfsiplus_b200/host/VtkDataB200.cpp for VtkData.cpp, vtk_xml_parser_b200.cpp for vtk_xml_parser.cpp).
// The functions below only fill / read the reference's mshType / faceType / ComMod.
#include "ComMod.h"
#include "Simulation.h"
#include "vtk_xml.h"
#include "vtk_xml_parser.h"
#include "Parameters.h"
#include "read_msh.h"

#include <unistd.h>

#include <cstring>
#include <stdexcept>
#include <string>

// write_vtus' post-processing calls are outside what is driven here; they only have to link (vtk_xml.cpp references them)
namespace post {
void fib_dir_post(Simulation*, const mshType&, const int, Array<double>&, const Array<double>&, const int) { throw std::runtime_error("[oracle] post:: is outside the I/O path"); }
void fib_algn_post(Simulation*, const mshType&, Array<double>&, const Array<double>&, const int) { throw std::runtime_error("[oracle] post:: is outside the I/O path"); }
void post(Simulation*, const mshType&, Array<double>&, const Array<double>&, const Array<double>&, consts::OutputNameType, const int) { throw std::runtime_error("[oracle] post:: is outside the I/O path"); }
void bpost(Simulation*, const mshType&, Array<double>&, const Array<double>&, const Array<double>&, consts::OutputNameType) { throw std::runtime_error("[oracle] post:: is outside the I/O path"); }
void tpost(Simulation*, const mshType&, const int, Array<double>&, Vector<double>&, const Array<double>&, const Array<double>&, const int, consts::OutputNameType) { throw std::runtime_error("[oracle] post:: is outside the I/O path"); }
void div_post(Simulation*, const mshType&, Array<double>&, const Array<double>&, const Array<double>&, const int) { throw std::runtime_error("[oracle] post:: is outside the I/O path"); }
void shl_post(Simulation*, const mshType&, const int, Array<double>&, Vector<double>&, const Array<double>&, const int, consts::OutputNameType) { throw std::runtime_error("[oracle] post:: is outside the I/O path"); }
}

namespace { std::string g_err; }

extern "C" {

const char* vx_last_error() { return g_err.c_str(); }

// vtk_xml::read_vtu into a fresh mshType.  sizes = {gnNo, gnEl, eNoN, len(gN), number of rows of mesh.ordering, its row length};
// null outputs: sizes only.
int vx_read_vtu(const char* path, int* sizes, double* x, int* gIEN, int* gN, int* ordering)
{
  try {
    mshType mesh;
    mesh.name = "msh";
    vtk_xml::read_vtu(path, mesh);
    sizes[0] = mesh.gnNo; sizes[1] = mesh.gnEl; sizes[2] = mesh.eNoN; sizes[3] = mesh.gN.size();
    sizes[4] = int(mesh.ordering.size()); sizes[5] = mesh.ordering.empty() ? 0 : int(mesh.ordering[0].size());
    if (!x) return 0;
    std::memcpy(x, mesh.x.data(), sizeof(double)*3*size_t(mesh.gnNo));
    std::memcpy(gIEN, mesh.gIEN.data(), sizeof(int)*size_t(mesh.eNoN)*mesh.gnEl);
    if (mesh.gN.size()) std::memcpy(gN, mesh.gN.data(), sizeof(int)*size_t(mesh.gN.size()));
    for (size_t f = 0; f < mesh.ordering.size(); f++)
      for (size_t k = 0; k < mesh.ordering[f].size(); k++) ordering[f*mesh.ordering[0].size() + k] = mesh.ordering[f][k];
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// vtk_xml::read_vtp into a fresh faceType.  sizes = {nNo, nEl, eNoN, len(gN), len(gE)}; gebc is (eNoN+1) x nEl.
int vx_read_vtp(const char* path, int* sizes, double* x, int* IEN, int* gN, int* gE, int* gebc)
{
  try {
    faceType face;
    face.name = "face";
    vtk_xml::read_vtp(path, face);
    sizes[0] = face.nNo; sizes[1] = face.nEl; sizes[2] = face.eNoN; sizes[3] = face.gN.size(); sizes[4] = face.gE.size();
    if (!x) return 0;
    std::memcpy(x, face.x.data(), sizeof(double)*3*size_t(face.nNo));
    std::memcpy(IEN, face.IEN.data(), sizeof(int)*size_t(face.eNoN)*face.nEl);
    if (face.gN.size()) std::memcpy(gN, face.gN.data(), sizeof(int)*size_t(face.gN.size()));
    if (face.gE.size()) {
      std::memcpy(gE, face.gE.data(), sizeof(int)*size_t(face.gE.size()));
      std::memcpy(gebc, face.gebc.data(), sizeof(int)*size_t(face.eNoN + 1)*face.nEl);
    }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// vtk_xml::write_vtu (points lM.x, connectivity lM.gIEN) and vtk_xml::write_vtp (lFa.x, lFa.IEN, GlobalNodeID, GlobalElementID).
int vx_write_vtu(const char* path, int nNo, const double* x, int eNoN, int nEl, const int* gIEN)
{
  try {
    ComMod com_mod;
    com_mod.nsd = 3;
    mshType lM;
    lM.x.resize(3, nNo);
    std::memcpy(lM.x.data(), x, sizeof(double)*3*size_t(nNo));
    lM.gIEN.resize(eNoN, nEl);
    std::memcpy(lM.gIEN.data(), gIEN, sizeof(int)*size_t(eNoN)*nEl);
    vtk_xml::write_vtu(com_mod, lM, path);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int vx_write_vtp(const char* path, int nNo, const double* x, int eNoN, int nEl, const int* IEN, const int* gN, const int* gE)
{
  try {
    ComMod com_mod;
    com_mod.nsd = 3;
    faceType lFa;
    lFa.x.resize(3, nNo);
    std::memcpy(lFa.x.data(), x, sizeof(double)*3*size_t(nNo));
    lFa.IEN.resize(eNoN, nEl);
    std::memcpy(lFa.IEN.data(), IEN, sizeof(int)*size_t(eNoN)*nEl);
    if (gN) { lFa.gN.resize(nNo); std::memcpy(lFa.gN.data(), gN, sizeof(int)*size_t(nNo)); }
    if (gE) { lFa.gE.resize(nEl); std::memcpy(lFa.gE.data(), gE, sizeof(int)*size_t(nEl)); }
    vtk_xml::write_vtp(com_mod, lFa, path);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// vtk_xml::read_vtu_pdata with m != nsd: the whole named point array into mesh.x (m x gnNo), as the prestress / fibre readers use it.
int vx_read_vtu_pdata(const char* path, const char* kwrd, int m, int gnNo, double* out)
{
  try {
    mshType mesh;
    mesh.name = "msh";
    mesh.gnNo = gnNo;
    mesh.x.resize(m, gnNo);
    vtk_xml::read_vtu_pdata(path, kwrd, 3, m, 0, mesh);
    std::memcpy(out, mesh.x.data(), sizeof(double)*size_t(m)*gnNo);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// vtk_xml_parser::load_fiber_direction_vtu (cell array -> mesh.fN rows idx*3..) and load_time_varying_field_vtu (-> mesh.Ys).
int vx_load_fibers(const char* path, const char* name, int idx, int nFn, int gnEl, double* fN)
{
  try {
    mshType mesh;
    mesh.name = "msh";
    mesh.gnEl = gnEl;
    mesh.fN.resize(3*nFn, gnEl);
    vtk_xml_parser::load_fiber_direction_vtu(path, name, idx, 3, mesh);
    std::memcpy(fN, mesh.fN.data(), sizeof(double)*3*size_t(nFn)*gnEl);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int vx_load_time_field(const char* path, const char* field, int* dims3, double* Ys, int cap)
{
  try {
    mshType mesh;
    mesh.name = "msh";
    vtk_xml_parser::load_time_varying_field_vtu(path, field, mesh);
    dims3[0] = mesh.Ys.nrows(); dims3[1] = mesh.Ys.ncols(); dims3[2] = mesh.Ys.nslices();
    const size_t n = size_t(dims3[0])*dims3[1]*dims3[2];
    if (Ys && n <= size_t(cap)) std::memcpy(Ys, mesh.Ys.data(), sizeof(double)*n);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// Parameters::read_xml (Code/Source/solver/Parameters.cpp:147): the reference's own parser on a solver.xml.  out = {time steps,
// meshes, faces of mesh 0, equations, BCs of equation 0, LS max_iterations, NS_GM max, NS_CG max, Krylov dimension};
// dout = {dt, density, Newton tolerance, LS tolerance, LS absolute tolerance}; ls_type / la_type: the <LS type> and <Linear_algebra type>.
int vx_parse_solver_xml(const char* path, int* out, double* dout, char* ls_type, char* la_type, int cap)
{
  try {
    Parameters params;
    params.read_xml(path);
    out[0] = params.general_simulation_parameters.number_of_time_steps();
    out[1] = int(params.mesh_parameters.size());
    out[2] = params.mesh_parameters.empty() ? 0 : int(params.mesh_parameters[0]->face_parameters.size());
    out[3] = int(params.equation_parameters.size());
    auto* eq = params.equation_parameters.at(0);
    out[4] = int(eq->boundary_conditions.size());
    out[5] = eq->linear_solver.max_iterations();
    out[6] = eq->linear_solver.ns_gm_max_iterations();
    out[7] = eq->linear_solver.ns_cg_max_iterations();
    out[8] = eq->linear_solver.krylov_space_dimension();
    dout[0] = params.general_simulation_parameters.time_step_size();
    dout[1] = eq->default_domain ? eq->default_domain->density() : 0.0;      // <Density> of the equation (its default domain)
    dout[2] = eq->tolerance();
    dout[3] = eq->linear_solver.tolerance();
    dout[4] = eq->linear_solver.absolute_tolerance();
    snprintf(ls_type, size_t(cap), "%s", eq->linear_solver.type().c_str());
    snprintf(la_type, size_t(cap), "%s", eq->linear_solver.linear_algebra.type().c_str());
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// The reference's mesh ingestion on a case directory: Simulation::read_parameters + set_module_parameters + read_msh_ns::read_msh
// (Code/Source/solver/read_files.cpp:1649-1652, read_msh.cpp:1077: read_sv -> read_vtu / read_vtp for the mesh and every face, face /
// element matching, check_ien, global coordinates) - all unmodified reference code, running on the VTK-free replacements.
// sizes = {nsd, nMsh, gtnNo, gnEl, eNoN, nFa}; face_sizes = nFa x {nNo, nEl, eNoN}.  x (3 x gtnNo) and gIEN (eNoN x gnEl) may be null.
int vx_read_case(const char* dir, const char* xml, int* sizes, int* face_sizes, int max_faces, double* x, int* gIEN)
{
  try {
    if (chdir(dir) != 0) throw std::runtime_error(std::string("cannot enter '") + dir + "'");
    Simulation sim;
    sim.read_parameters(xml);
    sim.set_module_parameters();
    read_msh_ns::read_msh(&sim);
    auto& com_mod = sim.com_mod;
    auto& msh = com_mod.msh.at(0);
    sizes[0] = com_mod.nsd; sizes[1] = com_mod.nMsh; sizes[2] = com_mod.gtnNo; sizes[3] = msh.gnEl; sizes[4] = msh.eNoN; sizes[5] = msh.nFa;
    for (int i = 0; i < msh.nFa && i < max_faces; i++) {
      face_sizes[3*i] = msh.fa[i].nNo; face_sizes[3*i + 1] = msh.fa[i].nEl; face_sizes[3*i + 2] = msh.fa[i].eNoN;
    }
    if (x) std::memcpy(x, com_mod.x.data(), sizeof(double)*size_t(com_mod.nsd)*com_mod.gtnNo);
    if (gIEN) std::memcpy(gIEN, msh.gIEN.data(), sizeof(int)*size_t(msh.eNoN)*msh.gnEl);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

} // extern "C"
