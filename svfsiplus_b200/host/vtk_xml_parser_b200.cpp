// vtk_xml_parser_b200.cpp — the reference's vtk_xml_parser namespace (Code/Source/solver/vtk_xml_parser.h:36-62: load_vtu / load_vtp
// for meshes and faces, load_fiber_direction_vtu, load_time_varying_field_vtu) on the VTK-free I/O library (include/svb200_io.h).
// It REPLACES Code/Source/solver/vtk_xml_parser.cpp in a build without VTK, next to VtkDataB200.cpp (which replaces VtkData.cpp):
// compiled against the unmodified vtk_xml_parser.h / ComMod.h, so the reference's own vtk_xml.cpp (read_vtu :568, read_vtp :438,
// read_vtu_pdata :667, ...) links against it unchanged.
//
// What the reference's functions leave in mshType / faceType is reproduced field by field (vtk_xml_parser.cpp, file:line):
//   mesh <- .vtu  (:794-835)  gnNo, x(3,gnNo), gN = GlobalNodeID as stored (:557-571), gnEl, eNoN, gIEN(eNoN,gnEl), ordering = the
//                             face-node table of the mesh's VTK cell type (:83-147, 149-292, 393-447)
//   mesh <- .vtp  (:742-792)  gnNo, x, gN = GlobalNodeID - 1 (:609-623), gnEl, eNoN, gIEN (:351-391); no ordering
//   face <- .vtp  (:699-740)  nNo, x, gN = GlobalNodeID - 1 (:594-607), nEl, eNoN, IEN (:294-327), gE = GlobalElementID - 1 (:466-488)
//   face <- .vtu  (:837-886)  the same with gN = GlobalNodeID as stored (:573-592)
//   missing GlobalNodeID: gN stays empty; missing GlobalElementID on a face: "No 'GlobalElementID' data of type Int32 found in VTK
//   mesh."; a file without points: "Failed reading the VTK file '<name>'."
//   fibres (:643-697): cell array `data_name` -> mesh.fN(i + idx*nsd, e); element-count and missing-array messages as the reference
//   time-varying field (:888-950): every point array whose name contains `field_name`, ordered by its trailing number, -> mesh.Ys
// Difference, on purpose: the id arrays may be of any integer type (the reference only down-casts vtkIntArray and treats an Int64
// GlobalNodeID as missing).
#include "vtk_xml_parser.h"

#include "svb200_io.h"

#include <algorithm>
#include <cctype>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace vtk_xml_parser {

const std::string VtkFileExtentions::VTK_VTU_EXTENSION = "vtu";
const std::string VtkFileExtentions::VTK_VTP_EXTENSION = "vtp";

namespace {

const std::string NODE_IDS_NAME("GlobalNodeID");
const std::string ELEMENT_IDS_NAME("GlobalElementID");

// nodes of each face of a VTK cell, in VTK's node numbering: what mshType::ordering holds (used to match boundary faces to their
// parent elements).  One row per face; for the quadratic cells the mid-side (and mid-face) nodes follow the corners' cycle.
const std::map<int, std::vector<std::vector<int>>>& face_tables()
{
  static const std::map<int, std::vector<std::vector<int>>> t = {
    {B200IO_VTK_LINE, {{0}, {1}}},
    {B200IO_VTK_TRIANGLE, {{0, 1}, {1, 2}, {2, 0}}},
    {B200IO_VTK_QUAD, {{0, 1}, {1, 2}, {2, 3}, {3, 0}}},
    {B200IO_VTK_TETRA, {{0, 1, 2}, {0, 1, 3}, {1, 2, 3}, {2, 0, 3}}},
    {B200IO_VTK_HEXAHEDRON, {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}}},
    {B200IO_VTK_WEDGE, {{0, 1, 2}, {3, 4, 5}, {0, 1, 4, 3}, {1, 2, 5, 4}, {2, 0, 3, 5}}},
    {B200IO_VTK_QUADRATIC_TRIANGLE, {{0, 3, 1}, {1, 4, 2}, {2, 5, 0}}},
    {34 /* VTK_BIQUADRATIC_TRIANGLE */, {{0, 3, 1}, {1, 4, 2}, {2, 5, 0}}},
    {B200IO_VTK_QUADRATIC_QUAD, {{0, 4, 1}, {1, 5, 2}, {2, 6, 3}, {3, 7, 0}}},
    {B200IO_VTK_BIQUADRATIC_QUAD, {{0, 4, 1}, {1, 5, 2}, {2, 6, 3}, {3, 7, 0}}},
    {B200IO_VTK_QUADRATIC_TETRA, {{0, 1, 2, 4, 5, 6}, {0, 3, 1, 7, 8, 4}, {1, 3, 2, 8, 9, 5}, {2, 3, 0, 9, 7, 6}}},
    {B200IO_VTK_QUADRATIC_HEXAHEDRON, {{0, 11, 3, 10, 2, 9, 1, 8}, {4, 12, 5, 13, 6, 14, 7, 15}, {0, 8, 1, 17, 5, 12, 4, 16},
                                       {1, 9, 2, 18, 6, 13, 5, 17}, {2, 10, 3, 19, 7, 14, 6, 18}, {3, 11, 0, 16, 4, 15, 7, 19}}},
    {B200IO_VTK_TRIQUADRATIC_HEXAHEDRON, {{0, 11, 3, 10, 2, 9, 1, 8, 24}, {4, 12, 5, 13, 6, 14, 7, 15, 25}, {0, 8, 1, 17, 5, 12, 4, 16, 22},
                                          {1, 9, 2, 18, 6, 13, 5, 17, 21}, {2, 10, 3, 19, 7, 14, 6, 18, 23}, {3, 11, 0, 16, 4, 15, 7, 19, 20}}},
  };
  return t;
}

// The reference scans the cell types in this order and lets the LAST type present win (get_mesh_ordering, :149-292).
const int kTypeScan[] = {B200IO_VTK_LINE, B200IO_VTK_HEXAHEDRON, B200IO_VTK_QUAD, B200IO_VTK_TETRA, B200IO_VTK_TRIANGLE, B200IO_VTK_WEDGE,
                         B200IO_VTK_QUADRATIC_TRIANGLE, 34, B200IO_VTK_QUADRATIC_QUAD, B200IO_VTK_BIQUADRATIC_QUAD, B200IO_VTK_QUADRATIC_TETRA,
                         B200IO_VTK_QUADRATIC_HEXAHEDRON, B200IO_VTK_TRIQUADRATIC_HEXAHEDRON};

struct File {
  b200io_vtk* h = nullptr;
  std::string name;
  explicit File(const std::string& file_name) : name(file_name)
  {
    if (b200io_vtk_read(file_name.c_str(), &h) != 0 || b200io_vtk_num_points(h) == 0) {
      if (h) { b200io_vtk_free(h); h = nullptr; }
      throw std::runtime_error("Failed reading the VTK file '" + file_name + "'.");
    }
  }
  ~File() { if (h) b200io_vtk_free(h); }
  File(const File&) = delete;
  File& operator=(const File&) = delete;

  int num_points() const { return b200io_vtk_num_points(h); }
  int num_elems() const { return b200io_vtk_num_cells(h); }
  Array<double> coords() const
  {
    Array<double> x(3, num_points());
    b200io_vtk_points(h, x.data());
    return x;
  }
  // np_elem of the first cell, as the reference takes it (GetCell(0)); every cell must have it
  Array<int> conn(int& np_elem, const std::string& what) const
  {
    np_elem = b200io_vtk_nodes_per_cell(h);
    if (num_elems() > 0 && np_elem <= 0) throw std::runtime_error("[store_element_conn] Error in VTK mesh data for mesh '" + what + "'.");
    if (np_elem < 0) np_elem = 0;
    Array<int> ien(np_elem, num_elems());
    if (num_elems() > 0) b200io_vtk_connectivity(h, ien.data());
    return ien;
  }
  bool ids(int where, const std::string& name_, int shift, Vector<int>& out) const
  {
    int nc = 0, nt = 0;
    if (b200io_vtk_array_info(h, where, name_.c_str(), &nc, &nt, nullptr) != 0) return false;
    std::vector<int> tmp(size_t(nc)*nt);
    if (b200io_vtk_array_i32(h, where, name_.c_str(), tmp.data()) != 0) throw std::runtime_error(std::string("[vtk_xml_parser/b200io] ") + b200io_last_error());
    out = Vector<int>(nt);
    for (int i = 0; i < nt; i++) out(i) = tmp[size_t(i)*nc] + shift;
    return true;
  }
};

void element_ids(const File& f, faceType& face)
{
  if (!f.ids(B200IO_CELL_DATA, ELEMENT_IDS_NAME, -1, face.gE))
    throw std::runtime_error("No '" + ELEMENT_IDS_NAME + "' data of type Int32 found in VTK mesh.");
}

void load_face(const std::string& file_name, faceType& face, int node_id_shift)
{
  File f(file_name);
  face.nNo = f.num_points();
  face.x = f.coords();
  f.ids(B200IO_POINT_DATA, NODE_IDS_NAME, node_id_shift, face.gN);
  int np = 0;
  face.IEN = f.conn(np, face.name);
  face.nEl = f.num_elems();
  face.eNoN = np;
  element_ids(f, face);
}

} // namespace

void load_vtp(const std::string& file_name, faceType& face) { load_face(file_name, face, -1); }
void load_vtu(const std::string& file_name, faceType& face) { load_face(file_name, face, 0); }

void load_vtp(const std::string& file_name, mshType& mesh)
{
  File f(file_name);
  mesh.gnNo = f.num_points();
  mesh.x = f.coords();
  f.ids(B200IO_POINT_DATA, NODE_IDS_NAME, -1, mesh.gN);
  int np = 0;
  mesh.gIEN = f.conn(np, mesh.name);
  mesh.gnEl = f.num_elems();
  mesh.eNoN = np;
}

void load_vtu(const std::string& file_name, mshType& mesh)
{
  File f(file_name);
  mesh.gnNo = f.num_points();
  mesh.x = f.coords();
  f.ids(B200IO_POINT_DATA, NODE_IDS_NAME, 0, mesh.gN);
  int np = 0;
  mesh.gIEN = f.conn(np, mesh.name);
  mesh.gnEl = f.num_elems();
  mesh.eNoN = np;
  // the face-node table of the cell type (the last one of the reference's scan order that occurs)
  std::vector<unsigned char> types(size_t(std::max(mesh.gnEl, 1)));
  if (mesh.gnEl > 0) b200io_vtk_cell_types(f.h, types.data());
  std::vector<char> present(256, 0);
  for (int e = 0; e < mesh.gnEl; e++) present[types[e]] = 1;
  mesh.ordering.clear();
  for (int t : kTypeScan) if (present[t]) mesh.ordering = face_tables().at(t);
}

void load_fiber_direction_vtu(const std::string& file_name, const std::string& data_name, const int idx, const int nsd, mshType& mesh)
{
  if (FILE* file = fopen(file_name.c_str(), "r")) fclose(file);
  else throw std::runtime_error("The fiber direction VTK file '" + file_name + "' can't be read.");
  File f(file_name);
  const int num_elems = f.num_elems();
  if (mesh.gnEl != num_elems) {
    throw std::runtime_error("The number of elements (" + std::to_string(num_elems) + ") in the fiber direction VTK file '" + file_name +
        "' is not equal to the number of elements (" + std::to_string(mesh.gnEl) + ") for the mesh named '" + mesh.name + "'.");
  }
  int nc = 0, nt = 0, is_int = 0;
  if (b200io_vtk_array_info(f.h, B200IO_CELL_DATA, data_name.c_str(), &nc, &nt, &is_int) != 0 || is_int)
    throw std::runtime_error("No '" + data_name + "' data found in the fiber direction VTK file '" + file_name + "'");
  std::vector<double> tmp(size_t(nc)*nt);
  b200io_vtk_array_f64(f.h, B200IO_CELL_DATA, data_name.c_str(), tmp.data());
  const int offset = idx*nsd;
  for (int e = 0; e < mesh.gnEl; e++)
    for (int i = 0; i < nsd && i < nc; i++) mesh.fN(i + offset, e) = tmp[size_t(e)*nc + i];
}

void load_time_varying_field_vtu(const std::string file_name, const std::string field_name, mshType& mesh)
{
  File f(file_name);
  const int num_nodes = f.num_points();
  std::vector<std::pair<std::string, int>> names;
  const int n_arrays = b200io_vtk_num_arrays(f.h, B200IO_POINT_DATA);
  for (int i = 0; i < n_arrays; i++) {
    const std::string name = b200io_vtk_array_name(f.h, B200IO_POINT_DATA, i);
    if (name.find(field_name) == std::string::npos) continue;
    auto it = std::find_if(name.rbegin(), name.rend(), [](char c) { return !std::isdigit((unsigned char)c); });
    const std::string step(it.base(), name.end());
    names.push_back({name, step.empty() ? 0 : std::stoi(step)});
  }
  if (names.empty()) throw std::runtime_error("No '" + field_name + "' data found in the VTK file '" + file_name + "'.");
  std::sort(names.begin(), names.end(), [](const std::pair<std::string, int>& a, const std::pair<std::string, int>& b) { return a.second < b.second; });
  int ncomp = 0;
  b200io_vtk_array_info(f.h, B200IO_POINT_DATA, names[0].first.c_str(), &ncomp, nullptr, nullptr);
  mesh.Ys.resize(ncomp, num_nodes, int(names.size()));
  std::vector<double> tmp(size_t(ncomp)*num_nodes);
  for (size_t i = 0; i < names.size(); i++) {
    int nc = 0;
    b200io_vtk_array_info(f.h, B200IO_POINT_DATA, names[i].first.c_str(), &nc, nullptr, nullptr);
    if (nc != ncomp)
      throw std::runtime_error("The number of components in the field '" + names[i].first + "' is not equal to the number of components in the first field.");
    b200io_vtk_array_f64(f.h, B200IO_POINT_DATA, names[i].first.c_str(), tmp.data());
    for (int j = 0; j < num_nodes; j++) for (int k = 0; k < ncomp; k++) mesh.Ys(k, j, int(i)) = tmp[size_t(j)*ncomp + k];
  }
}

} // namespace vtk_xml_parser
