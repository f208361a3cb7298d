// VtkDataB200.cpp — the reference's VtkData / VtkVtuData / VtkVtpData classes (Code/Source/solver/VtkData.h:38-160) implemented on
// the VTK-free I/O library (include/svb200_io.h) instead of the VTK library.  It REPLACES Code/Source/solver/VtkData.cpp in a
// build without VTK: same class names, same methods, same container conventions, compiled against the unmodified VtkData.h.
//
// Conventions kept from the reference implementation (VtkData.cpp, file:line):
//   * get_connectivity() -> Array<int>(np_elem, num_elems), 0-based ids (:560-580, :834-850); get_points() / copy_points ->
//     Array<double>(3, num_points) (:657-672, :935-950, :1000-1010)
//   * copy_point_data(name, Array<double>&) fills mesh_data(comp, point); get_point_data(name) returns Array(num_points, num_comp)
//     (:592-613, :869-891, :965-985); a missing array leaves the argument untouched / returns an empty Array
//   * set_connectivity(nsd, conn): the VTK cell type follows from nsd and conn.nrows() (:97-181, :303-389); node ids outside the
//     points throw "[VtkVtuData.set_connectivity] Element <e> has the non-valid node ID <id>."
//   * set_point_data / set_element_data(name, Array(comp, n)) (:183-226, :391-485); VtkVtuData::set_point_data(Vector<int>) throws as
//     in the reference (:480-485), VtkVtpData's stores one component
//   * create_reader / create_writer choose the class from the file extension (:520-545); write() writes file_name
// Differences, on purpose: a per-element Vector<int> handed to set_point_data (write_vtp's GlobalElementID) is stored as cell data;
// arrays of any numeric type are accepted where the reference only down-casts vtkDoubleArray / vtkIntArray
// (a Float32 field or an Int64 GlobalNodeID reads instead of being silently skipped); files are written appended-raw + zlib.
#include "VtkData.h"

#include "svb200_io.h"

#include <stdexcept>
#include <string>
#include <vector>

namespace {

void ck(int rc, const char* what)
{
  if (rc != 0) throw std::runtime_error(std::string("[VtkData/b200io] ") + what + ": " + b200io_last_error());
}

// what both implementation classes hold: one library handle, in reader or writer mode
struct Impl {
  b200io_vtk* h = nullptr;
  bool polydata = false;
  ~Impl() { if (h) b200io_vtk_free(h); }

  void reset_writer() { if (!h) h = b200io_vtk_new(polydata ? 1 : 0); }

  void read(const std::string& file_name)
  {
    if (h) { b200io_vtk_free(h); h = nullptr; }
    ck(b200io_vtk_read(file_name.c_str(), &h), "read_file");
  }
  int num_elems() const { return h ? b200io_vtk_num_cells(h) : 0; }
  int num_points() const { return h ? b200io_vtk_num_points(h) : 0; }
  int np_elem() const { const int n = h ? b200io_vtk_nodes_per_cell(h) : 0; return n < 0 ? 0 : n; }

  Array<int> connectivity() const
  {
    const int ne = num_elems(), np = np_elem();
    if (ne > 0 && np == 0) throw std::runtime_error("[VtkData/b200io] the file mixes cell types; one element type per mesh is expected");
    Array<int> conn(np, ne);
    if (ne > 0) ck(b200io_vtk_connectivity(h, conn.data()), "get_connectivity");
    return conn;
  }
  Array<double> points() const
  {
    Array<double> x(3, num_points());
    if (num_points() > 0) ck(b200io_vtk_points(h, x.data()), "get_points");
    return x;
  }
  bool has(int where, const std::string& name, int* ncomp = nullptr, int* ntup = nullptr) const
  {
    return h && b200io_vtk_array_info(h, where, name.c_str(), ncomp, ntup, nullptr) == 0;
  }
  std::vector<std::string> names(int where) const
  {
    std::vector<std::string> out;
    const int n = h ? b200io_vtk_num_arrays(h, where) : 0;
    for (int i = 0; i < n; i++) out.push_back(b200io_vtk_array_name(h, where, i));
    return out;
  }
  // (comp, point) layout
  void copy(const std::string& name, Array<double>& mesh_data) const
  {
    int nc = 0, nt = 0;
    if (!has(B200IO_POINT_DATA, name, &nc, &nt) || nt == 0) return;
    std::vector<double> tmp(size_t(nc)*nt);
    ck(b200io_vtk_array_f64(h, B200IO_POINT_DATA, name.c_str(), tmp.data()), "copy_point_data");
    for (int i = 0; i < nt; i++) for (int j = 0; j < nc; j++) mesh_data(j, i) = tmp[size_t(i)*nc + j];
  }
  void copy(const std::string& name, Vector<double>& mesh_data) const
  {
    int nc = 0, nt = 0;
    if (!has(B200IO_POINT_DATA, name, &nc, &nt) || nt == 0) return;
    std::vector<double> tmp(size_t(nc)*nt);
    ck(b200io_vtk_array_f64(h, B200IO_POINT_DATA, name.c_str(), tmp.data()), "copy_point_data");
    for (int i = 0; i < nt; i++) mesh_data[i] = tmp[size_t(i)*nc];
  }
  void copy(const std::string& name, Vector<int>& mesh_data) const
  {
    int nc = 0, nt = 0;
    if (!has(B200IO_POINT_DATA, name, &nc, &nt) || nt == 0) return;
    std::vector<int> tmp(size_t(nc)*nt);
    ck(b200io_vtk_array_i32(h, B200IO_POINT_DATA, name.c_str(), tmp.data()), "copy_point_data");
    for (int i = 0; i < nt; i++) mesh_data[i] = tmp[size_t(i)*nc];
  }
  Array<double> get(const std::string& name) const
  {
    int nc = 0, nt = 0;
    if (!has(B200IO_POINT_DATA, name, &nc, &nt) || nt == 0) return Array<double>();
    std::vector<double> tmp(size_t(nc)*nt);
    ck(b200io_vtk_array_f64(h, B200IO_POINT_DATA, name.c_str(), tmp.data()), "get_point_data");
    Array<double> data(nt, nc);
    for (int i = 0; i < nt; i++) for (int j = 0; j < nc; j++) data(i, j) = tmp[size_t(i)*nc + j];
    return data;
  }

  void set_points(const Array<double>& points)
  {
    reset_writer();
    const int n = points.ncols(), nr = points.nrows();
    std::vector<double> x(size_t(3)*n, 0.0);
    for (int i = 0; i < n; i++) for (int j = 0; j < nr && j < 3; j++) x[size_t(i)*3 + j] = points(j, i);
    ck(b200io_vtk_set_points(h, n, x.data()), "set_points");
  }
  void set_conn(const char* cls, const Array<int>& conn, int vtk_type)
  {
    reset_writer();
    const int ne = conn.ncols(), np = conn.nrows(), nNo = b200io_vtk_num_points(h);
    for (int i = 0; i < ne; i++)
      for (int j = 0; j < np; j++)
        if (conn(j, i) < 0 || conn(j, i) >= nNo)
          throw std::runtime_error(std::string("[") + cls + ".set_connectivity] Element " + std::to_string(i + 1) +
                                   " has the non-valid node ID " + std::to_string(conn(j, i)) + ".");
    ck(b200io_vtk_set_cells(h, ne, np, conn.data(), vtk_type), "set_connectivity");
  }
  void add(int where, const std::string& name, const Array<double>& data)
  {
    reset_writer();
    ck(b200io_vtk_add_array_f64(h, where, name.c_str(), data.nrows(), data.ncols(), data.data()), "set data");   // (comp, n) column-major = tuple-major
  }
  void add(int where, const std::string& name, const Array<int>& data)
  {
    reset_writer();
    ck(b200io_vtk_add_array_i32(h, where, name.c_str(), data.nrows(), data.ncols(), data.data()), "set data");
  }
  void add(int where, const std::string& name, const Vector<int>& data)
  {
    reset_writer();
    // write_vtp hands the face's GlobalElementID (one value per ELEMENT) to set_point_data (vtk_xml.cpp:846-848); VTK stores such an
    // array as it is and the reference's own read_vtp then finds no element ids in the file.  An array that has the length of the
    // cells and not of the points is stored where it belongs, as cell data - the file then reads back through read_vtp.
    if (where == B200IO_POINT_DATA && data.size() != b200io_vtk_num_points(h) && data.size() == b200io_vtk_num_cells(h)) where = B200IO_CELL_DATA;
    ck(b200io_vtk_add_array_i32(h, where, name.c_str(), 1, data.size(), data.data()), "set data");
  }
  void write(const std::string& file_name) const
  {
    if (!h) throw std::runtime_error("[VtkData/b200io] write: nothing to write");
    ck(b200io_vtk_write(h, file_name.c_str(), B200IO_APPENDED_RAW, 1, 1), "write");
  }
};

int surface_cell_type(int np_elem)      // VtkVtpDataImpl::set_connectivity (VtkData.cpp:97-181): the same table for nsd 2 and 3
{
  switch (np_elem) {
    case 2: return B200IO_VTK_LINE;
    case 3: return B200IO_VTK_TRIANGLE;
    case 4: return B200IO_VTK_QUAD;
    case 6: return B200IO_VTK_QUADRATIC_TRIANGLE;
    case 8: return B200IO_VTK_QUADRATIC_QUAD;
    case 9: return B200IO_VTK_BIQUADRATIC_QUAD;
    default: throw std::runtime_error("[VtkVtpData.set_connectivity] no VTK cell type for " + std::to_string(np_elem) + " nodes per face element");
  }
}

int volume_cell_type(int nsd, int np_elem)   // VtkVtuDataImpl::set_connectivity (VtkData.cpp:303-365)
{
  if (np_elem == 2) return B200IO_VTK_LINE;
  if (nsd == 2) return surface_cell_type(np_elem);
  switch (np_elem) {
    case 3: return B200IO_VTK_TRIANGLE;
    case 4: return B200IO_VTK_TETRA;
    case 6: return B200IO_VTK_WEDGE;
    case 8: return B200IO_VTK_HEXAHEDRON;
    case 10: return B200IO_VTK_QUADRATIC_TETRA;
    case 20: return B200IO_VTK_QUADRATIC_HEXAHEDRON;
    case 27: return B200IO_VTK_TRIQUADRATIC_HEXAHEDRON;
    default: throw std::runtime_error("[VtkVtuData.set_connectivity] no VTK cell type for " + std::to_string(np_elem) + " nodes per element");
  }
}

} // namespace

class VtkVtpData::VtkVtpDataImpl : public Impl { public: VtkVtpDataImpl() { polydata = true; } };
class VtkVtuData::VtkVtuDataImpl : public Impl { public: VtkVtuDataImpl() { polydata = false; } };

// ---- VtkData ---------------------------------------------------------------------------------------------------------------
VtkData::VtkData() {}
VtkData::~VtkData() {}

VtkData* VtkData::create_reader(const std::string& file_name)
{
  const auto ext = file_name.substr(file_name.find_last_of(".") + 1);
  if (ext == "vtp") return new VtkVtpData(file_name);
  if (ext == "vtu") return new VtkVtuData(file_name);
  throw std::runtime_error("[VtkData.create_reader] '" + file_name + "' is neither a .vtp nor a .vtu file");
}

VtkData* VtkData::create_writer(const std::string& file_name)
{
  const auto ext = file_name.substr(file_name.find_last_of(".") + 1);
  if (ext == "vtp") return new VtkVtpData(file_name, false);
  if (ext == "vtu") return new VtkVtuData(file_name, false);
  throw std::runtime_error("[VtkData.create_writer] '" + file_name + "' is neither a .vtp nor a .vtu file");
}

// ---- VtkVtpData ------------------------------------------------------------------------------------------------------------
VtkVtpData::VtkVtpData() { impl = new VtkVtpDataImpl; }
VtkVtpData::VtkVtpData(const std::string& file_name, bool reader)
{
  this->file_name = file_name;
  impl = new VtkVtpDataImpl;
  if (reader) read_file(file_name);
}
VtkVtpData::~VtkVtpData() { delete impl; }

Array<int> VtkVtpData::get_connectivity() { return impl->connectivity(); }
Array<double> VtkVtpData::get_points() { return impl->points(); }
int VtkVtpData::num_elems() { return impl->num_elems(); }
int VtkVtpData::np_elem() { return impl->np_elem(); }
int VtkVtpData::num_points() { return impl->num_points(); }
void VtkVtpData::read_file(const std::string& file_name) { impl->read(file_name); }
void VtkVtpData::copy_points(Array<double>& points)
{
  const Array<double> x = impl->points();
  for (int i = 0; i < x.ncols(); i++) for (int j = 0; j < 3; j++) points(j, i) = x(j, i);
}
void VtkVtpData::copy_point_data(const std::string& data_name, Array<double>& mesh_data) { impl->copy(data_name, mesh_data); }
void VtkVtpData::copy_point_data(const std::string& data_name, Vector<double>& mesh_data) { impl->copy(data_name, mesh_data); }
void VtkVtpData::copy_point_data(const std::string& data_name, Vector<int>& mesh_data) { impl->copy(data_name, mesh_data); }
Array<double> VtkVtpData::get_point_data(const std::string& data_name) { return impl->get(data_name); }
std::vector<std::string> VtkVtpData::get_point_data_names() { return impl->names(B200IO_POINT_DATA); }
bool VtkVtpData::has_point_data(const std::string& data_name) { return impl->has(B200IO_POINT_DATA, data_name); }
void VtkVtpData::set_connectivity(const int nsd, const Array<int>& conn, const int pid)
{
  (void)nsd; (void)pid;
  impl->set_conn("VtkVtpData", conn, surface_cell_type(conn.nrows()));
}
void VtkVtpData::set_element_data(const std::string& data_name, const Array<double>& data) { impl->add(B200IO_CELL_DATA, data_name, data); }
void VtkVtpData::set_element_data(const std::string& data_name, const Array<int>& data) { impl->add(B200IO_CELL_DATA, data_name, data); }
void VtkVtpData::set_point_data(const std::string& data_name, const Array<double>& data) { impl->add(B200IO_POINT_DATA, data_name, data); }
void VtkVtpData::set_point_data(const std::string& data_name, const Array<int>& data) { impl->add(B200IO_POINT_DATA, data_name, data); }
void VtkVtpData::set_point_data(const std::string& data_name, const Vector<int>& data) { impl->add(B200IO_POINT_DATA, data_name, data); }
void VtkVtpData::set_points(const Array<double>& points) { impl->set_points(points); }
void VtkVtpData::write() { impl->write(file_name); }

// ---- VtkVtuData ------------------------------------------------------------------------------------------------------------
VtkVtuData::VtkVtuData() { impl = new VtkVtuDataImpl; }
VtkVtuData::VtkVtuData(const std::string& file_name, bool reader)
{
  this->file_name = file_name;
  impl = new VtkVtuDataImpl;
  if (reader) read_file(file_name);
}
VtkVtuData::~VtkVtuData() { delete impl; }

Array<int> VtkVtuData::get_connectivity() { return impl->connectivity(); }
Array<double> VtkVtuData::get_points() { return impl->points(); }
int VtkVtuData::num_elems() { return impl->num_elems(); }
int VtkVtuData::np_elem() { return impl->np_elem(); }
int VtkVtuData::num_points() { return impl->num_points(); }
void VtkVtuData::read_file(const std::string& file_name) { impl->read(file_name); }
void VtkVtuData::copy_points(Array<double>& points)
{
  const Array<double> x = impl->points();
  for (int i = 0; i < x.ncols(); i++) for (int j = 0; j < 3; j++) points(j, i) = x(j, i);
}
void VtkVtuData::copy_point_data(const std::string& data_name, Array<double>& mesh_data) { impl->copy(data_name, mesh_data); }
void VtkVtuData::copy_point_data(const std::string& data_name, Vector<double>& mesh_data) { impl->copy(data_name, mesh_data); }
void VtkVtuData::copy_point_data(const std::string& data_name, Vector<int>& mesh_data) { impl->copy(data_name, mesh_data); }
Array<double> VtkVtuData::get_point_data(const std::string& data_name) { return impl->get(data_name); }
std::vector<std::string> VtkVtuData::get_point_data_names() { return impl->names(B200IO_POINT_DATA); }
bool VtkVtuData::has_point_data(const std::string& data_name) { return impl->has(B200IO_POINT_DATA, data_name); }
void VtkVtuData::set_connectivity(const int nsd, const Array<int>& conn, const int pid)
{
  (void)pid;
  impl->set_conn("VtkVtuData", conn, volume_cell_type(nsd, conn.nrows()));
}
void VtkVtuData::set_element_data(const std::string& data_name, const Array<double>& data) { impl->add(B200IO_CELL_DATA, data_name, data); }
void VtkVtuData::set_element_data(const std::string& data_name, const Array<int>& data) { impl->add(B200IO_CELL_DATA, data_name, data); }
void VtkVtuData::set_point_data(const std::string& data_name, const Array<double>& data) { impl->add(B200IO_POINT_DATA, data_name, data); }
void VtkVtuData::set_point_data(const std::string& data_name, const Array<int>& data) { impl->add(B200IO_POINT_DATA, data_name, data); }
void VtkVtuData::set_point_data(const std::string& data_name, const Vector<int>& data)
{
  (void)data_name; (void)data;
  throw std::runtime_error("[VtkVtuData] set_point_data for Vector<int> not implemented.");      // as the reference, VtkData.cpp:480-485
}
void VtkVtuData::set_points(const Array<double>& points) { impl->set_points(points); }
void VtkVtuData::write() { impl->write(file_name); }
