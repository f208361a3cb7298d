// TEST INFRASTRUCTURE ONLY.  Drives svfsiplus_b200/host/VtkDataB200.cpp - the product's implementation of the reference's
// VtkData / VtkVtuData / VtkVtpData classes - through the reference's OWN header (Code/Source/solver/VtkData.h) and containers
// (Array, Vector), the way vtk_xml.cpp does (read_vtu :568-596, write_vtu :855-873, write_vtp :827-853, read_vtu_pdata :667-714).
#include "VtkData.h"

#include <cstring>
#include <memory>
#include <string>

namespace { std::string g_err; }

extern "C" {

const char* vd_last_error() { return g_err.c_str(); }

// write_vtu / write_vtus style: points, connectivity, one double point field (ncomp x nNo), GlobalNodeID (Array<int>(1,nNo)),
// one int element field.  For .vtp: GlobalNodeID goes through the Vector<int> overload, as write_vtp does.
int vd_write(const char* path, int nsd, int nNo, const double* x, int eNoN, int nEl, const int* conn, const char* fname, int ncomp,
             const double* field, const int* gnid, const char* ename, const int* edata)
{
  try {
    std::unique_ptr<VtkData> w(VtkData::create_writer(path));
    Array<double> pts(nsd, nNo);
    for (int a = 0; a < nNo; a++) for (int i = 0; i < nsd; i++) pts(i, a) = x[size_t(a)*3 + i];
    Array<int> ien(eNoN, nEl);
    std::memcpy(ien.data(), conn, sizeof(int)*size_t(eNoN)*nEl);
    w->set_points(pts);
    w->set_connectivity(nsd, ien);
    if (field) {
      Array<double> f(ncomp, nNo);
      std::memcpy(f.data(), field, sizeof(double)*size_t(ncomp)*nNo);
      w->set_point_data(fname, f);
    }
    if (gnid) {
      const std::string p(path);
      if (p.substr(p.find_last_of(".") + 1) == "vtp") {
        Vector<int> g(nNo);
        std::memcpy(g.data(), gnid, sizeof(int)*size_t(nNo));
        w->set_point_data("GlobalNodeID", g);
      } else {
        Array<int> g(1, nNo);
        std::memcpy(g.data(), gnid, sizeof(int)*size_t(nNo));
        w->set_point_data("GlobalNodeID", g);
      }
    }
    if (edata) {
      Array<int> d(1, nEl);
      std::memcpy(d.data(), edata, sizeof(int)*size_t(nEl));
      w->set_element_data(ename, d);
    }
    w->write();
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// read_vtu style.  sizes[3] = {num_points, num_elems, np_elem}; pass null outputs to query the sizes only.
int vd_read(const char* path, int* sizes, double* x, int* conn, const char* fname, int ncomp, double* field_copy, double* field_get,
            int* gnid, int* has_field, int* has_missing)
{
  try {
    std::unique_ptr<VtkData> r(VtkData::create_reader(path));
    sizes[0] = r->num_points(); sizes[1] = r->num_elems(); sizes[2] = r->np_elem();
    if (!x) return 0;
    const Array<double> pts = r->get_points();
    std::memcpy(x, pts.data(), sizeof(double)*3*size_t(sizes[0]));
    const Array<int> ien = r->get_connectivity();
    std::memcpy(conn, ien.data(), sizeof(int)*size_t(sizes[1])*sizes[2]);
    *has_field = r->has_point_data(fname) ? 1 : 0;
    *has_missing = r->has_point_data("no_such_array") ? 1 : 0;
    Array<double> f(ncomp, sizes[0]);
    r->copy_point_data(fname, f);
    std::memcpy(field_copy, f.data(), sizeof(double)*size_t(ncomp)*sizes[0]);
    Array<double> untouched(ncomp, sizes[0]);
    untouched = -7.0;
    r->copy_point_data("no_such_array", untouched);          // a missing array leaves the argument as it was
    if (untouched(0, 0) != -7.0) throw std::runtime_error("copy_point_data touched the output for a missing array");
    // the typed classes also offer get_point_data (num_points x num_comp) and the Vector<int> overload
    Array<double> g;
    Vector<int> ids(sizes[0]);
    if (auto* vtu = dynamic_cast<VtkVtuData*>(r.get())) { g = vtu->get_point_data(fname); vtu->copy_point_data("GlobalNodeID", ids); }
    else if (auto* vtp = dynamic_cast<VtkVtpData*>(r.get())) { g = vtp->get_point_data(fname); vtp->copy_point_data("GlobalNodeID", ids); }
    if (g.size() != 0) std::memcpy(field_get, g.data(), sizeof(double)*size_t(ncomp)*sizes[0]);
    std::memcpy(gnid, ids.data(), sizeof(int)*size_t(sizes[0]));
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

} // extern "C"
