// TEST INFRASTRUCTURE ONLY.  The reference's vtk_xml.cpp includes three VTK headers for one debugging function (do_test,
// Code/Source/solver/vtk_xml.cpp:54-160) that nothing calls.  These stand-ins declare just enough for that function to compile, so
// that the UNMODIFIED vtk_xml.cpp can be built without the VTK library and its read_vtu / read_vtp / write_vtu / write_vtp run on
// the product's VTK-free VtkData / vtk_xml_parser replacements.  Calling anything here aborts.
#pragma once
#include <cstdlib>
typedef long long vtkIdType;
template <class T> class vtkSmartPointer {
  public:
    vtkSmartPointer() : p(nullptr) {}
    vtkSmartPointer(T* q) : p(q) {}
    vtkSmartPointer& operator=(T* q) { p = q; return *this; }
    T* operator->() const { std::abort(); return p; }
  private:
    T* p;
};
