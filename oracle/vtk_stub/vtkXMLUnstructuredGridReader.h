// TEST INFRASTRUCTURE ONLY: see vtkSmartPointer.h.
#pragma once
#include "vtkUnstructuredGrid.h"
class vtkXMLUnstructuredGridReader {
  public:
    static vtkXMLUnstructuredGridReader* New() { std::abort(); return nullptr; }
    void SetFileName(const char*) { std::abort(); }
    void Update() { std::abort(); }
    vtkUnstructuredGrid* GetOutput() { std::abort(); return nullptr; }
};
