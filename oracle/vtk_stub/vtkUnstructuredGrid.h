// TEST INFRASTRUCTURE ONLY: see vtkSmartPointer.h.
#pragma once
#include "vtkSmartPointer.h"
class vtkPoints {
  public:
    void GetPoint(vtkIdType, double*) { std::abort(); }
};
class vtkUnstructuredGrid {
  public:
    vtkIdType GetNumberOfPoints() { std::abort(); return 0; }
    vtkIdType GetNumberOfCells() { std::abort(); return 0; }
    vtkPoints* GetPoints() { std::abort(); return nullptr; }
};
