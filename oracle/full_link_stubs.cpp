// TEST INFRASTRUCTURE ONLY.  Link-time stand-ins that let the COMPLETE reference solver (every file of Code/Source/solver except the
// PETSc / Trilinos shims and the two VTK-bound files the product replaces) be linked here and run on ONE rank:
//   split_            the ParMETIS wrapper (Code/ThirdParty/parmetis_internal, needs MPI): a one-process run returns from part_msh
//                     before it is reached (distribute.cpp:1476-1510)
//   remesh3d_tetgen   TetGen remeshing (remeshTet.cpp needs tetgen.h): never reached without <Remesher>
#include <array>
#include <stdexcept>

extern "C" int split_(int*, int*, int*, int*, int*, int*, float*, int*)
{
  throw std::runtime_error("[oracle] split_ (ParMETIS) is not available: the full reference runs on one rank here");
}

void remesh3d_tetgen(const int, const int, const double*, const int*, const std::array<double,3>&, int*)
{
  throw std::runtime_error("[oracle] remesh3d_tetgen (TetGen) is not available");
}
