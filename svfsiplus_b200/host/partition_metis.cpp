// partition_metis.cpp — element partition of a mesh by METIS' k-way partition of the dual graph: what the reference obtains from
// ParMETIS_V3_PartMeshKway through split_ (Code/Source/solver/distribute.cpp:1683-1706, Code/Source/solver/SPLIT.c:
// ncommonnodes = eNoNb, the node count of a boundary element, so two volume elements are neighbours when they share a face).
// Set-up side only (runs once per mesh, on the host), not on the hot path.  The library is the serial METIS 5 that ships with the
// CUDA toolkit (libmetis_static.a, the one cuSOLVER's sparse orderings use; idx_t is 64-bit there, probed at build time by
// tests/test_partition.py).  The parallel ParMETIS result depends on the rank count of the run that made it; METIS_PartMeshDual is
// the same objective (edge cut of the dual graph, balanced element counts) computed serially.
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef SVB200_WITH_METIS
extern "C" {
typedef int64_t metis_idx_t;
int METIS_SetDefaultOptions(metis_idx_t* options);
int METIS_PartMeshDual(metis_idx_t* ne, metis_idx_t* nn, metis_idx_t* eptr, metis_idx_t* eind, metis_idx_t* vwgt, metis_idx_t* vsize,
                       metis_idx_t* ncommon, metis_idx_t* nparts, void* tpwgts, metis_idx_t* options, metis_idx_t* objval,
                       metis_idx_t* epart, metis_idx_t* npart);
}
#endif

// part[e] in [0, nparts); returns the edge cut of the dual graph (number of element faces between parts)
extern "C" long long svb200_partition_metis_impl(int nEl, int eNoN, int nNo, const int* IEN, int ncommon, int nparts, int* part)
{
#ifdef SVB200_WITH_METIS
  if (nEl <= 0 || eNoN <= 0 || nNo <= 0 || nparts <= 0) throw std::runtime_error("partition_metis: empty mesh");
  if (nparts == 1) { for (int e = 0; e < nEl; e++) part[e] = 0; return 0; }
  std::vector<metis_idx_t> eptr(size_t(nEl) + 1), eind(size_t(nEl)*eNoN), epart(nEl), npart(nNo);
  for (int e = 0; e <= nEl; e++) eptr[e] = metis_idx_t(e)*eNoN;
  for (size_t k = 0; k < eind.size(); k++) {
    if (IEN[k] < 0 || IEN[k] >= nNo) throw std::runtime_error("partition_metis: node id out of range");
    eind[k] = IEN[k];
  }
  metis_idx_t ne = nEl, nn = nNo, nc = ncommon, np = nparts, objval = 0;
  metis_idx_t options[40];
  METIS_SetDefaultOptions(options);             // numbering defaults to 0-based
  const int rc = METIS_PartMeshDual(&ne, &nn, eptr.data(), eind.data(), nullptr, nullptr, &nc, &np, nullptr, options, &objval,
                                    epart.data(), npart.data());
  if (rc != 1) throw std::runtime_error("partition_metis: METIS_PartMeshDual failed with code " + std::to_string(rc));
  for (int e = 0; e < nEl; e++) part[e] = int(epart[e]);
  return (long long)objval;
#else
  (void)nEl; (void)eNoN; (void)nNo; (void)IEN; (void)ncommon; (void)nparts; (void)part;
  throw std::runtime_error("partition_metis: this build has no METIS (libmetis_static.a of the CUDA toolkit was not found at build time)");
#endif
}
