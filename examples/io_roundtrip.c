/* Minimal C client of the VTK-free I/O library (include/svb200_io.h): writes a one-tet mesh with a point field, reads it back,
 * writes and re-reads a restart record.  Plain C99 - what a binding from any language has to do.
 *   gcc -std=c99 -Iinclude examples/io_roundtrip.c -Lsvfsiplus_b200 -lsvb200io -Wl,-rpath,$PWD/svfsiplus_b200 -o /tmp/io_roundtrip */
#include "svb200_io.h"

#include <stdio.h>
#include <string.h>

#define CK(call) do { if ((call) != 0) { fprintf(stderr, "%s failed: %s\n", #call, b200io_last_error()); return 1; } } while (0)

int main(int argc, char** argv)
{
  const char* dir = argc > 1 ? argv[1] : "/tmp";
  char path[1024];
  const double x[12] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1};
  const int ien[4] = {0, 1, 2, 3};
  const double vel[12] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12};

  snprintf(path, sizeof(path), "%s/one_tet.vtu", dir);
  b200io_vtk* w = b200io_vtk_new(0);
  CK(b200io_vtk_set_points(w, 4, x));
  CK(b200io_vtk_set_cells(w, 1, 4, ien, B200IO_VTK_TETRA));
  CK(b200io_vtk_add_array_f64(w, B200IO_POINT_DATA, "Velocity", 3, 4, vel));
  CK(b200io_vtk_write(w, path, B200IO_APPENDED_RAW, 1, 1));
  b200io_vtk_free(w);

  b200io_vtk* r = NULL;
  CK(b200io_vtk_read(path, &r));
  double v2[12];
  int conn[4];
  CK(b200io_vtk_array_f64(r, B200IO_POINT_DATA, "Velocity", v2));
  CK(b200io_vtk_connectivity(r, conn));
  if (b200io_vtk_num_points(r) != 4 || b200io_vtk_num_cells(r) != 1 || b200io_vtk_nodes_per_cell(r) != 4 ||
      memcmp(v2, vel, sizeof(vel)) != 0 || memcmp(conn, ien, sizeof(ien)) != 0) { fprintf(stderr, "vtu round trip differs\n"); return 1; }
  b200io_vtk_free(r);

  /* restart record of a 4-node, tDof 4 fluid state */
  snprintf(path, sizeof(path), "%s/stFile_001.bin", dir);
  double Yn[16], An[16], Y2[16], A2[16], iNorm[1] = {0.5}, iN2[1];
  for (int i = 0; i < 16; i++) { Yn[i] = i; An[i] = -i; }
  b200io_restart rec;
  memset(&rec, 0, sizeof(rec));
  rec.stamp[0] = 1; rec.stamp[1] = 1; rec.stamp[2] = 1; rec.stamp[3] = 4; rec.stamp[5] = 4;
  rec.cTS = 1; rec.time = 0.005; rec.nEq = 1; rec.iNorm = iNorm; rec.tDof = 4; rec.tnNo = 4; rec.Yn = Yn; rec.An = An;
  const long long recLn = b200io_restart_record_bytes(&rec);
  CK(b200io_restart_write(path, 0, recLn, &rec, 1));
  b200io_restart in = rec;
  in.iNorm = iN2; in.Yn = Y2; in.An = A2;
  CK(b200io_restart_read(path, 0, recLn, &in));
  if (in.cTS != 1 || memcmp(Y2, Yn, sizeof(Yn)) != 0 || memcmp(A2, An, sizeof(An)) != 0 || iN2[0] != 0.5) { fprintf(stderr, "restart round trip differs\n"); return 1; }
  printf("io_roundtrip ok (record length %lld bytes)\n", recLn);
  return 0;
}
