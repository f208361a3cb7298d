/* svb200_io.h — VTK-free mesh / result / restart I/O for the B200 backend.
 *
 * The reference reads its meshes (.vtu volume meshes, .vtp boundary faces) and writes its results through the VTK library
 * (Code/Source/solver/VtkData.cpp: vtkXMLUnstructuredGridReader / vtkXMLPolyDataReader :79-95, :285-301;
 * vtkXMLUnstructuredGridWriter / vtkXMLPolyDataWriter :228-235, :503-509; call sites Code/Source/solver/vtk_xml.cpp:
 * read_vtu :568, read_vtp :438, read_vtu_pdata :667, read_vtus :716, write_vtus :913) and its restart / history files with
 * plain C++ streams (Code/Source/solver/output.cpp: output_result :46-180, read_restart_header :182-198, write_restart
 * :202-345, write_restart_header :347-362; Code/Source/solver/initialize.cpp:360-520 for the record length).
 *
 * This library restates the three formats with no dependency but zlib:
 *   - VTK XML UnstructuredGrid / PolyData: reader for ascii, inline base64 ("binary") and appended (raw or base64) data
 *     arrays, with or without the vtkZLibDataCompressor, UInt32 / UInt64 headers, every numeric VTK type; writer for the
 *     same four encodings.  A file written here opens in ParaView / VTK; a file written by VTK's XML writers (any of their
 *     data modes) is read here.
 *   - restart records ("stFiles", <stem>_NNN.bin / <stem>_last.bin): byte-identical to output::write_restart.
 *   - history lines ("histor.dat"): character-identical to output::output_result.
 *
 * Plain C ABI: opaque handle, pointers and sizes, int status (0 = ok) + b200io_last_error().  Host only: nothing here
 * touches the GPU, and nothing on the hot path (include/svb200.h) depends on it.
 */
#ifndef SVB200_IO_H
#define SVB200_IO_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200io_vtk b200io_vtk;

const char* b200io_last_error(void);

/* where an array lives */
#define B200IO_POINT_DATA 0
#define B200IO_CELL_DATA  1

/* data modes of the writer (vtkXMLWriter::SetDataMode + SetEncodeAppendedData) */
#define B200IO_ASCII            0
#define B200IO_BINARY           1   /* inline base64 */
#define B200IO_APPENDED_RAW     2
#define B200IO_APPENDED_BASE64  3   /* what vtkXMLWriter does by default */

/* VTK cell types the solver's elements map to (VtkData.cpp:97-181, :303-389) */
#define B200IO_VTK_LINE 3
#define B200IO_VTK_TRIANGLE 5
#define B200IO_VTK_QUAD 9
#define B200IO_VTK_TETRA 10
#define B200IO_VTK_HEXAHEDRON 12
#define B200IO_VTK_WEDGE 13
#define B200IO_VTK_QUADRATIC_EDGE 21
#define B200IO_VTK_QUADRATIC_TRIANGLE 22
#define B200IO_VTK_QUADRATIC_QUAD 23
#define B200IO_VTK_QUADRATIC_TETRA 24
#define B200IO_VTK_QUADRATIC_HEXAHEDRON 25
#define B200IO_VTK_BIQUADRATIC_QUAD 28
#define B200IO_VTK_TRIQUADRATIC_HEXAHEDRON 29

/* ---- reading (replaces VtkVtuData / VtkVtpData read_file + copy_* accessors, VtkData.cpp:560-1074) ---- */
int b200io_vtk_read(const char* path, b200io_vtk** out);
/* 0 = UnstructuredGrid (.vtu), 1 = PolyData (.vtp) */
int b200io_vtk_is_polydata(const b200io_vtk* h);
int b200io_vtk_num_points(const b200io_vtk* h);
int b200io_vtk_num_cells(const b200io_vtk* h);
/* nodes per cell when every cell has the same count (what the reference requires, VtkData.cpp np_elem), else -1 */
int b200io_vtk_nodes_per_cell(const b200io_vtk* h);
int b200io_vtk_points(const b200io_vtk* h, double* x /* 3 x nNo, node-major */);
int b200io_vtk_connectivity(const b200io_vtk* h, int* ien /* eNoN x nEl, cell-major, 0-based */);
int b200io_vtk_cell_types(const b200io_vtk* h, unsigned char* types /* nEl; PolyData: triangle / quad / polygon by count */);
int b200io_vtk_num_arrays(const b200io_vtk* h, int where);
const char* b200io_vtk_array_name(const b200io_vtk* h, int where, int i);
/* returns 0 and fills ncomp / ntuples / is_integer when the array exists, 1 (no error text) when it does not */
int b200io_vtk_array_info(const b200io_vtk* h, int where, const char* name, int* ncomp, int* ntuples, int* is_integer);
int b200io_vtk_array_f64(const b200io_vtk* h, int where, const char* name, double* out /* ncomp x ntuples */);
int b200io_vtk_array_i32(const b200io_vtk* h, int where, const char* name, int* out);
void b200io_vtk_free(b200io_vtk* h);

/* ---- writing (replaces set_points / set_connectivity / set_point_data / set_element_data / write) ---- */
b200io_vtk* b200io_vtk_new(int is_polydata);
int b200io_vtk_set_points(b200io_vtk* h, int nNo, const double* x /* 3 x nNo */);
/* all cells of one type (the reference writes one element type per mesh) */
int b200io_vtk_set_cells(b200io_vtk* h, int nEl, int eNoN, const int* ien /* eNoN x nEl, 0-based */, int vtk_type);
int b200io_vtk_add_array_f64(b200io_vtk* h, int where, const char* name, int ncomp, int ntuples, const double* data);
int b200io_vtk_add_array_i32(b200io_vtk* h, int where, const char* name, int ncomp, int ntuples, const int* data);
/* compress: 0 none, 1 vtkZLibDataCompressor (ignored for ascii); header64: 0 UInt32, 1 UInt64 block headers */
int b200io_vtk_write(const b200io_vtk* h, const char* path, int mode, int compress, int header64);

/* ---- restart records (output::write_restart, output.cpp:202-345; read side initialize.cpp:81-230) ----
 * One record per rank at byte offset rank * recLn:  stamp[7] (int), cTS (int), time, cpu_time (double), iNorm[nEq],
 * xn[nXn] (cplBC.xn), Yn(tDof,tnNo), An(tDof,tnNo), then by flags:  Dn  [dFlag]  ( + Ad(nsd,tnNo) [sstEq] ), or pS0(nsymd,tnNo)
 * [pstEq].  The arrays are the column-major (node-contiguous) buffers of the solver as they are.
 * With dFlag && !sstEq && !pstEq the reference writes Dn a SECOND time after the record (output.cpp:306-320); its reader
 * (initialize.cpp:118-160) and its record length (initialize.cpp:505-518) do not know about that copy, and in a multi-rank
 * file it lands on the next rank's record.  trailing_Dn = 1 reproduces it (a one-rank file is then byte-identical to the
 * reference's), trailing_Dn = 0 ends the record where recLn says it ends. */
typedef struct {
  int stamp[7];
  int cTS;
  double time, cpu_time;
  int nEq;            const double* iNorm;
  int nXn;            const double* xn;
  int tDof, tnNo;     const double* Yn; const double* An;
  int dFlag;          const double* Dn;
  int trailing_Dn;
  int sstEq, nsd;     const double* Ad;     /* nsd x tnNo */
  int pstEq, nsymd;   const double* pS0;    /* nsymd x tnNo */
} b200io_restart;

/* bytes of one record with these sizes / flags (initialize.cpp:360-518 computes recLn the same way, then takes the max
 * over ranks) */
long long b200io_restart_record_bytes(const b200io_restart* r);
/* create = 1 truncates / creates the file first (what the master rank does), then every rank writes its record */
int b200io_restart_write(const char* path, int rank, long long recLn, const b200io_restart* r, int create);
/* reads into the caller's buffers (the pointers of r must be writable and sized by the fields of r) */
int b200io_restart_read(const char* path, int rank, long long recLn, b200io_restart* r);
/* "<stem>_%03d.bin" (or "%d" from 1000 on), the reference's naming (output.cpp:265-273); returns the length */
int b200io_restart_name(const char* stem, int cTS, char* out, int cap);

/* ---- history lines (output::output_result, output.cpp:46-180) ----
 * header = 1: writes the separator / column header block (co == 1; nEq decides the second separator) into out.
 * Otherwise one line for equation `sym` at time step cTS, Newton iteration itr; saved = 1 marks a step whose results were
 * written (co == 3, "s").  elapsed = seconds since the start of the run, since_last = seconds since the previous line
 * (both measured by the caller — the reference takes them from the CPU clock).  Returns the number of characters. */
typedef struct {
  const char* sym; int cTS, itr, saved;
  double elapsed, since_last;
  double eq_iNorm, eq_pNorm;               /* eq.iNorm, eq.pNorm */
  double ri_iNorm, ri_fNorm, ri_dB, ri_callD; int ri_itr, ri_suc;   /* eq.FSILS.RI */
} b200io_history;
int b200io_history_header(int nEq, char* out, int cap);
int b200io_history_line(const b200io_history* hst, char* out, int cap);

#ifdef __cplusplus
}
#endif
#endif
