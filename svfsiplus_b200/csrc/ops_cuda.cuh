// ops_cuda.cuh — the device policy behind krylov.hpp: owns the device CSR (solver ordering), a
// stack arena for Krylov work vectors, the deterministic reduction buffers, the halo lists and the
// NCCL communicator, and launches the kernels of kernels.cuh on ONE stream.  There is no host
// implementation of any of these operations in the product.
#pragma once

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "krylov.hpp"
#include "lhs_layout.hpp"
#include "peer_comm.cuh"
#include "spmv_tiled.cuh"
#include "fused_halo.cuh"

namespace svb200 {

#define CU_CHECK(call)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + \
                               __FILE__ + ":" + std::to_string(__LINE__));                    \
  } while (0)

// ---- NCCL through dlopen (the library torch already loaded, or the system one) -------------------
struct Nccl {
  typedef struct { char internal[128]; } UniqueId;
  typedef void* Comm;
  void* lib = nullptr;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, Comm, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  static constexpr int kFloat64 = 8, kSum = 0, kInt32 = 2, kMax = 2, kInt8 = 0, kMin = 3;

  void load()
  {
    if (lib) return;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (auto n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) throw std::runtime_error("NCCL library not found (libnccl.so.2)");
    auto sym = [&](const char* s) { void* p = dlsym(lib, s); if (!p) throw std::runtime_error(std::string("NCCL symbol missing: ") + s); return p; };
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
    AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
    Send = reinterpret_cast<decltype(Send)>(sym("ncclSend"));
    Recv = reinterpret_cast<decltype(Recv)>(sym("ncclRecv"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
  }
  void check(int r, const char* what)
  {
    if (r != 0) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + (GetErrorString ? GetErrorString(r) : "?"));
  }
};

struct DevFace {               // lhs.face[faIn] (liner_solver/fils_struct.hpp:116-161)
  int nNo = 0, dof = 0, bGrp = B200_BC_DIR;
  bool shared = false, inc = true, coupled = false, set = false;
  double res = 0.0, nS = 0.0;
  int* glob = nullptr;
  double* val = nullptr;
  double* valM = nullptr;
};

struct HaloReq { int peer = -1, n = 0; int* ptr = nullptr; double* sbuf = nullptr; double* rbuf = nullptr; };

// kernel classes for live (CUDA-event) timing inside a step: see b200_profile_read in svb200.h
enum KClass { KC_SPMV_VV4 = 0, KC_SPMV_VV3, KC_SPMV_SS, KC_SPMV_SV, KC_SPMV_VS, KC_MULTI_DOT, KC_CGS_UPDATE,
              KC_BLAS1, KC_SCALE_VAL, KC_DEPART, KC_ASSEMBLY, KC_HALO, KC_COUNT };

class CudaOps {
 public:
  cudaStream_t st = nullptr;
  long long launches = 0;
  double phase_ms[4] = {0, 0, 0, 0};

  // ---- live kernel timing: event pairs on the launch stream, resolved after the step ---------------
  bool profiling = false;
  unsigned profile_mask = 0xffffffffu;      // kernel classes that record events while profiling
  struct Span { cudaEvent_t a, b; int cls; };
  std::vector<Span> spans;
  size_t span_used = 0;
  double cls_ms[KC_COUNT] = {0};
  double cls_bytes[KC_COUNT] = {0};
  long long cls_launches[KC_COUNT] = {0};

  struct Scope {
    CudaOps& o; int idx;
    Scope(CudaOps& o_, int cls, double bytes, int nl = 1) : o(o_), idx(-1)
    {
      if (!o.profiling || !((o.profile_mask >> cls) & 1u)) return;
      if (o.span_used == o.spans.size()) {
        Span s; cudaEventCreate(&s.a); cudaEventCreate(&s.b); s.cls = cls; o.spans.push_back(s);
      }
      idx = int(o.span_used++);
      o.spans[idx].cls = cls;
      o.cls_bytes[cls] += bytes;
      o.cls_launches[cls] += nl;
      cudaEventRecord(o.spans[idx].a, o.st);
    }
    ~Scope() { if (idx >= 0) cudaEventRecord(o.spans[idx].b, o.st); }
  };
  // fold the recorded spans into cls_ms (call after a stream synchronize)
  void profile_resolve()
  {
    for (size_t i = 0; i < span_used; i++) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, spans[i].a, spans[i].b) == cudaSuccess) cls_ms[spans[i].cls] += ms;
    }
    span_used = 0;
  }
  void profile_reset()
  {
    span_used = 0;
    for (int c = 0; c < KC_COUNT; c++) { cls_ms[c] = 0; cls_bytes[c] = 0; cls_launches[c] = 0; }
  }
  // Device-resident loops enqueue iterations past convergence that return at once (skip flag): their algorithmic bytes and launch
  // counts must not be credited to the kernel classes.  The loops snapshot the counters around every enqueued iteration and take the
  // skipped iterations' share back once the device has said where it stopped.
  struct ProfCount { double bytes[KC_COUNT]; long long launches[KC_COUNT]; };
  ProfCount prof_snapshot() const
  {
    ProfCount p;
    for (int c = 0; c < KC_COUNT; c++) { p.bytes[c] = cls_bytes[c]; p.launches[c] = cls_launches[c]; }
    return p;
  }
  void prof_uncredit(const ProfCount& before, const ProfCount& after)
  {
    for (int c = 0; c < KC_COUNT; c++) { cls_bytes[c] -= after.bytes[c] - before.bytes[c]; cls_launches[c] -= after.launches[c] - before.launches[c]; }
  }
  // algorithmic bytes (SURVEY.md par. 8d)
  double bytes_vv(int d) const { return double(nnz_)*(8.0*d*d + 4.0) + double(nNo_)*(16.0*d + 8.0); }
  double bytes_ss() const { return double(nnz_)*12.0 + double(nNo_)*24.0; }
  double bytes_svs(int d) const { return double(nnz_)*(8.0*d + 4.0) + double(nNo_)*(8.0*d + 16.0); }

  // structure (solver ordering)
  int gnNo_ = 0, nNo_ = 0, mynNo_ = 0, nnz_ = 0;
  int* rowPtr = nullptr;     // nNo+1
  int* col = nullptr;        // nnz (solver ids, row entries keep the assembly order)
  int* diag = nullptr;       // nNo
  int* tpos = nullptr;       // nnz transpose positions
  std::vector<DevFace> faces;
  std::vector<HaloReq> reqs;
  int halo_dof_cap = 0;

  // row tiles of the TMA-staged SpMV kernels (spmv_tiled.cuh): tile t = rows [tile_row[t], tile_row[t+1]); tiles never straddle
  // ovA / ovB, tile_at[] = first tile of the row ranges [0, ovA), [ovA, ovB), [ovB, nNo) and the end
  int* tile_row = nullptr;
  int tile_at[4] = {0, 0, 0, 0};
  bool tiles_ok = false;
  void build_tiles(const std::vector<int>& rp)          // rp: host copy of rowPtr (solver ordering)
  {
    cudaFree(tile_row); tile_row = nullptr; tiles_ok = false;
    const int cuts[4] = {0, overlap_ok ? ovA : 0, overlap_ok ? ovB : nNo_, nNo_};
    const std::vector<int> tr = row_tiles(rp, cuts, kTileRows, kTileCap, tile_at);     // lhs_layout.hpp (pure host, CPU-tested)
    if (tr.empty()) return;                              // a single row longer than a tile: keep the per-lane kernels
    CU_CHECK(cudaMalloc(&tile_row, sizeof(int)*tr.size()));
    CU_CHECK(cudaMemcpyAsync(tile_row, tr.data(), sizeof(int)*tr.size(), cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaStreamSynchronize(st));
    tiles_ok = true;
  }
  // tile range of a row range handed out by rows_then_halo (always one of the three segments or the whole matrix)
  bool tiles_for(int r0, int r1, int& t0, int& t1) const
  {
    if (!tiles_ok) return false;
    if (r0 == 0 && r1 == nNo_) { t0 = tile_at[0]; t1 = tile_at[3]; return true; }
    if (!overlap_ok) return false;
    if (r0 == 0 && r1 == ovA) { t0 = tile_at[0]; t1 = tile_at[1]; return true; }
    if (r0 == ovA && r1 == ovB) { t0 = tile_at[1]; t1 = tile_at[2]; return true; }
    if (r0 == ovB && r1 == nNo_) { t0 = tile_at[2]; t1 = tile_at[3]; return true; }
    return false;
  }
  unsigned tiled_attr_mask = 0;        // shapes whose kernel already carries the dynamic shared memory attribute on this device
  template <class S>
  void launch_tiled(int t0, int t1, const double* K, const S& shape)
  {
    if (!((tiled_attr_mask >> S::ID) & 1u)) {
      CU_CHECK(cudaFuncSetAttribute(k_spmv_tiled<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(tile_smem_bytes<S>())));
      tiled_attr_mask |= (1u << S::ID);
    }
    if (t1 <= t0) return;
    const int g = std::min(t1 - t0, kSmCount);
    k_spmv_tiled<S><<<g, kTileThreads, tile_smem_bytes<S>(), st>>>(skip_flag, t0, t1, tile_row, rowPtr, col, K, shape);
    post();
  }

  // overlap of the halo exchange with the interior rows: in the FSILS ordering the rows that appear in an
  // overlap list are [0, ovA) (shared with lower ranks) and [ovB, nNo) (shared with higher ranks)
  cudaStream_t st2 = nullptr;          // communication stream (highest priority)
  cudaEvent_t ev_b = nullptr, ev_c = nullptr;
  int ovA = 0, ovB = 0;
  bool overlap_ok = false;

  // communicator
  Nccl nccl;
  Nccl::Comm comm = nullptr;
  int rank = 0, nranks = 1;

  // peer-mapped transport (peer_comm.cuh): set up collectively by peer_setup() at the end of b200_lhs_create; when it is
  // not available (no CUDA IPC between the ranks, SVB200_P2P=0) the NCCL send/recv + all-reduce path below is used.
  bool p2p = false;
  std::string p2p_why = "single rank";
  char* win = nullptr;                       // own window
  size_t win_bytes = 0;
  std::vector<char*> peer_win;               // every rank's window as mapped in this process
  PeerRedArgs red_args{};
  PeerState* peer_state = nullptr;
  PeerHaloReq* d_peer_reqs = nullptr;
  int* d_halo_ptr_all = nullptr;             // concatenated overlap lists
  int halo_tot = 0;
  int* d_hn_node = nullptr; int* d_hn_ptr = nullptr; int2* d_hn_src = nullptr;
  int halo_nh = 0;

  // reductions
  static constexpr int kMaxSlots = 1024;
  static constexpr size_t kDotsTwoStageMaxN = size_t(1) << 21;   // ... and only for vectors up to 2 M entries per rank: on longer ones the
                                                                 // first stage hides the in-kernel tail and the extra launch costs more (P10, one GPU:
                                                                 // 0.137 vs 0.133 ms per launch; 8 GPUs: 47.9 vs 52.8 us)
  static constexpr int kDotsTwoStage = 12;   // from this many dots per launch the ordered sum of the partials runs as its own kernel (warp per dot)
  double* red_d = nullptr;
  double* red_h = nullptr;        // pinned
  double* partial_d = nullptr;
  unsigned int* counter_d = nullptr;   // [0] multi-dot, [1] face dot, [2..3] one-launch face update (arrivals, finished)
  double* face_partial_d = nullptr;
  CgState* cg_d = nullptr;             // device-resident CG scalars
  CgState* cg_h = nullptr;             // pinned: [0..1] polling slots, [2] initial / final state
  cudaEvent_t cg_ev[2] = {nullptr, nullptr};
  int cg_batch = 8;
  const int* skip_flag = nullptr;      // != null inside a device-resident loop: heavy kernels return at once when set
  // device-resident Arnoldi loop (gmres_device_cycle)
  GmresState* gm_d = nullptr;
  GmresState* gm_h = nullptr;          // pinned: [0..1] polling slots, [2] initial / final state
  cudaEvent_t gm_ev[2] = {nullptr, nullptr};
  double* gm_arr = nullptr;            // h((sD+1) x sD), c(sD), s(sD), err(sD+1)
  double* gm_host = nullptr;           // pinned copy of h and err for the back substitution
  int gm_sD = 0;
  int gm_batch = 8;
  int variant_face_fused = 2;          // 2 (default): one multi-CTA launch with a grid barrier (same sums, bit for bit, as the two
                                       // launches of 0); 1: single-CTA one-launch form, measured SLOWER at P10
                                       // (1 665 vs 1 593 ms per Newton iteration, profiles/r02_ab_device_loop.txt): one CTA walking
                                       // the 28 k face values costs ~40 us against two multi-CTA launches of ~5 us
  int variant_gmres_device = 1;        // b200_tune("gmres_device", 0): host-driven Arnoldi loop (one D2H sync per iteration)

  // arena
  struct Chunk { char* p; size_t cap, top; };
  std::vector<Chunk> chunks;
  int cur_chunk = 0;
  struct Mark { int chunk; size_t top; };

  explicit CudaOps(int device)
  {
    CU_CHECK(cudaSetDevice(device));
    CU_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    {
      int lo = 0, hi = 0;
      CU_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CU_CHECK(cudaStreamCreateWithPriority(&st2, cudaStreamNonBlocking, hi));
      CU_CHECK(cudaEventCreateWithFlags(&ev_b, cudaEventDisableTiming));
      CU_CHECK(cudaEventCreateWithFlags(&ev_c, cudaEventDisableTiming));
    }
    if (const char* e = getenv("SVB200_FUSED")) variant_fused = std::atoi(e);      // A/B of the fused product + exchange kernel
    if (const char* e = getenv("SVB200_GMRES_DEVICE")) variant_gmres_device = std::atoi(e);
    if (const char* e = getenv("SVB200_SCHUR_GP")) variant_gp = std::atoi(e);
    if (const char* e = getenv("SVB200_FACE_FUSED")) variant_face_fused = std::atoi(e);
    if (const char* e = getenv("SVB200_GM_BATCH")) gm_batch = std::max(1, std::atoi(e));
    CU_CHECK(cudaMalloc(&red_d, sizeof(double)*kMaxSlots));
    CU_CHECK(cudaMallocHost(&red_h, sizeof(double)*kMaxSlots));
    CU_CHECK(cudaMalloc(&partial_d, sizeof(double)*kRedBlocks*kMaxDots));
    CU_CHECK(cudaMalloc(&counter_d, 4*sizeof(unsigned int)));
    CU_CHECK(cudaMemset(counter_d, 0, 4*sizeof(unsigned int)));
    CU_CHECK(cudaMalloc(&face_partial_d, sizeof(double)*kFaceBlocks));
    CU_CHECK(cudaMalloc(&cg_d, sizeof(CgState)));
    CU_CHECK(cudaMallocHost(&cg_h, 3*sizeof(CgState)));
    CU_CHECK(cudaEventCreateWithFlags(&cg_ev[0], cudaEventDisableTiming));
    CU_CHECK(cudaEventCreateWithFlags(&cg_ev[1], cudaEventDisableTiming));
  }
  ~CudaOps()
  {
    peer_teardown();
    for (auto& c : chunks) cudaFree(c.p);
    for (auto& sp : spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto& f : faces) { cudaFree(f.glob); cudaFree(f.val); cudaFree(f.valM); }
    for (auto& r : reqs) { cudaFree(r.ptr); cudaFree(r.sbuf); cudaFree(r.rbuf); }
    cudaFree(rowPtr); cudaFree(col); cudaFree(diag); cudaFree(tpos); cudaFree(tile_row);
    cudaFree(cg_d); cudaFreeHost(cg_h);
    cudaFree(gm_d); cudaFreeHost(gm_h); cudaFree(gm_arr); cudaFreeHost(gm_host);
    if (gm_ev[0]) cudaEventDestroy(gm_ev[0]);
    if (gm_ev[1]) cudaEventDestroy(gm_ev[1]);
    if (cg_ev[0]) cudaEventDestroy(cg_ev[0]);
    if (cg_ev[1]) cudaEventDestroy(cg_ev[1]);
    cudaFree(red_d); cudaFreeHost(red_h); cudaFree(partial_d); cudaFree(counter_d); cudaFree(face_partial_d);
    if (comm) nccl.CommDestroy(comm);
    if (ev_b) cudaEventDestroy(ev_b);
    if (ev_c) cudaEventDestroy(ev_c);
    if (st2) cudaStreamDestroy(st2);
    if (st) cudaStreamDestroy(st);
  }

  int nNo() const { return nNo_; }
  int mynNo() const { return mynNo_; }
  size_t nnz() const { return size_t(nnz_); }
  bool is_master() const { return rank == 0; }

  // ---- arena ------------------------------------------------------------------------------------
  Mark mark() const { return Mark{cur_chunk, chunks.empty() ? 0 : chunks[cur_chunk].top}; }
  void release(Mark m)
  {
    if (chunks.empty()) return;
    for (int c = m.chunk + 1; c < int(chunks.size()); c++) chunks[c].top = 0;
    cur_chunk = m.chunk;
    chunks[cur_chunk].top = m.top;
  }
  double* vec(size_t n)
  {
    const size_t bytes = ((n*sizeof(double) + 16 + 255)/256)*256;     // +16: the tiled SpMV's bulk copies round their size up to 16 bytes
    for (int c = cur_chunk; c < int(chunks.size()); c++) {
      if (c > cur_chunk && chunks[c].top != 0) continue;
      if (chunks[c].cap - chunks[c].top >= bytes) {
        char* p = chunks[c].p + chunks[c].top;
        chunks[c].top += bytes;
        cur_chunk = c;
        return reinterpret_cast<double*>(p);
      }
    }
    Chunk nc;
    nc.cap = std::max(bytes, size_t(64) << 20);
    nc.top = bytes;
    CU_CHECK(cudaMalloc(&nc.p, nc.cap));
    chunks.push_back(nc);
    cur_chunk = int(chunks.size()) - 1;
    return reinterpret_cast<double*>(nc.p);
  }

  // ---- launch helpers -----------------------------------------------------------------------------
  static int grid_for(size_t n, int threads, int per_thread = 4)
  {
    size_t b = (n + size_t(threads)*per_thread - 1)/(size_t(threads)*per_thread);
    const size_t cap = size_t(kSmCount)*16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return int(b);
  }
  // row-per-4-lanes kernels: enough CTAs to cover all rows, capped at 16 waves of 148 CTAs
  static int grid_rows(int nNo)
  {
    size_t groups_per_block = 256/4;
    size_t b = (size_t(nNo) + groups_per_block - 1)/groups_per_block;
    const size_t cap = size_t(kSmCount)*32;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return int(b);
  }
  void post() { launches++; }

  // ---- BLAS-1 -------------------------------------------------------------------------------------
  void zero(size_t n, double* x) { CU_CHECK(cudaMemsetAsync(x, 0, n*sizeof(double), st)); }
  void copy(size_t n, const double* x, double* y) { CU_CHECK(cudaMemcpyAsync(y, x, n*sizeof(double), cudaMemcpyDeviceToDevice, st)); }
  void fill(size_t n, double a, double* x) { Scope sc(*this, KC_BLAS1, 8.0*n); k_fill<<<grid_for(n, 256), 256, 0, st>>>(n, a, x); post(); }
  void axpy(size_t n, double a, const double* x, double* y) { Scope sc(*this, KC_BLAS1, 24.0*n); k_axpy<<<grid_for(n, 256), 256, 0, st>>>(n, a, x, y); post(); }
  void scal(size_t n, double a, double* x) { Scope sc(*this, KC_BLAS1, 16.0*n); k_scal<<<grid_for(n, 256), 256, 0, st>>>(n, a, x); post(); }
  void divs(size_t n, double d, double* x) { Scope sc(*this, KC_BLAS1, 16.0*n); k_divs<<<grid_for(n, 256), 256, 0, st>>>(n, d, x); post(); }
  void sub(size_t n, const double* a, const double* b, double* out) { Scope sc(*this, KC_BLAS1, 24.0*n); k_sub<<<grid_for(n, 256), 256, 0, st>>>(n, a, b, out); post(); }
  void mul_inplace(size_t n, const double* w, double* x) { Scope sc(*this, KC_BLAS1, 24.0*n); k_mul<<<grid_for(n, 256), 256, 0, st>>>(n, w, x); post(); }
  void lin2(size_t n, double* out, double a, const double* x, double b, const double* y) { Scope sc(*this, KC_BLAS1, 24.0*n); k_lin2<<<grid_for(n, 256), 256, 0, st>>>(n, out, a, x, b, y); post(); }
  void axpy2(size_t n, double* X, double a, const double* P, double b, const double* S) { Scope sc(*this, KC_BLAS1, 32.0*n); k_axpy2<<<grid_for(n, 256), 256, 0, st>>>(n, X, a, P, b, S); post(); }
  void bicg_p_update(size_t n, double* P, const double* R, const double* V, double beta, double omega) { Scope sc(*this, KC_BLAS1, 32.0*n); k_bicg_p<<<grid_for(n, 256), 256, 0, st>>>(n, P, R, V, beta, omega); post(); }

  // out = base + sum_j coef[j] V[(j0+j)*stride], sequential in j (base may be null, out may alias base)
  void lin_comb(size_t n, double* out, const double* base, int k, const double* V, size_t stride, int j0, const double* coef)
  {
    const double* b = base;
    int done = 0;
    if (k == 0) { if (base == nullptr) zero(n, out); else if (base != out) copy(n, base, out); return; }
    while (done < k) {
      const int m = std::min(kMaxComb, k - done);
      CombArgs a;
      for (int j = 0; j < m; j++) a.coef[j] = coef[done + j];
      Scope sc(*this, KC_BLAS1, 8.0*double(n)*(m + 2));
      k_lin_comb<<<grid_for(n, 256), 256, 0, st>>>(n, out, b, m, V + size_t(j0 + done)*stride, stride, a);
      post();
      b = out;
      done += m;
    }
  }

  // ---- reductions ---------------------------------------------------------------------------------
  // red[slot0 + j] = <base + j*stride, w> over the first mynNo nodes, j < count (local part only)
  void dots_local(int dof, int count, const double* base, size_t stride, const double* w, int slot0)
  {
    if (slot0 + count > kMaxSlots) throw std::runtime_error("reduction slot overflow");
    const size_t n = size_t(dof)*mynNo_;
    int done = 0;
    while (done < count) {
      const int m = std::min(kMaxDots, count - done);
      Scope sc(*this, KC_MULTI_DOT, 8.0*double(n)*(m + 1));
      // enough CTAs to fill the machine, never more than one per 1024 entries (tiny systems)
      const int g = int(std::min<size_t>(size_t(kRedBlocks), (n + 1023)/1024 + 1));
      if (m >= kDotsTwoStage && n <= kDotsTwoStageMaxN) {
        k_multi_dot<<<g, kRedThreads, 0, st>>>(skip_flag, n, base + size_t(done)*stride, stride, w, m, partial_d, counter_d, red_d, slot0 + done,
                                               PeerRedArgs(), nullptr, 1);
        post();
        k_multi_dot_final<<<(m + kRedThreads/32 - 1)/(kRedThreads/32), kRedThreads, 0, st>>>(skip_flag, g, m, partial_d, counter_d, red_d, slot0 + done);
        post();
      } else {
        k_multi_dot<<<g, kRedThreads, 0, st>>>(skip_flag, n, base + size_t(done)*stride, stride, w, m, partial_d, counter_d, red_d, slot0 + done);
        post();
      }
      done += m;
    }
  }
  // all-reduce of n doubles at device pointer v (MPI_Allreduce of dot.cpp / norm.cpp / bcast.cpp:51-58)
  void allreduce(double* v, int n, bool is_max = false)
  {
    if (nranks == 1) return;
    if (p2p) {
      if (n > kPeerSlots) throw std::runtime_error("allreduce: more slots than the peer mailbox holds");
      if (is_max) k_peer_allreduce<1><<<1, 256, 0, st>>>(red_args, peer_state, v, n);
      else k_peer_allreduce<0><<<1, 256, 0, st>>>(red_args, peer_state, v, n);
      post();
      return;
    }
    nccl.check(nccl.AllReduce(v, v, size_t(n), Nccl::kFloat64, is_max ? Nccl::kMax : Nccl::kSum, comm, st), "AllReduce");
  }
  void reduce_begin(int nslots) { allreduce(red_d, nslots); }
  // dots_local + the all-reduce of the same slots in ONE launch: the last CTA of the reduction kernel runs the all-reduce as its
  // epilogue (peer transport, one launch of <= kMaxDots dots); otherwise the two separate steps
  bool fuse_reduce() const { return variant_fused != 0 && nranks > 1 && p2p; }
  void dots_reduce(int dof, int count, const double* base, size_t stride, const double* w, int slot0)
  {
    if (!fuse_reduce() || count > kMaxDots) { dots_local(dof, count, base, stride, w, slot0); allreduce(red_d + slot0, count); return; }
    if (slot0 + count > kMaxSlots) throw std::runtime_error("reduction slot overflow");
    const size_t n = size_t(dof)*mynNo_;
    Scope sc(*this, KC_MULTI_DOT, 8.0*double(n)*(count + 1));
    const int g = int(std::min<size_t>(size_t(kRedBlocks), (n + 1023)/1024 + 1));
    if (count >= kDotsTwoStage && n <= kDotsTwoStageMaxN) {
      k_multi_dot<<<g, kRedThreads, 0, st>>>(skip_flag, n, base, stride, w, count, partial_d, counter_d, red_d, slot0, PeerRedArgs(), nullptr, 1);
      post();
      k_multi_dot_final<<<(count + kRedThreads/32 - 1)/(kRedThreads/32), kRedThreads, 0, st>>>(skip_flag, g, count, partial_d, counter_d, red_d, slot0,
                                                                                                 red_args, peer_state);
      post();
    } else {
      k_multi_dot<<<g, kRedThreads, 0, st>>>(skip_flag, n, base, stride, w, count, partial_d, counter_d, red_d, slot0, red_args, peer_state);
      post();
    }
  }
  void reduce_fetch(int nslots, double* out)
  {
    CU_CHECK(cudaMemcpyAsync(red_h, red_d, sizeof(double)*nslots, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    std::memcpy(out, red_h, sizeof(double)*nslots);
  }
  double dot(int dof, const double* a, const double* b)
  {
    dots_local(dof, 1, a, 0, b, 0);
    reduce_begin(1);
    double r;
    reduce_fetch(1, &r);
    return r;
  }
  double norm(int dof, const double* a) { return std::sqrt(dot(dof, a, a)); }

  void cgs_update_scale(int dof, int k, const double* base, size_t stride, double* w, int slot0)
  {
    const size_t n = size_t(dof)*nNo_;
    Scope sc(*this, KC_CGS_UPDATE, 8.0*double(n)*(k + 2));
    k_cgs_update_scale<<<grid_for(n, 256), 256, sizeof(double)*(k+1), st>>>(n, k, base, stride, w, red_d, slot0, skip_flag);
    post();
  }

  // entries in flight per lane in the two Schur passes (A/B on B200, profiles/r01_tour_b.jsonl): pass 1 is
  // fastest with two (0.151 vs 0.163 ms at P10), pass 2 with four (0.193 vs 0.217 ms)
  int variant_gp = 3, variant_sp = 1;   // 2: TMA-staged row tiles (spmv_tiled.cuh); variant_gp 3 (default): component-wise copy of G
                                        // (k_schur_gp_soa: 0.843 vs 0.785 of the HBM peak stand-alone, same sums bit for bit; profiles/r02_schur_gp_soa.txt)
  int variant_vv3 = 4;       // (default 4: 0.853 vs 0.813 of the HBM peak at P10, profiles/r02_vv3_variants_d.jsonl)  0: lane = component, 1: lanes stride over the row's blocks, 2: TMA-staged row tiles, 3: as 0 with an L2 evict-last policy on the gathered vector, 4 / 5: column-owner lanes with / without that policy (A/B by op_bench)
  int variant_narrow = 0;    // spmv_ss / sv / vs: 0 per-lane loads, 2 TMA-staged row tiles

  // ---- SpMV (+ overlap-node add) --------------------------------------------------------------------
  // Every product is launched on row ranges: launch(r0, r1) computes rows [r0, r1).  With more than one rank the
  // boundary rows go first, their exchange (pack -> ncclSend/ncclRecv) runs on the communication stream while the
  // interior rows are computed, and the received contributions are added in request order afterwards.
  template <class Launch>
  void rows_then_halo(int dof, double* out, int ld, Launch&& launch)
  {
    if (nranks == 1 || reqs.empty()) { launch(0, nNo_); return; }
    if (!overlap_ok) { launch(0, nNo_); halo_add(dof, out, ld); return; }
    if (p2p) {
      // one stream, no library call: boundary rows -> push into the neighbours' windows -> interior rows (the NVLink
      // stores fly meanwhile) -> acquire the neighbours' flags and add
      if (ovA > 0) launch(0, ovA);
      if (ovB < nNo_) launch(ovB, nNo_);
      { Scope hs(*this, KC_HALO, 0.0, 1); halo_push(dof, out, ld ? ld : dof); }          // profiled separately: push ...
      if (ovB > ovA) launch(ovA, ovB);
      { Scope hs(*this, KC_HALO, 0.0, 1); halo_wait_add(dof, out, ld ? ld : dof); }      // ... and wait + add (incl. the wait for the neighbour)
      return;
    }
    if (ovA > 0) launch(0, ovA);
    if (ovB < nNo_) launch(ovB, nNo_);
    CU_CHECK(cudaEventRecord(ev_b, st));
    CU_CHECK(cudaStreamWaitEvent(st2, ev_b, 0));
    halo_exchange(dof, out, ld ? ld : dof, st2);
    CU_CHECK(cudaEventRecord(ev_c, st2));
    if (ovB > ovA) launch(ovA, ovB);
    CU_CHECK(cudaStreamWaitEvent(st, ev_c, 0));
    halo_accumulate(dof, out, ld ? ld : dof);
  }

  // ---- fused product + overlap exchange (fused_halo.cuh): one launch instead of four -------------------------------
  int variant_fused = 1;               // b200_tune("fused", 0) selects the unfused peer path (A/B, parity tests)
  int fused_grid[4] = {0, 0, 0, 0};    // resident CTAs per shape (occupancy x SMs), filled at first use
  bool use_fused() const { return variant_fused != 0 && nranks > 1 && p2p && overlap_ok && !reqs.empty(); }
  template <class Rows>
  void launch_fused(int shape_id, const Rows& rows, int dof, double* out, int ld)
  {
    if (dof > halo_dof_cap) throw std::runtime_error("halo buffers too small for dof");
    if (fused_grid[shape_id] == 0) {
      int per_sm = 0;
      CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rows_halo<Rows>, 256, 0));
      if (per_sm < 1) throw std::runtime_error("fused halo kernel cannot be resident");
      fused_grid[shape_id] = per_sm*kSmCount;
    }
    FusedHaloArgs f;
    f.skip = skip_flag; f.nNo = nNo_; f.ovA = ovA; f.ovB = ovB; f.dof = dof; f.ld = ld; f.dofcap = halo_dof_cap; f.out = out;
    f.nreq = int(reqs.size()); f.reqs = d_peer_reqs; f.ptr_all = d_halo_ptr_all; f.halo_tot = halo_tot;
    f.nh = halo_nh; f.hn_node = d_hn_node; f.hn_ptr = d_hn_ptr; f.hn_src = d_hn_src; f.ps = peer_state;
    k_rows_halo<Rows><<<fused_grid[shape_id], 256, 0, st>>>(rows, f);
    post();
  }

  void spmv_vv(int dof, const double* K, const double* U, double* KU)
  {
    if (dof < 1 || dof > 4) throw std::runtime_error("spmv_vv: dof > 4 is not a supported FSILS path");
    Scope sc(*this, dof == 4 ? KC_SPMV_VV4 : KC_SPMV_VV3, bytes_vv(dof));
    if (use_fused() && dof == 4) { launch_fused(1, RowsVV4{rowPtr, col, K, U, KU}, 4, KU, 4); return; }
    if (use_fused() && dof == 3 && variant_vv3 == 4) { launch_fused(0, RowsVV3{rowPtr, col, K, U, KU}, 3, KU, 3); return; }
    rows_then_halo(dof, KU, dof, [&](int r0, int r1) {
      const int n = r1 - r0, g = grid_rows(n);
      const int* rp = rowPtr + r0;
      double* out = KU + size_t(r0)*dof;
      int t0, t1;
      if (dof == 3 && variant_vv3 == 2 && tiles_for(r0, r1, t0, t1)) { launch_tiled(t0, t1, K, TileVV3{U, KU}); return; }
      switch (dof) {
        case 4: k_spmv_vv4<<<g, 256, 0, st>>>(skip_flag, n, rp, col, K, U, out); break;
        case 3: if (variant_vv3 == 1) k_spmv_vv3s<<<g, 256, 0, st>>>(skip_flag, n, rp, col, K, U, out);
                else if (variant_vv3 == 4) k_spmv_vv3c<true><<<g, 256, 0, st>>>(skip_flag, n, rp, col, K, U, out);
                else if (variant_vv3 == 5) k_spmv_vv3c<false><<<g, 256, 0, st>>>(skip_flag, n, rp, col, K, U, out);
                else if (variant_vv3 == 6) k_spmv_vv3c<true, 8><<<g, 256, 0, st>>>(skip_flag, n, rp, col, K, U, out);   // <= 32 registers: 64 warps per SM
                else if (variant_vv3 == 3) k_spmv_vv<3, true><<<g, 256, 0, st>>>(skip_flag, n, rp, col, K, U, out);
                else k_spmv_vv<3><<<g, 256, 0, st>>>(skip_flag, n, rp, col, K, U, out);
                break;
        case 2: k_spmv_vv<2><<<g, 256, 0, st>>>(skip_flag, n, rp, col, K, U, out); break;
        default: k_spmv_vv<1><<<g, 256, 0, st>>>(skip_flag, n, rp, col, K, U, out); break;
      }
      post();
    });
  }
  void spmv_ss(const double* K, const double* U, double* KU)
  {
    Scope sc(*this, KC_SPMV_SS, bytes_ss());
    rows_then_halo(1, KU, 1, [&](int r0, int r1) {
      int t0, t1;
      if (variant_narrow == 2 && tiles_for(r0, r1, t0, t1)) { launch_tiled(t0, t1, K, TileSS{U, KU}); return; }
      k_spmv_ss<<<grid_rows(r1 - r0), 256, 0, st>>>(skip_flag, r1 - r0, rowPtr + r0, col, K, U, KU + r0);
      post();
    });
  }
  void spmv_sv(int dof, const double* K, const double* U, double* KU)
  {
    if (dof != 2 && dof != 3) throw std::runtime_error("spmv_sv: nsd must be 2 or 3");
    Scope sc(*this, KC_SPMV_SV, bytes_svs(dof));
    rows_then_halo(dof, KU, dof, [&](int r0, int r1) {
      const int n = r1 - r0, g = grid_rows(n);
      int t0, t1;
      if (dof == 3 && variant_narrow == 2 && tiles_for(r0, r1, t0, t1)) { launch_tiled(t0, t1, K, TileSV3{U, KU}); return; }
      if (dof == 3) k_spmv_sv<3><<<g, 256, 0, st>>>(skip_flag, n, rowPtr + r0, col, K, U, KU + size_t(r0)*3);
      else k_spmv_sv<2><<<g, 256, 0, st>>>(skip_flag, n, rowPtr + r0, col, K, U, KU + size_t(r0)*2);
      post();
    });
  }
  void spmv_vs(int dof, const double* K, const double* U, double* KU)
  {
    if (dof != 2 && dof != 3) throw std::runtime_error("spmv_vs: nsd must be 2 or 3");
    Scope sc(*this, KC_SPMV_VS, bytes_svs(dof));
    rows_then_halo(1, KU, 1, [&](int r0, int r1) {
      const int n = r1 - r0, g = grid_rows(n);
      int t0, t1;
      if (dof == 3 && variant_narrow == 2 && tiles_for(r0, r1, t0, t1)) { launch_tiled(t0, t1, K, TileVS3{U, KU}); return; }
      if (dof == 3) k_spmv_vs<3><<<g, 256, 0, st>>>(skip_flag, n, rowPtr + r0, col, K, U, KU + r0);
      else k_spmv_vs<2><<<g, 256, 0, st>>>(skip_flag, n, rowPtr + r0, col, K, U, KU + r0);
      post();
    });
  }

  // fsils_commuv / fsils_commus: pack -> grouped ncclSend/ncclRecv -> add in request order
  void halo_exchange(int dof, const double* V, int ld, cudaStream_t s)
  {
    if (dof > halo_dof_cap) throw std::runtime_error("halo buffers too small for dof");
    for (auto& r : reqs) {
      k_halo_pack<<<grid_for(size_t(r.n)*dof, 256, 1), 256, 0, s>>>(r.n, dof, ld, r.ptr, V, r.sbuf);
      post();
    }
    nccl.check(nccl.GroupStart(), "GroupStart");
    for (auto& r : reqs) {
      nccl.check(nccl.Recv(r.rbuf, size_t(r.n)*dof, Nccl::kFloat64, r.peer, comm, s), "Recv");
      nccl.check(nccl.Send(r.sbuf, size_t(r.n)*dof, Nccl::kFloat64, r.peer, comm, s), "Send");
    }
    nccl.check(nccl.GroupEnd(), "GroupEnd");
  }
  void halo_accumulate(int dof, double* V, int ld)
  {
    for (auto& r : reqs) {
      k_halo_add<<<grid_for(size_t(r.n)*dof, 256, 1), 256, 0, st>>>(r.n, dof, ld, r.ptr, r.rbuf, V);
      post();
    }
  }
  void halo_push(int dof, const double* V, int ld)
  {
    if (dof > halo_dof_cap) throw std::runtime_error("halo buffers too small for dof");
    const int g = std::max(1, std::min(kSmCount, (halo_tot*dof + 255)/256));
    k_halo_push<<<g, 256, 0, st>>>(int(reqs.size()), d_peer_reqs, d_halo_ptr_all, halo_tot, dof, halo_dof_cap, ld, V, peer_state);
    post();
  }
  void halo_wait_add(int dof, double* V, int ld)
  {
    const int g = std::max(1, std::min(kSmCount, (halo_nh*dof + 255)/256));
    k_halo_wait_add<<<g, 256, 0, st>>>(int(reqs.size()), d_peer_reqs, halo_nh, d_hn_node, d_hn_ptr, d_hn_src, dof, halo_dof_cap, ld, V, peer_state);
    post();
  }
  void halo_add(int dof, double* V, int ld = 0)
  {
    if (nranks == 1 || reqs.empty()) return;
    if (ld == 0) ld = dof;
    double hb = 0; for (auto& r : reqs) hb += 32.0*r.n*dof;
    Scope sc(*this, KC_HALO, hb, p2p ? 2 : int(reqs.size())*2);
    if (p2p) { halo_push(dof, V, ld); halo_wait_add(dof, V, ld); return; }
    halo_exchange(dof, V, ld, st);
    halo_accumulate(dof, V, ld);
  }

  // ---- peer-mapped transport: collective set-up (every rank of the communicator calls it, also ranks without neighbours) ----
  struct PeerRec {                          // what every rank publishes (gathered with ncclAllGather)
    cudaIpcMemHandle_t handle;
    long long halo_off[kPeerMaxRanks];      // byte offset in MY window where rank p's values arrive, -1: no exchange with p
    int req_idx[kPeerMaxRanks];             // my request index for rank p (= my flag slot for it)
    int req_n[kPeerMaxRanks];
    int ok;                                 // 0: this rank cannot use the peer transport
    int pad;
  };
  void peer_teardown()
  {
    for (int p = 0; p < int(peer_win.size()); p++)
      if (peer_win[p] && p != rank) cudaIpcCloseMemHandle(peer_win[p]);
    peer_win.clear();
    cudaFree(win); win = nullptr; win_bytes = 0;
    cudaFree(peer_state); peer_state = nullptr;
    cudaFree(d_peer_reqs); d_peer_reqs = nullptr;
    cudaFree(d_halo_ptr_all); d_halo_ptr_all = nullptr;
    cudaFree(d_hn_node); cudaFree(d_hn_ptr); cudaFree(d_hn_src);
    d_hn_node = nullptr; d_hn_ptr = nullptr; d_hn_src = nullptr;
    p2p = false;
  }
  // host_lists[i]: the overlap list of request i (solver row ids)
  void peer_setup(const std::vector<std::vector<int>>& host_lists)
  {
    peer_teardown();
    if (nranks == 1) { p2p_why = "single rank"; return; }
    const char* env = getenv("SVB200_P2P");
    PeerRec me;
    std::memset(&me, 0, sizeof(me));
    me.ok = 1;
    std::string why;
    if (env && std::string(env) == "0") { me.ok = 0; why = "disabled by SVB200_P2P=0"; }
    if (nranks > kPeerMaxRanks || int(reqs.size()) > kPeerMaxReq) { me.ok = 0; why = "more ranks / neighbours than the window holds"; }
    for (int p = 0; p < kPeerMaxRanks; p++) { me.halo_off[p] = -1; me.req_idx[p] = -1; }
    size_t bytes = PeerLayout::halo_data_off;
    for (int i = 0; i < int(reqs.size()) && me.ok; i++) {
      const int p = reqs[i].peer;
      if (p < 0 || p >= nranks || p == rank || me.req_idx[p] >= 0) { me.ok = 0; why = "overlap lists are not one per neighbour rank"; break; }
      me.halo_off[p] = (long long)bytes;
      me.req_idx[p] = i;
      me.req_n[p] = reqs[i].n;
      bytes += ((sizeof(double)*2*size_t(std::max(1, reqs[i].n))*halo_dof_cap + 255)/256)*256;
    }
    if (me.ok) {
      if (cudaMalloc(&win, bytes) != cudaSuccess) { me.ok = 0; why = "window allocation failed"; win = nullptr; cudaGetLastError(); }
      else {
        win_bytes = bytes;
        CU_CHECK(cudaMemset(win, 0, bytes));
        CU_CHECK(cudaDeviceSynchronize());
        if (cudaIpcGetMemHandle(&me.handle, win) != cudaSuccess) { me.ok = 0; why = "cudaIpcGetMemHandle failed"; cudaGetLastError(); }
      }
    }
    // gather everybody's record
    PeerRec* d_rec = nullptr;
    CU_CHECK(cudaMalloc(&d_rec, sizeof(PeerRec)*(nranks + 1)));
    CU_CHECK(cudaMemcpyAsync(d_rec + nranks, &me, sizeof(PeerRec), cudaMemcpyHostToDevice, st));
    nccl.check(nccl.AllGather(d_rec + nranks, d_rec, sizeof(PeerRec), Nccl::kInt8, comm, st), "AllGather");
    std::vector<PeerRec> rec(nranks);
    CU_CHECK(cudaMemcpyAsync(rec.data(), d_rec, sizeof(PeerRec)*nranks, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    int ok = 1;
    for (int p = 0; p < nranks; p++) if (!rec[p].ok) ok = 0;
    // the two sides of every exchange must agree
    if (ok) {
      for (int i = 0; i < int(reqs.size()); i++) {
        const int p = reqs[i].peer;
        if (rec[p].req_idx[rank] < 0 || rec[p].req_n[rank] != reqs[i].n) { ok = 0; why = "a neighbour's overlap list does not match ours"; }
      }
    }
    // map the other ranks' windows
    peer_win.assign(nranks, nullptr);
    if (ok) {
      peer_win[rank] = win;
      for (int p = 0; p < nranks && ok; p++) {
        if (p == rank) continue;
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, rec[p].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          ok = 0; why = std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(cudaGetLastError());
        } else peer_win[p] = static_cast<char*>(q);
      }
    }
    // every rank must take the same decision
    {
      int* d_ok = reinterpret_cast<int*>(d_rec);
      CU_CHECK(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, st));
      nccl.check(nccl.AllReduce(d_ok, d_ok, 1, Nccl::kInt32, Nccl::kMin, comm, st), "AllReduce");
      int all_ok = 0;
      CU_CHECK(cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, st));
      CU_CHECK(cudaStreamSynchronize(st));
      if (ok && !all_ok) why = "another rank could not set the peer transport up";
      ok = all_ok;
    }
    cudaFree(d_rec);
    if (!ok) {
      const std::string keep = why.empty() ? std::string("another rank could not set the peer transport up") : why;
      peer_teardown();
      p2p_why = keep;
      return;
    }
    // device tables
    CU_CHECK(cudaMalloc(&peer_state, sizeof(PeerState)));
    CU_CHECK(cudaMemset(peer_state, 0, sizeof(PeerState)));
    std::memset(&red_args, 0, sizeof(red_args));
    red_args.rank = rank; red_args.nranks = nranks;
    for (int p = 0; p < nranks; p++) red_args.win[p] = peer_win[p];
    std::vector<PeerHaloReq> hr(reqs.size());
    std::vector<int> all;
    for (int i = 0; i < int(reqs.size()); i++) {
      const int p = reqs[i].peer;
      hr[i].off = int(all.size());
      hr[i].n = reqs[i].n;
      hr[i].rdata = reinterpret_cast<double*>(peer_win[p] + rec[p].halo_off[rank]);
      hr[i].rflag = reinterpret_cast<unsigned long long*>(peer_win[p] + PeerLayout::halo_flag_off) + rec[p].req_idx[rank];
      hr[i].ldata = reinterpret_cast<const double*>(win + me.halo_off[p]);
      hr[i].lflag = reinterpret_cast<const unsigned long long*>(win + PeerLayout::halo_flag_off) + i;
      all.insert(all.end(), host_lists[i].begin(), host_lists[i].end());
    }
    halo_tot = int(all.size());
    // node-centric source lists: every distinct overlap row with its (request, position) sources in request order
    const HaloSources hsrc = halo_source_lists(nNo_, host_lists);                       // lhs_layout.hpp (pure host, CPU-tested)
    const std::vector<int>& hn_node = hsrc.node;
    const std::vector<int>& hn_ptr = hsrc.ptr;
    std::vector<int2> hn_src(hsrc.src_req.size());
    for (size_t e = 0; e < hn_src.size(); e++) hn_src[e] = make_int2(hsrc.src_req[e], hsrc.src_pos[e]);
    halo_nh = int(hn_node.size());
    auto up = [&](const void* src, size_t nbytes) { void* d = nullptr; CU_CHECK(cudaMalloc(&d, std::max<size_t>(nbytes, 16))); if (nbytes) CU_CHECK(cudaMemcpyAsync(d, src, nbytes, cudaMemcpyHostToDevice, st)); return d; };
    d_peer_reqs = static_cast<PeerHaloReq*>(up(hr.data(), sizeof(PeerHaloReq)*hr.size()));
    d_halo_ptr_all = static_cast<int*>(up(all.data(), sizeof(int)*all.size()));
    d_hn_node = static_cast<int*>(up(hn_node.data(), sizeof(int)*hn_node.size()));
    d_hn_ptr = static_cast<int*>(up(hn_ptr.data(), sizeof(int)*hn_ptr.size()));
    d_hn_src = static_cast<int2*>(up(hn_src.data(), sizeof(int2)*hn_src.size()));
    CU_CHECK(cudaStreamSynchronize(st));
    p2p = true;
    p2p_why = "peer-mapped windows (CUDA IPC over NVLink)";
  }
  // raises when a wait of the peer transport timed out (a rank died or the launch order diverged)
  void peer_check()
  {
    if (!p2p) return;
    PeerState s;
    CU_CHECK(cudaMemcpyAsync(&s, peer_state, sizeof(PeerState), cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    if (s.error) {
      CU_CHECK(cudaMemsetAsync(&peer_state->error, 0, sizeof(int), st));
      throw std::runtime_error(s.error == 1 ? "overlap-node exchange timed out waiting for a neighbour rank"
                               : s.error == 2 ? "all-reduce timed out waiting for another rank"
                                              : "fused product + exchange kernel: grid barrier timed out (not every CTA resident)");
    }
  }

  // ---- faces ----------------------------------------------------------------------------------------
  int n_faces() const { return int(faces.size()); }
  bool face_coupled(int f) const { return faces[f].coupled; }
  bool face_inc(int f) const { return faces[f].inc; }
  int face_bgrp(int f) const { return faces[f].bGrp; }
  void face_set_inc(int f, bool v) { faces[f].inc = v; }
  void face_set_coupled(int f, bool c, double res) { faces[f].coupled = c; if (c) faces[f].res = res; }

  void face_dot(const DevFace& fa, int m, int ld, int lim, const double* X, double* out)
  {
    const int n = fa.nNo*m;
    const int g = std::max(1, std::min(kFaceBlocks, (n + 511)/512));
    k_face_dot<<<g, 256, 0, st>>>(fa.nNo, m, fa.dof, ld, lim, fa.glob, fa.valM, X, face_partial_d, counter_d + 1, out);
    post();
  }
  // face.nS = ||valM||^2 (ns_solver.cpp:56-87, gmres.cpp:50-85)
  void bc_pre(int nsd)
  {
    int nslot = 0;
    std::vector<int> which;
    for (int f = 0; f < n_faces(); f++) {
      auto& fa = faces[f];
      if (!fa.coupled) continue;
      const int m = std::min(fa.dof, nsd);
      const int lim = fa.shared ? mynNo_ : nNo_;
      face_dot(fa, m, nsd, lim, nullptr, red_d + nslot);
      which.push_back(f);
      nslot++;
    }
    if (nslot == 0) return;
    // shared faces complete their norm with an all-reduce; for a face owned by one rank the other
    // ranks hold no nodes of it (nNo = 0 -> 0 contribution), so one reduction serves both cases
    // only when every coupled face is shared; otherwise reduce face by face.
    if (nranks > 1) {
      for (int k = 0; k < nslot; k++) {
        if (faces[which[k]].shared) allreduce(red_d + k, 1);
      }
    }
    std::vector<double> v(nslot);
    reduce_fetch(nslot, v.data());
    for (int k = 0; k < nslot; k++) faces[which[k]].nS = v[k];
  }

  // Y += coef * v (v^T X), v = valM on the face nodes (add_bc_mul.cpp:53-121); X may alias Y
  // (ld: leading dimension of X and Y when it differs from dof)
  void add_bc_mul(int op, int dof, const double* X, double* Y, int ld = 0)
  {
    if (ld == 0) ld = dof;
    for (int f = 0; f < n_faces(); f++) {
      auto& fa = faces[f];
      if (!fa.coupled) continue;
      const double coef = (op == BCOP_ADD) ? fa.res : -fa.res/(1.0 + fa.res*fa.nS);
      const int m = std::min(fa.dof, dof);
      const int lim = fa.shared ? mynNo_ : nNo_;
      if (!(fa.shared && nranks > 1)) {
        // the face lives on this rank alone: ranks that hold none of it have nothing to do, the owner does both stages in one launch
        if (fa.nNo == 0) continue;
        if (variant_face_fused == 2) {             // one multi-CTA launch: same partial sums and order as the two-launch path
          const int n = fa.nNo*m;
          const int g = std::max(1, std::min(kFaceBlocks, (n + 511)/512));
          k_face_rank1_grid<<<g, 256, 0, st>>>(fa.nNo, m, fa.dof, ld, lim, fa.glob, fa.valM, X, coef, Y, face_partial_d, counter_d + 2);
          post();
          continue;
        }
        if (variant_face_fused == 1 && size_t(fa.nNo)*m <= 131072) {
          k_face_rank1<<<1, 1024, 0, st>>>(fa.nNo, m, fa.dof, ld, lim, fa.glob, fa.valM, X, coef, Y);
          post();
          continue;
        }
      }
      double* S = red_d + (kMaxSlots - 1);
      face_dot(fa, m, ld, lim, X, S);
      if (fa.shared && nranks > 1) allreduce(S, 1);
      if (fa.nNo > 0) {
        k_face_axpy<<<grid_for(size_t(fa.nNo)*m, 256, 1), 256, 0, st>>>(fa.nNo, m, fa.dof, ld, fa.glob, fa.valM, coef, S, Y);
        post();
      }
    }
  }

  // ---- preconditioners --------------------------------------------------------------------------------
  // precond_diag (precond.cpp:122-256): W = diag -> overlap add -> 0->1 -> 1/sqrt|W| -> Dirichlet mask;
  // Val <- W Val W (one pass), R <- W R, valM = val W for coupled faces.
  void precond_diag(int dof, double* Val, double* R, double* W)
  {
    const size_t n = size_t(dof)*nNo_;
    k_diag_extract<<<grid_for(n, 256), 256, 0, st>>>(nNo_, dof, diag, Val, W); post();
    halo_add(dof, W);
    k_w_invsqrt<<<grid_for(n, 256), 256, 0, st>>>(n, W); post();
    for (auto& fa : faces) {
      if (!fa.inc || fa.bGrp != B200_BC_DIR || fa.nNo == 0) continue;
      const int m = std::min(fa.dof, dof);
      k_face_mask<<<grid_for(size_t(fa.nNo)*m, 256, 1), 256, 0, st>>>(fa.nNo, m, fa.dof, dof, fa.glob, fa.val, W); post();
    }
    scale_val(dof, W, W, Val);
    mul_inplace(n, W, R);
    for (auto& fa : faces) {
      if (!fa.coupled || fa.nNo == 0) continue;
      const int m = std::min(fa.dof, dof);
      k_face_valM<<<grid_for(size_t(fa.nNo)*m, 256, 1), 256, 0, st>>>(fa.nNo, m, fa.dof, dof, fa.glob, fa.val, W, fa.valM); post();
    }
  }
  void scale_val(int dof, const double* Wr, const double* Wc, double* Val)
  {
    const int g = grid_rows(nNo_);
    Scope sc(*this, KC_SCALE_VAL, double(nnz_)*(16.0*dof*dof + 4.0) + double(nNo_)*(16.0*dof + 8.0));
    switch (dof) {
      case 4: k_scale_val<4><<<g, 256, 0, st>>>(nNo_, rowPtr, col, Wr, Wc, Val); break;
      case 3: k_scale_val<3><<<g, 256, 0, st>>>(nNo_, rowPtr, col, Wr, Wc, Val); break;
      case 2: k_scale_val<2><<<g, 256, 0, st>>>(nNo_, rowPtr, col, Wr, Wc, Val); break;
      case 1: k_scale_val<1><<<g, 256, 0, st>>>(nNo_, rowPtr, col, Wr, Wc, Val); break;
      default: throw std::runtime_error("scale_val: dof > 4");
    }
    post();
  }
  // precond_rcs (precond.cpp:266-540): Dirichlet rows/columns killed and given a unit diagonal, then up
  // to 10 sweeps of row / column max-norm scaling.  W1 (row) and W2 (column) accumulate the scalings;
  // R <- W1 R at the end, the caller multiplies the solution by W2.  Like the reference it never
  // touches face.valM (a coupled face therefore contributes nothing under this preconditioner).
  void precond_rcs(int dof, double* Val, double* R, double* W1, double* W2)
  {
    if (dof < 1 || dof > 4) throw std::runtime_error("precond_rcs: dof > 4");
    const size_t n = size_t(dof)*nNo_;
    auto mk = mark();
    double* Wr = vec(n);
    double* Wc = vec(n);
    fill(n, 1.0, W1);
    fill(n, 1.0, W2);
    fill(n, 1.0, Wr);
    for (auto& fa : faces) {
      if (!fa.inc || fa.bGrp != B200_BC_DIR || fa.nNo == 0) continue;
      const int m = std::min(fa.dof, dof);
      k_face_mask<<<grid_for(size_t(fa.nNo)*m, 256, 1), 256, 0, st>>>(fa.nNo, m, fa.dof, dof, fa.glob, fa.val, Wr); post();
    }
    halo_add(dof, Wr);
    k_rcs_renorm<<<grid_for(n, 256), 256, 0, st>>>(n, Wr); post();
    scale_val(dof, Wr, Wr, Val);                      // pre_mul + pos_mul in one pass, same rounding
    mul_inplace(n, Wr, R);
    k_rcs_unit_diag<<<grid_for(n, 256), 256, 0, st>>>(nNo_, dof, diag, Wr, Val); post();

    const int maxiter = 10;
    const double tol = 2.0;
    int iter = 0;
    bool flag = true;
    const int g = grid_rows(nNo_);
    while (flag) {
      zero(n, Wc);
      iter++;
      if (iter >= maxiter) flag = false;     // (the reference prints a warning here)
      {
        Scope sc(*this, KC_SCALE_VAL, double(nnz_)*(8.0*dof*dof + 4.0) + double(nNo_)*(16.0*dof + 8.0));
        switch (dof) {
          case 4: k_rcs_norms<4><<<g, 256, 0, st>>>(nNo_, rowPtr, col, Val, Wr, Wc); break;
          case 3: k_rcs_norms<3><<<g, 256, 0, st>>>(nNo_, rowPtr, col, Val, Wr, Wc); break;
          case 2: k_rcs_norms<2><<<g, 256, 0, st>>>(nNo_, rowPtr, col, Val, Wr, Wc); break;
          default: k_rcs_norms<1><<<g, 256, 0, st>>>(nNo_, rowPtr, col, Val, Wr, Wc); break;
        }
        post();
      }
      halo_add(dof, Wr);
      halo_add(dof, Wc);
      zero(2, red_d);
      k_rcs_dev1<<<grid_for(n, 256), 256, 0, st>>>(n, Wr, red_d); post();
      k_rcs_dev1<<<grid_for(n, 256), 256, 0, st>>>(n, Wc, red_d + 1); post();
      double dev[2];
      reduce_fetch(2, dev);
      // NaN-safe like the reference's `max(...) < tol` (a NaN keeps the flag)
      if ((dev[0] < tol) && (dev[1] < tol)) flag = false;
      k_rcs_invsqrt_accum<<<grid_for(n, 256), 256, 0, st>>>(n, Wr, W1); post();
      k_rcs_invsqrt_accum<<<grid_for(n, 256), 256, 0, st>>>(n, Wc, W2); post();
      scale_val(dof, Wr, Wc, Val);
      if (nranks > 1) {
        // MPI_Allgather of the flags + any() (:529-534)
        fill(1, flag ? 1.0 : 0.0, red_d);
        allreduce(red_d, 1, true);
        double f;
        reduce_fetch(1, &f);
        flag = f > 0.5;
      }
    }
    mul_inplace(n, W1, R);
    release(mk);
  }

  // ---- NS helpers ---------------------------------------------------------------------------------------
  // packed copies made by depart for the fused Schur operator (valid until the arena mark of the
  // NS solve is released): GtL(4,nnz) = [Gt, L], V4(4,nNo) = [G P, P]
  const double* packed_Gt = nullptr;
  double* GtL = nullptr;
  double* V4 = nullptr;
  double* Gs = nullptr;       // component-wise copy of mG (variant_gp == 3)

  void depart(int nsd, const double* Val, double* Gt, double* mK, double* mG, double* mD, double* mL)
  {
    const size_t nz = size_t(nnz_);
    packed_Gt = nullptr; GtL = nullptr; V4 = nullptr;
    Gs = nullptr;
    if (nsd == 3) {
      GtL = vec(4*nz);
      V4 = vec(4*size_t(nNo_));
      packed_Gt = Gt;
      if (variant_gp == 3) Gs = vec(3*nz);             // component-wise copy of mG for pass 1
    }
    Scope sc(*this, KC_DEPART, double(nnz_)*(8.0*(nsd+1)*(nsd+1)*2 + 8.0*nsd + 4.0 + (nsd == 3 ? 32.0 : 0.0)));
    if (nsd == 3) k_depart3<<<grid_for(nz, 256, 1), 256, 0, st>>>(nz, tpos, Val, Gt, mK, mG, mD, mL, GtL, Gs);
    else if (nsd == 2) k_depart_generic<2><<<grid_for(nz, 256, 1), 256, 0, st>>>(nz, tpos, Val, Gt, mK, mG, mD, mL);
    else throw std::runtime_error("FSILS: Not defined nsd for DEPART");
    post();
  }
  double bytes_schur_gp() const { return double(nnz_)*28.0 + double(nNo_)*(8.0 + 8.0 + 32.0 + 8.0); }
  double bytes_schur_sp() const { return double(nnz_)*36.0 + double(nNo_)*(32.0 + 8.0 + 8.0); }

  // SP = L P - Gt (G P), with the resistance preconditioner applied to G P when a face is coupled
  // (cgrad.cpp:96-113).  GP, DGP: caller work vectors used by the unfused path only.
  void schur_op(int nsd, const double* Gt, const double* G, const double* L, const double* P, double* GP, double* DGP,
                double* SP, bool coupled)
  {
    if (nsd == 3 && Gt == packed_Gt && GtL) {
      if (use_fused() && (variant_gp == 0 || (variant_gp == 3 && Gs))) {
        Scope sc(*this, KC_SPMV_SV, bytes_schur_gp());
        if (variant_gp == 3) launch_fused(2, RowsGP{rowPtr, col, Gs, P, V4, size_t(nnz_)}, 3, V4, 4);
        else launch_fused(2, RowsGP{rowPtr, col, G, P, V4}, 3, V4, 4);
      } else {
        Scope sc(*this, KC_SPMV_SV, bytes_schur_gp());
        rows_then_halo(3, V4, 4, [&](int r0, int r1) {
          const int n = r1 - r0, g = grid_rows(n);
          int t0, t1;
          if (variant_gp == 2 && tiles_for(r0, r1, t0, t1)) { launch_tiled(t0, t1, G, TileGP{P, P, V4}); return; }
          if (variant_gp == 3 && Gs) k_schur_gp_soa<<<g, 256, 0, st>>>(skip_flag, n, rowPtr + r0, col, Gs, size_t(nnz_), P, P + r0, V4 + size_t(r0)*4);
          else if (variant_gp == 1) k_schur_gp4<<<g, 256, 0, st>>>(skip_flag, n, rowPtr + r0, col, G, P, P + r0, V4 + size_t(r0)*4);
          else k_schur_gp<<<g, 256, 0, st>>>(skip_flag, n, rowPtr + r0, col, G, P, P + r0, V4 + size_t(r0)*4);
          post();
        });
      }
      if (coupled) add_bc_mul(BCOP_PRE, 3, V4, V4, 4);
      if (use_fused() && variant_sp == 1) {
        Scope sc(*this, KC_SPMV_VS, bytes_schur_sp());
        launch_fused(3, RowsSP{rowPtr, col, GtL, V4, SP}, 1, SP, 1);
      } else {
        Scope sc(*this, KC_SPMV_VS, bytes_schur_sp());
        rows_then_halo(1, SP, 1, [&](int r0, int r1) {
          const int n = r1 - r0, g = grid_rows(n);
          int t0, t1;
          if (variant_sp == 2 && tiles_for(r0, r1, t0, t1)) { launch_tiled(t0, t1, GtL, TileSP{V4, SP}); return; }
          if (variant_sp == 1) k_schur_sp4<<<g, 256, 0, st>>>(skip_flag, n, rowPtr + r0, col, GtL, V4, SP + r0);
          else k_schur_sp<<<g, 256, 0, st>>>(skip_flag, n, rowPtr + r0, col, GtL, V4, SP + r0);
          post();
        });
      }
      return;
    }
    spmv_sv(nsd, G, P, GP);
    if (coupled) add_bc_mul(BCOP_PRE, nsd, GP, GP);
    spmv_vs(nsd, Gt, GP, DGP);
    spmv_ss(L, P, SP);
    axpy(size_t(nNo_), -1.0, DGP, SP);
  }
  // ---- device-resident CG loops --------------------------------------------------------------------------
  // Generic driver: `apply(P, SP)` enqueues SP = A P (with its overlap adds); the scalars stay on the device
  // (k_cg_head / k_cg_update / k_cg_pupdate) and the host polls the state every cg_batch iterations, one batch
  // behind, so the stream never drains.  Same arithmetic as the host-driven loop of krylov.hpp.
  template <class Apply>
  void cg_device(SubLs& ls, int dof, double* R, double* X, double* P, double* SP, Apply&& apply, int& last_i, double& err, double& errO)
  {
    const size_t n = size_t(dof)*nNo_, nOwn = size_t(dof)*mynNo_;
    CgState init;
    init.err = err; init.errO = errO; init.eps = std::pow(std::max(ls.absTol, ls.relTol*ls.iNorm), 2.0);
    init.done = 0; init.suc = 0; init.last_i = 0; init.pad = 0;
    cg_h[2] = init;
    CU_CHECK(cudaMemcpyAsync(cg_d, &cg_h[2], sizeof(CgState), cudaMemcpyHostToDevice, st));
    const int g = int(std::min<size_t>(size_t(kRedBlocks), (n + 1023)/1024 + 1));
    int enq = 0, slot = 0, pending = -1;
    bool stop = (ls.mItr <= 0);
    skip_flag = &cg_d->done;
    std::vector<ProfCount> marks;                 // counters before iteration i (profiling only)
    while (!stop) {
      const int nb = std::min(cg_batch, ls.mItr - enq);
      for (int k = 0; k < nb; k++) {
        if (profiling) marks.push_back(prof_snapshot());
        k_cg_head<<<1, 1, 0, st>>>(cg_d, enq + k); post();
        apply(P, SP);
        dots_reduce(dof, 1, P, 0, SP, 0);
        {
          Scope sc(*this, KC_BLAS1, 48.0*double(n));
          if (fuse_reduce()) k_cg_update<<<g, kRedThreads, 0, st>>>(n, nOwn, cg_d, red_d, P, SP, X, R, partial_d, counter_d, red_d + 1, red_args, peer_state);
          else k_cg_update<<<g, kRedThreads, 0, st>>>(n, nOwn, cg_d, red_d, P, SP, X, R, partial_d, counter_d, red_d + 1);
          post();
        }
        if (!fuse_reduce()) allreduce(red_d + 1, 1);
        {
          Scope sc(*this, KC_BLAS1, 24.0*double(n));
          k_cg_pupdate<<<grid_for(n, 256), 256, 0, st>>>(n, cg_d, red_d + 1, R, P); post();
        }
      }
      enq += nb;
      CU_CHECK(cudaMemcpyAsync(&cg_h[slot], cg_d, sizeof(CgState), cudaMemcpyDeviceToHost, st));
      CU_CHECK(cudaEventRecord(cg_ev[slot], st));
      if (pending >= 0) {
        CU_CHECK(cudaEventSynchronize(cg_ev[pending]));
        if (cg_h[pending].done) stop = true;
      }
      pending = slot; slot ^= 1;
      if (enq >= ls.mItr) stop = true;
    }
    // the reference leaves the loop with last_i = mItr - 1 when it never converges; a final head would only
    // matter for the `suc` flag of an iterate that converged in the very last iteration, which the reference
    // does not report either (it tests at the top of an iteration)
    skip_flag = nullptr;
    CU_CHECK(cudaMemcpyAsync(&cg_h[2], cg_d, sizeof(CgState), cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    ls.suc = cg_h[2].suc != 0;
    last_i = cg_h[2].last_i;
    {
      // bodies ran for the iterations before the one whose head found convergence (all enqueued ones when it never did)
      const int executed = ls.suc ? last_i : last_i + 1;
      if (profiling && int(marks.size()) > executed) prof_uncredit(marks[executed], prof_snapshot());
    }
    err = cg_h[2].err;
    errO = cg_h[2].errO;
  }

  // ---- device-resident Arnoldi loop ---------------------------------------------------------------------------------------------
  // One restart cycle of gmres / gmres_v (krylov.hpp) without a host round trip per iteration: `step(i)` enqueues the product, the
  // resistance-face terms, the dots (+ all-reduce) and the Gram-Schmidt update of iteration i; k_gmres_givens then does the Givens
  // bookkeeping and the convergence test on the device.  The host enqueues iterations in batches and polls the state one batch
  // behind (as the CG loops do); iterations enqueued past convergence return at once (skip flag).  On return hs holds the columns
  // 0..last_i of the Hessenberg matrix and the residual estimates, as the host loop would have left them.
  bool gmres_device_ok() const { return variant_gmres_device != 0; }
  static constexpr int kGivensMaxHost = kGivensMax;
  template <class Step>
  int gmres_device_cycle(int sD, double eps, double err0, Step&& step, Hessenberg& hs, bool& suc)
  {
    if (sD > gm_sD || !gm_d) {
      if (!gm_d) {
        CU_CHECK(cudaMalloc(&gm_d, sizeof(GmresState)));
        CU_CHECK(cudaMallocHost(&gm_h, 3*sizeof(GmresState)));
        CU_CHECK(cudaEventCreateWithFlags(&gm_ev[0], cudaEventDisableTiming));
        CU_CHECK(cudaEventCreateWithFlags(&gm_ev[1], cudaEventDisableTiming));
      }
      cudaFree(gm_arr); cudaFreeHost(gm_host);
      const size_t nd = size_t(sD + 1)*sD + 3*size_t(sD) + 2;
      CU_CHECK(cudaMalloc(&gm_arr, sizeof(double)*nd));
      CU_CHECK(cudaMallocHost(&gm_host, sizeof(double)*nd));
      gm_sD = sD;
    }
    double* d_h = gm_arr;
    double* d_c = d_h + size_t(gm_sD + 1)*gm_sD;
    double* d_s = d_c + gm_sD;
    double* d_err = d_s + gm_sD;
    GmresState init;
    init.eps = eps; init.err0 = err0; init.done = 0; init.suc = 0; init.last_i = 0; init.pad = 0;
    gm_h[2] = init;
    CU_CHECK(cudaMemcpyAsync(gm_d, &gm_h[2], sizeof(GmresState), cudaMemcpyHostToDevice, st));
    const int* outer_skip = skip_flag;
    skip_flag = &gm_d->done;
    int enq = 0, slot = 0, pending = -1;
    bool stop = (sD <= 0);
    std::vector<ProfCount> marks;                 // counters before iteration i (profiling only) + one after the last
    while (!stop) {
      const int nb = std::min(gm_batch, sD - enq);
      for (int k = 0; k < nb; k++) {
        if (profiling) marks.push_back(prof_snapshot());
        // (letting the Givens bookkeeping ride on the Gram-Schmidt update kernel as an extra CTA was measured and rejected: the 24 KB
        // of static shared memory it brings into that kernel cost the streaming update 50 % of its bandwidth, 0.39 vs 0.26 ms per
        // launch at P10; profiles/r02_bench_n1_givens_ride.json)
        step(enq + k);
        k_gmres_givens<<<1, 256, 0, st>>>(gm_d, enq + k, sD, red_d, d_h, d_c, d_s, d_err); post();
      }
      enq += nb;
      CU_CHECK(cudaMemcpyAsync(&gm_h[slot], gm_d, sizeof(GmresState), cudaMemcpyDeviceToHost, st));
      CU_CHECK(cudaEventRecord(gm_ev[slot], st));
      if (pending >= 0) {
        CU_CHECK(cudaEventSynchronize(gm_ev[pending]));
        if (gm_h[pending].done) stop = true;
      }
      pending = slot; slot ^= 1;
      if (enq >= sD) stop = true;
    }
    skip_flag = outer_skip;
    CU_CHECK(cudaMemcpyAsync(&gm_h[2], gm_d, sizeof(GmresState), cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    const int last_i = gm_h[2].last_i;
    suc = gm_h[2].suc != 0;
    if (profiling && int(marks.size()) > last_i + 1) prof_uncredit(marks[last_i + 1], prof_snapshot());    // iterations > last_i were skipped
    // the columns the back substitution needs + the residual estimates
    CU_CHECK(cudaMemcpyAsync(gm_host, d_h, sizeof(double)*size_t(last_i + 1)*(sD + 1), cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaMemcpyAsync(gm_host + size_t(sD + 1)*sD, d_err, sizeof(double)*(last_i + 2), cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    std::memcpy(hs.h.data(), gm_host, sizeof(double)*size_t(last_i + 1)*(sD + 1));
    std::memcpy(hs.err.data(), gm_host + size_t(sD + 1)*sD, sizeof(double)*(last_i + 2));
    return last_i;
  }

  void split_mc(int dof, const double* Ri, double* Rm, double* Rc)
  {
    k_split_mc<<<grid_for(size_t(nNo_)*dof, 256), 256, 0, st>>>(nNo_, dof, Ri, Rm, Rc); post();
  }
  void join_mc(int dof, const double* Rm, const double* Rc, double* Ri)
  {
    k_join_mc<<<grid_for(size_t(nNo_)*dof, 256), 256, 0, st>>>(nNo_, dof, Rm, Rc, Ri); post();
  }

  void phase_mark(int which, double t0)
  {
    CU_CHECK(cudaStreamSynchronize(st));
    phase_ms[which] = (wall_s() - t0)*1e3;
  }
};

} // namespace svb200
