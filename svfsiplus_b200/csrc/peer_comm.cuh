// peer_comm.cuh — overlap-node add (fsils_commuv / fsils_commus, liner_solver/in_commu.cpp:111-170) and the Krylov
// all-reduces (MPI_Allreduce in liner_solver/dot.cpp, norm.cpp, bcast.cpp:51-58) done by OUR kernels over peer-mapped
// memory: every rank owns one window in its HBM that all other ranks of the node map through CUDA IPC, and the
// exchange is plain NVLink stores into the receiver's window followed by a release-store of an epoch flag.  No
// library call, no second stream, no event hop sits between a boundary-row SpMV and its exchange any more:
//
//   boundary rows -> k_halo_push (pack straight into the neighbours' windows, flag) -> interior rows
//                 -> k_halo_wait_add (acquire the neighbours' flags, add in request order)
//
//   k_peer_allreduce: one CTA writes its partial sums into everybody's mailbox (slot = own rank), flags, waits for the
//   other ranks' flags and sums the mailbox in RANK ORDER - every rank computes bit-identical results, which the
//   device-resident convergence flags of the CG loops rely on.
//
// Buffers are double-buffered on the epoch's parity: rank A can only start epoch k+2 after its wait of epoch k+1, which
// needs B's push of k+1, which B's stream orders after B's reads of epoch k.  Communication kernels therefore NEVER skip
// (the `skip` flag of a finished device-resident loop only silences compute kernels): every rank executes every epoch.
// Epochs live in device memory (PeerState), so the launches carry no per-call arguments that change.
//
// A wait that does not see its flag within kSpinLimit clock ticks (~seconds) records an error in PeerState instead of
// hanging the GPU; the host reports it at the end of the solve.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace svb200 {

constexpr int kPeerMaxRanks = 16;
constexpr int kPeerMaxReq = 16;
constexpr int kPeerSlots = 1024;                    // doubles per rank in the reduction mailbox (= CudaOps::kMaxSlots)
constexpr long long kSpinLimit = 8000000000ll;      // clock64 ticks (~4 s at 1.9 GHz)

struct PeerState {                 // device memory, one per CudaOps
  unsigned long long red_epoch;    // last completed all-reduce epoch
  unsigned long long halo_epoch;   // last completed halo epoch
  unsigned int push_count;         // CTA counters of the two halo kernels
  unsigned int wait_count;
  int error;                       // 1: halo wait timed out, 2: all-reduce wait timed out
  unsigned int bar_count;          // grid barrier of the fused product + exchange kernel (fused_halo.cuh)
};

// Layout of a rank's window (all offsets in bytes, 256-byte aligned):
//   red_flag[2][kPeerMaxRanks]   unsigned long long   (written by rank r into slot [parity][r])
//   halo_flag[2][kPeerMaxReq]    unsigned long long   (written by the sender of request j into [parity][j])
//   red_mail[2][kPeerMaxRanks][kPeerSlots] double
//   halo data: request j at halo_off[j], 2 parities x n_j x dofcap doubles
struct PeerLayout {
  static constexpr size_t red_flag_off = 0;
  static constexpr size_t halo_flag_off = 256;
  static constexpr size_t red_mail_off = 512;
  static constexpr size_t halo_data_off = red_mail_off + sizeof(double)*2*kPeerMaxRanks*kPeerSlots;
};

struct PeerRedArgs {               // kernel argument of k_peer_allreduce (by value)
  int rank, nranks;
  char* win[kPeerMaxRanks];        // every rank's window as mapped here (own window: the local pointer)
};

struct PeerHaloReq {               // device array, one entry per request (neighbour)
  int off, n;                      // this request's slice of the concatenated node list
  double* rdata;                   // neighbour's window: where OUR values for it go (parity 0; parity 1 = + n*dofcap)
  unsigned long long* rflag;       // neighbour's flag slot for this exchange (parity 0; parity 1 = + kPeerMaxReq)
  const double* ldata;             // own window: where the neighbour's values arrive
  const unsigned long long* lflag; // own flag slot
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p)
{
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// data written by a peer over NVLink lands in this GPU's L2: read it past L1
__device__ __forceinline__ double ld_peer_written(const double* p)
{
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// once a wait has timed out (ps->error set) later waits do not spin again: the solve finishes quickly with garbage and the host
// raises the error (CudaOps::peer_check) instead of the GPU sitting in 4-second waits for the rest of the solve
__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long epoch, const PeerState* ps)
{
  if (*reinterpret_cast<const volatile int*>(&ps->error) != 0) return false;
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < epoch) {
    if (clock64() - t0 > kSpinLimit) return false;
    __nanosleep(20);
  }
  return true;
}

// ---- all-reduce of n <= kPeerSlots doubles at `v` (sum or max), one CTA -------------------------------------------------
// The body is a device function so that the LAST CTA of a reduction kernel (k_multi_dot, k_cg_update) can run it as its
// epilogue: local reduction + all-reduce in one launch.  Every thread of the CTA must call it.
template <int OP>   // 0 sum, 1 max
__device__ __forceinline__ void peer_allreduce_body(const PeerRedArgs& a, PeerState* ps, double* __restrict__ v, int n)
{
  __shared__ unsigned long long s_epoch;
  __shared__ int s_ok;
  __syncthreads();
  if (threadIdx.x == 0) { s_epoch = ps->red_epoch + 1; s_ok = 1; }
  __syncthreads();
  const unsigned long long epoch = s_epoch;
  const int par = int(epoch & 1ull);
  const size_t mail = PeerLayout::red_mail_off + sizeof(double)*(size_t(par)*kPeerMaxRanks + a.rank)*kPeerSlots;
  // 1. own partials into every rank's mailbox (including our own: the sum below then treats all ranks alike)
  for (int t = threadIdx.x; t < n*a.nranks; t += blockDim.x) {
    const int p = t / n, i = t - p*n;
    reinterpret_cast<double*>(a.win[p] + mail)[i] = __ldcg(v + i);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < a.nranks) {
    const int p = threadIdx.x;
    st_release_sys(reinterpret_cast<unsigned long long*>(a.win[p] + PeerLayout::red_flag_off) + par*kPeerMaxRanks + a.rank, epoch);
    // 2. wait for rank p's partials to arrive in OUR mailbox
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(a.win[a.rank] + PeerLayout::red_flag_off) + par*kPeerMaxRanks + p;
    if (!spin_until(f, epoch, ps)) s_ok = 0;
  }
  __syncthreads();
  // 3. rank-ordered reduction: bit-identical on every rank
  const double* m = reinterpret_cast<const double*>(a.win[a.rank] + PeerLayout::red_mail_off) + size_t(par)*kPeerMaxRanks*kPeerSlots;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = ld_peer_written(m + i);
    for (int p = 1; p < a.nranks; p++) {
      const double x = ld_peer_written(m + size_t(p)*kPeerSlots + i);
      if (OP == 0) s += x; else s = (x > s) ? x : s;
    }
    v[i] = s;
  }
  if (threadIdx.x == 0) { ps->red_epoch = epoch; if (!s_ok) ps->error = 2; }
}

template <int OP>
__global__ void __launch_bounds__(256) k_peer_allreduce(PeerRedArgs a, PeerState* ps, double* __restrict__ v, int n)
{
  peer_allreduce_body<OP>(a, ps, v, n);
}

// ---- halo push: pack the overlap rows of V straight into the neighbours' windows -----------------------------------------
__global__ void __launch_bounds__(256) k_halo_push(int nreq, const PeerHaloReq* __restrict__ reqs, const int* __restrict__ ptr_all,
                                                   int tot, int dof, int dofcap, int ld, const double* __restrict__ V, PeerState* ps)
{
  const unsigned long long epoch = ps->halo_epoch + 1;     // advanced by the matching k_halo_wait_add
  const int par = int(epoch & 1ull);
  const int total = tot*dof;
  for (int t = blockIdx.x*blockDim.x + threadIdx.x; t < total; t += gridDim.x*blockDim.x) {
    const int j = t / dof, l = t - j*dof;
    int r = 0;
    while (r + 1 < nreq && j >= reqs[r+1].off) r++;
    const PeerHaloReq q = reqs[r];
    q.rdata[size_t(par)*q.n*dofcap + size_t(j - q.off)*dof + l] = V[size_t(ptr_all[j])*ld + l];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(&ps->push_count, 1u);
    s_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if (threadIdx.x < nreq) st_release_sys(reqs[threadIdx.x].rflag + par*kPeerMaxReq, epoch);
    if (threadIdx.x == 0) ps->push_count = 0;
  }
}

// ---- halo wait + add: V[node] += sum over the requests that hold the node, in request order (in_commu.cpp:150-168) --------
// hn_node[nh]: the distinct overlap rows; hn_ptr[nh+1] / hn_src: for each, its (request, position) sources in request order,
// encoded as request << 26 | position... positions can exceed 2^26 on large interfaces, so two ints per source are used.
__global__ void __launch_bounds__(256) k_halo_wait_add(int nreq, const PeerHaloReq* __restrict__ reqs, int nh, const int* __restrict__ hn_node,
                                                       const int* __restrict__ hn_ptr, const int2* __restrict__ hn_src, int dof, int dofcap,
                                                       int ld, double* __restrict__ V, PeerState* ps)
{
  const unsigned long long epoch = ps->halo_epoch + 1;
  const int par = int(epoch & 1ull);
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < nreq) {
    if (!spin_until(reqs[threadIdx.x].lflag + par*kPeerMaxReq, epoch, ps)) s_ok = 0;
  }
  __syncthreads();
  const int total = nh*dof;
  for (int t = blockIdx.x*blockDim.x + threadIdx.x; t < total; t += gridDim.x*blockDim.x) {
    const int k = t / dof, l = t - k*dof;
    double* dst = V + size_t(hn_node[k])*ld + l;
    double s = *dst;
    for (int e = hn_ptr[k]; e < hn_ptr[k+1]; e++) {
      const int2 src = hn_src[e];
      const PeerHaloReq& q = reqs[src.x];
      s += ld_peer_written(q.ldata + size_t(par)*q.n*dofcap + size_t(src.y)*dof + l);
    }
    *dst = s;
  }
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    if (!s_ok) ps->error = 1;
    __threadfence();
    const unsigned int done = atomicAdd(&ps->wait_count, 1u);
    s_last = (done == gridDim.x - 1);
    if (s_last) { ps->wait_count = 0; ps->halo_epoch = epoch; }
  }
}

} // namespace svb200
