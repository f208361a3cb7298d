// TEST INFRASTRUCTURE ONLY (oracle).  In-process, thread-per-rank implementation of the
// MPI subset declared in mpi.h.  It exists so that the reference's own FSILS sources
// (/root/reference/Code/Source/liner_solver/*.cpp) can be run here with 1..N "ranks"
// without an MPI runtime: N host threads of one process, collectives through a shared
// barrier, point-to-point through buffered mailboxes.  Reductions are evaluated in rank
// order (0,1,..,N-1) on every rank so all ranks see bit-identical results.
#include "mpi.h"

#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <stdexcept>
#include <tuple>
#include <vector>

namespace {

int g_size = 1;
thread_local int t_rank = 0;

std::mutex g_mtx;
std::condition_variable g_cv;
int g_bar_count = 0;
long g_bar_gen = 0;
std::vector<const void*> g_slots;   // per-rank pointer published for a collective
std::vector<int> g_slot_n;

struct Msg { std::vector<char> data; };
std::map<std::tuple<int,int,int>, std::deque<Msg>> g_mail;   // (src,dst,tag) -> fifo

struct Req { bool is_recv; void* buf; size_t bytes; int src; int tag; bool done; };
thread_local std::vector<Req> t_reqs;

size_t tsize(MPI_Datatype t)
{
  switch (t) {
    case MPI_INTEGER: return sizeof(int);
    case MPI_DOUBLE_PRECISION: return sizeof(double);
    case MPI_LOGICAL: return sizeof(int);
    case MPI_CXX_BOOL: return sizeof(bool);
    case MPI_CHARACTER: return 1;
  }
  throw std::runtime_error("mpi_stub: unknown datatype");
}

void barrier()
{
  if (g_size == 1) return;
  std::unique_lock<std::mutex> lk(g_mtx);
  long gen = g_bar_gen;
  if (++g_bar_count == g_size) {
    g_bar_count = 0;
    g_bar_gen++;
    g_cv.notify_all();
  } else {
    g_cv.wait(lk, [&]{ return g_bar_gen != gen; });
  }
}

template <class T> void reduce_T(void* r, int n, MPI_Op op)
{
  T* out = static_cast<T*>(r);
  for (int i = 0; i < n; i++) {
    T acc = static_cast<const T*>(g_slots[0])[i];
    for (int p = 1; p < g_size; p++) {
      T v = static_cast<const T*>(g_slots[p])[i];
      if (op == MPI_SUM) acc = acc + v;
      else if (op == MPI_MAX) acc = (v > acc) ? v : acc;
      else if (op == MPI_MIN) acc = (v < acc) ? v : acc;
    }
    out[i] = acc;
  }
}

void recv_blocking(void* buf, size_t bytes, int src, int tag)
{
  std::unique_lock<std::mutex> lk(g_mtx);
  auto key = std::make_tuple(src, t_rank, tag);
  g_cv.wait(lk, [&]{ auto it = g_mail.find(key); return it != g_mail.end() && !it->second.empty(); });
  auto& q = g_mail[key];
  Msg m = std::move(q.front());
  q.pop_front();
  if (m.data.size() > bytes) throw std::runtime_error("mpi_stub: message truncated");
  std::memcpy(buf, m.data.data(), m.data.size());
}

void send_buffered(const void* buf, size_t bytes, int dst, int tag)
{
  Msg m;
  m.data.assign(static_cast<const char*>(buf), static_cast<const char*>(buf) + bytes);
  {
    std::lock_guard<std::mutex> lk(g_mtx);
    g_mail[std::make_tuple(t_rank, dst, tag)].push_back(std::move(m));
  }
  g_cv.notify_all();
}

} // namespace

extern "C" {

void mpistub_set_world(int size)
{
  g_size = size;
  g_slots.assign(size, nullptr);
  g_slot_n.assign(size, 0);
  g_bar_count = 0;
  g_mail.clear();
}

void mpistub_bind_rank(int rank) { t_rank = rank; t_reqs.clear(); }

int MPI_Init(int*, char***) { return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Comm_rank(MPI_Comm, int* r) { *r = t_rank; return 0; }
int MPI_Comm_size(MPI_Comm, int* s) { *s = g_size; return 0; }
int MPI_Barrier(MPI_Comm) { barrier(); return 0; }

int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm)
{
  if (g_size == 1) { if (s != r) std::memcpy(r, s, n*tsize(t)); return 0; }
  g_slots[t_rank] = s;
  barrier();
  if (t == MPI_DOUBLE_PRECISION) reduce_T<double>(r, n, op);
  else if (t == MPI_INTEGER || t == MPI_LOGICAL) reduce_T<int>(r, n, op);
  else throw std::runtime_error("mpi_stub: allreduce type");
  barrier();
  return 0;
}

int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rc, const int* disp, MPI_Datatype rt, MPI_Comm)
{
  size_t sz = tsize(st);
  if (g_size == 1) { std::memcpy(static_cast<char*>(r) + disp[0]*sz, s, sn*sz); return 0; }
  g_slots[t_rank] = s;
  g_slot_n[t_rank] = sn;
  barrier();
  for (int p = 0; p < g_size; p++) {
    std::memcpy(static_cast<char*>(r) + size_t(disp[p])*sz, g_slots[p], size_t(g_slot_n[p])*sz);
  }
  barrier();
  (void)rc; (void)rt;
  return 0;
}

int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm)
{
  size_t sz = tsize(st);
  if (g_size == 1) { std::memcpy(r, s, sn*sz); return 0; }
  g_slots[t_rank] = s;
  barrier();
  for (int p = 0; p < g_size; p++) {
    std::memcpy(static_cast<char*>(r) + size_t(p)*rn*tsize(rt), g_slots[p], size_t(sn)*sz);
  }
  barrier();
  return 0;
}

int MPI_Gatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rc, const int* disp, MPI_Datatype rt, int root, MPI_Comm)
{
  size_t sz = tsize(st);
  if (g_size == 1) { std::memcpy(static_cast<char*>(r) + disp[0]*sz, s, sn*sz); return 0; }
  g_slots[t_rank] = s;
  g_slot_n[t_rank] = sn;
  barrier();
  if (t_rank == root) {
    for (int p = 0; p < g_size; p++) {
      std::memcpy(static_cast<char*>(r) + size_t(disp[p])*sz, g_slots[p], size_t(g_slot_n[p])*sz);
    }
  }
  barrier();
  (void)rc; (void)rt;
  return 0;
}

int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm)
{
  if (g_size == 1) return 0;
  g_slots[t_rank] = b;
  barrier();
  if (t_rank != root) std::memcpy(b, g_slots[root], size_t(n)*tsize(t));
  barrier();
  return 0;
}

int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm)
{
  send_buffered(b, size_t(n)*tsize(t), dst, tag);
  return 0;
}

int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm, MPI_Status*)
{
  recv_blocking(b, size_t(n)*tsize(t), src, tag);
  return 0;
}

int MPI_Isend(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm, MPI_Request* rq)
{
  send_buffered(b, size_t(n)*tsize(t), dst, tag);   // buffered: complete on return
  rq->idx = -1;
  return 0;
}

int MPI_Irecv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm, MPI_Request* rq)
{
  t_reqs.push_back({true, b, size_t(n)*tsize(t), src, tag, false});
  rq->idx = int(t_reqs.size()) - 1;
  return 0;
}

int MPI_Wait(MPI_Request* rq, MPI_Status*)
{
  if (rq->idx < 0) return 0;
  Req& q = t_reqs[rq->idx];
  if (!q.done) {
    if (q.is_recv) recv_blocking(q.buf, q.bytes, q.src, q.tag);
    q.done = true;
  }
  // drop the request table once everything in it is complete
  bool all = true;
  for (auto& r : t_reqs) all = all && r.done;
  if (all) t_reqs.clear();
  return 0;
}

int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{
  // every rank gets the result; the root is the one that reads it
  (void)root;
  return MPI_Allreduce(s, r, n, t, op, c);
}

int MPI_Gather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c)
{
  (void)root;
  return MPI_Allgather(s, sn, st, r, rn, rt, c);
}

int MPI_Scatterv(const void* s, const int* sc, const int* disp, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm)
{
  const size_t sz = tsize(st);
  if (g_size == 1) { std::memcpy(r, static_cast<const char*>(s) + size_t(disp[0])*sz, size_t(sc[0])*sz); return 0; }
  // the root publishes its buffer and layout; every rank copies its own piece
  static const int* s_sc = nullptr;
  static const int* s_disp = nullptr;
  if (t_rank == root) { g_slots[root] = s; s_sc = sc; s_disp = disp; }
  barrier();
  std::memcpy(r, static_cast<const char*>(g_slots[root]) + size_t(s_disp[t_rank])*sz, size_t(s_sc[t_rank])*sz);
  barrier();
  (void)rn; (void)rt;
  return 0;
}

// MPI-IO: only the reference's partition cache file uses it; never found, never written (every run partitions afresh)
int MPI_File_open(MPI_Comm, const char*, int, MPI_Info, MPI_File* f) { f->unused = 0; return 1; }
int MPI_File_set_view(MPI_File, MPI_Offset, MPI_Datatype, MPI_Datatype, const char*, MPI_Info) { return 0; }
int MPI_File_read(MPI_File, void*, int, MPI_Datatype, MPI_Status*) { return 1; }
int MPI_File_write(MPI_File, const void*, int, MPI_Datatype, MPI_Status*) { return 0; }
int MPI_File_close(MPI_File*) { return 0; }

} // extern "C"
