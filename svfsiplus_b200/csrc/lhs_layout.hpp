// lhs_layout.hpp — node renumbering and overlap lists of a partitioned linear system, host side (no device needed).
// Replaces the part of fsils_lhs_create that is not a plain copy (Code/Source/liner_solver/lhs.cpp:57-376): given every
// rank's global node list in local order (what the MPI_Allgatherv at lhs.cpp:156 collects), rank r's solver ordering is
//   [ nodes shared with LOWER ranks | interior nodes | nodes shared with HIGHER ranks ]
// built by walking the other ranks from the highest id down (lhs.cpp:171-224): a node of r that rank i also holds and that
// no earlier rank claimed goes to the front block (i < r, in rank i's local order) or to the back block (i > r, written from
// the END backwards, lhs.cpp:198-199); interior nodes keep their relative order.  mynNo = nNo - |back block| (rows this rank
// counts in dot products), shnNo = |front block|.  The overlap list of a pair is ordered as the HIGHER rank of the pair walks
// its renumbered node list (lhs.cpp:304-360).  Integer for integer what the reference leaves in lhs.map / lhs.cS[].ptr
// (tests/test_partition.py, against the compiled reference on threads-as-ranks).
//
// Cost: O(gnNo + sum of all node lists) per renumbered list, one list for the rank itself and one per higher rank it shares
// nodes with; two gnNo-sized int tables at a time (54 MB each at 13.5 M global nodes).
#pragma once

#include <cstdint>
#include <stdexcept>
#include <utility>
#include <vector>

namespace svb200 {

struct LhsLayout {
  int nNo = 0, mynNo = 0, shnNo = 0;
  std::vector<int> map;                                    // local assembly id -> solver id
  std::vector<std::pair<int, std::vector<int>>> reqs;      // (peer rank, solver ids), ascending peer
};

namespace detail {

// renumbered global node list of rank r (lhs.cpp:171-224); gtl is scratch of size gnNo
inline void renumbered_list(int r, int nT, int gnNo, const int* counts, const int* const* gnodes, std::vector<int>& gtl,
                            std::vector<int>& ltg, int& shnNo, int& mynNo)
{
  const int n = counts[r];
  const int* g = gnodes[r];
  gtl.assign(size_t(gnNo), -1);
  for (int a = 0; a < n; a++) {
    if (g[a] < 0 || g[a] >= gnNo) throw std::runtime_error("lhs_layout: global node id outside [0, gnNo)");
    if (gtl[g[a]] != -1) throw std::runtime_error("lhs_layout: a rank lists a global node twice");
    gtl[g[a]] = a;
  }
  std::vector<char> taken(size_t(n), 0);
  std::vector<int> low, high;
  for (int i = nT - 1; i >= 0; i--) {
    if (i == r) continue;
    std::vector<int>& dst = (i < r) ? low : high;
    const int* gi = gnodes[i];
    for (int b = 0; b < counts[i]; b++) {
      const int v = gi[b];
      if (v < 0 || v >= gnNo) throw std::runtime_error("lhs_layout: global node id outside [0, gnNo)");
      const int a = gtl[v];
      if (a >= 0 && !taken[a]) { taken[a] = 1; dst.push_back(v); }
    }
  }
  ltg.clear();
  ltg.reserve(size_t(n));
  ltg.insert(ltg.end(), low.begin(), low.end());
  for (int a = 0; a < n; a++) if (!taken[a]) ltg.push_back(g[a]);
  ltg.insert(ltg.end(), high.rbegin(), high.rend());
  shnNo = int(low.size());
  mynNo = n - int(high.size());
}

} // namespace detail

inline LhsLayout lhs_layout(int rank, int nT, int gnNo, const int* counts, const int* const* gnodes)
{
  if (nT < 1 || rank < 0 || rank >= nT || gnNo < 0) throw std::runtime_error("lhs_layout: bad rank / size arguments");
  LhsLayout L;
  const int n = counts[rank];
  L.nNo = n;
  L.map.resize(size_t(n));
  if (nT == 1) {                                           // lhs.cpp:96-121
    for (int a = 0; a < n; a++) L.map[a] = a;
    L.mynNo = n;
    return L;
  }
  std::vector<int> gtl, ltg;
  detail::renumbered_list(rank, nT, gnNo, counts, gnodes, gtl, ltg, L.shnNo, L.mynNo);
  std::vector<int> pos(size_t(gnNo), -1);                  // global id -> solver id on this rank
  for (int k = 0; k < n; k++) pos[ltg[k]] = k;
  for (int a = 0; a < n; a++) L.map[a] = pos[gnodes[rank][a]];
  std::vector<int> ltg_hi;
  for (int i = 0; i < nT; i++) {
    if (i == rank) continue;
    bool common = false;
    for (int b = 0; b < counts[i] && !common; b++) common = pos[gnodes[i][b]] >= 0;
    if (!common) continue;                                 // lhs.cpp:300-302
    std::vector<int> ptr;
    if (i < rank) {
      // this rank is the higher one of the pair: its own renumbered order, restricted to the peer's nodes
      gtl.assign(size_t(gnNo), -1);
      for (int b = 0; b < counts[i]; b++) gtl[gnodes[i][b]] = b;
      for (int k = 0; k < n; k++) if (gtl[ltg[k]] >= 0) ptr.push_back(k);
    } else {
      int sh, my;
      detail::renumbered_list(i, nT, gnNo, counts, gnodes, gtl, ltg_hi, sh, my);
      for (int v : ltg_hi) if (pos[v] >= 0) ptr.push_back(pos[v]);
    }
    L.reqs.emplace_back(i, std::move(ptr));
  }
  return L;
}

} // namespace svb200

// ---------------------------------------------------------------------------------------------------------------------------
// Element partition by recursive coordinate bisection of the element centroids.  The reference partitions the dual graph
// with ParMETIS (Code/Source/solver/distribute.cpp:1683-1700, SPLIT.c) - a third-party library whose output depends on the
// rank count and only changes summation order downstream; this is the stand-in for meshes that are not generated slab by
// slab: balanced to one element, deterministic (ties broken by element id), O(n log P), no graph needed.  Parts are numbered
// along the cuts, so neighbouring ranks are neighbouring parts for elongated domains (a pipe is cut across its axis).
// ---------------------------------------------------------------------------------------------------------------------------
#include <algorithm>
#include <numeric>

namespace svb200 {

namespace detail {

inline void rcb(const double* c, int* idx, int n, int p0, int np, int* part)
{
  if (np == 1 || n == 0) {
    for (int i = 0; i < n; i++) part[idx[i]] = p0;
    return;
  }
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) {
      const double v = c[size_t(idx[i])*3 + d];
      lo[d] = std::min(lo[d], v); hi[d] = std::max(hi[d], v);
    }
  int ax = 0;
  for (int d = 1; d < 3; d++) if (hi[d] - lo[d] > hi[ax] - lo[ax]) ax = d;
  const int npL = np/2;
  const int nL = int((long long)n*npL/np);
  auto less = [c, ax](int a, int b) {
    const double va = c[size_t(a)*3 + ax], vb = c[size_t(b)*3 + ax];
    return va < vb || (va == vb && a < b);
  };
  std::nth_element(idx, idx + nL, idx + n, less);
  rcb(c, idx, nL, p0, npL, part);
  rcb(c, idx + nL, n - nL, p0 + npL, np - npL, part);
}

} // namespace detail

inline void partition_rcb(int nEl, const double* centroids, int nParts, int* part)
{
  if (nEl < 0 || nParts < 1 || (nEl && (!centroids || !part))) throw std::runtime_error("partition_rcb: bad arguments");
  std::vector<int> idx(size_t(nEl), 0);
  std::iota(idx.begin(), idx.end(), 0);
  detail::rcb(centroids, idx.data(), nEl, 0, nParts, part);
}


// ---- host-side tables of the device transport and of the row-tile kernels (pure functions: tests/hostlogic drives them on the CPU) ----

// Node-centric source lists of the overlap add (fsils_commuv adds the received values request by request, in_commu.cpp:150-168): for
// every distinct overlap row, its (request, position) sources in REQUEST order.  lists[i] = solver rows of request i.
struct HaloSources { std::vector<int> node, ptr; std::vector<int> src_req, src_pos; };
inline HaloSources halo_source_lists(int nNo, const std::vector<std::vector<int>>& lists)
{
  HaloSources h;
  std::vector<int> cnt(size_t(nNo) + 1, 0);
  size_t total = 0;
  for (const auto& l : lists) for (int v : l) { if (v < 0 || v >= nNo) throw std::runtime_error("overlap list entry out of range"); cnt[v + 1]++; total++; }
  std::vector<int> slot(nNo, -1);
  h.ptr.push_back(0);
  for (int r = 0; r < nNo; r++) if (cnt[r + 1]) { slot[r] = int(h.node.size()); h.node.push_back(r); h.ptr.push_back(h.ptr.back() + cnt[r + 1]); }
  h.src_req.assign(total, 0); h.src_pos.assign(total, 0);
  std::vector<int> fill(h.node.size(), 0);
  for (int i = 0; i < int(lists.size()); i++)
    for (int j = 0; j < int(lists[i].size()); j++) {
      const int k = slot[lists[i][j]];
      const int e = h.ptr[k] + fill[k]++;
      h.src_req[e] = i; h.src_pos[e] = j;
    }
  return h;
}

// Row tiles of the TMA-staged SpMV kernels: consecutive rows, at most max_rows rows and cap - 4 entries counted from the tile's first
// entry rounded down to a multiple of 4 (the bulk copies are 16-byte aligned), never across cuts[1] / cuts[2].  Returns the tile start
// rows + the end row, tile_at[4] = first tile of each of the three segments and the tile count; empty when a single row exceeds a tile.
inline std::vector<int> row_tiles(const std::vector<int>& rowPtr, const int cuts[4], int max_rows, int cap, int tile_at[4])
{
  std::vector<int> tr;
  for (int sgm = 0; sgm < 3; sgm++) {
    tile_at[sgm] = int(tr.size());
    int r = cuts[sgm];
    while (r < cuts[sgm+1]) {
      tr.push_back(r);
      const int p0 = rowPtr[r] & ~3;
      int q = r;
      while (q < cuts[sgm+1] && q - r < max_rows && rowPtr[q+1] - p0 <= cap - 4) q++;
      if (q == r) return std::vector<int>();
      r = q;
    }
  }
  tile_at[3] = int(tr.size());
  tr.push_back(cuts[3]);
  return tr;
}

} // namespace svb200
