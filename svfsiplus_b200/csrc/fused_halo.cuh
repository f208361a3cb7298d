// fused_halo.cuh — ONE kernel per partitioned block product: the rows of the SpMV, the overlap-node exchange over peer-mapped
// memory (peer_comm.cuh) and the ordered add of the received contributions, i.e. fsils_spar_mul_* + fsils_commuv
// (liner_solver/spar_mul.cpp, in_commu.cpp:111-170) without a second launch.
//
//   phase A   boundary rows [0, ovA) and [ovB, nNo) (the rows that appear in an overlap list), spread over the whole grid
//   barrier   grid-wide arrive/wait on a device counter (the grid never exceeds what is resident: occupancy x SMs)
//   push      every CTA packs its slice of the overlap lists straight into the neighbours' windows (NVLink stores); the last CTA
//             to finish release-stores the epoch flags
//   phase B   interior rows [ovA, ovB) - the exchange is in flight meanwhile
//   add       every CTA acquires the neighbours' flags and adds its slice of the overlap rows in request order
//
// Against the four launches of the unfused path (boundary rows, k_halo_push, interior rows, k_halo_wait_add) this removes three
// kernel boundaries with their drain / fill gaps and lets the push start the moment the last boundary row is done.
// The row bodies are the same arithmetic as the stand-alone kernels of kernels.cuh (k_spmv_vv3c, k_spmv_vv4, k_schur_gp,
// k_schur_sp4), restated as device functions on global row indices.
#pragma once

#include "kernels.cuh"
#include "peer_comm.cuh"

namespace svb200 {

// ---- row bodies: rows [r0, r1), quad `gq` of `nq` quads in the grid ---------------------------------------------------------------
struct RowsVV3 {            // dof-3 block rows, column-owner lanes (k_spmv_vv3c)
  const int* rowPtr; const int* col; const double* K; const double* U; double* KU;
  __device__ __forceinline__ void run(int r0, int r1, int gq, int nq, int lane4) const
  {
    const int li = lane4 < 3 ? lane4 : 0;
    const uint64_t pol_u = l2_policy_evict_last();
    const int n = r1 - r0, nrounds = (n + nq - 1)/nq;
    for (int r = 0; r < nrounds; r++) {
      const int row = r0 + gq + r*nq;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      if (row < r1) {
        const int s = __ldg(rowPtr + row), e = __ldg(rowPtr + row + 1);
#pragma unroll 4
        for (int p = s; p < e; p++) {
          const int c = __ldg(col + p);
          const double* k = K + size_t(p)*9 + li;
          const double u = ld_hint(U + size_t(c)*3 + li, pol_u);
          a0 = fma(__ldg(k), u, a0);
          a1 = fma(__ldg(k + 3), u, a1);
          a2 = fma(__ldg(k + 6), u, a2);
        }
      }
      if (lane4 == 3) { a0 = 0.0; a1 = 0.0; a2 = 0.0; }
      a0 += __shfl_xor_sync(0xffffffffu, a0, 1); a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
      a0 += __shfl_xor_sync(0xffffffffu, a0, 2); a1 += __shfl_xor_sync(0xffffffffu, a1, 2); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
      if (row < r1 && lane4 < 3) KU[size_t(row)*3 + lane4] = (lane4 == 0) ? a0 : (lane4 == 1) ? a1 : a2;
    }
  }
};

struct RowsVV4 {            // dof-4 block rows (k_spmv_vv4)
  const int* rowPtr; const int* col; const double* K; const double* U; double* KU;
  __device__ __forceinline__ void run(int r0, int r1, int gq, int nq, int lane4) const
  {
    for (int row = r0 + gq; row < r1; row += nq) {
      const int s = __ldg(rowPtr + row), e = __ldg(rowPtr + row + 1);
      double acc = 0.0;
      int p = s;
      for (; p + 2 <= e; p += 2) {
        const int c0 = ld_stream_i(col + p), c1 = ld_stream_i(col + p + 1);
        const d4 k0 = ld256_stream(K + (size_t(p)*16 + lane4*4));
        const d4 k1 = ld256_stream(K + (size_t(p+1)*16 + lane4*4));
        const d4 u0 = ld256_keep(U + size_t(c0)*4);
        const d4 u1 = ld256_keep(U + size_t(c1)*4);
        acc = acc + k0.x*u0.x + k0.y*u0.y + k0.z*u0.z + k0.w*u0.w;
        acc = acc + k1.x*u1.x + k1.y*u1.y + k1.z*u1.z + k1.w*u1.w;
      }
      for (; p < e; p++) {
        const int c0 = ld_stream_i(col + p);
        const d4 k0 = ld256_stream(K + (size_t(p)*16 + lane4*4));
        const d4 u0 = ld256_keep(U + size_t(c0)*4);
        acc = acc + k0.x*u0.x + k0.y*u0.y + k0.z*u0.z + k0.w*u0.w;
      }
      KU[size_t(row)*4 + lane4] = acc;
    }
  }
};

struct RowsGP {             // pass 1 of the Schur operator (k_schur_gp / k_schur_gp_soa): V4(i) = [sum_j G(:,j) P(col_j), P(i)]
  const int* rowPtr; const int* col; const double* G; const double* P; double* V4;
  size_t soa = 0;             // 0: G(3,nnz) entry-wise; else G is the component-wise copy with this component stride (= nnz)
  __device__ __forceinline__ void run(int r0, int r1, int gq, int nq, int lane4) const
  {
    const int n = r1 - r0, nrounds = (n + nq - 1)/nq;
    for (int r = 0; r < nrounds; r++) {
      const int row = r0 + gq + r*nq;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      if (row < r1) {
        const int s = __ldg(rowPtr + row), e = __ldg(rowPtr + row + 1);
#pragma unroll 2
        for (int p = s + lane4; p < e; p += 4) {
          const double u = __ldg(P + __ldg(col + p));
          const double* g = soa ? G + p : G + size_t(p)*3;
          const size_t st = soa ? soa : 1;
          a0 = fma(__ldg(g), u, a0);
          a1 = fma(__ldg(g + st), u, a1);
          a2 = fma(__ldg(g + 2*st), u, a2);
        }
      }
      a0 += __shfl_xor_sync(0xffffffffu, a0, 1); a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
      a0 += __shfl_xor_sync(0xffffffffu, a0, 2); a1 += __shfl_xor_sync(0xffffffffu, a1, 2); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
      if (row < r1 && lane4 == 0) {
        d4 o; o.x = a0; o.y = a1; o.z = a2; o.w = __ldg(P + row);
        st256(V4 + size_t(row)*4, o);
      }
    }
  }
};

struct RowsSP {             // pass 2 of the Schur operator (k_schur_sp): SP(i) = sum_j L(j) V4(3,col_j) - sum_j Gt(:,j).V4(0:2,col_j)
  const int* rowPtr; const int* col; const double* GtL; const double* V4; double* SP;
  __device__ __forceinline__ void run(int r0, int r1, int gq, int nq, int lane4) const
  {
    const int n = r1 - r0, nrounds = (n + nq - 1)/nq;
    for (int r = 0; r < nrounds; r++) {
      const int row = r0 + gq + r*nq;
      double aL = 0.0, aD = 0.0;
      if (row < r1) {
        const int s = __ldg(rowPtr + row), e = __ldg(rowPtr + row + 1);
        // four entries in flight per lane (k_schur_sp4); V4 was completed by an earlier launch, cached loads are fine
        for (int base = s + lane4; base < e; base += 16) {
          int c[4];
          d4 k[4], v[4];
#pragma unroll
          for (int q = 0; q < 4; q++) { const int p = base + 4*q; c[q] = (p < e) ? ld_stream_i(col + p) : -1; }
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const int p = base + 4*q;
            if (c[q] >= 0) { k[q] = ld256_stream(GtL + size_t(p)*4); v[q] = ld256_keep(V4 + size_t(c[q])*4); }
            else { k[q].x = k[q].y = k[q].z = k[q].w = 0.0; v[q] = k[q]; }
          }
#pragma unroll
          for (int q = 0; q < 4; q++) {
            aL = fma(k[q].w, v[q].w, aL);
            aD = aD + (k[q].x*v[q].x + k[q].y*v[q].y + k[q].z*v[q].z);
          }
        }
      }
      aL += __shfl_xor_sync(0xffffffffu, aL, 1); aD += __shfl_xor_sync(0xffffffffu, aD, 1);
      aL += __shfl_xor_sync(0xffffffffu, aL, 2); aD += __shfl_xor_sync(0xffffffffu, aD, 2);
      if (row < r1 && lane4 == 0) SP[row] = aL - aD;
    }
  }
};

struct FusedHaloArgs {
  const int* skip;
  int nNo, ovA, ovB;
  int dof, ld, dofcap;                 // exchanged components per overlap row, leading dimension of `out`
  double* out;                         // the product's result vector (rows with leading dimension ld)
  int nreq; const PeerHaloReq* reqs; const int* ptr_all; int halo_tot;
  int nh; const int* hn_node; const int* hn_ptr; const int2* hn_src;
  PeerState* ps;
};

template <class Rows>
__global__ void __launch_bounds__(256) k_rows_halo(Rows rows, FusedHaloArgs f)
{
  const bool skip = f.skip && *f.skip;                      // compute may be skipped, the exchange never is (peer_comm.cuh)
  const int lane4 = threadIdx.x & 3;
  const int gq = (blockIdx.x*blockDim.x + threadIdx.x) >> 2;
  const int nq = (gridDim.x*blockDim.x) >> 2;
  const unsigned long long epoch = f.ps->halo_epoch + 1;    // advanced by the last CTA of this launch
  const int par = int(epoch & 1ull);

  // phase A: boundary rows
  if (!skip) {
    if (f.ovA > 0) rows.run(0, f.ovA, gq, nq, lane4);
    if (f.ovB < f.nNo) rows.run(f.ovB, f.nNo, gq, nq, lane4);
  }
  // grid-wide barrier: every CTA of this launch is resident (the grid is sized by the occupancy calculator and the stream runs one
  // kernel at a time), arrivals are counted on a device counter that the last CTA of the launch resets
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&f.ps->bar_count, 1u);
    const long long t0 = clock64();
    while (ld_acquire_gpu(&f.ps->bar_count) < gridDim.x) {
      if (clock64() - t0 > kSpinLimit) { f.ps->error = 3; break; }       // not every CTA resident: reported by peer_check, no hang
      __nanosleep(32);
    }
  }
  __syncthreads();

  // push: this CTA's slice of the overlap lists into the neighbours' windows
  {
    const int total = f.halo_tot*f.dof;
    for (int t = blockIdx.x*blockDim.x + threadIdx.x; t < total; t += gridDim.x*blockDim.x) {
      const int j = t / f.dof, l = t - j*f.dof;
      int r = 0;
      while (r + 1 < f.nreq && j >= f.reqs[r+1].off) r++;
      const PeerHaloReq q = f.reqs[r];
      q.rdata[size_t(par)*q.n*f.dofcap + size_t(j - q.off)*f.dof + l] = __ldcg(f.out + size_t(f.ptr_all[j])*f.ld + l);
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) s_last = (atomicAdd(&f.ps->push_count, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {
      __threadfence_system();
      if (threadIdx.x < f.nreq) st_release_sys(f.reqs[threadIdx.x].rflag + par*kPeerMaxReq, epoch);
      if (threadIdx.x == 0) f.ps->push_count = 0;
    }
  }

  // phase B: interior rows
  if (!skip && f.ovB > f.ovA) rows.run(f.ovA, f.ovB, gq, nq, lane4);

  // add: wait for the neighbours' values, then this CTA's slice of the overlap rows, sources in request order
  {
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (threadIdx.x < f.nreq) {
      if (!spin_until(f.reqs[threadIdx.x].lflag + par*kPeerMaxReq, epoch, f.ps)) s_ok = 0;
    }
    __syncthreads();
    const int total = f.nh*f.dof;
    for (int t = blockIdx.x*blockDim.x + threadIdx.x; t < total; t += gridDim.x*blockDim.x) {
      const int k = t / f.dof, l = t - k*f.dof;
      double* dst = f.out + size_t(f.hn_node[k])*f.ld + l;
      double s = __ldcg(dst);
      for (int e = f.hn_ptr[k]; e < f.hn_ptr[k+1]; e++) {
        const int2 src = f.hn_src[e];
        const PeerHaloReq& q = f.reqs[src.x];
        s += ld_peer_written(q.ldata + size_t(par)*q.n*f.dofcap + size_t(src.y)*f.dof + l);
      }
      *dst = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (!s_ok) f.ps->error = 1;
      __threadfence();
      if (atomicAdd(&f.ps->wait_count, 1u) == gridDim.x - 1) { f.ps->wait_count = 0; f.ps->bar_count = 0; f.ps->halo_epoch = epoch; }
    }
  }
}

} // namespace svb200
