// gmpe_flat.cuh -- the consumer-side edge list in the layout the consumer takes.
//
// GNNBase.forward (onpolicy/algorithms/utils/gnn_new.py:492-500) runs process_adj (:328-358) on
// the (B*N, E, E) stack of adjacency matrices -- one graph per (world, agent), the N graphs of a
// world being the same matrix -- and gets ONE flat COO list over all graphs:
// edge_index (2, nnz) int64 with node ids graph * E + row / col, in (graph, row, col) order, and
// edge_attr (nnz, 1).  The step kernel already emits the per-world list (gnn_edge_index /
// gnn_edge_attr / gnn_num_edges: row-major, strict <, fp32 cast); what is left is an exclusive
// scan of the per-world counts and a gather that replicates each world's list N times with the
// graph offsets added.  Three small launches, no host synchronisation, caller-owned buffers.
#pragma once
#include "gmpe_device.cuh"

namespace gmpe {

constexpr int kFlatBlock = 1024;

// exclusive prefix of n[] within blocks of kFlatBlock worlds; block totals to block_tot[]
__global__ void __launch_bounds__(kFlatBlock) flat_scan_blocks(const int32_t* __restrict__ n, int64_t B,
                                                               int64_t* __restrict__ local, int64_t* __restrict__ block_tot) {
  __shared__ int32_t wsum[kFlatBlock / 32];
  const int64_t b = (int64_t)blockIdx.x * kFlatBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t v = b < B ? n[b] : 0;
  int32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int32_t s = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    wsum[lane] = s;  // inclusive prefix of the warp sums
  }
  __syncthreads();
  const int32_t before = warp ? wsum[warp - 1] : 0;
  if (b < B) local[b] = (int64_t)(before + x - v);
  if (threadIdx.x == kFlatBlock - 1) block_tot[blockIdx.x] = (int64_t)(before + x);
}

// exclusive scan of the block totals in place (one CTA); total[0] = N * (sum of all counts)
__global__ void __launch_bounds__(kFlatBlock) flat_scan_tops(int64_t* __restrict__ block_tot, int nblk, int N, int64_t* __restrict__ total) {
  __shared__ int64_t wsum[kFlatBlock / 32];
  __shared__ int64_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += kFlatBlock) {
    const int q = base + threadIdx.x;
    const int64_t v = q < nblk ? block_tot[q] : 0;
    int64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int64_t s = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      wsum[lane] = s;
    }
    __syncthreads();
    const int64_t carry = carry_s;
    const int64_t before = carry + (warp ? wsum[warp - 1] : 0);
    if (q < nblk) block_tot[q] = before + x - v;
    __syncthreads();
    if (threadIdx.x == kFlatBlock - 1) carry_s = before + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] = carry_s * (int64_t)N;
}

// one warp per world: its list, N times, with the graph offsets (graph = world * N + agent)
template <typename Real>
__global__ void __launch_bounds__(256) flat_expand(const int32_t* __restrict__ gidx, const Real* __restrict__ gattr,
                                                   const int32_t* __restrict__ n, const int64_t* __restrict__ local,
                                                   const int64_t* __restrict__ block_pre, int64_t B, int N, int E, int cap,
                                                   int64_t* __restrict__ out_row, int64_t* __restrict__ out_col,
                                                   Real* __restrict__ out_attr, int64_t capacity) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < B; b += warps) {
    const int nb = n[b];
    const int64_t base = (block_pre[b / kFlatBlock] + local[b]) * (int64_t)N;
    if (base + (int64_t)nb * N > capacity) continue;  // the caller sees total > capacity and raises
    const int32_t* r = gidx + (size_t)b * 2 * cap;
    const int32_t* c = r + cap;
    const Real* a = gattr + (size_t)b * cap;
    for (int e = lane; e < nb; e += 32) {
      const int64_t re = r[e], ce = c[e];
      const Real ae = a[e];
      for (int k = 0; k < N; ++k) {
        const int64_t node0 = (b * N + k) * (int64_t)E;
        const int64_t o = base + (int64_t)k * nb + e;
        out_row[o] = node0 + re;
        out_col[o] = node0 + ce;
        out_attr[o] = ae;
      }
    }
  }
}

}  // namespace gmpe
