// Persistent proximity kernels (one per query kind): every CTA (2 per SM, 4 warps in lockstep each) pulls CTA-sized
// tiles from a global queue ordered by class pair and dispatches to the class-pair specialisation of the tile's
// bin (device-linked entries from dcol_kernels_c*.cu; 21 unordered class pairs per query kind).
#include "dcol_common.cuh"
#include <cuda_runtime.h>

#ifndef DCOL_MAXTHREADS
#define DCOL_MAXTHREADS 256
#endif
namespace dcol {

template <int C1, int C2, int VAR>
__device__ void solve_entry(const QueryParams* p, const TileCtx* tc, int idx);   // dcol_kernels_c*.cu

// only class pairs C1 <= C2 exist (dcol_kernels_inst.cuh); the bins are keyed by the unordered class pair
template <int C1, int VAR>
__device__ __forceinline__ void dispatch_c2(int c2, const QueryParams* p, const TileCtx* tc, int idx) {
  switch (c2) {
    case 0: if constexpr (C1 <= 0) solve_entry<C1, 0, VAR>(p, tc, idx); break;
    case 1: if constexpr (C1 <= 1) solve_entry<C1, 1, VAR>(p, tc, idx); break;
    case 2: if constexpr (C1 <= 2) solve_entry<C1, 2, VAR>(p, tc, idx); break;
    case 3: if constexpr (C1 <= 3) solve_entry<C1, 3, VAR>(p, tc, idx); break;
    case 4: if constexpr (C1 <= 4) solve_entry<C1, 4, VAR>(p, tc, idx); break;
    case 5: if constexpr (C1 <= 5) solve_entry<C1, 5, VAR>(p, tc, idx); break;
  }
}

template <int VAR>
__global__ void __launch_bounds__(DCOL_MAXTHREADS, 1) proximity_persistent(LaunchTable* __restrict__ tt, int* __restrict__ counter) {
  __shared__ int s_tile;
  const int n_tiles = tt->n_tiles;
  const int tile_sz = blockDim.x;
  for (;;) {
    // one tile per atomic: the whole GPU walks the queue together, so at any moment (nearly) all CTAs run
    // the same class-pair specialisation and share it in the GPC-level instruction cache
    __syncthreads();                       // previous tile done: row store and s_tile can be reused
    if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1);
    __syncthreads();
    const int tile = s_tile;
    if (tile >= n_tiles) break;
    int b = 0;
    while (tile >= __ldg(&tt->tile_prefix[b + 1])) b++;
    // lanes past the end of the bin redo its last pair (identical stores), so every thread of the CTA
    // runs the same specialisation and reaches the same barriers
    const int last = __ldg(&tt->end[b]) - 1;
    const int idx = min(__ldg(&tt->begin[b]) + (tile - __ldg(&tt->tile_prefix[b])) * tile_sz + (int)threadIdx.x, last);
    const int cp = __ldg(&tt->cp[b]);
    TileCtx tc;
    tc.mode = tt->mode; tc.spill_t = (tt->ov_cap[b] > 0) ? tt->spill_t : 0;
    tc.ov = tt->ov_base + tt->ov_off[b]; tc.ov_cap = tt->ov_cap[b]; tc.ov_count = &tt->ov_count[b];
    tc.in = tt->in_base + tt->in_off[b]; tc.in_cap = tt->in_cap[b];
    const int c1 = cp / NCLS, c2 = cp - c1 * NCLS;
    switch (c1) {
      case 0: dispatch_c2<0, VAR>(c2, &tt->P, &tc, idx); break;
      case 1: dispatch_c2<1, VAR>(c2, &tt->P, &tc, idx); break;
      case 2: dispatch_c2<2, VAR>(c2, &tt->P, &tc, idx); break;
      case 3: dispatch_c2<3, VAR>(c2, &tt->P, &tc, idx); break;
      case 4: dispatch_c2<4, VAR>(c2, &tt->P, &tc, idx); break;
      case 5: dispatch_c2<5, VAR>(c2, &tt->P, &tc, idx); break;
    }
  }
}

const void* persistent_kernel_ptr(int var) {
  return var == 2 ? (const void*)proximity_persistent<2> : (var == 1 ? (const void*)proximity_persistent<1> : (const void*)proximity_persistent<0>);
}
cudaError_t launch_persistent(int var, int grid, int block, size_t smem, cudaStream_t st, LaunchTable* tt, int* counter) {
  if (var == 2) proximity_persistent<2><<<grid, block, smem, st>>>(tt, counter);
  else if (var == 1) proximity_persistent<1><<<grid, block, smem, st>>>(tt, counter);
  else proximity_persistent<0><<<grid, block, smem, st>>>(tt, counter);
  return cudaGetLastError();
}
}  // namespace dcol
