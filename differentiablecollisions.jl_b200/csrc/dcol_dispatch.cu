// Persistent proximity kernels: every warp pulls 32-pair tiles from a global queue and dispatches to
// the class-pair specialisation of the tile's bin (device-linked entries from dcol_kernels_c*.cu).
#include "dcol_common.cuh"
#include <cuda_runtime.h>

#ifndef DCOL_MAXTHREADS
#define DCOL_MAXTHREADS 256
#endif
namespace dcol {

template <int C1, int C2, bool JAC>
__device__ void solve_entry(const QueryParams* p, const TileCtx* tc, int idx);   // dcol_kernels_c*.cu

template <int C1, bool JAC>
__device__ __forceinline__ void dispatch_c2(int c2, const QueryParams* p, const TileCtx* tc, int idx) {
  switch (c2) {
    case 0: solve_entry<C1, 0, JAC>(p, tc, idx); break;
    case 1: solve_entry<C1, 1, JAC>(p, tc, idx); break;
    case 2: solve_entry<C1, 2, JAC>(p, tc, idx); break;
    case 3: solve_entry<C1, 3, JAC>(p, tc, idx); break;
    case 4: solve_entry<C1, 4, JAC>(p, tc, idx); break;
    case 5: solve_entry<C1, 5, JAC>(p, tc, idx); break;
  }
}

template <bool JAC>
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
      case 0: dispatch_c2<0, JAC>(c2, &tt->P, &tc, idx); break;
      case 1: dispatch_c2<1, JAC>(c2, &tt->P, &tc, idx); break;
      case 2: dispatch_c2<2, JAC>(c2, &tt->P, &tc, idx); break;
      case 3: dispatch_c2<3, JAC>(c2, &tt->P, &tc, idx); break;
      case 4: dispatch_c2<4, JAC>(c2, &tt->P, &tc, idx); break;
      case 5: dispatch_c2<5, JAC>(c2, &tt->P, &tc, idx); break;
    }
  }
}

const void* persistent_kernel_ptr(bool jac) {
  return jac ? (const void*)proximity_persistent<true> : (const void*)proximity_persistent<false>;
}
cudaError_t launch_persistent(bool jac, int grid, int block, size_t smem, cudaStream_t st, LaunchTable* tt, int* counter) {
  if (jac) proximity_persistent<true><<<grid, block, smem, st>>>(tt, counter);
  else proximity_persistent<false><<<grid, block, smem, st>>>(tt, counter);
  return cudaGetLastError();
}
}  // namespace dcol
