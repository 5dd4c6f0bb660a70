// Persistent proximity kernel of ONE query kind (DCOL_VAR = 0 proximity, 1 gradient, 2 Jacobian): every CTA (2 per
// SM, 4 warps in lockstep each) pulls CTA-sized tiles from a global queue ordered by class pair and dispatches to the
// class-pair specialisation of the tile's bin (21 unordered class pairs).
//
// Each query kind is ONE translation unit compiled in whole-program mode (no -rdc): the kernel and its 21
// specialisations.  The specialisations stay out of line (one instruction stream per class pair, the kernel itself is
// a few hundred instructions), but because ptxas sees every caller it allocates registers across the call: no
// callee-saved register traffic around an entry and no re-materialised memory descriptors inside it.  Measured against
// the device-linked build of round 1 (-rdc, six translation units): 500-900 B of spills per entry -> 0-320 B, and
// -11 % on the 1 M mixed batch (profiles/r02_*).  DCOL_INLINE=1 inlines the specialisations into the kernel instead
// (same speed, minutes of compile time).
#pragma once
#include "dcol_solver.cuh"
#include "dcol_shared.cuh"
#include <cuda_runtime.h>

#ifndef DCOL_MAXTHREADS
#define DCOL_MAXTHREADS 256
#endif
#ifndef DCOL_MINBLOCKS
#define DCOL_MINBLOCKS 1
#endif
#ifndef DCOL_INLINE
#define DCOL_INLINE 0
#endif
namespace dcol {

template <int C1, int C2, int VAR>
__device__ __forceinline__ void run_entry(const QueryParams* p, const TileCtx* tc, int idx) {
#if DCOL_INLINE
  solve_pair<C1, C2, VAR>(*p, *tc, idx);
#else
  solve_entry<C1, C2, VAR>(p, tc, idx);
#endif
}

// only class pairs C1 <= C2 exist; the bins are keyed by the unordered class pair
template <int C1, int VAR>
__device__ __forceinline__ void dispatch_c2(int c2, const QueryParams* p, const TileCtx* tc, int idx) {
  switch (c2) {
    case 0: if constexpr (C1 <= 0) run_entry<C1, 0, VAR>(p, tc, idx); break;
    case 1: if constexpr (C1 <= 1) run_entry<C1, 1, VAR>(p, tc, idx); break;
    case 2: if constexpr (C1 <= 2) run_entry<C1, 2, VAR>(p, tc, idx); break;
    case 3: if constexpr (C1 <= 3) run_entry<C1, 3, VAR>(p, tc, idx); break;
    case 4: if constexpr (C1 <= 4) run_entry<C1, 4, VAR>(p, tc, idx); break;
    case 5: if constexpr (C1 <= 5) run_entry<C1, 5, VAR>(p, tc, idx); break;
  }
}

// DCOL_MAXNREG (tuning builds): an exact register budget instead of the one derived from the launch bounds
#ifdef DCOL_MAXNREG
#define DCOL_KERNEL_BOUNDS __maxnreg__(DCOL_MAXNREG)
#else
#define DCOL_KERNEL_BOUNDS __launch_bounds__(DCOL_MAXTHREADS, DCOL_MINBLOCKS)
#endif
template <int VAR>
__global__ void DCOL_KERNEL_BOUNDS proximity_persistent(LaunchTable* __restrict__ tt, int* __restrict__ counter) {
  __shared__ int s_tile;
  // Small batches launch this kernel with programmatic stream serialisation behind the preparation kernel: its blocks are
  // already resident when that kernel ends and wait here for its tables (a no-op for ordinary launches).
  asm volatile("griddepcontrol.wait;" ::: "memory");
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
    // lanes past the end of the bin load its last pair, so every thread of the CTA runs the same specialisation and
    // reaches the same barriers; they are masked out of the iteration vote, the hand-over and the result stores
    const int last = __ldg(&tt->end[b]) - 1;
    const int want = __ldg(&tt->begin[b]) + (tile - __ldg(&tt->tile_prefix[b])) * tile_sz + (int)threadIdx.x;
    const int idx = min(want, last);
    const int cp = __ldg(&tt->cp[b]);
    TileCtx tc;
    tc.mode = tt->mode; tc.spill_t = (tt->ov_cap[b] > 0) ? tt->spill_t : 0;
    tc.ov = tt->ov_base + tt->ov_off[b]; tc.ov_cap = tt->ov_cap[b]; tc.ov_count = &tt->ov_count[b];
    tc.in = tt->in_base + tt->in_off[b]; tc.in_cap = tt->in_cap[b];
    tc.dup = want > last;
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

// A handful of pairs whose packed records the host wrote itself (host-pointer calls of up to DCOL_PACKED_MAX pairs, above all the
// single-pair calls; dcol_capi.cu): no binning pass and no launch table, one launch in all.  One block per class pair
// present; lanes past the group's last pair redo it for the barriers only.
template <int VAR>
__global__ void __launch_bounds__(DCOL_PACKED_MAX) proximity_packed(QueryParams P, PackedGroups G) {
  const int g = blockIdx.x, cnt = G.count[g];
  TileCtx tc;
  tc.mode = 0; tc.spill_t = 0; tc.ov = nullptr; tc.ov_cap = 0; tc.ov_count = nullptr; tc.in = nullptr; tc.in_cap = 0;
  tc.dup = (int)threadIdx.x >= cnt;
  const int idx = G.first[g] + min((int)threadIdx.x, cnt - 1);
  const int cp = G.cp[g];
  const int c1 = cp / NCLS, c2 = cp - c1 * NCLS;
  switch (c1) {
    case 0: dispatch_c2<0, VAR>(c2, &P, &tc, idx); break;
    case 1: dispatch_c2<1, VAR>(c2, &P, &tc, idx); break;
    case 2: dispatch_c2<2, VAR>(c2, &P, &tc, idx); break;
    case 3: dispatch_c2<3, VAR>(c2, &P, &tc, idx); break;
    case 4: dispatch_c2<4, VAR>(c2, &P, &tc, idx); break;
    case 5: dispatch_c2<5, VAR>(c2, &P, &tc, idx); break;
  }
}

// host-side handles of this translation unit's kernel (dcol_capi.cu)
#define DCOL_CAT2(a, b) a##b
#define DCOL_CAT(a, b) DCOL_CAT2(a, b)
const void* DCOL_CAT(persistent_kernel_ptr_v, DCOL_VAR)() { return (const void*)proximity_persistent<DCOL_VAR>; }
cudaError_t DCOL_CAT(launch_persistent_v, DCOL_VAR)(int grid, int block, size_t smem, cudaStream_t st, LaunchTable* tt, int* counter, int early) {
  if (early) {      // may start before the previous kernel of the stream has finished (it waits at griddepcontrol.wait)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, proximity_persistent<DCOL_VAR>, tt, counter);
  }
  proximity_persistent<DCOL_VAR><<<grid, block, smem, st>>>(tt, counter);
  return cudaGetLastError();
}
cudaError_t DCOL_CAT(launch_packed_v, DCOL_VAR)(int block, size_t smem, cudaStream_t st, const QueryParams& P, const PackedGroups& G) {
  if (smem > 48 * 1024) {      // (row stores of large polytopes only)
    cudaError_t e = cudaFuncSetAttribute(proximity_packed<DCOL_VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  proximity_packed<DCOL_VAR><<<G.n, block, smem, st>>>(P, G);
  return cudaGetLastError();
}
}  // namespace dcol
