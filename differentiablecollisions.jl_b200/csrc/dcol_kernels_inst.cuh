// Explicit instantiation of the per-pair solver entries: the 21 class pairs (C1 <= C2) x {proximity, gradient, Jacobian}
// variants, dealt round-robin over six translation units (DCOL_TU = 0..5; compiled in parallel, relocatable device code)
// and device-linked.  Pairs arriving as (C2, C1) are solved by the same entry with their primitives exchanged.
#include "dcol_solver.cuh"

namespace dcol {
#define DCOL_INST(C1, C2) \
  template __device__ void solve_entry<C1, C2, 0>(const QueryParams*, const TileCtx*, int); \
  template __device__ void solve_entry<C1, C2, 1>(const QueryParams*, const TileCtx*, int); \
  template __device__ void solve_entry<C1, C2, 2>(const QueryParams*, const TileCtx*, int);
#if DCOL_TU == 0
DCOL_INST(0, 0) DCOL_INST(1, 1) DCOL_INST(2, 3) DCOL_INST(4, 4)
#elif DCOL_TU == 1
DCOL_INST(0, 1) DCOL_INST(1, 2) DCOL_INST(2, 4) DCOL_INST(4, 5)
#elif DCOL_TU == 2
DCOL_INST(0, 2) DCOL_INST(1, 3) DCOL_INST(2, 5) DCOL_INST(5, 5)
#elif DCOL_TU == 3
DCOL_INST(0, 3) DCOL_INST(1, 4) DCOL_INST(3, 3)
#elif DCOL_TU == 4
DCOL_INST(0, 4) DCOL_INST(1, 5) DCOL_INST(3, 4)
#elif DCOL_TU == 5
DCOL_INST(0, 5) DCOL_INST(2, 2) DCOL_INST(3, 5)
#else
#error "DCOL_TU must be 0..5"
#endif
}  // namespace dcol
