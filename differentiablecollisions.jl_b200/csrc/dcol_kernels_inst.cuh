// Explicit instantiation of the per-pair solver entries for one first-primitive class (DCOL_C1):
// 6 second classes x {proximity, gradient, Jacobian} variants.  Split over six
// translation units (compiled in parallel, relocatable device code) and device-linked.
#include "dcol_solver.cuh"

namespace dcol {
#define DCOL_INST(C2) \
  template __device__ void solve_entry<DCOL_C1, C2, 0>(const QueryParams*, const TileCtx*, int); \
  template __device__ void solve_entry<DCOL_C1, C2, 1>(const QueryParams*, const TileCtx*, int); \
  template __device__ void solve_entry<DCOL_C1, C2, 2>(const QueryParams*, const TileCtx*, int);
DCOL_INST(0) DCOL_INST(1) DCOL_INST(2) DCOL_INST(3) DCOL_INST(4) DCOL_INST(5)
}  // namespace dcol
