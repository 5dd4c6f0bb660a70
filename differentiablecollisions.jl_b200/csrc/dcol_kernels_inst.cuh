// Explicit instantiation of the proximity kernels for one first-primitive class (DCOL_C1):
// 6 second classes x {solver(+gradient epilogue), solver+Jacobian epilogue}.  Split over six
// translation units so nvcc compiles them in parallel.
#include "dcol_solver.cuh"

namespace dcol {
typedef void (*SolveFn)(QueryParams, int, int);
#define DCOL_CAT2(a, b) a##b
#define DCOL_CAT(a, b) DCOL_CAT2(a, b)
SolveFn DCOL_CAT(solve_row_, DCOL_C1)(int c2, bool jac) {
  switch (c2) {
    case 0: return jac ? solve_kernel<DCOL_C1, 0, true> : solve_kernel<DCOL_C1, 0, false>;
    case 1: return jac ? solve_kernel<DCOL_C1, 1, true> : solve_kernel<DCOL_C1, 1, false>;
    case 2: return jac ? solve_kernel<DCOL_C1, 2, true> : solve_kernel<DCOL_C1, 2, false>;
    case 3: return jac ? solve_kernel<DCOL_C1, 3, true> : solve_kernel<DCOL_C1, 3, false>;
    case 4: return jac ? solve_kernel<DCOL_C1, 4, true> : solve_kernel<DCOL_C1, 4, false>;
    case 5: return jac ? solve_kernel<DCOL_C1, 5, true> : solve_kernel<DCOL_C1, 5, false>;
  }
  return nullptr;
}
}  // namespace dcol
