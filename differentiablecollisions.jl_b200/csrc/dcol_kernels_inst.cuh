// Explicit instantiation of the solver / derivative kernels for one first-primitive
// class (DCOL_C1).  Split over six translation units so nvcc compiles them in parallel.
#include "dcol_solver.cuh"
#include "dcol_deriv.cuh"

namespace dcol {
typedef void (*SolveFn)(QueryParams, int, int);
#define DCOL_CAT2(a, b) a##b
#define DCOL_CAT(a, b) DCOL_CAT2(a, b)
SolveFn DCOL_CAT(solve_row_, DCOL_C1)(int c2) {
  switch (c2) {
    case 0: return solve_kernel<DCOL_C1, 0>;
    case 1: return solve_kernel<DCOL_C1, 1>;
    case 2: return solve_kernel<DCOL_C1, 2>;
    case 3: return solve_kernel<DCOL_C1, 3>;
    case 4: return solve_kernel<DCOL_C1, 4>;
    case 5: return solve_kernel<DCOL_C1, 5>;
  }
  return nullptr;
}
SolveFn DCOL_CAT(deriv_row_, DCOL_C1)(int c2) {
  switch (c2) {
    case 0: return deriv_kernel<DCOL_C1, 0>;
    case 1: return deriv_kernel<DCOL_C1, 1>;
    case 2: return deriv_kernel<DCOL_C1, 2>;
    case 3: return deriv_kernel<DCOL_C1, 3>;
    case 4: return deriv_kernel<DCOL_C1, 4>;
    case 5: return deriv_kernel<DCOL_C1, 5>;
  }
  return nullptr;
}
}  // namespace dcol
