// Kernel lookup by (class1, class2).
#include "dcol_common.cuh"
namespace dcol {
typedef void (*SolveFn)(QueryParams, int, int);
SolveFn solve_row_0(int); SolveFn solve_row_1(int); SolveFn solve_row_2(int);
SolveFn solve_row_3(int); SolveFn solve_row_4(int); SolveFn solve_row_5(int);
SolveFn deriv_row_0(int); SolveFn deriv_row_1(int); SolveFn deriv_row_2(int);
SolveFn deriv_row_3(int); SolveFn deriv_row_4(int); SolveFn deriv_row_5(int);
SolveFn get_solve_kernel(int c1, int c2) {
  switch (c1) {
    case 0: return solve_row_0(c2); case 1: return solve_row_1(c2); case 2: return solve_row_2(c2);
    case 3: return solve_row_3(c2); case 4: return solve_row_4(c2); case 5: return solve_row_5(c2);
  }
  return nullptr;
}
SolveFn get_deriv_kernel(int c1, int c2) {
  switch (c1) {
    case 0: return deriv_row_0(c2); case 1: return deriv_row_1(c2); case 2: return deriv_row_2(c2);
    case 3: return deriv_row_3(c2); case 4: return deriv_row_4(c2); case 5: return deriv_row_5(c2);
  }
  return nullptr;
}
}  // namespace dcol
