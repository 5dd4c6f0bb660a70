// Kernel lookup by (class1, class2, variant).
#include "dcol_common.cuh"
namespace dcol {
typedef void (*SolveFn)(QueryParams, int, int);
SolveFn solve_row_0(int, bool); SolveFn solve_row_1(int, bool); SolveFn solve_row_2(int, bool);
SolveFn solve_row_3(int, bool); SolveFn solve_row_4(int, bool); SolveFn solve_row_5(int, bool);
SolveFn get_solve_kernel(int c1, int c2, bool jac) {
  switch (c1) {
    case 0: return solve_row_0(c2, jac); case 1: return solve_row_1(c2, jac); case 2: return solve_row_2(c2, jac);
    case 3: return solve_row_3(c2, jac); case 4: return solve_row_4(c2, jac); case 5: return solve_row_5(c2, jac);
  }
  return nullptr;
}
}  // namespace dcol
