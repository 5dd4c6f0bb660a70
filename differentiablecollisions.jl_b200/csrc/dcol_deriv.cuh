// placeholder until the derivative kernel lands
#pragma once
#include "dcol_common.cuh"
namespace dcol {
template <int C1, int C2>
__global__ void deriv_kernel(QueryParams p, int bin_begin, int bin_count) {}
}
