#define DCOL_C1 2
#include "dcol_kernels_inst.cuh"
