#define DCOL_C1 3
#include "dcol_kernels_inst.cuh"
