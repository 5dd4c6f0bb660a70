#define DCOL_C1 1
#include "dcol_kernels_inst.cuh"
