#define DCOL_C1 0
#include "dcol_kernels_inst.cuh"
