#define DCOL_C1 4
#include "dcol_kernels_inst.cuh"
