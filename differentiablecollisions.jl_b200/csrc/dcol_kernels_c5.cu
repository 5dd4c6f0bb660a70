#define DCOL_C1 5
#include "dcol_kernels_inst.cuh"
