#define DCOL_TU 3
#include "dcol_kernels_inst.cuh"
