#define DCOL_TU 1
#include "dcol_kernels_inst.cuh"
