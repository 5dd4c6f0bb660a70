#define DCOL_TU 0
#include "dcol_kernels_inst.cuh"
