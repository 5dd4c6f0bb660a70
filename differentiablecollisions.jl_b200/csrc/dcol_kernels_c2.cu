#define DCOL_TU 2
#include "dcol_kernels_inst.cuh"
