#define DCOL_TU 4
#include "dcol_kernels_inst.cuh"
