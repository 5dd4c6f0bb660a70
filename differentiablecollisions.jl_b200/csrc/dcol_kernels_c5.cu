#define DCOL_TU 5
#include "dcol_kernels_inst.cuh"
