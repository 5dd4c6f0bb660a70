// query kind 1 of the persistent proximity kernel (see dcol_persistent.cuh)
#define DCOL_VAR 1
#include "dcol_persistent.cuh"
