// query kind 0 of the persistent proximity kernel (see dcol_persistent.cuh)
#define DCOL_VAR 0
#include "dcol_persistent.cuh"
