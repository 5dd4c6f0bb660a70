// query kind 2 of the persistent proximity kernel (see dcol_persistent.cuh)
#define DCOL_VAR 2
#include "dcol_persistent.cuh"
