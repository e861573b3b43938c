// FP32 instantiation of the step / reset kernels (sm_100a)
#define DOCKAUV_REAL float
#include "dockauv_kernels.inl"
