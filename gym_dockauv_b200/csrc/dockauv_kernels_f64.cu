// FP64 instantiation of the step / reset kernels (sm_100a)
#define DOCKAUV_REAL double
#include "dockauv_kernels.inl"
