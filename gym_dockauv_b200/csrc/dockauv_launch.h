// dockauv_launch.h -- host-side launch entry points; one translation unit per precision instantiates them
// (dockauv_kernels_f64.cu, dockauv_kernels_f32.cu) so the two compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include "dockauv_kparams.h"

namespace dockauv {

template <typename T>
cudaError_t launch_step(const KParams<T> &k, int vehicle, int layout, cudaStream_t st);

template <typename T>
cudaError_t launch_reset(const KParams<T> &k, const uint8_t *mask_dev, cudaStream_t st);

}  // namespace dockauv
