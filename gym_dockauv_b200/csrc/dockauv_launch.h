// dockauv_launch.h -- host-side launch entry points; one translation unit per precision instantiates them
// (dockauv_kernels_f64.cu, dockauv_kernels_f32.cu) so the two compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include "dockauv_kparams.h"

namespace dockauv {

// marks (nullable): kMaxStepLaunches + 1 events; when given, the multi-launch layouts record one before their first
// launch and one after each launch, *n_marks receives how many were recorded (0 for single-launch layouts)
constexpr int kMaxStepLaunches = 6;
template <typename T>
cudaError_t launch_step(const KParams<T> &k, int vehicle, int layout, cudaStream_t st, cudaEvent_t *marks = nullptr,
                        int *n_marks = nullptr);

template <typename T>
cudaError_t launch_reset(const KParams<T> &k, const uint8_t *mask_dev, cudaStream_t st);

// float obstacle records of the cull launch from the bound obstacle / goal buffers (no-op for handles without them)
template <typename T>
cudaError_t launch_refresh_obstacles(const KParams<T> &k, cudaStream_t st);

}  // namespace dockauv
