// dockauv_kernels.inl -- included by dockauv_kernels_f64.cu / _f32.cu with DOCKAUV_REAL defined.
#include "dockauv_launch.h"
#include "dockauv_step_tpe.cuh"
#include "dockauv_step_warp.cuh"
#include "dockauv_step_pipe.cuh"

namespace dockauv {

template <typename T, int VEH, int NU>
static cudaError_t launch_variant(const KParams<T> &k, int layout, cudaStream_t st, cudaEvent_t *marks, int *n_marks) {
    const int64_t n = k.env_end - k.env_begin;
    if (n_marks) *n_marks = 0;
    if (n <= 0) return cudaSuccess;
    if (layout == DOCKAUV_LAYOUT_WARP_RAYS) return launch_step_warp<T, VEH, NU>(k, st);
    if (layout == DOCKAUV_LAYOUT_PIPELINE) return launch_step_pipe<T, VEH, NU>(k, st, marks, n_marks);
    const int threads = 128;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    step_tpe_kernel<T, VEH, NU><<<blocks, threads, 0, st>>>(k);
    return cudaGetLastError();
}

template <>
cudaError_t launch_step<DOCKAUV_REAL>(const KParams<DOCKAUV_REAL> &k, int vehicle, int layout, cudaStream_t st,
                                      cudaEvent_t *marks, int *n_marks) {
    if (vehicle == DOCKAUV_VEHICLE_LAUV) return launch_variant<DOCKAUV_REAL, DOCKAUV_VEHICLE_LAUV, 3>(k, layout, st, marks, n_marks);
    if (k.n_u == 8) return launch_variant<DOCKAUV_REAL, DOCKAUV_VEHICLE_BLUEROV2, 8>(k, layout, st, marks, n_marks);
    return launch_variant<DOCKAUV_REAL, DOCKAUV_VEHICLE_BLUEROV2, 6>(k, layout, st, marks, n_marks);
}

template <>
cudaError_t launch_refresh_obstacles<DOCKAUV_REAL>(const KParams<DOCKAUV_REAL> &k, cudaStream_t st) {
    if (k.obsf == nullptr || k.n_envs <= 0) return cudaSuccess;
    const int threads = 256;
    refresh_obstacles_kernel<DOCKAUV_REAL><<<(unsigned)((k.n_envs + threads - 1) / threads), threads, 0, st>>>(k);
    return cudaGetLastError();
}

template <>
cudaError_t launch_reset<DOCKAUV_REAL>(const KParams<DOCKAUV_REAL> &k, const uint8_t *mask_dev, cudaStream_t st) {
    const int64_t n = k.env_end - k.env_begin;
    if (n <= 0) return cudaSuccess;
    reset_kernel<DOCKAUV_REAL><<<(unsigned)((n + 31) / 32), kResetCta, 0, st>>>(k, mask_dev);
    return cudaGetLastError();
}

}  // namespace dockauv
