// dockauv_cull.cuh -- the cull of one env: which obstacles can a ray reach, does the vehicle touch one (cull launch of
// the pipeline layout).  Running it at the end of the dynamics launch instead (pose still in registers, obstacle rows
// prefetched into L2) was measured: 0.632 ms per step against 0.566 ms -- the cull wants its own 32 warps per SM.
#pragma once
#include "dockauv_env.cuh"

namespace dockauv {

#ifndef DOCKAUV_CULL_F32
#define DOCKAUV_CULL_F32 1        // culls + collision pre-test of the cull launch in float with conservative slack (0 = all in T)
#endif

constexpr uint32_t kViewCollision = 1u << 16;   // view_info bits: 0..15 in-view mask (capsules first), 16 collision,
constexpr uint32_t kViewListed = 1u << 17;      // 17 the env is on the ray list

// Cull of one env (thread per env; every lane of the warp must call it, `active` tells whether the lane has an env):
// walks over the env's obstacles, view word to HBM, listed envs appended to the compact list (one atomic per warp).
template <typename T>
__device__ __forceinline__ void cull_env(const KParams<T> &p, int64_t i, bool active, const T pos[3], const T Rm[9], T poison) {
    const int64_t N = p.n_envs;
    bool listed = false;
    uint32_t info = 0;
    if (active) {
#if DOCKAUV_CULL_F32
        float Rf[9];
#pragma unroll
        for (int c = 0; c < 9; c++) Rf[c] = (float)Rm[c];
#endif
        const int n_caps = p.n_caps, n_sph = p.n_sph;
        // next obstacle's words are requested before the current one is evaluated
        T ob[7], nx[7];
#pragma unroll
        for (int c = 0; c < 7; c++) nx[c] = T(0);
        const int n_obst = n_caps + n_sph;
        auto load = [&](int k, T o[7]) {
            if (k < n_caps) {
                const T *g = p.capsules + (int64_t)(k * 7) * N + i;
#pragma unroll
                for (int c = 0; c < 7; c++) o[c] = g[(int64_t)c * N];
            } else if (k < n_obst) {
                const T *g = p.spheres + (int64_t)((k - n_caps) * 4) * N + i;
#pragma unroll
                for (int c = 0; c < 4; c++) o[c] = g[(int64_t)c * N];
            }
        };
        load(0, nx);
#pragma unroll 1
        for (int k = 0; k < n_obst; k++) {
#pragma unroll
            for (int c = 0; c < 7; c++) ob[c] = nx[c];
            load(k + 1, nx);
            bool hit, view;
#if DOCKAUV_CULL_F32
            // float fast path (conservative culls, collision decided unless within 2 mm of the threshold)
            int hit3;
            cull_pair_f32<T>(p, pos, Rf, ob, k < n_caps, hit3, view);
            hit = hit3 == 1;
            if (hit3 == 2) {
                bool view64;
                obstacle_pair<T, false>(p, pos, Rm, ob, k < n_caps, nullptr, hit, view64);
            }
#else
            obstacle_pair<T, false>(p, pos, Rm, ob, k < n_caps, nullptr, hit, view);
#endif
            info |= view ? (1u << k) : 0u;
            info |= hit ? kViewCollision : 0u;
        }
        // a non-finite pose poisons the rays like the reference's NaN propagation: such envs go through the ray launch
        listed = (info & 0xffffu) != 0u || !(poison == T(0));
        if (listed) info |= kViewListed;
        p.view_info[i] = info;
    }
#ifdef DOCKAUV_VIEW_STATS   // tuning builds: in-view (env, obstacle) pairs and listed envs -> stats[11], [12]
    if (active) {
        atomicAdd(&p.stats[11], (double)__popc(info & 0xffffu));
        atomicAdd(&p.stats[12], listed ? 1.0 : 0.0);
    }
#endif
    // warp-aggregated append: one atomic per warp
    const unsigned lm = __ballot_sync(0xffffffffu, listed);
    if (lm) {
        const int lane = threadIdx.x & 31;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(p.view_count, (unsigned)__popc(lm));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (listed) {
            const unsigned k = base + __popc(lm & ((1u << lane) - 1u));
            p.view_list[k] = (unsigned long long)(uint32_t)(i - p.env_begin) | ((unsigned long long)(info & 0xffffu) << 32);
        }
    }
}

}  // namespace dockauv
