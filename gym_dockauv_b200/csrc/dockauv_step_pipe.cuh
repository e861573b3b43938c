// dockauv_step_pipe.cuh -- layout DOCKAUV_LAYOUT_PIPELINE: one batched step as four specialised launches.
//
//   1. dynamics   step_warp_kernel<MODE 1> (dockauv_step_warp.cuh): thread per env, writes the post-step pose and the
//                 radar-independent reward terms to the hand-off buffer.
//   2. cull       cull_kernel: thread per env.  Walks over the env's obstacles (coalesced SoA loads), does the body
//                 collision test and the range / field-of-view culls -- in float with conservative slack, the collision
//                 re-decided in T when it is within 2 mm of the threshold (cull_pair_f32) -- and appends the envs that
//                 have anything in view (28 % of them on the C4 workload) to a compact list.
//   3. rays       rays_kernel: persistent grid, one warp per LISTED env (lanes = rays), data of the next list entry
//                 prefetched while the current one is cast.  Writes the pooled ray cells of the observation row and the
//                 obstacle-avoidance sum.
//   4. finish     finish_kernel: thread per env: reward, done, counters, statistics, observation cells of envs with an
//                 empty view; the ~1 % of envs whose episode ended are compacted per CTA and re-initialised by
//                 neighbouring threads instead of one lane per warp.
//
// Each launch has its own register budget and occupancy (the fused / split kernels carry the radar's ~90 registers
// through everything and run at 4 warps per scheduler), the ray warps never idle on envs with nothing in view, and the
// three thread-per-env launches are plain data-parallel code.  Results are identical to the other layouts (same
// device functions, same order of operations per env); debug outputs are served by the fused kernel.
#pragma once
#include "dockauv_step_warp.cuh"

namespace dockauv {

#ifndef DOCKAUV_MINB_CULL
#define DOCKAUV_MINB_CULL 4
#endif
#ifndef DOCKAUV_RAY_CTAS_PER_SM
#define DOCKAUV_RAY_CTAS_PER_SM 4   // persistent grid of the ray launch, in 4-warp CTAs per SM (= what is resident; 8: +0.7 %, 16: +2 % time)
#endif
#ifndef DOCKAUV_MINB_RAYS
#define DOCKAUV_MINB_RAYS 4
#endif

// ------------------------------------------------------------------------------------------------------- 2. cull
template <typename T>
__global__ void __launch_bounds__(256, DOCKAUV_MINB_CULL) cull_kernel(const __grid_constant__ KParams<T> p) {
    const int64_t N = p.n_envs;
    const int64_t i = p.env_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < p.env_end;
    T pos[3] = {T(0), T(0), T(0)}, Rm[9] = {T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)}, poison = T(0);
    if (active) {
        const T *hf = p.handoff + i;
#pragma unroll
        for (int c = 0; c < 3; c++) pos[c] = hf[(int64_t)c * N];
#pragma unroll
        for (int c = 0; c < 9; c++) Rm[c] = hf[(int64_t)(3 + c) * N];
        poison = hf[(int64_t)12 * N];
    }
    cull_env<T>(p, i, active, pos, Rm, poison);
}

// ------------------------------------------------------------------------------------------------------- 3. rays
template <typename T>
struct RaysSmem {
    int warp_words;      // T words per warp: pose[14] + rec[16][kPreStride] + rays[ray_stride]
    int ray_stride;
    __host__ __device__ RaysSmem(int n_rays) {
        ray_stride = (n_rays + 2) & ~1;
        warp_words = kPoseStride + 16 * kPreStride + ray_stride;
    }
};

#ifndef DOCKAUV_RAY_WARPS
#define DOCKAUV_RAY_WARPS 4
#endif
constexpr int kRayWarps = DOCKAUV_RAY_WARPS;     // warps per CTA of the ray launch

template <typename T, int RPL>
__global__ void __launch_bounds__(kRayWarps * 32, DOCKAUV_MINB_RAYS * 4 / kRayWarps) rays_kernel(const __grid_constant__ KParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using P2 = typename Pair<T>::type;
    const RaysSmem<T> L(p.n_rays);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T *s_pose = reinterpret_cast<T *>(smem_raw) + warp * L.warp_words;
    T *s_pre = s_pose + kPoseStride;
    T *s_ray = s_pre + 16 * kPreStride;
    const int64_t N = p.n_envs;
    const unsigned count = *p.view_count;
    const unsigned n_warps = gridDim.x * kRayWarps;
    unsigned idx = blockIdx.x * kRayWarps + warp;
    if (idx >= count) return;

    const int n_caps = p.n_caps, n_sph = p.n_sph, n_obst = n_caps + n_sph, n_r = p.n_rays;
    const T dmax = p.radar_max_dist, inv_dmax = T(1) / dmax;
    // this lane's rays: body-frame direction and obstacle-avoidance weight stay in registers
    T rb[RPL][3], bw[RPL];
#pragma unroll
    for (int j = 0; j < RPL; j++) {
        const int ir = lane + 32 * j;
        const bool ok = ir < n_r;
#pragma unroll
        for (int c = 0; c < 3; c++) rb[j][c] = ok ? p.ray_tab[c * n_r + ir] : T(0);
        bw[j] = ok ? p.ray_tab[3 * n_r + ir] : T(0);
    }
    // 2x2 pooling (at most one pooled cell per lane): the four source slots of this lane's cell; cells beyond the ray
    // grid read the zero slot s_ray[n_r] (block_reduce pads with cval = 0, sensor.py:131-137)
    const bool fast_pool = (p.block == 2) && (p.n_rr <= 32);
    int pidx[4] = {n_r, n_r, n_r, n_r};
    if (fast_pool && lane < p.n_rr) {
        const int pr = lane / p.n_hr, pcol = lane - pr * p.n_hr;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int rv = 2 * pr + (q >> 1), rh = 2 * pcol + (q & 1);
            if (rv < p.n_vert && rh < p.n_horiz) pidx[q] = rv * p.n_horiz + rh;
        }
    }
    if (lane == 0) s_ray[n_r] = T(0);

    // software pipeline over the list: entry n + 2 and the data of entry n + 1 are in flight while entry n is cast.
    // data of one entry: lane c < 13 holds pose word c, lane k < n_obst with bit k of the mask holds obstacle k.
    const bool lane_is_cap = lane < n_caps;
    const T *obst_row = lane_is_cap ? p.capsules + (int64_t)(lane * 7) * N : p.spheres + (int64_t)((lane - n_caps) * 4) * N;
    auto fetch = [&](uint64_t entry, T &pose_w, T ob[7]) {
        const int64_t e = p.env_begin + (int64_t)(uint32_t)entry;
        const unsigned mask = (unsigned)(entry >> 32);
        if (lane < 13) pose_w = p.handoff[(int64_t)lane * N + e];
        if (lane < n_obst && ((mask >> lane) & 1u)) {
            const T *g = obst_row + e;
            const int n_words = lane_is_cap ? 7 : 4;
#pragma unroll
            for (int c = 0; c < 7; c++)
                if (c < n_words) ob[c] = g[(int64_t)c * N];
        }
    };
    uint64_t cur = p.view_list[idx];
    uint64_t nxt = (idx + n_warps < count) ? p.view_list[idx + n_warps] : 0;
    T pose_w = T(0), ob[7];
#pragma unroll
    for (int c = 0; c < 7; c++) ob[c] = T(0);
    fetch(cur, pose_w, ob);

    for (; idx < count; idx += n_warps) {
        const int64_t ie = p.env_begin + (int64_t)(uint32_t)cur;
        const unsigned mask = (unsigned)(cur >> 32);
        // ---- stage the prefetched data, start the next fetches
        if (lane < 13) s_pose[lane] = pose_w;
        __syncwarp();
        if (lane < n_obst && ((mask >> lane) & 1u)) {
            const T pos[3] = {s_pose[0], s_pose[1], s_pose[2]};
            bool hit, view;
            obstacle_pair<T, true>(p, pos, s_pose + 3, ob, lane_is_cap, s_pre + lane * kPreStride, hit, view);
        }
        const uint64_t nn = (idx + 2 * n_warps < count) ? p.view_list[idx + 2 * n_warps] : 0;
        if (idx + n_warps < count) fetch(nxt, pose_w, ob);
        __syncwarp();

        // ---- cast rays against the in-view obstacles (uniform loop, broadcast shared reads)
        T best[RPL];
#pragma unroll
        for (int j = 0; j < RPL; j++) best[j] = Mth<T>::inf();
        if (mask) {
            T rd[RPL][3];
            {
                T R[9];
#pragma unroll
                for (int c = 0; c < 9; c++) R[c] = s_pose[3 + c];
#pragma unroll
                for (int j = 0; j < RPL; j++) {
#pragma unroll
                    for (int c = 0; c < 3; c++)
                        rd[j][c] = R[3 * c] * rb[j][0] + R[3 * c + 1] * rb[j][1] + R[3 * c + 2] * rb[j][2];
                }
            }
            unsigned cap_mask = mask & ((1u << n_caps) - 1u);
            unsigned sph_mask = (mask >> n_caps) & ((1u << n_sph) - 1u);
            while (cap_mask) {
                const int k = __ffs(cap_mask) - 1;
                cap_mask &= cap_mask - 1;
                const P2 *w2 = reinterpret_cast<const P2 *>(s_pre + k * kPreStride);
                const P2 v0 = w2[0], v1 = w2[1], v2 = w2[2], v3 = w2[3], v4 = w2[4], v5 = w2[5];
                const T ba[3] = {v0.x, v0.y, v1.x}, oa[3] = {v1.y, v2.x, v2.y};
                const T baba = v3.x, baoa = v3.y, cc = v4.x, c2a = v4.y, c2b = v5.x;
#pragma unroll
                for (int j = 0; j < RPL; j++) {
                    // shape.py:341-390 for one ray: cylinder root, body hit if 0 < y < baba, else end cap
                    const T bard = rd[j][0] * ba[0] + rd[j][1] * ba[1] + rd[j][2] * ba[2];
                    const T rdoa = rd[j][0] * oa[0] + rd[j][1] * oa[1] + rd[j][2] * oa[2];
                    const T a = baba - bard * bard;
                    const T b = baba * rdoa - baoa * bard;
                    const T h = b * b - a * cc;
                    if (h > T(0)) {
                        const T t = (-b - Mth<T>::sqrt_pos(h)) * Mth<T>::rcp_(a);
                        const T y = baoa + t * bard;
                        T v = t;
                        if (!(y > T(0) && y < baba)) {
                            const bool far_end = y >= T(0);
                            const T b2 = far_end ? rdoa - bard : rdoa;     // rd . (pos - cap end)
                            const T h2 = b2 * b2 - (far_end ? c2b : c2a);
                            v = (h2 > T(0)) ? (-b2 - Mth<T>::sqrt_pos(h2 > T(0) ? h2 : T(1))) : T(-1);
                        }
                        if (v > T(0) && v < best[j]) best[j] = v;
                    }
                }
            }
            while (sph_mask) {
                const int k = __ffs(sph_mask) - 1;
                sph_mask &= sph_mask - 1;
                const P2 *w2 = reinterpret_cast<const P2 *>(s_pre + (n_caps + k) * kPreStride);
                const P2 v0 = w2[0], v1 = w2[1];
#pragma unroll
                for (int j = 0; j < RPL; j++) {
                    // shape.py:252-263: nearest root of the ray / sphere quadratic
                    const T b = v0.x * rd[j][0] + v0.y * rd[j][1] + v1.x * rd[j][2];
                    const T h = b * b - v1.y;
                    if (h >= T(0)) {
                        const T v = -b - (h > T(0) ? Mth<T>::sqrt_pos(h) : T(0));
                        if (v > T(0) && v < best[j]) best[j] = v;
                    }
                }
            }
        }
        // ---- clamp (sensor.py:117), obstacle-avoidance sum (docking3d.py:767-792), stash for pooling
        const T poison = s_pose[12];
        T oa_part = T(0);
#pragma unroll
        for (int j = 0; j < RPL; j++) {
            const int ir = lane + 32 * j;
            if (ir < n_r) {
                // min positive distance over obstacles (docking3d.py:438-439), max_dist if none or farther
                const T d = (best[j] > dmax ? dmax : best[j]) + poison;
                s_ray[ir] = d;
                // (gamma_c (1 - c))^2 with c = clip(1 - d/d_max, 0, 1): 1 - c = d/d_max for d in [0, d_max]
                const T x = d * inv_dmax;
                const T qq = x * x;
                const T mx = !(qq <= T(0.001)) ? qq : T(0.001);     // np.maximum, NaN propagates
                oa_part += mx * bw[j];
            }
        }
        const T oa_dot = warp_sum<T>(oa_part);
        if (lane == 0) p.oa_dot[ie] = oa_dot;
        __syncwarp();
        // ---- 2x2 max-pool with zero padding (sensor.py:131-137) -> obs[16:]
        float *orow = p.obs + ie * p.n_obs + 16;
        if (fast_pool) {
            if (lane < p.n_rr) {
                T mx = s_ray[pidx[0]];
#pragma unroll
                for (int q = 1; q < 4; q++) {
                    const T v = s_ray[pidx[q]];
                    mx = !(v <= mx) ? v : mx;            // np.max, NaN propagates
                }
                T o = mx * inv_dmax;                     // clip(d / max_dist, 0, 1), docking3d.py:487
                o = o > T(1) ? T(1) : o;
                orow[lane] = (float)o;
            }
        } else {
            for (int pc = lane; pc < p.n_rr; pc += 32) {
                const int pr = pc / p.n_hr, pcol = pc - pr * p.n_hr;
                T mx = T(0);
                for (int dv = 0; dv < p.block; dv++)
                    for (int dh = 0; dh < p.block; dh++) {
                        const int rv = pr * p.block + dv, rh = pcol * p.block + dh;
                        if (rv < p.n_vert && rh < p.n_horiz) {
                            const T v = s_ray[rv * p.n_horiz + rh];
                            mx = !(v <= mx) ? v : mx;
                        }
                    }
                T o = mx * inv_dmax;
                o = o > T(1) ? T(1) : o;
                orow[pc] = (float)o;
            }
        }
        __syncwarp();
        cur = nxt;
        nxt = nn;
    }
}

// ------------------------------------------------------------------------------------------------------- 4. finish
#ifndef DOCKAUV_MINB_FINISH
#define DOCKAUV_MINB_FINISH 4      // 64 registers; 5 (48 registers, 40 B spills) measured the same
#endif
#ifndef DOCKAUV_FINISH_THREADS
#define DOCKAUV_FINISH_THREADS 256
#endif
constexpr int kFinishThreads = DOCKAUV_FINISH_THREADS;

template <typename T>
__global__ void __launch_bounds__(kFinishThreads, DOCKAUV_MINB_FINISH) finish_kernel(const __grid_constant__ KParams<T> p) {
    __shared__ int s_n_reset;
    __shared__ int s_reset[kFinishThreads];
    const int64_t N = p.n_envs;
    const int64_t i0 = p.env_begin + (int64_t)blockIdx.x * kFinishThreads;
    const int64_t i = i0 + threadIdx.x;
    const bool active = i < p.env_end;
    if (threadIdx.x == 0) s_n_reset = 0;
    __syncthreads();
    WarpStats bs;
    if (active) {
        const T *hf = p.handoff + i;
        StepCarry<T> cy;
#pragma unroll
        for (int c = 0; c < 8; c++) cy.rarr[c] = hf[(int64_t)(13 + c) * N];
        cy.delta_d = hf[(int64_t)21 * N];
        cy.cond = p.handoff_cond[i];
        cy.t_steps = p.t_steps[i];
        cy.ep_return = p.ep_return[i];
        const uint32_t info = p.view_info[i];
        const bool listed = (info & kViewListed) != 0u;
        // envs with an empty view: every ray reads max_dist (sensor.py:113-117) -> pooled cells all ones, r_oa = 0
        T oa = p.sum_beta_oa;
        float *row = p.obs + i * p.n_obs;
        if (listed) {
            oa = p.oa_dot[i];
        } else {
            float *cells = row + 16;
            if ((p.n_obs & 3) == 0 && (p.n_rr & 3) == 0) {
                float4 *c4 = reinterpret_cast<float4 *>(cells);
                for (int c = 0; c < (p.n_rr >> 2); c++) c4[c] = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
            } else {
                for (int c = 0; c < p.n_rr; c++) cells[c] = 1.0f;
            }
        }
        const T r_oa = p.sum_beta_oa / oa - T(1);      // docking3d.py:792
        const bool done = step_finish<T, false, true>(p, i, cy, r_oa, (info & kViewCollision) != 0u, bs);
        if (done) s_reset[atomicAdd(&s_n_reset, 1)] = threadIdx.x;     // ~1 % of the envs per step
    }
    bs.flush(p.stats, threadIdx.x == 0 ? (int)min((int64_t)kFinishThreads, p.env_end - i0) : 0);
    // ---- finished envs of this CTA, compacted:
    __syncthreads();
    const int n_done = s_n_reset;
    // one warp per finished env (lanes = columns / reset roles; a single thread walking its own row is a chain of 32
    // dependent HBM round trips, a single-thread reset a ~3000-instruction chain, and the whole CTA would wait):
    //  (a) the last observation is kept as terminal_observation and the all-zero reset observation is handed back
    //      (docking3d.py:269,322);  (b) the env is re-initialised (reset_env_warp).
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_obs = p.n_obs;
        for (int e = warp; e < n_done; e += kFinishThreads / 32) {
            const int64_t ie = i0 + s_reset[e];
            float *row = p.obs + ie * n_obs;
            float *trow = p.terminal_obs ? p.terminal_obs + ie * n_obs : nullptr;
            for (int c = lane; c < n_obs; c += 32) {
                if (trow) trow[c] = row[c];
                if (p.auto_reset) row[c] = 0.0f;
            }
            if (p.auto_reset) reset_env_warp<T>(p, ie, lane);
        }
    }
}

// ------------------------------------------------------------------------------------------------------- launcher
template <typename T, int VEH, int NU>
static cudaError_t launch_step_pipe(const KParams<T> &k, cudaStream_t st, cudaEvent_t *marks = nullptr, int *n_marks = nullptr) {
    if (wants_debug(k) || k.handoff == nullptr) return launch_step_warp<T, VEH, NU>(k, st);   // debug outputs: fused kernel
    if (k.split_chunk > 0 && k.split_chunk < k.env_end - k.env_begin && marks == nullptr) {
        // chunked: the four launches per chunk of envs, so that a chunk's hand-off can still be in L2 when it is read
        const int64_t chunk = ((k.split_chunk + kWarpEnvs - 1) / kWarpEnvs) * kWarpEnvs;
        for (int64_t b = k.env_begin; b < k.env_end; b += chunk) {
            KParams<T> kb = k;
            kb.env_begin = b;
            kb.env_end = b + chunk < k.env_end ? b + chunk : k.env_end;
            kb.split_chunk = 0;
            cudaError_t e = launch_step_pipe<T, VEH, NU>(kb, st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    const int64_t n = k.env_end - k.env_begin;
    KParams<T> kc = k;
    kc.view_count = k.view_count + (k.env_begin / kWarpEnvs);   // one list counter per concurrently stepped env range
    kc.view_list = k.view_list + k.env_begin;
    int n_mark = 0;
    auto mark = [&]() {
        if (marks) cudaEventRecord(marks[n_mark++], st);
    };
    mark();
    // scenarios without obstacles: no cull, no rays (the view words stay at their initial 0 = nothing in view, no collision)
    const bool has_obstacles = k.n_caps + k.n_sph > 0;
    cudaError_t e = launch_step_warp_rpl<T, VEH, NU, 2, 1, false>(kc, st);
    if (e != cudaSuccess) return e;
    mark();
    if (has_obstacles) cull_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(kc);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    mark();
    if (has_obstacles) {
        const RaysSmem<T> L(k.n_rays);
        const int smem = kRayWarps * L.warp_words * (int)sizeof(T);
        int64_t blocks = (int64_t)(k.sm_count > 0 ? k.sm_count : 148) * (DOCKAUV_RAY_CTAS_PER_SM * 4 / kRayWarps);
        const int64_t most = (n + kRayWarps - 1) / kRayWarps;
        if (blocks > most) blocks = most;
        if (k.n_rays <= 64) {
            auto kern = rays_kernel<T, 2>;
            if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
            kern<<<(unsigned)blocks, kRayWarps * 32, smem, st>>>(kc);
        } else {
            auto kern = rays_kernel<T, 8>;
            if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
            kern<<<(unsigned)blocks, kRayWarps * 32, smem, st>>>(kc);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    mark();
    finish_kernel<T><<<(unsigned)((n + kFinishThreads - 1) / kFinishThreads), kFinishThreads, 0, st>>>(kc);
    mark();
    if (n_marks) *n_marks = n_mark;
    return cudaGetLastError();
}

}  // namespace dockauv
