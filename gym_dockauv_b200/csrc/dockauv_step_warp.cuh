// dockauv_step_warp.cuh -- layout DOCKAUV_LAYOUT_WARP_RAYS.
//
// One CTA of 128 threads owns 128 consecutive envs.
//   phase A (thread per env) : current, command filter, RKF45, angle wrap, nav errors, obs[0:16], done bits 0..3 and
//                              the radar-independent reward terms; the post-step pose goes to shared memory.
//   phase B (warp per env)   : each warp walks over the 32 envs its own threads just integrated.  Lanes 0..K-1
//                              load one obstacle each, do the ray-independent algebra and the body-collision
//                              test, and park the result in a per-warp shared scratch; then every lane casts
//                              its rays (ray = lane + 32 j) against the obstacles that are within radar range
//                              (uniform loop, broadcast shared reads), the per-ray minimum is lane-local, the
//                              obstacle-avoidance sum is a shuffle reduction and the 2x2 max-pool goes through
//                              the per-warp scratch.
//   phase C (thread per env) : reward, done, counters, statistics, auto-reset; then the CTA streams its
//                              128 x n_obs float32 observation tile to HBM with fully coalesced stores.
#pragma once
#include "dockauv_step_tpe.cuh"

namespace dockauv {

constexpr int kWarpEnvs = 128;       // envs (= threads) per CTA
constexpr int kPreCap = 12;          // shared words per capsule: ba[3] oa[3] baba baoa c c2a c2b (+1 pad)
constexpr int kPreSph = 4;           // shared words per sphere: oc[3] c

template <typename T>
struct WarpSmem {
    // layout of the dynamic shared memory block (all offsets in bytes, computed on host and device alike)
    int pose_off, oa_off, obs_off, pre_off, ray_off, flag_off, total;
    int obs_stride;     // floats per staged observation row (odd -> conflict-free column writes)
    int pre_stride;     // T words per warp of obstacle scratch
    int ray_stride;     // T words per warp of ray-distance scratch
    __host__ __device__ WarpSmem(int n_obs, int n_rays) {
        int off = 0;
        pose_off = off; off += kWarpEnvs * 12 * (int)sizeof(T);
        oa_off = off;   off += kWarpEnvs * (int)sizeof(T);
        pre_stride = DOCKAUV_MAX_CAPSULES * kPreCap + DOCKAUV_MAX_SPHERES * kPreSph;
        pre_off = off;  off += 4 * pre_stride * (int)sizeof(T);
        ray_stride = (n_rays + 1) & ~1;
        ray_off = off;  off += 4 * ray_stride * (int)sizeof(T);
        obs_stride = n_obs | 1;
        obs_off = off;  off += kWarpEnvs * obs_stride * (int)sizeof(float);
        flag_off = off; off += 2 * kWarpEnvs;
        off = (off + 15) & ~15;
        total = off + DOCKAUV_N_STATS * (int)sizeof(double);
    }
};

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T, int VEH, int NU, int RPL>
__global__ void __launch_bounds__(kWarpEnvs) step_warp_kernel(const __grid_constant__ KParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const WarpSmem<T> L(p.n_obs, p.n_rays);
    T *s_pose = reinterpret_cast<T *>(smem_raw + L.pose_off);
    T *s_oa = reinterpret_cast<T *>(smem_raw + L.oa_off);
    T *s_pre_all = reinterpret_cast<T *>(smem_raw + L.pre_off);
    T *s_ray_all = reinterpret_cast<T *>(smem_raw + L.ray_off);
    float *s_obs = reinterpret_cast<float *>(smem_raw + L.obs_off);
    unsigned char *s_col = smem_raw + L.flag_off;
    unsigned char *s_done = s_col + kWarpEnvs;
    BlockStats bs{reinterpret_cast<double *>(smem_raw + L.total - DOCKAUV_N_STATS * (int)sizeof(double))};
    bs.init();

    const int64_t N = p.n_envs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t i0 = p.env_begin + (int64_t)blockIdx.x * kWarpEnvs;
    const int64_t i = i0 + tid;
    const bool active = i < p.env_end;
    int64_t left64 = p.env_end - i0;
    const int n_here = (int)(left64 < (int64_t)kWarpEnvs ? left64 : (int64_t)kWarpEnvs);

    // ------------------------------------------------------------------ phase A
    StepCarry<T> cy;
    if (active) {
        T spsi, cpsi, att[3];
        float obs16[16];
        step_dynamics<T, VEH, NU>(p, i, cy, spsi, cpsi, obs16, att);
#pragma unroll
        for (int c = 0; c < 3; c++) s_pose[tid * 12 + c] = cy.pos[c];
#pragma unroll
        for (int c = 0; c < 9; c++) s_pose[tid * 12 + 3 + c] = cy.R[c];
#pragma unroll
        for (int c = 0; c < 16; c++) s_obs[tid * L.obs_stride + c] = obs16[c];
    }
    __syncwarp();

    // ------------------------------------------------------------------ phase B
    {
        T *s_pre = s_pre_all + warp * L.pre_stride;
        T *s_ray = s_ray_all + warp * L.ray_stride;
        const int n_caps = p.n_caps, n_sph = p.n_sph, n_obst = n_caps + n_sph;
        const int n_r = p.n_rays;
        const T dmax = p.radar_max_dist, R_safe = p.safety_radius;
        const T cull = dmax * T(1.000001);
        // this lane's rays: body-frame direction and obstacle-avoidance weight stay in registers
        T rb[RPL][3], bw[RPL];
#pragma unroll
        for (int j = 0; j < RPL; j++) {
            int ir = lane + 32 * j;
            bool ok = ir < n_r;
#pragma unroll
            for (int c = 0; c < 3; c++) rb[j][c] = ok ? p.ray_tab[c * n_r + ir] : T(0);
            bw[j] = ok ? p.ray_tab[3 * n_r + ir] : T(0);
        }
        const int e_begin = warp * 32;
        const int e_end = min(e_begin + 32, n_here);
        for (int e = e_begin; e < e_end; e++) {
            const int64_t ie = i0 + e;
            const T *pose = s_pose + e * 12;
            // ---- ray-independent algebra + body collision, one obstacle per lane
            bool hit_body = false, in_range = false;
            if (lane < n_obst) {
                T pos[3] = {pose[0], pose[1], pose[2]};
                if (lane < n_caps) {
                    T bot[3], top[3], rad;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        bot[c] = p.capsules[(int64_t)(lane * 7 + c) * N + ie];
                        top[c] = p.capsules[(int64_t)(lane * 7 + 3 + c) * N + ie];
                    }
                    rad = p.capsules[(int64_t)(lane * 7 + 6) * N + ie];
                    CapPre<T> q;
                    capsule_pre<T>(pos, bot, top, rad, q);
                    T *w = s_pre + lane * kPreCap;
                    w[0] = q.ba[0]; w[1] = q.ba[1]; w[2] = q.ba[2];
                    w[3] = q.oa[0]; w[4] = q.oa[1]; w[5] = q.oa[2];
                    w[6] = q.baba; w[7] = q.baoa; w[8] = q.c; w[9] = q.c2a; w[10] = q.c2b;
                    T dist = dist_segment_point<T>(pos, bot, top);
                    hit_body = dist <= rad + R_safe;                       // shape.py:195-210
                    in_range = !(dist - rad > cull);
                } else {
                    const int k = lane - n_caps;
                    T oc[3], rad, d2 = T(0);
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        oc[c] = pos[c] - p.spheres[(int64_t)(k * 4 + c) * N + ie];
                        d2 += oc[c] * oc[c];
                    }
                    rad = p.spheres[(int64_t)(k * 4 + 3) * N + ie];
                    T *w = s_pre + DOCKAUV_MAX_CAPSULES * kPreCap + k * kPreSph;
                    w[0] = oc[0]; w[1] = oc[1]; w[2] = oc[2]; w[3] = d2 - rad * rad;
                    T dist = Mth<T>::sqrt_(d2);
                    hit_body = dist <= R_safe + rad;                       // shape.py:182-192
                    in_range = !(dist - rad > cull);
                }
            }
            const unsigned col_mask = __ballot_sync(0xffffffffu, hit_body);
            unsigned near_mask = __ballot_sync(0xffffffffu, in_range);
            __syncwarp();

            // ---- cast this lane's rays
            T R[9];
#pragma unroll
            for (int c = 0; c < 9; c++) R[c] = pose[3 + c];
            T rd[RPL][3], best[RPL], first[RPL];
#pragma unroll
            for (int j = 0; j < RPL; j++) {
#pragma unroll
                for (int c = 0; c < 3; c++)
                    rd[j][c] = R[3 * c] * rb[j][0] + R[3 * c + 1] * rb[j][1] + R[3 * c + 2] * rb[j][2];
                best[j] = Mth<T>::inf();
                first[j] = -Mth<T>::inf();
            }
            unsigned cap_mask = near_mask & ((1u << n_caps) - 1u);
            unsigned sph_mask = (n_caps < 32 ? (near_mask >> n_caps) : 0u) & ((1u << n_sph) - 1u);
            while (cap_mask) {
                const int k = __ffs(cap_mask) - 1;
                cap_mask &= cap_mask - 1;
                const T *w = s_pre + k * kPreCap;
                CapPre<T> q;
                q.ba[0] = w[0]; q.ba[1] = w[1]; q.ba[2] = w[2];
                q.oa[0] = w[3]; q.oa[1] = w[4]; q.oa[2] = w[5];
                q.baba = w[6]; q.baoa = w[7]; q.c = w[8]; q.c2a = w[9]; q.c2b = w[10];
#pragma unroll
                for (int c = 0; c < 3; c++) q.oc2[c] = q.oa[c] - q.ba[c];
                q.r = T(0);
#pragma unroll
                for (int j = 0; j < RPL; j++) {
                    T v = ray_capsule<T>(q, rd[j]);
                    if (k == 0) first[j] = v;
                    if (v > T(0) && v < best[j]) best[j] = v;
                }
            }
            if (n_sph > 0) {
                T sbest[RPL], sfirst[RPL];
#pragma unroll
                for (int j = 0; j < RPL; j++) {
                    sbest[j] = Mth<T>::inf();
                    sfirst[j] = -Mth<T>::inf();
                }
                while (sph_mask) {
                    const int k = __ffs(sph_mask) - 1;
                    sph_mask &= sph_mask - 1;
                    const T *w = s_pre + DOCKAUV_MAX_CAPSULES * kPreCap + k * kPreSph;
                    T oc[3] = {w[0], w[1], w[2]};
                    T c = w[3];
#pragma unroll
                    for (int j = 0; j < RPL; j++) {
                        T v = ray_sphere<T>(oc, c, rd[j]);
                        if (k == 0) sfirst[j] = v;
                        if (v > T(0) && v < sbest[j]) sbest[j] = v;
                    }
                }
#pragma unroll
                for (int j = 0; j < RPL; j++) {
                    T v = (sbest[j] < Mth<T>::inf()) ? sbest[j] : sfirst[j];   // shape.py:264
                    if (n_caps == 0) first[j] = v;
                    if (v > T(0) && v < best[j]) best[j] = v;
                }
            }
            // ---- clamp (sensor.py:117), obstacle-avoidance partial sum (docking3d.py:792), stash for pooling
            T oa_part = T(0);
#pragma unroll
            for (int j = 0; j < RPL; j++) {
                const int ir = lane + 32 * j;
                if (ir < n_r) {
                    T d = dmax;
                    if (n_obst > 0) {
                        d = (best[j] < Mth<T>::inf()) ? best[j] : first[j];      // docking3d.py:438-439
                        if (d < T(0) || d > dmax) d = dmax;
                    }
                    s_ray[ir] = d;
                    if (p.dbg_ray_dist) p.dbg_ray_dist[(int64_t)ir * N + ie] = d;
                    T c = clipv(T(1) - d / dmax, T(0), T(1));
                    T qq = (T(1) - c) * (T(1) - c);
                    T mx = (qq != qq) ? qq : (qq > T(0.001) ? qq : T(0.001));
                    oa_part += mx * bw[j];
                }
            }
            const T oa_dot = warp_sum<T>(oa_part);
            __syncwarp();
            // ---- 2x2 max-pool with zero padding (sensor.py:131-137) -> obs[16:]
            for (int pc = lane; pc < p.n_rr; pc += 32) {
                const int pr = pc / p.n_hr, pcol = pc - pr * p.n_hr;
                T mx = T(0);
                bool nan = false;
                for (int dv = 0; dv < p.block; dv++)
                    for (int dh = 0; dh < p.block; dh++) {
                        const int rv = pr * p.block + dv, rh = pcol * p.block + dh;
                        if (rv < p.n_vert && rh < p.n_horiz) {
                            T v = s_ray[rv * p.n_horiz + rh];
                            nan |= (v != v);
                            mx = v > mx ? v : mx;
                        }
                    }
                T o = nan ? Mth<T>::nan() : clipv(mx / dmax, T(0), T(1));
                s_obs[e * L.obs_stride + 16 + pc] = (float)o;
                if (p.dbg_obs) p.dbg_obs[(int64_t)(16 + pc) * N + ie] = o;
            }
            if (lane == 0) {
                s_oa[e] = p.sum_beta_oa / oa_dot - T(1);
                s_col[e] = col_mask != 0u;
            }
            __syncwarp();
        }
    }

    // ------------------------------------------------------------------ phase C
    if (active) {
        bool done = step_finish<T>(p, i, cy, s_oa[tid], s_col[tid] != 0, bs);
        s_done[tid] = done;
    }
    bs.flush(p.stats, n_here);   // contains the __syncthreads that also publishes s_obs / s_done

    // ---- stream the observation tile out: rows are contiguous in HBM, so the copy is a flat coalesced store
    {
        const int n_obs = p.n_obs;
        const int total = n_here * n_obs;
        float *gobs = p.obs + i0 * n_obs;
        float *gterm = p.terminal_obs ? p.terminal_obs + i0 * n_obs : nullptr;
        const bool ar = p.auto_reset != 0;
        for (int f = tid; f < total; f += kWarpEnvs) {
            const int row = f / n_obs, col = f - row * n_obs;
            const float v = s_obs[row * L.obs_stride + col];
            const bool dn = s_done[row] != 0;
            if (dn && gterm) gterm[f] = v;
            gobs[f] = (dn && ar) ? 0.0f : v;     // reset() hands back the all-zero observation (docking3d.py:269,322)
        }
    }
}

template <typename T, int VEH, int NU, int RPL>
static cudaError_t launch_step_warp_rpl(const KParams<T> &k, cudaStream_t st) {
    const int64_t n = k.env_end - k.env_begin;
    const WarpSmem<T> L(k.n_obs, k.n_rays);
    auto kern = step_warp_kernel<T, VEH, NU, RPL>;
    if (L.total > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
        if (e != cudaSuccess) return e;
    }
    const unsigned blocks = (unsigned)((n + kWarpEnvs - 1) / kWarpEnvs);
    kern<<<blocks, kWarpEnvs, L.total, st>>>(k);
    return cudaGetLastError();
}

template <typename T, int VEH, int NU>
static cudaError_t launch_step_warp(const KParams<T> &k, cudaStream_t st) {
    // rays per lane: 2 covers the stock 63-ray and the 64-ray radar; 8 covers everything up to DOCKAUV_MAX_RAYS
    if (k.n_rays <= 64) return launch_step_warp_rpl<T, VEH, NU, 2>(k, st);
    return launch_step_warp_rpl<T, VEH, NU, 8>(k, st);
}

}  // namespace dockauv
