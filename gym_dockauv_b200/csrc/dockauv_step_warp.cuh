// dockauv_step_warp.cuh -- layout DOCKAUV_LAYOUT_WARP_RAYS.
//
// One CTA of 128 threads owns 128 consecutive envs; each warp works only on the 32 envs of its own lanes, so the
// phases below are separated by __syncwarp() only.
//   phase A (thread per env) : current, command filter, RKF45, angle wrap, nav errors, obs[0:16], done bits 0..3 and
//                              the radar-independent reward terms; the post-step pose goes to shared memory.
//   phase B (warp per env)   : the 32 envs are taken in sub-batches of 32 / slots envs (slots = 8 or 16 obstacle
//                              slots per env).  Pass 1 of a sub-batch uses ALL lanes, one (env, obstacle) pair per
//                              lane: coalesced-by-sector obstacle loads, the ray-independent algebra, the
//                              body-collision test and the radar-range cull; two ballots publish the collision and
//                              in-range bits of the whole sub-batch.  Pass 2 walks over the envs of the sub-batch:
//                              every lane casts its rays (ray = lane + 32 j) against the in-range obstacles (uniform
//                              loop, broadcast shared reads); the per-ray minimum is lane-local, the
//                              obstacle-avoidance sum is a shuffle reduction whose result stays in a register of
//                              the env's owner lane, the 2x2 max-pool goes through a per-warp scratch.
//   phase C (thread per env) : reward, done, counters, statistics, auto-reset; then each warp streams the
//                              observation rows of its envs to HBM (one coalesced store per row).
#pragma once
#include "dockauv_step_tpe.cuh"

namespace dockauv {

constexpr int kWarpEnvs = 128;       // envs (= threads) per CTA
constexpr int kPoseStride = 14;      // shared words per env: pos[3] R[9] poison pad
constexpr int kPreStride = 14;       // shared words per (env, obstacle) record, 16-byte aligned for 128-bit reads
                                     // capsule: ba[3] oa[3] baba baoa c c2a c2b ; sphere: oc[3] c

template <typename T>
struct WarpSmem {
    // layout of the dynamic shared memory block (offsets in bytes, computed on host and device alike)
    int pose_off, pre_off, ray_off, total;
    int ray_stride;     // T words per warp of ray-distance scratch (+1 zero slot for the pooling padding)
    __host__ __device__ WarpSmem(int n_rays) {
        int off = 0;
        pose_off = off; off += kWarpEnvs * kPoseStride * (int)sizeof(T);
        pre_off = off;  off += 4 * 32 * kPreStride * (int)sizeof(T);
        ray_stride = (n_rays + 2) & ~1;
        ray_off = off;  off += 4 * ray_stride * (int)sizeof(T);
        total = (off + 15) & ~15;
    }
};

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// two shared words with one 128-bit (double) / 64-bit (float) load
template <typename T>
struct Pair;
template <>
struct Pair<double> {
    using type = double2;
};
template <>
struct Pair<float> {
    using type = float2;
};

// no min-CTAs hint on purpose: __launch_bounds__(128, 3) ends at the same 168 registers but a 6 % slower schedule,
// (128, 4) and (128, 5) spill (measured on B200, profiles/r01/NOTES.md)
// MODE 0: fused step (phases A, B, C in one launch).
// MODE 1 / 2: the same step as two launches (layout DOCKAUV_LAYOUT_SPLIT): MODE 1 runs phase A and parks the post-step
// pose and the radar-independent reward terms in a hand-off buffer (kHandoffWords words + one flag word per env, SoA);
// MODE 2 picks them up and runs phases B and C.  Each launch then gets its own register allocation (phase A needs
// 168 registers, phases B/C far fewer -> more resident warps for the latency-bound ray pass) and the ray loop no
// longer shares the instruction cache with the 70 KB of straight-line integrator code.
constexpr int kHandoffWords = 22;    // pos[3] R[9] poison | reward terms r[0..7] | delta_d

#ifndef DOCKAUV_MINB_A
#define DOCKAUV_MINB_A 4
#endif
#ifndef DOCKAUV_A_PREFETCH
#define DOCKAUV_A_PREFETCH 37      // CTAs ahead whose inputs the dynamics launch prefetches into L2 (0 = off)
#endif
#ifndef DOCKAUV_MINB_B
#define DOCKAUV_MINB_B 4
#endif
// min-CTAs hints (0 = none): the fused kernel is fastest without one; measured for the split pair in profiles/r01/NOTES.md
// DBG: debug outputs compiled in (fused kernel only; the split pair is always built without them).
template <typename T, int VEH, int NU, int RPL, int MODE, bool DBG>
__global__ void __launch_bounds__(kWarpEnvs, MODE == 1 ? DOCKAUV_MINB_A : (MODE == 2 ? DOCKAUV_MINB_B : 0))
step_warp_kernel(const __grid_constant__ KParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using P2 = typename Pair<T>::type;
    const WarpSmem<T> L(p.n_rays);
    T *s_pose = reinterpret_cast<T *>(smem_raw + L.pose_off);
    WarpStats bs;

    const int64_t N = p.n_envs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t i0 = p.env_begin + (int64_t)blockIdx.x * kWarpEnvs;
    const int64_t i = i0 + tid;
    const bool active = i < p.env_end;
    int64_t left64 = p.env_end - i0;
    const int n_here = (int)(left64 < (int64_t)kWarpEnvs ? left64 : (int64_t)kWarpEnvs);
    const int e_warp = warp * 32;                         // first env (CTA-local) of this warp
    const int n_warp = max(0, min(32, n_here - e_warp));  // envs this warp really has

    // the pipeline layout appends to a list in its second launch: the first one empties it
    if (MODE == 1 && p.view_count != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *p.view_count = 0u;
#if DOCKAUV_A_PREFETCH > 0
    // CTAs are dispatched in blockIdx order: the inputs of a CTA a little further down the order are pulled into L2
    // now, while this one integrates (a third of a dynamics warp's lifetime was spent waiting for its own first HBM
    // round trip: 31 % of the launch's stall samples).  Measured: distance 8..148 CTAs 218 us, 300 235 us, 592 (one
    // resident wave) and beyond no gain, none 253 us.  The same trick does not help the cull / finish launches.
    if (MODE != 2) {
        const int64_t j = i + (int64_t)DOCKAUV_A_PREFETCH * kWarpEnvs;
        if (j < p.env_end) {
            auto pf = [](const void *a) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a)); };
#pragma unroll
            for (int c = 0; c < 12; c++) pf(p.state + (int64_t)c * N + j);
#pragma unroll
            for (int c = 0; c < NU; c++) pf(p.u_prev + (int64_t)c * N + j);
#pragma unroll
            for (int c = 0; c < 3; c++) pf(p.goal + (int64_t)c * N + j);
            if ((lane & 3) == 0) pf((const char *)p.actions + (p.act_f32 ? 4 : 8) * NU * j);
            if ((lane & 7) == 0) pf(p.t_steps + j);
        }
    }
#endif
    // ------------------------------------------------------------------ phase A
    StepCarry<T> cy;
    if (MODE != 2 && active) {
        T spsi, cpsi, att[3];
        float obs16[16];
        step_dynamics<T, VEH, NU, DBG>(p, i, cy, spsi, cpsi, obs16, att);
        // 0 for a finite pose, NaN otherwise: added to every ray distance so that a blown-up state poisons the
        // radar outputs exactly like the reference's NaN propagation does
        const T poison = (((cy.pos[0] + cy.pos[1]) + (cy.pos[2] + att[0])) + (att[1] + att[2])) * T(0);
        if (MODE == 0) {
            T *ps = s_pose + tid * kPoseStride;
#pragma unroll
            for (int c = 0; c < 3; c++) ps[c] = cy.pos[c];
#pragma unroll
            for (int c = 0; c < 9; c++) ps[3 + c] = cy.R[c];
            ps[12] = poison;
        } else {
            T *hf = p.handoff + i;
#pragma unroll
            for (int c = 0; c < 3; c++) hf[(int64_t)c * N] = cy.pos[c];
#pragma unroll
            for (int c = 0; c < 9; c++) hf[(int64_t)(3 + c) * N] = cy.R[c];
            hf[(int64_t)12 * N] = poison;
#pragma unroll
            for (int c = 0; c < 8; c++) hf[(int64_t)(13 + c) * N] = cy.rarr[c];
            hf[(int64_t)21 * N] = cy.delta_d;
            p.handoff_cond[i] = cy.cond;
        }
        // obs[0:16] goes straight to its HBM row (four 16-byte stores); phase C zeroes the row if the env is reset
        float4 *orow4 = reinterpret_cast<float4 *>(p.obs + i * p.n_obs);
        if ((p.n_obs & 3) == 0) {
#pragma unroll
            for (int c = 0; c < 4; c++)
                orow4[c] = make_float4(obs16[4 * c], obs16[4 * c + 1], obs16[4 * c + 2], obs16[4 * c + 3]);
        } else {
            float *orow = p.obs + i * p.n_obs;
#pragma unroll
            for (int c = 0; c < 16; c++) orow[c] = obs16[c];
        }
    }
    if (MODE == 1) return;
    if (MODE == 2 && active) {
        const T *hf = p.handoff + i;
        T *ps = s_pose + tid * kPoseStride;
#pragma unroll
        for (int c = 0; c < 13; c++) ps[c] = hf[(int64_t)c * N];
        // the reward terms and counters of the hand-off are only read in phase C: they are loaded there, so that they
        // do not occupy ~24 registers during the whole radar phase
    }
    __syncwarp();

    // ------------------------------------------------------------------ phase B
    T my_oa_dot = p.sum_beta_oa;     // owner-lane copies of the per-env radar results (neutral values: r_oa = 0)
    bool my_col = false;
    bool my_view_empty = true;       // nothing within range and view: the owner lane writes the all-ones ray cells
    const bool no_dbg = !DBG || (p.dbg_ray_dist == nullptr && p.dbg_obs == nullptr);
    {
        T *s_pre = reinterpret_cast<T *>(smem_raw + L.pre_off) + warp * 32 * kPreStride;
        T *s_ray = reinterpret_cast<T *>(smem_raw + L.ray_off) + warp * L.ray_stride;
        const int n_caps = p.n_caps, n_sph = p.n_sph, n_obst = n_caps + n_sph;
        const int n_r = p.n_rays;
        const T dmax = p.radar_max_dist, inv_dmax = T(1) / dmax;
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
        // pooling fast path (2x2 blocks, at most one pooled cell per lane): the four source slots of this lane's
        // cell; cells beyond the ray grid read the zero slot s_ray[n_r] (block_reduce pads with cval = 0)
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
        // obstacle slot of this lane in pass 1
        const int slots = n_obst <= 8 ? 8 : 16;
        const int epp = 32 / slots;                        // envs per sub-batch
        const unsigned slot_mask = (1u << slots) - 1u;
        const int my_slot = lane & (slots - 1), my_sub = lane / slots;
        const bool slot_is_cap = my_slot < n_caps, slot_used = my_slot < n_obst;
        const T *obst_row = slot_is_cap ? p.capsules + (int64_t)(my_slot * 7) * N
                                        : p.spheres + (int64_t)((my_slot - n_caps) * 4) * N;

        // raw obstacle of this lane's (env, slot) pair: loaded one sub-batch ahead so that the HBM latency of the
        // next pre-pass is covered by the ray loop of the current one
        auto load_obstacle = [&](int eb_, T ob[7]) {
            const int e_ = eb_ + my_sub;
            if (slot_used && e_ < n_warp) {
                const T *g = obst_row + (i0 + e_warp + e_);
                const int n_words = slot_is_cap ? 7 : 4;
#pragma unroll
                for (int c = 0; c < 7; c++)
                    if (c < n_words) ob[c] = g[(int64_t)c * N];
            }
        };
        T ob_next[7];
#pragma unroll
        for (int c = 0; c < 7; c++) ob_next[c] = T(0);
        load_obstacle(0, ob_next);

        for (int eb = 0; eb < n_warp; eb += epp) {
            T ob[7];
#pragma unroll
            for (int c = 0; c < 7; c++) ob[c] = ob_next[c];
            if (eb + epp < n_warp) load_obstacle(eb + epp, ob_next);
            // ---- pass 1: one (env, obstacle) pair per lane
            bool hit_body = false, in_range = false;
            {
                const int e = eb + my_sub;
                if (slot_used && e < n_warp) {
                    const T *pose = s_pose + (e_warp + e) * kPoseStride;
                    const T pos[3] = {pose[0], pose[1], pose[2]};
                    T *w = s_pre + lane * kPreStride;
                    obstacle_pair<T, true>(p, pos, pose + 3, ob, slot_is_cap, w, hit_body, in_range);
                }
            }
            const unsigned colb = __ballot_sync(0xffffffffu, hit_body);
            const unsigned nearb = __ballot_sync(0xffffffffu, in_range);
#ifdef DOCKAUV_VIEW_STATS   // tuning builds: in-view (env, obstacle) pairs and envs with a non-empty view -> stats[11], [12]
            if (lane == 0) {
                int n_env_view = 0;
                for (int s = 0; s < epp; s++) n_env_view += ((nearb >> (slots * s)) & slot_mask) != 0u;
                atomicAdd(&p.stats[11], (double)__popc(nearb));
                atomicAdd(&p.stats[12], (double)n_env_view);
            }
#endif
            {   // the owner lane of each env of this sub-batch keeps its collision flag for phase C
                const int s = lane - eb;
                if (s >= 0 && s < epp) {
                    my_col = ((colb >> (slots * s)) & slot_mask) != 0u;
                    my_view_empty = ((nearb >> (slots * s)) & slot_mask) == 0u;
                }
            }
            __syncwarp();

            // ---- pass 2: cast rays, env by env
            for (int s = 0; s < epp && eb + s < n_warp; s++) {
                const int e = eb + s;
                const int64_t ie = i0 + e_warp + e;
                const T *pose = s_pose + (e_warp + e) * kPoseStride;
                const unsigned near_mask = (nearb >> (slots * s)) & slot_mask;
                T best[RPL];
#pragma unroll
                for (int j = 0; j < RPL; j++) best[j] = Mth<T>::inf();
                if (near_mask) {
                    T R[9];
#pragma unroll
                    for (int c = 0; c < 9; c++) R[c] = pose[3 + c];
                    T rd[RPL][3];
#pragma unroll
                    for (int j = 0; j < RPL; j++) {
#pragma unroll
                        for (int c = 0; c < 3; c++)
                            rd[j][c] = R[3 * c] * rb[j][0] + R[3 * c + 1] * rb[j][1] + R[3 * c + 2] * rb[j][2];
                    }
                    const T *pre_env = s_pre + (s * slots) * kPreStride;
                    unsigned cap_mask = near_mask & ((1u << n_caps) - 1u);
                    unsigned sph_mask = (near_mask >> n_caps) & ((1u << n_sph) - 1u);
                    while (cap_mask) {
                        const int k = __ffs(cap_mask) - 1;
                        cap_mask &= cap_mask - 1;
                        const P2 *w2 = reinterpret_cast<const P2 *>(pre_env + k * kPreStride);
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
                        const P2 *w2 = reinterpret_cast<const P2 *>(pre_env + (n_caps + k) * kPreStride);
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
                const T poison = pose[12];
                if (near_mask == 0u && poison == T(0) && no_dbg) {
                    // nothing within range and view: every ray reads max_dist (sensor.py:113-117), the pooled
                    // observation is all ones (written by the owner lane in phase C) and the obstacle-avoidance
                    // sum keeps its neutral value sum(beta) (r_oa = 0)
                    continue;
                }
                // ---- clamp (sensor.py:117), obstacle-avoidance partial sum (docking3d.py:767-792), stash for pooling
                T oa_part = T(0);
#pragma unroll
                for (int j = 0; j < RPL; j++) {
                    const int ir = lane + 32 * j;
                    if (ir < n_r) {
                        // min positive distance over obstacles (docking3d.py:438-439), max_dist if none or farther
                        T d = (best[j] > dmax ? dmax : best[j]) + poison;
                        s_ray[ir] = d;
                        if (DBG && p.dbg_ray_dist) p.dbg_ray_dist[(int64_t)ir * N + ie] = d;
                        // (gamma_c (1 - c))^2 with c = clip(1 - d/d_max, 0, 1): 1 - c = d/d_max for d in [0, d_max]
                        const T x = d * inv_dmax;
                        const T qq = x * x;
                        const T mx = !(qq <= T(0.001)) ? qq : T(0.001);     // np.maximum, NaN propagates
                        oa_part += mx * bw[j];
                    }
                }
                const T oa_dot = warp_sum<T>(oa_part);
                if (lane == e) my_oa_dot = oa_dot;
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
                        if (DBG && p.dbg_obs) p.dbg_obs[(int64_t)(16 + lane) * N + ie] = o;
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
                        if (DBG && p.dbg_obs) p.dbg_obs[(int64_t)(16 + pc) * N + ie] = o;
                    }
                }
                __syncwarp();
            }
        }
    }

    // ------------------------------------------------------------------ phase C
    bool done = false;
    if (MODE == 2 && active) {
        const T *hf = p.handoff + i;
#pragma unroll
        for (int c = 0; c < 8; c++) cy.rarr[c] = hf[(int64_t)(13 + c) * N];
        cy.delta_d = hf[(int64_t)21 * N];
        cy.cond = p.handoff_cond[i];
        cy.t_steps = p.t_steps[i];
        cy.ep_return = p.ep_return[i];
    }
    if (active) {
        if (my_view_empty && no_dbg && s_pose[tid * kPoseStride + 12] == T(0)) {
            float *cells = p.obs + i * p.n_obs + 16;
            if ((p.n_obs & 3) == 0 && (p.n_rr & 3) == 0) {
                float4 *c4 = reinterpret_cast<float4 *>(cells);
                for (int c = 0; c < (p.n_rr >> 2); c++) c4[c] = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
            } else {
                for (int c = 0; c < p.n_rr; c++) cells[c] = 1.0f;
            }
        }
        const T r_oa = p.sum_beta_oa / my_oa_dot - T(1);      // docking3d.py:792
        done = step_finish<T, DBG>(p, i, cy, r_oa, my_col, bs);
    }
    __syncwarp();     // orders the pooled-cell stores of the other lanes before the row fix-up below
    // ---- rows of envs whose episode ended (~1 % per step): keep the last observation as terminal_observation and
    //      hand back the all-zero reset observation (docking3d.py:269,322); the warp moves each such row together
    {
        unsigned dm = __ballot_sync(0xffffffffu, done);
        const int n_obs = p.n_obs;
        while (dm) {
            const int e = __ffs(dm) - 1;
            dm &= dm - 1;
            float *row = p.obs + (i0 + e_warp + e) * n_obs;
            float *trow = p.terminal_obs ? p.terminal_obs + (i0 + e_warp + e) * n_obs : nullptr;
            for (int c = lane; c < n_obs; c += 32) {
                if (trow) trow[c] = row[c];
                if (p.auto_reset) row[c] = 0.0f;
            }
        }
    }
    bs.flush(p.stats, warp == 0 ? n_here : 0);
}

template <typename T, int VEH, int NU, int RPL, int MODE, bool DBG>
static cudaError_t launch_step_warp_rpl(const KParams<T> &k, cudaStream_t st) {
    const int64_t n = k.env_end - k.env_begin;
    const WarpSmem<T> L(k.n_rays);
    const int smem = MODE == 1 ? 0 : L.total;
    auto kern = step_warp_kernel<T, VEH, NU, RPL, MODE, DBG>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    const unsigned blocks = (unsigned)((n + kWarpEnvs - 1) / kWarpEnvs);
    kern<<<blocks, kWarpEnvs, smem, st>>>(k);
    return cudaGetLastError();
}

template <typename T>
static bool wants_debug(const KParams<T> &k) {
    return k.dbg_ray_dist || k.dbg_reward_arr || k.dbg_euler_dot || k.dbg_nu_c || k.dbg_nav || k.dbg_obs || k.dbg_state_dot;
}

// rays per lane: 2 covers the stock 63-ray and the 64-ray radar; 8 covers everything up to DOCKAUV_MAX_RAYS
template <typename T, int VEH, int NU>
static cudaError_t launch_step_warp(const KParams<T> &k, cudaStream_t st) {
    if (wants_debug(k))
        return k.n_rays <= 64 ? launch_step_warp_rpl<T, VEH, NU, 2, 0, true>(k, st)
                              : launch_step_warp_rpl<T, VEH, NU, 8, 0, true>(k, st);
    if (k.n_rays <= 64) return launch_step_warp_rpl<T, VEH, NU, 2, 0, false>(k, st);
    return launch_step_warp_rpl<T, VEH, NU, 8, 0, false>(k, st);
}

// layout DOCKAUV_LAYOUT_SPLIT: dynamics launch + radar launch per chunk of `chunk` envs (default: one pair over the
// whole batch).  Debug outputs are only compiled into the fused kernel, which then serves the call.
template <typename T, int VEH, int NU>
static cudaError_t launch_step_split(const KParams<T> &k, int64_t chunk, cudaStream_t st) {
    if (wants_debug(k) || k.handoff == nullptr) return launch_step_warp<T, VEH, NU>(k, st);   // no obstacles: no hand-off buffer
    if (chunk <= 0) chunk = k.env_end - k.env_begin;
    chunk = ((chunk + kWarpEnvs - 1) / kWarpEnvs) * kWarpEnvs;
    for (int64_t b = k.env_begin; b < k.env_end; b += chunk) {
        KParams<T> kc = k;
        kc.env_begin = b;
        kc.env_end = b + chunk < k.env_end ? b + chunk : k.env_end;
        cudaError_t e = launch_step_warp_rpl<T, VEH, NU, 2, 1, false>(kc, st);
        if (e != cudaSuccess) return e;
        // the radar launch does not depend on the vehicle: one instantiation serves all of them
        e = (k.n_rays <= 64) ? launch_step_warp_rpl<T, DOCKAUV_VEHICLE_BLUEROV2, 6, 2, 2, false>(kc, st)
                             : launch_step_warp_rpl<T, DOCKAUV_VEHICLE_BLUEROV2, 6, 8, 2, false>(kc, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace dockauv
