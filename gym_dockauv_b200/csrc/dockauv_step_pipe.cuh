// dockauv_step_pipe.cuh -- layout DOCKAUV_LAYOUT_PIPELINE (default): one batched step as specialised launches.
//
//   1. dynamics     dynamics_kernel: thread per env.  Current, command filter, RKF45, angle wrap, navigation errors,
//      + cull        obs[0:16], done conditions 0..2, the reward terms that need no radar -- a 16-word record per env
//      + finish      (post-step attitude sines / cosines, position relative to the goal, the reward terms already combined
//                   in numpy's summation order, delta_d, condition bits) that stays in REGISTERS for the second half of the
//                   thread (FUSE): the walk over the env's obstacles in FLOAT (float4 records relative to the goal, written
//                   at reset, waiting in shared memory since the start of the thread; cull_pair_rec): body collision
//                   (re-decided in T when within 2 mm of the threshold) and the range / field-of-view culls.  Envs with
//                   something in view (24 % on the C4 workload) are appended to a compact list and their record goes to
//                   KParams::rec.  Done flag, condition bits and counters are final here for EVERY env (coalesced stores);
//                   an env with nothing in view is finished completely -- all ray cells read max_dist, r_oa = 0: reward,
//                   running return, statistics -- and goes on a second list if its episode ended.
//                   (Obstacle-free scenarios: FIN, the env is finished right after the dynamics.  More float records than
//                   fit in shared memory, or coordinates too large for them: the cull + finish code runs as a launch of its
//                   own, cull_finish_kernel, on records read back from KParams::rec.)
//   2. rays+finish  rays_thread_kernel: persistent grid.  The listed envs with one or two obstacles in view with LANES =
//                   ENVS (tiles of 128), the others by one warp each (lanes = rays, next entry in flight while the current
//                   one is cast; radar_env, dockauv_rays.cuh) on the first CTAs of the same launch; then the reward with its
//                   obstacle-avoidance term and the running return.
//   3. episode end  episode_end_kernel: the ENDED envs (~1 % of the batch, compacted by launches 1 and 2): terminal-
//                   observation row, zero row, re-initialisation (warp = reset role, lane = env).
//
// Each launch has its own register budget and occupancy, the ray code never idles on envs with nothing in view, and no
// env is touched by a launch that has nothing to do for it.  Results are identical to the other layouts (same device
// functions, same order of operations per env; the ray tiles rotate obstacle records into the body frame instead of rays
// into NED, which moves ray distances in their last bits); debug outputs are served by the single-launch kernel.
#pragma once
#include "dockauv_step_warp.cuh"

namespace dockauv {

#ifndef DOCKAUV_DYN_THREADS
#define DOCKAUV_DYN_THREADS 256     // CTA size of the dynamics launch (measured at 1M envs, fused form: 64 or 128 threads 252 us, 256 threads 239 us, 512 threads 266 us: the warps of a CTA run in step and share their instruction-cache lines, a 512-thread CTA drains the whole SM at once)
#endif
#ifndef DOCKAUV_MINB_A
#define DOCKAUV_MINB_A 2            // its min-CTAs hint: (256, 2) = 128 registers, 16 warps per SM
#endif
constexpr int kDynThreads = DOCKAUV_DYN_THREADS;
#ifndef DOCKAUV_MINB_CULL
#define DOCKAUV_MINB_CULL 4
#endif
#ifndef DOCKAUV_MINB_TPE
#define DOCKAUV_MINB_TPE 4          // thread-per-env ray launch: (128, 4) = 128 registers
#endif
#ifndef DOCKAUV_TPE_CTAS_PER_SM
#define DOCKAUV_TPE_CTAS_PER_SM 8   // persistent grid of the thread-per-env ray launch (twice what is resident: finer balance)
#endif
#ifndef DOCKAUV_TPE_WARP_CTAS_PER_SM
#define DOCKAUV_TPE_WARP_CTAS_PER_SM 1   // CTAs per SM of that launch that run the warp-per-env loop over class 3
#endif
#ifndef DOCKAUV_TPE_SPLIT
#define DOCKAUV_TPE_SPLIT 1         // lanes per env in the thread-per-env ray launch (each takes a run of pooled cells; 2: +8 %, 4: +11 % time)
#endif
#ifndef DOCKAUV_CULL_PREFETCH
#define DOCKAUV_CULL_PREFETCH 1     // (without staging) L2 prefetch of the records
#endif
#ifndef DOCKAUV_TPE_BODY
#define DOCKAUV_TPE_BODY 1          // thread-per-env ray tiles: obstacle records rotated into the body frame once instead of every ray into NED
#endif
#ifndef DOCKAUV_MINB_RAYS
#define DOCKAUV_MINB_RAYS 6         // ray launch: (128, 6) = 80 registers, 24 warps per SM
#endif
#ifndef DOCKAUV_RAY_CTAS_PER_SM
#define DOCKAUV_RAY_CTAS_PER_SM DOCKAUV_MINB_RAYS   // persistent grid of the ray launch, in 4-warp CTAs per SM (= what is resident)
#endif

// work-list counters per stepped env range (KParams::view_count): words 0..3 the live counters (view lists of class 1, 2, 3,
// ended list); steps with the cull code fused into the dynamics launch: 4..7 their values at the end of the most recent
// step (dockauv_last_list_counts), 8 the ticket of the episode-end launch, whose last CTA saves and zeroes the live counters
constexpr int kListCounters = 4, kCounterStride = 16, kCounterLast = 4, kCounterTicket = 8;

// ---- the per-env record written by the dynamics launch
constexpr int kRecWords = 16;
constexpr int REC_TRIG = 0, REC_PREL = 6, REC_A = 9, REC_B = 10, REC_R7 = 11, REC_LPD = 12, REC_DD = 13, REC_COND = 14,
              REC_POISON = 15;

// 16-byte vector access to a record: double -> 8 x double2, float -> 4 x float4
template <typename T>
struct RecIO;
template <>
struct RecIO<double> {
    static constexpr int kVecWords = 2;
    static __device__ __forceinline__ void store(double *dst, const double w[kRecWords]) {
        double2 *d = reinterpret_cast<double2 *>(dst);
#pragma unroll
        for (int c = 0; c < 8; c++) d[c] = make_double2(w[2 * c], w[2 * c + 1]);
    }
    // words [2 * first_vec, 2 * (first_vec + n_vec))
    template <int FIRST, int COUNT>
    static __device__ __forceinline__ void load(const double *src, double *w) {
        const double2 *s = reinterpret_cast<const double2 *>(src);
#pragma unroll
        for (int c = 0; c < COUNT; c++) {
            const double2 v = s[FIRST + c];
            w[2 * c] = v.x;
            w[2 * c + 1] = v.y;
        }
    }
};
template <>
struct RecIO<float> {
    static constexpr int kVecWords = 4;
    static __device__ __forceinline__ void store(float *dst, const float w[kRecWords]) {
        float4 *d = reinterpret_cast<float4 *>(dst);
#pragma unroll
        for (int c = 0; c < 4; c++) d[c] = make_float4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
    }
    // same word ranges as the double version: FIRST / COUNT are in units of two words
    template <int FIRST, int COUNT>
    static __device__ __forceinline__ void load(const float *src, float *w) {
        const float2 *s = reinterpret_cast<const float2 *>(src);
#pragma unroll
        for (int c = 0; c < COUNT; c++) {
            const float2 v = s[FIRST + c];
            w[2 * c] = v.x;
            w[2 * c + 1] = v.y;
        }
    }
};

#ifndef DOCKAUV_PDL
#define DOCKAUV_PDL 1               // ray and episode-end launches as programmatic dependents of the launch before them
#endif
// Launch as a PROGRAMMATIC DEPENDENT of the previous launch in the stream: the grid is set up while that launch drains and
// its threads wait in grid_dependency_wait() until everything the previous launch wrote is visible (no launch before it
// triggers early, so the wait is for its completion; an early trigger at the top of the launch before -- dependents set up
// during its last wave -- measured no faster).  What it saves is part of the launch-to-launch gap: 22.5 -> 22.3 us per
// 65,536-env SimpleDocking3d step, 0.391 -> 0.390 ms per 1M-env C4 step.
__device__ __forceinline__ void grid_dependency_wait() {
#if DOCKAUV_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
// one word global -> shared without a register in between (completion: cp_async_wait_all); dst = shared-window address
template <typename T>
__device__ __forceinline__ void cp_async_word(unsigned dst_shared, const T *src) {
    if (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_shared), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_shared), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------------- finish helpers
// is_done (docking3d.py:597-631): the five condition bits once the collision flag is known; t_steps is the counter before
// its increment (:612 is evaluated pre-increment, so max_timesteps = 1000 ends an episode at step 1001)
template <typename T>
__device__ __forceinline__ uint32_t done_conditions(const KParams<T> &p, uint32_t cond012, int32_t t_steps, bool collision) {
    return cond012 | ((t_steps >= p.max_timesteps) ? 8u : 0u) | (collision ? 16u : 0u);
}

// reward_step (docking3d.py:560-595) from the record words of the dynamics launch and the radar term r_oa
// (Reward.obstacle_avoidance, :767-792).  np.sum's order for 13 terms: ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)) + r8 .. r12.
template <typename T>
__device__ __forceinline__ T step_reward(const KParams<T> &p, T A, T B, T r7, T lp_d, T r_oa, uint32_t cond) {
    T r6;
    if (p.reward_set == 1) r6 = -p.w_oa * r_oa;
    else r6 = -p.w_oa * cont_goal_constraints<T>(Mth<T>::abs_(r_oa), T(1), lp_d);
    T reward = A + (B + (r6 + r7));
    if (cond != 0u) {      // ~1 % of the env-steps; the other terms are exact zeros
#pragma unroll
        for (int k = 0; k < 5; k++) reward += ((cond >> k) & 1u) ? p.w_done[k] : T(0);
    }
    return reward;
}

// ------------------------------------------------------------------------------------------------------- 1. dynamics
// CUR: the ocean current is evaluated (scenario with a current, injected current or noise); SPM: sparse M_inv / C / G;
// FIN: the scenario has no obstacles -- every ray reads max_dist and nothing can collide, so the env is finished right
// here (reward, done, counters, statistics, all-ones ray cells; no record, no cull / ray launch);
// FUSE: the cull + finish code (section 2) runs in this launch, on the record while it is still in registers.

// shared memory of one dynamics CTA: park[kParkWords][kDynThreads] words of T that are only needed after the integration
// (position, goal; FIN / FUSE: running return, step counter), filled by cp.async into the thread's own slots -- no register
// during the integration and no memory round trip after it
constexpr int kParkWords = 8;

// one env of the dynamics launch.  FIN: the env is finished here (steps_here: env-steps this call accounts for in the
// statistics -- the callers pass the CTA's count from one thread, 0 from the others).  Otherwise the env's record is left
// in w[kRecWords] and its post-step position in pos[3].
//   stage(): the caller's own cp.async copies for this env (issued after the plain loads)
template <typename T, int VEH, int NU, bool CUR, bool SPM, bool FIN, bool LATE, typename STAGE>
__device__ __forceinline__ void dynamics_env(const KParams<T> &p, const int64_t i, const int steps_here, T *dyn_park, T w[kRecWords],
                                             T pos[3], STAGE stage) {
    const int64_t N = p.n_envs;
    // ---- what the integration starts from goes first: plain loads; the asynchronous copies of everything that is needed
    //      later queue up behind them (issued ahead of the loads, the 21 copies of a thread delayed its first use of a loaded
    //      value: 10 % of the fused launch's stall samples sat on that one wait)
    T y[9];
#pragma unroll
    for (int c = 0; c < 9; c++) y[c] = p.state[(int64_t)(3 + c) * N + i];
    CommandIn<T, NU> cin;
    load_command<T, NU>(p, i, cin);
    {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(dyn_park + threadIdx.x);
        constexpr unsigned kPlane = kDynThreads * (unsigned)sizeof(T);
#pragma unroll
        for (int c = 0; c < 3; c++) cp_async_word<T>(sa + c * kPlane, p.state + (int64_t)c * N + i);
#pragma unroll
        for (int c = 0; c < 3; c++) cp_async_word<T>(sa + (3 + c) * kPlane, p.goal + (int64_t)c * N + i);
        if (LATE) {
            cp_async_word<T>(sa + 6 * kPlane, p.ep_return + i);
            cp_async_word<int32_t>(sa + 7 * kPlane, p.t_steps + i);
        }
    }
    stage();
    T tr0[6];
    Mth<T>::sincos_(y[0], &tr0[0], &tr0[1]);
    Mth<T>::sincos_(y[1], &tr0[2], &tr0[3]);
    Mth<T>::sincos_(y[2], &tr0[4], &tr0[5]);
    T nu_c[3] = {T(0), T(0), T(0)};
    if (CUR) dyn_current<T>(p, i, tr0, nu_c);
    T tau[6], penalty;
    {
        T u[NU];
        penalty = command_and_penalty<T, NU>(p, cin, u);
#pragma unroll
        for (int k = 0; k < NU; k++) p.u_prev[(int64_t)k * N + i] = u[k];
        dyn_tau<T, VEH, NU>(p, u, tau);
    }
    T pacc[3], tr1[6];
    rkf45_step<T, VEH, SPM, CUR>(p, y, tr0, tau, nu_c, pacc, tr1);
#pragma unroll
    for (int c = 0; c < 3; c++) y[c] = ssa<T>(y[c]);
    // position and goal are only needed from here on: they have been waiting in shared memory
    T goal[3];
    cp_async_wait_all();
#pragma unroll
    for (int c = 0; c < 3; c++) pos[c] = dyn_park[c * kDynThreads + threadIdx.x];
#pragma unroll
    for (int c = 0; c < 3; c++) goal[c] = dyn_park[(3 + c) * kDynThreads + threadIdx.x];
#pragma unroll
    for (int c = 0; c < 3; c++) pos[c] += pacc[c];
#pragma unroll
    for (int c = 0; c < 3; c++) p.state[(int64_t)c * N + i] = pos[c];
#pragma unroll
    for (int c = 0; c < 9; c++) p.state[(int64_t)(3 + c) * N + i] = y[c];

    DynOut<T> q;
    dyn_outputs<T>(p, pos, y, tr1, goal, nu_c, penalty, q);
    // obs[0:16] goes straight to its HBM row (four 16-byte stores); a later launch zeroes the row if the env is reset
    if ((p.n_obs & 3) == 0) {
        float4 *orow4 = reinterpret_cast<float4 *>(p.obs + i * p.n_obs);
#pragma unroll
        for (int c = 0; c < 4; c++)
            orow4[c] = make_float4((float)q.o[4 * c], (float)q.o[4 * c + 1], (float)q.o[4 * c + 2], (float)q.o[4 * c + 3]);
    } else {
        float *orow = p.obs + i * p.n_obs;
#pragma unroll
        for (int c = 0; c < 16; c++) orow[c] = (float)q.o[c];
    }
    // np.sum of the 13 reward terms is ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)) + r8 + .. + r12 (reward_sum13): the two
    // sums that need no radar are formed here, in that order
    const T A = (q.r[0] + q.r[1]) + (q.r[2] + q.r[3]), B = q.r[4] + q.r[5];
    if (FIN) {
        // ---- every ray reads max_dist (docking3d.py:441, sensor.py:113-117) -> pooled cells all ones, r_oa = 0
        float *cells = p.obs + i * p.n_obs + 16;
        if ((p.n_obs & 3) == 0 && (p.n_rr & 3) == 0) {
            float4 *c4 = reinterpret_cast<float4 *>(cells);
            for (int c = 0; c < (p.n_rr >> 2); c++) c4[c] = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
        } else {
            for (int c = 0; c < p.n_rr; c++) cells[c] = 1.0f;
        }
        const int32_t t_steps = reinterpret_cast<const int32_t *>(dyn_park + 7 * kDynThreads + threadIdx.x)[0];
        const T ep_before = dyn_park[6 * kDynThreads + threadIdx.x];
        const uint32_t cond = done_conditions<T>(p, q.cond, t_steps, false);
        const bool done = cond != 0;
        const int32_t t_new = t_steps + 1;
        const T r_oa = p.sum_beta_oa / p.sum_beta_oa - T(1);
        const T reward = step_reward<T>(p, A, B, q.r[7], q.r[6], r_oa, cond);
        const T ep_ret = ep_before + reward;
        p.reward[i] = reward;
        p.done[i] = done ? 1 : 0;
        if (p.cond_bits) p.cond_bits[i] = (uint8_t)cond;
        if (p.delta_d_out) p.delta_d_out[i] = q.delta_d;
        WarpStats bs;
        if (done) {
            if (p.ep_len_out) p.ep_len_out[i] = t_new;
            if (p.ep_return_out) p.ep_return_out[i] = ep_ret;
            bs.done = true;
            bs.cond = cond;
            bs.length = t_new;
            bs.ep_return = (double)ep_ret;
            bs.delta_d = (double)q.delta_d;
            bs.nan = reward != reward;
        }
        if (!(done && p.auto_reset)) {
            p.ep_return[i] = ep_ret;
            p.t_steps[i] = t_new;
        }
        {   // ended episodes -> the list of the episode-end launch (aggregated over the lanes that are still here)
            const unsigned am = __activemask();
            const unsigned em = __ballot_sync(am, done);
            if (em) {
                const int lane = threadIdx.x & 31, leader = __ffs(am) - 1;
                unsigned base = 0;
                if (lane == leader) base = atomicAdd(&p.view_count[3], (unsigned)__popc(em));
                base = __shfl_sync(am, base, leader);
                if (done) p.ended_list[base + __popc(em & ((1u << lane) - 1u))] = (uint32_t)(i - p.env_begin);
            }
        }
        bs.flush_direct(p.stats, steps_here);
        return;
    }
#pragma unroll
    for (int c = 0; c < 6; c++) w[REC_TRIG + c] = tr1[c];
#pragma unroll
    for (int c = 0; c < 3; c++) w[REC_PREL + c] = pos[c] - goal[c];
    w[REC_A] = A;
    w[REC_B] = B;
    w[REC_R7] = q.r[7];
    w[REC_LPD] = q.r[6];
    w[REC_DD] = q.delta_d;
    w[REC_COND] = (T)q.cond;
    // 0 for a finite pose, NaN otherwise: added to every ray distance so that a blown-up state poisons the radar
    // outputs exactly like the reference's NaN propagation does
    w[REC_POISON] = (((pos[0] + pos[1]) + (pos[2] + y[0])) + (y[1] + y[2])) * T(0);
}

// ------------------------------------------------------------------------------------------------------- 2. cull + finish
// One env through the float culls and everything that can be finished without rays; called by EVERY thread of a CTA of
// CTA threads (the list appends are warp-collective; `active` = the thread has an env).
//   trig / prel      sin / cos of the post-step attitude, position relative to the goal (record words 0..8)
//   fetch(slot)      float obstacle record `slot` of this env (KParams::obsf layout), from wherever the caller keeps it
//   late()           CullLate: the record words 9..15, running return and step counter -- asked for AFTER the obstacle loop,
//                    so that a caller that has them in memory does not hold them in registers through the loop
//   pos_of(q)        post-step position (only for pairs within 2 mm of the collision threshold)
// Returns whether the env is on a view list (its record must then be in KParams::rec for the ray launch).
template <typename T>
struct CullLate {
    T A, B, r7, lp_d, delta_d, cond012, poison, ep_return;
    int32_t t_steps;
};

template <typename T, int CTA, typename FETCH, typename LATE, typename POS>
__device__ __forceinline__ bool cull_finish_env(const KParams<T> &p, const int64_t i, const int64_t i0, const bool active,
                                                const T trig[6], const T prel_T[3], FETCH fetch, LATE late, POS pos_of) {
    const int64_t N = p.n_envs;
    const int lane = threadIdx.x & 31;
    WarpStats bs;
    bool listed = false, ended = false;      // ended: episode over and nothing in view -> on the list of the episode-end launch
    bool poison_free = true;
    uint32_t info = 0;                       // bits 0..15 in-view mask (capsules first), 16 collision
    uint32_t cond = 0;
    if (active) {
        const int n_caps = p.n_caps, n_sph = p.n_sph, n_obst = n_caps + n_sph;
        if (n_obst > 0 && !p.cull_exact) {
            // ---- float culls + collision pre-test (cull_pair_rec)
            float Rf[9], prel[3];
            rzyx<float>((float)trig[0], (float)trig[1], (float)trig[2], (float)trig[3], (float)trig[4], (float)trig[5], Rf);
#pragma unroll
            for (int c = 0; c < 3; c++) prel[c] = (float)prel_T[c];
            // the next obstacle's record is requested before the current one is evaluated
            float4 q0 = fetch(0), q1 = n_caps > 0 ? fetch(1) : make_float4(0.f, 0.f, 0.f, 0.f);
            int slot = n_caps > 0 ? 2 : 1;
#pragma unroll 1
            for (int k = 0; k < n_obst; k++) {
                const bool is_cap = k < n_caps;
                const float4 c0 = q0, c1 = q1;
                if (k + 1 < n_obst) {
                    q0 = fetch(slot);
                    if (k + 1 < n_caps) q1 = fetch(slot + 1);
                    slot += (k + 1 < n_caps) ? 2 : 1;
                }
                int hit3;
                bool view;
                cull_pair_rec<T>(p, prel, Rf, c0, c1, is_cap, hit3, view);
                bool hit = hit3 == 1;
                if (hit3 == 2) {      // within 2 mm of the collision threshold (or NaN): decided in T like the other layouts
                    T pos[3], o7[7];
                    pos_of(pos);
                    const T *g = is_cap ? p.capsules + (int64_t)(k * 7) * N + i : p.spheres + (int64_t)((k - n_caps) * 4) * N + i;
                    const int n_words = is_cap ? 7 : 4;
#pragma unroll
                    for (int c = 0; c < 7; c++) o7[c] = c < n_words ? g[(int64_t)c * N] : T(0);
                    hit = obstacle_body_hit<T>(p, pos, o7, is_cap);
                }
                info |= view ? (1u << k) : 0u;
                info |= hit ? (1u << 16) : 0u;
            }
        } else if (n_obst > 0) {
            // ---- coordinates too large for float records: everything in T (obstacle_pair)
            T Rm[9], pos[3];
            rzyx<T>(trig[0], trig[1], trig[2], trig[3], trig[4], trig[5], Rm);
            pos_of(pos);
#pragma unroll 1
            for (int k = 0; k < n_obst; k++) {
                const bool is_cap = k < n_caps;
                T o7[7];
                const T *g = is_cap ? p.capsules + (int64_t)(k * 7) * N + i : p.spheres + (int64_t)((k - n_caps) * 4) * N + i;
                const int n_words = is_cap ? 7 : 4;
#pragma unroll
                for (int c = 0; c < 7; c++) o7[c] = c < n_words ? g[(int64_t)c * N] : T(0);
                bool hit, view;
                obstacle_pair<T, false>(p, pos, Rm, o7, is_cap, nullptr, hit, view);
                info |= view ? (1u << k) : 0u;
                info |= hit ? (1u << 16) : 0u;
            }
        }
        const CullLate<T> L = late();
        // a non-finite pose poisons the rays like the reference's NaN propagation: such envs go through the ray launch
        // (without obstacles every ray reads max_dist whatever the pose, docking3d.py:441)
        poison_free = L.poison == T(0);
        listed = (info & 0xffffu) != 0u || (n_obst > 0 && !poison_free);
        // ---- everything that does not depend on the rays is final for EVERY env: done, condition bits, counters.
        //      Stores of all lanes of the warp -> full sectors (a listed env only lacks its obstacle-avoidance term: its
        //      reward word is provisional here and rewritten by the ray launch together with the running return)
        const int32_t t_steps = L.t_steps;
        cond = done_conditions<T>(p, (uint32_t)L.cond012, t_steps, (info >> 16) != 0u);
        const bool done = cond != 0;
        const int32_t t_new = t_steps + 1;
        const T r_oa = p.sum_beta_oa / p.sum_beta_oa - T(1);      // docking3d.py:792 with every ray at max_dist: exactly 0
        const T reward = step_reward<T>(p, L.A, L.B, L.r7, L.lp_d, r_oa, cond);
        p.reward[i] = reward;
        p.done[i] = done ? 1 : 0;
        if (p.cond_bits) p.cond_bits[i] = (uint8_t)cond;
        if (p.delta_d_out) p.delta_d_out[i] = L.delta_d;
        if (done && p.ep_len_out) p.ep_len_out[i] = t_new;
        if (listed) {
            p.t_steps[i] = t_new;        // the ray launch reads it back (and zeroes it if it re-initialises the env)
        } else {
            // ---- nothing in view: every ray reads max_dist (sensor.py:113-117) -> pooled cells all ones, r_oa = 0
            float *cells = p.obs + i * p.n_obs + 16;
            if ((p.n_obs & 3) == 0 && (p.n_rr & 3) == 0) {
                float4 *c4 = reinterpret_cast<float4 *>(cells);
                for (int c = 0; c < (p.n_rr >> 2); c++) c4[c] = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
            } else {
                for (int c = 0; c < p.n_rr; c++) cells[c] = 1.0f;
            }
            const T ep_ret = L.ep_return + reward;
            if (done) {
                if (p.ep_return_out) p.ep_return_out[i] = ep_ret;
                bs.done = true;
                bs.cond = cond;
                bs.length = t_new;
                bs.ep_return = (double)ep_ret;
                bs.delta_d = (double)L.delta_d;
                bs.nan = reward != reward;      // episodes that ended on a NaN reward
                ended = true;
            }
            if (!(done && p.auto_reset)) {
                p.ep_return[i] = ep_ret;
                p.t_steps[i] = t_new;
            }
        }
    }
    // ---- warp-aggregated appends to the work lists
    {
        int cls = 0;
        if (listed) {
            const int pc = __popc(info & 0xffffu);
            cls = (p.tpe_rays && pc >= 1 && pc <= 2 && poison_free) ? pc : 3;
        }
        const unsigned long long entry = (unsigned long long)(uint32_t)(i - p.env_begin) | ((unsigned long long)(info & 0xffffu) << 32) |
                                         ((unsigned long long)cond << 48);
        // the four appends of a warp (three view classes, ended list) reserve their slots with ONE atomic instruction: lane c
        // adds the warp's count to counter c, so the four round trips overlap (as four dependent atomics they were 15 % of
        // this launch's stall samples)
        const unsigned m1 = __ballot_sync(0xffffffffu, cls == 1), m2 = __ballot_sync(0xffffffffu, cls == 2);
        const unsigned m3 = __ballot_sync(0xffffffffu, cls == 3), m4 = __ballot_sync(0xffffffffu, ended);
        const unsigned mine = lane == 0 ? m1 : (lane == 1 ? m2 : (lane == 2 ? m3 : (lane == 3 ? m4 : 0u)));
        unsigned base = 0;
        if (mine) base = atomicAdd(&p.view_count[lane], (unsigned)__popc(mine));
        const unsigned b1 = __shfl_sync(0xffffffffu, base, 0), b2 = __shfl_sync(0xffffffffu, base, 1);
        const unsigned b3 = __shfl_sync(0xffffffffu, base, 2), b4 = __shfl_sync(0xffffffffu, base, 3);
        const unsigned below = (1u << lane) - 1u;
        if (cls == 1) p.view_list[b1 + __popc(m1 & below)] = entry;
        else if (cls == 2) p.view_list[p.n_envs + b2 + __popc(m2 & below)] = entry;
        else if (cls == 3) p.view_list[2 * p.n_envs + b3 + __popc(m3 & below)] = entry;
        if (ended) p.ended_list[b4 + __popc(m4 & below)] = (uint32_t)(i - p.env_begin);
    }
    bs.flush_direct(p.stats, threadIdx.x == 0 ? (int)min((int64_t)CTA, p.env_end - i0) : 0);
    return listed;
}

// ---- the dynamics launch (with the cull + finish code in it when FUSE)
// FUSE: the env's float obstacle records travel global -> shared by cp.async at the very start ([slot][thread], dynamic
// shared memory: 16 n_obsf bytes per thread) and wait there through the integration; the record never leaves the registers
// unless the env ends up on a view list.  Against the separate cull launch this saves the record's round trip through HBM
// (128 B written for every env, 128 B read back) and every memory wait of the obstacle loop.
template <typename T, int VEH, int NU, bool CUR, bool SPM, bool FIN, bool FUSE>
__global__ void
#ifdef DOCKAUV_DYN_MAXNREG
__maxnreg__(DOCKAUV_DYN_MAXNREG)
#else
__launch_bounds__(kDynThreads, DOCKAUV_MINB_A)
#endif
dynamics_kernel(const __grid_constant__ KParams<T> p) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];      // FUSE: float4 [n_obsf][kDynThreads]
    __shared__ __align__(8) T dyn_park[kParkWords * kDynThreads];
    const int64_t i0 = p.env_begin + (int64_t)blockIdx.x * kDynThreads;
    const int64_t i = i0 + threadIdx.x;
    const bool active = i < p.env_end;
    prefetch_dynamics_inputs<T, NU>(p, i, threadIdx.x & 31, true);
    T w[kRecWords], pos[3];
    if (!FUSE) {
        // the cull launch appends to the lists: this launch empties them
        if (blockIdx.x == 0 && threadIdx.x < kListCounters) p.view_count[threadIdx.x] = 0u;
        if (!active) return;
        dynamics_env<T, VEH, NU, CUR, SPM, FIN, FIN>(p, i, threadIdx.x == 0 ? (int)min((int64_t)kDynThreads, p.env_end - i0) : 0, dyn_park, w, pos,
                                                     []() {});
        if (!FIN) RecIO<T>::store(p.rec + i * kRecWords, w);
        return;
    }
    const float4 *s_obs = reinterpret_cast<const float4 *>(dyn_smem) + threadIdx.x;
    if (active) {
        dynamics_env<T, VEH, NU, CUR, SPM, false, true>(p, i, 0, dyn_park, w, pos, [&]() {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(s_obs);
            for (int sl = 0; sl < p.n_obsf; sl++)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + (unsigned)(sl * kDynThreads * 16)),
                             "l"(p.obsf + (int64_t)sl * p.n_envs + i) : "memory");
        });      // (the cp.async wait after the integration covers the records)
    }
    const bool listed = cull_finish_env<T, kDynThreads>(
        p, i, i0, active, w + REC_TRIG, w + REC_PREL, [&](int sl) { return s_obs[sl * kDynThreads]; },
        [&]() {
            CullLate<T> L;
            L.A = w[REC_A]; L.B = w[REC_B]; L.r7 = w[REC_R7]; L.lp_d = w[REC_LPD]; L.delta_d = w[REC_DD];
            L.cond012 = w[REC_COND]; L.poison = w[REC_POISON];
            L.ep_return = dyn_park[6 * kDynThreads + threadIdx.x];
            L.t_steps = reinterpret_cast<const int32_t *>(dyn_park + 7 * kDynThreads + threadIdx.x)[0];
            return L;
        },
        [&](T q[3]) { q[0] = pos[0]; q[1] = pos[1]; q[2] = pos[2]; });
    if (listed) RecIO<T>::store(p.rec + i * kRecWords, w);      // the ray launch reads it
}

// ---- the cull + finish launch on its own (records from KParams::rec; used when the dynamics launch cannot take it in)
constexpr int kCullThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kCullThreads, DOCKAUV_MINB_CULL) cull_finish_kernel(const __grid_constant__ KParams<T> p) {
    // the words needed only after the obstacle loop (reward terms, condition / poison words, running return, step counter)
    // travel global -> shared by cp.async into the thread's own slots, [slot][thread]: no register, no wait up front.  As
    // plain loads these values were spilled to the stack under the 64-register cap, the spill stores waited for the loads
    // (13 % of this launch's stall samples) and the reloads after the loop for lines that had left L1 again (another 15 %)
    __shared__ __align__(16) unsigned char cull_park[4 * kCullThreads * 16];
    const int64_t N = p.n_envs;
    const int64_t i0 = p.env_begin + (int64_t)blockIdx.x * kCullThreads;
    const int64_t i = i0 + threadIdx.x;
    const bool active = i < p.env_end;
    const T *rec = p.rec + i * kRecWords;
    T w[10];
#pragma unroll
    for (int c = 0; c < 10; c++) w[c] = T(0);
    if (active) {
#if DOCKAUV_CULL_PREFETCH
        if (!p.cull_exact)
            for (int sl = 0; sl < p.n_obsf; sl++) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.obsf + (int64_t)sl * N + i));
#endif
        RecIO<T>::template load<0, 5>(rec, w);       // trig, prel, A
        const unsigned sa = (unsigned)__cvta_generic_to_shared(cull_park) + (unsigned)threadIdx.x * 16u;
        constexpr unsigned kPlane = kCullThreads * 16u;
        if (sizeof(T) == 8) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(rec + 10) : "memory");              // B, r7
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + kPlane), "l"(rec + 12) : "memory");     // lp_d, delta_d
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + 2 * kPlane), "l"(rec + 14) : "memory"); // cond, poison
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa + 3 * kPlane), "l"(p.ep_return + i) : "memory");
        } else {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(rec + 8) : "memory");               // prel_z, A, B, r7
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + kPlane), "l"(rec + 12) : "memory");     // lp_d, delta_d, cond, poison
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa + 3 * kPlane), "l"(p.ep_return + i) : "memory");
        }
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa + 3 * kPlane + 8u), "l"(p.t_steps + i) : "memory");
    }
    const T A = w[REC_A];
    const float4 *ob = p.obsf + i;
    cull_finish_env<T, kCullThreads>(
        p, i, i0, active, w + REC_TRIG, w + REC_PREL, [&](int sl) { return ob[(int64_t)sl * N]; },
        [&]() {
            CullLate<T> L;
            cp_async_wait_all();      // this thread's own slots: no barrier
            const unsigned char *sp = cull_park + threadIdx.x * 16;
            constexpr int kPlane = kCullThreads * 16;
            if (sizeof(T) == 8) {
                const double2 a = *reinterpret_cast<const double2 *>(sp), b = *reinterpret_cast<const double2 *>(sp + kPlane);
                const double2 c = *reinterpret_cast<const double2 *>(sp + 2 * kPlane);
                L.B = (T)a.x; L.r7 = (T)a.y; L.lp_d = (T)b.x; L.delta_d = (T)b.y; L.cond012 = (T)c.x; L.poison = (T)c.y;
                L.ep_return = (T)*reinterpret_cast<const double *>(sp + 3 * kPlane);
            } else {
                const float4 a = *reinterpret_cast<const float4 *>(sp), b = *reinterpret_cast<const float4 *>(sp + kPlane);
                L.B = (T)a.z; L.r7 = (T)a.w; L.lp_d = (T)b.x; L.delta_d = (T)b.y; L.cond012 = (T)b.z; L.poison = (T)b.w;
                L.ep_return = (T)*reinterpret_cast<const float *>(sp + 3 * kPlane);
            }
            L.t_steps = *reinterpret_cast<const int32_t *>(sp + 3 * kPlane + 8);
            L.A = A;
            return L;
        },
        [&](T q[3]) {
#pragma unroll
            for (int c = 0; c < 3; c++) q[c] = p.state[(int64_t)c * N + i];
        });
}

// ------------------------------------------------------------------------------------------------------- 3. rays + finish
// Shared memory of one warp of the ray launch (T words):
//   stage[2][kStageWords]   the data of two list entries (double buffer, filled by cp.async one entry ahead):
//                           0..15 the env's record, 16..18 post-step position, 19 running return, 20 step counter (int),
//                           24 + 8 k .. obstacle k as stored (7 or 4 words), k < 16
//   pre[16][kPreStride]     ray-test records of the in-view obstacles
//   ray[ray_stride]         clamped ray distances (pooling scratch)
//   lane[RayLaneShared]     per-lane ray constants
// Nothing of the software pipeline lives in registers (round 1 kept the next entry's 24 words there), and neither do the
// per-lane ray constants: 118 -> 80 registers, six CTAs of four warps per SM instead of four.
constexpr int kStageObst = 24, kStageWords = kStageObst + 16 * 8;
template <typename T, int RPL>
struct RaysSmem {
    int warp_words, ray_stride, pre_off, ray_off, lane_off;
    __host__ __device__ RaysSmem(int n_rays) {
        ray_stride = (n_rays + 2) & ~1;
        pre_off = 2 * kStageWords;
        ray_off = pre_off + 16 * kPreStride;
        lane_off = ray_off + ray_stride;
        warp_words = (lane_off + RayLaneShared<T, RPL>::words() + 1) & ~1;
    }
};

#ifndef DOCKAUV_RAY_WARPS
#define DOCKAUV_RAY_WARPS 4
#endif
constexpr int kRayWarps = DOCKAUV_RAY_WARPS;     // warps per CTA of the ray launch


// the loop of ONE warp over the class-3 list: warp w_global of n_warps (called by all 32 lanes; smem_raw = the CTA's
// dynamic shared memory, kRayWarps * RaysSmem::warp_words words)
template <typename T, int RPL>
__device__ __forceinline__ void rays_warp_loop(const KParams<T> &p, unsigned char *smem_raw, unsigned w_global, unsigned n_warps) {
    const RaysSmem<T, RPL> L(p.n_rays);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T *s_warp = reinterpret_cast<T *>(smem_raw) + warp * L.warp_words;
    T *s_pre = s_warp + L.pre_off;
    T *s_ray = s_warp + L.ray_off;
    const int64_t N = p.n_envs;
    const unsigned count = p.view_count[2];
    const unsigned long long *list = p.view_list + 2 * p.n_envs;
    unsigned idx = w_global;
    if (idx >= count) return;

    const int n_caps = p.n_caps, n_sph = p.n_sph, n_obst = n_caps + n_sph;
    RayLaneShared<T, RPL> rl;
    rl.init(p, lane, s_ray, s_warp + L.lane_off);

    // software pipeline over the list: the data of entry n + 1 is on its way into the other stage buffer while entry n is
    // cast.  What a lane fetches: lane c < 20 one word (record, position, running return) through a per-lane pointer and
    // stride, lane 20 the step counter, lane k < n_obst with bit k of the mask obstacle k.
    const bool lane_is_cap = lane < n_caps;
    const T *obst_row = lane_is_cap ? p.capsules + (int64_t)(lane * 7) * N : p.spheres + (int64_t)((lane - n_caps) * 4) * N;
    const int obst_words = lane_is_cap ? 7 : 4;
    const T *word_src = p.rec + lane;            // lanes 0..15: record word `lane`, stride 16
    int64_t word_stride = kRecWords;
    if (lane >= 16) {
        word_stride = 1;
        word_src = lane < 19 ? p.state + (int64_t)(lane - 16) * N : p.ep_return;
    }
    // shared-window addresses of this lane's slots in stage buffer 0 (buffer 1: + kStageWords words)
    const unsigned sa_word = (unsigned)__cvta_generic_to_shared(s_warp + lane);
    const unsigned sa_obst = (unsigned)__cvta_generic_to_shared(s_warp + kStageObst + 8 * (lane & 15));
    auto fetch = [&](uint64_t entry, int b) {
        const int64_t e = p.env_begin + (int64_t)(uint32_t)entry;
        const unsigned mask = (unsigned)(entry >> 32) & 0xffffu;
        const unsigned boff = b ? (unsigned)(kStageWords * sizeof(T)) : 0u;
        if (lane < 20) cp_async_word<T>(sa_word + boff, word_src + e * word_stride);
        if (lane == 20) cp_async_word<int32_t>(sa_word + boff, p.t_steps + e);
        if (lane < n_obst && ((mask >> lane) & 1u)) {
            const T *g = obst_row + e;
#pragma unroll
            for (int c = 0; c < 7; c++)
                if (c < obst_words) cp_async_word<T>(sa_obst + boff + c * (unsigned)sizeof(T), g + (int64_t)c * N);
        }
    };
    uint64_t cur = list[idx];
    uint64_t nxt = (idx + n_warps < count) ? list[idx + n_warps] : 0;
    fetch(cur, 0);
    double stat_acc = 0.0;       // lane k < DOCKAUV_STAT_ENV_STEPS accumulates statistic k of the episodes this warp ended
    int buf = 0;

    for (; idx < count; idx += n_warps, buf ^= 1) {
        const int64_t ie = p.env_begin + (int64_t)(uint32_t)cur;
        const unsigned mask = (unsigned)(cur >> 32) & 0xffffu;
        const uint32_t cond = (uint32_t)(cur >> 48) & 31u;
        const T *s_ent = s_warp + buf * kStageWords;
        // ---- the staged data of this entry is complete; start the next fetch into the other buffer (last read one
        //      iteration ago, before that iteration's closing __syncwarp)
        cp_async_wait_all();
        __syncwarp();
        const uint64_t nn = (idx + 2 * n_warps < count) ? list[idx + 2 * n_warps] : 0;
        if (idx + n_warps < count) fetch(nxt, buf ^ 1);
        const T pos[3] = {s_ent[16], s_ent[17], s_ent[18]};
        if (lane < n_obst && ((mask >> lane) & 1u)) {
            T ob[7];
#pragma unroll
            for (int c = 0; c < 7; c++) ob[c] = c < obst_words ? s_ent[kStageObst + 8 * lane + c] : T(0);
            obstacle_ray_record<T>(pos, ob, lane_is_cap, s_pre + lane * kPreStride);
        }
        T R[9];
        rzyx<T>(s_ent[0], s_ent[1], s_ent[2], s_ent[3], s_ent[4], s_ent[5], R);
        const T poison = s_ent[REC_POISON];
        __syncwarp();

        // ---- cast rays against the in-view obstacles, pooled cells to the observation row
        const T oa_dot = radar_env<T, RPL, false, RayLaneShared<T, RPL>>(p, rl, R, poison, mask, s_pre, s_ray, lane, ie);

        // ---- what the cull code left open: the reward with its obstacle-avoidance term and the running return
        //      (every lane computes the same values from the staged record, lane 0 stores)
        const T r_oa = p.sum_beta_oa / oa_dot - T(1);      // docking3d.py:792 (IEEE division: the other layouts' bits)
        const T reward = step_reward<T>(p, s_ent[REC_A], s_ent[REC_B], s_ent[REC_R7], s_ent[REC_LPD], r_oa, cond);
        const T ep_ret = s_ent[19] + reward;
        const bool done = cond != 0;
        if (lane == 0) {
            p.reward[ie] = reward;
            if (done && p.ep_return_out) p.ep_return_out[ie] = ep_ret;
            if (!(done && p.auto_reset)) p.ep_return[ie] = ep_ret;
        }
        if (done) {       // warp-uniform, ~1 % of the listed envs
            const int32_t t_new = reinterpret_cast<const int32_t *>(s_ent + 20)[0];     // already incremented by the cull code
            double mine = 0.0;
            if (lane == DOCKAUV_STAT_EPISODES) mine = 1.0;
            else if (lane == DOCKAUV_STAT_SUM_RETURN) mine = (double)ep_ret;
            else if (lane == DOCKAUV_STAT_SUM_LENGTH) mine = (double)t_new;
            else if (lane >= DOCKAUV_STAT_COND0 && lane < DOCKAUV_STAT_COND0 + 5) mine = ((cond >> (lane - DOCKAUV_STAT_COND0)) & 1u) ? 1.0 : 0.0;
            else if (lane == DOCKAUV_STAT_SUM_FINAL_DELTA_D) mine = (double)s_ent[REC_DD];
            else if (lane == DOCKAUV_STAT_NAN_ENVS) mine = (reward != reward) ? 1.0 : 0.0;
            stat_acc += mine;
            if (lane == 0) p.ended_list[atomicAdd(&p.view_count[3], 1u)] = (uint32_t)cur;
        }
        __syncwarp();
        cur = nxt;
        nxt = nn;
    }
    if (lane < DOCKAUV_STAT_ENV_STEPS && stat_acc != 0.0)
        atomicAdd(&p.stats[(blockIdx.x & (DOCKAUV_STAT_COPIES - 1)) * DOCKAUV_N_STATS + lane], stat_acc);
}

// the warp-per-env ray launch on its own (radars the thread mapping does not cover: every listed env is class 3)
template <typename T, int RPL>
__global__ void __launch_bounds__(kRayWarps * 32, DOCKAUV_MINB_RAYS * 4 / kRayWarps) rays_finish_kernel(const __grid_constant__ KParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    rays_warp_loop<T, RPL>(p, smem_raw, blockIdx.x * kRayWarps + (threadIdx.x >> 5), gridDim.x * kRayWarps);
}

// ------------------------------------------------------------------------------------------------------- 3b. rays, thread per env
// The envs with exactly one or two obstacles in view -- nine in ten of the listed ones -- with LANES = ENVS: every lane
// walks over the rays of its own env (SPLIT lanes share one env, each takes a contiguous run of pooled cells), the
// ray-test records of its obstacles stay in registers.  Against the warp-per-env mapping, everything that is per-env
// scalar work there (list entry, record fetch, Rzyx, obstacle records, reward: ~450 of its ~600 warp instructions per env,
// done by 32 lanes redundantly) is amortised over the envs of a warp here.  The ray tests are straight-line code
// (cast_capsule_bf; spheres take the same path as degenerate capsules, so lanes never diverge on the obstacle type) and
// the four rays of a pooled cell are independent chains the compiler interleaves.  The body-frame ray table is walked in
// pooled-cell order (shared memory), so a cell's maximum is complete after four rays and goes straight into the
// observation row.  One persistent launch covers both classes: tiles of class 2 (the longer ones) first, then class 1.
constexpr int kTpeThreads = kRayWarps * 32;      // one CTA shape for both ray mappings of the launch
constexpr int kTpeMaxCells = 64;      // pooled cells the shared ray table holds

template <typename T, int NOB, int SPLIT>
__device__ __forceinline__ void rays_thread_tile(const KParams<T> &p, const T *s_tab, unsigned tile, unsigned count) {
    const int64_t N = p.n_envs;
    const int n_rr = p.n_rr;
    const unsigned tg = tile * kTpeThreads + threadIdx.x;
    const unsigned eidx = tg / SPLIT;
    const int part = (int)(tg % SPLIT);
    const bool valid = eidx < count;
    const unsigned long long entry = valid ? p.view_list[(int64_t)(NOB - 1) * N + eidx] : 0ull;
    const int64_t ie = p.env_begin + (int64_t)(uint32_t)entry;
    unsigned mask = (unsigned)(entry >> 32) & 0xffffu;
    const uint32_t cond = (uint32_t)(entry >> 48) & 31u;
    const T *rec = p.rec + ie * kRecWords;
    // ---- pose and the ray-test records of the in-view obstacles (registers)
    T R[9], w[NOB][11];
    bool sph[NOB];
    {
        T trig[6], pos[3];
        RecIO<T>::template load<0, 3>(rec, trig);
#pragma unroll
        for (int c = 0; c < 3; c++) pos[c] = p.state[(int64_t)c * N + ie];
        rzyx<T>(trig[0], trig[1], trig[2], trig[3], trig[4], trig[5], R);
#pragma unroll
        for (int o = 0; o < NOB; o++) {
            const int k = mask ? __ffs(mask) - 1 : 0;
            mask &= mask - 1;
            const bool is_cap = k < p.n_caps;
            sph[o] = !is_cap;
            const T *g = is_cap ? p.capsules + (int64_t)(k * 7) * N + ie : p.spheres + (int64_t)((k - p.n_caps) * 4) * N + ie;
            T ob[7];
#pragma unroll
            for (int c = 0; c < 7; c++) ob[c] = (c < 4 || is_cap) ? g[(int64_t)c * N] : T(0);
#pragma unroll
            for (int c = 0; c < 11; c++) w[o][c] = T(0);
            obstacle_ray_record<T>(pos, ob, is_cap, w[o]);
            if (!is_cap) sphere_as_capsule<T>(w[o]);
#if DOCKAUV_TPE_BODY
            // the record's two vectors into the BODY frame (R^T v): a ray direction then is its table entry as it stands,
            // (R b) . v = b . (R^T v) -- 18 products per obstacle instead of 9 per ray
#pragma unroll
            for (int v = 0; v < 2; v++) {
                const T x = w[o][3 * v], y = w[o][3 * v + 1], z = w[o][3 * v + 2];
#pragma unroll
                for (int c = 0; c < 3; c++) w[o][3 * v + c] = R[c] * x + R[3 + c] * y + R[6 + c] * z;
            }
#endif
        }
    }
    T h_min[NOB];      // see cast_capsule_bf
#pragma unroll
    for (int o = 0; o < NOB; o++) h_min[o] = sph[o] ? -Mth<T>::min_normal() : T(0);
    // ---- this lane's run of pooled cells: multiples of four, so the row is written in 16-byte pieces
    const int cpp = (((n_rr + SPLIT - 1) / SPLIT) + 3) & ~3;
    const int c_begin = part * cpp;
    const T dmax = p.radar_max_dist, inv_dmax = T(1) / dmax;
    float *orow = p.obs + ie * p.n_obs + 16;
    const bool vec_row = (p.n_obs & 3) == 0 && (n_rr & 3) == 0;
    T oa_part = T(0);
    float o4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    // (every lane runs all cpp trips: cast_capsule_bf contains a warp vote; cells beyond the grid are computed and dropped)
#pragma unroll 1
    for (int cc = 0; cc < cpp; cc++) {
        const bool live = c_begin + cc < n_rr;
        const int cell = live ? c_begin + cc : 0;
        T dq[4], bwq[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const T *tb = s_tab + 4 * (4 * cell + q);
            const T b0 = tb[0], b1 = tb[1], b2 = tb[2];
            bwq[q] = tb[3];
            T rd[3];
#if DOCKAUV_TPE_BODY
            rd[0] = b0; rd[1] = b1; rd[2] = b2;
#else
#pragma unroll
            for (int c = 0; c < 3; c++) rd[c] = R[3 * c] * b0 + R[3 * c + 1] * b1 + R[3 * c + 2] * b2;
#endif
            T best = Mth<T>::inf();
#pragma unroll
            for (int o = 0; o < NOB; o++) {
                const T ba[3] = {w[o][0], w[o][1], w[o][2]}, oa[3] = {w[o][3], w[o][4], w[o][5]};
                best = cast_capsule_bf<T>(rd, ba, oa, w[o][6], w[o][7], w[o][8], w[o][9], w[o][10], h_min[o], best);
            }
            dq[q] = best > dmax ? dmax : best;        // clamp (sensor.py:117)
        }
        T mx = T(0);                                  // block_reduce pads with cval = 0 (sensor.py:131-137)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (bwq[q] < T(0) || !live) continue;     // padding slot of the ray table: no ray
            // obstacle-avoidance partial sum (docking3d.py:767-792), 2x2 max-pool
            const T x = dq[q] * inv_dmax;
            const T qq = x * x;
            const T mq = !(qq <= T(0.001)) ? qq : T(0.001);     // np.maximum, NaN propagates
            oa_part += mq * bwq[q];
            mx = !(dq[q] <= mx) ? dq[q] : mx;         // np.max, NaN propagates
        }
        T o = mx * inv_dmax;                          // clip(d / max_dist, 0, 1), docking3d.py:487
        o = o > T(1) ? T(1) : o;
        if (vec_row) {
            o4[cell & 3] = (float)o;
            if ((cell & 3) == 3 && valid && live) reinterpret_cast<float4 *>(orow)[cell >> 2] = make_float4(o4[0], o4[1], o4[2], o4[3]);
        } else if (valid && live) {
            orow[cell] = (float)o;
        }
    }
    // ---- the env's obstacle-avoidance sum: add up the SPLIT lanes of the env
    T oa_dot = oa_part;
#pragma unroll
    for (int s = 1; s < SPLIT; s <<= 1) oa_dot += __shfl_xor_sync(0xffffffffu, oa_dot, s);
    if (valid && part == 0) {
        // ---- what the cull code left open: the reward with its obstacle-avoidance term and the running return
        T wf[6];
        RecIO<T>::template load<4, 3>(rec, wf);       // prel_z, A, B, r7, lp_d, delta_d
        const T ep_before = p.ep_return[ie];
        const T r_oa = p.sum_beta_oa / oa_dot - T(1);      // docking3d.py:792
        const T reward = step_reward<T>(p, wf[1], wf[2], wf[3], wf[4], r_oa, cond);
        const T ep_ret = ep_before + reward;
        const bool done = cond != 0;
        p.reward[ie] = reward;
        if (done && p.ep_return_out) p.ep_return_out[ie] = ep_ret;
        if (!(done && p.auto_reset)) p.ep_return[ie] = ep_ret;
        if (done) {      // ~1 % of the listed envs
            WarpStats bs;
            bs.done = true;
            bs.cond = cond;
            bs.length = p.t_steps[ie];                // already incremented by the cull code
            bs.ep_return = (double)ep_ret;
            bs.delta_d = (double)wf[5];
            bs.nan = reward != reward;
            bs.flush_direct(p.stats, 0);
            p.ended_list[atomicAdd(&p.view_count[3], 1u)] = (uint32_t)entry;
        }
    }
}

// One launch for all three classes: the first `warp_ctas` CTAs run the warp-per-env loop over class 3 (few entries, each a
// long latency chain: started first, they run alongside the tiles instead of as a launch of their own that leaves the GPU
// 60 % idle for 11 us), the others are the persistent tile loop of classes 2 and 1.
template <typename T, int SPLIT, int RPL>
__global__ void __launch_bounds__(kTpeThreads, DOCKAUV_MINB_TPE) rays_thread_kernel(const __grid_constant__ KParams<T> p, unsigned warp_ctas) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(16) T s_tab[kTpeMaxCells * 4 * 4];      // slot = 4 * cell + q: rb[3], bw; bw < 0 marks zero padding
    static_assert(kTpeThreads == kRayWarps * 32, "the two ray mappings share one CTA shape");
    grid_dependency_wait();
    if (blockIdx.x < warp_ctas) {
        rays_warp_loop<T, RPL>(p, smem_raw, blockIdx.x * kRayWarps + (threadIdx.x >> 5), warp_ctas * kRayWarps);
        return;
    }
    const unsigned bid = blockIdx.x - warp_ctas, n_ctas = gridDim.x - warp_ctas;
    const unsigned c1 = p.view_count[0], c2 = p.view_count[1];
    constexpr unsigned per_tile = kTpeThreads / SPLIT;
    const unsigned t2 = (c2 + per_tile - 1) / per_tile, t1 = (c1 + per_tile - 1) / per_tile;
    if (bid >= t1 + t2) return;
    const int n_rr = p.n_rr, n_r = p.n_rays;
    for (int t = threadIdx.x; t < 4 * n_rr; t += kTpeThreads) {
        const int cell = t >> 2, q = t & 3;
        const int pr = cell / p.n_hr, pcol = cell - pr * p.n_hr;
        const int rv = 2 * pr + (q >> 1), rh = 2 * pcol + (q & 1);
        const bool ok = rv < p.n_vert && rh < p.n_horiz;
        const int ir = ok ? rv * p.n_horiz + rh : 0;
#pragma unroll
        for (int c = 0; c < 3; c++) s_tab[4 * t + c] = ok ? p.ray_tab[c * n_r + ir] : T(0);
        s_tab[4 * t + 3] = ok ? p.ray_tab[3 * n_r + ir] : T(-1);
    }
    __syncthreads();
    for (unsigned tile = bid; tile < t1 + t2; tile += n_ctas) {
        if (tile < t2) rays_thread_tile<T, 2, SPLIT>(p, s_tab, tile, c2);
        else rays_thread_tile<T, 1, SPLIT>(p, s_tab, tile - t2, c1);
    }
}

// ------------------------------------------------------------------------------------------------------- 4. episode end
// The ENDED envs of the step (compact list, ~1 % of the batch): the last observation is kept as terminal_observation, the
// all-zero reset observation is handed back (docking3d.py:269,322) and the env is re-initialised.  One CTA of eight warps
// per 32 list entries: lane = env, warp = reset role (reset_envs_cta) and 16-byte piece of the observation row.
// Measured alternatives (1M-env step): one WARP per ended env (round 1, and the first version of this pipeline, inside the
// ray launch): ~4000 warp instructions per reset at 1..8 active lanes, 28 M per half batch, a third of the ray launch; one
// THREAD per ended env: 40x leaner but a 5500-instruction dependent chain, 70 us for a launch that occupies 5 % of the GPU;
// eight LANES per env: the roles serialise inside the warp, same 70 us.
template <typename T>
__global__ void __launch_bounds__(kResetCta) episode_end_kernel(const __grid_constant__ KParams<T> p) {
    grid_dependency_wait();
    const unsigned n_ended = p.view_count[3];
    const int n_obs = p.n_obs;
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    // every thread of a CTA runs the same number of trips (reset_envs_cta has a barrier)
    const unsigned trips = (n_ended + gridDim.x * 32 - 1) / (gridDim.x * 32);
    unsigned k = blockIdx.x * 32 + lane;
    for (unsigned trip = 0; trip < trips; trip++, k += gridDim.x * 32) {
        const bool valid = k < n_ended;
        int64_t ie = 0;
        if (valid) {
            ie = p.env_begin + (int64_t)p.ended_list[k];
            float *row = p.obs + ie * n_obs;
            float *trow = p.terminal_obs ? p.terminal_obs + ie * n_obs : nullptr;
            if ((n_obs & 3) == 0) {
                float4 *r4 = reinterpret_cast<float4 *>(row), *t4 = reinterpret_cast<float4 *>(trow);
                for (int c = role; c < (n_obs >> 2); c += kResetRoles) {
                    if (trow) t4[c] = r4[c];
                    if (p.auto_reset) r4[c] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
            } else {
                for (int c = role; c < n_obs; c += kResetRoles) {
                    if (trow) trow[c] = row[c];
                    if (p.auto_reset) row[c] = 0.0f;
                }
            }
        }
        if (p.auto_reset) reset_envs_cta<T>(p, ie, valid);
    }
    // ---- steps whose dynamics launch appends to the lists itself (cull code fused in) cannot have that launch empty
    //      them: the last CTA to get here saves the list counters of this step and zeroes them for the next one (every CTA
    //      has read the ended count and its list entries by now).  Other steps skip this: the ticket's round trip is on
    //      the critical path of small batches (65,536-env SimpleDocking3d step: 24.3 against 22.6 us)
    if (!p.counters_zeroed_at_end) return;
    // (no fence: what must precede the ticket are this CTA's READS of the counter and of its list entries, and their values
    // have been consumed by now; a __threadfence() here waited for every reset store of the CTA: 37 % of this launch's samples)
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&p.view_count[kCounterTicket], 1u) == gridDim.x - 1) {
#pragma unroll
            for (int c = 0; c < kListCounters; c++) {
                p.view_count[kCounterLast + c] = p.view_count[c];
                p.view_count[c] = 0u;
            }
            p.view_count[kCounterTicket] = 0u;
        }
    }
}

// ------------------------------------------------------------------------------------------------------- launcher
template <typename... KArgs, typename... Args>
static cudaError_t launch_dependent(void (*kern)(KArgs...), unsigned blocks, unsigned threads, size_t smem, cudaStream_t st, Args... args) {
#if DOCKAUV_PDL
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
#else
    kern<<<blocks, threads, smem, st>>>(args...);
    return cudaGetLastError();
#endif
}

template <typename T, int VEH, int NU, bool FIN, bool FUSE>
static cudaError_t launch_dynamics(const KParams<T> &k, unsigned blocks, cudaStream_t st) {
    const bool cur = k.has_current != 0, spm = k.sparse_minv != 0;
    const int smem = FUSE ? k.n_obsf * kDynThreads * 16 : 0;
    auto go = [&](auto kern) {
        if (smem > 32 * 1024) {
            const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<blocks, kDynThreads, smem, st>>>(k);
        return cudaGetLastError();
    };
    if (cur && spm) return go(dynamics_kernel<T, VEH, NU, true, true, FIN, FUSE>);
    if (cur) return go(dynamics_kernel<T, VEH, NU, true, false, FIN, FUSE>);
    if (spm) return go(dynamics_kernel<T, VEH, NU, false, true, FIN, FUSE>);
    return go(dynamics_kernel<T, VEH, NU, false, false, FIN, FUSE>);
}

template <typename T, int VEH, int NU>
static cudaError_t launch_step_pipe(const KParams<T> &k, cudaStream_t st, cudaEvent_t *marks = nullptr, int *n_marks = nullptr) {
    if (wants_debug(k) || k.rec == nullptr) return launch_step_warp<T, VEH, NU>(k, st);   // debug outputs: fused kernel
    if (k.chunk_envs > 0 && k.chunk_envs < k.env_end - k.env_begin && marks == nullptr) {
        // chunked: the launches per chunk of envs (measured slower than the whole batch: partial waves of short launches
        // cost more than an L2-resident record saves; kept for experiments and as a test handle)
        const int64_t chunk = ((k.chunk_envs + kWarpEnvs - 1) / kWarpEnvs) * kWarpEnvs;
        for (int64_t b = k.env_begin; b < k.env_end; b += chunk) {
            KParams<T> kb = k;
            kb.env_begin = b;
            kb.env_end = b + chunk < k.env_end ? b + chunk : k.env_end;
            kb.chunk_envs = 0;
            cudaError_t e = launch_step_pipe<T, VEH, NU>(kb, st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    const int64_t n = k.env_end - k.env_begin;
    KParams<T> kc = k;
    kc.view_count = k.view_count + kCounterStride * (k.env_begin / kWarpEnvs);   // counters per concurrently stepped env range
    kc.view_list = k.view_list + k.env_begin;
    kc.ended_list = k.ended_list + k.env_begin;
    int n_mark = 0;
    auto mark = [&]() {
        if (marks) cudaEventRecord(marks[n_mark++], st);
    };
    mark();
    const bool has_obstacles = k.n_caps + k.n_sph > 0;
    const unsigned dyn_blocks = (unsigned)((n + kDynThreads - 1) / kDynThreads);
    // scenarios without obstacles are finished by the dynamics launch itself (no cull, no rays); with obstacles the cull +
    // finish code runs inside it whenever the float records of a CTA fit in shared memory
    const bool fuse = pipe_fuses_cull(k);
    kc.counters_zeroed_at_end = fuse ? 1 : 0;
    cudaError_t e = !has_obstacles ? launch_dynamics<T, VEH, NU, true, false>(kc, dyn_blocks, st)
                    : fuse       ? launch_dynamics<T, VEH, NU, false, true>(kc, dyn_blocks, st)
                                 : launch_dynamics<T, VEH, NU, false, false>(kc, dyn_blocks, st);
    if (e != cudaSuccess) return e;
    mark();
    if (has_obstacles) {
        if (!fuse) {
            cull_finish_kernel<T><<<(unsigned)((n + kCullThreads - 1) / kCullThreads), kCullThreads, 0, st>>>(kc);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            mark();
        }
        const int smem = kRayWarps * (k.n_rays <= 64 ? RaysSmem<T, 2>(k.n_rays).warp_words : RaysSmem<T, 8>(k.n_rays).warp_words) * (int)sizeof(T);
        const int64_t sms = k.sm_count > 0 ? k.sm_count : 148;
        if (k.tpe_rays) {
            // ---- rays: one launch; thread per env for the envs with one or two obstacles in view (persistent grid over tiles
            //      of 128 lanes; the counts are only known on the device), warp per env for the rest on its first CTAs
            int64_t tb = (n * DOCKAUV_TPE_SPLIT + kTpeThreads - 1) / kTpeThreads;      // tiles if every env were listed
            if (tb > sms * DOCKAUV_TPE_CTAS_PER_SM) tb = sms * DOCKAUV_TPE_CTAS_PER_SM;
            int64_t wb = (n + kRayWarps - 1) / kRayWarps;
            if (wb > sms * DOCKAUV_TPE_WARP_CTAS_PER_SM) wb = sms * DOCKAUV_TPE_WARP_CTAS_PER_SM;
            if (k.n_rays <= 64) {
                auto kern = rays_thread_kernel<T, DOCKAUV_TPE_SPLIT, 2>;
                if (smem > 40 * 1024 && (e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
                if ((e = launch_dependent(kern, (unsigned)(wb + tb), kTpeThreads, smem, st, kc, (unsigned)wb)) != cudaSuccess) return e;
            } else {
                auto kern = rays_thread_kernel<T, DOCKAUV_TPE_SPLIT, 8>;
                if (smem > 40 * 1024 && (e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
                if ((e = launch_dependent(kern, (unsigned)(wb + tb), kTpeThreads, smem, st, kc, (unsigned)wb)) != cudaSuccess) return e;
            }
        } else {
            // ---- rays: warp per env for every listed env (persistent grid)
            int64_t blocks = sms * (DOCKAUV_RAY_CTAS_PER_SM * 4 / kRayWarps);
            const int64_t most = (n + kRayWarps - 1) / kRayWarps;
            if (blocks > most) blocks = most;
            if (k.n_rays <= 64) {
                auto kern = rays_finish_kernel<T, 2>;
                if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
                kern<<<(unsigned)blocks, kRayWarps * 32, smem, st>>>(kc);
            } else {
                auto kern = rays_finish_kernel<T, 8>;
                if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
                kern<<<(unsigned)blocks, kRayWarps * 32, smem, st>>>(kc);
            }
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        mark();
    }
    {
        // the list length is only known on the device: a grid for ~3 % of the range (32 entries per CTA), grid-stride beyond
        int64_t blocks = n / 1024 + 1;
        const int64_t cap = (int64_t)(k.sm_count > 0 ? k.sm_count : 148) * 4;
        if (blocks > cap) blocks = cap;
        if ((e = launch_dependent(episode_end_kernel<T>, (unsigned)blocks, kResetCta, 0, st, kc)) != cudaSuccess) return e;
        mark();
    }
    if (n_marks) *n_marks = n_mark;
    return cudaGetLastError();
}

}  // namespace dockauv
