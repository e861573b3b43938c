// dockauv_env.cuh -- env-level pieces shared by the step kernels: reset, command filter, current, the
// non-radar part of observe / is_done / reward_step, episode statistics.
#pragma once
#include "dockauv_device.cuh"

namespace dockauv {

// ------------------------------------------------------------------------------------------- reset
// Float obstacle records of env i for the cull code (KParams::obsf, see cull_pair_rec), from the obstacles and the
// goal as they are stored (T).  Capsule k -> slots 2k, 2k+1; sphere s -> slot 2 n_caps + s.
template <typename T>
__device__ __forceinline__ void store_capsule_record(const KParams<T> &p, int64_t i, int k, const T cap[7], const T goal[3]) {
    if (p.obsf == nullptr) return;
    const double ob[7] = {(double)cap[0], (double)cap[1], (double)cap[2], (double)cap[3], (double)cap[4], (double)cap[5], (double)cap[6]};
    const double g[3] = {(double)goal[0], (double)goal[1], (double)goal[2]};
    float4 q0, q1;
    obstacle_record_f32(ob, g, true, q0, q1);
    p.obsf[(int64_t)(2 * k) * p.n_envs + i] = q0;
    p.obsf[(int64_t)(2 * k + 1) * p.n_envs + i] = q1;
}

template <typename T>
__device__ __forceinline__ void store_sphere_record(const KParams<T> &p, int64_t i, int s, const T sph[4], const T goal[3]) {
    if (p.obsf == nullptr) return;
    const double ob[7] = {(double)sph[0], (double)sph[1], (double)sph[2], (double)sph[3], 0.0, 0.0, 0.0};
    const double g[3] = {(double)goal[0], (double)goal[1], (double)goal[2]};
    float4 q0, q1;
    obstacle_record_f32(ob, g, false, q0, q1);
    p.obsf[(int64_t)(2 * p.n_caps + s) * p.n_envs + i] = q0;
}

// Rebuilds the float obstacle records of env i from the bound buffers (after the caller wrote obstacles or goals itself:
// dockauv_refresh_obstacles).
template <typename T>
__global__ void refresh_obstacles_kernel(const __grid_constant__ KParams<T> p) {
    const int64_t N = p.n_envs;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const T goal[3] = {p.goal[i], p.goal[N + i], p.goal[2 * N + i]};
    for (int k = 0; k < p.n_caps; k++) {
        T cap[7];
#pragma unroll
        for (int j = 0; j < 7; j++) cap[j] = p.capsules[(int64_t)(k * 7 + j) * N + i];
        store_capsule_record<T>(p, i, k, cap, goal);
    }
    for (int s = 0; s < p.n_sph; s++) {
        T sph[4];
#pragma unroll
        for (int j = 0; j < 4; j++) sph[j] = p.spheres[(int64_t)(s * 4 + j) * N + i];
        store_sphere_record<T>(p, i, s, sph, goal);
    }
}

// BaseDocking3d.reset (docking3d.py:222-322) + <Scenario>.generate_environment (:803-988) for env i.
// Distributions are the reference's; the random stream is Philox4x32-10 keyed by (seed, global env id,
// episode) instead of the reference's global MT19937 (DESIGN.md "reset").  Draw slots: 0 heading, 1-3
// position, 4-6 attitude, 7 goal angle, 8 goal depth, 9 pillar phase, 10-11 current direction, 12 current
// speed, 13+3s.. synthetic sphere s.
// The work of one reset is cut into kResetRoles independent ROLES so that eight warps can share it (a reset done by one
// thread is a ~5500-instruction dependent chain): 0 pose / goal / counters, 1..4 pillar k, 5 dock capsule + unused capsule
// slots + current + command, 6..7 spheres (s = role - 6, role - 4, ..).  Every role draws what it needs from the
// counter-based stream itself; roles never exchange values, so any mapping of roles to threads gives identical bits.
//   ep: the env's episode counter BEFORE this reset (the caller bumps it once all roles have read it)
constexpr int kResetRoles = 8;

template <typename T>
__device__ __forceinline__ void reset_env_role(const KParams<T> &p, int64_t i, int role, uint32_t ep) {
    const int64_t N = p.n_envs;
    const uint64_t gid = p.env_id0 + (uint64_t)i;
    const double PI = 3.141592653589793;
    // one Philox block yields two draws: the blocks a role needs are evaluated once each (role 0, the longest chain of the
    // eight, took nine block evaluations for its nine draws when every draw evaluated its own)
    auto U = [&](uint32_t idx) { return philox_uniform(p.seed, gid, ep, idx); };
    auto U2 = [&](uint32_t blk, double u[2]) { philox_uniform_pair(p.seed, gid, ep, blk, u); };
    const int scn = p.scenario;
    const bool has_goal_ring = scn >= DOCKAUV_SCN_CAPSULE;
    const bool has_dock = has_goal_ring && scn != DOCKAUV_SCN_OBSTACLES_NOCAP && p.n_caps > 0;
    const bool has_pillars = scn >= DOCKAUV_SCN_OBSTACLES;
    const int first_pillar = has_dock ? 1 : 0;
    const int n_pillars = has_pillars ? max(0, min(4, p.n_caps - first_pillar)) : 0;
    // the goal (draws 7 and 8): stored by role 0, and the origin of every float obstacle record
    double goal[3] = {0.0, 0.0, 0.0};
    double u67[2] = {0.0, 0.0}, u89[2] = {0.0, 0.0};      // draws 6..9: goal angle / depth here, yaw (role 0) and pillar phase below
    if (has_goal_ring || role == 0) U2(3, u67);
    if (has_goal_ring) {                                                       // :860-886
        U2(4, u89);
        double theta = u67[1] * 2 * PI;
        double radius = 1.0 + (double)p.safety_radius;
        double s, c;
        sincos(theta, &s, &c);
        goal[0] = c * radius;
        goal[1] = s * radius;
        goal[2] = (u89[0] - 0.5) * 4.0;
    }
    const T goal_t[3] = {(T)goal[0], (T)goal[1], (T)goal[2]};
    // obstacle rows (T) + their float records for the cull code
    auto put_capsule = [&](int k, const double cap[7]) {
        T c[7];
#pragma unroll
        for (int j = 0; j < 7; j++) {
            c[j] = (T)cap[j];
            p.capsules[(int64_t)(k * 7 + j) * N + i] = c[j];
        }
        store_capsule_record<T>(p, i, k, c, goal_t);
    };
    auto put_sphere = [&](int k, const double sph[4]) {
        T c[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            c[j] = (T)sph[j];
            p.spheres[(int64_t)(k * 4 + j) * N + i] = c[j];
        }
        store_sphere_record<T>(p, i, k, c, goal_t);
    };
    if (role == 0) {
        double u01[2], u23[2], u45[2];
        U2(0, u01);
        U2(1, u23);
        U2(2, u45);
        double heading = (u01[0] - 0.5) * PI;                                 // :814
        double r[3] = {u01[1] - 0.5, u23[0] - 0.5, u23[1] - 0.5};             // :694-696
        {
            double sg = (r[2] > 0.0) - (r[2] < 0.0);
            r[2] = fabs(r[0] + r[1]) / 3 * sg;
        }
        double sc = 15.0 / sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        double pos[3] = {r[0] * sc, r[1] * sc, r[2] * sc};
        double max_att = (double)p.max_attitude;
        double att[3] = {(u45[0] - 0.5) * 2 * (max_att * 0.7), (u45[1] - 0.5) * 2 * (max_att * 0.7),
                         (u67[0] - 0.5) * 2 * PI};                             // :699-703
        // vec_line_point(goal, top=(0,0,-2), bot=(0,0,2)) = (-gx, -gy, 0)  (shape.py:420-433)
        if (has_goal_ring) heading = (double)ssa<double>(atan2(0.0 - goal[1], 0.0 - goal[0]));
#pragma unroll
        for (int c = 0; c < 3; c++) {
            p.state[(int64_t)c * N + i] = (T)pos[c];
            p.state[(int64_t)(3 + c) * N + i] = (T)att[c];
            p.goal[(int64_t)c * N + i] = goal_t[c];
        }
#pragma unroll
        for (int c = 6; c < 12; c++) p.state[(int64_t)c * N + i] = T(0);     // auvsim.py:55-65
        p.heading_goal[i] = (T)heading;
        p.t_steps[i] = 0;
        p.ep_return[i] = T(0);
    } else if (role <= 4) {                                                    // pillars, :919-946
        const int k = role - 1;
        if (k < n_pillars) {
            double theta = (has_goal_ring ? u89[1] : U(9)) * 2 * PI;
            for (int q = 0; q < k; q++) theta += 2 * PI / 4;
            const double half = 2.0 * (double)p.max_dist_from_goal / 2.0;
            double s, c;
            sincos(theta, &s, &c);
            const double x = c * 6, y = s * 6;
            const double cap[7] = {x, y, half, x, y, -half, 1.0};
            put_capsule(first_pillar + k, cap);
        }
    } else if (role == 5) {
        if (has_dock) {
            const double cap[7] = {0, 0, 2.0, 0, 0, -2.0, 1.0};              // bot = 2*position - top, shape.py:105-108
            put_capsule(0, cap);
        }
        for (int kc = first_pillar + n_pillars; kc < p.n_caps; kc++) {         // unused slots: far away, zero radius
            const double cap[7] = {1e6, 1e6, 1e6, 1e6, 1e6, 1e6 + 1.0, 0.0};
            put_capsule(kc, cap);
        }
        double cur[5] = {0, 0, 0, 0, 0};
        if (scn == DOCKAUV_SCN_SIMPLE_CURRENT || scn == DOCKAUV_SCN_CAPSULE_CURRENT || scn == DOCKAUV_SCN_OBSTACLES_CURRENT) {
            cur[1] = (U(10) - 0.5) * 2 * (PI / 2);                             // :843-848, :903-907, :983-987
            cur[2] = (U(11) - 0.5) * 2 * PI;
            double speed = (scn == DOCKAUV_SCN_SIMPLE_CURRENT) ? U(12) * 1.0 : 0.5;
            cur[0] = 0.5;
            cur[3] = cur[4] = speed;
        }
#pragma unroll
        for (int c = 0; c < 5; c++) p.current[(int64_t)c * N + i] = (T)cur[c];
        for (int k = 0; k < p.n_u; k++) p.u_prev[(int64_t)k * N + i] = T(0);
    } else {
        const int n_synth = min(p.n_synth_sph, p.n_sph);
        for (int ks = role - 6; ks < p.n_sph; ks += 2) {
            if (ks < n_synth) {                                                 // BASELINE C4 extension: random unit spheres
                double z = 2 * U(13 + 3 * ks) - 1;
                double az = 2 * PI * U(14 + 3 * ks);
                double rr = 4.0 + 6.0 * U(15 + 3 * ks);
                double q = sqrt(1 - z * z), s, c;
                sincos(az, &s, &c);
                const double sph[4] = {rr * q * c, rr * q * s, rr * z, 1.0};
                put_sphere(ks, sph);
            } else {
                const double sph[4] = {1e6, 1e6, 1e6, 0.0};
                put_sphere(ks, sph);
            }
        }
    }
}

// the whole reset by one thread (single-launch layouts: the env's own thread re-initialises it inside the step)
template <typename T>
static __device__ __noinline__ void reset_env(const KParams<T> &p, int64_t i) {
    const uint32_t ep = (uint32_t)p.episode[i];
#pragma unroll 1
    for (int role = 0; role < kResetRoles; role++) reset_env_role<T>(p, i, role, ep);
    p.episode[i] = (int32_t)(ep + 1);
}

// The reset of up to 32 envs by one CTA of kResetRoles warps: warp = role, lane = env.  Roles are different code paths,
// so they must not share a warp (eight roles on eight lanes of one warp run one after the other: measured, no faster than
// one thread per reset); as warps they run side by side and every lane of a warp does the same thing for another env.
//   valid: this lane has an env.  Must be called by all threads of the CTA.
constexpr int kResetCta = 32 * kResetRoles;

template <typename T>
__device__ __forceinline__ void reset_envs_cta(const KParams<T> &p, int64_t i, bool valid) {
    const int role = threadIdx.x >> 5;
    uint32_t ep = 0;
    if (valid) ep = (uint32_t)p.episode[i];
    __syncthreads();                                 // every role has read the episode counter before role 0 bumps it
    if (valid) {
        reset_env_role<T>(p, i, role, ep);
        if (role == 0) p.episode[i] = (int32_t)(ep + 1);
    }
}

// dockauv_reset: one CTA of kResetRoles warps per 32 envs
template <typename T>
__global__ void __launch_bounds__(kResetCta) reset_kernel(const __grid_constant__ KParams<T> p, const uint8_t *mask) {
    const int64_t i = p.env_begin + (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
    const bool valid = i < p.env_end && (mask == nullptr || mask[i] != 0);
    reset_envs_cta<T>(p, i, valid);
}

// ------------------------------------------------------------------------------------------- step pieces
// Everything the non-radar part of one step hands to the radar / finalisation part.
template <typename T>
struct StepCarry {
    T pos[3];          // post-step position
    T R[9];            // post-step Rzyx
    T partial_reward;  // reward terms that do not depend on the radar or the collision flag, summed later
    T rarr[13];        // reward_arr (only [0..5], [7] filled here)
    uint32_t cond;     // bits 0..3
    T delta_d;
    int32_t t_steps;   // episode step counter before this step
    T ep_return;       // cumulative reward before this step
};

// what command_and_penalty reads of one env: the raw action (float32 or float64, as the caller handed it over) and the
// previous low-passed command.  Loaded as one batch before anything is used: the float32 division of the penalty has a
// slow-path branch per actuator, loads are not moved across it, and with the loads inside the loop a thread paid NU memory
// round trips one after the other (ncu: 5 % of the dynamics launch's stall samples on these six waits)
template <typename T, int NU>
struct CommandIn {
    float a32[NU];
    double a64[NU];
    T up[NU];
};

template <typename T, int NU>
__device__ __forceinline__ void load_command(const KParams<T> &p, int64_t i, CommandIn<T, NU> &in) {
    const int64_t N = p.n_envs;
#pragma unroll
    for (int k = 0; k < NU; k++) in.up[k] = p.u_prev[(int64_t)k * N + i];
    if (p.act_f32) {
        const float *a = (const float *)p.actions + i * NU;
#pragma unroll
        for (int k = 0; k < NU; k++) in.a32[k] = a[k];
    } else {
        const double *a = (const double *)p.actions + i * NU;
#pragma unroll
        for (int k = 0; k < NU; k++) in.a64[k] = a[k];
    }
}

// action -> low-passed command (auvsim.py:67-87, lowpassfilter.py:29-42) and the action penalty
// (docking3d.py:584-585).  Returns -(sum((|a|/n_u)^2 * w)).
template <typename T, int NU>
__device__ __forceinline__ T command_and_penalty(const KParams<T> &p, const CommandIn<T, NU> &in, T u[NU]) {
    T pen;
    if (p.act_f32) {
        float pen32 = 0.0f;
        double pen64 = 0.0;
#pragma unroll
        for (int k = 0; k < NU; k++) {
            float ak = in.a32[k];
            float c = ak < -1.0f ? -1.0f : (ak > 1.0f ? 1.0f : ak);
            float frac = (c + 1.0f) / 2.0f;                       // numpy keeps this in float32
            T x = p.u_lo[k] + p.u_span[k] * (T)frac;
            u[k] = p.lp_alpha * x + (T(1) - p.lp_alpha) * in.up[k];
            // numpy evaluates (|a| / n_u) ** 2 * w and the sum in float32 with one rounding per operation:
            // explicit _rn intrinsics keep the compiler from contracting them into FMAs
            float q = __fdiv_rn(fabsf(ak), (float)NU);
            q = __fmul_rn(q, q);
            pen32 = __fadd_rn(pen32, __fmul_rn(q, p.arf_f32[k]));
            pen64 += (double)q * (double)p.arf[k];
        }
        pen = p.action_factor_is_scalar ? (T)pen32 : (T)pen64;
    } else {
        T s = T(0);
#pragma unroll
        for (int k = 0; k < NU; k++) {
            T ak = (T)in.a64[k];
            T frac = (clipv(ak, T(-1), T(1)) + T(1)) / T(2);
            T x = p.u_lo[k] + p.u_span[k] * frac;
            u[k] = p.lp_alpha * x + (T(1) - p.lp_alpha) * in.up[k];
            T q = Mth<T>::abs_(ak) / T(NU);
            s += (q * q) * p.arf[k];
        }
        pen = s;
    }
    return -pen;
}

template <typename T, int NU>
__device__ __forceinline__ T command_and_penalty(const KParams<T> &p, int64_t i, T u[NU]) {
    CommandIn<T, NU> in;
    load_command<T, NU>(p, i, in);
    return command_and_penalty<T, NU>(p, in, u);
}

// np.sum over the 13 reward terms in numpy's pairwise order for n = 13 (8 unrolled accumulators combined
// as a tree, then the 5 remaining terms added sequentially).
template <typename T>
__device__ __forceinline__ T reward_sum13(const T r[13]) {
    T s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
#pragma unroll
    for (int k = 8; k < 13; k++) s += r[k];
    return s;
}

// Episode statistics of one warp: each lane remembers the outcome of its env; if any episode of the warp ended,
// a shuffle reduction folds the contributions and lane 0 issues one global atomic per non-zero statistic.
struct WarpStats {
    bool done = false, nan = false;
    uint32_t cond = 0;
    int32_t length = 0;
    double ep_return = 0.0, delta_d = 0.0;

    // n_steps = env-steps to account for (callers pass the CTA's count from one warp only, 0 from the others:
    // one same-address atomic per CTA instead of one per warp)
    __device__ __forceinline__ void flush(double *g, int n_steps) const {
        const int lane = threadIdx.x & 31;
        g += (blockIdx.x & (DOCKAUV_STAT_COPIES - 1)) * DOCKAUV_N_STATS;   // replica of this CTA (dockauv_fold_stats)
        if (__any_sync(0xffffffffu, done)) {
            double v[DOCKAUV_STAT_ENV_STEPS];
            v[DOCKAUV_STAT_EPISODES] = done ? 1.0 : 0.0;
            v[DOCKAUV_STAT_SUM_RETURN] = done ? ep_return : 0.0;
            v[DOCKAUV_STAT_SUM_LENGTH] = done ? (double)length : 0.0;
#pragma unroll
            for (int k = 0; k < 5; k++) v[DOCKAUV_STAT_COND0 + k] = (done && ((cond >> k) & 1u)) ? 1.0 : 0.0;
            v[DOCKAUV_STAT_SUM_FINAL_DELTA_D] = done ? delta_d : 0.0;
            v[DOCKAUV_STAT_NAN_ENVS] = (done && nan) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < DOCKAUV_STAT_ENV_STEPS; k++) {
                double x = v[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                if (lane == 0 && x != 0.0) atomicAdd(&g[k], x);
            }
        }
        if (lane == 0 && n_steps > 0) atomicAdd(&g[DOCKAUV_STAT_ENV_STEPS], (double)n_steps);
    }

    // the same without the shuffle reduction: every lane whose episode ended adds its own contributions (fire-and-forget
    // atomics into the CTA's replica).  ~1 % of the lanes end an episode, but 27 % of the warps contain one: the reduction
    // costs those warps ~150 instructions, the direct form ~8 atomics per ended episode
    __device__ __forceinline__ void flush_direct(double *g, int n_steps) const {
        g += (blockIdx.x & (DOCKAUV_STAT_COPIES - 1)) * DOCKAUV_N_STATS;
        if (done) {
            atomicAdd(&g[DOCKAUV_STAT_EPISODES], 1.0);
            atomicAdd(&g[DOCKAUV_STAT_SUM_RETURN], ep_return);
            atomicAdd(&g[DOCKAUV_STAT_SUM_LENGTH], (double)length);
#pragma unroll
            for (int k = 0; k < 5; k++)
                if ((cond >> k) & 1u) atomicAdd(&g[DOCKAUV_STAT_COND0 + k], 1.0);
            atomicAdd(&g[DOCKAUV_STAT_SUM_FINAL_DELTA_D], delta_d);
            if (nan) atomicAdd(&g[DOCKAUV_STAT_NAN_ENVS], 1.0);
        }
        if (n_steps > 0) atomicAdd(&g[DOCKAUV_STAT_ENV_STEPS], (double)n_steps);
    }
};

}  // namespace dockauv
