// dockauv_step_tpe.cuh -- layout DOCKAUV_LAYOUT_THREAD_PER_ENV: one thread carries one env through the whole
// step (docking3d.py:346-402).  Ray directions and every constant come from the constant bank (uniform
// reads); per-env capsule pre-computations live in a small local array.  6x slower than the pipeline on the C4
// workload; kept as the independently written cross-check (the reference's per-ray bookkeeping, library hypot and
// divisions, no culls).
#pragma once
#include "dockauv_rays.cuh"

namespace dockauv {

// ---- pieces of the per-env step that do not need the radar; shared by every layout (same expressions, same bits)

// Current.sim (current.py:78-96) then nu_c from the PRE-step attitude (docking3d.py:348-349).
//   tr0: sin/cos of the pre-step attitude
template <typename T>
__device__ __forceinline__ void dyn_current(const KParams<T> &p, int64_t i, const T tr0[6], T nu_c[3]) {
    const int64_t N = p.n_envs;
    T Vc = p.current[i];
    T alpha = p.current[N + i], beta = p.current[2 * N + i];
    T vmin = p.current[3 * N + i], vmax = p.current[4 * N + i];
    T w = T(0);
    if (p.has_noise) {
        w = (p.noise != nullptr) ? p.noise[i]
                                 : (T)((double)p.cur_sigma *
                                       philox_normal(p.seed, p.env_id0 + (uint64_t)i, (uint32_t)p.episode[i],
                                                     (uint32_t)p.t_steps[i]));
    }
    T Vc_dot = -p.cur_mu * Vc + w;
    Vc = Vc + Vc_dot * p.h;
    Vc = clipv(Vc, vmin, vmax);
    p.current[i] = Vc;
    T sa, ca, sb, cb;
    Mth<T>::sincos_(alpha, &sa, &ca);
    Mth<T>::sincos_(beta, &sb, &cb);
    T vn[3] = {Vc * ca * cb, Vc * sb, Vc * sa * cb};                 // current.py:71-73
    T R[9];
    rzyx(tr0[0], tr0[1], tr0[2], tr0[3], tr0[4], tr0[5], R);
#pragma unroll
    for (int c = 0; c < 3; c++) nu_c[c] = R[c] * vn[0] + R[3 + c] * vn[1] + R[6 + c] * vn[2];   // R^T v
}

// generalised force of the low-passed command: BlueROV2 B u (B constant, BlueROV2.py:74-75); LAUV keeps u (B(nu) u per stage)
template <typename T, int VEH, int NU>
__device__ __forceinline__ void dyn_tau(const KParams<T> &p, const T u[NU], T tau[6]) {
    if (VEH == DOCKAUV_VEHICLE_BLUEROV2) {
#pragma unroll
        for (int r = 0; r < 6; r++) {
            T s = T(0);
#pragma unroll
            for (int k = 0; k < NU; k++) s += p.B[r * NU + k] * u[k];
            tau[r] = s;
        }
    } else {
        tau[0] = u[0]; tau[1] = u[1]; tau[2] = u[2]; tau[3] = tau[4] = tau[5] = T(0);
    }
}

// Everything of update_navigation_errors / observe / is_done / reward_step (docking3d.py:404-631) that needs no radar,
// no collision flag and no step counter.
template <typename T>
struct DynOut {
    T o[16];            // observation entries 0..15 before the float32 cast
    T r[8];             // reward terms 0..5 and 7; r[6] holds log_precision(delta_d) for reward_set 2
    T ed[3];            // post-step Theta_dot (auvsim.py:108)
    T delta_d, delta_theta, delta_psi;
    uint32_t cond;      // done conditions 0..2
};

//   pos, y: post-step position and (Theta after ssa, nu_r); tr1: sin/cos of the post-step attitude
template <typename T>
__device__ __forceinline__ void dyn_outputs(const KParams<T> &p, const T pos[3], const T y[9], const T tr1[6], const T goal[3],
                                            const T nu_c[3], T penalty, DynOut<T> &q) {
    const T sphi = tr1[0], cphi = tr1[1], sth = tr1[2], cth = tr1[3], spsi = tr1[4], cpsi = tr1[5];
    const T *nu = y + 3;
    {
        T inv_cth = Mth<T>::rcp_(cth), tth = sth * inv_cth;
        T qs = sphi * nu[4] + cphi * nu[5];
        q.ed[0] = nu[3] + tth * qs;
        q.ed[1] = cphi * nu[4] - sphi * nu[5];
        q.ed[2] = qs * inv_cth;
    }
    // ---- navigation errors (docking3d.py:404-413)
    T diff[3];
#pragma unroll
    for (int c = 0; c < 3; c++) diff[c] = goal[c] - pos[c];
    T dxy2 = diff[0] * diff[0] + diff[1] * diff[1];
    T delta_d = Mth<T>::sqrt_(dxy2 + diff[2] * diff[2]);
    T delta_theta = y[1] + ssa<T>(Mth<T>::atan2_(diff[2], Mth<T>::sqrt_(dxy2)));
    T delta_psi = ssa<T>(Mth<T>::atan2_(diff[1], diff[0]) - y[2]);
    q.delta_d = delta_d;
    q.delta_theta = delta_theta;
    q.delta_psi = delta_psi;

    // ---- observe (docking3d.py:462-488), entries 0..15
    T *o = q.o;
    T lg = Mth<T>::log_(delta_d * p.inv_max_dist_from_goal);
    o[0] = clipv(T(1) - lg * p.inv_log_den_obs, T(0), T(1));
    o[1] = clipv(delta_theta * Mth<T>::inv_half_pi, T(-1), T(1));
    o[2] = clipv(delta_psi * Mth<T>::inv_pi, T(-1), T(1));
    o[3] = clipv(nu[0] * p.inv_u_max, T(-1), T(1));
    o[4] = clipv(nu[1] * p.inv_v_max, T(-1), T(1));
    o[5] = clipv(nu[2] * p.inv_w_max, T(-1), T(1));
    o[6] = clipv(y[0] * p.inv_max_attitude, T(-1), T(1));
    o[7] = clipv(y[1] * p.inv_max_attitude, T(-1), T(1));
    o[8] = clipv(spsi, T(-1), T(1));
    o[9] = clipv(cpsi, T(-1), T(1));
    o[10] = clipv(nu[3] * p.inv_p_max, T(-1), T(1));
    o[11] = clipv(nu[4] * p.inv_q_max, T(-1), T(1));
    o[12] = clipv(nu[5] * p.inv_r_max, T(-1), T(1));
    o[13] = clipv(nu_c[0] / T(2), T(-1), T(1));
    o[14] = clipv(nu_c[1] / T(2), T(-1), T(1));
    o[15] = clipv(nu_c[2] / T(2), T(-1), T(1));

    // ---- is_done conditions 0..2 (docking3d.py:606-611)
    uint32_t cond = 0;
    cond |= (delta_d < p.dist_goal_reached_tol) ? 1u : 0u;
    cond |= (delta_d > p.max_dist_from_goal) ? 2u : 0u;
    cond |= ((Mth<T>::abs_(y[0]) > p.max_attitude) || (Mth<T>::abs_(y[1]) > p.max_attitude)) ? 4u : 0u;
    q.cond = cond;

    // ---- reward terms that need no radar (docking3d.py:512-558, 584-588)
    T *r = q.r;
    T lp_d;
    {
        // log_precision(delta_d, tol, max): shares the logarithm with obs[0] unless the epsilon guard bites
        T lgr = (delta_d < T(0.001)) ? log_cold<T>(T(0.001) / p.max_dist_from_goal) : lg;
        lp_d = T(1) - clipv(lgr * p.inv_log_den_rew, T(0), T(1));
    }
    r[0] = -p.w_d * lp_d;
    if (p.reward_set == 1) {
        T a = delta_theta * Mth<T>::inv_half_pi, b = delta_psi * Mth<T>::inv_pi;
        r[1] = -p.w_delta_theta * (a * a);
        r[2] = -p.w_delta_psi * (b * b);
    } else {
        r[1] = -p.w_delta_theta * cont_goal_constraints<T>(Mth<T>::abs_(delta_theta), Mth<T>::half_pi, lp_d);
        r[2] = -p.w_delta_psi * cont_goal_constraints<T>(Mth<T>::abs_(delta_psi), Mth<T>::pi, lp_d);
    }
    {
        T a = y[0] * Mth<T>::inv_half_pi, b = y[1] * Mth<T>::inv_half_pi;
        r[3] = -p.w_phi * (a * a);
        r[4] = -p.w_theta * (b * b);
        T nrm = Mth<T>::sqrt_(q.ed[0] * q.ed[0] + q.ed[1] * q.ed[1] + q.ed[2] * q.ed[2]) * p.inv_p_max;
        r[5] = -p.w_Thetadot * (nrm * nrm);
    }
    r[6] = lp_d;   // parked here for reward_set 2 (replaced by the obstacle-avoidance term later)
    r[7] = penalty;
}

// Dynamics + everything that does not need the radar, for the single-launch layouts.
//   DBG: compile the optional debug outputs in (parity tests); the throughput kernels are built without them.
template <typename T, int VEH, int NU, bool DBG>
__device__ __forceinline__ void step_dynamics(const KParams<T> &p, int64_t i, StepCarry<T> &cy, T &spsi_out, T &cpsi_out,
                                              float obs16[16], T att_out[3]) {
    const int64_t N = p.n_envs;
    T pos[3], y[9];
#pragma unroll
    for (int c = 0; c < 3; c++) pos[c] = p.state[(int64_t)c * N + i];
#pragma unroll
    for (int c = 0; c < 9; c++) y[c] = p.state[(int64_t)(3 + c) * N + i];
    // everything else this step reads is requested up front so that one HBM round trip covers all of it
    T goal[3];
#pragma unroll
    for (int c = 0; c < 3; c++) goal[c] = p.goal[(int64_t)c * N + i];
    const int32_t t_steps = p.t_steps[i];
    cy.t_steps = t_steps;
    cy.ep_return = p.ep_return[i];

    // sin/cos of the pre-step attitude: the only library sincos calls of the step (every later attitude is a small
    // shift of this one, sincos_shift); shared by the current rotation and the first Runge-Kutta stage
    T tr0[6];
    Mth<T>::sincos_(y[0], &tr0[0], &tr0[1]);
    Mth<T>::sincos_(y[1], &tr0[2], &tr0[3]);
    Mth<T>::sincos_(y[2], &tr0[4], &tr0[5]);
    T nu_c[3] = {T(0), T(0), T(0)};
    if (p.has_current) dyn_current<T>(p, i, tr0, nu_c);

    // ---- command
    T u[NU];
    T penalty = command_and_penalty<T, NU>(p, i, u);
#pragma unroll
    for (int k = 0; k < NU; k++) p.u_prev[(int64_t)k * N + i] = u[k];
    T tau[6];
    dyn_tau<T, VEH, NU>(p, u, tau);

    // ---- integrate (auvsim.py:89-108)
    T tr1[6];
    {
        T pacc[3];
        rkf45_step<T, VEH>(p, y, tr0, tau, nu_c, pacc, tr1);
#pragma unroll
        for (int c = 0; c < 3; c++) pos[c] += pacc[c];
    }
#pragma unroll
    for (int c = 0; c < 3; c++) y[c] = ssa<T>(y[c]);
#pragma unroll
    for (int c = 0; c < 3; c++) p.state[(int64_t)c * N + i] = pos[c];
#pragma unroll
    for (int c = 0; c < 9; c++) p.state[(int64_t)(3 + c) * N + i] = y[c];

    DynOut<T> q;
    dyn_outputs<T>(p, pos, y, tr1, goal, nu_c, penalty, q);
    rzyx(tr1[0], tr1[1], tr1[2], tr1[3], tr1[4], tr1[5], cy.R);
#pragma unroll
    for (int c = 0; c < 3; c++) cy.pos[c] = pos[c];
    spsi_out = tr1[4];
    cpsi_out = tr1[5];
#pragma unroll
    for (int c = 0; c < 3; c++) att_out[c] = y[c];
    cy.delta_d = q.delta_d;
#pragma unroll
    for (int c = 0; c < 16; c++) obs16[c] = (float)q.o[c];
    cy.cond = q.cond | ((t_steps >= p.max_timesteps) ? 8u : 0u);      // docking3d.py:612, pre-increment
#pragma unroll
    for (int c = 0; c < 8; c++) cy.rarr[c] = q.r[c];

    if (DBG) {
        if (p.dbg_euler_dot)
            for (int c = 0; c < 3; c++) p.dbg_euler_dot[(int64_t)c * N + i] = q.ed[c];
        if (p.dbg_nu_c)
            for (int c = 0; c < 3; c++) p.dbg_nu_c[(int64_t)c * N + i] = nu_c[c];
        if (p.dbg_nav) {
            p.dbg_nav[i] = q.delta_d;
            p.dbg_nav[N + i] = q.delta_theta;
            p.dbg_nav[2 * N + i] = q.delta_psi;
        }
        if (p.dbg_obs)
            for (int c = 0; c < 16; c++) p.dbg_obs[(int64_t)c * N + i] = q.o[c];
        if (p.dbg_state_dot) {
            // the full auv._state_dot (auvsim.py:108): right-hand side at the post-step state with the pre-step nu_c
            T sd_pos[3] = {T(0), T(0), T(0)}, sd[9];
            rhs9<T, VEH, true>(p, y, tr1, tau, nu_c, T(1), sd_pos, sd);
            for (int c = 0; c < 3; c++) p.dbg_state_dot[(int64_t)c * N + i] = sd_pos[c];
            for (int c = 0; c < 9; c++) p.dbg_state_dot[(int64_t)(3 + c) * N + i] = sd[c];
        }
    }
}

// Final bookkeeping of one env once the radar term `r_oa` (Reward.obstacle_avoidance, docking3d.py:767-792) and
// the collision flag are known: reward (docking3d.py:560-595), done, counters (:380-385), statistics and the
// SB3-VecEnv style auto-reset.  Returns the done flag.
//   DEFER_RESET: the caller re-initialises finished envs itself (pipeline layout: compacted per CTA).
template <typename T, bool DBG, bool DEFER_RESET = false>
__device__ __forceinline__ bool step_finish(const KParams<T> &p, int64_t i, StepCarry<T> &cy, T r_oa, bool collision,
                                            WarpStats &bs) {
    const int64_t N = p.n_envs;
    T *r = cy.rarr;
    T lp_d = r[6];
    if (p.reward_set == 1) r[6] = -p.w_oa * r_oa;
    else r[6] = -p.w_oa * cont_goal_constraints<T>(Mth<T>::abs_(r_oa), T(1), lp_d);
    uint32_t cond = cy.cond | (collision ? 16u : 0u);
#pragma unroll
    for (int k = 0; k < 5; k++) r[8 + k] = ((cond >> k) & 1u) ? p.w_done[k] : T(0);
    T reward = reward_sum13<T>(r);
    bool done = cond != 0;
    p.reward[i] = reward;
    p.done[i] = done ? 1 : 0;
    if (p.cond_bits) p.cond_bits[i] = (uint8_t)cond;
    if (p.delta_d_out) p.delta_d_out[i] = cy.delta_d;
    T ep_ret = cy.ep_return + reward;
    int32_t t_new = cy.t_steps + 1;
    if (DBG && p.dbg_reward_arr)
        for (int k = 0; k < 13; k++) p.dbg_reward_arr[(int64_t)k * N + i] = r[k];
    if (done) {
        if (p.ep_return_out) p.ep_return_out[i] = ep_ret;
        if (p.ep_len_out) p.ep_len_out[i] = t_new;
        bs.done = true;
        bs.cond = cond;
        bs.length = t_new;
        bs.ep_return = (double)ep_ret;
        bs.delta_d = (double)cy.delta_d;
        bs.nan = reward != reward;      // episodes that ended on a NaN reward
    }
    if (done && p.auto_reset) {
        if (!DEFER_RESET) reset_env<T>(p, i);
    } else {
        p.ep_return[i] = ep_ret;
        p.t_steps[i] = t_new;
    }
    return done;
}

// Writes one observation row (and the terminal-observation row of a finished episode).
template <typename T>
__device__ __forceinline__ void write_obs_row(const KParams<T> &p, int64_t i, const float *row, int n, int offset,
                                              bool done) {
    float *dst = p.obs + i * p.n_obs + offset;
    if (done && p.terminal_obs) {
        float *t = p.terminal_obs + i * p.n_obs + offset;
        for (int c = 0; c < n; c++) t[c] = row[c];
    }
    if (done && p.auto_reset) {
        for (int c = 0; c < n; c++) dst[c] = 0.0f;      // reset() returns the all-zero observation
    } else {
        for (int c = 0; c < n; c++) dst[c] = row[c];
    }
}

template <typename T, int VEH, int NU>
__global__ void __launch_bounds__(128) step_tpe_kernel(const __grid_constant__ KParams<T> p) {
    WarpStats bs;
    bool done = false;
    const int64_t N = p.n_envs;
    const int64_t i0 = p.env_begin + (int64_t)blockIdx.x * blockDim.x;
    const int64_t i = i0 + threadIdx.x;
    if (i < p.env_end) {
        StepCarry<T> cy;
        T spsi, cpsi, att[3];
        float obs16[16];
        step_dynamics<T, VEH, NU, true>(p, i, cy, spsi, cpsi, obs16, att);

        // ---- obstacles: ray-independent pre-computation + body collision (docking3d.py:444-460)
        const T R_safe = p.safety_radius;
        bool collision = false;
        CapPre<T> cp[DOCKAUV_MAX_CAPSULES];
        T soc[DOCKAUV_MAX_SPHERES][3], sc[DOCKAUV_MAX_SPHERES];
        for (int k = 0; k < p.n_sph; k++) {
            T cen[3], rad;
            for (int c = 0; c < 3; c++) cen[c] = p.spheres[(int64_t)(k * 4 + c) * N + i];
            rad = p.spheres[(int64_t)(k * 4 + 3) * N + i];
            T d2 = T(0);
            for (int c = 0; c < 3; c++) {
                soc[k][c] = cy.pos[c] - cen[c];
                d2 += soc[k][c] * soc[k][c];
            }
            sc[k] = d2 - rad * rad;
            collision |= (Mth<T>::sqrt_(d2) <= R_safe + rad);              // shape.py:182-192
        }
        for (int k = 0; k < p.n_caps; k++) {
            T bot[3], top[3], rad;
            for (int c = 0; c < 3; c++) {
                bot[c] = p.capsules[(int64_t)(k * 7 + c) * N + i];
                top[c] = p.capsules[(int64_t)(k * 7 + 3 + c) * N + i];
            }
            rad = p.capsules[(int64_t)(k * 7 + 6) * N + i];
            capsule_pre<T>(cy.pos, bot, top, rad, cp[k]);
            collision |= (dist_segment_point<T>(cy.pos, bot, top) <= rad + R_safe);   // shape.py:195-210
        }

        // ---- radar: rays, min positive distance over obstacles, clamp, 2x2 max-pool, OA reward
        const int n_r = p.n_rays, n_h = p.n_horiz, blk = p.block;
        const T dmax = p.radar_max_dist;
        const bool any_obstacle = (p.n_caps + p.n_sph) > 0;
        T pooled[DOCKAUV_MAX_RAYS / 4 + 32];
        for (int c = 0; c < p.n_rr; c++) pooled[c] = T(0);     // block_reduce pads with cval = 0
        T oa_dot = T(0);
        for (int ir = 0; ir < n_r; ir++) {
            T d = dmax;
            if (any_obstacle) {
                const T b[3] = {p.ray_tab[ir], p.ray_tab[n_r + ir], p.ray_tab[2 * n_r + ir]};   // uniform loads
                T rd[3];
#pragma unroll
                for (int c = 0; c < 3; c++) rd[c] = cy.R[3 * c] * b[0] + cy.R[3 * c + 1] * b[1] + cy.R[3 * c + 2] * b[2];
                T best = Mth<T>::inf(), first = T(0);
                bool have_first = false;
                for (int k = 0; k < p.n_caps; k++) {
                    T v = ray_capsule<T>(cp[k], rd);
                    if (!have_first) { first = v; have_first = true; }
                    if (v > T(0) && v < best) best = v;
                }
                if (p.n_sph > 0) {
                    T sbest = Mth<T>::inf(), sfirst = T(0);
                    for (int k = 0; k < p.n_sph; k++) {
                        T v = ray_sphere<T>(soc[k], sc[k], rd);
                        if (k == 0) sfirst = v;
                        if (v > T(0) && v < sbest) sbest = v;
                    }
                    T v = (sbest < Mth<T>::inf()) ? sbest : sfirst;   // shape.py:264
                    if (!have_first) { first = v; have_first = true; }
                    if (v > T(0) && v < best) best = v;
                }
                d = (best < Mth<T>::inf()) ? best : first;             // docking3d.py:438-439
                if (d < T(0) || d > dmax) d = dmax;                    // sensor.py:117
            }
            if (p.dbg_ray_dist) p.dbg_ray_dist[(int64_t)ir * N + i] = d;
            int iv = ir / n_h, ih = ir - iv * n_h;
            int pc = (iv / blk) * p.n_hr + ih / blk;
            // np.max semantics: NaN propagates
            T cur = pooled[pc];
            pooled[pc] = (d != d || cur != cur) ? (d + cur) : (d > cur ? d : cur);
            T c = clipv(T(1) - d / dmax, T(0), T(1));
            T q = (T(1) - c) * (T(1) - c);
            T mx = (q != q) ? q : (q > T(0.001) ? q : T(0.001));
            oa_dot += mx * p.ray_tab[3 * n_r + ir];
        }
        T r_oa = p.sum_beta_oa / oa_dot - T(1);

        done = step_finish<T, true>(p, i, cy, r_oa, collision, bs);

        // ---- observation row
        write_obs_row<T>(p, i, obs16, 16, 0, done);
        {
            float *dst = p.obs + i * p.n_obs + 16;
            float *tdst = (done && p.terminal_obs) ? p.terminal_obs + i * p.n_obs + 16 : nullptr;
            bool zero = done && p.auto_reset;
            for (int c = 0; c < p.n_rr; c++) {
                T v = clipv(pooled[c] / dmax, T(0), T(1));
                if (p.dbg_obs) p.dbg_obs[(int64_t)(16 + c) * N + i] = v;
                if (tdst) tdst[c] = (float)v;
                dst[c] = zero ? 0.0f : (float)v;
            }
        }
    }
    {
        int64_t left = p.env_end - i0;
        bs.flush(p.stats, threadIdx.x < 32 ? (int)(left < (int64_t)blockDim.x ? left : (int64_t)blockDim.x) : 0);
    }
}

}  // namespace dockauv
