/* dockauv.h -- C ABI of the B200-native batched docking-AUV simulator (libdockauv_b200.so).
 *
 * This is the drop-in boundary for the `env.step()` hot path of Erikx3/gym_dockauv.  Every entry point
 * names the reference interface it replaces (paths relative to the reference repo root):
 *
 *   dockauv_create / dockauv_destroy   BaseDocking3d.__init__            gym_dockauv/envs/docking3d.py:48-220
 *                                      (vehicle + env_config + Radar constants frozen into one handle)
 *   dockauv_bind                       the per-env attributes that persist across steps:
 *                                      auv.state, auv.u, goal_location, heading_goal_reached, current,
 *                                      capsules / spheres, t_steps, cumulative_reward
 *   dockauv_reset                      BaseDocking3d.reset + <Scenario>.generate_environment
 *                                                                          docking3d.py:222-322, 803-988
 *   dockauv_step                       BaseDocking3d.step                 docking3d.py:346-402
 *                                      (Current.sim, AUVSim.step, Radar.update, update_radar_collision,
 *                                       update_body_collision, update_navigation_errors, observe, is_done,
 *                                       reward_step, and SB3-VecEnv style auto-reset)
 *   dockauv_step_host                  the same call with HOST buffers (what SB3's DummyVecEnv.step_wait
 *                                      hands over, train.py:64-71): H2D, kernel(s), D2H inside the call
 *   dockauv_get_stats / _clear_stats   FullDataStorage.update bookkeeping  utils/datastorage.py:65-74
 *   dockauv_rollout                    the step loop of the caller: model.learn() -> collect_rollouts of SB3's
 *                                      on-policy algorithms (train.py:64-71) and predict()'s while-loop
 *                                      (train.py:107-118), for action sequences that are known up front
 *   dockauv_gae                        RolloutBuffer.compute_returns_and_advantage of the PPO caller
 *                                      (stable-baselines3 1.5.0, requirements.txt; selected at train.py:64)
 *   DockauvDebugOut (state_dot, ...)   EpisodeDataStorage.update          utils/datastorage.py:268-288
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross this boundary.
 *   - all `*_dev` pointers are device pointers on the handle's device.  The library never allocates, frees
 *     or keeps caller buffers beyond what dockauv_bind() registered; the caller (PyTorch) owns them.  The
 *     handle owns its ray table, the statistics vector, the buffers between the launches of a step
 *     (a 16-word record, the ray list entry and a float copy of the obstacles: ~350 bytes per env) and, lazily, device
 *     staging for dockauv_step_host.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls are asynchronous and
 *     stream-ordered; one handle must not be used from two threads at once.  Large batches (>= 131,072 envs)
 *     are stepped as two halves on two handle-owned streams that are forked from / joined to `stream` by
 *     events, so everything stays ordered with respect to `stream` (and capturable into a CUDA graph).
 *   - every function returns 0 on success, a negative DOCKAUV_E* code otherwise; dockauv_last_error()
 *     returns a human-readable description of the last failure on the calling thread.
 *   - there is NO CPU fallback: without a usable CUDA device dockauv_create fails with DOCKAUV_ECUDA.
 *
 * Memory layout (HBM): structure-of-arrays, component-major, env index fastest, e.g. state[c * n_envs + i].
 * `real` below is double for DOCKAUV_F64 handles and float for DOCKAUV_F32 handles.
 */
#ifndef DOCKAUV_H
#define DOCKAUV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DOCKAUV_ABI_VERSION 3

#if defined(__GNUC__)
#define DOCKAUV_API __attribute__((visibility("default")))
#else
#define DOCKAUV_API
#endif

#define DOCKAUV_MAX_U 8          /* control inputs: BlueROV2 joystick 6, direct 8; LAUV 3 */
#define DOCKAUV_MAX_CAPSULES 8
#define DOCKAUV_MAX_SPHERES 8
#define DOCKAUV_MAX_RAYS 256
#define DOCKAUV_N_REWARDS 13     /* docking3d.py:153 */
#define DOCKAUV_N_STATS 16

/* error codes */
#define DOCKAUV_OK 0
#define DOCKAUV_EINVAL (-1)      /* bad argument */
#define DOCKAUV_ECUDA (-2)       /* CUDA runtime error (see dockauv_last_error) */
#define DOCKAUV_ESTATE (-3)      /* call order (e.g. step before bind) */

/* precision of the persistent state and of the arithmetic */
#define DOCKAUV_F64 0
#define DOCKAUV_F32 1

/* vehicles (gym_dockauv/objects/vehicles/) */
#define DOCKAUV_VEHICLE_BLUEROV2 0
#define DOCKAUV_VEHICLE_LAUV 1

/* scenarios = the generate_environment() variants, docking3d.py:795-988 */
#define DOCKAUV_SCN_SIMPLE 0
#define DOCKAUV_SCN_SIMPLE_CURRENT 1
#define DOCKAUV_SCN_CAPSULE 2
#define DOCKAUV_SCN_CAPSULE_CURRENT 3
#define DOCKAUV_SCN_OBSTACLES 4
#define DOCKAUV_SCN_OBSTACLES_CURRENT 5
#define DOCKAUV_SCN_OBSTACLES_NOCAP 6

/* action dtypes accepted by dockauv_step (the reference's arithmetic depends on it, auvsim.py:74-75,
 * docking3d.py:584-585: a float32 action array keeps parts of the computation in float32) */
#define DOCKAUV_ACT_F64 0
#define DOCKAUV_ACT_F32 1

/* kernel layouts */
#define DOCKAUV_LAYOUT_AUTO 0
#define DOCKAUV_LAYOUT_THREAD_PER_ENV 1   /* one thread does everything for one env */
#define DOCKAUV_LAYOUT_WARP_RAYS 2        /* one launch: dynamics thread-per-env, radar one warp per env (lanes = rays) */
#define DOCKAUV_LAYOUT_PIPELINE 4         /* dynamics + cull + finish of the envs with nothing in view (thread per env), rays +
                                             finish (a thread or a warp per env that has an obstacle in view, from compact
                                             lists), episode end (per finished env, from a compact list) */

/* indices into the stats vector (sums since the last clear; reduce over ranks with one all-reduce) */
#define DOCKAUV_STAT_EPISODES 0
#define DOCKAUV_STAT_SUM_RETURN 1
#define DOCKAUV_STAT_SUM_LENGTH 2
#define DOCKAUV_STAT_COND0 3            /* ..COND0+4: Done-Goal_reached, out_pos, out_att, max_t, collision */
#define DOCKAUV_STAT_SUM_FINAL_DELTA_D 8
#define DOCKAUV_STAT_NAN_ENVS 9
#define DOCKAUV_STAT_ENV_STEPS 10

/* Everything that is constant for the lifetime of a handle.  Filled by the host side from the reference's
 * env_config dict (config/env_config.py:20-91) and vehicle XML (objects/vehicles/ *.xml). */
typedef struct DockauvParams {
    int32_t abi_version;         /* DOCKAUV_ABI_VERSION */
    int32_t precision;           /* DOCKAUV_F64 / DOCKAUV_F32 */
    int32_t vehicle;             /* DOCKAUV_VEHICLE_* */
    int32_t n_u;                 /* number of control inputs */
    int32_t scenario;            /* DOCKAUV_SCN_* (used by reset / auto-reset) */
    int32_t n_capsules;          /* capsules per env (0..8); row order = reference's self.capsules */
    int32_t n_spheres;           /* spheres per env (0..8) */
    int32_t n_synthetic_spheres; /* reset extension: random unit spheres per env (BASELINE config C4), <= n_spheres */
    int32_t max_timesteps;
    int32_t reward_set;          /* 1 or 2, docking3d.py:515-548 */
    int32_t n_rays, n_vert, n_horiz, block_reduce;   /* sensor.py:43-71; n_rays = n_vert * n_horiz */
    int32_t action_factor_is_scalar;   /* config "action_reward_factors" was a python scalar */
    int32_t layout;              /* DOCKAUV_LAYOUT_* */
    int32_t force_current;       /* evaluate the ocean current even if the scenario spawns none (injected currents) */
    int32_t split_chunk_envs;    /* DOCKAUV_LAYOUT_PIPELINE: envs per launch group (0 = library default: the whole batch, in two
                                    halves on two streams from 131,072 envs on) */
    /* rigid body + hydrodynamics (statespace.py) */
    double m;
    double r_G[3];
    double I_b[9];               /* statespace.py:105-117, row-major */
    double MA_diag[6];           /* diag(M_A) = -(X_udot..N_rdot), statespace.py:164-187 */
    double M_inv[36];            /* inv(M_RB + M_A), statespace.py:190-197, row-major */
    double D_lin[10];            /* linear damping: 6 diagonal entries then [1,5],[2,4],[4,2],[5,1] (LAUV.py:78-83) */
    double D_quad[10];           /* quadratic damping, same order; entry [i,j] multiplies |nu_j| (LAUV.py:84-89) */
    double D_lift[10];           /* lift, same order; every entry multiplies |u| (LAUV.py:90-95,101) */
    double G_WB;                 /* W - BY */
    double G_r[3];               /* (x_G W - x_B BY, y_G W - y_B BY, z_G W - z_B BY), statespace.py:387-396 */
    double B[6 * DOCKAUV_MAX_U]; /* BlueROV2: constant 6 x n_u, row-major (BlueROV2.py:34-72) */
    double lauv_B[4];            /* LAUV: Y_uudr, Z_uuds, M_uuds, N_uudr (LAUV.py:59-67) */
    double u_lo[DOCKAUV_MAX_U], u_hi[DOCKAUV_MAX_U];   /* u_bound */
    double lp_alpha;             /* h / (h + T1), lowpassfilter.py:27 */
    double h;                    /* t_step_size */
    double safety_radius;        /* auvsim.py:43 */
    /* env_config */
    double max_dist_from_goal, max_attitude, dist_goal_reached_tol;
    double u_max, v_max, w_max, p_max, q_max, r_max;
    double w_d, w_delta_psi, w_delta_theta, w_phi, w_theta, w_Thetadot, w_oa;
    double w_done[5];            /* w_goal, w_deltad_max, w_Theta_max, w_t_max, w_col */
    double action_reward_factors[DOCKAUV_MAX_U];
    /* current (objects/current.py) */
    double cur_mu, cur_sigma;
    /* radar */
    double radar_max_dist;
    double rd_b[DOCKAUV_MAX_RAYS * 3];   /* body-frame unit ray directions, sensor.py:63-71 */
    double beta_oa[DOCKAUV_MAX_RAYS];    /* obstacle-avoidance ray weights, docking3d.py:789-790 */
    /* reset */
    uint64_t seed;               /* Philox key */
    uint64_t env_id0;            /* global id of local env 0 (rank offset when the batch is sharded) */
} DockauvParams;

/* Caller-owned persistent per-env state (device pointers). */
typedef struct DockauvBuffers {
    void *state;          /* real[12][N]: x y z phi theta psi | u v w p q r (relative velocity) */
    void *u_prev;         /* real[n_u][N]: low-passed command */
    void *goal;           /* real[3][N] */
    void *heading_goal;   /* real[N] */
    void *current;        /* real[5][N]: V_c, alpha, beta, V_min, V_max */
    void *capsules;       /* real[n_capsules*7][N]: vec_bot[3], vec_top[3], radius per capsule */
    void *spheres;        /* real[n_spheres*4][N]: centre[3], radius per sphere */
    void *ep_return;      /* real[N]: cumulative reward of the running episode */
    int32_t *t_steps;     /* int32[N] */
    int32_t *episode;     /* int32[N]: episodes started so far (Philox counter) */
} DockauvBuffers;

/* Per-step outputs (device pointers; nullable ones are skipped).  When n_obs is a multiple of 4 (every stock radar
 * geometry) observation rows are written with 128-bit stores: obs and terminal_obs must then be 16-byte aligned
 * (DOCKAUV_EINVAL otherwise). */
typedef struct DockauvStepOut {
    float *obs;            /* f32[N][n_obs] row-major; all-zero row when the env was auto-reset (docking3d.py:269,322) */
    void *reward;          /* real[N] */
    uint8_t *done;         /* u8[N] */
    uint8_t *cond_bits;    /* u8[N], nullable: bit k = done condition k (docking3d.py:606-617) */
    float *terminal_obs;   /* f32[N][n_obs], nullable: last observation of episodes that ended this step
                              (rows of envs that are not done are left untouched) */
    void *ep_return_out;   /* real[N], nullable: return of the episode that ended this step (Monitor 'r') */
    int32_t *ep_len_out;   /* int32[N], nullable: length of the episode that ended this step (Monitor 'l') */
    void *delta_d_out;     /* real[N], nullable: distance to the goal after the step = info["delta_d"] (docking3d.py:400) */
} DockauvStepOut;

/* Optional per-step intermediate values for the parity tests (device pointers, all nullable). */
typedef struct DockauvDebugOut {
    void *ray_dist;        /* real[n_rays][N], clamped distances (sensor.py:113-118) */
    void *reward_arr;      /* real[13][N] */
    void *euler_dot;       /* real[3][N], post-step Theta_dot (auvsim.py:108) */
    void *nu_c;            /* real[3][N], pre-step body-frame current (docking3d.py:349) */
    void *nav;             /* real[3][N]: delta_d, delta_theta, delta_psi */
    void *obs_f64;         /* real[n_obs][N]: observation before the float32 cast */
    void *state_dot;       /* real[12][N]: auv._state_dot = state_dot(post-step state, pre-step nu_c), auvsim.py:108
                              (the step itself only needs its Theta_dot part; the full vector is what
                              EpisodeDataStorage logs, datastorage.py:277) */
} DockauvDebugOut;

/* Outputs of dockauv_rollout: the per-step outputs stacked over T steps (device pointers). */
typedef struct DockauvRolloutOut {
    float *obs;            /* f32[T][N][n_obs]: row t = observation returned by step t */
    void *reward;          /* real[T][N] */
    uint8_t *done;         /* u8[T][N] */
    uint8_t *cond_bits;    /* u8[T][N], nullable */
    float *terminal_obs;   /* f32[T][N][n_obs], nullable: rows of episodes that ended at step t (others untouched) */
    void *ep_return_out;   /* real[T][N], nullable: only entries of episodes that ended at step t are written */
    int32_t *ep_len_out;   /* int32[T][N], nullable: same; the call zero-fills it first, so length > 0 marks an end */
    void *delta_d_out;     /* real[T][N], nullable */
} DockauvRolloutOut;

typedef struct DockauvHandle DockauvHandle;

DOCKAUV_API int dockauv_abi_version(void);
DOCKAUV_API const char *dockauv_last_error(void);
DOCKAUV_API size_t dockauv_sizeof_params(void);

/* n_obs = 16 + ceil(n_vert/block) * ceil(n_horiz/block) (docking3d.py:115, sensor.py:135-137) */
DOCKAUV_API int dockauv_n_obs(const DockauvParams *p);

DOCKAUV_API int dockauv_create(const DockauvParams *params, int64_t n_envs, int device, DockauvHandle **out);
DOCKAUV_API int dockauv_destroy(DockauvHandle *h);
DOCKAUV_API int dockauv_bind(DockauvHandle *h, const DockauvBuffers *buffers);

/* Re-key the counter-based random stream (gym's reset(seed=...), docking3d.py:296-298). */
DOCKAUV_API int dockauv_set_seed(DockauvHandle *h, uint64_t seed);

/* Re-initialise envs whose mask byte is non-zero (all envs if mask_dev == NULL). */
DOCKAUV_API int dockauv_reset(DockauvHandle *h, const uint8_t *mask_dev, void *stream);

/* The cull code reads a float copy of the obstacles relative to the goal, which the library keeps next to the bound
 * buffers: dockauv_bind and every reset (dockauv_reset, auto-reset inside a step) write it.  A caller that writes the
 * bound `capsules`, `spheres` or `goal` buffers ITSELF (exact-state injection: the reference's env.capsules = [...] /
 * env.goal_location = ..., docking3d.py:860-946) calls this afterwards, on the stream that did the writes. */
DOCKAUV_API int dockauv_refresh_obstacles(DockauvHandle *h, void *stream);

/* One batched env.step().  actions_dev: [N][n_u] row-major, float or double per action_dtype.
 * noise_dev (nullable): real[N] N(0, sigma) draws for Current.sim; when NULL and cur_sigma > 0 the draw
 * comes from the handle's Philox stream.  auto_reset != 0 re-initialises finished envs in the same launch. */
DOCKAUV_API int dockauv_step(DockauvHandle *h, const void *actions_dev, int action_dtype, const void *noise_dev,
                 const DockauvStepOut *out, const DockauvDebugOut *debug_or_null, int auto_reset, void *stream);

/* dockauv_step replays a CUDA graph of its launch sequence, captured once per set of pointers (actions, noise, outputs) and
 * cached per handle (96 sets; calls with debug outputs, with timing enabled, or on a stream that is itself being captured
 * issue plain launches).  On by default; dockauv_enable_step_graph(h, 0) switches to plain launches. */
DOCKAUV_API int dockauv_enable_step_graph(DockauvHandle *h, int enabled);
DOCKAUV_API int dockauv_step_graph_captures(DockauvHandle *h, int64_t *n_captures);

/* Same step with HOST buffers (pinned memory recommended): actions are copied in, obs/reward/done (and
 * cond_bits if non-NULL) are copied out, pipelined in chunks over the handle's internal streams; returns
 * after the results are in host memory.  The call is ordered after work already issued on the legacy default
 * stream; a caller that drives the handle from another stream synchronises that stream first.
 * reward_host is double[N] or float[N] per the handle precision.
 * aux_dev_or_null: optional DEVICE buffers for the per-episode extras (only terminal_obs, ep_return_out and
 * ep_len_out are read from it); they stay on the device, the caller fetches the few finished rows it needs. */
DOCKAUV_API int dockauv_step_host(DockauvHandle *h, const void *actions_host, int action_dtype, float *obs_host,
                      void *reward_host, uint8_t *done_host, uint8_t *cond_bits_host, int auto_reset,
                      const DockauvStepOut *aux_dev_or_null);

/* T consecutive batched steps with actions known up front (random-action rollouts, replayed action logs):
 * actions_dev is [T][N][n_u] row-major.  Equivalent to T dockauv_step calls writing into row t of `out`, but
 * issued from C (no per-step host round trip); with use_graph != 0 the launch sequence is captured once into a
 * CUDA graph per (pointers, T) and replayed, which is what makes small batches launch-bound no longer.
 * Stochastic-current handles draw their noise from the handle's Philox stream. */
DOCKAUV_API int dockauv_rollout(DockauvHandle *h, const void *actions_dev, int action_dtype, int n_steps,
                    const DockauvRolloutOut *out, int auto_reset, int use_graph, void *stream);

/* Generalised advantage estimation over a stacked rollout, one thread per env walking t = T-1 .. 0
 * (float32 like the caller's RolloutBuffer):
 *   delta_t = r_t + gamma * V_{t+1} * (1 - done_t) - V_t ;  A_t = delta_t + gamma * lambda * (1 - done_t) * A_{t+1}
 * with V_T = last_values, returns = A + V.  rewards_dev is real[T][N] in the precision given by reward_precision
 * (DOCKAUV_F64 / DOCKAUV_F32), i.e. the `reward` rows dockauv_rollout / dockauv_step wrote; done_t is the done flag
 * returned by step t (the next observation starts a new episode).  Needs no handle. */
DOCKAUV_API int dockauv_gae(const void *rewards_dev, int reward_precision, const float *values_dev,
                const float *last_values_dev, const uint8_t *dones_dev, int n_steps, int64_t n_envs, float gamma,
                float gae_lambda, float *advantages_dev, float *returns_dev, void *stream);

/* Episode statistics accumulated on the device since the last clear (DOCKAUV_STAT_*).  stats_dev points at
 * double[DOCKAUV_N_STATS] on the device: it is the send buffer of the per-rollout NCCL all-reduce.  The kernels
 * accumulate into DOCKAUV_STAT_COPIES replicas of the vector (same-address atomics of 10^4 warps per step would
 * serialise in one L2 slice); dockauv_fold_stats adds the replicas into the vector stats_dev points at -- call it on
 * the stream before reading through the pointer.  dockauv_get_stats folds by itself. */
#define DOCKAUV_STAT_COPIES 32
DOCKAUV_API int dockauv_stats_ptr(DockauvHandle *h, double **stats_dev);
DOCKAUV_API int dockauv_fold_stats(DockauvHandle *h, void *stream);
DOCKAUV_API int dockauv_get_stats(DockauvHandle *h, double *stats_host, void *stream);
DOCKAUV_API int dockauv_clear_stats(DockauvHandle *h, void *stream);

/* Micro-benchmarks for the roofline denominators: sustained FP64 / FP32 FMA rate in
 * TFLOP/s and device copy bandwidth in GB/s, measured with CUDA events on `device`. */
DOCKAUV_API int dockauv_measure_peaks(int device, double *fp64_tflops, double *fp32_tflops, double *copy_gbs);

/* Kernel-level accounting for bench.py: number of kernels launched by this handle so far, and CUDA-event
 * time (ms) of the most recent dockauv_step launch when timing is enabled. */
DOCKAUV_API int dockauv_launch_count(DockauvHandle *h, int64_t *n_launches);
/* Work-list lengths of the most recent step of the pipeline layout, summed over the stepped env ranges: envs that had an
 * obstacle in view (the ray launch's work) and envs whose episode ended (the episode-end launch's work).  Synchronises
 * `stream`; diagnostics for bench.py (the counters are re-used by the next step). */
DOCKAUV_API int dockauv_last_list_counts(DockauvHandle *h, int64_t *n_listed, int64_t *n_ended, void *stream);
/* how many times dockauv_rollout(use_graph) had to capture its launch sequence (a replay with the same pointers does not) */
DOCKAUV_API int dockauv_rollout_captures(DockauvHandle *h, int64_t *n_captures);
DOCKAUV_API int dockauv_enable_timing(DockauvHandle *h, int enabled);
DOCKAUV_API int dockauv_last_step_ms(DockauvHandle *h, float *ms);
/* Per-launch CUDA-event times (ms) of the most recent timed dockauv_step of a multi-launch layout, in launch order
 * (DOCKAUV_LAYOUT_PIPELINE: dynamics + cull + finish [, cull + finish when it runs as a launch of its own], rays + finish
 * (only with obstacles), episode end); *n_launches = 0
 * for the single-launch layouts. */
DOCKAUV_API int dockauv_last_step_launch_ms(DockauvHandle *h, float *ms, int capacity, int *n_launches);

#ifdef __cplusplus
}
#endif
#endif /* DOCKAUV_H */
