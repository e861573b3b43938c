/* dockauv_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the gym_dockauv `env.step()` hot path (reference: Erikx3/gym_dockauv,
 * gym_dockauv/envs/docking3d.py:346-402 and everything it calls).  It is the parity checker for the CUDA
 * path in gym_dockauv_b200/: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product package never imports, links or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against (a) the known
 * answers of the reference's own unit tests (tests/objects/test_shape.py, tests/utils/test_geomutils.py,
 * tests/objects/test_BlueROV2.py, tests/objects/test_current.py) and (b) step-by-step traces recorded from
 * the unmodified reference running in the build container (tests/golden/*.npz, made by
 * tests/golden/make_golden.py).
 *
 * The restatement deliberately stays close to the reference's formulation (full 6x6 matrices, the same
 * evaluation order, six RK stages) so that it is easy to audit against the numpy source; the CUDA kernels
 * use a different, matrix-free formulation, which makes the comparison a real cross-check.
 */
#ifndef DOCKAUV_ORACLE_H
#define DOCKAUV_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_U 8
#define ORC_MAX_CAPS 8
#define ORC_MAX_SPH 8
#define ORC_MAX_RAYS 1024
#define ORC_N_REWARDS 13

/* Everything that is constant for one env instance (vehicle + env_config + radar). */
typedef struct {
    /* vehicle (objects/statespace.py, objects/vehicles/*.py) */
    int32_t vehicle;            /* 0 = BlueROV2 (constant B), 1 = LAUV (B(nu) with u^2 terms) */
    int32_t n_u;
    double m, W, BY;
    double r_G[3], r_B[3];
    double I_b[9];              /* statespace.py:105-117 */
    double M_A[36];             /* statespace.py:164-187 */
    double M_inv[36];           /* statespace.py:190-197 (numpy.linalg.inv on the host) */
    double D_lin[36];           /* D = -(D_lin + D_quad .* |nu_col| + L_lift * |u|), statespace.py:288-351, LAUV.py:69-101 */
    double D_quad[36];
    double L_lift[36];
    double B_const[6 * ORC_MAX_U];  /* row-major 6 x n_u, BlueROV2.py:34-43 / 53-62 */
    double lauv_B[4];           /* Y_uudr, Z_uuds, M_uuds, N_uudr, LAUV.py:59-67 */
    double u_lo[ORC_MAX_U], u_hi[ORC_MAX_U];
    double lp_alpha;            /* lowpassfilter.py:13-27 */
    double h;                   /* t_step_size */
    double safety_radius;       /* auvsim.py:43 */
    /* env_config (config/env_config.py:20-91) */
    int32_t max_timesteps;
    int32_t reward_set;
    double max_dist_from_goal, max_attitude, dist_goal_reached_tol;
    double u_max, v_max, w_max, p_max, q_max, r_max;
    double w_d, w_delta_psi, w_delta_theta, w_phi, w_theta, w_Thetadot, w_oa;
    double w_done[5];           /* w_goal, w_deltad_max, w_Theta_max, w_t_max, w_col */
    double action_reward_factors[ORC_MAX_U];
    int32_t action_factor_is_scalar;   /* python float (weak scalar) vs float64 array: matters for f32 actions */
    /* current (objects/current.py) */
    double cur_mu, cur_sigma;
    /* radar (objects/sensor.py) */
    int32_t n_rays, n_vert, n_horiz, block, n_rays_reduced;
    double radar_max_dist;
    double rd_b[ORC_MAX_RAYS * 3];
    double beta_oa[ORC_MAX_RAYS];      /* docking3d.py:789-790 */
    int32_t n_obs;
} OrcParams;

/* Per-env persistent state. */
typedef struct {
    double state[12];
    double u[ORC_MAX_U];
    double goal[3];
    double heading_goal;
    double cur[5];              /* V_c, alpha, beta, V_min, V_max */
    int32_t n_caps, n_sph;
    double caps[ORC_MAX_CAPS][7];   /* vec_bot[3], vec_top[3], radius */
    double sph[ORC_MAX_SPH][4];     /* centre[3], radius */
    int32_t t_steps;
    int32_t episode;
    double cum_reward;
} OrcEnv;

/* Everything one step produces (superset of what gym returns; used by the parity tests). */
typedef struct {
    float obs[16 + ORC_MAX_RAYS / 4 + 64];
    double reward;
    double reward_arr[ORC_N_REWARDS];
    uint8_t cond[5];
    uint8_t collision, done, goal_reached;
    double ray_dist[ORC_MAX_RAYS];
    double state_dot[12];
    double nu_c[6];
    double delta_d, delta_theta, delta_psi, delta_heading_goal;
} OrcStepOut;

/* building blocks (exported so the tests can pin each one separately) */
double orc_ssa(double x);                                                   /* geomutils.py:4-11 */
void orc_Rzyx(double phi, double theta, double psi, double R[9]);           /* geomutils.py:14-43 */
void orc_Tzyx(double phi, double theta, double T[9]);                       /* geomutils.py:46-75 */
void orc_C(const OrcParams *p, const double nu[6], double C[36]);           /* statespace.py:199-286 */
void orc_D(const OrcParams *p, const double nu[6], double D[36]);           /* statespace.py:288-351, LAUV.py:69-101 */
void orc_G(const OrcParams *p, const double eta[6], double G[6]);           /* statespace.py:353-397 */
void orc_B(const OrcParams *p, const double nu[6], double B[6 * ORC_MAX_U]); /* vehicles */
void orc_state_dot(const OrcParams *p, const double y[12], const double u[], const double nu_c[6],
                   double out[12]);                                         /* auvsim.py:110-160 */
void orc_unnormalize(const OrcParams *p, const void *action, int action_is_f32, double x[]); /* auvsim.py:67-75 */
void orc_auv_step(const OrcParams *p, double state[12], double u[], const void *action, int action_is_f32,
                  const double nu_c[6], double state_dot[12]);              /* auvsim.py:77-108 */
void orc_current_nu_c(const double cur[5], const double att[3], double nu_c[6]);   /* current.py:33-76 */
double orc_ray_capsule(const double l1[3], const double ld[3], const double cap1[3], const double cap2[3],
                       double rad);                                         /* shape.py:327-390 (per ray) */
double orc_ray_spheres(const double l1[3], const double ld[3], const double *centres, const double *rads,
                       int n);                                              /* shape.py:235-264 (per ray) */
double orc_dist_line_point(const double po[3], const double l1[3], const double l2[3]);  /* shape.py:393-417 */
int orc_collision_capsule_sphere(const double c1[3], const double c2[3], double cr, const double sp[3],
                                 double sr);                                /* shape.py:195-210 */
int orc_collision_sphere_spheres(const double p1[3], double r1, const double *p2, const double *r2, int n); /* shape.py:182-192 */
void orc_block_reduce_max(const double *d, int n_v, int n_h, int block, double *out);   /* sensor.py:131-137 */
double orc_obstacle_avoidance(const OrcParams *p, const double *d);         /* docking3d.py:767-792 */
double orc_log_precision(double x, double x_goal, double x_max);            /* docking3d.py:712-723 */

/* One env.step() (docking3d.py:346-402).  `action` points at n_u float (action_is_f32) or double values.
 * `noise_w` is the N(0, sigma) draw of Current.sim (current.py:88); pass 0 when sigma == 0. */
void orc_step(const OrcParams *p, OrcEnv *e, const void *action, int action_is_f32, double noise_w,
              OrcStepOut *out);

/* Deterministic counter-based re-initialisation used for auto-reset (distributions of docking3d.py:687-703 and
 * the generate_environment() of each scenario, 803-988; the random stream is Philox4x32-10 keyed by
 * (seed, env id, episode), NOT the reference's global MT19937 -- see DESIGN.md). scenario ids in dockauv.h. */
void orc_reset_env(const OrcParams *p, OrcEnv *e, int scenario, uint64_t seed, uint64_t env_id);

/* Batched driver for the CPU baseline: steps n envs (AoS) with OpenMP, auto-resetting finished episodes.
 * actions: [n][n_u] float or double.  Returns the number of episodes that finished. */
int64_t orc_step_batch(const OrcParams *p, OrcEnv *envs, int64_t n, const void *actions, int action_is_f32,
                       int scenario, uint64_t seed, uint64_t env_id0, float *obs, double *reward, uint8_t *done,
                       int n_threads);

/* The same for an arbitrary sample of a larger batch: env i has the global id env_ids[i] (NULL: env_id0 + i);
 * cond_bits (nullable): bit k = done condition k of the step. */
int64_t orc_step_batch_ids(const OrcParams *p, OrcEnv *envs, int64_t n, const void *actions, int action_is_f32,
                           int scenario, uint64_t seed, uint64_t env_id0, const uint64_t *env_ids, float *obs,
                           double *reward, uint8_t *done, uint8_t *cond_bits, int n_threads);

int orc_sizeof_params(void);
int orc_sizeof_env(void);
int orc_sizeof_stepout(void);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
