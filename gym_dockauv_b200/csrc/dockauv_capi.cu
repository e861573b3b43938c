// dockauv_capi.cu -- C ABI (include/dockauv.h) over the sm_100a step kernels.
//
// No torch types, no global state besides a thread-local error string.  All device memory that crosses the
// boundary is caller-owned; the handle owns only its small ray table, the statistics vector and (lazily) the
// device staging buffers + streams of the host-buffer entry point.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "dockauv_launch.h"

using namespace dockauv;

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(DOCKAUV_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                                       \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

#ifndef DOCKAUV_HOST_CHUNKS
#define DOCKAUV_HOST_CHUNKS 12     // chunks of dockauv_step_host (copies of chunk c + 1 overlap the launches of chunk c)
#endif
static const int kHostStreams = 3;
static const size_t kStatsBytes = sizeof(double) * DOCKAUV_N_STATS * DOCKAUV_STAT_COPIES;   // replica 0 = the public vector

struct DockauvHandle {
    DockauvParams params;
    int64_t n_envs = 0;
    int device = 0;
    int n_obs = 0;
    bool bound = false;
    KParams<double> kd;
    KParams<float> kf;
    void *ray_tab = nullptr;
    void *pipe_buf = nullptr;      // pipeline layout: per-env records, float obstacle records, ray list + counters
    double *stats = nullptr;
    int64_t launches = 0;
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
    cudaEvent_t marks[kMaxStepLaunches + 1] = {};   // per-launch marks of the multi-launch layouts (timing enabled)
    int n_marks = 0;
    // host pipeline
    cudaStream_t hs[kHostStreams] = {nullptr, nullptr, nullptr};
    cudaEvent_t hev[kHostStreams] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr;   // device step in parts on the internal streams (see step_parts)
    void *st_actions = nullptr;
    size_t st_actions_bytes = 0;
    float *st_obs = nullptr;
    void *st_reward = nullptr;
    uint8_t *st_done = nullptr, *st_cond = nullptr;
    // rollout graph cache: one captured launch sequence, re-captured when the key changes
    cudaGraphExec_t rg_exec = nullptr;
    int64_t rg_launches = 0;       // kernel nodes in the captured sequence
    struct RolloutKey {
        const void *actions = nullptr;
        DockauvRolloutOut out = {};
        int64_t n_steps = 0, action_dtype = 0, auto_reset = 0;      // 64-bit fields: no padding bytes in the key
        uint64_t seed = 0;
        bool operator==(const RolloutKey &o) const {
            return actions == o.actions && out.obs == o.out.obs && out.reward == o.out.reward && out.done == o.out.done &&
                   out.cond_bits == o.out.cond_bits && out.terminal_obs == o.out.terminal_obs &&
                   out.ep_return_out == o.out.ep_return_out && out.ep_len_out == o.out.ep_len_out &&
                   out.delta_d_out == o.out.delta_d_out && n_steps == o.n_steps && action_dtype == o.action_dtype &&
                   auto_reset == o.auto_reset && seed == o.seed;
        }
    } rg_key;
    int64_t rg_captures = 0;       // how many times a rollout graph was captured (tests: replays must not re-capture)
    // step graph cache: the launch sequence of ONE dockauv_step (launches of both halves, fork / join events) captured per
    // set of pointers and replayed; trainers alternate between a handful of action / output tensors
    struct StepKey {
        const void *actions = nullptr, *noise = nullptr;
        DockauvStepOut out = {};
        int64_t action_dtype = 0, auto_reset = 0;
        uint64_t seed = 0;
        bool operator==(const StepKey &o) const {
            return actions == o.actions && noise == o.noise && out.obs == o.out.obs && out.reward == o.out.reward &&
                   out.done == o.out.done && out.cond_bits == o.out.cond_bits && out.terminal_obs == o.out.terminal_obs &&
                   out.ep_return_out == o.out.ep_return_out && out.ep_len_out == o.out.ep_len_out &&
                   out.delta_d_out == o.out.delta_d_out && action_dtype == o.action_dtype && auto_reset == o.auto_reset &&
                   seed == o.seed;
        }
    };
    struct StepGraph {
        StepKey key;
        cudaGraphExec_t exec = nullptr;
        int64_t launches = 0;
        std::vector<int64_t> begins;
    };
    std::vector<int64_t> last_begins;     // first env of every range the most recent step call stepped (their list counters)
    std::vector<StepGraph> sg;
    size_t sg_next = 0;            // slot replaced next once the cache is full
    bool sg_enabled = true;
    int64_t sg_captures = 0;
};

static int pooled_dim(int n, int b) { return (n + b - 1) / b; }

static void drop_graphs(DockauvHandle *h) {      // captured launch sequences hold the old pointers / seed
    if (h->rg_exec) {
        cudaGraphExecDestroy(h->rg_exec);
        h->rg_exec = nullptr;
    }
    for (auto &g : h->sg)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    h->sg.clear();
    h->sg_next = 0;
}

extern "C" int dockauv_abi_version(void) { return DOCKAUV_ABI_VERSION; }
extern "C" const char *dockauv_last_error(void) { return g_err; }
extern "C" size_t dockauv_sizeof_params(void) { return sizeof(DockauvParams); }

extern "C" int dockauv_n_obs(const DockauvParams *p) {
    if (!p || p->block_reduce <= 0) return DOCKAUV_EINVAL;
    return 16 + pooled_dim(p->n_vert, p->block_reduce) * pooled_dim(p->n_horiz, p->block_reduce);
}

template <typename T>
static void fill_kparams(const DockauvParams &s, int64_t n_envs, KParams<T> &k) {
    memset(&k, 0, sizeof(k));
    k.n_envs = n_envs;
    k.env_begin = 0;
    k.env_end = n_envs;
    k.n_u = s.n_u;
    k.n_caps = s.n_capsules;
    k.n_sph = s.n_spheres;
    k.n_synth_sph = s.n_synthetic_spheres;
    k.scenario = s.scenario;
    k.max_timesteps = s.max_timesteps;
    k.reward_set = s.reward_set;
    k.n_rays = s.n_rays;
    k.n_vert = s.n_vert;
    k.n_horiz = s.n_horiz;
    k.block = s.block_reduce;
    k.n_hr = pooled_dim(s.n_horiz, s.block_reduce);
    k.n_rr = pooled_dim(s.n_vert, s.block_reduce) * k.n_hr;
    k.n_obs = 16 + k.n_rr;
    k.action_factor_is_scalar = s.action_factor_is_scalar;
    k.seed = s.seed;
    k.env_id0 = s.env_id0;
    k.m = (T)s.m;
    for (int i = 0; i < 3; i++) k.r_G[i] = (T)s.r_G[i];
    for (int i = 0; i < 9; i++) k.I_b[i] = (T)s.I_b[i];
    for (int i = 0; i < 6; i++) k.MA[i] = (T)s.MA_diag[i];
    for (int i = 0; i < 36; i++) k.M_inv[i] = (T)s.M_inv[i];
    for (int i = 0; i < 10; i++) {
        k.D_lin[i] = (T)s.D_lin[i];
        k.D_quad[i] = (T)s.D_quad[i];
        k.D_lift[i] = (T)s.D_lift[i];
    }
    k.G_WB = (T)s.G_WB;
    for (int i = 0; i < 3; i++) k.G_r[i] = (T)s.G_r[i];
    for (int i = 0; i < 6 * DOCKAUV_MAX_U; i++) k.B[i] = (T)s.B[i];
    for (int i = 0; i < 4; i++) k.lauv_B[i] = (T)s.lauv_B[i];
    for (int i = 0; i < DOCKAUV_MAX_U; i++) {
        k.u_lo[i] = (T)s.u_lo[i];
        k.u_span[i] = (T)(s.u_hi[i] - s.u_lo[i]);
        k.arf[i] = (T)s.action_reward_factors[i];
        k.arf_f32[i] = (float)s.action_reward_factors[i];
    }
    k.lp_alpha = (T)s.lp_alpha;
    k.h = (T)s.h;
    k.safety_radius = (T)s.safety_radius;
    k.max_dist_from_goal = (T)s.max_dist_from_goal;
    k.max_attitude = (T)s.max_attitude;
    k.dist_goal_reached_tol = (T)s.dist_goal_reached_tol;
    k.u_max = (T)s.u_max; k.v_max = (T)s.v_max; k.w_max = (T)s.w_max;
    k.p_max = (T)s.p_max; k.q_max = (T)s.q_max; k.r_max = (T)s.r_max;
    k.inv_u_max = (T)(1.0 / s.u_max); k.inv_v_max = (T)(1.0 / s.v_max); k.inv_w_max = (T)(1.0 / s.w_max);
    k.inv_p_max = (T)(1.0 / s.p_max); k.inv_q_max = (T)(1.0 / s.q_max); k.inv_r_max = (T)(1.0 / s.r_max);
    k.inv_max_attitude = (T)(1.0 / s.max_attitude);
    k.log_den_obs = (T)std::log(s.dist_goal_reached_tol / s.max_dist_from_goal);
    k.log_den_rew = (T)std::log(std::fmax(s.dist_goal_reached_tol, 0.001) / s.max_dist_from_goal);
    k.inv_max_dist_from_goal = (T)(1.0 / s.max_dist_from_goal);
    k.inv_log_den_obs = (T)(1.0 / std::log(s.dist_goal_reached_tol / s.max_dist_from_goal));
    k.inv_log_den_rew = (T)(1.0 / std::log(std::fmax(s.dist_goal_reached_tol, 0.001) / s.max_dist_from_goal));
    k.w_d = (T)s.w_d; k.w_delta_psi = (T)s.w_delta_psi; k.w_delta_theta = (T)s.w_delta_theta;
    k.w_phi = (T)s.w_phi; k.w_theta = (T)s.w_theta; k.w_Thetadot = (T)s.w_Thetadot; k.w_oa = (T)s.w_oa;
    for (int i = 0; i < 5; i++) k.w_done[i] = (T)s.w_done[i];
    k.cur_mu = (T)s.cur_mu;
    k.cur_sigma = (T)s.cur_sigma;
    k.has_noise = s.cur_sigma > 0.0;
    k.radar_max_dist = (T)s.radar_max_dist;
    double sb = 0.0;
    for (int i = 0; i < s.n_rays; i++) sb += s.beta_oa[i];
    k.sum_beta_oa = (T)sb;
    // bounding pyramid of the ray fan (used by the warp layout's field-of-view cull); a ray with x <= 0 disables it
    double ty = 0.0, tz = 0.0;
    bool fan_ok = true;
    for (int i = 0; i < s.n_rays; i++) {
        const double x = s.rd_b[3 * i];
        if (!(x > 1e-6)) {
            fan_ok = false;
            break;
        }
        ty = std::fmax(ty, std::fabs(s.rd_b[3 * i + 1]) / x);
        tz = std::fmax(tz, std::fabs(s.rd_b[3 * i + 2]) / x);
    }
    if (!fan_ok) ty = tz = 1e30;
    ty *= 1.0 + 1e-9;
    tz *= 1.0 + 1e-9;
    // M_inv with nothing outside the diagonal and [0,4] [4,0] [1,3] [3,1] (centre of gravity offset along z only: both
    // stock vehicles): the kernels then skip the 26 exact zeros of the dense product
    bool sparse = true;
    for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) {
            const bool allowed = r == c || (r == 0 && c == 4) || (r == 4 && c == 0) || (r == 1 && c == 3) || (r == 3 && c == 1);
            if (!allowed && s.M_inv[6 * r + c] != 0.0) sparse = false;
        }
    // ... and the same structure in C(nu) and G(eta): r_G and the buoyancy lever on the z axis, I_b diagonal
    if (s.r_G[0] != 0.0 || s.r_G[1] != 0.0 || s.G_r[0] != 0.0 || s.G_r[1] != 0.0) sparse = false;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            if (r != c && s.I_b[3 * r + c] != 0.0) sparse = false;
#ifdef DOCKAUV_FORCE_DENSE_MINV      // tuning builds
    sparse = false;
#endif
    k.sparse_minv = sparse ? 1 : 0;
    // the cull code works on float records relative to the goal: every coordinate that matters is within
    // max_dist_from_goal + max_dist of it; beyond ~2 km the float rounding would eat the 2 mm slack -> exact culls
    k.cull_exact = (s.max_dist_from_goal + s.radar_max_dist > 2000.0) ? 1 : 0;
    k.fov_ty = (T)ty;
    k.fov_tz = (T)tz;
    k.fov_ny = (T)std::sqrt(1.0 + ty * ty);
    k.fov_nz = (T)std::sqrt(1.0 + tz * tz);
}

extern "C" int dockauv_create(const DockauvParams *p, int64_t n_envs, int device, DockauvHandle **out) {
    if (!p || !out) return fail(DOCKAUV_EINVAL, "null argument");
    *out = nullptr;
    if (p->abi_version != DOCKAUV_ABI_VERSION)
        return fail(DOCKAUV_EINVAL, "ABI version mismatch: caller %d, library %d", p->abi_version, DOCKAUV_ABI_VERSION);
    if (n_envs <= 0) return fail(DOCKAUV_EINVAL, "n_envs must be positive");
    if (p->precision != DOCKAUV_F64 && p->precision != DOCKAUV_F32) return fail(DOCKAUV_EINVAL, "bad precision");
    if (p->vehicle == DOCKAUV_VEHICLE_BLUEROV2) {
        if (p->n_u != 6 && p->n_u != 8) return fail(DOCKAUV_EINVAL, "BlueROV2 needs n_u = 6 (joystick) or 8 (direct)");
    } else if (p->vehicle == DOCKAUV_VEHICLE_LAUV) {
        if (p->n_u != 3) return fail(DOCKAUV_EINVAL, "LAUV needs n_u = 3");
    } else {
        return fail(DOCKAUV_EINVAL, "unknown vehicle %d", p->vehicle);
    }
    if (p->n_capsules < 0 || p->n_capsules > DOCKAUV_MAX_CAPSULES || p->n_spheres < 0 ||
        p->n_spheres > DOCKAUV_MAX_SPHERES)
        return fail(DOCKAUV_EINVAL, "obstacle counts out of range");
    if (p->n_rays <= 0 || p->n_rays > DOCKAUV_MAX_RAYS || p->n_rays != p->n_vert * p->n_horiz || p->block_reduce <= 0)
        return fail(DOCKAUV_EINVAL, "bad radar geometry (n_rays=%d, %d x %d, block %d)", p->n_rays, p->n_vert,
                    p->n_horiz, p->block_reduce);
    if (p->n_synthetic_spheres < 0 || p->n_synthetic_spheres > p->n_spheres)
        return fail(DOCKAUV_EINVAL, "n_synthetic_spheres must be in 0..n_spheres");
    if (!(p->h > 0.0)) return fail(DOCKAUV_EINVAL, "t_step_size must be positive");
    if (p->max_timesteps <= 0) return fail(DOCKAUV_EINVAL, "max_timesteps must be positive");
    if (!(p->u_max > 0.0) || !(p->v_max > 0.0) || !(p->w_max > 0.0) || !(p->p_max > 0.0) || !(p->q_max > 0.0) ||
        !(p->r_max > 0.0) || !(p->max_attitude > 0.0))
        return fail(DOCKAUV_EINVAL, "u_max .. r_max and max_attitude must be positive (they normalise the observation)");
    if (!(p->max_dist_from_goal > 0.0) || !(p->dist_goal_reached_tol > 0.0) || !(p->radar_max_dist > 0.0))
        return fail(DOCKAUV_EINVAL, "max_dist_from_goal, dist_goal_reached_tol and the radar's max_dist must be positive");
    if (p->reward_set != 1 && p->reward_set != 2) return fail(DOCKAUV_EINVAL, "reward_set must be 1 or 2");
    if (p->scenario < 0 || p->scenario > DOCKAUV_SCN_OBSTACLES_NOCAP) return fail(DOCKAUV_EINVAL, "bad scenario");
    if (p->layout != DOCKAUV_LAYOUT_AUTO && p->layout != DOCKAUV_LAYOUT_THREAD_PER_ENV && p->layout != DOCKAUV_LAYOUT_WARP_RAYS &&
        p->layout != DOCKAUV_LAYOUT_PIPELINE)
        return fail(DOCKAUV_EINVAL, "unknown layout %d", p->layout);
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(DOCKAUV_ECUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                    cudaGetErrorString(e));
    if (device < 0 || device >= n_dev) return fail(DOCKAUV_EINVAL, "device %d out of range (0..%d)", device, n_dev - 1);
    DeviceGuard guard(device);
    DockauvHandle *h = new (std::nothrow) DockauvHandle();
    if (!h) return fail(DOCKAUV_EINVAL, "out of host memory");
    h->params = *p;
    h->n_envs = n_envs;
    h->device = device;
    h->n_obs = dockauv_n_obs(p);
    fill_kparams<double>(*p, n_envs, h->kd);
    fill_kparams<float>(*p, n_envs, h->kf);
    // ray table in global memory: rd_b[3][n_rays] then beta_oa[n_rays], in the handle's precision
    {
        const int n = p->n_rays;
        const size_t esz = p->precision == DOCKAUV_F64 ? 8 : 4;
        std::vector<char> host(esz * 4 * n);
        for (int i = 0; i < n; i++)
            for (int c = 0; c < 4; c++) {
                double v = c < 3 ? p->rd_b[3 * i + c] : p->beta_oa[i];
                if (esz == 8) ((double *)host.data())[c * n + i] = v;
                else ((float *)host.data())[c * n + i] = (float)v;
            }
        cudaError_t e1 = cudaMalloc(&h->ray_tab, host.size());
        cudaError_t e2 = e1 == cudaSuccess ? cudaMemcpy(h->ray_tab, host.data(), host.size(), cudaMemcpyHostToDevice) : e1;
        cudaError_t e3 = e2 == cudaSuccess ? cudaMalloc((void **)&h->stats, kStatsBytes) : e2;
        cudaError_t e4 = e3 == cudaSuccess ? cudaMemset(h->stats, 0, kStatsBytes) : e3;
        if (e4 != cudaSuccess) {
            if (h->ray_tab) cudaFree(h->ray_tab);
            if (h->stats) cudaFree(h->stats);
            delete h;
            return fail(DOCKAUV_ECUDA, "device allocation failed: %s", cudaGetErrorString(e4));
        }
        h->kd.ray_tab = (const double *)h->ray_tab;
        h->kf.ray_tab = (const float *)h->ray_tab;
        h->kd.stats = h->stats;
        h->kf.stats = h->stats;
    }
    if (p->layout == DOCKAUV_LAYOUT_PIPELINE || p->layout == DOCKAUV_LAYOUT_AUTO) {
        // library-owned buffers of the pipeline layout: per-env record T[16] (128-byte aligned), float obstacle records
        // float4[2 n_caps + n_sph], the ray list u64 and one list counter per 128 envs
        const size_t esz = p->precision == DOCKAUV_F64 ? 8 : 4;
        const size_t n = (size_t)n_envs;
        const int n_obsf = 2 * p->n_capsules + p->n_spheres;
        const size_t off_rec = 0, off_obsf = off_rec + 16 * esz * n, off_list = off_obsf + 16 * (size_t)n_obsf * n;
        const size_t off_end = off_list + 3 * 8 * n, off_cnt = off_end + 4 * ((n + 1) & ~(size_t)1);      // three view lists
        const size_t n_cnt = 16 * (n / 128 + 2);     // counter block (kCounterStride words) per concurrently stepped env range
        cudaError_t e5 = cudaMalloc(&h->pipe_buf, off_cnt + 4 * n_cnt);
        if (e5 == cudaSuccess) e5 = cudaMemset(h->pipe_buf, 0, off_cnt + 4 * n_cnt);
        if (e5 != cudaSuccess) {
            cudaFree(h->ray_tab);
            cudaFree(h->stats);
            if (h->pipe_buf) cudaFree(h->pipe_buf);
            delete h;
            return fail(DOCKAUV_ECUDA, "device allocation of the pipeline buffers failed: %s", cudaGetErrorString(e5));
        }
        char *base = (char *)h->pipe_buf;
        h->kd.rec = (double *)(base + off_rec);
        h->kf.rec = (float *)(base + off_rec);
        h->kd.obsf = h->kf.obsf = n_obsf > 0 ? (float4 *)(base + off_obsf) : nullptr;
        h->kd.n_obsf = h->kf.n_obsf = n_obsf;
        h->kd.view_list = h->kf.view_list = (unsigned long long *)(base + off_list);
        h->kd.ended_list = h->kf.ended_list = (uint32_t *)(base + off_end);
        h->kd.view_count = h->kf.view_count = (unsigned int *)(base + off_cnt);
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        h->kd.sm_count = h->kf.sm_count = sms;
        h->kd.tpe_rays = h->kf.tpe_rays = (p->block_reduce == 2 && h->kd.n_rr <= 64) ? 1 : 0;
#ifdef DOCKAUV_NO_TPE_RAYS      // tuning builds: every listed env through the warp-per-env ray launch
        h->kd.tpe_rays = h->kf.tpe_rays = 0;
#endif
        // one launch group over the whole batch by default: smaller chunks would keep the records in L2 but lose more
        // to partial waves than they gain (measured, profiles/r01/NOTES.md)
        h->kd.chunk_envs = h->kf.chunk_envs = p->split_chunk_envs > 0 ? p->split_chunk_envs : 0;
    }
    *out = h;
    return DOCKAUV_OK;
}

extern "C" int dockauv_destroy(DockauvHandle *h) {
    if (!h) return DOCKAUV_OK;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    if (h->ray_tab) cudaFree(h->ray_tab);
    if (h->pipe_buf) cudaFree(h->pipe_buf);
    if (h->stats) cudaFree(h->stats);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (cudaEvent_t m : h->marks)
        if (m) cudaEventDestroy(m);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (int s = 0; s < kHostStreams; s++) {
        if (h->hs[s]) cudaStreamDestroy(h->hs[s]);
        if (h->hev[s]) cudaEventDestroy(h->hev[s]);
    }
    if (h->st_actions) cudaFree(h->st_actions);
    if (h->st_obs) cudaFree(h->st_obs);
    if (h->st_reward) cudaFree(h->st_reward);
    if (h->st_done) cudaFree(h->st_done);
    if (h->st_cond) cudaFree(h->st_cond);
    drop_graphs(h);
    delete h;
    return DOCKAUV_OK;
}

template <typename T>
static void bind_k(KParams<T> &k, const DockauvBuffers &b) {
    k.state = (T *)b.state;
    k.u_prev = (T *)b.u_prev;
    k.goal = (T *)b.goal;
    k.heading_goal = (T *)b.heading_goal;
    k.current = (T *)b.current;
    k.capsules = (T *)b.capsules;
    k.spheres = (T *)b.spheres;
    k.ep_return = (T *)b.ep_return;
    k.t_steps = b.t_steps;
    k.episode = b.episode;
}

extern "C" int dockauv_bind(DockauvHandle *h, const DockauvBuffers *b) {
    if (!h || !b) return fail(DOCKAUV_EINVAL, "null argument");
    if (!b->state || !b->u_prev || !b->goal || !b->heading_goal || !b->current || !b->ep_return || !b->t_steps ||
        !b->episode)
        return fail(DOCKAUV_EINVAL, "a required state buffer is null");
    if ((h->params.n_capsules > 0 && !b->capsules) || (h->params.n_spheres > 0 && !b->spheres))
        return fail(DOCKAUV_EINVAL, "obstacle buffer is null but the handle has obstacles");
    bind_k<double>(h->kd, *b);
    bind_k<float>(h->kf, *b);
    h->bound = true;
    {   // float obstacle records of whatever the buffers hold now (later writers: dockauv_reset, dockauv_refresh_obstacles)
        DeviceGuard guard(h->device);
        if (h->params.precision == DOCKAUV_F64) CUDA_TRY(launch_refresh_obstacles<double>(h->kd, (cudaStream_t)0));
        else CUDA_TRY(launch_refresh_obstacles<float>(h->kf, (cudaStream_t)0));
        CUDA_TRY(cudaStreamSynchronize((cudaStream_t)0));
    }
    drop_graphs(h);
    return DOCKAUV_OK;
}

// ------------------------------------------------------------------------------------------- launch dispatch
static bool scenario_has_current(int scn) {
    return scn == DOCKAUV_SCN_SIMPLE_CURRENT || scn == DOCKAUV_SCN_CAPSULE_CURRENT || scn == DOCKAUV_SCN_OBSTACLES_CURRENT;
}

static int resolve_layout(const DockauvHandle *h) {
    int layout = h->params.layout;
    // measured on B200 (1M envs of C4): fused kernel 0.91 ms, three-launch pipeline 0.5 ms; without obstacles the pipeline
    // is dynamics + finish only; thread-per-env stays as the independently written cross-check
    if (layout == DOCKAUV_LAYOUT_AUTO) layout = DOCKAUV_LAYOUT_PIPELINE;
    return layout;
}

template <typename T>
static void set_io(KParams<T> &k, const void *actions, bool act_f32, const void *noise, const DockauvStepOut &o,
                   const DockauvDebugOut *d, int auto_reset) {
    k.actions = actions;
    k.noise = (const T *)noise;
    k.obs = o.obs;
    k.terminal_obs = o.terminal_obs;
    k.reward = (T *)o.reward;
    k.ep_return_out = (T *)o.ep_return_out;
    k.done = o.done;
    k.cond_bits = o.cond_bits;
    k.ep_len_out = o.ep_len_out;
    k.delta_d_out = (T *)o.delta_d_out;
    k.auto_reset = auto_reset ? 1 : 0;
    k.act_f32 = act_f32 ? 1 : 0;
    k.dbg_ray_dist = d ? (T *)d->ray_dist : nullptr;
    k.dbg_reward_arr = d ? (T *)d->reward_arr : nullptr;
    k.dbg_euler_dot = d ? (T *)d->euler_dot : nullptr;
    k.dbg_nu_c = d ? (T *)d->nu_c : nullptr;
    k.dbg_nav = d ? (T *)d->nav : nullptr;
    k.dbg_obs = d ? (T *)d->obs_f64 : nullptr;
    k.dbg_state_dot = d ? (T *)d->state_dot : nullptr;
}

static int step_range(DockauvHandle *h, const void *actions, int action_dtype, const void *noise,
                      const DockauvStepOut *out, const DockauvDebugOut *dbg, int auto_reset, int64_t begin,
                      int64_t end, cudaStream_t st, bool with_marks = false) {
    const bool actf32 = action_dtype == DOCKAUV_ACT_F32;
    const int layout = resolve_layout(h);
    // the current is "on" when the scenario spawns one, the caller asked for it, or noise is configured
    const bool has_current = scenario_has_current(h->params.scenario) || h->params.force_current != 0 ||
                             h->params.cur_sigma > 0.0;
    cudaError_t e;
    if (h->params.precision == DOCKAUV_F64) {
        KParams<double> k = h->kd;
        set_io<double>(k, actions, actf32, noise, *out, dbg, auto_reset);
        k.env_begin = begin;
        k.env_end = end;
        k.has_current = has_current;
        e = launch_step<double>(k, h->params.vehicle, layout, st, with_marks ? h->marks : nullptr, with_marks ? &h->n_marks : nullptr);
    } else {
        KParams<float> k = h->kf;
        set_io<float>(k, actions, actf32, noise, *out, dbg, auto_reset);
        k.env_begin = begin;
        k.env_end = end;
        k.has_current = has_current;
        e = launch_step<float>(k, h->params.vehicle, layout, st, with_marks ? h->marks : nullptr, with_marks ? &h->n_marks : nullptr);
    }
    if (e != cudaSuccess) return fail(DOCKAUV_ECUDA, "step kernel launch failed: %s", cudaGetErrorString(e));
    h->last_begins.push_back(begin);
    const bool staged = dbg == nullptr;   // else: the fused kernel
    if (layout == DOCKAUV_LAYOUT_PIPELINE && staged) {
        const int64_t chunk = h->kd.chunk_envs > 0 ? h->kd.chunk_envs : (end - begin);
        h->launches += (int64_t)dockauv::pipe_launches(h->kd) * ((end - begin + chunk - 1) / chunk);
    } else {
        h->launches += 1;
    }
    return DOCKAUV_OK;
}

extern "C" int dockauv_set_seed(DockauvHandle *h, uint64_t seed) {
    if (!h) return fail(DOCKAUV_EINVAL, "null handle");
    h->params.seed = seed;
    h->kd.seed = seed;
    h->kf.seed = seed;
    return DOCKAUV_OK;
}

extern "C" int dockauv_reset(DockauvHandle *h, const uint8_t *mask_dev, void *stream) {
    if (!h) return fail(DOCKAUV_EINVAL, "null handle");
    if (!h->bound) return fail(DOCKAUV_ESTATE, "dockauv_bind must be called before dockauv_reset");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (h->params.precision == DOCKAUV_F64) CUDA_TRY(launch_reset<double>(h->kd, mask_dev, st));
    else CUDA_TRY(launch_reset<float>(h->kf, mask_dev, st));
    h->launches += 1;
    return DOCKAUV_OK;
}

extern "C" int dockauv_refresh_obstacles(DockauvHandle *h, void *stream) {
    if (!h) return fail(DOCKAUV_EINVAL, "null handle");
    if (!h->bound) return fail(DOCKAUV_ESTATE, "dockauv_bind must be called before dockauv_refresh_obstacles");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (h->kd.obsf == nullptr) return DOCKAUV_OK;
    if (h->params.precision == DOCKAUV_F64) CUDA_TRY(launch_refresh_obstacles<double>(h->kd, st));
    else CUDA_TRY(launch_refresh_obstacles<float>(h->kf, st));
    h->launches += 1;
    return DOCKAUV_OK;
}

static int ensure_streams(DockauvHandle *h) {
    for (int s = 0; s < kHostStreams; s++) {
        if (!h->hs[s]) CUDA_TRY(cudaStreamCreateWithFlags(&h->hs[s], cudaStreamNonBlocking));
        if (!h->hev[s]) CUDA_TRY(cudaEventCreateWithFlags(&h->hev[s], cudaEventDisableTiming));
    }
    if (!h->ev_fork) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    return DOCKAUV_OK;
}

#ifndef DOCKAUV_PARTS_MIN_ENVS
#define DOCKAUV_PARTS_MIN_ENVS (1 << 17)   // smallest batch that is stepped in parts (262,144 envs: 0.164 ms in two parts, 0.182 ms in one)
#endif
#ifndef DOCKAUV_STEP_PARTS
#define DOCKAUV_STEP_PARTS 2      // 1 = the whole batch in the caller's stream; at most kHostStreams
#endif
// The whole batch as DOCKAUV_STEP_PARTS parts on internal streams, forked from / joined to the caller's stream by
// events: each launch of one part fills the tail waves of the other parts' launches (a four-launch step has four
// partially filled last waves; in one stream the next launch cannot start before the last CTA of the previous one has
// finished).  Measured at 1M envs: 1 part 0.594 ms, 2 parts 0.566 ms.
static bool steps_in_parts(const DockauvHandle *h) {
    return DOCKAUV_STEP_PARTS > 1 && h->n_envs >= (int64_t)DOCKAUV_PARTS_MIN_ENVS && resolve_layout(h) == DOCKAUV_LAYOUT_PIPELINE &&
           h->params.split_chunk_envs == 0;
}

// obs / terminal_obs rows are written in 16-byte pieces when n_obs is a multiple of 4
static int check_rows(const DockauvHandle *h, const float *obs, const float *terminal_obs) {
    if ((h->n_obs & 3) == 0 && ((((uintptr_t)obs) & 15u) != 0 || (((uintptr_t)terminal_obs) & 15u) != 0))
        return fail(DOCKAUV_EINVAL, "obs and terminal_obs must be 16-byte aligned (n_obs = %d is a multiple of 4: rows are "
                                    "written with 128-bit stores)", h->n_obs);
    return DOCKAUV_OK;
}

static int step_parts(DockauvHandle *h, const void *actions, int action_dtype, const void *noise,
                      const DockauvStepOut *out, int auto_reset, cudaStream_t st) {
    int rc = ensure_streams(h);
    if (rc != DOCKAUV_OK) return rc;
    const int64_t N = h->n_envs;
    const int64_t part = (((N + DOCKAUV_STEP_PARTS - 1) / DOCKAUV_STEP_PARTS + 4095) / 4096) * 4096;   // at most PARTS ranges
    CUDA_TRY(cudaEventRecord(h->ev_fork, st));
    int s = 0;
    for (int64_t b = 0; b < N; b += part, s++) {
        CUDA_TRY(cudaStreamWaitEvent(h->hs[s], h->ev_fork, 0));
        rc = step_range(h, actions, action_dtype, noise, out, nullptr, auto_reset, b, b + part < N ? b + part : N, h->hs[s]);
        if (rc != DOCKAUV_OK) return rc;
        CUDA_TRY(cudaEventRecord(h->hev[s], h->hs[s]));
        CUDA_TRY(cudaStreamWaitEvent(st, h->hev[s], 0));
    }
    return DOCKAUV_OK;
}

// one batched step on the device: in parts for large batches of the pipeline layout, else one launch group
static int step_device(DockauvHandle *h, const void *actions, int action_dtype, const void *noise,
                       const DockauvStepOut *out, const DockauvDebugOut *dbg, int auto_reset, cudaStream_t st, bool timing) {
    if (!timing && dbg == nullptr && steps_in_parts(h)) return step_parts(h, actions, action_dtype, noise, out, auto_reset, st);
    return step_range(h, actions, action_dtype, noise, out, dbg, auto_reset, 0, h->n_envs, st, timing);
}

#ifndef DOCKAUV_STEP_GRAPHS
#define DOCKAUV_STEP_GRAPHS 96     // cached step graphs per handle (a trainer alternates between a few tensors; bench.py's pool: 64)
#endif
// One dockauv_step as a CUDA graph: the launch sequence (both halves, their fork / join events) is captured once per set
// of pointers and replayed -- one graph launch instead of six kernel launches and four event operations, and no
// launch-to-launch gaps on the device (what makes a 65,536-env step 24 us instead of 29).
static int step_graph(DockauvHandle *h, const void *actions, int action_dtype, const void *noise, const DockauvStepOut *out,
                      int auto_reset, cudaStream_t st) {
    DockauvHandle::StepKey key;
    key.actions = actions;
    key.noise = noise;
    key.out = *out;
    key.action_dtype = action_dtype;
    key.auto_reset = auto_reset;
    key.seed = h->params.seed;
    DockauvHandle::StepGraph *g = nullptr;
    for (auto &c : h->sg)
        if (c.key == key) {
            g = &c;
            break;
        }
    if (g == nullptr) {
        // capture on a private stream (the caller's may be the legacy default stream, which cannot be captured)
        cudaStream_t cs = nullptr;
        CUDA_TRY(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
        int rc = DOCKAUV_OK;
        const int64_t before = h->launches;
        h->last_begins.clear();
        if (e == cudaSuccess) {
            rc = step_device(h, actions, action_dtype, noise, out, nullptr, auto_reset, cs, false);
            e = cudaStreamEndCapture(cs, &graph);
        }
        const int64_t n_launch = h->launches - before;
        h->launches = before;       // captured, not launched; every replay counts them
        if (e == cudaSuccess && rc == DOCKAUV_OK) e = cudaGraphInstantiate(&exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        cudaStreamDestroy(cs);
        if (rc != DOCKAUV_OK) return rc;
        if (e != cudaSuccess) return fail(DOCKAUV_ECUDA, "step graph capture failed: %s", cudaGetErrorString(e));
        DockauvHandle::StepGraph fresh;
        fresh.key = key;
        fresh.exec = exec;
        fresh.launches = n_launch;
        fresh.begins = h->last_begins;
        if (h->sg.size() < (size_t)DOCKAUV_STEP_GRAPHS) {
            h->sg.push_back(fresh);
            g = &h->sg.back();
        } else {
            DockauvHandle::StepGraph &victim = h->sg[h->sg_next];
            h->sg_next = (h->sg_next + 1) % h->sg.size();
            cudaGraphExecDestroy(victim.exec);
            victim = fresh;
            g = &victim;
        }
        h->sg_captures += 1;
    }
    CUDA_TRY(cudaGraphLaunch(g->exec, st));
    h->launches += g->launches;
    h->last_begins = g->begins;
    return DOCKAUV_OK;
}

extern "C" int dockauv_step(DockauvHandle *h, const void *actions_dev, int action_dtype, const void *noise_dev,
                            const DockauvStepOut *out, const DockauvDebugOut *dbg, int auto_reset, void *stream) {
    if (!h || !actions_dev || !out) return fail(DOCKAUV_EINVAL, "null argument");
    if (!h->bound) return fail(DOCKAUV_ESTATE, "dockauv_bind must be called before dockauv_step");
    if (!out->obs || !out->reward || !out->done) return fail(DOCKAUV_EINVAL, "obs, reward and done outputs are required");
    if (action_dtype != DOCKAUV_ACT_F32 && action_dtype != DOCKAUV_ACT_F64) return fail(DOCKAUV_EINVAL, "bad action dtype");
    if (check_rows(h, out->obs, out->terminal_obs) != DOCKAUV_OK) return DOCKAUV_EINVAL;
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (h->timing) {
        if (!h->ev0) {
            CUDA_TRY(cudaEventCreate(&h->ev0));
            CUDA_TRY(cudaEventCreate(&h->ev1));
            for (cudaEvent_t &m : h->marks) CUDA_TRY(cudaEventCreate(&m));
        }
        CUDA_TRY(cudaEventRecord(h->ev0, st));
    }
    h->last_begins.clear();
    if (!h->timing && dbg == nullptr && h->sg_enabled) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone)
            return step_graph(h, actions_dev, action_dtype, noise_dev, out, auto_reset, st);
    }
    int rc = step_device(h, actions_dev, action_dtype, noise_dev, out, dbg, auto_reset, st, h->timing);
    if (rc != DOCKAUV_OK) return rc;
    if (h->timing) {
        CUDA_TRY(cudaEventRecord(h->ev1, st));
        h->ev_valid = true;
    }
    return DOCKAUV_OK;
}

extern "C" int dockauv_enable_step_graph(DockauvHandle *h, int enabled) {
    if (!h) return fail(DOCKAUV_EINVAL, "null handle");
    h->sg_enabled = enabled != 0;
    return DOCKAUV_OK;
}

extern "C" int dockauv_step_graph_captures(DockauvHandle *h, int64_t *n) {
    if (!h || !n) return fail(DOCKAUV_EINVAL, "null argument");
    *n = h->sg_captures;
    return DOCKAUV_OK;
}

// ------------------------------------------------------------------------------------------- host-buffer step
static int ensure_host_pipeline(DockauvHandle *h, size_t action_bytes) {
    int rc0 = ensure_streams(h);
    if (rc0 != DOCKAUV_OK) return rc0;
    const size_t esz = h->params.precision == DOCKAUV_F64 ? 8 : 4;
    if (h->st_actions_bytes < action_bytes) {
        if (h->st_actions) CUDA_TRY(cudaFree(h->st_actions));
        h->st_actions = nullptr;
        CUDA_TRY(cudaMalloc(&h->st_actions, action_bytes));
        h->st_actions_bytes = action_bytes;
    }
    if (!h->st_obs) CUDA_TRY(cudaMalloc((void **)&h->st_obs, sizeof(float) * h->n_obs * (size_t)h->n_envs));
    if (!h->st_reward) CUDA_TRY(cudaMalloc(&h->st_reward, esz * (size_t)h->n_envs));
    if (!h->st_done) CUDA_TRY(cudaMalloc((void **)&h->st_done, (size_t)h->n_envs));
    if (!h->st_cond) CUDA_TRY(cudaMalloc((void **)&h->st_cond, (size_t)h->n_envs));
    return DOCKAUV_OK;
}

extern "C" int dockauv_step_host(DockauvHandle *h, const void *actions_host, int action_dtype, float *obs_host,
                                 void *reward_host, uint8_t *done_host, uint8_t *cond_bits_host, int auto_reset,
                                 const DockauvStepOut *aux) {
    if (!h || !actions_host || !obs_host || !reward_host || !done_host) return fail(DOCKAUV_EINVAL, "null argument");
    if (!h->bound) return fail(DOCKAUV_ESTATE, "dockauv_bind must be called before dockauv_step_host");
    if (action_dtype != DOCKAUV_ACT_F32 && action_dtype != DOCKAUV_ACT_F64) return fail(DOCKAUV_EINVAL, "bad action dtype");
    DeviceGuard guard(h->device);
    const int64_t N = h->n_envs;
    const size_t asz = (action_dtype == DOCKAUV_ACT_F32 ? 4 : 8) * (size_t)h->params.n_u;
    const size_t esz = h->params.precision == DOCKAUV_F64 ? 8 : 4;
    int rc = ensure_host_pipeline(h, asz * (size_t)N);
    if (rc != DOCKAUV_OK) return rc;
    h->last_begins.clear();
    // chunks: multiples of 4096 envs, at most 12 per call so copies of chunk c+1 overlap the kernel of chunk c
    int64_t chunk = (N + DOCKAUV_HOST_CHUNKS - 1) / DOCKAUV_HOST_CHUNKS;
    chunk = ((chunk + 4095) / 4096) * 4096;
    if (chunk < 16384) chunk = 16384;
    DockauvStepOut out;
    memset(&out, 0, sizeof(out));
    out.obs = h->st_obs;
    out.reward = h->st_reward;
    out.done = h->st_done;
    out.cond_bits = h->st_cond;
    if (aux) {
        out.terminal_obs = aux->terminal_obs;
        out.ep_return_out = aux->ep_return_out;
        out.ep_len_out = aux->ep_len_out;
        out.delta_d_out = aux->delta_d_out;
        if (check_rows(h, out.obs, out.terminal_obs) != DOCKAUV_OK) return DOCKAUV_EINVAL;
    }
    // Per chunk: actions in, the step's launches, the observation rows out (the bulk of the bytes, one copy).  The
    // small outputs (reward, done, cond_bits) leave once per group of 2 * kHostStreams consecutive chunks (waiting for a
    // group is one event per stream): 36 small copies per step cost ~0.3 ms of copy-engine overhead on top of the
    // 2.6 ms the bytes need.
    // work issued earlier on the legacy default stream (PyTorch's default current stream: a preceding dockauv_step,
    // reset or set_state) is ordered before the internal streams start; callers on other streams synchronise themselves
    CUDA_TRY(cudaEventRecord(h->ev_fork, (cudaStream_t)0));
    for (int s = 0; s < kHostStreams; s++) CUDA_TRY(cudaStreamWaitEvent(h->hs[s], h->ev_fork, 0));
    int c = 0;
    int64_t group_begin = 0;
    auto flush_small = [&](int64_t gb, int64_t ge, int last_stream) -> int {
        cudaStream_t st = h->hs[last_stream];
        for (int s = 0; s < kHostStreams; s++) {
            if (s == last_stream) continue;
            CUDA_TRY(cudaEventRecord(h->hev[s], h->hs[s]));
            CUDA_TRY(cudaStreamWaitEvent(st, h->hev[s], 0));
        }
        CUDA_TRY(cudaMemcpyAsync((char *)reward_host + esz * gb, (char *)h->st_reward + esz * gb, esz * (ge - gb),
                                 cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(done_host + gb, h->st_done + gb, (size_t)(ge - gb), cudaMemcpyDeviceToHost, st));
        if (cond_bits_host)
            CUDA_TRY(cudaMemcpyAsync(cond_bits_host + gb, h->st_cond + gb, (size_t)(ge - gb), cudaMemcpyDeviceToHost, st));
        return DOCKAUV_OK;
    };
    for (int64_t b = 0; b < N; b += chunk, c++) {
        const int64_t e = b + chunk < N ? b + chunk : N;
        const int si = c % kHostStreams;
        cudaStream_t st = h->hs[si];
        CUDA_TRY(cudaMemcpyAsync((char *)h->st_actions + asz * b, (const char *)actions_host + asz * b, asz * (e - b),
                                 cudaMemcpyHostToDevice, st));
        rc = step_range(h, h->st_actions, action_dtype, nullptr, &out, nullptr, auto_reset, b, e, st);
        if (rc != DOCKAUV_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(obs_host + (size_t)h->n_obs * b, h->st_obs + (size_t)h->n_obs * b,
                                 sizeof(float) * h->n_obs * (e - b), cudaMemcpyDeviceToHost, st));
        if ((c + 1) % (2 * kHostStreams) == 0 || e == N) {
            rc = flush_small(group_begin, e, si);
            if (rc != DOCKAUV_OK) return rc;
            group_begin = e;
        }
    }
    for (int s = 0; s < kHostStreams; s++) CUDA_TRY(cudaStreamSynchronize(h->hs[s]));
    return DOCKAUV_OK;
}

// ------------------------------------------------------------------------------------------- stacked rollout
static int rollout_issue(DockauvHandle *h, const void *actions, int action_dtype, int n_steps,
                         const DockauvRolloutOut &o, int auto_reset, cudaStream_t st) {
    const int64_t N = h->n_envs;
    const size_t esz = h->params.precision == DOCKAUV_F64 ? 8 : 4;
    const size_t arow = (action_dtype == DOCKAUV_ACT_F32 ? 4 : 8) * (size_t)h->params.n_u * (size_t)N;
    const size_t orow = (size_t)h->n_obs * (size_t)N;
    auto out_of_step = [&](int t) {
        DockauvStepOut so;
        so.obs = o.obs + orow * t;
        so.reward = (char *)o.reward + esz * (size_t)N * t;
        so.done = o.done + (size_t)N * t;
        so.cond_bits = o.cond_bits ? o.cond_bits + (size_t)N * t : nullptr;
        so.terminal_obs = o.terminal_obs ? o.terminal_obs + orow * t : nullptr;
        so.ep_return_out = o.ep_return_out ? (char *)o.ep_return_out + esz * (size_t)N * t : nullptr;
        so.ep_len_out = o.ep_len_out ? o.ep_len_out + (size_t)N * t : nullptr;
        so.delta_d_out = o.delta_d_out ? (char *)o.delta_d_out + esz * (size_t)N * t : nullptr;
        return so;
    };
    if (steps_in_parts(h)) {
        // envs never interact and the actions are all known: each part of the batch runs its T steps as ONE chain on
        // its own stream, forked from / joined to the caller's stream once per rollout instead of once per step
        int rc = ensure_streams(h);
        if (rc != DOCKAUV_OK) return rc;
        const int64_t part = (((N + DOCKAUV_STEP_PARTS - 1) / DOCKAUV_STEP_PARTS + 4095) / 4096) * 4096;   // at most PARTS ranges
        CUDA_TRY(cudaEventRecord(h->ev_fork, st));
        int s = 0;
        for (int64_t b = 0; b < N; b += part, s++) {
            const int64_t e = b + part < N ? b + part : N;
            CUDA_TRY(cudaStreamWaitEvent(h->hs[s], h->ev_fork, 0));
            for (int t = 0; t < n_steps; t++) {
                const DockauvStepOut so = out_of_step(t);
                rc = step_range(h, (const char *)actions + arow * t, action_dtype, nullptr, &so, nullptr, auto_reset, b, e, h->hs[s]);
                if (rc != DOCKAUV_OK) return rc;
            }
            CUDA_TRY(cudaEventRecord(h->hev[s], h->hs[s]));
            CUDA_TRY(cudaStreamWaitEvent(st, h->hev[s], 0));
        }
        return DOCKAUV_OK;
    }
    for (int t = 0; t < n_steps; t++) {
        const DockauvStepOut so = out_of_step(t);
        int rc = step_range(h, (const char *)actions + arow * t, action_dtype, nullptr, &so, nullptr, auto_reset, 0, N, st);
        if (rc != DOCKAUV_OK) return rc;
    }
    return DOCKAUV_OK;
}

extern "C" int dockauv_rollout(DockauvHandle *h, const void *actions_dev, int action_dtype, int n_steps,
                               const DockauvRolloutOut *out, int auto_reset, int use_graph, void *stream) {
    if (!h || !actions_dev || !out) return fail(DOCKAUV_EINVAL, "null argument");
    if (!h->bound) return fail(DOCKAUV_ESTATE, "dockauv_bind must be called before dockauv_rollout");
    if (!out->obs || !out->reward || !out->done) return fail(DOCKAUV_EINVAL, "obs, reward and done outputs are required");
    if (action_dtype != DOCKAUV_ACT_F32 && action_dtype != DOCKAUV_ACT_F64) return fail(DOCKAUV_EINVAL, "bad action dtype");
    if (n_steps <= 0) return fail(DOCKAUV_EINVAL, "n_steps must be positive");
    if (check_rows(h, out->obs, out->terminal_obs) != DOCKAUV_OK) return DOCKAUV_EINVAL;
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (out->ep_len_out)
        CUDA_TRY(cudaMemsetAsync(out->ep_len_out, 0, sizeof(int32_t) * (size_t)h->n_envs * (size_t)n_steps, st));
    if (!use_graph) return rollout_issue(h, actions_dev, action_dtype, n_steps, *out, auto_reset, st);
    DockauvHandle::RolloutKey key;
    key.actions = actions_dev;
    key.out = *out;
    key.n_steps = n_steps;
    key.action_dtype = action_dtype;
    key.auto_reset = auto_reset;
    key.seed = h->params.seed;
    if (!h->rg_exec || !(key == h->rg_key)) {
        if (h->rg_exec) {
            cudaGraphExecDestroy(h->rg_exec);
            h->rg_exec = nullptr;
        }
        // capture on a private stream (the caller's may be the legacy default stream, which cannot be captured)
        cudaStream_t cs = nullptr;
        CUDA_TRY(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
        int rc = DOCKAUV_OK;
        const int64_t launches_before = h->launches;
        if (e == cudaSuccess) {
            rc = rollout_issue(h, actions_dev, action_dtype, n_steps, *out, auto_reset, cs);
            e = cudaStreamEndCapture(cs, &g);
        }
        h->rg_launches = h->launches - launches_before;
        h->launches = launches_before;      // captured, not launched; every replay below counts them
        if (e == cudaSuccess && rc == DOCKAUV_OK) e = cudaGraphInstantiate(&h->rg_exec, g, 0);
        if (g) cudaGraphDestroy(g);
        cudaStreamDestroy(cs);
        if (rc != DOCKAUV_OK) return rc;
        if (e != cudaSuccess) {
            h->rg_exec = nullptr;
            return fail(DOCKAUV_ECUDA, "rollout graph capture failed: %s", cudaGetErrorString(e));
        }
        h->rg_key = key;
        h->rg_captures += 1;
    }
    CUDA_TRY(cudaGraphLaunch(h->rg_exec, st));
    h->launches += h->rg_launches;
    return DOCKAUV_OK;
}

// ------------------------------------------------------------------------------------------- GAE
template <typename R>
__global__ void gae_kernel(const R *__restrict__ rewards, const float *__restrict__ values,
                           const float *__restrict__ last_values, const uint8_t *__restrict__ dones, int T, int64_t N,
                           float gamma, float lam, float *__restrict__ adv, float *__restrict__ ret) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float next_v = last_values[i], a = 0.0f;
    // rows are env-fastest, so each step is one coalesced request per warp; the next rows are in flight while the
    // recurrence of the current one (3 dependent FMAs) resolves
#pragma unroll 4
    for (int t = T - 1; t >= 0; t--) {
        const int64_t k = (int64_t)t * N + i;
        const float nt = dones[k] ? 0.0f : 1.0f;
        const float v = values[k];
        const float delta = (float)rewards[k] + gamma * next_v * nt - v;
        a = delta + gamma * lam * nt * a;
        adv[k] = a;
        ret[k] = a + v;
        next_v = v;
    }
}

extern "C" int dockauv_gae(const void *rewards_dev, int reward_precision, const float *values_dev,
                           const float *last_values_dev, const uint8_t *dones_dev, int n_steps, int64_t n_envs,
                           float gamma, float gae_lambda, float *advantages_dev, float *returns_dev, void *stream) {
    if (!rewards_dev || !values_dev || !last_values_dev || !dones_dev || !advantages_dev || !returns_dev)
        return fail(DOCKAUV_EINVAL, "null argument");
    if (n_steps <= 0 || n_envs <= 0) return fail(DOCKAUV_EINVAL, "n_steps and n_envs must be positive");
    if (reward_precision != DOCKAUV_F64 && reward_precision != DOCKAUV_F32) return fail(DOCKAUV_EINVAL, "bad reward precision");
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n_envs + threads - 1) / threads);
    if (reward_precision == DOCKAUV_F64)
        gae_kernel<double><<<blocks, threads, 0, st>>>((const double *)rewards_dev, values_dev, last_values_dev, dones_dev,
                                                       n_steps, n_envs, gamma, gae_lambda, advantages_dev, returns_dev);
    else
        gae_kernel<float><<<blocks, threads, 0, st>>>((const float *)rewards_dev, values_dev, last_values_dev, dones_dev,
                                                      n_steps, n_envs, gamma, gae_lambda, advantages_dev, returns_dev);
    CUDA_TRY(cudaGetLastError());
    return DOCKAUV_OK;
}

// ------------------------------------------------------------------------------------------- statistics
extern "C" int dockauv_stats_ptr(DockauvHandle *h, double **stats_dev) {
    if (!h || !stats_dev) return fail(DOCKAUV_EINVAL, "null argument");
    *stats_dev = h->stats;
    return DOCKAUV_OK;
}

__global__ void fold_stats_kernel(double *stats) {
    const int k = threadIdx.x;
    if (k >= DOCKAUV_N_STATS) return;
    // atomic swap / add: a step kernel on another stream may be accumulating while the replicas are folded
    double s = 0.0;
    for (int c = 1; c < DOCKAUV_STAT_COPIES; c++)
        s += __longlong_as_double((long long)atomicExch((unsigned long long *)&stats[c * DOCKAUV_N_STATS + k], 0ull));
    if (s != 0.0) atomicAdd(&stats[k], s);
}

extern "C" int dockauv_fold_stats(DockauvHandle *h, void *stream) {
    if (!h) return fail(DOCKAUV_EINVAL, "null handle");
    DeviceGuard guard(h->device);
    fold_stats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(h->stats);
    CUDA_TRY(cudaGetLastError());
    return DOCKAUV_OK;
}

extern "C" int dockauv_get_stats(DockauvHandle *h, double *stats_host, void *stream) {
    if (!h || !stats_host) return fail(DOCKAUV_EINVAL, "null argument");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = dockauv_fold_stats(h, stream);
    if (rc != DOCKAUV_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(stats_host, h->stats, sizeof(double) * DOCKAUV_N_STATS, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return DOCKAUV_OK;
}

extern "C" int dockauv_clear_stats(DockauvHandle *h, void *stream) {
    if (!h) return fail(DOCKAUV_EINVAL, "null handle");
    DeviceGuard guard(h->device);
    CUDA_TRY(cudaMemsetAsync(h->stats, 0, kStatsBytes, (cudaStream_t)stream));
    return DOCKAUV_OK;
}

extern "C" int dockauv_launch_count(DockauvHandle *h, int64_t *n) {
    if (!h || !n) return fail(DOCKAUV_EINVAL, "null argument");
    *n = h->launches;
    return DOCKAUV_OK;
}

extern "C" int dockauv_last_list_counts(DockauvHandle *h, int64_t *n_listed, int64_t *n_ended, void *stream) {
    if (!h || !n_listed || !n_ended) return fail(DOCKAUV_EINVAL, "null argument");
    *n_listed = *n_ended = 0;
    if (!h->pipe_buf) return DOCKAUV_OK;
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    // one counter pair per env range the most recent step call stepped (halves of the batch, chunks of step_host)
    const size_t n_cnt = 16 * ((size_t)h->n_envs / 128 + 2);
    std::vector<unsigned int> host(n_cnt);
    CUDA_TRY(cudaMemcpyAsync(host.data(), h->kd.view_count, n_cnt * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int64_t b : h->last_begins) {
        // fused cull: words 4..7 of a counter block hold the values at the end of the step (0..3 are zero again); else 0..3
        const size_t k = 16 * (size_t)(b / 128) + (pipe_fuses_cull(h->kd) ? 4 : 0);
        if (k + 3 < n_cnt) {
            *n_listed += (int64_t)host[k] + host[k + 1] + host[k + 2];
            *n_ended += host[k + 3];
        }
    }
    return DOCKAUV_OK;
}

extern "C" int dockauv_rollout_captures(DockauvHandle *h, int64_t *n) {
    if (!h || !n) return fail(DOCKAUV_EINVAL, "null argument");
    *n = h->rg_captures;
    return DOCKAUV_OK;
}

extern "C" int dockauv_enable_timing(DockauvHandle *h, int enabled) {
    if (!h) return fail(DOCKAUV_EINVAL, "null handle");
    h->timing = enabled != 0;
    h->ev_valid = false;
    return DOCKAUV_OK;
}

extern "C" int dockauv_last_step_ms(DockauvHandle *h, float *ms) {
    if (!h || !ms) return fail(DOCKAUV_EINVAL, "null argument");
    if (!h->ev_valid) return fail(DOCKAUV_ESTATE, "no timed step recorded (dockauv_enable_timing + dockauv_step first)");
    DeviceGuard guard(h->device);
    CUDA_TRY(cudaEventSynchronize(h->ev1));
    CUDA_TRY(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return DOCKAUV_OK;
}

extern "C" int dockauv_last_step_launch_ms(DockauvHandle *h, float *ms, int capacity, int *n_launches) {
    if (!h || !ms || !n_launches) return fail(DOCKAUV_EINVAL, "null argument");
    if (!h->ev_valid) return fail(DOCKAUV_ESTATE, "no timed step recorded (dockauv_enable_timing + dockauv_step first)");
    DeviceGuard guard(h->device);
    CUDA_TRY(cudaEventSynchronize(h->ev1));
    const int n = h->n_marks > 0 ? h->n_marks - 1 : 0;
    *n_launches = n;
    for (int k = 0; k < n && k < capacity; k++) CUDA_TRY(cudaEventElapsedTime(&ms[k], h->marks[k], h->marks[k + 1]));
    return DOCKAUV_OK;
}

// ------------------------------------------------------------------------------------------- roofline denominators
template <typename T>
__global__ void fma_peak_kernel(T *out, int iters, T a, T b) {
    T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
            x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
        }
    }
    T s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == (T)123456.789) out[0] = s;   // keeps the chains alive, never true in practice
}

template <typename T>
static int measure_fma(double *tflops) {
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    T *out = nullptr;
    CUDA_TRY(cudaMalloc((void **)&out, sizeof(T)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int threads = 256, blocks = sms * 8, iters = 4096;
    fma_peak_kernel<T><<<blocks, threads>>>(out, 64, (T)1.0000001, (T)1e-9);   // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        fma_peak_kernel<T><<<blocks, threads>>>(out, iters, (T)1.0000001, (T)1e-9);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 64.0 * (double)iters * (double)threads * (double)blocks;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return DOCKAUV_OK;
}

extern "C" int dockauv_measure_peaks(int device, double *fp64_tflops, double *fp32_tflops, double *copy_gbs) {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev)
        return fail(DOCKAUV_ECUDA, "no usable CUDA device %d", device);
    DeviceGuard guard(device);
    int rc;
    if (fp64_tflops && (rc = measure_fma<double>(fp64_tflops)) != DOCKAUV_OK) return rc;
    if (fp32_tflops && (rc = measure_fma<float>(fp32_tflops)) != DOCKAUV_OK) return rc;
    if (copy_gbs) {
        const size_t bytes = (size_t)1 << 30;
        void *a = nullptr, *b = nullptr;
        CUDA_TRY(cudaMalloc(&a, bytes));
        CUDA_TRY(cudaMalloc(&b, bytes));
        CUDA_TRY(cudaMemset(a, 1, bytes));
        cudaEvent_t e0, e1;
        CUDA_TRY(cudaEventCreate(&e0));
        CUDA_TRY(cudaEventCreate(&e1));
        double best = 0.0;
        for (int rep = 0; rep < 6; rep++) {
            CUDA_TRY(cudaEventRecord(e0));
            CUDA_TRY(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice));
            CUDA_TRY(cudaEventRecord(e1));
            CUDA_TRY(cudaEventSynchronize(e1));
            float ms = 0;
            CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
            double gbs = 2.0 * (double)bytes / (ms * 1e-3) / 1e9;
            if (rep > 0 && gbs > best) best = gbs;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaFree(a);
        cudaFree(b);
        *copy_gbs = best;
    }
    return DOCKAUV_OK;
}
