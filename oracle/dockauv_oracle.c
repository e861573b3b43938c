/* dockauv_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see dockauv_oracle.h).
 *
 * CPU restatement of the gym_dockauv step path, one function per reference function, each citing the
 * reference file:line it follows (paths relative to the reference repo root).
 */
#include "dockauv_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.141592653589793

/* np.clip(x, lo, hi) = minimum(maximum(x, lo), hi): propagates NaN */
static double clip(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* ---------------------------------------------------------------- gym_dockauv/utils/geomutils.py */

/* geomutils.py:4-11  ssa: (angle + pi) % (2 pi) - pi, numpy floor-mod on floats */
double orc_ssa(double x) {
    double b = 2.0 * PI;
    double a = x + PI;
    double mod = fmod(a, b);
    if (mod != 0.0) {
        if ((b < 0) != (mod < 0)) mod += b;
    } else {
        mod = copysign(0.0, b);
    }
    return mod - PI;
}

/* geomutils.py:14-43 */
void orc_Rzyx(double phi, double theta, double psi, double R[9]) {
    double cphi = cos(phi), sphi = sin(phi), cth = cos(theta), sth = sin(theta), cpsi = cos(psi), spsi = sin(psi);
    R[0] = cpsi * cth; R[1] = -spsi * cphi + cpsi * sth * sphi; R[2] = spsi * sphi + cpsi * cphi * sth;
    R[3] = spsi * cth; R[4] = cpsi * cphi + sphi * sth * spsi;  R[5] = -cpsi * sphi + sth * spsi * cphi;
    R[6] = -sth;       R[7] = cth * sphi;                       R[8] = cth * cphi;
}

/* geomutils.py:46-75 */
void orc_Tzyx(double phi, double theta, double T[9]) {
    double sphi = sin(phi), tth = tan(theta), cphi = cos(phi), cth = cos(theta);
    T[0] = 1; T[1] = sphi * tth; T[2] = cphi * tth;
    T[3] = 0; T[4] = cphi;       T[5] = -sphi;
    T[6] = 0; T[7] = sphi / cth; T[8] = cphi / cth;
}

/* geomutils.py:106-128 */
static void S_skew(const double a[3], double S[9]) {
    S[0] = 0;     S[1] = -a[2]; S[2] = a[1];
    S[3] = a[2];  S[4] = 0;     S[5] = -a[0];
    S[6] = -a[1]; S[7] = a[0];  S[8] = 0;
}

static void mat3_mul(const double A[9], const double B[9], double C[9]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * k + j];
            C[3 * i + j] = s;
        }
}

static void mat3_vec(const double A[9], const double v[3], double o[3]) {
    for (int i = 0; i < 3; i++) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}

static double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double norm3(const double a[3]) { return sqrt(dot3(a, a)); }

/* ---------------------------------------------------------------- gym_dockauv/objects/statespace.py */

/* statespace.py:199-228 C_RB, :230-276 C_A, :278-286 C = C_RB + C_A */
void orc_C(const OrcParams *p, const double nu[6], double C[36]) {
    const double *nu1 = nu, *nu2 = nu + 3;
    double Snu2[9], SrG[9], t[9], blk[9], Ibnu2[3], a1[3], a2[3];
    memset(C, 0, 36 * sizeof(double));
    S_skew(nu2, Snu2);
    S_skew(p->r_G, SrG);
    /* C_RB[0:3,0:3] = m S(nu2) */
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[6 * i + j] = p->m * Snu2[3 * i + j];
    /* C_RB[0:3,3:6] = -m S(nu2) S(r_G) */
    for (int i = 0; i < 9; i++) t[i] = -p->m * Snu2[i];
    mat3_mul(t, SrG, blk);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[6 * i + 3 + j] = blk[3 * i + j];
    /* C_RB[3:6,0:3] = m S(r_G) S(nu2) */
    for (int i = 0; i < 9; i++) t[i] = p->m * SrG[i];
    mat3_mul(t, Snu2, blk);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[6 * (3 + i) + j] = blk[3 * i + j];
    /* C_RB[3:6,3:6] = -S(I_b nu2) */
    mat3_vec(p->I_b, nu2, Ibnu2);
    S_skew(Ibnu2, blk);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[6 * (3 + i) + 3 + j] = -blk[3 * i + j];
    /* C_A: a1 = M_A11 nu1 + M_A12 nu2, a2 = M_A21 nu1 + M_A22 nu2 */
    for (int i = 0; i < 3; i++) {
        double s11 = 0, s12 = 0, s21 = 0, s22 = 0;
        for (int k = 0; k < 3; k++) {
            s11 += p->M_A[6 * i + k] * nu1[k];
            s12 += p->M_A[6 * i + 3 + k] * nu2[k];
            s21 += p->M_A[6 * (3 + i) + k] * nu1[k];
            s22 += p->M_A[6 * (3 + i) + 3 + k] * nu2[k];
        }
        a1[i] = s11 + s12;
        a2[i] = s21 + s22;
    }
    double Sa1[9], Sa2[9];
    S_skew(a1, Sa1);
    S_skew(a2, Sa2);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            C[6 * i + 3 + j] += -Sa1[3 * i + j];
            C[6 * (3 + i) + j] += -Sa1[3 * i + j];
            C[6 * (3 + i) + 3 + j] += -Sa2[3 * i + j];
        }
}

/* statespace.py:288-351 (diagonal D_L + D_NL); LAUV.py:69-101 (D + D_n + L*|u|).  One generic form:
 * D[i][j] = -(D_lin[i][j] + D_quad[i][j]*|nu_j| + L_lift[i][j]*|nu_0|). */
void orc_D(const OrcParams *p, const double nu[6], double D[36]) {
    double u = fabs(nu[0]);
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            double Dl = -p->D_lin[6 * i + j];
            double Dn = -(p->D_quad[6 * i + j] * fabs(nu[j]));
            double L = -p->L_lift[6 * i + j];
            D[6 * i + j] = Dl + Dn + L * u;
        }
}

/* statespace.py:353-397 */
void orc_G(const OrcParams *p, const double eta[6], double G[6]) {
    double phi = eta[3], theta = eta[4];
    double W = p->W, BY = p->BY;
    double xG = p->r_G[0], yG = p->r_G[1], zG = p->r_G[2], xB = p->r_B[0], yB = p->r_B[1], zB = p->r_B[2];
    G[0] = (W - BY) * sin(theta);
    G[1] = -(W - BY) * cos(theta) * sin(phi);
    G[2] = -(W - BY) * cos(theta) * cos(phi);
    G[3] = -(yG * W - yB * BY) * cos(theta) * cos(phi) + (zG * W - zB * BY) * cos(theta) * sin(phi);
    G[4] = (zG * W - zB * BY) * sin(theta) + (xG * W - xB * BY) * cos(theta) * cos(phi);
    G[5] = -(xG * W - xB * BY) * cos(theta) * sin(phi) - (yG * W - yB * BY) * sin(theta);
}

/* BlueROV2.py:34-43,74-75 (constant B); LAUV.py:59-67 (B(nu)) */
void orc_B(const OrcParams *p, const double nu[6], double B[6 * ORC_MAX_U]) {
    int n_u = p->n_u;
    if (p->vehicle == 0) {
        memcpy(B, p->B_const, sizeof(double) * 6 * n_u);
    } else {
        double u2 = nu[0] * nu[0];
        memset(B, 0, sizeof(double) * 6 * n_u);
        B[0 * 3 + 0] = 1;
        B[1 * 3 + 1] = p->lauv_B[0] * u2;  /* Y_uudr */
        B[2 * 3 + 2] = p->lauv_B[1] * u2;  /* Z_uuds */
        B[4 * 3 + 2] = p->lauv_B[2] * u2;  /* M_uuds */
        B[5 * 3 + 1] = p->lauv_B[3] * u2;  /* N_uudr */
    }
}

/* ---------------------------------------------------------------- gym_dockauv/objects/auvsim.py */

/* auvsim.py:110-160 */
void orc_state_dot(const OrcParams *p, const double y[12], const double u[], const double nu_c[6], double out[12]) {
    const double *eta = y, *nu_r = y + 6;
    double R[9], T[9], v[6];
    orc_Rzyx(eta[3], eta[4], eta[5], R);
    orc_Tzyx(eta[3], eta[4], T);
    for (int i = 0; i < 6; i++) v[i] = nu_r[i] + nu_c[i];
    /* geom.J(eta).dot(nu_r + nu_c), geomutils.py:78-103: block diag(R, T) */
    for (int i = 0; i < 3; i++) {
        out[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
        out[3 + i] = T[3 * i] * v[3] + T[3 * i + 1] * v[4] + T[3 * i + 2] * v[5];
    }
    double B[6 * ORC_MAX_U], D[36], C[36], G[6], rhs[6];
    orc_B(p, nu_r, B);
    orc_D(p, nu_r, D);
    orc_C(p, nu_r, C);
    orc_G(p, eta, G);
    for (int i = 0; i < 6; i++) {
        double bu = 0, dn = 0, cn = 0;
        for (int k = 0; k < p->n_u; k++) bu += B[i * p->n_u + k] * u[k];
        for (int k = 0; k < 6; k++) {
            dn += D[6 * i + k] * nu_r[k];
            cn += C[6 * i + k] * nu_r[k];
        }
        rhs[i] = bu - dn - cn - G[i];
    }
    for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int k = 0; k < 6; k++) s += p->M_inv[6 * i + k] * rhs[k];
        out[6 + i] = s;
    }
}

/* auvsim.py:67-75.  With a float32 action array numpy keeps clip, +1 and /2 in float32 (python scalars
 * are weak), and only the product with the float64 u_bound promotes to float64. */
void orc_unnormalize(const OrcParams *p, const void *action, int action_is_f32, double x[]) {
    for (int i = 0; i < p->n_u; i++) {
        double frac;
        if (action_is_f32) {
            float a = ((const float *)action)[i];
            float c = a < -1.0f ? -1.0f : (a > 1.0f ? 1.0f : a);
            float f = (c + 1.0f) / 2.0f;
            frac = (double)f;
        } else {
            double a = ((const double *)action)[i];
            frac = (clip(a, -1.0, 1.0) + 1.0) / 2.0;
        }
        x[i] = p->u_lo[i] + (p->u_hi[i] - p->u_lo[i]) * frac;
    }
}

/* utils/odesolver45.py:5-28 (all six stages, 4th-order result w kept: auvsim.py:98) */
static void odesolver45(const OrcParams *p, const double y[12], double h, const double u[], const double nu_c[6],
                        double w[12]) {
    double s1[12], s2[12], s3[12], s4[12], s5[12], s6[12], yt[12];
    orc_state_dot(p, y, u, nu_c, s1);
    for (int i = 0; i < 12; i++) yt[i] = y[i] + h * s1[i] / 4.0;
    orc_state_dot(p, yt, u, nu_c, s2);
    for (int i = 0; i < 12; i++) yt[i] = y[i] + 3.0 * h * s1[i] / 32.0 + 9.0 * h * s2[i] / 32.0;
    orc_state_dot(p, yt, u, nu_c, s3);
    for (int i = 0; i < 12; i++)
        yt[i] = y[i] + 1932.0 * h * s1[i] / 2197.0 - 7200.0 * h * s2[i] / 2197.0 + 7296.0 * h * s3[i] / 2197.0;
    orc_state_dot(p, yt, u, nu_c, s4);
    for (int i = 0; i < 12; i++)
        yt[i] = y[i] + 439.0 * h * s1[i] / 216.0 - 8.0 * h * s2[i] + 3680.0 * h * s3[i] / 513.0 -
                845.0 * h * s4[i] / 4104.0;
    orc_state_dot(p, yt, u, nu_c, s5);
    for (int i = 0; i < 12; i++)
        yt[i] = y[i] - 8.0 * h * s1[i] / 27.0 + 2 * h * s2[i] - 3544.0 * h * s3[i] / 2565 +
                1859.0 * h * s4[i] / 4104.0 - 11.0 * h * s5[i] / 40.0;
    orc_state_dot(p, yt, u, nu_c, s6);  /* feeds only the discarded 5th-order result q */
    (void)s6;
    for (int i = 0; i < 12; i++)
        w[i] = y[i] + h * (25.0 * s1[i] / 216.0 + 1408.0 * s3[i] / 2565.0 + 2197.0 * s4[i] / 4104.0 - s5[i] / 5.0);
}

/* auvsim.py:77-108 step + _sim; utils/lowpassfilter.py:29-42 */
void orc_auv_step(const OrcParams *p, double state[12], double u[], const void *action, int action_is_f32,
                  const double nu_c[6], double state_dot[12]) {
    double x[ORC_MAX_U], w[12];
    orc_unnormalize(p, action, action_is_f32, x);
    for (int i = 0; i < p->n_u; i++) u[i] = p->lp_alpha * x[i] + (1 - p->lp_alpha) * u[i];
    odesolver45(p, state, p->h, u, nu_c, w);
    memcpy(state, w, sizeof(w));
    for (int i = 3; i < 6; i++) state[i] = orc_ssa(state[i]);
    orc_state_dot(p, state, u, nu_c, state_dot);
}

/* ---------------------------------------------------------------- gym_dockauv/objects/current.py */

/* current.py:33-76 */
void orc_current_nu_c(const double cur[5], const double att[3], double nu_c[6]) {
    double Vc = cur[0], alpha = cur[1], beta = cur[2];
    double vn[3] = {Vc * cos(alpha) * cos(beta), Vc * sin(beta), Vc * sin(alpha) * cos(beta)};
    double R[9];
    orc_Rzyx(att[0], att[1], att[2], R);
    for (int i = 0; i < 3; i++) nu_c[i] = R[i] * vn[0] + R[3 + i] * vn[1] + R[6 + i] * vn[2]; /* R^T v */
    nu_c[3] = nu_c[4] = nu_c[5] = 0;
}

/* current.py:78-96 */
static void current_sim(const OrcParams *p, double cur[5], double w) {
    double Vc_dot = -p->cur_mu * cur[0] + w;
    cur[0] += Vc_dot * p->h;
    cur[0] = clip(cur[0], cur[3], cur[4]);
}

/* ---------------------------------------------------------------- gym_dockauv/objects/shape.py */

/* shape.py:327-390, one ray of the vectorised version (the one the env calls, docking3d.py:424-430) */
double orc_ray_capsule(const double l1[3], const double ld[3], const double cap1[3], const double cap2[3],
                       double cap_rad) {
    double ba[3], oa[3], rd[3], oc2[3];
    double n = norm3(ld);
    for (int i = 0; i < 3; i++) {
        ba[i] = cap2[i] - cap1[i];
        oa[i] = l1[i] - cap1[i];
        rd[i] = ld[i] / n;
        oc2[i] = l1[i] - cap2[i];
    }
    double baba = dot3(ba, ba), bard = dot3(rd, ba), baoa = dot3(oa, ba), rdoa = dot3(rd, oa), oaoa = dot3(oa, oa);
    double a = baba - bard * bard;
    double b = baba * rdoa - baoa * bard;
    double c = baba * oaoa - baoa * baoa - cap_rad * cap_rad * baba;
    double h = b * b - a * c;
    double res = 0.0;
    int mask_h = h >= 0;
    double t = mask_h ? (-b - sqrt(h)) / a : -INFINITY;
    double y = baoa + t * bard;
    int mask_body = mask_h && (y > 0) && (y < baba);
    if (mask_body) res = t;
    double oc[3] = {0, 0, 0};
    if (y <= 0.0) memcpy(oc, oa, sizeof(oc));
    if (y >= 0.0) memcpy(oc, oc2, sizeof(oc));
    double b2 = dot3(rd, oc);
    double c2 = dot3(oc, oc) - cap_rad * cap_rad;
    double h2 = b2 * b2 - c2;
    int mask_caps = mask_h && (h2 > 0.0) && !mask_body;
    if (mask_caps) res = -b2 - sqrt(h2);
    if ((h <= 0) || (res == 0)) res = -INFINITY;
    return res;
}

/* shape.py:235-264, one ray against n spheres: per sphere min(-b+h, -b-h) with h = sqrt(b^2-c) or -inf,
 * then the closest positive value, else the first sphere's value (argmin of an all-inf row is 0). */
double orc_ray_spheres(const double l1[3], const double ld[3], const double *centres, const double *rads, int n) {
    double rd[3];
    double nn = norm3(ld);
    for (int i = 0; i < 3; i++) rd[i] = ld[i] / nn;
    double best = INFINITY, first = 0;
    int best_idx = 0;
    double resv[ORC_MAX_SPH];
    for (int s = 0; s < n; s++) {
        double oc[3] = {l1[0] - centres[3 * s], l1[1] - centres[3 * s + 1], l1[2] - centres[3 * s + 2]};
        double b = dot3(oc, rd);
        double nrm = norm3(oc);
        double c = nrm * nrm - rads[s] * rads[s];
        double h = b * b - c;
        if (h < 0.0) h = -INFINITY; else h = sqrt(h);
        double r1 = -b + h, r2 = -b - h;
        double res = r1 < r2 ? r1 : r2;   /* np.minimum */
        if (isnan(r1) || isnan(r2)) res = NAN;
        resv[s] = res;
        double key = (res > 0) ? res : INFINITY;
        if (key < best) { best = key; best_idx = s; }
    }
    (void)first;
    return resv[best_idx];
}

/* shape.py:393-417 */
double orc_dist_line_point(const double po[3], const double l1[3], const double l2[3]) {
    double d[3], l[3] = {l2[0] - l1[0], l2[1] - l1[1], l2[2] - l1[2]};
    double n = norm3(l);
    for (int i = 0; i < 3; i++) d[i] = l[i] / n;
    double a[3] = {l1[0] - po[0], l1[1] - po[1], l1[2] - po[2]};
    double b[3] = {po[0] - l2[0], po[1] - l2[1], po[2] - l2[2]};
    double s = dot3(a, d), t = dot3(b, d);
    double h = s;                 /* np.maximum.reduce([s, t, 0]) */
    if (t > h || isnan(t)) h = t;
    if (0 > h) h = 0;
    double q[3] = {po[0] - l1[0], po[1] - l1[1], po[2] - l1[2]};
    double c[3] = {q[1] * d[2] - q[2] * d[1], q[2] * d[0] - q[0] * d[2], q[0] * d[1] - q[1] * d[0]};
    return hypot(h, norm3(c));
}

/* shape.py:195-210 */
int orc_collision_capsule_sphere(const double c1[3], const double c2[3], double cr, const double sp[3], double sr) {
    return orc_dist_line_point(sp, c1, c2) <= cr + sr;
}

/* shape.py:182-192 */
int orc_collision_sphere_spheres(const double p1[3], double r1, const double *p2, const double *r2, int n) {
    int any = 0;
    for (int s = 0; s < n; s++) {
        double d[3] = {p2[3 * s] - p1[0], p2[3 * s + 1] - p1[1], p2[3 * s + 2] - p1[2]};
        if (norm3(d) <= r1 + r2[s]) any = 1;
    }
    return any;
}

/* ---------------------------------------------------------------- gym_dockauv/objects/sensor.py */

/* sensor.py:131-137 + skimage.measure.block_reduce(block, np.max, cval=0): zero-pad to a multiple of the
 * block size, max over each block, flatten row-major. */
void orc_block_reduce_max(const double *d, int n_v, int n_h, int block, double *out) {
    int ov = (n_v + block - 1) / block, oh = (n_h + block - 1) / block;
    for (int i = 0; i < ov; i++)
        for (int j = 0; j < oh; j++) {
            double m = -INFINITY;
            int nan = 0;
            for (int di = 0; di < block; di++)
                for (int dj = 0; dj < block; dj++) {
                    int r = i * block + di, c = j * block + dj;
                    double v = (r < n_v && c < n_h) ? d[r * n_h + c] : 0.0;
                    if (isnan(v)) nan = 1;
                    if (v > m) m = v;
                }
            out[i * oh + j] = nan ? NAN : m;
        }
}

/* ---------------------------------------------------------------- gym_dockauv/envs/docking3d.py */

/* docking3d.py:712-723 */
double orc_log_precision(double x, double x_goal, double x_max) {
    double eps = 0.001;
    double xx = x > eps ? x : eps;            /* python max(x, eps): returns x unless eps > x; NaN stays */
    if (isnan(x)) xx = x;
    double gg = x_goal > eps ? x_goal : eps;
    return 1 - clip(log(xx / x_max) / log(gg / x_max), 0, 1);
}

/* docking3d.py:742-765 with the fixed arguments used at :523-548,:571-582 (x_des=0, exps=4, no reversal) */
static double cont_goal_constraints(double x, double delta_d, double x_des, double delta_d_des, double x_max,
                                    double delta_d_max, double x_exp, double delta_d_exp) {
    double r_x = pow(fabs(0.0 - orc_log_precision(x, x_des, x_max)), x_exp);
    double r_dd = pow(fabs(0.0 - orc_log_precision(delta_d, delta_d_des, delta_d_max)), delta_d_exp);
    return r_x * r_dd;
}

/* docking3d.py:767-792 with gamma_c=1, epsilon_c=0.001 (call site :560-563); beta_oa is precomputed */
double orc_obstacle_avoidance(const OrcParams *p, const double *d) {
    double sum_beta = 0, dotv = 0;
    for (int i = 0; i < p->n_rays; i++) sum_beta += p->beta_oa[i];
    for (int i = 0; i < p->n_rays; i++) {
        double c = clip(1 - d[i] / p->radar_max_dist, 0, 1);
        double q = (1.0 * (1 - c)) * (1.0 * (1 - c));
        double mx = q > 0.001 ? q : 0.001;   /* np.maximum */
        if (isnan(q)) mx = q;
        dotv += mx * p->beta_oa[i];
    }
    return sum_beta / dotv - 1;
}

/* docking3d.py:346-402 */
void orc_step(const OrcParams *p, OrcEnv *e, const void *action, int action_is_f32, double noise_w, OrcStepOut *o) {
    /* :348-349 current.sim(); nu_c from the pre-step attitude */
    current_sim(p, e->cur, noise_w);
    orc_current_nu_c(e->cur, e->state + 3, o->nu_c);
    /* :352 auv.step(action, current(attitude)) */
    orc_auv_step(p, e->state, e->u, action, action_is_f32, o->nu_c, o->state_dot);

    const double *pos = e->state, *att = e->state + 3;
    /* :355 radar.update, sensor.py:90-102 */
    double R[9];
    orc_Rzyx(att[0], att[1], att[2], R);
    int n_r = p->n_rays;
    static _Thread_local double rd_n[ORC_MAX_RAYS * 3];
    for (int i = 0; i < n_r; i++) {
        double v[3];
        mat3_vec(R, p->rd_b + 3 * i, v);
        double n = norm3(v);
        for (int k = 0; k < 3; k++) rd_n[3 * i + k] = v[k] / n;
    }
    /* :356 update_radar_collision, :415-442 */
    int n_cols = e->n_caps + (e->n_sph > 0 ? 1 : 0);
    for (int i = 0; i < n_r; i++) {
        double d;
        if (n_cols == 0) {
            d = p->radar_max_dist;            /* i_dist None -> fallback, sensor.py:113-114 */
        } else {
            double col[ORC_MAX_CAPS + 1];
            int nc = 0;
            for (int k = 0; k < e->n_caps; k++)
                col[nc++] = orc_ray_capsule(pos, rd_n + 3 * i, e->caps[k], e->caps[k] + 3, e->caps[k][6]);
            if (e->n_sph > 0) {
                double cen[ORC_MAX_SPH * 3], rad[ORC_MAX_SPH];
                for (int s = 0; s < e->n_sph; s++) {
                    memcpy(cen + 3 * s, e->sph[s], 3 * sizeof(double));
                    rad[s] = e->sph[s][3];
                }
                col[nc++] = orc_ray_spheres(pos, rd_n + 3 * i, cen, rad, e->n_sph);
            }
            /* :439 i_dist[np.where(i_dist > 0, i_dist, inf).argmin(axis=1)] */
            double best = INFINITY;
            int bi = 0;
            for (int k = 0; k < nc; k++) {
                double key = col[k] > 0 ? col[k] : INFINITY;
                if (key < best) { best = key; bi = k; }
            }
            d = col[bi];
            /* :357 radar.update_intersec, sensor.py:117 */
            if (d < 0 || d > p->radar_max_dist) d = p->radar_max_dist;
        }
        o->ray_dist[i] = d;
    }
    /* :360 update_body_collision, :444-460 */
    int col = 0;
    if (e->n_sph > 0) {
        double cen[ORC_MAX_SPH * 3], rad[ORC_MAX_SPH];
        for (int s = 0; s < e->n_sph; s++) {
            memcpy(cen + 3 * s, e->sph[s], 3 * sizeof(double));
            rad[s] = e->sph[s][3];
        }
        col |= orc_collision_sphere_spheres(pos, p->safety_radius, cen, rad, e->n_sph);
    }
    for (int k = 0; k < e->n_caps; k++)
        col |= orc_collision_capsule_sphere(e->caps[k], e->caps[k] + 3, e->caps[k][6], pos, p->safety_radius);
    o->collision = (uint8_t)col;

    /* :371 update_navigation_errors, :404-413 */
    double diff[3] = {e->goal[0] - pos[0], e->goal[1] - pos[1], e->goal[2] - pos[2]};
    o->delta_d = norm3(diff);
    o->delta_theta = att[1] + orc_ssa(atan2(diff[2], sqrt(diff[0] * diff[0] + diff[1] * diff[1])));
    o->delta_psi = orc_ssa(atan2(diff[1], diff[0]) - att[2]);
    o->delta_heading_goal = orc_ssa(e->heading_goal - att[2]);

    /* :374 observe, :462-488 (float64 expressions stored into a float32 array) */
    const double *nu_r = e->state + 6;
    float *obs = o->obs;
    obs[0] = (float)clip(1 - (log(o->delta_d / p->max_dist_from_goal) /
                              log(p->dist_goal_reached_tol / p->max_dist_from_goal)), 0, 1);
    obs[1] = (float)clip(o->delta_theta / (PI / 2), -1, 1);
    obs[2] = (float)clip(o->delta_psi / PI, -1, 1);
    obs[3] = (float)clip(nu_r[0] / p->u_max, -1, 1);
    obs[4] = (float)clip(nu_r[1] / p->v_max, -1, 1);
    obs[5] = (float)clip(nu_r[2] / p->w_max, -1, 1);
    obs[6] = (float)clip(att[0] / p->max_attitude, -1, 1);
    obs[7] = (float)clip(att[1] / p->max_attitude, -1, 1);
    obs[8] = (float)clip(sin(att[2]), -1, 1);
    obs[9] = (float)clip(cos(att[2]), -1, 1);
    obs[10] = (float)clip(nu_r[3] / p->p_max, -1, 1);
    obs[11] = (float)clip(nu_r[4] / p->q_max, -1, 1);
    obs[12] = (float)clip(nu_r[5] / p->r_max, -1, 1);
    obs[13] = (float)clip(o->nu_c[0] / 2, -1, 1);
    obs[14] = (float)clip(o->nu_c[1] / 2, -1, 1);
    obs[15] = (float)clip(o->nu_c[2] / 2, -1, 1);
    {
        double red[ORC_MAX_RAYS];
        orc_block_reduce_max(o->ray_dist, p->n_vert, p->n_horiz, p->block, red);
        for (int i = 0; i < p->n_rays_reduced; i++) obs[16 + i] = (float)clip(red[i] / p->radar_max_dist, 0, 1);
    }

    /* :377 is_done, :597-631 (t_steps is the value BEFORE the increment at :385) */
    o->cond[0] = o->delta_d < p->dist_goal_reached_tol;
    o->cond[1] = o->delta_d > p->max_dist_from_goal;
    o->cond[2] = (fabs(att[0]) > p->max_attitude) || (fabs(att[1]) > p->max_attitude);
    o->cond[3] = e->t_steps >= p->max_timesteps;
    o->cond[4] = o->collision;
    o->goal_reached = o->cond[0];
    o->done = o->cond[0] | o->cond[1] | o->cond[2] | o->cond[3] | o->cond[4];

    /* :380 reward_step, :490-595 */
    double *r = o->reward_arr;
    r[0] = -p->w_d * orc_log_precision(o->delta_d, p->dist_goal_reached_tol, p->max_dist_from_goal);
    if (p->reward_set == 1) {
        r[1] = -p->w_delta_theta * ((o->delta_theta / (PI / 2)) * (o->delta_theta / (PI / 2)));
        r[2] = -p->w_delta_psi * ((o->delta_psi / PI) * (o->delta_psi / PI));
    } else {
        r[1] = -p->w_delta_theta * cont_goal_constraints(fabs(o->delta_theta), o->delta_d, 0.0,
                                                         p->dist_goal_reached_tol, PI / 2, p->max_dist_from_goal, 4, 4);
        r[2] = -p->w_delta_psi * cont_goal_constraints(fabs(o->delta_psi), o->delta_d, 0.0,
                                                       p->dist_goal_reached_tol, PI, p->max_dist_from_goal, 4, 4);
    }
    r[3] = -p->w_phi * ((att[0] / (PI / 2)) * (att[0] / (PI / 2)));
    r[4] = -p->w_theta * ((att[1] / (PI / 2)) * (att[1] / (PI / 2)));
    {
        double nrm = norm3(o->state_dot + 3) / p->p_max;
        r[5] = -p->w_Thetadot * (nrm * nrm);
    }
    {
        double roa = orc_obstacle_avoidance(p, o->ray_dist);
        if (p->reward_set == 1) r[6] = -p->w_oa * roa;
        else r[6] = -p->w_oa * cont_goal_constraints(fabs(roa), o->delta_d, 0.0, p->dist_goal_reached_tol, 1.0,
                                                     p->max_dist_from_goal, 4, 4);
    }
    /* :584-585  -(sum((|a| / n_u)**2 * w_a)) with the RAW action; float32 arithmetic if the action array is
     * float32 and the factor is a python scalar (weak); float64 if the factor is a float64 array */
    if (action_is_f32 && p->action_factor_is_scalar) {
        float s = 0.0f;
        for (int i = 0; i < p->n_u; i++) {
            float a = fabsf(((const float *)action)[i]) / (float)p->n_u;
            s += (a * a) * (float)p->action_reward_factors[i];
        }
        r[7] = -(double)s;
    } else {
        double s = 0;
        for (int i = 0; i < p->n_u; i++) {
            double q;
            if (action_is_f32) {
                float a = fabsf(((const float *)action)[i]) / (float)p->n_u;
                q = (double)(a * a);
            } else {
                double a = fabs(((const double *)action)[i]) / p->n_u;
                q = a * a;
            }
            s += q * p->action_reward_factors[i];
        }
        r[7] = -s;
    }
    for (int k = 0; k < 5; k++) r[8 + k] = o->cond[k] * p->w_done[k];
    double sum = 0;
    for (int k = 0; k < ORC_N_REWARDS; k++) sum += r[k];
    o->reward = sum;
    e->cum_reward += sum;
    /* :384-385 */
    e->t_steps += 1;
}

/* ---------------------------------------------------------------- counter-based reset (ours, see header) */

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

/* uniform double in [0,1) number `idx` of the stream (seed, env_id, episode) */
static double uni(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t idx) {
    uint32_t c[4] = {(uint32_t)env_id, (uint32_t)(env_id >> 32), episode, idx >> 1};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint32_t hi = (idx & 1) ? c[2] : c[0], lo = (idx & 1) ? c[3] : c[1];
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) / 9007199254740992.0;
}

static double sign(double x) { return (x > 0) - (x < 0); }

void orc_reset_env(const OrcParams *p, OrcEnv *e, int scenario, uint64_t seed, uint64_t env_id) {
    int n_extra_sph = (scenario >> 8) & 0xff;
    scenario &= 0xff;
    uint32_t ep = (uint32_t)e->episode;
    e->episode += 1;
    memset(e->state, 0, sizeof(e->state));          /* auvsim.py:55-65 */
    memset(e->u, 0, sizeof(e->u));
    e->t_steps = 0;
    e->cum_reward = 0;
    e->n_caps = 0;
    e->n_sph = 0;
    /* SimpleDocking3d.generate_environment, docking3d.py:803-825 */
    e->goal[0] = e->goal[1] = e->goal[2] = 0;
    e->heading_goal = (uni(seed, env_id, ep, 0) - 0.5) * PI;
    {   /* generate_random_pos(d=15), :687-696 */
        double r[3] = {uni(seed, env_id, ep, 1) - 0.5, uni(seed, env_id, ep, 2) - 0.5, uni(seed, env_id, ep, 3) - 0.5};
        r[2] = fabs(r[0] + r[1]) / 3 * sign(r[2]);
        double s = 15.0 / norm3(r);
        for (int i = 0; i < 3; i++) e->state[i] = e->goal[i] + r[i] * s;
    }
    {   /* generate_random_att(0.7), :698-703 */
        double f[3] = {p->max_attitude * 0.7, p->max_attitude * 0.7, PI};
        for (int i = 0; i < 3; i++) e->state[3 + i] = (uni(seed, env_id, ep, 4 + i) - 0.5) * 2 * f[i];
    }
    e->cur[0] = 0; e->cur[1] = 0; e->cur[2] = 0; e->cur[3] = 0; e->cur[4] = 0;
    int has_capsule = scenario >= 2, has_pillars = scenario >= 4;
    if (has_capsule) {   /* CapsuleDocking3d, :860-886 */
        double theta = uni(seed, env_id, ep, 7) * 2 * PI;
        double radius = 1.0 + p->safety_radius;
        e->goal[0] = cos(theta) * radius;
        e->goal[1] = sin(theta) * radius;
        e->goal[2] = (uni(seed, env_id, ep, 8) - 0.5) * 4.0;
        double cap[7] = {0, 0, 2.0, 0, 0, -2.0, 1.0};   /* vec_bot = 2*position - vec_top, shape.py:105-108 */
        memcpy(e->caps[e->n_caps++], cap, sizeof(cap));
        /* vec_line_point(goal, top, bot), shape.py:420-433, then ssa(atan2(vec_y, vec_x)) */
        double dv[3] = {0, 0, 1.0};
        double v[3] = {e->goal[0] - 0, e->goal[1] - 0, e->goal[2] - (-2.0)};
        double t = dot3(v, dv);
        double pro[3] = {0 + t * dv[0], 0 + t * dv[1], -2.0 + t * dv[2]};
        e->heading_goal = orc_ssa(atan2(pro[1] - e->goal[1], pro[0] - e->goal[0]));
    }
    if (has_pillars) {   /* ObstaclesDocking3d, :919-946 */
        double theta = uni(seed, env_id, ep, 9) * 2 * PI;
        double half = 2 * p->max_dist_from_goal / 2.0;
        for (int i = 0; i < 4; i++) {
            double x = cos(theta) * 6, y = sin(theta) * 6;
            theta += 2 * PI / 4;
            double cap[7] = {x, y, half, x, y, -half, 1.0};
            memcpy(e->caps[e->n_caps++], cap, sizeof(cap));
        }
    }
    if (scenario == 6) {  /* ObstaclesNoCapDocking3d, :957-965: pop the dock capsule */
        memmove(e->caps[0], e->caps[1], sizeof(e->caps[0]) * 4);
        e->n_caps -= 1;
    }
    if (scenario == 1 || scenario == 3 || scenario == 5) {   /* *Current*, :837-849, :897-908, :977-988 */
        e->cur[1] = (uni(seed, env_id, ep, 10) - 0.5) * 2 * (PI / 2);
        e->cur[2] = (uni(seed, env_id, ep, 11) - 0.5) * 2 * PI;
        double speed = scenario == 1 ? uni(seed, env_id, ep, 12) * 1.0 : 0.5;
        e->cur[0] = 0.5;
        e->cur[3] = e->cur[4] = speed;
    }
    /* extension used by the BASELINE C4 workload: synthetic unit spheres, centres uniform in direction,
     * radius U[4, 10] from the origin */
    for (int s = 0; s < n_extra_sph && s < ORC_MAX_SPH; s++) {
        double z = 2 * uni(seed, env_id, ep, 13 + 3 * s) - 1;
        double az = 2 * PI * uni(seed, env_id, ep, 14 + 3 * s);
        double rr = 4.0 + 6.0 * uni(seed, env_id, ep, 15 + 3 * s);
        double q = sqrt(1 - z * z);
        e->sph[s][0] = rr * q * cos(az);
        e->sph[s][1] = rr * q * sin(az);
        e->sph[s][2] = rr * z;
        e->sph[s][3] = 1.0;
        e->n_sph++;
    }
}

/* env_ids == NULL: env i has the global id env_id0 + i (a contiguous shard); else env i has the id env_ids[i] (an
 * arbitrary sample of a larger batch: the Philox reset stream is keyed by the global id, so any subset can be
 * followed on its own). */
int64_t orc_step_batch_ids(const OrcParams *p, OrcEnv *envs, int64_t n, const void *actions, int action_is_f32,
                           int scenario, uint64_t seed, uint64_t env_id0, const uint64_t *env_ids, float *obs,
                           double *reward, uint8_t *done, uint8_t *cond_bits, int n_threads) {
    int64_t finished = 0;
    size_t astride = (size_t)p->n_u * (action_is_f32 ? 4 : 8);
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for num_threads(n_threads) schedule(static) reduction(+ : finished)
#endif
    for (int64_t i = 0; i < n; i++) {
        OrcStepOut o;
        orc_step(p, &envs[i], (const char *)actions + astride * i, action_is_f32, 0.0, &o);
        if (obs) {
            if (o.done) memset(obs + (size_t)p->n_obs * i, 0, sizeof(float) * p->n_obs);
            else memcpy(obs + (size_t)p->n_obs * i, o.obs, sizeof(float) * p->n_obs);
        }
        if (reward) reward[i] = o.reward;
        if (done) done[i] = o.done;
        if (cond_bits) {   /* bit k = done condition k, docking3d.py:606-617 */
            uint8_t b = 0;
            for (int k = 0; k < 5; k++) b |= (uint8_t)((o.cond[k] ? 1 : 0) << k);
            cond_bits[i] = b;
        }
        if (o.done) {
            finished++;
            orc_reset_env(p, &envs[i], scenario, seed, env_ids ? env_ids[i] : env_id0 + (uint64_t)i);
        }
    }
    return finished;
}

int64_t orc_step_batch(const OrcParams *p, OrcEnv *envs, int64_t n, const void *actions, int action_is_f32,
                       int scenario, uint64_t seed, uint64_t env_id0, float *obs, double *reward, uint8_t *done,
                       int n_threads) {
    return orc_step_batch_ids(p, envs, n, actions, action_is_f32, scenario, seed, env_id0, NULL, obs, reward, done,
                              NULL, n_threads);
}

int orc_sizeof_params(void) { return (int)sizeof(OrcParams); }
int orc_sizeof_env(void) { return (int)sizeof(OrcEnv); }
int orc_sizeof_stepout(void) { return (int)sizeof(OrcStepOut); }
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
