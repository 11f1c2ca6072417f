/*
 * shll_oracle.c -- CPU restatement of the SHLL split-flux Euler time-march.
 *
 * TEST INFRASTRUCTURE ONLY (see shll_oracle.h).  Parity status: PINNED against the
 * compiled reference (oracle/_ref, tests/golden).
 *
 * Structure differs from the reference on purpose: the reference materialises 22/49/81
 * global arrays and sweeps them in five functions; here one step is
 *   (1) per-cell primitive recompute + split fluxes      -> flux scratch
 *   (2) per-cell limited slopes of the split fluxes       -> slope scratch (order 2)
 *   (3) per-cell conservative update, neighbours and boundary rules evaluated inline.
 * The floating-point expression of every value is the reference's (cited per function);
 * that is what makes the result bit-identical.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math [-fopenmp] -shared -fPIC
 */
#include "shll_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* base_shll.c:17-19 / base_shll_2d.c:24-26: float constants, CV evaluated in double. */
static const float kR = 1.0;
static const float kGamma = 1.4;
#define K_CV ((float)(kR / (kGamma - 1.0)))

float shll_oracle_cv(void) { return K_CV; }
float shll_oracle_gamma(void) { return kGamma; }

/* base_shll.c:200,208,216-217 -- float clock. */
long shll_oracle_count_steps(float dt, float total_time)
{
    float t = 0.0;
    long n = 0;
    while (t < total_time) {
        float t_next = t + dt;
        if (t_next == t) return -1; /* clock stalled: the reference would never terminate */
        t = t_next;
        n++;
    }
    return n;
}

/* 2nd_order_base_shll.c:191-201.  The sign test is on the float product. */
float shll_oracle_minmod(float left, float right)
{
    if (left * right < 0.0) return 0.0;
    return (fabs(left) < fabs(right)) ? left : right;
}

/* base-omp/2nd_order_base_shll.c:317 -- MC limiter built from two minmods. */
float shll_oracle_mc(float fm1, float f0, float fp1, float alpha)
{
    float central = 0.5 * (fp1 - fm1);
    float inner = shll_oracle_minmod(f0 - fm1, fp1 - f0);
    return shll_oracle_minmod(central, alpha * inner);
}

static inline float limited_slope(int limiter, float alpha, float fm1, float f0, float fp1)
{
    if (limiter == SHLL_O_LIM_MC) return shll_oracle_mc(fm1, f0, fp1, alpha);
    return shll_oracle_minmod(f0 - fm1, fp1 - f0);
}

/* ---------------------------------------------------------------- primitives */

typedef struct { float rho, ux, uy, T, a; } prim_t;

/* base_shll_2d.c:313-317 (identical in 2nd_order_base_shll.c:535-539). */
static inline prim_t prim_2d(float u0, float u1, float u2, float u3)
{
    const float CV = K_CV;
    prim_t q;
    q.rho = u0;
    q.ux = u1 / u0;
    q.uy = u2 / u0;
    q.T = ((u3 / u0) - 0.5 * (q.ux * q.ux + q.uy * q.uy)) / CV;
    q.a = sqrt(kGamma * kR * q.T);
    return q;
}

/* base_shll.c:172-175 (tform 1D) or the 2D expression with uy = 0 (tform 2D,
 * the derived 1D 2nd-order program). */
static inline prim_t prim_1d(float u0, float u1, float u2, int tform)
{
    const float CV = K_CV;
    prim_t q;
    q.rho = u0;
    q.ux = u1 / u0;
    q.uy = 0.0f;
    if (tform == SHLL_O_TFORM_1D)
        q.T = ((u2 / u0) - 0.5 * q.ux * q.ux) / CV;
    else
        q.T = ((u2 / u0) - 0.5 * (q.ux * q.ux + q.uy * q.uy)) / CV;
    q.a = sqrt(kGamma * kR * q.T);
    return q;
}

/* ---------------------------------------------------------------- split fluxes */

/* base_shll_2d.c:246-298.  dir 1 = x (F), dir 2 = y (H). */
static inline void split_2d(const prim_t *q, float u0, float u1, float u2, float u3, int dir,
                            float fp[4], float fm[4])
{
    float P = q->rho * kR * q->T;
    float un = (dir == 1) ? q->ux : q->uy;
    float M = un / q->a;
    float Z1 = 0.5 * (M + 1.0);
    float Z2 = 0.5 * q->a * (1.0 - M * M);
    float Z3 = 0.5 * (M - 1.0);
    float f[4];
    if (dir == 1) {
        f[0] = u1;
        f[1] = u1 * q->ux + P;
        f[2] = u1 * q->uy;
        f[3] = q->ux * (u3 + P);
    } else {
        f[0] = u2;
        f[1] = u2 * q->ux;
        f[2] = u2 * q->uy + P;
        f[3] = q->uy * (u3 + P);
    }
    const float u[4] = {u0, u1, u2, u3};
    for (int k = 0; k < 4; k++) {
        fp[k] = f[k] * Z1 + u[k] * Z2;
        fm[k] = -f[k] * Z3 - u[k] * Z2;
    }
}

/* base_shll.c:135-157. */
static inline void split_1d(const prim_t *q, float u0, float u1, float u2, float fp[3], float fm[3])
{
    float M = q->ux / q->a;
    float Z1 = 0.5 * (M + 1.0);
    float Z2 = 0.5 * q->a * (1.0 - M * M);
    float Z3 = 0.5 * (M - 1.0);
    float P = q->rho * kR * q->T;
    float f[3];
    f[0] = u1;
    f[1] = u1 * q->ux + P;
    f[2] = q->ux * (u2 + P);
    const float u[3] = {u0, u1, u2};
    for (int k = 0; k < 3; k++) {
        fp[k] = f[k] * Z1 + u[k] * Z2;
        fm[k] = -f[k] * Z3 - u[k] * Z2;
    }
}

void shll_oracle_split_flux_2d(const float u[4], int dir, float fplus[4], float fminus[4])
{
    prim_t q = prim_2d(u[0], u[1], u[2], u[3]);
    split_2d(&q, u[0], u[1], u[2], u[3], dir, fplus, fminus);
}

void shll_oracle_split_flux_1d(const float u[3], int tform, float fplus[3], float fminus[3])
{
    prim_t q = prim_1d(u[0], u[1], u[2], tform);
    split_1d(&q, u[0], u[1], u[2], fplus, fminus);
}

/* ---------------------------------------------------------------- IC / P<->U */

static int check_cfg(const shll_oracle_cfg *c)
{
    if (!c) return -1;
    if (c->dims == 1) {
        if (c->nx < 2 || c->ny != 1) return -1;
        if (c->tform != SHLL_O_TFORM_1D && c->tform != SHLL_O_TFORM_2D) return -1;
    } else if (c->dims == 2) {
        if (c->nx < 2 || c->ny < 2) return -1;
    } else {
        return -1;
    }
    if (c->order != 1 && c->order != 2) return -1;
    if (c->bc != SHLL_O_BC_REFLECT && c->bc != SHLL_O_BC_OUTFLOW) return -1;
    if (c->limiter != SHLL_O_LIM_MINMOD && c->limiter != SHLL_O_LIM_MC) return -1;
    return 0;
}

int shll_oracle_init(const shll_oracle_cfg *cfg, int ic, float *const p[4])
{
    if (check_cfg(cfg)) return -1;
    const int NX = cfg->nx, NY = cfg->ny;
    if (cfg->dims == 1) {
        if (ic != SHLL_O_IC_SOD_1D) return -2;
        for (int i = 0; i < NX; i++) { /* base_shll.c:54-60 */
            p[0][i] = (i < 0.5 * NX) ? 10.0 : 1.0;
            p[1][i] = 0.0;
            p[2][i] = 1.0;
        }
        return 0;
    }
    long c = 0;
    for (int i = 0; i < NX; i++) {
        for (int j = 0; j < NY; j++, c++) {
            float r, vx, vy, T;
            switch (ic) {
            case SHLL_O_IC_IMPLOSION: /* base_shll_2d.c:96-100 */
                if ((i > 0.2 * NX) && (i < 0.8 * NX) && (j > 0.2 * NY) && (j < 0.8 * NY)) r = 1.0;
                else r = 10.0;
                vx = 0.0; vy = 0.0; T = 1.0;
                break;
            case SHLL_O_IC_FOUR_SHOCK: /* 2nd_order_base_shll.c:137-145 */
                if ((i < 0.75 * NX) && (j < 0.75 * NY)) { r = 0.138; vx = 1.206; vy = 1.206; T = (0.029 / (r * kR)); }
                else if ((i > 0.75 * NX) && (j < 0.75 * NY)) { r = 0.5323; vx = 0.0; vy = 1.206; T = (0.3 / (r * kR)); }
                else if ((i < 0.75 * NX) && (j > 0.75 * NY)) { r = 0.5323; vx = 1.206; vy = 0.0; T = (0.3 / (r * kR)); }
                else { r = 1.5; vx = 0.0; vy = 0.0; T = (1.5 / (r * kR)); }
                break;
            case SHLL_O_IC_CONFIG6: /* base-omp/2nd_order_base_shll.c:149-167 */
                if ((i < 0.5 * NX) && (j < 0.5 * NY)) { r = 1.0; vx = -0.75; vy = 0.5; T = (1.0 / (r * kR)); }
                else if ((i > 0.5 * NX) && (j < 0.5 * NY)) { r = 3.0; vx = -0.75; vy = -0.5; T = (1.0 / (r * kR)); }
                else if ((i < 0.5 * NX) && (j > 0.5 * NY)) { r = 2.0; vx = 0.75; vy = 0.5; T = (1.0 / (r * kR)); }
                else { r = 1.0; vx = 0.75; vy = -0.5; T = (1.0 / (r * kR)); }
                break;
            case SHLL_O_IC_SOD_X_2D: /* base-omp/2nd_order_base_shll.c:138-142 (commented-out IC) */
                r = (i < 0.5 * NX) ? 10.0 : 1.0;
                vx = 0.0; vy = 0.0; T = 1.0;
                break;
            default:
                return -2;
            }
            p[0][c] = r; p[1][c] = vx; p[2][c] = vy; p[3][c] = T;
        }
    }
    return 0;
}

int shll_oracle_cons_from_prim(const shll_oracle_cfg *cfg, const float *const p[4], float *const u[4])
{
    if (check_cfg(cfg)) return -1;
    const float CV = K_CV;
    const long n = (long)cfg->nx * cfg->ny;
    if (cfg->dims == 1) {
        for (long c = 0; c < n; c++) { /* base_shll.c:77-79 */
            u[0][c] = p[0][c];
            u[1][c] = p[0][c] * p[1][c];
            u[2][c] = p[0][c] * (p[2][c] * CV + 0.5 * p[1][c] * p[1][c]);
        }
    } else {
        for (long c = 0; c < n; c++) { /* base_shll_2d.c:127-130 */
            u[0][c] = p[0][c];
            u[1][c] = p[0][c] * p[1][c];
            u[2][c] = p[0][c] * p[2][c];
            u[3][c] = p[0][c] * (p[3][c] * CV + 0.5 * (p[1][c] * p[1][c] + p[2][c] * p[2][c]));
        }
    }
    return 0;
}

int shll_oracle_prim_from_cons(const shll_oracle_cfg *cfg, const float *const u[4], float *const p[4], float *a)
{
    if (check_cfg(cfg)) return -1;
    const long n = (long)cfg->nx * cfg->ny;
    if (cfg->dims == 1) {
        for (long c = 0; c < n; c++) {
            prim_t q = prim_1d(u[0][c], u[1][c], u[2][c], cfg->tform);
            p[0][c] = q.rho; p[1][c] = q.ux; p[2][c] = q.T;
            if (a) a[c] = q.a;
        }
    } else {
        for (long c = 0; c < n; c++) {
            prim_t q = prim_2d(u[0][c], u[1][c], u[2][c], u[3][c]);
            p[0][c] = q.rho; p[1][c] = q.ux; p[2][c] = q.uy; p[3][c] = q.T;
            if (a) a[c] = q.a;
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- time march */

/* Reflective-wall sign pattern of the ghost flux.
 * 1D (rho, rho*u, E): (-,+,-)        base_shll.c:95-97,108-110
 * 2D x walls: (-,+,-,-)              base_shll_2d.c:152-155,168-171
 * 2D y walls: (-,-,+,-)              base_shll_2d.c:186-189,202-205 */
static inline float wall_sign(int ncomp, int dir, int k)
{
    if (ncomp == 3) return (k == 1) ? 1.0f : -1.0f;
    return (k == dir) ? 1.0f : -1.0f;
}

static int run_1d(const shll_oracle_cfg *cfg, float *const u[4], long nsteps)
{
    const int N = cfg->nx;
    const int o2 = (cfg->order == 2);
    float *buf = (float *)malloc(sizeof(float) * (size_t)N * 12);
    if (!buf) return -3;
    float *fp[3], *fm[3], *dfp[3], *dfm[3];
    for (int k = 0; k < 3; k++) {
        fp[k] = buf + (size_t)N * k;
        fm[k] = buf + (size_t)N * (3 + k);
        dfp[k] = buf + (size_t)N * (6 + k);
        dfm[k] = buf + (size_t)N * (9 + k);
    }
    const float dtdx = cfg->dt_on_dx;
    const int nth = cfg->nthreads > 0 ? cfg->nthreads : 1;
    (void)nth;
    for (long s = 0; s < nsteps; s++) {
#pragma omp parallel num_threads(nth)
        {
#pragma omp for
            for (int i = 0; i < N; i++) {
                prim_t q = prim_1d(u[0][i], u[1][i], u[2][i], cfg->tform);
                float a[3], b[3];
                split_1d(&q, u[0][i], u[1][i], u[2][i], a, b);
                for (int k = 0; k < 3; k++) { fp[k][i] = a[k]; fm[k][i] = b[k]; }
            }
            if (o2) {
#pragma omp for
                for (int i = 0; i < N; i++) {
                    for (int k = 0; k < 3; k++) {
                        if (i == 0 || i == N - 1) { /* 2nd_order_base_shll.c:226-234,248-256 */
                            dfp[k][i] = 0.0; dfm[k][i] = 0.0;
                        } else {                    /* :268-276 */
                            dfp[k][i] = limited_slope(cfg->limiter, cfg->alpha, fp[k][i - 1], fp[k][i], fp[k][i + 1]);
                            dfm[k][i] = limited_slope(cfg->limiter, cfg->alpha, fm[k][i - 1], fm[k][i], fm[k][i + 1]);
                        }
                    }
                }
            }
#pragma omp for
            for (int i = 0; i < N; i++) {
                for (int k = 0; k < 3; k++) {
                    float left, right;
                    if (i == 0) left = (cfg->bc == SHLL_O_BC_REFLECT) ? wall_sign(3, 1, k) * fm[k][i] : fp[k][i];
                    else left = fp[k][i - 1];
                    if (i == N - 1) right = (cfg->bc == SHLL_O_BC_REFLECT) ? wall_sign(3, 1, k) * fp[k][i] : fm[k][i];
                    else right = fm[k][i + 1];
                    float v = u[k][i];
                    v = v - dtdx * (fp[k][i] - fm[k][i] + right - left); /* base_shll.c:124 */
                    if (o2) {
                        float ldf = (i == 0) ? 0.0f : dfp[k][i - 1];     /* 2nd_order_base_shll.c:360-391 */
                        float rdf = (i == N - 1) ? 0.0f : dfm[k][i + 1];
                        v = v - 0.5 * dtdx * (dfp[k][i] + dfm[k][i] - rdf - ldf); /* :443 */
                    }
                    u[k][i] = v;
                }
            }
        }
    }
    free(buf);
    return 0;
}

static int run_2d(const shll_oracle_cfg *cfg, float *const u[4], long nsteps)
{
    const int NX = cfg->nx, NY = cfg->ny;
    const size_t n = (size_t)NX * NY;
    const int o2 = (cfg->order == 2);
    const int reflect = (cfg->bc == SHLL_O_BC_REFLECT);
    float *buf = (float *)malloc(sizeof(float) * n * (o2 ? 32 : 16));
    if (!buf) return -3;
    float *fp[4], *fm[4], *hp[4], *hm[4], *dfp[4], *dfm[4], *dhp[4], *dhm[4];
    for (int k = 0; k < 4; k++) {
        fp[k] = buf + n * k;
        fm[k] = buf + n * (4 + k);
        hp[k] = buf + n * (8 + k);
        hm[k] = buf + n * (12 + k);
        dfp[k] = o2 ? buf + n * (16 + k) : NULL;
        dfm[k] = o2 ? buf + n * (20 + k) : NULL;
        dhp[k] = o2 ? buf + n * (24 + k) : NULL;
        dhm[k] = o2 ? buf + n * (28 + k) : NULL;
    }
    const float dtdx = cfg->dt_on_dx, dtdy = cfg->dt_on_dy;
    const int nth = cfg->nthreads > 0 ? cfg->nthreads : 1;
    (void)nth;
    for (long s = 0; s < nsteps; s++) {
#pragma omp parallel num_threads(nth)
        {
#pragma omp for
            for (int i = 0; i < NX; i++) {
                for (int j = 0; j < NY; j++) {
                    size_t c = (size_t)i * NY + j;
                    prim_t q = prim_2d(u[0][c], u[1][c], u[2][c], u[3][c]);
                    float a[4], b[4];
                    split_2d(&q, u[0][c], u[1][c], u[2][c], u[3][c], 1, a, b);
                    for (int k = 0; k < 4; k++) { fp[k][c] = a[k]; fm[k][c] = b[k]; }
                    split_2d(&q, u[0][c], u[1][c], u[2][c], u[3][c], 2, a, b);
                    for (int k = 0; k < 4; k++) { hp[k][c] = a[k]; hm[k][c] = b[k]; }
                }
            }
            if (o2) {
#pragma omp for
                for (int i = 0; i < NX; i++) {
                    for (int j = 0; j < NY; j++) {
                        size_t c = (size_t)i * NY + j;
                        for (int k = 0; k < 4; k++) {
                            if (i == 0 || i == NX - 1) { dfp[k][c] = 0.0; dfm[k][c] = 0.0; }
                            else {
                                dfp[k][c] = limited_slope(cfg->limiter, cfg->alpha, fp[k][c - NY], fp[k][c], fp[k][c + NY]);
                                dfm[k][c] = limited_slope(cfg->limiter, cfg->alpha, fm[k][c - NY], fm[k][c], fm[k][c + NY]);
                            }
                            if (j == 0 || j == NY - 1) { dhp[k][c] = 0.0; dhm[k][c] = 0.0; }
                            else {
                                dhp[k][c] = limited_slope(cfg->limiter, cfg->alpha, hp[k][c - 1], hp[k][c], hp[k][c + 1]);
                                dhm[k][c] = limited_slope(cfg->limiter, cfg->alpha, hm[k][c - 1], hm[k][c], hm[k][c + 1]);
                            }
                        }
                    }
                }
            }
#pragma omp for
            for (int i = 0; i < NX; i++) {
                for (int j = 0; j < NY; j++) {
                    size_t c = (size_t)i * NY + j;
                    for (int k = 0; k < 4; k++) {
                        float left, right, bottom, top;
                        if (i == 0) left = reflect ? wall_sign(4, 1, k) * fm[k][c] : fp[k][c];
                        else left = fp[k][c - NY];
                        if (i == NX - 1) right = reflect ? wall_sign(4, 1, k) * fp[k][c] : fm[k][c];
                        else right = fm[k][c + NY];
                        if (j == 0) bottom = reflect ? wall_sign(4, 2, k) * hm[k][c] : hp[k][c];
                        else bottom = hp[k][c - 1];
                        if (j == NY - 1) top = reflect ? wall_sign(4, 2, k) * hp[k][c] : hm[k][c];
                        else top = hm[k][c + 1];
                        float v = u[k][c];
                        v = v - dtdx * (fp[k][c] - fm[k][c] + right - left);           /* base_shll_2d.c:227 */
                        if (o2) {
                            float ldf = (i == 0) ? 0.0f : dfp[k][c - NY];
                            float rdf = (i == NX - 1) ? 0.0f : dfm[k][c + NY];
                            v = v - 0.5 * dtdx * (dfp[k][c] + dfm[k][c] - rdf - ldf);  /* 2nd_order_base_shll.c:443 */
                        }
                        v = v - dtdy * (hp[k][c] - hm[k][c] + top - bottom);           /* base_shll_2d.c:232 */
                        if (o2) {
                            float bdf = (j == 0) ? 0.0f : dhp[k][c - 1];
                            float tdf = (j == NY - 1) ? 0.0f : dhm[k][c + 1];
                            v = v - 0.5 * dtdy * (dhp[k][c] + dhm[k][c] - tdf - bdf);  /* 2nd_order_base_shll.c:454 */
                        }
                        u[k][c] = v;
                    }
                }
            }
        }
    }
    free(buf);
    return 0;
}

int shll_oracle_run(const shll_oracle_cfg *cfg, float *const u[4], long nsteps)
{
    if (check_cfg(cfg)) return -1;
    if (nsteps < 0) return -1;
    return (cfg->dims == 1) ? run_1d(cfg, u, nsteps) : run_2d(cfg, u, nsteps);
}

/* ---------------------------------------------------------------- results.dat */

int shll_oracle_save_results(const shll_oracle_cfg *cfg, const float *const p[4], const char *path)
{
    if (check_cfg(cfg)) return -1;
    FILE *f = fopen(path, "w");
    if (!f) return -4;
    const float L = 1.0, H = 1.0;
    if (cfg->dims == 1) { /* base_shll.c:186-189 */
        const float DX = L / cfg->nx;
        for (int i = 0; i < cfg->nx; i++) {
            float cx = (i + 0.5) * DX;
            fprintf(f, "%e\t%e\t%e\t%e\n", cx, p[0][i], p[1][i], p[2][i]);
        }
    } else { /* base_shll_2d.c:329-336 */
        const float DX = L / cfg->nx, DY = H / cfg->ny;
        size_t c = 0;
        for (int i = 0; i < cfg->nx; i++) {
            for (int j = 0; j < cfg->ny; j++, c++) {
                float cx = (i + 0.5) * DX;
                float cy = (j + 0.5) * DY;
                fprintf(f, "%e\t%e\t%e\t%e\t%e\t%e\n", cx, cy, p[0][c], p[1][c], p[2][c], p[3][c]);
            }
        }
    }
    fclose(f);
    return 0;
}
