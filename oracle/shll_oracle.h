/*
 * shll_oracle.h -- CPU restatement of the SHLL time-march (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for the B200 kernels.  It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product path (libshll_b200.so) never links or calls it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors
 * (SURVEY.md section 4), so the oracle is pinned against the reference itself: the
 * unmodified reference sources are compiled in this container into oracle/_ref/
 * (oracle/build_ref.py) and the restatement must reproduce their raw float32 state
 * bit for bit (tests/test_oracle_vs_ref.py, fixtures under tests/golden/).
 *
 * Arithmetic contract: IEEE-754 binary32 state, C usual arithmetic
 * conversions (unsuffixed literals promote to double), no FMA contraction
 * (build with -ffp-contract=off), denormals preserved.
 */
#ifndef SHLL_ORACLE_H
#define SHLL_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { SHLL_O_BC_REFLECT = 0, SHLL_O_BC_OUTFLOW = 1 };
enum { SHLL_O_LIM_MINMOD = 0, SHLL_O_LIM_MC = 1 };
/* Temperature expression: the 1D file and the 2D files round differently
 * (base-c/base_shll.c:174 vs base-c/base_shll_2d.c:316). */
enum { SHLL_O_TFORM_1D = 1, SHLL_O_TFORM_2D = 2 };
/* Initial conditions, one per reference program. */
enum {
    SHLL_O_IC_SOD_1D = 0,      /* base-c/base_shll.c:54-60 */
    SHLL_O_IC_IMPLOSION = 1,   /* base-c/base_shll_2d.c:93-103 */
    SHLL_O_IC_FOUR_SHOCK = 2,  /* base-c/2nd_order_base_shll.c:133-148 */
    SHLL_O_IC_CONFIG6 = 3,     /* base-omp/2nd_order_base_shll.c:149-167 */
    SHLL_O_IC_SOD_X_2D = 4     /* base-omp/2nd_order_base_shll.c:135-147 (commented IC); y-uniform Sod */
};

typedef struct {
    int dims;        /* 1 or 2 */
    int nx, ny;      /* 2D index = i*ny + j (i = x, slow); 1D: ny = 1 */
    int order;       /* 1 or 2 */
    int bc;          /* SHLL_O_BC_* */
    int limiter;     /* SHLL_O_LIM_* (order 2 only) */
    int tform;       /* SHLL_O_TFORM_* (1D only; 2D always uses the 2D form) */
    float alpha;     /* MC limiter parameter (base-omp: 1.25f) */
    float dt_on_dx;
    float dt_on_dy;
    int nthreads;    /* OpenMP threads for the sweep (results do not depend on it) */
} shll_oracle_cfg;

/* Gas constants exactly as the reference computes them (base_shll.c:17-19). */
float shll_oracle_cv(void);
float shll_oracle_gamma(void);

/* Number of iterations of `while (time < total) time += dt` with a float clock
 * (base_shll.c:200,208,216).  Returns -1 if the float clock stalls. */
long shll_oracle_count_steps(float dt, float total_time);

/* Fill primitive arrays with one of the reference's initial conditions.
 * p[k] has nx*ny floats; 1D uses p[0..2] = rho,u,T; 2D uses p[0..3] = rho,ux,uy,T. */
int shll_oracle_init(const shll_oracle_cfg *cfg, int ic, float *const p[4]);

/* Compute_U_from_P (base_shll.c:74-80, base_shll_2d.c:124-131). */
int shll_oracle_cons_from_prim(const shll_oracle_cfg *cfg, const float *const p[4], float *const u[4]);

/* Compute_P_from_U (base_shll.c:161-177, base_shll_2d.c:302-319). a may be NULL. */
int shll_oracle_prim_from_cons(const shll_oracle_cfg *cfg, const float *const u[4], float *const p[4], float *a);

/* nsteps of Compute_F_from_P -> Update_U_from_F (-> Compute_P_from_U), in place on u. */
int shll_oracle_run(const shll_oracle_cfg *cfg, float *const u[4], long nsteps);

/* Save_Results text format (base_shll.c:180-192, base_shll_2d.c:322-340). */
int shll_oracle_save_results(const shll_oracle_cfg *cfg, const float *const p[4], const char *path);

/* Per-cell pieces, exported for unit tests. */
float shll_oracle_minmod(float left, float right);
float shll_oracle_mc(float fm1, float f0, float fp1, float alpha);
/* One direction's split flux for one 2D cell: un = normal-velocity component index (1=x, 2=y). */
void shll_oracle_split_flux_2d(const float u[4], int dir, float fplus[4], float fminus[4]);
void shll_oracle_split_flux_1d(const float u[3], int tform, float fplus[3], float fminus[3]);

#ifdef __cplusplus
}
#endif
#endif
