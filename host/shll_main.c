/*
 * shll_main.c -- C host program: the reference's program contract on top of libshll_b200.so.
 *
 * One source, five executables (host/Makefile), each a drop-in for one reference program:
 *
 *   -DPROGRAM=1  base_shll               base-c/base_shll.c            1D Sod, 1st order, reflective, N=256, t=0.2
 *   -DPROGRAM=2  base_shll_2d            base-c/base_shll_2d.c         2D implosion, 1st order, reflective, 256^2, t=0.1
 *   -DPROGRAM=3  2nd_order_base_shll     base-c/2nd_order_base_shll.c  2D four-shock, 2nd order (minmod), outflow, 256^2, t=0.8
 *   -DPROGRAM=4  2nd_order_base_shll_1d  derived: the x-sweep of program 3 on a 1D Sod tube, t=0.2
 *   -DPROGRAM=5  base_omp_2nd_order      base-omp/2nd_order_base_shll.c  2D "configuration 6", 2nd order with the MC limiter
 *                                        (alpha = 1.25, :44,317-325), outflow, 1024^2, t=0.3 (README Tables 16/17).  The
 *                                        reference prints `Completed in %d steps` once per OpenMP thread and hard-codes 16
 *                                        threads (:620,652): this program prints the line 16 times too (SHLL_OMP_LINES=n).
 *
 * Same shape as the reference's main() (base_shll.c:198-225): Allocate_and_Init_Memory, Compute_U_from_P, the float
 * clock, `Completed in %d steps`, Save_Results (results.dat), Free_Memory.  The three calls inside the reference's time
 * loop -- Compute_F_from_P, Update_U_from_F, Compute_P_from_U -- run on the GPU: the loop only replays the float clock
 * to count NO_STEPS, then Run_Time_Steps() hands all of them to shll_run().
 *
 * The reference fixes its sizes with `const` globals and is re-edited per run; here they default to the reference's
 * values and can be overridden without recompiling:  ./base_shll [N]   ./base_shll_2d [NX [NY]]   env SHLL_STEPS=<k>
 * (fixed step count, needed for N >= 2^24 where the float clock stalls), SHLL_MODE=fast, SHLL_SAVE=0/1.
 *
 * Multi-GPU, still ONE process like the reference programs: SHLL_NGPUS=<n> cuts the domain into n slabs along x, one per
 * GPU (SHLL_DEVICES=0,1,... picks them; a device may be listed twice), halo rows exchanged by the step kernels over
 * NVLink (shll_group_* in include/shll_b200.h).  stdout and results.dat are byte-identical whatever n is.
 *
 * Additions next to the reference's contract (stdout and results.dat stay byte-identical, everything
 * below goes to stderr or to its own file):
 *   SHLL_SAVE_BIN=1          results.bin: the same primitives as results.dat as raw little-endian float32 planes behind a
 *                            64-byte header -- formatting 268 M lines with fprintf("%e") dwarfs the GPU time at 16384^2
 *   SHLL_SNAPSHOT_EVERY=k    snapshot_<step>.bin in the same format every k steps (device-side Compute_P_from_U)
 *   SHLL_MONITOR=1           max CFL number and the sums of the conserved variables at every snapshot and at the end
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/shll_b200.h"

#ifndef PROGRAM
#define PROGRAM 1
#endif

#if PROGRAM == 1
#define DIMS 1
#define ORDER 1
#define BC SHLL_BC_REFLECT
#define DEFAULT_NX 256
#define DEFAULT_TOTAL_TIME 0.2
#define SAVE_BY_DEFAULT 1 /* base_shll.c:221 calls Save_Results() */
#elif PROGRAM == 2
#define DIMS 2
#define ORDER 1
#define BC SHLL_BC_REFLECT
#define DEFAULT_NX 256
#define DEFAULT_TOTAL_TIME 0.1
#define SAVE_BY_DEFAULT 0 /* commented out in base_shll_2d.c:366 */
#elif PROGRAM == 3
#define DIMS 2
#define ORDER 2
#define BC SHLL_BC_OUTFLOW
#define DEFAULT_NX 256
#define DEFAULT_TOTAL_TIME 0.8
#define SAVE_BY_DEFAULT 0 /* commented out in 2nd_order_base_shll.c:588 */
#elif PROGRAM == 5
#define DIMS 2
#define ORDER 2
#define BC SHLL_BC_OUTFLOW
#define DEFAULT_NX 1024        /* base-omp/2nd_order_base_shll.c:29-31 */
#define DEFAULT_TOTAL_TIME 0.3 /* :43 */
#define SAVE_BY_DEFAULT 0      /* commented out in base-omp/2nd_order_base_shll.c:654 */
#else
#define DIMS 1
#define ORDER 2
#define BC SHLL_BC_OUTFLOW
#define DEFAULT_NX 256
#define DEFAULT_TOTAL_TIME 0.2
#define SAVE_BY_DEFAULT 1
#endif
#if PROGRAM == 5
#define LIMITER SHLL_LIM_MC /* base-omp/2nd_order_base_shll.c:317-325 */
#else
#define LIMITER SHLL_LIM_MINMOD
#endif
#define NCOMP (DIMS == 1 ? 3 : 4)

/* Problem constants, same names and values as the reference (base_shll.c:16-26). */
static int NX = DEFAULT_NX, NY = (DIMS == 2 ? DEFAULT_NX : 1), N;
static const float R = 1.0;
static const float GAMMA = 1.4;
static float CV;
static const float L = 1.0;
static const float H = 1.0;
static float DX, DY;
static const float CFL = 0.25;
static float DT, DT_ON_DX, DT_ON_DY;
static int NO_STEPS = 0;
static float TOTAL_TIME = DEFAULT_TOTAL_TIME;

static float *p[4], *u[4], *a; /* primitives, conserved variables, sound speed: SoA like p0..p3 / u0..u3 */
static shll_group *grp; /* the whole domain: one slab per GPU (a single slab unless SHLL_NGPUS > 1) */

static void die(const char *what, int rc)
{
    const char *msg = grp ? shll_group_last_error(grp) : "";
    if (!msg || !*msg) msg = shll_group_last_error(NULL); /* errors raised without a group (shll_group_create) */
    if (!msg || !*msg) msg = shll_last_error(NULL);       /* ... or without a context (shll_count_steps) */
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, msg);
    exit(1);
}

void Allocate_and_Init_Memory(void)
{
    size_t alignment = 32;
    for (int k = 0; k < NCOMP; k++) {
        if (posix_memalign((void **)&p[k], alignment, (size_t)N * sizeof(float)) ||
            posix_memalign((void **)&u[k], alignment, (size_t)N * sizeof(float))) {
            fprintf(stderr, "out of memory\n");
            exit(1);
        }
    }
    if (posix_memalign((void **)&a, alignment, (size_t)N * sizeof(float))) exit(1);

    long cell = 0;
    for (int i = 0; i < NX; i++) {
        for (int j = 0; j < NY; j++, cell++) {
            float rho, vx = 0.0, vy = 0.0, T = 1.0;
#if PROGRAM == 1 || PROGRAM == 4
            rho = (i < 0.5 * NX) ? 10.0 : 1.0; /* Sod tube, base_shll.c:55-59 */
#elif PROGRAM == 2
            int inside = (i > 0.2 * NX) && (i < 0.8 * NX) && (j > 0.2 * NY) && (j < 0.8 * NY);
            rho = inside ? 1.0 : 10.0; /* implosion, base_shll_2d.c:96-100 */
#elif PROGRAM == 5
            /* "Configuration 6", base-omp/2nd_order_base_shll.c:149-167 (cells exactly on a 1/2 line fall to the last state) */
            int lo_i = i < 0.5 * NX, hi_i = i > 0.5 * NX, lo_j = j < 0.5 * NY, hi_j = j > 0.5 * NY;
            if (lo_i && lo_j)      { rho = 1.0; vx = -0.75; vy = 0.5;  T = (1.0 / (rho * R)); }
            else if (hi_i && lo_j) { rho = 3.0; vx = -0.75; vy = -0.5; T = (1.0 / (rho * R)); }
            else if (lo_i && hi_j) { rho = 2.0; vx = 0.75;  vy = 0.5;  T = (1.0 / (rho * R)); }
            else                   { rho = 1.0; vx = 0.75;  vy = -0.5; T = (1.0 / (rho * R)); }
#else
            /* Euler four-shock problem, 2nd_order_base_shll.c:137-145 (cells exactly on a 3/4 line fall to the last state) */
            int lo_i = i < 0.75 * NX, hi_i = i > 0.75 * NX, lo_j = j < 0.75 * NY, hi_j = j > 0.75 * NY;
            if (lo_i && lo_j)      { rho = 0.138;  vx = 1.206; vy = 1.206; T = (0.029 / (rho * R)); }
            else if (hi_i && lo_j) { rho = 0.5323; vx = 0.0;   vy = 1.206; T = (0.3 / (rho * R)); }
            else if (lo_i && hi_j) { rho = 0.5323; vx = 1.206; vy = 0.0;   T = (0.3 / (rho * R)); }
            else                   { rho = 1.5;    vx = 0.0;   vy = 0.0;   T = (1.5 / (rho * R)); }
#endif
            p[0][cell] = rho;
            p[1][cell] = vx;
            if (DIMS == 2) { p[2][cell] = vy; p[3][cell] = T; } else { p[2][cell] = T; }
        }
    }
}

void Free_Memory(void)
{
    for (int k = 0; k < NCOMP; k++) { free(p[k]); free(u[k]); }
    free(a);
}

void Compute_U_from_P(void)
{
    for (long c = 0; c < N; c++) {
        u[0][c] = p[0][c];
        u[1][c] = p[0][c] * p[1][c];
        if (DIMS == 1) {
            u[2][c] = p[0][c] * (p[2][c] * CV + 0.5 * p[1][c] * p[1][c]); /* base_shll.c:79 */
        } else {
            u[2][c] = p[0][c] * p[2][c];
            u[3][c] = p[0][c] * (p[3][c] * CV + 0.5 * (p[1][c] * p[1][c] + p[2][c] * p[2][c])); /* base_shll_2d.c:130 */
        }
    }
    /* Estimated CFL = ((R + 1)*DT)/DX  (base_shll.c:82-84; base_shll_2d.c:134-136) */
    DT = (CFL / (R + 1)) * DX;
    DT_ON_DX = (CFL / (R + 1));
    DT_ON_DY = DT / DY;
}

/* results.bin / snapshot_<step>.bin: 64-byte header {"SHLLBIN1", dims, nx, ny, ncomp, steps, 9 x int32 zero}, then ncomp
 * planes of nx*ny little-endian float32 primitives in the reference's index order (rho, u, [v,] T). */
static void Save_Binary(const char *path, int steps)
{
    FILE *f = fopen(path, "wb");
    if (!f) { perror(path); exit(1); }
    int32_t hdr[16];
    memset(hdr, 0, sizeof(hdr));
    memcpy(hdr, "SHLLBIN1", 8);
    hdr[2] = DIMS; hdr[3] = NX; hdr[4] = NY; hdr[5] = NCOMP; hdr[6] = steps;
    int ok = fwrite(hdr, sizeof(hdr), 1, f) == 1;
    for (int k = 0; k < NCOMP && ok; k++) ok = fwrite(p[k], sizeof(float), (size_t)N, f) == (size_t)N;
    if (fclose(f) || !ok) { fprintf(stderr, "short write on %s\n", path); exit(1); }
}

/* Diagnostics only (the reference has neither): they never feed back into DT. */
static void Monitor(int steps)
{
    float cfl = 0.0f;
    double sums[4];
    int rc;
    if ((rc = shll_group_max_cfl(grp, &cfl))) die("shll_group_max_cfl", rc);
    if ((rc = shll_group_conserved_sums(grp, sums))) die("shll_group_conserved_sums", rc);
    const double vol = (double)DX * (DIMS == 2 ? (double)DY : 1.0);
    fprintf(stderr, "monitor step %d: max CFL %.6f  mass %.12e  momentum %.12e", steps, cfl, sums[0] * vol, sums[1] * vol);
    if (DIMS == 2) fprintf(stderr, " %.12e", sums[2] * vol);
    fprintf(stderr, "  energy %.12e\n", sums[NCOMP - 1] * vol);
}

/* Replaces the body of the reference's time loop: NO_STEPS x {Compute_F_from_P; Update_U_from_F; Compute_P_from_U}. */
void Run_Time_Steps(void)
{
    int rc;
    const int every = getenv("SHLL_SNAPSHOT_EVERY") ? atoi(getenv("SHLL_SNAPSHOT_EVERY")) : 0;
    const int monitor = getenv("SHLL_MONITOR") ? atoi(getenv("SHLL_MONITOR")) : 0;
    if ((rc = shll_group_upload_u(grp, (const float *const *)u))) die("shll_group_upload_u", rc);
    if (monitor) Monitor(0);
    int done = 0;
    while (every > 0 && done + every < NO_STEPS) {
        if ((rc = shll_group_run(grp, every))) die("shll_group_run", rc);
        done += every;
        if ((rc = shll_group_download_p(grp, p, a))) die("shll_group_download_p", rc);
        char name[64];
        snprintf(name, sizeof(name), "snapshot_%08d.bin", done);
        Save_Binary(name, done);
        if (monitor) Monitor(done);
    }
    if ((rc = shll_group_run(grp, NO_STEPS - done))) die("shll_group_run", rc);
    if ((rc = shll_group_download_u(grp, u))) die("shll_group_download_u", rc);
    if ((rc = shll_group_download_p(grp, p, a))) die("shll_group_download_p", rc); /* the last Compute_P_from_U */
    if (monitor) Monitor(NO_STEPS);
}

void Save_Results(void)
{
    FILE *fptr = fopen("results.dat", "w");
    if (!fptr) { perror("results.dat"); exit(1); }
    if (DIMS == 2) printf("Saving to file\n");
    long index = 0;
    for (int i = 0; i < NX; i++) {
        for (int j = 0; j < NY; j++, index++) {
            float cx = (i + 0.5) * DX;
            if (DIMS == 1) {
                fprintf(fptr, "%e\t%e\t%e\t%e\n", cx, p[0][index], p[1][index], p[2][index]);
            } else {
                float cy = (j + 0.5) * DY;
                fprintf(fptr, "%e\t%e\t%e\t%e\t%e\t%e\n", cx, cy, p[0][index], p[1][index], p[2][index], p[3][index]);
            }
        }
    }
    fclose(fptr);
    if (DIMS == 2) printf("Completed saving data\n");
}

int main(int argc, char **argv)
{
    if (argc > 1) NX = atoi(argv[1]);
    if (DIMS == 2) NY = (argc > 2) ? atoi(argv[2]) : NX;
    if (getenv("SHLL_TOTAL_TIME")) TOTAL_TIME = (float)atof(getenv("SHLL_TOTAL_TIME"));
    N = NX * NY;
    CV = R / (GAMMA - 1.0);
    DX = L / NX;
    DY = (DIMS == 2) ? H / NY : 1.0f;

    float time = 0.0;
    Allocate_and_Init_Memory();
    Compute_U_from_P();

    shll_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.struct_size = sizeof(cfg);
    cfg.dims = DIMS; cfg.nx = NX; cfg.ny = NY; cfg.order = ORDER; cfg.bc = BC;
    cfg.limiter = LIMITER; cfg.alpha = 1.25f; /* alpha: base-omp/2nd_order_base_shll.c:44 */
    cfg.tform = (PROGRAM == 4) ? SHLL_TFORM_2D : SHLL_TFORM_AUTO;
    cfg.mode = (getenv("SHLL_MODE") && !strcmp(getenv("SHLL_MODE"), "fast")) ? SHLL_MODE_FAST : SHLL_MODE_STRICT;
    cfg.dt_on_dx = DT_ON_DX; cfg.dt_on_dy = DT_ON_DY;
    cfg.device = getenv("SHLL_DEVICE") ? atoi(getenv("SHLL_DEVICE")) : 0;
    cfg.rank = 0; cfg.nranks = 1;
    int ngpus = getenv("SHLL_NGPUS") ? atoi(getenv("SHLL_NGPUS")) : 1;
    int devices[64], ndevices = 0;
    if (getenv("SHLL_DEVICES")) { /* e.g. 0,1,2,3 -- one entry per slab */
        char list[256];
        snprintf(list, sizeof(list), "%s", getenv("SHLL_DEVICES"));
        for (char *tok = strtok(list, ","); tok && ndevices < 64; tok = strtok(NULL, ",")) devices[ndevices++] = atoi(tok);
        if (!getenv("SHLL_NGPUS")) ngpus = ndevices;
        if (ndevices != ngpus) { fprintf(stderr, "SHLL_DEVICES lists %d devices, SHLL_NGPUS is %d\n", ndevices, ngpus); return 1; }
    }
    int rc = shll_group_create(&grp, &cfg, ngpus, ndevices ? devices : NULL);
    if (rc) die("shll_group_create", rc);

    /* Take some timesteps: the float clock decides how many (base_shll.c:208-218) */
    if (getenv("SHLL_STEPS")) {
        NO_STEPS = atoi(getenv("SHLL_STEPS"));
    } else {
        long n = 0;
        if ((rc = shll_count_steps(DT, TOTAL_TIME, &n))) die("shll_count_steps", rc);
        while (time < TOTAL_TIME) { time += DT; NO_STEPS += 1; }
        if (n != NO_STEPS) { fprintf(stderr, "step count mismatch %ld vs %d\n", n, NO_STEPS); return 1; }
    }
    Run_Time_Steps();

    int lines = 1;
#if PROGRAM == 5
    lines = getenv("SHLL_OMP_LINES") ? atoi(getenv("SHLL_OMP_LINES")) : 16; /* one line per OpenMP thread, base-omp/...:620,652 */
#endif
    for (int l = 0; l < lines; l++) printf("Completed in %d steps\n", NO_STEPS);
    int save = getenv("SHLL_SAVE") ? atoi(getenv("SHLL_SAVE")) : SAVE_BY_DEFAULT;
    if (save) Save_Results();
    if (getenv("SHLL_SAVE_BIN") && atoi(getenv("SHLL_SAVE_BIN"))) Save_Binary("results.bin", NO_STEPS);

    shll_group_destroy(grp);
    Free_Memory();
    return 0;
}
