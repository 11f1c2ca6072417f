/*
 * shll_b200.h -- C ABI of libshll_b200.so: the SHLL split-flux Euler time-march on NVIDIA B200.
 *
 * The reference (archembaud/shll-sve-cfd) has no library or plugin interface: every solver is
 * one C file whose main() calls parameter-less void functions over file-scope SoA arrays
 * (base-c/base_shll.c:198-225).  The drop-in boundary is therefore the *program contract*:
 * the host program keeps Allocate_and_Init_Memory / Compute_U_from_P / Save_Results and its
 * float time loop, and the three per-step calls inside `while (time < TOTAL_TIME)`
 *
 *     Compute_F_from_P();   base_shll.c:130-159 . base_shll_2d.c:239-300 . 2nd_order_base_shll.c:461-522
 *     Update_U_from_F();    base_shll.c:87-128  . base_shll_2d.c:139-237 . 2nd_order_base_shll.c:203-459
 *     Compute_P_from_U();   base_shll.c:161-177 . base_shll_2d.c:302-319 . 2nd_order_base_shll.c:524-541
 *
 * collapse into shll_run(ctx, NO_STEPS).  Plain pointers and sizes only; no C++/torch types.
 * All functions return 0 on success, a negative SHLL_E_* code otherwise; the message is
 * available from shll_last_error().  There is no CPU fallback: without a CUDA device
 * shll_create fails with SHLL_E_CUDA.
 */
#ifndef SHLL_B200_H
#define SHLL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHLL_ABI_VERSION 1

#if defined(__GNUC__)
#define SHLL_API __attribute__((visibility("default")))
#else
#define SHLL_API
#endif

/* error codes */
#define SHLL_OK 0
#define SHLL_E_INVAL (-1)    /* bad argument / unsupported combination */
#define SHLL_E_CUDA (-2)     /* CUDA runtime error (message has the cudaError string) */
#define SHLL_E_NOMEM (-3)
#define SHLL_E_STATE (-4)    /* call order (e.g. run before upload, halo not connected) */
#define SHLL_E_TIMEOUT (-5)  /* a peer halo never arrived (multi-GPU) */

/* Boundary rule applied on all physical walls of the domain. */
enum shll_bc {
    SHLL_BC_REFLECT = 0, /* ghost flux = own opposite split flux with sign flips: base_shll.c:93-110, base_shll_2d.c:150-205 */
    SHLL_BC_OUTFLOW = 1  /* ghost flux = own same split flux, zero slopes:       2nd_order_base_shll.c:214-257,280-323 */
};

/* Slope limiter on the split fluxes (order 2 only). */
enum shll_limiter {
    SHLL_LIM_MINMOD = 0, /* 2nd_order_base_shll.c:191-201,268-276 */
    SHLL_LIM_MC = 1      /* base-omp/2nd_order_base_shll.c:317-325 (alpha = 1.25) */
};

/* Which source file's temperature expression a 1D run follows (they round differently). */
enum shll_tform {
    SHLL_TFORM_AUTO = 0, /* 1D order 1 -> 1D form; everything else -> 2D form */
    SHLL_TFORM_1D = 1,   /* base_shll.c:174              T = ((E/rho) - 0.5*u*u)/CV          */
    SHLL_TFORM_2D = 2    /* base_shll_2d.c:316 with v=0  T = ((E/rho) - 0.5*(u*u + v*v))/CV  */
};

/* Arithmetic mode. */
enum shll_mode {
    SHLL_MODE_STRICT = 0, /* bit-exact vs the reference built with gcc -O3 on x86-64 (no FMA, two FP64 islands) */
    SHLL_MODE_FAST = 1    /* pure FP32 + FMA + reciprocal reuse; |x-ref| <= 3e-5 + 3e-5|ref| on the parity configs */
};

typedef struct shll_config {
    uint32_t struct_size;  /* = sizeof(shll_config); lets the ABI grow */
    int32_t dims;          /* 1 or 2 */
    int32_t nx;            /* cells along x owned by THIS context (the whole domain when nranks == 1) */
    int32_t ny;            /* cells along y; 1 for dims == 1.  2D index = i*ny + j, i (x) slow: base_shll_2d.c:157 */
    int32_t order;         /* 1 or 2 */
    int32_t bc;            /* enum shll_bc */
    int32_t limiter;       /* enum shll_limiter */
    int32_t tform;         /* enum shll_tform */
    int32_t mode;          /* enum shll_mode */
    float alpha;           /* MC limiter parameter */
    float dt_on_dx;        /* DT_ON_DX, computed by the HOST exactly as the reference does (base_shll.c:83-84) */
    float dt_on_dy;        /* DT_ON_DY (base_shll_2d.c:136); ignored for dims == 1 */
    int32_t device;        /* CUDA device ordinal */
    int32_t rank;          /* slab index along x, 0 .. nranks-1 */
    int32_t nranks;        /* number of slabs; slab 0 owns the x=0 wall, slab nranks-1 the x=NX-1 wall */
    int32_t variant;       /* kernel variant: 0 = auto; otherwise a tuning id */
    int32_t halo_steps;    /* slabs (nranks > 1): time steps per halo exchange round, K; must be the same on every slab.
                              1D: neighbouring GPUs exchange K*order cells once every K steps instead of `order` cells every
                              step (temporal blocking; same bits); K*order <= 32 and <= the smallest slab.
                              2D: 1 or 2; 2 = the 1st-order FAST kernel advances two steps per launch and exchanges two rows.
                              0 = 1 (exchange every step).  shll_plan_halo_steps() gives the value the front ends use */
    int32_t reserved[6];
} shll_config;

typedef struct shll_ctx shll_ctx;

/* Version of this ABI (SHLL_ABI_VERSION of the built library). */
SHLL_API int shll_abi_version(void);

/* Message of the last error on this context (ctx == NULL: last error of a failed shll_create on this thread). */
SHLL_API const char *shll_last_error(const shll_ctx *ctx);

/* Replays the reference's float clock `while (time < total_time) time += dt` (base_shll.c:200,208,216)
 * and returns the iteration count in *nsteps; SHLL_E_INVAL if the float clock stalls (never terminates). */
SHLL_API int shll_count_steps(float dt, float total_time, long *nsteps);

/* halo_steps for a domain described by `whole` (nx = ALL rows / cells along x) cut into `nslabs` balanced slabs: the value
 * every slab's shll_config.halo_steps should carry (shll_group_create and shll_sve_cfd_b200/slabs.py call it). */
SHLL_API int shll_plan_halo_steps(const shll_config *whole, int nslabs);

/* Replaces Allocate (device side only).  Device buffers: ping-pong U, SoA, FP32. */
SHLL_API int shll_create(shll_ctx **out, const shll_config *cfg);
/* Replaces Free_Memory (device side). */
SHLL_API int shll_destroy(shll_ctx *ctx);

/* Host SoA arrays u[k], k < ncomp (3 for 1D: rho, rho*u, E; 4 for 2D: rho, rho*ux, rho*uy, E), nx*ny floats each:
 * what Compute_U_from_P left in u0..u3 (base_shll.c:74-80). */
SHLL_API int shll_upload_u(shll_ctx *ctx, const float *const u[4]);
SHLL_API int shll_download_u(shll_ctx *ctx, float *const u[4]);
/* Device-side Compute_P_from_U of the current state, downloaded into p0..p3 (and a[] if non-NULL). */
SHLL_API int shll_download_p(shll_ctx *ctx, float *const p[4], float *a);

/* nsteps iterations of {Compute_F_from_P; Update_U_from_F; Compute_P_from_U} fused on the device.
 * Returns after the work is enqueued; shll_sync (or any download) waits for it. */
SHLL_API int shll_run(shll_ctx *ctx, long nsteps);
SHLL_API int shll_sync(shll_ctx *ctx);
/* Same as shll_run + shll_sync, bracketed by CUDA events on the context's stream; *ms = device time. */
SHLL_API int shll_run_timed(shll_ctx *ctx, long nsteps, float *ms);

/* Diagnostic only (the reference has no CFL reduction, base_shll.c:167 is a comment):
 * max over owned cells of (|u| + a) * dt_on_dx (and the y analogue).  Never feeds back into DT. */
SHLL_API int shll_max_cfl(shll_ctx *ctx, float *cfl);

/* Diagnostic only: sums over the owned cells of each conserved component -- mass, momentum
 * (x, and y in 2D), total energy per unit cell volume; sums[k] = 0 for k >= ncomp.  FP64 accumulation in a fixed order
 * (reproducible run to run).  With reflective walls (base_shll.c:93-110) mass and energy are conserved by the scheme up
 * to FP32 rounding of the update; multi-GPU callers add the per-slab sums. */
SHLL_API int shll_conserved_sums(shll_ctx *ctx, double sums[4]);

/* Multi-GPU attribution: time the edge warps of this context spent spinning on a neighbour's halo flag since creation.
 * stats[0] = seconds waiting for the lower neighbour, [1] = for the upper neighbour, [2] = number of waits that had to spin. */
SHLL_API int shll_halo_wait_stats(shll_ctx *ctx, double stats[3]);

/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
SHLL_API long shll_launch_count(const shll_ctx *ctx);
/* Name of the kernel variant the context selected, for logs. */
SHLL_API const char *shll_variant_name(const shll_ctx *ctx);

/* Device self-test of the STRICT-mode exactness shortcuts (csrc/shll_math.cuh: shared-reciprocal float division,
 * division by the constant CV in double): `npairs` operand pairs drawn on the device (uniform bit patterns, physical
 * magnitudes, the guard edges, denormals, signed zeros, infinities, NaNs) are compared bit for bit with the IEEE
 * divisions the reference executes (base_shll.c:173-175, base_shll_2d.c:314-316).  counts[0..3] = mismatches of
 * div_rn_shared / div_rn_spec / div_by_cv / div_by_cv_spec (all must be 0), [4],[5] = draws the one-guard-per-cell forms
 * flagged for IEEE recomputation, [6] = draws that left div_rn_shared's fast path, [7],[8] = float pairs / doubles tested. */
SHLL_API int shll_selftest_exact_division(int device, unsigned long long npairs, unsigned long long seed, unsigned long long counts[9]);

/* ---- multi-GPU slabs (one context per GPU, same or different processes) -------------------------
 * Each context exports a 64-byte CUDA-IPC handle per resource; the host exchanges them (MPI,
 * torch.distributed, a pipe ...) and connects each context to its lower (rank-1) and upper (rank+1)
 * neighbour.  After that shll_run exchanges `order` halo rows per side per step over NVLink by
 * direct peer stores from the edge warps of the step kernel ("halo exchange"). */
#define SHLL_IPC_BYTES 64
typedef struct shll_peer_desc {
    uint8_t state_handle[SHLL_IPC_BYTES]; /* cudaIpcMemHandle_t of the state allocation */
    uint8_t flag_handle[SHLL_IPC_BYTES];  /* cudaIpcMemHandle_t of the halo-arrival flags */
    int64_t pid;                          /* exporting process (same pid => plain peer pointers are used) */
    uint64_t state_ptr;                   /* device address in the exporting process */
    uint64_t flag_ptr;
    int32_t device;
    int32_t nx, ny, dims, order;
    int32_t reserved[3];
} shll_peer_desc;

SHLL_API int shll_peer_export(shll_ctx *ctx, shll_peer_desc *desc);
/* side: -1 = lower neighbour (rank-1), +1 = upper neighbour (rank+1). */
SHLL_API int shll_peer_connect(shll_ctx *ctx, int side, const shll_peer_desc *desc);

/* ---- single-process multi-GPU: a group of slabs behind one handle ------------------------------------
 * What a C program needs to stay ONE process, as the reference programs are (base-c/base_shll_2d.c:343-371), and still
 * use every GPU of the box.  `cfg` describes the WHOLE domain (cfg->nx = all rows / cells along x; rank, nranks and --
 * unless ngpus == 1 -- device are ignored).  Slab r of ngpus owns x indices [r*nx/ngpus, (r+1)*nx/ngpus) on devices[r]
 * (devices == NULL: device r).  The same device may be listed more than once (slabs then share it; used by the
 * single-GPU tests of this path).  The group creates one context per slab, connects neighbours with
 * shll_peer_export / shll_peer_connect (same process => plain peer pointers over NVLink) and feeds every slab from its
 * own short-lived host thread per call.  Host arrays passed to the group calls are the GLOBAL SoA arrays of the
 * reference (u0..u3 / p0..p3 of nx*ny floats): each slab copies its own contiguous row block, nothing is staged.
 * Results are bitwise independent of ngpus in both arithmetic modes (tests/test_group.py, tests/test_multi_gpu.py). */
typedef struct shll_group shll_group;

SHLL_API int shll_group_create(shll_group **out, const shll_config *cfg, int ngpus, const int *devices);
SHLL_API int shll_group_destroy(shll_group *g);
/* Message of the last error on this group (g == NULL: last error of a failed shll_group_create on this thread). */
SHLL_API const char *shll_group_last_error(const shll_group *g);
SHLL_API int shll_group_size(const shll_group *g);
/* The context of one slab (owned by the group), e.g. for shll_variant_name; NULL if out of range. */
SHLL_API shll_ctx *shll_group_ctx(const shll_group *g, int slab);

SHLL_API int shll_group_upload_u(shll_group *g, const float *const u[4]);
SHLL_API int shll_group_download_u(shll_group *g, float *const u[4]);
SHLL_API int shll_group_download_p(shll_group *g, float *const p[4], float *a);
/* nsteps fused time steps on every slab; returns when all slabs have finished them. */
SHLL_API int shll_group_run(shll_group *g, long nsteps);
/* Same, *ms = device time (CUDA events on each slab's stream) of the slowest slab. */
SHLL_API int shll_group_run_timed(shll_group *g, long nsteps, float *ms);
/* Diagnostics over the whole domain: max over slabs / sum over slabs in slab order. */
SHLL_API int shll_group_max_cfl(shll_group *g, float *cfl);
SHLL_API int shll_group_conserved_sums(shll_group *g, double sums[4]);
SHLL_API long shll_group_launch_count(const shll_group *g);

#ifdef __cplusplus
}
#endif
#endif /* SHLL_B200_H */
