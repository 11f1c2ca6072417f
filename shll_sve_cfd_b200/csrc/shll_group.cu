// shll_group.cu -- single-process multi-GPU front end of the C ABI (include/shll_b200.h, "shll_group_*").
//
// The reference programs are single-process C programs (base-c/base_shll_2d.c:343-371); a drop-in for them that uses
// all GPUs of a box must not need a process launcher.  A group owns one shll_ctx per slab -- contiguous row blocks of
// the x axis, slab r on devices[r] -- wires neighbouring contexts together with the same peer
// descriptors the multi-process path exchanges (same pid => plain peer pointers, shll_peer_connect) and drives them
// with one short-lived host thread per slab per call: shll_run enqueues all steps of a slab on that slab's stream, and
// a step kernel of slab r spins on the halo flag written by slab r+-1's previous step, so every slab's launches have
// to be fed concurrently (one host thread feeding the slabs in turn would dead-lock as soon as a launch queue fills).
// The data path is untouched: halo rows travel by peer stores from the edge warps of the step kernels (halo_sync.cuh).
//
// Built on the public per-context entry points only.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/shll_b200.h"

struct shll_group {
    shll_config cfg;  // the WHOLE domain
    int n;
    long row;         // floats per x index: ny (2D) or 1 (1D)
    int ncomp;
    std::vector<shll_ctx *> ctx;
    std::vector<int> i0, nx;  // first global row and row count of each slab
    char err[600];
};

namespace {

thread_local char g_group_error[600] = "";

int gfail(shll_group *g, int code, const char *fmt, ...)
{
    char *dst = g ? g->err : g_group_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 600, fmt, ap);
    va_end(ap);
    return code;
}

// f(r) on one host thread per slab; first failing slab's code and message win.
template <class F>
int for_all_slabs(shll_group *g, const char *what, F &&f)
{
    std::vector<int> rc(g->n, SHLL_OK);
    if (g->n == 1) {
        rc[0] = f(0);
    } else {
        std::vector<std::thread> th;
        th.reserve(g->n);
        for (int r = 0; r < g->n; r++) th.emplace_back([&, r] { rc[r] = f(r); });
        for (auto &t : th) t.join();
    }
    for (int r = 0; r < g->n; r++)
        if (rc[r] != SHLL_OK) return gfail(g, rc[r], "%s: slab %d of %d: %s", what, r, g->n, shll_last_error(g->ctx[r]));
    return SHLL_OK;
}

void slab_ptrs(const shll_group *g, int r, float *const src[4], float *dst[4])
{
    for (int k = 0; k < 4; k++) dst[k] = (k < g->ncomp && src[k]) ? src[k] + (size_t)g->i0[r] * g->row : nullptr;
}

}  // namespace

extern "C" {

const char *shll_group_last_error(const shll_group *g) { return g ? g->err : g_group_error; }

int shll_group_size(const shll_group *g) { return g ? g->n : 0; }

shll_ctx *shll_group_ctx(const shll_group *g, int slab) { return (g && slab >= 0 && slab < g->n) ? g->ctx[slab] : nullptr; }

int shll_group_destroy(shll_group *g)
{
    if (!g) return SHLL_OK;
    // all slabs idle before any of them unmaps its neighbours
    for (shll_ctx *c : g->ctx)
        if (c) shll_sync(c);
    for (shll_ctx *c : g->ctx)
        if (c) shll_destroy(c);
    delete g;
    return SHLL_OK;
}

int shll_group_create(shll_group **out, const shll_config *cfg, int ngpus, const int *devices)
{
    if (!out || !cfg) return gfail(nullptr, SHLL_E_INVAL, "shll_group_create: null argument");
    *out = nullptr;
    if (cfg->struct_size != sizeof(shll_config)) return gfail(nullptr, SHLL_E_INVAL, "shll_group_create: shll_config.struct_size mismatch");
    if (ngpus < 1) return gfail(nullptr, SHLL_E_INVAL, "shll_group_create: ngpus = %d", ngpus);
    if (cfg->nx < ngpus) return gfail(nullptr, SHLL_E_INVAL, "shll_group_create: %d rows cannot be cut into %d slabs", cfg->nx, ngpus);
    if (!devices) {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            return gfail(nullptr, SHLL_E_CUDA, "no CUDA device available (%s); libshll_b200 has no CPU fallback",
                         e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        if (ngpus > ndev) return gfail(nullptr, SHLL_E_INVAL, "shll_group_create: %d slabs asked for, %d CUDA devices present", ngpus, ndev);
    }
    shll_group *g = new (std::nothrow) shll_group();
    if (!g) return gfail(nullptr, SHLL_E_NOMEM, "out of host memory");
    g->cfg = *cfg;
    g->n = ngpus;
    g->row = cfg->dims == 2 ? cfg->ny : 1;
    g->ncomp = cfg->dims == 1 ? 3 : 4;
    g->err[0] = 0;
    g->ctx.assign(ngpus, nullptr);
    g->i0.resize(ngpus);
    g->nx.resize(ngpus);
    for (int r = 0; r < ngpus; r++) {  // balanced contiguous partition: sizes differ by at most one row
        g->i0[r] = (int)(((long)r * cfg->nx) / ngpus);
        g->nx[r] = (int)(((long)(r + 1) * cfg->nx) / ngpus) - g->i0[r];
    }
    // steps per halo exchange round (include/shll_b200.h: halo_steps) -- one value for the whole chain of slabs
    const int halo_steps = shll_plan_halo_steps(cfg, ngpus);
    for (int r = 0; r < ngpus; r++) {
        shll_config c = *cfg;
        c.halo_steps = halo_steps;
        c.nx = g->nx[r];
        c.rank = r;
        c.nranks = ngpus;
        c.device = devices ? devices[r] : r;
        if (ngpus == 1 && !devices) c.device = cfg->device;
        int rc = shll_create(&g->ctx[r], &c);
        if (rc != SHLL_OK) {
            gfail(nullptr, rc, "shll_group_create: slab %d of %d (rows %d..%d on device %d): %s", r, ngpus, g->i0[r],
                  g->i0[r] + g->nx[r] - 1, c.device, shll_last_error(nullptr));
            shll_group_destroy(g);
            return rc;
        }
    }
    for (int r = 0; r + 1 < ngpus; r++) {
        shll_peer_desc lo, hi;
        int rc = shll_peer_export(g->ctx[r], &lo);
        if (rc == SHLL_OK) rc = shll_peer_export(g->ctx[r + 1], &hi);
        if (rc == SHLL_OK) rc = shll_peer_connect(g->ctx[r], +1, &hi);
        if (rc == SHLL_OK) rc = shll_peer_connect(g->ctx[r + 1], -1, &lo);
        if (rc != SHLL_OK) {
            const char *m = shll_last_error(g->ctx[r]);
            if (!m || !*m) m = shll_last_error(g->ctx[r + 1]);
            gfail(nullptr, rc, "shll_group_create: connecting slabs %d and %d: %s", r, r + 1, m);
            shll_group_destroy(g);
            return rc;
        }
    }
    *out = g;
    return SHLL_OK;
}

int shll_group_upload_u(shll_group *g, const float *const u[4])
{
    if (!g || !u) return gfail(g, SHLL_E_INVAL, "shll_group_upload_u: null argument");
    for (int k = 0; k < g->ncomp; k++)
        if (!u[k]) return gfail(g, SHLL_E_INVAL, "shll_group_upload_u: u[%d] is null", k);
    // every slab idle before the halos of a new state are pushed into its buffers
    int rc = for_all_slabs(g, "shll_sync", [&](int r) { return shll_sync(g->ctx[r]); });
    if (rc) return rc;
    return for_all_slabs(g, "shll_upload_u", [&](int r) {
        float *p[4];
        slab_ptrs(g, r, const_cast<float *const *>(u), p);
        return shll_upload_u(g->ctx[r], p);
    });
}

int shll_group_download_u(shll_group *g, float *const u[4])
{
    if (!g || !u) return gfail(g, SHLL_E_INVAL, "shll_group_download_u: null argument");
    return for_all_slabs(g, "shll_download_u", [&](int r) {
        float *p[4];
        slab_ptrs(g, r, u, p);
        return shll_download_u(g->ctx[r], p);
    });
}

int shll_group_download_p(shll_group *g, float *const p[4], float *a)
{
    if (!g || !p) return gfail(g, SHLL_E_INVAL, "shll_group_download_p: null argument");
    return for_all_slabs(g, "shll_download_p", [&](int r) {
        float *q[4];
        slab_ptrs(g, r, p, q);
        return shll_download_p(g->ctx[r], q, a ? a + (size_t)g->i0[r] * g->row : nullptr);
    });
}

int shll_group_run(shll_group *g, long nsteps)
{
    if (!g || nsteps < 0) return gfail(g, SHLL_E_INVAL, "shll_group_run: bad argument");
    return for_all_slabs(g, "shll_run", [&](int r) {
        int rc = shll_run(g->ctx[r], nsteps);
        return rc != SHLL_OK ? rc : shll_sync(g->ctx[r]);
    });
}

int shll_group_run_timed(shll_group *g, long nsteps, float *ms)
{
    if (!g || !ms || nsteps < 0) return gfail(g, SHLL_E_INVAL, "shll_group_run_timed: bad argument");
    std::vector<float> t(g->n, 0.0f);
    int rc = for_all_slabs(g, "shll_run_timed", [&](int r) { return shll_run_timed(g->ctx[r], nsteps, &t[r]); });
    if (rc) return rc;
    *ms = 0.0f;
    for (float v : t) *ms = v > *ms ? v : *ms;  // device time of the slowest slab
    return SHLL_OK;
}

int shll_group_max_cfl(shll_group *g, float *cfl)
{
    if (!g || !cfl) return gfail(g, SHLL_E_INVAL, "shll_group_max_cfl: null argument");
    std::vector<float> v(g->n, 0.0f);
    int rc = for_all_slabs(g, "shll_max_cfl", [&](int r) { return shll_max_cfl(g->ctx[r], &v[r]); });
    if (rc) return rc;
    *cfl = 0.0f;
    for (float x : v) *cfl = x > *cfl ? x : *cfl;
    return SHLL_OK;
}

int shll_group_conserved_sums(shll_group *g, double sums[4])
{
    if (!g || !sums) return gfail(g, SHLL_E_INVAL, "shll_group_conserved_sums: null argument");
    std::vector<double> v((size_t)g->n * 4, 0.0);
    int rc = for_all_slabs(g, "shll_conserved_sums", [&](int r) { return shll_conserved_sums(g->ctx[r], &v[(size_t)r * 4]); });
    if (rc) return rc;
    for (int k = 0; k < 4; k++) {
        sums[k] = 0.0;
        for (int r = 0; r < g->n; r++) sums[k] += v[(size_t)r * 4 + k];  // slab order: reproducible for a given ngpus
    }
    return SHLL_OK;
}

long shll_group_launch_count(const shll_group *g)
{
    long n = 0;
    if (g)
        for (shll_ctx *c : g->ctx) n += shll_launch_count(c);
    return n;
}

}  // extern "C"
