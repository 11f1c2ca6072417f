// shll_timeline.cpp -- launch / sync timeline of the launch-bound regime (BASELINE.json configs[1]: 65 536 cells, derived
// 1D 2nd-order program) from CUPTI activity records.  nsys is not in this image; CUPTI is (libcupti.so ships with the
// toolkit), and its CONCURRENT_KERNEL / RUNTIME records are what an Nsight Systems timeline is drawn from.
//
// For each way of issuing the steps -- plain launches, programmatic dependent launches, CUDA-graph replay, the persistent
// cooperative kernel -- the program marches NSTEPS steps through the C ABI (libshll_b200.so), collects every kernel record
// (GPU start / end time stamps) and every runtime-API record of the launching thread, and prints
//   kernels, mean kernel duration, mean GPU idle gap between consecutive kernels, GPU busy fraction,
//   launch API calls with their mean host duration, time in synchronisation calls,
// then writes the raw records to <out>/timeline_<mode>.csv (kind,name,start_ns,end_ns) for plotting.
//
// build: g++ -O2 tools/timeline/shll_timeline.cpp -Iinclude -I/usr/local/cuda/include -Lshll_sve_cfd_b200 -lshll_b200
//            -L/usr/local/cuda/lib64 -lcupti -Wl,-rpath,... -o tools/timeline/shll_timeline      (tools/timeline/Makefile)
// usage: shll_timeline [ncells=65536] [nsteps=4096] [outdir=.] [fast|strict]
#include <cupti.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/shll_b200.h"

struct Rec {
    int kind;  // 0 kernel, 1 runtime api
    std::string name;
    unsigned long long start, end;
    unsigned cbid;
};
static std::vector<Rec> g_recs;

static void CUPTIAPI buffer_requested(uint8_t **buffer, size_t *size, size_t *max_records)
{
    *size = 16u << 20;
    *buffer = static_cast<uint8_t *>(aligned_alloc(8, *size));
    *max_records = 0;
}
static void CUPTIAPI buffer_completed(CUcontext, uint32_t, uint8_t *buffer, size_t, size_t valid)
{
    CUpti_Activity *r = nullptr;
    while (cuptiActivityGetNextRecord(buffer, valid, &r) == CUPTI_SUCCESS) {
        if (r->kind == CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL || r->kind == CUPTI_ACTIVITY_KIND_KERNEL) {
            const CUpti_ActivityKernel9 *k = reinterpret_cast<const CUpti_ActivityKernel9 *>(r);
            g_recs.push_back({0, k->name ? k->name : "?", k->start, k->end, 0});
        } else if (r->kind == CUPTI_ACTIVITY_KIND_RUNTIME) {
            const CUpti_ActivityAPI *a = reinterpret_cast<const CUpti_ActivityAPI *>(r);
            const char *nm = nullptr;
            cuptiGetCallbackName(CUPTI_CB_DOMAIN_RUNTIME_API, a->cbid, &nm);
            g_recs.push_back({1, nm ? nm : "?", a->start, a->end, a->cbid});
        }
    }
    free(buffer);
}

static void ck(int rc, const char *what, shll_ctx *c)
{
    if (rc) { fprintf(stderr, "%s failed (%d): %s\n", what, rc, shll_last_error(c)); exit(1); }
}

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 65536;
    const long nsteps = argc > 2 ? atol(argv[2]) : 4096;
    const std::string out = argc > 3 ? argv[3] : ".";
    const bool fast = !(argc > 4 && !strcmp(argv[4], "strict"));
    struct Mode { const char *name, *persist, *graph, *pdl; } modes[] = {
        {"plain_launches", "0", "0", "0"}, {"programmatic_launches", "0", "0", "1"}, {"graph_replay", "0", "1", "1"}, {"persistent_kernel", "1", "1", "1"}};
    if (cuptiActivityRegisterCallbacks(buffer_requested, buffer_completed) != CUPTI_SUCCESS) { fprintf(stderr, "CUPTI unavailable\n"); return 2; }
    cuptiActivityEnable(CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL);
    cuptiActivityEnable(CUPTI_ACTIVITY_KIND_RUNTIME);

    // Sod tube, conserved variables exactly as the host program's Compute_U_from_P leaves them (host/shll_main.c)
    std::vector<float> u0(n), u1(n, 0.0f), u2(n);
    const float CV = 1.0 / (1.4f - 1.0);
    for (int i = 0; i < n; i++) { u0[i] = (i < 0.5 * n) ? 10.0f : 1.0f; u2[i] = u0[i] * (1.0f * CV); }
    const float *up[4] = {u0.data(), u1.data(), u2.data(), nullptr};

    printf("# launch / sync timeline from CUPTI activity records: %d cells, %ld steps, %s arithmetic\n", n, nsteps, fast ? "FAST" : "STRICT");
    printf("%-24s %8s %12s %12s %10s %10s %14s %12s %12s\n", "mode", "kernels", "kernel_us", "gap_us", "gpu_busy", "us/step", "launch_calls", "call_us", "sync_ms");
    for (const Mode &m : modes) {
        setenv("SHLL_PERSIST", m.persist, 1);
        setenv("SHLL_GRAPH", m.graph, 1);
        setenv("SHLL_PDL", m.pdl, 1);
        shll_config cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.struct_size = sizeof(cfg);
        cfg.dims = 1; cfg.nx = n; cfg.ny = 1; cfg.order = 2; cfg.bc = SHLL_BC_OUTFLOW; cfg.limiter = SHLL_LIM_MINMOD; cfg.alpha = 1.25f;
        cfg.tform = SHLL_TFORM_2D; cfg.mode = fast ? SHLL_MODE_FAST : SHLL_MODE_STRICT; cfg.dt_on_dx = 0.125f; cfg.dt_on_dy = 0.125f; cfg.nranks = 1;
        shll_ctx *c = nullptr;
        ck(shll_create(&c, &cfg), "shll_create", nullptr);
        ck(shll_upload_u(c, up), "shll_upload_u", c);
        ck(shll_run(c, 512), "warm-up", c);   // includes graph capture / first-launch costs
        ck(shll_sync(c), "shll_sync", c);
        cuptiActivityFlushAll(1);
        g_recs.clear();
        ck(shll_run(c, nsteps), "shll_run", c);
        ck(shll_sync(c), "shll_sync", c);
        cuptiActivityFlushAll(1);
        std::vector<Rec> ks, api;
        for (const Rec &r : g_recs) (r.kind == 0 ? ks : api).push_back(r);
        std::sort(ks.begin(), ks.end(), [](const Rec &a, const Rec &b) { return a.start < b.start; });
        double busy = 0, gap = 0;
        for (size_t i = 0; i < ks.size(); i++) {
            busy += (double)(ks[i].end - ks[i].start);
            if (i) gap += (double)ks[i].start - (double)ks[i - 1].end;   // negative when programmatic launches overlap
        }
        const double span = ks.empty() ? 0 : (double)(ks.back().end - ks.front().start);
        double launch_ns = 0, sync_ns = 0;
        long launches = 0;
        for (const Rec &r : api) {
            if (r.name.find("Launch") != std::string::npos) { launches++; launch_ns += (double)(r.end - r.start); }
            if (r.name.find("Synchronize") != std::string::npos) sync_ns += (double)(r.end - r.start);
        }
        printf("%-24s %8zu %12.3f %12.3f %9.1f%% %10.3f %14ld %12.3f %12.3f\n", m.name, ks.size(), ks.empty() ? 0 : busy / ks.size() * 1e-3,
               ks.size() > 1 ? gap / (ks.size() - 1) * 1e-3 : 0.0, span > 0 ? 100.0 * std::min(busy, span) / span : 0.0, span * 1e-3 / nsteps, launches,
               launches ? launch_ns / launches * 1e-3 : 0.0, sync_ns * 1e-6);
        const std::string path = out + "/timeline_" + m.name + ".csv";
        if (FILE *f = fopen(path.c_str(), "w")) {
            fprintf(f, "kind,name,start_ns,end_ns\n");
            const unsigned long long t0 = ks.empty() ? 0 : ks.front().start;
            size_t kept = 0;
            for (const Rec &r : ks) { if (kept++ < 2000) fprintf(f, "kernel,%.60s,%llu,%llu\n", r.name.c_str(), r.start - t0, r.end - t0); }
            kept = 0;
            for (const Rec &r : api) { if (r.start >= t0 && kept++ < 2000) fprintf(f, "api,%s,%llu,%llu\n", r.name.c_str(), r.start - t0, r.end - t0); }
            fclose(f);
        }
        shll_destroy(c);
    }
    return 0;
}
