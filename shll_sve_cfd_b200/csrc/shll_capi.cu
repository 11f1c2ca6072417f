// shll_capi.cu -- the C ABI of libshll_b200.so (include/shll_b200.h): context, device buffers, launches.
//
// Device memory of one context (one slab on one GPU):
//   state : 2 (ping-pong) x ncomp planes, FP32.  2D plane = (nx + 2*HALO_ROWS) rows of ny floats, rows -2,-1 and
//           nx,nx+1 being halo rows; 1D plane = PAD1D + roundup4(n) + PAD1D floats, cells -2,-1 / n,n+1 being halo
//           cells.  One cudaMalloc, so that one CUDA-IPC handle exposes both buffers to the neighbour slabs.
//   flags : halo-arrival flags / edge counters / error word (see halo_sync.cuh).
// There is no host fallback anywhere in this file: every entry point either runs CUDA work or returns an error.
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>
#include <unistd.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/shll_b200.h"
#include "shll_internal.h"

using namespace shll;

namespace {

constexpr int HALO_ROWS = 2;
// [0]=arrived from lower, [1]=arrived from upper, [2]=cnt_lo, [3]=cnt_hi, [4]=err, [8]=cfl, [16..21]=halo-wait statistics (3 x u64),
// [MAIL_OFF ..) = 1D halo mailboxes [side: 0 = filled by the lower neighbour, 1 = by the upper][round parity][3 comps][HALO1D_MAX]
constexpr int FLAG_WORDS = 1024;
constexpr int MAIL_OFF = 128;
constexpr int WAIT_STATS_OFF = 16;
constexpr int EARLY_MAX = 1536;  // blocks of a 1D step launch that may start on their neighbours' flags (~2.5 resident waves)

inline size_t mail_index(int side, unsigned round, int k) { return MAIL_OFF + (size_t)(((side * 2 + (int)(round & 1u)) * 3 + k) * HALO1D_MAX); }

thread_local char g_create_error[512] = "";

__global__ void fill_kernel(float *p, float v, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

}  // namespace

struct shll_ctx {
    shll_config cfg;
    int ncomp;
    long ncells;          // owned cells
    size_t plane_elems;   // floats per plane, including halos / padding
    size_t interior_off;  // floats from the plane start to owned cell 0
    float *state;         // 2 * ncomp * plane_elems floats
    unsigned *flags;
    float *scratch;       // lazily allocated: primitive download (ncomp + 1 planes of ncells)
    double *sums_dev;     // lazily allocated: per-block partial sums of shll_conserved_sums
    int cur;              // which ping-pong buffer holds the current state
    bool has_state;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    long launches;
    unsigned state_index;  // time steps taken since creation (monotonic; halo flags are expressed in it)
    unsigned epoch;        // step launches since creation
    // 1D slabs: temporal blocking of the halo exchange (step1d.cuh)
    int halo_K;            // steps per exchange round (1 = exchange every step)
    unsigned round;        // exchange rounds completed since creation (monotonic; the 1D halo flags are expressed in it)
    unsigned sends;        // send steps launched since creation
    unsigned origin;       // state_index at the last upload: rounds restart there
    // 1D: early start of the first wave of blocks (step1d.cuh)
    unsigned *done;        // per-block done flags, EARLY_MAX + 2 words
    unsigned early_epoch;  // value published by the last step launch that used the early-start protocol
    int early_prev_grid;   // ... and its grid size
    int sms;               // multiprocessors of the device
    bool fuse2;            // 2D FAST 1st order: two steps per launch (step2d_acc2_kernel) whenever at least two steps remain
    bool fuse1d;           // 1D FAST 2nd order: two steps per launch (step1d_acc2_kernel)
    KernelKey key;
    int ntiles, nchunks;
    CUtensorMap tmap[2];   // 2D TMA kernels: one 3D map {ny, nx+4, 4} per ping-pong buffer
    CUtensorMap tmap_out[2];  // face-flux kernel: the same tensors with the TMA-store box (60 owned columns x 4 rows x 4 planes)
    bool tma_store;
    CUtensorMap *tmap_dev; // the same two descriptors in device memory
    cudaGraphExec_t graph; // single-GPU small grids: GRAPH_STEPS consecutive steps captured once (launch-bound regime)
    bool graph_tried;
    bool capturing;        // launch_one_step is being recorded into the graph
    int graph_cur;         // ping-pong orientation the graph was captured with
    // persistent 1D march (persist1d.cuh)
    float *strips;
    unsigned *round_done;
    unsigned persist_rounds;  // value of round_done[] after the last persistent launch
    int persist_blocks, persist_threads, persist_K;
    bool persist_failed;      // allocation or cooperative launch failed once: use the per-step kernels from now on
    int tma_stages;
    size_t tma_smem;
    char variant[128];
    char err[512];
    // neighbours
    float *peer_state[2];     // [0] lower, [1] upper: base of the neighbour's state allocation (mapped)
    unsigned *peer_flags[2];
    bool peer_ipc[2];         // opened with cudaIpcOpenMemHandle (must be closed)
    shll_peer_desc peer_desc[2];
    bool connected[2];

    float *plane(int buf, int k) const { return state + ((size_t)buf * ncomp + k) * plane_elems + interior_off; }
};

namespace {

int fail(shll_ctx *c, int code, const char *fmt, ...)
{
    char *dst = c ? c->err : g_create_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(ctx, call)                                                                                        \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return fail(ctx, SHLL_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

bool is_pow2(float x)
{
    int e;
    return x > 0.0f && std::frexp(x, &e) == 0.5f;
}

int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// Choose vector width, tiles and row chunks for the 2D kernel.  no_tma: the caller has already found that the TMA
// descriptors cannot be made (plan again for the LDG kernel; never decided through the environment).
void plan_2d(shll_ctx *c, bool no_tma = false)
{
    const shll_config &g = c->cfg;
    int vec = g.variant > 0 ? g.variant : env_int("SHLL_VEC", 0);
    if (vec != 1 && vec != 2 && vec != 4) vec = (g.order == 1) ? 2 : 1;  // defaults from the B200 sweep in DESIGN.md
    while (vec > 1 && (g.ny % vec != 0 || g.ny < 32 * vec)) vec >>= 1;
    // TMA-fed kernel (step2d_tma.cuh) whenever the row stride is a multiple of 16 bytes; SHLL_TMA=0 forces the LDG kernel.
    c->key.tma = !no_tma && (g.ny % 4 == 0) && (g.ny >= 32) && env_int("SHLL_TMA", 1) != 0;
    if (c->key.tma) {
        // defaults from the B200 sweep (profiles/r01_sweep_2d.log): order 1 FAST 2 cells per lane, everything else 1
        if (g.variant <= 0 && env_int("SHLL_VEC", 0) <= 0) vec = (g.mode == SHLL_MODE_FAST) ? 2 : 1;
        if (vec > 2) vec = 2;  // instantiated TMA widths
        while (vec > 1 && (g.ny % (4 * vec) != 0 || g.ny < 32 * vec)) vec >>= 1;
    }
    c->key.vec = vec;
    // FAST arithmetic, 2 cells per lane: the face-flux accumulate kernel (step2d_acc.cuh); SHLL_ACC=0 keeps the window kernel.
    c->key.acc = c->key.tma && vec == 2 && g.mode == SHLL_MODE_FAST && env_int("SHLL_ACC", 1) != 0;
    c->key.acc_cfg = 1;  // (round-1/2 register-cap and stash experiments: only the winning variant is instantiated, step2d_acc_o2.cu)
    // two steps per launch (step2d_acc.cuh: step2d_acc2_kernel): single slab -> decided here; slabs -> only when the front end asked
    // for it on EVERY slab (halo_steps = 2), because neighbouring slabs must issue the same sequence of launches
    c->fuse2 = c->key.acc && g.order == 1 && g.nx >= 8 && env_int("SHLL_FUSE2", 1) != 0 && (g.nranks == 1 || g.halo_steps == 2);
    const int hl = (g.order + vec - 1) / vec;
    const int useful = (32 - 2 * hl) * vec;
    c->ntiles = (g.ny + useful - 1) / useful;
    // B200 sweeps (profiles/).  1st-order FAST: 18 rows = 4 whole boxes + the two halo rows of a one-step launch; with two-step
    // launches (4 halo rows per chunk) the curve is flat from 28 rows up, 44 rows = 11 whole boxes is its best point (229.9 Gcu/s at
    // 4096^2, profiles/r02_fused_two_step.log)
    int rpc = env_int("SHLL_ROWS_PER_CHUNK", c->key.tma ? (g.order == 1 ? (c->key.acc ? (c->fuse2 ? 44 : 18) : 24) : 64) : 64);
    if (c->key.tma && env_int("SHLL_ROWS_PER_CHUNK", 0) <= 0) {
        // Small and medium grids: the tuned chunk height leaves most of the GPU without a warp (256^2, 2nd order: 20 one-warp
        // blocks for 148 SMs).  Shrink the chunks, a TMA box of rows at a time, until about half of the resident warp slots
        // have an item -- thinner chunks recompute more halo rows, which only matters once the GPU is full
        // (profiles/r01_sweep_chunk_height_small_grids.log: 256^2 order 2 37.9 -> 9.4 us per step, 1024^2 order 2 42.6 -> 20.3).
        const int box_rows = (g.order == 1 && !c->key.acc) ? 3 : 4;
        const int lowest = c->key.acc ? ((g.order == 1 && !c->fuse2) ? 2 : 4) : (g.order == 1 ? 6 : 4);
        int sms = 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g.device) != cudaSuccess || sms < 1) { (void)cudaGetLastError(); sms = 148; }
        // Blocks wanted before the chunks stop shrinking, in units of one resident wave of one-warp blocks (B200 sweep over the
        // reference's own grid sizes 256^2 .. 2048^2 and the full sizes, profiles/r02_sweep_chunk_height_reference_sizes.log):
        // FAST kernels are content with 3/4 (order 1) or 1/2 (order 2) of a wave; the STRICT kernels, whose rows take ~3x longer,
        // balance best with ~5 (order 1: 1024^2 57.7 -> 34.7 us per step, 2048^2 96.5 -> 78.3) or ~2.25 (order 2: 1024^2 85 -> 51) waves.
        // The two-step 1st-order kernel (12 warps per SM) is best with ~0.6 of its wave (1024^2: 16-row chunks 124.5 Gcu/s, 8 rows 102).
        const long wave = (long)sms * ((g.order == 1 && !c->fuse2) ? 16 : 12);
        const long want = (g.mode == SHLL_MODE_FAST && c->key.acc) ? (g.order == 1 ? (c->fuse2 ? wave * 3 / 5 : wave * 3 / 4) : wave / 2)
                                                                    : (g.order == 1 ? wave * 5 : wave * 9 / 4);
        while (rpc - box_rows >= lowest && (long)c->ntiles * ((g.nx + rpc - 1) / rpc) < want) rpc -= box_rows;
    }
    if (rpc < 2) rpc = 2;
    int nchunks = (g.nx + rpc - 1) / rpc;
    if (env_int("SHLL_NCHUNKS", 0) > 0) nchunks = env_int("SHLL_NCHUNKS", 0);  // experiments: chunk count given directly
    if (nchunks < 1) nchunks = 1;
    while (nchunks > 1 && g.nx / nchunks < 2) nchunks--;
    c->nchunks = nchunks;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3D tensor maps {ny, nx+4, 4 planes} over each ping-pong buffer, origin at halo row -2 of plane 0; box = the warp's
// 32*vec columns x R rows x 4 planes.  The driver entry point is looked up at run time (no link dependency on libcuda).
int make_tensor_maps(shll_ctx *c)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || fn == nullptr || q != cudaDriverEntryPointSuccess) return -1;
    const shll_config &g = c->cfg;
    const int R = (g.order == 1 && !c->key.acc) ? 3 : 4;  // rows per box == unroll factor of the kernel's row loop
    for (int b = 0; b < 2; b++) {
        float *base = c->state + (size_t)b * c->ncomp * c->plane_elems;  // plane 0, halo row -2
        cuuint64_t dims[3] = {(cuuint64_t)g.ny, (cuuint64_t)(g.nx + 4), 4};
        cuuint64_t strides[2] = {(cuuint64_t)g.ny * 4, (cuuint64_t)c->plane_elems * 4};
        cuuint32_t box[3] = {(cuuint32_t)(32 * c->key.vec + 4), (cuuint32_t)R, 4};  // 4 spare columns: 16-byte aligned box start
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = ((encode_tiled_fn)fn)(&c->tmap[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return -(int)r - 100;
        if (c->key.acc) {
            cuuint32_t obox[3] = {60, 4, 4};
            r = ((encode_tiled_fn)fn)(&c->tmap_out[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, obox, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return -(int)r - 200;
        }
    }
    c->tma_store = c->key.acc && env_int("SHLL_TMA_STORE", 1) != 0;
    if (env_int("SHLL_TMAP_GLOBAL", 0)) {
        if (cudaMalloc(&c->tmap_dev, 2 * sizeof(CUtensorMap)) != cudaSuccess) return -2;
        if (cudaMemcpy(c->tmap_dev, c->tmap, 2 * sizeof(CUtensorMap), cudaMemcpyHostToDevice) != cudaSuccess) return -3;
    }
    c->tma_stages = env_int("SHLL_TMA_STAGES", c->key.acc ? 2 : (c->key.vec == 2 ? 3 : 4));
    if (c->tma_stages < 2) c->tma_stages = 2;
    if (c->tma_stages > 16) c->tma_stages = 16;
    const size_t stage_stride = ((size_t)4 * R * (32 * c->key.vec + 4) * 4 + 127) & ~(size_t)127;
    c->tma_smem = (size_t)c->tma_stages * stage_stride + 8 * c->tma_stages;
    if (c->tma_store) c->tma_smem += 3840 + 128;                   // per-warp store stage (step2d_acc.cuh: ACC_OUT_STAGE_BYTES)
    return 0;
}

void plan_1d(shll_ctx *c)
{
    c->key.vec = 4;
    c->key.acc = c->cfg.mode == SHLL_MODE_FAST && c->cfg.order == 2 && env_int("SHLL_ACC", 1) != 0;
    c->key.acc_cfg = env_int("SHLL_ACC_CFG", 0);
    // two steps per launch (step1d_acc.cuh): slabs need an even number of steps per exchange round so that a launch never straddles one
    c->fuse1d = c->key.acc && env_int("SHLL_FUSE1D", 1) != 0 && (c->cfg.nranks == 1 || c->cfg.halo_steps % 2 == 0);
    c->ntiles = (c->cfg.nx + 119) / 120;
    c->nchunks = 1;
}

}  // namespace

extern "C" {

int shll_abi_version(void) { return SHLL_ABI_VERSION; }

const char *shll_last_error(const shll_ctx *ctx) { return ctx ? ctx->err : g_create_error; }

int shll_count_steps(float dt, float total_time, long *nsteps)
{
    if (!nsteps || !(dt > 0.0f)) return fail(nullptr, SHLL_E_INVAL, "shll_count_steps: bad arguments");
    // base_shll.c:200,208,216-217 -- the reference's clock is a float accumulator.
    volatile float t = 0.0f;
    long n = 0;
    while (t < total_time) {
        float tn = t + dt;
        if (tn == t) return fail(nullptr, SHLL_E_INVAL, "float clock stalls at t=%g with dt=%g (the reference would never terminate)", (double)t, (double)dt);
        t = tn;
        n++;
    }
    *nsteps = n;
    return SHLL_OK;
}

int shll_plan_halo_steps(const shll_config *whole, int nslabs)
{
    if (!whole || nslabs <= 1 || whole->struct_size != sizeof(shll_config)) return 1;
    const int order = whole->order == 2 ? 2 : 1;
    const int smallest = whole->nx / nslabs;  // balanced partition: slab sizes are floor or ceil of nx / nslabs
    if (whole->dims == 1) {
        int k = whole->halo_steps > 0 ? whole->halo_steps : env_int("SHLL_HALO_K", 16);
        if (k > HALO1D_MAX / order) k = HALO1D_MAX / order;
        if (k > smallest / order) k = smallest / order;
        return k < 1 ? 1 : k;
    }
    // 2D: two-step launches where the 1st-order FAST face-flux kernel would be selected for a slab of the smallest size
    if (whole->halo_steps == 1 || smallest < 16) return 1;
    shll_ctx probe;
    memset(&probe, 0, sizeof(probe));
    probe.cfg = *whole;
    probe.cfg.nx = smallest;
    probe.cfg.nranks = nslabs;
    probe.cfg.halo_steps = 2;
    if (probe.cfg.tform == SHLL_TFORM_AUTO) probe.cfg.tform = SHLL_TFORM_2D;
    probe.key.order = probe.cfg.order; probe.key.mode = probe.cfg.mode;
    plan_2d(&probe);
    return probe.fuse2 ? 2 : 1;
}

int shll_create(shll_ctx **out, const shll_config *cfg)
{
    if (!out || !cfg) return fail(nullptr, SHLL_E_INVAL, "shll_create: null argument");
    *out = nullptr;
    if (cfg->struct_size != sizeof(shll_config))
        return fail(nullptr, SHLL_E_INVAL, "shll_create: struct_size %u != %zu (ABI mismatch)", cfg->struct_size, sizeof(shll_config));
    shll_config g = *cfg;
    if (g.dims != 1 && g.dims != 2) return fail(nullptr, SHLL_E_INVAL, "dims must be 1 or 2");
    if (g.dims == 1) g.ny = 1;
    if (g.nx < 2 || (g.dims == 2 && g.ny < 2)) return fail(nullptr, SHLL_E_INVAL, "grid too small: nx=%d ny=%d", g.nx, g.ny);
    if (g.order != 1 && g.order != 2) return fail(nullptr, SHLL_E_INVAL, "order must be 1 or 2");
    if (g.bc != SHLL_BC_REFLECT && g.bc != SHLL_BC_OUTFLOW) return fail(nullptr, SHLL_E_INVAL, "bad bc");
    if (g.limiter != SHLL_LIM_MINMOD && g.limiter != SHLL_LIM_MC) return fail(nullptr, SHLL_E_INVAL, "bad limiter");
    if (g.mode != SHLL_MODE_STRICT && g.mode != SHLL_MODE_FAST) return fail(nullptr, SHLL_E_INVAL, "bad mode");
    if (g.nranks < 1) g.nranks = 1;
    if (g.rank < 0 || g.rank >= g.nranks) return fail(nullptr, SHLL_E_INVAL, "rank %d outside [0,%d)", g.rank, g.nranks);
    if (g.nranks > 1 && g.nx < 2 * g.order) return fail(nullptr, SHLL_E_INVAL, "slab of %d rows is thinner than two halos", g.nx);
    if (g.tform == SHLL_TFORM_AUTO) g.tform = (g.dims == 1 && g.order == 1) ? SHLL_TFORM_1D : SHLL_TFORM_2D;
    if (g.dims == 2) g.tform = SHLL_TFORM_2D;
    if (!(g.dt_on_dx > 0.0f) || (g.dims == 2 && !(g.dt_on_dy > 0.0f))) return fail(nullptr, SHLL_E_INVAL, "dt_on_dx / dt_on_dy must be positive");
    if (g.halo_steps <= 0 || g.nranks == 1) g.halo_steps = 1;
    if (g.dims == 2 && g.halo_steps > 2) g.halo_steps = 2;  // 2D: 2 = two-step launches (1st-order FAST kernel), 2-row exchange
    if (g.dims == 1 && (g.halo_steps * g.order > HALO1D_MAX || g.halo_steps * g.order > g.nx))
        return fail(nullptr, SHLL_E_INVAL, "halo_steps = %d: %d halo cells per side exceed the limit of %d or the slab's %d cells", g.halo_steps,
                    g.halo_steps * g.order, HALO1D_MAX, g.nx);

    if ((double)(g.nx + 8) * (double)g.ny >= 2147483647.0)
        return fail(nullptr, SHLL_E_INVAL, "slab of %d x %d cells exceeds the 32-bit plane index used by the kernels; use more slabs", g.nx, g.ny);

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, SHLL_E_CUDA, "no CUDA device available (%s); libshll_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (g.device < 0 || g.device >= ndev) return fail(nullptr, SHLL_E_INVAL, "device %d outside [0,%d)", g.device, ndev);

    shll_ctx *c = new (std::nothrow) shll_ctx();
    if (!c) return fail(nullptr, SHLL_E_NOMEM, "out of host memory");
    memset(c, 0, sizeof(*c));
    c->cfg = g;
    c->ncomp = (g.dims == 1) ? 3 : 4;
    c->ncells = (long)g.nx * g.ny;
    if (g.dims == 1) {
        c->interior_off = PAD1D;
        c->plane_elems = round_up((size_t)PAD1D + round_up((size_t)g.nx, 4) + PAD1D + 4, 64);
    } else {
        c->interior_off = (size_t)HALO_ROWS * g.ny;
        c->plane_elems = round_up((size_t)(g.nx + 2 * HALO_ROWS) * g.ny, 64);
    }
    c->halo_K = g.halo_steps;
    c->sms = 148;
    if (cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, g.device) != cudaSuccess || c->sms < 1) { (void)cudaGetLastError(); c->sms = 148; }
    c->key.order = g.order;
    c->key.bc = g.bc;
    c->key.lim = g.limiter;
    c->key.mode = g.mode;
    c->key.tform = g.tform;
    c->key.pow2 = is_pow2(g.dt_on_dx) && (g.dims == 1 || is_pow2(g.dt_on_dy));
    if (g.dims == 1) plan_1d(c); else plan_2d(c);
    snprintf(c->variant, sizeof(c->variant), "step%dd_o%d_%s_%s%s_%s_vec%d_tiles%d_chunks%d", g.dims, g.order,
             g.bc == SHLL_BC_REFLECT ? "reflect" : "outflow", g.order == 2 ? (g.limiter == SHLL_LIM_MC ? "mc_" : "minmod_") : "",
             g.mode == SHLL_MODE_STRICT ? "strict" : "fast", c->key.pow2 ? "pow2" : "gendt", c->key.vec, c->ntiles, c->nchunks);

#define CKC(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            int rc_ = fail(nullptr, e_ == cudaErrorMemoryAllocation ? SHLL_E_NOMEM : SHLL_E_CUDA, \
                           "%s failed: %s", #call, cudaGetErrorString(e_));                       \
            shll_destroy(c);                                                                      \
            return rc_;                                                                           \
        }                                                                                         \
    } while (0)
    CKC(cudaSetDevice(g.device));
    CKC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CKC(cudaEventCreate(&c->ev0));
    CKC(cudaEventCreate(&c->ev1));
    const size_t state_bytes = (size_t)2 * c->ncomp * c->plane_elems * sizeof(float);
    CKC(cudaMalloc(&c->state, state_bytes));
    CKC(cudaMalloc(&c->flags, FLAG_WORDS * sizeof(unsigned)));
    CKC(cudaMemsetAsync(c->flags, 0, FLAG_WORDS * sizeof(unsigned), c->stream));
    if (g.dims == 1) {
        CKC(cudaMalloc(&c->done, (EARLY_MAX + 2) * sizeof(unsigned)));
        CKC(cudaMemsetAsync(c->done, 0, (EARLY_MAX + 2) * sizeof(unsigned), c->stream));
    }
    // Fill everything (halos, padding) with a benign gas state (1.0f bit pattern) so unused lanes never see NaNs.
    fill_kernel<<<1024, 256, 0, c->stream>>>(c->state, 1.0f, state_bytes / sizeof(float));
    CKC(cudaGetLastError());
    CKC(cudaStreamSynchronize(c->stream));
    if (g.dims == 2 && c->key.tma) {
        int rc = make_tensor_maps(c);
        if (rc != 0) {  // not fatal: the LDG kernel computes the same bits
            plan_2d(c, /*no_tma=*/true);
        }
    }
    if (g.dims == 2 && g.nranks > 1 && g.halo_steps == 2 && !c->fuse2) {
        int rc_ = fail(nullptr, SHLL_E_INVAL, "halo_steps = 2 needs the 2D 1st-order FAST kernel (ny %% 8 == 0, nx >= 8, SHLL_FUSE2 != 0)");
        shll_destroy(c);
        return rc_;
    }
    if (g.dims == 2 && c->key.tma) {  // early-start flags, one per (chunk, tile) block of a step launch (step2d_tma.cuh)
        const size_t nflags = (size_t)c->ntiles * c->nchunks;
        CKC(cudaMalloc(&c->done, nflags * sizeof(unsigned)));
        CKC(cudaMemsetAsync(c->done, 0, nflags * sizeof(unsigned), c->stream));
        CKC(cudaStreamSynchronize(c->stream));
    }
    snprintf(c->variant, sizeof(c->variant), "step%dd%s_o%d_%s_%s%s_%s_vec%d_tiles%d_chunks%d", g.dims, (g.dims == 2 && c->key.tma) ? (c->key.acc ? "_tma_acc" : "_tma") : ((g.dims == 1 && c->key.acc) ? "_acc" : ""),
             g.order, g.bc == SHLL_BC_REFLECT ? "reflect" : "outflow", g.order == 2 ? (g.limiter == SHLL_LIM_MC ? "mc_" : "minmod_") : "",
             g.mode == SHLL_MODE_STRICT ? "strict" : "fast", c->key.pow2 ? "pow2" : "gendt", c->key.vec, c->ntiles, c->nchunks);
    if (c->fuse2 || c->fuse1d) strncat(c->variant, "_x2", sizeof(c->variant) - strlen(c->variant) - 1);  // two time steps per launch
#undef CKC
    *out = c;
    return SHLL_OK;
}

int shll_destroy(shll_ctx *c)
{
    if (!c) return SHLL_OK;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int s = 0; s < 2; s++) {
        if (c->connected[s] && c->peer_ipc[s]) {
            if (c->peer_state[s]) cudaIpcCloseMemHandle(c->peer_state[s]);
            if (c->peer_flags[s]) cudaIpcCloseMemHandle(c->peer_flags[s]);
        }
    }
    if (c->scratch) cudaFree(c->scratch);
    if (c->tmap_dev) cudaFree(c->tmap_dev);
    if (c->sums_dev) cudaFree(c->sums_dev);
    if (c->graph) cudaGraphExecDestroy(c->graph);
    if (c->done) cudaFree(c->done);
    if (c->strips) cudaFree(c->strips);
    if (c->round_done) cudaFree(c->round_done);
    if (c->state) cudaFree(c->state);
    if (c->flags) cudaFree(c->flags);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return SHLL_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ halo push
namespace {

// Copies this slab's edge rows / cells of the *current* buffer into the neighbours' halo storage and raises their
// flags.  Used once after shll_upload_u (the step kernels do it themselves afterwards).
__global__ void push_halo_kernel(const float *src_lo, const float *src_hi, float *dst_lo, float *dst_hi, long count,
                                 int ncomp, size_t src_plane, size_t dst_plane_lo, size_t dst_plane_hi,
                                 unsigned *flag_lo, unsigned *flag_hi, unsigned value)
{
    for (int k = 0; k < ncomp; k++) {
        for (long t = threadIdx.x; t < count; t += blockDim.x) {
            if (dst_lo) dst_lo[k * dst_plane_lo + t] = src_lo[k * src_plane + t];
            if (dst_hi) dst_hi[k * dst_plane_hi + t] = src_hi[k * src_plane + t];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (flag_lo) st_release_sys(flag_lo, value);
        if (flag_hi) st_release_sys(flag_hi, value);
    }
}

// geometry of a neighbour's plane, from its descriptor
size_t peer_plane_elems(const shll_peer_desc &d)
{
    if (d.dims == 1) return round_up((size_t)PAD1D + round_up((size_t)d.nx, 4) + PAD1D + 4, 64);
    return round_up((size_t)(d.nx + 2 * HALO_ROWS) * d.ny, 64);
}
size_t peer_interior_off(const shll_peer_desc &d) { return d.dims == 1 ? (size_t)PAD1D : (size_t)HALO_ROWS * d.ny; }

// Pointer to the place in neighbour `side`'s buffer `buf`, plane k, where OUR edge data must land.
//   side 0 (lower neighbour): its upper halo = rows nx_peer .. ;  side 1 (upper neighbour): its lower halo = rows -order ..
float *peer_halo_ptr(const shll_ctx *c, int side, int buf, int k)
{
    const shll_peer_desc &d = c->peer_desc[side];
    const size_t pe = peer_plane_elems(d);
    float *base = c->peer_state[side] + ((size_t)buf * c->ncomp + k) * pe + peer_interior_off(d);
    const long row = c->cfg.dims == 1 ? 1 : c->cfg.ny;
    if (side == 0) return base + (long)d.nx * row;
    return base - (long)(c->fuse2 ? 2 : c->cfg.order) * row;
}

bool multi(const shll_ctx *c) { return c->cfg.nranks > 1; }

int check_connected(shll_ctx *c)
{
    if (!multi(c)) return SHLL_OK;
    if (c->cfg.rank > 0 && !c->connected[0]) return fail(c, SHLL_E_STATE, "rank %d: lower neighbour not connected (shll_peer_connect)", c->cfg.rank);
    if (c->cfg.rank < c->cfg.nranks - 1 && !c->connected[1]) return fail(c, SHLL_E_STATE, "rank %d: upper neighbour not connected", c->cfg.rank);
    return SHLL_OK;
}

// Programmatic dependent launch of the step kernels (halo_sync.cuh), outside stream capture.  Slabs too: a block of step
// n+1 that is placed early touches nothing but its own halo flag before griddepcontrol.wait returns, i.e. before step n of
// THIS GPU has completed -- the neighbour-facing protocol (halo_sync.cuh) sees the same order of events as without it.
int use_pdl(const shll_ctx *c)
{
    if (multi(c) && env_int("SHLL_PDL_MULTI", 1) == 0) return 0;
    // 1D step kernels let their successor be placed BEFORE their halo wait (step1d.cuh: launch_dependents first).  When two
    // slabs share one device (shll_group with a device listed twice) the parked successors of a slab that is waiting for
    // its neighbour could take the SM slots that neighbour needs: no programmatic launch there.  (2D kernels wait for the
    // halo before launch_dependents, so at most ~2 of their grids are ever resident.)
    if (multi(c) && c->cfg.dims == 1)
        for (int s = 0; s < 2; s++)
            if (c->connected[s] && c->peer_desc[s].pid == (int64_t)getpid() && c->peer_desc[s].device == c->cfg.device) return 0;
    return ((!c->capturing || env_int("SHLL_PDL_GRAPH", 0) != 0) && env_int("SHLL_PDL", 1) != 0) ? 1 : 0;
}

// nsub: time steps this launch advances (2: the two-step kernel of a fuse2 context)
int launch_one_step(shll_ctx *c, int nsub = 1)
{
    const shll_config &g = c->cfg;
    const int in = c->cur, outb = c->cur ^ 1;
    const bool lo_wall = (g.rank == 0), hi_wall = (g.rank == g.nranks - 1);
    HaloSync S;
    memset(&S, 0, sizeof(S));
    if (multi(c)) {
        S.enabled = 1;
        S.want = c->state_index + 1;          // flag value = state index of the halo data + 1
        S.post = c->state_index + nsub + 1;
        S.epoch = c->epoch + 1;
        S.wait_lo = lo_wall ? nullptr : c->flags + 0;
        S.wait_hi = hi_wall ? nullptr : c->flags + 1;
        S.sig_lo = lo_wall ? nullptr : c->peer_flags[0] + 1;  // we are the lower neighbour's UPPER neighbour
        S.sig_hi = hi_wall ? nullptr : c->peer_flags[1] + 0;
        S.cnt_lo = c->flags + 2;
        S.cnt_hi = c->flags + 3;
        S.err = c->flags + 4;
        S.timeout_ns = (unsigned long long)env_int("SHLL_HALO_TIMEOUT_MS", 5000) * 1000000ull;
        S.wait_ns = reinterpret_cast<unsigned long long *>(c->flags + WAIT_STATS_OFF);
    }
    bool sent = false, early2d = false;
    cudaError_t e;
    if (g.dims == 2) {
        Step2DParams P;
        memset(&P, 0, sizeof(P));
        for (int k = 0; k < 4; k++) {
            P.in[k] = c->plane(in, k);
            P.out[k] = c->plane(outb, k);
            P.lo_peer[k] = (multi(c) && !lo_wall) ? peer_halo_ptr(c, 0, outb, k) : nullptr;
            P.hi_peer[k] = (multi(c) && !hi_wall) ? peer_halo_ptr(c, 1, outb, k) : nullptr;
        }
        P.nx = g.nx; P.ny = g.ny;
        P.lo_wall = lo_wall; P.hi_wall = hi_wall;
        P.ntiles = c->ntiles; P.nchunks = c->nchunks;
        P.dtdx = g.dt_on_dx; P.dtdy = g.dt_on_dy;
        P.half_dtdx = 0.5f * g.dt_on_dx; P.half_dtdy = 0.5f * g.dt_on_dy;
        P.alpha = g.alpha;
        P.quarter = 0.25f;
        P.peer_depth = c->fuse2 ? 2 : g.order;
        S.edge_warps_lo = S.edge_warps_hi = (unsigned)c->ntiles;
        P.sync = S;
        const int warps = c->ntiles * c->nchunks;
        if (c->key.tma) {
            Step2DTmaParams T;
            T.base = P;
            T.tmap = c->tmap[in];
            T.tmap_out = c->tmap_out[outb];
            T.tma_store = c->tma_store ? 1 : 0;
            T.tmap_global = c->tmap_dev ? c->tmap_dev + in : nullptr;
            T.stages = c->tma_stages;
            // consecutive steps of a single-GPU run overlap their launch with the predecessor's tail (step2d_acc.cu)
            T.pdl = use_pdl(c);
            T.early_blocks = 0; T.early_want = 0; T.early_post = 0; T.done = nullptr; T.early_err = c->flags + 4;
            // Measured (profiles/r02_early_start.log), with the publishers restricted to the blocks an early block can depend on
            // (with every block publishing, the release fences cost the 14 us blocks of the 1st-order FAST kernel more than the
            // overlapped ramp-up saved: 89.3 -> 97.6 us): FAST 2nd order 112.8 -> 117.0 Gcu/s, FAST 1st order 188.1 -> 190.6,
            // STRICT 80.0 -> 86.8 / 42.4 -> 45.0.  Only on grids of at least 4 resident waves; launch-bound small grids keep the
            // plain wait.  SHLL_EARLY=2 forces it on, 0 off.
            const int early_mode = env_int("SHLL_EARLY", 1);
            const int wave = c->sms * ((g.order == 1 && !c->fuse2) ? 16 : 12);   // resident one-warp blocks
            const bool long_blocks = warps >= 4 * wave;
            if (T.pdl && c->done && (early_mode == 2 || (early_mode == 1 && long_blocks))) {
                T.early_blocks = env_int("SHLL_EARLY_BLOCKS", wave * 3 / 2);
                if (T.early_blocks > warps) T.early_blocks = warps;
                T.early_want = c->early_epoch;
                T.early_post = c->early_epoch + 1;
                T.done = c->done;
            }
            dim3 grid(warps);
            early2d = (T.done != nullptr);
            T.early_diag = (nsub == 2) ? 1 : 0;
            if (nsub == 2) e = launch_step2d_acc2(c->key, T, grid, c->tma_smem, c->stream);
            else if (c->key.acc) e = launch_step2d_acc(c->key, T, grid, c->tma_smem, c->stream);
            else if (g.order == 1) e = launch_step2d_tma_o1(c->key, T, grid, c->tma_smem, c->stream);
            else if (g.mode == SHLL_MODE_STRICT) e = launch_step2d_tma_o2_strict(c->key, T, grid, c->tma_smem, c->stream);
            else e = launch_step2d_tma_o2_fast(c->key, T, grid, c->tma_smem, c->stream);
        } else {
            dim3 block(128), grid((warps + 3) / 4);
            if (g.order == 1) e = launch_step2d_o1(c->key, P, grid, block, c->stream);
            else if (g.mode == SHLL_MODE_STRICT) e = launch_step2d_o2_strict(c->key, P, grid, block, c->stream);
            else e = launch_step2d_o2_fast(c->key, P, grid, block, c->stream);
        }
    } else {
        Step1DParams P;
        memset(&P, 0, sizeof(P));
        // Slabs: temporal blocking of the halo exchange (step1d.cuh).  Position of this step in its round of K steps:
        const bool m = multi(c);
        const int K = m ? c->halo_K : 1, H = K * g.order;
        const int pos = m ? (int)((c->state_index - c->origin) % (unsigned)K) : 0;
        const bool recv = m && pos == 0, send = m && pos + nsub - 1 == K - 1;
        // halo cells this launch still updates: the range of its FIRST step (a two-step launch stores the same range; what lies
        // outside the second step's range is garbage that no later step of the round reads)
        const int ext = (!m || H - g.order * (pos + 1) <= 0) ? 0 : (int)round_up((size_t)(H - g.order * (pos + 1)), 4);
        const int ext_lo = lo_wall ? 0 : ext, ext_hi = hi_wall ? 0 : ext;
        for (int k = 0; k < 3; k++) {
            P.in[k] = c->plane(in, k) - ext_lo;
            P.out[k] = c->plane(outb, k) - ext_lo;
            // destination = the neighbours' mailboxes of the round that CONSUMES this data (round + 1)
            P.lo_peer[k] = (send && !lo_wall) ? reinterpret_cast<float *>(c->peer_flags[0] + mail_index(1, c->round + 1, k)) : nullptr;
            P.hi_peer[k] = (send && !hi_wall) ? reinterpret_cast<float *>(c->peer_flags[1] + mail_index(0, c->round + 1, k)) : nullptr;
        }
        P.n = g.nx + ext_lo + ext_hi;
        P.n_real = g.nx;
        P.ext_lo = ext_lo;
        P.recv = recv ? 1 : 0;
        P.xch = send ? H : 0;
        P.hcells = H;
        P.mail_lo = (recv && !lo_wall) ? reinterpret_cast<const float *>(c->flags + mail_index(0, c->round, 0)) : nullptr;
        P.mail_hi = (recv && !hi_wall) ? reinterpret_cast<const float *>(c->flags + mail_index(1, c->round, 0)) : nullptr;
        P.interior_end = P.n;
        if (recv && !hi_wall && ext_lo + g.nx < P.interior_end) P.interior_end = ext_lo + g.nx;  // tiles reading the upper halo cells
        if (send && !hi_wall && g.nx - H + ext_lo < P.interior_end) P.interior_end = g.nx - H + ext_lo;  // tiles owning cells to be sent
        P.lo_wall = lo_wall; P.hi_wall = hi_wall;
        P.ntiles = (P.n + 119) / 120;
        P.dtdx = g.dt_on_dx; P.half_dtdx = 0.5f * g.dt_on_dx; P.alpha = g.alpha;
        P.quarter = 0.25f;
        // 4 (face-flux kernel: 8) consecutive tiles per warp on large tubes (B200 sweep: profiles/); small tubes keep one warp
        // per tile so that all SMs work
        const int tpw_big = c->key.acc ? 8 : 4;
        const int ntiles_real = (g.nx + 119) / 120;  // (from the slab itself, not from this step's extended range: the same every step)
        P.tiles_per_warp = env_int("SHLL_1D_TILES_PER_WARP", ntiles_real >= tpw_big * 4096 ? tpw_big : (ntiles_real >= 4096 ? ntiles_real / 4096 : 1));
        if (P.tiles_per_warp < 1) P.tiles_per_warp = 1;
        if (m) {  // the 1D flags count exchange ROUNDS, the arrival counters count send steps
            S.want = c->round + 1;
            S.post = c->round + 2;
            S.epoch = c->sends + 1;
            // tiles owning one of the H outermost real cells of a side (shifted coordinates: real cell j sits at j + ext_lo)
            S.edge_warps_lo = (unsigned)((H + ext_lo - 1) / 120 + 1);
            S.edge_warps_hi = (unsigned)((g.nx + ext_lo - 1) / 120 - (g.nx - H + ext_lo) / 120 + 1);
        }
        P.sync = S;
        P.pdl = use_pdl(c);
        const int warps = (P.ntiles + P.tiles_per_warp - 1) / P.tiles_per_warp;  // a warp marches through consecutive tiles
        dim3 block(128), grid((warps + 3) / 4);
        // first wave starts on its neighbours' flags (only meaningful in a chain of programmatic launches; >= 4 blocks of 480*tpw cells
        // so that a block's cells stay within its own and its neighbours' blocks from one launch to the next)
        // (decided from the slab itself so that it is the same for every launch of the context)
        const int grid_real = ((ntiles_real + P.tiles_per_warp - 1) / P.tiles_per_warp + 3) / 4;
        if (P.pdl && c->done && grid_real >= 6 && env_int("SHLL_EARLY", 1) != 0) {
            P.early_blocks = (int)grid.x - 1 < EARLY_MAX ? (int)grid.x - 1 : EARLY_MAX;
            const int cap = env_int("SHLL_EARLY_BLOCKS", EARLY_MAX);
            if (cap >= 1 && cap < P.early_blocks) P.early_blocks = cap;
            P.early_want = c->early_epoch;
            P.early_post = c->early_epoch + 1;
            P.early_prev_grid = c->early_prev_grid;
            P.done = c->done;
            P.early_err = c->flags + 4;
        }
        e = launch_step1d(c->key, P, grid, block, c->stream, nsub);
        if (e == cudaSuccess && P.early_blocks > 0) { c->early_epoch++; c->early_prev_grid = (int)grid.x; }
        sent = send;
    }
    if (e != cudaSuccess) return fail(c, SHLL_E_CUDA, "step kernel launch failed (%s): %s", c->variant, cudaGetErrorString(e));
    if (early2d) c->early_epoch++;
    c->cur = outb;
    c->state_index += (unsigned)nsub;
    c->epoch++;
    c->launches++;
    if (sent) { c->round++; c->sends++; }
    return SHLL_OK;
}

// Steps the next launch advances: 2 where a two-step kernel exists for this context and at least two steps remain (1D slabs: and the
// launch would not straddle the end of an exchange round).
int next_nsub(const shll_ctx *c, long remaining)
{
    if (remaining < 2) return 1;
    if (c->cfg.dims == 2) return c->fuse2 ? 2 : 1;
    if (!c->fuse1d) return 1;
    if (multi(c)) {
        const int K = c->halo_K;
        const int pos = (int)((c->state_index - c->origin) % (unsigned)K);
        if (pos > K - 2) return 1;
    }
    return 2;
}

// Persistent 1D march: returns SHLL_E_STATE (without touching the error string) when the grid does not fit the scheme.
int run_persistent_1d(shll_ctx *c, long nsteps)
{
    const shll_config &g = c->cfg;
    if (!c->persist_blocks) {
        int dev_sms = 0, coop = 0;
        CK(c, cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, g.device));
        CK(c, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, g.device));
        if (!coop) return SHLL_E_STATE;
        int K = env_int("SHLL_PERSIST_K", 16);  // steps per round: B200 sweep, 1.37 us/step at 16 vs 1.57 at 8 (FAST)
        if (K < 1) K = 1;
        const int h = K * g.order;
        int nb = dev_sms;
        while (nb > 1 && g.nx / nb < 2 * h) nb--;             // every block must own at least its two strips
        const int seg_max = (g.nx + nb - 1) / nb;
        const int threads = ((seg_max + 2 * h + 31) / 32) * 32;
        if (threads > 1024 || g.nx < 2 * h) return SHLL_E_STATE;  // too many cells for one cell per thread: streaming kernel
        // the context becomes "persistent" only once everything it needs exists
        float *strips = nullptr;
        unsigned *round_done = nullptr;
        cudaError_t ea = cudaMalloc(&strips, (size_t)2 * nb * 2 * 3 * h * sizeof(float));
        if (ea == cudaSuccess) ea = cudaMalloc(&round_done, (size_t)nb * sizeof(unsigned));
        if (ea == cudaSuccess) ea = cudaMemsetAsync(round_done, 0, (size_t)nb * sizeof(unsigned), c->stream);
        if (ea != cudaSuccess) {
            if (strips) cudaFree(strips);
            if (round_done) cudaFree(round_done);
            (void)cudaGetLastError();
            c->persist_failed = true;
            return SHLL_E_STATE;  // the per-step kernels need no extra memory: fall through to them
        }
        c->strips = strips; c->round_done = round_done;
        c->persist_blocks = nb; c->persist_threads = threads; c->persist_K = K;
        c->persist_rounds = 0;
    }
    Persist1DParams P;
    memset(&P, 0, sizeof(P));
    const int in = c->cur, outb = c->cur ^ 1;
    for (int k = 0; k < 3; k++) { P.in[k] = c->plane(in, k); P.out[k] = c->plane(outb, k); }
    P.strips = c->strips; P.round_done = c->round_done; P.err = c->flags + 4;
    P.round_base = c->persist_rounds;
    P.n = g.nx; P.nblocks = c->persist_blocks; P.K = c->persist_K; P.hmax = c->persist_K * g.order;
    P.nsteps = nsteps;
    P.dtdx = g.dt_on_dx; P.half_dtdx = 0.5f * g.dt_on_dx; P.alpha = g.alpha;
    P.quarter = 0.25f;
    P.timeout_ns = (unsigned long long)env_int("SHLL_HALO_TIMEOUT_MS", 5000) * 1000000ull;
    cudaError_t e = launch_persist1d(c->key, P, c->persist_blocks, c->persist_threads, c->stream);
    if (e != cudaSuccess) {
        // e.g. cudaErrorCooperativeLaunchTooLarge under MPS / a reduced SM count: nothing has run, so clear the launch error,
        // never try again on this context and let shll_run take the launch-per-step path.
        (void)cudaGetLastError();
        c->persist_failed = true;
        return SHLL_E_STATE;
    }
    const long nrounds = (nsteps + c->persist_K - 1) / c->persist_K;
    c->persist_rounds += (unsigned)(nrounds - 1);
    c->cur = outb;
    c->state_index += (unsigned)nsteps;
    c->epoch += 1;
    c->launches += 1;
    return SHLL_OK;
}

int check_halo_error(shll_ctx *c)
{
    if (!multi(c) && !c->persist_blocks && !c->early_epoch) return SHLL_OK;
    unsigned err = 0;
    CK(c, cudaMemcpyAsync(&err, c->flags + 4, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (err == 2u) return fail(c, SHLL_E_TIMEOUT, "a block waited more than 2 s for its neighbours of the previous step launch (early-start protocol, step1d.cuh)");
    if (err) return fail(c, SHLL_E_TIMEOUT, "rank %d: a neighbour's halo (or a neighbouring block's strip) did not arrive within the timeout", c->cfg.rank);
    return SHLL_OK;
}

}  // namespace

extern "C" {

int shll_upload_u(shll_ctx *c, const float *const u[4])
{
    if (!c || !u) return fail(c, SHLL_E_INVAL, "shll_upload_u: null argument");
    CK(c, cudaSetDevice(c->cfg.device));
    int rc = check_connected(c);
    if (rc) return rc;
    for (int k = 0; k < c->ncomp; k++) {
        if (!u[k]) return fail(c, SHLL_E_INVAL, "shll_upload_u: u[%d] is null", k);
        CK(c, cudaMemcpyAsync(c->plane(c->cur, k), u[k], (size_t)c->ncells * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    c->has_state = true;
    if (multi(c)) {
        // publish our edge rows of the freshly uploaded state to the neighbours' halos
        const shll_config &g = c->cfg;
        const bool lo_wall = (g.rank == 0), hi_wall = (g.rank == g.nranks - 1);
        if (g.dims == 1) {
            // 1D: a new exchange round starts at this state; the outermost H = K*order cells go into the neighbours' mailboxes
            const long H = (long)c->halo_K * g.order;
            c->origin = c->state_index;
            push_halo_kernel<<<1, 128, 0, c->stream>>>(c->plane(c->cur, 0), c->plane(c->cur, 0) + (g.nx - H),
                                                        lo_wall ? nullptr : reinterpret_cast<float *>(c->peer_flags[0] + mail_index(1, c->round, 0)),
                                                        hi_wall ? nullptr : reinterpret_cast<float *>(c->peer_flags[1] + mail_index(0, c->round, 0)),
                                                        H, c->ncomp, c->plane_elems, HALO1D_MAX, HALO1D_MAX,
                                                        lo_wall ? nullptr : c->peer_flags[0] + 1, hi_wall ? nullptr : c->peer_flags[1] + 0,
                                                        c->round + 1);
        } else {
        const long row = g.ny;
        const long depth = c->fuse2 ? 2 : g.order;
        const long count = depth * row;
        const float *src_lo = c->plane(c->cur, 0);
        const float *src_hi = c->plane(c->cur, 0) + ((long)g.nx - depth) * row;
        float *dst_lo = lo_wall ? nullptr : peer_halo_ptr(c, 0, c->cur, 0);
        float *dst_hi = hi_wall ? nullptr : peer_halo_ptr(c, 1, c->cur, 0);
        push_halo_kernel<<<1, 1024, 0, c->stream>>>(src_lo, src_hi, dst_lo, dst_hi, count, c->ncomp, c->plane_elems,
                                                     lo_wall ? 0 : peer_plane_elems(c->peer_desc[0]),
                                                     hi_wall ? 0 : peer_plane_elems(c->peer_desc[1]),
                                                     lo_wall ? nullptr : c->peer_flags[0] + 1,
                                                     hi_wall ? nullptr : c->peer_flags[1] + 0, c->state_index + 1);
        }
        CK(c, cudaGetLastError());
        c->launches++;
    }
    CK(c, cudaStreamSynchronize(c->stream));
    return SHLL_OK;
}

int shll_download_u(shll_ctx *c, float *const u[4])
{
    if (!c || !u) return fail(c, SHLL_E_INVAL, "shll_download_u: null argument");
    if (!c->has_state) return fail(c, SHLL_E_STATE, "shll_download_u before shll_upload_u");
    CK(c, cudaSetDevice(c->cfg.device));
    for (int k = 0; k < c->ncomp; k++) {
        if (!u[k]) return fail(c, SHLL_E_INVAL, "shll_download_u: u[%d] is null", k);
        CK(c, cudaMemcpyAsync(u[k], c->plane(c->cur, k), (size_t)c->ncells * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    CK(c, cudaStreamSynchronize(c->stream));
    return check_halo_error(c);
}

int shll_download_p(shll_ctx *c, float *const p[4], float *a)
{
    if (!c || !p) return fail(c, SHLL_E_INVAL, "shll_download_p: null argument");
    if (!c->has_state) return fail(c, SHLL_E_STATE, "shll_download_p before shll_upload_u");
    CK(c, cudaSetDevice(c->cfg.device));
    if (!c->scratch) CK(c, cudaMalloc(&c->scratch, (size_t)(c->ncomp + 1) * c->ncells * sizeof(float)));
    const float *uin[4] = {nullptr, nullptr, nullptr, nullptr};
    float *pout[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < c->ncomp; k++) {
        uin[k] = c->plane(c->cur, k);
        pout[k] = c->scratch + (size_t)k * c->ncells;
    }
    float *adev = c->scratch + (size_t)c->ncomp * c->ncells;
    CK(c, launch_prim(c->cfg.dims, c->cfg.mode, c->cfg.tform, uin, pout, adev, c->ncells, c->stream));
    c->launches++;
    for (int k = 0; k < c->ncomp; k++) {
        if (!p[k]) return fail(c, SHLL_E_INVAL, "shll_download_p: p[%d] is null", k);
        CK(c, cudaMemcpyAsync(p[k], pout[k], (size_t)c->ncells * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    if (a) CK(c, cudaMemcpyAsync(a, adev, (size_t)c->ncells * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return check_halo_error(c);
}

int shll_run(shll_ctx *c, long nsteps)
{
    if (!c || nsteps < 0) return fail(c, SHLL_E_INVAL, "shll_run: bad argument");
    if (!c->has_state) return fail(c, SHLL_E_STATE, "shll_run before shll_upload_u");
    CK(c, cudaSetDevice(c->cfg.device));
    int rc = check_connected(c);
    if (rc) return rc;
    // Launch-bound regime, 1D: one cooperative launch keeps every cell in a register for all nsteps (persist1d.cuh).
    if (c->cfg.dims == 1 && !multi(c) && nsteps >= 32 && !c->persist_failed && env_int("SHLL_PERSIST", 1) != 0) {
        rc = run_persistent_1d(c, nsteps);
        if (rc != SHLL_E_STATE) return rc;  // SHLL_E_STATE here means "not eligible": fall through to the launch-per-step paths
    }
    // Launch-bound regime (small grids, many steps): replay a CUDA graph of GRAPH_STEPS captured steps instead of
    // issuing every launch from the host.  GRAPH_STEPS is even, so the ping-pong buffers end where they started.
    // Multi-GPU runs are excluded: their per-step halo flags are kernel parameters that change every step.
    constexpr int GRAPH_STEPS = 128;
    const bool want_graph = !multi(c) && c->ncells <= (1L << 22) && nsteps >= 2 * GRAPH_STEPS && env_int("SHLL_GRAPH", 1) != 0;
    if (want_graph && !c->graph && !c->graph_tried) {
        c->graph_tried = true;
        cudaGraph_t g = nullptr;
        const int cur0 = c->cur;
        c->graph_cur = cur0;
        const unsigned si0 = c->state_index, ep0 = c->epoch;
        const long l0 = c->launches;
        if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            int crc = SHLL_OK;
            c->capturing = true;
            const int gsub = next_nsub(c, GRAPH_STEPS);   // (single slab: the same for every launch of the graph)
            for (int s = 0; s < GRAPH_STEPS && crc == SHLL_OK; s += gsub) crc = launch_one_step(c, gsub);
            c->capturing = false;
            cudaError_t e = cudaStreamEndCapture(c->stream, &g);
            if (crc == SHLL_OK && e == cudaSuccess && g) {
                if (cudaGraphInstantiate(&c->graph, g, 0) != cudaSuccess) c->graph = nullptr;
            }
            if (g) cudaGraphDestroy(g);
        }
        (void)cudaGetLastError();
        c->cur = cur0; c->state_index = si0; c->epoch = ep0; c->launches = l0;  // the capture executed nothing
    }
    if (want_graph && c->graph) {
        if (c->cur != c->graph_cur) {  // an odd number of single steps happened before: realign the ping-pong
            rc = launch_one_step(c);
            if (rc) return rc;
            nsteps--;
        }
        while (nsteps >= GRAPH_STEPS) {
            CK(c, cudaGraphLaunch(c->graph, c->stream));
            const int gsub = next_nsub(c, GRAPH_STEPS);
            c->state_index += GRAPH_STEPS; c->epoch += GRAPH_STEPS / gsub; c->launches += GRAPH_STEPS / gsub;
            nsteps -= GRAPH_STEPS;
        }
    }
    while (nsteps > 0) {
        const int nsub = next_nsub(c, nsteps);
        rc = launch_one_step(c, nsub);
        if (rc) return rc;
        nsteps -= nsub;
    }
    return SHLL_OK;
}

int shll_sync(shll_ctx *c)
{
    if (!c) return fail(c, SHLL_E_INVAL, "shll_sync: null context");
    CK(c, cudaSetDevice(c->cfg.device));
    CK(c, cudaStreamSynchronize(c->stream));
    return check_halo_error(c);
}

int shll_run_timed(shll_ctx *c, long nsteps, float *ms)
{
    if (!c || !ms) return fail(c, SHLL_E_INVAL, "shll_run_timed: null argument");
    CK(c, cudaSetDevice(c->cfg.device));
    CK(c, cudaEventRecord(c->ev0, c->stream));
    int rc = shll_run(c, nsteps);
    if (rc) return rc;
    CK(c, cudaEventRecord(c->ev1, c->stream));
    CK(c, cudaEventSynchronize(c->ev1));
    CK(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return check_halo_error(c);
}

int shll_max_cfl(shll_ctx *c, float *cfl)
{
    if (!c || !cfl) return fail(c, SHLL_E_INVAL, "shll_max_cfl: null argument");
    if (!c->has_state) return fail(c, SHLL_E_STATE, "shll_max_cfl before shll_upload_u");
    CK(c, cudaSetDevice(c->cfg.device));
    const float *uin[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < c->ncomp; k++) uin[k] = c->plane(c->cur, k);
    float *out_dev = reinterpret_cast<float *>(c->flags + 8);
    CK(c, cudaMemsetAsync(out_dev, 0, sizeof(float), c->stream));
    // 2D planes are contiguous over the owned rows; 1D over the owned cells.
    CK(c, launch_max_cfl(c->cfg.dims, c->cfg.mode, c->cfg.tform, uin, c->ncells, c->cfg.dt_on_dx, c->cfg.dt_on_dy, out_dev, c->stream));
    c->launches++;
    CK(c, cudaMemcpyAsync(cfl, out_dev, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return SHLL_OK;
}

int shll_conserved_sums(shll_ctx *c, double sums[4])
{
    if (!c || !sums) return fail(c, SHLL_E_INVAL, "shll_conserved_sums: null argument");
    if (!c->has_state) return fail(c, SHLL_E_STATE, "shll_conserved_sums before shll_upload_u");
    CK(c, cudaSetDevice(c->cfg.device));
    const int nb = conserved_sums_blocks();
    if (!c->sums_dev) CK(c, cudaMalloc(&c->sums_dev, (size_t)nb * 4 * sizeof(double)));
    const float *uin[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < c->ncomp; k++) uin[k] = c->plane(c->cur, k);
    CK(c, launch_conserved_sums(c->ncomp, uin, c->ncells, c->sums_dev, c->stream));
    c->launches++;
    std::vector<double> host((size_t)nb * 4);
    CK(c, cudaMemcpyAsync(host.data(), c->sums_dev, host.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 4; k++) {
        double t = 0.0;
        for (int b = 0; b < nb; b++) t += host[(size_t)b * 4 + k];
        sums[k] = (k < c->ncomp) ? t : 0.0;
    }
    return SHLL_OK;
}

int shll_halo_wait_stats(shll_ctx *c, double stats[3])
{
    if (!c || !stats) return fail(c, SHLL_E_INVAL, "shll_halo_wait_stats: null argument");
    CK(c, cudaSetDevice(c->cfg.device));
    unsigned long long raw[3] = {0, 0, 0};
    CK(c, cudaMemcpyAsync(raw, c->flags + WAIT_STATS_OFF, sizeof(raw), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    stats[0] = (double)raw[0] * 1e-9;
    stats[1] = (double)raw[1] * 1e-9;
    stats[2] = (double)raw[2];
    return SHLL_OK;
}

long shll_launch_count(const shll_ctx *c) { return c ? c->launches : 0; }
const char *shll_variant_name(const shll_ctx *c) { return c ? c->variant : ""; }

int shll_peer_export(shll_ctx *c, shll_peer_desc *d)
{
    if (!c || !d) return fail(c, SHLL_E_INVAL, "shll_peer_export: null argument");
    CK(c, cudaSetDevice(c->cfg.device));
    memset(d, 0, sizeof(*d));
    static_assert(sizeof(cudaIpcMemHandle_t) == SHLL_IPC_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CK(c, cudaIpcGetMemHandle(&h, c->state));
    memcpy(d->state_handle, &h, SHLL_IPC_BYTES);
    CK(c, cudaIpcGetMemHandle(&h, c->flags));
    memcpy(d->flag_handle, &h, SHLL_IPC_BYTES);
    d->pid = (int64_t)getpid();
    d->state_ptr = (uint64_t)(uintptr_t)c->state;
    d->flag_ptr = (uint64_t)(uintptr_t)c->flags;
    d->device = c->cfg.device;
    d->nx = c->cfg.nx; d->ny = c->cfg.ny; d->dims = c->cfg.dims; d->order = c->cfg.order;
    return SHLL_OK;
}

int shll_peer_connect(shll_ctx *c, int side, const shll_peer_desc *d)
{
    if (!c || !d || (side != -1 && side != 1)) return fail(c, SHLL_E_INVAL, "shll_peer_connect: bad argument");
    const int s = side < 0 ? 0 : 1;
    if (c->connected[s]) return fail(c, SHLL_E_STATE, "side %d already connected", side);
    if (d->dims != c->cfg.dims || d->ny != c->cfg.ny || d->order != c->cfg.order)
        return fail(c, SHLL_E_INVAL, "neighbour geometry mismatch (dims %d/%d ny %d/%d order %d/%d)", d->dims, c->cfg.dims, d->ny, c->cfg.ny, d->order, c->cfg.order);
    CK(c, cudaSetDevice(c->cfg.device));
    if (d->pid == (int64_t)getpid()) {
        // same process: plain peer pointers
        if (d->device != c->cfg.device) {
            int can = 0;
            CK(c, cudaDeviceCanAccessPeer(&can, c->cfg.device, d->device));
            if (!can) return fail(c, SHLL_E_CUDA, "device %d cannot access peer %d", c->cfg.device, d->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(d->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(c, e);
            (void)cudaGetLastError();
        }
        c->peer_state[s] = reinterpret_cast<float *>((uintptr_t)d->state_ptr);
        c->peer_flags[s] = reinterpret_cast<unsigned *>((uintptr_t)d->flag_ptr);
        c->peer_ipc[s] = false;
    } else {
        cudaIpcMemHandle_t h;
        void *p = nullptr;
        memcpy(&h, d->state_handle, SHLL_IPC_BYTES);
        CK(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_state[s] = static_cast<float *>(p);
        memcpy(&h, d->flag_handle, SHLL_IPC_BYTES);
        CK(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_flags[s] = static_cast<unsigned *>(p);
        c->peer_ipc[s] = true;
    }
    c->peer_desc[s] = *d;
    c->connected[s] = true;
    return SHLL_OK;
}

}  // extern "C"
