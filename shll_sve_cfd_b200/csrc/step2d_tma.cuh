// step2d_tma.cuh -- the TMA-fed variant of the fused 2D step kernel (same arithmetic, same register sliding window
// as step2d.cuh; only the way the state reaches the registers differs).
//
// Why: the LDG kernel keeps one row per warp in flight (~10 KB per SM), far below the ~35 KB per SM that Little's law
// asks for at HBM3e bandwidth x latency, and spends ~10 issue slots per row on address arithmetic.  Here every block is
// ONE warp with a private shared-memory ring of STAGES boxes; each box is R rows x 32*VEC columns x 4 planes, fetched by
// ONE `cp.async.bulk.tensor.3d` (TMA) copy issued by one lane and tracked by an mbarrier.  The warp reads its cells
// back with conflict-free LDS at constant offsets, so loads cost no address arithmetic, (STAGES-1)*R rows per warp are
// in flight, and out-of-domain halo lanes are zero-filled by the TMA unit (plant_y_ghosts then gives them a ghost / benign state).
// With one warp per block the tile / chunk / row bookkeeping is warp-uniform by construction and runs on the uniform
// datapath (no convergence barriers around the row predicates).
//
// R equals the unroll factor of the row loop (3 for order 1, 4 for order 2), which makes the row-within-box index a
// compile-time constant of each unrolled step.
//
// TMA alignment: global strides AND the first column of a box must be multiples of 16 bytes (a box starting at column -1
// traps with "illegal instruction" -- measured with tools/tma_selftest.cu).  The warp tile starts at column
// tile*USEFUL - HL*VEC, which is only 2- or 1-column aligned, so the box is 4 columns wider than the warp (32*VEC + 4),
// starts at that column rounded down to a multiple of 4, and each lane adds the 0..3 column remainder to its LDS address.
// Requires ny % 4 == 0; other shapes use the LDG kernel.
#pragma once
#include <cuda.h>

#include "step2d.cuh"

namespace shll {

struct Step2DTmaParams {
    Step2DParams base;
    CUtensorMap tmap;     // INPUT buffer as a 3D tensor {ny, nx+4, 4 planes}, origin = halo row -2 of plane 0
    CUtensorMap tmap_out; // step2d_acc.cuh: OUTPUT buffer, same tensor, box = 60 owned columns x 4 rows x 4 planes (TMA store)
    int tma_store;        // step2d_acc.cuh: interior boxes leave through one cp.async.bulk.tensor store per box
    // early start of the first wave (same idea as step1d.cuh: step1d_ring_march): every block publishes done[chunk*ntiles + tile]
    // = launch id when its rows are visible; the first early_blocks blocks of the NEXT launch start on the flags of the five
    // blocks they depend on (own, tile +-1, chunk +-1 -- the stencil is a cross, box corners are loaded but never used)
    // instead of griddepcontrol.wait, so the ramp-up of a step overlaps the drain of its predecessor.
    int early_blocks;
    unsigned early_want, early_post;
    unsigned *done;       // [nchunks * ntiles]
    unsigned *early_err;
    int early_diag;       // also wait for the four diagonal neighbours (two-step launches: the footprint of two cross stencils is a diamond)
    const CUtensorMap *tmap_global;  // optional copy of the same descriptor in device memory (debug switch SHLL_TMAP_GLOBAL)
    int stages;
    int pdl;              // launched with programmatic stream serialization: the kernel waits for its predecessor itself (halo_sync.cuh)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SHLL_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SHLL_DONE_%=;\n"
        "bra SHLL_WAIT_%=;\n"
        "SHLL_DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int x, int y, int z, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}

__device__ __forceinline__ unsigned ld_acquire_gpu2d(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Start of a 2D step block (one warp): let the successor launch be placed, then wait for the input -- the whole previous
// launch (griddepcontrol.wait), or, for the first wave, just the five blocks of it this block reads from / overwrites after.
__device__ __forceinline__ void step2d_wait_for_input(const Step2DTmaParams &T, int tile, int chunk, int lane)
{
    if (T.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // (launch order: chunk 0, the LAST chunk, 1, 2, ... -- the last chunk's lower neighbour finishes at the very end of a launch,
    // so its blocks keep the plain wait)
    const bool last_chunk_slot = T.base.nchunks > 2 && (int)blockIdx.x >= T.base.ntiles && (int)blockIdx.x < 2 * T.base.ntiles;
    if ((int)blockIdx.x < T.early_blocks && !last_chunk_slot) {
        const Step2DParams &P = T.base;
        if (lane < (T.early_diag ? 9 : 5)) {
            // lane 0: own block; 1, 2: tile -+ 1; 3, 4: chunk -+ 1; 5..8: the diagonals
            const int dt = (lane == 1 || lane == 5 || lane == 7) ? -1 : ((lane == 2 || lane == 6 || lane == 8) ? 1 : 0);
            const int dc = (lane == 3 || lane == 5 || lane == 6) ? -1 : ((lane == 4 || lane == 7 || lane == 8) ? 1 : 0);
            const int t = tile + dt, c = chunk + dc;
            if (t >= 0 && t < P.ntiles && c >= 0 && c < P.nchunks) {
                const unsigned *f = T.done + (size_t)c * P.ntiles + t;
                if ((int)(ld_acquire_gpu2d(f) - T.early_want) < 0) {
                    const unsigned long long tstart = globaltimer_ns();
                    unsigned polls = 0;
                    while ((int)(ld_acquire_gpu2d(f) - T.early_want) < 0) {
                        __nanosleep(20);
                        if ((++polls & 1023u) == 0 && globaltimer_ns() - tstart > 2000000000ull) {  // never hang the GPU
                            atomicExch(T.early_err, 2u);
                            break;
                        }
                    }
                }
            }
            asm volatile("fence.proxy.async;" ::: "memory");  // the rows are read through the async proxy (TMA)
        }
        __syncwarp();
    } else if (T.pdl) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
}
// End of a 2D step block: every lane's stores (and lane 0's TMA stores, already waited for) are issued.
__device__ __forceinline__ void step2d_publish_output(const Step2DTmaParams &T, int tile, int chunk, int lane)
{
    // publishers: every block an early block of the next launch may depend on (its own launch slot, tile +-1, chunk +-1).  In launch
    // order (chunk 0, the last chunk, 1, 2, ...) chunk 0's upper neighbour sits TWO chunk slots further, every other one ONE.
    if (T.done != nullptr && (int)blockIdx.x < T.early_blocks + 2 * T.base.ntiles + 2) {
        __syncwarp();
        if (lane == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(T.done + (size_t)chunk * T.base.ntiles + tile), "r"(T.early_post) : "memory");
        }
    }
}

// TMA store: shared -> global, tracked by the issuing thread's bulk async-groups.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, int x, int y, int z, uint32_t src)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(x), "r"(y), "r"(z),
                 "r"(src)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Row source + store sink of one warp for the TMA kernel.
template <int VEC, int R>
struct TmaCtx {
    const Step2DParams *P;
    const Step2DTmaParams *T;
    int ny, r0, r1, j0;
    int wall_lo_row, wall_hi_row, first_real_row, last_real_row, peer_lo_end, peer_hi_begin;
    bool owner;
    YEdge<VEC> Y;
    // ring
    uint32_t ring;       // shared address of stage 0
    uint32_t bars;       // shared address of mbarrier 0
    uint32_t lane_off;   // byte offset of this lane's first cell inside a box row
    int x0;              // first column of the warp's box: multiple of 4 (may be negative: zero-filled)
    int ybase;           // tensor row of box 0 = rbeg + 2
    int nboxes;          // boxes this warp consumes
    int stages;
    static constexpr uint32_t BOX_COLS = 32 * VEC + 4;
    static constexpr uint32_t ROW_BYTES = BOX_COLS * 4;
    static constexpr uint32_t PLANE_BYTES = R * ROW_BYTES;
    static constexpr uint32_t STAGE_BYTES = 4 * PLANE_BYTES;                 // bytes one box delivers
    static constexpr uint32_t STAGE_STRIDE = (STAGE_BYTES + 127u) & ~127u;   // TMA destinations are 128-byte aligned

    __device__ __forceinline__ void arm(int box, int stage) const
    {   // one lane: expect the bytes of the box, then launch ONE 3D copy (32*VEC columns x R rows x 4 planes)
        const uint32_t bar = bars + 8u * stage;
        mbar_expect_tx(bar, STAGE_BYTES);
        tma_load_3d(ring + STAGE_STRIDE * stage, T->tmap_global ? T->tmap_global : &T->tmap, x0, ybase + box * R, 0, bar);
    }
    // read row `within` of the box sitting in `stage`
    template <int WITHIN>
    __device__ __forceinline__ void read_row(int stage, float (&u)[VEC][4]) const
    {
        const uint32_t a = ring + STAGE_STRIDE * stage + WITHIN * ROW_BYTES + lane_off;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (VEC == 1) {
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(u[0][k]) : "r"(a + k * PLANE_BYTES));
            } else if (VEC == 2) {
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(u[0][k]), "=f"(u[VEC > 1 ? 1 : 0][k]) : "r"(a + k * PLANE_BYTES));
            } else {
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(u[0][k]), "=f"(u[VEC > 1 ? 1 : 0][k]), "=f"(u[VEC > 2 ? 2 : 0][k]), "=f"(u[VEC > 3 ? 3 : 0][k])
                             : "r"(a + k * PLANE_BYTES));
            }
        }
    }
    template <int ORDER>
    __device__ __forceinline__ void store_row(int i, const float (&u)[VEC][4]) const
    {
        if (!owner) return;
        const int idx = i * ny + j0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float t[VEC];
#pragma unroll
            for (int v = 0; v < VEC; v++) t[v] = u[v][k];
            VecIO<VEC>::store(P->out[k] + idx, t);
        }
        if (i < peer_lo_end || i >= peer_hi_begin) {  // halo exchange fused into the step (multi-GPU edge rows)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float t[VEC];
#pragma unroll
                for (int v = 0; v < VEC; v++) t[v] = u[v][k];
                if (i < peer_lo_end) VecIO<VEC>::store(P->lo_peer[k] + idx, t);
                else VecIO<VEC>::store(P->hi_peer[k] + ((i - peer_hi_begin) * ny + j0), t);
            }
        }
    }
};

template <int ORDER, int BC, int LIM, int MODE, int VEC, bool POW2>
__global__ void __launch_bounds__(32) step2d_tma_kernel(const __grid_constant__ Step2DTmaParams T)
{
    constexpr int R = (ORDER == 1) ? 3 : 4;      // rows per box == unroll factor of the row loop
    constexpr int HL = (ORDER + VEC - 1) / VEC;  // halo lanes per side of the warp tile
    constexpr int USEFUL = (32 - 2 * HL) * VEC;  // columns a warp owns
    extern __shared__ __align__(128) unsigned char smem[];
    const Step2DParams &P = T.base;
    const int lane = threadIdx.x;
    const int gw = blockIdx.x;  // one warp per block: everything derived from it is warp-uniform
    const int tile = gw % P.ntiles;
    int chunk = gw / P.ntiles;
    if (P.nchunks > 2) chunk = (chunk == 0) ? 0 : (chunk == 1 ? P.nchunks - 1 : chunk - 1);  // edge chunks first

    TmaCtx<VEC, R> X;
    X.P = &P;
    X.T = &T;
    const int nx = P.nx;
    X.ny = P.ny;
    const int xs = tile * USEFUL - HL * VEC;  // first column of the warp tile (halo lanes included)
    X.x0 = xs & ~3;                           // box start, 16-byte aligned (floor, also for negative xs)
    X.j0 = xs + lane * VEC;
    X.owner = (lane >= HL) && (lane < 32 - HL) && (X.j0 < X.ny);
    X.Y.tile_has_wall = (tile == 0) || (tile == P.ntiles - 1);
    X.Y.ghost_lo = (X.j0 + VEC - 1 == -1);
    X.Y.ghost_hi = (X.j0 == X.ny);
#pragma unroll
    for (int v = 0; v < VEC; v++) {
        X.Y.y_inner[v] = (X.j0 + v > 0 && X.j0 + v < X.ny - 1);
        X.Y.outside[v] = (X.j0 + v < 0 || X.j0 + v >= X.ny);
    }
    X.r0 = (int)(((long)chunk * nx) / P.nchunks);
    X.r1 = (int)(((long)(chunk + 1) * nx) / P.nchunks);
    const int never = -(1 << 30);
    X.wall_lo_row = P.lo_wall ? 0 : never;
    X.wall_hi_row = P.hi_wall ? nx - 1 : never;
    X.first_real_row = P.lo_wall ? 0 : never;
    X.last_real_row = P.hi_wall ? nx - 1 : -never;
    X.peer_lo_end = (P.sync.enabled && P.lo_peer[0] != nullptr) ? ORDER : 0;
    X.peer_hi_begin = (P.sync.enabled && P.hi_peer[0] != nullptr) ? nx - ORDER : 0x7fffffff;
    const int rmax = P.hi_wall ? nx - 1 : nx + 1;  // last row that exists in memory
    const int rmin = P.lo_wall ? 0 : -2;
    const bool touch_lo = (X.r0 < ORDER), touch_hi = (X.r1 > nx - ORDER);
    if (P.sync.enabled) {  // the neighbour GPUs' edge rows of the previous step must sit in our halo rows
        if (touch_lo) halo_wait(P.sync, P.sync.wait_lo, 0);
        if (touch_hi) halo_wait(P.sync, P.sync.wait_hi, 1);
    }

    // ---- ring set-up
    const int rbeg = X.r0 - ORDER;  // first row the warp consumes (may not exist at a wall: its box is still fetched)
    const int rlast = X.r1 - 1 + ORDER;
    X.stages = T.stages;
    X.ring = smem_u32(smem);
    X.bars = X.ring + TmaCtx<VEC, R>::STAGE_STRIDE * X.stages;
    X.lane_off = (uint32_t)(xs - X.x0 + lane * VEC) * 4u;
    X.ybase = rbeg + 2;
    X.nboxes = (rlast - rbeg) / R + 1;
    if (lane == 0) {
        for (int s = 0; s < X.stages; s++) mbar_init(X.bars + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    step2d_wait_for_input(T, tile, chunk, lane);
    if (lane == 0) {
        for (int b = 0; b < X.stages && b < X.nboxes; b++) X.arm(b, b);
    }
    __syncwarp();

    int stage = 0;        // stage holding the box currently being consumed
    uint32_t parity = 0;  // its mbarrier phase
    int box = 0;
    auto next_box = [&]() {  // the box in `stage` has been fully read: refill it, move on
        __syncwarp();
        if (lane == 0 && box + X.stages < X.nboxes) X.arm(box + X.stages, stage);
        box++;
        stage++;
        if (stage == X.stages) { stage = 0; parity ^= 1u; }
    };

    if (ORDER == 1) {
        // rows: n = r - rbeg;  box = n / 3, within = n % 3.  Prologue reads n = 0 (row r0-1) and n = 1 (row r0);
        // the step for row i = r0 + 3m + p reads row i+1, i.e. n = 3m + p + 2: within = 2, 0, 1 for p = 0, 1, 2.
        RowSlot<VEC> A, B, C;
#pragma unroll
        for (int v = 0; v < VEC; v++)
#pragma unroll
            for (int k = 0; k < 4; k++) A.fp[v][k] = C.fm[v][k] = 0.0f;
        mbar_wait(X.bars + 8u * stage, parity);
        X.template read_row<0>(stage, A.u);
        X.template read_row<1>(stage, B.u);
        if (rbeg >= rmin) row_compute<1, BC, LIM, MODE, VEC>(A, X.Y, P.alpha);
        row_compute<1, BC, LIM, MODE, VEC>(B, X.Y, P.alpha);
        if (X.r0 == X.wall_lo_row) ghost_below<BC, VEC>(A, B);
        for (int i = X.r0; i < X.r1; i += 3) {
            // p = 0: row i+1 is the last row of the current box
            X.template read_row<2>(stage, C.u);
            if (i + 1 <= rmax) row_compute<1, BC, LIM, MODE, VEC>(C, X.Y, P.alpha);
            else ghost_above<BC, VEC>(C, B);
            finish_o1<BC, MODE, VEC>(X, i, A, B, C);
            next_box();
            if (i + 1 >= X.r1) break;
            // p = 1: first row of the next box
            mbar_wait(X.bars + 8u * stage, parity);
            X.template read_row<0>(stage, A.u);
            if (i + 2 <= rmax) row_compute<1, BC, LIM, MODE, VEC>(A, X.Y, P.alpha);
            else ghost_above<BC, VEC>(A, C);
            finish_o1<BC, MODE, VEC>(X, i + 1, B, C, A);
            if (i + 2 >= X.r1) break;
            // p = 2
            X.template read_row<1>(stage, B.u);
            if (i + 3 <= rmax) row_compute<1, BC, LIM, MODE, VEC>(B, X.Y, P.alpha);
            else ghost_above<BC, VEC>(B, A);
            finish_o1<BC, MODE, VEC>(X, i + 2, C, A, B);
        }
    } else {
        // rows: the step at unrolled position p of iteration m handles row r = rbeg + 4m + p: box m, within p.
        RowSlot<VEC> A, B, C, D;
#pragma unroll
        for (int v = 0; v < VEC; v++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                A.fp[v][k] = A.dfp[v][k] = 0.0f;
                B.u[v][k] = B.fp[v][k] = B.fm[v][k] = B.dfp[v][k] = B.dfm[v][k] = B.s1[v][k] = B.s2[v][k] = 0.0f;
                C.u[v][k] = C.fp[v][k] = C.fm[v][k] = C.s1[v][k] = C.s2[v][k] = C.dfp[v][k] = C.dfm[v][k] = 0.0f;
                D.fp[v][k] = D.fm[v][k] = D.s1[v][k] = D.s2[v][k] = 0.0f;
            }
        for (int r = rbeg; r <= rlast; r += 4) {
            mbar_wait(X.bars + 8u * stage, parity);
            X.template read_row<0>(stage, D.u);
            if (r >= rmin && r <= rmax) row_compute<2, BC, LIM, MODE, VEC>(D, X.Y, P.alpha);
            finish_o2<BC, LIM, MODE, VEC, POW2>(X, r, A, B, C, D);
            if (r + 1 > rlast) break;
            X.template read_row<1>(stage, A.u);
            if (r + 1 >= rmin && r + 1 <= rmax) row_compute<2, BC, LIM, MODE, VEC>(A, X.Y, P.alpha);
            finish_o2<BC, LIM, MODE, VEC, POW2>(X, r + 1, B, C, D, A);
            if (r + 2 > rlast) break;
            X.template read_row<2>(stage, B.u);
            if (r + 2 >= rmin && r + 2 <= rmax) row_compute<2, BC, LIM, MODE, VEC>(B, X.Y, P.alpha);
            finish_o2<BC, LIM, MODE, VEC, POW2>(X, r + 2, C, D, A, B);
            if (r + 3 > rlast) break;
            X.template read_row<3>(stage, C.u);
            if (r + 3 >= rmin && r + 3 <= rmax) row_compute<2, BC, LIM, MODE, VEC>(C, X.Y, P.alpha);
            finish_o2<BC, LIM, MODE, VEC, POW2>(X, r + 3, D, A, B, C);
            next_box();
        }
    }
    step2d_publish_output(T, tile, chunk, lane);
    if (P.sync.enabled) {  // publish: our edge rows of this step have landed in the neighbours' halo rows
        if (touch_lo) halo_arrive(P.sync, P.sync.cnt_lo, P.sync.edge_warps_lo, P.sync.sig_lo);
        if (touch_hi) halo_arrive(P.sync, P.sync.cnt_hi, P.sync.edge_warps_hi, P.sync.sig_hi);
    }
}

}  // namespace shll
