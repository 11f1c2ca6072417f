// step1d.cuh -- one fused SHLL time step on a 1D slab (base_shll.c and the derived 1D 2nd-order program).
//
// Layout: three FP32 planes (rho, rho*u, E), each padded by PAD1D floats on both sides so that the first owned
// cell is 16-byte aligned and the halo cells of a slab (2 per side) live at indices -2,-1 and n,n+1.
//
// A warp owns 120 consecutive cells: every lane loads one float4 (4 cells) per plane -> 512 contiguous bytes per
// plane per warp; lanes 0 and 31 are halo lanes that recompute the neighbouring tiles' edge cells, so the +-1 / +-2
// cell stencil is closed with in-thread neighbours plus one shuffle per plane per direction, and no warp ever
// talks to another.  Algorithmic traffic: 3 planes read + 3 written = 24 B per cell per step.
#pragma once
#include <cstdint>

#include "halo_sync.cuh"
#include "shll_math.cuh"

namespace shll {

// Padding on both sides of a 1D plane: HALO1D_MAX halo cells of a slab (temporal blocking of the GPU-to-GPU exchange, see below)
// + the 4 cells of lane 0's float4 + 4 spare.  A multiple of 4 floats, so owned cell 0 stays 16-byte aligned.
constexpr int HALO1D_MAX = 32;
constexpr int PAD1D = HALO1D_MAX + 8;

struct Step1DParams {
    const float *in[3];  // plane base = local cell 0
    float *out[3];
    float *lo_peer[3];   // lower neighbour's cells n_peer.. (its upper halo), or NULL
    float *hi_peer[3];   // upper neighbour's cells -ORDER.. (its lower halo), or NULL
    int n;
    int lo_wall, hi_wall;
    int ntiles;
    float dtdx, half_dtdx, alpha;
    int tiles_per_warp;  // step1d_acc.cuh: consecutive tiles one warp marches through (register prefetch of the next one)
    float quarter;  // 0.25f as a parameter (register operand of the one-LOP3 sign transfer, step1d_acc.cuh)
    int pdl;        // launched with programmatic stream serialization (halo_sync.cuh: pdl_wait_for_previous_step)
    // ---- early start of the first wave (see step1d_ring_march) -----------------------------------------------------------
    int early_blocks;        // blocks [0, early_blocks) start on their neighbours' flags instead of griddepcontrol.wait (0 = off)
    unsigned early_want;     // value the previous step launch published in done[]
    unsigned early_post;     // value this launch publishes
    int early_prev_grid;     // grid size of the launch that published early_want (its blocks [0, early_prev_grid) exist)
    unsigned *done;          // [early_blocks + 2] per-block "my output of step launch #v is visible" flags (monotonic)
    unsigned *early_err;     // error word (shared with the halo protocol)
    // ---- multi-GPU slabs: temporal blocking of the halo exchange -------------------------------------------------------
    // Neighbouring GPUs exchange H = K*ORDER cells once every K steps instead of ORDER cells every step: a round starts with
    // the H halo cells of both sides valid, every step of the round the kernel ALSO updates the halo cells that are still
    // valid (the garbage front moves ORDER cells inward per step and reaches cell 0 exactly when the round ends), and the last
    // step of the round sends the slab's outermost H owned cells to the neighbours' mailboxes.  K-1 of K steps have no
    // cross-GPU dependency at all.  The host shifts in/out/n so that tile 0 starts at the first cell to be updated:
    //   in[k], out[k] = plane + real cell -ext_lo ;  n = n_real + ext_lo + ext_hi   (ext = 0 at a wall and on send steps).
    int recv;            // this step starts a round: edge tiles wait for the neighbours' flags and copy the mailboxes into the halo cells
    int xch;             // > 0: this step ends a round: send the outermost xch (= H) owned cells of each side (ext_lo == ext_hi == 0)
    int hcells;          // H
    int ext_lo, n_real;  // see above
    int interior_end;    // tiles reaching cell index interior_end - ORDER (shifted coordinates) or beyond take the EDGE path
    const float *mail_lo, *mail_hi;  // local mailboxes [3][HALO1D_MAX] the neighbours filled (read on recv steps), NULL at a wall
    HaloSync sync;       // multi-GPU only; lo_peer / hi_peer above point at the NEIGHBOURS' mailboxes (component stride HALO1D_MAX)
};

// recv step, edge tiles only (whole warp): wait for the neighbour, then copy its mailbox into the halo cells of the input
// plane.  Several warps may do this for the same side (every tile that reads a halo cell does): they write identical values.
__device__ __forceinline__ void step1d_recv_halo(const Step1DParams &P, bool lo, bool hi)
{
    const int lane = threadIdx.x & 31;
    if (lo && P.mail_lo != nullptr) {
        halo_wait(P.sync, P.sync.wait_lo, 0);
        if (lane < P.hcells) {
#pragma unroll
            for (int k = 0; k < 3; k++)
                const_cast<float *>(P.in[k])[P.ext_lo - P.hcells + lane] = *(const volatile float *)(P.mail_lo + k * HALO1D_MAX + lane);
        }
    }
    if (hi && P.mail_hi != nullptr) {
        halo_wait(P.sync, P.sync.wait_hi, 1);
        if (lane < P.hcells) {
#pragma unroll
            for (int k = 0; k < 3; k++)
                const_cast<float *>(P.in[k])[P.ext_lo + P.n_real + lane] = *(const volatile float *)(P.mail_hi + k * HALO1D_MAX + lane);
        }
    }
    __syncwarp();  // the tile's own loads of those cells come after every lane's stores
}

// ---- how a warp gets its tiles ------------------------------------------------------------------------------------
// A warp marches through `tiles_per_warp` consecutive tiles.  Their loads go through a per-warp shared-memory ring filled
// by cp.async (LDGSTS, 16 bytes per lane per plane, L1 bypassed): STAGES-1 tiles are in flight per warp without holding
// registers, which is what it takes to cover HBM latency at ~6.5 TB/s with 16-24 resident warps per SM (one tile per warp,
// or a register prefetch of ONE tile ahead, stalled on the long scoreboard for 6 of every 7 issue cycles and stopped at 77 %
// of the HBM roofline in this round's ncu capture of that version; the ring version is profiles/r01_1d_o2_acc_fast.ncu.txt).  Every lane reads back only the 16 bytes it copied itself,
// so no cross-lane synchronisation is needed.  The address is clamped to the last float4 inside the allocation
// [-PAD1D, roundup4(n) + PAD1D) (only the ragged last tile needs it).  EDGE tiles -- the only ones that can read halo cells
// a neighbour GPU has just written -- ignore the ring's copy and load directly after their halo wait.
constexpr int STEP1D_STAGES = 4;

__device__ __forceinline__ void step1d_prefetch(const Step1DParams &P, int tile, int lane, uint32_t slot)
{
    const int jl = min(tile * 120 + (lane - 1) * 4, ((P.n + 3) & ~3));
#pragma unroll
    for (int k = 0; k < 3; k++)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot + k * 512u), "l"(P.in[k] + jl) : "memory");
}

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// tile_fn(tile, lane, interior, cur[3]); blocks are 128 threads (4 warps); ORDER decides which tiles are interior.
//
// Between two step launches.  With programmatic dependent launch the blocks of step s+1 are placed while step s drains, but
// griddepcontrol.wait holds ALL of them until the LAST block of step s has finished and flushed: every step pays the drain
// of its predecessor plus one DRAM latency of ramp-up (4.5 us of a 36 us step at 2^23 cells, profiles/r02_1d_tiles_per_warp.log).
// A block only needs the output of three blocks of the previous step -- the one with its own index and its two neighbours
// (a block's cells never move by more than 36 cells from one launch to the next, step1d.cuh: ext_lo).  So the blocks of the
// first wave [0, early_blocks) skip griddepcontrol.wait and start as soon as those three blocks have published
// done[b] = launch id (st.release.gpu after a block barrier; ld.acquire.gpu on the consumer side): the ramp-up of step s+1
// overlaps the drain of step s.  Later blocks are placed when step s is long over and keep the plain wait.  Write-after-read
// is covered by the same flags: the block that overwrites cells of buffer X has seen its neighbours -- the only readers of
// those cells in the previous launch -- finish.  No deadlock: a launch only starts once every block of its predecessor has
// STARTED (they all call griddepcontrol.launch_dependents first), so the blocks it waits for are resident or done.
template <class F>
__device__ __forceinline__ void step1d_ring_march(const Step1DParams &P, F &&tile_fn, int order)
{
    __shared__ __align__(16) float4 ring[4][STEP1D_STAGES][3][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = blockIdx.x * (blockDim.x >> 5) + wib;
    const int t0 = w * P.tiles_per_warp, t1 = min(t0 + P.tiles_per_warp, P.ntiles);
    const int b = blockIdx.x;
    const bool early = b < P.early_blocks;                 // block-uniform
    if (P.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (early) {
        if (threadIdx.x < 3) {
            const int nb = b - 1 + (int)threadIdx.x;       // b-1, b, b+1: every one of them publishes (nb <= early_blocks)
            if (nb >= 0 && nb < P.early_prev_grid && (int)(ld_acquire_gpu_u32(P.done + nb) - P.early_want) < 0) {
                const unsigned long long tstart = globaltimer_ns();
                unsigned polls = 0;
                while ((int)(ld_acquire_gpu_u32(P.done + nb) - P.early_want) < 0) {
                    __nanosleep(20);
                    if ((++polls & 1023u) == 0 && globaltimer_ns() - tstart > 2000000000ull) {  // never hang the GPU
                        atomicExch(P.early_err, 2u);
                        break;
                    }
                }
            }
        }
        __syncthreads();
    } else if (P.pdl) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    if (t0 < t1) {
        const uint32_t base = (uint32_t)__cvta_generic_to_shared(&ring[wib][0][0][lane]);
        constexpr uint32_t STAGE_BYTES = 3 * 512;
#pragma unroll
        for (int s = 0; s < STEP1D_STAGES - 1; s++) {
            if (t0 + s < t1) step1d_prefetch(P, t0 + s, lane, base + s * STAGE_BYTES);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        int stage = 0;
        for (int tile = t0; tile < t1; tile++) {
            {   // refill the stage whose tile was consumed in the previous iteration
                const int ahead = tile + STEP1D_STAGES - 1;
                int fill = stage + STEP1D_STAGES - 1;
                if (fill >= STEP1D_STAGES) fill -= STEP1D_STAGES;
                if (ahead < t1) step1d_prefetch(P, ahead, lane, base + fill * STAGE_BYTES);
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            asm volatile("cp.async.wait_group %0;" ::"n"(STEP1D_STAGES - 1) : "memory");
            float4 cur[3];
#pragma unroll
            for (int k = 0; k < 3; k++)
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(cur[k].x), "=f"(cur[k].y), "=f"(cur[k].z), "=f"(cur[k].w)
                             : "r"(base + stage * STAGE_BYTES + k * 512u)
                             : "memory");
            // interior tile: all 128 loaded cells and their +-ORDER neighbours are real owned cells of this slab
            const bool interior = (tile > 0) && ((long)tile * 120 + 124 + order <= (long)P.interior_end - order);
            tile_fn(tile, lane, interior, cur);
            if (++stage == STEP1D_STAGES) stage = 0;
        }
    }
    if (b <= P.early_blocks && P.early_blocks > 0) {       // publishers: the first wave and the block after it
        __syncthreads();                                   // every warp of the block has issued its stores
        if (threadIdx.x == 0) st_release_gpu_u32(P.done + b, P.early_post);
    }
}

// One warp tile.  EDGE = false is the interior fast path: no wall selects, no ragged stores, no halo exchange -- the
// warp-uniform dispatch in step1d_kernel sends only the first tile and the tiles touching the upper end through
// EDGE = true, so >99.9 % of the warps of a large tube never execute a boundary instruction.
template <int ORDER, int BC, int LIM, int MODE, int TFORM, bool POW2, bool EDGE>
__device__ __forceinline__ void step1d_tile(const Step1DParams &P, int tile, int lane, const float4 (&in)[3])
{
    constexpr int VEC = 4;
    constexpr int USEFUL = 30 * VEC;
    const unsigned full = 0xffffffffu;
    const int n = P.n;
    const int j0 = tile * USEFUL + (lane - 1) * VEC;  // >= -4: inside the padding
    // last float4 that is still inside the allocation [-PAD1D, roundup4(n) + PAD1D)
    const int jl = EDGE ? min(j0, ((n + 3) & ~3)) : j0;
    const bool lo_wall = P.lo_wall != 0, hi_wall = P.hi_wall != 0;
    // warps owning one of the first / last ORDER cells exchange halos with the neighbour GPUs
    const int own_lo = tile * USEFUL, own_hi = min(own_lo + USEFUL, n);  // owned cells [own_lo, own_hi)
    // send step: warps owning one of the outermost xch cells; recv step: warps whose loads reach a halo cell
    const bool touch_lo = EDGE && (own_lo < P.xch), touch_hi = EDGE && (own_hi > n - P.xch) && (own_lo < n);
    if (EDGE && P.sync.enabled && P.recv)
        step1d_recv_halo(P, tile == 0, (long)tile * USEFUL + 124 > (long)P.ext_lo + P.n_real);

    float u[VEC][3], fp[VEC][3], fm[VEC][3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float4 t = in[k];
        // plain (coherent) load AFTER the halo wait above: halo cells are written by a peer GPU while the kernel is resident
        if (EDGE) t = *reinterpret_cast<const float4 *>(P.in[k] + jl);
        u[0][k] = t.x; u[1][k] = t.y; u[2][k] = t.z; u[3][k] = t.w;
    }
#pragma unroll
    for (int v = 0; v < VEC; v++) cell_flux_1d<MODE, TFORM>(u[v], fp[v], fm[v]);

    bool at_lo[VEC], at_hi[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) {
        at_lo[v] = EDGE && lo_wall && (j0 + v == 0);
        at_hi[v] = EDGE && hi_wall && (j0 + v == n - 1);
    }

    float uo[VEC][3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float fp_from_lo = __shfl_up_sync(full, fp[VEC - 1][k], 1);  // F+ of cell j-1
        float fm_from_hi = __shfl_down_sync(full, fm[0][k], 1);      // F- of cell j+1
        float fpL[VEC], fmR[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            fpL[v] = (v > 0) ? fp[v > 0 ? v - 1 : 0][k] : fp_from_lo;
            fmR[v] = (v < VEC - 1) ? fm[v < VEC - 1 ? v + 1 : 0][k] : fm_from_hi;
        }
        float t1[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            // reflective ends: (-,+,-) on (rho, rho*u, E), base_shll.c:95-97,108-110; outflow: own flux
            float left = fpL[v], right = fmR[v];
            if (EDGE) {
                if (BC == BC_REFLECT) {
                    left = at_lo[v] ? ((k == 1) ? fm[v][k] : -fm[v][k]) : fpL[v];
                    right = at_hi[v] ? ((k == 1) ? fp[v][k] : -fp[v][k]) : fmR[v];
                } else {
                    left = at_lo[v] ? fp[v][k] : fpL[v];
                    right = at_hi[v] ? fm[v][k] : fmR[v];
                }
            }
            t1[v] = apply_first<MODE>(u[v][k], P.dtdx, flux_sum<MODE>(fp[v][k], fm[v][k], right, left));  // base_shll.c:124
        }
        if (ORDER == 2) {
            float fm_from_lo = __shfl_up_sync(full, fm[VEC - 1][k], 1);
            float fp_from_hi = __shfl_down_sync(full, fp[0][k], 1);
            float dfp[VEC], dfm[VEC];
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                float fmL = (v > 0) ? fm[v > 0 ? v - 1 : 0][k] : fm_from_lo;
                float fpR = (v < VEC - 1) ? fp[v < VEC - 1 ? v + 1 : 0][k] : fp_from_hi;
                const bool edge = at_lo[v] || at_hi[v];
                dfp[v] = limited_slope<LIM>(fpL[v], fp[v][k], fpR, P.alpha);
                dfm[v] = limited_slope<LIM>(fmL, fm[v][k], fmR[v], P.alpha);
                if (EDGE) {
                    dfp[v] = edge ? 0.0f : dfp[v];
                    dfm[v] = edge ? 0.0f : dfm[v];
                }
            }
            float dfp_from_lo = __shfl_up_sync(full, dfp[VEC - 1], 1);
            float dfm_from_hi = __shfl_down_sync(full, dfm[0], 1);
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                float ldf = (v > 0) ? dfp[v > 0 ? v - 1 : 0] : dfp_from_lo;
                float rdf = (v < VEC - 1) ? dfm[v < VEC - 1 ? v + 1 : 0] : dfm_from_hi;
                if (EDGE) {
                    ldf = at_lo[v] ? 0.0f : ldf;
                    rdf = at_hi[v] ? 0.0f : rdf;
                }
                t1[v] = apply_second<MODE, POW2>(t1[v], P.half_dtdx, slope_sum(dfp[v], dfm[v], rdf, ldf));
            }
        }
#pragma unroll
        for (int v = 0; v < VEC; v++) uo[v][k] = t1[v];
    }

    if (!EDGE) {  // interior tile: every owner lane stores three full float4
        if (lane != 0 && lane != 31) {
#pragma unroll
            for (int k = 0; k < 3; k++)
                *reinterpret_cast<float4 *>(P.out[k] + j0) = make_float4(uo[0][k], uo[1][k], uo[2][k], uo[3][k]);
        }
        return;
    }
    const bool owner = !(lane == 0 || lane == 31 || j0 >= n);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (!owner) continue;
        if (j0 + VEC <= n) {
            *reinterpret_cast<float4 *>(P.out[k] + j0) = make_float4(uo[0][k], uo[1][k], uo[2][k], uo[3][k]);
        } else {
#pragma unroll
            for (int v = 0; v < VEC; v++)
                if (j0 + v < n) P.out[k][j0 + v] = uo[v][k];
        }
        // halo exchange fused into the step: the outermost xch cells go straight into the neighbour GPU's mailbox
        if (P.lo_peer[k] != nullptr && j0 < P.xch) {
#pragma unroll
            for (int v = 0; v < VEC; v++)
                if (j0 + v < P.xch && j0 + v < n) P.lo_peer[k][j0 + v] = uo[v][k];
        }
        if (P.hi_peer[k] != nullptr && j0 + VEC > n - P.xch) {
#pragma unroll
            for (int v = 0; v < VEC; v++)
                if (j0 + v >= n - P.xch && j0 + v < n) P.hi_peer[k][j0 + v - (n - P.xch)] = uo[v][k];
        }
    }
    if (P.sync.enabled) {
        if (touch_lo) halo_arrive(P.sync, P.sync.cnt_lo, P.sync.edge_warps_lo, P.sync.sig_lo);
        if (touch_hi) halo_arrive(P.sync, P.sync.cnt_hi, P.sync.edge_warps_hi, P.sync.sig_hi);
    }
}

template <int ORDER, int BC, int LIM, int MODE, int TFORM, bool POW2>
__global__ void __launch_bounds__(128) step1d_kernel(const Step1DParams P)
{
    step1d_ring_march(P, [&](int tile, int lane, bool interior, const float4(&cur)[3]) {
        if (interior) step1d_tile<ORDER, BC, LIM, MODE, TFORM, POW2, false>(P, tile, lane, cur);
        else step1d_tile<ORDER, BC, LIM, MODE, TFORM, POW2, true>(P, tile, lane, cur);
    }, ORDER);
}

}  // namespace shll
