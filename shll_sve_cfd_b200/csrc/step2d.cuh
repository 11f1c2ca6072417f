// step2d.cuh -- one fused SHLL time step on a 2D slab (replaces Compute_F_from_P + Update_U_from_F +
// Compute_P_from_U of base_shll_2d.c / 2nd_order_base_shll.c for one step).
//
// Data layout: SoA, one FP32 plane per conserved component, index = i*ny + j (i = x, slow axis), exactly the
// reference's layout (base_shll_2d.c:157).  Each plane carries 2 halo rows below row 0 and above row nx-1.
//
// Work decomposition -- no shared memory, no __syncthreads, every warp is independent:
//   * a warp owns a column tile of 32*VEC consecutive j and a chunk of `rows_per_chunk` consecutive i;
//   * it marches along i keeping the x-direction stencil (F+ of rows behind, F- of rows ahead, slopes) in a
//     register sliding window, so each cell's primitives and split fluxes are evaluated once per step
//     (plus ORDER halo rows per chunk end);
//   * the y-direction neighbours (H+ from j-1, H- from j+1, limited slopes) come from the adjacent lanes by
//     warp shuffle; the outermost HL lanes of the warp are halo lanes that recompute the neighbouring tile's
//     edge cells, so tiles overlap by 2*HL*VEC columns and nothing crosses a warp;
//   * lane -> VEC consecutive j, so a warp reads 128*VEC contiguous bytes per plane per row (float/float2/float4 loads);
//   * the next row's state is prefetched into registers one iteration ahead.
// Algorithmic traffic: 4 planes read + 4 planes written = 32 B per cell per step.
#pragma once
#include "halo_sync.cuh"
#include "shll_math.cuh"

namespace shll {

struct Step2DParams {
    const float *in[4];  // plane base = local row 0, column 0
    float *out[4];
    float *lo_peer[4];   // lower neighbour's upper-halo rows (its row nx_peer), or NULL
    float *hi_peer[4];   // upper neighbour's lower-halo rows (its row -ORDER), or NULL
    int nx, ny;
    int lo_wall, hi_wall;  // local row 0 / nx-1 is a physical wall of the global domain
    int ntiles, nchunks;   // column tiles x row chunks = warps of the launch; chunks are balanced (sizes differ by <= 1)
    float dtdx, dtdy, half_dtdx, half_dtdy, alpha;
    HaloSync sync;         // multi-GPU only
};

// Plain (coherent) loads, not ld.global.nc: halo rows are written by a peer GPU while the kernel may be resident.
template <int VEC>
struct VecIO;
template <>
struct VecIO<1> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[1]) { v[0] = *p; }
    static __device__ __forceinline__ void store(float *p, const float (&v)[1]) { *p = v[0]; }
};
template <>
struct VecIO<2> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[2])
    {
        float2 t = *reinterpret_cast<const float2 *>(p);
        v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[2])
    {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    }
};
template <>
struct VecIO<4> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[4])
    {
        float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[4])
    {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};

// Ghost flux at a physical wall.  REFLECT: own opposite split flux, sign + for the wall-normal momentum
// component and - otherwise (base_shll_2d.c:152-155,186-189).  OUTFLOW: own same split flux
// (2nd_order_base_shll.c:216-219).  `same` / `opposite` are the cell's own split fluxes.
template <int BC>
__device__ __forceinline__ float wall_flux(float same, float opposite, int k, int normal_comp)
{
    if (BC == BC_REFLECT) return (k == normal_comp) ? opposite : -opposite;
    return same;
}

template <int VEC>
struct Row2D {
    float u[VEC][4];
    float fp[VEC][4], fm[VEC][4];  // x split fluxes F+, F-
    float sy1[VEC][4];             // hp - hm + Top - Bottom
};

// Fluxes of one row of cells held by the warp + everything the y direction contributes to their update.
template <int ORDER, int BC, int LIM, int MODE, int VEC>
__device__ __forceinline__ void row_compute(const float (&u)[VEC][4], const bool (&at_lo)[VEC], const bool (&at_hi)[VEC],
                                            float alpha, float (&fp)[VEC][4], float (&fm)[VEC][4],
                                            float (&sy1)[VEC][4], float (&sy2)[VEC][4])
{
    const unsigned full = 0xffffffffu;
    float hp[VEC][4], hm[VEC][4];
#pragma unroll
    for (int v = 0; v < VEC; v++) cell_flux_2d<MODE>(u[v], fp[v], fm[v], hp[v], hm[v]);

#pragma unroll
    for (int k = 0; k < 4; k++) {
        // neighbours across the thread boundary
        float hp_from_lo = __shfl_up_sync(full, hp[VEC - 1][k], 1);   // H+ of cell j-1
        float hm_from_hi = __shfl_down_sync(full, hm[0][k], 1);       // H- of cell j+1
        float hpL[VEC], hmR[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            hpL[v] = (v > 0) ? hp[v > 0 ? v - 1 : 0][k] : hp_from_lo;
            hmR[v] = (v < VEC - 1) ? hm[v < VEC - 1 ? v + 1 : 0][k] : hm_from_hi;
        }
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            float bottom = at_lo[v] ? wall_flux<BC>(hp[v][k], hm[v][k], k, 2) : hpL[v];
            float top = at_hi[v] ? wall_flux<BC>(hm[v][k], hp[v][k], k, 2) : hmR[v];
            sy1[v][k] = flux_sum<MODE>(hp[v][k], hm[v][k], top, bottom);
        }
        if (ORDER == 2) {
            float hm_from_lo = __shfl_up_sync(full, hm[VEC - 1][k], 1);  // H- of cell j-1
            float hp_from_hi = __shfl_down_sync(full, hp[0][k], 1);      // H+ of cell j+1
            float dhp[VEC], dhm[VEC];
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                float hmL = (v > 0) ? hm[v > 0 ? v - 1 : 0][k] : hm_from_lo;
                float hpR = (v < VEC - 1) ? hp[v < VEC - 1 ? v + 1 : 0][k] : hp_from_hi;
                bool edge = at_lo[v] || at_hi[v];  // 2nd_order_base_shll.c:292-300,314-322: first order in wall cells
                dhp[v] = edge ? 0.0f : limited_slope<LIM>(hpL[v], hp[v][k], hpR, alpha);
                dhm[v] = edge ? 0.0f : limited_slope<LIM>(hmL, hm[v][k], hmR[v], alpha);
            }
            float dhp_from_lo = __shfl_up_sync(full, dhp[VEC - 1], 1);
            float dhm_from_hi = __shfl_down_sync(full, dhm[0], 1);
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                float bdf = (v > 0) ? dhp[v > 0 ? v - 1 : 0] : dhp_from_lo;        // Bottom_df = dhp[j-1]
                float tdf = (v < VEC - 1) ? dhm[v < VEC - 1 ? v + 1 : 0] : dhm_from_hi;  // Top_df = dhm[j+1]
                bdf = at_lo[v] ? 0.0f : bdf;  // 2nd_order_base_shll.c:396-399
                tdf = at_hi[v] ? 0.0f : tdf;  // :412-415
                sy2[v][k] = slope_sum(dhp[v], dhm[v], tdf, bdf);
            }
        }
    }
}

template <int ORDER, int BC, int LIM, int MODE, int VEC, bool POW2>
__global__ void __launch_bounds__(128) step2d_kernel(const Step2DParams P)
{
    constexpr int HL = (ORDER + VEC - 1) / VEC;  // halo lanes per side of the warp tile
    constexpr int USEFUL = (32 - 2 * HL) * VEC;  // columns a warp owns
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= P.ntiles * P.nchunks) return;  // whole warp exits together
    const int tile = gw % P.ntiles;
    // chunk order: both edge chunks first (they feed the neighbour GPUs), then the interior
    int chunk = gw / P.ntiles;
    if (P.nchunks > 2) chunk = (chunk == 0) ? 0 : (chunk == 1 ? P.nchunks - 1 : chunk - 1);
    const int nx = P.nx, ny = P.ny;
    const int j0 = tile * USEFUL + (lane - HL) * VEC;
    const int jl = min(max(j0, 0), ny - VEC);  // clamped load column (lanes outside the domain hold unused values)
    const bool owner = (lane >= HL) && (lane < 32 - HL) && (j0 < ny);
    bool at_lo[VEC], at_hi[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) {
        at_lo[v] = (j0 + v == 0);
        at_hi[v] = (j0 + v == ny - 1);
    }
    const int r0 = (int)(((long)chunk * nx) / P.nchunks);
    const int r1 = (int)(((long)(chunk + 1) * nx) / P.nchunks);
    const bool touch_lo = (r0 < ORDER), touch_hi = (r1 > nx - ORDER);
    if (P.sync.enabled) {  // wait until the neighbour GPUs' edge rows of the previous step sit in our halo rows
        if (touch_lo) halo_wait(P.sync, P.sync.wait_lo);
        if (touch_hi) halo_wait(P.sync, P.sync.wait_hi);
    }
    const bool lo_wall = P.lo_wall != 0, hi_wall = P.hi_wall != 0;
    const float alpha = P.alpha;

    auto row_exists = [&](int r) { return (r >= 0 || !lo_wall) && (r < nx || !hi_wall); };
    auto load_row = [&](int r, float (&u)[VEC][4]) {
        const long off = (long)r * ny + jl;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float t[VEC];
            VecIO<VEC>::load(P.in[k] + off, t);
#pragma unroll
            for (int v = 0; v < VEC; v++) u[v][k] = t[v];
        }
    };
    auto store_row = [&](int i, const float (&u)[VEC][4]) {
        if (!owner) return;
        const long off = (long)i * ny + j0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float t[VEC];
#pragma unroll
            for (int v = 0; v < VEC; v++) t[v] = u[v][k];
            VecIO<VEC>::store(P.out[k] + off, t);
            // halo exchange fused into the step: edge rows are also stored straight into the neighbour GPU's halo rows
            if (P.lo_peer[k] != nullptr && i < ORDER) VecIO<VEC>::store(P.lo_peer[k] + (long)i * ny + j0, t);
            if (P.hi_peer[k] != nullptr && i >= nx - ORDER) VecIO<VEC>::store(P.hi_peer[k] + (long)(i - (nx - ORDER)) * ny + j0, t);
        }
    };

    if (ORDER == 1) {
        // window: fpL = F+ of row i-1 ; C = row i ; N = row i+1
        float fpL[VEC][4];
        float uC[VEC][4], fpC[VEC][4], fmC[VEC][4], syC[VEC][4], dummy[VEC][4];
        float uN[VEC][4] = {};
#pragma unroll
        for (int v = 0; v < VEC; v++)
#pragma unroll
            for (int k = 0; k < 4; k++) fpL[v][k] = 0.0f;
        if (row_exists(r0 - 1)) {
            float uL[VEC][4], fmL[VEC][4], syL[VEC][4];
            load_row(r0 - 1, uL);
            row_compute<1, BC, LIM, MODE, VEC>(uL, at_lo, at_hi, alpha, fpL, fmL, syL, dummy);
        }
        load_row(r0, uC);
        if (row_exists(r0 + 1)) load_row(r0 + 1, uN);
        row_compute<1, BC, LIM, MODE, VEC>(uC, at_lo, at_hi, alpha, fpC, fmC, syC, dummy);

        for (int i = r0; i < r1; i++) {
            float uNN[VEC][4] = {};
            float fpN[VEC][4] = {}, fmN[VEC][4] = {}, syN[VEC][4] = {};
            const bool have_next = row_exists(i + 1);
            if (i + 2 <= r1 && row_exists(i + 2)) load_row(i + 2, uNN);  // prefetch one row ahead
            if (have_next) row_compute<1, BC, LIM, MODE, VEC>(uN, at_lo, at_hi, alpha, fpN, fmN, syN, dummy);
            const bool lo = (i == 0) && lo_wall, hi = (i == nx - 1) && hi_wall;
            float uo[VEC][4];
#pragma unroll
            for (int v = 0; v < VEC; v++) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    float left = lo ? wall_flux<BC>(fpC[v][k], fmC[v][k], k, 1) : fpL[v][k];
                    float right = hi ? wall_flux<BC>(fmC[v][k], fpC[v][k], k, 1) : fmN[v][k];
                    float s = flux_sum<MODE>(fpC[v][k], fmC[v][k], right, left);
                    float t = apply_first<MODE>(uC[v][k], P.dtdx, s);      // base_shll_2d.c:227-230
                    uo[v][k] = apply_first<MODE>(t, P.dtdy, syC[v][k]);    // base_shll_2d.c:232-235
                }
            }
            store_row(i, uo);
#pragma unroll
            for (int v = 0; v < VEC; v++) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    fpL[v][k] = fpC[v][k];
                    fpC[v][k] = fpN[v][k];
                    fmC[v][k] = fmN[v][k];
                    syC[v][k] = syN[v][k];
                    uC[v][k] = uN[v][k];
                    uN[v][k] = uNN[v][k];
                }
            }
        }
    } else {
        // window when row r arrives: A = r-3 (F+ and its slope only), B = r-2 (the row being updated),
        // C = r-1 (slopes computed now), N = r.
        float fpA[VEC][4], dfpA[VEC][4];
        float uB[VEC][4], fpB[VEC][4], fmB[VEC][4], dfpB[VEC][4], dfmB[VEC][4], s1B[VEC][4], s2B[VEC][4];
        float uC[VEC][4], fpC[VEC][4], fmC[VEC][4], s1C[VEC][4], s2C[VEC][4];
        float uN[VEC][4];
#pragma unroll
        for (int v = 0; v < VEC; v++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                fpA[v][k] = dfpA[v][k] = 0.0f;
                uB[v][k] = fpB[v][k] = fmB[v][k] = dfpB[v][k] = dfmB[v][k] = s1B[v][k] = s2B[v][k] = 0.0f;
                uC[v][k] = fpC[v][k] = fmC[v][k] = s1C[v][k] = s2C[v][k] = 0.0f;
                uN[v][k] = 1.0f;
            }
        }
        const int rbeg = r0 - 2, rend = r1 + 1;  // inclusive
        if (row_exists(rbeg)) load_row(rbeg, uN);
        for (int r = rbeg; r <= rend; r++) {
            float uNN[VEC][4] = {};
            float fpN[VEC][4] = {}, fmN[VEC][4] = {}, s1N[VEC][4] = {}, s2N[VEC][4] = {};
            if (r + 1 <= rend && row_exists(r + 1)) load_row(r + 1, uNN);  // prefetch one row ahead
            if (row_exists(r)) row_compute<2, BC, LIM, MODE, VEC>(uN, at_lo, at_hi, alpha, fpN, fmN, s1N, s2N);
            // limited x slopes of row r-1 (2nd_order_base_shll.c:268-276); first order in wall rows (:226-234,248-256)
            float dfpC[VEC][4], dfmC[VEC][4];
            {
                const int rc = r - 1;
                const bool wallrow = ((rc == 0) && lo_wall) || ((rc == nx - 1) && hi_wall);
#pragma unroll
                for (int v = 0; v < VEC; v++) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        dfpC[v][k] = wallrow ? 0.0f : limited_slope<LIM>(fpB[v][k], fpC[v][k], fpN[v][k], alpha);
                        dfmC[v][k] = wallrow ? 0.0f : limited_slope<LIM>(fmB[v][k], fmC[v][k], fmN[v][k], alpha);
                    }
                }
            }
            const int i = r - 2;
            if (i >= r0) {  // i < r1 by construction
                const bool lo = (i == 0) && lo_wall, hi = (i == nx - 1) && hi_wall;
                float uo[VEC][4];
#pragma unroll
                for (int v = 0; v < VEC; v++) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        float left = lo ? wall_flux<BC>(fpB[v][k], fmB[v][k], k, 1) : fpA[v][k];
                        float right = hi ? wall_flux<BC>(fmB[v][k], fpB[v][k], k, 1) : fmC[v][k];
                        float s = flux_sum<MODE>(fpB[v][k], fmB[v][k], right, left);
                        float t = apply_first<MODE>(uB[v][k], P.dtdx, s);                 // 2nd_order_base_shll.c:438
                        float ldf = lo ? 0.0f : dfpA[v][k];                               // :362-365
                        float rdf = hi ? 0.0f : dfmC[v][k];                               // :378-381
                        t = apply_second<MODE, POW2>(t, P.half_dtdx, slope_sum(dfpB[v][k], dfmB[v][k], rdf, ldf));  // :443
                        t = apply_first<MODE>(t, P.dtdy, s1B[v][k]);                      // :449
                        uo[v][k] = apply_second<MODE, POW2>(t, P.half_dtdy, s2B[v][k]);   // :454
                    }
                }
                store_row(i, uo);
            }
#pragma unroll
            for (int v = 0; v < VEC; v++) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    fpA[v][k] = fpB[v][k];
                    dfpA[v][k] = dfpB[v][k];
                    uB[v][k] = uC[v][k]; fpB[v][k] = fpC[v][k]; fmB[v][k] = fmC[v][k];
                    dfpB[v][k] = dfpC[v][k]; dfmB[v][k] = dfmC[v][k];
                    s1B[v][k] = s1C[v][k]; s2B[v][k] = s2C[v][k];
                    uC[v][k] = uN[v][k]; fpC[v][k] = fpN[v][k]; fmC[v][k] = fmN[v][k];
                    s1C[v][k] = s1N[v][k]; s2C[v][k] = s2N[v][k];
                    uN[v][k] = uNN[v][k];
                }
            }
        }
    }
    if (P.sync.enabled) {  // publish: our edge rows of this step have landed in the neighbours' halo rows
        if (touch_lo) halo_arrive(P.sync, P.sync.cnt_lo, P.sync.edge_warps_lo, P.sync.sig_lo);
        if (touch_hi) halo_arrive(P.sync, P.sync.cnt_hi, P.sync.edge_warps_hi, P.sync.sig_hi);
    }
}

}  // namespace shll
