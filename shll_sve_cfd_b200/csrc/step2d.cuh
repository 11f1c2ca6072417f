// step2d.cuh -- one fused SHLL time step on a 2D slab (replaces Compute_F_from_P + Update_U_from_F +
// Compute_P_from_U of base_shll_2d.c / 2nd_order_base_shll.c for one step).
//
// Data layout: SoA, one FP32 plane per conserved component, index = i*ny + j (i = x, slow axis), exactly the
// reference's layout (base_shll_2d.c:157).  Each plane carries 2 halo rows below row 0 and above row nx-1.
//
// Work decomposition -- no shared memory, no __syncthreads, every warp is independent:
//   * a warp owns a column tile of 32*VEC consecutive j and a chunk of consecutive rows i;
//   * it marches along i keeping the x-direction stencil (F+ of rows behind, F- of rows ahead, slopes) in a
//     register sliding window of 3 (order 1) or 4 (order 2) row slots whose roles rotate -- the loop is unrolled
//     by the window depth so no register is ever copied -- and each cell's primitives and split fluxes are
//     evaluated once per step (plus ORDER halo rows per chunk end);
//   * the y-direction neighbours (H+ from j-1, H- from j+1, limited slopes) come from the adjacent lanes by
//     warp shuffle; the outermost HL lanes of the warp are halo lanes that recompute the neighbouring tile's
//     edge cells, so tiles overlap by 2*HL*VEC columns and nothing crosses a warp;
//   * lane -> VEC consecutive j, so a warp reads 128*VEC contiguous bytes per plane per row (float/float2/float4);
//   * the next row's state is prefetched into the free `u` registers of the oldest slot one row ahead;
//   * walls are handled by patching the slot of the non-existent neighbour row (x) or the shuffled-in values (y)
//     inside warp-uniform branches, so interior warps / rows execute no boundary selects.
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
    float quarter;         // 0.25f, passed as a parameter so that it lives in a register (step2d_acc.cuh: one-LOP3 sign transfer)
    int peer_depth;        // step2d_acc.cuh: rows per side stored into the neighbour slabs' halos (0 = the scheme's order; 2 for a
                           // 1st-order context that also issues two-step launches, so that its halos are always two rows deep)
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

// One row of cells as held by a warp: VEC cells per lane.  Members that a given ORDER never touches are never
// materialised (everything is fully unrolled into registers).
template <int VEC>
struct RowSlot {
    float u[VEC][4];                 // conserved state (also the landing zone of the prefetch)
    float fp[VEC][4], fm[VEC][4];    // x split fluxes F+, F-
    float s1[VEC][4];                // hp - hm + Top - Bottom
    float s2[VEC][4];                // dhp + dhm - Top_df - Bottom_df          (order 2)
    float dfp[VEC][4], dfm[VEC][4];  // limited x slopes of F+, F-               (order 2)
};

// Per-thread, per-launch constants of the y direction.
template <int VEC>
struct YEdge {
    bool tile_has_wall;  // warp-uniform: this column tile contains j == 0 or j == ny-1
    bool ghost_lo;       // this lane's LAST cell is j = -1 (just below the y = 0 wall)
    bool ghost_hi;       // this lane's FIRST cell is j = ny (just above the y = ny-1 wall)
    bool outside[VEC];   // cell is outside the domain (TMA zero-fills those; the LDG kernel clamps the load)
    bool y_inner[VEC];   // order 2: false for wall cells (j = 0, ny-1) and for cells outside the domain
};

// y walls.  The reference builds the ghost flux of a wall cell from the cell's own split flux
// (REFLECT: Bottom = (-,-,+,-) * H-(own), base_shll_2d.c:186-189; OUTFLOW: Bottom = H+(own), 2nd_order_base_shll.c:282-285).
// Both are reproduced BIT FOR BIT by giving the lane just outside the domain a ghost *state* before the fluxes are
// evaluated -- the wall cell's state with the y momentum mirrored (REFLECT) or copied (OUTFLOW): every product in
// H+-(ghost) is then the corresponding product of H-+(own) with exact sign flips (RN is sign-symmetric), e.g.
// H+_0(ghost) = (-h0)(-Z3) + u0*Z2 = -H-_0(own).  The flux code therefore has no wall selects at all; only the first
// and last column tile execute this (warp-uniform branch).  tests: every golden fixture + random states, both BCs.
template <int BC, int VEC>
__device__ __forceinline__ void plant_y_ghosts(float (&u)[VEC][4], const YEdge<VEC> &Y)
{
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float from_hi = __shfl_down_sync(full, u[0][k], 1);      // state of cell j+1 (first cell of the next lane)
        const float from_lo = __shfl_up_sync(full, u[VEC - 1][k], 1);  // state of cell j-1 (last cell of the previous lane)
        const float sgn = (BC == BC_REFLECT && k == 2) ? -1.0f : 1.0f;
#pragma unroll
        for (int v = 0; v < VEC; v++)
            if (Y.outside[v]) u[v][k] = 1.0f;           // cells further out are never used: keep them a benign gas state
        if (Y.ghost_lo) u[VEC - 1][k] = sgn * from_hi;  // this lane's last cell is j = -1
        if (Y.ghost_hi) u[0][k] = sgn * from_lo;        // this lane's first cell is j = ny
    }
}

// ---- FAST mode, 2 cells per lane: the lane's two cells as the halves of packed FP32x2 operations (shll_math.cuh,
// "FAST mode on PAIRS OF CELLS").  Same storage (RowSlot<2>), same op order as the scalar code below, half the FP32 issue
// slots; only the y neighbours need re-pairing (one half comes from the other cell of the lane, one from a shuffle).
#define SHLL_PK(arr, k) v2mk((arr)[0][k], (arr)[1][k])
#define SHLL_UNPK(arr, k, val) do { const v2 t_ = (val); (arr)[0][k] = t_.x; (arr)[1][k] = t_.y; } while (0)

template <int ORDER, int BC, int LIM>
__device__ __forceinline__ void row_compute_fast_x2(RowSlot<2> &S, const YEdge<2> &Y, float alpha)
{
    const unsigned full = 0xffffffffu;
    if (Y.tile_has_wall) plant_y_ghosts<BC, 2>(S.u, Y);
    const v2 u[4] = {SHLL_PK(S.u, 0), SHLL_PK(S.u, 1), SHLL_PK(S.u, 2), SHLL_PK(S.u, 3)};
    v2 fp[4], fm[4], hp[4], hm[4];
    cell_flux_2d_fast_x2(u, fp, fm, hp, hm);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        SHLL_UNPK(S.fp, k, fp[k]);
        SHLL_UNPK(S.fm, k, fm[k]);
        const v2 bottom = v2mk(__shfl_up_sync(full, hp[k].y, 1), hp[k].x);  // H+ of cells j-1
        const v2 top = v2mk(hm[k].y, __shfl_down_sync(full, hm[k].x, 1));   // H- of cells j+1
        SHLL_UNPK(S.s1, k, v2sub(v2add(v2sub(hp[k], hm[k]), top), bottom));
        if (ORDER == 2) {
            const v2 hmL = v2mk(__shfl_up_sync(full, hm[k].y, 1), hm[k].x);
            const v2 hpR = v2mk(hp[k].y, __shfl_down_sync(full, hp[k].x, 1));
            v2 dhp = limited_slope_x2<LIM>(bottom, hp[k], hpR, alpha);
            v2 dhm = limited_slope_x2<LIM>(hmL, hm[k], top, alpha);
            dhp.x = Y.y_inner[0] ? dhp.x : 0.0f; dhp.y = Y.y_inner[1] ? dhp.y : 0.0f;
            dhm.x = Y.y_inner[0] ? dhm.x : 0.0f; dhm.y = Y.y_inner[1] ? dhm.y : 0.0f;
            const v2 bdf = v2mk(__shfl_up_sync(full, dhp.y, 1), dhp.x);
            const v2 tdf = v2mk(dhm.y, __shfl_down_sync(full, dhm.x, 1));
            SHLL_UNPK(S.s2, k, v2sub(v2sub(v2add(dhp, dhm), tdf), bdf));
        }
    }
}

template <class Ctx>
__device__ __forceinline__ void finish_o1_fast_x2(Ctx &X, int i, RowSlot<2> &A, RowSlot<2> &B, RowSlot<2> &C)
{
    const Step2DParams &P = *X.P;
    float uo[2][4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const v2 s = v2sub(v2add(v2sub(SHLL_PK(B.fp, k), SHLL_PK(B.fm, k)), SHLL_PK(C.fm, k)), SHLL_PK(A.fp, k));
        const v2 t = v2fma(v2bc(-P.dtdx), s, SHLL_PK(B.u, k));
        SHLL_UNPK(uo, k, v2fma(v2bc(-P.dtdy), SHLL_PK(B.s1, k), t));
    }
    X.template store_row<1>(i, uo);
}

template <int LIM>
__device__ __forceinline__ void slopes_x_fast_x2(const RowSlot<2> &B, RowSlot<2> &C, const RowSlot<2> &D, float alpha)
{
#pragma unroll
    for (int k = 0; k < 4; k++) {
        SHLL_UNPK(C.dfp, k, limited_slope_x2<LIM>(SHLL_PK(B.fp, k), SHLL_PK(C.fp, k), SHLL_PK(D.fp, k), alpha));
        SHLL_UNPK(C.dfm, k, limited_slope_x2<LIM>(SHLL_PK(B.fm, k), SHLL_PK(C.fm, k), SHLL_PK(D.fm, k), alpha));
    }
}

template <class Ctx>
__device__ __forceinline__ void update_o2_fast_x2(Ctx &X, int i, RowSlot<2> &A, RowSlot<2> &B, RowSlot<2> &C)
{
    const Step2DParams &P = *X.P;
    float uo[2][4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const v2 s = v2sub(v2add(v2sub(SHLL_PK(B.fp, k), SHLL_PK(B.fm, k)), SHLL_PK(C.fm, k)), SHLL_PK(A.fp, k));
        v2 t = v2fma(v2bc(-P.dtdx), s, SHLL_PK(B.u, k));
        const v2 d = v2sub(v2sub(v2add(SHLL_PK(B.dfp, k), SHLL_PK(B.dfm, k)), SHLL_PK(C.dfm, k)), SHLL_PK(A.dfp, k));
        t = v2fma(v2bc(-P.half_dtdx), d, t);
        t = v2fma(v2bc(-P.dtdy), SHLL_PK(B.s1, k), t);
        SHLL_UNPK(uo, k, v2fma(v2bc(-P.half_dtdy), SHLL_PK(B.s2, k), t));
    }
    X.template store_row<2>(i, uo);
}

// Fluxes of one row of cells held by the warp + everything the y direction contributes to their update.
template <int ORDER, int BC, int LIM, int MODE, int VEC>
__device__ __forceinline__ void row_compute(RowSlot<VEC> &S, const YEdge<VEC> &Y, float alpha)
{
    if constexpr (VEC == 2 && MODE == MODE_FAST) {
        row_compute_fast_x2<ORDER, BC, LIM>(S, Y, alpha);
        return;
    }
    const unsigned full = 0xffffffffu;
    if (Y.tile_has_wall) plant_y_ghosts<BC, VEC>(S.u, Y);  // warp-uniform: first / last column tile only
    float hp[VEC][4], hm[VEC][4];
#pragma unroll
    for (int v = 0; v < VEC; v++) cell_flux_2d<MODE, ORDER == 1>(S.u[v], S.fp[v], S.fm[v], hp[v], hm[v]);

    // neighbours across the lane boundary: H+ of cell j-1 ("bottom"), H- of cell j+1 ("top")
    float bottom[VEC][4], top[VEC][4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float hp_from_lo = __shfl_up_sync(full, hp[VEC - 1][k], 1);
        const float hm_from_hi = __shfl_down_sync(full, hm[0][k], 1);
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            bottom[v][k] = (v > 0) ? hp[v > 0 ? v - 1 : 0][k] : hp_from_lo;
            top[v][k] = (v < VEC - 1) ? hm[v < VEC - 1 ? v + 1 : 0][k] : hm_from_hi;
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; v++) flux_sum4(hp[v], hm[v], top[v], bottom[v], S.s1[v]);
    if (ORDER == 2) {
        float hmL[VEC][4], hpR[VEC][4];  // H- of j-1, H+ of j+1
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float hm_from_lo = __shfl_up_sync(full, hm[VEC - 1][k], 1);
            const float hp_from_hi = __shfl_down_sync(full, hp[0][k], 1);
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                hmL[v][k] = (v > 0) ? hm[v > 0 ? v - 1 : 0][k] : hm_from_lo;
                hpR[v][k] = (v < VEC - 1) ? hp[v < VEC - 1 ? v + 1 : 0][k] : hp_from_hi;
            }
        }
        float dhp[VEC][4], dhm[VEC][4];
#pragma unroll
        for (int v = 0; v < VEC; v++) {  // 2nd_order_base_shll.c:336-343; first order in wall cells (:292-300,314-322)
            limited_slope4<LIM>(bottom[v], hp[v], hpR[v], alpha, dhp[v]);
            limited_slope4<LIM>(hmL[v], hm[v], top[v], alpha, dhm[v]);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                dhp[v][k] = Y.y_inner[v] ? dhp[v][k] : 0.0f;
                dhm[v][k] = Y.y_inner[v] ? dhm[v][k] : 0.0f;
            }
        }
        // Bottom_df = dhp[j-1], Top_df = dhm[j+1]; both are 0 beyond a wall (:396-399,412-415) -- automatically,
        // because the ghost lane's y_inner is false as well.
        float bdf[VEC][4], tdf[VEC][4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float dhp_from_lo = __shfl_up_sync(full, dhp[VEC - 1][k], 1);
            const float dhm_from_hi = __shfl_down_sync(full, dhm[0][k], 1);
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                bdf[v][k] = (v > 0) ? dhp[v > 0 ? v - 1 : 0][k] : dhp_from_lo;
                tdf[v][k] = (v < VEC - 1) ? dhm[v < VEC - 1 ? v + 1 : 0][k] : dhm_from_hi;
            }
        }
#pragma unroll
        for (int v = 0; v < VEC; v++) slope_sum4(dhp[v], dhm[v], tdf[v], bdf[v], S.s2[v]);
    }
}

// Everything a warp needs to know about its rows.  All row predicates are reduced to comparisons against a few
// per-warp integers computed once, and all addressing is 32-bit (plane sizes are checked on the host).
template <int VEC>
struct RowCtx {
    const Step2DParams *P;
    int ny, r0, r1;
    int jl, j0;          // clamped load column, own column
    int rmin, rmax;      // rows that exist in memory: [rmin, rmax]
    int first_real_row, last_real_row;  // rows beyond a physical wall are ghosts: never recomputed, never sloped
    int wall_lo_row;     // 0 if local row 0 is a physical wall, else a row index that never occurs
    int wall_hi_row;     // nx-1 if local row nx-1 is a physical wall, else never
    int peer_lo_end;     // rows [0, peer_lo_end) are also stored into the lower neighbour's halo (0 if none)
    int peer_hi_begin;   // rows [peer_hi_begin, nx) are also stored into the upper neighbour's halo (INT_MAX if none)
    bool owner;
    YEdge<VEC> Y;

    __device__ __forceinline__ bool row_exists(int r) const { return r >= rmin && r <= rmax; }
    __device__ __forceinline__ void load_row(int r, float (&u)[VEC][4]) const
    {
        const int idx = r * ny + jl;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float t[VEC];
            VecIO<VEC>::load(P->in[k] + idx, t);
#pragma unroll
            for (int v = 0; v < VEC; v++) u[v][k] = t[v];
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
        // halo exchange fused into the step: edge rows also go straight into the neighbour GPU's halo rows
        if (i < peer_lo_end || i >= peer_hi_begin) {
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

// ---- order 1: finish row i.  A = row i-1 (F+ valid), B = row i (complete), C = row i+1 (fluxes just computed).
// X supplies the wall rows and the store (RowCtx for the LDG kernel, TmaCtx for the TMA kernel).
template <int BC, int VEC>
__device__ __forceinline__ void ghost_below(RowSlot<VEC> &G, const RowSlot<VEC> &B)
{   // physical wall below row B: the ghost row's F+ is B's own wall flux, its slope is zero
    // (base_shll_2d.c:152-155; 2nd_order_base_shll.c:216-219,362-365)
#pragma unroll
    for (int v = 0; v < VEC; v++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            G.fp[v][k] = wall_flux<BC>(B.fp[v][k], B.fm[v][k], k, 1);
            G.dfp[v][k] = 0.0f;
        }
}
template <int BC, int VEC>
__device__ __forceinline__ void ghost_above(RowSlot<VEC> &G, const RowSlot<VEC> &B)
{   // physical wall above row B (base_shll_2d.c:168-171; 2nd_order_base_shll.c:243-246,378-381)
#pragma unroll
    for (int v = 0; v < VEC; v++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            G.fm[v][k] = wall_flux<BC>(B.fm[v][k], B.fp[v][k], k, 1);
            G.dfm[v][k] = 0.0f;
        }
}

// ---- order 1: finish row i.  A = row i-1 (F+ valid), B = row i (complete), C = row i+1 (fluxes just computed).
// At a physical wall the missing neighbour slot already holds the ghost flux (ghost_below / ghost_above are applied
// where that row would have been computed), so the hot path has no boundary selects.
template <int BC, int MODE, int VEC, class Ctx>
__device__ __forceinline__ void finish_o1(const Ctx &X, int i, RowSlot<VEC> &A, RowSlot<VEC> &B, RowSlot<VEC> &C)
{
    if constexpr (VEC == 2 && MODE == MODE_FAST) {
        finish_o1_fast_x2(X, i, A, B, C);
    } else {
        const Step2DParams &P = *X.P;
        float uo[VEC][4];
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            float s[4], t[4];
            flux_sum4(B.fp[v], B.fm[v], C.fm[v], A.fp[v], s);
            apply_first4<MODE>(B.u[v], P.dtdx, s, t);         // base_shll_2d.c:227-230
            apply_first4<MODE>(t, P.dtdy, B.s1[v], uo[v]);    // base_shll_2d.c:232-235
        }
        X.template store_row<1>(i, uo);
    }
}

// ---- order 2: row r has just been computed into D.  A = row r-3 (F+, dF+ valid), B = row r-2 (finished now),
//      C = row r-1 (its x slopes are computed here).
template <int BC, int LIM, int MODE, int VEC, bool POW2, class Ctx>
__device__ __forceinline__ void finish_o2(const Ctx &X, int r, RowSlot<VEC> &A, RowSlot<VEC> &B, RowSlot<VEC> &C,
                                          RowSlot<VEC> &D)
{
    const Step2DParams &P = *X.P;
    {  // limited x slopes of row r-1 (2nd_order_base_shll.c:268-276); first order in wall rows (:226-234,248-256)
        const int rc = r - 1;
        if ((rc == X.wall_lo_row) || (rc == X.wall_hi_row)) {
#pragma unroll
            for (int v = 0; v < VEC; v++)
#pragma unroll
                for (int k = 0; k < 4; k++) C.dfp[v][k] = C.dfm[v][k] = 0.0f;
            // the rows beyond the wall do not exist: their slots get the ghost fluxes of the wall row (rarely taken branch)
            if (rc == X.wall_lo_row) ghost_below<BC, VEC>(B, C);   // B = row -1
            if (rc == X.wall_hi_row) ghost_above<BC, VEC>(D, C);   // D = row nx
        } else if (rc >= X.first_real_row && rc <= X.last_real_row) {
            if constexpr (VEC == 2 && MODE == MODE_FAST) {
                slopes_x_fast_x2<LIM>(B, C, D, P.alpha);
            } else {
#pragma unroll
                for (int v = 0; v < VEC; v++) {
                    limited_slope4<LIM>(B.fp[v], C.fp[v], D.fp[v], P.alpha, C.dfp[v]);
                    limited_slope4<LIM>(B.fm[v], C.fm[v], D.fm[v], P.alpha, C.dfm[v]);
                }
            }
        }
    }
    const int i = r - 2;
    if (i < X.r0) return;  // still filling the window
    if constexpr (VEC == 2 && MODE == MODE_FAST) {
        update_o2_fast_x2(X, i, A, B, C);
    } else {
        float uo[VEC][4];
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            float s[4], d[4], t1[4], t2[4], t3[4];
            flux_sum4(B.fp[v], B.fm[v], C.fm[v], A.fp[v], s);
            apply_first4<MODE>(B.u[v], P.dtdx, s, t1);                       // 2nd_order_base_shll.c:438
            slope_sum4(B.dfp[v], B.dfm[v], C.dfm[v], A.dfp[v], d);
            apply_second4<MODE, POW2>(t1, P.half_dtdx, d, t2);               // :443
            apply_first4<MODE>(t2, P.dtdy, B.s1[v], t3);                     // :449
            apply_second4<MODE, POW2>(t3, P.half_dtdy, B.s2[v], uo[v]);      // :454
        }
        X.template store_row<2>(i, uo);
    }
}

// LDG kernel steps: prefetch one row ahead into the oldest slot's free `u` registers, compute, finish.
template <int BC, int LIM, int MODE, int VEC>
__device__ __forceinline__ void step_o1(const RowCtx<VEC> &X, int i, RowSlot<VEC> &A, RowSlot<VEC> &B, RowSlot<VEC> &C)
{
    if (i >= X.r1) return;
    if (i + 2 <= X.r1 && X.row_exists(i + 2)) X.load_row(i + 2, A.u);
    if (X.row_exists(i + 1)) row_compute<1, BC, LIM, MODE, VEC>(C, X.Y, X.P->alpha);
    else ghost_above<BC, VEC>(C, B);  // i is the wall row nx-1
    finish_o1<BC, MODE, VEC>(X, i, A, B, C);
}

template <int BC, int LIM, int MODE, int VEC, bool POW2>
__device__ __forceinline__ void step_o2(const RowCtx<VEC> &X, int r, RowSlot<VEC> &A, RowSlot<VEC> &B, RowSlot<VEC> &C,
                                        RowSlot<VEC> &D)
{
    const int rend = X.r1 + 1;
    if (r > rend) return;
    if (r + 1 <= rend && X.row_exists(r + 1)) X.load_row(r + 1, A.u);
    if (X.row_exists(r)) row_compute<2, BC, LIM, MODE, VEC>(D, X.Y, X.P->alpha);
    finish_o2<BC, LIM, MODE, VEC, POW2>(X, r, A, B, C, D);
}

template <int ORDER, int BC, int LIM, int MODE, int VEC, bool POW2>
__global__ void __launch_bounds__(128) step2d_kernel(const __grid_constant__ Step2DParams P)
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

    RowCtx<VEC> X;
    X.P = &P;
    const int nx = P.nx;
    X.ny = P.ny;
    X.j0 = tile * USEFUL + (lane - HL) * VEC;
    X.jl = min(max(X.j0, 0), X.ny - VEC);  // clamped load column (lanes outside the domain hold unused values)
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
    X.rmin = P.lo_wall ? 0 : -2;
    X.rmax = P.hi_wall ? nx - 1 : nx + 1;
    X.wall_lo_row = P.lo_wall ? 0 : never;
    X.wall_hi_row = P.hi_wall ? nx - 1 : never;
    X.first_real_row = P.lo_wall ? 0 : never;
    X.last_real_row = P.hi_wall ? nx - 1 : -never;
    X.peer_lo_end = (P.sync.enabled && P.lo_peer[0] != nullptr) ? ORDER : 0;
    X.peer_hi_begin = (P.sync.enabled && P.hi_peer[0] != nullptr) ? nx - ORDER : 0x7fffffff;
    const bool touch_lo = (X.r0 < ORDER), touch_hi = (X.r1 > nx - ORDER);
    if (P.sync.enabled) {  // wait until the neighbour GPUs' edge rows of the previous step sit in our halo rows
        if (touch_lo) halo_wait(P.sync, P.sync.wait_lo);
        if (touch_hi) halo_wait(P.sync, P.sync.wait_hi, 1);
    }

    if (ORDER == 1) {
        RowSlot<VEC> A, B, C;
#pragma unroll
        for (int v = 0; v < VEC; v++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                A.fp[v][k] = 0.0f;
                C.fm[v][k] = 0.0f;
                C.u[v][k] = 1.0f;
            }
        if (X.row_exists(X.r0 - 1)) {  // F+ of the row below the chunk
            X.load_row(X.r0 - 1, A.u);
            row_compute<1, BC, LIM, MODE, VEC>(A, X.Y, P.alpha);
        }
        X.load_row(X.r0, B.u);
        if (X.row_exists(X.r0 + 1)) X.load_row(X.r0 + 1, C.u);
        row_compute<1, BC, LIM, MODE, VEC>(B, X.Y, P.alpha);
        if (X.r0 == X.wall_lo_row) ghost_below<BC, VEC>(A, B);
        for (int i = X.r0; i < X.r1; i += 3) {  // roles rotate: no register copies
            step_o1<BC, LIM, MODE, VEC>(X, i, A, B, C);
            step_o1<BC, LIM, MODE, VEC>(X, i + 1, B, C, A);
            step_o1<BC, LIM, MODE, VEC>(X, i + 2, C, A, B);
        }
    } else {
        RowSlot<VEC> A, B, C, D;
#pragma unroll
        for (int v = 0; v < VEC; v++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                A.fp[v][k] = A.dfp[v][k] = 0.0f;
                B.u[v][k] = B.fp[v][k] = B.fm[v][k] = B.dfp[v][k] = B.dfm[v][k] = B.s1[v][k] = B.s2[v][k] = 0.0f;
                C.u[v][k] = C.fp[v][k] = C.fm[v][k] = C.s1[v][k] = C.s2[v][k] = C.dfp[v][k] = C.dfm[v][k] = 0.0f;
                D.u[v][k] = 1.0f;
                D.fp[v][k] = D.fm[v][k] = D.s1[v][k] = D.s2[v][k] = 0.0f;
            }
        const int rbeg = X.r0 - 2;
        if (X.row_exists(rbeg)) X.load_row(rbeg, D.u);
        for (int r = rbeg; r <= X.r1 + 1; r += 4) {
            step_o2<BC, LIM, MODE, VEC, POW2>(X, r, A, B, C, D);
            step_o2<BC, LIM, MODE, VEC, POW2>(X, r + 1, B, C, D, A);
            step_o2<BC, LIM, MODE, VEC, POW2>(X, r + 2, C, D, A, B);
            step_o2<BC, LIM, MODE, VEC, POW2>(X, r + 3, D, A, B, C);
        }
    }
    if (P.sync.enabled) {  // publish: our edge rows of this step have landed in the neighbours' halo rows
        if (touch_lo) halo_arrive(P.sync, P.sync.cnt_lo, P.sync.edge_warps_lo, P.sync.sig_lo);
        if (touch_hi) halo_arrive(P.sync, P.sync.cnt_hi, P.sync.edge_warps_hi, P.sync.sig_hi);
    }
}

}  // namespace shll
