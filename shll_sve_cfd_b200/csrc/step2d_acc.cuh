// step2d_acc.cuh -- FAST-mode fused 2D step in FACE-FLUX ACCUMULATE form: 2 cells per lane, TMA-fed.
//
// Same scheme as step2d.cuh (Compute_F_from_P + Update_U_from_F + Compute_P_from_U of base_shll_2d.c /
// 2nd_order_base_shll.c), arranged for the issue-bound regime the 2nd-order kernel lives in:
//
//  * The reference's two updates per direction (2nd_order_base_shll.c:436-458)
//        U -= DT_ON_DX*(F+ - F- + Right - Left);  U -= 0.5*DT_ON_DX*(dF+ + dF- - Right_df - Left_df)
//    are the difference of the reconstructed face fluxes  Phi+ = F+ + dF+/2,  Phi- = F- - dF-/2:
//        U -= DT_ON_DX*(Phi+[i] - Phi-[i] + Phi-[i+1] - Phi+[i-1]).
//    The same real-number expression with fewer roundings (this is FAST mode: tolerance-matched, not bit-exact), and it
//    means a row only has to carry Phi+ of the row behind it and the part of the flux difference that is already known
//    -- 8 values per component instead of the 4-row x 7-array window of the bit-exact kernel -- so 2 cells per lane fit
//    in registers without spills and every flux / slope / update operation is a packed FP32x2 instruction on the lane's
//    cell pair.
//  * The kernel works with G = -F- = f*Z3 + U*Z2 (and its face value Gamma = G - dG/2 = -Phi-): no negations anywhere,
//    the signs fold into the operand modifiers of the packed adds.
//  * minmod(l, r) (2nd_order_base_shll.c:191-201) = (sign(l) + sign(r))/2 * min(|l|, |r|): ONE LOP3 per difference
//    builds +-0.25 with the sign of the difference, the sum of two of them is the weight w in {+-0.5, 0} and
//    Phi = F + w*min(|l|,|r|) is one FFMA2.  3 ALU-pipe operations per slope instead of 2 FSETP + 2 FSEL + a product.
//    (Differs from the reference only when l*r underflows to zero: |slope| < 1e-19.)
//    MC limiter (base-omp/2nd_order_base_shll.c:317-325): w * min(|l + r|/2, alpha*min(|l|,|r|)).
//  * Backward differences are reused as the next cell's / next row's forward difference.
//  * Wall code is compiled out of the hot path twice over: column tiles that touch a y wall run their own instantiation
//    (WALLTILE), and only the boxes that contain a physical x wall run the EDGE row routine.
//
// Rows march as in step2d_tma.cuh: one warp per block, a private shared-memory ring of TMA boxes (4 rows x 68 columns x
// 4 planes per cp.async.bulk.tensor.3d), conflict-free LDS.64, STG.64 of the finished row; multi-GPU edge rows are
// stored a second time into the neighbour's halo (halo_sync.cuh).
#pragma once
#include "step2d_tma.cuh"

namespace shll {

// cell_flux_2d_fast_x2 with the minus fluxes un-negated: gm = -F-, hgm = -H-.
__device__ __forceinline__ void cell_flux_2d_fast_x2g(const v2 (&u)[4], v2 (&fp)[4], v2 (&gm)[4], v2 (&hp)[4], v2 (&hgm)[4])
{
    const v2 r = v2rcp_newton(u[0]);
    const v2 ux = v2mul(u[1], r), uy = v2mul(u[2], r);
    const v2 k = v2fma(ux, ux, v2mul(uy, uy));
    const v2 T = v2mul(v2fma(v2bc(-0.5f), k, v2mul(u[3], r)), v2bc(1.0f / SHLL_CV_F));
    const v2 g = v2mul(v2bc(SHLL_GAMMA_F), T);
    const v2 inv_a = v2rsqrt_newton(g);
    const v2 a = v2mul(g, inv_a);
    const v2 P = v2mul(u[0], T);
    const v2 eP = v2add(u[3], P);
    v2 f[4], h[4];
    f[0] = u[1]; f[1] = v2fma(u[1], ux, P); f[2] = v2mul(u[1], uy); f[3] = v2mul(ux, eP);
    h[0] = u[2]; h[1] = v2mul(u[2], ux); h[2] = v2fma(u[2], uy, P); h[3] = v2mul(uy, eP);
    const v2 ha = v2mul(v2bc(0.5f), a);
    {
        const v2 M = v2mul(ux, inv_a);
        const v2 z1 = v2fma(v2bc(0.5f), M, v2bc(0.5f)), z3 = v2fma(v2bc(0.5f), M, v2bc(-0.5f));
        const v2 z2 = v2mul(ha, v2fma(v2neg(M), M, v2bc(1.0f)));
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const v2 uz = v2mul(u[c], z2);
            fp[c] = v2fma(f[c], z1, uz);
            gm[c] = v2fma(f[c], z3, uz);
        }
    }
    {
        const v2 M = v2mul(uy, inv_a);
        const v2 z1 = v2fma(v2bc(0.5f), M, v2bc(0.5f)), z3 = v2fma(v2bc(0.5f), M, v2bc(-0.5f));
        const v2 z2 = v2mul(ha, v2fma(v2neg(M), M, v2bc(1.0f)));
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const v2 uz = v2mul(u[c], z2);
            hp[c] = v2fma(h[c], z1, uz);
            hgm[c] = v2fma(h[c], z3, uz);
        }
    }
}

// Wall ghost fluxes at an x wall, in (F+, G = -F-) form.  REFLECT (base_shll_2d.c:152-155,168-171): the ghost flux is the
// cell's own opposite split flux, + for the wall-normal momentum component (k == 1) and - otherwise.  OUTFLOW
// (2nd_order_base_shll.c:216-219,243-246): the cell's own same split flux.
template <int BC>
__device__ __forceinline__ v2 ghost_fp_below(v2 fp, v2 g, int k)   // F+ of the ghost row below a wall row
{
    if (BC == BC_REFLECT) return (k == 1) ? v2neg(g) : g;   // (k==1) ? F- : -F-
    return fp;
}
template <int BC>
__device__ __forceinline__ v2 ghost_g_above(v2 fp, v2 g, int k)    // G = -F- of the ghost row above a wall row
{
    if (BC == BC_REFLECT) return (k == 1) ? v2neg(fp) : fp;  // F-ghost = (k==1) ? F+ : -F+
    return g;
}

// What a warp carries from row to row (per conserved component, per cell pair).  The x update of row i is applied in one
// piece, U'' = U' - DT_ON_DX*(D[i] - Gamma[i+1]) with D[i] = (Phi+[i] - Phi+[i-1]) + Gamma[i]: the bracket is a small
// difference (exactly zero in a uniform gas, like the reference's sum), so a gas at rest stays bit-for-bit at rest.
// (Applying D[i] and Gamma[i+1] to U in two steps was measured 5-10x noisier against the reference: every step then
// rounds U - DT_ON_DX*Gamma, which is not small, everywhere in the domain.)
template <int ORDER>
struct AccState {
    v2 fpP[4];   // F+ of the previous row
    v2 D[4];     // D of row r-ORDER
    v2 accB[4];  // row r-ORDER: state minus its y-direction update
    v2 gP[4];    // order 2: G of the previous row
    v2 ep[4];    //          F+[r-1] - F+[r-2]
    v2 em[4];    //          G[r-1] - G[r-2]
    v2 PhiP[4];  //          Phi+ of row r-2
    v2 acc0[4];  //          row r-1: state minus its y-direction update
};

struct AccRows {   // warp-uniform row bookkeeping
    int r0, r1;                    // rows [r0, r1) are stored by this warp
    int rmin, rmax;                // rows that exist in memory
    int wall_lo_row, wall_hi_row;  // 0 / nx-1 where that row is a physical wall, else a value no row takes
    int noslope_lo, noslope_hi;    // rows strictly between them get limited x slopes
    float quarter, nquarter;       // 0.25f / -0.25f from kernel parameters: registers, so the sign transfer is one LOP3
    uint32_t stash;                // STASH: shared address of this lane's slot in the 2-row stash of (state - y update)
    uint32_t sstage_base;          // TMA store: shared address of the warp's store stage (128-byte aligned)
    uint32_t sstage;               //            ... of this lane's cell pair in row 0, plane 0 of it
    bool stager;                   //            this lane owns columns of the tile (lanes 1..30)
};

// Order 2 keeps `state minus y update` of rows r-1 and r-2 until their x update completes.  With STASH these 16 registers
// live in a per-warp shared-memory stash instead (2 rows x 4 components x 32 lanes x 8 bytes; row r uses slot r & 1);
// STASH level 2 also keeps the backward differences of F+ and G there (16 more registers).  8 / 24 LDS.64 + STS.64 per row
// in exchange for more resident warps per scheduler.
__device__ __forceinline__ v2 stash_load(uint32_t a)
{
    v2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void stash_store(uint32_t a, v2 v)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}

// Store stage of a warp: the 4 rows an interior box finishes, dense as the TMA store box wants them --
// [4 planes][4 rows][60 columns] floats.  Lane l (1..30) owns columns 2(l-1), 2(l-1)+1: conflict-free STS.64 at constant
// offsets, no address arithmetic, no bounds predicate (the TMA unit clips the ragged last tile).
constexpr uint32_t ACC_OUT_COLS = 60, ACC_OUT_ROW_BYTES = ACC_OUT_COLS * 4, ACC_OUT_PLANE_BYTES = 4 * ACC_OUT_ROW_BYTES;
constexpr uint32_t ACC_OUT_STAGE_BYTES = 4 * ACC_OUT_PLANE_BYTES;  // 3840
template <int WOUT>
__device__ __forceinline__ void acc_stage_row(const AccRows &W, const v2 (&o)[4])
{
    if (W.stager) {
#pragma unroll
        for (int k = 0; k < 4; k++) stash_store(W.sstage + k * ACC_OUT_PLANE_BYTES + WOUT * ACC_OUT_ROW_BYTES, o[k]);
    }
}

// Split fluxes of row r and its state minus the y-direction update (accN).
template <int ORDER, int BC, int LIM, bool WALLTILE>
__device__ __forceinline__ void acc_row_y(const YEdge<2> &Y, const Step2DParams &P, const AccRows &W, float (&uin)[2][4],
                                          v2 (&fp)[4], v2 (&g)[4], v2 (&accN)[4])
{
    const unsigned full = 0xffffffffu;
    if (WALLTILE) {
        // y walls as ghost STATES (see plant_y_ghosts): mirrored / copied wall cell just outside the domain, and a gas at
        // rest in the cells further out (their values are never used, they only have to stay finite).
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float from_hi = __shfl_down_sync(full, uin[0][k], 1);
            const float from_lo = __shfl_up_sync(full, uin[1][k], 1);
            const float sgn = (BC == BC_REFLECT && k == 2) ? -1.0f : 1.0f;
            const float rest = (k == 0 || k == 3) ? 1.0f : 0.0f;
            if (Y.outside[0]) uin[0][k] = rest;
            if (Y.outside[1]) uin[1][k] = rest;
            if (Y.ghost_lo) uin[1][k] = sgn * from_hi;
            if (Y.ghost_hi) uin[0][k] = sgn * from_lo;
        }
    }
    const v2 u[4] = {v2mk(uin[0][0], uin[1][0]), v2mk(uin[0][1], uin[1][1]), v2mk(uin[0][2], uin[1][2]), v2mk(uin[0][3], uin[1][3])};
    v2 hp[4], hg[4];
    cell_flux_2d_fast_x2g(u, fp, g, hp, hg);
    // slopes vanish in wall cells and beyond (2nd_order_base_shll.c:292-300,314-322,396-399,412-415)
    v2 q = v2bc(W.quarter), nq = v2bc(W.nquarter);
    if (WALLTILE) {
        q = v2mk(Y.y_inner[0] ? W.quarter : 0.0f, Y.y_inner[1] ? W.quarter : 0.0f);
        nq = v2mk(Y.y_inner[0] ? W.nquarter : 0.0f, Y.y_inner[1] ? W.nquarter : 0.0f);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        v2 ysum;
        if (ORDER == 1) {
            // (H+ - H-) - (H+ of cell j-1 - H- of cell j+1), base_shll_2d.c:232-235; the neighbour terms as scalar sums so that
            // no register pair has to be formed around the shuffled values
            const float up = __shfl_up_sync(full, hp[k].y, 1), dn = __shfl_down_sync(full, hg[k].x, 1);
            ysum = v2sub(v2add(hp[k], hg[k]), v2mk(fadd(hg[k].y, up), fadd(dn, hp[k].x)));
        } else {
            // backward differences d[j] = H[j] - H[j-1]; the forward difference of cell j is d[j+1].  The forward difference of
            // the lane's upper cell is computed here from the neighbour lane's H (the very subtraction that lane performs for its
            // own backward difference: same operands, same bits) rather than fetched from it -- all four shuffles of a component
            // then depend on the fluxes only, one level of shuffle latency less on the critical path of the row.
            const float hp_up = __shfl_up_sync(full, hp[k].y, 1), hp_dn = __shfl_down_sync(full, hp[k].x, 1);
            const float hg_up = __shfl_up_sync(full, hg[k].y, 1), hg_dn = __shfl_down_sync(full, hg[k].x, 1);
            const v2 dlp = v2mk(fsub(hp[k].x, hp_up), fsub(hp[k].y, hp[k].x));
            const v2 dlm = v2mk(fsub(hg[k].x, hg_up), fsub(hg[k].y, hg[k].x));
            const v2 drp = v2mk(dlp.y, fsub(hp_dn, hp[k].y));
            const v2 drm = v2mk(dlm.y, fsub(hg_dn, hg[k].y));
            const v2 psi = v2fma(limiter_weight(dlp, drp, q), limiter_magnitude<LIM>(dlp, drp, P.alpha), hp[k]);      // H+ + dH+/2
            const v2 gam = v2fma(limiter_weight_neg(dlm, drm, nq), limiter_magnitude<LIM>(dlm, drm, P.alpha), hg[k]);  // -(H- - dH-/2)
            const float up = __shfl_up_sync(full, psi.y, 1), dn = __shfl_down_sync(full, gam.x, 1);
            ysum = v2sub(v2add(psi, gam), v2mk(fadd(gam.y, up), fadd(dn, psi.x)));
        }
        accN[k] = v2fma(v2bc(-P.dtdy), ysum, u[k]);
    }
}

// Finished row i leaves: local store, and for the multi-GPU edge rows a second store into the neighbour's halo (kept
// out of line: it runs for ORDER rows per slab side only).
static __device__ __noinline__ void acc_store_peer(const Step2DParams *P, bool lo, int pidx, v2 o0, v2 o1, v2 o2, v2 o3)
{
    const v2 o[4] = {o0, o1, o2, o3};
#pragma unroll
    for (int k = 0; k < 4; k++) *reinterpret_cast<float2 *>((lo ? P->lo_peer[k] : P->hi_peer[k]) + pidx) = o[k];
}
template <class Ctx>
__device__ __forceinline__ void acc_store(const Ctx &X, int i, const v2 (&o)[4])
{
    if (X.owner) {
        const int idx = i * X.ny + X.j0;
#pragma unroll
        for (int k = 0; k < 4; k++) *reinterpret_cast<float2 *>(X.P->out[k] + idx) = o[k];
        if (i < X.peer_lo_end || i >= X.peer_hi_begin) {
            const bool lo = i < X.peer_lo_end;
            acc_store_peer(X.P, lo, lo ? idx : (i - X.peer_hi_begin) * X.ny + X.j0, o[0], o[1], o[2], o[3]);
        }
    }
}
// Interior boxes: the row is stored unconditionally by the owner lanes, no peer copy.
// (Four predicated STG.64 in one asm block were tried instead of the branch: all four addresses and values live at once
// cost 40+ registers of pressure and spills.)
template <class Ctx>
__device__ __forceinline__ void acc_store_owned(const Ctx &X, int i, const v2 (&o)[4])
{
    if (X.owner) {
        const int idx = i * X.ny + X.j0;
#pragma unroll
        for (int k = 0; k < 4; k++) *reinterpret_cast<float2 *>(X.P->out[k] + idx) = o[k];
    }
}

// One step of the march: row r arrives, row r-ORDER leaves.  EDGE = this box may contain a physical x wall.
// WOUT >= 0: the finished row is row WOUT of the box being staged for a TMA store; -1: plain STG.64.
template <int ORDER, int BC, int LIM, bool WALLTILE, bool EDGE, int STASH, int WOUT, class Ctx>
__device__ __forceinline__ void acc_row(const Ctx &X, const AccRows &W, AccState<ORDER> &S, int r, float (&uin)[2][4])
{
    const Step2DParams &P = *X.P;
    const v2 mdtdx = v2bc(-P.dtdx);
    v2 fp[4], g[4], accN[4], o[4];
    if (ORDER == 1) {
        if (EDGE && (r < W.rmin || r > W.rmax)) return;  // beyond a wall: the wall row was finished when it was computed
        acc_row_y<1, BC, LIM, WALLTILE>(X.Y, P, W, uin, fp, g, accN);
        if (EDGE && r == W.wall_lo_row) {  // Left of row 0 = its own wall flux (base_shll_2d.c:152-155)
#pragma unroll
            for (int k = 0; k < 4; k++) S.fpP[k] = ghost_fp_below<BC>(fp[k], g[k], k);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) o[k] = v2fma(mdtdx, v2sub(S.D[k], g[k]), S.accB[k]);  // Right of row r-1 = F-[r] = -G[r]
        if (EDGE) { if (r - 1 >= W.r0) acc_store(X, r - 1, o); }
        else if (WOUT >= 0) acc_stage_row<(WOUT >= 0 ? WOUT : 0)>(W, o);
        else acc_store_owned(X, r - 1, o);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            S.D[k] = v2add(v2sub(fp[k], S.fpP[k]), g[k]);
            S.accB[k] = accN[k];
            S.fpP[k] = fp[k];
        }
        if (EDGE && r == W.wall_hi_row && r < W.r1) {  // Right of row nx-1 = its own wall flux (base_shll_2d.c:168-171)
#pragma unroll
            for (int k = 0; k < 4; k++) o[k] = v2fma(mdtdx, v2sub(S.D[k], ghost_g_above<BC>(fp[k], g[k], k)), accN[k]);
            acc_store(X, r, o);
        }
    } else {
        if (EDGE && (r < W.rmin || r > W.rmax)) {
            // Row r does not exist.  r == nx above a wall: finish rows nx-2 and nx-1 here -- the wall row is first order
            // (2nd_order_base_shll.c:248-256) and its Right flux is its own wall flux, no slope beyond (:243-246,378-381).
            if (r == W.wall_hi_row + 1) {
                if (STASH) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        S.accB[k] = stash_load(W.stash + ((r & 1) << 10) + k * 256);        // row r-2
                        S.acc0[k] = stash_load(W.stash + (((r + 1) & 1) << 10) + k * 256);  // row r-1
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) o[k] = v2fma(mdtdx, v2sub(S.D[k], S.gP[k]), S.accB[k]);
                if (r - 2 >= W.r0) acc_store(X, r - 2, o);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const v2 d1 = v2add(v2sub(S.fpP[k], S.PhiP[k]), S.gP[k]);
                    o[k] = v2fma(mdtdx, v2sub(d1, ghost_g_above<BC>(S.fpP[k], S.gP[k], k)), S.acc0[k]);
                }
                if (r - 1 >= W.r0) acc_store(X, r - 1, o);
            }
            return;
        }
        acc_row_y<2, BC, LIM, WALLTILE>(X.Y, P, W, uin, fp, g, accN);
        const int rc = r - 1;  // the row whose face fluxes are completed by this step
        // wall rows are first order (2nd_order_base_shll.c:226-234,248-256): zero weight instead of a branch
        float qx = W.quarter, nqx = W.nquarter;
        if (EDGE && !(rc > W.noslope_lo && rc < W.noslope_hi)) qx = nqx = 0.0f;
        v2 php[4], gam[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const v2 eNp = v2sub(fp[k], S.fpP[k]), eNm = v2sub(g[k], S.gP[k]);
            const v2 ep = (STASH >= 2) ? stash_load(W.stash + 2048 + k * 256) : S.ep[k];
            const v2 em = (STASH >= 2) ? stash_load(W.stash + 3072 + k * 256) : S.em[k];
            php[k] = v2fma(limiter_weight(ep, eNp, v2bc(qx)), limiter_magnitude<LIM>(ep, eNp, P.alpha), S.fpP[k]);      // :268-276
            gam[k] = v2fma(limiter_weight_neg(em, eNm, v2bc(nqx)), limiter_magnitude<LIM>(em, eNm, P.alpha), S.gP[k]);
            if (STASH >= 2) {
                stash_store(W.stash + 2048 + k * 256, eNp);
                stash_store(W.stash + 3072 + k * 256, eNm);
            } else {
                S.ep[k] = eNp;
                S.em[k] = eNm;
            }
        }
        if (EDGE && rc == W.wall_lo_row) {  // Left of row 0 = its own wall flux, no slope beyond the wall (:216-219,362-365)
#pragma unroll
            for (int k = 0; k < 4; k++) S.PhiP[k] = ghost_fp_below<BC>(S.fpP[k], S.gP[k], k);
        }
        const uint32_t slot = W.stash + ((r & 1) << 10);  // holds row r-2, then row r
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const v2 yB = STASH ? stash_load(slot + k * 256) : S.accB[k];
            o[k] = v2fma(mdtdx, v2sub(S.D[k], gam[k]), yB);  // Right of row r-2 = Phi-[r-1] = -Gamma[r-1]
        }
        if (EDGE) { if (r - 2 >= W.r0) acc_store(X, r - 2, o); }
        else if (WOUT >= 0) acc_stage_row<(WOUT >= 0 ? WOUT : 0)>(W, o);
        else acc_store_owned(X, r - 2, o);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            S.D[k] = v2add(v2sub(php[k], S.PhiP[k]), gam[k]);
            if (STASH) {
                stash_store(slot + k * 256, accN[k]);
            } else {
                S.accB[k] = S.acc0[k];
                S.acc0[k] = accN[k];
            }
            S.PhiP[k] = php[k];
            S.fpP[k] = fp[k];
            S.gP[k] = g[k];
        }
    }
}

// read row `within` (run-time) of the box sitting in `stage`: 2 cells per lane
template <class Ctx>
__device__ __forceinline__ void acc_read_row(const Ctx &X, int stage, int within, float (&u)[2][4])
{
    const uint32_t a = X.ring + Ctx::STAGE_STRIDE * stage + within * Ctx::ROW_BYTES + X.lane_off;
#pragma unroll
    for (int k = 0; k < 4; k++)
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(u[0][k]), "=f"(u[1][k]) : "r"(a + k * Ctx::PLANE_BYTES));
}

template <int ORDER, int BC, int LIM, bool WALLTILE, int STASH, bool TSTORE, class Ctx>
__device__ __forceinline__ void acc_march(Ctx &X, const AccRows &W, int rbeg, int rlast, int lane, int tile)
{
    AccState<ORDER> S;
#pragma unroll
    for (int k = 0; k < 4; k++)
        S.fpP[k] = S.D[k] = S.accB[k] = S.gP[k] = S.ep[k] = S.em[k] = S.PhiP[k] = S.acc0[k] = v2bc(0.0f);
    int stage = 0;
    uint32_t parity = 0;
    float uin[2][4];
    const int nx = X.P->nx;
    bool store_pending = false;  // lane 0 has a TMA store in flight that may still be reading the store stage
    for (int box = 0; box < X.nboxes; box++) {
        const int r = rbeg + 4 * box;
        mbar_wait(X.bars + 8u * stage, parity);
        // a box needs the EDGE routine if one of its rows is missing, is a wall row, or completes / follows a wall row
        // ... or if one of the rows it finishes is not stored by this warp / is also stored into a neighbour GPU's halo
        const int ifirst = r - ORDER;  // rows finished by this box: ifirst .. ifirst + 3
        const bool edge = (r <= 1 && X.P->lo_wall) || (r + 3 >= nx - 1 && X.P->hi_wall) || (r + 3 > rlast) || (ifirst < W.r0) ||
                          (ifirst + 3 >= W.r1) || (ifirst < X.peer_lo_end) || (ifirst + 3 >= X.peer_hi_begin);
        if (edge) {
#pragma unroll 1
            for (int w = 0; w < 4; w++) {
                if (r + w > rlast) break;
                acc_read_row(X, stage, w, uin);
                acc_row<ORDER, BC, LIM, WALLTILE, true, STASH, -1>(X, W, S, r + w, uin);
            }
            __syncwarp();
            if (lane == 0 && box + X.stages < X.nboxes) X.arm(box + X.stages, stage);
        } else {
            if (TSTORE) {  // the previous box's TMA store must have finished READING the store stage before it is overwritten
                if (store_pending) {
                    if (lane == 0) tma_store_wait_read();
                    __syncwarp();
                }
            }
            X.template read_row<0>(stage, uin);
            acc_row<ORDER, BC, LIM, WALLTILE, false, STASH, TSTORE ? 0 : -1>(X, W, S, r, uin);
            X.template read_row<1>(stage, uin);
            acc_row<ORDER, BC, LIM, WALLTILE, false, STASH, TSTORE ? 1 : -1>(X, W, S, r + 1, uin);
            X.template read_row<2>(stage, uin);
            acc_row<ORDER, BC, LIM, WALLTILE, false, STASH, TSTORE ? 2 : -1>(X, W, S, r + 2, uin);
            X.template read_row<3>(stage, uin);
            acc_row<ORDER, BC, LIM, WALLTILE, false, STASH, TSTORE ? 3 : -1>(X, W, S, r + 3, uin);
            if (TSTORE) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this lane's STS -> visible to the TMA unit
            // the box has been fully read (and its results staged): store them, refill the ring stage
            __syncwarp();
            if (lane == 0) {
                if (TSTORE) tma_store_3d(&X.T->tmap_out, tile * (int)ACC_OUT_COLS, ifirst + 2, 0, W.sstage_base);
                if (box + X.stages < X.nboxes) X.arm(box + X.stages, stage);
            }
            store_pending = TSTORE;
        }
        stage++;
        if (stage == X.stages) { stage = 0; parity ^= 1u; }
    }
    if (TSTORE && lane == 0) tma_store_wait_all();  // shared memory may not be released under a store that is still reading it
}

// MINB = resident warps per SM the register allocation is capped for (launch bounds); STASH see above.
template <int ORDER, int BC, int LIM, int MINB, int STASH>
__global__ void __launch_bounds__(32, MINB) step2d_acc_kernel(const __grid_constant__ Step2DTmaParams T)
{
    constexpr int R = 4, VEC = 2, HL = 1;
    constexpr int USEFUL = (32 - 2 * HL) * VEC;
    extern __shared__ __align__(128) unsigned char smem[];
    const Step2DParams &P = T.base;
    const int lane = threadIdx.x;
    const int gw = blockIdx.x;
    const int tile = gw % P.ntiles;
    int chunk = gw / P.ntiles;
    if (P.nchunks > 2) chunk = (chunk == 0) ? 0 : (chunk == 1 ? P.nchunks - 1 : chunk - 1);  // edge chunks first

    TmaCtx<VEC, R> X;
    X.P = &P;
    X.T = &T;
    const int nx = P.nx;
    X.ny = P.ny;
    const int xs = tile * USEFUL - HL * VEC;
    X.x0 = xs & ~3;
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
    X.wall_lo_row = X.wall_hi_row = X.first_real_row = X.last_real_row = never;  // (window-kernel fields, unused here)
    const int depth = P.peer_depth > 0 ? P.peer_depth : ORDER;  // rows exchanged per side
    X.peer_lo_end = (P.sync.enabled && P.lo_peer[0] != nullptr) ? depth : 0;
    X.peer_hi_begin = (P.sync.enabled && P.hi_peer[0] != nullptr) ? nx - depth : 0x7fffffff;
    AccRows W;
    W.r0 = X.r0;
    W.r1 = X.r1;
    W.rmin = P.lo_wall ? 0 : -2;
    W.rmax = P.hi_wall ? nx - 1 : nx + 1;
    W.wall_lo_row = P.lo_wall ? 0 : never;
    W.wall_hi_row = P.hi_wall ? nx - 1 : never;
    W.noslope_lo = P.lo_wall ? 0 : never;
    W.noslope_hi = P.hi_wall ? nx - 1 : -never;
    W.quarter = P.quarter;
    W.nquarter = -P.quarter;
    W.stash = 0;
    const bool touch_lo = (X.r0 < depth), touch_hi = (X.r1 > nx - depth);
    if (P.sync.enabled) {
        if (touch_lo) halo_wait(P.sync, P.sync.wait_lo, 0);
        if (touch_hi) halo_wait(P.sync, P.sync.wait_hi, 1);
    }

    const int rbeg = X.r0 - ORDER;
    const int rlast = X.r1 - 1 + ORDER;
    X.stages = T.stages;
    X.ring = smem_u32(smem);
    X.bars = X.ring + TmaCtx<VEC, R>::STAGE_STRIDE * X.stages;
    X.lane_off = (uint32_t)(xs - X.x0 + lane * VEC) * 4u;
    X.ybase = rbeg + 2;
    X.nboxes = (rlast - rbeg) / R + 1;
    uint32_t smem_top = X.bars + 8u * X.stages;
    if (STASH) {  // the host adds 2 KB (level 1) / 4 KB (level 2) + 16 bytes to the dynamic shared memory
        W.stash = ((smem_top + 15u) & ~15u) + 8u * lane;
        for (int i = 0; i < (STASH >= 2 ? 16 : 8); i++) stash_store(W.stash + i * 256, v2bc(0.0f));
        smem_top = ((smem_top + 15u) & ~15u) + (STASH >= 2 ? 4096u : 2048u);
    }
    // store stage (the host adds ACC_OUT_STAGE_BYTES + 128 when T.tma_store is set)
    W.sstage_base = (smem_top + 127u) & ~127u;
    W.sstage = W.sstage_base + 8u * (uint32_t)(lane - HL);
    W.stager = (lane >= HL) && (lane < 32 - HL);
    if (lane == 0) {
        for (int s = 0; s < X.stages; s++) mbar_init(X.bars + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    step2d_wait_for_input(T, tile, chunk, lane);  // everything above only touched registers and shared memory
    if (lane == 0) {
        for (int b = 0; b < X.stages && b < X.nboxes; b++) X.arm(b, b);
    }
    __syncwarp();

    if (X.Y.tile_has_wall) acc_march<ORDER, BC, LIM, true, STASH, false>(X, W, rbeg, rlast, lane, tile);  // (few tiles: keep one code path)
    else if (T.tma_store) acc_march<ORDER, BC, LIM, false, STASH, true>(X, W, rbeg, rlast, lane, tile);
    else acc_march<ORDER, BC, LIM, false, STASH, false>(X, W, rbeg, rlast, lane, tile);

    step2d_publish_output(T, tile, chunk, lane);
    if (P.sync.enabled) {
        if (touch_lo) halo_arrive(P.sync, P.sync.cnt_lo, P.sync.edge_warps_lo, P.sync.sig_lo);
        if (touch_hi) halo_arrive(P.sync, P.sync.cnt_hi, P.sync.edge_warps_hi, P.sync.sig_hi);
    }
}

// =========================================================================================================================
// Two time steps per pass (order 1, FAST): step2d_acc2_kernel.
//
// The 1st-order kernel is not short of HBM bandwidth alone -- it waits on it (long-scoreboard stalls at the mbarriers) while
// its issue slots are 55 % used.  Here a warp chains TWO steps in registers: stage A turns the rows of U^n arriving from the
// ring into rows of U^(n+1) exactly as the one-step kernel does, and instead of storing them hands each finished row to stage B
// -- the same row routine with its own carried state -- which finishes the rows of U^(n+2) that leave.  HBM sees one read and
// one write per TWO steps; a row of U^(n+1) never exists in memory.  The lane layout needs no change: a halo lane holds two
// cells, which is exactly the two-cell reach of two 1st-order steps (stage A is valid on 62 of the warp's 64 columns, stage B
// on the 60 owned ones); a chunk reads two halo rows per side instead of one (the same redundancy per step).  Every row goes
// through the same arithmetic as in the one-step kernel, so the results are bit-for-bit the ones of two one-step launches
// (tests/test_gpu_fast_parity.py::test_fused_two_step_launches_give_the_bits_of_single_steps).
//
// One arriving row: returns which finished rows it produced -- bit 0: `lo` = row r-1, bit 1: `hi` = row r itself (r is the
// physical wall row nx-1, whose Right flux is its own wall flux, base_shll_2d.c:168-171).
template <int BC, bool WALLTILE, bool EDGE>
__device__ __forceinline__ int acc1_advance(const YEdge<2> &Y, const Step2DParams &P, const AccRows &W, AccState<1> &S, int r,
                                            float (&uin)[2][4], v2 (&lo)[4], v2 (&hi)[4])
{
    if (EDGE && (r < W.rmin || r > W.rmax)) return 0;
    const v2 mdtdx = v2bc(-P.dtdx);
    v2 fp[4], g[4], accN[4];
    acc_row_y<1, BC, LIM_MINMOD, WALLTILE>(Y, P, W, uin, fp, g, accN);
    if (EDGE && r == W.wall_lo_row) {  // Left of row 0 = its own wall flux (base_shll_2d.c:152-155)
#pragma unroll
        for (int k = 0; k < 4; k++) S.fpP[k] = ghost_fp_below<BC>(fp[k], g[k], k);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) lo[k] = v2fma(mdtdx, v2sub(S.D[k], g[k]), S.accB[k]);  // Right of row r-1 = F-[r] = -G[r]
#pragma unroll
    for (int k = 0; k < 4; k++) {
        S.D[k] = v2add(v2sub(fp[k], S.fpP[k]), g[k]);
        S.accB[k] = accN[k];
        S.fpP[k] = fp[k];
    }
    int made = 1;
    if (EDGE && r == W.wall_hi_row) {
#pragma unroll
        for (int k = 0; k < 4; k++) hi[k] = v2fma(mdtdx, v2sub(S.D[k], ghost_g_above<BC>(fp[k], g[k], k)), accN[k]);
        made = 3;
    }
    return made;
}

__device__ __forceinline__ void acc_as_input(const v2 (&o)[4], float (&u)[2][4])
{
#pragma unroll
    for (int k = 0; k < 4; k++) { u[0][k] = o[k].x; u[1][k] = o[k].y; }
}

template <int BC, bool WALLTILE, bool TSTORE, class Ctx>
__device__ __forceinline__ void acc2_march(Ctx &X, const AccRows &W, int rbeg, int rlast, int lane, int tile)
{
    AccState<1> SA, SB;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        SA.fpP[k] = SA.D[k] = SA.accB[k] = SA.gP[k] = SA.ep[k] = SA.em[k] = SA.PhiP[k] = SA.acc0[k] = v2bc(0.0f);
        SB.fpP[k] = SB.D[k] = SB.accB[k] = SB.gP[k] = SB.ep[k] = SB.em[k] = SB.PhiP[k] = SB.acc0[k] = v2bc(0.0f);
    }
    const Step2DParams &P = *X.P;
    const int nx = P.nx;
    // first row of U^(n+1) stage A can make: the row after its first input row -- or row 0 itself at a physical wall (ghost flux)
    const int a0 = (P.lo_wall && rbeg <= 0) ? 0 : rbeg + 1;
    // ... and the first row of U^(n+2) stage B can make, by the same rule applied to the rows it receives
    const int b0 = (P.lo_wall && a0 == 0) ? 0 : a0 + 1;
    int stage = 0;
    uint32_t parity = 0;
    float uin[2][4], umid[2][4];
    v2 oA[4], oA2[4], oB[4], oB2[4];
    bool store_pending = false;
    // stage B: row i of U^(n+1) arrives; rows of U^(n+2) inside [r0, r1) leave (EDGE form)
    auto feed_b = [&](int i, const v2(&mid)[4]) {
        acc_as_input(mid, umid);
        const int made = acc1_advance<BC, WALLTILE, true>(X.Y, P, W, SB, i, umid, oB, oB2);
        if ((made & 1) && i - 1 >= b0 && i - 1 >= W.r0 && i - 1 < W.r1) acc_store(X, i - 1, oB);
        if ((made & 2) && i < W.r1) acc_store(X, i, oB2);
    };
    for (int box = 0; box < X.nboxes; box++) {
        const int r = rbeg + 4 * box;
        mbar_wait(X.bars + 8u * stage, parity);
        const int ifirst = r - 2;  // rows of U^(n+2) finished by this box: ifirst .. ifirst + 3
        const bool edge = (P.lo_wall && r <= 1) || (P.hi_wall && r + 3 >= nx - 1) || (r + 3 > rlast) || (ifirst < W.r0) || (ifirst + 3 >= W.r1) ||
                          (ifirst < X.peer_lo_end) || (ifirst + 3 >= X.peer_hi_begin);
        if (edge) {
#pragma unroll 1
            for (int w = 0; w < 4; w++) {
                const int rr = r + w;
                if (rr > rlast) break;
                acc_read_row(X, stage, w, uin);
                const int made = acc1_advance<BC, WALLTILE, true>(X.Y, P, W, SA, rr, uin, oA, oA2);
                if ((made & 1) && rr - 1 >= a0) feed_b(rr - 1, oA);
                if (made & 2) feed_b(rr, oA2);
            }
            __syncwarp();
            if (lane == 0 && box + X.stages < X.nboxes) X.arm(box + X.stages, stage);
        } else {
            if (TSTORE) {
                if (store_pending) {
                    if (lane == 0) tma_store_wait_read();
                    __syncwarp();
                }
            }
#define SHLL_ACC2_ROW(WI)                                                                               \
            X.template read_row<WI>(stage, uin);                                                        \
            acc1_advance<BC, WALLTILE, false>(X.Y, P, W, SA, r + WI, uin, oA, oA2);                     \
            acc_as_input(oA, umid);                                                                     \
            acc1_advance<BC, WALLTILE, false>(X.Y, P, W, SB, r + WI - 1, umid, oB, oB2);                \
            if (TSTORE) acc_stage_row<WI>(W, oB); else acc_store_owned(X, r + WI - 2, oB);
            SHLL_ACC2_ROW(0)
            SHLL_ACC2_ROW(1)
            SHLL_ACC2_ROW(2)
            SHLL_ACC2_ROW(3)
#undef SHLL_ACC2_ROW
            if (TSTORE) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (TSTORE) tma_store_3d(&X.T->tmap_out, tile * (int)ACC_OUT_COLS, ifirst + 2, 0, W.sstage_base);
                if (box + X.stages < X.nboxes) X.arm(box + X.stages, stage);
            }
            store_pending = TSTORE;
        }
        stage++;
        if (stage == X.stages) { stage = 0; parity ^= 1u; }
    }
    if (TSTORE && lane == 0) tma_store_wait_all();
}

template <int BC, int MINB>
__global__ void __launch_bounds__(32, MINB) step2d_acc2_kernel(const __grid_constant__ Step2DTmaParams T)
{
    constexpr int R = 4, VEC = 2, HL = 1, DEPTH = 2;  // DEPTH: halo rows per side = steps per pass
    constexpr int USEFUL = (32 - 2 * HL) * VEC;
    extern __shared__ __align__(128) unsigned char smem[];
    const Step2DParams &P = T.base;
    const int lane = threadIdx.x;
    const int gw = blockIdx.x;
    const int tile = gw % P.ntiles;
    int chunk = gw / P.ntiles;
    if (P.nchunks > 2) chunk = (chunk == 0) ? 0 : (chunk == 1 ? P.nchunks - 1 : chunk - 1);  // edge chunks first

    TmaCtx<VEC, R> X;
    X.P = &P;
    X.T = &T;
    const int nx = P.nx;
    X.ny = P.ny;
    const int xs = tile * USEFUL - HL * VEC;
    X.x0 = xs & ~3;
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
    X.wall_lo_row = X.wall_hi_row = X.first_real_row = X.last_real_row = never;
    X.peer_lo_end = (P.sync.enabled && P.lo_peer[0] != nullptr) ? DEPTH : 0;
    X.peer_hi_begin = (P.sync.enabled && P.hi_peer[0] != nullptr) ? nx - DEPTH : 0x7fffffff;
    AccRows W;
    W.r0 = X.r0;
    W.r1 = X.r1;
    W.rmin = P.lo_wall ? 0 : -2;
    W.rmax = P.hi_wall ? nx - 1 : nx + 1;
    W.wall_lo_row = P.lo_wall ? 0 : never;
    W.wall_hi_row = P.hi_wall ? nx - 1 : never;
    W.noslope_lo = never;
    W.noslope_hi = -never;
    W.quarter = P.quarter;
    W.nquarter = -P.quarter;
    W.stash = 0;
    const bool touch_lo = (X.r0 < DEPTH), touch_hi = (X.r1 > nx - DEPTH);
    if (P.sync.enabled) {
        if (touch_lo) halo_wait(P.sync, P.sync.wait_lo, 0);
        if (touch_hi) halo_wait(P.sync, P.sync.wait_hi, 1);
    }
    const int rbeg = X.r0 - DEPTH;
    const int rlast = X.r1 - 1 + DEPTH;
    X.stages = T.stages;
    X.ring = smem_u32(smem);
    X.bars = X.ring + TmaCtx<VEC, R>::STAGE_STRIDE * X.stages;
    X.lane_off = (uint32_t)(xs - X.x0 + lane * VEC) * 4u;
    X.ybase = rbeg + 2;
    X.nboxes = (rlast - rbeg) / R + 1;
    const uint32_t smem_top = X.bars + 8u * X.stages;
    W.sstage_base = (smem_top + 127u) & ~127u;
    W.sstage = W.sstage_base + 8u * (uint32_t)(lane - HL);
    W.stager = (lane >= HL) && (lane < 32 - HL);
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

    if (X.Y.tile_has_wall) acc2_march<BC, true, false>(X, W, rbeg, rlast, lane, tile);
    else if (T.tma_store) acc2_march<BC, false, true>(X, W, rbeg, rlast, lane, tile);
    else acc2_march<BC, false, false>(X, W, rbeg, rlast, lane, tile);

    step2d_publish_output(T, tile, chunk, lane);
    if (P.sync.enabled) {
        if (touch_lo) halo_arrive(P.sync, P.sync.cnt_lo, P.sync.edge_warps_lo, P.sync.sig_lo);
        if (touch_hi) halo_arrive(P.sync, P.sync.cnt_hi, P.sync.edge_warps_hi, P.sync.sig_hi);
    }
}

}  // namespace shll
