// step1d_acc.cuh -- FAST-mode 1D 2nd-order step in face-flux form (the 1D twin of step2d_acc.cuh).
//
// Same warp tile as step1d.cuh (120 owned cells, one float4 per lane per plane, lanes 0 / 31 are halo lanes), but
//   * the lane's four cells are two packed FP32x2 pairs (c0,c1), (c2,c3) through the flux evaluation, the face fluxes and
//     the update -- the float4 load already leaves them in aligned register pairs;
//   * limited slopes as weight * magnitude (sign transfer by one LOP3, see shll_math.cuh) on backward differences that
//     every cell shares with its right neighbour;
//   * the two updates of the reference (2nd_order_base_shll.c:436-446, x direction) as one difference of face fluxes
//         U -= DT_ON_DX * ((Phi+[j] - Phi+[j-1]) + (Gamma[j] - Gamma[j+1])),   Phi+ = F+ + dF+/2,  Gamma = -(F- - dF-/2),
//     whose bracket is exactly zero in a uniform gas.
// Everything that crosses a pair boundary (differences, neighbour face fluxes) is done with scalar adds writing straight
// into the halves of the result pairs, so no register pair is ever formed by moves.
#pragma once
#include "step1d.cuh"

namespace shll {

// Split fluxes of a pair of 1D cells (rho, rho*u, E): F+ and G = -F-.  base_shll.c:135-157, arithmetic of prim1d_fast.
__device__ __forceinline__ void cell_flux_1d_fast_x2g(const v2 (&u)[3], v2 (&fp)[3], v2 (&gm)[3])
{
    const v2 r = v2rcp_newton(u[0]);
    const v2 ux = v2mul(u[1], r);
    const v2 T = v2mul(v2fma(v2mul(v2bc(-0.5f), ux), ux, v2mul(u[2], r)), v2bc(1.0f / SHLL_CV_F));
    const v2 g = v2mul(v2bc(SHLL_GAMMA_F), T);
    const v2 inv_a = v2rsqrt_newton(g);
    const v2 a = v2mul(g, inv_a);
    const v2 P = v2mul(u[0], T);
    const v2 f[3] = {u[1], v2fma(u[1], ux, P), v2mul(ux, v2add(u[2], P))};
    const v2 M = v2mul(ux, inv_a);
    const v2 z1 = v2fma(v2bc(0.5f), M, v2bc(0.5f)), z3 = v2fma(v2bc(0.5f), M, v2bc(-0.5f));
    const v2 z2 = v2mul(v2mul(v2bc(0.5f), a), v2fma(v2neg(M), M, v2bc(1.0f)));
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const v2 uz = v2mul(u[c], z2);
        fp[c] = v2fma(f[c], z1, uz);
        gm[c] = v2fma(f[c], z3, uz);
    }
}

// One 2nd-order step of a lane's four cells (two packed pairs): split fluxes, limited face fluxes, update.  `u` -> `uo`.
// Lanes 0 and 31 produce garbage in their outer two cells (no neighbour beyond the warp); their inner two are valid.
template <int BC, int LIM, bool EDGE>
__device__ __forceinline__ void step1d_acc_update(const Step1DParams &P, int j0, int n, const v2 (&u)[2][3], v2 (&uo)[2][3])
{
    constexpr int VEC = 4;
    const unsigned full = 0xffffffffu;
    const bool lo_wall = P.lo_wall != 0, hi_wall = P.hi_wall != 0;
    v2 fp[2][3], g[2][3];  // [pair][component]: pair 0 = cells j0, j0+1; pair 1 = cells j0+2, j0+3
    cell_flux_1d_fast_x2g(u[0], fp[0], g[0]);
    cell_flux_1d_fast_x2g(u[1], fp[1], g[1]);

    // wall cells: first order (no slope), ghost flux from the cell's own split fluxes -- only EDGE tiles can hold one
    bool at_lo[VEC], at_hi[VEC];
    v2 q[2] = {v2bc(P.quarter), v2bc(P.quarter)}, nq[2] = {v2bc(-P.quarter), v2bc(-P.quarter)};
    if (EDGE) {
        float qs[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            at_lo[v] = lo_wall && (j0 + v == 0);
            at_hi[v] = hi_wall && (j0 + v == n - 1);
            qs[v] = (at_lo[v] || at_hi[v]) ? 0.0f : P.quarter;
        }
        q[0] = v2mk(qs[0], qs[1]); q[1] = v2mk(qs[2], qs[3]);
        nq[0] = v2neg(q[0]); nq[1] = v2neg(q[1]);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        // backward differences d[j] = F[j] - F[j-1]; the forward difference of cell j is d[j+1]
        const float fp_lo = __shfl_up_sync(full, fp[1][k].y, 1), g_lo = __shfl_up_sync(full, g[1][k].y, 1);
        const v2 dp0 = v2mk(fsub(fp[0][k].x, fp_lo), fsub(fp[0][k].y, fp[0][k].x));
        const v2 dp1 = v2mk(fsub(fp[1][k].x, fp[0][k].y), fsub(fp[1][k].y, fp[1][k].x));
        const v2 dg0 = v2mk(fsub(g[0][k].x, g_lo), fsub(g[0][k].y, g[0][k].x));
        const v2 dg1 = v2mk(fsub(g[1][k].x, g[0][k].y), fsub(g[1][k].y, g[1][k].x));
        const v2 rp0 = v2mk(dp0.y, dp1.x), rp1 = v2mk(dp1.y, __shfl_down_sync(full, dp0.x, 1));
        const v2 rg0 = v2mk(dg0.y, dg1.x), rg1 = v2mk(dg1.y, __shfl_down_sync(full, dg0.x, 1));
        // face fluxes (2nd_order_base_shll.c:268-276 with the 0.5 of :443 folded into the weight)
        const v2 phi0 = v2fma(limiter_weight(dp0, rp0, q[0]), limiter_magnitude<LIM>(dp0, rp0, P.alpha), fp[0][k]);
        const v2 phi1 = v2fma(limiter_weight(dp1, rp1, q[1]), limiter_magnitude<LIM>(dp1, rp1, P.alpha), fp[1][k]);
        const v2 gam0 = v2fma(limiter_weight_neg(dg0, rg0, nq[0]), limiter_magnitude<LIM>(dg0, rg0, P.alpha), g[0][k]);
        const v2 gam1 = v2fma(limiter_weight_neg(dg1, rg1, nq[1]), limiter_magnitude<LIM>(dg1, rg1, P.alpha), g[1][k]);
        float phi_lo[VEC] = {__shfl_up_sync(full, phi1.y, 1), phi0.x, phi0.y, phi1.x};   // Phi+ of cell j-1
        float gam_hi[VEC] = {gam0.y, gam1.x, gam1.y, __shfl_down_sync(full, gam0.x, 1)};  // Gamma of cell j+1
        if (EDGE) {
            const float fpv[VEC] = {fp[0][k].x, fp[0][k].y, fp[1][k].x, fp[1][k].y};
            const float gv[VEC] = {g[0][k].x, g[0][k].y, g[1][k].x, g[1][k].y};
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                // reflective ends: (-,+,-) on (rho, rho*u, E), base_shll.c:95-97,108-110; outflow: own flux (2nd_order_base_shll.c:216-219)
                const float ghost_lo = (BC == BC_REFLECT) ? ((k == 1) ? -gv[v] : gv[v]) : fpv[v];   // F+ beyond the lower wall
                const float ghost_hi = (BC == BC_REFLECT) ? ((k == 1) ? -fpv[v] : fpv[v]) : gv[v];  // G = -F- beyond the upper wall
                phi_lo[v] = at_lo[v] ? ghost_lo : phi_lo[v];
                gam_hi[v] = at_hi[v] ? ghost_hi : gam_hi[v];
            }
        }
        const v2 s0 = v2mk(fsub(phi0.x, phi_lo[0]), fsub(phi0.y, phi_lo[1])), s1 = v2mk(fsub(phi1.x, phi_lo[2]), fsub(phi1.y, phi_lo[3]));
        const v2 t0 = v2mk(fsub(gam0.x, gam_hi[0]), fsub(gam0.y, gam_hi[1])), t1 = v2mk(fsub(gam1.x, gam_hi[2]), fsub(gam1.y, gam_hi[3]));
        uo[0][k] = v2fma(v2bc(-P.dtdx), v2add(s0, t0), u[0][k]);
        uo[1][k] = v2fma(v2bc(-P.dtdx), v2add(s1, t1), u[1][k]);
    }
}

// The tile's three float4 are passed in (the kernel below streams them through a shared-memory ring).  NSUB = 2: two time steps
// per launch -- the lane's four halo cells are exactly the reach of two 2nd-order steps, so the second step runs on the first
// one's result in registers (valid on the 120 owned cells); HBM sees one read and one write per two steps.  The slab machinery
// (step1d.cuh) is unchanged: such a launch sits at round positions (p, p + 1), the host extends the range for position p.
template <int BC, int LIM, bool EDGE, int NSUB>
__device__ __forceinline__ void step1d_acc_tile(const Step1DParams &P, int tile, int lane, const float4 (&in)[3])
{
    constexpr int VEC = 4;
    constexpr int USEFUL = 30 * VEC;
    const int n = P.n;
    const int j0 = tile * USEFUL + (lane - 1) * VEC;
    const int own_lo = tile * USEFUL, own_hi = min(own_lo + USEFUL, n);
    // real (unshifted) cell indices: send steps may run on an extended range (ext_lo > 0 when two steps share a launch)
    const int e = P.ext_lo, nr = P.n_real;
    const bool touch_lo = EDGE && P.xch > 0 && (own_lo - e < P.xch) && (own_hi - e > 0);
    const bool touch_hi = EDGE && P.xch > 0 && (own_hi - e > nr - P.xch) && (own_lo - e < nr);
    if (EDGE && P.sync.enabled && P.recv)
        step1d_recv_halo(P, tile == 0, (long)tile * USEFUL + 124 > (long)P.ext_lo + P.n_real);

    v2 u[2][3], uo[2][3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float4 t = in[k];
        // an EDGE tile may read halo cells a neighbour GPU has only just written: load it (again) AFTER the halo wait above
        if (EDGE) t = *reinterpret_cast<const float4 *>(P.in[k] + min(j0, ((n + 3) & ~3)));
        u[0][k] = v2mk(t.x, t.y);
        u[1][k] = v2mk(t.z, t.w);
    }
    if (NSUB == 1) {
        step1d_acc_update<BC, LIM, EDGE>(P, j0, n, u, uo);
    } else {
        v2 um[2][3];
        step1d_acc_update<BC, LIM, EDGE>(P, j0, n, u, um);
        step1d_acc_update<BC, LIM, EDGE>(P, j0, n, um, uo);
    }

    if (!EDGE) {  // interior tile: every owner lane stores three full float4
        if (lane != 0 && lane != 31) {
#pragma unroll
            for (int k = 0; k < 3; k++)
                *reinterpret_cast<float4 *>(P.out[k] + j0) = make_float4(uo[0][k].x, uo[0][k].y, uo[1][k].x, uo[1][k].y);
        }
        return;
    }
    const bool owner = !(lane == 0 || lane == 31 || j0 >= n);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (!owner) continue;
        const float o[VEC] = {uo[0][k].x, uo[0][k].y, uo[1][k].x, uo[1][k].y};
        if (j0 + VEC <= n) {
            *reinterpret_cast<float4 *>(P.out[k] + j0) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int v = 0; v < VEC; v++)
                if (j0 + v < n) P.out[k][j0 + v] = o[v];
        }
        // halo exchange fused into the step: the outermost xch real cells go straight into the neighbour GPU's mailbox
        const int jr = j0 - e;
        if (P.lo_peer[k] != nullptr && jr < P.xch) {
#pragma unroll
            for (int v = 0; v < VEC; v++)
                if (jr + v >= 0 && jr + v < P.xch && jr + v < nr) P.lo_peer[k][jr + v] = o[v];
        }
        if (P.hi_peer[k] != nullptr && jr + VEC > nr - P.xch) {
#pragma unroll
            for (int v = 0; v < VEC; v++)
                if (jr + v >= nr - P.xch && jr + v < nr) P.hi_peer[k][jr + v - (nr - P.xch)] = o[v];
        }
    }
    if (P.sync.enabled) {
        if (touch_lo) halo_arrive(P.sync, P.sync.cnt_lo, P.sync.edge_warps_lo, P.sync.sig_lo);
        if (touch_hi) halo_arrive(P.sync, P.sync.cnt_hi, P.sync.edge_warps_hi, P.sync.sig_hi);
    }
}

// MINB: resident 128-thread blocks per SM the register allocation is capped for.
template <int BC, int LIM, int MINB>
__global__ void __launch_bounds__(128, MINB) step1d_acc_kernel(const Step1DParams P)
{
    step1d_ring_march(P, [&](int tile, int lane, bool interior, const float4(&cur)[3]) {
        if (interior) step1d_acc_tile<BC, LIM, false, 1>(P, tile, lane, cur);
        else step1d_acc_tile<BC, LIM, true, 1>(P, tile, lane, cur);
    }, 2);
}

// Two time steps per launch (see step1d_acc_tile).  `order` = 4 for the interior test: a tile is interior when the two-step
// reach of its 128 loaded cells stays inside the owned cells.
template <int BC, int LIM, int MINB>
__global__ void __launch_bounds__(128, MINB) step1d_acc2_kernel(const Step1DParams P)
{
    step1d_ring_march(P, [&](int tile, int lane, bool interior, const float4(&cur)[3]) {
        if (interior) step1d_acc_tile<BC, LIM, false, 2>(P, tile, lane, cur);
        else step1d_acc_tile<BC, LIM, true, 2>(P, tile, lane, cur);
    }, 4);
}

}  // namespace shll
