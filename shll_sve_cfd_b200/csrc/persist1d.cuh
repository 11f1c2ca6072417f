// persist1d.cuh -- persistent, register-resident 1D time march for the launch-bound regime
// (BASELINE.json configs[1]: 65 536 cells x 104 858 steps, where one step is ~100x shorter than a kernel launch).
//
// One cooperative launch runs ALL time steps.  The tube is cut into one segment per block (one block per SM); every
// thread keeps ONE cell's conserved state in registers for the whole run -- HBM is touched at the start, at the end and
// for the halo strips, never per step.
//
//   * inside a block, neighbours come from warp shuffles; the four edge lanes of every warp publish their split fluxes
//     in a double-buffered shared-memory mailbox, so one __syncthreads per step closes the +-1 (order 1) or +-2
//     (order 2, flux slopes recomputed redundantly for the neighbouring warp's edge cell) stencil across warps;
//   * between blocks, temporal blocking: a block carries K*ORDER halo cells per side and advances K steps without
//     talking to anyone (the halo's garbage front moves ORDER cells inward per step and never reaches the owned cells),
//     then writes its 2 boundary strips to global memory, raises a per-block round counter (st.release.gpu) and waits for
//     its two neighbours' counters (ld.acquire.gpu) -- a point-to-point handshake, not a grid-wide barrier;
//   * strips are double-buffered by round parity; a neighbour can never be more than one round ahead.
// The per-cell arithmetic is the same device code as the streaming kernel (cell_flux_1d, limited_slope, apply_*), so
// the result is bit-identical to it (and, in STRICT mode, to the reference).
#pragma once
#include "halo_sync.cuh"
#include "shll_math.cuh"

namespace shll {

struct Persist1DParams {
    const float *in[3];   // plane base = cell 0
    float *out[3];
    float *strips;        // [2 parities][nblocks][2 sides][3 comps][hmax] floats
    unsigned *round_done; // [nblocks] rounds completed by each block (monotonic over the life of the context)
    unsigned *err;
    unsigned round_base;  // value of round_done[] before this launch
    int n;                // cells
    int nblocks;
    int K;                // steps per round
    int hmax;             // strip capacity in cells (= K * ORDER)
    long nsteps;
    float dtdx, half_dtdx, alpha;
    float quarter;        // 0.25f (see step1d_acc.cuh)
    unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int ORDER, int BC, int LIM, int MODE, int TFORM, bool POW2>
__global__ void __launch_bounds__(1024, 1) persist1d_kernel(const Persist1DParams P)
{
    extern __shared__ float mailbox[];  // [2][nwarps][4 edge lanes][6] : fp[3], fm[3] of lanes 0, 1, 30, 31
    __shared__ unsigned sh_abort;
    const unsigned full = 0xffffffffu;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;
    const int b = blockIdx.x;
    const int h = P.K * ORDER;                       // halo cells per side
    const int s0 = (int)(((long)b * P.n) / P.nblocks);        // first owned cell (balanced partition)
    const int s1 = (int)(((long)(b + 1) * P.n) / P.nblocks);  // one past the last owned cell
    const int c = s0 - h + t;                        // this thread's cell
    const int ext = (s1 - s0) + 2 * h;               // threads that carry a cell
    const bool carries = (t < ext);
    const bool in_domain = carries && c >= 0 && c < P.n;
    const bool owned = carries && c >= s0 && c < s1;
    const bool at_lo = (c == 0), at_hi = (c == P.n - 1);  // physical walls (single GPU: both ends are walls)
    const bool have_lo = (b > 0), have_hi = (b < P.nblocks - 1);

    float u[3];
#pragma unroll
    for (int k = 0; k < 3; k++) u[k] = in_domain ? P.in[k][c] : 1.0f;

    const long nrounds = (P.nsteps + P.K - 1) / P.K;
    for (long r = 0; r < nrounds; r++) {
        const int ksteps = (int)min((long)P.K, P.nsteps - r * P.K);
        if (r > 0) {
            // ---- halo refresh from the neighbours' strips of round r-1
            const unsigned want = P.round_base + (unsigned)r;
            if (t < 2) {  // thread 0 watches the lower neighbour, thread 1 the upper one; the clock is read every 1024 polls only
                const bool have = (t == 0) ? have_lo : have_hi;
                const unsigned *flag = P.round_done + (t == 0 ? b - 1 : b + 1);
                if (have && (int)(ld_acquire_gpu(flag) - want) < 0) {
                    const unsigned long long t0 = globaltimer_ns();
                    unsigned polls = 0;
                    while ((int)(ld_acquire_gpu(flag) - want) < 0) {
                        if ((++polls & 1023u) == 0 && (globaltimer_ns() - t0 > P.timeout_ns || *(volatile unsigned *)P.err)) {
                            atomicExch(P.err, 1u);
                            break;
                        }
                    }
                }
            }
            __syncthreads();
            if (t == 0) sh_abort = *(volatile unsigned *)P.err;
            __syncthreads();
            if (sh_abort) return;  // a neighbour never showed up: the whole block bails out (the host reports the error)
            const int par = (int)((r - 1) & 1);
            const size_t side_sz = (size_t)3 * P.hmax;
            if (have_lo && t < h) {  // left halo = lower neighbour's LAST h owned cells (its side 1)
                const float *src = P.strips + (((size_t)par * P.nblocks + (b - 1)) * 2 + 1) * side_sz;
#pragma unroll
                for (int k = 0; k < 3; k++) u[k] = __ldcg(src + k * P.hmax + t);
            }
            if (have_hi && t >= ext - h && t < ext) {  // right halo = upper neighbour's FIRST h owned cells (its side 0)
                const float *src = P.strips + (((size_t)par * P.nblocks + (b + 1)) * 2 + 0) * side_sz;
#pragma unroll
                for (int k = 0; k < 3; k++) u[k] = __ldcg(src + k * P.hmax + (t - (ext - h)));
            }
        }

        // ---- K steps on registers
        for (int s = 0; s < ksteps; s++) {
            float fp[3], fm[3];
            cell_flux_1d<MODE, TFORM>(u, fp, fm);
            float *box = mailbox + (size_t)(s & 1) * nwarps * 24;
            if (lane < 2 || lane >= 30) {
                float *slot = box + (warp * 4 + (lane < 2 ? lane : lane - 28)) * 6;
#pragma unroll
                for (int k = 0; k < 3; k++) { slot[k] = fp[k]; slot[3 + k] = fm[k]; }
            }
            __syncthreads();
            const float *lo_w = box + ((warp > 0 ? warp - 1 : 0) * 4) * 6;           // previous warp: slots 2,3 = its lanes 30,31
            const float *hi_w = box + ((warp < nwarps - 1 ? warp + 1 : warp) * 4) * 6;  // next warp: slots 0,1 = its lanes 0,1
            if constexpr (MODE == MODE_FAST && ORDER == 2) {
                // FAST 2nd order: face-flux form, operation for operation the arithmetic of step1d_acc.cuh (scalar twins of its
                // packed instructions), so the persistent and the streaming march give the same bits.
                const bool edge = at_lo || at_hi;
                const float q = edge ? 0.0f : P.quarter;  // wall cells are first order
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const float g = -fm[k];  // G = -F-
                    float fpL = __shfl_up_sync(full, fp[k], 1), gL = __shfl_up_sync(full, g, 1);      // cell c-1
                    float fpR = __shfl_down_sync(full, fp[k], 1), gR = __shfl_down_sync(full, g, 1);  // cell c+1
                    if (lane == 0) { fpL = lo_w[3 * 6 + k]; gL = -lo_w[3 * 6 + 3 + k]; }
                    if (lane == 31) { fpR = hi_w[0 * 6 + k]; gR = -hi_w[0 * 6 + 3 + k]; }
                    const float dp = fsub(fp[k], fpL), dg = fsub(g, gL);  // backward differences of this cell ...
                    const float rp = fsub(fpR, fp[k]), rg = fsub(gR, g);  // ... and of cell c+1 (its forward ones)
                    const float phi = face_flux_plus<LIM>(dp, rp, q, fp[k], P.alpha);
                    const float gam = face_flux_minus<LIM>(dg, rg, q, g, P.alpha);
                    float phiL = __shfl_up_sync(full, phi, 1);    // Phi+ of cell c-1
                    float gamR = __shfl_down_sync(full, gam, 1);  // Gamma of cell c+1
                    if (lane == 0) {   // face flux of the previous warp's lane 31, from its lanes 30, 31 and our lane 0
                        const float qn = ((c - 1 == 0) || (c - 1 == P.n - 1)) ? 0.0f : P.quarter;
                        phiL = face_flux_plus<LIM>(fsub(lo_w[3 * 6 + k], lo_w[2 * 6 + k]), dp, qn, lo_w[3 * 6 + k], P.alpha);
                    }
                    if (lane == 31) {  // face flux of the next warp's lane 0, from our lane 31 and its lanes 0, 1
                        const float qn = ((c + 1 == 0) || (c + 1 == P.n - 1)) ? 0.0f : P.quarter;
                        const float g1 = -hi_w[0 * 6 + 3 + k], g2 = -hi_w[1 * 6 + 3 + k];
                        gamR = face_flux_minus<LIM>(rg, fsub(g2, g1), qn, g1, P.alpha);
                    }
                    // wall ghosts: reflective (-,+,-) of the cell's own opposite flux (base_shll.c:95-97,108-110), outflow its own flux
                    const float ghost_lo = (BC == BC_REFLECT) ? ((k == 1) ? -g : g) : fp[k];
                    const float ghost_hi = (BC == BC_REFLECT) ? ((k == 1) ? -fp[k] : fp[k]) : g;
                    phiL = at_lo ? ghost_lo : phiL;
                    gamR = at_hi ? ghost_hi : gamR;
                    const float v = __fmaf_rn(-P.dtdx, fadd(fsub(phi, phiL), fsub(gam, gamR)), u[k]);
                    u[k] = in_domain ? v : 1.0f;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    float fpL = __shfl_up_sync(full, fp[k], 1);    // F+ of cell c-1
                    float fmR = __shfl_down_sync(full, fm[k], 1);  // F- of cell c+1
                    if (lane == 0) fpL = lo_w[3 * 6 + k];
                    if (lane == 31) fmR = hi_w[0 * 6 + 3 + k];
                    float left, right;
                    if (BC == BC_REFLECT) {
                        left = at_lo ? ((k == 1) ? fm[k] : -fm[k]) : fpL;
                        right = at_hi ? ((k == 1) ? fp[k] : -fp[k]) : fmR;
                    } else {
                        left = at_lo ? fp[k] : fpL;
                        right = at_hi ? fm[k] : fmR;
                    }
                    float v = apply_first<MODE>(u[k], P.dtdx, flux_sum<MODE>(fp[k], fm[k], right, left));
                    if (ORDER == 2) {
                        float fmL = __shfl_up_sync(full, fm[k], 1);
                        float fpR = __shfl_down_sync(full, fp[k], 1);
                        if (lane == 0) fmL = lo_w[3 * 6 + 3 + k];
                        if (lane == 31) fpR = hi_w[0 * 6 + k];
                        const bool edge = at_lo || at_hi;
                        float dfp = edge ? 0.0f : limited_slope<LIM>(fpL, fp[k], fpR, P.alpha);
                        float dfm = edge ? 0.0f : limited_slope<LIM>(fmL, fm[k], fmR, P.alpha);
                        float ldf = __shfl_up_sync(full, dfp, 1);    // dF+ of cell c-1
                        float rdf = __shfl_down_sync(full, dfm, 1);  // dF- of cell c+1
                        if (lane == 0) {   // slope of the previous warp's lane 31, from its lanes 30, 31 and our lane 0
                            const bool nb_edge = (c - 1 == 0) || (c - 1 == P.n - 1);
                            ldf = nb_edge ? 0.0f : limited_slope<LIM>(lo_w[2 * 6 + k], lo_w[3 * 6 + k], fp[k], P.alpha);
                        }
                        if (lane == 31) {  // slope of the next warp's lane 0, from our lane 31 and its lanes 0, 1
                            const bool nb_edge = (c + 1 == 0) || (c + 1 == P.n - 1);
                            rdf = nb_edge ? 0.0f : limited_slope<LIM>(fm[k], hi_w[0 * 6 + 3 + k], hi_w[1 * 6 + 3 + k], P.alpha);
                        }
                        ldf = at_lo ? 0.0f : ldf;
                        rdf = at_hi ? 0.0f : rdf;
                        v = apply_second<MODE, POW2>(v, P.half_dtdx, slope_sum(dfp, dfm, rdf, ldf));
                    }
                    u[k] = in_domain ? v : 1.0f;
                }
            }
        }

        // ---- publish the boundary strips of this round and raise the round counter
        if (r + 1 < nrounds) {
            const int par = (int)(r & 1);
            const size_t side_sz = (size_t)3 * P.hmax;
            float *mine = P.strips + ((size_t)par * P.nblocks + b) * 2 * side_sz;
            if (owned) {
                const int o = c - s0, no = s1 - s0;
                if (o < h) {
#pragma unroll
                    for (int k = 0; k < 3; k++) __stcg(mine + 0 * side_sz + k * P.hmax + o, u[k]);
                }
                if (o >= no - h) {
#pragma unroll
                    for (int k = 0; k < 3; k++) __stcg(mine + 1 * side_sz + k * P.hmax + (o - (no - h)), u[k]);
                }
            }
            __threadfence();
            __syncthreads();
            if (t == 0) st_release_gpu(P.round_done + b, P.round_base + (unsigned)r + 1u);
        }
    }
    if (owned) {
#pragma unroll
        for (int k = 0; k < 3; k++) P.out[k][c] = u[k];
    }
}

}  // namespace shll
