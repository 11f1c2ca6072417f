// step2d_o2_strict.cu -- instantiations of the fused 2D step kernels, second order, STRICT arithmetic
// (2nd_order_base_shll.c; MC limiter: base-omp/2nd_order_base_shll.c): LDG and TMA variants.
#include "shll_internal.h"

namespace shll {

template <int BC, int LIM, int VEC, bool POW2>
static cudaError_t go(const Step2DParams &p, dim3 grid, dim3 block, cudaStream_t s)
{
    step2d_kernel<2, BC, LIM, MODE_STRICT, VEC, POW2><<<grid, block, 0, s>>>(p);
    return cudaGetLastError();
}
template <int BC, int LIM, int VEC, bool POW2>
static cudaError_t go_tma(const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    return launch_pdl(step2d_tma_kernel<2, BC, LIM, MODE_STRICT, VEC, POW2>, grid, dim3(32), smem, s, p.pdl != 0, p);
}

template <int BC, int LIM>
static cudaError_t by_vec(const KernelKey &k, const Step2DParams &p, dim3 grid, dim3 block, cudaStream_t s)
{
    if (k.pow2) {
        if (k.vec == 1) return go<BC, LIM, 1, true>(p, grid, block, s);
        if (k.vec == 2) return go<BC, LIM, 2, true>(p, grid, block, s);
    } else {
        if (k.vec == 1) return go<BC, LIM, 1, false>(p, grid, block, s);
        if (k.vec == 2) return go<BC, LIM, 2, false>(p, grid, block, s);
    }
    return cudaErrorInvalidValue;
}
template <int BC, int LIM>
static cudaError_t by_vec_tma(const KernelKey &k, const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    if (k.vec == 1) return k.pow2 ? go_tma<BC, LIM, 1, true>(p, grid, smem, s) : go_tma<BC, LIM, 1, false>(p, grid, smem, s);
    if (k.vec == 2) return k.pow2 ? go_tma<BC, LIM, 2, true>(p, grid, smem, s) : go_tma<BC, LIM, 2, false>(p, grid, smem, s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_step2d_o2_strict(const KernelKey &k, const Step2DParams &p, dim3 grid, dim3 block, cudaStream_t s)
{
    if (k.bc == BC_REFLECT && k.lim == LIM_MINMOD) return by_vec<BC_REFLECT, LIM_MINMOD>(k, p, grid, block, s);
    if (k.bc == BC_REFLECT && k.lim == LIM_MC) return by_vec<BC_REFLECT, LIM_MC>(k, p, grid, block, s);
    if (k.bc == BC_OUTFLOW && k.lim == LIM_MINMOD) return by_vec<BC_OUTFLOW, LIM_MINMOD>(k, p, grid, block, s);
    if (k.bc == BC_OUTFLOW && k.lim == LIM_MC) return by_vec<BC_OUTFLOW, LIM_MC>(k, p, grid, block, s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_step2d_tma_o2_strict(const KernelKey &k, const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    if (k.bc == BC_REFLECT && k.lim == LIM_MINMOD) return by_vec_tma<BC_REFLECT, LIM_MINMOD>(k, p, grid, smem, s);
    if (k.bc == BC_REFLECT && k.lim == LIM_MC) return by_vec_tma<BC_REFLECT, LIM_MC>(k, p, grid, smem, s);
    if (k.bc == BC_OUTFLOW && k.lim == LIM_MINMOD) return by_vec_tma<BC_OUTFLOW, LIM_MINMOD>(k, p, grid, smem, s);
    if (k.bc == BC_OUTFLOW && k.lim == LIM_MC) return by_vec_tma<BC_OUTFLOW, LIM_MC>(k, p, grid, smem, s);
    return cudaErrorInvalidValue;
}

}  // namespace shll
