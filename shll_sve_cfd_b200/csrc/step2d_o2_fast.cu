// step2d_o2_fast.cu -- instantiations of the fused 2D step kernels, second order, FAST arithmetic: LDG and TMA variants.
#include "shll_internal.h"

namespace shll {

template <int BC, int LIM, int VEC>
static cudaError_t go(const Step2DParams &p, dim3 grid, dim3 block, cudaStream_t s)
{
    step2d_kernel<2, BC, LIM, MODE_FAST, VEC, true><<<grid, block, 0, s>>>(p);
    return cudaGetLastError();
}
template <int BC, int LIM, int VEC>
static cudaError_t go_tma(const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    return launch_pdl(step2d_tma_kernel<2, BC, LIM, MODE_FAST, VEC, true>, grid, dim3(32), smem, s, p.pdl != 0, p);
}

template <int BC, int LIM>
static cudaError_t by_vec(const KernelKey &k, const Step2DParams &p, dim3 grid, dim3 block, cudaStream_t s)
{
    if (k.vec == 1) return go<BC, LIM, 1>(p, grid, block, s);
    if (k.vec == 2) return go<BC, LIM, 2>(p, grid, block, s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_step2d_o2_fast(const KernelKey &k, const Step2DParams &p, dim3 grid, dim3 block, cudaStream_t s)
{
    if (k.bc == BC_REFLECT && k.lim == LIM_MINMOD) return by_vec<BC_REFLECT, LIM_MINMOD>(k, p, grid, block, s);
    if (k.bc == BC_REFLECT && k.lim == LIM_MC) return by_vec<BC_REFLECT, LIM_MC>(k, p, grid, block, s);
    if (k.bc == BC_OUTFLOW && k.lim == LIM_MINMOD) return by_vec<BC_OUTFLOW, LIM_MINMOD>(k, p, grid, block, s);
    if (k.bc == BC_OUTFLOW && k.lim == LIM_MC) return by_vec<BC_OUTFLOW, LIM_MC>(k, p, grid, block, s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_step2d_tma_o2_fast(const KernelKey &k, const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    if (k.vec == 1) {
        if (k.bc == BC_REFLECT && k.lim == LIM_MINMOD) return go_tma<BC_REFLECT, LIM_MINMOD, 1>(p, grid, smem, s);
        if (k.bc == BC_REFLECT && k.lim == LIM_MC) return go_tma<BC_REFLECT, LIM_MC, 1>(p, grid, smem, s);
        if (k.bc == BC_OUTFLOW && k.lim == LIM_MINMOD) return go_tma<BC_OUTFLOW, LIM_MINMOD, 1>(p, grid, smem, s);
        if (k.bc == BC_OUTFLOW && k.lim == LIM_MC) return go_tma<BC_OUTFLOW, LIM_MC, 1>(p, grid, smem, s);
    } else if (k.vec == 2) {
        if (k.bc == BC_REFLECT && k.lim == LIM_MINMOD) return go_tma<BC_REFLECT, LIM_MINMOD, 2>(p, grid, smem, s);
        if (k.bc == BC_REFLECT && k.lim == LIM_MC) return go_tma<BC_REFLECT, LIM_MC, 2>(p, grid, smem, s);
        if (k.bc == BC_OUTFLOW && k.lim == LIM_MINMOD) return go_tma<BC_OUTFLOW, LIM_MINMOD, 2>(p, grid, smem, s);
        if (k.bc == BC_OUTFLOW && k.lim == LIM_MC) return go_tma<BC_OUTFLOW, LIM_MC, 2>(p, grid, smem, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace shll
