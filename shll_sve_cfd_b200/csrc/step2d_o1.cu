// step2d_o1.cu -- instantiations of the fused 2D step kernels, first order (base_shll_2d.c): LDG and TMA variants.
#include "shll_internal.h"

namespace shll {

template <int BC, int MODE, int VEC>
static cudaError_t go(const Step2DParams &p, dim3 grid, dim3 block, cudaStream_t s)
{
    step2d_kernel<1, BC, LIM_MINMOD, MODE, VEC, true><<<grid, block, 0, s>>>(p);
    return cudaGetLastError();
}
template <int BC, int MODE, int VEC>
static cudaError_t go_tma(const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    return launch_pdl(step2d_tma_kernel<1, BC, LIM_MINMOD, MODE, VEC, true>, grid, dim3(32), smem, s, p.pdl != 0, p);
}

template <int BC, int MODE>
static cudaError_t by_vec(int vec, const Step2DParams &p, dim3 grid, dim3 block, cudaStream_t s)
{
    switch (vec) {
    case 1: return go<BC, MODE, 1>(p, grid, block, s);
    case 2: return go<BC, MODE, 2>(p, grid, block, s);
    case 4: return go<BC, MODE, 4>(p, grid, block, s);
    }
    return cudaErrorInvalidValue;
}
template <int BC, int MODE>
static cudaError_t by_vec_tma(int vec, const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    switch (vec) {
    case 1: return go_tma<BC, MODE, 1>(p, grid, smem, s);
    case 2: return go_tma<BC, MODE, 2>(p, grid, smem, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_step2d_o1(const KernelKey &k, const Step2DParams &p, dim3 grid, dim3 block, cudaStream_t s)
{
    if (k.bc == BC_REFLECT && k.mode == MODE_STRICT) return by_vec<BC_REFLECT, MODE_STRICT>(k.vec, p, grid, block, s);
    if (k.bc == BC_REFLECT && k.mode == MODE_FAST) return by_vec<BC_REFLECT, MODE_FAST>(k.vec, p, grid, block, s);
    if (k.bc == BC_OUTFLOW && k.mode == MODE_STRICT) return by_vec<BC_OUTFLOW, MODE_STRICT>(k.vec, p, grid, block, s);
    if (k.bc == BC_OUTFLOW && k.mode == MODE_FAST) return by_vec<BC_OUTFLOW, MODE_FAST>(k.vec, p, grid, block, s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_step2d_tma_o1(const KernelKey &k, const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    if (k.bc == BC_REFLECT && k.mode == MODE_STRICT) return by_vec_tma<BC_REFLECT, MODE_STRICT>(k.vec, p, grid, smem, s);
    if (k.bc == BC_REFLECT && k.mode == MODE_FAST) return by_vec_tma<BC_REFLECT, MODE_FAST>(k.vec, p, grid, smem, s);
    if (k.bc == BC_OUTFLOW && k.mode == MODE_STRICT) return by_vec_tma<BC_OUTFLOW, MODE_STRICT>(k.vec, p, grid, smem, s);
    if (k.bc == BC_OUTFLOW && k.mode == MODE_FAST) return by_vec_tma<BC_OUTFLOW, MODE_FAST>(k.vec, p, grid, smem, s);
    return cudaErrorInvalidValue;
}

}  // namespace shll
