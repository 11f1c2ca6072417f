// step2d_acc_o2.cu -- instantiations of the FAST-mode face-flux accumulate kernel (step2d_acc.cuh), 2nd order.
#include "shll_internal.h"

namespace shll {

// 12 resident warps per SM (168 registers, nothing spilled).  Every other register cap was measured slower on B200
// (profiles/r02_sweep_chunks_and_register_caps.log: 8 warps 105.2, 10 warps 112.9, 14 / 16 warps 97.0 with spills, vs 116.6 Gcu/s),
// and so were the shared-memory stash variants of round 1 (STASH = 1, 2 in step2d_acc.cuh: 14 - 16 warps, slower than 12 with
// everything in registers); only the winner is instantiated.
template <int BC, int LIM>
static cudaError_t go(const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    return launch_pdl(step2d_acc_kernel<2, BC, LIM, 12, 0>, grid, dim3(32), smem, s, p.pdl != 0, p);
}

cudaError_t launch_step2d_acc_o2(const KernelKey &k, const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    if (k.bc == BC_REFLECT && k.lim == LIM_MINMOD) return go<BC_REFLECT, LIM_MINMOD>(p, grid, smem, s);
    if (k.bc == BC_REFLECT && k.lim == LIM_MC) return go<BC_REFLECT, LIM_MC>(p, grid, smem, s);
    if (k.bc == BC_OUTFLOW && k.lim == LIM_MINMOD) return go<BC_OUTFLOW, LIM_MINMOD>(p, grid, smem, s);
    if (k.bc == BC_OUTFLOW && k.lim == LIM_MC) return go<BC_OUTFLOW, LIM_MC>(p, grid, smem, s);
    return cudaErrorInvalidValue;
}

}  // namespace shll
