// step2d_acc.cu -- instantiations of the FAST-mode face-flux accumulate kernel (step2d_acc.cuh), 1st order: the one-step kernel
// and the two-steps-per-launch kernel.  (2nd order: step2d_acc_o2.cu -- separate translation units compile in parallel.)
#include "shll_internal.h"

namespace shll {

cudaError_t launch_step2d_acc_o2(const KernelKey &k, const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s);

// two 1st-order steps per launch (step2d_acc.cuh: step2d_acc2_kernel)
cudaError_t launch_step2d_acc2(const KernelKey &k, const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    if (k.mode != MODE_FAST || k.vec != 2 || k.order != 1) return cudaErrorInvalidValue;
    if (k.bc == BC_REFLECT) return launch_pdl(step2d_acc2_kernel<BC_REFLECT, 12>, grid, dim3(32), smem, s, p.pdl != 0, p);
    return launch_pdl(step2d_acc2_kernel<BC_OUTFLOW, 12>, grid, dim3(32), smem, s, p.pdl != 0, p);
}

cudaError_t launch_step2d_acc(const KernelKey &k, const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    if (k.mode != MODE_FAST || k.vec != 2) return cudaErrorInvalidValue;
    if (k.order == 2) return launch_step2d_acc_o2(k, p, grid, smem, s);
    // 16 resident warps per SM (126 registers).  Caps of 18 / 20 warps (112 / 96 registers) were measured slower: 178.7 / 174.0 vs
    // 185.9 Gcu/s at 4096^2 (profiles/r02_tma_store_and_occupancy.log).
    if (k.bc == BC_REFLECT) return launch_pdl(step2d_acc_kernel<1, BC_REFLECT, LIM_MINMOD, 16, 0>, grid, dim3(32), smem, s, p.pdl != 0, p);
    if (k.bc == BC_OUTFLOW) return launch_pdl(step2d_acc_kernel<1, BC_OUTFLOW, LIM_MINMOD, 16, 0>, grid, dim3(32), smem, s, p.pdl != 0, p);
    return cudaErrorInvalidValue;
}

}  // namespace shll
