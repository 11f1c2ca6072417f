// step2d_acc.cu -- instantiations of the FAST-mode face-flux accumulate kernel (step2d_acc.cuh): both orders.
#include "shll_internal.h"

namespace shll {

template <int ORDER, int BC, int LIM, int MINB, int STASH>
static cudaError_t go(const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    return launch_pdl(step2d_acc_kernel<ORDER, BC, LIM, MINB, STASH>, grid, dim3(32), smem, s, p.pdl != 0, p);
}

// order 2: register cap (resident warps per SM) / stash level chosen by the host, KernelKey::acc_cfg
template <int BC, int LIM>
static cudaError_t go_o2(int cfg, const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    switch (cfg) {
    case 0: return go<2, BC, LIM, 8, 0>(p, grid, smem, s);
    case 1: return go<2, BC, LIM, 12, 0>(p, grid, smem, s);
    case 2: return go<2, BC, LIM, 14, 2>(p, grid, smem, s);
    case 3: return go<2, BC, LIM, 16, 2>(p, grid, smem, s);
    case 4: return go<2, BC, LIM, 14, 1>(p, grid, smem, s);
    case 5: return go<2, BC, LIM, 16, 0>(p, grid, smem, s);
    case 6: return go<2, BC, LIM, 14, 0>(p, grid, smem, s);
    case 7: return go<2, BC, LIM, 10, 0>(p, grid, smem, s);
    }
    return cudaErrorInvalidValue;
}

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
    if (k.order == 1) {
        if (k.acc_cfg == 5) {   // experiments: more resident warps per SM (register cap 112 / 96)
            if (k.bc == BC_REFLECT) return go<1, BC_REFLECT, LIM_MINMOD, 18, 0>(p, grid, smem, s);
            return go<1, BC_OUTFLOW, LIM_MINMOD, 18, 0>(p, grid, smem, s);
        }
        if (k.acc_cfg == 6) {
            if (k.bc == BC_REFLECT) return go<1, BC_REFLECT, LIM_MINMOD, 20, 0>(p, grid, smem, s);
            return go<1, BC_OUTFLOW, LIM_MINMOD, 20, 0>(p, grid, smem, s);
        }
        if (k.bc == BC_REFLECT) return go<1, BC_REFLECT, LIM_MINMOD, 16, 0>(p, grid, smem, s);
        if (k.bc == BC_OUTFLOW) return go<1, BC_OUTFLOW, LIM_MINMOD, 16, 0>(p, grid, smem, s);
    } else {
        if (k.bc == BC_REFLECT && k.lim == LIM_MINMOD) return go_o2<BC_REFLECT, LIM_MINMOD>(k.acc_cfg, p, grid, smem, s);
        if (k.bc == BC_REFLECT && k.lim == LIM_MC) return go_o2<BC_REFLECT, LIM_MC>(k.acc_cfg, p, grid, smem, s);
        if (k.bc == BC_OUTFLOW && k.lim == LIM_MINMOD) return go_o2<BC_OUTFLOW, LIM_MINMOD>(k.acc_cfg, p, grid, smem, s);
        if (k.bc == BC_OUTFLOW && k.lim == LIM_MC) return go_o2<BC_OUTFLOW, LIM_MC>(k.acc_cfg, p, grid, smem, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace shll
