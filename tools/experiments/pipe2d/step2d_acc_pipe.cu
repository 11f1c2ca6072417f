// step2d_acc_pipe.cu -- instantiations of the multi-step pipelined launch of the FAST 2D kernel (step2d_acc_pipe.cuh).
#include <cstring>

#include "shll_internal.h"

namespace shll {

// ---- the multi-step pipelined launch (step2d_acc_pipe.cuh): same instantiation table ------------------------------------
template <int ORDER, int BC, int LIM, int MINB, int STASH>
static cudaError_t go_pipe(const Step2DPipeParams &p, dim3 grid, size_t smem, cudaStream_t s, int *occ)
{
    if (occ != nullptr) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, step2d_acc_pipe_kernel<ORDER, BC, LIM, MINB, STASH>, 32, smem);
    step2d_acc_pipe_kernel<ORDER, BC, LIM, MINB, STASH><<<grid, 32, smem, s>>>(p);
    return cudaGetLastError();
}

template <int BC, int LIM>
static cudaError_t go_pipe_o2(int cfg, const Step2DPipeParams &p, dim3 grid, size_t smem, cudaStream_t s, int *occ)
{
    if (cfg == 1) return go_pipe<2, BC, LIM, 12, 0>(p, grid, smem, s, occ);  // the default register cap (step2d_acc.cu); the
    return cudaErrorInvalidValue;                                            // tuning variants keep one launch per step
}

static cudaError_t dispatch_pipe(const KernelKey &k, const Step2DPipeParams &p, dim3 grid, size_t smem, cudaStream_t s, int *occ)
{
    if (k.mode != MODE_FAST || k.vec != 2) return cudaErrorInvalidValue;
    if (k.order == 1) {
        if (k.bc == BC_REFLECT) return go_pipe<1, BC_REFLECT, LIM_MINMOD, 16, 0>(p, grid, smem, s, occ);
        if (k.bc == BC_OUTFLOW) return go_pipe<1, BC_OUTFLOW, LIM_MINMOD, 16, 0>(p, grid, smem, s, occ);
    } else {
        if (k.bc == BC_REFLECT && k.lim == LIM_MINMOD) return go_pipe_o2<BC_REFLECT, LIM_MINMOD>(k.acc_cfg, p, grid, smem, s, occ);
        if (k.bc == BC_REFLECT && k.lim == LIM_MC) return go_pipe_o2<BC_REFLECT, LIM_MC>(k.acc_cfg, p, grid, smem, s, occ);
        if (k.bc == BC_OUTFLOW && k.lim == LIM_MINMOD) return go_pipe_o2<BC_OUTFLOW, LIM_MINMOD>(k.acc_cfg, p, grid, smem, s, occ);
        if (k.bc == BC_OUTFLOW && k.lim == LIM_MC) return go_pipe_o2<BC_OUTFLOW, LIM_MC>(k.acc_cfg, p, grid, smem, s, occ);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_step2d_acc_pipe(const KernelKey &k, const Step2DPipeParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    return dispatch_pipe(k, p, grid, smem, s, nullptr);
}

cudaError_t step2d_acc_pipe_blocks_per_sm(const KernelKey &k, size_t smem, int *blocks)
{
    Step2DPipeParams none;
    memset(&none, 0, sizeof(none));
    return dispatch_pipe(k, none, dim3(1), smem, nullptr, blocks);
}

}  // namespace shll
