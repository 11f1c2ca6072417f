// shll_internal.h -- launcher interface between the C-ABI layer (shll_capi.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>

#include "persist1d.cuh"
#include "step1d.cuh"
#include "step1d_acc.cuh"
#include "step2d.cuh"
#include "step2d_tma.cuh"
#include "step2d_acc.cuh"

namespace shll {

struct KernelKey {
    int order, bc, lim, mode, vec, tform;
    bool pow2;
    bool tma;  // 2D only: TMA-fed kernel (needs ny % 4 == 0)
    bool acc;  // FAST: face-flux kernels (2D: TMA + 2 cells per lane, step2d_acc.cuh; 1D order 2: step1d_acc.cuh)
    int acc_cfg;  // its register cap / stash variant (step2d_acc.cu)
};

// <<<grid, block, smem, s>>> with or without the programmatic-stream-serialization attribute (halo_sync.cuh:
// pdl_wait_for_previous_step): the launch latency and the block dispatch of step n+1 overlap the tail of step n.
template <class Params>
inline cudaError_t launch_pdl(void (*kernel)(Params), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, const Params &p)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, p);
}

// Each returns cudaSuccess / the launch error; cudaErrorInvalidValue for a combination that was not instantiated.
cudaError_t launch_step2d_o1(const KernelKey &k, const Step2DParams &p, dim3 grid, dim3 block, cudaStream_t s);
cudaError_t launch_step2d_o2_strict(const KernelKey &k, const Step2DParams &p, dim3 grid, dim3 block, cudaStream_t s);
cudaError_t launch_step2d_o2_fast(const KernelKey &k, const Step2DParams &p, dim3 grid, dim3 block, cudaStream_t s);
cudaError_t launch_step2d_tma_o1(const KernelKey &k, const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s);
cudaError_t launch_step2d_tma_o2_strict(const KernelKey &k, const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s);
cudaError_t launch_step2d_tma_o2_fast(const KernelKey &k, const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s);
cudaError_t launch_step2d_acc(const KernelKey &k, const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s);
cudaError_t launch_step2d_acc2(const KernelKey &k, const Step2DTmaParams &p, dim3 grid, size_t smem, cudaStream_t s);
cudaError_t launch_step1d(const KernelKey &k, const Step1DParams &p, dim3 grid, dim3 block, cudaStream_t s, int nsub = 1);

// Persistent register-resident 1D march (cooperative launch; grid = nblocks, one block per SM).
cudaError_t launch_persist1d(const KernelKey &k, const Persist1DParams &p, int nblocks, int threads, cudaStream_t s);

// Compute_P_from_U on the device (for shll_download_p) and the diagnostic CFL reduction.
cudaError_t launch_prim(int dims, int mode, int tform, const float *const u[4], float *const p[4], float *a, long ncells,
                        cudaStream_t s);
cudaError_t launch_max_cfl(int dims, int mode, int tform, const float *const u[4], long ncells, float dtdx, float dtdy,
                           float *out_dev, cudaStream_t s);

// Per-block FP64 partial sums of the conserved components (conserved_sums_blocks() x 4 doubles), fixed order.
int conserved_sums_blocks();
cudaError_t launch_conserved_sums(int ncomp, const float *const u[4], long ncells, double *partial_dev, cudaStream_t s);

}  // namespace shll
