// step1d.cu -- instantiations of the fused 1D step kernel (base_shll.c; derived 1D 2nd-order program),
// plus the element-wise kernels: device-side Compute_P_from_U and the diagnostic CFL reduction.
#include "shll_internal.h"

namespace shll {

template <int ORDER, int BC, int LIM, int MODE, int TFORM>
static cudaError_t go(const KernelKey &k, const Step1DParams &p, dim3 grid, dim3 block, cudaStream_t s)
{
    if (MODE == MODE_STRICT && ORDER == 2 && !k.pow2)
        return launch_pdl(step1d_kernel<ORDER, BC, LIM, MODE, TFORM, false>, grid, block, 0, s, p.pdl != 0, p);
    return launch_pdl(step1d_kernel<ORDER, BC, LIM, MODE, TFORM, true>, grid, block, 0, s, p.pdl != 0, p);
}

template <int ORDER, int BC, int LIM>
static cudaError_t by_mode(const KernelKey &k, const Step1DParams &p, dim3 grid, dim3 block, cudaStream_t s, int nsub)
{
    if (nsub == 2) {  // two time steps per launch: FAST 2nd order only (step1d_acc.cuh: step1d_acc2_kernel)
        if (!(k.mode == MODE_FAST && ORDER == 2 && k.acc)) return cudaErrorInvalidValue;
        return launch_pdl(step1d_acc2_kernel<BC, LIM, 4>, grid, block, 0, s, p.pdl != 0, p);
    }
    if (k.mode == MODE_FAST && ORDER == 2 && k.acc) {  // face-flux form, packed cell pairs (step1d_acc.cuh)
        const dim3 g4 = grid, b4 = block;
        switch (k.acc_cfg) {
        case 0: return launch_pdl(step1d_acc_kernel<BC, LIM, 4>, g4, b4, 0, s, p.pdl != 0, p);
        case 2: return launch_pdl(step1d_acc_kernel<BC, LIM, 8>, g4, b4, 0, s, p.pdl != 0, p);
        default: return launch_pdl(step1d_acc_kernel<BC, LIM, 6>, g4, b4, 0, s, p.pdl != 0, p);
        }
    }
    if (k.mode == MODE_FAST) return go<ORDER, BC, LIM, MODE_FAST, TFORM_2D>(k, p, grid, block, s);  // FAST ignores tform
    if (k.tform == TFORM_1D) return go<ORDER, BC, LIM, MODE_STRICT, TFORM_1D>(k, p, grid, block, s);
    return go<ORDER, BC, LIM, MODE_STRICT, TFORM_2D>(k, p, grid, block, s);
}

cudaError_t launch_step1d(const KernelKey &k, const Step1DParams &p, dim3 grid, dim3 block, cudaStream_t s, int nsub)
{
    if (k.order == 1) {
        if (k.bc == BC_REFLECT) return by_mode<1, BC_REFLECT, LIM_MINMOD>(k, p, grid, block, s, nsub);
        return by_mode<1, BC_OUTFLOW, LIM_MINMOD>(k, p, grid, block, s, nsub);
    }
    if (k.bc == BC_REFLECT && k.lim == LIM_MINMOD) return by_mode<2, BC_REFLECT, LIM_MINMOD>(k, p, grid, block, s, nsub);
    if (k.bc == BC_REFLECT && k.lim == LIM_MC) return by_mode<2, BC_REFLECT, LIM_MC>(k, p, grid, block, s, nsub);
    if (k.bc == BC_OUTFLOW && k.lim == LIM_MINMOD) return by_mode<2, BC_OUTFLOW, LIM_MINMOD>(k, p, grid, block, s, nsub);
    if (k.bc == BC_OUTFLOW && k.lim == LIM_MC) return by_mode<2, BC_OUTFLOW, LIM_MC>(k, p, grid, block, s, nsub);
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------- persistent 1D march
template <int ORDER, int BC, int LIM, int MODE, int TFORM>
static cudaError_t go_persist(const KernelKey &k, const Persist1DParams &p, int nblocks, int threads, cudaStream_t s)
{
    void *args[] = {const_cast<Persist1DParams *>(&p)};
    const size_t smem = (size_t)2 * (threads / 32) * 24 * sizeof(float);
    const void *fn = (MODE == MODE_STRICT && ORDER == 2 && !k.pow2)
                         ? (const void *)persist1d_kernel<ORDER, BC, LIM, MODE, TFORM, false>
                         : (const void *)persist1d_kernel<ORDER, BC, LIM, MODE, TFORM, true>;
    return cudaLaunchCooperativeKernel(fn, dim3(nblocks), dim3(threads), args, smem, s);
}

template <int ORDER, int BC, int LIM>
static cudaError_t persist_by_mode(const KernelKey &k, const Persist1DParams &p, int nblocks, int threads, cudaStream_t s)
{
    if (k.mode == MODE_FAST) return go_persist<ORDER, BC, LIM, MODE_FAST, TFORM_2D>(k, p, nblocks, threads, s);
    if (k.tform == TFORM_1D) return go_persist<ORDER, BC, LIM, MODE_STRICT, TFORM_1D>(k, p, nblocks, threads, s);
    return go_persist<ORDER, BC, LIM, MODE_STRICT, TFORM_2D>(k, p, nblocks, threads, s);
}

cudaError_t launch_persist1d(const KernelKey &k, const Persist1DParams &p, int nblocks, int threads, cudaStream_t s)
{
    if (k.order == 1) {
        if (k.bc == BC_REFLECT) return persist_by_mode<1, BC_REFLECT, LIM_MINMOD>(k, p, nblocks, threads, s);
        return persist_by_mode<1, BC_OUTFLOW, LIM_MINMOD>(k, p, nblocks, threads, s);
    }
    if (k.bc == BC_REFLECT && k.lim == LIM_MINMOD) return persist_by_mode<2, BC_REFLECT, LIM_MINMOD>(k, p, nblocks, threads, s);
    if (k.bc == BC_REFLECT && k.lim == LIM_MC) return persist_by_mode<2, BC_REFLECT, LIM_MC>(k, p, nblocks, threads, s);
    if (k.bc == BC_OUTFLOW && k.lim == LIM_MINMOD) return persist_by_mode<2, BC_OUTFLOW, LIM_MINMOD>(k, p, nblocks, threads, s);
    if (k.bc == BC_OUTFLOW && k.lim == LIM_MC) return persist_by_mode<2, BC_OUTFLOW, LIM_MC>(k, p, nblocks, threads, s);
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------- Compute_P_from_U
struct PlanePtrs {
    const float *u[4];
    float *p[4];
    float *a;
};

template <int DIMS, int MODE, int TFORM>
__global__ void prim_kernel(PlanePtrs P, long n)
{
    for (long c = blockIdx.x * (long)blockDim.x + threadIdx.x; c < n; c += (long)gridDim.x * blockDim.x) {
        Prim q;
        if (DIMS == 1) {
            q = (MODE == MODE_STRICT) ? prim1d_strict<TFORM>(P.u[0][c], P.u[1][c], P.u[2][c]) : prim1d_fast(P.u[0][c], P.u[1][c], P.u[2][c]);
            P.p[0][c] = q.rho; P.p[1][c] = q.ux; P.p[2][c] = q.T;
        } else {
            q = (MODE == MODE_STRICT) ? prim2d_strict(P.u[0][c], P.u[1][c], P.u[2][c], P.u[3][c])
                                      : prim2d_fast(P.u[0][c], P.u[1][c], P.u[2][c], P.u[3][c]);
            P.p[0][c] = q.rho; P.p[1][c] = q.ux; P.p[2][c] = q.uy; P.p[3][c] = q.T;
        }
        if (P.a) P.a[c] = q.a;
    }
}

cudaError_t launch_prim(int dims, int mode, int tform, const float *const u[4], float *const p[4], float *a, long ncells,
                        cudaStream_t s)
{
    PlanePtrs P;
    for (int k = 0; k < 4; k++) { P.u[k] = u[k]; P.p[k] = p[k]; }
    P.a = a;
    const int block = 256;
    const int grid = (int)((ncells + block - 1) / block < 148 * 16 ? (ncells + block - 1) / block : 148 * 16);
    if (dims == 1) {
        if (mode == MODE_FAST) prim_kernel<1, MODE_FAST, TFORM_2D><<<grid, block, 0, s>>>(P, ncells);
        else if (tform == TFORM_1D) prim_kernel<1, MODE_STRICT, TFORM_1D><<<grid, block, 0, s>>>(P, ncells);
        else prim_kernel<1, MODE_STRICT, TFORM_2D><<<grid, block, 0, s>>>(P, ncells);
    } else {
        if (mode == MODE_FAST) prim_kernel<2, MODE_FAST, TFORM_2D><<<grid, block, 0, s>>>(P, ncells);
        else prim_kernel<2, MODE_STRICT, TFORM_2D><<<grid, block, 0, s>>>(P, ncells);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- CFL diagnostic
// max over cells of (|u| + a) * DT/DX.  The reference never computes this (base_shll.c:167 is a comment, DT is a
// constant); it is a monitor only.  Warp shuffle reduction, one atomicMax per block (values are >= 0, so the
// float bit patterns order like unsigned integers).
template <int DIMS, int TFORM>
__global__ void cfl_kernel(PlanePtrs P, long n, float dtdx, float dtdy, float *out)
{
    float m = 0.0f;
    for (long c = blockIdx.x * (long)blockDim.x + threadIdx.x; c < n; c += (long)gridDim.x * blockDim.x) {
        Prim q = (DIMS == 1) ? prim1d_strict<TFORM>(P.u[0][c], P.u[1][c], P.u[2][c])
                             : prim2d_strict(P.u[0][c], P.u[1][c], P.u[2][c], P.u[3][c]);
        float cx = (fabsf(q.ux) + q.a) * dtdx;
        float cy = (DIMS == 2) ? (fabsf(q.uy) + q.a) * dtdy : 0.0f;
        float v = fmaxf(cx, cy);
        m = (v > m) ? v : m;  // NaN-ignoring max
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    __shared__ float wmax[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wmax[w] = m;
    __syncthreads();
    if (w == 0) {
        m = (lane < (blockDim.x >> 5)) ? wmax[lane] : 0.0f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
        if (lane == 0) atomicMax(reinterpret_cast<unsigned *>(out), __float_as_uint(m));
    }
}

cudaError_t launch_max_cfl(int dims, int mode, int tform, const float *const u[4], long ncells, float dtdx, float dtdy,
                           float *out_dev, cudaStream_t s)
{
    (void)mode;
    PlanePtrs P;
    for (int k = 0; k < 4; k++) { P.u[k] = u[k]; P.p[k] = nullptr; }
    P.a = nullptr;
    const int block = 256;
    const int grid = (int)((ncells + block - 1) / block < 148 * 8 ? (ncells + block - 1) / block : 148 * 8);
    if (dims == 1) {
        if (tform == TFORM_1D) cfl_kernel<1, TFORM_1D><<<grid, block, 0, s>>>(P, ncells, dtdx, dtdy, out_dev);
        else cfl_kernel<1, TFORM_2D><<<grid, block, 0, s>>>(P, ncells, dtdx, dtdy, out_dev);
    } else {
        cfl_kernel<2, TFORM_2D><<<grid, block, 0, s>>>(P, ncells, dtdx, dtdy, out_dev);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- conservation sums
// Sum of every conserved component over the owned cells (mass, momentum, total energy per unit cell volume): the
// monitor SURVEY.md section 8(f) asks for next to the CFL number.  FP64 accumulation, fixed summation order (grid-stride
// per thread, shuffle tree per warp, warp order per block, block order on the host), so the result is reproducible
// run to run; it never feeds back into the march.
constexpr int SUM_BLOCKS = 148 * 4, SUM_THREADS = 256;

__global__ void __launch_bounds__(SUM_THREADS) sums_kernel(PlanePtrs P, int ncomp, long n, double *partial)
{
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (long c = blockIdx.x * (long)blockDim.x + threadIdx.x; c < n; c += (long)gridDim.x * blockDim.x)
        for (int k = 0; k < ncomp; k++) acc[k] += (double)P.u[k][c];
    __shared__ double wsum[SUM_THREADS / 32][4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], off);
        if (lane == 0) wsum[w][k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int i = 0; i < SUM_THREADS / 32; i++) t += wsum[i][threadIdx.x];
        partial[blockIdx.x * 4 + threadIdx.x] = t;
    }
}

int conserved_sums_blocks() { return SUM_BLOCKS; }

cudaError_t launch_conserved_sums(int ncomp, const float *const u[4], long ncells, double *partial_dev, cudaStream_t s)
{
    PlanePtrs P;
    for (int k = 0; k < 4; k++) { P.u[k] = u[k]; P.p[k] = nullptr; }
    P.a = nullptr;
    sums_kernel<<<SUM_BLOCKS, SUM_THREADS, 0, s>>>(P, ncomp, ncells, partial_dev);
    return cudaGetLastError();
}

}  // namespace shll
