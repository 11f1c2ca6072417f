// Microbenchmark: does packing FP32 work into FFMA2/FADD2/FMUL2 free issue slots on sm_100a?
// Each thread runs 8 independent FMA chains plus 8 integer adds per iteration (a ~50/50 FP/INT mix like the stencil).
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>  // 0: scalar FP + int, 1: packed FP + int, 2: scalar FP only, 3: packed FP only
__global__ void k(float *out, int iters, float a, float b)
{
    float x[8]; int n[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 0.001f + i; n[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; it++) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = __fmaf_rn(x[i], a, b);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float2 r = __ffma2_rn(make_float2(x[i], x[i + 1]), make_float2(a, a), make_float2(b, b));
                x[i] = r.x; x[i + 1] = r.y;
            }
        }
        if (MODE < 2) {  // 4 integer adds per 8 FMAs: issue-bound when scalar (12 slots), FP-pipe-bound when packed (8 slots)
#pragma unroll
            for (int i = 0; i < 4; i++) n[i] = n[i] + n[i + 4] + it;
        }
    }
    float s = 0; int m = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { s += x[i]; m += n[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + m;
}
int main()
{
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    const char *names[4] = {"scalar FFMA + int", "packed FFMA2 + int", "scalar FFMA only", "packed FFMA2 only"};
    for (int mode = 0; mode < 4; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f);
            if (mode == 1) k<1><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f);
            if (mode == 2) k<2><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f);
            if (mode == 3) k<3><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = 148.0 * 8 * 256 * 8.0 * iters;
        printf("%-20s %8.3f ms  %.1f TFMA-lane/s\n", names[mode], ms, fma / ms * 1e-9);
    }
    return 0;
}
