// Microbenchmark: do the FMA pipe (scalar FFMA, packed FFMA2) and the ALU pipe (LOP3, FMNMX) of an sm_100a scheduler overlap?
// Each thread runs NF independent FP chains and NA independent ALU chains per iteration; 8 warps per scheduler, so latency is
// hidden and the time per iteration is a pipe / dispatch throughput.  Reported: scheduler cycles per warp per iteration (from
// the SM clock) next to the count of instructions of each kind, so that "max(FMA, ALU)" (pipes overlap), "FMA + ALU" (they do
// not) and "one instruction per cycle" (issue bound) can be told apart.
#include <cuda_runtime.h>
#include <cstdio>

template <int KIND>  // FP kind: 0 none, 1 scalar FFMA, 2 packed FFMA2; ALU kind in bits 4..: 0 none, 1 LOP3, 2 FMNMX
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b, unsigned mask, long long *cycles)
{
    constexpr int FP = KIND & 15, ALU = KIND >> 4;
    float x[8];
    unsigned n[8];
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 0.001f + i; n[i] = threadIdx.x * 7 + i; m[i] = threadIdx.x * 0.5f - i; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (FP == 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = __fmaf_rn(x[i], a, b);
        } else if (FP == 2) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float2 r = __ffma2_rn(make_float2(x[i], x[i + 1]), make_float2(a, a), make_float2(b, b));
                x[i] = r.x; x[i + 1] = r.y;
            }
        }
        if (ALU == 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) n[i] = (n[i] & mask) | (n[(i + 1) & 7] ^ 0x3e800000u);
        } else if (ALU == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) m[i] = fminf(fabsf(m[i]), fabsf(m[(i + 3) & 7] ));
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + __uint_as_float(n[i] & 0x3fffffffu) + m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int KIND>
void run(const char *name, float *out, long long *cyc, int nfp, int nalu)
{
    const int iters = 40000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND><<<148 * 4, 256>>>(out, iters, 0.999f, 0.001f, 0x7fffffffu, cyc);   // 4 blocks x 8 warps per SM = 8 warps per scheduler
    cudaEventRecord(e0);
    k<KIND><<<148 * 4, 256>>>(out, iters, 0.999f, 0.001f, 0x7fffffffu, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    // every scheduler runs 8 warps: elapsed cycles / iterations / 8 = scheduler cycles per warp-iteration (at the nominal SM clock)
    const double per = (double)ms * 1e-3 * khz * 1e3 / iters / 8.0;
    printf("%-28s %2d FP + %2d ALU warp-instr / iteration : %6.2f scheduler cycles per warp-iteration (%.3f ms)\n", name, nfp, nalu, per, ms);
}

int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 4 * 256 * 4); cudaMalloc(&cyc, 8);
    run<0x01>("8 FFMA", out, cyc, 8, 0);
    run<0x02>("4 FFMA2", out, cyc, 4, 0);
    run<0x10>("8 LOP3", out, cyc, 0, 8);
    run<0x20>("8 FMNMX", out, cyc, 0, 8);
    run<0x11>("8 FFMA + 8 LOP3", out, cyc, 8, 8);
    run<0x12>("4 FFMA2 + 8 LOP3", out, cyc, 4, 8);
    run<0x21>("8 FFMA + 8 FMNMX", out, cyc, 8, 8);
    run<0x22>("4 FFMA2 + 8 FMNMX", out, cyc, 4, 8);
    return 0;
}
