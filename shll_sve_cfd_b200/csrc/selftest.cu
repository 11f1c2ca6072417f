// selftest.cu -- device self-test of the STRICT-mode exactness shortcuts (shll_math.cuh), exported as
// shll_selftest_exact_division().
//
// STRICT mode claims bit-exactness with the reference's IEEE divisions (base_shll.c:173-175, base_shll_2d.c:314-316:
// u1/u0, u2/u0, u3/u0, u/a in float; (...)/CV in double) while never executing nvcc's div.rn expansion on the hot path:
//   div_rn_shared / div_rn_spec  -- reciprocal shared over the numerators + remainder correction + range guard
//   div_by_cv / div_by_cv_spec   -- multiplication by RN53(1/CV) + Markstein correction + range guard
// Whole-run parity tests cover the operands a flow produces.  This kernel covers the operand SPACE: every thread draws
// operand pairs from a counter-based generator (uniform bit patterns, "physical" magnitudes, and the guard edges 2^+-60 /
// 2^+-40, denormals, +-0, +-inf, NaN) and compares each shortcut with __fdiv_rn / __ddiv_rn bit for bit (NaNs compare as
// NaNs).  For the *_spec variants a result only counts when the guard did not flag the cell (a flagged cell is recomputed
// with the IEEE operations by the caller); the number of flagged draws is reported so that "never flagged" cannot hide.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "../../include/shll_b200.h"
#include "shll_math.cuh"

namespace {

using namespace shll;

__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// float with a random mantissa / sign and an exponent near 2^e (e in [-149, 127])
__device__ __forceinline__ float float_near_pow2(int e, uint64_t r)
{
    const uint32_t mant = (uint32_t)(r & 0x7fffffu), sign = (uint32_t)((r >> 23) & 1u) << 31;
    int be = e + 127;
    if (be <= 0) return __uint_as_float(sign | ((mant | 0x800000u) >> min(1 - be, 24)));  // denormal
    if (be > 254) be = 254;
    // a few draws sit exactly on / one ulp around the power of two: the guard edges themselves
    const uint32_t sel = (uint32_t)(r >> 24) & 7u;
    const uint32_t m = sel == 0 ? 0u : (sel == 1 ? 1u : (sel == 2 ? 0x7fffffu : mant));
    return __uint_as_float(sign | ((uint32_t)be << 23) | m);
}

__device__ __forceinline__ bool same_f(float a, float b)
{
    return (__float_as_uint(a) == __float_as_uint(b)) || (a != a && b != b);
}
__device__ __forceinline__ bool same_d(double a, double b)
{
    return (__double_as_longlong(a) == __double_as_longlong(b)) || (a != a && b != b);
}

// counts: [0] div_rn_shared mismatches, [1] div_rn_spec mismatches (unflagged draws), [2] div_by_cv mismatches,
//         [3] div_by_cv_spec mismatches (unflagged), [4] float draws flagged by the spec guard, [5] double draws flagged,
//         [6] float draws taking div_rn_shared's slow path, [7] float pairs tested, [8] doubles tested
__global__ void __launch_bounds__(256) selftest_kernel(uint64_t seed, uint64_t per_thread, unsigned long long *counts)
{
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    unsigned long long c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    static const int edges[] = {-149, -140, -127, -126, -125, -100, -61, -60, -59, -41, -40, -39, -20, -1, 0, 1, 20, 39, 40, 41, 59, 60, 61, 100, 126, 127};
    for (uint64_t i = 0; i < per_thread; i++) {
        const uint64_t r0 = splitmix64(seed ^ (tid * 0x100000001B3ull + i)), r1 = splitmix64(r0), r2 = splitmix64(r1);
        float a, b;
        const uint32_t kind = (uint32_t)(r2 >> 60);
        if (kind < 6) {                 // uniform bit patterns: every exponent, denormals, infinities, NaNs
            a = __uint_as_float((uint32_t)r0);
            b = __uint_as_float((uint32_t)(r0 >> 32));
        } else if (kind < 11) {         // "physical": |a| in [2^-20, 2^20], b in [2^-10, 2^10] (densities, sound speeds)
            a = float_near_pow2((int)(r1 % 41) - 20, r0);
            b = fabsf(float_near_pow2((int)((r1 >> 8) % 21) - 10, r0 >> 32));
        } else if (kind < 15) {         // guard edges: operands and quotients around 2^+-60, 2^+-40, the denormal range
            const int ea = edges[(r1 >> 3) % 26], eb = edges[(r1 >> 11) % 26];
            a = float_near_pow2(ea, r0);
            b = float_near_pow2(eb, r0 >> 32);
        } else {                        // zero / signed-zero numerators (gas at rest), zero denominators
            a = (r1 & 1) ? 0.0f : -0.0f;
            b = ((r1 >> 1) & 15) == 0 ? ((r1 & 32) ? 0.0f : -0.0f) : float_near_pow2((int)((r1 >> 8) % 200) - 100, r0);
        }
        const float want = __fdiv_rn(a, b);
        {   // the per-division guard form (2D order-2 STRICT kernel, prim kernels)
            const Recip R = make_recip(b);
            const float q0 = __fmul_rn(a, R.r);
            const bool fast = R.ok && (fabsf(q0) >= 0x1p-40f) && (fabsf(q0) <= 0x1p40f);
            c[6] += !(fast || (a == 0.0f && R.ok));
            c[0] += !same_f(div_rn_shared(a, b, R), want);
        }
        {   // the one-guard-per-cell form
            bool bad = false;
            const float r = recip_spec(b, bad);
            const float q = div_rn_spec(a, b, r, bad);
            c[4] += bad;
            if (!bad) c[1] += !same_f(q, want);
        }
        c[7]++;
        // double: n / CV.  n = (double)e - 0.5*k style operands plus uniform bit patterns and the guard edges
        double n;
        if (kind < 5) n = __longlong_as_double((long long)r1);
        else if (kind < 12) n = (double)a - 0.5 * (double)b;
        else {
            const int e = (int)(r2 % 2098) - 1074;   // every binary64 exponent incl. denormals; the guard sits near -969 / 976
            n = scalbn(1.0 + (double)(r1 >> 12) * 0x1p-52, e) * ((r2 >> 59) & 1 ? -1.0 : 1.0);
            if (((r2 >> 40) & 63) == 0) n = (r1 & 1) ? 0.0 : -0.0;
        }
        const double wantd = __ddiv_rn(n, SHLL_CV_D);
        c[2] += !same_d(div_by_cv(n), wantd);
        {
            bool bad = false;
            const double q = div_by_cv_spec(n, bad);
            c[5] += bad;
            if (!bad) c[3] += !same_d(q, wantd);
        }
        c[8]++;
    }
#pragma unroll
    for (int k = 0; k < 9; k++) {
        unsigned long long v = c[k];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(counts + k, v);
    }
}

}  // namespace

extern "C" int shll_selftest_exact_division(int device, unsigned long long npairs, unsigned long long seed, unsigned long long counts[9])
{
    if (!counts) return SHLL_E_INVAL;
    if (cudaSetDevice(device) != cudaSuccess) { (void)cudaGetLastError(); return SHLL_E_CUDA; }
    unsigned long long *dev = nullptr;
    if (cudaMalloc(&dev, 9 * sizeof(unsigned long long)) != cudaSuccess) { (void)cudaGetLastError(); return SHLL_E_CUDA; }
    cudaMemset(dev, 0, 9 * sizeof(unsigned long long));
    const int blocks = 148 * 8, threads = 256;
    const uint64_t per_thread = (npairs + (uint64_t)blocks * threads - 1) / ((uint64_t)blocks * threads);
    selftest_kernel<<<blocks, threads>>>(seed, per_thread, dev);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(counts, dev, 9 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(dev);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return SHLL_E_CUDA; }
    return SHLL_OK;
}
