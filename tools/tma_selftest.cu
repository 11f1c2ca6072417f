// Minimal TMA probe: which forms of cp.async.bulk.tensor work on this box (debug aid, not part of the product).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, int z, float *out, int n, uint32_t bytes)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t bar = ring + 8192;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(ring), "l"(&map), "r"(x), "r"(y), "r"(bar) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(ring), "l"(&map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
    }
    __syncwarp();
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar) : "memory");
    const float *s = reinterpret_cast<const float *>(smem);
    for (int i = threadIdx.x; i < n; i += 32) out[i] = s[i];
}

int main()
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no entry point\n"); return 1; }
    const int ny = 64, nrows = 68, planes = 4;
    const size_t plane = (size_t)ny * nrows;
    float *d, *out;
    cudaMalloc(&d, plane * planes * 4);
    cudaMalloc(&out, 8192);
    float *h = new float[plane * planes];
    for (size_t i = 0; i < plane * planes; i++) h[i] = (float)i;
    cudaMemcpy(d, h, plane * planes * 4, cudaMemcpyHostToDevice);
    float ho[2048];
    struct Case { int rank; int x, y, z; cuuint32_t box[3]; const char *name; } cases[] = {
        {2, 0, 1, 0, {32, 3, 1}, "2d aligned"},
        {2, 28, 1, 0, {32, 3, 1}, "2d x=28"},
        {2, -4, 1, 0, {32, 3, 1}, "2d x=-4"},
        {2, 0, 1, 0, {36, 3, 1}, "2d box36 x=0"},
        {3, 0, 1, 0, {32, 3, 4}, "3d aligned"},
        {3, -4, 1, 0, {36, 3, 4}, "3d box36 x=-4"},
        {3, 28, 1, 0, {36, 3, 4}, "3d box36 x=28"},
        {3, 56, 66, 0, {36, 3, 4}, "3d box36 x=56 y=66 (clipped)"},
        {3, 0, -1, 0, {36, 4, 4}, "3d box36x4 y=-1"},
        {2, 2, 1, 0, {32, 3, 1}, "2d x=2"},
        {2, -1, 1, 0, {32, 3, 1}, "2d x=-1"},
    };
    for (auto &c : cases) {
        CUtensorMap m;
        cuuint64_t dims[3] = {ny, nrows, planes};
        cuuint64_t strides[2] = {ny * 4, plane * 4};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = ((encode_tiled_fn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, c.rank, d, dims, strides, c.box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
        uint32_t n = c.box[0] * c.box[1] * (c.rank == 3 ? c.box[2] : 1);
        cudaMemset(out, 0, 8192);
        if (c.rank == 2) probe<2><<<1, 32, 8192 + 64>>>(m, c.x, c.y, c.z, out, n, n * 4);
        else probe<3><<<1, 32, 8192 + 64>>>(m, c.x, c.y, c.z, out, n, n * 4);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); fflush(stdout); return 2; }
        cudaMemcpy(ho, out, n * 4, cudaMemcpyDeviceToHost);
        printf("%s: ok first=%g %g ... row1=%g last=%g\n", c.name, ho[0], ho[1], ho[32], ho[n - 1]);
    }
    return 0;
}
