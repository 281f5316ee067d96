// build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -o tools/tma_probe tools/tma_probe.cu -lcuda
// Developer probe: which 3-D TMA box shapes load correctly on this GPU (float32 tensor [nz][ny][2nx]).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int c0, int c1, int c2, unsigned bytes) {
    extern __shared__ __align__(128) float buf[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                         smem_u32(buf)), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (unsigned i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = buf[i];
}
int main() {
    const int nx = 64, ny = 8, nz = 64;
    std::vector<float> h((size_t)nz * ny * nx * 2);
    for (size_t i = 0; i < h.size(); ++i) h[i] = float(i);
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 1 << 20);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)f;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int boxes[][3] = {{96, 1, 48}};
    for (auto& b : boxes) {
        alignas(64) CUtensorMap tm;
        const cuuint64_t gdim[3] = {2 * nx, ny, nz};
        const cuuint64_t gstr[2] = {nx * 8, (cuuint64_t)ny * nx * 8};
        const cuuint32_t box[3] = {(cuuint32_t)b[0], (cuuint32_t)b[1], (cuuint32_t)b[2]}, es[3] = {1, 1, 1};
        CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const unsigned bytes = b[0] * b[1] * b[2] * 4;
        printf("box {%d,%d,%d} encode %d bytes %u: ", b[0], b[1], b[2], (int)r, bytes);
        if (r) { printf("\n"); continue; }
        const int coords[][3] = {{8, 3, 5}, {6, 3, 5}, {8, 3, 40}, {60, 3, 5}, {62, 7, 40}, {5, 3, 5}};
        for (auto& c : coords) {
            k<<<1, 256, bytes>>>(tm, o, c[0], c[1], c[2], bytes);
            cudaError_t e = cudaDeviceSynchronize();
            if (e) { printf("coords (%d,%d,%d) RUN ERROR %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 1; }
            std::vector<float> got(bytes / 4);
            cudaMemcpy(got.data(), o, bytes, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int z = 0; z < b[2]; ++z) for (int y = 0; y < b[1]; ++y) for (int x = 0; x < b[0]; ++x) {
                const int gz = c[2] + z, gy = c[1] + y, gx = c[0] + x;
                const float want = (gz < nz && gy < ny && gx < 2 * nx) ? h[((size_t)gz * ny + gy) * nx * 2 + gx] : 0.f;
                if (got[((size_t)z * b[1] + y) * b[0] + x] != want) ++bad;
            }
            printf("(%d,%d,%d) mismatches %d; ", c[0], c[1], c[2], bad);
        }
        printf("\n");
    }
    return 0;
}
