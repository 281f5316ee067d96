// Kernel and launcher of the mixed-radix line passes (genericfft.cuh).
#include "../../include/bdof.h"
#include "common.h"
#include "genericfft.cuh"

using namespace bdof;

struct WarpExec {
    int lane;
    template <class F> __device__ __forceinline__ void operator()(F f) { f(lane, 32); __syncwarp(); }
};

// smem: A[lpc][n] and B[lpc][n]; one warp per line
__global__ void __launch_bounds__(256) generic_line_kernel(const GenArgs a, const long long n_tiles) {
    extern __shared__ __align__(16) float2 g_smem[];
    const int n = a.n;
    float2* A = g_smem;
    float2* B = g_smem + a.lpc * n;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int warp = tid >> 5;
    WarpExec ex{tid & 31};
    const bool in_a = gen_result_in_a(a);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        gen_load(a, tile, tid, nthreads, A);
        __syncthreads();
        const long long line = tile * a.lpc + warp;
        if (line < a.n_lines) gen_transform_line(a, line, A + warp * n, B + warp * n, ex);
        __syncthreads();
        gen_store(a, tile, tid, nthreads, in_a ? A : B);
        __syncthreads();
    }
}

int bdof_generic_supported(int n) {
    int radix[GEN_MAX_STAGES];
    return gen_factorize(n, radix) > 0 ? 1 : 0;
}

int bdof_launch_line_generic(int n, int variant, const LineParams& p, long long n_lines, cudaStream_t st) {
    GenArgs a{};
    a.p = p;
    a.n = n;
    a.n_lines = n_lines;
    a.n_stages = gen_factorize(n, a.radix);
    if (a.n_stages == 0) return bdof_fail(BDOF_E_UNSUPPORTED, "FFT length %d is not of the form 2^a 3^b 5^c 7^d <= %d", n, GEN_MAX_N);
    if (!gen_set_variant(a, variant)) return bdof_fail(BDOF_E_BADARG, "bad variant %d for the mixed-radix pass", variant);
    const int lpc = gen_lines_per_tile(n);
    a.lpc = lpc;
    const size_t smem = size_t(2) * lpc * n * sizeof(float2);
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(generic_line_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096 * int(sizeof(float2))));
        attr_set = true;
    }
    const long long n_tiles = (n_lines + lpc - 1) / lpc;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        if (n_sm <= 0) n_sm = 148;
    }
    const long long slots = (long long)n_sm * 8;
    const unsigned grid = unsigned(n_tiles < slots ? n_tiles : slots);
    generic_line_kernel<<<grid, 32 * lpc, smem, st>>>(a, n_tiles);
    return bdof_launch_check("generic_line_kernel");
}
