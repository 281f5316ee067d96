// Microbenchmark: issue throughput of scalar vs packed (f32x2) FP32 ops on sm_100a, and mixes with LDS.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32x2 fp32x2.cu ; run on one B200.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float x, float y){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ float2 upk(u64 r){ float2 c; asm("mov.b64 {%0,%1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(r)); return c; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c){ float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float add1(float a, float b){ float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

constexpr int CH = 8;       // independent chains per thread
constexpr int IT = 2048;    // loop iterations
template <int MODE>
__global__ void __launch_bounds__(1024) bench(float* out, float seed, long long* cyc) {
    extern __shared__ float2 sm[];
    float a[CH], b[CH]; u64 A[CH], B[CH];
    for (int i = 0; i < CH; ++i) { a[i] = seed + i + threadIdx.x; b[i] = seed * 0.5f + i; A[i] = pk(a[i], b[i]); B[i] = pk(b[i], a[i]); }
    const float c = seed * 1.0001f; const u64 C = pk(c, c);
    if (MODE >= 6) { for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float2(seed, i); }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < IT; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) { a[i] = fma1(a[i], c, b[i]); b[i] = fma1(b[i], c, a[i]); }                 // 2 FFMA (3 reg)
            if (MODE == 1) { a[i] = add1(a[i], b[i]); b[i] = add1(b[i], a[i]); }                       // 2 FADD
            if (MODE == 2) { A[i] = fma2(A[i], C, B[i]); }                                             // 1 FFMA2 (= 2 FMA)
            if (MODE == 3) { A[i] = add2(A[i], B[i]); }                                                // 1 FADD2
            if (MODE == 4) { A[i] = add2(A[i], B[i]); B[i] = add2(B[i], A[i]); }                       // 2 FADD2
            if (MODE == 5) { A[i] = fma2(A[i], C, B[i]); a[i] = fma1(a[i], c, b[i]); }                 // FFMA2 + FFMA
            if (MODE == 6) { A[i] = add2(A[i], B[i]); B[i] = add2(B[i], A[i]);                         // 2 FADD2 + 1 LDS.64 per 2
                             if ((i & 1) == 0) { float2 v = sm[(threadIdx.x + i * 256 + it) & 4095]; a[i] += v.x; } }
            if (MODE == 7) { a[i] = add1(a[i], b[i]); b[i] = add1(b[i], a[i]); a[i] = add1(a[i], c); b[i] = add1(b[i], c);   // 4 FADD + LDS per 2
                             if ((i & 1) == 0) { float2 v = sm[(threadIdx.x + i * 256 + it) & 4095]; a[i] += v.x; } }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < CH; ++i) { float2 u = upk(A[i]); float2 w = upk(B[i]); s += a[i] + b[i] + u.x + u.y + w.x + w.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, double ops_per_iter_chain, int threads) {
    float* out; long long* cyc; int nb = 148;
    cudaMalloc(&out, nb * 1024 * 4); cudaMalloc(&cyc, nb * 8);
    cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    bench<MODE><<<nb, threads, 32768>>>(out, 1.0f, cyc); cudaDeviceSynchronize();
    bench<MODE><<<nb, threads, 32768>>>(out, 1.0f, cyc); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nb; ++i) avg += h[i]; avg /= nb;
    double warp_instr = double(IT) * CH * ops_per_iter_chain * (threads / 32);
    printf("%-30s warps/SM %2d  cycles %8.0f  warp-instr/clk/SM %.2f  (per SMSP %.2f)\n", name, threads / 32, avg, warp_instr / avg, warp_instr / avg / 4);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int th : {128, 256, 512, 1024}) {
        run<0>("FFMA 3-reg (scalar)", 2, th);
        run<1>("FADD (scalar)", 2, th);
        run<2>("FFMA2 (packed)", 1, th);
        run<3>("FADD2 (packed) 1 chain", 1, th);
        run<4>("FADD2 (packed) 2/chain", 2, th);
        run<5>("FFMA2 + FFMA mix", 2, th);
        run<6>("2 FADD2 + 0.5 LDS.64", 2, th);
        run<7>("4 FADD + 0.5 LDS.64", 4, th);
    }
    return 0;
}
