// Shared between bdof.cu (plan, C ABI) and line_inst.cu (one instantiation unit per FFT length).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace bdof {

struct LineParams {
    const float2* in;        // field in
    float2* out;             // field out (may alias in)
    const float2* h;         // frequency-domain multiplier (natural FFT order, 1/N folded in)
    const float2* tw;        // stage twiddles (LineCfg layout), forward sign
    const float2* db;        // (delta, beta) slice for PRE_TRANSMIT / POST_ADJ (row mode)
    float2* grad;            // POST_ADJ: (dL/ddelta, dL/dbeta) out (may alias db)
    const float2* psi;       // POST_ADJ: stored psi entering the slice
    long long batch_stride;  // elements between batch items in the field
    long long db_batch_stride;
    int lines_per_batch;     // rows per batch item (row mode) or columns (col mode)
    int elem_stride;         // 1 (row) or nx (col)
    int line_stride;         // nx (row) or 1 (col)
    int in_shift;            // circular shift applied to the load index (ifftshift), MODE_INV
    int out_shift;           // circular shift applied to the store index (fftshift), MODE_FWD
    float k_dz;              // 2 pi dz / lambda
    long long* dbg;          // phase-timing buffer (only read when built with -DBDOF_PHASE_TIMING)
    int dbg_flags;           // ablation switches, instrumented builds only (BDOF_DBG_FLAGS): 1 no stream wait, 2 no stores, 4 no transmission math
    // L2 prefetch of what the NEXT kernel streams from DRAM (issued by this kernel's CTAs at start)
    const void* pf0;         // nullable
    const void* pf1;         // nullable
    long long pf_bytes;      // bytes per prefetched array (multiple of 16)
    // Row passes reading their lines straight out of a larger pitched buffer (tiling: the windows of a block, fused cut):
    // item b's first line starts in_offsets[b] elements into `in` and consecutive lines are in_line_stride elements apart.
    // Null / 0 = the regular layout (b * batch_stride, line_stride).  The output and the side inputs keep the regular layout.
    const long long* in_offsets;
    int in_line_stride;
    int tune;                // runtime switches (BDOF_TUNE): 1 L2-prefetch the next tile's main input at tile start,
                             // 2 L2-prefetch the next line's delta/beta (forward row pass)
};

enum Variant { V_ROW_CONV_T = 0, V_ROW_CONV, V_ROW_CONV_ADJ, V_ROW_FWD, V_ROW_INV, V_COL_CONV, V_COL_FWD, V_COL_INV, V_COL_CONV2D,
               V_COL_CONV_PIPE, V_SWEEP_FWD, V_SWEEP_ADJ, V_RESIDENT_FWD, V_RESIDENT_ADJ, V_COUNT };

// Pipelined passes (pipefft.cuh): number of parts of the stage exchange per FFT length, 0 = not available.
// Lengths with T == R1 (4096) use the cyclic-shift scheme and a twiddle table that includes the all-ones row 0.
constexpr int pipe_parts(int n) { return n == 2048 ? 2 : (n == 4096 ? 4 : (n <= 1024 ? 1 : 0)); }
constexpr bool pipe_shift(int n) { return n == 4096; }

// sin/cos for |x| up to ~1e4 rad: 3-term Cody-Waite reduction by pi/2 and the cephes single-precision
// minimax polynomials (max error ~1 ulp on the reduced range).  Same arithmetic as the fast path of
// sincosf(), without its Payne-Hanek slow path (which drags a local-memory frame into every kernel);
// k*delta per slice is O(1e-4..1) rad for every physical configuration.
__device__ __forceinline__ void sincos_fast(float x, float* s, float* c) {
    const float q = rintf(x * 0.636619772367581343f);
    float r = fmaf(q, -1.57079601287841796875f, x);
    r = fmaf(q, -3.1391647326017846353e-7f, r);
    r = fmaf(q, -5.3903025299577647655e-15f, r);
    const int iq = __float2int_rn(q);
    const float r2 = r * r;
    float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fmaf(sp, r2, -1.6666654611e-1f);
    sp = fmaf(sp * r2, r, r);
    float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fmaf(cp, r2, 4.166664568298827e-2f);
    cp = fmaf(cp, r2, -0.5f);
    cp = fmaf(cp, r2, 1.0f);
    float ss = (iq & 1) ? cp : sp;
    float cc = (iq & 1) ? sp : cp;
    if (iq & 2) ss = -ss;
    if ((iq + 1) & 2) cc = -cc;
    *s = ss;
    *c = cc;
}

// Transmission of one slice, t = exp(i k delta) * exp(-k beta) (npfuncs.py:38), returned as tau = t - 1.
// Why t - 1: for X-ray objects k*delta per slice is ~1e-4, so Re(t) = 1 - 1e-8: fp32 drops the cos term of EVERY slice
// (the same sign each time) and the field's modulus grows by x^2/2 per slice -- a bias that compounds linearly with depth
// (measured: half of the 6.8e-6 intensity error after 512 slices of the config-2 phantom).  tau keeps the small terms, and
// psi * t is formed as psi + psi * tau with FMAs (cmul1p), at the same instruction count as a complex multiply.
// General path: 3-term Cody-Waite reduction + cephes polynomials (sincos_fast), tau = t - 1 directly.
__device__ __forceinline__ float2 transmission_m1(float2 db, float k) {
    float s, c;
    sincos_fast(k * db.x, &s, &c);
    float m = expf(-k * db.y);
    return make_float2(fmaf(m, c, -1.0f), m * s);
}

// Same function for |k delta| <= pi/4 and |k beta| <= 0.5 (every X-ray configuration: k*delta per slice
// is O(1e-4..1e-1)): no range reduction, no quadrant selects; cos - 1 and exp - 1 by polynomials without the leading 1
// (exp: degree-7 Taylor, truncation 0.5^8/8! = 1e-7).  Callers pick it with a warp-uniform vote.
__device__ __forceinline__ float2 transmission_small_m1(float2 db, float k) {
    const float x = k * db.x, y = -k * db.y;
    const float x2 = x * x;
    float sp = fmaf(x2, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fmaf(sp, x2, -1.6666654611e-1f);
    sp = fmaf(sp * x2, x, x);
    float cp = fmaf(x2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fmaf(cp, x2, 4.166664568298827e-2f);
    cp = fmaf(cp, x2, -0.5f);
    const float cm1 = cp * x2;                       // cos x - 1
    float m = fmaf(y, 1.98412698e-4f, 1.38888889e-3f);
    m = fmaf(m, y, 8.33333333e-3f);
    m = fmaf(m, y, 4.16666667e-2f);
    m = fmaf(m, y, 1.66666667e-1f);
    m = fmaf(m, y, 0.5f);
    m = fmaf(m, y, 1.0f);
    const float em1 = m * y;                         // exp y - 1
    return make_float2(fmaf(em1, cm1, em1 + cm1), fmaf(em1, sp, sp));
}
// |k delta| <= 2^-4 and |k beta| <= 2^-6 (one thin slice of any X-ray object): truncation errors
// x^7/5040 = 7e-13 (sin), x^6/720 = 8e-11 (cos), y^4/24 = 2.5e-9 (exp) are below fp32 resolution
__device__ __forceinline__ float2 transmission_tiny_m1(float2 db, float k) {
    const float x = k * db.x, y = -k * db.y;
    const float x2 = x * x;
    const float sp = fmaf(fmaf(x2, 8.33333333e-3f, -1.66666667e-1f) * x2, x, x);
    const float cm1 = fmaf(x2, 4.16666667e-2f, -0.5f) * x2;
    const float em1 = fmaf(fmaf(y, 1.66666667e-1f, 0.5f), y, 1.0f) * y;
    return make_float2(fmaf(em1, cm1, em1 + cm1), fmaf(em1, sp, sp));
}
__device__ __forceinline__ bool transmission_is_tiny(float2 db, float k) {
    return fabsf(k * db.x) <= 0.0625f && fabsf(k * db.y) <= 0.015625f;
}
__device__ __forceinline__ bool transmission_is_small(float2 db, float k) {
    return fabsf(k * db.x) <= 0.78539816f && fabsf(k * db.y) <= 0.5f;
}
// per-element tier selection (kernels off the hot path)
__device__ __forceinline__ float2 transmission_any_m1(float2 db, float k) {
    if (transmission_is_tiny(db, k)) return transmission_tiny_m1(db, k);
    if (transmission_is_small(db, k)) return transmission_small_m1(db, k);
    return transmission_m1(db, k);
}

}  // namespace bdof

int bdof_fail(int code, const char* fmt, ...);
// TMA descriptor of a row-major complex64 matrix [rows][cols] for boxes of box_cols x box_rows elements
// (cuTensorMapEncodeTiled through the runtime's driver entry point; no link against libcuda)
int bdof_make_tensor_map(CUtensorMap* out, const void* base, long long rows, long long cols, int box_cols, int box_rows);
int bdof_launch_check(const char* what);
int bdof_sm_reserve();    // SMs the persistent line kernels leave free (bdof_set_sm_reserve / BDOF_SM_RESERVE)
bool bdof_use_pdl();      // programmatic dependent launch of the line kernels (BDOF_PDL=0 disables)

namespace bdof { struct SweepParams; struct ResidentParams; }
// resident small-field kernels (residentfft.cuh, resident_inst.cu): one CTA per field through all slices
int bdof_resident_supported(int n);
int bdof_launch_resident(int n, int adj, const bdof::ResidentParams& p, cudaStream_t st);
// cluster-resident kernels (clusterfft.cuh): one cluster of 8 CTAs per 256 x 256 field; same parameters (BDOF_CLUSTER=0 disables)
int bdof_cluster_supported(int n);
int bdof_launch_cluster(int n, int adj, const bdof::ResidentParams& p, cudaStream_t st);
// sweep kernels (sweepfft.cuh): col = 0 x kernel (rows), 1 y kernel (columns); adj = 0 forward, 1 adjoint;
// the field is [rows][cols] complex64 row-major with rows = batch * ny
#define BDOF_DECL_LINE(N) int bdof_launch_line_##N(int variant, const bdof::LineParams& p, long long n_lines, cudaStream_t st); \
    int bdof_launch_sweep_##N(int col, int adj, const bdof::SweepParams& p, long long rows, int cols, cudaStream_t st);
BDOF_DECL_LINE(64) BDOF_DECL_LINE(128) BDOF_DECL_LINE(256) BDOF_DECL_LINE(512)
BDOF_DECL_LINE(1024) BDOF_DECL_LINE(2048) BDOF_DECL_LINE(4096) BDOF_DECL_LINE(8192)

// mixed-radix passes for the other lengths (genericfft.cu): 2^a 3^b 5^c 7^d <= 2048, p.tw = full table W_N^k
int bdof_generic_supported(int n);
int bdof_launch_line_generic(int n, int variant, const bdof::LineParams& p, long long n_lines, cudaStream_t st);

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess && (cudaGetLastError(), true))   /* clear the sticky last-error */ \
            return bdof_fail(int(_e), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define BDOF_TRY(expr)            \
    do {                          \
        int _r = (expr);          \
        if (_r != 0) return _r;   \
    } while (0)
