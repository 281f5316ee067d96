// libbdof: B200-native Fresnel multislice engine -- plan, pass scheduling and the C ABI.
// See include/bdof.h for the boundary and DESIGN.md for the data layout.
#include "../../include/bdof.h"
#include "common.h"
#include "regfft.cuh"
#include "sweepfft.cuh"
#include "residentfft.cuh"

#include <atomic>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

using namespace bdof;

// ------------------------------------------------------------------------------------------
// errors / bookkeeping
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

int bdof_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int bdof_launch_check(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bdof_fail(int(e), "launch of %s failed: %s", what, cudaGetErrorString(e));
    return 0;
}
static int g_sm_reserve = -1;
int bdof_sm_reserve() {
    if (g_sm_reserve < 0) { const char* e = getenv("BDOF_SM_RESERVE"); g_sm_reserve = e ? atoi(e) : 0; }
    return g_sm_reserve;
}
extern "C" int bdof_set_sm_reserve(int n_sms) {
    if (n_sms < 0) return bdof_fail(BDOF_E_BADARG, "negative SM count");
    g_sm_reserve = n_sms;
    return 0;
}
bool bdof_use_pdl() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("BDOF_PDL"); v = (e && e[0] == '0') ? 0 : 1; }     // sweep kernels: +2..14 % (prologue overlaps the previous kernel's tail)
    return v == 1;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int bdof_make_tensor_map(CUtensorMap* out, const void* base, long long rows, long long cols, int box_cols, int box_rows) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
        if (!f || q != cudaDriverEntryPointSuccess) return bdof_fail(BDOF_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(f);
    }
    // complex64 as pairs of fp32: inner dimension 2*cols floats
    const cuuint64_t gdim[2] = {cuuint64_t(2 * cols), cuuint64_t(rows)};
    const cuuint64_t gstride[1] = {cuuint64_t(cols) * sizeof(float2)};
    const cuuint32_t box[2] = {cuuint32_t(2 * box_cols), cuuint32_t(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return bdof_fail(BDOF_E_BADARG, "cuTensorMapEncodeTiled failed (CUresult %d)", int(r));
    return 0;
}
// general form: fp32 tensor of `rank` dimensions, dims[0] innermost (contiguous), strides_bytes[i] = stride of dims[i + 1]
static int make_tensor_map_nd(CUtensorMap* out, const void* base, int rank, const long long* dims, const long long* strides_bytes, const int* box) {
    CUtensorMap probe;
    if (int r = bdof_make_tensor_map(&probe, base, 1, 64, 16, 1)) return r;        // resolves the driver entry point
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t bx[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = cuuint64_t(dims[i]); bx[i] = cuuint32_t(box[i]); estr[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstride[i] = cuuint64_t(strides_bytes[i]);
    CUresult r = reinterpret_cast<EncodeTiledFn>(f)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, cuuint32_t(rank), const_cast<void*>(base), gdim, gstride, bx, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return bdof_fail(BDOF_E_BADARG, "cuTensorMapEncodeTiled (rank %d) failed (CUresult %d)", rank, int(r));
    return 0;
}
#define fail bdof_fail
#define launch_check bdof_launch_check

extern "C" int bdof_version(void) { return 100; }
extern "C" const char* bdof_last_error(void) { return g_err; }
extern "C" unsigned long long bdof_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------
// line-kernel dispatch: one translation unit per FFT length (line_inst.cu, -DBDOF_N=...)
// ------------------------------------------------------------------------------------------
extern "C" int bdof_size_supported(int n) {
    // 1: power-of-two lengths with register-resident kernels (and sweep kernels up to 4096); 2: mixed-radix passes
    switch (n) { case 64: case 128: case 256: case 512: case 1024: case 2048: case 4096: case 8192: return 1; }
    return bdof_generic_supported(n) ? 2 : 0;
}
static std::vector<float2> make_full_twiddles(int n) {
    std::vector<float2> tw(n);
    for (int k = 0; k < n; ++k) {
        const double a = -2.0 * M_PI * double(k) / double(n);
        tw[k] = make_float2(float(cos(a)), float(sin(a)));
    }
    return tw;
}

static int launch_variant(int n, int variant, const LineParams& p, long long n_lines, cudaStream_t st) {
    switch (n) {
        case 64:   return bdof_launch_line_64(variant, p, n_lines, st);
        case 128:  return bdof_launch_line_128(variant, p, n_lines, st);
        case 256:  return bdof_launch_line_256(variant, p, n_lines, st);
        case 512:  return bdof_launch_line_512(variant, p, n_lines, st);
        case 1024: return bdof_launch_line_1024(variant, p, n_lines, st);
        case 2048: return bdof_launch_line_2048(variant, p, n_lines, st);
        case 4096: return bdof_launch_line_4096(variant, p, n_lines, st);
        case 8192: return bdof_launch_line_8192(variant, p, n_lines, st);
    }
    return bdof_fail(BDOF_E_UNSUPPORTED, "FFT length %d is not supported", n);
}

static int launch_sweep_n(int n, int col, int adj, const SweepParams& p, long long rows, int cols, cudaStream_t st) {
    switch (n) {
        case 64:   return bdof_launch_sweep_64(col, adj, p, rows, cols, st);
        case 128:  return bdof_launch_sweep_128(col, adj, p, rows, cols, st);
        case 256:  return bdof_launch_sweep_256(col, adj, p, rows, cols, st);
        case 512:  return bdof_launch_sweep_512(col, adj, p, rows, cols, st);
        case 1024: return bdof_launch_sweep_1024(col, adj, p, rows, cols, st);
        case 2048: return bdof_launch_sweep_2048(col, adj, p, rows, cols, st);
        case 4096: return bdof_launch_sweep_4096(col, adj, p, rows, cols, st);
    }
    return bdof_fail(BDOF_E_UNSUPPORTED, "no sweep kernel for FFT length %d", n);
}

struct StageRadices { int r1, r2, r3; };
static StageRadices radices_for(int n) {
    switch (n) {
        case 64: return {8, 8, 1};
        case 128: return {16, 8, 1};
        case 256: return {16, 16, 1};
        case 512: return {32, 16, 1};
        case 1024: return {32, 32, 1};
#if !defined(BDOF_ALT) || BDOF_ALT == 0
        case 2048: return {64, 32, 1};
#elif BDOF_ALT == 1
        case 2048: return {32, 32, 2};
#else
        case 2048: return {16, 16, 8};
#endif
        case 4096: return {64, 64, 1};
        case 8192: return {64, 64, 2};
    }
    return {0, 0, 0};
}

// stage twiddles, forward sign, LineCfg layout, computed in double
static std::vector<float2> make_twiddles(int n) {
    StageRadices r = radices_for(n);
    std::vector<float2> tw;
    auto add_stage = [&](int R, int NS) {
        const double M = double(NS) * R;
        for (int rr = 1; rr < R; ++rr)
            for (int k = 0; k < NS; ++k) {
                double a = -2.0 * M_PI * double((long long)rr * k % (long long)M) / M;
                tw.push_back(make_float2(float(cos(a)), float(sin(a))));
            }
    };
    add_stage(r.r2, r.r1);
    if (r.r3 > 1) add_stage(r.r3, r.r1 * r.r2);
    return tw;
}
// table of the pipelined passes: the same, except that cyclic-shift plans also carry row 0 (all ones)
static std::vector<float2> make_twiddles_pipe(int n) {
    if (!pipe_shift(n)) return make_twiddles(n);
    StageRadices r = radices_for(n);
    std::vector<float2> tw;
    for (int rr = 0; rr < r.r2; ++rr)
        for (int k = 0; k < r.r1; ++k) {
            double a = -2.0 * M_PI * double((long long)rr * k % n) / double(n);
            tw.push_back(make_float2(float(cos(a)), float(sin(a))));
        }
    return tw;
}

// ------------------------------------------------------------------------------------------
// elementwise kernels
// ------------------------------------------------------------------------------------------
__global__ void k_broadcast_probe(const float2* __restrict__ probe, float2* __restrict__ out, long long n_per, int batch) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_per) return;
    float2 v = probe[i];
    for (int b = 0; b < batch; ++b) out[b * n_per + i] = v;
}

// out = in * t(db)    (slice that modulates without propagating, npfuncs.py:38-40)
__global__ void k_modulate(const float2* __restrict__ in, const float2* __restrict__ db, float2* __restrict__ out,
                           long long n, float k) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = cmul1p(in[i], transmission_any_m1(db[i], k));
}

// adjoint of k_modulate: grad = -k (Im, Re)(conj(G) psi t), G <- conj(t) G
__global__ void k_modulate_adj(float2* __restrict__ G, const float2* __restrict__ psi, const float2* __restrict__ db,
                               float2* __restrict__ grad, long long n, float k) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 tr = transmission_any_m1(db[i], k);         // tau = t - 1
    const float2 u = cmul1p(psi[i], tr);
    const float2 g = G[i];
    const float2 w = cmulc(u, g);
    grad[i] = make_float2(-k * w.y, -k * w.x);
    G[i] = cmulc1p(g, tr);
}

// out = in * (re + i im)
__global__ void k_scale_complex(const float2* __restrict__ in, float2* __restrict__ out, long long n, float re, float im) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = cmul(in[i], make_float2(re, im));
}

__global__ void k_sum_batch(const float2* __restrict__ in, float2* __restrict__ out, long long n_per, int batch) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_per) return;
    float sx = 0.f, sy = 0.f;
    for (int b = 0; b < batch; ++b) { float2 v = in[b * n_per + i]; sx += v.x; sy += v.y; }
    out[i] = make_float2(sx, sy);
}

// loss head: partial sums of (|psi| - y)^2 per block (double), G = scale*(2/M)(|psi|-y) psi/|psi|
constexpr int LOSS_BLOCKS = 1184;   // 148 SMs x 8
constexpr int LOSS_THREADS = 256;
__global__ void __launch_bounds__(LOSS_THREADS) k_loss_mag(const float2* __restrict__ psi, const float* __restrict__ target,
                                                            float2* __restrict__ G, long long n, float gscale,
                                                            double* __restrict__ partial) {
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 v = psi[i];
        const float mag = sqrtf(v.x * v.x + v.y * v.y);
        const float d = mag - target[i];
        acc += double(d) * double(d);
        if (G != nullptr) {
            const float s = mag > 0.f ? gscale * d / mag : 0.f;
            G[i] = make_float2(s * v.x, s * v.y);
        }
    }
    __shared__ double red[LOSS_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) s += red[w];
        partial[blockIdx.x] = s;
    }
}
__global__ void k_loss_final(const double* __restrict__ partial, int n_partial, double scale, double* __restrict__ out) {
    // single warp, fixed order -> deterministic
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += 32) acc += partial[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) *out = acc * scale;
}

// [B,Y,X,Z] planes -> [Z,B,Y,X,2]: per (b,y) a [X][Z] -> [Z][X] tile transpose through smem
__global__ void k_pack_db(const float* __restrict__ delta, const float* __restrict__ beta, float2* __restrict__ db,
                          int n_rows /*B*Y*/, int nx, int nz) {
    __shared__ float td[32][33], tb[32][33];
    const int row = blockIdx.z;
    const int x0 = blockIdx.x * 32, z0 = blockIdx.y * 32;
    const long long in_base = (long long)row * nx * nz;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int x = x0 + i, z = z0 + threadIdx.x;
        if (x < nx && z < nz) {
            td[i][threadIdx.x] = delta[in_base + (long long)x * nz + z];
            tb[i][threadIdx.x] = beta[in_base + (long long)x * nz + z];
        }
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int z = z0 + i, x = x0 + threadIdx.x;
        if (x < nx && z < nz)
            db[((long long)z * n_rows + row) * nx + x] = make_float2(td[threadIdx.x][i], tb[threadIdx.x][i]);
    }
}
__global__ void k_unpack_db(const float2* __restrict__ db, float* __restrict__ delta, float* __restrict__ beta,
                            int n_rows, int nx, int nz) {
    __shared__ float td[32][33], tb[32][33];
    const int row = blockIdx.z;
    const int x0 = blockIdx.x * 32, z0 = blockIdx.y * 32;
    const long long out_base = (long long)row * nx * nz;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int z = z0 + i, x = x0 + threadIdx.x;
        if (x < nx && z < nz) {
            const float2 v = db[((long long)z * n_rows + row) * nx + x];
            td[i][threadIdx.x] = v.x;
            tb[i][threadIdx.x] = v.y;
        }
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int x = x0 + i, z = z0 + threadIdx.x;
        if (x < nx && z < nz) {
            delta[out_base + (long long)x * nz + z] = td[threadIdx.x][i];
            beta[out_base + (long long)x * nz + z] = tb[threadIdx.x][i];
        }
    }
}

// ptychography windows (ptychography.py:62-76): zero outside the object
__global__ void k_patch_gather(const float2* __restrict__ obj, int oy, int ox, const int* __restrict__ pos, int n_pos,
                               int py, int px, float2* __restrict__ patches) {
    // grid: (ceil(px*py/256), n_pos, n_slice)
    const int z = blockIdx.z, ip = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= py * px) return;
    const int yy = i / px, xx = i - yy * px;
    const int y = pos[2 * ip] + yy, x = pos[2 * ip + 1] + xx;
    float2 v = make_float2(0.f, 0.f);
    if (y >= 0 && y < oy && x >= 0 && x < ox) v = obj[((long long)z * oy + y) * ox + x];
    patches[(((long long)z * n_pos + ip) * py + yy) * px + xx] = v;
}
__global__ void k_patch_scatter_add(const float2* __restrict__ gpatch, int oy, int ox, const int* __restrict__ pos,
                                    int n_pos, int py, int px, float2* __restrict__ gobj) {
    const int z = blockIdx.z, ip = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= py * px) return;
    const int yy = i / px, xx = i - yy * px;
    const int y = pos[2 * ip] + yy, x = pos[2 * ip + 1] + xx;
    if (y >= 0 && y < oy && x >= 0 && x < ox) {
        const float2 v = gpatch[(((long long)z * n_pos + ip) * py + yy) * px + xx];
        float* dst = reinterpret_cast<float*>(gobj + ((long long)z * oy + y) * ox + x);
        atomicAdd(dst, v.x);
        atomicAdd(dst + 1, v.y);
    }
}

// The same accumulation as a GATHER, in scan-position order: deterministic (bit-identical run to run and for any split of the
// positions into launches that keeps their order), which the fp32 atomics above are not.  A block owns PSCAT_BX consecutive x of
// one object row y and PSCAT_ZB consecutive slices (thread = one x, one of PSCAT_ZL slice lanes); it first compacts, IN ORDER,
// the positions whose window touches its pixels (ballot prefix sums; origins kept in shared memory), then every thread adds the
// window pixels that fall on its own object pixel.  Narrow x ranges keep the share of listed-but-not-covering positions low (a
// 64-wide window over a 32-wide range: 2/3 of the listed positions cover a given pixel).
constexpr int PSCAT_BX = 32, PSCAT_ZL = 4, PSCAT_THREADS = PSCAT_BX * PSCAT_ZL, PSCAT_ZB = 16, PSCAT_MAX_POS = 4096;
__global__ void __launch_bounds__(PSCAT_THREADS) k_patch_gather_add(const float2* __restrict__ gpatch, int n_slice, int oy, int ox,
                                                                     const int* __restrict__ pos, int n_pos, int py, int px,
                                                                     float2* __restrict__ gobj) {
    __shared__ int s_idx[PSCAT_MAX_POS];
    __shared__ short s_wy[PSCAT_MAX_POS], s_wx[PSCAT_MAX_POS];       // origins relative to (y, x0b): |value| < 32768 checked by the launcher
    __shared__ int s_warp_count[PSCAT_THREADS / 32];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0b = blockIdx.x * PSCAT_BX, y = blockIdx.y, z0 = blockIdx.z * PSCAT_ZB;
    if (tid == 0) s_total = 0;
    __syncthreads();
    for (int base = 0; base < n_pos; base += PSCAT_THREADS) {
        const int ip = base + tid;
        bool hit = false;
        int wy = 0, wx = 0;
        if (ip < n_pos) {
            wy = pos[2 * ip]; wx = pos[2 * ip + 1];
            hit = (y >= wy && y < wy + py) && (wx < x0b + PSCAT_BX && wx + px > x0b);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_warp_count[warp] = __popc(m);
        __syncthreads();
        int off = s_total;
        for (int w = 0; w < warp; ++w) off += s_warp_count[w];
        if (hit) {
            const int k = off + __popc(m & ((1u << lane) - 1u));
            s_idx[k] = ip; s_wy[k] = short(y - wy); s_wx[k] = short(wx - x0b);
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < PSCAT_THREADS / 32; ++w) t += s_warp_count[w]; s_total += t; }
        __syncthreads();
    }
    const int n_hit = s_total;
    const int xl = tid % PSCAT_BX, zl = tid / PSCAT_BX;
    const int x = x0b + xl;
    if (x >= ox) return;
    const long long tile_px = (long long)py * px;
    for (int z = z0 + zl; z < z0 + PSCAT_ZB && z < n_slice; z += PSCAT_ZL) {
        const float2* gz = gpatch + (long long)z * n_pos * tile_px;
        float ax = 0.f, ay = 0.f;
        for (int k = 0; k < n_hit; ++k) {
            const int xx = xl - int(s_wx[k]);
            if (xx >= 0 && xx < px) {
                const float2 v = gz[(long long)s_idx[k] * tile_px + int(s_wy[k]) * px + xx];
                ax += v.x; ay += v.y;
            }
        }
        float2* o = gobj + ((long long)z * oy + y) * ox + x;
        const float2 cur = *o;
        *o = make_float2(cur.x + ax, cur.y + ay);
    }
}

// real-space propagator step (propagation.py:85-99): out = conv2d_valid(pad(in * t, edge), kernel)
constexpr int CNN_MAX_K = 33;
__constant__ float2 c_cnn_kernel[CNN_MAX_K * CNN_MAX_K];
template <int TILE>
__global__ void k_cnn_step(const float2* __restrict__ in, const float2* __restrict__ db, float2* __restrict__ out,
                           int ny, int nx, int ks, float k_dz, float2 edge) {
    extern __shared__ float2 tile[];          // (TILE + ks - 1)^2
    const int pad = (ks - 1) / 2, W = TILE + ks - 1;
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * TILE, x0 = blockIdx.x * TILE;
    const long long fbase = (long long)b * ny * nx;
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < W * W; i += blockDim.x * blockDim.y) {
        const int ty = i / W, tx = i - ty * W;
        const int y = y0 + ty - pad, x = x0 + tx - pad;
        float2 v = edge;
        if (y >= 0 && y < ny && x >= 0 && x < nx) {
            const long long o = fbase + (long long)y * nx + x;
            v = cmul1p(in[o], transmission_any_m1(db[o], k_dz));
        }
        tile[i] = v;
    }
    __syncthreads();
    const int y = y0 + threadIdx.y, x = x0 + threadIdx.x;
    if (y >= ny || x >= nx) return;
    float ax = 0.f, ay = 0.f;
    // true convolution: out[y,x] = sum_{a,b} K[a,b] * padded[y + 2p - a, x + 2p - b]
    for (int a = 0; a < ks; ++a)
        for (int c = 0; c < ks; ++c) {
            const float2 kk = c_cnn_kernel[a * ks + c];
            const float2 v = tile[(threadIdx.y + 2 * pad - a) * W + (threadIdx.x + 2 * pad - c)];
            ax += kk.x * v.x - kk.y * v.y;
            ay += kk.x * v.y + kk.y * v.x;
        }
    out[fbase + (long long)y * nx + x] = make_float2(ax, ay);
}

// adjoint of k_cnn_step: G_u[y,x] = sum_ab conj(K[a,b]) G'[y - p + a, x - p + b] (zero outside the field: the edge value is a
// constant), grad = -k (Im, Re)(conj(G_u) psi t), G = conj(t) G_u
template <int TILE>
__global__ void k_cnn_step_adj(const float2* __restrict__ gin, const float2* __restrict__ psi, const float2* __restrict__ db,
                               float2* __restrict__ grad, float2* __restrict__ gout, int ny, int nx, int ks, float k_dz) {
    extern __shared__ float2 tile[];          // (TILE + ks - 1)^2
    const int pad = (ks - 1) / 2, W = TILE + ks - 1;
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * TILE, x0 = blockIdx.x * TILE;
    const long long fbase = (long long)b * ny * nx;
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < W * W; i += blockDim.x * blockDim.y) {
        const int ty = i / W, tx = i - ty * W;
        const int y = y0 + ty - pad, x = x0 + tx - pad;
        float2 v = make_float2(0.f, 0.f);
        if (y >= 0 && y < ny && x >= 0 && x < nx) v = gin[fbase + (long long)y * nx + x];
        tile[i] = v;
    }
    __syncthreads();
    const int y = y0 + threadIdx.y, x = x0 + threadIdx.x;
    if (y >= ny || x >= nx) return;
    float ax = 0.f, ay = 0.f;
    for (int a = 0; a < ks; ++a)
        for (int c = 0; c < ks; ++c) {
            const float2 kk = c_cnn_kernel[a * ks + c];
            const float2 v = tile[(threadIdx.y + a) * W + (threadIdx.x + c)];
            ax += kk.x * v.x + kk.y * v.y;      // conj(K) * v
            ay += kk.x * v.y - kk.y * v.x;
        }
    const long long o = fbase + (long long)y * nx + x;
    const float2 gu = make_float2(ax, ay);
    const float2 tau = transmission_any_m1(db[o], k_dz);
    const float2 u = cmul1p(psi[o], tau);
    const float2 w = cmulc(u, gu);              // u conj(G_u)
    grad[o] = make_float2(-k_dz * w.y, -k_dz * w.x);
    gout[o] = cmulc1p(gu, tau);
}

static inline unsigned blocks_for(long long n, int threads) { return unsigned((n + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
struct AxisTables {
    int n = 0;
    float2* tw = nullptr;      // stage twiddles
    float2* tw_pipe = nullptr; // stage twiddles in the layout of the pipelined passes
    float2* h = nullptr;       // ifftshift(h)/n, per-slice propagator, plain fp32 rounding (single steps: bdof_slice_step)
    float2* h_adj = nullptr;   // conj
    // Error-feedback sequence of the same multiplier, [n_seq][n] (build_h_sequence): entry e is the fp32 table of the e-th
    // kernel that applies this axis' propagator, rounded so that the PRODUCT of all tables applied so far stays within one
    // fp32 rounding of the exact power of h.  A fixed fp32 table has a fixed modulus / phase error per frequency (~3e-8) that
    // compounds linearly with depth (measured: 4.5e-5 intensity error after 512 slices of a strongly scattering object).
    float2* h_seq = nullptr;
    float2* h_seq_adj = nullptr;
    int n_seq = 0;
    float2* hf = nullptr;      // free-space propagator
    float2* hf_adj = nullptr;
};

struct bdof_plan {
    int ny, nx, batch, n_slice;
    uint32_t flags;
    cudaStream_t stream;
    long long F;               // batch*ny*nx
    AxisTables ax, ay;
    bool have_kernel = false, full_kernel = false;
    bool generic = false;      // a side is not one of the power-of-two lengths: every pass runs the mixed-radix kernel (genericfft.cu)
    bool sweep = true;         // one kernel per slice and direction (sweepfft.cuh); BDOF_SWEEP=0 at plan creation selects the
                               // per-pass kernels (row pass + column pass per slice) instead
    bool resident = true;      // small square fields: one kernel per direction with the field resident on chip (residentfft.cuh);
                               // BDOF_RESIDENT=0 at plan creation selects the sweep kernels instead
    bool row_prefetch = false; // row passes L2-prefetch their own side inputs at tile start (BDOF_ROW_PREFETCH=1 enables; measured slower)
    bool l2_prefetch = false;  // column passes prefetch the next row pass's DRAM inputs into L2 (BDOF_L2_PREFETCH=1 enables;
                               // measured slower on B200 at 2048^2: the prefetch traffic slows the column pass itself)
    float2* H2 = nullptr;      // general 2-D multiplier ifftshift2(H)/(nx ny) and its conjugate
    float2* H2_adj = nullptr;
    std::complex<double> phase0{1.0, 0.0}, phasef{1.0, 0.0};
    double k_dz = 0.0;
    int free_mode = BDOF_FREE_NONE;
    float2* tmp = nullptr;     // one field
    float2* work[2] = {nullptr, nullptr};   // ping-pong fields (no-store forward) / G buffer
    float2* slabs = nullptr;   // n_slice fields (STORE_SLICES)
    float2* t_stash = nullptr; // caller-owned [n_slice][batch][ny][nx] complex64: the sweep forward leaves t_i there for the adjoint
    // window mode (bdof_plan_set_windows): d_db of bdof_forward / bdof_adjoint is the OBJECT, the batch its windows
    const int* win_origin = nullptr;
    int win_oy = 0, win_ox = 0;
    bool grad_accumulate = false;   // bdof_adjoint ADDS to d_grad_out (bdof_plan_set_grad_accumulate)
    bool stash_valid = false;  // the last forward filled t_stash and nothing has overwritten it since
    double* partial = nullptr;
    std::complex<double> total_phase{1.0, 0.0};
    bool forward_done = false;
    // gradient buckets along z: events recorded by bdof_adjoint when a bucket's gradient is final
    std::vector<cudaEvent_t> bucket_events;
    // in-situ per-variant timing (bdof_profile_*): events around every line-kernel launch
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;      // pairs (start, stop)
    std::vector<int> prof_variant;
    // device time of the last forward / adjoint (bdof_plan_last_times): one event pair around each call's launch sequence
    cudaEvent_t t_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int t_launches[2] = {0, 0};
    // host staging for bdof_forward_host
    float* e2e_delta = nullptr; float* e2e_beta = nullptr; float2* e2e_db = nullptr;
    float2* e2e_probe = nullptr; float2* e2e_exit = nullptr;
};

static int upload(float2** dst, const std::vector<float2>& v) {
    if (*dst == nullptr) CUDA_TRY(cudaMalloc((void**)dst, v.size() * sizeof(float2)));
    CUDA_TRY(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float2), cudaMemcpyHostToDevice));
    return 0;
}

// centred complex128 factor -> ifftshift, scale, fp32 (and conjugate)
static std::vector<std::complex<double>> conv_diag_gain(int n);
static std::vector<std::complex<double>> generic_diag_gain(int n);
// `generic`: the pass runs on the mixed-radix kernel (its own gain model)
static void shift_factor(const double* h, int n, double scale, std::vector<float2>& out, std::vector<float2>& out_adj, bool generic) {
    out.resize(n); out_adj.resize(n);
    const int s = n / 2;                            // ifftshift(a)[i] = a[(i + n//2) % n]
    const std::vector<std::complex<double>> g = generic ? generic_diag_gain(n) : conv_diag_gain(n);
    for (int i = 0; i < n; ++i) {
        const int src = (i + s) % n;
        const std::complex<double> hh(h[2 * src] * scale, h[2 * src + 1] * scale);
        const std::complex<double> f = hh / g[i], a = std::conj(hh) / g[i];     // the realised pair multiplies bin i by g[i]
        out[i] = make_float2(float(f.real()), float(f.imag()));
        out_adj[i] = make_float2(float(a.real()), float(a.imag()));
    }
}


// ------------------------------------------------------------------------------------------
// diagonal gain of the REALISED fp32 transform pair
// ------------------------------------------------------------------------------------------
// The butterflies' twiddle constants and the stage-twiddle table are fp32 roundings of the exact roots of unity, so the
// transform the kernels compute is a fixed linear operator ~1e-7 away from the DFT.  Its effect on one convolution
// IFFT(h FFT(x)) is, to first order, a per-frequency complex gain g_k (plus incoherent leakage between frequencies): bin k
// is multiplied by h_k g_k instead of h_k, the SAME way in every slice, so the error compounds linearly with depth -- measured
// 1.3e-5 on the far-field intensity of BASELINE config 3 (64^2 Gaussian probe, 128 slices), 4.4e-5 for a strongly scattering
// object at 512 slices; tools/fp32_fft_error_model.py reproduces both in float64 arithmetic with fp32-rounded constants.
// g_k is computable exactly on the host: the two-stage plan is  X[k1 + R1 k2] = sum_j A2[k2,j] W[j,k1] sum_r A1[k1,r] x[j + R2 r]
// with A1, A2 the realised in-register DFTs (regfft.cuh, emulated here in double with the same fp32 constants) and W the
// fp32 table, and the inverse is the conj trick on the same operator.  Dividing the multiplier tables by g_k removes the
// coherent part (config 3: 1.3e-5 -> 2e-7 in the model).
typedef std::complex<double> cdbl;

static cdbl emu_tw32(int M, int R) {       // W_R^M = exp(-2 pi i M / R) as mul_tw<M, R, false> applies it
    M %= R; if (M < 0) M += R;
    if ((4 * M) % R == 0) {
        switch ((4 * M) / R) { case 0: return {1.0, 0.0}; case 1: return {0.0, -1.0}; case 2: return {-1.0, 0.0}; default: return {0.0, 1.0}; }
    }
    const double a = 2.0 * M_PI * double(M) / double(R);
    const float c = float(cos(a)), sn = float(sin(a));
    return {double(c), -double(sn)};
}
static void emu_regfft(std::vector<cdbl>& v) {      // RegFFT<R, false>::run in double with fp32-rounded constants
    const int R = int(v.size());
    const cdbl mi(0.0, -1.0);
    if (R == 1) return;
    if (R == 2) { const cdbl a = v[0], b = v[1]; v[0] = a + b; v[1] = a - b; return; }
    if (R == 4) {
        const cdbl t0 = v[0] + v[2], t1 = v[0] - v[2], t2 = v[1] + v[3], t3 = (v[1] - v[3]) * mi;
        v[0] = t0 + t2; v[2] = t0 - t2; v[1] = t1 + t3; v[3] = t1 - t3;
        return;
    }
    const int A = (R == 8) ? 2 : 4, B = R / A;
    for (int n2 = 0; n2 < B; ++n2) {
        std::vector<cdbl> t(A);
        for (int n1 = 0; n1 < A; ++n1) t[n1] = v[B * n1 + n2];
        emu_regfft(t);
        for (int k1 = 0; k1 < A; ++k1) v[B * k1 + n2] = t[k1] * emu_tw32((n2 * k1) % R, R);
    }
    std::vector<cdbl> out(R);
    for (int k1 = 0; k1 < A; ++k1) {
        std::vector<cdbl> u(B);
        for (int n2 = 0; n2 < B; ++n2) u[n2] = v[B * k1 + n2];
        emu_regfft(u);
        for (int k2 = 0; k2 < B; ++k2) out[k1 + A * k2] = u[k2];
    }
    v = out;
}
// rho[k][r] = realised / exact entry of the radix-R butterfly
static std::vector<cdbl> emu_small_ratio(int R) {
    std::vector<cdbl> rho((size_t)R * R);
    for (int r = 0; r < R; ++r) {
        std::vector<cdbl> e(R, cdbl(0.0, 0.0));
        e[r] = 1.0;
        emu_regfft(e);
        for (int k = 0; k < R; ++k) {
            const double a = 2.0 * M_PI * double((long long)r * k % R) / double(R);
            rho[(size_t)k * R + r] = e[k] * cdbl(cos(a), sin(a));      // divide by exp(-i a)
        }
    }
    return rho;
}
static StageRadices radices_for(int n);
// g_k per FFT bin (natural order) of one convolution of length n; all ones where the model does not apply
static std::vector<cdbl> conv_diag_gain(int n) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("BDOF_FFT_GAIN"); enabled = (e && e[0] == '0') ? 0 : 1; }
    std::vector<cdbl> g((size_t)n, cdbl(1.0, 0.0));
    const StageRadices rr = radices_for(n);
    if (!enabled || rr.r1 == 0 || rr.r3 != 1 || pipe_shift(n)) return g;
    const int R1 = rr.r1, R2 = rr.r2;
    const std::vector<cdbl> rho1 = emu_small_ratio(R1), rho2 = emu_small_ratio(R2);
    // tau[j][k1] = fp32 table entry / exact W_n^{j k1} (row 0 is the implicit 1)
    std::vector<cdbl> tau((size_t)R2 * R1, cdbl(1.0, 0.0));
    for (int j = 1; j < R2; ++j)
        for (int k1 = 0; k1 < R1; ++k1) {
            const double a = -2.0 * M_PI * double((long long)j * k1 % n) / double(n);
            const cdbl tab(double(float(cos(a))), double(float(sin(a))));
            tau[(size_t)j * R1 + k1] = tab * cdbl(cos(a), -sin(a));
        }
    std::vector<cdbl> s1(R1), c2(R2);                 // row sums of rho1, column sums of rho2
    for (int k1 = 0; k1 < R1; ++k1) { cdbl a = 0.0; for (int r = 0; r < R1; ++r) a += rho1[(size_t)k1 * R1 + r]; s1[k1] = a / double(R1); }
    for (int j = 0; j < R2; ++j) { cdbl a = 0.0; for (int k2 = 0; k2 < R2; ++k2) a += rho2[(size_t)k2 * R2 + j]; c2[j] = a; }
    for (int k = 0; k < n; ++k) {
        // forward: k is the OUTPUT index k1 + R1 k2
        const int k1 = k % R1, k2 = k / R1;
        cdbl gf = 0.0;
        for (int j = 0; j < R2; ++j) gf += rho2[(size_t)k2 * R2 + j] * tau[(size_t)j * R1 + k1];
        gf = gf / double(R2) * s1[k1];
        // inverse (conj trick): k is the INPUT index j + R2 r
        const int j = k % R2, r = k / R2;
        cdbl gi = 0.0;
        for (int q1 = 0; q1 < R1; ++q1) gi += tau[(size_t)j * R1 + q1] * rho1[(size_t)q1 * R1 + r];
        gi = std::conj(gi * c2[j] / double(n));
        g[k] = gf * gi;
    }
    return g;
}

// The same gain for the mixed-radix passes (genericfft.cuh): their Stockham stages are emulated in double with the fp32
// table W_n^q, one basis vector at a time, which yields the whole realised matrix F~; the forward gain of bin k is the mean
// of F~[k, m] / F[k, m] over the inputs m, the inverse (conj trick) gain the conjugate mean over the outputs of column k.
// O(n^2 * sum of radices) once per length (cached): 72 -> microseconds, 2000 -> ~0.3 s.
static std::vector<cdbl> generic_diag_gain(int n) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("BDOF_FFT_GAIN"); enabled = (e && e[0] == '0') ? 0 : 1; }
    std::vector<cdbl> g((size_t)n, cdbl(1.0, 0.0));
    int radix[16], n_stages = 0;
    {   // as gen_factorize(): 4s first, then 2, 3, 5, 7
        int m = n;
        while (m % 4 == 0 && n_stages < 16) { radix[n_stages++] = 4; m /= 4; }
        const int primes[4] = {2, 3, 5, 7};
        for (int pi = 0; pi < 4; ++pi) while (m % primes[pi] == 0 && n_stages < 16) { radix[n_stages++] = primes[pi]; m /= primes[pi]; }
        if (m != 1) return g;
    }
    if (!enabled || n < 2) return g;
    static std::mutex mu;
    static std::map<int, std::vector<cdbl>> cache;
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(n);
        if (it != cache.end()) return it->second;
    }
    std::vector<cdbl> tw(n), ex(n);                  // fp32 table as make_full_twiddles(), exact roots
    for (int k = 0; k < n; ++k) {
        const double a = -2.0 * M_PI * double(k) / double(n);
        tw[k] = cdbl(double(float(cos(a))), double(float(sin(a))));
        ex[k] = cdbl(cos(a), sin(a));
    }
    std::vector<cdbl> gf((size_t)n, 0.0), gi((size_t)n, 0.0), x(n), y(n);
    const cdbl I(0.0, 1.0);
    for (int m0 = 0; m0 < n; ++m0) {
        std::fill(x.begin(), x.end(), cdbl(0.0, 0.0));
        x[m0] = 1.0;
        int ns = 1;
        for (int s = 0; s < n_stages; ++s) {         // gen_stage() for every work item j
            const int R = radix[s], m = n / R, tstep = n / (ns * R), rstep = n / R;
            for (int j = 0; j < m; ++j) {
                const int k = j % ns;
                cdbl v[8];
                for (int r = 0; r < R; ++r) {
                    v[r] = x[j + r * m];
                    if (r > 0 && k > 0) v[r] *= tw[r * k * tstep];
                }
                const int o = (j / ns) * ns * R + k;
                if (R == 2) { y[o] = v[0] + v[1]; y[o + ns] = v[0] - v[1]; }
                else if (R == 4) {
                    const cdbl s0 = v[0] + v[2], d0 = v[0] - v[2], s1 = v[1] + v[3], d1 = v[1] - v[3];
                    y[o] = s0 + s1; y[o + ns] = d0 - I * d1; y[o + 2 * ns] = s0 - s1; y[o + 3 * ns] = d0 + I * d1;
                } else {
                    for (int q = 0; q < R; ++q) {
                        cdbl acc = v[0];
                        for (int r = 1; r < R; ++r) acc += v[r] * tw[((r * q) % R) * rstep];
                        y[o + q * ns] = acc;
                    }
                }
            }
            x.swap(y);
            ns *= R;
        }
        cdbl col = 0.0;
        for (int k = 0; k < n; ++k) {
            const cdbl ratio = x[k] * std::conj(ex[(long long)k * m0 % n]);
            gf[k] += ratio;
            col += ratio;
        }
        gi[m0] = std::conj(col / double(n));
    }
    for (int k = 0; k < n; ++k) g[k] = gf[k] / double(n) * gi[k];
    std::lock_guard<std::mutex> lk(mu);
    cache[n] = g;
    return g;
}

// ------------------------------------------------------------------------------------------
// error-feedback multiplier tables
// ------------------------------------------------------------------------------------------
static inline bool slice_propagates(const bdof_plan* p, int i);
static bool use_sweep(const bdof_plan* p);

// Applications of one axis' propagator per kernel of the schedule: entry i < Z = the kernel(s) of slice i, entry Z = the
// trailing half propagation of the TF semantics.  Sweep schedule: the kernel of slice i works along axis a(i) = i & 1 and
// applies it (i > 0) + propagates(i) times with ONE table; per-pass schedule: one row pass and one column pass per
// propagating slice.
static void h_schedule(const bdof_plan* p, int col, std::vector<int>& napp) {
    const int Z = p->n_slice;
    napp.assign(Z + 1, 0);
    if (use_sweep(p)) {
        for (int i = 0; i < Z; ++i)
            if ((i & 1) == col) napp[i] = (i > 0 ? 1 : 0) + (slice_propagates(p, i) ? 1 : 0);
        if (slice_propagates(p, Z - 1) && (Z & 1) == col) napp[Z] = 1;
    } else {
        for (int i = 0; i < Z; ++i) napp[i] = slice_propagates(p, i) ? 1 : 0;
    }
}

static int build_axis_sequence(bdof_plan* p, AxisTables& a, int col, const double* h_centred) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("BDOF_H_FEEDBACK"); enabled = (e && e[0] == '0') ? 0 : 1; }
    cudaFree(a.h_seq); cudaFree(a.h_seq_adj);
    a.h_seq = a.h_seq_adj = nullptr; a.n_seq = 0;
    if (!enabled) return 0;
    typedef std::complex<long double> cld;
    const int n = a.n, Z = p->n_slice, s = n / 2;
    std::vector<int> napp;
    h_schedule(p, col, napp);
    std::vector<float2> seq((size_t)(Z + 1) * n), seq_adj((size_t)(Z + 1) * n);
    const long double scale = 1.0L / (long double)n;
    const std::vector<cdbl> gain = p->generic ? generic_diag_gain(n) : conv_diag_gain(n);
    for (int i = 0; i < n; ++i) {
        const int src = (i + s) % n;                    // ifftshift, as shift_factor()
        const cld h((long double)h_centred[2 * src], (long double)h_centred[2 * src + 1]);
        const cld g((long double)gain[i].real(), (long double)gain[i].imag());      // the realised pair multiplies bin i by g
        const bool degenerate = std::abs(h) < 1e-30L;
        cld P(1.0L, 0.0L), Q(1.0L, 0.0L);               // exact and realised cumulative products
        for (int e = 0; e <= Z; ++e) {
            cld want = h;
            if (napp[e] > 0 && !degenerate) {
                cld hn = h;
                for (int k = 1; k < napp[e]; ++k) hn *= h;
                P *= hn;
                const cld r = P / Q / hn;               // = 1 + (accumulated relative error so far)
                want = h * (cld(1.0L, 0.0L) + (r - cld(1.0L, 0.0L)) / (long double)napp[e]);    // first-order root
            }
            const cld tab = want / g, tab_adj = std::conj(want) / g;
            const float re = float(tab.real() * scale), im = float(tab.imag() * scale);
            seq[(size_t)e * n + i] = make_float2(re, im);
            seq_adj[(size_t)e * n + i] = make_float2(float(tab_adj.real() * scale), float(tab_adj.imag() * scale));
            if (napp[e] > 0 && !degenerate) {
                const cld got = cld((long double)re * (long double)n, (long double)im * (long double)n) * g;   // effective multiplier
                for (int k = 0; k < napp[e]; ++k) Q *= got;
            }
        }
    }
    BDOF_TRY(upload(&a.h_seq, seq));
    BDOF_TRY(upload(&a.h_seq_adj, seq_adj));
    a.n_seq = Z + 1;
    return 0;
}
static int build_h_sequence(bdof_plan* p, const double* h_hy, const double* h_hx) {
    BDOF_TRY(build_axis_sequence(p, p->ax, 0, h_hx));
    return build_axis_sequence(p, p->ay, 1, h_hy);
}
// table of the e-th kernel of the schedule (e < 0: the plain table)
static const float2* h_entry(const AxisTables& a, int e, bool adj) {
    if (a.h_seq == nullptr || e < 0 || e >= a.n_seq) return adj ? a.h_adj : a.h;
    return (adj ? a.h_seq_adj : a.h_seq) + (size_t)e * a.n;
}

extern "C" int bdof_kernel_factors(double dist_nm, double lmbda_nm, const double* voxel_nm, int ny, int nx,
                                   double pi_const, double* hy_out, double* hx_out, double* phase0_out) {
    if (!voxel_nm || !hy_out || !hx_out || !phase0_out || ny < 1 || nx < 1) return fail(BDOF_E_BADARG, "bad argument");
    // util.py:176-181: u along axis 1 spans +-1/(2 voxel[0]) over nx points (endpoint inclusive),
    // v along axis 0 spans +-1/(2 voxel[1]) over ny points
    const double k = 2.0 * pi_const / lmbda_nm;
    const double u_max = 1.0 / (2.0 * voxel_nm[0]), v_max = 1.0 / (2.0 * voxel_nm[1]);
    auto fill = [&](double* out, int n, double mx) {
        for (int i = 0; i < n; ++i) {
            // numpy.linspace(-mx, mx, n): start + i*step, last point forced to stop
            double step = (n > 1) ? (2.0 * mx) / double(n - 1) : 0.0;
            double u = (i == n - 1 && n > 1) ? mx : -mx + double(i) * step;
            double ph = -pi_const * lmbda_nm * dist_nm * (u * u);
            out[2 * i] = cos(ph);
            out[2 * i + 1] = sin(ph);
        }
    };
    fill(hy_out, ny, v_max);
    fill(hx_out, nx, u_max);
    phase0_out[0] = cos(k * dist_nm);
    phase0_out[1] = sin(k * dist_nm);
    return 0;
}

extern "C" int bdof_plan_create(bdof_plan** out, int ny, int nx, int batch, int n_slice, uint32_t flags, void* cuda_stream) {
    if (!out || ny < 1 || nx < 1 || batch < 1 || n_slice < 1) return fail(BDOF_E_BADARG, "bad plan shape");
    if (!bdof_size_supported(ny) || !bdof_size_supported(nx))
        return fail(BDOF_E_UNSUPPORTED, "field %dx%d: each side must be a power of two in [64, 8192] or 2^a 3^b 5^c 7^d <= 2048", ny, nx);
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    bdof_plan* p = new bdof_plan();
    p->ny = ny; p->nx = nx; p->batch = batch; p->n_slice = n_slice; p->flags = flags;
    p->stream = (cudaStream_t)cuda_stream;
    p->F = (long long)batch * ny * nx;
    p->ax.n = nx; p->ay.n = ny;
    p->generic = bdof_size_supported(ny) != 1 || bdof_size_supported(nx) != 1;
    if (const char* e = getenv("BDOF_SWEEP")) p->sweep = (e[0] != '0');
    if (const char* e = getenv("BDOF_RESIDENT")) p->resident = (e[0] != '0');
    if (const char* e = getenv("BDOF_L2_PREFETCH")) p->l2_prefetch = (e[0] != '0');
    if (const char* e = getenv("BDOF_ROW_PREFETCH")) p->row_prefetch = (e[0] != '0');
    int r = 0;
    do {
        if (p->generic) {
            if ((r = upload(&p->ax.tw, make_full_twiddles(nx)))) break;
            if ((r = upload(&p->ay.tw, make_full_twiddles(ny)))) break;
        } else {
            if ((r = upload(&p->ax.tw, make_twiddles(nx)))) break;
            if ((r = upload(&p->ay.tw, make_twiddles(ny)))) break;
            if ((r = upload(&p->ax.tw_pipe, make_twiddles_pipe(nx)))) break;
            if ((r = upload(&p->ay.tw_pipe, make_twiddles_pipe(ny)))) break;
        }
        cudaError_t e;
        if ((e = cudaMalloc((void**)&p->tmp, p->F * sizeof(float2))) != cudaSuccess) { r = fail(int(e), "cudaMalloc tmp: %s", cudaGetErrorString(e)); break; }
        if ((e = cudaMalloc((void**)&p->work[0], p->F * sizeof(float2))) != cudaSuccess) { r = fail(int(e), "cudaMalloc work: %s", cudaGetErrorString(e)); break; }
        if (!(flags & BDOF_STORE_SLICES)) {
            if ((e = cudaMalloc((void**)&p->work[1], p->F * sizeof(float2))) != cudaSuccess) { r = fail(int(e), "cudaMalloc work: %s", cudaGetErrorString(e)); break; }
        } else {
            if ((e = cudaMalloc((void**)&p->slabs, (size_t)n_slice * p->F * sizeof(float2))) != cudaSuccess) { r = fail(int(e), "cudaMalloc slice store (%lld bytes): %s", (long long)n_slice * p->F * 8, cudaGetErrorString(e)); break; }
        }
        if ((e = cudaMalloc((void**)&p->partial, LOSS_BLOCKS * sizeof(double))) != cudaSuccess) { r = fail(int(e), "cudaMalloc: %s", cudaGetErrorString(e)); break; }
        for (int i = 0; i < 4 && r == 0; ++i)
            if ((e = cudaEventCreate(&p->t_ev[i])) != cudaSuccess) r = fail(int(e), "cudaEventCreate: %s", cudaGetErrorString(e));
    } while (0);
    if (r) { bdof_plan_destroy(p); return r; }
    *out = p;
    return 0;
}

static void free_axis(AxisTables& a) {
    cudaFree(a.tw); cudaFree(a.tw_pipe); cudaFree(a.h); cudaFree(a.h_adj); cudaFree(a.hf); cudaFree(a.hf_adj);
    cudaFree(a.h_seq); cudaFree(a.h_seq_adj);
}
extern "C" void bdof_plan_destroy(bdof_plan* p) {
    if (!p) return;
    free_axis(p->ax); free_axis(p->ay);
    cudaFree(p->H2); cudaFree(p->H2_adj);
    cudaFree(p->tmp); cudaFree(p->work[0]); cudaFree(p->work[1]); cudaFree(p->slabs); cudaFree(p->partial);
    cudaFree(p->e2e_delta); cudaFree(p->e2e_beta); cudaFree(p->e2e_db); cudaFree(p->e2e_probe); cudaFree(p->e2e_exit);
    for (int i = 0; i < 4; ++i) if (p->t_ev[i]) cudaEventDestroy(p->t_ev[i]);
    delete p;
}

extern "C" int bdof_plan_workspace_bytes(const bdof_plan* p, size_t* bytes_out) {
    if (!p || !bytes_out) return fail(BDOF_E_BADARG, "null");
    size_t f = size_t(p->F) * sizeof(float2);
    *bytes_out = f * 2 + ((p->flags & BDOF_STORE_SLICES) ? f * p->n_slice : f);
    return 0;
}

extern "C" int bdof_set_kernel(bdof_plan* p, const double* h_hy, const double* h_hx, double phase0_re, double phase0_im, double k_dz) {
    if (!p || !h_hy || !h_hx) return fail(BDOF_E_BADARG, "null");
    std::vector<float2> a, b;
    shift_factor(h_hx, p->nx, 1.0 / p->nx, a, b, p->generic);
    BDOF_TRY(upload(&p->ax.h, a)); BDOF_TRY(upload(&p->ax.h_adj, b));
    shift_factor(h_hy, p->ny, 1.0 / p->ny, a, b, p->generic);
    BDOF_TRY(upload(&p->ay.h, a)); BDOF_TRY(upload(&p->ay.h_adj, b));
    p->phase0 = {phase0_re, phase0_im};
    p->k_dz = k_dz;
    p->have_kernel = true; p->full_kernel = false;
    return build_h_sequence(p, h_hy, h_hx);
}

extern "C" int bdof_set_kernel_full(bdof_plan* p, const double* h_H, double k_dz) {
    if (!p || !h_H) return fail(BDOF_E_BADARG, "null");
    const int ny = p->ny, nx = p->nx;
    std::vector<float2> a((size_t)ny * nx), b((size_t)ny * nx);
    const double sc = 1.0 / (double(nx) * double(ny));
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) {
            const size_t src = (size_t)((y + ny / 2) % ny) * nx + (x + nx / 2) % nx;
            a[(size_t)y * nx + x] = make_float2(float(h_H[2 * src] * sc), float(h_H[2 * src + 1] * sc));
            b[(size_t)y * nx + x] = make_float2(float(h_H[2 * src] * sc), float(-h_H[2 * src + 1] * sc));
        }
    BDOF_TRY(upload(&p->H2, a)); BDOF_TRY(upload(&p->H2_adj, b));
    p->phase0 = {1.0, 0.0};
    p->k_dz = k_dz;
    p->have_kernel = true; p->full_kernel = true;
    return 0;
}

extern "C" int bdof_set_free_prop(bdof_plan* p, int mode, const double* h_hy, const double* h_hx, double phase0_re, double phase0_im) {
    if (!p) return fail(BDOF_E_BADARG, "null");
    if (mode != BDOF_FREE_NONE && mode != BDOF_FREE_INF && mode != BDOF_FREE_TF) return fail(BDOF_E_BADARG, "bad free-space mode %d", mode);
    p->phasef = {1.0, 0.0};
    if (mode == BDOF_FREE_TF) {
        if (!h_hy || !h_hx) return fail(BDOF_E_BADARG, "free-space factors missing");
        std::vector<float2> a, b;
        shift_factor(h_hx, p->nx, 1.0 / p->nx, a, b, p->generic);
        BDOF_TRY(upload(&p->ax.hf, a)); BDOF_TRY(upload(&p->ax.hf_adj, b));
        shift_factor(h_hy, p->ny, 1.0 / p->ny, a, b, p->generic);
        BDOF_TRY(upload(&p->ay.hf, a)); BDOF_TRY(upload(&p->ay.hf_adj, b));
        p->phasef = {phase0_re, phase0_im};
    }
    p->free_mode = mode;
    return 0;
}

static long long* g_dbg = nullptr;      // phase-timing buffer for instrumented builds (bdof_debug_set_buffer)
extern "C" int bdof_debug_set_buffer(void* d_buf) { g_dbg = reinterpret_cast<long long*>(d_buf); return 0; }

// ------------------------------------------------------------------------------------------
// pass helpers
// ------------------------------------------------------------------------------------------
static LineParams row_params(const bdof_plan* p, const float2* in, float2* out, const float2* h) {
    LineParams q{};
    q.in = in; q.out = out; q.h = h; q.tw = p->ax.tw;
    q.batch_stride = (long long)p->ny * p->nx; q.db_batch_stride = q.batch_stride;
    q.lines_per_batch = p->ny; q.elem_stride = 1; q.line_stride = p->nx;
    q.k_dz = float(p->k_dz);
    q.dbg = g_dbg;
    q.pf_bytes = p->row_prefetch ? 0 : -1;
    return q;
}
static LineParams col_params(const bdof_plan* p, const float2* in, float2* out, const float2* h) {
    LineParams q{};
    q.in = in; q.out = out; q.h = h; q.tw = p->ay.tw;
    q.batch_stride = (long long)p->ny * p->nx; q.db_batch_stride = q.batch_stride;
    q.lines_per_batch = p->nx; q.elem_stride = p->nx; q.line_stride = 1;
    q.k_dz = float(p->k_dz);
    q.dbg = g_dbg;
    q.pf_bytes = p->row_prefetch ? 0 : -1;
    return q;
}
static int timed_launch(bdof_plan* p, int n, int variant, const LineParams& q0, long long n_lines) {
    LineParams q = q0;
    const int pv = (variant == V_COL_CONV_PIPE) ? int(V_COL_CONV) : variant;    // reported under the pass it implements
    if (g_dbg) q.dbg = g_dbg + (long long)pv * (1 << 17);     // one region per pass variant
    { static int flags = -1; if (flags < 0) { const char* e = getenv("BDOF_DBG_FLAGS"); flags = e ? atoi(e) : 0; } q.dbg_flags = flags; }
    { static int tune = -1; if (tune < 0) { const char* e = getenv("BDOF_TUNE"); tune = e ? atoi(e) : 0; } q.tune = tune; }
    auto launch = [&]() { return p->generic ? bdof_launch_line_generic(n, variant, q, n_lines, p->stream) : launch_variant(n, variant, q, n_lines, p->stream); };
    if (!p->profile) return launch();
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    CUDA_TRY(cudaEventRecord(a, p->stream));
    int r = launch();
    CUDA_TRY(cudaEventRecord(b, p->stream));
    p->prof_events.push_back(a); p->prof_events.push_back(b);
    p->prof_variant.push_back(pv);
    return r;
}
static int row_pass(bdof_plan* p, int variant, const LineParams& q) {
    return timed_launch(p, p->nx, variant, q, (long long)p->batch * p->ny);
}
static bool use_pipe() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("BDOF_PIPE"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
static int col_pass(bdof_plan* p, int variant, const LineParams& q) {
    if (variant == V_COL_CONV && !p->generic && use_pipe() && pipe_parts(p->ny) > 0) {
        LineParams q2 = q;
        q2.tw = p->ay.tw_pipe;
        return timed_launch(p, p->ny, V_COL_CONV_PIPE, q2, (long long)p->batch * p->nx);
    }
    return timed_launch(p, p->ny, variant, q, (long long)p->batch * p->nx);
}

// ------------------------------------------------------------------------------------------
// sweep kernels: one launch per slice and direction (sweepfft.cuh)
// ------------------------------------------------------------------------------------------
static bool use_sweep(const bdof_plan* p) {
    return p->sweep && !(p->flags & BDOF_STEPWISE) && !p->generic && !p->full_kernel && p->n_slice >= 2 && pipe_parts(p->nx) > 0 && pipe_parts(p->ny) > 0;
}
// Small square fields run as ONE kernel per direction with the field resident on chip (residentfft.cuh); same schedule and
// multiplier tables as the sweep kernels.  BDOF_RESIDENT=0 selects the sweep kernels instead.
static bool use_resident(const bdof_plan* p) {
    return p->resident && use_sweep(p) && p->ny == p->nx && bdof_resident_supported(p->nx) && p->ax.h_seq != nullptr && p->ay.h_seq != nullptr;
}
static int resident_launch(bdof_plan* p, bool adj, ResidentParams q) {
    q.hx = adj ? p->ax.h_seq_adj : p->ax.h_seq;
    q.hy = adj ? p->ay.h_seq_adj : p->ay.h_seq;
    q.tw = p->ax.tw;
    q.slab = p->slabs;
    q.slice_stride = p->F;
    q.n_slice = p->n_slice; q.batch = p->batch;
    q.propagate_last = (p->flags & BDOF_PROPAGATE_LAST) ? 1 : 0;
    q.k_dz = float(p->k_dz);
    if (!p->profile) return bdof_launch_resident(p->nx, adj ? 1 : 0, q, p->stream);
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    CUDA_TRY(cudaEventRecord(a, p->stream));
    int r = bdof_launch_resident(p->nx, adj ? 1 : 0, q, p->stream);
    CUDA_TRY(cudaEventRecord(b, p->stream));
    p->prof_events.push_back(a); p->prof_events.push_back(b);
    p->prof_variant.push_back(adj ? V_RESIDENT_ADJ : V_RESIDENT_FWD);
    return r;
}

// kernel of slice i: x kernel (rows) for even i, y kernel (columns) for odd i
static int sweep_launch(bdof_plan* p, int i, bool adj, SweepParams q) {
    const bool col = (i & 1) != 0;
    const int n = col ? p->ny : p->nx;
    q.tw = col ? p->ay.tw_pipe : p->ax.tw_pipe;
    q.h = h_entry(col ? p->ay : p->ax, i, adj);
    q.k_dz = float(p->k_dz);
    { static int pf = -1; if (pf < 0) { const char* e = getenv("BDOF_SLAB_PREFETCH"); pf = (e && e[0] == '0') ? 0 : 1; } q.slab_prefetch = pf; }
    { static int sg = -1; if (sg < 0) { const char* e = getenv("BDOF_STAGGER_NS"); sg = e ? atoi(e) : 0; } q.stagger_ns = sg; }
    q.dbg = g_dbg ? g_dbg + (long long)((col ? 2 : 0) + (adj ? 1 : 0)) * (1 << 17) : nullptr;
    const long long rows = (long long)p->batch * p->ny;
    if (!p->profile) return launch_sweep_n(n, col, adj, q, rows, p->nx, p->stream);
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    CUDA_TRY(cudaEventRecord(a, p->stream));
    int r = launch_sweep_n(n, col, adj, q, rows, p->nx, p->stream);
    CUDA_TRY(cudaEventRecord(b, p->stream));
    p->prof_events.push_back(a); p->prof_events.push_back(b);
    p->prof_variant.push_back(adj ? V_SWEEP_ADJ : V_SWEEP_FWD);
    return r;
}

// one propagation of slice i: out = P(in * t(db))
static int propagate_slice(bdof_plan* p, const float2* in, const float2* db, float2* out, const float2* db_next, int seq) {
    if (!p->full_kernel) {
        LineParams r = row_params(p, in, p->tmp, h_entry(p->ax, seq, false));
        r.db = db;
        BDOF_TRY(row_pass(p, V_ROW_CONV_T, r));
        LineParams c = col_params(p, p->tmp, out, h_entry(p->ay, seq, false));
        if (db_next != nullptr && p->l2_prefetch) { c.pf0 = db_next; c.pf_bytes = p->F * (long long)sizeof(float2); }
        return col_pass(p, V_COL_CONV, c);
    }
    // general 2-D H: modulate, FFT_x, (FFT_y * H * IFFT_y), IFFT_x
    k_modulate<<<blocks_for(p->F, 256), 256, 0, p->stream>>>(in, db, p->tmp, p->F, float(p->k_dz));
    BDOF_TRY(launch_check("k_modulate"));
    BDOF_TRY(row_pass(p, V_ROW_FWD, row_params(p, p->tmp, p->tmp, nullptr)));
    LineParams c = col_params(p, p->tmp, p->tmp, p->H2);
    BDOF_TRY(col_pass(p, V_COL_CONV2D, c));
    return row_pass(p, V_ROW_INV, row_params(p, p->tmp, out, nullptr));
}

// adjoint of propagate_slice: G <- conj(t) P^H G, grad = -k (Im, Re)(conj(P^H G) psi t)
static int propagate_slice_adj(bdof_plan* p, float2* G, const float2* psi, const float2* db, float2* grad, int seq) {
    if (!p->full_kernel) {
        LineParams c = col_params(p, G, p->tmp, h_entry(p->ay, seq, true));
        if (p->l2_prefetch) { c.pf0 = db; c.pf1 = psi; c.pf_bytes = p->F * (long long)sizeof(float2); }
        BDOF_TRY(col_pass(p, V_COL_CONV, c));
        LineParams r = row_params(p, p->tmp, G, h_entry(p->ax, seq, true));
        r.db = db; r.psi = psi; r.grad = grad;
        return row_pass(p, V_ROW_CONV_ADJ, r);
    }
    BDOF_TRY(row_pass(p, V_ROW_FWD, row_params(p, G, p->tmp, nullptr)));
    LineParams c = col_params(p, p->tmp, p->tmp, p->H2_adj);
    BDOF_TRY(col_pass(p, V_COL_CONV2D, c));
    BDOF_TRY(row_pass(p, V_ROW_INV, row_params(p, p->tmp, G, nullptr)));
    k_modulate_adj<<<blocks_for(p->F, 256), 256, 0, p->stream>>>(G, psi, db, grad, p->F, float(p->k_dz));
    return launch_check("k_modulate_adj");
}

static inline bool slice_propagates(const bdof_plan* p, int i) {
    return (p->flags & BDOF_PROPAGATE_LAST) ? (p->n_slice > 1) : (i < p->n_slice - 1);
}

extern "C" int bdof_forward(bdof_plan* p, const float* d_db_f, const float* d_probe_f, float* d_exit_f) {
    if (!p || !d_db_f || !d_probe_f || !d_exit_f) return fail(BDOF_E_BADARG, "null");
    if (!p->have_kernel) return fail(BDOF_E_STATE, "bdof_set_kernel has not been called");
    const float2* d_db = reinterpret_cast<const float2*>(d_db_f);
    const float2* d_probe = reinterpret_cast<const float2*>(d_probe_f);
    float2* d_exit = reinterpret_cast<float2*>(d_exit_f);
    const bool store = p->flags & BDOF_STORE_SLICES;
    const long long per = (long long)p->ny * p->nx;
    const int Z = p->n_slice;

    float2* cur = store ? p->slabs : p->work[0];
    p->stash_valid = false;
    const unsigned long long launches0 = g_launches.load();
    CUDA_TRY(cudaEventRecord(p->t_ev[0], p->stream));
    const bool resident = use_resident(p);
    if (p->win_origin && !resident) return fail(BDOF_E_UNSUPPORTED, "window mode needs the resident small-field kernels (square 64 x 64 fields)");
    std::complex<double> phase{1.0, 0.0};
    if (resident) {
        // the whole object part of the chain in one launch: the field never leaves the SM
        float2* obj_out = (p->free_mode == BDOF_FREE_NONE) ? d_exit : p->work[0];
        ResidentParams q{};
        q.in = d_probe; q.out = obj_out; q.db = d_db;
        q.db_slice_stride = (p->flags & BDOF_Z_BROADCAST) ? 0 : p->F;
        q.stash = (store && p->t_stash) ? p->t_stash : nullptr;
        if (p->win_origin) {
            q.win = p->win_origin; q.oy = p->win_oy; q.ox = p->win_ox;
            q.db_slice_stride = (p->flags & BDOF_Z_BROADCAST) ? 0 : (long long)p->win_oy * p->win_ox;
        }
        q.store = store ? 1 : 0;
        BDOF_TRY(resident_launch(p, false, q));
        for (int i = 0; i < Z; ++i) if (slice_propagates(p, i)) phase *= p->phase0;
        cur = obj_out;
        p->stash_valid = store && p->t_stash != nullptr;
    } else {
        k_broadcast_probe<<<blocks_for(per, 256), 256, 0, p->stream>>>(d_probe, cur, per, p->batch);
        BDOF_TRY(launch_check("k_broadcast_probe"));
    }
    if (!resident && use_sweep(p)) {
        // slice i runs as ONE kernel along axis a(i) (x for even i, y for odd i): second half of the propagation
        // of slice i-1, modulation by slice i, first half of the propagation of slice i
        float2* obj_out = (p->free_mode == BDOF_FREE_NONE) ? d_exit : p->work[0];
        if (!store && obj_out == cur) obj_out = p->work[1];
        float2* A = p->tmp;
        for (int i = 0; i < Z; ++i) {
            const bool prop = slice_propagates(p, i);
            SweepParams q{};
            q.in = (i == 0) ? cur : A;
            q.out = prop ? A : obj_out;
            q.db = d_db + ((p->flags & BDOF_Z_BROADCAST) ? 0 : (long long)i * p->F);
            q.slab = (store && i > 0) ? p->slabs + (long long)i * p->F : nullptr;
            q.conv1 = (i > 0); q.conv2 = prop; q.store_slab = (store && i > 0); q.store_out = 1;
            q.grad = (store && p->t_stash) ? p->t_stash + (long long)i * p->F : nullptr;
            BDOF_TRY(sweep_launch(p, i, false, q));
            if (prop) phase *= p->phase0;
        }
        cur = obj_out;
        p->stash_valid = store && p->t_stash != nullptr;
        if (slice_propagates(p, Z - 1)) {
            // TF semantics: the last slice propagates too -> its second half along the other axis
            if (Z & 1) BDOF_TRY(col_pass(p, V_COL_CONV, col_params(p, A, obj_out, h_entry(p->ay, Z, false))));
            else       BDOF_TRY(row_pass(p, V_ROW_CONV, row_params(p, A, obj_out, h_entry(p->ax, Z, false))));
        }
    }
    // where the object part of the chain leaves its result
    float2* obj_out = (p->free_mode == BDOF_FREE_NONE) ? d_exit : (store ? p->work[0] : nullptr);
    for (int i = 0; i < Z && !use_sweep(p); ++i) {      // (resident implies use_sweep)
        const float2* db_i = d_db + ((p->flags & BDOF_Z_BROADCAST) ? 0 : (long long)i * p->F);
        const bool last = (i == Z - 1);
        float2* dst;
        if (last) dst = obj_out ? obj_out : (cur == p->work[0] ? p->work[1] : p->work[0]);
        else if (store) dst = p->slabs + (long long)(i + 1) * p->F;
        else dst = (cur == p->work[0]) ? p->work[1] : p->work[0];
        if (slice_propagates(p, i)) {
            const float2* db_next = (i + 1 < Z && !(p->flags & BDOF_Z_BROADCAST)) ? d_db + (long long)(i + 1) * p->F : nullptr;
            BDOF_TRY(propagate_slice(p, cur, db_i, dst, db_next, i));
            phase *= p->phase0;
        } else {
            k_modulate<<<blocks_for(p->F, 256), 256, 0, p->stream>>>(cur, db_i, dst, p->F, float(p->k_dz));
            BDOF_TRY(launch_check("k_modulate"));
        }
        cur = dst;
    }
    if (p->free_mode == BDOF_FREE_INF) {
        // fftshift(fft2(psi)) (npfuncs.py:46)
        LineParams r = row_params(p, cur, p->tmp, nullptr);
        r.out_shift = p->nx / 2;
        BDOF_TRY(row_pass(p, V_ROW_FWD, r));
        LineParams c = col_params(p, p->tmp, d_exit, nullptr);
        c.out_shift = p->ny / 2;
        BDOF_TRY(col_pass(p, V_COL_FWD, c));
    } else if (p->free_mode == BDOF_FREE_TF) {
        BDOF_TRY(row_pass(p, V_ROW_CONV, row_params(p, cur, p->tmp, p->ax.hf)));
        BDOF_TRY(col_pass(p, V_COL_CONV, col_params(p, p->tmp, d_exit, p->ay.hf)));
        phase *= p->phasef;
    }
    p->total_phase = phase;
    if (std::abs(phase - std::complex<double>(1.0, 0.0)) > 0.0) {
        k_scale_complex<<<blocks_for(p->F, 256), 256, 0, p->stream>>>(d_exit, d_exit, p->F, float(phase.real()), float(phase.imag()));
        BDOF_TRY(launch_check("k_scale_complex"));
    }
    p->forward_done = true;
    CUDA_TRY(cudaEventRecord(p->t_ev[1], p->stream));
    p->t_launches[0] = int(g_launches.load() - launches0);
    return 0;
}

extern "C" int bdof_loss_mag(bdof_plan* p, const float* d_exit, const float* d_target_mag, double loss_scale,
                             double* d_loss, float* d_grad_exit) {
    if (!p || !d_exit || !d_target_mag || !d_loss) return fail(BDOF_E_BADARG, "null");
    const double inv_m = 1.0 / double(p->F);
    k_loss_mag<<<LOSS_BLOCKS, LOSS_THREADS, 0, p->stream>>>(reinterpret_cast<const float2*>(d_exit), d_target_mag,
                                                            reinterpret_cast<float2*>(d_grad_exit), p->F,
                                                            float(2.0 * inv_m * loss_scale), p->partial);
    BDOF_TRY(launch_check("k_loss_mag"));
    k_loss_final<<<1, 32, 0, p->stream>>>(p->partial, LOSS_BLOCKS, inv_m * loss_scale, d_loss);
    return launch_check("k_loss_final");
}

extern "C" int bdof_adjoint(bdof_plan* p, float* d_db_inout, const float* d_grad_exit, float* d_grad_out, float* d_grad_probe) {
    if (!p || !d_db_inout || !d_grad_exit) return fail(BDOF_E_BADARG, "null");
    if (!(p->flags & BDOF_STORE_SLICES)) return fail(BDOF_E_STATE, "plan was created without BDOF_STORE_SLICES");
    if (!p->forward_done) return fail(BDOF_E_STATE, "bdof_adjoint before bdof_forward");
    const bool zb = p->flags & BDOF_Z_BROADCAST;
    if (zb && !d_grad_out) return fail(BDOF_E_BADARG, "BDOF_Z_BROADCAST needs d_grad_out");
    if (p->win_origin && !d_grad_out) return fail(BDOF_E_BADARG, "window mode needs d_grad_out (the per-window gradients)");
    if (p->grad_accumulate) {
        if (!d_grad_out) return fail(BDOF_E_BADARG, "gradient accumulation needs d_grad_out (the accumulator)");
        if (!use_sweep(p)) return fail(BDOF_E_UNSUPPORTED, "fused gradient accumulation is a feature of the sweep / resident kernels");
        if (p->win_origin) return fail(BDOF_E_UNSUPPORTED, "window mode writes per-window gradients; accumulate them with bdof_patch_gather_add");
        if (p->stash_valid && p->t_stash == reinterpret_cast<float2*>(d_grad_out))
            return fail(BDOF_E_STATE, "the transmission stash must not live in the accumulator");
    }
    if (p->win_origin && !use_resident(p)) return fail(BDOF_E_UNSUPPORTED, "window mode needs the resident small-field kernels");
    float2* db = reinterpret_cast<float2*>(d_db_inout);
    float2* gout = reinterpret_cast<float2*>(d_grad_out);
    float2* G = p->work[0];
    const int Z = p->n_slice;
    // psi_out = c * psi'_out with |c| = 1 (global phase kept out of the fp32 chain): G' = conj(c) G
    const std::complex<double> c = std::conj(p->total_phase);
    const unsigned long long launches0 = g_launches.load();
    CUDA_TRY(cudaEventRecord(p->t_ev[2], p->stream));
    k_scale_complex<<<blocks_for(p->F, 256), 256, 0, p->stream>>>(reinterpret_cast<const float2*>(d_grad_exit), G, p->F,
                                                                  float(c.real()), float(c.imag()));
    BDOF_TRY(launch_check("k_scale_complex"));
    if (p->free_mode == BDOF_FREE_INF) {
        // adjoint of fftshift(fft2(.)): unnormalised inverse transform of ifftshift(G)
        LineParams cc = col_params(p, G, p->tmp, nullptr);
        cc.in_shift = p->ny / 2;
        BDOF_TRY(col_pass(p, V_COL_INV, cc));
        LineParams r = row_params(p, p->tmp, G, nullptr);
        r.in_shift = p->nx / 2;
        BDOF_TRY(row_pass(p, V_ROW_INV, r));
    } else if (p->free_mode == BDOF_FREE_TF) {
        BDOF_TRY(col_pass(p, V_COL_CONV, col_params(p, G, p->tmp, p->ay.hf_adj)));
        BDOF_TRY(row_pass(p, V_ROW_CONV, row_params(p, p->tmp, G, p->ax.hf_adj)));
    }
    const int n_buckets = int(p->bucket_events.size());
    const bool resident = use_resident(p);
    const bool sweep = use_sweep(p) && !resident;
    if (resident) {
        ResidentParams q{};
        q.in = G; q.out = d_grad_probe ? G : nullptr;
        q.db = db; q.db_slice_stride = zb ? 0 : p->F;
        q.tstash = p->stash_valid ? p->t_stash : nullptr;
        q.grad = gout ? gout : db;
        q.accumulate = p->grad_accumulate ? 1 : 0;
        if (p->win_origin) {
            q.win = p->win_origin; q.oy = p->win_oy; q.ox = p->win_ox;
            q.db_slice_stride = zb ? 0 : (long long)p->win_oy * p->win_ox;
        }
        BDOF_TRY(resident_launch(p, true, q));
        for (int j = 0; j < n_buckets; ++j) CUDA_TRY(cudaEventRecord(p->bucket_events[j], p->stream));
        if (p->stash_valid && (p->t_stash == (gout ? gout : db))) p->stash_valid = false;
    }
    if (sweep) {
        if (slice_propagates(p, Z - 1)) {
            // adjoint of the trailing half propagation (TF semantics), into the work field the sweep runs on
            if (Z & 1) BDOF_TRY(col_pass(p, V_COL_CONV, col_params(p, G, p->tmp, h_entry(p->ay, Z, true))));
            else       BDOF_TRY(row_pass(p, V_ROW_CONV, row_params(p, G, p->tmp, h_entry(p->ax, Z, true))));
            G = p->tmp;
        }
    }
    for (int i = Z - 1; i >= 0 && sweep; --i) {
        SweepParams q{};
        q.in = G; q.out = G;
        q.db = db + (zb ? 0 : (long long)i * p->F);
        if (p->stash_valid) { q.db = p->t_stash + (long long)i * p->F; q.db_is_t = 1; }
        q.grad = gout ? gout + (long long)i * p->F : db + (long long)i * p->F;
        q.slab = p->slabs + (long long)i * p->F;
        q.conv1 = slice_propagates(p, i); q.conv2 = (i > 0);
        q.grad_accumulate = p->grad_accumulate ? 1 : 0;
        q.store_out = (i > 0 || d_grad_probe != nullptr);
        BDOF_TRY(sweep_launch(p, i, true, q));
        if (n_buckets > 0) {
            const int per = (Z + n_buckets - 1) / n_buckets;
            const int from_top = Z - 1 - i;
            if ((from_top + 1) % per == 0 || i == 0) {
                const int j = from_top / per;
                if (j < n_buckets) CUDA_TRY(cudaEventRecord(p->bucket_events[j], p->stream));
            }
        }
    }
    for (int i = Z - 1; i >= 0 && !sweep && !resident; --i) {
        const float2* db_i = db + (zb ? 0 : (long long)i * p->F);
        float2* grad_i = gout ? gout + (long long)i * p->F : db + (long long)i * p->F;
        const float2* psi_i = p->slabs + (long long)i * p->F;
        if (slice_propagates(p, i)) {
            BDOF_TRY(propagate_slice_adj(p, G, psi_i, db_i, grad_i, i));
        } else {
            k_modulate_adj<<<blocks_for(p->F, 256), 256, 0, p->stream>>>(G, psi_i, db_i, grad_i, p->F, float(p->k_dz));
            BDOF_TRY(launch_check("k_modulate_adj"));
        }
        if (n_buckets > 0) {
            // bucket j covers slices [z_lo(j), z_lo(j-1)) counted from the top: it is final once slice z_lo(j) is done
            const int per = (Z + n_buckets - 1) / n_buckets;
            const int from_top = Z - 1 - i;
            if ((from_top + 1) % per == 0 || i == 0) {
                const int j = from_top / per;
                if (j < n_buckets) CUDA_TRY(cudaEventRecord(p->bucket_events[j], p->stream));
            }
        }
    }
    if (sweep && p->stash_valid && (p->t_stash == (gout ? gout : db))) p->stash_valid = false;      // t has been replaced by the gradient
    if (d_grad_probe) {
        const long long per = (long long)p->ny * p->nx;
        k_sum_batch<<<blocks_for(per, 256), 256, 0, p->stream>>>(G, reinterpret_cast<float2*>(d_grad_probe), per, p->batch);
        BDOF_TRY(launch_check("k_sum_batch"));
    }
    CUDA_TRY(cudaEventRecord(p->t_ev[3], p->stream));
    p->t_launches[1] = int(g_launches.load() - launches0);
    return 0;
}

// ------------------------------------------------------------------------------------------
// layout conversion, ptychography windows, real-space propagator
// ------------------------------------------------------------------------------------------
extern "C" int bdof_pack_db(const float* d_delta, const float* d_beta, float* d_db, int batch, int ny, int nx, int n_slice, void* st) {
    if (!d_delta || !d_beta || !d_db || batch < 1 || ny < 1 || nx < 1 || n_slice < 1) return fail(BDOF_E_BADARG, "bad argument");
    const long long rows = (long long)batch * ny;
    if (rows > 65535LL * 1024) return fail(BDOF_E_UNSUPPORTED, "too many rows");
    // gridDim.z <= 65535: fold rows
    for (long long r0 = 0; r0 < rows; r0 += 65535) {
        const int nr = int(std::min<long long>(65535, rows - r0));
        dim3 grid((nx + 31) / 32, (n_slice + 31) / 32, nr), block(32, 8);
        k_pack_db<<<grid, block, 0, (cudaStream_t)st>>>(d_delta + r0 * nx * n_slice, d_beta + r0 * nx * n_slice,
                                                      reinterpret_cast<float2*>(d_db) + r0 * nx, int(rows), nx, n_slice);
        BDOF_TRY(launch_check("k_pack_db"));
    }
    return 0;
}
// row-chunked variants for pipelined host transfers: the chunk holds rows [row0, row0 + n_rows) of the B*Y rows
extern "C" int bdof_pack_db_rows(const float* d_delta_chunk, const float* d_beta_chunk, float* d_db, long long total_rows, long long row0,
                                 int n_rows, int nx, int n_slice, void* st) {
    if (!d_delta_chunk || !d_beta_chunk || !d_db || total_rows < 1 || row0 < 0 || n_rows < 1 || row0 + n_rows > total_rows || nx < 1 || n_slice < 1)
        return fail(BDOF_E_BADARG, "bad argument");
    if (total_rows > 0x7fffffffLL) return fail(BDOF_E_UNSUPPORTED, "too many rows");
    for (int r0 = 0; r0 < n_rows; r0 += 65535) {
        const int nr = std::min(65535, n_rows - r0);
        dim3 grid((nx + 31) / 32, (n_slice + 31) / 32, nr), block(32, 8);
        k_pack_db<<<grid, block, 0, (cudaStream_t)st>>>(d_delta_chunk + (long long)r0 * nx * n_slice, d_beta_chunk + (long long)r0 * nx * n_slice,
                                                      reinterpret_cast<float2*>(d_db) + (row0 + r0) * nx, int(total_rows), nx, n_slice);
        BDOF_TRY(launch_check("k_pack_db"));
    }
    return 0;
}
extern "C" int bdof_unpack_db_rows(const float* d_db, float* d_delta_chunk, float* d_beta_chunk, long long total_rows, long long row0,
                                   int n_rows, int nx, int n_slice, void* st) {
    if (!d_delta_chunk || !d_beta_chunk || !d_db || total_rows < 1 || row0 < 0 || n_rows < 1 || row0 + n_rows > total_rows || nx < 1 || n_slice < 1)
        return fail(BDOF_E_BADARG, "bad argument");
    if (total_rows > 0x7fffffffLL) return fail(BDOF_E_UNSUPPORTED, "too many rows");
    for (int r0 = 0; r0 < n_rows; r0 += 65535) {
        const int nr = std::min(65535, n_rows - r0);
        dim3 grid((nx + 31) / 32, (n_slice + 31) / 32, nr), block(32, 8);
        k_unpack_db<<<grid, block, 0, (cudaStream_t)st>>>(reinterpret_cast<const float2*>(d_db) + (row0 + r0) * nx,
                                                        d_delta_chunk + (long long)r0 * nx * n_slice, d_beta_chunk + (long long)r0 * nx * n_slice,
                                                        int(total_rows), nx, n_slice);
        BDOF_TRY(launch_check("k_unpack_db"));
    }
    return 0;
}

extern "C" int bdof_unpack_db(const float* d_db, float* d_delta, float* d_beta, int batch, int ny, int nx, int n_slice, void* st) {
    if (!d_delta || !d_beta || !d_db || batch < 1 || ny < 1 || nx < 1 || n_slice < 1) return fail(BDOF_E_BADARG, "bad argument");
    const long long rows = (long long)batch * ny;
    for (long long r0 = 0; r0 < rows; r0 += 65535) {
        const int nr = int(std::min<long long>(65535, rows - r0));
        dim3 grid((nx + 31) / 32, (n_slice + 31) / 32, nr), block(32, 8);
        k_unpack_db<<<grid, block, 0, (cudaStream_t)st>>>(reinterpret_cast<const float2*>(d_db) + r0 * nx,
                                                        d_delta + r0 * nx * n_slice, d_beta + r0 * nx * n_slice, int(rows), nx, n_slice);
        BDOF_TRY(launch_check("k_unpack_db"));
    }
    return 0;
}

extern "C" int bdof_patch_gather(const float* d_db_obj, int n_slice, int oy, int ox, const int* d_pos_yx, int n_pos,
                                 int py, int px, float* d_db_patches, void* st) {
    if (!d_db_obj || !d_pos_yx || !d_db_patches || n_pos < 1 || n_slice < 1) return fail(BDOF_E_BADARG, "bad argument");
    if (n_pos > 65535 || n_slice > 65535) return fail(BDOF_E_UNSUPPORTED, "n_pos / n_slice > 65535");
    dim3 grid((py * px + 255) / 256, n_pos, n_slice);
    k_patch_gather<<<grid, 256, 0, (cudaStream_t)st>>>(reinterpret_cast<const float2*>(d_db_obj), oy, ox, d_pos_yx, n_pos, py, px,
                                                     reinterpret_cast<float2*>(d_db_patches));
    return launch_check("k_patch_gather");
}
extern "C" int bdof_patch_scatter_add(const float* d_grad_patches, int n_slice, int oy, int ox, const int* d_pos_yx, int n_pos,
                                      int py, int px, float* d_grad_obj, void* st) {
    if (!d_grad_patches || !d_pos_yx || !d_grad_obj || n_pos < 1 || n_slice < 1) return fail(BDOF_E_BADARG, "bad argument");
    if (n_pos > 65535 || n_slice > 65535) return fail(BDOF_E_UNSUPPORTED, "n_pos / n_slice > 65535");
    dim3 grid((py * px + 255) / 256, n_pos, n_slice);
    k_patch_scatter_add<<<grid, 256, 0, (cudaStream_t)st>>>(reinterpret_cast<const float2*>(d_grad_patches), oy, ox, d_pos_yx, n_pos,
                                                          py, px, reinterpret_cast<float2*>(d_grad_obj));
    return launch_check("k_patch_scatter_add");
}

extern "C" int bdof_patch_gather_add(const float* d_grad_patches, int n_slice, int oy, int ox, const int* d_pos_yx, int n_pos, int py, int px,
                                     float* d_grad_obj, void* st) {
    if (!d_grad_patches || !d_pos_yx || !d_grad_obj || n_pos < 1 || n_slice < 1) return fail(BDOF_E_BADARG, "bad argument");
    if (n_pos > PSCAT_MAX_POS) return fail(BDOF_E_UNSUPPORTED, "more than %d scan positions per call", PSCAT_MAX_POS);
    if (oy > 65535 || (n_slice + PSCAT_ZB - 1) / PSCAT_ZB > 65535) return fail(BDOF_E_UNSUPPORTED, "object too large");
    if (oy > 30000 || ox > 30000 || py > 30000 || px > 30000) return fail(BDOF_E_UNSUPPORTED, "object or window side > 30000");
    dim3 grid((ox + PSCAT_BX - 1) / PSCAT_BX, oy, (n_slice + PSCAT_ZB - 1) / PSCAT_ZB);
    k_patch_gather_add<<<grid, PSCAT_THREADS, 0, (cudaStream_t)st>>>(reinterpret_cast<const float2*>(d_grad_patches), n_slice, oy, ox, d_pos_yx,
                                                                    n_pos, py, px, reinterpret_cast<float2*>(d_grad_obj));
    return launch_check("k_patch_gather_add");
}

extern "C" int bdof_cnn_forward(const float* d_db, const float* d_probe, float* d_exit, float* d_work, int batch, int ny, int nx,
                                int n_slice, const double* h_kernel, int ks, double k_dz, void* st_) {
    if (!d_db || !d_probe || !d_exit || !d_work || !h_kernel) return fail(BDOF_E_BADARG, "null");
    if (ks < 1 || ks % 2 == 0 || ks > CNN_MAX_K) return fail(BDOF_E_BADARG, "kernel_size must be odd and <= %d", CNN_MAX_K);
    cudaStream_t st = (cudaStream_t)st_;
    std::vector<float2> kf((size_t)ks * ks);
    std::complex<double> ksum = 0.0;
    for (int i = 0; i < ks * ks; ++i) {
        kf[i] = make_float2(float(h_kernel[2 * i]), float(h_kernel[2 * i + 1]));
        ksum += std::complex<double>(h_kernel[2 * i], h_kernel[2 * i + 1]);
    }
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_cnn_kernel, kf.data(), kf.size() * sizeof(float2), 0, cudaMemcpyHostToDevice, st));
    const long long per = (long long)ny * nx, F = per * batch;
    float2* bufs[2] = {reinterpret_cast<float2*>(d_work), reinterpret_cast<float2*>(d_work) + F};
    k_broadcast_probe<<<blocks_for(per, 256), 256, 0, st>>>(reinterpret_cast<const float2*>(d_probe), bufs[0], per, batch);
    BDOF_TRY(launch_check("k_broadcast_probe"));
    constexpr int TILE = 16;
    const size_t smem = size_t(TILE + ks - 1) * (TILE + ks - 1) * sizeof(float2);
    dim3 grid((nx + TILE - 1) / TILE, (ny + TILE - 1) / TILE, batch), block(TILE, TILE);
    std::complex<double> edge = 1.0;
    int cur = 0;
    for (int i = 0; i < n_slice; ++i) {
        float2* dst = (i == n_slice - 1) ? reinterpret_cast<float2*>(d_exit) : bufs[cur ^ 1];
        k_cnn_step<TILE><<<grid, block, smem, st>>>(bufs[cur], reinterpret_cast<const float2*>(d_db) + (long long)i * F, dst, ny, nx,
                                                   ks, float(k_dz), make_float2(float(edge.real()), float(edge.imag())));
        BDOF_TRY(launch_check("k_cnn_step"));
        edge *= ksum;                         // propagation.py:99
        cur ^= 1;
    }
    return 0;
}

static int cnn_upload_kernel(const double* h_kernel, int ks, cudaStream_t st, std::complex<double>* ksum_out) {
    if (ks < 1 || ks % 2 == 0 || ks > CNN_MAX_K) return fail(BDOF_E_BADARG, "kernel_size must be odd and <= %d", CNN_MAX_K);
    std::vector<float2> kf((size_t)ks * ks);
    std::complex<double> ksum = 0.0;
    for (int i = 0; i < ks * ks; ++i) {
        kf[i] = make_float2(float(h_kernel[2 * i]), float(h_kernel[2 * i + 1]));
        ksum += std::complex<double>(h_kernel[2 * i], h_kernel[2 * i + 1]);
    }
    // synchronous with respect to the host buffer (kf dies at return)
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_cnn_kernel, kf.data(), kf.size() * sizeof(float2), 0, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (ksum_out) *ksum_out = ksum;
    return 0;
}

// forward that keeps the field entering every slice: d_slices [n_slice + 1][batch][ny][nx], slices[0] = the probe, slices[i + 1] =
// the field after slice i; the (unscaled) chain output is slices[n_slice]
extern "C" int bdof_cnn_forward_store(const float* d_db, const float* d_probe, float* d_slices, int batch, int ny, int nx, int n_slice,
                                      const double* h_kernel, int ks, double k_dz, void* st_) {
    if (!d_db || !d_probe || !d_slices || !h_kernel || batch < 1 || ny < 1 || nx < 1 || n_slice < 1) return fail(BDOF_E_BADARG, "bad argument");
    cudaStream_t st = (cudaStream_t)st_;
    std::complex<double> ksum;
    BDOF_TRY(cnn_upload_kernel(h_kernel, ks, st, &ksum));
    const long long per = (long long)ny * nx, F = per * batch;
    float2* sl = reinterpret_cast<float2*>(d_slices);
    k_broadcast_probe<<<blocks_for(per, 256), 256, 0, st>>>(reinterpret_cast<const float2*>(d_probe), sl, per, batch);
    BDOF_TRY(launch_check("k_broadcast_probe"));
    constexpr int TILE = 16;
    const size_t smem = size_t(TILE + ks - 1) * (TILE + ks - 1) * sizeof(float2);
    dim3 grid((nx + TILE - 1) / TILE, (ny + TILE - 1) / TILE, batch), block(TILE, TILE);
    std::complex<double> edge = 1.0;
    for (int i = 0; i < n_slice; ++i) {
        k_cnn_step<TILE><<<grid, block, smem, st>>>(sl + (long long)i * F, reinterpret_cast<const float2*>(d_db) + (long long)i * F,
                                                   sl + (long long)(i + 1) * F, ny, nx, ks, float(k_dz),
                                                   make_float2(float(edge.real()), float(edge.imag())));
        BDOF_TRY(launch_check("k_cnn_step"));
        edge *= ksum;
    }
    return 0;
}

// back-propagate d_G (gradient w.r.t. the unscaled chain output slices[n_slice]; [2][batch][ny][nx], the second field is work
// space) through the stored chain: d_grad_out [n_slice][batch][ny][nx][2]; on return d_G[0] holds the gradient w.r.t. the probe
// broadcast (per batch element)
extern "C" int bdof_cnn_adjoint(const float* d_db, const float* d_slices, float* d_G, float* d_grad_out, int batch, int ny, int nx, int n_slice,
                                const double* h_kernel, int ks, double k_dz, void* st_) {
    if (!d_db || !d_slices || !d_G || !d_grad_out || !h_kernel || batch < 1 || ny < 1 || nx < 1 || n_slice < 1) return fail(BDOF_E_BADARG, "bad argument");
    cudaStream_t st = (cudaStream_t)st_;
    BDOF_TRY(cnn_upload_kernel(h_kernel, ks, st, nullptr));
    const long long F = (long long)ny * nx * batch;
    const float2* sl = reinterpret_cast<const float2*>(d_slices);
    float2* G[2] = {reinterpret_cast<float2*>(d_G), reinterpret_cast<float2*>(d_G) + F};
    constexpr int TILE = 16;
    const size_t smem = size_t(TILE + ks - 1) * (TILE + ks - 1) * sizeof(float2);
    dim3 grid((nx + TILE - 1) / TILE, (ny + TILE - 1) / TILE, batch), block(TILE, TILE);
    int cur = 0;
    for (int i = n_slice - 1; i >= 0; --i) {
        k_cnn_step_adj<TILE><<<grid, block, smem, st>>>(G[cur], sl + (long long)i * F, reinterpret_cast<const float2*>(d_db) + (long long)i * F,
                                                       reinterpret_cast<float2*>(d_grad_out) + (long long)i * F, G[cur ^ 1], ny, nx, ks, float(k_dz));
        BDOF_TRY(launch_check("k_cnn_step_adj"));
        cur ^= 1;
    }
    if (cur != 0) CUDA_TRY(cudaMemcpyAsync(G[0], G[1], (size_t)F * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int bdof_forward_host(bdof_plan* p, const float* h_delta, const float* h_beta, const float* h_probe, float* h_exit) {
    if (!p || !h_delta || !h_beta || !h_probe || !h_exit) return fail(BDOF_E_BADARG, "null");
    const size_t vol = (size_t)p->F * p->n_slice;
    const size_t per = (size_t)p->ny * p->nx;
    if (!p->e2e_delta) {
        CUDA_TRY(cudaMalloc((void**)&p->e2e_delta, vol * sizeof(float)));
        CUDA_TRY(cudaMalloc((void**)&p->e2e_beta, vol * sizeof(float)));
        CUDA_TRY(cudaMalloc((void**)&p->e2e_db, vol * sizeof(float2)));
        CUDA_TRY(cudaMalloc((void**)&p->e2e_probe, per * sizeof(float2)));
        CUDA_TRY(cudaMalloc((void**)&p->e2e_exit, (size_t)p->F * sizeof(float2)));
    }
    CUDA_TRY(cudaMemcpyAsync(p->e2e_delta, h_delta, vol * sizeof(float), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->e2e_beta, h_beta, vol * sizeof(float), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->e2e_probe, h_probe, per * sizeof(float2), cudaMemcpyHostToDevice, p->stream));
    BDOF_TRY(bdof_pack_db(p->e2e_delta, p->e2e_beta, reinterpret_cast<float*>(p->e2e_db), p->batch, p->ny, p->nx, p->n_slice, p->stream));
    BDOF_TRY(bdof_forward(p, reinterpret_cast<const float*>(p->e2e_db), reinterpret_cast<const float*>(p->e2e_probe),
                          reinterpret_cast<float*>(p->e2e_exit)));
    CUDA_TRY(cudaMemcpyAsync(h_exit, p->e2e_exit, (size_t)p->F * sizeof(float2), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    // the staging volumes are as large as the object (17 GB at 2048^2 x 256): do not keep them for the life of the plan
    cudaFree(p->e2e_delta); cudaFree(p->e2e_beta); cudaFree(p->e2e_db); cudaFree(p->e2e_probe); cudaFree(p->e2e_exit);
    p->e2e_delta = p->e2e_beta = nullptr; p->e2e_db = nullptr; p->e2e_probe = p->e2e_exit = nullptr;
    return 0;
}

// out[b] = in[b] * m   (complex64; m [n_per] shared by the batch) -- the kernel multiply of the IR free-space step
__global__ void k_field_multiply(const float2* __restrict__ in, const float2* __restrict__ m, float2* __restrict__ out, long long n_per,
                                 long long n_total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_total; i += (long long)gridDim.x * blockDim.x)
        out[i] = cmul(in[i], m[i % n_per]);
}
extern "C" int bdof_field_multiply(const float* d_in, const float* d_mult, float* d_out, int batch, long long n_per, void* st) {
    if (!d_in || !d_mult || !d_out || batch < 1 || n_per < 1) return fail(BDOF_E_BADARG, "bad argument");
    k_field_multiply<<<LOSS_BLOCKS, 256, 0, (cudaStream_t)st>>>(reinterpret_cast<const float2*>(d_in), reinterpret_cast<const float2*>(d_mult),
                                                              reinterpret_cast<float2*>(d_out), n_per, n_per * batch);
    return launch_check("k_field_multiply");
}

extern "C" int bdof_debug_fft_gain(int n, double* gain_out) {
    if (n < 1 || !gain_out) return fail(BDOF_E_BADARG, "bad argument");
    const std::vector<cdbl> g = bdof_size_supported(n) == 1 ? conv_diag_gain(n) : generic_diag_gain(n);
    for (int k = 0; k < n; ++k) { gain_out[2 * k] = g[k].real(); gain_out[2 * k + 1] = g[k].imag(); }
    return 0;
}

extern "C" int bdof_plan_set_stream(bdof_plan* p, void* cuda_stream) {
    if (!p) return fail(BDOF_E_BADARG, "null");
    p->stream = (cudaStream_t)cuda_stream;
    return 0;
}

// ------------------------------------------------------------------------------------------
// the steps either side of the hot path (cnn_propagator drivers)
// ------------------------------------------------------------------------------------------
// obj_rot[z][y][x] = obj[z_old(z,x)][y][x_old(z,x)]   (apply_rotation, cnn_propagator/util.py:374-402; the table is the
// reference's nearest-neighbour lookup of one angle re-ordered slice-major: lookup[z][x] = (x_old, z_old))
// The rotation is about y, so a run of x in the rotated frame reads a run along an oblique line of the (z, x) plane: near 90
// degrees consecutive x read consecutive SLICES (8 useful bytes per 32-byte sector).  Both directions therefore go through
// shared memory: a CTA owns a ROT_T x ROT_T tile of (z, x) cells, finds where the cells on the other side of the table lie (a
// box of at most ROT_T (|cos| + |sin|) + 2 <= ROT_BOX on a side), and for every y of its chunk ONE TMA box load
// (cp.async.bulk.tensor, ROT_BOX x 1 x ROT_BOX of the [z][y][x] tensor, double buffered on two mbarriers) brings the box in; the
// tile is then served from shared memory with row-contiguous writes.  Anything outside the box -- clipped table entries at the
// borders, arbitrary user tables -- is read from global memory directly, so the result never depends on the tiling; sides that
// TMA cannot describe (odd nx, sides below the box) load the box with ordinary row-contiguous reads (TMA = false).  A tiled TMA
// load faults ("illegal instruction") unless its innermost start coordinate is a multiple of 16 bytes (tools/tma_probe.cu), so
// boxes start on even x.
constexpr int ROT_T = 32, ROT_BOX = 48, ROT_THREADS = 256, ROT_CPT = ROT_T * ROT_T / ROT_THREADS, ROT_YB = 8, ROT_YA = 4;   // rows of y per CTA: gather, transpose
constexpr int ROT_BPT = ROT_BOX * ROT_BOX / ROT_THREADS;       // box elements per thread
constexpr unsigned ROT_BOX_BYTES = ROT_BOX * ROT_BOX * sizeof(float2);
static_assert(ROT_BOX * ROT_BOX % ROT_THREADS == 0, "box elements divide over the threads");

// element idx = tid + ROT_THREADS i of the box is (r, c) = (idx / ROT_BOX, idx % ROT_BOX): a warp covers 32 consecutive
// elements of at most two rows
__device__ __forceinline__ void rot_box_load(const float2* __restrict__ origin, long long row_stride, int bh, int bw, int tid, float2 (&r)[ROT_BPT]) {
#pragma unroll
    for (int i = 0; i < ROT_BPT; ++i) {
        const int idx = tid + ROT_THREADS * i, rr = idx / ROT_BOX, cc = idx - rr * ROT_BOX;
        r[i] = (rr < bh && cc < bw) ? __ldg(origin + (long long)rr * row_stride + cc) : make_float2(0.f, 0.f);
    }
}
__device__ __forceinline__ void rot_box_store(float2* __restrict__ buf, int tid, const float2 (&r)[ROT_BPT]) {
#pragma unroll
    for (int i = 0; i < ROT_BPT; ++i) buf[tid + ROT_THREADS * i] = r[i];
}
__device__ __forceinline__ void tma_load_3d(void* dst_smem, const CUtensorMap* tm, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst_smem, const CUtensorMap* tm, int c0, int c1, int c2, int c3, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}

static bool rot_use_tma() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("BDOF_ROT_TMA"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
template <bool TMA>
__global__ void __launch_bounds__(ROT_THREADS) k_rotate_gather(const float2* __restrict__ obj, const int2* __restrict__ lookup, float2* __restrict__ out,
                                long long out_slice_stride, int ny, int nx, int nz, const __grid_constant__ CUtensorMap tm_obj) {
    __shared__ __align__(128) float2 buf[2][ROT_BOX * ROT_BOX];
    __shared__ __align__(8) unsigned long long bar[2];
    __shared__ int s_mm[4];                              // z_min, x_min, z_max, x_max of the sources
    const int tid = threadIdx.x;
    // y chunks are the fastest grid dimension: the CTAs resident at one time then span whole 2-D slabs
    const int y0 = blockIdx.x * ROT_YB, y1 = min(ny, y0 + ROT_YB), x0 = blockIdx.y * ROT_T, z0 = blockIdx.z * ROT_T;
    if (tid == 0) {
        s_mm[0] = s_mm[1] = 0x7fffffff; s_mm[2] = s_mm[3] = -1;
        if constexpr (TMA) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    }
    __syncthreads();
    int2 src[ROT_CPT];
    bool ok[ROT_CPT];
    int zmn = 0x7fffffff, xmn = 0x7fffffff, zmx = -1, xmx = -1;
#pragma unroll
    for (int i = 0; i < ROT_CPT; ++i) {
        const int c = tid + ROT_THREADS * i, z = z0 + (c >> 5), x = x0 + (c & 31);
        ok[i] = z < nz && x < nx;
        src[i] = ok[i] ? lookup[(long long)z * nx + x] : make_int2(0, 0);
        if (ok[i]) { zmn = min(zmn, src[i].y); zmx = max(zmx, src[i].y); xmn = min(xmn, src[i].x); xmx = max(xmx, src[i].x); }
    }
    zmn = __reduce_min_sync(0xffffffffu, zmn); xmn = __reduce_min_sync(0xffffffffu, xmn);
    zmx = __reduce_max_sync(0xffffffffu, zmx); xmx = __reduce_max_sync(0xffffffffu, xmx);
    if ((tid & 31) == 0) { atomicMin(&s_mm[0], zmn); atomicMin(&s_mm[1], xmn); atomicMax(&s_mm[2], zmx); atomicMax(&s_mm[3], xmx); }
    __syncthreads();
    zmn = s_mm[0]; xmn = s_mm[1] & ~1;                   // TMA boxes start on 16 bytes of the innermost dimension: even x
    const int bh = s_mm[2] - zmn + 1, bw = s_mm[3] - xmn + 1;
    if (s_mm[2] < 0) return;                             // no cell of this tile is inside the array
    float2* o[ROT_CPT];
#pragma unroll
    for (int i = 0; i < ROT_CPT; ++i) {
        const int c = tid + ROT_THREADS * i;
        o[i] = out + (long long)(z0 + (c >> 5)) * out_slice_stride + x0 + (c & 31);
    }
    if (bh > ROT_BOX || bw > ROT_BOX) {                  // not a rotation-like table: plain gather
        for (int y = y0; y < y1; ++y)
#pragma unroll
            for (int i = 0; i < ROT_CPT; ++i)
                if (ok[i]) __stcs(o[i] + (long long)y * nx, __ldg(obj + ((long long)src[i].y * ny + y) * nx + src[i].x));
        return;
    }
    int soff[ROT_CPT];
#pragma unroll
    for (int i = 0; i < ROT_CPT; ++i) soff[i] = (src[i].y - zmn) * ROT_BOX + (src[i].x - xmn);
    if constexpr (TMA) {
        if (tid == 0) { mbar_expect_tx(&bar[0], ROT_BOX_BYTES); tma_load_3d(buf[0], &tm_obj, 2 * xmn, y0, zmn, &bar[0]); }
        for (int y = y0; y < y1; ++y) {
            const int it = y - y0, cur = it & 1;
            if (tid == 0 && y + 1 < y1) { mbar_expect_tx(&bar[cur ^ 1], ROT_BOX_BYTES); tma_load_3d(buf[cur ^ 1], &tm_obj, 2 * xmn, y + 1, zmn, &bar[cur ^ 1]); }
            mbar_wait(&bar[cur], (it >> 1) & 1);
#pragma unroll
            for (int i = 0; i < ROT_CPT; ++i)
                if (ok[i]) __stcs(o[i] + (long long)y * nx, buf[cur][soff[i]]);
            __syncthreads();                             // buf[cur] is free for the load of y + 2
        }
    } else {
        const float2* origin = obj + (long long)zmn * ny * nx + xmn;      // box element (r, c) of row y: origin + (r ny + y) nx + c
        const long long row_stride = (long long)ny * nx;
        float2 r[ROT_BPT];
        rot_box_load(origin + (long long)y0 * nx, row_stride, bh, bw, tid, r);
        rot_box_store(buf[0], tid, r);
        __syncthreads();
        for (int y = y0; y < y1; ++y) {
            const int cur = (y - y0) & 1;
            if (y + 1 < y1) rot_box_load(origin + (long long)(y + 1) * nx, row_stride, bh, bw, tid, r);
#pragma unroll
            for (int i = 0; i < ROT_CPT; ++i)
                if (ok[i]) __stcs(o[i] + (long long)y * nx, buf[cur][soff[i]]);
            if (y + 1 < y1) rot_box_store(buf[cur ^ 1], tid, r);
            __syncthreads();
        }
    }
}
// transpose of the gather (what autograd does to the fancy index): grad_obj[z_old][y][x_old] += grad_rot[z][y][x]
__global__ void k_rotate_scatter_add(const float2* __restrict__ grot, long long slice_stride, const int2* __restrict__ lookup,
                                     float2* __restrict__ gobj, int ny, int nx, int nz) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, z = blockIdx.z;
    if (x >= nx) return;
    const int2 src = lookup[(long long)z * nx + x];
    const float2 g = grot[(long long)z * slice_stride + (long long)y * nx + x];
    float* dst = reinterpret_cast<float*>(gobj + ((long long)src.y * ny + y) * nx + src.x);
    atomicAdd(dst, g.x);
    atomicAdd(dst + 1, g.y);
}
// the same transpose without atomics: every source pixel (z0, x0) sums the rotated pixels that read from it (CSR lists
// built on the host from the lookup table; deterministic, and ~10x faster than fp32 atomics to scattered addresses)
// `n_ang` angles at once: element a of the minibatch has its lists at offsets[a] / dest[a] and its rotated gradient at
// grot + a * batch_stride; the sum over angles stays in registers (angles in order, list entries in order: bit-reproducible), so
// gobj is read and written once.  Tiled like the gather: the CTA owns ROT_T x ROT_T SOURCE cells; per angle the box is centred
// on the mean position of the cells' first readers; list entries outside it (the long lists of clipped border cells) and
// entries beyond the second of a cell are read from global memory.
struct RotLists { const int* offsets[BDOF_ROT_MAX_ANGLES]; const int* dest[BDOF_ROT_MAX_ANGLES]; };
// Cells with more than two readers are the clipped border cells of the table (rotation_lookup clips the source coordinates, so
// at 45 degrees 17 % of the rotated pixels pile onto the border cells, hundreds on the corner cells): their tails go to a
// per-CTA work list that the warps sum cooperatively -- lanes stride the list, fixed shuffle tree, so still bit-reproducible.
// (One thread walking such a list serially made the back-rotation 2.0-2.6 ms at 30-60 degrees against 0.9 ms below 30.)
constexpr int ROT_LONG = 160;
__global__ void __launch_bounds__(ROT_THREADS, 2) k_rotate_adjoint_csr(const float2* __restrict__ grot, long long slice_stride, long long batch_stride,
                                     const RotLists lists, int n_ang, int accumulate, float2* __restrict__ gobj, int ny, int nx, int nz, int z_tile0) {
    __shared__ float2 buf[2][ROT_BOX * ROT_BOX];
    __shared__ float2 s_extra[ROT_LONG][ROT_YA];
    __shared__ int s_long_beg[ROT_LONG], s_long_cnt[ROT_LONG];
    __shared__ int s_sum[4];                             // sum of z, sum of x, count over the first readers; number of long cells
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int y0 = blockIdx.x * ROT_YA, x0 = blockIdx.y * ROT_T, z0 = (blockIdx.z + z_tile0) * ROT_T;      // y fastest, as the gather
    const int nyc = min(ROT_YA, ny - y0);
    float2 acc[ROT_CPT][ROT_YA];
    bool ok[ROT_CPT];
    int cell[ROT_CPT];
#pragma unroll
    for (int i = 0; i < ROT_CPT; ++i) {
        const int c = tid + ROT_THREADS * i, z = z0 + (c >> 5), x = x0 + (c & 31);
        ok[i] = z < nz && x < nx;
        cell[i] = ok[i] ? z * nx + x : 0;
#pragma unroll
        for (int j = 0; j < ROT_YA; ++j) acc[i][j] = make_float2(0.f, 0.f);
    }
    for (int a = 0; a < n_ang; ++a) {
        const int* __restrict__ off = lists.offsets[a];
        const int* __restrict__ dst = lists.dest[a];
        const float2* __restrict__ base = grot + (long long)a * batch_stride;
        if (tid == 0) s_sum[0] = s_sum[1] = s_sum[2] = s_sum[3] = 0;
        __syncthreads();
        int beg[ROT_CPT], cnt[ROT_CPT], slot[ROT_CPT];
        int sz = 0, sx = 0, sc = 0;
#pragma unroll
        for (int i = 0; i < ROT_CPT; ++i) {
            beg[i] = ok[i] ? off[cell[i]] : 0;
            cnt[i] = ok[i] ? off[cell[i] + 1] - beg[i] : 0;
            slot[i] = -1;
            if (cnt[i] > 0 && cnt[i] <= 2) { const int d = dst[beg[i]], z = d / nx; sz += z; sx += d - z * nx; ++sc; }    // ordinary cells centre the box
            if (cnt[i] > 2) {
                const int sl = atomicAdd(&s_sum[3], 1);              // slot order is arbitrary; every cell's sum is its own
                if (sl < ROT_LONG) { slot[i] = sl; s_long_beg[sl] = beg[i] + 2; s_long_cnt[sl] = cnt[i] - 2; }
            }
        }
        sz = __reduce_add_sync(0xffffffffu, sz); sx = __reduce_add_sync(0xffffffffu, sx); sc = __reduce_add_sync(0xffffffffu, sc);
        if (lane == 0 && sc > 0) { atomicAdd(&s_sum[0], sz); atomicAdd(&s_sum[1], sx); atomicAdd(&s_sum[2], sc); }
        __syncthreads();
        const int n_first = max(s_sum[2], 1);
        if (s_sum[2] == 0 && s_sum[3] == 0) { __syncthreads(); continue; }  // nothing reads from this tile at this angle
        const int zmn = max(0, min(s_sum[0] / n_first - ROT_BOX / 2, nz - ROT_BOX)), xmn = max(0, min(s_sum[1] / n_first - ROT_BOX / 2, nx - ROT_BOX)) & ~1;   // even: TMA boxes start on 16 bytes
        const int bh = min(ROT_BOX, nz - zmn), bw = min(ROT_BOX, nx - xmn);
        int s0[ROT_CPT], s1[ROT_CPT];                    // shared-memory offset of the first two readers, -1: outside the box
#pragma unroll
        for (int i = 0; i < ROT_CPT; ++i) {
            s0[i] = s1[i] = -1;
            if (cnt[i] > 0) {
                const int d = dst[beg[i]], z = d / nx, x = d - z * nx;
                if (z >= zmn && z < zmn + bh && x >= xmn && x < xmn + bw) s0[i] = (z - zmn) * ROT_BOX + (x - xmn);
            }
            if (cnt[i] > 1) {
                const int d = dst[beg[i] + 1], z = d / nx, x = d - z * nx;
                if (z >= zmn && z < zmn + bh && x >= xmn && x < xmn + bw) s1[i] = (z - zmn) * ROT_BOX + (x - xmn);
            }
        }
        auto add_readers = [&](int j, const float2* __restrict__ box) __attribute__((always_inline)) {
            const int y = y0 + j;
#pragma unroll
            for (int i = 0; i < ROT_CPT; ++i) {
                if (cnt[i] > 0) {
                    float2 v;
                    if (s0[i] >= 0) v = box[s0[i]];
                    else { const int d = dst[beg[i]], z = d / nx, x = d - z * nx; v = __ldg(base + (long long)z * slice_stride + (long long)y * nx + x); }
                    acc[i][j].x += v.x; acc[i][j].y += v.y;
                }
                if (cnt[i] > 1) {
                    float2 v;
                    if (s1[i] >= 0) v = box[s1[i]];
                    else { const int d = dst[beg[i] + 1], z = d / nx, x = d - z * nx; v = __ldg(base + (long long)z * slice_stride + (long long)y * nx + x); }
                    acc[i][j].x += v.x; acc[i][j].y += v.y;
                }
                if (cnt[i] > 2 && slot[i] < 0)           // work list overflow (not a rotation table): serial tail
                    for (int k = 2; k < cnt[i]; ++k) {
                        const int d = dst[beg[i] + k], z = d / nx, x = d - z * nx;
                        const float2 v = __ldg(base + (long long)z * slice_stride + (long long)y * nx + x);
                        acc[i][j].x += v.x; acc[i][j].y += v.y;
                    }
            }
        };
        {
            const float2* origin = base + (long long)zmn * slice_stride + xmn;        // box element (r, c) of row y: origin + r slice_stride + y nx + c
            float2 r[ROT_BPT];
            rot_box_load(origin + (long long)y0 * nx, slice_stride, bh, bw, tid, r);
            rot_box_store(buf[0], tid, r);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < ROT_YA; ++j) {
                if (j < nyc) {                           // uniform over the CTA
                    const int cur = j & 1;
                    if (j + 1 < nyc) rot_box_load(origin + (long long)(y0 + j + 1) * nx, slice_stride, bh, bw, tid, r);
                    add_readers(j, buf[cur]);
                    if (j + 1 < nyc) rot_box_store(buf[cur ^ 1], tid, r);
                    __syncthreads();
                }
            }
        }
        const int n_long = min(s_sum[3], ROT_LONG);
        if (n_long > 0) {                                // uniform over the CTA
            for (int sl = warp; sl < n_long; sl += ROT_THREADS / 32) {
                const int lb = s_long_beg[sl], lc = s_long_cnt[sl];
                float2 part[ROT_YA];
#pragma unroll
                for (int j = 0; j < ROT_YA; ++j) part[j] = make_float2(0.f, 0.f);
                for (int k = lane; k < lc; k += 32) {
                    const int d = dst[lb + k], z = d / nx, x = d - z * nx;
                    const float2* g = base + (long long)z * slice_stride + (long long)y0 * nx + x;
#pragma unroll
                    for (int j = 0; j < ROT_YA; ++j)
                        if (j < nyc) { const float2 v = __ldg(g + (long long)j * nx); part[j].x += v.x; part[j].y += v.y; }
                }
#pragma unroll
                for (int j = 0; j < ROT_YA; ++j) {
#pragma unroll
                    for (int m = 16; m > 0; m >>= 1) {
                        part[j].x += __shfl_xor_sync(0xffffffffu, part[j].x, m);
                        part[j].y += __shfl_xor_sync(0xffffffffu, part[j].y, m);
                    }
                    if (lane == 0) s_extra[sl][j] = part[j];
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < ROT_CPT; ++i)
                if (slot[i] >= 0) {
#pragma unroll
                    for (int j = 0; j < ROT_YA; ++j) { const float2 v = s_extra[slot[i]][j]; acc[i][j].x += v.x; acc[i][j].y += v.y; }
                }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < ROT_CPT; ++i) {
        if (!ok[i]) continue;
        const int c = tid + ROT_THREADS * i;
        float2* o = gobj + ((long long)(z0 + (c >> 5)) * ny + y0) * nx + x0 + (c & 31);
#pragma unroll
        for (int j = 0; j < ROT_YA; ++j)
            if (j < nyc) {
                float2 v = acc[i][j];
                if (accumulate) { const float2 cur = o[(long long)j * nx]; v.x += cur.x; v.y += cur.y; }
                o[(long long)j * nx] = v;
            }
    }
}

// ---- the TMA form of the same kernel.  The box origins of every (angle, tile) come from a small kernel of their own
// (k_rot_origins), so the box loads do not wait for the lists: one thread streams the (angle, y) boxes of the CTA through a
// ring of ROT_NBUF shared-memory buffers, ROT_NBUF loads in flight, while all threads read the next angle's lists.
constexpr int ROT_NBUF = 4;
// grow-only scratch per (device, stream): calls on one stream are ordered, calls on different streams never share a buffer
// (a stream-ordered allocation per call cost 0.3-2 ms of host time)
static int rot_scratch(int2** out, size_t bytes, cudaStream_t st) {
    struct Buf { void* p; size_t n; };
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, Buf> bufs;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    Buf& b = bufs[std::make_pair(dev, st)];
    if (b.n < bytes) {
        if (b.p) { CUDA_TRY(cudaStreamSynchronize(st)); cudaFree(b.p); b.p = nullptr; b.n = 0; }
        const size_t want = bytes < 65536 ? 65536 : bytes;
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) { b.p = nullptr; return bdof_fail(int(e), "cudaMalloc rotation scratch: %s", cudaGetErrorString(e)); }
        b.n = want;
    }
    *out = static_cast<int2*>(b.p);
    return 0;
}
__global__ void __launch_bounds__(ROT_THREADS) k_rot_origins(const RotLists lists, int nx, int nz, int z_tile0, int2* __restrict__ origins) {
    __shared__ int s_sum[4];
    const int tid = threadIdx.x, a = blockIdx.z;
    const int x0 = blockIdx.x * ROT_T, z0 = (blockIdx.y + z_tile0) * ROT_T;
    const int* __restrict__ off = lists.offsets[a];
    const int* __restrict__ dst = lists.dest[a];
    if (tid == 0) s_sum[0] = s_sum[1] = s_sum[2] = s_sum[3] = 0;
    __syncthreads();
    int sz = 0, sx = 0, sc = 0, sl = 0;
#pragma unroll
    for (int i = 0; i < ROT_CPT; ++i) {
        const int c = tid + ROT_THREADS * i, z = z0 + (c >> 5), x = x0 + (c & 31);
        if (z < nz && x < nx) {
            const int beg = off[z * nx + x], cnt = off[z * nx + x + 1] - beg;
            if (cnt > 0 && cnt <= 2) { const int d = dst[beg], zz = d / nx; sz += zz; sx += d - zz * nx; ++sc; }    // ordinary cells centre the box
            if (cnt > 2) ++sl;
        }
    }
    sz = __reduce_add_sync(0xffffffffu, sz); sx = __reduce_add_sync(0xffffffffu, sx);
    sc = __reduce_add_sync(0xffffffffu, sc); sl = __reduce_add_sync(0xffffffffu, sl);
    if ((tid & 31) == 0) { atomicAdd(&s_sum[0], sz); atomicAdd(&s_sum[1], sx); atomicAdd(&s_sum[2], sc); atomicAdd(&s_sum[3], sl); }
    __syncthreads();
    if (tid == 0) {
        int2 o = make_int2(-1, -1);                      // (z, x) of the box; x = -1: nothing reads from this tile at this angle
        if (s_sum[2] > 0 || s_sum[3] > 0) {
            const int n = max(s_sum[2], 1);
            o.x = max(0, min(s_sum[0] / n - ROT_BOX / 2, nz - ROT_BOX));
            o.y = max(0, min(s_sum[1] / n - ROT_BOX / 2, nx - ROT_BOX)) & ~1;         // even: TMA boxes start on 16 bytes
        }
        origins[((long long)a * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = o;
    }
}
// d / nx for 0 <= d < 2^31 by multiplication: mul = ceil(2^(31 + sh) / nx), sh = ceil(log2 nx) (round-up method; the lists hold
// z * nx + x, and a runtime integer division is ~25 instructions -- the per-angle code of the kernel is fetched, not looped)
struct RotDiv { unsigned mul; int sh; int nx; };
static RotDiv rot_div_make(int nx) {
    RotDiv d; d.nx = nx; d.sh = 0;
    while ((1LL << d.sh) < nx) ++d.sh;
    d.mul = unsigned(((1ULL << (31 + d.sh)) + (unsigned long long)nx - 1) / (unsigned long long)nx);
    return d;
}
__host__ __device__ __forceinline__ void rot_split(const RotDiv& dv, int d, int* z, int* x) {
    const int q = int(((unsigned long long)(unsigned)d * dv.mul) >> (31 + dv.sh));
    *z = q; *x = d - q * dv.nx;
}
// the same arithmetic on the host, for the CPU tests (d = z * nx + x of a list entry -> z, x)
extern "C" int bdof_debug_rot_split(int nx, int d, int* z, int* x) {
    if (nx < 1 || d < 0 || !z || !x) return bdof_fail(BDOF_E_BADARG, "bad argument");
    const RotDiv dv = rot_div_make(nx);
    rot_split(dv, d, z, x);
    return 0;
}
// serial tail of one cell (work list overflow: not a rotation table); rare, kept out of line
__device__ __noinline__ float2 rot_serial_tail(const float2* __restrict__ base, const int* __restrict__ dst, int beg, int n, long long slice_stride, int y, int nx) {
    float2 a = make_float2(0.f, 0.f);
    for (int k = 0; k < n; ++k) {
        const int d = dst[beg + k], z = d / nx, x = d - z * nx;
        const float2 v = __ldg(base + (long long)z * slice_stride + (long long)y * nx + x);
        a.x += v.x; a.y += v.y;
    }
    return a;
}
__global__ void __launch_bounds__(ROT_THREADS, 2) k_rotate_adjoint_tma(const float2* __restrict__ grot, long long slice_stride, long long batch_stride,
                                     const RotLists lists, int n_ang, int accumulate, float2* __restrict__ gobj, int ny, int nx, int nz,
                                     int z_tile0, const int2* __restrict__ origins, const RotDiv dv, const __grid_constant__ CUtensorMap tm_grot) {
    extern __shared__ __align__(128) float2 ring[];      // ROT_NBUF boxes
    __shared__ __align__(8) unsigned long long bar[ROT_NBUF];
    __shared__ float2 s_extra[ROT_LONG][ROT_YA];
    __shared__ int s_long_beg[ROT_LONG], s_long_cnt[ROT_LONG];
    __shared__ int2 s_org[BDOF_ROT_MAX_ANGLES];          // box origin (z, x) of the angles that have readers, in angle order
    __shared__ int s_ang[BDOF_ROT_MAX_ANGLES];
    __shared__ int s_nv, s_nlong;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int y0 = blockIdx.x * ROT_YA, x0 = blockIdx.y * ROT_T, z0 = (blockIdx.z + z_tile0) * ROT_T;      // y fastest, as the gather
    const int nyc = min(ROT_YA, ny - y0);
    if (tid == 0) {
        int nv = 0;
        for (int a = 0; a < n_ang; ++a) {
            const int2 o = origins[((long long)a * gridDim.z + blockIdx.z) * gridDim.y + blockIdx.y];
            if (o.x >= 0) { s_org[nv] = o; s_ang[nv] = a; ++nv; }
        }
        s_nv = nv;
        for (int b = 0; b < ROT_NBUF; ++b) mbar_init(&bar[b], 1);
        fence_mbar_init();
    }
    float2 acc[ROT_CPT][ROT_YA];
    bool ok[ROT_CPT];
    int cell[ROT_CPT];
#pragma unroll
    for (int i = 0; i < ROT_CPT; ++i) {
        const int c = tid + ROT_THREADS * i, z = z0 + (c >> 5), x = x0 + (c & 31);
        ok[i] = z < nz && x < nx;
        cell[i] = ok[i] ? z * nx + x : 0;
#pragma unroll
        for (int j = 0; j < ROT_YA; ++j) acc[i][j] = make_float2(0.f, 0.f);
    }
    __syncthreads();
    const int nv = s_nv, n_loads = nv * nyc;
    // load q = (valid angle q / nyc, row q % nyc) goes to buffer q % ROT_NBUF; its mbarrier completes phase q / ROT_NBUF
    auto issue = [&](int q) __attribute__((always_inline)) {
        const int vi = q / nyc, j = q - vi * nyc, b = q % ROT_NBUF;
        const int2 o = s_org[vi];
        mbar_expect_tx(&bar[b], ROT_BOX_BYTES);
        tma_load_4d(ring + b * (ROT_BOX * ROT_BOX), &tm_grot, 2 * o.y, y0 + j, s_ang[vi], o.x, &bar[b]);
    };
    if (tid == 0)
        for (int q = 0; q < ROT_NBUF && q < n_loads; ++q) issue(q);
    int q = 0;
#pragma unroll 1
    for (int vi = 0; vi < nv; ++vi) {
        const int a = s_ang[vi], zmn = s_org[vi].x, xmn = s_org[vi].y;
        const int* __restrict__ off = lists.offsets[a];
        const int* __restrict__ dst = lists.dest[a];
        const float2* __restrict__ base = grot + (long long)a * batch_stride;
        if (tid == 0) s_nlong = 0;
        __syncthreads();
        // the first one or two readers of a cell come from the box; whatever is left (readers outside the box, the tails of
        // the clipped border cells) goes to the CTA's work list
        int s0[ROT_CPT], s1[ROT_CPT], slot[ROT_CPT], tail_beg[ROT_CPT], tail_n[ROT_CPT];
#pragma unroll
        for (int i = 0; i < ROT_CPT; ++i) {
            const int beg = ok[i] ? off[cell[i]] : 0;
            const int cnt = ok[i] ? off[cell[i] + 1] - beg : 0;
            s0[i] = -1; s1[i] = -1; slot[i] = -1;
            int nf = 0;                                  // leading readers served from the box
            if (cnt > 0) {
                int z, x;
                rot_split(dv, dst[beg], &z, &x);
                if (z >= zmn && z < zmn + ROT_BOX && x >= xmn && x < xmn + ROT_BOX) { s0[i] = (z - zmn) * ROT_BOX + (x - xmn); nf = 1; }
            }
            if (cnt > 1 && nf == 1) {
                int z, x;
                rot_split(dv, dst[beg + 1], &z, &x);
                if (z >= zmn && z < zmn + ROT_BOX && x >= xmn && x < xmn + ROT_BOX) { s1[i] = (z - zmn) * ROT_BOX + (x - xmn); nf = 2; }
            }
            tail_beg[i] = beg + nf; tail_n[i] = cnt - nf;
            if (tail_n[i] > 0) {
                const int sl = atomicAdd(&s_nlong, 1);              // slot order is arbitrary; every cell's sum is its own
                if (sl < ROT_LONG) { slot[i] = sl; s_long_beg[sl] = tail_beg[i]; s_long_cnt[sl] = tail_n[i]; }
            }
        }
#pragma unroll
        for (int j = 0; j < ROT_YA; ++j) {
            if (j < nyc) {                               // uniform over the CTA
                const int b = q % ROT_NBUF;
                const float2* __restrict__ box = ring + b * (ROT_BOX * ROT_BOX);
                mbar_wait(&bar[b], (q / ROT_NBUF) & 1);
#pragma unroll
                for (int i = 0; i < ROT_CPT; ++i) {
                    if (s0[i] >= 0) { const float2 v = box[s0[i]]; acc[i][j].x += v.x; acc[i][j].y += v.y; }
                    if (s1[i] >= 0) { const float2 v = box[s1[i]]; acc[i][j].x += v.x; acc[i][j].y += v.y; }
                }
                __syncthreads();                         // the buffer is free
                if (tid == 0 && q + ROT_NBUF < n_loads) issue(q + ROT_NBUF);
                ++q;
            }
        }
        const int n_long = s_nlong;
        if (n_long > 0) {                                // uniform over the CTA
            for (int sl = warp; sl < min(n_long, ROT_LONG); sl += ROT_THREADS / 32) {
                const int lb = s_long_beg[sl], lc = s_long_cnt[sl];
                float2 part[ROT_YA];
#pragma unroll
                for (int j = 0; j < ROT_YA; ++j) part[j] = make_float2(0.f, 0.f);
                for (int k = lane; k < lc; k += 32) {
                    int z, x;
                    rot_split(dv, dst[lb + k], &z, &x);
                    const float2* g = base + (long long)z * slice_stride + (long long)y0 * nx + x;
#pragma unroll
                    for (int j = 0; j < ROT_YA; ++j)
                        if (j < nyc) { const float2 v = __ldg(g + (long long)j * nx); part[j].x += v.x; part[j].y += v.y; }
                }
#pragma unroll
                for (int j = 0; j < ROT_YA; ++j) {
#pragma unroll
                    for (int m = 16; m > 0; m >>= 1) {
                        part[j].x += __shfl_xor_sync(0xffffffffu, part[j].x, m);
                        part[j].y += __shfl_xor_sync(0xffffffffu, part[j].y, m);
                    }
                    if (lane == 0) s_extra[sl][j] = part[j];
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < ROT_CPT; ++i) {
                if (slot[i] >= 0) {
#pragma unroll
                    for (int j = 0; j < ROT_YA; ++j) { const float2 v = s_extra[slot[i]][j]; acc[i][j].x += v.x; acc[i][j].y += v.y; }
                } else if (tail_n[i] > 0) {              // work list overflow
#pragma unroll
                    for (int j = 0; j < ROT_YA; ++j)
                        if (j < nyc) { const float2 v = rot_serial_tail(base, dst, tail_beg[i], tail_n[i], slice_stride, y0 + j, nx); acc[i][j].x += v.x; acc[i][j].y += v.y; }
                }
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < ROT_CPT; ++i) {
        if (!ok[i]) continue;
        const int c = tid + ROT_THREADS * i;
        float2* o = gobj + ((long long)(z0 + (c >> 5)) * ny + y0) * nx + x0 + (c & 31);
#pragma unroll
        for (int j = 0; j < ROT_YA; ++j)
            if (j < nyc) {
                float2 v = acc[i][j];
                if (accumulate) { const float2 cur = o[(long long)j * nx]; v.x += cur.x; v.y += cur.y; }
                o[(long long)j * nx] = v;
            }
    }
}
extern "C" int bdof_rotate_adjoint_csr_batch_range(const float* d_grad_rot_db, long long slice_stride_px, long long batch_stride_px, int n_angles,
                                                   const int32_t* const* d_offsets, const int32_t* const* d_dest, float* d_grad_obj_db,
                                                   int accumulate, int ny, int nx, int nz, int z_begin, int z_end, void* st) {
    if (!d_grad_rot_db || !d_offsets || !d_dest || !d_grad_obj_db || ny < 1 || nx < 1 || nz < 1 || n_angles < 1) return fail(BDOF_E_BADARG, "bad argument");
    if (z_begin < 0 || z_end > nz || z_begin >= z_end || z_begin % ROT_T != 0 || (z_end % ROT_T != 0 && z_end != nz))
        return fail(BDOF_E_BADARG, "z range [%d, %d): bounds must be multiples of %d (or the last slice)", z_begin, z_end, ROT_T);
    if (nx > 65535 * ROT_T || nz > 65535 * ROT_T) return fail(BDOF_E_UNSUPPORTED, "nx / nz too large");
    const int z_tile0 = z_begin / ROT_T;
    dim3 grid((ny + ROT_YA - 1) / ROT_YA, (nx + ROT_T - 1) / ROT_T, (z_end - z_begin + ROT_T - 1) / ROT_T);
    // TMA box loads need a tensor the descriptor can express: even nx (16-byte row pitch), sides at least one box, the
    // minibatch laid out [z][angle][y][x] (a plan's db) or a single angle
    const bool layout_ok = n_angles == 1 || (batch_stride_px == (long long)ny * nx && slice_stride_px >= (long long)n_angles * ny * nx);
    const bool tma = rot_use_tma() && (nx % 2 == 0) && nx >= ROT_BOX && nz >= ROT_BOX && layout_ok && (reinterpret_cast<uintptr_t>(d_grad_rot_db) % 16 == 0) &&
                     (slice_stride_px % 2 == 0) && (batch_stride_px % 2 == 0);
    for (int a0 = 0; a0 < n_angles; a0 += BDOF_ROT_MAX_ANGLES) {
        RotLists l;
        const int n = n_angles - a0 < BDOF_ROT_MAX_ANGLES ? n_angles - a0 : BDOF_ROT_MAX_ANGLES;
        for (int a = 0; a < n; ++a) {
            if (!d_offsets[a0 + a] || !d_dest[a0 + a]) return fail(BDOF_E_BADARG, "null list");
            l.offsets[a] = d_offsets[a0 + a]; l.dest[a] = d_dest[a0 + a];
        }
        const float2* base = reinterpret_cast<const float2*>(d_grad_rot_db) + (long long)a0 * batch_stride_px;
        const int acc = (accumulate || a0 > 0) ? 1 : 0;
        if (tma) {
            alignas(64) CUtensorMap tm;
            const long long bstride = n == 1 ? (long long)ny * nx : batch_stride_px;             // one angle: the dimension has extent 1
            const long long dims[4] = {2LL * nx, ny, n, nz};
            const long long strides[3] = {(long long)nx * 8, bstride * 8, slice_stride_px * 8};
            const int box[4] = {2 * ROT_BOX, 1, 1, ROT_BOX};
            BDOF_TRY(make_tensor_map_nd(&tm, base, 4, dims, strides, box));
            const int ring_bytes = ROT_NBUF * int(ROT_BOX_BYTES);
            CUDA_TRY(cudaFuncSetAttribute(k_rotate_adjoint_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes));   // per device: set every time
            int2* origins = nullptr;                     // scratch [angle][tile z][tile x], one buffer per (device, stream)
            BDOF_TRY(rot_scratch(&origins, sizeof(int2) * n * grid.y * grid.z, (cudaStream_t)st));
            k_rot_origins<<<dim3(grid.y, grid.z, n), ROT_THREADS, 0, (cudaStream_t)st>>>(l, nx, nz, z_tile0, origins);
            int r = launch_check("k_rot_origins");
            if (!r) {
                k_rotate_adjoint_tma<<<grid, ROT_THREADS, ring_bytes, (cudaStream_t)st>>>(base, slice_stride_px, batch_stride_px, l, n, acc,
                                                                                        reinterpret_cast<float2*>(d_grad_obj_db), ny, nx, nz, z_tile0, origins, rot_div_make(nx), tm);
                r = launch_check("k_rotate_adjoint_tma");
            }
            if (r) return r;
        } else {
            k_rotate_adjoint_csr<<<grid, ROT_THREADS, 0, (cudaStream_t)st>>>(base, slice_stride_px, batch_stride_px, l, n, acc,
                                                                           reinterpret_cast<float2*>(d_grad_obj_db), ny, nx, nz, z_tile0);
        }
        if (int r = launch_check("k_rotate_adjoint_csr")) return r;
    }
    return 0;
}
extern "C" int bdof_rotate_adjoint_csr_batch(const float* d_grad_rot_db, long long slice_stride_px, long long batch_stride_px, int n_angles,
                                             const int32_t* const* d_offsets, const int32_t* const* d_dest, float* d_grad_obj_db,
                                             int accumulate, int ny, int nx, int nz, void* st) {
    return bdof_rotate_adjoint_csr_batch_range(d_grad_rot_db, slice_stride_px, batch_stride_px, n_angles, d_offsets, d_dest, d_grad_obj_db, accumulate,
                                               ny, nx, nz, 0, nz, st);
}
extern "C" int bdof_rotate_adjoint_csr(const float* d_grad_rot_db, long long slice_stride_px, const int32_t* d_offsets,
                                       const int32_t* d_dest, float* d_grad_obj_db, int ny, int nx, int nz, void* st) {
    return bdof_rotate_adjoint_csr_batch(d_grad_rot_db, slice_stride_px, 0, 1, &d_offsets, &d_dest, d_grad_obj_db, 1, ny, nx, nz, st);
}
extern "C" int bdof_rotate_gather(const float* d_obj_db, const int32_t* d_lookup_zx, float* d_out_db, long long out_slice_stride_px,
                                  int ny, int nx, int nz, void* st) {
    if (!d_obj_db || !d_lookup_zx || !d_out_db || ny < 1 || nx < 1 || nz < 1) return fail(BDOF_E_BADARG, "bad argument");
    if (nx > 65535 * ROT_T || nz > 65535 * ROT_T) return fail(BDOF_E_UNSUPPORTED, "nx / nz too large");
    dim3 grid((ny + ROT_YB - 1) / ROT_YB, (nx + ROT_T - 1) / ROT_T, (nz + ROT_T - 1) / ROT_T);
    const bool tma = rot_use_tma() && (nx % 2 == 0) && nx >= ROT_BOX && nz >= ROT_BOX && (reinterpret_cast<uintptr_t>(d_obj_db) % 16 == 0);
    alignas(64) CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    if (tma) {
        const long long dims[3] = {2LL * nx, ny, nz};
        const long long strides[2] = {(long long)nx * 8, (long long)ny * nx * 8};
        const int box[3] = {2 * ROT_BOX, 1, ROT_BOX};
        BDOF_TRY(make_tensor_map_nd(&tm, d_obj_db, 3, dims, strides, box));
        k_rotate_gather<true><<<grid, ROT_THREADS, 0, (cudaStream_t)st>>>(reinterpret_cast<const float2*>(d_obj_db), reinterpret_cast<const int2*>(d_lookup_zx),
                                                                        reinterpret_cast<float2*>(d_out_db), out_slice_stride_px, ny, nx, nz, tm);
    } else {
        k_rotate_gather<false><<<grid, ROT_THREADS, 0, (cudaStream_t)st>>>(reinterpret_cast<const float2*>(d_obj_db), reinterpret_cast<const int2*>(d_lookup_zx),
                                                                         reinterpret_cast<float2*>(d_out_db), out_slice_stride_px, ny, nx, nz, tm);
    }
    return launch_check("k_rotate_gather");
}
extern "C" int bdof_rotate_scatter_add(const float* d_grad_rot_db, long long slice_stride_px, const int32_t* d_lookup_zx,
                                       float* d_grad_obj_db, int ny, int nx, int nz, void* st) {
    if (!d_grad_rot_db || !d_lookup_zx || !d_grad_obj_db || ny < 1 || nx < 1 || nz < 1) return fail(BDOF_E_BADARG, "bad argument");
    if (ny > 65535 || nz > 65535) return fail(BDOF_E_UNSUPPORTED, "ny / nz > 65535");
    dim3 grid((nx + 127) / 128, ny, nz);
    k_rotate_scatter_add<<<grid, 128, 0, (cudaStream_t)st>>>(reinterpret_cast<const float2*>(d_grad_rot_db), slice_stride_px,
                                                           reinterpret_cast<const int2*>(d_lookup_zx),
                                                           reinterpret_cast<float2*>(d_grad_obj_db), ny, nx, nz);
    return launch_check("k_rotate_scatter_add");
}

// Bilinear rotation of the TF drivers: tf.contrib.image.rotate(stack([delta, beta], -1), theta, 'BILINEAR') on [Y, X, Z, 2]
// (tensorflow_recon/fullfield.py:96, ptychography.py:39), i.e. images of height X and width Z, one per y.  TF semantics
// (contrib/image: angles_to_projective_transforms + ProjectiveGenerator): output pixel (x_o, z_o) samples the input at
//   z_i = cos z_o - sin x_o + ((W-1) - (cos (W-1) - sin (H-1))) / 2,   x_i = sin z_o + cos x_o + ((H-1) - (sin (W-1) + cos (H-1))) / 2
// (W = nz, H = nx) with bilinear weights over floor / floor + 1 and ZERO outside the image.
struct BilinearTaps { int x0, z0; float wx0, wx1, wz0, wz1; };
__device__ __forceinline__ BilinearTaps bilinear_taps(int x_o, int z_o, float c, float s, float off_z, float off_x) {
    const float zi = c * float(z_o) - s * float(x_o) + off_z;
    const float xi = s * float(z_o) + c * float(x_o) + off_x;
    const float zf = floorf(zi), xf = floorf(xi);
    BilinearTaps t;
    t.x0 = int(xf); t.z0 = int(zf);
    t.wx1 = xi - xf; t.wx0 = (xf + 1.f) - xi;
    t.wz1 = zi - zf; t.wz0 = (zf + 1.f) - zi;
    return t;
}
__global__ void k_rotate_bilinear(const float2* __restrict__ obj, float2* __restrict__ out, long long out_slice_stride, int ny, int nx,
                                  int nz, float c, float s, float off_z, float off_x) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, z = blockIdx.z;
    if (x >= nx) return;
    const BilinearTaps t = bilinear_taps(x, z, c, s, off_z, off_x);
    auto rd = [&](int xi, int zi) {
        if (xi < 0 || xi >= nx || zi < 0 || zi >= nz) return make_float2(0.f, 0.f);
        return obj[((long long)zi * ny + y) * nx + xi];
    };
    // TF order: interpolate along the width (z) first, then along the height (x)
    const float2 a0 = rd(t.x0, t.z0), a1 = rd(t.x0, t.z0 + 1), b0 = rd(t.x0 + 1, t.z0), b1 = rd(t.x0 + 1, t.z0 + 1);
    const float vx0d = t.wz0 * a0.x + t.wz1 * a1.x, vx0b = t.wz0 * a0.y + t.wz1 * a1.y;
    const float vx1d = t.wz0 * b0.x + t.wz1 * b1.x, vx1b = t.wz0 * b0.y + t.wz1 * b1.y;
    out[(long long)z * out_slice_stride + (long long)y * nx + x] = make_float2(t.wx0 * vx0d + t.wx1 * vx1d, t.wx0 * vx0b + t.wx1 * vx1b);
}
// transpose: every rotated pixel adds its weighted gradient to its (up to) four source pixels
__global__ void k_rotate_bilinear_adj(const float2* __restrict__ grot, long long slice_stride, float2* __restrict__ gobj, int ny, int nx,
                                      int nz, float c, float s, float off_z, float off_x) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, z = blockIdx.z;
    if (x >= nx) return;
    const BilinearTaps t = bilinear_taps(x, z, c, s, off_z, off_x);
    const float2 g = grot[(long long)z * slice_stride + (long long)y * nx + x];
    auto add = [&](int xi, int zi, float w) {
        if (xi < 0 || xi >= nx || zi < 0 || zi >= nz || w == 0.f) return;
        float* dst = reinterpret_cast<float*>(gobj + ((long long)zi * ny + y) * nx + xi);
        atomicAdd(dst, w * g.x);
        atomicAdd(dst + 1, w * g.y);
    };
    add(t.x0, t.z0, t.wx0 * t.wz0);
    add(t.x0, t.z0 + 1, t.wx0 * t.wz1);
    add(t.x0 + 1, t.z0, t.wx1 * t.wz0);
    add(t.x0 + 1, t.z0 + 1, t.wx1 * t.wz1);
}
static void bilinear_consts(double theta, int nx, int nz, float* c, float* s, float* off_z, float* off_x) {
    const double cs = std::cos(theta), sn = std::sin(theta), W = double(nz), H = double(nx);
    *c = float(cs); *s = float(sn);
    *off_z = float(((W - 1) - (cs * (W - 1) - sn * (H - 1))) / 2.0);
    *off_x = float(((H - 1) - (sn * (W - 1) + cs * (H - 1))) / 2.0);
}
extern "C" int bdof_rotate_bilinear(const float* d_obj_db, float* d_out_db, long long out_slice_stride_px, double theta, int ny, int nx,
                                    int nz, void* st) {
    if (!d_obj_db || !d_out_db || ny < 1 || nx < 1 || nz < 1) return fail(BDOF_E_BADARG, "bad argument");
    if (ny > 65535 || nz > 65535) return fail(BDOF_E_UNSUPPORTED, "ny / nz > 65535");
    float c, s, oz, ox;
    bilinear_consts(theta, nx, nz, &c, &s, &oz, &ox);
    dim3 grid((nx + 127) / 128, ny, nz);
    k_rotate_bilinear<<<grid, 128, 0, (cudaStream_t)st>>>(reinterpret_cast<const float2*>(d_obj_db), reinterpret_cast<float2*>(d_out_db),
                                                        out_slice_stride_px, ny, nx, nz, c, s, oz, ox);
    return launch_check("k_rotate_bilinear");
}
extern "C" int bdof_rotate_bilinear_adjoint(const float* d_grad_rot_db, long long slice_stride_px, float* d_grad_obj_db, double theta, int ny,
                                            int nx, int nz, void* st) {
    if (!d_grad_rot_db || !d_grad_obj_db || ny < 1 || nx < 1 || nz < 1) return fail(BDOF_E_BADARG, "bad argument");
    if (ny > 65535 || nz > 65535) return fail(BDOF_E_UNSUPPORTED, "ny / nz > 65535");
    float c, s, oz, ox;
    bilinear_consts(theta, nx, nz, &c, &s, &oz, &ox);
    dim3 grid((nx + 127) / 128, ny, nz);
    k_rotate_bilinear_adj<<<grid, 128, 0, (cudaStream_t)st>>>(reinterpret_cast<const float2*>(d_grad_rot_db), slice_stride_px,
                                                            reinterpret_cast<float2*>(d_grad_obj_db), ny, nx, nz, c, s, oz, ox);
    return launch_check("k_rotate_bilinear_adj");
}

// Adam (apply_gradient_adam, cnn_propagator/util.py:280-291): one fused pass over x, g, m, v
__global__ void k_adam(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                       float b1, float b2, float omb1, float omb2, float inv_c1, float inv_c2, float step, float eps) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = omb1 * gi + b1 * m[i];          // 1 - b formed in double on the host (1.f - 0.999f is 1.3e-5 off)
        const float vi = omb2 * (gi * gi) + b2 * v[i];
        m[i] = mi;
        v[i] = vi;
        x[i] = x[i] - step * (mi * inv_c1) / (sqrtf(vi * inv_c2) + eps);
    }
}
extern "C" int bdof_adam_step(float* d_x, const float* d_g, float* d_m, float* d_v, long long n, int i_batch, double step_size,
                              double b1, double b2, double eps, void* st) {
    if (!d_x || !d_g || !d_m || !d_v || n < 1 || i_batch < 0) return fail(BDOF_E_BADARG, "bad argument");
    const double c1 = 1.0 - std::pow(b1, i_batch + 1), c2 = 1.0 - std::pow(b2, i_batch + 1);
    k_adam<<<LOSS_BLOCKS, 256, 0, (cudaStream_t)st>>>(d_x, d_g, d_m, d_v, n, float(b1), float(b2), float(1.0 - b1), float(1.0 - b2), float(1.0 / c1), float(1.0 / c2),
                                                    float(step_size), float(eps));
    return launch_check("k_adam");
}

// L1 and 3-D total-variation regularisers of the TF driver (tensorflow_recon/fullfield.py:389-396, util.py:913-923 /
// cnn_propagator/util.py:61-70: periodic first differences, L1) on the native object [nz][ny][nx][2]: ONE pass adds
//   alpha_d sign(delta) + gamma dTV/ddelta   to the delta gradient,   alpha_b sign(beta)   to the beta gradient
// and accumulates  alpha_d |delta|_1 + alpha_b |beta|_1 + gamma TV(delta)  (double partial sums, fixed order: deterministic).
// dTV/dx_i = sum over axes of s_{i+1} - s_i with s_i = sign(x_{i-1} - x_i).
__device__ __forceinline__ float sgnf(float v) { return float(v > 0.f) - float(v < 0.f); }
__global__ void __launch_bounds__(LOSS_THREADS) k_regularizers(const float2* __restrict__ obj, float2* __restrict__ grad, int nz, int ny, int nx,
                                                                float alpha_d, float alpha_b, float gamma, double* __restrict__ partial) {
    const long long n = (long long)nz * ny * nx;
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = int(i % nx);
        const long long r = i / nx;
        const int y = int(r % ny), z = int(r / ny);
        const float2 v = obj[i];
        float gd = alpha_d * sgnf(v.x);
        const float gb = alpha_b * sgnf(v.y);
        acc += double(alpha_d) * fabsf(v.x) + double(alpha_b) * fabsf(v.y);
        if (gamma != 0.f) {
            const long long sx = 1, sy = nx, sz = (long long)ny * nx;
            const float xm = obj[i + (x > 0 ? -sx : sx * (nx - 1))].x, xp = obj[i + (x < nx - 1 ? sx : -sx * (nx - 1))].x;
            const float ym = obj[i + (y > 0 ? -sy : sy * (ny - 1))].x, yp = obj[i + (y < ny - 1 ? sy : -sy * (ny - 1))].x;
            const float zm = obj[i + (z > 0 ? -sz : sz * (nz - 1))].x, zp = obj[i + (z < nz - 1 ? sz : -sz * (nz - 1))].x;
            // s_i = sign(x_{i-1} - x_i), s_{i+1} = sign(x_i - x_{i+1})
            gd += gamma * ((sgnf(v.x - xp) - sgnf(xm - v.x)) + (sgnf(v.x - yp) - sgnf(ym - v.x)) + (sgnf(v.x - zp) - sgnf(zm - v.x)));
            acc += double(gamma) * (fabsf(xm - v.x) + fabsf(ym - v.x) + fabsf(zm - v.x));
        }
        if (grad != nullptr) {
            float2 g = grad[i];
            g.x += gd; g.y += gb;
            grad[i] = g;
        }
    }
    __shared__ double red[LOSS_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) t += red[w];
        partial[blockIdx.x] = t;
    }
}
__global__ void k_add_partials(const double* __restrict__ partial, int n_partial, double* __restrict__ inout) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += 32) acc += partial[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) *inout += acc;
}
extern "C" int bdof_regularizers(const float* d_obj_db, float* d_grad_db, int nz, int ny, int nx, double alpha_d, double alpha_b, double gamma,
                                 double* d_loss_inout, double* d_partial_work, void* st) {
    if (!d_obj_db || !d_loss_inout || !d_partial_work || nz < 1 || ny < 1 || nx < 1) return fail(BDOF_E_BADARG, "bad argument");
    k_regularizers<<<LOSS_BLOCKS, LOSS_THREADS, 0, (cudaStream_t)st>>>(reinterpret_cast<const float2*>(d_obj_db), reinterpret_cast<float2*>(d_grad_db), nz, ny,
                                                                      nx, float(alpha_d), float(alpha_b), float(gamma), d_partial_work);
    BDOF_TRY(launch_check("k_regularizers"));
    k_add_partials<<<1, 32, 0, (cudaStream_t)st>>>(d_partial_work, LOSS_BLOCKS, d_loss_inout);
    return launch_check("k_add_partials");
}

// finite support + non-negativity + shrink-wrap (cnn_propagator/fullfield.py:359-368): x <- clip(x * mask, 0, inf) on both
// channels of the interleaved object; shrink_threshold >= 0 also updates mask <- mask * (delta > threshold)
__global__ void k_finite_support(float2* __restrict__ x, float* __restrict__ mask, long long npx, float shrink_threshold) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        float2 v = x[i];
        if (mask != nullptr) {
            const float mk = mask[i];
            v.x *= mk; v.y *= mk;
        }
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f);
        x[i] = v;
        if (mask != nullptr && shrink_threshold >= 0.f) mask[i] = mask[i] * (v.x > shrink_threshold ? 1.f : 0.f);
    }
}
extern "C" int bdof_finite_support(float* d_x_db, float* d_mask, long long n_px, double shrink_threshold, void* st) {
    if (!d_x_db || n_px < 1) return fail(BDOF_E_BADARG, "bad argument");
    k_finite_support<<<LOSS_BLOCKS, 256, 0, (cudaStream_t)st>>>(reinterpret_cast<float2*>(d_x_db), d_mask, n_px, float(shrink_threshold));
    return launch_check("k_finite_support");
}

// ------------------------------------------------------------------------------------------
// in-situ kernel timing and the free-space step on its own
// ------------------------------------------------------------------------------------------
extern "C" int bdof_profile_begin(bdof_plan* p) {
    if (!p) return fail(BDOF_E_BADARG, "null");
    for (cudaEvent_t e : p->prof_events) cudaEventDestroy(e);
    p->prof_events.clear(); p->prof_variant.clear();
    p->profile = true;
    return 0;
}
extern "C" int bdof_profile_end(bdof_plan* p, int n_variants, int* counts, double* ms_total) {
    if (!p || !counts || !ms_total || n_variants < 1) return fail(BDOF_E_BADARG, "bad argument");
    p->profile = false;
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    for (int v = 0; v < n_variants; ++v) { counts[v] = 0; ms_total[v] = 0.0; }
    for (size_t i = 0; i < p->prof_variant.size(); ++i) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, p->prof_events[2 * i], p->prof_events[2 * i + 1]));
        const int v = p->prof_variant[i];
        if (v >= 0 && v < n_variants) { counts[v] += 1; ms_total[v] += double(ms); }
    }
    for (cudaEvent_t e : p->prof_events) cudaEventDestroy(e);
    p->prof_events.clear(); p->prof_variant.clear();
    return 0;
}

// adjoint of the plan's free-space step on its own (the head of bdof_adjoint): out = free_prop^H(in)
extern "C" int bdof_free_prop_adjoint(bdof_plan* p, const float* d_in_f, float* d_out_f) {
    if (!p || !d_in_f || !d_out_f) return fail(BDOF_E_BADARG, "null");
    const float2* in = reinterpret_cast<const float2*>(d_in_f);
    float2* out = reinterpret_cast<float2*>(d_out_f);
    if (p->free_mode == BDOF_FREE_INF) {
        LineParams cc = col_params(p, in, p->tmp, nullptr);
        cc.in_shift = p->ny / 2;
        BDOF_TRY(col_pass(p, V_COL_INV, cc));
        LineParams r = row_params(p, p->tmp, out, nullptr);
        r.in_shift = p->nx / 2;
        return row_pass(p, V_ROW_INV, r);
    }
    if (p->free_mode == BDOF_FREE_TF) {
        const std::complex<double> c = std::conj(p->phasef);
        k_scale_complex<<<blocks_for(p->F, 256), 256, 0, p->stream>>>(in, out, p->F, float(c.real()), float(c.imag()));
        BDOF_TRY(launch_check("k_scale_complex"));
        BDOF_TRY(col_pass(p, V_COL_CONV, col_params(p, out, p->tmp, p->ay.hf_adj)));
        return row_pass(p, V_ROW_CONV, row_params(p, p->tmp, out, p->ax.hf_adj));
    }
    if (in != out) CUDA_TRY(cudaMemcpyAsync(out, in, (size_t)p->F * sizeof(float2), cudaMemcpyDeviceToDevice, p->stream));
    return 0;
}

extern "C" int bdof_plan_last_times(bdof_plan* p, double* ms_out, int* launches_out) {
    if (!p || !ms_out || !launches_out) return fail(BDOF_E_BADARG, "null");
    if (!p->forward_done) return fail(BDOF_E_STATE, "no forward has run on this plan");
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, p->t_ev[0], p->t_ev[1]));
    ms_out[0] = ms; launches_out[0] = p->t_launches[0];
    ms_out[1] = 0.0; launches_out[1] = p->t_launches[1];
    if (p->t_launches[1] > 0) {
        CUDA_TRY(cudaEventElapsedTime(&ms, p->t_ev[2], p->t_ev[3]));
        ms_out[1] = ms;
    }
    return 0;
}

extern "C" int bdof_free_prop(bdof_plan* p, const float* d_in_f, float* d_out_f) {
    if (!p || !d_in_f || !d_out_f) return fail(BDOF_E_BADARG, "null");
    const float2* in = reinterpret_cast<const float2*>(d_in_f);
    float2* out = reinterpret_cast<float2*>(d_out_f);
    if (p->free_mode == BDOF_FREE_INF) {
        LineParams r = row_params(p, in, p->tmp, nullptr);
        r.out_shift = p->nx / 2;
        BDOF_TRY(row_pass(p, V_ROW_FWD, r));
        LineParams c = col_params(p, p->tmp, out, nullptr);
        c.out_shift = p->ny / 2;
        return col_pass(p, V_COL_FWD, c);
    }
    if (p->free_mode == BDOF_FREE_TF) {
        BDOF_TRY(row_pass(p, V_ROW_CONV, row_params(p, in, p->tmp, p->ax.hf)));
        BDOF_TRY(col_pass(p, V_COL_CONV, col_params(p, p->tmp, out, p->ay.hf)));
        const std::complex<double> c = p->phasef;
        k_scale_complex<<<blocks_for(p->F, 256), 256, 0, p->stream>>>(out, out, p->F, float(c.real()), float(c.imag()));
        return launch_check("k_scale_complex");
    }
    if (in != out) CUDA_TRY(cudaMemcpyAsync(out, in, (size_t)p->F * sizeof(float2), cudaMemcpyDeviceToDevice, p->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------
// single-slice stepping (tiling with halo exchange) and gradient buckets (overlapped all-reduce)
// ------------------------------------------------------------------------------------------
extern "C" int bdof_slice_step(bdof_plan* p, const float* d_in, const float* d_db_slice, float* d_out, int propagate) {
    return bdof_slice_step_seq(p, d_in, d_db_slice, d_out, propagate, -1);
}
extern "C" int bdof_slice_step_seq(bdof_plan* p, const float* d_in, const float* d_db_slice, float* d_out, int propagate, int slice_index) {
    if (!p || !d_in || !d_db_slice || !d_out) return fail(BDOF_E_BADARG, "null");
    if (slice_index >= 0 && !(p->flags & BDOF_STEPWISE)) return fail(BDOF_E_STATE, "bdof_slice_step_seq needs a plan created with BDOF_STEPWISE");
    if (!p->have_kernel) return fail(BDOF_E_STATE, "bdof_set_kernel has not been called");
    const float2* in = reinterpret_cast<const float2*>(d_in);
    const float2* db = reinterpret_cast<const float2*>(d_db_slice);
    float2* out = reinterpret_cast<float2*>(d_out);
    if (propagate) return propagate_slice(p, in, db, out, nullptr, slice_index);     // -1: a step on its own, the plain table
    k_modulate<<<blocks_for(p->F, 256), 256, 0, p->stream>>>(in, db, out, p->F, float(p->k_dz));
    return launch_check("k_modulate");
}

extern "C" int bdof_plan_set_windows(bdof_plan* p, int oy, int ox, const int* d_origin_yx) {
    if (!p || (d_origin_yx && (oy < 1 || ox < 1))) return fail(BDOF_E_BADARG, "bad argument");
    p->win_origin = d_origin_yx; p->win_oy = oy; p->win_ox = ox;
    p->stash_valid = false;
    return 0;
}
extern "C" int bdof_plan_is_resident(const bdof_plan* p) {
    if (!(p && p->have_kernel && use_resident(p))) return 0;
    return p->nx == 64 ? 1 : 2;                          // 1: one CTA per field (window mode available), 2: one cluster per field
}

// One propagating slice whose input lines are read straight out of a larger pitched buffer (tiling: the windows of a block; the
// cut is fused into the row pass): window b's first row starts d_in_offsets[b] elements into d_buf, rows are `pitch` elements
// apart.  Row lengths >= 1024 are fetched by bulk copies and need 16-byte aligned rows: even offsets and an even pitch.
extern "C" int bdof_slice_step_windows(bdof_plan* p, const float* d_buf, long long pitch, const long long* d_in_offsets, const float* d_db_tiles,
                                       float* d_out_tiles, int slice_index) {
    if (!p || !d_buf || !d_in_offsets || !d_db_tiles || !d_out_tiles || pitch < p->nx) return fail(BDOF_E_BADARG, "bad argument");
    if (!p->have_kernel) return fail(BDOF_E_STATE, "bdof_set_kernel has not been called");
    if (p->full_kernel || p->generic) return fail(BDOF_E_UNSUPPORTED, "windowed stepping needs a separable kernel and power-of-two window sides");
    if (slice_index >= 0 && !(p->flags & BDOF_STEPWISE)) return fail(BDOF_E_STATE, "a slice index needs a plan created with BDOF_STEPWISE");
    if (p->nx >= 1024 && (pitch % 2 != 0)) return fail(BDOF_E_BADARG, "rows of 1024 pixels and more need an even pitch (16-byte aligned bulk copies)");
    LineParams r = row_params(p, reinterpret_cast<const float2*>(d_buf), p->tmp, h_entry(p->ax, slice_index, false));
    r.db = reinterpret_cast<const float2*>(d_db_tiles);
    r.in_offsets = d_in_offsets;
    r.in_line_stride = int(pitch);
    BDOF_TRY(row_pass(p, V_ROW_CONV_T, r));
    return col_pass(p, V_COL_CONV, col_params(p, p->tmp, reinterpret_cast<float2*>(d_out_tiles), h_entry(p->ay, slice_index, false)));
}

extern "C" int bdof_plan_set_grad_accumulate(bdof_plan* p, int on) {
    if (!p) return fail(BDOF_E_BADARG, "null");
    p->grad_accumulate = on != 0;
    return 0;
}

extern "C" int bdof_plan_set_t_stash(bdof_plan* p, float* d_stash) {
    if (!p) return fail(BDOF_E_BADARG, "null");
    p->t_stash = reinterpret_cast<float2*>(d_stash);
    p->stash_valid = false;
    return 0;
}

extern "C" int bdof_plan_set_bucket_events(bdof_plan* p, int n_buckets, void** cuda_events) {
    if (!p || n_buckets < 0 || (n_buckets > 0 && !cuda_events)) return fail(BDOF_E_BADARG, "bad argument");
    p->bucket_events.clear();
    for (int j = 0; j < n_buckets; ++j) p->bucket_events.push_back(reinterpret_cast<cudaEvent_t>(cuda_events[j]));
    return 0;
}
