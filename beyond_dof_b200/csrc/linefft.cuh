// Line kernels: one pass of the separable Fresnel step over a bundle of lines of the field.
//
// A "line" is a row (contiguous) or a column (stride nx) of one batch element.  A persistent CTA
// loops over tiles of LPC lines; a tile lives entirely on chip: every thread keeps E = N/T elements
// of one line in registers (element t + T*q in register q), the Stockham stages run as in-register
// radix-R butterflies (regfft.cuh, packed FFMA2/FADD2 math) and the inter-stage exchanges go through
// padded shared memory.  The stage twiddles and the frequency-domain multiplier h are staged once per
// CTA in shared memory.  One kernel does
//     load [x transmission exp(k(i delta - beta))] -> FFT -> x h -> IFFT -> [adjoint epilogue] -> store
// so a pass is exactly one HBM read and one HBM write of the field (plus delta/beta once).
//
// Only the forward transform is instantiated: IFFT(z) = conj(FFT(conj(z))), and the two transforms
// of a convolution run through the same code in a two-trip loop (halves the instruction footprint).
#pragma once
#include "regfft.cuh"
#include "common.h"

namespace bdof {

enum LineMode { MODE_CONV = 0, MODE_FWD = 1, MODE_INV = 2, MODE_CONV2D = 3 };
enum LinePre { PRE_NONE = 0, PRE_TRANSMIT = 1 };
enum LinePost { POST_NONE = 0, POST_ADJ = 1 };

// FFT plan of one line length: N = R1*R2*R3 (R3 may be 1), T threads per line.
template <int N_, int T_, int R1_, int R2_, int R3_>
struct LineCfg {
    static constexpr int N = N_, T = T_, R1 = R1_, R2 = R2_, R3 = R3_;
    static constexpr int E = N / T;
    static_assert(R1 * R2 * R3 == N, "radix product");
    static_assert(E * T == N, "threads");
    static_assert(E % R1 == 0 && E % R2 == 0 && E % R3 == 0, "radix must divide elements/thread");
    static_assert(R1 >= 2 && R2 >= 2, "at least two stages");
    static constexpr int PADDED = N + N / R1;     // one pad element every R1
    // twiddle table layout: stage 2 at [0, (R2-1)*R1), stage 3 after it
    static constexpr int TW2 = (R2 - 1) * R1;
    static constexpr int TW3 = (R3 > 1) ? (R3 - 1) * R1 * R2 : 0;
    static constexpr int TW_TOTAL = TW2 + TW3;
};

template <class Cfg, int LPC, bool COL, int MODE>
struct LineSmem {
    // line stride in float2: banks of the LPC interleaved lines must not collide in col mode
    static constexpr int ADJ = COL ? ((((16 / LPC) - Cfg::PADDED) % 16) + 16) % 16 : 0;
    static constexpr int STRIDE = Cfg::PADDED + ADJ;
    static constexpr int TW_ELEMS = (Cfg::TW_TOTAL + 1) & ~1;
    // h is staged in shared memory unless that would overflow the 227 KB CTA limit (8192-long columns)
    static constexpr bool H_IN_SMEM = (MODE == MODE_CONV) &&
        size_t(TW_ELEMS + Cfg::N + STRIDE * LPC) * sizeof(float2) <= 227 * 1024;
    static constexpr int H_ELEMS = H_IN_SMEM ? Cfg::N : 0;
    static constexpr size_t BYTES = size_t(TW_ELEMS + H_ELEMS + STRIDE * LPC) * sizeof(float2);
};

template <class Cfg, int LPC, bool COL>
__device__ __forceinline__ void line_sync() {
    if constexpr (!COL && Cfg::T <= 32) __syncwarp();
    else __syncthreads();
}

// butterflies of radix R on the register file: E/R independent butterflies per thread
template <class Cfg, int R>
__device__ __forceinline__ void reg_butterflies(float2 (&v)[Cfg::E]) {
    constexpr int E = Cfg::E, M = E / R;
    static_for<M>([&](auto MM) {
        constexpr int m = decltype(MM)::value;
        float2 a[R];
        static_for<R>([&](auto RR) { constexpr int r = decltype(RR)::value; a[r] = v[m + r * M]; });
        RegFFT<R, false>::run(a);
        static_for<R>([&](auto RR) { constexpr int r = decltype(RR)::value; v[m + r * M] = a[r]; });
    });
}

// Stockham exchange after a radix-R stage with sub-transform size NS (NS = product of earlier radices)
template <class Cfg, int R, int NS>
__device__ __forceinline__ void exchange(float2 (&v)[Cfg::E], int t, float2* sm) {
    constexpr int E = Cfg::E, T = Cfg::T, M = E / R, R1 = Cfg::R1;
    static_for<M>([&](auto MM) {
        constexpr int m = decltype(MM)::value;
        const int j = t + T * m;
        int base;
        if constexpr (NS == 1) {
            base = j * (R1 + 1);                        // pad(j*R + r) with R == R1
        } else {
            const int o = (j / NS) * (NS * R) + (j % NS);
            base = o + o / R1;                          // r*NS adds r*NS + r*NS/R1 exactly (R1 | NS)
        }
        static_for<R>([&](auto RR) {
            constexpr int r = decltype(RR)::value;
            sm[base + r * (NS + NS / R1)] = v[m + r * M];
        });
    });
}

// read element t + T*q back and multiply by the next stage's twiddle W_{NS*R}^{r*k}, k = j mod NS
// (table row r-1 at tw[(r-1)*NS + k], r = q / (E/R), j = t + T*(q mod E/R))
template <class Cfg, int R, int NS>
__device__ __forceinline__ void exchange_read_twiddle(float2 (&v)[Cfg::E], int t, const float2* sm, const float2* tw) {
    constexpr int E = Cfg::E, T = Cfg::T, R1 = Cfg::R1, M = E / R;
    static_for<E>([&](auto Q) {
        constexpr int q = decltype(Q)::value;
        constexpr int m = q % M, r = q / M;
        int idx;
        if constexpr (T % R1 == 0) idx = (t + t / R1) + q * (T + T / R1);
        else { const int i = t + T * q; idx = i + i / R1; }
        const float2 x = sm[idx];
        if constexpr (r == 0) v[q] = x;
        else {
            const int k = (t + T * m) % NS;
            v[q] = cmul(x, tw[(r - 1) * NS + k]);
        }
    });
}

// forward length-N transform of the line held in v (natural order in and out)
template <class Cfg, int LPC, bool COL>
__device__ __forceinline__ void line_fft(float2 (&v)[Cfg::E], int t, float2* sm, const float2* tw) {
    constexpr int R1 = Cfg::R1, R2 = Cfg::R2, R3 = Cfg::R3;
    reg_butterflies<Cfg, R1>(v);
    line_sync<Cfg, LPC, COL>();                 // previous readers of the exchange buffer are done
    exchange<Cfg, R1, 1>(v, t, sm);
    line_sync<Cfg, LPC, COL>();
    exchange_read_twiddle<Cfg, R2, R1>(v, t, sm, tw);
    reg_butterflies<Cfg, R2>(v);
    if constexpr (R3 > 1) {
        line_sync<Cfg, LPC, COL>();
        exchange<Cfg, R2, R1>(v, t, sm);
        line_sync<Cfg, LPC, COL>();
        exchange_read_twiddle<Cfg, R3, R1 * R2>(v, t, sm, tw + Cfg::TW2);
        reg_butterflies<Cfg, R3>(v);
    }
}

template <class Cfg, int LPC, bool COL, int MODE, int PRE, int POST>
__global__ void __launch_bounds__(Cfg::T* LPC) line_kernel(const LineParams p, const int n_tiles) {
    constexpr int N = Cfg::N, T = Cfg::T, E = Cfg::E;
    using SM = LineSmem<Cfg, LPC, COL, MODE>;
    constexpr int CH = (E < 16) ? E : 16;          // prologue / epilogue load batch
    extern __shared__ float2 smem[];
    float2* s_tw = smem;
    float2* s_h = smem + SM::TW_ELEMS;
    float2* s_x = s_h + SM::H_ELEMS;

    const int tid = threadIdx.x;
    for (int i = tid; i < Cfg::TW_TOTAL; i += T * LPC) s_tw[i] = p.tw[i];
    if constexpr (SM::H_IN_SMEM)
        for (int i = tid; i < N; i += T * LPC) s_h[i] = p.h[i];
    __syncthreads();

    int l, t;
    if constexpr (COL) { l = tid % LPC; t = tid / LPC; }
    else               { l = tid / T;   t = tid % T; }
    float2* sm = s_x + l * SM::STRIDE;
    const long long estride = COL ? (long long)p.elem_stride : 1LL;
    const long long step = (long long)T * estride;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long line = (long long)tile * LPC + l;
        const int b = int(line / p.lines_per_batch);
        const int li = int(line - (long long)b * p.lines_per_batch);
        const long long base = (long long)b * p.batch_stride + (long long)li * p.line_stride + (long long)t * estride;

        float2 v[E];
        // ---- load (element t + T*q -> register q)
        {
            const float2* __restrict__ src = p.in + base;
            if constexpr (MODE == MODE_INV) {
                // conj + circular input shift (ifftshift), far-field adjoint only
                const float2* __restrict__ src0 = p.in + (base - (long long)t * estride);
                static_for<E>([&](auto Q) {
                    constexpr int q = decltype(Q)::value;
                    int e = t + T * q + p.in_shift;
                    if (e >= N) e -= N;
                    v[q] = conjf2(src0[(long long)e * estride]);
                });
            } else if constexpr (!COL) {
                static_for<E>([&](auto Q) { constexpr int q = decltype(Q)::value; v[q] = src[T * q]; });
            } else {
                const float2* ptr = src;
                static_for<E>([&](auto Q) { v[decltype(Q)::value] = *ptr; ptr += step; });
            }
        }
        if constexpr (PRE == PRE_TRANSMIT) {
            static_assert(PRE == PRE_NONE || !COL, "transmission is fused into the row pass");
            const float2* __restrict__ dbp = p.db + (long long)b * p.db_batch_stride + (long long)li * p.line_stride + t;
            static_for<E / CH>([&](auto C) {
                constexpr int c = decltype(C)::value;
                float2 d[CH];
                static_for<CH>([&](auto I) { constexpr int i = decltype(I)::value; d[i] = dbp[T * (c * CH + i)]; });
                static_for<CH>([&](auto I) {
                    constexpr int i = decltype(I)::value;
                    v[c * CH + i] = cmul(v[c * CH + i], transmission(d[i], p.k_dz));
                });
            });
        }
        // ---- transform(s)
        if constexpr (MODE == MODE_FWD || MODE == MODE_INV) {
            line_fft<Cfg, LPC, COL>(v, t, sm, s_tw);
        } else {
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                line_fft<Cfg, LPC, COL>(v, t, sm, s_tw);
                if (pass == 0) {
                    // v <- conj(v * h): the second trip then computes conj(IFFT(v h))
                    if constexpr (MODE == MODE_CONV) {
                        const float2* hs = SM::H_IN_SMEM ? s_h : p.h;
                        static_for<E>([&](auto Q) {
                            constexpr int q = decltype(Q)::value;
                            v[q] = cmul_conj(v[q], hs[t + T * q]);
                        });
                    } else {
                        // general 2-D multiplier H[ky][kx] (column pass): this line is column li
                        const float2* hp = p.h + li + (long long)t * estride;
                        static_for<E>([&](auto Q) {
                            constexpr int q = decltype(Q)::value;
                            v[q] = cmul_conj(v[q], __ldg(hp));
                            hp += step;
                        });
                    }
                }
            }
        }
        // ---- store.  Except for MODE_FWD the register file holds the CONJUGATE of the result.
        float2* __restrict__ dst = p.out + base;
        if constexpr (POST == POST_ADJ) {
            // G_u = conj(v) is the gradient w.r.t. u_i = psi_i t_i.  SURVEY.md 7.1:
            //   dL/ddelta = -k Im(conj(G_u) u),  dL/dbeta = -k Re(conj(G_u) u),  G_i = conj(t) G_u
            const long long dbase = (long long)b * p.db_batch_stride + (long long)li * p.line_stride + t;
            const float2* __restrict__ dbp = p.db + dbase;
            const float2* __restrict__ psip = p.psi + base;
            float2* __restrict__ gp = p.grad + dbase;
            static_for<E / CH>([&](auto C) {
                constexpr int c = decltype(C)::value;
                float2 d[CH], ps[CH];
                static_for<CH>([&](auto I) {
                    constexpr int i = decltype(I)::value;
                    d[i] = dbp[T * (c * CH + i)];
                    ps[i] = psip[T * (c * CH + i)];
                });
                static_for<CH>([&](auto I) {
                    constexpr int i = decltype(I)::value;
                    constexpr int q = c * CH + i;
                    const float2 tr = transmission(d[i], p.k_dz);
                    const float2 u = cmul(ps[i], tr);
                    const float2 w = cmul(u, v[q]);              // u * conj(G_u) = u * v
                    gp[T * q] = make_float2(-p.k_dz * w.y, -p.k_dz * w.x);
                    dst[T * q] = cmul_conj(v[q], tr);           // conj(v) conj(t) = G_u conj(t)
                });
            });
        } else if constexpr (MODE == MODE_FWD) {
            // circular output shift (fftshift), far field only
            float2* __restrict__ dst0 = p.out + (base - (long long)t * estride);
            static_for<E>([&](auto Q) {
                constexpr int q = decltype(Q)::value;
                int e = t + T * q + p.out_shift;
                if (e >= N) e -= N;
                dst0[(long long)e * estride] = v[q];
            });
        } else if constexpr (!COL) {
            static_for<E>([&](auto Q) { constexpr int q = decltype(Q)::value; dst[T * q] = conjf2(v[q]); });
        } else {
            float2* ptr = dst;
            static_for<E>([&](auto Q) { *ptr = conjf2(v[decltype(Q)::value]); ptr += step; });
        }
    }
}

}  // namespace bdof
