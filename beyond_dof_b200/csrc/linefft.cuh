// Line kernels: one pass of the separable Fresnel step over a bundle of lines of the field.
//
// A "line" is a row (contiguous) or a column (stride nx) of one batch element.  A CTA holds
// LPC lines entirely on chip: every thread keeps E = N/T elements of one line in registers
// (element t + T*q in register q), the Stockham stages run as in-register radix-R butterflies
// (regfft.cuh) and the inter-stage exchanges go through padded shared memory.  One kernel does
//     load [x transmission exp(k(i delta - beta))] -> FFT -> x h -> IFFT -> [adjoint epilogue] -> store
// so a pass is exactly one HBM read and one HBM write of the field (plus delta/beta once).
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

template <class Cfg, int LPC, bool COL>
struct LineSmem {
    // line stride in float2: banks of the LPC interleaved lines must not collide in col mode
    static constexpr int ADJ = COL ? ((((16 / LPC) - Cfg::PADDED) % 16) + 16) % 16 : 0;
    static constexpr int STRIDE = Cfg::PADDED + ADJ;
    static constexpr size_t BYTES = size_t(STRIDE) * LPC * sizeof(float2);
};

template <class Cfg, int LPC, bool COL>
__device__ __forceinline__ void line_sync() {
    if constexpr (!COL && Cfg::T <= 32) __syncwarp();
    else __syncthreads();
}

// butterflies of radix R on the register file: E/R independent butterflies per thread
template <class Cfg, int R, bool INV>
__device__ __forceinline__ void reg_butterflies(float2 (&v)[Cfg::E]) {
    constexpr int E = Cfg::E, M = E / R;
    static_for<M>([&](auto MM) {
        constexpr int m = decltype(MM)::value;
        float2 a[R];
        static_for<R>([&](auto RR) { constexpr int r = decltype(RR)::value; a[r] = v[m + r * M]; });
        RegFFT<R, INV>::run(a);
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

template <class Cfg>
__device__ __forceinline__ void exchange_read(float2 (&v)[Cfg::E], int t, const float2* sm) {
    constexpr int E = Cfg::E, T = Cfg::T, R1 = Cfg::R1;
    if constexpr (T % R1 == 0) {
        const int base = t + t / R1;
        static_for<E>([&](auto Q) { constexpr int q = decltype(Q)::value; v[q] = sm[base + q * (T + T / R1)]; });
    } else {
        static_for<E>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            const int i = t + T * q;
            v[q] = sm[i + i / R1];
        });
    }
}

// multiply by the stage twiddles W_{NS*R}^{r*k}, k = j mod NS, table row r-1 at tw[(r-1)*NS + k]
template <class Cfg, int R, int NS, bool INV>
__device__ __forceinline__ void stage_twiddle(float2 (&v)[Cfg::E], int t, const float2* __restrict__ tw) {
    constexpr int E = Cfg::E, T = Cfg::T, M = E / R;
    static_for<M>([&](auto MM) {
        constexpr int m = decltype(MM)::value;
        const int k = (t + T * m) % NS;
        static_for<R - 1>([&](auto RR) {
            constexpr int r = decltype(RR)::value + 1;
            const float2 w = __ldg(tw + (r - 1) * NS + k);
            v[m + r * M] = INV ? cmulc(v[m + r * M], w) : cmul(v[m + r * M], w);
        });
    });
}

// full length-N transform of the line held in v (natural order in and out)
template <class Cfg, int LPC, bool COL, bool INV>
__device__ __forceinline__ void line_fft(float2 (&v)[Cfg::E], int t, float2* sm, const float2* __restrict__ tw) {
    constexpr int R1 = Cfg::R1, R2 = Cfg::R2, R3 = Cfg::R3;
    reg_butterflies<Cfg, R1, INV>(v);
    line_sync<Cfg, LPC, COL>();                 // previous readers of the exchange buffer are done
    exchange<Cfg, R1, 1>(v, t, sm);
    line_sync<Cfg, LPC, COL>();
    exchange_read<Cfg>(v, t, sm);
    stage_twiddle<Cfg, R2, R1, INV>(v, t, tw);
    reg_butterflies<Cfg, R2, INV>(v);
    if constexpr (R3 > 1) {
        line_sync<Cfg, LPC, COL>();
        exchange<Cfg, R2, R1>(v, t, sm);
        line_sync<Cfg, LPC, COL>();
        exchange_read<Cfg>(v, t, sm);
        stage_twiddle<Cfg, R3, R1 * R2, INV>(v, t, tw + Cfg::TW2);
        reg_butterflies<Cfg, R3, INV>(v);
    }
}

// Element accessor for one line: element e of the line lives at ptr[e * stride].  In row mode the
// stride is the compile-time constant 1 so every access is base + immediate; in column mode the
// pointer is stepped by a uniform stride so no per-element 64-bit address is kept live.
template <bool COL, int T, int E, class F>
__device__ __forceinline__ void for_each_elem(long long first_off, long long step, F&& f) {
    if constexpr (!COL) {
        static_for<E>([&](auto Q) { constexpr int q = decltype(Q)::value; f(Q, first_off + T * q); });
    } else {
        long long off = first_off;
        static_for<E>([&](auto Q) { f(Q, off); off += step; });
    }
}

template <class Cfg, int LPC, bool COL, int MODE, int PRE, int POST>
__global__ void __launch_bounds__(Cfg::T* LPC) line_kernel(const LineParams p) {
    constexpr int N = Cfg::N, T = Cfg::T, E = Cfg::E;
    using SM = LineSmem<Cfg, LPC, COL>;
    extern __shared__ float2 smem[];

    const int tid = threadIdx.x;
    int l, t;
    if constexpr (COL) { l = tid % LPC; t = tid / LPC; }
    else               { l = tid / T;   t = tid % T; }
    const long long line = (long long)blockIdx.x * LPC + l;
    const int b = int(line / p.lines_per_batch);
    const int li = int(line - (long long)b * p.lines_per_batch);
    const long long base = (long long)b * p.batch_stride + (long long)li * p.line_stride;
    // offset of element t and the step between this thread's consecutive elements (t + T*q)
    const long long first = COL ? (long long)t * p.elem_stride : (long long)t;
    const long long step = COL ? (long long)T * p.elem_stride : (long long)T;
    float2* sm = smem + l * SM::STRIDE;

    float2 v[E];
    // ---- load (element t + T*q -> register q)
    {
        const float2* __restrict__ src = p.in + base;
        if constexpr (MODE == MODE_INV) {
            // circular input shift (ifftshift), far-field adjoint only
            static_for<E>([&](auto Q) {
                constexpr int q = decltype(Q)::value;
                int e = t + T * q + p.in_shift;
                if (e >= N) e -= N;
                v[q] = src[(long long)e * p.elem_stride];
            });
        } else {
            for_each_elem<COL, T, E>(first, step, [&](auto Q, long long off) { v[decltype(Q)::value] = src[off]; });
        }
    }
    if constexpr (PRE == PRE_TRANSMIT) {
        const float2* __restrict__ dbp = p.db + (long long)b * p.db_batch_stride + (long long)li * p.line_stride;
        for_each_elem<COL, T, E>(first, step, [&](auto Q, long long off) {
            constexpr int q = decltype(Q)::value;
            v[q] = cmul(v[q], transmission(dbp[off], p.k_dz));
        });
    }
    // ---- transform
    if constexpr (MODE != MODE_INV) line_fft<Cfg, LPC, COL, false>(v, t, sm, p.tw);
    if constexpr (MODE == MODE_CONV) {
        const float2* __restrict__ hp = p.h + t;
        static_for<E>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            v[q] = cmul(v[q], __ldg(hp + T * q));
        });
    }
    if constexpr (MODE == MODE_CONV2D) {
        // general 2-D multiplier H[ky][kx] (column pass): this line is column li
        const float2* __restrict__ hp = p.h + li;
        for_each_elem<true, T, E>(first, step, [&](auto Q, long long off) {
            constexpr int q = decltype(Q)::value;
            v[q] = cmul(v[q], __ldg(hp + off));
        });
    }
    if constexpr (MODE != MODE_FWD) line_fft<Cfg, LPC, COL, true>(v, t, sm, p.tw);
    // ---- store
    float2* __restrict__ dst = p.out + base;
    if constexpr (POST == POST_ADJ) {
        // v = G_u (gradient w.r.t. u_i = psi_i t_i).  SURVEY.md 7.1:
        //   dL/ddelta = -k Im(conj(G_u) u),  dL/dbeta = -k Re(conj(G_u) u),  G_i = conj(t) G_u
        const long long dbase = (long long)b * p.db_batch_stride + (long long)li * p.line_stride;
        const float2* __restrict__ dbp = p.db + dbase;
        const float2* __restrict__ psip = p.psi + base;
        float2* __restrict__ gp = p.grad + dbase;
        for_each_elem<COL, T, E>(first, step, [&](auto Q, long long off) {
            constexpr int q = decltype(Q)::value;
            const float2 d = dbp[off];
            const float2 ps = psip[off];
            const float2 tr = transmission(d, p.k_dz);
            const float2 u = cmul(ps, tr);
            const float2 w = cmulc(u, v[q]);            // u * conj(G_u)
            gp[off] = make_float2(-p.k_dz * w.y, -p.k_dz * w.x);
            dst[off] = cmulc(v[q], tr);                 // G_u * conj(t)
        });
    } else if constexpr (MODE == MODE_FWD) {
        // circular output shift (fftshift), far field only
        static_for<E>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            int e = t + T * q + p.out_shift;
            if (e >= N) e -= N;
            dst[(long long)e * p.elem_stride] = v[q];
        });
    } else {
        for_each_elem<COL, T, E>(first, step, [&](auto Q, long long off) { dst[off] = v[decltype(Q)::value]; });
    }
}

}  // namespace bdof
