// Resident small-field kernels: ONE CTA carries one field of the batch through ALL slices.
//
// Ptychography probes (64 x 64 in BASELINE config 3; cnn_propagator/ptychography.py:61-76 cuts one such window per scan
// position) are 32 KB: the whole wavefield fits in the registers of one CTA, so psi never leaves the SM between slices and
// the per-slice HBM traffic collapses to the side arrays -- (delta, beta) in, the stored psi_i and the transmission stash out
// (forward); stash, stored psi_i in, gradient out (adjoint).  One launch per direction replaces 2 x n_slice launches of the
// sweep kernels, which at this size are pure launch latency (25 Gpx*slice/s for 128 x 64^2 x 128, round 1).
//
// Schedule = the sweep schedule (sweepfft.cuh): slice i works along ONE axis a(i) (x for even i, y for odd i) and applies
// that axis' convolution twice -- second half of the propagation of slice i-1, modulation, first half of the propagation of
// slice i -- with the multiplier table of entry i of the error-feedback sequence (bdof.cu: build_h_sequence).
//
// Thread layout (T threads per line, E = N/T elements per thread, N lines -> N*T threads):
//   x steps: l = tid / T (row),  t = tid % T: element q is (y = l, x = t + T q); a line's exchange buffer is private to the
//            4 lines of a warp (line_fft in row mode: __syncwarp only)
//   y steps: l = tid % N (column), t = tid / N: element q is (y = t + T q, x = l); a warp is 32 adjacent columns, so global
//            accesses are 256-byte rows and the interleaved exchange buffer Y[index][column] is conflict free
// Between steps the field is transposed through the same shared-memory buffer (pitch N + N/R1: conflict free both ways).
// (delta, beta) / the stash of the NEXT slice are prefetched into registers at the top of every step.
#pragma once
#include "linefft.cuh"

namespace bdof {

// fire-and-forget vector reduction at L2: *p += (a, b)
__device__ __forceinline__ void red_add_f32x2_res(float2* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

struct ResidentParams {
    const float2* in;          // forward: probe [n][n] (shared by the batch); adjoint: G [batch][n][n]
    float2* out;               // forward: field after the object [batch][n][n]; adjoint: G at the entrance plane (nullable)
    const float2* db;          // (delta, beta) [n_slice][batch][n][n]; with `win`: the OBJECT [n_slice][oy][ox] the fields are windows of
    const int* win;            // nullable: (y0, x0) window origin per field (ptychography.py:62-76); pixels outside the object are vacuum
    int oy, ox;                //   object sides (window mode)
    float2* stash;             // forward: nullable, tau_i = t_i - 1 out, layout of db
    const float2* tstash;      // adjoint: nullable, tau_i in (then db is not read)
    float2* grad;              // adjoint: (dL/ddelta, dL/dbeta) out, layout of db (may alias db / tstash)
    float2* slab;              // psi entering slice i, [n_slice][batch][E][N*T] (register order; private to the plan)
    const float2* hx;          // multiplier sequences [n_seq][n], 1/n folded in (adjoint: the conjugate tables)
    const float2* hy;
    const float2* tw;          // stage twiddles, LineCfg layout
    long long db_slice_stride; // elements between slices of db (0: axially repeating object)
    long long slice_stride;    // batch * n * n
    int n_slice, batch;
    int propagate_last;        // TF semantics: the last slice propagates too
    int store;                 // forward: write the slab
    int accumulate;            // adjoint: ADD the gradient to `grad` (bdof_plan_set_grad_accumulate: red.global.add.v2.f32 at L2)
    float k_dz;
};

// forward transform of N interleaved lines (column l of Y[index][l]); natural order in and out
template <class Cfg>
__device__ __forceinline__ void fft_interleaved(float2 (&v)[Cfg::E], int t, int l, float2* Y, const float2* tw) {
    constexpr int N = Cfg::N, E = Cfg::E, T = Cfg::T, R1 = Cfg::R1, R2 = Cfg::R2;
    static_assert(Cfg::R3 == 1, "two-stage plans only");
    constexpr int M1 = E / R1, M2 = E / R2;
    reg_butterflies<Cfg, R1>(v);
    __syncthreads();                                   // earlier readers of Y are done
    static_for<M1>([&](auto MM) __attribute__((always_inline)) {
        constexpr int m = decltype(MM)::value;
        const int j = t + T * m;
        static_for<R1>([&](auto RR) __attribute__((always_inline)) {
            constexpr int r = decltype(RR)::value;
            Y[(j * R1 + r) * N + l] = v[m + r * M1];
        });
    });
    __syncthreads();
    static_for<E>([&](auto Q) __attribute__((always_inline)) {
        constexpr int q = decltype(Q)::value;
        constexpr int m = q % M2, r = q / M2;
        const float2 x = Y[(t + T * q) * N + l];
        if constexpr (r == 0) v[q] = x;
        else v[q] = cmul(x, tw[(r - 1) * R1 + (t + T * m) % R1]);
    });
    reg_butterflies<Cfg, R2>(v);
}

template <class Cfg>
struct ResidentSmem {
    static constexpr int N = Cfg::N;
    static constexpr int PITCH = Cfg::PADDED;                    // N + N/R1: == 8 (mod 16) for 64 and 128
    static constexpr int X_ELEMS = N * PITCH;
    static constexpr int TW_ELEMS = (Cfg::TW_TOTAL + 1) & ~1;
    static constexpr size_t BYTES = size_t(X_ELEMS + TW_ELEMS + 2 * N) * sizeof(float2);
};

template <class Cfg, bool COL>
struct ResidentMap {
    static constexpr int N = Cfg::N, T = Cfg::T;
    int l, t;
    __device__ __forceinline__ ResidentMap(int tid) : l(COL ? tid % N : tid / T), t(COL ? tid / N : tid % T) {}
    // row-major index of element q, and its row / column
    __device__ __forceinline__ int g(int q) const { return COL ? (t + T * q) * N + l : l * N + t + T * q; }
    __device__ __forceinline__ int y(int q) const { return COL ? t + T * q : l; }
    __device__ __forceinline__ int x(int q) const { return COL ? l : t + T * q; }
};

// (delta, beta) of element q of field b in slice s: from the per-field array, or straight from the object through the
// field's window (the object slice is L2 resident: 512 KB at 256^2, and no window copy of the object ever exists in HBM)
template <class Map>
__device__ __forceinline__ float2 resident_load_db(const ResidentParams& p, int s, long long fbase, int wy0, int wx0, const Map& m, int q) {
    if (p.win == nullptr) return __ldg(p.db + (long long)s * p.db_slice_stride + fbase + m.g(q));
    const int yy = wy0 + m.y(q), xx = wx0 + m.x(q);
    if (yy < 0 || yy >= p.oy || xx < 0 || xx >= p.ox) return make_float2(0.f, 0.f);
    return __ldg(p.db + (long long)s * p.db_slice_stride + (long long)yy * p.ox + xx);
}

template <class Cfg, bool COL>
__device__ __forceinline__ void resident_fft(float2 (&v)[Cfg::E], const ResidentMap<Cfg, COL>& m, float2* X, const float2* s_tw) {
    if constexpr (COL) fft_interleaved<Cfg>(v, m.t, m.l, X, s_tw);
    else line_fft<Cfg, Cfg::N, false>(v, m.t, m.l, X + m.l * Cfg::PADDED, s_tw);
}

// v <- IFFT(h FFT(v)) along the lines of the current layout
template <class Cfg, bool COL>
__device__ __forceinline__ void resident_conv(float2 (&v)[Cfg::E], const ResidentMap<Cfg, COL>& m, float2* X, const float2* s_tw,
                                              const float2* s_h) {
    constexpr int E = Cfg::E, T = Cfg::T;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        resident_fft<Cfg, COL>(v, m, X, s_tw);
        if (pass == 0) {
            static_for<E>([&](auto Q) __attribute__((always_inline)) {
                constexpr int q = decltype(Q)::value;
                v[q] = cmul_conj(v[q], s_h[m.t + T * q]);
            });
        }
    }
    static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; v[q] = conjf2(v[q]); });
}

// switch the register layout from the lines of one axis to the lines of the other
template <class Cfg, bool FROM_COL>
__device__ __forceinline__ void resident_transpose(float2 (&v)[Cfg::E], int tid, float2* X) {
    constexpr int E = Cfg::E, T = Cfg::T, N = Cfg::N, P = ResidentSmem<Cfg>::PITCH;
    const ResidentMap<Cfg, FROM_COL> a(tid);
    const ResidentMap<Cfg, !FROM_COL> b(tid);
    __syncthreads();                                   // the transforms are done with X
    static_for<E>([&](auto Q) __attribute__((always_inline)) {
        constexpr int q = decltype(Q)::value;
        if constexpr (FROM_COL) X[(a.t + T * q) * P + a.l] = v[q];
        else X[a.l * P + a.t + T * q] = v[q];
    });
    __syncthreads();
    static_for<E>([&](auto Q) __attribute__((always_inline)) {
        constexpr int q = decltype(Q)::value;
        if constexpr (FROM_COL) v[q] = X[b.l * P + b.t + T * q];
        else v[q] = X[(b.t + T * q) * P + b.l];
    });
    __syncthreads();                                   // X is free for the next exchange
}

// tau = t - 1 of E (delta, beta) pairs, series picked by a warp-uniform vote (common.h)
template <int E>
__device__ __forceinline__ void resident_tau(const float2 (&d)[E], float2 (&tau)[E], float k) {
    bool tiny = true, small = true;
#pragma unroll
    for (int q = 0; q < E; ++q) { tiny = tiny && transmission_is_tiny(d[q], k); small = small && transmission_is_small(d[q], k); }
    if (__all_sync(0xffffffffu, tiny)) {
#pragma unroll
        for (int q = 0; q < E; ++q) tau[q] = transmission_tiny_m1(d[q], k);
    } else if (__all_sync(0xffffffffu, small)) {
#pragma unroll
        for (int q = 0; q < E; ++q) tau[q] = transmission_small_m1(d[q], k);
    } else {
#pragma unroll
        for (int q = 0; q < E; ++q) tau[q] = transmission_m1(d[q], k);
    }
}

// stage the multiplier table of step s into s_h[s & 1] (read after the next block-wide barrier)
template <class Cfg>
__device__ __forceinline__ void resident_stage_h(const ResidentParams& p, int s, int tid, float2* s_h) {
    constexpr int N = Cfg::N;
    if (tid < N) s_h[(s & 1) * N + tid] = ((s & 1) ? p.hy : p.hx)[(long long)s * N + tid];
}

// ------------------------------------------------------------------------------------------------------------------
// forward: step s < n_slice = slice s; step n_slice = the trailing half propagation of the TF semantics
// ------------------------------------------------------------------------------------------------------------------
template <class Cfg, bool COL>
__device__ __forceinline__ void resident_forward_step(const ResidentParams& p, int s, int n_steps, long long fbase, int wy0, int wx0,
                                                      int tid, float2 (&v)[Cfg::E], float2 (&d)[Cfg::E], float2* X, const float2* s_tw,
                                                      float2* s_h) {
    constexpr int E = Cfg::E, N = Cfg::N, NT = Cfg::N * Cfg::T;
    const ResidentMap<Cfg, COL> m(tid);
    const int Z = p.n_slice;
    // prefetch for the NEXT step, in that step's layout
    float2 dn[E] = {};
    if (s + 1 < Z) {
        const ResidentMap<Cfg, !COL> mn(tid);
#pragma unroll
        for (int q = 0; q < E; ++q) dn[q] = resident_load_db(p, s + 1, fbase, wy0, wx0, mn, q);
    }
    if (s + 1 < n_steps) resident_stage_h<Cfg>(p, s + 1, tid, s_h);
    if (s > 0) resident_conv<Cfg, COL>(v, m, X, s_tw, s_h + (s & 1) * N);
    if (s < Z) {
        if (p.store) {
            float2* sp = p.slab + (long long)s * p.slice_stride + fbase + tid;
#pragma unroll
            for (int q = 0; q < E; ++q) sp[q * NT] = v[q];
        }
        float2 tau[E];
        resident_tau<E>(d, tau, p.k_dz);
        if (p.stash != nullptr) {
            float2* tp = p.stash + (long long)s * p.slice_stride + fbase;
#pragma unroll
            for (int q = 0; q < E; ++q) tp[m.g(q)] = tau[q];
        }
#pragma unroll
        for (int q = 0; q < E; ++q) v[q] = cmul1p(v[q], tau[q]);
        const bool prop = p.propagate_last ? (Z > 1) : (s < Z - 1);
        if (prop) resident_conv<Cfg, COL>(v, m, X, s_tw, s_h + (s & 1) * N);
    }
    if (s + 1 < n_steps) resident_transpose<Cfg, COL>(v, tid, X);
#pragma unroll
    for (int q = 0; q < E; ++q) d[q] = dn[q];
}

// MINB = 2: registers capped at 64 so that two CTAs share an SM (experiment, BDOF_RESIDENT_2CTA=1: measured 4-6 % slower than one
// CTA per SM -- the kernels are bound by issue slots and shared-memory wavefronts, not by latency)
template <class Cfg, int MINB = 1>
__global__ void __launch_bounds__(Cfg::N* Cfg::T, MINB) resident_forward_kernel(const ResidentParams p) {
    using SM = ResidentSmem<Cfg>;
    constexpr int E = Cfg::E, N = Cfg::N;
    extern __shared__ __align__(16) float2 smem_res[];
    float2* X = smem_res;
    float2* s_tw = X + SM::X_ELEMS;
    float2* s_h = s_tw + SM::TW_ELEMS;
    const int tid = threadIdx.x;
    for (int i = tid; i < Cfg::TW_TOTAL; i += blockDim.x) s_tw[i] = p.tw[i];
    const int Z = p.n_slice;
    const bool trail = p.propagate_last && Z > 1;
    const int n_steps = Z + (trail ? 1 : 0);
    for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
        const long long fbase = (long long)b * N * N;
        __syncthreads();                               // tables staged / the previous field is done with the shared buffers
        resident_stage_h<Cfg>(p, 0, tid, s_h);
        const int wy0 = p.win ? p.win[2 * b] : 0, wx0 = p.win ? p.win[2 * b + 1] : 0;
        float2 v[E], d[E];
        {
            const ResidentMap<Cfg, false> m0(tid);
#pragma unroll
            for (int q = 0; q < E; ++q) { v[q] = __ldg(p.in + m0.g(q)); d[q] = resident_load_db(p, 0, fbase, wy0, wx0, m0, q); }
        }
        __syncthreads();
#pragma unroll 1
        for (int s = 0; s < n_steps; ++s) {
            if (s & 1) resident_forward_step<Cfg, true>(p, s, n_steps, fbase, wy0, wx0, tid, v, d, X, s_tw, s_h);
            else       resident_forward_step<Cfg, false>(p, s, n_steps, fbase, wy0, wx0, tid, v, d, X, s_tw, s_h);
        }
        float2* op = p.out + fbase;
        if ((n_steps - 1) & 1) {
            const ResidentMap<Cfg, true> m(tid);
#pragma unroll
            for (int q = 0; q < E; ++q) op[m.g(q)] = v[q];
        } else {
            const ResidentMap<Cfg, false> m(tid);
#pragma unroll
            for (int q = 0; q < E; ++q) op[m.g(q)] = v[q];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// adjoint: the same steps backwards with the conjugate tables
//   B_s --C_a^H--> G_u --[G = conj(t_s) G_u ; grad_s = -k (Im, Re)(psi_s conj(G))]--> --C_a^H--> B_{s-1}
// ------------------------------------------------------------------------------------------------------------------
// AHEAD: the side arrays of the next step are prefetched into registers one step ahead (one CTA per SM); otherwise they are
// loaded at the top of their own step and land during its first convolution (register-capped variant, two CTAs per SM)
template <class Cfg, bool COL, bool AHEAD>
__device__ __forceinline__ void resident_adjoint_step(const ResidentParams& p, int s, long long fbase, int wy0, int wx0, int tid,
                                                      float2 (&v)[Cfg::E], float2 (&d)[Cfg::E], float2 (&psi)[Cfg::E], float2* X,
                                                      const float2* s_tw, float2* s_h) {
    constexpr int E = Cfg::E, N = Cfg::N, NT = Cfg::N * Cfg::T;
    const ResidentMap<Cfg, COL> m(tid);
    const int Z = p.n_slice;
    const bool from_stash = p.tstash != nullptr;
    if constexpr (!AHEAD) {
        if (s < Z) {
            const float2* tsrc = p.tstash + (long long)s * p.slice_stride + fbase;
            const float2* sp = p.slab + (long long)s * p.slice_stride + fbase + tid;
#pragma unroll
            for (int q = 0; q < E; ++q) {
                d[q] = from_stash ? __ldg(tsrc + m.g(q)) : resident_load_db(p, s, fbase, wy0, wx0, m, q);
                psi[q] = __ldg(sp + q * NT);
            }
        }
    }
    // prefetch for the NEXT step (slice s - 1), in that step's layout
    [[maybe_unused]] float2 dn[AHEAD ? E : 1] = {}, pn[AHEAD ? E : 1] = {};
    if (AHEAD && s >= 1 && s - 1 < Z) {
        const ResidentMap<Cfg, !COL> mn(tid);
        const float2* tsrc = p.tstash + (long long)(s - 1) * p.slice_stride + fbase;
        const float2* sp = p.slab + (long long)(s - 1) * p.slice_stride + fbase + tid;
        if constexpr (AHEAD) {
#pragma unroll
            for (int q = 0; q < E; ++q) {
                dn[q] = from_stash ? __ldg(tsrc + mn.g(q)) : resident_load_db(p, s - 1, fbase, wy0, wx0, mn, q);
                pn[q] = __ldg(sp + q * NT);
            }
        }
    }
    if (s >= 1) resident_stage_h<Cfg>(p, s - 1, tid, s_h);
    const float2* h = s_h + (s & 1) * N;
    if (s == Z) {
        resident_conv<Cfg, COL>(v, m, X, s_tw, h);     // adjoint of the trailing half propagation
    } else {
        const bool prop = p.propagate_last ? (Z > 1) : (s < Z - 1);
        if (prop) resident_conv<Cfg, COL>(v, m, X, s_tw, h);
        float2 tau[E];
        if (from_stash) {
#pragma unroll
            for (int q = 0; q < E; ++q) tau[q] = d[q];
        } else {
            resident_tau<E>(d, tau, p.k_dz);
        }
        float2* gp = p.grad + (long long)s * p.slice_stride + fbase;
        const float kdz = p.k_dz;
#pragma unroll
        for (int q = 0; q < E; ++q) {
            v[q] = cmulc1p(v[q], tau[q]);              // G = G_u conj(t)
            const float2 w = cmulc(psi[q], v[q]);      // psi conj(G)
            tau[q] = make_float2(-kdz * w.y, -kdz * w.x);
        }
        if (p.accumulate) {                            // the branch outside the loops: no asm volatile among the arithmetic
#pragma unroll
            for (int q = 0; q < E; ++q) red_add_f32x2_res(gp + m.g(q), tau[q].x, tau[q].y);
        } else {
#pragma unroll
            for (int q = 0; q < E; ++q) gp[m.g(q)] = tau[q];
        }
        if (s > 0) resident_conv<Cfg, COL>(v, m, X, s_tw, h);
    }
    if (s > 0) resident_transpose<Cfg, COL>(v, tid, X);
    if constexpr (AHEAD) {
#pragma unroll
        for (int q = 0; q < E; ++q) { d[q] = dn[q]; psi[q] = pn[q]; }
    }
}

template <class Cfg, int MINB = 1>
__global__ void __launch_bounds__(Cfg::N* Cfg::T, MINB) resident_adjoint_kernel(const ResidentParams p) {
    constexpr bool AHEAD = (MINB == 1);
    using SM = ResidentSmem<Cfg>;
    constexpr int E = Cfg::E, N = Cfg::N, NT = Cfg::N * Cfg::T;
    extern __shared__ __align__(16) float2 smem_res[];
    float2* X = smem_res;
    float2* s_tw = X + SM::X_ELEMS;
    float2* s_h = s_tw + SM::TW_ELEMS;
    const int tid = threadIdx.x;
    for (int i = tid; i < Cfg::TW_TOTAL; i += blockDim.x) s_tw[i] = p.tw[i];
    const int Z = p.n_slice;
    const bool trail = p.propagate_last && Z > 1;
    const int s0 = trail ? Z : Z - 1;                  // first step executed
    const bool from_stash = p.tstash != nullptr;
    for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
        const long long fbase = (long long)b * N * N;
        __syncthreads();
        resident_stage_h<Cfg>(p, s0, tid, s_h);
        const int wy0 = p.win ? p.win[2 * b] : 0, wx0 = p.win ? p.win[2 * b + 1] : 0;
        float2 v[E], d[E], psi[E];
        auto first_loads = [&](auto map) __attribute__((always_inline)) {
#pragma unroll
            for (int q = 0; q < E; ++q) v[q] = __ldg(p.in + fbase + map.g(q));
            if (AHEAD && s0 < Z) {
                const float2* tsrc = p.tstash + (long long)s0 * p.slice_stride + fbase;
                const float2* sp = p.slab + (long long)s0 * p.slice_stride + fbase + tid;
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    d[q] = from_stash ? __ldg(tsrc + map.g(q)) : resident_load_db(p, s0, fbase, wy0, wx0, map, q);
                    psi[q] = __ldg(sp + q * NT);
                }
            } else {
#pragma unroll
                for (int q = 0; q < E; ++q) { d[q] = make_float2(0.f, 0.f); psi[q] = make_float2(0.f, 0.f); }
            }
        };
        if (s0 & 1) first_loads(ResidentMap<Cfg, true>(tid));
        else        first_loads(ResidentMap<Cfg, false>(tid));
        __syncthreads();
#pragma unroll 1
        for (int s = s0; s >= 0; --s) {
            if (s & 1) resident_adjoint_step<Cfg, true, AHEAD>(p, s, fbase, wy0, wx0, tid, v, d, psi, X, s_tw, s_h);
            else       resident_adjoint_step<Cfg, false, AHEAD>(p, s, fbase, wy0, wx0, tid, v, d, psi, X, s_tw, s_h);
        }
        if (p.out != nullptr) {
            const ResidentMap<Cfg, false> m(tid);       // step 0 is an x step
            float2* op = p.out + fbase;
#pragma unroll
            for (int q = 0; q < E; ++q) op[m.g(q)] = v[q];
        }
    }
}

}  // namespace bdof
