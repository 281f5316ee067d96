// Cluster-resident kernels: ONE thread-block cluster carries one field of the batch through ALL slices.
//
// The resident kernels (residentfft.cuh) keep a 64 x 64 field in the registers of one CTA.  A 256 x 256 field (BASELINE config 4:
// 256^3 tomography, tensorflow_recon/reconstruct_fullfield.py) is 512 KB -- too large for one SM, but not for the C = 8 SMs of a
// cluster: CTA c owns N/C = 32 lines of the field, 16 elements per thread in registers, and between two slices the field is
// transposed ACROSS THE CLUSTER through distributed shared memory: every thread pushes its elements into the receive buffers
// of the CTAs that own them in the other layout (st.async to shared::cluster, 128-byte / 256-byte contiguous runs per warp) and
// reports the bytes to the receiver's mbarrier; a CTA starts its next step when its own 64 KB have arrived -- no cluster-wide
// barrier in the slice loop.  psi never touches HBM between slices; per slice the cluster
// streams only the side arrays: (delta, beta) in, the stored psi_i and the transmission stash out (forward); stash and psi_i in,
// gradient out (adjoint).  One launch per direction replaces 2 x n_slice sweep-kernel launches, which at this size are bound by
// launch latency (10.5 us per launch for 10 fields of 256^2, the same for 5 fields: tools/two_stream_probe.py).
//
// Schedule, multiplier tables and semantics are those of the sweep / resident kernels: slice i works along ONE axis a(i) (x for
// even i, y for odd i) and applies that axis' convolution twice with entry i of the error-feedback table sequence.
//
// Thread layout in CTA c (R = N/C lines, T threads per line, E = N/T elements per thread, R*T threads):
//   x steps: l = tid / T (row c R + l), t = tid % T: element q is (y = c R + l, x = t + T q); row-mode line_fft, the exchange
//            buffer of a line is private to its warp (__syncwarp only)
//   y steps: l = tid % R (column c R + l), t = tid / R: element q is (y = t + T q, x = c R + l); a warp is 32 adjacent columns,
//            global accesses are 256-byte rows, interleaved exchange buffer Y[index][column]
// Receive buffers (two, alternating by step parity; the one a step read its field from is that step's FFT scratch):
//   for an x step: [R rows][PADDED] (what line_fft's exchange uses anyway); for a y step: [N rows][R columns].
// A peer may be one step ahead at most (it needs every CTA's data of a step to finish that step), so the pushes of step s + 1
// land in the buffer of the other parity while slow CTAs still work in this one (cluster_transpose).
#pragma once
#include "residentfft.cuh"

namespace bdof {

__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_id_x() {
    unsigned r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned n_clusters_x() {
    unsigned r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
// shared::cluster address of `smem_addr` (a shared::cta address of THIS CTA) in CTA `rank` of the cluster
__device__ __forceinline__ unsigned mapa_shared(unsigned smem_addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f2(unsigned addr, float2 v) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
// asynchronous remote store of 8 bytes that also reports them to the mbarrier `bar` of the destination CTA (both shared::cluster
// addresses): the receiver waits for its bytes, nobody waits for anybody's global stores
__device__ __forceinline__ void st_async_f2(unsigned addr, float2 v, unsigned bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(addr), "f"(v.x), "f"(v.y), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <class Cfg, int C>
struct ClusterSmem {
    static constexpr int N = Cfg::N, R = N / C;
    static_assert(N % C == 0 && Cfg::T <= R, "T consecutive elements of a line stay inside one CTA's share");
    static constexpr int XP = Cfg::PADDED;                        // pitch of the x-step buffer [R][XP]
    static constexpr int X_ELEMS = R * XP, Y_ELEMS = N * R;
    static constexpr int BUF_ELEMS = ((X_ELEMS > Y_ELEMS ? X_ELEMS : Y_ELEMS) + 15) & ~15;
    static constexpr int TW_ELEMS = (Cfg::TW_TOTAL + 1) & ~1;
    static constexpr size_t BYTES = size_t(2 * BUF_ELEMS + TW_ELEMS + 2 * N) * sizeof(float2);
    static_assert(BYTES <= 227 * 1024, "cluster-resident buffers do not fit in shared memory");
};

template <class Cfg, int C, bool COL>
struct ClusterMap {
    static constexpr int N = Cfg::N, T = Cfg::T, R = N / C;
    int l, t, c;
    __device__ __forceinline__ ClusterMap(int tid, int rank) : l(COL ? tid % R : tid / T), t(COL ? tid / R : tid % T), c(rank) {}
    __device__ __forceinline__ int y(int q) const { return COL ? t + T * q : c * R + l; }
    __device__ __forceinline__ int x(int q) const { return COL ? c * R + l : t + T * q; }
    __device__ __forceinline__ int g(int q) const { return y(q) * N + x(q); }
};

// forward transform of LINES interleaved lines (column l of Y[index][l]); natural order in and out
template <class Cfg, int LINES>
__device__ __forceinline__ void fft_interleaved_n(float2 (&v)[Cfg::E], int t, int l, float2* Y, const float2* tw) {
    constexpr int E = Cfg::E, T = Cfg::T, R1 = Cfg::R1, R2 = Cfg::R2;
    static_assert(Cfg::R3 == 1, "two-stage plans only");
    constexpr int M1 = E / R1, M2 = E / R2;
    reg_butterflies<Cfg, R1>(v);
    __syncthreads();                                   // earlier readers of Y are done
    static_for<M1>([&](auto MM) __attribute__((always_inline)) {
        constexpr int m = decltype(MM)::value;
        const int j = t + T * m;
        static_for<R1>([&](auto RR) __attribute__((always_inline)) {
            constexpr int r = decltype(RR)::value;
            Y[(j * R1 + r) * LINES + l] = v[m + r * M1];
        });
    });
    __syncthreads();
    static_for<E>([&](auto Q) __attribute__((always_inline)) {
        constexpr int q = decltype(Q)::value;
        constexpr int m = q % M2, r = q / M2;
        const float2 x = Y[(t + T * q) * LINES + l];
        if constexpr (r == 0) v[q] = x;
        else v[q] = cmul(x, tw[(r - 1) * R1 + (t + T * m) % R1]);
    });
    reg_butterflies<Cfg, R2>(v);
}

// v <- IFFT(h FFT(v)) along the lines of the current layout; `buf` is the step's scratch buffer
template <class Cfg, int C, bool COL>
__device__ __forceinline__ void cluster_conv(float2 (&v)[Cfg::E], const ClusterMap<Cfg, C, COL>& m, float2* buf, const float2* s_tw,
                                             const float2* s_h) {
    constexpr int E = Cfg::E, T = Cfg::T, R = Cfg::N / C;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        if constexpr (COL) fft_interleaved_n<Cfg, R>(v, m.t, m.l, buf, s_tw);
        else line_fft<Cfg, R, false>(v, m.t, m.l, buf + m.l * Cfg::PADDED, s_tw);
        if (pass == 0) {
            static_for<E>([&](auto Q) __attribute__((always_inline)) {
                constexpr int q = decltype(Q)::value;
                v[q] = cmul_conj(v[q], s_h[m.t + T * q]);
            });
        }
    }
    static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; v[q] = conjf2(v[q]); });
}

// Hand the field over to the layout of the other axis: every element goes into the receive buffer `next` (a shared::cta address,
// the same in every CTA) of the CTA that owns it there, as an asynchronous store that reports its 8 bytes to that CTA's
// mbarrier `bar_next`; then wait until the own share (N * R elements) has arrived and read it.  T <= R, so element q of any
// thread goes to CTA (T q) / R -- known at compile time.
// No cluster-wide barrier: barrier.cluster.arrive.release is MEMBAR.ALL.GPU (it waited for the slab / stash stores of the step)
// and its acquire side invalidates L1.  The data flow alone keeps the skew below one step: a CTA finishes step s + 1 -- and only
// then pushes into the buffers of parity s -- after it received the step-(s + 1) data of EVERY peer, which a peer sends at the
// end of its step s, when it is done with its parity-s buffer (its scratch during step s).  `phase` = uses of the barrier so far.
template <class Cfg, int C, bool FROM_COL>
__device__ __forceinline__ void cluster_transpose(float2 (&v)[Cfg::E], int tid, int rank, float2* next, unsigned long long* bar_next, unsigned phase) {
    constexpr int E = Cfg::E, T = Cfg::T, N = Cfg::N, R = N / C, XP = ClusterSmem<Cfg, C>::XP;
    const ClusterMap<Cfg, C, FROM_COL> a(tid, rank);
    const ClusterMap<Cfg, C, !FROM_COL> b(tid, rank);
    const unsigned base = smem_u32(next), bar = smem_u32(bar_next);
    static_for<E>([&](auto Q) __attribute__((always_inline)) {
        constexpr int q = decltype(Q)::value;
        constexpr int peer = (T * q) / R, off = (T * q) % R;      // a.t + off < R
        const unsigned remote = mapa_shared(base, peer), rbar = mapa_shared(bar, peer);
        if constexpr (FROM_COL) {
            // element (row k = a.t + T q, column rank R + a.l) -> x-step buffer of CTA k / R: [k % R][column]
            st_async_f2(remote + unsigned(((a.t + off) * XP + rank * R + a.l) * sizeof(float2)), v[q], rbar);
        } else {
            // element (row rank R + a.l, column k = a.t + T q) -> y-step buffer of CTA k / R: [row][k % R]
            st_async_f2(remote + unsigned(((rank * R + a.l) * R + a.t + off) * sizeof(float2)), v[q], rbar);
        }
    });
    mbar_wait(bar_next, phase & 1);
    // the next use of this barrier: armed before any of its bytes can complete it (they need this arrival)
    if (tid == 0) mbar_expect_tx(bar_next, unsigned(N * R * sizeof(float2)));
    static_for<E>([&](auto Q) __attribute__((always_inline)) {
        constexpr int q = decltype(Q)::value;
        if constexpr (FROM_COL) v[q] = next[b.l * XP + b.t + T * q];           // now an x step: row b.l, column b.t + T q
        else v[q] = next[(b.t + T * q) * R + b.l];                              // now a y step: row b.t + T q, column b.l
    });
    __syncthreads();                                   // step boundary inside the CTA (table staging, scratch of the y steps)
}

template <class Cfg>
__device__ __forceinline__ void cluster_stage_h(const ResidentParams& p, int s, int tid, float2* s_h) {
    constexpr int N = Cfg::N;
    if (tid < N) s_h[(s & 1) * N + tid] = ((s & 1) ? p.hy : p.hx)[(long long)s * N + tid];
}

// pull the side arrays of a step towards L2 ahead of their use (no registers held): one contiguous run per warp and q
template <class Cfg, int C, bool COL>
__device__ __forceinline__ void cluster_prefetch_l2(const float2* base, const ClusterMap<Cfg, C, COL>& m) {
    if ((threadIdx.x & 3) == 0) {                      // 32-byte sectors: every fourth lane
        static_for<Cfg::E>([&](auto Q) __attribute__((always_inline)) {
            constexpr int q = decltype(Q)::value;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + m.g(q)));
        });
    }
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
template <class Cfg, int C, bool COL>
__device__ __forceinline__ void cluster_forward_step(const ResidentParams& p, int s, int n_steps, long long fbase, int tid, int rank,
                                                     float2 (&v)[Cfg::E], float2* buf, float2* next, unsigned long long* bar_next, unsigned phase,
                                                     const float2* s_tw, float2* s_h) {
    constexpr int E = Cfg::E, N = Cfg::N, R = N / C, NT = R * Cfg::T;
    const ClusterMap<Cfg, C, COL> m(tid, rank);
    const int Z = p.n_slice;
    if (s + 1 < Z) {
        const ClusterMap<Cfg, C, !COL> mn(tid, rank);
        cluster_prefetch_l2<Cfg, C, !COL>(p.db + (long long)(s + 1) * p.db_slice_stride + fbase, mn);
    }
    if (s + 1 < n_steps) cluster_stage_h<Cfg>(p, s + 1, tid, s_h);
    if (s > 0) cluster_conv<Cfg, C, COL>(v, m, buf, s_tw, s_h + (s & 1) * N);
    if (s < Z) {
        // (delta, beta) of this slice: in L2 since the previous step; loaded where it is used (16 elements per thread leave no
        // registers to hold it across a convolution under the 128-register cap of a 512-thread CTA)
        float2 d[E];
        {
            const float2* dp = p.db + (long long)s * p.db_slice_stride + fbase;
#pragma unroll
            for (int q = 0; q < E; ++q) d[q] = __ldg(dp + m.g(q));
        }
        if (p.store) {
            float2* sp = p.slab + (long long)s * p.slice_stride + fbase + (long long)rank * (NT * E) + tid;
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
        if (prop) cluster_conv<Cfg, C, COL>(v, m, buf, s_tw, s_h + (s & 1) * N);
    }
    if (s + 1 < n_steps) cluster_transpose<Cfg, C, COL>(v, tid, rank, next, bar_next, phase);
}

template <class Cfg, int C>
__global__ void __launch_bounds__((Cfg::N / C) * Cfg::T, ((Cfg::N / C) * Cfg::T <= 256 ? 2 : 1)) cluster_forward_kernel(const ResidentParams p) {
    using SM = ClusterSmem<Cfg, C>;
    constexpr int E = Cfg::E, N = Cfg::N;
    extern __shared__ __align__(16) float2 smem_cl[];
    __shared__ __align__(8) unsigned long long bars[2];           // bytes received into the buffer of each parity
    float2* bufs = smem_cl;
    float2* s_tw = bufs + 2 * SM::BUF_ELEMS;
    float2* s_h = s_tw + SM::TW_ELEMS;
    const int tid = threadIdx.x, rank = int(cluster_ctarank());
    for (int i = tid; i < Cfg::TW_TOTAL; i += blockDim.x) s_tw[i] = p.tw[i];
    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
        fence_mbar_init();
        mbar_expect_tx(&bars[0], unsigned(N * (N / C) * sizeof(float2)));
        mbar_expect_tx(&bars[1], unsigned(N * (N / C) * sizeof(float2)));
    }
    unsigned uses0 = 0u, uses1 = 0u;                   // completed phases of the two barriers
    const int Z = p.n_slice;
    const bool trail = p.propagate_last && Z > 1;
    const int n_steps = Z + (trail ? 1 : 0);
    for (int b = int(cluster_id_x()); b < p.batch; b += int(n_clusters_x())) {
        const long long fbase = (long long)b * N * N;
        // nobody pushes into a CTA that is still busy with the previous field or has not armed its barriers yet (also orders the
        // table staging); the only cluster-wide barriers of the kernel are this one per field and the one before exit
        cluster_arrive();
        cluster_wait();
        cluster_stage_h<Cfg>(p, 0, tid, s_h);
        float2 v[E];
        {
            const ClusterMap<Cfg, C, false> m0(tid, rank);
#pragma unroll
            for (int q = 0; q < E; ++q) v[q] = __ldg(p.in + m0.g(q));
        }
        __syncthreads();
#pragma unroll 1
        for (int s = 0; s < n_steps; ++s) {
            const int pn = (s + 1) & 1;
            float2* buf = bufs + (s & 1) * SM::BUF_ELEMS;
            float2* next = bufs + pn * SM::BUF_ELEMS;
            if (s & 1) cluster_forward_step<Cfg, C, true>(p, s, n_steps, fbase, tid, rank, v, buf, next, &bars[pn], pn ? uses1 : uses0, s_tw, s_h);
            else       cluster_forward_step<Cfg, C, false>(p, s, n_steps, fbase, tid, rank, v, buf, next, &bars[pn], pn ? uses1 : uses0, s_tw, s_h);
            if (s + 1 < n_steps) { if (pn) ++uses1; else ++uses0; }
        }
        float2* op = p.out + fbase;
        if ((n_steps - 1) & 1) {
            const ClusterMap<Cfg, C, true> m(tid, rank);
#pragma unroll
            for (int q = 0; q < E; ++q) op[m.g(q)] = v[q];
        } else {
            const ClusterMap<Cfg, C, false> m(tid, rank);
#pragma unroll
            for (int q = 0; q < E; ++q) op[m.g(q)] = v[q];
        }
    }
    // a CTA must not exit while peers may still address its shared memory
    cluster_arrive();
    cluster_wait();
}

// ------------------------------------------------------------------------------------------------------------------
// adjoint: the same steps backwards with the conjugate tables (residentfft.cuh)
// ------------------------------------------------------------------------------------------------------------------
template <class Cfg, int C, bool COL>
__device__ __forceinline__ void cluster_adjoint_step(const ResidentParams& p, int s, long long fbase, int tid, int rank,
                                                     float2 (&v)[Cfg::E], float2* buf, float2* next, unsigned long long* bar_next, unsigned phase,
                                                     const float2* s_tw, float2* s_h) {
    constexpr int E = Cfg::E, N = Cfg::N, R = N / C, NT = R * Cfg::T;
    const ClusterMap<Cfg, C, COL> m(tid, rank);
    const int Z = p.n_slice;
    const bool from_stash = p.tstash != nullptr;
    // side arrays of THIS step towards L2 now (they are read after the first convolution), of the NEXT step as well
    if (s < Z) {
        cluster_prefetch_l2<Cfg, C, COL>((from_stash ? p.tstash + (long long)s * p.slice_stride : p.db + (long long)s * p.db_slice_stride) + fbase, m);
        if (tid == 0) bulk_prefetch_l2(p.slab + (long long)s * p.slice_stride + fbase + (long long)rank * (NT * E), unsigned(NT * E * sizeof(float2)));
    }
    if (s >= 1) cluster_stage_h<Cfg>(p, s - 1, tid, s_h);
    const float2* h = s_h + (s & 1) * N;
    if (s == Z) {
        cluster_conv<Cfg, C, COL>(v, m, buf, s_tw, h);     // adjoint of the trailing half propagation
    } else {
        const bool prop = p.propagate_last ? (Z > 1) : (s < Z - 1);
        if (prop) cluster_conv<Cfg, C, COL>(v, m, buf, s_tw, h);
        float2 tau[E];
        {
            const float2* tsrc = (from_stash ? p.tstash + (long long)s * p.slice_stride : p.db + (long long)s * p.db_slice_stride) + fbase;
#pragma unroll
            for (int q = 0; q < E; ++q) tau[q] = __ldg(tsrc + m.g(q));
            if (!from_stash) {
                float2 d[E];
#pragma unroll
                for (int q = 0; q < E; ++q) d[q] = tau[q];
                resident_tau<E>(d, tau, p.k_dz);
            }
        }
        const float2* sp = p.slab + (long long)s * p.slice_stride + fbase + (long long)rank * (NT * E) + tid;
        float2* gp = p.grad + (long long)s * p.slice_stride + fbase;
        const float kdz = p.k_dz;
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const float2 psi = __ldg(sp + q * NT);
            v[q] = cmulc1p(v[q], tau[q]);              // G = G_u conj(t)
            const float2 w = cmulc(psi, v[q]);         // psi conj(G)
            tau[q] = make_float2(-kdz * w.y, -kdz * w.x);
        }
        // the branch stays outside the loops: an asm volatile (the reduction) inside the loop above would pin the order of its
        // loads and serialise their latencies (measured: adjoint 2.4 -> 3.2 ms)
        if (p.accumulate) {
#pragma unroll
            for (int q = 0; q < E; ++q) red_add_f32x2_res(gp + m.g(q), tau[q].x, tau[q].y);
        } else {
#pragma unroll
            for (int q = 0; q < E; ++q) gp[m.g(q)] = tau[q];
        }
        if (s > 0) cluster_conv<Cfg, C, COL>(v, m, buf, s_tw, h);
    }
    if (s > 0) cluster_transpose<Cfg, C, COL>(v, tid, rank, next, bar_next, phase);
}

template <class Cfg, int C>
__global__ void __launch_bounds__((Cfg::N / C) * Cfg::T, ((Cfg::N / C) * Cfg::T <= 256 ? 2 : 1)) cluster_adjoint_kernel(const ResidentParams p) {
    using SM = ClusterSmem<Cfg, C>;
    constexpr int E = Cfg::E, N = Cfg::N;
    extern __shared__ __align__(16) float2 smem_cl[];
    __shared__ __align__(8) unsigned long long bars[2];           // bytes received into the buffer of each parity
    float2* bufs = smem_cl;
    float2* s_tw = bufs + 2 * SM::BUF_ELEMS;
    float2* s_h = s_tw + SM::TW_ELEMS;
    const int tid = threadIdx.x, rank = int(cluster_ctarank());
    for (int i = tid; i < Cfg::TW_TOTAL; i += blockDim.x) s_tw[i] = p.tw[i];
    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
        fence_mbar_init();
        mbar_expect_tx(&bars[0], unsigned(N * (N / C) * sizeof(float2)));
        mbar_expect_tx(&bars[1], unsigned(N * (N / C) * sizeof(float2)));
    }
    unsigned uses0 = 0u, uses1 = 0u;                   // completed phases of the two barriers
    const int Z = p.n_slice;
    const bool trail = p.propagate_last && Z > 1;
    const int s0 = trail ? Z : Z - 1;                  // first step executed
    for (int b = int(cluster_id_x()); b < p.batch; b += int(n_clusters_x())) {
        const long long fbase = (long long)b * N * N;
        cluster_arrive();
        cluster_wait();
        cluster_stage_h<Cfg>(p, s0, tid, s_h);
        float2 v[E];
        if (s0 & 1) {
            const ClusterMap<Cfg, C, true> m(tid, rank);
#pragma unroll
            for (int q = 0; q < E; ++q) v[q] = __ldg(p.in + fbase + m.g(q));
        } else {
            const ClusterMap<Cfg, C, false> m(tid, rank);
#pragma unroll
            for (int q = 0; q < E; ++q) v[q] = __ldg(p.in + fbase + m.g(q));
        }
        __syncthreads();
#pragma unroll 1
        for (int s = s0; s >= 0; --s) {
            // buffer parity follows the step index, as in the forward kernel
            const int pn = (s + 1) & 1;
            float2* buf = bufs + (s & 1) * SM::BUF_ELEMS;
            float2* next = bufs + pn * SM::BUF_ELEMS;
            if (s & 1) cluster_adjoint_step<Cfg, C, true>(p, s, fbase, tid, rank, v, buf, next, &bars[pn], pn ? uses1 : uses0, s_tw, s_h);
            else       cluster_adjoint_step<Cfg, C, false>(p, s, fbase, tid, rank, v, buf, next, &bars[pn], pn ? uses1 : uses0, s_tw, s_h);
            if (s > 0) { if (pn) ++uses1; else ++uses0; }
        }
        if (p.out != nullptr) {
            const ClusterMap<Cfg, C, false> m(tid, rank);  // step 0 is an x step
            float2* op = p.out + fbase;
#pragma unroll
            for (int q = 0; q < E; ++q) op[m.g(q)] = v[q];
        }
    }
    cluster_arrive();
    cluster_wait();
}

}  // namespace bdof
