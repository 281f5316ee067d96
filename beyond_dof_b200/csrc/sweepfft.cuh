// Sweep kernels: ONE kernel per slice and direction.
//
// The propagator is separable and the two 1-D convolutions commute, so consecutive slices alternate the
// axis order: slice i is propagated as C_{a(i+1)} C_{a(i)} with a(i) = x for even i, y for odd i.  The
// kernel of slice i then works along ONE axis a(i):
//
//   forward   A_{i-1} --C_a--> psi_i --[store psi_i]--> x t_i(delta,beta) --C_a--> A_i
//   adjoint   B_i --C_a^H--> G_u --[G = conj(t_i) G_u, grad_i = -k (Im,Re)(psi_i conj(G))]--> --C_a^H--> B_{i-1}
//
// i.e. two convolutions (four transforms) per load/store of the field, and every side array (delta/beta,
// the stored psi_i, the gradient) is touched in the MIDDLE of the kernel, where its DRAM latency hides
// behind the first convolution.  A tile is LPC lines (rows for the x kernels, adjacent columns for the y
// kernels) held in registers; all loads are asynchronous (TMA) into one tile-sized landing buffer L that
// is reused in turn for the tile itself, delta/beta, the stored psi and -- y kernels -- as the staging
// area of the strided stores (TMA tensor stores; results are SWAPPED into L against the next tile).
// Stage exchanges run in parts through a small buffer (pipefft.cuh).
#pragma once
#include "pipefft.cuh"

namespace bdof {

struct SweepParams {
    const float2* in;        // field in  (row-major [batch][ny][nx]); x kernels only, y kernels use tensor maps
    float2* out;             // field out (may alias in)
    const float2* db;        // (delta, beta) of this slice, row-major
    float2* grad;            // adjoint: gradient of this slice, row-major (may alias db)
                             // forward: nullable transmission stash -- tau_i = exp(k(i delta - beta)) - 1 of this slice is written here
                             // (row-major, may alias db) so that the adjoint kernel of the slice lands t instead of recomputing it
    float2* slab;            // psi entering this slice in TILE layout (forward: written, adjoint: read)
    const float2* h;         // multiplier of this axis (forward) or its conjugate (adjoint), 1/N folded in
    const float2* tw;        // stage twiddles, pipelined layout
    int n_tiles;
    int lines_per_batch;     // columns per batch item (y kernels)
    int lines_per_cta;       // x kernels with whole-warp lines: CTA c owns rows [c * lines_per_cta, ...) and walks them in
    long long total_lines;   //   tiles of LPC rows, the last one partial (0: tiles dealt round-robin as in the y kernels)
    int conv1, conv2;        // run the first / second convolution
    int store_slab;          // forward: write psi_i to the slab
    int store_out;           // write the final field
    int slab_prefetch;       // adjoint: L2-prefetch the slab tile at tile start
    int stagger_ns;          // x kernels: the upper half of the lines starts every convolution this much later, so that
                             // its stage exchanges (shared-memory pipe) overlap the other half's butterflies (FP pipe)
    int grad_accumulate;     // adjoint: ADD the gradient to `grad` (fused accumulation over the fields of a minibatch): vector
                             // reductions at L2 from the x kernels (red.global.add.v2.f32), TMA reduce-stores from the y kernels
    int db_is_t;             // adjoint: `db` holds the stashed transmission tau_i = t_i - 1, not (delta, beta)
    float k_dz;
    long long* dbg;
};

enum { LAND_IN = 0, LAND_DB = 1, LAND_SLAB = 2 };

// ACC (adjoint only): the gradient is ADDED to p.grad (fused accumulation over a minibatch); a template parameter so that the
// plain kernels carry neither the branch nor the reduction instructions
template <class Cfg, int LPC, int P, bool COL, bool ADJ, bool ACC = false>
__global__ void __launch_bounds__(Cfg::T* LPC)
    sweep_kernel(const SweepParams p, const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                 const __grid_constant__ CUtensorMap tm_db, const __grid_constant__ CUtensorMap tm_grad) {
    using PC = PipeCfg<Cfg, P>;
    using SM = PipeSmem<PC, LPC, COL>;
    constexpr int N = Cfg::N, T = Cfg::T, E = Cfg::E;
    constexpr unsigned TILE_BYTES = unsigned(SM::LAND_ELEMS * sizeof(float2));
    extern __shared__ __align__(128) float2 smem_sweep[];
    float2* s_tw = smem_sweep;
    float2* s_h = smem_sweep + SM::TW_ELEMS;
    float2* s_x = s_h + SM::H_ELEMS;
    float2* L = smem_sweep + SM::LAND_OFF;
    __shared__ unsigned long long table_bar, land_bar;

    const int tid = threadIdx.x;
#ifdef BDOF_PHASE_TIMING
#define SWEEP_GSTAMP(slot)                                                                        \
    do {                                                                                          \
        if (p.dbg != nullptr && (threadIdx.x & 31) == 0) {                                        \
            unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));          \
            p.dbg[((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32 + (slot)] = (long long)gt; \
        }                                                                                         \
    } while (0)
#else
#define SWEEP_GSTAMP(slot) do { } while (0)
#endif
    SWEEP_GSTAMP(27);
    constexpr unsigned TW_BYTES = PC::TW_ELEMS * sizeof(float2);
    constexpr unsigned H_BYTES = N * sizeof(float2);
    if (tid == 0) {
        mbar_init(&table_bar, 1);
        mbar_init(&land_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    // ---- asynchronous landing of one tile-sized array into L (one thread talks to the TMA)
    // x kernels: `tile` is the first row of the tile and the CTA owns the contiguous row range [.., row_end)
    constexpr bool RANGES = !COL && T >= 32;
    const long long row_end = RANGES ? ((long long)(blockIdx.x + 1) * p.lines_per_cta < p.total_lines
                                            ? (long long)(blockIdx.x + 1) * p.lines_per_cta : p.total_lines) : 0;
    auto tile_lines = [&](long long tile) __attribute__((always_inline)) {
        if constexpr (RANGES) return int(row_end - tile < LPC ? row_end - tile : LPC);
        else return LPC;
    };
    auto land = [&](int what, long long tile) __attribute__((always_inline)) {
        if (tid == 0) {
            if constexpr (RANGES) {
                const int nl = tile_lines(tile);
                mbar_expect_tx(&land_bar, unsigned(nl) * N * (unsigned)sizeof(float2));
                const float2* src = (what == LAND_IN ? p.in : (what == LAND_DB ? p.db : p.slab)) + tile * (long long)N;
#pragma unroll 1
                for (int j = 0; j < nl; ++j) bulk_g2s(L + j * N, src + j * N, N * (unsigned)sizeof(float2), &land_bar);
                return;
            }
            mbar_expect_tx(&land_bar, TILE_BYTES);
            if (COL && what != LAND_SLAB) {
                const long long tl = tile * LPC;              // first column of the tile
                const int bb = int(tl / p.lines_per_batch);
                const int c0 = int(tl - (long long)bb * p.lines_per_batch);
                const CUtensorMap* tm = (what == LAND_IN) ? &tm_in : &tm_db;
#pragma unroll 1
                for (int j = 0; j < SM::NBOX; ++j) tma_load_2d(L + j * SM::BOXR * LPC, tm, 2 * c0, bb * N + j * SM::BOXR, &land_bar);
            } else {
                // contiguous tile: rows of a row-major array (x kernels) or a slab tile
                const float2* src = (what == LAND_IN ? p.in : (what == LAND_DB ? p.db : p.slab)) + tile * (long long)(N * LPC);
#pragma unroll 1
                for (int j = 0; j < LPC; ++j) bulk_g2s(L + j * N, src + j * N, N * (unsigned)sizeof(float2), &land_bar);
            }
        }
    };
    // adjoint: pull the stored psi of this tile into L2 while the first convolution runs (it lands in L later)
    auto slab_prefetch = [&](long long tile) __attribute__((always_inline)) {
        if (tid == 0 && p.slab_prefetch) {
            const char* src = reinterpret_cast<const char*>(p.slab + tile * (long long)(RANGES ? N : N * LPC));
            const int nl = tile_lines(tile);
#pragma unroll 1
            for (int j = 0; j < nl; ++j) bulk_prefetch_l2(src + (size_t)j * N * sizeof(float2), N * (unsigned)sizeof(float2));
        }
    };
    unsigned land_seq = 0;
    auto land_wait = [&]() __attribute__((always_inline)) { mbar_wait(&land_bar, land_seq & 1); ++land_seq; };

    long long tile = RANGES ? (long long)blockIdx.x * p.lines_per_cta : (long long)blockIdx.x;
    const long long tile_end = RANGES ? row_end : (long long)p.n_tiles;
    const long long tile_step = RANGES ? (long long)LPC : (long long)gridDim.x;
    if (tid == 0) {
        mbar_expect_tx(&table_bar, TW_BYTES + H_BYTES);
        bulk_g2s(s_tw, p.tw, TW_BYTES, &table_bar);
        bulk_g2s(s_h, p.h, H_BYTES, &table_bar);
    }
    // Programmatic dependent launch: everything above touches only constant tables, so it may overlap the
    // tail of the previous kernel of the stream; from here on we read what that kernel wrote.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    SWEEP_GSTAMP(28);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    SWEEP_GSTAMP(29);
    if (tile < tile_end) land(LAND_IN, tile);
    int l, t;
    if constexpr (COL) { l = tid % LPC; t = tid / LPC; }
    else               { l = tid / T;   t = tid % T; }
    float2* sm = s_x + l * SM::STRIDE;
    // my elements e = t + T q live at Lme[q * LQ] in every tile-layout array
    float2* Lme = COL ? (L + t * LPC + l) : (L + l * N + t);
    constexpr int LQ = COL ? T * LPC : T;
    ShiftState<PC> st;
    shift_init<PC>(st, t);

    // y kernels: strided stores leave through L by tensor copies
    auto tma_store_tile = [&](const CUtensorMap* tm, long long done_tile, auto reduce_tag) __attribute__((always_inline)) {
        constexpr bool reduce_add = decltype(reduce_tag)::value;
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            const long long tl = done_tile * LPC;
            const int bb = int(tl / p.lines_per_batch);
            const int c0 = int(tl - (long long)bb * p.lines_per_batch);
            if constexpr (reduce_add) {
#pragma unroll 1
                for (int j = 0; j < SM::NBOX; ++j) tma_reduce_add_2d(tm, 2 * c0, bb * N + j * SM::BOXR, L + j * SM::BOXR * LPC);
            } else {
#pragma unroll 1
                for (int j = 0; j < SM::NBOX; ++j) tma_store_2d(tm, 2 * c0, bb * N + j * SM::BOXR, L + j * SM::BOXR * LPC);
            }
            bulk_commit_group();
        }
    };

#ifdef BDOF_PHASE_TIMING
    int tile_iter = -1;
#define SWEEP_STAMP(slot)                                                                         \
    do {                                                                                          \
        if (p.dbg != nullptr && (threadIdx.x & 31) == 0 && tile_iter < 2)                         \
            p.dbg[((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32 + tile_iter * 16 + (slot)] = clock64(); \
    } while (0)
#else
#define SWEEP_STAMP(slot) do { } while (0)
#endif
    float2 v[E];
    bool pending = false;               // y kernels: the previous tile's result still sits in registers
    bool tables_ready = false;
    const float kdz = p.k_dz;
    const float akdz = fabsf(kdz);
    for (; tile < tile_end; tile += tile_step) {
        const long long tile_off = tile * (long long)(RANGES ? N : N * LPC);
        const bool has_next = tile + tile_step < tile_end;
        const bool active = !RANGES || l < tile_lines(tile);      // partial last tile of an x kernel: whole warps sit out
#ifdef BDOF_PHASE_TIMING
        ++tile_iter;
#endif
        SWEEP_STAMP(0);
        land_wait();                    // this tile's field has landed
        SWEEP_STAMP(1);
        if (COL && pending) {
            static_for<E>([&](auto Q) __attribute__((always_inline)) {
                constexpr int q = decltype(Q)::value;
                const float2 x = Lme[q * LQ];
                Lme[q * LQ] = v[q];
                v[q] = x;
            });
            tma_store_tile(&tm_out, tile - gridDim.x, std::false_type{});
            pending = false;
        } else {
            if (active) static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; v[q] = Lme[q * LQ]; });
            __syncthreads();            // everybody holds its elements: L is free
        }
        SWEEP_STAMP(2);
        // y kernels: L is still being read by a tensor store here; the next landing is issued from inside the
        // following convolution (or right away if there is none), once that store has drained
        int deferred = -1;
        if constexpr (!COL) land(LAND_DB, tile);
        else deferred = LAND_DB;
        if constexpr (ADJ && !COL) slab_prefetch(tile);
        auto deferred_landing = [&]() __attribute__((always_inline)) {
            if ((COL || !ADJ) && deferred >= 0) {
                if (tid == 0) bulk_wait_group_read0();
                land(deferred, deferred == LAND_DB ? tile : tile + tile_step);
                if (ADJ && deferred == LAND_DB) slab_prefetch(tile);
                deferred = -1;
            }
        };
        if (!tables_ready) { mbar_wait(&table_bar, 0); tables_ready = true; }
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            if (half == 1) {
                // ------------------------------------------------ middle: the pointwise part of the slice
                SWEEP_STAMP(3);
                land_wait();            // delta/beta of this tile
                SWEEP_STAMP(4);
                // tau = exp(k(i delta - beta)) - 1 in place (common.h: why t - 1).  ROLLED on purpose: a straight-line version (64 x 20 instructions)
                // pushed the kernel past the instruction cache and cost ~10k cycles on the first tile of every launch
                const int q_end = (active && !(ADJ && p.db_is_t)) ? E : 0;
                // software pipeline: the next group's (delta, beta) are loaded while this group is evaluated (its shared-memory
                // latency hides behind the arithmetic); not in the y kernels of long lines, which have no registers to spare
                constexpr bool PREF = !COL || N <= 1024;
                // G elements per trip: the trip is a chain of dependent latencies (shared-memory load -> ~12 FP ops -> store), so
                // the forward x kernels evaluate 8 pixels per trip for twice the instruction-level parallelism (BDOF_TGROUP)
#ifndef BDOF_TGROUP
#define BDOF_TGROUP 8
#endif
                constexpr int G = (PREF && !ADJ && E % BDOF_TGROUP == 0) ? BDOF_TGROUP : 4;
                [[maybe_unused]] float2 dn[G];
                if constexpr (PREF) {
#pragma unroll
                    for (int i = 0; i < G; ++i) dn[i] = Lme[i * LQ];
                }
#pragma unroll 1
                for (int q0 = 0; q0 < q_end; q0 += G) {
                    float2 d[G];
                    if constexpr (PREF) {
#pragma unroll
                        for (int i = 0; i < G; ++i) d[i] = dn[i];
                        if (q0 + G < E) {
#pragma unroll
                            for (int i = 0; i < G; ++i) dn[i] = Lme[(q0 + G + i) * LQ];
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < G; ++i) d[i] = Lme[(q0 + i) * LQ];
                    }
                    // the largest |k delta|, |k beta| of the group pick the series: the same decision as testing every element
                    // (rounding is monotonic), without the chain of dependent predicate updates
                    float md = 0.f, mb = 0.f;
#pragma unroll
                    for (int i = 0; i < G; ++i) { md = fmaxf(md, fabsf(d[i].x)); mb = fmaxf(mb, fabsf(d[i].y)); }
                    md *= akdz; mb *= akdz;
                    const bool tiny = md <= 0.0625f && mb <= 0.015625f;
                    const bool small = md <= 0.78539816f && mb <= 0.5f;
                    if (__all_sync(0xffffffffu, tiny)) {
#pragma unroll
                        for (int i = 0; i < G; ++i) Lme[(q0 + i) * LQ] = transmission_tiny_m1(d[i], kdz);
                    } else if (__all_sync(0xffffffffu, small)) {
#pragma unroll
                        for (int i = 0; i < G; ++i) Lme[(q0 + i) * LQ] = transmission_small_m1(d[i], kdz);
                    } else {
#pragma unroll
                        for (int i = 0; i < G; ++i) Lme[(q0 + i) * LQ] = transmission_m1(d[i], kdz);
                    }
                }
                SWEEP_STAMP(5);
                if constexpr (!ADJ) {
                    if (p.store_slab && active) {
                        float2* sp = p.slab + tile_off + (Lme - L);
                        static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; sp[q * LQ] = v[q]; });
                    }
                    if (active) static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; v[q] = cmul1p(v[q], Lme[q * LQ]); });
                    if (p.grad != nullptr) {
                        // stash t for the adjoint: L leaves by TMA (tensor store for column tiles, bulk rows otherwise); the next
                        // landing is issued from inside the following convolution once the store has finished reading L
                        if constexpr (COL) {
                            tma_store_tile(&tm_grad, tile, std::false_type{});
                        } else {
                            fence_proxy_async();
                            __syncthreads();
                            if (tid == 0) {
                                const int nl = tile_lines(tile);
                                float2* dst = p.grad + tile_off;
#pragma unroll 1
                                for (int j = 0; j < nl; ++j) bulk_s2g(dst + j * N, L + j * N, N * (unsigned)sizeof(float2));
                                bulk_commit_group();
                            }
                        }
                        if (has_next) deferred = LAND_IN;
                    } else {
                        __syncthreads();    // L is free again
                        if (has_next) land(LAND_IN, tile + tile_step);
                    }
                } else {
                    // G = G_u conj(t), t = 1 + tau
                    if (active) static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; v[q] = cmulc1p(v[q], Lme[q * LQ]); });
                    __syncthreads();
                    land(LAND_SLAB, tile);
                    SWEEP_STAMP(6);
                    land_wait();        // psi_i
                    SWEEP_STAMP(7);
                    // grad = -k (Im, Re)(psi conj(G))
                    if constexpr (COL) {
                        static_for<E>([&](auto Q) __attribute__((always_inline)) {
                            constexpr int q = decltype(Q)::value;
                            const float2 w = cmulc(Lme[q * LQ], v[q]);
                            Lme[q * LQ] = make_float2(-kdz * w.y, -kdz * w.x);
                        });
                        tma_store_tile(&tm_grad, tile, std::integral_constant<bool, ACC>{});
                        if (has_next) deferred = LAND_IN;
                    } else {
                        float2* gp = p.grad + tile_off + (Lme - L);
                        if (active) static_for<E>([&](auto Q) __attribute__((always_inline)) {
                            constexpr int q = decltype(Q)::value;
                            const float2 w = cmulc(Lme[q * LQ], v[q]);
                            if constexpr (ACC) red_add_f32x2(gp + q * LQ, -kdz * w.y, -kdz * w.x);
                            else gp[q * LQ] = make_float2(-kdz * w.y, -kdz * w.x);
                        });
                        __syncthreads();
                        if (has_next) land(LAND_IN, tile + tile_step);
                    }
                }
            }
            if (half == 1) SWEEP_STAMP(8);
            const bool conv = (half == 0) ? (p.conv1 != 0) : (p.conv2 != 0);
            if (conv && active) {
                if constexpr (!COL && LPC >= 2) {
                    if (p.stagger_ns > 0 && l >= LPC / 2) __nanosleep(p.stagger_ns);
                }
                if constexpr (PC::SHIFT) {
                    static_for<E>([&](auto Q) __attribute__((always_inline)) {
                        constexpr int q = decltype(Q)::value;
                        if constexpr ((q % P) != 0) v[q] = cmul(v[q], st.cmod[q % P]);
                    });
                }
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    pipe_fft<PC, LPC, COL>(v, t, l, sm, s_tw, st);
                    if (pass == 0) {
                        static_for<E>([&](auto Q) __attribute__((always_inline)) {
                            constexpr int q = decltype(Q)::value;
                            v[q] = cmul_conj(v[q], s_h[t + T * q]);
                        });
                        deferred_landing();
                    }
                }
                // registers hold the conjugate of the result (times the shift modulation)
                static_for<E>([&](auto Q) __attribute__((always_inline)) {
                    constexpr int q = decltype(Q)::value;
                    if constexpr (PC::SHIFT && (q % P) != 0) v[q] = cmul_conj(v[q], st.cmod[q % P]);
                    else v[q] = conjf2(v[q]);
                });
            } else if (!conv) {
                deferred_landing();
            }
        }
        // ---- result
#ifdef BDOF_PHASE_TIMING
        asm volatile("" ::"f"(v[0].x), "f"(v[E - 1].y));
#endif
        SWEEP_STAMP(9);
        if (p.store_out) {
            if constexpr (COL) {
                pending = true;
            } else {
                float2* op = p.out + tile_off + (Lme - L);
                if (active) static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; op[q * LQ] = v[q]; });
            }
        }
        SWEEP_STAMP(10);
    }
    if (COL && pending) {
        // drain: the last tile of this CTA (L is free: no landing is outstanding)
        if (tid == 0) bulk_wait_group_read0();
        __syncthreads();
        static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; Lme[q * LQ] = v[q]; });
        tma_store_tile(&tm_out, tile - gridDim.x, std::false_type{});
    }
    if ((COL || !ADJ) && tid == 0) bulk_wait_group_read0();
    SWEEP_GSTAMP(30);
}

}  // namespace bdof
