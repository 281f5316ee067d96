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

// Streaming of the per-pixel side inputs of the row passes (delta/beta for the transmission; delta/beta
// and the stored psi_i for the adjoint epilogue): every line owns NBUF small shared-memory buffers that
// cp.async.bulk (TMA) refills as soon as the line has consumed them, so the DRAM streams run ahead of
// the math by NBUF-1 chunks (and by a whole FFT pair across tiles) without holding registers.
template <class Cfg, int NSTREAM>
struct StreamCfg {
    static constexpr int T = Cfg::T, E = Cfg::E;
    static constexpr int CHK_WANT = (NSTREAM == 2 ? 256 : 512) / T;                 // elements per lane per chunk
    static constexpr int CHK = NSTREAM == 0 ? E : (CHK_WANT < 1 ? 1 : (CHK_WANT > E ? E : CHK_WANT));
    static constexpr int NCH = E / CHK;
    static constexpr int NBUF = NSTREAM == 0 ? 0 : (NCH < 2 ? NCH : 2);
    static constexpr int CHUNK_ELEMS = CHK * T;                                     // per stream
    static constexpr unsigned CHUNK_BYTES = CHUNK_ELEMS * sizeof(float2);
    static constexpr int LINE_ELEMS = NBUF * NSTREAM * CHUNK_ELEMS;                 // float2 per line
    static_assert(NSTREAM == 0 || (NCH % NBUF == 0), "chunks per tile must be a multiple of the buffer count");
    static_assert(NSTREAM == 0 || (CHUNK_BYTES % 16 == 0), "bulk copies move multiples of 16 bytes");
};

template <class Cfg, int LPC, bool COL, int MODE, int NSTREAM = 0>
struct LineSmem {
    using ST = StreamCfg<Cfg, NSTREAM>;
    // line stride in float2: banks of the LPC interleaved lines must not collide in col mode
    static constexpr int ADJ = COL ? ((((16 / LPC) - Cfg::PADDED) % 16) + 16) % 16 : 0;
    static constexpr int STRIDE = Cfg::PADDED + ADJ;
    static constexpr int TW_ELEMS = (Cfg::TW_TOTAL + 1) & ~1;
    // h is staged in shared memory unless that would overflow the 227 KB CTA limit (8192-long columns)
    static constexpr bool H_IN_SMEM = (MODE == MODE_CONV) &&
        size_t(TW_ELEMS + Cfg::N + STRIDE * LPC + ST::LINE_ELEMS * LPC) * sizeof(float2) + size_t(ST::NBUF * LPC) * 8 <= 227 * 1024;
    static constexpr int H_ELEMS = H_IN_SMEM ? Cfg::N : 0;
    static constexpr int STAGE_ELEMS = ST::LINE_ELEMS * LPC;
    static constexpr int NBARS = ST::NBUF * LPC;
    static constexpr size_t BYTES = size_t(TW_ELEMS + H_ELEMS + STRIDE * LPC + STAGE_ELEMS) * sizeof(float2) + size_t(NBARS) * 8;
};

// ---- mbarrier + bulk async copy (TMA) helpers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Ampere-style 16-byte asynchronous copies (LDGSTS): column tiles are 32..64-byte row segments, too
// small for bulk copies without a tensor map
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// bulk prefetch of a global range into L2 (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Synchronise the threads that share one exchange buffer: a warp (row mode, T <= 32), the warps of one
// line (row mode, T > 32: named barrier 1 + line index) or the whole CTA (column mode, lines interleaved).
template <class Cfg, int LPC, bool COL>
__device__ __forceinline__ void line_sync(int l) {
    if constexpr (!COL && Cfg::T <= 32) __syncwarp();
    else if constexpr (!COL && LPC > 1 && LPC <= 8) {
        // compile-time barrier ids so that ptxas reserves only LPC + 1 of the 16 hardware barriers
        static_for<LPC>([&](auto I) __attribute__((always_inline)) {
            constexpr int i = decltype(I)::value;
            if (l == i) asm volatile("bar.sync %0, %1;" ::"n"(i + 1), "n"(Cfg::T) : "memory");
        });
    } else __syncthreads();
}

// butterflies of radix R on the register file: E/R independent butterflies per thread
template <class Cfg, int R>
__device__ __forceinline__ void reg_butterflies(float2 (&v)[Cfg::E]) {
    constexpr int E = Cfg::E, M = E / R;
    static_for<M>([&](auto MM) __attribute__((always_inline)) {
        constexpr int m = decltype(MM)::value;
        float2 a[R];
        static_for<R>([&](auto RR) __attribute__((always_inline)) { constexpr int r = decltype(RR)::value; a[r] = v[m + r * M]; });
        RegFFT<R, false>::run(a);
        static_for<R>([&](auto RR) __attribute__((always_inline)) { constexpr int r = decltype(RR)::value; v[m + r * M] = a[r]; });
    });
}

// Stockham exchange after a radix-R stage with sub-transform size NS (NS = product of earlier radices)
template <class Cfg, int R, int NS>
__device__ __forceinline__ void exchange(float2 (&v)[Cfg::E], int t, float2* sm) {
    constexpr int E = Cfg::E, T = Cfg::T, M = E / R, R1 = Cfg::R1;
    static_for<M>([&](auto MM) __attribute__((always_inline)) {
        constexpr int m = decltype(MM)::value;
        const int j = t + T * m;
        int base;
        if constexpr (NS == 1) {
            base = j * (R1 + 1);                        // pad(j*R + r) with R == R1
        } else {
            const int o = (j / NS) * (NS * R) + (j % NS);
            base = o + o / R1;                          // r*NS adds r*NS + r*NS/R1 exactly (R1 | NS)
        }
        static_for<R>([&](auto RR) __attribute__((always_inline)) {
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
    static_for<E>([&](auto Q) __attribute__((always_inline)) {
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
struct NoHook { __device__ __forceinline__ void operator()() const {} };

// `after_last_read` runs once all threads sharing the exchange buffer have finished their LAST read of
// it in this transform: from then on the buffer is free (used to prefetch the next tile into it).
template <class Cfg, int LPC, bool COL, class Hook = NoHook>
__device__ __forceinline__ void line_fft(float2 (&v)[Cfg::E], int t, int l, float2* sm, const float2* tw,
                                         [[maybe_unused]] long long* stamps = nullptr, Hook after_last_read = Hook()) {
    constexpr int R1 = Cfg::R1, R2 = Cfg::R2, R3 = Cfg::R3;
    reg_butterflies<Cfg, R1>(v);
#ifdef BDOF_PHASE_TIMING
    asm volatile("" ::"f"(v[0].x), "f"(v[Cfg::E - 1].y));
    if (stamps) stamps[0] = clock64();
#endif
    line_sync<Cfg, LPC, COL>(l);                 // previous readers of the exchange buffer are done
    exchange<Cfg, R1, 1>(v, t, sm);
    line_sync<Cfg, LPC, COL>(l);
    exchange_read_twiddle<Cfg, R2, R1>(v, t, sm, tw);
    if constexpr (R3 == 1) after_last_read();
#ifdef BDOF_PHASE_TIMING
    asm volatile("" ::"f"(v[0].x), "f"(v[Cfg::E - 1].y));
    if (stamps) stamps[1] = clock64();
#endif
    reg_butterflies<Cfg, R2>(v);
    if constexpr (R3 > 1) {
        line_sync<Cfg, LPC, COL>(l);
        exchange<Cfg, R2, R1>(v, t, sm);
        line_sync<Cfg, LPC, COL>(l);
        exchange_read_twiddle<Cfg, R3, R1 * R2>(v, t, sm, tw + Cfg::TW2);
        after_last_read();
        reg_butterflies<Cfg, R3>(v);
    }
}

// Optional per-phase timestamps (-DBDOF_PHASE_TIMING): lane 0 of every warp records clock64() at the
// phase boundaries of its first two tiles into p.dbg[(cta * warps + warp) * 32 + slot].
#ifdef BDOF_PHASE_TIMING
#define BDOF_STAMP(slot)                                                                         \
    do {                                                                                          \
        if (p.dbg != nullptr && (threadIdx.x & 31) == 0 && tile_iter < 2)                         \
            p.dbg[((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32 + tile_iter * 12 + (slot)] = clock64(); \
    } while (0)
#else
#define BDOF_STAMP(slot) do { } while (0)
#endif

template <class Cfg, int LPC, bool COL, int MODE, int PRE, int POST>
__global__ void __launch_bounds__(Cfg::T* LPC) line_kernel(const LineParams p, const int n_tiles) {
    constexpr int N = Cfg::N, T = Cfg::T, E = Cfg::E;
    constexpr int NSTREAM = (PRE == PRE_TRANSMIT) ? 1 : (POST == POST_ADJ ? 2 : 0);
    static_assert(NSTREAM == 0 || !COL, "side-input streaming is a row-pass feature");
    using SM = LineSmem<Cfg, LPC, COL, MODE, NSTREAM>;
    using ST = typename SM::ST;
    extern __shared__ __align__(16) float2 smem[];
    float2* s_tw = smem;
    float2* s_h = smem + SM::TW_ELEMS;
    float2* s_x = s_h + SM::H_ELEMS;
    [[maybe_unused]] float2* s_stage = s_x + SM::STRIDE * LPC;
    [[maybe_unused]] unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_stage + SM::STAGE_ELEMS);

    const int tid = threadIdx.x;
    // Stage the twiddle and h tables with bulk async copies (TMA, cp.async.bulk): one thread issues
    // them, completion is tracked by an mbarrier that is only waited on right before the first use,
    // so the copy overlaps the first tile's loads and stage-1 butterflies.
    __shared__ unsigned long long table_bar;
    // Phase staggering (row passes): the lines that share SM sub-partitions with lines 0..LPC/2-1 start
    // one load phase later, so that in steady state one of them streams memory while the other computes.
#ifndef BDOF_NO_STAGGER
    constexpr bool STAGGER_ON = true;
#else
    constexpr bool STAGGER_ON = false;
#endif
    constexpr bool STAGGER = STAGGER_ON && !COL && (PRE == PRE_TRANSMIT || POST == POST_ADJ) && (LPC % 2 == 0) && (T >= 32) && ((LPC / 2) * (T / 32) % 4 == 0);
    __shared__ volatile int stagger_flag[STAGGER ? LPC / 2 : 1];
    // Next-tile prefetch of the main input into the exchange buffers (convolution passes only): column
    // tiles by cp.async 16-byte pieces issued by every thread, rows by one bulk copy per line.
#ifdef BDOF_NO_ROW_TILE_PREFETCH
    constexpr bool PF_ROW = false;
#else
    constexpr bool PF_ROW = true;
#endif
#ifdef BDOF_NO_COL_TILE_PREFETCH
    constexpr bool PF_COL = false;
#else
    constexpr bool PF_COL = true;
#endif
    constexpr bool PREFETCH = (MODE == MODE_CONV) && (COL ? (PF_COL && LPC >= 2 && (N * LPC / 2) % (T * LPC) == 0) : (PF_ROW && T >= 32));
    __shared__ unsigned long long line_bar[(PREFETCH && !COL) ? LPC : 1];
    constexpr unsigned TW_BYTES = Cfg::TW_TOTAL * sizeof(float2);
    constexpr unsigned H_BYTES = SM::H_IN_SMEM ? N * sizeof(float2) : 0;
    static_assert(TW_BYTES % 16 == 0 && H_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
    if (tid == 0) {
        mbar_init(&table_bar, 1);
        for (int i = 0; i < SM::NBARS; ++i) mbar_init(&s_bar[i], 1);
        if constexpr (PREFETCH && !COL) { for (int i = 0; i < LPC; ++i) mbar_init(&line_bar[i], 1); }
        fence_mbar_init();
    }
    if constexpr (STAGGER) { if (tid < LPC / 2) stagger_flag[tid] = 0; }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&table_bar, TW_BYTES + H_BYTES);
        bulk_g2s(s_tw, p.tw, TW_BYTES, &table_bar);
        if constexpr (SM::H_IN_SMEM) bulk_g2s(s_h, p.h, H_BYTES, &table_bar);
    }
    bool tables_ready = false;
#ifdef BDOF_PDL
    // Programmatic dependent launch (experiment, off: no gain measured on B200 and griddepcontrol.wait
    // itself costs time): everything above touches only constant tables, so it may run while the
    // previous kernel of the stream drains; from here on we read what that kernel wrote.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
    // Cross-kernel overlap: this pass leaves DRAM mostly idle (the field is L2 resident), so its CTAs pull
    // the side inputs of the NEXT pass (delta/beta, stored psi) into L2 while they compute.
    if (p.pf0 != nullptr && tid < 32) {
        constexpr long long PF_CHUNK = 32768;
        const long long per_cta = ((p.pf_bytes / gridDim.x) + PF_CHUNK - 1) / PF_CHUNK * PF_CHUNK;
        const long long begin = (long long)blockIdx.x * per_cta;
        const long long end = begin + per_cta < p.pf_bytes ? begin + per_cta : p.pf_bytes;
        for (long long o = begin + (long long)tid * PF_CHUNK; o < end; o += 32 * PF_CHUNK) {
            const unsigned n = unsigned(end - o < PF_CHUNK ? end - o : PF_CHUNK);
            bulk_prefetch_l2(static_cast<const char*>(p.pf0) + o, n);
            if (p.pf1 != nullptr) bulk_prefetch_l2(static_cast<const char*>(p.pf1) + o, n);
        }
    }

    int l, t;
    if constexpr (COL) { l = tid % LPC; t = tid / LPC; }
    else               { l = tid / T;   t = tid % T; }
    float2* sm = s_x + l * SM::STRIDE;
    const long long estride = COL ? (long long)p.elem_stride : 1LL;
    const long long step = (long long)T * estride;

    [[maybe_unused]] int tile_iter = -1;
    // ---- side-input streaming state (row passes only)
    [[maybe_unused]] float2* my_stage = s_stage + l * ST::LINE_ELEMS;
    [[maybe_unused]] unsigned long long* my_bar = s_bar + l * ST::NBUF;
    [[maybe_unused]] int chunk_seq = 0;             // chunks consumed so far by this line (all tiles)
    // issue chunk c of line `ln` (only the line's first thread talks to the TMA)
    auto stream_issue = [&](long long ln, int c) __attribute__((always_inline)) {
        if constexpr (NSTREAM > 0) {
            if (t == 0) {
                const int bb = int(ln / p.lines_per_batch);
                const int lli = int(ln - (long long)bb * p.lines_per_batch);
                const long long drow = (long long)bb * p.db_batch_stride + (long long)lli * p.line_stride + (long long)c * ST::CHUNK_ELEMS;
                const int buf = c % ST::NBUF;
                float2* dstb = my_stage + buf * (NSTREAM * ST::CHUNK_ELEMS);
                mbar_expect_tx(&my_bar[buf], NSTREAM * ST::CHUNK_BYTES);
                bulk_g2s(dstb, p.db + drow, ST::CHUNK_BYTES, &my_bar[buf]);
                if constexpr (NSTREAM == 2) {
                    const long long prow = (long long)bb * p.batch_stride + (long long)lli * p.line_stride + (long long)c * ST::CHUNK_ELEMS;
                    bulk_g2s(dstb + ST::CHUNK_ELEMS, p.psi + prow, ST::CHUNK_BYTES, &my_bar[buf]);
                }
            }
        }
    };
    // pull the side-input rows (delta/beta [, psi_i]) of line `ln` from DRAM into L2 ahead of their use
    auto stream_prefetch_l2 = [&](long long ln, bool enabled) __attribute__((always_inline)) {
        if constexpr (NSTREAM > 0) {
            if (t == 0 && enabled) {
                const int bb = int(ln / p.lines_per_batch);
                const int lli = int(ln - (long long)bb * p.lines_per_batch);
                const long long drow = (long long)bb * p.db_batch_stride + (long long)lli * p.line_stride;
                constexpr unsigned ROW_BYTES = N * sizeof(float2);
                constexpr unsigned PF = ROW_BYTES < 16384 ? ROW_BYTES : 16384;
                for (unsigned o = 0; o < ROW_BYTES; o += PF) {
                    bulk_prefetch_l2(reinterpret_cast<const char*>(p.db + drow) + o, PF);
                    if constexpr (NSTREAM == 2) {
                        const long long prow = (long long)bb * p.batch_stride + (long long)lli * p.line_stride;
                        bulk_prefetch_l2(reinterpret_cast<const char*>(p.psi + prow) + o, PF);
                    }
                }
            }
        }
    };
    // Line assignment.  Column mode: tile k of this CTA is LPC adjacent columns (they share sectors).
    // Row mode: lines are dealt round-robin over CTAs first and line slots second, so that the last,
    // partially filled round leaves every SM with fewer active warps instead of some SMs with none.
    const long long n_lines = (long long)n_tiles * LPC;
    const long long line_step = (long long)gridDim.x * LPC;
    // (lines narrower than a warp keep the contiguous assignment: the lines sharing a warp must run the
    //  same number of trips because they meet in __syncwarp)
    constexpr bool INTERLEAVE = !COL && T >= 32;
    const long long line0 = INTERLEAVE ? (long long)blockIdx.x + (long long)gridDim.x * l : (long long)blockIdx.x * LPC + l;
    // after chunk c of line `ln` has been consumed by the whole line: refill its buffer with the chunk
    // NBUF ahead (possibly belonging to this line slot's next line)
    auto stream_advance = [&](long long ln, int c) __attribute__((always_inline)) {
        if constexpr (NSTREAM > 0) {
            line_sync<Cfg, LPC, COL>(l);
            const int c2 = c + ST::NBUF;
            if (c2 < ST::NCH) stream_issue(ln, c2);
            else if (ln + line_step < n_lines) stream_issue(ln + line_step, c2 - ST::NCH);
        }
    };
    if constexpr (NSTREAM > 0) {
        if (line0 < n_lines)
            for (int c = 0; c < ST::NBUF; ++c) stream_issue(line0, c);
    }
    // start of row `ln` of the input field (row mode)
    // (row mode; `in_offsets` / `in_line_stride`: lines read straight out of a larger pitched buffer, common.h)
    const long long in_lstride = (!COL && p.in_line_stride != 0) ? (long long)p.in_line_stride : (long long)p.line_stride;
    auto row_ptr = [&](long long ln) __attribute__((always_inline)) {
        const int bb = int(ln / p.lines_per_batch);
        const int lli = int(ln - (long long)bb * p.lines_per_batch);
        const long long b0 = (!COL && p.in_offsets != nullptr) ? p.in_offsets[bb] : (long long)bb * p.batch_stride;
        return p.in + b0 + (long long)lli * in_lstride;
    };
    if constexpr (PREFETCH && !COL) {
        // the first row of every line slot is fetched the same way as all later ones
        if (line0 < n_lines && t == 0) {
            mbar_expect_tx(&line_bar[l], N * (unsigned)sizeof(float2));
            bulk_g2s(sm, row_ptr(line0), N * (unsigned)sizeof(float2), &line_bar[l]);
        }
    }

    if constexpr (STAGGER) {
        if (l >= LPC / 2) { while (stagger_flag[l - LPC / 2] == 0) __nanosleep(64); }
        else if (line0 >= n_lines && t == 0) stagger_flag[l] = 1;      // nothing to do: release the partner
    }
    for (long long line = line0; line < n_lines; line += line_step) {
        ++tile_iter;
        BDOF_STAMP(0);
        const int b = int(line / p.lines_per_batch);
        const int li = int(line - (long long)b * p.lines_per_batch);
        const long long base = (long long)b * p.batch_stride + (long long)li * p.line_stride + (long long)t * estride;

        float2 v[E];
        // ---- load (element t + T*q -> register q)
        constexpr bool ROW_TMA = PREFETCH && !COL;        // every row arrives by bulk copy in the line's exchange buffer
        bool loaded = false;
        if constexpr (ROW_TMA) {
            mbar_wait(&line_bar[l], tile_iter & 1);
            if constexpr (PRE != PRE_TRANSMIT) {
                static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; v[q] = sm[t + T * q]; });
            }
            loaded = true;
        } else if constexpr (PREFETCH) {
            if (tile_iter > 0) {
                // this column tile was prefetched into the exchange buffers while the previous one finished
                cp_async_wait_all();
                __syncthreads();
                const float2* sp = s_x + t * LPC + l;                 // staged as [row][LPC]
                static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; v[q] = sp[T * q * LPC]; });
                loaded = true;
            }
        }
        if (!loaded) {
            const float2* __restrict__ src = COL ? p.in + base : row_ptr(line) + t;
            if constexpr (MODE == MODE_INV) {
                // conj + circular input shift (ifftshift), far-field adjoint only
                const float2* __restrict__ src0 = p.in + (base - (long long)t * estride);
                static_for<E>([&](auto Q) __attribute__((always_inline)) {
                    constexpr int q = decltype(Q)::value;
                    int e = t + T * q + p.in_shift;
                    if (e >= N) e -= N;
                    v[q] = conjf2(src0[(long long)e * estride]);
                });
            } else if constexpr (!COL) {
                static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; v[q] = src[T * q]; });
            } else {
                const float2* ptr = src;
                static_for<E>([&](auto Q) __attribute__((always_inline)) { v[decltype(Q)::value] = *ptr; ptr += step; });
            }
        }
        // pull the NEXT tile's main input from DRAM into L2 while this tile is transformed (the cp.async /
        // bulk copies that fetch it at the end of this tile then hit L2)
        if constexpr (MODE == MODE_CONV) {
            const long long nl = line + line_step;
            if ((p.tune & 1) && nl < n_lines) {
                if constexpr (COL) {
                    const long long tl = nl - l;
                    const int bb = int(tl / p.lines_per_batch);
                    const int c0 = int(tl - (long long)bb * p.lines_per_batch);
                    const float2* sp = p.in + (long long)bb * p.batch_stride + c0;
#pragma unroll 4
                    for (int r = tid; r < N; r += T * LPC)
                        bulk_prefetch_l2(sp + (long long)r * p.elem_stride, LPC * (unsigned)sizeof(float2));
                } else {
                    if (t == 0) {
                        constexpr unsigned ROW_BYTES = N * sizeof(float2);
                        constexpr unsigned PF = ROW_BYTES < 16384 ? ROW_BYTES : 16384;
                        const char* rp = reinterpret_cast<const char*>(row_ptr(nl));
                        for (unsigned o = 0; o < ROW_BYTES; o += PF) bulk_prefetch_l2(rp + o, PF);
                    }
                }
            }
            if constexpr (PRE == PRE_TRANSMIT) {
                if (nl < n_lines) stream_prefetch_l2(nl, (p.tune & 2) != 0);
            }
        }
        // issued from inside the second transform, right after its last read of the exchange buffer
        auto prefetch_next = [&]() __attribute__((always_inline)) {
            if constexpr (PREFETCH) {
                const long long nl = line + line_step;
                if (nl < n_lines && (COL || POST != POST_ADJ)) {
                    line_sync<Cfg, LPC, COL>(l);                          // every reader of the buffer is done
                    if constexpr (COL) {
                        const long long tl = nl - l;                      // first column of the next tile
                        const int bb = int(tl / p.lines_per_batch);
                        const int c0 = int(tl - (long long)bb * p.lines_per_batch);
                        constexpr int PPR = LPC / 2;                      // 16-byte pieces per row segment
                        constexpr int RPI = T * LPC / PPR;                // rows covered per sweep of the CTA
                        constexpr int NP = N / RPI;                       // pieces per thread
                        const int r = tid / PPR, j = tid % PPR;
                        const char* sp = reinterpret_cast<const char*>(p.in + (long long)bb * p.batch_stride + c0) +
                                         ((long long)r * p.elem_stride) * (long long)sizeof(float2) + j * 16;
                        char* dp = reinterpret_cast<char*>(s_x) + r * (LPC * (int)sizeof(float2)) + j * 16;
                        const long long sstep = (long long)RPI * p.elem_stride * (long long)sizeof(float2);
#pragma unroll 8
                        for (int m = 0; m < NP; ++m) {
                            cp_async16(dp, sp);
                            sp += sstep;
                            dp += RPI * LPC * (int)sizeof(float2);
                        }
                        cp_async_commit();
                    } else if constexpr (POST != POST_ADJ) {
                        // (the adjoint epilogue still needs the buffer: it prefetches chunk by chunk itself)
                        if (t == 0) {
                            mbar_expect_tx(&line_bar[l], N * (unsigned)sizeof(float2));
                            bulk_g2s(sm, row_ptr(nl), N * (unsigned)sizeof(float2), &line_bar[l]);
                        }
                    }
                }
            }
        };
#ifndef BDOF_NO_ROWPF
        if constexpr (POST == POST_ADJ) {
            stream_prefetch_l2(line, p.pf_bytes >= 0);      // lands in L2 while the transforms run
            if (line + line_step < n_lines) stream_prefetch_l2(line + line_step, (p.tune & 4) != 0);
        }
#endif
        if constexpr (PRE == PRE_TRANSMIT) {
            if constexpr (ROW_TMA) {
                // ROLLED chunk loop (keeps the instruction footprint small: the 64x unrolled version made
                // instruction-cache misses the top stall): psi sits in the exchange buffer, delta/beta arrive
                // through the streaming buffers; u = psi * t is formed in place, then read into registers
#pragma unroll 1
                for (int c = 0; c < ST::NCH; ++c) {
                    const int buf = c % ST::NBUF;
                    mbar_wait(&my_bar[buf], (chunk_seq / ST::NBUF) & 1);
                    const float2* sb = my_stage + buf * ST::CHUNK_ELEMS + t;
                    float2* sx = sm + c * ST::CHUNK_ELEMS + t;
                    float2 d[ST::CHK];
                    bool small = true;
                    static_for<ST::CHK>([&](auto I) __attribute__((always_inline)) {
                        constexpr int i = decltype(I)::value;
                        d[i] = sb[T * i];
                        small = small && transmission_is_small(d[i], p.k_dz);
                    });
                    if (__all_sync(0xffffffffu, small)) {
                        static_for<ST::CHK>([&](auto I) __attribute__((always_inline)) {
                            constexpr int i = decltype(I)::value;
                            sx[T * i] = cmul1p(sx[T * i], transmission_small_m1(d[i], p.k_dz));
                        });
                    } else {
                        static_for<ST::CHK>([&](auto I) __attribute__((always_inline)) {
                            constexpr int i = decltype(I)::value;
                            sx[T * i] = cmul1p(sx[T * i], transmission_m1(d[i], p.k_dz));
                        });
                    }
                    ++chunk_seq;
                    stream_advance(line, c);
                }
                // every thread touched only its own elements of the buffer: no barrier needed before reading them
                static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; v[q] = sm[t + T * q]; });
            } else {
                // short lines (T < 32): fully unrolled register version
                static_for<ST::NCH>([&](auto C) __attribute__((always_inline)) {
                    constexpr int c = decltype(C)::value;
                    constexpr int buf = c % ST::NBUF;
                    mbar_wait(&my_bar[buf], (chunk_seq / ST::NBUF) & 1);
                    const float2* sb = my_stage + buf * ST::CHUNK_ELEMS + t;
                    static_for<ST::CHK>([&](auto I) __attribute__((always_inline)) {
                        constexpr int i = decltype(I)::value;
                        v[c * ST::CHK + i] = cmul1p(v[c * ST::CHK + i], transmission_any_m1(sb[T * i], p.k_dz));
                    });
                    ++chunk_seq;
                    stream_advance(line, c);
                });
            }
        }
        if constexpr (STAGGER && PRE == PRE_TRANSMIT) {
            // forward row pass: release the partner line once my memory-heavy prologue is done
            if (tile_iter == 0 && l < LPC / 2) {
                asm volatile("" ::"f"(v[0].x), "f"(v[E - 1].y), "f"(v[E / 2].x));
                if (t == 0) stagger_flag[l] = 1;
            }
        }
        BDOF_STAMP(1);
        if (!tables_ready) { mbar_wait(&table_bar, 0); tables_ready = true; }
        BDOF_STAMP(2);
        // ---- transform(s)
        if constexpr (MODE == MODE_FWD || MODE == MODE_INV) {
            line_fft<Cfg, LPC, COL>(v, t, l, sm, s_tw);
        } else {
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                auto hook = [&]() __attribute__((always_inline)) { if (pass == 1) prefetch_next(); };
#ifdef BDOF_PHASE_TIMING
                long long* st = (p.dbg != nullptr && (threadIdx.x & 31) == 0 && tile_iter < 2)
                    ? p.dbg + ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32 + tile_iter * 12 + 3 + pass * 4 : nullptr;
                // ONE call site: a second instantiation would double the instruction footprint
                line_fft<Cfg, LPC, COL>(v, t, l, sm, s_tw, st, hook);
                asm volatile("" ::"f"(v[0].x), "f"(v[E - 1].y));
                if (st) st[2] = clock64();
#else
                line_fft<Cfg, LPC, COL>(v, t, l, sm, s_tw, nullptr, hook);
#endif
                if constexpr (STAGGER && POST == POST_ADJ) {
                    // adjoint row pass: the memory-heavy epilogue is at the END of a tile, so the partner
                    // line is released half a tile in (after the first transform)
                    if (pass == 0 && tile_iter == 0 && l < LPC / 2) {
                        asm volatile("" ::"f"(v[0].x), "f"(v[E - 1].y), "f"(v[E / 2].x));
                        if (t == 0) stagger_flag[l] = 1;
                    }
                }
                if (pass == 0) {
                    // v <- conj(v * h): the second trip then computes conj(IFFT(v h))
                    if constexpr (MODE == MODE_CONV) {
                        const float2* hs = SM::H_IN_SMEM ? s_h : p.h;
                        static_for<E>([&](auto Q) __attribute__((always_inline)) {
                            constexpr int q = decltype(Q)::value;
                            v[q] = cmul_conj(v[q], hs[t + T * q]);
                        });
                    } else {
                        // general 2-D multiplier H[ky][kx] (column pass): this line is column li
                        const float2* hp = p.h + li + (long long)t * estride;
                        static_for<E>([&](auto Q) __attribute__((always_inline)) {
                            constexpr int q = decltype(Q)::value;
                            v[q] = cmul_conj(v[q], __ldg(hp));
                            hp += step;
                        });
                    }
#ifdef BDOF_PHASE_TIMING
                    asm volatile("" ::"f"(v[0].x), "f"(v[E - 1].y));
                    if (st) st[3] = clock64();
#endif
                }
            }
        }
        // ---- store.  Except for MODE_FWD the register file holds the CONJUGATE of the result.
        float2* __restrict__ dst = p.out + base;
        if constexpr (POST == POST_ADJ) {
            // G_u = conj(v) is the gradient w.r.t. u_i = psi_i t_i.  SURVEY.md 7.1:
            //   dL/ddelta = -k Im(conj(G_u) u),  dL/dbeta = -k Re(conj(G_u) u),  G_i = conj(t) G_u
            const long long dbase = (long long)b * p.db_batch_stride + (long long)li * p.line_stride + t;
            float2* __restrict__ gp = p.grad + dbase;
            if constexpr (ROW_TMA) {
                // park conj(G_u) in the (now idle) exchange buffer, then a ROLLED chunk loop streams delta/beta
                // and psi_i; as soon as chunk c has been consumed its part of the buffer receives chunk c of the
                // NEXT row (bulk copy), so the next tile's input lands during this epilogue
                line_sync<Cfg, LPC, COL>(l);
                static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; sm[t + T * q] = v[q]; });
                const long long nl = line + line_step;
                const bool has_next = nl < n_lines;
                const float2* next_row = has_next ? row_ptr(nl) : nullptr;
                if (has_next && t == 0) mbar_expect_tx(&line_bar[l], N * (unsigned)sizeof(float2));
                float2* __restrict__ dstc = dst;
#pragma unroll 1
                for (int c = 0; c < ST::NCH; ++c) {
                    const int buf = c % ST::NBUF;
                    mbar_wait(&my_bar[buf], (chunk_seq / ST::NBUF) & 1);
                    const float2* sb = my_stage + buf * (2 * ST::CHUNK_ELEMS) + t;
                    const float2* sx = sm + c * ST::CHUNK_ELEMS + t;
                    float2 d[ST::CHK];
                    bool small = true;
                    static_for<ST::CHK>([&](auto I) __attribute__((always_inline)) {
                        constexpr int i = decltype(I)::value;
                        d[i] = sb[T * i];
                        small = small && transmission_is_small(d[i], p.k_dz);
                    });
                    float2 trs[ST::CHK];
                    if (__all_sync(0xffffffffu, small)) {
                        static_for<ST::CHK>([&](auto I) __attribute__((always_inline)) { constexpr int i = decltype(I)::value; trs[i] = transmission_small_m1(d[i], p.k_dz); });
                    } else {
                        static_for<ST::CHK>([&](auto I) __attribute__((always_inline)) { constexpr int i = decltype(I)::value; trs[i] = transmission_m1(d[i], p.k_dz); });
                    }
                    static_for<ST::CHK>([&](auto I) __attribute__((always_inline)) {
                        constexpr int i = decltype(I)::value;
                        const float2 vq = sx[T * i];                        // conj(G_u)
                        const float2 u = cmul1p(sb[ST::CHUNK_ELEMS + T * i], trs[i]);      // trs = tau = t - 1
                        const float2 w = cmul(u, vq);                       // u * conj(G_u)
                        gp[c * ST::CHUNK_ELEMS + T * i] = make_float2(-p.k_dz * w.y, -p.k_dz * w.x);
                        dstc[c * ST::CHUNK_ELEMS + T * i] = cmul_conj1p(vq, trs[i]);   // G_u conj(t)
                    });
                    ++chunk_seq;
                    stream_advance(line, c);                                // (line barrier inside)
                    if (has_next && t == 0)
                        bulk_g2s(sm + c * ST::CHUNK_ELEMS, next_row + c * ST::CHUNK_ELEMS, ST::CHUNK_BYTES, &line_bar[l]);
                }
            } else {
                static_for<ST::NCH>([&](auto C) __attribute__((always_inline)) {
                    constexpr int c = decltype(C)::value;
                    constexpr int buf = c % ST::NBUF;
                    mbar_wait(&my_bar[buf], (chunk_seq / ST::NBUF) & 1);
                    const float2* sb = my_stage + buf * (2 * ST::CHUNK_ELEMS) + t;
                    static_for<ST::CHK>([&](auto I) __attribute__((always_inline)) {
                        constexpr int i = decltype(I)::value;
                        constexpr int q = c * ST::CHK + i;
                        const float2 tr = transmission_any_m1(sb[T * i], p.k_dz);     // tau = t - 1
                        const float2 u = cmul1p(sb[ST::CHUNK_ELEMS + T * i], tr);
                        const float2 w = cmul(u, v[q]);              // u * conj(G_u) = u * v
                        gp[T * q] = make_float2(-p.k_dz * w.y, -p.k_dz * w.x);
                        dst[T * q] = cmul_conj1p(v[q], tr);         // conj(v) conj(t) = G_u conj(t)
                    });
                    ++chunk_seq;
                    stream_advance(line, c);
                });
            }
        } else if constexpr (MODE == MODE_FWD) {
            // circular output shift (fftshift), far field only
            float2* __restrict__ dst0 = p.out + (base - (long long)t * estride);
            static_for<E>([&](auto Q) __attribute__((always_inline)) {
                constexpr int q = decltype(Q)::value;
                int e = t + T * q + p.out_shift;
                if (e >= N) e -= N;
                dst0[(long long)e * estride] = v[q];
            });
        } else if constexpr (!COL) {
            static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; dst[T * q] = conjf2(v[q]); });
        } else {
            float2* ptr = dst;
            static_for<E>([&](auto Q) __attribute__((always_inline)) { *ptr = conjf2(v[decltype(Q)::value]); ptr += step; });
        }
        BDOF_STAMP(11);
    }
}

}  // namespace bdof
