// Pipelined convolution passes: the NEXT tile lands in a dedicated shared-memory buffer (asynchronous
// copies issued at the start of the current tile) while the current tile is transformed in registers.
//
// The landing buffer is a whole tile (N x LPC complex64, 128 KB for the large plans), so the Stockham
// exchange buffer has to shrink: the exchange between the two butterfly stages runs in P parts through a
// buffer of 1/P line.  Two schemes keep every register index a compile-time constant:
//
//  * natural (R1 == P*T, e.g. 2048 = 64 x 32 with 32 threads): part s carries the stage-1 outputs
//    r in [s*R1/P, (s+1)*R1/P); they are exactly the elements the readers hold in registers q with
//    q mod P == s.
//  * cyclic shift (T == R1 == R2, e.g. 4096 = 64 x 64 with 64 threads): the exchange is a full
//    thread-to-thread transpose, reader t' needs r == t' from everybody.  Thread block a = t / (R1/P)
//    works on cyclically shifted data: its stage-1 outputs sit in registers shifted by a*R1/P (obtained
//    for free from the DFT shift theorem by modulating the INPUT with exp(2 pi i a n / P)), its stage-2
//    inputs arrive shifted the same way (which modulates the OUTPUT with the conjugate factor).  In step
//    s everybody sends register block s and receives register block s; who talks to whom is hidden in
//    two per-thread address offsets.  Inside a convolution the output modulation of the first transform
//    is the input modulation the second one needs (conj trick), so only the tile load and the tile store
//    pay one complex multiply per element with (n mod P) != 0.
#pragma once
#include "linefft.cuh"

namespace bdof {

template <class Cfg_, int P_>
struct PipeCfg {
    using Cfg = Cfg_;
    static constexpr int N = Cfg::N, T = Cfg::T, E = Cfg::E, R1 = Cfg::R1, R2 = Cfg::R2, P = P_;
    static_assert(Cfg::R3 == 1, "two-stage plans only");
    static_assert(E == R1, "one stage-1 butterfly per thread");
    static constexpr bool SHIFT = (P > 1) && (T == R1);
    static_assert(P == 1 || SHIFT || R1 == P * T, "natural partial exchange needs R1 == P*T");
    static_assert(!SHIFT || E == R2, "cyclic-shift exchange needs a square plan");
    static_assert((P & (P - 1)) == 0 && R1 % P == 0, "parts");
    static constexpr int RB = R1 / P;                       // stage-1 outputs per thread per part
    static constexpr int PART = P == 1 ? Cfg::PADDED : (N / R1) * (RB + 1);    // float2 per line per part (one pad per RB)
    static constexpr int TW_ELEMS = SHIFT ? R2 * R1 : Cfg::TW_TOTAL;           // SHIFT: table with the all-ones row 0
};

template <class PC, int LPC, bool COL>
struct PipeSmem {
    static constexpr int ADJ = COL ? ((((16 / LPC) - PC::PART) % 16) + 16) % 16 : 0;
    static constexpr int STRIDE = PC::PART + ADJ;
    static constexpr int TW_ELEMS = (PC::TW_ELEMS + 1) & ~1;
    static constexpr int H_ELEMS = PC::N;
    static constexpr int LAND_ELEMS = PC::N * LPC;
    static constexpr int LAND_OFF = (TW_ELEMS + H_ELEMS + STRIDE * LPC + 15) & ~15;     // TMA destinations: 128-byte aligned
    static constexpr int BOXR = PC::N < 256 ? PC::N : 256;                              // rows per TMA box
    static constexpr int NBOX = PC::N / BOXR;
    static constexpr size_t BYTES = size_t(LAND_OFF + LAND_ELEMS) * sizeof(float2);
    static_assert(BYTES + 256 <= 227 * 1024, "pipelined pass does not fit in shared memory");
};

// per-thread state of the cyclic-shift scheme
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tm, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, const void* src_smem) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1),
                 "r"(smem_u32(src_smem))
                 : "memory");
}
// the same as a reduction: global[...] += shared (fp32 add performed by the TMA unit / L2)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, int c0, int c1, const void* src_smem) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1),
                 "r"(smem_u32(src_smem))
                 : "memory");
}
// fire-and-forget vector reduction at L2: *p += (a, b)
__device__ __forceinline__ void red_add_f32x2(float2* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
// contiguous shared -> global bulk copy (bytes a multiple of 16), tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

template <class PC>
struct ShiftState {
    int a;                     // block of this thread, t / RB
    int rlo;                   // t mod RB
    float2 cmod[PC::P];        // exp(2 pi i a m / P)
};

template <class PC>
__device__ __forceinline__ void shift_init(ShiftState<PC>& st, int t) {
    if constexpr (PC::SHIFT) {
        st.a = t / PC::RB;
        st.rlo = t % PC::RB;
        static_for<PC::P>([&](auto M) __attribute__((always_inline)) {
            constexpr int m = decltype(M)::value;
            float s, c;
            sincospif(2.0f * float((st.a * m) % PC::P) / float(PC::P), &s, &c);
            st.cmod[m] = make_float2(c, s);
        });
    }
}

// forward transform of the line in v (natural order in and out, up to the cyclic-shift modulation)
template <class PC, int LPC, bool COL>
__device__ __forceinline__ void pipe_fft(float2 (&v)[PC::E], int t, int l, float2* sm, const float2* tw,
                                         [[maybe_unused]] const ShiftState<PC>& st) {
    using Cfg = typename PC::Cfg;
    constexpr int E = PC::E, T = PC::T, R1 = PC::R1, R2 = PC::R2, P = PC::P, RB = PC::RB;
    if constexpr (P == 1) {
        line_fft<Cfg, LPC, COL>(v, t, l, sm, tw);
    } else {
        reg_butterflies<Cfg, R1>(v);
        float2 w[E];
        static_for<P>([&](auto S) __attribute__((always_inline)) {
            constexpr int s = decltype(S)::value;
            line_sync<Cfg, LPC, COL>(l);                      // readers of the previous part are done
            // ---- send register block s: outputs of butterfly j = t, local index i0, at j*(RB+1) + i0
            {
                float2* wp = sm + t * (RB + 1);
                static_for<RB>([&](auto I) __attribute__((always_inline)) {
                    constexpr int i0 = decltype(I)::value;
                    wp[i0] = v[s * RB + i0];
                });
            }
            line_sync<Cfg, LPC, COL>(l);
            // ---- receive
            if constexpr (PC::SHIFT) {
                // register block s <- logical stage-2 inputs q = qoff + i0 (written by thread q), r = t
                const int qoff = ((s - st.a) & (P - 1)) * RB;
                const float2* rp = sm + qoff * (RB + 1) + st.rlo;
                const float2* twp = tw + qoff * R1 + t;
                static_for<RB>([&](auto I) __attribute__((always_inline)) {
                    constexpr int i0 = decltype(I)::value;
                    w[s * RB + i0] = cmul(rp[i0 * (RB + 1)], twp[i0 * R1]);
                });
            } else {
                // registers q with q mod P == s: element t + T q = butterfly j = q / P, output r = t + T s
                constexpr int M2 = E / R2;
                static_for<E / P>([&](auto I) __attribute__((always_inline)) {
                    constexpr int q = decltype(I)::value * P + s;
                    constexpr int j = q / P;
                    constexpr int r2 = q / M2, m = q % M2;
                    const float2 x = sm[j * (RB + 1) + t];
                    if constexpr (r2 == 0) w[q] = x;
                    else w[q] = cmul(x, tw[(r2 - 1) * R1 + (t + T * m) % R1]);
                });
            }
        });
        static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; v[q] = w[q]; });
        reg_butterflies<Cfg, R2>(v);
    }
}

// Column convolution pass: out[:, c] = IFFT(h * FFT(in[:, c])) for LPC adjacent columns per tile.
template <class Cfg, int LPC, int P>
__global__ void __launch_bounds__(Cfg::T* LPC) pipe_col_conv_kernel(const LineParams p, const int n_tiles,
                                                                    const __grid_constant__ CUtensorMap tm_in,
                                                                    const __grid_constant__ CUtensorMap tm_out) {
    using PC = PipeCfg<Cfg, P>;
    using SM = PipeSmem<PC, LPC, true>;
    constexpr int N = Cfg::N, T = Cfg::T, E = Cfg::E;
    extern __shared__ __align__(128) float2 smem_pipe[];
    float2* smem = smem_pipe;
    float2* s_tw = smem;
    float2* s_h = smem + SM::TW_ELEMS;
    float2* s_x = s_h + SM::H_ELEMS;
    float2* s_land = smem + SM::LAND_OFF;
    __shared__ unsigned long long table_bar, land_bar;

    const int tid = threadIdx.x;
    constexpr unsigned TW_BYTES = PC::TW_ELEMS * sizeof(float2);
    constexpr unsigned H_BYTES = N * sizeof(float2);
    static_assert(TW_BYTES % 16 == 0 && H_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
    if (tid == 0) {
        mbar_init(&table_bar, 1);
        mbar_init(&land_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    // the tile lands as [row][LPC] through NBOX tensor copies (TMA), issued by one thread
    auto land_issue = [&](long long tile) __attribute__((always_inline)) {
        if (tid == 0) {
            const long long tl = tile * LPC;              // first column of the tile
            const int bb = int(tl / p.lines_per_batch);
            const int c0 = int(tl - (long long)bb * p.lines_per_batch);
            mbar_expect_tx(&land_bar, unsigned(SM::LAND_ELEMS * sizeof(float2)));
#pragma unroll 1
            for (int j = 0; j < SM::NBOX; ++j)
                tma_load_2d(s_land + j * SM::BOXR * LPC, &tm_in, 2 * c0, bb * N + j * SM::BOXR, &land_bar);
        }
    };
    long long tile = blockIdx.x;
    if (tile < n_tiles) land_issue(tile);
    if (tid == 0) {
        mbar_expect_tx(&table_bar, TW_BYTES + H_BYTES);
        bulk_g2s(s_tw, p.tw, TW_BYTES, &table_bar);
        bulk_g2s(s_h, p.h, H_BYTES, &table_bar);
    }
    const int l = tid % LPC, t = tid / LPC;
    float2* sm = s_x + l * SM::STRIDE;
    ShiftState<PC> st;
    shift_init<PC>(st, t);

    bool tables_ready = false;
    [[maybe_unused]] int tile_iter = -1;
#ifdef BDOF_PHASE_TIMING
#define PIPE_STAMP(slot)                                                                          \
    do {                                                                                          \
        if (p.dbg != nullptr && (threadIdx.x & 31) == 0 && tile_iter < 4)                         \
            p.dbg[((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32 + tile_iter * 8 + (slot)] = clock64(); \
    } while (0)
#else
#define PIPE_STAMP(slot) do { } while (0)
#endif
    // Results leave through the landing buffer too: when tile k+1 has landed, every thread SWAPS its
    // finished elements of tile k with its new elements of tile k+1 (same addresses), and one thread
    // sends the buffer to global memory with tensor copies (TMA store) that drain while tile k+1 is
    // transformed.  The load of tile k+2 is issued once those copies have read the buffer.
    float2 v[E];
    auto store_issue = [&](long long done_tile) __attribute__((always_inline)) {
        fence_proxy_async();                          // my generic-proxy writes -> visible to the TMA
        __syncthreads();
        if (tid == 0) {
            const long long tl = done_tile * LPC;
            const int bb = int(tl / p.lines_per_batch);
            const int c0 = int(tl - (long long)bb * p.lines_per_batch);
#pragma unroll 1
            for (int j = 0; j < SM::NBOX; ++j)
                tma_store_2d(&tm_out, 2 * c0, bb * N + j * SM::BOXR, s_land + j * SM::BOXR * LPC);
            bulk_commit_group();
        }
    };
    for (; tile < n_tiles; tile += gridDim.x) {
        ++tile_iter;
        PIPE_STAMP(0);
        mbar_wait(&land_bar, tile_iter & 1);          // the whole tile has landed
        PIPE_STAMP(1);
        {
            float2* sp = s_land + t * LPC + l;
            if (tile_iter == 0) {
                static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; v[q] = sp[T * q * LPC]; });
            } else {
                static_for<E>([&](auto Q) __attribute__((always_inline)) {
                    constexpr int q = decltype(Q)::value;
                    const float2 in = sp[T * q * LPC];
                    sp[T * q * LPC] = v[q];
                    v[q] = in;
                });
            }
            if constexpr (PC::SHIFT) {
                static_for<E>([&](auto Q) __attribute__((always_inline)) {
                    constexpr int q = decltype(Q)::value;
                    if constexpr ((q % P) != 0) v[q] = cmul(v[q], st.cmod[q % P]);
                });
            }
        }
        if (tile_iter > 0) store_issue(tile - gridDim.x);
        else __syncthreads();                         // everybody holds its elements: the buffer is free
        PIPE_STAMP(2);
        if (tile_iter == 0 && tile + gridDim.x < n_tiles) land_issue(tile + gridDim.x);
        PIPE_STAMP(3);
        if (!tables_ready) { mbar_wait(&table_bar, 0); tables_ready = true; }
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            pipe_fft<PC, LPC, true>(v, t, l, sm, s_tw, st);
            if (pass == 0) {
                static_for<E>([&](auto Q) __attribute__((always_inline)) {
                    constexpr int q = decltype(Q)::value;
                    v[q] = cmul_conj(v[q], s_h[t + T * q]);
                });
                // by now the store of the previous tile has long read the buffer: fetch the next tile
                if (tile_iter > 0 && tile + gridDim.x < n_tiles) {
                    if (tid == 0) bulk_wait_group_read0();
                    land_issue(tile + gridDim.x);
                }
            }
#ifdef BDOF_PHASE_TIMING
            asm volatile("" ::"f"(v[0].x), "f"(v[E - 1].y));
            if (pass == 0) PIPE_STAMP(4); else PIPE_STAMP(5);
#endif
        }
        // the registers hold the conjugate of the result (times the shift modulation)
        static_for<E>([&](auto Q) __attribute__((always_inline)) {
            constexpr int q = decltype(Q)::value;
            if constexpr (PC::SHIFT && (q % P) != 0) v[q] = cmul_conj(v[q], st.cmod[q % P]);
            else v[q] = conjf2(v[q]);
        });
        PIPE_STAMP(6);
    }
    // drain: the last tile of this CTA
    if (tile_iter >= 0) {
        const long long last = tile - gridDim.x;
        if (tid == 0) bulk_wait_group_read0();        // (a previous store may still be reading the buffer)
        __syncthreads();
        float2* sp = s_land + t * LPC + l;
        static_for<E>([&](auto Q) __attribute__((always_inline)) { constexpr int q = decltype(Q)::value; sp[T * q * LPC] = v[q]; });
        store_issue(last);
        if (tid == 0) bulk_wait_group_read0();
    }
}

}  // namespace bdof
