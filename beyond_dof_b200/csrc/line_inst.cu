// Instantiation unit for the line kernels of ONE FFT length: compile with -DBDOF_N=<length>.
#include "../../include/bdof.h"
#include "common.h"
#include <cstring>
#include "linefft.cuh"
#include "pipefft.cuh"
#include "sweepfft.cuh"

using namespace bdof;

#ifndef BDOF_N
#error "compile with -DBDOF_N=<fft length>"
#endif

template <int N> struct CfgFor;
//                                          N    T   R1  R2  R3        row LPC, col LPC
template <> struct CfgFor<64>   { using C = LineCfg<64, 8, 8, 8, 1>;      static constexpr int RL = 16, CL = 16; };
template <> struct CfgFor<128>  { using C = LineCfg<128, 8, 16, 8, 1>;    static constexpr int RL = 16, CL = 16; };
template <> struct CfgFor<256>  { using C = LineCfg<256, 16, 16, 16, 1>;  static constexpr int RL = 8,  CL = 8; };
template <> struct CfgFor<512>  { using C = LineCfg<512, 16, 32, 16, 1>;  static constexpr int RL = 8,  CL = 8; };
template <> struct CfgFor<1024> { using C = LineCfg<1024, 32, 32, 32, 1>; static constexpr int RL = 4,  CL = 4; };
#if !defined(BDOF_ALT) || BDOF_ALT == 0
template <> struct CfgFor<2048> { using C = LineCfg<2048, 32, 64, 32, 1>; static constexpr int RL = 8,  CL = 8; };
#elif BDOF_ALT == 1     // experiment: 32 elements/thread, three stages
template <> struct CfgFor<2048> { using C = LineCfg<2048, 64, 32, 32, 2>; static constexpr int RL = 4,  CL = 8; };
#elif BDOF_ALT == 2     // experiment: 16 elements/thread, three stages
template <> struct CfgFor<2048> { using C = LineCfg<2048, 128, 16, 16, 8>; static constexpr int RL = 2,  CL = 8; };
#endif
template <> struct CfgFor<4096> { using C = LineCfg<4096, 64, 64, 64, 1>; static constexpr int RL = 4,  CL = 4; };
template <> struct CfgFor<8192> { using C = LineCfg<8192, 128, 64, 64, 2>; static constexpr int RL = 1, CL = 2; };

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// persistent launch: one CTA per resident slot, each looping over tiles of LPC lines
template <class Cfg, int LPC, bool COL, int MODE, int PRE, int POST>
static int launch_line(const LineParams& p, long long n_lines, cudaStream_t st) {
    constexpr int NSTREAM = (PRE == PRE_TRANSMIT) ? 1 : (POST == POST_ADJ ? 2 : 0);
    using SM = LineSmem<Cfg, LPC, COL, MODE, NSTREAM>;
    auto kern = line_kernel<Cfg, LPC, COL, MODE, PRE, POST>;
    static int ctas_per_sm = 0;           // per instantiation
    if (ctas_per_sm == 0) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SM::BYTES)));
        int occ = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::T * LPC, SM::BYTES));
        if (occ < 1) return bdof_fail(BDOF_E_UNSUPPORTED, "line kernel does not fit on an SM (%zu bytes smem)", SM::BYTES);
        ctas_per_sm = occ;
    }
    if (n_lines % LPC != 0) return bdof_fail(BDOF_E_UNSUPPORTED, "line count %lld not a multiple of %d", n_lines, LPC);
    const long long n_tiles = n_lines / LPC;
    // leave `bdof_sm_reserve()` SMs free (for NCCL kernels that must run concurrently with the sweep)
    const long long slots = (long long)ctas_per_sm * (sm_count() > bdof_sm_reserve() ? sm_count() - bdof_sm_reserve() : 1);
    const unsigned grid = unsigned(n_tiles < slots ? n_tiles : slots);
    kern<<<grid, Cfg::T * LPC, SM::BYTES, st>>>(p, int(n_tiles));
    return bdof_launch_check("line_kernel");
}

// pipelined column convolution (pipefft.cuh)
template <class Cfg, int LPC, int P>
static int launch_pipe_col(const LineParams& p, long long n_lines, cudaStream_t st) {
    if constexpr (P == 0) {
        return bdof_fail(BDOF_E_UNSUPPORTED, "no pipelined pass for this FFT length");
    } else {
        using SM = PipeSmem<PipeCfg<Cfg, P>, LPC, true>;
        auto kern = pipe_col_conv_kernel<Cfg, LPC, P>;
        static int ctas_per_sm = 0;
        if (ctas_per_sm == 0) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SM::BYTES)));
            int occ = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::T * LPC, SM::BYTES));
            if (occ < 1) return bdof_fail(BDOF_E_UNSUPPORTED, "column pass does not fit on an SM (%zu bytes smem)", SM::BYTES);
            ctas_per_sm = occ > 8 ? 8 : occ;
        }
        if (n_lines % LPC != 0) return bdof_fail(BDOF_E_UNSUPPORTED, "line count %lld not a multiple of %d", n_lines, LPC);
        const long long n_tiles = n_lines / LPC;
        const long long slots = (long long)ctas_per_sm * (sm_count() > bdof_sm_reserve() ? sm_count() - bdof_sm_reserve() : 1);
        const unsigned grid = unsigned(n_tiles < slots ? n_tiles : slots);
        // the field as a matrix [batch * N rows][lines_per_batch columns]; a tile lands in boxes of LPC columns
        alignas(64) CUtensorMap tm, tmo;
        BDOF_TRY(bdof_make_tensor_map(&tm, p.in, (n_lines / p.lines_per_batch) * (long long)Cfg::N, p.lines_per_batch, LPC, SM::BOXR));
        BDOF_TRY(bdof_make_tensor_map(&tmo, p.out, (n_lines / p.lines_per_batch) * (long long)Cfg::N, p.lines_per_batch, LPC, SM::BOXR));
        kern<<<grid, Cfg::T * LPC, SM::BYTES, st>>>(p, int(n_tiles), tm, tmo);
        return bdof_launch_check("pipe_col_conv_kernel");
    }
}

// sweep kernels (sweepfft.cuh): one kernel per slice and direction
template <class Cfg, int LPC, int P, bool COL, bool ADJ, bool ACC = false>
static int launch_sweep(const SweepParams& p0, long long rows, int cols, cudaStream_t st) {
    if constexpr (P == 0) {
        return bdof_fail(BDOF_E_UNSUPPORTED, "no sweep kernel for this FFT length");
    } else {
        using SM = PipeSmem<PipeCfg<Cfg, P>, LPC, COL>;
        auto kern = sweep_kernel<Cfg, LPC, P, COL, ADJ, ACC>;
        static int ctas_per_sm = 0;           // per instantiation: short lines leave room for several CTAs per SM
        if (ctas_per_sm == 0) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SM::BYTES)));
            int occ = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::T * LPC, SM::BYTES));
            if (occ < 1) return bdof_fail(BDOF_E_UNSUPPORTED, "sweep kernel does not fit on an SM (%zu bytes smem)", SM::BYTES);
            ctas_per_sm = occ > 8 ? 8 : occ;
        }
        SweepParams p = p0;
        const long long n_lines = COL ? (rows / Cfg::N) * (long long)cols : rows;
        if (n_lines % LPC != 0) return bdof_fail(BDOF_E_UNSUPPORTED, "line count %lld not a multiple of %d", n_lines, LPC);
        p.n_tiles = int(n_lines / LPC);
        p.lines_per_batch = cols;
        const long long slots = (long long)ctas_per_sm * (sm_count() > bdof_sm_reserve() ? sm_count() - bdof_sm_reserve() : 1);
        const unsigned grid = unsigned(p.n_tiles < slots ? p.n_tiles : slots);
        // x kernels whose lines are whole warps: contiguous row ranges per CTA, last tile partial (2048 rows on 148 SMs:
        // 14 rows per CTA = a tile of 8 and a tile of 6 instead of two rounds of 8 with a quarter of the SMs idle)
        p.total_lines = n_lines;
        p.lines_per_cta = int((n_lines + grid - 1) / grid);
        alignas(64) CUtensorMap tm_in, tm_out, tm_db, tm_grad;
        if constexpr (COL) {
            BDOF_TRY(bdof_make_tensor_map(&tm_in, p.in, rows, cols, LPC, SM::BOXR));
            BDOF_TRY(bdof_make_tensor_map(&tm_out, p.out ? p.out : p.in, rows, cols, LPC, SM::BOXR));
            BDOF_TRY(bdof_make_tensor_map(&tm_db, p.db, rows, cols, LPC, SM::BOXR));
            BDOF_TRY(bdof_make_tensor_map(&tm_grad, p.grad ? (const void*)p.grad : (const void*)p.db, rows, cols, LPC, SM::BOXR));
        } else {
            memset(&tm_in, 0, sizeof(tm_in)); memset(&tm_out, 0, sizeof(tm_out));
            memset(&tm_db, 0, sizeof(tm_db)); memset(&tm_grad, 0, sizeof(tm_grad));
        }
        if (bdof_use_pdl()) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(Cfg::T * LPC); cfg.dynamicSmemBytes = SM::BYTES; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p, tm_in, tm_out, tm_db, tm_grad));
        } else {
            kern<<<grid, Cfg::T * LPC, SM::BYTES, st>>>(p, tm_in, tm_out, tm_db, tm_grad);
        }
        return bdof_launch_check("sweep_kernel");
    }
}

#define BDOF_CAT2(a, b) a##b
#define BDOF_CAT(a, b) BDOF_CAT2(a, b)

int BDOF_CAT(bdof_launch_line_, BDOF_N)(int variant, const LineParams& p, long long n_lines, cudaStream_t st) {
    using C = typename CfgFor<BDOF_N>::C;
    constexpr int RL = CfgFor<BDOF_N>::RL, CL = CfgFor<BDOF_N>::CL;
    switch (variant) {
        case V_ROW_CONV_T:   return launch_line<C, RL, false, MODE_CONV, PRE_TRANSMIT, POST_NONE>(p, n_lines, st);
        case V_ROW_CONV:     return launch_line<C, RL, false, MODE_CONV, PRE_NONE, POST_NONE>(p, n_lines, st);
        case V_ROW_CONV_ADJ: return launch_line<C, RL, false, MODE_CONV, PRE_NONE, POST_ADJ>(p, n_lines, st);
        case V_ROW_FWD:      return launch_line<C, RL, false, MODE_FWD, PRE_NONE, POST_NONE>(p, n_lines, st);
        case V_ROW_INV:      return launch_line<C, RL, false, MODE_INV, PRE_NONE, POST_NONE>(p, n_lines, st);
        case V_COL_CONV:     return launch_line<C, CL, true, MODE_CONV, PRE_NONE, POST_NONE>(p, n_lines, st);
        case V_COL_FWD:      return launch_line<C, CL, true, MODE_FWD, PRE_NONE, POST_NONE>(p, n_lines, st);
        case V_COL_INV:      return launch_line<C, CL, true, MODE_INV, PRE_NONE, POST_NONE>(p, n_lines, st);
        case V_COL_CONV2D:   return launch_line<C, CL, true, MODE_CONV2D, PRE_NONE, POST_NONE>(p, n_lines, st);
        case V_COL_CONV_PIPE: return launch_pipe_col<C, CL, pipe_parts(BDOF_N)>(p, n_lines, st);
    }
    return bdof_fail(BDOF_E_BADARG, "bad variant %d", variant);
}

int BDOF_CAT(bdof_launch_sweep_, BDOF_N)(int col, int adj, const SweepParams& p, long long rows, int cols, cudaStream_t st) {
    using C = typename CfgFor<BDOF_N>::C;
    constexpr int RL = CfgFor<BDOF_N>::RL, CL = CfgFor<BDOF_N>::CL, PP = pipe_parts(BDOF_N);
    if (adj && p.grad_accumulate)
        return col ? launch_sweep<C, CL, PP, true, true, true>(p, rows, cols, st) : launch_sweep<C, RL, PP, false, true, true>(p, rows, cols, st);
    if (col) return adj ? launch_sweep<C, CL, PP, true, true>(p, rows, cols, st) : launch_sweep<C, CL, PP, true, false>(p, rows, cols, st);
    return adj ? launch_sweep<C, RL, PP, false, true>(p, rows, cols, st) : launch_sweep<C, RL, PP, false, false>(p, rows, cols, st);
}
