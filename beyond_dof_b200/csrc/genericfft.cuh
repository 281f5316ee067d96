// Mixed-radix line passes for field sides that are not one of the power-of-two lengths of linefft.cuh
// (the reference's ptychography probes are 72 x 72 and 18 x 18: tensorflow_recon/reconstruct_ptycho.py).
//
// Same contract as line_kernel (LineParams, the V_ROW_* / V_COL_* pass variants), any length
// N = 2^a 3^b 5^c 7^d <= GEN_MAX_N.  A CTA owns a tile of LPC lines in shared memory (cooperative, coalesced
// load and store; column tiles are LPC adjacent columns); one warp transforms one line with a Stockham
// autosort FFT that ping-pongs between two buffers, radices 4, 2, 3, 5, 7, twiddles W_N^k from a full table
// generated in float64.  The inverse transform is conj(FFT(conj(.))) as in the power-of-two kernels.
//
// Every function that carries index arithmetic is __host__ __device__ and takes (thread id, thread count), so
// that the CPU test suite can step the same code thread by thread (tests/genfft_host.cu) -- test
// infrastructure only: libbdof.so exports no host path.
#pragma once
#include "common.h"
#include <math.h>

namespace bdof {

constexpr int GEN_MAX_N = 2048;
constexpr int GEN_MAX_STAGES = 12;
constexpr int GEN_MAX_RADIX = 7;

enum GenMode { GEN_CONV = 0, GEN_FWD = 1, GEN_INV = 2, GEN_CONV2D = 3 };

struct GenArgs {
    LineParams p;            // p.tw = full table W_N^k = exp(-2 pi i k / N), k in [0, N)
    long long n_lines;
    int n;                   // line length
    int lpc;                 // lines per tile (= warps per CTA)
    int mode;                // GenMode
    int pre_transmit;        // multiply the input by t(delta, beta) (row passes)
    int post_adj;            // adjoint epilogue (row passes)
    int col;                 // lines are columns
    int n_stages;
    int radix[GEN_MAX_STAGES];
};

#define GEN_HD __host__ __device__ __forceinline__

GEN_HD float2 g_cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
GEN_HD float2 g_conj(float2 a) { return make_float2(a.x, -a.y); }

// tau = exp(i k delta) exp(-k beta) - 1: the device build uses the same routine as every other kernel (common.h)
GEN_HD float2 g_transmission_m1(float2 db, float k) {
#ifdef __CUDA_ARCH__
    return transmission_any_m1(db, k);
#else
    const double m = exp(-double(k) * db.y), x = double(k) * db.x;
    return make_float2(float(m * cos(x) - 1.0), float(m * sin(x)));
#endif
}
// a * (1 + tau)
GEN_HD float2 g_cmul1p(float2 a, float2 tau) { return make_float2(a.x + (a.x * tau.x - a.y * tau.y), a.y + (a.x * tau.y + a.y * tau.x)); }

// radices of n: 4s first, then 2, 3, 5, 7; returns the number of stages or 0 when n has another prime factor
inline __host__ int gen_factorize(int n, int* radix) {
    if (n < 2 || n > GEN_MAX_N) return 0;
    int ns = 0;
    while (n % 4 == 0) { if (ns == GEN_MAX_STAGES) return 0; radix[ns++] = 4; n /= 4; }
    const int primes[4] = {2, 3, 5, 7};
    for (int pi = 0; pi < 4; ++pi)
        while (n % primes[pi] == 0) { if (ns == GEN_MAX_STAGES) return 0; radix[ns++] = primes[pi]; n /= primes[pi]; }
    return n == 1 ? ns : 0;
}

// (line, element) handled by cooperative slot idx of a tile: column tiles interleave the LPC adjacent columns so that
// consecutive threads touch consecutive addresses
GEN_HD void gen_slot(const GenArgs& a, int idx, int* l, int* e) {
    if (a.col) { *l = idx % a.lpc; *e = idx / a.lpc; }
    else       { *l = idx / a.n;   *e = idx % a.n; }
}

// tile load: global -> A[l * n + e]
GEN_HD void gen_load(const GenArgs& a, long long tile, int tid, int nthreads, float2* A) {
    const LineParams& p = a.p;
    const int n = a.n;
    for (int idx = tid; idx < a.lpc * n; idx += nthreads) {
        int l, e;
        gen_slot(a, idx, &l, &e);
        const long long line = tile * a.lpc + l;
        if (line >= a.n_lines) continue;
        const int b = int(line / p.lines_per_batch);
        const int li = int(line - (long long)b * p.lines_per_batch);
        const long long base = (long long)b * p.batch_stride + (long long)li * p.line_stride;
        float2 v;
        if (a.mode == GEN_INV) {
            int es = e + p.in_shift;                  // circular input shift (ifftshift) + conj: far-field adjoint
            if (es >= n) es -= n;
            v = g_conj(p.in[base + (long long)es * p.elem_stride]);
        } else {
            v = p.in[base + (long long)e * p.elem_stride];
        }
        if (a.pre_transmit) {
            const long long dbase = (long long)b * p.db_batch_stride + (long long)li * p.line_stride;
            v = g_cmul1p(v, g_transmission_m1(p.db[dbase + e], p.k_dz));
        }
        A[l * n + e] = v;
    }
}

// one Stockham stage of radix R with sub-transform size ns on ONE line: x -> y, work items j = lane, lane + nlanes, ...
GEN_HD void gen_stage(const float2* x, float2* y, int n, int R, int ns, const float2* tw, int lane, int nlanes) {
    const int m = n / R;
    const int tstep = n / (ns * R);                  // W_{ns R}^{q} = W_n^{q * tstep}
    const int rstep = n / R;                         // W_R^{q}     = W_n^{q * rstep}
    for (int j = lane; j < m; j += nlanes) {
        const int k = j % ns;
        float2 v[GEN_MAX_RADIX];
        for (int r = 0; r < R; ++r) {
            v[r] = x[j + r * m];
            if (r > 0 && k > 0) v[r] = g_cmul(v[r], tw[r * k * tstep]);
        }
        const int o = (j / ns) * ns * R + k;
        if (R == 2) {
            y[o] = make_float2(v[0].x + v[1].x, v[0].y + v[1].y);
            y[o + ns] = make_float2(v[0].x - v[1].x, v[0].y - v[1].y);
        } else if (R == 4) {
            const float2 s0 = make_float2(v[0].x + v[2].x, v[0].y + v[2].y), d0 = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
            const float2 s1 = make_float2(v[1].x + v[3].x, v[1].y + v[3].y), d1 = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
            y[o]          = make_float2(s0.x + s1.x, s0.y + s1.y);
            y[o + ns]     = make_float2(d0.x + d1.y, d0.y - d1.x);          // d0 - i d1
            y[o + 2 * ns] = make_float2(s0.x - s1.x, s0.y - s1.y);
            y[o + 3 * ns] = make_float2(d0.x - d1.y, d0.y + d1.x);          // d0 + i d1
        } else {
            for (int q = 0; q < R; ++q) {
                float2 acc = v[0];
                for (int r = 1; r < R; ++r) {
                    const float2 w = tw[((r * q) % R) * rstep];
                    const float2 t = g_cmul(v[r], w);
                    acc.x += t.x; acc.y += t.y;
                }
                y[o + q * ns] = acc;
            }
        }
    }
}

// between the two transforms of a convolution: x <- conj(x * h)   (h natural order, 1/N folded in)
GEN_HD void gen_mul_h(const GenArgs& a, long long line, float2* x, int lane, int nlanes) {
    const LineParams& p = a.p;
    if (a.mode == GEN_CONV) {
        for (int e = lane; e < a.n; e += nlanes) x[e] = g_conj(g_cmul(x[e], p.h[e]));
    } else {
        // general 2-D multiplier H[k_line][k_other] on a column pass: this line is column li
        const int b = int(line / p.lines_per_batch);
        const int li = int(line - (long long)b * p.lines_per_batch);
        for (int e = lane; e < a.n; e += nlanes) x[e] = g_conj(g_cmul(x[e], p.h[li + (long long)e * p.elem_stride]));
    }
}

// The transform(s) of one line: A_l holds the loaded line, B_l is its second buffer.  `ex(f)` runs f(lane, n_lanes) for
// every lane of the line's warp and synchronises them (device: each lane calls f for itself, then __syncwarp(); the CPU
// stepper calls f for lane 0..31 in turn).  Returns true when the result is in A_l, false when it is in B_l -- the same
// for every line, since it only depends on the stage count.
template <class Exec>
GEN_HD bool gen_transform_line(const GenArgs& a, long long line, float2* A_l, float2* B_l, Exec& ex) {
    float2* cur = A_l;
    float2* oth = B_l;
    bool in_a = true;
    const int n_pass = (a.mode == GEN_CONV || a.mode == GEN_CONV2D) ? 2 : 1;
    for (int pass = 0; pass < n_pass; ++pass) {
        int ns = 1;
        for (int s = 0; s < a.n_stages; ++s) {
            const int R = a.radix[s];
            ex([&](int lane, int nlanes) { gen_stage(cur, oth, a.n, R, ns, a.p.tw, lane, nlanes); });
            float2* t = cur; cur = oth; oth = t;
            in_a = !in_a;
            ns *= R;
        }
        if (pass == 0 && n_pass == 2) ex([&](int lane, int nlanes) { gen_mul_h(a, line, cur, lane, nlanes); });
    }
    return in_a;
}
GEN_HD bool gen_result_in_a(const GenArgs& a) {
    const int n_pass = (a.mode == GEN_CONV || a.mode == GEN_CONV2D) ? 2 : 1;
    return ((n_pass * a.n_stages) & 1) == 0;
}

// pass variant -> GenArgs switches; false for variants the mixed-radix pass does not implement
inline __host__ bool gen_set_variant(GenArgs& a, int variant) {
    a.col = 0; a.pre_transmit = 0; a.post_adj = 0;
    switch (variant) {
        case V_ROW_CONV_T:   a.mode = GEN_CONV; a.pre_transmit = 1; return true;
        case V_ROW_CONV:     a.mode = GEN_CONV; return true;
        case V_ROW_CONV_ADJ: a.mode = GEN_CONV; a.post_adj = 1; return true;
        case V_ROW_FWD:      a.mode = GEN_FWD; return true;
        case V_ROW_INV:      a.mode = GEN_INV; return true;
        case V_COL_CONV:     a.col = 1; a.mode = GEN_CONV; return true;
        case V_COL_FWD:      a.col = 1; a.mode = GEN_FWD; return true;
        case V_COL_INV:      a.col = 1; a.mode = GEN_INV; return true;
        case V_COL_CONV2D:   a.col = 1; a.mode = GEN_CONV2D; return true;
    }
    return false;
}
// lines per tile (= warps per CTA): both buffers of a tile fit in 64 KB of shared memory
inline __host__ int gen_lines_per_tile(int n) {
    const int lpc = 4096 / n;
    return lpc < 1 ? 1 : (lpc > 8 ? 8 : lpc);
}

// tile store: R[l * n + e] (the transform output; the CONJUGATE of the result except for GEN_FWD) -> global
GEN_HD void gen_store(const GenArgs& a, long long tile, int tid, int nthreads, const float2* Rb) {
    const LineParams& p = a.p;
    const int n = a.n;
    for (int idx = tid; idx < a.lpc * n; idx += nthreads) {
        int l, e;
        gen_slot(a, idx, &l, &e);
        const long long line = tile * a.lpc + l;
        if (line >= a.n_lines) continue;
        const int b = int(line / p.lines_per_batch);
        const int li = int(line - (long long)b * p.lines_per_batch);
        const long long base = (long long)b * p.batch_stride + (long long)li * p.line_stride;
        const float2 v = Rb[l * n + e];
        if (a.post_adj) {
            // G_u = conj(v) is the gradient w.r.t. u = psi t:  dL/ddelta = -k Im(conj(G_u) u), dL/dbeta = -k Re(conj(G_u) u),
            // G = conj(t) G_u   (same epilogue as line_kernel POST_ADJ)
            const long long dbase = (long long)b * p.db_batch_stride + (long long)li * p.line_stride;
            const float2 tr = g_transmission_m1(p.db[dbase + e], p.k_dz);      // tau = t - 1
            const float2 u = g_cmul1p(p.psi[base + e], tr);
            const float2 w = g_cmul(u, v);
            p.grad[dbase + e] = make_float2(-p.k_dz * w.y, -p.k_dz * w.x);
            p.out[base + e] = g_conj(g_cmul1p(v, tr));
        } else if (a.mode == GEN_FWD) {
            int es = e + p.out_shift;                 // circular output shift (fftshift): far field
            if (es >= n) es -= n;
            p.out[base + (long long)es * p.elem_stride] = v;
        } else {
            p.out[base + (long long)e * p.elem_stride] = g_conj(v);
        }
    }
}

}  // namespace bdof
