// CPU stepper for the mixed-radix line pass (csrc/genericfft.cuh): runs the SAME __host__ __device__ index arithmetic
// as generic_line_kernel thread by thread on the host, so that tests/test_generic_host.py can check every pass variant
// against numpy.fft without a GPU.  Test infrastructure only -- built into tests/_build/, never into libbdof.so.
#include "../beyond_dof_b200/csrc/genericfft.cuh"
#include <vector>
#include <cmath>

using namespace bdof;

struct HostExec {
    template <class F> void operator()(F f) { for (int lane = 0; lane < 32; ++lane) f(lane, 32); }
};

extern "C" int genfft_host_factorize(int n, int* radix) { return gen_factorize(n, radix); }

extern "C" int genfft_host_run(int n, int variant, long long n_lines, int lines_per_batch, long long batch_stride, int elem_stride,
                               int line_stride, int in_shift, int out_shift, float k_dz, const float* in, float* out, const float* h,
                               const float* db, float* grad, const float* psi) {
    GenArgs a{};
    a.n = n; a.n_lines = n_lines;
    a.n_stages = gen_factorize(n, a.radix);
    if (a.n_stages == 0) return -2;
    if (!gen_set_variant(a, variant)) return -1;
    a.lpc = gen_lines_per_tile(n);
    std::vector<float2> tw(n);
    for (int k = 0; k < n; ++k) {
        const double ang = -2.0 * M_PI * double(k) / double(n);
        tw[k] = make_float2(float(cos(ang)), float(sin(ang)));
    }
    LineParams& p = a.p;
    p.in = reinterpret_cast<const float2*>(in); p.out = reinterpret_cast<float2*>(out); p.h = reinterpret_cast<const float2*>(h);
    p.tw = tw.data(); p.db = reinterpret_cast<const float2*>(db); p.grad = reinterpret_cast<float2*>(grad);
    p.psi = reinterpret_cast<const float2*>(psi);
    p.batch_stride = batch_stride; p.db_batch_stride = batch_stride; p.lines_per_batch = lines_per_batch;
    p.elem_stride = elem_stride; p.line_stride = line_stride; p.in_shift = in_shift; p.out_shift = out_shift; p.k_dz = k_dz;
    const int nthreads = 32 * a.lpc;
    std::vector<float2> A(size_t(a.lpc) * n), B(size_t(a.lpc) * n);
    const long long n_tiles = (n_lines + a.lpc - 1) / a.lpc;
    HostExec ex;
    const bool in_a = gen_result_in_a(a);
    for (long long tile = 0; tile < n_tiles; ++tile) {
        for (int tid = 0; tid < nthreads; ++tid) gen_load(a, tile, tid, nthreads, A.data());
        for (int w = 0; w < a.lpc; ++w) {
            const long long line = tile * a.lpc + w;
            if (line < n_lines) {
                const bool r = gen_transform_line(a, line, A.data() + size_t(w) * n, B.data() + size_t(w) * n, ex);
                if (r != in_a) return -3;
            }
        }
        for (int tid = 0; tid < nthreads; ++tid) gen_store(a, tile, tid, nthreads, in_a ? A.data() : B.data());
    }
    return 0;
}
