// Instantiation unit and launcher of the resident small-field kernels (residentfft.cuh).
#include "../../include/bdof.h"
#include "common.h"
#include "residentfft.cuh"
#include <cstdlib>

using namespace bdof;

template <int N> struct ResCfg;
template <> struct ResCfg<64>  { using C = LineCfg<64, 8, 8, 8, 1>; };

template <class Cfg>
static int launch_resident(int adj, const ResidentParams& p, cudaStream_t st) {
    using SM = ResidentSmem<Cfg>;
    auto kf = resident_forward_kernel<Cfg, 1>;
    auto kf2 = resident_forward_kernel<Cfg, 2>;
    auto ka = resident_adjoint_kernel<Cfg, 1>;
    auto ka2 = resident_adjoint_kernel<Cfg, 2>;
    static int slots[2] = {0, 0};
    static int slots_f2 = 0, slots_a2 = 0, n_sm_cached = 0;
    if (slots_f2 == 0) {
        int dev = 0, occ = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&n_sm_cached, cudaDevAttrMultiProcessorCount, dev));
        CUDA_TRY(cudaFuncSetAttribute(kf2, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SM::BYTES)));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kf2, Cfg::N * Cfg::T, SM::BYTES));
        slots_f2 = (occ >= 2 ? occ : -1) * (n_sm_cached > 0 ? n_sm_cached : 148);
        CUDA_TRY(cudaFuncSetAttribute(ka2, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SM::BYTES)));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ka2, Cfg::N * Cfg::T, SM::BYTES));
        slots_a2 = (occ >= 2 ? occ : -1) * (n_sm_cached > 0 ? n_sm_cached : 148);
    }
    if (slots[adj] == 0) {
        int dev = 0, n_sm = 0, occ = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        if (adj) {
            CUDA_TRY(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SM::BYTES)));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ka, Cfg::N * Cfg::T, SM::BYTES));
        } else {
            CUDA_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SM::BYTES)));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kf, Cfg::N * Cfg::T, SM::BYTES));
        }
        if (occ < 1) return bdof_fail(BDOF_E_UNSUPPORTED, "resident kernel does not fit on an SM (%zu bytes smem)", SM::BYTES);
        slots[adj] = occ * (n_sm > 0 ? n_sm : 148);
    }
    unsigned grid = unsigned(p.batch < slots[adj] ? p.batch : slots[adj]);
    static int two = -1;
    // measured (1024 fields of 64^2 x 128, B200): two register-capped CTAs per SM are 4-6 % SLOWER than one (forward 4.07 vs
    // 3.92 ms, adjoint 3.77 vs 3.55 ms): the kernels are bound by issue slots and shared-memory wavefronts, not by latency.
    // Kept as a switch (BDOF_RESIDENT_2CTA=1).
    if (two < 0) { const char* e = getenv("BDOF_RESIDENT_2CTA"); two = (e && e[0] == '1') ? 1 : 0; }
    if (adj && two && slots_a2 > 0 && p.batch > slots[1]) {
        grid = unsigned(p.batch < slots_a2 ? p.batch : slots_a2);
        ka2<<<grid, Cfg::N * Cfg::T, SM::BYTES, st>>>(p);
    } else if (adj) {
        ka<<<grid, Cfg::N * Cfg::T, SM::BYTES, st>>>(p);
    } else if (two && slots_f2 > 0 && p.batch > slots[0]) {
        // more fields than SMs: the register-capped variant, two CTAs per SM
        grid = unsigned(p.batch < slots_f2 ? p.batch : slots_f2);
        kf2<<<grid, Cfg::N * Cfg::T, SM::BYTES, st>>>(p);
    } else {
        kf<<<grid, Cfg::N * Cfg::T, SM::BYTES, st>>>(p);
    }
    return bdof_launch_check(adj ? "resident_adjoint_kernel" : "resident_forward_kernel");
}

int bdof_resident_supported(int n) { return (n == 64 || bdof_cluster_supported(n)) ? 1 : 0; }

int bdof_launch_resident(int n, int adj, const ResidentParams& p, cudaStream_t st) {
    switch (n) {
        case 64: return launch_resident<ResCfg<64>::C>(adj, p, st);
    }
    if (bdof_cluster_supported(n)) return bdof_launch_cluster(n, adj, p, st);
    return bdof_fail(BDOF_E_UNSUPPORTED, "no resident kernel for %d x %d fields", n, n);
}
