// Instantiation unit and launcher of the cluster-resident kernels (clusterfft.cuh): 256 x 256 fields on clusters of 8 CTAs,
// 128 x 128 fields on clusters of 4 (32 lines per CTA either way).
#include "../../include/bdof.h"
#include "common.h"
#include "clusterfft.cuh"
#include <cstdlib>

using namespace bdof;

template <class Cfg, int C>
static int launch_cluster(int adj, const ResidentParams& p, cudaStream_t st) {
    using SM = ClusterSmem<Cfg, C>;
    constexpr int THREADS = (Cfg::N / C) * Cfg::T;
    if (p.win != nullptr) return bdof_fail(BDOF_E_UNSUPPORTED, "window mode is a feature of the 64 x 64 resident kernels");
    auto kf = cluster_forward_kernel<Cfg, C>;
    auto ka = cluster_adjoint_kernel<Cfg, C>;
    static int max_clusters[2] = {0, 0};
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SM::BYTES; cfg.stream = st;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (max_clusters[adj] == 0) {
        if (adj) CUDA_TRY(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SM::BYTES)));
        else     CUDA_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SM::BYTES)));
        int n = 0;
        cfg.gridDim = dim3(C);
        if (adj) CUDA_TRY(cudaOccupancyMaxActiveClusters(&n, ka, &cfg));
        else     CUDA_TRY(cudaOccupancyMaxActiveClusters(&n, kf, &cfg));
        if (n < 1) return bdof_fail(BDOF_E_UNSUPPORTED, "no cluster of %d CTAs with %zu bytes of shared memory fits on this GPU", C, SM::BYTES);
        max_clusters[adj] = n;
    }
    const int n_cl = p.batch < max_clusters[adj] ? p.batch : max_clusters[adj];
    cfg.gridDim = dim3(unsigned(n_cl * C));
    if (adj) CUDA_TRY(cudaLaunchKernelEx(&cfg, ka, p));
    else     CUDA_TRY(cudaLaunchKernelEx(&cfg, kf, p));
    return bdof_launch_check(adj ? "cluster_adjoint_kernel" : "cluster_forward_kernel");
}

int bdof_cluster_supported(int n) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("BDOF_CLUSTER"); enabled = (e && e[0] == '0') ? 0 : 1; }
    return (enabled && (n == 256 || n == 128)) ? 1 : 0;
}

int bdof_launch_cluster(int n, int adj, const ResidentParams& p, cudaStream_t st) {
    switch (n) {
        case 256: return launch_cluster<LineCfg<256, 16, 16, 16, 1>, 8>(adj, p, st);
        case 128: return launch_cluster<LineCfg<128, 8, 16, 8, 1>, 4>(adj, p, st);
    }
    return bdof_fail(BDOF_E_UNSUPPORTED, "no cluster-resident kernel for %d x %d fields", n, n);
}
