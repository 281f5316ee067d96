// Tiling-based multislice across GPUs (BASELINE config 5; the paper's scheme, whose enabling idea in this snapshot is the finite
// real-space kernel of cnn_propagator/propagation.py:41-47,91-99: finite support => finite halo).
//
// The global field [NY][NX] is split into a gy x gx grid of blocks, one per rank (one process per GPU).  A rank keeps its block
// with an APRON of `apron` pixels on every side in two buffers [by + 2 apron][bx + 2 apron] complex64 (ping-pong between slices),
// allocated here so that they can be exported by CUDA IPC.  Per slice a rank
//   1. cuts its local FFT windows (tiles with halo) out of the current buffer                      bdof_tiles_cut
//   2. steps them with the exact FFT propagator                                                    bdof_slice_step
//   3. pastes every tile's OWNED rectangle into the interior of the next buffer                    bdof_tiles_paste
//   4. pushes the border strips of that interior straight into the aprons of its 8 neighbours' next buffers -- plain stores to
//      peer-mapped memory over NVLink from one kernel -- then raises a flag word in every neighbour with a stream memory
//      operation, and makes its own stream wait for the 8 flags of its neighbours                  bdof_tiles_halo_exchange
// Nothing synchronises with the host: the whole slice loop is enqueued asynchronously.  Neighbours wrap around at the global
// border (the oracle's FFT propagator is periodic).  Why double buffering is enough: a rank pushes slice i's strips only after its
// own cut of slice i, which waited for the neighbours' flags of slice i-1, which they raised after THEIR cut of slice i-1 -- so
// nobody is still reading the apron that is being overwritten.
#include "../../include/bdof.h"
#include "common.h"

#include <cstdint>
#include <cstring>
#include <vector>

typedef CUresult (*TileStreamValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static TileStreamValue32Fn t_wait32 = nullptr, t_write32 = nullptr;

static int tiles_load_memops() {
    if (t_wait32 && t_write32) return 0;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_TRY(cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q));
    if (!f || q != cudaDriverEntryPointSuccess) return bdof_fail(BDOF_E_UNSUPPORTED, "cuStreamWaitValue32 is not available in this driver");
    t_wait32 = reinterpret_cast<TileStreamValue32Fn>(f);
    f = nullptr;
    CUDA_TRY(cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &q));
    if (!f || q != cudaDriverEntryPointSuccess) return bdof_fail(BDOF_E_UNSUPPORTED, "cuStreamWriteValue32 is not available in this driver");
    t_write32 = reinterpret_cast<TileStreamValue32Fn>(f);
    return 0;
}

struct bdof_tiles {
    int rank = 0, gy = 1, gx = 1, by = 0, bx = 0, apron = 0;
    int world() const { return gy * gx; }
    long long pitch() const { return (long long)bx + 2 * apron; }
    size_t buf_bytes() const { return size_t(by + 2 * apron) * size_t(pitch()) * sizeof(float2); }
    size_t off_flags() const { return 2 * ((buf_bytes() + 255) / 256 * 256); }
    char* base = nullptr;                 // buffer 0 | buffer 1 | flags[8]
    std::vector<char*> peer_base;
    std::vector<bool> opened;
    uint32_t seq = 0;
};

// one strip: src rectangle of my buffer -> dst rectangle of a neighbour's buffer (same pitch on every rank)
struct HaloStrip { float2* dst; int sy, sx, dy, dx, h, w; };
struct HaloPush { HaloStrip s[8]; long long first[9]; };

__global__ void __launch_bounds__(256) k_halo_push(const float2* __restrict__ src, long long pitch, const HaloPush hp) {
    const long long total = hp.first[8];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int k = 0;
#pragma unroll
        for (int j = 1; j < 8; ++j) k += (i >= hp.first[j]) ? 1 : 0;
        const HaloStrip st = hp.s[k];
        const long long o = i - hp.first[k];
        const int y = int(o / st.w), x = int(o - (long long)y * st.w);
        st.dst[(long long)(st.dy + y) * pitch + st.dx + x] = src[(long long)(st.sy + y) * pitch + st.sx + x];
    }
}

// tiles[t][y][x] = buf[apron + oy_t + y][apron + ox_t + x]   (origins relative to the block interior)
__global__ void __launch_bounds__(256) k_tiles_cut(const float2* __restrict__ buf, long long pitch, int apron, const int* __restrict__ origin,
                                                    int ly, int lx, float2* __restrict__ tiles) {
    const int t = blockIdx.z, y = blockIdx.y;
    const int oy = origin[2 * t], ox = origin[2 * t + 1];
    const float2* row = buf + (long long)(apron + oy + y) * pitch + apron + ox;
    float2* out = tiles + ((long long)t * ly + y) * lx;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < lx; x += gridDim.x * blockDim.x) out[x] = row[x];
}
// buf interior[own_y + y][own_x + x] = tiles[t][own_y - oy_t + y][own_x - ox_t + x] for the owned rectangle of every tile
__global__ void __launch_bounds__(256) k_tiles_paste(float2* __restrict__ buf, long long pitch, int apron, const float2* __restrict__ tiles,
                                                      const int* __restrict__ origin, const int* __restrict__ own, int ly, int lx) {
    const int t = blockIdx.z;
    const int oy = origin[2 * t], ox = origin[2 * t + 1];
    const int y0 = own[4 * t], x0 = own[4 * t + 1], h = own[4 * t + 2], w = own[4 * t + 3];
    const int y = blockIdx.y;
    if (y >= h) return;
    const float2* row = tiles + ((long long)t * ly + (y0 - oy + y)) * lx + (x0 - ox);
    float2* out = buf + (long long)(apron + y0 + y) * pitch + apron + x0;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < w; x += gridDim.x * blockDim.x) out[x] = row[x];
}

extern "C" int bdof_tiles_create(bdof_tiles** out, int rank, int gy, int gx, int by, int bx, int apron) {
    if (!out || gy < 1 || gx < 1 || rank < 0 || rank >= gy * gx || by < 1 || bx < 1 || apron < 0 || apron > by || apron > bx)
        return bdof_fail(BDOF_E_BADARG, "bad tile grid (the apron may not exceed the block: a strip comes from ONE neighbour)");
    BDOF_TRY(tiles_load_memops());
    bdof_tiles* c = new bdof_tiles();
    c->rank = rank; c->gy = gy; c->gx = gx; c->by = by; c->bx = bx; c->apron = apron;
    const size_t total = c->off_flags() + 256;
    cudaError_t e = cudaMalloc((void**)&c->base, total);
    if (e != cudaSuccess) { delete c; cudaGetLastError(); return bdof_fail(int(e), "cudaMalloc of the tile buffers (%zu bytes): %s", total, cudaGetErrorString(e)); }
    e = cudaMemset(c->base, 0, total);
    if (e != cudaSuccess) { cudaFree(c->base); delete c; cudaGetLastError(); return bdof_fail(int(e), "cudaMemset: %s", cudaGetErrorString(e)); }
    c->peer_base.assign(c->world(), nullptr);
    c->opened.assign(c->world(), false);
    c->peer_base[rank] = c->base;
    *out = c;
    return 0;
}

extern "C" void bdof_tiles_destroy(bdof_tiles* c) {
    if (!c) return;
    cudaDeviceSynchronize();
    for (int p = 0; p < c->world(); ++p)
        if (p != c->rank && c->opened[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
    cudaFree(c->base);
    cudaGetLastError();
    delete c;
}

extern "C" int bdof_tiles_handle_bytes(void) { return int(sizeof(cudaIpcMemHandle_t)); }

extern "C" int bdof_tiles_export(bdof_tiles* c, void* h_handle_out) {
    if (!c || !h_handle_out) return bdof_fail(BDOF_E_BADARG, "null");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, c->base));
    memcpy(h_handle_out, &h, sizeof(h));
    return 0;
}

extern "C" int bdof_tiles_connect(bdof_tiles* c, const void* h_all_handles) {
    if (!c || !h_all_handles) return bdof_fail(BDOF_E_BADARG, "null");
    const char* hs = static_cast<const char*>(h_all_handles);
    for (int p = 0; p < c->world(); ++p) {
        if (p == c->rank || c->opened[p]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs + size_t(p) * sizeof(h), sizeof(h));
        void* ptr = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_base[p] = static_cast<char*>(ptr);
        c->opened[p] = true;
    }
    return 0;
}

extern "C" int bdof_tiles_block_ptr(bdof_tiles* c, int which, void** d_out) {
    if (!c || !d_out || which < 0 || which > 1) return bdof_fail(BDOF_E_BADARG, "bad argument");
    *d_out = c->base + size_t(which) * (c->off_flags() / 2);
    return 0;
}

extern "C" int bdof_tiles_cut(bdof_tiles* c, int which, const int* d_origin_yx, int n_tiles, int ly, int lx, float* d_tiles, void* st) {
    if (!c || !d_origin_yx || !d_tiles || which < 0 || which > 1 || n_tiles < 1 || ly < 1 || lx < 1) return bdof_fail(BDOF_E_BADARG, "bad argument");
    if (ly > 65535 || n_tiles > 65535) return bdof_fail(BDOF_E_UNSUPPORTED, "tile too tall / too many tiles");
    const float2* buf = reinterpret_cast<const float2*>(c->base + size_t(which) * (c->off_flags() / 2));
    dim3 grid((lx + 1023) / 1024, ly, n_tiles);
    k_tiles_cut<<<grid, 256, 0, (cudaStream_t)st>>>(buf, c->pitch(), c->apron, d_origin_yx, ly, lx, reinterpret_cast<float2*>(d_tiles));
    return bdof_launch_check("k_tiles_cut");
}

extern "C" int bdof_tiles_paste(bdof_tiles* c, int which, const float* d_tiles, const int* d_origin_yx, const int* d_own_yxhw, int n_tiles,
                                int ly, int lx, void* st) {
    if (!c || !d_origin_yx || !d_own_yxhw || !d_tiles || which < 0 || which > 1 || n_tiles < 1) return bdof_fail(BDOF_E_BADARG, "bad argument");
    if (ly > 65535 || n_tiles > 65535) return bdof_fail(BDOF_E_UNSUPPORTED, "tile too tall / too many tiles");
    float2* buf = reinterpret_cast<float2*>(c->base + size_t(which) * (c->off_flags() / 2));
    dim3 grid((lx + 1023) / 1024, ly, n_tiles);
    k_tiles_paste<<<grid, 256, 0, (cudaStream_t)st>>>(buf, c->pitch(), c->apron, reinterpret_cast<const float2*>(d_tiles), d_origin_yx, d_own_yxhw, ly, lx);
    return bdof_launch_check("k_tiles_paste");
}

extern "C" int bdof_tiles_halo_exchange(bdof_tiles* c, int which, void* st_) {
    if (!c || which < 0 || which > 1) return bdof_fail(BDOF_E_BADARG, "bad argument");
    cudaStream_t st = (cudaStream_t)st_;
    const int a = c->apron;
    if (a == 0) return 0;
    for (int p = 0; p < c->world(); ++p)
        if (!c->peer_base[p]) return bdof_fail(BDOF_E_STATE, "bdof_tiles_connect has not been called");
    const int ry = c->rank / c->gx, rx = c->rank % c->gx;
    const int by = c->by, bx = c->bx;
    const size_t boff = size_t(which) * (c->off_flags() / 2);
    HaloPush hp;
    long long n = 0;
    int k = 0;
    int nb_rank[8], nb_slot[8];
    // direction (dy, dx): my strip on that side goes to the neighbour there, into its apron on the OPPOSITE side
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            if (dy == 0 && dx == 0) continue;
            const int ny_ = (ry + dy + c->gy) % c->gy, nx_ = (rx + dx + c->gx) % c->gx;
            const int nb = ny_ * c->gx + nx_;
            HaloStrip s;
            s.h = (dy == 0) ? by : a;
            s.w = (dx == 0) ? bx : a;
            // source: interior edge facing the neighbour (buffer coordinates: interior starts at (a, a))
            s.sy = (dy < 0) ? a : (dy > 0 ? a + by - a : a);
            s.sx = (dx < 0) ? a : (dx > 0 ? a + bx - a : a);
            // destination in the neighbour's buffer: its apron on the side facing me
            s.dy = (dy < 0) ? a + by : (dy > 0 ? 0 : a);
            s.dx = (dx < 0) ? a + bx : (dx > 0 ? 0 : a);
            s.dst = reinterpret_cast<float2*>(c->peer_base[nb] + boff);
            hp.s[k] = s;
            hp.first[k] = n;
            n += (long long)s.h * s.w;
            nb_rank[k] = nb;
            nb_slot[k] = (1 - dy) * 3 + (1 - dx);       // the slot of the opposite direction (-dy, -dx) in the neighbour's flag array
            ++k;
        }
    hp.first[8] = n;
    const float2* src = reinterpret_cast<const float2*>(c->base + boff);
    const unsigned grid = unsigned((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);
    k_halo_push<<<grid, 256, 0, st>>>(src, c->pitch(), hp);
    BDOF_TRY(bdof_launch_check("k_halo_push"));
    const uint32_t seq = ++c->seq;
    // flags: slot index (dy + 1) * 3 + (dx + 1) of the direction the data CAME from, 9 words per buffer set (slot 4 unused)
    for (int j = 0; j < 8; ++j) {
        CUdeviceptr f = (CUdeviceptr)(c->peer_base[nb_rank[j]] + c->off_flags() + size_t(nb_slot[j]) * sizeof(uint32_t));
        if (t_write32((CUstream)st, f, seq, 0) != CUDA_SUCCESS) return bdof_fail(BDOF_E_STATE, "cuStreamWriteValue32 failed");
    }
    for (int slot = 0; slot < 9; ++slot) {
        if (slot == 4) continue;
        CUdeviceptr f = (CUdeviceptr)(c->base + c->off_flags() + size_t(slot) * sizeof(uint32_t));
        if (t_wait32((CUstream)st, f, seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) return bdof_fail(BDOF_E_STATE, "cuStreamWaitValue32 failed");
    }
    return 0;
}
