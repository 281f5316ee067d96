// Data-parallel gradient exchange over NVLink peer memory with the COPY ENGINES (no communication kernels on the SMs).
//
// Why not NCCL here: the sweep kernels are persistent and own every SM (227 KB of shared memory, 255 registers x 256
// threads), so NCCL's all-reduce CTAs displace sweep CTAs for as long as a bucket is in flight (24.5 ms of
// compute + 14 ms of all-reduce overlapped to only 34.3 ms on two B200).  The exchange below moves every byte with
// cudaMemcpyAsync between peer-mapped buffers (CUDA IPC), signals with 4-byte copies into the peer's flag array, waits
// with stream memory operations (cuStreamWaitValue32) and touches the SMs only for the short sum of the N partial shards.
//
// One context per rank (one process per GPU).  Symmetric buffers, allocated here so that they can be exported:
//   grad     [grad_bytes]                the object gradient the adjoint writes (wrapped as a tensor by the host side)
//   staging  [(N-1) x grad_bytes / N]    slot s receives the partial shard of the s-th OTHER rank
//   flags    [2 x K x N] uint32          A[j][p]: rank p's partial of bucket j has landed here (value = step number)
//                                        B[j][p]: rank p's reduced shard of bucket j has landed in my grad
// Bucket j (a contiguous z range of the slice-major gradient, final as soon as the adjoint sweep has passed it) is cut
// in N equal shards; rank r owns shard r:
//   1. push   my partial of shard p -> staging of rank p, then flag A[j][me] on rank p          (copy engine, per-peer stream)
//   2. reduce once A[j][*] have arrived: grad[shard me] = (mine + sum of the staged partials) / N  (one small kernel)
//   3. gather my reduced shard -> grad of every peer, then flag B[j][me] there                   (copy engine)
//   finish:   wait for B[*][*] and for my own copies.
// The reduced shard is computed once and broadcast, so every rank ends with bit-identical gradients.
#include "../../include/bdof.h"
#include "common.h"

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef CUresult (*StreamValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static StreamValue32Fn g_wait32 = nullptr, g_write32 = nullptr;

static int load_stream_memops() {
    if (g_wait32 && g_write32) return 0;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_TRY(cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q));
    if (!f || q != cudaDriverEntryPointSuccess) return bdof_fail(BDOF_E_UNSUPPORTED, "cuStreamWaitValue32 is not available in this driver");
    g_wait32 = reinterpret_cast<StreamValue32Fn>(f);
    f = nullptr;
    CUDA_TRY(cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &q));
    if (!f || q != cudaDriverEntryPointSuccess) return bdof_fail(BDOF_E_UNSUPPORTED, "cuStreamWriteValue32 is not available in this driver");
    g_write32 = reinterpret_cast<StreamValue32Fn>(f);
    return 0;
}
#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        CUresult _r = (expr);                                                                     \
        if (_r != CUDA_SUCCESS) return bdof_fail(BDOF_E_STATE, "%s failed (CUresult %d)", #expr, int(_r)); \
    } while (0)

struct bdof_dp {
    int rank = 0, world = 1, n_buckets = 0;
    size_t grad_bytes = 0, staging_bytes = 0, flags_bytes = 0, total_bytes = 0;
    char* base = nullptr;                    // my allocation: grad | staging | flags | scratch
    std::vector<char*> peer_base;            // [world], peer_base[rank] == base
    std::vector<bool> opened;
    std::vector<cudaStream_t> send;          // [world], pushes of partial shards, one stream per peer (unused at [rank])
    std::vector<cudaStream_t> gath;          // [world], gathers of reduced shards (own streams: a gather waits for the peers'
                                             // partials, the next bucket's push must not queue behind it)
    std::vector<cudaEvent_t> send_done;      // [world]
    std::vector<cudaEvent_t> gath_done;      // [world]
    // One copy engine moves ~550 GB/s over NVLink 5 (measured, 2 x B200); a shard is therefore cut in `split` pieces that
    // travel on their own streams (sub[p][k], k >= 1; piece 0 stays on send[p] / gath[p], which also carries the flag).
    int split = 1;
    bool flag_memop = true;                  // flags written straight into the peer's flag word by a stream memory operation
                                             // instead of a 4-byte copy (which queues behind the data copies on the copy engines)
    std::vector<std::vector<cudaStream_t>> sub_send, sub_gath;     // [world][split - 1]
    std::vector<std::vector<cudaEvent_t>> sub_send_ev, sub_gath_ev;
    cudaStream_t gseq = nullptr;             // gather-only mode: ONE stream, peers served one after the other (concurrent copy-engine
    cudaEvent_t gseq_done = nullptr;         // transfers to several peers do not add up: 340 GB/s aggregate at 4 GPUs vs 550 GB/s for one)
    bool gather_only = false;                // no staging area: the reduce-scatter half is done elsewhere (NCCL)
    cudaStream_t red = nullptr;
    std::vector<cudaEvent_t> red_done;       // [n_buckets]
    cudaEvent_t fin = nullptr;
    uint32_t epoch = 0;
    int buckets_issued = 0;
    size_t off_staging() const { return grad_bytes; }
    size_t off_flags() const { return grad_bytes + staging_bytes; }
    size_t off_scratch() const { return grad_bytes + staging_bytes + flags_bytes; }
};
static constexpr int DP_SCRATCH_RING = 4;

// own[i] = (own[i] + sum_s staged[s][i]) * scale, 16 bytes per thread and step
__global__ void __launch_bounds__(256) k_dp_reduce(float4* __restrict__ own, const float4* __restrict__ staged, long long n4,
                                                   long long slot_stride4, int n_slots, float scale) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 a = own[i];
        for (int s = 0; s < n_slots; ++s) {
            const float4 b = __ldcs(staged + s * slot_stride4 + i);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
        own[i] = a;
    }
}

extern "C" int bdof_dp_create(bdof_dp** out, int rank, int world, size_t grad_bytes, int n_buckets, int gather_only) {
    if (!out || world < 1 || rank < 0 || rank >= world || grad_bytes == 0 || n_buckets < 1)
        return bdof_fail(BDOF_E_BADARG, "bad exchange shape");
    if (grad_bytes % (size_t(world) * 16) != 0)
        return bdof_fail(BDOF_E_UNSUPPORTED, "gradient of %zu bytes does not split into %d shards of whole 16-byte units", grad_bytes, world);
    BDOF_TRY(load_stream_memops());
    bdof_dp* c = new bdof_dp();
    c->rank = rank; c->world = world; c->n_buckets = n_buckets;
    c->grad_bytes = grad_bytes;
    c->gather_only = gather_only != 0;
    c->staging_bytes = c->gather_only ? 0 : (grad_bytes / world) * size_t(world - 1);
    c->flags_bytes = ((size_t(2) * n_buckets * world * sizeof(uint32_t)) + 255) / 256 * 256;
    const size_t scratch_bytes = 256 * size_t(world);
    c->total_bytes = c->grad_bytes + c->staging_bytes + c->flags_bytes + scratch_bytes;
    cudaError_t e = cudaMalloc((void**)&c->base, c->total_bytes);
    if (e != cudaSuccess) { delete c; cudaGetLastError(); return bdof_fail(int(e), "cudaMalloc of the exchange buffers (%zu bytes): %s", c->total_bytes, cudaGetErrorString(e)); }
    e = cudaMemset(c->base + c->off_flags(), 0, c->flags_bytes + scratch_bytes);
    if (e != cudaSuccess) { cudaFree(c->base); delete c; cudaGetLastError(); return bdof_fail(int(e), "cudaMemset: %s", cudaGetErrorString(e)); }
    c->peer_base.assign(world, nullptr);
    c->opened.assign(world, false);
    c->peer_base[rank] = c->base;
    c->send.assign(world, nullptr);
    c->gath.assign(world, nullptr);
    c->send_done.assign(world, nullptr);
    c->gath_done.assign(world, nullptr);
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    c->split = 1;                            // measured: more pieces do not raise the ~550 GB/s of one peer-to-peer direction
    if (const char* ev = getenv("BDOF_DP_SPLIT")) c->split = atoi(ev);
    if (const char* ev = getenv("BDOF_DP_FLAG")) c->flag_memop = atoi(ev) != 0;
    c->split = c->split < 1 ? 1 : (c->split > 8 ? 8 : c->split);
    c->sub_send.resize(world); c->sub_gath.resize(world); c->sub_send_ev.resize(world); c->sub_gath_ev.resize(world);
    for (int p = 0; p < world; ++p) {
        if (p == rank) continue;
        for (int k = 1; k < c->split; ++k) {
            cudaStream_t s1 = nullptr, s2 = nullptr;
            cudaEvent_t e1 = nullptr, e2 = nullptr;
            cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
            cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
            cudaEventCreateWithFlags(&e1, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&e2, cudaEventDisableTiming);
            c->sub_send[p].push_back(s1); c->sub_gath[p].push_back(s2);
            c->sub_send_ev[p].push_back(e1); c->sub_gath_ev[p].push_back(e2);
        }
        cudaStreamCreateWithFlags(&c->send[p], cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&c->gath[p], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&c->send_done[p], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->gath_done[p], cudaEventDisableTiming);
    }
    // the reduction is short and on the critical path of the gather: let its CTAs in ahead of the next sweep kernel
    cudaStreamCreateWithPriority(&c->red, cudaStreamNonBlocking, hi);
    cudaStreamCreateWithFlags(&c->gseq, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&c->gseq_done, cudaEventDisableTiming);
    c->red_done.resize(n_buckets);
    for (auto& ev : c->red_done) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->fin, cudaEventDisableTiming);
    e = cudaGetLastError();
    if (e != cudaSuccess) { bdof_dp_destroy(c); return bdof_fail(int(e), "stream/event creation: %s", cudaGetErrorString(e)); }
    *out = c;
    return 0;
}

extern "C" void bdof_dp_destroy(bdof_dp* c) {
    if (!c) return;
    cudaDeviceSynchronize();
    for (int p = 0; p < c->world; ++p) {
        if (p != c->rank && c->opened[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
        for (auto st : c->sub_send[p]) cudaStreamDestroy(st);
        for (auto st : c->sub_gath[p]) cudaStreamDestroy(st);
        for (auto ev : c->sub_send_ev[p]) cudaEventDestroy(ev);
        for (auto ev : c->sub_gath_ev[p]) cudaEventDestroy(ev);
        if (c->send[p]) cudaStreamDestroy(c->send[p]);
        if (c->gath[p]) cudaStreamDestroy(c->gath[p]);
        if (c->send_done[p]) cudaEventDestroy(c->send_done[p]);
        if (c->gath_done[p]) cudaEventDestroy(c->gath_done[p]);
    }
    if (c->red) cudaStreamDestroy(c->red);
    if (c->gseq) cudaStreamDestroy(c->gseq);
    if (c->gseq_done) cudaEventDestroy(c->gseq_done);
    for (auto ev : c->red_done) cudaEventDestroy(ev);
    if (c->fin) cudaEventDestroy(c->fin);
    cudaFree(c->base);
    cudaGetLastError();
    delete c;
}

extern "C" int bdof_dp_handle_bytes(void) { return int(sizeof(cudaIpcMemHandle_t)); }

extern "C" int bdof_dp_export(bdof_dp* c, void* h_handle_out) {
    if (!c || !h_handle_out) return bdof_fail(BDOF_E_BADARG, "null");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, c->base));
    memcpy(h_handle_out, &h, sizeof(h));
    return 0;
}

extern "C" int bdof_dp_connect(bdof_dp* c, const void* h_all_handles) {
    if (!c || !h_all_handles) return bdof_fail(BDOF_E_BADARG, "null");
    const char* hs = static_cast<const char*>(h_all_handles);
    for (int p = 0; p < c->world; ++p) {
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

extern "C" int bdof_dp_grad_ptr(bdof_dp* c, void** d_grad_out) {
    if (!c || !d_grad_out) return bdof_fail(BDOF_E_BADARG, "null");
    *d_grad_out = c->base;
    return 0;
}

// dst <- src (n bytes, multiple of 16) in `split` pieces: piece 0 on `main` (after `after`), the others on their own streams;
// on return `main` is ordered after every piece
static int split_copy(bdof_dp* c, char* dst, const char* src, size_t n, cudaStream_t main, std::vector<cudaStream_t>& subs,
                      std::vector<cudaEvent_t>& evs, cudaEvent_t after) {
    const int K = (n >= (size_t(1) << 22)) ? c->split : 1;            // small shards: one piece
    const size_t piece = ((n / K) + 255) / 256 * 256;
    CUDA_TRY(cudaStreamWaitEvent(main, after, 0));
    for (int k = 1; k < K; ++k) {
        const size_t o = size_t(k) * piece;
        if (o >= n) break;
        const size_t len = (o + piece <= n) ? piece : n - o;
        CUDA_TRY(cudaStreamWaitEvent(subs[k - 1], after, 0));
        CUDA_TRY(cudaMemcpyAsync(dst + o, src + o, len, cudaMemcpyDeviceToDevice, subs[k - 1]));
        CUDA_TRY(cudaEventRecord(evs[k - 1], subs[k - 1]));
    }
    CUDA_TRY(cudaMemcpyAsync(dst, src, piece < n ? piece : n, cudaMemcpyDeviceToDevice, main));
    for (int k = 1; k < K; ++k) {
        if (size_t(k) * piece >= n) break;
        CUDA_TRY(cudaStreamWaitEvent(main, evs[k - 1], 0));
    }
    return 0;
}

// Exchange of one bucket = bytes [offset, offset + n_bytes) of the gradient, valid on this rank once `ready_event` (recorded
// on the stream that runs the adjoint) has fired.  Buckets of one step must be issued in the same order on every rank.
extern "C" int bdof_dp_bucket(bdof_dp* c, size_t offset, size_t n_bytes, void* ready_event) {
    if (!c || !ready_event) return bdof_fail(BDOF_E_BADARG, "null");
    const int N = c->world, me = c->rank;
    if (c->gather_only) return bdof_fail(BDOF_E_STATE, "the context was created gather-only: use bdof_dp_gather");
    if (c->buckets_issued >= c->n_buckets) return bdof_fail(BDOF_E_STATE, "more buckets than the context was created for");
    if (offset + n_bytes > c->grad_bytes || n_bytes % (size_t(N) * 16) != 0 || offset % (size_t(N) * 16) != 0)
        return bdof_fail(BDOF_E_BADARG, "bucket [%zu, +%zu) does not split into %d aligned shards", offset, n_bytes, N);
    for (int p = 0; p < N; ++p)
        if (p != me && !c->opened[p]) return bdof_fail(BDOF_E_STATE, "bdof_dp_connect has not been called");
    const int j = c->buckets_issued++;
    if (j == 0) ++c->epoch;
    const uint32_t e = c->epoch;
    const size_t shard = n_bytes / N, slot_stride = c->grad_bytes / N;
    cudaEvent_t ready = reinterpret_cast<cudaEvent_t>(ready_event);
    uint32_t* my_flags = reinterpret_cast<uint32_t*>(c->base + c->off_flags());
    auto flag_at = [&](int rank_of_buffer, int which, int bucket, int src) {
        return c->peer_base[rank_of_buffer] + c->off_flags() + (size_t(which) * c->n_buckets * N + size_t(bucket) * N + src) * sizeof(uint32_t);
    };
    // 1. push my partial shards; the flag value sits in a per-peer scratch word written in stream order
    for (int p = 0; p < N; ++p) {
        if (p == me) continue;
        cudaStream_t s = c->send[p];
        const int slot = me < p ? me : me - 1;                       // index of me among the other ranks of p
        char* dst = c->peer_base[p] + c->off_staging() + size_t(slot) * slot_stride + offset / N;
        BDOF_TRY(split_copy(c, dst, c->base + offset + size_t(p) * shard, shard, s, c->sub_send[p], c->sub_send_ev[p], ready));
        if (c->flag_memop) {
            CU_TRY(g_write32(s, (CUdeviceptr)flag_at(p, 0, j, me), e, 0));
        } else {
            char* scratch = c->base + c->off_scratch() + size_t(p) * 256 + (e % DP_SCRATCH_RING) * sizeof(uint32_t);
            CU_TRY(g_write32(s, (CUdeviceptr)scratch, e, 0));
            CUDA_TRY(cudaMemcpyAsync(flag_at(p, 0, j, me), scratch, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        }
    }
    // 2. reduce my shard once every partial has landed
    CUDA_TRY(cudaStreamWaitEvent(c->red, ready, 0));
    for (int p = 0; p < N; ++p) {
        if (p == me) continue;
        CU_TRY(g_wait32(c->red, (CUdeviceptr)(my_flags + size_t(j) * N + p), e, CU_STREAM_WAIT_VALUE_GEQ));
    }
    {
        const long long n4 = (long long)(shard / 16);
        float4* own = reinterpret_cast<float4*>(c->base + offset + size_t(me) * shard);
        const float4* staged = reinterpret_cast<const float4*>(c->base + c->off_staging() + offset / N);
        long long blocks = (n4 + 256 * 8 - 1) / (256 * 8);            // 32 KB of output per CTA: short-lived CTAs that slip in between sweep kernels
        if (blocks > 148 * 16) blocks = 148 * 16;
        if (blocks < 1) blocks = 1;
        k_dp_reduce<<<unsigned(blocks), 256, 0, c->red>>>(own, staged, n4, (long long)(slot_stride / 16), N - 1, 1.0f / float(N));
        BDOF_TRY(bdof_launch_check("k_dp_reduce"));
    }
    CUDA_TRY(cudaEventRecord(c->red_done[j], c->red));
    // 3. gather: my reduced shard into every peer's gradient
    for (int p = 0; p < N; ++p) {
        if (p == me) continue;
        cudaStream_t s = c->gath[p];
        BDOF_TRY(split_copy(c, c->peer_base[p] + offset + size_t(me) * shard, c->base + offset + size_t(me) * shard, shard, s,
                            c->sub_gath[p], c->sub_gath_ev[p], c->red_done[j]));
        if (c->flag_memop) {
            CU_TRY(g_write32(s, (CUdeviceptr)flag_at(p, 1, j, me), e, 0));
        } else {
            char* scratch = c->base + c->off_scratch() + size_t(p) * 256 + 128 + (e % DP_SCRATCH_RING) * sizeof(uint32_t);
            CU_TRY(g_write32(s, (CUdeviceptr)scratch, e, 0));
            CUDA_TRY(cudaMemcpyAsync(flag_at(p, 1, j, me), scratch, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        }
    }
    return 0;
}

// Gather half only: my shard of the bucket (shard `rank` of its `world` equal parts) already holds the reduced values once
// `ready_event` has fired (e.g. after an in-place NCCL reduce-scatter); copy it into every peer's gradient, one peer after the
// other starting with rank + 1, and flag B[j][me] there.  bdof_dp_finish waits for the peers' shards as usual.
extern "C" int bdof_dp_gather(bdof_dp* c, size_t offset, size_t n_bytes, void* ready_event) {
    if (!c || !ready_event) return bdof_fail(BDOF_E_BADARG, "null");
    const int N = c->world, me = c->rank;
    if (c->buckets_issued >= c->n_buckets) return bdof_fail(BDOF_E_STATE, "more buckets than the context was created for");
    if (offset + n_bytes > c->grad_bytes || n_bytes % (size_t(N) * 16) != 0 || offset % 16 != 0)
        return bdof_fail(BDOF_E_BADARG, "bucket [%zu, +%zu) does not split into %d aligned shards", offset, n_bytes, N);
    for (int p = 0; p < N; ++p)
        if (p != me && !c->opened[p]) return bdof_fail(BDOF_E_STATE, "bdof_dp_connect has not been called");
    const int j = c->buckets_issued++;
    if (j == 0) ++c->epoch;
    const uint32_t e = c->epoch;
    const size_t shard = n_bytes / N;
    CUDA_TRY(cudaStreamWaitEvent(c->gseq, reinterpret_cast<cudaEvent_t>(ready_event), 0));
    for (int k = 1; k < N; ++k) {
        const int p = (me + k) % N;
        CUDA_TRY(cudaMemcpyAsync(c->peer_base[p] + offset + size_t(me) * shard, c->base + offset + size_t(me) * shard, shard,
                                 cudaMemcpyDeviceToDevice, c->gseq));
        char* flag = c->peer_base[p] + c->off_flags() + (size_t(1) * c->n_buckets * N + size_t(j) * N + me) * sizeof(uint32_t);
        CU_TRY(g_write32(c->gseq, (CUdeviceptr)flag, e, 0));
    }
    return 0;
}

// Make `stream` (the one the optimiser / next forward runs on) wait until every bucket of this step is complete on this rank:
// all reduced shards of the peers have landed in my gradient and my own copies have left.
extern "C" int bdof_dp_finish(bdof_dp* c, void* stream) {
    if (!c) return bdof_fail(BDOF_E_BADARG, "null");
    const int N = c->world, me = c->rank;
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t e = c->epoch;
    uint32_t* flags_b = reinterpret_cast<uint32_t*>(c->base + c->off_flags()) + size_t(c->n_buckets) * N;
    for (int j = 0; j < c->buckets_issued; ++j)
        for (int p = 0; p < N; ++p) {
            if (p == me) continue;
            CU_TRY(g_wait32(c->red, (CUdeviceptr)(flags_b + size_t(j) * N + p), e, CU_STREAM_WAIT_VALUE_GEQ));
        }
    CUDA_TRY(cudaEventRecord(c->fin, c->red));
    CUDA_TRY(cudaStreamWaitEvent(st, c->fin, 0));
    for (int p = 0; p < N; ++p) {
        if (p == me) continue;
        CUDA_TRY(cudaEventRecord(c->send_done[p], c->send[p]));
        CUDA_TRY(cudaStreamWaitEvent(st, c->send_done[p], 0));
        CUDA_TRY(cudaEventRecord(c->gath_done[p], c->gath[p]));
        CUDA_TRY(cudaStreamWaitEvent(st, c->gath_done[p], 0));
    }
    CUDA_TRY(cudaEventRecord(c->gseq_done, c->gseq));
    CUDA_TRY(cudaStreamWaitEvent(st, c->gseq_done, 0));
    c->buckets_issued = 0;
    return 0;
}
