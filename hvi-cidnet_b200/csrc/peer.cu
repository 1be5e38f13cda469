// Peer-memory transport of the row-strip sharded forward (BASELINE.json configs[4]: one 4K image over the GPUs of a box).
//
// Every rank's workspace is one cudaMalloc'ed block that the other ranks of the node map through CUDA IPC
// (cidnet_peer_alloc / cidnet_peer_open), so a kernel on GPU r can read GPU r+-1's activations directly over
// NVLink / NVSwitch.  Two kernels replace the host callbacks (NCCL send/recv pairs, all-reduce) of cidnet_forward_sharded:
//
//   peer_halo_kernel       pulls the neighbours' boundary rows of a list of NHWC tensors into this rank's halo rows
//   peer_allreduce_kernel  sums the ranks' partial [Gram | sum q^2 | sum k^2] vectors in RANK ORDER (every rank reads
//                          every peer's vector: 8 x <= 23 KB) -- identical bits on every rank, run to run
//
// Both are ordinary kernels on the forward's stream -- the whole sharded forward is one CUDA graph of this library's own
// kernels, no NCCL operation on the critical path (15 of them, ~50 us each, in round 1).  Synchronisation is a handshake
// on flags in the workspace headers, numbered by a per-rank exchange counter that the kernels themselves advance (so a
// replayed graph needs no changing argument):
//
//   1. "ready": rank r writes v into ready[r] of every partner -- its tensors for exchange v are complete (all
//      producing kernels precede this kernel in stream order);
//   2. every CTA waits until its partners' ready flags have reached v, then reads their memory;
//   3. "ack": the LAST CTA to finish writes v into ack[r] of every partner -- and waits for the partners' acks before the
//      kernel completes, so nothing this rank launches afterwards can overwrite rows a partner is still reading.
//
// All ranks execute the same sequence of exchanges (the schedule depends only on the shard geometry), so the counters
// stay in lockstep.  A spin that does not complete within ~2 s sets hdr.error and gives up instead of hanging the GPU.
#include "peer.cuh"
#include "ptx_sm100.cuh"

namespace cidnet {

namespace {

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// wait until *flag >= v (wrap-safe); false on timeout
__device__ __forceinline__ bool spin_until(const uint32_t* flag, uint32_t v, PeerHdr* me) {
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(flag) - v) < 0) {
        if (clock64() - t0 > 4000000000ll) { me->error = 1u; return false; }
        __nanosleep(64);
    }
    return true;
}

// first step of both kernels: the exchange number, the ready signal, the wait for the partners' ready signals
__device__ __forceinline__ uint32_t handshake_begin(const PeerSync& s) {
    __shared__ uint32_t s_v;
    if (threadIdx.x == 0) {
        const uint32_t v = *reinterpret_cast<volatile uint32_t*>(&s.me->seq) + 1u;
        if (blockIdx.x == 0) {
            __threadfence_system();
            for (int i = 0; i < s.npartners; ++i) st_release_sys(&s.partner[i]->ready[s.rank], v);
        }
        for (int i = 0; i < s.npartners; ++i) spin_until(&s.me->ready[s.partner_rank[i]], v, s.me);
        s_v = v;
    }
    __syncthreads();
    return s_v;
}

// last step: the last CTA acknowledges, waits for the partners' acknowledgements and publishes the new exchange number
__device__ __forceinline__ void handshake_end(const PeerSync& s, uint32_t v) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t t = atomicAdd(&s.me->arrive, 1u);
        if (t == gridDim.x - 1) {
            __threadfence_system();
            for (int i = 0; i < s.npartners; ++i) st_release_sys(&s.partner[i]->ack[s.rank], v);
            for (int i = 0; i < s.npartners; ++i) spin_until(&s.me->ack[s.partner_rank[i]], v, s.me);
            s.me->arrive = 0u;
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(&s.me->seq) = v;
        }
    }
}

__global__ void __launch_bounds__(256)
peer_halo_kernel(const PeerHaloArgs a) {
    ptx::pdl_wait();                       // (launched without the PDL attribute; harmless)
    const uint32_t v = handshake_begin(a.sync);
    // the jobs are byte ranges (multiples of 16 B): grid-stride over each of them
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int j = 0; j < a.njobs; ++j) {
        const long long n16 = a.job[j].bytes >> 4;
        const uint4* src = reinterpret_cast<const uint4*>(a.job[j].src);
        uint4* dst = reinterpret_cast<uint4*>(a.job[j].dst);
        long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n16; i += 4 * stride) {       // four NVLink round trips in flight per thread
            const uint4 v0 = ld_relaxed_sys_v4(src + i), v1 = ld_relaxed_sys_v4(src + i + stride);
            const uint4 v2 = ld_relaxed_sys_v4(src + i + 2 * stride), v3 = ld_relaxed_sys_v4(src + i + 3 * stride);
            dst[i] = v0; dst[i + stride] = v1; dst[i + 2 * stride] = v2; dst[i + 3 * stride] = v3;
        }
        for (; i < n16; i += stride) dst[i] = ld_relaxed_sys_v4(src + i);
    }
    handshake_end(a.sync, v);
}

__global__ void __launch_bounds__(256)
peer_allreduce_kernel(const PeerReduceArgs a) {
    ptx::pdl_wait();
    const uint32_t v = handshake_begin(a.sync);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.count; i += gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int r = 0; r < a.nranks; ++r) acc += ld_relaxed_sys_f32(a.src[r] + i);    // rank order: same bits everywhere
        a.dst[i] = acc;
    }
    handshake_end(a.sync, v);
}

}  // namespace

int launch_peer_halo(const PeerHaloArgs& a, cudaStream_t stream) {
    CIDNET_CHECK(a.njobs >= 0 && a.njobs <= kPeerMaxJobs, CIDNET_ERR_INVALID, "peer halo: too many jobs");
    long long bytes = 0;
    for (int j = 0; j < a.njobs; ++j) {
        CIDNET_CHECK((a.job[j].bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(a.job[j].src) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(a.job[j].dst) & 15) == 0, CIDNET_ERR_INVALID, "peer halo: 16-byte alignment");
        bytes += a.job[j].bytes;
    }
    int grid = (int)std::min<long long>(64, std::max<long long>(1, bytes / (256 * 16 * 4)));
    peer_halo_kernel<<<grid, 256, 0, stream>>>(a);
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

int launch_peer_allreduce(const PeerReduceArgs& a, cudaStream_t stream) {
    CIDNET_CHECK(a.nranks >= 1 && a.nranks <= kPeerMaxRanks && a.count > 0, CIDNET_ERR_INVALID, "peer allreduce: bad arguments");
    const int grid = std::max(1, std::min(16, a.count / 512));
    peer_allreduce_kernel<<<grid, 256, 0, stream>>>(a);
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

}  // namespace cidnet

// ------------------------------------------------------------------ C ABI: IPC plumbing ---
using namespace cidnet;

extern "C" int cidnet_peer_alloc(int device, int64_t bytes, void** dev_ptr, void* handle64) {
    CIDNET_CHECK(dev_ptr && handle64 && bytes > 0, CIDNET_ERR_INVALID, "peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    int prev = 0;
    CIDNET_CUDA_OK(cudaGetDevice(&prev));
    CIDNET_CUDA_OK(cudaSetDevice(device));
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);          // flags and counters start at zero
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaSetDevice(prev);
    if (e != cudaSuccess) { if (p) cudaFree(p); return fail(CIDNET_ERR_CUDA, std::string("peer_alloc: ") + cudaGetErrorString(e)); }
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return CIDNET_OK;
}

extern "C" int cidnet_peer_open(int device, const void* handle64, void** dev_ptr) {
    CIDNET_CHECK(dev_ptr && handle64, CIDNET_ERR_INVALID, "peer_open: bad arguments");
    int prev = 0;
    CIDNET_CUDA_OK(cudaGetDevice(&prev));
    CIDNET_CUDA_OK(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(CIDNET_ERR_CUDA, std::string("peer_open: ") + cudaGetErrorString(e));
    *dev_ptr = p;
    return CIDNET_OK;
}

extern "C" int cidnet_peer_close(void* dev_ptr) {
    if (dev_ptr) CIDNET_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
    return CIDNET_OK;
}

extern "C" int cidnet_peer_free(void* dev_ptr) {
    if (dev_ptr) CIDNET_CUDA_OK(cudaFree(dev_ptr));
    return CIDNET_OK;
}

extern "C" int cidnet_peer_error(const void* own_ws, int* error_out) {
    CIDNET_CHECK(own_ws && error_out, CIDNET_ERR_INVALID, "peer_error: bad arguments");
    PeerHdr h;
    CIDNET_CUDA_OK(cudaMemcpy(&h, own_ws, sizeof h, cudaMemcpyDeviceToHost));
    *error_out = (int)h.error;
    return CIDNET_OK;
}
