// p2p.cu -- the validation flow's one exchange (all-gather of the padded detections for AP) as direct peer-memory
// stores over NVLink 5 / NVSwitch instead of a ring collective.
//
// One process per GPU.  Every rank owns a receive buffer [N blocks] and a flag array [N] in its own HBM, allocated with
// jabd_p2p_alloc (cudaMalloc + cudaIpcGetMemHandle); the 64-byte handles travel over the job's host channel (torch.distributed
// all_gather_object in sharding.PeerGather) and every other rank maps them with jabd_p2p_open.  An exchange is then ONE kernel
// per rank: p2p_allgather_kernel copies the rank's block into slot `rank` of all N receive buffers with 16-byte stores (N - 1 of
// them cross the NVSwitch, a single hop each -- a 1.44 MB message on an 8-rank ring is seven latency-bound steps), and the last
// CTA per peer publishes `seq` in that peer's flag word with a system-scope release.  The consumer enqueues p2p_wait_kernel on
// whichever stream reads the gathered rows: one thread per peer spins (acquire, system scope) until its flag reaches `seq`.
// Nothing here synchronises the host; the producer stream goes straight on to the next batch.
//
// Reuse of a receive slot is the caller's protocol (PeerGather: as many buffer sets as exchanges may be in flight, and a rank
// only starts exchange s + depth after it has waited for exchange s + 1 .. of every peer on the same stream).
#include <cstring>

#include "common.cuh"

namespace jabd {

struct P2PArgs {
    void *dst[JABD_P2P_MAX_PEERS];
    unsigned long long *flag[JABD_P2P_MAX_PEERS];
    unsigned long long *ack[JABD_P2P_MAX_PEERS];   // peer j's acknowledgement array (null: no handshake)
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// grid (ctas_per_peer, N).  counters: [N] unsigned, zero before the first launch; every launch leaves them zero again.
// Handshake (ack_seq != 0): before a block of peer j's receive buffer is overwritten, peer j must have said that it has read the
// previous contents.  CTA (0, j) tells peer j "I have read what you sent me last time" (ack[j][rank] = ack_seq: this kernel is
// stream-ordered behind this rank's reads), then every CTA of column j waits for own_acks[j] >= ack_seq before it stores.
__global__ void __launch_bounds__(256) p2p_allgather_kernel(const uint4 *__restrict__ src, size_t n16, size_t tail_bytes, P2PArgs a,
                                                            size_t dst_offset, int rank, unsigned long long seq,
                                                            unsigned int *counters, const unsigned long long *own_acks,
                                                            unsigned long long ack_seq, long long timeout_cycles, int *status)
{
    const int peer = blockIdx.y;
    if (ack_seq != 0ull) {
        if (threadIdx.x == 0) {
            if (blockIdx.x == 0) st_release_sys(a.ack[peer] + rank, ack_seq);
            const long long t0 = clock64();
            while (ld_acquire_sys(own_acks + peer) < ack_seq) {
                if (clock64() - t0 > timeout_cycles) { // never hang the device on a peer that died
                    if (status) atomicMax(status, 101 + peer);
                    break;
                }
                __nanosleep(100);
            }
        }
        __syncthreads();
    }
    char *base = static_cast<char *>(a.dst[peer]) + dst_offset;
    uint4 *dst = reinterpret_cast<uint4 *>(base);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // four independent 16-byte transfers in flight per thread
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        const uint4 v0 = __ldg(src + i), v1 = __ldg(src + i + stride), v2 = __ldg(src + i + 2 * stride), v3 = __ldg(src + i + 3 * stride);
        dst[i] = v0; dst[i + stride] = v1; dst[i + 2 * stride] = v2; dst[i + 3 * stride] = v3;
    }
    for (; i < n16; i += stride) dst[i] = __ldg(src + i);
    if (blockIdx.x == 0 && threadIdx.x < tail_bytes) // a size that is not a multiple of 16
        base[n16 * 16 + threadIdx.x] = reinterpret_cast<const char *>(src)[n16 * 16 + threadIdx.x];
    __threadfence_system(); // this thread's peer stores are performed before the arrival below can be observed
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int arrived = atomicAdd(&counters[peer], 1u);
        if (arrived == gridDim.x - 1) {          // last CTA for this peer: every other CTA fenced before it arrived
            counters[peer] = 0u;
            __threadfence_system();
            st_release_sys(a.flag[peer] + rank, seq);
        }
    }
}

// one CTA, one thread per peer.  status (optional): set to 1 + peer if that peer's flag did not arrive in time.
__global__ void p2p_wait_kernel(const unsigned long long *flags, int n, unsigned long long seq, long long timeout_cycles, int *status)
{
    const int t = threadIdx.x;
    if (t >= n) return;
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + t) < seq) {
        if (clock64() - t0 > timeout_cycles) { // never hang the device on a peer that died
            if (status) atomicMax(status, 1 + t);
            break;
        }
        __nanosleep(200);
    }
}

} // namespace jabd

using namespace jabd;

extern "C" {

int jabd_p2p_alloc(size_t bytes, void **dev_ptr, unsigned char *handle64)
{
    JABD_REQUIRE(bytes > 0 && dev_ptr && handle64, JABD_EINVAL, "p2p_alloc: zero size or null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == JABD_P2P_HANDLE_BYTES, "IPC handle size");
    void *p = nullptr;
    JABD_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(e, "p2p_alloc");
    }
    memcpy(handle64, &h, sizeof(h));
    *dev_ptr = p;
    return JABD_OK;
}

int jabd_p2p_open(const unsigned char *handle64, void **dev_ptr)
{
    JABD_REQUIRE(handle64 && dev_ptr, JABD_EINVAL, "p2p_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    JABD_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return JABD_OK;
}

int jabd_p2p_close(void *dev_ptr)
{
    if (dev_ptr) JABD_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return JABD_OK;
}

int jabd_p2p_free(void *dev_ptr)
{
    if (dev_ptr) JABD_CUDA(cudaFree(dev_ptr));
    return JABD_OK;
}

// the time-outs are counted in SM clocks; 2.5 GHz bounds every sm_100 part from above (a wait then lasts at least the asked time).
// Not queried: cudaDevAttrClockRate is a millisecond-class driver call.
static long long timeout_cycles_of(double timeout_s)
{
    if (!(timeout_s > 0.0)) timeout_s = 2.0;
    return (long long)(timeout_s * 2.5e9);
}

int jabd_p2p_allgather(const void *src, size_t bytes, void *const *peer_bufs, size_t dst_offset,
                       unsigned long long *const *peer_flags, unsigned long long *const *peer_acks,
                       const unsigned long long *own_acks, int n_ranks, int rank, unsigned long long seq, unsigned long long ack_seq,
                       unsigned int *counters, double timeout_s, int *status, jabd_stream_t stream)
{
    JABD_REQUIRE(n_ranks >= 1 && n_ranks <= JABD_P2P_MAX_PEERS && rank >= 0 && rank < n_ranks, JABD_EINVAL,
                 "p2p_allgather: %d ranks (at most %d), rank %d", n_ranks, JABD_P2P_MAX_PEERS, rank);
    JABD_REQUIRE((src || bytes == 0) && peer_bufs && peer_flags && counters, JABD_EINVAL, "p2p_allgather: null pointer");
    JABD_REQUIRE(ack_seq == 0 || (peer_acks && own_acks), JABD_EINVAL, "p2p_allgather: a handshake needs peer_acks and own_acks");
    JABD_REQUIRE(aligned_to(src, 16) && (dst_offset & 15u) == 0, JABD_EALIGN, "p2p_allgather: src and dst_offset must be 16-byte aligned");
    // bytes == 0: nothing is copied, the flags are still raised
    P2PArgs a;
    for (int i = 0; i < JABD_P2P_MAX_PEERS; ++i) { a.dst[i] = nullptr; a.flag[i] = nullptr; a.ack[i] = nullptr; }
    for (int i = 0; i < n_ranks; ++i) {
        JABD_REQUIRE(peer_bufs[i] && peer_flags[i] && aligned_to(peer_bufs[i], 16) && aligned_to(peer_flags[i], 8), JABD_EINVAL,
                     "p2p_allgather: peer %d buffer / flag pointer null or misaligned", i);
        a.dst[i] = peer_bufs[i];
        a.flag[i] = peer_flags[i];
        if (ack_seq != 0) {
            JABD_REQUIRE(peer_acks[i] && aligned_to(peer_acks[i], 8), JABD_EINVAL, "p2p_allgather: peer %d ack pointer null or misaligned", i);
            a.ack[i] = peer_acks[i];
        }
    }
    const size_t n16 = bytes / 16, tail = bytes % 16;
    // enough CTAs to keep every link busy, few enough to slip in beside a resident detect kernel (256 threads, no shared memory)
    unsigned ctas = (unsigned)((n16 + 256 * 16 - 1) / (256 * 16));
    ctas = ctas < 1 ? 1 : (ctas > 16 ? 16 : ctas);
    p2p_allgather_kernel<<<dim3(ctas, (unsigned)n_ranks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4 *>(src), n16, tail, a, dst_offset, rank, seq, counters, own_acks, ack_seq, timeout_cycles_of(timeout_s),
        status);
    JABD_LAUNCH_CHECK("p2p_allgather_kernel");
    return JABD_OK;
}

int jabd_p2p_wait(const unsigned long long *flags, int n_ranks, unsigned long long seq, double timeout_s, int *status,
                  jabd_stream_t stream)
{
    JABD_REQUIRE(flags && n_ranks >= 1 && n_ranks <= JABD_P2P_MAX_PEERS, JABD_EINVAL, "p2p_wait: null flags or bad rank count");
    const long long cycles = timeout_cycles_of(timeout_s);
    p2p_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(flags, n_ranks, seq, cycles, status);
    JABD_LAUNCH_CHECK("p2p_wait_kernel");
    return JABD_OK;
}

} // extern "C"
