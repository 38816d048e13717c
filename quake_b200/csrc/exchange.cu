// Shard exchange for lists sharded over the GPUs of one box: partial top-k of every rank -> merged top-k on every
// rank, as ONE kernel per rank over NVLink peer memory (no NCCL call on the query path).
//
// Replaces, across GPUs, the reference's merge of per-core partial results into the global TopkBuffer
// (/root/reference/src/cpp/src/query_coordinator.cpp:167-173, 752-758). Every rank owns a "peer
// buffer" (cudaMalloc + CUDA IPC handle, opened by all other ranks of the box):
//     header   arrive[2]  u32   counters, one per parity, bumped by the peers' CTAs after their stores
//              epoch      u32   number of exchanges this rank has completed (device-side: graph-replayable)
//     data     [2 parities][world sources][Q x k] ids (i64) + [Q x k] distances (f32)
// exchange_merge_kernel (persistent grid, the same on every rank):
//   phase 1  push: every CTA stores its share of the local partial straight into slot [parity][rank] of EVERY peer's
//            buffer (remote stores over NVLink; 120 KB per peer at Q = 1024, k = 10), then one system-scope fence and
//            one remote atomicAdd per peer
//   phase 2  wait until the world * grid arrivals of this epoch have landed in the local header, then merge the
//            world partial lists of each query by (distance, id) -- the order of the single-GPU refine step, so the
//            result is bit-identical to the unsharded search -- and write the top-k
// No CTA waits before it has pushed, and the grid is co-resident (<= 2 CTAs per SM), so ranks cannot deadlock each
// other. Two parities: a rank can run at most one exchange ahead of the slowest rank (it needs every peer's push of
// exchange n to finish n), so the buffer of exchange n is never overwritten before everyone has merged it.
#include "common.cuh"
#include <cstring>

namespace qk {

struct PeerHeader {
    unsigned int arrive[2];
    unsigned int epoch;
    unsigned int pad_[13];
};
static_assert(sizeof(PeerHeader) == 64, "peer header");
static constexpr int XCHG_MAX_WORLD = 16;
static constexpr int XCHG_THREADS = 256;

struct ExchangeArgs {
    const int64_t* ids;   // [Q x k] local partial
    const float* dist;
    int64_t Q;
    int k, rank, world;
    unsigned char* peers[XCHG_MAX_WORLD];  // peer buffers (device pointers valid on this GPU); peers[rank] is the local one
    int64_t* out_ids;
    float* out_dist;
};

__host__ __device__ inline size_t xchg_slot_bytes(int64_t Q, int k) { return (size_t)Q * k * 12; }
__host__ __device__ inline size_t xchg_buffer_bytes(int64_t Q, int k, int world) {
    return sizeof(PeerHeader) + 2 * (size_t)world * xchg_slot_bytes(Q, k);
}
__device__ __forceinline__ int64_t* slot_ids(unsigned char* buf, int parity, int src, int64_t Q, int k, int world) {
    return reinterpret_cast<int64_t*>(buf + sizeof(PeerHeader) + ((size_t)parity * world + src) * xchg_slot_bytes(Q, k));
}
__device__ __forceinline__ float* slot_dist(unsigned char* buf, int parity, int src, int64_t Q, int k, int world) {
    return reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(slot_ids(buf, parity, src, Q, k, world)) + (size_t)Q * k * 8);
}

template <bool kIP>
__global__ void __launch_bounds__(XCHG_THREADS) exchange_merge_kernel(const ExchangeArgs a) {
    extern __shared__ __align__(16) unsigned char xsm[];
    const int tid = threadIdx.x;
    const int world = a.world, k = a.k;
    const int64_t Q = a.Q;
    PeerHeader* local = reinterpret_cast<PeerHeader*>(a.peers[a.rank]);
    const unsigned epoch = *reinterpret_cast<volatile unsigned*>(&local->epoch);
    const int parity = epoch & 1u;
    const int64_t nk = Q * k;

    // ---- phase 1: push this CTA's share of the local partial into every rank's buffer (the local one included)
    for (int p = 0; p < world; ++p) {
        int64_t* di = slot_ids(a.peers[p], parity, a.rank, Q, k, world);
        float* dd = slot_dist(a.peers[p], parity, a.rank, Q, k, world);
        for (int64_t i = (int64_t)blockIdx.x * XCHG_THREADS + tid; i < nk; i += (int64_t)gridDim.x * XCHG_THREADS) {
            di[i] = a.ids[i];
            dd[i] = a.dist[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    // one arrival per CTA on EVERY rank's counter, the local one included (the local slot is written by other CTAs too)
    if (tid < world) atomicAdd_system(&reinterpret_cast<PeerHeader*>(a.peers[tid])->arrive[parity], 1u);

    // ---- phase 2: wait for the peers' pushes of this epoch, merge
    if (tid == 0) {
        const unsigned want = (epoch / 2u + 1u) * (unsigned)world * gridDim.x;
        volatile unsigned* f = &local->arrive[parity];
        unsigned spins = 0;
        while ((int)(*f - want) < 0) {
            __nanosleep(100);
            if (++spins == 300000000u) {  // >= 30 s: a peer never arrived (ranks may be seconds apart at a first call)
                printf("quake_b200: shard exchange timed out on rank %d (epoch %u, %u of %u arrivals)\n", a.rank, epoch, *f, want);
                __trap();
            }
        }
        __threadfence_system();
    }
    __syncthreads();

    const int n = world * k;
    int np = 1;
    while (np < n) np <<= 1;
    uint64_t* key = reinterpret_cast<uint64_t*>(xsm);      // [np] distkey << 32 | slot
    int64_t* ids = reinterpret_cast<int64_t*>(key + np);   // [n]
    float* dst = reinterpret_cast<float*>(ids + n);        // [n]
    int* rank_of = reinterpret_cast<int*>(dst + n);        // [n]
    unsigned char* lb = a.peers[a.rank];
    for (int64_t q = blockIdx.x; q < Q; q += gridDim.x) {
        __syncthreads();
        for (int i = tid; i < np; i += XCHG_THREADS) {
            uint64_t kv = COMP_MAX;
            if (i < n) {
                const int s = i / k, j = i - s * k;
                // volatile: the data was written by other GPUs; do not let it be served from a stale L1 line
                const int64_t id = *reinterpret_cast<volatile int64_t*>(slot_ids(lb, parity, s, Q, k, world) + q * k + j);
                const float d = *reinterpret_cast<volatile float*>(slot_dist(lb, parity, s, Q, k, world) + q * k + j);
                ids[i] = id;
                dst[i] = d;
                if (id >= 0) kv = ((uint64_t)f2key(kIP ? -d : d) << 32) | (uint32_t)i;
            }
            key[i] = kv;
        }
        __syncthreads();
        if (n <= XCHG_THREADS) {
            // rank by counting: position of entry i among the valid entries ordered by (distance key, id)
            if (tid < n) {
                const uint64_t ki = key[tid];
                int r = -1;
                if (ki != COMP_MAX) {
                    r = 0;
                    const uint32_t di = (uint32_t)(ki >> 32);
                    const int64_t idi = ids[tid];
                    for (int j = 0; j < n; ++j) {
                        const uint64_t kj = key[j];
                        if (kj == COMP_MAX) continue;
                        const uint32_t dj = (uint32_t)(kj >> 32);
                        const int64_t idj = ids[j];
                        r += (dj < di || (dj == di && (idj < idi || (idj == idi && j < tid)))) ? 1 : 0;
                    }
                }
                rank_of[tid] = r;
            }
            __syncthreads();
            if (tid < k) { a.out_ids[q * k + tid] = -1; a.out_dist[q * k + tid] = kIP ? -INFINITY : INFINITY; }
            __syncthreads();
            if (tid < n) {
                const int r = rank_of[tid];
                if (r >= 0 && r < k) { a.out_ids[q * k + r] = ids[tid]; a.out_dist[q * k + r] = dst[tid]; }
            }
        } else {
            const int64_t* idc = ids;
            for (int size = 2; size <= np; size <<= 1) {
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    __syncthreads();
                    for (int i = tid; i < (np >> 1); i += XCHG_THREADS) {
                        int lo = 2 * i - (i & (stride - 1));
                        int hi = lo + stride;
                        bool up = ((lo & size) == 0);
                        uint64_t x = key[lo], y = key[hi];
                        auto less = [idc](uint64_t u, uint64_t v) {
                            const uint32_t du = (uint32_t)(u >> 32), dv = (uint32_t)(v >> 32);
                            if (du != dv) return du < dv;
                            if (u == COMP_MAX || v == COMP_MAX) return u < v;
                            return idc[(uint32_t)u] < idc[(uint32_t)v];
                        };
                        if (up ? less(y, x) : less(x, y)) { key[lo] = y; key[hi] = x; }
                    }
                }
            }
            __syncthreads();
            for (int i = tid; i < k; i += XCHG_THREADS) {
                int64_t id = -1;
                float d = kIP ? -INFINITY : INFINITY;
                if (i < n && key[i] != COMP_MAX) {
                    const uint32_t slot = (uint32_t)key[i];
                    id = ids[slot];
                    d = dst[slot];
                }
                a.out_ids[q * k + i] = id;
                a.out_dist[q * k + i] = d;
            }
        }
    }
    // ---- this exchange is complete on this rank once every CTA has merged its queries: the last one bumps the epoch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&local->pad_[0], 1u) + 1u;
        if (done == gridDim.x) {
            local->pad_[0] = 0u;
            __threadfence();
            *reinterpret_cast<volatile unsigned*>(&local->epoch) = epoch + 1u;
        }
    }
}

}  // namespace qk

using namespace qk;

extern "C" size_t qk_peer_buffer_bytes(int64_t num_queries, int k, int world) {
    if (num_queries <= 0 || k <= 0 || world <= 0 || world > XCHG_MAX_WORLD) return 0;
    return xchg_buffer_bytes(num_queries, k, world);
}

extern "C" int qk_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_64) {
    QK_REQUIRE(bytes >= sizeof(PeerHeader) && dev_ptr && ipc_handle_64, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    QK_CUDA(cudaMalloc(&p, bytes));
    QK_CUDA(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(e, "cudaIpcGetMemHandle");
    }
    memcpy(ipc_handle_64, &h, 64);
    *dev_ptr = p;
    return QK_OK;
}

extern "C" int qk_peer_open(const void* ipc_handle_64, void** dev_ptr) {
    QK_REQUIRE(ipc_handle_64 && dev_ptr, "bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle_64, 64);
    void* p = nullptr;
    QK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = p;
    return QK_OK;
}

extern "C" int qk_peer_close(void* dev_ptr) {
    if (dev_ptr) QK_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return QK_OK;
}

extern "C" int qk_peer_free(void* dev_ptr) {
    if (dev_ptr) QK_CUDA(cudaFree(dev_ptr));
    return QK_OK;
}

extern "C" int qk_exchange_merge_topk(const int64_t* ids, const float* distances, int64_t Q, int k, int metric, int rank,
                                      int world, void* const* peer_buffers, int64_t* out_ids, float* out_distances,
                                      void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(ids && distances && peer_buffers && out_ids && out_distances, "null argument");
    QK_REQUIRE(world >= 1 && world <= XCHG_MAX_WORLD && rank >= 0 && rank < world && Q > 0 && k > 0, "bad argument");
    QK_REQUIRE((int64_t)world * k <= 8192, "world * k = %lld exceeds 8192", (long long)world * k);
    ExchangeArgs a;
    memset(&a, 0, sizeof(a));
    a.ids = ids; a.dist = distances; a.Q = Q; a.k = k; a.rank = rank; a.world = world;
    for (int p = 0; p < world; ++p) {
        QK_REQUIRE(peer_buffers[p] != nullptr, "peer buffer %d is null", p);
        a.peers[p] = (unsigned char*)peer_buffers[p];
    }
    a.out_ids = out_ids; a.out_dist = out_distances;
    const int n = world * k;
    int np = 1;
    while (np < n) np <<= 1;
    const size_t smem = (size_t)np * 8 + (size_t)n * (8 + 4 + 4) + 16;
    // the same grid on every rank (the arrival count depends on it) and co-resident (a CTA that waits must not keep a
    // CTA that still has to push from starting): two CTAs per SM of a B200, one when the merge buffers are large
    const int grid = smem > 100 * 1024 ? 148 : 2 * 148;
    if (metric == QK_METRIC_INNER_PRODUCT) {
        QK_CUDA(cudaFuncSetAttribute(exchange_merge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        exchange_merge_kernel<true><<<grid, XCHG_THREADS, smem, stream>>>(a);
    } else {
        QK_CUDA(cudaFuncSetAttribute(exchange_merge_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        exchange_merge_kernel<false><<<grid, XCHG_THREADS, smem, stream>>>(a);
    }
    QK_LAUNCHED();
    return QK_OK;
}
