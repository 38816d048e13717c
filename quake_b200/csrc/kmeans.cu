// k-means update / scatter kernels and small utilities for sm_100a.
//
// The k-means ASSIGN step (reference: faiss::IndexFlat::search(k=1) inside faiss::Clustering::train,
// /root/reference/src/cpp/src/clustering.cpp:54-65; batched_scan_list(k=1) in
// kmeans_refine_partitions, clustering.cpp:152-159) runs on the partition-scan kernel of scan.cu with
// the centroid matrix as a flat store and the points as queries (k = 1); see qk_kmeans_assign below.
#include "common.cuh"

namespace qk {

// ------------------------------------------------------------------------------------------------
// counting sort of point indices by assignment (torch::sort + bincount + split, clustering.cpp:68-84)
// ------------------------------------------------------------------------------------------------
__global__ void assign_hist_kernel(const int32_t* __restrict__ assign, int64_t n, int64_t K,
                                   unsigned long long* __restrict__ counts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a = assign[i];
    if (a >= 0 && a < K) atomicAdd(&counts[a], 1ull);
}

// exclusive scan of K int64 counts by one CTA -> offsets[K+1]; also clears the cursors
__global__ void __launch_bounds__(1024) offsets_kernel(const unsigned long long* __restrict__ counts, int64_t K,
                                                       int64_t* __restrict__ offsets,
                                                       unsigned long long* __restrict__ cursor) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < K; base += 1024) {
        int64_t s = base + tid;
        unsigned long long c = (s < K) ? counts[s] : 0ull;
        unsigned long long x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        unsigned long long cy = carry;
        unsigned long long wprev = warp ? warp_tot[warp - 1] : 0ull;
        if (s < K) {
            offsets[s] = (int64_t)(cy + wprev + x - c);
            cursor[s] = 0ull;
        }
        __syncthreads();
        if (tid == 1023) carry = cy + wprev + x;
        __syncthreads();
    }
    if (tid == 0) offsets[K] = (int64_t)carry;
}

__global__ void assign_scatter_kernel(const int32_t* __restrict__ assign, int64_t n, int64_t K,
                                      const int64_t* __restrict__ offsets, unsigned long long* __restrict__ cursor,
                                      int64_t* __restrict__ order) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a = assign[i];
    if (a < 0 || a >= K) return;
    unsigned long long pos = atomicAdd(&cursor[a], 1ull);
    order[offsets[a] + (int64_t)pos] = i;
}

// restore ascending point order inside every list (the atomic scatter is unordered): one CTA per
// list. Lists of <= 4096 members are bitonic-sorted in shared memory; larger ones are sorted in runs
// of 4096 and the runs merged pairwise through `scratch` (rank by binary search; members are distinct).
__device__ void smem_sort_i64(int64_t* s, int np) {
    for (int size = 2; size <= np; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < (np >> 1); i += blockDim.x) {
                int lo = 2 * i - (i & (stride - 1));
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                int64_t x = s[lo], y = s[hi];
                if (up ? (y < x) : (x < y)) { s[lo] = y; s[hi] = x; }
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) sort_lists_kernel(const int64_t* __restrict__ offsets, int64_t K,
                                                         int64_t* __restrict__ order, int64_t* __restrict__ scratch) {
    __shared__ int64_t s[4096];
    for (int64_t c = blockIdx.x; c < K; c += gridDim.x) {
        const int64_t b = offsets[c], e = offsets[c + 1];
        const int64_t n = e - b;
        if (n <= 1) continue;
        // sort runs of 4096
        for (int64_t r0 = 0; r0 < n; r0 += 4096) {
            const int m = (int)((n - r0) < 4096 ? (n - r0) : 4096);
            int np = 1;
            while (np < m) np <<= 1;
            for (int i = threadIdx.x; i < np; i += blockDim.x) s[i] = (i < m) ? order[b + r0 + i] : INT64_MAX;
            smem_sort_i64(s, np);
            for (int i = threadIdx.x; i < m; i += blockDim.x) order[b + r0 + i] = s[i];
            __syncthreads();
        }
        if (n <= 4096) continue;
        int64_t* src = order + b;
        int64_t* dst = scratch + b;
        for (int64_t w = 4096; w < n; w <<= 1) {
            __threadfence_block();
            __syncthreads();
            for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
                const int64_t pair0 = i / (2 * w) * (2 * w);
                const int64_t mid = (pair0 + w) < n ? (pair0 + w) : n;
                const int64_t end = (pair0 + 2 * w) < n ? (pair0 + 2 * w) : n;
                const int64_t v = src[i];
                int64_t lo, hi, pos;
                if (i < mid) {  // left run: count right-run elements smaller than v
                    lo = mid; hi = end;
                    while (lo < hi) { int64_t m2 = (lo + hi) >> 1; if (src[m2] < v) lo = m2 + 1; else hi = m2; }
                    pos = pair0 + (i - pair0) + (lo - mid);
                } else {        // right run: count left-run elements smaller than v
                    lo = pair0; hi = mid;
                    while (lo < hi) { int64_t m2 = (lo + hi) >> 1; if (src[m2] < v) lo = m2 + 1; else hi = m2; }
                    pos = pair0 + (i - mid) + (lo - pair0);
                }
                dst[pos] = v;
            }
            int64_t* t = src; src = dst; dst = t;
        }
        __threadfence_block();
        __syncthreads();
        if (src != order + b)
            for (int64_t i = threadIdx.x; i < n; i += blockDim.x) order[b + i] = src[i];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// per-centroid sums in ascending point order (faiss compute_centroids order,
// third_party/faiss/faiss/Clustering.cpp:123-192): one WARP per centroid, lane l owns dimensions 4l .. 4l+3 of a
// 128-wide dimension block (one 16-byte load per member row and lane, a warp reads 512 contiguous bytes of the row).
// The member indices of 32 rows are fetched with one coalesced load and handed round with shuffles; eight row loads
// are in flight per lane; every dimension still adds its members one by one in list order, so the sums are
// bit-identical to the sequential loop of the reference. HBM-bound: N*d*4 + N*8 + K*d*4 bytes.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) accumulate_kernel(const float* __restrict__ points, int64_t pitch, int d,
                                                         const int64_t* __restrict__ order,
                                                         const int64_t* __restrict__ offsets, int64_t K,
                                                         float* __restrict__ sums, int64_t out_pitch) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool vec = (pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0);
    for (int64_t c = warp; c < K; c += nwarps) {
        const int64_t b = offsets[c], e = offsets[c + 1];
        for (int j0 = 0; j0 < d; j0 += 128) {
            const int j = j0 + 4 * lane;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            for (int64_t base = b; base < e; base += 32) {
                const int m = (int)((e - base) < 32 ? (e - base) : 32);
                const int64_t mine = lane < m ? order[base + lane] : 0;
                for (int t0 = 0; t0 < m; t0 += 8) {
                    float4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int64_t r = __shfl_sync(0xffffffffu, mine, (t0 + u) & 31);
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (t0 + u < m && j < d) {
                            const float* rp = points + r * pitch + j;
                            if (vec && j + 3 < d) {
                                v[u] = *reinterpret_cast<const float4*>(rp);
                            } else {
                                v[u].x = rp[0];
                                if (j + 1 < d) v[u].y = rp[1];
                                if (j + 2 < d) v[u].z = rp[2];
                                if (j + 3 < d) v[u].w = rp[3];
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (t0 + u < m) {  // warp-uniform
                            s0 = __fadd_rn(s0, v[u].x); s1 = __fadd_rn(s1, v[u].y);
                            s2 = __fadd_rn(s2, v[u].z); s3 = __fadd_rn(s3, v[u].w);
                        }
                    }
                }
            }
            if (j < d) {
                float* o = sums + c * out_pitch + j;
                o[0] = s0;
                if (j + 1 < d) o[1] = s1;
                if (j + 2 < d) o[2] = s2;
                if (j + 3 < d) o[3] = s3;
            }
        }
    }
}

__global__ void gather_rows_kernel(const float* __restrict__ src, int64_t src_pitch, const int64_t* __restrict__ src_ids,
                                   const int64_t* __restrict__ order, int64_t n, int d, float* __restrict__ dst,
                                   int64_t dst_pitch, int64_t* __restrict__ dst_ids) {
    // one warp per row
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const int64_t s = order ? order[row] : row;
    const float* sp = src + s * src_pitch;
    float* dp = dst + row * dst_pitch;
    for (int j = lane; j < dst_pitch; j += 32) dp[j] = (j < d) ? sp[j] : 0.f;
    if (lane == 0 && dst_ids) dst_ids[row] = src_ids ? src_ids[s] : s;
}

__global__ void scatter_rows_kernel(const float* __restrict__ src, int64_t src_pitch, const int64_t* __restrict__ src_ids,
                                    const int64_t* __restrict__ order, const int64_t* __restrict__ dst_rows, int64_t n,
                                    int d, float* __restrict__ dst, int64_t dst_pitch, int64_t* __restrict__ dst_ids) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int64_t s = order ? order[i] : i;
    const int64_t r = dst_rows[i];
    const float* sp = src + s * src_pitch;
    float* dp = dst + r * dst_pitch;
    for (int j = lane; j < dst_pitch; j += 32) dp[j] = (j < d) ? sp[j] : 0.f;
    if (lane == 0 && dst_ids) dst_ids[r] = src_ids ? src_ids[s] : s;
}

__global__ void normalize_rows_kernel(float* __restrict__ rows, int64_t n, int64_t pitch, int d) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    float* r = rows + row * pitch;
    float s = 0.f;
    for (int j = lane; j < d; j += 32) s = fmaf(r[j], r[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float nrm = __fsqrt_rn(s);
    for (int j = lane; j < d; j += 32) r[j] = __fdiv_rn(r[j], nrm);
}

__global__ void row_sqnorms_kernel(const float* __restrict__ rows, int64_t n, int64_t pitch, int d,
                                   float* __restrict__ out) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float* r = rows + row * pitch;
    float s = 0.f;
    for (int j = lane; j < d; j += 32) s = fmaf(r[j], r[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = s;
}

__global__ void max_row_norm_kernel(const float* __restrict__ rows, int64_t n, int64_t pitch, int d,
                                    float* __restrict__ out) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    float nrm = 0.f;
    if (row < n) {
        const float* r = rows + row * pitch;
        float s = 0.f;
        for (int j = lane; j < d; j += 32) s = fmaf(r[j], r[j], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        nrm = sqrtf(s) * 1.000001f;  // round up a little: this is an upper bound
    }
    __shared__ float smax[32];
    const int warp = threadIdx.x >> 5;
    if (lane == 0) smax[warp] = nrm;
    __syncthreads();
    if (warp == 0) {
        float m = (lane < (blockDim.x >> 5)) ? smax[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        // non-negative floats order like their bit patterns
        if (lane == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
    }
}

__global__ void map_ids_kernel(const int64_t* __restrict__ ids, int64_t n, const int32_t* __restrict__ table,
                               int64_t table_size, int32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t id = ids[i];
    int32_t s = -1;
    if (id >= 0 && id < table_size) s = table[id];
    out[i] = s < 0 ? -1 : s;
}

// ------------------------------------------------------------------------------------------------
// merge of S partial top-k lists per query (multi-GPU shards / APS rounds)
// ------------------------------------------------------------------------------------------------
template <bool kIP>
__global__ void __launch_bounds__(256) merge_topk_kernel(const float* __restrict__ pd, const int64_t* __restrict__ pi,
                                                         int S, int64_t Q, int k, int64_t* __restrict__ out_ids,
                                                         float* __restrict__ out_dist) {
    extern __shared__ __align__(16) unsigned char msm[];
    const int n = S * k;
    int np = 1;
    while (np < n) np <<= 1;
    uint64_t* key = reinterpret_cast<uint64_t*>(msm);   // [np] distkey<<32 | slot
    int64_t* ids = reinterpret_cast<int64_t*>(key + np);  // [n]
    const int64_t q = blockIdx.x;
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
        uint64_t kv = COMP_MAX;
        if (i < n) {
            const int s = i / k, j = i - s * k;
            const int64_t src = ((int64_t)s * Q + q) * k + j;
            const int64_t id = pi[src];
            ids[i] = id;
            if (id >= 0) {
                const float d = pd[src];
                kv = ((uint64_t)f2key(kIP ? -d : d) << 32) | (uint32_t)i;
            }
        }
        key[i] = kv;
    }
    const int64_t* idc = ids;
    for (int size = 2; size <= np; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < (np >> 1); i += blockDim.x) {
                int lo = 2 * i - (i & (stride - 1));
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                uint64_t x = key[lo], y = key[hi];
                auto less = [idc](uint64_t a, uint64_t b) {
                    const uint32_t da = (uint32_t)(a >> 32), db = (uint32_t)(b >> 32);
                    if (da != db) return da < db;
                    if (a == COMP_MAX || b == COMP_MAX) return a < b;
                    return idc[(uint32_t)a] < idc[(uint32_t)b];
                };
                if (up ? less(y, x) : less(x, y)) { key[lo] = y; key[hi] = x; }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        int64_t id = -1;
        float d = kIP ? -INFINITY : INFINITY;
        if (i < n && key[i] != COMP_MAX) {
            const uint32_t slot = (uint32_t)key[i];
            const int s = slot / k, j = slot - s * k;
            id = ids[slot];
            d = pd[((int64_t)s * Q + q) * k + j];
        }
        out_ids[q * k + i] = id;
        out_dist[q * k + i] = d;
    }
}

// final argmin rows -> int32 assignment (+ optional Euclidean distance / inner product)
__global__ void assign_finish_kernel(const int64_t* __restrict__ rows, const float* __restrict__ dist, int64_t n,
                                     int32_t* __restrict__ out_assign, float* __restrict__ out_dist) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out_assign[i] = (int32_t)rows[i];
    if (out_dist) out_dist[i] = dist[i];
}

}  // namespace qk

using namespace qk;

extern "C" size_t qk_partition_workspace_bytes(int64_t n, int64_t K) { return (size_t)(K + 32) * 8 + (size_t)(n + 32) * 8; }

extern "C" int qk_partition_by_assignment(const int32_t* assign, int64_t n, int64_t K, int64_t* out_counts,
                                          int64_t* out_offsets, int64_t* out_order, void* workspace,
                                          size_t workspace_bytes, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(assign && out_counts && out_offsets && out_order && K > 0 && n >= 0, "bad argument");
    if (workspace_bytes < qk_partition_workspace_bytes(n, K) || !workspace) {
        set_error("workspace too small");
        return QK_ERR_WORKSPACE;
    }
    unsigned long long* cursor = (unsigned long long*)workspace;
    int64_t* scratch = (int64_t*)workspace + (K + 32);
    unsigned long long* counts = (unsigned long long*)out_counts;
    QK_CUDA(cudaMemsetAsync(counts, 0, (size_t)K * 8, stream));
    if (n > 0) {
        assign_hist_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(assign, n, K, counts);
        QK_LAUNCHED();
    }
    offsets_kernel<<<1, 1024, 0, stream>>>(counts, K, out_offsets, cursor);
    QK_LAUNCHED();
    if (n > 0) {
        assign_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(assign, n, K, out_offsets, cursor, out_order);
        QK_LAUNCHED();
        int grid = (int)(K < 65535 ? K : 65535);
        sort_lists_kernel<<<grid, 256, 0, stream>>>(out_offsets, K, out_order, scratch);
        QK_LAUNCHED();
    }
    return QK_OK;
}

extern "C" int qk_kmeans_accumulate(const float* points, int64_t point_pitch, int d, const int64_t* order,
                                    const int64_t* offsets, int64_t K, float* out_sums, int64_t out_pitch,
                                    void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(points && order && offsets && out_sums && K > 0 && d > 0, "bad argument");
    const int64_t ctas = (K + 7) / 8;  // one warp per centroid, eight per CTA
    int grid = (int)(ctas < 65535 ? ctas : 65535);
    accumulate_kernel<<<grid, 256, 0, stream>>>(points, point_pitch, d, order, offsets, K, out_sums, out_pitch);
    QK_LAUNCHED();
    return QK_OK;
}

extern "C" int qk_gather_rows(const float* src, int64_t src_pitch, const int64_t* src_ids, const int64_t* order,
                              int64_t n, int d, float* dst, int64_t dst_pitch, int64_t* dst_ids, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(src && dst && d > 0 && dst_pitch >= d, "bad argument");
    if (n == 0) return QK_OK;
    int64_t threads = n * 32;
    gather_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(src, src_pitch, src_ids, order, n, d, dst,
                                                                               dst_pitch, dst_ids);
    QK_LAUNCHED();
    return QK_OK;
}

extern "C" int qk_scatter_rows(const float* src, int64_t src_pitch, const int64_t* src_ids, const int64_t* order,
                               const int64_t* dst_rows, int64_t n, int d, float* dst, int64_t dst_pitch,
                               int64_t* dst_ids, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(src && dst && dst_rows && d > 0 && dst_pitch >= d, "bad argument");
    if (n == 0) return QK_OK;
    int64_t threads = n * 32;
    scatter_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(src, src_pitch, src_ids, order, dst_rows, n,
                                                                                d, dst, dst_pitch, dst_ids);
    QK_LAUNCHED();
    return QK_OK;
}

extern "C" int qk_normalize_rows(float* rows, int64_t n, int64_t pitch, int d, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(rows && d > 0, "bad argument");
    if (n == 0) return QK_OK;
    int64_t threads = n * 32;
    normalize_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(rows, n, pitch, d);
    QK_LAUNCHED();
    return QK_OK;
}

extern "C" int qk_row_sqnorms(const float* rows, int64_t n, int64_t pitch, int d, float* out, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(rows && out && d > 0, "bad argument");
    if (n == 0) return QK_OK;
    int64_t threads = n * 32;
    row_sqnorms_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(rows, n, pitch, d, out);
    QK_LAUNCHED();
    return QK_OK;
}

extern "C" int qk_max_row_norm(const float* rows, int64_t n, int64_t pitch, int d, float* out, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(rows && out && d > 0, "bad argument");
    if (n == 0) return QK_OK;
    int64_t threads = n * 32;
    max_row_norm_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(rows, n, pitch, d, out);
    QK_LAUNCHED();
    return QK_OK;
}

extern "C" int qk_map_ids_to_slots(const int64_t* ids, int64_t n, const int32_t* id_to_slot, int64_t table_size,
                                   int32_t* out_slots, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(ids && id_to_slot && out_slots, "bad argument");
    if (n == 0) return QK_OK;
    map_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ids, n, id_to_slot, table_size, out_slots);
    QK_LAUNCHED();
    return QK_OK;
}

extern "C" int qk_merge_topk(const float* part_distances, const int64_t* part_ids, int num_parts, int64_t Q, int k,
                             int metric, int64_t* out_ids, float* out_distances, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(part_distances && part_ids && out_ids && out_distances && num_parts > 0 && k > 0, "bad argument");
    QK_REQUIRE((int64_t)num_parts * k <= 8192, "num_parts * k = %lld exceeds 8192", (long long)num_parts * k);
    if (Q == 0) return QK_OK;
    int n = num_parts * k, np = 1;
    while (np < n) np <<= 1;
    size_t smem = (size_t)np * 8 + (size_t)n * 8;
    if (metric == QK_METRIC_INNER_PRODUCT) {
        QK_CUDA(cudaFuncSetAttribute(merge_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        merge_topk_kernel<true><<<(unsigned)Q, 256, smem, stream>>>(part_distances, part_ids, num_parts, Q, k, out_ids, out_distances);
    } else {
        QK_CUDA(cudaFuncSetAttribute(merge_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        merge_topk_kernel<false><<<(unsigned)Q, 256, smem, stream>>>(part_distances, part_ids, num_parts, Q, k, out_ids, out_distances);
    }
    QK_LAUNCHED();
    return QK_OK;
}

// ---- k-means assign on the scan path -----------------------------------------------------------
namespace {
struct AssignLayout {
    size_t off_seg_row0, off_seg_rows, off_list_seg0, off_list_nseg, off_probe, off_rows, off_dist, off_ids, off_norm,
        off_cnorms, off_stats, off_scan, scan_bytes, total;
    int nseg;
};
int64_t assign_batch(int64_t n) { return n < 65536 ? n : 65536; }
int assign_layout(int64_t n, int64_t K, int d, AssignLayout* L) {
    const int64_t B = assign_batch(n);
    L->nseg = (int)((K + QK_SEGMENT_ROWS - 1) / QK_SEGMENT_ROWS);
    qk_store_t st;
    memset(&st, 0, sizeof(st));
    st.d = d;
    st.pitch = (d + 3) / 4 * 4;
    st.num_lists = 1;
    st.num_segments = L->nseg;
    st.max_list_segments = L->nseg;
    st.num_rows = K;   // the plan (dense mode, seed sample) depends on the list's geometry: same fields as the real store
    st.flat_rows = K;
    L->scan_bytes = qk_scan_workspace_bytes(&st, B, 1, 1);
    if (L->scan_bytes == 0) return QK_ERR_INVALID_ARGUMENT;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = (o + bytes + 255) / 256 * 256; return r; };
    L->off_seg_row0 = take((size_t)L->nseg * 8);
    L->off_seg_rows = take((size_t)L->nseg * 4);
    L->off_list_seg0 = take(4);
    L->off_list_nseg = take(4);
    L->off_probe = take((size_t)B * 4);
    L->off_rows = take((size_t)B * 8);
    L->off_ids = take((size_t)B * 8);
    L->off_dist = take((size_t)B * 4);
    L->off_norm = take(4);
    L->off_cnorms = take((size_t)K * 4);
    L->off_stats = take(64);
    L->off_scan = take(L->scan_bytes);
    L->total = o;
    return QK_OK;
}
__global__ void assign_tables_kernel(int64_t K, int nseg, int64_t* seg_row0, int32_t* seg_rows, int32_t* list_seg0,
                                     int32_t* list_nseg, int32_t* probe, int64_t B, float* norm) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nseg) {
        seg_row0[i] = i * QK_SEGMENT_ROWS;
        int64_t r = K - i * QK_SEGMENT_ROWS;
        seg_rows[i] = (int32_t)(r < QK_SEGMENT_ROWS ? r : QK_SEGMENT_ROWS);
    }
    if (i == 0) { list_seg0[0] = 0; list_nseg[0] = nseg; norm[0] = 0.f; }
    if (i < B) probe[i] = 0;
}
}  // namespace

extern "C" size_t qk_kmeans_assign_workspace_bytes(int64_t n, int64_t K, int d) {
    AssignLayout L;
    if (n <= 0 || K <= 0 || d <= 0) return 0;
    if (assign_layout(n, K, d, &L) != QK_OK) return 0;
    return L.total;
}

namespace {
// out[0] += queries the batch re-scanned exactly, out[1] = max(out[1], most candidates of one query)
__global__ void assign_stats_kernel(const int32_t* __restrict__ batch_stats, int32_t* __restrict__ out) {
    out[0] += batch_stats[0];
    out[1] = max(out[1], batch_stats[1]);
}
}  // namespace

extern "C" int qk_kmeans_assign(const float* points, int64_t n, int64_t point_pitch, int d, const float* centroids,
                                int64_t K, int64_t centroid_pitch, int metric, int32_t* out_assign,
                                float* out_distances, void* workspace, size_t workspace_bytes, void* stream_v) {
    return qk_kmeans_assign_filtered(points, n, point_pitch, d, centroids, K, centroid_pitch, metric, 3, out_assign,
                                     out_distances, nullptr, workspace, workspace_bytes, stream_v);
}

extern "C" int qk_kmeans_assign_filtered(const float* points, int64_t n, int64_t point_pitch, int d, const float* centroids,
                                         int64_t K, int64_t centroid_pitch, int metric, int filter_terms, int32_t* out_assign,
                                         float* out_distances, int32_t* out_stats, void* workspace, size_t workspace_bytes,
                                         void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(filter_terms == 2 || filter_terms == 3, "filter_terms must be 2 or 3");
    QK_REQUIRE(points && centroids && out_assign && n > 0 && K > 0 && d > 0, "bad argument");
    AssignLayout L;
    int rc = assign_layout(n, K, d, &L);
    if (rc) return rc;
    if (!workspace || workspace_bytes < L.total) {
        set_error("workspace too small: need %zu bytes, have %zu", L.total, workspace_bytes);
        return QK_ERR_WORKSPACE;
    }
    char* ws = (char*)workspace;
    const int64_t B = assign_batch(n);
    int64_t* seg_row0 = (int64_t*)(ws + L.off_seg_row0);
    int32_t* seg_rows = (int32_t*)(ws + L.off_seg_rows);
    int32_t* list_seg0 = (int32_t*)(ws + L.off_list_seg0);
    int32_t* list_nseg = (int32_t*)(ws + L.off_list_nseg);
    int32_t* probe = (int32_t*)(ws + L.off_probe);
    int64_t* rows = (int64_t*)(ws + L.off_rows);
    int64_t* ids = (int64_t*)(ws + L.off_ids);
    float* dist = (float*)(ws + L.off_dist);
    float* norm = (float*)(ws + L.off_norm);
    {
        int64_t m = B > L.nseg ? B : L.nseg;
        assign_tables_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(K, L.nseg, seg_row0, seg_rows, list_seg0,
                                                                               list_nseg, probe, B, norm);
        QK_LAUNCHED();
    }
    rc = qk_max_row_norm(centroids, K, centroid_pitch, d, norm, stream);
    if (rc) return rc;
    float* cnorms = (float*)(ws + L.off_cnorms);
    rc = qk_row_sqnorms(centroids, K, centroid_pitch, d, cnorms, stream);
    if (rc) return rc;
    qk_store_t st;
    memset(&st, 0, sizeof(st));
    st.vectors = centroids;
    st.ids = nullptr;
    st.pitch = centroid_pitch;
    st.d = d;
    st.num_lists = 1;
    st.list_seg0 = list_seg0;
    st.list_nseg = list_nseg;
    st.num_segments = L.nseg;
    st.max_list_segments = L.nseg;
    st.seg_row0 = seg_row0;
    st.seg_rows = seg_rows;
    st.max_row_norm = 0.f;  // the bound stays on the device (ScanExtras::max_row_norm_dev): no host synchronisation
    st.filter_terms = filter_terms;
    st.row_norms = cnorms;
    st.num_rows = K;
    st.flat_row0 = 0;
    st.flat_rows = K;
    st.max_segment_rows = 0;
    for (int64_t b = 0; b < n; b += B) {
        const int64_t cnt = (n - b) < B ? (n - b) : B;
        ScanExtras ex;
        ex.rank_squared = 1;
        ex.max_row_norm_dev = norm;
        // flat mode (no probe table): every point scans the whole centroid list
        int32_t* batch_stats = out_stats ? (int32_t*)(ws + L.off_stats) : nullptr;
        rc = scan_partitions_impl(&st, points + b * point_pitch, cnt, point_pitch, nullptr, 1, metric, 1, ids, dist, rows,
                                  ws + L.off_scan, L.scan_bytes, batch_stats, stream, ex);
        if (rc) return rc;
        if (out_stats) {
            assign_stats_kernel<<<1, 1, 0, stream>>>(batch_stats, out_stats);
            QK_LAUNCHED();
        }
        assign_finish_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, stream>>>(rows, dist, cnt, out_assign + b,
                                                                                out_distances ? out_distances + b : nullptr);
        QK_LAUNCHED();
    }
    return QK_OK;
}
