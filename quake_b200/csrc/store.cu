// Device-resident bookkeeping of the dynamic partition store: id -> arena row hash table, removal by id.
//
// Replaces, on the GPU, the reference's per-list linear searches and the std::set-driven removal loop
// (/root/reference/src/cpp/src/dynamic_inverted_list.cpp:137-149, 302-321; src/index_partition.cpp:79-98, 129-145;
// src/partition_manager.cpp:264-320). The reference removes an id by scanning EVERY list (O(ntotal) per call); here
//   * an open-addressing hash table (64-bit id -> arena row, linear probing, tombstones) answers "where does this id
//     live" in O(1) per id,
//   * qk_store_remove erases the ids and flags their rows, and
//   * qk_store_compact_lists, one CTA per list, replays the reference's swap-with-last loop for all flagged rows of
//     the list at once: with n' = n - (rows removed), the holes below n' are filled, in ascending order, by the
//     surviving rows at or above n' in DESCENDING order -- exactly the content the sequential loop leaves behind
//     (dynamic_inverted_list.cpp:127-133: position i takes the current last element and is examined again).
#include "common.cuh"

namespace qk {

static constexpr long long HASH_EMPTY = (long long)0x8000000000000000ull;      // INT64_MIN
static constexpr long long HASH_TOMB = (long long)0x8000000000000001ull;       // INT64_MIN + 1

__device__ __forceinline__ uint64_t hash_mix(uint64_t x) {  // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

__global__ void hash_fill_kernel(long long* __restrict__ keys, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = HASH_EMPTY;
}

// insert or overwrite ids[i] -> rows[i]. The whole probe chain (up to the first EMPTY slot) is searched for the id
// before a free slot (the first tombstone met, else that EMPTY slot) is claimed, so an id never sits in two slots.
__global__ void hash_insert_kernel(long long* __restrict__ keys, long long* __restrict__ vals, uint64_t mask,
                                   const int64_t* __restrict__ ids, const int64_t* __restrict__ rows, int64_t n,
                                   int* __restrict__ failed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long id = ids[i];
    const long long row = rows ? rows[i] : i;
    if (row < 0) return;
    for (int attempt = 0; attempt < 64; ++attempt) {
        uint64_t h = hash_mix((uint64_t)id) & mask;
        long long free_slot = -1, free_val = 0;
        for (uint64_t probe = 0; probe <= mask; ++probe, h = (h + 1) & mask) {
            const long long cur = *reinterpret_cast<volatile long long*>(&keys[h]);
            if (cur == id) { vals[h] = row; return; }
            if (cur == HASH_TOMB && free_slot < 0) { free_slot = (long long)h; free_val = cur; }
            if (cur == HASH_EMPTY) {
                if (free_slot < 0) { free_slot = (long long)h; free_val = cur; }
                break;
            }
        }
        if (free_slot < 0) break;
        const long long old = (long long)atomicCAS((unsigned long long*)&keys[free_slot], (unsigned long long)free_val,
                                                   (unsigned long long)id);
        if (old == free_val || old == id) { vals[free_slot] = row; return; }
        // somebody else claimed the slot: search again
    }
    atomicExch(failed, 1);  // table full (or hopelessly contended)
}

__device__ __forceinline__ long long hash_find_slot(const long long* __restrict__ keys, uint64_t mask, long long id) {
    uint64_t h = hash_mix((uint64_t)id) & mask;
    for (uint64_t probe = 0; probe <= mask; ++probe, h = (h + 1) & mask) {
        const long long cur = keys[h];
        if (cur == id) return (long long)h;
        if (cur == HASH_EMPTY) return -1;
    }
    return -1;
}

__global__ void hash_lookup_kernel(const long long* __restrict__ keys, const long long* __restrict__ vals, uint64_t mask,
                                   const int64_t* __restrict__ ids, int64_t n, int64_t* __restrict__ out_rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long s = hash_find_slot(keys, mask, ids[i]);
    out_rows[i] = s < 0 ? -1 : vals[s];
}

// erase ids[i]: out_rows[i] = its row, or -1 if absent (or erased by a duplicate of the same id in this batch: the
// tombstone CAS succeeds once). The row is flagged for the list compaction.
__global__ void hash_erase_kernel(long long* __restrict__ keys, const long long* __restrict__ vals, uint64_t mask,
                                  const int64_t* __restrict__ ids, int64_t n, int64_t* __restrict__ out_rows,
                                  uint8_t* __restrict__ row_flags, unsigned long long* __restrict__ n_erased) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long id = ids[i];
    long long row = -1;
    const long long s = hash_find_slot(keys, mask, id);
    if (s >= 0) {
        const long long r = vals[s];
        const long long old = (long long)atomicCAS((unsigned long long*)&keys[s], (unsigned long long)id, (unsigned long long)HASH_TOMB);
        if (old == id) {
            row = r;
            row_flags[r] = 1;
            atomicAdd(n_erased, 1ull);
        }
    }
    if (out_rows) out_rows[i] = row;
}

// ---- per-list compaction after a removal --------------------------------------------------------------
struct CompactArgs {
    float* vectors;
    int64_t pitch;
    int64_t* ids;
    float* norms;
    const int64_t* list_row0;   // [num_lists]
    const int64_t* list_size;   // [num_lists] sizes BEFORE the removal
    int64_t* new_size;          // [num_lists] sizes after it (written for every list)
    uint8_t* row_flags;         // [arena rows] 1 = removed; cleared here
    int32_t* scratch_holes;     // [arena rows]
    int32_t* scratch_surv;      // [arena rows]
    long long* hkeys;
    long long* hvals;
    uint64_t hmask;
};

__global__ void __launch_bounds__(256) compact_lists_kernel(const CompactArgs a) {
    __shared__ int s_scan[256];
    __shared__ int s_total, s_base;
    const int tid = threadIdx.x;
    const int64_t l = blockIdx.x;
    const int64_t r0 = a.list_row0[l];
    const int n = (int)a.list_size[l];
    // rows removed from this list
    int cnt = 0;
    for (int p = tid; p < n; p += 256) cnt += a.row_flags[r0 + p] ? 1 : 0;
    if (tid == 0) s_total = 0;
    __syncthreads();
    if (cnt) atomicAdd(&s_total, cnt);
    __syncthreads();
    const int m = s_total;
    if (tid == 0) a.new_size[l] = n - m;
    if (m == 0) return;
    const int n2 = n - m;
    auto block_excl_scan = [&](int v) {  // exclusive prefix of v over the block; returns (prefix, block total in s_total)
        s_scan[tid] = v;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const int t = tid >= o ? s_scan[tid - o] : 0;
            __syncthreads();
            s_scan[tid] += t;
            __syncthreads();
        }
        const int incl = s_scan[tid];
        __syncthreads();
        return incl - v;
    };
    // holes below n2, ascending
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n2; c0 += 256) {
        const int p = c0 + tid;
        const int f = (p < n2 && a.row_flags[r0 + p]) ? 1 : 0;
        const int ex = block_excl_scan(f);
        const int base = s_base;
        if (f) a.scratch_holes[r0 + base + ex] = p;
        __syncthreads();
        if (tid == 255) s_base = base + ex + f;
        __syncthreads();
    }
    const int h = s_base;  // holes to fill == survivors at or above n2
    __syncthreads();       // everybody has read h before the counter is reused
    // survivors at or above n2, descending
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < m; c0 += 256) {  // positions n-1-c0-tid, down to n2
        const int p = n - 1 - (c0 + tid);
        const int f = (p >= n2 && !a.row_flags[r0 + p]) ? 1 : 0;
        const int ex = block_excl_scan(f);
        const int base = s_base;
        if (f) a.scratch_surv[r0 + base + ex] = p;
        __syncthreads();
        if (tid == 255) s_base = base + ex + f;
        __syncthreads();
    }
    __syncthreads();
    // clear the flags of this list, then move: one warp per (hole, survivor) pair
    for (int p = tid; p < n; p += 256) a.row_flags[r0 + p] = 0;
    const int lane = tid & 31, warp = tid >> 5;
    for (int i = warp; i < h; i += 8) {
        const int64_t dst = r0 + a.scratch_holes[r0 + i], src = r0 + a.scratch_surv[r0 + i];
        for (int j = lane; j < a.pitch; j += 32) a.vectors[dst * a.pitch + j] = a.vectors[src * a.pitch + j];
        if (lane == 0) {
            const int64_t id = a.ids[src];
            a.ids[dst] = id;
            a.norms[dst] = a.norms[src];
            if (a.hkeys) {
                const long long s = hash_find_slot(a.hkeys, a.hmask, id);
                if (s >= 0) a.hvals[s] = dst;
            }
        }
    }
}

}  // namespace qk

using namespace qk;

extern "C" int64_t qk_hash_capacity(int64_t n) {
    int64_t c = 1024;
    while (c < 2 * n + 16) c <<= 1;  // load factor <= 0.5
    return c;
}

extern "C" int qk_hash_clear(int64_t* keys, int64_t capacity, void* stream_v) {
    QK_REQUIRE(keys && capacity > 0 && (capacity & (capacity - 1)) == 0, "capacity must be a power of two");
    cudaStream_t stream = (cudaStream_t)stream_v;
    hash_fill_kernel<<<(unsigned)((capacity + 255) / 256), 256, 0, stream>>>((long long*)keys, capacity);
    QK_LAUNCHED();
    return QK_OK;
}

extern "C" int qk_hash_insert(int64_t* keys, int64_t* vals, int64_t capacity, const int64_t* ids, const int64_t* rows,
                              int64_t n, int32_t* failed_flag, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(keys && vals && ids && failed_flag && capacity > 0 && (capacity & (capacity - 1)) == 0, "bad argument");
    if (n <= 0) return QK_OK;
    hash_insert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((long long*)keys, (long long*)vals, (uint64_t)capacity - 1,
                                                                        ids, rows, n, failed_flag);
    QK_LAUNCHED();
    return QK_OK;
}

extern "C" int qk_hash_lookup(const int64_t* keys, const int64_t* vals, int64_t capacity, const int64_t* ids, int64_t n,
                              int64_t* out_rows, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(keys && vals && ids && out_rows && capacity > 0, "bad argument");
    if (n <= 0) return QK_OK;
    hash_lookup_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const long long*)keys, (const long long*)vals,
                                                                        (uint64_t)capacity - 1, ids, n, out_rows);
    QK_LAUNCHED();
    return QK_OK;
}

extern "C" int qk_store_remove(int64_t* keys, const int64_t* vals, int64_t capacity, const int64_t* ids, int64_t n,
                               int64_t* out_rows, uint8_t* row_flags, unsigned long long* n_erased, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(keys && vals && ids && row_flags && n_erased && capacity > 0, "bad argument");
    if (n <= 0) return QK_OK;
    hash_erase_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((long long*)keys, (const long long*)vals,
                                                                       (uint64_t)capacity - 1, ids, n, out_rows, row_flags, n_erased);
    QK_LAUNCHED();
    return QK_OK;
}

extern "C" int qk_store_compact_lists(float* vectors, int64_t pitch, int64_t* ids, float* norms,
                                      const int64_t* list_row0, const int64_t* list_size, int64_t num_lists,
                                      int64_t* new_size, uint8_t* row_flags, int32_t* scratch_holes, int32_t* scratch_surv,
                                      int64_t* hash_keys, int64_t* hash_vals, int64_t hash_capacity, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(vectors && ids && norms && list_row0 && list_size && new_size && row_flags && scratch_holes && scratch_surv,
               "null argument");
    if (num_lists <= 0) return QK_OK;
    CompactArgs a;
    a.vectors = vectors; a.pitch = pitch; a.ids = ids; a.norms = norms;
    a.list_row0 = list_row0; a.list_size = list_size; a.new_size = new_size; a.row_flags = row_flags;
    a.scratch_holes = scratch_holes; a.scratch_surv = scratch_surv;
    a.hkeys = (long long*)hash_keys; a.hvals = (long long*)hash_vals;
    a.hmask = hash_capacity > 0 ? (uint64_t)hash_capacity - 1 : 0;
    compact_lists_kernel<<<(unsigned)num_lists, 256, 0, stream>>>(a);
    QK_LAUNCHED();
    return QK_OK;
}
