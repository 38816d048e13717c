// Per-query candidate selection, exact refine, proof and (rare) exact re-scan -- included by scan.cu.
//
// Second half of the filter -> refine split: one CTA per query
//   1. candidates: the survivors the filter kernel appended to the query's buffer (IVF stores), or -- dense mode,
//      the coarse centroid scan -- the query's row of the [Q x rows] filter-key matrix; a block-wide radix select
//      keeps the kc best by filter key (plus ties)
//   2. exact refine of those in the REFERENCE'S summation order (common.cuh: ref_pair_distance_g8), sorted by
//      (distance, id): TopkBuffer's order (/root/reference/src/cpp/include/list_scanning.h:151-203) with ties
//      resolved by ascending id
//   3. proof, from the filter's rounding-error bound, that no rejected row can enter the top-k
//   4. if the proof fails (ties at the boundary, duplicates, buffer overflow): the same CTA re-scans the query's
//      probed rows exhaustively in exact arithmetic (the reference's own loop, list_scanning.h:241-311)
// Steps 1-4 used to be four launches (dense_select, merge_refine, exact_rescan + the flags hand-off); they are one.
#pragma once

namespace qk {

// 128 threads per query: with ~32 KB of shared memory seven CTAs share an SM, so a batch of 1024 queries is ONE wave of
// 148 x 7 slots (with 256 threads it was 1.4 waves of five: the second wave ran at a third of the machine). Every loop
// below strides by MERGE_THREADS / blockDim.x; nothing assumes a particular count beyond "a multiple of 32, <= 256".
#ifndef QK_MERGE_THREADS
#define QK_MERGE_THREADS 128
#endif
static constexpr int MERGE_THREADS = QK_MERGE_THREADS;
static constexpr int MERGE_SORT_CAP = 4096;   // survivors gathered per query before the select (more => exact re-scan)
static constexpr int MERGE_RANK_SORT_MAX = MERGE_THREADS;  // candidate counts up to this are ordered by a one-pass rank sort

// ------------------------------------------------------------------------------------------------
// block-wide bitonic sort of n (power of two) uint64 keys in shared memory (large k only)
// ------------------------------------------------------------------------------------------------
template <typename Less>
__device__ void block_bitonic_sort(uint64_t* s, int n, Less less) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
                int lo = 2 * i - (i & (stride - 1));
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                uint64_t x = s[lo], y = s[hi];
                bool sw = up ? less(y, x) : less(x, y);
                if (sw) { s[lo] = y; s[hi] = x; }
            }
        }
    }
    __syncthreads();
}

__host__ __device__ __forceinline__ int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

struct MergeArgs {
    const float* vecs;
    int64_t pitch;
    const int64_t* ids;
    int d;
    const int64_t* seg_row0;
    const int32_t* seg_rows;
    const float* queries;
    int64_t q_pitch;
    const int32_t* pair_seg;   // [Q x P] segments probed by every query (-1 = none); unused when flat_nseg > 0
    int flat_nseg;             // single-list store scanned without a probe table: every query probes segments 0..flat_nseg-1
    const uint32_t* gthr;
    const uint64_t* qbuf;
    const int32_t* qcount;
    int qcap, qstride;
    int32_t* flags;
    int32_t* ctrl;
    int P, kc, k;
    float max_row_norm;
    const float* max_row_norm_dev;  // optional: the bound lives on the device (k-means assign); overrides max_row_norm
    double filter_gam;  // extra relative error bound of the filter's dot products (tensor-core path), vs |q||v|
    double probe_gam;   // > 0: also evaluate the proof with THIS filter bound and count the queries that would fail it
                        // (ctrl[6]): how the host decides whether the cheaper 2xTF32 filter would do for this data
    int64_t* out_ids;
    float* out_dist;
    int64_t* out_rows;
    int force_rescan;
    int sort_cap;      // survivors the shared-memory gather buffer holds (more => exact re-scan)
    int rank_squared;  // l2 only: order by the squared distance (k-means assign: faiss Top1 on squared l2)
    // dense mode (dense_refine_kernel)
    const uint32_t* dense;  // [Q x dense_rows] filter keys
    int dense_rows;
    long long dense_row0;
    // fused pair expansion (qk_search_ivf): the ids this kernel emits are the partitions the NEXT scan probes; rank i of
    // query q becomes pair slot (q, i) of that scan -- segment of the list (single-segment lists only), histogram, and
    // the query's threshold reset -- instead of a separate expand_pairs launch re-reading the ids. x_pair_seg == NULL: off
    int32_t* x_pair_seg;          // [Q x k]
    int32_t* x_seg_count;         // [S] zeroed by the caller
    uint32_t* x_gthr;             // [Q]
    const int32_t* x_id_to_slot;  // dense partition id -> list slot table
    int64_t x_table_size;
    const int32_t* x_list_seg0;
    const int32_t* x_list_nseg;
    int x_num_lists, x_shard_rank, x_shard_world;
    // set mode (dense_refine_kernel, qk_search_ivf's coarse scan): the caller only needs WHICH k rows are the best, not
    // their order or exact distances -- see set_select_and_emit
    int set_mode;
};

// the fused expansion of one emitted result (see MergeArgs::x_pair_seg)
__device__ __forceinline__ void emit_expand(const MergeArgs& a, int64_t q, int i, int64_t id) {
    if (!a.x_pair_seg) return;
    int seg = -1;
    if (id >= 0 && id < a.x_table_size && !(a.x_shard_world > 1 && (int)(id % a.x_shard_world) != a.x_shard_rank)) {
        const int l = a.x_id_to_slot[id];
        if (l >= 0 && l < a.x_num_lists && a.x_list_nseg[l] > 0) seg = a.x_list_seg0[l];
    }
    a.x_pair_seg[q * a.k + i] = seg;
    if (seg >= 0) atomicAdd(&a.x_seg_count[seg], 1);
    if (i == 0) a.x_gthr[q] = KEY_MAX;
}

// |q|^2 in double for the proof's bounds: warp 0, four partial sums per lane and a shuffle tree (one thread adding d
// terms in sequence was a ~0.5 us chain at d = 128 that the CTA's first barrier waited for). Any accurate value will do:
// the bound it enters carries relative slack of 2^-17 and more.
__device__ __forceinline__ void block_query_sqnorm(const float* qs, int d, double* out) {
    if (threadIdx.x < 32) {
        double t0 = 0.0, t1 = 0.0;
        int i = threadIdx.x;
        for (; i + 32 < d; i += 64) {
            t0 += (double)qs[i] * (double)qs[i];
            t1 += (double)qs[i + 32] * (double)qs[i + 32];
        }
        if (i < d) t0 += (double)qs[i] * (double)qs[i];
        double t = t0 + t1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) *out = t;
    }
}

struct RescanSmem {
    unsigned hist[256];
    int scan[256];
    unsigned prefix, need, less, eq_total;
    unsigned long long prefix64;
    int base_lt, base_eq;
};

// segment j of query q's probe list, or -1
__device__ __forceinline__ int probed_segment(const MergeArgs& a, int64_t q, int j) {
    if (a.flat_nseg > 0) return j < a.flat_nseg ? j : -1;
    return a.pair_seg[q * a.P + j];
}
__device__ __forceinline__ int probed_slots(const MergeArgs& a) { return a.flat_nseg > 0 ? a.flat_nseg : a.P; }

// ------------------------------------------------------------------------------------------------
// The kc entries with the smallest 32-bit keys among entry_at(i) = key << 32 | row, i < n (n >= kc), written to
// out[0..kc) in no particular order; *t_out = the largest selected key (every entry left out has a key >= it).
// Exactly kc entries are selected: ties at the boundary are broken by position, which is fine for a FILTER -- the
// proof downstream only needs "everything rejected has a filter key >= T".
//
// Filter keys of one query share their leading bits (scores of similar magnitude), so a plain MSB-first radix pass
// would pile every key into one or two histogram bins (serialised shared-memory atomics: most of the old 16 us
// dense_select). Two scans of the data instead of six:
//   1. histogram passes, 8 bits of v = key - kmin at a time from the top of the range [kmin, kmax] the caller measured
//      while it staged the data, per-warp histograms summed, narrowing to the bin that holds rank kc -- until that
//      boundary bin has at most SELECT_BOUNDARY_CAP entries (typically one or two passes)
//   2. one scan: entries below the boundary bin go straight to `out`, the bin's entries to a small list
//   3. the boundary list is ranked by counting; its `need` smallest complete the selection
// Returns kc.
// `whist` = (MERGE_THREADS / 32) x 256 words, `sc` = 4 words, `blist` = SELECT_BOUNDARY_CAP entries.
// ------------------------------------------------------------------------------------------------
static constexpr int SELECT_BOUNDARY_CAP = 512;
template <typename EntryAt>
__device__ int block_select_smallest(EntryAt entry_at, int n, int kc, uint32_t kmin, uint32_t kmax, uint64_t* out,
                                     uint32_t* t_out, unsigned* whist, unsigned* sc, uint64_t* blist) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = MERGE_THREADS / 32;
    const uint32_t range = kmax - kmin;
    // The current bin is [lo, lo + 2^hi) in v = key - kmin space and holds rank kc; `below` entries lie under it.
    // Order-preserving float keys are logarithmic in the score (sign and exponent on top), so one 8-bit pass may leave
    // a whole binade -- hundreds of keys -- in the boundary bin: refine it 8 bits at a time until it is small.
    int hi = 32 - __clz(range);  // significant bits of v (0: all keys equal)
    uint32_t lo = 0;
    int below = 0, nb = n;
    while (hi > 0 && nb > SELECT_BOUNDARY_CAP) {
        const int w = hi < 8 ? hi : 8, sh = hi - w;
        for (int i = tid; i < NW * 256; i += MERGE_THREADS) whist[i] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += MERGE_THREADS) {
            const uint32_t v = (uint32_t)(entry_at(i) >> 32) - kmin;
            const bool in_bin = hi >= 32 || (v >> hi) == (lo >> hi);
            if (in_bin) atomicAdd(&whist[warp * 256 + ((v >> sh) & ((1u << w) - 1u))], 1u);
        }
        __syncthreads();
        for (int b = tid; b < 256; b += MERGE_THREADS) {
            unsigned t = 0;
#pragma unroll
            for (int ww = 0; ww < NW; ++ww) t += whist[ww * 256 + b];
            whist[b] = t;  // column b is only touched by this thread
        }
        __syncthreads();
        if (tid < 32) {  // lane l owns bins 8l .. 8l+7
            uint32_t h[8], sum = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) { h[b] = whist[lane * 8 + b]; sum += h[b]; }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            const uint32_t need = (uint32_t)(kc - below);
            const unsigned owner = __ballot_sync(0xffffffffu, incl >= need);
            const int ol = __ffs(owner) - 1;
            if (lane == ol) {
                uint32_t before = incl - sum;
                int bin = 0;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    if (before + h[b] >= need) { bin = lane * 8 + b; break; }
                    before += h[b];
                }
                sc[0] = (uint32_t)bin;
                sc[1] = before;
                sc[2] = h[bin & 7];
            }
        }
        __syncthreads();
        lo |= sc[0] << sh;
        below += (int)sc[1];
        nb = (int)sc[2];
        hi = sh;
        __syncthreads();  // sc is rewritten by the next pass
    }
    // one scan: entries under the boundary bin go straight to `out`, the boundary bin's to the list
    if (tid == 0) { sc[2] = 0u; sc[3] = 0u; }
    __syncthreads();
    const int need = kc - below;  // 1 .. nb entries of the boundary bin complete the selection
    const bool all_equal = nb > SELECT_BOUNDARY_CAP;  // hi == 0: one key value, any `need` of its entries will do
    for (int i = tid; i < n; i += MERGE_THREADS) {
        const uint64_t e = entry_at(i);
        const uint32_t v = (uint32_t)(e >> 32) - kmin;
        const bool in_bin = hi >= 32 || (v >> hi) == (lo >> hi);
        if (in_bin) {
            const unsigned pos = atomicAdd(&sc[3], 1u);
            if (all_equal) {
                if (pos < (unsigned)need) out[below + pos] = e;
            } else if (pos < (unsigned)SELECT_BOUNDARY_CAP) {
                blist[pos] = e;
            }
        } else if (v < lo) {
            out[atomicAdd(&sc[2], 1u)] = e;
        }
    }
    __syncthreads();
    if (all_equal) {
        if (tid == 0) *t_out = kmin + lo;
        __syncthreads();
        return kc;
    }
    for (int i = tid; i < nb; i += MERGE_THREADS) {
        const uint64_t e = blist[i];
        const uint32_t ke = (uint32_t)(e >> 32);
        int rank = 0;
        for (int j = 0; j < nb; ++j) {
            const uint32_t kj = (uint32_t)(blist[j] >> 32);
            rank += (kj < ke || (kj == ke && j < i)) ? 1 : 0;
        }
        if (rank < need) out[below + rank] = e;
        if (rank == need - 1) *t_out = ke;
    }
    __syncthreads();
    return kc;
}

// ------------------------------------------------------------------------------------------------
// exhaustive exact re-scan of one query (radix select on exact distance keys), output written here
// ------------------------------------------------------------------------------------------------
template <bool kIP>
__device__ void exact_rescan_body(const MergeArgs& a, int64_t q, const float* qs, uint64_t* rkey, int64_t* rid,
                                  uint32_t* rrow, RescanSmem& sm) {
    const int kp = next_pow2(a.k);
    const int tid = threadIdx.x;
    const float inf_pad = kIP ? -INFINITY : INFINITY;
    const int nslots = probed_slots(a);
    int64_t total = 0;
    for (int j = 0; j < nslots; ++j) {
        const int seg = probed_segment(a, q, j);
        if (seg >= 0) total += a.seg_rows[seg];
    }
    const int kk = (int)(total < a.k ? total : a.k);
    __syncthreads();
    if (tid == 0) { sm.prefix = 0; sm.need = kk; sm.less = 0; }
    __syncthreads();

    auto dist_key = [&](int64_t row) {
        const float dist = ref_pair_distance<kIP>(qs, a.vecs + row * a.pitch, a.d);
        return f2key(kIP ? -dist : (a.rank_squared ? dist : __fsqrt_rn(dist)));
    };
    auto id_key = [&](int64_t row) {  // ascending signed id order as unsigned
        const int64_t id = a.ids ? a.ids[row] : row;
        return (uint64_t)id ^ 0x8000000000000000ull;
    };

    if (kk > 0) {
        // radix select of the kk-th smallest exact key, most significant byte first
        for (int pass = 3; pass >= 0; --pass) {
            for (int b = tid; b < 256; b += blockDim.x) sm.hist[b] = 0;
            __syncthreads();
            const unsigned prefix = sm.prefix;
            for (int j = 0; j < nslots; ++j) {
                const int seg = probed_segment(a, q, j);
                if (seg < 0) continue;
                const int64_t r0 = a.seg_row0[seg];
                const int n = a.seg_rows[seg];
                for (int r = tid; r < n; r += blockDim.x) {
                    const uint32_t dk = dist_key(r0 + r);
                    const bool match = (pass == 3) || ((dk >> (8 * (pass + 1))) == (prefix >> (8 * (pass + 1))));
                    if (match) atomicAdd(&sm.hist[(dk >> (8 * pass)) & 255u], 1u);
                }
            }
            __syncthreads();
            if (tid == 0) {
                unsigned need = sm.need, cum = 0;
                int b = 0;
                for (; b < 256; ++b) {
                    if (cum + sm.hist[b] >= need) break;
                    cum += sm.hist[b];
                }
                sm.prefix = prefix | ((unsigned)b << (8 * pass));
                sm.need = need - cum;
                sm.less += cum;
                sm.eq_total = b < 256 ? sm.hist[b] : 0u;
            }
            __syncthreads();
        }
        const uint32_t T = sm.prefix;
        const unsigned n_less = sm.less;        // keys strictly below T
        const unsigned take_eq = sm.need;       // how many keys equal to T to take
        const unsigned eq_total = sm.eq_total;  // how many keys equal T
        __syncthreads();
        // A distance tie that straddles the k-th boundary is resolved by ascending id (the order the
        // oracle fixes for the reference's distance-only comparator): radix-select the take_eq-th
        // smallest id among the rows whose key equals T.
        uint64_t id_thr = ~0ull;
        if (eq_total > take_eq) {
            if (tid == 0) { sm.prefix64 = 0ull; sm.need = take_eq; }
            __syncthreads();
            for (int pass = 7; pass >= 0; --pass) {
                for (int b = tid; b < 256; b += blockDim.x) sm.hist[b] = 0;
                __syncthreads();
                const uint64_t prefix = sm.prefix64;
                for (int j = 0; j < nslots; ++j) {
                    const int seg = probed_segment(a, q, j);
                    if (seg < 0) continue;
                    const int64_t r0 = a.seg_row0[seg];
                    const int n = a.seg_rows[seg];
                    for (int r = tid; r < n; r += blockDim.x) {
                        if (dist_key(r0 + r) != T) continue;
                        const uint64_t ik = id_key(r0 + r);
                        const bool match = (pass == 7) || ((ik >> (8 * (pass + 1))) == (prefix >> (8 * (pass + 1))));
                        if (match) atomicAdd(&sm.hist[(unsigned)(ik >> (8 * pass)) & 255u], 1u);
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    unsigned need = sm.need, cum = 0;
                    int b = 0;
                    for (; b < 256; ++b) {
                        if (cum + sm.hist[b] >= need) break;
                        cum += sm.hist[b];
                    }
                    sm.prefix64 = prefix | ((unsigned long long)(b & 255) << (8 * pass));
                    sm.need = need - cum;
                }
                __syncthreads();
            }
            id_thr = sm.prefix64;
        }
        if (tid == 0) { sm.base_lt = 0; sm.base_eq = 0; }
        __syncthreads();
        // deterministic ordered collection
        for (int j = 0; j < nslots; ++j) {
            const int seg = probed_segment(a, q, j);
            if (seg < 0) continue;
            const int64_t r0 = a.seg_row0[seg];
            const int n = a.seg_rows[seg];
            for (int base = 0; base < n; base += blockDim.x) {
                const int r = base + tid;
                uint32_t dk = KEY_MAX;
                bool lt = false, eq = false;
                if (r < n) {
                    dk = dist_key(r0 + r);
                    lt = dk < T;
                    eq = (dk == T) && (id_key(r0 + r) <= id_thr);
                }
                // block exclusive scan of (lt, eq) packed
                int v = (lt ? 1 : 0) | (eq ? (1 << 16) : 0);
                sm.scan[tid] = v;
                __syncthreads();
                for (int o = 1; o < (int)blockDim.x; o <<= 1) {
                    int t = (tid >= o) ? sm.scan[tid - o] : 0;
                    __syncthreads();
                    sm.scan[tid] += t;
                    __syncthreads();
                }
                const int incl = sm.scan[tid];
                const int excl = incl - v;
                const int blt = sm.base_lt, beq = sm.base_eq;
                int slot = -1;
                if (lt) slot = blt + (excl & 0xffff);
                else if (eq) {
                    const int e = beq + (excl >> 16);
                    if (e < (int)take_eq) slot = (int)n_less + e;
                }
                if (slot >= 0 && slot < kp) {
                    const int64_t row = r0 + r;
                    rkey[slot] = ((uint64_t)dk << 32) | (uint32_t)slot;
                    rid[slot] = a.ids ? a.ids[row] : row;
                    rrow[slot] = (uint32_t)row;
                }
                __syncthreads();
                if (tid == (int)blockDim.x - 1) {
                    sm.base_lt = blt + (incl & 0xffff);
                    sm.base_eq = beq + (incl >> 16);
                }
                __syncthreads();
            }
        }
    }
    __syncthreads();
    for (int i = kk + tid; i < kp; i += blockDim.x) rkey[i] = COMP_MAX;
    const int64_t* ridc = rid;
    block_bitonic_sort(rkey, kp, [ridc](uint64_t x, uint64_t y) {
        const uint32_t dx = (uint32_t)(x >> 32), dy = (uint32_t)(y >> 32);
        if (dx != dy) return dx < dy;
        if (x == COMP_MAX || y == COMP_MAX) return x < y;
        return ridc[(uint32_t)x] < ridc[(uint32_t)y];
    });
    for (int i = tid; i < a.k; i += blockDim.x) {
        int64_t id = -1, row = -1;
        float dist = inf_pad;
        if (i < kk) {
            const uint64_t rk = rkey[i];
            const uint32_t slot = (uint32_t)rk;
            const float v = key2f((uint32_t)(rk >> 32));
            dist = kIP ? -v : (a.rank_squared ? __fsqrt_rn(v) : v);
            id = rid[slot];
            row = rrow[slot];
        }
        a.out_ids[q * a.k + i] = id;
        a.out_dist[q * a.k + i] = dist;
        if (a.out_rows) a.out_rows[q * a.k + i] = row;
        emit_expand(a, q, i, id);
    }
}

// ------------------------------------------------------------------------------------------------
// refine `m` candidates (composites key << 32 | arena row in cand[0..m), any order, m <= kcp), order them by
// (exact distance, id), prove that the rejected rows (all of which have filter key > a_key) cannot enter the
// top-k, emit -- or fall through to the exact re-scan.
// ------------------------------------------------------------------------------------------------
struct RefineSmem {
    float* qs;        // [d padded to 4]
    uint64_t* rkey;   // [kcp]  exact distance key << 32 | candidate slot
    int64_t* rid;     // [kcp]
    uint32_t* rrow;   // [kcp]
    uint32_t* order;  // [kcp]  candidate slot at every output rank
    int* rank;        // [kcp]
};

template <bool kIP>
__device__ void refine_and_emit(const MergeArgs& a, int64_t q, const uint64_t* cand, int m, uint32_t a_key,
                                bool have_rejects, bool rescan, const RefineSmem& s, double qn2, RescanSmem& rs,
                                int* s_flag) {
    const int tid = threadIdx.x;
    const int kcp = next_pow2(a.kc);
    const float inf_pad = kIP ? -INFINITY : INFINITY;
    int nc = 0;
    if (!rescan) {
        nc = m;
        // ---- exact refine in the reference's summation order: eight lanes per candidate, one per accumulator of
        //      the reference's 8-wide loop
        for (int base = 0; base < nc; base += MERGE_THREADS / 8) {
            const int i = base + (tid >> 3), j = tid & 7;
            if (i < nc) {  // uniform inside every 8-lane group
                const uint32_t row = (uint32_t)cand[i];
                const float dist = ref_pair_distance_g8<kIP>(s.qs, a.vecs + (int64_t)row * a.pitch, a.d, j);
                if (j == 0) {
                    // order by the value the reference orders by: sqrt'ed for l2 (list_scanning.h:260)
                    const uint32_t dk = f2key(kIP ? -dist : (a.rank_squared ? dist : __fsqrt_rn(dist)));
                    s.rkey[i] = ((uint64_t)dk << 32) | (uint32_t)i;
                    s.rid[i] = a.ids ? a.ids[row] : (int64_t)row;
                    s.rrow[i] = row;
                }
            }
        }
        __syncthreads();
        if (kcp <= MERGE_RANK_SORT_MAX) {
            // one-pass rank sort: rank(i) = number of candidates ordered before i; P threads share one candidate
            const int P = MERGE_THREADS / kcp;  // kcp is a power of two <= MERGE_THREADS
            for (int i = tid; i < kcp; i += MERGE_THREADS) s.rank[i] = 0;
            __syncthreads();
            const int i = tid / P, part = tid - i * P;
            if (i < nc) {
                const uint64_t ki = s.rkey[i];
                const uint32_t di = (uint32_t)(ki >> 32);
                const int64_t idi = s.rid[i];
                const int span = (nc + P - 1) / P;
                const int j0 = part * span, j1 = min(nc, j0 + span);
                int cnt = 0;
                for (int j = j0; j < j1; ++j) {
                    const uint32_t dj = (uint32_t)(s.rkey[j] >> 32);
                    const int64_t idj = s.rid[j];
                    const bool before = dj < di || (dj == di && (idj < idi || (idj == idi && j < i)));
                    cnt += before ? 1 : 0;
                }
                if (cnt) atomicAdd(&s.rank[i], cnt);
            }
            __syncthreads();
            if (tid < nc) s.order[s.rank[tid]] = (uint32_t)tid;
            __syncthreads();
        } else {
            for (int i = nc + tid; i < kcp; i += MERGE_THREADS) s.rkey[i] = COMP_MAX;
            const int64_t* ridc = s.rid;
            block_bitonic_sort(s.rkey, kcp, [ridc](uint64_t x, uint64_t y) {
                const uint32_t dx = (uint32_t)(x >> 32), dy = (uint32_t)(y >> 32);
                if (dx != dy) return dx < dy;
                if (x == COMP_MAX || y == COMP_MAX) return x < y;
                return ridc[(uint32_t)x] < ridc[(uint32_t)y];
            });
            // after the in-place sort rkey[r] carries the slot of rank r in its low word; look the key up through it
            for (int i = tid; i < nc; i += MERGE_THREADS) s.order[i] = (uint32_t)s.rkey[i];
            __syncthreads();
        }
        // ---- proof that nothing outside the refined set can enter the top-k
        if (have_rejects && nc >= 1) {
            if (tid == 0) {
                const int kk = a.k < nc ? a.k : nc;
                const float a_score = key2f(a_key);  // every rejected row's filter score is above this
                // exact k-th (l2: distance, ip: -ip). In the bitonic path rkey is sorted in place (rank r at r); in
                // the rank-sort path it is indexed by candidate slot.
                const uint64_t kth = kcp <= MERGE_RANK_SORT_MAX ? s.rkey[s.order[kk - 1]] : s.rkey[kk - 1];
                const float rk = key2f((uint32_t)(kth >> 32));
                const double U = (double)(a.max_row_norm_dev ? *a.max_row_norm_dev : a.max_row_norm);
                const double qn = qn2, qnorm = sqrt(qn);
                const double eps = 5.960464477539063e-08;  // 2^-24
                const double gam = (a.d + 8) * eps;
                const double e2 = (a.d / 8 + 12) * eps;
                bool ok;
                if (!kIP) {
                    const double e1 = gam * U * U + 2.0 * (gam + a.filter_gam) * qnorm * U + 4.0 * eps * fabs((double)a_score);
                    // lb bounds the reference-order SQUARED distance of every rejected row from below; the
                    // reference compares sqrt'ed values, and sqrt_rn is monotone, so a rejected row cannot
                    // tie or beat the k-th as soon as sqrt_rn(round_down(lb)) is strictly above it.
                    const double lb = (qn * (1.0 - gam) + (double)a_score - e1) * (1.0 - e2);
                    const float lbf = __double2float_rd(lb);
                    ok = a.rank_squared ? (lb > (double)rk) : (lbf > 0.f && __fsqrt_rn(lbf) > rk);
                    if (a.probe_gam > 0.0) {
                        const double e1p = gam * U * U + 2.0 * (gam + a.probe_gam) * qnorm * U + 4.0 * eps * fabs((double)a_score);
                        const double lbp = (qn * (1.0 - gam) + (double)a_score - e1p) * (1.0 - e2);
                        const float lbpf = __double2float_rd(lbp);
                        const bool okp = a.rank_squared ? (lbp > (double)rk) : (lbpf > 0.f && __fsqrt_rn(lbpf) > rk);
                        if (!okp) atomicAdd(&a.ctrl[6], 1);
                    }
                } else {
                    // scores are -<q,v>: any rejected v has ip <= -a_score + err; need that below the k-th exact ip
                    const double err = (gam + a.filter_gam + e2) * qnorm * U + 4.0 * eps * fabs((double)a_score);
                    ok = (-(double)a_score + err) < -(double)rk;
                    if (a.probe_gam > 0.0) {
                        const double errp = (gam + a.probe_gam + e2) * qnorm * U + 4.0 * eps * fabs((double)a_score);
                        if (!((-(double)a_score + errp) < -(double)rk)) atomicAdd(&a.ctrl[6], 1);
                    }
                }
                *s_flag = ok ? 0 : -1;
            }
            __syncthreads();
            rescan = (*s_flag < 0);
        }
    }
    if (rescan) {
        if (tid == 0) {
            a.flags[q] = 1;
            atomicAdd(&a.ctrl[2], 1);
        }
        exact_rescan_body<kIP>(a, q, s.qs, s.rkey, s.rid, s.rrow, rs);
        return;
    }
    if (tid == 0) a.flags[q] = 0;
    for (int i = tid; i < a.k; i += blockDim.x) {
        int64_t id = -1, row = -1;
        float dist = inf_pad;
        if (i < nc) {
            const uint32_t slot = s.order[i];
            const float v = key2f((uint32_t)(s.rkey[kcp <= MERGE_RANK_SORT_MAX ? slot : i] >> 32));
            dist = kIP ? -v : (a.rank_squared ? __fsqrt_rn(v) : v);
            id = s.rid[slot];
            row = s.rrow[slot];
        }
        a.out_ids[q * a.k + i] = id;
        a.out_dist[q * a.k + i] = dist;
        if (a.out_rows) a.out_rows[q * a.k + i] = row;
        emit_expand(a, q, i, id);
    }
}

// ------------------------------------------------------------------------------------------------
// Set mode: the k best rows as a SET (fixed-nprobe coarse scan: the partition scan that follows does not care in which
// order a query's lists were probed, query_coordinator.cpp:628-645 only hands the ids on). Membership is decided from
// the filter scores wherever the filter's error bound allows it and by exact arithmetic only where it does not:
//   * the m > k candidates (everything else has a filter key >= a_key) are ranked by filter key; s_lo / s_hi = the
//     k-th / (k+1)-th smallest score
//   * [lo(s), hi(s)] bounds the value the REFERENCE orders by (sqrt'ed l2 distance in its summation order, or -ip) of
//     any row whose filter score is s -- the same error model as the proof in refine_and_emit, used in both directions
//   * a candidate with hi(s) < lo(s_hi) beats every row ranked k+1 or worse: a sure member; one with lo(s) > hi(s_lo)
//     loses to k rows: surely out; the rows outside the candidate set must all be surely out (else: return false and
//     the ordinary path runs). What remains -- typically the two or three candidates around the boundary, none at all
//     for most queries -- is refined exactly and ordered by (distance, id) like the ordinary path does
// Emits exactly k ids, best filter score first (the seed kernel samples a query's first list), with distances derived
// from the filter score (approximate; nobody reads them). Saves the exact evaluation of ~k candidates and the
// (distance, id) rank sort: a third of dense_refine_kernel's instructions at nprobe 64.
// Returns false (nothing emitted) when the bounds do not settle it; all threads take the same decision.
// ------------------------------------------------------------------------------------------------
template <bool kIP>
__device__ __forceinline__ void exact_value_bounds(float sc, double qn, double qnorm, double U, double fgam, int d, bool squared,
                                                   float* lo, float* hi) {
    const double eps = 5.960464477539063e-08;  // 2^-24
    const double gam = (d + 8) * eps;
    const double e2 = (d / 8 + 12) * eps;
    if (!kIP) {
        const double e1 = gam * U * U + 2.0 * (gam + fgam) * qnorm * U + 4.0 * eps * fabs((double)sc);
        const float lbf = __double2float_rd((qn * (1.0 - gam) + (double)sc - e1) * (1.0 - e2));
        const float ubf = __double2float_ru((qn * (1.0 + gam) + (double)sc + e1) * (1.0 + e2));
        *lo = lbf > 0.f ? (squared ? lbf : __fsqrt_rn(lbf)) : 0.f;
        *hi = ubf > 0.f ? (squared ? ubf : __fsqrt_rn(ubf)) : 0.f;
        if (!(ubf == ubf) || !(lbf == lbf)) { *lo = 0.f; *hi = INFINITY; }  // NaN: nothing is settled
    } else {
        const double err = (gam + fgam + e2) * qnorm * U + 4.0 * eps * fabs((double)sc);
        *lo = __double2float_rd((double)sc - err);
        *hi = __double2float_ru((double)sc + err);
        if (!(*lo == *lo) || !(*hi == *hi)) { *lo = -INFINITY; *hi = INFINITY; }
    }
}

template <bool kIP>
__device__ bool set_select_and_emit(const MergeArgs& a, int64_t q, const uint64_t* cand, int m, uint32_t a_key, bool have_rejects,
                                    const RefineSmem& s, double qn2) {
    __shared__ int s_nm, s_na;
    __shared__ int s_wsum[MERGE_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = a.k;
    // 1. rank by filter key (ties by position): order[r] = candidate at rank r
    for (int i = tid; i < m; i += MERGE_THREADS) {
        const uint32_t ki = (uint32_t)(cand[i] >> 32);
        int r = 0;
        for (int j = 0; j < m; ++j) {
            const uint32_t kj = (uint32_t)(cand[j] >> 32);
            r += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
        }
        s.order[r] = (uint32_t)i;
    }
    if (tid == 0) { s_nm = 0; s_na = 0; }
    __syncthreads();
    // 2. classify (every thread evaluates the same boundary bounds: the decision is uniform without a broadcast)
    const double U = (double)(a.max_row_norm_dev ? *a.max_row_norm_dev : a.max_row_norm);
    const double qnorm = sqrt(qn2);
    const bool sq = a.rank_squared != 0;
    float lo_klo, hi_klo, lo_khi, hi_khi;
    exact_value_bounds<kIP>(key2f((uint32_t)(cand[s.order[k - 1]] >> 32)), qn2, qnorm, U, a.filter_gam, a.d, sq, &lo_klo, &hi_klo);
    exact_value_bounds<kIP>(key2f((uint32_t)(cand[s.order[k]] >> 32)), qn2, qnorm, U, a.filter_gam, a.d, sq, &lo_khi, &hi_khi);
    if (have_rejects) {
        float lo_r, hi_r;
        exact_value_bounds<kIP>(key2f(a_key), qn2, qnorm, U, a.filter_gam, a.d, sq, &lo_r, &hi_r);
        if (!(lo_r > hi_klo)) return false;  // a row outside the candidate set might belong to the k best
    }
    for (int r = tid; r < m; r += MERGE_THREADS) {
        float lo_i, hi_i;
        exact_value_bounds<kIP>(key2f((uint32_t)(cand[s.order[r]] >> 32)), qn2, qnorm, U, a.filter_gam, a.d, sq, &lo_i, &hi_i);
        int cls = 2;                       // undecided
        if (hi_i < lo_khi) cls = 1;        // beats every row ranked k+1 or worse
        else if (lo_i > hi_klo) cls = 0;   // loses to the k rows ranked k or better
        if (cls == 1) atomicAdd(&s_nm, 1);
        else if (cls == 2) s.rrow[atomicAdd(&s_na, 1)] = (uint32_t)r;
        s.rank[r] = cls;
    }
    __syncthreads();
    const int na = s_na, need = k - s_nm;
    if (need < 0 || need > na) return false;  // cannot happen with monotone bounds; the ordinary path sorts it out
    // 3. the undecided candidates: exact values, (distance, id) order, the `need` best are members
    if (need > 0 && need < na) {
        for (int base = 0; base < na; base += MERGE_THREADS / 8) {
            const int x = base + (tid >> 3), j = tid & 7;
            if (x < na) {  // uniform inside every 8-lane group
                const uint32_t row = (uint32_t)cand[s.order[s.rrow[x]]];
                const float dist = ref_pair_distance_g8<kIP>(s.qs, a.vecs + (int64_t)row * a.pitch, a.d, j);
                if (j == 0) {
                    s.rkey[x] = (uint64_t)f2key(kIP ? -dist : (sq ? dist : __fsqrt_rn(dist))) << 32;
                    s.rid[x] = a.ids ? a.ids[row] : (int64_t)row;
                }
            }
        }
        __syncthreads();
        for (int x = tid; x < na; x += MERGE_THREADS) {
            const uint32_t dx = (uint32_t)(s.rkey[x] >> 32);
            const int64_t idx = s.rid[x];
            int before = 0;
            for (int y = 0; y < na; ++y) {
                const uint32_t dy = (uint32_t)(s.rkey[y] >> 32);
                const int64_t idy = s.rid[y];
                before += (dy < dx || (dy == dx && (idy < idx || (idy == idx && y < x)))) ? 1 : 0;
            }
            s.rank[s.rrow[x]] = before < need ? 1 : 0;
        }
    } else {
        for (int x = tid; x < na; x += MERGE_THREADS) s.rank[s.rrow[x]] = need > 0 ? 1 : 0;
    }
    __syncthreads();
    // 4. emit the members in filter-rank order
    int carry = 0;
    for (int base = 0; base < m; base += MERGE_THREADS) {
        const int r = base + tid;
        const bool member = r < m && s.rank[r] == 1;
        const unsigned bal = __ballot_sync(0xffffffffu, member);
        if (lane == 0) s_wsum[warp] = __popc(bal);
        __syncthreads();
        int off = carry, tot = 0;
#pragma unroll
        for (int w = 0; w < MERGE_THREADS / 32; ++w) {
            const int c = s_wsum[w];
            if (w < warp) off += c;
            tot += c;
        }
        if (member) {
            const int slot = off + __popc(bal & ((1u << lane) - 1u));
            const uint64_t e = cand[s.order[r]];
            const uint32_t row = (uint32_t)e;
            const float sc = key2f((uint32_t)(e >> 32));
            const int64_t id = a.ids ? a.ids[row] : (int64_t)row;
            const float d2 = fmaxf((float)qn2 + sc, 0.f);
            a.out_ids[q * k + slot] = id;
            a.out_dist[q * k + slot] = kIP ? -sc : (sq ? d2 : sqrtf(d2));
            if (a.out_rows) a.out_rows[q * k + slot] = row;
            emit_expand(a, q, slot, id);
        }
        carry += tot;
        __syncthreads();
    }
    if (tid == 0) a.flags[q] = 0;
    return true;
}

// shared-memory carve-up common to both kernels: [front buffer][cbuf kcp u64][qs][rkey][rid][rrow][order][rank]
__host__ __device__ inline size_t refine_tail_bytes(int d, int kcp) {
    return (size_t)kcp * 8 + (size_t)((d + 3) & ~3) * 4 + 8 + (size_t)kcp * (8 + 8 + 4 + 4 + 4);
}
__device__ __forceinline__ uint64_t* carve_refine(unsigned char* p, int d, int kcp, RefineSmem& s) {
    uint64_t* cbuf = reinterpret_cast<uint64_t*>(p);
    s.rkey = cbuf + kcp;
    s.rid = reinterpret_cast<int64_t*>(s.rkey + kcp);
    s.rrow = reinterpret_cast<uint32_t*>(s.rid + kcp);
    s.order = s.rrow + kcp;
    s.rank = reinterpret_cast<int*>(s.order + kcp);
    s.qs = reinterpret_cast<float*>(s.rank + kcp);
    return cbuf;
}

// ------------------------------------------------------------------------------------------------
// IVF stores: candidates from the per-query buffers the filter kernel appended to
// ------------------------------------------------------------------------------------------------
template <bool kIP>
__global__ void __launch_bounds__(MERGE_THREADS, MERGE_THREADS == 128 ? 8 : 5) merge_refine_kernel(const MergeArgs a) {
    extern __shared__ __align__(16) unsigned char msm[];
    uint64_t* sbuf = reinterpret_cast<uint64_t*>(msm);  // [sort_cap]
    const int kcp = next_pow2(a.kc);
    RefineSmem s;
    uint64_t* cbuf = carve_refine(msm + (size_t)a.sort_cap * 8, a.d, kcp, s);
    // the exact re-scan only starts after the select is over: its scratch shares the histograms' storage
    __shared__ __align__(16) unsigned s_whist[(MERGE_THREADS / 32) * 256 > (int)(sizeof(RescanSmem) / 4 + 1) ? (MERGE_THREADS / 32) * 256
                                                                                                         : (int)(sizeof(RescanSmem) / 4 + 1)];
    RescanSmem& rs = *reinterpret_cast<RescanSmem*>(s_whist);
    __shared__ uint64_t s_blist[SELECT_BOUNDARY_CAP];
    __shared__ int s_n, s_tot, s_flag;
    __shared__ unsigned s_sc[4], s_mm[2], s_T;
    __shared__ double s_qn;

    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x;
    if (tid == 0) { s_n = 0; s_tot = 0; s_mm[0] = 0xffffffffu; s_mm[1] = 0u; }
    pdl_wait();  // launched as a programmatic dependent of the filter kernel: nothing of its output is read above
    for (int i = tid; i < a.d; i += blockDim.x) s.qs[i] = a.queries[q * a.q_pitch + i];
    __syncthreads();
    block_query_sqnorm(s.qs, a.d, &s_qn);

    // ---- gather survivors: appended candidates whose filter key is within the final threshold
    const uint32_t gthr = a.gthr[q];
    const int appended = a.qcount[(size_t)q * a.qstride];
    if (tid == 0) {  // statistics for qk_scan_partitions' `stats`
        atomicMax(&a.ctrl[3], appended);
        atomicAdd(reinterpret_cast<unsigned long long*>(a.ctrl + 4), (unsigned long long)appended);
    }
    bool overflow = appended > a.qcap;  // entries were dropped: only the exact re-scan can answer
    {
        const int n = appended < a.qcap ? appended : a.qcap;
        const uint64_t* c = a.qbuf + (size_t)q * a.qcap;
        uint32_t mn = 0xffffffffu, mx = 0u;
        // four loads in flight per thread (a few hundred entries per query: one at a time they were dependent L2 trips)
        for (int base = 0; base < n; base += 4 * MERGE_THREADS) {
            uint64_t v4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * MERGE_THREADS + tid;
                v4[u] = i < n ? __ldcs(reinterpret_cast<const unsigned long long*>(c) + i) : ~0ull;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint64_t v = v4[u];
                const uint32_t key = (uint32_t)(v >> 32);
                if (key > gthr || key == KEY_MAX) continue;
                mn = min(mn, key);
                mx = max(mx, key);
                const int pos = atomicAdd(&s_n, 1);
                if (pos < a.sort_cap) sbuf[pos] = v;
                else overflow = true;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if ((tid & 31) == 0 && mn <= mx) { atomicMin(&s_mm[0], mn); atomicMax(&s_mm[1], mx); }
    }
    overflow = __syncthreads_or(overflow);
    const int ns = s_n;
    if (ns < a.k && !overflow) {
        // Fewer survivors than results wanted is only legitimate when the query probed fewer than k rows altogether.
        // Anything else (rows that scored NaN, a threshold that somehow got too tight) goes to the exact re-scan.
        // Between k and kc survivors is fine: all of them are refined and the proof below decides -- thresholds that
        // follow the running minimum (top-1 mode) legitimately leave only a handful.
        int total = 0;
        const int nslots = probed_slots(a);
        for (int j = tid; j < nslots; j += blockDim.x) {
            const int seg = probed_segment(a, q, j);
            if (seg >= 0) total += a.seg_rows[seg];
        }
        if (total) atomicAdd(&s_tot, total);
        __syncthreads();
        overflow = s_tot > ns;
    }
    bool rescan = overflow || a.force_rescan;
    const uint64_t* cand = sbuf;
    int m = ns;
    uint32_t a_key = gthr;  // every row that is not a survivor has a filter key above the final threshold
    bool have_rejects = gthr != KEY_MAX;  // KEY_MAX: no threshold ever existed, every valid row is a survivor
    if (!rescan && ns > kcp) {
        // more survivors than the refine holds: keep the kc best by filter key
        const uint64_t* sb = sbuf;
        m = block_select_smallest([sb](int i) { return sb[i]; }, ns, a.kc, s_mm[0], s_mm[1], cbuf, &s_T, s_whist, s_sc, s_blist);
        if (m < 0) { rescan = true; m = 0; }  // a pile of equal filter keys at the boundary
        cand = cbuf;
        a_key = s_T;
        have_rejects = true;
    }
    refine_and_emit<kIP>(a, q, cand, m, a_key, have_rejects, rescan, s, s_qn, rs, &s_flag);
}

// ------------------------------------------------------------------------------------------------
// dense mode (single-list stores whose [Q x rows] key matrix is small: the coarse centroid scan): the filter stored
// the key of every (query, row); select the kc best here -- no thresholds, seeds, atomics or buffers in the filter
// ------------------------------------------------------------------------------------------------
template <bool kIP>
__global__ void __launch_bounds__(MERGE_THREADS, MERGE_THREADS == 128 ? 7 : 5) dense_refine_kernel(const MergeArgs a) {
    extern __shared__ __align__(16) unsigned char msm[];
    uint32_t* dkeys = reinterpret_cast<uint32_t*>(msm);  // [dense_rows rounded up to 2]
    const int kcp = next_pow2(a.kc);
    RefineSmem s;
    uint64_t* cbuf = carve_refine(msm + (size_t)((a.dense_rows + 1) & ~1) * 4, a.d, kcp, s);
    // the exact re-scan only starts after the select is over: its scratch shares the histograms' storage
    __shared__ __align__(16) unsigned s_whist[(MERGE_THREADS / 32) * 256 > (int)(sizeof(RescanSmem) / 4 + 1) ? (MERGE_THREADS / 32) * 256
                                                                                                         : (int)(sizeof(RescanSmem) / 4 + 1)];
    RescanSmem& rs = *reinterpret_cast<RescanSmem*>(s_whist);
    __shared__ uint64_t s_blist[SELECT_BOUNDARY_CAP];
    __shared__ int s_flag;
    __shared__ unsigned s_sc[4], s_mm[2], s_T, s_valid;
    __shared__ double s_qn;

    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x;
    const int rows = a.dense_rows;
    const uint32_t* src = a.dense + (size_t)q * rows;
    pdl_launch_dependents();  // the grouping kernel may become resident (it waits for this grid's completion itself)
    if (tid == 0) { s_mm[0] = 0xffffffffu; s_mm[1] = 0u; s_valid = 0u; }
    for (int i = tid; i < a.d; i += blockDim.x) s.qs[i] = a.queries[q * a.q_pitch + i];
    __syncthreads();
    {
        // stage the query's keys; key range and the number of valid (non-NaN) scores on the way
        uint32_t mn = 0xffffffffu, mx = 0u, valid = 0;
        auto note = [&](uint32_t key) {
            if (key != KEY_MAX) { mn = min(mn, key); mx = max(mx, key); ++valid; }
        };
        if ((rows & 3) == 0) {
            // 16-byte loads, four in flight per thread (one at a time this loop was a chain of L2 round trips:
            // 28 % of the kernel's stall samples, profiles/r02i_ncu_details_refine_seed_merge.txt)
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(dkeys);
            const int n4 = rows >> 2;
            for (int base = 0; base < n4; base += 4 * MERGE_THREADS) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = base + u * MERGE_THREADS + tid;
                    v[u] = i < n4 ? __ldcs(s4 + i) : make_uint4(KEY_MAX, KEY_MAX, KEY_MAX, KEY_MAX);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = base + u * MERGE_THREADS + tid;
                    if (i < n4) {
                        d4[i] = v[u];
                        note(v[u].x); note(v[u].y); note(v[u].z); note(v[u].w);
                    }
                }
            }
        } else {
#pragma unroll 4
            for (int i = tid; i < rows; i += MERGE_THREADS) {
                const uint32_t key = src[i];
                dkeys[i] = key;
                note(key);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            valid += __shfl_xor_sync(0xffffffffu, valid, o);
        }
        if ((tid & 31) == 0 && valid) { atomicMin(&s_mm[0], mn); atomicMax(&s_mm[1], mx); atomicAdd(&s_valid, valid); }
    }
    block_query_sqnorm(s.qs, a.d, &s_qn);
    __syncthreads();
    bool rescan = a.force_rescan != 0;
    int m = 0;
    if ((int)s_valid < a.kc) {
        rescan = true;  // fewer than kc rows with a valid score
    } else {
        // invalid keys (KEY_MAX) are clamped into the top bin: they are never among the kc smallest of >= kc valid ones
        const uint32_t* dk = dkeys;
        const uint32_t kmax = s_mm[1];
        const uint32_t r0 = (uint32_t)a.dense_row0;
        m = block_select_smallest(
            [dk, kmax, r0](int i) { const uint32_t key = dk[i]; return ((uint64_t)(key < kmax ? key : kmax) << 32) | (r0 + (uint32_t)i); },
            rows, a.kc, s_mm[0], kmax, cbuf, &s_T, s_whist, s_sc, s_blist);
        if (m < 0) { rescan = true; m = 0; }
    }
    if (tid == 0) {
        atomicMax(&a.ctrl[3], m);
        atomicAdd(reinterpret_cast<unsigned long long*>(a.ctrl + 4), (unsigned long long)m);
    }
    if (a.set_mode && !rescan && m > a.k && set_select_and_emit<kIP>(a, q, cbuf, m, s_T, /*have_rejects=*/rows > m, s, s_qn)) return;
    refine_and_emit<kIP>(a, q, cbuf, m, s_T, /*have_rejects=*/rows > m, rescan, s, s_qn, rs, &s_flag);
}

// ------------------------------------------------------------------------------------------------
// collect mode (APS rounds): ALL survivors of a scan with given, fixed thresholds are refined exactly and handed out
// grouped by the probe rank of the list they came from, each group best first (distance, then id) and cut at k -- the
// per-(query, partition) result lists the sequential APS loop consumes (query_coordinator.cpp:537-580), without a
// pseudo-query per partition.
// ------------------------------------------------------------------------------------------------
static constexpr int COLLECT_CAP = 4096;
struct CollectArgs {
    const float* vecs;
    int64_t pitch;
    const int64_t* ids;
    int d;
    const int64_t* seg_row0;
    const int32_t* seg_rows;
    const float* queries;
    int64_t q_pitch;
    const int32_t* pair_seg;  // [Q x R] (single-segment lists: pair slot == probe rank)
    int R, k;
    const uint64_t* qbuf;
    const int32_t* qcount;
    int qstride;
    int qcap, cap, np;        // cap = candidates handled per query (<= COLLECT_CAP), np = next power of two
    int64_t* out_ids;         // [Q x R x k]
    float* out_dist;
    int32_t* out_cnt;         // [Q x R]
    int32_t* overflow;        // [Q]
};

template <bool kIP>
__global__ void __launch_bounds__(MERGE_THREADS) collect_refine_kernel(const CollectArgs a) {
    extern __shared__ __align__(16) unsigned char csm[];
    uint64_t* comp = reinterpret_cast<uint64_t*>(csm);                 // [np] rank << 44 | distance key << 12 | slot
    int64_t* cid = reinterpret_cast<int64_t*>(comp + a.np);            // [cap]
    float* qs = reinterpret_cast<float*>(cid + a.cap);                 // [d padded]
    long long* srow0 = reinterpret_cast<long long*>(qs + ((a.d + 3) & ~3) + 2);
    srow0 = reinterpret_cast<long long*>((reinterpret_cast<uintptr_t>(srow0) + 7) & ~(uintptr_t)7);  // [R]
    int* srows = reinterpret_cast<int*>(srow0 + a.R);                  // [R]
    int* first = srows + a.R;                                          // [R] first sorted position of every rank
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x;
    const int appended = a.qcount[(size_t)q * a.qstride];
    for (int j = tid; j < a.R; j += MERGE_THREADS) a.out_cnt[q * a.R + j] = 0;
    if (appended > a.cap) {
        if (tid == 0) a.overflow[q] = 1;
        return;
    }
    if (tid == 0) a.overflow[q] = 0;
    const int n = appended;
    for (int i = tid; i < a.d; i += MERGE_THREADS) qs[i] = a.queries[q * a.q_pitch + i];
    for (int j = tid; j < a.R; j += MERGE_THREADS) {
        const int seg = a.pair_seg[q * a.R + j];
        srow0[j] = seg >= 0 ? a.seg_row0[seg] : -1;
        srows[j] = seg >= 0 ? a.seg_rows[seg] : 0;
        first[j] = n;
    }
    __syncthreads();
    const uint64_t* cand = a.qbuf + (size_t)q * a.qcap;
    for (int base = 0; base < a.np; base += MERGE_THREADS / 8) {
        const int i = base + (tid >> 3), j8 = tid & 7;
        if (i < n) {  // uniform inside every 8-lane group
            const uint32_t row = (uint32_t)cand[i];
            const float dist = ref_pair_distance_g8<kIP>(qs, a.vecs + (int64_t)row * a.pitch, a.d, j8);
            if (j8 == 0) {
                int rank = 0;
                for (int j = 0; j < a.R; ++j)
                    if (srow0[j] >= 0 && (long long)row >= srow0[j] && (long long)row < srow0[j] + srows[j]) { rank = j; break; }
                const uint32_t dk = f2key(kIP ? -dist : __fsqrt_rn(dist));
                comp[i] = ((uint64_t)rank << 44) | ((uint64_t)dk << 12) | (uint64_t)i;
                cid[i] = a.ids ? a.ids[row] : (int64_t)row;
            }
        } else if (i < a.np && j8 == 0) {
            comp[i] = COMP_MAX;
        }
    }
    // (rank, distance key) order; ties between equal keys of one rank by id
    const int64_t* cidc = cid;
    block_bitonic_sort(comp, a.np, [cidc](uint64_t x, uint64_t y) {
        const uint64_t hx = x >> 12, hy = y >> 12;
        if (hx != hy) return hx < hy;
        if (x == COMP_MAX || y == COMP_MAX) return x < y;
        return cidc[x & 0xfffu] < cidc[y & 0xfffu];
    });
    for (int i = tid; i < n; i += MERGE_THREADS) {
        const int r = (int)(comp[i] >> 44);
        if (i == 0 || (int)(comp[i - 1] >> 44) != r) first[r] = i;
    }
    __syncthreads();
    for (int i = tid; i < n; i += MERGE_THREADS) {
        const uint64_t c = comp[i];
        const int r = (int)(c >> 44);
        const int idx = i - first[r];
        if (idx < a.k) {
            const float v = key2f((uint32_t)(c >> 12));
            const size_t o = ((size_t)q * a.R + r) * a.k + idx;
            a.out_ids[o] = cid[c & 0xfffu];
            a.out_dist[o] = kIP ? -v : v;
        }
        if (i + 1 == n || (int)(comp[i + 1] >> 44) != r) a.out_cnt[q * a.R + r] = idx + 1 < a.k ? idx + 1 : a.k;
    }
}

}  // namespace qk
