// Partition scan + per-query top-k for sm_100a.
//
// Replaces, on the GPU, the reference's scan_list / batched_scan_list + TopkBuffer
// (/root/reference/src/cpp/include/list_scanning.h:241-366, 41-204) as driven by
// QueryCoordinator::serial_scan / batched_serial_scan (src/cpp/src/query_coordinator.cpp:471-611,
// 675-799).
//
// Pipeline (all on one stream, no host synchronisation):
//   1. expand_pairs      (query, probed list) -> (query, segment) pairs; histogram of pairs per segment
//   2. prefix_segments   exclusive scan -> per-segment group offsets and the work-item list size
//   3. scatter_pairs     group the pairs by segment (the reference's batched_serial_scan grouping,
//                        query_coordinator.cpp:708-721) and emit work items (segment, query chunk)
//   4. scan_kernel       FILTER: persistent, warp-specialised CTAs (one per SM). A producer warp pulls work
//                        items (segment, chunk of <= 32 queries) from an atomic counter and streams the
//                        segment's rows through a 4-stage shared-memory ring with TMA bulk copies
//                        (mbarrier full/empty pipeline); two compute groups of 4 warps take alternate
//                        64-row tiles and score them against the staged query chunk with fp32 FFMA2
//                        (||v||^2 - 2<q,v> for l2 with precomputed row norms, -<q,v> for ip), writing
//                        order-preserving keys to shared memory; four select warps keep, per (query,
//                        segment), a running top-kc in a warp-resident sorted array, pruned by a per-query
//                        global threshold (atomicMin). No CTA-wide barrier inside the loop.
//   5. merge_refine      REFINE: per query, the kc best candidates by filter score are re-evaluated in
//                        the reference's exact summation order (common.cuh: ref_pair_distance), sorted
//                        by (distance, id), and a rigorous rounding-error bound proves no rejected
//                        vector could belong to the top-k. If the proof fails (duplicates, ties at the
//                        boundary) the query is flagged ...
//   6. exact_rescan      ... and re-scanned exhaustively in exact arithmetic (rare).
#include "common.cuh"

namespace qk {

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct ScanPlan {
    int P;        // (query, segment) pair slots per query
    int kc;       // candidates kept per (query, segment) and refined per query
    int gq;       // queries per work item
    int nq;       // query-chunk ring depth
    int dp;       // padded dimension (multiple of 4)
    size_t smem;  // dynamic shared memory of scan_kernel
    // workspace offsets (bytes)
    size_t off_pair_seg, off_seg_count, off_seg_fill, off_seg_start, off_item_start, off_seg_pairs, off_items,
        off_ctrl, off_gthr, off_flags, off_cand_n, off_cand, total;
};

static constexpr int SCAN_DC = 128;     // floats of a row staged per pipeline unit
static constexpr int SCAN_VP = SCAN_DC + 4;  // padded smem row stride (floats): 8 rows x 16 B hit 32 distinct banks
static constexpr int SCAN_STAGES = 4;
static constexpr int SCAN_GQ = 32;      // max queries per work item
static constexpr int SCAN_TV = 64;      // rows per tile
static constexpr int SCAN_KP = SCAN_TV + 4;  // padded key row stride (u32)
static constexpr int SCAN_COMPUTE_WARPS = 8;  // two groups of four
static constexpr int SCAN_SELECT_WARPS = 4;
static constexpr int SCAN_THREADS = 32 * (SCAN_COMPUTE_WARPS + SCAN_SELECT_WARPS + 1);
static constexpr int SCAN_MAX_NQ = 4;   // query-chunk ring depth
static constexpr int MERGE_THREADS = 256;
static constexpr int MERGE_SORT_CAP = 4096;
static constexpr size_t SCAN_SMEM_LIMIT = 227 * 1024;

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int candidate_count(int k) { return k + (k / 16 > 6 ? k / 16 : 6); }

struct ItemDesc {  // published by the producer warp for every work item
    int seg;     // -1: no more work
    int g_begin, g_cnt;
    int nrows;
    long long row0;
};

static size_t scan_smem_bytes(int dp, int kc, int gq, int nq) {
    size_t b = 512;                                                          // mbarriers + item descriptors
    b += (size_t)SCAN_STAGES * SCAN_TV * SCAN_VP * sizeof(float);            // V ring
    b += (size_t)nq * gq * (dp + 4) * sizeof(float);                         // query-chunk ring
    b += (size_t)4 * gq * SCAN_KP * sizeof(uint32_t);                        // score keys: 2 groups x 2 buffers
    b += (size_t)gq * kc * sizeof(uint64_t);                                 // sorted candidate arrays
    return b;
}

static int make_plan(const qk_store_t* st, int64_t Q, int nprobe, int k, ScanPlan* p) {
    QK_REQUIRE(k >= 1 && k <= QK_MAX_K, "k=%d out of range [1, %d]", k, QK_MAX_K);
    QK_REQUIRE(st->d >= 1 && st->pitch >= st->d && st->pitch % 4 == 0, "bad store d=%d pitch=%lld", st->d,
               (long long)st->pitch);
    p->dp = (st->d + 3) / 4 * 4;
    // a query probes each list at most once: the segments beyond the first of its probed lists number
    // at most min(nprobe * (max_list_segments - 1), num_segments)
    int64_t extra = (int64_t)nprobe * (st->max_list_segments > 1 ? st->max_list_segments - 1 : 0);
    if (extra > st->num_segments) extra = st->num_segments;
    p->P = nprobe + (int)extra;
    p->kc = candidate_count(k);
    // queries per work item: bounded by shared memory (candidate arrays + staged query rows)
    int gq = SCAN_GQ;
    while (gq > 1 && scan_smem_bytes(p->dp, p->kc, gq, 2) > SCAN_SMEM_LIMIT) gq >>= 1;
    QK_REQUIRE(scan_smem_bytes(p->dp, p->kc, gq, 2) <= SCAN_SMEM_LIMIT,
               "scan kernel shared memory exceeds 227 KB (k=%d d=%d)", k, st->d);
    p->gq = gq;
    p->nq = 2;
    while (p->nq < SCAN_MAX_NQ && scan_smem_bytes(p->dp, p->kc, gq, p->nq + 1) <= SCAN_SMEM_LIMIT) p->nq++;
    p->smem = scan_smem_bytes(p->dp, p->kc, gq, p->nq);
    QK_REQUIRE(Q * (int64_t)p->P < (int64_t)1 << 30, "too many (query, segment) pairs; split the query batch");
    const size_t QP = (size_t)Q * p->P;
    const size_t S = (size_t)st->num_segments;
    size_t o = 0;
    p->off_pair_seg = o;   o = align_up(o + QP * 4, 256);
    p->off_seg_count = o;  o = align_up(o + (S + 1) * 4, 256);
    p->off_seg_fill = o;   o = align_up(o + (S + 1) * 4, 256);
    p->off_flags = o;      o = align_up(o + (size_t)Q * 4, 256);
    p->off_ctrl = o;       o = align_up(o + 64, 256);
    p->off_seg_start = o;  o = align_up(o + (S + 1) * 4, 256);
    p->off_item_start = o; o = align_up(o + (S + 1) * 4, 256);
    p->off_seg_pairs = o;  o = align_up(o + QP * 4, 256);
    size_t max_items = QP / gq + (QP < S ? QP : S) + 1;
    p->off_items = o;      o = align_up(o + max_items * 8, 256);
    p->off_gthr = o;       o = align_up(o + (size_t)Q * 4, 256);
    p->off_cand_n = o;     o = align_up(o + QP * 4, 256);
    p->off_cand = o;       o = align_up(o + QP * p->kc * 8, 256);
    p->total = o;
    return QK_OK;
}

// ------------------------------------------------------------------------------------------------
// 1. expand (query, list) -> (query, segment) pairs + histogram
// ------------------------------------------------------------------------------------------------
__global__ void expand_pairs_kernel(const int32_t* __restrict__ probe, int64_t Q, int nprobe, int P,
                                    const int32_t* __restrict__ list_seg0, const int32_t* __restrict__ list_nseg,
                                    int num_lists, int32_t* __restrict__ pair_seg, int32_t* __restrict__ seg_count,
                                    uint32_t* __restrict__ gthr, bool single_segment_lists) {
    if (single_segment_lists) {
        // one thread per (query, probe)
        int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= Q * nprobe) return;
        int64_t q = i / nprobe;
        int j = (int)(i - q * nprobe);
        if (j == 0) gthr[q] = KEY_MAX;
        int l = probe[i];
        int seg = -1;
        if (l >= 0 && l < num_lists && list_nseg[l] > 0) seg = list_seg0[l];
        pair_seg[q * P + j] = seg;  // P == nprobe here
        if (seg >= 0) atomicAdd(&seg_count[seg], 1);
    } else {
        // one thread per query, sequential over its probes
        int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (q >= Q) return;
        gthr[q] = KEY_MAX;
        int pos = 0;
        for (int j = 0; j < nprobe; ++j) {
            int l = probe[q * nprobe + j];
            if (l < 0 || l >= num_lists) continue;
            int s0 = list_seg0[l], ns = list_nseg[l];
            for (int s = 0; s < ns && pos < P; ++s) {
                pair_seg[q * P + pos++] = s0 + s;
                atomicAdd(&seg_count[s0 + s], 1);
            }
        }
        for (; pos < P; ++pos) pair_seg[q * P + pos] = -1;
    }
}

// ------------------------------------------------------------------------------------------------
// 2. exclusive scan of the per-segment pair counts (one CTA)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) prefix_segments_kernel(const int32_t* __restrict__ seg_count, int S, int gq,
                                                               int32_t* __restrict__ seg_start,
                                                               int32_t* __restrict__ item_start,
                                                               int32_t* __restrict__ ctrl) {
    __shared__ int2 warp_tot[32];
    __shared__ int2 carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = make_int2(0, 0);
    __syncthreads();
    for (int base = 0; base < S; base += 1024) {
        int s = base + tid;
        int c = (s < S) ? seg_count[s] : 0;
        int it = (c + gq - 1) / gq;
        int xc = c, xi = it;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int yc = __shfl_up_sync(0xffffffffu, xc, o);
            int yi = __shfl_up_sync(0xffffffffu, xi, o);
            if (lane >= o) { xc += yc; xi += yi; }
        }
        if (lane == 31) warp_tot[warp] = make_int2(xc, xi);
        __syncthreads();
        if (warp == 0) {
            int2 w = warp_tot[lane];
            int wc = w.x, wi = w.y;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int yc = __shfl_up_sync(0xffffffffu, wc, o);
                int yi = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) { wc += yc; wi += yi; }
            }
            warp_tot[lane] = make_int2(wc, wi);  // inclusive
        }
        __syncthreads();
        int2 cy = carry;
        int2 wprev = warp ? warp_tot[warp - 1] : make_int2(0, 0);
        if (s < S) {
            seg_start[s] = cy.x + wprev.x + xc - c;
            item_start[s] = cy.y + wprev.y + xi - it;
        }
        __syncthreads();
        if (tid == 1023) carry = make_int2(cy.x + wprev.x + xc, cy.y + wprev.y + xi);
        __syncthreads();
    }
    if (tid == 0) {
        seg_start[S] = carry.x;
        item_start[S] = carry.y;
        ctrl[0] = 0;        // work counter
        ctrl[1] = carry.y;  // number of work items
        ctrl[2] = 0;        // queries sent to exact_rescan
    }
}

// ------------------------------------------------------------------------------------------------
// 3. group pairs by segment, emit work items
// ------------------------------------------------------------------------------------------------
__global__ void scatter_pairs_kernel(const int32_t* __restrict__ pair_seg, int64_t QP, int S,
                                     const int32_t* __restrict__ seg_start, int32_t* __restrict__ seg_fill,
                                     int32_t* __restrict__ seg_pairs, const int32_t* __restrict__ item_start,
                                     int2* __restrict__ items) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < QP) {
        int seg = pair_seg[i];
        if (seg >= 0) {
            int pos = seg_start[seg] + atomicAdd(&seg_fill[seg], 1);
            seg_pairs[pos] = (int32_t)i;
        }
    }
    if (i < S) {
        int b = item_start[i], e = item_start[i + 1];
        for (int c = b; c < e; ++c) items[c] = make_int2((int)i, c - b);
    }
}

// ------------------------------------------------------------------------------------------------
// warp-resident sorted arrays (ascending composite keys), blocked layout: lane holds E consecutive
// elements. insert() requires c < current maximum.
// ------------------------------------------------------------------------------------------------
template <int E>
__device__ __forceinline__ void warp_insert(uint64_t (&a)[E], uint64_t c, int lane) {
    int cnt = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) cnt += (a[e] < c) ? 1 : 0;
    unsigned open = __ballot_sync(0xffffffffu, cnt < E);
    uint64_t carry = a[E - 1];
    {
        uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)carry, 1);
        uint32_t hi = __shfl_up_sync(0xffffffffu, (uint32_t)(carry >> 32), 1);
        carry = ((uint64_t)hi << 32) | lo;
    }
    if (open == 0) return;
    const int li = __ffs(open) - 1;
    if (lane > li) {
#pragma unroll
        for (int e = E - 1; e >= 1; --e) a[e] = a[e - 1];
        a[0] = carry;
    } else if (lane == li) {
#pragma unroll
        for (int e = E - 1; e >= 0; --e) {
            if (e > cnt) a[e] = a[e - 1 < 0 ? 0 : e - 1];
            else if (e == cnt) a[e] = c;
        }
    }
}

template <int E>
__device__ __forceinline__ uint64_t warp_element(const uint64_t (&a)[E], int idx) {
    const int src = idx / E, slot = idx % E;
    uint64_t v = a[0];
#pragma unroll
    for (int e = 1; e < E; ++e)
        if (slot == e) v = a[e];
    return shfl_u64(v, src);
}

// generic (any n) insertion into a shared-memory sorted array, warp-cooperative
__device__ __forceinline__ void smem_insert(uint64_t* arr, int n, uint64_t c, int lane) {
    int lo = 0, hi = n - 1;  // arr[n-1] > c guaranteed
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (arr[mid] > c) hi = mid; else lo = mid + 1;
    }
    const int p = lo;
    for (int base = n - 1; base > p; base -= 32) {
        int i = base - lane;
        uint64_t v = 0;
        bool act = i > p;
        if (act) v = arr[i - 1];
        __syncwarp();
        if (act) arr[i] = v;
        __syncwarp();
    }
    if (lane == 0) arr[p] = c;
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// 4. the scan (filter) kernel
// ------------------------------------------------------------------------------------------------
struct ScanArgs {
    const float* vecs;
    const float* norms;  // squared row norms (l2 only)
    int64_t pitch;
    int dp;
    const int64_t* seg_row0;
    const int32_t* seg_rows;
    const float* queries;
    int64_t q_pitch;
    const int32_t* seg_start;
    const int32_t* seg_pairs;
    const int2* items;
    int32_t* ctrl;
    uint32_t* gthr;
    uint64_t* cand;
    int32_t* cand_n;
    int P, kc, gq, nq;
};

// process the keys of one tile for one query with a register-resident sorted array
template <int E>
__device__ __forceinline__ void select_tile(uint64_t* top_g, int kc, const uint32_t* keys, int tile_row0, int lane,
                                            uint32_t& gthr_q, uint32_t* gthr_global) {
    uint64_t a[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        int i = lane * E + e;
        a[e] = (i < kc) ? top_g[i] : COMP_MAX;
    }
    uint64_t thr = warp_element<E>(a, kc - 1);
    const uint64_t thr_in = thr;
    bool changed = false;  // warp-uniform
#pragma unroll
    for (int j = 0; j < SCAN_TV / 32; ++j) {
        const int r = j * 32 + lane;
        const uint32_t key = keys[r];
        const uint64_t comp = ((uint64_t)key << 32) | (uint32_t)(tile_row0 + r);
        bool pass = (key != KEY_MAX) && (key <= gthr_q) && (comp < thr);
        unsigned m = __ballot_sync(0xffffffffu, pass);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint64_t c = shfl_u64(comp, src);
            if (c < thr) {
                warp_insert<E>(a, c, lane);
                thr = warp_element<E>(a, kc - 1);
                changed = true;
            }
        }
    }
    if (changed) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            int i = lane * E + e;
            if (i < kc) top_g[i] = a[e];
        }
        if (thr != thr_in && thr != COMP_MAX) {
            const uint32_t tk = (uint32_t)(thr >> 32);
            if (tk < gthr_q) {
                uint32_t old = 0;
                if (lane == 0) old = atomicMin(gthr_global, tk);
                old = __shfl_sync(0xffffffffu, old, 0);
                gthr_q = old < tk ? old : tk;
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void select_tile_generic(uint64_t* top_g, int kc, const uint32_t* keys, int tile_row0,
                                                    int lane, uint32_t& gthr_q, uint32_t* gthr_global) {
    uint64_t thr = top_g[kc - 1];
    const uint64_t thr_in = thr;
    for (int j = 0; j < SCAN_TV / 32; ++j) {
        const int r = j * 32 + lane;
        const uint32_t key = keys[r];
        const uint64_t comp = ((uint64_t)key << 32) | (uint32_t)(tile_row0 + r);
        bool pass = (key != KEY_MAX) && (key <= gthr_q) && (comp < thr);
        unsigned m = __ballot_sync(0xffffffffu, pass);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint64_t c = shfl_u64(comp, src);
            if (c < thr) {
                smem_insert(top_g, kc, c, lane);
                thr = top_g[kc - 1];
            }
        }
    }
    if (thr != thr_in && thr != COMP_MAX) {
        const uint32_t tk = (uint32_t)(thr >> 32);
        if (tk < gthr_q) {
            uint32_t old = 0;
            if (lane == 0) old = atomicMin(gthr_global, tk);
            old = __shfl_sync(0xffffffffu, old, 0);
            gthr_q = old < tk ? old : tk;
        }
    }
}

// Shared-memory map (dynamic):
//   [0, 512)        mbarriers: full[4] empty[4] qfull[4] qempty[4] kfull[2][2] kempty[2][2]; ItemDesc[4] at +256
//   Vs   [STAGES][TV][VP] f32      row ring, one stage = one tile x one d-chunk
//   Qs   [nq][gq][dp+4]   f32      query-chunk ring, one slot per work item in flight
//   Ks   [2 groups][2][gq][KP] u32 score keys
//   top  [gq][kc] u64              per-query sorted candidate arrays of the current item (select warps only)
//
// Roles: warps 0-3 compute group 0 (even tiles), warps 4-7 compute group 1 (odd tiles), warps 8-11 select,
// warp 12 producer. Inside a compute group, warp gw covers rows rb*32.. (rb = gw & 1) x query half qh = gw >> 1;
// lane (lr = lane & 7, lq = lane >> 3) owns rows rb*32 + lr + 8i (i < 4) and queries qh + 2*(lq + 4t) (t < 4), so
// every shared-memory load is one wavefront (8 distinct rows or 4 distinct queries, the rest broadcast).
template <bool kIP>
__global__ void __launch_bounds__(SCAN_THREADS, 1) scan_kernel(const ScanArgs a) {
    constexpr int TV = SCAN_TV, DC = SCAN_DC, VP = SCAN_VP, KP = SCAN_KP, NS = SCAN_STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* full = bars;             // [NS]   producer -> compute (tx)
    uint64_t* empty = bars + 4;        // [NS]   compute  -> producer (4 warps of the owning group)
    uint64_t* qfull = bars + 8;        // [nq]   producer -> all consumers (tx + descriptor)
    uint64_t* qempty = bars + 12;      // [nq]   12 consumer warps -> producer
    uint64_t* kfull = bars + 16;       // [2][2] compute group -> select (4 warps)
    uint64_t* kempty = bars + 20;      // [2][2] select (4 warps) -> compute group
    ItemDesc* descs = reinterpret_cast<ItemDesc*>(smem_raw + 256);  // [SCAN_MAX_NQ]
    float* Vs = reinterpret_cast<float*>(smem_raw + 512);
    const int dp = a.dp, kc = a.kc, gq = a.gq, nq = a.nq;
    const int QP = dp + 4;
    float* Qs = Vs + (size_t)NS * TV * VP;
    uint32_t* Ks = reinterpret_cast<uint32_t*>(Qs + (size_t)nq * gq * QP);
    uint64_t* top = reinterpret_cast<uint64_t*>(Ks + (size_t)4 * gq * KP);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 4); }
        for (int s = 0; s < SCAN_MAX_NQ; ++s) {
            mbar_init(qfull + s, 1);
            mbar_init(qempty + s, SCAN_COMPUTE_WARPS + SCAN_SELECT_WARPS);
        }
        for (int s = 0; s < 4; ++s) { mbar_init(kfull + s, 4); mbar_init(kempty + s, SCAN_SELECT_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    const int ndc = (dp + DC - 1) / DC;

    if (warp == SCAN_COMPUTE_WARPS + SCAN_SELECT_WARPS) {
        // ===================================================================== producer warp
        const int n_items = a.ctrl[1];
        uint32_t U = 0;
        for (uint32_t n = 0;; ++n) {
            const int ib = n % nq;
            mbar_wait(qempty + ib, ((n / nq) & 1u) ^ 1u);
            int it = 0;
            if (lane == 0) it = atomicAdd(&a.ctrl[0], 1);
            it = __shfl_sync(0xffffffffu, it, 0);
            if (it >= n_items) {
                if (lane == 0) {
                    descs[ib].seg = -1;
                    mbar_arrive(qfull + ib);
                }
                break;
            }
            const int2 item = a.items[it];
            const int seg = item.x;
            const int g_begin = a.seg_start[seg] + item.y * gq;
            int g_cnt = a.seg_start[seg + 1] - g_begin;
            g_cnt = g_cnt < gq ? g_cnt : gq;
            const int64_t row0 = a.seg_row0[seg];
            const int nrows = a.seg_rows[seg];
            if (lane == 0) {
                ItemDesc d;
                d.seg = seg; d.g_begin = g_begin; d.g_cnt = g_cnt; d.nrows = nrows; d.row0 = row0;
                descs[ib] = d;
                mbar_expect_tx(qfull + ib, (uint32_t)(g_cnt * dp * 4));
            }
            __syncwarp();
            if (lane < g_cnt) {
                const int pair = a.seg_pairs[g_begin + lane];
                const int64_t q = pair / a.P;
                bulk_g2s(Qs + ((size_t)ib * gq + lane) * QP, a.queries + q * a.q_pitch, (uint32_t)(dp * 4), qfull + ib);
            }
            const int ntiles = (nrows + TV - 1) / TV;
            for (int tile = 0; tile < ntiles; ++tile) {
                const int tr = min(TV, nrows - tile * TV);
                for (int dc = 0; dc < ndc; ++dc, ++U) {
                    const int st = U % NS;
                    mbar_wait(empty + st, ((U / NS) & 1u) ^ 1u);
                    const int dcur = min(DC, dp - dc * DC);
                    if (lane == 0) mbar_expect_tx(full + st, (uint32_t)(tr * dcur * 4));
                    __syncwarp();
                    for (int r = lane; r < tr; r += 32)
                        bulk_g2s(Vs + ((size_t)st * TV + r) * VP,
                                 a.vecs + (row0 + (int64_t)tile * TV + r) * a.pitch + dc * DC, (uint32_t)(dcur * 4),
                                 full + st);
                }
            }
        }
    } else if (warp >= SCAN_COMPUTE_WARPS) {
        // ===================================================================== select warps
        const int sw = warp - SCAN_COMPUTE_WARPS;
        uint32_t T = 0;
        for (uint32_t n = 0;; ++n) {
            const int ib = n % nq;
            mbar_wait(qfull + ib, (n / nq) & 1u);
            const ItemDesc d = descs[ib];
            if (d.seg < 0) break;
            // lane j < 8 keeps the (pair, threshold) of query g = sw + 4j
            int my_pair = -1;
            uint32_t my_gthr = KEY_MAX;
            {
                const int g = sw + 4 * lane;
                if (lane < 8 && g < d.g_cnt) {
                    my_pair = a.seg_pairs[d.g_begin + g];
                    my_gthr = a.gthr[my_pair / a.P];
                }
            }
            for (int g = sw; g < d.g_cnt; g += 4)
                for (int i = lane; i < kc; i += 32) top[(size_t)g * kc + i] = COMP_MAX;
            __syncwarp();
            const int ntiles = (d.nrows + TV - 1) / TV;
            for (int tile = 0; tile < ntiles; ++tile, ++T) {
                const int grp = T & 1u, kb = (T >> 1) & 1u;
                mbar_wait(kfull + grp * 2 + kb, (T >> 2) & 1u);
                const uint32_t* kbase = Ks + (size_t)(grp * 2 + kb) * gq * KP;
                for (int j = 0; sw + 4 * j < d.g_cnt; ++j) {
                    const int g = sw + 4 * j;
                    uint32_t gthr_q = __shfl_sync(0xffffffffu, my_gthr, j);
                    const int pair = __shfl_sync(0xffffffffu, my_pair, j);
                    uint32_t* gthr_global = a.gthr + pair / a.P;
                    if (kc <= 32)
                        select_tile<1>(top + (size_t)g * kc, kc, kbase + g * KP, tile * TV, lane, gthr_q, gthr_global);
                    else if (kc <= 128)
                        select_tile<4>(top + (size_t)g * kc, kc, kbase + g * KP, tile * TV, lane, gthr_q, gthr_global);
                    else
                        select_tile_generic(top + (size_t)g * kc, kc, kbase + g * KP, tile * TV, lane, gthr_q, gthr_global);
                    if (lane == j) my_gthr = gthr_q;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(kempty + grp * 2 + kb);
            }
            // emit the per-(query, segment) candidates
            for (int j = 0; sw + 4 * j < d.g_cnt; ++j) {
                const int g = sw + 4 * j;
                const int pair = __shfl_sync(0xffffffffu, my_pair, j);
                const uint64_t* tg = top + (size_t)g * kc;
                uint64_t* out = a.cand + (size_t)pair * kc;
                int cnt = 0;
                for (int base = 0; base < kc; base += 32) {
                    int i = base + lane;
                    uint64_t v = (i < kc) ? tg[i] : COMP_MAX;
                    if (v != COMP_MAX) out[i] = v;
                    cnt += __popc(__ballot_sync(0xffffffffu, v != COMP_MAX));
                }
                if (lane == 0) a.cand_n[pair] = cnt;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(qempty + ib);
        }
    } else {
        // ===================================================================== compute warps
        const int grp = warp >> 2, gw = warp & 3;
        const int rb = gw & 1, qh = gw >> 1;
        const int lr = lane & 7, lq = lane >> 3;
        uint32_t T = 0, U = 0;
        for (uint32_t n = 0;; ++n) {
            const int ib = n % nq;
            mbar_wait(qfull + ib, (n / nq) & 1u);
            const ItemDesc d = descs[ib];
            if (d.seg < 0) break;
            const int g_cnt = d.g_cnt;
            // queries of this lane: g = qh + 2*(lq + 4t); t < nt is warp-uniform
            const int jmax = (g_cnt - qh + 1) >> 1;            // number of j = lq + 4t with g < g_cnt
            const int nt = (jmax + 3) >> 2;                    // 0..4
            const float* qbase = Qs + ((size_t)ib * gq + qh + 2 * lq) * QP;
            const int ntiles = (d.nrows + TV - 1) / TV;
            for (int tile = 0; tile < ntiles; ++tile, ++T, U += ndc) {
                if ((int)(T & 1u) != grp) continue;
                const int tr = min(TV, d.nrows - tile * TV);
                float nrm[4];
                if (!kIP) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = rb * 32 + lr + 8 * i;
                        nrm[i] = (r < tr) ? __ldg(a.norms + d.row0 + (int64_t)tile * TV + r) : 0.f;
                    }
                }
                float2 acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int t = 0; t < 4; ++t) acc[i][t] = make_float2(0.f, 0.f);
                for (int dc = 0; dc < ndc; ++dc) {
                    const uint32_t u = U + dc;
                    const int st = u % NS;
                    mbar_wait(full + st, (u / NS) & 1u);
                    const int dcur4 = min(DC, dp - dc * DC) >> 2;
                    const float4* vrow = reinterpret_cast<const float4*>(Vs + ((size_t)st * TV + rb * 32 + lr) * VP);
                    const float4* qrow = reinterpret_cast<const float4*>(qbase + dc * DC);
                    constexpr int VSTEP = 8 * VP / 4;          // float4 stride between this lane's rows
                    const int qstep = 8 * QP / 4;              // float4 stride between this lane's queries
                    if (nt == 4) {
#pragma unroll 4
                        for (int c = 0; c < dcur4; ++c) {
                            float4 v[4], q[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[i] = vrow[i * VSTEP + c];
#pragma unroll
                            for (int t = 0; t < 4; ++t) q[t] = qrow[t * qstep + c];
#pragma unroll
                            for (int t = 0; t < 4; ++t)
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    acc[i][t] = ffma2(make_float2(v[i].x, v[i].y), make_float2(q[t].x, q[t].y), acc[i][t]);
                                    acc[i][t] = ffma2(make_float2(v[i].z, v[i].w), make_float2(q[t].z, q[t].w), acc[i][t]);
                                }
                        }
                    } else if (nt >= 1) {
#pragma unroll 2
                        for (int c = 0; c < dcur4; ++c) {
                            float4 v[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[i] = vrow[i * VSTEP + c];
#pragma unroll
                            for (int t = 0; t < 3; ++t) {
                                if (t < nt) {
                                    const float4 q4 = qrow[t * qstep + c];
#pragma unroll
                                    for (int i = 0; i < 4; ++i) {
                                        acc[i][t] = ffma2(make_float2(v[i].x, v[i].y), make_float2(q4.x, q4.y), acc[i][t]);
                                        acc[i][t] = ffma2(make_float2(v[i].z, v[i].w), make_float2(q4.z, q4.w), acc[i][t]);
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty + st);
                }
                // ---- tile epilogue: scores -> order-preserving keys -> Ks[grp][kb]
                const int kb = (T >> 1) & 1u;
                mbar_wait(kempty + grp * 2 + kb, ((T >> 2) & 1u) ^ 1u);
                uint32_t* kbase = Ks + (size_t)(grp * 2 + kb) * gq * KP;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int g = qh + 2 * (lq + 4 * t);
                    if (g < g_cnt) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = rb * 32 + lr + 8 * i;
                            const float dot = acc[i][t].x + acc[i][t].y;
                            const float sc = kIP ? -dot : fmaf(-2.f, dot, nrm[i]);
                            kbase[g * KP + r] = (r < tr) ? f2key(sc) : KEY_MAX;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(kfull + grp * 2 + kb);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(qempty + ib);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// block-wide bitonic sort of n (power of two) uint64 keys in shared memory
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

__device__ __forceinline__ int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

// ------------------------------------------------------------------------------------------------
// 5. merge + exact refine, one CTA per query
// ------------------------------------------------------------------------------------------------
struct MergeArgs {
    const float* vecs;
    int64_t pitch;
    const int64_t* ids;
    int d;
    const int64_t* seg_row0;
    const float* queries;
    int64_t q_pitch;
    const int32_t* pair_seg;
    const uint32_t* gthr;
    const uint64_t* cand;
    const int32_t* cand_n;
    int32_t* flags;
    int32_t* ctrl;
    int P, kc, k;
    float max_row_norm;
    int64_t* out_ids;
    float* out_dist;
    int64_t* out_rows;
    int force_rescan;
    int rank_squared;  // l2 only: order by the squared distance (k-means assign: faiss Top1 on squared l2)
};

template <bool kIP>
__global__ void __launch_bounds__(MERGE_THREADS) merge_refine_kernel(const MergeArgs a) {
    extern __shared__ __align__(16) unsigned char msm[];
    uint64_t* sbuf = reinterpret_cast<uint64_t*>(msm);                     // [MERGE_SORT_CAP]
    float* qs = reinterpret_cast<float*>(sbuf + MERGE_SORT_CAP);           // [d]
    const int kcp = next_pow2(a.kc);
    uint64_t* rkey = reinterpret_cast<uint64_t*>(qs + ((a.d + 3) & ~3));   // [kcp] (distkey<<32 | slot)
    int64_t* rid = reinterpret_cast<int64_t*>(rkey + kcp);                 // [kcp]
    uint32_t* rrow = reinterpret_cast<uint32_t*>(rid + kcp);               // [kcp]
    __shared__ int s_n;
    __shared__ double s_qn;

    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x;
    const float inf_pad = kIP ? -INFINITY : INFINITY;
    if (tid == 0) { s_n = 0; }
    for (int i = tid; i < a.d; i += blockDim.x) qs[i] = a.queries[q * a.q_pitch + i];
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < a.d; ++i) s += (double)qs[i] * (double)qs[i];
        s_qn = s;
    }

    // ---- gather survivors: emitted candidates whose filter key is within the final threshold
    const uint32_t gthr = a.gthr[q];
    bool overflow = false;
    for (int j = tid; j < a.P; j += blockDim.x) {
        const int seg = a.pair_seg[q * a.P + j];
        if (seg < 0) continue;
        const int64_t pair = q * a.P + j;
        const int n = a.cand_n[pair];
        const uint64_t r0 = (uint64_t)a.seg_row0[seg];
        const uint64_t* c = a.cand + (size_t)pair * a.kc;
        for (int i = 0; i < n; ++i) {
            const uint64_t v = c[i];
            const uint32_t key = (uint32_t)(v >> 32);
            if (key > gthr) break;  // runs are sorted ascending
            const int pos = atomicAdd(&s_n, 1);
            if (pos < MERGE_SORT_CAP) sbuf[pos] = ((uint64_t)key << 32) | (uint32_t)(r0 + (uint32_t)v);
            else overflow = true;
        }
    }
    overflow = __syncthreads_or(overflow);
    int ns = s_n;
    bool rescan = overflow || a.force_rescan;
    int nc = 0;
    if (!rescan) {
        const int np = next_pow2(ns > 1 ? ns : 1);
        for (int i = ns + tid; i < np; i += blockDim.x) sbuf[i] = COMP_MAX;
        block_bitonic_sort(sbuf, np, [](uint64_t x, uint64_t y) { return x < y; });
        nc = ns < a.kc ? ns : a.kc;
        // ---- exact refine in the reference's summation order
        for (int i = tid; i < kcp; i += blockDim.x) {
            if (i < nc) {
                const uint32_t row = (uint32_t)sbuf[i];
                const float dist = ref_pair_distance<kIP>(qs, a.vecs + (int64_t)row * a.pitch, a.d);
                // order by the value the reference orders by: sqrt'ed for l2 (list_scanning.h:260)
                const uint32_t dk = f2key(kIP ? -dist : (a.rank_squared ? dist : __fsqrt_rn(dist)));
                rkey[i] = ((uint64_t)dk << 32) | (uint32_t)i;
                rid[i] = a.ids ? a.ids[row] : (int64_t)row;
                rrow[i] = row;
            } else {
                rkey[i] = COMP_MAX;
            }
        }
        const int64_t* ridc = rid;
        block_bitonic_sort(rkey, kcp, [ridc](uint64_t x, uint64_t y) {
            const uint32_t dx = (uint32_t)(x >> 32), dy = (uint32_t)(y >> 32);
            if (dx != dy) return dx < dy;
            if (x == COMP_MAX || y == COMP_MAX) return x < y;
            return ridc[(uint32_t)x] < ridc[(uint32_t)y];
        });
        // ---- proof that nothing outside the refined set can enter the top-k
        if (ns >= a.kc && nc >= 1) {  // ns < kc: every probed row was a survivor, nothing was rejected
            if (tid == 0) {
                const int kk = a.k < nc ? a.k : nc;
                const float a_score = key2f((uint32_t)(sbuf[a.kc - 1] >> 32));  // filter score of the kc-th candidate
                const float rk = key2f((uint32_t)(rkey[kk - 1] >> 32));         // exact k-th (l2: distance, ip: -ip)
                const double qn = s_qn, qnorm = sqrt(qn), U = (double)a.max_row_norm;
                const double eps = 5.960464477539063e-08;  // 2^-24
                const double gam = (a.d + 8) * eps;
                const double e2 = (a.d / 8 + 12) * eps;
                bool ok;
                if (!kIP) {
                    const double e1 = gam * (U * U + 2.0 * qnorm * U) + 4.0 * eps * fabs((double)a_score);
                    // lb bounds the reference-order SQUARED distance of every rejected row from below; the
                    // reference compares sqrt'ed values, and sqrt_rn is monotone, so a rejected row cannot
                    // tie or beat the k-th as soon as sqrt_rn(round_down(lb)) is strictly above it.
                    const double lb = (qn * (1.0 - gam) + (double)a_score - e1) * (1.0 - e2);
                    const float lbf = __double2float_rd(lb);
                    ok = a.rank_squared ? (lb > (double)rk) : (lbf > 0.f && __fsqrt_rn(lbf) > rk);
                } else {
                    // scores are -<q,v>: any rejected v has ip <= -a_score + err; need that below the k-th exact ip
                    const double err = (gam + e2) * qnorm * U + 4.0 * eps * fabs((double)a_score);
                    ok = (-(double)a_score + err) < -(double)rk;
                }
                s_n = ok ? 0 : -1;
            }
            __syncthreads();
            rescan = (s_n < 0);
        }
    }
    if (rescan) {
        if (tid == 0) {
            a.flags[q] = 1;
            atomicAdd(&a.ctrl[2], 1);
        }
        return;
    }
    if (tid == 0) a.flags[q] = 0;
    for (int i = tid; i < a.k; i += blockDim.x) {
        int64_t id = -1, row = -1;
        float dist = inf_pad;
        if (i < nc) {
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
    }
}

// ------------------------------------------------------------------------------------------------
// 6. exhaustive exact re-scan of flagged queries (radix select on exact distance keys)
// ------------------------------------------------------------------------------------------------
template <bool kIP>
__global__ void __launch_bounds__(256) exact_rescan_kernel(const MergeArgs a, const int32_t* __restrict__ seg_rows) {
    const int64_t q = blockIdx.x;
    if (a.flags[q] == 0) return;
    extern __shared__ __align__(16) unsigned char msm[];
    float* qs = reinterpret_cast<float*>(msm);                                  // [d]
    const int kp = next_pow2(a.k);
    uint64_t* rkey = reinterpret_cast<uint64_t*>(qs + ((a.d + 3) & ~3) + 2);     // [kp]
    rkey = reinterpret_cast<uint64_t*>(((uintptr_t)rkey + 7) & ~(uintptr_t)7);
    int64_t* rid = reinterpret_cast<int64_t*>(rkey + kp);                       // [kp]
    uint32_t* rrow = reinterpret_cast<uint32_t*>(rid + kp);                     // [kp]
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_need, s_less, s_eq_total;
    __shared__ unsigned long long s_prefix64;
    __shared__ int s_scan[256];
    __shared__ int s_base_lt, s_base_eq;
    const int tid = threadIdx.x;
    const float inf_pad = kIP ? -INFINITY : INFINITY;

    for (int i = tid; i < a.d; i += blockDim.x) qs[i] = a.queries[q * a.q_pitch + i];
    int64_t total = 0;
    for (int j = 0; j < a.P; ++j) {
        const int seg = a.pair_seg[q * a.P + j];
        if (seg >= 0) total += seg_rows[seg];
    }
    const int kk = (int)(total < a.k ? total : a.k);
    if (tid == 0) { s_prefix = 0; s_need = kk; s_less = 0; }
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
            hist[tid] = 0;
            __syncthreads();
            const unsigned prefix = s_prefix;
            for (int j = 0; j < a.P; ++j) {
                const int seg = a.pair_seg[q * a.P + j];
                if (seg < 0) continue;
                const int64_t r0 = a.seg_row0[seg];
                const int n = seg_rows[seg];
                for (int r = tid; r < n; r += blockDim.x) {
                    const uint32_t dk = dist_key(r0 + r);
                    const bool match = (pass == 3) || ((dk >> (8 * (pass + 1))) == (prefix >> (8 * (pass + 1))));
                    if (match) atomicAdd(&hist[(dk >> (8 * pass)) & 255u], 1u);
                }
            }
            __syncthreads();
            if (tid == 0) {
                unsigned need = s_need, cum = 0;
                int b = 0;
                for (; b < 256; ++b) {
                    if (cum + hist[b] >= need) break;
                    cum += hist[b];
                }
                s_prefix = prefix | ((unsigned)b << (8 * pass));
                s_need = need - cum;
                s_less += cum;
                s_eq_total = b < 256 ? hist[b] : 0u;
            }
            __syncthreads();
        }
        const uint32_t T = s_prefix;
        const unsigned n_less = s_less;        // keys strictly below T
        const unsigned take_eq = s_need;       // how many keys equal to T to take
        const unsigned eq_total = s_eq_total;  // how many keys equal T
        __syncthreads();
        // A distance tie that straddles the k-th boundary is resolved by ascending id (the order the
        // oracle fixes for the reference's distance-only comparator): radix-select the take_eq-th
        // smallest id among the rows whose key equals T.
        uint64_t id_thr = ~0ull;
        if (eq_total > take_eq) {
            if (tid == 0) { s_prefix64 = 0ull; s_need = take_eq; }
            __syncthreads();
            for (int pass = 7; pass >= 0; --pass) {
                hist[tid] = 0;
                __syncthreads();
                const uint64_t prefix = s_prefix64;
                for (int j = 0; j < a.P; ++j) {
                    const int seg = a.pair_seg[q * a.P + j];
                    if (seg < 0) continue;
                    const int64_t r0 = a.seg_row0[seg];
                    const int n = seg_rows[seg];
                    for (int r = tid; r < n; r += blockDim.x) {
                        if (dist_key(r0 + r) != T) continue;
                        const uint64_t ik = id_key(r0 + r);
                        const bool match = (pass == 7) || ((ik >> (8 * (pass + 1))) == (prefix >> (8 * (pass + 1))));
                        if (match) atomicAdd(&hist[(unsigned)(ik >> (8 * pass)) & 255u], 1u);
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    unsigned need = s_need, cum = 0;
                    int b = 0;
                    for (; b < 256; ++b) {
                        if (cum + hist[b] >= need) break;
                        cum += hist[b];
                    }
                    s_prefix64 = prefix | ((unsigned long long)(b & 255) << (8 * pass));
                    s_need = need - cum;
                }
                __syncthreads();
            }
            id_thr = s_prefix64;
        }
        if (tid == 0) { s_base_lt = 0; s_base_eq = 0; }
        __syncthreads();
        // deterministic ordered collection
        for (int j = 0; j < a.P; ++j) {
            const int seg = a.pair_seg[q * a.P + j];
            if (seg < 0) continue;
            const int64_t r0 = a.seg_row0[seg];
            const int n = seg_rows[seg];
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
                s_scan[tid] = v;
                __syncthreads();
                for (int o = 1; o < 256; o <<= 1) {
                    int t = (tid >= o) ? s_scan[tid - o] : 0;
                    __syncthreads();
                    s_scan[tid] += t;
                    __syncthreads();
                }
                const int incl = s_scan[tid];
                const int excl = incl - v;
                const int blt = s_base_lt, beq = s_base_eq;
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
                if (tid == 255) {
                    s_base_lt = blt + (incl & 0xffff);
                    s_base_eq = beq + (incl >> 16);
                }
                __syncthreads();
            }
        }
    }
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
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int g_scan_variant = -1;
static int g_force_rescan = 0;

// optional per-launch timing of the filter kernel (qk_profile_*): CUDA events recorded on the launch
// stream right around scan_kernel, read back by the caller after it has synchronised.
struct ProfileRecord {
    cudaEvent_t start, stop;
    int64_t queries;
    int nprobe, k, used;
};
static ProfileRecord* g_prof = nullptr;
static int g_prof_cap = 0, g_prof_n = 0;

static int launch_scan(const ScanArgs& sa, int metric, size_t smem, cudaStream_t stream) {
    const int grid = sm_count();
    if (metric == QK_METRIC_INNER_PRODUCT) {
        auto kern = scan_kernel<true>;
        QK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, SCAN_THREADS, smem, stream>>>(sa);
    } else {
        auto kern = scan_kernel<false>;
        QK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, SCAN_THREADS, smem, stream>>>(sa);
    }
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}

}  // namespace qk

using namespace qk;

extern "C" size_t qk_scan_workspace_bytes(const qk_store_t* store, int64_t num_queries, int nprobe, int k) {
    ScanPlan p;
    if (!store || num_queries <= 0 || nprobe <= 0) return 0;
    if (make_plan(store, num_queries, nprobe, k, &p) != QK_OK) return 0;
    return p.total;
}

extern "C" int qk_scan_partitions(const qk_store_t* st, const float* queries, int64_t Q, int64_t q_pitch,
                                  const int32_t* probe_lists, int nprobe, int metric, int k, int64_t* out_ids,
                                  float* out_dist, int64_t* out_rows, void* workspace, size_t workspace_bytes,
                                  int32_t* stats, void* stream_v) {
    return qk::scan_partitions_impl(st, queries, Q, q_pitch, probe_lists, nprobe, metric, k, out_ids, out_dist, out_rows,
                                    workspace, workspace_bytes, stats, stream_v, 0);
}

int qk::scan_partitions_impl(const qk_store_t* st, const float* queries, int64_t Q, int64_t q_pitch,
                             const int32_t* probe_lists, int nprobe, int metric, int k, int64_t* out_ids,
                             float* out_dist, int64_t* out_rows, void* workspace, size_t workspace_bytes,
                             int32_t* stats, void* stream_v, int rank_squared) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(st && queries && probe_lists && out_ids && out_dist, "null argument");
    QK_REQUIRE(metric == QK_METRIC_L2 || metric == QK_METRIC_INNER_PRODUCT, "metric %d not supported", metric);
    QK_REQUIRE(Q > 0 && nprobe > 0, "empty query batch");
    QK_REQUIRE(st->num_segments > 0 && st->num_lists > 0, "store has no segments");
    ScanPlan p;
    int rc = make_plan(st, Q, nprobe, k, &p);
    if (rc) return rc;
    QK_REQUIRE(q_pitch >= p.dp && q_pitch % 4 == 0, "query pitch %lld must be a multiple of 4 and >= %d",
               (long long)q_pitch, p.dp);
    QK_REQUIRE(((uintptr_t)queries % 16 == 0) && ((uintptr_t)st->vectors % 16 == 0), "vectors must be 16-byte aligned");
    if (workspace_bytes < p.total || !workspace) {
        set_error("workspace too small: need %zu bytes, have %zu", p.total, workspace_bytes);
        return QK_ERR_WORKSPACE;
    }
    if (g_scan_variant < 0) {
        g_scan_variant = 0;
        const char* f = getenv("QK_FORCE_RESCAN");
        g_force_rescan = f ? atoi(f) : 0;
    }
    char* ws = (char*)workspace;
    int32_t* pair_seg = (int32_t*)(ws + p.off_pair_seg);
    int32_t* seg_count = (int32_t*)(ws + p.off_seg_count);
    int32_t* seg_fill = (int32_t*)(ws + p.off_seg_fill);
    int32_t* flags = (int32_t*)(ws + p.off_flags);
    int32_t* ctrl = (int32_t*)(ws + p.off_ctrl);
    int32_t* seg_start = (int32_t*)(ws + p.off_seg_start);
    int32_t* item_start = (int32_t*)(ws + p.off_item_start);
    int32_t* seg_pairs = (int32_t*)(ws + p.off_seg_pairs);
    int2* items = (int2*)(ws + p.off_items);
    uint32_t* gthr = (uint32_t*)(ws + p.off_gthr);
    int32_t* cand_n = (int32_t*)(ws + p.off_cand_n);
    uint64_t* cand = (uint64_t*)(ws + p.off_cand);
    const int S = st->num_segments;
    const int64_t QP = Q * p.P;

    // seg_count, seg_fill, flags, ctrl are contiguous: one memset
    QK_CUDA(cudaMemsetAsync(ws + p.off_seg_count, 0, p.off_seg_start - p.off_seg_count, stream));
    const bool single = (p.P == nprobe);
    if (single) {
        int64_t n = Q * nprobe;
        expand_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(probe_lists, Q, nprobe, p.P, st->list_seg0,
                                                                              st->list_nseg, st->num_lists, pair_seg,
                                                                              seg_count, gthr, true);
    } else {
        expand_pairs_kernel<<<(unsigned)((Q + 127) / 128), 128, 0, stream>>>(probe_lists, Q, nprobe, p.P, st->list_seg0,
                                                                              st->list_nseg, st->num_lists, pair_seg,
                                                                              seg_count, gthr, false);
    }
    QK_CUDA(cudaGetLastError());
    prefix_segments_kernel<<<1, 1024, 0, stream>>>(seg_count, S, p.gq, seg_start, item_start, ctrl);
    QK_CUDA(cudaGetLastError());
    {
        int64_t n = QP > S ? QP : S;
        scatter_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(pair_seg, QP, S, seg_start, seg_fill,
                                                                               seg_pairs, item_start, items);
        QK_CUDA(cudaGetLastError());
    }
    ScanArgs sa;
    sa.vecs = st->vectors; sa.pitch = st->pitch; sa.dp = p.dp;
    sa.seg_row0 = st->seg_row0; sa.seg_rows = st->seg_rows;
    sa.queries = queries; sa.q_pitch = q_pitch;
    sa.seg_start = seg_start; sa.seg_pairs = seg_pairs; sa.items = items; sa.ctrl = ctrl;
    sa.gthr = gthr; sa.cand = cand; sa.cand_n = cand_n;
    sa.P = p.P; sa.kc = p.kc; sa.gq = p.gq; sa.nq = p.nq;
    sa.norms = st->row_norms;
    ProfileRecord* rec = nullptr;
    if (g_prof && g_prof_n < g_prof_cap) {
        rec = &g_prof[g_prof_n++];
        rec->queries = Q; rec->nprobe = nprobe; rec->k = k; rec->used = 1;
        QK_CUDA(cudaEventRecord(rec->start, stream));
    }
    rc = launch_scan(sa, metric, p.smem, stream);
    if (rc) return rc;
    if (rec) QK_CUDA(cudaEventRecord(rec->stop, stream));

    MergeArgs ma;
    ma.vecs = st->vectors; ma.pitch = st->pitch; ma.ids = st->ids; ma.d = st->d;
    ma.seg_row0 = st->seg_row0; ma.queries = queries; ma.q_pitch = q_pitch;
    ma.pair_seg = pair_seg; ma.gthr = gthr; ma.cand = cand; ma.cand_n = cand_n;
    ma.flags = flags; ma.ctrl = ctrl; ma.P = p.P; ma.kc = p.kc; ma.k = k;
    ma.max_row_norm = st->max_row_norm;
    ma.out_ids = out_ids; ma.out_dist = out_dist; ma.out_rows = out_rows;
    ma.force_rescan = g_force_rescan;
    ma.rank_squared = rank_squared;
    {
        int kcp = 1;
        while (kcp < p.kc) kcp <<= 1;
        size_t msmem = MERGE_SORT_CAP * 8 + (size_t)((st->d + 3) & ~3) * 4 + (size_t)kcp * (8 + 8 + 4) + 16;
        if (metric == QK_METRIC_INNER_PRODUCT) {
            QK_CUDA(cudaFuncSetAttribute(merge_refine_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
            merge_refine_kernel<true><<<(unsigned)Q, MERGE_THREADS, msmem, stream>>>(ma);
        } else {
            QK_CUDA(cudaFuncSetAttribute(merge_refine_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
            merge_refine_kernel<false><<<(unsigned)Q, MERGE_THREADS, msmem, stream>>>(ma);
        }
        QK_CUDA(cudaGetLastError());
        int kp = 1;
        while (kp < k) kp <<= 1;
        size_t rsmem = (size_t)((st->d + 3) & ~3) * 4 + 32 + (size_t)kp * (8 + 8 + 4);
        if (metric == QK_METRIC_INNER_PRODUCT) {
            QK_CUDA(cudaFuncSetAttribute(exact_rescan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
            exact_rescan_kernel<true><<<(unsigned)Q, 256, rsmem, stream>>>(ma, st->seg_rows);
        } else {
            QK_CUDA(cudaFuncSetAttribute(exact_rescan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
            exact_rescan_kernel<false><<<(unsigned)Q, 256, rsmem, stream>>>(ma, st->seg_rows);
        }
        QK_CUDA(cudaGetLastError());
    }
    if (stats) QK_CUDA(cudaMemcpyAsync(stats, ctrl + 2, sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
    return QK_OK;
}

// ---- per-launch timing of the filter kernel ------------------------------------------------------
extern "C" int qk_profile_begin(int max_records) {
    QK_REQUIRE(max_records > 0 && max_records <= (1 << 20), "bad max_records");
    if (g_prof) {
        for (int i = 0; i < g_prof_cap; ++i) { cudaEventDestroy(g_prof[i].start); cudaEventDestroy(g_prof[i].stop); }
        delete[] g_prof;
        g_prof = nullptr;
    }
    g_prof = new ProfileRecord[max_records];
    g_prof_cap = max_records;
    g_prof_n = 0;
    for (int i = 0; i < max_records; ++i) {
        QK_CUDA(cudaEventCreate(&g_prof[i].start));
        QK_CUDA(cudaEventCreate(&g_prof[i].stop));
        g_prof[i].used = 0;
    }
    return QK_OK;
}

extern "C" int qk_profile_count(void) { return g_prof ? g_prof_n : 0; }

extern "C" int qk_profile_read(int index, float* ms, int64_t* queries, int* nprobe, int* k) {
    QK_REQUIRE(g_prof && index >= 0 && index < g_prof_n, "profile record %d out of range", index);
    ProfileRecord& r = g_prof[index];
    QK_CUDA(cudaEventSynchronize(r.stop));
    float t = 0.f;
    QK_CUDA(cudaEventElapsedTime(&t, r.start, r.stop));
    if (ms) *ms = t;
    if (queries) *queries = r.queries;
    if (nprobe) *nprobe = r.nprobe;
    if (k) *k = r.k;
    return QK_OK;
}

extern "C" int qk_profile_end(void) {
    if (g_prof) {
        for (int i = 0; i < g_prof_cap; ++i) { cudaEventDestroy(g_prof[i].start); cudaEventDestroy(g_prof[i].stop); }
        delete[] g_prof;
    }
    g_prof = nullptr;
    g_prof_cap = g_prof_n = 0;
    return QK_OK;
}
