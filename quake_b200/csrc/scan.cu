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
#include <cstring>
#include <cstdlib>
#include <mutex>

namespace qk {

static int g_scan_variant = -1;  // 0: tensor-core filter for d <= 128 (default), 1: FP32-pipe filter everywhere
static int g_force_rescan = 0;
static int g_filter_terms_env = 0;  // QK_FILTER_TERMS: overrides the store's setting (experiments)
// relative error bound of the tensor-core filter's dot products against |q||v|: 3xTF32 measured ~2^-20 of sum|q_i v_i|
// (scripts/umma_probe.cu), bounded by 2^-17; 2xTF32 drops a_lo = a - tf32(a), |a_lo| < 2^-10 |a| element-wise
static double filter_gamma(int terms) { return terms == 2 ? 9.765625e-04 + 7.62939453125e-06 : 7.62939453125e-06; }
static void read_scan_env() {
    if (g_scan_variant >= 0) return;
    const char* e = getenv("QK_SCAN_PATH");
    g_scan_variant = (e && strcmp(e, "ffma") == 0) ? 1 : 0;
    const char* f = getenv("QK_FORCE_RESCAN");
    g_force_rescan = f ? atoi(f) : 0;
    const char* t = getenv("QK_FILTER_TERMS");
    g_filter_terms_env = t ? atoi(t) : 0;
}

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
        off_ctrl, off_gthr, off_qdelta, off_flags, off_qcount, off_qbuf, total;
    int qcap;     // entries of the per-query candidate buffer
    int qstride;  // ints between the fill counters of consecutive queries (see plan_qstride)
    int sample;   // rows sampled per query for the threshold seed (0: no seeding)
    int flat_seed;  // every query samples the same rows (single-list store): seed scores as one small GEMM
    size_t off_skeys;
    int dense;    // flat store, small enough: the filter writes every score, one select per query replaces thresholds
    size_t off_dense;
};

static constexpr int SCAN_DC = 128;     // floats of a row staged per pipeline unit
static constexpr int SCAN_STAGES = 4;
static constexpr int SCAN_GQ = 32;      // max queries per work item
static constexpr int SCAN_TV = 64;      // rows per tile
static constexpr int SCAN_KP = SCAN_TV + 4;  // padded key row stride (u32)
static constexpr int SCAN_COMPUTE_WARPS = 8;  // two groups of four
static constexpr int SCAN_SELECT_WARPS = 7;   // 8 + 7 + 1 producer = 16 warps: 128 registers per thread
static constexpr int SCAN_THREADS = 32 * (SCAN_COMPUTE_WARPS + SCAN_SELECT_WARPS + 1);
static constexpr int SCAN_MAX_NQ = 4;   // query-chunk ring depth
static constexpr size_t SCAN_SMEM_LIMIT = 227 * 1024;

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// The per-query fill counters take one returning atomicAdd per appended candidate (~500 k per batch at C2). L2 atomics
// on one 128-byte line are serialised by its slice: with the counters packed (32 per line) all of them funnelled
// through 32 lines. One counter per line (per 32-byte sector for very large batches, where the clear would cost more
// than the contention).
static int plan_qstride(int64_t Q) {
    static const int env = getenv("QK_QSTRIDE") ? atoi(getenv("QK_QSTRIDE")) : 0;  // experiments
    if (env > 0) return env;
    return Q <= 16384 ? 32 : (Q <= 262144 ? 8 : 1);
}

static int candidate_count(int k) { return k + (k / 16 > 6 ? k / 16 : 6); }

struct WorkItem {  // one (segment, query chunk) unit of work; written by scatter_pairs_kernel
    int seg;     // -1: no more work (only in shared memory)
    int g_begin, g_cnt;
    int nrows;
    long long row0;
    long long pad_;
};
static_assert(sizeof(WorkItem) == 32, "WorkItem layout");

struct ItemDesc {  // published in shared memory by the producer warp for every work item in flight
    WorkItem w;
    int pair[SCAN_GQ];        // (query, segment) pair index of every query slot
    uint32_t gthr[SCAN_GQ];   // the queries' global filter-key thresholds when the item was issued
};
static constexpr int SCAN_SMEM_HEADER = 2048;  // mbarriers (256 B) + SCAN_MAX_NQ descriptors
static_assert(256 + SCAN_MAX_NQ * sizeof(ItemDesc) <= SCAN_SMEM_HEADER, "header");

// Entries of the per-query candidate buffer in global memory. With an exact running threshold a query appends
// about kc * (1 + ln(rows / kc)) rows over a whole scan; the threshold is refreshed every SCAN_REFRESH_STEP-ish
// appends and several pairs of one query can be in flight, hence the slack. Overflow is not an error: the query
// is handed to the exact re-scan.
static int candidate_buffer_cap(int kc, double rows_per_query) {
    // small k: thresholds converge within a few hundred appends. Large k needs many more rows before a useful
    // threshold exists (the kc-th best of what was seen so far) and floods the refresh warps meanwhile: give it room
    // (8 B per entry) rather than pay the exact re-scan.
    int c = kc <= 32 ? 2048 : 256 * kc;
    // long scans: the appends of a query grow with the rows it scans (measured at k = 10: mean 534 / max ~1900 of
    // 15.6k rows, mean 609 / max > 2048 now and then of 62k rows) and ONE overflowing query costs an exact re-scan of
    // all of them (~1.4 ms for 62k rows): scale the buffer with the expected rows per query
    if (rows_per_query > 16384.0) {
        double f = rows_per_query / 16384.0;
        if (f > 16.0) f = 16.0;
        c = (int)(c * f) / 1024 * 1024;
    }
    static const int env_cap = getenv("QK_QCAP") ? atoi(getenv("QK_QCAP")) : 0;  // experiments
    if (env_cap > 0) c = env_cap;
    if (c > 65536) c = 65536;
    if (c < 32 * kc) c = 32 * kc;
    if (c > 131072) c = 131072;
    return c;
}

static size_t scan_smem_bytes(int dp, int kc, int gq, int nq) {
    size_t b = SCAN_SMEM_HEADER;                                             // mbarriers + item descriptors
    b += (size_t)SCAN_STAGES * SCAN_TV * SCAN_DC * sizeof(float);            // row ring (TMA, swizzled)
    b += (size_t)nq * gq * (dp + 4) * sizeof(float);                         // query-chunk ring
    b += (size_t)4 * gq * SCAN_KP * sizeof(uint32_t);                        // score keys: 2 groups x 2 buffers
    b += (size_t)SCAN_SELECT_WARPS * 256 * sizeof(uint32_t);                 // radix-select histograms
    (void)kc;
    return b + 1024;                                                         // slack to align the base to 1 KB
}

static int make_plan(const qk_store_t* st, int64_t Q, int nprobe, int k, ScanPlan* p, bool allow_dense = true) {
    read_scan_env();
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
    // enough work items for every SM: shrink the query chunk (down to 8) while the batch yields fewer than
    // two items per SM (a coarse scan has one pair per query)
    {
        const int64_t est_pairs = Q * (int64_t)p->P;
        while (gq > 8 && (est_pairs + gq - 1) / gq < 2 * (int64_t)sm_count()) gq >>= 1;
    }
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
    p->qstride = plan_qstride(Q);
    p->off_qcount = o;     o = align_up(o + (size_t)Q * p->qstride * 4, 256);
    p->off_ctrl = o;       o = align_up(o + 64, 256);
    p->off_seg_start = o;  o = align_up(o + (S + 1) * 4, 256);
    p->off_item_start = o; o = align_up(o + (S + 1) * 4, 256);
    p->off_seg_pairs = o;  o = align_up(o + QP * 4, 256);
    size_t max_items = QP / gq + (QP < S ? QP : S) + 1;
    p->off_items = o;      o = align_up(o + max_items * sizeof(WorkItem), 256);
    p->off_gthr = o;       o = align_up(o + (size_t)Q * 4, 256);
    p->off_qdelta = o;     o = align_up(o + (size_t)Q * 4, 256);
    {
        const double mean_list = st->num_lists > 0 ? (double)st->num_rows / st->num_lists : 0.0;
        p->qcap = candidate_buffer_cap(p->kc, st->num_lists == 1 ? 0.0 : mean_list * nprobe);  // single-list stores: one seed sample fits all rows
    }
    if (st->max_segment_rows > 0) {
        // a query never appends more rows than it scans
        const int64_t bound = ((int64_t)p->P * st->max_segment_rows + 63) / 64 * 64;
        if (bound < p->qcap) p->qcap = (int)(bound < 64 ? 64 : bound);
    }
    p->off_qbuf = o;       o = align_up(o + (size_t)Q * p->qcap * 8, 256);
    // threshold seeds from a sample of the query's first probed rows: about 8 * kc of them (pass rate of the seed
    // ~ 1/8), bounded by the read traffic it costs -- unless the whole store is small enough to sit in L2 (a flat
    // index / the centroid list, shared by every query)
    {
        int sample = 32;
        while (sample < 1024 && sample < 8 * p->kc) sample <<= 1;
        const size_t store_bytes = (size_t)st->num_rows * st->pitch * sizeof(float);
        if (store_bytes > ((size_t)32 << 20)) {
            // budget: 96 MB, or a tenth of what the scan itself is expected to stream for this batch
            const double lists = (double)(Q * (int64_t)nprobe < st->num_lists ? Q * (int64_t)nprobe : st->num_lists);
            const double scan_bytes = lists * ((double)st->num_rows / (st->num_lists > 0 ? st->num_lists : 1)) * p->dp * 4.0;
            size_t budget = (size_t)96 << 20;
            if (scan_bytes / 10.0 > (double)budget) budget = (size_t)(scan_bytes / 10.0);
            while (sample > 32 && sample / 2 >= p->kc && (size_t)Q * sample * p->dp * sizeof(float) > budget) sample >>= 1;
        }
        p->flat_seed = (st->num_lists == 1 && nprobe == 1) ? 1 : 0;
        static const int env_sample = getenv("QK_SEED_SAMPLE") ? atoi(getenv("QK_SEED_SAMPLE")) : -1;  // experiments
        if (env_sample >= 0 && st->num_lists > 1) sample = env_sample;
        if (sample < p->kc) sample = 0;
        // dense mode: [Q x rows] keys, at most 512 MB and at most 32768 rows (the select keeps a query's keys in smem).
        // Only a flat-mode call (no probe table) uses it; the workspace is sized for either.
        static const bool no_dense = getenv("QK_NO_DENSE") != nullptr;
        const bool dense_ok = p->flat_seed && g_scan_variant == 0 && p->dp <= 128 && Q <= 8192 && st->flat_rows >= p->kc &&
                              st->flat_rows <= 32768 && (size_t)Q * (size_t)st->flat_rows * 4 <= ((size_t)512 << 20) && !no_dense;
        p->dense = (dense_ok && allow_dense) ? 1 : 0;
        p->sample = p->dense ? 0 : sample;
        p->off_skeys = o;
        if (p->flat_seed && sample) o = align_up(o + (size_t)Q * sample * 4, 256);
        p->off_dense = o;
        if (dense_ok) o = align_up(o + (size_t)Q * (size_t)st->flat_rows * 4, 256);
    }
    p->total = o;
    return QK_OK;
}

// ------------------------------------------------------------------------------------------------
// 1. expand (query, list) -> (query, segment) pairs + histogram
// ------------------------------------------------------------------------------------------------
struct ProbeSource {
    const int32_t* slots;       // [Q x nprobe] list slots, or null:
    const int64_t* ids;         // [Q x nprobe] partition ids (the coarse scan's output) mapped through ...
    const int32_t* id_to_slot;  // ... this dense table (ids outside it, or mapped to < 0, are skipped)
    int64_t table_size;
    int shard_rank, shard_world;  // shard_world > 1: only partitions with id % shard_world == shard_rank are scanned
};
__device__ __forceinline__ int probe_slot(const ProbeSource& ps, int64_t i) {
    if (ps.slots) return ps.slots[i];
    const int64_t id = ps.ids[i];
    if (id < 0 || id >= ps.table_size) return -1;
    if (ps.shard_world > 1 && (int)(id % ps.shard_world) != ps.shard_rank) return -1;
    return ps.id_to_slot[id];
}

__global__ void expand_pairs_kernel(const ProbeSource probe, int64_t Q, int nprobe, int P,
                                    const int32_t* __restrict__ list_seg0, const int32_t* __restrict__ list_nseg,
                                    int num_lists, int32_t* __restrict__ pair_seg, int32_t* __restrict__ seg_count,
                                    uint32_t* __restrict__ gthr, bool single_segment_lists, bool reset_thresholds) {
    if (single_segment_lists) {
        // one thread per (query, probe)
        int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= Q * nprobe) return;
        int64_t q = i / nprobe;
        int j = (int)(i - q * nprobe);
        if (j == 0 && reset_thresholds) gthr[q] = KEY_MAX;
        int l = probe_slot(probe, i);
        int seg = -1;
        if (l >= 0 && l < num_lists && list_nseg[l] > 0) seg = list_seg0[l];
        pair_seg[q * P + j] = seg;  // P == nprobe here
        if (seg >= 0) atomicAdd(&seg_count[seg], 1);
    } else if (nprobe == 1) {
        // one probed list per query: one thread per (query, segment slot)
        int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= Q * P) return;
        int64_t q = i / P;
        int s2 = (int)(i - q * P);
        if (s2 == 0 && reset_thresholds) gthr[q] = KEY_MAX;
        int l = probe_slot(probe, q);
        int seg = -1;
        if (l >= 0 && l < num_lists && s2 < list_nseg[l]) seg = list_seg0[l] + s2;
        pair_seg[i] = seg;
        if (seg >= 0) atomicAdd(&seg_count[seg], 1);
    } else {
        // one thread per query, sequential over its probes
        int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (q >= Q) return;
        if (reset_thresholds) gthr[q] = KEY_MAX;
        int pos = 0;
        for (int j = 0; j < nprobe; ++j) {
            int l = probe_slot(probe, q * nprobe + j);
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
    pdl_launch_dependents();
    pdl_wait();
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
                                     const int64_t* __restrict__ seg_row0, const int32_t* __restrict__ seg_rows, int gq,
                                     WorkItem* __restrict__ items) {
    pdl_wait();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < QP) {
        int seg = pair_seg[i];
        if (seg >= 0) {
            int pos = seg_start[seg] + atomicAdd(&seg_fill[seg], 1);
            seg_pairs[pos] = (int32_t)i;
        }
    }
    if (i < S) {
        const int b = item_start[i], e = item_start[i + 1];
        if (e > b) {
            const int p0 = seg_start[i], p1 = seg_start[i + 1];
            WorkItem w;
            w.seg = (int)i;
            w.nrows = seg_rows[i];
            w.row0 = seg_row0[i];
            w.pad_ = 0;
            for (int c = b; c < e; ++c) {
                w.g_begin = p0 + (c - b) * gq;
                w.g_cnt = min(gq, p1 - w.g_begin);
                items[c] = w;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// running per-query filter threshold
// ------------------------------------------------------------------------------------------------
// kc-th smallest 32-bit filter key among buf[0..n) (composite entries key << 32 | row, read past L1), by a
// warp-level radix select: 4 passes, 256-bin histogram in shared memory. Slots that were reserved but not
// yet written hold 0xff..ff and count as +inf, so the result is always a valid upper bound on the kc-th
// smallest key of everything appended so far. Returns KEY_MAX while fewer than kc entries are visible.
template <typename KeyAt>
__device__ __forceinline__ uint32_t radix_select(KeyAt key_at, int n, int kc, uint32_t* hist, int lane) {
    if (n < kc) return KEY_MAX;
    uint32_t prefix = 0;
    int need = kc;  // rank (1-based) of the wanted key among the keys matching `prefix` so far
    for (int pass = 3; pass >= 0; --pass) {
#pragma unroll
        for (int b = 0; b < 8; ++b) hist[lane * 8 + b] = 0;
        __syncwarp();
        const int sh = 8 * pass;
        for (int i = lane; i < n; i += 32) {
            const uint32_t key = key_at(i);
            const bool match = (pass == 3) || ((key >> (sh + 8)) == (prefix >> (sh + 8)));
            if (match) atomicAdd(&hist[(key >> sh) & 255u], 1u);
        }
        __syncwarp();
        uint32_t h[8], sum = 0;  // lane l owns bins 8l..8l+7
#pragma unroll
        for (int b = 0; b < 8; ++b) { h[b] = hist[lane * 8 + b]; sum += h[b]; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const uint32_t excl = incl - sum;
        const unsigned owner = __ballot_sync(0xffffffffu, incl >= (uint32_t)need);
        const int ol = __ffs(owner) - 1;  // first lane whose cumulative count reaches `need`
        int bin = 0;
        uint32_t before = excl;
        if (lane == ol) {
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                if (before + h[b] >= (uint32_t)need) { bin = lane * 8 + b; break; }
                before += h[b];
            }
        }
        bin = __shfl_sync(0xffffffffu, bin, ol);
        before = __shfl_sync(0xffffffffu, before, ol);
        prefix |= (uint32_t)bin << sh;
        need -= (int)before;
        __syncwarp();
    }
    return prefix;
}

// ------------------------------------------------------------------------------------------------
// 3b. threshold seeds
// ------------------------------------------------------------------------------------------------
// One CTA per query: the kc-th smallest filter score over a small sample of the query's probed rows (the first
// `sample` rows of its first probed lists, scored on the FP32 pipe) is an upper bound on the kc-th smallest over all
// of them, so the scan starts with a useful threshold instead of admitting everything. The sample is scored in
// different arithmetic than the scan kernel (plain FMA chain vs. split TF32 / FFMA2 tiles), so the bound is widened
// by more than the two error bounds together.
// The sampled rows are contiguous runs of the arena (one per probed list): they are fetched with TMA bulk copies into
// shared memory, SEED_CHUNK rows at a time, so the kernel is a stream of a few large asynchronous copies per SM instead
// of register-limited dependent loads (it was latency-bound: 39-63 us for 67 MB at C2).
// 128 threads and 24 KB of staging per CTA (three 8 KB buffers: two bulk copies in flight behind the chunk being
// scored): eight CTAs share an SM, so a batch of 1024 queries is ONE wave of 148 x 8 slots (with 256 threads / 32 KB it
// was 1.7 waves of four, each CTA mostly waiting for its own bulk copy: 41 us at C2). Plain 16-byte loads into
// registers, eight rows in flight per warp, were slower (32 us).
// Scoring: EIGHT LANES PER ROW, four rows per warp step -- lane (r, j) sums the 16-byte pieces j, j + 8, ... of row r
// (a quarter-warp reads 128 contiguous bytes: conflict-free) and three shuffles finish the row; one row per warp step
// with a five-shuffle reduction spent three times the instructions and the kernel was issue-bound (56 % issue-active).
static constexpr int SEED_THREADS = 128;
static constexpr int SEED_WARPS = SEED_THREADS / 32;
static constexpr int SEED_NBUF = 3;
static int seed_chunk_rows(int dp, int sample) {  // rows per staging buffer: at most 8 KB (16 rows at d = 128)
    int c = (8 * 1024) / (dp * 4);
    if (c < 4) c = 4;
    return sample < c ? sample : c;
}
static size_t seed_smem_bytes(int dp, int sample) {
    return (size_t)SEED_NBUF * seed_chunk_rows(dp, sample) * dp * 4 + (size_t)dp * 4 + (size_t)sample * 8 + 16;
}
template <bool kIP>
__global__ void __launch_bounds__(SEED_THREADS, 8) seed_thresholds_kernel(const float* __restrict__ vecs, int64_t pitch,
                                                              const float* __restrict__ norms, int d, int dp,
                                                              const float* __restrict__ queries, int64_t q_pitch, int64_t Q,
                                                              const int32_t* __restrict__ pair_seg, int P,
                                                              const int64_t* __restrict__ seg_row0,
                                                              const int32_t* __restrict__ seg_rows, int kc, int sample,
                                                              int chunk_rows, float max_row_norm, float rel_margin,
                                                              uint32_t* __restrict__ gthr) {
    extern __shared__ __align__(128) unsigned char seed_raw[];
    float* rows = reinterpret_cast<float*>(seed_raw);                 // [SEED_NBUF][chunk_rows][dp]: staging ring
    float* qs = rows + (size_t)SEED_NBUF * chunk_rows * dp;           // [dp]
    uint32_t* keys = reinterpret_cast<uint32_t*>(qs + dp);            // [sample]
    float* nrm = reinterpret_cast<float*>(keys + sample);             // [sample] squared norms of the sampled rows
    __shared__ int s_start[33];      // exclusive prefix of the sampled rows over the first 32 probes
    __shared__ long long s_r0[32];   // first arena row of each of them
    __shared__ __align__(8) uint64_t s_bar[SEED_NBUF];
    __shared__ uint32_t s_hist[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q = blockIdx.x;
    for (int i = tid; i < dp; i += SEED_THREADS) qs[i] = i < d ? queries[q * q_pitch + i] : 0.f;
    if (warp == 0) {
        int nrows = 0;
        long long r0 = 0;
        if (lane < P) {
            const int seg = pair_seg[q * P + lane];
            if (seg >= 0) { nrows = seg_rows[seg]; r0 = seg_row0[seg]; }
        }
        int incl = nrows;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        s_start[lane] = incl - nrows;
        s_r0[lane] = r0;
        if (lane == 31) s_start[32] = incl;
        if (lane == 0) {
#pragma unroll
            for (int b = 0; b < SEED_NBUF; ++b) mbar_init(&s_bar[b], 1);
            mbar_fence_init();
        }
    }
    __syncthreads();
    const int have = min(s_start[32], sample);
    const int dp4 = dp >> 2;
    const uint32_t row_bytes = (uint32_t)dp * 4u;
    const int nchunks = (have + chunk_rows - 1) / chunk_rows;
    // chunk c = rows [c * chunk_rows, ...) of the sample -> buffer c % SEED_NBUF: one bulk copy per probed list they
    // touch (contiguous when the arena has no row padding), else one per row. Issued by warp 0, SEED_NBUF - 1 chunks
    // ahead of the scoring.
    auto issue = [&](int c) {
        if (warp != 0) return;
        const int base = c * chunk_rows, cn = min(chunk_rows, have - base);
        uint64_t* bar = &s_bar[c % SEED_NBUF];
        float* buf = rows + (size_t)(c % SEED_NBUF) * chunk_rows * dp;
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)cn * row_bytes);
        __syncwarp();
        const int lo = max(s_start[lane], base), hi = min(s_start[lane + 1], base + cn);
        if (hi > lo) {
            const float* src = vecs + (s_r0[lane] + (lo - s_start[lane])) * pitch;
            float* dst = buf + (size_t)(lo - base) * dp;
            if (pitch == dp) {
                bulk_g2s(dst, src, (uint32_t)(hi - lo) * row_bytes, bar);
            } else {
                for (int r = 0; r < hi - lo; ++r) bulk_g2s(dst + (size_t)r * dp, src + (size_t)r * pitch, row_bytes, bar);
            }
        }
    };
#pragma unroll
    for (int c = 0; c < SEED_NBUF - 1; ++c)
        if (c < nchunks) issue(c);
    // the rows' squared norms, one coalesced load per sampled row, in flight together with the first bulk copies
    // (read one by one inside the scoring loop they were a chain of dependent cache misses)
    if (!kIP) {
        for (int i = tid; i < have; i += SEED_THREADS) {
            int j = 0;
            while (j < 31 && i >= s_start[j + 1]) ++j;
            nrm[i] = norms[s_r0[j] + (i - s_start[j])];
        }
    }
    __syncthreads();
    const int sub = lane >> 3, j8 = lane & 7;  // row of the warp's group of four, 16-byte piece within the row
    for (int c = 0; c < nchunks; ++c) {
        // the buffer of chunk c + SEED_NBUF - 1 held chunk c - 1: released by the barrier that ended iteration c - 1
        if (c + SEED_NBUF - 1 < nchunks) issue(c + SEED_NBUF - 1);
        const int base = c * chunk_rows, cn = min(chunk_rows, have - base);
        const float* buf = rows + (size_t)(c % SEED_NBUF) * chunk_rows * dp;
        mbar_wait(&s_bar[c % SEED_NBUF], (uint32_t)(c / SEED_NBUF) & 1u);
        for (int i0 = warp * 4; i0 < cn; i0 += SEED_WARPS * 4) {
            const int i = i0 + sub;
            const bool live = i < cn;
            const float4* rp = reinterpret_cast<const float4*>(buf + (size_t)(live ? i : i0) * dp);
            float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
            for (int cc = j8; cc < dp4; cc += 8) {
                const float4 x = rp[cc], y = reinterpret_cast<const float4*>(qs)[cc];
                a0 = fmaf(x.x, y.x, a0); a1 = fmaf(x.y, y.y, a1);
                a0 = fmaf(x.z, y.z, a0); a1 = fmaf(x.w, y.w, a1);
            }
            float acc = a0 + a1;
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            if (live && j8 == 0) keys[base + i] = f2key(kIP ? -acc : fmaf(-2.f, acc, nrm[base + i]));
        }
        __syncthreads();  // the buffer is refilled SEED_NBUF - 1 chunks later
    }
    __syncthreads();
    if (warp != 0 || have < kc) return;  // fewer sampled rows than candidates wanted: no bound
    // kc-th smallest key of the sample: warp-level radix select over the keys in shared memory (the bisection on key
    // bits this used to be -- 32 rounds of one ballot per 32 keys, by one warp -- was ~40 % of the kernel's samples)
    float qq = 0.f;
    for (int i = lane; i < dp; i += 32) qq = fmaf(qs[i], qs[i], qq);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
    const uint32_t* kk = keys;
    const uint32_t lo = radix_select([kk](int i) { return kk[i]; }, have, kc, s_hist, lane);
    if (lane == 0 && lo < KEY_MAX) {
        const float t = key2f(lo);
        const float margin = __fadd_ru(__fmul_ru(rel_margin, __fmul_ru(sqrtf(qq), max_row_norm)), fabsf(t) * 9.5367431640625e-07f);
        const float tm = __fadd_ru(t, margin);
        if (tm == tm) atomicMin(gthr + q, f2key(tm));
    }
}

// Flat stores (one list: a flat index, the centroid list of the coarse scan, the k-means assign): every query
// samples the SAME rows -- the first `sample` rows of the list -- so the seed scores are one small dense
// [Q x sample x d] contraction: a register-tiled FP32 kernel (64 queries x 64 rows per CTA, 4 x 4 per thread)
// writes the keys, a second kernel selects the kc-th smallest per query.
template <bool kIP>
__global__ void __launch_bounds__(256) seed_scores_flat_kernel(const float* __restrict__ vecs, int64_t pitch,
                                                               const float* __restrict__ norms, int d,
                                                               const float* __restrict__ queries, int64_t q_pitch, int64_t Q,
                                                               int64_t row0, int nrows, int sample,
                                                               uint32_t* __restrict__ skeys) {
    __shared__ float qs[32][65];  // [k][query]
    __shared__ float rs[32][65];  // [k][row]
    const int tid = threadIdx.x;
    const int64_t qb = (int64_t)blockIdx.x * 64;
    const int rb = blockIdx.y * 64;
    const int tq = tid & 15, tr = tid >> 4;  // thread tile: queries tq*4.., rows tr*4..
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < d; k0 += 32) {
        // 64 x 32 floats each: thread loads 8 + 8 scalars (row-major sources, coalesced over k)
        for (int e = tid; e < 64 * 32; e += 256) {
            const int r = e >> 5, k = e & 31;
            const int64_t q = qb + r;
            qs[k][r] = (q < Q && k0 + k < d) ? queries[q * q_pitch + k0 + k] : 0.f;
            const int row = rb + r;
            rs[k][r] = (row < nrows && row < sample && k0 + k < d) ? vecs[(row0 + row) * pitch + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = qs[k][tq * 4 + i]; b[i] = rs[k][tr * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t q = qb + tq * 4 + i;
        if (q >= Q) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int row = rb + tr * 4 + j;
            if (row >= sample) continue;
            uint32_t key = KEY_MAX;
            if (row < nrows) key = f2key(kIP ? -acc[i][j] : fmaf(-2.f, acc[i][j], norms[row0 + row]));
            skeys[q * sample + row] = key;
        }
    }
}

__global__ void __launch_bounds__(256) seed_select_flat_kernel(const uint32_t* __restrict__ skeys, int sample, int have,
                                                               const float* __restrict__ queries, int64_t q_pitch, int d,
                                                               int64_t Q, int kc, float max_row_norm,
                                                               const float* __restrict__ max_row_norm_dev, float rel_margin,
                                                               uint32_t* __restrict__ gthr) {
    const int lane = threadIdx.x & 31;
    if (max_row_norm_dev) max_row_norm = *max_row_norm_dev;
    const int64_t q = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= Q || have < kc) return;
    uint32_t kreg[32];
#pragma unroll
    for (int s2 = 0; s2 < 32; ++s2) {
        const int i = s2 * 32 + lane;
        kreg[s2] = i < have ? skeys[q * sample + i] : KEY_MAX;
    }
    const int nslot = (have + 31) >> 5;
    float qq = 0.f;
    for (int i = lane; i < d; i += 32) {
        const float x = queries[q * q_pitch + i];
        qq = fmaf(x, x, qq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
    uint32_t lo = 0;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = lo | (1u << bit);
        int c = 0;
#pragma unroll
        for (int s2 = 0; s2 < 32; ++s2)
            if (s2 < nslot) c += __popc(__ballot_sync(0xffffffffu, kreg[s2] < cand));
        if (c < kc) lo = cand;
    }
    if (lane == 0 && lo < KEY_MAX) {
        const float t = key2f(lo);
        const float margin = __fadd_ru(__fmul_ru(rel_margin, __fmul_ru(sqrtf(qq), max_row_norm)), fabsf(t) * 9.5367431640625e-07f);
        const float tm = __fadd_ru(t, margin);
        if (tm == tm) atomicMin(gthr + q, f2key(tm));
    }
}

// Top-1 mode margins: delta[q] = 3 x the filter/refine error bound of query q (the proof of refine.cuh needs the
// rejected rows' scores to exceed the best one by more than 2 x that bound), rounded up.
__global__ void top1_delta_kernel(const float* __restrict__ queries, int64_t q_pitch, int d, int64_t Q, float max_row_norm,
                                  const float* __restrict__ max_row_norm_dev, double filter_gam, int ip,
                                  float* __restrict__ qdelta) {
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= Q) return;
    double qn = 0.0;
    for (int i = lane; i < d; i += 32) {
        const double x = queries[q * q_pitch + i];
        qn += x * x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qn += __shfl_xor_sync(0xffffffffu, qn, o);
    if (lane) return;
    const double U = (double)(max_row_norm_dev ? *max_row_norm_dev : max_row_norm), qnorm = sqrt(qn);
    const double eps = 5.960464477539063e-08, gam = (d + 8) * eps, e2 = (d / 8 + 12) * eps;
    double e;
    if (!ip) e = gam * U * U + 2.0 * (gam + filter_gam) * qnorm * U + (8.0 * eps + e2) * (qnorm + U) * (qnorm + U);
    else e = (gam + filter_gam + e2) * qnorm * U + 8.0 * eps * qnorm * U;
    qdelta[q] = __double2float_ru(3.0 * e) * 1.0001f + 1e-30f;
}

// ------------------------------------------------------------------------------------------------
// 4. the scan (filter) kernel
// ------------------------------------------------------------------------------------------------
struct ScanArgs {
    const float* norms;  // squared row norms (l2 only)
    int dp;
    const float* queries;
    int64_t q_pitch;
    const int32_t* seg_pairs;
    const WorkItem* items;
    int32_t* ctrl;
    uint32_t* gthr;    // [Q] running filter-key threshold of every query (upper bound on its kc-th best key)
    int32_t* qcount;   // [Q] entries appended to the query's candidate buffer (may exceed qcap: overflow)
    uint64_t* qbuf;    // [Q][qcap] candidates: key << 32 | arena row; unwritten slots are 0xff..ff
    int P, kc, gq, nq;
    int qcap;
    int qstride;       // ints between consecutive queries' fill counters
    int refresh_first;                 // an extra, early refresh request when a query's fill reaches this (0: none)
    int refresh_boxes, refresh_fresh;  // mailboxes per epilogue warp - 1 (0, 1 or 3); re-read the fill when serving
    int refresh_step, refresh_window;  // tensor-core path: threshold refresh cadence (appends, power of two) / entries looked at
    uint32_t* dense;      // dense mode: [Q x dense_rows] filter keys of every (query, row); null otherwise
    long long dense_row0;
    int dense_rows;
    // flat mode (single-list store scanned without a probe table): the work items are the grid (segment x chunk of gq
    // consecutive queries), item it = seg * flat_nchunks + chunk, computed in the kernel -- no pair tables, no grouping
    int fixed_thr;  // collect mode: thresholds are given and never refreshed
    int top1;       // k = 1: thresholds follow the running minimum + qdelta[query] (no seeds, no kc-th refresh needed)
    const float* qdelta;  // [Q] top-1 margins
    int terms;  // tensor-core filter: 3 = 3xTF32 (a_hi b_hi + a_lo b_hi + a_hi b_lo), 2 = 2xTF32 (no a_lo term)
    int flat;
    int flat_nchunks;
    int flat_items;
    int64_t Q;
    const int64_t* seg_row0;
    const int32_t* seg_rows;
    int dbg;  // profiling aid (-DQK_STAGE_DEBUG builds only): skip pipeline stages to find the floor of the others
};
#ifdef QK_STAGE_DEBUG
#define QK_DBG(a, bit) ((a).dbg & (bit))
#else
#define QK_DBG(a, bit) 0
#endif

// Work item `it` of a flat-mode scan (see ScanArgs::flat)
__device__ __forceinline__ void flat_item(const ScanArgs& a, int it, int& seg, int& g_begin, int& g_cnt) {
    seg = it / a.flat_nchunks;
    const int chunk = it - seg * a.flat_nchunks;
    g_begin = chunk * a.gq;
    const int64_t left = a.Q - g_begin;
    g_cnt = (int)(left < a.gq ? left : a.gq);
}

// One d-chunk (<= 128 floats) of one 64-row tile for NT query slots per lane: 4 rows x NT queries x float4
// per step. The rows sit in shared memory as the TMA engine wrote them with the 128-byte swizzle: sub-tile b
// (32 floats = 128 B of every row) at b * 8 KB, row r at r * 128 B inside it, 16-byte chunk cc stored at
// position cc ^ (r & 7). This lane's rows all have r & 7 == lr, so `lrx` = lr << 4 un-swizzles them.
template <int NT>
__device__ __forceinline__ void fma_chunk(float2 (&acc)[4][4], const unsigned char* __restrict__ vbase, uint32_t lrx,
                                          const float4* __restrict__ qrow, int qstep, int dcur4) {
    auto step = [&](int c, uint32_t off) {
        float4 v[4], q[NT];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(vbase + i * 1024 + off);
#pragma unroll
        for (int t = 0; t < NT; ++t) q[t] = qrow[t * qstep + c];
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][t] = ffma2(make_float2(v[i].x, v[i].y), make_float2(q[t].x, q[t].y), acc[i][t]);
                acc[i][t] = ffma2(make_float2(v[i].z, v[i].w), make_float2(q[t].z, q[t].w), acc[i][t]);
            }
    };
    int c = 0;
    for (; c + 8 <= dcur4; c += 8) {
        const uint32_t sub = (uint32_t)(c >> 3) * 8192u;
#pragma unroll
        for (int u = 0; u < 8; ++u) step(c + u, sub + (((uint32_t)u << 4) ^ lrx));
    }
    for (; c < dcur4; ++c) step(c, (uint32_t)(c >> 3) * 8192u + ((((uint32_t)c & 7u) << 4) ^ lrx));
}

// Shared-memory map (dynamic, 1 KB aligned):
//   [0, 2048)       mbarriers: full[4] empty[4] qfull[4] qempty[4] kfull[2][2] kempty[2][2]; ItemDesc[4] at +256
//   Vs   [STAGES][4][TV][32] f32   row ring, one stage = one 64-row tile x one d-chunk of <= 128 floats, written
//                                  by TMA tensor copies (box 64 rows x 128 B, 128-byte swizzle)
//   Qs   [nq][gq][dp+4]   f32      query-chunk ring, one slot per work item in flight (cp.async gathers)
//   Ks   [2 groups][2][gq][KP] u32 score keys
//   hist [select warps][256] u32   radix-select histograms
//
// Roles: warps 0-3 compute group 0 (even tiles), warps 4-7 compute group 1 (odd tiles), warps 8-14 select
// (warp sw owns queries sw, sw+7, ...), warp 15 producer. Inside a compute group, warp gw covers rows rb*32..
// (rb = gw & 1) x query half qh = gw >> 1; lane (lr = lane & 7, lq = lane >> 3) owns rows rb*32 + lr + 8i (i < 4)
// and queries qh + 2*(lq + 4t) (t < 4), so every shared-memory load is one wavefront (8 distinct rows or 4
// distinct queries, the rest broadcast).
template <bool kIP>
__global__ void __launch_bounds__(SCAN_THREADS, 1) scan_kernel(const ScanArgs a, const __grid_constant__ CUtensorMap vmap) {
    constexpr int TV = SCAN_TV, DC = SCAN_DC, KP = SCAN_KP, NS = SCAN_STAGES;
    constexpr int STAGE_BYTES = TV * DC * 4;  // 32 KB
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    // the 128-byte TMA swizzle is a function of the shared-memory address: align the map to 1 KB
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* full = bars;             // [NS]   producer -> compute (tx)
    uint64_t* empty = bars + 4;        // [NS]   compute  -> producer (4 warps of the owning group)
    uint64_t* qfull = bars + 8;        // [nq]   producer (32 lanes, after their cp.async landed) -> consumers
    uint64_t* qempty = bars + 12;      // [nq]   15 consumer warps -> producer
    uint64_t* kfull = bars + 16;       // [2][2] compute group (4 warps) -> select
    uint64_t* kempty = bars + 20;      // [2][2] select (7 warps) -> compute group
    ItemDesc* descs = reinterpret_cast<ItemDesc*>(smem_raw + 256);  // [SCAN_MAX_NQ]
    unsigned char* Vs = smem_raw + SCAN_SMEM_HEADER;
    const int dp = a.dp, kc = a.kc, gq = a.gq, nq = a.nq;
    const int QP = dp + 4;
    float* Qs = reinterpret_cast<float*>(Vs + (size_t)NS * STAGE_BYTES);
    uint32_t* Ks = reinterpret_cast<uint32_t*>(Qs + (size_t)nq * gq * QP);
    uint32_t* hists = Ks + (size_t)4 * gq * KP;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 4); }
        for (int s = 0; s < SCAN_MAX_NQ; ++s) {
            mbar_init(qfull + s, 32);
            mbar_init(qempty + s, SCAN_COMPUTE_WARPS + SCAN_SELECT_WARPS);
        }
        for (int s = 0; s < 4; ++s) { mbar_init(kfull + s, 4); mbar_init(kempty + s, SCAN_SELECT_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    const int ndc = (dp + DC - 1) / DC;

    if (warp == SCAN_COMPUTE_WARPS + SCAN_SELECT_WARPS) {
        // ===================================================================== producer warp
        // Software-pipelined over work items so that the dependent metadata loads (work counter -> item ->
        // pair ids -> thresholds) of the next items are in flight while the current item is issued:
        //   it3: index reserved for item n+3      m2: WorkItem of item n+2
        //   m1 + pair1: item n+1 and its pair ids  m0 + pair0 + gthr0: item n, complete
        const int n_items = a.flat ? a.flat_items : a.ctrl[1];
        auto fetch_index = [&]() {
            int it = 0;
            if (lane == 0) it = atomicAdd(&a.ctrl[0], 1);
            return __shfl_sync(0xffffffffu, it, 0);
        };
        auto fetch_item = [&](int it) {
            WorkItem w;
            w.seg = -1; w.g_begin = 0; w.g_cnt = 0; w.nrows = 0; w.row0 = 0; w.pad_ = 0;
            if (it < n_items) {
                if (a.flat) {
                    flat_item(a, it, w.seg, w.g_begin, w.g_cnt);
                    w.nrows = a.seg_rows[w.seg];
                    w.row0 = a.seg_row0[w.seg];
                } else {
                    w = a.items[it];
                }
            }
            return w;
        };
        // pair index = query * P + slot; only the query (pair / P) is ever used downstream
        auto fetch_pair = [&](const WorkItem& w) {
            if (w.seg < 0 || lane >= w.g_cnt) return -1;
            return a.flat ? (w.g_begin + lane) * a.P : a.seg_pairs[w.g_begin + lane];
        };
        // thresholds move during the kernel (atomicMin from every SM): read them past the non-coherent L1
        auto fetch_gthr = [&](int pair) { return pair >= 0 ? __ldcg(a.gthr + pair / a.P) : KEY_MAX; };
        WorkItem m0 = fetch_item(fetch_index());
        WorkItem m1 = fetch_item(fetch_index());
        WorkItem m2 = fetch_item(fetch_index());
        int pair0 = fetch_pair(m0);
        int pair1 = fetch_pair(m1);
        uint32_t gthr0 = fetch_gthr(pair0);
        uint32_t U = 0;
        const int dp4 = dp >> 2;
        for (uint32_t n = 0;; ++n) {
            const int ib = n % nq;
            const int it3 = fetch_index();
            mbar_wait(qempty + ib, ((n / nq) & 1u) ^ 1u);
            if (m0.seg < 0) {
                if (lane == 0) descs[ib].w.seg = -1;
                __syncwarp();
                mbar_arrive(qfull + ib);  // all 32 lanes
                break;
            }
            const int g_cnt = m0.g_cnt, nrows = m0.nrows;
            const int64_t row0 = m0.row0;
            descs[ib].pair[lane] = pair0;
            descs[ib].gthr[lane] = gthr0;
            if (lane == 0) descs[ib].w = m0;
            __threadfence_block();  // the descriptor is published by the (asynchronous) arrivals below
            // gather the item's query rows (16-byte cp.async, L2 only); the slot's barrier counts one arrival
            // per lane, delivered when that lane's copies have landed
            {
                float* qdst = Qs + (size_t)ib * gq * QP;
                for (int g = 0; g < g_cnt; ++g) {
                    const int64_t q = __shfl_sync(0xffffffffu, pair0, g) / a.P;
                    const float* src = a.queries + q * a.q_pitch;
                    for (int c = lane; c < dp4; c += 32) cp_async16(qdst + (size_t)g * QP + 4 * c, src + 4 * c);
                }
                cp_async_mbar_arrive_noinc(qfull + ib);
            }
            const int ntiles = (nrows + TV - 1) / TV;
            for (int tile = 0; tile < ntiles; ++tile) {
                for (int dc = 0; dc < ndc; ++dc, ++U) {
                    const int st = U % NS;
                    mbar_wait(empty + st, ((U / NS) & 1u) ^ 1u);
                    if (lane == 0) {
                        const int dcur = min(DC, dp - dc * DC);
                        const int nbox = (dcur + 31) >> 5;
                        mbar_expect_tx(full + st, (uint32_t)(nbox * TV * 128));
                        for (int b = 0; b < nbox; ++b)
                            tma_load_2d(Vs + (size_t)st * STAGE_BYTES + (size_t)b * (TV * 128), &vmap, dc * DC + b * 32,
                                        (int)(row0 + (int64_t)tile * TV), full + st);
                    }
                    __syncwarp();
                }
            }
            // rotate the pipeline; the thresholds of the next item are read as late as possible
            const int pair2 = fetch_pair(m2);
            m0 = m1; pair0 = pair1;
            m1 = m2; pair1 = pair2;
            m2 = fetch_item(it3);
            gthr0 = fetch_gthr(pair0);
        }
    } else if (warp >= SCAN_COMPUTE_WARPS) {
        // ===================================================================== select warps
        // Every key at or below its query's running threshold is appended to the query's candidate buffer in
        // global memory (one atomicAdd per (query, tile), all of the warp's queries in one instruction). Each
        // time a query's fill passes a multiple of `step`, the warp that crossed it re-derives the threshold
        // (kc-th smallest key appended so far) and publishes it with atomicMin.
        constexpr int NSEL = SCAN_SELECT_WARPS;
        const int sw = warp - SCAN_COMPUTE_WARPS;
        uint32_t* hist = hists + sw * 256;
        const int qcap = a.qcap;
        const int step = kc > 64 ? kc : 64;
        const unsigned below = (1u << lane) - 1u;
        uint32_t T = 0;
        for (uint32_t n = 0;; ++n) {
            const int ib = n % nq;
            mbar_wait(qfull + ib, (n / nq) & 1u);
            const WorkItem d = descs[ib].w;
            if (d.seg < 0) break;
            // lane j keeps the state of query slot g = sw + NSEL*j: query index and threshold
            int my_q = -1;
            uint32_t my_lim = 0;
            {
                const int g = sw + NSEL * lane;
                if (g < d.g_cnt) {
                    my_q = descs[ib].pair[g] / a.P;
                    const uint32_t t = descs[ib].gthr[g];
                    my_lim = t < KEY_MAX ? t : KEY_MAX - 1;  // KEY_MAX marks an invalid row
                }
            }
            // Per tile: (1) pass masks + the passing keys into registers, key buffer released at once;
            // (2) one atomicAdd per query reserves buffer slots; (3) the entries of the PREVIOUS tile are
            // written -- its atomics were issued one tile ago, so their latency is off the critical path.
            constexpr int JMAX = (SCAN_GQ + NSEL - 1) / NSEL;
            struct Pending {
                uint32_t k0[JMAX], k1[JMAX];  // this lane's two keys of every query slot
                unsigned m0, m1;              // lane j: pass masks of query slot j
                int base, nn;                 // lane j: reserved slot range
                uint32_t arow0;
                bool live;
            };
            Pending pend;
            pend.live = false;
            pend.m0 = pend.m1 = 0; pend.base = pend.nn = 0; pend.arow0 = 0;
            auto flush = [&](const Pending& pd) {
                unsigned todo = __ballot_sync(0xffffffffu, pd.nn > 0);
                unsigned cross = 0;  // queries whose fill passed a multiple of `step`
#pragma unroll
                for (int j = 0; j < JMAX; ++j) {
                    if (!((todo >> j) & 1u)) continue;
                    const unsigned m0 = __shfl_sync(0xffffffffu, pd.m0, j), m1 = __shfl_sync(0xffffffffu, pd.m1, j);
                    const int b = __shfl_sync(0xffffffffu, pd.base, j);
                    const int q = __shfl_sync(0xffffffffu, my_q, j);
                    uint64_t* qb = a.qbuf + (size_t)q * qcap;
                    const int c0 = __popc(m0), c1 = __popc(m1);
                    if ((m0 >> lane) & 1u) {
                        const int slot = b + __popc(m0 & below);
                        if (slot < qcap) qb[slot] = ((uint64_t)pd.k0[j] << 32) | (pd.arow0 + lane);
                    }
                    if ((m1 >> lane) & 1u) {
                        const int slot = b + c0 + __popc(m1 & below);
                        if (slot < qcap) qb[slot] = ((uint64_t)pd.k1[j] << 32) | (pd.arow0 + 32 + lane);
                    }
                    const int e = b + c0 + c1;
                    if (b / step != e / step && e >= kc && !a.fixed_thr) cross |= 1u << j;
                }
                // refresh the thresholds of the queries that crossed. The entries this warp just stored are
                // read back by other lanes: order them first.
                if (cross) __threadfence();
                while (cross) {
                    const int j = __ffs(cross) - 1;
                    cross &= cross - 1;
                    const int q = __shfl_sync(0xffffffffu, my_q, j);
                    const int fill = __shfl_sync(0xffffffffu, pd.base + pd.nn, j);
                    const unsigned long long* qb = reinterpret_cast<const unsigned long long*>(a.qbuf) + (size_t)q * qcap;
                    const uint32_t t = radix_select([qb](int i) { return (uint32_t)(__ldcg(qb + i) >> 32); },
                                                    fill < qcap ? fill : qcap, kc, hist, lane);
                    if (t < KEY_MAX) {
                        if (lane == 0) atomicMin(a.gthr + q, t);
                        if (lane == j && t < my_lim) my_lim = t;
                    }
                }
            };
            const int ntiles = (d.nrows + TV - 1) / TV;
            for (int tile = 0; tile < ntiles; ++tile, ++T) {
                const int grp = T & 1u, kb = (T >> 1) & 1u;
                // what every other SM has learnt about my queries meanwhile (used from the next tile on)
                const uint32_t g_now = my_q >= 0 ? __ldcg(a.gthr + my_q) : 0u;
                mbar_wait(kfull + grp * 2 + kb, (T >> 2) & 1u);
                const uint32_t* kbase = Ks + (size_t)(grp * 2 + kb) * gq * KP;
                Pending cur;
                cur.live = true;
                cur.m0 = cur.m1 = 0;
                cur.arow0 = (uint32_t)(d.row0 + (int64_t)tile * TV);
#pragma unroll
                for (int j = 0; j < JMAX; ++j) {
                    const int g = sw + NSEL * j;
                    cur.k0[j] = cur.k1[j] = KEY_MAX;
                    if (g >= d.g_cnt) continue;  // warp-uniform
                    uint32_t lim = __shfl_sync(0xffffffffu, my_lim, j);
                    const uint32_t k0 = kbase[g * KP + lane], k1 = kbase[g * KP + 32 + lane];
                    unsigned m0 = __ballot_sync(0xffffffffu, k0 <= lim);
                    unsigned m1 = __ballot_sync(0xffffffffu, k1 <= lim);
                    if (kc <= 48 && !a.fixed_thr && __popc(m0) + __popc(m1) > kc + 8) {
                        // a loose (stale or missing) threshold: this tile alone bounds the kc-th best key
                        const uint32_t* kq = kbase + g * KP;
                        const uint32_t t = radix_select([kq](int i) { return kq[i]; }, TV, kc, hist, lane);
                        if (t < lim) {
                            lim = t;
                            m0 = __ballot_sync(0xffffffffu, k0 <= lim);
                            m1 = __ballot_sync(0xffffffffu, k1 <= lim);
                            if (lane == 0) atomicMin(a.gthr + descs[ib].pair[g] / a.P, lim);
                        }
                    }
                    cur.k0[j] = k0; cur.k1[j] = k1;
                    if (lane == j) { cur.m0 = m0; cur.m1 = m1; my_lim = lim; }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(kempty + grp * 2 + kb);
                cur.nn = __popc(cur.m0) + __popc(cur.m1);
                cur.base = 0;
                if (cur.nn > 0) cur.base = atomicAdd(&a.qcount[(size_t)my_q * a.qstride], cur.nn);
                if (pend.live) flush(pend);
                pend = cur;
                if (g_now < my_lim) my_lim = g_now;
            }
            if (pend.live) flush(pend);
            __syncwarp();
            if (lane == 0) mbar_arrive(qempty + ib);
        }
    } else {
        // ===================================================================== compute warps
        const int grp = warp >> 2, gw = warp & 3;
        const int rb = gw & 1, qh = gw >> 1;
        const int lr = lane & 7, lq = lane >> 3;
        const uint32_t lrx = (uint32_t)lr << 4;
        uint32_t T = 0, U = 0;
        for (uint32_t n = 0;; ++n) {
            const int ib = n % nq;
            mbar_wait(qfull + ib, (n / nq) & 1u);
            const WorkItem d = descs[ib].w;
            if (d.seg < 0) break;
            const int g_cnt = d.g_cnt;
            // queries of this lane: g = qh + 2*(lq + 4t); t < nt is warp-uniform
            const int jmax = (g_cnt - qh + 1) >> 1;            // number of j = lq + 4t with g < g_cnt
            const int nt = (jmax + 3) >> 2;                    // 0..4
            const float* qbase = Qs + ((size_t)ib * gq + qh + 2 * lq) * QP;
            const int qstep = 8 * QP / 4;                      // float4 stride between this lane's queries
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
                    const unsigned char* vbase = Vs + (size_t)st * STAGE_BYTES + (size_t)(rb * 32 + lr) * 128;
                    const float4* qrow = reinterpret_cast<const float4*>(qbase + dc * DC);
                    switch (nt) {
                        case 1: fma_chunk<1>(acc, vbase, lrx, qrow, qstep, dcur4); break;
                        case 2: fma_chunk<2>(acc, vbase, lrx, qrow, qstep, dcur4); break;
                        case 3: fma_chunk<3>(acc, vbase, lrx, qrow, qstep, dcur4); break;
                        case 4: fma_chunk<4>(acc, vbase, lrx, qrow, qstep, dcur4); break;
                        default: break;
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

}  // namespace qk
#include "scan_mma.cuh"
#include "refine.cuh"
namespace qk {

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// optional per-launch timing of the filter kernel (qk_profile_*): CUDA events recorded on the launch
// stream right around scan_kernel, read back by the caller after it has synchronised.
struct ProfileRecord {
    cudaEvent_t start, stop;
    int64_t queries;
    int nprobe, k, used;
};
static ProfileRecord* g_prof = nullptr;
static int g_prof_cap = 0, g_prof_n = 0;

// Fork/join partner of the caller's stream (threshold seeds run beside the grouping kernels): one per device, created
// on first use under a lock -- an index on a second GPU of the same process gets its own.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaEvent_t fork2 = nullptr, join2 = nullptr;  // qk_search_ivf: workspace clears beside the coarse scan
};
static constexpr int QK_MAX_DEVICES = 64;
static SideStream g_side[QK_MAX_DEVICES];
static std::mutex g_side_mutex;
static int side_stream_for_current_device(SideStream** out) {
    int dev = 0;
    QK_CUDA(cudaGetDevice(&dev));
    QK_REQUIRE(dev >= 0 && dev < QK_MAX_DEVICES, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_side_mutex);
    SideStream& ss = g_side[dev];
    if (!ss.stream) {
        QK_CUDA(cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking));
        QK_CUDA(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
        QK_CUDA(cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming));
        QK_CUDA(cudaEventCreateWithFlags(&ss.fork2, cudaEventDisableTiming));
        QK_CUDA(cudaEventCreateWithFlags(&ss.join2, cudaEventDisableTiming));
    }
    *out = &ss;
    return QK_OK;
}

// Tensor map of the row arena for the scan kernel's TMA loads: [num_rows x pitch] f32, box = 64 rows x 32
// floats (128 B), 128-byte swizzle, out-of-bounds elements read as zero.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_row_tensor_map(const qk_store_t* st, int box_rows, CUtensorMap* out) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        QK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
            return QK_ERR_CUDA;
        }
        encode = (EncodeTiledFn)fn;
    }
    QK_REQUIRE(st->num_rows > 0, "store.num_rows must be set");
    cuuint64_t dims[2] = {(cuuint64_t)st->pitch, (cuuint64_t)st->num_rows};
    cuuint64_t strides[1] = {(cuuint64_t)st->pitch * sizeof(float)};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)st->vectors, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld pitch=%lld)", (int)r,
                  (long long)st->num_rows, (long long)st->pitch);
        return QK_ERR_CUDA;
    }
    return QK_OK;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel and size instead of before every launch
static int ensure_smem_impl(const void* kern, size_t bytes) {
    static std::mutex m;
    static const void* fn[256];
    static int fdev[256];
    static size_t granted[256];
    static int n = 0;
    int dev = 0;
    QK_CUDA(cudaGetDevice(&dev));  // the attribute is per function AND per device
    std::lock_guard<std::mutex> lock(m);
    int i = 0;
    while (i < n && (fn[i] != kern || fdev[i] != dev)) ++i;
    if (i == n) {
        if (n == 256) {
            QK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            return QK_OK;
        }
        fn[n] = kern;
        fdev[n] = dev;
        granted[n] = 0;  // static + dynamic shared memory together may pass 48 KB: always opt in
        ++n;
        // these kernels live on shared memory, not L1: ask for the largest carve-out so that occupancy is not capped
        // by the default split
        QK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    }
    if (bytes > granted[i]) {
        QK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        granted[i] = bytes;
    }
    return QK_OK;
}
template <typename K>
static int ensure_smem(K kern, size_t bytes) { return ensure_smem_impl((const void*)kern, bytes); }

static int launch_scan(const ScanArgs& sa, const CUtensorMap& vmap, const CUtensorMap& vmap_sub, int metric, size_t smem, bool mma,
                       cudaStream_t stream) {
    const int grid = sm_count();
    int rc;
    if (mma) {
        const size_t msmem = scan_mma_smem_bytes();
        if (metric == QK_METRIC_INNER_PRODUCT) {
            if ((rc = ensure_smem(scan_mma_kernel<true>, msmem))) return rc;
            scan_mma_kernel<true><<<grid, MMA_THREADS, msmem, stream>>>(sa, vmap, vmap_sub);
        } else {
            if ((rc = ensure_smem(scan_mma_kernel<false>, msmem))) return rc;
            scan_mma_kernel<false><<<grid, MMA_THREADS, msmem, stream>>>(sa, vmap, vmap_sub);
        }
    } else if (metric == QK_METRIC_INNER_PRODUCT) {
        if ((rc = ensure_smem(scan_kernel<true>, smem))) return rc;
        scan_kernel<true><<<grid, SCAN_THREADS, smem, stream>>>(sa, vmap);
    } else {
        if ((rc = ensure_smem(scan_kernel<false>, smem))) return rc;
        scan_kernel<false><<<grid, SCAN_THREADS, smem, stream>>>(sa, vmap);
    }
    QK_LAUNCHED();
    return QK_OK;
}

}  // namespace qk

using namespace qk;

// The zero / +inf initialisation a scan needs of its workspace: seg_count, seg_fill, flags, qcount, ctrl are contiguous
// (one memset); candidate slots start as +inf (a refresh may read a slot that was reserved but not written yet).
static int clear_scan_workspace(const ScanPlan& p, char* ws, int64_t Q, cudaStream_t stream) {
    QK_CUDA(cudaMemsetAsync(ws + p.off_seg_count, 0, p.off_seg_start - p.off_seg_count, stream));
    if (!p.dense) QK_CUDA(cudaMemsetAsync(ws + p.off_qbuf, 0xff, (size_t)Q * p.qcap * 8, stream));
    return QK_OK;
}

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
    ScanExtras ex;
    return qk::scan_partitions_impl(st, queries, Q, q_pitch, probe_lists, nprobe, metric, k, out_ids, out_dist, out_rows,
                                    workspace, workspace_bytes, stats, stream_v, ex);
}

int qk::scan_partitions_impl(const qk_store_t* st, const float* queries, int64_t Q, int64_t q_pitch,
                             const int32_t* probe_lists, int nprobe, int metric, int k, int64_t* out_ids,
                             float* out_dist, int64_t* out_rows, void* workspace, size_t workspace_bytes,
                             int32_t* stats, void* stream_v, const ScanExtras& ex) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    const bool collect = ex.preset_thresholds != nullptr;
    QK_REQUIRE(st && queries && (collect || (out_ids && out_dist)), "null argument");
    QK_REQUIRE(!collect || (ex.collect_ids && ex.collect_dist && ex.collect_cnt && ex.collect_overflow), "collect outputs");
    QK_REQUIRE(metric == QK_METRIC_L2 || metric == QK_METRIC_INNER_PRODUCT, "metric %d not supported", metric);
    QK_REQUIRE(Q > 0 && nprobe > 0, "empty query batch");
    QK_REQUIRE(st->num_segments > 0 && st->num_lists > 0, "store has no segments");
    // flat mode: a single-list store scanned without a probe table -- every query scans the whole list
    const bool flat = (probe_lists == nullptr && ex.probe_ids == nullptr);
    QK_REQUIRE(!flat || (st->num_lists == 1 && nprobe == 1), "a probe table is required unless the store has one list");
    QK_REQUIRE(probe_lists || flat || (ex.id_to_slot && ex.table_size > 0), "probe ids need an id -> slot table");
    ScanPlan p;
    int rc = make_plan(st, Q, nprobe, k, &p, /*allow_dense=*/flat);  // the dense key matrix is indexed by (query, row of THE list)
    if (rc) return rc;
    // top-1 mode: k = 1 on the tensor-core path -- thresholds follow the running minimum, no seeds
    const bool top1 = (k == 1) && (g_scan_variant == 0) && p.dp <= 128 && !collect && !p.dense && getenv("QK_NO_TOP1") == nullptr;
    if (top1) p.sample = 0;
    if (collect) {
        QK_REQUIRE(!flat && p.P == nprobe, "collect mode needs single-segment lists and a probe table");
        if (p.P != nprobe) return QK_ERR_UNSUPPORTED;
        p.sample = 0;  // thresholds are given
    }
    QK_REQUIRE(q_pitch >= p.dp && q_pitch % 4 == 0, "query pitch %lld must be a multiple of 4 and >= %d",
               (long long)q_pitch, p.dp);
    QK_REQUIRE(((uintptr_t)queries % 16 == 0) && ((uintptr_t)st->vectors % 16 == 0), "vectors must be 16-byte aligned");
    if (workspace_bytes < p.total || !workspace) {
        set_error("workspace too small: need %zu bytes, have %zu", p.total, workspace_bytes);
        return QK_ERR_WORKSPACE;
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
    WorkItem* items = (WorkItem*)(ws + p.off_items);
    uint32_t* gthr = (uint32_t*)(ws + p.off_gthr);
    int32_t* qcount = (int32_t*)(ws + p.off_qcount);
    uint64_t* qbuf = (uint64_t*)(ws + p.off_qbuf);
    const int S = st->num_segments;
    const int64_t QP = Q * p.P;
    const bool ip = metric == QK_METRIC_INNER_PRODUCT;
    const bool use_mma = (g_scan_variant == 0) && p.dp <= 128;
    int terms = g_filter_terms_env ? g_filter_terms_env : st->filter_terms;
    if (terms != 2) terms = 3;
    const double fgam = use_mma ? filter_gamma(terms) : 0.0;

    if (!ex.ws_precleared && (rc = clear_scan_workspace(p, ws, Q, stream))) return rc;
    if (flat) {
        if (!p.dense) QK_CUDA(cudaMemsetAsync(gthr, 0xff, (size_t)Q * 4, stream));  // KEY_MAX: no threshold yet
    } else {
        ProbeSource ps;
        ps.slots = probe_lists; ps.ids = ex.probe_ids; ps.id_to_slot = ex.id_to_slot; ps.table_size = ex.table_size;
        ps.shard_rank = ex.shard_rank; ps.shard_world = ex.shard_world;
        const bool single = (p.P == nprobe);
        if (ex.pairs_preexpanded) {
            QK_REQUIRE(single && !collect, "pre-expanded pairs need single-segment lists");
        } else if (single) {
            int64_t n = Q * nprobe;
            expand_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ps, Q, nprobe, p.P, st->list_seg0, st->list_nseg,
                                                                                  st->num_lists, pair_seg, seg_count, gthr, true, !collect);
        } else if (nprobe == 1) {
            int64_t n = Q * p.P;
            expand_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ps, Q, nprobe, p.P, st->list_seg0, st->list_nseg,
                                                                                  st->num_lists, pair_seg, seg_count, gthr, false, !collect);
        } else {
            expand_pairs_kernel<<<(unsigned)((Q + 127) / 128), 128, 0, stream>>>(ps, Q, nprobe, p.P, st->list_seg0, st->list_nseg,
                                                                                  st->num_lists, pair_seg, seg_count, gthr, false, !collect);
        }
        if (!ex.pairs_preexpanded) QK_LAUNCHED();
        if (collect) QK_CUDA(cudaMemcpyAsync(gthr, ex.preset_thresholds, (size_t)Q * 4, cudaMemcpyDeviceToDevice, stream));
    }
    // The seeds only need the pair table; the grouping kernels (prefix, scatter) only the histogram: run them side by
    // side (fork / join through a second stream -- inside a CUDA-graph capture this becomes two parallel branches).
    SideStream* side = nullptr;
    cudaStream_t seed_stream = stream;
    const bool forked = p.sample && !flat;  // flat mode has no grouping kernels to overlap with
    if (forked) {
        if ((rc = side_stream_for_current_device(&side))) return rc;
        QK_CUDA(cudaEventRecord(side->fork, stream));
        QK_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
        seed_stream = side->stream;
    }
    if (p.sample) {
        const int sample = p.sample;
        const float rel_margin = 4.f * (float)(st->d + 8) * 5.9604645e-08f + 4.f * (float)fgam;
        if (p.flat_seed) {
            // the single list's first rows (its segments are consecutive in the arena); host-known geometry
            if (st->flat_rows > 0) {
                uint32_t* skeys = (uint32_t*)(ws + p.off_skeys);
                const int have = (int)(st->flat_rows < sample ? st->flat_rows : sample);
                dim3 grid((unsigned)((Q + 63) / 64), (unsigned)((have + 63) / 64));
                if (ip)
                    seed_scores_flat_kernel<true><<<grid, 256, 0, seed_stream>>>(st->vectors, st->pitch, st->row_norms, st->d, queries,
                                                                            q_pitch, Q, st->flat_row0, (int)st->flat_rows, sample, skeys);
                else
                    seed_scores_flat_kernel<false><<<grid, 256, 0, seed_stream>>>(st->vectors, st->pitch, st->row_norms, st->d, queries,
                                                                             q_pitch, Q, st->flat_row0, (int)st->flat_rows, sample, skeys);
                QK_LAUNCHED();
                seed_select_flat_kernel<<<(unsigned)((Q + 7) / 8), 256, 0, seed_stream>>>(skeys, sample, have, queries, q_pitch, st->d, Q,
                                                                                     p.kc, st->max_row_norm, ex.max_row_norm_dev,
                                                                                     rel_margin, gthr);
                QK_LAUNCHED();
            }
        } else if (!flat) {
            const size_t ssm = seed_smem_bytes(p.dp, sample);
            const unsigned grid = (unsigned)Q;
            if ((rc = ip ? ensure_smem(seed_thresholds_kernel<true>, ssm) : ensure_smem(seed_thresholds_kernel<false>, ssm))) return rc;
            if (ip)
                seed_thresholds_kernel<true><<<grid, SEED_THREADS, ssm, seed_stream>>>(st->vectors, st->pitch, st->row_norms, st->d, p.dp,
                                                                         queries, q_pitch, Q, pair_seg, p.P, st->seg_row0,
                                                                         st->seg_rows, p.kc, sample, seed_chunk_rows(p.dp, sample),
                                                                         st->max_row_norm, rel_margin, gthr);
            else
                seed_thresholds_kernel<false><<<grid, SEED_THREADS, ssm, seed_stream>>>(st->vectors, st->pitch, st->row_norms, st->d, p.dp,
                                                                          queries, q_pitch, Q, pair_seg, p.P, st->seg_row0,
                                                                          st->seg_rows, p.kc, sample, seed_chunk_rows(p.dp, sample),
                                                                          st->max_row_norm, rel_margin, gthr);
            QK_LAUNCHED();
        }
    }
    if (forked) QK_CUDA(cudaEventRecord(side->join, side->stream));
    ScanArgs sa;
    memset(&sa, 0, sizeof(sa));
    if (!flat) {
        const int64_t n = QP > S ? QP : S;
        QK_CUDA(launch_pdl(1, prefix_segments_kernel, dim3(1), dim3(1024), 0, stream, (const int32_t*)seg_count, S, p.gq, seg_start,
                           item_start, ctrl));
        QK_LAUNCHED();
        QK_CUDA(launch_pdl(1, scatter_pairs_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream,
                           (const int32_t*)pair_seg, QP, S, (const int32_t*)seg_start, seg_fill, seg_pairs,
                           (const int32_t*)item_start, st->seg_row0, st->seg_rows, p.gq, items));
        QK_LAUNCHED();
    } else {
        sa.flat = 1;
        sa.flat_nchunks = (int)((Q + p.gq - 1) / p.gq);
        const int64_t n_items = (int64_t)S * sa.flat_nchunks;
        QK_REQUIRE(n_items < ((int64_t)1 << 31), "too many work items; split the query batch");
        sa.flat_items = (int)n_items;
    }
    sa.norms = st->row_norms; sa.dp = p.dp;
    sa.queries = queries; sa.q_pitch = q_pitch;
    sa.seg_pairs = seg_pairs; sa.items = items; sa.ctrl = ctrl;
    sa.gthr = gthr; sa.qcount = qcount; sa.qbuf = qbuf;
    sa.P = p.P; sa.kc = p.kc; sa.gq = p.gq; sa.nq = p.nq; sa.qcap = p.qcap; sa.qstride = p.qstride;
    sa.Q = Q; sa.seg_row0 = st->seg_row0; sa.seg_rows = st->seg_rows;
    sa.terms = terms;
    {
        static const int env_step = getenv("QK_REFRESH_STEP") ? atoi(getenv("QK_REFRESH_STEP")) : 0;      // experiments
        static const int env_win = getenv("QK_REFRESH_WINDOW") ? atoi(getenv("QK_REFRESH_WINDOW")) : 0;
        sa.refresh_step = env_step > 0 ? env_step : 64;
        sa.refresh_window = env_win > 0 ? env_win : 256;
        static const int env_boxes = getenv("QK_REFRESH_BOXES") ? atoi(getenv("QK_REFRESH_BOXES")) : 0;
        static const int env_fresh = getenv("QK_REFRESH_FRESH") ? atoi(getenv("QK_REFRESH_FRESH")) : 0;
        sa.refresh_boxes = env_boxes == 4 ? 3 : (env_boxes == 2 ? 1 : 0);
        sa.refresh_fresh = env_fresh;
        static const int env_first = getenv("QK_REFRESH_FIRST") ? atoi(getenv("QK_REFRESH_FIRST")) : 0;
        sa.refresh_first = env_first;
    }
    sa.fixed_thr = collect ? 1 : 0;
    sa.top1 = top1 ? 1 : 0;
    sa.qdelta = (const float*)(ws + p.off_qdelta);
    if (top1) {
        top1_delta_kernel<<<(unsigned)((Q * 32 + 255) / 256), 256, 0, stream>>>(queries, q_pitch, st->d, Q, st->max_row_norm,
                                                                               ex.max_row_norm_dev, fgam, ip ? 1 : 0,
                                                                               (float*)(ws + p.off_qdelta));
        QK_LAUNCHED();
    }
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("QK_SCAN_DBG"); dbg = e ? atoi(e) : 0; }
        sa.dbg = dbg;
        sa.dense = p.dense ? (uint32_t*)(ws + p.off_dense) : nullptr;
        sa.dense_row0 = st->flat_row0;
        sa.dense_rows = (int)st->flat_rows;
    }
    if (forked) QK_CUDA(cudaStreamWaitEvent(stream, side->join, 0));
    CUtensorMap vmap;
    // d <= 128: tensor-core filter (tcgen05, 3xTF32 split); otherwise the FP32-pipe kernel. QK_SCAN_PATH=ffma
    // forces the latter (tests cross-check the two).
    rc = make_row_tensor_map(st, use_mma ? MMA_TM : SCAN_TV, &vmap);
    if (rc) return rc;
    CUtensorMap vmap_sub = vmap;  // tensor-core path: a half-height box for the short last tile of a list
    if (use_mma && (rc = make_row_tensor_map(st, MMA_SUB_ROWS, &vmap_sub))) return rc;
    ProfileRecord* rec = nullptr;
    if (g_prof && g_prof_n < g_prof_cap) {
        rec = &g_prof[g_prof_n++];
        rec->queries = Q; rec->nprobe = nprobe; rec->k = k; rec->used = 1;
        QK_CUDA(cudaEventRecord(rec->start, stream));
    }
    rc = launch_scan(sa, vmap, vmap_sub, metric, p.smem, use_mma, stream);
    if (rc) return rc;
    if (rec) QK_CUDA(cudaEventRecord(rec->stop, stream));

    if (collect) {
        CollectArgs ca;
        memset(&ca, 0, sizeof(ca));
        ca.vecs = st->vectors; ca.pitch = st->pitch; ca.ids = st->ids; ca.d = st->d;
        ca.seg_row0 = st->seg_row0; ca.seg_rows = st->seg_rows; ca.queries = queries; ca.q_pitch = q_pitch;
        ca.pair_seg = pair_seg; ca.R = nprobe; ca.k = k; ca.qbuf = qbuf; ca.qcount = qcount; ca.qcap = p.qcap; ca.qstride = p.qstride;
        ca.out_ids = ex.collect_ids; ca.out_dist = ex.collect_dist; ca.out_cnt = ex.collect_cnt; ca.overflow = ex.collect_overflow;
        int cap = p.qcap < COLLECT_CAP ? p.qcap : COLLECT_CAP;
        int np = 1;
        while (np < cap) np <<= 1;
        ca.cap = cap; ca.np = np;
        const size_t csm = (size_t)np * 8 + (size_t)cap * 8 + (size_t)((st->d + 3) & ~3) * 4 + (size_t)nprobe * 16 + 64;
        if (ip) {
            if ((rc = ensure_smem(collect_refine_kernel<true>, csm))) return rc;
            collect_refine_kernel<true><<<(unsigned)Q, MERGE_THREADS, csm, stream>>>(ca);
        } else {
            if ((rc = ensure_smem(collect_refine_kernel<false>, csm))) return rc;
            collect_refine_kernel<false><<<(unsigned)Q, MERGE_THREADS, csm, stream>>>(ca);
        }
        QK_LAUNCHED();
        return QK_OK;
    }
    MergeArgs ma;
    memset(&ma, 0, sizeof(ma));
    ma.vecs = st->vectors; ma.pitch = st->pitch; ma.ids = st->ids; ma.d = st->d;
    ma.seg_row0 = st->seg_row0; ma.seg_rows = st->seg_rows; ma.queries = queries; ma.q_pitch = q_pitch;
    ma.pair_seg = pair_seg; ma.flat_nseg = flat ? S : 0;
    ma.gthr = gthr; ma.qbuf = qbuf; ma.qcount = qcount; ma.qcap = p.qcap; ma.qstride = p.qstride;
    ma.flags = flags; ma.ctrl = ctrl; ma.P = p.P; ma.kc = p.kc; ma.k = k;
    ma.max_row_norm = st->max_row_norm;
    ma.max_row_norm_dev = ex.max_row_norm_dev;
    ma.filter_gam = fgam;
    // a 3-term scan also reports how many of its queries the 2-term filter's looser bound would have sent to the
    // exact re-scan (stats[4]): the host's filter precision policy starts safe and relaxes on this evidence
    ma.probe_gam = (use_mma && terms == 3 && !flat) ? filter_gamma(2) : 0.0;
    ma.out_ids = out_ids; ma.out_dist = out_dist; ma.out_rows = out_rows;
    ma.force_rescan = g_force_rescan;
    ma.rank_squared = ex.rank_squared;
    ma.set_mode = (ex.set_mode && p.dense) ? 1 : 0;
    if (ex.fused_expand) {
        const FusedExpand& fx = *ex.fused_expand;
        ma.x_pair_seg = fx.pair_seg; ma.x_seg_count = fx.seg_count; ma.x_gthr = fx.gthr;
        ma.x_id_to_slot = fx.id_to_slot; ma.x_table_size = fx.table_size;
        ma.x_list_seg0 = fx.list_seg0; ma.x_list_nseg = fx.list_nseg;
        ma.x_num_lists = fx.num_lists; ma.x_shard_rank = fx.shard_rank; ma.x_shard_world = fx.shard_world;
    }
    if (ex.pre_refine_event) QK_CUDA(cudaStreamWaitEvent(stream, (cudaEvent_t)ex.pre_refine_event, 0));
    const int kcp = next_pow2(p.kc);
    if (p.dense) {
        ma.dense = sa.dense; ma.dense_rows = (int)st->flat_rows; ma.dense_row0 = st->flat_row0;
        const size_t dsm = (size_t)((st->flat_rows + 1) & ~(int64_t)1) * 4 + refine_tail_bytes(st->d, kcp);
        if (ip) {
            if ((rc = ensure_smem(dense_refine_kernel<true>, dsm))) return rc;
            dense_refine_kernel<true><<<(unsigned)Q, MERGE_THREADS, dsm, stream>>>(ma);
        } else {
            if ((rc = ensure_smem(dense_refine_kernel<false>, dsm))) return rc;
            dense_refine_kernel<false><<<(unsigned)Q, MERGE_THREADS, dsm, stream>>>(ma);
        }
    } else {
        // the gather buffer never needs more than the candidate buffer holds; small buffers let more CTAs share an SM
        int sort_cap = MERGE_SORT_CAP;
        while (sort_cap / 2 >= p.qcap && sort_cap > 256) sort_cap >>= 1;
        if (sort_cap < kcp) sort_cap = kcp;
        ma.sort_cap = sort_cap;
        const size_t msmem = (size_t)sort_cap * 8 + refine_tail_bytes(st->d, kcp);
        if (ip) {
            if ((rc = ensure_smem(merge_refine_kernel<true>, msmem))) return rc;
            QK_CUDA(launch_pdl(2, merge_refine_kernel<true>, dim3((unsigned)Q), dim3(MERGE_THREADS), msmem, stream, ma));
        } else {
            if ((rc = ensure_smem(merge_refine_kernel<false>, msmem))) return rc;
            QK_CUDA(launch_pdl(2, merge_refine_kernel<false>, dim3((unsigned)Q), dim3(MERGE_THREADS), msmem, stream, ma));
        }
    }
    QK_LAUNCHED();
    if (stats) QK_CUDA(cudaMemcpyAsync(stats, ctrl + 2, 8 * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
    return QK_OK;
}

// ---- collect mode: all rows under given per-query thresholds, refined exactly, grouped by probe rank ------------
extern "C" int qk_scan_collect(const qk_store_t* st, const float* queries, int64_t Q, int64_t q_pitch,
                               const int32_t* probe_lists, int nprobe, int metric, int k, const uint32_t* threshold_keys,
                               int64_t* out_ids, float* out_dist, int32_t* out_cnt, int32_t* out_overflow,
                               void* workspace, size_t workspace_bytes, void* stream_v) {
    QK_REQUIRE(probe_lists && threshold_keys && out_ids && out_dist && out_cnt && out_overflow, "null argument");
    ScanExtras ex;
    ex.preset_thresholds = threshold_keys;
    ex.collect_ids = out_ids;
    ex.collect_dist = out_dist;
    ex.collect_cnt = out_cnt;
    ex.collect_overflow = out_overflow;
    return qk::scan_partitions_impl(st, queries, Q, q_pitch, probe_lists, nprobe, metric, k, nullptr, nullptr, nullptr,
                                    workspace, workspace_bytes, nullptr, stream_v, ex);
}

// ---- two-level fixed-nprobe search in one call ------------------------------------------------------
namespace {
struct IvfLayout {
    size_t off_pids, off_pdist, off_coarse, off_part, coarse_bytes, part_bytes, total;
    int np;  // partitions probed per query = min(nprobe, number of centroids)
};
int ivf_layout(const qk_store_t* parent, const qk_store_t* store, int64_t Q, int nprobe, int k, IvfLayout* L) {
    QK_REQUIRE(parent && store && Q > 0 && nprobe > 0, "bad argument");
    QK_REQUIRE(parent->num_lists == 1 && parent->flat_rows > 0, "the parent must be a flat (single-list) store");
    L->np = (int)(nprobe < parent->flat_rows ? nprobe : parent->flat_rows);
    L->coarse_bytes = qk_scan_workspace_bytes(parent, Q, 1, L->np);
    L->part_bytes = store->num_segments > 0 ? qk_scan_workspace_bytes(store, Q, L->np, k) : 256;
    if (L->coarse_bytes == 0 || L->part_bytes == 0) return QK_ERR_INVALID_ARGUMENT;
    // the two scans have their own regions: the partition scan's is cleared (and its pair table filled by the coarse
    // scan's refine kernel) while the coarse scan still runs
    size_t o = 0;
    L->off_pids = o;   o = align_up(o + (size_t)Q * L->np * 8, 256);
    L->off_pdist = o;  o = align_up(o + (size_t)Q * L->np * 4, 256);
    L->off_coarse = o; o = align_up(o + L->coarse_bytes, 256);
    L->off_part = o;   o += L->part_bytes;
    L->total = o;
    return QK_OK;
}
__global__ void fill_empty_result_kernel(int64_t n, int64_t* ids, float* dist, float pad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { ids[i] = -1; dist[i] = pad; }
}
}  // namespace

extern "C" size_t qk_search_ivf_workspace_bytes(const qk_store_t* parent, const qk_store_t* store, int64_t num_queries,
                                                int nprobe, int k) {
    IvfLayout L;
    if (ivf_layout(parent, store, num_queries, nprobe, k, &L) != QK_OK) return 0;
    return L.total;
}

extern "C" int qk_search_ivf(const qk_store_t* parent, const qk_store_t* store, const int32_t* id_to_slot,
                             int64_t table_size, const float* queries, int64_t Q, int64_t q_pitch, int nprobe,
                             int metric, int k, int shard_rank, int shard_world, int64_t* out_ids, float* out_dist,
                             int64_t* out_probe_ids, void* workspace, size_t workspace_bytes, int32_t* stats,
                             void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    IvfLayout L;
    int rc = ivf_layout(parent, store, Q, nprobe, k, &L);
    if (rc) return rc;
    QK_REQUIRE(id_to_slot && table_size > 0 && queries && out_ids && out_dist, "null argument");
    if (!workspace || workspace_bytes < L.total) {
        set_error("workspace too small: need %zu bytes, have %zu", L.total, workspace_bytes);
        return QK_ERR_WORKSPACE;
    }
    char* ws = (char*)workspace;
    int64_t* p_ids = out_probe_ids ? out_probe_ids : (int64_t*)(ws + L.off_pids);
    float* p_dist = (float*)(ws + L.off_pdist);
    if (store->num_segments == 0) {  // every list is empty: padded results (query_coordinator.cpp:589-601)
        ScanExtras cx;
        rc = scan_partitions_impl(parent, queries, Q, q_pitch, nullptr, 1, metric, L.np, p_ids, p_dist, nullptr,
                                  ws + L.off_coarse, L.coarse_bytes, nullptr, stream, cx);
        if (rc) return rc;
        const int64_t n = Q * k;
        fill_empty_result_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
            n, out_ids, out_dist, metric == QK_METRIC_INNER_PRODUCT ? -INFINITY : INFINITY);
        QK_LAUNCHED();
        return QK_OK;
    }
    // 0. the partition scan's workspace clears run on a parallel branch beside the coarse scan (16 MB of candidate
    //    slots at C2: ~6 us that used to sit on the critical path)
    ScanPlan pp;
    if ((rc = make_plan(store, Q, L.np, k, &pp, /*allow_dense=*/false))) return rc;
    char* pws = ws + L.off_part;
    SideStream* side = nullptr;
    if ((rc = side_stream_for_current_device(&side))) return rc;
    QK_CUDA(cudaEventRecord(side->fork2, stream));
    QK_CUDA(cudaStreamWaitEvent(side->stream, side->fork2, 0));
    if ((rc = clear_scan_workspace(pp, pws, Q, side->stream))) return rc;
    QK_CUDA(cudaEventRecord(side->join2, side->stream));
    // 1. coarse centroid scan (query_coordinator.cpp:644): top-np centroids per query, nearest first; flat mode. With
    //    single-segment lists its refine kernel also writes the partition scan's pair table (id -> slot map, shard
    //    filter, per-segment histogram): no separate expansion launch
    const bool fuse = (pp.P == L.np);
    FusedExpand fx;
    fx.pair_seg = (int32_t*)(pws + pp.off_pair_seg);
    fx.seg_count = (int32_t*)(pws + pp.off_seg_count);
    fx.gthr = (uint32_t*)(pws + pp.off_gthr);
    fx.id_to_slot = id_to_slot; fx.table_size = table_size;
    fx.list_seg0 = store->list_seg0; fx.list_nseg = store->list_nseg; fx.num_lists = store->num_lists;
    fx.shard_rank = shard_rank; fx.shard_world = shard_world < 1 ? 1 : shard_world;
    ScanExtras cx;
    if (fuse) cx.fused_expand = &fx;
    // the partition scan does not depend on the order of a query's probes: the coarse scan only has to settle WHICH np
    // centroids are the nearest (out_probe_ids: nearest by filter score first, the rest in no guaranteed order)
    static const bool no_set_mode = getenv("QK_SET_MODE") && atoi(getenv("QK_SET_MODE")) == 0;
    cx.set_mode = no_set_mode ? 0 : 1;
    cx.pre_refine_event = side->join2;  // joins the clearing branch back into the stream
    rc = scan_partitions_impl(parent, queries, Q, q_pitch, nullptr, 1, metric, L.np, p_ids, p_dist, nullptr, ws + L.off_coarse,
                              L.coarse_bytes, nullptr, stream, cx);
    if (rc) return rc;
    // 2. partition scan of the probed lists; the id -> slot map (and the shard filter) run inside the pair expansion
    ScanExtras px;
    px.probe_ids = p_ids;
    px.id_to_slot = id_to_slot;
    px.table_size = table_size;
    px.shard_rank = shard_rank;
    px.shard_world = shard_world < 1 ? 1 : shard_world;
    px.ws_precleared = 1;
    px.pairs_preexpanded = fuse ? 1 : 0;
    return scan_partitions_impl(store, queries, Q, q_pitch, nullptr, L.np, metric, k, out_ids, out_dist, nullptr, pws,
                                L.part_bytes, stats, stream, px);
}

#ifdef QK_STAGE_DEBUG
// debug builds only (not part of the C ABI): per-role wait cycles of the tensor-core scan kernel, see scan_mma.cuh
extern "C" int qk_debug_times(unsigned long long* out, int reset) {
    if (out) QK_CUDA(cudaMemcpyFromSymbol(out, qk::g_dbg_times, sizeof(qk::g_dbg_times)));
    if (reset) {
        void* p = nullptr;
        QK_CUDA(cudaGetSymbolAddress(&p, qk::g_dbg_times));
        QK_CUDA(cudaMemset(p, 0, sizeof(qk::g_dbg_times)));
    }
    return QK_OK;
}
#endif

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
