// Tensor-core variant of the scan (filter) kernel for d <= 128 -- included by scan.cu.
//
// The grouped partition scan is a small dense contraction per work item: [<=32 queries x d] . [rows x d]^T.
// Scoring it on the FP32 pipe is bound by shared-memory operand traffic (an 8x8 register tile per thread would
// be needed to feed FFMA at rate), so this kernel hands the contraction to the 5th-generation tensor cores:
// tcgen05.mma kind::tf32 with M = 128 rows, N = 32 queries, K = 8, operands read straight from the
// TMA-written (128-byte swizzled) shared-memory tiles, accumulators in tensor memory. TF32 keeps only 10
// mantissa bits, far too few for a distance filter, so every product is split as
//       a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo,      x_hi = x truncated to TF32 (what the tensor core reads
//                                                       from an fp32 word), x_lo = x - x_hi (exact in fp32)
// which leaves a relative error of about 2^-20 per dot product (measured: scripts/umma_probe.cu) -- the filter
// only has to be good enough for the exact refine + error-bound proof that follows it (merge_refine_kernel).
// The a_lo tile never touches shared memory: the split warps write it to tensor memory (tcgen05.st) and the
// second MMA of each K-step takes its A operand from there. b_hi and b_lo sit side by side in one K-major tile
// (rows 0-31 / 32-63), so a_hi * [b_hi | b_lo] is ONE MMA with N = 64 whose two halves the epilogue adds: two
// MMAs per K-step instead of three (the single issuing thread is the scarce resource, not the tensor pipe).
//
// Roles (16 warps, one persistent CTA per SM):
//   warp 0      producer: work items (atomic counter, metadata pipelined; or the arithmetic grid of a flat-mode scan),
//               TMA tensor loads of the row tiles (box 128 rows x 32 floats, ring of 8)
//   warps 1, 15 MMA issuers (alternate tiles, one elected lane each): per row box the a_hi products
//               a_hi . [b_hi | b_lo] (N = 2 npad, A from shared memory) as soon as the TMA data has landed, then
//               a_lo . b_hi (N = npad, A from tensor memory) once the split warps delivered a_lo; tcgen05.commit onto
//               the pipeline mbarriers. npad = the item's query slots padded to 16 or 32
//   warps 2-5   split: a_lo = a - trunc(a) -> tensor memory; gather of the item's query chunk (prefetched one item
//               ahead) into the swizzled K-major B tiles, b_hi and b_lo
//   warps 6-13  epilogue + selection, two groups of four taking alternate tiles: tcgen05.ld of the
//               accumulators (thread = row, npad query scores in registers), score vs the query's threshold in the
//               float domain (one FFMA + one compare per score), thread-level append of the rare survivors to the
//               per-query candidate buffers (one global atomic per survivor, issued in batches of four)
//   warp 14     threshold refresh, off the critical path: when a query's fill passes a multiple of 64 the
//               appending thread posts (query, fill) in its warp's mailbox; the refresh warp re-derives the
//               kc-th smallest key appended so far (radix select over the buffer) and publishes it with atomicMin
#pragma once

namespace qk {

static constexpr int MMA_TM = 128;            // rows per tile (UMMA M)
static constexpr int MMA_NQ = 32;             // query slots per item (UMMA N)
static constexpr int MMA_BOX = 32;            // floats per TMA box row (128 B, one swizzle atom row)
static constexpr int MMA_STAGES = 8;          // ring of A boxes, 16 KB each (two whole 128-row tiles at d = 128)
static constexpr int MMA_BOX_BYTES = MMA_TM * 128;
static constexpr int MMA_SUB_ROWS = 64;       // rows of the half-height box a list's short last tile is loaded with
static constexpr int MMA_NB = 2;              // B-operand slots (query chunk hi + lo, 32 KB each)
static constexpr int MMA_ND = 5;              // work-item descriptor slots (the selection warps lag the MMAs)
static constexpr int MMA_BBOX_BYTES = 2 * MMA_NQ * 128;  // 8 KB: (32 queries hi + 32 queries lo) x 32 floats
static constexpr int MMA_THREADS = 32 * 16;
static constexpr int MMA_SMEM_HEADER = 4096;   // mbarriers, mailboxes, descriptor ring
static constexpr int MMA_REFRESH_CAP = 1024;   // candidates a refresh looks at (any subset gives a valid bound)
static constexpr int MMA_TMEM_COLS = 512;
static constexpr int MMA_NACC = 4;            // accumulator buffers: the epilogue may lag the MMAs by that many tiles
static constexpr int MMA_TMEM_D = 0;          // MMA_NACC accumulator buffers x 64 columns (a.b_hi | a_hi.b_lo)
static constexpr int MMA_TMEM_ALO = 256;      // MMA_STAGES a_lo boxes x 32 columns

// -DQK_STAGE_DEBUG builds (scripts/gpu_roles.sh): cycles every role spends in its waits, accumulated per CTA in a
// device array -- [launch kind: 0 partition scan, 1 flat / coarse][CTA][64 counters]
#ifdef QK_STAGE_DEBUG
__device__ unsigned long long g_dbg_times[2 * 148 * 64];
#define QK_TDECL long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long dbg_t00 = clock64();
#define QK_TWAIT(slot, stmt) { const long long t0_ = clock64(); stmt; dbg_acc[slot] += clock64() - t0_; }
#define QK_TFLUSH(base) { dbg_acc[0] = clock64() - dbg_t00; if (lane == 0 && blockIdx.x < 148) { \
    for (int i_ = 0; i_ < 8; ++i_) atomicAdd(&g_dbg_times[((a.flat ? 148 : 0) + blockIdx.x) * 64 + (base) + i_], (unsigned long long)dbg_acc[i_]); } }
#else
#define QK_TDECL
#define QK_TWAIT(slot, stmt) { stmt; }
#define QK_TFLUSH(base) {}
#endif

struct MmaDesc {  // published in shared memory by the producer warp for every work item in flight
    WorkItem w;
    int q[MMA_NQ];       // query index of every query slot (-1: unused slot)
    float limf[MMA_NQ];  // the queries' filter-score thresholds (float domain; +inf = none yet, -inf = unused slot)
    float delta[MMA_NQ]; // top-1 mode: the margin added to a running minimum to make it a threshold (see ScanArgs::top1)
};

static size_t scan_mma_smem_bytes() {
    return (size_t)MMA_SMEM_HEADER + (size_t)MMA_STAGES * MMA_BOX_BYTES + (size_t)MMA_NB * 4 * MMA_BBOX_BYTES +
           (size_t)2 * 256 * sizeof(uint32_t) + (size_t)2 * MMA_REFRESH_CAP * sizeof(uint32_t) + 1024;
}

// ---- tcgen05 wrappers --------------------------------------------------------------------------------------
// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor: D = f32, A = B = tf32, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// The MMA / commit wrappers are called by ALL lanes of the (converged) issuing warp; one elected lane issues.
// Keeping the election inside the asm spares the per-instruction "elect / retry" loop the compiler wraps around a
// warp-uniform instruction that sits in a lane-divergent branch.
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
// All four K-steps of one full row box (32 floats) in ONE asm block: a single election, the descriptors of the later
// K-steps derived inside (a K-step advances both start addresses by 32 B = 2 descriptor units), no branch in between.
// Issued one by one from C++ every MMA cost ~25 instructions (election, divergence check, five R2UR moves, 64-bit
// descriptor adds, a bounds branch) and the issuing warps were the longest stage a tile went through.
__device__ __forceinline__ void umma_box_ss4(uint32_t tmem_d, uint64_t da0, uint64_t db0, uint32_t idesc, uint32_t acc0) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "add.u64 a1, %1, 2;\n\tadd.u64 a2, %1, 4;\n\tadd.u64 a3, %1, 6;\n\t"
        "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a1, b1, %3, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a2, b2, %3, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a3, b3, %3, 1;\n\t}\n" ::"r"(tmem_d),
        "l"(da0), "l"(db0), "r"(idesc), "r"(acc0)
        : "memory");
}
// the same for the 3-term filter: a_hi . [b_hi | b_lo] (A from shared memory) and a_lo . b_hi (A from tensor memory, 8
// columns per K-step) interleaved per K-step
__device__ __forceinline__ void umma_box_ss4_ts4(uint32_t tmem_d, uint64_t da0, uint64_t db0, uint32_t talo, uint32_t idesc_hl,
                                                 uint32_t idesc_h, uint32_t acc0) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t.reg .b32 t1, t2, t3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "add.u64 a1, %1, 2;\n\tadd.u64 a2, %1, 4;\n\tadd.u64 a3, %1, 6;\n\t"
        "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
        "add.u32 t1, %3, 8;\n\tadd.u32 t2, %3, 16;\n\tadd.u32 t3, %3, 24;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %4, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%3], %2, %5, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a1, b1, %4, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [t1], b1, %5, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a2, b2, %4, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [t2], b2, %5, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a3, b3, %4, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [t3], b3, %5, 1;\n\t}\n" ::"r"(tmem_d),
        "l"(da0), "l"(db0), "r"(talo), "r"(idesc_hl), "r"(idesc_h), "r"(acc0)
        : "memory");
}
// arrive on `bar` once every tcgen05 operation issued so far by the elected thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
          "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// kc-th smallest of one key per lane (kc <= 32), by bisection on the key bits
__device__ __forceinline__ uint32_t warp_kth_smallest(uint32_t key, int kc) {
    uint32_t lo = 0;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = lo | (1u << bit);
        if (__popc(__ballot_sync(0xffffffffu, key < cand)) < kc) lo = cand;
    }
    return lo;
}

// v[g] for a run-time g without spilling v[] to local memory: a 5-level select tree (31 SELs)
__device__ __forceinline__ uint32_t pick32(const uint32_t (&v)[32], int g) {
    uint32_t a[16], b[8], c[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (g & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (g & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = (g & 4) ? b[2 * i + 1] : b[2 * i];
    const uint32_t d0 = (g & 8) ? c[1] : c[0], d1 = (g & 8) ? c[3] : c[2];
    return (g & 16) ? d1 : d0;
}

// threshold key -> float-domain limit: KEY_MAX (no threshold yet) admits everything but NaN
__device__ __forceinline__ float key2lim(uint32_t t) { return t == KEY_MAX ? INFINITY : key2f(t); }

template <bool kIP>
__global__ void __launch_bounds__(MMA_THREADS, 1) scan_mma_kernel(const ScanArgs a, const __grid_constant__ CUtensorMap vmap,
                                                                const __grid_constant__ CUtensorMap vmap_sub) {
    constexpr int TM = MMA_TM, NB = MMA_NB, ND = MMA_ND;
    static_assert(MMA_STAGES <= 16 && 768 + MMA_ND * sizeof(MmaDesc) <= 3072 && 3072 + 32 * 8 <= MMA_SMEM_HEADER, "descriptor ring");
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* a_full = bars;           // [NS] producer (tx)                 -> split warps, MMA issuer
    uint64_t* a_empty = bars + 16;     // [NS] 4 split warps + MMA commit    -> producer
    uint64_t* alo_full = bars + 32;   // [NS] 4 split warps                 -> MMA issuer
    uint64_t* b_ready = bars + 48;     // [NB] 4 split warps (b_lo written)  -> MMA issuer
    uint64_t* b_empty = bars + 50;     // [NB] MMA commit                    -> producer
    uint64_t* i_full = bars + 52;      // [ND] producer (descriptor written) -> every consumer warp
    uint64_t* i_empty = bars + 60;     // [ND] 13 consumer warps             -> producer
    uint64_t* d_full = bars + 68;      // [NACC] MMA commit                  -> epilogue group
    uint64_t* d_empty = bars + 72;     // [NACC] 4 epilogue warps            -> MMA issuer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 640);
    volatile uint32_t* ep_done = reinterpret_cast<volatile uint32_t*>(smem_raw + 644);       // epilogue warps that left
    // [32] refresh requests: four mailboxes per epilogue warp (a lane posts into box lane & 3), so that a burst of
    // requests -- the first lists of many queries, while the thresholds are still loose -- queues up instead of being
    // dropped (with one box per warp more than half of the requests were)
    volatile unsigned long long* mbox = reinterpret_cast<volatile unsigned long long*>(smem_raw + 3072);
    MmaDesc* descs = reinterpret_cast<MmaDesc*>(smem_raw + 768);    // [ND]
    unsigned char* As = smem_raw + MMA_SMEM_HEADER;                 // [NS][128 rows][128 B]
    unsigned char* Bs = As + (size_t)MMA_STAGES * MMA_BOX_BYTES;    // [NB][4 boxes][hi: 32 rows | lo: 32 rows][128 B]
    uint32_t* hists = reinterpret_cast<uint32_t*>(Bs + (size_t)NB * 4 * MMA_BBOX_BYTES);  // [2 refresh warps][256]
    uint32_t* rscratch = hists + 2 * 256;                                                  // [2][MMA_REFRESH_CAP] keys

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dp = a.dp, kc = a.kc;
    const int nbox = (dp + MMA_BOX - 1) / MMA_BOX;  // 1..4
    // the refine kernel (a programmatic dependent, common.cuh) may be scheduled as this grid's CTAs leave; it waits for
    // the grid's completion itself before it reads a candidate
    pdl_launch_dependents();
    // Ring slots in use: a multiple of two tiles' worth of boxes (8, 8, 6, 8 for 1..4 boxes per tile), so that each of
    // the two MMA issuers (alternate tiles) meets every slot it uses in EVERY round. With 3 boxes per tile on 8 slots an
    // issuer would skip rounds of a slot; its parity wait could then be satisfied by the completion of an older round
    // (mbarrier waits only tell "the phase of this parity is over") -- seen as a pipeline deadlock at d = 96 in the
    // 2-term mode, where nothing else orders the issuer behind the data.
    const int NS = (MMA_STAGES / (2 * nbox)) * 2 * nbox;
    if (tid == 0) {
        // a row box is released by the MMA commit and -- when the a_lo term is computed (3 terms) -- by the 4 split warps
        for (int s = 0; s < MMA_STAGES; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, a.terms == 2 ? 1 : 5); mbar_init(alo_full + s, 4); }
        for (int s = 0; s < NB; ++s) { mbar_init(b_ready + s, 4); mbar_init(b_empty + s, 2); }
        for (int s = 0; s < ND; ++s) { mbar_init(i_full + s, 1); mbar_init(i_empty + s, 14); }
        for (int s = 0; s < MMA_NACC; ++s) { mbar_init(d_full + s, 1); mbar_init(d_empty + s, 4); }
        mbar_fence_init();
        *ep_done = 0;
        for (int s = 0; s < 32; ++s) mbox[s] = 0ull;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(MMA_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================================================================== producer
        // Work items are taken from an atomic counter several items ahead; every metadata load of the chain
        // index -> item -> query ids -> thresholds is consumed one iteration after it was issued, so none of them
        // is waited for. The descriptor of item n+1 is published BEFORE the row tiles of item n are issued, so
        // the split warps can prefetch its query chunk (B operand) while item n streams.
        QK_TDECL
        const int n_items = a.flat ? a.flat_items : a.ctrl[1];
        auto fetch_index = [&]() {  // lane 0 holds the result; broadcast where it is consumed
            int it = 0;
            if (lane == 0) it = atomicAdd(&a.ctrl[0], 1);
            return it;
        };
        auto fetch_item = [&](int it) {
            WorkItem w;
            w.seg = -1; w.g_begin = 0; w.g_cnt = 0; w.nrows = 0; w.row0 = 0; w.pad_ = 0;
            if (it < n_items) {
                if (a.flat) {  // the grid (segment x query chunk), chunk fastest: concurrent CTAs share the segment's rows
                    flat_item(a, it, w.seg, w.g_begin, w.g_cnt);
                    w.nrows = a.seg_rows[w.seg];
                    w.row0 = a.seg_row0[w.seg];
                } else {
                    w = a.items[it];
                }
            }
            return w;
        };
        // query index of this lane's slot (the pair index is query * P + slot)
        auto fetch_query = [&](const WorkItem& w) {
            if (w.seg < 0 || lane >= w.g_cnt) return -1;
            return a.flat ? w.g_begin + lane : a.seg_pairs[w.g_begin + lane] / a.P;
        };
        auto fetch_gthr = [&](int q) { return q >= 0 ? __ldcg(a.gthr + q) : KEY_MAX; };
        auto fetch_delta = [&](int q) { return (a.top1 && q >= 0) ? __ldg(a.qdelta + q) : 0.f; };
        // descriptor of item n; false when n is past the last item (the sentinel is published instead)
        auto issue_desc_b = [&](uint32_t n, const WorkItem& m, int q, uint32_t gthr, float delta) {
            const int id = n % ND;
            QK_TWAIT(2, mbar_wait(i_empty + id, ((n / ND) & 1u) ^ 1u));
            if (m.seg < 0) {
                if (lane == 0) {
                    descs[id].w.seg = -1;
                    mbar_arrive(i_full + id);
                }
                return false;
            }
            descs[id].q[lane] = q;
            descs[id].limf[lane] = q >= 0 ? key2lim(gthr) : -INFINITY;
            descs[id].delta[lane] = delta;
            if (lane == 0) descs[id].w = m;
            __syncwarp();
            if (lane == 0) mbar_arrive(i_full + id);
            return true;
        };
        // Metadata pipeline, one stage per loop iteration: index (atomic) -> item -> query ids -> thresholds. Every load
        // is issued a whole iteration (one item's worth of TMA issue) before its result is used -- with the thresholds
        // fetched in the same iteration as their use (as this loop used to) the producer stalled for a full L2/DRAM
        // round trip per item, ~3 us of ~4, and the ring ran half empty (scripts/role_probe.py).
        WorkItem cur = fetch_item(__shfl_sync(0xffffffffu, fetch_index(), 0));
        WorkItem m1 = fetch_item(__shfl_sync(0xffffffffu, fetch_index(), 0));
        WorkItem m2 = fetch_item(__shfl_sync(0xffffffffu, fetch_index(), 0));
        WorkItem m3 = fetch_item(__shfl_sync(0xffffffffu, fetch_index(), 0));
        WorkItem m4 = fetch_item(__shfl_sync(0xffffffffu, fetch_index(), 0));
        int i5 = fetch_index();
        int q1 = fetch_query(m1), q2 = fetch_query(m2), q3 = fetch_query(m3);
        uint32_t g1 = fetch_gthr(q1), g2 = fetch_gthr(q2);
        float d1 = fetch_delta(q1), d2 = fetch_delta(q2);
        bool live;
        {
            const int q0 = fetch_query(cur);
            live = issue_desc_b(0, cur, q0, fetch_gthr(q0), fetch_delta(q0));
        }
        uint32_t U = 0;
        for (uint32_t n = 0; live; ++n) {
            const int i6 = fetch_index();  // consumed two iterations from now
            bool next_live;
            QK_TWAIT(4, next_live = issue_desc_b(n + 1, m1, q1, g1, d1));
            // ---- row tiles of item n
            // lane b issues box b of the tile: the (up to four) slot waits and TMA issues of a tile run side by side
            // instead of one after the other (~160 cycles each: the producer was busy, not blocked, most of the time)
            const int ntiles = (cur.nrows + TM - 1) / TM;
            for (int tile = 0; tile < ntiles; ++tile, U += nbox) {
                QK_TWAIT(1, if (lane < nbox) {
                    const uint32_t Ub = U + lane;
                    const int st = Ub % NS;
                    mbar_wait(a_empty + st, ((Ub / NS) & 1u) ^ 1u);
                    const int tr = cur.nrows - tile * TM;  // rows of this tile that belong to the list
                    unsigned char* dst = As + (size_t)st * MMA_BOX_BYTES;
                    const int y0 = (int)(cur.row0 + (int64_t)tile * TM);
                    if (tr > MMA_SUB_ROWS) {
                        mbar_expect_tx(a_full + st, (uint32_t)MMA_BOX_BYTES);
                        tma_load_2d(dst, &vmap, lane * MMA_BOX, y0, a_full + st);
                    } else {
                        // a short last tile of a list: ONE half-height box. The rows after it keep whatever an earlier
                        // tile left in the slot (finite data, or anything at all in the first round: every row's scores
                        // depend on that row alone, and the epilogue masks them); loading them would re-read the head
                        // of the next list. (Finer sub-boxes -- 32 rows, up to three per slot -- cut the DRAM traffic
                        // further but cost more in TMA issues than they saved: 134.7 vs 131.6 us at C2.)
                        mbar_expect_tx(a_full + st, (uint32_t)(MMA_SUB_ROWS * 128));
                        tma_load_2d(dst, &vmap_sub, lane * MMA_BOX, y0, a_full + st);
                    }
                }
                __syncwarp());
            }
            // ---- rotate the metadata pipeline (every right-hand side was loaded at least one iteration ago)
            QK_TWAIT(5,
            cur = m1;
            m1 = m2; q1 = q2; g1 = g2; d1 = d2;
            m2 = m3; q2 = q3; g2 = fetch_gthr(q2); d2 = fetch_delta(q2);
            m3 = m4; q3 = fetch_query(m3);
            m4 = fetch_item(__shfl_sync(0xffffffffu, i5, 0));
            i5 = i6;
            live = next_live);
        }
        QK_TFLUSH(0)
    } else if (warp == 1 || warp == 15) {
        // ===================================================================== MMA issuers
        // Issuer mi takes the tiles with (T & 1) == mi. tcgen05.commit only tracks the issuing thread's own MMAs: the
        // row boxes and the accumulator of a tile are released by the issuer that consumed them, the B slot of an item
        // by both (count 2).
        // Everything this role computes with is the same in all lanes; values that come out of shared memory or the
        // thread index are broadcast from lane 0 (the compiler treats the result of such a shuffle as warp-uniform), so
        // that the tile / ring counters and the MMA descriptors live in uniform registers. Without it every MMA was
        // preceded by ~20 instructions of R2UR moves and divergence bookkeeping: the two issuers spent 65 of the
        // kernel's 114 us issuing (scripts/role_probe.py), and a tile sat in the two-tile ring for that long.
        const int mi = __shfl_sync(0xffffffffu, warp == 1 ? 0 : 1, 0);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        QK_TDECL
        uint32_t T = 0, U = 0;
        for (uint32_t n = 0;; ++n) {
            const int ib = n % NB, id = n % ND;
            QK_TWAIT(1, mbar_wait(i_full + id, (n / ND) & 1u));
            WorkItem d;
            d.seg = __shfl_sync(0xffffffffu, descs[id].w.seg, 0);
            d.nrows = __shfl_sync(0xffffffffu, descs[id].w.nrows, 0);
            d.g_cnt = __shfl_sync(0xffffffffu, descs[id].w.g_cnt, 0);
            if (d.seg < 0) break;
            __syncwarp();
            if (lane == 0) mbar_arrive(i_empty + id);  // only the descriptor header was needed
            // query slots padded to 16 or 32: b_hi in rows [0, npad) of every B box, b_lo in rows [npad, 2 npad)
            const int npad = (d.g_cnt <= 16 && !QK_DBG(a, 16)) ? 16 : 32;
            const uint32_t idesc_hl = umma_idesc_tf32(MMA_TM, 2 * npad);  // a_hi . [b_hi | b_lo]
            const uint32_t idesc_h = umma_idesc_tf32(MMA_TM, npad);       // a_lo . b_hi
            QK_TWAIT(2, mbar_wait(b_ready + ib, (n / NB) & 1u));
            const uint32_t bs = smem_u32(Bs + (size_t)ib * 4 * MMA_BBOX_BYTES);
            const int ntiles = (d.nrows + TM - 1) / TM;
            for (int tile = 0; tile < ntiles; ++tile, ++T) {
                const int db = T % MMA_NACC;
                if ((int)(T & 1u) != mi) { U += nbox; continue; }
                QK_TWAIT(3, mbar_wait(d_empty + db, ((T / MMA_NACC) & 1u) ^ 1u));
                const uint32_t tmem_d = tmem_u + MMA_TMEM_D + db * (2 * MMA_NQ);
                for (int b = 0; b < nbox; ++b, ++U) {
                    const int st = U % NS;
                    const uint64_t da0 = umma_desc_sw128(smem_u32(As + (size_t)st * MMA_BOX_BYTES));
                    const uint64_t db0 = umma_desc_sw128(bs + b * MMA_BBOX_BYTES);
                    const uint32_t talo = tmem_u + MMA_TMEM_ALO + st * MMA_BOX;
                    const int ksteps = min(4, (dp - b * MMA_BOX + 7) >> 3);
                    if (a.terms == 2) {
                        // 2xTF32: a_hi . (b_hi + b_lo) only -- no a_lo term, nothing to wait for but the TMA data
                        QK_TWAIT(4, mbar_wait(a_full + st, (U / NS) & 1u));
                        tc_fence_after();
                        if (ksteps == 4) {
                            umma_box_ss4(tmem_d, da0, db0, idesc_hl, b ? 1u : 0u);
                        } else {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                if (kk < ksteps) umma_ss(tmem_d, da0 + 2 * kk, db0 + 2 * kk, idesc_hl, (b | kk) ? 1u : 0u);
                        }
                    } else if (QK_DBG(a, 8)) {
                        // the a_hi products only need the TMA data: they run while the split warps still derive a_lo
                        mbar_wait(a_full + st, (U / NS) & 1u);
                        tc_fence_after();
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)  // a K-step advances both start addresses by 32 B (2 descriptor units)
                            if (kk < ksteps) umma_ss(tmem_d, da0 + 2 * kk, db0 + 2 * kk, idesc_hl, (b | kk) ? 1u : 0u);
                        mbar_wait(alo_full + st, (U / NS) & 1u);
                        tc_fence_after();
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            if (kk < ksteps) umma_ts(tmem_d, talo + kk * 8, db0 + 2 * kk, idesc_h, 1u);
                    } else {
                        // the split warps waited for the TMA data of this box themselves: a_lo ready => a_hi ready
                        QK_TWAIT(4, mbar_wait(alo_full + st, (U / NS) & 1u));
                        tc_fence_after();
                        if (ksteps == 4 && !(QK_DBG(a, 2))) {
                            umma_box_ss4_ts4(tmem_d, da0, db0, talo, idesc_hl, idesc_h, b ? 1u : 0u);
                        } else {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                if (kk < ksteps && !(QK_DBG(a, 2))) {  // a K-step advances both start addresses by 32 B (2 descriptor units)
                                    umma_ss(tmem_d, da0 + 2 * kk, db0 + 2 * kk, idesc_hl, (b | kk) ? 1u : 0u);
                                    umma_ts(tmem_d, talo + kk * 8, db0 + 2 * kk, idesc_h, 1u);
                                }
                            }
                        }
                    }
                    umma_commit(a_empty + st);  // the row box (and its a_lo columns) may be refilled
                }
                umma_commit(d_full + db);
            }
            umma_commit(b_empty + ib);
        }
        QK_TFLUSH(mi ? 16 : 8)
    } else if (warp < 6) {
        // ===================================================================== split warps
        // Besides a_lo, these 128 threads build the B operand of every item: thread (warp w, lane c) owns the
        // 16-byte chunk c of queries g = 4j + w (j < 8). The chunks of item n+1 are loaded into registers before
        // the row boxes of item n are processed and written (b_hi and b_lo = b - tf32(b)) at the top of the next
        // iteration, so the gather latency is hidden behind a whole item.
        //   B tile layout: box b = 32 floats of every query; b_hi of query g at g * 128 B, b_lo at (32 + g) * 128 B,
        //   16-byte chunk cc stored at cc ^ (g & 7) (canonical K-major SWIZZLE_128B); padding chunks are zero
        const int q4 = warp & 3;  // the tensor-memory lane quadrant this warp may access
        const int sw = warp - 2;
        const int dp4 = dp >> 2;
        const bool c_used = lane < nbox * 8;  // chunk column inside the boxes the MMA reads
        uint32_t U = 0;
        float4 bq[8];
        QK_TDECL
        auto prefetch_b = [&](uint32_t n, WorkItem& w) {  // reads descriptor n, issues the loads, releases it
            const int id = n % ND;
            QK_TWAIT(1, mbar_wait(i_full + id, (n / ND) & 1u));
            w = descs[id].w;
            if (w.seg >= 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int g = 4 * j + sw;
                    bq[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (g < w.g_cnt && lane < dp4)
                        bq[j] = __ldg(reinterpret_cast<const float4*>(a.queries + (int64_t)descs[id].q[g] * a.q_pitch) + lane);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(i_empty + id);
        };
        WorkItem d;
        prefetch_b(0, d);
        for (uint32_t n = 0;; ++n) {
            if (d.seg < 0) break;
            const int ib = n % NB;
            QK_TWAIT(2, mbar_wait(b_empty + ib, ((n / NB) & 1u) ^ 1u));  // the MMAs of item n - NB have completed
            if (c_used) {
                unsigned char* bs = Bs + (size_t)ib * 4 * MMA_BBOX_BYTES + (size_t)(lane >> 3) * MMA_BBOX_BYTES;
                const int npad = (d.g_cnt <= 16 && !QK_DBG(a, 16)) ? 16 : 32;
                // slots g_cnt .. npad-1 are read by the MMA too: whatever an earlier item left there is finite query
                // data (or the zeros of the first fill below), and their scores are masked in the epilogue
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int g = 4 * j + sw;
                    if (g < d.g_cnt || (n < (uint32_t)NB && g < npad)) {
                        unsigned char* hp = bs + g * 128 + (((lane & 7) ^ (g & 7)) << 4);
                        const float4 x = g < d.g_cnt ? bq[j] : make_float4(0.f, 0.f, 0.f, 0.f);
                        *reinterpret_cast<float4*>(hp) = x;
                        *reinterpret_cast<float4*>(hp + npad * 128) =
                            make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
                    }
                }
            }
            fence_proxy_async();  // the tensor core reads B through the async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(b_ready + ib);
            const int nrows = d.nrows;
            prefetch_b(n + 1, d);  // d now describes item n+1
            const int ntiles = a.terms == 2 ? 0 : (nrows + TM - 1) / TM;  // 2xTF32: no a_lo term, the boxes are not touched here
            const int r = q4 * 32 + lane;  // this thread's row of the tile == its tensor-memory lane
            for (int tile = 0; tile < ntiles; ++tile) {
                for (int b = 0; b < nbox; ++b, ++U) {
                    const int st = U % NS;
                    QK_TWAIT(3, mbar_wait(a_full + st, (U / NS) & 1u));
                    const unsigned char* rowp = As + (size_t)st * MMA_BOX_BYTES + r * 128;
                    uint32_t v[32];
                    if (!(QK_DBG(a, 4)))
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        const float4 x = *reinterpret_cast<const float4*>(rowp + ((cc ^ (r & 7)) << 4));
                        v[4 * cc + 0] = __float_as_uint(tf32_lo(x.x));
                        v[4 * cc + 1] = __float_as_uint(tf32_lo(x.y));
                        v[4 * cc + 2] = __float_as_uint(tf32_lo(x.z));
                        v[4 * cc + 3] = __float_as_uint(tf32_lo(x.w));
                    }
                    if (!(QK_DBG(a, 4))) tmem_st32(tmem + MMA_TMEM_ALO + st * MMA_BOX + ((uint32_t)(q4 * 32) << 16), v);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(alo_full + st); mbar_arrive(a_empty + st); }
                }
            }
        }
        if (warp == 2) QK_TFLUSH(24)
    } else if (warp < 14) {
        // ===================================================================== epilogue + selection
        const int eg = (warp - 6) >> 2;  // group: takes tiles with (T & 1) == eg
        const int q4 = warp & 3;
        const int qcap = a.qcap;
        volatile unsigned long long* my_box = mbox + (warp - 6) * 4 + (lane & a.refresh_boxes);
        const int rmask = a.refresh_step - 1;  // a refresh is requested whenever a query's fill passes a multiple of this
        int dropped = 0;
        QK_TDECL
        uint32_t T = 0;
        for (uint32_t n = 0;; ++n) {
            const int id = n % ND;
            QK_TWAIT(1, mbar_wait(i_full + id, (n / ND) & 1u));
            const WorkItem d = descs[id].w;
            if (d.seg < 0) break;
            const int g_cnt = d.g_cnt;
            const int npad = (g_cnt <= 16 && !QK_DBG(a, 16)) ? 16 : 32;
            const uint32_t gvalid = g_cnt >= 32 ? 0xffffffffu : ((1u << g_cnt) - 1u);
            const int* dq = descs[id].q;
            float* limf = descs[id].limf;
            const int my_q = dq[lane];  // lane g watches query slot g's global threshold
            const int ntiles = (d.nrows + TM - 1) / TM;
            for (int tile = 0; tile < ntiles; ++tile, ++T) {
                if ((int)(T & 1u) != eg) continue;
                const int tr = min(TM, d.nrows - tile * TM);
                const int r = q4 * 32 + lane;
                // what every other SM has learnt about the queries meanwhile (folded in after this tile)
                const uint32_t g_now = my_q >= 0 ? __ldcg(a.gthr + my_q) : KEY_MAX;
                float nrm = 0.f;
                if (!kIP && r < tr) nrm = __ldg(a.norms + d.row0 + (int64_t)tile * TM + r);
                const int db = T % MMA_NACC;  // group eg sees the buffers with db & 1 == eg
                QK_TWAIT(2, mbar_wait(d_full + db, (T / MMA_NACC) & 1u));
                tc_fence_after();
#ifdef QK_STAGE_DEBUG
                const long long t_tile0 = clock64();
#endif
                // dot = (a_hi + a_lo) . b_hi [columns 0 .. npad-1] + a_hi . b_lo [columns npad .. 2 npad - 1]
                uint32_t v[32];
                {
                    const uint32_t td = tmem + MMA_TMEM_D + db * (2 * MMA_NQ) + ((uint32_t)(q4 * 32) << 16);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (h * 16 < npad) {  // warp-uniform
                            uint32_t x[16], y[16];
                            tmem_ld16_nowait(td + h * 16, x);
                            tmem_ld16_nowait(td + npad + h * 16, y);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[h * 16 + i] = __float_as_uint(__uint_as_float(x[i]) + __uint_as_float(y[i]));
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[h * 16 + i] = 0u;
                        }
                    }
                }
                tc_fence_before();
                // fold in what every SM has learnt about the queries meanwhile (read before the wait above) BEFORE this
                // tile is scored; every warp does it for itself -- all values ever written are valid upper bounds, so
                // the races between the warps of an item are benign
                if (my_q >= 0 && !a.dense) {
                    const float f = key2lim(g_now);
                    if (f < limf[lane]) limf[lane] = f;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(d_empty + db);
                if (QK_DBG(a, 1)) continue;
                if (a.top1 && !a.dense) {
                    // top-1 mode (k = 1: the k-means assign, add()'s nearest-centroid search): the running minimum of a
                    // query's scores plus a margin of a few filter error bounds IS a valid threshold -- every row that
                    // can still be the exact nearest scores under it -- and it is known after the first tile instead of
                    // after kc survivors. Per slot: warp-wide minimum of this tile's 32 rows (one REDUX), folded into
                    // the item's limit and published with atomicMin when it improves it.
                    const float* delta = descs[id].delta;
                    uint32_t my_min = KEY_MAX;
#pragma unroll
                    for (int g = 0; g < MMA_NQ; ++g) {
                        if (g >= npad) break;  // warp-uniform
                        const float dot = __uint_as_float(v[g]);
                        const float sc = kIP ? -dot : fmaf(-2.f, dot, nrm);
                        const uint32_t key = r < tr ? f2key(sc) : KEY_MAX;
                        const uint32_t wmin = __reduce_min_sync(0xffffffffu, key);
                        if (lane == g) my_min = wmin;
                    }
                    if (lane < g_cnt && my_min != KEY_MAX) {
                        const float cand = __fadd_ru(key2f(my_min), delta[lane]);
                        if (cand < limf[lane]) {
                            limf[lane] = cand;  // benign race between the warps of the item: every value written is valid
                            atomicMin(a.gthr + dq[lane], f2key(cand));
                        }
                    }
                    __syncwarp();
                }
                // ---- scores of this thread's row against the 32 query slots; bit g of pm: the score passes
                uint32_t pm = 0;
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4) {
                    if (4 * g4 >= npad) break;  // warp-uniform: only the padded slot count is scored
                    const float4 L = reinterpret_cast<const float4*>(limf)[g4];
                    const float lim[4] = {L.x, L.y, L.z, L.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int g = 4 * g4 + j;
                        const float dot = __uint_as_float(v[g]);
                        const float sc = kIP ? -dot : fmaf(-2.f, dot, nrm);
                        v[g] = __float_as_uint(sc);
                        pm |= (sc <= lim[j]) ? (1u << g) : 0u;
                    }
                }
                if (a.dense) {  // dense mode: every score goes out, the per-query select follows the kernel
                    if (r < tr) {
                        const long long lrow = d.row0 + (long long)tile * TM + r - a.dense_row0;
#pragma unroll
                        for (int g = 0; g < MMA_NQ; ++g)
                            if (g < g_cnt) a.dense[(size_t)dq[g] * a.dense_rows + lrow] = f2key(__uint_as_float(v[g]));
                    }
#ifdef QK_STAGE_DEBUG
                    dbg_acc[3] += clock64() - t_tile0;
#endif
                    continue;
                }
                pm &= gvalid;
                if (r >= tr) pm = 0;
                // ---- survivors: one atomic each (issued four at a time), then the entry stores
                const uint32_t arow = (uint32_t)(d.row0 + (int64_t)tile * TM + r);
                while (__any_sync(0xffffffffu, pm != 0)) {
                    int qs[4], base[4];
                    uint32_t key[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        qs[i] = -1;
                        key[i] = 0;
                        if (pm) {
                            const int g = __ffs(pm) - 1;
                            pm &= pm - 1;
                            qs[i] = dq[g];
                            key[i] = f2key(__uint_as_float(pick32(v, g)));
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) base[i] = qs[i] >= 0 ? atomicAdd(&a.qcount[(size_t)qs[i] * a.qstride], 1) : 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (qs[i] >= 0) {
                            const int slot = base[i];
                            if (slot < qcap) a.qbuf[(size_t)qs[i] * qcap + slot] = ((uint64_t)key[i] << 32) | arow;
                            // the fill passed a multiple of 64: ask a refresh warp for a new threshold (a busy
                            // mailbox just drops the request -- thresholds are an optimisation)
                            if (((slot & rmask) == rmask || slot + 1 == a.refresh_first) && slot + 1 >= kc && !a.fixed_thr && !a.top1) {
                                if (*my_box == 0ull) *my_box = ((unsigned long long)(qs[i] + 1) << 32) | (uint32_t)(slot + 1);
                                else ++dropped;
                            }
                        }
                    }
                }
#ifdef QK_STAGE_DEBUG
                dbg_acc[3] += clock64() - t_tile0;
#endif
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(i_empty + id);
        }
        if (warp == 6) QK_TFLUSH(32)
        if (warp == 10) QK_TFLUSH(40)
        dropped = __reduce_add_sync(0xffffffffu, dropped);
        if (lane == 0 && dropped) atomicAdd(&a.ctrl[9], dropped);  // statistics: refresh requests that found the mailbox busy
        __syncwarp();
        if (lane == 0) atomicAdd(const_cast<uint32_t*>(ep_done), 1u);
    } else {
        // ===================================================================== threshold refresh (off the critical path)
        // Up to 256 entries (8 keys per lane) are selected in registers: common prefix of the keys from a warp min / max,
        // then one bit per round -- 8 compares and one REDUX -- with no shared memory, histogram or atomics. Two
        // requests are taken per sweep so that their candidate loads (an L2 round trip each) overlap. This warp used
        // to spend ~6 us per request on a 4-pass shared-memory radix select and more than half of the requests found
        // their mailbox busy (stats[6], stats[7]); larger windows (kc > 64) still take that path.
        uint32_t* hist = hists;
        uint32_t* keys = rscratch;
        const int qcap = a.qcap;
        int served = 0, rr = 0;
        int ncap = 4 * kc > a.refresh_window ? 4 * kc : a.refresh_window;
        if (ncap > MMA_REFRESH_CAP) ncap = MMA_REFRESH_CAP;
        auto decode = [&](unsigned long long req, int& q, int& n) {
            q = (int)(req >> 32) - 1;
            int fill = (int)(uint32_t)req;
            if (a.refresh_fresh) fill = __ldcg(a.qcount + (size_t)q * a.qstride);  // what the buffer holds NOW
            if (fill > qcap) fill = qcap;
            // the most recent entries carry the tightest keys; any subset yields a valid upper bound
            n = fill < ncap ? fill : ncap;
            return reinterpret_cast<const unsigned long long*>(a.qbuf) + (size_t)q * qcap + (fill - n);
        };
        auto load8 = [&](const unsigned long long* qb, int n, unsigned long long (&e)[8]) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int idx = j * 32 + lane;
                e[j] = idx < n ? __ldcg(qb + idx) : ~0ull;
            }
        };
        auto select8 = [&](const unsigned long long (&e)[8]) -> uint32_t {  // kc-th smallest key of <= 256 entries
            uint32_t key[8], mn = KEY_MAX, mx = 0u;
            int valid = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                key[j] = (uint32_t)(e[j] >> 32);
                if (key[j] != KEY_MAX) { mn = min(mn, key[j]); mx = max(mx, key[j]); ++valid; }
            }
            mn = __reduce_min_sync(0xffffffffu, mn);
            mx = __reduce_max_sync(0xffffffffu, mx);
            valid = __reduce_add_sync(0xffffffffu, valid);
            if (valid < kc) return KEY_MAX;  // slots reserved but not written yet count as +inf
            const uint32_t diff = mn ^ mx;
            if (diff == 0u) return mn;
            const int top = 31 - __clz(diff);
            uint32_t lo = mn & ~((2u << top) - 1u);  // the bits every valid key shares
#pragma unroll 1
            for (int bit = top; bit >= 0; --bit) {
                const uint32_t cand = lo | (1u << bit);
                int c = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) c += key[j] < cand ? 1 : 0;
                if (__reduce_add_sync(0xffffffffu, c) < kc) lo = cand;  // the kc-th smallest is >= cand
            }
            return lo;
        };
        for (;;) {
            // up to two pending requests: lane l looks at mailbox l; round-robin start so that no box starves
            const unsigned long long mine = mbox[lane];
            const unsigned pend = __ballot_sync(0xffffffffu, mine != 0ull);
            int ia = -1, ib = -1;
            unsigned long long ra = 0ull, rb = 0ull;
            if (pend) {
                const unsigned rot = __funnelshift_r(pend, pend, rr);  // bit j of rot = box (j + rr) & 31
                ia = (__ffs(rot) - 1 + rr) & 31;
                const unsigned rot2 = rot & (rot - 1);
                if (rot2) ib = (__ffs(rot2) - 1 + rr) & 31;
                ra = __shfl_sync(0xffffffffu, mine, ia);
                if (ib >= 0) rb = __shfl_sync(0xffffffffu, mine, ib);
                rr = (ia + 1) & 31;
            }
            if (ia < 0) {
                if (*ep_done >= 8u) break;
                __nanosleep(100);
                continue;
            }
            int qa, na, qb2 = 0, nb = 0;
            const unsigned long long* pa = decode(ra, qa, na);
            if (na > 256) {
                // large window: shared-memory radix select
                for (int base = 0; base < na; base += 256) {
                    unsigned long long e[8];
                    load8(pa + base, na - base, e);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int idx = base + j * 32 + lane;
                        if (idx < na) keys[idx] = (uint32_t)(e[j] >> 32);
                    }
                }
                __syncwarp();
                const uint32_t* kk = keys;
                const uint32_t t = radix_select([kk](int i2) { return kk[i2]; }, na, kc, hist, lane);
                if (t < KEY_MAX && lane == 0) atomicMin(a.gthr + qa, t);
                __syncwarp();
                if (lane == 0) mbox[ia] = 0ull;
                ++served;
                continue;
            }
            unsigned long long ea[8], eb[8];
            load8(pa, na, ea);
            if (ib >= 0) {
                const unsigned long long* pb = decode(rb, qb2, nb);
                load8(pb, nb, eb);
            }
            {
                const uint32_t t = select8(ea);
                if (t < KEY_MAX && lane == 0) atomicMin(a.gthr + qa, t);
                if (lane == 0) mbox[ia] = 0ull;
                ++served;
            }
            if (ib >= 0) {
                const uint32_t t = select8(eb);
                if (t < KEY_MAX && lane == 0) atomicMin(a.gthr + qb2, t);
                if (lane == 0) mbox[ib] = 0ull;
                ++served;
            }
        }
        if (lane == 0 && served) atomicAdd(&a.ctrl[8], served);  // statistics: refreshes served
    }
    // ---- teardown: all tensor-memory traffic of this CTA has completed once every role has left its loop
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(MMA_TMEM_COLS));
}

}  // namespace qk
