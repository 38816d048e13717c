// Shared device helpers for the quake_b200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "../../include/quake_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "quake_b200 kernels are written for sm_100a (B200) only"
#endif

namespace qk {

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define QK_CUDA(call)                                              \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return ::qk::cuda_fail(_e, #call);  \
    } while (0)
#define QK_REQUIRE(cond, ...)                     \
    do {                                          \
        if (!(cond)) {                            \
            ::qk::set_error(__VA_ARGS__);         \
            return QK_ERR_INVALID_ARGUMENT;       \
        }                                         \
    } while (0)

int sm_count();

// every kernel launch of the library is followed by QK_LAUNCHED(): launch-error check + the launch counter that
// qk_launch_count() reports (bench.py's gpu_launches)
void count_launch();
#define QK_LAUNCHED()                          \
    do {                                       \
        ::qk::count_launch();                  \
        QK_CUDA(cudaGetLastError());           \
    } while (0)

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may start -- block scheduling, prologue --
// while the previous kernel of its stream is still draining; it must call pdl_wait() before it touches anything that
// kernel (or any earlier one) wrote. A primary that calls pdl_launch_dependents() early lets the dependent's CTAs become
// resident as soon as every primary CTA has got that far. Inside a stream capture these become programmatic graph edges.
// Both device calls are no-ops in a kernel launched the ordinary way. QK_PDL is a mask of the launch sites that use it
// (1: grouping kernels after the coarse refine, 2: refine kernel after the partition scan; default 0 = none: no
// measurable gain at C2, kept as an experiment switch).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
int pdl_mask();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int site, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl_mask() & site) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// qk_scan_partitions with more knobs (library-internal callers: k-means assign, qk_search_ivf)
struct ScanExtras {
    int rank_squared = 0;                    // l2: order results by the squared distance (k-means assign: faiss's Top1
                                             // handler compares squared distances)
    const float* max_row_norm_dev = nullptr; // the store's norm bound lives on the device (overrides store->max_row_norm)
    const int64_t* probe_ids = nullptr;      // probes given as partition ids [Q x nprobe] (probe_lists == NULL) ...
    const int32_t* id_to_slot = nullptr;     // ... mapped through this dense table
    int64_t table_size = 0;
    int shard_rank = 0, shard_world = 1;     // shard_world > 1: scan only the partitions with id % world == rank
    int set_mode = 0;                        // dense mode only: the k best as a set (refine.cuh: set_select_and_emit)
    // collect mode (APS rounds, qk_scan_collect): the per-query filter thresholds are GIVEN (device keys, [Q]) and
    // stay fixed -- every row at or below them is kept -- and instead of a top-k the survivors are refined exactly and
    // handed out grouped by the probe rank of their list
    const uint32_t* preset_thresholds = nullptr;
    int64_t* collect_ids = nullptr;          // [Q x nprobe x k]
    float* collect_dist = nullptr;           // [Q x nprobe x k]
    int32_t* collect_cnt = nullptr;          // [Q x nprobe] entries per (query, rank), <= k, best first
    int32_t* collect_overflow = nullptr;     // [Q] 1 = the candidate buffer overflowed: this query's lists are unusable
    // qk_search_ivf plumbing: the coarse scan's refine kernel expands its results straight into the partition scan's
    // pair table (fused_expand, on the coarse call), the partition scan then skips its own expansion
    // (pairs_preexpanded) and the workspace clears the caller already issued on a parallel branch (ws_precleared;
    // pre_refine_event = a cudaEvent_t the coarse call's stream waits for before its refine kernel touches that table)
    const struct FusedExpand* fused_expand = nullptr;
    void* pre_refine_event = nullptr;
    int pairs_preexpanded = 0;
    int ws_precleared = 0;
};
struct FusedExpand {
    int32_t* pair_seg;          // [Q x nprobe] of the NEXT scan
    int32_t* seg_count;         // [S]
    uint32_t* gthr;             // [Q]
    const int32_t* id_to_slot;
    int64_t table_size;
    const int32_t* list_seg0;
    const int32_t* list_nseg;
    int num_lists, shard_rank, shard_world;
};
// probe_lists == NULL and probe_ids == NULL: flat mode -- the store has ONE list and every query scans it.
int scan_partitions_impl(const qk_store_t* st, const float* queries, int64_t Q, int64_t q_pitch,
                         const int32_t* probe_lists, int nprobe, int metric, int k, int64_t* out_ids, float* out_dist,
                         int64_t* out_rows, void* workspace, size_t workspace_bytes, int32_t* stats, void* stream,
                         const ScanExtras& extras);

// ---------------------------------------------------------------------------------------------
// order-preserving float <-> uint32 keys (ascending float order == ascending unsigned order)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f2key(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    // NaN (either sign) never wins a comparison in the reference (faiss result handlers use strict
    // comparisons; a NaN centroid comes out of kmeans_refine_partitions for an emptied cluster,
    // clustering.cpp:122-124): it maps to the "invalid" key.
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float key2f(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
static constexpr uint32_t KEY_MAX = 0xffffffffu;
static constexpr uint64_t COMP_MAX = 0xffffffffffffffffull;

// ---------------------------------------------------------------------------------------------
// Pairwise distances in the REFERENCE'S summation order.
//
// The reference's per-pair arithmetic is faiss fvec_L2sqr / fvec_inner_product
// (third_party/faiss/faiss/utils/distances_simd.cpp:188-224), plain loops that GCC -O3 (the reference's
// Release flag, CMakeLists.txt:37) vectorises 8-wide for AVX2 without contracting the main loop to
// FMA: lane j accumulates elements i == j (mod 8) with separately rounded sub, mul, add; the lanes
// are folded as ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)); a 4-wide FMA step and scalar FMA steps
// absorb the remainder when d is not a multiple of 8. The functions below reproduce exactly that
// evaluation order with IEEE round-to-nearest intrinsics, so the refined distances are
// bit-identical to the reference build in oracle/_ref (checked in tests/).
// ---------------------------------------------------------------------------------------------
template <bool kIP>
__device__ __forceinline__ float ref_term(float x, float y) {
    if (kIP) return __fmul_rn(x, y);
    float t = __fsub_rn(x, y);
    return __fmul_rn(t, t);
}
template <bool kIP>
__device__ __forceinline__ float ref_fma_term(float x, float y, float acc) {
    if (kIP) return __fmaf_rn(x, y, acc);
    float t = __fsub_rn(x, y);
    return __fmaf_rn(t, t, acc);
}

template <bool kIP>
__device__ float ref_pair_distance(const float* __restrict__ x, const float* __restrict__ y, int d) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f;
    const int nb = d >> 3;
    for (int b = 0; b < nb; ++b) {
        const float* xb = x + 8 * b;
        const float* yb = y + 8 * b;
        a0 = __fadd_rn(a0, ref_term<kIP>(xb[0], yb[0]));
        a1 = __fadd_rn(a1, ref_term<kIP>(xb[1], yb[1]));
        a2 = __fadd_rn(a2, ref_term<kIP>(xb[2], yb[2]));
        a3 = __fadd_rn(a3, ref_term<kIP>(xb[3], yb[3]));
        a4 = __fadd_rn(a4, ref_term<kIP>(xb[4], yb[4]));
        a5 = __fadd_rn(a5, ref_term<kIP>(xb[5], yb[5]));
        a6 = __fadd_rn(a6, ref_term<kIP>(xb[6], yb[6]));
        a7 = __fadd_rn(a7, ref_term<kIP>(xb[7], yb[7]));
    }
    float s0 = __fadd_rn(a0, a4), s1 = __fadd_rn(a1, a5), s2 = __fadd_rn(a2, a6), s3 = __fadd_rn(a3, a7);
    float res = __fadd_rn(__fadd_rn(s0, s2), __fadd_rn(s1, s3));
    int o = nb << 3;
    int r = d - o;
    if (r >= 4) {
        float f0 = ref_fma_term<kIP>(x[o + 0], y[o + 0], s0);
        float f1 = ref_fma_term<kIP>(x[o + 1], y[o + 1], s1);
        float f2 = ref_fma_term<kIP>(x[o + 2], y[o + 2], s2);
        float f3 = ref_fma_term<kIP>(x[o + 3], y[o + 3], s3);
        res = __fadd_rn(__fadd_rn(f0, f2), __fadd_rn(f1, f3));
        o += 4;
        r -= 4;
    }
    for (int i = 0; i < r; ++i) res = ref_fma_term<kIP>(x[o + i], y[o + i], res);
    return res;
}

// The same evaluation spread over an aligned group of 8 lanes: lane j (= lane & 7) carries accumulator
// a_j; the fold uses shuffles inside the group. Every lane of the group must call it; the result is valid
// on the group's lane 0. Bit-identical to ref_pair_distance (float addition is commutative).
template <bool kIP>
__device__ __forceinline__ float ref_pair_distance_g8(const float* __restrict__ x, const float* __restrict__ y, int d,
                                                      int j) {
    const unsigned gmask = 0xffu << ((threadIdx.x & 31u) & 24u);  // the group may sit in a divergent branch
    float a = 0.f;
    const int nb = d >> 3;
    for (int b = 0; b < nb; ++b) a = __fadd_rn(a, ref_term<kIP>(x[8 * b + j], y[8 * b + j]));
    // s_j = a_j + a_{j+4} on lanes j < 4
    float s = __fadd_rn(a, __shfl_down_sync(gmask, a, 4, 8));
    int o = nb << 3;
    int r = d - o;
    float f = s;
    if (r >= 4) {
        if (j < 4) f = ref_fma_term<kIP>(x[o + j], y[o + j], s);
        o += 4;
        r -= 4;
    }
    // (f0 + f2) on lane 0, (f1 + f3) on lane 1, then their sum on lane 0
    float t = __fadd_rn(f, __shfl_down_sync(gmask, f, 2, 8));
    float res = __fadd_rn(t, __shfl_down_sync(gmask, t, 1, 8));
    for (int i = 0; i < r; ++i) res = ref_fma_term<kIP>(x[o + i], y[o + i], res);
    return res;
}

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk-copy (TMA, UBLKCP) wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Waits until the phase with the given parity has completed. try_wait carries a suspend-time hint, so a waiting warp
// is parked by the hardware instead of spinning through the issue slots of the warps that have work (16 warps share
// 4 schedulers here, and most of them are waiting at any time). A wait that lasts longer than any legitimate one
// (seconds) means a pipeline protocol error: report where and trap instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t a = smem_u32(bar);
    uint32_t spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity), "r"(1000000u)  // parked for up to 1 ms per attempt (a 20 us limit cost 14 us at C2:
                                                   // the parked warps of a 16-warp CTA wake up and spin)
            : "memory");
        if (!done && ++spins == 8000u) {  // seconds: a pipeline protocol error, not a slow neighbour
            printf("quake_b200: mbarrier wait timed out: block %d warp %d lane %d barrier@%u parity %u\n", blockIdx.x,
                   threadIdx.x >> 5, threadIdx.x & 31, a & 0xffffu, parity);
            __trap();
        }
    }
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, 16B aligned).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 2-D tiled TMA load (tensor map in kernel parameter space), completion on an mbarrier.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
// 16-byte asynchronous copy global -> shared that bypasses L1 (LDGSTS)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// one (pre-counted) arrival on `bar` once all cp.async of this thread issued so far have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

}  // namespace qk
