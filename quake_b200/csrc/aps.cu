// Adaptive Partition Scanning (APS) on the device.
//
// Replaces the per-query APS logic of QueryCoordinator::serial_scan
// (/root/reference/src/cpp/src/query_coordinator.cpp:521-579) and the geometry it calls
// (src/cpp/include/geometry.h: compute_boundary_distances :57-113, incomplete_beta :115-161,
// incomplete_beta_lookup :163-211, log_hyperspherical_cap_volume :247-295, compute_recall_profile :345-407).
//
// The reference scans a query's candidate partitions one at a time and, after each, re-estimates the recall
// from the current k-th distance. Here the partition scans of a ROUND (R consecutive probe ranks of every still
// active query) run as one batched qk_scan_partitions call that keeps the per-(query, rank) top-k apart, and
// aps_advance_kernel then replays the reference's sequential loop for every query over those R ranks: merge the
// rank's top-k into the running top-k, take the k-th distance, recompute the recall profile when the radius
// moved by more than recompute_threshold, stop when the estimate reaches the target. Whatever was scanned past
// a query's stopping point is discarded, so results and partitions_scanned are those of the sequential loop.
#include "common.cuh"
#include <cmath>
#include <cfloat>

namespace qk {

static constexpr int APS_TABLE = 1001;  // NUM_X_VALUES, geometry.h:26
static constexpr int APS_THREADS = 128;

// geometry.h:115-161 -- regularised incomplete beta function, Lentz continued fraction (double)
__host__ __device__ inline double aps_incomplete_beta(double a, double b, double x) {
    if (x < 0.0 || x > 1.0) return INFINITY;
    bool flip = false;
    if (x > (a + 1.0) / (a + b + 2.0)) {  // I_x(a,b) = 1 - I_{1-x}(b,a); one level, as the recursion in the reference
        const double t = a; a = b; b = t;
        x = 1.0 - x;
        flip = true;
    }
    const double lbeta_ab = lgamma(a) + lgamma(b) - lgamma(a + b);
    const double front = exp(log(x) * a + log(1.0 - x) * b - lbeta_ab) / a;
    double f = 1.0, c = 1.0, dd = 0.0, res = INFINITY;
    for (int i = 0; i <= 200; ++i) {
        const int m = i / 2;
        double numerator;
        if (i == 0) numerator = 1.0;
        else if (i % 2 == 0) numerator = (m * (b - m) * x) / ((a + 2.0 * m - 1.0) * (a + 2.0 * m));
        else numerator = -((a + m) * (a + b + m) * x) / ((a + 2.0 * m) * (a + 2.0 * m + 1));
        dd = 1.0 + numerator * dd;
        if (fabs(dd) < 1.0e-30) dd = 1.0e-30;
        dd = 1.0 / dd;
        c = 1.0 + numerator / c;
        if (fabs(c) < 1.0e-30) c = 1.0e-30;
        const double cd = c * dd;
        f *= cd;
        if (fabs(1.0 - cd) < 1.0e-8) { res = front * (f - 1.0); break; }
    }
    return flip ? 1.0 - res : res;
}

// geometry.h:188-211 -- linear interpolation in the precomputed table
__device__ inline double aps_beta_lookup(const double* __restrict__ table, double x) {
    x = fmax(0.0, fmin(1.0, x));
    const double scaled = x * (APS_TABLE - 1);
    int xi = (int)scaled;
    if (xi < 0) xi = 0;
    if (xi > APS_TABLE - 2) xi = APS_TABLE - 2;
    const double y1 = table[xi], y2 = table[xi + 1];
    const double dx = 1.0 / (APS_TABLE - 1);
    const double x1 = xi * dx;
    return y1 + (x - x1) * (y2 - y1) / dx;
}

// geometry.h:247-295 with ratio = true
__device__ inline double aps_log_cap_volume(double radius, double boundary, int d, bool use_precomputed, bool euclid,
                                            const double* __restrict__ table) {
    double h = radius - boundary;
    h = fmax(0.0, fmin(2 * radius, h));
    if (euclid) {
        const double x = sqrt((2 * radius * h - h * h) / (radius * radius));
        const double ib = (use_precomputed && table) ? aps_beta_lookup(table, x) : aps_incomplete_beta((d + 1.0) / 2.0, 0.5, x);
        if (ib <= 0.0 || isnan(ib) || isinf(ib)) return -INFINITY;
        return log(0.5) + log(ib);
    }
    const double s1 = sin(radius / 2.0), s2 = sin(boundary / 2.0);
    const double l1 = log(aps_incomplete_beta((d - 1) / 2.0, 0.5, s1 * s1));
    const double l2 = log(aps_incomplete_beta((d - 1) / 2.0, 0.5, s2 * s2));
    return log(0.5) + l1 - l2;
}

// ------------------------------------------------------------------------------------------------
// boundary distances (geometry.h:57-113): eight lanes per (query, candidate) pair, every inner product in the
// summation order of faiss fvec_inner_product (see common.cuh)
// ------------------------------------------------------------------------------------------------
// Inner product of two element streams in the reference's order; x(i), y(i) are evaluated on the fly.
template <typename FX, typename FY>
__device__ __forceinline__ float ref_ip_g8(FX x, FY y, int d, int j) {
    const unsigned gmask = 0xffu << ((threadIdx.x & 31u) & 24u);
    float a = 0.f;
    const int nb = d >> 3;
    for (int b = 0; b < nb; ++b) a = __fadd_rn(a, __fmul_rn(x(8 * b + j), y(8 * b + j)));
    float s = __fadd_rn(a, __shfl_down_sync(gmask, a, 4, 8));
    int o = nb << 3;
    int r = d - o;
    float f = s;
    if (r >= 4) {
        if (j < 4) f = __fmaf_rn(x(o + j), y(o + j), s);
        o += 4;
        r -= 4;
    }
    float t = __fadd_rn(f, __shfl_down_sync(gmask, f, 2, 8));
    float res = __fadd_rn(t, __shfl_down_sync(gmask, t, 1, 8));
    for (int i = 0; i < r; ++i) res = __fmaf_rn(x(o + i), y(o + i), res);
    return __shfl_sync(gmask, res, (threadIdx.x & 24u), 32);  // broadcast lane 0 of the group
}

__global__ void __launch_bounds__(256) aps_boundary_kernel(const float* __restrict__ queries, int64_t Q, int64_t q_pitch, int d,
                                                           const float* __restrict__ cents, int64_t c_pitch,
                                                           const int64_t* __restrict__ cand_rows, int m, int euclid,
                                                           float* __restrict__ out) {
    const int64_t pair = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3);
    const int j8 = threadIdx.x & 7;
    if (pair >= Q * m) return;  // uniform inside an 8-lane group
    const int64_t q = pair / m;
    const int j = (int)(pair - q * m);
    if (j == 0) {
        if (j8 == 0) out[pair] = -1.0f;
        return;
    }
    const int64_t r0 = cand_rows[q * m], rj = cand_rows[pair];
    if (r0 < 0 || rj < 0) {  // the coarse scan returned fewer candidates than asked for
        if (j8 == 0) out[pair] = -1.0f;
        return;
    }
    const float* qv = queries + q * q_pitch;
    const float* c0 = cents + r0 * c_pitch;
    const float* cj = cents + rj * c_pitch;
    float res;
    if (euclid) {
        auto line = [&](int i) { return __fsub_rn(cj[i], c0[i]); };
        auto resid = [&](int i) { return __fsub_rn(qv[i], c0[i]); };
        const float A2 = ref_ip_g8(line, line, d, j8);
        const float A = __fsqrt_rn(A2);
        const float dot = ref_ip_g8(resid, line, d, j8);
        res = __fdiv_rn(fabsf(__fsub_rn(dot, __fmul_rn(0.5f, A2))), A);
    } else {
        auto mid = [&](int i) { return __fadd_rn(c0[i], __fdiv_rn(__fsub_rn(cj[i], c0[i]), 2.0f)); };
        const float nrm = __fsqrt_rn(ref_ip_g8(mid, mid, d, j8));
        auto midn = [&](int i) { return __fdiv_rn(mid(i), nrm); };
        auto qf = [&](int i) { return qv[i]; };
        const float ang = ref_ip_g8(qf, midn, d, j8);
        res = (float)acos((double)ang);
    }
    if (j8 == 0) out[pair] = res;
}

// ------------------------------------------------------------------------------------------------
// one round of the APS loop
// ------------------------------------------------------------------------------------------------
struct ApsArgs {
    const int32_t* active;  // [num_active] query indices
    int R, p0, m, k, d, ip;
    const int32_t* slots;       // [Q x m] list slot of every candidate (-1: skip, query_coordinator.cpp:540)
    const int64_t* round_ids;   // [num_active x R x k]
    const float* round_dist;
    const int32_t* round_cnt;   // optional [num_active x R]: valid entries of every list (else: padded with id -1)
    const float* boundary;      // [Q x m]
    const double* table;        // device [APS_TABLE] or null
    float recall_target, recompute_threshold;
    int use_precomputed;
    int64_t* run_ids;           // [Q x k]
    float* run_dist;
    int32_t* run_cnt;           // [Q]
    float* radius;              // [Q]
    int32_t* have_probs;        // [Q]
    float* probs;               // [Q x m]
    int32_t* done;              // [Q]
    int32_t* scanned;           // [Q]
    int32_t* still_active;      // device counter
};

__device__ __forceinline__ bool aps_less(uint32_t ka, int64_t ia, uint32_t kb, int64_t ib) {
    return ka != kb ? ka < kb : ia < ib;
}

__global__ void __launch_bounds__(APS_THREADS) aps_advance_kernel(const ApsArgs a) {
    extern __shared__ __align__(16) unsigned char aps_sm[];
    const int k = a.k, m = a.m;
    int64_t* idA = reinterpret_cast<int64_t*>(aps_sm);          // running top-k
    int64_t* idB = idA + k;                                     // the rank's top-k
    int64_t* idO = idB + k;                                     // merged
    uint32_t* kA = reinterpret_cast<uint32_t*>(idO + k);
    uint32_t* kB = kA + k;
    uint32_t* kO = kB + k;
    __shared__ int s_cnt, s_stop, s_recompute;
    const int tid = threadIdx.x;
    const int64_t slot_a = blockIdx.x;
    const int64_t q = a.active[slot_a];
    if (a.done[q]) return;
    int cnt = a.run_cnt[q];
    for (int i = tid; i < cnt; i += blockDim.x) {
        const float dv = a.run_dist[q * k + i];
        kA[i] = f2key(a.ip ? -dv : dv);
        idA[i] = a.run_ids[q * k + i];
    }
    float radius = a.radius[q];
    int have = a.have_probs[q];
    int scanned = a.scanned[q];
    float* probs = a.probs + q * m;
    const float* bnd = a.boundary + q * m;
    bool stop = false;
    __syncthreads();
    for (int r = 0; r < a.R && !stop; ++r) {
        const int p = a.p0 + r;
        if (p >= m) break;
        if (a.slots[q * m + p] < 0) continue;  // invalid partition: skipped entirely (query_coordinator.cpp:540-542)
        // ---- the rank's top-k (padded entries have id -1)
        const int64_t* rid = a.round_ids + ((size_t)slot_a * a.R + r) * k;
        const float* rdv = a.round_dist + ((size_t)slot_a * a.R + r) * k;
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        if (a.round_cnt) {
            const int c = a.round_cnt[(size_t)slot_a * a.R + r];
            for (int i = tid; i < c; i += blockDim.x) {
                idB[i] = rid[i];
                kB[i] = f2key(a.ip ? -rdv[i] : rdv[i]);
            }
            if (tid == 0) s_cnt = c;
        } else {
            int local = 0;
            for (int i = tid; i < k; i += blockDim.x) {
                const int64_t id = rid[i];
                idB[i] = id;
                kB[i] = f2key(a.ip ? -rdv[i] : rdv[i]);
                if (id >= 0) ++local;
            }
            if (local) atomicAdd(&s_cnt, local);
        }
        __syncthreads();
        const int cntB = s_cnt;  // valid entries are a prefix (best first, padding last)
        // ---- merge by rank: position of x in the union = own index + number of smaller elements of the other list
        const int newcnt = min(k, cnt + cntB);
        for (int i = tid; i < cnt + cntB; i += blockDim.x) {
            const bool fromA = i < cnt;
            const int own = fromA ? i : i - cnt;
            const uint32_t key = fromA ? kA[own] : kB[own];
            const int64_t id = fromA ? idA[own] : idB[own];
            const uint32_t* ok = fromA ? kB : kA;
            const int64_t* oid = fromA ? idB : idA;
            int lo = 0, hi = fromA ? cntB : cnt;
            while (lo < hi) {  // first index of the other list that is NOT smaller than (key, id)
                const int mid = (lo + hi) >> 1;
                if (aps_less(ok[mid], oid[mid], key, id)) lo = mid + 1;
                else hi = mid;
            }
            const int pos = own + lo;
            if (pos < newcnt) { kO[pos] = key; idO[pos] = id; }
        }
        __syncthreads();
        for (int i = tid; i < newcnt; i += blockDim.x) { kA[i] = kO[i]; idA[i] = idO[i]; }
        cnt = newcnt;
        ++scanned;
        __syncthreads();
        // ---- the reference's APS step (query_coordinator.cpp:557-579)
        float curr;
        if (cnt >= k) {
            const float v = key2f(kA[k - 1]);
            curr = a.ip ? -v : v;
        } else {
            curr = a.ip ? -INFINITY : FLT_MAX;  // TopkBuffer sentinel (list_scanning.h:187-191)
        }
        if (tid == 0) {
            const float change = __fdiv_rn(fabsf(__fsub_rn(curr, radius)), curr);
            s_recompute = (change > a.recompute_threshold) ? 1 : 0;
        }
        __syncthreads();
        if (s_recompute) {
            radius = curr;
            // compute_recall_profile (geometry.h:345-407)
            for (int j = 1 + tid; j < m; j += blockDim.x) {
                const float b = bnd[j];
                float pj = 0.0f;
                if (!(b >= radius)) {
                    const double v = exp(aps_log_cap_volume((double)radius, (double)b, a.d, a.use_precomputed != 0, !a.ip, a.table));
                    pj = (float)((v > 0.0) ? v : 0.0);
                }
                probs[j] = pj;
            }
            __syncthreads();
            if (tid == 0) {
                probs[0] = (float)(2.0 * (double)probs[1]);
                double sum = 0.0;
                for (int j = 0; j < m; ++j) sum += (double)probs[j];
                if (sum > 0.0) {
                    for (int j = 0; j < m; ++j) probs[j] = (float)((double)probs[j] / sum);
                } else {
                    for (int j = 0; j < m; ++j) probs[j] = (float)(1.0 / m);
                }
            }
            have = 1;
            __syncthreads();
        }
        if (tid == 0) {
            float est = 0.0f;
            if (have)
                for (int i = 0; i < p; ++i) est = __fadd_rn(est, probs[i]);
            s_stop = (est >= a.recall_target) ? 1 : 0;
        }
        __syncthreads();
        stop = s_stop != 0;
        __syncthreads();
    }
    const bool finished = stop || (a.p0 + a.R >= m);
    // ---- write the state back
    for (int i = tid; i < k; i += blockDim.x) {
        float dv = a.ip ? -INFINITY : INFINITY;
        int64_t id = -1;
        if (i < cnt) {
            const float v = key2f(kA[i]);
            dv = a.ip ? -v : v;
            id = idA[i];
        }
        a.run_dist[q * k + i] = dv;
        a.run_ids[q * k + i] = id;
    }
    if (tid == 0) {
        a.run_cnt[q] = cnt;
        a.radius[q] = radius;
        a.have_probs[q] = have;
        a.scanned[q] = scanned;
        if (finished) a.done[q] = 1;
        else atomicAdd(a.still_active, 1);
    }
}

// Per active query: the filter-key threshold below which EVERY row whose exact distance can still enter the running
// top-k must fall (collect mode of the scan). With r = the current k-th distance (the reference's
// TopkBuffer::get_kth_distance, +-inf while fewer than k results are held) and s the filter score:
//   l2: exact d^2 >= (|q|^2 (1 - g) + s - e1)(1 - e2)   =>   d <= r implies s <= r^2 / (1 - e2) - |q|^2 (1 - g) + e1
//   ip: |s + <q,v>| <= err                               =>   <q,v> >= r implies s <= -r + err
// (the same error terms as the refine step's proof, csrc/refine.cuh), rounded up generously.
__global__ void aps_thresholds_kernel(const int32_t* __restrict__ active, int64_t num_active, const float* __restrict__ queries,
                                      int64_t q_pitch, int d, const float* __restrict__ run_dist, const int32_t* __restrict__ run_cnt,
                                      int k, int ip, float max_row_norm, double filter_gam, uint32_t* __restrict__ out_keys) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= num_active) return;
    const int64_t q = active[w];
    double qn = 0.0;
    for (int i = lane; i < d; i += 32) {
        const double x = queries[q * q_pitch + i];
        qn += x * x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qn += __shfl_xor_sync(0xffffffffu, qn, o);
    if (lane) return;
    uint32_t key = f2key(INFINITY);  // everything with a valid score
    if (run_cnt[q] >= k) {
        const double r = (double)run_dist[q * k + k - 1];
        const double eps = 5.960464477539063e-08, U = (double)max_row_norm, qnorm = sqrt(qn);
        const double gam = (d + 8) * eps, e2 = (d / 8 + 12) * eps;
        double t;
        if (!ip) {
            const double base = r * r / (1.0 - e2) - qn * (1.0 - gam) + gam * U * U + 2.0 * (gam + filter_gam) * qnorm * U;
            t = base + 16.0 * eps * (fabs(base) + r * r + qn);
        } else {
            const double err = (gam + filter_gam + e2) * qnorm * U;
            t = -r + err + 16.0 * eps * (fabs(r) + qnorm * U);
        }
        float tf = __double2float_ru(t);
        tf = nextafterf(tf, INFINITY);
        if (tf == tf) key = f2key(tf);
    }
    out_keys[w] = key;
}

}  // namespace qk

using namespace qk;

extern "C" int qk_aps_thresholds(const int32_t* active, int64_t num_active, const float* queries, int64_t q_pitch, int d,
                                 const float* run_distances, const int32_t* run_count, int k, int metric,
                                 float max_row_norm, int filter_terms, uint32_t* out_keys, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(active && queries && run_distances && run_count && out_keys && num_active > 0 && k > 0, "bad argument");
    const double fg = filter_terms == 2 ? 9.765625e-04 + 7.62939453125e-06 : 7.62939453125e-06;
    aps_thresholds_kernel<<<(unsigned)((num_active * 32 + 255) / 256), 256, 0, stream>>>(
        active, num_active, queries, q_pitch, d, run_distances, run_count, k, metric == QK_METRIC_INNER_PRODUCT ? 1 : 0,
        max_row_norm, fg, out_keys);
    QK_LAUNCHED();
    return QK_OK;
}

extern "C" int qk_host_beta_table(int d, double* table) {
    QK_REQUIRE(d > 0 && table, "qk_host_beta_table: bad argument");
    const double dx = 1.0 / (APS_TABLE - 1);
    const double a = (d + 1.0) / 2.0, b = 0.5;
    for (int i = 0; i < APS_TABLE; ++i) table[i] = aps_incomplete_beta(a, b, i * dx);
    return QK_OK;
}

extern "C" int qk_aps_boundary_distances(const float* queries, int64_t Q, int64_t q_pitch, int d, const float* centroids,
                                         int64_t centroid_pitch, const int64_t* cand_rows, int m, int metric,
                                         float* out_boundary, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(queries && centroids && cand_rows && out_boundary && Q > 0 && m > 0 && d > 0, "bad argument");
    QK_REQUIRE(metric == QK_METRIC_L2 || metric == QK_METRIC_INNER_PRODUCT, "metric %d not supported", metric);
    const int64_t pairs = Q * m;
    aps_boundary_kernel<<<(unsigned)((pairs + 31) / 32), 256, 0, stream>>>(queries, Q, q_pitch, d, centroids, centroid_pitch,
                                                                          cand_rows, m, metric == QK_METRIC_L2 ? 1 : 0,
                                                                          out_boundary);
    QK_LAUNCHED();
    return QK_OK;
}

extern "C" int qk_aps_advance(const int32_t* active, int64_t num_active, int R, int p0, int m, int k, int d, int metric,
                              const int32_t* slots, const int64_t* round_ids, const float* round_distances, const int32_t* round_cnt,
                              const float* boundary, const double* beta_table, float recall_target,
                              float recompute_threshold, int use_precomputed, int64_t* run_ids, float* run_distances,
                              int32_t* run_count, float* radius, int32_t* have_probs, float* probs, int32_t* done,
                              int32_t* scanned, int32_t* still_active, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    QK_REQUIRE(active && slots && round_ids && round_distances && boundary && run_ids && run_distances && run_count &&
                   radius && have_probs && probs && done && scanned && still_active,
               "null argument");
    QK_REQUIRE(num_active > 0 && R > 0 && m >= 2 && k >= 1 && k <= 1024, "qk_aps_advance: bad sizes (k <= 1024, m >= 2)");
    ApsArgs a;
    a.active = active; a.R = R; a.p0 = p0; a.m = m; a.k = k; a.d = d; a.ip = metric == QK_METRIC_INNER_PRODUCT;
    a.slots = slots; a.round_ids = round_ids; a.round_dist = round_distances; a.round_cnt = round_cnt; a.boundary = boundary;
    a.table = beta_table;
    a.recall_target = recall_target; a.recompute_threshold = recompute_threshold; a.use_precomputed = use_precomputed;
    a.run_ids = run_ids; a.run_dist = run_distances; a.run_cnt = run_count; a.radius = radius; a.have_probs = have_probs;
    a.probs = probs; a.done = done; a.scanned = scanned; a.still_active = still_active;
    const size_t smem = (size_t)3 * k * (sizeof(int64_t) + sizeof(uint32_t));
    QK_CUDA(cudaMemsetAsync(still_active, 0, sizeof(int32_t), stream));
    aps_advance_kernel<<<(unsigned)num_active, APS_THREADS, smem, stream>>>(a);
    QK_LAUNCHED();
    return QK_OK;
}
