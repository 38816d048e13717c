/*
 * quake_oracle.c -- TEST INFRASTRUCTURE ONLY. CPU restatement of the reference's search hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this
 * library, and only as the checker / the reported baseline; nothing under quake_b200/ links, imports or
 * calls it.
 *
 * Parity pinning: every function here is checked against the compiled, unmodified reference
 * (oracle/_ref, built by oracle/build_ref.sh from /root/reference) and against golden vectors generated
 * from it (tests/golden/, generator tests/golden/make_golden.py); see tests/test_oracle.py.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------------------------------------
 * Pairwise kernels: faiss fvec_L2sqr / fvec_inner_product
 * (src/cpp/third_party/faiss/faiss/utils/distances_simd.cpp:188-224). The source is a plain
 * `res += tmp*tmp` loop under `#pragma GCC optimize("unroll-loops,associative-math,no-signed-zeros")`,
 * i.e. the summation order is whatever the compiler picks. The reference's Release flags (-O3,
 * CMakeLists.txt:37) make GCC vectorise it 8-wide: lane j accumulates elements i == j (mod 8) with
 * separately rounded sub / mul / add; the lanes fold as ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)); a 4-wide
 * FMA step and scalar FMA steps absorb d mod 8. This restatement spells that order out (compiled with
 * -ffp-contract=off so nothing else is fused), which makes it bit-identical to oracle/_ref.
 * --------------------------------------------------------------------------------------------- */
static inline float term_l2(float x, float y) { float t = x - y; return t * t; }
static inline float term_ip(float x, float y) { return x * y; }

ORC_API float orc_l2sqr(const float* x, const float* y, size_t d) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t nb = d / 8;
    for (size_t b = 0; b < nb; b++)
        for (int j = 0; j < 8; j++) a[j] = a[j] + term_l2(x[8 * b + j], y[8 * b + j]);
    float s0 = a[0] + a[4], s1 = a[1] + a[5], s2 = a[2] + a[6], s3 = a[3] + a[7];
    float res = (s0 + s2) + (s1 + s3);
    size_t o = nb * 8, r = d - o;
    if (r >= 4) {
        float t0 = x[o] - y[o], t1 = x[o + 1] - y[o + 1], t2 = x[o + 2] - y[o + 2], t3 = x[o + 3] - y[o + 3];
        float f0 = fmaf(t0, t0, s0), f1 = fmaf(t1, t1, s1), f2 = fmaf(t2, t2, s2), f3 = fmaf(t3, t3, s3);
        res = (f0 + f2) + (f1 + f3);
        o += 4;
        r -= 4;
    }
    for (size_t i = 0; i < r; i++) {
        float t = x[o + i] - y[o + i];
        res = fmaf(t, t, res);
    }
    return res;
}

ORC_API float orc_ip(const float* x, const float* y, size_t d) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t nb = d / 8;
    for (size_t b = 0; b < nb; b++)
        for (int j = 0; j < 8; j++) a[j] = a[j] + term_ip(x[8 * b + j], y[8 * b + j]);
    float s0 = a[0] + a[4], s1 = a[1] + a[5], s2 = a[2] + a[6], s3 = a[3] + a[7];
    float res = (s0 + s2) + (s1 + s3);
    size_t o = nb * 8, r = d - o;
    if (r >= 4) {
        float f0 = fmaf(x[o], y[o], s0), f1 = fmaf(x[o + 1], y[o + 1], s1);
        float f2 = fmaf(x[o + 2], y[o + 2], s2), f3 = fmaf(x[o + 3], y[o + 3], s3);
        res = (f0 + f2) + (f1 + f3);
        o += 4;
        r -= 4;
    }
    for (size_t i = 0; i < r; i++) res = fmaf(x[o + i], y[o + i], res);
    return res;
}

ORC_API void orc_pairwise(const float* x, int64_t nx, const float* y, int64_t ny, int64_t d, int ip, float* out) {
    for (int64_t i = 0; i < nx; i++)
        for (int64_t j = 0; j < ny; j++)
            out[i * ny + j] = ip ? orc_ip(x + i * d, y + j * d, d) : orc_l2sqr(x + i * d, y + j * d, d);
}

/* ---------------------------------------------------------------------------------------------
 * TypedTopKBuffer<float,int64_t> (src/cpp/include/list_scanning.h:41-204): append buffer, flush =
 * partial_sort by distance only (ascending for l2, descending for ip), so ties are unordered in the
 * reference. This restatement orders ties by ascending id, one valid outcome of that comparator.
 * --------------------------------------------------------------------------------------------- */
typedef struct {
    int k, cap, n, desc;
    float* dist;
    int64_t* id;
} orc_topk;

typedef struct { float d; int64_t id; } orc_pair;
static int g_desc = 0;
static int cmp_pair(const void* a, const void* b) {
    const orc_pair* x = (const orc_pair*)a;
    const orc_pair* y = (const orc_pair*)b;
    if (x->d != y->d) {
        if (g_desc) return x->d > y->d ? -1 : 1;
        return x->d < y->d ? -1 : 1;
    }
    return x->id < y->id ? -1 : (x->id > y->id ? 1 : 0);
}

static void topk_init(orc_topk* t, int k, int desc, int cap) {
    t->k = k; t->cap = cap < k ? k : cap; t->n = 0; t->desc = desc;
    t->dist = (float*)malloc(sizeof(float) * (size_t)t->cap);
    t->id = (int64_t*)malloc(sizeof(int64_t) * (size_t)t->cap);
}
static void topk_free(orc_topk* t) { free(t->dist); free(t->id); }

/* list_scanning.h:151-173 */
static void topk_flush(orc_topk* t) {
    orc_pair* p = (orc_pair*)malloc(sizeof(orc_pair) * (size_t)(t->n > 0 ? t->n : 1));
    for (int i = 0; i < t->n; i++) { p[i].d = t->dist[i]; p[i].id = t->id[i]; }
    g_desc = t->desc;
    qsort(p, (size_t)t->n, sizeof(orc_pair), cmp_pair);
    if (t->n > t->k) t->n = t->k;
    for (int i = 0; i < t->n; i++) { t->dist[i] = p[i].d; t->id[i] = p[i].id; }
    free(p);
}
/* list_scanning.h:117-122 */
static void topk_add(orc_topk* t, float d, int64_t id) {
    if (t->n >= t->cap) topk_flush(t);
    t->dist[t->n] = d; t->id[t->n] = id; t->n++;
}
/* list_scanning.h:187-191: the k-th distance, or the sentinel while fewer than k entries exist */
static float topk_kth(orc_topk* t) {
    topk_flush(t);
    if (t->n >= t->k) return t->dist[t->k - 1];
    return t->desc ? -INFINITY : FLT_MAX;
}

ORC_API void orc_topk_stream(const float* dist, const int64_t* ids, int64_t n, int k, int desc, int cap,
                             int64_t* out_ids, float* out_dist, int* out_n, float* out_kth) {
    orc_topk t;
    topk_init(&t, k, desc, cap);
    for (int64_t i = 0; i < n; i++) topk_add(&t, dist[i], ids[i]);
    *out_kth = topk_kth(&t);
    *out_n = t.n;
    for (int i = 0; i < t.n; i++) { out_ids[i] = t.id[i]; out_dist[i] = t.dist[i]; }
    topk_free(&t);
}

/* scan_list (list_scanning.h:241-311): one query x one list; l2 distances are sqrt'ed (:260, :286) */
static void scan_list(const float* q, const float* vecs, const int64_t* ids, int64_t n, int64_t d, int ip, orc_topk* t) {
    for (int64_t l = 0; l < n; l++) {
        float v = ip ? orc_ip(q, vecs + l * d, d) : sqrtf(orc_l2sqr(q, vecs + l * d, d));
        topk_add(t, v, ids ? ids[l] : l);
    }
}

ORC_API void orc_scan_list(const float* q, const float* vecs, const int64_t* ids, int64_t n, int64_t d, int ip, int k,
                           int64_t* out_ids, float* out_dist, int* out_n) {
    orc_topk t;
    topk_init(&t, k, ip, 8192); /* TOP_K_BUFFER_CAPACITY, list_scanning.h:39 */
    scan_list(q, vecs, ids, n, d, ip, &t);
    topk_flush(&t);
    *out_n = t.n;
    for (int i = 0; i < t.n; i++) { out_ids[i] = t.id[i]; out_dist[i] = t.dist[i]; }
    topk_free(&t);
}

/* ---------------------------------------------------------------------------------------------
 * APS geometry (src/cpp/include/geometry.h)
 * --------------------------------------------------------------------------------------------- */
#define NUM_X_VALUES 1001
#define STOP 1.0e-8
#define TINY 1.0e-30

/* geometry.h:115-161 (Lentz continued fraction) */
ORC_API double orc_incomplete_beta(double a, double b, double x) {
    if (x < 0.0 || x > 1.0) return INFINITY;
    if (x > (a + 1.0) / (a + b + 2.0)) return 1.0 - orc_incomplete_beta(b, a, 1.0 - x);
    const double lbeta_ab = lgamma(a) + lgamma(b) - lgamma(a + b);
    const double front = exp(log(x) * a + log(1.0 - x) * b - lbeta_ab) / a;
    double f = 1.0, c = 1.0, dd = 0.0;
    for (int i = 0; i <= 200; ++i) {
        int m = i / 2;
        double numerator;
        if (i == 0) numerator = 1.0;
        else if (i % 2 == 0) numerator = (m * (b - m) * x) / ((a + 2.0 * m - 1.0) * (a + 2.0 * m));
        else numerator = -((a + m) * (a + b + m) * x) / ((a + 2.0 * m) * (a + 2.0 * m + 1));
        dd = 1.0 + numerator * dd;
        if (fabs(dd) < TINY) dd = TINY;
        dd = 1.0 / dd;
        c = 1.0 + numerator / c;
        if (fabs(c) < TINY) c = TINY;
        const double cd = c * dd;
        f *= cd;
        if (fabs(1.0 - cd) < STOP) return front * (f - 1.0);
    }
    return INFINITY;
}

/* geometry.h:163-211. The reference keeps ONE process-global table initialised with the first d it
 * sees; this restatement keeps one table per call site by taking the table as an argument. */
ORC_API void orc_beta_table(int d, double* table /* [NUM_X_VALUES] */) {
    double dx = 1.0 / (NUM_X_VALUES - 1);
    double a = (d + 1.0) / 2.0, b = 0.5;
    for (int i = 0; i < NUM_X_VALUES; i++) table[i] = orc_incomplete_beta(a, b, i * dx);
}
static double beta_lookup(const double* table, double x) {
    x = fmax(0.0, fmin(1.0, x));
    double scaled = x * (NUM_X_VALUES - 1);
    int xi = (int)scaled;
    if (xi < 0) xi = 0;
    if (xi > NUM_X_VALUES - 2) xi = NUM_X_VALUES - 2;
    double y1 = table[xi], y2 = table[xi + 1];
    double dx = 1.0 / (NUM_X_VALUES - 1);
    double x1 = xi * dx;
    return y1 + (x - x1) * (y2 - y1) / dx;
}

/* geometry.h:57-113 */
ORC_API void orc_boundary_distances(const float* q, const float* const* cents, int m, int d, int euclid, float* out) {
    float* line = (float*)malloc(sizeof(float) * (size_t)d);
    float* mid = (float*)malloc(sizeof(float) * (size_t)d);
    float* resid = (float*)malloc(sizeof(float) * (size_t)d);
    for (int j = 0; j < m; j++) out[j] = -1.0f;
    const float* c0 = cents[0];
    if (euclid) {
        for (int i = 0; i < d; i++) resid[i] = q[i] - c0[i]; /* faiss::fvec_sub */
        for (int j = 1; j < m; j++) {
            for (int i = 0; i < d; i++) line[i] = cents[j][i] - c0[i];
            float A2 = orc_ip(line, line, (size_t)d);
            float A = sqrtf(A2);
            float dot = orc_ip(resid, line, (size_t)d);
            out[j] = fabsf(dot - 0.5f * A2) / A;
        }
    } else {
        for (int j = 1; j < m; j++) {
            for (int i = 0; i < d; i++) line[i] = cents[j][i] - c0[i];
            for (int i = 0; i < d; i++) mid[i] = line[i] / 2.0f;
            for (int i = 0; i < d; i++) mid[i] = c0[i] + mid[i];
            float norm = sqrtf(orc_ip(mid, mid, (size_t)d));
            for (int i = 0; i < d; i++) mid[i] = mid[i] / norm;
            float ang = orc_ip(q, mid, (size_t)d);
            out[j] = acosf(ang);
        }
    }
    free(line); free(mid); free(resid);
}

/* geometry.h:247-295 */
static double log_cap_volume(double radius, double boundary, int d, int use_precomputed, int euclid, const double* table) {
    double h = radius - boundary;
    h = fmax(0.0, fmin(2 * radius, h));
    if (euclid) {
        double x = sqrt((2 * radius * h - h * h) / (radius * radius));
        double ib = use_precomputed ? beta_lookup(table, x) : orc_incomplete_beta((d + 1.0) / 2.0, 0.5, x);
        if (ib <= 0.0 || isnan(ib) || isinf(ib)) return -INFINITY;
        return log(0.5) + log(ib);
    } else {
        double l1 = log(orc_incomplete_beta((d - 1) / 2.0, 0.5, sin(radius / 2.0) * sin(radius / 2.0)));
        double l2 = log(orc_incomplete_beta((d - 1) / 2.0, 0.5, sin(boundary / 2.0) * sin(boundary / 2.0)));
        return log(0.5) + l1 - l2;
    }
}

/* geometry.h:345-407. Returns 0, or -1 when the reference would throw (fewer than 2 partitions). */
ORC_API int orc_recall_profile(const float* boundary, int m, float query_radius, int d, int use_precomputed, int euclid,
                               const double* table, float* probs) {
    if (m < 2) return -1;
    for (int j = 0; j < m; j++) probs[j] = 0.0f;
    for (int j = 1; j < m; j++) {
        float b = boundary[j];
        if (b >= query_radius) { probs[j] = 0.0; continue; }
        double v = exp(log_cap_volume(query_radius, b, d, use_precomputed, euclid, table));
        probs[j] = (v > 0.0) ? v : 0.0;
    }
    probs[0] = 2.0 * probs[1];
    double sum = 0.0;
    for (int j = 0; j < m; j++) sum += probs[j];
    if (sum > 0.0f) for (int j = 0; j < m; j++) probs[j] /= sum;
    else for (int j = 0; j < m; j++) probs[j] = 1.0 / m;
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * serial_scan (src/cpp/src/query_coordinator.cpp:471-611): per query, scan the probed lists in rank
 * order; APS early exit (:557-579); pad with -1 / +-inf (:586-601).
 *
 * lists: list l has vectors list_vecs[l] ([list_n[l] x d]) and ids list_ids[l]; probe [Q x nprobe]
 * holds list indices (-1 = skip). cents [Q x nprobe] centroid pointers are only read under APS.
 * --------------------------------------------------------------------------------------------- */
ORC_API int orc_serial_scan(const float* queries, int64_t Q, int64_t d, const float* const* list_vecs,
                            const int64_t* const* list_ids, const int64_t* list_n, const int64_t* probe, int nprobe,
                            int ip, int k, float recall_target, float recompute_threshold, int use_precomputed,
                            const float* const* cents, int64_t* out_ids, float* out_dist, int* out_scanned) {
    int use_aps = recall_target > 0.0f && cents != NULL;
    double* table = NULL;
    if (use_aps && use_precomputed && !ip) {
        table = (double*)malloc(sizeof(double) * NUM_X_VALUES);
        orc_beta_table((int)d, table);
    }
    float* boundary = (float*)malloc(sizeof(float) * (size_t)(nprobe > 0 ? nprobe : 1));
    float* probs = (float*)calloc((size_t)(nprobe > 0 ? nprobe : 1), sizeof(float));
    int rc = 0;
    for (int64_t q = 0; q < Q && rc == 0; q++) {
        orc_topk t;
        topk_init(&t, k, ip, 8192);
        const float* qv = queries + q * d;
        float query_radius = ip ? -1000000.0f : 1000000.0f;
        int have_probs = 0, scanned = 0;
        if (use_aps) orc_boundary_distances(qv, cents + q * nprobe, nprobe, (int)d, !ip, boundary);
        for (int p = 0; p < nprobe; p++) {
            int64_t pi = probe[q * nprobe + p];
            if (pi == -1) continue;
            scan_list(qv, list_vecs[pi], list_ids[pi], list_n[pi], d, ip, &t);
            scanned++;
            float curr = topk_kth(&t);
            float change = fabsf(curr - query_radius) / curr;
            if (use_aps) {
                if (change > recompute_threshold) {
                    query_radius = curr;
                    if (orc_recall_profile(boundary, nprobe, query_radius, (int)d, use_precomputed, !ip, table, probs)) {
                        rc = -1; /* "Boundary distances must have at least 2 partitions" */
                        break;
                    }
                    have_probs = 1;
                }
                float est = 0.0f;
                if (have_probs) for (int i = 0; i < p; i++) est += probs[i];
                if (est >= recall_target) break;
            }
        }
        topk_flush(&t);
        for (int i = 0; i < k; i++) {
            if (i < t.n) { out_ids[q * k + i] = t.id[i]; out_dist[q * k + i] = t.dist[i]; }
            else { out_ids[q * k + i] = -1; out_dist[q * k + i] = ip ? -INFINITY : INFINITY; }
        }
        if (out_scanned) out_scanned[q] = scanned;
        topk_free(&t);
    }
    free(boundary); free(probs); free(table);
    return rc;
}

/* ---------------------------------------------------------------------------------------------
 * batched_scan_list, exact branch (list_scanning.h:313-366 with faiss knn_L2sqr / knn_inner_product
 * for nx < 20, faiss/utils/distances.cpp:133-203): per query top-min(k, n), l2 sqrt'ed afterwards.
 * (For nx >= 20 the reference goes through BLAS sgemm + norms, distances.cpp:262-343; its last-bit
 * rounding depends on the BLAS build and is compared with a tolerance, see tests/.)
 * --------------------------------------------------------------------------------------------- */
ORC_API void orc_batched_scan_list(const float* queries, int64_t nq, const float* vecs, const int64_t* ids, int64_t n,
                                   int64_t d, int ip, int k, int64_t* out_ids, float* out_dist, int* out_n) {
    int kmax = k < n ? k : (int)n;
    for (int64_t q = 0; q < nq; q++) {
        orc_topk t;
        topk_init(&t, kmax > 0 ? kmax : 1, ip, 8192);
        if (kmax > 0) {
            for (int64_t l = 0; l < n; l++) {
                float v = ip ? orc_ip(queries + q * d, vecs + l * d, d) : orc_l2sqr(queries + q * d, vecs + l * d, d);
                topk_add(&t, v, l);
            }
            topk_flush(&t);
        }
        int cnt = kmax > 0 ? t.n : 0;
        out_n[q] = cnt;
        for (int i = 0; i < k; i++) {
            if (i < cnt) {
                out_ids[q * k + i] = ids ? ids[t.id[i]] : t.id[i];
                out_dist[q * k + i] = ip ? t.dist[i] : sqrtf(t.dist[i]);
            } else {
                out_ids[q * k + i] = -1;
                out_dist[q * k + i] = ip ? -INFINITY : INFINITY;
            }
        }
        topk_free(&t);
    }
}

/* ---------------------------------------------------------------------------------------------
 * k-means pieces
 * --------------------------------------------------------------------------------------------- */
/* nearest centroid, ties to the lowest index (faiss Top1BlockResultHandler uses a strict comparison,
 * faiss/impl/ResultHandler.h:143) */
ORC_API void orc_assign(const float* x, int64_t n, const float* c, int64_t K, int64_t d, int ip, int32_t* assign) {
    for (int64_t i = 0; i < n; i++) {
        int64_t best = 0;
        float bv = ip ? -INFINITY : INFINITY;
        for (int64_t j = 0; j < K; j++) {
            float v = ip ? orc_ip(x + i * d, c + j * d, d) : orc_l2sqr(x + i * d, c + j * d, d);
            if (ip ? (v > bv) : (v < bv)) { bv = v; best = j; }
        }
        assign[i] = (int32_t)best;
    }
}

/* faiss compute_centroids (faiss/Clustering.cpp:123-192): per-centroid fp32 sum in data order, then
 * multiply by 1/count; also the accumulation of kmeans_refine_partitions (clustering.cpp:162-175),
 * which divides instead (clustering.cpp:122-124). */
ORC_API void orc_centroid_sums(const float* x, int64_t n, int64_t d, const int32_t* assign, int64_t K, float* sums,
                               int64_t* counts) {
    memset(sums, 0, sizeof(float) * (size_t)(K * d));
    memset(counts, 0, sizeof(int64_t) * (size_t)K);
    for (int64_t i = 0; i < n; i++) {
        int64_t c = assign[i];
        for (int64_t j = 0; j < d; j++) sums[c * d + j] += x[i * d + j];
        counts[c]++;
    }
}
