// Library-level entry points: version, error reporting, device check.
#include "common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>

namespace qk {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return QK_ERR_CUDA;
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int pdl_mask() {
    // default off: measured at C2 (round 2, graph replay) the step is the same with and without it, 238-242 us
    static const int mask = getenv("QK_PDL") ? atoi(getenv("QK_PDL")) : 0;
    return mask;
}
long long launches() { return g_launches.load(std::memory_order_relaxed); }
int sm_count() {
    static int n[64] = {0};  // per device ordinal
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (n[dev] == 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        n[dev] = v > 0 ? v : 148;
    }
    return n[dev];
}
}  // namespace qk

extern "C" const char* qk_version(void) { return "quake_b200 0.1.0 (sm_100a)"; }
extern "C" const char* qk_last_error(void) { return qk::g_err; }
namespace qk { long long launches(); }
extern "C" long long qk_launch_count(void) { return qk::launches(); }

extern "C" int qk_device_check(int* sms, int* cc_major, int* cc_minor) {
    int dev = 0, n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        qk::set_error("no CUDA device available (%s); quake_b200 has no CPU fallback",
                      e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return QK_ERR_CUDA;
    }
    QK_CUDA(cudaGetDevice(&dev));
    int maj = 0, min = 0, sm = 0;
    QK_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    QK_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    QK_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    if (sms) *sms = sm;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    if (maj != 10) {
        qk::set_error("device has compute capability %d.%d; quake_b200 kernels are built for sm_100a only", maj, min);
        return QK_ERR_UNSUPPORTED;
    }
    return QK_OK;
}

// ---- host-side helpers mirroring vendored faiss control logic (no GPU work) ----------------------
#include <random>
#include <vector>
#include <unordered_map>

// First m entries of faiss::rand_perm(n, seed) (third_party/faiss/faiss/utils/random.cpp:153-163):
// Fisher-Yates with std::mt19937(seed) and `mt() % (n - i)`. Entry i of the permutation is final after
// step i, so only m steps are simulated, with a sparse map standing in for the untouched identity.
extern "C" int qk_host_rand_perm_prefix(int64_t n, int64_t seed, int64_t m, int64_t* out) {
    if (n < 0 || m < 0 || m > n || !out) {
        qk::set_error("qk_host_rand_perm_prefix: bad argument");
        return QK_ERR_INVALID_ARGUMENT;
    }
    std::mt19937 mt((unsigned int)seed);
    std::unordered_map<int64_t, int64_t> moved;
    moved.reserve((size_t)m * 2 + 16);
    auto get = [&](int64_t i) {
        auto it = moved.find(i);
        return it == moved.end() ? i : it->second;
    };
    for (int64_t i = 0; i < m; ++i) {
        if (i + 1 < n) {
            int64_t i2 = i + (int64_t)(mt() % (unsigned long)(n - i));
            int64_t a = get(i), b = get(i2);
            moved[i] = b;
            moved[i2] = a;
        }
        out[i] = get(i);
    }
    return QK_OK;
}

// faiss split_clusters (third_party/faiss/faiss/Clustering.cpp:204-251) on host arrays: every empty
// cluster takes a copy of a cluster drawn by size-proportional roulette (RandomGenerator(1234)) and both
// are perturbed by +-1/1024 on alternating dimensions. Returns the number of splits in *nsplit.
extern "C" int qk_host_split_clusters(int64_t d, int64_t k, int64_t n, float* hassign, float* centroids,
                                      int64_t pitch, int64_t* nsplit_out) {
    if (!hassign || !centroids || d <= 0 || k <= 0) {
        qk::set_error("qk_host_split_clusters: bad argument");
        return QK_ERR_INVALID_ARGUMENT;
    }
    const double EPS = 1 / 1024.;
    int64_t nsplit = 0;
    std::mt19937 mt(1234u);
    for (int64_t ci = 0; ci < k; ci++) {
        if (hassign[ci] == 0) {
            int64_t cj;
            int64_t guard = 0;
            for (cj = 0; true; cj = (cj + 1) % k) {
                float p = (hassign[cj] - 1.0) / (float)(n - k);
                float r = mt() / float(mt.max());
                if (r < p) break;
                if (++guard > (int64_t)1 << 40) break;
            }
            memcpy(centroids + ci * pitch, centroids + cj * pitch, sizeof(float) * d);
            for (int64_t j = 0; j < d; j++) {
                if (j % 2 == 0) {
                    centroids[ci * pitch + j] *= 1 + EPS;
                    centroids[cj * pitch + j] *= 1 - EPS;
                } else {
                    centroids[ci * pitch + j] *= 1 - EPS;
                    centroids[cj * pitch + j] *= 1 + EPS;
                }
            }
            hassign[ci] = hassign[cj] / 2;
            hassign[cj] -= hassign[ci];
            nsplit++;
        }
    }
    if (nsplit_out) *nsplit_out = nsplit;
    return QK_OK;
}
