// Rate of small tcgen05.mma kind::tf32 instructions (M = 128, K = 8) as the scan kernel issues them: cycles per MMA for
// N in {16, 32, 64, 128, 256}, A from shared memory (SS) or from tensor memory (TS), issued back to back by one thread.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_rate scripts/umma_rate.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// out[0] = cycles from first issue to completion, out[1] = cycles spent issuing
__global__ void __launch_bounds__(128) rate(int N, int ts, int reps, int nacc, long long* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char* base = sm + ((1024u - (smem_u32(sm) & 1023u)) & 1023u);
    float* As = (float*)base;            // 16 KB
    float* Bs = (float*)(base + 16384);  // 32 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) ((float*)base)[i] = 1.0f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, N);
        const uint64_t da0 = make_desc(smem_u32(As)), db0 = make_desc(smem_u32(Bs));
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t dcol = tmem + (uint32_t)(((r * 4 + k) % nacc) * 64);  // independent accumulators, round robin
                if (ts) mma_ts(dcol, tmem + 256 + k * 8, db0 + 2 * k, idesc, (r * 4 + k) >= nacc ? 1u : 0u);
                else mma_ss(dcol, da0 + 2 * k, db0 + 2 * k, idesc, (r * 4 + k) >= nacc ? 1u : 0u);
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[0] = t2 - t0;
        out[1] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}
int main() {
    long long* d;
    CK(cudaMalloc(&d, 16));
    CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const int reps = 256;
    for (int nacc : {1, 2, 4})
    for (int ts = 0; ts < 2; ++ts)
        for (int N : {16, 32, 64, 128, 256}) {
            if (N * nacc > 256 && nacc > 1) continue;  // accumulators live in columns [0, 256)
            long long h[2];
            for (int it = 0; it < 2; ++it) {
                rate<<<1, 128, 64 * 1024>>>(N, ts, reps, nacc, d);
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
            printf("%d accumulators, %s N=%3d: %.1f cycles per MMA to completion, %.1f cycles per MMA to issue (%d MMAs)\n", nacc, ts ? "TS" : "SS", N,
                   (double)h[0] / (4 * reps), (double)h[1] / (4 * reps), 4 * reps);
        }
    return 0;
}
