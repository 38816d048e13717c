// Issue cost of small tcgen05.mma kind::tf32 (M = 128, N = 64, K = 8) under different issue patterns (one CTA):
//   0: one thread (if tid == 0), one asm block per MMA                       (the probe's first version)
//   1: one thread, four MMAs per asm block, descriptors precomputed
//   2: whole warp converged, election inside every asm block (the scan kernel's pattern)
//   3: whole warp converged, one election + four MMAs per asm block
//   4: like 3, descriptors advanced with immediates inside the asm block (one base per operand)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_issue scripts/umma_issue.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma1(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma1_elect(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma4(uint32_t d, uint64_t a0, uint64_t a1, uint64_t a2, uint64_t a3, uint64_t b0, uint64_t b1, uint64_t b2, uint64_t b3, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %10, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %5, %9, p;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %2, %6, %9, 1;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %3, %7, %9, 1;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %4, %8, %9, 1;\n\t}\n" ::"r"(d), "l"(a0), "l"(a1), "l"(a2), "l"(a3), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma4_elect(uint32_t d, uint64_t a0, uint64_t a1, uint64_t a2, uint64_t a3, uint64_t b0, uint64_t b1, uint64_t b2, uint64_t b3, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %10, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %5, %9, p;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %2, %6, %9, 1;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %3, %7, %9, 1;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %4, %8, %9, 1;\n\t}\n" ::"r"(d), "l"(a0), "l"(a1), "l"(a2), "l"(a3), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma4_elect_imm(uint32_t d, uint64_t a0, uint64_t b0, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, e;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "add.u64 a1, %1, 2;\n\tadd.u64 a2, %1, 4;\n\tadd.u64 a3, %1, 6;\n\tadd.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a1, b1, %3, 1;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a2, b2, %3, 1;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], a3, b3, %3, 1;\n\t}\n" ::"r"(d), "l"(a0), "l"(b0), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(128) rate(int reps, long long* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char* base = sm + ((1024u - (smem_u32(sm) & 1023u)) & 1023u);
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
    constexpr uint32_t idesc = make_idesc(128, 64);
    const uint64_t da0 = make_desc(smem_u32(base)), db0 = make_desc(smem_u32(base + 16384));
    long long t0 = 0, t1 = 0;
    if (MODE <= 1) {
        if (tid == 0) {
            t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                const uint32_t d = tmem + (r & 3) * 64;  // four accumulators round robin (like the scan kernel's tiles)
                if (MODE == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) mma1(d, da0 + 2 * k, db0 + 2 * k, idesc, k ? 1u : 0u);
                } else {
                    mma4(d, da0, da0 + 2, da0 + 4, da0 + 6, db0, db0 + 2, db0 + 4, db0 + 6, idesc, 0u);
                }
            }
            t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
    } else if (warp == 1) {
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t d = tmem + (r & 3) * 64;
            if (MODE == 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k) mma1_elect(d, da0 + 2 * k, db0 + 2 * k, idesc, k ? 1u : 0u);
            } else if (MODE == 3) {
                mma4_elect(d, da0, da0 + 2, da0 + 4, da0 + 6, db0, db0 + 2, db0 + 4, db0 + 6, idesc, 0u);
            } else {
                mma4_elect_imm(d, da0, db0, idesc, 0u);
            }
        }
        t1 = clock64();
        asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if ((MODE <= 1 && tid == 0) || (MODE > 1 && tid == 32)) { out[0] = t2 - t0; out[1] = t1 - t0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}
template <int MODE>
void run(long long* d, int reps) {
    CK(cudaFuncSetAttribute(rate<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    long long h[2];
    for (int it = 0; it < 2; ++it) {
        rate<MODE><<<1, 128, 64 * 1024>>>(reps, d);
        CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("mode %d: %.1f cycles per MMA to completion, %.1f to issue\n", MODE, (double)h[0] / (4 * reps), (double)h[1] / (4 * reps));
}
int main() {
    long long* d;
    CK(cudaMalloc(&d, 16));
    run<0>(d, 256); run<1>(d, 256); run<2>(d, 256); run<3>(d, 256); run<4>(d, 256);
    return 0;
}
