// Standalone probe of the tcgen05 building blocks used by the scan kernel (sm_100a):
//   - kind::tf32 MMA, M=128, N=32, K=8, A and B K-major in shared memory with the 128-byte swizzle
//   - the same with A read from tensor memory (written with tcgen05.st 32x32b)
//   - accumulator read-back with tcgen05.ld 32x32b
//   - accuracy of the 3-pass split  A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  against a double reference
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_probe scripts/umma_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 B (32 floats), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);        // start address
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}

// instruction descriptor: D f32, A/B tf32, K-major both, N, M
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;             // c_format = F32
    d |= 2u << 7;             // a_format = TF32
    d |= 2u << 10;            // b_format = TF32
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// mode bit 0: pass hi*hi (SS), bit 1: lo*hi with A_lo in TMEM (TS), bit 2: hi*lo (SS), bit 3: lo*hi with A_lo in smem (SS)
__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* out, int mode) {
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char* base = sm + ((1024u - (smem_u32(sm) & 1023u)) & 1023u);
    float* As = (float*)base;                 // [128][32] swizzled, 16 KB
    float* Al = (float*)(base + 16384);       // A_lo, same layout
    float* Bs = (float*)(base + 32768);       // [32][32] swizzled, 4 KB
    float* Bl = (float*)(base + 36864);       // B_lo
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    auto sw = [](int r, int c) { return r * 32 + ((((c >> 2) ^ (r & 7)) << 2) | (c & 3)); };
    for (int i = tid; i < 128 * 32; i += 128) {
        int r = i / 32, c = i % 32;
        float a = A[i];
        float hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
        As[sw(r, c)] = a;
        Al[sw(r, c)] = a - hi;
    }
    for (int i = tid; i < 32 * 32; i += 128) {
        int r = i / 32, c = i % 32;
        float b = B[i];
        float hi = __uint_as_float(__float_as_uint(b) & 0xffffe000u);
        Bs[sw(r, c)] = b;
        Bl[sw(r, c)] = b - hi;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic smem writes -> async proxy (MMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t tmem_d = tmem;            // columns 0..31
    const uint32_t tmem_alo = tmem + 64;     // columns 64..95 (K = 32)

    // A_lo -> TMEM: thread = row (lane of TMEM), 32 columns = 32 k values
    {
        const int row = warp * 32 + lane;
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            float a = A[row * 32 + c];
            float hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
            v[c] = __float_as_uint(a - hi);
        }
        const uint32_t taddr = tmem_alo + ((uint32_t)(warp * 32) << 16);
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
            "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
            "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
            "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
            "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, 32);
        uint32_t acc = 0;
        for (int k = 0; k < 4; ++k) {  // 4 K-steps of 8 floats (32 B) inside the 128-byte swizzle row
            const uint64_t da = make_desc(smem_u32(As) + k * 32), dal = make_desc(smem_u32(Al) + k * 32);
            const uint64_t db = make_desc(smem_u32(Bs) + k * 32), dbl = make_desc(smem_u32(Bl) + k * 32);
            if (mode & 1) { mma_ss(tmem_d, da, db, idesc, acc); acc = 1; }
            if (mode & 2) { mma_ts(tmem_d, tmem_alo + k * 8, db, idesc, acc); acc = 1; }
            if (mode & 4) { mma_ss(tmem_d, da, dbl, idesc, acc); acc = 1; }
            if (mode & 8) { mma_ss(tmem_d, dal, db, idesc, acc); acc = 1; }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        uint32_t v[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int row = warp * 32 + lane;
#pragma unroll
        for (int n = 0; n < 32; ++n) out[row * 32 + n] = __uint_as_float(v[n]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
}

int main() {
    const int M = 128, N = 32, K = 32;
    std::vector<float> A(M * K), B(N * K), out(M * N);
    srand(1);
    for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& x : B) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dO;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dO, out.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
    auto trunc = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y; };
    const int modes[] = {1, 1 | 2 | 4, 1 | 8 | 4, 2, 8};
    for (int mode : modes) {
        CK(cudaMemset(dO, 0, out.size() * 4));
        probe<<<1, 128, 48 * 1024>>>(dA, dB, dO, mode);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
        double err_full = 0, err_model = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double full = 0, model = 0;
                for (int k = 0; k < K; ++k) {
                    float a = A[m * K + k], b = B[n * K + k];
                    float ah = trunc(a), bh = trunc(b), al = a - ah, bl = b - bh;
                    full += (double)a * b;
                    if (mode & 1) model += (double)ah * bh;
                    if (mode & (2 | 8)) model += (double)trunc(al) * bh;
                    if (mode & 4) model += (double)ah * trunc(bl);
                }
                err_full = fmax(err_full, fabs(out[m * N + n] - full));
                err_model = fmax(err_model, fabs(out[m * N + n] - model));
            }
        printf("mode %2d: max |D - exact A.B^T| = %.3e   max |D - truncation model| = %.3e   D[5][7]=%f\n", mode, err_full,
               err_model, out[5 * N + 7]);
    }
    return 0;
}
