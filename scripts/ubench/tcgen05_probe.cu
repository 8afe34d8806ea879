// Developer probe (not product): one tcgen05.mma kind::tf32 tile, D[128 x 32] = A[128 x K] * B[32 x K]^T, operands in
// shared memory in the no-swizzle K-major canonical layout, accumulator in TMEM, read back with tcgen05.ld and checked
// against the CPU.  Purpose: pin the descriptor encodings on this toolchain for a tensor-core FIR.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcgen05_probe tcgen05_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 32, K = 16;            // K = 2 MMA steps of 8
constexpr int KC = K / 4;                          // 16-byte chunks along K

// canonical no-swizzle K-major layout (units of 16 bytes): (row % 8) + (row / 8) * SBO + kchunk * LBO
constexpr uint32_t LBO_B = 128;                    // next core matrix along K
constexpr uint32_t SBO_B = KC * 128;               // next 8-row group
__device__ __host__ inline uint32_t canon_off(int row, int k) {   // byte offset of element (row, k), 4-byte elements
    return (row % 8) * 16 + (row / 8) * SBO_B + (k / 4) * LBO_B + (k % 4) * 4;
}

__device__ inline uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);            // start address
    d |= (uint64_t)((LBO_B >> 4) & 0x3fff) << 16;          // leading byte offset
    d |= (uint64_t)((SBO_B >> 4) & 0x3fff) << 32;          // stride byte offset
    d |= (uint64_t)1 << 46;                                // version = 1 (sm_100)
    // base_offset = 0, lbo_mode = 0, layout_type = 0 (no swizzle)
    return d;
}

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                               // M * K * 4 bytes
    unsigned char* sB = smem + M * K * 4;                   // N * K * 4 bytes
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x, warp = t >> 5;
    for (int e = t; e < M * K; e += 128) *reinterpret_cast<float*>(sA + canon_off(e / K, e % K)) = A[e];
    for (int e = t; e < N * K; e += 128) *reinterpret_cast<float*>(sB + canon_off(e / K, e % K)) = B[e];
    if (t == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // generic-proxy writes of the operands must be visible to the tensor-core (async) proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (t == 0) {
        // instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(sA), b0 = (uint32_t)__cvta_generic_to_shared(sB);
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t da = make_desc(a0 + ks * 2 * LBO_B), db = make_desc(b0 + ks * 2 * LBO_B);
            const uint32_t acc = ks > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    }
    // wait for the MMAs
    {
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
        asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@!p bra W;\n\t}" ::"r"(b) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 32; ++c) D[t * N + c] = __uint_as_float(v[c]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tm));
}

int main() {
    std::vector<float> A(M * K), B(N * K), D(M * N), R(M * N, 0.f);
    for (int i = 0; i < M * K; ++i) A[i] = (float)((i * 7) % 13 - 6);           // small integers: exact in TF32
    for (int i = 0; i < N * K; ++i) B[i] = (float)((i * 5) % 11 - 5) * 0.5f;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k]; R[m * N + n] = s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, D.size() * 4);
    probe<<<1, 128, (M + N) * K * 4>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0; double maxd = 0;
    for (int i = 0; i < M * N; ++i) { double d = fabs((double)D[i] - R[i]); if (!(d <= 1e-3)) ++bad; if (d > maxd) maxd = d; }
    printf("mismatches %d of %d, max |diff| %.3g; D[0..3] = %g %g %g %g  ref %g %g %g %g; D[row 64] %g ref %g\n", bad, M * N, maxd,
           D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3], D[64 * N], R[64 * N]);
    return bad != 0;
}
