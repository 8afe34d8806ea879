// Developer probe (not product): can one shared-memory copy of the mixed signal serve all five K chunks of the
// Toeplitz-GEMM FIR?  Row (g, s) of the A operand for chunk c must be the 32-sample block c rows further down the same
// mini-stream, i.e. the descriptor start address moves by c ROWS while the 8-row groups stay SBO apart (SBO = 12 rows:
// every group of 8 blocks carries its own 4 history blocks).  Variants:
//   v = 0  128-byte swizzle, rows 128 B, SBO = 1536, start + c * 128, data swizzled by ABSOLUTE address bits, base_offset 0
//   v = 1  same, base_offset = (start >> 7) & 7
//   v = 2  no swizzle, 16-byte K chunks in planes LBO apart, rows 16 B, SBO = 192, start + c * 16
// and the time of a 60-MMA chain (the FIR's K = 160, three TF32 terms) in each layout, with one and two accumulators.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcgen05_shift_probe tcgen05_shift_probe.cu
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 32, KC = 32;           // one K chunk = 4 MMA steps of 8
constexpr int GROUPS = 16, GROWS = 12;            // 16 groups of 8 rows, each with 4 rows of history in front
constexpr int ROWS = GROUPS * GROWS;              // 192 rows
constexpr int NV = 3, NC = 5, NT = 6;

__host__ __device__ inline float a_val(int row, int kk) { return (float)((row * 7 + kk * 3) % 17 - 8); }
__host__ __device__ inline float b_val(int n, int kk) { return (float)((n * 5 + kk * 11) % 13 - 6) * 0.5f; }

__device__ inline uint64_t desc_sw128(uint32_t addr, uint32_t sbo, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ inline uint64_t desc_none(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ inline void mma(uint32_t tm, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

constexpr uint32_t LBO_P = ROWS * 16 + 16;        // plane pitch of the no-swizzle layout (3088 B)

__global__ void __launch_bounds__(128) probe(float* D, float* cycles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = smem;                                  // swizzled rows: 192 x 128 B = 24 KB
    unsigned char* sP = sA + ROWS * 128;                       // no-swizzle planes: 8 x LBO_P
    unsigned char* sB = sP + 25 * 1024;                        // B: 32 x 32 floats, 128-byte swizzle, 4 KB
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x, warp = t >> 5;
    const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(sA), p0 = (uint32_t)__cvta_generic_to_shared(sP),
                   b0 = (uint32_t)__cvta_generic_to_shared(sB);
    for (int e = t; e < ROWS * KC; e += 128) {
        const int row = e / KC, kk = e % KC;
        uint32_t lin = a0 + row * 128 + kk * 4;
        lin ^= ((lin >> 7) & 7) << 4;                          // swizzle by absolute shared-memory address bits
        *reinterpret_cast<float*>(sA + (lin - a0)) = a_val(row, kk);
        *reinterpret_cast<float*>(sP + (kk / 4) * LBO_P + row * 16 + (kk % 4) * 4) = a_val(row, kk);
    }
    for (int e = t; e < N * KC; e += 128) {
        const int n = e / KC, kk = e % KC;
        *reinterpret_cast<float*>(sB + (n / 8) * 1024 + (n % 8) * 128 + (((kk / 4) ^ (n % 8)) * 16) + (kk % 4) * 4) = b_val(n, kk);
    }
    if (t == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t barp = (uint32_t)__cvta_generic_to_shared(&bar);
    uint32_t phase = 0;
    auto wait_all = [&]() {
        asm volatile("{\n\t.reg .pred p;\n\tWL_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WL_%=;\n\t}" ::"r"(barp), "r"(phase) : "memory");
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };
    auto commit = [&]() {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barp) : "memory");
    };
    auto desc_a = [&](int v, int c, int j) -> uint64_t {
        if (v == 2) return desc_none(p0 + 2 * j * LBO_P + c * 16, LBO_P, GROWS * 16);
        const uint32_t start = a0 + c * 128 + j * 32;
        return desc_sw128(start, GROWS * 128, v == 1 ? (start >> 7) & 7 : 0);
    };
    // ---- correctness: every variant, every shift ---------------------------------------------------------------
    for (int v = 0; v < NV; ++v) {
        for (int c = 0; c < NC; ++c) {
            if (t == 0) {
                for (int j = 0; j < KC / 8; ++j) mma(tm, desc_a(v, c, j), desc_sw128(b0 + j * 32, 1024, 0), idesc, j > 0);
                commit();
            }
            wait_all();
            uint32_t r[32];
            const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float* out = D + (size_t)(v * NC + c) * M * N;
            for (int n = 0; n < 32; ++n) out[t * N + n] = __uint_as_float(r[n]);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
    }
    // ---- timing: the FIR's 60-MMA chain (5 chunks x 4 K steps x 3 terms) --------------------------------------------
    for (int tt = 0; tt < NT; ++tt) {
        const int v = (tt % 3 == 2) ? 2 : 0;          // tests 0,1,3,4: swizzled; 2,5: planes
        const int nacc = (tt % 3 == 1) ? 2 : 1;       // tests 1,4: two accumulators (two independent tiles back to back)
        const int reps = (tt >= 3) ? 4 : 1;           // tests 3..5: four chains before the commit (steady-state rate)
        __syncthreads();
        const long long c0 = clock64();
        if (t == 0) {
            for (int rep = 0; rep < reps; ++rep)
                for (int a = 0; a < nacc; ++a)
                    for (int c = 0; c < NC; ++c)
                        for (int j = 0; j < KC / 8; ++j)
                            for (int s = 0; s < 3; ++s)
                                mma(tm + (uint32_t)((a & 1) * 32), desc_a(v, c, j), desc_sw128(b0 + j * 32, 1024, 0), idesc, (c | j | s) != 0);
            commit();
        }
        wait_all();
        const long long c1 = clock64();
        if (t == 0) cycles[tt] = (float)(c1 - c0) / (float)(reps * nacc);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm));
}

int main() {
    float *dD, *dC;
    const size_t nd = (size_t)NV * NC * M * N;
    cudaMalloc(&dD, nd * 4); cudaMalloc(&dC, NT * 4);
    cudaMemset(dD, 0xff, nd * 4);
    const int smem = 24 * 1024 + 25 * 1024 + 4 * 1024 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(dD, dC);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> D(nd), C(NT);
    cudaMemcpy(D.data(), dD, nd * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(C.data(), dC, NT * 4, cudaMemcpyDeviceToHost);
    const char* names[NV] = {"sw128 abs-swizzle base_off=0", "sw128 abs-swizzle base_off=(addr>>7)&7", "no-swizzle planes SBO=192"};
    int ok_any = 0;
    for (int v = 0; v < NV; ++v)
        for (int c = 0; c < NC; ++c) {
            int bad = 0;
            for (int r = 0; r < M; ++r)
                for (int n = 0; n < N; ++n) {
                    const int g = r / 8, s = r % 8;
                    float ref = 0;
                    for (int kk = 0; kk < KC; ++kk) ref += a_val(g * GROWS + s + c, kk) * b_val(n, kk);
                    if (!(fabsf(D[((size_t)(v * NC + c) * M + r) * N + n] - ref) <= 1e-3f)) ++bad;
                }
            printf("variant %d (%s) shift %d: %d mismatches of %d\n", v, names[v], c, bad, M * N);
            if (!bad && c > 0) ok_any |= 1 << v;
        }
    const char* tn[NT] = {"sw128 1 acc", "sw128 2 acc", "planes 1 acc", "sw128 1 acc x4", "sw128 2 acc x4", "planes 1 acc x4"};
    for (int i = 0; i < NT; ++i) printf("60-MMA chain, %s: %.0f cycles per chain (128 rows x 32 outputs)\n", tn[i], C[i]);
    printf("usable variants mask: %d\n", ok_any);
    return 0;
}
