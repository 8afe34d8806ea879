// Developer probe (not product): the FIR of the demodulator as a Toeplitz GEMM on tcgen05, to pin layout and accuracy
// before building the kernel.  out[r][n] = sum_t h[t] z_r[126 + n - t], n < 32, for 128 rows
// (= 64 channels x re/im), K = 160 (126 history + 32 new samples + 2 pad), three TF32 MMAs per K step
// (A_hi B_hi + A_hi B_lo + A_lo B_hi, split by cvt.rna.tf32), operands in the 128-byte-swizzle K-major layout (a row
// of 32 samples is 128 contiguous bytes: conflict-free row-wise stores), accumulator in TMEM.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 32, K = 160, T = 127;


// 128B-swizzle K-major: atom (katom a, row group g) at (a * G + g) * 1024 bytes, G = rows / 8; inside the atom row r8 at
// r8 * 128, 16-byte chunk c at ((c ^ r8) * 16)
__host__ __device__ inline uint32_t sw_off(int row, int k, int groups) {
    const int a = k / 32, kk = k % 32, g = row / 8, r8 = row % 8, c = kk / 4;
    return (uint32_t)((a * groups + g) * 1024 + r8 * 128 + ((c ^ r8) * 16) + (kk % 4) * 4);
}

__device__ inline uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;                                   // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                         // SBO: next 8-row group
    d |= (uint64_t)1 << 46;                                   // version 1
    d |= (uint64_t)2 << 61;                                   // SWIZZLE_128B
    return d;
}

__device__ inline float tf32_hi(float x) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return __uint_as_float(u); }

__global__ void __launch_bounds__(128) probe(const float* Z, const float* H, float* D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sAh = smem;                                // 128 x 160 x 4 = 80 KB
    unsigned char* sAl = sAh + M * K * 4;
    unsigned char* sBh = sAl + M * K * 4;                     // 32 x 160 x 4 = 20 KB
    unsigned char* sBl = sBh + N * K * 4;
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x, warp = t >> 5;
    // A: row r, column k = sample z_r[k]  (each warp writes rows, lanes along k: 128 contiguous bytes per row)
    for (int e = t; e < M * K; e += 128) {
        const int r = e / K, k = e % K;
        const float x = Z[r * K + k], hi = tf32_hi(x);
        *reinterpret_cast<float*>(sAh + sw_off(r, k, M / 8)) = hi;
        *reinterpret_cast<float*>(sAl + sw_off(r, k, M / 8)) = x - hi;
    }
    // B: row n, column k = h[n + 126 - k] (Toeplitz of the taps), zero outside 0..126
    for (int e = t; e < N * K; e += 128) {
        const int n = e / K, k = e % K, tau = n + 126 - k;
        const float x = (tau >= 0 && tau < T) ? H[tau] : 0.f, hi = tf32_hi(x);
        *reinterpret_cast<float*>(sBh + sw_off(n, k, N / 8)) = hi;
        *reinterpret_cast<float*>(sBl + sw_off(n, k, N / 8)) = x - hi;
    }
    if (t == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    const long long c0 = clock64();
    if (t == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t ah = (uint32_t)__cvta_generic_to_shared(sAh), al = (uint32_t)__cvta_generic_to_shared(sAl);
        const uint32_t bh = (uint32_t)__cvta_generic_to_shared(sBh), bl = (uint32_t)__cvta_generic_to_shared(sBl);
        int first = 1;
        for (int j = 0; j < K / 8; ++j) {                     // K step j: atom j / 4, 32 bytes per step inside the 128-byte row
            const uint32_t oa = (uint32_t)((j / 4) * (M / 8) * 1024 + (j % 4) * 32);
            const uint32_t ob = (uint32_t)((j / 4) * (N / 8) * 1024 + (j % 4) * 32);
            const uint32_t pa[3] = {ah + oa, ah + oa, al + oa}, pb[3] = {bh + ob, bl + ob, bh + ob};
            for (int s = 0; s < 3; ++s) {
                const uint64_t da = make_desc_sw128(pa[s]), db = make_desc_sw128(pb[s]);
                const uint32_t acc = first ? 0u : 1u;
                first = 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    }
    {
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
        asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@!p bra W;\n\t}" ::"r"(b) : "memory");
    }
    const long long c1 = clock64();
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
    if (t == 0) D[M * N] = (float)(c1 - c0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tm));
}

int main() {
    std::vector<float> Z(M * K), H(T), D(M * N + 1);
    srand(7);
    for (auto& x : Z) x = (float)((rand() / (double)RAND_MAX * 2 - 1) * 3000.0);
    double hs = 0;
    for (int i = 0; i < T; ++i) {                             // Blackman-windowed sinc, the demodulator's recipe
        const double u = i - 63.0, fl = 1200.0 / 12000.0;
        const double sinc = (u == 0) ? 1.0 : sin(M_PI * 2 * fl * u) / (M_PI * 2 * fl * u);
        const double w = 0.42 - 0.5 * cos(2 * M_PI * i / 126.0) + 0.08 * cos(4 * M_PI * i / 126.0);
        H[i] = (float)(sinc * w); hs += H[i];
    }
    for (auto& x : H) x = (float)(x / hs);
    float *dZ, *dH, *dD;
    cudaMalloc(&dZ, Z.size() * 4); cudaMalloc(&dH, H.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dZ, Z.data(), Z.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dH, H.data(), H.size() * 4, cudaMemcpyHostToDevice);
    const int smem = (2 * M + 2 * N) * K * 4 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(dZ, dH, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double num = 0, den = 0, maxd = 0;
    for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) {
        double s = 0; for (int tt = 0; tt < T; ++tt) s += (double)H[tt] * (double)Z[r * K + 126 + n - tt];
        const double d = D[r * N + n] - s; num += d * d; den += s * s; if (fabs(d) > maxd) maxd = fabs(d);
    }
    printf("FIR on tcgen05 (3 x TF32): relative RMS error %.3e, max |diff| %.3e (signal rms %.1f); 60 MMAs in %.0f cycles\n",
           sqrt(num / den), maxd, sqrt(den / (M * N)), D[M * N]);
    return !(sqrt(num / den) < 1e-5);
}
