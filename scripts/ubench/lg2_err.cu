// Developer tool: exhaustive error of the quantiser's byte estimate v = fma(lg2.approx.ftz(P), c1, c0) against
// float64, over every float32 mantissa for a range of exponents (sets kQEps).
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
__device__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__global__ void k(int e, float c1, float c0, double c1d, double c0d, double* maxerr) {
    double m = 0.0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (1u << 23); i += gridDim.x * blockDim.x) {
        const float P = __uint_as_float(((unsigned)(e + 127) << 23) | i);
        const float v = __fmaf_rn(lg2_ftz(P), c1, c0);
        const double ex = log2((double)P) * c1d + c0d;
        m = fmax(m, fabs((double)v - ex));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)maxerr, __double_as_longlong(m));
}
int main() {
    double* d; cudaMalloc(&d, 8);
    for (int N : {256, 1024, 16384, 65536}) {
        double ref = (double)N * 32768.0 * 0.5; ref *= ref;
        const double c1d = 10.0 * log10(2.0), c0d = -10.0 * log10(ref) - 10.0 + 255.0;
        const float c1 = (float)c1d, c0 = (float)c0d;
        double worst = 0; int worst_e = 0;
        for (int e = -60; e <= 100; ++e) {
            const double vmid = (e + 0.5) * c1d + c0d;
            if (vmid < -2 || vmid > 257) continue;           // only exponents whose bytes are in range
            cudaMemset(d, 0, 8);
            k<<<592, 256>>>(e, c1, c0, c1d, c0d, d);
            double h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            if (h > worst) { worst = h; worst_e = e; }
        }
        printf("N=%6d  c1=%.9g c0=%.9g  max |v - exact| = %.3e at exponent %d\n", N, c1, c0, worst, worst_e);
    }
    return 0;
}
