// Micro-benchmark (developer tool): issue rate of the legacy mma.sync m16n8k8 TF32 path on sm_100a, cycles per
// warp-instruction per SM sub-partition (clock64), to decide whether a Toeplitz-GEMM FIR on mma.sync is worth building.
#include <cuda_runtime.h>
#include <cstdio>
constexpr int ITERS = 20000, CH = 8;
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc) {
    float d[CH][4];
    unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f900000u, 0x3fa00000u, 0x3fb00000u}, b[2] = {0x3f800000u, 0x3f400000u + threadIdx.x};
    for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) d[i][j] = (float)(i + j);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    float* out; long long* cyc; cudaMalloc(&out, sizeof(float) * sms * 1024); cudaMalloc(&cyc, sizeof(long long) * sms);
    for (int threads : {128, 256, 512, 1024}) {
        for (int rep = 0; rep < 2; ++rep) { k<<<sms, threads>>>(out, cyc); cudaDeviceSynchronize(); }
        long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
        const double wps = threads / 32 / 4.0;
        const double cyc_per = avg / ((double)ITERS * CH * wps);
        printf("warps/SMSP %4.1f  mma.sync m16n8k8 tf32: %6.2f cycles per warp-instr per SMSP  => %.1f TFLOP/s dense (148 SMs, 1.965 GHz)\n",
               wps, cyc_per, 2.0 * 16 * 8 * 8 / cyc_per * 4 * sms * 1.965e9 / 1e12);
    }
    return 0;
}
