// Micro-benchmark (developer tool, not product): cycles per warp-instruction per SM sub-partition for packed /
// scalar fp32 and mixes, measured with clock64() inside the kernel (independent of the SM clock), one CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_cycles pipe_cycles.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int ITERS = 20000;
constexpr int CH = 8;

template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc, float seed) {
    __shared__ float2 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float2(i, -i);
    __syncthreads();
    float a[CH], b[CH]; float2 p[CH]; int n[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = seed + i + threadIdx.x; b[i] = seed * 0.5f + i; p[i] = make_float2(a[i], b[i]); n[i] = threadIdx.x + i; }
    const float2 c2 = make_float2(seed, seed * 1.0001f), d2 = make_float2(0.999f, 1.001f);
    const float2* q = sm + (threadIdx.x & 1023);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (OP == 0) p[i] = __fadd2_rn(p[i], c2);
            if (OP == 1) p[i] = __ffma2_rn(p[i], d2, c2);
            if (OP == 2) p[i] = __fmul2_rn(p[i], d2);
            if (OP == 3) a[i] = __fadd_rn(a[i], b[i]);
            if (OP == 4) a[i] = __fmaf_rn(a[i], seed, b[i]);
            if (OP == 5) { p[i] = __fadd2_rn(p[i], c2); a[i] = __fadd_rn(a[i], b[i]); }
            if (OP == 6) { p[i] = __fadd2_rn(p[i], c2); n[i] = (n[i] + it) ^ i; }
            if (OP == 7) { p[i] = __fadd2_rn(p[i], c2); n[i] = (n[i] + it) ^ i; a[i] = fminf(a[i], b[i] + 0.f * n[i]); }
            if (OP == 8) { p[i] = __fadd2_rn(p[i], c2); if (i == 0) { float2 v = q[(it & 1) * 1024]; p[1].x += v.x * 0.f; } }
            if (OP == 9) { p[i] = __fadd2_rn(p[i], c2); if ((i & 1) == 0) { float2 v = q[(it & 1) * 1024 + (i >> 1) * 0]; a[i] = v.x; } }
            if (OP == 10) { p[i] = __fadd2_rn(p[i], c2); p[i] = __ffma2_rn(p[i], d2, c2); }
            if (OP == 11) { p[i] = __fadd2_rn(p[i], make_float2(p[(i + 1) % CH].y, -p[(i + 1) % CH].x)); }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i] + b[i] + p[i].x + p[i].y + n[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    float* out; long long* cyc; CK(cudaMalloc(&out, sizeof(float) * sms * 1024)); CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    const char* names[] = {"FADD2", "FFMA2", "FMUL2", "FADD", "FFMA", "FADD2+FADD", "FADD2+2ALU", "FADD2+2ALU+FMNMX..", "8FADD2+1LDS64", "8FADD2+4LDS64", "FADD2+FFMA2", "FADD2 swap/neg"};
    const int fma_per[] = {1, 1, 1, 1, 1, 2, 1, 1, 1, 1, 2, 1};
    for (int threads : {128, 256, 512, 1024}) {
        for (int op = 0; op < 12; ++op) {
            for (int rep = 0; rep < 2; ++rep) {
                switch (op) {
#define C(n) case n: k<n><<<sms, threads>>>(out, cyc, 1.0001f); break;
                    C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11)
                }
                CK(cudaDeviceSynchronize());
            }
            long long h[256]; CK(cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
            double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
            const double warps_per_smsp = threads / 32 / 4.0;
            const double fma_instr = (double)ITERS * CH * fma_per[op] * warps_per_smsp;   // per SMSP
            printf("warps/SMSP %4.1f  %-20s %7.3f cycles per fma-pipe warp-instr per SMSP\n", warps_per_smsp, names[op], avg / fma_instr);
        }
    }
    return 0;
}
