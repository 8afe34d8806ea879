// Micro-benchmark (developer tool, not product): cost of the operand modifiers of the packed fp32 instructions
// (swap .LO_HI, half negation .NP/.PN, scalar broadcast .F32, whole negation) in cycles per warp-instruction per SM
// sub-partition, clock64() inside the kernel, one CTA of 512 threads per SM (4 warps per sub-partition), 8 independent
// chains per thread.  Check the SASS (cuobjdump -sass) for the modifiers each variant really gets.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ffma2_modifiers ffma2_modifiers.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int ITERS = 20000;
constexpr int CH = 8;

template <int OP>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, float seed) {
    float2 p[CH], q[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { p[i] = make_float2(seed + i + threadIdx.x, seed * 0.5f + i); q[i] = make_float2(0.5f * i + seed, 1.0f + i); }
    const float c = seed * 0.999f, s = seed * 0.001f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            float2& a = p[i]; const float2 b = q[i];
            if (OP == 0) a = __ffma2_rn(a, make_float2(c, s), b);                                            // plain, register pair
            if (OP == 1) a = __ffma2_rn(a, make_float2(c, c), b);                                            // scalar broadcast
            if (OP == 2) a = __ffma2_rn(make_float2(a.y, a.x), make_float2(c, c), b);                        // swap + broadcast
            if (OP == 3) a = __ffma2_rn(a, make_float2(c, c), make_float2(-b.x, b.y));                       // half negation of the addend
            if (OP == 4) a = __ffma2_rn(make_float2(a.y, a.x), make_float2(c, c), make_float2(-b.x, b.y));   // swap + half negation
            if (OP == 5) a = __ffma2_rn(b, make_float2(2.0f, 2.0f), make_float2(-a.x, -a.y));                // immediate 2, whole negation
            if (OP == 6) a = __fadd2_rn(a, b);                                                               // FADD2 plain
            if (OP == 7) a = __fadd2_rn(a, make_float2(b.y, -b.x));                                          // FADD2 swap + half negation (independent operand)
            if (OP == 8) a = __fadd2_rn(a, make_float2(-b.x, -b.y));                                         // FADD2 whole negation
            if (OP == 9) a = __fmul2_rn(make_float2(a.y, a.x), make_float2(c, c));                           // FMUL2 swap + broadcast (v1 cmul, first half)
            if (OP == 10) a = __ffma2_rn(a, make_float2(0.9238795f, 0.9238795f), make_float2(-b.x, b.y));    // immediate + half negation (v1 cmul, second half)
            if (OP == 11) { const float2 t2 = __ffma2_rn(make_float2(b.y, b.x), make_float2(s, s), make_float2(-a.x, a.y));      // v2 butterfly: 3 instructions
                            const float2 x = __ffma2_rn(b, make_float2(c, c), make_float2(-t2.x, t2.y));
                            q[i] = __ffma2_rn(a, make_float2(2.0f, 2.0f), make_float2(-x.x, -x.y)); a = x; }
            if (OP == 12) { const float2 x = __fadd2_rn(a, b), y = __fadd2_rn(a, make_float2(-b.x, -b.y));                       // v1 butterfly: add, sub, cmul = 4 instructions
                            const float2 t = __fmul2_rn(make_float2(y.y, y.x), make_float2(s, s));
                            q[i] = __ffma2_rn(y, make_float2(c, c), make_float2(-t.x, t.y)); a = x; }
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) r += p[i].x + p[i].y + q[i].x + q[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    float* out; long long* cyc; CK(cudaMalloc(&out, sizeof(float) * sms * 512)); CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    const char* names[] = {"FFMA2 pair operand", "FFMA2 scalar broadcast", "FFMA2 swap + broadcast", "FFMA2 addend half-neg", "FFMA2 swap + half-neg",
                           "FFMA2 imm 2, addend negated", "FADD2 plain", "FADD2 swap + half-neg", "FADD2 negated", "FMUL2 swap + broadcast",
                           "FFMA2 imm + half-neg", "v2 butterfly (3 instr)", "v1 butterfly (4 instr)"};
    const int per[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 3, 4};
    for (int op = 0; op < 13; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (op) {
#define C(n) case n: k<n><<<sms, 512>>>(out, cyc, 1.0001f); break;
                C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12)
            }
            CK(cudaDeviceSynchronize());
        }
        long long h[256]; CK(cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
        const double instr = (double)ITERS * CH * per[op] * 4.0;      // per sub-partition (4 warps)
        printf("%-32s %7.3f cycles per packed instruction per sub-partition   (%7.3f per unit)\n", names[op], avg / instr, avg / ((double)ITERS * CH * 4.0));
    }
    return 0;
}
