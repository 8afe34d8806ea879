// Micro-benchmark (developer tool, not product): per-SM issue rates of scalar vs packed fp32 ops
// and shared-memory LDS/STS widths on B200.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

template <int OP>
__global__ void __launch_bounds__(1024) k_fp(float* out, float seed) {
    float a[CHAINS], b[CHAINS];
    float2 p[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { a[i] = seed + i + threadIdx.x; b[i] = seed * 0.5f + i; p[i] = make_float2(a[i], b[i]); }
    const float2 c2 = make_float2(seed, seed * 1.0001f), d2 = make_float2(0.999f, 1.001f);
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == 0) { a[i] = __fmaf_rn(a[i], seed, b[i]); }                       // FFMA
            else if (OP == 1) { a[i] = __fadd_rn(a[i], b[i]); }                        // FADD
            else if (OP == 2) { a[i] = __fmul_rn(a[i], seed); }                        // FMUL
            else if (OP == 3) { p[i] = __ffma2_rn(p[i], d2, c2); }                     // FFMA2
            else if (OP == 4) { p[i] = __fadd2_rn(p[i], c2); }                         // FADD2
            else if (OP == 5) { p[i] = __fmul2_rn(p[i], d2); }                         // FMUL2
            else if (OP == 6) { a[i] = __fadd_rn(a[i], b[i]); b[i] = __fmaf_rn(b[i], seed, a[i]); }  // FADD + FFMA mix (2 instr)
            else if (OP == 7) { p[i] = __fadd2_rn(p[i], c2); a[i] = __fadd_rn(a[i], b[i]); }         // FADD2 + FADD (2 instr)
            else if (OP == 8) { p[i] = __ffma2_rn(p[i], d2, c2); a[i] = (float)(__float_as_int(a[i]) ^ 0x80000000) ; } // FFMA2 + LOP
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i] + b[i] + p[i].x + p[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared memory: each thread reads / writes W-byte words at conflict-free addresses
template <int W, bool STORE>
__global__ void __launch_bounds__(1024) k_smem(float* out) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int t = threadIdx.x;
    float acc = 0.f;
    for (int i = t; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = (float)i;
    __syncthreads();
    const int nw = 48 * 1024 / W;
#pragma unroll 1
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            int idx = (t + u * blockDim.x + it) % nw;
            if (W == 4) { float* q = reinterpret_cast<float*>(sm) + idx; if (STORE) *q = acc; else acc += *q; }
            if (W == 8) { float2* q = reinterpret_cast<float2*>(sm) + idx; if (STORE) *q = make_float2(acc, acc); else { float2 v = *q; acc += v.x + v.y; } }
            if (W == 16) { float4* q = reinterpret_cast<float4*>(sm) + idx; if (STORE) *q = make_float4(acc, acc, acc, acc); else { float4 v = *q; acc += v.x + v.y + v.z + v.w; } }
        }
    }
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + reinterpret_cast<float*>(sm)[t];
}

template <class F>
float time_ms(F f) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("%s, %d SMs, clockRate attr %.0f MHz\n", p.name, sms, clk_khz / 1e3);
    float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 2 * 1024));
    const char* names[] = {"FFMA", "FADD", "FMUL", "FFMA2", "FADD2", "FMUL2", "FADD+FFMA", "FADD2+FADD", "FFMA2+LOP"};
    const int instr_per[] = {1, 1, 1, 1, 1, 1, 2, 2, 2};
    for (int threads : {256, 512, 1024}) {
        for (int op = 0; op < 9; ++op) {
            auto launch = [&]() {
                dim3 g(sms * (2048 / threads) / 1), b(threads);   // fill each SM to 2048 threads where registers allow
                switch (op) {
                    case 0: k_fp<0><<<g, b>>>(out, 1.0001f); break; case 1: k_fp<1><<<g, b>>>(out, 1.0001f); break;
                    case 2: k_fp<2><<<g, b>>>(out, 1.0001f); break; case 3: k_fp<3><<<g, b>>>(out, 1.0001f); break;
                    case 4: k_fp<4><<<g, b>>>(out, 1.0001f); break; case 5: k_fp<5><<<g, b>>>(out, 1.0001f); break;
                    case 6: k_fp<6><<<g, b>>>(out, 1.0001f); break; case 7: k_fp<7><<<g, b>>>(out, 1.0001f); break;
                    case 8: k_fp<8><<<g, b>>>(out, 1.0001f); break;
                }
            };
            float ms = time_ms(launch);
            double warp_instr = (double)sms * (2048 / threads) * (threads / 32) * ITERS * CHAINS * instr_per[op];
            // warp-instructions per ns per SM; x (clock GHz) gives per-cycle
            printf("threads/CTA %4d  %-11s %8.3f ms  %6.3f warp-instr/ns/SM  (%.2f lane-ops/ns/SM)\n", threads, names[op], ms,
                   warp_instr / sms / (ms * 1e6), warp_instr * 32 / sms / (ms * 1e6));
        }
    }
    CK(cudaFuncSetAttribute(k_smem<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
    for (int threads : {512, 1024}) {
        auto run = [&](const char* nm, int W, auto kern) {
            auto launch = [&]() { kern<<<sms * (2048 / threads), threads, 48 * 1024>>>(out); };
            float ms = time_ms(launch);
            double bytes = (double)sms * (2048 / threads) * threads * ITERS * W;
            printf("threads/CTA %4d  %-8s %8.3f ms  %7.1f B/ns/SM\n", threads, nm, ms, bytes / sms / (ms * 1e6));
        };
        run("LDS.32", 4, k_smem<4, false>); run("LDS.64", 8, k_smem<8, false>); run("LDS.128", 16, k_smem<16, false>);
        run("STS.32", 4, k_smem<4, true>); run("STS.64", 8, k_smem<8, true>); run("STS.128", 16, k_smem<16, true>);
    }
    return 0;
}
