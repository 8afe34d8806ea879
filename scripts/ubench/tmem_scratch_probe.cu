// Micro-benchmark / probe (developer tool, not product): TMEM as thread-private scratch.
//   * 512 threads (16 warps), one CTA per SM, tcgen05.alloc of all 512 columns; warp w owns lanes 32 (w % 4) .. +31 and the
//     128 columns starting at 128 (w / 4): 128 private 32-bit words per thread.
//   * correctness: every thread writes a pattern with tcgen05.st.32x32b.x32 / .x16, reads it back with tcgen05.ld.
//   * cost: cycles per (LDTM.x16 + 16 dependent FFMA2) iteration against the same with an LDS.128 x 4 column.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_scratch_probe tmem_scratch_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tm_st16(unsigned addr, const unsigned (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                 "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tm_ld16(unsigned addr, unsigned (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(unsigned* bad, long long* cyc, float* sink, int iters) {
    __shared__ unsigned tmem_base;
    extern __shared__ uint4 col[];                 // MODE 1: thread-private uint4[4] columns
    const int w = threadIdx.x >> 5;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned base = tmem_base + ((unsigned)(32 * (w & 3)) << 16) + 128u * (unsigned)(w >> 2);
    // ---- correctness: 128 words per thread ----
    unsigned nbad = 0;
    for (int c = 0; c < 8; ++c) {
        unsigned v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (threadIdx.x * 1000003u) ^ (unsigned)(c * 16 + i) * 2654435761u ^ blockIdx.x;
        tm_st16(base + 16 * c, v);
    }
    tm_wait_st();
    __syncthreads();
    for (int c = 7; c >= 0; --c) {
        unsigned v[16];
        tm_ld16(base + 16 * c, v);
        tm_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) nbad += v[i] != ((threadIdx.x * 1000003u) ^ (unsigned)(c * 16 + i) * 2654435761u ^ blockIdx.x);
    }
    atomicAdd(bad, nbad);
    // ---- cost ----
    float2 p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = make_float2(1.0f + threadIdx.x, 0.5f * i);
    uint4* mycol = col + threadIdx.x;
    for (int c = 0; c < 4; ++c) mycol[c * 512] = make_uint4(c, c + 1, c + 2, c + 3);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        unsigned v[16];
        if (MODE == 0) { tm_ld16(base + 16 * (it & 7), v); tm_wait_ld(); }
        else {
#pragma unroll
            for (int c = 0; c < 4; ++c) { const uint4 a = mycol[c * 512]; v[4 * c] = a.x + it; v[4 * c + 1] = a.y; v[4 * c + 2] = a.z; v[4 * c + 3] = a.w; }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                p[i] = __ffma2_rn(p[i], make_float2(0.999f, 1.001f), make_float2(__uint_as_float(v[2 * i] & 0x3fffffffu), __uint_as_float(v[2 * i + 1] & 0x3fffffffu)));
        if (MODE == 2) { unsigned u[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) u[i] = __float_as_uint(i & 1 ? p[i >> 1].y : p[i >> 1].x);
            tm_st16(base + 16 * (it & 7), u); tm_wait_st(); tm_ld16(base + 16 * ((it + 1) & 7), v); tm_wait_ld(); p[0].x += __uint_as_float(v[3] & 0x3fffffffu); }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += p[i].x + p[i].y;
    sink[blockIdx.x * 512 + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount, iters = 4000;
    unsigned* bad; long long* cyc; float* sink;
    CK(cudaMalloc(&bad, 4)); CK(cudaMemset(bad, 0, 4)); CK(cudaMalloc(&cyc, 8 * sms)); CK(cudaMalloc(&sink, 4 * sms * 512));
    const char* names[] = {"LDTM.x16 + 32 FFMA2", "4 x LDS.128 column + 32 FFMA2", "32 FFMA2 + STTM.x16 + wait + LDTM.x16 + wait"};
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            if (mode == 0) k<0><<<sms, 512, 32768>>>(bad, cyc, sink, iters);
            if (mode == 1) k<1><<<sms, 512, 32768>>>(bad, cyc, sink, iters);
            if (mode == 2) k<2><<<sms, 512, 32768>>>(bad, cyc, sink, iters);
            CK(cudaDeviceSynchronize());
        }
        long long h[256]; CK(cudaMemcpy(h, cyc, 8 * sms, cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
        unsigned hb; CK(cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost));
        printf("%-48s %8.1f cycles/iteration (16 warps/SM; 32 FFMA2 alone = ~%d at 2.06 x 4 warps/SMSP)   readback mismatches so far: %u\n",
               names[mode], avg / iters, (int)(32 * 2.06 * 4), hb);
    }
    return 0;
}
