// Shared host-side helpers for libssdr_b200 (error reporting, launch accounting).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/ssdr_b200.h"

namespace ssdr {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    return SSDR_E_CUDA;
}

#define SSDR_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return ssdr::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define SSDR_ARG(cond, ...)                                                    \
    do {                                                                       \
        if (!(cond)) { ssdr::set_error(__VA_ARGS__); return SSDR_E_ARG; }      \
    } while (0)

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int sm_count();

}  // namespace ssdr
