// Tensor memory (TMEM, 256 KB per SM, sm_100a) as THREAD-PRIVATE scratch for loop-invariant per-thread tables and
// accumulators: the waterfall kernel fills the SM's register file (128 registers x 512 threads) and its shared memory
// (frame buffer), while the tensor memory of a kernel without MMAs sits idle.  tcgen05.ld writes registers without
// reading any (the kernel is bound by register-file READ bandwidth, profiles/r2c_ubench_ffma2_modifiers.txt), and
// takes the traffic off the shared-memory pipe.  Probe: scripts/ubench/tmem_scratch_probe.cu (profiles/r2a_ubench_tmem_scratch.txt).
#pragma once
#include <cuda_runtime.h>

#include "fft_radix.cuh"

namespace ssdr {

// ---- tensor memory as thread-private scratch ---------------------------------------------------------
// tcgen05.st / tcgen05.ld .32x32b: lane i of the warp <-> TMEM lane (32 (warp % 4) + i), consecutive registers <->
// consecutive columns.  Warps w, w + 4, w + 8, .. share a lane quarter and take disjoint column ranges.
SSDR_DEV unsigned tmem_alloc_cols(unsigned* slot_smem, int cols) {      // one warp calls; cols: power of two >= 32
    const unsigned sa = (unsigned)__cvta_generic_to_shared(slot_smem);
    switch (cols) {
        case 32: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(sa)); break;
        case 64: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(sa)); break;
        case 128: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(sa)); break;
        case 256: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(sa)); break;
        default: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sa)); break;
    }
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    return 0;
}
SSDR_DEV void tmem_free_cols(unsigned base, int cols) {
    switch (cols) {
        case 32: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(base)); break;
        case 64: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(base)); break;
        case 128: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(base)); break;
        case 256: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(base)); break;
        default: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base)); break;
    }
}
SSDR_DEV void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SSDR_DEV void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
SSDR_DEV void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
SSDR_DEV void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

SSDR_DEV void tmem_st8(unsigned addr, const unsigned (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
SSDR_DEV void tmem_ld8(unsigned addr, unsigned (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr) : "memory");
}
SSDR_DEV void tmem_st16(unsigned addr, const unsigned (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                 "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
SSDR_DEV void tmem_ld16(unsigned addr, unsigned (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(addr) : "memory");
}

}  // namespace ssdr
