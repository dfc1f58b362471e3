// sm_100a tensor-core plumbing shared by the tcgen05 kernels: shared-memory matrix descriptors, instruction
// descriptors, TMEM allocation / load, mbarrier and proxy fences.  Raw PTX only (no CUTLASS dependency).
//
// Operand layouts used throughout (no swizzle; unit = one 16-byte chunk, "core matrix" = 8 chunks = 128 contiguous B):
//   K-major  (rows r along M/N, k along K, T = 16/sizeof(elt) elements per chunk)
//       addr(r, k) = (r % 8) * 16 + (r / 8) * SBO + (k / T) * LBO + (k % T) * sizeof(elt)
//   MN-major (i along M/N, j along K)
//       addr(i, j) = (i % T) * sizeof(elt) + (i / T) * SBO + (j % 8) * 16 + (j / 8) * LBO
// so ONE shared-memory tile written as K-major [r][k] is at the same time an MN-major [k][r]-transposed operand
// with LBO and SBO exchanged — used to read the genotype tile as X (encoder forward) and as X^T (encoder backward),
// and the G tile as G (dQ) and G^T (dP).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nadm {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared memory matrix descriptor (tcgen05, version 1, SWIZZLE_NONE) ------------------------------------------
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version for sm_100
    return d;
}

// Advance a descriptor's start address by `units16` 16-byte units.  Only the low word changes (the 14-bit address field
// of a valid shared-memory address cannot carry out of bits 0..13), so this is ONE 32-bit add instead of a 64-bit
// add-with-carry pair per descriptor on the issuer's instruction stream.
__device__ __forceinline__ uint64_t desc_add(uint64_t d, uint32_t units16) {
    return (d & 0xFFFFFFFF00000000ull) | (uint64_t)((uint32_t)d + units16);
}

// ---- instruction descriptor ---------------------------------------------------------------------------------------
enum : uint32_t { kFmtF16 = 0, kFmtBF16 = 1, kFmtTF32 = 2, kFmtU8 = 0, kFmtS8 = 1 };
enum : uint32_t { kAccF16 = 0, kAccF32 = 1, kAccS32 = 2 };
constexpr uint32_t instr_desc(uint32_t cfmt, uint32_t afmt, uint32_t bfmt, bool a_mn_major, bool b_mn_major, int M,
                              int N) {
    return (cfmt << 4) | (afmt << 7) | (bfmt << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- MMA issue (one thread) -----------------------------------------------------------------------------------------
#define NADM_DEF_MMA_SS(NAME, KIND)                                                                              \
    __device__ __forceinline__ void NAME(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,       \
                                         uint32_t accumulate) {                                                  \
        asm volatile(                                                                                            \
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                                     \
            "tcgen05.mma.cta_group::1.kind::" KIND " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),                  \
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)                                                  \
            : "memory");                                                                                         \
    }
NADM_DEF_MMA_SS(mma_i8_ss, "i8")
NADM_DEF_MMA_SS(mma_tf32_ss, "tf32")
NADM_DEF_MMA_SS(mma_f16_ss, "f16")
#undef NADM_DEF_MMA_SS

// A operand from tensor memory
#define NADM_DEF_MMA_TS(NAME, KIND)                                                                              \
    __device__ __forceinline__ void NAME(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,      \
                                         uint32_t accumulate) {                                                  \
        asm volatile(                                                                                            \
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                                     \
            "tcgen05.mma.cta_group::1.kind::" KIND " [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),                \
            "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)                                                 \
            : "memory");                                                                                         \
    }
NADM_DEF_MMA_TS(mma_i8_ts, "i8")
NADM_DEF_MMA_TS(mma_tf32_ts, "tf32")
NADM_DEF_MMA_TS(mma_f16_ts, "f16")
#undef NADM_DEF_MMA_TS

// all previously issued MMAs of this thread arrive on the mbarrier when they have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
                 : "memory");
}

// knock-out measurement build (-DNADM_KO_MMA): the predicated MMA issue below never fires (issue flag compared with 2),
// commits still arrive: shows what the rest of a kernel costs without its tensor-core work.  Garbage results by design.
#ifdef NADM_KO_MMA
#define NADM_KO_MMA_EXTRA "setp.eq.b32 q, %5, 2;\n\t"
#else
#define NADM_KO_MMA_EXTRA
#endif
// ---- predicated issue: the WHOLE warp executes the (convergent) descriptor arithmetic, only the lane whose `issue`
// flag is set executes the MMA / commit.  Keeping the issuing code free of a divergent `if (elected)` region lets the
// compiler hold descriptors and counters in uniform registers instead of moving them there (R2UR) per instruction.
#define NADM_DEF_MMA_SS_P(NAME, KIND)                                                                            \
    __device__ __forceinline__ void NAME(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,       \
                                         uint32_t accumulate, uint32_t issue) {                                  \
        asm volatile(                                                                                            \
            "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t" NADM_KO_MMA_EXTRA             \
            "@q tcgen05.mma.cta_group::1.kind::" KIND " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),              \
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(issue)                                      \
            : "memory");                                                                                         \
    }
NADM_DEF_MMA_SS_P(mma_i8_ss_p, "i8")
NADM_DEF_MMA_SS_P(mma_f16_ss_p, "f16")
#undef NADM_DEF_MMA_SS_P
#define NADM_DEF_MMA_TS_P(NAME, KIND)                                                                            \
    __device__ __forceinline__ void NAME(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,      \
                                         uint32_t accumulate, uint32_t issue) {                                  \
        asm volatile(                                                                                            \
            "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t" NADM_KO_MMA_EXTRA             \
            "@q tcgen05.mma.cta_group::1.kind::" KIND " [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),            \
            "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(issue)                                     \
            : "memory");                                                                                         \
    }
NADM_DEF_MMA_TS_P(mma_f16_ts_p, "f16")
NADM_DEF_MMA_TS_P(mma_i8_ts_p, "i8")
#undef NADM_DEF_MMA_TS_P
__device__ __forceinline__ void mma_commit_p(uint64_t* bar, uint32_t issue) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(issue)
        : "memory");
}

// ---- tensor memory -----------------------------------------------------------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_result)),
                 "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // the same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// 32 lanes x 32 bit, 8 / 16 / 32 consecutive columns: thread t of the warp gets lane (quadrant base + t)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(
            taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// ---- mbarrier / fences ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Hang watchdog: every wait of the pipelines below is bounded.  A failed try_wait suspends the thread for >= 25 ns, so
// 2^26 failed polls are > 1.5 s on a barrier that is normally reached within microseconds: the kernel then traps
// (the launch fails with a CUDA error that the C ABI reports) instead of spinning forever and taking the GPU with it.
// One predicated add per FAILED poll; nothing on the success path.  -DNADM_NO_WATCHDOG removes it.
#ifndef NADM_NO_WATCHDOG
#define NADM_WATCHDOG_POLL(n) do { if (++(n) > (1u << 26)) __trap(); } while (0)
#else
#define NADM_WATCHDOG_POLL(n) do { } while (0)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) NADM_WATCHDOG_POLL(polls);
}
// Wait of a warp that is NOT on the kernel's critical path (epilogues, operand producers running ahead): failed polls
// are spaced by nanosleep so that the polling loop does not compete for issue slots with the warps of its SM
// sub-partition (back-to-back try_wait polls were measured to be 10-20 % of all instructions issued; the try_wait
// suspend-time hint does not space them: the hardware wakes the thread after ~25 ns regardless).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(ns);
        NADM_WATCHDOG_POLL(polls);
    }
}
// named barrier among `nthreads` threads (whole warps) of the CTA; id 1..15 (0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

}  // namespace tc
}  // namespace nadm
