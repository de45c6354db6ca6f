// umma.cuh — sm_100a device primitives: mbarrier, TMEM allocation, tcgen05.mma/ld/commit, descriptors.
//
// Everything here is inline PTX for compute_100a. No CUTLASS. Conventions used by every kernel in this
// library:
//
//  * Shared-memory operand tiles use ONE physical layout ("chunk-column" layout, no swizzle):
//        addr(row r, feature f) = (f / CE) * CS + r * 16 + (f % CE) * ES
//    ES = element bytes (2: bf16/fp16, 4: tf32), CE = 16 / ES elements per 16-byte chunk, CS = chunk
//    stride in bytes = (rows incl. halo) * 16.  A 16-byte chunk of 8 consecutive rows is one UMMA
//    "core matrix" (8 x 16 B = 128 B, rows at 16 B stride).
//  * The same tile is a K-major operand (MMA M/N index = row, K = feature):
//        SBO = 128 (next 8-row group), LBO = CS (next 16-byte K chunk);  k-step = 2 chunks = 2*CS
//    and an MN-major operand (MMA M/N index = feature, K = row):
//        SBO = CS (next 8 features), LBO = 128 (next 8 rows);            k-step = 16 rows (bf16) = 256 B
//    Because rows are linear at 16 B stride, a row-shifted view (temporal convolution taps) is the same
//    descriptor with start address + 16*shift.
//  * Accumulators: TMEM lane = tile row (M = 128), column = N index, fp32.
//
// Waits on mbarriers are time-bounded (NBSS_WAIT_TIMEOUT_NS of %globaltimer) and raise a sticky device-side error flag
// instead of hanging the GPU.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nbss {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#ifndef NBSS_WAIT_TIMEOUT_NS
#define NBSS_WAIT_TIMEOUT_NS 200000000ull  // 0.2 s: far beyond any legitimate wait in these kernels
#endif
// Bounded wait. Returns false (and sets *err_flag) when the barrier did not complete within the timeout; once the
// flag is set every later wait in the grid bails out immediately, so a protocol bug cannot hang the GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag = nullptr) {
    if (mbar_try_wait(bar, parity)) return true;
    const uint64_t t0 = global_timer_ns();
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++polls & 255u) == 0) {
            if (err_flag && *reinterpret_cast<volatile int*>(err_flag) != 0) return false;
            if (global_timer_ns() - t0 > NBSS_WAIT_TIMEOUT_NS) {
                if (err_flag) atomicExch(err_flag, 0x7001);
                return false;
            }
        }
    }
    return true;
}

// ---------------------------------------------------------------- fences
__device__ __forceinline__ void fence_async_smem() {  // generic-proxy smem writes -> visible to UMMA/TMA
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---------------------------------------------------------------- TMEM alloc (call from ONE full warp)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (sm_100 format, version 1, no swizzle).
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
    return d;                             // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// K-major view of a chunk-column tile: M/N index = row.
__device__ __forceinline__ uint64_t sdesc_kmajor(uint32_t saddr, uint32_t cs_bytes) {
    return make_sdesc(saddr, /*lbo=*/cs_bytes, /*sbo=*/128);
}
// MN-major view of a chunk-column tile: M/N index = feature, K index = row.
__device__ __forceinline__ uint64_t sdesc_mnmajor(uint32_t saddr, uint32_t cs_bytes) {
    return make_sdesc(saddr, /*lbo=*/128, /*sbo=*/cs_bytes);
}

enum : uint32_t { FMT_F16 = 0, FMT_BF16 = 1, FMT_TF32 = 2 };
// Instruction descriptor for kind::f16 / kind::tf32, fp32 accumulate.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt_ab, uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                  uint32_t b_mn_major) {
    return (1u << 4)                 // D format: F32
           | (fmt_ab << 7)           // A format
           | (fmt_ab << 10)          // B format
           | (a_mn_major << 15)      // A major: 0 = K, 1 = MN
           | (b_mn_major << 16)      // B major
           | ((N >> 3) << 17)        // N / 8
           | ((M >> 4) << 24);       // M / 16
}

// ---------------------------------------------------------------- MMA issue (one thread)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One lane of a fully converged warp (elect.sync).  The MMA-issuing code is written so that the WHOLE warp runs the
// descriptor arithmetic (warp-uniform values -> uniform registers) and only the tcgen05.mma / commit instructions
// themselves are predicated on the elected lane: a single-thread `if (tid == 0)` block makes the compiler stage every
// operand through R2UR.BROADCAST loops (about 20-30 instructions per MMA, more than a small MMA takes to execute).
// The issuing warp must ALSO be selected through a value ptxas knows to be warp-uniform: `warp = __shfl_sync(~0u, tid >> 5, 0)`
// instead of `tid >> 5`.  With the plain shift `if (warp == 0)` counts as a divergent branch, every descriptor lives in vector
// registers and each tcgen05.mma pays five R2UR.BROADCAST moves; with the shuffle the whole issue loop runs on the uniform
// datapath (ffn_fwd: 429 -> 19 R2UR, 14 uniform instructions per MMA).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// A operand from TENSOR MEMORY (lane = row, two 16-bit K elements per 32-bit column, 8 columns per K = 16 step), B from a
// shared-memory descriptor.  Lets a thread-per-row epilogue hand its result (e.g. softmax probabilities) to the next MMA
// without a shared-memory tile.  Convention pinned by tests/test_umma_selftest.py (a_mn = 2).
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// registers -> TMEM, 32x32b: thread i of the warp writes lane (32*(warp%4) + i), 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Arrive on an mbarrier when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- TMEM -> registers
// 32x32b: thread i of the warp reads lane (32*(warp%4) + i), N consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// TMEM address helpers: bits [31:16] lane, [15:0] column.
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) {
    return base + (lane << 16) + col;
}

// ---------------------------------------------------------------- small math / packing helpers
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <int FMT>
__device__ __forceinline__ uint32_t pack16(float lo, float hi) {
    if constexpr (FMT == FMT_F16) return pack_f16(lo, hi);
    else return pack_bf16(lo, hi);
}
__device__ __forceinline__ float bf16lo_to_f32(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }

}  // namespace nbss
