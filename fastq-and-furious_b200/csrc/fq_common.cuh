// fq_common.cuh -- shared device helpers for the sm_100a FASTQ kernels.
//
// PTX wrappers (mbarrier, 1-D TMA bulk copy, relaxed gpu-scope loads/stores), byte-SIMD newline
// detection and the decoupled look-back used by the single-pass newline-rank scan.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fqb200.h"

namespace fqb {

// ---- status codes of the reference (src/_fastqandfurious.c:7-15, src/fastqandfurious.py:19-27) ----
constexpr int ST_INVALID = FQB_INVALID;
constexpr int ST_NO_HEAD_BEG = FQB_MISSING_SEQHEADER_BEGIN;
constexpr int ST_NO_HEAD_END = FQB_MISSING_SEQHEADER_END;
constexpr int ST_NO_SEQ_END = FQB_MISSING_SEQ_END;
constexpr int ST_NO_QUAL_END = FQB_MISSING_QUAL_END;
constexpr int ST_COMPLETE = FQB_COMPLETE;
constexpr int ST_NO_QUALHEAD_END = FQB_MISSING_QUALHEADER_END;

// class of the byte that follows a newline (2 bits)
constexpr uint32_t CLS_OTHER = 0, CLS_AT = 1, CLS_PLUS = 2, CLS_NL = 3;

__device__ __forceinline__ uint32_t classify(uint8_t b)
{
    return b == '@' ? CLS_AT : (b == '+' ? CLS_PLUS : (b == '\n' ? CLS_NL : CLS_OTHER));
}

// Device-side state shared by the kernels of one parse call (lives at the start of the workspace).
struct ParseState {
    unsigned long long first_bad;  // smallest record index that failed a fast-path check (~0 = none)
    int fast_fail;                 // 1: the 4-line fast path cannot represent this input
    int need_general;              // 1: the general path must (re)compute the result
    int error;                     // FQB_ERR_* raised by a kernel
    unsigned int done_counter;     // CTAs of the finalize kernel that have finished
    unsigned long long n_lines;    // visible newlines (+ sentinel) found by the scan
    // general path
    unsigned int head;             // first candidate line (NONE_T: none)
    unsigned int terminal;         // line on which the chain stopped / NONE_E / NONE_X (no chain)
    unsigned long long n_chain;    // COMPLETE records on the chain
};

// ---- PTX helpers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 1-D TMA bulk copy global -> shared (SASS: UBLKCP), completion on an mbarrier.
// dst/src 16-byte aligned, bytes a non-zero multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- byte SIMD ----------------------------------------------------------------------------------
// bit i of the result is set iff byte i of the 16-byte vector equals '\n'.
__device__ __forceinline__ uint32_t newline_nibble(uint32_t w)
{
    const uint32_t t = w ^ 0x0a0a0a0au;
    const uint32_t m = ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t) & 0x80808080u;  // 0x80 where byte == '\n'
    return (m * 0x00204081u) >> 28;                                               // gather bits 7,15,23,31
}

__device__ __forceinline__ uint32_t newline_mask16(const uint4& v)
{
    return newline_nibble(v.x) | (newline_nibble(v.y) << 4) | (newline_nibble(v.z) << 8) |
           (newline_nibble(v.w) << 12);
}

// ---- decoupled look-back ------------------------------------------------------------------------
// One 64-bit descriptor per tile: bits 63:62 = state, bits 61:0 = count.  A single word keeps state
// and value coherent without fences.
constexpr unsigned long long LB_AGG = 1ull << 62;   // value = this tile's own count
constexpr unsigned long long LB_INCL = 2ull << 62;  // value = inclusive prefix up to this tile
constexpr unsigned long long LB_VALUE = (1ull << 62) - 1;

// Called by ALL 32 lanes of one warp, for tile t >= 1 whose aggregate is already published.
// Returns the exclusive prefix (sum of the counts of tiles 0..t-1).
__device__ __forceinline__ unsigned long long lookback_exclusive(const unsigned long long* desc, long long t,
                                                                 int lane)
{
    unsigned long long excl = 0;
    long long idx = t - 1 - lane;
    for (;;) {
        unsigned long long d = (idx >= 0) ? ld_relaxed_u64(desc + idx) : LB_INCL;  // virtual tile -1: prefix 0
        const unsigned st = static_cast<unsigned>(d >> 62);
        const unsigned incl = __ballot_sync(0xffffffffu, st == 2);
        const unsigned inval = __ballot_sync(0xffffffffu, st == 0);
        const int first_incl = incl ? (__ffs(incl) - 1) : 32;
        const unsigned need = (first_incl >= 31) ? 0xffffffffu : ((2u << first_incl) - 1u);
        if (inval & need) {  // a needed predecessor has not published yet
            __nanosleep(40);
            continue;
        }
        unsigned long long v = (lane <= first_incl) ? (d & LB_VALUE) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (first_incl < 32) break;
        idx -= 32;
    }
    return excl;
}

}  // namespace fqb
