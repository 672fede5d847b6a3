// fq_common.cuh -- shared device helpers for the sm_100a FASTQ kernels.
//
// PTX wrappers (mbarrier, 1-D TMA bulk copy), byte-SIMD newline detection, the per-call device state
// and the view of the per-tile newline lists that the scan kernel leaves for its consumers.
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

// Device-side state shared by the kernels of one parse call (lives at the start of the workspace,
// zeroed by a memset at the start of every call).
struct ParseState {
    unsigned long long first_bad_inv;  // ~(smallest record index that failed a fast-path check); 0 = none
    int fast_fail;                     // 1: the 4-line fast path cannot represent this input
    int need_general;                  // 1: the general path must (re)compute the result
    int error;                         // FQB_ERR_* raised by a kernel
    unsigned int scan_done;            // CTAs of the scan kernel that have finished
    unsigned int emit_done;            // CTAs of the emit kernel that have finished
    unsigned int cls0;                 // class of the buffer's first byte (follows the virtual sentinel)
    unsigned long long n_lines;        // visible newlines (+ sentinel) found by the scan
    // general path
    unsigned int head;                 // first candidate line (NONE_T: none)
    unsigned int terminal;             // line on which the chain stopped / NONE_E / 0 (no chain yet)
    unsigned long long n_chain;        // COMPLETE records on the chain
    unsigned int n_cand;               // candidate lines ('@'-class) collected
    unsigned int n_list1;              // distinct level-1 exits collected
    unsigned int n_list2;              // distinct level-2 exits collected
    unsigned int pad_general;
    // fast path: classification of the open last record, computed while the rows are being written
    long long tail_n;                  // records of the fast-path result
    long long tail_resume;
    long long tail_pos[6];
    int tail_status;
    int tail_error;
    // general path, byte-range sharding: what the previous shard handed over
    unsigned long long shard_resume_abs;      // absolute stream position the chain resumes its search at
    unsigned long long shard_records_before;  // records emitted by the earlier shards
    int shard_ended;                          // 1: the chain ended in an earlier shard (nothing to emit here)
    int shard_pad;
    // general path, speculative single pass (fq_gspec.cuh)
    unsigned int spec_ticket;                 // next chunk to hand out
    unsigned int spec_done;                   // CTAs that have finished
    int spec_fail;                            // 1: a chunk could not be resolved from its window
    int general_done;                         // 1: the speculative pass produced the result; the exact path is skipped
    int spec_tail_status;                     // the call that ends the chain (last chunk)
    int spec_pad;
    long long spec_tail_pos[6];
};

// The scan kernel's output: for every tile of TILE input bytes the list of its visible newlines,
// entry = (offset inside the tile << 2) | class of the following byte, plus newline counts as
// prefixes: CTA b of the scan owns the contiguous tiles [b*T, (b+1)*T); lprefix[t] is the inclusive
// count inside that range and rprefix[b] the number of newlines in all earlier ranges.
// The virtual sentinel newline (blob position 0, rank 0) is not stored: consumers see it as entry 0
// of an "augmented" list of tile 0.
struct ListView {
    const unsigned short* lists;       // [n_tiles][slot_cap]
    const unsigned int* lprefix;       // [n_tiles]
    const unsigned long long* rprefix; // [n_ranges + 1]
    int n_tiles;
    int T;                             // tiles per range
    int slot_cap;                      // entries reserved per tile
    int tile;                          // bytes per tile
    int virt;                          // 1: virtual sentinel present
    int mis;
    unsigned int cls0;
};

__device__ __forceinline__ unsigned int lv_count(const ListView& v, int t)  // augmented
{
    const unsigned int hi = v.lprefix[t];
    const unsigned int lo = ((unsigned int)t % (unsigned int)v.T) ? v.lprefix[t - 1] : 0u;
    return hi - lo + ((t == 0) ? (unsigned int)v.virt : 0u);
}

__device__ __forceinline__ unsigned long long lv_base(const ListView& v, int t)  // rank of augmented entry 0
{
    if (t == 0) return 0ull;
    const unsigned int b = (unsigned int)t / (unsigned int)v.T;
    const unsigned int r = (unsigned int)t - b * (unsigned int)v.T;
    return (unsigned long long)v.virt + v.rprefix[b] + (r ? v.lprefix[t - 1] : 0u);
}

// augmented entry jj of tile t -> position (byte index from `base`) and class
__device__ __forceinline__ void lv_entry(const ListView& v, int t, unsigned int jj, long long* a, unsigned int* cls)
{
    if (t == 0 && v.virt) {
        if (jj == 0) {
            *a = (long long)v.mis - 1;
            *cls = v.cls0;
            return;
        }
        jj -= 1;
    }
    const unsigned int e = v.lists[(size_t)t * (unsigned int)v.slot_cap + jj];
    *a = (long long)t * v.tile + (long long)(e >> 2);
    *cls = e & 3u;
}

// cursor over the global newline sequence
struct LvCursor {
    int t;
    unsigned int jj, n;  // n = augmented count of tile t
};

__device__ __forceinline__ bool lv_next(const ListView& v, LvCursor& c)  // false: ran off the end
{
    if (++c.jj < c.n) return true;
    for (;;) {
        if (++c.t >= v.n_tiles) return false;
        c.n = lv_count(v, c.t);
        if (c.n) {
            c.jj = 0;
            return true;
        }
    }
}

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

// ---- shared memory through 32-bit shared-space addresses (no generic-pointer arithmetic in the loop) ----
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds_128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_128(uint32_t addr, const uint4& v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
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

}  // namespace fqb
