// fq_general_logic.h -- record logic of the GENERAL path, shared by the CUDA kernels and the
// host-side simulation used by the CPU tests (tests/hostsim).  Plain functions over the line table.
//
// Line table: entry i = (blob position of the i-th visible newline << 2) | class of the byte that
// follows it.  "Visible" = every newline of the blob except one in its last byte, which the
// reference can never match (memchr windows exclude it, src/_fastqandfurious.c:71,103; "\n@" /
// "\n+" need a second byte, :62,:88).  With the sentinel, entry 0 is blob position 0.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FQ_HD __host__ __device__ __forceinline__
#else
#define FQ_HD static inline
#endif

namespace fqb {

constexpr unsigned int G_CLS_OTHER = 0, G_CLS_AT = 1, G_CLS_PLUS = 2, G_CLS_NL = 3;

constexpr unsigned int NONE_T = 0xFFFFFFFFu;  // chain stops ON a node whose entrypos call is not COMPLETE
constexpr unsigned int NONE_E = 0xFFFFFFFEu;  // chain stops AFTER a COMPLETE record: no further "\n@"
constexpr unsigned int NONE_X = 0xFFFFFFFDu;  // not a candidate line
constexpr unsigned int NONE_MIN = 0xFFFFFFF0u;  // every value >= this is "outside"

constexpr int G_BLK = 256;                   // lines per summary block (next-'+' / next-'@' lookups)
constexpr int G_GRP = 1024;                  // summary blocks per group (second level of the lookups)
constexpr int G_S1 = 1024;                   // lines per level-1 chunk (resolved in shared memory)
constexpr int G_FAN = 64;                    // fan-out of levels 2 and 3
constexpr long long G_S2 = (long long)G_S1 * G_FAN;  // lines per level-2 block
constexpr long long G_S3 = G_S2 * G_FAN;             // lines per level-3 block

struct LineView {
    const unsigned long long* nlt;  // [M]
    const unsigned int* sumP;       // [nblk] first '+'-class line in blocks b .. end of b's group (NONE_T if none)
    const unsigned int* sumA;       // [nblk] same for '@'
    const unsigned int* gsufP;      // [ngrp + 1] first '+'-class line in groups >= g
    const unsigned int* gsufA;      // [ngrp + 1] same for '@'
    unsigned long long M;           // number of lines
    long long L;                    // blob length
    // optional window of the table held in fast (shared) memory: win[d] == nlt[win_lo + d] for d < win_n
    const unsigned long long* win;
    unsigned long long win_lo;
    unsigned int win_n;
};

FQ_HD unsigned long long line_entry(const LineView& v, unsigned long long i)
{
    const unsigned long long d = i - v.win_lo;  // wraps to a huge value below the window
    return (d < (unsigned long long)v.win_n) ? v.win[d] : v.nlt[i];
}
FQ_HD long long line_pos(const LineView& v, unsigned long long i) { return (long long)(line_entry(v, i) >> 2); }
FQ_HD unsigned int line_cls(const LineView& v, unsigned long long i) { return (unsigned int)(line_entry(v, i) & 3ull); }

// first line index >= i whose class is `cls`, or NONE_T
FQ_HD unsigned int next_of_class(const LineView& v, unsigned long long i, unsigned int cls, const unsigned int* sum,
                                 const unsigned int* gsuf)
{
    if (i >= v.M) return NONE_T;
    const unsigned long long b = i / G_BLK;
    unsigned long long end = (b + 1) * G_BLK;
    if (end > v.M) end = v.M;
    for (unsigned long long j = i; j < end; ++j)
        if (line_cls(v, j) == cls) return (unsigned int)j;
    const unsigned long long b1 = b + 1;
    if (b1 * G_BLK >= v.M) return NONE_T;
    if (b1 % G_GRP) {  // rest of the same group
        const unsigned int s = sum[b1];
        if (s != NONE_T) return s;
        return gsuf[b1 / G_GRP + 1];
    }
    return gsuf[b1 / G_GRP];
}

// One entrypos call anchored on candidate line i (class '@'): src/_fastqandfurious.c:57-136 with
// the memmem / memchr searches answered from the line table.  pos[] is blob relative, -1 filled.
// *succ (when want_succ) = line of the next call's "\n@" (first '@'-class line at a position
// >= pos5 - 1, src/fastqandfurious.py:254), NONE_E if there is none; NONE_T when not COMPLETE.
FQ_HD int general_rec(const LineView& v, unsigned long long i, long long* pos, bool want_succ, unsigned int* succ)
{
    for (int q = 0; q < 6; ++q) pos[q] = -1;
    if (want_succ) *succ = NONE_T;
    const long long p0 = line_pos(v, i) + 1;
    pos[0] = p0;
    if (i + 1 >= v.M) return 1;  // no header '\n' (:70-77)
    const long long p1 = line_pos(v, i + 1);
    pos[1] = p1;
    const long long p2 = p1 + 1;
    pos[2] = p2;
    // "\n+" from p2 + 1 (:87-88): a newline AT p2 (empty first sequence line) is skipped
    const unsigned long long kmin = i + 2 + (line_cls(v, i + 1) == G_CLS_NL ? 1 : 0);
    const unsigned int k = next_of_class(v, kmin, G_CLS_PLUS, v.sumP, v.gsufP);
    if (k == NONE_T) return 3;
    const long long p3 = line_pos(v, k);
    pos[3] = p3;
    if (p3 + 2 >= v.L) return 7;                     // (:97-101)
    if ((unsigned long long)k + 1 >= v.M) return 7;  // no end of the '+' line (:102-107)
    const long long h = line_pos(v, (unsigned long long)k + 1);
    if ((h - p3 - 1) > 1 && (h - p3) != (p1 - p0 + 1)) return -1;  // (:109-117)
    const long long p4 = h + 1;
    pos[4] = p4;
    const long long p5 = p4 + p3 - p1 - 1;  // (:129)
    if (p5 + 2 >= v.L) return 5;
    pos[5] = p5;
    if (want_succ) {
        const long long target = p5 - 1;
        unsigned long long lb = (unsigned long long)k + 2;
        if (lb < v.M && line_pos(v, lb) < target) {
            unsigned long long lo = lb, step = 1;  // line_pos(lo) < target
            while (lo + step < v.M && line_pos(v, lo + step) < target) {
                lo += step;
                step <<= 1;
            }
            unsigned long long hi = lo + step;  // line_pos(hi) >= target or hi >= M
            if (hi > v.M) hi = v.M;
            while (hi - lo > 1) {
                const unsigned long long mid = (lo + hi) >> 1;
                if (line_pos(v, mid) < target)
                    lo = mid;
                else
                    hi = mid;
            }
            lb = hi;
        }
        const unsigned int s = next_of_class(v, lb, G_CLS_AT, v.sumA, v.gsufA);
        *succ = (s == NONE_T) ? NONE_E : s;
    }
    return 6;
}

FQ_HD unsigned long long pack_jump(unsigned int exit, unsigned int hops)
{
    return ((unsigned long long)exit << 32) | hops;
}
FQ_HD unsigned int jump_exit(unsigned long long j) { return (unsigned int)(j >> 32); }
FQ_HD unsigned int jump_hops(unsigned long long j) { return (unsigned int)(j & 0xffffffffull); }

}  // namespace fqb
