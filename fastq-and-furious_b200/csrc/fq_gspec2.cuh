// fq_gspec2.cuh -- the general path's speculative single pass, one WARP per chunk.
//
// Same formulation and the same verification as fq_gspec.cuh (candidates = '@'-class lines, every candidate makes its
// entrypos call against a window of lines, the chain is speculated per chunk and verified by continuity: the entry a
// chunk assumed must be the exit of the chunk before it; anything the window cannot decide declines the whole buffer to
// the exact resolution).  What differs is the shape of the work: fq_gspec.cuh gives a chunk to a CTA of eight warps that
// meet at five barriers per chunk (43 % of its stall samples are barrier waits, one warp walks while seven wait); here
// every warp owns a chunk from the ticket to the rows, nothing but __syncwarp in the loop:
//
//   * the WINDOW of a warp: the last G2_LBL lines of the tile before its own tiles (look-behind), the own tiles, the
//     first G2_LAL lines of the tile behind them (look-ahead) -- the raw 16-bit list entries (offset in tile << 2 |
//     class), copied, not unpacked: 2 bytes per line, 2 048 lines per warp, 32 warps per SM.  A line's position is
//     rebuilt from the tile it belongs to (first line of every tile + the tile of every 32nd line);
//   * one pass over the window builds a bit mask of the '@'-class lines and one of the '+'-class lines (ballots), the
//     running candidate count per 32 lines and the candidate list in line order.  "First '+' at or behind line x" and
//     "first '@' at or behind line x" are then a masked word + ffs instead of a linear scan, and "candidate index of
//     line s" a popcount;
//   * a lane per candidate makes the call (src/_fastqandfurious.c:57-136): header end = next line, '+' line from the
//     mask, and the successor GUESSED from the number of sequence lines: with nseq = k - i - 1 the quality should end
//     on line 2k - i; when that line's newline sits exactly at pos5 and an '@' follows, it is the successor (no earlier
//     line can qualify: the only one at or behind pos5 - 1 would be followed by the newline at pos5).  Otherwise a
//     bounded linear search.  Six shared loads for a clean record, whatever its number of lines;
//   * the chain through 32 consecutive candidates is resolved by POINTER DOUBLING over the lanes: M = set of nodes
//     visited (a 32-bit mask), J = where the chain leaves the group; five rounds of two shuffles.  The group at the
//     last 16 look-behind candidates speculates the chunk's entry: the first start whose call is COMPLETE, that one of
//     the three candidates before it points to, and whose chain reaches the own lines;
//   * the on-chain lanes of a group rebuild their four positions from the window and PARK the row (16 bytes) in the
//     free upper half of the own tiles' list slots; the chunk publishes its exit and its record count;
//   * record counts: decoupled look-back in TWO levels (chunks inside a block of 32, blocks: with 4 736 chunks in
//     flight a one-level look-back would walk back ~150 rounds at the start), DEFERRED by one chunk: a warp resolves
//     its next chunk before it counts the records in front of the previous one and turns the parked rows into table
//     rows, so the counts it needs have been published long ago (a warp that waited right after its own chunk spent
//     a fifth of the kernel's instructions polling).  The continuity check is distributed: chunk c reads the exit
//     chunk c - 1 published with its count.
// Sequential model with the same decline rules: tests/algo_model.py:model_general_spec2.
#pragma once
#include "fq_gspec.cuh"

namespace fqb {

#ifndef G2_FILL
#define G2_FILL 85  // per cent of a window the staged lines should fill on average (tuning)
#endif
constexpr int G2_WL = 2048;          // lines a warp's window can hold
constexpr int G2_WORDS = G2_WL / 32;
constexpr int G2_CW = 448;           // candidates a chunk can hold (more: declined); 4-line records: one per four lines
constexpr int G2_LBL = 160;          // look-behind, lines
constexpr int G2_LAL = 128;          // look-ahead, lines
constexpr int G2_RUNUP = 16;         // look-behind candidates the entry is speculated from
constexpr int G2_MAXG = 24;          // groups of 32 candidates the chain of one chunk may pass through
constexpr int G2_TC = 8;             // tiles per chunk at most
constexpr int G2_NT = G2_TC + 2;
constexpr int G2_WARPS = 8;
constexpr int G2_THREADS = G2_WARPS * 32;
// successor codes (16 bits): candidate index (< G2_CW) | G2_LINE + window line behind the candidates | how the chain ends
constexpr unsigned int G2_LINE = 0x4000u, G2_NONE_E = 0xFFFEu, G2_NONE_T = 0xFFFDu, G2_UNRES = 0xFFFCu;

struct __align__(16) G2Warp {
    unsigned short lines[G2_WL + 8];     // raw list entries of the window, in line order
    unsigned int atm[G2_WORDS + 2];      // bit x & 31 of word x >> 5: line x is '@'-class
    unsigned int plm[G2_WORDS + 2];      // ... '+'-class
    unsigned short atb[G2_WORDS + 2];    // candidates before line 32 w
    unsigned short tstart[G2_NT + 2];    // window index of the first line of every staged tile; [nt] = lines in the window
    unsigned short cand[G2_CW];          // candidate lines, ascending
    unsigned short csucc[G2_CW];         // successor code of every candidate
    unsigned short tv_end[G2_NT + 2];    // staging: 16-byte vectors of the tiles' lists, running total ...
    unsigned short t_raw0[G2_NT + 2];    // ... first raw entry wanted, number of entries, window index of the first one
    unsigned short t_rawn[G2_NT + 2];
    unsigned short t_d0[G2_NT + 2];
    unsigned char tq[G2_WORDS + 8];      // tile (window relative) of line 32 w
};
constexpr size_t G2_SMEM = sizeof(G2Warp) * G2_WARPS;

// Tiles per chunk from the average number of lines per tile: look-behind + own + look-ahead lines should fill about
// 85 % of a window.
__host__ __device__ __forceinline__ int g2_tiles_per_chunk(unsigned long long n_lines, long long n_tiles)
{
    if (n_tiles <= 0) return 1;
    const unsigned long long per_tile = n_lines / (unsigned long long)n_tiles + 1;
    long long tc = (long long)((unsigned long long)(G2_WL * G2_FILL / 100 - G2_LBL - G2_LAL) / per_tile);
    if (tc > G2_TC) tc = G2_TC;
    if (tc < 1) tc = 1;
    return int(tc);
}

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// arrival counter: everything the thread wrote before is visible to whoever reads the count (release), and what the
// earlier arrivals published is visible to the thread after it (acquire) -- one instruction instead of fence, add, fence
__device__ __forceinline__ unsigned int atom_add_acq_rel_gpu(unsigned int* p, unsigned int v)
{
    unsigned int old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void atom_max_release_gpu(unsigned long long* p, unsigned long long v)
{
    asm volatile("red.release.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct G2Win {
    int nw, nwords;
    bool at_end;      // the window reaches the end of the buffer: what it does not show does not exist
    bool virt0;       // window line 0 is the virtual sentinel
    long long l_rel;  // blob length in window coordinates
    int tile, tshift, mis;  // bytes per tile (a power of two) and its logarithm
    int nt;           // staged tiles
};

// tile (window relative) of line x
__device__ __forceinline__ int g2_tile_of(const G2Warp& w, int x)
{
    int q = w.tq[x >> 5];
    while (x >= int(w.tstart[q + 1])) ++q;  // tstart[nt] = nw > x
    return q;
}
// position of line x (in tile q) in window coordinates: byte index from the window's first tile + 1
__device__ __forceinline__ int g2_rel(const G2Warp& w, const G2Win& g, int x, int q)
{
    const int r = q * g.tile + int(w.lines[x] >> 2) + 1;
    return (g.virt0 && x == 0) ? g.mis : r;
}
// first set bit of the mask at or behind line x (nw: none)
__device__ __forceinline__ int g2_next_bit(const unsigned int* mask, int x, int nw, int nwords)
{
    int wd = x >> 5;
    if (wd >= nwords) return nw;
    unsigned int m = mask[wd] & (0xffffffffu << (x & 31));
    while (m == 0u && ++wd < nwords) m = mask[wd];
    return m ? wd * 32 + __ffs(m) - 1 : nw;
}

// One entrypos call anchored on window line i ('@'-class).  Returns the status (GS_ST_UNRES: the window cannot tell);
// *succ = window line of the next call's "\n@" / G2_NONE_E (COMPLETE, no further "\n@") / G2_NONE_T (not COMPLETE: the
// chain stops on this node) / G2_UNRES; *k_out = line of the '+'.  FULL: rel[] = the six positions (-1: not set).
template <bool FULL>
__device__ __forceinline__ int g2_call(const G2Warp& w, const G2Win& g, int i, int* k_out, unsigned int* succ, int* rel)
{
    if (FULL) {
#pragma unroll
        for (int q = 0; q < 6; ++q) rel[q] = -1;
    }
    *succ = G2_UNRES;
    *k_out = 0;
    const int nw = g.nw;
    int q = g2_tile_of(w, i);
    const int p0 = g2_rel(w, g, i, q) + 1;
    if (FULL) rel[0] = p0;
    if (i + 1 >= nw) {
        if (!g.at_end) return GS_ST_UNRES;
        *succ = G2_NONE_T;
        return ST_NO_HEAD_END;
    }
    const unsigned int e1 = w.lines[i + 1];
    while (i + 1 >= int(w.tstart[q + 1])) ++q;
    const int p1 = g2_rel(w, g, i + 1, q);
    if (FULL) {
        rel[1] = p1;
        rel[2] = p1 + 1;
    }
    // "\n+" from p2 + 1: a newline AT p2 is skipped (:87-88)
    const int k = g2_next_bit(w.plm, i + 2 + ((e1 & 3u) == CLS_NL ? 1 : 0), nw, g.nwords);
    if (k >= nw) {
        if (!g.at_end) return GS_ST_UNRES;
        *succ = G2_NONE_T;
        return ST_NO_SEQ_END;
    }
    *k_out = k;
    while (k >= int(w.tstart[q + 1])) ++q;
    const int p3 = g2_rel(w, g, k, q);
    if (FULL) rel[3] = p3;
    if ((long long)p3 + 2 >= g.l_rel) {  // (:97-101)
        *succ = G2_NONE_T;
        return ST_NO_QUALHEAD_END;
    }
    if (k + 1 >= nw) {
        if (!g.at_end) return GS_ST_UNRES;
        *succ = G2_NONE_T;
        return ST_NO_QUALHEAD_END;
    }
    while (k + 1 >= int(w.tstart[q + 1])) ++q;
    const int h = g2_rel(w, g, k + 1, q);
    if ((h - p3 - 1) > 1 && (h - p3) != (p1 - p0 + 1)) {  // (:109-117)
        *succ = G2_NONE_T;
        return ST_INVALID;
    }
    const int p4 = h + 1;
    if (FULL) rel[4] = p4;
    const int p5 = p4 + p3 - p1 - 1;  // (:129)
    if ((long long)p5 + 2 >= g.l_rel) {
        *succ = G2_NONE_T;
        return ST_NO_QUAL_END;
    }
    if (FULL) rel[5] = p5;
    // the next call starts at p5 - 1 (src/fastqandfurious.py:254): first '@'-class line whose newline is at or behind it.
    // Guess: as many quality lines as sequence lines.
    int j = 2 * k - i;
    bool hit = false;
    if (j < nw) {
        int qj = q;
        while (j >= int(w.tstart[qj + 1])) ++qj;
        hit = (g2_rel(w, g, j, qj) == p5) && ((w.lines[j] & 3u) == CLS_AT);
    }
    if (!hit) {
        // first line at or behind p5 - 1: the lines are sorted, so it is the lower bound inside the tile that holds
        // that byte (or the first line of the tiles behind it), and never before line k + 2
        const int T = p5 - 1;  // rel = tile * q + offset + 1
        const int qt = (T - 1) >> g.tshift;
        int jl = nw;
        if (qt < g.nt) {
            int lo = int(w.tstart[qt]), hi = int(w.tstart[qt + 1]);
            const int toff = T - 1 - (qt << g.tshift);  // offset inside tile qt
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (int(w.lines[mid] >> 2) >= toff) hi = mid;
                else lo = mid + 1;
            }
            jl = lo;  // the end of the tile: all its lines lie before the target, the next line is the answer
        }
        if (jl < k + 2) jl = k + 2;
        j = g2_next_bit(w.atm, jl, nw, g.nwords);
    }
    if (j >= nw) {
        if (!g.at_end) return GS_ST_UNRES;
        *succ = G2_NONE_E;
        return ST_COMPLETE;
    }
    *succ = (unsigned int)j;
    return ST_COMPLETE;
}

// The chain through the 32 candidates gb .. gb + 31, one per lane: J0 = my successor code.  Afterwards M = the
// candidates the chain from mine visits inside the group (bit = lane), J = the first code it meets outside (a candidate
// behind the group, a look-ahead line, an end).  Successors point forward, so five doublings cover the group.
__device__ __forceinline__ void g2_double(unsigned int gb, unsigned int J0, int lane, unsigned int& M, unsigned int& J)
{
    M = 1u << lane;
    J = J0;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const unsigned int d = J - gb;
        const bool in = d < 32u;
        const unsigned int src = in ? d : (unsigned int)lane;
        const unsigned int Jn = __shfl_sync(0xffffffffu, J, src);
        const unsigned int Mn = __shfl_sync(0xffffffffu, M, src);
        if (in) {
            M |= Mn;
            J = Jn;
        }
    }
}

__device__ __forceinline__ unsigned long long g2_warp_sum(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// '@'-class and '+'-class flags of the eight list entries of one 16-byte vector (bit k = entry k): classes are the low
// two bits of every 16-bit entry ('@' = 01, '+' = 10), two entries per word
__device__ __forceinline__ void g2_class_bits(const uint4& v, unsigned int& at8, unsigned int& pl8)
{
    const unsigned int k1 = 0x00010001u;
    const unsigned int x[4] = {v.x, v.y, v.z, v.w};
    unsigned int a = 0u, b = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const unsigned int hi = x[q] >> 1;
        a |= (x[q] & ~hi & k1) << (2 * q);  // bit 0 and not bit 1
        b |= (hi & ~x[q] & k1) << (2 * q);  // bit 1 and not bit 0
    }
    at8 = (a | (a >> 15)) & 0xffu;  // the upper entry of word q (bit 16 + 2q) belongs at bit 2q + 1
    pl8 = (b | (b >> 15)) & 0xffu;
}

// Where the rows of a chunk wait until the records before the chunk have been counted: the upper half of the list
// slots of its own tiles (a slot holds tile / 8 = 2 048 entries or more, a staged tile at most 1 024 lines), 128 rows
// of four 32-bit window positions (p0 p1 p3 p4) per tile.
constexpr int G2_PARK_ROWS = 128, G2_PARK_AT = 1024;
__device__ __forceinline__ uint4* g2_park(const ListView& lv, int t0, unsigned int r)
{
    unsigned short* base = const_cast<unsigned short*>(lv.lists) +
                           (size_t)(t0 + int(r >> 7)) * (unsigned int)lv.slot_cap + G2_PARK_AT + (r & 127u) * 8u;
    return reinterpret_cast<uint4*>(base);
}

// SpecParams as for fq_gspec_kernel; `pe` is not an array of entries here but scratch for the block level of the
// look-back: [n_blocks + 1] descriptors (state << 62 | records), then [n_blocks + 1] 32-bit arrival counters; `desc`
// and `pe` are zeroed before the launch.
__global__ void __launch_bounds__(G2_THREADS, 4) fq_gspec2_kernel(const SpecParams p)
{
    ParseState* st = p.st;
    if (*((volatile int*)&st->need_general) == 0 || *((volatile int*)&st->error) != 0) return;
    extern __shared__ __align__(16) uint8_t g2_smem[];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31;
    G2Warp& w = reinterpret_cast<G2Warp*>(g2_smem)[threadIdx.x >> 5];
    ListView lv = p.lv;
    lv.cls0 = *((volatile unsigned int*)&st->cls0);
    const long long L = (p.A > 0 ? p.A - p.mis : 0) + p.sentinel;
    const unsigned long long VMASK = (1ull << 62) - 1ull;
    const int tc = g2_tiles_per_chunk(*((volatile unsigned long long*)&st->n_lines), lv.n_tiles);
    const int n_chunks = (lv.n_tiles + tc - 1) / tc;  // <= p.n_chunks (sized for one tile per chunk)
    const int n_blocks = (n_chunks + 31) >> 5;
    unsigned long long* bdesc = p.pe;
    unsigned int* bcnt = reinterpret_cast<unsigned int*>(p.pe + n_blocks + 1);
    const unsigned int lt_mask = (1u << lane) - 1u;
    const bool geometry_ok = lv.slot_cap >= G2_PARK_AT + G2_PARK_ROWS * 8 && (lv.tile & (lv.tile - 1)) == 0;

    // The chunk whose rows this warp still owes: they wait in the parking area while the warp resolves its NEXT chunk,
    // so that the look-back below finds the counts of the chunks before it published long ago instead of spinning
    // (a warp that waited right after its own chunk spent a fifth of the kernel's instructions polling).
    int pd_c = -1, pd_n = 0, pd_t0 = 0;
    long long pd_ob = 0;
    unsigned long long pd_pe = 0;
    auto flush_pending = [&]() {
        if (pd_c < 0) return;  // uniform
        const int c = pd_c, b = c >> 5, b_first = b << 5;
        unsigned long long base = 0;
        if (c > 0) {
            {   // chunks of my block
                const int j = c - 1 - lane;
                unsigned long long d = 0;
                if (j >= b_first) {
                    do {
                        d = ld_acquire_gpu(&p.desc[j]);
                    } while ((d >> 62) == 0);
                    d &= VMASK;
                }
                base = g2_warp_sum(d);
            }
            for (int bj0 = b - 1; bj0 >= 0; bj0 -= 32) {  // whole blocks
                const int bj = bj0 - lane;
                unsigned long long d = 3ull << 62;  // lanes before block 0: neutral
                if (bj >= 0) {
                    do {
                        d = ld_acquire_gpu(&bdesc[bj]);
                    } while ((d >> 62) == 0);
                }
                const unsigned int has_prefix = __ballot_sync(0xffffffffu, (d >> 62) == 2);
                const int stop = has_prefix ? __ffs(has_prefix) - 1 : 32;  // nearest block with an inclusive prefix
                base += g2_warp_sum((lane <= stop && bj >= 0) ? (d & VMASK) : 0ull);
                if (has_prefix) break;
            }
            // continuity: the entry this chunk assumed is the exit of the chunk before it
            if (lane == 0) {
                while ((ld_acquire_gpu(&p.desc[c - 1]) >> 62) == 0) {
                }
                const unsigned long long xp = *((volatile unsigned long long*)&p.xx[c - 1]);
                if (xp != pd_pe || pd_pe >= GX_FAIL) st->spec_fail = 1;
            }
        }
        if (lane == 0) {
            if ((c & 31) == 31)  // the last chunk of a block: the block's inclusive prefix
                atom_max_release_gpu(&bdesc[b], (2ull << 62) | (base + (unsigned long long)pd_n));
            if (c == n_chunks - 1) st->n_chain = base + (unsigned long long)pd_n;
        }
        for (int r = lane; r < pd_n; r += 32) {
            const unsigned long long k = base + (unsigned long long)r;
            const uint4 q = *g2_park(lv, pd_t0, (unsigned int)r);  // p0 p1 p3 p4, window coordinates
            if ((long long)k < p.cap) {
                longlong2* row = reinterpret_cast<longlong2*>(p.table + k * 6);
                const long long p1 = (long long)q.y, p3 = (long long)q.z, p4 = (long long)q.w;
                row[0] = make_longlong2(pd_ob + (long long)q.x, pd_ob + p1);
                row[1] = make_longlong2(pd_ob + p1 + 1, pd_ob + p3);
                row[2] = make_longlong2(pd_ob + p4, pd_ob + p4 + p3 - p1 - 1);  // pos5 (:129)
            }
        }
        pd_c = -1;
    };

    for (;;) {
        int c = 0;
        if (lane == 0) c = int(atomicAdd(&st->spec_ticket, 1u));
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= n_chunks) break;
        // a chunk that has been handed out is always published (the chunks behind it wait for its count), resolved or not
        bool fail = !geometry_ok || *((volatile int*)&st->spec_fail) != 0;  // declined already: nothing left to resolve
        const int t0 = c * tc;
        const int t1 = (t0 + tc < lv.n_tiles) ? t0 + tc : lv.n_tiles;
        const int tb = c > 0 ? t0 - 1 : t0;
        const int te = (t1 + 1 < lv.n_tiles) ? t1 + 1 : lv.n_tiles;
        const int nt = te - tb;  // <= G2_NT
        // ---- the window: which lines of which tiles ----
        unsigned int a0 = 0, sc = 0;  // first staged (augmented) entry of my tile, staged entries
        bool trunc = false, big = false;
        if (lane < nt) {
            const unsigned int cf = lv_count(lv, tb + lane);
            sc = cf;
            if (c > 0 && lane == 0 && cf > (unsigned int)G2_LBL) {
                a0 = cf - (unsigned int)G2_LBL;
                sc = (unsigned int)G2_LBL;
            }
            if (tb + lane >= t1 && cf > (unsigned int)G2_LAL) {
                sc = (unsigned int)G2_LAL;
                trunc = true;
            }
            big = tb + lane >= t0 && tb + lane < t1 && cf > (unsigned int)G2_PARK_AT;  // its slot has no room to park rows
        }
        unsigned int inc = sc;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        const unsigned int off = inc - sc;  // window index of my tile's first line (lane nt: lines of the window)
        const int nw = int(__shfl_sync(0xffffffffu, inc, 15));
        const bool at_end = (te == lv.n_tiles) && !__any_sync(0xffffffffu, trunc);
        const unsigned int a0_0 = __shfl_sync(0xffffffffu, a0, 0);
        const int nb = c > 0 ? int(__shfl_sync(0xffffffffu, sc, 0)) : 0;  // look-behind lines
        const int nbo = int(__shfl_sync(0xffffffffu, off, t1 - tb));       // look-behind + own lines
        const unsigned long long R0 = lv_base(lv, tb) + a0_0;              // global rank of window line 0
        if (nw > G2_WL || __any_sync(0xffffffffu, big)) fail = true;
        G2Win g;
        g.nw = nw;
        g.nwords = (nw + 31) >> 5;
        g.at_end = at_end;
        g.virt0 = (tb == 0 && lv.virt && a0_0 == 0u);
        // blob position = rel + bias, rel = (byte index from `base`) - tb * tile + 1
        const long long bias = (long long)tb * lv.tile - 1 - p.mis + p.sentinel;
        g.l_rel = L - bias;
        g.tile = lv.tile;
        g.tshift = __ffs(lv.tile) - 1;
        g.mis = lv.mis;
        g.nt = nt;
        const int nwords = g.nwords;
        __syncwarp();  // the previous chunk's window is no longer read
        if (lane <= nt) w.tstart[lane] = (unsigned short)off;
        if (lane == nt + 1) w.tstart[lane] = 0xFFFFu;
        if (!fail)
            for (int wd = lane; wd <= nwords; wd += 32) {  // the masks are assembled with atomicOr
                w.atm[wd] = 0u;
                w.plm[wd] = 0u;
            }

        // ---- A. the raw list entries: the 16-byte vectors of all staged tiles as ONE flat list over the lanes, four
        //      independent loads per lane in flight; a lane copies the entries it wants and ORs their class flags into
        //      the window's '@' and '+' masks ----
        unsigned int V = 0;
        {
            unsigned int raw0 = a0, rawn = (lane < nt) ? sc : 0u, d0 = off;
            if (lane < nt && tb + lane == 0 && lv.virt) {  // the virtual sentinel leads tile 0's augmented list
                if (a0 == 0u) {
                    d0 += 1u;
                    rawn = sc > 0u ? sc - 1u : 0u;
                } else {
                    raw0 = a0 - 1u;
                }
            }
            const unsigned int nv = rawn ? ((raw0 + rawn + 7u) >> 3) - (raw0 >> 3) : 0u;
            unsigned int ve = nv;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const unsigned int v = __shfl_up_sync(0xffffffffu, ve, o);
                if (lane >= o) ve += v;
            }
            V = __shfl_sync(0xffffffffu, ve, 15);
            if (lane < G2_NT + 2) {
                w.tv_end[lane] = (unsigned short)(lane < nt ? ve : 0xFFFFu);
                w.t_raw0[lane] = (unsigned short)raw0;
                w.t_rawn[lane] = (unsigned short)rawn;
                w.t_d0[lane] = (unsigned short)d0;
            }
        }
        __syncwarp();
        if (!fail) {
            if (g.virt0 && lane == 0) {  // window line 0: class of the buffer's first byte (its position: g2_rel)
                w.lines[0] = (unsigned short)lv.cls0;
                if (lv.cls0 == CLS_AT) atomicOr(&w.atm[0], 1u);
                if (lv.cls0 == CLS_PLUS) atomicOr(&w.plm[0], 1u);
            }
            int qq = 0;  // my cursor over the tiles (my vectors ascend)
            for (unsigned int f0 = 0; f0 < V; f0 += 128u) {
                uint4 v[4];
                int dbase[4], klo[4], khi[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned int f = f0 + 32u * u + (unsigned int)lane;
                    klo[u] = 0;
                    khi[u] = 0;
                    dbase[u] = 0;
                    v[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (f < V) {
                        while (f >= (unsigned int)w.tv_end[qq]) ++qq;  // tv_end of the last staged tile = V > f
                        const unsigned int beg = qq ? (unsigned int)w.tv_end[qq - 1] : 0u;
                        const int raw0 = int(w.t_raw0[qq]), rawn = int(w.t_rawn[qq]);
                        const int e0 = int(((unsigned int)raw0 >> 3) + (f - beg)) * 8;  // first entry of my vector
                        v[u] = __ldg(reinterpret_cast<const uint4*>(lv.lists + (size_t)(tb + qq) * (unsigned int)lv.slot_cap + e0));
                        dbase[u] = int(w.t_d0[qq]) + e0 - raw0;
                        klo[u] = raw0 - e0;          // entries k of the vector with klo <= k < khi are wanted
                        khi[u] = raw0 + rawn - e0;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (khi[u] <= 0) continue;
                    const unsigned int ee[8] = {v[u].x & 0xffffu, v[u].x >> 16, v[u].y & 0xffffu, v[u].y >> 16,
                                                v[u].z & 0xffffu, v[u].z >> 16, v[u].w & 0xffffu, v[u].w >> 16};
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k >= klo[u] && k < khi[u]) w.lines[dbase[u] + k] = (unsigned short)ee[k];
                    unsigned int at8, pl8;
                    g2_class_bits(v[u], at8, pl8);
                    unsigned int vm = khi[u] >= 8 ? 0xffu : ((1u << khi[u]) - 1u);
                    if (klo[u] > 0) vm &= ~((1u << klo[u]) - 1u);  // (klo < 8: the vector holds a wanted entry)
                    at8 &= vm;
                    pl8 &= vm;
                    // bit k belongs to window line dbase + k; dbase >= -7, the unwanted low bits are zero
                    const int pos = dbase[u] + 32, sh = pos & 31, w0 = (pos >> 5) - 1;
                    const unsigned int alo = at8 << sh, plo = pl8 << sh;
                    if (alo) atomicOr(&w.atm[w0], alo);
                    if (plo) atomicOr(&w.plm[w0], plo);
                    if (sh > 24) {
                        const unsigned int ahi = at8 >> (32 - sh), phi = pl8 >> (32 - sh);
                        if (ahi) atomicOr(&w.atm[w0 + 1], ahi);
                        if (phi) atomicOr(&w.plm[w0 + 1], phi);
                    }
                }
            }
        }
        __syncwarp();

        // ---- B. candidates before every 32 lines, the candidates in line order, the tile of every 32nd line ----
        int nc = 0;
        if (!fail) {
            for (int w0 = 0; w0 <= nwords; w0 += 32) {  // (word nwords: the total, an empty mask)
                const int wd = w0 + lane;
                const int lim = nbo - wd * 32;  // the candidates: '@'-class lines below the look-ahead
                unsigned int cm = (wd < nwords && lim > 0) ? w.atm[wd] : 0u;
                if (lim < 32 && lim > 0) cm &= (1u << lim) - 1u;
                const int cnt = __popc(cm);
                int ic = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, ic, o);
                    if (lane >= o) ic += v;
                }
                int idx = nc + ic - cnt;
                if (wd <= nwords) {
                    w.atb[wd] = (unsigned short)idx;
                    if (wd < nwords) {
                        int q = 0;
                        while (wd * 32 >= int(w.tstart[q + 1])) ++q;  // wd * 32 < nw = tstart[nt]
                        w.tq[wd] = (unsigned char)q;
                    }
                }
                while (cm) {
                    const int bpos = __ffs(cm) - 1;
                    cm &= cm - 1u;
                    if (idx < G2_CW) w.cand[idx] = (unsigned short)(wd * 32 + bpos);
                    ++idx;
                }
                nc += __shfl_sync(0xffffffffu, ic, 31);
            }
            if (nc > G2_CW) fail = true;
        }
        __syncwarp();

        // ---- C. every candidate makes its call ----
        int q_own = 0;  // candidates among the look-behind lines
        if (!fail) {
            for (int gb = 0; gb < nc; gb += 32) {
                const int q = gb + lane;
                if (q < nc) {
                    int k;
                    unsigned int s;
                    g2_call<false>(w, g, int(w.cand[q]), &k, &s, nullptr);
                    if (s < G2_LINE) {  // a window line: a candidate below the look-ahead, a plain line inside it
                        if (int(s) < nbo)
                            s = (unsigned int)w.atb[s >> 5] + (unsigned int)__popc(w.atm[s >> 5] & ((1u << (s & 31u)) - 1u));
                        else
                            s |= G2_LINE;
                    }
                    w.csucc[q] = (unsigned short)s;
                }
            }
            if (c > 0) q_own = int(w.atb[nb >> 5]) + __popc(w.atm[nb >> 5] & ((1u << (nb & 31)) - 1u));
        }
        __syncwarp();

        // ---- D. the chain, 32 candidates at a time; the rows of its nodes are parked ----
        int n = 0, term_line = -1;
        unsigned long long x = GX_FAIL, pe_rank = GX_FAIL;
        if (!fail) {
            unsigned int cur = G2_UNRES;  // where the chain is: a candidate (< G2_LINE), a look-ahead line, an end
            int ng = 0;
            const int park_cap = (t1 - t0) * G2_PARK_ROWS;
            // the chain from candidate `cur` through the group at gb: Mc = its nodes, Jc = where it leaves
            auto take_group = [&](unsigned int gb, unsigned int Mc, unsigned int Jc, unsigned int J0) {
                const unsigned int unresm = __ballot_sync(0xffffffffu, J0 == G2_UNRES);
                const unsigned int endtm = __ballot_sync(0xffffffffu, J0 == G2_NONE_T);
                const unsigned int rows = Mc & ~endtm;  // a node the chain stops ON is not a row
                if ((Mc & unresm) || ++ng > G2_MAXG || n + __popc(rows) > park_cap) {
                    fail = true;
                    return;
                }
                if (Mc & endtm) term_line = int(w.cand[gb + 31u - (unsigned int)__clz(Mc)]);
                if ((rows >> lane) & 1u) {  // my candidate is a record of the chain: its positions, from the window
                    const int i = int(w.cand[gb + (unsigned int)lane]);
                    // (the line of the '+' as in the call: "\n+" from p2 + 1)
                    const int k = g2_next_bit(w.plm, i + 2 + ((w.lines[i + 1] & 3u) == CLS_NL ? 1 : 0), nw, nwords);
                    int tq = g2_tile_of(w, i);
                    const int p0 = g2_rel(w, g, i, tq) + 1;
                    while (i + 1 >= int(w.tstart[tq + 1])) ++tq;
                    const int p1 = g2_rel(w, g, i + 1, tq);
                    while (k >= int(w.tstart[tq + 1])) ++tq;
                    const int p3 = g2_rel(w, g, k, tq);
                    while (k + 1 >= int(w.tstart[tq + 1])) ++tq;
                    const int p4 = g2_rel(w, g, k + 1, tq) + 1;
                    *g2_park(lv, t0, (unsigned int)(n + __popc(rows & lt_mask))) =
                        make_uint4((unsigned int)p0, (unsigned int)p1, (unsigned int)p3, (unsigned int)p4);
                }
                n += __popc(rows);
                cur = Jc;
            };
            if (c == 0) {
                if (nc > 0) {
                    cur = 0u;  // the head of the whole chain: the first "\n@" of the buffer
                    pe_rank = R0 + w.cand[0];
                } else {  // no candidate among the own lines: the head may still lie in the look-ahead lines
                    const int j = g2_next_bit(w.atm, nbo, nw, nwords);
                    if (j < nw) {
                        cur = G2_LINE | (unsigned int)j;
                        pe_rank = R0 + (unsigned long long)j;
                    } else if (at_end) {
                        cur = G2_NONE_E;  // no "\n@" at all: an empty chain
                        pe_rank = GX_NONE_E;
                    } else {
                        fail = true;
                    }
                }
            } else {
                // the entry, speculated from the last look-behind candidates
                const int g0 = q_own > G2_RUNUP ? q_own - G2_RUNUP : 0;
                const int nlb = q_own - g0;
                if (nlb == 0) {
                    fail = true;
                } else {
                    const int q = g0 + lane;
                    const unsigned int J0 = (q < nc) ? (unsigned int)w.csucc[q] : G2_UNRES;
                    unsigned int M, J;
                    g2_double((unsigned int)g0, J0, lane, M, J);
                    bool viable = lane < nlb && J0 != G2_NONE_T && J0 != G2_UNRES;
                    // a true record start is nearly always pointed to by the record before it, a quality line that
                    // begins with '@' hardly ever (the first three candidates and the last start are exempt)
                    if (viable && q >= 3 && lane + 1 < nlb)
                        viable = (w.csucc[q - 1] == q) || (w.csucc[q - 2] == q) || (w.csucc[q - 3] == q);
                    const unsigned int ownb = M & ~((1u << nlb) - 1u);  // nodes among the own candidates
                    const unsigned int E = ownb ? (unsigned int)(g0 + __ffs(ownb) - 1) : J;
                    const unsigned int gm = __ballot_sync(0xffffffffu, viable && E < 0x8000u);
                    if (gm == 0u) {
                        fail = true;
                    } else {
                        cur = __shfl_sync(0xffffffffu, E, __ffs(gm) - 1);
                        pe_rank = R0 + (cur < G2_LINE ? (unsigned long long)w.cand[cur] : (unsigned long long)(cur & 0x3FFFu));
                        if (cur < G2_LINE && cur - (unsigned int)g0 < 32u) {  // the chain through this group is known already
                            const unsigned int Mc = __shfl_sync(0xffffffffu, M, cur - (unsigned int)g0);
                            const unsigned int Jc = __shfl_sync(0xffffffffu, J, cur - (unsigned int)g0);
                            take_group((unsigned int)g0, Mc, Jc, J0);
                        }
                    }
                }
            }
            while (!fail && cur < G2_LINE) {
                const unsigned int gb = cur;
                const unsigned int q = gb + (unsigned int)lane;
                const unsigned int J0 = (q < (unsigned int)nc) ? (unsigned int)w.csucc[q] : G2_UNRES;
                unsigned int M, J;
                g2_double(gb, J0, lane, M, J);
                take_group(gb, __shfl_sync(0xffffffffu, M, 0), __shfl_sync(0xffffffffu, J, 0), J0);
            }
            if (!fail) {
                if (cur == G2_NONE_E) x = GX_NONE_E;
                else if (cur == G2_NONE_T) x = GX_NONE_T;
                else if (cur >= G2_LINE && cur < 0x8000u) x = R0 + (unsigned long long)(cur & 0x3FFFu);
                else fail = true;
            }
        }
        if (fail) {
            n = 0;
            x = GX_FAIL;
            pe_rank = GX_FAIL;
        }

        // ---- publish: exit, count; the block's aggregate by whoever completes the block ----
        {
            const int b = c >> 5, b_first = b << 5;
            const int b_n = (n_chunks - b_first < 32) ? n_chunks - b_first : 32;
            unsigned int arrived = 0;
            if (lane == 0) {
                if (fail) st->spec_fail = 1;
                p.xx[c] = x;
                if (c == n_chunks - 1 && !fail) {  // the end of the chain: the call that is not COMPLETE
                    int rel[6];
                    int status = ST_NO_HEAD_BEG;
                    for (int q = 0; q < 6; ++q) rel[q] = -1;
                    if (term_line >= 0) {
                        int k;
                        unsigned int s;
                        status = g2_call<true>(w, g, term_line, &k, &s, rel);
                    }
                    st->spec_tail_status = status;
                    for (int q = 0; q < 6; ++q) st->spec_tail_pos[q] = rel[q] >= 0 ? (long long)rel[q] + bias : -1;
                }
                st_release_gpu(&p.desc[c], (1ull << 62) | (unsigned long long)n);
                arrived = atom_add_acq_rel_gpu(&bcnt[b], 1u);
            }
            arrived = __shfl_sync(0xffffffffu, arrived, 0);
            if (arrived == (unsigned int)(b_n - 1)) {  // every chunk of the block has published its count
                unsigned long long v = (lane < b_n) ? (ld_acquire_gpu(&p.desc[b_first + lane]) & VMASK) : 0ull;
                v = g2_warp_sum(v);
                if (lane == 0) atom_max_release_gpu(&bdesc[b], (1ull << 62) | v);
            }
        }
        // the rows of the chunk before this one (its predecessors have published by now), then this one waits
        flush_pending();
        pd_c = c;
        pd_n = n;
        pd_t0 = t0;
        pd_ob = bias + p.goff;
        pd_pe = pe_rank;
    }
    flush_pending();

    // ---- last CTA: result header ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->spec_done, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    if (*((volatile int*)&st->spec_fail) != 0) return;  // declined: the exact path runs
    const unsigned long long xl = *((volatile unsigned long long*)&p.xx[n_chunks - 1]);
    if (xl != GX_NONE_T && xl != GX_NONE_E) return;
    const long long n = (long long)(*((volatile unsigned long long*)&st->n_chain));
    long long pos[6];
    for (int q = 0; q < 6; ++q) pos[q] = *((volatile long long*)&st->spec_tail_pos[q]);
    const int status = *((volatile int*)&st->spec_tail_status);
    const unsigned long long fbi = *((volatile unsigned long long*)&st->first_bad_inv);
    const long long first_bad = fbi ? (long long)~fbi : -1;
    int error = FQB_OK;
    long long resume = 0;
    if (n + 1 > p.cap)
        error = FQB_ERR_CAPACITY;
    else if (n >= 1)
        resume = *((volatile long long*)&p.table[(n - 1) * 6 + 5]) - p.goff - 1;
    write_result(p.res, n, resume, status, pos, FQB_PATH_GENERAL, error, 0, (long long)st->n_lines, first_bad);
    p.res->reserved[1] = 1;  // resolved by the speculative pass
    __threadfence();
    st->general_done = 1;
}

}  // namespace fqb
