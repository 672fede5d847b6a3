// fq_gspec.cuh -- the general path, speculative single pass ("G-spec").
//
// The exact general path (fq_general.cuh) reproduces the reference's entrypos chain for ANY input by building a global
// line table and resolving the successor forest hierarchically: 11 kernels, ~4 passes over 38 bytes per line.  Most
// input that needs the general path is not hostile, it just is not 4-line FASTQ: wrapped multi-line records
// (data/test_multiline.fq), or a 4-line file with a few damaged entries.  For those this kernel gets the same table in
// ONE pass over the scan kernel's per-tile newline lists (2 bytes per line, the input is not touched):
//
//   * a CHUNK is up to GS_TC consecutive tiles; a CTA stages the newline lists of the tile before it (look-behind), its
//     own tiles and the tile after it (look-ahead) in shared memory: the WINDOW;
//   * the '@'-class lines of the own tiles (and of the last GS_LB lines before them) are the CANDIDATES, compacted in
//     line order; every candidate makes its entrypos call against the window (src/_fastqandfurious.c:57-136 with the
//     "\n+" / "\n@" searches answered by bounded linear scans over the window's lines) and keeps its four positions
//     and its successor: the line the next call would find (src/fastqandfurious.py:254);
//   * chains started anywhere merge with the true chain within a record or two (a false start lands on "the first
//     '\n@' at or after some position", which is a true record start unless a quality line begins with '@' right
//     there).  So the chain is SPECULATED at two levels: 32 walkers (the lanes of one warp) split the own candidates
//     into consecutive regions; each follows the successors from GS_LBQ candidates before its region to the first node
//     inside it (its entry), then through the region (its nodes, its exit).  The first walker's run-up lies in the
//     look-behind lines: its entry is the speculated entry of the chunk;
//   * VERIFICATION makes it exact: inside the chunk every walker's entry must be the exit of the walker before it;
//     across chunks every chunk publishes its speculated entry pe(c) and its exit x(c) (first chain node behind its
//     own lines).  pe(0) is the head by construction; if x(c-1) == pe(c) for every c, induction over walkers and
//     chunks shows that every node emitted lies on the reference's chain and none is missing.  Any mismatch, any
//     lookup that leaves the window or exceeds its bound, a window that does not fit, a chain that ends before the last
//     chunk: the kernel declines and the exact path runs (ParseState::general_done stays 0) -- results are identical
//     either way, only the time differs;
//   * the record count goes through a decoupled look-back (single-pass prefix sum over chunks) whose wait is deferred
//     by one chunk: the rows of a chunk stay in registers while the CTA resolves its next chunk.
// Sequential model with the same decline rules: tests/algo_model.py:model_general_spec (property-tested against the
// oracle on CPU).
#pragma once
#include "fq_common.cuh"
#include "fq_emit.cuh"

namespace fqb {

constexpr int GS_TC = 8;          // tiles per chunk at most (fewer when the lines are dense, see gs_tiles_per_chunk)
constexpr int GS_W = 4096;        // lines a window can hold (look-behind + own + look-ahead)
constexpr int GS_THREADS = 256;
constexpr int GS_CMAX = 768;      // candidates a chunk can hold (more: declined)
constexpr int GS_SCAN = 192;      // bound of the linear searches inside one call
constexpr int GS_LB = 160;        // look-behind of the chunk, in lines
constexpr int GS_LBQ = 12;        // run-up of a walker, in candidates
constexpr unsigned short GS_UNRES = 0xFFFD, GS_NONE_E = 0xFFFE, GS_NONE_T = 0xFFFF;  // successor codes (window indices < GS_W)
constexpr unsigned short GS_INF = 0xFFFF;
constexpr unsigned long long GX_NONE_T = ~0ull, GX_NONE_E = ~0ull - 1, GX_FAIL = ~0ull - 2;
constexpr int GS_ST_UNRES = 100;
// walker results: a window line (< GS_W), or the line the chain ended on tagged END_E (COMPLETE, no further "\n@") /
// END_T (a call that is not COMPLETE), or UNRES
constexpr unsigned int GW_END_E = 0x10000u, GW_END_T = 0x20000u, GW_UNRES = 0x40000u, GW_NONE = 0x80000u;
// lines, line -> candidate index, candidate lines, successor (candidate index / line), rows (p0 p1 p3 p4)
constexpr size_t GS_SMEM = size_t(GS_W) * (4 + 2) + size_t(GS_CMAX) * (2 + 2 + 2 + 16);

struct SpecParams {
    const uint8_t* base;
    long long A;
    int mis, sentinel;
    long long goff;
    long long* table;
    long long cap;
    ParseState* st;
    fqb_result* res;
    ListView lv;                // cls0 is filled in on the device
    unsigned long long* desc;   // [n_chunks] look-back descriptors: state << 62 | records (zeroed before the launch)
    unsigned long long* pe;     // [n_chunks] speculated entry (global line rank) / GX_FAIL
    unsigned long long* xx;     // [n_chunks] exit: rank of the first chain node behind the chunk / GX_NONE_* / GX_FAIL
    int n_chunks;
};

// Tiles per chunk, from the average number of lines per tile: look-behind + own + look-ahead tiles should fill about
// 70 % of a window.  Every CTA (and the host, for the number of chunks) derives the same value from the line count.
__host__ __device__ __forceinline__ int gs_tiles_per_chunk(unsigned long long n_lines, long long n_tiles)
{
    if (n_tiles <= 0) return 1;
    const unsigned long long per_tile = n_lines / (unsigned long long)n_tiles + 1;
    long long tc = (long long)((unsigned long long)(GS_W * 7 / 10) / per_tile) - 2;
    if (tc > GS_TC) tc = GS_TC;
    if (tc < 1) tc = 1;
    return int(tc);
}

struct SpecWin {
    uint32_t e_s;                // shared-space address of the lines: (rel << 2) | class, rel = byte index from the window's first tile + 1
    int nw;
    bool at_end;                 // the window reaches the last tile: what it does not show does not exist
    long long l_rel;             // blob length in window coordinates
};

// One entrypos call anchored on window line i (class '@').  Returns the status (GS_ST_UNRES: the window cannot tell),
// rel[] = the six positions in window coordinates (-1: not set), *succ = window index of the next call's "\n@" /
// GS_NONE_E (COMPLETE, no further "\n@") / GS_NONE_T (not COMPLETE: the chain stops on this node) / GS_UNRES.
// The searches (the "\n+", then the first "\n@" at or behind the resume position) are linear scans over the window's
// lines, bounded by GS_SCAN steps each.
__device__ __forceinline__ int spec_rec(const SpecWin& w, int i, int* rel, unsigned short* succ)
{
#pragma unroll
    for (int q = 0; q < 6; ++q) rel[q] = -1;
    *succ = GS_UNRES;
    const int nw = w.nw;
    const uint32_t es = w.e_s;
    const int p0 = int(lds_u32(es + 4u * i) >> 2) + 1;
    rel[0] = p0;
    if (i + 1 >= nw) {
        if (!w.at_end) return GS_ST_UNRES;
        *succ = GS_NONE_T;
        return ST_NO_HEAD_END;
    }
    const unsigned int e1 = lds_u32(es + 4u * (i + 1));
    const int p1 = int(e1 >> 2);
    rel[1] = p1;
    rel[2] = p1 + 1;
    int k = i + 2 + ((e1 & 3u) == CLS_NL ? 1 : 0);  // "\n+" from p2 + 1: a newline AT p2 is skipped (:87-88)
    const int kend = (k + GS_SCAN < nw) ? k + GS_SCAN : nw;
    unsigned int ek = 0;
    for (; k < kend; ++k) {
        ek = lds_u32(es + 4u * k);
        if ((ek & 3u) == CLS_PLUS) break;
    }
    if (k >= nw) {
        if (!w.at_end) return GS_ST_UNRES;
        *succ = GS_NONE_T;
        return ST_NO_SEQ_END;
    }
    if (k >= kend) return GS_ST_UNRES;
    const int p3 = int(ek >> 2);
    rel[3] = p3;
    if ((long long)p3 + 2 >= w.l_rel) {  // (:97-101)
        *succ = GS_NONE_T;
        return ST_NO_QUALHEAD_END;
    }
    if (k + 1 >= nw) {
        if (!w.at_end) return GS_ST_UNRES;
        *succ = GS_NONE_T;
        return ST_NO_QUALHEAD_END;
    }
    const int h = int(lds_u32(es + 4u * (k + 1)) >> 2);
    if ((h - p3 - 1) > 1 && (h - p3) != (p1 - p0 + 1)) {  // (:109-117)
        *succ = GS_NONE_T;
        return ST_INVALID;
    }
    const int p4 = h + 1;
    rel[4] = p4;
    const int p5 = p4 + p3 - p1 - 1;  // (:129)
    if ((long long)p5 + 2 >= w.l_rel) {
        *succ = GS_NONE_T;
        return ST_NO_QUAL_END;
    }
    rel[5] = p5;
    // the next call starts at p5 - 1 (src/fastqandfurious.py:254): first '@'-class line whose newline is at or behind it
    const unsigned int tgt = ((unsigned int)(p5 - 1) << 2) | CLS_AT;  // (rel << 2 | class) >= tgt and class == '@'
    int j = k + 2;
    const int jend = (j + GS_SCAN < nw) ? j + GS_SCAN : nw;
    for (; j < jend; ++j) {
        const unsigned int ej = lds_u32(es + 4u * j);
        if (ej >= tgt && (ej & 3u) == CLS_AT) break;
    }
    if (j >= nw) {
        if (!w.at_end) return GS_ST_UNRES;
        *succ = GS_NONE_E;
        return ST_COMPLETE;
    }
    if (j >= jend) return GS_ST_UNRES;
    *succ = (unsigned short)j;
    return ST_COMPLETE;
}

__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// min over the CTA (GS_THREADS threads, `scratch` = 8 words of shared memory; every thread calls)
__device__ __forceinline__ unsigned int gs_block_min(unsigned int v, unsigned int* scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = __reduce_min_sync(0xffffffffu, v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    unsigned int m = 0xffffffffu;
#pragma unroll
    for (int q = 0; q < GS_THREADS / 32; ++q) m = min(m, scratch[q]);
    return m;
}

__global__ void __launch_bounds__(GS_THREADS, 5) fq_gspec_kernel(const SpecParams p)
{
    ParseState* st = p.st;
    if (*((volatile int*)&st->need_general) == 0 || *((volatile int*)&st->error) != 0) return;
    extern __shared__ __align__(16) uint8_t gs_smem[];
    // shared-space addresses (no generic-pointer arithmetic in the loops)
    const uint32_t we_s = smem_u32(gs_smem);                                   // [GS_W] u32 lines
    const uint32_t rows_s = we_s + uint32_t(GS_W) * 4u;                        // [GS_CMAX] uint4 p0 p1 p3 p4
    const uint32_t lq_s = rows_s + uint32_t(GS_CMAX) * 16u;                    // [GS_W] u16 line -> candidate index
    const uint32_t cand_s = lq_s + uint32_t(GS_W) * 2u;                        // [GS_CMAX] u16 candidate lines, ascending
    const uint32_t nq_s = cand_s + uint32_t(GS_CMAX) * 2u;                     // [GS_CMAX] u16 successor as candidate index
    const uint32_t nl_s = nq_s + uint32_t(GS_CMAX) * 2u;                       // [GS_CMAX] u16 successor as line / code
    const uint32_t ord_s = lq_s;  // on-chain candidates of the own lines, in order (the line -> candidate map is dead by then)
    __shared__ unsigned int s_cnt[2][GS_TC + 3], s_off[2][GS_TC + 3];
    __shared__ unsigned long long s_r0[2];
    __shared__ int s_tick[2];
    __shared__ int s_tc[GS_TC + 3], s_tc2[GS_TC + 3];
    __shared__ unsigned int s_scr[8];
    __shared__ int s_n, s_fail, s_term;
    __shared__ unsigned int s_entry, s_exit;
    __shared__ unsigned long long s_base;
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    ListView lv = p.lv;
    lv.cls0 = *((volatile unsigned int*)&st->cls0);
    const long long L = (p.A > 0 ? p.A - p.mis : 0) + p.sentinel;
    const unsigned long long VMASK = (1ull << 62) - 1ull;
    const int tc = gs_tiles_per_chunk(*((volatile unsigned long long*)&st->n_lines), lv.n_tiles);
    const int n_chunks = (lv.n_tiles + tc - 1) / tc;  // <= p.n_chunks (sized for one tile per chunk)

    // records before chunk c: decoupled look-back over the descriptors by one warp; publishes the inclusive prefix
    auto lookback = [&](int c, int n) -> unsigned long long {
        if (tid < 32) {
            unsigned long long base = 0;
            if (c > 0) {
                int j0 = c - 1;
                for (;;) {
                    const int j = j0 - lane;
                    unsigned long long d = 3ull << 62;  // lanes before chunk 0: neutral
                    if (j >= 0) {
                        do {
                            d = ld_relaxed_gpu(&p.desc[j]);
                        } while ((d >> 62) == 0);
                    }
                    const unsigned int has_prefix = __ballot_sync(0xffffffffu, (d >> 62) == 2);
                    const int stop = has_prefix ? __ffs(has_prefix) - 1 : 32;  // nearest chunk with an inclusive prefix
                    unsigned long long part = (lane <= stop && j >= 0) ? (d & VMASK) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    base += part;
                    if (has_prefix || j0 - 32 < 0) break;
                    j0 -= 32;
                }
                if (lane == 0) st_relaxed_gpu(&p.desc[c], (2ull << 62) | (base + (unsigned long long)n));
            }
            if (lane == 0) s_base = base;
        }
        __syncthreads();
        const unsigned long long b = s_base;
        __syncthreads();  // s_base may be rewritten by the next call
        return b;
    };
    auto store_row = [&](unsigned long long k, long long ob, const uint4& r) {  // r = p0 p1 p3 p4, window coordinates
        if ((long long)k < p.cap) {
            longlong2* row = reinterpret_cast<longlong2*>(p.table + k * 6);
            const long long p1 = (long long)r.y, p3 = (long long)r.z, p4 = (long long)r.w;
            row[0] = make_longlong2(ob + (long long)r.x, ob + p1);
            row[1] = make_longlong2(ob + p1 + 1, ob + p3);
            row[2] = make_longlong2(ob + p4, ob + p4 + p3 - p1 - 1);  // pos5 (:129)
        }
    };
    int pd_c = -1, pd_n = 0;  // the chunk whose rows this CTA still holds (one per thread, in registers)
    long long pd_ob = 0;
    uint4 pd_row = make_uint4(0, 0, 0, 0);
    auto flush_pending = [&]() {
        if (pd_c < 0) return;  // uniform
        const unsigned long long base = lookback(pd_c, pd_n);
        if (tid < pd_n) store_row(base + (unsigned long long)tid, pd_ob, pd_row);
        pd_c = -1;
    };
    // the next chunk of this CTA and the line counts of its window's tiles (the last warp, one chunk ahead: the
    // ticket's round trip and the prefix loads overlap the current chunk's work)
    auto fetch_next = [&](int slot) {  // called by the last warp
        int c = 0;
        if (lane == 0) c = int(atomicAdd(&st->spec_ticket, 1u));
        c = __shfl_sync(0xffffffffu, c, 0);
        if (lane == 0) s_tick[slot] = c;
        if (c >= n_chunks) return;
        const int t0 = c * tc;
        const int t1 = (t0 + tc < lv.n_tiles) ? t0 + tc : lv.n_tiles;
        const int tb = c > 0 ? t0 - 1 : t0;
        const int te = (t1 + 1 < lv.n_tiles) ? t1 + 1 : lv.n_tiles;
        const int nt = te - tb;  // <= GS_TC + 2
        const unsigned int cnt = (lane < nt) ? lv_count(lv, tb + lane) : 0u;
        unsigned int inc = cnt;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane <= GS_TC + 2) {
            s_cnt[slot][lane] = cnt;
            s_off[slot][lane] = inc - cnt;  // s_off[nt] = all lines of the window
        }
        if (lane == 0) s_r0[slot] = lv_base(lv, tb);
    };
    if (warp == GS_THREADS / 32 - 1) fetch_next(0);

    for (int it = 0;; ++it) {
        __syncthreads();  // the previous chunk's shared memory is no longer needed; the prefetched chunk is visible
        const int slot = it & 1;
        const int c = s_tick[slot];
        if (c >= n_chunks) break;
        if (tid == 0) {
            s_n = 0;
            s_fail = 0;
            s_term = -1;
            s_entry = GW_NONE;
            s_exit = GW_NONE;
        }
        const int t0 = c * tc;
        const int t1 = (t0 + tc < lv.n_tiles) ? t0 + tc : lv.n_tiles;
        const int tb = c > 0 ? t0 - 1 : t0;
        const int te = (t1 + 1 < lv.n_tiles) ? t1 + 1 : lv.n_tiles;
        const int nt = te - tb;  // <= GS_TC + 2
        const int nb = c > 0 ? int(s_cnt[slot][0]) : 0;   // look-behind lines
        const int nbo = int(s_off[slot][t1 - tb]);         // look-behind + own lines
        const int nw = int(s_off[slot][nt]);
        const int clo = (nb > GS_LB) ? nb - GS_LB : 0;  // first line whose call is needed
        bool fail = nw > GS_W;  // uniform
        for (int q = 0; q < nt; ++q)
            if (s_cnt[slot][q] > 1024u) fail = true;  // A keeps one candidate bit per entry for four rounds of 256
        SpecWin w;
        w.e_s = we_s;
        w.nw = nw;
        w.at_end = (te == lv.n_tiles);
        // blob position = rel + bias, rel = (byte index from `base`) - tb * tile + 1
        const long long bias = (long long)tb * lv.tile - 1 - p.mis + p.sentinel;
        w.l_rel = L - bias;
        const unsigned long long R0 = s_r0[slot];  // global rank of window line 0
        int n = 0;                 // rows of this chunk
        unsigned long long x = GX_FAIL, pe_rank = GX_FAIL;
        int term_line = -1;
        if (warp == GS_THREADS / 32 - 1) fetch_next(slot ^ 1);
        unsigned int cmask[2] = {0u, 0u};  // candidate bits of my tiles (q = warp and warp + 8), from A to B
        if (!fail) {
            // ---- A. the window's lines: a warp per tile, a lane per 8 list entries (one 16-byte load); which of a
            //      lane's entries are candidates ('@'-class lines in [clo, nbo)) stays in a register: one bit per entry,
            //      up to four rounds of 256 entries per tile (a denser tile declines) ----
#pragma unroll
            for (int qi = 0; qi < 2; ++qi) {  // (nt <= GS_TC + 2 <= 16 tiles: two per warp at most)
                const int q = warp + qi * (GS_THREADS / 32);
                if (q >= nt) break;
                const int t = tb + q;
                unsigned int cnt = s_cnt[slot][q];
                const unsigned short* src = lv.lists + (size_t)t * (unsigned int)lv.slot_cap;
                unsigned int i0 = s_off[slot][q];  // window index of the tile's first line
                const unsigned int relbase4 = ((unsigned int)q * (unsigned int)lv.tile + 1u) << 2;
                int nc_t = 0;
                if (t == 0 && lv.virt) {  // the virtual sentinel leads tile 0: byte index mis - 1 (tb == 0)
                    if (lane == 0) sts_u32(we_s + 4u * i0, ((unsigned int)lv.mis << 2) | lv.cls0);
                    i0 += 1;
                    cnt -= 1;
                }
                const uint32_t dst_s = we_s + 4u * i0;
                // entries v of this tile with c_lo <= v < c_hi are inside [clo, nbo)
                const unsigned int c_lo = (unsigned int)clo > i0 ? (unsigned int)clo - i0 : 0u;
                const unsigned int c_hi = (unsigned int)nbo > i0 ? (unsigned int)nbo - i0 : 0u;
                unsigned int cm = 0u;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const unsigned int v = (unsigned int)it * 256u + (unsigned int)lane * 8u;
                    if ((unsigned int)it * 256u >= cnt) break;  // uniform
                    if (v < cnt) {
                        const uint4 x4 = *reinterpret_cast<const uint4*>(src + v);  // the slot is 16-byte aligned and padded
                        const unsigned int ee[8] = {x4.x & 0xffffu, x4.x >> 16, x4.y & 0xffffu, x4.y >> 16,
                                                    x4.z & 0xffffu, x4.z >> 16, x4.w & 0xffffu, x4.w >> 16};
                        unsigned int m8 = 0u;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            if (v + k < cnt) {
                                // ((relbase + (e >> 2)) << 2) | (e & 3) = (relbase << 2) + e
                                sts_u32(dst_s + 4u * (v + k), relbase4 + ee[k]);
                                if ((ee[k] & 3u) == CLS_AT) m8 |= 1u << k;
                            }
                        }
                        // keep the candidates inside [c_lo, c_hi): entries v .. v + 7 against the two bounds
                        if (v < c_lo) m8 &= (c_lo - v >= 8u) ? 0u : (0xffu << (c_lo - v));
                        if (v + 8u > c_hi) m8 &= (c_hi <= v) ? 0u : ((1u << (c_hi - v)) - 1u);
                        cm |= m8 << (8 * it);
                    }
                }
                nc_t = __reduce_add_sync(0xffffffffu, __popc(cm));
                if (t == 0 && lv.virt && lv.cls0 == CLS_AT && i0 - 1u >= (unsigned int)clo && i0 - 1u < (unsigned int)nbo) nc_t += 1;
                cmask[qi] = cm;
                if (lane == 0) {
                    s_tc[q] = nc_t;
                    s_tc2[q] = (c > 0 && q == 0) ? nc_t : 0;  // the look-behind lines are exactly the first staged tile's
                }
            }
        }
        __syncthreads();
        int nc = 0, qa = 0;  // all candidates, those before the own lines
        if (!fail) {
            for (int q = 0; q < nt; ++q) {
                nc += s_tc[q];
                qa += s_tc2[q];
            }
            if (nc > GS_CMAX) fail = true;  // uniform
        }
        if (!fail) {
            // ---- B. the candidates, compacted in line order, from the bits kept in A: a warp-wide prefix over the lanes'
            //      counts per round, then every lane stores its (few) candidates ----
            int cbase = 0;
            for (int q = 0; q < warp && q < nt; ++q) cbase += s_tc[q];
#pragma unroll
            for (int qi = 0; qi < 2; ++qi) {
                const int q = warp + qi * (GS_THREADS / 32);
                if (q >= nt) break;
                unsigned int i0 = s_off[slot][q];
                const unsigned int cm = cmask[qi];
                if (tb + q == 0 && lv.virt) {  // the virtual sentinel: the tile's first line, a candidate of its own
                    if (lv.cls0 == CLS_AT && i0 >= (unsigned int)clo && i0 < (unsigned int)nbo) {
                        if (lane == 0) {
                            sts_u16(cand_s + 2u * (unsigned int)cbase, i0);
                            sts_u16(lq_s + 2u * i0, (unsigned int)cbase);
                        }
                        cbase += 1;
                    }
                    i0 += 1;
                }
                if (s_tc[q] > 0) {
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        unsigned int m8 = (cm >> (8 * it)) & 0xffu;
                        if (__ballot_sync(0xffffffffu, m8 != 0u) == 0u) continue;  // uniform
                        const int mine = __popc(m8);
                        int inc = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int nb_ = __shfl_up_sync(0xffffffffu, inc, o);
                            if (lane >= o) inc += nb_;
                        }
                        unsigned int qq = (unsigned int)(cbase + inc - mine);
                        const unsigned int ibase = i0 + (unsigned int)it * 256u + (unsigned int)lane * 8u;
                        while (m8) {
                            const unsigned int k = __ffs(m8) - 1;
                            m8 &= m8 - 1;
                            sts_u16(cand_s + 2u * qq, ibase + k);
                            sts_u16(lq_s + 2u * (ibase + k), qq);
                            ++qq;
                        }
                        cbase += __shfl_sync(0xffffffffu, inc, 31);
                    }
                }
                for (int q2 = q + 1; q2 < q + GS_THREADS / 32 && q2 < nt; ++q2) cbase += s_tc[q2];  // tiles of the other warps
            }
        }
        __syncthreads();
        if (fail) nc = 0;
        // ---- C. every candidate makes its call ----
        for (int q = tid; q < nc; q += GS_THREADS) {
            const int i = int(lds_u16(cand_s + 2u * q));
            int rel[6];
            unsigned short s;
            spec_rec(w, i, rel, &s);
            sts_u16(nl_s + 2u * q, s);
            sts_u16(nq_s + 2u * q, (s < (unsigned short)nbo) ? lds_u16(lq_s + 2u * s) : (unsigned int)GS_INF);  // successors inside [clo, nbo) are candidates
            sts_128(rows_s + 16u * q, make_uint4((unsigned int)rel[0], (unsigned int)rel[1], (unsigned int)rel[3], (unsigned int)rel[4]));
        }
        // chunk 0: the head of the whole chain is the first "\n@" of the window; without a candidate among the
        // own lines it may still lie in the look-ahead tile
        unsigned int head_line = 0xffffffffu;
        if (c == 0 && nc == 0 && !fail) {
            unsigned int mine = 0xffffffffu;
            for (int i = nbo + tid; i < nw; i += GS_THREADS)
                if ((lds_u32(we_s + 4u * i) & 3u) == CLS_AT) {
                    mine = (unsigned int)i;
                    break;
                }
            head_line = gs_block_min(mine, s_scr);
        }
        __syncthreads();  // the line -> candidate map is dead from here on (the ordered list reuses it)
        // ---- D. the chain through the own candidates: 32 walkers (warp 0), one region of candidates each ----
        if (warp == 0 && !fail) {
            const int q_own = (c == 0) ? 0 : qa;           // first candidate of the own lines
            const int cper = (nc - q_own + 31) >> 5;       // candidates per region
            int rlo = q_own + lane * cper, rhi = rlo + cper;
            if (rlo > nc) rlo = nc;
            if (rhi > nc) rhi = nc;
            const bool active = (lane == 0) || rlo < nc;   // walker 0 also stands for "no own candidate at all"
            // follows the successors from candidate q until a candidate >= limit: returns that candidate's line, or
            // the line behind the candidates the chain leaves to, or how it ended
            auto run = [&](int q, int limit, int* q_out) -> unsigned int {
                for (;;) {
                    if (q >= limit) {
                        *q_out = q;
                        return lds_u16(cand_s + 2u * q);
                    }
                    const unsigned int nq = lds_u16(nq_s + 2u * q);
                    if (nq == GS_INF) {
                        const unsigned int sl = lds_u16(nl_s + 2u * q);
                        *q_out = -1;
                        if (sl < GS_UNRES) return sl;  // a line at or behind nbo
                        if (sl == GS_NONE_E) return GW_END_E | lds_u16(cand_s + 2u * q);
                        if (sl == GS_NONE_T) return GW_END_T | lds_u16(cand_s + 2u * q);
                        return GW_UNRES;
                    }
                    q = int(nq);
                }
            };
            unsigned int a = GW_NONE;  // entry: first chain node at or behind my region's first candidate
            int qa_in = -1;            // ... as a candidate index when it is one
            if (active) {
                if (c == 0 && lane == 0) {
                    if (nc > 0) {
                        a = lds_u16(cand_s);
                        qa_in = 0;
                    } else {
                        a = (head_line != 0xffffffffu) ? head_line : (GW_END_E | 0xffffu);  // no "\n@" among the own lines
                    }
                } else {
                    // run-up: chains started on the candidates before the region; the first start whose call is
                    // COMPLETE decides (a false start that ends before the region: the next one)
                    int s0 = rlo - GS_LBQ;
                    if (s0 < 0) s0 = 0;
                    for (; s0 < rlo; ++s0) {
                        const unsigned int sl = lds_u16(nl_s + 2u * s0);
                        if (sl >= GS_UNRES && sl != GS_NONE_E) continue;  // not COMPLETE: no chain from here
                        // prefer a start that one of the three candidates before it points to: a true record start
                        // nearly always is (by the record before it), a quality line that begins with '@' hardly ever
                        if (s0 >= 3 && s0 + 1 < rlo && lds_u16(nq_s + 2u * (s0 - 1)) != (unsigned int)s0 &&
                            lds_u16(nq_s + 2u * (s0 - 2)) != (unsigned int)s0 && lds_u16(nq_s + 2u * (s0 - 3)) != (unsigned int)s0)
                            continue;
                        int qo;
                        const unsigned int r = run(s0, rlo, &qo);
                        a = r;  // reached my region / passed over it / how the chain ended before it
                        if (r < GW_END_E) {
                            qa_in = qo;
                            break;
                        }
                    }
                }
            }
            // my region: the chain's nodes in [rlo, rhi), my exit
            int cnt = 0;
            unsigned int xw = a;  // nothing of mine on the chain: the entry is the exit
            int term = -1;
            // my nodes wait in a scratch area (behind the ordered list, 32 entries per walker: a region holds at most
            // GS_CMAX / 32 = 24 candidates) until the walkers' counts have been summed
            const uint32_t tmp_s = ord_s + 2u * (unsigned int)(GS_CMAX + 32 * lane);
            if (active && qa_in >= 0 && qa_in < rhi) {
                int q = qa_in;
                for (;;) {
                    const unsigned int sl = lds_u16(nl_s + 2u * q);
                    if (sl == GS_UNRES) {
                        xw = GW_UNRES;
                        break;
                    }
                    if (sl == GS_NONE_T) {  // the chain stops ON this node: not a row
                        term = int(lds_u16(cand_s + 2u * q));
                        xw = GW_END_T | (unsigned int)term;
                        break;
                    }
                    sts_u16(tmp_s + 2u * (unsigned int)cnt, (unsigned int)q);
                    ++cnt;
                    if (sl == GS_NONE_E) {
                        xw = GW_END_E | lds_u16(cand_s + 2u * q);
                        break;
                    }
                    const unsigned int nq = lds_u16(nq_s + 2u * q);
                    if (nq == GS_INF) {
                        xw = sl;  // leaves the candidates
                        break;
                    }
                    if (int(nq) >= rhi) {
                        xw = lds_u16(cand_s + 2u * nq);
                        break;
                    }
                    q = int(nq);
                }
            }
            // continuity: my entry is the exit of the active walker before me
            const unsigned int xprev = __shfl_up_sync(0xffffffffu, xw, 1);
            bool bad = active && (a == GW_NONE || a == GW_UNRES || xw == GW_UNRES);
            if (active && lane > 0 && a != xprev) bad = true;
            const unsigned int act = __ballot_sync(0xffffffffu, active);
            const int last_w = 31 - __clz(act);  // act has bit 0
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            const int total = __shfl_sync(0xffffffffu, inc, 31);
            const bool anybad = __any_sync(0xffffffffu, bad);
            if (!anybad && cnt > 0) {  // my nodes into the ordered list (independent copies, no second walk)
                const uint32_t o_s = ord_s + 2u * (unsigned int)(inc - cnt);
                for (int k = 0; k < cnt; ++k) sts_u16(o_s + 2u * k, lds_u16(tmp_s + 2u * k));
            }
            const unsigned int x_last = __shfl_sync(0xffffffffu, xw, last_w);
            const unsigned int a_first = __shfl_sync(0xffffffffu, a, 0);
            if (lane == 0) {
                s_fail = anybad ? 1 : 0;
                s_n = total;
                s_entry = a_first;
                s_exit = x_last;
            }
            // at most one walker when the walk is consistent (the chain stops there); an inconsistent walk is declined
            const unsigned int tmask = __ballot_sync(0xffffffffu, term >= 0);
            if (term >= 0 && lane == 31 - __clz(tmask)) s_term = term;
        }
        __syncthreads();
        if (!fail) {
            if (s_fail) {
                fail = true;
            } else {
                n = s_n;
                const unsigned int en = s_entry, ex = s_exit;
                term_line = s_term;
                if (en < GW_END_E) {
                    pe_rank = R0 + en;
                } else if (c == 0 && en == (GW_END_E | 0xffffu) && w.at_end) {
                    pe_rank = GX_NONE_E;  // no "\n@" at all: an empty chain
                } else {
                    fail = true;  // the chain ended before this chunk (or chunk 0 cannot see its head)
                }
                if (ex < GW_END_E) x = R0 + ex;
                else if (ex & GW_END_E) x = GX_NONE_E;
                else if (ex & GW_END_T) x = GX_NONE_T;
                else fail = true;
            }
        }
        if (fail) {
            n = 0;
            x = GX_FAIL;
            pe_rank = GX_FAIL;
        }
        if (tid == 0) {
            if (fail) st->spec_fail = 1;
            p.pe[c] = pe_rank;
            p.xx[c] = x;
            if (c == n_chunks - 1 && !fail) {  // the end of the chain: the call that is not COMPLETE
                int rel[6];
                int status = ST_NO_HEAD_BEG;
                for (int q = 0; q < 6; ++q) rel[q] = -1;
                if (term_line >= 0) {
                    unsigned short s;
                    status = spec_rec(w, term_line, rel, &s);
                }
                st->spec_tail_status = status;
                for (int q = 0; q < 6; ++q) st->spec_tail_pos[q] = rel[q] >= 0 ? (long long)rel[q] + bias : -1;
            }
        }
        // ---- record count: the aggregate is published at once; the look-back for the exclusive prefix and the row
        //      stores are DEFERRED by one chunk (the rows wait in registers), so that a CTA never sits waiting for the
        //      chunks before it: by the time it has resolved its next chunk they have published long ago ----
        if (tid == 0) st_relaxed_gpu(&p.desc[c], ((c == 0 ? 2ull : 1ull) << 62) | (unsigned long long)n);
        uint4 cur_row = make_uint4(0, 0, 0, 0);
        if (n <= GS_THREADS && tid < n) cur_row = lds_128(rows_s + 16u * lds_u16(ord_s + 2u * tid));
        flush_pending();
        if (n > GS_THREADS) {  // many short records: stored now, straight from shared memory
            const unsigned long long base = lookback(c, n);
            const long long ob = bias + p.goff;
            for (int q = tid; q < n; q += GS_THREADS) store_row(base + (unsigned long long)q, ob, lds_128(rows_s + 16u * lds_u16(ord_s + 2u * q)));
        } else {
            pd_c = c;
            pd_n = n;
            pd_ob = bias + p.goff;
            pd_row = cur_row;
        }
    }
    flush_pending();

    // ---- last CTA: verification + result header ----
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&st->spec_done, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    int bad = (*((volatile int*)&st->spec_fail) != 0) ? 1 : 0;
    const volatile unsigned long long* pe = p.pe;
    const volatile unsigned long long* xx = p.xx;
    for (int c = 1 + tid; c < n_chunks; c += GS_THREADS)
        if (xx[c - 1] != pe[c] || pe[c] >= GX_FAIL) bad = 1;
    bad = __syncthreads_or(bad);
    if (tid != 0) return;
    const unsigned long long xl = xx[n_chunks - 1];
    if (bad || (xl != GX_NONE_T && xl != GX_NONE_E) || pe[0] == GX_FAIL) return;  // declined: the exact path runs
    const long long n = (long long)(*((volatile unsigned long long*)&p.desc[n_chunks - 1]) & VMASK);
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
    st->n_chain = (unsigned long long)n;
    __threadfence();
    st->general_done = 1;
}

}  // namespace fqb
