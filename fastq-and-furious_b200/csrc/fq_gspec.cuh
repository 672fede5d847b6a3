// fq_gspec.cuh -- the general path, speculative single pass ("G-spec").
//
// The exact general path (fq_general.cuh) reproduces the reference's entrypos chain for ANY input by building a global
// line table and resolving the successor forest hierarchically: 11 kernels, ~4 passes over 38 bytes per line.  Most
// input that needs the general path is not hostile, it just is not 4-line FASTQ: wrapped multi-line records
// (data/test_multiline.fq), or a 4-line file with a few damaged entries.  For those this kernel gets the same table in
// ONE pass over the scan kernel's per-tile newline lists (2 bytes per line, the input is not touched):
//
//   * a CHUNK is GS_TC consecutive tiles; a CTA stages the newline lists of the tile before it (look-behind), its own
//     tiles and the tile after it (look-ahead) in shared memory: the WINDOW;
//   * every '@'-class line of the own tiles (and of the last GS_LB lines before them) makes its entrypos call against
//     the window (src/_fastqandfurious.c:57-136 with the "\n+" / "\n@" searches answered from next-'+' / next-'@'
//     arrays built by two backward scans, one bounded linear search for the resume position) and gets its successor:
//     the line the next call would find (src/fastqandfurious.py:254);
//   * the ENTRY of the chain into the chunk is SPECULATED: chains started anywhere merge with the true chain within a
//     record or two (a false start lands on "the first '\n@' at or after some position", which is a true record start
//     unless a quality line begins with '@' right there), so the chain started GS_LB lines before the chunk, followed to
//     the first node inside the chunk, is the true entry with overwhelming probability;
//   * the nodes of that chain are found by pointer doubling over the successors (log2 rounds, all candidates in
//     parallel: no serial walk), the record count goes through a decoupled look-back (single-pass prefix sum over
//     chunks), the rows are written in parallel;
//   * VERIFICATION makes it exact: every chunk publishes its speculated entry pe(c) and its exit x(c) (first chain
//     node behind its own lines).  pe(0) is the head by construction; if x(c-1) == pe(c) for every c, induction over c
//     shows that every chunk walked the reference's chain.  Any mismatch, any lookup that leaves the window, a window
//     that does not fit, a chain that ends before the last chunk: the kernel declines and the exact path runs
//     (ParseState::general_done stays 0) -- results are identical either way, only the time differs.
// Sequential model with the same decline rules: tests/algo_model.py:model_general_spec (property-tested against the
// oracle on CPU).
#pragma once
#include "fq_common.cuh"
#include "fq_emit.cuh"

namespace fqb {

constexpr int GS_TC = 8;          // tiles per chunk at most (fewer when the lines are dense, see gs_tiles_per_chunk)
constexpr int GS_W = 4096;        // lines a window can hold (look-behind + own + look-ahead)
constexpr int GS_THREADS = 256;
constexpr int GS_STRIP = GS_W / GS_THREADS;  // consecutive lines per thread in the next-'+' / next-'@' scans
constexpr int GS_SCAN = 192;      // bound of the linear search inside one call
constexpr int GS_LB = 160;        // the speculated entry comes from a chain started this many lines before the chunk
constexpr int GS_CPT = 4;        // candidates per thread the pointer doubling holds in registers (more: declined)
constexpr int GS_STARTS = 8;      // chains tried (a false start may end on INVALID before it reaches the chunk)
constexpr unsigned short GS_UNRES = 0xFFFD, GS_NONE_E = 0xFFFE, GS_NONE_T = 0xFFFF;  // successor codes (window indices < GS_W)
constexpr unsigned short GS_INF = 0xFFFF;
constexpr unsigned long long GX_NONE_T = ~0ull, GX_NONE_E = ~0ull - 1, GX_FAIL = ~0ull - 2;
constexpr int GS_ST_UNRES = 100;
constexpr size_t GS_SMEM = size_t(GS_W) * (4 + 2 + 2 + 2 + 2);  // lines, successors, two scratch arrays, candidates

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
    const unsigned int* e;       // (rel << 2) | class, rel = byte index from the window's first tile + 1
    const unsigned short* nxp;   // first '+'-class line at or after i (GS_INF: none in the window); [nw + 1]
    const unsigned short* nxa;   // ... '@'-class
    int nw;
    bool at_end;                 // the window reaches the last tile: what it does not show does not exist
    long long l_rel;             // blob length in window coordinates
};

// One entrypos call anchored on window line i (class '@').  Returns the status (GS_ST_UNRES: the window cannot tell),
// rel[] = the six positions in window coordinates (-1: not set), *succ = window index of the next call's "\n@" /
// GS_NONE_E (COMPLETE, no further "\n@") / GS_NONE_T (not COMPLETE: the chain stops on this node) / GS_UNRES.
template <bool TABLES>  // TABLES: '+' / '@' searches answered from nxp / nxa; else linear scans (row emission)
__device__ __forceinline__ int spec_rec(const SpecWin& w, int i, int* rel, unsigned short* succ)
{
#pragma unroll
    for (int q = 0; q < 6; ++q) rel[q] = -1;
    *succ = GS_UNRES;
    const int nw = w.nw;
    const int p0 = int(w.e[i] >> 2) + 1;
    rel[0] = p0;
    if (i + 1 >= nw) {
        if (!w.at_end) return GS_ST_UNRES;
        *succ = GS_NONE_T;
        return ST_NO_HEAD_END;
    }
    const unsigned int e1 = w.e[i + 1];
    const int p1 = int(e1 >> 2);
    rel[1] = p1;
    rel[2] = p1 + 1;
    const int kmin = i + 2 + ((e1 & 3u) == CLS_NL ? 1 : 0);  // "\n+" from p2 + 1: a newline AT p2 is skipped (:87-88)
    int k;
    if (TABLES) {
        k = (kmin < nw) ? int(w.nxp[kmin]) : int(GS_INF);
    } else {
        k = kmin;
        while (k < nw && (w.e[k] & 3u) != CLS_PLUS) ++k;
    }
    if (k >= nw) {
        if (!w.at_end) return GS_ST_UNRES;
        *succ = GS_NONE_T;
        return ST_NO_SEQ_END;
    }
    const int p3 = int(w.e[k] >> 2);
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
    const int h = int(w.e[k + 1] >> 2);
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
    const int target = p5 - 1;  // the next call starts here (src/fastqandfurious.py:254)
    int j = k + 2, steps = 0;
    while (j < nw && int(w.e[j] >> 2) < target) {
        ++j;
        if (++steps > GS_SCAN) return GS_ST_UNRES;
    }
    if (TABLES) {
        j = (j < nw) ? int(w.nxa[j]) : int(GS_INF);
    } else {
        while (j < nw && (w.e[j] & 3u) != CLS_AT) ++j;
    }
    if (j >= nw) {
        if (!w.at_end) return GS_ST_UNRES;
        *succ = GS_NONE_E;
        return ST_COMPLETE;
    }
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

// block-wide helpers (GS_THREADS threads, `scratch` = 8 words of shared memory; every thread calls)
__device__ __forceinline__ int gs_block_excl_sum(int v, int* scratch, int* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int nb = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += nb;
    }
    __syncthreads();  // scratch may still be read from the previous call
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < GS_THREADS / 32; ++q) {
        const int x = scratch[q];
        if (q < warp) base += x;
        tot += x;
    }
    *total = tot;
    return base + inc - v;
}
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
// min over the threads AFTER me (exclusive suffix minimum)
__device__ __forceinline__ unsigned int gs_block_suffix_min(unsigned int v, unsigned int* scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int inc = v;  // inclusive suffix min inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int nb = __shfl_down_sync(0xffffffffu, inc, o);
        if (lane + o < 32) inc = min(inc, nb);
    }
    __syncthreads();
    if (lane == 0) scratch[warp] = inc;
    __syncthreads();
    unsigned int later = 0xffffffffu;
#pragma unroll
    for (int q = 0; q < GS_THREADS / 32; ++q)
        if (q > warp) later = min(later, scratch[q]);
    const unsigned int nxt = __shfl_down_sync(0xffffffffu, inc, 1);
    return min(later, lane < 31 ? nxt : 0xffffffffu);
}

__global__ void __launch_bounds__(GS_THREADS) fq_gspec_kernel(const SpecParams p)
{
    ParseState* st = p.st;
    if (*((volatile int*)&st->need_general) == 0 || *((volatile int*)&st->error) != 0) return;
    extern __shared__ __align__(16) uint8_t gs_smem[];
    unsigned int* w_e = reinterpret_cast<unsigned int*>(gs_smem);                         // [GS_W] lines
    unsigned short* s_succ = reinterpret_cast<unsigned short*>(gs_smem + size_t(GS_W) * 4);  // [GS_W] by line
    unsigned short* s_a = s_succ + GS_W;     // next '+' line, then the jump pointers
    unsigned short* s_b = s_a + GS_W;        // next '@' line, then: reach flags (bytes [0, GS_W)) + on-chain rows
    unsigned short* s_cand = s_b + GS_W;     // candidate lines, ascending
    uint8_t* s_reach = reinterpret_cast<uint8_t*>(s_b);
    unsigned short* s_ord = s_b + GS_W / 2;  // [GS_W / 2] (a chain advances >= 4 lines per record)
    __shared__ unsigned int s_cnt[GS_TC + 3], s_off[GS_TC + 3];
    __shared__ unsigned long long s_r0;
    __shared__ unsigned int s_scr[8];
    __shared__ int s_chunk, s_term, s_failflag;
    __shared__ unsigned int s_x;
    __shared__ unsigned long long s_base;
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31;
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
    auto store_row = [&](unsigned long long k, long long ob, const int* rel) {
        if ((long long)k < p.cap) {
            longlong2* row = reinterpret_cast<longlong2*>(p.table + k * 6);
            row[0] = make_longlong2(ob + rel[0], ob + rel[1]);
            row[1] = make_longlong2(ob + rel[2], ob + rel[3]);
            row[2] = make_longlong2(ob + rel[4], ob + rel[5]);
        }
    };
    int pd_c = -1, pd_n = 0;  // the chunk whose rows this CTA still holds (one per thread, in registers)
    long long pd_ob = 0;
    int pd_rel[6] = {0, 0, 0, 0, 0, 0};
    auto flush_pending = [&]() {
        if (pd_c < 0) return;  // uniform
        const unsigned long long base = lookback(pd_c, pd_n);
        if (tid < pd_n) store_row(base + (unsigned long long)tid, pd_ob, pd_rel);
        pd_c = -1;
    };

    for (;;) {
        __syncthreads();  // the previous chunk's shared memory is no longer needed
        if (tid == 0) {
            s_chunk = int(atomicAdd(&st->spec_ticket, 1u));
            s_term = -1;
            s_failflag = 0;
            s_x = 0xffffffffu;
        }
        __syncthreads();
        const int c = s_chunk;
        if (c >= n_chunks) break;
        const int t0 = c * tc;
        const int t1 = (t0 + tc < lv.n_tiles) ? t0 + tc : lv.n_tiles;
        const int tb = c > 0 ? t0 - 1 : t0;
        const int te = (t1 + 1 < lv.n_tiles) ? t1 + 1 : lv.n_tiles;
        const int nt = te - tb;  // <= GS_TC + 2
        if (tid < 32) {  // lines per staged tile and their prefix sums (nt <= 10 tiles: one warp)
            const unsigned int cnt = (lane < nt) ? lv_count(lv, tb + lane) : 0u;
            unsigned int inc = cnt;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (lane <= GS_TC + 2) {
                s_cnt[lane] = cnt;
                s_off[lane] = inc - cnt;  // s_off[nt] = all lines of the window
            }
            if (lane == 0) s_r0 = lv_base(lv, tb);
        }
        __syncthreads();
        const int nb = c > 0 ? int(s_cnt[0]) : 0;   // look-behind lines
        const int nbo = int(s_off[t1 - tb]);         // look-behind + own lines
        const int nw = int(s_off[nt]);
        const int clo = (nb > GS_LB) ? nb - GS_LB : 0;  // first line whose call is needed
        bool fail = nw > GS_W;  // uniform
        SpecWin w;
        w.e = w_e;
        w.nxp = s_a;
        w.nxa = s_b;
        w.nw = nw;
        w.at_end = (te == lv.n_tiles);
        // blob position = rel + bias, rel = (byte index from `base`) - tb * tile + 1
        const long long bias = (long long)tb * lv.tile - 1 - p.mis + p.sentinel;
        w.l_rel = L - bias;
        const unsigned long long R0 = s_r0;  // global rank of window line 0
        int n = 0;                 // rows of this chunk
        unsigned long long x = GX_FAIL, pe_rank = GX_FAIL;
        if (!fail) {
            // ---- A. the window's lines: a warp per tile, a lane per 8 list entries (one 16-byte load) ----
            for (int q = tid >> 5; q < nt; q += GS_THREADS / 32) {
                const int t = tb + q;
                unsigned int cnt = s_cnt[q];
                const unsigned short* src = lv.lists + (size_t)t * (unsigned int)lv.slot_cap;
                unsigned int* dst = w_e + s_off[q];
                const unsigned int relbase = (unsigned int)q * (unsigned int)lv.tile + 1u;
                if (t == 0 && lv.virt) {  // the virtual sentinel leads tile 0: byte index mis - 1 (tb == 0)
                    if (lane == 0) dst[0] = ((unsigned int)lv.mis << 2) | lv.cls0;
                    dst += 1;
                    cnt -= 1;
                }
                for (unsigned int v = lane * 8; v < cnt; v += 256) {
                    const uint4 x = *reinterpret_cast<const uint4*>(src + v);  // the slot is 16-byte aligned and padded
                    const unsigned int ee[8] = {x.x & 0xffffu, x.x >> 16, x.y & 0xffffu, x.y >> 16,
                                                x.z & 0xffffu, x.z >> 16, x.w & 0xffffu, x.w >> 16};
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (v + k < cnt) dst[v + k] = ((relbase + (ee[k] >> 2)) << 2) | (ee[k] & 3u);
                }
            }
            __syncthreads();
            // ---- B. next '+' / next '@' line for every line (backward scans), candidate list ----
            const int per = (nw + GS_THREADS - 1) / GS_THREADS;  // <= GS_STRIP
            const int lo = tid * per < nw ? tid * per : nw;
            const int hi = lo + per < nw ? lo + per : nw;
            unsigned int np = 0xffffu, na = 0xffffu;
            int ncand = 0;
            for (int i = hi - 1; i >= lo; --i) {
                const unsigned int cls = w_e[i] & 3u;
                if (cls == CLS_PLUS) np = (unsigned int)i;
                if (cls == CLS_AT) {
                    na = (unsigned int)i;
                    if (i >= clo && i < nbo) ++ncand;
                }
                s_a[i] = (unsigned short)np;
                s_b[i] = (unsigned short)na;
            }
            const unsigned int after_p = gs_block_suffix_min(np, s_scr);
            const unsigned int after_a = gs_block_suffix_min(na, s_scr);
            int nc;
            int cpos = gs_block_excl_sum(ncand, reinterpret_cast<int*>(s_scr), &nc);
            for (int i = lo; i < hi; ++i) {
                if (s_a[i] == GS_INF) s_a[i] = (unsigned short)after_p;
                if (s_b[i] == GS_INF) s_b[i] = (unsigned short)after_a;
                if ((w_e[i] & 3u) == CLS_AT && i >= clo && i < nbo) s_cand[cpos++] = (unsigned short)i;
            }
            __syncthreads();
            if (nc > GS_CPT * GS_THREADS) fail = true;  // uniform; (the phases below see nc = 0)
            if (fail) nc = 0;
            // ---- C. every candidate makes its call ----
            for (int q = tid; q < nc; q += GS_THREADS) {
                const int i = s_cand[q];
                int rel[6];
                unsigned short s;
                spec_rec<true>(w, i, rel, &s);
                s_succ[i] = s;
            }
            // the head of the whole chain (chunk 0): first "\n@" of the window, it may lie in the look-ahead tile
            const unsigned int head = (nw > 0) ? (unsigned int)s_b[0] : 0xffffu;
            __syncthreads();  // nxp / nxa are dead from here on: s_a = jump pointers, s_b = reach flags + rows
            // ---- D. the chain: start node, reachability by pointer doubling ----
            int rounds = 1;
            while ((1 << rounds) < (nbo - clo) / 4 + 2) ++rounds;
            unsigned int e = 0xffffffffu;   // entry: first chain node at or behind line nb
            int qstart = 0;
            for (int attempt = 0; attempt < GS_STARTS && e == 0xffffffffu; ++attempt) {
                unsigned int start;
                if (c == 0) {
                    start = head;
                    if (head >= (unsigned int)nbo) {  // no "\n@" in the own lines (0xffff: none at all)
                        e = head;
                        break;
                    }
                } else {  // first candidate from qstart on whose call is COMPLETE with a successor in the window
                    unsigned int mine = 0xffffffffu;
                    for (int q = qstart + tid; q < nc; q += GS_THREADS) {
                        const int i = s_cand[q];
                        if (i >= nb) break;
                        if (s_succ[i] < GS_UNRES) {
                            mine = (unsigned int)q;
                            break;
                        }
                    }
                    const unsigned int qs = gs_block_min(mine, s_scr);
                    if (qs == 0xffffffffu) break;  // nothing to start from
                    qstart = int(qs) + 1;
                    start = s_cand[qs];
                }
                for (int q = tid; q < nc; q += GS_THREADS) {
                    const int i = s_cand[q];
                    const unsigned short s = s_succ[i];
                    s_a[i] = (s < (unsigned short)nbo) ? s : GS_INF;  // jumps stay inside look-behind + own lines
                    s_reach[i] = (unsigned int)i == start ? 1 : 0;
                }
                __syncthreads();
                for (int r = 0; r < rounds; ++r) {
                    unsigned short jj[GS_CPT];
#pragma unroll
                    for (int k = 0; k < GS_CPT; ++k) {
                        if (k * GS_THREADS >= nc) break;  // uniform: a chunk of clean records has < GS_THREADS candidates
                        const int q = tid + k * GS_THREADS;
                        jj[k] = GS_INF;
                        if (q < nc) {
                            const int i = s_cand[q];
                            const unsigned short j = s_a[i];
                            if (j != GS_INF) {
                                jj[k] = s_a[j];
                                if (s_reach[i]) s_reach[j] = 1;
                            }
                        }
                    }
                    __syncthreads();
#pragma unroll
                    for (int k = 0; k < GS_CPT; ++k) {
                        if (k * GS_THREADS >= nc) break;
                        const int q = tid + k * GS_THREADS;
                        if (q < nc) s_a[s_cand[q]] = jj[k];
                    }
                    __syncthreads();
                }
                if (c == 0) {
                    e = start;
                } else {  // the chain's edge that crosses into the chunk
                    unsigned int mine = 0xffffffffu;
                    for (int q = tid; q < nc; q += GS_THREADS) {
                        const int i = s_cand[q];
                        if (i >= nb) break;
                        const unsigned short s = s_succ[i];
                        if (s_reach[i] && s < GS_UNRES && s >= (unsigned short)nb) mine = s;
                    }
                    e = gs_block_min(mine, s_scr);  // 0xffffffff: the chain ended before the chunk, try another start
                }
            }
            if (e == 0xffffffffu || e == 0xffffu) {
                if (c == 0 && e == 0xffffu && w.at_end) {
                    x = GX_NONE_E;      // no "\n@" at all: an empty chain
                    pe_rank = GX_NONE_E;
                } else {
                    fail = true;
                }
            } else {
                pe_rank = R0 + e;
                // ---- E. rows: reached candidates of the own lines, in order ----
                const int per_c = (nc + GS_THREADS - 1) / GS_THREADS;
                const int qlo = tid * per_c < nc ? tid * per_c : nc;
                const int qhi = qlo + per_c < nc ? qlo + per_c : nc;
                const int own_lo = (c == 0) ? 0 : nb;
                int rows = 0;
                if (e < (unsigned int)nbo) {
                    for (int q = qlo; q < qhi; ++q) {
                        const int i = s_cand[q];
                        if (i < own_lo || !s_reach[i]) continue;
                        const unsigned short s = s_succ[i];
                        if (s == GS_UNRES) s_failflag = 1;
                        else if (s == GS_NONE_T) s_term = i;          // the chain stops ON this node: not a row
                        else ++rows;
                        if (s >= GS_UNRES || s >= (unsigned short)nbo) s_x = s;  // the one edge that leaves the own lines
                    }
                }
                int total;
                int rpos = gs_block_excl_sum(rows, reinterpret_cast<int*>(s_scr), &total);
                if (e < (unsigned int)nbo) {
                    for (int q = qlo; q < qhi; ++q) {
                        const int i = s_cand[q];
                        if (i < own_lo || !s_reach[i]) continue;
                        if (s_succ[i] != GS_NONE_T && s_succ[i] != GS_UNRES) s_ord[rpos++] = (unsigned short)i;
                    }
                }
                __syncthreads();
                n = total;
                if (s_failflag) {
                    fail = true;
                } else if (e >= (unsigned int)nbo) {
                    x = R0 + e;  // the chain passes over the own lines
                } else {
                    const unsigned int sx = s_x;
                    x = (sx == GS_NONE_T) ? GX_NONE_T : (sx == GS_NONE_E) ? GX_NONE_E : (sx < GS_UNRES ? R0 + sx : GX_FAIL);
                    if (x == GX_FAIL) fail = true;
                }
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
                if (s_term >= 0) {
                    unsigned short s;
                    status = spec_rec<false>(w, s_term, rel, &s);
                }
                st->spec_tail_status = status;
                for (int q = 0; q < 6; ++q) st->spec_tail_pos[q] = rel[q] >= 0 ? (long long)rel[q] + bias : -1;
            }
        }
        // ---- record count: the aggregate is published at once; the look-back for the exclusive prefix and the row
        //      stores are DEFERRED by one chunk (the rows wait in registers), so that a CTA never sits waiting for the
        //      chunks before it: by the time it has resolved its next chunk they have published long ago ----
        if (tid == 0) st_relaxed_gpu(&p.desc[c], ((c == 0 ? 2ull : 1ull) << 62) | (unsigned long long)n);
        int cur_rel[6];
        if (n <= GS_THREADS && tid < n) {
            unsigned short s;
            spec_rec<false>(w, int(s_ord[tid]), cur_rel, &s);
        }
        flush_pending();
        if (n > GS_THREADS) {  // many short records: stored now, straight from shared memory
            const unsigned long long base = lookback(c, n);
            const long long ob = bias + p.goff;
            for (int q = tid; q < n; q += GS_THREADS) {
                int rel[6];
                unsigned short s;
                spec_rec<false>(w, int(s_ord[q]), rel, &s);
                store_row(base + (unsigned long long)q, ob, rel);
            }
        } else {
            pd_c = c;
            pd_n = n;
            pd_ob = bias + p.goff;
#pragma unroll
            for (int q = 0; q < 6; ++q) pd_rel[q] = cur_rel[q];
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
