// fq_general.cuh -- the GENERAL path: exact reproduction of the reference's entrypos chain
// (src/fastqandfurious.py:251-255 over src/_fastqandfurious.c:25-153) for inputs the 4-line fast
// path cannot represent: multi-line sequence / quality, damaged records that make the reference
// resynchronise on the next "\n@", leading garbage, very short lines.
//
// Formulation.  The scan kernel's per-tile newline lists are gathered into a LINE TABLE: position of
// every visible newline plus the class of the byte after it (the input is not read again).  Every '@'-class line is a CANDIDATE record start; one
// entrypos call anchored there is a pure function of the table (fq_general_logic.h: general_rec),
// and so is its successor (the candidate the next call would find).  The reference's output is the
// path from the first candidate through this successor forest.  It is resolved hierarchically:
//   level 1  chunks of 1024 lines: pointer jumping in shared memory gives every candidate its exit
//            from the chunk and the number of records on the way;
//   level 2/3  only nodes that ARE the exit of something walk (<= 64 steps) to the end of their
//            65 536-line / 4 Mi-line block;
//   top      one thread walks the (few) level-3 blocks, then the entries and record-index bases are
//            pushed back down level by level, and each chunk emits its rows in parallel.
// All kernels are enqueued unconditionally and return at once unless the fast path asked for them
// (ParseState::need_general) -- the host never synchronises.
#pragma once
#include "fq_common.cuh"
#include "fq_emit.cuh"
#include "fq_general_logic.h"

namespace fqb {

struct GeneralArrays {
    unsigned long long* nlt;    // [max_lines]
    unsigned int* sumP;         // [nblk + 1]
    unsigned int* sumA;         // [nblk + 1]
    unsigned int* succ;         // [max_lines]
    unsigned long long* jump1;  // [max_lines] exit from the level-1 chunk << 32 | records on the way
    unsigned long long* jump2;  // [max_lines] (flagged nodes only)
    unsigned long long* jump3;  // [max_lines] (flagged nodes only)
    unsigned int* flag1;        // bit set: node is the level-1 exit of some candidate
    unsigned int* flag2;        // bit set: node is the level-2 exit of some listed node
    unsigned int* gminP;        // [ngrp] first '+' / '@' line of each group of G_GRP summary blocks
    unsigned int* gminA;
    unsigned int* gsufP;        // [ngrp + 1] ... of all groups >= g
    unsigned int* gsufA;
    unsigned int* list1;        // [max_lines] distinct level-1 exits (+ head)
    unsigned int* list2;        // [max_lines] distinct level-2 exits (+ head)
    unsigned int* entry1;       // [n1] first chain node inside each level-1 chunk (NONE_T: none)
    unsigned long long* base1;  // [n1] record index of that node
    unsigned int* entry2;       // [n2]
    unsigned long long* base2;
    unsigned int* entry3;       // [n3]
    unsigned long long* base3;
};

inline size_t g_align256(size_t x) { return (x + 255) & ~size_t(255); }

inline size_t carve_general(GeneralArrays& g, uint8_t* b, size_t off, long long max_lines)
{
    memset(&g, 0, sizeof(g));
    if (max_lines <= 0) return off;
    const size_t ml = size_t(max_lines);
    const size_t nblk = (ml + G_BLK - 1) / G_BLK + 1;
    const size_t n1 = (ml + G_S1 - 1) / G_S1 + 1;
    const size_t n2 = (ml + G_S2 - 1) / G_S2 + 1;
    const size_t n3 = (ml + G_S3 - 1) / G_S3 + 1;
    auto take = [&](size_t bytes) {
        uint8_t* p = b + off;
        off += g_align256(bytes);
        return p;
    };
    g.nlt = reinterpret_cast<unsigned long long*>(take(ml * 8));
    g.sumP = reinterpret_cast<unsigned int*>(take(nblk * 4));
    g.sumA = reinterpret_cast<unsigned int*>(take(nblk * 4));
    g.succ = reinterpret_cast<unsigned int*>(take(ml * 4));
    g.jump1 = reinterpret_cast<unsigned long long*>(take(ml * 8));
    g.jump2 = reinterpret_cast<unsigned long long*>(take(ml * 8));
    g.jump3 = reinterpret_cast<unsigned long long*>(take(ml * 8));
    const size_t nflag = (ml + 31) / 32 + 1;
    const size_t ngrp = (nblk + G_GRP - 1) / G_GRP + 2;
    g.flag1 = reinterpret_cast<unsigned int*>(take(nflag * 4));
    g.flag2 = reinterpret_cast<unsigned int*>(take(nflag * 4));
    g.gminP = reinterpret_cast<unsigned int*>(take(ngrp * 4));
    g.gminA = reinterpret_cast<unsigned int*>(take(ngrp * 4));
    g.gsufP = reinterpret_cast<unsigned int*>(take(ngrp * 4));
    g.gsufA = reinterpret_cast<unsigned int*>(take(ngrp * 4));
    g.list1 = reinterpret_cast<unsigned int*>(take(ml * 4));
    g.list2 = reinterpret_cast<unsigned int*>(take(ml * 4));
    g.entry1 = reinterpret_cast<unsigned int*>(take(n1 * 4));
    g.base1 = reinterpret_cast<unsigned long long*>(take(n1 * 8));
    g.entry2 = reinterpret_cast<unsigned int*>(take(n2 * 4));
    g.base2 = reinterpret_cast<unsigned long long*>(take(n2 * 8));
    g.entry3 = reinterpret_cast<unsigned int*>(take(n3 * 4));
    g.base3 = reinterpret_cast<unsigned long long*>(take(n3 * 8));
    return off;
}

struct GeneralParams {
    const uint8_t* base;
    long long A;
    int mis;
    int sentinel;
    long long goff;
    long long* table;
    long long cap;
    ParseState* st;
    fqb_result* res;
    GeneralArrays g;
    unsigned long long max_lines;
    int8_t* qual;
    uint8_t qual_add;
    ListView lv;  // cls0 is filled in on the device
    // byte-range sharding of the general path (fqb_shard_general): the chain enters this shard where the
    // previous one says it resumes, and the records whose leading newline lies before own_end_blob are this
    // shard's; the hand-over {resume, records so far, ended, epoch} travels through peer memory
    int sharded, is_first, is_last;
    long long own_end_blob;                  // blob index of the first byte this shard does not own
    const unsigned long long* entry_slot;    // local [8]: words 0-3 written by the previous shard, word 4 = the epoch
                                             // the NEXT shard has consumed (its acknowledgement)
    unsigned long long* exit_slot;           // the next shard's slot (peer-mapped; nullptr: last shard)
    unsigned long long* ack_left;            // word 4 of the previous shard's slot (peer-mapped; nullptr: first shard)
    unsigned long long epoch, prev_epoch;    // prev_epoch: the hand-over this shard published last (0: none)
};

__device__ __forceinline__ bool general_active(const ParseState* st)
{
    return *((volatile const int*)&st->need_general) != 0 && *((volatile const int*)&st->error) == 0 &&
           *((volatile const int*)&st->general_done) == 0;  // the speculative pass (fq_gspec.cuh) may have answered already
}

__device__ __forceinline__ LineView line_view(const GeneralParams& p)
{
    LineView v;
    v.nlt = p.g.nlt;
    v.sumP = p.g.sumP;
    v.sumA = p.g.sumA;
    v.gsufP = p.g.gsufP;
    v.gsufA = p.g.gsufA;
    v.M = p.st->n_lines;
    v.L = (p.A > 0 ? p.A - p.mis : 0) + p.sentinel;
    v.win = nullptr;
    v.win_lo = 0;
    v.win_n = 0;
    return v;
}

// ---- G0: line table from the scan kernel's per-tile lists (one warp per tile), and the first '+' / '@'
//      line of every block of G_BLK lines (atomicMin into sumP / sumA, preset to NONE_T by launch_general):
//      32 consecutive ranks touch at most two blocks ----
__global__ void __launch_bounds__(256) fq_g_lines_kernel(const GeneralParams p)
{
    if (!general_active(p.st)) return;
    ListView lv = p.lv;
    lv.cls0 = *((volatile unsigned int*)&p.st->cls0);
    const int lane = threadIdx.x & 31;
    const int warp = int((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarps = int((gridDim.x * blockDim.x) >> 5);
    const long long blob_bias = (long long)p.sentinel - p.mis;
    for (int t = warp; t < lv.n_tiles; t += nwarps) {
        const unsigned int n = lv_count(lv, t);
        if (n == 0) continue;
        const unsigned long long B = lv_base(lv, t);
        for (unsigned int j0 = 0; j0 < n; j0 += 32) {
            const unsigned int jj = j0 + lane;
            unsigned int cls = G_CLS_OTHER;
            const unsigned long long r = B + jj;
            const bool live = jj < n && r < p.max_lines;
            if (live) {
                long long a;
                lv_entry(lv, t, jj, &a, &cls);
                p.g.nlt[r] = ((unsigned long long)(a + blob_bias) << 2) | cls;
            }
            const unsigned long long r0 = B + j0;                 // rank of lane 0
            const unsigned long long b_lo = r0 / G_BLK;           // block of lane 0
            const unsigned int split = (unsigned int)((b_lo + 1) * G_BLK - r0);  // lanes >= split sit in the next block
            const unsigned int lo_mask = split >= 32 ? 0xffffffffu : ((1u << split) - 1u);
            const unsigned int bp = __ballot_sync(0xffffffffu, live && cls == G_CLS_PLUS);
            const unsigned int ba = __ballot_sync(0xffffffffu, live && cls == G_CLS_AT);
            if (lane == 0) {
                if (bp & lo_mask) atomicMin(&p.g.sumP[b_lo], (unsigned int)(r0 + __ffs(bp & lo_mask) - 1));
                if (bp & ~lo_mask) atomicMin(&p.g.sumP[b_lo + 1], (unsigned int)(r0 + __ffs(bp & ~lo_mask) - 1));
                if (ba & lo_mask) atomicMin(&p.g.sumA[b_lo], (unsigned int)(r0 + __ffs(ba & lo_mask) - 1));
                if (ba & ~lo_mask) atomicMin(&p.g.sumA[b_lo + 1], (unsigned int)(r0 + __ffs(ba & ~lo_mask) - 1));
            }
        }
    }
}

// append `node` to a list once (first setter of its flag bit wins)
__device__ __forceinline__ void list_add_once(unsigned int* flags, unsigned int* list, unsigned int* count, unsigned int node)
{
    const unsigned int bit = 1u << (node & 31u);
    if (!(atomicOr(&flags[node >> 5], bit) & bit)) list[atomicAdd(count, 1u)] = node;
}

// ---- G2a: suffix-min of the block summaries inside each group of G_GRP blocks (one CTA per group) ----
__global__ void __launch_bounds__(G_GRP) fq_g_suffix_local_kernel(const GeneralParams p)
{
    if (!general_active(p.st)) return;
    const unsigned long long M = p.st->n_lines;
    if (M > p.max_lines || M > 0xfffffff0ull) return;
    const long long nblk = (long long)((M + G_BLK - 1) / G_BLK);
    const long long ngrp = (nblk + G_GRP - 1) / G_GRP;
    __shared__ unsigned int s_p[G_GRP], s_a[G_GRP];
    const int t = threadIdx.x;
    for (long long g = blockIdx.x; g < ngrp; g += gridDim.x) {
        const long long b = g * G_GRP + t;
        s_p[t] = (b < nblk) ? p.g.sumP[b] : NONE_T;
        s_a[t] = (b < nblk) ? p.g.sumA[b] : NONE_T;
        __syncthreads();
        for (int o = 1; o < G_GRP; o <<= 1) {  // Hillis-Steele suffix-min
            unsigned int vp = s_p[t], va = s_a[t];
            if (t + o < G_GRP) {
                vp = min(vp, s_p[t + o]);
                va = min(va, s_a[t + o]);
            }
            __syncthreads();
            s_p[t] = vp;
            s_a[t] = va;
            __syncthreads();
        }
        if (b < nblk) {
            p.g.sumP[b] = s_p[t];
            p.g.sumA[b] = s_a[t];
        }
        if (t == 0) {
            p.g.gminP[g] = s_p[0];
            p.g.gminA[g] = s_a[0];
        }
        __syncthreads();
    }
}

// ---- G2b: suffix-min over the groups (single CTA), head of the chain, error checks ----
__global__ void __launch_bounds__(1024) fq_g_suffix_top_kernel(const GeneralParams p)
{
    if (!general_active(p.st)) return;
    const unsigned long long M = p.st->n_lines;
    if (M > p.max_lines || M > 0xfffffff0ull) {
        if (threadIdx.x == 0) {
            p.st->error = (M > 0xfffffff0ull) ? FQB_ERR_TOO_MANY_LINES : FQB_ERR_WORKSPACE;
        }
        return;
    }
    const long long nblk = (long long)((M + G_BLK - 1) / G_BLK);
    const long long ngrp = (nblk + G_GRP - 1) / G_GRP;
    __shared__ unsigned int s_p[1024], s_a[1024];
    const int t = threadIdx.x;
    const long long seg = (ngrp + 1023) / 1024;
    const long long lo = (long long)t * seg;
    long long hi = lo + seg;
    if (hi > ngrp) hi = ngrp;
    unsigned int mp = NONE_T, ma = NONE_T;
    for (long long g = lo; g < hi; ++g) {
        mp = min(mp, p.g.gminP[g]);
        ma = min(ma, p.g.gminA[g]);
    }
    s_p[t] = mp;
    s_a[t] = ma;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        unsigned int vp = s_p[t], va = s_a[t];
        if (t + o < 1024) {
            vp = min(vp, s_p[t + o]);
            va = min(va, s_a[t + o]);
        }
        __syncthreads();
        s_p[t] = vp;
        s_a[t] = va;
        __syncthreads();
    }
    unsigned int cp = (t + 1 < 1024) ? s_p[t + 1] : NONE_T;  // everything to the right of my segment
    unsigned int ca = (t + 1 < 1024) ? s_a[t + 1] : NONE_T;
    for (long long g = hi - 1; g >= lo; --g) {
        cp = min(cp, p.g.gminP[g]);
        ca = min(ca, p.g.gminA[g]);
        p.g.gsufP[g] = cp;
        p.g.gsufA[g] = ca;
    }
    if (t == 0) {
        p.g.gsufP[ngrp] = NONE_T;
        p.g.gsufA[ngrp] = NONE_T;
        // first '@'-class line (NONE_T if there is none); a later shard learns its entry from its predecessor
        const unsigned int head = (p.sharded && !p.is_first) ? NONE_T : s_a[0];
        p.st->head = head;
        p.st->terminal = NONE_X;
        p.st->n_chain = 0;
        if (head < NONE_MIN) {  // the head walks on every level
            list_add_once(p.g.flag1, p.g.list1, &p.st->n_list1, head);
            list_add_once(p.g.flag2, p.g.list2, &p.st->n_list2, head);
        }
    }
}

// Window of the line table staged per level-1 chunk: the chunk's own lines plus a tail, so that a record
// that starts near the end of the chunk still finds its lines in shared memory (anything beyond falls back
// to the global table).
constexpr int G_TAIL = 128;
constexpr int G_WIN = G_S1 + G_TAIL;

template <int THREADS>
__device__ __forceinline__ void stage_window(const GeneralParams& p, unsigned long long* win, unsigned long long lo,
                                             unsigned long long M, LineView& v)
{
    unsigned long long n = M - lo;
    if (n > G_WIN) n = G_WIN;
    // all loads of a thread are issued before the first store: one round trip per chunk, not one per line
    constexpr int N = (G_WIN + THREADS - 1) / THREADS;
    unsigned long long r[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const unsigned int l = k * THREADS + threadIdx.x;
        r[k] = (l < (unsigned int)n) ? p.g.nlt[lo + l] : 0ull;
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const unsigned int l = k * THREADS + threadIdx.x;
        if (l < (unsigned int)n) win[l] = r[k];
    }
    v.win = win;
    v.win_lo = lo;
    v.win_n = (unsigned int)n;
}

// ---- G3+G4: per chunk of G_S1 lines -- successor of every candidate (one entrypos call answered from the
//      staged window), then level 1: exit from the chunk and records on the way for EVERY candidate.
//      Successors always lie ahead, so the candidates (kept sorted by line) are resolved right to left in
//      blocks of 32 by ONE warp: pointers into later blocks are final after one hop, pointers inside the
//      block are resolved by <= 6 rounds of warp-synchronous pointer jumping -- no CTA barrier per round. ----
constexpr int G_CHUNK_THREADS = 128;

__global__ void __launch_bounds__(G_CHUNK_THREADS, 12) fq_g_chunk_kernel(const GeneralParams p)
{
    if (!general_active(p.st)) return;
    LineView v = line_view(p);
    const unsigned long long M = v.M;
    const unsigned long long n1 = (M + G_S1 - 1) / G_S1;
    __shared__ unsigned long long win[G_WIN];
    __shared__ unsigned int nxt[G_S1];          // successor, then exit from the chunk (indexed by line)
    __shared__ unsigned short cnt[G_S1];        // records on the way to the exit (<= G_S1 / 2)
    __shared__ unsigned int cbits[G_S1 / 32];   // candidate bitmap
    __shared__ unsigned short cbase[G_S1 / 32 + 1];  // candidates before each bitmap word
    __shared__ unsigned short clist[G_S1];      // candidate lines, ascending
    constexpr int PER = G_S1 / G_CHUNK_THREADS;
    const int lane = threadIdx.x & 31;
    for (unsigned long long c = blockIdx.x; c < n1; c += gridDim.x) {
        const unsigned long long lo = c * G_S1;
        unsigned long long hi = lo + G_S1;
        if (hi > M) hi = M;
        stage_window<G_CHUNK_THREADS>(p, win, lo, M, v);
        __syncthreads();
        // the chunk's candidates ('@'-class lines)
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int l = q * G_CHUNK_THREADS + threadIdx.x;
            const bool is_c = (lo + l < hi) && ((win[l] & 3ull) == G_CLS_AT);
            const unsigned int b = __ballot_sync(0xffffffffu, is_c);
            if (lane == 0) cbits[l >> 5] = b;
        }
        __syncthreads();
        if (threadIdx.x < 32) {  // candidates before each word (G_S1 / 32 == 32 words: one per lane)
            const int pc = __popc(cbits[lane]);
            int inc = pc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int nb = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += nb;
            }
            cbase[lane] = (unsigned short)(inc - pc);
            if (lane == 31) cbase[32] = (unsigned short)inc;
        }
        __syncthreads();
        const unsigned int nc = cbase[G_S1 / 32];
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int l = q * G_CHUNK_THREADS + threadIdx.x;
            const unsigned int b = cbits[l >> 5];
            if ((b >> lane) & 1u) clist[cbase[l >> 5] + __popc(b & ((1u << lane) - 1u))] = (unsigned short)l;
        }
        __syncthreads();
        // every candidate makes its entrypos call (one per thread)
        for (unsigned int q = threadIdx.x; q < nc; q += G_CHUNK_THREADS) {
            const unsigned int l = clist[q];
            unsigned int s;
            long long pos[6];
            general_rec(v, lo + l, pos, true, &s);
            p.g.succ[lo + l] = s;  // the emission walks along these
            nxt[l] = s;
            cnt[l] = (s == NONE_T) ? 0 : 1;  // NONE_T: the call is not COMPLETE, the chain stops ON this node
        }
        __syncthreads();
        if (threadIdx.x < 32 && nc > 0) {  // level 1, right to left in blocks of 32 candidates
            for (int blk = int((nc - 1) >> 5); blk >= 0; --blk) {
                const unsigned int q = (unsigned int)blk * 32 + lane;
                const bool mine = q < nc;
                const unsigned int l = mine ? clist[q] : 0;
                for (int r = 0; r < 6; ++r) {  // 1 hop into the resolved blocks + log2(32) inside this one
                    unsigned int n = NONE_T, n2 = 0, c2 = 0;
                    if (mine) {
                        n = nxt[l];
                        if (n < hi) {  // still a line of the chunk: follow it
                            n2 = nxt[n - lo];
                            c2 = cnt[n - lo];
                        }
                    }
                    if (!__any_sync(0xffffffffu, n < hi)) break;
                    __syncwarp();
                    if (mine && n < hi) {
                        nxt[l] = n2;
                        cnt[l] = (unsigned short)(cnt[l] + c2);
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        for (unsigned int q = threadIdx.x; q < nc; q += G_CHUNK_THREADS) {
            const unsigned int l = clist[q];
            const unsigned int e = nxt[l];
            p.g.jump1[lo + l] = pack_jump(e, cnt[l]);
            // one list insertion per distinct exit and warp (most candidates of a chunk share their exit)
            const unsigned int peers = __match_any_sync(__activemask(), e);
            if (e < NONE_MIN && (unsigned int)lane == (unsigned int)(__ffs(peers) - 1))
                list_add_once(p.g.flag1, p.g.list1, &p.st->n_list1, e);
        }
        __syncthreads();
    }
}

// ---- G5/G6: levels 2 and 3 -- listed nodes walk to the end of their block ----
__global__ void __launch_bounds__(256) fq_g_walk_kernel(const GeneralParams p, int level)
{
    if (!general_active(p.st)) return;
    const unsigned int* list = (level == 2) ? p.g.list1 : p.g.list2;
    const unsigned int nl = (level == 2) ? p.st->n_list1 : p.st->n_list2;
    const unsigned long long* jin = (level == 2) ? p.g.jump1 : p.g.jump2;
    unsigned long long* jout = (level == 2) ? p.g.jump2 : p.g.jump3;
    const unsigned long long S = (level == 2) ? (unsigned long long)G_S2 : (unsigned long long)G_S3;
    const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int nthreads = gridDim.x * blockDim.x;
    for (unsigned int q = tid; q < nl; q += nthreads) {
        const unsigned long long i = list[q];
        const unsigned long long end = (i / S + 1) * S;
        unsigned long long cur = i;
        unsigned int hops = 0, e;
        for (;;) {
            const unsigned long long j = jin[cur];
            e = jump_exit(j);
            hops += jump_hops(j);
            if (e >= NONE_MIN || e >= end) break;
            cur = e;
        }
        jout[i] = pack_jump(e, hops);
        if (level == 2 && e < NONE_MIN) list_add_once(p.g.flag2, p.g.list2, &p.st->n_list2, e);
    }
}

// ---- G7: top -- one thread walks the level-3 blocks ----
__global__ void __launch_bounds__(1024) fq_g_top_kernel(const GeneralParams p)
{
    if (!general_active(p.st)) return;
    const unsigned long long M = p.st->n_lines;
    const unsigned long long n3 = (M + G_S3 - 1) / G_S3;
    for (unsigned long long b = threadIdx.x; b < n3; b += blockDim.x) p.g.entry3[b] = NONE_T;
    __syncthreads();
    if (threadIdx.x != 0) return;
    unsigned long long total = 0;
    unsigned int cur = p.st->head;
    while (cur < NONE_MIN) {
        const unsigned long long b = cur / G_S3;
        p.g.entry3[b] = cur;
        p.g.base3[b] = total;
        const unsigned long long j = p.g.jump3[cur];
        total += jump_hops(j);
        cur = jump_exit(j);
    }
    p.st->n_chain = total;
}

// ---- G8/G9: push entries and record-index bases down one level ----
__global__ void __launch_bounds__(64) fq_g_down_kernel(const GeneralParams p, int level /* parent level: 3 or 2 */)
{
    if (!general_active(p.st)) return;
    const unsigned long long M = p.st->n_lines;
    const unsigned long long SP = (level == 3) ? (unsigned long long)G_S3 : (unsigned long long)G_S2;
    const unsigned long long SC = (level == 3) ? (unsigned long long)G_S2 : (unsigned long long)G_S1;
    const unsigned int* pentry = (level == 3) ? p.g.entry3 : p.g.entry2;
    const unsigned long long* pbase = (level == 3) ? p.g.base3 : p.g.base2;
    unsigned int* centry = (level == 3) ? p.g.entry2 : p.g.entry1;
    unsigned long long* cbase = (level == 3) ? p.g.base2 : p.g.base1;
    const unsigned long long* jmp = (level == 3) ? p.g.jump2 : p.g.jump1;
    const unsigned long long np = (M + SP - 1) / SP;
    const unsigned long long nc = (M + SC - 1) / SC;
    // one CTA (64 threads) per parent block: clear its children, then thread 0 walks
    for (unsigned long long b = blockIdx.x; b < np; b += gridDim.x) {
        const unsigned long long c0 = b * G_FAN;
        if (c0 + threadIdx.x < nc) centry[c0 + threadIdx.x] = NONE_T;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int cur = pentry[b];
            if (cur != NONE_T) {
                unsigned long long r = pbase[b];
                const unsigned long long end = (b + 1) * SP;
                while (cur < NONE_MIN && cur < end) {
                    const unsigned long long c = cur / SC;
                    centry[c] = cur;
                    cbase[c] = r;
                    const unsigned long long j = jmp[cur];
                    r += jump_hops(j);
                    cur = jump_exit(j);
                }
            }
        }
        __syncthreads();
    }
}

// ---- G10: emission -- every level-1 chunk writes the rows of the chain nodes it holds ----
constexpr int G_EMIT_THREADS = 128;  // small CTAs: the serial walks of many chunks overlap on an SM

__global__ void __launch_bounds__(G_EMIT_THREADS) fq_g_emit_kernel(const GeneralParams p)
{
    if (!general_active(p.st)) return;
    LineView v = line_view(p);
    const unsigned long long M = v.M;
    const unsigned long long n1 = (M + G_S1 - 1) / G_S1;
    __shared__ unsigned long long win[G_WIN];
    __shared__ unsigned int s_succ[G_S1];
    __shared__ unsigned short s_ord[G_S1 / 2];
    __shared__ int s_n;
    for (unsigned long long c = blockIdx.x; c < n1; c += gridDim.x) {
        const unsigned int entry = p.g.entry1[c];
        if (entry == NONE_T) continue;  // uniform for the CTA
        const unsigned long long lo = c * G_S1;
        unsigned long long hi = lo + G_S1;
        if (hi > M) hi = M;
        // successors of the chunk's lines (only those of candidates are defined -- and only those are read)
        {
            unsigned int r[G_S1 / G_EMIT_THREADS];
#pragma unroll
            for (int k = 0; k < G_S1 / G_EMIT_THREADS; ++k) {
                const int l = k * G_EMIT_THREADS + threadIdx.x;
                r[k] = (lo + l < hi) ? p.g.succ[lo + l] : NONE_X;
            }
#pragma unroll
            for (int k = 0; k < G_S1 / G_EMIT_THREADS; ++k) s_succ[k * G_EMIT_THREADS + threadIdx.x] = r[k];
        }
        stage_window<G_EMIT_THREADS>(p, win, lo, M, v);
        __syncthreads();
        if (threadIdx.x == 0) {
            int n = 0;
            unsigned long long cur = entry;
            for (;;) {
                const unsigned int s = s_succ[cur - lo];
                if (s == NONE_T) {  // the chain stops on this node
                    p.st->terminal = (unsigned int)cur;
                    break;
                }
                s_ord[n++] = (unsigned short)(cur - lo);
                if (s == NONE_E) {  // COMPLETE, and no further "\n@"
                    p.st->terminal = NONE_E;
                    break;
                }
                if (s >= hi) break;
                cur = s;
            }
            s_n = n;
        }
        __syncthreads();
        const int n = s_n;
        const unsigned long long r0 = p.g.base1[c];
        for (int q = threadIdx.x; q < n; q += blockDim.x) {
            long long pos[6];
            general_rec(v, lo + s_ord[q], pos, false, nullptr);
            const unsigned long long k = r0 + (unsigned long long)q;
            if ((long long)k < p.cap) {
                longlong2* row = reinterpret_cast<longlong2*>(p.table + k * 6);
                row[0] = make_longlong2(pos[0] + p.goff, pos[1] + p.goff);
                row[1] = make_longlong2(pos[2] + p.goff, pos[3] + p.goff);
                row[2] = make_longlong2(pos[4] + p.goff, pos[5] + p.goff);
            }
        }
        __syncthreads();
    }
}

// ---- G11: result header ----
__global__ void fq_g_result_kernel(const GeneralParams p)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ParseState* st = p.st;
    if (!*((volatile int*)&st->need_general)) return;  // the fast path's result stands
    if (*((volatile int*)&st->general_done)) return;   // so does the speculative pass's
    const long long first_bad = st->first_bad_inv ? (long long)~st->first_bad_inv : -1;
    if (st->error) {
        write_result(p.res, 0, 0, ST_NO_HEAD_BEG, nullptr, FQB_PATH_GENERAL, st->error, 0, (long long)st->n_lines,
                     first_bad);
        return;
    }
    const LineView v = line_view(p);
    const long long n = (long long)st->n_chain;
    long long pos[6] = {-1, -1, -1, -1, -1, -1};
    int status = ST_NO_HEAD_BEG;
    const unsigned int term = st->terminal;
    if (term < NONE_MIN) status = general_rec(v, term, pos, false, nullptr);
    int error = FQB_OK;
    long long resume = 0;
    if (n + 1 > p.cap)
        error = FQB_ERR_CAPACITY;
    else if (n >= 1)
        resume = p.table[(n - 1) * 6 + 5] - p.goff - 1;
    write_result(p.res, n, resume, status, pos, FQB_PATH_GENERAL, error, 0, (long long)v.M, first_bad);
}

// ---- G12: Phred decode of the stored records (one warp per record) ----
__global__ void __launch_bounds__(256) fq_g_decode_kernel(const GeneralParams p)
{
    // (also after the speculative pass: it leaves n_chain and the rows like the exact path does)
    if (*((volatile int*)&p.st->need_general) == 0 || *((volatile int*)&p.st->error) != 0 || !p.qual) return;
    long long n = (long long)p.st->n_chain;
    if (n > p.cap) n = p.cap;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    const uint8_t* buf = p.base + p.mis;  // caller's byte i; blob index = i + sentinel
    for (long long k = warp; k < n; k += nwarps) {
        const long long b = p.table[k * 6 + 4] - p.goff - p.sentinel;
        const long long e = p.table[k * 6 + 5] - p.goff - p.sentinel;
        for (long long i = b + lane; i < e; i += 32) p.qual[i] = int8_t(uint8_t(buf[i] + p.qual_add));
    }
}

// ---- sharded general path: entry of the chain into this shard ----
// Thread 0 waits for the previous shard's hand-over (acquire loads of a local slot it writes over NVLink, 10 s
// timeout -> FQB_ERR_PEER) and turns "the chain resumes its search at absolute position X" into the head node:
// the first '@'-class line at a blob position >= X - goff -- what the next entrypos call of the reference
// would find.  The first shard keeps the head fq_g_suffix_top_kernel chose.
__global__ void fq_g_head_kernel(const GeneralParams p)
{
    if (!p.sharded || threadIdx.x != 0 || blockIdx.x != 0) return;
    ParseState* st = p.st;
    if (!general_active(st)) return;
    if (p.is_first) {
        st->shard_resume_abs = (unsigned long long)p.goff;  // blob index 0
        st->shard_records_before = 0;
        st->shard_ended = 0;
        return;
    }
    const unsigned long long t0 = global_timer_ns();
    unsigned int spins = 0;
    while (ld_acquire_sys(p.entry_slot + 3) != p.epoch) {
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 10000000000ull) {
            st->error = FQB_ERR_PEER;
            return;
        }
    }
    const unsigned long long resume_abs = ld_acquire_sys(p.entry_slot + 0);
    st->shard_resume_abs = resume_abs;
    st->shard_records_before = ld_acquire_sys(p.entry_slot + 1);
    st->shard_ended = (int)ld_acquire_sys(p.entry_slot + 2);
    // acknowledge: the previous shard may overwrite the slot with its next hand-over
    if (p.ack_left) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.ack_left), "l"(p.epoch) : "memory");
    if (st->shard_ended) return;  // head stays NONE_T: nothing to emit
    const LineView v = line_view(p);
    const long long xr = (long long)resume_abs - p.goff;  // blob index the search starts at
    unsigned long long lo = 0, hi = v.M;                  // first line with position >= xr
    while (lo < hi) {
        const unsigned long long mid = (lo + hi) >> 1;
        if (line_pos(v, mid) < xr)
            lo = mid + 1;
        else
            hi = mid;
    }
    const unsigned int head = next_of_class(v, lo, G_CLS_AT, v.sumA, v.gsufA);
    st->head = head;
    if (head < NONE_MIN) {
        list_add_once(p.g.flag1, p.g.list1, &st->n_list1, head);
        list_add_once(p.g.flag2, p.g.list2, &st->n_list2, head);
    }
}

// ---- sharded general path: ownership, hand-over to the next shard, result header ----
// The chain was resolved over own bytes + halo; the rows whose leading newline lies in the own range are this
// shard's (rows are in chain order, so they are a prefix of the table).  The next shard resumes the search at
// pos5 - 1 of the last owned record (src/fastqandfurious.py:254), exactly like the next entrypos call.
__global__ void fq_g_shard_result_kernel(const GeneralParams p)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ParseState* st = p.st;
    auto publish = [&](unsigned long long resume_abs, unsigned long long records, unsigned long long ended) {
        if (!p.exit_slot) return;
        if (p.prev_epoch) {  // the next shard must have consumed our previous hand-over (its slot is not a queue)
            const unsigned long long t0 = global_timer_ns();
            unsigned int spins = 0;
            while (ld_acquire_sys(p.entry_slot + 4) < p.prev_epoch) {
                if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 10000000000ull) break;  // give up: it will time out too
            }
        }
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p.exit_slot + 0), "l"(resume_abs) : "memory");
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p.exit_slot + 1), "l"(records) : "memory");
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p.exit_slot + 2), "l"(ended) : "memory");
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.exit_slot + 3), "l"(p.epoch) : "memory");
    };
    const long long M = (long long)st->n_lines;
    if (st->error) {
        write_result(p.res, 0, 0, ST_NO_HEAD_BEG, nullptr, FQB_PATH_GENERAL, st->error, 0, M, -1);
        publish(0, 0, 2);  // later shards stop instead of waiting
        return;
    }
    const unsigned long long before = st->shard_records_before;
    if (st->shard_ended) {  // the chain ended (or failed) in an earlier shard
        write_result(p.res, 0, 0, ST_NO_HEAD_BEG, nullptr, FQB_PATH_GENERAL, st->shard_ended == 2 ? FQB_ERR_PEER : FQB_OK, 0, M,
                     -1);
        p.res->reserved[0] = (long long)before;
        p.res->reserved[1] = 1;  // "ended upstream"
        publish(st->shard_resume_abs, before, (unsigned long long)st->shard_ended);
        return;
    }
    const LineView v = line_view(p);
    const long long n_chain = (long long)st->n_chain;
    if (n_chain + 1 > p.cap) {
        write_result(p.res, n_chain, 0, ST_NO_HEAD_BEG, nullptr, FQB_PATH_GENERAL, FQB_ERR_CAPACITY, 0, M, -1);
        publish(0, 0, 2);
        return;
    }
    const unsigned int term = st->terminal;
    long long pos[6] = {-1, -1, -1, -1, -1, -1};
    if (p.is_last) {  // the end of the stream: the ordinary result
        int status = ST_NO_HEAD_BEG;
        if (term < NONE_MIN) status = general_rec(v, term, pos, false, nullptr);
        const long long resume = n_chain >= 1 ? p.table[(n_chain - 1) * 6 + 5] - p.goff - 1 : 0;
        write_result(p.res, n_chain, resume, status, pos, FQB_PATH_GENERAL, FQB_OK, 0, M, -1);
        p.res->reserved[0] = (long long)before;
        return;
    }
    // rows owned: leading newline (pos0 - 1, absolute) before the end of the own range
    const long long own_end_abs = p.goff + p.own_end_blob;
    long long lo = 0, hi = n_chain;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (p.table[mid * 6] - 1 < own_end_abs)
            lo = mid + 1;
        else
            hi = mid;
    }
    const long long n_owned = lo;
    const unsigned long long resume_abs =
        n_owned >= 1 ? (unsigned long long)(p.table[(n_owned - 1) * 6 + 5] - 1) : st->shard_resume_abs;
    int status = ST_COMPLETE;  // the chain continues in the next shard
    int error = FQB_OK;
    unsigned long long ended = 0;
    if (n_owned == n_chain && term < NONE_MIN && line_pos(v, term) < p.own_end_blob) {
        // the chain stops ON a node this shard owns: a broken record, or one that needs more than the halo
        status = general_rec(v, term, pos, false, nullptr);
        if (status == ST_INVALID) {
            ended = 1;
        } else {
            error = FQB_ERR_HALO;
            ended = 2;
        }
    }
    const long long resume_blob = n_owned >= 1 ? p.table[(n_owned - 1) * 6 + 5] - p.goff - 1 : 0;
    write_result(p.res, n_owned, ended ? resume_blob : 0, status, ended == 1 ? pos : nullptr, FQB_PATH_GENERAL, error, 0, M, -1);
    p.res->reserved[0] = (long long)before;
    publish(resume_abs, before + (unsigned long long)n_owned, ended);
}

inline cudaError_t launch_general(const GeneralParams& gp, int sms, cudaStream_t stream)
{
    cudaError_t e;
    {  // block summaries start at "none", the list-membership flags at 0
        const size_t ml = size_t(gp.max_lines);
        const size_t nblk = (ml + G_BLK - 1) / G_BLK + 1;
        const size_t nflag = (ml + 31) / 32 + 1;
        if ((e = cudaMemsetAsync(gp.g.sumP, 0xff, nblk * 4, stream)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(gp.g.sumA, 0xff, nblk * 4, stream)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(gp.g.flag1, 0, nflag * 4, stream)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(gp.g.flag2, 0, nflag * 4, stream)) != cudaSuccess) return e;
    }
    fq_g_lines_kernel<<<sms * 8, 256, 0, stream>>>(gp);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    fq_g_suffix_local_kernel<<<sms * 2, G_GRP, 0, stream>>>(gp);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    fq_g_suffix_top_kernel<<<1, 1024, 0, stream>>>(gp);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    fq_g_chunk_kernel<<<sms * 12, G_CHUNK_THREADS, 0, stream>>>(gp);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (gp.sharded) {  // the one place a shard depends on its predecessor
        fq_g_head_kernel<<<1, 32, 0, stream>>>(gp);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    fq_g_walk_kernel<<<sms * 8, 256, 0, stream>>>(gp, 2);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    fq_g_walk_kernel<<<sms * 8, 256, 0, stream>>>(gp, 3);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    fq_g_top_kernel<<<1, 1024, 0, stream>>>(gp);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    fq_g_down_kernel<<<sms * 4, 64, 0, stream>>>(gp, 3);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    fq_g_down_kernel<<<sms * 8, 64, 0, stream>>>(gp, 2);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    fq_g_emit_kernel<<<sms * 12, G_EMIT_THREADS, 0, stream>>>(gp);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (gp.sharded)
        fq_g_shard_result_kernel<<<1, 32, 0, stream>>>(gp);
    else
        fq_g_result_kernel<<<1, 32, 0, stream>>>(gp);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (gp.qual) {
        fq_g_decode_kernel<<<sms * 8, 256, 0, stream>>>(gp);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace fqb
