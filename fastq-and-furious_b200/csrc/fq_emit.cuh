// fq_emit.cuh -- second kernel of the 4-line fast path: offset rows from the newline lists.
//
// With four lines per record the newline of global rank r belongs to record r >> 2, field r & 3:
//   field 0 -> '\n' before '@' (pos0 = p + 1)      field 1 -> header '\n' (pos1 = p, pos2 = p + 1)
//   field 2 -> '\n' before '+' (pos3 = p)          field 3 -> end of the '+' line (pos4 = p + 1)
// and pos5 = pos4 + pos3 - pos2 (the reference never searches for it, src/_fastqandfurious.c:129).
// One warp per tile; a lane takes one record whose field-0 newline lies in the tile, pulls the next
// four newlines from the lists (running into the following tiles where the record straddles), writes
// the 6 x int64 row with three 16-byte stores and checks the conditions under which this "rank mod
// 4" parse is provably identical to the reference's sequential memmem/memchr chain
// (src/_fastqandfurious.c:62-136, DESIGN.md "equivalence conditions").  Any failed check hands the
// whole buffer to the general path.  Optional Phred decode of the record's quality span.
// The last CTA to finish classifies the still-open last record exactly like one more entrypos call
// would (status code + posbuffer) and writes the result header.
#pragma once
#include "fq_common.cuh"

namespace fqb {

struct EmitParams {
    const uint8_t* base;
    long long A;
    int mis;
    int sentinel;
    long long out_bias;  // emitted = a + out_bias
    long long goff;      // emitted = blob + goff
    long long* table;
    long long cap;
    ListView lv;         // cls0 is filled in on the device
    ParseState* st;
    fqb_result* res;
    int8_t* qual;        // mirror of the caller's buffer or nullptr
    unsigned int qual_add;
    int force_general;   // 1: skip the fast path, hand over to the general path
    // byte-range sharding (fqb_shard_*): this buffer is one shard (+ halo) of a longer stream
    const unsigned long long* line_base;  // device: lines owned by all earlier shards (nullptr: 0)
    long long own_end;   // newlines at byte index (from base) < own_end are owned by this shard
    int is_last;         // 1: the buffer ends where the stream ends (always 1 without sharding)
    int sharded;         // 1: shard mode (no general path; halo / ownership rules apply)
    // fused exchange (fqb_shard_emit_wait): instead of line_base, the line counts of the earlier shards
    // arrive in LOCAL memory, stored there by the peers' scans over NVLink as {count, epoch} pairs; the
    // kernel itself waits for them (no collective call, no barrier kernel between scan and emit)
    const unsigned long long* wait_slots;  // [n_wait][2]
    int n_wait;
    unsigned long long epoch;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Lines owned by all earlier shards: either given (line_base) or collected from the slots the peers write.
// Thread 0 of every CTA polls; a peer that never publishes ends the wait after 10 s with FQB_ERR_PEER.
__device__ inline unsigned long long shard_line_base(const EmitParams& p, bool* ok)
{
    *ok = true;
    if (!p.wait_slots) return p.line_base ? *p.line_base : 0ull;
    __shared__ unsigned long long s_base;
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        unsigned long long sum = 0;
        int good = 1;
        const unsigned long long t0 = global_timer_ns();
        for (int r = 0; r < p.n_wait && good; ++r) {
            const unsigned long long* slot = p.wait_slots + 2 * r;
            unsigned int spins = 0;
            while (ld_acquire_sys(slot + 1) != p.epoch) {
                if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 10000000000ull) {
                    good = 0;
                    break;
                }
            }
            if (good) sum += ld_acquire_sys(slot);
        }
        s_base = sum;
        s_ok = good;
    }
    __syncthreads();
    *ok = s_ok != 0;
    return s_base;
}

// number of (augmented) list entries at byte index < a_end
__device__ inline unsigned long long lv_count_before(const ListView& lv, long long a_end)
{
    if (a_end <= 0) return (lv.virt && a_end > (long long)lv.mis - 1) ? 1ull : 0ull;
    long long te = a_end / lv.tile;
    if (te >= lv.n_tiles) {
        if (lv.n_tiles == 0) return 0ull;
        te = lv.n_tiles - 1;
    }
    const int t = int(te);
    unsigned long long total = lv_base(lv, t);
    const unsigned int n = lv_count(lv, t);
    for (unsigned int jj = 0; jj < n; ++jj) {
        long long a;
        unsigned int cls;
        lv_entry(lv, t, jj, &a, &cls);
        if (a < a_end) ++total;
    }
    return total;
}

// entrypos on the open last record.  nl[0..cnt) are ALL visible newlines (blob coordinates) at or
// after the search offset, in order.  Mirrors src/_fastqandfurious.c:57-136 with the memmem/memchr
// calls answered from that list.  Returns the status; pos[] is -1 filled like :57-59.
__device__ inline int classify_tail(const uint8_t* blob0 /* address of blob[0] (may be virtual) */,
                                    long long L, const long long* nl, int cnt, long long* pos)
{
    for (int i = 0; i < 6; ++i) pos[i] = -1;
    int i = 0;
    while (i < cnt && blob0[nl[i] + 1] != '@') ++i;  // first "\n@"  (:62)
    if (i == cnt) return ST_NO_HEAD_BEG;
    const long long p0 = nl[i] + 1;
    pos[0] = p0;
    if (i + 1 >= cnt) return ST_NO_HEAD_END;  // header '\n' (:70-77)
    const long long p1 = nl[i + 1];
    pos[1] = p1;
    const long long p2 = p1 + 1;
    pos[2] = p2;
    int j = i + 2;  // "\n+" at or after p2 + 1 (:87-94)
    while (j < cnt && !(nl[j] >= p2 + 1 && blob0[nl[j] + 1] == '+')) ++j;
    if (j >= cnt) return ST_NO_SEQ_END;
    const long long p3 = nl[j];
    pos[3] = p3;
    if (p3 + 2 >= L) return ST_NO_QUALHEAD_END;   // (:97-101)
    if (j + 1 >= cnt) return ST_NO_QUALHEAD_END;  // end of the '+' line (:102-107)
    const long long h = nl[j + 1];
    if ((h - p3 - 1) > 1 && (h - p3) != (p1 - p0 + 1)) return ST_INVALID;  // (:109-117)
    const long long p4 = h + 1;
    pos[4] = p4;
    const long long p5 = p4 + p3 - p1 - 1;  // (:129)
    if (p5 + 2 >= L) return ST_NO_QUAL_END;
    pos[5] = p5;
    return ST_COMPLETE;
}

__device__ inline void write_result(fqb_result* r, long long n, long long resume, int status, const long long* pos,
                                    int path, int error, int need_general, long long n_lines, long long first_bad)
{
    r->n_records = n;
    r->resume_offset = resume;
    for (int i = 0; i < 6; ++i) r->tail_pos[i] = pos ? pos[i] : -1;
    r->tail_status = status;
    r->path = path;
    r->error = error;
    r->need_general = need_general;
    r->n_lines = n_lines;
    r->first_bad = first_bad;
    for (int i = 0; i < 4; ++i) r->reserved[i] = 0;
}

// The last (up to) 9 newlines of the buffer in blob coordinates, nl[8] = newest, gathered by ONE WARP: a lane per tile
// for the counts of the last 32 tiles (one round of loads), then a lane per wanted newline (a second round) --
// instead of one thread walking tile by tile, entry by entry (a dozen dependent round trips: with long reads, where
// the rows take a few microseconds, that walk was the longest thing in the kernel).  Every lane gets all nine.
__device__ inline void gather_last9(const EmitParams& p, const ListView& lv, int lane, long long* nl)
{
    for (int q = 0; q < 9; ++q) nl[q] = 0;
    const int t = lv.n_tiles - 1 - lane;  // lane 0: the last tile
    const unsigned int cnt = (t >= 0) ? lv_count(lv, t) : 0u;
    unsigned int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    const unsigned int total = __shfl_sync(0xffffffffu, inc, 31);
    const unsigned int excl = inc - cnt;  // newlines in the tiles behind mine
    if (total >= 9u || lv.n_tiles <= 32) {
        // newline m (0 = newest) lives in the lane with excl <= m < excl + cnt
        long long mine = 0;
        const unsigned int m = (unsigned int)lane;
        unsigned int pos = 0;
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
            const unsigned int cnd = pos + (unsigned int)sft;
            const unsigned int b = __shfl_sync(0xffffffffu, excl, cnd & 31u);
            if (cnd < 32u && b <= m) pos = cnd;
        }
        const unsigned int e_pos = __shfl_sync(0xffffffffu, excl, pos);
        const unsigned int c_pos = __shfl_sync(0xffffffffu, cnt, pos);
        if (m < 9u && m < total && m - e_pos < c_pos) {
            unsigned int cls;
            lv_entry(lv, lv.n_tiles - 1 - int(pos), c_pos - 1u - (m - e_pos), &mine, &cls);
            mine = mine - p.mis + p.sentinel;
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) nl[8 - q] = __shfl_sync(0xffffffffu, mine, q);
        return;
    }
    // fewer than 9 newlines in the last 32 tiles (reads beyond 50 kb): the walk
    int have = 0;
    for (int tt = lv.n_tiles - 1; tt >= 0 && have < 9; --tt) {
        const unsigned int c = lv_count(lv, tt);
        for (unsigned int jj = c; jj > 0 && have < 9; --jj) {
            long long a;
            unsigned int cls;
            lv_entry(lv, tt, jj - 1, &a, &cls);
            nl[8 - have] = a - p.mis + p.sentinel;
            ++have;
        }
    }
}

// Classification of the end of the buffer, from the newline lists alone (runs concurrently with the
// row emission): number of records, status / posbuffer / offset of the first entrypos call that is
// not COMPLETE.  Stored in ParseState; fast4_finish turns it into the result header.
__device__ inline void fast4_tail(const EmitParams& p, const ListView& lv, unsigned long long M,
                                  unsigned long long gbase, const long long* nl)
{
    ParseState* st = p.st;
    const long long k0 = (long long)((gbase + 3) >> 2);  // global index of this shard's first record
    const long long L = (p.A > 0 ? p.A - p.mis : 0) + p.sentinel;
    const uint8_t* blob0 = p.base + p.mis - p.sentinel;  // address of blob[0]; virtual when sentinel
    long long pos[6] = {-1, -1, -1, -1, -1, -1};
    auto put = [&](long long n, long long resume, int status, int error) {
        st->tail_n = n;
        st->tail_resume = resume;
        for (int i = 0; i < 6; ++i) st->tail_pos[i] = pos[i];
        st->tail_status = status;
        st->tail_error = error;
    };
    if (p.sharded && !p.is_last) {
        // every owned record is closed inside shard + halo (else FQB_ERR_HALO is raised by the row
        // emission): the chain simply continues in the next shard
        const unsigned long long n_own = lv_count_before(lv, p.own_end);
        const long long n = (long long)((gbase + n_own + 3) >> 2) - k0;
        put(n, 0, ST_COMPLETE, (n + 1 > p.cap) ? FQB_ERR_CAPACITY : FQB_OK);
        return;
    }
    M += gbase;                       // global line count (the last shard sees the end of the stream)
    if (M == 0) {  // no visible newline at all: entrypos finds no "\n@"
        put(0, 0, ST_NO_HEAD_BEG, FQB_OK);
        return;
    }
    const long long Kg = (long long)((M - 1) >> 2);  // records of the whole stream closed by a newline
    const long long K = Kg - k0;                     // ... as a row of this shard's table
    const int m = int((M - 1) & 3ull);               // newlines after the last closing one
    // the last closed record is only COMPLETE if pos5 + 2 < L (src/_fastqandfurious.c:130), i.e. its
    // closing newline is not blob[L-2]
    const bool last_is_5 = (Kg >= 1 && m == 0 && blob0[L - 2] == '\n');
    if (K < 0 || (last_is_5 && K == 0)) {
        // the record the end-of-stream rules apply to starts in an earlier shard: the last shard is
        // shorter than a record
        put(0, 0, ST_NO_HEAD_BEG, FQB_ERR_HALO);
        return;
    }
    long long n = K - (last_is_5 ? 1 : 0);
    if (K + 1 > p.cap) {
        put(n, 0, ST_NO_HEAD_BEG, FQB_ERR_CAPACITY);
        return;
    }
    // the last (up to) 9 newlines of the buffer, blob coordinates: nl[8] = newest (gathered by the caller's warp).
    // Ranks 4(K-1) .. M-1 are at most 8 newlines when m <= 3.
    const long long* open_nl = nl + 9 - (m + 1);  // ranks 4K .. M-1 (global), the open record's newlines
    int status;
    long long resume = 0;
    if (last_is_5) {
        // ranks 4(K-1) .. 4K: the record that fails only the "pos5 + 2 < L" test
        const long long* r = nl + 9 - 5;
        pos[0] = r[0] + 1;
        pos[1] = r[1];
        pos[2] = r[1] + 1;
        pos[3] = r[2];
        pos[4] = r[3] + 1;
        pos[5] = -1;
        status = ST_NO_QUAL_END;
        if (n >= 1) resume = r[0] - 1;  // pos5 - 1 of record n-1 = K-2: its closing newline is rank 4(K-1)
    } else {
        status = classify_tail(blob0, L, open_nl, m + 1, pos);
        if (status == ST_COMPLETE) {  // last record without a newline after its quality string
            long long* row = p.table + K * 6;
            for (int i = 0; i < 6; ++i) row[i] = pos[i] + p.goff;
            n = K + 1;
            resume = pos[5] - 1;
            status = ST_NO_HEAD_BEG;  // the next call finds no further "\n@"
            for (int i = 0; i < 6; ++i) pos[i] = -1;
        } else if (n >= 1) {
            resume = open_nl[0] - 1;  // pos5 - 1 of record n-1: its closing newline is rank 4K
        }
    }
    put(n, resume, status, FQB_OK);
}

// result header: flags raised by the row emission + the stored tail classification
__device__ inline void fast4_finish(const EmitParams& p, unsigned long long M, unsigned long long gbase)
{
    ParseState* st = p.st;
    const long long k0 = (long long)((gbase + 3) >> 2);
    const unsigned long long fbi = *((volatile unsigned long long*)&st->first_bad_inv);
    const long long first_bad = fbi ? (long long)~fbi : -1;
    const long long n_lines = (long long)(M + ((p.sharded && !p.is_last) ? 0ull : gbase));
    const int err = *((volatile int*)&st->error);
    if (err) {
        write_result(p.res, 0, 0, ST_NO_HEAD_BEG, nullptr, FQB_PATH_FAST4, err, 0, n_lines, -1);
        return;
    }
    if (*((volatile int*)&st->fast_fail)) {
        st->need_general = 1;
        write_result(p.res, 0, 0, ST_NO_HEAD_BEG, nullptr, FQB_PATH_FAST4, p.sharded ? FQB_ERR_SHARD_GENERAL : FQB_OK, 1,
                     n_lines, first_bad);
        return;
    }
    long long pos[6];
    for (int i = 0; i < 6; ++i) pos[i] = *((volatile long long*)&st->tail_pos[i]);
    const int terr = *((volatile int*)&st->tail_error);
    const long long tn = *((volatile long long*)&st->tail_n);
    long long resume = *((volatile long long*)&st->tail_resume);
    int status = *((volatile int*)&st->tail_status);
    if (terr) {
        resume = 0;
        status = 0;  // ST_NO_HEAD_BEG
    }
    write_result(p.res, tn, resume, status, terr ? nullptr : pos, FQB_PATH_FAST4, terr, 0, n_lines, -1);
    p.res->reserved[0] = k0;
}

constexpr int EMIT_WIN = 256;  // list entries of a tile staged per warp (+4 of the following tile)
constexpr int EMIT_OWN = 24;  // long-read mode: tiles of a 32-tile group whose records the group emits
constexpr unsigned long long EMIT_SPARSE = 8;  // average lines per tile below which a warp takes 32 tiles at a time

__global__ void __launch_bounds__(256, 4) fq_emit_kernel(const EmitParams p)
{
    // launched with programmatic stream serialization: the CTAs may become resident while the scan kernel drains; nothing
    // the scan wrote is read before this returns (the scan has completed and its memory operations are visible)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (p.force_general) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            p.st->need_general = 1;
            write_result(p.res, 0, 0, ST_NO_HEAD_BEG, nullptr, FQB_PATH_FAST4, FQB_OK, 1, (long long)p.st->n_lines, -1);
        }
        return;
    }
    __shared__ __align__(16) unsigned short s_win[8][EMIT_WIN + 8];
    ListView lv = p.lv;
    lv.cls0 = *((volatile unsigned int*)&p.st->cls0);
    const unsigned long long M = *((volatile unsigned long long*)&p.st->n_lines);  // lines of this buffer
    bool peers_ok;
    const unsigned long long gbase = shard_line_base(p, &peers_ok);  // lines of the earlier shards
    if (!peers_ok) {  // uniform for the CTA
        if (threadIdx.x == 0) {
            p.st->error = FQB_ERR_PEER;
            if (blockIdx.x == 0)
                write_result(p.res, 0, 0, ST_NO_HEAD_BEG, nullptr, FQB_PATH_FAST4, FQB_ERR_PEER, 0, (long long)p.st->n_lines, -1);
        }
        return;
    }
    const long long k0 = (long long)((gbase + 3) >> 2);
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warp = int((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarps = int((gridDim.x * blockDim.x) >> 5);
    const bool dense_err = *((volatile int*)&p.st->error) != 0;
    unsigned short* win = s_win[wib];
    __shared__ __align__(16) uint4 s_stage[8][96];  // per warp: 32 rows of 48 bytes on their way to the table
    uint4* stage = s_stage[wib];
    // the end-of-buffer classification runs on one thread of the last CTA while the rows are written
    if (blockIdx.x == gridDim.x - 1 && wib == 7 && !dense_err) {
        long long nl9[9];
        gather_last9(p, lv, lane, nl9);
        if (lane == 31) {
            fast4_tail(p, lv, M, gbase, nl9);
            __threadfence();  // the classification is read by whichever CTA finishes last
        }
    }
    bool bad = false;
    unsigned long long bad_k = ~0ull;

    // Everything a tile needs from global memory -- the first EMIT_WIN entries of its list, the first four
    // of the next tile's, the count prefixes -- is one round of independent loads, issued one tile AHEAD:
    // the loads of a warp's next tile are in flight while it emits the rows of the current one.
    struct TileIn {
        unsigned int w[EMIT_WIN / 64];
        unsigned int wn, lp_t, lp_prev, lp_next, rq;
        unsigned long long rp;
    };
    // range (bq) and index inside the range (rq) of the tile being fetched, advanced without divisions
    const unsigned int T_u = (unsigned int)lv.T;
    unsigned int f_bq = (unsigned int)warp / T_u, f_rq = (unsigned int)warp % T_u;
    const unsigned int step_b = (unsigned int)nwarps / T_u, step_r = (unsigned int)nwarps % T_u;
    auto fetch = [&](int t, TileIn& in) {
        const unsigned int* own32 = reinterpret_cast<const unsigned int*>(lv.lists + (size_t)t * (unsigned int)lv.slot_cap);
#pragma unroll
        for (int k = 0; k < EMIT_WIN / 64; ++k) in.w[k] = __ldg(own32 + lane + 32 * k);
        const bool has_next = t + 1 < lv.n_tiles;
        in.wn = 0;
        if (has_next && lane < 2) in.wn = __ldg(own32 + (unsigned int)lv.slot_cap / 2 + lane);
        in.rq = f_rq;
        in.lp_t = lv.lprefix[t];
        in.lp_prev = f_rq ? lv.lprefix[t - 1] : 0u;
        in.lp_next = has_next ? lv.lprefix[t + 1] : 0u;
        in.rp = lv.rprefix[f_bq];
        f_bq += step_b;
        f_rq += step_r;
        if (f_rq >= T_u) {
            f_rq -= T_u;
            ++f_bq;
        }
    };
    // the equivalence conditions of one record and its row (positions of its five newlines, classes of the first three)
    auto finish = [&](long long k, long long s0, long long s1, long long s2, long long s3, long long s4, unsigned int c0,
                      unsigned int c1, unsigned int c2) {
        bool ok = (c0 == CLS_AT) && (c1 != CLS_NL) && (c2 == CLS_PLUS);
        const long long plus_len = s3 - s2;  // '+' line incl. its newline
        if (plus_len > 2 && plus_len != s1 - s0) ok = false;  // src/_fastqandfurious.c:109-117
        if (s4 - s3 != s2 - s1) ok = false;  // quality line as long as the sequence line
        if (k < p.cap) {
            const long long ob = p.out_bias;
            longlong2* row = reinterpret_cast<longlong2*>(p.table + k * 6);
            row[0] = make_longlong2(ob + s0 + 1, ob + s1);
            row[1] = make_longlong2(ob + s1 + 1, ob + s2);
            row[2] = make_longlong2(ob + s3 + 1, ob + s3 + s2 - s1);
        }
        if (!ok) {
            bad = true;
            if ((unsigned long long)k < bad_k) bad_k = (unsigned long long)k;
        }
    };

    // One record through the general route: its first newline is augmented entry jj of tile t (n entries, local rank
    // Bl of entry 0).  use_win: the tile's window is staged in shared memory (tile-at-a-time path).
    auto record = [&](int t, unsigned int jj, unsigned int n, unsigned int virt0, unsigned long long Bl, long long tb,
                      bool use_win, unsigned int n_next) {
        const unsigned long long B = gbase + Bl;
        const long long k = (long long)((B + jj) >> 2) - k0;  // row of this shard's table
        const bool closed = Bl + jj + 4 <= M - 1;           // all five newlines are in this buffer
        bool mine = true;
        if (p.sharded) {  // the shard that holds a record's first newline owns the record
            long long a0 = (long long)lv.mis - 1;
            if (jj >= virt0) {
                unsigned int cx;
                lv_entry(lv, t, jj, &a0, &cx);
            }
            mine = a0 < p.own_end;
            if (mine && !closed && !p.is_last) p.st->error = FQB_ERR_HALO;  // record runs past the halo
        }
        if (!(closed && mine)) return;
        long long s0, s1, s2, s3, s4;
        unsigned int c0, c1, c2;
        const unsigned int last = jj + 4;  // augmented index of the closing newline
        const bool in_win = use_win && (jj >= virt0) &&
                            ((last < n) ? (last - virt0 < EMIT_WIN) : (n - virt0 <= EMIT_WIN && last - n < 4 && last - n < n_next));
        if (in_win) {
            unsigned int e[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const unsigned int idx = jj + q;
                const bool own = idx < n;
                const unsigned int v = own ? win[idx - virt0] : win[EMIT_WIN + (idx - n)];
                e[q] = (v >> 2) + (own ? 0u : (unsigned int)lv.tile);
                if (q == 0) c0 = v & 3u;
                if (q == 1) c1 = v & 3u;
                if (q == 2) c2 = v & 3u;
            }
            s0 = tb + e[0];
            s1 = tb + e[1];
            s2 = tb + e[2];
            s3 = tb + e[3];
            s4 = tb + e[4];
        } else {  // virtual sentinel, long lists, records that run over several tiles
            LvCursor c = {t, jj, n};
            unsigned int cx;
            lv_entry(lv, c.t, c.jj, &s0, &c0);
            lv_next(lv, c);
            lv_entry(lv, c.t, c.jj, &s1, &c1);
            lv_next(lv, c);
            lv_entry(lv, c.t, c.jj, &s2, &c2);
            lv_next(lv, c);
            lv_entry(lv, c.t, c.jj, &s3, &cx);
            lv_next(lv, c);
            lv_entry(lv, c.t, c.jj, &s4, &cx);
        }
        finish(k, s0, s1, s2, s3, s4, c0, c1, c2);
    };

    // all records that start in tile t (one warp)
    auto emit_tile = [&](int t, const TileIn& in) {
        const unsigned int (&w)[EMIT_WIN / 64] = in.w;
        const unsigned int wn = in.wn;
        const bool has_next = t + 1 < lv.n_tiles;
        const unsigned int rq = in.rq;
        const unsigned int lp_t = in.lp_t, lp_prev = in.lp_prev, lp_next = in.lp_next;
        const unsigned long long rp = in.rp;
        const unsigned int virt0 = (t == 0) ? (unsigned int)lv.virt : 0u;
        const unsigned int n = lp_t - lp_prev + virt0;  // augmented count
        const unsigned long long Bl = (t == 0) ? 0ull : (unsigned long long)lv.virt + rp + lp_prev;  // local rank
        const unsigned long long B = gbase + Bl;                                                     // global rank
        const long long tb = (long long)t * lv.tile;
        if (n == 0) return;
        const unsigned int n_next = has_next ? (lp_next - ((rq + 1 == (unsigned int)lv.T) ? 0u : lp_t)) : 0u;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < EMIT_WIN / 64; ++k) reinterpret_cast<unsigned int*>(win)[lane + 32 * k] = w[k];
        if (lane < 2) reinterpret_cast<unsigned int*>(win)[EMIT_WIN / 2 + lane] = wn;
        __syncwarp();
        const unsigned int j0 = (4u - (unsigned int)(B & 3ull)) & 3u;  // first field-0 newline of the tile
        // ---- common case: every record that starts in this tile is closed inside this tile or by the first
        //      four newlines of the next one, all five list entries sit in the window (own entries followed
        //      by the neighbour's), the rows fit the table and the tile's output offsets share one high
        //      word: 32-bit tile-relative arithmetic, no per-record branches ----
        if (virt0 == 0 && n <= EMIT_WIN && n_next >= 4 && Bl + n + 4 <= M &&
            (!p.sharded || tb + lv.tile <= p.own_end)) {
            const unsigned int nrec = (n - j0 + 3u) >> 2;  // records whose first newline lies in this tile
            const long long k_first = (long long)((B + j0) >> 2) - k0;
            const unsigned long long obt = (unsigned long long)(p.out_bias + tb);
            const unsigned int ob_lo = (unsigned int)obt, ob_hi = (unsigned int)(obt >> 32);
            if (k_first + (long long)nrec <= p.cap && ob_lo <= 0xffffffffu - 2u * (unsigned int)lv.tile - 8u) {
                const unsigned int wv = __shfl_sync(0xffffffffu, wn, lane >> 1);
                if (lane < 4) win[n + lane] = (unsigned short)(wv >> ((lane & 1) * 16));
                __syncwarp();
                uint4* rows = reinterpret_cast<uint4*>(p.table + k_first * 6);
                const unsigned int tile_u = (unsigned int)lv.tile;
                for (unsigned int rb = 0; rb < nrec; rb += 32) {
                    // lanes past the last record redo it (their rows are not stored)
                    const unsigned int r = (rb + lane < nrec) ? rb + lane : nrec - 1;
                    const unsigned int jj = j0 + 4u * r;
                    const unsigned short* e = win + jj;
                    const unsigned int v0 = e[0], v1 = e[1], v2 = e[2], v3 = e[3], v4 = e[4];
                    const unsigned int r0 = v0 >> 2;  // always an own entry
                    const unsigned int r1 = (v1 >> 2) + (jj + 1 >= n ? tile_u : 0u);
                    const unsigned int r2 = (v2 >> 2) + (jj + 2 >= n ? tile_u : 0u);
                    const unsigned int r3 = (v3 >> 2) + (jj + 3 >= n ? tile_u : 0u);
                    const unsigned int r4 = (v4 >> 2) + (jj + 4 >= n ? tile_u : 0u);
                    bool ok = ((v0 & 3u) == CLS_AT) && ((v1 & 3u) != CLS_NL) && ((v2 & 3u) == CLS_PLUS);
                    const unsigned int plus_len = r3 - r2;                // '+' line incl. its newline
                    if (plus_len > 2 && plus_len != r1 - r0) ok = false;  // src/_fastqandfurious.c:109-117
                    if (r4 - r3 != r2 - r1) ok = false;                   // quality line as long as the sequence line
                    // the 32 rows of the warp are 1536 contiguous bytes of the table: transposed through
                    // shared memory so that every store instruction writes 512 contiguous bytes (whole
                    // sectors) instead of 32 half sectors 48 bytes apart
                    __syncwarp();
                    uint4* mine = stage + 3 * lane;  // 48-byte stride: conflict-free 16-byte accesses
                    mine[0] = make_uint4(ob_lo + r0 + 1, ob_hi, ob_lo + r1, ob_hi);
                    mine[1] = make_uint4(ob_lo + r1 + 1, ob_hi, ob_lo + r2, ob_hi);
                    mine[2] = make_uint4(ob_lo + r3 + 1, ob_hi, ob_lo + r3 + r2 - r1, ob_hi);
                    __syncwarp();
                    const unsigned int r_base = rb;
                    const unsigned int nv = (nrec - r_base < 32u ? nrec - r_base : 32u) * 3u;  // valid 16-byte units
                    uint4* out = rows + 3u * r_base;
#pragma unroll
                    for (int q = 0; q < 3; ++q)
                        if (q * 32u + lane < nv) out[q * 32 + lane] = stage[q * 32 + lane];
                    if (!ok) {
                        bad = true;
                        const unsigned long long k = (unsigned long long)(k_first + r);
                        if (k < bad_k) bad_k = k;
                    }
                }
                return;
            }
        }
        for (unsigned int jb = j0; jb < n; jb += 128) {
            const unsigned int jj = jb + 4u * lane;
            if (jj < n) record(t, jj, n, virt0, Bl, tb, true, n_next);
        }
    };

    // Long reads (fewer than EMIT_SPARSE lines per tile on average, e.g. 10 kb reads: 3 newlines per 16 KiB): a tile at
    // a time would spend its fixed cost (window loads, prefixes: ~290 instructions) on tiles that start less than one
    // record.  Instead a warp takes 32 consecutive tiles, a lane per tile: counts and ranks from the prefixes, then
    // every lane emits the (few) records that start in its tile through the general route.
    const bool sparse = M < (unsigned long long)lv.n_tiles * EMIT_SPARSE;
    if (sparse && !dense_err) {
        // a group hands out the records of its first EMIT_OWN tiles; the tiles behind them are only looked into (a record
        // that starts in the group then finds its five newlines inside it unless it spans more than 8 tiles)
        for (long long g0 = (long long)warp * EMIT_OWN; g0 < lv.n_tiles; g0 += (long long)nwarps * EMIT_OWN) {
            const int t = int(g0) + lane;
            unsigned int n = 0;
            unsigned long long Bl = 0;
            if (t < lv.n_tiles) {
                n = lv_count(lv, t);
                Bl = lv_base(lv, t);
            }
            if (__any_sync(0xffffffffu, lane < EMIT_OWN && n > 32u)) {  // a dense stretch inside a sparse buffer: tile at a time
                for (int q = 0; q < EMIT_OWN && g0 + q < lv.n_tiles; ++q) {
                    const int tq = int(g0) + q;
                    f_bq = (unsigned int)tq / T_u;
                    f_rq = (unsigned int)tq % T_u;
                    TileIn in;
                    fetch(tq, in);
                    emit_tile(tq, in);
                }
                continue;
            }
            const unsigned int virt0 = (t == 0) ? (unsigned int)lv.virt : 0u;
            const unsigned int j0 = (4u - (unsigned int)((gbase + Bl) & 3ull)) & 3u;
            // The five newlines of a record lie in the tiles behind its first one, nearly always inside the group:
            // the lane that owns each of them is found by a jump search over the lanes' ranks (shuffles, no memory),
            // so that the five list entries are five independent loads instead of a walk from tile to tile.
            const unsigned long long Bl0 = __shfl_sync(0xffffffffu, Bl, 0);
            const unsigned int rel = (t < lv.n_tiles) ? (unsigned int)(Bl - Bl0) : 0xffffffffu;  // ranks inside the group
            const unsigned int n_own = (lane < EMIT_OWN) ? n : 0u;  // my tile's records are mine only in the group's front part
            for (unsigned int jj = j0; __any_sync(0xffffffffu, jj < n_own); jj += 4) {
                const bool act = jj < n_own;
                unsigned int own_l[5], own_i[5];
                bool in_group = true;
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const unsigned int r = rel + jj + (unsigned int)q;  // rank of the record's q-th newline, group relative
                    unsigned int pos = (unsigned int)lane;
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) {
                        const unsigned int cnd = pos + (unsigned int)sft;
                        const unsigned int b = __shfl_sync(0xffffffffu, rel, cnd & 31u);
                        if (cnd < 32u && b <= r) pos = cnd;
                    }
                    const unsigned int bpos = __shfl_sync(0xffffffffu, rel, pos);
                    const unsigned int npos = __shfl_sync(0xffffffffu, n, pos);
                    own_l[q] = pos;
                    own_i[q] = r - bpos;
                    if (own_i[q] >= npos) in_group = false;  // behind the group's last line
                }
                if (!act) continue;
                const long long k = (long long)((gbase + Bl + jj) >> 2) - k0;
                if (!in_group || p.sharded) {  // (shards: the ownership and halo rules of the general route)
                    record(t, jj, n, virt0, Bl, (long long)t * lv.tile, false, 0u);
                    continue;
                }
                if (!(Bl + jj + 4 <= M - 1)) continue;  // not closed inside this buffer
                long long sp[5];
                unsigned int cl[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) lv_entry(lv, int(g0) + int(own_l[q]), own_i[q], &sp[q], &cl[q]);
                finish(k, sp[0], sp[1], sp[2], sp[3], sp[4], cl[0], cl[1], cl[2]);
            }
        }
    } else {
        TileIn ahead;
        if (warp < lv.n_tiles && !dense_err) fetch(warp, ahead);
        for (int t = warp; t < lv.n_tiles && !dense_err; t += nwarps) {
            const TileIn in = ahead;
            if (t + nwarps < lv.n_tiles) fetch(t + nwarps, ahead);
            emit_tile(t, in);
        }
    }
    if (bad) {
        atomicMax(&p.st->first_bad_inv, ~bad_k);
        p.st->fast_fail = 1;
        __threadfence();
    }

    // ---- last CTA done: tail classification + result header ----
    // Only the flags above (and the tail classification) are read by the last CTA; the threads that raised them have
    // fenced.  The rows need no fence: nobody reads them before the kernel ends (a fence by every thread made each
    // wait for its own row stores to drain: 7 % of the kernel's stall samples).
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(&p.st->emit_done, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    fast4_finish(p, M, gbase);
}


// ---- Phred decode of the fast path for mirrors the scan cannot write itself ----------------------------
// qual[i] = buf[i] + qual_add over every quality line (the arrayadd_b recipe, src/_fastqandfurious.c:180-182
// applied as in src/demo/benchmark.py:161-163) when the mirror is NOT congruent to the buffer modulo 16 (the
// congruent case is fused into the scan, fq_scan.cuh).  One CTA per tile, from the newline lists alone: the
// line after a newline is a quality line iff the number of newlines before it is a positive multiple of 4.
// Byte-exact: nothing outside the quality spans is written.  Results are only meaningful when the fast path
// accepts the buffer (the general path decodes on its own otherwise).
constexpr int DEC_THREADS = 128;

__global__ void __launch_bounds__(DEC_THREADS) fq_decode_kernel(const EmitParams p)
{
    if (p.force_general || !p.qual) return;
    if (*((volatile int*)&p.st->error) != 0) return;
    ListView lv = p.lv;
    lv.cls0 = *((volatile unsigned int*)&p.st->cls0);
    const int tid = threadIdx.x;
    int8_t* qbase = p.qual - p.mis;  // qbase[a] mirrors base[a]
    const unsigned int add = p.qual_add & 0xffu;
    // byte-exact copy of [b, e) by the whole CTA
    auto span = [&](long long b, long long e) {
        if (b < p.mis) b = p.mis;
        if (e > p.A) e = p.A;
        for (long long a = b + tid; a < e; a += DEC_THREADS) qbase[a] = int8_t(uint8_t(p.base[a] + add));
    };
    for (int t = blockIdx.x; t < lv.n_tiles; t += gridDim.x) {
        const unsigned int n = lv_count(lv, t);          // augmented count
        const unsigned long long B = lv_base(lv, t);     // rank of augmented entry 0
        const long long tb = (long long)t * lv.tile;
        if (n == 0) {
            if (B > 0 && (B & 3ull) == 0) span(tb, tb + lv.tile);  // the whole tile lies inside a quality line
            continue;
        }
        for (int jj = -1; jj < int(n); ++jj) {  // line after entry jj (jj = -1: the tile's first bytes); uniform
            const unsigned long long before = B + (unsigned long long)(jj + 1);
            if (before == 0 || (before & 3ull)) continue;
            long long b = tb, e = tb + lv.tile;
            unsigned int cx;
            if (jj >= 0) {
                lv_entry(lv, t, (unsigned int)jj, &b, &cx);
                b += 1;
            }
            if (jj + 1 < int(n)) lv_entry(lv, t, (unsigned int)(jj + 1), &e, &cx);
            span(b, e);
        }
    }
}

}  // namespace fqb
