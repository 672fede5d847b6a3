// fq_fasta.cuh -- FASTA on the newline-rank machinery (last row): the chain of
// entrypos_fasta calls (src/fastqandfurious.py:103-143), each starting at pos3 of the previous record.
//
// The scan kernel runs with '>' as class 1 and with the blob's last byte visible (bytes.find can match a
// single '\n' there; the pair "\n>" cannot start there).  Over the global newline sequence (rank r, blob
// position P[r]) a CANDIDATE is a newline followed by '>'.  One call anchored on candidate r uses newline
// r+1 as the end of the header and then searches "\n>" from P[r+1] + 1, i.e. among ranks >= r + 2: the
// chain takes the first candidate and then always the first candidate at least two ranks further.  Hence,
// inside a run of consecutive candidate ranks the chain takes every other one starting with the first, and
// it always enters a run at its first element (runs are separated by a non-candidate rank):
//     on(r) = cand(r) and the number of consecutive candidates immediately before r is even.
// Row of on-chain r: pos0 = P[r]+1, pos1 = P[r+1], pos2 = P[r+1]+1, pos3 = P[next on-chain rank]; the last
// on-chain rank is the call that is not COMPLETE (status 1, 2 or 3).  Steps: last non-candidate rank per tile and
// its running maximum (run lengths without walking the runs), the number of on-chain ranks per TILE, exclusive prefix
// sum over the tiles (record index of a tile's first on-chain rank; inside the tile the index comes from ballots),
// rows -- every on-chain rank also writes pos3 of the record before it.  Nothing is stored per rank: the rows kernel
// computes the flags again, with the exact carry.  (Round 1 kept an int64 flag per rank and ran the prefix sum over all
// of them: 155 MB read and written three times over for 19 M lines per GiB, 160 of the 690 us; the first round-2 form
// kept a flag byte per rank and 64-bit positions per lane: 228 us for the two passes over the lists.)
#pragma once
#include "fq_common.cuh"
#include "fq_consume.cuh"
#include "fq_emit.cuh"

namespace fqb {

struct FastaParams {
    const uint8_t* base;
    long long A;
    int mis;
    int sentinel;
    long long goff;
    long long* table;  // [cap][4]
    long long cap;
    ListView lv;
    ParseState* st;
    fqb_result* res;
    unsigned char* flags;  // [max_lines + 1]: unused since round 2 (a flag byte per rank; the rows kernel recomputes the flags)
    long long* tile_on;    // [n_tiles + 1]: on-chain ranks per tile, then (in place) their exclusive prefix sums
    unsigned long long max_lines;
    long long* tilemax;   // [n_tiles]: rank of the last newline that is not a candidate in the tiles of t's group up to t, -1: none
    long long* groupmax;  // [groups of FA_GROUP tiles]: the same over all tiles up to the end of the group
    unsigned int* lead;   // [n_tiles]: candidates at the start of the tile's list before its first non-candidate
};

__device__ __forceinline__ bool fa_is_cand(const FastaParams& p, const ListView& lv, int t, unsigned int jj, long long L,
                                           long long* pos_out)
{
    long long a;
    unsigned int cls;
    lv_entry(lv, t, jj, &a, &cls);
    const long long P = a - p.mis + p.sentinel;  // blob position
    if (pos_out) *pos_out = P;
    return cls == CLS_AT && P + 1 < L;  // class 1 is '>' here; "\n>" needs its second byte inside the blob
}

// "Consecutive candidates immediately before rank r" = r - 1 - (last non-candidate rank before r): linear work
// however long a run of header-only records is (a walk back from every rank would be quadratic in the run).
// Inside a tile that rank comes from ballots; across tiles from the running maximum of the tiles' last
// non-candidate ranks (fq_fa_groupscan_kernel / fq_fa_topscan_kernel), applied afterwards to the only flags that
// depend on it: the run of candidates a tile's list starts with (fq_fa_fixup_kernel).
__device__ __forceinline__ long long fa_last_noncand(unsigned int cand_mask, unsigned int valid_mask, long long rank0)
{
    const unsigned int nc = ~cand_mask & valid_mask;
    return nc ? rank0 + (31 - __clz(nc)) : -1;
}

// Running maximum in two levels: groups of FA_GROUP tiles are scanned in place by one CTA each (coalesced loads, the
// group's maximum goes to groupmax[g]), one CTA turns groupmax into its running maximum, and the flags kernel
// combines the two (fa_carry).  A single-CTA scan over all tiles cost 80 us per GiB in dependent loads.
constexpr int FA_GROUP = 256;

__device__ __forceinline__ long long fa_block_scan_max(long long v, long long* sh /* [blockDim.x] */)
{
    const int n = blockDim.x, i = threadIdx.x;
    sh[i] = v;
    __syncthreads();
    for (int d = 1; d < n; d <<= 1) {
        const long long o = (i >= d) ? sh[i - d] : -1;
        __syncthreads();
        v = max(v, o);
        sh[i] = v;
        __syncthreads();
    }
    return v;
}

__global__ void __launch_bounds__(FA_GROUP) fq_fa_groupscan_kernel(const FastaParams p)
{
    if (*((volatile int*)&p.st->error) != 0) return;
    __shared__ long long sh[FA_GROUP];
    const int t = blockIdx.x * FA_GROUP + threadIdx.x;
    const long long v = fa_block_scan_max(t < p.lv.n_tiles ? p.tilemax[t] : -1, sh);
    if (t < p.lv.n_tiles) p.tilemax[t] = v;
    if (threadIdx.x == FA_GROUP - 1) p.groupmax[blockIdx.x] = v;
}

__global__ void __launch_bounds__(1024) fq_fa_topscan_kernel(const FastaParams p)
{
    if (*((volatile int*)&p.st->error) != 0) return;
    __shared__ long long sh[1024];
    const int n_groups = (p.lv.n_tiles + FA_GROUP - 1) / FA_GROUP;
    long long carry = -1;
    for (int g0 = 0; g0 < n_groups; g0 += 1024) {
        const int g = g0 + threadIdx.x;
        const long long v = max(carry, fa_block_scan_max(g < n_groups ? p.groupmax[g] : -1, sh));
        if (g < n_groups) p.groupmax[g] = v;
        carry = max(carry, sh[1023]);
        __syncthreads();
    }
}

// rank of the last non-candidate newline in tiles 0 .. t-1 (-1: none)
__device__ __forceinline__ long long fa_carry(const FastaParams& p, int t)
{
    if (t == 0) return -1;
    const int g = t / FA_GROUP;
    long long c = (t % FA_GROUP) ? p.tilemax[t - 1] : -1;
    if (g > 0) c = max(c, p.groupmax[g - 1]);
    return c;
}

// One tile's list as a warp sees it: EIGHT raw entries per lane (one 16-byte load), 256 per round; candidate, non-candidate
// and on-chain flags are 8-bit masks per lane, the run parities come from bit arithmetic instead of a test per line.
// Positions are relative to the tile (P = tileP + offset); the virtual sentinel of tile 0 (augmented entry 0, one byte
// before the buffer) is handled apart, before the raw entries.
struct FaTile {
    const unsigned short* src;  // raw entries of the tile (16-byte aligned slot)
    int nraw;                   // raw entries
    int virt0;                  // 1: the virtual sentinel precedes them
    int lim;                    // a '>'-class entry is a candidate iff offset < lim ("\n>" needs its second byte inside the blob)
    bool lim_active;            // only the last tile(s): lim inside the tile
    long long tileP;            // blob position of the tile's byte 0
};

__device__ __forceinline__ void fa_tile(const FastaParams& p, const ListView& lv, int t, unsigned int n, long long L, FaTile& ft)
{
    ft.virt0 = (t == 0 && lv.virt) ? 1 : 0;
    ft.nraw = int(n) - ft.virt0;
    ft.src = lv.lists + (size_t)t * (unsigned int)lv.slot_cap;
    ft.tileP = (long long)t * lv.tile - p.mis + p.sentinel;
    const long long lim = L - 1 - ft.tileP;  // P + 1 < L  <=>  offset < L - 1 - tileP
    ft.lim = lim > 0x40000000ll ? 0x40000000 : (lim < 0 ? 0 : int(lim));
    ft.lim_active = lim < (long long)lv.tile;
}

// my eight entries of round rr (zero behind the end of the list: stale bytes of the slot are never read as entries)
__device__ __forceinline__ uint4 fa_load8(const FaTile& ft, int rr, int lane)
{
    const int raw0 = rr * 256 + lane * 8;
    return raw0 < ft.nraw ? __ldg(reinterpret_cast<const uint4*>(ft.src + raw0)) : make_uint4(0u, 0u, 0u, 0u);
}
__device__ __forceinline__ unsigned int fa_entry(const uint4& x, int k)  // entry k (0..7) of a vector
{
    const unsigned long long lo = ((unsigned long long)x.y << 32) | x.x, hi = ((unsigned long long)x.w << 32) | x.z;
    return (unsigned int)(((k < 4 ? lo : hi) >> (16 * (k & 3))) & 0xffffu);
}

// Flags of my eight entries.  run_in = consecutive candidates immediately before the round's first entry (uniform; only
// its parity matters).  on(r) = cand(r) and an even number of consecutive candidates immediately before r: inside a run
// of candidates every other one, starting with the first.  The run my group starts in gets its parity from the nearest
// non-candidate below (lower lanes: one ballot + one shuffle); the runs that start inside the group from the carry
// trick -- adding a run's first bit to the mask clears exactly that run, so `rest & ~(rest + even_starts)` are the runs
// that start on an even bit, and an on-chain bit has the parity of its run's start.
__device__ __forceinline__ void fa_flags8(const FaTile& ft, const uint4& x, int rr, int lane, int run_in, unsigned int& cand8,
                                          unsigned int& nc8, unsigned int& on8, unsigned int& hb)
{
    const int nv = ft.nraw - (rr * 256 + lane * 8);
    const unsigned int valid8 = nv >= 8 ? 0xffu : (nv <= 0 ? 0u : ((1u << nv) - 1u));
    const unsigned int k1 = 0x00010001u;
    const unsigned int xs[4] = {x.x, x.y, x.z, x.w};
    unsigned int a = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) a |= (xs[q] & ~(xs[q] >> 1) & k1) << (2 * q);  // class 01 ('>')
    unsigned int at8 = (a | (a >> 15)) & 0xffu;
    if (ft.lim_active) {  // the end of the buffer: "\n>" needs its second byte
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (int(fa_entry(x, k) >> 2) >= ft.lim) at8 &= ~(1u << k);
    }
    cand8 = at8 & valid8;
    nc8 = ~cand8 & valid8;
    hb = __ballot_sync(0xffffffffu, nc8 != 0u);
    const unsigned int below = hb & ((1u << lane) - 1u);
    const int ln = below ? 31 - __clz(below) : 0;
    const unsigned int ncl = __shfl_sync(0xffffffffu, nc8, ln);
    // consecutive candidates immediately before my entry 0
    const int cnt = below ? (8 * lane - 1 - (8 * ln + (31 - __clz(ncl)))) : (8 * lane + run_in);
    const unsigned int lead_run = cand8 & ~(cand8 + 1u);  // the run my group starts in (trailing ones)
    const unsigned int rest = cand8 & ~lead_run;
    const unsigned int starts = rest & ~(rest << 1);
    const unsigned int runs_e = rest & ~(rest + (starts & 0x55u));
    on8 = (lead_run & ((cnt & 1) ? 0xAAu : 0x55u)) | (runs_e & 0x55u) | (rest & ~runs_e & 0xAAu);
}

// after a round: consecutive candidates at its end (for the next round) and its last non-candidate (raw index, -1: none)
__device__ __forceinline__ void fa_round_end(const FaTile& ft, int rr, unsigned int nc8, unsigned int hb, int& run_in, int& last_nc)
{
    const int lines = (ft.nraw - rr * 256 < 256) ? ft.nraw - rr * 256 : 256;
    if (hb) {
        const int lt = 31 - __clz(hb);
        const unsigned int nct = __shfl_sync(0xffffffffu, nc8, lt);
        const int pos = 8 * lt + (31 - __clz(nct));
        run_in = lines - 1 - pos;
        last_nc = rr * 256 + pos;
    } else {
        run_in += lines;
    }
}

// ---- F1: on-chain ranks, last non-candidate rank and leading run of candidates of every tile ----
// The rank before the tile is taken to be a non-candidate; fq_fa_fixup_kernel corrects the count of the (few) tiles
// whose LEADING run of candidates (the only flags that depend on earlier tiles) starts with the other parity, once the
// running maximum over the earlier tiles is known.  The rows kernel computes the flags again with the exact carry:
// no flag per rank is stored (rounds 1 and 2 wrote and re-read a byte per rank).
__global__ void __launch_bounds__(256) fq_fa_count_kernel(const FastaParams p)
{
    if (*((volatile int*)&p.st->error) != 0) return;
    ListView lv = p.lv;
    lv.cls0 = *((volatile unsigned int*)&p.st->cls0);
    const unsigned long long M = *((volatile unsigned long long*)&p.st->n_lines);
    if (M > p.max_lines) {
        if (threadIdx.x == 0 && blockIdx.x == 0) p.st->error = FQB_ERR_WORKSPACE;
        return;
    }
    const long long L = (p.A > 0 ? p.A - p.mis : 0) + p.sentinel;
    const int lane = threadIdx.x & 31;
    const int warp = int((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarps = int((gridDim.x * blockDim.x) >> 5);
    for (int t = warp; t < lv.n_tiles; t += nwarps) {
        const unsigned int n = lv_count(lv, t);
        const unsigned long long B = lv_base(lv, t);
        FaTile ft;
        fa_tile(p, lv, t, n, L, ft);
        // the virtual sentinel (rank 0, blob position 0): a candidate iff the buffer starts with '>' (and holds a 2nd byte)
        const bool cand_v = ft.virt0 && lv.cls0 == CLS_AT && L > 1;
        int run_in = cand_v ? 1 : 0, last_nc = -1, lead = 0;
        unsigned int n_on = cand_v ? 1u : 0u;
        bool in_lead = true;
        uint4 x_nxt = fa_load8(ft, 0, lane);
        for (int rr = 0; rr * 256 < ft.nraw; ++rr) {
            const uint4 x = x_nxt;
            x_nxt = fa_load8(ft, rr + 1, lane);
            unsigned int cand8, nc8, on8, hb;
            fa_flags8(ft, x, rr, lane, run_in, cand8, nc8, on8, hb);
            n_on += (unsigned int)__reduce_add_sync(0xffffffffu, __popc(on8));
            if (in_lead) {
                if (hb) {
                    const int lf = __ffs(hb) - 1;
                    const unsigned int ncf = __shfl_sync(0xffffffffu, nc8, lf);
                    lead += 8 * lf + __ffs(ncf) - 1;
                    in_lead = false;
                } else {
                    lead += (ft.nraw - rr * 256 < 256) ? ft.nraw - rr * 256 : 256;
                }
            }
            fa_round_end(ft, rr, nc8, hb, run_in, last_nc);
        }
        if (lane == 0) {
            long long tm = -1;
            if (last_nc >= 0) tm = (long long)B + ft.virt0 + last_nc;
            else if (ft.virt0 && !cand_v) tm = 0;
            p.tilemax[t] = tm;
            p.lead[t] = (unsigned int)lead;  // (tile 0 is never fixed up)
            p.tile_on[t] = n_on;
        }
    }
}

// ---- F2: the counts of the tiles whose leading run starts with the other parity ----
__global__ void __launch_bounds__(256) fq_fa_fixup_kernel(const FastaParams p)
{
    if (*((volatile int*)&p.st->error) != 0) return;
    const ListView& lv = p.lv;
    const int t = int(blockIdx.x * blockDim.x + threadIdx.x);
    if (t <= 0 || t >= lv.n_tiles) return;
    const unsigned int n_lead = p.lead[t];
    if ((n_lead & 1u) == 0) return;  // an even run holds as many on-chain ranks either way
    const unsigned long long B = lv_base(lv, t);
    // the run's flags alternate 1 0 1 0 ...: started one rank later, an odd run holds one on-chain rank less
    if ((((long long)B - 1 - fa_carry(p, t)) & 1) != 0) p.tile_on[t] -= 1;
}

// ---- F3: rows ----
__global__ void __launch_bounds__(256) fq_fa_rows_kernel(const FastaParams p)
{
    if (*((volatile int*)&p.st->error) != 0) return;
    ListView lv = p.lv;
    lv.cls0 = *((volatile unsigned int*)&p.st->cls0);
    const unsigned long long M = *((volatile unsigned long long*)&p.st->n_lines);
    const long long L = (p.A > 0 ? p.A - p.mis : 0) + p.sentinel;
    const long long total = p.tile_on[lv.n_tiles];  // on-chain ranks = calls of the chain that found a header
    const int lane = threadIdx.x & 31;
    const int warp = int((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarps = int((gridDim.x * blockDim.x) >> 5);
    // on-chain rank number k at blob position P, its header line ends at P1 (-1: no further newline)
    auto emit_row = [&](long long k, long long P, long long P1) {
        if (k >= 1 && k - 1 < p.cap) p.table[(k - 1) * 4 + 3] = P + p.goff;  // closes the record before
        if (k + 1 < total) {  // COMPLETE: a later on-chain rank exists (so do P1 and the byte after it)
            if (k < p.cap) {
                long long* row = p.table + k * 4;
                row[0] = P + 1 + p.goff;
                row[1] = P1 + p.goff;
                row[2] = P1 + 1 + p.goff;
            }
        } else {  // the call that is not COMPLETE: src/fastqandfurious.py:120-139
            long long pos[6] = {P + 1, -1, -1, -1, -1, -1};
            int status;
            if (P1 < 0) {
                status = FQB_MISSING_SEQHEADER_END;
            } else {
                pos[1] = P1;
                if (P1 + 1 >= L) {
                    status = FQB_MISSING_SEQ_BEG;
                } else {
                    pos[2] = P1 + 1;
                    pos[3] = (p.base[p.A - 1] == '\n') ? L - 1 : L;
                    status = FQB_MISSING_SEQ_END;
                }
            }
            const long long n_rec = total - 1;
            write_result(p.res, n_rec, n_rec >= 1 ? P : 0, status, pos, FQB_PATH_FAST4,
                         (total > p.cap) ? FQB_ERR_CAPACITY : FQB_OK, 0, (long long)M, -1);
        }
    };
    for (int t = warp; t < lv.n_tiles; t += nwarps) {
        const unsigned int n = lv_count(lv, t);
        if (n == 0) continue;
        const unsigned long long B = lv_base(lv, t);
        FaTile ft;
        fa_tile(p, lv, t, n, L, ft);
        long long kbase = p.tile_on[t];  // index of the tile's first on-chain rank
        // consecutive candidates immediately before the tile's first rank, exact now (only the parity matters)
        int run_in = (t > 0) ? int(((long long)B - 1 - fa_carry(p, t)) & 1) : 0, last_nc = -1;
        // position of the line behind augmented entry jj when it lies in another tile (or nowhere: -1)
        auto next_elsewhere = [&](unsigned int jj) -> long long {
            long long P1 = -1;
            LvCursor cur = {t, jj, n};
            if (lv_next(lv, cur)) fa_is_cand(p, lv, cur.t, cur.jj, L, &P1);
            return P1;
        };
        if (ft.virt0) {  // the virtual sentinel: blob position 0, on the chain iff it is a candidate (nothing before it)
            const bool cand_v = lv.cls0 == CLS_AT && L > 1;
            if (cand_v) {
                if (lane == 0)
                    emit_row(kbase, 0, ft.nraw > 0 ? ft.tileP + (long long)(__ldg(ft.src) >> 2) : next_elsewhere(0u));
                kbase += 1;
                run_in = 1;
            }
        }
        uint4 x_nxt = fa_load8(ft, 0, lane);
        for (int rr = 0; rr * 256 < ft.nraw; ++rr) {
            const uint4 x = x_nxt;
            x_nxt = fa_load8(ft, rr + 1, lane);
            unsigned int cand8, nc8, on8, hb;
            fa_flags8(ft, x, rr, lane, run_in, cand8, nc8, on8, hb);
            fa_round_end(ft, rr, nc8, hb, run_in, last_nc);
            const int mine = __popc(on8);
            int inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            const unsigned int nx0 = __shfl_down_sync(0xffffffffu, x.x & 0xffffu, 1);  // the entry behind my last one
            long long k = kbase + inc - mine;
            kbase += __shfl_sync(0xffffffffu, inc, 31);
            const int raw0 = rr * 256 + lane * 8;
            for (unsigned int m = on8; m; m &= m - 1u, ++k) {
                const int b = __ffs(m) - 1, raw = raw0 + b;
                const long long P = ft.tileP + (long long)(fa_entry(x, b) >> 2);
                long long P1;
                if (raw + 1 < ft.nraw) {
                    const unsigned int en = b < 7 ? fa_entry(x, b + 1) : (lane < 31 ? nx0 : (unsigned int)__ldg(ft.src + raw + 1));
                    P1 = ft.tileP + (long long)(en >> 2);
                } else {
                    P1 = next_elsewhere((unsigned int)(raw + ft.virt0));
                }
                emit_row(k, P, P1);
            }
        }
    }
}

// ---- F4: result header when there is no call to describe (no "\n>" at all) or an error ----
__global__ void fq_fa_result_kernel(const FastaParams p)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int err = *((volatile int*)&p.st->error);
    const unsigned long long M = *((volatile unsigned long long*)&p.st->n_lines);
    if (err) {
        write_result(p.res, 0, 0, FQB_MISSING_SEQHEADER_BEGIN, nullptr, FQB_PATH_FAST4, err, 0, (long long)M, -1);
        return;
    }
    if (p.tile_on[p.lv.n_tiles] == 0)
        write_result(p.res, 0, 0, FQB_MISSING_SEQHEADER_BEGIN, nullptr, FQB_PATH_FAST4, FQB_OK, 0, (long long)M, -1);
}

}  // namespace fqb
