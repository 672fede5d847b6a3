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

// One tile's list as a warp sees it: a lane per (augmented) entry, 32 at a time, the next 32 one round ahead.  Positions
// are kept relative to the tile (P = tileP + off, off >= -1: the virtual sentinel of tile 0 sits one byte before the
// buffer), ranks relative to the tile's first rank; 64-bit arithmetic only where a row is written.
struct FaTile {
    const unsigned short* src;  // raw entries of the tile
    unsigned int n, virt0;      // augmented count; 1: entry 0 is the virtual sentinel
    int voff;                   // ... its offset
    unsigned int vcls;
    int lim;                    // entry is a candidate iff class '>' and off < lim ("\n>" needs its second byte inside the blob)
};

__device__ __forceinline__ void fa_tile(const FastaParams& p, const ListView& lv, int t, unsigned int n, long long L, FaTile& ft,
                                        long long* tileP)
{
    ft.n = n;
    ft.virt0 = (t == 0 && lv.virt) ? 1u : 0u;
    ft.src = lv.lists + (size_t)t * (unsigned int)lv.slot_cap;
    ft.voff = lv.mis - 1;
    ft.vcls = lv.cls0;
    *tileP = (long long)t * lv.tile - p.mis + p.sentinel;  // blob position of the tile's byte 0
    const long long lim = L - 1 - *tileP;                    // P + 1 < L  <=>  off < L - 1 - tileP
    ft.lim = lim > 0x40000000ll ? 0x40000000 : (lim < -2 ? -2 : int(lim));
}

// entry jj of the tile: off (bits 31..2, signed) and class (bits 1..0); jj >= n: class 0 behind every real offset
__device__ __forceinline__ int fa_load(const FaTile& ft, unsigned int jj)
{
    if (jj >= ft.n) return 0x7ffffffc;
    if (jj < ft.virt0) return (ft.voff << 2) | int(ft.vcls);
    return int(__ldg(ft.src + (jj - ft.virt0)));
}

// on-chain flags of one round: c = my entry is a candidate, vm = valid lanes, carry_rel = tile-relative rank of the last
// non-candidate before the round (only its parity matters while it lies before the tile)
__device__ __forceinline__ bool fa_on(bool c, unsigned int cm, unsigned int vm, int r, int lane, int carry_rel)
{
    const unsigned int nc = ~cm & vm & ((1u << lane) - 1u);
    const int lastnc = nc ? (r - lane) + (31 - __clz(nc)) : carry_rel;
    return c && !((r - 1 - lastnc) & 1);  // an even number of consecutive candidates immediately before this rank
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
        long long tileP;
        fa_tile(p, lv, t, n, L, ft, &tileP);
        int carry_rel = -1, last_rel = -1;
        unsigned int lead = 0, n_on = 0;
        bool in_lead = true;
        int e_nxt = fa_load(ft, (unsigned int)lane);
        for (unsigned int j0 = 0; j0 < n; j0 += 32) {
            const int e = e_nxt;
            e_nxt = fa_load(ft, j0 + 32u + (unsigned int)lane);
            const bool valid = j0 + (unsigned int)lane < n;
            const bool c = valid && (e & 3) == int(CLS_AT) && (e >> 2) < ft.lim;
            const unsigned int cm = __ballot_sync(0xffffffffu, c), vm = __ballot_sync(0xffffffffu, valid);
            const bool on = fa_on(c, cm, vm, int(j0) + lane, lane, carry_rel);
            n_on += __popc(__ballot_sync(0xffffffffu, on));
            const unsigned int nc = ~cm & vm;
            if (in_lead) {
                lead += nc ? (unsigned int)(__ffs(nc) - 1) : (unsigned int)__popc(vm);
                in_lead = nc == 0;
            }
            if (nc) {
                last_rel = int(j0) + (31 - __clz(nc));
                carry_rel = last_rel;
            }
        }
        if (lane == 0) {
            p.tilemax[t] = last_rel >= 0 ? (long long)B + last_rel : -1;
            p.lead[t] = lead;
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
    for (int t = warp; t < lv.n_tiles; t += nwarps) {
        const unsigned int n = lv_count(lv, t);
        if (n == 0) continue;
        const unsigned long long B = lv_base(lv, t);
        FaTile ft;
        long long tileP;
        fa_tile(p, lv, t, n, L, ft, &tileP);
        // the last non-candidate rank before the tile, exact: only the parity of its distance matters
        int carry_rel = -1 - int(((long long)B - 1 - fa_carry(p, t)) & 1);
        long long kbase = p.tile_on[t];  // index of the tile's first on-chain rank
        int e_nxt = fa_load(ft, (unsigned int)lane);
        for (unsigned int j0 = 0; j0 < n; j0 += 32) {
            const int e = e_nxt;
            e_nxt = fa_load(ft, j0 + 32u + (unsigned int)lane);
            const unsigned int jj = j0 + (unsigned int)lane;
            const bool valid = jj < n;
            const int off = e >> 2;
            const bool c = valid && (e & 3) == int(CLS_AT) && off < ft.lim;
            const unsigned int cm = __ballot_sync(0xffffffffu, c), vm = __ballot_sync(0xffffffffu, valid);
            const bool on = fa_on(c, cm, vm, int(jj), lane, carry_rel);
            const unsigned int nc = ~cm & vm;
            if (nc) carry_rel = int(j0) + (31 - __clz(nc));
            const unsigned int om = __ballot_sync(0xffffffffu, on);
            // the line behind mine ends the header: my neighbour's entry, the next round's first one, or another tile's
            const int off_r = __shfl_down_sync(0xffffffffu, off, 1);
            const int off_n = __shfl_sync(0xffffffffu, e_nxt >> 2, 0);
            const long long k = kbase + __popc(om & ((1u << lane) - 1u));
            kbase += __popc(om);
            if (!on) continue;  // not on the chain
            const long long P = tileP + off;
            long long P1 = -1;  // end of the header line
            if (jj + 1 < n) {
                P1 = tileP + (lane < 31 ? off_r : off_n);
            } else {
                LvCursor cur = {t, jj, n};
                if (lv_next(lv, cur)) fa_is_cand(p, lv, cur.t, cur.jj, L, &P1);
            }
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
