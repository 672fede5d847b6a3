// fq_scan.cuh -- the single-pass newline-rank scan (the HBM-bound kernel of the parser).
//
// Persistent CTAs (cooperative launch: all co-resident) walk the byte buffer in TILE-sized steps:
//   1. tiles are staged global -> shared with 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) into a
//      STAGES-deep ring guarded by mbarriers, so several tiles per CTA are always in flight;
//   2. every thread reads its 16-byte chunks from shared memory (conflict-free LDS.128), turns them
//      into 16-bit newline masks with byte-SIMD arithmetic and counts them;
//   3. a warp-shuffle scan of the packed per-chunk counts plus a scan of the warp totals gives every
//      newline its rank inside the tile; the tile's rank base among ALL tiles comes from a decoupled
//      look-back over one 64-bit descriptor per tile (single pass over HBM: each input byte is read
//      exactly once);
//   4. MODE_FAST4: newline positions are compacted into shared memory and one thread per record
//      turns five consecutive newlines into a 6 x int64 table row (three 16-byte stores), checking
//      on the fly the conditions under which "newline rank mod 4" is provably identical to the
//      reference's sequential memmem/memchr chain (src/_fastqandfurious.c:62-136), see DESIGN.md;
//      with QUAL the bytes of every quality line (line index mod 4 == 0) are written, + qual_add,
//      to the mirror buffer from the staged tile -- the arrayadd_b recipe
//      (src/demo/benchmark.py:161-163, src/_fastqandfurious.c:180-182) without a second read;
//      MODE_LINES: every newline is written to the global line table (position | class of the
//      following byte) for the general path.
//
// Coordinates: a = byte index from `base` (the 16-byte aligned address at or below the caller's
// buffer); the caller's byte i is a = mis + i; the reference's blob index is a - mis + sentinel.
#pragma once
#include "fq_common.cuh"

namespace fqb {

constexpr int MODE_FAST4 = 0;
constexpr int MODE_LINES = 1;

struct ScanParams {
    const uint8_t* base;        // 16-byte aligned
    long long A;                // bytes addressable from base (mis + len)
    int mis;                    // leading bytes of `base` that are not part of the buffer
    int sentinel;               // 1: a virtual '\n' precedes the buffer
    long long out_bias;         // FAST4: emitted position = a + out_bias
    long long* table;           // FAST4: [cap][6]
    long long cap;
    unsigned long long* desc;   // [n_tiles] look-back descriptors, zero-initialised
    long long n_tiles;
    ParseState* st;
    unsigned long long* nlt;    // LINES: [max_lines] (blob position << 2) | class
    unsigned long long max_lines;
    int8_t* qual;               // FAST4 + QUAL: mirror of the caller's buffer (qual[i] <-> byte i)
    unsigned int qual_add4;     // qual_add replicated into 4 bytes
    int qual_vec;               // 1: (qual - mis) is 16-byte aligned, whole chunks go out as STG.128
};

template <int THREADS, int CPT, int STAGES>
struct ScanConfig {
    static constexpr int TILE = THREADS * CPT * 16;
    static constexpr int STAGE_BYTES = TILE + 128;  // 16 look-ahead bytes, padded to keep 128-B alignment
    static constexpr int NLCAP = TILE / 8;          // FAST4: newlines a tile may hold (mean line >= 8 bytes)
    static constexpr int NW = THREADS / 32;
    static constexpr size_t SMEM = size_t(STAGES) * STAGE_BYTES + size_t(NLCAP + 8) * 4;
};

// Field stores of one newline that could not be emitted as part of a whole row (its record has
// newlines in another tile, or is the still-open last record).
__device__ __forceinline__ void store_field(long long* table, long long cap, long long k, int f, long long pos)
{
    if (k >= cap) return;
    long long* row = table + k * 6;
    if (f == 0) {
        row[0] = pos + 1;  // '@' follows the closing newline of the previous record
    } else if (f == 1) {
        row[1] = pos;      // header '\n'
        row[2] = pos + 1;  // first sequence byte
    } else if (f == 2) {
        row[3] = pos;      // '\n' before '+'
    } else {
        row[4] = pos + 1;  // first quality byte
    }
}

// bits of `m` (newlines of one 16-byte chunk) -> bits of the bytes that lie on a quality line.
// cnt0 = number of visible newlines (incl. the sentinel) before the chunk.  A byte is on a quality
// line iff the count of newlines before it is a positive multiple of 4 and it is not a newline.
__device__ __forceinline__ uint32_t quality_bits(uint32_t m, unsigned long long cnt0)
{
    if (m == 0) return ((cnt0 & 3ull) == 0 && cnt0 != 0) ? 0xffffu : 0u;
    uint32_t q = 0, rest = m, start = 0;
    unsigned long long cnt = cnt0;
    for (;;) {
        const uint32_t e = rest ? uint32_t(__ffs(rest) - 1) : 16u;
        if ((cnt & 3ull) == 0 && cnt != 0) q |= ((1u << e) - 1u) & ~((1u << start) - 1u);
        if (!rest) break;
        rest &= rest - 1;
        start = e + 1;
        ++cnt;
    }
    return q;
}

template <int THREADS, int CPT, int STAGES, int MODE, bool QUAL>
__global__ void __launch_bounds__(THREADS) fq_scan_kernel(const ScanParams p)
{
    using Cfg = ScanConfig<THREADS, CPT, STAGES>;
    constexpr int TILE = Cfg::TILE;
    constexpr int NW = Cfg::NW;
    constexpr int NLCAP = Cfg::NLCAP;
    static_assert(CPT >= 1 && CPT <= 4, "packed 16-bit counts need CPT <= 4");
    static_assert(NW <= 32, "one warp scans the warp totals");
    static_assert(TILE + 16 < (1 << 24), "24-bit tile-local positions");

    if (MODE == MODE_LINES) {  // enqueued unconditionally, needed only when the fast path declined
        if (*((volatile const int*)&p.st->need_general) == 0 || *((volatile const int*)&p.st->error) != 0) return;
    }
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* nl_s = reinterpret_cast<uint32_t*>(smem + size_t(STAGES) * Cfg::STAGE_BYTES);
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ int s_wtot[32];
    __shared__ int s_wbase[32];
    __shared__ int s_nt;
    __shared__ unsigned long long s_base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long lo = p.mis;    // first visible byte
    const long long hi = p.A - 1;  // the last byte of the blob is never seen as a newline by the
                                   // reference (memchr windows exclude it; pairs need a 2nd byte)

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue_load = [&](long long it) {  // called by thread 0
        const long long t = (long long)blockIdx.x + it * gridDim.x;
        if (t >= p.n_tiles) return;
        const int s = int(it % STAGES);
        const long long tile_base = t * TILE;
        long long avail = p.A - tile_base;
        if (avail > TILE + 16) avail = TILE + 16;
        const uint32_t bytes = uint32_t(avail) & ~15u;
        if (bytes) {
            mbar_arrive_expect_tx(&full_bar[s], bytes);
            tma_load_1d(smem + size_t(s) * Cfg::STAGE_BYTES, p.base + tile_base, bytes, &full_bar[s]);
        }
    };

    if (tid == 0) {
        for (int it = 0; it < STAGES; ++it) issue_load(it);
    }

    for (long long it = 0;; ++it) {
        const long long t = (long long)blockIdx.x + it * gridDim.x;
        if (t >= p.n_tiles) break;
        const int s = int(it % STAGES);
        const uint32_t parity = uint32_t(it / STAGES) & 1u;
        uint8_t* tile = smem + size_t(s) * Cfg::STAGE_BYTES;
        const long long tile_base = t * TILE;
        long long avail = p.A - tile_base;
        if (avail > TILE + 16) avail = TILE + 16;
        const int full16 = int(avail) & ~15;
        const int rem = int(avail) - full16;
        // the last <16 bytes of the buffer are fetched with plain loads (a bulk copy moves whole
        // 16-byte units and must not run past the caller's allocation)
        if (tid < rem) tile[full16 + tid] = p.base[tile_base + full16 + tid];
        if (full16) mbar_wait(&full_bar[s], parity);
        if (rem) __syncthreads();

        // ---- phase A: newline masks and counts ----
        const bool edge = (tile_base < lo) || (tile_base + TILE > hi);  // first / last tiles only
        uint32_t masks[CPT];
        unsigned long long packed = 0;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int chunk = warp * (32 * CPT) + c * 32 + lane;
            const uint4 v = *reinterpret_cast<const uint4*>(tile + chunk * 16);
            uint32_t m = newline_mask16(v);
            if (edge) {
                const long long a0 = tile_base + chunk * 16;
                const long long b_lo = lo - a0, b_hi = hi - a0;
                uint32_t keep = 0xffffu;
                if (b_lo > 0) keep &= (b_lo >= 16) ? 0u : (0xffffu << int(b_lo));
                if (b_hi < 16) keep &= (b_hi <= 0) ? 0u : ((1u << int(b_hi)) - 1u);
                m &= keep;
            }
            masks[c] = m;
            packed += (unsigned long long)__popc(m) << (16 * c);
        }

        // ---- phase B: ranks inside the tile ----
        unsigned long long inc = packed;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long nb = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += nb;
        }
        const unsigned long long wtot_packed = __shfl_sync(0xffffffffu, inc, 31);
        const unsigned long long exc = inc - packed;
        int pre[CPT];  // rank of this thread's first newline of chunk c inside the warp
        int wtot = 0;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            pre[c] = wtot + int((exc >> (16 * c)) & 0xffffu);
            wtot += int((wtot_packed >> (16 * c)) & 0xffffu);
        }
        // the virtual sentinel newline occupies local index 0 of tile 0
        const int virt = (t == 0 && p.sentinel && p.A > p.mis) ? 1 : 0;

        if (lane == 0) s_wtot[warp] = wtot;
        __syncthreads();  // S1
        if (warp == 0) {
            const int v = (lane < NW) ? s_wtot[lane] : 0;
            int incw = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int nb = __shfl_up_sync(0xffffffffu, incw, o);
                if (lane >= o) incw += nb;
            }
            if (lane < NW) s_wbase[lane] = incw - v + virt;
            const int n_t = __shfl_sync(0xffffffffu, incw, 31) + virt;
            // publish this tile's count as early as possible, then look back
            unsigned long long excl = 0;
            if (t == 0) {
                if (lane == 0) st_relaxed_u64(p.desc, LB_INCL | (unsigned long long)n_t);
            } else {
                if (lane == 0) st_relaxed_u64(p.desc + t, LB_AGG | (unsigned long long)n_t);
                excl = lookback_exclusive(p.desc, t, lane);
                if (lane == 0) st_relaxed_u64(p.desc + t, LB_INCL | (excl + (unsigned long long)n_t));
            }
            if (lane == 0) {
                s_nt = n_t;
                s_base = excl;
            }
        }
        __syncthreads();  // S2
        const int n_t = s_nt;
        const unsigned long long B = s_base;  // rank of the tile's first newline (sentinel = rank 0)

        if (MODE == MODE_FAST4) {
            // ---- phase C: compact newline positions (+ class of the following byte) ----
            const bool fits = n_t <= NLCAP;
            if (fits) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    uint32_t m = masks[c];
                    const int chunk = warp * (32 * CPT) + c * 32 + lane;
                    int idx = s_wbase[warp] + pre[c];
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        const int lp = chunk * 16 + b;
                        nl_s[idx++] = uint32_t(lp + 1) | (classify(tile[lp + 1]) << 24);
                    }
                }
                if (virt && tid == 0) nl_s[0] = uint32_t(p.mis - 1 + 1) | (classify(tile[p.mis]) << 24);
            } else if (tid == 0) {
                p.st->fast_fail = 1;  // lines shorter than 8 bytes on average: not the fast path's business
            }

            // ---- fused Phred decode: bytes on quality lines, + qual_add, to the mirror buffer ----
            if (QUAL) {
                int8_t* qbase = p.qual - p.mis;  // qbase[a] mirrors base[a]
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const int chunk = warp * (32 * CPT) + c * 32 + lane;
                    const long long a0 = tile_base + chunk * 16;
                    const unsigned long long cnt0 = B + (unsigned long long)(s_wbase[warp] + pre[c]);
                    uint32_t q = quality_bits(masks[c], cnt0);
                    if (edge) {  // never write outside [mis, A)
                        const long long b_lo = lo - a0, b_hi = p.A - a0;
                        uint32_t keep = 0xffffu;
                        if (b_lo > 0) keep &= (b_lo >= 16) ? 0u : (0xffffu << int(b_lo));
                        if (b_hi < 16) keep &= (b_hi <= 0) ? 0u : ((1u << int(b_hi)) - 1u);
                        q &= keep;
                        // the blob's last byte is invisible as a newline but is not a quality byte either
                        if (b_hi >= 1 && b_hi <= 16 && tile[chunk * 16 + int(b_hi) - 1] == '\n')
                            q &= ~(1u << (int(b_hi) - 1));
                    }
                    if (q) {
                        const uint4 v = *reinterpret_cast<const uint4*>(tile + chunk * 16);
                        uint4 d;
                        d.x = __vadd4(v.x, p.qual_add4);
                        d.y = __vadd4(v.y, p.qual_add4);
                        d.z = __vadd4(v.z, p.qual_add4);
                        d.w = __vadd4(v.w, p.qual_add4);
                        if (q == 0xffffu && p.qual_vec) {
                            *reinterpret_cast<uint4*>(qbase + a0) = d;
                        } else {
                            const uint32_t w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                            for (int b = 0; b < 16; ++b)
                                if (q & (1u << b)) qbase[a0 + b] = int8_t((w[b >> 2] >> (8 * (b & 3))) & 0xffu);
                        }
                    }
                }
            }
            __syncthreads();  // S3: nl_s complete, the tile's bytes are no longer needed
            if (tid == 0) issue_load(it + STAGES);

            // ---- phase D: rows ----
            if (fits) {
                const int j0 = (4 - int(B & 3ull)) & 3;                 // first record-closing newline
                const int F = (n_t > j0) ? ((n_t - 1 - j0) >> 2) : 0;  // records with all 5 newlines here
                const long long obase = tile_base + p.out_bias;
                for (int q = tid; q < F; q += THREADS) {
                    const int j = j0 + 4 * q;
                    const long long k = (long long)((B + (unsigned long long)j) >> 2);
                    const uint32_t e0 = nl_s[j], e1 = nl_s[j + 1], e2 = nl_s[j + 2], e3 = nl_s[j + 3],
                                   e4 = nl_s[j + 4];
                    const int s0 = int(e0 & 0xffffffu) - 1, s1 = int(e1 & 0xffffffu) - 1,
                              s2 = int(e2 & 0xffffffu) - 1, s3 = int(e3 & 0xffffffu) - 1,
                              s4 = int(e4 & 0xffffffu) - 1;
                    bool ok = ((e0 >> 24) == CLS_AT) && ((e1 >> 24) != CLS_NL) && ((e2 >> 24) == CLS_PLUS);
                    const int plus_len = s3 - s2;  // '+' line incl. its newline
                    if (plus_len > 2 && plus_len != s1 - s0) ok = false;  // src/_fastqandfurious.c:109-117
                    if (s4 - s3 != s2 - s1) ok = false;  // quality line as long as the sequence line
                    if (k < p.cap) {
                        longlong2* row = reinterpret_cast<longlong2*>(p.table + k * 6);
                        row[0] = make_longlong2(obase + s0 + 1, obase + s1);
                        row[1] = make_longlong2(obase + s1 + 1, obase + s2);
                        row[2] = make_longlong2(obase + s3 + 1, obase + s3 + s2 - s1);
                    }
                    if (!ok) {
                        atomicMin(&p.st->first_bad, (unsigned long long)k);
                        p.st->fast_fail = 1;
                    }
                }
                // newlines whose record is not whole inside this tile: at most 3 at the head of the
                // tile and 4 at its end
                if (tid >= THREADS - 8) {
                    const int u = tid - (THREADS - 8);
                    const int nh = (j0 < n_t) ? j0 : n_t;
                    int j = n_t;
                    if (u < nh)
                        j = u;
                    else if (n_t > j0)
                        j = j0 + 4 * F + (u - nh);
                    if (j < n_t) {
                        const unsigned long long r = B + (unsigned long long)j;
                        const long long pos = obase + (long long)(nl_s[j] & 0xffffffu) - 1;
                        store_field(p.table, p.cap, (long long)(r >> 2), int(r & 3ull), pos);
                    }
                }
            }
        } else {
            // ---- MODE_LINES: global line table ----
            const long long blob_bias = (long long)p.sentinel - p.mis;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                uint32_t m = masks[c];
                const int chunk = warp * (32 * CPT) + c * 32 + lane;
                unsigned long long idx = B + (unsigned long long)(s_wbase[warp] + pre[c]);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const int lp = chunk * 16 + b;
                    if (idx < p.max_lines)
                        p.nlt[idx] = ((unsigned long long)(tile_base + lp + blob_bias) << 2) | classify(tile[lp + 1]);
                    ++idx;
                }
            }
            if (virt && tid == 0 && p.max_lines > 0) p.nlt[0] = classify(tile[p.mis]);  // blob position 0
            __syncthreads();  // S3
            if (tid == 0) issue_load(it + STAGES);
        }
    }
}

}  // namespace fqb
