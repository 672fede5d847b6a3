// fq_scan.cuh -- the newline scan: the one kernel that reads the input, the HBM-bound kernel of the
// parser.  Every input byte is read from HBM exactly once, by this kernel.
//
// Persistent CTAs, each owning a CONTIGUOUS range of TILE-byte tiles (no inter-CTA dependency on the
// hot path -- an earlier single-pass version with a decoupled look-back between co-resident CTAs
// was latency bound at 0.9 TB/s, see profiles/r01_v0_*):
//   1. tiles are staged global -> shared with 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) into a
//      STAGES-deep ring guarded by mbarriers, so several tiles per CTA are always in flight;
//   2. every thread reads its 16-byte chunks from shared memory (conflict-free LDS.128) and tests them for a
//      possible newline with one subtraction per word (no false negatives);
//   3. the (few) chunks that pass are queued per warp in position order (ballot + popc); one lane per queued
//      chunk then computes its exact 16-bit newline mask with byte-SIMD arithmetic; the counts of the queued chunks
//      plus the sum of the warp totals (one __syncthreads per tile) give every newline its index inside the tile;
//   4. each newline is written to the tile's slot of the global newline list as a 16-bit entry
//      (offset in tile << 2 | class of the following byte: '@', '+', '\n', other) and the running
//      count of the CTA's range to lprefix[tile];
//   5. the last CTA to finish turns the per-range totals into exclusive prefixes (rprefix).
// The lists are ~2.5 % of the input size for 150 bp reads; everything downstream (row emission,
// tail classification, the general path's line table) works from them and never re-reads the input
// except for Phred decoding.
//
// Coordinates: a = byte index from `base` (the 16-byte aligned address at or below the caller's
// buffer); the caller's byte i is a = mis + i; the reference's blob index is a - mis + sentinel.
#pragma once
#include "fq_common.cuh"

namespace fqb {

struct ScanParams {
    const uint8_t* base;          // 16-byte aligned
    long long A;                  // bytes addressable from base (mis + len), 0 for an empty buffer
    int mis;                      // leading bytes of `base` that are not part of the buffer
    int sentinel;                 // 1: a virtual '\n' precedes the buffer
    unsigned short* lists;        // [n_tiles][slot_cap]
    unsigned int* lprefix;        // [n_tiles]
    unsigned int* rangetot;       // [gridDim.x]
    unsigned long long* rprefix;  // [gridDim.x + 1]
    long long n_tiles;
    long long T;                  // tiles per CTA
    int slot_cap;
    ParseState* st;
    // optional Phred mirror fused into the scan (DEC kernels): qual[a] = base[a] + add for EVERY byte a of
    // the buffer.  The bytes inside quality spans are what fqb_parse promises, the rest of the mirror is
    // unspecified -- and a second pass that picks the quality lines out would fetch the whole input from
    // DRAM again (the gaps between 150-byte lines are shorter than the DRAM fetch granularity).
    int8_t* qual;        // mirror of `base` (qual[a] belongs to base[a]), 16-byte aligned; nullptr: off
    unsigned int add4;   // the byte to add, replicated
    // line classes: the byte after a newline is class 1 if it equals cls1 ('@'; '>' for FASTA), class 2 if it
    // equals cls2 ('+'), class 3 if it is a newline.  last_visible: a newline in the blob's last byte counts
    // (bytes.find semantics of entrypos_fasta); the C entrypos can never see it.
    unsigned char cls1, cls2;
    int last_visible;
};

// Sharded parse, fused exchange: what the LAST CTA of a shard's scan hands on once the count prefixes are final --
// the number of lines at offsets below own_end (the shard's contribution to the line rank of the shards behind
// it), stored locally and as {count, epoch} into the memory of every later shard (peer-mapped pointers, NVLink),
// and the "my bytes of the next parse are in place" signal to the left neighbour.  One kernel instead of three
// (scan, fq_own_lines_kernel, fq_signal_ready_kernel).
struct ShardTail {
    long long own_end;                // byte index from ScanParams::base behind the shard's own bytes
    unsigned long long* own_lines;    // local result (may be nullptr)
    unsigned long long* pub[16];      // slot of this shard in the memory of later shards
    int n_pub;
    unsigned long long epoch;
    unsigned long long* ready_left;   // nullptr: no signal
    unsigned long long ready_epoch;
};
struct NoTail {};
template <bool SHARD>
struct TailOf {
    using type = NoTail;
};
template <>
struct TailOf<true> {
    using type = ShardTail;
};

template <int THREADS, int CPT, int STAGES>
struct ScanConfig {
    static constexpr int TILE = THREADS * CPT * 16;     // bytes per loop iteration ("super tile")
    static constexpr int LT = TILE > 16384 ? 16384 : TILE;  // bytes per LIST tile (16-bit entries: 14-bit offsets)
    static constexpr int TPI = TILE / LT;               // list tiles per iteration
    static constexpr int STAGE_BYTES = TILE + 128;  // 16 look-ahead bytes, padded to keep 128-B alignment
    static constexpr int NW = THREADS / 32;
    static constexpr int WPL = NW / TPI;                // warps per list tile
    static constexpr int QCAP = 32 * CPT + 1;       // one queue entry per 16-byte chunk of the warp + a dummy word
    static constexpr size_t SMEM = size_t(STAGES) * STAGE_BYTES + size_t(NW) * QCAP * 4;
};

// (w & k_and) ^ k_xor in one LOP3
__device__ __forceinline__ uint32_t and_xor(uint32_t w, uint32_t k_and, uint32_t k_xor)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(r) : "r"(w), "r"(k_and), "r"(k_xor));
    return r;
}

// 0x80 in every byte of w that equals '\n' (exact, no cross-byte carries)
__device__ __forceinline__ uint32_t newline_flags(uint32_t w)
{
    const uint32_t k7 = 0x7f7f7f7fu;
    return ~((and_xor(w, k7, 0x0a0a0a0au) + k7) | w) & 0x80808080u;
}

// flags of the four words of a 16-byte chunk -> bit i set iff byte i is a newline.  Two words share one
// multiply: A = (f0 >> 4) | f1 has word 0 at bits 3,11,19,27 and word 1 at bits 7,15,23,31; times
// 2^21 + 2^14 + 2^7 + 1 every partial product lands on its own bit and bits 24..31 collect
// [w0b0 w0b1 w0b2 w0b3 w1b0 w1b1 w1b2 w1b3] (verified exhaustively in tests/test_algo_model.py).
__device__ __forceinline__ uint32_t gather_flags16(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3)
{
    const uint32_t kg = 0x00204081u;
    const uint32_t a = ((f0 >> 4) | f1) * kg, b = ((f2 >> 4) | f3) * kg;
    return (a >> 24) | ((b >> 16) & 0xff00u);
}

__device__ __forceinline__ bool mbar_try_wait_s(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}

// a | b | c in one LOP3
__device__ __forceinline__ uint32_t or3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xfe;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// Rows of one tile, phase 1: every 16-byte chunk that MAY hold a newline goes to the warp's queue (shared-space
// byte address `q0`), in position order, as its chunk index inside the warp's part of the tile (lane | row << 5).
// The test is one subtraction per word: a byte equal to '\n' (0x0a) leaves bit 7 of its byte of w - 0x0b0b0b0b
// set whatever borrow arrives from the bytes below it (0x0a - 0x0b - {0, 1} = 0xff / 0xfe): no newline is missed.
// Other bytes that set the bit (values below 0x0b or above 0x8a, a 0x0b behind a smaller byte -- none of them
// FASTQ text) only cost a queue entry that turns out empty.  The exact position mask is computed for the queued
// chunks alone (queue_block: about a fifth of all chunks at 150 bp), one lane per chunk, instead of by every lane
// for every chunk.  `my_chunk` is the shared-space address of my chunk of row 0.  Returns the number of queued
// chunks.  ~18 instructions per row of 512 bytes (exact flags + gather in the row loop: 33).
template <int CPT, bool DEC>
__device__ __forceinline__ int scan_rows(uint32_t my_chunk, uint32_t q0, uint32_t lt_mask, uint32_t lane, int8_t* qrow,
                                         unsigned int add4)
{
    uint32_t qa = q0;
    const uint32_t q_dummy = q0 + 4u * 32u * CPT;  // word 32*CPT of the warp's queue
    const uint32_t kb = 0x0b0b0b0bu;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const uint4 v = lds_128(my_chunk + c * 512);
        if (DEC && qrow) {  // block uniform: the Phred mirror of my chunk (qrow: my chunk of row 0)
            uint4 d;
            d.x = __vadd4(v.x, add4);
            d.y = __vadd4(v.y, add4);
            d.z = __vadd4(v.z, add4);
            d.w = __vadd4(v.w, add4);
            *reinterpret_cast<uint4*>(qrow + c * 512) = d;
        }
        const bool any = ((or3(v.x - kb, v.y - kb, v.z - kb) | (v.w - kb)) & 0x80808080u) != 0;
        const uint32_t nz = __ballot_sync(0xffffffffu, any);
        if (nz) {  // warp uniform
#ifdef FQB_NO_DUMMY_STORE  // racecheck build (tools/gpu_sanitize.sh): the stores below are the only intended write-write overlap
            if (any) sts_u32(qa + 4u * __popc(nz & lt_mask), lane | uint32_t(c) << 5);
#else
            // lanes without a candidate store to the warp's dummy word (no divergent branch; the word is never read)
            sts_u32(any ? qa + 4u * __popc(nz & lt_mask) : q_dummy, lane | uint32_t(c) << 5);
#endif
            qa += 4u * __popc(nz);
        }
    }
    return int((qa - q0) >> 2);
}

// Block k of the warp's queue (entries 32k .. 32k+31), phase 2: my chunk's exact newline mask (from the staged
// tile at `warp_tile_s`), the entry (mask | chunk index << 16), the index of its first newline inside the warp's
// part of the tile, running warp total.  With one or two newlines in every queued chunk (the common case) the
// prefix is lane + (chunks with two below me), no shuffle scan.  Edge tiles (`special`): newlines outside the
// visible bytes are struck from the mask; rel_lo / rel_hi = first visible / first invisible byte relative to the
// warp's first byte of the tile.
__device__ __forceinline__ void queue_block(uint32_t q0, int k, int nq, int lane, uint32_t lt_mask, bool special,
                                            uint32_t warp_tile_s, long long rel_lo, long long rel_hi, uint32_t& qe, int& qpre,
                                            int& wtot)
{
    const int q = k * 32 + lane;
    const bool have = q < nq;
    uint32_t e = 0u;
    if (have) {
        const uint32_t ci = lds_u32(q0 + 4u * q);
        const uint4 v = lds_128(warp_tile_s + ci * 16u);
        uint32_t m = gather_flags16(newline_flags(v.x), newline_flags(v.y), newline_flags(v.z), newline_flags(v.w));
        if (special) {
            const long long b_lo = rel_lo - (long long)ci * 16, b_hi = rel_hi - (long long)ci * 16;
            uint32_t keep = 0xffffu;
            if (b_lo > 0) keep &= (b_lo >= 16) ? 0u : (0xffffu << int(b_lo));
            if (b_hi < 16) keep &= (b_hi <= 0) ? 0u : ((1u << int(b_hi)) - 1u);
            m &= keep;
        }
        e = m | (ci << 16);
    }
    const int cnt = __popc(e & 0xffffu);
    const uint32_t two = __ballot_sync(0xffffffffu, cnt >= 2);
    qe = e;
    // entries of this block that hold neither one nor two newlines (none: a false candidate or an edge tile)
    const bool odd = have && (unsigned(cnt - 1) >= 2u);
    if (!__any_sync(0xffffffffu, odd)) {
        const int nk = (nq - k * 32 < 32) ? nq - k * 32 : 32;  // entries of this block
        qpre = wtot + lane + __popc(two & lt_mask);
        wtot += nk + __popc(two);
    } else {
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int nb = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += nb;
        }
        qpre = wtot + inc - cnt;
        wtot += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// list entries of one queued chunk: (offset in list tile << 2) | class of the byte after the newline
__device__ __forceinline__ void emit_entries(uint32_t qe, unsigned short* slot, unsigned int idx, uint32_t warp_tile_s,
                                             int warp_off_lt, uint32_t cls_s)
{
    uint32_t m = qe & 0xffffu;
    const uint32_t coff = (qe >> 16) * 16u;
    const uint32_t nxt = warp_tile_s + coff + 1u;  // address of the byte after chunk byte 0
    const uint32_t e0 = (uint32_t(warp_off_lt) + coff) << 2;
    unsigned short* dst = slot + idx;
    if (m) {  // first and second newline of the chunk without loop-carried pointer arithmetic
        const uint32_t b = __ffs(m) - 1;
        m &= m - 1;
        dst[0] = (unsigned short)(e0 + (b << 2) + lds_u8(cls_s + lds_u8(nxt + b)));
        if (m) {
            const uint32_t b1 = __ffs(m) - 1;
            m &= m - 1;
            dst[1] = (unsigned short)(e0 + (b1 << 2) + lds_u8(cls_s + lds_u8(nxt + b1)));
            dst += 2;
            while (m) {  // three or more newlines in 16 bytes: rare
                const uint32_t b2 = __ffs(m) - 1;
                m &= m - 1;
                *dst++ = (unsigned short)(e0 + (b2 << 2) + lds_u8(cls_s + lds_u8(nxt + b2)));
            }
        }
    }
}

// warp 0 of the last CTA, after the count prefixes have been written: lines below own_end, published
template <int LT>
__device__ __forceinline__ void shard_tail_publish(const ScanParams& p, const ShardTail& t, int lane)
{
    ListView lv;
    lv.lists = p.lists;
    lv.lprefix = p.lprefix;
    lv.rprefix = p.rprefix;
    lv.n_tiles = int(p.n_tiles);
    lv.T = int(p.T);
    lv.slot_cap = p.slot_cap;
    lv.tile = LT;
    lv.virt = (p.sentinel && p.A > p.mis) ? 1 : 0;
    lv.mis = p.mis;
    lv.cls0 = 0;
    unsigned long long total = 0;
    int part = 0;
    if (t.own_end <= 0) {
        total = (lv.virt && t.own_end > (long long)lv.mis - 1) ? 1ull : 0ull;
    } else if (lv.n_tiles > 0) {
        long long te = t.own_end / lv.tile;
        if (te >= lv.n_tiles) te = lv.n_tiles - 1;
        const int tt = int(te);
        total = lv_base(lv, tt);
        const unsigned int n = lv_count(lv, tt);
        for (unsigned int jj = lane; jj < n; jj += 32) {
            long long a;
            unsigned int cls;
            lv_entry(lv, tt, jj, &a, &cls);
            if (a < t.own_end) ++part;
        }
    }
    part = __reduce_add_sync(0xffffffffu, part);
    const unsigned long long count = total + (unsigned long long)part;
    if (lane == 0 && t.own_lines) *t.own_lines = count;
    if (lane < t.n_pub) {
        unsigned long long* slot = t.pub[lane];
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(slot), "l"(count) : "memory");
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(slot + 1), "l"(t.epoch) : "memory");
    }
    if (lane == 0 && t.ready_left)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(t.ready_left), "l"(t.ready_epoch) : "memory");
}

template <int THREADS, int CPT, int STAGES, bool DEC, bool FASTA = false, bool SHARD = false>
// (the instance with the shard epilogue is held to the register budget of 6 CTAs per SM -- it shares its grid geometry
// with the plain instance; an explicit bound on the plain instance itself changes its code and costs it 3 %)
__global__ void __launch_bounds__(THREADS, (SHARD && THREADS == 256 && CPT == 4 && STAGES == 2) ? 6 : 0) fq_scan_kernel(const ScanParams p, const typename TailOf<SHARD>::type tail)
{
    using Cfg = ScanConfig<THREADS, CPT, STAGES>;
    constexpr int TILE = Cfg::TILE;
    constexpr int NW = Cfg::NW;
    constexpr int QCAP = Cfg::QCAP;
    static_assert(CPT >= 1 && CPT <= 8, "rows per warp and iteration");
    static_assert(NW <= 32, "one warp sums the warp totals");
    constexpr int LT = Cfg::LT, TPI = Cfg::TPI, WPL = Cfg::WPL;
    static_assert(TILE % LT == 0 && NW % TPI == 0 && LT % (32 * CPT * 16) == 0, "warps do not straddle list tiles");
    static_assert(TPI <= 2, "lprefix bookkeeping is unrolled for at most two list tiles per iteration");
    static_assert(size_t(THREADS) * 8 <= Cfg::SMEM, "range-prefix scan reuses the staging ring");

    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ int s_wtot[2][32];       // warp totals of the current / previous tile (entries >= NW stay 0)
    __shared__ uint8_t s_cls[256];      // class of a byte that follows a newline
    __shared__ bool s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // the kernel behind this one in the stream (fq_emit_kernel, launched with programmatic stream serialization) may be
    // scheduled as soon as SMs free up; it waits for this grid's completion before it reads anything
    asm volatile("griddepcontrol.launch_dependents;");
    const uint32_t lt_mask = (1u << lane) - 1u;
    const long long lo = p.mis;    // first visible byte
    // the last byte of the blob is never seen as a newline by the reference's C entrypos (memchr windows
    // exclude it; pairs need a 2nd byte)
    const long long hi = (FASTA && p.last_visible) ? p.A : p.A - 1;
    const long long t_begin = (long long)blockIdx.x * p.T;
    long long t_end = t_begin + p.T;
    if (t_end > p.n_tiles) t_end = p.n_tiles;
    const int ntl = t_end > t_begin ? int((t_end - t_begin + TPI - 1) / TPI) : 0;  // iterations (p.T % TPI == 0)
    const unsigned int slot_cap = (unsigned int)p.slot_cap;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    for (int b = tid; b < 256; b += THREADS) {
        if (FASTA)  // classes from the parameters ('>' is class 1)
            s_cls[b] = uint8_t(b == p.cls1 ? CLS_AT : (b == p.cls2 ? CLS_PLUS : (b == '\n' ? CLS_NL : CLS_OTHER)));
        else
            s_cls[b] = uint8_t(classify(uint8_t(b)));
    }
    if (tid < 64) s_wtot[tid >> 5][tid & 31] = 0;
    __syncthreads();

    auto issue_load = [&](int i) {  // called by thread 0
        if (i >= ntl) return;
        const int s = i % STAGES;
        const long long tile_base = t_begin * LT + (long long)i * TILE;
        long long avail = p.A - tile_base;
        if (avail > TILE + 16) avail = TILE + 16;
        if (avail <= 0) return;
        const uint32_t bytes = uint32_t(avail) & ~15u;
        if (bytes) {
            mbar_arrive_expect_tx(&full_bar[s], bytes);
            tma_load_1d(smem + size_t(s) * Cfg::STAGE_BYTES, p.base + tile_base, bytes, &full_bar[s]);
        }
#ifdef FQB_SCAN_L2PF  // experiment: the tile FQB_SCAN_L2PF iterations further on into L2 (HBM latency off the bulk copy's path)
        {
            const long long pf_base = tile_base + (long long)FQB_SCAN_L2PF * TILE;
            if (i + FQB_SCAN_L2PF < ntl && pf_base + TILE <= p.A)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.base + pf_base), "r"(uint32_t(TILE)) : "memory");
        }
#endif
    };

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) issue_load(i);
    }

    // tiles that need care: the first one (bytes before the buffer / virtual sentinel) and those that
    // touch the end of the buffer (partial bulk copy, invisible last byte); all others run unmasked
    const int i_first = (t_begin == 0 && lo > 0) ? 0 : -1;
    // first local iteration with tile_base + TILE + 16 > hi, tile_base = t_begin * LT + i * TILE
    const long long room = hi - TILE - 16 - t_begin * LT;
    long long i_end_ll = room < 0 ? 0 : room / TILE + 1;
    const int i_end = i_end_ll < 0 ? 0 : (i_end_ll > ntl ? ntl : int(i_end_ll));

    // shared-space addresses
    const uint32_t smem_s = smem_u32(smem), bar_s = smem_u32(full_bar), wtot_s = smem_u32(s_wtot), cls_s = smem_u32(s_cls);
    const uint32_t q0 = smem_s + uint32_t(STAGES) * Cfg::STAGE_BYTES + uint32_t(warp * QCAP) * 4u;  // my warp's queue
    // list tile of this warp inside the iteration and the warp's offset inside that list tile
    const int my_lt = warp / WPL;
    const uint32_t warp_off = uint32_t(warp) * (32 * CPT * 16);  // super-tile offset of the warp's first byte
    const int warp_off_lt = int(warp_off) - my_lt * LT;          // ... inside its list tile
    const uint32_t my_off = warp_off + uint32_t(lane) * 16;      // super-tile offset of my chunk of row 0
    // this CTA's list slots: element offsets from `slots` stay below 2^32 (checked by the host)
    unsigned short* const slots = p.lists + (size_t)t_begin * slot_cap;
    unsigned int slot_off = (unsigned int)my_lt * slot_cap;
    // lane 0 of the last warp keeps the books of the range (its wbase + wtot is the tile total)
    const bool book = (tid == THREADS - 32);
    unsigned int run = 0;  // newlines of this CTA's range so far (book keeper only)
    bool overflow = false;
    uint32_t stage_s = smem_s, bar = bar_s, parity = 0;
    int s = 0;
    for (int i = 0; i < ntl; ++i) {
        __syncwarp();  // the warp's queue entries of the previous tile have been consumed
        const bool special = (i == i_first) || (i >= i_end);
        if (!special) {
            while (!mbar_try_wait_s(bar, parity)) {
            }
        } else {
            const long long tile_base = t_begin * LT + (long long)i * TILE;
            const long long left = p.A - tile_base;  // > 0
            const long long availb = left < 0 ? 0 : (left < TILE + 16 ? left : TILE + 16);
            const int full16 = int(availb) & ~15, rem = int(availb) - full16;
            if (full16) {
                while (!mbar_try_wait_s(bar, parity)) {
                }
            }
            // the last <16 bytes of the buffer are fetched with plain loads (a bulk copy moves whole
            // 16-byte units and must not run past the caller's allocation)
            uint8_t* tile = smem + size_t(s) * Cfg::STAGE_BYTES;
            if (tid < rem) tile[full16 + tid] = p.base[tile_base + full16 + tid];
            __syncthreads();
            if (DEC && p.qual) {  // edge tiles: the mirror byte by byte, inside the buffer only
                const uint8_t add = uint8_t(p.add4 & 0xffu);
                for (int o = tid; o < TILE; o += THREADS) {
                    const long long a = tile_base + o;
                    if (a >= lo && a < p.A) p.qual[a] = int8_t(uint8_t(tile[o] + add));
                }
            }
        }
        int8_t* qrow = nullptr;
        if (DEC && !special && p.qual) qrow = p.qual + (t_begin * LT + (long long)i * TILE) + my_off;
#ifdef FQB_SCAN_LOADS_ONLY  // measurement build (tools/scan_skeleton.py): staging pipeline + per-tile skeleton, no row scan
        const int nq = (lds_u32(stage_s + my_off) == 0x12345678u && p.slot_cap == 1) ? 1 : 0;
#else
        const int nq = scan_rows<CPT, DEC>(stage_s + my_off, q0, lt_mask, uint32_t(lane), qrow, p.add4);
#endif
        // edge tiles: newlines outside the visible bytes [lo, hi) are struck from the masks of the queued chunks
        long long rel_lo = 0, rel_hi = 0;
        if (special) {
            const long long w0 = t_begin * LT + (long long)i * TILE + warp_off;  // the warp's first byte of the tile
            rel_lo = lo - w0;
            rel_hi = hi - w0;
        }
        __syncwarp();

        // ---- consumer, part 1: count the queued newlines, index of each chunk's first one ----
        const uint32_t warp_tile_s = stage_s + warp_off;
        uint32_t qe0 = 0;
        int qpre0 = 0, wtot = 0;
        if (nq > 0) queue_block(q0, 0, nq, lane, lt_mask, special, warp_tile_s, rel_lo, rel_hi, qe0, qpre0, wtot);
        uint32_t qe[CPT > 1 ? CPT - 1 : 1];  // blocks 1.. (rare: more than 32 chunks with newlines)
        int qpre[CPT > 1 ? CPT - 1 : 1];
        if (CPT > 1 && nq > 32) {
#pragma unroll
            for (int k = 1; k < CPT; ++k) {
                qe[k - 1] = 0;
                qpre[k - 1] = 0;
                if (k * 32 < nq)
                    queue_block(q0, k, nq, lane, lt_mask, special, warp_tile_s, rel_lo, rel_hi, qe[k - 1], qpre[k - 1], wtot);
            }
        }
        const uint32_t wt_par = wtot_s + uint32_t(i & 1) * 128u;
        if (lane == 0) sts_u32(wt_par + 4u * warp, uint32_t(wtot));
        __syncthreads();  // the only barrier per tile: warp totals visible, previous tile fully consumed
        if (STAGES > 1 && tid == 0 && i >= 1) issue_load(i - 1 + STAGES);  // refill the stage of the previous tile
        const int wv = int(lds_u32(wt_par + 4u * lane));
        const int wbase = __reduce_add_sync(0xffffffffu, (lane < warp && lane >= my_lt * WPL) ? wv : 0);

        // ---- consumer, part 2: list entries ----
        if ((unsigned int)(wbase + wtot) <= slot_cap) {
            unsigned short* slot = slots + slot_off;
            if (nq > 0) emit_entries(qe0, slot, (unsigned int)(wbase + qpre0), warp_tile_s, warp_off_lt, cls_s);
            if (CPT > 1 && nq > 32) {
#pragma unroll
                for (int k = 1; k < CPT; ++k)
                    if (k * 32 < nq) emit_entries(qe[k - 1], slot, (unsigned int)(wbase + qpre[k - 1]), warp_tile_s, warp_off_lt, cls_s);
            }
        } else {  // the slot is too small for this tile (reported as FQB_ERR_DENSE below): clip
            unsigned short* slot = slots + slot_off;
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                if (k * 32 < nq) {
                    const uint32_t e = (k == 0) ? qe0 : qe[k > 0 ? k - 1 : 0];
                    uint32_t m = e & 0xffffu;
                    const uint32_t coff = (e >> 16) * 16u;
                    unsigned int idx = (unsigned int)(wbase + ((k == 0) ? qpre0 : qpre[k > 0 ? k - 1 : 0]));
                    while (m) {
                        const uint32_t b = __ffs(m) - 1;
                        m &= m - 1;
                        if (idx < slot_cap)
                            slot[idx] = (unsigned short)(((uint32_t(warp_off_lt) + coff + b) << 2) | lds_u8(cls_s + lds_u8(warp_tile_s + coff + b + 1)));
                        ++idx;
                    }
                }
            }
        }
        // ---- range bookkeeping ----
        int n_t0 = 0;
        if (TPI > 1) n_t0 = __reduce_add_sync(0xffffffffu, (lane < WPL) ? wv : 0);  // first list tile of two
        if (book) {
            const int n_last = wbase + wtot;  // total of the last (or only) list tile of the iteration
            const long long lt0 = t_begin + (long long)i * TPI;
            if (TPI > 1) {
                p.lprefix[lt0] = run + (unsigned int)n_t0;
                p.lprefix[lt0 + 1] = run + (unsigned int)(n_t0 + n_last);
            } else {
                p.lprefix[lt0] = run + (unsigned int)n_last;
            }
            if ((unsigned int)n_t0 > slot_cap || (unsigned int)n_last > slot_cap) overflow = true;
            run += (unsigned int)(n_t0 + n_last);
        }
        if (i == 0 && t_begin == 0 && tid == 0) p.st->cls0 = FASTA ? (unsigned int)s_cls[smem[p.mis]] : classify(smem[p.mis]);  // stage 0 holds tile 0
        if (STAGES == 1) {  // single buffer: other CTAs of the SM cover the load latency
            __syncthreads();
            if (tid == 0) issue_load(i + 1);
        }
        slot_off += slot_cap * TPI;
        stage_s += Cfg::STAGE_BYTES;
        bar += 8;
        if (++s == STAGES) {
            s = 0;
            stage_s = smem_s;
            bar = bar_s;
            parity ^= 1u;
        }
    }
    if (overflow) p.st->error = FQB_ERR_DENSE;  // more newlines than the list slots hold

    // ---- range totals -> exclusive prefixes, by the last CTA to finish ----
    if (book) {
        p.rangetot[blockIdx.x] = run;
        __threadfence();
        s_last = (atomicAdd(&p.st->scan_done, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    unsigned long long* sc = reinterpret_cast<unsigned long long*>(smem);
    const int G = int(gridDim.x);
    const int per = (G + THREADS - 1) / THREADS;
    unsigned long long mine = 0;
    for (int q = 0; q < per; ++q) {
        const int b = tid * per + q;
        if (b < G) mine += *((volatile unsigned int*)&p.rangetot[b]);
    }
    sc[tid] = mine;
    __syncthreads();
    for (int o = 1; o < THREADS; o <<= 1) {
        const unsigned long long v = (tid >= o) ? sc[tid - o] : 0ull;
        __syncthreads();
        sc[tid] += v;
        __syncthreads();
    }
    unsigned long long acc = sc[tid] - mine;  // exclusive prefix of my first range
    for (int q = 0; q < per; ++q) {
        const int b = tid * per + q;
        if (b < G) {
            p.rprefix[b] = acc;
            acc += *((volatile unsigned int*)&p.rangetot[b]);
        }
    }
    if (tid == THREADS - 1) {
        const unsigned long long total = sc[THREADS - 1];
        p.rprefix[G] = total;
        const int virt = (p.sentinel && p.A > p.mis) ? 1 : 0;
        p.st->n_lines = total + (unsigned long long)virt;
    }
    if constexpr (SHARD) {
        __syncthreads();  // the prefixes written above are what the count below reads
        if (warp == 0) shard_tail_publish<LT>(p, tail, lane);
    }
}

}  // namespace fqb
