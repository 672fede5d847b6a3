// fq_consume.cuh -- device-side consumers of the offset table: what the reference's users do per record
// inside an `entryfunc` (doc/user-guide.rst:153-204, src/demo/benchmark.py:47-83,159-168), done for the
// whole table at once so records never travel back to Python one by one:
//   * field lengths / length filter      lengthfilter_entryfunc: posarray[3] - posarray[2] < THRESHOLD
//   * index replay (gather)              benchmark_faf_c_index: buf[pos0:pos1], buf[pos2:pos3], buf[pos4:pos5]
//                                        of the rows read back from the on-disk index, packed contiguously
//   * Phred decode on gather / sums      biopython_entryfunc: frombytes(buf[pos4:pos5]); arrayadd_b(q, -33)
// plus the exclusive prefix sum that turns lengths into output offsets.
// Fields: 0 = header buf[pos0+1:pos1] (entryfunc, src/fastqandfurious.py:161-171), 1 = sequence
// buf[pos2:pos3], 2 = quality buf[pos4:pos5].
#pragma once
#include "fq_common.cuh"

namespace fqb {

constexpr int CONS_ERR_SEL = 1;   // a selected row index is outside the table
constexpr int CONS_ERR_SPAN = 2;  // a row's span is reversed or leaves the buffer

__device__ __forceinline__ void field_span(const long long* row, int field, long long& b, long long& e)
{
    if (field == 0) {
        b = row[0] + 1;
        e = row[1];
    } else if (field == 1) {
        b = row[2];
        e = row[3];
    } else {
        b = row[4];
        e = row[5];
    }
}

// out[i] = length of `field` of row sel[i] (sel == nullptr: row i).  With use_range the output is the
// 0/1 flag "min_len <= length <= max_len" (the length filter), ready for the prefix sum.
__global__ void __launch_bounds__(256) fq_field_lengths_kernel(const long long* table, long long n_rows, const long long* sel,
                                                               long long n_sel, int field, int use_range, long long min_len,
                                                               long long max_len, long long* out, int* status)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_sel; i += stride) {
        const long long r = sel ? sel[i] : i;
        long long len = 0;
        if (r < 0 || r >= n_rows) {
            if (status) atomicOr(status, CONS_ERR_SEL);
        } else {
            long long b, e;
            field_span(table + r * 6, field, b, e);
            len = e - b;
            if (len < 0) {
                len = 0;
                if (status) atomicOr(status, CONS_ERR_SPAN);
            }
        }
        out[i] = use_range ? ((len >= min_len && len <= max_len) ? 1 : 0) : len;
    }
}

// ---- exclusive prefix sum of int64 (three small kernels; in == out is allowed) ----
constexpr int PS_THREADS = 256, PS_ITEMS = 8, PS_BLOCK = PS_THREADS * PS_ITEMS;

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long* total)  // 256 threads
{
    __shared__ long long s_w[PS_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long nb = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += nb;
    }
    __syncthreads();  // s_w may still be read by the previous call
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    long long base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < PS_THREADS / 32; ++w) {
        const long long x = s_w[w];
        if (w < warp) base += x;
        tot += x;
    }
    *total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(PS_THREADS) fq_ps_sums_kernel(const long long* in, long long n, long long* bsum)
{
    const long long base = (long long)blockIdx.x * PS_BLOCK;
    long long s = 0;
#pragma unroll
    for (int k = 0; k < PS_ITEMS; ++k) {
        const long long i = base + k * PS_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    long long tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(PS_THREADS) fq_ps_top_kernel(long long* bsum, long long nb)  // one block
{
    long long carry = 0;
    for (long long c = 0; c < nb; c += PS_THREADS) {
        const long long i = c + threadIdx.x;
        const long long v = (i < nb) ? bsum[i] : 0;
        long long tot;
        const long long ex = block_exclusive_scan(v, &tot);
        if (i < nb) bsum[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) bsum[nb] = carry;
}

__global__ void __launch_bounds__(PS_THREADS) fq_ps_apply_kernel(const long long* in, long long n, const long long* bsum,
                                                                 long long nb, long long* out)
{
    // thread t owns the PS_ITEMS consecutive items base + t*PS_ITEMS ...
    const long long base = (long long)blockIdx.x * PS_BLOCK + (long long)threadIdx.x * PS_ITEMS;
    long long v[PS_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < PS_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    long long tot;
    long long run = bsum[blockIdx.x] + block_exclusive_scan(s, &tot);  // all reads of `in` are done: in == out is safe
#pragma unroll
    for (int k = 0; k < PS_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = bsum[nb];
}

// note: fq_ps_sums_kernel reads items strided by thread, fq_ps_apply_kernel consecutively; both orders
// cover the same block of PS_BLOCK items, so the block sums agree.

// idx_out[excl[i]] = i for every i with excl[i+1] != excl[i]  (order-preserving compaction of 0/1 flags)
__global__ void __launch_bounds__(256) fq_compact_kernel(const long long* excl, long long n, long long* idx_out)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long a = excl[i];
        if (excl[i + 1] != a) idx_out[a] = i;
    }
}

struct GatherParams {
    const uint8_t* buf;      // buffer the table indexes: table position p is byte buf[p - sub]
    long long len;
    long long sub;
    const long long* table;  // [n_rows][6]
    long long n_rows;
    const long long* sel;    // [n_sel] or nullptr (all rows in order)
    long long n_sel;
    int field;
    const long long* offsets;  // [n_sel + 1] exclusive prefix of the field lengths
    uint8_t* out;
    unsigned int add4;       // byte added to every copied byte, replicated (0: plain copy)
    int* status;
};

// Per-record metadata of a batch of 32 selected records, one record per lane: one round of independent
// loads per 32 records instead of a dependent chain per record.
struct RecMeta {
    long long b;    // first byte of the span inside buf, or -1 for a record that is skipped
    long long off;  // output offset
    int len;        // bytes
};

__device__ __forceinline__ RecMeta load_meta(const GatherParams& p, long long i, bool need_offsets)
{
    RecMeta m;
    m.b = -1;
    m.off = 0;
    m.len = 0;
    if (i >= p.n_sel) return m;
    const long long r = p.sel ? p.sel[i] : i;
    if (r < 0 || r >= p.n_rows) {
        if (p.status) atomicOr(p.status, CONS_ERR_SEL);
        return m;
    }
    long long b, e;
    field_span(p.table + r * 6, p.field, b, e);
    b -= p.sub;
    e -= p.sub;
    bool ok = !(b < 0 || e < b || e > p.len || e - b > 0x7fffffffll);
    if (ok && need_offsets) {
        m.off = p.offsets[i];
        ok = (p.offsets[i + 1] - m.off == e - b);
    }
    if (!ok) {
        if (p.status) atomicOr(p.status, CONS_ERR_SPAN);
        return m;
    }
    m.b = b;
    m.len = int(e - b);
    return m;
}

// out[offsets[i] : offsets[i+1]] = buf[b:e] (+ add) for every selected record.  A warp takes 32 records at
// a time (metadata loaded by one lane each, then handed to the copying lanes by shuffle); the body of a copy
// moves 4-byte words aligned to the DESTINATION, the source words being assembled from two aligned loads with a
// funnel shift, so every global access of the body is an aligned 4-byte access.
// GROUP lanes copy one record, 32 / GROUP records at a time: with the whole warp on one 150-byte record (38 words)
// the second of two rounds ran on 6 lanes and the per-record set-up (~60 instructions) was paid by all 32 -- 339 M
// warp instructions per GiB of 150 bp reads, issue bound at 0.53 ms.  Eight lanes per record (short fields) fill
// 95 % of the lanes and share the set-up between four records; long fields keep the whole warp.
template <int GROUP>
__device__ __forceinline__ void gather_records(const GatherParams& p, const RecMeta& mine, int nrec, int lane)
{
    constexpr int PER = 32 / GROUP;  // records copied at the same time
    const unsigned int add = p.add4 & 0xffu;
    const int g = lane / GROUP, gl = lane % GROUP;
    for (int j0 = 0; j0 < nrec; j0 += PER) {
        const int j = j0 + g;  // my group's record (lanes of a group agree)
        const long long b = __shfl_sync(0xffffffffu, mine.b, j & 31);
        const long long off = __shfl_sync(0xffffffffu, mine.off, j & 31);
        const int L = __shfl_sync(0xffffffffu, mine.len, j & 31);
        if (j >= nrec || b < 0) continue;
        const uint8_t* src = p.buf + b;
        uint8_t* dst = p.out + off;
        int head = int((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3);
        if (head > L) head = L;
        if (gl < head) dst[gl] = uint8_t(src[gl] + add);  // head < 4 <= GROUP
        const int nwords = (L - head) >> 2;
        const uint8_t* s0 = src + head;
        const unsigned int sh = (unsigned int)(reinterpret_cast<uintptr_t>(s0) & 3) * 8;
        const unsigned int* sa = reinterpret_cast<const unsigned int*>(s0 - (sh >> 3));
        unsigned int* da = reinterpret_cast<unsigned int*>(dst + head);
        for (int w = gl; w < nwords; w += GROUP) {
            const unsigned int lo = sa[w];
            const unsigned int hi = sh ? sa[w + 1] : 0u;  // aligned sources never look past their last word
            da[w] = __vadd4(__funnelshift_r(lo, hi, sh), p.add4);
        }
        const int done = head + (nwords << 2);
        if (done + gl < L) dst[done + gl] = uint8_t(src[done + gl] + add);  // < 4 bytes
    }
}

__global__ void __launch_bounds__(256) fq_gather_fields_kernel(const GatherParams p)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i0 = warp * 32; i0 < p.n_sel; i0 += nwarps * 32) {
        const RecMeta mine = load_meta(p, i0 + lane, true);
        const int nrec = (p.n_sel - i0 < 32) ? int(p.n_sel - i0) : 32;
        const int longest = __reduce_max_sync(0xffffffffu, mine.len);
        if (longest <= 1024)
            gather_records<8>(p, mine, nrec, lane);
        else
            gather_records<32>(p, mine, nrec, lane);
    }
}

// sum of (int8)(byte + add) over buf[b:e): aligned 4-byte loads + dp4a, single bytes at the ragged ends
__device__ __forceinline__ int span_sum(const uint8_t* buf, long long b, long long e, unsigned int add4)
{
    const unsigned int add = add4 & 0xffu;
    int s = 0;
    long long a = b;
    while (a < e && (reinterpret_cast<uintptr_t>(buf + a) & 3)) s += int(int8_t(uint8_t(buf[a++] + add)));
    for (; a + 4 <= e; a += 4)
        s = __dp4a(int(__vadd4(*reinterpret_cast<const unsigned int*>(buf + a), add4)), 0x01010101, s);
    while (a < e) s += int(int8_t(uint8_t(buf[a++] + add)));
    return s;
}

// sums[i] = sum over the bytes of `field` of row sel[i] of (int8)(byte + add)   (add = -33: the sum of the
// Phred scores; the caller divides by the length for the mean quality).  32 records per warp and step: short
// spans (<= 1 KiB) are summed by one lane each, long ones by the whole warp.
__global__ void __launch_bounds__(256) fq_field_sums_kernel(const GatherParams p, long long* sums)
{
    constexpr int LANE_MAX = 1024;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i0 = warp * 32; i0 < p.n_sel; i0 += nwarps * 32) {
        const RecMeta mine = load_meta(p, i0 + lane, false);
        long long my_total = 0;
        const bool is_long = mine.b >= 0 && mine.len > LANE_MAX;
        if (mine.b >= 0 && !is_long) my_total = span_sum(p.buf, mine.b, mine.b + mine.len, p.add4);
        unsigned int todo = __ballot_sync(0xffffffffu, is_long);
        while (todo) {  // warp uniform
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const long long b = __shfl_sync(0xffffffffu, mine.b, j);
            const int L = __shfl_sync(0xffffffffu, mine.len, j);
            // lane l sums the slice [b + l*per, b + (l+1)*per), per a multiple of 4
            const int per = ((L + 31) / 32 + 3) & ~3;
            long long lo = b + (long long)lane * per, hi = lo + per;
            if (hi > b + L) hi = b + L;
            const int s = (lo < hi) ? span_sum(p.buf, lo, hi, p.add4) : 0;
            const int tot = __reduce_add_sync(0xffffffffu, s);
            if (lane == j) my_total = tot;
        }
        if (i0 + lane < p.n_sel) sums[i0 + lane] = my_total;
    }
}

// ---- 2-bit packed sequences ("packed (2-bit) sequence extraction") -----------------------------
// Bases of the sequence field, embedded newlines of wrapped records skipped, four per byte, base i of a record in
// bits 2(i % 4) .. 2(i % 4) + 1 of byte i / 4 of the record's slot: A/a = 0, C/c = 1, G/g = 2, T/t/U/u = 3; any
// other byte is encoded by the same bit formula (((b >> 1) & 3) ^ ((b >> 2) & 1), 'N' -> 2) and counted in
// n_other[i], so that records with ambiguity codes can be told apart and fetched as bytes (fq_gather_fields_kernel).
// Slot of record i: out[offsets[i] : offsets[i+1]), 4 * ceil(L / 16) bytes for a field of L bytes (whole 32-bit
// words: every store is an aligned word); the words behind the last base are zero.
// codes of the 4 bytes of w in the low 2 bits of every byte
__device__ __forceinline__ unsigned int base_codes4(unsigned int w)
{
    return ((w >> 1) & 0x03030303u) ^ ((w >> 2) & 0x01010101u);
}
// byte-wise codes (2 bits at the bottom of each byte) -> 8 bits; partial products land on distinct bits
__device__ __forceinline__ unsigned int squeeze_codes4(unsigned int c)
{
    return (c * 0x01041040u) >> 24;
}
// 0x80 in every byte of x that equals the byte replicated in k (exact)
__device__ __forceinline__ unsigned int eq_flags4(unsigned int x, unsigned int k)
{
    const unsigned int k7 = 0x7f7f7f7fu;
    const unsigned int y = x ^ k;
    return ~(((y & k7) + k7) | y) & 0x80808080u;
}
__device__ __forceinline__ unsigned int acgtu_flags4(unsigned int w)
{
    const unsigned int x = w & 0xdfdfdfdfu;  // upper case
    return eq_flags4(x, 0x41414141u) | eq_flags4(x, 0x43434343u) | eq_flags4(x, 0x47474747u) | eq_flags4(x, 0x54545454u) |
           eq_flags4(x, 0x55555555u);
}

struct Pack2State {
    unsigned long long acc;  // packed bits not yet stored
    int nbits;
    unsigned int* dst;       // next output word
    long long bases, other;
};

// 0x80 flags of a word -> bit j for byte j (partial products land on distinct bits)
__device__ __forceinline__ unsigned int flags_to_mask4(unsigned int f)
{
    return (((f >> 7) * 0x00204081u) >> 21) & 0xfu;
}

// 16 bytes of the field starting at byte a (any alignment), of which the first `take` (1..16) belong to the field
__device__ __forceinline__ void pack2_step(const uint8_t* buf, long long a, int take, Pack2State& st)
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(buf + a);
    const unsigned int* w32 = reinterpret_cast<const unsigned int*>(addr & ~uintptr_t(3));
    const int mis = int(addr & 3);
    // aligned words that cover the wanted bytes: none past the word that holds the last of them
    const int nw = (mis + take + 3) >> 2;  // 1..5
    unsigned int r[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) r[q] = (q < nw) ? w32[q] : 0u;
    unsigned int nl = 0, ok = 0, codes = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        unsigned int x = __funnelshift_r(r[q], r[q + 1], mis * 8);
        const int have = take - 4 * q;  // bytes of this word that belong to the field; the others read as 'A'
        if (have <= 0)
            x = 0x41414141u;
        else if (have < 4)
            x = (x & (0xffffffffu >> (32 - 8 * have))) | (0x41414141u << (8 * have));
        nl |= flags_to_mask4(eq_flags4(x, 0x0a0a0a0au)) << (4 * q);
        ok |= flags_to_mask4(acgtu_flags4(x)) << (4 * q);
        codes |= squeeze_codes4(base_codes4(x)) << (8 * q);
    }
    const unsigned int in_field = (take >= 16) ? 0xffffu : ((1u << take) - 1u);
    const unsigned int valid = in_field & ~nl;
    const int cnt = __popc(valid);
    st.other += cnt - __popc(ok & valid);
    unsigned int cw = codes;
    if (valid != in_field) {  // embedded newline(s): squeeze their codes out
        cw = 0;
        int k = 0;
        for (unsigned int m = valid; m; m &= m - 1) {
            const int b = __ffs(m) - 1;
            cw |= ((codes >> (2 * b)) & 3u) << (2 * k);
            ++k;
        }
    } else if (take < 16) {
        cw &= (1u << (2 * take)) - 1u;
    }
    st.bases += cnt;
    st.acc |= (unsigned long long)cw << st.nbits;
    st.nbits += 2 * cnt;
    if (st.nbits >= 32) {
        *st.dst++ = (unsigned int)st.acc;
        st.acc >>= 32;
        st.nbits -= 32;
    }
}

__device__ __forceinline__ void pack2_span(const uint8_t* buf, long long b, long long e, Pack2State& st)
{
    for (long long a = b; a < e; a += 16) pack2_step(buf, a, (e - a < 16) ? int(e - a) : 16, st);
}

__global__ void __launch_bounds__(256) fq_pack2_kernel(const GatherParams p, long long* n_bases, long long* n_other)
{
    constexpr int LANE_MAX = 2048;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i0 = warp * 32; i0 < p.n_sel; i0 += nwarps * 32) {
        RecMeta mine = load_meta(p, i0 + lane, false);
        if (mine.b >= 0) {  // the slot must be the whole words of the field's byte length, at a 4-byte aligned offset
            mine.off = p.offsets[i0 + lane];
            const long long slot = p.offsets[i0 + lane + 1] - mine.off;
            if (slot != 4ll * ((mine.len + 15) / 16) || (mine.off & 3)) {
                if (p.status) atomicOr(p.status, CONS_ERR_SPAN);
                mine.b = -1;
            }
        }
        long long my_bases = 0, my_other = 0;
        const bool is_long = mine.b >= 0 && mine.len > LANE_MAX;
        if (mine.b >= 0 && !is_long) {  // short fields: one lane each
            Pack2State st = {0ull, 0, reinterpret_cast<unsigned int*>(p.out + mine.off), 0, 0};
            unsigned int* const end = st.dst + (mine.len + 15) / 16;
            pack2_span(p.buf, mine.b, mine.b + mine.len, st);
            if (st.nbits > 0 && st.dst < end) *st.dst++ = (unsigned int)st.acc;
            while (st.dst < end) *st.dst++ = 0u;  // words behind the last base (embedded newlines shorten the record)
            my_bases = st.bases;
            my_other = st.other;
        }
        unsigned int todo = __ballot_sync(0xffffffffu, is_long);
        while (todo) {  // long fields: the warp shares one, 16-byte aligned slices -> word aligned output
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const long long b = __shfl_sync(0xffffffffu, mine.b, j);
            const int L = __shfl_sync(0xffffffffu, mine.len, j);
            const long long off = __shfl_sync(0xffffffffu, mine.off, j);
            const int per = ((L + 31) / 32 + 15) & ~15;
            long long lo = (long long)lane * per, hi = lo + per;
            if (hi > L) hi = L;
            Pack2State st = {0ull, 0, reinterpret_cast<unsigned int*>(p.out + off) + lo / 16, 0, 0};
            if (lo < hi) {
                pack2_span(p.buf, b + lo, b + hi, st);
                if (st.nbits > 0) *st.dst++ = (unsigned int)st.acc;
            }
            // a slice that met a newline shifts every base behind it: such a field is redone by one lane
            const bool shifted = lo < hi && st.bases != hi - lo;
            long long tb, to;
            if (__any_sync(0xffffffffu, shifted)) {
                __syncwarp();
                tb = to = 0;
                if (lane == j) {
                    Pack2State s1 = {0ull, 0, reinterpret_cast<unsigned int*>(p.out + off), 0, 0};
                    unsigned int* const end = s1.dst + (L + 15) / 16;
                    pack2_span(p.buf, b, b + L, s1);
                    if (s1.nbits > 0 && s1.dst < end) *s1.dst++ = (unsigned int)s1.acc;
                    while (s1.dst < end) *s1.dst++ = 0u;
                    tb = s1.bases;
                    to = s1.other;
                }
            } else {
                tb = st.bases;
                to = st.other;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    tb += __shfl_xor_sync(0xffffffffu, tb, o);
                    to += __shfl_xor_sync(0xffffffffu, to, o);
                }
            }
            if (lane == j) {
                my_bases = tb;
                my_other = to;
            }
        }
        if (i0 + lane < p.n_sel) {
            if (n_bases) n_bases[i0 + lane] = my_bases;
            if (n_other) n_other[i0 + lane] = my_other;
        }
    }
}

}  // namespace fqb
