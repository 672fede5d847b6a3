// fq_synth.cuh -- synthetic FASTQ streams with VARIABLE record geometry (BASELINE.json configs[2..4]):
// Illumina-like 150 bp reads with variable-width headers, ONT-like long reads (10 kb mean), wrapped multi-line
// records with long '+' lines.  Not part of the reference; bench.py and the full-size parity tests need inputs of
// 8-64 GiB that exist only on the device, together with the TRUE offset table.
//
// Everything about record k is a pure function of (kind, seed, k); byte g of the stream additionally of (seed, g).
// So any window of the stream can be generated on any GPU (one shard per rank) once the record offsets are known:
//   1. fq_synth_meta_kernel   record k -> header / read / field / '+' line lengths and the record's total bytes;
//   2. (caller) exclusive prefix sum of the record bytes = stream offset of every record (= pos0 of the truth table);
//   3. fq_synth_fill_kernel   bytes [first_byte, first_byte + n_bytes) of the stream, 16 per thread.
// Record = '@' header \n sequence \n '+' [header] \n quality \n ; sequence and quality of wrapped records carry a
// '\n' after every `wrap` letters (identically, so that pos5 lands on the closing newline like data/test_multiline.fq).
// numpy twin: tests/fqgen.py:synth_records_np (bit identical, checked by the GPU tests).
#pragma once
#include "fq_common.cuh"
#include "fq_misc.cuh"

namespace fqb {

constexpr int SYNTH_ILLUMINA = 0, SYNTH_ONT = 1, SYNTH_MULTILINE = 2;
constexpr int SYNTH_QT_BITS = 12;  // read-length quantile table: 2^12 + 1 entries
constexpr int SYNTH_MAX_HEADER = 160;

struct SynthRec {
    int hl;    // header line: '@' + text, without the newline
    int rl;    // letters of the read
    int sb;    // bytes of the sequence (= quality) field: rl + embedded newlines
    int pl;    // '+' line without the newline: 1, or hl when the header is repeated
    int wrap;  // 0: single line
    int qmode; // 0 uniform '!'..'I' | 1 binned "#,:F" | 2 uniform '"'..'S'
    unsigned long long h[6];
};

__host__ __device__ __forceinline__ int dec_digits(unsigned long long v)
{
    int d = 1;
    while (v >= 10) {
        v /= 10;
        ++d;
    }
    return d;
}

__host__ __device__ __forceinline__ void synth_rec(int kind, unsigned long long seed, unsigned long long k, const int* qtable, SynthRec& r)
{
    // record-level hashes live in the half of the counter space the byte hashes (seed ^ g, g < 2^63) never reach
    r.h[0] = splitmix64(seed ^ (0x8000000000000000ull | k));
    for (int i = 1; i < 6; ++i) r.h[i] = splitmix64(r.h[i - 1]);
    r.wrap = 0;
    if (kind == SYNTH_ILLUMINA) {
        // @A00123:45:HXXXXXXXX:<lane 1-4>:<tile 1101-2678>:<x 1000-32000>:<y 1000-50000> 1:N:0:ACGTACGT
        const unsigned long long x = 1000 + ((r.h[0] >> 18) & 0xfffffull) % 31001ull;
        const unsigned long long y = 1000 + ((r.h[0] >> 38) & 0xfffffull) % 49001ull;
        r.hl = 44 + dec_digits(x) + dec_digits(y);
        r.rl = 150;
        r.pl = 1;
        r.qmode = ((r.h[1] & 0xffffull) % 10ull == 0) ? 0 : 1;  // 10 % of the records: uniform qualities ('@' / '+' hazards)
    } else if (kind == SYNTH_ONT) {
        // @<uuid> runid=<40 hex> read=<k> ch=<1-512> start_time=2026-01-01T00:00:00Z
        const unsigned long long ch = 1 + (r.h[5] & 0xffffull) % 512ull;
        r.hl = 126 + dec_digits(k) + dec_digits(ch);
        const unsigned int u = (unsigned int)(r.h[5] >> 40);  // 24 bits
        const int idx = int(u >> 12), frac = int(u & 4095u);
        const long long a = qtable[idx], b = qtable[idx + 1];
        r.rl = int(a + (((b - a) * frac) >> 12));
        r.pl = 1;
        r.qmode = 2;
    } else {
        // @SIM:%09d:<0-999999> len ; '+' repeats the header for half of the records; wrapped at 60 columns
        const unsigned long long v = (r.h[0] >> 20) % 1000000ull;
        r.hl = 19 + dec_digits(v);
        r.rl = 150 + int((r.h[1] >> 8) % 151ull);
        r.pl = (r.h[1] & 1ull) ? r.hl : 1;
        r.wrap = 60;
        r.qmode = 0;
    }
    r.sb = r.rl + (r.wrap ? (r.rl - 1) / r.wrap : 0);
}

__host__ __device__ __forceinline__ long long synth_rec_bytes(const SynthRec& r)
{
    return (long long)r.hl + 1 + r.sb + 1 + r.pl + 1 + r.sb + 1;
}

__host__ __device__ __forceinline__ int put_str(char* dst, const char* s)
{
    int n = 0;
    while (s[n]) {
        dst[n] = s[n];
        ++n;
    }
    return n;
}
__host__ __device__ __forceinline__ int put_dec(char* dst, unsigned long long v, int width /* 0: as many digits as needed */)
{
    const int n = width ? width : dec_digits(v);
    for (int i = n - 1; i >= 0; --i) {
        dst[i] = char('0' + v % 10);
        v /= 10;
    }
    return n;
}
__host__ __device__ __forceinline__ int put_hex(char* dst, unsigned long long v, int n)  // the low 4n bits, most significant first
{
    for (int i = n - 1; i >= 0; --i) {
        dst[i] = "0123456789abcdef"[v & 15ull];
        v >>= 4;
    }
    return n;
}

// header line of record k ('@' included, no newline) into hdr[0 .. r.hl)
__host__ __device__ inline void synth_header(int kind, unsigned long long k, const SynthRec& r, char* hdr)
{
    int n = 0;
    if (kind == SYNTH_ILLUMINA) {
        n += put_str(hdr + n, "@A00123:45:HXXXXXXXX:");
        n += put_dec(hdr + n, 1 + (r.h[0] & 3ull), 0);
        hdr[n++] = ':';
        n += put_dec(hdr + n, 1101 + ((r.h[0] >> 2) & 0xffffull) % 1578ull, 0);
        hdr[n++] = ':';
        n += put_dec(hdr + n, 1000 + ((r.h[0] >> 18) & 0xfffffull) % 31001ull, 0);
        hdr[n++] = ':';
        n += put_dec(hdr + n, 1000 + ((r.h[0] >> 38) & 0xfffffull) % 49001ull, 0);
        n += put_str(hdr + n, " 1:N:0:ACGTACGT");
    } else if (kind == SYNTH_ONT) {
        hdr[n++] = '@';
        n += put_hex(hdr + n, r.h[0] >> 32, 8);
        hdr[n++] = '-';
        n += put_hex(hdr + n, r.h[0] >> 16, 4);
        hdr[n++] = '-';
        n += put_hex(hdr + n, r.h[0], 4);
        hdr[n++] = '-';
        n += put_hex(hdr + n, r.h[1] >> 48, 4);
        hdr[n++] = '-';
        n += put_hex(hdr + n, r.h[1], 12);
        n += put_str(hdr + n, " runid=");
        n += put_hex(hdr + n, r.h[2], 16);
        n += put_hex(hdr + n, r.h[3], 16);
        n += put_hex(hdr + n, r.h[4], 8);
        n += put_str(hdr + n, " read=");
        n += put_dec(hdr + n, k, 0);
        n += put_str(hdr + n, " ch=");
        n += put_dec(hdr + n, 1 + (r.h[5] & 0xffffull) % 512ull, 0);
        n += put_str(hdr + n, " start_time=2026-01-01T00:00:00Z");
    } else {
        n += put_str(hdr + n, "@SIM:");
        n += put_dec(hdr + n, k % 1000000000ull, 9);
        hdr[n++] = ':';
        n += put_dec(hdr + n, (r.h[0] >> 20) % 1000000ull, 0);
        n += put_str(hdr + n, " len");
    }
}

// byte at offset o of record k (stream position g); hdr / have_hdr: lazily built header text of the record
__host__ __device__ inline uint8_t synth_byte(int kind, unsigned long long seed, unsigned long long k, long long g, long long o,
                                              const SynthRec& r, char* hdr, bool& have_hdr)
{
    const long long s0 = (long long)r.hl + 1;  // sequence field
    const long long p0 = s0 + r.sb + 1;        // '+' line
    const long long q0 = p0 + r.pl + 1;        // quality field
    if (o < r.hl || (o >= p0 + 1 && o < p0 + r.pl)) {  // header text (repeated behind the '+' of long '+' lines)
        if (!have_hdr) {
            synth_header(kind, k, r, hdr);
            have_hdr = true;
        }
        return uint8_t(hdr[o < r.hl ? o : o - p0]);
    }
    if (o == r.hl || o == s0 + r.sb || o == p0 + r.pl || o == q0 + r.sb) return '\n';
    if (o == p0) return '+';
    const bool is_q = o >= q0;
    const long long f = o - (is_q ? q0 : s0);  // position inside the field
    if (r.wrap && (f % (r.wrap + 1)) == r.wrap) return '\n';
    const unsigned int v = (unsigned int)(splitmix64(seed ^ (unsigned long long)g) >> 33);
    if (!is_q) return uint8_t("ACGT"[v & 3u]);
    if (r.qmode == 0) return uint8_t(33 + v % 41u);
    if (r.qmode == 1) return uint8_t("#,:F"[v & 3u]);
    return uint8_t(34 + v % 50u);
}

// ---- 1. per-record lengths ------------------------------------------------------------------------------------
// d_len[i] = bytes of record k0 + i; d_meta[i] = {hl, rl, sb, pl} (optional).
__global__ void __launch_bounds__(256) fq_synth_meta_kernel(int kind, unsigned long long seed, long long k0, long long n,
                                                            const int* qtable, int4* meta, long long* len)
{
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        SynthRec r;
        synth_rec(kind, seed, (unsigned long long)(k0 + i), qtable, r);
        if (meta) meta[i] = make_int4(r.hl, r.rl, r.sb, r.pl);
        if (len) len[i] = synth_rec_bytes(r);
    }
}

// ---- 3. the bytes ---------------------------------------------------------------------------------------------
// off[0 .. n]: stream offsets of records k0 .. k0 + n (off[n] = end of the last one); the window must lie inside
// [off[0], off[n]).  A thread writes 16 consecutive bytes; lane 0 of a warp finds the record of the warp's first
// byte by bisection, the lanes walk forward from there.
__global__ void __launch_bounds__(256) fq_synth_fill_kernel(int kind, unsigned long long seed, long long k0, long long n,
                                                            const long long* __restrict__ off, const int* qtable,
                                                            uint8_t* buf, long long first_byte, long long n_bytes)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long n_slices = (n_bytes + 511) / 512;
    char hdr[SYNTH_MAX_HEADER];
    for (long long s = warp; s < n_slices; s += nwarps) {
        const long long g_warp = first_byte + s * 512;
        long long j = 0;
        if (lane == 0) {  // largest j with off[j] <= g_warp
            long long lo = 0, hi = n;
            while (hi - lo > 1) {
                const long long mid = (lo + hi) >> 1;
                if (off[mid] <= g_warp)
                    lo = mid;
                else
                    hi = mid;
            }
            j = lo;
        }
        j = __shfl_sync(0xffffffffu, j, 0);
        const long long i0 = s * 512 + lane * 16;
        if (i0 >= n_bytes) continue;
        long long g = first_byte + i0;
        while (j + 1 < n && off[j + 1] <= g) ++j;
        SynthRec r;
        long long rec_lo = off[j], rec_hi = off[j + 1];
        synth_rec(kind, seed, (unsigned long long)(k0 + j), qtable, r);
        bool have_hdr = false;
        const int nb = (n_bytes - i0 < 16) ? int(n_bytes - i0) : 16;
        uint8_t out[16];
        for (int b = 0; b < nb; ++b, ++g) {
            if (g >= rec_hi) {
                ++j;
                rec_lo = rec_hi;
                rec_hi = off[j + 1];
                synth_rec(kind, seed, (unsigned long long)(k0 + j), qtable, r);
                have_hdr = false;
            }
            out[b] = synth_byte(kind, seed, (unsigned long long)(k0 + j), g, g - rec_lo, r, hdr, have_hdr);
        }
        uint8_t* dst = buf + i0;
        if (nb == 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            uint4 v;
            memcpy(&v, out, 16);
            *reinterpret_cast<uint4*>(dst) = v;
        } else {
            for (int b = 0; b < nb; ++b) dst[b] = out[b];
        }
    }
}

}  // namespace fqb
