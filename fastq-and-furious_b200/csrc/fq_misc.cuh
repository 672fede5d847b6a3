// fq_misc.cuh -- array helpers of the reference's C extension and the synthetic-input generators.
#pragma once
#include "fq_common.cuh"
#include "fq_emit.cuh"

namespace fqb {

// arrayadd_b (src/_fastqandfurious.c:161-185): int8 a[i] += (int8)value, two's-complement wrap.
// 16 bytes per thread per step where the pointer allows it, scalar head / tail otherwise.
__global__ void __launch_bounds__(256) fq_arrayadd_b_kernel(int8_t* a, long long n, unsigned int add4)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(a);
    long long head = (long long)((16 - (addr & 15)) & 15);
    if (head > n) head = n;
    const long long nvec = (n - head) >> 4;
    uint4* v = reinterpret_cast<uint4*>(a + head);
    for (long long i = tid; i < nvec; i += nthreads) {
        uint4 x = v[i];
        x.x = __vadd4(x.x, add4);
        x.y = __vadd4(x.y, add4);
        x.z = __vadd4(x.z, add4);
        x.w = __vadd4(x.w, add4);
        v[i] = x;
    }
    const uint8_t add = uint8_t(add4 & 0xffu);
    const long long tail0 = head + (nvec << 4);
    const long long nscalar = head + (n - tail0);
    for (long long i = tid; i < nscalar; i += nthreads) {
        const long long j = (i < head) ? i : (tail0 + (i - head));
        a[j] = int8_t(uint8_t(uint8_t(a[j]) + add));
    }
}

// arrayadd_q (src/_fastqandfurious.c:193-217): int64 a[i] += value (wrapping).
__global__ void __launch_bounds__(256) fq_arrayadd_q_kernel(long long* a, long long n, long long value)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (long long i = tid; i < n; i += nthreads)
        a[i] = (long long)((unsigned long long)a[i] + (unsigned long long)value);
}

__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// lines (visible newlines + virtual sentinel) of a shard at buffer offsets < own_len: what the shard
// contributes to the global line rank of the shards after it.  One warp: everything before the tile that
// holds own_end comes from the count prefixes, the entries of that tile are compared by one lane each.
struct PubList {
    unsigned long long* p[16];
};
__global__ void fq_own_lines_kernel(ListView lv, const ParseState* st, long long own_end, unsigned long long* out,
                                    PubList pub, int n_pub, unsigned long long epoch, unsigned long long* ready_left,
                                    unsigned long long ready_epoch)
{
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    // the emit kernel behind this one (programmatic stream serialization) may be dispatched already; it waits for this
    // grid's completion before it reads anything
    asm volatile("griddepcontrol.launch_dependents;");
    // "my bytes of the next parse are in place" to the left neighbour (it was a kernel of its own: one launch less)
    if (lane == 0 && ready_left)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(ready_left), "l"(ready_epoch) : "memory");
    lv.cls0 = st->cls0;
    unsigned long long total = 0;
    int part = 0;
    if (own_end <= 0) {
        total = (lv.virt && own_end > (long long)lv.mis - 1) ? 1ull : 0ull;
    } else if (lv.n_tiles > 0) {
        long long te = own_end / lv.tile;
        if (te >= lv.n_tiles) te = lv.n_tiles - 1;
        const int t = int(te);
        total = lv_base(lv, t);
        unsigned int n = lv_count(lv, t);  // augmented
        if (t == 0 && lv.virt) {           // the virtual sentinel (byte index mis - 1) is not in the stored list
            if (lane == 0 && (long long)lv.mis - 1 < own_end) ++part;
            n -= 1;
        }
        // the tile's stored entries below own_end: eight per lane and load (the slot is 16-byte aligned and padded), all
        // loads of a round independent -- this kernel sits between the scan and the emit of every sharded step
        const unsigned short* src = lv.lists + (size_t)t * (unsigned int)lv.slot_cap;
        const long long rel_end = own_end - (long long)t * lv.tile;  // > 0
        for (unsigned int v = lane * 8; v < n; v += 256) {
            const uint4 x = *reinterpret_cast<const uint4*>(src + v);
            const unsigned int ee[8] = {x.x & 0xffffu, x.x >> 16, x.y & 0xffffu, x.y >> 16,
                                        x.z & 0xffffu, x.z >> 16, x.w & 0xffffu, x.w >> 16};
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (v + k < n && (long long)(ee[k] >> 2) < rel_end) ++part;
        }
    }
    part = __reduce_add_sync(0xffffffffu, part);
    const unsigned long long count = total + (unsigned long long)part;
    if (lane == 0) *out = count;
    // fused exchange: store {count, epoch} into the slot this shard owns in every later shard's memory
    // (peer-mapped pointers, NVLink stores); the epoch is released after the count
    if (lane < n_pub) {
        unsigned long long* slot = pub.p[lane];
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(slot), "l"(count) : "memory");
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(slot + 1), "l"(epoch) : "memory");
    }
}

// *out = sum of the values behind up to 16 device pointers -- peer-mapped memory of other GPUs included
// (NVLink loads): the line base of a shard from the counts its left neighbours published
struct PtrList {
    const unsigned long long* p[16];
};
__global__ void fq_sum_ptrs_kernel(PtrList pl, int n, unsigned long long* out)
{
    unsigned long long v = 0;
    if ((int)threadIdx.x < n) v = *((volatile const unsigned long long*)pl.p[threadIdx.x]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) *out = v;
}

// Halo pull over peer memory (fused exchange, step 1): ONE kernel replaces the device-side barrier and the
// peer copy.  Thread 0 of CTA 0 first tells the LEFT neighbour "my bytes of this epoch are in place" (a
// release store into its ready slot through a peer-mapped pointer); every CTA then waits until the RIGHT
// neighbour has said the same to us (acquire loads of our local slot, 10 s timeout -> *status = 1) and
// copies its slice of the neighbour's first `n` bytes over NVLink into the halo behind our own bytes.
// "my bytes of `epoch` are in place": release store into the left neighbour's ready slot (peer-mapped pointer)
__global__ void fq_signal_ready_kernel(unsigned long long* ready_left, unsigned long long epoch)
{
    if (blockIdx.x == 0 && threadIdx.x == 0)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(ready_left), "l"(epoch) : "memory");
}

// The waiting half of the halo pull on its own, ONE warp: spins until the right neighbour has announced `epoch` (10 s
// timeout -> *status = 1).  Small enough to sit next to a full grid of scan CTAs on a second stream, so that the bytes
// themselves can then travel by a copy engine (peer copy) while the SMs scan the previous buffer.
__global__ void __launch_bounds__(32) fq_wait_ready_kernel(const unsigned long long* ready_local, unsigned long long epoch,
                                                           int* status)
{
    if (threadIdx.x != 0) return;
    const unsigned long long t0 = global_timer_ns();
    unsigned int spins = 0;
    while (ld_acquire_sys(ready_local) < epoch) {
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 10000000000ull) {
            if (status) *status = 1;
            return;
        }
    }
}

__global__ void __launch_bounds__(256) fq_halo_pull_kernel(uint8_t* dst, const uint8_t* src, long long n,
                                                           const unsigned long long* ready_local,
                                                           unsigned long long* ready_left, unsigned long long epoch,
                                                           int* status)
{
    if (ready_left && blockIdx.x == 0 && threadIdx.x == 0)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(ready_left), "l"(epoch) : "memory");
    if (n <= 0 || !src) return;
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        int good = 1;
        const unsigned long long t0 = global_timer_ns();
        unsigned int spins = 0;
        while (ld_acquire_sys(ready_local) < epoch) {
            if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 10000000000ull) {
                good = 0;
                break;
            }
        }
        s_ok = good;
    }
    __syncthreads();
    if (!s_ok) {
        if (threadIdx.x == 0 && status) *status = 1;
        return;
    }
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
        const long long nvec = n >> 4;
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (long long i = tid; i < nvec; i += nthreads) d4[i] = s4[i];
        for (long long i = (nvec << 4) + tid; i < n; i += nthreads) dst[i] = src[i];
    } else {
        for (long long i = tid; i < n; i += nthreads) dst[i] = src[i];
    }
}

// Fixed-geometry synthetic FASTQ (SURVEY.md 8d cfg 2): byte g depends only on (seed, g).
// Record = '@SIM:' zero-padded decimal index ' 1:N:0:ACGTACGT' \n bases \n + \n quals \n ;
// bases uniform ACGT, qualities uniform '!'..'I' (so '+' and '@' occur).  numpy twin:
// tests/fqgen.py:fixed_records_np.
__global__ void __launch_bounds__(256) fq_synth_fixed_kernel(uint8_t* buf, long long n_bytes, long long first_byte,
                                                             int header_len, int read_len, unsigned long long seed)
{
    const long long rec = (long long)header_len + 1 + read_len + 1 + 2 + read_len + 1;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const int width = header_len - 20;
    const char* prefix = "@SIM:";
    const char* suffix = " 1:N:0:ACGTACGT";
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_bytes; i += nthreads) {
        const long long g = first_byte + i;  // byte index in the (unbounded) synthetic stream
        const long long k = g / rec;
        const int o = int(g - k * rec);
        const unsigned long long h = splitmix64(seed ^ (unsigned long long)g);
        uint8_t c;
        const int s0 = header_len + 1, q0 = s0 + read_len + 3;
        if (o < 5) {
            c = uint8_t(prefix[o]);
        } else if (o < 5 + width) {
            unsigned long long v = (unsigned long long)k;
            for (int d = width - 1 - (o - 5); d > 0; --d) v /= 10;
            c = uint8_t('0' + v % 10);
        } else if (o < header_len) {
            c = uint8_t(suffix[o - 5 - width]);
        } else if (o == header_len) {
            c = '\n';
        } else if (o < s0 + read_len) {
            c = uint8_t("ACGT"[(h >> 33) & 3]);
        } else if (o == s0 + read_len) {
            c = '\n';
        } else if (o == s0 + read_len + 1) {
            c = '+';
        } else if (o == s0 + read_len + 2) {
            c = '\n';
        } else if (o < q0 + read_len) {
            c = uint8_t(33 + (h >> 33) % 41);
        } else {
            c = '\n';
        }
        buf[i] = c;
    }
}

}  // namespace fqb
