// fqb200.cu -- C ABI of libfqb200.so (see include/fqb200.h): argument checks and kernel launches.
//
// Everything is enqueued on the caller's stream; the library allocates nothing and keeps no state
// between calls except a per-device cache of launch geometry (SM count, resident CTAs per SM).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "../../include/fqb200.h"
#include "fq_common.cuh"
#include "fq_consume.cuh"
#include "fq_emit.cuh"
#include "fq_fasta.cuh"
#include "fq_general.cuh"
#include "fq_gspec.cuh"
#include "fq_gspec2.cuh"
#include "fq_misc.cuh"
#include "fq_scan.cuh"
#include "fq_synth.cuh"

using namespace fqb;

namespace {

// ---- scan kernel configurations (flags bits 8..11 select one; 0 = default) ----------------------
struct ScanCfgInfo {
    int threads, cpt, stages;
};
constexpr int N_CFG = 8;
constexpr ScanCfgInfo kCfg[N_CFG] = {{256, 4, 2}, {256, 4, 3}, {256, 8, 2}, {256, 8, 1}, {256, 4, 1}, {128, 4, 4},
                                        {128, 8, 2}, {128, 8, 3}};
constexpr int MAX_GRID = 148 * 16;  // upper bound of scan CTAs (sizes rangetot / rprefix)

constexpr int MAX_DEV = 32;
struct DevCache {
    bool ready;
    int sms;
    int occ[N_CFG];
    int occ_emit[2];  // resident CTAs per SM of fq_emit_kernel / fq_decode_kernel
    int occ_spec;     // ... of fq_gspec_kernel
    int occ_spec2;    // ... of fq_gspec2_kernel
};
DevCache g_dev[MAX_DEV];
std::mutex g_dev_mutex;  // the cache is filled once per device; callers may come from several host threads

template <int T, int C, int S>
cudaError_t prep_kernel(int* occ)
{
    auto kern = fq_scan_kernel<T, C, S, false>;
    auto kern_dec = fq_scan_kernel<T, C, S, true>;
    const size_t smem = ScanConfig<T, C, S>::SMEM;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(kern_dec, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))) != cudaSuccess) return e;
    if (T == 256 && C == 4 && S == 2) {  // the default geometry also exists as FASTA scan and with the shard epilogue
        const int sm = int(ScanConfig<256, 4, 2>::SMEM);
        e = cudaFuncSetAttribute(fq_scan_kernel<256, 4, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(fq_scan_kernel<256, 4, 2, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(fq_scan_kernel<256, 4, 2, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        if (e != cudaSuccess) return e;
    }
    if (T == 256 && C == 8 && S == 2) {  // ... and as FASTA scan
        e = cudaFuncSetAttribute(fq_scan_kernel<256, 8, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
    }
    int occ_dec = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dec, kern_dec, T, smem)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, T, smem)) != cudaSuccess) return e;
    if (occ_dec < *occ) *occ = occ_dec;  // one grid geometry for both variants
    return cudaSuccess;
}

cudaError_t device_cache(DevCache** out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= MAX_DEV) return cudaErrorInvalidDevice;
    DevCache& d = g_dev[dev];
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    if (!d.ready) {
        if ((e = cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        if ((e = prep_kernel<256, 4, 2>(&d.occ[0])) != cudaSuccess) return e;
        if ((e = prep_kernel<256, 4, 3>(&d.occ[1])) != cudaSuccess) return e;
        if ((e = prep_kernel<256, 8, 2>(&d.occ[2])) != cudaSuccess) return e;
        if ((e = prep_kernel<256, 8, 1>(&d.occ[3])) != cudaSuccess) return e;
        if ((e = prep_kernel<256, 4, 1>(&d.occ[4])) != cudaSuccess) return e;
        if ((e = prep_kernel<128, 4, 4>(&d.occ[5])) != cudaSuccess) return e;
        if ((e = prep_kernel<128, 8, 2>(&d.occ[6])) != cudaSuccess) return e;
        if ((e = prep_kernel<128, 8, 3>(&d.occ[7])) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_emit[0], fq_emit_kernel, 256, 0)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_emit[1], fq_decode_kernel, DEC_THREADS, 0)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(fq_gspec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(GS_SMEM))) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_spec, fq_gspec_kernel, GS_THREADS, GS_SMEM)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(fq_gspec2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(G2_SMEM))) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_spec2, fq_gspec2_kernel, G2_THREADS, G2_SMEM)) != cudaSuccess) return e;
        d.ready = true;
    }
    *out = &d;
    return cudaSuccess;
}

template <int T, int C, int S>
cudaError_t launch_scan_t(const ScanParams& p, int grid, cudaStream_t stream)
{
    if (p.qual)
        fq_scan_kernel<T, C, S, true><<<grid, T, ScanConfig<T, C, S>::SMEM, stream>>>(p, NoTail());
    else
        fq_scan_kernel<T, C, S, false><<<grid, T, ScanConfig<T, C, S>::SMEM, stream>>>(p, NoTail());
    return cudaGetLastError();
}

// the default configuration with the shard epilogue (ShardTail): count + publication + ready signal by the last CTA
cudaError_t launch_scan_shard(const ScanParams& p, const ShardTail& tail, int grid, cudaStream_t stream)
{
    if (p.qual)
        fq_scan_kernel<256, 4, 2, true, false, true><<<grid, 256, ScanConfig<256, 4, 2>::SMEM, stream>>>(p, tail);
    else
        fq_scan_kernel<256, 4, 2, false, false, true><<<grid, 256, ScanConfig<256, 4, 2>::SMEM, stream>>>(p, tail);
    return cudaGetLastError();
}

cudaError_t launch_scan(int cfg, const ScanParams& p, int grid, cudaStream_t stream)
{
    if (p.last_visible) {  // FASTA: the default geometry, or (cfg 2) 32 KiB per iteration
        if (cfg == 2)
            fq_scan_kernel<256, 8, 2, false, true><<<grid, 256, ScanConfig<256, 8, 2>::SMEM, stream>>>(p, NoTail());
        else
            fq_scan_kernel<256, 4, 2, false, true><<<grid, 256, ScanConfig<256, 4, 2>::SMEM, stream>>>(p, NoTail());
        return cudaGetLastError();
    }
    switch (cfg) {
        case 0: return launch_scan_t<256, 4, 2>(p, grid, stream);
        case 1: return launch_scan_t<256, 4, 3>(p, grid, stream);
        case 2: return launch_scan_t<256, 8, 2>(p, grid, stream);
        case 3: return launch_scan_t<256, 8, 1>(p, grid, stream);
        case 4: return launch_scan_t<256, 4, 1>(p, grid, stream);
        case 5: return launch_scan_t<128, 4, 4>(p, grid, stream);
        case 6: return launch_scan_t<128, 8, 2>(p, grid, stream);
        default: return launch_scan_t<128, 8, 3>(p, grid, stream);
    }
}

// ---- optional timing of the dominant (scan) kernel, for bench.py's roofline record ----
constexpr int PROF_RING = 64;
struct Profile {
    bool on;
    int dev;
    cudaEvent_t start[PROF_RING], stop[PROF_RING];
    bool made[PROF_RING];
    int pending;       // event pairs recorded and not yet read
    double total_ms;   // accumulated by fqb_profile_read / ring wrap
    long long count;
};
Profile g_prof;

cudaError_t prof_flush()
{
    for (int i = 0; i < g_prof.pending; ++i) {
        cudaError_t e = cudaEventSynchronize(g_prof.stop[i]);
        if (e != cudaSuccess) return e;
        float ms = 0.f;
        if ((e = cudaEventElapsedTime(&ms, g_prof.start[i], g_prof.stop[i])) != cudaSuccess) return e;
        g_prof.total_ms += ms;
        g_prof.count += 1;
    }
    g_prof.pending = 0;
    return cudaSuccess;
}

cudaError_t prof_slot(int* slot)
{
    if (g_prof.pending == PROF_RING) {
        cudaError_t e = prof_flush();
        if (e != cudaSuccess) return e;
    }
    const int i = g_prof.pending;
    if (!g_prof.made[i]) {
        cudaError_t e;
        if ((e = cudaEventCreate(&g_prof.start[i])) != cudaSuccess) return e;
        if ((e = cudaEventCreate(&g_prof.stop[i])) != cudaSuccess) return e;
        g_prof.made[i] = true;
    }
    *slot = i;
    return cudaSuccess;
}

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

inline long long tiles_for(long long A, int tile) { return A <= 0 ? 0 : (A + tile - 1) / tile; }

// bytes per list tile / list tiles per scan iteration of a configuration (ScanConfig::LT / TPI)
inline int list_tile_of(int cfg)
{
    const int st = kCfg[cfg].threads * kCfg[cfg].cpt * 16;
    return st > 16384 ? 16384 : st;
}
inline int tpi_of(int cfg) { return kCfg[cfg].threads * kCfg[cfg].cpt * 16 / list_tile_of(cfg); }

inline int cfg_of(uint32_t flags)
{
    const int cfg = int((flags >> 8) & 15u);
    return cfg < N_CFG ? cfg : 0;
}

// workspace layout
struct Workspace {
    ParseState* st;
    unsigned int* rangetot;
    unsigned long long* rprefix;
    unsigned int* lprefix;
    unsigned short* lists;
    int slot_cap;
    GeneralArrays g;
    unsigned long long *spec_desc, *spec_pe, *spec_xx;  // speculative general pass: one entry per chunk of GS_TC tiles
    long long spec_chunks;
    size_t total;
};

Workspace carve(void* base, long long len, long long max_lines, uint32_t flags)
{
    Workspace w;
    size_t off = 0;
    uint8_t* b = static_cast<uint8_t*>(base);
    const int cfg = cfg_of(flags);
    const int tile = list_tile_of(cfg);
    w.slot_cap = (flags & FQB_FLAG_DENSE) ? tile : tile / 8;
    w.st = reinterpret_cast<ParseState*>(b + off);
    off += align256(sizeof(ParseState));
    w.rangetot = reinterpret_cast<unsigned int*>(b + off);
    off += align256(size_t(MAX_GRID) * 4);
    w.rprefix = reinterpret_cast<unsigned long long*>(b + off);
    off += align256(size_t(MAX_GRID + 1) * 8);
    const long long nt = tiles_for(len + 16, tile) + 1;
    w.lprefix = reinterpret_cast<unsigned int*>(b + off);
    off += align256(size_t(nt) * 4);
    w.lists = reinterpret_cast<unsigned short*>(b + off);
    off += align256(size_t(nt) * size_t(w.slot_cap) * 2);
    off = carve_general(w.g, b, off, max_lines);
    w.spec_desc = w.spec_pe = w.spec_xx = nullptr;
    w.spec_chunks = 0;
    if (max_lines > 0 || (flags & FQB_FLAG_SPEC_ONLY)) {  // the general path is wanted: room for its speculative pass
        w.spec_chunks = nt + 1;  // one tile per chunk at worst (the device picks the chunk size from the line density)
        w.spec_desc = reinterpret_cast<unsigned long long*>(b + off);
        off += align256(size_t(w.spec_chunks) * 8);
        w.spec_pe = reinterpret_cast<unsigned long long*>(b + off);
        off += align256(size_t(w.spec_chunks) * 8);
        w.spec_xx = reinterpret_cast<unsigned long long*>(b + off);
        off += align256(size_t(w.spec_chunks) * 8);
    }
    w.total = off;
    return w;
}

}  // namespace

extern "C" {

size_t fqb_workspace_bytes(int64_t len, int64_t max_lines, uint32_t flags)
{
    if (len < 0) len = 0;
    if (max_lines < 0) max_lines = 0;
    return carve(nullptr, len, max_lines, flags).total;
}

}  // extern "C"

namespace {

// everything both halves of a parse (scan / emit) derive from the call arguments
struct Geometry {
    Workspace w;
    DevCache* dc;
    int cfg, tile, mis, grid;
    const uint8_t* base;
    long long A, n_tiles, T;
    ListView lv;
};

cudaError_t make_geometry(Geometry& g, const uint8_t* d_buf, int64_t len, int32_t sentinel, void* d_workspace,
                          size_t workspace_bytes, int64_t max_lines, uint32_t flags)
{
    if (len < 0 || max_lines < 0 || !d_workspace) return cudaErrorInvalidValue;
    if (len > 0 && !d_buf) return cudaErrorInvalidValue;
    if (reinterpret_cast<uintptr_t>(d_workspace) & 255) return cudaErrorInvalidValue;
    g.w = carve(d_workspace, len, max_lines, flags);
    if (g.w.total > workspace_bytes) return cudaErrorInvalidValue;
    cudaError_t e = device_cache(&g.dc);
    if (e != cudaSuccess) return e;
    g.cfg = cfg_of(flags);
    g.tile = list_tile_of(g.cfg);
    const uintptr_t addr = reinterpret_cast<uintptr_t>(d_buf);
    g.mis = len > 0 ? int(addr & 15) : 0;
    g.base = len > 0 ? d_buf - g.mis : nullptr;
    g.A = len > 0 ? (long long)g.mis + len : 0;
    g.n_tiles = tiles_for(g.A, g.tile);
    if (g.n_tiles > 0x7ffffff0ll) return cudaErrorInvalidValue;  // > 16 TiB in one call
    int grid = g.dc->sms * g.dc->occ[g.cfg];
    if (grid > MAX_GRID) grid = MAX_GRID;
    if (grid > g.n_tiles) grid = int(g.n_tiles);
    if (grid < 1) grid = 1;
    const int tpi = tpi_of(g.cfg);
    g.T = g.n_tiles > 0 ? (g.n_tiles + grid - 1) / grid : 1;
    g.T = (g.T + tpi - 1) / tpi * tpi;  // whole scan iterations per range
    if (g.n_tiles > 0) grid = int((g.n_tiles + g.T - 1) / g.T);  // no empty ranges
    g.grid = grid;
    memset(&g.lv, 0, sizeof(g.lv));
    g.lv.lists = g.w.lists;
    g.lv.lprefix = g.w.lprefix;
    g.lv.rprefix = g.w.rprefix;
    g.lv.n_tiles = int(g.n_tiles);
    g.lv.T = int(g.T);
    g.lv.slot_cap = g.w.slot_cap;
    g.lv.tile = g.tile;
    g.lv.virt = (sentinel && len > 0) ? 1 : 0;
    g.lv.mis = g.mis;
    return cudaSuccess;
}

// the only pass over the input: newline lists + count prefixes (state is reset first)
// With a Phred mirror whose address is congruent to the buffer's modulo 16 the scan writes the mirror itself
// (fused_decode); otherwise fq_decode_kernel copies the quality spans byte by byte after the emit.
inline bool fused_decode(const Geometry& g, const int8_t* d_qual)
{
    return d_qual && g.A > 0 && ((reinterpret_cast<uintptr_t>(d_qual - g.mis) & 15) == 0);
}

cudaError_t run_scan(const Geometry& g, int32_t sentinel, cudaStream_t stream, int8_t* d_qual = nullptr, int32_t qual_add = 0,
                     bool fasta = false, const ShardTail* tail = nullptr)
{
    cudaError_t e;
    if ((e = cudaMemsetAsync(g.w.st, 0, sizeof(ParseState), stream)) != cudaSuccess) return e;
    ScanParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.base = g.base;
    sp.A = g.A;
    sp.mis = g.mis;
    sp.sentinel = sentinel;
    sp.lists = g.w.lists;
    sp.lprefix = g.w.lprefix;
    sp.rangetot = g.w.rangetot;
    sp.rprefix = g.w.rprefix;
    sp.n_tiles = g.n_tiles;
    sp.T = g.T;
    sp.slot_cap = g.w.slot_cap;
    sp.st = g.w.st;
    sp.qual = fused_decode(g, d_qual) ? d_qual - g.mis : nullptr;
    sp.add4 = (unsigned(qual_add) & 0xffu) * 0x01010101u;
    sp.cls1 = fasta ? '>' : '@';
    sp.cls2 = fasta ? '>' : '+';
    sp.last_visible = fasta ? 1 : 0;
    int slot = -1;
    if (g_prof.on) {
        if ((e = prof_slot(&slot)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(g_prof.start[slot], stream)) != cudaSuccess) return e;
    }
    if ((e = tail ? launch_scan_shard(sp, *tail, g.grid, stream) : launch_scan(g.cfg, sp, g.grid, stream)) != cudaSuccess)
        return e;
    if (slot >= 0) {
        if ((e = cudaEventRecord(g_prof.stop[slot], stream)) != cudaSuccess) return e;
        g_prof.pending += 1;
    }
    return cudaSuccess;
}

// rows from the lists (4-line fast path), tail classification, result header
cudaError_t run_emit(const Geometry& g, int32_t sentinel, int64_t goff, int64_t* d_table, int64_t cap, int8_t* d_qual,
                     int32_t qual_add, fqb_result* d_result, bool want_fast, bool sharded, int64_t own_len,
                     int32_t is_last, const uint64_t* d_line_base, cudaStream_t stream,
                     const uint64_t* d_wait_slots = nullptr, int32_t n_wait = 0, uint64_t epoch = 0)
{
    EmitParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.base = g.base;
    ep.A = g.A;
    ep.mis = g.mis;
    ep.sentinel = sentinel;
    ep.out_bias = (long long)sentinel + goff - g.mis;
    ep.goff = goff;
    ep.table = reinterpret_cast<long long*>(d_table);
    ep.cap = cap;
    ep.lv = g.lv;
    ep.st = g.w.st;
    ep.res = d_result;
    ep.qual = d_qual;
    ep.qual_add = unsigned(qual_add) & 0xffu;
    ep.force_general = want_fast ? 0 : 1;
    ep.line_base = reinterpret_cast<const unsigned long long*>(d_line_base);
    ep.own_end = sharded ? (long long)g.mis + own_len : g.A;
    ep.is_last = sharded ? (is_last ? 1 : 0) : 1;
    ep.sharded = sharded ? 1 : 0;
    ep.wait_slots = reinterpret_cast<const unsigned long long*>(d_wait_slots);
    ep.n_wait = n_wait;
    ep.epoch = epoch;
    long long warps = g.n_tiles > 0 ? g.n_tiles : 1;
    long long blocks = (warps + 7) / 8;
    // persistent grid: one wave of resident CTAs, every warp walks its tiles with the next one prefetched
    const int occ = g.dc->occ_emit[0];
    const long long maxb = (long long)g.dc->sms * (occ > 0 ? occ : 4);
    if (blocks > maxb) blocks = maxb;
    if (!want_fast) blocks = 1;
    // programmatic dependent launch: the CTAs are dispatched while the kernel before this one (the scan) drains and wait
    // there (griddepcontrol.wait) -- no launch gap between scan and emit
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3((unsigned int)blocks);
    lc.blockDim = dim3(256);
    lc.dynamicSmemBytes = 0;
    lc.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&lc, fq_emit_kernel, ep);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess || !d_qual || !want_fast || g.n_tiles == 0 || fused_decode(g, d_qual)) return e;
    // Phred decode of an unaligned mirror: one CTA per tile, one resident wave
    const int occ_d = g.dc->occ_emit[1];
    long long dblocks = (long long)g.dc->sms * (occ_d > 0 ? occ_d : 4);
    if (dblocks > g.n_tiles) dblocks = g.n_tiles;
    fq_decode_kernel<<<int(dblocks), DEC_THREADS, 0, stream>>>(ep);
    return cudaGetLastError();
}

// which speculative pass: FQB_FLAG_SPEC_V1 or the environment variable FQB200_SPEC=1 pick the CTA-per-chunk kernel
bool spec_v1(uint32_t flags)
{
    static const int env = [] {
        const char* v = getenv("FQB200_SPEC");
        return (v && v[0] == '1') ? 1 : 0;
    }();
    return (flags & FQB_FLAG_SPEC_V1) || env;
}

}  // namespace

extern "C" {

int fqb_parse(const uint8_t* d_buf, int64_t len, int32_t sentinel, int64_t goff, int64_t* d_table, int64_t cap,
              int8_t* d_qual, int32_t qual_add, fqb_result* d_result, void* d_workspace, size_t workspace_bytes,
              int64_t max_lines, uint32_t flags, void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (cap < 0 || !d_result) return cudaErrorInvalidValue;
    if (cap > 0 && (!d_table || (reinterpret_cast<uintptr_t>(d_table) & 15))) return cudaErrorInvalidValue;
    if (max_lines > 0xfffffff0ll) max_lines = 0xfffffff0ll;
    sentinel = sentinel ? 1 : 0;
    Geometry g;
    cudaError_t e = make_geometry(g, d_buf, len, sentinel, d_workspace, workspace_bytes, max_lines, flags);
    if (e != cudaSuccess) return e;
    const bool want_fast = !(flags & FQB_FLAG_FORCE_GENERAL);
    const bool want_general = !(flags & (FQB_FLAG_FAST_ONLY | FQB_FLAG_SPEC_ONLY)) && max_lines > 0;
    const bool want_spec = !(flags & (FQB_FLAG_FAST_ONLY | FQB_FLAG_NO_SPEC)) && (want_general || (flags & FQB_FLAG_SPEC_ONLY));

    if ((e = run_scan(g, sentinel, stream, d_qual, qual_add)) != cudaSuccess) return e;
    if ((e = run_emit(g, sentinel, goff, d_table, cap, d_qual, qual_add, d_result, want_fast, false, 0, 1, nullptr,
                      stream)) != cudaSuccess)
        return e;
    if (want_spec && g.n_tiles > 0) {
        // speculative single pass over the newline lists (fq_gspec.cuh); declines -> the exact path below runs
        SpecParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.base = g.base;
        sp.A = g.A;
        sp.mis = g.mis;
        sp.sentinel = sentinel;
        sp.goff = goff;
        sp.table = reinterpret_cast<long long*>(d_table);
        sp.cap = cap;
        sp.st = g.w.st;
        sp.res = d_result;
        sp.lv = g.lv;
        sp.desc = g.w.spec_desc;
        sp.pe = g.w.spec_pe;
        sp.xx = g.w.spec_xx;
        sp.n_chunks = int(g.n_tiles);  // upper bound; the kernel derives the real number from the line count
        // (the warp-per-chunk pass parks rows in the upper half of the list slots: 16 KiB list tiles)
        if (spec_v1(flags) || sp.lv.slot_cap < G2_PARK_AT + G2_PARK_ROWS * 8) {  // one CTA per chunk (fq_gspec.cuh)
            if ((e = cudaMemsetAsync(sp.desc, 0, size_t(sp.n_chunks) * 8, stream)) != cudaSuccess) return e;
            int blocks = g.dc->sms * (g.dc->occ_spec > 0 ? g.dc->occ_spec : 4);
            if (blocks > sp.n_chunks) blocks = sp.n_chunks;
            fq_gspec_kernel<<<blocks, GS_THREADS, GS_SMEM, stream>>>(sp);
        } else {  // one warp per chunk (fq_gspec2.cuh): the chunk descriptors and the block-level scratch (sp.pe) start at zero
            const size_t zero = size_t(reinterpret_cast<const char*>(sp.xx) - reinterpret_cast<const char*>(sp.desc));
            if ((e = cudaMemsetAsync(sp.desc, 0, zero, stream)) != cudaSuccess) return e;
            int blocks = g.dc->sms * (g.dc->occ_spec2 > 0 ? g.dc->occ_spec2 : 4);
            const int need = (sp.n_chunks + G2_WARPS - 1) / G2_WARPS;
            if (blocks > need) blocks = need;
            fq_gspec2_kernel<<<blocks, G2_THREADS, G2_SMEM, stream>>>(sp);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (want_general || (want_spec && d_qual && g.n_tiles > 0)) {
        GeneralParams gp;
        memset(&gp, 0, sizeof(gp));
        gp.base = g.base;
        gp.A = g.A;
        gp.mis = g.mis;
        gp.sentinel = sentinel;
        gp.goff = goff;
        gp.table = reinterpret_cast<long long*>(d_table);
        gp.cap = cap;
        gp.st = g.w.st;
        gp.res = d_result;
        gp.g = g.w.g;
        gp.max_lines = (unsigned long long)max_lines;
        gp.qual = d_qual;
        gp.qual_add = uint8_t(unsigned(qual_add) & 0xffu);
        gp.lv = g.lv;
        if (want_general) {
            if ((e = launch_general(gp, g.dc->sms, stream)) != cudaSuccess) return e;
        } else {  // FQB_FLAG_SPEC_ONLY with a Phred mirror: the general path's decode of the stored records
            fq_g_decode_kernel<<<g.dc->sms * 8, 256, 0, stream>>>(gp);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

// FQB_FLAG_SHARD_TAIL: count / publish / signal in the epilogue of the scan's last CTA instead of two small kernels
// after it.  Same results (tests/test_shard.py runs both), same step time on 2 GPUs (0.2768 against 0.2771 ms: the
// small kernels hide in the launch queue), so the form validated at 4 and 8 GPUs stays the default.
static int shard_scan_impl(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, uint64_t* d_own_lines,
                           uint64_t* const* pub_slots, int32_t n_pub, uint64_t epoch, void* d_workspace,
                           size_t workspace_bytes, uint32_t flags, void* stream_, int8_t* d_qual = nullptr,
                           int32_t qual_add = 0, uint64_t* d_ready_left = nullptr, uint64_t ready_epoch = 0)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!d_own_lines || own_len < 0 || own_len > len) return cudaErrorInvalidValue;
    if (n_pub < 0 || n_pub > 16 || (n_pub > 0 && (!pub_slots || epoch == 0))) return cudaErrorInvalidValue;
    sentinel = sentinel ? 1 : 0;
    Geometry g;
    cudaError_t e = make_geometry(g, d_buf, len, sentinel, d_workspace, workspace_bytes, 0, flags);
    if (e != cudaSuccess) return e;
    // Phred mirror of own bytes + halo: written by the scan itself, which needs a mirror congruent to the buffer
    if (d_qual && g.A > 0 && !fused_decode(g, d_qual)) return cudaErrorInvalidValue;
    if (g.cfg == 0 && g.n_tiles > 0 && (flags & FQB_FLAG_SHARD_TAIL)) {  // one kernel: the scan's last CTA counts, publishes, signals
        ShardTail tail;
        memset(&tail, 0, sizeof(tail));
        tail.own_end = (long long)g.mis + own_len;
        tail.own_lines = reinterpret_cast<unsigned long long*>(d_own_lines);
        for (int i = 0; i < n_pub; ++i) tail.pub[i] = reinterpret_cast<unsigned long long*>(pub_slots[i]);
        tail.n_pub = n_pub;
        tail.epoch = epoch;
        tail.ready_left = reinterpret_cast<unsigned long long*>(d_ready_left);
        tail.ready_epoch = ready_epoch;
        return run_scan(g, sentinel, stream, d_qual, qual_add, false, &tail);
    }
    if ((e = run_scan(g, sentinel, stream, d_qual, qual_add)) != cudaSuccess) return e;
    PubList pub;
    memset(&pub, 0, sizeof(pub));
    for (int i = 0; i < n_pub; ++i) pub.p[i] = reinterpret_cast<unsigned long long*>(pub_slots[i]);
    // counts, publishes and (d_ready_left) signals: one small kernel
    fq_own_lines_kernel<<<1, 32, 0, stream>>>(g.lv, g.w.st, (long long)g.mis + own_len,
                                              reinterpret_cast<unsigned long long*>(d_own_lines), pub, n_pub, epoch,
                                              reinterpret_cast<unsigned long long*>(d_ready_left), ready_epoch);
    return cudaGetLastError();
}

static int shard_emit_impl(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, int32_t is_last, int64_t goff,
                           const uint64_t* d_line_base, const uint64_t* d_wait_slots, int32_t n_wait, uint64_t epoch,
                           int64_t* d_table, int64_t cap, fqb_result* d_result, void* d_workspace, size_t workspace_bytes,
                           uint32_t flags, void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (cap < 0 || !d_result || own_len < 0 || own_len > len) return cudaErrorInvalidValue;
    if (cap > 0 && (!d_table || (reinterpret_cast<uintptr_t>(d_table) & 15))) return cudaErrorInvalidValue;
    sentinel = sentinel ? 1 : 0;
    Geometry g;
    cudaError_t e = make_geometry(g, d_buf, len, sentinel, d_workspace, workspace_bytes, 0, flags);
    if (e != cudaSuccess) return e;
    return run_emit(g, sentinel, goff, d_table, cap, nullptr, 0, d_result, true, true, own_len, is_last, d_line_base, stream,
                    d_wait_slots, n_wait, epoch);
}

int fqb_shard_scan(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, uint64_t* d_own_lines,
                   void* d_workspace, size_t workspace_bytes, uint32_t flags, void* stream)
{
    return shard_scan_impl(d_buf, len, own_len, sentinel, d_own_lines, nullptr, 0, 0, d_workspace, workspace_bytes, flags,
                           stream);
}

int fqb_shard_scan_publish(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, uint64_t* d_own_lines,
                           uint64_t* const* pub_slots, int32_t n_pub, uint64_t epoch, void* d_workspace,
                           size_t workspace_bytes, uint32_t flags, void* stream)
{
    return shard_scan_impl(d_buf, len, own_len, sentinel, d_own_lines, pub_slots, n_pub, epoch, d_workspace,
                           workspace_bytes, flags, stream);
}

int fqb_shard_scan_publish_ready(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, uint64_t* d_own_lines,
                                 uint64_t* const* pub_slots, int32_t n_pub, uint64_t epoch, uint64_t* d_ready_left,
                                 uint64_t ready_epoch, int8_t* d_qual, int32_t qual_add, void* d_workspace,
                                 size_t workspace_bytes, uint32_t flags, void* stream)
{
    return shard_scan_impl(d_buf, len, own_len, sentinel, d_own_lines, pub_slots, n_pub, epoch, d_workspace,
                           workspace_bytes, flags, stream, d_qual, qual_add, d_ready_left, ready_epoch);
}

int fqb_shard_scan_decode(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, uint64_t* d_own_lines,
                          uint64_t* const* pub_slots, int32_t n_pub, uint64_t epoch, int8_t* d_qual, int32_t qual_add,
                          void* d_workspace, size_t workspace_bytes, uint32_t flags, void* stream)
{
    if (!d_qual) return cudaErrorInvalidValue;
    return shard_scan_impl(d_buf, len, own_len, sentinel, d_own_lines, pub_slots, n_pub, epoch, d_workspace,
                           workspace_bytes, flags, stream, d_qual, qual_add);
}

int fqb_shard_emit(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, int32_t is_last, int64_t goff,
                   const uint64_t* d_line_base, int64_t* d_table, int64_t cap, fqb_result* d_result, void* d_workspace,
                   size_t workspace_bytes, uint32_t flags, void* stream)
{
    if (!d_line_base) return cudaErrorInvalidValue;
    return shard_emit_impl(d_buf, len, own_len, sentinel, is_last, goff, d_line_base, nullptr, 0, 0, d_table, cap, d_result,
                           d_workspace, workspace_bytes, flags, stream);
}

int fqb_shard_emit_wait(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, int32_t is_last, int64_t goff,
                        const uint64_t* d_wait_slots, int32_t n_wait, uint64_t epoch, int64_t* d_table, int64_t cap,
                        fqb_result* d_result, void* d_workspace, size_t workspace_bytes, uint32_t flags, void* stream)
{
    if (n_wait < 0 || n_wait > 16 || epoch == 0 || (n_wait > 0 && !d_wait_slots)) return cudaErrorInvalidValue;
    // n_wait == 0 (the first shard): nothing to wait for, its line base is 0
    return shard_emit_impl(d_buf, len, own_len, sentinel, is_last, goff, nullptr, n_wait ? d_wait_slots : nullptr, n_wait,
                           epoch, d_table, cap, d_result, d_workspace, workspace_bytes, flags, stream);
}

int fqb_shard_general(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, int32_t is_first, int32_t is_last,
                      int64_t goff, const uint64_t* d_slot_local, uint64_t* d_slot_right, uint64_t* d_slot_left, uint64_t epoch,
                      uint64_t prev_epoch, int64_t* d_table, int64_t cap, fqb_result* d_result, void* d_workspace,
                      size_t workspace_bytes, int64_t max_lines, uint32_t flags, void* stream_)
{
    const uint64_t* d_entry_slot = d_slot_local;
    uint64_t* d_exit_slot = is_last ? nullptr : d_slot_right;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (cap < 0 || !d_result || own_len < 0 || own_len > len || max_lines < 1 || epoch == 0) return cudaErrorInvalidValue;
    if (cap > 0 && (!d_table || (reinterpret_cast<uintptr_t>(d_table) & 15))) return cudaErrorInvalidValue;
    if (!d_slot_local || (!is_last && !d_slot_right) || (!is_first && !d_slot_left) || prev_epoch >= epoch)
        return cudaErrorInvalidValue;
    if (max_lines > 0xfffffff0ll) max_lines = 0xfffffff0ll;
    sentinel = sentinel ? 1 : 0;
    Geometry g;
    cudaError_t e = make_geometry(g, d_buf, len, sentinel, d_workspace, workspace_bytes, max_lines, flags);
    if (e != cudaSuccess) return e;
    if ((e = run_scan(g, sentinel, stream)) != cudaSuccess) return e;
    // hands the buffer to the general path (need_general = 1)
    if ((e = run_emit(g, sentinel, goff, d_table, cap, nullptr, 0, d_result, false, false, 0, 1, nullptr, stream)) != cudaSuccess)
        return e;
    GeneralParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.base = g.base;
    gp.A = g.A;
    gp.mis = g.mis;
    gp.sentinel = sentinel;
    gp.goff = goff;
    gp.table = reinterpret_cast<long long*>(d_table);
    gp.cap = cap;
    gp.st = g.w.st;
    gp.res = d_result;
    gp.g = g.w.g;
    gp.max_lines = (unsigned long long)max_lines;
    gp.lv = g.lv;
    gp.sharded = 1;
    gp.is_first = is_first ? 1 : 0;
    gp.is_last = is_last ? 1 : 0;
    gp.own_end_blob = (long long)sentinel + own_len;
    gp.entry_slot = reinterpret_cast<const unsigned long long*>(d_entry_slot);
    gp.exit_slot = reinterpret_cast<unsigned long long*>(d_exit_slot);
    gp.ack_left = is_first ? nullptr : reinterpret_cast<unsigned long long*>(d_slot_left) + 4;
    gp.epoch = epoch;
    gp.prev_epoch = prev_epoch;
    return launch_general(gp, g.dc->sms, stream);
}

int fqb_shard_pull_halo(uint8_t* d_halo_dst, const uint8_t* d_peer_src, int64_t halo_bytes, const uint64_t* d_ready_local,
                        uint64_t* d_ready_left, uint64_t epoch, int32_t* d_status, void* stream)
{
    if (halo_bytes < 0 || epoch == 0) return cudaErrorInvalidValue;
    if (halo_bytes > 0 && (!d_halo_dst || !d_peer_src || !d_ready_local)) return cudaErrorInvalidValue;
    if (halo_bytes == 0 && !d_ready_left) return cudaSuccess;
    // one 16-byte load per thread in flight where possible: the copy is bound by the NVLink round trip
    int blocks = int((halo_bytes / 16 + 255) / 256);
    if (blocks < 1) blocks = 1;
    if (blocks > 1024) blocks = 1024;
    fq_halo_pull_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_halo_dst, d_peer_src, halo_bytes, reinterpret_cast<const unsigned long long*>(d_ready_local),
        reinterpret_cast<unsigned long long*>(d_ready_left), epoch, d_status);
    return cudaGetLastError();
}

int fqb_shard_wait_ready(const uint64_t* d_ready_local, uint64_t epoch, int32_t* d_status, void* stream)
{
    if (!d_ready_local || epoch == 0) return cudaErrorInvalidValue;
    // A stream memory operation (cuStreamWaitValue64, >=): the wait occupies no SM at all.  A resident spinning CTA --
    // even one warp -- takes a CTA slot's worth of shared memory, and the scan kernel's grid is sized to fill every SM
    // to the last kilobyte: with the waiting CTA resident one scan CTA has to run behind the others (0.197 -> 0.265 ms
    // per scan on 2 GPUs, profiles/r02_*).  The kernel is the fallback (no driver entry point / not supported; it has
    // the 10 s timeout the memory operation lacks).
    typedef int (*WaitValue64)(cudaStream_t, unsigned long long, unsigned long long, unsigned int);
    static WaitValue64 wait64 = nullptr;
    static int tried = 0;  // 0: not yet, 1: available, 2: unavailable
    {
        std::lock_guard<std::mutex> lock(g_dev_mutex);
        if (tried == 0) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            const char* force = getenv("FQB_WAIT_KERNEL");
            if (!(force && force[0] == '1') &&
                cudaGetDriverEntryPoint("cuStreamWaitValue64", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
                qres == cudaDriverEntryPointSuccess) {
                wait64 = reinterpret_cast<WaitValue64>(fn);
                tried = 1;
            } else {
                (void)cudaGetLastError();
                tried = 2;
            }
        }
    }
    if (tried == 1) {
        const int rc = wait64(static_cast<cudaStream_t>(stream), (unsigned long long)reinterpret_cast<uintptr_t>(d_ready_local),
                              (unsigned long long)epoch, 0u /* CU_STREAM_WAIT_VALUE_GEQ */);
        if (rc == 0) return cudaSuccess;
    }
    fq_wait_ready_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const unsigned long long*>(d_ready_local), epoch, d_status);
    return cudaGetLastError();
}

int fqb_shard_signal_ready(uint64_t* d_ready_left, uint64_t epoch, void* stream)
{
    if (!d_ready_left || epoch == 0) return cudaErrorInvalidValue;
    fq_signal_ready_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned long long*>(d_ready_left),
                                                                            epoch);
    return cudaGetLastError();
}

int fqb_sum_u64_ptrs(const uint64_t* const* ptrs, int32_t n, uint64_t* d_out, void* stream)
{
    if (n < 0 || n > 16 || !d_out || (n > 0 && !ptrs)) return cudaErrorInvalidValue;
    PtrList pl;
    memset(&pl, 0, sizeof(pl));
    for (int i = 0; i < n; ++i) pl.p[i] = reinterpret_cast<const unsigned long long*>(ptrs[i]);
    fq_sum_ptrs_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(pl, n, reinterpret_cast<unsigned long long*>(d_out));
    return cudaGetLastError();
}

int fqb_arrayadd_b(int8_t* d_a, int64_t n, int32_t value, void* stream)
{
    if (n < 0 || (n > 0 && !d_a)) return cudaErrorInvalidValue;
    if (n == 0) return cudaSuccess;
    long long blocks = (n / 16 + 255) / 256 + 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    fq_arrayadd_b_kernel<<<int(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_a, n, (unsigned(value) & 0xffu) * 0x01010101u);
    return cudaGetLastError();
}

int fqb_arrayadd_q(int64_t* d_a, int64_t n, int64_t value, void* stream)
{
    if (n < 0 || (n > 0 && !d_a)) return cudaErrorInvalidValue;
    if (n == 0) return cudaSuccess;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    fq_arrayadd_q_kernel<<<int(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<long long*>(d_a), n, value);
    return cudaGetLastError();
}

int fqb_synth_fixed(uint8_t* d_buf, int64_t n_bytes, int64_t first_byte, int32_t header_len, int32_t read_len,
                    uint64_t seed, void* stream)
{
    if (n_bytes < 0 || first_byte < 0 || header_len < 21 || header_len > 38 || read_len < 1) return cudaErrorInvalidValue;
    if (n_bytes == 0) return cudaSuccess;
    if (!d_buf) return cudaErrorInvalidValue;
    fq_synth_fixed_kernel<<<148 * 16, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_buf, n_bytes, first_byte,
                                                                                   header_len, read_len, seed);
    return cudaGetLastError();
}

int fqb_synth_meta(int32_t kind, uint64_t seed, int64_t k0, int64_t n, const int32_t* d_qtable, int32_t* d_meta,
                   int64_t* d_len, void* stream)
{
    if (kind < 0 || kind > 2 || k0 < 0 || n < 0 || (kind == SYNTH_ONT && !d_qtable)) return cudaErrorInvalidValue;
    if (n == 0) return cudaSuccess;
    if (d_meta && (reinterpret_cast<uintptr_t>(d_meta) & 15)) return cudaErrorInvalidValue;
    fq_synth_meta_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        kind, seed, k0, n, d_qtable, reinterpret_cast<int4*>(d_meta), reinterpret_cast<long long*>(d_len));
    return cudaGetLastError();
}

int fqb_synth_fill(int32_t kind, uint64_t seed, int64_t k0, int64_t n, const int64_t* d_off, const int32_t* d_qtable,
                   uint8_t* d_buf, int64_t first_byte, int64_t n_bytes, void* stream)
{
    if (kind < 0 || kind > 2 || k0 < 0 || n < 0 || n_bytes < 0 || first_byte < 0 || (kind == SYNTH_ONT && !d_qtable))
        return cudaErrorInvalidValue;
    if (n_bytes == 0) return cudaSuccess;
    if (n < 1 || !d_off || !d_buf) return cudaErrorInvalidValue;
    fq_synth_fill_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        kind, seed, k0, n, reinterpret_cast<const long long*>(d_off), d_qtable, d_buf, first_byte, n_bytes);
    return cudaGetLastError();
}

int64_t fqb_synth_host_record(int32_t kind, uint64_t seed, int64_t k, int64_t stream_offset, const int32_t* qtable,
                              uint8_t* out, int64_t cap, int32_t* meta4)
{
    if (kind < 0 || kind > 2 || k < 0 || stream_offset < 0 || cap < 0 || (kind == SYNTH_ONT && !qtable)) return -1;
    SynthRec r;
    synth_rec(kind, seed, (unsigned long long)k, qtable, r);
    const long long n = synth_rec_bytes(r);
    if (meta4) {
        meta4[0] = r.hl;
        meta4[1] = r.rl;
        meta4[2] = r.sb;
        meta4[3] = r.pl;
    }
    char hdr[SYNTH_MAX_HEADER];
    bool have_hdr = false;
    for (long long o = 0; o < n && o < cap; ++o)
        out[o] = synth_byte(kind, seed, (unsigned long long)k, stream_offset + o, o, r, hdr, have_hdr);
    return n;
}

int fqb_kernel_info(int32_t cfg, int32_t* tile_bytes, int32_t* threads, int32_t* stages, int32_t* ctas_per_sm)
{
    if (cfg < 0 || cfg >= N_CFG) return cudaErrorInvalidValue;
    DevCache* dc = nullptr;
    cudaError_t e = device_cache(&dc);
    if (e != cudaSuccess) return e;
    if (tile_bytes) *tile_bytes = kCfg[cfg].threads * kCfg[cfg].cpt * 16;  // bytes per scan iteration
    if (threads) *threads = kCfg[cfg].threads;
    if (stages) *stages = kCfg[cfg].stages;
    if (ctas_per_sm) *ctas_per_sm = dc->occ[cfg];
    return cudaSuccess;
}

int fqb_profile_enable(int32_t on)
{
    cudaError_t e = prof_flush();
    g_prof.on = on != 0;
    g_prof.total_ms = 0.0;
    g_prof.count = 0;
    return e;
}

int fqb_profile_read(double* total_ms, int64_t* launches)
{
    cudaError_t e = prof_flush();
    if (total_ms) *total_ms = g_prof.total_ms;
    if (launches) *launches = g_prof.count;
    return e;
}

const char* fqb_version(void) { return "fqb200 0.1 (sm_100a)"; }

}  // extern "C"

// ---- consumers of the offset table (fq_consume.cuh) ----------------------------------------------
namespace {
inline int blocks_for(long long n, int per_block, int cap)
{
    long long b = (n + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return int(b);
}
inline bool bad_field(int32_t f) { return f < 0 || f > 2; }
}  // namespace

extern "C" {

int fqb_field_lengths(const int64_t* d_table, int64_t n_rows, const int64_t* d_sel, int64_t n_sel, int32_t field,
                      int64_t* d_len, int32_t* d_status, void* stream)
{
    if (n_rows < 0 || n_sel < 0 || bad_field(field) || (n_sel > 0 && (!d_len || (!d_table && n_rows > 0))))
        return cudaErrorInvalidValue;
    if (n_sel == 0) return cudaSuccess;
    if (!d_sel && n_sel != n_rows) return cudaErrorInvalidValue;
    fq_field_lengths_kernel<<<blocks_for(n_sel, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(d_table), n_rows, reinterpret_cast<const long long*>(d_sel), n_sel, field, 0, 0,
        0, reinterpret_cast<long long*>(d_len), d_status);
    return cudaGetLastError();
}

int fqb_length_flags(const int64_t* d_table, int64_t n_rows, int32_t field, int64_t min_len, int64_t max_len,
                     int64_t* d_flags, int32_t* d_status, void* stream)
{
    if (n_rows < 0 || bad_field(field) || (n_rows > 0 && (!d_flags || !d_table))) return cudaErrorInvalidValue;
    if (n_rows == 0) return cudaSuccess;
    fq_field_lengths_kernel<<<blocks_for(n_rows, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(d_table), n_rows, nullptr, n_rows, field, 1, min_len, max_len,
        reinterpret_cast<long long*>(d_flags), d_status);
    return cudaGetLastError();
}

size_t fqb_scan_workspace_bytes(int64_t n)
{
    if (n < 0) n = 0;
    const long long nb = (n + PS_BLOCK - 1) / PS_BLOCK;
    return align256(size_t(nb + 1) * 8);
}

int fqb_exclusive_scan(const int64_t* d_in, int64_t n, int64_t* d_out, void* d_workspace, size_t workspace_bytes,
                       void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n < 0 || !d_out || (n > 0 && !d_in)) return cudaErrorInvalidValue;
    if (n == 0) return cudaMemsetAsync(d_out, 0, 8, stream);
    if (!d_workspace || workspace_bytes < fqb_scan_workspace_bytes(n) || (reinterpret_cast<uintptr_t>(d_workspace) & 7))
        return cudaErrorInvalidValue;
    const long long nb = (n + PS_BLOCK - 1) / PS_BLOCK;
    if (nb > 0x7fffffffll) return cudaErrorInvalidValue;
    long long* bsum = static_cast<long long*>(d_workspace);
    fq_ps_sums_kernel<<<int(nb), PS_THREADS, 0, stream>>>(reinterpret_cast<const long long*>(d_in), n, bsum);
    fq_ps_top_kernel<<<1, PS_THREADS, 0, stream>>>(bsum, nb);
    fq_ps_apply_kernel<<<int(nb), PS_THREADS, 0, stream>>>(reinterpret_cast<const long long*>(d_in), n, bsum, nb,
                                                            reinterpret_cast<long long*>(d_out));
    return cudaGetLastError();
}

int fqb_compact_indices(const int64_t* d_excl, int64_t n, int64_t* d_idx, void* stream)
{
    if (n < 0 || (n > 0 && (!d_excl || !d_idx))) return cudaErrorInvalidValue;
    if (n == 0) return cudaSuccess;
    fq_compact_kernel<<<blocks_for(n, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(d_excl), n, reinterpret_cast<long long*>(d_idx));
    return cudaGetLastError();
}

static int fill_gather(GatherParams& gp, const uint8_t* d_buf, int64_t len, int64_t sub, const int64_t* d_table,
                       int64_t n_rows, const int64_t* d_sel, int64_t n_sel, int32_t field, int32_t add, int32_t* d_status)
{
    if (len < 0 || n_rows < 0 || n_sel < 0 || bad_field(field)) return cudaErrorInvalidValue;
    if (n_sel > 0 && !d_sel && n_sel != n_rows) return cudaErrorInvalidValue;
    if (n_sel > 0 && ((!d_table && n_rows > 0) || (!d_buf && len > 0))) return cudaErrorInvalidValue;
    memset(&gp, 0, sizeof(gp));
    gp.buf = d_buf;
    gp.len = len;
    gp.sub = sub;
    gp.table = reinterpret_cast<const long long*>(d_table);
    gp.n_rows = n_rows;
    gp.sel = reinterpret_cast<const long long*>(d_sel);
    gp.n_sel = n_sel;
    gp.field = field;
    gp.add4 = (unsigned(add) & 0xffu) * 0x01010101u;
    gp.status = d_status;
    return cudaSuccess;
}

int fqb_gather_fields(const uint8_t* d_buf, int64_t len, int64_t table_base, const int64_t* d_table, int64_t n_rows,
                      const int64_t* d_sel, int64_t n_sel, int32_t field, const int64_t* d_offsets, uint8_t* d_out,
                      int32_t add, int32_t* d_status, void* stream)
{
    GatherParams gp;
    int e = fill_gather(gp, d_buf, len, table_base, d_table, n_rows, d_sel, n_sel, field, add, d_status);
    if (e) return e;
    if (n_sel == 0) return cudaSuccess;
    if (!d_offsets) return cudaErrorInvalidValue;
    gp.offsets = reinterpret_cast<const long long*>(d_offsets);
    gp.out = d_out;
    fq_gather_fields_kernel<<<blocks_for(n_sel, 256, 148 * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(gp);
    return cudaGetLastError();
}

int fqb_field_sums(const uint8_t* d_buf, int64_t len, int64_t table_base, const int64_t* d_table, int64_t n_rows,
                   const int64_t* d_sel, int64_t n_sel, int32_t field, int32_t add, int64_t* d_sums, int32_t* d_status,
                   void* stream)
{
    GatherParams gp;
    int e = fill_gather(gp, d_buf, len, table_base, d_table, n_rows, d_sel, n_sel, field, add, d_status);
    if (e) return e;
    if (n_sel == 0) return cudaSuccess;
    if (!d_sums) return cudaErrorInvalidValue;
    fq_field_sums_kernel<<<blocks_for(n_sel, 256, 148 * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        gp, reinterpret_cast<long long*>(d_sums));
    return cudaGetLastError();
}

int fqb_pack_2bit(const uint8_t* d_buf, int64_t len, int64_t table_base, const int64_t* d_table, int64_t n_rows,
                  const int64_t* d_sel, int64_t n_sel, const int64_t* d_offsets, uint8_t* d_out, int64_t* d_n_bases,
                  int64_t* d_n_other, int32_t* d_status, void* stream)
{
    GatherParams gp;
    int e = fill_gather(gp, d_buf, len, table_base, d_table, n_rows, d_sel, n_sel, 1, 0, d_status);
    if (e) return e;
    if (n_sel == 0) return cudaSuccess;
    if (!d_offsets || (reinterpret_cast<uintptr_t>(d_out) & 3)) return cudaErrorInvalidValue;
    gp.offsets = reinterpret_cast<const long long*>(d_offsets);
    gp.out = d_out;
    fq_pack2_kernel<<<blocks_for(n_sel, 256, 148 * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        gp, reinterpret_cast<long long*>(d_n_bases), reinterpret_cast<long long*>(d_n_other));
    return cudaGetLastError();
}

}  // extern "C"

// ---- FASTA (fq_fasta.cuh) ------------------------------------------------------------------------
extern "C" {

// FASTA runs scan configuration 2 (32 KiB per iteration: 2.5 % faster at FASTA's line density) unless the caller
// asks for the 16 KiB geometry with FQB_FLAG_CFG(1); its scan instance exists for these two
static uint32_t fasta_flags(uint32_t flags)
{
    const uint32_t cfg = (flags >> 8) & 15u;
    return (flags & ~FQB_FLAG_CFG(15)) | (cfg == 1 ? 0u : FQB_FLAG_CFG(2));
}

size_t fqb_fasta_workspace_bytes(int64_t len, int64_t max_lines, uint32_t flags)
{
    flags = fasta_flags(flags);
    if (len < 0) len = 0;
    if (max_lines < 0) max_lines = 0;
    const size_t nt = size_t(tiles_for(len + 16, list_tile_of(0)) + 1);
    const size_t tilemax = align256(nt * 8) + align256((nt / FA_GROUP + 2) * 8) + align256(nt * 4);
    // parse workspace | flag byte per rank | on-chain ranks per tile (+ total) | its prefix-sum workspace | running maxima
    return carve(nullptr, len, 0, flags).total + align256(size_t(max_lines + 1)) + align256((nt + 1) * 8) +
           align256(fqb_scan_workspace_bytes(int64_t(nt + 1))) + tilemax;
}

int fqb_parse_fasta(const uint8_t* d_buf, int64_t len, int32_t sentinel, int64_t goff, int64_t* d_table, int64_t cap,
                    fqb_result* d_result, void* d_workspace, size_t workspace_bytes, int64_t max_lines, uint32_t flags,
                    void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (cap < 0 || !d_result || max_lines < 1 || max_lines > 0x7ffffff0ll) return cudaErrorInvalidValue;
    if (cap > 0 && !d_table) return cudaErrorInvalidValue;
    if (workspace_bytes < fqb_fasta_workspace_bytes(len, max_lines, flags)) return cudaErrorInvalidValue;
    sentinel = sentinel ? 1 : 0;
    flags = fasta_flags(flags);
    Geometry g;
    cudaError_t e = make_geometry(g, d_buf, len, sentinel, d_workspace, workspace_bytes, 0, flags);
    if (e != cudaSuccess) return e;
    if ((e = run_scan(g, sentinel, stream, nullptr, 0, true)) != cudaSuccess) return e;
    uint8_t* extra = static_cast<uint8_t*>(d_workspace) + g.w.total;
    FastaParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.base = g.base;
    fp.A = g.A;
    fp.mis = g.mis;
    fp.sentinel = sentinel;
    fp.goff = goff;
    fp.table = reinterpret_cast<long long*>(d_table);
    fp.cap = cap;
    fp.lv = g.lv;
    fp.st = g.w.st;
    fp.res = d_result;
    const size_t nt_ws = size_t(tiles_for(len + 16, list_tile_of(0)) + 1);
    fp.flags = extra;
    fp.max_lines = (unsigned long long)max_lines;
    fp.tile_on = reinterpret_cast<long long*>(extra + align256(size_t(max_lines + 1)));
    void* scan_ws = reinterpret_cast<uint8_t*>(fp.tile_on) + align256((nt_ws + 1) * 8);
    fp.tilemax = reinterpret_cast<long long*>(static_cast<uint8_t*>(scan_ws) + align256(fqb_scan_workspace_bytes(int64_t(nt_ws + 1))));
    const int blocks = g.dc->sms * 8;
    fp.groupmax = reinterpret_cast<long long*>(reinterpret_cast<uint8_t*>(fp.tilemax) + align256(nt_ws * 8));
    fp.lead = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(fp.groupmax) + align256((nt_ws / FA_GROUP + 2) * 8));
    const int n_groups = int((g.n_tiles + FA_GROUP - 1) / FA_GROUP);
    fq_fa_count_kernel<<<blocks, 256, 0, stream>>>(fp);
    if (n_groups > 0) {
        fq_fa_groupscan_kernel<<<n_groups, FA_GROUP, 0, stream>>>(fp);
        fq_fa_topscan_kernel<<<1, 1024, 0, stream>>>(fp);
        fq_fa_fixup_kernel<<<int((g.n_tiles + 255) / 256), 256, 0, stream>>>(fp);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // record index of every tile's first on-chain rank (and, in the entry behind the last tile, the number of calls)
    int rc = fqb_exclusive_scan(reinterpret_cast<const int64_t*>(fp.tile_on), int64_t(g.n_tiles),
                                reinterpret_cast<int64_t*>(fp.tile_on), scan_ws,
                                fqb_scan_workspace_bytes(int64_t(nt_ws + 1)), stream_);
    if (rc) return rc;
    fq_fa_rows_kernel<<<blocks, 256, 0, stream>>>(fp);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    fq_fa_result_kernel<<<1, 32, 0, stream>>>(fp);
    return cudaGetLastError();
}

}  // extern "C"
