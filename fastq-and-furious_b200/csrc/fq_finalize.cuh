// fq_finalize.cuh -- second (tiny) kernel of the 4-line fast path.
//
//  * seam fix-up: one thread per tile boundary completes (pos5) and validates the single record
//    whose newlines are spread over more than one tile;
//  * tail classifier: the last CTA to finish reproduces, for the still-open last record, the status
//    code and posbuffer the reference's entrypos would return (src/_fastqandfurious.c:57-136), from
//    the <= 4 newline positions the scan already stored in the table -- no byte is re-scanned.
#pragma once
#include "fq_common.cuh"

namespace fqb {

struct FinalizeParams {
    const uint8_t* base;
    long long A;
    int mis;
    int sentinel;
    long long out_bias;  // emitted = a + out_bias
    long long goff;      // emitted = blob + goff
    long long* table;
    long long cap;
    const unsigned long long* desc;
    long long n_tiles;
    ParseState* st;
    fqb_result* res;
    unsigned int flags;
    int force_general;  // 1: the fast scan was skipped, hand over to the general path
};

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
    if (p3 + 2 >= L) return ST_NO_QUALHEAD_END;  // (:97-101)
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

__global__ void __launch_bounds__(256) fq_fast4_finalize_kernel(const FinalizeParams p)
{
    if (p.force_general) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            p.st->need_general = 1;
            write_result(p.res, 0, 0, ST_NO_HEAD_BEG, nullptr, FQB_PATH_FAST4, FQB_OK, 1, 0, -1);
        }
        return;
    }
    // visible newlines incl. sentinel
    const unsigned long long M = p.n_tiles > 0 ? (p.desc[p.n_tiles - 1] & LB_VALUE) : 0ull;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;

    // ---- seam fix-up: record straddling into tile t ----
    if (t >= 1 && t < p.n_tiles) {
        const unsigned long long Bt = p.desc[t - 1] & LB_VALUE;
        const unsigned long long Et = p.desc[t] & LB_VALUE;
        if (Bt >= 1 && Et > Bt) {
            const unsigned long long k = (Bt - 1) >> 2;
            if (4 * k + 4 <= M - 1 && (long long)(k + 1) < p.cap) {  // closed record, rows k and k+1 stored
                long long* row = p.table + k * 6;
                const long long p0 = row[0], p1 = row[1], p3 = row[3], p4 = row[4];
                const long long d = row[6] - 1;  // closing newline = pos0 of the next record - 1
                const long long p5 = p4 + p3 - p1 - 1;
                row[5] = p5;
                const uint8_t* b = p.base;  // b[emitted position - out_bias] = that byte
                const long long ob = p.out_bias;
                bool ok = b[p0 - ob] == '@' && b[p1 + 1 - ob] != '\n' && b[p3 + 1 - ob] == '+';
                const long long plus_len = (p4 - 1) - p3;
                if (plus_len > 2 && plus_len != p1 - p0 + 1) ok = false;
                if (d != p5) ok = false;
                if (!ok) {
                    atomicMin(&p.st->first_bad, k);
                    p.st->fast_fail = 1;
                }
            }
        }
    }

    // ---- last CTA done: tail classification + result header ----
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&p.st->done_counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();

    ParseState* st = p.st;
    const long long L = (p.A - p.mis) + p.sentinel;
    const uint8_t* blob0 = p.base + p.mis - p.sentinel;  // address of blob[0]; virtual when sentinel
    st->n_lines = M;
    int fail = *((volatile int*)&st->fast_fail);
    const long long first_bad = (st->first_bad == ~0ull) ? -1 : (long long)st->first_bad;
    if (fail) {
        st->need_general = 1;
        write_result(p.res, 0, 0, ST_NO_HEAD_BEG, nullptr, FQB_PATH_FAST4, FQB_OK, 1, (long long)M, first_bad);
        return;
    }
    st->need_general = 0;
    if (M == 0) {  // no visible newline at all: entrypos finds no "\n@"
        write_result(p.res, 0, 0, ST_NO_HEAD_BEG, nullptr, FQB_PATH_FAST4, FQB_OK, 0, 0, -1);
        return;
    }
    const long long K = (long long)((M - 1) >> 2);  // records closed by a newline
    const int m = int((M - 1) & 3ull);               // newlines after the last closing one
    // the last closed record is only COMPLETE if pos5 + 2 < L (src/_fastqandfurious.c:130), i.e. its
    // closing newline is not blob[L-2]
    const bool last_is_5 = (K >= 1 && m == 0 && blob0[L - 2] == '\n');
    long long n = K - (last_is_5 ? 1 : 0);
    if (K + 1 > p.cap) {
        write_result(p.res, n, 0, ST_NO_HEAD_BEG, nullptr, FQB_PATH_FAST4, FQB_ERR_CAPACITY, 0, (long long)M, -1);
        return;
    }
    long long pos[6];
    int status;
    if (last_is_5) {
        const long long* row = p.table + (K - 1) * 6;
        for (int i = 0; i < 5; ++i) pos[i] = row[i] - p.goff;
        pos[5] = -1;
        status = ST_NO_QUAL_END;
    } else {
        long long* row = p.table + K * 6;
        long long nl[4];
        nl[0] = row[0] - 1 - p.goff;
        if (m >= 1) nl[1] = row[1] - p.goff;
        if (m >= 2) nl[2] = row[3] - p.goff;
        if (m >= 3) nl[3] = row[4] - 1 - p.goff;
        status = classify_tail(blob0, L, nl, m + 1, pos);
        if (status == ST_COMPLETE) {  // last record without a newline after its quality string
            row[5] = pos[5] + p.goff;
            n = K + 1;
            status = ST_NO_HEAD_BEG;  // the next call finds no further "\n@"
            for (int i = 0; i < 6; ++i) pos[i] = -1;
        }
    }
    const long long resume = (n >= 1) ? (p.table[(n - 1) * 6 + 5] - p.goff - 1) : 0;
    write_result(p.res, n, resume, status, pos, FQB_PATH_FAST4, FQB_OK, 0, (long long)M, -1);
}

}  // namespace fqb
