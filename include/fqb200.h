/*
 * fqb200.h -- C ABI of libfqb200.so, the B200 (sm_100a) FASTQ-buffer parser.
 *
 * This is the drop-in boundary for the ONE hot path of lgautier/fastq-and-furious that the
 * library replaces: the per-record loop of `readfastq_iter` (src/fastqandfurious.py:251-255),
 * i.e. repeated calls of the C extension's `entrypos` (src/_fastqandfurious.c:25-153) followed
 * by `entryfunc_abspos` (src/fastqandfurious.py:186-195), plus the two array helpers
 * `arrayadd_b` / `arrayadd_q` (src/_fastqandfurious.c:161-217).
 *
 * A per-record device call makes no sense, so the boundary is BATCHED: one call walks the whole
 * entrypos chain of a byte buffer that is already resident in device memory and emits the table
 * of 6 x int64 positions per record that the reference would have produced one record at a time.
 *
 * Conventions
 *   - plain C types only; every pointer prefixed d_ is DEVICE memory owned by the caller;
 *   - all work is enqueued on `stream` and is asynchronous with respect to the host; the library
 *     keeps no state between calls and never allocates;
 *   - functions return 0 (cudaSuccess) or a cudaError_t value for launch/argument errors; data
 *     dependent conditions are reported on the device in `fqb_result`.
 */
#ifndef FQB200_H
#define FQB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes, identical to the reference (src/_fastqandfurious.c:7-15,
 * src/fastqandfurious.py:19-27). */
#define FQB_INVALID (-1)
#define FQB_MISSING_SEQHEADER_BEGIN 0 /* POS_HEAD_BEG */
#define FQB_MISSING_SEQHEADER_END 1   /* POS_HEAD_END */
#define FQB_MISSING_SEQ_BEG 2         /* POS_SEQ_BEG  */
#define FQB_MISSING_SEQ_END 3         /* POS_SEQ_END  */
#define FQB_MISSING_QUAL_BEGIN 4      /* POS_QUAL_BEG */
#define FQB_MISSING_QUAL_END 5        /* POS_QUAL_END */
#define FQB_COMPLETE 6
#define FQB_MISSING_QUALHEADER_END 7

/* fqb_result.error */
#define FQB_OK 0
#define FQB_ERR_CAPACITY 1  /* cap < n_records + 1: table content unspecified, n_records is exact */
#define FQB_ERR_WORKSPACE 2 /* general path: more lines than max_lines; n_lines holds the need */
#define FQB_ERR_TOO_MANY_LINES 3 /* general path: more than 2^32 - 16 lines in one call */
#define FQB_ERR_DENSE 4 /* a tile holds more newlines than its list slot: call again with FQB_FLAG_DENSE */
#define FQB_ERR_HALO 5  /* sharded parse: a record runs past the halo (or the last shard is shorter than a record) */
#define FQB_ERR_SHARD_GENERAL 6 /* sharded parse: the input needs the general path (single-buffer calls only) */
#define FQB_ERR_PEER 7 /* sharded parse, fused exchange: an earlier shard did not publish its line count within 10 s */
#define FQB_ERR_OVERRUN 8 /* sharded parse, fused exchange: a count slot already carries a LATER epoch (a peer ran more
                             parses ahead than the ring of slots holds) */

/* fqb_result.path */
#define FQB_PATH_FAST4 1   /* single-pass 4-line kernel, validated */
#define FQB_PATH_GENERAL 2 /* chain resolution from the newline lists (multi-line records, resync, ...): the speculative
                              single pass where its verification holds (fqb_result.reserved[1] == 1), else line table +
                              hierarchical resolution; the results are identical */

/* fqb_parse flags */
#define FQB_FLAG_FORCE_GENERAL 1u /* skip the 4-line fast path */
#define FQB_FLAG_FAST_ONLY 2u     /* do not enqueue the general path; result.need_general tells */
#define FQB_FLAG_DENSE 4u         /* size the per-tile newline lists for one newline per byte */
#define FQB_FLAG_NO_SPEC 8u        /* general path: skip the speculative single pass, resolve the chain exactly */
#define FQB_FLAG_SPEC_ONLY 16u      /* general path: ONLY the speculative pass (no line table, max_lines may be 0); if it
                                      declines, result.need_general = 1 and the caller repeats the call without this flag */
#define FQB_FLAG_SPEC_V1 32u       /* general path: the speculative pass as one CTA per chunk (fq_gspec.cuh) instead of one
                                      warp per chunk (fq_gspec2.cuh); same results, kept for comparison */
#define FQB_FLAG_CFG(i) (((uint32_t)(i) & 15u) << 8) /* scan kernel configuration (tuning) */
#define FQB_FLAG_SHARD_TAIL 0x10000u /* fqb_shard_scan*: count / publish / signal in the scan's epilogue (one kernel) */

/*
 * Device-resident result header written by fqb_parse (128 bytes).
 *
 * It describes the FIRST entrypos call of the chain that did not return COMPLETE -- exactly the
 * information `readfastq_iter` needs to apply its end-of-stream / refill rules
 * (src/fastqandfurious.py:256-279).
 */
typedef struct fqb_result {
    int64_t n_records;     /* COMPLETE records on the chain = rows of the table */
    int64_t resume_offset; /* blob offset that non-COMPLETE call was made with
                              (= pos5 - 1 of the last COMPLETE record, or 0) */
    int64_t tail_pos[6];   /* its posbuffer, blob relative, -1 filled (src/_fastqandfurious.c:57-59) */
    int32_t tail_status;   /* its return value */
    int32_t path;          /* FQB_PATH_* that produced the result */
    int32_t error;         /* FQB_OK or FQB_ERR_* */
    int32_t need_general;  /* FQB_FLAG_FAST_ONLY: 1 if the fast path could not represent the input */
    int64_t n_lines;       /* visible newlines counted by the scan (incl. the sentinel) */
    int64_t first_bad;     /* fast path: first record index that failed validation, or -1 */
    int64_t reserved[4];
} fqb_result;

/*
 * Bytes of device workspace fqb_parse needs for a buffer of `len` bytes with these `flags`
 * (FQB_FLAG_DENSE and FQB_FLAG_CFG matter).  `max_lines` bounds the number of lines the GENERAL
 * path can index (0 = fast path only).  Default sizing: about len/4 + 49 * max_lines bytes.
 */
size_t fqb_workspace_bytes(int64_t len, int64_t max_lines, uint32_t flags);

/*
 * Walk the entrypos chain of one buffer (replaces the loop src/fastqandfurious.py:251-255 with
 * entrypos = _fastqandfurious.entrypos and entryfunc = entryfunc_abspos).
 *
 *   d_buf, len   raw bytes.  The blob the reference would see is
 *                    blob = (sentinel ? "\n" : "") + d_buf[0:len]
 *                (readfastq_iter prepends that '\n' to the first chunk, :245); the sentinel is
 *                virtual, d_buf is never copied.  d_buf may have any alignment.  The 16-byte
 *                aligned blocks containing d_buf[0] and d_buf[len-1] must be readable.
 *   goff         added to every emitted position (the reference's `globaloffset`, -1 for a
 *                stream start) -- this is also what arrayadd_q is for.
 *   d_table,cap  out: row k = [pos0..pos5] + goff of the k-th COMPLETE record, int64[cap][6],
 *                16-byte aligned.  cap must be >= n_records + 1 (one scratch row), else
 *                FQB_ERR_CAPACITY.
 *   d_qual       optional (NULL = off): int8[len]; for every byte i of d_buf inside the quality
 *                span of a stored record, d_qual[i] = (int8)(d_buf[i] + qual_add)  (the
 *                arrayadd_b recipe, src/demo/benchmark.py:161-163, qual_add = -33 for Phred+33).
 *                Other bytes of d_qual are UNSPECIFIED (when d_qual and d_buf are congruent modulo 16 the
 *                scan writes the whole mirror while it has the bytes on chip: cheaper than fetching the
 *                input a second time to pick the quality lines out).
 *   d_result     out: header above.
 *   d_workspace  fqb_workspace_bytes(len, max_lines, flags) bytes, 256-byte aligned.
 */
int fqb_parse(const uint8_t* d_buf, int64_t len, int32_t sentinel, int64_t goff, int64_t* d_table,
              int64_t cap, int8_t* d_qual, int32_t qual_add, fqb_result* d_result, void* d_workspace,
              size_t workspace_bytes, int64_t max_lines, uint32_t flags, void* stream);

/*
 * Byte-range sharding of one stream over several GPUs (one process per GPU).  Shard g
 * holds the bytes [c_g, c_g + own_len) of the stream followed by a HALO: the first bytes of shard
 * g+1 (at least one maximal record + 2 bytes; the reference has the same contract for fbufsize,
 * src/fastqandfurious.py:219-223), received from the neighbour with one NCCL send/recv.  A record
 * belongs to the shard that holds the newline before its '@' (the virtual sentinel for record 0).
 *
 *   fqb_shard_scan  scans own bytes + halo (len = own_len + halo bytes) and writes to *d_own_lines the
 *                   number of lines (visible newlines + sentinel) at offsets < own_len.
 *   -- caller: exclusive prefix of own_lines over the shards (one all-gather of 8 bytes per rank) --
 *   fqb_shard_emit  *d_line_base = lines owned by all earlier shards.  Emits the rows of the records this
 *                   shard owns (positions + goff; pass goff = c_g - sentinel for absolute offsets),
 *                   row 0 = global record fqb_result.reserved[0].  is_last: the buffer ends at the end
 *                   of the stream (no halo), the end-of-stream classification of fqb_parse applies;
 *                   otherwise tail_status is FQB_COMPLETE (the chain continues in the next shard).
 * Same buffer, workspace and flags in both calls (max_lines = 0 sizing).  4-line fast path only:
 * input that needs the general path reports FQB_ERR_SHARD_GENERAL.
 */
int fqb_shard_scan(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, uint64_t* d_own_lines,
                   void* d_workspace, size_t workspace_bytes, uint32_t flags, void* stream);
int fqb_shard_emit(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, int32_t is_last, int64_t goff,
                   const uint64_t* d_line_base, int64_t* d_table, int64_t cap, fqb_result* d_result, void* d_workspace,
                   size_t workspace_bytes, uint32_t flags, void* stream);

/*
 * Fused exchange over peer memory (NVLink / NVSwitch), replacing the all-gather AND the kernels around it:
 *   fqb_shard_scan_publish  = fqb_shard_scan, and the kernel that counts the owned lines also stores
 *                             {count, epoch} (two uint64) through each of the `n_pub` (<= 16) peer-mapped
 *                             pointers in `pub_slots` (HOST array of DEVICE pointers: this shard's slot in the
 *                             memory of every LATER shard), count first, epoch with release semantics.
 *   fqb_shard_emit_wait     = fqb_shard_emit, but instead of *d_line_base the kernel itself waits until the
 *                             `n_wait` slots at d_wait_slots (LOCAL memory, [n_wait][2] uint64, one per EARLIER
 *                             shard) carry `epoch`, and sums their counts.  A peer that never publishes ends
 *                             the wait after 10 s with FQB_ERR_PEER.
 * Use a fresh `epoch` (> 0, increasing) for every parse and a ring of at least as many slot sets as there are
 * shards, indexed by epoch modulo the ring size: when the ready signals (fqb_shard_signal_ready) are sent early,
 * the first shard can run up to (shards - 1) parses ahead of the last one.  The halo has to be in place before
 * the scan (fqb_shard_pull_halo).
 */
/* Step 1 of the fused exchange: tell the left neighbour that this shard's bytes of `epoch` are in place (release
 * store of `epoch` through d_ready_left, a peer-mapped pointer to ITS ready slot; NULL for the first shard), wait
 * until the right neighbour has said so into d_ready_local (this shard's slot; value >= epoch) and copy its first
 * `halo_bytes` bytes (d_peer_src, peer-mapped) to d_halo_dst = d_buf + own_len.  One kernel instead of a barrier
 * and a peer copy.  *d_status (optional) is set to 1 if the neighbour does not show up within 10 s.  The right
 * neighbour must keep its head bytes unchanged until this call has run (as with any halo exchange). */
int fqb_shard_pull_halo(uint8_t* d_halo_dst, const uint8_t* d_peer_src, int64_t halo_bytes, const uint64_t* d_ready_local,
                        uint64_t* d_ready_left, uint64_t epoch, int32_t* d_status, void* stream);
/* The signalling half of fqb_shard_pull_halo on its own (pass d_ready_left = NULL there): enqueue it as soon as the
 * shard's bytes of `epoch` are in place -- e.g. right after the host->device copy of the NEXT buffer, while the
 * current parse is still running -- so that the left neighbour's pull never waits for this shard's pipeline. */
int fqb_shard_signal_ready(uint64_t* d_ready_left, uint64_t epoch, void* stream);
/* The waiting half on its own, one warp: returns (on the stream) once *d_ready_local >= epoch (10 s timeout ->
 * *d_status = 1).  With double-buffered shards the halo of the NEXT parse is brought in on a second stream while the
 * current one is scanned: this call, then a plain peer copy (copy engine, no SMs) of the neighbour's head bytes. */
int fqb_shard_wait_ready(const uint64_t* d_ready_local, uint64_t epoch, int32_t* d_status, void* stream);
int fqb_shard_scan_publish(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, uint64_t* d_own_lines,
                           uint64_t* const* pub_slots, int32_t n_pub, uint64_t epoch, void* d_workspace,
                           size_t workspace_bytes, uint32_t flags, void* stream);
int fqb_shard_emit_wait(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, int32_t is_last, int64_t goff,
                        const uint64_t* d_wait_slots, int32_t n_wait, uint64_t epoch, int64_t* d_table, int64_t cap,
                        fqb_result* d_result, void* d_workspace, size_t workspace_bytes, uint32_t flags, void* stream);
/*
 * Sharded parse with Phred decode (arrayadd_b over every quality string, src/_fastqandfurious.c:161-185 as used in
 * src/demo/benchmark.py:161-163): fqb_shard_scan_publish (n_pub = 0: fqb_shard_scan) whose scan also writes the mirror
 * d_qual[i] = (int8)(d_buf[i] + qual_add) for all `len` bytes of own range + halo -- every record the shard owns lies
 * inside them, so d_qual[pos4 - c_g .. pos5 - c_g) of each emitted row is its decoded quality string (the rest of the
 * mirror is unspecified, as for fqb_parse).  d_qual must be congruent to d_buf modulo 16 (cudaErrorInvalidValue
 * otherwise): the mirror costs one 16-byte store per 16-byte load of the only pass over the input.
 */
int fqb_shard_scan_decode(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, uint64_t* d_own_lines,
                          uint64_t* const* pub_slots, int32_t n_pub, uint64_t epoch, int8_t* d_qual, int32_t qual_add,
                          void* d_workspace, size_t workspace_bytes, uint32_t flags, void* stream);
/*
 * fqb_shard_scan_publish [+ fqb_shard_scan_decode when d_qual != NULL] + fqb_shard_signal_ready(d_ready_left,
 * ready_epoch) in one call, and optionally as ONE kernel: with the default scan configuration the last CTA of the scan, which has just written
 * the count prefixes, counts the shard's own lines, stores {count, epoch} into the later shards and releases the
 * ready signal (d_ready_left may be NULL: no signal) -- when FQB_FLAG_SHARD_TAIL is passed; by default,
 * for other configurations and for empty shards the count and the signal are two small kernels behind the scan.  The
 * results are the same and so is the measured step time; fqb_shard_scan / _publish / _decode follow the same switch.
 */
int fqb_shard_scan_publish_ready(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, uint64_t* d_own_lines,
                                 uint64_t* const* pub_slots, int32_t n_pub, uint64_t epoch, uint64_t* d_ready_left,
                                 uint64_t ready_epoch, int8_t* d_qual, int32_t qual_add, void* d_workspace,
                                 size_t workspace_bytes, uint32_t flags, void* stream);

/*
 * Sharded GENERAL path (multi-line records, damaged entries): every shard resolves the candidate forest of its
 * own bytes + halo independently (scan, line table, successors, level 1); only the entry of the chain into a
 * shard depends on its predecessor.  Every shard has a slot of 8 uint64 in memory its neighbours can reach:
 * words 0-3 = the hand-over {absolute position the search resumes at, records emitted so far, ended (0 no / 1
 * the chain ended / 2 error), epoch}, stored by the PREVIOUS shard's result kernel and awaited by one thread of
 * this shard's head kernel (10 s timeout -> FQB_ERR_PEER); word 4 = the epoch the NEXT shard has consumed (it
 * acknowledges into the slot of the shard that wrote), which the result kernel waits for before it overwrites
 * the next shard's slot (`prev_epoch` = the epoch of this shard's previous hand-over, 0 if none).
 *   d_slot_local   this shard's slot;   d_slot_right / d_slot_left   peer-mapped pointers to the neighbours' slots
 *                  (NULL where there is no neighbour).
 * Rows: the records whose leading newline lies in the shard's own range (fqb_result.n_records of them, global
 * index of row 0 in reserved[0]); tail_status FQB_COMPLETE = the chain continues in the next shard, anything
 * else (or the last shard's result) is the tail of the whole stream.  Workspace as fqb_parse with the same
 * max_lines; the walk / push-down / emission of a shard start when its predecessor has finished, so only the
 * first half of the work overlaps across shards.
 */
int fqb_shard_general(const uint8_t* d_buf, int64_t len, int64_t own_len, int32_t sentinel, int32_t is_first, int32_t is_last,
                      int64_t goff, const uint64_t* d_slot_local, uint64_t* d_slot_right, uint64_t* d_slot_left, uint64_t epoch,
                      uint64_t prev_epoch, int64_t* d_table, int64_t cap, fqb_result* d_result, void* d_workspace,
                      size_t workspace_bytes, int64_t max_lines, uint32_t flags, void* stream);

/* *d_out = sum of the uint64 values behind `n` (<= 16) device pointers (`ptrs` is a HOST array).  The
 * pointers may be peer-mapped memory of other GPUs (NVLink loads): with the counts every shard publishes
 * in symmetric memory this replaces the all-gather before fqb_shard_emit. */
int fqb_sum_u64_ptrs(const uint64_t* const* ptrs, int32_t n, uint64_t* d_out, void* stream);

/* In-place int8 add with two's-complement wrap: d_a[i] += (int8)value
 * (arrayadd_b, src/_fastqandfurious.c:161-185; value -33 decodes Phred+33). */
int fqb_arrayadd_b(int8_t* d_a, int64_t n, int32_t value, void* stream);

/* In-place int64 add: d_a[i] += value  (arrayadd_q, src/_fastqandfurious.c:193-217). */
int fqb_arrayadd_q(int64_t* d_a, int64_t n, int64_t value, void* stream);

/* ---- Consumers of the offset table (index replay, device-side entryfunc work) ----------
 * What users of the reference do per record inside an `entryfunc`, for the whole table at once.
 * `field`: 0 = header buf[pos0+1:pos1], 1 = sequence buf[pos2:pos3], 2 = quality buf[pos4:pos5]
 * (entryfunc, src/fastqandfurious.py:161-171).  `d_sel` = optional int64 row indices (NULL: all rows in
 * order, n_sel == n_rows).  `d_status` = optional int32, OR-ed with 1 (row index outside the table) /
 * 2 (span reversed or outside the buffer); such records yield length 0 / are skipped. */

/* d_len[i] = length of the field of row d_sel[i]  (lengthfilter_entryfunc, doc/user-guide.rst:162-167). */
int fqb_field_lengths(const int64_t* d_table, int64_t n_rows, const int64_t* d_sel, int64_t n_sel, int32_t field,
                      int64_t* d_len, int32_t* d_status, void* stream);
/* d_flags[i] = 1 if min_len <= length of the field of row i <= max_len else 0. */
int fqb_length_flags(const int64_t* d_table, int64_t n_rows, int32_t field, int64_t min_len, int64_t max_len,
                     int64_t* d_flags, int32_t* d_status, void* stream);
/* d_out[i] = d_in[0] + ... + d_in[i-1] for i in [0, n]  (n + 1 outputs; d_out may alias d_in if it has room). */
size_t fqb_scan_workspace_bytes(int64_t n);
int fqb_exclusive_scan(const int64_t* d_in, int64_t n, int64_t* d_out, void* d_workspace, size_t workspace_bytes,
                       void* stream);
/* d_excl = exclusive scan (n + 1 values) of 0/1 flags: d_idx[d_excl[i]] = i for every set flag. */
int fqb_compact_indices(const int64_t* d_excl, int64_t n, int64_t* d_idx, void* stream);
/* Index replay (src/demo/benchmark.py:47-83): d_out[d_offsets[i] : d_offsets[i+1]] = the field bytes of row
 * d_sel[i], each plus (int8)add (add = -33 on field 2: the arrayadd_b recipe, src/demo/benchmark.py:161-163).
 * Table positions are indices into the stream; d_buf[0] is stream position `table_base`.  d_offsets = the
 * exclusive scan of fqb_field_lengths (n_sel + 1 values). */
int fqb_gather_fields(const uint8_t* d_buf, int64_t len, int64_t table_base, const int64_t* d_table, int64_t n_rows,
                      const int64_t* d_sel, int64_t n_sel, int32_t field, const int64_t* d_offsets, uint8_t* d_out,
                      int32_t add, int32_t* d_status, void* stream);
/* d_sums[i] = sum over the field bytes of row d_sel[i] of (int8)(byte + add): the record's total Phred score
 * with add = -33 on field 2 (mean quality = sum / length). */
int fqb_field_sums(const uint8_t* d_buf, int64_t len, int64_t table_base, const int64_t* d_table, int64_t n_rows,
                   const int64_t* d_sel, int64_t n_sel, int32_t field, int32_t add, int64_t* d_sums, int32_t* d_status,
                   void* stream);
/* 2-bit packed sequences (the reference slices the bytes, doc/user-guide.rst:153-180 -- there is no
 * upstream packing to match, the layout is defined here): the bases of field 1 (buf[pos2:pos3], the newlines inside
 * wrapped records skipped) of row d_sel[i], four per byte, base k in bits 2(k % 4).. of byte k / 4 of
 * d_out[d_offsets[i] : d_offsets[i+1]); A/a = 0, C/c = 1, G/g = 2, T/t/U/u = 3, any other byte by the same bit
 * formula (((b >> 1) & 3) ^ ((b >> 2) & 1)) and counted in d_n_other[i] (may be NULL).  d_n_bases[i] (may be NULL) =
 * bases of the record.  A slot is 4 * ceil(L / 16) bytes for a field of L bytes (d_offsets = exclusive scan of that,
 * n_sel + 1 values; d_out 4-byte aligned); the bytes behind the last base are zero. */
int fqb_pack_2bit(const uint8_t* d_buf, int64_t len, int64_t table_base, const int64_t* d_table, int64_t n_rows,
                  const int64_t* d_sel, int64_t n_sel, const int64_t* d_offsets, uint8_t* d_out, int64_t* d_n_bases,
                  int64_t* d_n_other, int32_t* d_status, void* stream);

/* ---- FASTA (entrypos_fasta, src/fastqandfurious.py:103-143) -----------------------------
 * The chain of entrypos_fasta calls over one buffer, each starting at pos3 of the previous record (the '\n'
 * of its closing "\n>").  d_table: int64[cap][4] = [pos0 '>', pos1 header '\n', pos2 first sequence byte,
 * pos3 '\n' before the next '>'] + goff of every COMPLETE call (cap >= n_records + 1); d_result describes the
 * first call that is not COMPLETE (tail_status 0 / 1 / 2 / 3, tail_pos[0..3] with -1 for the entries the
 * reference leaves unassigned, resume_offset = the offset of that call).  `max_lines` >= the number of visible
 * newlines + sentinel (else FQB_ERR_WORKSPACE with n_lines = the need); sentinel / goff as in fqb_parse.
 * flags: FQB_FLAG_DENSE as in fqb_parse; the scan runs 32 KiB per iteration (FASTA lines are dense: 60-80 columns,
 * 2.5 % faster than the 16 KiB geometry of the FASTQ fast path), FQB_FLAG_CFG(1) selects the 16 KiB geometry. */
size_t fqb_fasta_workspace_bytes(int64_t len, int64_t max_lines, uint32_t flags);
int fqb_parse_fasta(const uint8_t* d_buf, int64_t len, int32_t sentinel, int64_t goff, int64_t* d_table, int64_t cap,
                    fqb_result* d_result, void* d_workspace, size_t workspace_bytes, int64_t max_lines, uint32_t flags,
                    void* stream);

/* Synthetic FASTQ generator used by bench.py and the full-size parity tests (not part of the
 * reference): d_buf[i] = byte first_byte + i of an unbounded stream of fixed-geometry records
 * (any window of it can be generated independently, e.g. one shard per GPU).  See DESIGN.md. */
int fqb_synth_fixed(uint8_t* d_buf, int64_t n_bytes, int64_t first_byte, int32_t header_len, int32_t read_len,
                    uint64_t seed, void* stream);

/* Synthetic streams with VARIABLE record geometry (bench.py, full-size parity tests; not part of the reference):
 * BASELINE.json configs[2..4] -- FQB_SYNTH_ILLUMINA: 150 bp reads, '@A00123:45:HXXXXXXXX:<lane>:<tile>:<x>:<y>
 * 1:N:0:ACGTACGT' headers of variable width, NovaSeq-binned qualities with 10 % of the records uniform over
 * '!'..'I'; FQB_SYNTH_ONT: long reads whose length comes from a quantile table (2^12 + 1 ascending int32 entries:
 * clip(Gamma(2, 5000), 200, 500000) for the 10 kb mean of the config), ~130-byte headers, qualities '"'..'S';
 * FQB_SYNTH_MULTILINE: 150-300 bp reads wrapped at 60 columns, the '+' line repeats the header for half of the
 * records.  Everything about record k is a function of (kind, seed, k), byte g of the stream of (seed, g):
 *   fqb_synth_meta  d_len[i] = bytes of record k0 + i, d_meta[i] = {header line, read, field bytes, '+' line}
 *                   lengths (either may be NULL);
 *   -- caller: exclusive prefix sum of d_len = stream offset of every record (pos0 of the true offset table) --
 *   fqb_synth_fill  d_buf[i] = byte first_byte + i of the stream, for a window inside [d_off[0], d_off[n]) where
 *                   d_off[0..n] are the stream offsets of records k0 .. k0 + n (one shard per GPU);
 *   fqb_synth_host_record  the same bytes of ONE record on the host (no device; the CPU tests compare it with the
 *                   numpy twin in tests/fqgen.py): writes min(cap, record bytes) bytes, returns the record bytes. */
#define FQB_SYNTH_ILLUMINA 0
#define FQB_SYNTH_ONT 1
#define FQB_SYNTH_MULTILINE 2
int fqb_synth_meta(int32_t kind, uint64_t seed, int64_t k0, int64_t n, const int32_t* d_qtable, int32_t* d_meta,
                   int64_t* d_len, void* stream);
int fqb_synth_fill(int32_t kind, uint64_t seed, int64_t k0, int64_t n, const int64_t* d_off, const int32_t* d_qtable,
                   uint8_t* d_buf, int64_t first_byte, int64_t n_bytes, void* stream);
int64_t fqb_synth_host_record(int32_t kind, uint64_t seed, int64_t k, int64_t stream_offset, const int32_t* qtable,
                              uint8_t* out, int64_t cap, int32_t* meta4);

/* Library / kernel configuration introspection (for bench.py's roofline record).  `cfg` is the scan
 * kernel configuration selected by bits 8..11 of fqb_parse's flags (0 = default). */
int fqb_kernel_info(int32_t cfg, int32_t* tile_bytes, int32_t* threads, int32_t* stages, int32_t* ctas_per_sm);
const char* fqb_version(void);

/* Timing of the dominant kernel (the FAST4 scan) for the roofline record: while enabled, fqb_parse
 * brackets that kernel with CUDA events on the caller's stream; fqb_profile_read waits for them and
 * returns the summed device time and the number of launches since fqb_profile_enable(1).  Events are
 * created lazily and kept; not thread safe; one device at a time. */
int fqb_profile_enable(int32_t on);
int fqb_profile_read(double* total_ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* FQB200_H */
