"""Sequential Python MODEL of the two device algorithms (test infrastructure only).

The CUDA kernels cannot run in the CPU-only container, so the *algorithms* they implement --
the tile/rank decomposition of the 4-line fast path with its validation predicates, seam fix-up
and tail classifier (csrc/fq_scan.cuh, csrc/fq_finalize.cuh) and the line-table / successor /
chain formulation of the general path (csrc/fq_general.cuh) -- are restated here step by step and
property-tested against the oracle (tests/test_algo_model.py).  Nothing in the product imports this.
"""

CLS_OTHER, CLS_AT, CLS_PLUS, CLS_NL = 0, 1, 2, 3
NONE = None


def classify(b):
    return CLS_AT if b == 0x40 else CLS_PLUS if b == 0x2b else CLS_NL if b == 0x0a else CLS_OTHER


def visible_newlines(data, sentinel):
    """(positions, classes) of the newlines the reference can see, blob coordinates.

    blob = ('\\n' if sentinel else '') + data; the last byte of the blob is never seen as a newline
    (src/_fastqandfurious.c:71,103: memchr windows exclude it; "\\n@"/"\\n+" need a second byte)."""
    blob = (b'\n' if sentinel else b'') + bytes(data)
    L = len(blob)
    pos = [i for i in range(L - 1) if blob[i] == 0x0a]
    cls = [classify(blob[i + 1]) for i in pos]
    return blob, pos, cls


def classify_tail(blob, nl):
    """entrypos on the open last record given ALL visible newlines from the search offset on."""
    L = len(blob)
    pos = [-1] * 6
    cnt = len(nl)
    i = 0
    while i < cnt and blob[nl[i] + 1] != 0x40:
        i += 1
    if i == cnt:
        return 0, pos
    p0 = nl[i] + 1
    pos[0] = p0
    if i + 1 >= cnt:
        return 1, pos
    p1 = nl[i + 1]
    pos[1] = p1
    p2 = p1 + 1
    pos[2] = p2
    j = i + 2
    while j < cnt and not (nl[j] >= p2 + 1 and blob[nl[j] + 1] == 0x2b):
        j += 1
    if j >= cnt:
        return 3, pos
    p3 = nl[j]
    pos[3] = p3
    if p3 + 2 >= L:
        return 7, pos
    if j + 1 >= cnt:
        return 7, pos
    h = nl[j + 1]
    if (h - p3 - 1) > 1 and (h - p3) != (p1 - p0 + 1):
        return -1, pos
    p4 = h + 1
    pos[4] = p4
    p5 = p4 + p3 - p1 - 1
    if p5 + 2 >= L:
        return 5, pos
    pos[5] = p5
    return 6, pos


def model_fast4(data, sentinel, goff, tile, nlcap=None):
    """The fast path.  Returns None when validation fails (general path needed), else
    (rows, tail_status, tail_pos, resume_offset)."""
    blob, NL, CL = visible_newlines(data, sentinel)
    L = len(blob)
    M = len(NL)
    # tiles are cut over blob coordinates here (the kernel cuts over aligned addresses; any cut works)
    n_tiles = max(1, -(-L // tile))
    table = {}
    fail = False
    tile_first_rank = []
    r = 0
    for t in range(n_tiles):
        lo, hi = t * tile, (t + 1) * tile
        idx = [i for i in range(M) if lo <= NL[i] < hi]
        B = r
        tile_first_rank.append(B)
        n = len(idx)
        r += n
        if nlcap is not None and n > nlcap:
            fail = True
            continue
        s = [NL[i] for i in idx]
        c = [CL[i] for i in idx]
        j0 = (4 - (B & 3)) & 3
        F = (n - 1 - j0) >> 2 if n > j0 else 0
        for q in range(F):
            j = j0 + 4 * q
            k = (B + j) >> 2
            s0, s1, s2, s3, s4 = s[j:j + 5]
            ok = c[j] == CLS_AT and c[j + 1] != CLS_NL and c[j + 2] == CLS_PLUS
            plus_len = s3 - s2
            if plus_len > 2 and plus_len != s1 - s0:
                ok = False
            if s4 - s3 != s2 - s1:
                ok = False
            table[k] = [s0 + 1, s1, s1 + 1, s2, s3 + 1, s3 + s2 - s1]
            if not ok:
                fail = True
        unc = list(range(0, min(j0, n))) + (list(range(j0 + 4 * F, n)) if n > j0 else [])
        assert len(unc) <= 7
        for j in unc:
            rr = B + j
            k, f = rr >> 2, rr & 3
            row = table.setdefault(k, [None] * 6)
            if f == 0:
                row[0] = s[j] + 1
            elif f == 1:
                row[1] = s[j]
                row[2] = s[j] + 1
            elif f == 2:
                row[3] = s[j]
            else:
                row[4] = s[j] + 1
    tile_first_rank.append(r)
    assert r == M
    # seam fix-up
    for t in range(1, n_tiles):
        Bt, Et = tile_first_rank[t], tile_first_rank[t + 1]
        if Bt >= 1 and Et > Bt:
            k = (Bt - 1) >> 2
            if 4 * k + 4 <= M - 1:
                row = table[k]
                p0, p1, p3, p4 = row[0], row[1], row[3], row[4]
                d = table[k + 1][0] - 1
                p5 = p4 + p3 - p1 - 1
                row[5] = p5
                ok = blob[p0] == 0x40 and blob[p1 + 1] != 0x0a and blob[p3 + 1] == 0x2b
                plus_len = (p4 - 1) - p3
                if plus_len > 2 and plus_len != p1 - p0 + 1:
                    ok = False
                if d != p5:
                    ok = False
                if not ok:
                    fail = True
    if fail:
        return None
    if M == 0:
        return [], 0, [-1] * 6, 0
    K = (M - 1) >> 2
    m = (M - 1) & 3
    last_is_5 = K >= 1 and m == 0 and blob[L - 2] == 0x0a
    n = K - (1 if last_is_5 else 0)
    if last_is_5:
        pos = table[K - 1][:5] + [-1]
        status = 5
    else:
        row = table[K]
        nl = [row[0] - 1]
        if m >= 1:
            nl.append(row[1])
        if m >= 2:
            nl.append(row[3])
        if m >= 3:
            nl.append(row[4] - 1)
        status, pos = classify_tail(blob, nl)
        if status == 6:
            row[5] = pos[5]
            n = K + 1
            status, pos = 0, [-1] * 6
    resume = table[n - 1][5] - 1 if n >= 1 else 0
    rows = [[x + goff for x in table[k]] for k in range(n)]
    return rows, status, pos, resume


def compute_rec(NL, CL, nxp, nxa, L, i):
    """Record anchored at candidate node i (CL[i] == '@'): (status, pos[6], successor node)."""
    M = len(NL)
    pos = [-1] * 6
    p0 = NL[i] + 1
    pos[0] = p0
    if i + 1 >= M:
        return 1, pos, NONE
    p1 = NL[i + 1]
    pos[1] = p1
    p2 = p1 + 1
    pos[2] = p2
    kmin = i + 2 + (1 if CL[i + 1] == CLS_NL else 0)
    k = nxp[kmin] if kmin < M else NONE
    if k is NONE:
        return 3, pos, NONE
    p3 = NL[k]
    pos[3] = p3
    if p3 + 2 >= L:
        return 7, pos, NONE
    if k + 1 >= M:
        return 7, pos, NONE
    h = NL[k + 1]
    if (h - p3 - 1) > 1 and (h - p3) != (p1 - p0 + 1):
        return -1, pos, NONE
    p4 = h + 1
    pos[4] = p4
    p5 = p4 + p3 - p1 - 1
    if p5 + 2 >= L:
        return 5, pos, NONE
    pos[5] = p5
    target = p5 - 1
    lb = k + 2
    if lb < M and NL[lb] < target:
        lo, step = lb, 1  # NL[lo] < target
        while lo + step < M and NL[lo + step] < target:
            lo += step
            step <<= 1
        hi = min(lo + step, M)  # NL[hi] >= target or hi == M
        while hi - lo > 1:
            mid = (lo + hi) >> 1
            if NL[mid] < target:
                lo = mid
            else:
                hi = mid
        lb = hi
    succ = nxa[lb] if lb < M else NONE
    return 6, pos, succ


def model_general(data, sentinel, goff, chunk=8):
    """The general path: line table, next-'+'/'@' arrays, per-candidate successor, chunked jump
    (exit / hops), chunk-level walk, per-chunk emission.  Returns (rows, status, pos, resume)."""
    blob, NL, CL = visible_newlines(data, sentinel)
    L = len(blob)
    M = len(NL)
    nxp = [NONE] * (M + 1)
    nxa = [NONE] * (M + 1)
    for i in range(M - 1, -1, -1):
        nxp[i] = i if CL[i] == CLS_PLUS else nxp[i + 1]
        nxa[i] = i if CL[i] == CLS_AT else nxa[i + 1]
    if M == 0 or nxa[0] is NONE:
        return [], 0, [-1] * 6, 0
    recs = {}
    for i in range(M):
        if CL[i] == CLS_AT:
            recs[i] = compute_rec(NL, CL, nxp, nxa, L, i)
    # chunked pointer jumping: exit[u] / hops[u] for every candidate
    n_chunks = -(-M // chunk)
    exit_, hops = {}, {}
    for g in range(n_chunks):
        lo, hi = g * chunk, min((g + 1) * chunk, M)
        for u in range(hi - 1, lo - 1, -1):  # (the kernel does this with log-step pointer jumping)
            if u not in recs:
                continue
            st, _, su = recs[u]
            if st != 6:
                exit_[u], hops[u] = NONE, 0
            elif su is NONE or su >= hi:
                exit_[u], hops[u] = su, 1
            else:
                exit_[u], hops[u] = exit_[su], 1 + hops[su]
    # level 2: walk over chunk entries
    entry = [NONE] * n_chunks
    cnt = [0] * n_chunks
    cur = nxa[0]
    while cur is not NONE:
        g = cur // chunk
        entry[g] = cur
        cnt[g] = hops[cur]
        cur = exit_[cur]
    base = [0] * n_chunks
    acc = 0
    for g in range(n_chunks):
        base[g] = acc
        acc += cnt[g]
    n = acc
    # emission per chunk
    rows = [None] * n
    terminal = None
    for g in range(n_chunks):
        if entry[g] is NONE:
            continue
        hi = min((g + 1) * chunk, M)
        u = entry[g]
        r = base[g]
        while True:
            st, pos, su = recs[u]
            if st != 6:
                terminal = (st, pos)
                break
            rows[r] = [x + goff for x in pos]
            r += 1
            if su is NONE:
                terminal = (0, [-1] * 6)
                break
            if su >= hi:
                break
            u = su
        assert r == base[g] + cnt[g]
    assert terminal is not None and all(x is not None for x in rows)
    resume = rows[n - 1][5] - goff - 1 if n >= 1 else 0
    return rows, terminal[0], terminal[1], resume


# ---- 2-bit packing: word-level model of csrc/fq_consume.cuh (pack2_step / pack2_span) -------------------------
M32 = 0xffffffff


def base_codes4(w):
    return ((w >> 1) & 0x03030303) ^ ((w >> 2) & 0x01010101)


def squeeze_codes4(c):
    return ((c * 0x01041040) & M32) >> 24


def eq_flags4(x, k):
    k7 = 0x7f7f7f7f
    y = x ^ k
    return ~((((y & k7) + k7) & M32) | y) & 0x80808080


def acgtu_flags4(w):
    x = w & 0xdfdfdfdf
    f = 0
    for k in (0x41, 0x43, 0x47, 0x54, 0x55):
        f |= eq_flags4(x, k * 0x01010101)
    return f


def flags_to_mask4(f):
    return ((((f >> 7) * 0x00204081) & M32) >> 21) & 0xf


def pack2_model(data, b, e):
    """Words (list of 32-bit ints), bases, other for the field data[b:e], the way one lane of fq_pack2_kernel does it."""
    acc, nbits, out, bases, other = 0, 0, [], 0, 0
    a = b
    while a < e:
        take = min(16, e - a)
        x16 = bytes(data[a:a + take]) + b'A' * (16 - take)
        nl = ok = codes = 0
        for q in range(4):
            x = int.from_bytes(x16[4 * q:4 * q + 4], 'little')
            nl |= flags_to_mask4(eq_flags4(x, 0x0a0a0a0a)) << (4 * q)
            ok |= flags_to_mask4(acgtu_flags4(x)) << (4 * q)
            codes |= squeeze_codes4(base_codes4(x)) << (8 * q)
        in_field = 0xffff if take >= 16 else (1 << take) - 1
        valid = in_field & ~nl
        cnt = bin(valid).count('1')
        other += cnt - bin(ok & valid).count('1')
        cw = codes
        if valid != in_field:
            cw, k = 0, 0
            for j in range(16):
                if (valid >> j) & 1:
                    cw |= ((codes >> (2 * j)) & 3) << (2 * k)
                    k += 1
        elif take < 16:
            cw &= (1 << (2 * take)) - 1
        bases += cnt
        acc |= cw << nbits
        nbits += 2 * cnt
        if nbits >= 32:
            out.append(acc & M32)
            acc >>= 32
            nbits -= 32
        a += 16
    n_words = (e - b + 15) // 16
    if nbits > 0 and len(out) < n_words:
        out.append(acc & M32)
    out += [0] * (n_words - len(out))
    return out, bases, other


# ---- byte-range sharding: the ownership / line-base rule of the sharded fast path (csrc/fq_emit.cuh) ---------
def model_shard_fast4(buf, own_len, sentinel, is_last, stream_offset, line_base):
    """Rows (absolute stream offsets) a shard emits from its buffer = own bytes + halo, given the number of lines the
    earlier shards own.  Newline i of the buffer has global rank line_base + i; rank 4k opens record k; a record
    belongs to the shard whose OWN range holds that newline (the virtual sentinel for record 0).  Returns
    (first global record index, rows, error) with error in (None, 'halo', 'general')."""
    blob, NL, CL = visible_newlines(buf, sentinel)
    L, M = len(blob), len(NL)
    goff = stream_offset - sentinel
    k0 = (line_base + 3) >> 2
    rows = []
    for i in range(M):
        if (line_base + i) & 3:
            continue
        if NL[i] - sentinel >= own_len:  # opened in the halo: the next shard's record
            break
        closed = i + 4 <= M - 1
        if not closed:
            if not is_last:
                return k0, rows, 'halo'
            break
        s0, s1, s2, s3, s4 = NL[i:i + 5]
        ok = CL[i] == CLS_AT and CL[i + 1] != CLS_NL and CL[i + 2] == CLS_PLUS
        plus_len = s3 - s2
        if plus_len > 2 and plus_len != s1 - s0:
            ok = False
        if s4 - s3 != s2 - s1:
            ok = False
        if not ok:
            return k0, rows, 'general'
        assert len(rows) == ((line_base + i) >> 2) - k0
        rows.append([s0 + 1 + goff, s1 + goff, s1 + 1 + goff, s2 + goff, s3 + 1 + goff, s3 + s2 - s1 + goff])
    if is_last and rows and M >= 1:
        # the last closed record is COMPLETE only if pos5 + 2 < L (src/_fastqandfurious.c:130)
        Mg = line_base + M
        if ((Mg - 1) & 3) == 0 and blob[L - 2] == 0x0a and len(rows) == ((Mg - 1) >> 2) - k0:
            rows.pop()
    return k0, rows, None


# ---- FASTA: the chain of entrypos_fasta calls from newline ranks (csrc/fq_fasta.cuh) --------------------------
def model_fasta(data, sentinel, goff, tile=64, group=4):
    """Rows / status / positions / resume offset the way the FASTA kernels compute them: on-chain flags per newline
    rank from the run parity, the parity from the last non-candidate rank (per tile, running maximum in two levels,
    fix-up of every tile's leading run), exclusive prefix sum, rows written by the on-chain ranks."""
    blob = (b'\n' if sentinel else b'') + bytes(data)
    L = len(blob)
    P = [i for i in range(L) if blob[i] == 0x0a]  # FASTA: a newline in the last byte is visible (bytes.find)
    M = len(P)
    cand = [P[r] + 1 < L and blob[P[r] + 1] == 0x3e for r in range(M)]
    n_tiles = max(1, -(-L // tile))
    tiles = [[] for _ in range(n_tiles)]
    for r in range(M):
        tiles[P[r] // tile].append(r)
    flags = [0] * M
    tilemax, lead = [-1] * n_tiles, [0] * n_tiles
    for t, ranks in enumerate(tiles):  # fq_fa_flags_kernel
        if not ranks:
            continue
        carry = ranks[0] - 1  # assumed: the rank before the tile is not a candidate
        in_lead = True
        for r in ranks:
            before = r - 1 - carry
            flags[r] = 1 if cand[r] and not (before & 1) else 0
            if cand[r]:
                if in_lead:
                    lead[t] += 1
            else:
                in_lead = False
                carry = r
                tilemax[t] = r
    n_groups = -(-n_tiles // group)
    groupmax = [-1] * n_groups
    for g in range(n_groups):  # fq_fa_groupscan_kernel: in place inside the group
        run = -1
        for t in range(g * group, min(n_tiles, (g + 1) * group)):
            run = max(run, tilemax[t])
            tilemax[t] = run
        groupmax[g] = run
    for g in range(1, n_groups):  # fq_fa_topscan_kernel
        groupmax[g] = max(groupmax[g], groupmax[g - 1])

    def fa_carry(t):
        if t == 0:
            return -1
        c = tilemax[t - 1] if t % group else -1
        if t // group > 0:
            c = max(c, groupmax[t // group - 1])
        return c
    for t, ranks in enumerate(tiles):  # fq_fa_fixup_kernel
        if t == 0 or not lead[t]:
            continue
        B = ranks[0]
        if (B - 1 - fa_carry(t)) & 1:
            for r in ranks[:lead[t]]:
                flags[r] ^= 1
    chain = [r for r in range(M) if flags[r]]
    total = len(chain)
    if total == 0:
        return [], 0, [-1] * 4, 0
    rows = [[None] * 4 for _ in range(total - 1)]
    status, pos, resume = None, None, 0
    for k, r in enumerate(chain):  # fq_fa_rows_kernel
        p = P[r]
        p1 = P[r + 1] if r + 1 < M else -1
        if k >= 1:
            rows[k - 1][3] = p + goff
        if k + 1 < total:
            rows[k][:3] = [p + 1 + goff, p1 + goff, p1 + 1 + goff]
        else:
            pos = [p + 1, -1, -1, -1]
            if p1 < 0:
                status = 1
            else:
                pos[1] = p1
                if p1 + 1 >= L:
                    status = 2
                else:
                    pos[2] = p1 + 1
                    pos[3] = L - 1 if blob[L - 1] == 0x0a else L
                    status = 3
            resume = p if total - 1 >= 1 else 0
    return rows, status, pos, resume


# ---- byte-range sharding of the GENERAL path: the hand-over of the chain from shard to shard ------------------
def model_shard_general(data, cuts, halo, entrypos):
    """fqb_shard_general's protocol with a per-call `entrypos(blob, offset, pos) -> status` (the oracle's) standing in
    for the device's candidate forest: every shard walks the chain over own bytes + halo from the position its
    predecessor hands over (pos5 - 1 of that shard's last owned record), emits the records whose leading newline
    lies in its own range and hands the position on.  Returns (rows in global blob coordinates, status of the call
    the chain ended on or None, error in (None, 'halo'))."""
    data = bytes(data)
    n = len(data)
    bounds = [0] + list(cuts) + [n]
    world = len(bounds) - 1
    rows = []
    resume = 0  # global blob coordinate ('\n' + data) the next search starts at
    for g in range(world):
        lo, hi = bounds[g], bounds[g + 1]
        last = g == world - 1
        hend = n if last else min(n, hi + halo)
        sentinel = 1 if g == 0 else 0
        blob = (b'\n' if sentinel else b'') + data[lo:hend]
        shift = 0 if g == 0 else lo + 1  # global blob coordinate of local blob index 0
        while True:
            pos = [-1] * 6
            st = entrypos(blob, max(0, resume - shift), pos)
            if pos[0] >= 0 and (pos[0] - 1) + shift - 1 >= hi:
                break  # the next record opens behind this shard's own bytes: the next shard's
            if st == 6:
                rows.append([p + shift for p in pos])
                resume = pos[5] - 1 + shift
                continue
            if last:
                return rows, st, None
            if st == 0:
                break  # nothing more in own bytes + halo: the chain continues (if at all) further right
            if st == -1:
                return rows, st, None  # INVALID on an owned record: the chain ends here
            return rows, st, 'halo'  # an owned record does not close inside own bytes + halo
    return rows, None, None


# ---- speculative general path (csrc/fq_gspec.cuh): chunks of tiles resolved independently, verified by continuity --------
S_UNRES, S_NONE_T, S_NONE_E = 'U', 'T', 'E'


def spec_rec(NL, CL, L, R0, nw, at_end, i, scan_max):
    """One entrypos call anchored on window line i (class '@'), answered from the window's lines only
    (global ranks R0 .. R0 + nw - 1).  Returns (status, pos[6], succ) with succ a window index, S_NONE_E (COMPLETE, no
    further '\\n@'), S_NONE_T (not COMPLETE: the chain stops ON this node) or S_UNRES (the window cannot tell)."""
    P = lambda j: NL[R0 + j]  # noqa: E731
    C = lambda j: CL[R0 + j]  # noqa: E731
    pos = [-1] * 6
    p0 = P(i) + 1
    pos[0] = p0
    if i + 1 >= nw:
        return (1, pos, S_NONE_T) if at_end else (None, pos, S_UNRES)
    p1 = P(i + 1)
    pos[1] = p1
    p2 = p1 + 1
    pos[2] = p2
    k = i + 2 + (1 if C(i + 1) == CLS_NL else 0)
    steps = 0
    while k < nw and C(k) != CLS_PLUS:
        k += 1
        steps += 1
        if steps > scan_max:
            return None, pos, S_UNRES
    if k >= nw:
        return (3, pos, S_NONE_T) if at_end else (None, pos, S_UNRES)
    p3 = P(k)
    pos[3] = p3
    if p3 + 2 >= L:
        return 7, pos, S_NONE_T
    if k + 1 >= nw:
        return (7, pos, S_NONE_T) if at_end else (None, pos, S_UNRES)
    h = P(k + 1)
    if (h - p3 - 1) > 1 and (h - p3) != (p1 - p0 + 1):
        return -1, pos, S_NONE_T
    p4 = h + 1
    pos[4] = p4
    p5 = p4 + p3 - p1 - 1
    if p5 + 2 >= L:
        return 5, pos, S_NONE_T
    pos[5] = p5
    target = p5 - 1
    j = k + 2
    steps = 0
    while j < nw and P(j) < target:
        j += 1
        steps += 1
        if steps > scan_max:
            return None, pos, S_UNRES
    while j < nw and C(j) != CLS_AT:
        j += 1
        steps += 1
        if steps > scan_max:
            return None, pos, S_UNRES
    if j >= nw:
        return (6, pos, S_NONE_E) if at_end else (None, pos, S_UNRES)
    return 6, pos, j


def model_general_spec(data, sentinel, goff, tile=64, tc=2, wmax=1 << 30, lookback=160, scan_max=1 << 30, starts=8,
                       walkers=32, runup=16, cmax=1 << 30):
    """The speculative general path (csrc/fq_gspec.cuh).  Returns None when it declines (the exact general path
    takes over), else (rows, status, pos, resume) -- which must then equal the reference's chain.
    Per chunk: the candidates ('@'-class lines of the last `lookback` look-behind lines and of the own lines) make
    their calls; `walkers` walkers split the own candidates into consecutive regions, each follows the successors
    from `runup` candidates before its region to its entry, then through its region; a walker's entry must be the
    exit of the walker before it (walker 0's entry is the speculated entry of the chunk)."""
    import bisect
    blob, NL, CL = visible_newlines(data, sentinel)
    L, M = len(blob), len(NL)
    n_tiles = -(-L // tile)
    if n_tiles == 0 or M == 0:
        return None
    tile_of = [p // tile for p in NL]
    first = [bisect.bisect_left(tile_of, t) for t in range(n_tiles + 1)]
    n_chunks = -(-n_tiles // tc)
    pe, xx, rows_of = [None] * n_chunks, [None] * n_chunks, [None] * n_chunks
    tail = None
    END_E, END_T, UNRES = 'end_e', 'end_t', 'unres'
    for c in range(n_chunks):
        t0, t1 = c * tc, min((c + 1) * tc, n_tiles)
        tb = t0 - 1 if c > 0 else t0
        te = min(t1 + 1, n_tiles)
        R0, nb, no, nw = first[tb], first[t0] - first[tb], first[t1] - first[t0], first[te] - first[tb]
        nbo = nb + no
        at_end = te == n_tiles
        if nw > wmax:
            return None
        clo = max(0, nb - lookback)
        cand = [i for i in range(clo, nbo) if CL[R0 + i] == CLS_AT]
        nc = len(cand)
        if nc > cmax:
            return None
        q_of = {i: q for q, i in enumerate(cand)}
        nl, recs = [], []
        for i in cand:
            st, pos, s = spec_rec(NL, CL, L, R0, nw, at_end, i, scan_max)
            nl.append(s)
            recs.append((st, pos))
        nq = [q_of[s] if isinstance(s, int) and s < nbo else None for s in nl]  # successors inside [clo, nbo) are candidates

        def run(q, limit):
            """(node, candidate index or None): first chain node with candidate index >= limit, or where the chain
            leaves the candidates / how it ended."""
            while True:
                if q >= limit:
                    return cand[q], q
                if nq[q] is None:
                    s = nl[q]
                    if isinstance(s, int):
                        return s, None
                    return ((END_E if s == S_NONE_E else END_T if s == S_NONE_T else UNRES), cand[q]), None
                q = nq[q]

        q_own = 0 if c == 0 else sum(1 for i in cand if i < nb)
        cper = -(-(nc - q_own) // walkers)
        ents, exits, out, term = [], [], [], None
        for wk in range(walkers):
            rlo = min(q_own + wk * cper, nc)
            rhi = min(rlo + cper, nc)
            if not (wk == 0 or rlo < nc):
                break
            a, qa_in = None, None
            if c == 0 and wk == 0:
                if nc > 0:
                    a, qa_in = cand[0], 0
                else:
                    ahead = [i for i in range(nbo, nw) if CL[R0 + i] == CLS_AT]  # the head may lie in the look-ahead tile
                    a = ahead[0] if ahead else (END_E, None)
            else:
                for s0 in range(max(0, rlo - runup), rlo):
                    if not (isinstance(nl[s0], int) or nl[s0] == S_NONE_E):
                        continue
                    # prefer a start that one of the three candidates before it points to (the last one is always tried)
                    if s0 >= 3 and s0 + 1 < rlo and s0 not in (nq[s0 - 1], nq[s0 - 2], nq[s0 - 3]):
                        continue
                    a, qo = run(s0, rlo)
                    if isinstance(a, int):
                        qa_in = qo
                        break
            if a is None or (isinstance(a, tuple) and a[0] == UNRES):
                return None
            xw = a
            if qa_in is not None and qa_in < rhi:
                q = qa_in
                while True:
                    s = nl[q]
                    if s == S_UNRES:
                        return None
                    if s == S_NONE_T:
                        xw = (END_T, cand[q])
                        term = q
                        break
                    out.append([v + goff for v in recs[q][1]])
                    if s == S_NONE_E:
                        xw = (END_E, cand[q])
                        break
                    if nq[q] is None:
                        xw = s
                        break
                    if nq[q] >= rhi:
                        xw = cand[nq[q]]
                        break
                    q = nq[q]
            if ents and a != exits[-1]:
                return None  # a walker's entry is not the exit of the walker before it
            ents.append(a)
            exits.append(xw)
        en, ex = ents[0], exits[-1]
        if isinstance(en, int):
            pe[c] = R0 + en
        elif c == 0 and en == (END_E, None) and at_end:
            return [], 0, [-1] * 6, 0  # no "\n@" at all
        else:
            return None
        if isinstance(ex, int):
            xx[c] = R0 + ex
        elif ex[0] == END_E:
            xx[c] = S_NONE_E
            tail = (c, 0, [-1] * 6)
        else:
            xx[c] = S_NONE_T
            tail = (c, recs[term][0], recs[term][1])
        rows_of[c] = out
    # verification: the speculated entries are the exits of the chunks before them; the chain ends in the last chunk
    for c in range(1, n_chunks):
        if xx[c - 1] != pe[c]:
            return None
    if xx[-1] not in (S_NONE_T, S_NONE_E) or tail is None or tail[0] != n_chunks - 1:
        return None
    rows = [r for part in rows_of for r in part]
    n = len(rows)
    resume = rows[n - 1][5] - goff - 1 if n >= 1 else 0
    return rows, tail[1], tail[2], resume


# ---- speculative general path, warp-autonomous form (csrc/fq_gspec2.cuh) ---------------------------------------------
def spec_rec2(NL, CL, L, R0, nw, at_end, i, scan_max):
    """spec_rec with the kernel's successor guess: with nseq = k - i - 1 sequence lines the quality is expected to end
    on line jg = 2k - i; when that line's newline sits exactly at pos5 and an '@' follows it, jg is the successor (no
    earlier line can qualify: the only one at or behind pos5 - 1 would be followed by the newline at pos5).  Otherwise
    the bounded linear search.  Also returns k (the line of the '+')."""
    P = lambda j: NL[R0 + j]  # noqa: E731
    C = lambda j: CL[R0 + j]  # noqa: E731
    pos = [-1] * 6
    p0 = P(i) + 1
    pos[0] = p0
    if i + 1 >= nw:
        return ((1, pos, S_NONE_T) if at_end else (None, pos, S_UNRES)) + (None,)
    p1 = P(i + 1)
    pos[1] = p1
    pos[2] = p1 + 1
    k = i + 2 + (1 if C(i + 1) == CLS_NL else 0)
    while k < nw and C(k) != CLS_PLUS:  # (the kernel: a bit mask of the '+'-class lines, no bound but the window)
        k += 1
    if k >= nw:
        return ((3, pos, S_NONE_T) if at_end else (None, pos, S_UNRES)) + (None,)
    p3 = P(k)
    pos[3] = p3
    if p3 + 2 >= L:
        return 7, pos, S_NONE_T, k
    if k + 1 >= nw:
        return ((7, pos, S_NONE_T) if at_end else (None, pos, S_UNRES)) + (k,)
    h = P(k + 1)
    if (h - p3 - 1) > 1 and (h - p3) != (p1 - p0 + 1):
        return -1, pos, S_NONE_T, k
    p4 = h + 1
    pos[4] = p4
    p5 = p4 + p3 - p1 - 1
    if p5 + 2 >= L:
        return 5, pos, S_NONE_T, k
    pos[5] = p5
    jg = 2 * k - i
    if jg < nw and P(jg) == p5 and C(jg) == CLS_AT:
        return 6, pos, jg, k
    j = k + 2
    steps = 0
    while j < nw and P(j) < p5 - 1:
        j += 1
        steps += 1
        if steps > scan_max:
            return None, pos, S_UNRES, k
    while j < nw and C(j) != CLS_AT:  # (the kernel: a bit mask of the '@'-class lines)
        j += 1
    if j >= nw:
        return ((6, pos, S_NONE_E) if at_end else (None, pos, S_UNRES)) + (k,)
    return 6, pos, j, k


def model_general_spec2(data, sentinel, goff, tile=64, tc=2, wl=1 << 30, lbl=160, lal=128, cw=1 << 30, runup=16,
                        maxg=1 << 30, scan_max=1 << 30, gw=32, park=1 << 30, big=1 << 30):
    """The speculative general path as one WARP per chunk (csrc/fq_gspec2.cuh).  None = declined.
    Window of chunk c: the last `lbl` lines of the tile before its own tiles, the own tiles, the first `lal` lines of
    the tile behind them.  All '@'-class lines below the look-ahead are candidates and make their call.  The chain is
    followed in groups of `gw` consecutive candidates (the kernel: pointer doubling over the lanes of the warp); the
    entry of chunk c > 0 is speculated from the last `runup` look-behind candidates: the first of them whose call is
    COMPLETE, that one of the three candidates before it points to (the first three and the last one are exempt),
    and whose chain reaches a candidate of the own lines or a look-ahead line.  Verified by continuity.
    The rows of a chunk are parked in the upper half of its own tiles' list slots until the records before it are
    counted: `park` rows per own tile, no own tile with more than `big` lines.  (`scan_max` bounds the model's linear
    successor search; the kernel searches the sorted lines and needs no bound.)"""
    import bisect
    blob, NL, CL = visible_newlines(data, sentinel)
    L, M = len(blob), len(NL)
    n_tiles = -(-L // tile)
    if n_tiles == 0 or M == 0:
        return None
    tile_of = [p // tile for p in NL]
    first = [bisect.bisect_left(tile_of, t) for t in range(n_tiles + 1)]
    n_chunks = -(-n_tiles // tc)
    pe, xx, rows_of = [None] * n_chunks, [None] * n_chunks, [None] * n_chunks
    tail = None
    for c in range(n_chunks):
        t0, t1 = c * tc, min((c + 1) * tc, n_tiles)
        tb = t0 - 1 if c > 0 else t0
        te = min(t1 + 1, n_tiles)
        skip = max(0, (first[t0] - first[tb]) - lbl) if c > 0 else 0
        R0 = first[tb] + skip
        nb = first[t0] - R0
        nbo = first[t1] - R0
        la_full = first[te] - first[t1]
        la = min(la_full, lal)
        nw = nbo + la
        at_end = te == n_tiles and la == la_full
        if nw > wl or any(first[t + 1] - first[t] > big for t in range(t0, t1)):
            return None
        cand = [i for i in range(nbo) if CL[R0 + i] == CLS_AT]
        nc = len(cand)
        if nc > cw:
            return None
        q_of = {i: q for q, i in enumerate(cand)}
        calls = [spec_rec2(NL, CL, L, R0, nw, at_end, i, scan_max) for i in cand]
        # successor codes: candidate index / ('line', i) behind the candidates / 'E' / 'T' / 'U'
        succ = []
        for st, pos, s, k in calls:
            succ.append(q_of[s] if isinstance(s, int) and s < nbo else (('line', s) if isinstance(s, int) else s))
        q_own = sum(1 for i in cand if i < nb)

        def group(gb):
            """chain from candidate gb through candidates [gb, gb + gw): (nodes, first value outside)"""
            nodes, q = [], gb
            while isinstance(q, int) and gb <= q < gb + gw:
                nodes.append(q)
                q = succ[q]
            return nodes, q

        cur = None
        if c == 0:
            if nc > 0:
                cur = 0
            else:
                ahead = [i for i in range(nbo, nw) if CL[R0 + i] == CLS_AT]
                if ahead:
                    cur = ('line', ahead[0])
                elif at_end:
                    return [], 0, [-1] * 6, 0  # no "\n@" at all
                else:
                    return None
        else:
            g0 = max(0, q_own - runup)
            for s0 in range(g0, q_own):
                if succ[s0] in (S_NONE_T, S_UNRES):
                    continue
                if s0 >= 3 and s0 + 1 < q_own and s0 not in (succ[s0 - 1], succ[s0 - 2], succ[s0 - 3]):
                    continue
                # first node of the chain from s0 at or behind q_own, inside the group [g0, g0 + gw) or the value that leaves it
                nodes, out = _chain_in(succ, s0, g0, gw)
                own = [q for q in nodes if q >= q_own]
                e = own[0] if own else out
                if isinstance(e, int) or (isinstance(e, tuple) and e[0] == 'line'):
                    cur = e
                    break
            if cur is None:
                return None
        pe[c] = R0 + (cand[cur] if isinstance(cur, int) else cur[1])
        out_rows, term, ng = [], None, 0
        while isinstance(cur, int):
            nodes, nxt = group(cur)
            if any(succ[q] == S_UNRES for q in nodes):
                return None
            ng += 1
            if ng > maxg:
                return None
            for q in nodes:
                if succ[q] == S_NONE_T:
                    term = q
                else:
                    out_rows.append([v + goff for v in calls[q][1]])
            if len(out_rows) > park * (t1 - t0):
                return None
            cur = nxt
        if cur == S_UNRES:
            return None
        if isinstance(cur, tuple):
            xx[c] = R0 + cur[1]
        elif cur == S_NONE_E:
            xx[c] = S_NONE_E
            tail = (c, 0, [-1] * 6)
        else:
            xx[c] = S_NONE_T
            tail = (c, calls[term][0], calls[term][1])
        rows_of[c] = out_rows
    for c in range(1, n_chunks):
        if xx[c - 1] != pe[c]:
            return None
    if xx[-1] not in (S_NONE_T, S_NONE_E) or tail is None or tail[0] != n_chunks - 1:
        return None
    rows = [r for part in rows_of for r in part]
    n = len(rows)
    resume = rows[n - 1][5] - goff - 1 if n >= 1 else 0
    return rows, tail[1], tail[2], resume


def _chain_in(succ, q, gb, gw):
    nodes = []
    while isinstance(q, int) and gb <= q < gb + gw:
        nodes.append(q)
        q = succ[q]
    return nodes, q


# ---- two-level decoupled look-back with deferred flush (csrc/fq_gspec2.cuh) -------------------------------------------
def simulate_lookback(counts, n_warps, rng, block=32):
    """Discrete-event model of the record-count protocol of fq_gspec2_kernel under a random schedule: warps take chunk
    tickets in order; a warp PUBLISHES its chunk's count (desc[c], state 1), the last arrival of a block of `block`
    chunks publishes the block's aggregate (bdesc[b], state 1); a warp FLUSHES the chunk before its current one only
    after publishing the current one: exclusive prefix = counts of the chunks of its block before it (all must be
    published) + blocks before it, walked back `block` at a time until one with an inclusive prefix (state 2; every
    entry read must be published); the last chunk of a block then publishes the block's inclusive prefix.  Publishing
    never waits, flushing waits only for publications: returns the bases the chunks computed, or None on a deadlock."""
    n = len(counts)
    desc = [None] * n                      # published counts
    nb = -(-n // block)
    bdesc = [None] * nb                    # (state, value)
    bcnt = [0] * nb
    ticket = 0
    base_of = [None] * n
    warps = [{'cur': None, 'pending': None, 'state': 'take'} for _ in range(n_warps)]

    def try_flush(c):
        b, b_first = c // block, (c // block) * block
        if any(desc[j] is None for j in range(b_first, c)):
            return None
        base = sum(desc[b_first:c])
        bj0 = b - 1
        while bj0 >= 0:
            window = list(range(bj0, max(-1, bj0 - block), -1))
            if any(bdesc[bj] is None for bj in window):
                return None
            stop = next((k for k, bj in enumerate(window) if bdesc[bj][0] == 2), None)
            for k, bj in enumerate(window):
                if stop is None or k <= stop:
                    base += bdesc[bj][1]
            if stop is not None:
                break
            bj0 -= block
        return base

    steps = 0
    while True:
        steps += 1
        if steps > 200 * (n + n_warps) + 1000:
            return None
        live = [w for w in warps if w['state'] != 'done']
        if not live:
            break
        w = rng.choice(live)
        if w['state'] == 'take':
            if ticket < n:
                w['cur'] = ticket
                ticket += 1
                w['state'] = 'publish'
            else:
                w['cur'] = None
                w['state'] = 'flush' if w['pending'] is not None else 'done'
        elif w['state'] == 'publish':
            c = w['cur']
            desc[c] = counts[c]
            b = c // block
            bcnt[b] += 1
            if bcnt[b] == min(block, n - b * block):
                agg = sum(desc[b * block:min(n, (b + 1) * block)])
                if bdesc[b] is None or bdesc[b][0] < 2:  # atomicMax: a prefix is never replaced by an aggregate
                    bdesc[b] = (1, agg)
            w['state'] = 'flush' if w['pending'] is not None else 'keep'
        elif w['state'] == 'flush':
            c = w['pending']
            base = try_flush(c)
            if base is None:
                continue  # still spinning
            base_of[c] = base
            if c % block == block - 1:
                bdesc[c // block] = (2, base + counts[c])
            w['pending'] = None
            w['state'] = 'keep' if w['cur'] is not None else 'done'
        elif w['state'] == 'keep':
            w['pending'] = w['cur']
            w['state'] = 'take'
    return base_of
