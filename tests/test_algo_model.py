"""The device algorithms, restated sequentially in tests/am.py, against the oracle."""
import random

import pytest

import algo_model as am
import fqgen


def _oracle_chain(oracle, data, sentinel, goff):
    blob = (b'\n' if sentinel else b'') + data
    table, st, tail, resume = oracle.parse_chain(blob, 0, goff)
    return table.tolist(), st, tail.tolist(), resume


@pytest.mark.parametrize('seed', range(8))
def test_general_model_matches_oracle(oracle, seed):
    for data in fqgen.corpus(1000 + seed, 400):
        for sentinel in (1, 0):
            want = _oracle_chain(oracle, data, sentinel, -1)
            for chunk in (1, 3, 8, 1024):
                got = am.model_general(data, sentinel, -1, chunk=chunk)
                assert (got[0], got[1], list(got[2]), got[3]) == want, (data, sentinel, chunk)


@pytest.mark.parametrize('seed', range(8))
def test_fast4_model_exact_or_declines(oracle, seed):
    accepted = 0
    for data in fqgen.corpus(2000 + seed, 400):
        for sentinel in (1, 0):
            want = _oracle_chain(oracle, data, sentinel, 7)
            for tile in (5, 16, 64, 4096):
                got = am.model_fast4(data, sentinel, 7, tile)
                if got is None:
                    continue
                accepted += 1
                assert (got[0], got[1], list(got[2]), got[3]) == want, (data, sentinel, tile)
    assert accepted > 100


def test_fast4_model_accepts_clean_four_line(oracle):
    rng = random.Random(5)
    for _ in range(200):
        data = fqgen.fastq_bytes(rng, rng.randint(1, 10), long_plus=0.3, trailing_newlines=rng.randint(0, 3),
                                 at_plus_bias=0.5)
        for tile in (7, 64, 4096):
            got = am.model_fast4(data, 1, -1, tile)
            assert got is not None, data
            assert (got[0], got[1], list(got[2]), got[3]) == _oracle_chain(oracle, data, 1, -1)


def test_scan_kernel_bit_tricks():
    """The byte-SIMD arithmetic of fq_scan.cuh, restated with Python ints: exact newline flags for every
    byte value next to every neighbour, and the two-multiply gather of 16 flags into a position mask."""
    M32 = 0xffffffff

    def newline_flags(w):
        k7 = 0x7f7f7f7f
        t = ((w & k7) ^ 0x0a0a0a0a)
        return ~(((t + k7) & M32) | w) & 0x80808080 & M32

    rng = random.Random(3)
    words = [int.from_bytes(bytes([a, b, 10, c]), 'little') for a in (0, 9, 10, 11, 0x7f, 0x80, 0x8a, 0xff)
             for b in (10, 0x8a, 0xff, 0) for c in (10, 11, 0x8a)]
    words += [rng.getrandbits(32) for _ in range(20000)]
    for w in words:
        want = sum(0x80 << (8 * j) for j in range(4) if (w >> (8 * j)) & 0xff == 10)
        assert newline_flags(w) == want, hex(w)

    # the row loop's candidate test (scan_rows): bit 7 of every byte of w - 0x0b0b0b0b that holds a newline is set,
    # whatever the bytes below it are (no false negatives); exhaustive over byte pairs, then random words
    def maybe_newline(w):
        return ((w - 0x0b0b0b0b) & M32) & 0x80808080

    for lo in range(256):
        for hi in range(256):
            for w in (lo | hi << 8 | 0x41 << 16 | 0x0a << 24, 0x0a | lo << 8 | hi << 16 | 0x0a << 24):
                t = maybe_newline(w)
                for j in range(4):
                    if (w >> (8 * j)) & 0xff == 10:
                        assert t & (0x80 << (8 * j)), hex(w)
    for w in words:
        assert newline_flags(w) & ~maybe_newline(w) == 0, hex(w)
    # FASTQ text (printable ASCII and newlines) raises no false candidate except a byte 0x0b after a newline
    for _ in range(20000):
        w = int.from_bytes(bytes(rng.choice(b'ACGTN@+!I~ 0:#\n') for _ in range(4)), 'little')
        assert maybe_newline(w) == newline_flags(w), hex(w)

    kg = 0x00204081

    def gather(f0, f1, f2, f3):
        a = ((((f0 >> 4) | f1) * kg) & M32)
        b = ((((f2 >> 4) | f3) * kg) & M32)
        return (a >> 24) | ((b >> 16) & 0xff00)

    for m in range(1 << 16):
        f = [sum(0x80 << (8 * j) for j in range(4) if (m >> (4 * i + j)) & 1) for i in range(4)]
        assert gather(*f) == m


def test_pack2_bit_tricks_and_model_match_oracle(oracle):
    """The word-level arithmetic of fq_pack2_kernel (tests/am.py mirrors it): byte codes, the multiply that
    squeezes four codes into a byte, exact byte-equality flags, flag -> position masks; and the step / accumulator
    logic against the oracle's definition of the packed layout on random sequences (wrapped, with other letters)."""
    import numpy as np
    for b, c in ((ord('A'), 0), (ord('C'), 1), (ord('G'), 2), (ord('T'), 3), (ord('U'), 3), (ord('a'), 0), (ord('c'), 1),
                 (ord('g'), 2), (ord('t'), 3), (ord('u'), 3)):
        assert am.base_codes4(b) & 3 == c
    for v in range(256):  # four 2-bit codes in byte lanes -> one byte
        c = (v & 3) | (((v >> 2) & 3) << 8) | (((v >> 4) & 3) << 16) | (((v >> 6) & 3) << 24)
        assert am.squeeze_codes4(c) == v
    for m in range(16):
        f = sum(0x80 << (8 * j) for j in range(4) if (m >> j) & 1)
        assert am.flags_to_mask4(f) == m
    rng = random.Random(9)
    for k in range(256):  # exact equality flags for every byte value in every lane, random neighbours
        for lane in range(4):
            w = rng.getrandbits(32) & ~(0xff << (8 * lane)) | (k << (8 * lane))
            for target in (0x0a, 0x41, 0x55):
                f = am.eq_flags4(w, target * 0x01010101)
                want = sum(0x80 << (8 * j) for j in range(4) if ((w >> (8 * j)) & 0xff) == target)
                assert f == want
    for trial in range(200):
        n = rng.choice([0, 1, 3, 15, 16, 17, 31, 32, 33, 150, 301])
        seq = bytes(rng.choice(b'ACGTacgtNUuRY.') if rng.random() < 0.1 else rng.choice(b'ACGT') for _ in range(n))
        wrap = rng.choice([0, 0, 7, 60])
        field = b'\n'.join(seq[i:i + wrap] for i in range(0, len(seq), wrap)) if wrap and seq else seq
        lead = bytes(rng.choice(b'xyz\n') for _ in range(rng.randrange(5)))
        data = lead + field + b'\n+\n'
        table = np.array([[0, 0, len(lead), len(lead) + len(field), 0, 0]], dtype=np.int64)
        packed, offsets, nb, no = oracle.pack_2bit(data, table)
        words, bases, other = am.pack2_model(data, len(lead), len(lead) + len(field))
        assert bases == nb[0] == len(seq) and other == no[0]
        assert b''.join(w.to_bytes(4, 'little') for w in words) == packed.tobytes()
        assert offsets[1] == 4 * ((len(field) + 15) // 16)
        for i, ch in enumerate(seq[:40]):  # the layout itself: base i in bits 2(i % 4).. of byte i // 4
            if ch in b'ACGTUacgtu':
                assert (packed[i // 4] >> (2 * (i % 4))) & 3 == 'ACGT'.index(chr(ch).upper().replace('U', 'T'))


@pytest.mark.parametrize('seed', range(6))
def test_fasta_model_matches_oracle(oracle, seed):
    """The FASTA formulation of csrc/fq_fasta.cuh (run parity from last non-candidate ranks, per-tile assumption +
    fix-up, two-level running maximum) against the oracle's chain of entrypos_fasta calls, for tiles and groups small
    enough that runs of header-only records cross several of both."""
    rng = random.Random(700 + seed)
    cases = [fqgen.fasta_bytes(rng) for _ in range(150)]
    cases += [b'>h\n' * rng.randint(1, 200) + b'>x\nACGT\n' + b'>\n' * rng.randint(0, 9) + b'>y\nAC\n>z\n' for _ in range(10)]
    cases += [b'AC\n' * rng.randint(0, 3) + b'>' + b'h' * rng.randint(0, 300) + b'\n' + b'>\n' * rng.randint(0, 70) + b'>q\nA\n'
              for _ in range(10)]
    for data in cases:
        for sentinel in (1, 0):
            goff = rng.choice([-1, 0, 1000])
            tile, group = rng.choice([(3, 2), (16, 4), (64, 4), (1 << 20, 256)])
            want, st, tail, resume = oracle.fasta_chain((b'\n' if sentinel else b'') + data, 0, goff)
            rows, gst, gpos, gres = am.model_fasta(data, sentinel, goff, tile, group)
            ctx = (data[:80], sentinel, tile, group)
            assert rows == want.tolist(), ctx
            assert (gst, gpos, gres) == (st, tail.tolist(), resume), ctx


@pytest.mark.parametrize('seed', range(6))
def test_shard_general_handover_model_matches_oracle(oracle, seed):
    """The sharded general path's protocol (ownership by the leading newline, hand-over of pos5 - 1, halo check) on CPU:
    with the oracle's entrypos inside every shard, the shards' rows concatenate to the single-buffer chain."""
    from array import array
    rng = random.Random(900 + seed)

    def entrypos(blob, offset, pos):
        buf = array('q', [-1] * 6)
        st = oracle.entrypos(blob, offset, buf)
        pos[:] = list(buf)
        return st
    n_ok = 0
    for trial in range(60):
        kind = trial % 3
        if kind == 0:
            data = fqgen.fastq_bytes(rng, rng.randint(20, 120), read_len=(20, 90), header_len=(3, 20), wrap=rng.choice([0, 9, 30]),
                                     long_plus=0.3, trailing_newlines=rng.randint(0, 2), at_plus_bias=0.3)
        elif kind == 1:
            data = fqgen.mutate(rng, fqgen.fastq_bytes(rng, rng.randint(20, 120), read_len=(20, 90), header_len=(3, 20),
                                                        trailing_newlines=1, at_plus_bias=0.3), rng.randint(1, 4))
        else:
            data = fqgen.soup(rng, rng.randint(0, 30)) + fqgen.fastq_bytes(rng, rng.randint(10, 60), wrap=rng.choice([0, 5]))
        if len(data) < 400 or data.endswith(b'\n@'):
            continue
        world = rng.choice([2, 3, 5])
        cuts = sorted(rng.sample(range(50, len(data) - 50), world - 1))
        halo = rng.choice([120, 400, 5000])
        want, st, tail, resume = oracle.parse_chain(b'\n' + data, 0, 0)
        rows, end_st, err = am.model_shard_general(data, cuts, halo, entrypos)
        if err == 'halo':
            continue
        n_ok += 1
        assert rows == want.tolist(), (data[:60], cuts, halo)
        assert end_st is None or end_st == st, (end_st, st)
    assert n_ok > 20


@pytest.mark.parametrize('seed', range(3))
def test_general_spec_model_exact_or_declines(oracle, seed):
    """The speculative general path (chunks of tiles resolved independently from a window of lines, entries
    speculated from a look-behind, verified by continuity) either declines or equals the reference's chain -- for
    every tile size, chunk size, look-behind and scan bound, on clean, wrapped, damaged and garbage inputs."""
    rng = random.Random(700 + seed)
    accepted = 0
    for data in fqgen.corpus(8100 + seed, 260):
        for sentinel in (1, 0):
            want = _oracle_chain(oracle, data, sentinel, -1)
            tile = rng.choice([16, 32, 64, 256])
            got = am.model_general_spec(data, sentinel, -1, tile=tile, tc=rng.choice([1, 2, 4]),
                                                wmax=rng.choice([24, 200, 1 << 30]), lookback=rng.choice([2, 8, 1 << 30]),
                                                scan_max=rng.choice([3, 16, 1 << 30]), walkers=rng.choice([1, 2, 3, 32]),
                                                runup=rng.choice([1, 2, 16]), cmax=rng.choice([6, 1 << 30]))
            if got is not None:
                accepted += 1
                assert (got[0], got[1], list(got[2]), got[3]) == want, (data, sentinel, tile)
    assert accepted > 100


def test_general_spec_model_accepts_clean_wrapped_records(oracle):
    """Config 5's shape (wrapped reads, long '+' lines, '@' / '+' at line starts) is resolved without the exact path."""
    rng = random.Random(5)
    ok = 0
    for trial in range(30):
        data = fqgen.fastq_bytes(rng, 400, read_len=(100, 300), header_len=(8, 30), wrap=60, long_plus=0.5,
                                 trailing_newlines=1, at_plus_bias=0.02)
        got = am.model_general_spec(data, 1, -1, tile=8192, tc=4, wmax=6144, scan_max=192)
        want = _oracle_chain(oracle, data, 1, -1)
        if got is not None:
            ok += 1
            assert (got[0], got[1], list(got[2]), got[3]) == want
    assert ok == 30


@pytest.mark.parametrize('seed', range(3))
def test_general_spec2_model_exact_or_declines(oracle, seed):
    """The warp-per-chunk form of the speculative pass (csrc/fq_gspec2.cuh: partial look-behind / look-ahead tiles,
    the successor guessed from the number of sequence lines, the chain followed in groups of consecutive candidates,
    the entry speculated from the last look-behind candidates) either declines or equals the reference's chain."""
    rng = random.Random(900 + seed)
    accepted = 0
    for data in fqgen.corpus(9100 + seed, 260):
        for sentinel in (1, 0):
            want = _oracle_chain(oracle, data, sentinel, -1)
            tile = rng.choice([16, 32, 64, 256])
            got = am.model_general_spec2(data, sentinel, -1, tile=tile, tc=rng.choice([1, 2, 4]),
                                         wl=rng.choice([24, 200, 1 << 30]), lbl=rng.choice([2, 8, 160]),
                                         lal=rng.choice([1, 4, 16, 128]), cw=rng.choice([6, 1 << 30]),
                                         runup=rng.choice([1, 2, 16]), maxg=rng.choice([2, 1 << 30]),
                                         scan_max=rng.choice([3, 16, 1 << 30]), gw=rng.choice([1, 2, 5, 32]),
                                         park=rng.choice([2, 1 << 30, 1 << 30]), big=rng.choice([12, 1 << 30, 1 << 30]))
            if got is not None:
                accepted += 1
                assert (got[0], got[1], list(got[2]), got[3]) == want, (data, sentinel, tile)
    assert accepted > 100


def test_general_spec2_model_accepts_clean_wrapped_records(oracle):
    """Config 5's shape is resolved by the warp-per-chunk form with the kernel's window sizes."""
    rng = random.Random(6)
    ok = 0
    for trial in range(30):
        data = fqgen.fastq_bytes(rng, 400, read_len=(100, 300), header_len=(8, 30), wrap=60, long_plus=0.5,
                                 trailing_newlines=1, at_plus_bias=0.02)
        got = am.model_general_spec2(data, 1, -1, tile=8192, tc=4, wl=2048, lbl=160, lal=128, cw=448, maxg=24,
                                     park=128, big=1024)
        want = _oracle_chain(oracle, data, 1, -1)
        if got is not None:
            ok += 1
            assert (got[0], got[1], list(got[2]), got[3]) == want
    assert ok == 30


def test_list_consumer_bit_tricks():
    """The word-level arithmetic of the kernels that read the newline lists eight entries per lane (fq_gspec2.cuh,
    fq_fasta.cuh), restated with Python ints and checked against the definitions: class flags of a 16-byte vector of
    list entries, on-chain flags of FASTA runs from the carry trick (exhaustive over all 8-bit candidate masks and
    both entry parities), placement of an 8-bit mask in the window's 32-bit words, pointer doubling over 32 lanes."""
    M32 = 0xffffffff
    rng = random.Random(17)

    # class flags: entries are 16 bits, class in the low two bits ('@' / '>' = 1, '+' = 2), two entries per word
    def class_bits(words):
        a = b = 0
        for q, x in enumerate(words):
            hi = x >> 1
            a |= ((x & ~hi & 0x00010001) << (2 * q)) & M32
            b |= ((hi & ~x & 0x00010001) << (2 * q)) & M32
        return (a | (a >> 15)) & 0xff, (b | (b >> 15)) & 0xff

    for _ in range(20000):
        ent = [rng.getrandbits(16) for _ in range(8)]
        words = [ent[2 * q] | ent[2 * q + 1] << 16 for q in range(4)]
        at8, pl8 = class_bits(words)
        assert at8 == sum(1 << k for k in range(8) if ent[k] & 3 == 1)
        assert pl8 == sum(1 << k for k in range(8) if ent[k] & 3 == 2)

    # FASTA: on(r) = cand(r) and an even number of consecutive candidates immediately before r (fa_flags8)
    def on8_kernel(cand8, cnt):
        lead_run = cand8 & ~(cand8 + 1)
        rest = cand8 & ~lead_run
        starts = rest & ~(rest << 1)
        runs_e = rest & ~(rest + (starts & 0x55))
        return (lead_run & (0xAA if cnt & 1 else 0x55)) | (runs_e & 0x55) | (rest & ~runs_e & 0xAA)

    for cand8 in range(256):
        for cnt in (0, 1, 2, 7):
            want, run = 0, cnt  # run = consecutive candidates immediately before the bit
            for k in range(8):
                if cand8 >> k & 1:
                    if run % 2 == 0:
                        want |= 1 << k
                    run += 1
                else:
                    run = 0
            assert on8_kernel(cand8, cnt) == want, (bin(cand8), cnt)

    # an 8-bit mask whose bit k belongs to window line dbase + k (dbase >= -7, bits of lines < 0 are zero) lands in the
    # words w0 / w0 + 1 (fq_gspec2.cuh, staging)
    for _ in range(5000):
        dbase = rng.randint(-7, 300)
        m8 = rng.getrandbits(8)
        if dbase < 0:
            m8 &= ~((1 << -dbase) - 1) & 0xff
        words = {}
        pos = dbase + 32
        sh, w0 = pos & 31, (pos >> 5) - 1
        lo = (m8 << sh) & M32
        if lo:
            words[w0] = words.get(w0, 0) | lo
        if sh > 24 and m8 >> (32 - sh):
            words[w0 + 1] = words.get(w0 + 1, 0) | (m8 >> (32 - sh))
        want = {}
        for k in range(8):
            if m8 >> k & 1:
                line = dbase + k
                want[line >> 5] = want.get(line >> 5, 0) | 1 << (line & 31)
        assert words == want and all(w >= 0 for w in words), (dbase, m8)

    # pointer doubling over the 32 candidates of a group (g2_double): M = nodes the chain from a lane visits inside the
    # group, J = the first value it meets outside; successors point forward
    for _ in range(300):
        gb = rng.randrange(0, 400)
        J0 = []
        for lane in range(32):
            r = rng.random()
            if r < 0.75:
                J0.append(gb + lane + rng.randint(1, 3))               # a candidate a little further on
            elif r < 0.9:
                J0.append(gb + 32 + rng.randrange(40))                 # behind the group
            else:
                J0.append(rng.choice([0x4000 | rng.randrange(2048), 0xFFFE, 0xFFFD, 0xFFFC]))  # a line / an end
        M, J = [1 << lane for lane in range(32)], list(J0)
        for _round in range(5):
            Mn, Jn = list(M), list(J)
            for lane in range(32):
                d = (J[lane] - gb) & M32
                if d < 32:
                    Mn[lane] = M[lane] | M[d]
                    Jn[lane] = J[d]
            M, J = Mn, Jn
        for lane in range(32):
            nodes, q = 0, gb + lane
            while gb <= q < gb + 32:
                nodes |= 1 << (q - gb)
                q = J0[q - gb]
            assert (M[lane], J[lane]) == (nodes, q), (gb, lane)


def test_fasta_round_flags_model():
    """One 256-entry round of the FASTA kernels (fq_fasta.cuh: fa_flags8 + fa_round_end) as a sequential model: eight
    entries per lane, the parity of the run a lane's group starts in from the nearest non-candidate in the lanes below
    (else from the candidates before the round), the rest from the carry trick -- against the definition, for random
    candidate patterns (dense runs included), partial rounds and every entry state."""
    rng = random.Random(23)

    def clz32(x):
        return 32 - x.bit_length()

    for trial in range(4000):
        nraw = rng.choice([1, 7, 8, 9, 40, 200, 255, 256])
        p = rng.choice([0.1, 0.5, 0.9, 0.99])
        cand = [rng.random() < p for _ in range(nraw)]
        run_in = rng.choice([0, 1, 2, 5])
        cand8, nc8 = [], []
        for lane in range(32):
            nv = nraw - lane * 8
            valid8 = 0xff if nv >= 8 else (0 if nv <= 0 else (1 << nv) - 1)
            c8 = sum(1 << k for k in range(8) if lane * 8 + k < nraw and cand[lane * 8 + k])
            cand8.append(c8 & valid8)
            nc8.append(~c8 & valid8 & 0xff)
        hb = sum(1 << lane for lane in range(32) if nc8[lane])
        got = []
        for lane in range(32):
            below = hb & ((1 << lane) - 1)
            if below:
                ln = 31 - clz32(below)
                cnt = 8 * lane - 1 - (8 * ln + (31 - clz32(nc8[ln])))
            else:
                cnt = 8 * lane + run_in
            c8 = cand8[lane]
            lead_run = c8 & ~(c8 + 1)
            rest = c8 & ~lead_run
            starts = rest & ~(rest << 1)
            runs_e = rest & ~(rest + (starts & 0x55))
            on8 = (lead_run & (0xAA if cnt & 1 else 0x55)) | (runs_e & 0x55) | (rest & ~runs_e & 0xAA)
            got += [bool(on8 >> k & 1) for k in range(8)]
        want, run = [], run_in
        for c in cand:
            want.append(c and run % 2 == 0)
            run = run + 1 if c else 0
        assert got[:nraw] == want and not any(got[nraw:]), (nraw, run_in)
        # the state handed to the next round
        if hb:
            lt = 31 - clz32(hb)
            pos = 8 * lt + (31 - clz32(nc8[lt]))
            nxt = nraw - 1 - pos
        else:
            nxt = run_in + nraw
        assert nxt == run, (nraw, run_in)


def test_lookback_protocol_model():
    """The record-count protocol of the warp-per-chunk speculative pass (two-level decoupled look-back, the flush of a
    chunk deferred behind the publication of the warp's next chunk) under random schedules: no deadlock, and every
    chunk gets the number of records before it -- for few and many warps, partial last blocks, single chunks."""
    rng = random.Random(31)
    for trial in range(150):
        n = rng.choice([1, 2, 31, 32, 33, 64, 100, 257])
        counts = [rng.randint(0, 9) for _ in range(n)]
        n_warps = rng.choice([1, 2, 7, 40, 300])
        bases = am.simulate_lookback(counts, n_warps, rng, block=rng.choice([4, 32]))
        assert bases is not None, (n, n_warps)
        want, acc = [], 0
        for c in counts:
            want.append(acc)
            acc += c
        assert bases == want, (n, n_warps)
