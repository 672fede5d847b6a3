"""Seeded FASTQ / byte-soup generators shared by the CPU and GPU parity tests."""
import random

import numpy as np

BASES = b'ACGTN'
QUALS = bytes(range(33, 74))  # '!'..'I' : includes '+' (43) and '@' (64)


def _wrap(s, width):
    return b'\n'.join(s[i:i + width] for i in range(0, len(s), width)) if width else s


def fastq_bytes(rng, n_records, read_len=(1, 40), header_len=(1, 12), wrap=0, long_plus=0.0,
                trailing_newlines=1, at_plus_bias=0.0):
    """Well-formed FASTQ.  wrap > 0 wraps sequence and quality identically at that column."""
    out = []
    for k in range(n_records):
        hl = rng.randint(*header_len)
        rl = rng.randint(*read_len)
        head = bytes(rng.choice(b'abcXYZ09:/ #@+') for _ in range(hl))
        seq = bytes(rng.choice(BASES) for _ in range(rl))
        if rng.random() < at_plus_bias:
            qual = bytes(rng.choice(b'@+') for _ in range(rl))
        else:
            qual = bytes(rng.choice(QUALS) for _ in range(rl))
        plus = b'+' + (head if rng.random() < long_plus else b'')
        out.append(b'@' + head + b'\n' + _wrap(seq, wrap) + b'\n' + plus + b'\n' + _wrap(qual, wrap))
    data = b'\n'.join(out)
    if n_records:
        data += b'\n' * trailing_newlines
    return data


def mutate(rng, data, n_mut=1):
    """Damage a FASTQ buffer: delete / insert / replace a few bytes, or truncate."""
    b = bytearray(data)
    for _ in range(n_mut):
        if not b:
            break
        op = rng.randrange(5)
        i = rng.randrange(len(b))
        if op == 0:
            del b[i]
        elif op == 1:
            b.insert(i, rng.choice(b'\n@+AI'))
        elif op == 2:
            b[i] = rng.choice(b'\n@+AI\r')
        elif op == 3:
            del b[i:]
        else:
            j = min(len(b), i + rng.randint(1, 30))
            del b[i:j]
    return bytes(b)


def fasta_bytes(rng, n_records=None, leading_newline=None):
    """FASTA-like input: wrapped / unwrapped / empty sequences, '>' inside lines, runs of header-only records,
    optional damage (soup, truncation), with or without a trailing newline."""
    kind = rng.randrange(10)
    if kind == 0:
        return soup(rng, rng.randint(0, 60), b'\n\n>>ACGT')
    n = rng.randint(0, 12) if n_records is None else n_records
    out = []
    for _ in range(n):
        head = bytes(rng.choice(b'abcXYZ09:/ #>') for _ in range(rng.randint(0, 14)))
        length = rng.choice([0, 0, 1, 5, 30, 61, 200])
        seq = bytes(rng.choice(b'ACGTN>') if rng.random() < 0.02 else rng.choice(b'ACGTN') for _ in range(length))
        wrap = rng.choice([0, 0, 7, 60])
        out.append(b'>' + head + b'\n' + _wrap(seq, wrap))
    lead = rng.random() < 0.7 if leading_newline is None else leading_newline
    data = (b'\n' if lead else b'') + b'\n'.join(out) + (b'\n' * rng.choice([0, 1, 1, 2]) if out else b'')
    if kind == 1 and data:
        data = mutate(rng, data, rng.randint(1, 3))
    return data


def soup(rng, n, alphabet=b'\n\n@+AI'):
    return bytes(rng.choice(alphabet) for _ in range(n))


def corpus(seed, n_cases):
    """A mixed bag of small inputs: clean, multi-line, long '+', damaged, garbage-led, byte soup."""
    rng = random.Random(seed)
    for c in range(n_cases):
        kind = rng.randrange(10)
        n = rng.randint(0, 12)
        if kind <= 2:
            d = fastq_bytes(rng, n, trailing_newlines=rng.randint(0, 3), at_plus_bias=rng.choice([0, 0.5]))
        elif kind == 3:
            d = fastq_bytes(rng, n, long_plus=0.5, trailing_newlines=rng.randint(0, 3))
        elif kind == 4:
            d = fastq_bytes(rng, n, read_len=(1, 60), wrap=rng.randint(3, 20), long_plus=0.3,
                            trailing_newlines=rng.randint(0, 2), at_plus_bias=rng.choice([0, 0.5]))
        elif kind == 5:
            d = mutate(rng, fastq_bytes(rng, n, trailing_newlines=1, at_plus_bias=0.3), rng.randint(1, 3))
        elif kind == 6:
            d = soup(rng, rng.randint(0, 8)) + fastq_bytes(rng, n, trailing_newlines=rng.randint(0, 2))
        elif kind == 7:
            d = soup(rng, rng.randint(0, 80))
        elif kind == 8:
            d = mutate(rng, fastq_bytes(rng, n, wrap=rng.randint(3, 9), long_plus=0.5), rng.randint(1, 2))
        else:
            d = fastq_bytes(rng, n, read_len=(1, 6), header_len=(0, 2), trailing_newlines=rng.randint(0, 3),
                            at_plus_bias=0.7)
        yield d


# ---- numpy generators for the larger GPU parity cases (BASELINE.json config shapes) ----------------
def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def fixed_records_np(n_records, header_len=32, read_len=150, seed=0xB2000002):
    """numpy twin of the device generator fqb_synth_fixed (csrc/fq_synth.cuh): byte g of the buffer
    depends only on (seed, g).  Record = '@SIM:' + zero padded decimal index + ' 1:N:0:ACGTACGT' \\n
    bases \\n + \\n quals \\n ; bases uniform ACGT, qualities uniform '!'..'I' (so '+' and '@' occur)."""
    rec = header_len + 1 + read_len + 1 + 2 + read_len + 1
    with np.errstate(over='ignore'):
        g = np.arange(n_records * rec, dtype=np.uint64)
        k = g // np.uint64(rec)
        o = (g % np.uint64(rec)).astype(np.int64)
        h = splitmix64(np.uint64(seed) ^ g)
    out = np.empty(n_records * rec, dtype=np.uint8)
    width = header_len - 20
    suffix = np.frombuffer(b' 1:N:0:ACGTACGT', dtype=np.uint8)
    prefix = np.frombuffer(b'@SIM:', dtype=np.uint8)
    # header
    m = o < 5
    out[m] = prefix[o[m]]
    m = (o >= 5) & (o < 5 + width)
    digit_pos = (width - 1 - (o[m] - 5)).astype(np.uint64)
    out[m] = (48 + (k[m] // (np.uint64(10) ** digit_pos)) % np.uint64(10)).astype(np.uint8)
    m = (o >= 5 + width) & (o < header_len)
    out[m] = suffix[o[m] - 5 - width]
    out[o == header_len] = 10
    s0 = header_len + 1
    m = (o >= s0) & (o < s0 + read_len)
    out[m] = np.frombuffer(b'ACGT', dtype=np.uint8)[(h[m] >> np.uint64(33)) & np.uint64(3)]
    out[o == s0 + read_len] = 10
    out[o == s0 + read_len + 1] = 43
    out[o == s0 + read_len + 2] = 10
    q0 = s0 + read_len + 3
    m = (o >= q0) & (o < q0 + read_len)
    out[m] = (33 + (h[m] >> np.uint64(33)) % np.uint64(41)).astype(np.uint8)
    out[o == q0 + read_len] = 10
    return out


def variable_records_np(n_records, seed, kind):
    """Illumina-like ('illumina'), ONT-like long reads ('ont') and wrapped multi-line + long '+'
    header ('multiline') buffers, SURVEY.md 8d configs 3-5, built record by record with numpy."""
    rng = np.random.default_rng(seed)
    parts = []
    acgt = np.frombuffer(b'ACGT', dtype=np.uint8)
    for k in range(n_records):
        if kind == 'illumina':
            head = ('@A00123:45:HXXXXXXXX:%d:%d:%d:%d 1:N:0:ACGTACGT' % (
                rng.integers(1, 5), rng.integers(1101, 2679), rng.integers(1000, 32001),
                rng.integers(1000, 50001))).encode()
            rl = 150
            if rng.random() < 0.1:
                q = rng.integers(33, 74, rl).astype(np.uint8)
            else:
                q = np.frombuffer(b'#,:F', dtype=np.uint8)[rng.integers(0, 4, rl)]
            wrap, plus = 0, b'+'
        elif kind == 'ont':
            head = ('@%08x-%04x-%04x-%04x-%012x runid=%040x read=%d ch=%d start_time=2026-01-01T00:00:00Z' % (
                rng.integers(0, 2 ** 32), rng.integers(0, 2 ** 16), rng.integers(0, 2 ** 16),
                rng.integers(0, 2 ** 16), rng.integers(0, 2 ** 48), int(rng.integers(0, 2 ** 62)),
                k, rng.integers(1, 513))).encode()
            rl = int(np.clip(rng.gamma(2.0, 5000.0), 200, 500000))
            q = rng.integers(34, 84, rl).astype(np.uint8)
            wrap, plus = 0, b'+'
        else:
            head = ('@SIM:%09d:%d len' % (k, rng.integers(0, 10 ** 6))).encode()
            rl = int(rng.integers(150, 301))
            q = rng.integers(33, 74, rl).astype(np.uint8)
            wrap = 60
            plus = (b'+' + head[1:]) if rng.random() < 0.5 else b'+'
        s = acgt[rng.integers(0, 4, rl)]
        if wrap:
            s = _wrap(s.tobytes(), wrap)
            q = _wrap(q.tobytes(), wrap)
        else:
            s, q = s.tobytes(), q.tobytes()
        parts.append(head + b'\n' + s + b'\n' + plus + b'\n' + q + b'\n')
    return np.frombuffer(b''.join(parts), dtype=np.uint8)


# ---- numpy twin of the device generator for variable record geometry (csrc/fq_synth.cuh) ----------------
SYNTH_SEEDS = {'illumina': 0xB2000003, 'ont': 0xB2000004, 'multiline': 0xB2000005}


def _splitmix_int(x):
    """splitmix64 on a Python int."""
    m = (1 << 64) - 1
    x = (x + 0x9E3779B97F4A7C15) & m
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return z ^ (z >> 31)


def synth_record_meta(kind, k, seed, qtable=None):
    """(header line bytes incl. '@', read length, wrap, '+' line bytes, quality mode) of record k -- restates
    synth_rec / synth_header of csrc/fq_synth.cuh with Python integers and string formatting."""
    h = [_splitmix_int(seed ^ ((1 << 63) | k))]
    for _ in range(5):
        h.append(_splitmix_int(h[-1]))
    if kind == 'illumina':
        head = '@A00123:45:HXXXXXXXX:%d:%d:%d:%d 1:N:0:ACGTACGT' % (
            1 + (h[0] & 3), 1101 + ((h[0] >> 2) & 0xffff) % 1578, 1000 + ((h[0] >> 18) & 0xfffff) % 31001,
            1000 + ((h[0] >> 38) & 0xfffff) % 49001)
        return head.encode(), 150, 0, b'+', (0 if (h[1] & 0xffff) % 10 == 0 else 1)
    if kind == 'ont':
        head = '@%08x-%04x-%04x-%04x-%012x runid=%016x%016x%08x read=%d ch=%d start_time=2026-01-01T00:00:00Z' % (
            (h[0] >> 32) & 0xffffffff, (h[0] >> 16) & 0xffff, h[0] & 0xffff, (h[1] >> 48) & 0xffff, h[1] & 0xffffffffffff,
            h[2], h[3], h[4] & 0xffffffff, k, 1 + (h[5] & 0xffff) % 512)
        u = h[5] >> 40
        idx, frac = u >> 12, u & 4095
        a, b = int(qtable[idx]), int(qtable[idx + 1])
        return head.encode(), a + (((b - a) * frac) >> 12), 0, b'+', 2
    head = ('@SIM:%09d:%d len' % (k % 1000000000, (h[0] >> 20) % 1000000)).encode()
    return head, 150 + (h[1] >> 8) % 151, 60, (b'+' + head[1:]) if (h[1] & 1) else b'+', 0


def synth_records_np(kind, n_records, seed=None, qtable=None):
    """Records 0 .. n_records-1 of the synthetic stream `kind` exactly as fqb_synth_fill writes them.
    Returns (uint8 array of the bytes, int64 [n,6] true offset table, int64 [n+1] record offsets)."""
    seed = SYNTH_SEEDS[kind] if seed is None else seed
    metas = [synth_record_meta(kind, k, seed, qtable) for k in range(n_records)]
    lens, rows = [], []
    off = 0
    for head, rl, wrap, plus, qmode in metas:
        sb = rl + ((rl - 1) // wrap if wrap else 0)
        p0, p1 = off, off + len(head)
        p3 = p1 + 1 + sb
        p4 = p3 + 1 + len(plus) + 1
        rows.append((p0, p1, p1 + 1, p3, p4, p4 + sb))
        off = p4 + sb + 1
        lens.append(off - p0)
    total = off
    out = np.empty(total, dtype=np.uint8)
    with np.errstate(over='ignore'):
        g = np.arange(total, dtype=np.uint64)
        v = (splitmix64(np.uint64(seed) ^ g) >> np.uint64(33)).astype(np.uint32)
    base = np.frombuffer(b'ACGT', dtype=np.uint8)[v & 3]
    q_modes = (
        (33 + v % 41).astype(np.uint8),
        np.frombuffer(b'#,:F', dtype=np.uint8)[v & 3],
        (34 + v % 50).astype(np.uint8),
    )
    for (head, rl, wrap, plus, qmode), (p0, p1, p2, p3, p4, p5) in zip(metas, rows):
        out[p0:p1] = np.frombuffer(head, dtype=np.uint8)
        out[p1] = 10
        out[p2:p3] = base[p2:p3]
        out[p3] = 10
        out[p3 + 1:p3 + 1 + len(plus)] = np.frombuffer(plus, dtype=np.uint8)
        out[p4 - 1] = 10
        out[p4:p5] = q_modes[qmode][p4:p5]
        out[p5] = 10
        if wrap:
            f = np.arange(p3 - p2)
            nl = f[(f % (wrap + 1)) == wrap]
            out[p2 + nl] = 10
            out[p4 + nl] = 10
    offs = np.array([r[0] for r in rows] + [total], dtype=np.int64)
    return out, np.array(rows, dtype=np.int64).reshape(-1, 6), offs
