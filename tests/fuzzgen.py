"""Adversarial FASTA / FASTQ generators shared by the CPU (emulation) and GPU parity tests."""
import random


def rand_fasta(rng: random.Random) -> bytes:
    out = bytearray()
    nrec = rng.randint(1, 6)
    style = rng.choice(['wrap', 'wrap', 'short', 'unwrapped', 'mixed', 'crlf', 'blank'])
    for _ in range(nrec):
        hl = rng.choice([5, 20, 100, 600, 1500]) if rng.random() < 0.3 else rng.randint(1, 60)
        hdr = ''.join(rng.choice('ACGTacgtN >|_.0123456789xyz') for _ in range(hl))
        out += b'>' + hdr.encode() + b'\n'
        L = rng.choice([0, 1, 3, 6, 7, 8, 15, 16, 17, 100, 511, 512, 513, 3000, 20000]) if rng.random() < 0.5 \
            else rng.randint(0, 5000)
        alphabet = 'ACGT' * 20 + 'acgt' * 5 + 'N' * (3 if rng.random() < 0.5 else 0) + \
                   ('RYKM-*.>' if rng.random() < 0.3 else '')
        seq = ''.join(rng.choice(alphabet) for _ in range(L))
        if rng.random() < 0.3 and L > 50:
            p = rng.randint(0, L - 10)
            n = rng.randint(1, 40)
            seq = seq[:p] + 'N' * n + seq[p + n:]
        if style == 'unwrapped':
            w = 10 ** 9
        elif style == 'short':
            w = rng.randint(1, 20)
        elif style == 'mixed':
            w = None
        else:
            w = rng.choice([60, 70, 80, 61, 15, 16, 17, 31, 32, 33])
        i = 0
        while i < len(seq):
            ww = w if w else rng.randint(1, 100)
            line = seq[i:i + ww]
            i += ww
            out += line.encode() + (b'\r\n' if style == 'crlf' else b'\n')
            if style == 'blank' and rng.random() < 0.2:
                out += b'\n' * rng.randint(1, 3)
    if rng.random() < 0.3 and out.endswith(b'\n'):
        out = out[:-1]
    return bytes(out)


def rand_fastq(rng: random.Random) -> bytes:
    """Well-formed 4-line FASTQ with nasty content: N runs, lower case, qualities starting with '@'/'+',
    empty reads, reads shorter than k, variable read lengths."""
    out = bytearray()
    nrec = rng.randint(1, 60)
    for r in range(nrec):
        L = rng.choice([0, 1, 6, 7, 8, 15, 16, 17, 31, 32, 33, 100, 150, 151, 500, 513]) if rng.random() < 0.5 \
            else rng.randint(0, 300)
        alphabet = 'ACGT' * 20 + 'acgt' * 3 + 'N' * (2 if rng.random() < 0.5 else 0)
        seq = ''.join(rng.choice(alphabet) for _ in range(L))
        qual = ''.join(chr(rng.randint(33, 74)) for _ in range(L))
        if L and rng.random() < 0.3:
            qual = rng.choice('@+>') + qual[1:]
        hdr = '@r%d %s' % (r, ''.join(rng.choice('ACGT@+>: /') for _ in range(rng.randint(0, 40))))
        plus = '+' + (hdr[1:] if rng.random() < 0.2 else '')
        out += (hdr + '\n' + seq + '\n' + plus + '\n' + qual + '\n').encode()
    if rng.random() < 0.3:
        out = out[:-1]
    return bytes(out)


def rand_fasta_grid(rng: random.Random) -> bytes:
    """Mostly fixed-width FASTA (60/70/80 columns) with the irregularities real files have: short last
    lines, several records, N runs, lower case, IUPAC, an occasional record in another width, blank lines,
    CRLF records, long headers, missing final newline."""
    out = bytearray()
    lw = rng.choice([60, 70, 80, 60, 70, 80, 50, 100, 120])   # (120: no line-kernel instantiation -> generic kernel)
    nrec = rng.randint(1, 5)
    for r in range(nrec):
        hl = rng.choice([5, 20, 79, 80, 81, 100, 600]) if rng.random() < 0.4 else rng.randint(1, 60)
        hdr = ''.join(rng.choice('ACGTacgtN >|_.0123456789xyz') for _ in range(hl))
        out += b'>' + hdr.encode() + b'\n'
        L = rng.choice([0, 1, 7, 79, 80, 81, 160, 161, 2400, 2592, 5000, 20000, 40000]) if rng.random() < 0.6 \
            else rng.randint(0, 30000)
        alphabet = 'ACGT' * 30 + 'acgt' * 3 + ('N' if rng.random() < 0.5 else '') + ('RY-' if rng.random() < 0.2 else '')
        seq = ''.join(rng.choice(alphabet) for _ in range(L))
        if rng.random() < 0.4 and L > 200:
            p = rng.randint(0, L - 100)
            n = rng.randint(1, 90)
            seq = seq[:p] + 'N' * n + seq[p + n:]
        w = lw if rng.random() < 0.85 else rng.choice([60, 70, 80, 50, 100])
        eol = b'\r\n' if rng.random() < 0.05 else b'\n'
        for i in range(0, len(seq), w):
            out += seq[i:i + w].encode() + eol
            if rng.random() < 0.01:
                out += b'\n'
    if rng.random() < 0.3 and out.endswith(b'\n'):
        out = out[:-1]
    return bytes(out)


def rand_fasta_long(rng: random.Random) -> bytes:
    """Long-line ("unwrapped") FASTA as assemblers write it -- one line per contig, kilobases long -- with everything
    that can go wrong around it: contigs shorter than k, N runs, lower case, IUPAC codes, records split over a few long
    lines (k-mers span the '\\n'), blank lines, CRLF, headers of up to 3,000 bytes made of A/C/G/T text (longer than the
    kernel's look-back: a unit boundary inside one must be found out), '>' inside sequence lines, missing final newline.
    At least 32 KiB, so that the width probe takes it for a long-line file."""
    out = bytearray()
    big = rng.random() < 0.5
    while len(out) < (rng.choice([33000, 60000, 150000]) if big else 33000):
        kind = rng.random()
        if kind < 0.15:
            hl = rng.choice([300, 800, 3000])
            hdr = ''.join(rng.choice('ACGT') for _ in range(hl))          # looks like sequence
        elif kind < 0.3:
            hl = rng.choice([79, 80, 81, 255, 256, 257])
            hdr = ''.join(rng.choice('ACGTacgtN >|_.0123456789xyz') for _ in range(hl))
        else:
            hdr = 'contig_%d length=%d' % (len(out), rng.randint(1, 10 ** 6))
        out += b'>' + hdr.encode() + b'\n'
        L = rng.choice([0, 1, 6, 7, 8, 79, 80, 81, 86, 87, 160, 2559, 2560, 2561, 2566, 2640, 5120, 12000, 40000]) if rng.random() < 0.6 \
            else rng.randint(0, 30000)
        alphabet = 'ACGT' * 40 + 'acgt' * 3 + ('N' if rng.random() < 0.4 else '') + ('RY>-' if rng.random() < 0.1 else '')
        seq = ''.join(rng.choice(alphabet) for _ in range(L))
        for _ in range(rng.choice([0, 0, 1, 3])):
            if L > 200:
                p = rng.randint(0, L - 100)
                n = rng.choice([1, 5, 79, 80, 81, 200])
                seq = seq[:p] + 'N' * n + seq[p + n:]
        seq = seq[:L]
        eol = b'\r\n' if rng.random() < 0.04 else b'\n'
        nlines = 1 if rng.random() < 0.8 else rng.randint(2, 4)
        cuts = sorted(rng.randint(0, len(seq)) for _ in range(nlines - 1)) + [len(seq)]
        a = 0
        for c in cuts:
            out += seq[a:c].encode() + eol
            a = c
            if rng.random() < 0.05:
                out += b'\n' * rng.randint(1, 2)
    if rng.random() < 0.3 and out.endswith(b'\n'):
        out = out[:-1]
    return bytes(out)


def rand_fastq_multiline(rng: random.Random) -> bytes:
    """FASTQ as old tools wrote it: sequence and quality wrapped over several lines (widths of their own), qualities that
    begin a line with '@' or '+', reads shorter than k, N runs, lower case, blank lines between records, missing final
    newline -- what Jellyfish reads front to back (oracle: walk_fastq)."""
    out = bytearray()
    for r in range(rng.randint(1, 25)):
        L = rng.choice([0, 1, 6, 7, 8, 13, 60, 61, 150, 500]) if rng.random() < 0.5 else rng.randint(0, 400)
        alphabet = 'ACGT' * 20 + 'acgt' * 3 + 'N' * (2 if rng.random() < 0.5 else 0)
        seq = ''.join(rng.choice(alphabet) for _ in range(L))
        qual = ''.join(chr(rng.randint(33, 74)) for _ in range(L))
        ws, wq = rng.choice([1, 3, 5, 7, 60, 70, 80, 1000]), rng.choice([1, 4, 60, 61, 80, 1000])
        out += ('@r%d %s\n' % (r, ''.join(rng.choice('ACGT@+>: /') for _ in range(rng.randint(0, 30))))).encode()
        for i in range(0, max(L, 1), ws):
            out += seq[i:i + ws].encode() + b'\n'
        out += b'+' + (b'r%d' % r if rng.random() < 0.2 else b'') + b'\n'
        for i in range(0, max(L, 1), wq):
            line = qual[i:i + wq]
            if line and rng.random() < 0.2:
                line = rng.choice('@+') + line[1:]
            out += line.encode() + b'\n'
        if rng.random() < 0.1:
            out += b'\n'
    if rng.random() < 0.3 and out.endswith(b'\n'):
        out = out[:-1]
    return bytes(out)
