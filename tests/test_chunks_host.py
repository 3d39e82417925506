"""Host-side logic of the chunked-genome mode (kf2vecfsw_b200/chunks.py) against the oracle's restatement of
kf2vec/main.py:654-929 -- linearise, N-run collapse, gap strip, 10-kbp filter, sliding-window plan, row labels.
No GPU: the counting call itself is covered by tests/test_gpu_parity.py."""
import random

import numpy as np

import kf_oracle as o
from kf2vecfsw_b200 import chunks


def test_window_plan_equals_oracle_and_covers_contig():
    for L in list(range(10000, 10050)) + [19999, 20000, 20001, 25000, 99999, 100000, 123457, 1234567]:
        w = chunks.window_plan(L)
        assert w == o.window_plan(L)
        assert w[0][0] == 1 and all(b - a + 1 == chunks.CHUNK_SZ for a, b in w)
        assert w[-1][1] <= L and L - w[-1][1] < chunks.CHUNK_SZ


def test_plan_genome_equals_oracle_on_toy(toy_inputs):
    for s in ("G000402355", "G000830275", "G000830295"):
        seq, offs, lens, labels = chunks.plan_genome(s, toy_inputs[s])
        ref = o.chunk_rows(s, toy_inputs[s], 7)
        assert labels == [l for l, _ in ref]
        assert len(labels) in (117, 125, 116)
        # window bytes are what the oracle counts: spot-check rows through the oracle's own counter
        for i in (0, len(labels) // 2, len(labels) - 1):
            sub = seq[int(offs[i]): int(offs[i]) + int(lens[i])]
            cnt = o.fold_canonical(o.forward_counts(o._CODE_LUT[np.frombuffer(sub, dtype=np.uint8)], 7), 7)
            assert np.array_equal(cnt, ref[i][1])


def test_plan_genome_text_rules():
    rng = random.Random(5)
    body = "".join(rng.choice("ACGT") for _ in range(12000))
    gapped = body[:3000] + "NNNNnnNN|N" + body[3000:6000] + "-. -" + body[6000:]
    wrapped = "\n".join(gapped[i:i + 70] for i in range(0, len(gapped), 70))
    data = (">c1 some description\n" + wrapped + "\n>c2\nACGTACGT\n>c3\n" + body[:9999] + "\n").encode()
    seq, offs, lens, labels = chunks.plan_genome("s", data)
    assert seq.tobytes() == (body[:3000] + "N" + body[3000:]).encode()          # run collapsed to one N, gaps removed
    assert labels == ["s.part_c1.part_c1_sliding__1-10000", "s.part_c1.part_c1_sliding__2002-12001"]
    assert list(offs) == [0, 2001] and list(lens) == [10000, 10000]
    assert chunks.plan_genome("s", b"")[3] == [] and chunks.plan_genome("s", b"@fq\nACGT\n+\nIIII\n")[3] == []


def test_linearise_fasta_equals_oracle_text_rules():
    """kf_linearise_fasta (one C++ pass) against the oracle's three separate steps on nasty records: N runs that span
    line breaks, '|' in the run class, gaps between Ns, CRLF, empty records, a missing final newline."""
    from kf2vecfsw_b200 import engine
    rng = random.Random(9)
    for trial in range(40):
        recs = []
        for r in range(rng.randint(1, 6)):
            L = rng.choice([0, 5, 9999, 10000, 10001, 25000])
            s = "".join(rng.choice("ACGT" * 8 + "Nn|-. ") for _ in range(L))
            w = rng.choice([60, 70, 80, 10 ** 9])
            eol = "\r\n" if rng.random() < 0.2 else "\n"
            recs.append(">c%d desc %d" % (r, trial) + eol + eol.join(s[i:i + w] for i in range(0, len(s), w)) + (eol if s else ""))
        data = "".join(recs).encode()
        if rng.random() < 0.3 and data.endswith(b"\n"):
            data = data[:-1]
        seq, got = engine.linearise_fasta(data, 10000)
        want = []
        for header, s in o.fasta_records(data):
            s = o.strip_gaps(o.collapse_n_runs(s))
            if len(s) >= 10000:
                want.append((header, s))
        assert [h for h, _, _ in got] == [h.rstrip("\r") for h, _ in want]   # (the CR of a CRLF header is not part of the id)
        for (h, off, ln), (_, s) in zip(got, want):
            assert seq[off:off + ln].tobytes() == s
