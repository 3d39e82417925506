"""The oracle (oracle/kf_oracle.py, oracle/kf_oracle.c) against the reference's own committed outputs.

Golden vectors: reference toy_example/{train_tree_kf,test_kf}/*.kf (7 reproducible files) and
toy_example/train_tree_chunks/*.kf (358 rows), committed under tests/golden by tools/make_golden.py.
"""
import hashlib
import json
import os
import random

import numpy as np
import pytest

import kf_oracle as o
import c_oracle
from fuzzgen import rand_fasta, rand_fastq


def test_fixture_integrity(toy_inputs, toy_golden_kf, golden_dir):
    man = json.load(open(os.path.join(golden_dir, "manifest.json")))
    assert len(man) == 7
    for s, m in man.items():
        assert hashlib.sha256(toy_inputs[s]).hexdigest() == m["fna_sha256"]
        assert hashlib.sha256(toy_golden_kf[s].encode()).hexdigest() == m["kf_sha256"]


@pytest.mark.parametrize("k", [3, 4, 5, 6, 7, 8, 9])
def test_vocabulary_matches_reference_files(k, golden_dir):
    sha = json.load(open(os.path.join(golden_dir, "vocab_sha256.json")))[str(k)]
    text = "".join(w + "\n" for w in o.vocab(k))
    assert hashlib.sha256(text.encode()).hexdigest() == sha
    assert len(o.vocab(k)) == o.vocab_size(k)


def test_numpy_oracle_reproduces_golden_kf_byte_exact(toy_inputs, toy_golden_kf):
    for s, data in toy_inputs.items():
        counts = o.canonical_counts_bytes(data, 7)
        vals, int_mode = o.row_values(counts, pseudocount=False, raw_cnt=False)
        assert o.format_kf_line(s, vals, int_mode) == toy_golden_kf[s], s


def test_c_oracle_matches_numpy_oracle_on_toy(toy_inputs):
    for s, data in toy_inputs.items():
        for k in (5, 7):
            assert np.array_equal(c_oracle.count_buffer(data, k), o.canonical_counts_bytes(data, k)), (s, k)


def test_chunk_rows_match_golden(toy_inputs, golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "chunks_golden.json")))
    total = 0
    for s, rows in gold.items():
        ours = o.chunk_rows(s, toy_inputs[s], 7)
        assert len(ours) == len(rows)
        gd = dict(rows)
        for label, cnt in ours:
            vals, int_mode = o.row_values(cnt, pseudocount=False, raw_cnt=True)
            line = o.format_kf_line(label, vals, int_mode)
            assert hashlib.sha256(line.encode()).hexdigest() == gd[label], label
            total += 1
    assert total == 358


def test_slow_python_cross_check():
    rng = random.Random(7)
    for _ in range(10):
        data = rand_fasta(rng)
        for k in (3, 7):
            assert np.array_equal(o.canonical_counts_slow(data, k), o.canonical_counts_bytes(data, k))


def test_c_oracle_matches_numpy_oracle_on_fuzz():
    rng = random.Random(11)
    for i in range(60):
        data = rand_fasta(rng) if i % 3 else rand_fastq(rng)
        k = rng.choice([1, 3, 5, 7, 9])
        assert np.array_equal(c_oracle.count_buffer(data, k), o.canonical_counts_bytes(data, k)), i


def test_fastq_semantics_small():
    # qualities that look like headers / plus lines are skipped by length; N breaks the window
    fq = b"@r1\nACGTACGTAC\n+\n@+@+@+@+@+\n@r2\nACGTNACGTACG\n+r2\n++++++++++++\n"
    c = o.canonical_counts_bytes(fq, 4)
    assert int(c.sum()) == (10 - 3) + (4 - 3) + (7 - 3)
    # multi-line FASTQ (jellyfish tolerates it): same k-mers as the joined read
    ml = b"@r1\nACGTAC\nGTAC\n+\n@+@+@+\n@+@+\n"
    sl = b"@r1\nACGTACGTAC\n+\n@+@+@+@+@+\n"
    assert np.array_equal(o.canonical_counts_bytes(ml, 4), o.canonical_counts_bytes(sl, 4))


def test_unpinned_semantics_are_as_documented():
    up = b">a\nACGTACGTAGGCTA\n"
    assert np.array_equal(o.canonical_counts_bytes(up.lower().replace(b">A", b">a"), 5), o.canonical_counts_bytes(up, 5))
    # CRLF: '\r' is a non-ACGT byte, so k-mers do not span lines
    lf = b">a\nACGTACGT\nACGTACGT\n"
    crlf = lf.replace(b"\n", b"\r\n")
    assert int(o.canonical_counts_bytes(lf, 5).sum()) == 12
    assert int(o.canonical_counts_bytes(crlf, 5).sum()) == 8
    # records never join; '>' only opens a header at a line start
    assert int(o.canonical_counts_bytes(b">a\nACGT\n>b\nACGT\n", 5).sum()) == 0
    assert int(o.canonical_counts_bytes(b">a\nACGT>ACGT\n", 4).sum()) == 2
    with pytest.raises(o.FormatError):
        o.symbols_from_bytes(b"ACGT\n")
    with pytest.raises(o.FormatError):
        o.symbols_from_bytes(b"")


def test_row_values_and_format_quirks():
    c = np.array([1, 2, 3], dtype=np.uint64)
    v, im = o.row_values(c, False, True)
    assert im and o.format_kf_line("s", v, im) == "s,1,2,3\n"
    v, im = o.row_values(np.array([1, 0, 3], dtype=np.uint64), False, True)
    assert not im and o.format_kf_line("s", v, im) == "s,1.0,0.0,3.0\n"
    v, im = o.row_values(c, True, True)
    assert o.format_kf_line("s", v, im) == "s,1.5,2.5,3.5\n"
    v, im = o.row_values(np.zeros(3, dtype=np.uint64), False, False)
    assert o.format_kf_line("s", v, im) == "s,nan,nan,nan\n"
