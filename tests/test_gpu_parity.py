"""GPU parity: the CUDA path (through the C ABI of libkfcount.so) against the oracle and against the
reference's committed golden .kf files.  Integer counts bit-exact; frequencies bit-exact fp64 (the bar
in BASELINE.json is 1e-7 relative; one correctly-rounded IEEE division gives equality)."""
import argparse
import os
import random

import numpy as np
import pytest

import kf_oracle as o
import c_oracle
import kfsynth
from fuzzgen import rand_fasta, rand_fasta_grid, rand_fastq

pytestmark = pytest.mark.gpu

REL_TOL = 1e-7   # BASELINE.json north_star tolerance for normalised frequencies


@pytest.fixture(scope="module")
def eng():
    from kf2vecfsw_b200 import engine
    assert os.path.exists(engine.lib_path()), "libkfcount.so missing: the CUDA path must be the one under test"
    engine.init(0)
    return engine


def check_against_oracle(eng, bufs, k, **kw):
    counts, freq, totals, status = eng.count_buffers(bufs, k=k, **kw)
    assert (status == 0).all(), status
    for i, b in enumerate(bufs):
        ref = o.canonical_counts_bytes(bytes(b), k)
        assert np.array_equal(counts[i], ref), (i, k, int(ref.sum()), int(counts[i].sum()))
        assert int(totals[i]) == int(ref.sum())
        vals, _ = o.row_values(ref, kw.get("pseudocount", False), kw.get("raw_cnt", False))
        if ref.sum() > 0 or kw.get("pseudocount", False) or kw.get("raw_cnt", False):
            assert np.allclose(freq[i], vals, rtol=REL_TOL, atol=0)
            assert np.array_equal(freq[i], vals)          # in practice bit-exact
        else:
            assert np.isnan(freq[i]).all()
    return counts, freq


def test_get_frequencies_reproduces_reference_golden_kf(eng, toy_inputs, toy_golden_kf, tmp_path, capsys):
    from kf2vecfsw_b200 import get_frequencies
    ind, outd = tmp_path / "in", tmp_path / "out"
    ind.mkdir(); outd.mkdir()
    for s, data in toy_inputs.items():
        (ind / (s + ".fna")).write_bytes(data)
    get_frequencies(argparse.Namespace(input_dir=str(ind), output_dir=str(outd), k=7, p=4, pseudocount=False,
                                       raw_cnt=False))
    for s in toy_inputs:
        assert (outd / (s + ".kf")).read_text() == toy_golden_kf[s], s
    out = capsys.readouterr().out
    assert "==> Starting k-mer counting for" in out and ">>> Normalizing. Sample:" in out and "==> Done processing" in out
    assert sorted(os.listdir(outd)) == sorted(s + ".kf" for s in toy_inputs)    # no .jf/.dump left behind


def test_wrapper_namespace_without_raw_cnt(eng, toy_inputs, toy_golden_kf, tmp_path):
    """process_query_data / build_library parsers carry no raw_cnt attribute (main.py:1253-1353)."""
    from kf2vecfsw_b200 import get_frequencies
    ind, outd = tmp_path / "in", tmp_path / "out"
    ind.mkdir(); outd.mkdir()
    (ind / "G000830275sub.fna").write_bytes(toy_inputs["G000830275sub"])
    get_frequencies(argparse.Namespace(input_dir=str(ind), output_dir=str(outd), k=7, p=1, pseudocount=False))
    assert (outd / "G000830275sub.kf").read_text() == toy_golden_kf["G000830275sub"]


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
def test_toy_counts_all_k(eng, toy_inputs, k):
    names = ["G000830275sub", "G000402355sub", "G000830295"]
    check_against_oracle(eng, [toy_inputs[s] for s in names], k)


def test_flags_pseudocount_and_raw(eng, toy_inputs):
    bufs = [toy_inputs["G000830275sub"], toy_inputs["G001020955"]]
    check_against_oracle(eng, bufs, 7, pseudocount=True)
    check_against_oracle(eng, bufs, 7, raw_cnt=True)
    check_against_oracle(eng, bufs, 7, pseudocount=True, raw_cnt=True)


def test_walker_and_fast_path_agree(eng, toy_inputs):
    bufs = list(toy_inputs.values())
    a = eng.count_buffers(bufs, k=7)[0]
    b = eng.count_buffers(bufs, k=7, force_walker=True)[0]
    assert np.array_equal(a, b)


@pytest.mark.parametrize("seed0", [0, 5000])
def test_fuzz_fasta(eng, seed0):
    for s in range(seed0, seed0 + 40):
        rng = random.Random(s)
        bufs = [rand_fasta(rng) for _ in range(rng.randint(1, 12))]
        k = rng.choice([1, 3, 5, 7, 7, 8, 9])
        check_against_oracle(eng, bufs, k)


def test_edge_inputs(eng):
    bufs = [b">only header", b">h\n", b">h\nACGTAC", b">h\nACGTACG", b">h\nACGTACG\n", b">h\n" + b"A" * 100000,
            b">h\n" + b"N" * 5000 + b"\n", b">a\n>b\n>c\nACGTTTTGGA\n", b">h\r\nACGTACGTAA\r\nACGTACGTAA\r\n",
            b">h\n" + b"\n" * 2000 + b"ACGTACGTACGT" + b"\n" * 700 + b"ACGTTGCA\n",
            b">" + b"ACGT" * 1000 + b"\nACGTACGTACGTAAA\n"]
    check_against_oracle(eng, bufs, 7)
    check_against_oracle(eng, bufs, 3)
    # unsupported / empty inputs are reported per file, the rest of the batch still counts
    counts, freq, totals, status = eng.count_buffers([b"ACGT\n", b"", b">h\nACGTACGTAC\n"], k=5)
    assert list(status) == [-5, -9, 0] and int(totals[2]) == 6


def test_synthetic_genomes_vs_c_oracle_and_properties(eng):
    """Config-2 style genomes at a size the oracle finishes in seconds, plus size-independent properties."""
    gen = [kfsynth.synth_fasta(20261018, i, 2_000_000) for i in range(6)]
    counts, freq, totals, status = eng.count_buffers(gen, k=7)
    assert (status == 0).all()
    ref, _, st = c_oracle.count_buffers_mt(gen, 7, threads=4, want_freq=False)
    assert np.array_equal(counts, ref)
    # property: total = sum over maximal ACGT runs of max(0, L-6)
    for i, gbuf in enumerate(gen):
        sym = o.symbols_from_bytes(gbuf.tobytes())
        good = np.concatenate(([0], (sym >= 0).astype(np.int8), [0]))
        d = np.diff(good)
        runs = np.flatnonzero(d == -1) - np.flatnonzero(d == 1)
        assert int(totals[i]) == int(np.maximum(runs - 6, 0).sum())
        assert abs(freq[i].sum() - 1.0) < 1e-12
    # property: permuting the batch permutes the rows; duplicating a file duplicates its row
    perm = [3, 0, 5, 1, 1]
    c2 = eng.count_buffers([gen[j] for j in perm], k=7)[0]
    assert np.array_equal(c2, counts[perm])


def test_device_arena_path_matches_host_path(eng, toy_inputs):
    import torch
    bufs = [np.frombuffer(toy_inputs[s], dtype=np.uint8) for s in sorted(toy_inputs)]
    arena = eng.DeviceArena(bufs)
    V = eng.vocab_size(7)
    counts = torch.zeros((arena.n, V), dtype=torch.int64, device="cuda")
    freq = torch.zeros((arena.n, V), dtype=torch.float64, device="cuda")
    feat = torch.zeros((arena.n, V), dtype=torch.float32, device="cuda")
    totals = torch.zeros(arena.n, dtype=torch.int64, device="cuda")
    eng.count_device(arena, k=7, counts=counts, freq=freq, feat=feat, totals=totals)
    torch.cuda.synchronize()
    assert eng.last_launch_count() >= 2
    assert eng.last_count_kernel_ms() > 0
    hc, hf, ht, _ = eng.count_buffers(bufs, k=7)
    assert np.array_equal(counts.cpu().numpy().astype(np.uint64), hc)
    assert np.array_equal(freq.cpu().numpy(), hf)
    # the matrix the trainers see: fp32(fp64 freq * 1e4)  (train_classifier_model.py:149,323)
    assert np.array_equal(feat.cpu().numpy(), (hf * 1e4).astype(np.float32))


def test_linegrid_and_generic_kernels_agree(eng, toy_inputs):
    bufs = list(toy_inputs.values())
    a = eng.count_buffers(bufs, k=7)[0]
    b = eng.count_buffers(bufs, k=7, no_linegrid=True)[0]
    assert np.array_equal(a, b)
    for i, data in enumerate(bufs):
        assert np.array_equal(a[i], o.canonical_counts_bytes(data, 7))


@pytest.mark.parametrize("seed0", [0, 7000])
def test_fuzz_fixed_width_fasta(eng, seed0):
    for s in range(seed0, seed0 + 40):
        rng = random.Random(s)
        bufs = [rand_fasta_grid(rng) for _ in range(rng.randint(1, 12))]
        check_against_oracle(eng, bufs, 7)


def test_linegrid_u16_overflow_recount(eng):
    """Every CTA sees far more than 65,535 identical 8-mer pairs: the pair histogram's 16-bit halves wrap, the
    checksum catches it and the exact recount path must still give the oracle's numbers."""
    seq = b"A" * 40_000_000 + b"ACGTTGCAAGGCTTAACCGGTTAA" * 1000 + b"T" * 1_000_003
    lines = np.frombuffer(seq, dtype=np.uint8)
    n = len(seq)
    full = n // 80
    body = np.empty((full, 81), dtype=np.uint8)
    body[:, :80] = lines[: full * 80].reshape(full, 80)
    body[:, 80] = 10
    data = b">poly\n" + body.tobytes() + seq[full * 80:] + b"\n"
    counts, freq, totals, status = eng.count_buffers([data], k=7)
    ref = c_oracle.count_buffer(data, 7)
    assert status[0] == 0 and np.array_equal(counts[0], ref)
    assert int(ref.max()) > 40_000_000
    # same for the rare-path ("singles") histogram: every line holds an N, so all of it goes through the byte walker
    data = b">polyA_with_N\n" + (b"A" * 40 + b"N" + b"A" * 39 + b"\n") * 400_000
    counts, freq, totals, status = eng.count_buffers([data], k=7)
    ref = c_oracle.count_buffer(data, 7)
    assert status[0] == 0 and np.array_equal(counts[0], ref)


# ---- FASTQ (BASELINE.json configs[3]: process_query_data preprocessing path) -----------------------------------
@pytest.mark.parametrize("seed0", [0, 9000])
def test_fuzz_fastq(eng, seed0):
    for s in range(seed0, seed0 + 40):
        rng = random.Random(s)
        bufs = [rand_fastq(rng) if rng.random() < 0.8 else rand_fasta(rng) for _ in range(rng.randint(1, 12))]
        k = rng.choice([1, 3, 5, 7, 7, 8, 9])
        check_against_oracle(eng, bufs, k)


def test_synthetic_fastq_reads_vs_c_oracle(eng):
    """Config-4 style reads: 150 bp, random strand, 0.2 % N per base, 1 % reads with an N run, qualities that may
    start with '@' or '+'; sized so the C oracle finishes in seconds."""
    samples = [kfsynth.synth_fastq(20261018, i, 1_000_000, 60_000, 150) for i in range(3)]
    assert any(b"\n@" in bytes(s[:200000]).replace(b"\n@g", b"") for s in samples)   # a quality line starts with '@'
    for k in (7, 9):
        counts, freq, totals, status = eng.count_buffers(samples, k=k)
        assert (status == 0).all()
        ref, _, st = c_oracle.count_buffers_mt(samples, k, threads=4, want_freq=False)
        assert np.array_equal(counts, ref)
    # property: reads are independent records -- reversing the record order leaves the row unchanged
    recs = bytes(samples[0]).split(b"\n")
    recs = [b"\n".join(recs[i:i + 4]) for i in range(0, len(recs) - 1, 4)]
    rev = b"\n".join(reversed(recs)) + b"\n"
    a = eng.count_buffers([samples[0]], k=7)[0]
    b = eng.count_buffers([rev], k=7)[0]
    assert np.array_equal(a, b)


def test_fastq_edge_inputs_and_layout_check(eng):
    ok = [b"@r\nACGTACGTAC\n+\nIIIIIIIIII\n", b"@r\nACGTACGTAC\n+\nIIIIIIIIII", b"@r\n\n+\n\n", b"@r\nACGTACG\n+r\n@@@@@@@\n",
          b"@r\nACGTACGTAC\n+\n+IIIIIIIII\n@s\nTTTTTTTTTT\n+\n@IIIIIIIII\n\n\n", b"@only header", b"@h\nACGTACGTACGT",
          b"@r\n" + b"ACGTTGCA" * 5000 + b"\n+\n" + b"I" * 40000 + b"\n"]
    check_against_oracle(eng, ok, 7)
    check_against_oracle(eng, ok, 4)
    # multi-line FASTQ is not in 4-line layout: the record-chasing kernel reports it and the file is read front to back
    # by one warp, the way Jellyfish does (k-mers run on over the line ends of a record)
    ml = b"@r\nACGTACGTAC\nACGTACGTAC\n+\nIIIIIIIIIIIIIIIIIIII\n@s\nACGTACGTAC\n+\nIIIIIIIIII\n"
    check_against_oracle(eng, [ok[0], ml, ok[4]], 7)
    check_against_oracle(eng, [ml], 9)


@pytest.mark.parametrize("k", [5, 7, 8, 11])
def test_fuzz_multiline_fastq(eng, k):
    from fuzzgen import rand_fastq_multiline
    rng = random.Random(5000 + k)
    bufs = [rand_fastq_multiline(rng) for _ in range(40)] + [rand_fastq(rng) for _ in range(10)] + [rand_fasta(rng) for _ in range(5)]
    rng.shuffle(bufs)
    counts, freq, totals, status = eng.count_buffers(bufs, k=k)
    assert (status == 0).all(), status
    for i, b in enumerate(bufs):
        ref = c_oracle.count_buffer(bytes(b), k) if len(b) else None
        assert np.array_equal(counts[i], ref), (k, i)


# ---- chunked-genome mode (get_chunks, main.py:654-929; BASELINE.json configs[4]) -------------------------------
def test_get_chunks_reproduces_reference_golden_rows(eng, toy_inputs, golden_dir, tmp_path):
    """All 358 rows of the reference's committed toy_example/train_tree_chunks/*.kf, text-exact (sha256 of every
    line); contig order is the input file's (the reference's is os.listdir order), order within a contig is checked."""
    import hashlib
    import json
    from kf2vecfsw_b200 import get_chunks
    gold = json.load(open(os.path.join(golden_dir, "chunks_golden.json")))
    ind, outd = tmp_path / "in", tmp_path / "out"
    ind.mkdir(); outd.mkdir()
    for s in gold:
        (ind / (s + ".fna")).write_bytes(toy_inputs[s])
    (ind / "tiny.fna").write_bytes(b">c1\n" + b"ACGT" * 2500 + b"\n")       # 1 chunk < 5: excluded (main.py:845)
    (ind / "short.fna").write_bytes(b">c1\n" + b"ACGT" * 100 + b"\n")       # no contig >= 10 kbp: excluded (main.py:761)
    get_chunks(argparse.Namespace(input_dir=str(ind), output_dir=str(outd), k=7, p=4, pseudocount=False))
    total = 0
    for s, rows in gold.items():
        lines = (outd / (s + ".kf")).read_bytes().splitlines(keepends=True)
        assert len(lines) == len(rows)
        gd = dict(rows)
        for ln in lines:
            label = ln.split(b",", 1)[0].decode()
            assert hashlib.sha256(ln).hexdigest() == gd[label], label
            total += 1
        def contig_order(labels):
            seen = {}
            for l in labels:
                seen.setdefault(l.split(".part_")[1], []).append(l)
            return seen
        ours, ref = contig_order([l.split(b",", 1)[0].decode() for l in lines]), contig_order([r[0] for r in rows])
        assert ours == ref
    assert total == 358
    assert not (outd / "tiny.kf").exists() and not (outd / "short.kf").exists()
    log = (outd / "get_chunks_in.log").read_text()
    assert "Excluded tiny.fna. 1 chunks is too low. 5 is required." in log
    assert "Excluded short.fna. No contigs above threshold length." in log
    assert sorted(os.listdir(outd)) == sorted([s + ".kf" for s in gold] + ["get_chunks_in.log"])   # no tmp dirs left


@pytest.mark.parametrize("k", [5, 7, 9])
def test_count_windows_vs_oracle(eng, k):
    rng = random.Random(k)
    seq = "".join(rng.choice("ACGT" * 30 + "Nn" + "acgt" * 3 + "R") for _ in range(60000)).encode()
    offs = [0, 1, 15, 16, 17, 511, 512, 513, 9999, 20000, 59000, 59990, 60000 - k, 0]
    lens = [10000, 10000, 10000, 10000, 10000, 10000, 10000, 10000, 10000, 3, 1000, 10, k, 60000]
    counts, freq, totals = eng.count_windows(np.frombuffer(seq, dtype=np.uint8), offs, lens, k=k)
    for i, (a, n) in enumerate(zip(offs, lens)):
        sym = o._CODE_LUT[np.frombuffer(seq[a:a + n], dtype=np.uint8)]
        ref = o.fold_canonical(o.forward_counts(sym, k), k)
        assert np.array_equal(counts[i], ref), i
        assert int(totals[i]) == int(ref.sum())


# ---- BASELINE.json configs[1] at FULL size -------------------------------------------------------------------------
def test_full_size_config_every_row_vs_c_oracle(eng):
    """1,000 synthetic genomes x 5 Mbp (the bench workload, 5.06 GB) through the device-arena path: every one of the
    1,000 rows bit-exact against the multi-threaded C oracle, plus size-independent properties (totals from the run
    structure of a sample, frequencies sum to one, a checksum of checksums over the whole matrix)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    G, NB = 1000, 5_000_000
    threads = len(os.sched_getaffinity(0))
    with ThreadPoolExecutor(threads) as ex:
        gen = list(ex.map(lambda i: kfsynth.synth_fasta(20261018, i, NB), range(G)))
    arena = eng.DeviceArena(gen)
    V = eng.vocab_size(7)
    counts = torch.empty((G, V), dtype=torch.int64, device="cuda")
    freq = torch.empty((G, V), dtype=torch.float64, device="cuda")
    totals = torch.empty(G, dtype=torch.int64, device="cuda")
    eng.count_device(arena, k=7, counts=counts, freq=freq, totals=totals)
    torch.cuda.synchronize()
    got = counts.cpu().numpy().astype(np.uint64)
    ref, _, st = c_oracle.count_buffers_mt(gen, 7, threads=threads, want_freq=False)
    assert np.array_equal(got, ref)
    assert np.array_equal(totals.cpu().numpy().astype(np.uint64), ref.sum(axis=1))
    # checksum of checksums: one number for the whole [1000, 8192] matrix
    w = (np.arange(V, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1)) & np.uint64(0xFFFFFFFF)
    assert int((got * w).sum(dtype=np.uint64)) == int((ref * w).sum(dtype=np.uint64))
    f = freq.cpu().numpy()
    assert np.all(np.abs(f.sum(axis=1) - 1.0) < 1e-12)
    assert np.array_equal(f, ref.astype(np.float64) / ref.sum(axis=1, keepdims=True).astype(np.float64))
    for i in (0, 499, 999):   # totals = sum over maximal ACGT runs of max(0, L - 6)
        sym = o.symbols_from_bytes(gen[i].tobytes())
        good = np.concatenate(([0], (sym >= 0).astype(np.int8), [0]))
        d = np.diff(good)
        runs = np.flatnonzero(d == -1) - np.flatnonzero(d == 1)
        assert int(totals[i]) == int(np.maximum(runs - 6, 0).sum())


def test_big_single_files_and_mixed_batch(eng):
    """One 150 Mbp single-contig genome (every line-kernel CTA holds a piece of the same file: 148 rows summed by the
    fold), one 60-column and one 70-column genome, an unwrapped one, FASTQ reads, an empty and an unsupported file in
    the same batch -- each row against the C oracle."""
    big = kfsynth.synth_fasta(7, 1, 150_000_000, max_contigs=1, n_runs=25)
    g60 = kfsynth.synth_fasta(7, 2, 20_000_000, line_width=60)
    g70 = kfsynth.synth_fasta(7, 3, 20_000_000, line_width=70)
    flat = kfsynth.synth_fasta(7, 4, 3_000_000, line_width=10 ** 8, max_contigs=5)
    fq = kfsynth.synth_fastq(7, 5, 1_000_000, 100_000, 150)
    bufs = [big, b"", g60, b"ACGT\n", g70, flat, fq]
    counts, freq, totals, status = eng.count_buffers(bufs, k=7)
    assert list(status) == [0, -9, 0, -5, 0, 0, 0]
    for i in (0, 2, 4, 5, 6):
        ref = c_oracle.count_buffer(bytes(bufs[i]) if not isinstance(bufs[i], np.ndarray) else bufs[i].tobytes(), 7)
        assert np.array_equal(counts[i], ref), i
        assert int(totals[i]) == int(ref.sum())
    assert not counts[1].any() and not counts[3].any()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [8, 9, 10])
def test_partitioned_kernel_large_k(eng, k):
    """k = 8..10: the partitioned shared-memory kernel ((file, partition) work items, u16 halves, plain read-add-write
    of the file's row) against the oracle and against the global-RED kernel; small and large files, FASTA and FASTQ in
    one batch; every file size through the partitioned kernel (part_all); a half that wraps (poly-A) recounted."""
    import random
    from fuzzgen import rand_fasta, rand_fasta_grid, rand_fastq
    rng = random.Random(1234 + k)
    bufs = [rand_fasta(rng), kfsynth.synth_fasta(7, 0, 700_000).tobytes(), rand_fasta_grid(rng), rand_fastq(rng),
            kfsynth.synth_fasta(7, 1, 3_000_000).tobytes(), b">only a header\n"]
    ref = [o.canonical_counts_bytes(bytes(b), k) for b in bufs]
    for kw in ({}, {"part_all": True}, {"no_linegrid": True}):
        counts, freq, totals, status = eng.count_buffers(bufs, k=k, **kw)
        for i in range(len(bufs)):
            assert np.array_equal(counts[i], ref[i]), (k, kw, i)
    seq = "A" * 5_000_000 + "ACGTTGCAAGGCTTAACCGGTTAA" * 500 + "N" * 50 + "C" * 1_600_001
    data = (">polyA\n" + "\n".join(seq[i:i + 80] for i in range(0, len(seq), 80)) + "\n").encode()
    r = c_oracle.count_buffer(data, k)
    assert int(r.max()) > 4_000_000
    counts, freq, totals, status = eng.count_buffers([data], k=k)
    assert np.array_equal(counts[0], r)


@pytest.mark.gpu
def test_files_to_kf_pipeline_matches_buffer_path(eng, toy_inputs, tmp_path):
    """kf_files_to_kf (reads / GPU / writes pipelined over batches) writes byte-for-byte what count_buffers + write_kf
    write, over several batches (tiny batch_bytes), FASTA and FASTQ, with unreadable / empty / non-sequence files
    reported per file and skipped; -raw_cnt rows switch to integers only when no k-mer is missing."""
    rng = random.Random(77)
    data = {"a": toy_inputs["G000830275sub"], "b": rand_fastq(rng), "c": toy_inputs["G000402355sub"], "d": rand_fasta_grid(rng),
            "e": b"", "f": b"not a sequence file\n", "g": kfsynth.synth_fasta(3, 0, 400_000).tobytes()}
    ind, outd = tmp_path / "in", tmp_path / "out"
    ind.mkdir(); outd.mkdir()
    names = sorted(data)
    paths = []
    for s in names:
        p = ind / (s + ".fa")
        p.write_bytes(data[s])
        paths.append(str(p))
    paths.append(str(ind / "missing.fa"))
    names.append("missing")
    for raw in (False, True):
        outs = [str(outd / ("%s_%d.kf" % (s, raw))) for s in names]
        status, totals, secs = eng.files_to_kf(paths, outs, names, k=7, raw_cnt=raw, threads=3, batch_bytes=600_000)
        ok = [s for s in names if s not in ("e", "f", "missing")]
        assert [int(status[names.index(s)]) for s in ok] == [0] * len(ok)
        assert all(int(status[names.index(s)]) != 0 for s in ("e", "f", "missing"))
        counts, freq, tot, st = eng.count_buffers([data[s] for s in ok], k=7, raw_cnt=raw)
        for j, s in enumerate(ok):
            ref = str(outd / "ref.kf")
            eng.write_kf(ref, s, freq[j], int_mode=bool(raw and np.all(counts[j] > 0)))
            assert open(outs[names.index(s)], "rb").read() == open(ref, "rb").read(), (s, raw)
            assert int(totals[names.index(s)]) == int(tot[j])
        for s in ("e", "f", "missing"):
            assert not os.path.exists(outs[names.index(s)])


@pytest.mark.gpu
@pytest.mark.parametrize("k,G", [(8, 160), (9, 160), (10, 96)])
def test_large_k_full_size_genomes_every_row_vs_c_oracle(eng, k, G):
    """BASELINE.json configs[4] shape: 5 Mbp genomes at k = 8 / 9 / 10 through the device-arena path (text pass + stream
    passes of the partitioned kernel, more work items than SMs): every row bit-exact against the multi-threaded C
    oracle, totals from the run structure, and the canonical fold property sum(counts) = number of valid k-mers."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    NB = 5_000_000
    threads = len(os.sched_getaffinity(0))
    with ThreadPoolExecutor(threads) as ex:
        gen = list(ex.map(lambda i: kfsynth.synth_fasta(777, i, NB), range(G)))
    arena = eng.DeviceArena(gen)
    V = eng.vocab_size(k)
    counts = torch.empty((G, V), dtype=torch.int64, device="cuda")
    totals = torch.empty(G, dtype=torch.int64, device="cuda")
    eng.count_device(arena, k=k, counts=counts, totals=totals)
    torch.cuda.synchronize()
    got = counts.cpu().numpy().astype(np.uint64)
    ref, _, st = c_oracle.count_buffers_mt(gen, k, threads=threads, want_freq=False)
    assert np.array_equal(got, ref)
    assert np.array_equal(totals.cpu().numpy().astype(np.uint64), ref.sum(axis=1))
    sym = o.symbols_from_bytes(gen[G - 1].tobytes())
    good = np.concatenate(([0], (sym >= 0).astype(np.int8), [0]))
    d = np.diff(good)
    runs = np.flatnonzero(d == -1) - np.flatnonzero(d == 1)
    assert int(totals[G - 1]) == int(np.maximum(runs - (k - 1), 0).sum())
    # a second call on the same layout (cached plan, reused stream and counters) gives the same matrix
    eng.count_device(arena, k=k, counts=counts, totals=totals)
    torch.cuda.synchronize()
    assert np.array_equal(counts.cpu().numpy().astype(np.uint64), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [9, 10])
def test_large_k_file_batching_over_a_small_workspace(eng, k, monkeypatch):
    """k >= 8 rows live in a bounded workspace (12 GB): larger batches are run in groups of files.  With the workspace
    squeezed to a few rows (KF_WS_LIMIT_BYTES) the groups hold 1-3 files: batch-global file ids in tiles and work
    items, rows relative to the group's first file, the decoded stream indexed by arena chunk."""
    rng = random.Random(99 + k)
    bufs = [kfsynth.synth_fasta(11, i, 400_000).tobytes() for i in range(4)] + [rand_fasta(rng), rand_fastq(rng), rand_fasta_grid(rng)] + \
           [kfsynth.synth_fasta(11, 9, 700_000).tobytes()]
    ref = [o.canonical_counts_bytes(bytes(b), k) for b in bufs]
    row_bytes = 4 << (2 * k)
    for nrows in (1, 3):
        monkeypatch.setenv("KF_WS_LIMIT_BYTES", str(nrows * row_bytes))
        counts, freq, totals, status = eng.count_buffers(bufs, k=k)
        for i in range(len(bufs)):
            assert np.array_equal(counts[i], ref[i]), (k, nrows, i)
    monkeypatch.delenv("KF_WS_LIMIT_BYTES")


@pytest.mark.gpu
def test_frequency_matrix_fast_path_equals_the_kf_round_trip(eng, toy_inputs, tmp_path):
    """frequency_matrix (files -> [N, V] float32 on the device through kf_files_to_device: pipelined reads, no text) holds
    exactly the tensor the trainers build from the .kf files: float32(float64 frequency * 1e4)
    (train_classifier_model.py:144-150,323), in os.listdir order, over several pipeline batches."""
    from kf2vecfsw_b200 import frequency_matrix, frequencies
    ind = tmp_path / "in"
    ind.mkdir()
    data = {"a": toy_inputs["G000830275sub"], "b": toy_inputs["G000402355sub"], "c": kfsynth.synth_fasta(5, 0, 500_000).tobytes(),
            "d": kfsynth.synth_fastq(5, 0, 100_000, 3_000, 150).tobytes()}
    for s, b in data.items():
        (ind / (s + (".fastq" if s == "d" else ".fna"))).write_bytes(b)
    old = frequencies.BATCH_BYTES
    frequencies.BATCH_BYTES = 700_000
    try:
        names, X = frequency_matrix(str(ind), k=7)
    finally:
        frequencies.BATCH_BYTES = old
    assert sorted(names) == sorted(data)
    counts, freq, totals, status = eng.count_buffers([data[n] for n in names], k=7)
    want = (freq * 1e4).astype(np.float32)
    assert X.dtype.__str__() == "torch.float32" and tuple(X.shape) == want.shape
    assert np.array_equal(X.cpu().numpy(), want)


# ---- round 2: k = 11 / 12 on the dense path, the sparse (sort-and-run-length) path up to k = 31, get_kmers ----
@pytest.mark.parametrize("k", [11, 12])
def test_dense_k11_k12_vs_c_oracle(eng, toy_inputs, k):
    """The ABI's dense rows at k = 11 and 12 (4^k global bins, 2,098,176 / 8,390,656 canonical columns): toy inputs and
    fuzz files against the C oracle's rolling canonical counter."""
    rng = random.Random(1100 + k)
    bufs = [toy_inputs["G000830275sub"], toy_inputs["G000402355sub"], rand_fasta(rng), rand_fasta_grid(rng), rand_fastq(rng)]
    counts, freq, totals, status = eng.count_buffers(bufs, k=k)
    assert (status == 0).all(), status
    for i, b in enumerate(bufs):
        ref = c_oracle.count_buffer(bytes(b), k)
        assert np.array_equal(counts[i], ref), (k, i)
        assert int(totals[i]) == int(ref.sum())
        if ref.sum() > 0:
            assert np.array_equal(freq[i], ref.astype(np.float64) / np.float64(ref.sum()))


def check_sparse(eng, bufs, k, **kw):
    codes, counts, row_off, totals, status = eng.sparse_count(bufs, k, **kw)
    assert row_off[0] == 0 and int(row_off[-1]) == codes.size == counts.size
    for i, b in enumerate(bufs):
        b = bytes(b)
        if len(b) == 0:
            assert status[i] == -9 and row_off[i] == row_off[i + 1]
            continue
        if b[:1] != b">":
            assert status[i] == (-10 if b[:1] == b"@" else -5) and row_off[i] == row_off[i + 1]
            continue
        assert status[i] == 0
        rc, rn, rt = c_oracle.count_sparse(b, k)
        a, e = int(row_off[i]), int(row_off[i + 1])
        assert int(totals[i]) == rt, (k, i)
        assert np.array_equal(codes[a:e], rc), (k, i, e - a, rc.size)
        assert np.array_equal(counts[a:e].astype(np.uint64), rn), (k, i)
    return codes, counts, row_off


@pytest.mark.parametrize("k", [6, 7, 8, 9, 10, 11, 12, 13, 15, 16, 17, 21, 31])
def test_sparse_counts_vs_c_oracle(eng, toy_inputs, k):
    """kf_sparse_count: observed canonical k-mers, ascending by code, against the C oracle (sort + run lengths of the
    rolling canonical mers): toy genomes, fuzz files (N runs, lower case, IUPAC, CRLF, blank lines, ragged widths),
    an empty buffer, a FASTQ buffer (KF_ERR_UNSUPPORTED) and a non-sequence buffer (KF_ERR_FORMAT)."""
    rng = random.Random(2000 + k)
    bufs = [toy_inputs["G000830275sub"], b"", toy_inputs["G000402355sub"], rand_fastq(rng), b"hello\n", toy_inputs["G000830295"]]
    bufs += [rand_fasta(rng) for _ in range(6)] + [rand_fasta_grid(rng) for _ in range(4)]
    check_sparse(eng, bufs, k)
    eng.sparse_release()


def test_sparse_equals_dense_at_k12_and_sub_batches(eng, toy_inputs, monkeypatch):
    """Two independent device paths at k = 12: the non-zero columns of the dense row (global REDs + fold) are the sparse
    path's entries; and several sub-batches (KF_SPARSE_BATCH_BYTES) give the same result as one."""
    names = ["G000830275sub", "G000402355sub", "G000830295"]
    bufs = [toy_inputs[s] for s in names]
    codes, counts, row_off = check_sparse(eng, bufs, 12)
    dense, _, _, status = eng.count_buffers(bufs, k=12, want_freq=False)
    vc = eng.vocab_codes(12).astype(np.uint64)
    for i in range(len(bufs)):
        nz = np.flatnonzero(dense[i])
        a, e = int(row_off[i]), int(row_off[i + 1])
        assert np.array_equal(vc[nz], codes[a:e]) and np.array_equal(dense[i][nz], counts[a:e].astype(np.uint64))
    monkeypatch.setenv("KF_SPARSE_BATCH_BYTES", "300000")
    c2, n2, r2, _, _ = eng.sparse_count(bufs, 12)
    assert len(eng.sparse_chunks()) >= 2
    assert np.array_equal(c2, codes) and np.array_equal(n2, counts) and np.array_equal(r2, row_off)
    eng.sparse_release()


def test_sparse_synthetic_genomes_k21_and_skewed_buckets(eng):
    """5 Mbp synthetic genomes at k = 21 (64-bit keys) and k = 12, plus a low-complexity file whose k-mers crowd a few
    buckets (poly-A / short repeats: one bucket far larger than shared memory -> the in-place global sort)."""
    import kfsynth
    bufs = [kfsynth.synth_fasta(7, i, 5_000_000) for i in range(3)]
    seq = "A" * 400000 + "ACGTTGCA" * 30000 + "N" * 7 + "AAAAAAAAAAAAC" * 20000
    bufs.append(np.frombuffer((">low\n" + "\n".join(seq[i:i + 70] for i in range(0, len(seq), 70)) + "\n").encode(), dtype=np.uint8))
    for k in (12, 10, 21):   # (k = 10, 12: the poly-A run overflows the packed 16-bit counters of its bucket -> exact u32 redo)
        check_sparse(eng, bufs, k)
    eng.sparse_release()


def test_get_kmers_end_to_end(eng, toy_inputs, tmp_path, capsys):
    """The FSW fork's get_kmers (main.py:112-184) on top of the sparse path: the .npy of every *.fna holds exactly the
    observed canonical k-mers as k base codes (A0 T1 C2 G3) + float32 normalised counts.  The reference lists rows in
    Jellyfish's hash order; the set of rows is what is defined (train_model_set.py:192-204 embeds it as a set)."""
    from kf2vecfsw_b200 import get_kmers
    ind, outd = tmp_path / "in", tmp_path / "out"
    ind.mkdir()
    names = ["G000830275sub", "G000402355sub"]
    for s in names:
        (ind / (s + ".fna")).write_bytes(toy_inputs[s])
    (ind / "noseq.fna").write_bytes(b">only_n\nNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN\n")
    (ind / "ignored.fa").write_bytes(toy_inputs[names[0]])          # get_kmers globs *.fna only (main.py:123)
    for k in (4, 7, 12, 21):
        get_kmers(argparse.Namespace(input_dir=str(ind), output_dir=str(outd), k=k))
        for s in names:
            got = np.load(outd / ("%s_k%d.npy" % (s, k)))
            ref = o.kmer_matrix(toy_inputs[s], k)
            assert got.dtype == np.float32 and got.shape == ref.shape == (ref.shape[0], k + 1)
            assert np.array_equal(got, ref), (s, k)                 # same row order here (ascending), so plain equality
            assert abs(float(got[:, k].sum()) - 1.0) < 1e-3
        assert not (outd / ("noseq_k%d.npy" % k)).exists() and not (outd / ("ignored_k%d.npy" % k)).exists()
    out = capsys.readouterr().out
    assert "--- Processing G000830275sub ---" in out and "Warning: No valid ATCG k-mers found in noseq" in out and "Saved: " in out


# ---- round 2: long-line ("unwrapped") FASTA through virtual lines ----
@pytest.mark.parametrize("seed0", [0, 100])
def test_fuzz_long_line_fasta(eng, seed0):
    """One line per contig (what assemblers write) and everything around it -- the virtual-line path of the line kernel
    against the oracle: 60 files per call, mixed with wrapped and generic files in one batch."""
    from fuzzgen import rand_fasta_long
    rng = random.Random(4000 + seed0)
    bufs = [rand_fasta_long(rng) for _ in range(60)] + [rand_fasta_grid(rng) for _ in range(4)] + [rand_fasta(rng) for _ in range(4)]
    rng.shuffle(bufs)
    check_against_oracle(eng, bufs, 7)


def test_unwrapped_genomes_vs_c_oracle(eng):
    """Unwrapped synthetic genomes: 1 contig (one 5 Mbp line: pieces that begin megabytes into a line) and up to 50 contigs,
    every row against the C oracle; and the same bases wrapped at 80 columns give the same counts."""
    flat1 = [kfsynth.synth_fasta(31, i, 5_000_000, line_width=10 ** 9, max_contigs=1) for i in range(12)]
    flat50 = [kfsynth.synth_fasta(32, i, 5_000_000, line_width=10 ** 9, max_contigs=50, n_runs=40) for i in range(12)]
    small = [kfsynth.synth_fasta(33, i, 60_000, line_width=10 ** 9, max_contigs=200) for i in range(8)]
    bufs = flat1 + flat50 + small
    counts, freq, totals, status = eng.count_buffers(bufs, k=7)
    assert (status == 0).all()
    ref, _, _ = c_oracle.count_buffers_mt(bufs, 7, threads=8, want_freq=False)
    assert np.array_equal(counts, ref)
    wrapped = [kfsynth.synth_fasta(31, i, 5_000_000, line_width=80, max_contigs=1) for i in range(3)]
    cw = eng.count_buffers(wrapped, k=7)[0]
    assert np.array_equal(cw, counts[:3])


def _peer_gather_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    from kf2vecfsw_b200.dist import PeerGather
    rows = [5, 3][:world] if world == 2 else [4] * world
    pg = PeerGather(rows, 64, torch.float32, dev)
    ok = True
    for step in range(5):                      # more steps than matrices: the release / reuse handshake is exercised
        pg.slot().fill_(float(100 * step + rank))
        pg.submit()
        full = pg.wait()
        torch.cuda.synchronize(dev)
        for r in range(world):
            ok = ok and bool((full[sum(rows[:r]):sum(rows[:r + 1])] == float(100 * step + r)).all())
        pg.release()
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_gather_two_gpus():
    """dist.PeerGather (copy-engine all-gather through symmetric memory, flag words + cuStreamWaitValue32) on two GPUs of one
    node; skipped on a single-GPU box (bench.py --gpus N uses it for every N > 1 and checks the gathered matrix there)."""
    import torch
    import torch.multiprocessing as mp
    import socket
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_fastq_pair_histogram_overflow_recount(eng):
    """k = 7 FASTQ counts 8-mer pairs in u16 halves; a CTA that issues more than 65,535 identical pairs for one file must
    notice (low-half checksum) and recount its tiles exactly."""
    rec = b"@r\n" + b"A" * 150 + b"\n+\n" + b"I" * 150 + b"\n"
    odd = b"@s\n" + b"A" * 70 + b"N" + b"ACGTACGTTTGCA" + b"\n+\n" + b"I" * 84 + b"\n@t\nACGTAC\n+\nIIIIII\n"
    data = rec * 40000 + odd + rec * 9000           # ~15 MB: several CTAs, each far beyond 65,535 identical pairs
    rng = random.Random(77)
    check_against_oracle(eng, [data, rand_fastq(rng), data[: len(rec) * 700]], 7)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [7, 9])
def test_count_buffers_sub_batch_pipeline(eng, toy_inputs, monkeypatch, k):
    """kf_count_buffers pipelines sub-batches of files (input copies of j + 1 beside the counting and the result copies
    of j).  A tiny sub-batch size makes every file or two its own sub-batch: results, totals and the per-file status
    (FASTQ layout reports are fetched per sub-batch) must not depend on the cut."""
    rng = random.Random(4242)
    ml = b"@r\nACGTACGTAC\nACGTACGTAC\n+\nIIIIIIIIIIIIIIIIIIII\n@s\nACGTACGTAC\n+\nIIIIIIIIII\n"   # multi-line: exact slow path
    bufs = [toy_inputs["G000830275sub"], rand_fastq(rng), b"", rand_fasta(rng), ml, b"hello\n", rand_fasta_grid(rng), rand_fastq(rng),
            toy_inputs["G000402355sub"], b"@r\nACGT\n+\nIII\n@s\nACGTACGTACGT\n+\nIIIIIIIIIIII\n" + b"x" * 600, rand_fasta(rng), rand_fastq(rng)]
    one = eng.count_buffers(bufs, k=k)
    monkeypatch.setenv("KF_SUB_BATCH_BYTES", "3000")
    cut = eng.count_buffers(bufs, k=k)
    monkeypatch.delenv("KF_SUB_BATCH_BYTES")
    assert np.array_equal(one[3], cut[3]), (one[3], cut[3])
    assert one[3][2] != 0 and one[3][5] != 0 and one[3][4] == 0   # empty, not a sequence file; multi-line FASTQ is counted (exact slow path)
    for i, b in enumerate(bufs):
        if one[3][i] != 0:
            continue
        ref = o.canonical_counts_bytes(bytes(b), k)
        assert np.array_equal(cut[0][i], ref), i
        assert np.array_equal(one[0][i], ref), i
        assert int(cut[2][i]) == int(ref.sum())
        assert np.array_equal(cut[1][i], one[1][i], equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [9, 12])
def test_sparse_bucket_paths_agree_and_device_kmer_matrix(eng, toy_inputs, monkeypatch, k):
    """k = 9 .. 12 have two independent device paths: buckets of 65,536 codes with 16-bit keys (write-combined partition, packed
    counters) and the 4,096-bucket path behind KF_SPARSE_NO16 -- same entries.  kf_sparse_kmer_matrix (the FSW rows expanded
    on the device) equals the NumPy expansion of the fetched entries bit for bit (main.py:147-169)."""
    from kf2vecfsw_b200.kmers import kmer_matrix_sparse
    rng = random.Random(77 + k)
    bufs = [toy_inputs["G000830275sub"], rand_fasta(rng), toy_inputs["G000402355sub"], rand_fasta_grid(rng), kfsynth.synth_fasta(3, 1, 1_200_000).tobytes()]
    a = eng.sparse_count(bufs, k)
    mats = []
    for i in range(len(bufs)):
        lo, hi = int(a[2][i]), int(a[2][i + 1])
        mats.append(eng.sparse_kmer_matrix(i, k, hi - lo, np.sum(a[1][lo:hi].astype(np.float32))))
    monkeypatch.setenv("KF_SPARSE_NO16", "1")
    b = eng.sparse_count(bufs, k)
    monkeypatch.delenv("KF_SPARSE_NO16")
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    for i in range(len(bufs)):
        lo, hi = int(a[2][i]), int(a[2][i + 1])
        ref = kmer_matrix_sparse(a[0][lo:hi], a[1][lo:hi], k)
        assert mats[i].dtype == np.float32 and mats[i].shape == ref.shape
        assert np.array_equal(mats[i], ref), i
    eng.sparse_release()
