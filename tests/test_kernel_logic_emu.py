"""Kernel LOGIC on the CPU: the real device code of kf2vecfsw_b200/csrc/kf_kernels.cuh (decode, line-state
resolution, byte walker, tiling, flush, fold/normalise) compiled for the host against tests/emu/cuda_emu.h
(one OS thread per CUDA thread) and checked against the oracle.  This is not a product path -- the
product is the nvcc build of the same header -- it exists so kernel bugs are caught before GPU time."""
import os
import random
import subprocess

import numpy as np
import pytest

import kf_oracle as o
from fuzzgen import rand_fasta, rand_fasta_grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
BIN = os.path.join(EMU, "emu_main")


@pytest.fixture(scope="module", autouse=True)
def emu_binary():
    srcs = [os.path.join(EMU, "emu_main.cpp"), os.path.join(ROOT, "kf2vecfsw_b200", "csrc", "kf_host.cpp")]
    deps = srcs + [os.path.join(EMU, "cuda_emu.h"), os.path.join(ROOT, "kf2vecfsw_b200", "csrc", "kf_kernels.cuh")]
    if not os.path.exists(BIN) or any(os.path.getmtime(d) > os.path.getmtime(BIN) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-DKF_EMU", "-I", EMU, "-I", os.path.join(ROOT, "include"),
                               "-pthread"] + srcs + ["-o", BIN])


def run_emu(k, threads, grid, force_walker, tile_chunks, files, linegrid=True, mode=None):
    out = subprocess.run([BIN, str(k), str(threads), str(grid), str(int(force_walker)), str(tile_chunks),
                          str(int(linegrid) if mode is None else mode)] + files,
                         capture_output=True, text=True, check=True).stdout.strip().split("\n")
    res = []
    for i in range(len(files)):
        row = np.array(out[2 * i].split(), dtype=np.uint64)
        freq = np.array(out[2 * i + 1].split()[1:], dtype=np.float64)
        res.append((int(row[0]), row[1:], freq))
    return res


def test_emulated_kernels_reproduce_toy_goldens(toy_inputs, tmp_path):
    files = []
    for s in ("G000830275sub", "G000402355sub", "G000830295"):   # 1 record / 43 records / 100 N
        p = str(tmp_path / (s + ".fna"))
        open(p, "wb").write(toy_inputs[s])
        files.append((s, p))
    res = run_emu(7, 64, 5, False, 64, [p for _, p in files])
    for (s, _), (tot, counts, freq) in zip(files, res):
        ref = o.canonical_counts_bytes(toy_inputs[s], 7)
        vals, _ = o.row_values(ref, False, False)
        assert np.array_equal(counts, ref), s
        assert tot == int(ref.sum())
        assert np.array_equal(freq, vals), s   # bit-exact fp64


@pytest.mark.parametrize("seed0", [0, 1000])
def test_emulated_kernels_fuzz(seed0, tmp_path):
    for s in range(seed0, seed0 + 25):
        rng = random.Random(s)
        files = []
        for i in range(rng.randint(1, 3)):
            p = str(tmp_path / ("s%d_%d.fa" % (s, i)))
            open(p, "wb").write(rand_fasta(rng))
            files.append(p)
        k = rng.choice([3, 4, 5, 7])
        grid, thr, tile = rng.randint(1, 4), rng.choice([32, 64]), rng.choice([1, 2, 3, 5, 64])
        for fw in (False, True):
            res = run_emu(k, thr, grid, fw, tile, files)
            for f, (tot, counts, _) in zip(files, res):
                ref = o.canonical_counts_bytes(open(f, "rb").read(), k)
                assert np.array_equal(counts, ref), (s, fw, k, grid, thr, tile, f)


def test_emulated_linegrid_fuzz(tmp_path):
    """Fixed-width FASTA with real-file irregularities through the line-grid kernel (TMA staging emulated by memcpy)."""
    for s in range(300, 330):
        rng = random.Random(s)
        files = []
        for i in range(rng.randint(1, 3)):
            p = str(tmp_path / ("g%d_%d.fa" % (s, i)))
            open(p, "wb").write(rand_fasta_grid(rng))
            files.append(p)
        grid, thr, tile = rng.randint(1, 4), rng.choice([32, 64]), rng.choice([3, 8, 64])
        res = run_emu(7, thr, grid, False, tile, files, linegrid=True)
        for f, (tot, counts, _) in zip(files, res):
            ref = o.canonical_counts_bytes(open(f, "rb").read(), 7)
            assert np.array_equal(counts, ref), (s, grid, thr, tile, f)


def test_emulated_linegrid_u16_overflow_is_detected_and_recounted(tmp_path):
    """More than 65,535 identical 8-mer pairs in one flush interval wrap a 16-bit half of the pair histogram;
    the low-half checksum must catch it and the CTA must recount exactly."""
    seq = "A" * 300000 + "ACGTTGCAAGGCTTAACCGGTTAA" * 500 + "N" * 50 + "C" * 160001
    data = (">polyA\n" + "\n".join(seq[i:i + 80] for i in range(0, len(seq), 80)) + "\n").encode()
    p = str(tmp_path / "a.fa")
    open(p, "wb").write(data)
    ref = o.canonical_counts_bytes(data, 7)
    assert int(ref.max()) > 2 * 65535
    for grid, thr in ((1, 64), (3, 32)):
        tot, counts, _ = run_emu(7, thr, grid, False, 64, [p], linegrid=True)[0]
        assert np.array_equal(counts, ref)
    # the rare-path ("singles") histogram has 16-bit halves too: 2,500 lines that each hold an N send > 65,535
    # identical 7-mers through the byte walker
    data = (">polyA_with_N\n" + ("A" * 40 + "N" + "A" * 39 + "\n") * 2500).encode()
    open(p, "wb").write(data)
    ref = o.canonical_counts_bytes(data, 7)
    assert int(ref.max()) > 2 * 65535
    tot, counts, _ = run_emu(7, 64, 1, False, 64, [p], linegrid=True)[0]
    assert np.array_equal(counts, ref)


def test_emulated_fastq_fuzz(tmp_path):
    """4-line FASTQ through the FASTQ kernels (tile newline counts -> tile line types -> mask-based counting):
    N runs, lower case, qualities that start with '@' / '+' / '>', empty and shorter-than-k reads, files without
    a final newline, FASTA and FASTQ mixed in one batch, tiles of 1..64 chunks."""
    from fuzzgen import rand_fastq
    for s in range(500, 540):
        rng = random.Random(s)
        files = []
        for i in range(rng.randint(1, 3)):
            fq = rng.random() < 0.8
            p = str(tmp_path / ("q%d_%d.%s" % (s, i, "fq" if fq else "fa")))
            open(p, "wb").write(rand_fastq(rng) if fq else rand_fasta(rng))
            files.append(p)
        k = rng.choice([3, 4, 5, 7])
        grid, thr, tile = rng.randint(1, 4), rng.choice([32, 64]), rng.choice([1, 2, 3, 5, 64])
        res = run_emu(k, thr, grid, False, tile, files)
        for f, (tot, counts, _) in zip(files, res):
            ref = o.canonical_counts_bytes(open(f, "rb").read(), k)
            assert np.array_equal(counts, ref), (s, k, grid, thr, tile, f)


def test_emulated_more_ctas_than_chunks(toy_inputs, tmp_path):
    """A 14 KB file cut over 40 CTAs: most CTAs hold nothing, the others one 512-byte piece each -- every piece owns a
    row of the file (rank among the CTAs that hold the file, with gaps between them), the fold sums the rows."""
    files = []
    for s in ("G000830275sub", "G000402355sub"):
        p = str(tmp_path / (s + ".fna"))
        open(p, "wb").write(toy_inputs[s][:20000] if s != "G000830275sub" else toy_inputs[s])
        files.append(p)
    for grid in (40, 97):
        res = run_emu(7, 32, grid, False, 1, files)
        for f, (tot, counts, _) in zip(files, res):
            assert np.array_equal(counts, o.canonical_counts_bytes(open(f, "rb").read(), 7)), (grid, f)


def test_emulated_fastq_many_lanes(tmp_path):
    """FASTQ files of a few hundred KB: every lane of a warp gets its own 4 KiB range, has to find its first record
    with the local rule ('@' line whose second-next line starts with '+') and chases records by the quality length."""
    from fuzzgen import rand_fastq

    def piece(seed):
        b = rand_fastq(random.Random(seed))
        return b + b"\n" if b.count(b"\n") % 4 else b   # the generator drops the final newline at times

    files = []
    for j in range(3):
        p = str(tmp_path / ("big%d.fq" % j))
        open(p, "wb").write(b"".join(piece(1000 * j + i) for i in range(40)))
        files.append(p)
    for k, thr, grid, tile in ((7, 64, 2, 64), (5, 32, 3, 2), (3, 64, 4, 5)):
        res = run_emu(k, thr, grid, False, tile, files)
        for f, (tot, counts, _) in zip(files, res):
            assert np.array_equal(counts, o.canonical_counts_bytes(open(f, "rb").read(), k)), (k, thr, grid, tile, f)


def test_emulated_last_line_plus_padding_plus_next_header_is_one_slot(tmp_path):
    """Found by the full-size GPU test (1 k-mer in 5e9): a short last line (53 bases + '\\n'), 6 bytes of arena padding
    and the next file's 21-byte header add up to exactly one 81-byte slot whose byte 80 is a '\\n' -- the slot must not
    pass for a line on the grid, or the byte walker runs on into the next file."""
    rng = random.Random(835)
    n = next(n for n in range(200, 2000) if (57 + 81 * n) % 512 == 506)     # header 3 + n full lines + 54 -> 6 bytes of padding
    seq = "".join(rng.choice("ACGT") for _ in range(80 * n + 53))
    a = (">a\n" + "\n".join(seq[i:i + 80] for i in range(0, len(seq), 80)) + "\n").encode()
    assert len(a) % 512 == 506 and seq[-1] in "ACGT"
    b = (">g00836_c0 syntheti" + seq[-1].lower() + "\n" + "\n".join(seq[i:i + 80] for i in range(0, 8000, 80)) + "\n").encode()
    assert b.index(b"\n") == 20
    pa, pb = str(tmp_path / "a.fna"), str(tmp_path / "b.fna")
    open(pa, "wb").write(a)
    open(pb, "wb").write(b)
    for grid, thr in ((1, 64), (2, 32)):
        res = run_emu(7, thr, grid, False, 1024, [pa, pb])
        for f, (tot, counts, _) in zip((pa, pb), res):
            assert np.array_equal(counts, o.canonical_counts_bytes(open(f, "rb").read(), 7)), (grid, thr, f)
    # same file without its final newline: the file ends inside the slot
    open(pa, "wb").write(a[:-1])
    res = run_emu(7, 64, 1, False, 1024, [pa, pb])
    for f, (tot, counts, _) in zip((pa, pb), res):
        assert np.array_equal(counts, o.canonical_counts_bytes(open(f, "rb").read(), 7)), f


def test_emulated_partitioned_kernel_fuzz(tmp_path):
    """The k = 8..10 partitioned shared-memory kernel, instantiated small for the emulation (k = 4: 4 partitions of 64
    bins, k = 5: 16 partitions): (file, partition) items from a counter, u16 halves with the low-half checksum, drains
    every PART_FLUSH_TILES tiles, plain read-add-write of the row, exact recount after a wrapped half."""
    for s in range(800, 830):
        rng = random.Random(s)
        files = []
        for i in range(rng.randint(1, 3)):
            p = str(tmp_path / ("p%d_%d.fa" % (s, i)))
            open(p, "wb").write(rand_fasta(rng) if rng.random() < 0.5 else rand_fasta_grid(rng))
            files.append(p)
        k = rng.choice([3, 4, 5])   # one partition (the k = 8 shape) / text pass + stream passes over 3 / 15 partitions
        grid, thr, tile = rng.randint(1, 4), rng.choice([32, 64]), rng.choice([1, 3, 8, 64])
        res = run_emu(k, thr, grid, False, tile, files, mode=2)
        for f, (tot, counts, _) in zip(files, res):
            ref = o.canonical_counts_bytes(open(f, "rb").read(), k)
            assert np.array_equal(counts, ref), (s, k, grid, thr, tile, f)
    seq = "A" * 300000 + "ACGTTGCAAGGCTTAACCGGTTAA" * 500 + "N" * 50 + "C" * 160001 + "T" * 200000 + "ACGGT" * 1000 + "G" * 150000
    data = (">polyA\n" + "\n".join(seq[i:i + 80] for i in range(0, len(seq), 80)) + "\n").encode()
    p = str(tmp_path / "pa.fa")
    open(p, "wb").write(data)
    for k in (3, 4, 5):   # halves wrap in the text pass (A, C: gray codes 0, 1 ...) and in the stream passes
        ref = o.canonical_counts_bytes(data, k)
        assert int(ref.max()) > 2 * 65535
        for grid, thr, tile in ((1, 64, 64), (3, 32, 8)):
            tot, counts, _ = run_emu(k, thr, grid, False, tile, [p], mode=2)[0]
            assert np.array_equal(counts, ref), (k, grid, thr, tile)


def run_emu_sparse(k, threads, grid, tile_chunks, files):
    out = subprocess.run([BIN, str(k), str(threads), str(grid), "0", str(tile_chunks), "3"] + files,
                         capture_output=True, text=True, check=True).stdout.strip("\n").split("\n")
    res = []
    for i in range(len(files)):
        parts = out[2 * i].split()
        ent = [p.split(":") for p in parts[1:]]
        res.append((int(parts[0]), np.array([int(c) for c, _ in ent], dtype=np.uint64), np.array([int(n) for _, n in ent], dtype=np.uint64)))
    return res


def test_emulated_sparse_16bit_buckets_fuzz(tmp_path, monkeypatch):
    """k = 9 .. 12 on the sparse path: buckets of 65,536 codes, 16-bit keys, write-combined partition (per-warp slots,
    blocks of 8 from a run's front, single codes from its end), 65,536-bit map and two halves of 32,768 bins per
    (file, bucket).  A poly-A stretch overfills one bucket's slots within a chunk (single-code stores), counts above
    65,535 need the u32 bins."""
    for s, k in zip(range(950, 954), (12, 9, 11, 10)):
        rng = random.Random(s)
        files = []
        for i in range(rng.randint(1, 2)):
            p = str(tmp_path / ("w%d_%d.fa" % (s, i)))
            data = rand_fasta(rng) if rng.random() < 0.5 else rand_fasta_grid(rng)
            if s >= 953:
                seq = "A" * 70001 + "ACGTTGCAAGGCTTAACCGGTTAA" * 40 + "T" * 300
                data += (">poly\n" + "\n".join(seq[j:j + 80] for j in range(0, len(seq), 80)) + "\n").encode()
            open(p, "wb").write(data)
            files.append(p)
        grid, thr, tile = rng.randint(1, 3), rng.choice([32, 64]), rng.choice([1, 3, 64])
        res = run_emu_sparse(k, thr, grid, tile, files)
        for f, (tot, codes, counts) in zip(files, res):
            rc, rn, rt = o.sparse_counts_bytes(open(f, "rb").read(), k)
            assert tot == rt, (s, k, f)
            assert np.array_equal(codes, rc) and np.array_equal(counts, rn), (s, k, grid, thr, tile, f)
    # the 4,096-bucket path stays behind KF_SPARSE_NO16 (and serves k = 6 .. 8)
    monkeypatch.setenv("KF_SPARSE_NO16", "1")
    res = run_emu_sparse(12, 32, 2, 3, files[:1])
    rc, rn, rt = o.sparse_counts_bytes(open(files[0], "rb").read(), 12)
    assert res[0][0] == rt and np.array_equal(res[0][1], rc) and np.array_equal(res[0][2], rn)


def test_emulated_sparse_sort_rle_fuzz(tmp_path):
    """The sparse (sort-and-run-length) path for large k: two extraction passes (bucket histogram, bucket scatter),
    per-bucket bitonic sort, run-length emit -- against the NumPy oracle's observed canonical k-mers, k = 6 .. 31
    (32-bit keys up to k = 16 with the window fast path, 64-bit keys through the canonical byte walker above)."""
    for s, k in zip(range(900, 903), (12, 16, 31)):   # (histogram path / 32-bit sort / 64-bit sort + walker; ~1 min each emulated)
        rng = random.Random(s)
        files = []
        for i in range(1 if k == 12 else rng.randint(1, 2)):
            p = str(tmp_path / ("sp%d_%d.fa" % (s, i)))
            open(p, "wb").write(rand_fasta(rng) if rng.random() < 0.5 else rand_fasta_grid(rng))
            files.append(p)
        grid, thr, tile = rng.randint(1, 3), rng.choice([32, 64]), rng.choice([1, 3, 64])
        res = run_emu_sparse(k, thr, grid, tile, files)
        for f, (tot, codes, counts) in zip(files, res):
            rc, rn, rt = o.sparse_counts_bytes(open(f, "rb").read(), k)
            assert tot == rt, (s, k, f)
            assert np.array_equal(codes, rc) and np.array_equal(counts, rn), (s, k, grid, thr, tile, f)


BIN_RU = os.path.join(EMU, "emu_main_ru")


@pytest.fixture(scope="module")
def emu_binary_random_units(emu_binary):
    srcs = [os.path.join(EMU, "emu_main.cpp"), os.path.join(ROOT, "kf2vecfsw_b200", "csrc", "kf_host.cpp")]
    deps = srcs + [os.path.join(EMU, "cuda_emu.h"), os.path.join(ROOT, "kf2vecfsw_b200", "csrc", "kf_kernels.cuh")]
    if not os.path.exists(BIN_RU) or any(os.path.getmtime(d) > os.path.getmtime(BIN_RU) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-DKF_EMU", "-DKF_EMU_RANDOM_UNITS", "-I", EMU, "-I", os.path.join(ROOT, "include"),
                               "-pthread"] + srcs + ["-o", BIN_RU])


def run_emu_bin(binary, k, threads, grid, tile_chunks, files, seed=1):
    env = dict(os.environ, KF_EMU_SEED=str(seed))
    p = subprocess.run([binary, str(k), str(threads), str(grid), "0", str(tile_chunks), "1"] + files, capture_output=True, text=True, check=True, env=env,
                       timeout=120)
    out = p.stdout.strip().split("\n")
    return [(int(out[2 * i].split()[0]), np.array(out[2 * i].split()[1:], dtype=np.uint64)) for i in range(len(files))], p.stderr


def test_emulated_virtual_lines_fuzz(tmp_path, emu_binary_random_units):
    """Long-line FASTA through the virtual-line path of the line kernel: grid anywhere inside a line, unit states from
    the 256-byte look-back or assumed and verified at the end of the piece (headers longer than the look-back force the
    exact recount), breaks walked by the warp, pieces cut anywhere (tiles of 3 .. 64 chunks, 1 .. 4 CTAs).  Both the
    production unit sizes and random ones (1 .. 5 windows: many unit boundaries)."""
    from fuzzgen import rand_fasta_long
    n_virtual = n_files = 0
    for s in range(1200, 1236):
        rng = random.Random(s)
        files = []
        for i in range(rng.randint(1, 2)):
            p = str(tmp_path / ("v%d_%d.fa" % (s, i)))
            open(p, "wb").write(rand_fasta_long(rng))
            files.append(p)
        grid, thr, tile = rng.randint(1, 4), rng.choice([32, 64]), rng.choice([3, 8, 17, 64])
        res, err = run_emu_bin(BIN_RU if s % 2 else BIN, 7, thr, grid, tile, files, seed=s)
        n_virtual += int(err.split("virtual lines: ")[1].split(")")[0])      # files the probe took for long-line files
        n_files += len(files)
        for f, (tot, counts) in zip(files, res):
            ref = o.canonical_counts_bytes(open(f, "rb").read(), 7)
            assert np.array_equal(counts, ref), (s, grid, thr, tile, f, int(ref.sum()), int(counts.sum()))
    assert n_virtual >= 0.6 * n_files, (n_virtual, n_files)   # (a file whose first lines are all short goes to the generic kernel)


def test_emulated_multiline_fastq(tmp_path):
    """FASTQ out of 4-line layout: the record-chasing kernel reports it, the one-warp-per-file walk recounts it exactly."""
    from fuzzgen import rand_fastq, rand_fastq_multiline
    for s in range(1400, 1420):
        rng = random.Random(s)
        files = []
        for i in range(rng.randint(1, 3)):
            p = str(tmp_path / ("m%d_%d.fq" % (s, i)))
            open(p, "wb").write(rand_fastq_multiline(rng) if rng.random() < 0.8 else rand_fastq(rng))
            files.append(p)
        k = rng.choice([3, 4, 5, 7])
        res = run_emu(k, rng.choice([32, 64]), rng.randint(1, 3), False, rng.choice([1, 3, 64]), files)
        for f, (tot, counts, _) in zip(files, res):
            ref = o.canonical_counts_bytes(open(f, "rb").read(), k)
            assert np.array_equal(counts, ref), (s, k, f)


def test_emulated_fastq_pair_histogram_overflow_recount(tmp_path):
    """k = 7 FASTQ counts 8-mer pairs in u16 halves: more than 65,535 identical pairs between two flushes wrap a half, the
    low-half checksum must catch it and the CTA recount its tiles exactly (reads of poly-A with an N and a short read)."""
    rec = b"@r\n" + b"A" * 150 + b"\n+\n" + b"I" * 150 + b"\n"
    odd = b"@s\n" + b"A" * 70 + b"N" + b"ACGTACGTTTGCA" + b"\n+\n" + b"I" * 84 + b"\n@t\nACGTAC\n+\nIIIIII\n"
    data = rec * 1200 + odd + rec * 300
    p = str(tmp_path / "polyA.fq")
    open(p, "wb").write(data)
    ref = o.canonical_counts_bytes(data, 7)
    assert int(ref.max()) > 2 * 65535
    for grid, thr, tile in ((1, 64, 64), (2, 32, 64)):
        tot, counts, _ = run_emu(7, thr, grid, False, tile, [p])[0]
        assert np.array_equal(counts, ref), (grid, thr, tile)
