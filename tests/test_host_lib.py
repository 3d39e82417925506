"""CPU-side tests of libkfcount.so: it loads, exports every symbol include/kfcount.h declares, and its
host-only parts (vocabulary, Python-repr formatting, synthetic generators, error behaviour without a
GPU) match the oracle.  No compute entry point is exercised here."""
import ctypes
import math
import os
import random
import re
import struct

import numpy as np
import pytest

import kf_oracle as o
import kfsynth
from kf2vecfsw_b200 import build as kfbuild
from kf2vecfsw_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    kfbuild.build()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "kfcount.h")).read()
    names = sorted(set(re.findall(r"\b(kf_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 18
    L = ctypes.CDLL(engine.lib_path())
    for n in names:
        assert hasattr(L, n), n
    assert L.kf_abi_version() == 1


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = ctypes.CDLL(engine.lib_path())
    assert L.kf_init(0) == -2
    with pytest.raises(engine.KfError):
        engine.count_buffers([b">a\nACGT\n"], k=3)
    # the file pipeline and the SM limit refuse as well: nothing is read, counted or written without the device
    assert L.kf_set_sm_limit(100) == -2
    with pytest.raises(engine.KfError):
        engine.files_to_kf(["/nonexistent.fa"], ["/tmp/never_written.kf"], ["x"], k=7)
    assert not os.path.exists("/tmp/never_written.kf")


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
def test_vocab_equals_oracle(k):
    assert engine.vocab_size(k) == o.vocab_size(k)
    assert engine.vocab(k) == o.vocab(k)
    assert np.array_equal(engine.vocab_codes(k).astype(np.uint64), o.canonical_codes(k))


def test_format_row_is_python_repr():
    rng = random.Random(3)
    vals = [0.0, 1.0, 0.5, 1e-5, 1.5e-5, 9.999e-5, 1e-4, 0.0016335492570220753, 123456789.125, 1e15, 1e16,
            1.2345e20, 5e-324, 1.7976931348623157e308, 2.0 / 3.0, 1e22, 1e23, 100.0, 16.0, 0.1, float("nan")]
    for _ in range(3000):
        vals.append(struct.unpack("<d", struct.pack("<Q", rng.getrandbits(63)))[0])
        vals.append(rng.random() * 10 ** rng.randint(-12, 3))
        vals.append(rng.randint(0, 10 ** 6) / rng.randint(1, 10 ** 7))
    vals = [v for v in vals if not math.isinf(v)]
    row = np.array(vals, dtype=np.float64)
    want = "smp," + ",".join(repr(float(v)) for v in row) + "\n"
    assert engine.format_row("smp", row) == want
    ints = np.array([0, 1, 5, 123456789, 2 ** 40], dtype=np.float64)
    assert engine.format_row("s", ints, int_mode=True) == "s,0,1,5,123456789,1099511627776\n"


def test_format_row_reproduces_golden_kf(toy_inputs, toy_golden_kf):
    for s, data in toy_inputs.items():
        vals, int_mode = o.row_values(o.canonical_counts_bytes(data, 7), False, False)
        assert engine.format_row(s, vals, int_mode) == toy_golden_kf[s]


def test_write_kf(tmp_path):
    p = str(tmp_path / "a.kf")
    engine.write_kf(p, "a", np.array([0.25, 0.75]))
    engine.write_kf(p, "b", np.array([3.0, 4.0]), int_mode=True, append=True)
    assert open(p).read() == "a,0.25,0.75\nb,3,4\n"


def test_synth_fasta_is_deterministic_and_well_formed():
    a = kfsynth.synth_fasta(1234, 7, 200000)
    b = kfsynth.synth_fasta(1234, 7, 200000)
    c = kfsynth.synth_fasta(1234, 8, 200000)
    assert a.tobytes() == b.tobytes() and a.tobytes() != c.tobytes()
    assert a.size == kfsynth.synth_fasta_size(1234, 7, 200000)
    recs = o.fasta_records(a.tobytes())
    assert 1 <= len(recs) <= 50
    seq = b"".join(s for _, s in recs)
    assert len(seq) == 200000 and set(seq) <= set(b"ACGTN") and 1 <= seq.count(b"N") <= 1000
    assert all(len(l) <= 80 for l in a.tobytes().split(b"\n"))
    sym = o.symbols_from_bytes(a.tobytes())
    assert int((sym >= 0).sum()) == 200000 - seq.count(b"N")


def test_synth_fastq_is_four_line_and_hostile():
    fq = kfsynth.synth_fastq(99, 3, 50000, 2000, 150).tobytes()
    lines = fq.split(b"\n")
    assert lines[-1] == b"" and (len(lines) - 1) == 4 * 2000
    assert all(l.startswith(b"@g00003.") for l in lines[0:-1:4])
    assert all(l == b"+" for l in lines[2:-1:4])
    assert all(len(l) == 150 for l in lines[1:-1:4]) and all(len(l) == 150 for l in lines[3:-1:4])
    assert any(l[:1] in (b"@", b"+") for l in lines[3:-1:4])
    assert any(b"N" in l for l in lines[1:-1:4])


def test_list_inputs_naming(tmp_path):
    from kf2vecfsw_b200.frequencies import list_inputs
    for n in ("a.fna", "b.fa", "c.fasta", "d.fq", "e.fastq", "x.txt", "y.f.fna", "z.fna.gz"):
        (tmp_path / n).write_bytes(b">x\nA\n")
    files, samples = list_inputs(str(tmp_path))
    m = dict(zip(files, samples))
    assert set(files) == {"a.fna", "b.fa", "c.fasta", "d.fq", "e.fastq", "y.f.fna"}
    assert m["y.f.fna"] == "y.f" and m["c.fasta"] == "c" and m["e.fastq"] == "e"
    assert (files, samples) == tuple(map(list, zip(*o.list_inputs(str(tmp_path)))))


def test_parse_kf_round_trips_golden_rows_and_matches_trainer_tensor(toy_golden_kf):
    """The .kf reader (utils.py:436-437 and friends): values are exactly float(text) and the trainers' tensor
    float32(value * 1e4) equals what pandas + numpy build (train_classifier_model.py:144-150)."""
    import io
    import pandas as pd
    text = "".join(toy_golden_kf[s] for s in sorted(toy_golden_kf)).encode()
    labels, rows, feat = engine.parse_kf(text, 8192, want_feat=True)
    assert labels == sorted(toy_golden_kf)
    for i, s in enumerate(labels):
        exact = np.array([float(x) for x in toy_golden_kf[s].strip().split(",")[1:]])
        assert np.array_equal(rows[i], exact)
        assert engine.format_row(s, rows[i]) == toy_golden_kf[s]          # writer(reader(x)) == x
    df = pd.read_csv(io.BytesIO(text), index_col=0, header=None, sep=",")   # the reference's reader
    assert list(df.index) == labels
    assert np.array_equal(feat, (df.values * 1e4).astype(np.float32))
    assert np.allclose(rows, df.values, rtol=1e-11, atol=0)   # pandas xstrtod drops digits (1e-12 relative); gone after float32


def test_parse_kf_formats_and_errors():
    labels, rows, _ = engine.parse_kf(b"a,1,2,3\nb,1.0,0.0,nan\r\n\nc.part_x,5e-05,1e+16,0.5", 3)
    assert labels == ["a", "b", "c.part_x"]
    assert np.array_equal(rows[0], [1, 2, 3]) and np.isnan(rows[1][2]) and rows[2][0] == 5e-05 and rows[2][1] == 1e16
    for bad in (b"a,1,2\n", b"a,1,2,3,4\n", b"a,1,x,3\n", b"a\n"):
        with pytest.raises(engine.KfError):
            engine.parse_kf(bad, 3)
    assert engine.parse_kf(b"", 3)[0] == []


def test_loader_and_chunk_readers(tmp_path, toy_golden_kf):
    from kf2vecfsw_b200 import loader
    for s, t in toy_golden_kf.items():
        (tmp_path / (s + ".kf")).write_text(t)
    (tmp_path / "chunks.kf").write_text("g.part_c.part_c_sliding__1-10000,16.0,0.0,300.0\ng.part_c.part_c_sliding__5-10004,1.0,255.0,256.0\n")
    labels, rows = loader.read_kf(str(tmp_path / "G000830275sub.kf"))
    assert labels == ["G000830275sub"] and rows.shape == (1, 8192)
    names, feat = loader.load_kf_files([str(tmp_path / (s + ".kf")) for s in sorted(toy_golden_kf)])
    assert names == sorted(toy_golden_kf) and tuple(feat.shape) == (7, 8192) and str(feat.dtype) == "torch.float32"
    assert abs(float(feat[0].double().sum()) - 1e4) < 1e-2
    lab, u16 = loader.read_chunk_kf(str(tmp_path / "chunks.kf"))
    assert u16.dtype == np.uint16 and u16.tolist() == [[16, 0, 300], [1, 255, 256]]
    lab, u8 = loader.read_chunk_kf(str(tmp_path / "chunks.kf"), cap_uint8=True)
    assert u8.dtype == np.uint8 and u8.tolist() == [[16, 0, 255], [1, 255, 255]]       # utils.py:416-431


def test_kmer_matrix_is_the_fsw_feature_layout():
    """get_kmers (main.py:112-184): observed canonical k-mers as base codes A0 T1 C2 G3 + float32 normalised count."""
    from kf2vecfsw_b200 import kmers
    k = 3
    counts = o.canonical_counts_bytes(b">a\nACGTTGCAAT\n", k)
    m = kmers.kmer_matrix(counts, k)
    assert m.dtype == np.float32 and m.shape == (int((counts > 0).sum()), k + 1)
    code = {"A": 0.0, "T": 1.0, "C": 2.0, "G": 3.0}
    words = [w for w, c in zip(o.vocab(k), counts) if c]
    assert [[code[ch] for ch in w] for w in words] == m[:, :k].tolist()
    c32 = counts[counts > 0].astype(np.float32)
    assert np.array_equal(m[:, k], c32 / np.sum(c32))
