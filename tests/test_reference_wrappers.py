"""The reference's own wrappers on top of the replacement: kf2vec/main.py is loaded UNCHANGED from /root/reference (its
tree / FSW dependencies, which are not in this image, stubbed in sys.modules), ``get_frequencies`` is rebound to this
package's, and ``process_query_data`` (main.py:629-655) and ``build_library`` (main.py:569-626) are run with the Namespace
their argparse sub-commands build (main.py:1253-1353: no ``raw_cnt`` attribute).  The stages behind the frequency step
(classify / query / divide_tree / get_distances / trainers) are replaced by recorders: what is checked is that the
reference code drives the replacement through its own call, that the .kf files it leaves are the reference's committed
golden files byte for byte, and that the next stage is pointed at them.

/root/reference exists only in the build container, which has no GPU, so the counting itself is served here by the
oracle behind ``engine.files_to_kf`` (test infrastructure; the text rows are written by the product's host code);
tests/test_gpu_parity.py::test_get_frequencies_reproduces_reference_golden_kf runs the same call on the CUDA path."""
import argparse
import os
import sys
import types

import numpy as np
import pytest

import kf_oracle as o

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "kf2vec", "main.py")), reason="reference tree not present on this box")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None})


@pytest.fixture(scope="module")
def ref_main():
    for name in ("treeswift", "fswlib", "treecluster"):
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    sys.path.insert(0, REF)
    try:
        import kf2vec.main as M
    finally:
        sys.path.remove(REF)
    return M


@pytest.fixture()
def oracle_backed_engine(monkeypatch):
    """engine.files_to_kf with the counting served by the oracle (no GPU in this container); rows formatted and written by
    the library's own host code (kf_write_kf)."""
    from kf2vecfsw_b200 import engine

    def files_to_kf(in_paths, out_paths, samples, k=7, pseudocount=False, raw_cnt=False, threads=0, batch_bytes=0):
        status = np.zeros(len(in_paths), dtype=np.int32)
        totals = np.zeros(len(in_paths), dtype=np.uint64)
        for i, (p, q, s) in enumerate(zip(in_paths, out_paths, samples)):
            counts = o.canonical_counts_bytes(open(p, "rb").read(), k)
            vals, int_mode = o.row_values(counts, pseudocount, raw_cnt)
            engine.write_kf(q, s, vals, int_mode=int_mode)
            totals[i] = counts.sum()
        return status, totals, np.zeros(4)

    monkeypatch.setattr(engine, "files_to_kf", files_to_kf)
    return engine


def _dirs(tmp_path, toy_inputs, names):
    ind, outd = tmp_path / "in", tmp_path / "out"
    ind.mkdir()
    outd.mkdir()
    for s in names:
        (ind / (s + ".fna")).write_bytes(toy_inputs[s])
    return str(ind), str(outd)


def test_reference_process_query_data_runs_on_the_replacement(ref_main, oracle_backed_engine, toy_inputs, toy_golden_kf, tmp_path, monkeypatch, capsys):
    import kf2vecfsw_b200
    names = ["G000830275sub", "G000402355sub"]
    ind, outd = _dirs(tmp_path, toy_inputs, names)
    seen = {}
    monkeypatch.setattr(ref_main, "get_frequencies", kf2vecfsw_b200.get_frequencies)            # the drop-in
    monkeypatch.setattr(ref_main, "classify", lambda a: seen.setdefault("classify", (a.input_dir, a.model, a.o, sorted(os.listdir(a.input_dir)))))
    monkeypatch.setattr(ref_main, "query", lambda a: seen.setdefault("query", (a.input_dir, a.model, a.classes)))
    # the Namespace of the `process_query_data` sub-command (main.py:1323-1353)
    args = argparse.Namespace(input_dir=ind, output_dir=outd, k=7, p=4, pseudocount=False, classifier_model="cl.ckpt",
                              distance_model="di_models", cl_seed=28, di_seed=28)
    ref_main.process_query_data(args)
    for s in names:
        assert open(os.path.join(outd, s + ".kf")).read() == toy_golden_kf[s], s
    assert seen["classify"] == (outd, "cl.ckpt", outd, sorted(s + ".kf" for s in names))
    assert seen["query"] == (outd, "di_models", outd)
    out = capsys.readouterr().out
    assert "==> Computing k-mer frequences" in out and "==> Starting k-mer counting for" in out and "==> Query processing step is completed!" in out


def test_reference_build_library_runs_on_the_replacement(ref_main, oracle_backed_engine, toy_inputs, toy_golden_kf, tmp_path, monkeypatch):
    import kf2vecfsw_b200
    names = ["G000830275sub", "G000830295"]
    ind, outd = _dirs(tmp_path, toy_inputs, names)
    order = []
    monkeypatch.setattr(ref_main, "get_frequencies", kf2vecfsw_b200.get_frequencies)
    monkeypatch.setattr(ref_main, "divide_tree", lambda a: order.append(("divide_tree", sorted(os.listdir(a.output_dir)))))
    monkeypatch.setattr(ref_main, "get_distances", lambda a: order.append(("get_distances", a.subtrees)))
    monkeypatch.setattr(ref_main, "train_classifier", lambda a: order.append(("train_classifier", a.input_dir, a.e)))
    monkeypatch.setattr(ref_main, "train_model_set", lambda a: order.append(("train_model_set", a.input_dir, a.true_dist)))
    tree = str(tmp_path / "tree" / "backbone.nwk")
    args = argparse.Namespace(input_dir=ind, output_dir=outd, k=7, p=2, pseudocount=False, tree=tree, cl_epochs=3, cl_hidden_sz=8, cl_batch_sz=2,
                              cl_lr=1e-5, cl_lr_min=3e-6, cl_lr_decay=10, cl_seed=1, di_epochs=4, di_hidden_sz=8, di_embed_sz=4, di_batch_sz=2,
                              di_lr=1e-5, di_lr_min=3e-6, di_lr_decay=10, di_seed=1)
    ref_main.build_library(args)
    for s in names:
        assert open(os.path.join(outd, s + ".kf")).read() == toy_golden_kf[s], s
    assert order[0] == ("divide_tree", sorted(s + ".kf" for s in names))                  # the frequency step ran first and left only .kf files
    assert order[1] == ("get_distances", os.path.join(str(tmp_path / "tree"), "backbone.subtrees"))
    assert order[2] == ("train_classifier", outd, 3) and order[3] == ("train_model_set", outd, str(tmp_path / "tree"))


def test_reference_loader_reads_what_the_replacement_writes(ref_main, oracle_backed_engine, toy_inputs, tmp_path):
    """The trainers' own reader (train_classifier_model.py:144-150: pandas, header=None, index_col=0, * 1e4 -> float32)
    on the files get_frequencies wrote, against this package's loader."""
    import pandas as pd
    import kf2vecfsw_b200
    names = ["G000830275sub", "G000402355sub"]
    ind, outd = _dirs(tmp_path, toy_inputs, names)
    kf2vecfsw_b200.get_frequencies(argparse.Namespace(input_dir=ind, output_dir=outd, k=7, p=1, pseudocount=False))
    frames = [pd.read_csv(os.path.join(outd, s + ".kf"), header=None, index_col=0) for s in names]
    ref = (pd.concat(frames).to_numpy(dtype=np.float64) * 1e4).astype(np.float32)
    text = b"".join(open(os.path.join(outd, s + ".kf"), "rb").read() for s in names)
    got_labels, rows, got = kf2vecfsw_b200.engine.parse_kf(text, 8192, want_rows=True, want_feat=True)
    assert got_labels == names
    assert np.array_equal(got, ref)
