#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference checkout's toy_example (run in the build container,
where /root/reference exists).  The fixtures are the reference's own inputs and committed outputs:

  tests/golden/fna/<sample>.fna.xz      toy_example/{train_tree_fna,test_fna}/*.fna (inputs, xz)
  tests/golden/kf/<sample>.kf.gz        toy_example/{train_tree_kf,test_kf}/<sample>.kf (golden rows)
  tests/golden/chunks_golden.json       per-row (label, sha256 of the full text line) of
                                        toy_example/train_tree_chunks/*.kf (358 rows)
  tests/golden/vocab_sha256.json        sha256 of kf2vec/data/<vocabulary file> for k = 3..9

Nothing here is produced by the oracle or the product: these pin both.
"""
import glob
import gzip
import hashlib
import json
import lzma
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

os.makedirs(os.path.join(OUT, "fna"), exist_ok=True)
os.makedirs(os.path.join(OUT, "kf"), exist_ok=True)

manifest = {}
for d_in, d_out in (("train_tree_fna", "train_tree_kf"), ("test_fna", "test_kf")):
    for f in sorted(glob.glob(os.path.join(REF, "toy_example", d_in, "*.fna"))):
        sample = os.path.basename(f).rsplit(".f", 1)[0]
        data = open(f, "rb").read()
        gold = open(os.path.join(REF, "toy_example", d_out, sample + ".kf"), "rb").read()
        with open(os.path.join(OUT, "fna", sample + ".fna.xz"), "wb") as fh:
            fh.write(lzma.compress(data, preset=9))
        with gzip.GzipFile(os.path.join(OUT, "kf", sample + ".kf.gz"), "wb", mtime=0) as fh:
            fh.write(gold)
        manifest[sample] = {"source": d_in, "bytes": len(data), "fna_sha256": hashlib.sha256(data).hexdigest(),
                            "kf_sha256": hashlib.sha256(gold).hexdigest()}
json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1, sort_keys=True)

chunks = {}
for f in sorted(glob.glob(os.path.join(REF, "toy_example", "train_tree_chunks", "*.kf"))):
    sample = os.path.basename(f)[:-3]
    rows = []
    for line in open(f, "rb"):
        rows.append([line.split(b",", 1)[0].decode(), hashlib.sha256(line).hexdigest()])
    chunks[sample] = rows
json.dump(chunks, open(os.path.join(OUT, "chunks_golden.json"), "w"))

vocab = {}
for k, name in ((3, "vocab_generator_k3C_fin.fa"), (4, "vocab_generator_k4C_fin.fa"), (5, "vocab_generator_k5C_fin.fa"),
                (6, "test_kmers_6_sorted"), (7, "test_kmers_7_sorted"), (8, "vocab_generator_k8C_fin.fa"),
                (9, "vocab_generator_k9C_fin.fa")):
    vocab[str(k)] = hashlib.sha256(open(os.path.join(REF, "kf2vec", "data", name), "rb").read()).hexdigest()
json.dump(vocab, open(os.path.join(OUT, "vocab_sha256.json"), "w"), indent=1, sort_keys=True)
print("wrote", OUT)
