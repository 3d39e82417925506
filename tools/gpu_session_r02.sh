#!/bin/bash
# One GPU-box session (round 2): parity tests, bench lines (own arm + reference arm), line widths, sparse path, files -> .kf,
# ncu launch list + full captures.  Usage (from the repo root, under gpurun): bash tools/gpu_session_r02.sh
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>> gpurun_out/bench.err; echo "ref rc=$?"
python tools/exp_widths.py 1000 80 70 60 50 100 120 0 > gpurun_out/widths.jsonl 2> gpurun_out/widths.err; echo "widths rc=$?"
KF_CONTIGS=1 python tools/exp_widths.py 1000 0 >> gpurun_out/widths.jsonl 2>> gpurun_out/widths.err
python tools/exp_sparse.py 296 12 11 10 9 > gpurun_out/sparse.jsonl 2> gpurun_out/sparse.err; python tools/exp_sparse.py 64 15 21 31 >> gpurun_out/sparse.jsonl 2>> gpurun_out/sparse.err; echo "sparse rc=$?"
python tools/exp_files.py 1000 > gpurun_out/files.jsonl 2> gpurun_out/files.err; echo "files rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/ncu.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --import-source on --clock-control none -k regex:count_fasta_lines_kernel -s 1 -c 1 -f -o gpurun_out/prof_ln_bench \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc=$?"
ncu --set full --import-source on --clock-control none -k regex:count_fasta_lines_kernel -s 2 -c 1 -f -o gpurun_out/prof_vl \
    python tools/exp_widths.py 1000 0 > gpurun_out/ncu3.log 2>&1; echo "ncu3 rc=$?"
python tools/ncu_summary.py gpurun_out/prof_vl.ncu-rep > gpurun_out/ncu_vl.txt
ncu --set full --import-source on --clock-control none -k regex:sparse -s 6 -c 6 -f -o gpurun_out/prof_sparse \
    python tools/exp_sparse.py 148 12 > gpurun_out/ncu4.log 2>&1; echo "ncu4 rc=$?"
python tools/ncu_summary.py gpurun_out/prof_sparse.ncu-rep > gpurun_out/ncu_sparse.txt
tools/ubench 2048 > gpurun_out/ubench.log 2>&1; echo "ub rc=$?"
