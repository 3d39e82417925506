#!/bin/bash
# One GPU-box session: parity tests, bench lines (own arm + reference arm), configs 3/4, ncu launch list + full captures.
# Usage (from the repo root, under gpurun): bash tools/gpu_session.sh
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>> gpurun_out/bench.err; echo "ref rc=$?"
python tools/run_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "cfg rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --import-source on --clock-control none -k regex:count_fasta_lines_kernel -s 3 -c 1 -f -o gpurun_out/prof_ln_bench \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc=$?"
ncu --set full --import-source on --clock-control none -k regex:count_fasta_part_kernel -s 1 -c 1 -f -o gpurun_out/prof_part_k9 \
    python tools/exp_largek.py 9 > gpurun_out/ncu3.log 2>&1; echo "ncu3 rc=$?"
tools/ubench 2048 > gpurun_out/ubench.log 2>&1; echo "ub rc=$?"
