#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), ncu launch list, one full ncu capture of the step kernel.
# usage: scripts/gpu_check.sh <tag>      outputs under gpurun_out/<tag>_*
set -u
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_gpu_tests.log
tail -5 $OUT/${TAG}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
cat $OUT/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench_ref.json
# launch list of the bench command, restricted to this library's kernels (the set-up alone is > 200 ATen fill / copy launches)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:elg_ -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 50 --warmup 3 --e2e-steps 5 --no-cpu-baseline --no-secondary > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:elg_step_fast -s 10 -c 3 -f -o $OUT/${TAG}_step \
    python bench.py --steps 20 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-secondary > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -20
