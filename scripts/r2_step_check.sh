#!/bin/bash
# step-kernel iteration pass: parity tests of the step, in-kernel stamps, A/B timing (this build vs the round-1 library), launch floor
set -u
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests/test_step_parity.py tests/test_fused_reset.py tests/test_actuator_net.py tests/test_rollout_clone.py -m gpu -x -q > $OUT/${TAG}_step_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_step_tests.log
tail -15 $OUT/${TAG}_step_tests.log
timeout 300 python scripts/step_stamps.py > $OUT/${TAG}_stamps.txt 2>&1; cat $OUT/${TAG}_stamps.txt
timeout 600 python scripts/step_ab.py ${2:-quick} > $OUT/${TAG}_ab.txt 2>&1; cat $OUT/${TAG}_ab.txt
if [ -f scripts/ab/libelg_v5.so ]; then
  echo "--- round-1 library (v5) on the same box"
  ELG_LIB_PATH=scripts/ab/libelg_v5.so timeout 600 python scripts/step_ab.py ${2:-quick} > $OUT/${TAG}_ab_v5.txt 2>&1; cat $OUT/${TAG}_ab_v5.txt
fi
