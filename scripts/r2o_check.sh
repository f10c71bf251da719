python scripts/norm_time.py > gpurun_out/r2o_norm_time.log 2>&1; cat gpurun_out/r2o_norm_time.log
timeout 300 python -m pytest tests/test_normalizer.py -m gpu -x -q 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:elg_reset_kernel -s 3 -c 2 -f -o gpurun_out/r2o_reset python scripts/config3_probe.py > gpurun_out/r2o_ncu_reset.log 2>&1; tail -3 gpurun_out/r2o_ncu_reset.log
