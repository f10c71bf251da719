python scripts/norm_time.py > gpurun_out/r2p_norm_time.log 2>&1; cat gpurun_out/r2p_norm_time.log
timeout 600 python -m pytest tests/test_normalizer.py tests/test_fused_reset.py tests/test_rollout_step.py tests/test_rollout_clone.py tests/test_sensor_envs.py -m gpu -x -q 2>&1 | tail -3
python scripts/config3_probe.py 2>&1 | tail -7
