set -u
OUT=gpurun_out
TAG=${1:-r2l}
timeout 900 python -m pytest tests/test_normalizer.py tests/test_fused_reset.py tests/test_step_parity.py tests/test_rollout_step.py tests/test_multi_gpu.py -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; tail -5 $OUT/${TAG}_tests.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err
python - $TAG <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/{sys.argv[1]}_bench.json').read().strip().splitlines()[-1])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['value'])
print('norm', d['secondary']['obs_normalize_store'])
print('config3', d['secondary']['config3_strong']['us_per_step'])
print('value', d['value'], d['ms_per_step'], d['roofline']['frac'])
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:elg_ -c 60 --csv --log-file $OUT/${TAG}_config3_launches.csv python scripts/config3_probe.py > $OUT/${TAG}_config3_probe.log 2>&1; tail -2 $OUT/${TAG}_config3_probe.log
