set -u
OUT=gpurun_out
timeout 600 python -m pytest tests/test_normalizer.py -m gpu -x -q > $OUT/r2k_norm_tests.log 2>&1; tail -5 $OUT/r2k_norm_tests.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/r2k_bench.json 2> $OUT/r2k_bench.err; echo "bench rc=$?"; tail -3 $OUT/r2k_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['value'])
print('norm', d['secondary']['obs_normalize_store'])
print('value', d['value'], d['roofline']['frac'])
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:elg_ -c 60 --csv --log-file $OUT/r2k_config3_launches.csv python scripts/config3_probe.py > $OUT/r2k_config3_probe.log 2>&1; tail -2 $OUT/r2k_config3_probe.log
