set -x
timeout 900 python -m pytest tests/test_phaser_gpu.py tests/test_render_gpu.py tests/test_long_gpu.py tests/test_data_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:phaser --csv --log-file gpurun_out/ph_launches.csv python scripts/prof_phaser.py > /dev/null 2>&1; tail -4 gpurun_out/ph_launches.csv | awk -F'","' '{print $5, $NF}'
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ph3.json; python - <<'PY'
import json;d=json.load(open('gpurun_out/bench_ph3.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],{k:v['ms'] for k,v in d['roofline']['kernels'].items()},d['e2e']['value'])
PY
