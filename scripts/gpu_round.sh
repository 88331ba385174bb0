# One GPU visit: full GPU test-suite, smoke, reference arm, bench, ncu launch list of the bench's timed steps.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | grep -v -i warn | tail -3
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref_final.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench_final.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_all.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extractor > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out
