# One GPU visit: full GPU test-suite, smoke, reference arm, bench, ncu launch list.
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -8
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 4000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
