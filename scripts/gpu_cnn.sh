# N3 visit: CNN tests (bounded), then timings.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cnn_gpu.py -q 2>&1 | grep -E "^(FAILED|ERROR)|passed|failed|^E +(assert|Assert|Runtime)" | cut -c1-220 | head -60
timeout 300 python scripts/quick_cnn_bench.py 32 tf32 2>&1 | grep -v Warn | tee gpurun_out/cnn_bench.txt | tail -30
