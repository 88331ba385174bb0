# N3 visit: CNN tests (bounded), then timings.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cnn_gpu.py -q 2>&1 | grep -E "^(FAILED|ERROR)|passed|failed|^E +(assert|Assert)" | cut -c1-220 | head -60
