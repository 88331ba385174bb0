set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -m pytest tests -x -q -m gpu 2>&1 | tail -30
