# Collect the round's evidence under gpurun_out/ (copied into profiles/ afterwards):
#  - launch list of the bench command (ncu gpu__time_duration, serialised, cold cache)
#  - one `ncu --set full` capture per heavy kernel
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 66 -c 44 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:logmel_kernel -s 1 -c 1 -o gpurun_out/full_logmel \
    python scripts/prof_logmel.py 2048 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fc_kernel -s 1 -c 1 -o gpurun_out/full_flanger \
    python scripts/prof_fc_bench.py flanger 4096 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fc_kernel -s 1 -c 1 -o gpurun_out/full_chorus \
    python scripts/prof_fc_bench.py chorus 4096 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:phaser -s 4 -c 4 -o gpurun_out/full_phaser \
    python scripts/prof_phaser.py 1365 > /dev/null 2>&1
ls -la gpurun_out
