# Round-2 evidence, collected under gpurun_out/ (summaries are copied into profiles/ afterwards):
#  full GPU test-suite and smoke(), the bench line of every BASELINE config and the reference arm, the ncu launch list of
#  the timed steps of the default bench, one `ncu --set full` capture per heavy kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/r02_bench.json
for c in 1 2 3 5; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r02_bench_config$c.json 2> gpurun_out/bench_c$c.err; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_all.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logmel_kernel -s 2 -c 1 -o gpurun_out/r02_full_logmel \
    python scripts/prof_logmel.py 2048 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fc_cta_kernel -s 1 -c 1 -o gpurun_out/r02_full_flanger \
    python scripts/prof_fc_bench.py flanger 4096 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fc_wide_kernel -s 1 -c 1 -o gpurun_out/r02_full_chorus \
    python scripts/prof_fc_bench.py chorus 4096 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:phaser_fused -s 2 -c 1 -o gpurun_out/r02_full_phaser \
    python scripts/prof_phaser.py 1365 > /dev/null 2>&1
ls -la gpurun_out | tail -20
