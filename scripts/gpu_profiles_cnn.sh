# ncu evidence for the extractor (N3): launch list of one forward at B = 32, full captures of both tcgen05 kernels.
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 20 --csv --log-file gpurun_out/launches_extractor.csv \
    python scripts/prof_extractor.py 32 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tf32 -s 1 -c 1 -o gpurun_out/full_conv_tf32 \
    python scripts/prof_cnn.py 32 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1_tf32 -s 1 -c 1 -o gpurun_out/full_conv1_tf32 \
    python scripts/prof_extractor.py 32 > /dev/null 2>&1
timeout 300 python scripts/quick_cnn_bench.py 128 tf32 2>&1 | grep -v Warn > gpurun_out/cnn_bench_b128.txt
timeout 300 python scripts/quick_cnn_bench.py 32 2>&1 | grep -v Warn > gpurun_out/cnn_bench_b32.txt
ls -la gpurun_out
