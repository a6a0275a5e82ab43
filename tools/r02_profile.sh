#!/bin/bash
# round-2 profile set: launch list of the default bench command + full captures of the kernels that carry the configs
mkdir -p gpurun_out/r2
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_ring_rows -s 2 -c 1 -o gpurun_out/r2/ncu_cfg2_ring -f python tools/quick_bench.py 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_band -s 2 -c 1 -o gpurun_out/r2/ncu_cfg3_band -f python tools/quick_bench.py 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_ring_rows -s 2 -c 1 -o gpurun_out/r2/ncu_cfg3_rows -f python tools/quick_bench.py 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_lines -s 4 -c 1 -o gpurun_out/r2/ncu_cfg1_lines -f python tools/quick_bench.py 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_ringcol -s 1 -c 1 -o gpurun_out/r2/ncu_cfg5_ringcol -f python tools/quick_bench.py 5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_ringcol -s 1 -c 1 -o gpurun_out/r2/ncu_cfg4_ringcol_tw -f python tools/quick_bench.py 4 > /dev/null 2>&1
ls -la gpurun_out/r2/*.ncu-rep
