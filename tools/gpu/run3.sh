set -x
mkdir -p gpurun_out/r2
timeout 300 python tools/gpu/band_check.py > gpurun_out/r2/band_check.txt 2>&1; tail -12 gpurun_out/r2/band_check.txt
timeout 120 python tools/quick_bench.py 3 -d > gpurun_out/r2/band_cfg3.txt 2>&1; cat gpurun_out/r2/band_cfg3.txt
