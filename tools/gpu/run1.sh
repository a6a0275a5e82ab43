set -x
mkdir -p gpurun_out/r2
python tools/quick_bench.py all > gpurun_out/r2/quick_all.txt 2>&1
B200FFT_FUSED=1 python tools/quick_bench.py 3 -d > gpurun_out/r2/quick_fused3.txt 2>&1
B200FFT_FUSED=1 B200FFT_BAND_MB=4 python tools/quick_bench.py 3 -d >> gpurun_out/r2/quick_fused3.txt 2>&1
B200FFT_FUSED=1 python tools/quick_bench.py 4 -d > gpurun_out/r2/quick_fused4.txt 2>&1
B200FFT_FUSED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused -c 1 -o gpurun_out/r2/fused3 python tools/ncu_one.py 2d f 1 8192 8192 > gpurun_out/r2/ncu_fused3.log 2>&1
ls -la gpurun_out/r2
