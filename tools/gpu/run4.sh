set -x
mkdir -p gpurun_out/r2
rm -f gpurun_out/r2/band_sweep.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:band -c 1 -f -o gpurun_out/r2/band3 python tools/ncu_one.py 2d f 1 8192 8192 > gpurun_out/r2/ncu_band3.log 2>&1
for la in 4 8 12; do for sl in 4 8; do
B200FFT_BAND_LA=$la B200FFT_BAND_SLOTS=$((la+sl)) python tools/quick_bench.py 3 | sed "s/^/la=$la slots=+$sl /" >> gpurun_out/r2/band_sweep.txt 2>&1
done; done
B200FFT_BAND_COLS=64 python tools/quick_bench.py 3 | sed "s/^/cols=64 /" >> gpurun_out/r2/band_sweep.txt 2>&1
cat gpurun_out/r2/band_sweep.txt
