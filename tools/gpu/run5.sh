mkdir -p gpurun_out/r2
timeout 60 python tools/gpu/band_check.py > gpurun_out/r2/band_check.txt 2>&1; grep -c OK gpurun_out/r2/band_check.txt; grep FAIL gpurun_out/r2/band_check.txt; tail -3 gpurun_out/r2/band_check.txt
if [ "$(grep -c OK gpurun_out/r2/band_check.txt)" != "9" ]; then echo "band_check failed: stopping"; exit 1; fi
rm -f gpurun_out/r2/band_sweep.txt
for dbg in 1 2 3; do
B200FFT_BAND_DEBUG=$dbg B200FFT_BAND_SLOTS=8 timeout 40 python tools/quick_bench.py 3 2>&1 | tail -1 | sed "s/^/debug=$dbg /" >> gpurun_out/r2/band_sweep.txt 2>&1
done
for sl in 6 8 12 16; do
B200FFT_BAND_SLOTS=$sl timeout 40 python tools/quick_bench.py 3 2>&1 | tail -1 | sed "s/^/slots=$sl /" >> gpurun_out/r2/band_sweep.txt 2>&1
done
B200FFT_BAND_COLS=64 B200FFT_BAND_SLOTS=6 timeout 40 python tools/quick_bench.py 3 2>&1 | tail -1 | sed "s/^/cols=64 slots=6 /" >> gpurun_out/r2/band_sweep.txt 2>&1
cat gpurun_out/r2/band_sweep.txt
B200FFT_BAND_SLOTS=12 timeout 120 ncu --set full --clock-control none --import-source on -k regex:band -c 1 -f -o gpurun_out/r2/band3 python tools/ncu_one.py 2d f 1 8192 8192 > gpurun_out/r2/ncu_band3.log 2>&1
