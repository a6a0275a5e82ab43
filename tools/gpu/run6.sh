mkdir -p gpurun_out/r2
rm -f gpurun_out/r2/band_dbg.txt
for dbg in 1 2 3; do
B200FFT_BAND_DEBUG=$dbg B200FFT_BAND_SLOTS=8 timeout 100 python tools/quick_bench.py 3 | sed "s/^/debug=$dbg /" >> gpurun_out/r2/band_dbg.txt 2>&1
done
cat gpurun_out/r2/band_dbg.txt
