mkdir -p gpurun_out/r2
timeout 60 python tools/gpu/band_check.py > gpurun_out/r2/band_check.txt 2>&1
if [ "$(grep -c OK gpurun_out/r2/band_check.txt)" != "9" ]; then echo "band_check failed: stopping"; tail -5 gpurun_out/r2/band_check.txt; exit 1; fi
rm -f gpurun_out/r2/band_var.txt
for v in 0 1 2 3; do
for dbg in 1 0; do
B200FFT_BAND_VARIANT=$v B200FFT_BAND_DEBUG=$dbg B200FFT_BAND_SLOTS=16 timeout 40 python tools/quick_bench.py 3 2>&1 | tail -1 | sed "s/^/variant=$v debug=$dbg /" >> gpurun_out/r2/band_var.txt 2>&1
done; done
cat gpurun_out/r2/band_var.txt
