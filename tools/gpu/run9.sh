mkdir -p gpurun_out/r2
timeout 90 python tools/gpu/band_check.py > gpurun_out/r2/band_check.txt 2>&1
if [ "$(grep -c OK gpurun_out/r2/band_check.txt)" != "9" ]; then echo "band_check failed: stopping"; tail -5 gpurun_out/r2/band_check.txt; exit 1; fi
rm -f gpurun_out/r2/band_slots.txt
for sl in 20 24 28 48 24 48; do
B200FFT_BAND_SLOTS=$sl timeout 60 python tools/quick_bench.py 3 2>&1 | tail -1 | sed "s/^/slots=$sl /" >> gpurun_out/r2/band_slots.txt 2>&1
done
cat gpurun_out/r2/band_slots.txt
