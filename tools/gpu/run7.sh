mkdir -p gpurun_out/r2
for dbg in 2 3; do
B200FFT_BAND_DEBUG=$dbg B200FFT_BAND_SLOTS=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:band -c 1 -f -o gpurun_out/r2/band_dbg$dbg python tools/ncu_one.py 2d f 1 8192 8192 > gpurun_out/r2/ncu_band_dbg$dbg.log 2>&1
done
