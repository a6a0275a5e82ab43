#!/bin/bash
# band kernel loader / group cycle accounting on cfg3 (quick_bench prints the stderr lines of the last execs)
mkdir -p gpurun_out/r2
for d in 0 2 3 10 11; do
  echo "== debug=$d"
  B200FFT_BAND_PROF=1 B200FFT_BAND_DEBUG=$d python tools/quick_bench.py 3 2>&1 | tail -3 | cut -c1-400
done
