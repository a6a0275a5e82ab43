#!/bin/bash
for v in 0 1; do
  for c in 3 4; do
    echo -n "variant=$v  "
    B200FFT_BAND_VARIANT=$v python tools/quick_bench.py $c | cut -c1-130
  done
done
echo "== variant=0 prof"
B200FFT_BAND_PROF=1 python tools/quick_bench.py 3 2>&1 | tail -2 | cut -c1-400
