#!/bin/bash
# band kernel stream isolation on cfg3 (rows pass ~200 us in every line): debug = mode + 8 nomath + 16 nostore + 32 noload
for d in 0 2 3 10 11 26 27 42 43 58 59; do
  echo -n "debug=$d  "; B200FFT_BAND_DEBUG=$d python tools/quick_bench.py 3 | cut -c1-60
done
