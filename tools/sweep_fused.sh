python tools/quick_bench.py 3 -d | grep -v "^$"
for mb in 1 2 4; do for la in 2 4 8; do sl=$((la+4)); echo "band=$mb la=$la slots=$sl: $(B200FFT_BAND_MB=$mb B200FFT_FUSED_LA=$la B200FFT_FUSED_SLOTS=$sl python tools/quick_bench.py 3 | grep cfg)"; done; done
python tools/quick_bench.py 4 | grep cfg
python tools/quick_bench.py 5 | grep cfg
