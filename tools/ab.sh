nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv 2>&1 | head -3
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown --format=csv,noheader,nounits -lms 100 > /tmp/c.csv 2>&1 &
PID=$!
for h in 0 1 0 1; do echo "stream_hint=$h"; B200FFT_STREAM_HINT=$h python tools/quick_bench.py 2 | tail -1; done
kill $PID; head -3 /tmp/c.csv; wc -l /tmp/c.csv
