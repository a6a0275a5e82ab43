mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.txt 2>&1; tail -4 gpurun_out/r2/pytest_gpu.txt
cat > /tmp/q4096.py <<'PY'
import sys; sys.path.insert(0,"."); sys.argv=["x","none"]
exec(open("tools/quick_bench.py").read().split("which = sys.argv")[0])
bench("c64 4096x4096", "2d", [4096,4096], af.C2C, 1, 2, 20)
bench("c64 4096x16384", "2d", [4096,16384], af.C2C, 1, 2, 10)
bench("c64 8192x4096", "2d", [8192,4096], af.C2C, 1, 2, 10)
bench("c64 8192x8192", "2d", [8192,8192], af.C2C, 1, 2, 10)
PY
timeout 100 python /tmp/q4096.py > gpurun_out/r2/q4096.txt 2>&1
B200FFT_BAND=0 timeout 100 python /tmp/q4096.py 2>&1 | sed "s/^/BAND=0 /" >> gpurun_out/r2/q4096.txt
cat gpurun_out/r2/q4096.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench_n1.json 2> gpurun_out/r2/bench_n1.err; tail -c 300 gpurun_out/r2/bench_n1.err
