set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2/pytest_gpu.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench_n1.json 2> gpurun_out/r2/bench_n1.err; tail -c 600 gpurun_out/r2/bench_n1.err
python bench.py --steps 20 --warmup 3 --config cfg1 --no-configs --no-cpu > gpurun_out/r2/bench_cfg1.json 2>> gpurun_out/r2/bench_n1.err
python bench.py --steps 20 --warmup 3 --config cfg1 --no-configs --no-cpu --no-graph > gpurun_out/r2/bench_cfg1_nograph.json 2>> gpurun_out/r2/bench_n1.err
