mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2g_pytest.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_C3_20.log 2>&1
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2g_bench_C3_200.log 2>&1
python bench.py --config C4 --steps 10 --warmup 3 > gpurun_out/r2g_bench_C4.log 2>&1
python bench.py --config C5 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench_C5.log 2>&1
python bench.py --config C1 --steps 50 --warmup 5 > gpurun_out/r2g_bench_C1.log 2>&1
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --walkers-per-gpu 1024 > gpurun_out/r2g_bench_C3_w1024.log 2>&1
python bench.py --steps 50 --warmup 10 --no-cpu-baseline --walkers-per-gpu 4096 > gpurun_out/r2g_bench_C3_w4096.log 2>&1
tail -5 gpurun_out/r2g_pytest.log
