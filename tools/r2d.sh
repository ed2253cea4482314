mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2d_pytest.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_C3_20.log 2>&1
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2d_bench_C3_200.log 2>&1
python bench.py --config C4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_C4_wt16.log 2>&1
NB_SSC_WT=8 python bench.py --config C4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_C4_wt8.log 2>&1
python bench.py --config C5 --steps 50 --warmup 5 > gpurun_out/r2d_bench_C5.log 2>&1
tail -5 gpurun_out/r2d_pytest.log
