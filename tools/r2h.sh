mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2h_pytest.log 2>&1
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2h_bench_C3_self.log 2>&1
NB_SELF_TABLE=0 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2h_bench_C3_noself.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench_C3_20.log 2>&1
python bench.py --config C2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench_C2.log 2>&1
python bench.py --config C4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_C4.log 2>&1
python tools/debug_r2.py profile C3 > gpurun_out/r2h_profile.log 2>&1
tail -5 gpurun_out/r2h_pytest.log; head -12 gpurun_out/r2h_profile.log
