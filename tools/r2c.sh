mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r2c_pytest.log 2>&1
python tools/debug_r2.py c5step > gpurun_out/r2c_c5step.log 2>&1
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2c_bench_C3_pdl.log 2>&1
NB_PDL=0 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2c_bench_C3_nopdl.log 2>&1
python bench.py --config C4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_C4.log 2>&1
python bench.py --config C2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_C2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ssc_inner|contract_kernel" --launch-skip 8 -c 2 -o gpurun_out/r2c_c4 python bench.py --config C4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c_ncu_c4.log 2>&1
tail -5 gpurun_out/r2c_pytest.log; tail -30 gpurun_out/r2c_c5step.log
