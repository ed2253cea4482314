mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2b_pytest.log 2>&1
python tools/debug_r2.py c5nan > gpurun_out/r2b_c5nan.log 2>&1
python tools/debug_r2.py profile C3 > gpurun_out/r2b_profile.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_C3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"contract_kernel|synchrotron_fused|walker_prep|combine_lnprob" --launch-skip 60 -c 8 -o gpurun_out/r2b_c3 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b_ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ssc_inner|contract_kernel|ssc_outer" --launch-skip 6 -c 3 -o gpurun_out/r2b_c4 python bench.py --config C4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b_ncu_c4.log 2>&1
tail -5 gpurun_out/r2b_pytest.log; tail -30 gpurun_out/r2b_c5nan.log; head -40 gpurun_out/r2b_profile.log
