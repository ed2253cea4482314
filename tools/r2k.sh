mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log
python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline --no-flush --timeline gpurun_out/r2k_tl1_noflush_ > gpurun_out/r2k_n1_tl_nf.log 2>&1
python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline --timeline gpurun_out/r2k_tl1_flush_ > gpurun_out/r2k_n1_tl.log 2>&1
timeout 200 $TR bench.py --gpus 2 --steps 200 --warmup 10 --no-e2e --no-flush --timeline gpurun_out/r2k_tl2_noflush_ > gpurun_out/r2k_n2_tl_nf.log 2>&1
timeout 200 $TR bench.py --gpus 2 --steps 200 --warmup 10 --no-e2e --timeline gpurun_out/r2k_tl2_flush_ > gpurun_out/r2k_n2_tl.log 2>&1
for f in gpurun_out/r2k_n*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d['ms_per_step'],5), d.get('step_ms_min_median_max'), d.get('sharded_chain_bitwise'))
except Exception as e: print('ERR', open(sys.argv[1]).read()[-600:])
PY
done
python tools/timeline_report.py gpurun_out/r2k_tl1_noflush_ 1 300
python tools/timeline_report.py gpurun_out/r2k_tl2_noflush_ 2 300
python tools/timeline_report.py gpurun_out/r2k_tl1_flush_ 1 300
python tools/timeline_report.py gpurun_out/r2k_tl2_flush_ 2 300
