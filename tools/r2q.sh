mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547"
B="bench.py --steps 200 --warmup 12 --no-e2e --no-cpu-baseline"
for c in C2 C3 C5; do
  timeout 300 python $B --config $c > gpurun_out/r2q_${c}_new.log 2>&1
  NB_ONE_LAUNCH=0 timeout 300 python $B --config $c > gpurun_out/r2q_${c}_old.log 2>&1
done
timeout 300 $TR $B --gpus 2 > gpurun_out/r2q_C3n2_new.log 2>&1
NB_ONE_LAUNCH=0 timeout 300 $TR $B --gpus 2 > gpurun_out/r2q_C3n2_old.log 2>&1
timeout 300 $TR tests/multi/check_sharded.py > gpurun_out/r2q_check.log 2>&1; tail -4 gpurun_out/r2q_check.log
for f in gpurun_out/r2q_C*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d['ms_per_step'],5), d.get('step_ms_min_median_max'), d.get('gpu_launches'), d.get('acceptance_fraction'), d.get('sharded_chain_bitwise'))
except Exception as e: print('ERR', open(sys.argv[1]).read()[-1500:])
PY
done
