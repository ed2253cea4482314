mkdir -p gpurun_out
for rt in 8 4; do echo "=== RT cap $rt, order 2"; NB_PROGRAM_RT=$rt NB_PROGRAM_ORDER=2 python tools/program_trace.py C3 128 2>&1 | tail -9; done
B="python bench.py --steps 200 --warmup 12 --no-e2e --no-cpu-baseline"
for rt in 8 4; do for o in 0 2; do
  NB_PROGRAM_RT=$rt NB_PROGRAM_ORDER=$o timeout 300 $B --no-flush > gpurun_out/r2o_rt${rt}_o${o}_nf.log 2>&1
  NB_PROGRAM_RT=$rt NB_PROGRAM_ORDER=$o timeout 300 $B > gpurun_out/r2o_rt${rt}_o${o}.log 2>&1
done; done
for f in gpurun_out/r2o_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d['ms_per_step'],5), d.get('step_ms_min_median_max'), d.get('gpu_launches'), d.get('acceptance_fraction'))
except Exception as e: print('ERR', open(sys.argv[1]).read()[-1500:])
PY
done
