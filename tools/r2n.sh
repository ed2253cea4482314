mkdir -p gpurun_out
for o in 0 1 2; do echo "=== order $o"; NB_PROGRAM_ORDER=$o python tools/program_trace.py C3 128 2>&1 | tail -12; done
echo "=== order 0 flushed"; python tools/program_trace.py C3 128 flush 2>&1 | tail -12
B="python bench.py --steps 200 --warmup 12 --no-e2e --no-cpu-baseline"
for c in C2 C5; do
  timeout 300 $B --config $c > gpurun_out/r2n_${c}_new.log 2>&1
  NB_ONE_LAUNCH=0 timeout 300 $B --config $c > gpurun_out/r2n_${c}_old.log 2>&1
done
for f in gpurun_out/r2n_C*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d['ms_per_step'],5), d.get('step_ms_min_median_max'), d.get('gpu_launches'), d.get('acceptance_fraction'))
except Exception as e: print('ERR', open(sys.argv[1]).read()[-1500:])
PY
done
