mkdir -p gpurun_out
export NB_PROGRAM_RT=4
for cfg in "0 8" "3 8" "2 8" "0 16" "3 16"; do set -- $cfg; echo "=== order $1 wpc $2"; NB_PROGRAM_WPC=$2 NB_PROGRAM_ORDER=$1 python tools/program_trace.py C3 128 2>&1 | tail -9; done
B="python bench.py --steps 200 --warmup 12 --no-e2e --no-cpu-baseline"
for cfg in "0 8" "3 8" "2 8"; do set -- $cfg
  NB_PROGRAM_WPC=$2 NB_PROGRAM_ORDER=$1 timeout 300 $B --no-flush > gpurun_out/r2p_o$1_w$2_nf.log 2>&1
  NB_PROGRAM_WPC=$2 NB_PROGRAM_ORDER=$1 timeout 300 $B > gpurun_out/r2p_o$1_w$2.log 2>&1
done
for f in gpurun_out/r2p_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d['ms_per_step'],5), d.get('step_ms_min_median_max'), d.get('gpu_launches'), d.get('acceptance_fraction'))
except Exception as e: print('ERR', open(sys.argv[1]).read()[-1500:])
PY
done
