mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2m_pytest.log
B="python bench.py --steps 200 --warmup 12 --no-e2e --no-cpu-baseline"
for o in 0 1 2; do
  NB_PROGRAM_ORDER=$o timeout 300 $B --no-flush --timeline gpurun_out/r2m_o${o}_ > gpurun_out/r2m_o${o}_nf.log 2>&1
  NB_PROGRAM_ORDER=$o timeout 300 $B > gpurun_out/r2m_o${o}.log 2>&1
done
NB_ONE_LAUNCH=0 timeout 300 $B > gpurun_out/r2m_old.log 2>&1
for f in gpurun_out/r2m_o*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d['ms_per_step'],5), d.get('step_ms_min_median_max'), d.get('gpu_launches'), d.get('acceptance_fraction'))
except Exception as e: print('ERR', open(sys.argv[1]).read()[-1500:])
PY
done
for o in 0 1 2; do echo "== order $o"; python tools/timeline_report.py gpurun_out/r2m_o${o}_ 1 300; done
