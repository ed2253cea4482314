# Last check of the round on one B200: GPU tests, smoke(), the default bench line, C4.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/ev_pytest.log 2>&1; tail -2 gpurun_out/ev_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/ev_bench_C3.json 2> gpurun_out/ev_bench_C3.err
python bench.py --config C4 > gpurun_out/ev_bench_C4.json 2> gpurun_out/ev_bench_C4.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ev_bench_C3_reference.json 2>/dev/null
for f in gpurun_out/ev_bench_C3.json gpurun_out/ev_bench_C4.json gpurun_out/ev_bench_C3_reference.json; do python - "$f" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1]); e=d.get('e2e') or {}
print(sys.argv[1], round(d['ms_per_step'],5), round(d['value']), 'e2e', round(e.get('value',0)), d.get('gpu_launches'), d.get('clocks'))
PY
done
