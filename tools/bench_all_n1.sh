# bench lines of every configuration on one B200 (no ncu): gpurun_out/ev_bench_*.json
mkdir -p gpurun_out
python bench.py > gpurun_out/ev_bench_C3.json 2> gpurun_out/ev_bench_C3.err
for c in C1 C2 C4 C5; do python bench.py --config $c > gpurun_out/ev_bench_$c.json 2> gpurun_out/ev_bench_$c.err; done
for w in 1024 4096; do python bench.py --walkers-per-gpu $w --no-cpu-baseline > gpurun_out/ev_bench_C3_w$w.json 2> gpurun_out/ev_bench_C3_w$w.err; done
for f in gpurun_out/ev_bench_C[1-5].json gpurun_out/ev_bench_C3_w*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1]); e=d.get('e2e') or {}
print(sys.argv[1].split('ev_bench_')[1], round(d['ms_per_step'],5), round(d['value']), 'e2e', round(e.get('value',0)), round(e.get('ms_per_step',0),4))
PY
done
