# Round-2 evidence on ONE B200: GPU tests, bench lines of every configuration, the walkers-per-
# GPU sweep, the launch list and one `ncu --set full` capture per hot kernel.  Everything goes
# to gpurun_out/ev_*; tools/evidence_collect.py turns it into profiles/r02_*.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/ev_pytest.log 2>&1; tail -2 gpurun_out/ev_pytest.log
python bench.py > gpurun_out/ev_bench_C3.json 2> gpurun_out/ev_bench_C3.err
for c in C1 C2 C4 C5; do python bench.py --config $c > gpurun_out/ev_bench_$c.json 2> gpurun_out/ev_bench_$c.err; done
for w in 1024 4096; do python bench.py --walkers-per-gpu $w --no-cpu-baseline > gpurun_out/ev_bench_C3_w$w.json 2> gpurun_out/ev_bench_C3_w$w.err; done
python bench.py --steps-per-graph 4 --steps 200 --warmup 12 --no-cpu-baseline --no-e2e > gpurun_out/ev_bench_C3_spg4.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ev_bench_C3_reference.json 2>/dev/null
NQ="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev_launches_c3.csv python bench.py $NQ > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev_launches_c4.csv python bench.py --config C4 $NQ > /dev/null 2>&1
for k in contract_kernel synchrotron_fused_kernel walker_prep_kernel combine_lnprob_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 8 -c 2 -o gpurun_out/ev_c3_$k python bench.py $NQ > gpurun_out/ev_ncu_c3_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:ssc_inner --launch-skip 4 -c 1 -o gpurun_out/ev_c4_ssc_inner_wt8 python bench.py --config C4 $NQ > gpurun_out/ev_ncu_c4_wt8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ssc_outer|ssc_seed" --launch-skip 8 -c 2 -o gpurun_out/ev_c4_ssc_rest python bench.py --config C4 $NQ > gpurun_out/ev_ncu_c4_rest.log 2>&1
ls -la gpurun_out/ev_* | awk '{print $5, $9}'
for f in gpurun_out/ev_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get('e2e') or {}
    print(sys.argv[1].split('ev_bench_')[1], 'ms/step', round(d.get('ms_per_step',0),5), 'value', round(d.get('value',0)), 'e2e', round(e.get('value',0)), (d.get('clocks') or {}).get('sm_mhz'))
except Exception as ex: print(sys.argv[1], 'ERR', ex)
PY
done
